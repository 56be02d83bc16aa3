"""Pins the CPU oracle to numbers produced by the REAL reference (lanl/artemis @ 6c2a7a8,
Kokkos-OpenMP build, measured during the survey; BASELINE.md / SURVEY.md 8c) and to the
reference's own regression thresholds (tst/scripts/hydro/linwave.py:96-143)."""
import json
import os

import numpy as np
import pytest

from artemis_b200 import pgen
from artemis_b200.enums import Coordinates, Fluid, ReconstructionMethod, RSolver
from artemis_b200.mesh import UniformMesh
from artemis_b200.params import FluidParams
from oracle.oracle_py import OracleSim

GOLDEN_FILE = os.path.join(os.path.dirname(__file__), "golden", "reference_linwave.json")
with open(GOLDEN_FILE) as fh:
    GOLDEN = json.load(fh)


def _run(res, recon, rs="hllc", wave_flag=0, amp=1e-6, ng=4, nx3=None, cfl=0.9, vflow=0.0):
    if nx3 == 1:   # the deck-default 2D case: 128x64, 32^2 blocks, domain 3.0 x 1.5
        mesh = UniformMesh(nx=(res, res // 2, 1), xmin=(0, 0, 0), xmax=(3.0, 1.5, 1.5),
                           block_nx=(32, 32, 1), nghost=ng)
    else:          # the regression-test geometry 3.0 x 1.5 x 1.5
        mesh = UniformMesh(nx=(res, res // 2, res // 2), xmin=(0, 0, 0), xmax=(3.0, 1.5, 1.5),
                           block_nx=(res // 4,) * 3, nghost=ng)
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod[recon], RSolver[rs],
                     cfl=cfl, nspecies=1, dfloor=1e-20, gamma=1.66666666667)
    prim, lw = pgen.linear_wave(mesh, gp.gamma, wave_flag, amp, vflow)
    sim = OracleSim(mesh, gas=gp)
    sim.gas.prim[:] = prim
    sim.tlim, sim.nlim = lw.tlim, 100000
    sim.initialize()
    sim.run()
    rms, l1 = pgen.linear_wave_errors(mesh, lw, sim.gas.u0)
    return rms, l1, sim.ncycle


@pytest.mark.parametrize("recon", ["plm", "ppm"])
def test_rms_l1_matches_the_reference_to_all_published_digits(recon, oracle_lib):
    g = GOLDEN["test_geometry"][recon]
    errs = {}
    for res in (16, 32):
        rms, l1, ncycle = _run(res, recon)
        assert ncycle == g[str(res)]["ncycles"]
        assert f"{rms:.6e}" == f"{g[str(res)]['rms_l1']:.6e}"
        errs[res] = rms
    ratio = errs[32] / errs[16]
    assert abs(ratio - g["ratio"]) < 1e-4
    thr = GOLDEN["reference_test_thresholds"][recon]
    assert errs[32] <= thr["err_max_sound"] and ratio <= thr["ratio_max_sound"]


def test_per_variable_l1_matches_the_reference(oracle_lib):
    rms, l1, _ = _run(32, "plm")
    want = GOLDEN["test_geometry"]["plm"]["32"]["l1_per_variable"]
    for got, w in zip(l1, want):
        assert f"{got:.6e}" == f"{w:.6e}"


def test_left_and_right_going_sound_waves_have_identical_errors(oracle_lib):
    """tst/scripts/hydro/linwave.py:135-143."""
    for recon in ("plm", "ppm"):
        l = _run(16, recon, wave_flag=0)[0]
        r = _run(16, recon, wave_flag=4)[0]
        assert l == r


def test_deck_default_2d_entropy_wave(oracle_lib):
    """inputs/linwave/linear_wave.in + nx3=1: 128x64, nghost=2, PLM+HLLC, entropy wave, amp
    1e-3 -> 213 cycles, RMS-L1 1.235051e-05 in the reference."""
    g = GOLDEN["deck_default_2d"]
    rms, _, ncycle = _run(128, "plm", wave_flag=g["wave_flag"], amp=g["amp"], ng=2, nx3=1,
                         vflow=g["vflow"])
    assert ncycle == g["ncycles"]
    assert f"{rms:.6e}" == f"{g['rms_l1']:.6e}"


@pytest.mark.parametrize("rs", ["hlle", "llf"])
@pytest.mark.parametrize("recon", ["plm", "ppm"])
def test_other_solvers_meet_the_reference_thresholds(recon, rs, oracle_lib):
    """The reference's regression runs every {plm,ppm} x {hllc,hlle,llf} combination against
    one threshold pair (linwave.py:96-106)."""
    e16 = _run(16, recon, rs)[0]
    e32 = _run(32, recon, rs)[0]
    thr = GOLDEN["reference_test_thresholds"][recon]
    assert e32 <= thr["err_max_sound"] and e32 / e16 <= thr["ratio_max_sound"]
