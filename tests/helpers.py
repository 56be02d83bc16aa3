"""Shared test helpers: seeded states, oracle/CUDA twins, comparison utilities."""
from __future__ import annotations

import numpy as np

from artemis_b200.enums import (BoundaryFlag, Coordinates, Fluid, ReconstructionMethod,
                                RSolver)
from artemis_b200.mesh import UniformMesh
from artemis_b200.params import FluidParams

GEOM_DOMAINS = {
    # (xmin, xmax) per coordinate system chosen away from coordinate singularities
    Coordinates.cartesian: ((0.0, 0.0, 0.0), (1.0, 0.8, 0.6)),
    Coordinates.cylindrical: ((0.5, 0.0, -0.4), (2.0, 1.2, 0.4)),
    Coordinates.axisymmetric: ((0.5, -0.4, 0.0), (2.0, 0.4, 1.0)),
    Coordinates.spherical3D: ((0.6, 0.7, 0.0), (2.2, 2.3, 1.1)),
    Coordinates.spherical2D: ((0.6, 0.7, 0.0), (2.2, 2.3, 1.0)),
    Coordinates.spherical1D: ((0.6, 0.0, 0.0), (2.2, 1.0, 1.0)),
}


def make_mesh(coords=Coordinates.cartesian, ndim=3, nblk=(2, 2, 2), bnx=(8, 6, 4), ng=4,
              bcs=None):
    if coords == Coordinates.spherical2D:
        ndim = 2
    if coords == Coordinates.spherical1D:
        ndim = 1
    if coords == Coordinates.spherical3D:
        ndim = 3
    nblk = tuple(nblk[d] if d < ndim else 1 for d in range(3))
    bnx = tuple(bnx[d] if d < ndim else 1 for d in range(3))
    nx = tuple(nblk[d] * bnx[d] for d in range(3))
    xmin, xmax = GEOM_DOMAINS[coords]
    if bcs is None:
        bcs = (BoundaryFlag.periodic,) * 6
    return UniformMesh(nx=nx, xmin=xmin, xmax=xmax, block_nx=bnx, nghost=ng, bcs=tuple(bcs),
                       coords=coords)


def gas_params(coords, recon="ppm", rs="hllc", S=1, gamma=1.4, cfl=0.3, de_switch=0.0):
    return FluidParams(Fluid.gas, coords, ReconstructionMethod[recon], RSolver[rs], cfl=cfl,
                       nspecies=S, dfloor=1e-10, gamma=gamma, siefloor=1e-10,
                       de_switch=de_switch)


def dust_params(coords, recon="plm", rs="hlle", S=2, cfl=0.3):
    return FluidParams(Fluid.dust, coords, ReconstructionMethod[recon], RSolver[rs], cfl=cfl,
                       nspecies=S, dfloor=1e-10)


def random_prim(mesh, fp, seed=0, smooth=True, shocks=True):
    """Seeded, physically admissible primitives over the ENTIRE domain (ghosts included).

    A smooth multi-mode field plus a few sharp jumps so limiter and wave-speed branches are
    exercised (SURVEY 8d config 3 note)."""
    rng = np.random.default_rng(seed)
    S = fp.nspecies
    gas = fp.fluid_type == Fluid.gas
    shp = mesh.shape(fp.nvar)
    prim = np.zeros(shp)
    nb, _, nk, nj, ni = shp
    kk, jj, ii = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
    for b in range(nb):
        base = rng.uniform(0, 2 * np.pi, size=6)
        wave = (np.sin(0.9 * ii + base[0]) * np.cos(0.7 * jj + base[1]) *
                np.cos(0.5 * kk + base[2]))
        wave2 = np.cos(0.6 * ii + base[3]) * np.sin(0.8 * jj + base[4] + 0.4 * kk)
        noise = rng.normal(size=(nk, nj, ni))
        step = (ii > ni // 2).astype(float) * (jj > nj // 3) if shocks else 0.0
        for n in range(S):
            prim[b, n] = 1.0 + 0.3 * wave + 0.05 * noise + 0.8 * step + 0.1 * n
            for d in range(3):
                prim[b, S + 3 * n + d] = (0.4 * wave2 * (d + 1) / 2 + 0.1 * noise * (d - 1)
                                          - 0.5 * step * (d == 0) + 0.05 * n)
            if gas:
                prim[b, 5 * S + n] = 1.5 + 0.5 * wave2 + 0.05 * noise + 1.2 * step
                prim[b, 4 * S + n] = fp.gm1 * prim[b, n] * prim[b, 5 * S + n]
    return prim


def rel_err(a, b, floor=None):
    """Element-wise relative difference with an absolute floor.  Used only where no per-zone
    physical scale exists (flux arrays, refinement buffers, strided golden samples); state
    comparisons use `zone_rel_err`."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.maximum(np.abs(a), np.abs(b))
    if floor is None:
        floor = max(np.max(scale) * 1e-3, 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(scale, floor)))


def zone_rel_err(got, want, fp, kind, vref=None, per_var=False):
    """north_star's parity metric: the PER-ZONE relative difference of a state array
    [nb][nvar][nk][nj][ni] (pack order IDN=n, IVX=S+3n+d, IPR|IEN=4S+n, ISE|IU=5S+n) against
    the reference values `want`.  No global-maximum floor:

      * positive scalars (density; pressure / total energy; sie / internal energy): the
        difference divided by the zone's own reference value (floors: dfloor, siefloor*dfloor);
      * vector components (velocity; momentum): the difference divided by the zone's own
        vector magnitude, floored by the zone's own sound speed (gas: c_s = sqrt(gamma gm1 sie)
        of that zone, times the zone's density for momenta).  A component that vanishes by
        symmetry has no relative error of its own; the zone's characteristic speed is the
        per-zone scale of its velocity vector.  Pressureless dust has no sound speed: its floor
        is `vref` (per-zone array or scalar, e.g. the co-located gas sound speed) or, failing
        that, 1e-6 of the largest dust speed on the mesh.
    kind = "cons" | "prim".  Curvilinear momenta carry the scale factor h_d (O(1) on the test
    domains); it is ignored in the floor only."""
    from artemis_b200.enums import Fluid
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape and kind in ("cons", "prim")
    S = fp.nspecies
    gas = fp.fluid_type == Fluid.gas
    worst = []
    tiny = 1e-300
    for n in range(S):
        rho = np.maximum(np.abs(want[:, n]), fp.dfloor)
        vec = want[:, S + 3 * n:S + 3 * n + 3]
        mag = np.sqrt(np.sum(vec * vec, axis=1))
        if gas:
            ie = np.abs(want[:, 5 * S + n])                       # sie (prim) or u (cons)
            sie = ie / rho if kind == "cons" else ie
            cs = np.sqrt((fp.gm1 + 1.0) * fp.gm1 * np.maximum(sie, fp.siefloor))
            vfloor = cs * rho if kind == "cons" else cs
        else:
            if vref is None:
                vf = 1e-6 * float(np.max(mag / rho if kind == "cons" else mag))
            else:
                vf = vref
            vfloor = vf * rho if kind == "cons" else vf * np.ones_like(rho)
        vscale = np.maximum(np.maximum(mag, vfloor), tiny)
        e = [np.max(np.abs(got[:, n] - want[:, n]) / rho)]
        for d in range(3):
            e.append(np.max(np.abs(got[:, S + 3 * n + d] - vec[:, d]) / vscale))
        if gas:
            efl = fp.siefloor * fp.dfloor if kind == "cons" else fp.siefloor
            for v in (4 * S + n, 5 * S + n):
                sc = np.maximum(np.abs(want[:, v]), max(efl, tiny))
                e.append(np.max(np.abs(got[:, v] - want[:, v]) / sc))
        worst.append([float(x) for x in e])
    if per_var:
        return worst
    return max(max(w) for w in worst)


def interior(mesh, a):
    sl = mesh.interior()
    return a[(slice(None), slice(None)) + sl]
