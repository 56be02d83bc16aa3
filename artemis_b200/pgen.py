"""Problem generators (host side, numpy): they FEED the hot path, they are not on it.

Restated from the reference so parity runs start from identical primitives:
  * linear_wave -- src/pgen/linear_wave.hpp:62-258 (ICs) and :266-330 (L1 errors)
  * blast       -- src/pgen/blast.hpp:138-230
  * drag-like constant state + seeded perturbation
Primitive layout: [nb][6S][nk][nj][ni] gas (rho | v | P | sie), [nb][4S][...] dust.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .enums import Coordinates
from .mesh import UniformMesh


def cell_centers(mesh: UniformMesh, b: int):
    """Cartesian cell centroids 0.5*(xf[i]+xf[i+1]) (geometry.hpp:166-168)."""
    out = []
    for d in range(3):
        xf = mesh.face_positions(b, d)
        out.append(0.5 * (xf[:-1] + xf[1:]))
    return out  # x1v[ni], x2v[nj], x3v[nk]


def cell_volumes(mesh: UniformMesh, b: int):
    xf = [mesh.face_positions(b, d) for d in range(3)]
    d1, d2, d3 = (x[1:] - x[:-1] for x in xf)
    return (d1[None, None, :] * d2[None, :, None]) * d3[:, None, None]


@dataclass
class LinWave:
    """LinWaveVariables (linear_wave.hpp:45-53) + the analytic solution."""
    wave_flag: int
    amp: float
    vflow: float
    gamma: float
    lam: float
    k_par: float
    cos_a2: float
    cos_a3: float
    sin_a2: float
    sin_a3: float
    rem: np.ndarray
    ev: np.ndarray
    d0: float = 1.0
    tlim: float = 0.0

    @property
    def p0(self):
        return 1.0 / self.gamma

    def conserved(self, x1v, x2v, x3v):
        """(cd, cm1, cm2, cm3, ce) on a broadcast grid; linear_wave.hpp:226-246."""
        wf = self.wave_flag
        x = self.cos_a2 * (x1v * self.cos_a3 + x2v * self.sin_a3) + x3v * self.sin_a2
        sn = np.sin(self.k_par * x)
        mx = self.d0 * self.vflow + self.amp * sn * self.rem[1][wf]
        my = self.amp * sn * self.rem[2][wf]
        mz = self.amp * sn * self.rem[3][wf]
        cd = self.d0 + self.amp * sn * self.rem[0][wf]
        cm1 = mx * self.cos_a2 * self.cos_a3 - my * self.sin_a3 - mz * self.sin_a2 * self.cos_a3
        cm2 = mx * self.cos_a2 * self.sin_a3 + my * self.cos_a3 - mz * self.sin_a2 * self.sin_a3
        cm3 = mx * self.sin_a2 + mz * self.cos_a2
        gm1 = self.gamma - 1.0
        ce = (self.p0 / gm1 + 0.5 * self.d0 * self.vflow * self.vflow
              + self.amp * sn * self.rem[4][wf])
        return cd, cm1, cm2, cm3, ce


def _hydro_eigensystem(d, v1, v2, v3, p, gamma):
    """linear_wave.hpp:62-111."""
    vsq = v1 * v1 + v2 * v2 + v3 * v3
    h = (p / (gamma - 1.0) + 0.5 * d * vsq + p) / d
    a = math.sqrt(gamma * p / d)
    ev = np.array([v1 - a, v1, v1, v1, v1 + a])
    rem = np.zeros((5, 5))
    rem[:, 0] = [1.0, v1 - a, v2, v3, h - v1 * a]
    rem[:, 1] = [0.0, 0.0, 1.0, 0.0, v2]
    rem[:, 2] = [0.0, 0.0, 0.0, 1.0, v3]
    rem[:, 3] = [1.0, v1, v2, v3, 0.5 * vsq]
    rem[:, 4] = [1.0, v1 + a, v2, v3, h + v1 * a]
    return ev, rem


def linear_wave(mesh: UniformMesh, gamma, wave_flag, amp, vflow=0.0, nperiod=1.0,
                along_x1=False, along_x2=False, along_x3=False):
    """Returns (prim[nb,6,nk,nj,ni], LinWave).  linear_wave.hpp:117-258."""
    if mesh.coords != Coordinates.cartesian:
        raise ValueError("linear_wave pgen requires Cartesian geometry!")
    ndim = mesh.ndim
    x1s, x2s, x3s = (mesh.xmax[d] - mesh.xmin[d] for d in range(3))
    cos_a3, sin_a3, cos_a2, sin_a2 = 1.0, 0.0, 1.0, 0.0
    if ndim > 1 and not along_x1:
        ang3 = math.atan(x1s / x2s)
        sin_a3, cos_a3 = math.sin(ang3), math.cos(ang3)
    if ndim > 2 and not along_x1:
        ang2 = math.atan(0.5 * (x1s * cos_a3 + x2s * sin_a3) / x3s)
        sin_a2, cos_a2 = math.sin(ang2), math.cos(ang2)
    if along_x2:
        cos_a3, sin_a3, cos_a2, sin_a2 = 0.0, 1.0, 1.0, 0.0
    if along_x3:
        cos_a3, sin_a3, cos_a2, sin_a2 = 0.0, 1.0, 0.0, 1.0
    lam = float(np.finfo(np.float32).max)
    if cos_a2 * cos_a3 > 0.0:
        lam = min(lam, x1s * cos_a2 * cos_a3)
    if cos_a2 * sin_a3 > 0.0:
        lam = min(lam, x2s * cos_a2 * sin_a3)
    if sin_a2 > 0.0:
        lam = min(lam, x3s * sin_a2)
    k_par = 2.0 * math.pi / lam
    d0, p0 = 1.0, 1.0 / gamma
    ev, rem = _hydro_eigensystem(d0, vflow, 0.0, 0.0, p0, gamma)
    lw = LinWave(wave_flag, amp, vflow, gamma, lam, k_par, cos_a2, cos_a3, sin_a2, sin_a3,
                 rem, ev, d0, tlim=nperiod * abs(lam / ev[wave_flag]))
    prim = np.zeros(mesh.shape(6))
    for b in range(mesh.nb):
        x1v, x2v, x3v = cell_centers(mesh, b)
        cd, cm1, cm2, cm3, ce = lw.conserved(x1v[None, None, :], x2v[None, :, None],
                                             x3v[:, None, None])
        cu = ce - 0.5 * (cm1 * cm1 + cm2 * cm2 + cm3 * cm3) / cd
        prim[b, 0] = cd
        prim[b, 1] = cm1 / cd
        prim[b, 2] = cm2 / cd
        prim[b, 3] = cm3 / cd
        prim[b, 5] = cu / cd
    return prim, lw


def linear_wave_errors(mesh: UniformMesh, lw: LinWave, cons: np.ndarray):
    """UserWorkAfterLoop L1 errors (linear_wave.hpp:266-330): returns (rms, l1[5])."""
    l1 = np.zeros(5)
    sl = mesh.interior()
    for b in range(mesh.nb):
        x1v, x2v, x3v = cell_centers(mesh, b)
        ref = lw.conserved(x1v[None, None, :], x2v[None, :, None], x3v[:, None, None])
        vol = cell_volumes(mesh, b)[sl]
        for n, comp in enumerate((0, 1, 2, 3, 4)):
            r = np.broadcast_to(ref[n], (mesh.nk, mesh.nj, mesh.ni))[sl]
            l1[n] += np.sum(vol * np.abs(cons[b, comp][sl] - r))
    tot = np.prod([mesh.xmax[d] - mesh.xmin[d] for d in range(3)])
    l1 /= tot
    return float(np.sqrt(np.sum(l1 * l1))), l1


def blast(mesh: UniformMesh, gamma, d0=1.0, p0=1.0, internal_energy=1.0, radius=1.0,
          x0=(0.0, 0.0, 0.0), symmetry="spherical", samples=-1):
    """blast.hpp:138-230, Cartesian; returns prim[nb,6,nk,nj,ni]."""
    if mesh.coords != Coordinates.cartesian:
        raise ValueError("blast pgen: only the Cartesian branch is restated")
    gm1 = gamma - 1.0
    e0 = p0 / gm1
    prim = np.zeros(mesh.shape(6))
    for b in range(mesh.nb):
        xf = [mesh.face_positions(b, d) for d in range(3)]
        x1v, x2v, x3v = cell_centers(mesh, b)
        tv = cell_volumes(mesh, b)
        xc = x1v[None, None, :] - x0[0]
        yc = x2v[None, :, None] - x0[1]
        zc = x3v[:, None, None] - x0[2]
        if samples > 0:
            s = (np.arange(samples) + 0.5)
            ov = np.zeros_like(tv)
            d1 = (xf[0][1:] - xf[0][:-1]) / samples
            d2 = (xf[1][1:] - xf[1][:-1]) / samples
            d3 = (xf[2][1:] - xf[2][:-1]) / samples
            xs = xf[0][:-1, None] + s[None, :] * d1[:, None]      # [ni, samples]
            ys = xf[1][:-1, None] + s[None, :] * d2[:, None]
            zs = xf[2][:-1, None] + s[None, :] * d3[:, None]
            if symmetry == "spherical":
                for k in range(mesh.nk):
                    r2 = (xs[None, :, None, :, None] ** 2 + ys[:, None, :, None, None] ** 2
                          + zs[k][None, None, None, None, :] ** 2)
                    cnt = np.sum(r2 <= radius * radius, axis=(2, 3, 4))
                    ov[k] = cnt * d1[None, :] * d2[:, None] * d3[k]
            else:
                r2 = xs[None, :, None, :] ** 2 + ys[:, None, :, None] ** 2
                cnt = np.sum(r2 <= radius * radius, axis=(2, 3))
                ov[:] = (cnt * d1[None, :] * d2[:, None])[None]
            vol = ov
        else:
            vol = np.where(xc * xc + yc * yc + zc * zc < radius * radius, tv, 0.0)
        if symmetry == "spherical":
            norm = 4.0 * math.pi / 3.0 * radius * radius * radius
        elif symmetry == "cylindrical":
            norm = math.pi * radius * radius
        else:
            raise ValueError("Bad blast wave symmetry parameter in <problem>!")
        ie = e0 * (1.0 - vol / tv) + internal_energy * vol / tv / norm
        prim[b, 0] = d0
        prim[b, 5] = ie / d0
    return prim


def perturbed_constant(mesh: UniformMesh, nvar_gas=6, nspecies_dust=0, gas_rho=10.0,
                       gas_v=(1.0, 0.0, 0.0), gas_sie=1.0, dust_rho=0.01, amp=1e-3, seed=1234):
    """Constant gas (+dust) state with a seeded smooth perturbation."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=8)
    prim = np.zeros(mesh.shape(nvar_gas))
    dprim = np.zeros(mesh.shape(4 * nspecies_dust)) if nspecies_dust else None
    for b in range(mesh.nb):
        x1v, x2v, x3v = cell_centers(mesh, b)
        X = 2 * np.pi * (x1v[None, None, :] - mesh.xmin[0]) / (mesh.xmax[0] - mesh.xmin[0])
        Y = 2 * np.pi * (x2v[None, :, None] - mesh.xmin[1]) / (mesh.xmax[1] - mesh.xmin[1])
        Z = 2 * np.pi * (x3v[:, None, None] - mesh.xmin[2]) / (mesh.xmax[2] - mesh.xmin[2])
        s1 = np.sin(X + ph[0]) * np.cos(Y + ph[1]) * np.cos(Z + ph[2])
        s2 = np.cos(2 * X + ph[3]) * np.sin(Y + ph[4]) * np.cos(Z + ph[5])
        s3 = np.sin(X + ph[6]) * np.sin(2 * Y + ph[7]) * np.sin(Z + ph[0])
        prim[b, 0] = gas_rho * (1.0 + amp * s1)
        prim[b, 1] = gas_v[0] + amp * s2
        prim[b, 2] = gas_v[1] + amp * s3
        prim[b, 3] = gas_v[2] + amp * s1 * s2
        prim[b, 5] = gas_sie * (1.0 + amp * s3)
        if dprim is not None:
            S = nspecies_dust
            for n in range(S):
                dprim[b, n] = dust_rho * (1.0 + amp * s2 * (n + 1))
                dprim[b, S + 3 * n + 0] = amp * s1 * (n + 1)
                dprim[b, S + 3 * n + 1] = amp * s3
                dprim[b, S + 3 * n + 2] = amp * s2
    return prim, dprim
