// trio.cuh -- the WARP-SPECIALISED single-pass stage kernel for 3-D Cartesian meshes.
//
// Same contract as k_sweep_stage (sweep.cuh): ONE kernel per stage does, for every interior
// zone, reconstruct (x1, x2, x3) -> Riemann (x1, x2, x3) -> flux divergence -> ApplyUpdate ->
// FluxSource -> SetAuxillaryFields -> ConsToPrim -> interior PrimToCons (-> CFL dt), i.e. the
// task chain src/artemis_driver.cpp:184-255 (+ interior of :261, + K13 on the last stage), in
// the reference's summation order (artemis_integrator.hpp:95-106, fluid_fluxes.hpp:365-392), so
// the strict build stays bit-identical to the reference.  Primitives and conserved state cross
// HBM once per stage; the kernel reads one primitive set and writes the other (ping-pong owned
// by sweep_host.cu).
//
// What is different: k_sweep_stage runs ONE role on 10 warps of 168 registers -- three CTA-wide
// barriers per plane, 36 % issue-slot utilisation (profiles/r01d).  Here a CTA still owns a
// 16 x 16 tile of (i, j) columns of one MeshBlock and sweeps it plane by plane along k through
// a 5-slot TMA ring (planes k .. k+3 are live during a step, the fifth slot gives the loads a
// full step of lead), but the work of a plane is split over THREE warp groups that run
// as a software pipeline, 26 warps of <= 72 registers per SM:
//
//   Z  (256 threads, one per column)  x3 march: interface value I(k+1|k+2) from the ring,
//        monotonise cell k+1, Riemann at face k+1 (left state / lower-face flux carried in
//        private shared-memory slots); then -- once the X and Y groups have published the
//        in-plane face fluxes of plane k -- gathers the six faces of its zone, updates,
//        C2P / P2C, stores, CFL dt.  Thread 0 also drives the TMA ring.
//   X  (288 threads = 16 rows x 18 cells)  x1: every thread reconstructs ONE cell of a row
//        (cells -1 .. 16: the tile plus one halo cell on either side, so there are no separate
//        halo warps and no idle phases), publishes its upper-edge state, then solves the
//        Riemann problem at its lower face (faces 0 .. 16) and publishes the 8 face quantities.
//   Y  (288 threads = 18 rows x 16 cells)  x2: the same along j.
//
// Both PPM interface values of a cell are computed by the cell's own thread (the value shared
// by two cells is the same expression of the same four operands, ppm.hpp:39-46, so it is
// bit-identical on both sides); that removes the interface-value exchange, one barrier and all
// halo special cases of k_sweep_stage for ~50 extra instructions per zone and direction.
//
// Synchronisation.  TMA ring: one "full" mbarrier per slot (transaction bytes) and one "empty"
// mbarrier per slot (one arrival per warp when the warp has read the plane).  Inside a group:
// one named barrier (bar.sync id, nthreads) between "publish upper-edge states" and "Riemann".
// Between groups: named-barrier producer/consumer pairs -- FULL_X / FULL_Y (X / Y arrive, Z
// syncs) and EMPTY_X / EMPTY_Y (Z arrives after the gather, X / Y sync before overwriting the
// flux planes).  A thread loads its left state BEFORE it syncs on EMPTY, so when the group
// passes that barrier every upper-edge state of the plane has been consumed and the next plane's
// may be published into the same slots.  There is no CTA-wide barrier inside the sweep.
#pragma once
#include "sweep.cuh"

namespace ab200 {

constexpr int kTrNZ = kSwTI * kSwTJ;         // one thread per column
constexpr int kTrNX = (kSwTI + 2) * kSwTJ;   // cells -1 .. TI of every row
constexpr int kTrNY = kSwTI * (kSwTJ + 2);   // cells -1 .. TJ of every column
constexpr int kTrThreads = kTrNZ + kTrNX + kTrNY;
constexpr int kTrWarps = kTrThreads / 32;
constexpr int kTrRing = 5;
static_assert(kTrNZ % 32 == 0 && kTrNX % 32 == 0 && kTrNY % 32 == 0, "whole warps per role");
static_assert(kTrThreads <= 1024, "CTA too large");

// named barriers (0 is __syncthreads)
enum : int { kBarX = 1, kBarY = 2, kBarFullX = 3, kBarEmptyX = 4, kBarFullY = 5, kBarEmptyY = 6 };

AB_D void nbar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
AB_D void nbar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
AB_D void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory carve-up, in doubles
template <int NV, int NF>
struct TrSmem {
  static constexpr int full = 0, empty = kTrRing;  // mbarriers (8 bytes each)
  static constexpr int ptrs = 2 * kTrRing;        // 3*NV CTA-uniform base pointers
  static constexpr int ring = 32;
  static constexpr int fxs = kSwTJ * (kSwTI + 1);   // x1 faces 0..TI of every row
  static constexpr int fys = (kSwTJ + 1) * kSwTI;   // x2 faces 0..TJ of every column
  static constexpr int qlx = ring + kTrRing * NV * kSwTile;  // [NV][fxs] upper-edge states
  static constexpr int qly = qlx + NV * fxs;                 // [NV][fys]
  static constexpr int fx = qly + NV * fys;                  // [NF][fxs] face quantities
  static constexpr int fy = fx + NF * fxs;                   // [NF][fys]
  static constexpr int fzs = fy + NF * fys;                  // [NF][NZ] x3 lower-face flux (private)
  static constexpr int qus = fzs + NF * kTrNZ;               // [NV][NZ] x3 upper-edge state (private)
  static constexpr int total = qus + NV * kTrNZ;
  static constexpr size_t bytes = (size_t)total * 8;
};
static_assert(TrSmem<6, 8>::bytes <= 227 * 1024, "trio tile does not fit shared memory");
static_assert(2 * kTrRing + 3 * 6 <= 32, "header overflow");

// both limited edge states of one cell from its 5-point (PPM) / 3-point (PLM) neighbourhood in
// shared memory; q points at the cell, s = stride along the direction
template <int RC>
AB_D void tr_recon(const double *q, int s, double &ql_up, double &qr_lo) {
  if (RC == AB200_PPM) {
    const double qm2 = q[-2 * s], qm1 = q[-s], q0 = q[0], qp1 = q[s], qp2 = q[2 * s];
    const double ilo = ppm_iface(qm2, qm1, q0, qp1);
    const double iup = ppm_iface(qm1, q0, qp1, qp2);
    ppm_mono(ilo, q0, iup, ql_up, qr_lo);
  } else if (RC == AB200_PLM) {
    plm(q[-s], q[0], q[s], ql_up, qr_lo);
  } else {
    ql_up = q[0];
    qr_lo = q[0];
  }
}

// MODE: ApplyUpdate base term at compile time, as in k_sweep_stage
//   0  stage 1 (gam0 = 0, gam1 = 1, u1 == u0 on entry): base = u0, u1 <- u0 folded in
//   1  gam0 == 0: base = gam1 * u1        2  general: base = gam0 * u0 + gam1 * u1
template <int FLUID, int RS, int RC, int MODE>
__global__ void __launch_bounds__(kTrThreads, 1)
k_trio_stage(GridDev g, FluidDev f, SweepArgs a) {
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr int NV = gas ? 6 : 4;
  constexpr int NF = gas ? 8 : 4;
  constexpr bool PPM = (RC == AB200_PPM);
  using SM = TrSmem<NV, NF>;
  constexpr int TI = kSwTI, TJ = kSwTJ, H = kSwH, PI = kSwPI;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  double *sm = reinterpret_cast<double *>(smem_raw);
  double **s_ptr = reinterpret_cast<double **>(sm + SM::ptrs);  // u0 | u1 | prim_out, NV each

  const int tid = threadIdx.x;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int b = blockIdx.y, n = blockIdx.z;
  const int S = f.S, nvar = f.nvar;
  const int i0 = g.is + tx * TI, j0 = g.js + ty * TJ;
  const int HX = H + ((g.is - H) & 1);  // even TMA start column (see sweep.cuh)
  const int nkr = g.ke - g.ks + 1;
  const int nplanes = nkr + 2 * H;  // planes ks-3 .. ke+3
  const int nsteps = nkr + H;       // step st works on plane k = ks-3+st

  const int pv0 = n, pv1 = S + 3 * n, pv4 = 4 * S + n, pv5 = 5 * S + n;
  auto pvx = [&](int m) { return m == 0 ? pv0 : (m < 4 ? pv1 + m - 1 : (m == 4 ? pv4 : pv5)); };

  if (tid < 3 * NV) {
    const int kind = tid / NV, m = tid - kind * NV;
    double *const *tab = kind == 0 ? f.u0 : (kind == 1 ? f.u1 : a.prim_out);
    s_ptr[tid] = tab ? tab[(size_t)b * nvar + pvx(m)] : nullptr;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTrRing; ++s) {
      mbar_init(bar + SM::full + s, 1);
      mbar_init(bar + SM::empty + s, kTrWarps);
    }
  }
  __syncthreads();

  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};

  if (tid < kTrNZ) {
    // ======================================= Z group ========================================
    const int ci = tid % TI, cj = tid / TI;
    const int i = i0 + ci, j = j0 + cj;
    const bool active = i <= g.ie && j <= g.je;
    const int ii = active ? i : g.is, jj = active ? j : g.js;
    const int pc = (cj + H) * PI + (ci + HX);  // own column inside a staged variable tile
    const int lane = tid & 31;
    const int plane = g.nj * g.ni;
    int offk = ((g.ks - H) * g.nj + jj) * g.ni + ii;
    const double dt = a.dt_dev ? *a.dt_dev : a.dt;
    const double bdt = a.beta * dt;

    auto issue = [&](int p, int slot) {  // thread 0: TMA loads of plane p into slot p % 5
      mbar_expect_tx(bar + SM::full + slot, (uint32_t)(kSwPI * kSwPJ * 8) * NV);
      double *dst = sm + SM::ring + slot * NV * kSwTile;
      const CUtensorMap *mp = a.maps + (size_t)b * nvar;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        tma_load_3d(dst + v * kSwTile, mp + pvx(v), bar + SM::full + slot, i0 - HX, j0 - H,
                    g.ks - H + p);
    };
    constexpr int kAhead = 3;
    auto prefetch = [&](int p) {
      const CUtensorMap *mp = a.maps + (size_t)b * nvar;
#pragma unroll
      for (int v = 0; v < NV; ++v) tma_prefetch_3d(mp + pvx(v), i0 - HX, j0 - H, g.ks - H + p);
    };
    if (tid == 0) {
      for (int p = 0; p < kTrRing && p < nplanes; ++p) issue(p, p);
      for (int p = kTrRing; p < kTrRing + kAhead && p < nplanes; ++p) prefetch(p);
    }

#ifdef AB200_FAST_MATH
    // Cartesian: A_d / V = 1 / dx_d, one reciprocal per thread and direction (the per-cell
    // value differs only in the last bit of the xmin + idx*dx differences)
    const double *x1f = g.t.x1f + (size_t)b * (g.ni + 1), *x2f = g.t.x2f + (size_t)b * (g.nj + 1);
    const double *x3f0 = g.t.x3f + (size_t)b * (g.nk + 1);
    const double rx = drcp(x1f[ii + 1] - x1f[ii]);
    const double ry = drcp(x2f[jj + 1] - x2f[jj]);
    const double rz = drcp(x3f0[g.ks + 1] - x3f0[g.ks]);
    double tden = 0.0;
#endif
    double Ilo[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) Ilo[v] = 0.0;
    double tmin = 1.79769313486231570815e+308;

    // planes 0..2 arrive before the first step; plane st+3 is waited for inside step st
    for (int p = 0; p < 3; ++p) mbar_wait(bar + SM::full + p, 0);
    int s0 = 0;         // ring slot of plane st (= st % 5), kept incrementally
    uint32_t ph0 = 0;   // phase parity of plane st (= (st / 5) & 1)

    for (int st = 0; st < nsteps; ++st) {
      const bool INP = st >= H;   // plane k is an interior plane: the zone is updated
      const bool ZR = st >= 2;    // the x3 Riemann problem at face k+1 is solved
      const int k = g.ks - H + st;
      const int s1 = s0 + 1 >= kTrRing ? s0 + 1 - kTrRing : s0 + 1;
      const int s2 = s0 + 2 >= kTrRing ? s0 + 2 - kTrRing : s0 + 2;
      const int s3 = s0 + 3 >= kTrRing ? s0 + 3 - kTrRing : s0 + 3;
      // ring refill: plane st-1 is dead once every warp has read it; its slot takes plane st+4,
      // which is first needed one step from now
      if (tid == 0 && st >= 1) {
        const int p = st - 1;
        const int sp = s0 == 0 ? kTrRing - 1 : s0 - 1;
        if (p + kTrRing < nplanes) {
          mbar_wait(bar + SM::empty + sp, s0 == 0 ? ph0 ^ 1u : ph0);
          issue(p + kTrRing, sp);
        }
        if (p + kTrRing + kAhead < nplanes) prefetch(p + kTrRing + kAhead);
      }
      // the conserved state of the NEXT plane's zones is pulled into L1 now, a whole step before
      // the update reads it (the loads sit on the Z group's critical path)
      if (st + 1 >= H && st + 1 < nsteps) {
        const int offn = offk + plane;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          if (MODE != 1) prefetch_l1(s_ptr[m] + offn);
          if (MODE != 0) prefetch_l1(s_ptr[NV + m] + offn);
        }
      }
      mbar_wait(bar + SM::full + s3, s0 + 3 >= kTrRing ? ph0 ^ 1u : ph0);
      const double *R0 = sm + SM::ring + s0 * NV * kSwTile + pc;  // plane k
      const double *R1 = sm + SM::ring + s1 * NV * kSwTile + pc;
      const double *R2 = sm + SM::ring + s2 * NV * kSwTile + pc;
      const double *R3 = sm + SM::ring + s3 * NV * kSwTile + pc;

      // ---- x3: interface value I(k+1|k+2), cell k+1, Riemann at face k+1 -------------------
      double QupN[NV], qrz[NV], FzN[NF];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double W1 = R1[v * kSwTile];
        if (PPM) {
          const double Inew = ppm_iface(R0[v * kSwTile], W1, R2[v * kSwTile], R3[v * kSwTile]);
          ppm_mono(Ilo[v], W1, Inew, QupN[v], qrz[v]);
          Ilo[v] = Inew;
        } else if (RC == AB200_PLM) {
          plm(R0[v * kSwTile], W1, R2[v * kSwTile], QupN[v], qrz[v]);
        } else {
          QupN[v] = W1;
          qrz[v] = W1;
        }
      }
      // plane k is not needed by this warp any more
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + SM::empty + s0);
#pragma unroll
      for (int m = 0; m < NF; ++m) FzN[m] = 0.0;
      if (ZR) {  // recon order (rho, v3, v1, v2, P, sie)
        double wl[NV], wr[NV], out[8];
        const double *Qup = sm + SM::qus + tid;
        wl[0] = Qup[0]; wl[1] = Qup[3 * kTrNZ]; wl[2] = Qup[1 * kTrNZ]; wl[3] = Qup[2 * kTrNZ];
        wr[0] = qrz[0]; wr[1] = qrz[3]; wr[2] = qrz[1]; wr[3] = qrz[2];
        if (gas) {
          wl[4] = Qup[4 * kTrNZ]; wl[5] = Qup[5 * kTrNZ];
          wr[4] = qrz[4]; wr[5] = qrz[5];
        }
        Riemann<RS, FLUID>::solve(eos, wl, wr, out);
        FzN[0] = out[0]; FzN[3] = out[1]; FzN[1] = out[2]; FzN[2] = out[3];
        if (gas) { FzN[4] = out[4]; FzN[5] = out[5]; FzN[6] = out[6]; FzN[7] = out[7]; }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) sm[SM::qus + v * kTrNZ + tid] = QupN[v];

      if (INP) {
        const int off = offk;
        // in-plane face quantities of plane k are complete
        nbar_sync(kBarFullX, kTrNX + kTrNZ);
        nbar_sync(kBarFullY, kTrNY + kTrNZ);
        const int ox = cj * (TI + 1) + ci, oy = cj * TI + ci;
        double u[6];
#ifdef AB200_FAST_MATH
        double ub[6];
#else
        Coords<AB200_CARTESIAN> cc(g, b, k, jj, ii);
        const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
        const double ax2[2] = {cc.area2(0), cc.area2(1)};
        const double ax3[2] = {cc.area3(), cc.area3()};
        const double vol = cc.volume();
#endif
        double Fz6 = 0.0, Fz7 = 0.0;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const double xl = sm[SM::fx + m * SM::fxs + ox], xh = sm[SM::fx + m * SM::fxs + ox + 1];
          const double yl = sm[SM::fy + m * SM::fys + oy], yh = sm[SM::fy + m * SM::fys + oy + TI];
          const double zl = sm[SM::fzs + m * kTrNZ + tid];
          double base;
          if (MODE == 0) {
            base = __ldg(s_ptr[m] + off);
            if (active) __stcg(s_ptr[NV + m] + off, base);  // u1 <- u0
          } else if (MODE == 1) {
            base = a.gam1 * __ldg(s_ptr[NV + m] + off);
          } else {
            base = a.gam0 * __ldg(s_ptr[m] + off) + a.gam1 * __ldg(s_ptr[NV + m] + off);
          }
          // ApplyUpdate (artemis_integrator.hpp:95-106)
#ifdef AB200_FAST_MATH
          u[m] = (xl - xh) * rx + (yl - yh) * ry + (zl - FzN[m]) * rz;
          ub[m] = base;
#else
          double divf = (ax1[0] * xl - ax1[1] * xh);
          divf += (ax2[0] * yl - ax2[1] * yh);
          divf += (ax3[0] * zl - ax3[1] * FzN[m]);
          u[m] = base + divf * bdt / vol;
#endif
        }
        if (gas) {  // FluxSource (fluid_fluxes.hpp:365-392), direction by direction
          Fz6 = sm[SM::fzs + 6 * kTrNZ + tid];
          Fz7 = sm[SM::fzs + 7 * kTrNZ + tid];
          const double pxl = sm[SM::fx + 6 * SM::fxs + ox], pxh = sm[SM::fx + 6 * SM::fxs + ox + 1];
          const double vxl = sm[SM::fx + 7 * SM::fxs + ox], vxh = sm[SM::fx + 7 * SM::fxs + ox + 1];
          const double pyl = sm[SM::fy + 6 * SM::fys + oy], pyh = sm[SM::fy + 6 * SM::fys + oy + TI];
          const double vyl = sm[SM::fy + 7 * SM::fys + oy], vyh = sm[SM::fy + 7 * SM::fys + oy + TI];
#ifdef AB200_FAST_MATH
          u[1] += rx * (pxl - pxh);
          u[5] -= rx * 0.5 * (pxl + pxh) * (vxh - vxl);
          u[2] += ry * (pyl - pyh);
          u[5] -= ry * 0.5 * (pyl + pyh) * (vyh - vyl);
          u[3] += rz * (Fz6 - FzN[6]);
          u[5] -= rz * 0.5 * (Fz6 + FzN[6]) * (FzN[7] - Fz7);
#else
          const double dx1 = cc.x1[1] - cc.x1[0], dx2 = cc.x2[1] - cc.x2[0],
                       dx3 = cc.x3[1] - cc.x3[0];
          u[1] += bdt / dx1 * (pxl - pxh);
          u[5] -= bdt / vol * 0.5 * (pxl + pxh) * (ax1[1] * vxh - ax1[0] * vxl);
          u[2] += bdt / dx2 * (pyl - pyh);
          u[5] -= bdt / vol * 0.5 * (pyl + pyh) * (ax2[1] * vyh - ax2[0] * vyl);
          u[3] += bdt / dx3 * (Fz6 - FzN[6]);
          u[5] -= bdt / vol * 0.5 * (Fz6 + FzN[6]) * (ax3[1] * FzN[7] - ax3[0] * Fz7);
#endif
        }
        // the flux planes of this step are consumed: X / Y may publish the next plane
        if (st + 1 < nsteps) {
          nbar_arrive(kBarEmptyX, kTrNX + kTrNZ);
          nbar_arrive(kBarEmptyY, kTrNY + kTrNZ);
        }
#ifdef AB200_FAST_MATH
#pragma unroll
        for (int m = 0; m < NV; ++m) u[m] = fma(bdt, u[m], ub[m]);
#endif
        if (active) {
          double **pp = s_ptr + 2 * NV, **pu = s_ptr;
#ifdef AB200_FAST_MATH
          // SetAuxillaryFields + ConsToPrim + PrimToCons (fill_derived.cpp:55-72, 129-164,
          // 217-274) with ONE reciprocal of the floored density (hx == 1 on Cartesian meshes)
          const double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
          const double rwd = drcp(w_d);
          const double v1 = u[1] * rwd, v2 = u[2] * rwd, v3 = u[3] * rwd;
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
          double w_s = 0.0;
          if (gas) {
            const double ue = u[4] - ke;
            double sie = ((ue > f.de_switch * u[4]) ? ue : u[5]) * rwd;
            sie = dmax(sie, f.siefloor);
            w_s = sie;  // == (sie * w_d) / w_d up to rounding, floors already applied
          }
#else
          const double hx[3] = {1.0, 1.0, 1.0};
          if (gas)  // SetAuxillaryFields (fill_derived.cpp:55-72)
            u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor,
                                f.de_switch);
          // ConsToPrim (fill_derived.cpp:129-164)
          double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
          const double v1 = u[1] / (w_d * hx[0]), v2 = u[2] / (w_d * hx[1]),
                       v3 = u[3] / (w_d * hx[2]);
          w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
          double w_s = 0.0;
          if (gas) {
            w_s = u[5] / ((u[0] > f.dfloor) ? u[0] : f.dfloor);
            w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          }
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
#endif
          // PrimToCons on the just-computed primitives (fill_derived.cpp:217-274)
          __stcg(pp[0] + off, w_d);
          __stcg(pp[1] + off, v1);
          __stcg(pp[2] + off, v2);
          __stcg(pp[3] + off, v3);
          __stcg(pu[0] + off, w_d);
          __stcg(pu[1] + off, w_d * v1);
          __stcg(pu[2] + off, w_d * v2);
          __stcg(pu[3] + off, w_d * v3);
          if (gas) {
            const double u_u = w_s * w_d;
            __stcg(pp[5] + off, w_s);
            __stcg(pp[4] + off, dmax(0.0, f.gm1 * w_d * w_s));
            __stcg(pu[5] + off, u_u);
            __stcg(pu[4] + off, u_u + ke);
          }
          if (a.dt_min) {  // EstimateTimestepMesh folded in (src/gas/gas.cpp:411-433)
#ifdef AB200_FAST_MATH
            double cs = 0.0;
            if (gas) cs = dsqrt(dmax(0.0, (f.gm1 + 1) * f.gm1 * w_d * w_s) * rwd);
            tden = dmax(tden, (fabs(v1) + cs) * rx + (fabs(v2) + cs) * ry + (fabs(v3) + cs) * rz);
#else
            Coords<AB200_CARTESIAN> cd(g, b, k, j, i);
            const double vel[3] = {v1, v2, v3};
            tmin = dmin(tmin, cell_dt<AB200_CARTESIAN, FLUID>(g, f, cd, w_d, vel, w_s));
#endif
          }
        }
      }
      if (ZR) {
#pragma unroll
        for (int m = 0; m < NF; ++m) sm[SM::fzs + m * kTrNZ + tid] = FzN[m];
      }
      offk += plane;
      if (++s0 == kTrRing) { s0 = 0; ph0 ^= 1u; }
    }
    if (a.dt_min) {  // warp-shuffle min, one atomic per warp
#ifdef AB200_FAST_MATH
      if (tden > 0.0) tmin = drcp(tden);
#endif
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmin = dmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
      if (lane == 0) atomicMin(a.dt_min, (unsigned long long)__double_as_longlong(tmin));
    }
  } else {
    // ==================================== X and Y groups ====================================
    const bool isx = tid < kTrNZ + kTrNX;
    const int t = isx ? tid - kTrNZ : tid - kTrNZ - kTrNX;
    const int lane = tid & 31;
    // X: row r = t / 18, cell c = t % 18 - 1 along i.  Y: column r = t % 16, cell c = t / 16 - 1
    // along j.  `c` runs along the group's direction; faces 0 .. T of a pencil need the upper
    // edge of cells -1 .. T-1 and the lower edge of cells 0 .. T.
    const int r = isx ? t / (TI + 2) : t % TI;
    const int c = isx ? t % (TI + 2) - 1 : t / TI - 1;
    const int T = isx ? TI : TJ;
    const int sdir = isx ? 1 : PI;  // stride along the direction inside a staged tile
    const int pcell = isx ? (r + H) * PI + (c + HX) : (c + H) * PI + (r + HX);
    // slot of face `fc` of this pencil: X [row][face], Y [face][column]
    const int fs = isx ? SM::fxs : SM::fys;
    const int fslot = isx ? r * (TI + 1) + c : c * TI + r;       // own lower face (c >= 0)
    const int uslot = isx ? r * (TI + 1) + c + 1 : (c + 1) * TI + r;  // own upper face (c < T)
    const int qlb = isx ? SM::qlx : SM::qly;
    const int fb = isx ? SM::fx : SM::fy;
    const int bar_grp = isx ? kBarX : kBarY, n_grp = isx ? kTrNX : kTrNY;
    const int bar_full = isx ? kBarFullX : kBarFullY, bar_empty = isx ? kBarEmptyX : kBarEmptyY;
    const int n_pc = n_grp + kTrNZ;

    // the three warm-up planes are read by the Z group only
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar + SM::empty + 0);
      mbar_arrive(bar + SM::empty + 1);
      mbar_arrive(bar + SM::empty + 2);
    }
    int s0 = H;        // ring slot / phase parity of plane st, kept incrementally
    uint32_t ph0 = 0;
    double *ql_pub = sm + qlb;
    for (int st = H; st < nsteps; ++st) {
      mbar_wait(bar + SM::full + s0, ph0);
      const double *R0 = sm + SM::ring + s0 * NV * kSwTile + pcell;
      double qr[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        double ql;
        tr_recon<RC>(R0 + v * kSwTile, sdir, ql, qr[v]);
        if (c < T) ql_pub[v * fs + uslot] = ql;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + SM::empty + s0);
      if (++s0 == kTrRing) { s0 = 0; ph0 ^= 1u; }
      nbar_sync(bar_grp, n_grp);                    // upper-edge states of the plane published
      // left state of the own face, in recon order: X (rho, v1, v2, v3, P, sie) = pack order;
      // Y (rho, v2, v3, v1, P, sie).  Loaded BEFORE the EMPTY sync: once the group has passed it,
      // every published state has been consumed and the slots may be overwritten.
      double wl[NV], wr[NV];
      {
        const int sl = c >= 0 ? fslot : uslot;  // cell -1 has no lower face: any valid slot
        if (isx) {
#pragma unroll
          for (int v = 0; v < NV; ++v) { wl[v] = ql_pub[v * fs + sl]; wr[v] = qr[v]; }
        } else {
          wl[0] = ql_pub[0 * fs + sl]; wl[1] = ql_pub[2 * fs + sl];
          wl[2] = ql_pub[3 * fs + sl]; wl[3] = ql_pub[1 * fs + sl];
          wr[0] = qr[0]; wr[1] = qr[2]; wr[2] = qr[3]; wr[3] = qr[1];
          if (gas) {
            wl[4] = ql_pub[4 * fs + sl]; wl[5] = ql_pub[5 * fs + sl];
            wr[4] = qr[4]; wr[5] = qr[5];
          }
        }
      }
      if (st > H) nbar_sync(bar_empty, n_pc);       // Z has gathered the previous flux plane
      else nbar_sync(bar_grp, n_grp);               // first plane: "all left states loaded"
      if (c >= 0) {
        double out[8];
        Riemann<RS, FLUID>::solve(eos, wl, wr, out);
        double *fp = sm + fb + fslot;
        if (isx) {
#pragma unroll
          for (int m = 0; m < NF; ++m) fp[m * fs] = out[m];
        } else {
          fp[0 * fs] = out[0]; fp[2 * fs] = out[1]; fp[3 * fs] = out[2]; fp[1 * fs] = out[3];
          if (gas) {
#pragma unroll
            for (int m = 4; m < 8; ++m) fp[m * fs] = out[m];
          }
        }
      }
      nbar_arrive(bar_full, n_pc);
    }
  }
}

}  // namespace ab200
