// sweep.cuh -- the single-pass stage kernel for 3-D Cartesian meshes ("plane sweep").
//
// ONE kernel per stage does, for every interior zone,
//   reconstruct (x1, x2, x3) -> Riemann (x1, x2, x3) -> flux divergence -> ApplyUpdate ->
//   FluxSource -> SetAuxillaryFields -> ConsToPrim -> interior PrimToCons (-> CFL dt)
// i.e. the whole task chain src/artemis_driver.cpp:184-255 (+ interior of :261, + K13 when the
// stage is the last of a cycle), in the reference's own summation order: the three directional
// flux differences are added into ONE divf before the update exactly like
// artemis_integrator.hpp:95-106, and the three FluxSource terms follow in direction order
// (fluid_fluxes.hpp:365-392).  Primitives and conserved state cross HBM once per stage.
//
// Decomposition.  A CTA owns a TI x TJ tile of (i, j) columns of ONE MeshBlock and sweeps it
// plane by plane along k.  The primitive planes (tile + 3 halo cells in i and j) are brought
// into a 4-slot shared-memory ring by TMA (cp.async.bulk.tensor.3d, one box per variable,
// completion on an mbarrier per slot); plane k+4 is in flight while plane k is worked on.
// Thread (ci, cj) owns one column:
//   x3: marched.  The PPM interface value I(k+1|k+2) is formed from the ring planes k..k+3,
//       cell k+1 is monotonised, the Riemann problem at face k+1 is solved; the lower-face
//       flux F3(k), the interface value and the upper-edge state are carried in registers
//       (20 doubles), so every x3 interface value / limiter / Riemann solve is done once.
//   x1, x2: in-plane through shared memory, every quantity computed once per face:
//       P1  lower interface value of the own cell          -> smem
//       P2  monotonise the own cell (upper neighbour's interface value from smem);
//           upper-edge state                               -> smem
//       P3  Riemann at the own lower face (left state from smem); 8 face quantities -> smem
//       P4  gather the four in-plane faces + the two x3 faces, update, C2P/P2C, store.
//   Two extra "halo" warps do the tile-edge work (cells -1 and TI/TJ of every row/column, and
//   the last face), so the main warps stay convergent and the phases stay balanced.
// u0 / u1 of the own zone are fetched with cp.async at the top of the step into a private
// shared-memory slot, so no register is held across the Riemann phase for them.
//
// The kernel READS one primitive set and WRITES another (prim_out): tiles of one MeshBlock
// are swept by different CTAs in no particular order and read each other's cells as halo, so
// the stage cannot update the primitives in place.  The host side ping-pongs between the
// caller's arrays and a library-owned alternate set (sweep.cu).
#pragma once
#include <type_traits>

#include "march.cuh"

namespace ab200 {

#ifndef AB200_SW_TI
#define AB200_SW_TI 16
#endif
#ifndef AB200_SW_TJ
#define AB200_SW_TJ 16
#endif
#ifndef AB200_SW_HALO_THREADS
#define AB200_SW_HALO_THREADS 64
#endif
constexpr int kSwTI = AB200_SW_TI, kSwTJ = AB200_SW_TJ, kSwH = 3;
// TMA needs a 16-byte aligned global start address along i (an odd fp64 start coordinate is an
// illegal instruction on sm_100a), so the staged row starts at the even column i0-3 or i0-4
// and is TI+8 wide; HX (3 or 4, CTA-uniform) is the own tile's column offset inside the row.
constexpr int kSwPI = kSwTI + 2 * kSwH + 2, kSwPJ = kSwTJ + 2 * kSwH;
constexpr int kSwMain = kSwTI * kSwTJ;
constexpr int kSwHalo = AB200_SW_HALO_THREADS;
constexpr int kSwThreads = kSwMain + kSwHalo;
constexpr int kSwRing = 4;
// doubles per staged variable tile (TMA destinations are 128-byte aligned)
constexpr int kSwTile = ((kSwPI * kSwPJ * 8 + 127) / 128) * 16;
static_assert(kSwMain % 32 == 0 && kSwHalo % 32 == 0, "whole warps per role");
static_assert(kSwTI % 2 == 0, "tile origin parity must not depend on the tile index");
static_assert((kSwPI * 8) % 16 == 0, "TMA inner box extent must be a multiple of 16 bytes");

struct SweepArgs {
  double gam0, gam1, beta, dt;
  const double *dt_dev;  // if non-null: dt = *dt_dev
  int copy_u1;           // stage 1: u1 <- u0 (DeepCopyConservedData folded in)
  int tiles_x, tiles_y;
  const CUtensorMap *maps;   // [nb*nvar] tensor maps of the INPUT primitive set, box {PI,PJ,1}
  double *const *prim_out;   // [nb*nvar] OUTPUT primitive set
  unsigned long long *dt_min;
};

// shared-memory carve-up, in doubles
template <int NV, int NF>
struct SwSmem {
  static constexpr int ptrs = 8;    // after the 4 mbarriers: 3*NV CTA-uniform base pointers
  static constexpr int ring = 32;   // 256 bytes of header
  static constexpr int ix = ring + kSwRing * NV * kSwTile;
  static constexpr int ix_vs = kSwTJ * (kSwTI + 3);       // lower iface of cells -1..TI+1
  static constexpr int iy = ix + NV * ix_vs;
  static constexpr int iy_vs = (kSwTJ + 3) * kSwTI;
  // Left states at faces 0..TI (P2 -> P3) and, in the SAME slots, the face quantities the
  // Riemann solve of that face produces (P3 -> P4): a face slot is read and then overwritten
  // by the one thread that owns the face, so the two generations can share memory.
  static constexpr int qlx = iy + NV * iy_vs;
  static constexpr int qlx_vs = kSwTJ * (kSwTI + 1);
  static constexpr int fx = qlx, fx_vs = qlx_vs;
  static constexpr int qly = qlx + NF * qlx_vs;
  static constexpr int qly_vs = (kSwTJ + 1) * kSwTI;
  static constexpr int fy = qly, fy_vs = qly_vs;
  static constexpr int qhx = qly + NF * qly_vs;           // right state at face TI
  static constexpr int qhx_vs = kSwTJ;
  static constexpr int qhy = qhx + NV * qhx_vs;
  static constexpr int qhy_vs = kSwTI;
  static constexpr int ust = qhy + NV * qhy_vs;           // [2*NV][kSwMain] u0 | u1 of own zone
  // x3 march state kept out of the register file (private slot per thread): face quantities
  // at the lower x3 face of the current cell, upper-edge state of the cell below
  static constexpr int fzs = ust + 2 * NV * kSwMain;      // [NF][kSwMain]
  static constexpr int qus = fzs + NF * kSwMain;          // [NV][kSwMain]
  static constexpr int total = qus + NV * kSwMain;
  static constexpr size_t bytes = (size_t)total * 8;
};
static_assert(SwSmem<6, 8>::bytes <= 227 * 1024, "sweep tile does not fit shared memory");
static_assert(8 + 3 * 6 <= 32, "header overflow");

AB_D void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no completion to wait for)
AB_D void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
AB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
AB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// reconstruct one variable of one cell from shared memory (q points at the cell, s = stride
// along the direction); PPM takes its two limited interface values
template <int RC>
AB_D void sw_recon(const double *q, int s, double ilo, double iup, double &ql_up, double &qr_lo) {
  if (RC == AB200_PPM) {
    ppm_mono(ilo, q[0], iup, ql_up, qr_lo);
  } else if (RC == AB200_PLM) {
    plm(q[-s], q[0], q[s], ql_up, qr_lo);
  } else {
    ql_up = q[0];
    qr_lo = q[0];
  }
}

// MODE selects the ApplyUpdate base term (artemis_integrator.hpp:95-106) at compile time:
//   0  stage 1 of every integrator here (gam0 = 0, gam1 = 1, u1 == u0 on entry): base = u0 and
//      DeepCopyConservedData is folded in (u1 <- u0); u1 is never read
//   1  gam0 == 0: base = gam1 * u1; u0 is never read
//   2  general:   base = gam0 * u0 + gam1 * u1
template <int FLUID, int RS, int RC, int MODE>
__global__ void __launch_bounds__(kSwThreads, 1)
k_sweep_stage(GridDev g, FluidDev f, SweepArgs a) {
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr int NV = gas ? 6 : 4;   // reconstructed variables per species
  constexpr int NF = gas ? 8 : 4;   // face quantities per species
  constexpr bool PPM = (RC == AB200_PPM);
  constexpr bool need_u0 = (MODE != 1), need_u1 = (MODE != 0);
  using SM = SwSmem<NV, NF>;
  constexpr int TI = kSwTI, TJ = kSwTJ, H = kSwH, PI = kSwPI;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  double *sm = reinterpret_cast<double *>(smem_raw);
  double **s_ptr = reinterpret_cast<double **>(sm + SM::ptrs);  // u0 | u1 | prim_out, NV each

  const int tid = threadIdx.x;
  const bool is_main = tid < kSwMain;
  const int ci = tid % TI, cj = tid / TI;
  const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
  const int b = blockIdx.y, n = blockIdx.z;
  const int S = f.S, nvar = f.nvar;
  const int i0 = g.is + tx * TI, j0 = g.js + ty * TJ;
  const int HX = H + ((g.is - H) & 1);  // even TMA start column
  const int i = i0 + ci, j = j0 + cj;
  const bool active = is_main && i <= g.ie && j <= g.je;
  const int nkr = g.ke - g.ks + 1;
  const int nplanes = nkr + 2 * H;  // planes ks-3 .. ke+3
  const int nsteps = nkr + H;       // k = ks-3 .. ke (the first three steps only warm up x3)

  // pack order: rho, v1|m1, v2|m2, v3|m3, (P|E, sie|u)
  const int pv0 = n, pv1 = S + 3 * n, pv4 = 4 * S + n, pv5 = 5 * S + n;
  auto pvx = [&](int m) { return m == 0 ? pv0 : (m < 4 ? pv1 + m - 1 : (m == 4 ? pv4 : pv5)); };

  if (tid < 3 * NV) {
    const int kind = tid / NV, m = tid - kind * NV;
    double *const *tab = kind == 0 ? f.u0 : (kind == 1 ? f.u1 : a.prim_out);
    s_ptr[tid] = tab ? tab[(size_t)b * nvar + pvx(m)] : nullptr;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSwRing; ++s) mbar_init(bar + s, 1);
  }
  __syncthreads();

  auto issue = [&](int p) {  // elected thread: TMA loads of plane p into slot p % 4
    const int slot = p & (kSwRing - 1);
    mbar_expect_tx(bar + slot, (uint32_t)(kSwPI * kSwPJ * 8) * NV);
    double *dst = sm + SM::ring + slot * NV * kSwTile;
    const CUtensorMap *mp = a.maps + (size_t)b * nvar;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      tma_load_3d(dst + v * kSwTile, mp + pvx(v), bar + slot, i0 - HX, j0 - H, g.ks - H + p);
  };
  // planes further ahead than the ring can hold are pulled into L2 so that the TMA load that
  // eventually fills a freed slot does not pay the DRAM latency
  constexpr int kAhead = 3;
  auto prefetch = [&](int p) {
    const CUtensorMap *mp = a.maps + (size_t)b * nvar;
#pragma unroll
    for (int v = 0; v < NV; ++v) tma_prefetch_3d(mp + pvx(v), i0 - HX, j0 - H, g.ks - H + p);
  };
  if (tid == 0) {
    for (int p = 0; p < kSwRing && p < nplanes; ++p) issue(p);
    for (int p = kSwRing; p < kSwRing + kAhead && p < nplanes; ++p) prefetch(p);
  }

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};

  const int pc = (cj + H) * PI + (ci + HX);  // own column inside a staged variable tile
  const int h = tid - kSwMain;               // halo-thread index
  const int ii = active ? i : g.is, jj = active ? j : g.js;
  const double *x3f = g.t.x3f + (size_t)b * (g.nk + 1);
  const int plane = g.nj * g.ni;
  int offk = ((g.ks - H) * g.nj + jj) * g.ni + ii;  // own zone in the current plane (carried)

#ifdef AB200_FAST_MATH
  // Cartesian: A_d / V = 1 / dx_d; one reciprocal per thread (x1, x2) and per plane (x3).
  // The flux divergence and the FluxSource terms of a component are summed first and applied
  // with one fma(beta*dt, sum, base).
  double rx = 0.0, ry = 0.0;
  if (is_main) {
    const double *x1f = g.t.x1f + (size_t)b * (g.ni + 1), *x2f = g.t.x2f + (size_t)b * (g.nj + 1);
    rx = drcp(x1f[ii + 1] - x1f[ii]);
    ry = drcp(x2f[jj + 1] - x2f[jj]);
  }
#endif

  // carried across planes (x3 march): interface value I(k|k+1) in registers; the upper-edge
  // state of cell k (qus) and the face quantities at the lower face of cell k (fzs, momentum
  // fluxes in pack order) in private shared-memory slots
  double Ilo[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) Ilo[v] = 0.0;
  double tmin = 1.79769313486231570815e+308;
  double tden = 0.0;  // fast build: max over zones of sum_d (|v_d| + c_s) / dx_d

  // One plane step.  INP: the plane k = ks-3+st is an interior plane (in-plane x1/x2 work and
  // the zone update happen); ZR: the x3 Riemann problem at face k+1 is solved (st >= 2).
  // The first three steps (planes ks-3 .. ks-1) only warm up the x3 march and are instantiated
  // without any in-plane code, so the steady-state loop body carries no `inplane` branches.
  auto step = [&](auto inp_tag, auto zr_tag, const int st) {
    constexpr bool INP = decltype(inp_tag)::value;
    constexpr bool ZR = decltype(zr_tag)::value;
    const int k = g.ks - H + st;
    const double *R0 = sm + SM::ring + ((st + 0) & 3) * NV * kSwTile;  // plane k
    const double *R1 = sm + SM::ring + ((st + 1) & 3) * NV * kSwTile;
    const double *R2 = sm + SM::ring + ((st + 2) & 3) * NV * kSwTile;
    const double *R3 = sm + SM::ring + ((st + 3) & 3) * NV * kSwTile;
    const int off = offk;

    double W0[NV], Ixl[NV], Iyl[NV];
    // =========================== P1: interface values ==========================================
    if (is_main) {
#pragma unroll
      for (int v = 0; v < NV; ++v) W0[v] = R0[v * kSwTile + pc];
      if (INP) {
        if (active) {  // own zone's u0 / u1 -> private shared-memory slots, asynchronously
#pragma unroll
          for (int m = 0; m < NV; ++m) {
            if (need_u0) cp_async8(sm + SM::ust + m * kSwMain + tid, s_ptr[m] + off);
            if (need_u1) cp_async8(sm + SM::ust + (NV + m) * kSwMain + tid, s_ptr[NV + m] + off);
          }
          cp_async_commit();
        }
        if (PPM) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const double *q = R0 + v * kSwTile + pc;
            Ixl[v] = ppm_iface(q[-2], q[-1], W0[v], q[1]);
            sm[SM::ix + v * SM::ix_vs + cj * (TI + 3) + ci + 1] = Ixl[v];
            Iyl[v] = ppm_iface(q[-2 * PI], q[-PI], W0[v], q[PI]);
            sm[SM::iy + v * SM::iy_vs + (cj + 1) * TI + ci] = Iyl[v];
          }
        }
      }
    } else if (INP && PPM) {
      // halo: lower interface value of cells -1, TI, TI+1 of every row (x1) and of rows
      // -1, TJ, TJ+1 of every column (x2)
      for (int t = h; t < 3 * TJ + 3 * TI; t += kSwHalo) {
        int p, s, dst, vs;
        if (t < 3 * TJ) {
          const int r = t / 3, e = t - 3 * r, c = e == 0 ? -1 : TI + e - 1;
          p = (r + H) * PI + c + HX; s = 1;
          dst = SM::ix + r * (TI + 3) + c + 1; vs = SM::ix_vs;
        } else {
          const int tt = t - 3 * TJ, c = tt % TI, e = tt / TI, r = e == 0 ? -1 : TJ + e - 1;
          p = (r + H) * PI + c + HX; s = PI;
          dst = SM::iy + (r + 1) * TI + c; vs = SM::iy_vs;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double *q = R0 + v * kSwTile + p;
          sm[dst + v * vs] = ppm_iface(q[-2 * s], q[-s], q[0], q[s]);
        }
      }
    }
    // (also orders P4 of the previous step, which reads the face slots, before P2 refills them)
    if (INP) __syncthreads();

    // =========================== P2: limit the cells ===========================================
    // plane k+3 (needed by the x3 interface value only) is waited for here so that its TMA
    // latency hides behind P3/P4 of the previous step and P1 of this one
    double Inew[NV], QupN[NV], qrx[NV], qry[NV], FzN[NF];
    mbar_wait(bar + ((st + 3) & (kSwRing - 1)), (uint32_t)(((st + 3) >> 2) & 1));
    if (is_main) {
#pragma unroll
      double qrz[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {  // x3: interface value I(k+1|k+2), then cell k+1
        const double W1 = R1[v * kSwTile + pc];
        if (PPM) {
          Inew[v] = ppm_iface(W0[v], W1, R2[v * kSwTile + pc], R3[v * kSwTile + pc]);
          ppm_mono(Ilo[v], W1, Inew[v], QupN[v], qrz[v]);
        } else if (RC == AB200_PLM) {
          plm(W0[v], W1, R2[v * kSwTile + pc], QupN[v], qrz[v]);
        } else {
          QupN[v] = W1;
          qrz[v] = W1;
        }
      }
      // the x3 Riemann problem is solved here, before the in-plane limiting, so that Qup / qrz
      // are dead when the x1 / x2 states become live (register pressure of P3)
      if (ZR) {  // x3, face k+1: recon order (rho, v3, v1, v2, P, sie)
        double wl[NV], wr[NV], out[8];
        const double *Qup = sm + SM::qus + tid;
        wl[0] = Qup[0]; wl[1] = Qup[3 * kSwMain]; wl[2] = Qup[1 * kSwMain];
        wl[3] = Qup[2 * kSwMain];
        wr[0] = qrz[0]; wr[1] = qrz[3]; wr[2] = qrz[1]; wr[3] = qrz[2];
        if (gas) {
          wl[4] = Qup[4 * kSwMain]; wl[5] = Qup[5 * kSwMain];
          wr[4] = qrz[4]; wr[5] = qrz[5];
        }
        Riemann<RS, FLUID>::solve(eos, wl, wr, out);
        FzN[0] = out[0]; FzN[3] = out[1]; FzN[1] = out[2]; FzN[2] = out[3];
        if (gas) { FzN[4] = out[4]; FzN[5] = out[5]; FzN[6] = out[6]; FzN[7] = out[7]; }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) sm[SM::qus + v * kSwMain + tid] = QupN[v];
      if (INP) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double *q = R0 + v * kSwTile + pc;
          double ql, iup = 0.0;
          if (PPM) iup = sm[SM::ix + v * SM::ix_vs + cj * (TI + 3) + ci + 2];
          sw_recon<RC>(q, 1, Ixl[v], iup, ql, qrx[v]);
          sm[SM::qlx + v * SM::qlx_vs + cj * (TI + 1) + ci + 1] = ql;
          if (PPM) iup = sm[SM::iy + v * SM::iy_vs + (cj + 2) * TI + ci];
          sw_recon<RC>(q, PI, Iyl[v], iup, ql, qry[v]);
          sm[SM::qly + v * SM::qly_vs + (cj + 1) * TI + ci] = ql;
        }
      }
    } else if (INP) {
      // halo: cells -1 (upper edge -> left state of face 0) and TI / TJ (lower edge -> right
      // state of the last face)
      for (int t = h; t < 2 * TJ + 2 * TI; t += kSwHalo) {
        int p, s, ilo, iup, ivs, out, ovs, side;
        if (t < 2 * TJ) {
          const int r = t >> 1;
          side = t & 1;
          const int c = side ? TI : -1;
          p = (r + H) * PI + c + HX; s = 1;
          ilo = SM::ix + r * (TI + 3) + c + 1; iup = ilo + 1; ivs = SM::ix_vs;
          out = side ? SM::qhx + r : SM::qlx + r * (TI + 1);
          ovs = side ? SM::qhx_vs : SM::qlx_vs;
        } else {
          const int tt = t - 2 * TJ, c = tt % TI;
          side = tt / TI;
          const int r = side ? TJ : -1;
          p = (r + H) * PI + c + HX; s = PI;
          ilo = SM::iy + (r + 1) * TI + c; iup = ilo + TI; ivs = SM::iy_vs;
          out = side ? SM::qhy + c : SM::qly + c;
          ovs = side ? SM::qhy_vs : SM::qly_vs;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          double ql, qr;
          sw_recon<RC>(R0 + v * kSwTile + p, s, PPM ? sm[ilo + v * ivs] : 0.0,
                       PPM ? sm[iup + v * ivs] : 0.0, ql, qr);
          sm[out + v * ovs] = side ? qr : ql;
        }
      }
    }
    __syncthreads();
    // plane k is dead: refill its slot with plane k+4
    if (tid == 0) {
      if (st + kSwRing < nplanes) issue(st + kSwRing);
      if (st + kSwRing + kAhead < nplanes) prefetch(st + kSwRing + kAhead);
    }

    // =========================== P3: Riemann solves ============================================
#ifdef AB200_FAST_MATH
    double z0 = 0.0, z1 = 1.0;
    if (INP && is_main) {  // issued here so the table latency hides behind the Riemann solves
      z0 = x3f[k];
      z1 = x3f[k + 1];
    }
#endif
    if (is_main) {
      if (INP) {
        // x1 (face ci, recon order == pack order) and x2 (face cj, recon order rho, v2, v3, v1,
        // P, sie): both left states are loaded first and both results stored last, so the two
        // independent solves form one straight-line block the scheduler can interleave
        double wlx[NV], wly[NV], wry[NV], ox_[8], oy_[8];
        const int o = cj * TI + ci;
#pragma unroll
        for (int v = 0; v < NV; ++v) wlx[v] = sm[SM::qlx + v * SM::qlx_vs + cj * (TI + 1) + ci];
        wly[0] = sm[SM::qly + 0 * SM::qly_vs + o]; wly[1] = sm[SM::qly + 2 * SM::qly_vs + o];
        wly[2] = sm[SM::qly + 3 * SM::qly_vs + o]; wly[3] = sm[SM::qly + 1 * SM::qly_vs + o];
        wry[0] = qry[0]; wry[1] = qry[2]; wry[2] = qry[3]; wry[3] = qry[1];
        if (gas) {
          wly[4] = sm[SM::qly + 4 * SM::qly_vs + o]; wly[5] = sm[SM::qly + 5 * SM::qly_vs + o];
          wry[4] = qry[4]; wry[5] = qry[5];
        }
        Riemann<RS, FLUID>::solve(eos, wlx, qrx, ox_);
        Riemann<RS, FLUID>::solve(eos, wly, wry, oy_);
#pragma unroll
        for (int m = 0; m < NF; ++m) sm[SM::fx + m * SM::fx_vs + cj * (TI + 1) + ci] = ox_[m];
        sm[SM::fy + 0 * SM::fy_vs + o] = oy_[0];
        sm[SM::fy + 2 * SM::fy_vs + o] = oy_[1];
        sm[SM::fy + 3 * SM::fy_vs + o] = oy_[2];
        sm[SM::fy + 1 * SM::fy_vs + o] = oy_[3];
        if (gas) {
#pragma unroll
          for (int m = 4; m < 8; ++m) sm[SM::fy + m * SM::fy_vs + o] = oy_[m];
        }
      }
    } else if (INP) {
      // halo: the last face of every row (x1, face TI) and of every column (x2, face TJ)
      for (int t = h; t < TJ + TI; t += kSwHalo) {
        const bool isx = t < TJ;
        const int d = isx ? 0 : 1;
        int wlb, wlvs, wrb, wrvs, fb, fvs;
        if (isx) {
          wlb = SM::qlx + t * (TI + 1) + TI; wlvs = SM::qlx_vs;
          wrb = SM::qhx + t; wrvs = SM::qhx_vs;
          fb = SM::fx + t * (TI + 1) + TI; fvs = SM::fx_vs;
        } else {
          const int c = t - TJ;
          wlb = SM::qly + TJ * TI + c; wlvs = SM::qly_vs;
          wrb = SM::qhy + c; wrvs = SM::qhy_vs;
          fb = SM::fy + TJ * TI + c; fvs = SM::fy_vs;
        }
        const int v1 = 1 + d, v2 = 1 + (d + 1) % 3, v3 = 1 + (d + 2) % 3;
        double wl[NV], wr[NV], out[8];
        wl[0] = sm[wlb]; wl[1] = sm[wlb + v1 * wlvs]; wl[2] = sm[wlb + v2 * wlvs];
        wl[3] = sm[wlb + v3 * wlvs];
        wr[0] = sm[wrb]; wr[1] = sm[wrb + v1 * wrvs]; wr[2] = sm[wrb + v2 * wrvs];
        wr[3] = sm[wrb + v3 * wrvs];
        if (gas) {
          wl[4] = sm[wlb + 4 * wlvs]; wl[5] = sm[wlb + 5 * wlvs];
          wr[4] = sm[wrb + 4 * wrvs]; wr[5] = sm[wrb + 5 * wrvs];
        }
        Riemann<RS, FLUID>::solve(eos, wl, wr, out);
        sm[fb] = out[0];
        sm[fb + v1 * fvs] = out[1];
        sm[fb + v2 * fvs] = out[2];
        sm[fb + v3 * fvs] = out[3];
        if (gas) {
#pragma unroll
          for (int m = 4; m < 8; ++m) sm[fb + m * fvs] = out[m];
        }
      }
    }
    if (INP) __syncthreads();

    // =========================== P4: update the zone ===========================================
    if (INP && is_main) {
      cp_async_wait_all();
      const int ox = cj * (TI + 1) + ci, oy = cj * TI + ci;
      double u[6], Fz[NF];
#ifdef AB200_FAST_MATH
      double ub[6];
#endif
#pragma unroll
      for (int m = 0; m < NF; ++m) Fz[m] = sm[SM::fzs + m * kSwMain + tid];
#ifdef AB200_FAST_MATH
      const double rz = drcp(z1 - z0);
#else
      Coords<AB200_CARTESIAN> cc(g, b, k, jj, ii);
      const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
      const double ax2[2] = {cc.area2(0), cc.area2(1)};
      const double ax3[2] = {cc.area3(), cc.area3()};
      const double vol = cc.volume();
#endif
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const double xl = sm[SM::fx + m * SM::fx_vs + ox], xh = sm[SM::fx + m * SM::fx_vs + ox + 1];
        const double yl = sm[SM::fy + m * SM::fy_vs + oy], yh = sm[SM::fy + m * SM::fy_vs + oy + TI];
        double base;
        if (MODE == 0) {
          base = sm[SM::ust + m * kSwMain + tid];
          if (active) __stcg(s_ptr[NV + m] + off, base);  // u1 <- u0
        } else if (MODE == 1) {
          base = a.gam1 * sm[SM::ust + (NV + m) * kSwMain + tid];
        } else {
          base = a.gam0 * sm[SM::ust + m * kSwMain + tid] +
                 a.gam1 * sm[SM::ust + (NV + m) * kSwMain + tid];
        }
        // ApplyUpdate (artemis_integrator.hpp:95-106)
#ifdef AB200_FAST_MATH
        u[m] = (xl - xh) * rx + (yl - yh) * ry + (Fz[m] - FzN[m]) * rz;
        ub[m] = base;
#else
        double divf = (ax1[0] * xl - ax1[1] * xh);
        divf += (ax2[0] * yl - ax2[1] * yh);
        divf += (ax3[0] * Fz[m] - ax3[1] * FzN[m]);
        u[m] = base + divf * bdt / vol;
#endif
      }
      if (gas) {  // FluxSource (fluid_fluxes.hpp:365-392), direction by direction
        const double pxl = sm[SM::fx + 6 * SM::fx_vs + ox], pxh = sm[SM::fx + 6 * SM::fx_vs + ox + 1];
        const double vxl = sm[SM::fx + 7 * SM::fx_vs + ox], vxh = sm[SM::fx + 7 * SM::fx_vs + ox + 1];
        const double pyl = sm[SM::fy + 6 * SM::fy_vs + oy], pyh = sm[SM::fy + 6 * SM::fy_vs + oy + TI];
        const double vyl = sm[SM::fy + 7 * SM::fy_vs + oy], vyh = sm[SM::fy + 7 * SM::fy_vs + oy + TI];
#ifdef AB200_FAST_MATH
        u[1] += rx * (pxl - pxh);
        u[5] -= rx * 0.5 * (pxl + pxh) * (vxh - vxl);
        u[2] += ry * (pyl - pyh);
        u[5] -= ry * 0.5 * (pyl + pyh) * (vyh - vyl);
        u[3] += rz * (Fz[6] - FzN[6]);
        u[5] -= rz * 0.5 * (Fz[6] + FzN[6]) * (FzN[7] - Fz[7]);
#pragma unroll
        for (int m = 0; m < NV; ++m) u[m] = fma(bdt, u[m], ub[m]);
#else
        const double dx1 = cc.x1[1] - cc.x1[0], dx2 = cc.x2[1] - cc.x2[0],
                     dx3 = cc.x3[1] - cc.x3[0];
        u[1] += bdt / dx1 * (pxl - pxh);
        u[5] -= bdt / vol * 0.5 * (pxl + pxh) * (ax1[1] * vxh - ax1[0] * vxl);
        u[2] += bdt / dx2 * (pyl - pyh);
        u[5] -= bdt / vol * 0.5 * (pyl + pyh) * (ax2[1] * vyh - ax2[0] * vyl);
        u[3] += bdt / dx3 * (Fz[6] - FzN[6]);
        u[5] -= bdt / vol * 0.5 * (Fz[6] + FzN[6]) * (ax3[1] * FzN[7] - ax3[0] * Fz[7]);
#endif
      }
#ifdef AB200_FAST_MATH
      if (!gas) {
#pragma unroll
        for (int m = 0; m < NV; ++m) u[m] = fma(bdt, u[m], ub[m]);
      }
#endif
      if (active) {
        const double hx[3] = {1.0, 1.0, 1.0};
        if (gas)  // SetAuxillaryFields (fill_derived.cpp:55-72)
          u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor,
                              f.de_switch);
        // ConsToPrim (fill_derived.cpp:129-164)
        double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
#ifdef AB200_FAST_MATH
        const double rwd = drcp(w_d);
        const double v1 = u[1] * rwd, v2 = u[2] * rwd, v3 = u[3] * rwd;
#else
        const double v1 = u[1] / (w_d * hx[0]), v2 = u[2] / (w_d * hx[1]),
                     v3 = u[3] / (w_d * hx[2]);
#endif
        // PrimToCons on the just-computed primitives (fill_derived.cpp:217-274)
        w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
        double **pp = s_ptr + 2 * NV, **pu = s_ptr;
        __stcg(pp[0] + off, w_d);
        __stcg(pp[1] + off, v1);
        __stcg(pp[2] + off, v2);
        __stcg(pp[3] + off, v3);
        __stcg(pu[0] + off, w_d);
        __stcg(pu[1] + off, w_d * v1 * hx[0]);
        __stcg(pu[2] + off, w_d * v2 * hx[1]);
        __stcg(pu[3] + off, w_d * v3 * hx[2]);
        double w_s = 0.0;
        if (gas) {
#ifdef AB200_FAST_MATH
          w_s = u[5] * rwd;
#else
          w_s = u[5] / ((u[0] > f.dfloor) ? u[0] : f.dfloor);
#endif
          w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          const double u_u = w_s * w_d;
          __stcg(pp[5] + off, w_s);
          __stcg(pp[4] + off, dmax(0.0, f.gm1 * w_d * w_s));
          __stcg(pu[5] + off, u_u);
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
          __stcg(pu[4] + off, u_u + ke);
        }
        if (a.dt_min) {  // EstimateTimestepMesh folded in (src/gas/gas.cpp:411-433)
#ifdef AB200_FAST_MATH
          double cs = 0.0;
          if (gas) cs = dsqrt(ddiv(dmax(0.0, (f.gm1 + 1) * f.gm1 * w_d * w_s), w_d));
          tden = dmax(tden, (fabs(v1) + cs) * rx + (fabs(v2) + cs) * ry + (fabs(v3) + cs) * rz);
#else
          Coords<AB200_CARTESIAN> cd(g, b, k, j, i);
          const double vel[3] = {v1, v2, v3};
          tmin = dmin(tmin, cell_dt<AB200_CARTESIAN, FLUID>(g, f, cd, w_d, vel, w_s));
#endif
        }
      }
    }
    // ---- rotate the x3 march state ---------------------------------------------------------------
    if (is_main) {
      if (PPM) {
#pragma unroll
        for (int v = 0; v < NV; ++v) Ilo[v] = Inew[v];
      }
      if (ZR) {
#pragma unroll
        for (int m = 0; m < NF; ++m) sm[SM::fzs + m * kSwMain + tid] = FzN[m];
      }
    }
    offk += plane;
  };

  // planes ks-3 .. ks-1 arrive before the first step; plane ks is waited for inside it
  for (int p = 0; p < kSwRing - 1; ++p) mbar_wait(bar + p, 0);
  using T_ = std::true_type;
  using F_ = std::false_type;
  step(F_{}, F_{}, 0);
  step(F_{}, F_{}, 1);
  step(F_{}, T_{}, 2);
  for (int st = H; st < nsteps; ++st) step(T_{}, T_{}, st);

  if (a.dt_min) {  // warp-shuffle min, one atomic per warp
#ifdef AB200_FAST_MATH
    if (tden > 0.0) tmin = drcp(tden);
#endif
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmin = dmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    if ((tid & 31) == 0 && is_main)
      atomicMin(a.dt_min, (unsigned long long)__double_as_longlong(tmin));
  }
}

}  // namespace ab200
