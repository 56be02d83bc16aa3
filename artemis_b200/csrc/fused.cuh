// fused.cuh -- the fused fast path: ONE kernel per direction that does
//   reconstruct -> Riemann -> flux difference -> conserved update (+ coordinate sources)
// without ever materialising a flux array, and whose LAST pass also folds in
// SetAuxillaryFields + ConsToPrim + the interior part of PrimToCons.
//
// It replaces, per stage, K1-K7 and the interior of K12, i.e. the
// task chain src/artemis_driver.cpp:184-255 (+ interior of :261).
//
// Work decomposition ("pencil tiles"): a CTA owns NP pencils of ONE MeshBlock along the pass
// direction (blockIdx.y = block).  Every cell of a pencil -- interior cells 0..L-1 plus one
// halo cell at each end -- gets one thread.  Data flow inside the CTA:
//   stage-in (TMA=true) one elected thread issues cp.async.bulk.tensor loads that bring the
//            pencil tile (all halo cells included) of every primitive variable, of u0 and,
//            when the stage needs it, of u1 into shared memory; completion is signalled on an
//            mbarrier.  All loads of a tile are in flight at once and 2 CTAs per SM overlap
//            one tile's load latency with another tile's arithmetic.
//   phase A  each thread reconstructs ITS cell once (all variables) from the staged tile;
//            keeps the lower-face state qr in registers, publishes the upper-face state ql;
//   phase B  threads of cells 0..L solve the Riemann problem at their LOWER face once
//            (ql from shared memory, qr from registers) and publish the 8 face quantities;
//   phase C  threads of cells 0..L-1 pick up the UPPER face from shared memory and update.
// Every reconstruction and every Riemann solve is done exactly once; the only redundancy is
// 2 idle lanes per pencil in phase C.  Stores are coalesced along i in all three directions
// (thread->item order is i-fastest).  TMA=false is the fallback that reads the stencil
// straight from global memory through L1 (odd ni, unaligned arrays, no driver entry point).
#pragma once
#include <cuda.h>

#include "tasks.cuh"

namespace ab200 {

constexpr int kFusedMaxThreads = 512;  // L1-staged fallback: 1 CTA / SM
constexpr int kTmaMaxThreads = 288;    // TMA-staged: 2 CTAs / SM (<= 113 registers)

struct FusedArgs {
  double gam0, gam1, beta, dt, omf;
  const double *dt_dev;  // if non-null: dt = *dt_dev
  int first, last, copy_u1;
  int defer_c2p;  // AB200_STAGE_DEFER_C2P: no pass is the LAST one (source terms follow)
  int tap;        // AB200_STAGE_TAP_DFLUX: every pass also stores its MASS flux into f.dflux
  // block subset of this launch (AB200_STAGE_SURFACE / _INTERIOR): the launch covers blocks
  // blist[0 .. nbl-1]; nullptr = all g.nb blocks
  const int *blist;
  int nbl;
  int subset;  // 0 all, 1 surface call, 2 interior call (host-side bookkeeping)
  int np;        // pencils per CTA
  int npencils;  // pencils per MeshBlock
  int tiles_per_row;    // TMA: tiles along the transverse index that is tiled
  int tiles_per_block;  // TMA: tiles per MeshBlock
  int nwork;            // TMA: nb * tiles_per_block * nspecies
  const CUtensorMap *maps;  // TMA: [3 kinds][nb*nvar] for this direction
  // last pass of the last stage: min over zones of the CFL timestep (positive doubles order
  // like unsigned integers), folded into the kernel that writes the new primitives
  unsigned long long *dt_min;
};

AB_D void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
AB_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
AB_D void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
AB_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
AB_D void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AB_DONE;\n"
      "bra AB_WAIT;\n"
      "AB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
AB_D void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int GEOM, int FLUID, int RS, int RC, int DIR, bool TMA>
__global__ void __launch_bounds__(TMA ? kTmaMaxThreads : kFusedMaxThreads, TMA ? 2 : 1)
k_fused_pass(GridDev g, FluidDev f, FusedArgs a) {
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr int NV = gas ? 6 : 4;   // reconstructed variables per species
  constexpr int NF = gas ? 8 : 4;   // face quantities per species
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int L = DIR == 1 ? nir : (DIR == 2 ? njr : nkr);
  const int nL = DIR == 1 ? g.ni : (DIR == 2 ? g.nj : g.nk);  // allocated extent along DIR
  const int s0 = DIR == 1 ? g.is : (DIR == 2 ? g.js : g.ks);
  const int NP = a.np;
  const int nslots = NP * (L + 1);
  const bool need_u1 = a.first && !a.copy_u1;
  const int S = f.S;
  const int nvar = f.nvar;
  // ---- shared-memory carve-up -----------------------------------------------------------
  // TMA: [2 mbarriers | 2 stages x {prim, u0, (u1)} tiles | ql | fx | pointers]
  const int tile_stride = TMA ? ((NP * nL * 8 + 127) / 128) * 16 : 0;  // doubles per variable
  const int stage_stride = (need_u1 ? 3 : 2) * NV * tile_stride;        // doubles per stage
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  double *s_tiles = reinterpret_cast<double *>(smem_raw + (TMA ? 128 : 0));
  double *s_ql = s_tiles + 2 * stage_stride;  // [NV][nslots]
  double *s_fx = s_ql + NV * nslots;          // [NF][nslots]
  // CTA-uniform array base pointers of the current block: prim | u0 | u1, nvar each
  double **s_ptr = reinterpret_cast<double **>(s_fx + NF * nslots);

  // ---- thread -> (pencil p, cell c): fixed for the lifetime of the CTA ----------------------
  const int t = threadIdx.x;
  int p, c;
  if (DIR == 1) { p = t / (L + 2); c = t % (L + 2) - 1; }
  else { p = t % NP; c = t / NP - 1; }
  const bool in_tile = (t < NP * (L + 2));
  const int st = DIR == 1 ? 1 : (DIR == 2 ? g.ni : g.ni * g.nj);
  // position of this cell inside a staged tile, and the stride along DIR there
  const int tctr = DIR == 1 ? p * nL + (s0 + c) : (s0 + c) * NP + p;
  const int tst = DIR == 1 ? 1 : NP;
  // slot of face `fc` (0..L) of pencil p; consecutive lanes -> consecutive doubles
  auto slot = [&](int fc) { return DIR == 1 ? p * (L + 1) + fc : fc * NP + p; };

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};

  // ---- work items ---------------------------------------------------------------------------
  // TMA: persistent CTAs; item w = (tile, species), tiles enumerated block-major.
  // fallback: one tile per CTA (blockIdx), items = species.
  const int nwork = TMA ? a.nwork : S;
  const int wstep = TMA ? (int)gridDim.x : 1;
  auto tile_of = [&](int w, int &b, int &n, int &tr, int &row) {
    if (TMA) {
      n = w % S;
      const int tile = w / S;
      b = tile / a.tiles_per_block;
      const int tb = tile - b * a.tiles_per_block;
      row = tb / a.tiles_per_row;
      tr = tb - row * a.tiles_per_row;
    } else {
      n = w; b = blockIdx.y; tr = 0; row = 0;
    }
  };
  // elected thread: launch the TMA loads of work item w into pipeline stage `stg`
  auto issue = [&](int w, int stg) {
    int b, n, tr, row;
    tile_of(w, b, n, tr, row);
    int bc0, bc1, bc2;  // box start (i, j, k); tiles never straddle a transverse row
    if (DIR == 1) { bc0 = 0; bc1 = g.js + tr * NP; bc2 = g.ks + row; }
    if (DIR == 2) { bc0 = g.is + tr * NP; bc1 = 0; bc2 = g.ks + row; }
    if (DIR == 3) { bc0 = g.is + tr * NP; bc1 = g.js + row; bc2 = 0; }
    const int idx[6] = {n, S + 3 * n + (DIR - 1), S + 3 * n + ((DIR - 1) + 1) % 3,
                        S + 3 * n + ((DIR - 1) + 2) % 3, 4 * S + n, 5 * S + n};
    const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};
    const uint32_t bytes = (uint32_t)(NP * nL * 8) * NV * (need_u1 ? 3 : 2);
    mbar_expect_tx(bar + stg, bytes);
    const CUtensorMap *mp = a.maps + (size_t)b * nvar;
    const size_t kind = (size_t)g.nb * nvar;
    double *dst = s_tiles + stg * stage_stride;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      tma_load_3d(dst + v * tile_stride, mp + idx[v], bar + stg, bc0, bc1, bc2);
#pragma unroll
    for (int m = 0; m < NV; ++m)
      tma_load_3d(dst + (NV + m) * tile_stride, mp + kind + ci[m], bar + stg, bc0, bc1, bc2);
    if (need_u1) {
#pragma unroll
      for (int m = 0; m < NV; ++m)
        tma_load_3d(dst + (2 * NV + m) * tile_stride, mp + 2 * kind + ci[m], bar + stg, bc0, bc1,
                    bc2);
    }
  };

  int w = TMA ? (int)blockIdx.x : 0;
  if (TMA) {
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_init(bar + 1, 1);
    }
    __syncthreads();
    if (t == 0 && w < nwork) issue(w, 0);
  }
  int cur_b = -1;
  for (int it = 0; w < nwork; w += wstep, ++it) {
    int b, n, tr, row;
    tile_of(w, b, n, tr, row);
    const int stg = it & 1;
    if (TMA && t == 0 && w + wstep < nwork) issue(w + wstep, stg ^ 1);  // prefetch next item
    if (b != cur_b) {  // (re)load the block's array base pointers (CTA-uniform branch)
      cur_b = b;
      for (int q = threadIdx.x; q < 3 * nvar; q += blockDim.x) {
        const int kind = q / nvar, e = q - kind * nvar;
        double *const *tab = kind == 0 ? f.prim : (kind == 1 ? f.u0 : f.u1);
        s_ptr[q] = tab ? tab[(size_t)b * nvar + e] : nullptr;
      }
      __syncthreads();
    }
    // ---- cell indices of this thread for this item ------------------------------------------
    int k = g.ks, j = g.js, i = g.is;
    bool active = in_tile;
    if (TMA) {
      if (DIR == 1) { j = g.js + tr * NP + p; k = g.ks + row; i = g.is + c; active = active && j <= g.je; }
      if (DIR == 2) { i = g.is + tr * NP + p; k = g.ks + row; j = g.js + c; active = active && i <= g.ie; }
      if (DIR == 3) { i = g.is + tr * NP + p; j = g.js + row; k = g.ks + c; active = active && i <= g.ie; }
    } else {
      const int pid = blockIdx.x * NP + p;
      active = active && (pid < a.npencils);
      if (active) {
        if (DIR == 1) { j = pid % njr + g.js; k = pid / njr + g.ks; i = g.is + c; }
        if (DIR == 2) { i = pid % nir + g.is; k = pid / nir + g.ks; j = g.js + c; }
        if (DIR == 3) { i = pid % nir + g.is; j = pid / nir + g.js; k = g.ks + c; }
      }
    }
    const int off = (k * g.nj + j) * g.ni + i;
    const bool interior = active && c >= 0 && c < L;
    const double *s_prim = s_tiles + stg * stage_stride;
    const double *s_u0 = s_prim + NV * tile_stride;
    const double *s_u1 = s_u0 + NV * tile_stride;
    // PLM_G geometry of this cell along DIR
    double gx[6] = {0, 0, 0, 0, 0, 0};
    if (!CART && RC == AB200_PLM && active)
      plmg_geom<GEOM, DIR>(g, b, k, j, i, gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);

    int idx[6];
    idx[0] = n;
    idx[1] = S + 3 * n + (DIR - 1);
    idx[2] = S + 3 * n + ((DIR - 1) + 1) % 3;
    idx[3] = S + 3 * n + ((DIR - 1) + 2) % 3;
    idx[4] = 4 * S + n;
    idx[5] = 5 * S + n;
    // conserved values in pack order: rho, m1, m2, m3, (E, u)
    const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};

    // ---- stage-in: wait for this item's TMA loads --------------------------------------------
    if (TMA) mbar_wait(bar + stg, (uint32_t)((it >> 1) & 1));

    // ---- phase A: reconstruct this cell ------------------------------------------------
    double qr[NV], wc0 = 0.0, wcv[3] = {0.0, 0.0, 0.0};
    if (active) {
      if (!TMA && interior) {  // warm L1 for phase C
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          prefetch_l1(s_ptr[nvar + ci[m]] + off);
          if (need_u1) prefetch_l1(s_ptr[2 * nvar + ci[m]] + off);
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double *q = TMA ? (s_prim + v * tile_stride + tctr) : (s_ptr[idx[v]] + off);
        double ql;
        recon_cell<RC, CART>(q, (ptrdiff_t)(TMA ? tst : st), ql, qr[v], gx[0], gx[1], gx[2],
                             gx[3], gx[4], gx[5]);
        if (!CART) {
          if (v == 0) wc0 = q[0];
          if (v >= 1 && v <= 3) wcv[v - 1] = q[0];
        }
        if (c < L) s_ql[v * nslots + slot(c + 1)] = ql;
      }
    }
    __syncthreads();

    // ---- phase B: Riemann solve at the lower face of cells 0..L --------------------------
    double lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active && c >= 0) {
      double wl[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) wl[v] = s_ql[v * nslots + slot(c)];
      Riemann<RS, FLUID>::solve(eos, wl, qr, lo);
      if (!CART) {  // ScaleMomentumFlux, fluid_fluxes.hpp:32-70
        Coords<GEOM> cf(g, b, k, j, i);
        double hs[3];
        cf.template face_scale<DIR>(hs);
#pragma unroll
        for (int m = 1; m <= 3; ++m) lo[m] *= hs[(DIR - 1 + (m - 1)) % 3];
      }
#pragma unroll
      for (int m = 0; m < NF; ++m) s_fx[m * nslots + slot(c)] = lo[m];
      if (!CART && a.tap) f.dflux[DIR - 1][(size_t)b * S + n][off] = lo[0];
    }
    __syncthreads();

    // ---- phase C: flux difference + sources + (last pass) C2P/P2C ------------------------
    if (interior) {
      double hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NF; ++m) hi[m] = s_fx[m * nslots + slot(c + 1)];
      Coords<GEOM> cc(g, b, k, j, i);
      double a0, a1;
      if (DIR == 1) { a0 = cc.area1(cc.x1[0]); a1 = cc.area1(cc.x1[1]); }
      else if (DIR == 2) { a0 = cc.area2(0); a1 = cc.area2(1); }
      else { a0 = cc.area3(); a1 = cc.area3(); }
      const double vol = cc.volume();
      double u[6];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const double v0 = TMA ? s_u0[m * tile_stride + tctr]
                              : ((a.first && !a.copy_u1 && a.gam0 == 0.0) ? 0.0
                                                                          : s_ptr[nvar + ci[m]][off]);
        if (a.first) {
          if (a.copy_u1) {  // stage 1 with DeepCopyConservedData folded in: u1 <- u0
            s_ptr[2 * nvar + ci[m]][off] = v0;
            u[m] = v0;
          } else {
            const double v1 = TMA ? s_u1[m * tile_stride + tctr] : s_ptr[2 * nvar + ci[m]][off];
            u[m] = (a.gam0 == 0.0) ? a.gam1 * v1 : a.gam0 * v0 + a.gam1 * v1;
          }
        } else {
          u[m] = v0;
        }
      }
      // ApplyUpdate, direction DIR (artemis_integrator.hpp:95-106): += divf * beta_dt / vol
#ifdef AB200_FAST_MATH
      const double wv = ddiv(bdt, vol);
#define AB_UPD(x) ((x) * wv)
#else
#define AB_UPD(x) ((x) * bdt / vol)
#endif
      u[0] += AB_UPD(a0 * lo[0] - a1 * hi[0]);
#pragma unroll
      for (int m = 1; m <= 3; ++m) {
        const int comp = (DIR - 1 + (m - 1)) % 3;
        u[1 + comp] += AB_UPD(a0 * lo[m] - a1 * hi[m]);
      }
      if (gas) {
        u[4] += AB_UPD(a0 * lo[4] - a1 * hi[4]);
        u[5] += AB_UPD(a0 * lo[5] - a1 * hi[5]);
        // FluxSource, direction DIR (fluid_fluxes.hpp:365-392)
        const double dxd = DIR == 1 ? cc.x1[1] - cc.x1[0]
                                    : (DIR == 2 ? cc.x2[1] - cc.x2[0] : cc.x3[1] - cc.x3[0]);
        u[1 + (DIR - 1)] += ddiv(bdt, dxd) * (lo[6] - hi[6]);
#ifdef AB200_FAST_MATH
        u[5] -= wv * 0.5 * (lo[6] + hi[6]) * (a1 * hi[7] - a0 * lo[7]);
#else
        u[5] -= bdt / vol * 0.5 * (lo[6] + hi[6]) * (a1 * hi[7] - a0 * lo[7]);
#endif
      }
#undef AB_UPD
      // coordinate source terms (fluid_fluxes.hpp:395-415), added once in the first pass
      if (!CART && a.first) {
        // wcv[] is in permuted order: velocities back to (1,2,3)
        double vel[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) vel[(DIR - 1 + m) % 3] = wcv[m];
        double vf[3];
        cc.rotation_velocity(a.omf, vf);
        const double rdt = wc0 * bdt;
        const double s0q = sqr(vel[0] + vf[0]), s1q = sqr(vel[1] + vf[1]),
                     s2q = sqr(vel[2] + vf[2]);
        if (Coords<GEOM>::x1dep) {
          double dh[3];
          cc.conn1(dh);
          u[1] += rdt * (dh[0] * s0q + dh[1] * s1q + dh[2] * s2q);
        }
        if (Coords<GEOM>::x2dep && g.ndim >= 2) {
          double dh[3];
          cc.conn2(dh);
          u[2] += rdt * (dh[0] * s0q + dh[1] * s1q + dh[2] * s2q);
        }
      }
      if (!a.last) {
#pragma unroll
        for (int m = 0; m < NV; ++m) s_ptr[nvar + ci[m]][off] = u[m];
      } else {
        const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
        if (gas)  // SetAuxillaryFields (fill_derived.cpp:55-72)
          u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor,
                              f.de_switch);
        // ConsToPrim (fill_derived.cpp:129-164)
        double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
#ifdef AB200_FAST_MATH
        const double rwd = drcp(w_d);
        const double v1 = CART ? u[1] * rwd : ddiv(u[1], w_d * hx[0]);
        const double v2 = CART ? u[2] * rwd : ddiv(u[2], w_d * hx[1]);
        const double v3 = CART ? u[3] * rwd : ddiv(u[3], w_d * hx[2]);
#else
        const double v1 = u[1] / (w_d * hx[0]), v2 = u[2] / (w_d * hx[1]),
                     v3 = u[3] / (w_d * hx[2]);
#endif
        // PrimToCons on the just-computed primitives (fill_derived.cpp:217-274)
        w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
        double **pp = s_ptr, **pu = s_ptr + nvar;
        pp[ci[0]][off] = w_d;
        pp[ci[1]][off] = v1;
        pp[ci[2]][off] = v2;
        pp[ci[3]][off] = v3;
        pu[ci[0]][off] = w_d;
        pu[ci[1]][off] = w_d * v1 * hx[0];
        pu[ci[2]][off] = w_d * v2 * hx[1];
        pu[ci[3]][off] = w_d * v3 * hx[2];
        if (gas) {
#ifdef AB200_FAST_MATH
          double w_s = u[5] * rwd;
#else
          double w_s = u[5] / ((u[0] > f.dfloor) ? u[0] : f.dfloor);
#endif
          w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          const double u_u = w_s * w_d;
          pp[ci[5]][off] = w_s;
          pp[ci[4]][off] = dmax(0.0, f.gm1 * w_d * w_s);
          pu[ci[5]][off] = u_u;
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
          pu[ci[4]][off] = u_u + ke;
        }
      }
    }
    // exchange buffers and (TMA) this stage's tiles are free for the next item
    if (w + wstep < nwork) __syncthreads();
  }
}

}  // namespace ab200
