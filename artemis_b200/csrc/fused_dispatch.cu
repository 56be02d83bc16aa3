// fused_dispatch.cu -- runtime coordinate-system dispatch for the fused directional passes.
#include <cstdlib>

#include "fused.cuh"

namespace ab200 {

template <int GEOM>
int launch_fused_geom(ab200_ctx *c, int fluid, const FusedArgs &a, int pcm);
template <> int launch_fused_geom<0>(ab200_ctx *, int, const FusedArgs &, int);
template <> int launch_fused_geom<1>(ab200_ctx *, int, const FusedArgs &, int);
template <> int launch_fused_geom<2>(ab200_ctx *, int, const FusedArgs &, int);
template <> int launch_fused_geom<3>(ab200_ctx *, int, const FusedArgs &, int);
template <> int launch_fused_geom<4>(ab200_ctx *, int, const FusedArgs &, int);
template <> int launch_fused_geom<5>(ab200_ctx *, int, const FusedArgs &, int);

bool fused_folds_dt(const ab200_ctx *c) {
  return c->g.ndim >= 2 && !getenv("AB200_NO_MARCH");
}

// block subsets exist for the streaming x1 pass + the marching x2 / x3 passes only
bool fused_supports_subsets(const ab200_ctx *c) {
  return c->g.ndim >= 2 && !getenv("AB200_NO_XCHUNK") && !getenv("AB200_NO_MARCH");
}

int launch_fused_stage(ab200_ctx *c, int fluid, double gam0, double gam1, double beta, double dt,
                       int pcm, int stage1_copy, int use_device_dt, unsigned long long *dt_min,
                       int defer_c2p, int subset, int tap) {
  FusedArgs a{};
  a.dt_min = dt_min;
  a.defer_c2p = defer_c2p;
  a.tap = tap;
  if (subset) {  // 1 = surface blocks (touch a face owned by another rank), 2 = the rest
    a.blist = c->d_blist[subset - 1];
    a.nbl = c->n_blist[subset - 1];
    a.subset = subset;
    if (a.nbl == 0 && subset == 2) return AB200_OK;
  }
  a.gam0 = gam0; a.gam1 = gam1; a.beta = beta; a.dt = dt; a.omf = c->omf;
  a.dt_dev = use_device_dt ? c->d_time : nullptr;
  a.copy_u1 = stage1_copy;
  switch (c->g.geom) {
  case 0: return launch_fused_geom<0>(c, fluid, a, pcm);
  case 1: return launch_fused_geom<1>(c, fluid, a, pcm);
  case 2: return launch_fused_geom<2>(c, fluid, a, pcm);
  case 3: return launch_fused_geom<3>(c, fluid, a, pcm);
  case 4: return launch_fused_geom<4>(c, fluid, a, pcm);
  case 5: return launch_fused_geom<5>(c, fluid, a, pcm);
  }
  set_error("Coordinate type not recognized!");
  return AB200_EINVAL;
}

}  // namespace ab200
