// api.cu -- extern "C" surface of libartemis_b200: context lifetime, binding, utilities.
// Every entry point is declared in include/ab200.h with the reference file:line it replaces.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "ab200_ctx.cuh"

namespace ab200 {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e,
           cudaGetErrorString(e), what, file, line);
  g_err = buf;
  return AB200_ECUDA;
}

static int dev_alloc(ab200_ctx *c, void **p, size_t bytes, std::vector<void *> *own) {
  AB_CUDA(cudaMalloc(p, bytes ? bytes : 8));
  if (own) own->push_back(*p);
  (void)c;
  return AB200_OK;
}

template <typename T>
static int upload(ab200_ctx *c, T **dptr, const T *h, size_t n, std::vector<void *> *own) {
  AB_TRY(dev_alloc(c, (void **)dptr, n * sizeof(T), own));
  AB_CUDA(cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  return AB200_OK;
}

static void free_all(std::vector<void *> &v) {
  for (void *p : v) cudaFree(p);
  v.clear();
}

// ---- host restatement of the centroid formulas feeding the metric tables ----------------
static double h_rface(double a, double b) {  // cylindrical.hpp:52-56
  return 2.0 / 3.0 * (a * a + a * b + b * b) / (a + b);
}
static double h_x1v(int geom, double a, double b) {
  if (geom == AB200_CYLINDRICAL || geom == AB200_AXISYMMETRIC) return h_rface(a, b);
  if (geom == AB200_SPHERICAL1D || geom == AB200_SPHERICAL2D || geom == AB200_SPHERICAL3D) {
    const double dr2 = a * a + b * b;  // spherical.hpp:57-60
    return 0.75 * (a + b) * dr2 / (dr2 + a * b);
  }
  return 0.5 * (a + b);
}
static double h_x2v(int geom, double a, double b) {
  if (geom == AB200_SPHERICAL2D || geom == AB200_SPHERICAL3D) {  // spherical.hpp:61-68
    const double ctm = std::cos(a), ctp = std::cos(b);
    const double dst = std::sin(b) - std::sin(a);
    return (dst - b * ctp + a * ctm) / std::fabs(ctm - ctp);
  }
  return 0.5 * (a + b);
}

// Metric tables of nb blocks of ni x nj x nk cells (xmin / dx: [nb][3]); used for the bound
// (fine) grid and, by refine.cu, for the coarse buffers of the multilevel operators.
int build_geom_tables_for(ab200_ctx *c, GeomTab &t, int geom, int nb, int ni, int nj, int nk,
                          const double *xmin_all, const double *dx_all) {
  std::vector<double> x1f((size_t)nb * (ni + 1)), x2f((size_t)nb * (nj + 1)),
      x3f((size_t)nb * (nk + 1)), x1v((size_t)nb * ni), x2v((size_t)nb * nj),
      x3v((size_t)nb * nk), cosf((size_t)nb * (nj + 1)), sinf((size_t)nb * (nj + 1)),
      sinv((size_t)nb * nj), sinc((size_t)nb * nj), cosv((size_t)nb * nj),
      sin3v((size_t)nb * nk), cos3v((size_t)nb * nk);
  for (int b = 0; b < nb; ++b) {
    const double *xm = xmin_all + 3 * b, *dx = dx_all + 3 * b;
    // P:coordinates/uniform_cartesian.hpp:153-157  Xf(idx) = xmin + idx*dx
    for (int i = 0; i <= ni; ++i) x1f[(size_t)b * (ni + 1) + i] = xm[0] + i * dx[0];
    for (int j = 0; j <= nj; ++j) x2f[(size_t)b * (nj + 1) + j] = xm[1] + j * dx[1];
    for (int k = 0; k <= nk; ++k) x3f[(size_t)b * (nk + 1) + k] = xm[2] + k * dx[2];
    for (int i = 0; i < ni; ++i)
      x1v[(size_t)b * ni + i] =
          h_x1v(geom, x1f[(size_t)b * (ni + 1) + i], x1f[(size_t)b * (ni + 1) + i + 1]);
    for (int j = 0; j <= nj; ++j) {
      cosf[(size_t)b * (nj + 1) + j] = std::cos(x2f[(size_t)b * (nj + 1) + j]);
      sinf[(size_t)b * (nj + 1) + j] = std::sin(x2f[(size_t)b * (nj + 1) + j]);
    }
    for (int j = 0; j < nj; ++j) {
      const double a = x2f[(size_t)b * (nj + 1) + j], bb = x2f[(size_t)b * (nj + 1) + j + 1];
      const double v = h_x2v(geom, a, bb);
      x2v[(size_t)b * nj + j] = v;
      sinv[(size_t)b * nj + j] = std::sin(v);
      cosv[(size_t)b * nj + j] = std::cos(v);
      sinc[(size_t)b * nj + j] = std::sin(0.5 * (a + bb));
    }
    for (int k = 0; k < nk; ++k) {
      const double v = 0.5 * (x3f[(size_t)b * (nk + 1) + k] + x3f[(size_t)b * (nk + 1) + k + 1]);
      x3v[(size_t)b * nk + k] = v;
      sin3v[(size_t)b * nk + k] = std::sin(v);
      cos3v[(size_t)b * nk + k] = std::cos(v);
    }
  }
  double *p;
#define UP(field, vec)                                                                   \
  AB_TRY(upload<double>(c, &p, vec.data(), vec.size(), &c->grid_allocs));                \
  t.field = p;
  UP(x1f, x1f) UP(x2f, x2f) UP(x3f, x3f) UP(x1v, x1v) UP(x2v, x2v) UP(x3v, x3v)
  UP(cosf, cosf) UP(sinf, sinf) UP(sinv, sinv) UP(sinc, sinc) UP(cosv, cosv)
  UP(sin3v, sin3v) UP(cos3v, cos3v)
#undef UP
  return AB200_OK;
}

static int build_geom_tables(ab200_ctx *c, const ab200_grid_desc *gd) {
  return build_geom_tables_for(c, c->g.t, gd->geom, gd->nblocks, gd->ni, gd->nj, gd->nk, gd->xmin,
                               gd->dx);
}

int ensure_scratch(ab200_ctx *c, int fluid, bool need_flux, bool need_u1) {
  FluidHost &f = c->fl[fluid];
  const GridDev &g = c->g;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  const size_t fcells = (size_t)g.fni * g.fnj * g.fnk;
  auto make_table = [&](double *const **slot, int nent, size_t elems) -> int {
    if (*slot) return AB200_OK;
    double *slab;
    AB_TRY(dev_alloc(c, (void **)&slab, sizeof(double) * elems * nent * g.nb, &f.owned_scratch));
    AB_CUDA(cudaMemsetAsync(slab, 0, sizeof(double) * elems * nent * g.nb, c->stream));
    std::vector<double *> tab((size_t)nent * g.nb);
    for (size_t e = 0; e < tab.size(); ++e) tab[e] = slab + e * elems;
    double **dt;
    AB_TRY(upload<double *>(c, &dt, tab.data(), tab.size(), &f.owned_tables));
    *slot = dt;
    return AB200_OK;
  };
  if (need_u1) AB_TRY(make_table(&f.d.u1, f.d.nvar, cells));
  if (need_flux) {
    for (int d = 0; d < g.ndim; ++d) {
      AB_TRY(make_table(&f.d.flux[d], f.d.nvar, cells));
      if (fluid == AB200_GAS) {
        AB_TRY(make_table(&f.d.pflux[d], f.d.S, cells));
        AB_TRY(make_table(&f.d.vface[d], f.d.S, fcells));
      }
    }
  }
  return AB200_OK;
}

int ensure_dflux(ab200_ctx *c, int fluid) {
  FluidHost &f = c->fl[fluid];
  const GridDev &g = c->g;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  for (int d = 0; d < g.ndim; ++d) {
    if (f.d.dflux[d]) continue;
    double *slab;
    AB_TRY(dev_alloc(c, (void **)&slab, sizeof(double) * cells * f.d.S * g.nb, &f.owned_scratch));
    AB_CUDA(cudaMemsetAsync(slab, 0, sizeof(double) * cells * f.d.S * g.nb, c->stream));
    std::vector<double *> tab((size_t)f.d.S * g.nb);
    for (size_t e = 0; e < tab.size(); ++e) tab[e] = slab + e * cells;
    double **dt;
    AB_TRY(upload<double *>(c, &dt, tab.data(), tab.size(), &f.owned_tables));
    f.d.dflux[d] = dt;
  }
  return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_abi_version(void) { return AB200_ABI_VERSION; }
const char *ab200_last_error(void) { return g_err.c_str(); }

int ab200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ab200_create(ab200_ctx **out, int device, void *cuda_stream) {
  AB_REQUIRE(out != nullptr, AB200_EINVAL, "ab200_create: null output pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("ab200_create: no CUDA device available (libartemis_b200 has no CPU fallback)");
    return AB200_ECUDA;
  }
  AB_REQUIRE(device >= 0 && device < n, AB200_EINVAL, "ab200_create: bad device ordinal");
  AB_CUDA(cudaSetDevice(device));
  ab200_ctx *c = new ab200_ctx();
  c->device = device;
  c->stream = (cudaStream_t)cuda_stream;
  // any failure below releases what was allocated so far (ab200_destroy tolerates nulls)
  auto init = [&]() -> int {
    AB_CUDA(cudaMalloc((void **)&c->d_time, 8 * sizeof(double)));
    AB_CUDA(cudaMemset(c->d_time, 0, 8 * sizeof(double)));
    AB_CUDA(cudaMalloc((void **)&c->d_red, 4096 * sizeof(double)));
    AB_CUDA(cudaMallocHost((void **)&c->h_pinned, 16 * sizeof(double)));
    // d_red[3080..3081]: bit pattern of Big<Real>() = DBL_MAX, the seed of the folded CFL
    // reduction (the same start value as k_estimate_dt and the reference's Kokkos::Min)
    const double big[2] = {1.79769313486231570815e+308, 1.79769313486231570815e+308};
    AB_CUDA(cudaMemcpy(c->d_red + 3080, big, sizeof big, cudaMemcpyHostToDevice));
    AB_CUDA(cudaEventCreate(&c->ev0));
    AB_CUDA(cudaEventCreate(&c->ev1));
    cudaDeviceProp prop;
    AB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    return AB200_OK;
  };
  const int rc = init();
  if (rc != AB200_OK) {
    ab200_destroy(c);
    return rc;
  }
  *out = c;
  return AB200_OK;
}

int ab200_destroy(ab200_ctx *c) {
  if (!c) return AB200_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  ab200_comm_destroy(c);
  for (int f = 0; f < 2; ++f) ab200_unbind(c, f);
  free_all(c->grid_allocs);
  free_all(c->host_path_allocs);
  clear_descriptor_cache(c);
  for (int q = 0; q < 2; ++q)
    if (c->d_blist[q]) cudaFree(c->d_blist[q]);
  for (int q = 0; q < 3; ++q)
    if (c->d_dflx[q]) cudaFree(c->d_dflx[q]);
  if (c->d_dcoef) cudaFree(c->d_dcoef);
  if (c->d_time) cudaFree(c->d_time);
  if (c->d_red) cudaFree(c->d_red);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  delete c;
  return AB200_OK;
}

int ab200_synchronize(ab200_ctx *c) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  return AB200_OK;
}

int ab200_set_grid(ab200_ctx *c, const ab200_grid_desc *gd) {
  AB_REQUIRE(c && gd, AB200_EINVAL, "ab200_set_grid: null argument");
  AB_REQUIRE(gd->geom >= 0 && gd->geom <= AB200_AXISYMMETRIC, AB200_EINVAL,
             "Coordinate type not recognized!");
  AB_REQUIRE(gd->ndim >= 1 && gd->ndim <= 3, AB200_EINVAL, "ab200_set_grid: bad ndim");
  AB_REQUIRE(gd->nblocks > 0 && gd->ni > 0 && gd->nj > 0 && gd->nk > 0, AB200_EINVAL,
             "ab200_set_grid: empty partition");
  AB_REQUIRE(gd->xmin && gd->dx, AB200_EINVAL, "ab200_set_grid: null coordinates");
  AB_REQUIRE(gd->is >= 0 && gd->ie < gd->ni && gd->js >= 0 && gd->je < gd->nj && gd->ks >= 0 &&
                 gd->ke < gd->nk && gd->is <= gd->ie && gd->js <= gd->je && gd->ks <= gd->ke,
             AB200_EINVAL, "ab200_set_grid: interior bounds outside the allocation");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  for (int f = 0; f < 2; ++f) ab200_unbind(c, f);
  free_all(c->grid_allocs);
  clear_descriptor_cache(c);  // halo / refinement descriptor lists describe the old mesh
  // the (block, variable) index sits in grid.y / grid.z of several launches
  AB_REQUIRE(gd->nblocks <= 65535, AB200_EINVAL,
             "ab200_set_grid: at most 65535 MeshBlocks per MeshData partition");
  GridDev &g = c->g;
  g.geom = gd->geom; g.ndim = gd->ndim; g.ng = gd->nghost; g.nb = gd->nblocks;
  g.ni = gd->ni; g.nj = gd->nj; g.nk = gd->nk;
  g.is = gd->is; g.ie = gd->ie; g.js = gd->js; g.je = gd->je; g.ks = gd->ks; g.ke = gd->ke;
  g.fni = gd->fni; g.fnj = gd->fnj; g.fnk = gd->fnk;
  c->h_xmin.assign(gd->xmin, gd->xmin + 3 * gd->nblocks);
  c->h_dx.assign(gd->dx, gd->dx + 3 * gd->nblocks);
  AB_TRY(build_geom_tables(c, gd));
  c->coarse_ready = false;
  c->grid_set = true;
  c->topo.set = false;
  c->host_path_ready = false;
  return AB200_OK;
}

int ab200_unbind(ab200_ctx *c, int fluid) {
  AB_REQUIRE(c && (fluid == 0 || fluid == 1), AB200_EINVAL, "ab200_unbind: bad argument");
  FluidHost &f = c->fl[fluid];
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_all(f.owned_tables);
  free_all(f.owned_scratch);
  if (f.ghost_vars) cudaFree(f.ghost_vars);
  if (f.ghost_vdir) cudaFree(f.ghost_vdir);
  release_tma(f);
  release_sweep(f);
  f = FluidHost();
  return AB200_OK;
}

int ab200_bind_pack(ab200_ctx *c, const ab200_fluid_desc *fd, const ab200_pack_desc *pk) {
  AB_REQUIRE(c && fd && pk, AB200_EINVAL, "ab200_bind_pack: null argument");
  AB_REQUIRE(c->grid_set, AB200_ESTATE, "ab200_bind_pack: call ab200_set_grid first");
  AB_REQUIRE(fd->fluid == AB200_GAS || fd->fluid == AB200_DUST, AB200_EINVAL,
             "Fluid type not recognized!");
  AB_REQUIRE(fd->nspecies >= 1, AB200_EINVAL, "ab200_bind_pack: nspecies < 1");
  AB_REQUIRE(fd->recon >= AB200_PCM && fd->recon <= AB200_PPM, AB200_EINVAL,
             "Reconstruction method not recognized!");
  AB_REQUIRE(fd->riemann >= AB200_HLLC && fd->riemann <= AB200_LLF, AB200_EINVAL,
             "Riemann solver not recognized!");
  AB_REQUIRE(!(fd->fluid == AB200_DUST && fd->riemann == AB200_HLLC), AB200_EINVAL,
             "Riemann solver (dust) not recognized.");
  const int need_ng = fd->recon == AB200_PPM ? 3 : (fd->recon == AB200_PLM ? 2 : 1);
  AB_REQUIRE(c->g.ng >= need_ng, AB200_EINVAL,
             fd->recon == AB200_PPM   ? "PPM requires at least 3 ghost cells."
             : fd->recon == AB200_PLM ? "PLM requires at least 2 ghost cells."
                                      : "PCM requires at least 1 ghost cell.");
  AB_REQUIRE(pk->prim && pk->cons0, AB200_EINVAL, "ab200_bind_pack: prim/cons0 tables required");
  AB_CUDA(cudaSetDevice(c->device));
  AB_TRY(ab200_unbind(c, fd->fluid));
  FluidHost &f = c->fl[fd->fluid];
  FluidDev &d = f.d;
  d.fluid = fd->fluid; d.S = fd->nspecies;
  d.nvar = (fd->fluid == AB200_GAS ? 6 : 4) * fd->nspecies;
  d.recon = fd->recon; d.riemann = fd->riemann;
  d.gm1 = fd->gm1; d.dfloor = fd->dfloor; d.siefloor = fd->siefloor;
  d.de_switch = fd->de_switch; d.cfl = fd->cfl;
  d.igm1 = 1.0 / fd->gm1;  // hllc.hpp:76-78
  d.gamma = fd->gm1 + 1.0;
  d.alpha = (d.gamma + 1.0) / (2.0 * d.gamma);
  const size_t nent = (size_t)c->g.nb * d.nvar, sent = (size_t)c->g.nb * d.S;
  auto up = [&](double *const **slot, double *const *src, size_t n) -> int {
    *slot = nullptr;
    if (!src) return AB200_OK;
    for (size_t e = 0; e < n; ++e)
      AB_REQUIRE(src[e] != nullptr, AB200_EINVAL, "ab200_bind_pack: null array in a pointer table");
    double **dt;
    AB_TRY(upload<double *>(c, &dt, (double *const *)src, n, &f.owned_tables));
    *slot = dt;
    return AB200_OK;
  };
  AB_TRY(up(&d.prim, pk->prim, nent));
  AB_TRY(up(&d.u0, pk->cons0, nent));
  AB_TRY(up(&d.u1, pk->cons1, nent));
  for (int dir = 0; dir < 3; ++dir) {
    AB_TRY(up(&d.flux[dir], pk->flux[dir], nent));
    AB_TRY(up(&d.pflux[dir], pk->pflux[dir], sent));
    AB_TRY(up(&d.vface[dir], pk->vface[dir], sent));
  }
  // FillGhost fields: gas prim rho, v, sie (pressure is not exchanged: src/gas/gas.cpp:243-270);
  // dust prim rho, v (src/dust/dust.cpp:200-212)
  std::vector<int> gv, gd;
  const int S = d.S;
  for (int v = 0; v < d.nvar; ++v) {
    if (d.fluid == AB200_GAS && v >= 4 * S && v < 5 * S) continue;
    gv.push_back(v);
    gd.push_back((v >= S && v < 4 * S) ? ((v - S) % 3 + 1) : 0);
  }
  f.n_ghost = (int)gv.size();
  AB_CUDA(cudaMalloc((void **)&f.ghost_vars, gv.size() * sizeof(int)));
  AB_CUDA(cudaMalloc((void **)&f.ghost_vdir, gv.size() * sizeof(int)));
  AB_CUDA(cudaMemcpy(f.ghost_vars, gv.data(), gv.size() * sizeof(int), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(f.ghost_vdir, gd.data(), gd.size() * sizeof(int), cudaMemcpyHostToDevice));
  f.bound = true;
  return AB200_OK;
}

int ab200_set_stage_path(ab200_ctx *c, int path) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_REQUIRE(path >= AB200_PATH_AUTO && path <= AB200_PATH_ROLE_SPLIT, AB200_EINVAL,
             "ab200_set_stage_path: unknown path");
  // the primitives may sit in the alternate set of the single-pass kernel: bring them home
  // before the other path takes over
  if (c->grid_set) {
    AB_CUDA(cudaSetDevice(c->device));
    for (int f = 0; f < 2; ++f) AB_TRY(sync_prim_home(c, f, 0));
  }
  c->stage_path = path;
  return AB200_OK;
}

int ab200_set_halo_stream(ab200_ctx *c, void *cuda_stream) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  c->halo_stream = (cudaStream_t)cuda_stream;
  c->halo_stream_set = cuda_stream != nullptr;
  return AB200_OK;
}

int ab200_get_stage_path(ab200_ctx *c, int fluid, int *path_out) {
  AB_REQUIRE(c && path_out, AB200_EINVAL, "ab200_get_stage_path: null argument");
  AB_REQUIRE(c->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");
  AB_REQUIRE(fluid == 0 || fluid == 1, AB200_EINVAL, "Fluid type not recognized!");
  AB_REQUIRE(c->fl[fluid].bound, AB200_ESTATE, "fluid pack not bound: call ab200_bind_pack");
  AB_CUDA(cudaSetDevice(c->device));
  *path_out = !sweep_eligible(c, fluid) ? AB200_PATH_THREE_PASS
              : sweep_uses_role_split(c, fluid) ? AB200_PATH_ROLE_SPLIT : AB200_PATH_SINGLE_PASS;
  return AB200_OK;
}

int ab200_set_rotating_frame(ab200_ctx *c, double omega) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  c->omf = omega;
  return AB200_OK;
}

#define AB_ENTER(c)                                                                      \
  AB_REQUIRE((c) != nullptr, AB200_EINVAL, "null context");                              \
  AB_REQUIRE((c)->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");         \
  AB_CUDA(cudaSetDevice((c)->device));
#define AB_FLUID(c, fluid)                                                               \
  AB_REQUIRE((fluid) == 0 || (fluid) == 1, AB200_EINVAL, "Fluid type not recognized!");  \
  AB_REQUIRE((c)->fl[fluid].bound, AB200_ESTATE, "fluid pack not bound: call ab200_bind_pack");

// Entry points that read or write the caller's primitive arrays directly first bring the
// current primitives home (no-op unless a ping-pong stage left them in the alternate set).
#define AB_PRIM_HOME(c)                                                                  \
  for (int f__ = 0; f__ < 2; ++f__) AB_TRY(sync_prim_home((c), f__, 0));

int ab200_sync_prim(ab200_ctx *c) {
  AB_ENTER(c)
  AB_PRIM_HOME(c)
  return AB200_OK;
}

int ab200_calculate_fluxes(ab200_ctx *c, int fluid, int pcm) {
  AB_ENTER(c) AB_FLUID(c, fluid)
  AB_PRIM_HOME(c)
  AB_TRY(ensure_scratch(c, fluid, true, false));
  c->fl[fluid].dflux_src = 2;  // the mass fluxes of this stage live in the full flux arrays
  return launch_calculate_fluxes(c, fluid, pcm);
}

int ab200_apply_update(ab200_ctx *c, double gam0, double gam1, double beta_dt) {
  AB_ENTER(c)
  for (int f = 0; f < 2; ++f) {
    if (!c->fl[f].bound) continue;
    AB_REQUIRE(c->fl[f].d.flux[0] != nullptr, AB200_ESTATE,
               "ab200_apply_update: no flux arrays (call ab200_calculate_fluxes first)");
    AB_TRY(ensure_scratch(c, f, false, true));
    AB_TRY(launch_apply_update(c, f, gam0, gam1, beta_dt));
  }
  return AB200_OK;
}

int ab200_flux_source(ab200_ctx *c, int fluid, double dt) {
  AB_ENTER(c) AB_FLUID(c, fluid)
  AB_PRIM_HOME(c)
  if (fluid == AB200_GAS)
    AB_REQUIRE(c->fl[fluid].d.pflux[0] != nullptr, AB200_ESTATE,
               "ab200_flux_source: no interface-pressure arrays (call ab200_calculate_fluxes)");
  return launch_flux_source(c, fluid, dt);
}

int ab200_set_auxillary_fields(ab200_ctx *c) {
  AB_ENTER(c)
  if (!c->fl[AB200_GAS].bound) return AB200_OK;  // fill_derived.cpp:37-38
  return launch_set_aux(c);
}

int ab200_cons_to_prim(ab200_ctx *c) {
  AB_ENTER(c)
  AB_PRIM_HOME(c)
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_cons_to_prim(c, f));
  return AB200_OK;
}

int ab200_prim_to_cons(ab200_ctx *c) {
  AB_ENTER(c)
  AB_PRIM_HOME(c)
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_prim_to_cons(c, f, 0));
  return AB200_OK;
}

int ab200_prim_to_cons_ghosts(ab200_ctx *c) {
  AB_ENTER(c)
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) {
      AB_TRY(launch_prim_to_cons(c, f, 1));
      c->fl[f].ghost_cons_stale = false;
    }
  return AB200_OK;
}

int ab200_set_ghost_cons_lazy(ab200_ctx *c, int lazy) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  c->ghost_cons_lazy = lazy != 0;
  return AB200_OK;
}

int ab200_sync_ghost_cons(ab200_ctx *c) {
  AB_ENTER(c)
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound && c->fl[f].ghost_cons_stale) {
      AB_TRY(launch_prim_to_cons(c, f, 1));
      c->fl[f].ghost_cons_stale = false;
    }
  return AB200_OK;
}

int ab200_deep_copy_conserved(ab200_ctx *c) {
  AB_ENTER(c)
  for (int f = 0; f < 2; ++f) {
    if (!c->fl[f].bound) continue;
    AB_TRY(ensure_scratch(c, f, false, true));
    AB_TRY(launch_deep_copy(c, f));
  }
  return AB200_OK;
}

int ab200_estimate_timestep(ab200_ctx *c, int fluid, double *dt_host) {
  AB_ENTER(c) AB_FLUID(c, fluid)
  AB_REQUIRE(dt_host, AB200_EINVAL, "ab200_estimate_timestep: null output");
  AB_PRIM_HOME(c)
  AB_TRY(launch_estimate_dt(c, fluid, c->d_red + 2048, 0));
  AB_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_red + 2048, sizeof(double), cudaMemcpyDeviceToHost,
                          c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  *dt_host = c->h_pinned[0];
  return AB200_OK;
}

int ab200_estimate_timestep_device(ab200_ctx *c) {
  AB_ENTER(c)
  int first = 1;
  for (int f = 0; f < 2; ++f) {
    if (!c->fl[f].bound) continue;
    AB_TRY(launch_estimate_dt(c, f, c->d_time + 1, first ? 0 : 1));
    first = 0;
  }
  AB_REQUIRE(!first, AB200_ESTATE, "ab200_estimate_timestep_device: no fluid bound");
  return AB200_OK;
}

int ab200_set_global_timestep_device(ab200_ctx *c, double tlim, int advance_time) {
  AB_ENTER(c)
  return launch_set_global_dt(c, tlim, advance_time);
}

double *ab200_dt_device(ab200_ctx *c) { return c ? c->d_time : nullptr; }

int ab200_read_time_state(ab200_ctx *c, double *host4) {
  AB_ENTER(c)
  AB_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_time, 4 * sizeof(double), cudaMemcpyDeviceToHost,
                          c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(host4, c->h_pinned, 4 * sizeof(double));
  return AB200_OK;
}

int ab200_write_time_state(ab200_ctx *c, const double *host4) {
  AB_ENTER(c)
  AB_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(c->h_pinned + 8, host4, 4 * sizeof(double));
  AB_CUDA(cudaMemcpyAsync(c->d_time, c->h_pinned + 8, 4 * sizeof(double),
                          cudaMemcpyHostToDevice, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  return AB200_OK;
}

int ab200_fused_stage(ab200_ctx *c, double gam0, double gam1, double beta, double dt, int pcm,
                      int stage1_copy, int flags) {
  AB_ENTER(c)
  AB_REQUIRE(!stage1_copy || (gam0 == 0.0 && gam1 == 1.0), AB200_EINVAL,
             "ab200_fused_stage: stage1_copy requires gam0 == 0 and gam1 == 1");
  AB_REQUIRE((flags & ~(AB200_STAGE_DEVICE_DT | AB200_STAGE_REDUCE_DT | AB200_STAGE_PINGPONG |
                        AB200_STAGE_DEFER_C2P | AB200_STAGE_SURFACE | AB200_STAGE_INTERIOR |
                        AB200_STAGE_TAP_DFLUX)) == 0,
             AB200_EINVAL, "ab200_fused_stage: unknown flag");
  const int subset = (flags & AB200_STAGE_SURFACE) ? 1 : ((flags & AB200_STAGE_INTERIOR) ? 2 : 0);
  AB_REQUIRE(!((flags & AB200_STAGE_SURFACE) && (flags & AB200_STAGE_INTERIOR)), AB200_EINVAL,
             "ab200_fused_stage: SURFACE and INTERIOR are two separate calls");
  AB_REQUIRE(!subset || (c->topo.set && fused_supports_subsets(c)), AB200_ESTATE,
             "ab200_fused_stage: block subsets need ab200_set_topology and a >= 2-D mesh");
  AB_REQUIRE(!subset || (c->n_blist[0] > 0 && c->n_blist[1] > 0), AB200_ESTATE,
             "ab200_fused_stage: SURFACE / INTERIOR need both block subsets to be non-empty");
  const int use_device_dt = (flags & AB200_STAGE_DEVICE_DT) != 0;
  const int defer = (flags & AB200_STAGE_DEFER_C2P) != 0;
  // deferred C2P: the timestep is estimated by ab200_finish_stage, from the final primitives
  const int reduce_dt = (flags & AB200_STAGE_REDUCE_DT) != 0 && !defer;
  const int pingpong = (flags & AB200_STAGE_PINGPONG) != 0;
  const int tap = (flags & AB200_STAGE_TAP_DFLUX) != 0;
  AB_REQUIRE(!tap || c->g.geom != AB200_CARTESIAN, AB200_EINVAL,
             "ab200_fused_stage: the mass-flux tap exists for the curvilinear systems only "
             "(its reader, RotatingFrameImpl, is never Cartesian: rotating_frame.cpp:67-82)");
  // per-fluid raw minimum (bit pattern of a positive double), reset to a huge finite value
  unsigned long long *slots = reinterpret_cast<unsigned long long *>(c->d_red + 3072);
  // (a SURFACE call opens the reduction, the INTERIOR call that follows closes it)
  if (reduce_dt && subset != 2)
    AB_CUDA(cudaMemcpyAsync(slots, c->d_red + 3080, 2 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToDevice, c->stream));
  int any = 0;
  for (int f = 0; f < 2; ++f) {
    if (!c->fl[f].bound) continue;
    AB_TRY(ensure_scratch(c, f, false, true));
    if (tap) AB_TRY(ensure_dflux(c, f));
    c->fl[f].dflux_src = tap ? 1 : 0;
    bool fold;
    if (!defer && !subset && !tap && sweep_eligible(c, f)) {
      // single-pass stage (sweep.cuh): reads the current primitive set, writes the other one
      fold = reduce_dt;
      AB_TRY(launch_sweep_stage(c, f, gam0, gam1, beta, dt, pcm, stage1_copy, use_device_dt,
                                fold ? slots + f : nullptr));
      if (!pingpong) AB_TRY(sync_prim_home(c, f, 1));  // interior zones back to the caller
    } else {
      AB_TRY(sync_prim_home(c, f, 0));
      fold = reduce_dt && fused_folds_dt(c);
      AB_TRY(launch_fused_stage(c, f, gam0, gam1, beta, dt, pcm, stage1_copy, use_device_dt,
                                fold ? slots + f : nullptr, defer, subset, tap));
    }
    if (reduce_dt && subset != 1) {  // new_dt = min over fluids of cfl * min dt  (EstimateTimestepMesh)
      if (fold) {
        AB_TRY(launch_finish_dt(c, reinterpret_cast<const double *>(slots + f), 1, c->fl[f].d.cfl,
                                c->d_time + 1, any));
        // Gas::EstimateTimestepMesh also applies the diffusive limits (gas.cpp:437-467)
        if (f == AB200_GAS && c->has_diffusion) AB_TRY(launch_diffusion_dt(c, c->d_time + 1, 1));
      } else {
        AB_TRY(launch_estimate_dt(c, f, c->d_time + 1, any));
      }
    }
    any = 1;
  }
  AB_REQUIRE(any, AB200_ESTATE, "ab200_fused_stage: no fluid bound");
  return AB200_OK;
}

// Second half of a stage whose C2P was deferred (AB200_STAGE_DEFER_C2P) so that out-of-scope or
// library source terms could act on the conserved state in between: SetAuxillaryFields ->
// ConsToPrim -> PrimToCons (src/artemis_driver.cpp:251-261 without the exchange), and with
// AB200_STAGE_REDUCE_DT the CFL timestep of the new primitives into ab200_dt_device()[1].
int ab200_finish_stage(ab200_ctx *c, int flags) {
  AB_ENTER(c)
  AB_REQUIRE((flags & ~AB200_STAGE_REDUCE_DT) == 0, AB200_EINVAL, "ab200_finish_stage: unknown flag");
  const int reduce_dt = (flags & AB200_STAGE_REDUCE_DT) != 0;
  unsigned long long *slots = reinterpret_cast<unsigned long long *>(c->d_red + 3072);
  if (reduce_dt)
    AB_CUDA(cudaMemcpyAsync(slots, c->d_red + 3080, 2 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToDevice, c->stream));
  int any = 0;
  for (int f = 0; f < 2; ++f) {
    if (!c->fl[f].bound) continue;
    AB_TRY(sync_prim_home(c, f, 0));
    // one pointwise kernel per fluid: SetAux + C2P + interior P2C (+ CFL minimum)
    AB_TRY(launch_finish_stage(c, f, reduce_dt ? slots + f : nullptr));
    if (reduce_dt) {
      AB_TRY(launch_finish_dt(c, reinterpret_cast<const double *>(slots + f), 1, c->fl[f].d.cfl,
                              c->d_time + 1, any));
      if (f == AB200_GAS && c->has_diffusion) AB_TRY(launch_diffusion_dt(c, c->d_time + 1, 1));
    }
    any = 1;
  }
  AB_REQUIRE(any, AB200_ESTATE, "ab200_finish_stage: no fluid bound");
  return AB200_OK;
}

int ab200_halo_pack(ab200_ctx *c, const ab200_bnd_desc *bnd, int n) {
  AB_ENTER(c)
  if (n == 0) return AB200_OK;
  AB_REQUIRE(bnd && n > 0, AB200_EINVAL, "ab200_halo_pack: bad descriptor list");
  return launch_halo(c, bnd, n, 0);
}
int ab200_halo_unpack(ab200_ctx *c, const ab200_bnd_desc *bnd, int n) {
  AB_ENTER(c)
  if (n == 0) return AB200_OK;
  AB_REQUIRE(bnd && n > 0, AB200_EINVAL, "ab200_halo_unpack: bad descriptor list");
  return launch_halo(c, bnd, n, 1);
}

int ab200_set_topology(ab200_ctx *c, int nbx, int nby, int nbz, const int bc[6]) {
  AB_ENTER(c)
  AB_REQUIRE(nbx > 0 && nby > 0 && nbz > 0 && nbx * nby * nbz == c->g.nb, AB200_EINVAL,
             "ab200_set_topology: lattice does not match the number of bound blocks");
  for (int i = 0; i < 6; ++i)
    AB_REQUIRE(bc[i] >= AB200_BC_PERIODIC && bc[i] <= AB200_BC_FIXED, AB200_EINVAL,
               "ab200_set_topology: unknown boundary flag");
  {
    bool fixed = false, remote = false;
    for (int i = 0; i < 6; ++i) {
      fixed = fixed || bc[i] == AB200_BC_FIXED;
      remote = remote || bc[i] == AB200_BC_NONE;
    }
    AB_REQUIRE(!(fixed && remote), AB200_EINVAL,
               "ab200_set_topology: AB200_BC_FIXED faces need a single-rank topology "
               "(no AB200_BC_NONE face)");
  }
  c->topo.set = true;
  c->topo.nbx = nbx; c->topo.nby = nby; c->topo.nbz = nbz;
  for (int i = 0; i < 6; ++i) c->topo.bc[i] = bc[i];
  // surface / interior block lists (the surface is staged first so that the remote exchange
  // overlaps the interior blocks' stage, ab200_run_cycles_mr)
  std::vector<int> lists[2];
  const int nbd[3] = {nbx, nby, nbz};
  for (int b = 0; b < c->g.nb; ++b) {
    const int l[3] = {b % nbx, (b / nbx) % nby, b / (nbx * nby)};
    bool surf = false;
    for (int d = 0; d < 3; ++d) {
      if (l[d] == 0 && bc[2 * d] == AB200_BC_NONE) surf = true;
      if (l[d] == nbd[d] - 1 && bc[2 * d + 1] == AB200_BC_NONE) surf = true;
    }
    lists[surf ? 0 : 1].push_back(b);
  }
  for (int q = 0; q < 2; ++q) {
    if (c->d_blist[q]) cudaFree(c->d_blist[q]);
    c->d_blist[q] = nullptr;
    c->n_blist[q] = (int)lists[q].size();
    if (lists[q].empty()) continue;
    AB_CUDA(cudaMalloc((void **)&c->d_blist[q], lists[q].size() * sizeof(int)));
    AB_CUDA(cudaMemcpy(c->d_blist[q], lists[q].data(), lists[q].size() * sizeof(int),
                       cudaMemcpyHostToDevice));
  }
  return AB200_OK;
}

int ab200_exchange_ghosts(ab200_ctx *c) {
  AB_ENTER(c)
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_exchange_ghosts: call ab200_set_topology first");
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_exchange(c, f));
  return AB200_OK;
}

int ab200_fill_ghosts(ab200_ctx *c) {
  AB_ENTER(c)
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_fill_ghosts: call ab200_set_topology first");
  AB_REQUIRE(topology_is_local(c), AB200_ESTATE,
             "ab200_fill_ghosts: a face is flagged AB200_BC_NONE (neighbour on another rank); "
             "use ab200_exchange_ghosts + halo pack/unpack + ab200_apply_physical_bcs");
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_fill_ghosts(c, f, 0));
  return AB200_OK;
}

int ab200_fill_ghosts_local(ab200_ctx *c) {
  AB_ENTER(c)
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_fill_ghosts_local: call ab200_set_topology first");
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_fill_ghosts(c, f, 0));
  return AB200_OK;
}

int ab200_finish_remote_ghosts(ab200_ctx *c) {
  AB_ENTER(c)
  AB_REQUIRE(c->topo.set, AB200_ESTATE,
             "ab200_finish_remote_ghosts: call ab200_set_topology first");
  if (topology_is_local(c)) return AB200_OK;  // no face belongs to another rank
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_fill_ghosts(c, f, 1));
  return AB200_OK;
}

int ab200_apply_physical_bcs(ab200_ctx *c) {
  AB_ENTER(c)
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_apply_physical_bcs: call ab200_set_topology first");
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(launch_physical_bcs(c, f));
  return AB200_OK;
}

int ab200_malloc(ab200_ctx *c, void **dptr, size_t bytes) {
  AB_REQUIRE(c && dptr, AB200_EINVAL, "ab200_malloc: null argument");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaMalloc(dptr, bytes ? bytes : 8));
  return AB200_OK;
}
int ab200_free(ab200_ctx *c, void *dptr) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaFree(dptr));
  return AB200_OK;
}
int ab200_memcpy_h2d(ab200_ctx *c, void *dst, const void *src, size_t bytes) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  return AB200_OK;
}
int ab200_memcpy_d2h(ab200_ctx *c, void *dst, const void *src, size_t bytes) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  return AB200_OK;
}
long long ab200_launch_count(ab200_ctx *c) { return c ? c->launches : -1; }

int ab200_timer_begin(ab200_ctx *c) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaEventRecord(c->ev0, c->stream));
  return AB200_OK;
}
int ab200_timer_end(ab200_ctx *c, float *ms) {
  AB_REQUIRE(c && ms, AB200_EINVAL, "null argument");
  AB_CUDA(cudaSetDevice(c->device));
  AB_CUDA(cudaEventRecord(c->ev1, c->stream));
  AB_CUDA(cudaEventSynchronize(c->ev1));
  AB_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return AB200_OK;
}

}  // extern "C"
