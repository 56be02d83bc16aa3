/* abi_roundtrip.c -- a plain C99 host that drives libartemis_b200 through include/ab200.h the
 * way the Parthenon glue TU would: no Python, no ctypes, no torch.
 *
 *   gcc -std=c99 -I include tests/c/abi_roundtrip.c -L artemis_b200/lib -lartemis_b200 \
 *       -Wl,-rpath,$PWD/artemis_b200/lib -lm -o abi_roundtrip
 *
 * Binds a 2-block (2 x 1 x 1 lattice of 8 x 6 x 4 zones, nghost 4) periodic gas mesh, fills a
 * deterministic smooth state, runs PrimToCons -> one rk2 cycle as two ab200_fused_stage +
 * ab200_fill_ghosts -> the device dt bookkeeping, downloads u0 / prim and prints an order-
 * dependent checksum of the raw bit patterns.  tests/test_c_abi.py builds the same state
 * through ctypes and requires the same checksum.  Without a CUDA device ab200_create fails
 * with AB200_ECUDA (there is no CPU fallback): the program reports that and exits 77. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ab200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != AB200_OK) {                                                           \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ab200_last_error());       \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

enum { NB = 2, NX1 = 8, NX2 = 6, NX3 = 4, NG = 4, NVAR = 6 };

static uint64_t fnv1a(const double *a, size_t n, uint64_t h) {
  const unsigned char *p = (const unsigned char *)a;
  for (size_t i = 0; i < 8 * n; ++i) {
    h ^= p[i];
    h *= 1099511628211ULL;
  }
  return h;
}

int main(int argc, char **argv) {
  const char *path = argc > 1 ? argv[1] : "auto";
  printf("abi_version %d\n", ab200_abi_version());
  if (ab200_abi_version() != AB200_ABI_VERSION) {
    fprintf(stderr, "header / library ABI mismatch\n");
    return 1;
  }
  ab200_ctx *ctx = NULL;
  int rc = ab200_create(&ctx, 0, NULL);
  if (rc == AB200_ECUDA) {
    printf("no CUDA device: %s\n", ab200_last_error());
    return 77;
  }
  if (rc != AB200_OK) {
    fprintf(stderr, "ab200_create failed (%d): %s\n", rc, ab200_last_error());
    return 1;
  }

  const int ni = NX1 + 2 * NG, nj = NX2 + 2 * NG, nk = NX3 + 2 * NG;
  const size_t cells = (size_t)ni * nj * nk;
  const double gamma = 1.4;
  /* mesh [0, 1.6] x [0, 0.6] x [0, 0.4], two blocks along x1; UniformCartesian per block */
  double xmin[NB][3], dx[NB][3];
  for (int b = 0; b < NB; ++b) {
    dx[b][0] = 0.8 / NX1; dx[b][1] = 0.6 / NX2; dx[b][2] = 0.4 / NX3;
    xmin[b][0] = 0.8 * b - NG * dx[b][0];
    xmin[b][1] = -NG * dx[b][1];
    xmin[b][2] = -NG * dx[b][2];
  }
  ab200_grid_desc g;
  memset(&g, 0, sizeof g);
  g.geom = AB200_CARTESIAN; g.ndim = 3; g.nghost = NG; g.nblocks = NB;
  g.ni = ni; g.nj = nj; g.nk = nk;
  g.is = NG; g.ie = NG + NX1 - 1; g.js = NG; g.je = NG + NX2 - 1; g.ks = NG; g.ke = NG + NX3 - 1;
  g.fni = ni + 1; g.fnj = nj + 1; g.fnk = nk + 1;
  g.xmin = &xmin[0][0]; g.dx = &dx[0][0];
  CHECK(ab200_set_grid(ctx, &g));

  /* device arrays in the MeshBlockPack layout: one dense [nk][nj][ni] array per (block, var) */
  double *slab[3];
  for (int a = 0; a < 3; ++a) CHECK(ab200_malloc(ctx, (void **)&slab[a], 8 * cells * NB * NVAR));
  double *tab[3][NB * NVAR];
  for (int a = 0; a < 3; ++a)
    for (int e = 0; e < NB * NVAR; ++e) tab[a][e] = slab[a] + (size_t)e * cells;
  ab200_fluid_desc fd = {AB200_GAS, 1, AB200_PPM, AB200_HLLC, gamma - 1.0, 1e-10, 1e-10, 0.0, 0.3};
  ab200_pack_desc pk;
  memset(&pk, 0, sizeof pk);
  pk.prim = tab[0]; pk.cons0 = tab[1]; pk.cons1 = tab[2];
  CHECK(ab200_bind_pack(ctx, &fd, &pk));
  int bc[6] = {AB200_BC_PERIODIC, AB200_BC_PERIODIC, AB200_BC_PERIODIC,
               AB200_BC_PERIODIC, AB200_BC_PERIODIC, AB200_BC_PERIODIC};
  CHECK(ab200_set_topology(ctx, NB, 1, 1, bc));
  int pcode = !strcmp(path, "three_pass") ? 1 : !strcmp(path, "single_pass") ? 2
              : !strcmp(path, "role_split") ? 3 : 0;
  CHECK(ab200_set_stage_path(ctx, pcode));

  /* smooth periodic primitives, entire domain (ghosts are overwritten by the exchange) */
  double *h = (double *)malloc(8 * cells * NB * NVAR);
  const double PI2 = 6.283185307179586;
  for (int b = 0; b < NB; ++b)
    for (int k = 0; k < nk; ++k)
      for (int j = 0; j < nj; ++j)
        for (int i = 0; i < ni; ++i) {
          const double x = xmin[b][0] + (i + 0.5) * dx[b][0], y = xmin[b][1] + (j + 0.5) * dx[b][1],
                       z = xmin[b][2] + (k + 0.5) * dx[b][2];
          const double s = sin(PI2 * x / 1.6) * cos(PI2 * y / 0.6), c = cos(PI2 * z / 0.4);
          const size_t o = ((size_t)k * nj + j) * ni + i;
          double *q = h + (size_t)b * NVAR * cells;
          q[0 * cells + o] = 1.0 + 0.2 * s * c;
          q[1 * cells + o] = 0.3 * c;
          q[2 * cells + o] = -0.2 * s;
          q[3 * cells + o] = 0.1 * s * c;
          q[5 * cells + o] = 1.5 + 0.3 * s;
          q[4 * cells + o] = (gamma - 1.0) * q[0 * cells + o] * q[5 * cells + o];
        }
  CHECK(ab200_memcpy_h2d(ctx, slab[0], h, 8 * cells * NB * NVAR));

  /* Mesh::Initialize sequence, then one rk2 cycle on the device-resident path */
  CHECK(ab200_prim_to_cons(ctx));
  CHECK(ab200_cons_to_prim(ctx));
  CHECK(ab200_fill_ghosts(ctx));
  double ts[4] = {1.79769313486231570815e+308, 0.0, 0.0, 0.0};
  CHECK(ab200_write_time_state(ctx, ts));
  CHECK(ab200_estimate_timestep_device(ctx));
  CHECK(ab200_set_global_timestep_device(ctx, 1.79769313486231570815e+308, 0));
  const double gam[2][3] = {{0.0, 1.0, 1.0}, {0.5, 0.5, 0.5}};
  for (int s = 0; s < 2; ++s) {
    const int flags = AB200_STAGE_DEVICE_DT | AB200_STAGE_PINGPONG | (s == 1 ? AB200_STAGE_REDUCE_DT : 0);
    CHECK(ab200_fused_stage(ctx, gam[s][0], gam[s][1], gam[s][2], 0.0, 0, s == 0, flags));
    CHECK(ab200_fill_ghosts(ctx));
  }
  CHECK(ab200_set_global_timestep_device(ctx, 1.79769313486231570815e+308, 1));
  CHECK(ab200_sync_prim(ctx));
  CHECK(ab200_read_time_state(ctx, ts));

  uint64_t hsum = 1469598103934665603ULL;
  CHECK(ab200_memcpy_d2h(ctx, h, slab[1], 8 * cells * NB * NVAR));
  hsum = fnv1a(h, cells * NB * NVAR, hsum);
  CHECK(ab200_memcpy_d2h(ctx, h, slab[0], 8 * cells * NB * NVAR));
  hsum = fnv1a(h, cells * NB * NVAR, hsum);
  int used = 0;
  CHECK(ab200_get_stage_path(ctx, AB200_GAS, &used));
  printf("stage_path %d\n", used);
  printf("ncycle %.0f time %.17g dt %.17g\n", ts[3], ts[2], ts[0]);
  printf("launches %lld\n", ab200_launch_count(ctx));
  printf("checksum %016llx\n", (unsigned long long)hsum);
  free(h);
  for (int a = 0; a < 3; ++a) CHECK(ab200_free(ctx, slab[a]));
  CHECK(ab200_destroy(ctx));
  return 0;
}
