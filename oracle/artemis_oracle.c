/*
 * artemis_oracle.c -- CPU restatement of the Artemis finite-volume gas/dust stage update.
 *
 * TEST INFRASTRUCTURE ONLY (see artemis_oracle.h).  Never linked into the product.
 *
 * Compile WITHOUT FMA contraction: gcc -O2 -ffp-contract=off -fopenmp (oracle/Makefile).
 * All citations are file:line in lanl/artemis @ 6c2a7a8 ("P:" = external/parthenon/src/).
 */
#include "artemis_oracle.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SQR(x) ((x) * (x))
#define IDX(g, nvar, b, n, k, j, i)                                                      \
  (((((size_t)(b) * (nvar) + (n)) * (g)->nk + (k)) * (g)->nj + (j)) * (g)->ni + (i))
#define FIDX(g, S, b, n, k, j, i)                                                        \
  (((((size_t)(b) * (S) + (n)) * (g)->fnk + (k)) * (g)->fnj + (j)) * (g)->fni + (i))

static inline double dmax(double a, double b) { return a > b ? a : b; } /* std::max */
static inline double dmin(double a, double b) { return b < a ? b : a; } /* std::min */

int ao_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ===================================================================================== */
/* Geometry: src/geometry/{geometry,cylindrical,spherical,axisymmetric}.hpp               */
/* ===================================================================================== */
typedef struct {
  double x1[2], x2[2], x3[2];
} bbox_t;

/* geometry.hpp:61-79 (BBox) with P:coordinates/uniform_cartesian.hpp:153-157 (Xf) */
static inline bbox_t make_bbox(const double *xmin, const double *dx, int k, int j, int i) {
  bbox_t b;
  b.x1[0] = xmin[0] + i * dx[0];
  b.x1[1] = xmin[0] + (i + 1) * dx[0];
  b.x2[0] = xmin[1] + j * dx[1];
  b.x2[1] = xmin[1] + (j + 1) * dx[1];
  b.x3[0] = xmin[2] + k * dx[2];
  b.x3[1] = xmin[2] + (k + 1) * dx[2];
  return b;
}

static inline int is_sph(int geom) {
  return geom == AO_SPHERICAL1D || geom == AO_SPHERICAL2D || geom == AO_SPHERICAL3D;
}
/* geometry.hpp:102-113 */
static inline int g_x1dep(int geom) { return geom != AO_CARTESIAN; }
static inline int g_x2dep(int geom) {
  return geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D;
}

/* <r> = d(r^3/3)/d(r^2/2): cylindrical.hpp:52-56, axisymmetric.hpp:48-52 */
static inline double rface_avg(const bbox_t *b) {
  return 2.0 / 3.0 * (b->x1[0] * b->x1[0] + b->x1[0] * b->x1[1] + b->x1[1] * b->x1[1]) /
         (b->x1[0] + b->x1[1]);
}

static inline double g_x1v(int geom, const bbox_t *b) {
  if (geom == AO_CYLINDRICAL || geom == AO_AXISYMMETRIC) return rface_avg(b);
  if (is_sph(geom)) { /* spherical.hpp:57-60, 261-264, 448-451 */
    const double dr2 = b->x1[0] * b->x1[0] + b->x1[1] * b->x1[1];
    return 0.75 * (b->x1[0] + b->x1[1]) * dr2 / (dr2 + b->x1[0] * b->x1[1]);
  }
  return 0.5 * (b->x1[0] + b->x1[1]); /* geometry.hpp:166 */
}
static inline double g_x2v(int geom, const bbox_t *b) {
  if (geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D) { /* spherical.hpp:61-68 */
    const double ctm = cos(b->x2[0]);
    const double ctp = cos(b->x2[1]);
    const double dst = sin(b->x2[1]) - sin(b->x2[0]);
    return (dst - b->x2[1] * ctp + b->x2[0] * ctm) / fabs(ctm - ctp);
  }
  return 0.5 * (b->x2[0] + b->x2[1]);
}
static inline double g_x3v(int geom, const bbox_t *b) {
  (void)geom;
  return 0.5 * (b->x3[0] + b->x3[1]);
}

/* scale factors h_i(x): geometry.hpp:172-180 + overrides */
static inline double g_hx1(int geom, double x1, double x2, double x3) {
  (void)geom; (void)x1; (void)x2; (void)x3;
  return 1.0;
}
static inline double g_hx2(int geom, double x1, double x2, double x3) {
  (void)x2; (void)x3;
  if (geom == AO_CYLINDRICAL || is_sph(geom)) return x1; /* cyl:58, sph:49,253,444 */
  return 1.0;
}
static inline double g_hx3(int geom, double x1, double x2, double x3) {
  (void)x3;
  if (geom == AO_AXISYMMETRIC) return x1; /* axisymmetric.hpp:54 */
  if (geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D) return x1 * sin(x2); /* sph:53 */
  return 1.0; /* NB spherical1D does NOT override hx3 */
}
/* volume-averaged scale factors: geometry.hpp:182-184 + overrides */
static inline double g_hx1v(int geom, const bbox_t *b) {
  (void)geom; (void)b;
  return 1.0;
}
static inline double g_hx2v(int geom, const bbox_t *b) {
  if (geom == AO_CYLINDRICAL || geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D)
    return g_x1v(geom, b); /* cyl:62, sph:70,274 ; spherical1D keeps the default 1.0 */
  return 1.0;
}
static inline double g_hx3v(int geom, const bbox_t *b) {
  if (geom == AO_AXISYMMETRIC) return g_x1v(geom, b); /* axisymmetric.hpp:58 */
  if (geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D) { /* spherical.hpp:71-85 */
    const double ctm = cos(b->x2[0]);
    const double ctp = cos(b->x2[1]);
    const double stm = sin(b->x2[0]);
    const double stp = sin(b->x2[1]);
    const double dsc = stp * ctp - stm * ctm;
    const double dx2 = b->x2[1] - b->x2[0];
    return g_x1v(geom, b) * 0.5 * (dx2 - dsc) / fabs(ctm - ctp);
  }
  return 1.0;
}

/* face centroids: geometry.hpp:186-201 + overrides */
static inline void g_facecen1(int geom, const bbox_t *b, int f, double xf[3]) {
  xf[0] = b->x1[f];
  xf[1] = g_x2v(geom, b);
  xf[2] = g_x3v(geom, b);
}
static inline void g_facecen2(int geom, const bbox_t *b, int f, double xf[3]) {
  if (geom == AO_AXISYMMETRIC || geom == AO_SPHERICAL3D) { /* axi:60-67, sph:87-94 */
    xf[0] = rface_avg(b);
    xf[1] = b->x2[f];
    xf[2] = 0.5 * (b->x3[0] + b->x3[1]);
  } else if (geom == AO_SPHERICAL2D) { /* spherical.hpp:291-298 */
    xf[0] = rface_avg(b);
    xf[1] = b->x2[f];
    xf[2] = 0.0;
  } else if (geom == AO_SPHERICAL1D) { /* spherical.hpp:453-460 */
    xf[0] = rface_avg(b);
    xf[1] = M_PI * 0.5;
    xf[2] = 0.0;
  } else {
    xf[0] = g_x1v(geom, b);
    xf[1] = b->x2[f];
    xf[2] = g_x3v(geom, b);
  }
}
static inline void g_facecen3(int geom, const bbox_t *b, int f, double xf[3]) {
  if (geom == AO_CYLINDRICAL || geom == AO_SPHERICAL3D) { /* cyl:64-71, sph:96-103 */
    xf[0] = rface_avg(b);
    xf[1] = 0.5 * (b->x2[0] + b->x2[1]);
    xf[2] = b->x3[f];
  } else if (geom == AO_SPHERICAL2D) { /* spherical.hpp:300-307 */
    xf[0] = rface_avg(b);
    xf[1] = 0.5 * (b->x2[0] + b->x2[1]);
    xf[2] = 0.0;
  } else if (geom == AO_SPHERICAL1D) { /* spherical.hpp:462-469 */
    xf[0] = rface_avg(b);
    xf[1] = M_PI * 0.5;
    xf[2] = 0.0;
  } else {
    xf[0] = g_x1v(geom, b);
    xf[1] = g_x2v(geom, b);
    xf[2] = b->x3[f];
  }
}

/* face areas: geometry.hpp:203-220 + overrides */
static inline double g_area1(int geom, const bbox_t *b, double x1f) {
  const double dx2 = b->x2[1] - b->x2[0];
  const double dx3 = b->x3[1] - b->x3[0];
  switch (geom) {
  case AO_CYLINDRICAL:
  case AO_AXISYMMETRIC: return x1f * dx2 * dx3; /* cyl:73-78, axi:69-74 */
  case AO_SPHERICAL3D:                          /* spherical.hpp:105-109 */
    return x1f * x1f * fabs(cos(b->x2[0]) - cos(b->x2[1])) * dx3;
  case AO_SPHERICAL2D: return x1f * x1f * fabs(cos(b->x2[0]) - cos(b->x2[1])); /* :309 */
  case AO_SPHERICAL1D: return x1f * x1f;                                       /* :471 */
  default: return dx2 * dx3;
  }
}
static inline double g_area2(int geom, const bbox_t *b, double x2f) {
  const double dx1 = b->x1[1] - b->x1[0];
  const double dx3 = b->x3[1] - b->x3[0];
  switch (geom) {
  case AO_AXISYMMETRIC: return (b->x1[0] + b->x1[1]) * 0.5 * dx1 * dx3; /* axi:75-80 */
  case AO_SPHERICAL3D: return 0.5 * (b->x1[1] + b->x1[0]) * sin(x2f) * dx1 * dx3; /*:110*/
  case AO_SPHERICAL2D: return 0.5 * (b->x1[1] + b->x1[0]) * sin(x2f) * dx1;       /*:313*/
  case AO_SPHERICAL1D: return 0.5 * (b->x1[1] + b->x1[0]) * dx1;                  /*:475*/
  default: return dx1 * dx3;
  }
}
static inline double g_area3(int geom, const bbox_t *b, double x3f) {
  (void)x3f;
  const double dx1 = b->x1[1] - b->x1[0];
  const double dx2 = b->x2[1] - b->x2[0];
  switch (geom) {
  case AO_CYLINDRICAL:
  case AO_SPHERICAL3D:
  case AO_SPHERICAL2D: return 0.5 * (b->x1[0] + b->x1[1]) * dx1 * dx2; /* cyl:79, sph:116 */
  case AO_SPHERICAL1D: return 0.5 * (b->x1[0] + b->x1[1]) * dx1;       /* sph:480 */
  default: return dx1 * dx2;
  }
}
/* geometry.hpp:222-228 + overrides */
static inline double g_volume(int geom, const bbox_t *b) {
  const double dx1 = b->x1[1] - b->x1[0];
  const double dx2 = b->x2[1] - b->x2[0];
  const double dx3 = b->x3[1] - b->x3[0];
  if (geom == AO_CYLINDRICAL || geom == AO_AXISYMMETRIC)
    return (b->x1[0] + b->x1[1]) * 0.5 * dx1 * dx2 * dx3; /* cyl:85-90, axi:82-87 */
  if (is_sph(geom)) {                                     /* sph:123-132,327-335,486 */
    const double rfac =
        (b->x1[0] * b->x1[0] + b->x1[0] * b->x1[1] + b->x1[1] * b->x1[1]) / 3.0;
    if (geom == AO_SPHERICAL1D) return rfac * dx1;
    const double dc = fabs(cos(b->x2[0]) - cos(b->x2[1]));
    if (geom == AO_SPHERICAL2D) return rfac * dx1 * dc;
    return rfac * dx1 * dc * dx3;
  }
  return dx1 * dx2 * dx3;
}
/* connection terms {dh1/dxd, dh2/dxd, dh3/dxd}: geometry.hpp:238-248 + overrides */
static inline void g_conn1(int geom, const bbox_t *b, double c[3]) {
  c[0] = c[1] = c[2] = 0.0;
  if (geom == AO_CYLINDRICAL) c[1] = 1.0 / (0.5 * (b->x1[0] + b->x1[1])); /* cyl:92 */
  if (geom == AO_AXISYMMETRIC) c[2] = 1.0 / (0.5 * (b->x1[0] + b->x1[1])); /* axi:89 */
  if (is_sph(geom)) { /* spherical.hpp:134-141 */
    const double v =
        3.0 / 2.0 * (b->x1[0] + b->x1[1]) /
        (b->x1[0] * b->x1[0] + b->x1[0] * b->x1[1] + b->x1[1] * b->x1[1]);
    c[1] = v;
    c[2] = v;
  }
}
static inline void g_conn2(int geom, const bbox_t *b, double c[3]) {
  c[0] = c[1] = c[2] = 0.0;
  if (geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D) /* spherical.hpp:142-145 */
    c[2] = (sin(b->x2[1]) - sin(b->x2[0])) / fabs(cos(b->x2[0]) - cos(b->x2[1]));
}
/* geometry.hpp:330-357 cell widths h_d(xv) * dx_d */
static inline void g_cell_widths(int geom, const bbox_t *b, double w[3]) {
  const double xv[3] = {g_x1v(geom, b), g_x2v(geom, b), g_x3v(geom, b)};
  w[0] = g_hx1(geom, xv[0], xv[1], xv[2]) * (b->x1[1] - b->x1[0]);
  w[1] = g_hx2(geom, xv[0], xv[1], xv[2]) * (b->x2[1] - b->x2[0]);
  w[2] = g_hx3(geom, xv[0], xv[1], xv[2]) * (b->x3[1] - b->x3[0]);
}
/* src/rotating_frame/rotating_frame.hpp:32-49 with the ConvertToCylWithVec of each geom */
static inline void g_rotation_velocity(int geom, const double xv[3], double omf,
                                       double vf[3]) {
  vf[0] = vf[1] = vf[2] = 0.0;
  switch (geom) {
  case AO_CARTESIAN: vf[1] = omf; break;
  case AO_CYLINDRICAL: vf[1] = 1.0 * (omf * xv[0]); break; /* cyl:134-141 ex2[1]=1 */
  case AO_AXISYMMETRIC: vf[2] = 1.0 * (omf * xv[0]); break; /* axi:137-147 ex3[1]=1 */
  case AO_SPHERICAL3D:
  case AO_SPHERICAL2D: vf[2] = 1.0 * (omf * (xv[0] * sin(xv[1]))); break; /* sph:208-230 */
  case AO_SPHERICAL1D: vf[2] = 1.0 * (omf * (xv[0] * 1.0)); break;
  }
}

void ao_geom_cell(int geom, const double *xmin, const double *dx, int k, int j, int i,
                  double *out) {
  bbox_t b = make_bbox(xmin, dx, k, j, i);
  double c[3], w[3], xf[3];
  out[0] = g_x1v(geom, &b);
  out[1] = g_x2v(geom, &b);
  out[2] = g_x3v(geom, &b);
  out[3] = g_hx1v(geom, &b);
  out[4] = g_hx2v(geom, &b);
  out[5] = g_hx3v(geom, &b);
  out[6] = g_volume(geom, &b);
  out[7] = g_area1(geom, &b, b.x1[0]);
  out[8] = g_area1(geom, &b, b.x1[1]);
  out[9] = g_area2(geom, &b, b.x2[0]);
  out[10] = g_area2(geom, &b, b.x2[1]);
  out[11] = g_area3(geom, &b, b.x3[0]);
  out[12] = g_area3(geom, &b, b.x3[1]);
  g_conn1(geom, &b, c);
  out[13] = c[0]; out[14] = c[1]; out[15] = c[2];
  g_conn2(geom, &b, c);
  out[16] = c[0]; out[17] = c[1]; out[18] = c[2];
  g_cell_widths(geom, &b, w);
  out[19] = w[0]; out[20] = w[1]; out[21] = w[2];
  g_facecen1(geom, &b, 0, xf);
  out[22] = g_hx1(geom, xf[0], xf[1], xf[2]);
  out[23] = g_hx2(geom, xf[0], xf[1], xf[2]);
  out[24] = g_hx3(geom, xf[0], xf[1], xf[2]);
  g_facecen2(geom, &b, 0, xf);
  out[25] = g_hx1(geom, xf[0], xf[1], xf[2]);
  out[26] = g_hx2(geom, xf[0], xf[1], xf[2]);
  out[27] = g_hx3(geom, xf[0], xf[1], xf[2]);
  g_facecen3(geom, &b, 0, xf);
  out[28] = g_hx1(geom, xf[0], xf[1], xf[2]);
  out[29] = g_hx2(geom, xf[0], xf[1], xf[2]);
  out[30] = g_hx3(geom, xf[0], xf[1], xf[2]);
  out[31] = 0.0;
}

/* ===================================================================================== */
/* Reconstruction: src/utils/fluxes/reconstruction/{plm,ppm}.hpp                          */
/* ===================================================================================== */
/* plm.hpp:31-47 */
void ao_plm(double q_im1, double q_i, double q_ip1, double *ql_ip1, double *qr_i) {
  double dql = (q_i - q_im1);
  double dqr = (q_ip1 - q_i);
  double dq2 = dql * dqr;
  double dqm = dq2 / (dql + dqr);
  if (dq2 <= 0.0) dqm = 0.0;
  *ql_ip1 = q_i + dqm;
  *qr_i = q_i - dqm;
}
/* plm.hpp:53-73 */
void ao_plm_g(double q_im1, double q_i, double q_ip1, double x_im1, double x_i,
              double x_ip1, double xf0, double xf1, double dx, double *ql_ip1,
              double *qr_i) {
  const double dql = (q_i - q_im1) * dx / (x_i - x_im1);
  const double dqr = (q_ip1 - q_i) * dx / (x_ip1 - x_i);
  const double dq2 = dql * dqr;
  const double cr = (x_ip1 - x_i) / (xf1 - x_i);
  const double cl = (x_i - x_im1) / (x_i - xf0);
  const double dqm = (dq2 <= 0.0) ? 0.0
                                  : dq2 * (cr * dql + cl * dqr) /
                                        (dql * dql + dqr * dqr + dq2 * (cl + cr - 2.0));
  *ql_ip1 = q_i + dqm * (xf1 - x_i) / dx;
  *qr_i = q_i - dqm * (x_i - xf0) / dx;
}
/* ppm.hpp:32-66 */
void ao_ppm4(double q_im2, double q_im1, double q_i, double q_ip1, double q_ip2,
             double *ql_ip1, double *qr_i) {
  double qlv = (7. * (q_i + q_im1) - (q_im2 + q_ip1)) / 12.0;
  double qrv = (7. * (q_i + q_ip1) - (q_im1 + q_ip2)) / 12.0;
  qlv = dmax(qlv, dmin(q_i, q_im1));
  qlv = dmin(qlv, dmax(q_i, q_im1));
  qrv = dmax(qrv, dmin(q_i, q_ip1));
  qrv = dmin(qrv, dmax(q_i, q_ip1));
  double qc = qrv - q_i;
  double qd = qlv - q_i;
  if ((qc * qd) >= 0.0) {
    qlv = q_i;
    qrv = q_i;
  } else {
    if (fabs(qc) >= 2.0 * fabs(qd)) qrv = q_i - 2.0 * qd;
    if (fabs(qd) >= 2.0 * fabs(qc)) qlv = q_i - 2.0 * qc;
  }
  *ql_ip1 = qrv;
  *qr_i = qlv;
}

/* Reconstruction<R,DIR,GEOM>::apply (pcm.hpp:30-88, plm.hpp:78-175, ppm.hpp:71-129):
 * reconstruct cells (k,j,i), i in [il,iu], along `dir`; cell i's upper-face value goes to
 * ql[n][i + (dir==1)] and its lower-face value to qr[n][i].  Scratch rows are ni+1 wide. */
static void recon_row(const ao_grid *g, int recon, int dir, int nvar, const double *prim,
                      int b, int k, int j, int il, int iu, double *ql, double *qr) {
  const int W = g->ni + 1;
  const ptrdiff_t s1 = 1, s2 = (ptrdiff_t)g->ni, s3 = (ptrdiff_t)g->ni * g->nj;
  const ptrdiff_t st = dir == 1 ? s1 : (dir == 2 ? s2 : s3);
  const int sh = (dir == 1) ? 1 : 0;
  const double *xmin = g->xmin + 3 * b, *dx = g->dx + 3 * b;
  for (int n = 0; n < nvar; ++n) {
    const double *q = prim + IDX(g, nvar, b, n, k, j, 0);
    double *qln = ql + (size_t)n * W, *qrn = qr + (size_t)n * W;
    for (int i = il; i <= iu; ++i) {
      if (recon == AO_PCM) {
        qln[i + sh] = q[i];
        qrn[i] = q[i];
      } else if (recon == AO_PLM) {
        if (g->geom == AO_CARTESIAN) {
          ao_plm(q[i - st], q[i], q[i + st], &qln[i + sh], &qrn[i]);
        } else {
          const int dk = (dir == 3), dj = (dir == 2), di = (dir == 1);
          bbox_t bm = make_bbox(xmin, dx, k - dk, j - dj, i - di);
          bbox_t bc = make_bbox(xmin, dx, k, j, i);
          bbox_t bp = make_bbox(xmin, dx, k + dk, j + dj, i + di);
          double xvm, xvc, xvp, w[3];
          const double *xf;
          g_cell_widths(g->geom, &bc, w);
          if (dir == 1) {
            xvm = g_x1v(g->geom, &bm); xvc = g_x1v(g->geom, &bc); xvp = g_x1v(g->geom, &bp);
            xf = bc.x1;
          } else if (dir == 2) {
            xvm = g_x2v(g->geom, &bm); xvc = g_x2v(g->geom, &bc); xvp = g_x2v(g->geom, &bp);
            xf = bc.x2;
          } else {
            xvm = g_x3v(g->geom, &bm); xvc = g_x3v(g->geom, &bc); xvp = g_x3v(g->geom, &bp);
            xf = bc.x3;
          }
          ao_plm_g(q[i - st], q[i], q[i + st], xvm, xvc, xvp, xf[0], xf[1], w[dir - 1],
                   &qln[i + sh], &qrn[i]);
        }
      } else {
        ao_ppm4(q[i - 2 * st], q[i - st], q[i], q[i + st], q[i + 2 * st], &qln[i + sh],
                &qrn[i]);
      }
    }
  }
}

/* ===================================================================================== */
/* Riemann solvers: src/utils/fluxes/riemann/{hllc,hlle,llf}.hpp                          */
/* wl/wr = {rho, vx(normal), vy, vz, P, sie}; out = {Frho,Fmx,Fmy,Fmz,FE,Fu,pface,vface}   */
/* ===================================================================================== */
/* hllc.hpp:76-180 */
static void solve_hllc(double gm1, const double *wl, const double *wr, double *out) {
  const double wl_idn = wl[0], wl_ivx = wl[1], wl_ivy = wl[2], wl_ivz = wl[3],
               wl_ipr = wl[4], wl_ise = wl[5];
  const double wr_idn = wr[0], wr_ivx = wr[1], wr_ivy = wr[2], wr_ivz = wr[3],
               wr_ipr = wr[4], wr_ise = wr[5];
  double igm1 = 1.0 / gm1;
  double gamma = gm1 + 1.0;
  double alpha = (gamma + 1.0) / (2.0 * gamma);
  double qa, qb, qc, qd, qe, qf;
  qa = sqrt(gamma * wl_ipr / wl_idn);
  qb = sqrt(gamma * wr_ipr / wr_idn);
  double el = wl_ipr * igm1 + 0.5 * wl_idn * (SQR(wl_ivx) + SQR(wl_ivy) + SQR(wl_ivz));
  double er = wr_ipr * igm1 + 0.5 * wr_idn * (SQR(wr_ivx) + SQR(wr_ivy) + SQR(wr_ivz));
  qc = 0.25 * (wl_idn + wr_idn) * (qa + qb);
  qd = 0.5 * (wl_ipr + wr_ipr + (wl_ivx - wr_ivx) * qc);
  qe = (qd <= wl_ipr) ? 1.0 : sqrt(1.0 + alpha * ((qd / wl_ipr) - 1.0));
  qf = (qd <= wr_ipr) ? 1.0 : sqrt(1.0 + alpha * ((qd / wr_ipr) - 1.0));
  double sl = wl_ivx - qa * qe;
  double sr = wr_ivx + qb * qf;
  qa = sr > 0.0 ? sr : 1.0e-20;
  qb = sl < 0.0 ? sl : -1.0e-20;
  qe = wl_ivx - sl;
  qf = wr_ivx - sr;
  qc = wl_ipr + qe * wl_idn * wl_ivx;
  qd = wr_ipr + qf * wr_idn * wr_ivx;
  double ml = wl_idn * qe;
  double mr = -(wr_idn * qf);
  double am = (qc - qd) / (ml + mr);
  double cp = (ml * qd + mr * qc) / (ml + mr);
  cp = cp > 0.0 ? cp : 0.0;
  qe = wl_idn * (wl_ivx - qb);
  qf = wr_idn * (wr_ivx - qa);
  double fld = qe, frd = qf;
  double flmx = qe * wl_ivx, frmx = qf * wr_ivx;
  double flmy = qe * wl_ivy, frmy = qf * wr_ivy;
  double flmz = qe * wl_ivz, frmz = qf * wr_ivz;
  double fle = el * (wl_ivx - qb) + wl_ipr * wl_ivx;
  double fre = er * (wr_ivx - qa) + wr_ipr * wr_ivx;
  if (am >= 0.0) {
    qc = am / (am - qb);
    qd = 0.0;
    qe = -qb / (am - qb);
  } else {
    qc = 0.0;
    qd = -am / (qa - am);
    qe = qa / (qa - am);
  }
  out[6] = qc * wl_ipr + qd * wr_ipr + qe * cp;
  const double frho = qc * fld + qd * frd;
  out[0] = frho;
  out[1] = qc * flmx + qd * frmx;
  out[2] = qc * flmy + qd * frmy;
  out[3] = qc * flmz + qd * frmz;
  out[4] = qc * fle + qd * fre + qe * cp * am;
  out[5] = frho * ((frho >= 0.0) ? wl_ise : wr_ise);
  out[7] = frho / ((frho >= 0.0) ? wl_idn : wr_idn);
}

/* hlle.hpp:92-220 */
static void solve_hlle(int fluid, double gm1, const double *wl, const double *wr,
                       double *out) {
  const int gas = (fluid == AO_GAS);
  const double wl_idn = wl[0], wl_ivx = wl[1], wl_ivy = wl[2], wl_ivz = wl[3];
  const double wr_idn = wr[0], wr_ivx = wr[1], wr_ivy = wr[2], wr_ivz = wr[3];
  double wl_ipr = 0, wr_ipr = 0, wl_ise = 0, wr_ise = 0, igm1 = 0, gamma = 0;
  if (gas) {
    wl_ipr = wl[4]; wl_ise = wl[5]; wr_ipr = wr[4]; wr_ise = wr[5];
    igm1 = 1.0 / gm1;
    gamma = gm1 + 1.0;
  }
  double sqrtdl = sqrt(wl_idn);
  double sqrtdr = sqrt(wr_idn);
  double isdlpdr = 1.0 / (sqrtdl + sqrtdr);
  double wroe_ivx = (sqrtdl * wl_ivx + sqrtdr * wr_ivx) * isdlpdr;
  double wroe_ivy = (sqrtdl * wl_ivy + sqrtdr * wr_ivy) * isdlpdr;
  double wroe_ivz = (sqrtdl * wl_ivz + sqrtdr * wr_ivz) * isdlpdr;
  double el = 0, er = 0, hroe = 0;
  if (gas) {
    el = wl_ipr * igm1 + 0.5 * wl_idn * (SQR(wl_ivx) + SQR(wl_ivy) + SQR(wl_ivz));
    er = wr_ipr * igm1 + 0.5 * wr_idn * (SQR(wr_ivx) + SQR(wr_ivy) + SQR(wr_ivz));
    hroe = ((el + wl_ipr) / sqrtdl + (er + wr_ipr) / sqrtdr) * isdlpdr;
  }
  double qa = 0, qb = 0, sl, sr;
  if (gas) {
    qa = sqrt(gamma * wl_ipr / wl_idn);
    qb = sqrt(gamma * wr_ipr / wr_idn);
    double a = hroe - 0.5 * (SQR(wroe_ivx) + SQR(wroe_ivy) + SQR(wroe_ivz));
    a = (a < 0.0) ? 0.0 : sqrt(gm1 * a);
    double sla = wroe_ivx - a;
    double slb = wl_ivx - qa;
    double sra = wroe_ivx + a;
    double srb = wr_ivx + qb;
    sl = dmin(sla, slb);
    sr = dmax(sra, srb);
  } else {
    sl = dmin(wroe_ivx, wl_ivx);
    sr = dmax(wroe_ivx, wr_ivx);
  }
  double bp = (sr > 0.0) ? sr : 1.0e-20;
  double bm = (sl < 0.0) ? sl : -1.0e-20;
  qa = wl_ivx - bm;
  qb = wr_ivx - bp;
  double fl_d = wl_idn * qa, fr_d = wr_idn * qb;
  double fl_mx = wl_idn * wl_ivx * qa, fr_mx = wr_idn * wr_ivx * qb;
  double fl_my = wl_idn * wl_ivy * qa, fr_my = wr_idn * wr_ivy * qb;
  double fl_mz = wl_idn * wl_ivz * qa, fr_mz = wr_idn * wr_ivz * qb;
  double fl_e = 0, fr_e = 0;
  if (gas) {
    fl_e = el * qa + wl_ipr * wl_ivx;
    fr_e = er * qb + wr_ipr * wr_ivx;
  }
  qa = 0.0;
  if (bp != bm) qa = 0.5 * (bp + bm) / (bp - bm);
  if (gas) out[6] = 0.5 * (wl_ipr + wr_ipr) + qa * (wl_ipr - wr_ipr);
  const double frho = 0.5 * (fl_d + fr_d) + qa * (fl_d - fr_d);
  out[0] = frho;
  out[1] = 0.5 * (fl_mx + fr_mx) + qa * (fl_mx - fr_mx);
  out[2] = 0.5 * (fl_my + fr_my) + qa * (fl_my - fr_my);
  out[3] = 0.5 * (fl_mz + fr_mz) + qa * (fl_mz - fr_mz);
  if (gas) {
    out[4] = 0.5 * (fl_e + fr_e) + qa * (fl_e - fr_e);
    out[5] = frho * ((frho >= 0.0) ? wl_ise : wr_ise);
    out[7] = frho / ((frho >= 0.0) ? wl_idn : wr_idn);
  }
}

/* llf.hpp:87-168 */
static void solve_llf(int fluid, double gm1, const double *wl, const double *wr,
                      double *out) {
  const int gas = (fluid == AO_GAS);
  const double wl_idn = wl[0], wl_ivx = wl[1], wl_ivy = wl[2], wl_ivz = wl[3];
  const double wr_idn = wr[0], wr_ivx = wr[1], wr_ivy = wr[2], wr_ivz = wr[3];
  double wl_ipr = 0, wr_ipr = 0, wl_ise = 0, wr_ise = 0, igm1 = 0, gamma = 0;
  if (gas) {
    wl_ipr = wl[4]; wl_ise = wl[5]; wr_ipr = wr[4]; wr_ise = wr[5];
    igm1 = 1.0 / gm1;
    gamma = gm1 + 1.0;
  }
  double qa = wl_idn * wl_ivx;
  double qb = wr_idn * wr_ivx;
  double fsum_d = qa + qb;
  double fsum_mx = qa * wl_ivx + qb * wr_ivx;
  double fsum_my = qa * wl_ivy + qb * wr_ivy;
  double fsum_mz = qa * wl_ivz + qb * wr_ivz;
  double el = 0, er = 0, fsum_e = 0;
  if (gas) {
    el = wl_ipr * igm1 + 0.5 * wl_idn * (SQR(wl_ivx) + SQR(wl_ivy) + SQR(wl_ivz));
    er = wr_ipr * igm1 + 0.5 * wr_idn * (SQR(wr_ivx) + SQR(wr_ivy) + SQR(wr_ivz));
    fsum_e = (el + wl_ipr) * wl_ivx + (er + wr_ipr) * wr_ivx;
  }
  double a;
  if (gas) {
    qa = sqrt(gamma * wl_ipr / wl_idn);
    qb = sqrt(gamma * wr_ipr / wr_idn);
    a = dmax((fabs(wl_ivx) + qa), (fabs(wr_ivx) + qb));
  } else {
    a = dmax(fabs(wl_ivx), fabs(wr_ivx));
  }
  double du_d = a * (wr_idn - wl_idn);
  double du_mx = a * (wr_idn * wr_ivx - wl_idn * wl_ivx);
  double du_my = a * (wr_idn * wr_ivy - wl_idn * wl_ivy);
  double du_mz = a * (wr_idn * wr_ivz - wl_idn * wl_ivz);
  double du_e = 0;
  if (gas) du_e = a * (er - el);
  if (gas) out[6] = 0.5 * (wl_ipr + wr_ipr);
  const double frho = 0.5 * (fsum_d - du_d);
  out[0] = frho;
  out[1] = 0.5 * (fsum_mx - du_mx);
  out[2] = 0.5 * (fsum_my - du_my);
  out[3] = 0.5 * (fsum_mz - du_mz);
  if (gas) {
    out[4] = 0.5 * (fsum_e - du_e);
    out[5] = frho * ((frho >= 0.0) ? wl_ise : wr_ise);
    out[7] = frho / ((frho >= 0.0) ? wl_idn : wr_idn);
  }
}

void ao_riemann(int solver, int fluid, double gm1, const double *wl, const double *wr,
                double *out) {
  if (solver == AO_HLLC)
    solve_hllc(gm1, wl, wr, out);
  else if (solver == AO_HLLE)
    solve_hlle(fluid, gm1, wl, wr, out);
  else
    solve_llf(fluid, gm1, wl, wr, out);
}

/* RiemannSolver::solve over one row of faces + ScaleMomentumFlux (fluid_fluxes.hpp:32-70) */
static void riemann_row(const ao_grid *g, const ao_fluid *f, int dir, int b, int k, int j,
                        int il, int iu, const double *wl, const double *wr, double *flux,
                        double *pflux, double *vface) {
  const int S = f->nspecies;
  const int gas = (f->fluid == AO_GAS);
  const int nvar = gas ? 6 * S : 4 * S;
  const int W = g->ni + 1;
  const double *xmin = g->xmin + 3 * b, *dx = g->dx + 3 * b;
  for (int n = 0; n < S; ++n) {
    const int IDN = n;
    const int ivx = S + (n * 3) + ((dir - 1));
    const int ivy = S + (n * 3) + ((dir - 1) + 1) % 3;
    const int ivz = S + (n * 3) + ((dir - 1) + 2) % 3;
    const int IPR = S * 4 + n, ISE = S * 5 + n;
    for (int i = il; i <= iu; ++i) {
      double l[6], r[6], out[8];
      l[0] = wl[(size_t)IDN * W + i]; r[0] = wr[(size_t)IDN * W + i];
      l[1] = wl[(size_t)ivx * W + i]; r[1] = wr[(size_t)ivx * W + i];
      l[2] = wl[(size_t)ivy * W + i]; r[2] = wr[(size_t)ivy * W + i];
      l[3] = wl[(size_t)ivz * W + i]; r[3] = wr[(size_t)ivz * W + i];
      if (gas) {
        l[4] = wl[(size_t)IPR * W + i]; r[4] = wr[(size_t)IPR * W + i];
        l[5] = wl[(size_t)ISE * W + i]; r[5] = wr[(size_t)ISE * W + i];
      }
      ao_riemann(f->riemann, f->fluid, f->gm1, l, r, out);
      flux[IDX(g, nvar, b, IDN, k, j, i)] = out[0];
      flux[IDX(g, nvar, b, ivx, k, j, i)] = out[1];
      flux[IDX(g, nvar, b, ivy, k, j, i)] = out[2];
      flux[IDX(g, nvar, b, ivz, k, j, i)] = out[3];
      if (gas) {
        flux[IDX(g, nvar, b, IPR, k, j, i)] = out[4];
        flux[IDX(g, nvar, b, ISE, k, j, i)] = out[5];
        pflux[IDX(g, S, b, n, k, j, i)] = out[6];
        vface[FIDX(g, S, b, n, k, j, i)] = out[7];
      }
    }
    /* ScaleMomentumFlux: fluid_fluxes.hpp:49-68 */
    if (g->geom != AO_CARTESIAN) {
      const int IVX = S + 3 * n, IVY = IVX + 1, IVZ = IVX + 2;
      for (int i = il; i <= iu; ++i) {
        bbox_t bb = make_bbox(xmin, dx, k, j, i);
        double xf[3];
        if (dir == 1) g_facecen1(g->geom, &bb, 0, xf);
        else if (dir == 2) g_facecen2(g->geom, &bb, 0, xf);
        else g_facecen3(g->geom, &bb, 0, xf);
        flux[IDX(g, nvar, b, IVX, k, j, i)] *= g_hx1(g->geom, xf[0], xf[1], xf[2]);
        flux[IDX(g, nvar, b, IVY, k, j, i)] *= g_hx2(g->geom, xf[0], xf[1], xf[2]);
        flux[IDX(g, nvar, b, IVZ, k, j, i)] *= g_hx3(g->geom, xf[0], xf[1], xf[2]);
      }
    }
  }
}

/* CalculateFluxesImpl: src/utils/fluxes/fluid_fluxes.hpp:76-213 */
void ao_calculate_fluxes(const ao_grid *g, const ao_fluid *f, int pcm, const double *prim,
                         double *flux1, double *flux2, double *flux3, double *pflux1,
                         double *pflux2, double *pflux3, double *vface1, double *vface2,
                         double *vface3) {
  const int S = f->nspecies;
  const int nvar = (f->fluid == AO_GAS) ? 6 * S : 4 * S;
  const int recon = pcm ? AO_PCM : f->recon; /* fluid_fluxes.hpp:225 */
  const int W = g->ni + 1;
  const size_t row = (size_t)nvar * W;
  const int nkr = g->ke - g->ks + 1, njr = g->je - g->js + 1;

  /* X1: fluid_fluxes.hpp:105-126 */
#pragma omp parallel
  {
    double *scr = (double *)malloc(sizeof(double) * row * 3);
    double *s1 = scr, *s2 = scr + row, *s3 = scr + 2 * row;
#pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < g->nb; ++b)
      for (int kj = 0; kj < nkr * njr; ++kj) {
        const int k = g->ks + kj / njr, j = g->js + kj % njr;
        recon_row(g, recon, 1, nvar, prim, b, k, j, g->is - 1, g->ie + 1, s1, s2);
        riemann_row(g, f, 1, b, k, j, g->is, g->ie + 1, s1, s2, flux1, pflux1, vface1);
      }
    /* X2: fluid_fluxes.hpp:129-168 */
    if (g->ndim > 1) {
#pragma omp for collapse(2) schedule(static)
      for (int b = 0; b < g->nb; ++b)
        for (int k = g->ks; k <= g->ke; ++k) {
          const int jl = g->js - 1, ju = g->je + 1;
          for (int j = jl; j <= ju; ++j) {
            double *wl = s1, *wl_jp1 = s2, *wr = s3;
            if ((j % 2) == 0) { wl = s2; wl_jp1 = s1; }
            recon_row(g, recon, 2, nvar, prim, b, k, j, g->is, g->ie, wl_jp1, wr);
            if (j > jl)
              riemann_row(g, f, 2, b, k, j, g->is, g->ie, wl, wr, flux2, pflux2, vface2);
          }
        }
    }
    /* X3: fluid_fluxes.hpp:171-210 */
    if (g->ndim > 2) {
#pragma omp for collapse(2) schedule(static)
      for (int b = 0; b < g->nb; ++b)
        for (int j = g->js; j <= g->je; ++j) {
          const int kl = g->ks - 1, ku = g->ke + 1;
          for (int k = kl; k <= ku; ++k) {
            double *wl = s1, *wl_kp1 = s2, *wr = s3;
            if ((k % 2) == 0) { wl = s2; wl_kp1 = s1; }
            recon_row(g, recon, 3, nvar, prim, b, k, j, g->is, g->ie, wl_kp1, wr);
            if (k > kl)
              riemann_row(g, f, 3, b, k, j, g->is, g->ie, wl, wr, flux3, pflux3, vface3);
          }
        }
    }
    free(scr);
  }
}

/* ===================================================================================== */
/* ApplyUpdate: src/utils/integrators/artemis_integrator.hpp:56-110                        */
/* ===================================================================================== */
void ao_apply_update(const ao_grid *g, int nvar, double *u0, const double *u1,
                     const double *flux1, const double *flux2, const double *flux3,
                     double gam0, double gam1, double beta_dt) {
  const int multi_d = g->ndim > 1, three_d = g->ndim > 2;
  const size_t sj = (size_t)g->ni, sk = (size_t)g->ni * g->nj;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double ax1[2] = {g_area1(g->geom, &bb, bb.x1[0]),
                                 g_area1(g->geom, &bb, bb.x1[1])};
          double ax2[2] = {0.0, 0.0}, ax3[2] = {0.0, 0.0};
          if (multi_d) {
            ax2[0] = g_area2(g->geom, &bb, bb.x2[0]);
            ax2[1] = g_area2(g->geom, &bb, bb.x2[1]);
          }
          if (three_d) {
            ax3[0] = g_area3(g->geom, &bb, bb.x3[0]);
            ax3[1] = g_area3(g->geom, &bb, bb.x3[1]);
          }
          const double vol = g_volume(g->geom, &bb);
          for (int n = 0; n < nvar; ++n) {
            const size_t c = IDX(g, nvar, b, n, k, j, i);
            double divf = (ax1[0] * flux1[c] - ax1[1] * flux1[c + 1]);
            if (multi_d) divf += (ax2[0] * flux2[c] - ax2[1] * flux2[c + sj]);
            if (three_d) divf += (ax3[0] * flux3[c] - ax3[1] * flux3[c + sk]);
            u0[c] = gam0 * u0[c] + gam1 * u1[c] + divf * beta_dt / vol;
          }
        }
}

/* DeepCopyConservedData: artemis_integrator.hpp:30-51 */
void ao_deep_copy(const ao_grid *g, int nvar, double *to, const double *from) {
  memcpy(to, from, sizeof(double) * (size_t)g->nb * nvar * g->nk * g->nj * g->ni);
}

/* ===================================================================================== */
/* FluxSourceImpl: src/utils/fluxes/fluid_fluxes.hpp:298-420                               */
/* NB the reference loops i over [is-2, ie+1]; rows outside the interior only touch ghost   */
/* conserved values that PrimToCons overwrites.  We restate the interior part and apply     */
/* the same update to ghost columns only where every operand exists (i >= 0, i+1 < ni).    */
/* ===================================================================================== */
void ao_flux_source(const ao_grid *g, const ao_fluid *f, const double *prim, double *cons,
                    const double *pflux1, const double *pflux2, const double *pflux3,
                    const double *vface1, const double *vface2, const double *vface3,
                    double omf, double dt) {
  const int S = f->nspecies;
  const int gas = (f->fluid == AO_GAS);
  const int nvar = gas ? 6 * S : 4 * S;
  const int multi_d = g->ndim >= 2, three_d = g->ndim == 3;
  const int x1dep = g_x1dep(g->geom);
  const int x2dep = g_x2dep(g->geom) && multi_d;
  const size_t sj = (size_t)g->ni, sk = (size_t)g->ni * g->nj;
  const size_t fsj = (size_t)g->fni, fsk = (size_t)g->fni * g->fnj;
  /* Dust::FluxSource early-out: src/dust/dust.cpp:311-312 */
  if (!gas && !(x1dep || x2dep)) return;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          double dhdx1[3] = {0, 0, 0}, dhdx2[3] = {0, 0, 0};
          if (x1dep) g_conn1(g->geom, &bb, dhdx1);
          if (x2dep) g_conn2(g->geom, &bb, dhdx2);
          const double ax1[2] = {g_area1(g->geom, &bb, bb.x1[0]),
                                 g_area1(g->geom, &bb, bb.x1[1])};
          double ax2[2] = {0.0, 0.0}, ax3[2] = {0.0, 0.0};
          if (multi_d) {
            ax2[0] = g_area2(g->geom, &bb, bb.x2[0]);
            ax2[1] = g_area2(g->geom, &bb, bb.x2[1]);
          }
          if (three_d) {
            ax3[0] = g_area3(g->geom, &bb, bb.x3[0]);
            ax3[1] = g_area3(g->geom, &bb, bb.x3[1]);
          }
          const double vol = g_volume(g->geom, &bb);
          const double dx[3] = {bb.x1[1] - bb.x1[0], bb.x2[1] - bb.x2[0],
                                bb.x3[1] - bb.x3[0]};
          const double xv[3] = {g_x1v(g->geom, &bb), g_x2v(g->geom, &bb),
                                g_x3v(g->geom, &bb)};
          double vf[3];
          g_rotation_velocity(g->geom, xv, omf, vf);
          for (int n = 0; n < S; ++n) {
            const size_t cmx = IDX(g, nvar, b, S + 3 * n + 0, k, j, i);
            const size_t cmy = IDX(g, nvar, b, S + 3 * n + 1, k, j, i);
            const size_t cmz = IDX(g, nvar, b, S + 3 * n + 2, k, j, i);
            if (gas) {
              const size_t ceg = IDX(g, nvar, b, 5 * S + n, k, j, i);
              const size_t p = IDX(g, S, b, n, k, j, i);
              const size_t v = FIDX(g, S, b, n, k, j, i);
              cons[cmx] += dt / dx[0] * (pflux1[p] - pflux1[p + 1]);
              cons[ceg] -= dt / vol * 0.5 * (pflux1[p] + pflux1[p + 1]) *
                           (ax1[1] * vface1[v + 1] - ax1[0] * vface1[v]);
              if (multi_d) {
                cons[cmy] += dt / dx[1] * (pflux2[p] - pflux2[p + sj]);
                cons[ceg] -= dt / vol * 0.5 * (pflux2[p] + pflux2[p + sj]) *
                             (ax2[1] * vface2[v + fsj] - ax2[0] * vface2[v]);
              }
              if (three_d) {
                cons[cmz] += dt / dx[2] * (pflux3[p] - pflux3[p + sk]);
                cons[ceg] -= dt / vol * 0.5 * (pflux3[p] + pflux3[p + sk]) *
                             (ax3[1] * vface3[v + fsk] - ax3[0] * vface3[v]);
              }
            }
            const double dens = prim[IDX(g, nvar, b, n, k, j, i)];
            const double rdt = dens * dt;
            const double vx = prim[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)];
            const double vy = prim[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)];
            const double vz = prim[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)];
            if (x1dep)
              cons[cmx] += rdt * (dhdx1[0] * SQR(vx + vf[0]) + dhdx1[1] * SQR(vy + vf[1]) +
                                  dhdx1[2] * SQR(vz + vf[2]));
            if (x2dep)
              cons[cmy] += rdt * (dhdx2[0] * SQR(vx + vf[0]) + dhdx2[1] * SQR(vy + vf[1]) +
                                  dhdx2[2] * SQR(vz + vf[2]));
          }
        }
}

/* ===================================================================================== */
/* src/derived/fill_derived.cpp + src/utils/artemis_utils.hpp:42-78                        */
/* ===================================================================================== */
/* SetAuxillaryFields: fill_derived.cpp:29-75 */
void ao_set_aux(const ao_grid *g, const ao_fluid *f, double *cons) {
  if (f->fluid != AO_GAS) return;
  const int S = f->nspecies, nvar = 6 * S;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double hx[3] = {g_hx1v(g->geom, &bb), g_hx2v(g->geom, &bb),
                                g_hx3v(g->geom, &bb)};
          for (int n = 0; n < S; ++n) {
            double u_d = cons[IDX(g, nvar, b, n, k, j, i)];
            u_d = (u_d > f->dfloor) ? u_d : f->dfloor;
            /* GetSpecificInternalEnergy: artemis_utils.hpp:50-66 */
            const double ud2 = dmax(cons[IDX(g, nvar, b, n, k, j, i)], f->dfloor);
            const double rv1 = cons[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)] / hx[0];
            const double rv2 = cons[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)] / hx[1];
            const double rv3 = cons[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)] / hx[2];
            const double ke = 0.5 * (SQR(rv1) + SQR(rv2) + SQR(rv3)) / ud2;
            const double e_cons = cons[IDX(g, nvar, b, 4 * S + n, k, j, i)];
            const double ue_cons = e_cons - ke;
            double *u_u = &cons[IDX(g, nvar, b, 5 * S + n, k, j, i)];
            double sie = (ue_cons > f->de_switch * e_cons) ? ue_cons / ud2 : *u_u / ud2;
            sie = dmax(sie, f->siefloor);
            *u_u = sie * u_d;
            const double uflr = f->siefloor * u_d;
            *u_u = (*u_u > uflr) ? *u_u : uflr;
          }
        }
}

/* ConsToPrim: fill_derived.cpp:81-167 (interior) */
void ao_cons_to_prim(const ao_grid *g, const ao_fluid *f, const double *cons, double *prim) {
  const int S = f->nspecies;
  const int gas = (f->fluid == AO_GAS);
  const int nvar = gas ? 6 * S : 4 * S;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double hx[3] = {g_hx1v(g->geom, &bb), g_hx2v(g->geom, &bb),
                                g_hx3v(g->geom, &bb)};
          for (int n = 0; n < S; ++n) {
            const double u_d = cons[IDX(g, nvar, b, n, k, j, i)];
            const double w_d = (u_d > f->dfloor) ? u_d : f->dfloor;
            prim[IDX(g, nvar, b, n, k, j, i)] = w_d;
            for (int d = 0; d < 3; ++d)
              prim[IDX(g, nvar, b, S + 3 * n + d, k, j, i)] =
                  cons[IDX(g, nvar, b, S + 3 * n + d, k, j, i)] / (w_d * hx[d]);
            if (gas) {
              const double w_s = cons[IDX(g, nvar, b, 5 * S + n, k, j, i)] / w_d;
              prim[IDX(g, nvar, b, 5 * S + n, k, j, i)] =
                  (w_s > f->siefloor) ? w_s : f->siefloor;
            }
          }
        }
}

/* PrimToCons: fill_derived.cpp:172-277 (entire domain, ghosts included) */
void ao_prim_to_cons(const ao_grid *g, const ao_fluid *f, double *prim, double *cons) {
  const int S = f->nspecies;
  const int gas = (f->fluid == AO_GAS);
  const int nvar = gas ? 6 * S : 4 * S;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = 0; k < g->nk; ++k)
      for (int j = 0; j < g->nj; ++j)
        for (int i = 0; i < g->ni; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double hx[3] = {g_hx1v(g->geom, &bb), g_hx2v(g->geom, &bb),
                                g_hx3v(g->geom, &bb)};
          for (int n = 0; n < S; ++n) {
            double *w_d = &prim[IDX(g, nvar, b, n, k, j, i)];
            *w_d = (*w_d > f->dfloor) ? *w_d : f->dfloor;
            const double u_d = *w_d;
            cons[IDX(g, nvar, b, n, k, j, i)] = u_d;
            const double vel1 = prim[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)];
            const double vel2 = prim[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)];
            const double vel3 = prim[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)];
            cons[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)] = *w_d * vel1 * hx[0];
            cons[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)] = *w_d * vel2 * hx[1];
            cons[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)] = *w_d * vel3 * hx[2];
            if (gas) {
              double *w_s = &prim[IDX(g, nvar, b, 5 * S + n, k, j, i)];
              *w_s = (*w_s > f->siefloor) ? *w_s : f->siefloor;
              const double u_u = *w_s * u_d;
              cons[IDX(g, nvar, b, 5 * S + n, k, j, i)] = u_u;
              /* singularity-eos 1.9.1 eos_ideal.hpp:91-95: MYMAX(0.0, gm1*rho*sie) */
              prim[IDX(g, nvar, b, 4 * S + n, k, j, i)] = dmax(0.0, f->gm1 * *w_d * *w_s);
              const double ke = 0.5 * *w_d * (SQR(vel1) + SQR(vel2) + SQR(vel3));
              cons[IDX(g, nvar, b, 4 * S + n, k, j, i)] = u_u + ke;
            }
          }
        }
}

/* Gas/Dust::EstimateTimestepMesh: src/gas/gas.cpp:391-468, src/dust/dust.cpp:238-276.
 * Returns cfl * min_dt (diffusion limits out of scope). */
double ao_estimate_dt(const ao_grid *g, const ao_fluid *f, const double *prim) {
  const int S = f->nspecies;
  const int gas = (f->fluid == AO_GAS);
  const int nvar = gas ? 6 * S : 4 * S;
  double min_dt = 1.79769313486231570815e+308; /* Big<Real>() */
#pragma omp parallel for collapse(2) schedule(static) reduction(min : min_dt)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          double dx[3];
          g_cell_widths(g->geom, &bb, dx);
          for (int n = 0; n < S; ++n) {
            double cs = 0.0;
            if (gas) {
              const double dens = prim[IDX(g, nvar, b, n, k, j, i)];
              const double sie = prim[IDX(g, nvar, b, 5 * S + n, k, j, i)];
              /* eos_ideal.hpp:136-140 */
              const double bulk = dmax(0.0, (f->gm1 + 1) * f->gm1 * dens * sie);
              cs = sqrt(bulk / dens);
            }
            double denom = 0.0;
            for (int d = 0; d < g->ndim; d++) {
              const double av = fabs(prim[IDX(g, nvar, b, S + 3 * n + d, k, j, i)]);
              if (gas) {
                const double ss = av + cs;
                denom += ss / dx[d];
              } else {
                denom += av / dx[d];
              }
            }
            min_dt = dmin(min_dt, 1.0 / denom);
          }
        }
  return f->cfl * min_dt;
}

/* ===================================================================================== */
/* Ghost exchange (same level) + physical BCs.                                             */
/* Index ranges: P:bvals/comms/bnd_info.cpp:152-213 (same-level branch); pack/unpack        */
/* P:bvals/comms/boundary_communication.cpp:95-140, 273-334; outflow/reflect                */
/* P:bvals/boundary_conditions_generic.hpp:178-256 in face order ix1,ox1,ix2,ox2,ix3,ox3.   */
/* ===================================================================================== */
static void range_send(int off, int s, int e, int ng, int *lo, int *hi) {
  if (off == 0) { *lo = s; *hi = e; }
  else if (off > 0) { *lo = e - ng + 1; *hi = e; }
  else { *lo = s; *hi = s + ng - 1; }
}
static void range_recv(int off, int s, int e, int ng, int *lo, int *hi) {
  if (off == 0) { *lo = s; *hi = e; }
  else if (off > 0) { *lo = e + 1; *hi = e + ng; }
  else { *lo = s - ng; *hi = s - 1; }
}

/* phases: bit 0 = neighbour -> ghost copies, bit 1 = physical boundaries.  Faces flagged
 * AO_BC_NONE belong to another rank: neither copied across nor filled (multi-rank tests). */
void ao_exchange_ghosts_phase(const ao_grid *g, int nbx, int nby, int nbz, const int *bc,
                              int nvar, double *a, int nv, const int *vars,
                              const int *vec_dir, int phases) {
  ao_exchange_ghosts_ic(g, nbx, nby, nbz, bc, nvar, a, nv, vars, vec_dir, phases, 0);
}

void ao_exchange_ghosts_ic(const ao_grid *g, int nbx, int nby, int nbz, const int *bc, int nvar,
                           double *a, int nv, const int *vars, const int *vec_dir, int phases,
                           const double *ic) {
  const int ng = g->ng;
  const int nbd[3] = {nbx, nby, nbz};
  const int ox3 = g->ndim > 2 ? 1 : 0, ox2 = g->ndim > 1 ? 1 : 0;
  /* 1. neighbour -> ghost copies (receiver-driven; all sources are interior cells) */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < ((phases & 1) ? g->nb : 0); ++b) {
    const int lb[3] = {b % nbx, (b / nbx) % nby, b / (nbx * nby)};
    for (int o3 = -ox3; o3 <= ox3; ++o3)
      for (int o2 = -ox2; o2 <= ox2; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          if (!o1 && !o2 && !o3) continue;
          const int off[3] = {o1, o2, o3};
          int ln[3], ok = 1;
          for (int d = 0; d < 3; ++d) {
            ln[d] = lb[d] + off[d];
            if (ln[d] < 0 || ln[d] >= nbd[d]) {
              if (bc[2 * d + (off[d] > 0)] == AO_BC_PERIODIC)
                ln[d] = (ln[d] + nbd[d]) % nbd[d];
              else
                ok = 0;
            }
          }
          if (!ok) continue;
          const int nbr = ln[0] + nbx * (ln[1] + nby * ln[2]);
          int rl[3], rh[3], sl[3], sh[3];
          range_recv(o1, g->is, g->ie, ng, &rl[0], &rh[0]);
          range_recv(o2, g->js, g->je, ng, &rl[1], &rh[1]);
          range_recv(o3, g->ks, g->ke, ng, &rl[2], &rh[2]);
          /* the sender sees this block at offset -off */
          range_send(-o1, g->is, g->ie, ng, &sl[0], &sh[0]);
          range_send(-o2, g->js, g->je, ng, &sl[1], &sh[1]);
          range_send(-o3, g->ks, g->ke, ng, &sl[2], &sh[2]);
          (void)sh;
          for (int v = 0; v < nv; ++v)
            for (int k = rl[2]; k <= rh[2]; ++k)
              for (int j = rl[1]; j <= rh[1]; ++j)
                for (int i = rl[0]; i <= rh[0]; ++i)
                  a[IDX(g, nvar, b, vars[v], k, j, i)] =
                      a[IDX(g, nvar, nbr, vars[v], sl[2] + (k - rl[2]), sl[1] + (j - rl[1]),
                            sl[0] + (i - rl[0]))];
        }
  }
  /* 2. physical boundaries */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < ((phases & 2) ? g->nb : 0); ++b) {
    const int lb[3] = {b % nbx, (b / nbx) % nby, b / (nbx * nby)};
    const int s[3] = {g->is, g->js, g->ks}, e[3] = {g->ie, g->je, g->ke};
    const int nt[3] = {g->ni, g->nj, g->nk};
    for (int face = 0; face < 2 * g->ndim; ++face) {
      const int d = face / 2, outer = face % 2;
      if (bc[face] == AO_BC_PERIODIC || bc[face] == AO_BC_NONE || bc[face] > AO_BC_IC) continue;
      if (outer ? (lb[d] != nbd[d] - 1) : (lb[d] != 0)) continue;
      const int ref = outer ? e[d] : s[d];
      const int offset = 2 * ref + (outer ? 1 : -1);
      int lo[3] = {0, 0, 0}, hi[3] = {nt[0] - 1, nt[1] - 1, nt[2] - 1};
      if (outer) { lo[d] = e[d] + 1; hi[d] = nt[d] - 1; }
      else { lo[d] = 0; hi[d] = s[d] - 1; }
      if (bc[face] == AO_BC_IC) { /* src/pgen/disk.hpp:595-633: profile at the zone itself */
        for (int v = 0; v < (ic ? nv : 0); ++v)
          for (int k = lo[2]; k <= hi[2]; ++k)
            for (int j = lo[1]; j <= hi[1]; ++j)
              for (int i = lo[0]; i <= hi[0]; ++i)
                a[IDX(g, nvar, b, vars[v], k, j, i)] = ic[IDX(g, nvar, b, vars[v], k, j, i)];
        continue;
      }
      for (int v = 0; v < nv; ++v) {
        const double sgn = (bc[face] == AO_BC_REFLECT && vec_dir[v] == d + 1) ? -1.0 : 1.0;
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
              int c[3] = {i, j, k};
              c[d] = (bc[face] == AO_BC_REFLECT) ? offset - c[d] : ref;
              a[IDX(g, nvar, b, vars[v], k, j, i)] =
                  sgn * a[IDX(g, nvar, b, vars[v], c[2], c[1], c[0])];
            }
      }
    }
  }
}

/* The shearing-box user boundary conditions of the strat / ssheet problem generators, applied to
 * ONE block array a[nvar][nk][nj][ni] (fine arrays or a coarse buffer; xmin / dx are those of
 * the index space, UniformCartesian::Xf(idx) = xmin + idx * dx).  s / e: interior bounds of the
 * three directions.  Every ghost zone of `face` over the FULL transverse extent, like
 * MeshBlock::par_for_bndry over IndexDomain::inner_x1 ... outer_x3.
 *   type AO_BC_EXTRAP, x1 faces: strat::ExtrapInnerX1 / ExtrapOuterX1  src/pgen/strat.hpp:154-297
 *   type AO_BC_INFLOW, x2 faces: strat::ShearInnerX2 / ShearOuterX2    src/pgen/strat.hpp:299-485
 *   type AO_BC_EXTRAP, x3 faces: strat::ExtrapInnerX3 / ExtrapOuterX3  src/pgen/strat.hpp:487-663
 * The gas branch writes gas::prim::{density, velocity, sie}(0) only; the dust branch every
 * species.  Returns 0, or 1 for a (type, face) pair the reference does not register
 * (src/pgen/problem_modifier.hpp:114-127). */
int ao_strat_bc(int geom, const double *xmin, const double *dx, int ni, int nj, int nk,
                const int *s, const int *e, int fluid, int S, double *a, int face, int type,
                double q, double om0) {
  const int d = face / 2, outer = face % 2;
  if (type == AO_BC_EXTRAP ? d == 1 : (type != AO_BC_INFLOW || d != 1)) return 1;
  const int nt[3] = {ni, nj, nk};
  int lo[3] = {0, 0, 0}, hi[3] = {ni - 1, nj - 1, nk - 1};
  if (outer) lo[d] = e[d] + 1; else hi[d] = s[d] - 1;
  const int ref = outer ? e[d] : s[d];        /* is / ie, js / je, ks / ke */
  const int ref1 = outer ? ref - 1 : ref + 1; /* is + 1 / ie - 1, ...       */
  const size_t cells = (size_t)ni * nj * nk;
  const int nsp = fluid == AO_GAS ? 1 : S;
  (void)nt;
#define SA(var, k, j, i) a[(size_t)(var) * cells + ((size_t)(k) * nj + (j)) * ni + (i)]
  for (int n = 0; n < nsp; ++n) {
    const int vd = n, v1 = S + 3 * n, v2 = v1 + 1, v3 = v1 + 2, ve = 5 * S + n;
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          int c0[3] = {i, j, k}, c1[3] = {i, j, k};
          c0[d] = ref; c1[d] = ref1;
          const bbox_t bb = make_bbox(xmin, dx, k, j, i);
          const bbox_t b0 = make_bbox(xmin, dx, c0[2], c0[1], c0[0]);
          const bbox_t b1 = make_bbox(xmin, dx, c1[2], c1[1], c1[0]);
          const double gv1 = SA(v1, c0[2], c0[1], c0[0]);
          const double gv2 = SA(v2, c0[2], c0[1], c0[0]);
          const double gv3 = SA(v3, c0[2], c0[1], c0[0]);
          const double gd = SA(vd, c0[2], c0[1], c0[0]);
          double vx1 = gv1, vx2 = gv2, vx3 = gv3, dens = gd;
          if (type == AO_BC_EXTRAP && d == 0) {
            const double x0 = g_x1v(geom, &b0), x1 = g_x1v(geom, &b1);
            const double ddx = outer ? x0 - x1 : x1 - x0;
            const double x = g_x1v(geom, &bb);
            const double gv2n = SA(v2, c1[2], c1[1], c1[0]);
            if (outer) {
              vx1 = (gv1 < 0.0) ? 0.0 : gv1;
              vx2 = gv2 + (gv2 - gv2n) * (x - x0) / ddx;
            } else {
              vx1 = (gv1 > 0.0) ? 0.0 : gv1;
              vx2 = gv2 + (gv2n - gv2) * (x - x0) / ddx;
            }
          } else if (type == AO_BC_EXTRAP) { /* x3 */
            const double z = g_x3v(geom, &bb);
            const double z0 = g_x3v(geom, &b0), z1 = g_x3v(geom, &b1);
            const double dz = outer ? z0 - z1 : z1 - z0;
            const double gdn = SA(vd, c1[2], c1[1], c1[0]);
            const double drho = outer ? gd / gdn : gdn / gd;
            vx3 = outer ? ((gv3 < 0.0) ? 0.0 : gv3) : ((gv3 > 0.0) ? 0.0 : gv3);
            dens = gd * pow(drho, (z - z0) / dz);
          } else { /* inflow, x2 */
            const double x = g_x1v(geom, &bb);
            const double xf = bb.x1[0];
            const double vy0 = -q * om0 * x;
            if (outer) vx2 = (xf < 0) ? ((gv2 < 0.0) ? 0.0 : gv2) : vy0;
            else vx2 = (xf >= 0) ? ((gv2 > 0.) ? 0.0 : gv2) : vy0;
          }
          const double sie = fluid == AO_GAS ? SA(ve, c0[2], c0[1], c0[0]) : 0.0;
          SA(v1, k, j, i) = vx1;
          SA(v2, k, j, i) = vx2;
          SA(v3, k, j, i) = vx3;
          SA(vd, k, j, i) = dens;
          if (fluid == AO_GAS) SA(ve, k, j, i) = sie;
        }
  }
#undef SA
  return 0;
}

/* ao_exchange_ghosts_ic with the shearing-box user conditions: faces flagged AO_BC_EXTRAP /
 * AO_BC_INFLOW run ao_strat_bc on the whole fluid (a = its primitive pack, S species), every
 * other face the generic condition, all in Parthenon's x1 -> x2 -> x3 face order. */
void ao_exchange_ghosts_user(const ao_grid *g, int nbx, int nby, int nbz, const int *bc, int nvar,
                             double *a, int nv, const int *vars, const int *vec_dir,
                             const double *ic, int fluid, int S, double q, double om0) {
  ao_exchange_ghosts_ic(g, nbx, nby, nbz, bc, nvar, a, nv, vars, vec_dir, 1, ic);
  const int nbd[3] = {nbx, nby, nbz};
  const int s[3] = {g->is, g->js, g->ks}, e[3] = {g->ie, g->je, g->ke};
  for (int face = 0; face < 2 * g->ndim; ++face) {
    const int d = face / 2, outer = face % 2;
    if (bc[face] == AO_BC_EXTRAP || bc[face] == AO_BC_INFLOW) {
#pragma omp parallel for schedule(static)
      for (int b = 0; b < g->nb; ++b) {
        const int lb[3] = {b % nbx, (b / nbx) % nby, b / (nbx * nby)};
        if (outer ? (lb[d] != nbd[d] - 1) : (lb[d] != 0)) continue;
        ao_strat_bc(g->geom, g->xmin + 3 * b, g->dx + 3 * b, g->ni, g->nj, g->nk, s, e, fluid, S,
                    a + (size_t)b * nvar * g->ni * g->nj * g->nk, face, bc[face], q, om0);
      }
    } else { /* one generic face at a time keeps the order */
      int one[6] = {AO_BC_NONE, AO_BC_NONE, AO_BC_NONE, AO_BC_NONE, AO_BC_NONE, AO_BC_NONE};
      one[face] = bc[face];
      ao_exchange_ghosts_ic(g, nbx, nby, nbz, one, nvar, a, nv, vars, vec_dir, 2, ic);
    }
  }
}

void ao_exchange_ghosts(const ao_grid *g, int nbx, int nby, int nbz, const int *bc,
                        int nvar, double *a, int nv, const int *vars, const int *vec_dir) {
  ao_exchange_ghosts_phase(g, nbx, nby, nbz, bc, nvar, a, nv, vars, vec_dir, 3);
}

/* ===================================================================================== */
/* Multilevel operators.                                               */
/* ===================================================================================== */
/* P:coordinates/uniform_cartesian.hpp:41-55 UniformCartesian(src, coarsen = 2): the coarse
 * buffer keeps ng ghost cells of twice the width in every active direction. */
static void coarse_coords(const ao_refine_geom *r, double cxmin[3], double cdx[3]) {
  const int act[3] = {1, r->ndim > 1, r->ndim > 2};
  for (int d = 0; d < 3; ++d) {
    const int istart = act[d] ? r->ng : 0;
    const int coarsen = 2;
    cdx[d] = r->dx[d];
    cxmin[d] = r->xmin[d];
    cxmin[d] += istart * cdx[d] * (1 - coarsen);
    cdx[d] *= (d == 0 ? coarsen : (istart > 0 ? coarsen : 1));
  }
}
#define RIDX(nk_, nj_, ni_, n, k, j, i) ((((size_t)(n) * (nk_) + (k)) * (nj_) + (j)) * (ni_) + (i))

/* restriction.hpp:41-114 (el = CC): coarse = sum(vol * fine) / sum(vol) over the 2^ndim fine
 * cells, both sums in the reference's pairing ((000+010)+(001+011))+((100+110)+(101+111)). */
/* the same operator on a face-centred (flux) field, el = 1..3 = x1 / x2 / x3 faces: the
 * average runs over the 2^(ndim-1) fine faces of the coarse face, weighted by the LOWER face
 * area of the fine cell (restriction.hpp:57-63, 88-95) */
void ao_restrict_average_face(const ao_refine_geom *r, int nvar, const double *fine,
                              double *coarse, const int *box, int el) {
  const int inc1 = r->ndim > 0 && el != 1, inc2 = r->ndim > 1 && el != 2,
            inc3 = r->ndim > 2 && el != 3;
  for (int n = 0; n < nvar; ++n)
    for (int ck = box[4]; ck <= box[5]; ++ck)
      for (int cj = box[2]; cj <= box[3]; ++cj)
        for (int ci = box[0]; ci <= box[1]; ++ci) {
          const int i = (r->ndim > 0) ? (ci - r->cib_s) * 2 + r->ib_s : r->ib_s;
          const int j = (r->ndim > 1) ? (cj - r->cjb_s) * 2 + r->jb_s : r->jb_s;
          const int k = (r->ndim > 2) ? (ck - r->ckb_s) * 2 + r->kb_s : r->kb_s;
          double vol[2][2][2], terms[2][2][2];
          for (int ok = 0; ok < 2; ++ok)
            for (int oj = 0; oj < 2; ++oj)
              for (int oi = 0; oi < 2; ++oi) vol[ok][oj][oi] = terms[ok][oj][oi] = 0;
          for (int ok = 0; ok < 1 + inc3; ++ok)
            for (int oj = 0; oj < 1 + inc2; ++oj)
              for (int oi = 0; oi < 1 + inc1; ++oi) {
                const bbox_t b = make_bbox(r->xmin, r->dx, k + ok, j + oj, i + oi);
                vol[ok][oj][oi] = el == 1 ? g_area1(r->geom, &b, b.x1[0])
                                  : el == 2 ? g_area2(r->geom, &b, b.x2[0])
                                            : g_area3(r->geom, &b, b.x3[0]);
                terms[ok][oj][oi] =
                    vol[ok][oj][oi] * fine[RIDX(r->nk, r->nj, r->ni, n, k + ok, j + oj, i + oi)];
              }
          const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                              ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
          coarse[RIDX(r->cnk, r->cnj, r->cni, n, ck, cj, ci)] =
              (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
               ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
              tvol;
        }
}

void ao_restrict_average(const ao_refine_geom *r, int nvar, const double *fine, double *coarse,
                         const int *box) {
  const int inc1 = r->ndim > 0, inc2 = r->ndim > 1, inc3 = r->ndim > 2;
  for (int n = 0; n < nvar; ++n)
    for (int ck = box[4]; ck <= box[5]; ++ck)
      for (int cj = box[2]; cj <= box[3]; ++cj)
        for (int ci = box[0]; ci <= box[1]; ++ci) {
          const int i = inc1 ? (ci - r->cib_s) * 2 + r->ib_s : r->ib_s;
          const int j = inc2 ? (cj - r->cjb_s) * 2 + r->jb_s : r->jb_s;
          const int k = inc3 ? (ck - r->ckb_s) * 2 + r->kb_s : r->kb_s;
          double vol[2][2][2], terms[2][2][2];
          for (int ok = 0; ok < 2; ++ok)
            for (int oj = 0; oj < 2; ++oj)
              for (int oi = 0; oi < 2; ++oi) vol[ok][oj][oi] = terms[ok][oj][oi] = 0;
          for (int ok = 0; ok < 1 + inc3; ++ok)
            for (int oj = 0; oj < 1 + inc2; ++oj)
              for (int oi = 0; oi < 1 + inc1; ++oi) {
                const bbox_t b = make_bbox(r->xmin, r->dx, k + ok, j + oj, i + oi);
                vol[ok][oj][oi] = g_volume(r->geom, &b);
                terms[ok][oj][oi] =
                    vol[ok][oj][oi] * fine[RIDX(r->nk, r->nj, r->ni, n, k + ok, j + oj, i + oi)];
              }
          const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                              ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
          coarse[RIDX(r->cnk, r->cnj, r->cni, n, ck, cj, ci)] =
              (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
               ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
              tvol;
        }
}

/* prolongation.hpp:72-79 GradMinMod, SIGN from P:config.hpp.in:86 */
static inline double grad_minmod(double fc, double fm, double fp, double dxm, double dxp) {
  const double gxm = (fc - fm) / dxm;
  const double gxp = (fp - fc) / dxp;
  const double sm = (gxm < 0.0) ? -1.0 : 1.0, sp = (gxp < 0.0) ? -1.0 : 1.0;
  return 0.5 * (sm + sp) * dmin(fabs(gxm), fabs(gxp));
}
/* prolongation.hpp:39-67 GetGridSpacings<C, DIM>: centroid distances on both levels */
static inline void grid_spacings(int geom, int dim, const double *xmin, const double *dx,
                                 const double *cxmin, const double *cdx, int k, int j, int i,
                                 int fk, int fj, int fi, double *dxm, double *dxp, double *dxfm,
                                 double *dxfp) {
  const int o3 = dim == 3, o2 = dim == 2, o1 = dim == 1;
  const bbox_t cc = make_bbox(cxmin, cdx, k, j, i);
  const bbox_t cm = make_bbox(cxmin, cdx, k - o3, j - o2, i - o1);
  const bbox_t cp = make_bbox(cxmin, cdx, k + o3, j + o2, i + o1);
  const bbox_t fm = make_bbox(xmin, dx, fk, fj, fi);
  const bbox_t fp = make_bbox(xmin, dx, fk + o3, fj + o2, fi + o1);
  double xm, xc, xp, fxm, fxp;
  if (dim == 1) {
    xm = g_x1v(geom, &cm); xc = g_x1v(geom, &cc); xp = g_x1v(geom, &cp);
    fxm = g_x1v(geom, &fm); fxp = g_x1v(geom, &fp);
  } else if (dim == 2) {
    xm = g_x2v(geom, &cm); xc = g_x2v(geom, &cc); xp = g_x2v(geom, &cp);
    fxm = g_x2v(geom, &fm); fxp = g_x2v(geom, &fp);
  } else {
    xm = g_x3v(geom, &cm); xc = g_x3v(geom, &cc); xp = g_x3v(geom, &cp);
    fxm = g_x3v(geom, &fm); fxp = g_x3v(geom, &fp);
  }
  *dxm = xc - xm;
  *dxp = xp - xc;
  *dxfm = xc - fxm;
  *dxfp = fxp - xc;
}

/* prolongation.hpp:82-184 ProlongateSharedMinMod<GEOM>::Do<DIM, CC> */
void ao_prolongate_minmod(const ao_refine_geom *r, int nvar, const double *coarse, double *fine,
                          const int *box) {
  const int inc1 = r->ndim > 0, inc2 = r->ndim > 1, inc3 = r->ndim > 2;
  double cxmin[3], cdx[3];
  coarse_coords(r, cxmin, cdx);
#define CO(n, k, j, i) coarse[RIDX(r->cnk, r->cnj, r->cni, n, k, j, i)]
#define FI(n, k, j, i) fine[RIDX(r->nk, r->nj, r->ni, n, k, j, i)]
  for (int n = 0; n < nvar; ++n)
    for (int k = box[4]; k <= box[5]; ++k)
      for (int j = box[2]; j <= box[3]; ++j)
        for (int i = box[0]; i <= box[1]; ++i) {
          const int fi = inc1 ? (i - r->cib_s) * 2 + r->ib_s : r->ib_s;
          const int fj = inc2 ? (j - r->cjb_s) * 2 + r->jb_s : r->jb_s;
          const int fk = inc3 ? (k - r->ckb_s) * 2 + r->kb_s : r->kb_s;
          const double fc = CO(n, k, j, i);
          double dx1fm = 0, dx1fp = 0, gx1m = 0, gx1p = 0;
          if (inc1) {
            double dx1m, dx1p;
            grid_spacings(r->geom, 1, r->xmin, r->dx, cxmin, cdx, k, j, i, fk, fj, fi, &dx1m,
                          &dx1p, &dx1fm, &dx1fp);
            const double g = grad_minmod(fc, CO(n, k, j, i - 1), CO(n, k, j, i + 1), dx1m, dx1p);
            gx1m = g; gx1p = g;
          }
          double dx2fm = 0, dx2fp = 0, gx2m = 0, gx2p = 0;
          if (inc2) {
            double dx2m, dx2p;
            grid_spacings(r->geom, 2, r->xmin, r->dx, cxmin, cdx, k, j, i, fk, fj, fi, &dx2m,
                          &dx2p, &dx2fm, &dx2fp);
            const double g = grad_minmod(fc, CO(n, k, j - 1, i), CO(n, k, j + 1, i), dx2m, dx2p);
            gx2m = g; gx2p = g;
          }
          double dx3fm = 0, dx3fp = 0, gx3m = 0, gx3p = 0;
          if (inc3) {
            double dx3m, dx3p;
            grid_spacings(r->geom, 3, r->xmin, r->dx, cxmin, cdx, k, j, i, fk, fj, fi, &dx3m,
                          &dx3p, &dx3fm, &dx3fp);
            const double g = grad_minmod(fc, CO(n, k - 1, j, i), CO(n, k + 1, j, i), dx3m, dx3p);
            gx3m = g; gx3p = g;
          }
          FI(n, fk, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
          if (inc1) FI(n, fk, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
          if (inc2) FI(n, fk, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
          if (inc2 && inc1)
            FI(n, fk, fj + 1, fi + 1) = fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
          if (inc3) FI(n, fk + 1, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
          if (inc3 && inc1)
            FI(n, fk + 1, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
          if (inc3 && inc2)
            FI(n, fk + 1, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
          if (inc3 && inc2 && inc1)
            FI(n, fk + 1, fj + 1, fi + 1) = fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
        }
#undef CO
#undef FI
}

/* ==========================================================================================
 * Pointwise source terms between FluxSource and SetAuxillaryFields (SURVEY 8f rank 1,
 * src/artemis_driver.cpp:217-248).  They read the STAGE-START primitives (C2P has not run yet)
 * and add to the conserved state of interior zones.
 * ========================================================================================== */

/* Gravity::UniformGravity<GEOM>, src/gravity/uniform.cpp:28-90 */
void ao_uniform_gravity(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                        const ao_fluid *dust, const double *dprim, double *dcons, double dt,
                        double gx1, double gx2, double gx3) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double hx[3] = {g_hx1v(g->geom, &bb), g_hx2v(g->geom, &bb), g_hx3v(g->geom, &bb)};
          const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
          if (gas) {
            const int S = gas->nspecies;
            const double *w = gprim + (size_t)b * 6 * S * cells;
            double *u = gcons + (size_t)b * 6 * S * cells;
            for (int n = 0; n < S; ++n) {
              const double rdt = dt * w[(size_t)n * cells + o];
              u[(size_t)(S + 3 * n + 0) * cells + o] += rdt * hx[0] * gx1;
              u[(size_t)(S + 3 * n + 1) * cells + o] += rdt * hx[1] * gx2;
              u[(size_t)(S + 3 * n + 2) * cells + o] += rdt * hx[2] * gx3;
              u[(size_t)(4 * S + n) * cells + o] +=
                  rdt * (w[(size_t)(S + 3 * n + 0) * cells + o] * gx1 +
                         w[(size_t)(S + 3 * n + 1) * cells + o] * gx2 +
                         w[(size_t)(S + 3 * n + 2) * cells + o] * gx3);
            }
          }
          if (dust) {
            const int S = dust->nspecies;
            const double *w = dprim + (size_t)b * 4 * S * cells;
            double *u = dcons + (size_t)b * 4 * S * cells;
            for (int n = 0; n < S; ++n) {
              const double rdt = dt * w[(size_t)n * cells + o];
              u[(size_t)(S + 3 * n + 0) * cells + o] += rdt * hx[0] * gx1;
              u[(size_t)(S + 3 * n + 1) * cells + o] += rdt * hx[1] * gx2;
              u[(size_t)(S + 3 * n + 2) * cells + o] += rdt * hx[2] * gx3;
            }
          }
        }
}

/* ---- coordinate-system conversions used by the point-mass and rotating-frame sources ------
 * ConvertTo{Cart,Cyl,Sph}WithVec(x) = { Convert*Coords*(x), rows ex1, ex2, ex3 of Convert*Vec*(x) }
 * (geometry.hpp:438-484).  e[r][c]: component c of row ex{r+1}. */
#define AO_FUZZ 1e-99 /* Fuzz<Real>(), src/artemis.hpp:112-118 */
static inline void e_set(double e[3][3], double a0, double a1, double a2, double b0, double b1,
                         double b2, double c0, double c1, double c2) {
  e[0][0] = a0; e[0][1] = a1; e[0][2] = a2;
  e[1][0] = b0; e[1][1] = b1; e[1][2] = b2;
  e[2][0] = c0; e[2][1] = c1; e[2][2] = c2;
}
/* geometry.hpp:246-260, cylindrical.hpp:95-109, spherical.hpp:171-189 / 375-393 / 527-545,
 * axisymmetric.hpp:98-113 */
static void g_to_cart(int geom, const double xi[3], double xo[3], double e[3][3]) {
  switch (geom) {
  case AO_CYLINDRICAL: {
    const double cp = cos(xi[1]), sp = sin(xi[1]);
    e_set(e, cp, sp, 0.0, -sp, cp, 0.0, 0.0, 0.0, 1.0);
    xo[0] = xi[0] * cp; xo[1] = xi[0] * sp; xo[2] = xi[2];
    break;
  }
  case AO_AXISYMMETRIC: {
    const double cp = cos(xi[2]), sp = sin(xi[2]);
    e_set(e, cp, 0.0, sp, -sp, 0.0, cp, 0.0, 1.0, 0.0);
    xo[0] = xi[0] * cp; xo[1] = xi[0] * sp; xo[2] = xi[1];
    break;
  }
  case AO_SPHERICAL3D:
  case AO_SPHERICAL2D:
  case AO_SPHERICAL1D: {
    const double cp = (geom == AO_SPHERICAL3D) ? cos(xi[2]) : 1.0;
    const double sp = (geom == AO_SPHERICAL3D) ? sin(xi[2]) : 0.0;
    const double ct = (geom == AO_SPHERICAL1D) ? 0.0 : cos(xi[1]);
    const double st = (geom == AO_SPHERICAL1D) ? 1.0 : sin(xi[1]);
    e_set(e, st * cp, st * sp, ct, ct * cp, ct * sp, -st, -sp, cp, 0.0);
    xo[0] = xi[0] * st * cp; xo[1] = xi[0] * st * sp; xo[2] = xi[0] * ct;
    break;
  }
  default:
    e_set(e, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0);
    xo[0] = xi[0]; xo[1] = xi[1]; xo[2] = xi[2];
  }
}
/* geometry.hpp:284-302, cylindrical.hpp:127-137, spherical.hpp:201-222 / 405-426 / 557-578,
 * axisymmetric.hpp:134-146 */
static void g_to_cyl(int geom, const double xi[3], double xo[3], double e[3][3]) {
  switch (geom) {
  case AO_CYLINDRICAL:
    e_set(e, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0);
    xo[0] = xi[0]; xo[1] = xi[1]; xo[2] = xi[2];
    break;
  case AO_AXISYMMETRIC:
    e_set(e, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0);
    xo[0] = xi[0]; xo[1] = xi[2]; xo[2] = xi[1];
    break;
  case AO_SPHERICAL3D:
  case AO_SPHERICAL2D:
  case AO_SPHERICAL1D: {
    const double ct = (geom == AO_SPHERICAL1D) ? 0.0 : cos(xi[1]);
    const double st = (geom == AO_SPHERICAL1D) ? 1.0 : sin(xi[1]);
    e_set(e, st, 0.0, ct, ct, 0.0, -st, 0.0, 1.0, 0.0);
    xo[0] = xi[0] * st; xo[1] = (geom == AO_SPHERICAL3D) ? xi[2] : 0.0; xo[2] = xi[0] * ct;
    break;
  }
  default: { /* Cartesian */
    const double R = sqrt(xi[0] * xi[0] + xi[1] * xi[1]);
    const double cp = xi[0] / (R + AO_FUZZ), sp = xi[1] / (R + AO_FUZZ);
    e_set(e, cp, -sp, 0.0, sp, cp, 0.0, 0.0, 0.0, 1.0);
    xo[0] = R; xo[1] = atan2(sp, cp); xo[2] = xi[2];
  }
  }
}
/* RFWeights: +-(<R^2>_face - <R^2>) of the cylindrical radius; geometry.hpp:228-232,
 * cylindrical.hpp:88-93, axisymmetric.hpp:91-96, spherical.hpp:148-169 / 352-373 / 514-525 */
static void g_rf_weights(int geom, const bbox_t *b, double bx[3][2]) {
  for (int d = 0; d < 3; ++d) bx[d][0] = bx[d][1] = 0.0;
  if (geom == AO_CYLINDRICAL || geom == AO_AXISYMMETRIC) {
    const double ans = 0.5 * (b->x1[0] + b->x1[1]) * (b->x1[1] - b->x1[0]);
    bx[0][0] = bx[0][1] = ans;
  } else if (geom == AO_SPHERICAL3D || geom == AO_SPHERICAL2D) {
    const double rv = g_x1v(geom, b);
    const double stv = sin(g_x2v(geom, b));
    const double rf = rface_avg(b);
    const double r2cyl = (rv * stv) * (rv * stv);
    bx[0][0] = r2cyl - (b->x1[0] * stv) * (b->x1[0] * stv);
    bx[0][1] = (b->x1[1] * stv) * (b->x1[1] * stv) - r2cyl;
    bx[1][0] = r2cyl - (rf * sin(b->x2[0])) * (rf * sin(b->x2[0]));
    bx[1][1] = (rf * sin(b->x2[1])) * (rf * sin(b->x2[1])) - r2cyl;
  } else if (geom == AO_SPHERICAL1D) {
    const double rv = g_x1v(geom, b);
    const double r2cyl = rv * rv;
    bx[0][0] = r2cyl - b->x1[0] * b->x1[0];
    bx[0][1] = b->x1[1] * b->x1[1] - r2cyl;
  }
}

/* Gravity::PointMassGravity<GEOM>, src/gravity/point_mass.cpp:26-196.
 * pm = {gm, x, y, z, soft, sink_rate, sink}: the gravity-package parameters it reads. */
void ao_point_mass_gravity(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                           double *gcons, const ao_fluid *dust, const double *dprim,
                           double *dcons, double dt, const double *pm) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const int geom = g->geom;
  const double gm = pm[0];
  const double pos[3] = {pm[1], pm[2], pm[3]};
  const double rsft2 = pm[4] * pm[4];
  const double sink_rate = dt * pm[5];
  const double sink_rad = pm[6];
  const int multi_d = (g->ndim >= 2), three_d = (g->ndim == 3);
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double dx[3] = {g_x1v(geom, &bb), g_x2v(geom, &bb), g_x3v(geom, &bb)};
          const double hx[3] = {g_hx1v(geom, &bb), g_hx2v(geom, &bb), g_hx3v(geom, &bb)};
          double gx1 = 0.0, gx2 = 0.0, gx3 = 0.0, dr;
          if (geom == AO_SPHERICAL1D || geom == AO_SPHERICAL2D) { /* :76-80 */
            const double rad2 = dx[0] * dx[0] + rsft2;
            gx1 = -gm / rad2;
            dr = sqrt(rad2);
          } else if (geom == AO_AXISYMMETRIC) { /* :81-88 with axisymmetric.hpp:115-133 */
            const double rs = sqrt(dx[0] * dx[0] + dx[1] * dx[1]);
            const double ct = dx[1] / (rs + AO_FUZZ);
            const double st = dx[0] / (rs + AO_FUZZ);
            dr = rs;
            const double rad2 = dr * dr + rsft2;
            const double gg = -gm / rad2;
            gx1 = gg * st; /* ex1[0] */
            gx2 = gg * ct; /* ex3[0] */
          } else { /* :89-114 */
            double dxc[3], e[3][3];
            g_to_cart(geom, dx, dxc, e);
            for (int n = 0; n < 3; n++) dxc[n] -= pos[n];
            /* Coords<cartesian>::ConvertToSph: only the radius is used (geometry.hpp:262-269) */
            const double R = sqrt(dxc[0] * dxc[0] + dxc[1] * dxc[1]);
            dr = sqrt(R * R + dxc[2] * dxc[2]);
            const double rad2 = dr * dr + rsft2;
            const double idr3 = 1.0 / (sqrt(rad2) * rad2);
            const double gv[3] = {-gm * dxc[0] * idr3, (multi_d) * (-gm * dxc[1] * idr3),
                                  (three_d) * (-gm * dxc[2] * idr3)};
            gx1 = gv[0] * e[0][0] + gv[1] * e[0][1] + gv[2] * e[0][2];
            gx2 = gv[0] * e[1][0] + gv[1] * e[1][1] + gv[2] * e[1][2];
            gx3 = gv[0] * e[2][0] + gv[1] * e[2][1] + gv[2] * e[2][2];
          }
          /* mass accretion :146-148 (quad_ramp(x) = x^2, gravity.hpp:116) */
          const double xr = (dr - sink_rad) / sink_rad;
          const double sramp = sink_rate * (xr * xr);
          double fd = dmin(0.5, sramp / (1.0 + sramp));
          fd *= ((sink_rate > 0.0) && (dr <= sink_rad));
          const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
          if (gas) {
            const int S = gas->nspecies;
            const double *w = gprim + (size_t)b * 6 * S * cells;
            double *u = gcons + (size_t)b * 6 * S * cells;
            for (int n = 0; n < S; ++n) {
              const double rho = w[(size_t)n * cells + o];
              const double v1 = w[(size_t)(S + 3 * n + 0) * cells + o];
              const double v2 = w[(size_t)(S + 3 * n + 1) * cells + o];
              const double v3 = w[(size_t)(S + 3 * n + 2) * cells + o];
              const double sie = w[(size_t)(5 * S + n) * cells + o];
              const double tote = rho * (sie + 0.5 * (v1 * v1 + v2 * v2 + v3 * v3));
              double *m1 = u + (size_t)(S + 3 * n + 0) * cells + o;
              double *m2 = u + (size_t)(S + 3 * n + 1) * cells + o;
              double *m3 = u + (size_t)(S + 3 * n + 2) * cells + o;
              double *en = u + (size_t)(4 * S + n) * cells + o;
              *m1 += dt * rho * hx[0] * gx1;
              *m2 += dt * rho * hx[1] * gx2;
              *m3 += dt * rho * hx[2] * gx3;
              *en += dt * rho * (v1 * gx1 + v2 * gx2 + v3 * gx3);
              u[(size_t)n * cells + o] -= fd * rho;
              *m1 -= fd * hx[0] * rho * v1;
              *m2 -= fd * hx[1] * rho * v2;
              *m3 -= fd * hx[2] * rho * v3;
              *en -= fd * tote;
            }
          }
          if (dust) {
            const int S = dust->nspecies;
            const double *w = dprim + (size_t)b * 4 * S * cells;
            double *u = dcons + (size_t)b * 4 * S * cells;
            for (int n = 0; n < S; ++n) {
              const double rho = w[(size_t)n * cells + o];
              const double v1 = w[(size_t)(S + 3 * n + 0) * cells + o];
              const double v2 = w[(size_t)(S + 3 * n + 1) * cells + o];
              const double v3 = w[(size_t)(S + 3 * n + 2) * cells + o];
              double *m1 = u + (size_t)(S + 3 * n + 0) * cells + o;
              double *m2 = u + (size_t)(S + 3 * n + 1) * cells + o;
              double *m3 = u + (size_t)(S + 3 * n + 2) * cells + o;
              *m1 += dt * rho * hx[0] * gx1;
              *m2 += dt * rho * hx[1] * gx2;
              *m3 += dt * rho * hx[2] * gx3;
              u[(size_t)n * cells + o] -= fd * rho;
              *m1 -= fd * hx[0] * rho * v1;
              *m2 -= fd * hx[1] * rho * v2;
              *m3 -= fd * hx[2] * rho * v3;
            }
          }
        }
}

/* RotatingFrame::RotatingFrameImpl<GEOM>, src/rotating_frame/rotating_frame_impl.hpp:96-199
 * (every non-Cartesian geometry, rotating_frame.cpp:69-82).  Reads the DENSITY fluxes of the
 * stage: gflux[d] / dflux[d] are the [nb][nvar][cells] flux slabs of ao_calculate_fluxes. */
static void rotating_frame_fluid(const ao_grid *g, int gasf, int S, int nvar, double *cons,
                                 const double *const fl[3], int b, int k, int j, int i,
                                 const double ax[3][2], const double bx[3][2], double vol,
                                 double omdt, double om2dt, double xcyl0, double e[3][3]) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const int multi_d = (g->ndim >= 2), three_d = (g->ndim == 3);
  const size_t sj = (size_t)g->ni, sk = (size_t)g->ni * g->nj;
  const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
  double *u = cons + (size_t)b * nvar * cells;
  for (int n = 0; n < S; ++n) {
    const double *f1 = fl[0] + ((size_t)b * nvar + n) * cells;
    const double *f2 = fl[1] ? fl[1] + ((size_t)b * nvar + n) * cells : f1;
    const double *f3 = fl[2] ? fl[2] + ((size_t)b * nvar + n) * cells : f1;
    const double f1m = f1[o], f1p = f1[o + 1];
    const double f2m = multi_d ? f2[o] : 0.0, f2p = multi_d ? f2[o + sj] : 0.0;
    const double f3m = three_d ? f3[o] : 0.0, f3p = three_d ? f3[o + sk] : 0.0;
    const double divf = (f1m * ax[0][0] * bx[0][0] + f1p * ax[0][1] * bx[0][1]) +
                        multi_d * (f2m * ax[1][0] * bx[1][0] + f2p * ax[1][1] * bx[1][1]) +
                        three_d * (f3m * ax[2][0] * bx[2][0] + f3p * ax[2][1] * bx[2][1]);
    u[(size_t)(S + 3 * n + 0) * cells + o] -= omdt * (divf / vol) * e[0][1];
    u[(size_t)(S + 3 * n + 1) * cells + o] -= omdt * (divf / vol) * e[1][1];
    u[(size_t)(S + 3 * n + 2) * cells + o] -= omdt * (divf / vol) * e[2][1];
    if (gasf) {
      const double fx[3] = {0.5 * (f1m + f1p), multi_d * 0.5 * (f2m + f2p),
                            three_d * 0.5 * (f3m + f3p)};
      u[(size_t)(4 * S + n) * cells + o] +=
          om2dt * xcyl0 * (fx[0] * e[0][0] + fx[1] * e[1][0] + fx[2] * e[2][0]);
    }
  }
}
void ao_rotating_frame(const ao_grid *g, const ao_fluid *gas, double *gcons,
                       const double *gflux1, const double *gflux2, const double *gflux3,
                       const ao_fluid *dust, double *dcons, const double *dflux1,
                       const double *dflux2, const double *dflux3, double dt, double om0) {
  const int geom = g->geom;
  if (geom == AO_CARTESIAN) return; /* the Cartesian frame is the shearing box */
  const int multi_d = (g->ndim >= 2), three_d = (g->ndim == 3);
  const double omdt = om0 * dt;
  const double om2dt = omdt * om0;
  const double *const gf[3] = {gflux1, gflux2, gflux3};
  const double *const df[3] = {dflux1, dflux2, dflux3};
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double xv[3] = {g_x1v(geom, &bb), g_x2v(geom, &bb), g_x3v(geom, &bb)};
          double xcyl[3], e[3][3], bx[3][2], ax[3][2];
          g_to_cyl(geom, xv, xcyl, e);
          g_rf_weights(geom, &bb, bx);
          ax[0][0] = g_area1(geom, &bb, bb.x1[0]);
          ax[0][1] = g_area1(geom, &bb, bb.x1[1]);
          ax[1][0] = multi_d ? g_area2(geom, &bb, bb.x2[0]) : 0.0;
          ax[1][1] = multi_d ? g_area2(geom, &bb, bb.x2[1]) : 0.0;
          ax[2][0] = three_d ? g_area3(geom, &bb, bb.x3[0]) : 0.0;
          ax[2][1] = three_d ? g_area3(geom, &bb, bb.x3[1]) : 0.0;
          const double vol = g_volume(geom, &bb);
          if (gas)
            rotating_frame_fluid(g, 1, gas->nspecies, 6 * gas->nspecies, gcons, gf, b, k, j, i,
                                 ax, bx, vol, omdt, om2dt, xcyl[0], e);
          if (dust)
            rotating_frame_fluid(g, 0, dust->nspecies, 4 * dust->nspecies, dcons, df, b, k, j, i,
                                 ax, bx, vol, omdt, om2dt, xcyl[0], e);
        }
}

/* ==========================================================================================
 * Diffusion operators: viscous stress and heat conduction of the gas.
 * Reference: src/utils/diffusion/{momentum_diffusion,thermal_diffusion,diffusion_coeff,
 * diffusion}.hpp driven by Gas::{ZeroDiffusionFlux,ViscousFlux,ThermalFlux,DiffusionUpdate}
 * (src/gas/gas.cpp:524-642).  The reference marches pencils with scratch rows; face by face the
 * arithmetic is the one restated here (same operands, same operation order).
 * ========================================================================================== */
typedef struct {
  double xv[3], hx[3];
  bbox_t bb;
} dcell_t;
static inline dcell_t dcell(const ao_grid *g, int b, int k, int j, int i) {
  dcell_t c;
  c.bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
  c.xv[0] = g_x1v(g->geom, &c.bb); c.xv[1] = g_x2v(g->geom, &c.bb); c.xv[2] = g_x3v(g->geom, &c.bb);
  c.hx[0] = g_hx1v(g->geom, &c.bb); c.hx[1] = g_hx2v(g->geom, &c.bb); c.hx[2] = g_hx3v(g->geom, &c.bb);
  return c;
}
/* CoordsBase::Distance, geometry.hpp:398-403 */
static inline double g_distance(int geom, const double x1[3], const double x2[3]) {
  double a[3], b[3], e[3][3];
  g_to_cart(geom, x1, a, e);
  g_to_cart(geom, x2, b, e);
  return sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) +
              (a[2] - b[2]) * (a[2] - b[2]));
}
/* radius of ConvertToSph(xv): geometry.hpp:262-269, cylindrical.hpp:110-115,
 * axisymmetric.hpp:115-120, spherical.hpp (identity) */
static inline double g_sph_radius(int geom, const double xi[3]) {
  if (geom == AO_CARTESIAN) {
    const double R = sqrt(xi[0] * xi[0] + xi[1] * xi[1]);
    return sqrt(R * R + xi[2] * xi[2]);
  }
  if (geom == AO_CYLINDRICAL) return sqrt(xi[0] * xi[0] + xi[2] * xi[2]);
  if (geom == AO_AXISYMMETRIC) return sqrt(xi[0] * xi[0] + xi[1] * xi[1]);
  return xi[0];
}
/* dh_D/dx_k: geometry.hpp:234-244 (all zero by default) + the overrides in g_conn1 / g_conn2 */
static inline void g_dh(int geom, const bbox_t *b, double dh[3][3]) {
  double c1[3], c2[3];
  g_conn1(geom, b, c1);
  g_conn2(geom, b, c2);
  for (int D = 0; D < 3; ++D) { dh[D][0] = c1[D]; dh[D][1] = c2[D]; dh[D][2] = 0.0; }
}
#define PR(v, kk, jj, ii) prim[IDX(g, nvar, b, (v), (kk), (jj), (ii))]
/* DiffusionCoeff<viscosity_plaw / viscosity_alpha>, diffusion_coeff.hpp:178-268 */
static double visc_mu_val(const ao_grid *g, const ao_fluid *f, const ao_diffusion *dd, int b, int k,
                          int j, int i, double dens, double sie) {
  const dcell_t c = dcell(g, b, k, j, i);
  if (dd->visc_type == AO_VISC_PLAW) {
    double xs[3], e[3][3];
    g_to_cyl(g->geom, c.xv, xs, e);
    return dd->nu * dens * pow(xs[0] / dd->r0, dd->r_exp);
  }
  const double rs = g_sph_radius(g->geom, c.xv);
  const double Omk = dd->omega0 * pow(rs / dd->r0, -1.5);
  const double blk = dmax(0.0, (f->gm1 + 1) * f->gm1 * dens * sie);
  return dd->alpha * blk / Omk;
}
static double visc_mu(const ao_grid *g, const ao_fluid *f, const ao_diffusion *dd, const double *prim,
                      int b, int n, int k, int j, int i) {
  const int S = f->nspecies, nvar = 6 * S;
  return visc_mu_val(g, f, dd, b, k, j, i, PR(n, k, j, i), PR(5 * S + n, k, j, i));
}
/* DiffusionCoeff<conductivity_plaw / thermaldiff_plaw>, diffusion_coeff.hpp:270-384 */
static double cond_kappa(const ao_grid *g, const ao_fluid *f, const ao_diffusion *dd,
                         const double *prim, int b, int n, int k, int j, int i) {
  const int S = f->nspecies, nvar = 6 * S;
  const double dens = PR(n, k, j, i), sie = PR(5 * S + n, k, j, i);
  const double T = dmax(0.0, sie / dd->cv);
  if (dd->cond_type == AO_COND_CONDUCTIVITY)
    return dd->cond * pow(T / dd->t_ref, dd->temp_exp) * pow(dens / dd->rho_ref, dd->rho_exp);
  return dd->kappa * pow(T / dd->t_ref, dd->temp_exp) * pow(dens / dd->rho_ref, dd->rho_exp) *
         dens * dd->cv;
}
/* FaceAverage selection of StressTensorFaceX* / ThermalFluxImpl: avg * arithmetic + havg * harmonic
 * (both evaluated), diffusion_coeff.hpp:148-160 */
static inline double face_avg(int avg_type, double m1, double m2) {
  const double avg = (avg_type == AO_AVG_ARITHMETIC), havg = (avg_type == AO_AVG_HARMONIC);
  return avg * (0.5 * (m1 + m2)) + havg * (2.0 * m1 * m2 / (m1 + m2));
}
/* VelocityDivergence, momentum_diffusion.hpp:553-590 */
static double vel_div(const ao_grid *g, const double *prim, int nvar, int S, int b, int n, int k,
                      int j, int i) {
  const int multid = (g->ndim >= 2), threed = (g->ndim == 3);
  const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
  const double vol = g_volume(g->geom, &bb);
  const double a1[2] = {g_area1(g->geom, &bb, bb.x1[0]), g_area1(g->geom, &bb, bb.x1[1])};
  const double a2[2] = {multid ? g_area2(g->geom, &bb, bb.x2[0]) : 0.0,
                        multid ? g_area2(g->geom, &bb, bb.x2[1]) : 0.0};
  const double a3[2] = {threed ? g_area3(g->geom, &bb, bb.x3[0]) : 0.0,
                        threed ? g_area3(g->geom, &bb, bb.x3[1]) : 0.0};
  const int v1 = S + 3 * n, v2 = S + 3 * n + 1, v3 = S + 3 * n + 2;
  const double divv =
      a1[1] * (PR(v1, k, j, i) + PR(v1, k, j, i + 1)) -
      a1[0] * (PR(v1, k, j, i) + PR(v1, k, j, i - 1)) +
      multid * a2[1] * (PR(v2, k, j, i) + PR(v2, k, j + multid, i)) -
      multid * a2[0] * (PR(v2, k, j, i) + PR(v2, k, j - multid, i)) +
      threed * a3[1] * (PR(v3, k, j, i) + PR(v3, k + threed, j, i)) -
      threed * a3[0] * (PR(v3, k, j, i) + PR(v3, k - threed, j, i));
  return divv / (2.0 * vol);
}
/* Viscous flux through the LOWER face of cell (k,j,i) in direction D (0,1,2):
 * StrainTensorFace<XDIR> + StressTensorFaceX{1,2,3}, momentum_diffusion.hpp:27-551.
 * out = {F(m1), F(m2), F(m3), F(E)}. */
static void visc_face(const ao_grid *g, const ao_fluid *f, const ao_diffusion *dd,
                      const double *prim, int D, int b, int n, int k, int j, int i, double out[4]) {
  const int S = f->nspecies, nvar = 6 * S, geom = g->geom;
  const int multid = (g->ndim >= 2), threed = (g->ndim == 3);
  const int off[3] = {1, multid, threed};   /* neighbour offset along each direction */
  const int vi[3] = {S + 3 * n, S + 3 * n + 1, S + 3 * n + 2};
  int dk[3] = {0, 0, 0}, dj[3] = {0, 0, 0}, di[3] = {0, 0, 0};
  di[0] = 1; dj[1] = 1; dk[2] = 1;          /* unit step of direction c */
  const dcell_t c0 = dcell(g, b, k, j, i);
  const int km = k - dk[D], jm = j - dj[D], im = i - di[D];  /* the cell across the face */
  const dcell_t cm = dcell(g, b, km, jm, im);
  double v[3], vm[3];
  for (int c = 0; c < 3; ++c) {
    v[c] = PR(vi[c], k, j, i) / c0.hx[c];
    vm[c] = PR(vi[c], km, jm, im) / cm.hx[c];
  }
  double xf[3];
  if (D == 0) g_facecen1(geom, &c0.bb, 0, xf);
  else if (D == 1) g_facecen2(geom, &c0.bb, 0, xf);
  else g_facecen3(geom, &c0.bb, 0, xf);
  const double hxf[3] = {g_hx1(geom, xf[0], xf[1], xf[2]), g_hx2(geom, xf[0], xf[1], xf[2]),
                         g_hx3(geom, xf[0], xf[1], xf[2])};
  const double dxD = g_distance(geom, c0.xv, cm.xv);
  double dh0[3][3], dhm[3][3];
  g_dh(geom, &c0.bb, dh0);
  g_dh(geom, &cm.bb, dhm);
  double flx[3];
  for (int c = 0; c < 3; ++c) {
    if (c == D) {
      /* T_D^D = 2 dv^D/dxD + v^k dhD/dxk / hD  (averaged over the two cells) */
      const double dv = v[D] - vm[D];
      const double src = v[0] * dh0[D][0] + v[1] * dh0[D][1] + v[2] * dh0[D][2];
      const double srm = vm[0] * dhm[D][0] + vm[1] * dhm[D][1] + vm[2] * dhm[D][2];
      flx[c] = 2 * dv / dxD + 0.5 * (src + srm);
    } else {
      /* T_c^D = dv^D/dxc (transverse, averaged over the two cells) + hc^2/hD^2 dv^c/dxD */
      const int o = off[c];
      const int kp = k + o * dk[c], jp = j + o * dj[c], ip = i + o * di[c];
      const int kq = k - o * dk[c], jq = j - o * dj[c], iq = i - o * di[c];
      const int kmp = km + o * dk[c], jmp = jm + o * dj[c], imp = im + o * di[c];
      const int kmq = km - o * dk[c], jmq = jm - o * dj[c], imq = im - o * di[c];
      const dcell_t cp = dcell(g, b, kp, jp, ip), cq = dcell(g, b, kq, jq, iq);
      const dcell_t cmp = dcell(g, b, kmp, jmp, imp), cmq = dcell(g, b, kmq, jmq, imq);
      const double dxc = o ? g_distance(geom, cq.xv, cp.xv) : AO_FUZZ;
      const double dxc_m = o ? g_distance(geom, cmq.xv, cmp.xv) : AO_FUZZ;
      const double dvt = PR(vi[D], kp, jp, ip) / cp.hx[D] - PR(vi[D], kq, jq, iq) / cq.hx[D];
      const double dvt_m = PR(vi[D], kmp, jmp, imp) / cmp.hx[D] - PR(vi[D], kmq, jmq, imq) / cmq.hx[D];
      const double dv = v[c] - vm[c];
      const double r = hxf[c] / hxf[D];
      flx[c] = o * 0.5 * (dvt / dxc + dvt_m / dxc_m) + (r * r) * dv / dxD;
    }
  }
  const double mus = face_avg(dd->visc_avg, visc_mu(g, f, dd, prim, b, n, k, j, i),
                              visc_mu(g, f, dd, prim, b, n, km, jm, im));
  const double divs = (D == 0) ? vel_div(g, prim, nvar, S, b, n, k, j, i) +
                                     vel_div(g, prim, nvar, S, b, n, km, jm, im)
                               : vel_div(g, prim, nvar, S, b, n, km, jm, im) +
                                     vel_div(g, prim, nvar, S, b, n, k, j, i);
  const double hDf = hxf[D];
  double fc[3];
  for (int c = 0; c < 3; ++c)
    fc[c] = (c == D) ? hDf * mus * (flx[c] - 1. / 3 * (1. - dd->eta) * divs) : hDf * mus * flx[c];
  out[0] = fc[0]; out[1] = fc[1]; out[2] = fc[2];
  out[3] = 0.5 * (PR(vi[0], k, j, i) / c0.hx[0] + PR(vi[0], km, jm, im) / cm.hx[0]) * fc[0] +
           0.5 * (PR(vi[1], k, j, i) / c0.hx[1] + PR(vi[1], km, jm, im) / cm.hx[1]) * fc[1] +
           0.5 * (PR(vi[2], k, j, i) / c0.hx[2] + PR(vi[2], km, jm, im) / cm.hx[2]) * fc[2];
}
/* Heat flux through the lower face of cell (k,j,i) in direction D, thermal_diffusion.hpp:62-218 */
static double cond_face(const ao_grid *g, const ao_fluid *f, const ao_diffusion *dd,
                        const double *prim, int D, int b, int n, int k, int j, int i) {
  const int S = f->nspecies, nvar = 6 * S;
  const int km = k - (D == 2), jm = j - (D == 1), im = i - (D == 0);
  const dcell_t c0 = dcell(g, b, k, j, i), cm = dcell(g, b, km, jm, im);
  const double dxD = g_distance(g->geom, c0.xv, cm.xv);
  const double T = dmax(0.0, PR(5 * S + n, k, j, i) / dd->cv);
  const double Tm = dmax(0.0, PR(5 * S + n, km, jm, im) / dd->cv);
  const double kc = face_avg(dd->cond_avg, cond_kappa(g, f, dd, prim, b, n, k, j, i),
                             cond_kappa(g, f, dd, prim, b, n, km, jm, im));
  return kc * (T - Tm) / dxD;
}
#undef PR

/* Gas::ZeroDiffusionFlux + Gas::ViscousFlux + Gas::ThermalFlux (src/gas/gas.cpp:524-603): every
 * face of every interior cell: x1 faces is..ie+1, x2 faces js..je+1, x3 faces ks..ke+1 */
void ao_diffusion_flux(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                       const ao_diffusion *dd, double *dflx1, double *dflx2, double *dflx3) {
  const int S = gas->nspecies, nv = 4 * S;
  const int multid = (g->ndim >= 2), threed = (g->ndim == 3);
  double *dflx[3] = {dflx1, dflx2, dflx3};
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke + threed; ++k)
      for (int j = g->js; j <= g->je + multid; ++j)
        for (int i = g->is; i <= g->ie + 1; ++i)
          for (int D = 0; D < g->ndim; ++D) {
            /* the face must bound an interior cell in the two transverse directions */
            if (D != 0 && i > g->ie) continue;
            if (D != 1 && j > g->je) continue;
            if (D != 2 && k > g->ke) continue;
            for (int n = 0; n < S; ++n) {
              double o[4] = {0.0, 0.0, 0.0, 0.0}, acc[4] = {0.0, 0.0, 0.0, 0.0};
              if (dd->visc_type != AO_DIFF_NONE) {
                visc_face(g, gas, dd, gprim, D, b, n, k, j, i, o);
                for (int m = 0; m < 4; ++m) acc[m] += o[m];
              }
              if (dd->cond_type != AO_COND_NONE)
                acc[3] += cond_face(g, gas, dd, gprim, D, b, n, k, j, i);
              dflx[D][FIDX(g, nv, b, 3 * n + 0, k, j, i)] = acc[0];
              dflx[D][FIDX(g, nv, b, 3 * n + 1, k, j, i)] = acc[1];
              dflx[D][FIDX(g, nv, b, 3 * n + 2, k, j, i)] = acc[2];
              dflx[D][FIDX(g, nv, b, 3 * S + n, k, j, i)] = acc[3];
            }
          }
}

/* Diffusion::DiffusionUpdateImpl, diffusion.hpp:113-242 */
void ao_diffusion_update(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                         const ao_diffusion *dd, const double *dflx1, const double *dflx2,
                         const double *dflx3, double dt) {
  const int S = gas->nspecies, nv = 4 * S, nvar = 6 * S, geom = g->geom;
  const int multi_d = (g->ndim > 1), three_d = (g->ndim > 2);
  const int do_viscosity = dd->visc_type != AO_DIFF_NONE;
  const int x1dep = g_x1dep(geom), x2dep = g_x2dep(geom) && multi_d, x3dep = 0;
  const double *F1 = dflx1, *F2 = multi_d ? dflx2 : dflx1, *F3 = three_d ? dflx3 : dflx1;
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double ax1[2] = {g_area1(geom, &bb, bb.x1[0]), g_area1(geom, &bb, bb.x1[1])};
          const double ax2[2] = {multi_d ? g_area2(geom, &bb, bb.x2[0]) : 0.0,
                                 multi_d ? g_area2(geom, &bb, bb.x2[1]) : 0.0};
          const double ax3[2] = {three_d ? g_area3(geom, &bb, bb.x3[0]) : 0.0,
                                 three_d ? g_area3(geom, &bb, bb.x3[1]) : 0.0};
          double dhdx1[3] = {0, 0, 0}, dhdx2[3] = {0, 0, 0}, dhdx3[3] = {0, 0, 0};
          if (x1dep) g_conn1(geom, &bb, dhdx1);
          if (x2dep) g_conn2(geom, &bb, dhdx2);
          const double hx[3] = {g_hx1v(geom, &bb), g_hx2v(geom, &bb), g_hx3v(geom, &bb)};
          const double vol = g_volume(geom, &bb);
#define F(A, v, kk, jj, ii) A[FIDX(g, nv, b, (v), (kk), (jj), (ii))]
          for (int n = 0; n < S; ++n) {
            const int m1 = 3 * n, m2 = 3 * n + 1, m3 = 3 * n + 2, ien = 3 * S + n;
            double divfxm = 0., divfym = 0., divfzm = 0.;
            if (do_viscosity) {
              const double s1 = 0.5 * (F(F1, m1, k, j, i) + F(F1, m1, k, j, i + 1));
              const double s2 = 0.5 * (F(F2, m2, k, j, i) + F(F2, m2, k, j + multi_d, i));
              const double s3 = 0.5 * (F(F3, m3, k, j, i) + F(F3, m3, k + three_d, j, i));
              divfxm = (ax1[0] * F(F1, m1, k, j, i) - ax1[1] * F(F1, m1, k, j, i + 1)) +
                       multi_d * (ax2[0] * F(F2, m1, k, j, i) - ax2[1] * F(F2, m1, k, j + multi_d, i)) +
                       three_d * (ax3[0] * F(F3, m1, k, j, i) - ax3[1] * F(F3, m1, k + three_d, j, i));
              divfxm /= vol;
              double src = dhdx1[0] * s1 + multi_d * dhdx1[1] * s2 + three_d * dhdx1[2] * s3;
              divfxm += x1dep * src;
              divfym = (ax1[0] * F(F1, m2, k, j, i) - ax1[1] * F(F1, m2, k, j, i + 1)) +
                       multi_d * (ax2[0] * F(F2, m2, k, j, i) - ax2[1] * F(F2, m2, k, j + multi_d, i)) +
                       three_d * (ax3[0] * F(F3, m2, k, j, i) - ax3[1] * F(F3, m2, k + three_d, j, i));
              divfym /= vol;
              src = dhdx2[0] * s1 + multi_d * dhdx2[1] * s2 + three_d * dhdx2[2] * s3;
              divfym += x2dep * src;
              divfzm = (ax1[0] * F(F1, m3, k, j, i) - ax1[1] * F(F1, m3, k, j, i + 1)) +
                       multi_d * (ax2[0] * F(F2, m3, k, j, i) - ax2[1] * F(F2, m3, k, j + multi_d, i)) +
                       three_d * (ax3[0] * F(F3, m3, k, j, i) - ax3[1] * F(F3, m3, k + three_d, j, i));
              divfzm /= vol;
              src = dhdx3[0] * s1 + multi_d * dhdx3[1] * s2 + three_d * dhdx3[2] * s3;
              divfzm += x3dep * src;
            }
            double divfe = (ax1[0] * F(F1, ien, k, j, i) - ax1[1] * F(F1, ien, k, j, i + 1)) +
                           multi_d * (ax2[0] * F(F2, ien, k, j, i) - ax2[1] * F(F2, ien, k, j + multi_d, i)) +
                           three_d * (ax3[0] * F(F3, ien, k, j, i) - ax3[1] * F(F3, ien, k + three_d, j, i));
            divfe /= vol;
            gcons[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)] -= dt * divfxm;
            gcons[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)] -= dt * divfym;
            gcons[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)] -= dt * divfzm;
            gcons[IDX(g, nvar, b, 4 * S + n, k, j, i)] -= dt * divfe;
            gcons[IDX(g, nvar, b, 5 * S + n, k, j, i)] -=
                dt * divfe - dt * (divfxm * gprim[IDX(g, nvar, b, S + 3 * n + 0, k, j, i)] / hx[0] +
                                   divfym * gprim[IDX(g, nvar, b, S + 3 * n + 1, k, j, i)] / hx[1] +
                                   divfzm * gprim[IDX(g, nvar, b, S + 3 * n + 2, k, j, i)] / hx[2]);
          }
#undef F
        }
}

/* Diffusion::EstimateTimestep for the configured viscosity / conduction, min of the two
 * (diffusion.hpp:64-111, src/gas/gas.cpp:437-464) */
double ao_diffusion_dt(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                       const ao_diffusion *dd) {
  const int S = gas->nspecies, nvar = 6 * S;
  const double big = 1.79769313486231570815e+308;
  double dtv = big, dtc = big;
#pragma omp parallel for collapse(2) schedule(static) reduction(min : dtv) reduction(min : dtc)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          double dx[3];
          g_cell_widths(g->geom, &bb, dx);
          double min_dx = big;
          for (int d = 0; d < g->ndim; d++) min_dx = dmin(min_dx, dx[d]);
          for (int n = 0; n < S; ++n) {
            const double dens = gprim[IDX(g, nvar, b, n, k, j, i)];
            if (dd->visc_type != AO_DIFF_NONE) {
              double mu = visc_mu(g, gas, dd, gprim, b, n, k, j, i);
              mu *= (1.0 + (dd->eta > 1.0) * (dd->eta - 1.0)) / dens;
              dtv = dmin(dtv, min_dx * min_dx / (mu + AO_FUZZ));
            }
            if (dd->cond_type != AO_COND_NONE) {
              double mu = cond_kappa(g, gas, dd, gprim, b, n, k, j, i);
              if (dd->cond_type == AO_COND_CONDUCTIVITY) mu /= (dens * dd->cv);
              dtc = dmin(dtc, min_dx * min_dx / (mu + AO_FUZZ));
            }
          }
        }
  if (dd->visc_type != AO_DIFF_NONE) dtv = dtv / (2.0 * g->ndim);
  if (dd->cond_type != AO_COND_NONE) dtc = dtc / (2.0 * g->ndim);
  return dmin(dtv, dtc);
}

/* quadratic damping ramp of one direction (drag.hpp:195-199): dt * (irate * [x < ix] *
 * ((x - ix) / (ix - xmin))^2 + orate * [x > ox] * ((x - ox) / (ox - xmax))^2) */
static inline double damp_ramp(double dt, double x, double ix, double ox, double irate,
                               double orate, double xmin, double xmax) {
  const double ri = (x - ix) / (ix - xmin), ro = (x - ox) / (ox - xmax);
  return dt * (irate * ((x < ix) * (ri * ri)) + orate * ((x > ox) * (ro * ro)));
}
/* ArtemisUtils::GetSpecificInternalEnergy, src/utils/artemis_utils.hpp:42-62 */
static inline double cons_sie(const double *ug, size_t cells, size_t o, int S, int n,
                              const double hx[3], double de_switch, double dflr, double sieflr) {
  const double u_d = dmax(ug[(size_t)n * cells + o], dflr);
  const double rv1 = ug[(size_t)(S + 3 * n + 0) * cells + o] / hx[0];
  const double rv2 = ug[(size_t)(S + 3 * n + 1) * cells + o] / hx[1];
  const double rv3 = ug[(size_t)(S + 3 * n + 2) * cells + o] / hx[2];
  const double ke = 0.5 * (rv1 * rv1 + rv2 * rv2 + rv3 * rv3) / u_d;
  const double e_cons = ug[(size_t)(4 * S + n) * cells + o];
  const double ue_cons = e_cons - ke;
  const double sie = (ue_cons > de_switch * e_cons) ? ue_cons / u_d
                                                    : ug[(size_t)(5 * S + n) * cells + o] / u_d;
  return dmax(sie, sieflr);
}

void ao_drag_source(const ao_grid *g, const ao_fluid *gas, double *gcons, const ao_fluid *dust,
                    double *dcons, const ao_drag *dp, const ao_diffusion *dd, double dt) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const int geom = g->geom;
  const int Sg = gas ? gas->nspecies : 0, Sd = dust ? dust->nspecies : 0;
  const double big = 1.79769313486231570815e+308; /* Big<Real>() */
  const int multi_d = (g->ndim >= 2), three_d = (g->ndim == 3);
  const int use_visc = dp->g_damp_to_visc && dd && dd->visc_type != AO_DIFF_NONE;
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double xv[3] = {g_x1v(geom, &bb), g_x2v(geom, &bb), g_x3v(geom, &bb)};
          const double hx[3] = {g_hx1v(geom, &bb), g_hx2v(geom, &bb), g_hx3v(geom, &bb)};
          double xcyl[3], e[3][3];
          g_to_cyl(geom, xv, xcyl, e);
          const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
          double *ug = gas ? gcons + (size_t)b * 6 * Sg * cells : NULL;
          double *ud = dust ? dcons + (size_t)b * 4 * Sd * cells : NULL;
          const double dsc[3] = {1.0, multi_d, three_d};
          double bg[3], bd[3];
          for (int d = 0; d < 3; ++d) {
            const double rg = damp_ramp(dt, xv[d], dp->g_ix[d], dp->g_ox[d], dp->g_irate[d],
                                        dp->g_orate[d], dp->xmin[d], dp->xmax[d]);
            const double rd = damp_ramp(dt, xv[d], dp->d_ix[d], dp->d_ox[d], dp->d_irate[d],
                                        dp->d_orate[d], dp->xmin[d], dp->xmax[d]);
            /* fx1 = dt * (..); fx2 = multi_d * dt * (..); fx3 = three_d * dt * (..) */
            bg[d] = (d == 0) ? rg : dsc[d] * rg;
            bd[d] = (d == 0) ? rd : dsc[d] * rd;
          }
          if (dp->coupling == AO_DRAG_SELF) { /* SelfDragSourceImpl, drag.hpp:150-291 */
            for (int n = 0; n < Sg; ++n) {
              const double dens = ug[(size_t)n * cells + o];
              const double vg[3] = {ug[(size_t)(Sg + 3 * n + 0) * cells + o] / (hx[0] * dens),
                                    ug[(size_t)(Sg + 3 * n + 1) * cells + o] / (hx[1] * dens),
                                    ug[(size_t)(Sg + 3 * n + 2) * cells + o] / (hx[2] * dens)};
              const double sieg = cons_sie(ug, cells, o, Sg, n, hx, gas->de_switch, gas->dfloor,
                                           gas->siefloor);
              const double mu = use_visc ? visc_mu_val(g, gas, dd, b, k, j, i, dens, sieg) : 0.0;
              const double vR = -1.5 * mu / (xcyl[0] * dens);
              const double vd[3] = {e[0][0] * vR, e[1][0] * vR, e[2][0] * vR};
              const double dm1 = -bg[0] * dens * (vg[0] - vd[0]) / (1.0 + bg[0]);
              const double dm2 = -bg[1] * dens * (vg[1] - vd[1]) / (1.0 + bg[1]);
              const double dm3 = -bg[2] * dens * (vg[2] - vd[2]) / (1.0 + bg[2]);
              ug[(size_t)(Sg + 3 * n + 0) * cells + o] += hx[0] * dm1;
              ug[(size_t)(Sg + 3 * n + 1) * cells + o] += hx[1] * dm2;
              ug[(size_t)(Sg + 3 * n + 2) * cells + o] += hx[2] * dm3;
              ug[(size_t)(4 * Sg + n) * cells + o] += dm1 * (vg[0] + 0.5 * dm1 / dens) +
                                                      dm2 * (vg[1] + 0.5 * dm2 / dens) +
                                                      dm3 * (vg[2] + 0.5 * dm3 / dens);
            }
            for (int n = 0; n < Sd; ++n)
              for (int d = 0; d < 3; ++d) {
                double *m = ud + (size_t)(Sd + 3 * n + d) * cells + o;
                const double mom = *m;
                *m -= bd[d] * mom / (1.0 + bd[d]);
              }
            continue;
          }
          /* SimpleDragSourceImpl, drag.hpp:296-482 */
          const double dg = ug[o];
          const double vg[3] = {ug[(size_t)(Sg + 0) * cells + o] / (hx[0] * dg),
                                ug[(size_t)(Sg + 1) * cells + o] / (hx[1] * dg),
                                ug[(size_t)(Sg + 2) * cells + o] / (hx[2] * dg)};
          const double sieg = cons_sie(ug, cells, o, Sg, 0, hx, gas->de_switch, gas->dfloor,
                                       gas->siefloor);
          const double mu = use_visc ? visc_mu_val(g, gas, dd, b, k, j, i, dg, sieg) : 0.0;
          const double vR = -1.5 * mu / (xcyl[0] * dg);
          const double vt[3] = {e[0][0] * vR, e[1][0] * vR, e[2][0] * vR};
          const double vdt[3] = {0.0, 0.0, 0.0};
          double vth = 0.0;
          if (dp->model == AO_DRAG_STOKES) vth = sqrt(8.0 / M_PI * gas->gm1 * sieg);
          double fd[3] = {0.0, 0.0, 0.0}, fvd[3] = {0.0, 0.0, 0.0};
          for (int n = 0; n < Sd; ++n) {
            const double dens = ud[(size_t)n * cells + o];
            const double vd[3] = {ud[(size_t)(Sd + 3 * n + 0) * cells + o] / (hx[0] * dens),
                                  ud[(size_t)(Sd + 3 * n + 1) * cells + o] / (hx[1] * dens),
                                  ud[(size_t)(Sd + 3 * n + 2) * cells + o] / (hx[2] * dens)};
            double tc = (dp->model == AO_DRAG_STOKES) ? dp->scale : dp->scale * dp->tau[n];
            if (dp->model == AO_DRAG_STOKES)
              tc = dp->scale * dp->grain_density / dg * dp->sizes[n] / vth;
            const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
            for (int d = 0; d < 3; d++) {
              const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
              fd[d] += rhop * (1.0 + bd[d]);
              fvd[d] += rhop * (vd[d] + bd[d] * vdt[d]);
            }
          }
          double vgp[3];
          for (int d = 0; d < 3; d++)
            vgp[d] = (dg * (vg[d] + bg[d] * vt[d]) + fvd[d]) / (dg * (1.0 + bg[d]) + fd[d]);
          double delta_g[3] = {0.0, 0.0, 0.0};
          for (int d = 0; d < 3; d++) fvd[d] = 0.;
          for (int n = 0; n < Sd; ++n) {
            const double dens = ud[(size_t)n * cells + o];
            const double vd[3] = {ud[(size_t)(Sd + 3 * n + 0) * cells + o] / (hx[0] * dens),
                                  ud[(size_t)(Sd + 3 * n + 1) * cells + o] / (hx[1] * dens),
                                  ud[(size_t)(Sd + 3 * n + 2) * cells + o] / (hx[2] * dens)};
            double tc = (dp->model == AO_DRAG_STOKES) ? dp->scale : dp->scale * dp->tau[n];
            if (dp->model == AO_DRAG_STOKES)
              tc = dp->scale * dp->grain_density / dg * dp->sizes[n] / vth;
            const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
            for (int d = 0; d < 3; d++) {
              double delta_d = 0.;
              const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
              const double delta = rhop * ((vgp[d] - vd[d] + bd[d] * (vgp[d] - vdt[d])));
              delta_d += delta;
              delta_g[d] -= delta;
              delta_d -= bd[d] * dens / (1. + alpha + bd[d]) *
                         (vd[d] - vdt[d] + alpha * (vgp[d] - vdt[d]));
              fvd[d] += rhop * (vd[d] - vt[d] + bd[d] * (vdt[d] - vt[d]));
              ud[(size_t)(Sd + 3 * n + d) * cells + o] += hx[d] * delta_d;
            }
          }
          for (int d = 0; d < 3; d++) {
            const double prefac = dg * bg[d] / (1.0 + bg[d] + fd[d]);
            delta_g[d] -= prefac * (dg * (vg[d] - vt[d]) + fvd[d]);
            ug[(size_t)(Sg + d) * cells + o] += hx[d] * delta_g[d];
            ug[(size_t)(4 * Sg) * cells + o] += 0.5 * (vg[d] + vgp[d]) * delta_g[d];
          }
        }
}

/* RotatingFrame::ShearingBoxImpl, src/rotating_frame/rotating_frame_impl.hpp:28-94
 * (Cartesian only: Coriolis + tidal potential differenced across the cell) */
void ao_shearing_box(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                     const ao_fluid *dust, const double *dprim, double *dcons, double dt,
                     double om0, double qshear) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const int three_d = (g->ndim == 3);
  const double omsq = om0 * om0;
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double dx = bb.x1[1] - bb.x1[0];
          const double dz = bb.x3[1] - bb.x3[0];
          const double phi_xm1 = -qshear * omsq * bb.x1[0] * bb.x1[0];
          const double phi_xp1 = -qshear * omsq * bb.x1[1] * bb.x1[1];
          const double phi_zm1 = 0.5 * omsq * bb.x3[0] * bb.x3[0];
          const double phi_zp1 = 0.5 * omsq * bb.x3[1] * bb.x3[1];
          const double dpx = (phi_xp1 - phi_xm1) / dx;
          const double dpz = three_d * ((phi_zp1 - phi_zm1) / dz);
          const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
          for (int fl = 0; fl < 2; ++fl) {
            const ao_fluid *f = fl ? dust : gas;
            if (!f) continue;
            const int S = f->nspecies, nv = (fl ? 4 : 6) * S;
            const double *w = (fl ? dprim : gprim) + (size_t)b * nv * cells;
            double *u = (fl ? dcons : gcons) + (size_t)b * nv * cells;
            for (int n = 0; n < S; ++n) {
              const double dens = w[(size_t)n * cells + o];
              const double v1 = w[(size_t)(S + 3 * n + 0) * cells + o];
              const double v2 = w[(size_t)(S + 3 * n + 1) * cells + o];
              const double v3 = w[(size_t)(S + 3 * n + 2) * cells + o];
              const double rdt = dens * dt;
              u[(size_t)(S + 3 * n + 0) * cells + o] -= rdt * (dpx - 2.0 * om0 * v2);
              u[(size_t)(S + 3 * n + 1) * cells + o] -= rdt * 2.0 * om0 * v1;
              u[(size_t)(S + 3 * n + 2) * cells + o] -= rdt * dpz;
              if (!fl) u[(size_t)(4 * S + n) * cells + o] -= rdt * (v1 * dpx + v3 * dpz);
            }
          }
        }
}

/* Drag::SimpleDragSourceImpl<DiffType::null, DragModel::constant, GEOM>,
 * src/drag/drag.hpp:296-482, with no damping zones (irate = orate = 0 => bg = bd = 0) and no
 * viscous target velocity (mu = 0 => vt = 0): the implicit gas <-> dust momentum exchange with
 * constant stopping times tau[n] (tp.tau, drag.hpp:112-129).  gas species 0 couples to every
 * dust species; total momentum is conserved to rounding (tst/scripts/drag/drag.py:135-137). */
void ao_drag_simple(const ao_grid *g, const ao_fluid *gas, double *gcons, const ao_fluid *dust,
                    double *dcons, double dt, const double *tau) {
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const int Sg = gas->nspecies, Sd = dust->nspecies;
  const double big = 1.79769313486231570815e+308; /* Big<Real>() */
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < g->nb; ++b)
    for (int k = g->ks; k <= g->ke; ++k)
      for (int j = g->js; j <= g->je; ++j)
        for (int i = g->is; i <= g->ie; ++i) {
          const bbox_t bb = make_bbox(g->xmin + 3 * b, g->dx + 3 * b, k, j, i);
          const double hx[3] = {g_hx1v(g->geom, &bb), g_hx2v(g->geom, &bb), g_hx3v(g->geom, &bb)};
          const size_t o = ((size_t)k * g->nj + j) * g->ni + i;
          double *ug = gcons + (size_t)b * 6 * Sg * cells;
          double *ud = dcons + (size_t)b * 4 * Sd * cells;
          const double bg[3] = {0.0, 0.0, 0.0}, bd[3] = {0.0, 0.0, 0.0};
          const double vt[3] = {0.0, 0.0, 0.0}, vdt[3] = {0.0, 0.0, 0.0};
          const double dg = ug[o];
          const double vg[3] = {ug[(size_t)(Sg + 0) * cells + o] / (hx[0] * dg),
                                ug[(size_t)(Sg + 1) * cells + o] / (hx[1] * dg),
                                ug[(size_t)(Sg + 2) * cells + o] / (hx[2] * dg)};
          double fd[3] = {0.0, 0.0, 0.0}, fvd[3] = {0.0, 0.0, 0.0};
          for (int n = 0; n < Sd; ++n) {
            const double dens = ud[(size_t)n * cells + o];
            const double vd[3] = {ud[(size_t)(Sd + 3 * n + 0) * cells + o] / (hx[0] * dens),
                                  ud[(size_t)(Sd + 3 * n + 1) * cells + o] / (hx[1] * dens),
                                  ud[(size_t)(Sd + 3 * n + 2) * cells + o] / (hx[2] * dens)};
            const double tc = tau[n];
            const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
            for (int d = 0; d < 3; d++) {
              const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
              fd[d] += rhop * (1.0 + bd[d]);
              fvd[d] += rhop * (vd[d] + bd[d] * vdt[d]);
            }
          }
          double vgp[3];
          for (int d = 0; d < 3; d++)
            vgp[d] = (dg * (vg[d] + bg[d] * vt[d]) + fvd[d]) / (dg * (1.0 + bg[d]) + fd[d]);
          double delta_g[3] = {0.0, 0.0, 0.0};
          for (int d = 0; d < 3; d++) fvd[d] = 0.;
          for (int n = 0; n < Sd; ++n) {
            const double dens = ud[(size_t)n * cells + o];
            const double vd[3] = {ud[(size_t)(Sd + 3 * n + 0) * cells + o] / (hx[0] * dens),
                                  ud[(size_t)(Sd + 3 * n + 1) * cells + o] / (hx[1] * dens),
                                  ud[(size_t)(Sd + 3 * n + 2) * cells + o] / (hx[2] * dens)};
            const double tc = tau[n];
            const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
            for (int d = 0; d < 3; d++) {
              double delta_d = 0.;
              const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
              const double delta = rhop * ((vgp[d] - vd[d] + bd[d] * (vgp[d] - vdt[d])));
              delta_d += delta;
              delta_g[d] -= delta;
              delta_d -= bd[d] * dens / (1. + alpha + bd[d]) *
                         (vd[d] - vdt[d] + alpha * (vgp[d] - vdt[d]));
              fvd[d] += rhop * (vd[d] - vt[d] + bd[d] * (vdt[d] - vt[d]));
              ud[(size_t)(Sd + 3 * n + d) * cells + o] += hx[d] * delta_d;
            }
          }
          for (int d = 0; d < 3; d++) {
            const double prefac = dg * bg[d] / (1.0 + bg[d] + fd[d]);
            delta_g[d] -= prefac * (dg * (vg[d] - vt[d]) + fvd[d]);
            ug[(size_t)(Sg + d) * cells + o] += hx[d] * delta_g[d];
            ug[(size_t)(4 * Sg) * cells + o] += 0.5 * (vg[d] + vgp[d]) * delta_g[d];
          }
        }
}
