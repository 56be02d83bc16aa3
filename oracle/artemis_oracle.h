/*
 * artemis_oracle.h -- CPU restatement of the Artemis finite-volume hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / reported CPU baseline.
 *
 * Every function cites the reference file:line (relative to lanl/artemis @ 6c2a7a8)
 * whose arithmetic it restates, operation for operation, so that the result is
 * bit-identical to the reference's Kokkos-OpenMP build when compiled without FMA
 * contraction (gcc -O2 -ffp-contract=off; the reference build uses no -march).
 *
 * Pinning: see oracle/README.md -- (1) linear-wave RMS-L1 golden numbers measured from
 * the real reference, (2) the reference's own regression thresholds
 * (tst/scripts/hydro/linwave.py), (3) oracle/_ref: the reference's own headers
 * compiled against a mock Parthenon, compared function by function.
 *
 * Data model ("MeshBlockPack layout"): every variable is a dense
 * array [nb][nvar][nk][nj][ni], i fastest, ghosts included.  Pack index conventions
 * follow src/utils/fluxes/riemann/hllc.hpp:66-73 (S = nspecies):
 *   gas  prim: rho n | vel S+3n+d | pressure 4S+n | sie 5S+n          (6S vars)
 *   gas  cons: rho n | mom S+3n+d | total_energy 4S+n | internal 5S+n (6S vars)
 *   dust prim: rho n | vel S+3n+d                                     (4S vars)
 *   dust cons: rho n | mom S+3n+d                                     (4S vars)
 * Fluxes: flux[d] has the cons layout; flux at index i is the LOWER face of cell i.
 * pflux[d]: [nb][S][nk][nj][ni] interface pressure (the prim-pressure flux slot).
 * vface[d]: [nb][S][fnk][fnj][fni] face velocity (gas.face.velocity, el = d).
 */
#ifndef ARTEMIS_ORACLE_H_
#define ARTEMIS_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* src/artemis.hpp:78-105 enum order */
enum { AO_CARTESIAN = 0, AO_CYLINDRICAL = 1, AO_SPHERICAL1D = 2, AO_SPHERICAL2D = 3,
       AO_SPHERICAL3D = 4, AO_AXISYMMETRIC = 5 };
enum { AO_HLLC = 0, AO_HLLE = 1, AO_LLF = 2 };
enum { AO_PCM = 0, AO_PLM = 1, AO_PPM = 2 };
enum { AO_GAS = 0, AO_DUST = 1 };
enum { AO_BC_PERIODIC = 0, AO_BC_OUTFLOW = 1, AO_BC_REFLECT = 2, AO_BC_NONE = 3,
       AO_BC_IC = 4 /* user condition `ic`: the initial-condition profile (src/pgen/disk.hpp:595-633) */,
       /* shearing-box user conditions of src/pgen/strat.hpp (inputs/ssheet/ssheet.in) */
       AO_BC_EXTRAP = 5, AO_BC_INFLOW = 6 };

typedef struct {
  int geom, ndim, ng, nb;
  int ni, nj, nk;             /* allocated cells per block (ghosts included)      */
  int is, ie, js, je, ks, ke; /* interior bounds (inclusive)                      */
  int fni, fnj, fnk;          /* allocated dims of the face-velocity field        */
  const double *xmin;         /* [nb][3] UniformCartesian::xmin_ (ghost-shifted)  */
  const double *dx;           /* [nb][3]                                          */
} ao_grid;

typedef struct {
  int fluid;    /* AO_GAS / AO_DUST */
  int nspecies;
  int recon, riemann;
  double gm1;   /* gamma - 1 (gas)  */
  double dfloor, siefloor, de_switch, cfl;
} ao_fluid;

void ao_calculate_fluxes(const ao_grid *g, const ao_fluid *f, int pcm, const double *prim,
                         double *flux1, double *flux2, double *flux3, double *pflux1,
                         double *pflux2, double *pflux3, double *vface1, double *vface2,
                         double *vface3);
void ao_apply_update(const ao_grid *g, int nvar, double *u0, const double *u1,
                     const double *flux1, const double *flux2, const double *flux3,
                     double gam0, double gam1, double beta_dt);
void ao_flux_source(const ao_grid *g, const ao_fluid *f, const double *prim, double *cons,
                    const double *pflux1, const double *pflux2, const double *pflux3,
                    const double *vface1, const double *vface2, const double *vface3,
                    double omf, double dt);
void ao_set_aux(const ao_grid *g, const ao_fluid *f, double *cons);
void ao_cons_to_prim(const ao_grid *g, const ao_fluid *f, const double *cons, double *prim);
void ao_prim_to_cons(const ao_grid *g, const ao_fluid *f, double *prim, double *cons);
double ao_estimate_dt(const ao_grid *g, const ao_fluid *f, const double *prim);
void ao_deep_copy(const ao_grid *g, int nvar, double *to, const double *from);

/* Same-level ghost exchange on a uniform nbx x nby x nbz block lattice + physical BCs.
 * vars: list of nv pack indices into `a` ([nb][nvar][nk][nj][ni]) to communicate;
 * vec_dir[v] = 1,2,3 if var v is that component of a vector (sign flip under reflect),
 * else 0.  bc[6] = {ix1, ox1, ix2, ox2, ix3, ox3}. */
void ao_exchange_ghosts(const ao_grid *g, int nbx, int nby, int nbz, const int *bc,
                        int nvar, double *a, int nv, const int *vars, const int *vec_dir);

/* Same, split in phases (bit 0: neighbour copies, bit 1: physical BCs); faces flagged
 * AO_BC_NONE (owned by another rank) are left untouched. */
void ao_exchange_ghosts_phase(const ao_grid *g, int nbx, int nby, int nbz, const int *bc,
                              int nvar, double *a, int nv, const int *vars,
                              const int *vec_dir, int phases);
/* the same with user `ic` faces (AO_BC_IC): `ic` holds the initial-condition profile in every
 * zone (same shape as `a`); an ic face sets all its ghost zones, over the full transverse extent,
 * to the profile at their own position -- Disk::DiskBoundaryIC's par_for_bndry */
void ao_exchange_ghosts_ic(const ao_grid *g, int nbx, int nby, int nbz, const int *bc, int nvar,
                           double *a, int nv, const int *vars, const int *vec_dir, int phases,
                           const double *ic);

/* strat.hpp:154-666 on one block array (see artemis_oracle.c); 0 = applied */
int ao_strat_bc(int geom, const double *xmin, const double *dx, int ni, int nj, int nk,
                const int *s, const int *e, int fluid, int S, double *a, int face, int type,
                double q, double om0);
/* neighbour copies + every physical face in x1 -> x2 -> x3 order, user faces included */
void ao_exchange_ghosts_user(const ao_grid *g, int nbx, int nby, int nbz, const int *bc, int nvar,
                             double *a, int nv, const int *vars, const int *vec_dir,
                             const double *ic, int fluid, int S, double q, double om0);

/* Geometry probes (used by tests to compare against oracle/_ref and the CUDA tables) */
void ao_geom_cell(int geom, const double *xmin, const double *dx, int k, int j, int i,
                  double *out /* [32] */);

/* Single-function probes */
void ao_plm(double qm, double q, double qp, double *ql_ip1, double *qr_i);
void ao_plm_g(double qm, double q, double qp, double xm, double xc, double xp,
              double xf0, double xf1, double dx, double *ql_ip1, double *qr_i);
void ao_ppm4(double qm2, double qm1, double q, double qp1, double qp2, double *ql_ip1,
             double *qr_i);
/* wl/wr: [6] (gas: rho,vx,vy,vz,P,sie) or [4] (dust); out: gas [8] = Frho,Fmx,Fmy,Fmz,
 * FE, Fu, pface, vface ; dust [4] */
void ao_riemann(int solver, int fluid, double gm1, const double *wl, const double *wr,
                double *out);

int ao_num_threads(void);

/* ---- multilevel operators ------------------------------------------ */
/* One MeshBlock: fine array [nvar][nk][nj][ni] and its coarse buffer [nvar][cnk][cnj][cni]
 * (Parthenon's c_cellbounds: nx/2 interior cells + ng ghosts in every active direction,
 * P:mesh/meshblock.cpp:205-228).  box = {cis, cie, cjs, cje, cks, cke}: inclusive COARSE index
 * range the operator loops over.  xmin/dx: the block's fine UniformCartesian. */
typedef struct {
  int geom, ndim, ng;
  int ni, nj, nk;
  int cni, cnj, cnk;
  int ib_s, jb_s, kb_s;     /* fine interior start   */
  int cib_s, cjb_s, ckb_s;  /* coarse interior start */
  double xmin[3], dx[3];
} ao_refine_geom;
/* restriction.hpp:41-114: volume-weighted average of the 2^ndim fine cells */
void ao_restrict_average(const ao_refine_geom *r, int nvar, const double *fine, double *coarse,
                         const int *box);
/* the same on a face-centred (flux) field, el = 1..3 (flux correction): area-weighted average
 * of the 2^(ndim-1) fine faces */
void ao_restrict_average_face(const ao_refine_geom *r, int nvar, const double *fine,
                              double *coarse, const int *box, int el);
/* prolongation.hpp:82-184: minmod-limited linear interpolation onto the 2^ndim fine cells */
void ao_prolongate_minmod(const ao_refine_geom *r, int nvar, const double *coarse, double *fine,
                          const int *box);

struct ao_diffusion_s;
typedef struct ao_diffusion_s ao_diffusion_fwd;
/* ---- pointwise source terms between FluxSource and SetAuxillaryFields ----
 * gas / dust may be NULL (fluid absent).  They read the stage-start primitives and update the
 * conserved state of interior zones in place. */
/* Gravity::UniformGravity<GEOM>, src/gravity/uniform.cpp:28-90 */
void ao_uniform_gravity(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                        const ao_fluid *dust, const double *dprim, double *dcons, double dt,
                        double gx1, double gx2, double gx3);
/* RotatingFrame::ShearingBoxImpl, src/rotating_frame/rotating_frame_impl.hpp:28-94 */
void ao_shearing_box(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                     const ao_fluid *dust, const double *dprim, double *dcons, double dt,
                     double om0, double qshear);
/* Gravity::PointMassGravity<GEOM>, src/gravity/point_mass.cpp:26-196;
 * pm = {gm, x, y, z, soft, sink_rate, sink} */
void ao_point_mass_gravity(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                           double *gcons, const ao_fluid *dust, const double *dprim,
                           double *dcons, double dt, const double *pm);
/* RotatingFrame::RotatingFrameImpl<GEOM> (non-Cartesian), rotating_frame_impl.hpp:96-199;
 * g/dflux{1,2,3}: the [nb][nvar][cells] flux slabs of the stage (unused directions may be NULL) */
void ao_rotating_frame(const ao_grid *g, const ao_fluid *gas, double *gcons,
                       const double *gflux1, const double *gflux2, const double *gflux3,
                       const ao_fluid *dust, double *dcons, const double *dflux1,
                       const double *dflux2, const double *dflux3, double dt, double om0);
/* Drag::SimpleDragSourceImpl (constant stopping times, no damping zones, no viscous target
 * velocity), src/drag/drag.hpp:296-482 */
void ao_drag_simple(const ao_grid *g, const ao_fluid *gas, double *gcons, const ao_fluid *dust,
                    double *dcons, double dt, const double *tau);

/* Drag::DragSource<GEOM> in full (src/drag/drag.cpp:88-165 dispatch, src/drag/drag.hpp:144-482):
 * coupling simple_dust (implicit gas-dust drag, constant or Stokes stopping times) or self
 * (damping zones only); quadratic damping ramps of gas and dust towards rest or, for the gas with
 * damp_to_visc, towards the viscous inflow velocity -1.5 nu / R. */
enum { AO_DRAG_SIMPLE_DUST = 0, AO_DRAG_SELF = 1 };
enum { AO_DRAG_CONSTANT = 0, AO_DRAG_STOKES = 1 };
typedef struct {
  int coupling, model;
  double tau[16], scale;             /* <dust/stopping_time> tau (constant), scale          */
  double grain_density, sizes[16];   /* <dust> grain_density, sizes (Stokes)                */
  double g_ix[3], g_ox[3], g_irate[3], g_orate[3]; /* <gas/damping> inner_x*, outer_x*, rates */
  int g_damp_to_visc;
  double d_ix[3], d_ox[3], d_irate[3], d_orate[3]; /* <dust/damping>                        */
  double xmin[3], xmax[3];           /* <parthenon/mesh> x*min, x*max                       */
} ao_drag;
/* dd: the viscosity parameters (read only with g_damp_to_visc; may be NULL otherwise) */
void ao_drag_source(const ao_grid *g, const ao_fluid *gas, double *gcons, const ao_fluid *dust,
                    double *dcons, const ao_drag *dp, const ao_diffusion_fwd *dd, double dt);

/* ---- diffusion operators: src/utils/diffusion/{diffusion,diffusion_coeff,
 * momentum_diffusion,thermal_diffusion}.hpp driven by Gas::{ZeroDiffusionFlux,ViscousFlux,
 * ThermalFlux,DiffusionUpdate} (src/gas/gas.cpp:524-642) and the diffusive timestep limits of
 * Gas::EstimateTimestepMesh (src/gas/gas.cpp:437-467). */
enum { AO_DIFF_NONE = 0, AO_VISC_PLAW = 1, AO_VISC_ALPHA = 2 };
enum { AO_COND_NONE = 0, AO_COND_CONDUCTIVITY = 1, AO_COND_DIFFUSIVITY = 2 };
enum { AO_AVG_ARITHMETIC = 0, AO_AVG_HARMONIC = 1 };
typedef struct ao_diffusion_s {
  int visc_type, visc_avg;           /* DiffCoeffParams of <gas/viscosity>            */
  double nu, eta, r0, r_exp;         /*   plaw: nu_s, eta_bulk, problem/r0, r_exp      */
  double alpha, omega0;              /*   alpha: alpha, Omega0 = sqrt(gm / r0^3)       */
  int cond_type, cond_avg;           /* DiffCoeffParams of <gas/conductivity>         */
  double cond, kappa, temp_exp, rho_exp, rho_ref, t_ref;
  double cv;                         /* specific heat of the ideal-gas EOS            */
} ao_diffusion;
/* diffusion flux slabs: dflx{1,2,3} = [nb][4S][fnk][fnj][fni] face arrays (gas.diff.momentum 3S
 * entries then gas.diff.energy S entries, Metadata::Face: src/gas/gas.cpp:277-285) */
void ao_diffusion_flux(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                       const ao_diffusion *dd, double *dflx1, double *dflx2, double *dflx3);
void ao_diffusion_update(const ao_grid *g, const ao_fluid *gas, const double *gprim, double *gcons,
                         const ao_diffusion *dd, const double *dflx1, const double *dflx2,
                         const double *dflx3, double dt);
/* min(visc_dt, cond_dt) BEFORE the cfl factor (src/gas/gas.cpp:437-464) */
double ao_diffusion_dt(const ao_grid *g, const ao_fluid *gas, const double *gprim,
                       const ao_diffusion *dd);

#ifdef __cplusplus
}
#endif
#endif
