/*
 * ab200.h -- C ABI of libartemis_b200: the B200-native (sm_100a, hand-written fp64 CUDA)
 * replacement for the data-parallel hot path of lanl/artemis.
 *
 * The reference has no FFI today: its hot path is a set of Parthenon *task functions*
 * (TaskStatus f(MeshData<Real>*, ...)) added to the per-stage TaskList in
 * src/artemis_driver.cpp:166-270.  Each entry point below replaces the body of one of those
 * task functions; the file:line it replaces is cited on every declaration.  A ~150-line
 * glue TU compiled inside Artemis flattens the SparsePack it already
 * builds into the pointer tables of ab200_pack_desc and calls these functions.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch/Kokkos types.
 *  - Every function returns 0 on success, non-zero on failure (AB200_E*); the message is
 *    available from ab200_last_error().  The glue maps non-zero to TaskStatus::fail /
 *    PARTHENON_FAIL (P:basic_types.hpp:59).  Nothing throws across the boundary.
 *  - The library BORROWS every state array (Parthenon owns them); it owns only its pointer
 *    tables, metric tables, halo slabs and reduction scratch.
 *  - All work is enqueued on the stream given at ab200_create (pass the raw cudaStream_t of
 *    Kokkos::Cuda().cuda_stream(), or NULL for the legacy default stream) and is
 *    asynchronous unless the function returns a host scalar.  Callable from any host
 *    thread: every entry point does cudaSetDevice(ctx->device).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Variable layout ("MeshBlockPack layout", P:interface/variable.cpp:112-128): each
 * (block, pack-index) entry of a pointer table addresses one dense 3-D array
 * [nk][nj][ni] (i fastest, ghosts included) in DEVICE memory.  Pack index conventions are
 * the reference's (src/utils/fluxes/riemann/hllc.hpp:66-73), S = nspecies:
 *    gas  prim: rho n | vel S+3n+d | pressure 4S+n | sie 5S+n            (6S entries)
 *    gas  cons: rho n | mom S+3n+d | total_energy 4S+n | internal 5S+n   (6S entries)
 *    dust prim: rho n | vel S+3n+d        dust cons: rho n | mom S+3n+d  (4S entries)
 * flux[d] uses the cons numbering; flux(i) is the LOWER face of cell i
 * (src/utils/fluxes/fluid_fluxes.hpp:114-121).  pflux[d] is the flux slot of
 * gas.prim.pressure (interface pressure), vface[d] is gas.face.velocity element d with
 * allocated dims fnk x fnj x fni (P:interface/metadata.cpp:378-387).
 */
#ifndef AB200_H_
#define AB200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB200_ABI_VERSION 12

/* src/artemis.hpp:78-105 */
enum { AB200_CARTESIAN = 0, AB200_CYLINDRICAL = 1, AB200_SPHERICAL1D = 2,
       AB200_SPHERICAL2D = 3, AB200_SPHERICAL3D = 4, AB200_AXISYMMETRIC = 5 };
enum { AB200_HLLC = 0, AB200_HLLE = 1, AB200_LLF = 2 };
enum { AB200_PCM = 0, AB200_PLM = 1, AB200_PPM = 2 };
enum { AB200_GAS = 0, AB200_DUST = 1 };
enum { AB200_BC_PERIODIC = 0, AB200_BC_OUTFLOW = 1, AB200_BC_REFLECT = 2,
       AB200_BC_NONE = 3 /* face handled by the caller (remote rank / user BC) */,
       /* User boundary condition that is a function of POSITION only -- `ic` of the disk and
        * strat problem generators (Disk::DiskBoundaryIC, src/pgen/disk.hpp:595-633: every ghost
        * zone of the face, over the full transverse extent, receives the initial-condition
        * profile at its own position; inputs/disk/disk_sph.in uses it on four faces).  Such
        * ghost zones never change after Mesh::Initialize applied the condition once, so the
        * library keeps them resident: zones beyond a FIXED face are never written, except
        * where a LATER face in Parthenon's x1 -> x2 -> x3 order (outflow / reflect) covers
        * them, which then copies from the fixed zones as the reference does.  Single-rank
        * topologies (no AB200_BC_NONE face). */
       AB200_BC_FIXED = 4,
       /* State-dependent user conditions of the shearing-box problem generators
        * (src/pgen/strat.hpp, registered in src/pgen/problem_modifier.hpp:114-127;
        * inputs/ssheet/ssheet.in uses all six).  They exist as ab200_block_bc_desc types only
        * (ab200_block_bcs applies them per block, on fine arrays or coarse buffers); in
        * ab200_set_topology such a face is AB200_BC_NONE, i.e. the caller's.
        *   EXTRAP on x1 faces: strat::ExtrapInnerX1 / ExtrapOuterX1 (strat.hpp:154-297)
        *   EXTRAP on x3 faces: strat::ExtrapInnerX3 / ExtrapOuterX3 (strat.hpp:487-663)
        *   INFLOW on x2 faces: strat::ShearInnerX2 / ShearOuterX2   (strat.hpp:299-485) */
       AB200_BC_EXTRAP = 5, AB200_BC_INFLOW = 6 };

enum { AB200_OK = 0, AB200_EINVAL = 1, AB200_ECUDA = 2, AB200_ESTATE = 3, AB200_ENOMEM = 4 };

typedef struct ab200_ctx ab200_ctx;

/* Mesh/index description of one MeshData partition (all blocks share one shape).
 * xmin/dx are Parthenon's UniformCartesian xmin_/dx_ per block
 * (P:coordinates/uniform_cartesian.hpp:30-36): Xf(idx) = xmin + idx*dx. */
typedef struct ab200_grid_desc {
  int geom, ndim, nghost, nblocks;
  int ni, nj, nk;             /* allocated cells per block, ghosts included */
  int is, ie, js, je, ks, ke; /* interior bounds, inclusive                 */
  int fni, fnj, fnk;          /* allocated dims of gas.face.velocity        */
  const double *xmin;         /* HOST [nblocks][3]                          */
  const double *dx;           /* HOST [nblocks][3]                          */
} ab200_grid_desc;

/* StateDescriptor params read by the hot path (src/gas/gas.cpp:59-208,
 * src/dust/dust.cpp:44-100). */
typedef struct ab200_fluid_desc {
  int fluid, nspecies, recon, riemann;
  double gm1, dfloor, siefloor, de_switch, cfl;
} ab200_fluid_desc;

/* HOST arrays of DEVICE pointers; entry [b*nvar + n] (or [b*S + n]). NULL tables are
 * allowed for cons1/flux/pflux/vface: the library then allocates its own scratch when an
 * entry point needs them. */
typedef struct ab200_pack_desc {
  double *const *prim;
  double *const *cons0; /* u0 */
  double *const *cons1; /* u1 */
  double *const *flux[3];
  double *const *pflux[3];
  double *const *vface[3];
} ab200_pack_desc;

/* One ghost-zone buffer, the analogue of parthenon::BndInfo (P:bvals/comms/bnd_info.hpp,
 * index ranges from CalcIndices P:bvals/comms/bnd_info.cpp:105-252).  The sub-box
 * [sk..ek][sj..ej][si..ei] of `ncomp` consecutive pack entries starting at `var0` of block
 * `block` is (un)packed to/from `buf` in [comp][k][j][i] order (i fastest,
 * P:utils/indexer.hpp:119-131). */
typedef struct ab200_bnd_desc {
  int fluid, block, var0, ncomp;
  int si, ei, sj, ej, sk, ek;
  double *buf; /* DEVICE */
} ab200_bnd_desc;

/* ---- lifetime ------------------------------------------------------------------------ */
int ab200_abi_version(void);
const char *ab200_last_error(void);
int ab200_device_count(void);
int ab200_create(ab200_ctx **ctx, int device, void *cuda_stream);
int ab200_destroy(ab200_ctx *ctx);
int ab200_synchronize(ab200_ctx *ctx);

/* ---- binding (rebuilt when the SparsePack cache is invalidated,
 *      P:interface/sparse_pack_base.cpp:320-344) ------------------------------------------ */
int ab200_set_grid(ab200_ctx *ctx, const ab200_grid_desc *grid);
int ab200_bind_pack(ab200_ctx *ctx, const ab200_fluid_desc *fluid, const ab200_pack_desc *pack);
int ab200_unbind(ab200_ctx *ctx, int fluid);
/* rotating_frame "omega" used by FluxSource (src/utils/fluxes/fluid_fluxes.hpp:433-437) */
int ab200_set_rotating_frame(ab200_ctx *ctx, double omega);

/* ---- task functions ---------------------------------------------------------------------- */
/* Gas::CalculateFluxes / Dust::CalculateFluxes -> ArtemisUtils::CalculateFluxesImpl
 * (src/gas/gas.cpp:473-494, src/dust/dust.cpp:281-298, fluid_fluxes.hpp:76-213) */
int ab200_calculate_fluxes(ab200_ctx *ctx, int fluid, int pcm);
/* ArtemisUtils::ApplyUpdate<GEOM> over every bound fluid
 * (src/utils/integrators/artemis_integrator.hpp:56-110) */
int ab200_apply_update(ab200_ctx *ctx, double gam0, double gam1, double beta_dt);
/* Gas::FluxSource / Dust::FluxSource -> FluxSourceImpl
 * (src/gas/gas.cpp:499-519, src/dust/dust.cpp:303-326, fluid_fluxes.hpp:298-420) */
int ab200_flux_source(ab200_ctx *ctx, int fluid, double dt);
/* ArtemisDerived::SetAuxillaryFields<GEOM> (src/derived/fill_derived.cpp:29-75) */
int ab200_set_auxillary_fields(ab200_ctx *ctx);
/* ArtemisDerived::ConsToPrim<GEOM> (fill_derived.cpp:81-167), interior */
int ab200_cons_to_prim(ab200_ctx *ctx);
/* ArtemisDerived::PrimToCons<T,GEOM> (fill_derived.cpp:172-277), entire domain */
int ab200_prim_to_cons(ab200_ctx *ctx);
/* ArtemisUtils::DeepCopyConservedData u1 <- u0 (artemis_integrator.hpp:30-51) */
int ab200_deep_copy_conserved(ab200_ctx *ctx);
/* Gas/Dust::EstimateTimestepMesh<GEOM> (src/gas/gas.cpp:391-468, src/dust/dust.cpp:238-276):
 * returns cfl * min over the partition in *dt_host (synchronises the stream). */
int ab200_estimate_timestep(ab200_ctx *ctx, int fluid, double *dt_host);

/* ---- fused fast path ----------------------------------------------------------------------
 * One call == CalculateFluxes + ApplyUpdate + FluxSource + SetAuxillaryFields + ConsToPrim +
 * (interior part of) PrimToCons for every bound fluid, i.e. tasks
 * src/artemis_driver.cpp:184-255 plus the interior of :261, valid when no out-of-scope
 * source term (drag/gravity/rotating-frame/cooling/diffusion, :217-248) sits in between.
 * One fused pass per direction: reconstruct -> Riemann -> flux difference -> update; flux
 * arrays are never materialised.  If stage1_copy != 0 the pass also writes u1 <- u0
 * (DeepCopyConservedData folded in; requires gam0 == 0, gam1 == 1).
 * flags: AB200_STAGE_DEVICE_DT  dt is read from the device scalar ab200_dt_device()[0];
 *        AB200_STAGE_REDUCE_DT  (last stage of a cycle) also leave cfl * min(dt) of every bound
 *                               fluid in ab200_dt_device()[1], i.e. Gas/Dust::
 *                               EstimateTimestepMesh (src/gas/gas.cpp:391-468) over the new
 *                               interior primitives; folded into the kernel that writes them.
 *        AB200_STAGE_PINGPONG   3-D Cartesian meshes run the stage as ONE kernel (all three
 *                               directions, primitives and conserved state cross HBM once) that
 *                               reads one primitive set and writes another, because tiles of a
 *                               MeshBlock read each other's zones as halo.  Without this flag the
 *                               new interior primitives are copied back into the caller's arrays
 *                               before the call returns.  With it they stay in a library-owned
 *                               alternate set that every ghost-zone / timestep entry point below
 *                               follows transparently, and the caller's arrays are current again
 *                               after an even number of stages or after ab200_sync_prim(). */
#define AB200_STAGE_DEVICE_DT 1
#define AB200_STAGE_REDUCE_DT 2
#define AB200_STAGE_PINGPONG 4
/*        AB200_STAGE_DEFER_C2P  the passes only update the conserved state (ApplyUpdate +
 *                               FluxSource); SetAuxillaryFields / ConsToPrim / PrimToCons and the
 *                               timestep wait for ab200_finish_stage, so that source terms
 *                               (src/artemis_driver.cpp:217-248: the library's ab200_uniform_gravity /
 *                               ab200_shearing_box / ab200_drag_simple, or the host's own Kokkos
 *                               kernels) can act on the conserved state in between.  Takes the
 *                               directional passes (the single-pass kernels always finish). */
#define AB200_STAGE_DEFER_C2P 8
/*        AB200_STAGE_SURFACE / AB200_STAGE_INTERIOR  the call covers only the MeshBlocks that
 *                               touch a face owned by another rank (AB200_BC_NONE in
 *                               ab200_set_topology) / only the others.  A stage is then TWO calls,
 *                               surface first: the remote ghost exchange (ab200_comm_exchange_begin)
 *                               starts as soon as the surface blocks are done and overlaps the
 *                               interior blocks' stage.  Directional passes only. */
#define AB200_STAGE_SURFACE 16
#define AB200_STAGE_INTERIOR 32
/*   AB200_STAGE_TAP_DFLUX      every directional pass also stores its MASS flux (one array per
 *                               species and direction, library-owned) for ab200_rotating_frame:
 *                               the only part of the reference's 21 flux arrays per zone that a
 *                               source term reads (rotating_frame_impl.hpp:135-163).  Curvilinear
 *                               systems only (AB200_EINVAL on a Cartesian mesh, whose rotating
 *                               frame is the shearing box and reads no flux). */
#define AB200_STAGE_TAP_DFLUX 64
int ab200_fused_stage(ab200_ctx *ctx, double gam0, double gam1, double beta, double dt,
                      int pcm, int stage1_copy, int flags);
/* SetAuxillaryFields -> ConsToPrim -> PrimToCons after a AB200_STAGE_DEFER_C2P stage and its
 * source terms; flags: AB200_STAGE_REDUCE_DT (last stage of a cycle: Gas/Dust::
 * EstimateTimestepMesh of the new primitives into ab200_dt_device()[1]). */
int ab200_finish_stage(ab200_ctx *ctx, int flags);

/* ---- pointwise source terms between FluxSource and SetAuxillaryFields --------
 * They read the stage-start primitives and add to the conserved state of interior zones of every
 * bound fluid; dt = beta * dt of the stage (src/artemis_driver.cpp:217-248).
 *   ab200_uniform_gravity  Gravity::UniformGravity<GEOM>   src/gravity/uniform.cpp:28-90
 *   ab200_shearing_box     RotatingFrame::ShearingBoxImpl  src/rotating_frame/rotating_frame_impl.hpp:28-94
 *                          (Cartesian)
 *   ab200_rotating_frame   RotatingFrame::RotatingFrameImpl<GEOM> :96-199, every curvilinear
 *                          system (rotating_frame.cpp:69-82).  Reads the mass fluxes of the stage:
 *                          call after ab200_fused_stage(.. | AB200_STAGE_DEFER_C2P |
 *                          AB200_STAGE_TAP_DFLUX) or after ab200_calculate_fluxes; AB200_ESTATE
 *                          otherwise.  Updates momentum and total energy only.
 *   ab200_point_mass_gravity  Gravity::PointMassGravity<GEOM>  src/gravity/point_mass.cpp:26-196:
 *                          softened point mass at a Cartesian position + the mass sink inside
 *                          `sink` (rate `sink_rate`; 0 disables it)
 *   ab200_drag_simple      Drag::SimpleDragSourceImpl      src/drag/drag.hpp:296-482 with constant
 *                          stopping times tau[n] per dust species (<drag/dust> type = constant),
 *                          no damping zones and no viscous target velocity (inputs/drag/simple_drag.in)
 *   ab200_drag_source      Drag::DragSource<GEOM> in full (see ab200_drag_desc below) */
int ab200_uniform_gravity(ab200_ctx *ctx, double dt, double gx1, double gx2, double gx3);
int ab200_shearing_box(ab200_ctx *ctx, double dt, double omega, double qshear);
int ab200_drag_simple(ab200_ctx *ctx, double dt, int ntau, const double *tau);
/* Drag::DragSource<GEOM> in full (src/drag/drag.cpp:88-165, src/drag/drag.hpp:144-482): the
 * parameters of <drag>, <dust/stopping_time>, <dust> sizes / grain_density, <gas/damping>,
 * <dust/damping> and the mesh bounds the damping ramps are normalised with.  Unused damping
 * directions: inner = -DBL_MAX, outer = +DBL_MAX, rates 0 (the deck defaults).  With
 * g_damp_to_visc the gas is damped towards the viscous inflow velocity -1.5 nu / R of the
 * viscosity configured through ab200_configure_diffusion (AB200_ESTATE without one). */
#define AB200_DRAG_SIMPLE_DUST 0  /* <drag> type = simple_dust: implicit gas-dust coupling */
#define AB200_DRAG_SELF 1         /* type = self: damping zones only                      */
#define AB200_DRAG_CONSTANT 0     /* <dust/stopping_time> type = constant                  */
#define AB200_DRAG_STOKES 1       /* type = stokes: tau = scale rho_s s / (rho_g v_th)     */
typedef struct ab200_drag_desc {
  int coupling, model;
  double tau[16], scale;
  double grain_density, sizes[16];
  double g_ix[3], g_ox[3], g_irate[3], g_orate[3];
  int g_damp_to_visc;
  double d_ix[3], d_ox[3], d_irate[3], d_orate[3];
  double xmin[3], xmax[3];
} ab200_drag_desc;
int ab200_drag_source(ab200_ctx *ctx, double dt, const ab200_drag_desc *drag);
typedef struct ab200_point_mass_desc {
  double gm;          /* <gravity/point> gm                                  */
  double x, y, z;     /* Cartesian position of the mass                      */
  double soft;        /* softening length                                    */
  double sink_rate;   /* mass removal rate inside the sink radius (per time) */
  double sink;        /* sink radius                                         */
} ab200_point_mass_desc;
int ab200_point_mass_gravity(ab200_ctx *ctx, double dt, const ab200_point_mass_desc *pm);
int ab200_rotating_frame(ab200_ctx *ctx, double dt, double omega);
/* Source terms the device-resident drivers (ab200_run_cycles, ab200_run_cycles_mr,
 * ab200_cycles_host) apply every stage, in the reference's task order gravity -> rotating frame
 * -> drag; NULL or all-zero switches them off (the stage then runs un-split). */
typedef struct ab200_sources_desc {
  int gravity;       double g[3];            /* Gravity::UniformGravity          */
  int shearing_box;  double omega, qshear;   /* RotatingFrame::ShearingBoxImpl   */
  int drag;          int ntau; double tau[16]; /* Drag::SimpleDragSourceImpl      */
  int point_mass;    ab200_point_mass_desc pm; /* Gravity::PointMassGravity (instead of `gravity`) */
  int rotating_frame; double rf_omega;       /* RotatingFrame::RotatingFrameImpl (curvilinear) */
  int drag_model;    ab200_drag_desc drag_desc; /* Drag::DragSource in full (instead of `drag`) */
} ab200_sources_desc;
int ab200_configure_sources(ab200_ctx *ctx, const ab200_sources_desc *src);

/* ---- diffusion operators of the gas -------------------------------------------
 * Viscous stress (constant / power-law / alpha viscosity, bulk viscosity) and heat conduction
 * (power-law conductivity or thermal diffusivity) of src/utils/diffusion/{diffusion_coeff,
 * momentum_diffusion,thermal_diffusion,diffusion}.hpp, as the reference's driver applies them
 * every stage (src/artemis_driver.cpp:188-196, 217-221):
 *   ab200_diffusion_flux      Gas::ZeroDiffusionFlux + Gas::ViscousFlux + Gas::ThermalFlux
 *                             (src/gas/gas.cpp:524-603) in ONE kernel: the fluxes of
 *                             gas.diff.momentum / gas.diff.energy through every face of the
 *                             interior zones from the current primitives (ghost zones valid to
 *                             depth 2), into library-owned face arrays
 *   ab200_diffusion_update    Gas::DiffusionUpdate (src/gas/gas.cpp:608-642): momentum, total and
 *                             internal energy of the interior zones -= dt * div(flux) (+ the
 *                             curvilinear stress source terms)
 *   ab200_diffusion_timestep  cfl * min(viscous, conductive limit), the diffusive part of
 *                             Gas::EstimateTimestepMesh (src/gas/gas.cpp:437-467)
 * Once configured, ab200_estimate_timestep[_device](AB200_GAS) includes the diffusive limits and
 * the device-resident drivers (ab200_run_cycles, ab200_run_cycles_mr, ab200_cycles_host) apply
 * flux + update every stage, after FluxSource and before the gravity / rotating-frame / drag
 * sources.  Parameters mirror Diffusion::DiffCoeffParams (diffusion_coeff.hpp:59-146). */
#define AB200_VISC_NONE 0
#define AB200_VISC_PLAW 1         /* <gas/viscosity> type = constant | powerlaw */
#define AB200_VISC_ALPHA 2        /* type = alpha: nu = alpha c_s^2 / Omega_K   */
#define AB200_COND_NONE 0
#define AB200_COND_CONDUCTIVITY 1 /* <gas/conductivity> type = conductivity     */
#define AB200_COND_DIFFUSIVITY 2  /* type = diffusivity                         */
#define AB200_AVG_ARITHMETIC 0
#define AB200_AVG_HARMONIC 1
typedef struct ab200_diffusion_desc {
  int visc_type, visc_avg;
  double nu, eta_bulk, r0, r_exp;  /* plaw: nu (r/r0)^r_exp; eta_bulk = bulk / shear          */
  double alpha, omega0;            /* alpha; omega0 = sqrt(gm / r0^3) (diffusion_coeff.hpp:113) */
  int cond_type, cond_avg;
  double cond, kappa, temp_exp, rho_exp, rho_ref, t_ref;
  double cv;                       /* specific heat of the ideal-gas EOS (T = sie / cv)       */
} ab200_diffusion_desc;
int ab200_configure_diffusion(ab200_ctx *ctx, const ab200_diffusion_desc *dd); /* NULL: off */
int ab200_diffusion_flux(ab200_ctx *ctx);
int ab200_diffusion_update(ab200_ctx *ctx, double dt);
int ab200_diffusion_timestep(ab200_ctx *ctx, double *dt_host);
/* device pointer and element count of the library-owned flux array of direction dir (1..3):
 * [nblocks][4 nspecies][fnk][fnj][fni], entries 3n..3n+2 = gas.diff.momentum of species n,
 * 3 nspecies + n = gas.diff.energy (Metadata::Face arrays, src/gas/gas.cpp:277-285) */
int ab200_diffusion_flux_array(ab200_ctx *ctx, int dir, double **dev_ptr, size_t *count);

/* ---- history reductions -------------------------------------------------
 * out_host[v] = sum over the interior zones of the partition of cons0[v] * Volume, for every
 * conserved pack entry v of `fluid` (nout = its pack size): the integrals behind gas_mass,
 * gas_momentum_x{1,2,3}, gas_energy, gas_internal_energy and the dust ones --
 * ArtemisUtils::ReduceSpeciesVolumeIntegral / ReduceSpeciesVectorVolumeIntegral
 * (src/utils/history.hpp:24-95, registered in src/gas/gas.cpp:650-675).  One launch for all
 * entries + one ordered finishing kernel (deterministic); synchronises the stream. */
int ab200_history_volume_integrals(ab200_ctx *ctx, int fluid, double *out_host, int nout);

/* Which kernels ab200_fused_stage runs on meshes where both exist (3-D Cartesian, TMA-able
 * arrays); every other mesh always takes the directional passes.
 *   AB200_PATH_AUTO         the faster of the two as measured on B200 for the bound fluid's
 *                           reconstruction / Riemann solver
 *   AB200_PATH_THREE_PASS   one kernel per direction (x1, x2, x3)
 *   AB200_PATH_SINGLE_PASS  the single-pass stage kernel, one role per CTA (sweep.cuh; requires
 *                           the alternate primitive set)
 *   AB200_PATH_ROLE_SPLIT   the single-pass stage kernel, warp-specialised: x1 / x2 / x3+update
 *                           warp groups pipelined over named barriers (trio.cuh; same
 *                           requirements, same results bit for bit as SINGLE_PASS)
 * The paths agree to the parity bar (1e-12 per zone and cycle); in the strict build the
 * single-pass kernel is bit-identical to the reference (it sums the flux divergence over the
 * three directions before the update, artemis_integrator.hpp:95-106), the directional passes
 * round once per direction. */
#define AB200_PATH_AUTO 0
#define AB200_PATH_THREE_PASS 1
#define AB200_PATH_SINGLE_PASS 2
#define AB200_PATH_ROLE_SPLIT 3
int ab200_set_stage_path(ab200_ctx *ctx, int path);
/* *path_out = AB200_PATH_THREE_PASS, AB200_PATH_SINGLE_PASS or AB200_PATH_ROLE_SPLIT: what ab200_fused_stage runs for
 * `fluid` on the bound mesh under the current setting (resolves AUTO and eligibility). */
int ab200_get_stage_path(ab200_ctx *ctx, int fluid, int *path_out);
/* Copies the current primitives (interior and ghosts) into the caller's arrays if a
 * AB200_STAGE_PINGPONG stage left them in the alternate set; no-op otherwise.  Entry points
 * that hand the caller's primitive arrays to task-level kernels call it implicitly. */
int ab200_sync_prim(ab200_ctx *ctx);
/* PrimToCons restricted to ghost zones (completes :261 after the exchange). */
int ab200_prim_to_cons_ghosts(ab200_ctx *ctx);
/* Lazy conserved ghost zones.  The fused stage kernels read conserved variables of interior
 * zones only, so between stages the ghost part of PrimToCons (fill_derived.cpp:217-274 over
 * the ENTIRE domain) is dead work.  With lazy != 0, ab200_fill_ghosts / _local /
 * ab200_finish_remote_ghosts write the ghost primitives only and mark the conserved ghost zones
 * stale; ab200_sync_ghost_cons converts them once (no-op when nothing is stale).
 * ab200_run_cycles and ab200_cycles_host do this internally and return complete arrays. */
int ab200_set_ghost_cons_lazy(ab200_ctx *ctx, int lazy);
int ab200_sync_ghost_cons(ab200_ctx *ctx);

/* ---- multilevel ghost exchange operators (config 5: static / adaptive refinement) ---------
 * Replace the stencils Parthenon's refinement::Restrict / Prolongate apply over the index
 * ranges of BndInfo (P:prolong_restrict/prolong_restrict.hpp):
 *   ab200_restrict    ArtemisUtils::RestrictAverage<GEOM>        src/utils/refinement/restriction.hpp:41-114
 *   ab200_prolongate  ArtemisUtils::ProlongateSharedMinMod<GEOM> src/utils/refinement/prolongation.hpp:82-184
 * A descriptor names pack entries var0 .. var0+nvar-1 of one MeshBlock's primitive (FillGhost)
 * or conserved arrays, the block's coarse buffer for those entries (device memory,
 * [nvar][cnk][cnj][cni], Parthenon's `coarse_s`), and the inclusive COARSE index box to loop
 * over.  The coarse buffer shape is Parthenon's c_cellbounds: nx/2 interior cells plus nghost
 * ghosts in every active direction (P:mesh/meshblock.cpp:205-228); ab200_coarse_shape returns
 * {cni, cnj, cnk, cis, cjs, cks}.  A whole descriptor list is ONE kernel launch. */
#define AB200_REFINE_PRIM 0
#define AB200_REFINE_CONS 1
typedef struct ab200_refine_desc {
  int fluid, block, var0, nvar, kind;
  int cis, cie, cjs, cje, cks, cke;
  double *coarse;
} ab200_refine_desc;
int ab200_coarse_shape(ab200_ctx *ctx, int *dims6);
int ab200_restrict(ab200_ctx *ctx, const ab200_refine_desc *descs, int n);
int ab200_prolongate(ab200_ctx *ctx, const ab200_refine_desc *descs, int n);

/* ---- multilevel ghost exchange: data movement (config 5) ------------------------------------
 * With ab200_restrict / ab200_prolongate above these execute Parthenon's multilevel
 * AddBoundaryExchangeTasks (P:bvals/comms/boundary_communication.cpp:406-445) for blocks on one
 * GPU: SendBoundBufs (ab200_restrict over ProResInfo::GetSend ranges, then the copies),
 * SetBounds (the copies; ab200_restrict over ProResInfo::GetSet ranges), physical BCs on the
 * coarse buffers (ab200_block_bcs), ProlongateBounds (ab200_prolongate), physical BCs on the
 * fine arrays (ab200_block_bcs).  The host supplies the index boxes from Parthenon's boundary
 * cache (BndInfo::idxer = CalcIndices, P:bvals/comms/bnd_info.cpp:105-252);
 * artemis_b200/multilevel.py restates that bookkeeping for hosts without Parthenon.
 *
 * ab200_box_desc: copy the box of extents (ni, nj, nk) with origin (ssi, ssj, ssk) of `ncomp`
 * consecutive PRIMITIVE pack entries from src_var0 of block src_block -- or, when src_coarse is
 * not NULL, from that coarse buffer ([ncomp][cnk][cnj][cni], ab200_coarse_shape) -- to origin
 * (dsi, dsj, dsk) of entries from dst_var0 of block dst_block or of the coarse buffer dst_coarse.
 * The three cases of BndInfo (bnd_info.cpp:273-304): same level fine -> fine, to a coarser
 * neighbour coarse -> fine, to a finer neighbour fine -> coarse.  A whole list is ONE launch; all
 * sources must be interior data and all destinations ghost regions (true of CalcIndices' boxes). */
typedef struct ab200_box_desc {
  int fluid, ncomp;
  int src_block, src_var0;
  const double *src_coarse; /* DEVICE or NULL */
  int dst_block, dst_var0;
  double *dst_coarse;       /* DEVICE or NULL */
  int ssi, ssj, ssk, dsi, dsj, dsk, ni, nj, nk;
} ab200_box_desc;
int ab200_box_copy(ab200_ctx *ctx, const ab200_box_desc *boxes, int n);
/* GenericBC outflow / reflect (P:bvals/boundary_conditions_generic.hpp:178-256) on face
 * (0..5 = ix1, ox1, ix2, ox2, ix3, ox3) of `ncomp` primitive pack entries from var0 of one block:
 * on its fine arrays, or on its coarse buffer when `coarse` is not NULL (the buffer slab of those
 * entries).  Applied over the full transverse extent, x1 faces of the whole list first, then
 * x2, then x3 (ApplyBoundaryConditionsOnCoarseOrFineMD). */
typedef struct ab200_block_bc_desc {
  int fluid, block, var0, ncomp, face, type; /* type: AB200_BC_OUTFLOW | AB200_BC_REFLECT, or a
                                              * user condition AB200_BC_EXTRAP | AB200_BC_INFLOW,
                                              * which takes the whole fluid (var0 = 0, ncomp =
                                              * every pack entry): the gas condition writes
                                              * species 0, the dust condition every species */
  double *coarse;                            /* DEVICE or NULL */
  /* User conditions only, optional: DEVICE table of ncomp pointers, the block's coarse array
   * of every pack entry.  Parthenon keeps ONE coarse buffer per Variable
   * (Variable::coarse_s, P:interface/variable.hpp:139), so the entries a user condition
   * couples (density, velocity, sie) are not one slab there; with this table set, `coarse`
   * is ignored and the face lives in the coarse index space.  NULL: entries contiguous from
   * `coarse` (or the fine arrays when that is NULL too). */
  double *const *coarse_entries;
} ab200_block_bc_desc;
int ab200_block_bcs(ab200_ctx *ctx, const ab200_block_bc_desc *bcs, int n);
/* StratParams::q and ::Om0 (src/pgen/strat.hpp:36-44, <rotating_frame> qshear / omega of the
 * deck) for AB200_BC_INFLOW: the background shear v2 = -q Om0 x1 imposed where it enters. */
int ab200_set_shear_bc_params(ab200_ctx *ctx, double q, double om0);

/* Flux correction (AddFluxCorrectionTasks, P:bvals/comms/boundary_communication.cpp:454-461;
 * src/artemis_driver.cpp:198-202): between ab200_calculate_fluxes and ab200_apply_update on a
 * multilevel mesh.  One descriptor per face a fine block shares with a coarser block: the box
 * [cis..cie][cjs..cje][cks..cke] of the fine block's coarse index space that covers the face
 * (CalcIndices with the flux element, prores = true), and the origin of the same cells in the
 * coarse block's own index space (CalcIndices, BoundaryExteriorRecv).  Every Metadata::Flux
 * field of the fluid (conserved fluxes; for the gas also the interface pressure) is restricted
 * with RestrictAverage<GEOM> on face elements (src/utils/refinement/restriction.hpp:41-114) and
 * written over the coarse block's flux.  The list is ONE launch.  With the diffusion operators
 * configured (ab200_configure_diffusion) and their fluxes computed, the gas descriptors also
 * correct gas.diff.momentum / gas.diff.energy -- Metadata::WithFluxes fields like the others
 * (src/gas/gas.cpp:277-285) -- in a second launch over the same list. */
typedef struct ab200_fluxcor_desc {
  int fluid, fine_block, coarse_block, dir; /* dir 0..2 */
  int cis, cie, cjs, cje, cks, cke;
  int dsi, dsj, dsk;
} ab200_fluxcor_desc;
int ab200_flux_correct(ab200_ctx *ctx, const ab200_fluxcor_desc *faces, int n);

/* ---- timestep on the device (replaces the per-cycle host round trip of
 *      P:driver/driver.cpp:210-269 + MPI_Allreduce :237) ------------------------------------ */
/* min over all bound fluids of cfl*min_dt -> device scalar new_dt (no sync) */
int ab200_estimate_timestep_device(ab200_ctx *ctx);
/* dt = min(2*dt, new_dt); clamp to tlim - time; time += old dt handled by caller flag.
 * Mirrors EvolutionDriver::SetGlobalTimeStep; runs on the device, no sync. */
int ab200_set_global_timestep_device(ab200_ctx *ctx, double tlim, int advance_time);
double *ab200_dt_device(ab200_ctx *ctx);     /* device double[4]: dt, new_dt, time, ncycle */
int ab200_read_time_state(ab200_ctx *ctx, double *host4); /* syncs */
int ab200_write_time_state(ab200_ctx *ctx, const double *host4);

/* ---- ghost zones --------------------------------------------------------------------------
 * Pack / unpack kernels of SendBoundBufs / SetBounds
 * (P:bvals/comms/boundary_communication.cpp:95-140, 273-334) driven by the caller's BndInfo. */
int ab200_halo_pack(ab200_ctx *ctx, const ab200_bnd_desc *bnd, int n);
int ab200_halo_unpack(ab200_ctx *ctx, const ab200_bnd_desc *bnd, int n);
/* Stream the two halo kernels above are launched on (NULL: the context stream).  Lets the
 * caller run pack -> transfer -> unpack of the remote neighbours concurrently with
 * ab200_fill_ghosts_local on the context stream; the caller orders the two streams with events
 * (the stage must have finished before the pack, both must have finished before
 * ab200_finish_remote_ghosts).  The two touch disjoint ghost cells. */
int ab200_set_halo_stream(ab200_ctx *ctx, void *cuda_stream);
/* Library-managed same-level exchange for a uniform nbx*nby*nbz lattice of the bound blocks
 * (block id = lx + nbx*(ly + nby*lz)).
 *  ab200_exchange_ghosts: same-GPU neighbours are filled ghost<-interior in ONE kernel (no
 *    intermediate buffer; the reference's BuffCommType::both case,
 *    P:bvals/comms/build_boundary_buffers.cpp:148-151).  Neighbours across a face flagged
 *    AB200_BC_NONE (another rank) are left to ab200_halo_pack/unpack.
 *  ab200_apply_physical_bcs: outflow / reflect (P:bvals/boundary_conditions_generic.hpp:
 *    178-256) on lattice faces flagged OUTFLOW/REFLECT, in the reference's face order
 *    ix1, ox1, ix2, ox2, ix3, ox3; call it after every neighbour (local and remote) is set. */
int ab200_set_topology(ab200_ctx *ctx, int nbx, int nby, int nbz, const int bc[6]);
int ab200_exchange_ghosts(ab200_ctx *ctx);
int ab200_apply_physical_bcs(ab200_ctx *ctx);
/* Single-rank fast path: ab200_exchange_ghosts + ab200_apply_physical_bcs +
 * ab200_prim_to_cons_ghosts in ONE kernel over the ghost cells only (tasks
 * src/artemis_driver.cpp:258-261 for a partition with no remote neighbour).  Fails with
 * AB200_ESTATE when a lattice face is flagged AB200_BC_NONE. */
int ab200_fill_ghosts(ab200_ctx *ctx);
/* Multi-rank split of the same fill (one process per GPU, block-spatial partition):
 *   ab200_fill_ghosts_local      every ghost cell whose neighbour chain stays on this GPU
 *                                (same-GPU neighbours, physical boundaries) + its PrimToCons;
 *                                cells that depend on a face flagged AB200_BC_NONE are skipped.
 *   [caller: ab200_halo_pack -> NCCL send/recv -> ab200_halo_unpack, one sweep per direction]
 *   ab200_finish_remote_ghosts   the skipped cells: directions that do not cross onto another
 *                                rank are resolved (neighbour shift / outflow / reflect /
 *                                periodic) against the just-unpacked ghost data, then
 *                                PrimToCons.  Together: src/artemis_driver.cpp:258-261. */
int ab200_fill_ghosts_local(ab200_ctx *ctx);
int ab200_finish_remote_ghosts(ab200_ctx *ctx);

/* ---- host-buffer entry point (a CPU-resident Parthenon build, and bench.py's e2e leg) ----
 * Runs `ncycles` full rk/vl cycles on state held in HOST memory: uploads the gas (and dust)
 * primitives [nblocks][nvar][nk][nj][ni], rebuilds conserved state with PrimToCons, advances
 * (fused path + library exchange + device dt), downloads primitives and conserved u0.
 * integrator: 0 rk1, 1 rk2, 2 vl2, 3 rk3.  dt_io: in = current dt (<=0: estimate), out = next.
 * Bytes that need not cross PCIe do not: the gas PRESSURE entries of the input are ignored
 * (PrimToCons recomputes P = EOS(rho, sie) over the entire domain, fill_derived.cpp:247) and
 * are not uploaded; *_cons_host may be NULL (cons is a pure function of prim), which halves
 * the download and skips the ghost PrimToCons.
 * Multi-rank (faces flagged AB200_BC_NONE, ab200_comm_set_layout done): every rank calls it
 * collectively on its own partition; the cycles run through ab200_run_cycles_mr and an estimated
 * first dt is all-reduced. */
int ab200_cycles_host(ab200_ctx *ctx, int integrator, int ncycles, double *dt_io,
                      double *gas_prim_host, double *gas_cons_host, double *dust_prim_host,
                      double *dust_cons_host);
/* Which zones of the host arrays cross PCIe in ab200_cycles_host (default 0: whole arrays).
 * Ghost zones are redundant at a cycle boundary -- the reference refills them from interior
 * zones at the end of every stage (src/artemis_driver.cpp:258-261) -- and they are 30 % of a
 * 64^3 MeshBlock with nghost = 4:
 *   AB200_HOST_INTERIOR_IN   upload interior zones only and rebuild the ghost zones on the device
 *                            with the library's exchange + physical boundaries before
 *                            PrimToCons.  Identical results whenever the caller's ghost zones
 *                            were consistent (any state a stage driver or Mesh::Initialize
 *                            left behind); the caller's ghost values are never read.
 *   AB200_HOST_INTERIOR_OUT  download interior zones only; the ghost zones of the host arrays
 *                            keep their old contents (the next call with _IN does not read
 *                            them; a host consumer that needs them refills them itself).
 *   AB200_HOST_ZERO_COPY_IN / _OUT   the interior rows of that direction are moved by copy
 *                            kernels that read / write the PINNED host arrays in place over PCIe
 *                            (cudaHostAlloc / cudaHostRegister memory; AB200_EINVAL otherwise)
 *                            instead of strided DMA (cudaMemcpy3DAsync, which also accepts
 *                            pageable memory).  Measured on B200 / PCIe 5 (256^3 zones in 64^3
 *                            MeshBlocks, profiles/r02_e2e_transfer_ab.json): strided DMA is the
 *                            better engine on the way out, the copy kernel by far on the way in
 *                            (host-to-device DMA of 512-byte rows runs at a fraction of the link
 *                            rate), so AB200_HOST_INTERIOR_IN | _OUT | AB200_HOST_ZERO_COPY_IN is
 *                            the combination bench.py's end-to-end leg uses. */
#define AB200_HOST_INTERIOR_IN 1
#define AB200_HOST_INTERIOR_OUT 2
#define AB200_HOST_ZERO_COPY_IN 4
#define AB200_HOST_ZERO_COPY_OUT 8
#define AB200_HOST_ZERO_COPY (AB200_HOST_ZERO_COPY_IN | AB200_HOST_ZERO_COPY_OUT)
int ab200_set_host_transfer(ab200_ctx *ctx, int flags);

/* Device-resident driver loop: `ncycles` cycles of {per stage: ab200_fused_stage ->
 * ab200_exchange_ghosts -> ab200_apply_physical_bcs -> ab200_prim_to_cons_ghosts} followed by
 * ab200_estimate_timestep_device + ab200_set_global_timestep_device, i.e.
 * ArtemisDriver::Step (src/artemis_driver.cpp:101-121) for a single-rank uniform mesh with
 * no host round trip.  State must already be bound; dt/time live in ab200_dt_device(). */
int ab200_run_cycles(ab200_ctx *ctx, int integrator, int ncycles, double tlim);
/* Opt-in: ab200_run_cycles replays a CUDA graph of one cycle when it can (no finite tlim, at
 * least three cycles, a launch sequence that is periodic over one cycle): the first cycle runs
 * eagerly, the second is captured (on a private stream when the context runs on the legacy
 * default stream, which cannot be captured), the rest are one cudaGraphLaunch each on the
 * context's stream.  Results are identical to the eager loop, bit for bit (tested).  Measured
 * on B200 (profiles/r02b_graph_ab.log): config 1's 8192-zone mesh 137 -> 127 us per cycle, a
 * 64^3 mesh 254 -> 239 us; nothing at BASELINE's 256^3, where a cycle is 4 ms of kernels.  The
 * eager loop never waits for the host either (every input of a cycle lives on the device), so
 * the gain is only the inter-kernel gap.  Off by default; AB200_GRAPH=1 in the environment
 * turns it on for every context. */
int ab200_set_graph_replay(ab200_ctx *ctx, int on);
/* cycles this context has executed as a graph launch so far */
int ab200_graph_replay_count(ab200_ctx *ctx, long long *count);

/* ---- multi-rank transport (one process per GPU; NCCL over NVLink / NVSwitch) ----------------
 * Replaces, for blocks whose neighbour lives on another rank, the MPI path of
 * parthenon::SendBoundBufs / ReceiveBoundBufs / SetBounds
 * (P:bvals/comms/boundary_communication.cpp:48-334, CommBuffer::Send / TryReceive
 * P:utils/communication_buffer.hpp:209-420: one MPI_Isend per (block, neighbour, variable), polled
 * with MPI_Iprobe / MPI_Test) and the MPI_Allreduce(MIN) of EvolutionDriver::SetGlobalTimeStep
 * (P:driver/driver.cpp:237).  NCCL is bound at run time (dlopen): single-rank hosts never load it.
 *
 *   rank 0: ab200_comm_unique_id(id)  -> host broadcasts the 128 bytes (MPI_Bcast in Parthenon)
 *   every rank: ab200_comm_init(ctx, nranks, rank, id)            ncclCommInitRank on ctx's device
 *               ab200_comm_set_layout(ctx, lx, ly, lz, periodic)  rank lattice of the block-spatial
 *                   partition (rank = x + lx*(y + ly*z)); plans the SINGLE-ROUND exchange -- one
 *                   aggregated message per peer rank (faces, rank edges and corner at once) --
 *                   for the bound fluids.  Faces towards other ranks carry AB200_BC_NONE in
 *                   ab200_set_topology.
 *   per stage:  ab200_fused_stage -> ab200_comm_exchange_begin (pack / grouped ncclSend+ncclRecv /
 *               unpack on a library-owned stream, ordered after the stage by an event)
 *               -> ab200_fill_ghosts_local (concurrently) -> ab200_comm_exchange_end
 *               -> ab200_finish_remote_ghosts
 *   per cycle:  ab200_allreduce_min(ctx, ab200_dt_device(ctx) + 1) -> ab200_set_global_timestep_device
 * ab200_run_cycles_mr is that loop (the multi-rank twin of ab200_run_cycles). */
int ab200_comm_unique_id(char *id128);
int ab200_comm_init(ab200_ctx *ctx, int nranks, int rank, const char *id128);
int ab200_comm_destroy(ab200_ctx *ctx);
int ab200_comm_set_layout(ab200_ctx *ctx, int layx, int layy, int layz, const int *periodic3);
long long ab200_comm_bytes_per_exchange(ab200_ctx *ctx);
/* 1 if the exchange runs over CUDA IPC peer mappings (the pack kernel stores straight into the
 * peers' receive slabs over NVLink, one flag per peer; NCCL only bootstraps the handles and does
 * the dt all-reduce), 0 if it goes through grouped ncclSend / ncclRecv (AB200_NO_DIRECT=1 or IPC
 * unavailable on some rank). */
int ab200_comm_is_direct(ab200_ctx *ctx);
int ab200_comm_exchange_begin(ab200_ctx *ctx);
int ab200_comm_exchange_end(ab200_ctx *ctx);
/* in-place MIN all-reduce of one DEVICE double on the context's stream; identity without a
 * communicator */
int ab200_allreduce_min(ab200_ctx *ctx, double *dev_scalar);
int ab200_run_cycles_mr(ab200_ctx *ctx, int integrator, int ncycles, double tlim);
/* The exchange planner on its own (no context, no GPU): rows of 13 values (peer, is_recv, fluid,
 * block, var0, ncomp, si, ei, sj, ej, sk, ek, offset-in-message) for the rank at lattice position
 * rl3 of lay3; index ranges are the same-level branch of CalcIndices
 * (P:bvals/comms/bnd_info.cpp:152-213).  *rows_out is malloc'd: free with ab200_comm_plan_free. */
int ab200_comm_plan_direct(const int *nblk3, const int *nt3, const int *s3, const int *e3,
                           const int *ng3, int nfluids, const int *fluid_type,
                           const int *nspecies, const int *lay3, const int *rl3,
                           const int *periodic3, long long **rows_out, int *nrows_out);
void ab200_comm_plan_free(long long *rows);

/* ---- utilities ---------------------------------------------------------------------------- */
int ab200_malloc(ab200_ctx *ctx, void **dptr, size_t bytes);
int ab200_free(ab200_ctx *ctx, void *dptr);
int ab200_memcpy_h2d(ab200_ctx *ctx, void *dst, const void *src, size_t bytes);
int ab200_memcpy_d2h(ab200_ctx *ctx, void *dst, const void *src, size_t bytes);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long long ab200_launch_count(ab200_ctx *ctx);
/* CUDA-event timing of the work enqueued between begin/end on the context's stream (ms) */
int ab200_timer_begin(ab200_ctx *ctx);
int ab200_timer_end(ab200_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* AB200_H_ */
