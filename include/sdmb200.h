/*
 * sdmb200.h -- C ABI of libsdmb200.so, the B200-native (sm_100a) implementation of the
 * Single-Decoupling / Alchemical-Transfer dual-state force path of
 * rajatkrpal/openmm_sdm_plugin.
 *
 * Plain pointers and sizes only; no C++/torch/OpenMM types cross this boundary.  Every
 * function returns SDM_OK (0) or a negative sdm_status; the message of the last failure on
 * the calling thread is sdm_last_error().  Nothing throws across the boundary.  There is NO
 * CPU fallback: without a CUDA device sdm_create() fails with SDM_ERR_NO_DEVICE.
 *
 * Two groups of entry points (citations are relative to the reference tree):
 *
 *  (A) the fused path -- replaces the whole force column of LangevinIntegratorSDM::step
 *      (openmmapi/src/LangevinIntegratorSDM.cpp:156-182): the two
 *      context->calcForcesAndEnergy(true,true,4) calls (:160,:168), SaveState1 / MakeState2 /
 *      SaveState2 / RestoreState1 (:162-173, platforms/reference/src/ReferenceSDMKernels.cpp:
 *      161-199) and the pre-integration half of execute() (:180,
 *      ReferenceSDMKernels.cpp:202-318: soft-core, bias, PotEnergy/BindE, non-equilibrium
 *      work, hybrid force).
 *
 *  (B) the literal kernel-interface operations -- one entry point per virtual of
 *      SDMPlugin::IntegrateLangevinStepSDMKernel (openmmapi/include/SDMKernels.h:60-111) on
 *      caller-owned DEVICE float4 buffers, for an OpenMM platform adapter that keeps using
 *      OpenMM's own NonbondedForce (platforms/opencl/src/kernels/langevin.cl:72-87,147-206).
 */
#ifndef SDMB200_H_
#define SDMB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDM_ABI_VERSION 2

typedef enum {
    SDM_OK = 0,
    SDM_ERR_INVALID = -1,      /* bad argument (null pointer, index out of range, ...)      */
    SDM_ERR_NO_DEVICE = -2,    /* no CUDA device / driver: there is no CPU fallback         */
    SDM_ERR_CUDA = -3,         /* a CUDA runtime call failed; see sdm_last_error()          */
    SDM_ERR_BOX = -4,          /* periodic box smaller than 2*cutoff (OpenMM throws there)  */
    SDM_ERR_SOFTCORE = -5,     /* "Unknown soft core method" (LangevinIntegratorSDM.cpp:147)*/
    SDM_ERR_STALE_LIST = -6,   /* an atom moved more than skin/2 since the list was built: reported in
                                  sdm_scalars.status; the list is rebuilt at the next sdm_eval(),
                                  repeat the evaluation.  The status is STICKY: it stays in
                                  sdm_scalars.status of later evaluations until the host has read it
                                  (sdm_get_scalars / sdm_read_results / sdm_enqueue_results), so a
                                  pipelined sequence of sdm_eval calls cannot lose it              */
    SDM_ERR_CAPACITY = -7,     /* internal capacity exceeded -- the per-hit scratch of the displaced-atom
                                  kernels, or a pair-list build (enqueued without host synchronisation,
                                  sized with bounds from the previous build) that outgrew its bounds:
                                  reported in sdm_scalars.status of every affected replica (sticky like
                                  STALE_LIST); the context grows what was too small / sizes the next build
                                  on the host; repeat the evaluation                              */
    SDM_ERR_CONSTRAINT = -8    /* a constraint cluster did not converge (sdm_md_step)              */
} sdm_status;

/* NonbondedForce::NonbondedMethod as the reference's reader sets it
 * (example/desmonddmsfile75.py:418-426). */
#define SDM_NOCUTOFF 0
#define SDM_CUTOFF_NONPERIODIC 1
#define SDM_CUTOFF_PERIODIC 2
/* NonbondedForce::Ewald / ::PME (what example/test_explicit.py:64 asks for).  A context created with one of them
 * computes the DIRECT-SPACE part of the Ewald sum -- erfc(alpha r)/r pair terms inside the cutoff, the
 * erf(alpha r)/r correction of the excluded pairs, 1-4 exceptions and the dispersion correction: OpenMM's
 * calculateEwaldIxn with includeDirect.  The reciprocal-space part (and the self energy OpenMM books with it) is
 * added by sdm_enable_reciprocal_pme() (smooth PME on the device, both states) or handed in through
 * sdm_set_external_dual() -- with OpenMM in the loop: NonbondedForce::setReciprocalSpaceForceGroup and one
 * evaluation of that group per state.  Both values select the same arithmetic. */
#define SDM_EWALD 3
#define SDM_PME 4

/* Lennard-Jones combining rule (sdm_system.lj_combining).  Lorentz-Berthelot is NonbondedForce's; the geometric
 * rule is what createSystem(OPLS=True) of the reference's reader builds in the nonbonded force group: NonbondedForce
 * with every epsilon zeroed plus CustomNonbondedForce("4 eps12 ((s12/r)^12 - (s12/r)^6); s12 = sqrt(s1 s2);
 * eps12 = sqrt(eps1 eps2)") with the same exclusions and cutoff, no long-range correction
 * (example/desmonddmsfile75.py:780-810, :427-438).  1-4 exceptions keep their own sigma / epsilon either way. */
#define SDM_LJ_LORENTZ_BERTHELOT 0
#define SDM_LJ_GEOMETRIC 1

/* LangevinIntegratorSDM.h:120-122 and :143-145 */
#define SDM_BIAS_LINEAR 0
#define SDM_BIAS_QUADRATIC 1
#define SDM_BIAS_ILOGISTIC 2
#define SDM_SOFTCORE_NONE 0
#define SDM_SOFTCORE_TANH 1
#define SDM_SOFTCORE_RATIONAL 2

/* Which force array sdm_get_forces() returns. */
#define SDM_FORCE_HYBRID 0     /* F  = F1 + sp*(F2-F1) + Fb  (ReferenceSDMKernels.cpp:309-318) */
#define SDM_FORCE_STATE1 1     /* F1 = State1Forces                                            */
#define SDM_FORCE_STATE2 2     /* F2 = State2Forces  (= F1 + dF)                               */
#define SDM_FORCE_DELTA 3      /* dF = F2 - F1, accumulated over moved pairs only              */

/* Pair-kernel selection (sdm_options.pair_mode). */
#define SDM_PAIR_AUTO 0
#define SDM_PAIR_ALLPAIRS 1    /* O(N^2) tile kernel, no list (small systems, cross-check)     */
#define SDM_PAIR_CLUSTER 2     /* cell-sorted 8-atom clusters + cluster-pair list              */

/* What reaches NonbondedForce in force group 2 (example/desmonddmsfile75.py:772-850) plus the
 * displacement map (LangevinIntegratorSDM.h:467-472,508).  The map is SNAPSHOTTED here, like
 * the reference does at initialize() (ReferenceSDMKernels.cpp:150-154); use
 * sdm_set_displacement() to change it later (the reference needs Context::reinitialize). */
typedef struct {
    int32_t n_atoms;
    int32_t method;                 /* SDM_NOCUTOFF / SDM_CUTOFF_NONPERIODIC / SDM_CUTOFF_PERIODIC / SDM_EWALD / SDM_PME */
    double cutoff;                  /* nm */
    double eps_rf;                  /* reaction-field dielectric, 78.3 is NonbondedForce's default */
    double box[3];                  /* orthorhombic box edges (nm); ignored unless periodic */
    int32_t use_dispersion_correction;
    int32_t n_exclusions;           /* every addException pair, zero-parameter or not */
    int32_t n_exceptions;           /* exceptions with chargeProd != 0 or epsilon != 0 (1-4) */
    int32_t n_replicas;             /* lambda-replicas resident in this context (>= 1) */
    const double* charge;           /* [n_atoms] e */
    const double* sigma;            /* [n_atoms] nm */
    const double* epsilon;          /* [n_atoms] kJ/mol */
    const int32_t* exclusions;      /* [2*n_exclusions] System particle indices */
    const int32_t* exceptions;      /* [2*n_exceptions] */
    const double* exception_params; /* [3*n_exceptions] chargeProd (e^2), sigma (nm), epsilon (kJ/mol) */
    const double* displacement;     /* [3*n_atoms] nm; NULL = all zero */
    double ewald_alpha;             /* SDM_EWALD / SDM_PME: splitting parameter (1/nm); 0 = OpenMM's rule
                                       sqrt(-log(2 tol)) / cutoff (NonbondedForceImpl::calcPMEParameters) */
    double ewald_tolerance;         /* tol of that rule; 0 = NonbondedForce's default 5e-4 */
    int32_t lj_combining;           /* SDM_LJ_LORENTZ_BERTHELOT (0, default) / SDM_LJ_GEOMETRIC; the geometric rule
                                       needs use_dispersion_correction = 0 with a periodic method */
    int32_t reserved_;
} sdm_system;

typedef struct {
    int32_t device;                 /* CUDA device ordinal; -1 = current device */
    int32_t pair_mode;              /* SDM_PAIR_* */
    double skin;                    /* list buffer rlist - cutoff (nm); <0 = default 0.06 */
    int32_t nstlist;                /* rebuild the cluster-pair list every nstlist evals; <=0 = default 20 */
    int32_t exact_cutoff;           /* 1: re-test pairs within 1 ulp-band of the cutoff in FP64 so
                                       the in-cutoff pair set is bit-identical to a double-precision
                                       evaluation (default 1) */
    int32_t use_graph;              /* 1: replay the per-eval kernel sequence of the cluster path as a
                                       CUDA graph between list rebuilds (default 1) */
    int32_t reserved[7];
} sdm_options;

/* Scalar state of LangevinIntegratorSDM that execute() reads
 * (ReferenceSDMKernels.cpp:205-258); field meaning and defaults as in the integrator's ctor
 * (LangevinIntegratorSDM.cpp:48-85).  Units: kJ/mol, alpha in (kJ/mol)^-1, ps. */
typedef struct {
    int32_t bias_method;            /* SDM_BIAS_* */
    int32_t softcore_method;        /* SDM_SOFTCORE_* */
    double lambdac, gammac, wbcoeff, w0coeff;
    double lambda1, lambda2, alpha, u0;
    double umax, acore, ubcore;
    int32_t nonequilibrium;         /* getNonEquilibrium() == 1 enables the lambda schedule */
    int32_t pad_;
    double noneq_tmax, work_value, time, step_size;
    double m_lambda1, m_lambda2, m_u0, m_w0;   /* slopes     (set*Slope)     */
    double b_lambda1, b_lambda2, b_u0, b_w0;   /* intercepts (set*intercept) */
} sdm_alch;

/* Everything execute() derives before it integrates. */
typedef struct {
    double E1, E2, Eb;              /* State1Energy, State2Energy (= E1 + u), RestraintEnergy */
    double u;                       /* E2 - E1, accumulated over moved pairs only in FP64 */
    double u_sc, fp;                /* SoftCoreF(u) and its derivative */
    double ebias, bfp;              /* bias energy W(u_sc) and slope */
    double sp;                      /* bfp*fp, the force-mixing coefficient */
    double pot_energy;              /* E1 + ebias + Eb   -> integrator.setPotEnergy */
    double bind_e;                  /* u_sc              -> integrator.setBindE     */
    double E1_pair, E1_exc, E1_disp;/* decomposition of E1: pair sum, 1-4 exceptions, dispersion */
    int64_t n_pairs1;               /* in-cutoff non-excluded pairs at state 1 */
    int64_t n_moved1, n_moved2;     /* in-cutoff moved pairs (>=1 displaced atom, different
                                       displacement) at state 1 / state 2 */
    int32_t status;                 /* SDM_OK, SDM_ERR_SOFTCORE, SDM_ERR_STALE_LIST, ... */
    int32_t list_age;               /* evals since the cluster-pair list was built */
} sdm_scalars;

typedef struct sdm_ctx sdm_ctx;

/* ---- library ---------------------------------------------------------------------------- */
int sdm_abi_version(void);
const char* sdm_last_error(void);
int sdm_device_count(void);          /* 0 when there is no usable CUDA device */
void sdm_default_options(sdm_options* opt);
void sdm_default_alch(sdm_alch* alch);   /* LangevinIntegratorSDM ctor defaults */

/* ---- (A) fused dual-state path ------------------------------------------------------------ */
/* Replaces kernel.initialize(system, integrator) (ReferenceSDMKernels.cpp:144-159): copies all
 * inputs (nothing is retained), uploads topology + displacement map, allocates per-replica
 * state.  The handle is owned by the caller and freed with sdm_destroy(). */
int sdm_create(const sdm_system* sys, const sdm_options* opt, sdm_ctx** out);
void sdm_destroy(sdm_ctx* ctx);

/* All work of a ctx is ordered on one CUDA stream (default: a stream the ctx owns).
 * `cuda_stream` is a cudaStream_t / CUstream passed as void* (NULL = legacy default stream). */
int sdm_set_stream(sdm_ctx* ctx, void* cuda_stream);
int sdm_synchronize(sdm_ctx* ctx);

/* Pinned host memory helpers for the end-to-end path (plain malloc'ed buffers also work,
 * they are just slower to copy). */
int sdm_host_alloc(void** ptr, uint64_t bytes);
int sdm_host_free(void* ptr);

/* Positions of one replica, [3*n_atoms] doubles in System particle order (Vec3 layout), nm.
 * Host version copies asynchronously on the ctx stream; device version takes a device pointer. */
int sdm_set_positions(sdm_ctx* ctx, int replica, const double* xyz);
int sdm_set_positions_device(sdm_ctx* ctx, int replica, const double* d_xyz);
/* Device address of the ctx-owned position buffer of a replica (for in-place integrators). */
int sdm_positions_device_ptr(sdm_ctx* ctx, int replica, double** d_xyz);
/* All replicas at once: xyz_all is [n_replicas][3*n_atoms] host doubles, one asynchronous copy. */
int sdm_set_positions_all(sdm_ctx* ctx, const double* xyz_all);
/* The same in single precision -- the precision positions and forces have on the reference's live
 * OpenCL path (posq / force are float4 there, platforms/opencl/src/OpenCLSDMKernels.cpp:96-103):
 * half the bytes over PCIe.  The copy lands in a device staging buffer and is widened to the ctx's
 * FP64 positions by a kernel on the ctx stream. */
int sdm_set_positions_all_f32(sdm_ctx* ctx, const float* xyz_all);

/* Result of the group-1 (bonded/restraint) evaluation the integrator does at :176: forces left
 * in the force buffer and RestraintEnergy.  fb may be NULL (zero).  Host pointers. */
int sdm_set_bonded_forces(sdm_ctx* ctx, int replica, const double* fb, double eb);

int sdm_set_alchemical(sdm_ctx* ctx, int replica, const sdm_alch* alch);
/* Reads back the state execute() writes into the integrator in non-equilibrium mode
 * (lambda, lambda1, lambda2, u0, w0, work value, time).  Synchronises. */
int sdm_get_alchemical(sdm_ctx* ctx, int replica, sdm_alch* alch);

/* Replace the displacement map ([3*n_atoms], nm) -- what Context::reinitialize() would do. */
int sdm_set_displacement(sdm_ctx* ctx, const double* displacement);

/* One dual-state evaluation for every resident replica: E1, u, u_sc, W, sp, PotEnergy and the
 * hybrid force, enqueued on the ctx stream; returns without synchronising. */
int sdm_eval(sdm_ctx* ctx);
/* Force a rebuild of the cluster-pair list at the next sdm_eval(). */
int sdm_invalidate_list(sdm_ctx* ctx);

int sdm_get_scalars(sdm_ctx* ctx, int replica, sdm_scalars* out);       /* synchronises */
int sdm_get_forces(sdm_ctx* ctx, int replica, int which, double* out);  /* [3*n_atoms], synchronises */
int sdm_forces_device_ptr(sdm_ctx* ctx, int replica, double** d_f);     /* hybrid force, device */
/* All replicas at once: hybrid forces into forces_all ([n_replicas][3*n_atoms], may be NULL) and
 * scalars into scalars_all ([n_replicas], may be NULL); two copies, one synchronisation. */
int sdm_read_results(sdm_ctx* ctx, double* forces_all, sdm_scalars* scalars_all);
/* The same in two halves, for callers that pipeline several contexts (replica groups) on
 * separate streams so that the copies of one group overlap the kernels of another:
 * sdm_enqueue_results() queues the device->host copies of the hybrid forces (into forces_all,
 * ideally pinned; may be NULL) and of the scalar blocks behind the last sdm_eval() and returns
 * at once; after sdm_synchronize(), sdm_collect_scalars() hands out the scalars that arrived. */
int sdm_enqueue_results(sdm_ctx* ctx, double* forces_all);
/* The same with the hybrid forces narrowed to single precision on the device first
 * ([n_replicas][3*n_atoms] floats, 12 B/atom instead of 24); the scalars stay FP64. */
int sdm_enqueue_results_f32(sdm_ctx* ctx, float* forces_all);
int sdm_collect_scalars(sdm_ctx* ctx, sdm_scalars* scalars_all);
/* Debug / parity: the sorted in-cutoff non-excluded (i<j) pair list the pair kernel evaluated at
 * state 1, System particle indices.  pairs may be NULL to query *n only.  Synchronises. */
int sdm_get_pairs(sdm_ctx* ctx, int replica, int32_t* pairs, int64_t max_pairs, int64_t* n);

/* Launch statistics since creation (own kernels only) and device time of the last eval. */
int sdm_get_launch_count(sdm_ctx* ctx, int64_t* n_launches);
/* Device time (ms) spent in the pair kernel during the last sdm_eval() that had timing
 * enabled; enable with sdm_set_timing(ctx, 1) (adds two cudaEventRecord per eval; the displaced-atom kernels then
 * follow the pair kernel instead of running beside it).  The kernel is launched as in production -- for large
 * batches with room left on every SM for the displaced-atom kernels; sdm_set_timing(ctx, 2) times it with every
 * resident block it can have instead (a diagnostic: that launch is not what an evaluation runs). */
int sdm_set_timing(sdm_ctx* ctx, int enabled);
int sdm_get_last_timing(sdm_ctx* ctx, float* pair_ms, float* total_ms);
/* Algorithmic work of the last eval, summed over replicas: in-cutoff pairs evaluated. */
int sdm_get_info(sdm_ctx* ctx, const char* key, double* value);

/* ---- (B) literal kernel-interface operations on DEVICE float4 buffers --------------------- */
/* MakeState2: posq[i] += displ[i]                   (langevin.cl:198-206) */
int sdm_k_make_state2(void* cuda_stream, int n, void* posq, const void* displ);
/* SaveState1: save_f = force, save_x = posq         (langevin.cl:161-178) */
int sdm_k_save_state1(void* cuda_stream, int n, const void* posq, const void* force,
                      void* save_f, void* save_x);
/* SaveState2: save_f = force                        (langevin.cl:183-192) */
int sdm_k_save_state2(void* cuda_stream, int n, const void* force, void* save_f);
/* RestoreState1: posq = saved                       (langevin.cl:147-155) */
int sdm_k_restore_state1(void* cuda_stream, int n, void* posq, const void* saved);
/* sdmForce: force = (1-sp)*f1 + sp*f2 + force, all four lanes, sp as float
 * (langevin.cl:72-87, OpenCLSDMKernels.cpp:273-275) */
int sdm_k_hybrid_force(void* cuda_stream, int n, const void* f1, const void* f2, void* force,
                       float sp);
/* The two integration kernels the plugin owns (langevin.cl:7-31 and :37-69, single-precision
 * form; constraints are applied by OpenMM between them, OpenCLSDMKernels.cpp:357-372).
 * velm = (vx, vy, vz, 1/mass); atoms with 1/mass == 0 are left untouched.  `random` holds
 * normally distributed numbers, one float4 per atom starting at random_index.
 *   part 1: v = vscale*v + fscale*(1/m)*F + noisescale*sqrt(1/m)*xi ; posDelta = stepSize*v
 *   part 2: posq.xyz += posDelta.xyz ; v.xyz = posDelta.xyz/stepSize */
int sdm_k_langevin_part1(void* cuda_stream, int n, void* velm, const void* force, void* pos_delta,
                         float vscale, float fscale, float noisescale, float step_size,
                         const void* random, uint32_t random_index);
int sdm_k_langevin_part2(void* cuda_stream, int n, void* posq, const void* pos_delta, void* velm,
                         float step_size);
/* vscale = exp(-dt*friction), fscale = (1-vscale)/friction (dt when friction == 0),
 * noisescale = sqrt(kT*(1-vscale^2)), kT = BOLTZ*temperature (OpenCLSDMKernels.cpp:331-336,
 * BOLTZ at :57-60).  Units: K, 1/ps, ps. */
int sdm_langevin_params(double temperature, double friction, double step_size, double* vscale,
                        double* fscale, double* noisescale);
/* ---- device-resident Langevin dynamics -------------------------------------
 * The update the reference applies after the hybrid force, Reference-platform algorithm
 * (platforms/reference/src/ReferenceStochasticDynamicsSDM.cpp:131-266; the OpenCL kernels
 * langevin.cl:7-69 are its float form): v = vscale*v + fscale*F/m + noisescale*xi/sqrt(m),
 * x' = x + dt*v, v = (x'-x)/dt.  Positions and velocities of all replicas stay in HBM, FP64, with
 * the reference's operation order, so a step needs no host<->device copy.  Distance constraints
 * are applied where the reference applies them (between x' and the velocity/position copy,
 * ReferenceStochasticDynamicsSDM.cpp:250-262) once sdm_md_set_constraints() has been called. */
/* masses [n_atoms] (0 = particle does not move); velocities start at zero.  friction (1/ps) must be
 * > 0 (the reference divides by it); seed keys the Philox4x32-10 noise stream. */
int sdm_md_init(sdm_ctx* ctx, const double* masses, double temperature, double friction,
                double step_size, uint64_t seed);
int sdm_md_set_velocities(sdm_ctx* ctx, int replica, const double* v);   /* [3*n_atoms] nm/ps */
int sdm_md_get_velocities(sdm_ctx* ctx, int replica, double* v);         /* synchronises */
int sdm_get_positions(sdm_ctx* ctx, int replica, double* xyz);           /* current positions, synchronises */
/* The System's distance constraints (System::addConstraint as the reference's reader fills it,
 * example/desmonddmsfile75.py:560-561,589-595,603-637): pairs [2*n], distances [n] (nm), relative
 * tolerance (Integrator::getConstraintTolerance, LangevinIntegratorSDM.cpp:52: 1e-5; <= 0 = 1e-5).
 * Rigid three-site molecules (three mutual constraints, two equal legs to atoms of equal mass: the
 * constraint_hoh waters) are solved analytically (SETTLE), every other cluster of coupled
 * constraints (a heavy atom and its hydrogens; at most 8 atoms / 16 constraints) by an in-thread
 * SHAKE iteration to the tolerance.  Call after sdm_md_init(); n = 0 removes them. */
int sdm_md_set_constraints(sdm_ctx* ctx, int32_t n_constraints, const int32_t* pairs,
                           const double* distances, double tolerance);
/* nsteps x (sdm_eval + Langevin update [+ constraints]) on the device; returns when they are done.
 * A step whose evaluation reports SDM_ERR_STALE_LIST / SDM_ERR_CAPACITY is NOT taken (nor are the
 * steps enqueued behind it): the list is rebuilt / the scratch grown and the missing steps are
 * repeated, so the trajectory never integrates a force that misses pairs.  Fails with
 * SDM_ERR_STALE_LIST if a freshly built list goes stale within a single step (skin too small for
 * the step size), with SDM_ERR_CONSTRAINT if a cluster does not converge. */
int sdm_md_step(sdm_ctx* ctx, int nsteps);
/* Steps taken since sdm_md_init() and how many of the enqueued steps had to be repeated because of
 * a stale list. */
int sdm_md_get_counters(sdm_ctx* ctx, uint64_t* steps_taken, uint64_t* steps_repeated);
/* The update alone, with the hybrid force already on the device (forces_all == NULL) or with
 * forces_all [R][n][3] uploaded first (test hook: the reference's forces in, its x and v out). */
int sdm_md_update(sdm_ctx* ctx, const double* forces_all);
/* Test hook: the NEXT update draws its normals from xi_all [R][n][3] (atom-major, x y z: the order
 * the reference consumes SimTKOpenMMUtilities::getNormallyDistributedRandomNumber) instead of the
 * Philox stream; NULL cancels. */
int sdm_md_set_noise(sdm_ctx* ctx, const double* xi_all);
/* 0.5 * sum m v^2 of one replica (ReferenceSDMKernels.cpp:105-137 without constraints). */
int sdm_md_kinetic_energy(sdm_ctx* ctx, int replica, double* ke);

/* Contributions of BOTH states that are computed outside this library, per replica: energies E1_ext, E2_ext
 * (kJ/mol) and forces F1_ext, F2_ext ([3*n_atoms], kJ/mol/nm, System order) of a term that cannot be reduced
 * to moved pairs -- the reciprocal-space part of PME (global in the charge density), a GB model.  They enter
 * like the nonbonded terms themselves: E1 += E1_ext, u += E2_ext - E1_ext, F1 += F1_ext, F2 - F1 += F2_ext -
 * F1_ext, and stay in force until replaced; NULL force pointers remove them.  (Row N4 of SURVEY.md 8f.) */
int sdm_set_external_dual(sdm_ctx* ctx, int replica, const double* f1_ext, const double* f2_ext,
                          double e1_ext, double e2_ext);

/* The reciprocal-space part of PME on the device (SDM_PME / SDM_EWALD contexts): smooth particle-mesh Ewald with
 * B-splines of order 5 (OpenMM 7.3 ReferencePME), both states of every replica per evaluation, plus the self
 * energy -- it then fills the external slots itself (sdm_set_external_dual is refused).  grid: [3] mesh sizes,
 * or NULL for OpenMM's rule ceil(2 alpha L / (3 tol^(1/5))) rounded up to a size with factors 2, 3, 5, 7
 * (sdm_get_info "pme_grid_x/y/z").  Needs cuFFT (libcufft.so.11) at run time; SDM_ERR_CUDA if it is absent. */
int sdm_enable_reciprocal_pme(sdm_ctx* ctx, const int32_t* grid);

/* The HCT generalized-Born model with the ACE surface-area term on the device: what the reference's DMS reader adds
 * for implicitSolvent=HCT (example/desmonddmsfile75.py:454-465: GBSAHCTForce(SA='ACE') in the nonbonded force group,
 * so LangevinIntegratorSDM::step, openmmapi/src/LangevinIntegratorSDM.cpp:160-170, evaluates it in both states).
 * Expressions of OpenMM 7.3's app/internal/customgbforces.py under CustomGBForce's NoCutoff / no-exclusion rules;
 * Born radii are global in the coordinates, so both states of every replica get the full model per evaluation
 * (FP64), and the library fills the external slots itself (sdm_set_external_dual is refused).  Arrays of n_atoms:
 * charge (NULL: the NonbondedForce charges of sdm_system), offset_radius = "or" (nm, radius - 0.009),
 * scaled_radius = "sr" (nm, scale * or) -- the per-particle parameters as the CustomGBForce holds them.
 * solute/solvent dielectric: 1.0 / 78.5 are GBSAHCTForce's defaults.  sa_ace != 0 adds
 * 28.3919551 (radius + 0.14)^2 (radius/B)^6.  Refused (SDM_ERR_INVALID) for periodic methods. */
int sdm_enable_hct_gb(sdm_ctx* ctx, const double* charge, const double* offset_radius, const double* scaled_radius,
                      double solute_dielectric, double solvent_dielectric, int sa_ace);
/* Born radii B_i of state 1 or 2 (x or x + d) of one replica in the last evaluation (zeros before the first one);
 * synchronises.  HCT has no guard against 1/or - I <= 0 for deeply buried atoms (neither has OpenMM's expression): such
 * a configuration gives a negative or infinite radius and non-finite energies, which show up in sdm_scalars as they are. */
int sdm_get_born_radii(sdm_ctx* ctx, int replica, int state, double* radii);

/* ---- restraint forces of SDMUtils (the SDMUtils part) -------------------------------
 * What python/SDMUtils.py builds as OpenMM Custom*Forces in force group 1, evaluated on the device for every
 * replica inside sdm_eval(): the energy enters sdm_scalars.pot_energy like Eb, the forces are added to the
 * hybrid force like Fb (F1 stays the nonbonded state-1 force).  Units: kJ/mol, nm, radians.  Terms are kept
 * until sdm_clear_restraints(); index arrays are copied. */
typedef struct sdm_centroid_restraint {     /* SDMUtils.addRestraintForce, python/SDMUtils.py:32-162 */
    int32_t n_lig_cm, n_rcpt_cm;
    const int32_t* lig_cm_atoms;            /* g1: ligand atoms of the centroid          (:102) */
    const int32_t* rcpt_cm_atoms;           /* g2: receptor atoms of the centroid        (:103) */
    const double* lig_cm_weights;           /* NULL = equal weights; OpenMM's default is the particle masses */
    const double* rcpt_cm_weights;
    double kfcm, tolcm, offset[3];          /* (kfcm/2) step(d12-tolcm) (d12-tolcm)^2, d12 = |g1 - offset - g2| (:61,:68) */
    int32_t do_angles;                      /* 0: the distance term only                 (:55-59) */
    int32_t rcpt_ref[3], lig_ref[3];        /* g3..g5, g6..g8                            (:122-128) */
    double kfcd[3], a[3], b[3];             /* angle(g3,g6,g7), dihedral(g4,g3,g6,g7), dihedral(g3,g6,g7,g8):
                                               force constants and flat-bottom windows [a, b] (:63-83, :137-149) */
} sdm_centroid_restraint;
typedef struct sdm_alignment_restraint {    /* SDMUtils.addAlignmentForce, python/SDMUtils.py:166-258 */
    int32_t liga_ref[3], ligb_ref[3];
    double kfdispl, ktheta, kpsi, offset[3];
} sdm_alignment_restraint;
int sdm_add_centroid_restraint(sdm_ctx* ctx, const sdm_centroid_restraint* r);
int sdm_add_alignment_restraint(sdm_ctx* ctx, const sdm_alignment_restraint* r);
int sdm_clear_restraints(sdm_ctx* ctx);
/* The global parameter "SDMRestraintControlParameter" (SDMUtils.py:7,97): scales the centroid restraints. */
int sdm_set_restraint_control(sdm_ctx* ctx, double value);
/* Restraint energy of the last evaluation of one replica (already inside pot_energy); synchronises. */
int sdm_get_restraint_energy(sdm_ctx* ctx, int replica, double* energy);

/* Scalar half of execute() (ReferenceSDMKernels.cpp:205-302): from E1, E2, Eb and the
 * integrator state compute u_sc, fp, ebias, bfp, sp, PotEnergy, BindE and update the
 * non-equilibrium state in *alch.  O(1) host arithmetic, exactly as the reference does it on
 * the host for both of its platforms. */
int sdm_execute_scalars(sdm_alch* alch, double E1, double E2, double Eb, sdm_scalars* out);

#ifdef __cplusplus
}
#endif
#endif /* SDMB200_H_ */
