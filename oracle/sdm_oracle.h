/*
 * sdm_oracle.h -- CPU oracle for the SDM / ATM dual-state force path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (openmm_sdm_plugin_b200/,
 * include/sdmb200.h) may include, link or call this.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, as the
 * checker.
 *
 * PARITY UNPINNED BY THE REFERENCE: rajatkrpal/openmm_sdm_plugin ships no tests, no
 * golden vectors and cannot be compiled here (every TU needs OpenMM, which is not
 * vendored and not installed).  The oracle is pinned instead on
 *   - the scalar known answers of SURVEY.md Appendix A.4 (soft-core, ILogistic),
 *   - the fixture energies of SURVEY.md Appendix C (independent numpy script),
 *   - an independent O(N^2) numpy re-derivation in tests/test_oracle.py.
 *
 * What is restated, double precision, same evaluation order as the reference:
 *   plugin arithmetic  : openmmapi/src/LangevinIntegratorSDM.cpp:125-149 (SoftCoreF),
 *                        :153-183 (step: two full evaluations),
 *                        platforms/reference/src/ReferenceSDMKernels.cpp:161-199
 *                        (SaveState1/SaveState2/RestoreState1/MakeState2), :202-318
 *                        (bias, bookkeeping, non-equilibrium work, hybrid force).
 *   pair arithmetic    : OpenMM 7.2.2/7.3.1 Reference platform NonbondedForce
 *                        (README.md:38 names the version; OpenMM is NOT vendored):
 *                        ReferenceLJCoulombIxn::calculateOneIxn, ReferenceLJCoulomb14,
 *                        ReferenceNeighborList, NonbondedForceImpl::
 *                        calcDispersionCorrection, ONE_4PI_EPS0 = 138.935456.
 */
#ifndef SDM_ORACLE_H_
#define SDM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NOCUTOFF 0
#define ORC_CUTOFF_NONPERIODIC 1
#define ORC_CUTOFF_PERIODIC 2

#define ORC_ONE_4PI_EPS0 138.935456
#define ORC_SQRT_PI 1.7724538509055160273

/* Flat description of what reaches OpenMM's NonbondedForce for force group 2
 * (example/desmonddmsfile75.py:772-850). */
typedef struct {
    int32_t n_atoms;
    int32_t method;              /* ORC_* */
    double cutoff;               /* nm */
    double eps_rf;               /* reaction-field dielectric (78.3 default) */
    double box[3];               /* orthorhombic box edges, nm (periodic only) */
    int32_t use_dispersion_correction;
    int32_t n_exclusions;        /* every addException pair, zero or not */
    int32_t n_exceptions;        /* the subset with chargeProd != 0 or epsilon != 0 */
    int32_t pad_;
    const double* charge;        /* [n] e */
    const double* sigma;         /* [n] nm */
    const double* epsilon;       /* [n] kJ/mol */
    const int32_t* exclusions;   /* [2*n_exclusions] */
    const int32_t* exceptions;   /* [2*n_exceptions] */
    const double* exception_params; /* [3*n_exceptions] chargeProd, sigma, epsilon */
    double ewald_alpha;          /* > 0 (with ORC_CUTOFF_PERIODIC): NonbondedForce::Ewald / ::PME, DIRECT-SPACE part
                                    only -- erfc(alpha r)/r pair terms and the erf(alpha r)/r correction of the
                                    excluded pairs (ReferenceLJCoulombIxn::calculateEwaldIxn, includeDirect) */
    int32_t lj_geometric;        /* 1: the force group as createSystem(OPLS=True) builds it
                                    (example/desmonddmsfile75.py:780-810): NonbondedForce keeps the charges (every
                                    epsilon 0) and a CustomNonbondedForce adds 4 eps12 ((s12/r)^12 - (s12/r)^6),
                                    s12 = sqrt(s1 s2), eps12 = sqrt(eps1 eps2), same exclusions and cutoff, no
                                    long-range correction.  (ReferenceCustomNonbondedIxn drops r >= cutoff where
                                    NonbondedForce drops r > cutoff: the sets differ only at r == cutoff exactly.) */
    int32_t pad2_;
} orc_system;

/* Alchemical / soft-core state held by LangevinIntegratorSDM
 * (openmmapi/include/LangevinIntegratorSDM.h:499-521). */
typedef struct {
    int32_t bias_method;         /* 0 linear, 1 quadratic, 2 ilogistic */
    int32_t softcore_method;     /* 0 none, 1 tanh, 2 rational */
    double lambdac, gammac, wbcoeff, w0coeff;
    double lambda1, lambda2, alpha, u0;
    double umax, acore, ubcore;
    /* non-equilibrium mode (ReferenceSDMKernels.cpp:221-245, 289-302) */
    int32_t nonequilibrium;
    int32_t pad_;
    double noneq_tmax, work_value, time, step_size;
    double m_lambda1, m_lambda2, m_u0, m_w0;
    double b_lambda1, b_lambda2, b_u0, b_w0;
} orc_alch;

typedef struct {
    double E1, E2, Eb;           /* state energies handed to execute()         */
    double u;                    /* E2 - E1                                    */
    double u_sc, fp;             /* SoftCoreF                                  */
    double ebias, bfp;           /* bias energy and slope                      */
    double sp;                   /* bfp*fp                                     */
    double pot_energy;           /* E1 + ebias + Eb  (setPotEnergy)            */
    double bind_e;               /* u_sc            (setBindE)                 */
    double E1_pair, E1_exc, E1_disp; /* decomposition of E1                    */
    int64_t n_pairs1, n_pairs2;  /* in-cutoff non-excluded pairs per state     */
} orc_result;

/* LangevinIntegratorSDM::SoftCoreF.  Returns u_sc, writes fp.  *err = 1 for an
 * unknown method (the reference throws OpenMMException there). */
double orc_softcore(int method, double u, double umax, double a, double ub,
                    double* fp, int* err);

/* Bias energy / slope of ReferenceSDMKernels.cpp:247-282 (reads and, in
 * non-equilibrium mode, updates *alch exactly like execute() does). */
void orc_bias(orc_alch* alch, double bind_e, double* ebias, double* bfp);

/* One NonbondedForce evaluation (Reference platform): list build + pair loop +
 * exceptions + dispersion correction.  forces is [3n] and is OVERWRITTEN.
 * pairs (may be NULL) receives up to max_pairs (i,j) with i<j, sorted.
 * nthreads <= 1 reproduces the single-threaded Reference platform.
 * Returns 0, or <0 on error (box < 2*cutoff like the reference's throw). */
int orc_nonbonded(const orc_system* sys, const double* pos, double* forces,
                  double* e_pair, double* e_exc, double* e_disp,
                  int32_t* pairs, int64_t max_pairs, int64_t* n_pairs,
                  int nthreads);

/* The whole per-step force path of LangevinIntegratorSDM::step + execute up to the
 * hybrid force (no Langevin update):  two full evaluations, state copies, scalars,
 * mix.  displ is [3n] (the displacement map), fb is [3n] bonded forces or NULL,
 * f_out [3n] hybrid force; f1_out / f2_out optional [3n].  pos is left unchanged
 * (RestoreState1).  alch is updated in non-equilibrium mode. */
int orc_sdm_eval(const orc_system* sys, orc_alch* alch, const double* displ,
                 double* pos, const double* fb, double eb,
                 double* f_out, double* f1_out, double* f2_out,
                 orc_result* res, int nthreads);

/* Constant of NonbondedForceImpl::calcDispersionCorrection (divide by volume). */
double orc_dispersion_coefficient(const orc_system* sys);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
