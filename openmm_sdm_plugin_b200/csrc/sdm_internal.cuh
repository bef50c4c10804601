// sdm_internal.cuh -- shared device/host declarations of libsdmb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdmb200.h"

#define SDM_K_COULOMB 138.935456              // ONE_4PI_EPS0, OpenMM 7.x
#define SDM_FORCE_SCALE 4294967296.0          // 2^32 fixed-point force accumulators
#define SDM_INV_FORCE_SCALE (1.0 / 4294967296.0)

namespace sdm {

// ---------------------------------------------------------------------------------------------
// Topology shared by all replicas of a context (device pointers, System particle order).
// ---------------------------------------------------------------------------------------------
struct Topology {
    int n;                 // atoms
    int method;            // SDM_NOCUTOFF / SDM_CUTOFF_NONPERIODIC / SDM_CUTOFF_PERIODIC
    int n_lig;             // displaced atoms (non-zero row of the displacement map)
    int n_exceptions;
    double rc, rc2, krf, crf;
    double box[3], inv_box[3];
    float rc2f, krff, crff, band;      // band: |r2-rc2| below which the FP64 re-test runs
    float boxf[3], inv_boxf[3];
    double e_disp;                      // dispersion correction energy (coefficient / volume)
    int ewald;                          // 1: SDM_EWALD / SDM_PME -- direct-space Ewald Coulomb instead of the reaction field
    double alpha;                       // Ewald splitting parameter (1/nm)
    float alphaf;
    int n_excl_pairs;                   // excluded pairs (i < j) that get the erf(alpha r)/r correction
    const int* excl_pairs;              // [2*n_excl_pairs]
    const double* q;       // [n] charge
    int lj_geom;           // 1: SDM_LJ_GEOMETRIC -- sigma_ij = sqrt(sigma_i sigma_j); hsig / parf.y then hold sigma itself
    const double* hsig;    // [n] sigma/2 (Lorentz-Berthelot) or sigma (geometric rule)
    const double* heps;    // [n] 2*sqrt(eps)
    const float4* parf;    // [n] (q*sqrt(K), sigma/2 or sigma, 2*sqrt(eps), 0) float
    const double* disp;    // [3n] displacement map
    const int* group;      // [n] id of the displacement vector (0 = not displaced)
    const int* lig_idx;    // [n_lig] displaced atoms, ascending
    const int* lig_flags;  // [n_lig] bit 0: has an exclusion with a NON-displaced atom
    const int* excl_start; // [n+1] CSR over both directions, rows ascending
    const int* excl_idx;
    const int* exc_pairs;      // [2*n_exceptions]
    const double* exc_params;  // [3*n_exceptions] chargeProd, sigma, epsilon
};

// Per-replica mutable state on the device.
struct ReplicaState {
    sdm_alch alch;
    sdm_scalars sc;
};

// ---------------------------------------------------------------------------------------------
// Scalar arithmetic of the plugin, shared by the device scalar kernel and the host entry point
// sdm_execute_scalars().  Restates LangevinIntegratorSDM::SoftCoreF
// (openmmapi/src/LangevinIntegratorSDM.cpp:125-149) and the bias block of execute()
// (platforms/reference/src/ReferenceSDMKernels.cpp:205-302).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double softcore(int method, double u, double umax, double a, double ub,
                                           double* fp, int* err) {
    *err = 0;
    if (u <= ub) { *fp = 1.0; return u; }
    if (method == SDM_SOFTCORE_NONE) { *fp = 1.0; return u; }
    if (method == SDM_SOFTCORE_TANH) {
        double x = (u - ub) / umax;
        double t = tanh(x);
        *fp = 1.0 - t * t;
        return umax * t + ub;
    }
    if (method == SDM_SOFTCORE_RATIONAL) {
        double gu = (u - ub) / (a * (umax - ub));
        double zeta = 1.0 + 2.0 * gu * (gu + 1.0);
        double zetap = pow(zeta, a);
        double s = 4.0 * (2.0 * gu + 1.0) / zeta;
        double d = 1.0 + zetap;
        *fp = s * zetap / (d * d);
        return (umax - ub) * (zetap - 1.0) / (zetap + 1.0) + ub;
    }
    *err = 1;
    *fp = 1.0;
    return u;
}

// Fills sc->u_sc .. sc->bind_e from sc->E1, sc->u, sc->Eb and updates the non-equilibrium
// fields of *al the way execute() writes them back into the integrator.
__host__ __device__ inline void execute_scalars(sdm_alch* al, sdm_scalars* sc) {
    double lambdac = al->lambdac;
    double dlambdac = 0.0;
    double gamma = 0.0, wbcoeff = lambdac, w0coeff = 0.0;
    double lambda1 = lambdac, lambda2 = lambdac, alpha = 1.0, u0 = 0.0;
    if (al->nonequilibrium == 1) {
        lambdac = al->time / al->noneq_tmax;
        al->lambdac = lambdac;
        al->lambda1 = al->m_lambda1 * lambdac + al->b_lambda1;
        al->lambda2 = al->m_lambda2 * lambdac + al->b_lambda2;
        al->u0 = al->m_u0 * lambdac + al->b_u0;
        al->w0coeff = al->m_w0 * lambdac + al->b_w0;
        dlambdac = al->step_size / al->noneq_tmax;
    }
    if (al->bias_method == SDM_BIAS_QUADRATIC) {
        gamma = al->gammac; wbcoeff = al->wbcoeff; w0coeff = al->w0coeff;
    } else if (al->bias_method == SDM_BIAS_ILOGISTIC) {
        lambda1 = al->lambda1; lambda2 = al->lambda2; alpha = al->alpha; u0 = al->u0;
        w0coeff = al->w0coeff;
    }
    int err = 0;
    double fp;
    double B = softcore(al->softcore_method, sc->u, al->umax, al->acore, al->ubcore, &fp, &err);
    if (err) sc->status = SDM_ERR_SOFTCORE;
    double bfp = 0.0, ebias = 0.0;
    if (al->bias_method == SDM_BIAS_QUADRATIC) {
        ebias = 0.5 * gamma * B * B + wbcoeff * B + w0coeff;
        bfp = gamma * B + wbcoeff;
    } else if (al->bias_method == SDM_BIAS_ILOGISTIC) {
        double ee = 1.0 + exp(-alpha * (B - u0));
        if (alpha > 0) ebias = ((lambda2 - lambda1) / alpha) * log(ee);
        ebias += lambda2 * B + w0coeff;
        bfp = (lambda2 - lambda1) / ee + lambda1;
    } else {
        ebias = lambdac * B;
        bfp = lambdac;
    }
    sc->u_sc = B;
    sc->fp = fp;
    sc->ebias = ebias;
    sc->bfp = bfp;
    sc->sp = bfp * fp;
    sc->pot_energy = sc->E1 + ebias + sc->Eb;
    sc->bind_e = B;
    if (al->nonequilibrium == 1) {
        double ee = 1.0 + exp(-alpha * (B - u0));
        double dwdl1 = -log(ee) / alpha;
        double dwdl2 = B + (log(ee) / alpha);
        double dwdu0 = (lambda2 - lambda1) * exp(-alpha * (B - u0)) / ee;
        double dwdlambda = (dwdl1 * al->m_lambda1) + (dwdl2 * al->m_lambda2) +
                           (dwdu0 * al->m_u0) + al->m_w0;
        al->work_value = al->work_value + dlambdac * dwdlambda;
    }
    al->time += al->step_size;  // data.time += stepSize (ReferenceSDMKernels.cpp:340)
}

// ---------------------------------------------------------------------------------------------
// Double-precision pair term, evaluation order of OpenMM 7.3's
// ReferenceLJCoulombIxn::calculateOneIxn.  d = x_i - x_j.
// Returns dEdR/r^2-scaled factor so that F_i += fs*d, F_j -= fs*d; *e is the pair energy.
// ---------------------------------------------------------------------------------------------
// alpha > 0: direct-space Ewald Coulomb (OpenMM 7.3 ReferenceLJCoulombIxn::calculateEwaldIxn, the
// "SHORT-RANGE ENERGY AND FORCES" loop): qq*erfc(alpha r)/r, dE/dr*r = qq*(erfc(alpha r) + 2 alpha r
// exp(-alpha^2 r^2)/sqrt(pi))/r.
__device__ __forceinline__ double pair_term_f64(double r2, double sig, double eps, double qq,
                                                bool cutoff, double krf, double crf, double* e, double alpha = 0.0) {
    double inverseR = rsqrt(r2);  // within 2 ulp of 1/sqrt(r2) and three times cheaper
    double sig2 = inverseR * sig;
    sig2 *= sig2;
    double sig6 = sig2 * sig2 * sig2;
    double dEdR = eps * (12.0 * sig6 - 6.0) * sig6;
    double en = eps * (sig6 - 1.0) * sig6;
    if (alpha > 0.0) {
        const double alphaR = alpha * r2 * inverseR;
        const double ec = erfc(alphaR);
        dEdR += qq * inverseR * (ec + alphaR * exp(-alphaR * alphaR) * 1.1283791670955126);   // 2/sqrt(pi)
        en += qq * inverseR * ec;
    } else if (cutoff) {
        dEdR += qq * (inverseR - 2.0 * krf * r2);
        en += qq * (inverseR + krf * r2 - crf);
    } else {
        dEdR += qq * inverseR;
        en += qq * inverseR;
    }
    *e = en;
    return dEdR * inverseR * inverseR;
}

// Exact (FP64) in-cutoff decision from the double positions, same expression as the oracle /
// OpenMM Reference neighbour list: r^2 <= rc^2 keeps the pair.
// No FMA contraction here: the host evaluates the same expression with separate roundings.
__device__ __forceinline__ double min_image_exact(double d, double L) {
    return __dsub_rn(d, __dmul_rn(floor(__dadd_rn(__ddiv_rn(d, L), 0.5)), L));
}

__device__ __forceinline__ double norm2_exact(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Same result as min_image_exact with the image INDEX chosen by a reciprocal multiplication (no
// FP64 division): the two can only disagree when d/L + 1/2 is within an ulp of an integer, i.e.
// |d_min| = L/2 >= cutoff, where the pair is out of range (or exactly at r = rc = L/2, where both
// images give the same r^2) either way.  Used by the displaced-atom kernels.
__device__ __forceinline__ double min_image_fast(double d, double L, double invL) {
    return __dsub_rn(d, __dmul_rn(floor(__fma_rn(d, invL, 0.5)), L));
}

__device__ __forceinline__ bool in_cutoff_f64(const Topology& T, const double* __restrict__ pos,
                                              int i, int j) {
    double dx = pos[3 * i] - pos[3 * j];
    double dy = pos[3 * i + 1] - pos[3 * j + 1];
    double dz = pos[3 * i + 2] - pos[3 * j + 2];
    if (T.method == SDM_CUTOFF_PERIODIC) {
        dx = min_image_exact(dx, T.box[0]);
        dy = min_image_exact(dy, T.box[1]);
        dz = min_image_exact(dz, T.box[2]);
    }
    return norm2_exact(dx, dy, dz) <= T.rc2;
}

// Is (i, j) an exclusion?  Rows are short (bonded neighbours), ascending.
__device__ __forceinline__ bool is_excluded(const Topology& T, int i, int j) {
    int lo = T.excl_start[i], hi = T.excl_start[i + 1];
    for (int k = lo; k < hi; k++) {
        int v = T.excl_idx[k];
        if (v == j) return true;
        if (v > j) return false;
    }
    return false;
}

__device__ __forceinline__ long long to_fixed(double f) {
    return __double2ll_rn(f * SDM_FORCE_SCALE);
}

__device__ __forceinline__ void atomic_add_fixed(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

}  // namespace sdm
