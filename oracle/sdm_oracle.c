/*
 * sdm_oracle.c -- CPU oracle (double precision) for the SDM / ATM dual-state force
 * path.  TEST INFRASTRUCTURE ONLY; see sdm_oracle.h for the rules and for the
 * "parity unpinned" statement.
 *
 * Every function cites the reference lines it restates.  "OpenMM 7.3" marks
 * arithmetic that lives in the un-vendored dependency named at README.md:38 and is
 * restated from its published algorithm.
 */
#include "sdm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------
 * LangevinIntegratorSDM::SoftCoreF  (openmmapi/src/LangevinIntegratorSDM.cpp:125-149)
 * The u <= ub test comes BEFORE the method switch (:126-129); the tanh asymptote is
 * umax+ub (:134-138); the rational form is the .cpp one (:139-145), not the header
 * comment.
 * ---------------------------------------------------------------------------------- */
double orc_softcore(int method, double u, double umax, double a, double ub,
                    double* fp, int* err) {
    if (err) *err = 0;
    if (u <= ub) {
        *fp = 1.;
        return u;
    }
    if (method == 0) {
        *fp = 1.;
        return u;
    } else if (method == 1) {
        double x = (u - ub) / umax;
        double t = tanh(x);
        *fp = 1. - t * t;
        return umax * t + ub;
    } else if (method == 2) {
        double gu = (u - ub) / (a * (umax - ub));
        double zeta = 1. + 2. * gu * (gu + 1.);
        double zetap = pow(zeta, a);
        double s = 4. * (2. * gu + 1.) / zeta;
        *fp = s * zetap / pow(1. + zetap, 2);
        return (umax - ub) * (zetap - 1.) / (zetap + 1.) + ub;
    }
    if (err) *err = 1; /* reference: throw OpenMMException("Unknown soft core method") */
    *fp = 1.;
    return u;
}

/* ------------------------------------------------------------------------------------
 * Bias part of ReferenceIntegrateLangevinStepSDMKernel::execute
 * (platforms/reference/src/ReferenceSDMKernels.cpp:205-258 parameter fetch incl. the
 * non-equilibrium schedule, :266-282 bias energy and slope, :289-302 work).
 * ---------------------------------------------------------------------------------- */
void orc_bias(orc_alch* al, double bind_e, double* ebias_out, double* bfp_out) {
    double lambdac = al->lambdac;
    double dlambdac = 0.0;
    double gamma = 0.0;
    double wbcoeff = lambdac;
    double w0coeff = 0.0;
    double lambda1 = lambdac;
    double lambda2 = lambdac;
    double alpha = 1.0;
    double u0 = 0.0;
    if (al->nonequilibrium == 1) {
        /* :231-245 -- schedule; the reference's locals shadow, values go back into the
         * integrator and are re-read below for ILogistic only. */
        lambdac = al->time / al->noneq_tmax;
        al->lambdac = lambdac;
        al->lambda1 = al->m_lambda1 * lambdac + al->b_lambda1;
        al->lambda2 = al->m_lambda2 * lambdac + al->b_lambda2;
        al->u0 = al->m_u0 * lambdac + al->b_u0;
        al->w0coeff = al->m_w0 * lambdac + al->b_w0;
        dlambdac = al->step_size / al->noneq_tmax;
    }
    if (al->bias_method == 1) {
        gamma = al->gammac;
        wbcoeff = al->wbcoeff;
        w0coeff = al->w0coeff;
    } else if (al->bias_method == 2) {
        lambda1 = al->lambda1;
        lambda2 = al->lambda2;
        alpha = al->alpha;
        u0 = al->u0;
        w0coeff = al->w0coeff;
    }
    double bfp = 0.0, ebias = 0.0;
    if (al->bias_method == 1) {
        ebias = 0.5 * gamma * bind_e * bind_e + wbcoeff * bind_e + w0coeff;
        bfp = gamma * bind_e + wbcoeff;
    } else if (al->bias_method == 2) {
        double ee = 1.0 + exp(-alpha * (bind_e - u0));
        if (alpha > 0) ebias = ((lambda2 - lambda1) / alpha) * log(ee);
        ebias += lambda2 * bind_e + w0coeff;
        bfp = (lambda2 - lambda1) / ee + lambda1;
    } else {
        ebias = lambdac * bind_e;
        bfp = lambdac;
    }
    if (al->nonequilibrium == 1) {
        double ee = 1.0 + exp(-alpha * (bind_e - u0));
        double dwdl1 = -log(ee) / alpha;
        double dwdl2 = bind_e + (log(ee) / alpha);
        double dwdu0 = (lambda2 - lambda1) * exp(-alpha * (bind_e - u0)) / ee;
        double dwdw0 = 1;
        double dwdlambda = (dwdl1 * al->m_lambda1) + (dwdl2 * al->m_lambda2) +
                           (dwdu0 * al->m_u0) + (dwdw0 * al->m_w0);
        al->work_value = al->work_value + dlambdac * dwdlambda;
    }
    *ebias_out = ebias;
    *bfp_out = bfp;
}

/* ------------------------------------------------------------------------------------
 * OpenMM 7.3 NonbondedForceImpl::calcDispersionCorrection (no switching function).
 * Classes are (sigma, epsilon) pairs in map order.
 * ---------------------------------------------------------------------------------- */
typedef struct { double s, e; double count; } disp_class;

static int cmp_class(const void* a, const void* b) {
    const disp_class* x = (const disp_class*)a;
    const disp_class* y = (const disp_class*)b;
    if (x->s < y->s) return -1;
    if (x->s > y->s) return 1;
    if (x->e < y->e) return -1;
    if (x->e > y->e) return 1;
    return 0;
}

double orc_dispersion_coefficient(const orc_system* sys) {
    if (sys->method != ORC_CUTOFF_PERIODIC) return 0.0;
    int n = sys->n_atoms;
    disp_class* all = (disp_class*)malloc(sizeof(disp_class) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        all[i].s = sys->sigma[i];
        all[i].e = sys->epsilon[i];
        all[i].count = 1;
    }
    qsort(all, (size_t)n, sizeof(disp_class), cmp_class);
    int nc = 0;
    for (int i = 0; i < n; i++) {
        if (nc > 0 && all[nc - 1].s == all[i].s && all[nc - 1].e == all[i].e)
            all[nc - 1].count += 1;
        else
            all[nc++] = all[i];
    }
    double sum1 = 0, sum2 = 0;
    for (int c = 0; c < nc; c++) {
        double sigma = all[c].s, epsilon = all[c].e, count = all[c].count;
        count *= (count + 1) / 2;
        double sigma2 = sigma * sigma;
        double sigma6 = sigma2 * sigma2 * sigma2;
        sum1 += count * epsilon * sigma6 * sigma6;
        sum2 += count * epsilon * sigma6;
    }
    for (int c1 = 0; c1 < nc; c1++)
        for (int c2 = 0; c2 < c1; c2++) {
            double sigma = 0.5 * (all[c1].s + all[c2].s);
            double epsilon = sqrt(all[c1].e * all[c2].e);
            double count = all[c1].count * all[c2].count;
            double sigma2 = sigma * sigma;
            double sigma6 = sigma2 * sigma2 * sigma2;
            sum1 += count * epsilon * sigma6 * sigma6;
            sum2 += count * epsilon * sigma6;
        }
    free(all);
    double np = (double)n;
    double ni = (np * (np + 1)) / 2;
    sum1 /= ni;
    sum2 /= ni;
    double rc = sys->cutoff;
    return 8 * np * np * M_PI * (sum1 / (9 * pow(rc, 9)) - sum2 / (3 * pow(rc, 3)));
}

/* ------------------------------------------------------------------------------------
 * Neighbour list (OpenMM 7.3 ReferenceNeighborList semantics): every i<j that is not
 * an exclusion and has r^2 <= rc^2 (pairs are dropped only when r^2 > rc^2), minimum
 * image per component for the orthorhombic box, rebuilt from scratch on every call,
 * no skin.  A cell grid replaces the voxel hash; the resulting SET is the same.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t* data;
    int64_t n, cap;
} pairvec;

static void pv_push(pairvec* v, int32_t i, int32_t j) {
    if (v->n + 2 > v->cap) {
        v->cap = v->cap ? v->cap * 2 : (1 << 16);
        v->data = (int32_t*)realloc(v->data, sizeof(int32_t) * (size_t)v->cap);
    }
    v->data[v->n++] = i;
    v->data[v->n++] = j;
}

static inline double min_image(double d, double L) { return d - floor(d / L + 0.5) * L; }

static int cmp_pair(const void* a, const void* b) {
    const int32_t* x = (const int32_t*)a;
    const int32_t* y = (const int32_t*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}

typedef struct {
    int32_t* start; /* CSR [n+1] */
    int32_t* idx;
} excl_csr;

static void build_excl(const orc_system* sys, excl_csr* ex) {
    int n = sys->n_atoms;
    ex->start = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    for (int k = 0; k < sys->n_exclusions; k++) {
        int a = sys->exclusions[2 * k], b = sys->exclusions[2 * k + 1];
        if (a == b) continue;
        ex->start[a + 1]++;
        ex->start[b + 1]++;
    }
    for (int i = 0; i < n; i++) ex->start[i + 1] += ex->start[i];
    ex->idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ex->start[n] > 0 ? ex->start[n] : 1));
    int32_t* fill = (int32_t*)calloc((size_t)n, sizeof(int32_t));
    for (int k = 0; k < sys->n_exclusions; k++) {
        int a = sys->exclusions[2 * k], b = sys->exclusions[2 * k + 1];
        if (a == b) continue;
        ex->idx[ex->start[a] + fill[a]++] = b;
        ex->idx[ex->start[b] + fill[b]++] = a;
    }
    free(fill);
}

/* Builds per-thread pair lists; returns them in lists[0..nt). */
static int build_neighbor_list(const orc_system* sys, const double* pos, const excl_csr* ex,
                               pairvec* lists, int nt) {
    const int n = sys->n_atoms;
    const int periodic = sys->method == ORC_CUTOFF_PERIODIC;
    const int nocut = sys->method == ORC_NOCUTOFF;
    const double rc2 = sys->cutoff * sys->cutoff;

    if (nocut) {
#pragma omp parallel num_threads(nt)
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            int32_t* mark = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
            for (int i = 0; i < n; i++) mark[i] = -1;
#pragma omp for schedule(dynamic, 16)
            for (int i = 0; i < n; i++) {
                for (int k = ex->start[i]; k < ex->start[i + 1]; k++) mark[ex->idx[k]] = i;
                for (int j = i + 1; j < n; j++)
                    if (mark[j] != i) pv_push(&lists[t], i, j);
            }
            free(mark);
        }
        return 0;
    }

    /* grid */
    double lo[3], ext[3];
    int nc[3];
    if (periodic) {
        for (int d = 0; d < 3; d++) {
            if (2 * sys->cutoff > sys->box[d]) return -2; /* OpenMM throws */
            lo[d] = 0;
            ext[d] = sys->box[d];
            nc[d] = (int)floor(sys->box[d] / sys->cutoff);
            if (nc[d] < 1) nc[d] = 1;
        }
    } else {
        double hi[3];
        for (int d = 0; d < 3; d++) { lo[d] = 1e300; hi[d] = -1e300; }
        for (int i = 0; i < n; i++)
            for (int d = 0; d < 3; d++) {
                if (pos[3 * i + d] < lo[d]) lo[d] = pos[3 * i + d];
                if (pos[3 * i + d] > hi[d]) hi[d] = pos[3 * i + d];
            }
        for (int d = 0; d < 3; d++) {
            ext[d] = hi[d] - lo[d];
            if (!(ext[d] > 0)) ext[d] = 1e-9;
            nc[d] = (int)floor(ext[d] / sys->cutoff);
            if (nc[d] < 1) nc[d] = 1;
            if (nc[d] > 256) nc[d] = 256;
        }
    }
    /* a cell must be at least cutoff wide so that the 27 neighbours suffice */
    double side[3];
    for (int d = 0; d < 3; d++) side[d] = ext[d] / nc[d];
    const int ncell = nc[0] * nc[1] * nc[2];
    int32_t* cell_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* cstart = (int32_t*)calloc((size_t)ncell + 1, sizeof(int32_t));
    int32_t* catoms = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            double x = pos[3 * i + d];
            if (periodic) x -= floor(x / ext[d]) * ext[d];
            int k = (int)floor((x - lo[d]) / side[d]);
            if (k < 0) k = 0;
            if (k >= nc[d]) k = nc[d] - 1;
            c[d] = k;
        }
        cell_of[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
        cstart[cell_of[i] + 1]++;
    }
    for (int c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    {
        int32_t* fill = (int32_t*)calloc((size_t)ncell, sizeof(int32_t));
        for (int i = 0; i < n; i++) catoms[cstart[cell_of[i]] + fill[cell_of[i]]++] = i;
        free(fill);
    }

#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        int32_t* mark = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
        for (int i = 0; i < n; i++) mark[i] = -1;
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            for (int k = ex->start[i]; k < ex->start[i + 1]; k++) mark[ex->idx[k]] = i;
            int ci = cell_of[i];
            int c0 = ci % nc[0], c1 = (ci / nc[0]) % nc[1], c2 = ci / (nc[0] * nc[1]);
            int seen[27], nseen = 0;
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        int a = c0 + dx, b = c1 + dy, c = c2 + dz;
                        if (periodic) {
                            a = (a + nc[0]) % nc[0];
                            b = (b + nc[1]) % nc[1];
                            c = (c + nc[2]) % nc[2];
                        } else if (a < 0 || b < 0 || c < 0 || a >= nc[0] || b >= nc[1] || c >= nc[2])
                            continue;
                        int cj = (c * nc[1] + b) * nc[0] + a;
                        int dup = 0;
                        for (int s = 0; s < nseen; s++)
                            if (seen[s] == cj) dup = 1;
                        if (dup) continue;
                        seen[nseen++] = cj;
                        for (int k = cstart[cj]; k < cstart[cj + 1]; k++) {
                            int j = catoms[k];
                            if (j <= i || mark[j] == i) continue;
                            double dx_ = pos[3 * i] - pos[3 * j];
                            double dy_ = pos[3 * i + 1] - pos[3 * j + 1];
                            double dz_ = pos[3 * i + 2] - pos[3 * j + 2];
                            if (periodic) {
                                dx_ = min_image(dx_, sys->box[0]);
                                dy_ = min_image(dy_, sys->box[1]);
                                dz_ = min_image(dz_, sys->box[2]);
                            }
                            double r2 = dx_ * dx_ + dy_ * dy_ + dz_ * dz_;
                            if (r2 > rc2) continue;
                            pv_push(&lists[t], i, j);
                        }
                    }
        }
        free(mark);
    }
    free(cell_of);
    free(cstart);
    free(catoms);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * One NonbondedForce evaluation, Reference platform:
 *   pair loop   -- OpenMM 7.3 ReferenceLJCoulombIxn::calculateOneIxn with per-atom
 *                  half-sigma / 2*sqrt(eps) parameters, reaction field when a cutoff is
 *                  used (krf, crf), plain Coulomb for NoCutoff, LJ not shifted.
 *   exceptions  -- OpenMM 7.3 ReferenceLJCoulomb14::calculateBondIxn (no cutoff, no RF,
 *                  plain delta).
 *   dispersion  -- + coefficient / volume for CutoffPeriodic.
 *   Ewald / PME -- sys->ewald_alpha > 0: the direct-space part of OpenMM 7.3
 *                  ReferenceLJCoulombIxn::calculateEwaldIxn: Coulomb term qq*erfc(alpha r)/r in
 *                  the pair loop, then "subtract off the exclusions": -qq*erf(alpha r)/r for every
 *                  excluded pair (minimum image), -qq*2*alpha/sqrt(pi) without a force where
 *                  erf(alpha r) <= 1e-6.  Reciprocal space and self energy (includeReciprocal)
 *                  are NOT part of this function; tests/test_oracle.py adds them in numpy to
 *                  check the split.
 * Call sites in the reference: LangevinIntegratorSDM.cpp:160,168
 * (context->calcForcesAndEnergy(true, true, 4)).
 * ---------------------------------------------------------------------------------- */
int orc_nonbonded(const orc_system* sys, const double* pos, double* forces,
                  double* e_pair, double* e_exc, double* e_disp,
                  int32_t* pairs, int64_t max_pairs, int64_t* n_pairs, int nthreads) {
    const int n = sys->n_atoms;
    int nt = nthreads < 1 ? 1 : nthreads;
#ifndef _OPENMP
    nt = 1;
#endif
    const int periodic = sys->method == ORC_CUTOFF_PERIODIC;
    const int cutoff = sys->method != ORC_NOCUTOFF;

    excl_csr ex;
    build_excl(sys, &ex);
    pairvec* lists = (pairvec*)calloc((size_t)nt, sizeof(pairvec));
    int rc = build_neighbor_list(sys, pos, &ex, lists, nt);
    free(ex.start);
    free(ex.idx);
    if (rc < 0) {
        for (int t = 0; t < nt; t++) free(lists[t].data);
        free(lists);
        return rc;
    }

    const double ewald_alpha = periodic ? sys->ewald_alpha : 0.0;
    double krf = 0, crf = 0;
    if (cutoff) {
        double rcut = sys->cutoff, es = sys->eps_rf;
        krf = pow(rcut, -3.0) * (es - 1.0) / (2.0 * es + 1.0);
        crf = (1.0 / rcut) * (3.0 * es) / (2.0 * es + 1.0);
    }

    /* per-atom parameters as the Reference kernel stores them */
    double* hsig = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double* heps = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        hsig[i] = 0.5 * sys->sigma[i];
        heps[i] = 2.0 * sqrt(sys->epsilon[i]);
    }

    double* ftmp = (double*)calloc((size_t)nt * 3 * (size_t)(n > 0 ? n : 1), sizeof(double));
    double* etmp = (double*)calloc((size_t)nt, sizeof(double));
    int64_t total_pairs = 0;
    for (int t = 0; t < nt; t++) total_pairs += lists[t].n / 2;

#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        double* f = ftmp + (size_t)t * 3 * n;
        double energy = 0;
        const pairvec* L = &lists[t];
        for (int64_t p = 0; p < L->n; p += 2) {
            int ii = L->data[p], jj = L->data[p + 1];
            double d[3];
            for (int k = 0; k < 3; k++) d[k] = pos[3 * ii + k] - pos[3 * jj + k];
            if (periodic)
                for (int k = 0; k < 3; k++) d[k] = min_image(d[k], sys->box[k]);
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double r = sqrt(r2);
            double inverseR = 1.0 / r;
            double sig = hsig[ii] + hsig[jj];
            double sig2 = inverseR * sig;
            sig2 *= sig2;
            double sig6 = sig2 * sig2 * sig2;
            double eps = heps[ii] * heps[jj];
            double qq = ORC_ONE_4PI_EPS0 * sys->charge[ii] * sys->charge[jj];
            double dEdR = eps * (12.0 * sig6 - 6.0) * sig6;
            double e = eps * (sig6 - 1.0) * sig6;
            if (sys->lj_geometric) {
                /* the CustomNonbondedForce expression of desmonddmsfile75.py:781 and its -r dE/dr */
                double sigma12 = sqrt(sys->sigma[ii] * sys->sigma[jj]);
                double epsilon12 = sqrt(sys->epsilon[ii] * sys->epsilon[jj]);
                double x6 = pow(sigma12 / r, 6.0), x12 = pow(sigma12 / r, 12.0);
                e = 4.0 * epsilon12 * (x12 - x6);
                dEdR = 4.0 * epsilon12 * (12.0 * x12 - 6.0 * x6);
            }
            if (ewald_alpha > 0) {
                double alphaR = ewald_alpha * r;
                dEdR += qq * inverseR * (erfc(alphaR) + 2.0 * alphaR * exp(-alphaR * alphaR) / ORC_SQRT_PI);
                e += qq * inverseR * erfc(alphaR);
            } else if (cutoff) {
                dEdR += qq * (inverseR - 2.0 * krf * r2);
                e += qq * (inverseR + krf * r2 - crf);
            } else {
                dEdR += qq * inverseR;
                e += qq * inverseR;
            }
            dEdR *= inverseR * inverseR;
            for (int k = 0; k < 3; k++) {
                double fk = dEdR * d[k];
                f[3 * ii + k] += fk;
                f[3 * jj + k] -= fk;
            }
            energy += e;
        }
        etmp[t] = energy;
    }
    double epair = 0;
    memset(forces, 0, sizeof(double) * 3 * (size_t)n);
    for (int t = 0; t < nt; t++) {
        epair += etmp[t];
        const double* f = ftmp + (size_t)t * 3 * n;
        for (int i = 0; i < 3 * n; i++) forces[i] += f[i];
    }
    free(ftmp);
    free(etmp);
    free(hsig);
    free(heps);

    /* exceptions (1-4) */
    double eexc = 0;
    for (int k = 0; k < sys->n_exceptions; k++) {
        int a = sys->exceptions[2 * k], b = sys->exceptions[2 * k + 1];
        double qq = sys->exception_params[3 * k];
        double sigma = sys->exception_params[3 * k + 1];
        double eps4 = 4.0 * sys->exception_params[3 * k + 2];
        if (qq == 0.0 && sys->exception_params[3 * k + 2] == 0.0) continue;
        double d[3];
        for (int c = 0; c < 3; c++) d[c] = pos[3 * a + c] - pos[3 * b + c];
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        double inverseR = 1.0 / sqrt(r2);
        double sig2 = inverseR * sigma;
        sig2 *= sig2;
        double sig6 = sig2 * sig2 * sig2;
        double dEdR = eps4 * (12.0 * sig6 - 6.0) * sig6;
        dEdR += ORC_ONE_4PI_EPS0 * qq * inverseR;
        dEdR *= inverseR * inverseR;
        for (int c = 0; c < 3; c++) {
            double fk = dEdR * d[c];
            forces[3 * a + c] += fk;
            forces[3 * b + c] -= fk;
        }
        eexc += eps4 * (sig6 - 1.0) * sig6 + ORC_ONE_4PI_EPS0 * qq * inverseR;
    }

    /* Ewald: take the erf part of the excluded pairs out again (booked with the exceptions) */
    if (ewald_alpha > 0 && sys->n_exclusions > 0) {
        /* every unordered pair once, whatever the input lists: sorted (min, max) keys */
        int32_t* key = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)sys->n_exclusions);
        for (int k = 0; k < sys->n_exclusions; k++) {
            int a = sys->exclusions[2 * k], b = sys->exclusions[2 * k + 1];
            key[2 * k] = a < b ? a : b;
            key[2 * k + 1] = a < b ? b : a;
        }
        qsort(key, (size_t)sys->n_exclusions, 2 * sizeof(int32_t), cmp_pair);
        for (int k = 0; k < sys->n_exclusions; k++) {
            int a = key[2 * k], b = key[2 * k + 1];
            if (a == b) continue;
            if (k > 0 && key[2 * k - 2] == a && key[2 * k - 1] == b) continue;
            double d[3];
            for (int c = 0; c < 3; c++) d[c] = min_image(pos[3 * a + c] - pos[3 * b + c], sys->box[c]);
            double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            double alphaR = ewald_alpha * r;
            double qq = ORC_ONE_4PI_EPS0 * sys->charge[a] * sys->charge[b];
            if (erf(alphaR) > 1e-6) {
                double inverseR = 1.0 / r;
                double dEdR = qq * inverseR * inverseR * inverseR;
                dEdR = dEdR * (erf(alphaR) - 2.0 * alphaR * exp(-alphaR * alphaR) / ORC_SQRT_PI);
                for (int c = 0; c < 3; c++) {
                    forces[3 * a + c] -= dEdR * d[c];
                    forces[3 * b + c] += dEdR * d[c];
                }
                eexc -= qq * inverseR * erf(alphaR);
            } else {
                eexc -= ewald_alpha * 2.0 / ORC_SQRT_PI * qq;
            }
        }
        free(key);
    }

    double edisp = 0;
    if (periodic && sys->use_dispersion_correction)
        edisp = orc_dispersion_coefficient(sys) / (sys->box[0] * sys->box[1] * sys->box[2]);

    if (e_pair) *e_pair = epair;
    if (e_exc) *e_exc = eexc;
    if (e_disp) *e_disp = edisp;
    if (n_pairs) *n_pairs = total_pairs;
    if (pairs) {
        int64_t w = 0;
        for (int t = 0; t < nt && w < max_pairs; t++)
            for (int64_t p = 0; p < lists[t].n && w < max_pairs; p += 2, w++) {
                pairs[2 * w] = lists[t].data[p];
                pairs[2 * w + 1] = lists[t].data[p + 1];
            }
        qsort(pairs, (size_t)w, 2 * sizeof(int32_t), cmp_pair);
    }
    for (int t = 0; t < nt; t++) free(lists[t].data);
    free(lists);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * LangevinIntegratorSDM::step up to the hybrid force
 * (openmmapi/src/LangevinIntegratorSDM.cpp:156-182) with the Reference kernel's state
 * copies (ReferenceSDMKernels.cpp:161-199) and execute() (:202-318).
 * ---------------------------------------------------------------------------------- */
int orc_sdm_eval(const orc_system* sys, orc_alch* alch, const double* displ, double* pos,
                 const double* fb, double eb, double* f_out, double* f1_out, double* f2_out,
                 orc_result* res, int nthreads) {
    const int n = sys->n_atoms;
    const size_t n3 = 3 * (size_t)n;
    double* forces = (double*)malloc(sizeof(double) * (n3 ? n3 : 1));
    double* State1Forces = (double*)malloc(sizeof(double) * (n3 ? n3 : 1));
    double* State2Forces = (double*)malloc(sizeof(double) * (n3 ? n3 : 1));
    double* State1Coordinates = (double*)malloc(sizeof(double) * (n3 ? n3 : 1));
    memset(res, 0, sizeof(*res));
    int rc;

    /* :160  State1Energy = calcForcesAndEnergy(true, true, 4) */
    double ep, ee, ed;
    rc = orc_nonbonded(sys, pos, forces, &ep, &ee, &ed, NULL, 0, &res->n_pairs1, nthreads);
    if (rc < 0) goto done;
    res->E1_pair = ep;
    res->E1_exc = ee;
    res->E1_disp = ed;
    res->E1 = ep + ee + ed;
    /* :162  SaveState1 */
    memcpy(State1Forces, forces, sizeof(double) * n3);
    memcpy(State1Coordinates, pos, sizeof(double) * n3);
    /* :165  MakeState2  (posData[p] += State2Displacement[p], every atom) */
    for (size_t k = 0; k < n3; k++) pos[k] += displ[k];
    /* :168  State2Energy */
    rc = orc_nonbonded(sys, pos, forces, &ep, &ee, &ed, NULL, 0, &res->n_pairs2, nthreads);
    if (rc < 0) {
        memcpy(pos, State1Coordinates, sizeof(double) * n3);
        goto done;
    }
    res->E2 = ep + ee + ed;
    /* :170  SaveState2 */
    memcpy(State2Forces, forces, sizeof(double) * n3);
    /* :173  RestoreState1 */
    memcpy(pos, State1Coordinates, sizeof(double) * n3);
    /* :176  bonded/restraint evaluation leaves Fb in the force buffer, returns Eb */
    if (fb)
        memcpy(forces, fb, sizeof(double) * n3);
    else
        memset(forces, 0, sizeof(double) * n3);
    res->Eb = eb;

    /* :180 execute() */
    {
        int err = 0;
        res->u = res->E2 - res->E1;
        res->u_sc = orc_softcore(alch->softcore_method, res->u, alch->umax, alch->acore,
                                 alch->ubcore, &res->fp, &err);
        if (err) { rc = -3; goto done; }
        orc_bias(alch, res->u_sc, &res->ebias, &res->bfp);
        res->pot_energy = res->E1 + res->ebias + res->Eb;
        res->bind_e = res->u_sc;
        double sp = res->bfp * res->fp;
        res->sp = sp;
        for (size_t k = 0; k < n3; k++)
            forces[k] = sp * State2Forces[k] + (1.0 - sp) * State1Forces[k] + forces[k];
        alch->time += alch->step_size; /* data.time += stepSize (:340) */
    }
    if (f_out) memcpy(f_out, forces, sizeof(double) * n3);
    if (f1_out) memcpy(f1_out, State1Forces, sizeof(double) * n3);
    if (f2_out) memcpy(f2_out, State2Forces, sizeof(double) * n3);
    rc = 0;
done:
    free(forces);
    free(State1Forces);
    free(State2Forces);
    free(State1Coordinates);
    return rc;
}
