// kernels_pme.cu -- reciprocal-space smooth particle-mesh Ewald for both states of every resident replica
// (what example/test_explicit.py:64 asks of OpenMM with nonbondedMethod=PME).
//
// Restated from OpenMM 7.3 ReferencePME.cpp (B-splines of order 5: update_bsplines, grid_spread_charge,
// reciprocal_convolution, grid_interpolate_force, pme_calculate_bsplines_moduli) and the self energy that
// ReferenceLJCoulombIxn::calculateEwaldIxn books with the reciprocal part; orthorhombic box; FP64.
//
// The reciprocal sum is global in the charge density, so the moved-pairs-only trick of the direct-space path
// does not apply: state 1 (x) and state 2 (x + d) each get a full pass
//     spread -> forward FFT -> multiply by the influence function (+ energy) -> backward FFT -> gather,
// 2R grids per evaluation, batched.  Real-to-complex transforms (half spectrum) and a real potential grid: the
// gather is bound by L2 traffic (125 grid values per atom and state) and reads 8 instead of 16 bytes per value.
// The charge grid of state 2 is the one of state 1 plus the difference of the displaced atoms (spreading is
// linear), so the second spreading pass touches 38 atoms instead of 20 446.  The results land in the
// external dual-state slots of the evaluation (EvalBuffers::ext_*): E1 += E_rec(x), u += E_rec(x+d) - E_rec(x), F1 += F_rec(x), F2 - F1 likewise -- the
// scalar stage and the mix kernel need no change.
//
// Determinism: charges are spread with 64-bit fixed-point atomics (2^-40 e per count), so the grid -- and with
// it every energy and force -- does not depend on the order in which atoms arrive; the energy partials are
// added in a fixed order.  The 3-D transforms are cuFFT's (a plain library FFT; loaded with dlopen when
// the reciprocal part is switched on, so the library does not depend on cuFFT otherwise).
#include <cufft.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "sdm_ctx.h"
#include "sdm_internal.cuh"

namespace sdm {

constexpr int kPmeOrder = 5;
constexpr double kPmeFix = 1099511627776.0;   // 2^40

struct PmeState {
    int K[3] = {0, 0, 0};
    int R = 0, n = 0;
    size_t glen = 0;                   // grid points per grid
    double alpha = 0, box[3] = {0, 0, 0};
    size_t slen = 0;                   // points of the half spectrum: K0 * K1 * (K2/2 + 1)
    long long* acc = nullptr;          // [2R][glen] fixed-point charge grids (state 2: difference to state 1)
    double* real = nullptr;            // [2R][glen] charge grid, later the potential
    cufftDoubleComplex* spec = nullptr;   // [2R][slen] half spectrum
    double* mod[3] = {nullptr, nullptr, nullptr};   // B-spline moduli per dimension
    double* epart = nullptr;           // [2R][nblk] energy partials of the convolution
    int nblk = 0;
    double self_energy = 0;            // -K alpha / sqrt(pi) sum q^2
    cufftHandle plan_fwd = 0, plan_bwd = 0;
    bool have_fwd = false, have_bwd = false;
    // cuFFT entry points (dlopen)
    void* lib = nullptr;
    cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, cufftType, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal*, cufftDoubleComplex*) = nullptr;
    cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex*, cufftDoubleReal*) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
};

namespace {

// update_bsplines of ReferencePME.cpp for one fraction w: weights th[0..4] and derivatives dth[0..4]
__device__ __forceinline__ void bspline5(const double w, double* th, double* dth) {
    th[kPmeOrder - 1] = 0.0;
    th[1] = w;
    th[0] = 1.0 - w;
#pragma unroll
    for (int k = 3; k < kPmeOrder; k++) {
        const double div = 1.0 / (k - 1.0);
        th[k - 1] = div * w * th[k - 2];
#pragma unroll
        for (int l = 1; l < k - 1; l++) th[k - l - 1] = div * ((w + l) * th[k - l - 2] + (k - l - w) * th[k - l - 1]);
        th[0] = div * (1.0 - w) * th[0];
    }
    dth[0] = -th[0];
#pragma unroll
    for (int k = 1; k < kPmeOrder; k++) dth[k] = th[k - 1] - th[k];
    const double div = 1.0 / (kPmeOrder - 1);
    th[kPmeOrder - 1] = div * w * th[kPmeOrder - 2];
#pragma unroll
    for (int l = 1; l < kPmeOrder - 1; l++)
        th[kPmeOrder - l - 1] = div * ((w + l) * th[kPmeOrder - l - 2] + (kPmeOrder - l - w) * th[kPmeOrder - l - 1]);
    th[0] = div * (1.0 - w) * th[0];
}

struct PmeDims {
    int K[3];
    double box[3], inv_box[3];
};

// grid index of the first spline point and the fraction, per dimension (update_grid_index_and_fraction)
__device__ __forceinline__ void grid_coord(const PmeDims& D, const double x, const int d, int* ti, double* w) {
    double f = x * D.inv_box[d];
    f -= floor(f);
    double t = f * D.K[d];
    int i = (int)t;
    if (i >= D.K[d]) i = D.K[d] - 1;
    *ti = i;
    *w = t - i;
}

// position of atom a of replica r in state s (0: x, 1: x + d)
__device__ __forceinline__ void state_pos(const Topology& T, const double* __restrict__ pos_all, int r, int a, int s,
                                          double* x) {
    const double* p = pos_all + ((size_t)r * T.n + a) * 3;
    x[0] = p[0]; x[1] = p[1]; x[2] = p[2];
    if (s) { x[0] += T.disp[3 * a]; x[1] += T.disp[3 * a + 1]; x[2] += T.disp[3 * a + 2]; }
}

__global__ void __launch_bounds__(128)
pme_spread_kernel(Topology T, PmeDims D, int R, const double* __restrict__ pos_all, long long* __restrict__ acc,
                  size_t glen) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;          // grid = 2 * replica + state
    if (a >= T.n) return;
    const double q0 = T.q[a];
    if (q0 == 0.0) return;
    // state-1 grids take every atom; state-2 grids hold the DIFFERENCE to state 1: the displaced atoms only,
    // minus their charge where they were, plus their charge where they are now
    const bool second = (g & 1) != 0;
    if (second && T.group[a] == 0) return;
    long long* G = acc + (size_t)g * glen;
    for (int pass = 0; pass < (second ? 2 : 1); pass++) {
        const double q = (second && pass == 0) ? -q0 : q0;
        double x[3];
        state_pos(T, pos_all, g >> 1, a, second ? pass : 0, x);
        int ti[3];
        double w, th[3][kPmeOrder], dth[kPmeOrder];
        for (int d = 0; d < 3; d++) {
            grid_coord(D, x[d], d, &ti[d], &w);
            bspline5(w, th[d], dth);
        }
        for (int ix = 0; ix < kPmeOrder; ix++) {
            const int gx = (ti[0] + ix) % D.K[0];
            for (int iy = 0; iy < kPmeOrder; iy++) {
                const int gy = (ti[1] + iy) % D.K[1];
                const double qxy = q * th[0][ix] * th[1][iy];
                long long* row = G + ((size_t)gx * D.K[1] + gy) * D.K[2];
                for (int iz = 0; iz < kPmeOrder; iz++) {
                    const int gz = (ti[2] + iz) % D.K[2];
                    atomicAdd(reinterpret_cast<unsigned long long*>(row + gz),
                              (unsigned long long)__double2ll_rn(qxy * th[2][iz] * kPmeFix));
                }
            }
        }
    }
}

// fixed point -> double for both states of a replica (state 2 = state 1 + its difference grid); the
// accumulators go back to zero for the next evaluation
__global__ void pme_to_real_kernel(int R, size_t glen, long long* __restrict__ acc, double* __restrict__ real) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= glen) return;
    long long* a1 = acc + (size_t)(2 * r) * glen + i;
    long long* a2 = a1 + glen;
    const long long q1 = *a1, q2 = q1 + *a2;
    real[(size_t)(2 * r) * glen + i] = (double)q1 * (1.0 / kPmeFix);
    real[(size_t)(2 * r + 1) * glen + i] = (double)q2 * (1.0 / kPmeFix);
    *a1 = 0;
    *a2 = 0;
}

// reciprocal_convolution: grid *= eterm(m), energy partial = sum eterm |grid|^2 (before the multiplication)
__global__ void __launch_bounds__(256)
pme_convolve_kernel(PmeDims D, double alpha, size_t slen, cufftDoubleComplex* __restrict__ grid,
                    const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ mz,
                    double* __restrict__ epart, int nblk) {
    __shared__ double s_red[8];
    const int g = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < slen) {
        const int hz = D.K[2] / 2 + 1;     // half spectrum along z (real input)
        const int kz = (int)(i % hz);
        const int ky = (int)((i / hz) % D.K[1]);
        const int kx = (int)(i / ((size_t)hz * D.K[1]));
        cufftDoubleComplex* c = grid + (size_t)g * slen + i;
        if (kx == 0 && ky == 0 && kz == 0) {
            c->x = 0.0; c->y = 0.0;
        } else {
            const double mhx = (kx < (D.K[0] + 1) / 2 ? kx : kx - D.K[0]) * D.inv_box[0];
            const double mhy = (ky < (D.K[1] + 1) / 2 ? ky : ky - D.K[1]) * D.inv_box[1];
            const double mhz = kz * D.inv_box[2];
            const double m2 = mhx * mhx + mhy * mhy + mhz * mhz;
            const double pi = 3.14159265358979323846;
            const double V = D.box[0] * D.box[1] * D.box[2];
            const double denom = m2 * mx[kx] * my[ky] * mz[kz];
            const double eterm = SDM_K_COULOMB / (pi * V) * exp(-(pi * pi / (alpha * alpha)) * m2) / denom;
            const double re = c->x, im = c->y;
            // the point stands for itself and for its mirror image -kz, unless it is its own mirror image
            const double mult = (kz == 0 || 2 * kz == D.K[2]) ? 1.0 : 2.0;
            e = mult * eterm * (re * re + im * im);
            c->x = re * eterm;
            c->y = im * eterm;
        }
    }
    // fixed-order block sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if (lane == 0) s_red[warp] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; k++) s += s_red[k];
        epart[(size_t)g * nblk + blockIdx.x] = s;
    }
}

// grid_interpolate_force: F = -q sum dtheta/dr * phi; written into the external dual-state force slots
__global__ void __launch_bounds__(128)
pme_gather_kernel(Topology T, PmeDims D, int R, const double* __restrict__ pos_all,
                  const double* __restrict__ grid, size_t glen, double* __restrict__ f1, double* __restrict__ f2) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (a >= T.n) return;
    const int r = g >> 1;
    double* out = ((g & 1) ? f2 : f1) + ((size_t)r * T.n + a) * 3;
    const double q = T.q[a];
    if (q == 0.0) { out[0] = out[1] = out[2] = 0.0; return; }
    double x[3];
    state_pos(T, pos_all, r, a, g & 1, x);
    int ti[3];
    double w, th[3][kPmeOrder], dth[3][kPmeOrder];
    for (int d = 0; d < 3; d++) {
        grid_coord(D, x[d], d, &ti[d], &w);
        bspline5(w, th[d], dth[d]);
    }
    const double* G = grid + (size_t)g * glen;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int ix = 0; ix < kPmeOrder; ix++) {
        const int gx = (ti[0] + ix) % D.K[0];
        for (int iy = 0; iy < kPmeOrder; iy++) {
            const int gy = (ti[1] + iy) % D.K[1];
            const double* row = G + ((size_t)gx * D.K[1] + gy) * D.K[2];
            for (int iz = 0; iz < kPmeOrder; iz++) {
                const int gz = (ti[2] + iz) % D.K[2];
                const double v = row[gz];
                fx += dth[0][ix] * th[1][iy] * th[2][iz] * v;
                fy += th[0][ix] * dth[1][iy] * th[2][iz] * v;
                fz += th[0][ix] * th[1][iy] * dth[2][iz] * v;
            }
        }
    }
    out[0] = -q * fx * D.K[0] * D.inv_box[0];
    out[1] = -q * fy * D.K[1] * D.inv_box[1];
    out[2] = -q * fz * D.K[2] * D.inv_box[2];
}

// one warp per grid: lane-strided partial sums, then a shuffle tree -- a fixed order
__global__ void __launch_bounds__(32)
pme_finalize_kernel(int nblk, const double* __restrict__ epart, double self_energy,
                    double* __restrict__ ext_e, int* __restrict__ ext_on) {
    const int g = blockIdx.x, lane = threadIdx.x;
    double s = 0.0;
    for (int k = lane; k < nblk; k += 32) s += epart[(size_t)g * nblk + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) {
        ext_e[g] = 0.5 * s + self_energy;   // ext_e[2 r + state]
        if ((g & 1) == 0) ext_on[g >> 1] = 1;
    }
}

// pme_calculate_bsplines_moduli on the host
std::vector<double> bspline_moduli(int K) {
    double data[kPmeOrder];
    {   // M_n at the knots: update_bsplines with w = 0
        const double w = 0.0;
        data[kPmeOrder - 1] = 0.0; data[1] = w; data[0] = 1.0 - w;
        for (int k = 3; k < kPmeOrder; k++) {
            const double div = 1.0 / (k - 1.0);
            data[k - 1] = div * w * data[k - 2];
            for (int l = 1; l < k - 1; l++) data[k - l - 1] = div * ((w + l) * data[k - l - 2] + (k - l - w) * data[k - l - 1]);
            data[0] = div * (1.0 - w) * data[0];
        }
        const double div = 1.0 / (kPmeOrder - 1);
        data[kPmeOrder - 1] = div * w * data[kPmeOrder - 2];
        for (int l = 1; l < kPmeOrder - 1; l++)
            data[kPmeOrder - l - 1] = div * ((w + l) * data[kPmeOrder - l - 2] + (kPmeOrder - l - w) * data[kPmeOrder - l - 1]);
        data[0] = div * (1.0 - w) * data[0];
    }
    std::vector<double> b(K, 0.0), mod(K, 0.0);
    for (int i = 0; i < kPmeOrder && i + 1 < K; i++) b[i + 1] = data[i];
    const double two_pi = 6.28318530717958647692;
    for (int m = 0; m < K; m++) {
        double sc = 0.0, ss = 0.0;
        for (int j = 0; j < K; j++) {
            const double arg = two_pi * m * j / K;
            sc += b[j] * std::cos(arg);
            ss += b[j] * std::sin(arg);
        }
        mod[m] = sc * sc + ss * ss;
    }
    for (int m = 0; m < K; m++)
        if (mod[m] < 1e-7) mod[m] = 0.5 * (mod[(m - 1 + K) % K] + mod[(m + 1) % K]);
    return mod;
}

int fft_friendly(int n) {
    for (;; n++) {
        int m = n;
        for (int p : {2, 3, 5, 7})
            while (m % p == 0) m /= p;
        if (m == 1) return n;
    }
}

}  // namespace
}  // namespace sdm

using namespace sdm;

#define PME_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return sdm_fail(SDM_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
    } while (0)

void sdm_ctx_free_pme(sdm_ctx* c) {
    PmeState* P = c->pme;
    if (!P) return;
    if (P->have_fwd && P->Destroy) P->Destroy(P->plan_fwd);
    if (P->have_bwd && P->Destroy) P->Destroy(P->plan_bwd);
    cudaFree(P->acc); cudaFree(P->real); cudaFree(P->spec); cudaFree(P->epart);
    for (int d = 0; d < 3; d++) cudaFree(P->mod[d]);
    if (P->lib) dlclose(P->lib);
    delete P;
    c->pme = nullptr;
}

int sdm_ctx_init_pme(sdm_ctx* c, const int32_t* grid_in) {
    if (!c->T.ewald) return sdm_fail(SDM_ERR_INVALID, "reciprocal-space PME needs method SDM_PME or SDM_EWALD");
    sdm_ctx_free_pme(c);
    PmeState* P = new PmeState();
    c->pme = P;
    P->R = c->R; P->n = c->n; P->alpha = c->T.alpha;
    const double tol = c->ewald_tol > 0 ? c->ewald_tol : 5e-4;
    for (int d = 0; d < 3; d++) {
        P->box[d] = c->T.box[d];
        int k = grid_in ? grid_in[d] : 0;
        if (k <= 0) {   // NonbondedForceImpl::calcPMEParameters, rounded up to an FFT-friendly size
            k = (int)std::ceil(2.0 * P->alpha * P->box[d] / (3.0 * std::pow(tol, 0.2)));
            k = fft_friendly(k < 6 ? 6 : k);
        }
        if (k < kPmeOrder + 1) { sdm_ctx_free_pme(c); return sdm_fail(SDM_ERR_INVALID, "PME grid smaller than the spline order"); }
        P->K[d] = k;
    }
    P->glen = (size_t)P->K[0] * P->K[1] * P->K[2];
    P->slen = (size_t)P->K[0] * P->K[1] * (P->K[2] / 2 + 1);
    // cuFFT, loaded on demand
    for (const char* name : {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so"}) {
        P->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (P->lib) break;
    }
    if (!P->lib) { sdm_ctx_free_pme(c); return sdm_fail(SDM_ERR_CUDA, "cuFFT (libcufft.so.11) not found: reciprocal-space PME is unavailable"); }
    P->PlanMany = (decltype(P->PlanMany))dlsym(P->lib, "cufftPlanMany");
    P->SetStream = (decltype(P->SetStream))dlsym(P->lib, "cufftSetStream");
    P->ExecD2Z = (decltype(P->ExecD2Z))dlsym(P->lib, "cufftExecD2Z");
    P->ExecZ2D = (decltype(P->ExecZ2D))dlsym(P->lib, "cufftExecZ2D");
    P->Destroy = (decltype(P->Destroy))dlsym(P->lib, "cufftDestroy");
    if (!P->PlanMany || !P->SetStream || !P->ExecD2Z || !P->ExecZ2D || !P->Destroy) {
        sdm_ctx_free_pme(c);
        return sdm_fail(SDM_ERR_CUDA, "cuFFT entry points missing");
    }
    const size_t total = 2 * (size_t)P->R * P->glen;
    PME_CUDA(cudaMalloc(&P->acc, sizeof(long long) * total));
    PME_CUDA(cudaMemset(P->acc, 0, sizeof(long long) * total));
    PME_CUDA(cudaMalloc(&P->real, sizeof(double) * total));
    PME_CUDA(cudaMalloc(&P->spec, sizeof(cufftDoubleComplex) * 2 * (size_t)P->R * P->slen));
    P->nblk = (int)((P->slen + 255) / 256);
    PME_CUDA(cudaMalloc(&P->epart, sizeof(double) * 2 * (size_t)P->R * P->nblk));
    for (int d = 0; d < 3; d++) {
        const std::vector<double> m = bspline_moduli(P->K[d]);
        PME_CUDA(cudaMalloc(&P->mod[d], sizeof(double) * m.size()));
        PME_CUDA(cudaMemcpy(P->mod[d], m.data(), sizeof(double) * m.size(), cudaMemcpyHostToDevice));
    }
    int dims[3] = {P->K[0], P->K[1], P->K[2]};
    if (P->PlanMany(&P->plan_fwd, 3, dims, nullptr, 1, (int)P->glen, nullptr, 1, (int)P->slen, CUFFT_D2Z, 2 * P->R) != CUFFT_SUCCESS) {
        sdm_ctx_free_pme(c);
        return sdm_fail(SDM_ERR_CUDA, "cufftPlanMany (D2Z) failed");
    }
    P->have_fwd = true;
    if (P->PlanMany(&P->plan_bwd, 3, dims, nullptr, 1, (int)P->slen, nullptr, 1, (int)P->glen, CUFFT_Z2D, 2 * P->R) != CUFFT_SUCCESS) {
        sdm_ctx_free_pme(c);
        return sdm_fail(SDM_ERR_CUDA, "cufftPlanMany (Z2D) failed");
    }
    P->have_bwd = true;
    double q2 = 0.0;
    for (double q : c->h_charge) q2 += q * q;
    P->self_energy = -SDM_K_COULOMB * P->alpha / std::sqrt(3.14159265358979323846) * q2;
    return SDM_OK;
}

// The reciprocal-space pass of one evaluation, on stream s (needs the positions only).
int sdm_ctx_pme_enqueue(sdm_ctx* c, cudaStream_t s) {
    PmeState* P = c->pme;
    if (!P) return SDM_OK;
    PmeDims D;
    for (int d = 0; d < 3; d++) { D.K[d] = P->K[d]; D.box[d] = P->box[d]; D.inv_box[d] = 1.0 / P->box[d]; }
    const int R = P->R, n = P->n;
    dim3 ga((n + 127) / 128, 2 * R);
    pme_spread_kernel<<<ga, 128, 0, s>>>(c->T, D, R, c->d_pos, P->acc, P->glen);
    dim3 gr((unsigned)((P->glen + 255) / 256), R);
    pme_to_real_kernel<<<gr, 256, 0, s>>>(R, P->glen, P->acc, P->real);
    if (P->SetStream(P->plan_fwd, s) != CUFFT_SUCCESS || P->SetStream(P->plan_bwd, s) != CUFFT_SUCCESS)
        return sdm_fail(SDM_ERR_CUDA, "cufftSetStream failed");
    if (P->ExecD2Z(P->plan_fwd, P->real, P->spec) != CUFFT_SUCCESS) return sdm_fail(SDM_ERR_CUDA, "cufftExecD2Z failed");
    dim3 gc(P->nblk, 2 * R);
    pme_convolve_kernel<<<gc, 256, 0, s>>>(D, P->alpha, P->slen, P->spec, P->mod[0], P->mod[1], P->mod[2], P->epart, P->nblk);
    if (P->ExecZ2D(P->plan_bwd, P->spec, P->real) != CUFFT_SUCCESS) return sdm_fail(SDM_ERR_CUDA, "cufftExecZ2D failed");
    pme_gather_kernel<<<ga, 128, 0, s>>>(c->T, D, R, c->d_pos, P->real, P->glen, c->d_ext_f1, c->d_ext_f2);
    pme_finalize_kernel<<<2 * R, 32, 0, s>>>(P->nblk, P->epart, P->self_energy, c->d_ext_e, c->d_ext_on);
    c->launches += 5;   // our kernels; the two transforms are cuFFT's
    PME_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_ctx_pme_info(sdm_ctx* c, const char* key, double* value) {
    if (!c->pme) return SDM_ERR_INVALID;
    const std::string k(key);
    if (k == "pme_grid_x") *value = c->pme->K[0];
    else if (k == "pme_grid_y") *value = c->pme->K[1];
    else if (k == "pme_grid_z") *value = c->pme->K[2];
    else if (k == "pme_self_energy") *value = c->pme->self_energy;
    else return SDM_ERR_INVALID;
    return SDM_OK;
}
