// kernels_gb.cu -- HCT generalized Born + ACE surface area for both states of every resident replica
// (the implicit-solvent model example/desmonddmsfile75.py:454-465 puts in force group 2:
// GBSAHCTForce(SA='ACE'), so LangevinIntegratorSDM::step evaluates it at x and at x + d).
//
// Restated from the expression strings of OpenMM 7.3's app/internal/customgbforces.py (GBSAHCTForce and
// _createEnergyTerms with cutoff=None, kappa=0) under CustomGBForce's rules: NoCutoff (plain distances), no
// exclusions, step(x) = 0 for x < 0 else 1, chain rule through the computed values:
//     I_i = sum_{j != i} H(r_ij; or_i, sr_j),   B_i = 1/(1/or_i - I_i)
//     E   = sum_i [-0.5 Kp q_i^2 / B_i + 28.3919551 (rad_i + 0.14)^2 (rad_i/B_i)^6] + sum_{i<j} -Kp q_i q_j / f_ij
//     f   = sqrt(r^2 + B_i B_j exp(-r^2 / (4 B_i B_j))),  Kp = 138.935485 (1/eps_solute - 1/eps_solvent)
//
// The Born radii are global in the coordinates (every atom's radius changes when the ligand is displaced), so the
// moved-pairs-only trick of the pair path does not apply: state 1 and state 2 each get the full three passes
//     born (I, B)  ->  pair (energy, direct force, dE/dB)  ->  chain (force through the Born radii)
// as 2R independent systems in one launch each.  FP64 throughout; one warp owns one atom of one system and walks
// its whole row, so every force is a fixed-order sum without atomics (bit-reproducible; each pair is visited from
// both sides, which is what an implicit-solvent system of a few hundred to a few thousand atoms affords).  The
// results land in the external dual-state slots like reciprocal-space PME: E1 += E_gb(x), u += E_gb(x+d) - E_gb(x),
// F1 += F_gb(x), F2 - F1 += F_gb(x+d) - F_gb(x).
#include <cmath>
#include <string>
#include <vector>

#include "sdm_ctx.h"
#include "sdm_internal.cuh"

namespace sdm {

constexpr double kGbCoulomb = 138.935485;   // the constant of customgbforces.py (not ONE_4PI_EPS0)
constexpr double kAceCoeff = 28.3919551;
constexpr double kAceProbe = 0.14;
constexpr double kHctOffset = 0.009;

struct GbState {
    int R = 0, n = 0;
    double kp = 0;                 // 138.935485 (1/solute - 1/solvent)
    int sa_ace = 1;
    double* par = nullptr;         // [n][4] charge, or, sr, ACE prefactor 28.39 (rad+0.14)^2 rad^6 (0 without SA)
    double* xs = nullptr;          // [2R][n][4] positions of system g = 2 r + state, .w = charge
    double* born = nullptr;        // [2R][n][2] B and 1/B
    double* dEdI = nullptr;        // [2R][n] dE/dI
    double* eatom = nullptr;       // [2R][n] energy booked on the atom
};

namespace {

constexpr int kGbWarps = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// H(r; or_i, sr_j) of the computed value I and its derivative with respect to r; ir = 1/r.  One division: 1/L and
// 1/U come from the reciprocal of their product.
__device__ __forceinline__ void hct_term(const double r, const double ir, const double o, const double s, double* h,
                                         double* dh) {
    if (r + s - o < 0.0) { *h = 0.0; *dh = 0.0; return; }
    const double U = r + s, D = fabs(r - s);
    const bool lo = o >= D;                         // L = max(or, D)
    const double L = lo ? o : D;
    const double dL = lo ? 0.0 : (r >= s ? 1.0 : -1.0);
    const double iLU = 1.0 / (L * U);
    const double iL = iLU * U, iU = iLU * L;
    const double iL2 = iL * iL, iU2 = iU * iU;
    const double lg = log(L * iU);
    const double s2ir = s * s * ir;
    const double a = r - s2ir;
    *h = 0.5 * (iL - iU + 0.25 * a * (iU2 - iL2) + 0.5 * lg * ir);
    *dh = 0.5 * (-dL * iL2 + iU2 + 0.25 * (1.0 + s2ir * ir) * (iU2 - iL2) +
                 0.5 * a * (dL * iL2 * iL - iU2 * iU) + 0.5 * ((dL * iL - iU) * ir - lg * ir * ir));
}

__global__ void gb_prep_kernel(Topology T, int R, const double* __restrict__ pos_all, const double* __restrict__ par,
                               double* __restrict__ xs) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (a >= T.n) return;
    const int r = g >> 1, st = g & 1;
    const double* p = pos_all + ((size_t)r * T.n + a) * 3;
    double x = p[0], y = p[1], z = p[2];
    if (st) { x += T.disp[3 * a]; y += T.disp[3 * a + 1]; z += T.disp[3 * a + 2]; }
    double* o = xs + ((size_t)g * T.n + a) * 4;
    o[0] = x; o[1] = y; o[2] = z; o[3] = par[4 * a];
}

// Born radii: one warp per atom of one system.
__global__ void __launch_bounds__(32 * kGbWarps)
gb_born_kernel(int n, const double* __restrict__ par, const double* __restrict__ xs, double* __restrict__ born) {
    const int i = blockIdx.x * kGbWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31, g = blockIdx.y;
    if (i >= n) return;
    const double* X = xs + (size_t)g * n * 4;
    const double xi = X[4 * i], yi = X[4 * i + 1], zi = X[4 * i + 2], oi = par[4 * i + 1];
    double I = 0.0;
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double dx = xi - X[4 * j], dy = yi - X[4 * j + 1], dz = zi - X[4 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double ir = rsqrt(r2), r = r2 * ir;
        double h, dh;
        hct_term(r, ir, oi, par[4 * j + 2], &h, &dh);
        I += h;
    }
    I = warp_sum(I);
    if (lane == 0) {
        const double ib = 1.0 / oi - I;
        reinterpret_cast<double2*>(born)[(size_t)g * n + i] = make_double2(1.0 / ib, ib);
    }
}

// Pair term at fixed Born radii: energy, force on i, dE/dB_i -> dE/dI_i; single-particle terms of atom i.
__global__ void __launch_bounds__(32 * kGbWarps)
gb_pair_kernel(int n, double kp, const double* __restrict__ par, const double* __restrict__ xs,
               const double* __restrict__ born, double* __restrict__ dEdI, double* __restrict__ eatom,
               double* __restrict__ f_state1, double* __restrict__ f_state2) {
    const int i = blockIdx.x * kGbWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31, g = blockIdx.y;
    if (i >= n) return;
    const double* X = xs + (size_t)g * n * 4;
    const double2* Bv = reinterpret_cast<const double2*>(born) + (size_t)g * n;   // (B, 1/B)
    const double xi = X[4 * i], yi = X[4 * i + 1], zi = X[4 * i + 2], qi = X[4 * i + 3], Bi = Bv[i].x;
    const double qiB = 0.25 * Bv[i].y;
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0, dB = 0.0;
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double dx = xi - X[4 * j], dy = yi - X[4 * j + 1], dz = zi - X[4 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double2 bj = Bv[j];
        const double Bj = bj.x, bb = Bi * Bj;
        const double w = r2 * (qiB * bj.y);                   // r^2 / (4 B_i B_j)
        const double ex = exp(-w);
        const double f2 = r2 + bb * ex;
        const double inv_f = rsqrt(f2);
        const double c = kp * qi * X[4 * j + 3] * inv_f;      // -E_ij
        e -= c;
        const double c3 = c * inv_f * inv_f;                  // Kp qi qj / f^3
        const double fr = -c3 * (1.0 - 0.25 * ex);            // force on i = fr * (xi - xj)
        fx += fr * dx; fy += fr * dy; fz += fr * dz;
        dB += 0.5 * c3 * Bj * ex * (1.0 + w);
    }
    e = warp_sum(e); fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz); dB = warp_sum(dB);
    if (lane == 0) {
        const double self = -0.5 * kp * qi * qi / Bi;
        const double iB = 1.0 / Bi, iB3 = iB * iB * iB;
        const double ace = par[4 * i + 3] * iB3 * iB3;
        dB += -self * iB - 6.0 * ace * iB;
        dEdI[(size_t)g * n + i] = dB * Bi * Bi;                // dB/dI = B^2
        eatom[(size_t)g * n + i] = 0.5 * e + self + ace;
        double* f = ((g & 1) ? f_state2 : f_state1) + ((size_t)(g >> 1) * n + i) * 3;   // system g = 2 r + state
        f[0] = fx; f[1] = fy; f[2] = fz;
    }
}

// Chain rule through the Born radii: F_i -= sum_j [dE/dI_i H'(r; or_i, sr_j) + dE/dI_j H'(r; or_j, sr_i)] (x_i - x_j)/r
__global__ void __launch_bounds__(32 * kGbWarps)
gb_chain_kernel(int n, const double* __restrict__ par, const double* __restrict__ xs, const double* __restrict__ dEdI,
                double* __restrict__ f_state1, double* __restrict__ f_state2) {
    const int i = blockIdx.x * kGbWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31, g = blockIdx.y;
    if (i >= n) return;
    const double* X = xs + (size_t)g * n * 4;
    const double* G = dEdI + (size_t)g * n;
    const double xi = X[4 * i], yi = X[4 * i + 1], zi = X[4 * i + 2], oi = par[4 * i + 1], si = par[4 * i + 2], gi = G[i];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double dx = xi - X[4 * j], dy = yi - X[4 * j + 1], dz = zi - X[4 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double ir = rsqrt(r2), r = r2 * ir;
        double h, dh_ij, dh_ji;
        hct_term(r, ir, oi, par[4 * j + 2], &h, &dh_ij);
        hct_term(r, ir, par[4 * j + 1], si, &h, &dh_ji);
        const double fr = -(gi * dh_ij + G[j] * dh_ji) * ir;
        fx += fr * dx; fy += fr * dy; fz += fr * dz;
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
        double* f = ((g & 1) ? f_state2 : f_state1) + ((size_t)(g >> 1) * n + i) * 3;
        f[0] += fx; f[1] += fy; f[2] += fz;
    }
}

// Energy of every system in a fixed order; switches the external slots of the replica on.
__global__ void __launch_bounds__(32)
gb_finalize_kernel(int n, const double* __restrict__ eatom, double* __restrict__ ext_e, int* __restrict__ ext_on) {
    const int g = blockIdx.x, lane = threadIdx.x;
    double s = 0.0;
    for (int k = lane; k < n; k += 32) s += eatom[(size_t)g * n + k];
    s = warp_sum(s);
    if (lane == 0) {
        ext_e[g] = s;                        // ext_e[2 r + state]
        if ((g & 1) == 0) ext_on[g >> 1] = 1;
    }
}

}  // namespace
}  // namespace sdm

using namespace sdm;

#define GB_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return sdm_fail(SDM_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
    } while (0)

void sdm_ctx_free_gb(sdm_ctx* c) {
    GbState* G = c->gb;
    if (!G) return;
    cudaFree(G->par); cudaFree(G->xs); cudaFree(G->born); cudaFree(G->dEdI); cudaFree(G->eatom);
    delete G;
    c->gb = nullptr;
}

int sdm_ctx_init_gb(sdm_ctx* c, const double* charge, const double* offset_radius, const double* scaled_radius,
                    double solute_dielectric, double solvent_dielectric, int sa_ace) {
    sdm_ctx_free_gb(c);
    GbState* G = new GbState();
    c->gb = G;
    G->R = c->R; G->n = c->n; G->sa_ace = sa_ace ? 1 : 0;
    G->kp = kGbCoulomb * (1.0 / solute_dielectric - 1.0 / solvent_dielectric);
    std::vector<double> par(4 * (size_t)c->n);
    for (int a = 0; a < c->n; a++) {
        const double rad = offset_radius[a] + kHctOffset;
        par[4 * a] = charge ? charge[a] : c->h_charge[a];
        par[4 * a + 1] = offset_radius[a];
        par[4 * a + 2] = scaled_radius[a];
        par[4 * a + 3] = sa_ace ? kAceCoeff * (rad + kAceProbe) * (rad + kAceProbe) * std::pow(rad, 6) : 0.0;
    }
    const size_t sys = 2 * (size_t)G->R * G->n;
    GB_CUDA(cudaMalloc(&G->par, sizeof(double) * par.size()));
    GB_CUDA(cudaMemcpy(G->par, par.data(), sizeof(double) * par.size(), cudaMemcpyHostToDevice));
    GB_CUDA(cudaMalloc(&G->xs, sizeof(double) * 4 * sys));
    GB_CUDA(cudaMalloc(&G->born, sizeof(double) * 2 * sys));
    GB_CUDA(cudaMemset(G->born, 0, sizeof(double) * 2 * sys));   // sdm_get_born_radii before the first evaluation: zeros
    GB_CUDA(cudaMalloc(&G->dEdI, sizeof(double) * sys));
    GB_CUDA(cudaMalloc(&G->eatom, sizeof(double) * sys));
    return SDM_OK;
}

// The GB pass of one evaluation, on stream s (needs the positions only).
int sdm_ctx_gb_enqueue(sdm_ctx* c, cudaStream_t s) {
    GbState* G = c->gb;
    if (!G) return SDM_OK;
    const int R = G->R, n = G->n;
    gb_prep_kernel<<<dim3((n + 127) / 128, 2 * R), 128, 0, s>>>(c->T, R, c->d_pos, G->par, G->xs);
    const dim3 grid((n + kGbWarps - 1) / kGbWarps, 2 * R);
    gb_born_kernel<<<grid, 32 * kGbWarps, 0, s>>>(n, G->par, G->xs, G->born);
    gb_pair_kernel<<<grid, 32 * kGbWarps, 0, s>>>(n, G->kp, G->par, G->xs, G->born, G->dEdI, G->eatom, c->d_ext_f1, c->d_ext_f2);
    gb_chain_kernel<<<grid, 32 * kGbWarps, 0, s>>>(n, G->par, G->xs, G->dEdI, c->d_ext_f1, c->d_ext_f2);
    gb_finalize_kernel<<<2 * R, 32, 0, s>>>(n, G->eatom, c->d_ext_e, c->d_ext_on);
    c->launches += 5;
    GB_CUDA(cudaGetLastError());
    return SDM_OK;
}

int sdm_ctx_gb_info(sdm_ctx* c, const char* key, double* value) {
    if (!c->gb) return SDM_ERR_INVALID;
    const std::string k(key);
    if (k == "gb_prefactor") *value = c->gb->kp;
    else if (k == "gb_sa_ace") *value = c->gb->sa_ace;
    else return SDM_ERR_INVALID;
    return SDM_OK;
}

// Born radii of one system (replica, state) of the last evaluation; synchronises (tests, diagnostics).
int sdm_ctx_gb_born_radii(sdm_ctx* c, int replica, int state, double* out, cudaStream_t s) {
    GbState* G = c->gb;
    if (!G) return sdm_fail(SDM_ERR_INVALID, "HCT-GB is not switched on");
    GB_CUDA(cudaMemcpy2DAsync(out, sizeof(double), G->born + 2 * (size_t)(2 * replica + state) * G->n, 2 * sizeof(double),
                              sizeof(double), G->n, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return SDM_OK;
}
