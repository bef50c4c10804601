// kernels_fused.cu -- stages of the fused dual-state evaluation that do not depend on the
// neighbour-list machinery: position prep, the all-pairs tile kernel (small systems and
// cross-check), the displaced-atom ("ligand") dual-state kernels, 1-4 exceptions, the device
// scalar stage (soft-core + bias) and the hybrid-force mix.  sm_100a.
//
// Reference behaviour restated (never copied): LangevinIntegratorSDM.cpp:153-183 (sequence),
// ReferenceSDMKernels.cpp:161-199 (state copies), :202-318 (execute), OpenMM 7.3 Reference
// NonbondedForce arithmetic.
#include <algorithm>
#include <cstdlib>

#include "sdm_kernels.h"

namespace sdm {

namespace {

constexpr int kTile = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-order block reduction (deterministic): warp trees, then warp 0 adds the warp totals in
// order.  Result valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
__device__ double block_sum(double v, double* smem /* >= 32 */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < nw; k++) t += smem[k];
    return t;
}

__device__ long long block_sum_ll(long long v, long long* smem) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_ll(v);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    long long t = 0;
    if (threadIdx.x == 0)
        for (int k = 0; k < nw; k++) t += smem[k];
    return t;
}

// ---------------------------------------------------------------------------------------------
// prep: double positions -> float4 (wrapped into the box when periodic; .w = q*sqrt(K)).
// ---------------------------------------------------------------------------------------------
__global__ void prep_posq_kernel(Topology T, const double* __restrict__ pos, float4* __restrict__ posq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (i >= T.n) return;
    const double* p = pos + (size_t)r * 3 * T.n + 3 * (size_t)i;
    double x = p[0], y = p[1], z = p[2];
    if (T.method == SDM_CUTOFF_PERIODIC) {
        x -= floor(x * T.inv_box[0]) * T.box[0];
        y -= floor(y * T.inv_box[1]) * T.box[1];
        z -= floor(z * T.inv_box[2]) * T.box[2];
    }
    posq[(size_t)r * T.n + i] = make_float4((float)x, (float)y, (float)z, T.parf[i].x);
}

// ---------------------------------------------------------------------------------------------
// all-pairs tile kernel: thread i walks every j (tiles of 128 staged in shared memory), keeps
// the force on i only, so every pair is visited from both sides (energy and counts are halved
// by the scalar stage).  No atomics, deterministic.  FP32 pair arithmetic; per-tile partial sums
// are flushed into FP64 accumulators.
// ---------------------------------------------------------------------------------------------
template <int METHOD>
__global__ void __launch_bounds__(kTile)
allpairs_kernel(Topology T, const float4* __restrict__ posq_all, const double* __restrict__ pos_all,
                long long* __restrict__ f1acc, int nslot, double* __restrict__ epart,
                long long* __restrict__ cpart, int n_epart, int exact, int* emit_counter,
                int* emit_pairs, int emit_cap, int emit_replica) {
    __shared__ float4 s_posq[kTile];
    __shared__ float2 s_par[kTile];
    __shared__ double s_red[32];
    __shared__ long long s_redl[32];

    const int n = T.n;
    const int r = blockIdx.y;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * kTile + tid;
    const bool valid = i < n;
    const float4* posq = posq_all + (size_t)r * n;
    const double* pos = pos_all + (size_t)r * 3 * n;
    const bool emit = emit_pairs != nullptr && r == emit_replica;

    float4 pi = valid ? posq[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pari = valid ? T.parf[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    int ep = valid ? T.excl_start[i] : 0;
    const int eend = valid ? T.excl_start[i + 1] : 0;
    int enext = (ep < eend) ? T.excl_idx[ep] : 0x7fffffff;

    double fx = 0.0, fy = 0.0, fz = 0.0, en = 0.0;
    long long cnt = 0;

    for (int jt = 0; jt < n; jt += kTile) {
        int jl = jt + tid;
        if (jl < n) {
            s_posq[tid] = posq[jl];
            float4 pj = T.parf[jl];
            s_par[tid] = make_float2(pj.y, pj.z);
        }
        __syncthreads();
        const int jmax = min(kTile, n - jt);
        float tfx = 0.f, tfy = 0.f, tfz = 0.f, te = 0.f;
        if (valid) {
            for (int jj = 0; jj < jmax; jj++) {
                const int j = jt + jj;
                if (j == enext) {
                    ep++;
                    enext = (ep < eend) ? T.excl_idx[ep] : 0x7fffffff;
                    continue;
                }
                if (j == i) continue;
                const float4 pj = s_posq[jj];
                float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                if (METHOD == SDM_CUTOFF_PERIODIC) {
                    dx -= T.boxf[0] * rintf(dx * T.inv_boxf[0]);
                    dy -= T.boxf[1] * rintf(dy * T.inv_boxf[1]);
                    dz -= T.boxf[2] * rintf(dz * T.inv_boxf[2]);
                }
                const float r2 = dx * dx + dy * dy + dz * dz;
                bool in = true;
                if (METHOD != SDM_NOCUTOFF) {
                    in = r2 <= T.rc2f;
                    if (exact && fabsf(r2 - T.rc2f) < T.band) in = in_cutoff_f64(T, pos, i, j);
                }
                if (!in) continue;
                const float2 parj = s_par[jj];
                const float rinv = rsqrtf(r2);
                const float rinv2 = rinv * rinv;
                const float sig = pari.y + parj.x;
                const float sr2 = (T.lj_geom ? pari.y * parj.x : sig * sig) * rinv2;   // geometric rule: sigma_i sigma_j
                const float sr6 = sr2 * sr2 * sr2;
                const float eps = pari.z * parj.y;
                const float qq = pi.w * pj.w;
                float dEdR = eps * (12.f * sr6 - 6.f) * sr6;
                float e = eps * (sr6 - 1.f) * sr6;
                if (METHOD == SDM_CUTOFF_PERIODIC && T.ewald) {   // direct-space Ewald (A&S 7.1.26 erfc, like the row kernel)
                    const float ar = T.alphaf * (r2 * rinv);
                    const float expo = __expf(-ar * ar);
                    const float t = 1.f / (1.f + 0.3275911f * ar);
                    const float ec = ((((1.061405429f * t - 1.453152027f) * t + 1.421413741f) * t - 0.284496736f) * t + 0.254829592f) * t * expo;
                    dEdR += qq * rinv * (ec + 1.1283791670955126f * ar * expo);
                    e += qq * rinv * ec;
                } else if (METHOD != SDM_NOCUTOFF) {
                    dEdR += qq * (rinv - 2.f * T.krff * r2);
                    e += qq * (rinv + T.krff * r2 - T.crff);
                } else {
                    dEdR += qq * rinv;
                    e += qq * rinv;
                }
                const float fs = dEdR * rinv2;
                tfx += fs * dx;
                tfy += fs * dy;
                tfz += fs * dz;
                te += e;
                cnt++;
                if (emit && i < j) {
                    int slot = atomicAdd(emit_counter, 1);
                    if (slot < emit_cap) {
                        emit_pairs[2 * slot] = i;
                        emit_pairs[2 * slot + 1] = j;
                    }
                }
            }
        }
        fx += tfx; fy += tfy; fz += tfz; en += te;
        __syncthreads();
    }
    if (valid) {
        long long* acc = f1acc + (size_t)r * 3 * nslot;  // all-pairs: acc_rstride == 3*nslot
        acc[i] = to_fixed(fx);
        acc[nslot + i] = to_fixed(fy);
        acc[2 * nslot + i] = to_fixed(fz);
    }
    double be = block_sum(en, s_red);
    long long bc = block_sum_ll(cnt, s_redl);
    if (tid == 0) {
        epart[(size_t)r * n_epart + blockIdx.x] = be;
        cpart[(size_t)r * n_epart + blockIdx.x] = bc;
    }
}

// ---------------------------------------------------------------------------------------------
// Displaced-atom kernels.  State 2 = x + d for EVERY atom (ReferenceSDMKernels.cpp:192-199); a
// pair changes iff the two displacement vectors differ, so u = E2 - E1 and dF = F2 - F1 only need
// the pairs between a displaced atom and an atom of another displacement group.
//
// Both kernels walk a spatially ordered scan list (EvalBuffers::scan_*) with a cheap FP32
// minimum-image prefilter (conservative: margin >> FP32 rounding) and evaluate the ~2% of
// survivors in FP64, positions straight from the double buffer, in-cutoff decisions with the
// contraction-free expression of the oracle.  Warps scan consecutive slots of one cell, so the
// FP64 branch is taken by whole warps near the displaced atoms and skipped elsewhere.  No
// atomics: every output element has one owner and a fixed summation order (bit-reproducible).
//
// probe kernel: one block per (displaced atom i, replica); walks all atoms k whose displacement
// differs from i's, evaluates the pair at state 1 and state 2 and reduces
//     dF_i = sum_k f_i(state 2) - f_i(state 1),   u_i = sum_k w_k (e2 - e1),
// w_k = 1/2 when k is displaced too (that pair is seen from k's block as well), else 1.
// ---------------------------------------------------------------------------------------------
struct PairGeom {
    double dx, dy, dz, r2;
};

__device__ __forceinline__ PairGeom geom(const Topology& T, double xi, double yi, double zi,
                                         double xk, double yk, double zk) {
    PairGeom g;
    g.dx = xi - xk; g.dy = yi - yk; g.dz = zi - zk;
    if (T.method == SDM_CUTOFF_PERIODIC) {
        g.dx = min_image_fast(g.dx, T.box[0], T.inv_box[0]);
        g.dy = min_image_fast(g.dy, T.box[1], T.inv_box[1]);
        g.dz = min_image_fast(g.dz, T.box[2], T.inv_box[2]);
    }
    g.r2 = norm2_exact(g.dx, g.dy, g.dz);
    return g;
}

// FP32 minimum-image distance^2 for the prefilter.  Both points lie in [0, L) when periodic
// (scan positions are wrapped at list-build / prep time, probes are wrapped when staged, and a
// list reuse lets atoms drift by less than the skin), so |d| < 2L and one conditional shift per
// dimension selects the nearest image -- no FRND.
__device__ __forceinline__ float wrap1(float d, float L, float hL) {
    d = d > hL ? d - L : d;
    return d < -hL ? d + L : d;
}

__device__ __forceinline__ float r2_prefilter(const Topology& T, const float3 hbox, float xi, float yi,
                                              float zi, const float4 p) {
    float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
    if (T.method == SDM_CUTOFF_PERIODIC) {
        dx = wrap1(dx, T.boxf[0], hbox.x);
        dy = wrap1(dy, T.boxf[1], hbox.y);
        dz = wrap1(dz, T.boxf[2], hbox.z);
    }
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ float wrap_into_box(double x, double L, double invL) {
    return (float)(x - floor(x * invL) * L);
}

__device__ __forceinline__ void scan_range(const Topology& T, const EvalBuffers& B, int r, int* begin,
                                           int* end) {
    if (B.scan_off) {
        *begin = B.scan_off[(size_t)r * B.scan_stride];
        *end = B.scan_off[(size_t)(r + 1) * B.scan_stride];
    } else {
        *begin = r * T.n;
        *end = (r + 1) * T.n;
    }
}

// ---- filter kernel -------------------------------------------------------------------------------
// One thread per scan index (a NON-displaced atom j).  FP32 only: tests j against the displaced
// atoms (staged in shared memory in groups of 8 with a bounding sphere per state) and writes the
// prefilter bitmap hitbits[r][m][w] (bit = lane) -- every word of every row, zero when nothing is
// near.  No forces are computed here.
struct ProbeF {
    float fx1, fy1, fz1, fx2, fy2, fz2;
};

constexpr int kFilterThreads = 128;
constexpr int kProbeChunk = 64;   // displaced atoms staged per pass
constexpr int kProbeGroup = 8;    // displaced atoms per bounding sphere

__global__ void __launch_bounds__(kFilterThreads)
ligand_filter_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    __shared__ ProbeF s_p[kProbeChunk];
    __shared__ float4 s_sph[kProbeChunk / kProbeGroup][2];   // (center, (radius + r_lim)^2) per state
    const int n = T.n, r = blockIdx.y;
    const double* pos = B.pos + (size_t)r * 3 * n;
    int begin, end;
    scan_range(T, B, r, &begin, &end);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sw = blockIdx.x * (kFilterThreads / 32) + warp;   // scan word (32 scan indices) of this warp
    const int idx = begin + sw * 32 + lane;
    int j = -1;
    if (idx < end) {
        j = idx - begin;
        if (B.scan_atom) {
            const int ga = B.scan_atom[idx];
            j = ga < 0 ? -1 : ga - r * n;
        }
    }
    const bool active = j >= 0 && T.group[j] == 0;
    const bool cutoff = T.method != SDM_NOCUTOFF;
    const bool periodic = T.method == SDM_CUTOFF_PERIODIC;
    const float rl = sqrtf(T.rc2f) + B.filter_skin;   // cluster path: + skin, the bitmap outlives this eval
    const float lim = rl * rl * 1.0001f + 1.0e-4f;
    const float rlim = sqrtf(lim);
    const float3 hbox = make_float3(0.5f * T.boxf[0], 0.5f * T.boxf[1], 0.5f * T.boxf[2]);
    const float4 pf = active ? B.scan_posq[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    // layout [replica][scan word][displaced atom]: the words of one warp are contiguous
    uint32_t* bits = B.hitbits + ((size_t)r * B.scan_words + sw) * T.n_lig;
    const bool store_bits = sw < B.scan_words;
    for (int m0 = 0; m0 < T.n_lig; m0 += kProbeChunk) {
        const int mc = min(kProbeChunk, T.n_lig - m0);
        __syncthreads();
        if (threadIdx.x < mc) {
            const int i = T.lig_idx[m0 + threadIdx.x];
            const double x1 = pos[3 * i], y1 = pos[3 * i + 1], z1 = pos[3 * i + 2];
            const double x2 = x1 + T.disp[3 * i], y2 = y1 + T.disp[3 * i + 1], z2 = z1 + T.disp[3 * i + 2];
            ProbeF p;
            if (periodic) {
                p.fx1 = wrap_into_box(x1, T.box[0], T.inv_box[0]); p.fx2 = wrap_into_box(x2, T.box[0], T.inv_box[0]);
                p.fy1 = wrap_into_box(y1, T.box[1], T.inv_box[1]); p.fy2 = wrap_into_box(y2, T.box[1], T.inv_box[1]);
                p.fz1 = wrap_into_box(z1, T.box[2], T.inv_box[2]); p.fz2 = wrap_into_box(z2, T.box[2], T.inv_box[2]);
            } else {
                p.fx1 = (float)x1; p.fy1 = (float)y1; p.fz1 = (float)z1;
                p.fx2 = (float)x2; p.fy2 = (float)y2; p.fz2 = (float)z2;
            }
            s_p[threadIdx.x] = p;
        }
        __syncthreads();
        const int ngroups = (mc + kProbeGroup - 1) / kProbeGroup;
        if (threadIdx.x < 2 * ngroups) {
            // bounding sphere of one group of displaced atoms in one state, centred on its first atom
            const int g = threadIdx.x >> 1, st = threadIdx.x & 1;
            const ProbeF& c0 = s_p[g * kProbeGroup];
            const float cx = st ? c0.fx2 : c0.fx1, cy = st ? c0.fy2 : c0.fy1, cz = st ? c0.fz2 : c0.fz1;
            float rmax2 = 0.f;
            for (int k = g * kProbeGroup + 1; k < min(mc, (g + 1) * kProbeGroup); k++) {
                const ProbeF& q = s_p[k];
                const float4 o = make_float4(st ? q.fx2 : q.fx1, st ? q.fy2 : q.fy1, st ? q.fz2 : q.fz1, 0.f);
                rmax2 = fmaxf(rmax2, r2_prefilter(T, hbox, cx, cy, cz, o));
            }
            const float rr = sqrtf(rmax2) * 1.0001f + rlim + 1.0e-4f;
            s_sph[g][st] = make_float4(cx, cy, cz, rr * rr);
        }
        __syncthreads();
        for (int g = 0; g < ngroups; g++) {
            const float4 s1 = s_sph[g][0], s2 = s_sph[g][1];
            const bool near = active && (!cutoff || r2_prefilter(T, hbox, s1.x, s1.y, s1.z, pf) <= s1.w ||
                                         r2_prefilter(T, hbox, s2.x, s2.y, s2.z, pf) <= s2.w);
            const int mend = min(mc, (g + 1) * kProbeGroup);
            if (!__any_sync(0xffffffffu, near)) {
                if (store_bits && lane < mend - g * kProbeGroup) bits[m0 + g * kProbeGroup + lane] = 0u;
                continue;
            }
            for (int m = g * kProbeGroup; m < mend; m++) {
                const ProbeF& p = s_p[m];
                const bool hit = near && (!cutoff || r2_prefilter(T, hbox, p.fx1, p.fy1, p.fz1, pf) <= lim ||
                                          r2_prefilter(T, hbox, p.fx2, p.fy2, p.fz2, pf) <= lim);
                const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                if (store_bits && lane == 0) bits[m0 + m] = ballot;
            }
        }
    }
}

// ---- probe kernel -------------------------------------------------------------------------------
// One block per (displaced atom i, replica).  Reads row (r, m) of the prefilter bitmap, spreads
// its set bits densely over the threads (prefix sum of popcounts, then hit h -> thread h mod 128)
// and evaluates those pairs ONCE, in FP64, at state 1 and state 2:
//     dF_i = sum_k f_i(state 2) - f_i(state 1),   u_i = sum_k w_k (e2 - e1),
// w_k = 1/2 when k is displaced too (that pair is seen from k's block as well), else 1.  The
// opposite force of every hit, pairf[row][h] = -(f_i(2) - f_i(1)), and the per-word prefix counts
// hitpre[row][w] are left for the gather kernel, which sums them per resting atom.  The other
// displaced atoms (different displacement group) are walked directly.
struct ProbeAcc {
    double fx, fy, fz, u;
    long long c1, c2;
};

struct ProbeAtom {
    int i, gi, flags;
    double x1, y1, z1, x2, y2, z2, q, hsig, heps;
};

// Exact (FP64) dual-state term of the pair (displaced atom P, atom k) added to A; returns the
// force difference on i, f_i(state 2) - f_i(state 1), in (px, py, pz).
__device__ __forceinline__ void probe_pair(const Topology& T, const double* __restrict__ pos,
                                           const ProbeAtom& P, int k, bool check_excl, ProbeAcc& A,
                                           double& px, double& py, double& pz) {
    px = py = pz = 0.0;
    const int gk = T.group[k];
    if (gk == P.gi) return;  // same displacement (includes k == i): pair unchanged
    const bool cutoff = T.method != SDM_NOCUTOFF;
    const double xk = pos[3 * k], yk = pos[3 * k + 1], zk = pos[3 * k + 2];
    PairGeom g1 = geom(T, P.x1, P.y1, P.z1, xk, yk, zk);
    double xk2 = xk, yk2 = yk, zk2 = zk;
    if (gk != 0) { xk2 += T.disp[3 * k]; yk2 += T.disp[3 * k + 1]; zk2 += T.disp[3 * k + 2]; }
    PairGeom g2 = geom(T, P.x2, P.y2, P.z2, xk2, yk2, zk2);
    const bool in1 = !cutoff || g1.r2 <= T.rc2;
    const bool in2 = !cutoff || g2.r2 <= T.rc2;
    if (!(in1 || in2)) return;
    if (check_excl && is_excluded(T, P.i, k)) return;
    const double sig = T.lj_geom ? sqrt(P.hsig * T.hsig[k]) : P.hsig + T.hsig[k], eps = P.heps * T.heps[k];
    const double qq = SDM_K_COULOMB * P.q * T.q[k];
    const double w = (gk != 0) ? 0.5 : 1.0;
    const int wc = (gk != 0) ? 1 : 2;
    if (in1) {
        double e;
        double fs = pair_term_f64(g1.r2, sig, eps, qq, cutoff, T.krf, T.crf, &e, T.ewald ? T.alpha : 0.0);
        px -= fs * g1.dx; py -= fs * g1.dy; pz -= fs * g1.dz;
        A.u -= w * e;
        A.c1 += wc;
    }
    if (in2) {
        double e;
        double fs = pair_term_f64(g2.r2, sig, eps, qq, cutoff, T.krf, T.crf, &e, T.ewald ? T.alpha : 0.0);
        px += fs * g2.dx; py += fs * g2.dy; pz += fs * g2.dz;
        A.u += w * e;
        A.c2 += wc;
    }
    A.fx += px; A.fy += py; A.fz += pz;
}

#ifndef SDM_PROBE_THREADS
#define SDM_PROBE_THREADS 512
#endif
constexpr int kProbeThreads = SDM_PROBE_THREADS;
constexpr int kProbeWarps = kProbeThreads / 32;
constexpr int kProbeWords = 1024;   // bitmap words handled per pass

// Second half of both probe kernels: the other displaced atoms (different displacement group;
// walked directly, no prefilter), then ONE fixed-order block reduction for all six sums (warp
// trees, thread 0 adds the warp totals in warp order) and the per-atom outputs.
__device__ __forceinline__ void probe_finish(const Topology& T, const EvalBuffers& B, const ProbeAtom& P,
                                             ProbeAcc& A, const double* __restrict__ pos, int m, int r,
                                             double* s_red /* 4*kProbeWarps */,
                                             long long* s_redl /* 2*kProbeWarps */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n = T.n;
    for (int mm = threadIdx.x; mm < T.n_lig; mm += kProbeThreads) {
        double px, py, pz;
        probe_pair(T, pos, P, T.lig_idx[mm], true, A, px, py, pz);
    }
    double sx = warp_sum(A.fx), sy = warp_sum(A.fy), sz = warp_sum(A.fz), su = warp_sum(A.u);
    long long sc1 = warp_sum_ll(A.c1), sc2 = warp_sum_ll(A.c2);
    __syncthreads();
    if (lane == 0) {
        s_red[warp] = sx; s_red[kProbeWarps + warp] = sy; s_red[2 * kProbeWarps + warp] = sz;
        s_red[3 * kProbeWarps + warp] = su;
        s_redl[warp] = sc1; s_redl[kProbeWarps + warp] = sc2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        sx = sy = sz = su = 0.0;
        sc1 = sc2 = 0;
        for (int k = 0; k < kProbeWarps; k++) {
            sx += s_red[k]; sy += s_red[kProbeWarps + k]; sz += s_red[2 * kProbeWarps + k];
            su += s_red[3 * kProbeWarps + k];
            sc1 += s_redl[k]; sc2 += s_redl[kProbeWarps + k];
        }
        double* dF = B.dF + (size_t)r * 3 * n;
        dF[3 * P.i] = sx; dF[3 * P.i + 1] = sy; dF[3 * P.i + 2] = sz;
        B.upart[(size_t)r * T.n_lig + m] = su;
        B.mcnt[((size_t)r * T.n_lig + m) * 2] = sc1;
        B.mcnt[((size_t)r * T.n_lig + m) * 2 + 1] = sc2;
    }
}

__device__ __forceinline__ ProbeAtom load_probe_atom(const Topology& T, const double* __restrict__ pos, int m) {
    ProbeAtom P;
    P.i = T.lig_idx[m];
    P.gi = T.group[P.i];
    P.flags = T.lig_flags[m];
    P.x1 = pos[3 * P.i]; P.y1 = pos[3 * P.i + 1]; P.z1 = pos[3 * P.i + 2];
    P.x2 = P.x1 + T.disp[3 * P.i]; P.y2 = P.y1 + T.disp[3 * P.i + 1]; P.z2 = P.z1 + T.disp[3 * P.i + 2];
    P.q = T.q[P.i]; P.hsig = T.hsig[P.i]; P.heps = T.heps[P.i];
    return P;
}

__global__ void __launch_bounds__(kProbeThreads)
ligand_probe_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    __shared__ double s_red[4 * kProbeWarps];
    __shared__ long long s_redl[2 * kProbeWarps];
    __shared__ uint32_t s_bits[kProbeWords];
    __shared__ int s_pre[kProbeWords + 1];
    __shared__ int s_wsum[kProbeThreads / 32 + 1];
    const int m = blockIdx.x, r = blockIdx.y, n = T.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* pos = B.pos + (size_t)r * 3 * n;
    const ProbeAtom P = load_probe_atom(T, pos, m);
    int begin, end;
    scan_range(T, B, r, &begin, &end);
    const size_t row = (size_t)r * T.n_lig + m;
    // bitmap and prefix layout [replica][scan word][displaced atom]: stride n_lig between words
    const uint32_t* bits = B.hitbits + (size_t)r * B.scan_words * T.n_lig + m;
    int* pre_out = B.hitpre + (size_t)r * B.scan_words * T.n_lig + m;
    const size_t wstride = (size_t)T.n_lig;
    double* pf_out = B.pairf + row * (size_t)B.pairf_cap * 3;
    const int nwords = min(B.scan_words, (end - begin + 31) / 32);
    ProbeAcc A{0, 0, 0, 0, 0, 0};
    int row_base = 0;   // hits in the words already handled

    // (1) resting atoms named by the prefilter bitmap
    for (int w0 = 0; w0 < nwords; w0 += kProbeWords) {
        const int nw = min(kProbeWords, nwords - w0);
        __syncthreads();
        // popcounts and their block-wide exclusive prefix (each thread owns a contiguous run)
        constexpr int kRun = kProbeWords / kProbeThreads;
        int run = 0;
        for (int k = 0; k < kRun; k++) {
            const int w = threadIdx.x * kRun + k;
            const uint32_t v = w < nw ? bits[(size_t)(w0 + w) * wstride] : 0u;
            s_bits[w] = v;
            run += __popc(v);
        }
        int incl = run;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp + 1] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_wsum[0] = 0;
            for (int k = 1; k <= kProbeThreads / 32; k++) s_wsum[k] += s_wsum[k - 1];
        }
        __syncthreads();
        int acc = s_wsum[warp] + incl - run;
        for (int k = 0; k < kRun; k++) {
            const int w = threadIdx.x * kRun + k;
            s_pre[w] = acc;
            if (w < nw) pre_out[(size_t)(w0 + w) * wstride] = row_base + acc;
            acc += __popc(s_bits[w]);
        }
        if (threadIdx.x == kProbeThreads - 1) s_pre[kProbeWords] = acc;
        __syncthreads();
        const int total = s_pre[kProbeWords];
        for (int h = threadIdx.x; h < total; h += kProbeThreads) {
            // word holding hit h: last w with s_pre[w] <= h
            int lo_ = 0, hi_ = kProbeWords;
            while (hi_ - lo_ > 1) {
                const int mid = (lo_ + hi_) >> 1;
                if (s_pre[mid] <= h) lo_ = mid; else hi_ = mid;
            }
            const int bit = __fns(s_bits[lo_], 0, h - s_pre[lo_] + 1);
            const int idx = begin + (w0 + lo_) * 32 + bit;
            int k = idx - begin;
            if (B.scan_atom) k = B.scan_atom[idx] - r * n;
            double px, py, pz;
            probe_pair(T, pos, P, k, (P.flags & 1) != 0, A, px, py, pz);
            const int hg = row_base + h;
            if (hg < B.pairf_cap) {
                pf_out[3 * (size_t)hg] = -px; pf_out[3 * (size_t)hg + 1] = -py; pf_out[3 * (size_t)hg + 2] = -pz;
            }
        }
        row_base += total;
    }
    if (row_base > B.pairf_cap && threadIdx.x == 0) atomicExch(B.flags + r, SDM_ERR_CAPACITY);
    probe_finish(T, B, P, A, pos, m, r, s_red, s_redl);
}

// ---- static candidates (cluster path) -------------------------------------------------------------
// At every list rebuild the prefilter runs once with the list's skin added to the cutoff; this
// kernel (one block per (displaced atom, replica)) expands the bitmap row into the candidate list
// cand[row][h] = resting atom of hit h, in scan order, and writes the per-word prefix hitpre the
// gather kernel needs.  Bitmap, prefix and candidates stay valid as long as the pair list does (no
// atom further than skin/2 from its position at build time), so an evaluation only runs the
// probe-list and gather kernels.
constexpr int kCompactThreads = 256;

__global__ void __launch_bounds__(kCompactThreads)
ligand_compact_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    __shared__ uint32_t s_bits[kProbeWords];
    __shared__ int s_wsum[kCompactThreads / 32 + 1];
    const int m = blockIdx.x, r = blockIdx.y, n = T.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int begin, end;
    scan_range(T, B, r, &begin, &end);
    const size_t row = (size_t)r * T.n_lig + m;
    const uint32_t* bits = B.hitbits + (size_t)r * B.scan_words * T.n_lig + m;
    int* pre_out = B.hitpre + (size_t)r * B.scan_words * T.n_lig + m;
    const size_t wstride = (size_t)T.n_lig;
    int* cand = B.cand + row * (size_t)B.pairf_cap;
    const int nwords = min(B.scan_words, (end - begin + 31) / 32);
    int row_base = 0;
    for (int w0 = 0; w0 < nwords; w0 += kProbeWords) {
        const int nw = min(kProbeWords, nwords - w0);
        __syncthreads();
        constexpr int kRun = kProbeWords / kCompactThreads;
        int run = 0;
        for (int k = 0; k < kRun; k++) {
            const int w = threadIdx.x * kRun + k;
            const uint32_t v = w < nw ? bits[(size_t)(w0 + w) * wstride] : 0u;
            s_bits[w] = v;
            run += __popc(v);
        }
        int incl = run;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp + 1] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_wsum[0] = 0;
            for (int k = 1; k <= kCompactThreads / 32; k++) s_wsum[k] += s_wsum[k - 1];
        }
        __syncthreads();
        int acc = row_base + s_wsum[warp] + incl - run;
        for (int k = 0; k < kRun; k++) {
            const int w = threadIdx.x * kRun + k;
            if (w >= nw) break;
            pre_out[(size_t)(w0 + w) * wstride] = acc;
            uint32_t v = s_bits[w];
            while (v) {
                const int bit = __ffs(v) - 1;
                v &= v - 1u;
                const int idx = begin + (w0 + w) * 32 + bit;
                if (acc < B.pairf_cap) cand[acc] = B.scan_atom ? B.scan_atom[idx] - r * n : idx - begin;
                acc++;
            }
        }
        row_base += s_wsum[kCompactThreads / 32];
    }
    if (threadIdx.x == 0) B.cand_count[row] = row_base;   // > pairf_cap: reported by every evaluation
}

// Per evaluation: the exact (FP64) dual-state terms of the candidates of one displaced atom, one
// block per (displaced atom, replica).  pairf[row][h] = -(f_i(2) - f_i(1)) for the gather kernel
// (zero for a candidate that is outside the cutoff in both states).
__global__ void __launch_bounds__(kProbeThreads)
ligand_probe_list_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    __shared__ double s_red[4 * kProbeWarps];
    __shared__ long long s_redl[2 * kProbeWarps];
    const int m = blockIdx.x, r = blockIdx.y, n = T.n;
    const double* pos = B.pos + (size_t)r * 3 * n;
    const ProbeAtom P = load_probe_atom(T, pos, m);
    const size_t row = (size_t)r * T.n_lig + m;
    const int* cand = B.cand + row * (size_t)B.pairf_cap;
    double* pf_out = B.pairf + row * (size_t)B.pairf_cap * 3;
    const int total = B.cand_count[row];
    if (total > B.pairf_cap && threadIdx.x == 0) atomicExch(B.flags + r, SDM_ERR_CAPACITY);
    const int cnt = min(total, B.pairf_cap);
    ProbeAcc A{0, 0, 0, 0, 0, 0};
    for (int h = threadIdx.x; h < cnt; h += kProbeThreads) {
        double px, py, pz;
        probe_pair(T, pos, P, cand[h], (P.flags & 1) != 0, A, px, py, pz);
        pf_out[3 * (size_t)h] = -px; pf_out[3 * (size_t)h + 1] = -py; pf_out[3 * (size_t)h + 2] = -pz;
    }
    probe_finish(T, B, P, A, pos, m, r, s_red, s_redl);
}

// The same work sized to run BESIDE the persistent pair kernel.  One 512-thread block of the kernel above holds
// 50 k registers: while it is resident the pair kernel gets a quarter of that SM, so at 16 replicas (608 rows, four
// waves) the displaced-atom kernels cost the evaluation their whole duration (measured: 44 us of 280).  Here a few
// small blocks (three per SM, 64 threads = 4.6 k registers each) walk the rows with a grid stride and the pair kernel leaves
// them the room (launch_pair_rows' reserve): both run at the same time, the FP64 latency chains of these rows in the
// shadow of the FP32 pipe.
// Summation order, independent of the block size (a replica must give the same bits in every batch size): hits in
// chunks of 32 consecutive candidates -- lane = hit % 32, butterfly per chunk -- then the other displaced atoms as
// further chunks; thread 0 adds the chunk sums in chunk order.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
ligand_probe_rows_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B, const int rows,
                         const int max_chunks) {
    extern __shared__ double s_chunk[];                 // [max_chunks][4] sums, then [max_chunks][2] counts
    long long* s_cnt = reinterpret_cast<long long*>(s_chunk + 4 * (size_t)max_chunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n = T.n;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int r = row / T.n_lig, m = row - r * T.n_lig;
        const double* pos = B.pos + (size_t)r * 3 * n;
        const ProbeAtom P = load_probe_atom(T, pos, m);
        const int* cand = B.cand + (size_t)row * (size_t)B.pairf_cap;
        double* pf_out = B.pairf + (size_t)row * (size_t)B.pairf_cap * 3;
        const int total = B.cand_count[row];
        if (total > B.pairf_cap && threadIdx.x == 0) atomicExch(B.flags + r, SDM_ERR_CAPACITY);
        const int cnt = min(total, B.pairf_cap);
        const int nch1 = (cnt + 31) >> 5, nch = min(nch1 + ((T.n_lig + 31) >> 5), max_chunks);
        for (int ch = warp; ch < nch; ch += THREADS / 32) {
            ProbeAcc A{0, 0, 0, 0, 0, 0};
            double px, py, pz;
            if (ch < nch1) {
                const int h = ch * 32 + lane;
                if (h < cnt) {
                    probe_pair(T, pos, P, cand[h], (P.flags & 1) != 0, A, px, py, pz);
                    pf_out[3 * (size_t)h] = -px; pf_out[3 * (size_t)h + 1] = -py; pf_out[3 * (size_t)h + 2] = -pz;
                }
            } else {
                const int mm = (ch - nch1) * 32 + lane;
                if (mm < T.n_lig) probe_pair(T, pos, P, T.lig_idx[mm], true, A, px, py, pz);
            }
            const double sx = warp_sum(A.fx), sy = warp_sum(A.fy), sz = warp_sum(A.fz), su = warp_sum(A.u);
            const long long sc1 = warp_sum_ll(A.c1), sc2 = warp_sum_ll(A.c2);
            if (lane == 0) {
                s_chunk[4 * ch] = sx; s_chunk[4 * ch + 1] = sy; s_chunk[4 * ch + 2] = sz; s_chunk[4 * ch + 3] = su;
                s_cnt[2 * ch] = sc1; s_cnt[2 * ch + 1] = sc2;
            }
        }
        __syncthreads();
        // one thread per quantity, chunks in order
        if (threadIdx.x < 4) {
            double v = 0.0;
            for (int ch = 0; ch < nch; ch++) v += s_chunk[4 * ch + threadIdx.x];
            if (threadIdx.x < 3) B.dF[(size_t)r * 3 * n + 3 * P.i + threadIdx.x] = v;
            else B.upart[row] = v;
        } else if (threadIdx.x < 6) {
            long long v = 0;
            for (int ch = 0; ch < nch; ch++) v += s_cnt[2 * ch + (threadIdx.x - 4)];
            B.mcnt[(size_t)row * 2 + (threadIdx.x - 4)] = v;
        }
        __syncthreads();
    }
}

// All-pairs path, small systems (a few hundred atoms, hundreds of replicas): ONE WARP per (displaced atom,
// replica) instead of a block -- lane = scan index of the current bitmap word, hit index = running prefix +
// popcount below the lane: no block barriers, no search for the n-th set bit, eight rows per block.  Same
// outputs and layouts as ligand_probe_kernel (sums in a different, still fixed, order).
constexpr int kProbeWarpRows = 8;
__global__ void __launch_bounds__(32 * kProbeWarpRows)
ligand_probe_warp_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    const int lane = threadIdx.x & 31;
    const int rowi = blockIdx.x * kProbeWarpRows + (threadIdx.x >> 5);
    if (rowi >= B.R * T.n_lig) return;
    const int r = rowi / T.n_lig, m = rowi - r * T.n_lig, n = T.n;
    const double* pos = B.pos + (size_t)r * 3 * n;
    const ProbeAtom P = load_probe_atom(T, pos, m);
    int begin, end;
    scan_range(T, B, r, &begin, &end);
    const size_t row = (size_t)rowi;
    const uint32_t* bits = B.hitbits + (size_t)r * B.scan_words * T.n_lig + m;
    int* pre_out = B.hitpre + (size_t)r * B.scan_words * T.n_lig + m;
    const size_t wstride = (size_t)T.n_lig;
    double* pf_out = B.pairf + row * (size_t)B.pairf_cap * 3;
    const int nwords = min(B.scan_words, (end - begin + 31) / 32);
    ProbeAcc A{0, 0, 0, 0, 0, 0};
    int pre = 0;
    for (int w = 0; w < nwords; w++) {
        const uint32_t v = bits[(size_t)w * wstride];
        if (lane == 0) pre_out[(size_t)w * wstride] = pre;
        if ((v >> lane) & 1u) {
            const int idx = begin + w * 32 + lane;
            int k = idx - begin;
            if (B.scan_atom) k = B.scan_atom[idx] - r * n;
            double px, py, pz;
            probe_pair(T, pos, P, k, (P.flags & 1) != 0, A, px, py, pz);
            const int hg = pre + __popc(v & ((1u << lane) - 1u));
            if (hg < B.pairf_cap) {
                pf_out[3 * (size_t)hg] = -px; pf_out[3 * (size_t)hg + 1] = -py; pf_out[3 * (size_t)hg + 2] = -pz;
            }
        }
        pre += __popc(v);
    }
    if (pre > B.pairf_cap && lane == 0) atomicExch(B.flags + r, SDM_ERR_CAPACITY);
    for (int mm = lane; mm < T.n_lig; mm += 32) {
        double px, py, pz;
        probe_pair(T, pos, P, T.lig_idx[mm], true, A, px, py, pz);
    }
    const double sx = warp_sum(A.fx), sy = warp_sum(A.fy), sz = warp_sum(A.fz), su = warp_sum(A.u);
    const long long sc1 = warp_sum_ll(A.c1), sc2 = warp_sum_ll(A.c2);
    if (lane == 0) {
        double* dF = B.dF + (size_t)r * 3 * n;
        dF[3 * P.i] = sx; dF[3 * P.i + 1] = sy; dF[3 * P.i + 2] = sz;
        B.upart[row] = su;
        B.mcnt[row * 2] = sc1;
        B.mcnt[row * 2 + 1] = sc2;
    }
}

// ---- gather kernel -------------------------------------------------------------------------------
// One thread per scan index (a NON-displaced atom j): dF_j = sum over the displaced atoms m (in
// index order: fixed summation order) of the force the probe kernel stored for the pair, found
// through the bitmap: h = hitpre[row][w] + popc(bits below this lane).  Pure loads; writes every
// dF_j (zero when nothing is near), so no memset is needed.
constexpr int kGatherThreads = 128;
constexpr int kGatherBatch = 4;

__global__ void __launch_bounds__(kGatherThreads)
ligand_gather_kernel(const __grid_constant__ Topology T, const __grid_constant__ EvalBuffers B) {
    const int n = T.n, r = blockIdx.y;
    int begin, end;
    scan_range(T, B, r, &begin, &end);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sw = blockIdx.x * (kGatherThreads / 32) + warp;
    if (sw >= B.scan_words) return;
    const int idx = begin + sw * 32 + lane;
    int j = -1;
    if (idx < end) {
        j = idx - begin;
        if (B.scan_atom) {
            const int ga = B.scan_atom[idx];
            j = ga < 0 ? -1 : ga - r * n;
        }
    }
    const bool active = j >= 0 && T.group[j] == 0;
    double fx = 0, fy = 0, fz = 0;
    const uint32_t below = (1u << lane) - 1u;
    const size_t wbase = ((size_t)r * B.scan_words + sw) * T.n_lig;
    for (int m0 = 0; m0 < T.n_lig; m0 += 32) {
        // the words (and their hit prefixes) of this warp for 32 displaced atoms: two coalesced
        // loads, most warps see zeros
        const bool mv = m0 + lane < T.n_lig;
        const uint32_t mine = mv ? B.hitbits[wbase + m0 + lane] : 0u;
        const int pre_mine = mv ? B.hitpre[wbase + m0 + lane] : 0;
        unsigned todo = __ballot_sync(0xffffffffu, mine != 0u);
        while (todo) {
            // four displaced atoms per round: their loads are in flight together; the sums keep
            // the displaced-atom order (a skipped term adds +0.0)
            double v[kGatherBatch][3];
#pragma unroll
            for (int b = 0; b < kGatherBatch; b++) {
                v[b][0] = v[b][1] = v[b][2] = 0.0;
                if (todo) {   // uniform
                    const int ml = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const uint32_t word = __shfl_sync(0xffffffffu, mine, ml);
                    const int h = __shfl_sync(0xffffffffu, pre_mine, ml) + __popc(word & below);
                    // h >= cap: overflow was flagged by the probe kernel
                    if (((word >> lane) & 1u) && h < B.pairf_cap) {
                        const double* f = B.pairf + (((size_t)r * T.n_lig + m0 + ml) * (size_t)B.pairf_cap + h) * 3;
                        v[b][0] = f[0]; v[b][1] = f[1]; v[b][2] = f[2];
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < kGatherBatch; b++) { fx += v[b][0]; fy += v[b][1]; fz += v[b][2]; }
        }
    }
    if (active) {
        double* dF = B.dF + (size_t)r * 3 * n;
        dF[3 * j] = fx; dF[3 * j + 1] = fy; dF[3 * j + 2] = fz;
    }
}

// ---------------------------------------------------------------------------------------------
// 1-4 exceptions (OpenMM 7.3 ReferenceLJCoulomb14: no cutoff, no reaction field, plain delta).
// State-independent unless the two atoms carry different displacements; that rare case also
// feeds dF and u (FP64 atomics on dF -- must run after the ligand kernels).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double exc_term(double dx, double dy, double dz, double qq,
                                           double sigma, double eps4, double* e) {
    double r2 = dx * dx + dy * dy + dz * dz;
    double inverseR = 1.0 / sqrt(r2);
    double sig2 = inverseR * sigma;
    sig2 *= sig2;
    double sig6 = sig2 * sig2 * sig2;
    double dEdR = eps4 * (12.0 * sig6 - 6.0) * sig6;
    dEdR += SDM_K_COULOMB * qq * inverseR;
    dEdR *= inverseR * inverseR;
    *e = eps4 * (sig6 - 1.0) * sig6 + SDM_K_COULOMB * qq * inverseR;
    return dEdR;
}

// -qq*erf(alpha r)/r of an excluded pair (energy returned, *fs: F_a += fs*d, F_b -= fs*d); for r -> 0 the
// limit -qq*2*alpha/sqrt(pi) without a force, as OpenMM does when erf(alpha r) <= 1e-6.
__device__ __forceinline__ double ewald_exclusion_term(double r2, double qq, double alpha, double* fs) {
    const double r = sqrt(r2);
    const double alphaR = alpha * r;
    const double ef = erf(alphaR);
    if (ef > 1e-6) {
        const double inverseR = 1.0 / r;
        *fs = -qq * inverseR * inverseR * inverseR * (ef - alphaR * exp(-alphaR * alphaR) * 1.1283791670955126);
        return -qq * inverseR * ef;
    }
    *fs = 0.0;
    return -alpha * 1.1283791670955126 * qq;
}

__global__ void __launch_bounds__(128)
exceptions_kernel(Topology T, const double* __restrict__ pos_all, long long* __restrict__ f1acc,
                  size_t acc_rstride, int nslot, const int* __restrict__ slot_of, double* __restrict__ dF_all,
                  double* __restrict__ eexc_part, double* __restrict__ uexc_part, int n_excpart) {
    __shared__ double s_red[32];
    const int n = T.n, r = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const double* pos = pos_all + (size_t)r * 3 * n;
    double e1 = 0.0, du = 0.0;
    if (k < T.n_exceptions) {
        const int a = T.exc_pairs[2 * k], b = T.exc_pairs[2 * k + 1];
        const double qq = T.exc_params[3 * k], sigma = T.exc_params[3 * k + 1];
        const double eps = T.exc_params[3 * k + 2];
        if (qq != 0.0 || eps != 0.0) {
            const double eps4 = 4.0 * eps;
            double dx = pos[3 * a] - pos[3 * b], dy = pos[3 * a + 1] - pos[3 * b + 1],
                   dz = pos[3 * a + 2] - pos[3 * b + 2];
            double fs = exc_term(dx, dy, dz, qq, sigma, eps4, &e1);
            long long* acc = f1acc + (size_t)r * acc_rstride;
            const int sa = slot_of ? slot_of[(size_t)r * n + a] : a;
            const int sb = slot_of ? slot_of[(size_t)r * n + b] : b;
            atomic_add_fixed(acc + sa, to_fixed(fs * dx));
            atomic_add_fixed(acc + nslot + sa, to_fixed(fs * dy));
            atomic_add_fixed(acc + 2 * nslot + sa, to_fixed(fs * dz));
            atomic_add_fixed(acc + sb, to_fixed(-fs * dx));
            atomic_add_fixed(acc + nslot + sb, to_fixed(-fs * dy));
            atomic_add_fixed(acc + 2 * nslot + sb, to_fixed(-fs * dz));
            if (T.group[a] != T.group[b]) {
                double dx2 = (pos[3 * a] + T.disp[3 * a]) - (pos[3 * b] + T.disp[3 * b]);
                double dy2 = (pos[3 * a + 1] + T.disp[3 * a + 1]) - (pos[3 * b + 1] + T.disp[3 * b + 1]);
                double dz2 = (pos[3 * a + 2] + T.disp[3 * a + 2]) - (pos[3 * b + 2] + T.disp[3 * b + 2]);
                double e2;
                double fs2 = exc_term(dx2, dy2, dz2, qq, sigma, eps4, &e2);
                double* dF = dF_all + (size_t)r * 3 * n;
                atomicAdd(dF + 3 * a, fs2 * dx2 - fs * dx);
                atomicAdd(dF + 3 * a + 1, fs2 * dy2 - fs * dy);
                atomicAdd(dF + 3 * a + 2, fs2 * dz2 - fs * dz);
                atomicAdd(dF + 3 * b, -(fs2 * dx2 - fs * dx));
                atomicAdd(dF + 3 * b + 1, -(fs2 * dy2 - fs * dy));
                atomicAdd(dF + 3 * b + 2, -(fs2 * dz2 - fs * dz));
                du = e2 - e1;
            }
        }
    }
    // Direct-space Ewald: the excluded pairs were implicitly included in the reciprocal-space sum, their
    // erf(alpha r)/r part is taken out here (OpenMM 7.3 ReferenceLJCoulombIxn::calculateEwaldIxn, "subtract
    // off the exclusions"): items n_exceptions .. n_exceptions + n_excl_pairs - 1 of this kernel.
    const int kx = k - T.n_exceptions;
    if (T.ewald && kx >= 0 && kx < T.n_excl_pairs) {
        const int a = T.excl_pairs[2 * kx], b = T.excl_pairs[2 * kx + 1];
        const double qq = SDM_K_COULOMB * T.q[a] * T.q[b];
        const PairGeom g = geom(T, pos[3 * a], pos[3 * a + 1], pos[3 * a + 2], pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
        double fs;
        e1 = ewald_exclusion_term(g.r2, qq, T.alpha, &fs);
        long long* acc = f1acc + (size_t)r * acc_rstride;
        const int sa = slot_of ? slot_of[(size_t)r * n + a] : a;
        const int sb = slot_of ? slot_of[(size_t)r * n + b] : b;
        atomic_add_fixed(acc + sa, to_fixed(fs * g.dx));
        atomic_add_fixed(acc + nslot + sa, to_fixed(fs * g.dy));
        atomic_add_fixed(acc + 2 * nslot + sa, to_fixed(fs * g.dz));
        atomic_add_fixed(acc + sb, to_fixed(-fs * g.dx));
        atomic_add_fixed(acc + nslot + sb, to_fixed(-fs * g.dy));
        atomic_add_fixed(acc + 2 * nslot + sb, to_fixed(-fs * g.dz));
        if (T.group[a] != T.group[b]) {   // an excluded pair with different displacements: state 2 differs
            const PairGeom g2 = geom(T, pos[3 * a] + T.disp[3 * a], pos[3 * a + 1] + T.disp[3 * a + 1],
                                     pos[3 * a + 2] + T.disp[3 * a + 2], pos[3 * b] + T.disp[3 * b],
                                     pos[3 * b + 1] + T.disp[3 * b + 1], pos[3 * b + 2] + T.disp[3 * b + 2]);
            double fs2;
            const double e2 = ewald_exclusion_term(g2.r2, qq, T.alpha, &fs2);
            double* dF = dF_all + (size_t)r * 3 * n;
            atomicAdd(dF + 3 * a, fs2 * g2.dx - fs * g.dx);
            atomicAdd(dF + 3 * a + 1, fs2 * g2.dy - fs * g.dy);
            atomicAdd(dF + 3 * a + 2, fs2 * g2.dz - fs * g.dz);
            atomicAdd(dF + 3 * b, -(fs2 * g2.dx - fs * g.dx));
            atomicAdd(dF + 3 * b + 1, -(fs2 * g2.dy - fs * g.dy));
            atomicAdd(dF + 3 * b + 2, -(fs2 * g2.dz - fs * g.dz));
            du = e2 - e1;
        }
    }
    double se = block_sum(e1, s_red);
    double sd = block_sum(du, s_red);
    if (threadIdx.x == 0) {
        eexc_part[(size_t)r * n_excpart + blockIdx.x] = se;
        uexc_part[(size_t)r * n_excpart + blockIdx.x] = sd;
    }
}

// ---------------------------------------------------------------------------------------------
// scalar stage: one block per replica.  Fixed-order reductions of the partials, then thread 0
// runs SoftCoreF + bias + bookkeeping on the device -- no host round trip (the reference pays
// three D->H energy reads per step).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scalars_kernel(Topology T, EvalBuffers B, double e_scale, int c_div) {
    // all seven sums in ONE pass: independent loads in flight together, one barrier, fixed order.
    // 1024 threads: the per-unit partials of a replica (a few thousand) are two or three loads per
    // thread instead of a chain of ten -- this kernel sits on the critical path of every evaluation
    __shared__ double s_d[4][32];
    __shared__ long long s_l[3][32];
    const int r = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = B.part_off[r], np_ = B.part_off[r + 1] - p0;
    const int nmax = max(np_, max(B.n_excpart, T.n_lig));
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long b0 = 0, b1 = 0, b2 = 0;
    for (int k = threadIdx.x; k < nmax; k += blockDim.x) {
        if (k < np_) { a0 += B.epart[p0 + k]; b0 += B.cpart[p0 + k]; }
        if (k < B.n_excpart) {
            a1 += B.eexc_part[(size_t)r * B.n_excpart + k];
            a2 += B.uexc_part[(size_t)r * B.n_excpart + k];
        }
        if (k < T.n_lig) {
            a3 += B.upart[(size_t)r * T.n_lig + k];
            b1 += B.mcnt[((size_t)r * T.n_lig + k) * 2];
            b2 += B.mcnt[((size_t)r * T.n_lig + k) * 2 + 1];
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    b0 = warp_sum_ll(b0); b1 = warp_sum_ll(b1); b2 = warp_sum_ll(b2);
    if (lane == 0) {
        s_d[0][warp] = a0; s_d[1][warp] = a1; s_d[2][warp] = a2; s_d[3][warp] = a3;
        s_l[0][warp] = b0; s_l[1][warp] = b1; s_l[2][warp] = b2;
    }
    __syncthreads();
    double ep = 0.0, ee = 0.0, ue = 0.0, ul = 0.0;
    long long c = 0, m1 = 0, m2 = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            ep += s_d[0][w]; ee += s_d[1][w]; ue += s_d[2][w]; ul += s_d[3][w];
            c += s_l[0][w]; m1 += s_l[1][w]; m2 += s_l[2][w];
        }
    if (threadIdx.x == 0) {
        ReplicaState* st = B.state + r;
        sdm_scalars* sc = &st->sc;
        // sticky status: the first stale-list / capacity report since the host last looked survives
        // later evaluations; a device-resident MD loop is told to stop advancing the state
        int status = B.flags[r];
        if (status != 0 && B.sticky[r] == 0) B.sticky[r] = status;
        if (B.sticky[r] != 0) status = B.sticky[r];
        if (B.md_ctl && (status == SDM_ERR_STALE_LIST || status == SDM_ERR_CAPACITY || status == SDM_ERR_CONSTRAINT))
            B.md_ctl[0] = 1ull;
        sc->status = status;
        sc->E1_pair = ep * e_scale;
        sc->E1_exc = ee;
        sc->E1_disp = T.e_disp;
        sc->E1 = sc->E1_pair + sc->E1_exc + sc->E1_disp;
        sc->u = ul + ue;
        if (B.ext_e) {   // contributions computed outside the library (sdm_set_external_dual)
            sc->E1 += B.ext_e[2 * r];
            sc->u += B.ext_e[2 * r + 1] - B.ext_e[2 * r];
        }
        sc->E2 = sc->E1 + sc->u;
        sc->n_pairs1 = c / c_div;
        sc->n_moved1 = m1 / 2;
        sc->n_moved2 = m2 / 2;
        sc->list_age = *B.list_age;
        // sc->Eb was stored by sdm_set_bonded_forces
        execute_scalars(&st->alch, sc);
    }
}

// ---------------------------------------------------------------------------------------------
// mix: F = F1 + sp*(F2 - F1) + Fb  ==  sp*F2 + (1-sp)*F1 + Fb  (ReferenceSDMKernels.cpp:309-318),
// sp read from the device scalar block.  Also converts F1 to double and (optionally) clears the
// fixed-point accumulators for the next evaluation.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mix_kernel(Topology T, EvalBuffers B, int zero_acc) {
    const int n = T.n, r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) B.flags[r] = 0;  // already copied into sc.status by the scalar stage
    if (i == 0 && r == 0 && B.work_counter) *B.work_counter = 0;
    const double sp = B.state[r].sc.sp;
    const int slot = B.slot_of ? B.slot_of[(size_t)r * n + i] : i;
    long long* acc = B.f1acc + (size_t)r * B.acc_rstride;
    const size_t o = (size_t)r * 3 * n + 3 * (size_t)i;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double f1 = (double)acc[(size_t)c * B.nslot + slot] * SDM_INV_FORCE_SCALE;
        double d = B.dF[o + c];
        if (B.ext_f1 && B.ext_on[r]) {   // external dual-state term: F1 += F1_ext, F2 - F1 += F2_ext - F1_ext
            f1 += B.ext_f1[o + c];
            d += B.ext_f2[o + c] - B.ext_f1[o + c];
            B.dF[o + c] = d;   // F2 - F1 as sdm_get_forces reports it (rewritten by the next evaluation)
        }
        const double f = f1 + sp * d + B.fb[o + c];
        B.F1[o + c] = f1;
        B.F[o + c] = f;
        if (zero_acc) acc[(size_t)c * B.nslot + slot] = 0;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int allpairs_num_blocks(int n) { return (n + kTile - 1) / kTile; }
int exceptions_num_blocks(int n_exceptions) { return n_exceptions > 0 ? (n_exceptions + 127) / 128 : 1; }

void launch_prep_posq(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    dim3 grid((T.n + 255) / 256, B.R);
    prep_posq_kernel<<<grid, 256, 0, s>>>(T, B.pos, B.posq);
}

void launch_allpairs(const Topology& T, const EvalBuffers& B, int exact, int* emit_counter,
                     int* emit_pairs, int emit_cap, int emit_replica, cudaStream_t s) {
    dim3 grid(allpairs_num_blocks(T.n), B.R);
    if (T.method == SDM_NOCUTOFF)
        allpairs_kernel<SDM_NOCUTOFF><<<grid, kTile, 0, s>>>(T, B.posq, B.pos, B.f1acc, B.nslot, B.epart, B.cpart, B.n_epart_allpairs, exact, emit_counter, emit_pairs, emit_cap, emit_replica);
    else if (T.method == SDM_CUTOFF_NONPERIODIC)
        allpairs_kernel<SDM_CUTOFF_NONPERIODIC><<<grid, kTile, 0, s>>>(T, B.posq, B.pos, B.f1acc, B.nslot, B.epart, B.cpart, B.n_epart_allpairs, exact, emit_counter, emit_pairs, emit_cap, emit_replica);
    else
        allpairs_kernel<SDM_CUTOFF_PERIODIC><<<grid, kTile, 0, s>>>(T, B.posq, B.pos, B.f1acc, B.nslot, B.epart, B.cpart, B.n_epart_allpairs, exact, emit_counter, emit_pairs, emit_cap, emit_replica);
}

void launch_ligand_probe(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    if (T.n_lig == 0) return;
    if (B.scan_words <= 64) {   // small systems: one warp per (displaced atom, replica)
        const int rows = T.n_lig * B.R;
        ligand_probe_warp_kernel<<<(rows + kProbeWarpRows - 1) / kProbeWarpRows, 32 * kProbeWarpRows, 0, s>>>(T, B);
        return;
    }
    dim3 grid(T.n_lig, B.R);
    ligand_probe_kernel<<<grid, kProbeThreads, 0, s>>>(T, B);
}

void launch_ligand_filter(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    // every scan word (32 scan indices) of a replica gets a warp: the bitmap rows are complete
    if (T.n_lig == 0) return;
    dim3 grid((B.scan_words * 32 + kFilterThreads - 1) / kFilterThreads, B.R);
    ligand_filter_kernel<<<grid, kFilterThreads, 0, s>>>(T, B);
}

void launch_ligand_gather(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    dim3 grid((B.scan_words * 32 + kGatherThreads - 1) / kGatherThreads, B.R);
    ligand_gather_kernel<<<grid, kGatherThreads, 0, s>>>(T, B);
}

void launch_ligand_compact(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    if (T.n_lig <= 0) return;
    dim3 grid(T.n_lig, B.R);
    ligand_compact_kernel<<<grid, kCompactThreads, 0, s>>>(T, B);
}

int ligand_rows_beside_pair_kernel(const Topology& T, const EvalBuffers& B, int num_sms) {
    static const int force = getenv("SDMB200_SIDE_SMALL") ? atoi(getenv("SDMB200_SIDE_SMALL")) : -1;   // development knob
    if (force >= 0) return force;
    // enough rows to occupy the small blocks, and a pair pass long enough to hide them (at low occupancy the rows
    // take about three times as long as in the one-block-per-row kernel)
    return (T.n_lig * B.R >= 2 * num_sms && T.n >= 200 * T.n_lig) ? 1 : 0;
}

void launch_ligand_probe_list(const Topology& T, const EvalBuffers& B, int num_sms, cudaStream_t s) {
    if (T.n_lig <= 0) return;
    const int rows = T.n_lig * B.R;
    const int max_chunks = (B.pairf_cap + 31) / 32 + (T.n_lig + 31) / 32;
    const size_t smem = (size_t)max_chunks * (4 * sizeof(double) + 2 * sizeof(long long));
    if (smem > 200 * 1024) {   // absurd capacities: the one-block-per-row kernel has no such table
        dim3 grid(T.n_lig, B.R);
        ligand_probe_list_kernel<<<grid, kProbeThreads, 0, s>>>(T, B);
        return;
    }
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(ligand_probe_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(ligand_probe_rows_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    if (ligand_rows_beside_pair_kernel(T, B, num_sms)) {
        static const int per_sm = getenv("SDMB200_SIDE_BLOCKS") ? std::max(1, atoi(getenv("SDMB200_SIDE_BLOCKS"))) : 3;
        ligand_probe_rows_kernel<64><<<std::min(rows, num_sms * per_sm), 64, smem, s>>>(T, B, rows, max_chunks);
    } else {
        ligand_probe_rows_kernel<512><<<rows, 512, smem, s>>>(T, B, rows, max_chunks);
    }
}

void launch_exceptions(const Topology& T, const EvalBuffers& B, cudaStream_t s) {
    dim3 grid(exceptions_num_blocks(T.n_exceptions + (T.ewald ? T.n_excl_pairs : 0)), B.R);
    exceptions_kernel<<<grid, 128, 0, s>>>(T, B.pos, B.f1acc, B.acc_rstride, B.nslot, B.slot_of, B.dF, B.eexc_part, B.uexc_part, B.n_excpart);
}

void launch_scalars(const Topology& T, const EvalBuffers& B, double e_scale, int c_div,
                    cudaStream_t s) {
    scalars_kernel<<<B.R, 1024, 0, s>>>(T, B, e_scale, c_div);
}

void launch_mix(const Topology& T, const EvalBuffers& B, int zero_acc, cudaStream_t s) {
    dim3 grid((T.n + 255) / 256, B.R);
    mix_kernel<<<grid, 256, 0, s>>>(T, B, zero_acc);
}

}  // namespace sdm
