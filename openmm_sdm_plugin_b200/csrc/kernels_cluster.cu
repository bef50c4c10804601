// kernels_cluster.cu -- the hot kernel of the fused path: cutoff LJ + reaction-field pair
// interactions over the cluster-pair list (sm_100a, FP32 SIMT pipe; not a dense contraction, so
// no tensor cores).
//
// One warp per work unit (a supercluster = up to 8 clusters x 8 atoms, and a chunk of its
// j-group list).  The 64 i-atoms are staged once in shared memory; lane (tj, ti) = (lane>>3,
// lane&7) keeps j-atom tj of the current j-group in registers and walks the 8 i-clusters, so
//   * the j force accumulates in registers across 8 pair steps and is reduced over ti with 3
//     shuffles per component per ENTRY (not per pair),
//   * the i forces accumulate in 24 registers across the whole unit and are reduced over tj
//     once per unit,
//   * each pair step costs two shared loads (float4 + float2, 8 distinct addresses per warp).
// Forces go to 64-bit fixed-point accumulators (2^32), so the result is independent of the
// order in which warps finish: bit-reproducible across runs, replicas-per-GPU and GPUs.
//
// Arithmetic restated from OpenMM 7.3 ReferenceLJCoulombIxn::calculateOneIxn (SURVEY.md
// Appendix B.3) in FP32: per-atom sigma/2 and 2*sqrt(eps), charges pre-scaled by
// sqrt(ONE_4PI_EPS0), reaction field krf/crf, LJ not shifted.
#include "pairlist.h"

namespace sdm {
namespace {

constexpr int kWarps = 4;

struct Acc {
    float fx, fy, fz;
};

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RSQ, no denormal fix-up
    return y;
}

// Rare path: pairs whose FP32 r^2 lies within T.band of the cutoff are left out by the hot loop
// and handled here, once per affected entry, with the FP64 in-cutoff test of the oracle.  Their
// force goes straight to the fixed-point accumulators (both atoms); energy and count are
// returned through en / cnt.  Out of line so that the hot loop stays small.
__device__ __noinline__ void fix_rare_pairs(const Topology& T, const PairListView& V,
                                            const double* __restrict__ pos_all,
                                            long long* __restrict__ f1acc, const float4* s_xi,
                                            const float2* s_pi, int ibase, int jslot, float4 xj,
                                            float2 pj, uint32_t imask, uint32_t midx, int lane,
                                            float* en, int* cnt, bool use_f64, bool all,
                                            bool emit, int* emit_counter, int* emit_pairs,
                                            int emit_cap) {
    const int ti = lane & 7;
    const size_t plane = (size_t)V.nslot_cap;
    for (int ci = 0; ci < nbl::kMaxCi; ci++) {
        if (!((imask >> ci) & 1u)) continue;
        const uint32_t w = midx ? V.masks[(size_t)midx * nbl::kMaxCi + ci] : 0xffffffffu;
        if (!((w >> lane) & 1u)) continue;
        const int il = ci * nbl::kClusterSize + ti;
        const float4 xi = s_xi[il];
        const float2 pi = s_pi[il];
        const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float t = r2 - T.rc2f;
        const bool near = fabsf(t) < T.band;
        if (!(near || (all && t <= 0.f))) continue;
        const int islot = ibase + il;
        const int ai = V.atom[islot], aj = V.atom[jslot];
        if (ai < 0 || aj < 0) continue;
        const int r = ai / T.n;
        if (near && use_f64) {
            if (!in_cutoff_f64(T, pos_all + (size_t)r * 3 * T.n, ai - r * T.n, aj - r * T.n)) continue;
        } else if (!(t <= 0.f)) {
            continue;
        }
        const float rinv = rsqrtf(r2), rinv2 = rinv * rinv;
        const float sig = pi.x + pj.x;
        const float sr2 = (sig * sig) * rinv2;
        const float sr6 = sr2 * sr2 * sr2;
        const float elj = (pi.y * pj.y) * sr6;
        const float qq = xi.w * xj.w;
        const float kr2 = T.krff * r2;
        const float fs = (elj * (12.f * sr6 - 6.f) + qq * (rinv - 2.f * kr2)) * rinv2;
        *en += elj * (sr6 - 1.f) + qq * (rinv + kr2 - T.crff);
        *cnt += 1;
        const float f[3] = {fs * dx, fs * dy, fs * dz};
        for (int c = 0; c < 3; c++) {
            const long long v = __float2ll_rn(f[c] * 4294967296.0f);
            atomic_add_fixed(f1acc + (size_t)c * plane + islot, v);
            atomic_add_fixed(f1acc + (size_t)c * plane + jslot, -v);
        }
        if (emit) {
            const int a = ai % T.n, b = aj % T.n;
            const int slot = atomicAdd(emit_counter, 1);
            if (slot < emit_cap) {
                emit_pairs[2 * slot] = a < b ? a : b;
                emit_pairs[2 * slot + 1] = a < b ? b : a;
            }
        }
    }
}

// One (i-atom, j-atom) pair of the hot loop.  `rare` accumulates "this lane met a pair inside
// the FP64 re-test band"; such pairs are skipped here and fixed up by fix_rare_pairs().
template <bool EXACT, bool ALL>
__device__ __forceinline__ void pair_step(const Topology& T, const float4 xi, const float2 pi,
                                          const float4 xj, const float2 pj, const bool allowed,
                                          bool& rare, Acc& fi, Acc& fj, float& en, int& cnt) {
    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float t = r2 - T.rc2f;
    bool in = allowed && (t <= 0.f);
    if (EXACT) {
        const bool near = allowed && (ALL ? t < T.band : fabsf(t) < T.band);
        rare |= near;
        in = in && !near;
    }
    const float rinv = rsqrt_approx(r2);
    const float rinv2 = rinv * rinv;
    const float sig = pi.x + pj.x;
    const float sr2 = (sig * sig) * rinv2;
    const float sr6 = sr2 * sr2 * sr2;
    const float elj = (pi.y * pj.y) * sr6;
    const float qq = xi.w * xj.w;
    const float kr2 = T.krff * r2;
    // dE/dr * r  and energy (OpenMM 7.3 ReferenceLJCoulombIxn, reaction field, LJ not shifted)
    const float dEdR = elj * (12.f * sr6 - 6.f) + qq * (rinv - 2.f * kr2);
    const float e = elj * (sr6 - 1.f) + qq * (rinv + kr2 - T.crff);
    const float fs = in ? dEdR * rinv2 : 0.f;
    en += in ? e : 0.f;
    cnt += in;
    fi.fx += fs * dx; fi.fy += fs * dy; fi.fz += fs * dz;
    fj.fx -= fs * dx; fj.fy -= fs * dy; fj.fz -= fs * dz;
}

template <bool PERIODIC, bool EXACT, bool EMIT>
__global__ void __launch_bounds__(kWarps * 32, 8)
pair_cluster_kernel(const __grid_constant__ Topology T, const __grid_constant__ PairListView V,
                    const double* __restrict__ pos_all,
                    long long* __restrict__ f1acc, double* __restrict__ epart,
                    long long* __restrict__ cpart, int exact, int* emit_counter, int* emit_pairs,
                    int emit_cap, int emit_replica) {
    __shared__ float4 s_xi[kWarps][64];
    __shared__ float2 s_pi[kWarps][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarps + warp;
    if (unit >= V.nunits) return;  // warp-uniform; no block-level barrier below
    const Unit u = V.units[unit];
    const nbl::SciDesc sd = V.sci[u.sci];
    const int ibase = sd.c0 * nbl::kClusterSize;
    const int ni = sd.nci * nbl::kClusterSize;
    const bool emit = EMIT && sd.replica == emit_replica;

#pragma unroll
    for (int k = lane; k < 64; k += 32) {
        float4 p = make_float4(-nbl::kFar, -nbl::kFar, -nbl::kFar, 0.f);
        float2 pr = make_float2(0.f, 0.f);
        if (k < ni) {
            const float4 q = V.posq[ibase + k];
            if (q.x < 0.5f * nbl::kFar) { p = q; pr = V.par[ibase + k]; }
        }
        s_xi[warp][k] = p;
        s_pi[warp][k] = pr;
    }
    __syncwarp();

    const int ti = lane & 7, tj = lane >> 3;
    const uint32_t lanebit = 1u << lane;
    Acc fi[nbl::kMaxCi];
#pragma unroll
    for (int ci = 0; ci < nbl::kMaxCi; ci++) fi[ci] = Acc{0.f, 0.f, 0.f};
    float en = 0.f;
    int cnt = 0;
    const size_t plane = (size_t)V.nslot_cap;

    for (int e = u.begin; e < u.end; e++) {
        const uint2 ent = V.entries[e];
        const int j4 = (int)(ent.x & 0x3ffffffu);
        const uint32_t code = ent.x >> 26;
        const uint32_t imask = ent.y & 0xffu;
        const uint32_t midx = ent.y >> 8;
        const int jslot = j4 * nbl::kJGroup + tj;
        float4 xj = V.posq[jslot];
        const float2 pj = V.par[jslot];
        if (PERIODIC) {
            xj.x += (float)nbl::shift_x(code) * T.boxf[0];
            xj.y += (float)nbl::shift_y(code) * T.boxf[1];
            xj.z += (float)nbl::shift_z(code) * T.boxf[2];
        }
        Acc fj{0.f, 0.f, 0.f};
        bool rare = false;
        // mask set 0 is "all ones": only masked entries (diagonal / bonded neighbours) load words
        const uint32_t* mw = V.masks + (size_t)midx * nbl::kMaxCi;
#pragma unroll
        for (int ci = 0; ci < nbl::kMaxCi; ci++) {
            if ((imask >> ci) & 1u) {
                const int il = ci * nbl::kClusterSize + ti;
                const uint32_t w = midx ? mw[ci] : 0xffffffffu;
                const bool allowed = (w & lanebit) != 0u;
                pair_step<EXACT || EMIT, EMIT>(T, s_xi[warp][il], s_pi[warp][il], xj, pj, allowed, rare,
                                         fi[ci], fj, en, cnt);
            }
        }
        if (EXACT || EMIT) {
            if (__any_sync(0xffffffffu, rare))
                fix_rare_pairs(T, V, pos_all, f1acc, s_xi[warp], s_pi[warp], ibase, jslot, xj, pj,
                               imask, midx, lane, &en, &cnt, exact != 0, EMIT, emit, emit_counter, emit_pairs,
                               emit_cap);
        }
        // j force: reduce over ti (lanes with equal tj), lanes ti = 0,1,2 write x,y,z
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            fj.fx += __shfl_xor_sync(0xffffffffu, fj.fx, o);
            fj.fy += __shfl_xor_sync(0xffffffffu, fj.fy, o);
            fj.fz += __shfl_xor_sync(0xffffffffu, fj.fz, o);
        }
        if (ti < 3) {
            const float v = ti == 0 ? fj.fx : (ti == 1 ? fj.fy : fj.fz);
            if (v != 0.f)
                atomic_add_fixed(f1acc + (size_t)ti * plane + jslot, __float2ll_rn(v * 4294967296.0f));
        }
    }

    // i forces: reduce over tj, lane (tj, ti) writes clusters tj and tj+4
#pragma unroll
    for (int ci = 0; ci < nbl::kMaxCi; ci++) {
        float x = fi[ci].fx, y = fi[ci].fy, z = fi[ci].fz;
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        y += __shfl_xor_sync(0xffffffffu, y, 8);
        z += __shfl_xor_sync(0xffffffffu, z, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        y += __shfl_xor_sync(0xffffffffu, y, 16);
        z += __shfl_xor_sync(0xffffffffu, z, 16);
        if ((ci & 3) == tj && ci < sd.nci) {
            const int islot = ibase + ci * nbl::kClusterSize + ti;
            if (x != 0.f) atomic_add_fixed(f1acc + islot, __float2ll_rn(x * 4294967296.0f));
            if (y != 0.f) atomic_add_fixed(f1acc + plane + islot, __float2ll_rn(y * 4294967296.0f));
            if (z != 0.f) atomic_add_fixed(f1acc + 2 * plane + islot, __float2ll_rn(z * 4294967296.0f));
        }
    }

    // energy / count partials of this unit (fixed-order warp tree)
    double de = (double)en;
    long long dc = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_down_sync(0xffffffffu, de, o);
        dc += __shfl_down_sync(0xffffffffu, dc, o);
    }
    if (lane == 0) {
        epart[unit] = de;
        cpart[unit] = dc;
    }
}

// ---------------------------------------------------------------------------------------------
// refresh: sorted float positions from the current double positions, keeping the periodic image
// chosen at build time; raises SDM_ERR_STALE_LIST when an atom moved more than skin/2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
refresh_kernel(Topology T, nbl::Grid G, int nslot, const double* __restrict__ pos_all,
               const int* __restrict__ atom, const int* __restrict__ img,
               const float4* __restrict__ posq_build, float4* __restrict__ posq, float half_skin2,
               int* flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslot) return;
    const int ga = atom[s];
    if (ga < 0) return;  // dummy slot keeps its far-away coordinates
    const int r = ga / T.n;
    const double* p = pos_all + 3 * (size_t)ga;  // ga = r*n + a
    const int im = img[s];
    const int ix = (im & 0x3ff) - 512, iy = ((im >> 10) & 0x3ff) - 512, iz = ((im >> 20) & 0x3ff) - 512;
    double x = p[0], y = p[1], z = p[2];
    if (G.periodic) {
        x += ix * G.box[0];
        y += iy * G.box[1];
        z += iz * G.box[2];
    }
    const float4 b = posq_build[s];
    const float fx = (float)x, fy = (float)y, fz = (float)z;
    const float dx = fx - b.x, dy = fy - b.y, dz = fz - b.z;
    if (dx * dx + dy * dy + dz * dz > half_skin2) atomicExch(flags + r, SDM_ERR_STALE_LIST);
    posq[s] = make_float4(fx, fy, fz, b.w);
}

}  // namespace

void launch_pair_cluster(const Topology& T, const PairListView& V, const double* pos_all,
                         long long* f1acc, double* epart, long long* cpart, int exact,
                         int* emit_counter, int* emit_pairs, int emit_cap, int emit_replica,
                         cudaStream_t s) {
    if (V.nunits <= 0) return;
    const int grid = (V.nunits + kWarps - 1) / kWarps;
    const bool periodic = T.method == SDM_CUTOFF_PERIODIC;
#define SDM_LAUNCH(P, X, E, TT)                                                                  \
    pair_cluster_kernel<P, X, E><<<grid, kWarps * 32, 0, s>>>(TT, V, pos_all, f1acc, epart, cpart, \
                                                             exact, emit_counter, emit_pairs,     \
                                                             emit_cap, emit_replica)
    if (emit_pairs) {
        // debug pass: every in-range pair takes the out-of-line path, which also records it
        if (periodic) SDM_LAUNCH(true, true, true, T);
        else SDM_LAUNCH(false, true, true, T);
    } else if (exact) {
        if (periodic) SDM_LAUNCH(true, true, false, T);
        else SDM_LAUNCH(false, true, false, T);
    } else {
        if (periodic) SDM_LAUNCH(true, false, false, T);
        else SDM_LAUNCH(false, false, false, T);
    }
#undef SDM_LAUNCH
}

void launch_refresh(const Topology& T, const nbl::Grid& G, int nslot, const double* pos_all,
                    const int* atom, const int* img, const float4* posq_build, float4* posq,
                    float half_skin2, int* flags, cudaStream_t s) {
    if (nslot <= 0) return;
    refresh_kernel<<<(nslot + 255) / 256, 256, 0, s>>>(T, G, nslot, pos_all, atom, img, posq_build,
                                                      posq, half_skin2, flags);
}

}  // namespace sdm
