// kernels_rows.cu -- the hot kernel of the fused path: cutoff LJ + reaction-field pair interactions
// over per-atom j rows (sm_100a, FP32 SIMT pipe; not a dense contraction, so no tensor cores).
//
// One warp per work unit = an i-group of NI = 8 atoms (one cluster; 16 = two clusters is kept as a
// measured alternative) and a chunk of its ROW of individual j-atoms (nblist_core.h stage 5).  Lane =
// j-atom: every step the warp takes the next 32 row entries, each lane fetches its own j-atom
// (position + image shift, parameters) and evaluates it against all NI i-atoms, two at a time in the
// halves of packed f32x2 registers: the whole pair term is FADD2 / FMUL2 / FFMA2 work (35 packed FP32
// instructions + 2 MUFU per two pairs, 24 when the j-atom has no Lennard-Jones term).  The i-atoms are
// staged once per unit in shared memory as NI/2 (even, odd) pairs and read by broadcast LDS.128.
//   * i forces: NI*3 packed partial sums per lane for the whole unit, transpose-reduced over the
//     32 lanes once per unit;
//   * j force: complete in the lane after the NI/2 tile steps -- no shuffle -- and added with three
//     64-bit fixed-point REDs (row entries of one j-cluster sit in neighbouring lanes, so a warp
//     RED touches few sectors);
//   * row entries two steps ahead and j-atom data one step ahead are requested before the current
//     step's arithmetic, so the gathers overlap the FP32 work of the same warp;
//   * a row is ordered [entries with an allow word (exclusions, the cluster against itself) | plain |
//     plain without Lennard-Jones (epsilon_j == 0: water hydrogens)]: only the first steps take the
//     masked path and load allow words, the last ones skip eleven of the 35 packed instructions.
// Lanes past the end of the row hold a far-away dummy atom.
//
// Why rows of atoms and not 8 x 8 cluster tiles: a j-atom that is out of reach of all i-atoms of the
// group costs nothing, so half of the lane pairs the kernel evaluates are inside the cutoff (35 % with
// tiles); and with a j-atom per lane there is no per-tile mask, no REDUX and no
// j-force shuffle reduction -- the per-entry work that took 45 % of the tile kernel's time.
//
// Forces go to 64-bit fixed-point accumulators (2^32), so the result is independent of the order in
// which warps finish: bit-reproducible across runs, replicas-per-GPU and GPUs.
//
// Exact cutoff: the hot loop decides r^2 <= rc^2 in FP32 and tracks min |r^2 - rc^2| per lane and
// step; a lane that saw it inside the FP32 uncertainty band re-decides its NI pairs of that step in
// FP64 after the loop, exactly like the oracle, and applies +/- corrections, so the in-cutoff pair set
// is bit-identical to a double-precision evaluation.  The debug build (template flag EMIT) records
// every pair the kernel accepts (sdm_get_pairs).
//
// Arithmetic restated from OpenMM 7.3 ReferenceLJCoulombIxn::calculateOneIxn (SURVEY.md
// Appendix B.3) in FP32: per-atom sigma/2 and 2*sqrt(eps), charges pre-scaled by
// sqrt(ONE_4PI_EPS0), reaction field krf/crf, LJ not shifted.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "f32x2.cuh"
#include "pairlist.h"

namespace sdm {
namespace {

#ifndef SDM_ROW_MADWIDE
#define SDM_ROW_MADWIDE 1   // gather / accumulator addresses as one mad.wide.u32
#endif
#ifndef SDM_ROW_RED1
#define SDM_ROW_RED1 1      // one non-zero test for the three j-force REDs of a step
#endif
#ifndef SDM_ROW_ROTATE_LATE
#define SDM_ROW_ROTATE_LATE 1
#endif
#ifndef SDM_ROW_LDG256
#define SDM_ROW_LDG256 1      // the 32-byte j record with one 256-bit load (sm_100: LDG.E.256)
#endif
#ifndef SDM_ROW_SHIFT_VOTE
#define SDM_ROW_SHIFT_VOTE 1
#endif
#ifndef SDM_PAIR_WARPS
#define SDM_PAIR_WARPS 1
#endif
#ifndef SDM_PAIR_SEL2
#define SDM_PAIR_SEL2 1
#endif
// One warp per block: a warp that finishes its unit frees its slot at once (units differ in
// length), which keeps the achieved occupancy at the register-limited maximum.
constexpr int kWarps = SDM_PAIR_WARPS;
constexpr float kFix = 4294967296.0f;  // 2^32

struct Acc2 {
    f2 x, y, z;
};

// Staged i-atom pair (atoms 2p and 2p+1 of the i-group), 48 bytes: three LDS.128.
struct __align__(16) IPair {
    f2 x, y;      // (x_lo, x_hi), (y_lo, y_hi)
    f2 z, q;      // (z_lo, z_hi), (q_lo, q_hi)
    f2 s, e;      // sigma/2 and 2*sqrt(eps) pairs
};

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RSQ, no denormal fix-up
    return y;
}

// r^2 of an (i, j) pair with a fixed operation order: the packed hot loop, the fix-up path and
// the debug pair dump must all see the same FP32 value (fma.rn.f32x2 rounds each half exactly
// like fma.rn.f32).
__device__ __forceinline__ float pair_r2(const float xi, const float yi, const float zi,
                                         const float4 xj, float& dx, float& dy, float& dz) {
    dx = __fsub_rn(xi, xj.x);
    dy = __fsub_rn(yi, xj.y);
    dz = __fsub_rn(zi, xj.z);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// (c_lo ? lo(v) : 0, c_hi ? hi(v) : 0) as two FSEL writing one register pair
__device__ __forceinline__ f2 sel2_or_zero(const f2 v, const bool c_lo, const bool c_hi) {
    f2 r;
    asm("{\n .reg .f32 a, b;\n .reg .pred p, q;\n mov.b64 {a, b}, %1;\n setp.ne.s32 p, %2, 0;\n"
        " setp.ne.s32 q, %3, 0;\n selp.f32 a, a, 0f00000000, p;\n selp.f32 b, b, 0f00000000, q;\n"
        " mov.b64 %0, {a, b};\n}" : "=l"(r) : "l"(v), "r"((int)c_lo), "r"((int)c_hi));
    return r;
}

// The 32-byte record of a slot (position + charge, sigma/2, 2*sqrt(eps)): two 128-bit loads from one sector.
__device__ __forceinline__ void load_jrec(const float4* __restrict__ jrec, const uint32_t slot, float4& xj, float2& pj) {
#if SDM_ROW_MADWIDE
    const float4* r;   // base + slot * 32 as ONE multiply-add (the compiler's rendering takes a shift, a mask and a 64-bit add)
    asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(r) : "r"(slot), "l"(jrec));
#else
    const float4* r = jrec + 2 * (size_t)slot;
#endif
#if SDM_ROW_LDG256
    float u0, u1;   // padding words of the record
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(xj.x), "=f"(xj.y), "=f"(xj.z), "=f"(xj.w), "=f"(pj.x), "=f"(pj.y), "=f"(u0), "=f"(u1) : "l"(r));
#else
    xj = r[0];
    const float4 p = r[1];
    pj = make_float2(p.x, p.y);
#endif
}

struct PairConsts {
    float rc2, krf, crf, band;
    float alpha;   // Ewald splitting parameter (EWALD instantiations)
    int geom;      // 1: sigma_ij^2 = sigma_i sigma_j (SDM_LJ_GEOMETRIC; GEOM instantiations)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erfc(x) for x >= 0 by Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7; the polynomial OpenMM's GPU platforms
// use for the direct-space Ewald term): returns erfc, and exp(-x^2) through `expo`.
constexpr float kErfcP = 0.3275911f, kErfcA1 = 0.254829592f, kErfcA2 = -0.284496736f, kErfcA3 = 1.421413741f,
                kErfcA4 = -1.453152027f, kErfcA5 = 1.061405429f, kLog2e = 1.4426950408889634f,
                kTwoOverSqrtPi = 1.1283791670955126f;
__device__ __forceinline__ float erfc_approx(const float x, float& expo) {
    expo = ex2_approx(-kLog2e * x * x);
    const float t = rcp_approx(fmaf(kErfcP, x, 1.f));
    return ((((kErfcA5 * t + kErfcA4) * t + kErfcA3) * t + kErfcA2) * t + kErfcA1) * t * expo;
}
__device__ __forceinline__ f2 erfc_approx2(const f2 x, f2& expo) {
    const f2 a = mul2(mul2(x, x), bc(-kLog2e));
    expo = pk(ex2_approx(lo(a)), ex2_approx(hi(a)));
    const f2 d = fma2(x, bc(kErfcP), bc(1.f));
    const f2 t = pk(rcp_approx(lo(d)), rcp_approx(hi(d)));
    f2 p = fma2(bc(kErfcA5), t, bc(kErfcA4));
    p = fma2(p, t, bc(kErfcA3));
    p = fma2(p, t, bc(kErfcA2));
    p = fma2(p, t, bc(kErfcA1));
    return mul2(mul2(p, t), expo);
}

// Debug build of the hot kernel (template flag EMIT): every pair the kernel ACCEPTS -- in the tile
// loop and in the band fix-up -- is also recorded with its System indices, so
// sdm_get_pairs() returns the hot kernel's own decisions.  All of it compiles away when !EMIT.
struct EmitCtx {
    PairEmit em;
    const int* atom;   // slot -> replica*n + atom
    int n;
};

__device__ __forceinline__ void emit_pair(const EmitCtx& ec, const int islot, const int jslot) {
    const int gi = ec.atom[islot], gj = ec.atom[jslot];
    if (gi < 0 || gj < 0 || gi / ec.n != ec.em.replica) return;
    const int a = gi % ec.n, b = gj % ec.n;
    const int k = atomicAdd(ec.em.counter, 1);
    if (k < ec.em.cap) {
        ec.em.pairs[2 * k] = a < b ? a : b;
        ec.em.pairs[2 * k + 1] = a < b ? b : a;
    }
}

// Scalar FP32 pair term (fix-up path only); returns fs with F_i += fs*d, F_j -= fs*d.
__device__ __forceinline__ float pair_term_f32(const float r2, const float qi, const float si,
                                               const float ei, const float qj, const float2 pj,
                                               const PairConsts& K, float& e) {
    const float rinv = rsqrt_approx(r2);
    const float rinv2 = rinv * rinv;
    const float sig = si + pj.x;
    const float sr2 = (K.geom ? si * pj.x : sig * sig) * rinv2;   // geometric rule: the parameter is sigma itself
    const float sr6 = sr2 * sr2 * sr2;
    const float elj = (ei * pj.y) * sr6;
    const float qq = qi * qj;
    const float a = elj * sr6;
    const float e_lj = a - elj;
    if (K.alpha > 0.f) {   // direct-space Ewald (EWALD instantiations set alpha)
        float expo;
        const float ar = K.alpha * (r2 * rinv);
        const float ec = erfc_approx(ar, expo);
        const float qr = qq * rinv;
        e = fmaf(qr, ec, e_lj);
        return fmaf(a + e_lj, 6.f, qr * fmaf(ar * expo, kTwoOverSqrtPi, ec)) * rinv2;
    }
    const float kr2 = K.krf * r2;
    const float dEdR = fmaf(a + e_lj, 6.f, qq * fmaf(-2.f, kr2, rinv));
    e = fmaf(qq, (rinv + kr2) - K.crf, e_lj);
    return dEdR * rinv2;
}

// One tile step of the hot loop: the two i-atoms of a staged pair against the lane's j atom.
//   dE/dr * r  and energy (OpenMM 7.3 ReferenceLJCoulombIxn, reaction field, LJ not shifted):
//   e_lj = elj*(sr6 - 1) = a - elj,   elj*(12*sr6 - 6) = 6*(a + e_lj)   with a = elj*sr6
// LJ = false: the j-atom has no Lennard-Jones term (epsilon_j == 0: elj, a and e_lj are exactly
// zero), so the eleven packed instructions that would compute them are left out -- same bits out.
template <bool MASKED, bool EXACT, bool EMIT, int HI_OFF, bool LJ = true, bool EWALD = false, bool GEOM = false>
__device__ __forceinline__ void tile_step(const IPair* __restrict__ ip, const float4 xj,
                                          const float2 pj, const bool allow_lo, const bool allow_hi,
                                          const PairConsts& K, Acc2& fi, Acc2& fj, f2& en, int& cnt,
                                          float& tmin, const EmitCtx* ec, const int islot_lo,
                                          const int jslot) {
    const ulonglong2 a0 = *reinterpret_cast<const ulonglong2*>(&ip->x);
    const ulonglong2 a1 = *reinterpret_cast<const ulonglong2*>(&ip->z);
    const ulonglong2 a2 = *reinterpret_cast<const ulonglong2*>(&ip->s);
    const f2 dx = sub2(a0.x, bc(xj.x));
    const f2 dy = sub2(a0.y, bc(xj.y));
    const f2 dz = sub2(a1.x, bc(xj.z));
    const f2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    const f2 t = sub2(r2, bc(K.rc2));
    const float t_lo = lo(t), t_hi = hi(t);
    const bool in_lo = MASKED ? (allow_lo && t_lo <= 0.f) : (t_lo <= 0.f);
    const bool in_hi = MASKED ? (allow_hi && t_hi <= 0.f) : (t_hi <= 0.f);
    if (EXACT) tmin = fminf(tmin, fminf(fabsf(t_lo), fabsf(t_hi)));
    if (EMIT) {   // pairs inside the band are recorded by the fix-up path, which re-decides them
        if (in_lo && !(EXACT && fabsf(t_lo) < K.band)) emit_pair(*ec, islot_lo, jslot);
        if (in_hi && !(EXACT && fabsf(t_hi) < K.band)) emit_pair(*ec, islot_lo + HI_OFF, jslot);
    }
    const f2 rinv = pk(rsqrt_approx(lo(r2)), rsqrt_approx(hi(r2)));
    const f2 rinv2 = mul2(rinv, rinv);
    const f2 qq = mul2(a1.y, bc(xj.w));
    // Coulomb part: (ce, cd) = energy and dE/dr*r per unit charge product -- reaction field
    // (rinv + krf r^2 - crf, rinv - 2 krf r^2) or direct-space Ewald (erfc(alpha r)/r,
    // (erfc(alpha r) + 2 alpha r exp(-alpha^2 r^2)/sqrt(pi))/r)
    f2 ce, cd;
    if (EWALD) {
        f2 expo;
        const f2 ar = mul2(mul2(r2, rinv), bc(K.alpha));
        const f2 ec = erfc_approx2(ar, expo);
        ce = mul2(rinv, ec);
        cd = mul2(rinv, fma2(mul2(ar, expo), bc(kTwoOverSqrtPi), ec));
    } else {
        const f2 kr2 = mul2(r2, bc(K.krf));
        ce = sub2(add2(rinv, kr2), bc(K.crf));
        cd = fma2(kr2, bc(-2.f), rinv);
    }
    f2 dEdR, e;
    if (LJ) {
        // Lorentz-Berthelot: the parameters are sigma/2 and sigma_ij their sum; geometric rule (the OPLS
        // CustomNonbondedForce of desmonddmsfile75.py:781): the parameters are sigma and sigma_ij^2 their product
        const f2 sig = add2(a2.x, bc(pj.x));
        const f2 sr2 = mul2(GEOM ? mul2(a2.x, bc(pj.x)) : mul2(sig, sig), rinv2);
        const f2 sr6 = mul2(mul2(sr2, sr2), sr2);
        const f2 elj = mul2(mul2(a2.y, bc(pj.y)), sr6);
        const f2 a = mul2(elj, sr6);
        const f2 e_lj = sub2(a, elj);
        dEdR = fma2(add2(a, e_lj), bc(6.f), mul2(qq, cd));
        e = fma2(qq, ce, e_lj);
    } else {
        dEdR = mul2(qq, cd);
        e = mul2(qq, ce);
    }
    const f2 fsr = mul2(dEdR, rinv2);
#if SDM_PAIR_SEL2
    const f2 fs = sel2_or_zero(fsr, in_lo, in_hi);
    en = add2(en, sel2_or_zero(e, in_lo, in_hi));
#else
    const f2 fs = pk(in_lo ? lo(fsr) : 0.f, in_hi ? hi(fsr) : 0.f);
    en = add2(en, pk(in_lo ? lo(e) : 0.f, in_hi ? hi(e) : 0.f));
#endif
    // cnt += in as ONE predicated add per half (the compiler's own rendering takes three
    // instructions); ptxas merges the setp with the one that feeds the selects above
    if (MASKED) {
        asm("{\n .reg .pred p, q;\n setp.ne.s32 q, %2, 0;\n setp.le.and.f32 p, %1, 0f00000000, q;\n"
            " @p add.s32 %0, %0, 1;\n}" : "+r"(cnt) : "f"(t_lo), "r"((int)allow_lo));
        asm("{\n .reg .pred p, q;\n setp.ne.s32 q, %2, 0;\n setp.le.and.f32 p, %1, 0f00000000, q;\n"
            " @p add.s32 %0, %0, 1;\n}" : "+r"(cnt) : "f"(t_hi), "r"((int)allow_hi));
    } else {
        asm("{\n .reg .pred p;\n setp.le.f32 p, %1, 0f00000000;\n @p add.s32 %0, %0, 1;\n}"
            : "+r"(cnt) : "f"(t_lo));
        asm("{\n .reg .pred p;\n setp.le.f32 p, %1, 0f00000000;\n @p add.s32 %0, %0, 1;\n}"
            : "+r"(cnt) : "f"(t_hi));
    }
    fi.x = fma2(fs, dx, fi.x); fi.y = fma2(fs, dy, fi.y); fi.z = fma2(fs, dz, fi.z);
    // the j force is accumulated with the i sign and negated once per entry
    fj.x = fma2(fs, dx, fj.x); fj.y = fma2(fs, dy, fj.y); fj.z = fma2(fs, dz, fj.z);
}

// add v = f * scale (fixed point) to *p unless f == 0: one predicated RED, no branch
__device__ __forceinline__ void red_fixed_nonzero(long long* p, const float f, const float scale) {
    const long long v = __float2ll_rn(f * scale);
    asm volatile("{\n .reg .pred p;\n setp.neu.f32 p, %2, 0f00000000;\n @p red.global.add.u64 [%0], %1;\n}"
                 :: "l"(p), "l"(v), "f"(f) : "memory");
}

#ifndef SDM_ROW_MINB
#define SDM_ROW_MINB 24
#endif
#ifndef SDM_ROW_JPREFETCH
#define SDM_ROW_JPREFETCH 1   // 1: j-atom data one step ahead (row entries two ahead); 0: entries one ahead only
#endif

// Rare path: the lane's pairs of one step whose FP32 r^2 lies within K.band of the cutoff are
// re-decided with the FP64 in-cutoff test of the oracle; where that differs from the FP32 decision
// of the hot loop the pair's force / energy / count is added or taken back (forces straight to the
// fixed-point accumulators of both atoms).  Out of line so that the hot loop stays small.
template <int NI, bool EMIT, bool EWALD>
__device__ __noinline__ void fix_band_row(const Topology& T, const PairListView& V,
                                          const double* __restrict__ pos_all,
                                          long long* __restrict__ f1acc, const IPair* s_ip, int ibase,
                                          int jslot, float4 xj, float2 pj, uint32_t allow, float* en,
                                          int* cnt, const EmitCtx* ec) {
    const size_t plane = (size_t)V.nslot_cap;
    const PairConsts K{T.rc2f, T.krff, T.crff, T.band, EWALD ? T.alphaf : 0.f, T.lj_geom};
    const int aj = V.atom[jslot];
    if (aj < 0) return;
    for (int a = 0; a < NI; a++) {
        if (!((allow >> a) & 1u)) continue;
        const IPair ip = s_ip[a >> 1];
        const int h = a & 1;
        const float xi = h ? hi(ip.x) : lo(ip.x), yi = h ? hi(ip.y) : lo(ip.y);
        const float zi = h ? hi(ip.z) : lo(ip.z), qi = h ? hi(ip.q) : lo(ip.q);
        const float si = h ? hi(ip.s) : lo(ip.s), ei = h ? hi(ip.e) : lo(ip.e);
        float dx, dy, dz;
        const float r2 = pair_r2(xi, yi, zi, xj, dx, dy, dz);
        const float t = r2 - K.rc2;
        if (!(fabsf(t) < K.band)) continue;
        const int islot = ibase + a;
        const int ai = V.atom[islot];
        if (ai < 0) continue;
        const int r = ai / T.n;
        const bool in64 = in_cutoff_f64(T, pos_all + (size_t)r * 3 * T.n, ai - r * T.n, aj - r * T.n);
        const bool in32 = t <= 0.f;
        if (EMIT && in64) emit_pair(*ec, islot, jslot);
        if (in64 == in32) continue;
        const float sgn = in64 ? 1.f : -1.f;
        float e;
        const float fs = sgn * pair_term_f32(r2, qi, si, ei, xj.w, pj, K, e);
        *en += sgn * e;
        *cnt += in64 ? 1 : -1;
        const float f[3] = {fs * dx, fs * dy, fs * dz};
        for (int c = 0; c < 3; c++) {
            const long long v = __float2ll_rn(f[c] * kFix);
            atomic_add_fixed(f1acc + (size_t)c * plane + islot, v);
            atomic_add_fixed(f1acc + (size_t)c * plane + jslot, -v);
        }
    }
}

// Sum of v[0..N) over the 32 lanes by halving: after the exchange with lane^W every lane keeps the
// half of the values its bit selects.  Ends with 3 values per lane (one atom's x, y, z), which the
// remaining lanes of the atom's group share through a butterfly.
template <int N, int W>
__device__ __forceinline__ void row_halve(float (&v)[N], const int lane) {
    const bool up = (lane & W) != 0;
#pragma unroll
    for (int k = 0; k < N / 2; k++) {
        const float send = up ? v[k] : v[k + N / 2];
        v[k] = (up ? v[k + N / 2] : v[k]) + __shfl_xor_sync(0xffffffffu, send, W);
    }
}

template <int NI, bool PERIODIC, bool EXACT, bool EMIT, bool EWALD, bool GEOM>
__device__ __forceinline__ void process_row_unit(const Topology& T, const PairListView& V,
                                                 const double* __restrict__ pos_all,
                                                 long long* __restrict__ f1acc, double* __restrict__ epart,
                                                 long long* __restrict__ cpart, const int unit,
                                                 const RowUnit u, const int lane, IPair* s_ip,
                                                 const float4* s_shift, const EmitCtx* ec) {
    constexpr int NP = NI / 2;
    const int ibase = (u.c0n & 0xfffffff) * nbl::kClusterSize;
    const int ni = (u.c0n >> 28) * nbl::kClusterSize;
    const PairConsts K{T.rc2f, T.krff, T.crff, T.band, EWALD ? T.alphaf : 0.f, T.lj_geom};
    const uint32_t dummy_ent = (uint32_t)V.dummy_slot | (nbl::kShiftZero << 26);

    // row entries two steps ahead, j-atom data one step ahead
    const int nsteps = (u.end - u.begin + 31) >> 5;
    const int mend = u.begin + (u.seg & 0xffff);             // [begin, mend): entries with an allow word
    const int msteps = ((u.seg & 0xffff) + 31) >> 5;         // steps that hold masked entries
    const int lsteps = ((u.seg >> 16) + 31) >> 5;            // steps that hold j-atoms with a Lennard-Jones term
    int idx = u.begin + lane;
    uint32_t ent1 = idx < u.end ? V.jent[idx] : dummy_ent;
    uint32_t ent2 = idx + 32 < u.end ? V.jent[idx + 32] : dummy_ent;

    // stage the i-atoms: atom a goes to half (a & 1) of pair a >> 1
    if (lane < NI) {
        float4 q = make_float4(-nbl::kFar, -nbl::kFar, -nbl::kFar, 0.f);
        float2 pr = make_float2(0.f, 0.f);
        if (lane < ni) {
            float4 g;
            float2 gp;
            load_jrec(V.jrec, (uint32_t)(ibase + lane), g, gp);
            if (g.x < 0.5f * nbl::kFar) { q = g; pr = gp; }
        }
        float* dst = reinterpret_cast<float*>(s_ip + (lane >> 1)) + (lane & 1);
        dst[0] = q.x; dst[2] = q.y; dst[4] = q.z; dst[6] = q.w; dst[8] = pr.x; dst[10] = pr.y;
    }
#if SDM_ROW_JPREFETCH
    float4 xj1;
    float2 pj1;
    load_jrec(V.jrec, ent1 & 0x3ffffffu, xj1, pj1);
#endif
    __syncwarp();

    Acc2 fi[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) fi[p] = Acc2{0ull, 0ull, 0ull};
    f2 en = 0ull;
    int cnt = 0;
    uint32_t fixmask = 0u;
    const size_t plane = (size_t)V.nslot_cap;

    for (int k = 0; k < nsteps; k++, idx += 32) {
        const uint32_t ent = ent1;
#if SDM_ROW_JPREFETCH
        float4 xj = xj1;
        const float2 pj = pj1;
        ent1 = ent2;
        ent2 = idx + 64 < u.end ? V.jent[idx + 64] : dummy_ent;
#if SDM_ROW_ROTATE_LATE
        // the next step's record lands in its own registers and is handed over at the END of this step
        // (volatile moves after the REDs): left to itself the compiler copies the landing registers into
        // the loop-carried ones right after issuing the load and waits for it there -- no prefetch at all
        float4 xjn;
        float2 pjn;
        load_jrec(V.jrec, ent1 & 0x3ffffffu, xjn, pjn);
#else
        load_jrec(V.jrec, ent1 & 0x3ffffffu, xj1, pj1);
#endif
#else
        float4 xj;
        float2 pj;
        load_jrec(V.jrec, ent & 0x3ffffffu, xj, pj);
        ent1 = ent2;
        ent2 = idx + 64 < u.end ? V.jent[idx + 64] : dummy_ent;
#endif
        const int jslot = (int)(ent & 0x3ffffffu);
        if (PERIODIC) {
#if SDM_ROW_SHIFT_VOTE
            // most steps hold central-image atoms only: one vote instead of the table look-up
            if (__any_sync(0xffffffffu, (ent >> 26) != nbl::kShiftZero))
#endif
            {
                const float4 sh = s_shift[ent >> 26];
                xj.x += sh.x; xj.y += sh.y; xj.z += sh.z;
            }
        }
        Acc2 fj{0ull, 0ull, 0ull};
        float tmin = 3.0e38f;
        if (k < msteps) {
            const uint32_t allow = idx < mend ? (uint32_t)V.jallow[idx] : 0xffffu;
#pragma unroll
            for (int p = 0; p < NP; p++)
                tile_step<true, EXACT, EMIT, 1, true, EWALD, GEOM>(s_ip + p, xj, pj, ((allow >> (2 * p)) & 1u) != 0u,
                                                ((allow >> (2 * p + 1)) & 1u) != 0u, K, fi[p], fj, en, cnt,
                                                tmin, ec, ibase + 2 * p, jslot);
        } else if (k < lsteps) {
#pragma unroll
            for (int p = 0; p < NP; p++)
                tile_step<false, EXACT, EMIT, 1, true, EWALD, GEOM>(s_ip + p, xj, pj, true, true, K, fi[p], fj, en, cnt, tmin, ec,
                                                 ibase + 2 * p, jslot);
        } else {   // the tail of the row: j-atoms without a Lennard-Jones term (water hydrogens)
#pragma unroll
            for (int p = 0; p < NP; p++)
                tile_step<false, EXACT, EMIT, 1, false, EWALD>(s_ip + p, xj, pj, true, true, K, fi[p], fj, en, cnt, tmin,
                                                        ec, ibase + 2 * p, jslot);
        }
        if (EXACT) fixmask |= (tmin < K.band ? 1u : 0u) << k;
        // j force: complete in this lane (sign: F_j = -sum)
        // a j-atom of the row that has no partner inside the cutoff (one in six: the list reaches to
        // rc + skin) and the padding lanes stay silent; ptxas renders each predicated RED as a short
        // branch region that also skips the 64-bit conversion
#if SDM_ROW_MADWIDE
        long long* fjp;
        asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(fjp) : "r"((uint32_t)jslot), "l"(f1acc));
#else
        long long* fjp = f1acc + jslot;
#endif
#if SDM_ROW_RED1
        {   // one test for the three components: a j-atom out of reach of all eight i-atoms has an exactly zero force
            const float fx = lo(fj.x) + hi(fj.x), fy = lo(fj.y) + hi(fj.y), fz = lo(fj.z) + hi(fj.z);
            if (fx != 0.f || fy != 0.f || fz != 0.f) {
                const long long vx = __float2ll_rn(fx * -kFix), vy = __float2ll_rn(fy * -kFix), vz = __float2ll_rn(fz * -kFix);
                asm volatile("red.global.add.u64 [%0], %1;" :: "l"(fjp), "l"(vx) : "memory");
                asm volatile("red.global.add.u64 [%0], %1;" :: "l"(fjp + plane), "l"(vy) : "memory");
                asm volatile("red.global.add.u64 [%0], %1;" :: "l"(fjp + 2 * plane), "l"(vz) : "memory");
            }
        }
#else
        red_fixed_nonzero(fjp, lo(fj.x) + hi(fj.x), -kFix);
        red_fixed_nonzero(fjp + plane, lo(fj.y) + hi(fj.y), -kFix);
        red_fixed_nonzero(fjp + 2 * plane, lo(fj.z) + hi(fj.z), -kFix);
#endif
#if SDM_ROW_JPREFETCH && SDM_ROW_ROTATE_LATE
        asm volatile("mov.b32 %0, %6;\n mov.b32 %1, %7;\n mov.b32 %2, %8;\n mov.b32 %3, %9;\n mov.b32 %4, %10;\n mov.b32 %5, %11;"
                     : "=f"(xj1.x), "=f"(xj1.y), "=f"(xj1.z), "=f"(xj1.w), "=f"(pj1.x), "=f"(pj1.y)
                     : "f"(xjn.x), "f"(xjn.y), "f"(xjn.z), "f"(xjn.w), "f"(pjn.x), "f"(pjn.y));
#endif
    }

    // i forces: v[3*a + d] of atom a, summed over the lanes
    {
        float v[3 * NI];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            v[6 * p + 0] = lo(fi[p].x); v[6 * p + 1] = lo(fi[p].y); v[6 * p + 2] = lo(fi[p].z);
            v[6 * p + 3] = hi(fi[p].x); v[6 * p + 4] = hi(fi[p].y); v[6 * p + 5] = hi(fi[p].z);
        }
        // halving stages pick the atom by the high lane bits; what is left is one atom per lane group
        if constexpr (NI == 16) {
            float a24[24], a12[12], a6[6];
            row_halve<48, 16>(v, lane);
#pragma unroll
            for (int k = 0; k < 24; k++) a24[k] = v[k];
            row_halve<24, 8>(a24, lane);
#pragma unroll
            for (int k = 0; k < 12; k++) a12[k] = a24[k];
            row_halve<12, 4>(a12, lane);
#pragma unroll
            for (int k = 0; k < 6; k++) a6[k] = a12[k];
            row_halve<6, 2>(a6, lane);
            float w[3];
#pragma unroll
            for (int d = 0; d < 3; d++) w[d] = a6[d] + __shfl_xor_sync(0xffffffffu, a6[d], 1);
            const int atom = lane >> 1;   // lane 0 of the pair writes x and y, lane 1 writes z
            if (atom < ni) {
                long long* fp = f1acc + ibase + atom;
                if ((lane & 1) == 0) {
                    red_fixed_nonzero(fp, w[0], kFix);
                    red_fixed_nonzero(fp + plane, w[1], kFix);
                } else {
                    red_fixed_nonzero(fp + 2 * plane, w[2], kFix);
                }
            }
        } else {
            float a12[12], a6[6];
            row_halve<24, 16>(v, lane);
#pragma unroll
            for (int k = 0; k < 12; k++) a12[k] = v[k];
            row_halve<12, 8>(a12, lane);
#pragma unroll
            for (int k = 0; k < 6; k++) a6[k] = a12[k];
            row_halve<6, 4>(a6, lane);
            float w[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                w[d] = a6[d] + __shfl_xor_sync(0xffffffffu, a6[d], 2);
                w[d] += __shfl_xor_sync(0xffffffffu, w[d], 1);
            }
            const int atom = lane >> 2, comp = lane & 3;
            if (atom < ni && comp < 3)
                red_fixed_nonzero(f1acc + (size_t)comp * plane + ibase + atom,
                                  comp == 0 ? w[0] : comp == 1 ? w[1] : w[2], kFix);
        }
    }

    float en1 = lo(en) + hi(en);
    if (EXACT && fixmask) {
        float en_fix = 0.f;   // separate variables: their address is taken by the call
        int cnt_fix = 0;
        while (fixmask) {
            const int k = __ffs(fixmask) - 1;
            fixmask &= fixmask - 1u;
            const int id = u.begin + 32 * k + lane;
            const uint32_t fe = V.jent[id];
            const int js = (int)(fe & 0x3ffffffu);
            float4 xj;
            float2 pjf;
            load_jrec(V.jrec, (uint32_t)js, xj, pjf);
            if (PERIODIC) {
                const float4 sh = s_shift[fe >> 26];
                xj.x += sh.x; xj.y += sh.y; xj.z += sh.z;
            }
            const uint32_t allow = id < mend ? (uint32_t)V.jallow[id] : 0xffffu;
            fix_band_row<NI, EMIT, EWALD>(T, V, pos_all, f1acc, s_ip, ibase, js, xj, pjf, allow, &en_fix, &cnt_fix, ec);
        }
        en1 += en_fix;
        cnt += cnt_fix;
    }
    __syncwarp();

    // energy / count partials of this unit (fixed-order warp tree)
    double de = (double)en1;
    int dc = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        de += __shfl_down_sync(0xffffffffu, de, o);
        dc += __shfl_down_sync(0xffffffffu, dc, o);
    }
    if (lane == 0) {
        epart[unit] = de;
        cpart[unit] = dc;
    }
    __syncwarp();   // the staging area is reused by the next unit of this warp
}

template <int NI, bool PERIODIC, bool EXACT, bool EMIT, bool EWALD = false, bool GEOM = false>
__global__ void __launch_bounds__(kWarps * 32, NI == 16 ? 16 : SDM_ROW_MINB)
pair_row_kernel(const __grid_constant__ Topology T, const __grid_constant__ PairListView V,
                const double* __restrict__ pos_all, long long* __restrict__ f1acc,
                double* __restrict__ epart, long long* __restrict__ cpart, int* unit_counter,
                const __grid_constant__ PairEmit em) {
    __shared__ IPair s_ip[kWarps][NI / 2];
    __shared__ float4 s_shift[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (PERIODIC) {
        for (uint32_t code = threadIdx.x; code < 64; code += kWarps * 32)
            s_shift[code] = make_float4((float)nbl::shift_x(code) * T.boxf[0],
                                        (float)nbl::shift_y(code) * T.boxf[1],
                                        (float)nbl::shift_z(code) * T.boxf[2], 0.f);
        __syncthreads();
    }
    EmitCtx ec_store;
    const EmitCtx* ec = nullptr;
    if (EMIT) {
        ec_store.em = em;
        ec_store.atom = V.atom;
        ec_store.n = T.n;
        ec = &ec_store;
    }
    const int nrunits = *V.nrunits;
    for (;;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(unit_counter, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= nrunits) break;
        if (V.runit_order) unit = V.runit_order[unit];   // longest units first
        process_row_unit<NI, PERIODIC, EXACT, EMIT, EWALD, GEOM>(T, V, pos_all, f1acc, epart, cpart, unit, V.runits[unit], lane,
                                                    s_ip[warp], s_shift, ec);
    }
}

// ---------------------------------------------------------------------------------------------
// refresh: sorted float positions from the current double positions, keeping the periodic image
// chosen at build time; raises SDM_ERR_STALE_LIST when an atom moved more than skin/2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
refresh_kernel(Topology T, nbl::Grid G, const int* __restrict__ d_nslot, const double* __restrict__ pos_all,
               const int* __restrict__ atom, const int* __restrict__ img,
               const float4* __restrict__ posq_build, float4* __restrict__ posq, float4* __restrict__ jrec,
               float half_skin2, int* flags, int* list_age, unsigned int* max_disp2,
               long long* __restrict__ acc, int nslot_cap) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) *list_age += 1;   // one more evaluation with this list (read by the scalar stage)
    // fresh state-1 accumulators for the pair pass that follows (replaces an 8 MB memset node; acc == nullptr
    // when the mix kernel of the previous evaluation has cleared them already)
    if (acc && s < nslot_cap) {
        acc[s] = 0;
        acc[(size_t)nslot_cap + s] = 0;
        acc[2 * (size_t)nslot_cap + s] = 0;
    }
    float d2 = 0.f;
    const int ga = s < *d_nslot ? atom[s] : -1;
    if (ga >= 0) {   // a dummy slot keeps its far-away coordinates
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
        d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > half_skin2) atomicExch(flags + r, SDM_ERR_STALE_LIST);
        posq[s] = make_float4(fx, fy, fz, b.w);
        jrec[2 * (size_t)s] = make_float4(fx, fy, fz, b.w);
    }
    // largest squared displacement since the list was built, over all replicas: the host plans the
    // next rebuild from its growth (non-negative floats order like their bit patterns)
    if (max_disp2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
        if ((threadIdx.x & 31) == 0 && __float_as_uint(d2) > *max_disp2) atomicMax(max_disp2, __float_as_uint(d2));
    }
}

}  // namespace

void launch_pair_rows(const Topology& T, const PairListView& V, const double* pos_all,
                      long long* f1acc, double* epart, long long* cpart, int exact,
                      int* unit_counter, int num_sms, const PairEmit* emit, int reserve, cudaStream_t s) {
    if (V.nrunits_ub <= 0) return;
    // the counter is zero when an evaluation starts (the mix kernel of the previous one put it back);
    // the debug launch stands outside that cycle
    if (emit) cudaMemsetAsync(unit_counter, 0, sizeof(int), s);
    const bool periodic = T.method == SDM_CUTOFF_PERIODIC;
    const PairEmit em = emit ? *emit : PairEmit{nullptr, nullptr, 0, -1};
#define SDM_LAUNCH(N, P, X, E, W, G)                                                               \
    do {                                                                                          \
        static int resident = 0; /* blocks per SM the hardware keeps resident (register limited) */ \
        if (!resident) {                                                                          \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pair_row_kernel<N, P, X, E, W, G>, \
                                                              kWarps * 32, 0) != cudaSuccess ||    \
                resident < 1)                                                                     \
                resident = SDM_ROW_MINB;                                                          \
            if (const char* e_ = getenv("SDMB200_PAIR_RESIDENT")) resident = std::max(1, atoi(e_)); \
        }                                                                                         \
        const int grid = std::min((V.nrunits_ub + kWarps - 1) / kWarps, num_sms * std::max(1, resident - reserve));       \
        pair_row_kernel<N, P, X, E, W, G><<<grid, kWarps * 32, 0, s>>>(T, V, pos_all, f1acc, epart, cpart, \
                                                                   unit_counter, em);             \
    } while (0)
    // geometric combining rule: its own instantiations of the production kernel; the debug build (pair records
    // only, forces discarded) has none
    const bool geom = T.lj_geom != 0 && !emit;
#define SDM_LAUNCH_G(N, P, X, E, W)                                                               \
    do {                                                                                          \
        if (!E && geom) SDM_LAUNCH(N, P, X, false, W, true);                                      \
        else SDM_LAUNCH(N, P, X, E, W, false);                                                    \
    } while (0)
#define SDM_LAUNCH_N(P, X, E, W)                                                                  \
    do {                                                                                          \
        if (V.row_group == 2) SDM_LAUNCH_G(16, P, X, E, W);                                       \
        else SDM_LAUNCH_G(8, P, X, E, W);                                                         \
    } while (0)
    if (T.ewald) {   // direct-space Ewald / PME: periodic by definition
        if (emit) { if (exact) SDM_LAUNCH_N(true, true, true, true); else SDM_LAUNCH_N(true, false, true, true); }
        else if (exact) SDM_LAUNCH_N(true, true, false, true);
        else SDM_LAUNCH_N(true, false, false, true);
    } else if (emit) {   // debug build of the same kernel: records the accepted pairs
        if (exact) { if (periodic) SDM_LAUNCH_N(true, true, true, false); else SDM_LAUNCH_N(false, true, true, false); }
        else { if (periodic) SDM_LAUNCH_N(true, false, true, false); else SDM_LAUNCH_N(false, false, true, false); }
    } else if (exact) {
        if (periodic) SDM_LAUNCH_N(true, true, false, false);
        else SDM_LAUNCH_N(false, true, false, false);
    } else {
        if (periodic) SDM_LAUNCH_N(true, false, false, false);
        else SDM_LAUNCH_N(false, false, false, false);
    }
#undef SDM_LAUNCH_N
#undef SDM_LAUNCH_G
#undef SDM_LAUNCH
}

void launch_refresh(const Topology& T, const nbl::Grid& G, const int* d_nslot, int nslot_ub, const double* pos_all,
                    const int* atom, const int* img, const float4* posq_build, float4* posq, float4* jrec,
                    float half_skin2, int* flags, int* list_age, unsigned int* max_disp2, long long* acc_to_clear,
                    cudaStream_t s) {
    if (nslot_ub <= 0) return;
    refresh_kernel<<<(nslot_ub + 255) / 256, 256, 0, s>>>(T, G, d_nslot, pos_all, atom, img, posq_build,
                                                      posq, jrec, half_skin2, flags, list_age, max_disp2,
                                                      acc_to_clear, nslot_ub);
}

}  // namespace sdm
