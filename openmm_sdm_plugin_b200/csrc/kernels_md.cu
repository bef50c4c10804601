// kernels_md.cu -- the device-resident Langevin step with distance constraints.
//
// What is restated: ReferenceStochasticDynamicsSDM::update
// (platforms/reference/src/ReferenceStochasticDynamicsSDM.cpp:216-266): updatePart1 (:131-169,
// velocity update with noise), updatePart2 (:183-201, xPrime = x + dt*v), then
// referenceConstraintAlgorithm->apply(x, xPrime, 1/m, tolerance) (:250-252), then
// v = (xPrime - x)/dt, x = xPrime (:256-262).  The constraint algorithm itself is OpenMM's
// (ReferenceSETTLEAlgorithm for rigid three-site waters, ReferenceCCMAAlgorithm for the rest; not
// vendored): here the waters get the analytic SETTLE solution (Miyamoto & Kollman 1992, general
// masses) and every other constraint cluster (a heavy atom with its hydrogens) an in-thread SHAKE
// iteration to the same relative tolerance.  Both satisfy the same equations -- positions on the
// constraint manifold reached along the OLD bond vectors -- so they agree with the oracle's
// tightly converged solution to within the tolerance.  All FP64, like the Reference platform.
//
// Stale lists: if the scalar stage of this step's evaluation reported SDM_ERR_STALE_LIST /
// SDM_ERR_CAPACITY it raises ctl[0]; every update kernel then returns without touching the state,
// for this and all later enqueued steps, until the host (sdm_md_step) has rebuilt the list and
// repeats the steps that did not happen.  ctl[1] is the number of steps really taken, ctl[2] is
// raised by a constraint cluster that did not converge (later steps are not taken either).
#include <curand_kernel.h>

#include "sdm_kernels.h"

namespace sdm {
namespace {

constexpr int kThreads = 256;
constexpr int kShakeMaxIter = 500;

// Part 1 + 2 of the update for one atom.  Atoms of a constraint cluster only get their xPrime; all
// others are finished here (x, v):
//   v  = vscale*v + fscale*invm*F + noisescale*sqrt(invm)*xi        (updatePart1, :156-162)
//   x' = x + dt*v                                                   (updatePart2, :196-200)
//   v  = (1/dt)*(x' - x);  x = x'                                    (update, :256-262)
// in double precision with the reference's operation order and no contraction, so that without
// constraints the result is bit-identical to the reference's compiled update (tests/test_gpu_md.py).
// The normals come from `noise` when given (test hook) and from a Philox4x32-10 stream keyed by
// (seed, replica*n + atom, step) otherwise; two curand_normal2_double calls consume 8 32-bit
// outputs of the atom's subsequence per step.  Massless particles (invm == 0) do not move.
__global__ void __launch_bounds__(kThreads)
md_part12_kernel(int n, int total, double* __restrict__ pos, double* __restrict__ vel,
                 const double* __restrict__ force, const double* __restrict__ invm, double vscale,
                 double fscale, double noisescale, double dt, double inv_dt,
                 const double* __restrict__ noise, unsigned long long seed, unsigned long long step,
                 const unsigned char* __restrict__ in_cluster, double* __restrict__ xprime,
                 unsigned long long* __restrict__ ctl) {
    // a stale list this step (or a failed constraint solve before it): nothing moves until the host
    // has dealt with it.  Nothing writes ctl[0] / ctl[2] while this kernel runs.
    if (ctl && (ctl[0] | ctl[2]) != 0ull) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // replica*n + atom
    if (i == 0 && ctl) ctl[1] = step + 1ull;
    if (i >= total) return;
    const int a = i % n;
    const double im = invm[a];
    const bool clustered = in_cluster && in_cluster[a];
    if (im == 0.0) {
        if (clustered) { xprime[3 * (size_t)i] = pos[3 * (size_t)i]; xprime[3 * (size_t)i + 1] = pos[3 * (size_t)i + 1]; xprime[3 * (size_t)i + 2] = pos[3 * (size_t)i + 2]; }
        return;
    }
    const double sim = sqrt(im);
    double xi[3];
    if (noise) {
        xi[0] = noise[3 * (size_t)i]; xi[1] = noise[3 * (size_t)i + 1]; xi[2] = noise[3 * (size_t)i + 2];
    } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, 8ull * step, &st);
        const double2 g = curand_normal2_double(&st), h = curand_normal2_double(&st);
        xi[0] = g.x; xi[1] = g.y; xi[2] = h.x;
    }
    const double fi = __dmul_rn(fscale, im), ns = __dmul_rn(noisescale, sim);
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const size_t k = 3 * (size_t)i + d;
        const double x = pos[k];
        const double v = __dadd_rn(__dadd_rn(__dmul_rn(vscale, vel[k]), __dmul_rn(fi, force[k])), __dmul_rn(ns, xi[d]));
        const double xp = __dadd_rn(x, __dmul_rn(dt, v));
        if (clustered) {
            xprime[k] = xp;
        } else {
            vel[k] = __dmul_rn(inv_dt, __dsub_rn(xp, x));
            pos[k] = xp;
        }
    }
}

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// SETTLE for one rigid three-site molecule: apex atom 0 at distance d1 from atoms 1 and 2, which
// are d2 apart.  x0..x2: positions before the step (on the constraint manifold), p0..p2: the
// unconstrained new positions, overwritten with the constrained ones.
__device__ void settle(const V3 x0, const V3 x1, const V3 x2, V3& p0, V3& p1, V3& p2, const double m0,
                       const double m1, const double m2, const double d1, const double d2) {
    const V3 xp0 = p0 - x0, xp1 = p1 - x1, xp2 = p2 - x2;
    const V3 xb0 = x1 - x0, xc0 = x2 - x0;
    const double inv = 1.0 / (m0 + m1 + m2);
    const V3 xcom = (xp0 * m0 + (xb0 + xp1) * m1 + (xc0 + xp2) * m2) * inv;
    const V3 xa1 = xp0 - xcom, xb1 = xb0 + xp1 - xcom, xc1 = xc0 + xp2 - xcom;
    V3 tz = cross(xb0, xc0), tx = cross(xa1, tz), ty = cross(tz, tx);
    tx = tx * rsqrt(dot(tx, tx));
    ty = ty * rsqrt(dot(ty, ty));
    tz = tz * rsqrt(dot(tz, tz));
    const double xb0d = dot(tx, xb0), yb0d = dot(ty, xb0), xc0d = dot(tx, xc0), yc0d = dot(ty, xc0);
    const double za1d = dot(tz, xa1);
    const double xb1d = dot(tx, xb1), yb1d = dot(ty, xb1), zb1d = dot(tz, xb1);
    const double xc1d = dot(tx, xc1), yc1d = dot(ty, xc1), zc1d = dot(tz, xc1);
    const double rc = 0.5 * d2;
    const double rb = sqrt(d1 * d1 - rc * rc) * m0 * inv;
    const double ra = rb * (m1 + m2) / m0;
    const double sinphi = za1d / ra, cosphi = sqrt(1.0 - sinphi * sinphi);
    const double sinpsi = (zb1d - zc1d) / (2.0 * rc * cosphi), cospsi = sqrt(1.0 - sinpsi * sinpsi);
    const double ya2d = ra * cosphi;
    double xb2d = -rc * cospsi;
    const double yb2d = -rb * cosphi - rc * sinpsi * sinphi, yc2d = -rb * cosphi + rc * sinpsi * sinphi;
    const double xb2d2 = xb2d * xb2d;
    const double hh2 = 4.0 * xb2d2 + (yb2d - yc2d) * (yb2d - yc2d) + (zb1d - zc1d) * (zb1d - zc1d);
    xb2d -= 0.5 * (2.0 * xb2d + sqrt(4.0 * xb2d2 - hh2 + d2 * d2));
    const double alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d;
    const double beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d;
    const double gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d;
    const double al2be2 = alpha * alpha + beta * beta;
    const double sintheta = (alpha * gamma - beta * sqrt(al2be2 - gamma * gamma)) / al2be2;
    const double costheta = sqrt(1.0 - sintheta * sintheta);
    const double xa3d = -ya2d * sintheta, ya3d = ya2d * costheta, za3d = za1d;
    const double xb3d = xb2d * costheta - yb2d * sintheta, yb3d = xb2d * sintheta + yb2d * costheta, zb3d = zb1d;
    const double xc3d = -xb2d * costheta - yc2d * sintheta, yc3d = -xb2d * sintheta + yc2d * costheta, zc3d = zc1d;
    const V3 base = x0 + xcom;
    p0 = base + tx * xa3d + ty * ya3d + tz * za3d;
    p1 = base + tx * xb3d + ty * yb3d + tz * zb3d;
    p2 = base + tx * xc3d + ty * yc3d + tz * zc3d;
}

// One thread per (replica, constraint unit): units [0, n_settle) are waters, the rest SHAKE
// clusters.  Finishes the step for the unit's atoms: x'' from the constraint solve, then
// v = (x'' - x)/dt and x = x'' (ReferenceStochasticDynamicsSDM.cpp:256-262).
__global__ void __launch_bounds__(128)
md_constrain_kernel(int n, int R, MdConstraints C, double* __restrict__ pos, double* __restrict__ vel,
                    double* __restrict__ xprime, const double* __restrict__ invm, double inv_dt,
                    unsigned long long step, unsigned long long* __restrict__ ctl, int* __restrict__ flags) {
    if (ctl && ctl[1] != step + 1ull) return;   // part 1 + 2 of this step did not run
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nunits = C.n_settle + C.n_shake;
    if (t >= nunits * R) return;
    const int r = t / nunits, u = t - r * nunits;
    double* X = pos + 3 * (size_t)r * n;
    double* V = vel + 3 * (size_t)r * n;
    double* P = xprime + 3 * (size_t)r * n;
    if (u < C.n_settle) {
        const int a0 = C.settle_atoms[3 * u], a1 = C.settle_atoms[3 * u + 1], a2 = C.settle_atoms[3 * u + 2];
        const V3 x0 = ld3(X + 3 * a0), x1 = ld3(X + 3 * a1), x2 = ld3(X + 3 * a2);
        V3 p0 = ld3(P + 3 * a0), p1 = ld3(P + 3 * a1), p2 = ld3(P + 3 * a2);
        settle(x0, x1, x2, p0, p1, p2, 1.0 / invm[a0], 1.0 / invm[a1], 1.0 / invm[a2], C.settle_par[2 * u],
               C.settle_par[2 * u + 1]);
        st3(V + 3 * a0, (p0 - x0) * inv_dt); st3(X + 3 * a0, p0);
        st3(V + 3 * a1, (p1 - x1) * inv_dt); st3(X + 3 * a1, p1);
        st3(V + 3 * a2, (p2 - x2) * inv_dt); st3(X + 3 * a2, p2);
        return;
    }
    const int k = u - C.n_settle;
    const int c0 = C.shake_off[k], c1 = C.shake_off[k + 1];
    const double lo = (1.0 - C.tol) * (1.0 - C.tol), hi = (1.0 + C.tol) * (1.0 + C.tol);
    bool done = false;
    for (int it = 0; it < kShakeMaxIter && !done; it++) {
        done = true;
        for (int c = c0; c < c1; c++) {
            const int i = C.shake_ij[2 * c], j = C.shake_ij[2 * c + 1];
            const double d2 = C.shake_d[c] * C.shake_d[c];
            const V3 rp = ld3(P + 3 * i) - ld3(P + 3 * j);
            const double rp2 = dot(rp, rp);
            if (rp2 >= lo * d2 && rp2 <= hi * d2) continue;
            done = false;
            const V3 r0 = ld3(X + 3 * i) - ld3(X + 3 * j);
            const double wi = invm[i], wj = invm[j];
            const double g = (d2 - rp2) / (2.0 * dot(r0, rp) * (wi + wj));
            st3(P + 3 * i, ld3(P + 3 * i) + r0 * (g * wi));
            st3(P + 3 * j, ld3(P + 3 * j) - r0 * (g * wj));
        }
    }
    if (!done) {
        atomicExch(flags + r, SDM_ERR_CONSTRAINT);
        if (ctl) ctl[2] = 1ull;
    }
    for (int m = C.shake_aoff[k]; m < C.shake_aoff[k + 1]; m++) {
        const int a = C.shake_atoms[m];
        if (invm[a] == 0.0) continue;
        const V3 x = ld3(X + 3 * a), p = ld3(P + 3 * a);
        st3(V + 3 * a, (p - x) * inv_dt);
        st3(X + 3 * a, p);
    }
}

}  // namespace

void launch_md_update(int n, int R, double* pos, double* vel, const double* force, const double* invm,
                      double vscale, double fscale, double noisescale, double dt, const double* noise,
                      unsigned long long seed, unsigned long long step, const MdConstraints* C,
                      double* xprime, unsigned long long* ctl, int* flags, cudaStream_t s) {
    const int total = n * R;
    if (total <= 0) return;
    const bool cons = C && (C->n_settle + C->n_shake) > 0;
    md_part12_kernel<<<(total + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        n, total, pos, vel, force, invm, vscale, fscale, noisescale, dt, 1.0 / dt, noise, seed, step,
        cons ? C->in_cluster : nullptr, xprime, ctl);
    if (cons) {
        const int units = (C->n_settle + C->n_shake) * R;
        md_constrain_kernel<<<(units + 127) / 128, 128, 0, s>>>(n, R, *C, pos, vel, xprime, invm, 1.0 / dt, step, ctl, flags);
    }
}

}  // namespace sdm
