// kernels_restraints.cu -- the restraint forces of the reference's SDMUtils (the
// `SDMUtils` part): what python/SDMUtils.py:32-162 (addRestraintForce: a CustomCentroidBondForce) and
// :166-258 (addAlignmentForce: three CustomCompoundBondForces) hand to OpenMM as algebraic expressions
// in force group 1, evaluated here on the device for every resident replica, in double precision.
//
// They belong to the "bonded" side of the path (row a10): their energy enters PotEnergy like Eb and
// their forces are added to the hybrid force like Fb -- after the mix kernel, by the one kernel of
// this file, so an evaluation gains a single launch and nothing when no restraint is defined.
//
// The expressions are restated from the strings of SDMUtils.py; the functions they call (step, max,
// floor, angle(), dihedral(), centroid = mass-weighted mean) follow OpenMM 7.3's
// ReferenceCustomCentroidBondIxn / ReferenceCustomCompoundBondIxn (angle at the middle point;
// dihedral = angle between (p1-p2)x(p3-p2) and (p3-p2)x(p3-p4), sign of (p1-p2).((p3-p2)x(p3-p4))).
// OpenMM is not available in this build environment: like the pair arithmetic this part is
// "restated, unpinned"; the oracle (oracle/restraints.py, autograd of the same expressions) checks
// the derivatives, not OpenMM's conventions.
//
// Gradients: forward-mode automatic differentiation over the (at most 8) points of a term, i.e. 24
// partials carried through the expression -- exact derivatives of exactly the restated expression,
// no hand-derived angle / dihedral gradients to get wrong.  A term costs microseconds per replica.
#include <cmath>

#include "sdm_internal.cuh"
#include "sdm_kernels.h"

namespace sdm {
namespace {

constexpr int kND = 3 * kRestraintPoints;   // partial derivatives carried

struct Dual {
    double v;
    double d[kND];
};

__device__ inline Dual dconst(double v) {
    Dual r;
    r.v = v;
    for (int k = 0; k < kND; k++) r.d[k] = 0.0;
    return r;
}
__device__ inline Dual operator+(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v + b.v;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] + b.d[k];
    return r;
}
__device__ inline Dual operator-(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v - b.v;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] - b.d[k];
    return r;
}
__device__ inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v -= b; return r; }
__device__ inline Dual operator*(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v * b.v;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
__device__ inline Dual operator*(const Dual& a, double b) {
    Dual r;
    r.v = a.v * b;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * b;
    return r;
}
__device__ inline Dual operator/(const Dual& a, const Dual& b) {
    Dual r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
    for (int k = 0; k < kND; k++) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    return r;
}
__device__ inline Dual dsqrt(const Dual& a) {
    Dual r;
    r.v = sqrt(a.v);
    const double g = r.v > 0.0 ? 0.5 / r.v : 0.0;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * g;
    return r;
}
__device__ inline Dual dacos(const Dual& a) {
    Dual r;
    const double x = fmin(1.0, fmax(-1.0, a.v));
    r.v = acos(x);
    const double s = 1.0 - x * x;
    const double g = s > 0.0 ? -1.0 / sqrt(s) : 0.0;
    for (int k = 0; k < kND; k++) r.d[k] = a.d[k] * g;
    return r;
}

struct DVec {
    Dual x, y, z;
};
__device__ inline DVec operator-(const DVec& a, const DVec& b) { return DVec{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ inline Dual dot(const DVec& a, const DVec& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ inline DVec cross(const DVec& a, const DVec& b) {
    return DVec{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ inline DVec scale(const DVec& a, const Dual& s) { return DVec{a.x * s, a.y * s, a.z * s}; }

// angle(p1, p2, p3): at the middle point
__device__ inline Dual angle3(const DVec& p1, const DVec& p2, const DVec& p3) {
    const DVec a = p1 - p2, b = p3 - p2;
    return dacos(dot(a, b) / dsqrt(dot(a, a) * dot(b, b)));
}
// dihedral(p1, p2, p3, p4), OpenMM's convention
__device__ inline Dual dihedral4(const DVec& p1, const DVec& p2, const DVec& p3, const DVec& p4) {
    const DVec v0 = p1 - p2, v1 = p3 - p2, v2 = p3 - p4;
    const DVec c0 = cross(v0, v1), c1 = cross(v1, v2);
    Dual a = dacos(dot(c0, c1) / dsqrt(dot(c0, c0) * dot(c1, c1)));
    if (dot(v0, c1).v < 0.0) a = a * -1.0;
    return a;
}
// x - period*floor(x/period + 0.5): the floor is piecewise constant
__device__ inline Dual wrap(const Dual& x, double period) { return x - period * floor(x.v / period + 0.5); }

// (kf/2)*(step(dm)*max(0,db)^2 + step(-dm)*max(0,-da)^2) with the window [a, b] (SDMUtils.py:66-83)
__device__ inline Dual flat_bottom(const Dual& value, double kf, double a, double b, double period) {
    const Dual db = wrap(value - b, period), da = wrap(value - a, period), dm = wrap(value - 0.5 * (a + b), period);
    Dual e = dconst(0.0);
    if (dm.v >= 0.0 && db.v > 0.0) e = e + db * db;             // step(0) = 1 in OpenMM
    if (-dm.v >= 0.0 && -da.v > 0.0) e = e + da * da;
    return e * (0.5 * kf);
}

// SDMUtils.addRestraintForce, python/SDMUtils.py:61-85.  Points: g1 ligand centroid, g2 receptor centroid,
// g3..g5 receptor reference atoms, g6..g8 ligand reference atoms.
__device__ Dual centroid_restraint_energy(const DVec* g, const RestraintTerm& t, double control) {
    const double kfcm = t.p[0], tolcm = t.p[1];
    DVec d = g[0] - g[1];
    d.x = d.x - t.p[2]; d.y = d.y - t.p[3]; d.z = d.z - t.p[4];
    const Dual d12 = dsqrt(dot(d, d));
    Dual e = dconst(0.0);
    if (d12.v - tolcm >= 0.0) {
        const Dual x = d12 - tolcm;
        e = x * x * (0.5 * kfcm);
    }
    if (t.npoints == 8) {
        const double pi = 3.14159265358979323846;
        e = e + flat_bottom(angle3(g[2], g[5], g[6]), t.p[5], t.p[6], t.p[7], pi);
        e = e + flat_bottom(dihedral4(g[3], g[2], g[5], g[6]), t.p[8], t.p[9], t.p[10], 2.0 * pi);
        e = e + flat_bottom(dihedral4(g[2], g[5], g[6], g[7]), t.p[11], t.p[12], t.p[13], 2.0 * pi);
    }
    return e * control;
}

// (k/2)*(1 - cos(v, w)), v / w = d0 / d3 with their components along dn1 removed (SDMUtils.py:218-244)
__device__ inline Dual psi_term(const DVec& x1, const DVec& x2, const DVec& x3, const DVec& x4, const DVec& x5, double k) {
    const DVec d1 = x2 - x1;
    const DVec dn1 = scale(d1, dconst(1.0) / dsqrt(dot(d1, d1)));
    const DVec d0 = x3 - x1, d3 = x5 - x4;
    const DVec v = d0 - scale(dn1, dot(d0, dn1)), w = d3 - scale(dn1, dot(d3, dn1));
    const Dual cosp = dot(v, w) / dsqrt(dot(v, v) * dot(w, w));
    return (dconst(1.0) - cosp) * (0.5 * k);
}

// SDMUtils.addAlignmentForce, python/SDMUtils.py:183-256.  Points: b1, b2, b3 (ligb_ref_particles), a1, a2,
// a3 (liga_ref_particles).
__device__ Dual alignment_energy(const DVec* q, const RestraintTerm& t) {
    const double kfdispl = t.p[0], ktheta = t.p[1], kpsi = t.p[2];
    const DVec &b1 = q[0], &b2 = q[1], &b3 = q[2], &a1 = q[3], &a2 = q[4], &a3 = q[5];
    DVec d = b1 - a1;
    d.x = d.x - t.p[3]; d.y = d.y - t.p[4]; d.z = d.z - t.p[5];
    Dual e = dot(d, d) * (0.5 * kfdispl);
    const DVec d1 = b2 - b1, d2 = a2 - a1;
    const Dual cost = dot(d1, d2) / dsqrt(dot(d1, d1) * dot(d2, d2));
    e = e + (dconst(1.0) - cost) * (0.5 * ktheta);
    e = e + psi_term(b1, b2, b3, a1, a3, 0.5 * kpsi);
    e = e + psi_term(a1, a2, a3, b1, b3, 0.5 * kpsi);   // symmetrised
    return e;
}

constexpr int kRThreads = 128;

// One block per replica; the terms one after the other, the points of a term one after the other when
// their forces are handed out (an atom may sit in two groups of a term), so every addition to F has a
// fixed place in a fixed order.
__global__ void __launch_bounds__(kRThreads)
restraints_kernel(RestraintTables RT, int n, const double* __restrict__ pos_all, double* __restrict__ F_all,
                  ReplicaState* state, double* erest) {
    __shared__ double s_part[kRThreads][3];
    __shared__ double s_pt[kRestraintPoints][3];
    __shared__ double s_grad[kND];
    __shared__ double s_e;
    const int r = blockIdx.x;
    const double* pos = pos_all + (size_t)r * 3 * n;
    double* F = F_all + (size_t)r * 3 * n;
    if (threadIdx.x == 0) s_e = 0.0;
    for (int ti = 0; ti < RT.n_terms; ti++) {
        const RestraintTerm& t = RT.terms[ti];
        for (int k = 0; k < t.npoints; k++) {
            const int b = t.grp_begin[k], e = t.grp_begin[k + 1];
            double sx = 0, sy = 0, sz = 0;
            for (int a = b + threadIdx.x; a < e; a += kRThreads) {
                const int at = RT.atoms[a];
                const double w = RT.weights[a];
                sx += w * pos[3 * at]; sy += w * pos[3 * at + 1]; sz += w * pos[3 * at + 2];
            }
            s_part[threadIdx.x][0] = sx; s_part[threadIdx.x][1] = sy; s_part[threadIdx.x][2] = sz;
            __syncthreads();
            if (threadIdx.x < 3) {
                double acc = 0.0;
                for (int j = 0; j < kRThreads; j++) acc += s_part[j][threadIdx.x];
                s_pt[k][threadIdx.x] = acc;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            DVec g[kRestraintPoints];
            for (int k = 0; k < kRestraintPoints; k++) {
                g[k].x = dconst(k < t.npoints ? s_pt[k][0] : 0.0);
                g[k].y = dconst(k < t.npoints ? s_pt[k][1] : 0.0);
                g[k].z = dconst(k < t.npoints ? s_pt[k][2] : 0.0);
                g[k].x.d[3 * k] = 1.0; g[k].y.d[3 * k + 1] = 1.0; g[k].z.d[3 * k + 2] = 1.0;
            }
            const Dual e = t.kind == 0 ? centroid_restraint_energy(g, t, RT.control) : alignment_energy(g, t);
            for (int k = 0; k < kND; k++) s_grad[k] = e.d[k];
            s_e += e.v;
        }
        __syncthreads();
        for (int k = 0; k < t.npoints; k++) {
            const int b = t.grp_begin[k], e = t.grp_begin[k + 1];
            for (int a = b + threadIdx.x; a < e; a += kRThreads) {
                const int at = RT.atoms[a];
                const double w = RT.weights[a];
                F[3 * at] -= w * s_grad[3 * k]; F[3 * at + 1] -= w * s_grad[3 * k + 1]; F[3 * at + 2] -= w * s_grad[3 * k + 2];
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        erest[r] = s_e;
        state[r].sc.pot_energy += s_e;   // like Eb: PotEnergy = E1 + ebias + Eb (+ the built-in restraints)
    }
}

}  // namespace

void launch_restraints(const RestraintTables& RT, int n, int R, const double* pos_all, double* F_all,
                       ReplicaState* state, double* erest, cudaStream_t s) {
    if (RT.n_terms <= 0) return;
    restraints_kernel<<<R, kRThreads, 0, s>>>(RT, n, pos_all, F_all, state, erest);
}

}  // namespace sdm
