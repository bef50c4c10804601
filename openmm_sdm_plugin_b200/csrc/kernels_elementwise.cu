// kernels_elementwise.cu -- the literal kernel-interface operations of
// SDMPlugin::IntegrateLangevinStepSDMKernel on device float4 buffers (sm_100a).
//
// Behaviour restated from platforms/opencl/src/kernels/langevin.cl:72-87 (sdmForce),
// :147-155 (RestoreState1), :161-178 (SaveState1), :183-192 (SaveState2), :198-206
// (MakeState2) and :7-69 (integrateLangevinPart1 / Part2, single-precision form).  These are HBM-bound streaming kernels: 128-bit accesses, four independent
// loads in flight per thread, grid sized in multiples of the SM count, no shared memory
// (there is no reuse), streaming cache hints for write-once data.
#include <curand_kernel.h>

#include <algorithm>

#include "sdm_kernels.h"

namespace sdm {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// posq[i] += displ[i]   (48 B/atom)
__global__ void __launch_bounds__(kThreads)
make_state2_kernel(int n, float4* __restrict__ posq, const float4* __restrict__ displ) {
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        float4 p[kUnroll], d[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; k++) { p[k] = posq[i + k * stride]; d[k] = ld_stream(displ + i + k * stride); }
#pragma unroll
        for (int k = 0; k < kUnroll; k++) posq[i + k * stride] = add4(p[k], d[k]);
    }
    for (; i < n; i += stride) posq[i] = add4(posq[i], displ[i]);
}

// dst = src   (32 B/atom); SaveState2 and RestoreState1
__global__ void __launch_bounds__(kThreads)
copy4_kernel(int n, const float4* __restrict__ src, float4* __restrict__ dst) {
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        float4 v[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; k++) v[k] = src[i + k * stride];
#pragma unroll
        for (int k = 0; k < kUnroll; k++) dst[i + k * stride] = v[k];
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

// save_f = force ; save_x = posq   (64 B/atom) in one pass
__global__ void __launch_bounds__(kThreads)
save_state1_kernel(int n, const float4* __restrict__ posq, const float4* __restrict__ force,
                   float4* __restrict__ save_f, float4* __restrict__ save_x) {
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        float4 f0 = force[i], f1 = force[i + stride], x0 = posq[i], x1 = posq[i + stride];
        save_f[i] = f0; save_f[i + stride] = f1;
        save_x[i] = x0; save_x[i + stride] = x1;
    }
    for (; i < n; i += stride) { save_f[i] = force[i]; save_x[i] = posq[i]; }
}

// force = (1-sp)*f1 + sp*f2 + force on all four lanes (64 B/atom)
__global__ void __launch_bounds__(kThreads)
hybrid_force_kernel(int n, const float4* __restrict__ f1, const float4* __restrict__ f2,
                    float4* __restrict__ force, float sp) {
    const float lmb = sp, lmb1 = 1.0f - sp;
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        float4 a0 = ld_stream(f1 + i), b0 = ld_stream(f2 + i), c0 = force[i];
        float4 a1 = ld_stream(f1 + i + stride), b1 = ld_stream(f2 + i + stride), c1 = force[i + stride];
        force[i] = make_float4(lmb1 * a0.x + lmb * b0.x + c0.x, lmb1 * a0.y + lmb * b0.y + c0.y,
                               lmb1 * a0.z + lmb * b0.z + c0.z, lmb1 * a0.w + lmb * b0.w + c0.w);
        force[i + stride] = make_float4(lmb1 * a1.x + lmb * b1.x + c1.x, lmb1 * a1.y + lmb * b1.y + c1.y,
                                        lmb1 * a1.z + lmb * b1.z + c1.z, lmb1 * a1.w + lmb * b1.w + c1.w);
    }
    for (; i < n; i += stride) {
        float4 a = f1[i], b = f2[i], c = force[i];
        force[i] = make_float4(lmb1 * a.x + lmb * b.x + c.x, lmb1 * a.y + lmb * b.y + c.y,
                               lmb1 * a.z + lmb * b.z + c.z, lmb1 * a.w + lmb * b.w + c.w);
    }
}

// integrateLangevinPart1 (langevin.cl:7-31): for atoms with velm.w (inverse mass) != 0
//   v = vscale*v + fscale*w*F + noisescale*sqrt(w)*xi ;  posDelta = stepSize*v   (all four lanes)
// 80 B/atom.  Products and sums are evaluated left to right without FMA contraction, so the
// result is the float32 expression as written (the tests compare bit for bit).
__global__ void __launch_bounds__(kThreads)
langevin_part1_kernel(int n, float4* __restrict__ velm, const float4* __restrict__ force,
                      float4* __restrict__ pos_delta, float vscale, float fscale, float noisescale,
                      float step_size, const float4* __restrict__ random, unsigned random_index) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 v = velm[i];
        if (v.w != 0.0f) {
            const float4 f = ld_stream(force + i), r = ld_stream(random + random_index + i);
            const float s = sqrtf(v.w), fw = __fmul_rn(fscale, v.w), ns = __fmul_rn(noisescale, s);
            v.x = __fadd_rn(__fadd_rn(__fmul_rn(vscale, v.x), __fmul_rn(fw, f.x)), __fmul_rn(ns, r.x));
            v.y = __fadd_rn(__fadd_rn(__fmul_rn(vscale, v.y), __fmul_rn(fw, f.y)), __fmul_rn(ns, r.y));
            v.z = __fadd_rn(__fadd_rn(__fmul_rn(vscale, v.z), __fmul_rn(fw, f.z)), __fmul_rn(ns, r.z));
            velm[i] = v;
            st_stream(pos_delta + i, make_float4(__fmul_rn(step_size, v.x), __fmul_rn(step_size, v.y),
                                                 __fmul_rn(step_size, v.z), __fmul_rn(step_size, v.w)));
        }
    }
}

// integrateLangevinPart2 (langevin.cl:37-69, single precision branch): for atoms with velm.w != 0
//   posq.xyz += posDelta.xyz ;  vel.xyz = invStep*delta.xyz + correction*delta.xyz
// with invStep = 1/dt and correction = (1 - invStep*dt)/dt.  80 B/atom.
__global__ void __launch_bounds__(kThreads)
langevin_part2_kernel(int n, float4* __restrict__ posq, const float4* __restrict__ pos_delta,
                      float4* __restrict__ velm, float step_size) {
    const float inv = __fdiv_rn(1.0f, step_size);
    const float corr = __fdiv_rn(__fsub_rn(1.0f, __fmul_rn(inv, step_size)), step_size);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 v = velm[i];
        if (v.w != 0.0f) {
            float4 p = posq[i];
            const float4 d = ld_stream(pos_delta + i);
            p.x = __fadd_rn(p.x, d.x); p.y = __fadd_rn(p.y, d.y); p.z = __fadd_rn(p.z, d.z);
            v.x = __fadd_rn(__fmul_rn(inv, d.x), __fmul_rn(corr, d.x));
            v.y = __fadd_rn(__fmul_rn(inv, d.y), __fmul_rn(corr, d.y));
            v.z = __fadd_rn(__fmul_rn(inv, d.z), __fmul_rn(corr, d.z));
            posq[i] = p;
            velm[i] = v;
        }
    }
}

int grid_for(int n, int per_thread) {
    long long want = ((long long)n + (long long)kThreads * per_thread - 1) / ((long long)kThreads * per_thread);
    const int cap = 148 * 8;  // 8 resident 256-thread CTAs per SM on B200
    if (want < 1) want = 1;
    return (int)(want > cap ? cap : want);
}

// 0.5 * sum m v^2 per replica (ReferenceSDMKernels.cpp:105-137 without constraints); one block per
// replica, fixed-order reduction.
__global__ void __launch_bounds__(256)
kinetic_energy_kernel(int n, const double* __restrict__ vel, const double* __restrict__ mass, double* ke) {
    __shared__ double sh[256];
    const int r = blockIdx.x;
    double acc = 0.0;
    for (int a = threadIdx.x; a < n; a += 256) {
        const double* v = vel + 3 * ((size_t)r * n + a);
        acc += mass[a] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) ke[r] = 0.5 * sh[0];
}

}  // namespace

void launch_make_state2(int n, float4* posq, const float4* displ, cudaStream_t s) {
    if (n <= 0) return;
    make_state2_kernel<<<grid_for(n, kUnroll), kThreads, 0, s>>>(n, posq, displ);
}
void launch_save_state1(int n, const float4* posq, const float4* force, float4* save_f,
                        float4* save_x, cudaStream_t s) {
    if (n <= 0) return;
    save_state1_kernel<<<grid_for(n, 2), kThreads, 0, s>>>(n, posq, force, save_f, save_x);
}
void launch_copy4(int n, const float4* src, float4* dst, cudaStream_t s) {
    if (n <= 0) return;
    copy4_kernel<<<grid_for(n, kUnroll), kThreads, 0, s>>>(n, src, dst);
}
void launch_hybrid_force(int n, const float4* f1, const float4* f2, float4* force, float sp,
                         cudaStream_t s) {
    if (n <= 0) return;
    hybrid_force_kernel<<<grid_for(n, 2), kThreads, 0, s>>>(n, f1, f2, force, sp);
}

void launch_langevin_part1(int n, float4* velm, const float4* force, float4* pos_delta, float vscale,
                           float fscale, float noisescale, float step_size, const float4* random,
                           unsigned random_index, cudaStream_t s) {
    if (n <= 0) return;
    langevin_part1_kernel<<<grid_for(n, 1), kThreads, 0, s>>>(n, velm, force, pos_delta, vscale, fscale,
                                                             noisescale, step_size, random, random_index);
}
void launch_langevin_part2(int n, float4* posq, const float4* pos_delta, float4* velm, float step_size,
                           cudaStream_t s) {
    if (n <= 0) return;
    langevin_part2_kernel<<<grid_for(n, 1), kThreads, 0, s>>>(n, posq, pos_delta, velm, step_size);
}

// FP32 <-> FP64 conversion of flat coordinate / force arrays for the single-precision transfer calls
// (sdm_set_positions_all_f32 / sdm_enqueue_results_f32): 12 B/element, HBM bound, a few microseconds.
namespace {
__global__ void __launch_bounds__(256) widen_kernel(size_t count, const float* __restrict__ src, double* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (double)src[i];
}
__global__ void __launch_bounds__(256) narrow_kernel(size_t count, const double* __restrict__ src, float* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}
}  // namespace
void launch_widen(size_t count, const float* src, double* dst, cudaStream_t s) {
    if (count == 0) return;
    widen_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 148 * 16), 256, 0, s>>>(count, src, dst);
}
void launch_narrow(size_t count, const double* src, float* dst, cudaStream_t s) {
    if (count == 0) return;
    narrow_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 148 * 16), 256, 0, s>>>(count, src, dst);
}

// Sustained FP32 FMA rate of the device this library runs on: the denominator of the pair kernel's
// roofline, measured instead of derived from the clock.  Eight independent packed FMA chains per
// thread (fma.rn.f32x2, the instruction the pair kernel is made of), 8 x 256 threads per SM.
namespace {
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float* sink) {
    unsigned long long p[8];
    const unsigned long long m = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.999f);
    const unsigned long long a = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed);
#pragma unroll
    for (int c = 0; c < 8; c++) p[c] = a + c;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(m), "l"(a));
    }
    unsigned long long x = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) x ^= p[c];
    if (x == 0x1234567ull) sink[0] = 1.f;
}
}  // namespace
double measure_fp32_fma_tflops(int num_sms, cudaStream_t s) {
    float* sink = nullptr;
    if (cudaMalloc(&sink, sizeof(float)) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 40000, blocks = num_sms * 8;
    fma_peak_kernel<<<blocks, 256, 0, s>>>(2000, 1.0f, sink);   // warm-up (clocks)
    cudaEventRecord(e0, s);
    fma_peak_kernel<<<blocks, 256, 0, s>>>(iters, 1.0f, sink);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (!(ms > 0.f)) return 0.0;
    const double flop = (double)blocks * 256 * 8 * 2 * 2 * (double)iters;   // 8 chains x 2 halves x (mul + add)
    return flop / (ms * 1e-3) / 1e12;
}

void launch_kinetic_energy(int n, int R, const double* vel, const double* mass, double* ke, cudaStream_t s) {
    if (R <= 0) return;
    kinetic_energy_kernel<<<R, 256, 0, s>>>(n, vel, mass, ke);
}

}  // namespace sdm
