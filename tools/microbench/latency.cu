// latency.cu -- dependent-issue latencies (cycles) of the instructions on the pair kernel's
// critical path, one warp per SM: FFMA, FFMA2, FMUL2, MUFU.RSQ, LDS.128, SHFL, REDUX.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 r) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); return a + b; }
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed, int n) {
    __shared__ float4 sh[64];
    sh[threadIdx.x & 63] = make_float4(seed, seed, 0.f, 0.f);
    __syncthreads();
    float a = seed; u64 p = pk(seed, seed), q = pk(seed * 0.5f, seed); int idx = threadIdx.x & 1; unsigned u = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(seed));
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));
        if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
        if (MODE == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a));
        if (MODE == 4) { float4 v = sh[idx]; idx = __float_as_int(v.z) + (idx & 63); asm volatile("" : "+r"(idx)); }
        if (MODE == 5) { u = __shfl_xor_sync(0xffffffffu, u, 1); }
        if (MODE == 6) { u = __reduce_or_sync(0xffffffffu, u) + 1; asm volatile("" : "+r"(u)); }
        if (MODE == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + lo(p) + idx + u;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, float* out, long long* cyc) {
    const int n = 4096;
    k<MODE><<<1, 32>>>(out, cyc, 1.0f, n);
    k<MODE><<<1, 32>>>(out, cyc, 1.0f, n);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-12s %.2f cycles per dependent instruction\n", name, (double)h / n);
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    run<0>("FFMA", out, cyc); run<1>("FFMA2", out, cyc); run<2>("FMUL2", out, cyc); run<7>("FADD2", out, cyc);
    run<3>("MUFU.RSQ", out, cyc); run<4>("LDS.128", out, cyc); run<5>("SHFL", out, cyc); run<6>("REDUX+IADD", out, cyc);
    return 0;
}
