// fp32_rates.cu -- issue-rate microbenchmarks that back the pair-kernel design:
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a), alone and interleaved with ALU / MUFU
// work.  Prints warp-instructions per clock per SM for each mix.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rates fp32_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 r) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); return a + b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ int iadd(int a, int b) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float rsq(float a) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

constexpr int CH = 8;     // independent chains per thread
constexpr int IT = 4096;  // loop trips

// MODE: 0 FFMA, 1 FFMA2, 2 FFMA+IADD 1:1, 3 FFMA2+IADD 1:1, 4 MUFU, 5 FFMA2+MUFU 4:1,
//       11 FFMA 3 distinct regs, 12 FMUL, 13 FADD, 14 FFMA2 + FFMA + ALU
//       6 FFMA2+LOP 1:1, 7 FFMA2 + 2 ALU, 8 FFMA + MUFU 8:1, 9 FFMA2 + FFMA 1:1, 10 LDS.128 + FFMA2 4:1
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[CH], x = seed * 1.0001f, y = seed * 0.9999f;
    u64 p[CH];
    int q[CH];
    for (int c = 0; c < CH; c++) { a[c] = seed + c; p[c] = pk(seed + c, seed - c); q[c] = c + (int)seed; }
    const u64 px = pk(x, x);
    long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < IT; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0) a[c] = ffma(a[c], x, x);
            if (MODE == 1) p[c] = fma2(p[c], px, px);
            if (MODE == 2) { a[c] = ffma(a[c], x, x); q[c] = iadd(q[c], it); }
            if (MODE == 3) { p[c] = fma2(p[c], px, px); q[c] = iadd(q[c], it); }
            if (MODE == 4) a[c] = rsq(a[c]);
            if (MODE == 5) { p[c] = fma2(p[c], px, px); if ((c & 3) == 0) a[c] = rsq(a[c]); }
            if (MODE == 6) { p[c] = fma2(p[c], px, px); q[c] = (int)lop((unsigned)q[c], (unsigned)it); }
            if (MODE == 7) { p[c] = fma2(p[c], px, px); q[c] = iadd(q[c], it); q[c] = (int)lop((unsigned)q[c], 0x55u); }
            if (MODE == 8) { a[c] = ffma(a[c], x, x); if (c == 0) a[c] = rsq(a[c]); }
            if (MODE == 9) { p[c] = fma2(p[c], px, px); a[c] = ffma(a[c], x, x); }
            if (MODE == 11) a[c] = ffma(a[c], x, y);
            if (MODE == 12) { float r_; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r_) : "f"(a[c]), "f"(x)); a[c] = r_; }
            if (MODE == 13) { float r_; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r_) : "f"(a[c]), "f"(x)); a[c] = r_; }
            if (MODE == 14) { p[c] = fma2(p[c], px, px); a[c] = ffma(a[c], x, y); q[c] = iadd(q[c], it); }
            if (MODE == 10) { p[c] = fma2(p[c], px, px); if ((c & 3) == 0) { float4 v = sh[(threadIdx.x + q[c]) & 255]; q[c] += __float_as_int(v.x) & 1; } }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < CH; c++) s += a[c] + lo(p[c]) + (float)q[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double instr_per_chain_step, float* out, long long* cyc, int blocks_per_sm) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * blocks_per_sm;
    k<MODE><<<grid, 256>>>(out, cyc, 1.0f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // rate from the kernel's own duration (CUDA events) and the SM clock the device reports: the
    // whole grid is resident at once (blocks_per_sm <= 8 blocks of 256 threads), so
    // warp-instructions per SM / (ms * clock) is the sustained issue rate per SM.  (Per-block clock64
    // deltas, used before, under-count the time of blocks that share a scheduler: rates above the
    // 4/clk/SM issue limit came out of that.)
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clocks = (double)ms * 1e-3 * (double)khz * 1e3;
    const double warp_instr = (double)IT * CH * instr_per_chain_step * 8.0 * blocks_per_sm;  // per SM
    printf("%-28s blocks/SM=%d  warp-instr/clk/SM=%6.3f  (%.3f ms at %.3f GHz)\n", name, blocks_per_sm, warp_instr / clocks, ms,
           khz * 1e-6);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 8 * 256 * 2);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    for (int b = 4; b <= 8; b += 4) {
        run<0>("FFMA", 1, out, cyc, b);
        run<1>("FFMA2", 1, out, cyc, b);
        run<2>("FFMA+IADD 1:1", 2, out, cyc, b);
        run<3>("FFMA2+IADD 1:1", 2, out, cyc, b);
        run<6>("FFMA2+LOP 1:1", 2, out, cyc, b);
        run<7>("FFMA2+IADD+LOP 1:1:1", 3, out, cyc, b);
        run<9>("FFMA2+FFMA 1:1", 2, out, cyc, b);
        run<11>("FFMA 3 distinct", 1, out, cyc, b);
        run<12>("FMUL", 1, out, cyc, b);
        run<13>("FADD", 1, out, cyc, b);
        run<14>("FFMA2+FFMA+IADD 1:1:1", 3, out, cyc, b);
        run<4>("MUFU.RSQ", 1, out, cyc, b);
        run<5>("FFMA2+MUFU 4:1", 1.25, out, cyc, b);
        run<8>("FFMA+MUFU 8:1", 1.125, out, cyc, b);
        run<10>("FFMA2+LDS.128 4:1", 1.25, out, cyc, b);
    }
    return 0;
}
