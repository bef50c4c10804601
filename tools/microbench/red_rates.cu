// red_rates.cu -- throughput of the memory-side operations a per-atom j list would need, chip-wide
// (all SMs busy, 16 warps per SM): 64-bit fixed-point REDs with different address patterns, vector
// FP32 REDs, shared-memory atomics, and 16-byte gathers from an L2-resident array.
// Prints lane-operations per ns and SM-cycles per lane-operation (at the clock it measures).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rates red_rates.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// MODE 0: RED.64, every lane a random slot            (spread)
// MODE 1: RED.64, runs of 8 consecutive 8-byte slots  (one j-cluster's atoms, plane layout)
// MODE 2: RED.64, runs of 8 slots at 16-byte stride   (xy packed + z interleaved per atom)
// MODE 3: RED.64, fully coalesced (32 consecutive slots)
// MODE 4: red.v4.f32, random 16-byte slots
// MODE 5: LDG.128 gather, random 16-byte slots
// MODE 6: LDG.128 gather, runs of 8 consecutive 16-byte slots
// MODE 7: ATOMS.64 spread over 16 KB of shared memory
// MODE 8: RED.32 (u32), random slot
// MODE 9: LDG.128 + LDG.64 gather from two arrays (posq + par), random slots
// MODE 10: one LDG.256 (two LDG.128 to the same 32-byte sector), random slots
template <int MODE>
__global__ void __launch_bounds__(128) k(u64* acc, const float4* src, const float2* src2, unsigned nslots, int iters, float* sink) {
    __shared__ u64 sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float s = 0.f;
    unsigned seed = gw * 0x9e3779b9U + 12345u;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        seed = seed * 1664525u + 1013904223u;
        const unsigned hl = hash32(seed + lane * 0x85ebca6bU);      // per lane
        const unsigned hg = hash32(seed + (lane >> 3) * 0xc2b2ae35U); // per group of 8 lanes
        const unsigned hw = hash32(seed);                           // per warp
        if (MODE == 0) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(acc + hl % nslots), "l"((u64)it) : "memory"); }
        if (MODE == 1) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(acc + (hg % (nslots / 8)) * 8 + (lane & 7)), "l"((u64)it) : "memory"); }
        if (MODE == 2) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(acc + ((hg % (nslots / 16)) * 8 + (lane & 7)) * 2), "l"((u64)it) : "memory"); }
        if (MODE == 3) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(acc + (hw % (nslots / 32)) * 32 + lane), "l"((u64)it) : "memory"); }
        if (MODE == 4) { float* p = reinterpret_cast<float*>(acc) + (size_t)(hl % (nslots / 2)) * 4;
                         asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory"); }
        if (MODE == 5) { const float4 v = src[hl % nslots]; s += v.x + v.w; }
        if (MODE == 6) { const float4 v = src[(hg % (nslots / 8)) * 8 + (lane & 7)]; s += v.x + v.w; }
        if (MODE == 7) { atomicAdd(&sh[hl & 2047], (u64)it); }
        if (MODE == 8) { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(reinterpret_cast<unsigned*>(acc) + hl % nslots), "r"((unsigned)it) : "memory"); }
        if (MODE == 9) { const unsigned a = hl % nslots; const float4 v = src[a]; const float2 w = src2[a]; s += v.x + v.w + w.x; }
        if (MODE == 10) { const unsigned a = (hl % (nslots / 2)) * 2; const float4 v = src[a]; const float4 w = src[a + 1]; s += v.x + v.w + w.x; }
    }
    if (MODE == 7) { __syncthreads(); s += (float)sh[threadIdx.x]; }
    if (s == 123.456f) sink[0] = s;
}

template <int MODE>
void run(const char* name, u64* acc, float4* src, float2* src2, unsigned nslots, float* sink, int sms, double ghz) {
    const int iters = 2000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 128>>>(acc, src, src2, nslots, 100, sink);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128>>>(acc, src, src2, nslots, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 128 * iters;
    printf("%-52s slots %8u  %8.3f ms  %7.2f lane-ops/ns  %6.3f SM-cycles per lane-op (at %.3f GHz)  err=%s\n", name, nslots, ms,
           ops / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * sms / ops, ghz, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    const unsigned big = 1u << 22;   // 4 M slots: 32 MB of u64 / 64 MB of float4 (L2 resident)
    u64* acc; float4* src; float2* src2; float* sink;
    cudaMalloc(&acc, (size_t)big * 16); cudaMemset(acc, 0, (size_t)big * 16);
    cudaMalloc(&src, (size_t)big * 16); cudaMemset(src, 0, (size_t)big * 16);
    cudaMalloc(&src2, (size_t)big * 8); cudaMemset(src2, 0, (size_t)big * 8);
    cudaMalloc(&sink, 4);
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
    for (unsigned ns : {1u << 18, 1u << 20, 1u << 22}) {
        run<0>("RED.64 random slot per lane", acc, src, src2, ns, sink, sms, ghz);
        run<1>("RED.64 runs of 8 consecutive slots", acc, src, src2, ns, sink, sms, ghz);
        run<2>("RED.64 runs of 8 slots, 16-byte stride", acc, src, src2, ns, sink, sms, ghz);
        run<3>("RED.64 32 consecutive slots", acc, src, src2, ns, sink, sms, ghz);
        run<8>("RED.32 random slot per lane", acc, src, src2, ns, sink, sms, ghz);
        run<4>("red.v4.f32 random 16-byte slot per lane", acc, src, src2, ns, sink, sms, ghz);
        run<5>("LDG.128 random slot per lane", acc, src, src2, ns, sink, sms, ghz);
        run<6>("LDG.128 runs of 8 consecutive slots", acc, src, src2, ns, sink, sms, ghz);
        run<9>("LDG.128 + LDG.64 (two arrays) random slot", acc, src, src2, ns, sink, sms, ghz);
        run<10>("2 x LDG.128 same 32-byte sector, random", acc, src, src2, ns, sink, sms, ghz);
    }
    run<7>("ATOMS.64 spread over 16 KB", acc, src, src2, big, sink, sms, ghz);
    return 0;
}
