// f32x2.cuh -- packed pairs of FP32 values for the sm_100a FADD2 / FMUL2 / FFMA2 instructions
// (PTX add/sub/mul/fma.rn.f32x2).  One packed instruction does the work of two scalar ones for a
// single issue slot; measured on B200 (tools/microbench/fp32_rates.cu): 2 FFMA2 per clock per SM
// = the full 128 FP32 lanes, while leaving half of the 4 issue slots per clock to ALU, MUFU,
// LDS and branch instructions.  ptxas folds pk(a, a) and immediates into broadcast operands
// (R.F32 / imm) and sub2 into a negate modifier, so broadcasts cost no instruction.
#pragma once

namespace sdm {

typedef unsigned long long f2;  // (lo, hi)

__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

}  // namespace sdm
