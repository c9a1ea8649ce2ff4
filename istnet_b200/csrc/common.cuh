// Shared helpers for the istnet_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/istnet_b200.h"

#define ISTNET_LAUNCH_CHECK()                      \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define ISTNET_CUDA_TRY(expr)                      \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Squared distance in the exact operation order of the reference's SASS:
// three FADDs, FMUL, FFMA, FFMA.  Explicit intrinsics so that nvcc neither adds nor removes a contraction.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

constexpr int kNumSMs = 148;

// FP32 -> bf16 operand pair: hi = bf16(x), lo = bf16(x - hi); x - (hi + lo) <= 2^-18 |x|.
// (A bf16-hi / fp16-lo pair would leave a 2^-21 residual, but tcgen05.mma kind::f16 rejects mixed A/B formats on
// sm_100a — measured: "illegal instruction" — so both halves stay bf16.)
#include <cuda_bf16.h>
__device__ __forceinline__ void split_hi_lo(float v, unsigned short &hi, unsigned short &lo) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}
