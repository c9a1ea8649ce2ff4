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
