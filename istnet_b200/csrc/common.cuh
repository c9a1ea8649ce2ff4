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

// Squared distance in the exact operation order of the reference's sm_100a SASS for `dx*dx + dy*dy + dz*dz`:
// FMUL dy*dy, FFMA dx*dx + ., FFMA dz*dz + .  (cuobjdump of the reference build under oracle/_ref; the plain multiply
// is on the middle term).  Explicit intrinsics so that nvcc neither adds nor removes a contraction.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

constexpr int kNumSMs = 148;

// FP32 -> bf16 operand planes: p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) ...; the residual after n
// planes is <= 2^(-9n) |x|.  Operand tensors are stored plane-major: planes[i] is a complete bf16 tensor,
// `plane_stride` elements after planes[i-1].  nsplit = 2 -> 3 tensor-core products (error ~3e-6 per product),
// nsplit = 3 -> 6 products (error ~1e-8, below FP32 accumulation noise).
// (A bf16-hi / fp16-lo pair would reach 2^-21 with 4 products, but tcgen05.mma kind::f16 rejects mixed A/B formats on
// sm_100a — measured: "illegal instruction" — and fp16 alone underflows on gradient-sized tensors.)
#include <cuda_bf16.h>
constexpr int kMaxPlanes = 3;
__device__ __forceinline__ void split_planes(float v, unsigned short (&out)[kMaxPlanes], int n) {
    float r = v;
#pragma unroll
    for (int i = 0; i < kMaxPlanes; ++i) {
        if (i < n) {
            __nv_bfloat16 h = __float2bfloat16_rn(r);
            out[i] = __bfloat16_as_ushort(h);
            r -= __bfloat162float(h);
        }
    }
}
// scalar / 4-wide stores of one value group into every plane
__device__ __forceinline__ void store_planes1(__nv_bfloat16 *base, long long plane_stride, int n, float v) {
    unsigned short o[kMaxPlanes];
    split_planes(v, o, n);
#pragma unroll
    for (int i = 0; i < kMaxPlanes; ++i)
        if (i < n) base[i * plane_stride] = __ushort_as_bfloat16(o[i]);
}
// two values at once: one packed round-to-nearest conversion per plane (F2FP.BF16.F32.PACK_AB), the packed word is the
// store payload and, shifted / masked back to FP32, the subtrahend of the residual — same bits as split_planes
__device__ __forceinline__ void split_planes_pair(float a, float b, uint32_t (&out)[kMaxPlanes], int n) {
    float ra = a, rb = b;
#pragma unroll
    for (int i = 0; i < kMaxPlanes; ++i) {
        if (i < n) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(ra, rb);  // .x (low half) = ra, .y (high half) = rb
            const uint32_t pk = *reinterpret_cast<const uint32_t *>(&h);
            out[i] = pk;
            ra -= __uint_as_float(pk << 16);
            rb -= __uint_as_float(pk & 0xffff0000u);
        }
    }
}
__device__ __forceinline__ void store_planes4(__nv_bfloat16 *base, long long plane_stride, int n, float4 v) {
    uint32_t lo[kMaxPlanes], hi[kMaxPlanes];
    split_planes_pair(v.x, v.y, lo, n); split_planes_pair(v.z, v.w, hi, n);
#pragma unroll
    for (int i = 0; i < kMaxPlanes; ++i)
        if (i < n) *reinterpret_cast<uint2 *>(base + i * plane_stride) = make_uint2(lo[i], hi[i]);
}
// 8 consecutive channels -> one 16-byte store per plane (base 16-byte aligned, plane_stride % 8 == 0)
__device__ __forceinline__ void store_planes8(__nv_bfloat16 *base, long long plane_stride, int n, const float (&v)[8]) {
    uint32_t a[kMaxPlanes], b[kMaxPlanes], c[kMaxPlanes], d[kMaxPlanes];
    split_planes_pair(v[0], v[1], a, n); split_planes_pair(v[2], v[3], b, n);
    split_planes_pair(v[4], v[5], c, n); split_planes_pair(v[6], v[7], d, n);
#pragma unroll
    for (int i = 0; i < kMaxPlanes; ++i)
        if (i < n) *reinterpret_cast<uint4 *>(base + i * plane_stride) = make_uint4(a[i], b[i], c[i], d[i]);
}
