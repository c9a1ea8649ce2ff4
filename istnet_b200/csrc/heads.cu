// Pose-head tail of the three estimators (ist_net.py:228-264 LightEstimator, :296-332 HeavyEstimator; duplicate posenet_gt.py:71-136):
//   AdaptiveAvgPool1d(1) over the N points -> three heads [Linear 512-512, ReLU, Linear 512-256, ReLU, Linear 256-k] (k = 6, 3, 3)
//   -> Ortho6d2Mat on the rotation head (utils/rotation_utils.py:4-28).
// The reference (and round 1) issue ~25 library launches forward and ~70 backward per estimator for this (cuBLAS SIMT sgemm with
// M = batch rows, bias/ReLU element-wise kernels, ~30 micro-kernels of the Gram-Schmidt).  Here: one pooling kernel, one launch per
// Linear LAYER for all three heads (blockIdx.y = head), one Ortho6d kernel — each with a hand-written backward.  M = B <= 64
// rows, so this is latency-bound SIMT work: a warp owns an output feature, its lanes stride the K input features (coalesced weight
// reads), the B batch rows are register accumulators, reduced over the lanes with shuffles.  All sums in a fixed order.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxB = 64;   // batch rows per launch (cfg3 uses 64)
constexpr int kBT = 16;     // batch rows per register tile

struct HeadPtrs {            // up to 3 heads per launch
    const float *x[3];       // input rows [B][K]   (the same pooled vector for layer 1)
    const float *w[3];       // [O][K]
    const float *b[3];       // [O]
    float *y[3];             // [B][O]
    int O[3];
};

// ------------------------------------------------------------------ pooling over the points
// pooled[b][c] = (1/N) sum_n feat[(b*N + n)*C + c]   (nn.AdaptiveAvgPool1d(1));  grid = (ceil(C/32), B): a CTA owns 32 channels of one
// instance (8 lanes x float4 = one 128-byte line per row), its 32 row groups stride the N rows with 4 independent loads in flight per
// thread; fixed-order combine in shared memory (deterministic).  [B=32, N=1024, C=512] -> 512 CTAs (the first version used 128 CTAs
// with one dependent load stream per thread: 31 us at 2.2 TB/s, profiles/r2_ncu_families.txt).
__global__ void __launch_bounds__(kThreads) rows_mean_kernel(int N, int C, const float *__restrict__ feat, float *__restrict__ pooled) {
    constexpr int kGroups = kThreads / 8;
    __shared__ float4 red[kThreads];
    const int b = blockIdx.y;
    const int q = threadIdx.x & 7, g = threadIdx.x >> 3;
    const int c = blockIdx.x * 32 + q * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        const float *src = feat + (size_t)b * N * C + c;
        int n = g;
        for (; n + 3 * kGroups < N; n += 4 * kGroups) {
            const float4 v0 = *reinterpret_cast<const float4 *>(src + (size_t)n * C);
            const float4 v1 = *reinterpret_cast<const float4 *>(src + (size_t)(n + kGroups) * C);
            const float4 v2 = *reinterpret_cast<const float4 *>(src + (size_t)(n + 2 * kGroups) * C);
            const float4 v3 = *reinterpret_cast<const float4 *>(src + (size_t)(n + 3 * kGroups) * C);
            s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
            s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
        }
        for (; n < N; n += kGroups) {
            const float4 v = *reinterpret_cast<const float4 *>(src + (size_t)n * C);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 8 && c < C) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < kGroups; ++k) {
            const float4 v = red[k * 8 + q];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        const float inv = 1.f / (float)N;
        *reinterpret_cast<float4 *>(pooled + (size_t)b * C + c) = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
    }
}
// out[g][c] = sum over the rows r of group g (r / group == g) of sum_planes pl[plane][r][c]: per-instance column sums of a gradient that
// exists only as bf16 operand planes (the dy operand of the estimators' first layer; its sum over an instance's rows is the gradient
// of the per-instance bias W_b mean(f) + b).  grid = (ceil(C/64), groups): 8 lanes x 8 channels (one 16-byte load per plane), 32 row
// groups, fixed-order combine.
__global__ void __launch_bounds__(kThreads) rows_group_sum_planes_kernel(const __nv_bfloat16 *__restrict__ pl, long long pl_stride, int nsplit, long long P,
                                                                         int C, int cs, int group, float *__restrict__ out) {
    constexpr int kGroups = kThreads / 8;
    __shared__ float red[kThreads][8];
    const int g = blockIdx.y;
    const int q = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int c = blockIdx.x * 64 + q * 8;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    if (c < C) {
        const long long r0 = (long long)g * group;
        const long long r1 = min(P, r0 + group);
        for (long long r = r0 + rg; r < r1; r += kGroups) {
            for (int p = 0; p < nsplit; ++p) {
                const uint4 v = *reinterpret_cast<const uint4 *>(pl + (size_t)p * pl_stride + (size_t)r * cs + c);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    s[2 * i] += __uint_as_float(w[i] << 16);
                    s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = s[i];
    __syncthreads();
    if (threadIdx.x < 64 && blockIdx.x * 64 + (int)threadIdx.x < C) {
        const int qq = threadIdx.x >> 3, i = threadIdx.x & 7;
        float t = 0.f;
        for (int k = 0; k < kGroups; ++k) t += red[k * 8 + qq][i];
        out[(size_t)g * C + blockIdx.x * 64 + threadIdx.x] = t;
    }
}
// d_feat[(b*N + n)*C + c] = d_pooled[b][c] / N
__global__ void __launch_bounds__(kThreads) rows_mean_bwd_kernel(int B, int N, int C, const float *__restrict__ dpooled, float *__restrict__ dfeat) {
    const int lanes = C >> 2;
    const long long total = (long long)B * N * lanes;
    const float inv = 1.f / (float)N;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int c = (int)(i % lanes) * 4;
        const int b = (int)(i / ((long long)N * lanes));
        const float4 v = *reinterpret_cast<const float4 *>(dpooled + (size_t)b * C + c);
        reinterpret_cast<float4 *>(dfeat)[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
}

// ------------------------------------------------------------------ Linear (+ReLU) for small batches, up to 3 heads per launch
// y[b][o] = act(sum_k w[o][k] x[b][k] + bias[o]);  grid = (ceil(Omax / kWarps), heads, ceil(B / kBT)); dynamic smem = kBT*K floats
__global__ void __launch_bounds__(kThreads) heads_linear_kernel(HeadPtrs p, int B, int K, int relu) {
    extern __shared__ __align__(16) float xs[];   // [kBT][K]
    const int h = blockIdx.y, b0 = blockIdx.z * kBT;
    const int O = p.O[h];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = min(kBT, B - b0);
    for (int i = threadIdx.x; i < kBT * K; i += kThreads) {
        const int r = i / K;
        xs[i] = r < nb ? p.x[h][(size_t)(b0 + r) * K + (i - r * K)] : 0.f;
    }
    __syncthreads();
    const int o = blockIdx.x * kWarps + warp;
    if (o >= O) return;
    float acc[kBT];
#pragma unroll
    for (int r = 0; r < kBT; ++r) acc[r] = 0.f;
    const float *wr = p.w[h] + (size_t)o * K;
    for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < kBT; ++r) acc[r] = __fmaf_rn(wv, xs[r * K + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kBT; ++r) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
    }
    if (lane < nb) {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < kBT; ++r) v = (lane == r) ? acc[r] : v;
        v += p.b[h] ? __ldg(p.b[h] + o) : 0.f;
        if (relu) v = fmaxf(v, 0.f);
        p.y[h][(size_t)(b0 + lane) * O + o] = v;
    }
}

struct HeadBwdPtrs {
    const float *x[3];      // layer input [B][K]
    const float *w[3];      // [O][K]
    const float *y[3];      // layer output [B][O] (ReLU mask) or null
    const float *dy[3];     // gradient w.r.t. the layer output [B][O]
    float *dx[3];           // [B][K]: gradient w.r.t. the input (null: not needed); ACCUMULATE = add into it (shared input of the 3 heads)
    float *dw[3];           // [O][K]
    float *db[3];           // [O]
    int O[3];
};
// role 0 (blocks < nblk_w): dW[o][k] = sum_b g[b][o] x[b][k], db[o] = sum_b g[b][o]   with g = dy * [y > 0]
//        grid.x covers O*K/kThreads elements (thread = one (o, k), k fastest: coalesced x reads and dW writes)
// role 1 (blocks >= nblk_w): dx[b][k] = sum_o g[b][o] w[o][k]     CTA = one batch row x 32 input features, 8 warps split the outputs
__global__ void __launch_bounds__(kThreads) heads_linear_bwd_kernel(HeadBwdPtrs p, int B, int K, int relu, int nblk_w) {
    const int h = blockIdx.y;
    const int O = p.O[h];
    const float *dy = p.dy[h], *y = p.y[h];
    if ((int)blockIdx.x < nblk_w) {
        const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
        if (i >= (long long)O * K) return;
        const int o = (int)(i / K), k = (int)(i - (long long)o * K);
        float acc = 0.f, accb = 0.f;
        for (int b = 0; b < B; ++b) {
            float g = dy[(size_t)b * O + o];
            if (relu && !(y[(size_t)b * O + o] > 0.f)) g = 0.f;
            acc = __fmaf_rn(g, __ldg(p.x[h] + (size_t)b * K + k), acc);
            accb += g;
        }
        p.dw[h][i] = acc;
        if (k == 0 && p.db[h]) p.db[h][o] = accb;
    } else {
        // one CTA per (batch row, 32 input features): 8 warps split the O outputs, partial sums combined in a fixed order
        __shared__ float red[kWarps][32];
        if (!p.dx[h]) return;
        const int kchunks = (K + 31) / 32;
        const int blk = (int)blockIdx.x - nblk_w;
        const int b = blk / kchunks, k = (blk - b * kchunks) * 32 + (threadIdx.x & 31);
        const int sl = threadIdx.x >> 5;
        float acc = 0.f;
        if (k < K) {
            for (int o = sl; o < O; o += kWarps) {
                float g = dy[(size_t)b * O + o];
                if (relu && !(y[(size_t)b * O + o] > 0.f)) g = 0.f;
                acc = __fmaf_rn(g, __ldg(p.w[h] + (size_t)o * K + k), acc);
            }
        }
        red[sl][threadIdx.x & 31] = acc;
        __syncthreads();
        if (sl == 0 && k < K) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < kWarps; ++q) t += red[q][threadIdx.x & 31];
            p.dx[h][(size_t)b * K + k] = t;
        }
    }
}
// dx_total[i] = a[i] + b[i] + c[i]  (the three heads share the pooled input)
__global__ void __launch_bounds__(kThreads) sum3_kernel(long long n, const float *a, const float *b, const float *c, float *out) {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) out[i] = (a[i] + b[i]) + c[i];
}

// ------------------------------------------------------------------ Ortho6d2Mat (utils/rotation_utils.py:4-28), rows r6[b][0:3] = x_raw, [3:6] = y_raw
__device__ __forceinline__ void cross3(const float *a, const float *b, float *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float norm3c(const float *v) { return fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-8f); }  // normalize_vector's clamp
// R[b] = [x y z] (columns): y = n(y_raw); z = n(x_raw x y); x = y x z
__global__ void ortho6d_kernel(int B, const float *__restrict__ r6, float *__restrict__ R) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *xr = r6 + (size_t)b * 6, *yr = xr + 3;
    float y[3], z[3], x[3], w[3];
    const float ny = norm3c(yr);
    for (int i = 0; i < 3; ++i) y[i] = yr[i] / ny;
    cross3(xr, y, w);
    const float nw = norm3c(w);
    for (int i = 0; i < 3; ++i) z[i] = w[i] / nw;
    cross3(y, z, x);
    for (int i = 0; i < 3; ++i) { R[(size_t)b * 9 + i * 3 + 0] = x[i]; R[(size_t)b * 9 + i * 3 + 1] = y[i]; R[(size_t)b * 9 + i * 3 + 2] = z[i]; }
}
// gradient of the above: dR [B][3][3] -> d_r6 [B][6]
__global__ void ortho6d_bwd_kernel(int B, const float *__restrict__ r6, const float *__restrict__ dR, float *__restrict__ dr6) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *xr = r6 + (size_t)b * 6, *yr = xr + 3;
    float y[3], z[3], w[3], gx[3], gy[3], gz[3], t[3], gw[3], gxr[3], gyr[3];
    const float sy = sqrtf(yr[0] * yr[0] + yr[1] * yr[1] + yr[2] * yr[2]), ny = fmaxf(sy, 1e-8f);
    for (int i = 0; i < 3; ++i) y[i] = yr[i] / ny;
    cross3(xr, y, w);
    const float sw = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), nw = fmaxf(sw, 1e-8f);
    for (int i = 0; i < 3; ++i) z[i] = w[i] / nw;
    for (int i = 0; i < 3; ++i) { gx[i] = dR[(size_t)b * 9 + i * 3 + 0]; gy[i] = dR[(size_t)b * 9 + i * 3 + 1]; gz[i] = dR[(size_t)b * 9 + i * 3 + 2]; }
    // x = y x z:  gy += z x gx,  gz += gx x y
    cross3(z, gx, t);
    for (int i = 0; i < 3; ++i) gy[i] += t[i];
    cross3(gx, y, t);
    for (int i = 0; i < 3; ++i) gz[i] += t[i];
    // z = w / max(|w|, 1e-8): above the clamp gw = (gz - z (z.gz)) / |w|, on the clamp the divisor is a constant
    if (sw > 1e-8f) {
        const float d = z[0] * gz[0] + z[1] * gz[1] + z[2] * gz[2];
        for (int i = 0; i < 3; ++i) gw[i] = (gz[i] - z[i] * d) / nw;
    } else {
        for (int i = 0; i < 3; ++i) gw[i] = gz[i] / nw;
    }
    // w = x_raw x y:  g_xraw = y x gw,  gy += gw x x_raw
    cross3(y, gw, gxr);
    cross3(gw, xr, t);
    for (int i = 0; i < 3; ++i) gy[i] += t[i];
    if (sy > 1e-8f) {
        const float d = y[0] * gy[0] + y[1] * gy[1] + y[2] * gy[2];
        for (int i = 0; i < 3; ++i) gyr[i] = (gy[i] - y[i] * d) / ny;
    } else {
        for (int i = 0; i < 3; ++i) gyr[i] = gy[i] / ny;
    }
    for (int i = 0; i < 3; ++i) { dr6[(size_t)b * 6 + i] = gxr[i]; dr6[(size_t)b * 6 + 3 + i] = gyr[i]; }
}
}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int istnet_rows_mean(int B, int N, int C, const float *feat, float *pooled, void *stream) {
    if (B <= 0 || N <= 0 || C <= 0 || (C & 3) || !feat || !pooled) return ISTNET_ERR_BAD_ARG;
    dim3 grid(ceil_div(C, 32), B);
    rows_mean_kernel<<<grid, kThreads, 0, ST>>>(N, C, feat, pooled);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_rows_group_sum_planes(const void *planes, long long plane_stride, int nsplit, long long P, int C, int cs, int group,
                                            float *out, void *stream) {
    if (!planes || !out || P <= 0 || C <= 0 || group <= 0 || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if ((C & 7) || (cs & 7) || (plane_stride & 7) || (reinterpret_cast<uintptr_t>(planes) & 15)) return ISTNET_ERR_UNSUPPORTED;
    const long long groups = (P + group - 1) / group;
    if (groups > 65535) return ISTNET_ERR_UNSUPPORTED;
    dim3 grid(ceil_div(C, 64), (unsigned)groups);
    rows_group_sum_planes_kernel<<<grid, kThreads, 0, ST>>>((const __nv_bfloat16 *)planes, plane_stride, nsplit, P, C, cs, group, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_rows_mean_bwd(int B, int N, int C, const float *dpooled, float *dfeat, void *stream) {
    if (B <= 0 || N <= 0 || C <= 0 || (C & 3) || !dpooled || !dfeat) return ISTNET_ERR_BAD_ARG;
    long long g = ((long long)B * N * (C / 4) + kThreads - 1) / kThreads;
    if (g > kNumSMs * 8) g = kNumSMs * 8;
    rows_mean_bwd_kernel<<<(unsigned)g, kThreads, 0, ST>>>(B, N, C, dpooled, dfeat);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_heads_linear(int nheads, int B, int K, const float *const *x, const float *const *w, const float *const *bias, float *const *y,
                                   const int *O, int relu, void *stream) {
    if (nheads < 1 || nheads > 3 || B <= 0 || B > kMaxB || K <= 0 || K > 2048 || !x || !w || !y || !O) return ISTNET_ERR_BAD_ARG;
    HeadPtrs p{};
    int omax = 0;
    for (int h = 0; h < nheads; ++h) {
        p.x[h] = x[h]; p.w[h] = w[h]; p.b[h] = bias ? bias[h] : nullptr; p.y[h] = y[h]; p.O[h] = O[h];
        if (!x[h] || !w[h] || !y[h] || O[h] <= 0) return ISTNET_ERR_BAD_ARG;
        if (O[h] > omax) omax = O[h];
    }
    const size_t smem = (size_t)kBT * K * sizeof(float);
    if (smem > 48 * 1024) ISTNET_CUDA_TRY(cudaFuncSetAttribute(heads_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(omax, kWarps), nheads, ceil_div(B, kBT));
    heads_linear_kernel<<<grid, kThreads, smem, ST>>>(p, B, K, relu);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_heads_linear_bwd(int nheads, int B, int K, const float *const *x, const float *const *w, const float *const *y,
                                       const float *const *dy, float *const *dx, float *const *dw, float *const *db, const int *O, int relu,
                                       void *stream) {
    if (nheads < 1 || nheads > 3 || B <= 0 || B > kMaxB || K <= 0 || !x || !w || !dy || !dw || !O || (relu && !y)) return ISTNET_ERR_BAD_ARG;
    HeadBwdPtrs p{};
    int omax = 0;
    bool any_dx = false;
    for (int h = 0; h < nheads; ++h) {
        p.x[h] = x[h]; p.w[h] = w[h]; p.y[h] = y ? y[h] : nullptr; p.dy[h] = dy[h]; p.dx[h] = dx ? dx[h] : nullptr; p.dw[h] = dw[h];
        p.db[h] = db ? db[h] : nullptr; p.O[h] = O[h];
        if (!x[h] || !w[h] || !dy[h] || !dw[h] || O[h] <= 0 || (relu && !p.y[h])) return ISTNET_ERR_BAD_ARG;
        if (O[h] > omax) omax = O[h];
        any_dx = any_dx || p.dx[h];
    }
    const int nblk_w = (int)(((long long)omax * K + kThreads - 1) / kThreads);
    const int nblk_x = any_dx ? B * ((K + 31) / 32) : 0;
    dim3 grid(nblk_w + nblk_x, nheads, 1);
    heads_linear_bwd_kernel<<<grid, kThreads, 0, ST>>>(p, B, K, relu, nblk_w);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_sum3(long long n, const float *a, const float *b, const float *c, float *out, void *stream) {
    if (n <= 0 || !a || !b || !c || !out) return ISTNET_ERR_BAD_ARG;
    long long g = (n + kThreads - 1) / kThreads;
    if (g > kNumSMs * 4) g = kNumSMs * 4;
    sum3_kernel<<<(unsigned)g, kThreads, 0, ST>>>(n, a, b, c, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_ortho6d(int B, const float *r6, float *R, void *stream) {
    if (B <= 0 || !r6 || !R) return ISTNET_ERR_BAD_ARG;
    ortho6d_kernel<<<ceil_div(B, 64), 64, 0, ST>>>(B, r6, R);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_ortho6d_bwd(int B, const float *r6, const float *dR, float *dr6, void *stream) {
    if (B <= 0 || !r6 || !dR || !dr6) return ISTNET_ERR_BAD_ARG;
    ortho6d_bwd_kernel<<<ceil_div(B, 64), 64, 0, ST>>>(B, r6, dR, dr6);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
