// Channels-last ("row") companions of the PointNet++ layers: the grouped neighbourhood tensor of a set-abstraction
// scale is produced directly as the bf16 operand pair of the first shared-MLP GEMM (ball-query indices -> gather of
// [xyz - centroid | features] rows), the max over the nsample axis is fused with the last BatchNorm + ReLU, and the
// three-NN interpolation of the feature-propagation layers reads / writes [points][channels] rows, so every gather
// moves contiguous channel vectors instead of the reference's strided (B,C,N) element gathers
// (group_points_gpu.cu:13-33, interpolate_gpu.cu:77-106, F.max_pool2d at pointnet2_modules.py:66-68).
#include "common.cuh"

namespace {
constexpr int kThreads = 256;

inline int grid1d(long long total) {
    long long g = (total + kThreads - 1) / kThreads;
    long long cap = (long long)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

// out row (b, j, l) = [ xyz[b, idx[b,j,l]] - new_xyz[b,j]  (3 ch, FIRST) | feats[b, idx[b,j,l], 0:C] ]   (pointnet2_utils.py:335-367)
__global__ void __launch_bounds__(kThreads) group_rows_split_kernel(int B, int N, int M, int ns, int C, const float *__restrict__ xyz,
                                                                     const float *__restrict__ new_xyz, const float *__restrict__ feats,
                                                                     const int32_t *__restrict__ idx, __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs) {
    const int K = 3 + C;
    const long long total = (long long)B * M * ns * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const long long row = i / K;  // (b*M + j)*ns + l
        const long long bj = row / ns;
        const int b = (int)(bj / M);
        const int src = idx[row];
        float v;
        if (k < 3) v = __fsub_rn(xyz[((long long)b * N + src) * 3 + k], new_xyz[bj * 3 + k]);
        else v = feats[((long long)b * N + src) * C + (k - 3)];
        store_planes1(pl + row * cs + k, pl_stride, nsplit, v);
    }
}
// d_feats[b, idx[row], c] += d_grouped[row, 3 + c]   (d_feats pre-zeroed; atomics as in group_points_grad)
__global__ void __launch_bounds__(kThreads) group_rows_bwd_kernel(int B, int N, int M, int ns, int C, const float *__restrict__ dg,
                                                                   const int32_t *__restrict__ idx, float *d_feats) {
    const int K = 3 + C;
    const long long total = (long long)B * M * ns * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const int b = (int)(row / ((long long)M * ns));
        atomicAdd(d_feats + ((long long)b * N + idx[row]) * C + c, dg[row * K + 3 + c]);
    }
}

// out[g, c] = max_l relu(bn(y[g*ns + l, c]));  argmax index l (first maximum)
__global__ void __launch_bounds__(kThreads) bn_relu_maxrows_kernel(long long G, int ns, int C, const float *__restrict__ y,
                                                                    const float *__restrict__ mean, const float *__restrict__ invstd,
                                                                    const float *__restrict__ gamma, const float *__restrict__ beta, float *out,
                                                                    int out_ld, int out_off, uint8_t *argmax) {
    const long long total = G * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long g = i / C;
        const float m = mean[c], s = invstd[c], ga = gamma[c], be = beta[c];
        float best = -INFINITY;
        int bi = 0;
        for (int l = 0; l < ns; ++l) {
            float u = (y[(g * ns + l) * C + c] - m) * s * ga + be;
            float v = fmaxf(u, 0.f);
            if (v > best) { best = v; bi = l; }
        }
        out[g * out_ld + out_off + c] = best;
        argmax[i] = (uint8_t)bi;
    }
}
// gsel[g*ns + l, c] = (l == argmax[g,c] && relu'(u) ) ? dz[g, c] : 0
__global__ void __launch_bounds__(kThreads) maxrows_bwd_kernel(long long G, int ns, int C, const float *__restrict__ y,
                                                                const float *__restrict__ mean, const float *__restrict__ invstd,
                                                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                const float *__restrict__ dz, int dz_ld, int dz_off, const uint8_t *__restrict__ argmax,
                                                                float *gsel) {
    const long long total = G * ns * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const long long g = row / ns;
        const int l = (int)(row % ns);
        float v = 0.f;
        if (argmax[g * C + c] == l) {
            float u = (y[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
            if (u > 0.f) v = dz[g * dz_ld + dz_off + c];
        }
        gsel[i] = v;
    }
}

// out[b, j, off + c] = fma(p3, w3, fma(p1, w1, p2*w2)),  p_q = feats[b, idx[b,j,q], c]   (interpolate_gpu.cu:77-106 on rows)
__global__ void __launch_bounds__(kThreads) interp_rows_kernel(int B, int m, int n, int C, const float *__restrict__ feats,
                                                                const int32_t *__restrict__ idx, const float *__restrict__ w, float *out, int out_ld,
                                                                int out_off) {
    const long long total = (long long)B * n * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bj = i / C;
        const int b = (int)(bj / n);
        const float *f = feats + (long long)b * m * C + c;
        const long long o = bj * 3;
        out[bj * out_ld + out_off + c] = __fmaf_rn(f[(long long)idx[o + 2] * C], w[o + 2],
                                                   __fmaf_rn(f[(long long)idx[o] * C], w[o], __fmul_rn(f[(long long)idx[o + 1] * C], w[o + 1])));
    }
}
// Operand planes of the first FP-layer GEMM in one pass: row (b,j) = [ three_interpolate(known feats)[0:C2] | skip feats[0:C1] ]
// (torch.cat([interpolated, unknow_feats], dim=1), pointnet2_modules.py:190-199) — replaces interp_rows + two split passes.
__global__ void __launch_bounds__(kThreads) interp_concat_split_kernel(int B, int m, int n, int C2, int C1, const float *__restrict__ feats,
                                                                        const int32_t *__restrict__ idx, const float *__restrict__ w,
                                                                        const float *__restrict__ skip, __nv_bfloat16 *pl, long long pl_stride,
                                                                        int nsplit, int cs) {
    const int lanes = (C2 + C1) >> 2;
    const long long total = (long long)B * n * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 4;
        const long long bj = i / lanes;
        float4 v;
        if (c < C2) {
            const int b = (int)(bj / n);
            const float *f = feats + (long long)b * m * C2 + c;
            const long long o = bj * 3;
            const float4 p1 = *reinterpret_cast<const float4 *>(f + (long long)idx[o] * C2);
            const float4 p2 = *reinterpret_cast<const float4 *>(f + (long long)idx[o + 1] * C2);
            const float4 p3 = *reinterpret_cast<const float4 *>(f + (long long)idx[o + 2] * C2);
            const float w1 = w[o], w2 = w[o + 1], w3 = w[o + 2];
            v.x = __fmaf_rn(p3.x, w3, __fmaf_rn(p1.x, w1, __fmul_rn(p2.x, w2)));
            v.y = __fmaf_rn(p3.y, w3, __fmaf_rn(p1.y, w1, __fmul_rn(p2.y, w2)));
            v.z = __fmaf_rn(p3.z, w3, __fmaf_rn(p1.z, w1, __fmul_rn(p2.z, w2)));
            v.w = __fmaf_rn(p3.w, w3, __fmaf_rn(p1.w, w1, __fmul_rn(p2.w, w2)));
        } else {
            v = *reinterpret_cast<const float4 *>(skip + bj * C1 + (c - C2));
        }
        store_planes4(pl + bj * cs + c, pl_stride, nsplit, v);
    }
}
// d_feats[b, idx[b,j,q], c] += dout[b, j, off + c] * w_q   (d_feats pre-zeroed)
__global__ void __launch_bounds__(kThreads) interp_rows_bwd_kernel(int B, int m, int n, int C, const float *__restrict__ dout, int d_ld, int d_off,
                                                                    const int32_t *__restrict__ idx, const float *__restrict__ w, float *d_feats) {
    const long long total = (long long)B * n * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bj = i / C;
        const int b = (int)(bj / n);
        float *f = d_feats + (long long)b * m * C + c;
        const long long o = bj * 3;
        const float d = dout[bj * d_ld + d_off + c];
        atomicAdd(f + (long long)idx[o] * C, __fmul_rn(d, w[o]));
        atomicAdd(f + (long long)idx[o + 1] * C, __fmul_rn(d, w[o + 1]));
        atomicAdd(f + (long long)idx[o + 2] * C, __fmul_rn(d, w[o + 2]));
    }
}
}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int istnet_group_rows_split(int B, int N, int M, int ns, int C, const float *xyz, const float *new_xyz, const float *feats,
                                       const int32_t *idx, void *planes, long long plane_stride, int nsplit, int cs, void *stream) {
    if (B <= 0 || M <= 0 || ns <= 0 || cs < 3 + C || (C > 0 && !feats) || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    group_rows_split_kernel<<<grid1d((long long)B * M * ns * (3 + C)), kThreads, 0, ST>>>(B, N, M, ns, C, xyz, new_xyz, feats, idx,
                                                                                         (__nv_bfloat16 *)planes, plane_stride, nsplit, cs);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_group_rows_bwd(int B, int N, int M, int ns, int C, const float *d_grouped, const int32_t *idx, float *d_feats, void *stream) {
    if (B <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    ISTNET_CUDA_TRY(cudaMemsetAsync(d_feats, 0, sizeof(float) * (size_t)B * N * C, ST));
    group_rows_bwd_kernel<<<grid1d((long long)B * M * ns * C), kThreads, 0, ST>>>(B, N, M, ns, C, d_grouped, idx, d_feats);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_bn_relu_maxrows(const float *y, long long G, int ns, int C, const float *mean, const float *invstd, const float *gamma,
                                      const float *beta, float *out, int out_ld, int out_off, uint8_t *argmax, void *stream) {
    if (G <= 0 || ns <= 0 || ns > 255 || C <= 0) return ISTNET_ERR_BAD_ARG;
    bn_relu_maxrows_kernel<<<grid1d(G * C), kThreads, 0, ST>>>(G, ns, C, y, mean, invstd, gamma, beta, out, out_ld, out_off, argmax);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_maxrows_bwd(const float *y, long long G, int ns, int C, const float *mean, const float *invstd, const float *gamma,
                                  const float *beta, const float *dz, int dz_ld, int dz_off, const uint8_t *argmax, float *gsel, void *stream) {
    if (G <= 0 || ns <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    maxrows_bwd_kernel<<<grid1d(G * ns * C), kThreads, 0, ST>>>(G, ns, C, y, mean, invstd, gamma, beta, dz, dz_ld, dz_off, argmax, gsel);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_interp_rows(int B, int m, int n, int C, const float *feats, const int32_t *idx, const float *weight, float *out, int out_ld,
                                  int out_off, void *stream) {
    if (B <= 0 || n <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    interp_rows_kernel<<<grid1d((long long)B * n * C), kThreads, 0, ST>>>(B, m, n, C, feats, idx, weight, out, out_ld, out_off);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_interp_rows_bwd(int B, int m, int n, int C, const float *dout, int d_ld, int d_off, const int32_t *idx, const float *weight,
                                      float *d_feats, void *stream) {
    if (B <= 0 || n <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    ISTNET_CUDA_TRY(cudaMemsetAsync(d_feats, 0, sizeof(float) * (size_t)B * m * C, ST));
    interp_rows_bwd_kernel<<<grid1d((long long)B * n * C), kThreads, 0, ST>>>(B, m, n, C, dout, d_ld, d_off, idx, weight, d_feats);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_interp_concat_split(int B, int m, int n, int C2, int C1, const float *feats, const int32_t *idx, const float *weight,
                                          const float *skip, void *planes, long long plane_stride, int nsplit, int cs, void *stream) {
    if (B <= 0 || n <= 0 || C2 <= 0 || (C2 & 3) || (C1 & 3) || C1 < 0 || (C1 > 0 && !skip) || cs < C2 + C1 || (cs & 3) || nsplit < 1 || nsplit > kMaxPlanes)
        return ISTNET_ERR_BAD_ARG;
    interp_concat_split_kernel<<<grid1d((long long)B * n * ((C2 + C1) / 4)), kThreads, 0, ST>>>(B, m, n, C2, C1, feats, idx, weight, skip,
                                                                                              (__nv_bfloat16 *)planes, plane_stride, nsplit, cs);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
