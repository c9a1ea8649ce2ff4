// Point-cloud operators of the PointNet++ set-abstraction / feature-propagation stack for sm_100a.
//
// Semantics follow the reference kernels under /root/reference/model/pointnet2/_ext_src/src
// (sampling_gpu.cu, ball_query_gpu.cu, group_points_gpu.cu, interpolate_gpu.cu) bit for bit for every
// index-producing operator; the implementation is new:
//   * FPS keeps the running min-distance array and the coordinates in REGISTERS, reduces with two
//     redux.sync per warp plus one double-buffered shared-memory exchange (ONE barrier per round instead of
//     ten and no global-memory read-modify-write), and chains all SA levels of an extractor in one launch.
//   * ball query runs one WARP per centroid (ballot + prefix popcount keeps the reference's "first nsample
//     hits in index order" rule) over a shared-memory copy of the cloud, instead of one thread per centroid.
//   * gather / group / interpolate are fully coalesced over the innermost output dimension.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// Furthest point sampling
// ---------------------------------------------------------------------------------------------------
namespace {

constexpr int kFpsThreads = 256;
constexpr int kFpsWarps = kFpsThreads / 32;

// Reference launch geometry (cuda_utils.h:18-24): S = clamp(2^floor(log2 n), 1, 512) threads; thread t owns
// k = t, t+S, ...  Among equal maxima the reference's shared-memory tree keeps the thread with the smallest
// bit-reversed id and, inside a thread, the smallest k (strict '>' scan).  We encode that preference as a
// 32-bit priority so the arg-max may be evaluated in ANY order:  prio = ~((bitrev_S(k mod S) << 8) | (k / S)).
__device__ __forceinline__ uint32_t fps_prio(int k, int log2S) {
    uint32_t t = (uint32_t)k & ((1u << log2S) - 1u);
    uint32_t br = log2S ? (__brev(t) >> (32 - log2S)) : 0u;
    return ~((br << 8) | ((uint32_t)k >> log2S));
}
__device__ __forceinline__ int fps_prio_to_index(uint32_t prio, int log2S) {
    uint32_t key = ~prio;
    uint32_t br = key >> 8, slot = key & 0xffu;
    uint32_t t = log2S ? (__brev(br) >> (32 - log2S)) : 0u;
    return (int)(t + (slot << log2S));
}

__device__ __forceinline__ int floor_log2_clamped(int n) {
    int l = 31 - __clz(n);
    return l > 9 ? 9 : l;
}

// One FPS level executed by the whole CTA.  `cloud` is the level's input in SHARED memory ([n][3] floats).
// Selected coordinates are appended to `next_cloud` (shared, may be nullptr) and to xyz_out (global, may be
// nullptr); indices go to idx_out (global).
template <int PPT>
__device__ void fps_level(const float *cloud, int n, int m, int32_t *__restrict__ idx_out, float *__restrict__ xyz_out,
                          float *next_cloud, unsigned long long (*slots)[kFpsWarps]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int log2S = floor_log2_clamped(n);
    float px[PPT], py[PPT], pz[PPT], mind[PPT];
    uint32_t prio[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = tid + i * kFpsThreads;
        bool ok = k < n;
        px[i] = ok ? cloud[k * 3 + 0] : 0.f;
        py[i] = ok ? cloud[k * 3 + 1] : 0.f;
        pz[i] = ok ? cloud[k * 3 + 2] : 0.f;
        mind[i] = 1e10f;  // sampling.cpp:78-80
        prio[i] = ok ? fps_prio(k, log2S) : 0u;
    }
    int old = 0;
    float x1 = cloud[0], y1 = cloud[1], z1 = cloud[2];
    if (tid == 0) {
        idx_out[0] = 0;
        if (xyz_out) { xyz_out[0] = x1; xyz_out[1] = y1; xyz_out[2] = z1; }
        if (next_cloud) { next_cloud[0] = x1; next_cloud[1] = y1; next_cloud[2] = z1; }
    }
    for (int j = 1; j < m; ++j) {
        uint32_t bv = 0u, bp = 0u;  // best value bits (non-negative float => monotone as uint), best priority
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float d = sqdist_ref(__fsub_rn(px[i], x1), __fsub_rn(py[i], y1), __fsub_rn(pz[i], z1));
            float d2 = fminf(d, mind[i]);
            mind[i] = d2;
            uint32_t v = __float_as_uint(d2), p = prio[i];
            bool take = (p != 0u) && (v > bv || (v == bv && p > bp));
            bv = take ? v : bv;
            bp = take ? p : bp;
        }
        uint32_t wv = __reduce_max_sync(0xffffffffu, bv);
        uint32_t wp = __reduce_max_sync(0xffffffffu, bv == wv ? bp : 0u);
        if (lane == 0) slots[j & 1][warp] = ((unsigned long long)wv << 32) | wp;
        __syncthreads();
        unsigned long long best = slots[j & 1][0];
#pragma unroll
        for (int w = 1; w < kFpsWarps; ++w) {
            unsigned long long c = slots[j & 1][w];
            best = c > best ? c : best;
        }
        old = fps_prio_to_index((uint32_t)best, log2S);
        x1 = cloud[old * 3 + 0];
        y1 = cloud[old * 3 + 1];
        z1 = cloud[old * 3 + 2];
        if (tid == 0) {
            idx_out[j] = old;
            if (xyz_out) { xyz_out[j * 3 + 0] = x1; xyz_out[j * 3 + 1] = y1; xyz_out[j * 3 + 2] = z1; }
            if (next_cloud) { next_cloud[j * 3 + 0] = x1; next_cloud[j * 3 + 1] = y1; next_cloud[j * 3 + 2] = z1; }
        }
    }
}

__device__ void fps_level_dispatch(const float *cloud, int n, int m, int32_t *idx_out, float *xyz_out, float *next_cloud,
                                   unsigned long long (*slots)[kFpsWarps]) {
    int ppt = (n + kFpsThreads - 1) / kFpsThreads;
    if (ppt <= 1) fps_level<1>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
    else if (ppt <= 2) fps_level<2>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
    else if (ppt <= 4) fps_level<4>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
    else if (ppt <= 8) fps_level<8>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
    else if (ppt <= 16) fps_level<16>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
    else fps_level<32>(cloud, n, m, idx_out, xyz_out, next_cloud, slots);
}

struct FpsChainArgs {
    int nlevels;
    int npoint[4];
    int32_t *idx_out[4];
    float *xyz_out[4];
};

// grid = b (one CTA per instance), block = kFpsThreads, dynamic smem = (n + max_l npoint[l]) * 12 bytes
// (two ping-pong clouds sized for the largest level).
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_chain_kernel(int n, const float *__restrict__ xyz, FpsChainArgs a, int cloud_b_capacity) {
    extern __shared__ float fps_smem[];
    __shared__ unsigned long long slots[2][kFpsWarps];
    const int bi = blockIdx.x;
    float *cloud_a = fps_smem;                               // capacity: max(n, level sizes)
    float *cloud_b = fps_smem + (size_t)cloud_b_capacity * 3;  // second buffer
    const float *src = xyz + (size_t)bi * n * 3;
    for (int i = threadIdx.x; i < n * 3; i += kFpsThreads) cloud_a[i] = src[i];
    __syncthreads();
    float *cur = cloud_a, *nxt = cloud_b;
    int cur_n = n;
    for (int l = 0; l < a.nlevels; ++l) {
        int m = a.npoint[l];
        if (m > 0) {
            fps_level_dispatch(cur, cur_n, m, a.idx_out[l] + (size_t)bi * m, a.xyz_out[l] ? a.xyz_out[l] + (size_t)bi * m * 3 : nullptr,
                               (l + 1 < a.nlevels) ? nxt : nullptr, slots);
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
        cur_n = m;
    }
}

}  // namespace

static int launch_fps_chain(int b, int n, const FpsChainArgs &a, const float *xyz, cudaStream_t st) {
    if (b <= 0) return ISTNET_OK;
    if (n <= 0) return ISTNET_ERR_BAD_ARG;
    if (n > kFpsThreads * 32) return ISTNET_ERR_UNSUPPORTED;  // 8192 points per cloud
    int cap = n;
    for (int l = 0; l < a.nlevels; ++l) {
        if (a.npoint[l] < 0) return ISTNET_ERR_BAD_ARG;
        if (a.npoint[l] > kFpsThreads * 32) return ISTNET_ERR_UNSUPPORTED;
        if (a.npoint[l] > cap) cap = a.npoint[l];
    }
    size_t smem = (size_t)cap * 3 * sizeof(float) * 2;
    if (smem > 48 * 1024) {
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(fps_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    fps_chain_kernel<<<b, kFpsThreads, smem, st>>>(n, xyz, a, cap);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_furthest_point_sampling(int b, int n, int m, const float *xyz, int32_t *idx, void *stream) {
    if (m <= 0) return ISTNET_OK;  // sampling_gpu.cu:78
    FpsChainArgs a{};
    a.nlevels = 1;
    a.npoint[0] = m;
    a.idx_out[0] = idx;
    a.xyz_out[0] = nullptr;
    return launch_fps_chain(b, n, a, xyz, (cudaStream_t)stream);
}

extern "C" int istnet_fps_chain(int b, int n, int nlevels, const int *npoint, const float *xyz, int32_t *const *idx_out,
                                float *const *xyz_out, void *stream) {
    if (nlevels < 1 || nlevels > 4) return ISTNET_ERR_BAD_ARG;
    FpsChainArgs a{};
    a.nlevels = nlevels;
    for (int l = 0; l < nlevels; ++l) {
        if (npoint[l] <= 0) return ISTNET_ERR_BAD_ARG;
        a.npoint[l] = npoint[l];
        a.idx_out[l] = idx_out[l];
        a.xyz_out[l] = xyz_out ? xyz_out[l] : nullptr;
    }
    return launch_fps_chain(b, n, a, xyz, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// Ball query: one warp per centroid
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int kBqThreads = 256;
constexpr int kBqCentroidsPerBlock = 32;

// grid = (ceil(m / 32), b); dynamic smem = n*3 floats (the instance's cloud)
__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                  int32_t *__restrict__ idx) {
    extern __shared__ float bq_cloud[];
    const int bi = blockIdx.y;
    const float *src = xyz + (size_t)bi * n * 3;
    for (int i = threadIdx.x; i < n * 3; i += kBqThreads) bq_cloud[i] = src[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r2 = __fmul_rn(radius, radius);  // ball_query_gpu.cu:27
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int c = warp; c < kBqCentroidsPerBlock; c += kBqThreads / 32) {
        int j = blockIdx.x * kBqCentroidsPerBlock + c;
        if (j >= m) break;
        const float *q = new_xyz + ((size_t)bi * m + j) * 3;
        float cx = q[0], cy = q[1], cz = q[2];
        int32_t *out = idx + ((size_t)bi * m + j) * nsample;
        int cnt = 0, first = 0;
        for (int base = 0; base < n && cnt < nsample; base += 32) {
            int k = base + lane;
            bool hit = false;
            if (k < n) {
                float d2 = sqdist_ref(__fsub_rn(cx, bq_cloud[k * 3 + 0]), __fsub_rn(cy, bq_cloud[k * 3 + 1]),
                                      __fsub_rn(cz, bq_cloud[k * 3 + 2]));
                hit = d2 < r2;
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
                if (cnt == 0) first = base + __ffs(mask) - 1;
                int pos = cnt + __popc(mask & lt_mask);
                if (hit && pos < nsample) out[pos] = k;
                cnt += __popc(mask);
            }
        }
        if (cnt > nsample) cnt = nsample;
        // tail: the reference pre-fills the row with the first hit (ball_query_gpu.cu:39-43); no hit => zeros
        for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;
    }
}
}  // namespace

extern "C" int istnet_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                                 int32_t *idx, void *stream) {
    if (b <= 0 || m <= 0 || nsample <= 0) return ISTNET_OK;
    if (n < 0) return ISTNET_ERR_BAD_ARG;
    size_t smem = (size_t)n * 3 * sizeof(float);
    if (smem > 200 * 1024) return ISTNET_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(m, kBqCentroidsPerBlock), b);
    ball_query_kernel<<<grid, kBqThreads, smem, (cudaStream_t)stream>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// ---------------------------------------------------------------------------------------------------
// gather / group (+ grads)
// ---------------------------------------------------------------------------------------------------
namespace {
// out[(bi*c + l)*m + j] = points[(bi*c + l)*n + idx[bi*m + j]]
__global__ void gather_points_kernel(long long total, int c, int n, int m, const float *__restrict__ points,
                                     const int32_t *__restrict__ idx, float *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % m);
        long long row = i / m;  // bi*c + l
        int bi = (int)(row / c);
        out[i] = points[row * n + idx[(long long)bi * m + j]];
    }
}
__global__ void gather_points_grad_kernel(long long total, int c, int n, int m, const float *__restrict__ grad_out,
                                          const int32_t *__restrict__ idx, float *__restrict__ grad_points) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % m);
        long long row = i / m;
        int bi = (int)(row / c);
        atomicAdd(grad_points + row * n + idx[(long long)bi * m + j], grad_out[i]);
    }
}
// out[((bi*c + l)*np + j)*ns + k] = points[(bi*c + l)*n + idx[(bi*np + j)*ns + k]]
__global__ void group_points_kernel(long long total, int c, int n, int np_ns, const float *__restrict__ points,
                                    const int32_t *__restrict__ idx, float *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int jk = (int)(i % np_ns);
        long long row = i / np_ns;
        int bi = (int)(row / c);
        out[i] = points[row * n + idx[(long long)bi * np_ns + jk]];
    }
}
__global__ void group_points_grad_kernel(long long total, int c, int n, int np_ns, const float *__restrict__ grad_out,
                                         const int32_t *__restrict__ idx, float *__restrict__ grad_points) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int jk = (int)(i % np_ns);
        long long row = i / np_ns;
        int bi = (int)(row / c);
        atomicAdd(grad_points + row * n + idx[(long long)bi * np_ns + jk], grad_out[i]);
    }
}
inline int grid_for(long long total, int threads) {
    long long g = ceil_div_ll(total, threads);
    long long cap = (long long)kNumSMs * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
}  // namespace

extern "C" int istnet_gather_points(int b, int c, int n, int m, const float *points, const int32_t *idx, float *out, void *stream) {
    long long total = (long long)b * c * m;
    if (total <= 0) return ISTNET_OK;
    gather_points_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, n, m, points, idx, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, float *grad_points,
                                         void *stream) {
    long long total = (long long)b * c * m;
    if (total <= 0) return ISTNET_OK;
    gather_points_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, n, m, grad_out, idx, grad_points);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx, float *out,
                                   void *stream) {
    long long total = (long long)b * c * npoints * nsample;
    if (total <= 0) return ISTNET_OK;
    group_points_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints * nsample, points, idx, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int32_t *idx,
                                        float *grad_points, void *stream) {
    long long total = (long long)b * c * npoints * nsample;
    if (total <= 0) return ISTNET_OK;
    group_points_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, n, npoints * nsample, grad_out, idx,
                                                                                      grad_points);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// ---------------------------------------------------------------------------------------------------
// three-NN and three-interpolate (+ grad)
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int kNnThreads = 128;
// grid = (ceil(n/128), b); dynamic smem = m*3 floats.  One thread per unknown point, broadcast reads of `known`.
__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known, float *__restrict__ dist2,
                int32_t *__restrict__ idx, float *__restrict__ weight) {
    extern __shared__ float nn_known[];
    const int bi = blockIdx.y;
    const float *src = known + (size_t)bi * m * 3;
    for (int i = threadIdx.x; i < m * 3; i += kNnThreads) nn_known[i] = src[i];
    __syncthreads();
    int j = blockIdx.x * kNnThreads + threadIdx.x;
    if (j >= n) return;
    const float *u = unknown + ((size_t)bi * n + j) * 3;
    float ux = u[0], uy = u[1], uz = u[2];
    // The reference keeps double 1e40 sentinels (interpolate_gpu.cu:32); +inf in float orders identically
    // against every float d (inf and NaN are never selected in either formulation).
    const float inf = __int_as_float(0x7f800000);
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < m; ++k) {
        float d = sqdist_ref(__fsub_rn(ux, nn_known[k * 3 + 0]), __fsub_rn(uy, nn_known[k * 3 + 1]), __fsub_rn(uz, nn_known[k * 3 + 2]));
        if (d < b1) {
            b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
        } else if (d < b2) {
            b3 = b2; i3 = i2; b2 = d; i2 = k;
        } else if (d < b3) {
            b3 = d; i3 = k;
        }
    }
    size_t o = ((size_t)bi * n + j) * 3;
    if (dist2) { dist2[o + 0] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3; }
    idx[o + 0] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    if (weight) {
        // PointnetFPModule.forward (pointnet2_modules.py:185-188) on top of pointnet2_utils.py:142: dist = sqrt(dist2);
        // recip = 1 / (dist + 1e-8); weight = recip / sum(recip) — the same IEEE FP32 operations torch issues, in one place
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
        const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
        weight[o + 0] = __fdiv_rn(r1, norm); weight[o + 1] = __fdiv_rn(r2, norm); weight[o + 2] = __fdiv_rn(r3, norm);
    }
}

// out[(bi*c + l)*n + j] = fma(p[i3], w3, fma(p[i1], w1, p[i2]*w2)) (the reference's SASS contraction),  p = points + (bi*c + l)*m
__global__ void three_interpolate_kernel(long long total, int c, int m, int n, const float *__restrict__ points,
                                         const int32_t *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % n);
        long long row = i / n;
        int bi = (int)(row / c);
        long long o = ((long long)bi * n + j) * 3;
        const float *p = points + row * m;
        out[i] = __fmaf_rn(p[idx[o + 2]], weight[o + 2], __fmaf_rn(p[idx[o + 0]], weight[o + 0], __fmul_rn(p[idx[o + 1]], weight[o + 1])));
    }
}
__global__ void three_interpolate_grad_kernel(long long total, int c, int n, int m, const float *__restrict__ grad_out,
                                              const int32_t *__restrict__ idx, const float *__restrict__ weight,
                                              float *__restrict__ grad_points) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % n);
        long long row = i / n;
        int bi = (int)(row / c);
        long long o = ((long long)bi * n + j) * 3;
        float g = grad_out[i];
        float *gp = grad_points + row * m;
        atomicAdd(gp + idx[o + 0], __fmul_rn(g, weight[o + 0]));
        atomicAdd(gp + idx[o + 1], __fmul_rn(g, weight[o + 1]));
        atomicAdd(gp + idx[o + 2], __fmul_rn(g, weight[o + 2]));
    }
}
}  // namespace

extern "C" int istnet_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx, void *stream) {
    if (b <= 0 || n <= 0) return ISTNET_OK;
    if (m < 0) return ISTNET_ERR_BAD_ARG;
    size_t smem = (size_t)m * 3 * sizeof(float);
    if (smem > 200 * 1024) return ISTNET_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n, kNnThreads), b);
    three_nn_kernel<<<grid, kNnThreads, smem, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx, nullptr);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
// three_nn + the inverse-distance interpolation weights of PointnetFPModule.forward in ONE launch (the reference: three_nn kernel +
// sqrt, add, reciprocal, sum, div as five ATen launches).  dist2 may be null.
extern "C" int istnet_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx, float *weight,
                                       void *stream) {
    if (b <= 0 || n <= 0) return ISTNET_OK;
    if (m < 0 || !weight) return ISTNET_ERR_BAD_ARG;
    size_t smem = (size_t)m * 3 * sizeof(float);
    if (smem > 200 * 1024) return ISTNET_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n, kNnThreads), b);
    three_nn_kernel<<<grid, kNnThreads, smem, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx, weight);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx, const float *weight,
                                        float *out, void *stream) {
    long long total = (long long)b * c * n;
    if (total <= 0) return ISTNET_OK;
    three_interpolate_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, m, n, points, idx, weight, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, const float *weight,
                                             float *grad_points, void *stream) {
    long long total = (long long)b * c * n;
    if (total <= 0) return ISTNET_OK;
    three_interpolate_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(total, c, n, m, grad_out, idx, weight,
                                                                                           grad_points);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" int istnet_version(void) { return 100; }
extern "C" const char *istnet_strerror(int status) {
    if (status == ISTNET_OK) return "ok";
    if (status == ISTNET_ERR_BAD_ARG) return "istnet_b200: bad argument";
    if (status == ISTNET_ERR_UNSUPPORTED) return "istnet_b200: unsupported size";
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "istnet_b200: unknown error";
}
