// Fused set-abstraction level (PointnetSAModuleMSG, pointnet2_modules.py:29-73 + QueryAndGroup pointnet2_utils.py:317-377 +
// SharedMLP pytorch_utils.py:25-206 + F.max_pool2d over nsample) for the two fine levels of PointNet2MSG (modules.py:249-275),
// whose shared MLPs are narrow (3 -> 16 -> 16 -> 32 and 67 -> 32 -> 32 -> 64) and run on 0.8 M / 0.4 M grouped rows per batch of 32.
//
// Round 1 ran each (radius, nsample) scale as ~15 launches (ball query, gather, 2 tensor-core GEMMs on 128 x 16 tiles, 3 BatchNorm
// finalize + apply passes, max) that wrote and re-read every grouped activation: 1.0 ms of the 2.3 ms extractor forward for
// 4 % of its FLOPs.  Here ONE launch per pass covers both scales of a level and nothing grouped ever reaches HBM:
//   pass 0: ball query (warp-ballot scan of the cloud in shared memory, bit-exact with ball_query_gpu.cu:14-49) -> idx;
//           layer 0 on the fly (y0 = u[idx] + Wx (xyz_j - c_i), u = F Wf^T precomputed on the points) -> BatchNorm-0 statistics
//   pass 1: recompute layer 0, BN-0 + ReLU, layer 1 from shared memory                               -> BatchNorm-1 statistics
//   pass 2: recompute layers 0-1, layer 2; statistics of y2 and, per (centroid, channel), the max (gamma >= 0) or min (gamma < 0)
//           of y2 over the neighbours: relu(bn(.)) is monotone, so max_k relu(bn(y_k)) = relu(bn(max_k y_k))   (SURVEY §7 hard part 2)
//   final : out = relu(bn2(selected y2))  on [B*M, 2*C2] values.
// Train-mode BatchNorm needs the statistics of layer l over ALL rows before layer l+1 can start, hence the passes; recomputing
// the narrow prefix (<= 3 K FMA per row) is cheaper than storing it, and each pass's last CTA finishes the statistics itself
// (ticket.cuh), so a level costs 4 launches forward.  Backward: pre (sums of the max-routed gradient on [B*M, C2]) and three
// passes A/B/C that recompute the forward tile, apply the BatchNorm backward of layers 2/1/0, and accumulate the weight
// gradients per CTA in registers (fixed-order ticket reduction, no float atomics except the scatter to the points that the
// reference also performs with atomics, group_points_gpu.cu:48-69).
//
// Mapping: one lane = one output channel; a warp works alone on 16-row tiles (one neighbourhood of 16, or half of one of 32) held in
// its own shared-memory slots: the weight row of the lane lives in registers, the rows of the tile are broadcast from shared
// memory (LDS.128, same address for all lanes), so BatchNorm statistics and the max over neighbours are lane-local running
// values — no shuffles, no atomics — and there is no block-wide barrier before the final flush: the 12-16 warps of an SM drift
// apart and hide each other's gather / shared-memory latencies (the first version synchronised the CTA per 128-256-row tile and
// was latency-bound: profiles/r2_sa_fused_v1_launches.txt).  The forward pass scans the cloud ONCE per centroid for both radii.
#include <stdlib.h>

#include "common.cuh"
#include "ticket.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxNs = 32;

struct SaBn {
    const float *mean, *invstd, *gamma, *beta;
};
struct SaScale {
    float radius;
    int ns;
    int cta0, ncta;             // CTAs [cta0, cta0 + ncta) of the launch work on this scale
    int32_t *idx;               // [B*M*ns] ball-query result (written by the query pass, read by every other pass)
    const float *w0;            // layer-0 weight [C0][ldw0]: columns 0..2 multiply (xyz_j - c_i), the rest are folded into u
    int ldw0;
    const float *w1, *w2;       // [C1][C0], [C2][C1]
    SaBn bn0, bn1, bn2;
    float *part;                // forward: statistics partials of the layer the pass ends with;  backward: BN-backward partials
    FinP fin;
    float *ysel;                // [B*M][C2] selected (max or min over the neighbours) pre-BN layer-2 output
    uint8_t *asel;              // [B*M][C2] its (first) position in the neighbourhood
    // ---- backward
    const float *dz;            // gradient of the level output [B*M][ld_dz], this scale's channels at off_dz
    int ld_dz, off_dz;
    const double *ws2, *ws1, *ws0;  // [3C] BatchNorm-backward sums (sum g | sum g*xhat | -) of layers 2, 1, 0
    float *g1, *g0;             // [rows][C1], [rows][C0]: ReLU-masked gradients w.r.t. the BN outputs of layers 1 / 0
    float *part_w;              // weight-gradient partials [ISTNET_FIN_ROWS][rows*cols]
    unsigned *tickets_w;
    float *dw;                  // destination of the finished weight gradient (row stride ld_dw)
    int ld_dw;
};
struct SaLevelP {
    int B, N, M;
    const float *xyz, *new_xyz;  // [B,N,3], [B,M,3]
    const float *u;              // [B*N][ldu] = F Wf^T of both scales (scale s at column s*C0); null: no input features
    int ldu;
    float *dU;                   // backward: [B*N][ldu], zeroed by the launcher
    SaScale sc[2];
};

__device__ __forceinline__ float bn_apply(float y, float m, float s, float g, float b) {
    return __fmaf_rn(__fmul_rn(__fsub_rn(y, m), s), g, b);  // (y - mean) * invstd * gamma + beta
}

// dot(z[0:K], w[0:K]) with z broadcast from shared memory (16-byte aligned) and w in registers.  Four independent partial sums
// (k mod 4), combined as (a0 + a1) + (a2 + a3): a fixed order, and 4 FMAs in flight per lane instead of one dependent chain.
template <int K>
__device__ __forceinline__ float row_dot(const float *__restrict__ z, const float (&w)[K]) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < K; k += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(z + k);
        a0 = __fmaf_rn(w[k], v.x, a0);
        a1 = __fmaf_rn(w[k + 1], v.y, a1);
        a2 = __fmaf_rn(w[k + 2], v.z, a2);
        a3 = __fmaf_rn(w[k + 3], v.w, a3);
    }
    return __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
}

// ---- per-lane BatchNorm parameters
struct LaneBn {
    float m, s, g, b;
};
__device__ __forceinline__ LaneBn lane_bn(const SaBn &bn, int c, bool ok) {
    LaneBn r{0.f, 1.f, 1.f, 0.f};
    if (ok && bn.mean) { r.m = bn.mean[c]; r.s = bn.invstd[c]; r.g = bn.gamma[c]; r.b = bn.beta[c]; }
    return r;
}
template <int K>
__device__ __forceinline__ void load_row(float (&w)[K], const float *src, int stride, bool ok) {
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = ok ? __ldg(src + (size_t)k * stride) : 0.f;
}
// Weight matrix W[CO][CI] (row-major, global) -> shared memory TRANSPOSED wT[ci*CO + co]: lane co then reads its row with
// consecutive-bank (conflict-free) loads.  Loading a row per lane straight from global memory touches 32 cache lines per
// instruction; repeated per 16-row tile that made the first warp-autonomous version L1-bound (profiles/r2_sa_fused_v2_launches.txt).
template <int CO, int CI>
__device__ __forceinline__ void stage_weight_t(float *wT, const float *__restrict__ w) {
    for (int i = threadIdx.x; i < CO * CI; i += blockDim.x) {
        const int co = i / CI, ci = i - co * CI;
        wT[ci * CO + co] = __ldg(w + i);
    }
}
template <int K>
__device__ __forceinline__ void load_row_t(float (&w)[K], const float *wT, int CO, int co, bool ok) {
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = ok ? wT[k * CO + co] : 0.f;
}
// y0 of one row for the lane's channel: same FMA chain as sa_gather_l0_kernel (elementwise.cu)
__device__ __forceinline__ float y0_val(float uv, const float *rel, const float (&wx)[3]) {
    return __fmaf_rn(wx[0], rel[0], __fmaf_rn(wx[1], rel[1], __fmaf_rn(wx[2], rel[2], uv)));
}

constexpr int kTileRows = 16;  // rows a warp works on at a time: one neighbourhood of 16, or half of one of 32

// The 16 rows [l0, l0 + 16) of neighbourhood bj: point rows and relative coordinates into the warp's shared-memory slots.
// `cloud` is the instance's point cloud in shared memory ([k*3 + d]: a stride of 3 words is conflict-free, while the same
// access pattern on global memory costs 12 sectors per load and made the ball query L1-bound).
__device__ __forceinline__ void tile_geometry(const SaLevelP &p, const float *cloud, int b, int bj, const int *idx16, int *src_s, float *rel_s) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < kTileRows) {
        const int k = idx16[lane];
        src_s[lane] = b * p.N + k;
        const float *c = p.new_xyz + (size_t)bj * 3;
        rel_s[lane * 4 + 0] = __fsub_rn(cloud[k * 3 + 0], __ldg(c + 0));
        rel_s[lane * 4 + 1] = __fsub_rn(cloud[k * 3 + 1], __ldg(c + 1));
        rel_s[lane * 4 + 2] = __fsub_rn(cloud[k * 3 + 2], __ldg(c + 2));
    }
    __syncwarp();
}
// all threads of the CTA: instance b's cloud -> shared memory (barriers on both sides)
__device__ __forceinline__ void stage_cloud(const SaLevelP &p, int b, float *cloud) {
    __syncthreads();
    const float *src = p.xyz + (size_t)b * p.N * 3;
    for (int i = threadIdx.x; i < p.N * 3; i += blockDim.x) cloud[i] = __ldg(src + i);
    __syncthreads();
}
// layer 0 of the 16 rows for the lane's channel: y[r] (the u gathers of all rows are issued before the first use)
__device__ __forceinline__ void tile_y0(const SaLevelP &p, int s_off, bool ok, const int *src_s, const float *rel_s, const float (&wx)[3], float (&y)[kTileRows]) {
    float uv[kTileRows];
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) uv[r] = (ok && p.u) ? __ldg(p.u + (size_t)src_s[r] * p.ldu + s_off) : 0.f;
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) y[r] = y0_val(uv[r], rel_s + r * 4, wx);
}

// Sum of per-lane partials over the CTA's warps (fixed order) into this CTA's partial row: lane l of every warp holds quantity a of
// channel q*32 + l in acc[q][a].  red: >= nwarps * NACC * NQ * 32 floats of shared memory.
template <int NACC, int NQ>
__device__ __forceinline__ void flush_sums(float *red, const float (&acc)[NQ][NACC], int C, float *part, int cta, int G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int a = 0; a < NACC; ++a) red[((warp * NACC + a) * NQ + q) * 32 + lane] = acc[q][a];
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
        const int a = i / C, c = i - a * C;
        float t = 0.f;
        for (int w = 0; w < nwarps; ++w) t += red[((w * NACC + a) * NQ + (c >> 5)) * 32 + (c & 31)];
        part[((size_t)a * G + cta) * C + c] = t;
    }
}

// ------------------------------------------------------------------------------------------------ forward
// PASS 0: layer-0 statistics;  PASS 1: layer-1 statistics;  PASS 2: layer-2 statistics + selection.  QUERY: run the ball query.
// Warp-autonomous: a warp takes one centroid, scans the cloud ONCE for both radii, then runs the two neighbourhoods (16 and 32
// rows) through the layers in 16-row tiles held in its own shared-memory slots.  No block-wide barrier until the final flush of
// the statistics, so the warps of an SM drift apart and hide each other's gather / shared-memory latencies.
constexpr int kFwdThreads = 256;
template <int C0, int C1, int C2, int PASS, bool QUERY>
__global__ void __launch_bounds__(kFwdThreads, 2) sa_fwd_kernel(const SaLevelP p) {
    static_assert(C0 <= 32 && C1 <= 32 && C2 % 32 == 0 && C2 <= 64, "fused set abstraction: narrow levels only");
    constexpr int NW = kFwdThreads / 32;
    constexpr int NS2 = C2 / 32;
    constexpr int WF = 48 + 16 + 64 + kTileRows * C0 + kTileRows * C1;  // floats of shared memory per warp
    constexpr int WT = C0 * C1 + C1 * C2;                                // transposed weights of one scale
    extern __shared__ __align__(16) float smem_dyn[];
    float *wts = smem_dyn;              // [2][WT]
    float *smem = smem_dyn + 2 * WT;    // [NW][WF] + red
    float *cloud = smem + NW * WF + NW * 2 * NS2 * 32;  // [N*3]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (PASS >= 1) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            stage_weight_t<C1, C0>(wts + s * WT, p.sc[s].w1);
            if (PASS >= 2) stage_weight_t<C2, C1>(wts + s * WT + C0 * C1, p.sc[s].w2);
        }
        __syncthreads();
    }
    float *ws = smem + warp * WF;
    int *idx_s = reinterpret_cast<int *>(ws);        // [48]: scale 0 rows 0..15, scale 1 rows 16..47
    int *src_s = reinterpret_cast<int *>(ws + 48);   // [16]
    float *rel_s = ws + 64;                          // [16][4]
    float *z0 = ws + 128;                            // [16][C0]
    float *z1 = z0 + kTileRows * C0;                 // [16][C1]
    float *red = smem + NW * WF;
    const int G = gridDim.x, cta = blockIdx.x;
    const int BM = p.B * p.M;
    const int per = (BM + G - 1) / G;
    const int j_begin = min(BM, cta * per), j_end = min(BM, j_begin + per);
    const bool ok0 = lane < C0, ok1 = lane < C1;
    const float r2a = __fmul_rn(p.sc[0].radius, p.sc[0].radius), r2b = __fmul_rn(p.sc[1].radius, p.sc[1].radius);  // ball_query_gpu.cu:27
    const unsigned lt_mask = (1u << lane) - 1u;
    float acc[2][NS2][2];  // [scale][slice][sum, sum of squares] of the pass's layer for the lane's channel(s)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int q = 0; q < NS2; ++q) acc[s][q][0] = acc[s][q][1] = 0.f;

    for (int seg = j_begin; seg < j_end;) {  // the CTA's centroids, one instance at a time (its cloud staged in shared memory)
    const int b = seg / p.M;
    const int seg_end = min(j_end, (b + 1) * p.M);
    stage_cloud(p, b, cloud);
    for (int bj = seg + warp; bj < seg_end; bj += NW) {
        __syncwarp();
        if (QUERY) {
            const float *q = p.new_xyz + (size_t)bj * 3;
            const float cx = __ldg(q), cy = __ldg(q + 1), cz = __ldg(q + 2);
            int32_t *out0 = p.sc[0].idx + (size_t)bj * 16, *out1 = p.sc[1].idx + (size_t)bj * 32;
            int cnt0 = 0, cnt1 = 0, first0 = 0, first1 = 0;
            for (int base = 0; base < p.N && (cnt0 < 16 || cnt1 < 32); base += 32) {
                const int k = base + lane;
                bool h0 = false, h1 = false;
                if (k < p.N) {
                    const float d2 = sqdist_ref(__fsub_rn(cx, cloud[k * 3 + 0]), __fsub_rn(cy, cloud[k * 3 + 1]), __fsub_rn(cz, cloud[k * 3 + 2]));
                    h0 = d2 < r2a;
                    h1 = d2 < r2b;
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                if (cnt0 < 16 && m0) {  // the reference stops looking once nsample hits are stored (ball_query_gpu.cu:34-46)
                    if (cnt0 == 0) first0 = base + __ffs(m0) - 1;
                    const int pos = cnt0 + __popc(m0 & lt_mask);
                    if (h0 && pos < 16) { out0[pos] = k; idx_s[pos] = k; }
                    cnt0 += __popc(m0);
                }
                if (cnt1 < 32 && m1) {
                    if (cnt1 == 0) first1 = base + __ffs(m1) - 1;
                    const int pos = cnt1 + __popc(m1 & lt_mask);
                    if (h1 && pos < 32) { out1[pos] = k; idx_s[16 + pos] = k; }
                    cnt1 += __popc(m1);
                }
            }
            // tail: the reference pre-fills the row with the first hit (ball_query_gpu.cu:39-43); no hit => zeros
            for (int l = min(cnt0, 16) + lane; l < 16; l += 32) { out0[l] = first0; idx_s[l] = first0; }
            for (int l = min(cnt1, 32) + lane; l < 32; l += 32) { out1[l] = first1; idx_s[16 + l] = first1; }
        } else {
            if (lane < 16) idx_s[lane] = p.sc[0].idx[(size_t)bj * 16 + lane];
            idx_s[16 + lane] = p.sc[1].idx[(size_t)bj * 32 + lane];
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const SaScale &sc = p.sc[s];
            float wx[3] = {0.f, 0.f, 0.f};
            if (ok0) { wx[0] = __ldg(sc.w0 + (size_t)lane * sc.ldw0); wx[1] = __ldg(sc.w0 + (size_t)lane * sc.ldw0 + 1); wx[2] = __ldg(sc.w0 + (size_t)lane * sc.ldw0 + 2); }
            const LaneBn b0 = lane_bn(sc.bn0, lane, ok0 && PASS >= 1);
            const LaneBn b1 = lane_bn(sc.bn1, lane, ok1 && PASS >= 2);
            float best[NS2];
            int bi[NS2];
#pragma unroll
            for (int q = 0; q < NS2; ++q) { best[q] = 0.f; bi[q] = 0; }
            for (int h = 0; h <= s; ++h) {  // scale 0: one 16-row tile; scale 1: two
                tile_geometry(p, cloud, b, bj, idx_s + s * 16 + h * kTileRows, src_s, rel_s);
                {
                    float y[kTileRows];
                    tile_y0(p, s * C0 + lane, ok0, src_s, rel_s, wx, y);
                    if (ok0) {
#pragma unroll
                        for (int r = 0; r < kTileRows; ++r) {
                            if (PASS == 0) { acc[s][0][0] += y[r]; acc[s][0][1] += y[r] * y[r]; }
                            else z0[r * C0 + lane] = fmaxf(bn_apply(y[r], b0.m, b0.s, b0.g, b0.b), 0.f);
                        }
                    }
                }
                if (PASS == 0) continue;
                __syncwarp();
                {
                    float w1[C0];
                    load_row_t<C0>(w1, wts + s * WT, C1, lane, ok1);
                    if (ok1) {
#pragma unroll 4
                        for (int r = 0; r < kTileRows; ++r) {
                            const float y = row_dot<C0>(z0 + r * C0, w1);
                            if (PASS == 1) { acc[s][0][0] += y; acc[s][0][1] += y * y; }
                            else z1[r * C1 + lane] = fmaxf(bn_apply(y, b1.m, b1.s, b1.g, b1.b), 0.f);
                        }
                    }
                }
                if (PASS == 1) continue;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < NS2; ++q) {
                    const int c2 = q * 32 + lane;
                    float w2[C1];
                    load_row_t<C1>(w2, wts + s * WT + C0 * C1, C2, c2, true);
                    const bool pick_max = __ldg(sc.bn2.gamma + c2) >= 0.f;
#pragma unroll 4
                    for (int r = 0; r < kTileRows; ++r) {
                        const float y = row_dot<C1>(z1 + r * C1, w2);
                        acc[s][q][0] += y; acc[s][q][1] += y * y;
                        const int l = h * kTileRows + r;
                        const bool take = (l == 0) || (pick_max ? (y > best[q]) : (y < best[q]));  // strict: the first extreme wins (max_pool2d)
                        if (take) { best[q] = y; bi[q] = l; }
                    }
                }
            }
            if (PASS == 2) {
#pragma unroll
                for (int q = 0; q < NS2; ++q) {
                    sc.ysel[(size_t)bj * C2 + q * 32 + lane] = best[q];
                    sc.asel[(size_t)bj * C2 + q * 32 + lane] = (uint8_t)bi[q];
                }
            }
        }
    }
    seg = seg_end;
    }
    constexpr int CS = PASS == 0 ? C0 : (PASS == 1 ? C1 : C2);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const SaScale &sc = p.sc[s];
        if (sc.part == nullptr) continue;  // running statistics (eval): nothing to reduce
        flush_sums<2, NS2>(red, acc[s], CS, sc.part, cta, G);
        ticket_finish<2>(sc.fin, sc.part, G, CS, cta);
    }
}

// out[bj][off + c] = relu(bn2(ysel[bj][c])) for both scales
template <int C2>
__global__ void __launch_bounds__(256) sa_final_kernel(int BM, SaScale s0, SaScale s1, float *out, int ld_out) {
    const int total = BM * 2 * C2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % (2 * C2), bj = i / (2 * C2);
        const SaScale &sc = c < C2 ? s0 : s1;
        const int cc = c < C2 ? c : c - C2;
        const float y = sc.ysel[(size_t)bj * C2 + cc];
        out[(size_t)bj * ld_out + c] = fmaxf(bn_apply(y, sc.bn2.mean[cc], sc.bn2.invstd[cc], sc.bn2.gamma[cc], sc.bn2.beta[cc]), 0.f);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// The CTAs of a backward launch are split between the two scales (cta0 / ncta) because the weight-gradient accumulators of a
// scale live in registers for the whole kernel.
struct BwdCtx {
    int s, cta, G;
};
__device__ __forceinline__ BwdCtx bwd_ctx(const SaLevelP &p) {
    BwdCtx c;
    c.s = ((int)blockIdx.x >= p.sc[1].cta0 && p.sc[1].ncta > 0) ? 1 : 0;
    c.cta = (int)blockIdx.x - p.sc[c.s].cta0;
    c.G = p.sc[c.s].ncta;
    return c;
}
// pre: BatchNorm-backward sums of layer 2.  The gradient of the max over the neighbours is non-zero on one row per (centroid,
// channel), so sum g and sum g*xhat run over [B*M, C2] values only: g = dz * [bn2(ysel) > 0], xhat = (ysel - mean) * invstd.
template <int C2>
__global__ void __launch_bounds__(256, 2) sa_bwd_pre_kernel(const SaLevelP p) {
    constexpr int NS2 = C2 / 32;
    __shared__ float red[8 * 3 * NS2 * 32];
    const BwdCtx bc = bwd_ctx(p);
    const SaScale &sc = p.sc[bc.s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[NS2][3];
    LaneBn b2[NS2];
#pragma unroll
    for (int q = 0; q < NS2; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = 0.f; b2[q] = lane_bn(sc.bn2, q * 32 + lane, true); }
    const int BM = p.B * p.M;
    for (int bj = bc.cta * 8 + warp; bj < BM; bj += bc.G * 8) {
#pragma unroll
        for (int q = 0; q < NS2; ++q) {
            const int c = q * 32 + lane;
            const float y = sc.ysel[(size_t)bj * C2 + c];
            const float xh = __fmul_rn(__fsub_rn(y, b2[q].m), b2[q].s);
            const float u = __fmaf_rn(xh, b2[q].g, b2[q].b);
            const float g = u > 0.f ? sc.dz[(size_t)bj * sc.ld_dz + sc.off_dz + c] : 0.f;
            acc[q][0] += g;
            acc[q][1] += g * xh;
        }
    }
    flush_sums<3, NS2>(red, acc, C2, sc.part, bc.cta, bc.G);
    ticket_finish<3>(sc.fin, sc.part, bc.G, C2, bc.cta);
}

// Weight-gradient accumulators of one warp (lane = input channel ci < CI, one register per output channel co < CO) are combined
// across the CTA's warps in a fixed order and written as this CTA's partial row part[cta][co*CI + ci]; the launcher then sums the
// rows with the channel-parallel finalize kernel (a CO*CI-wide reduction is too long for a one-CTA tail).
template <int CO, int CI>
__device__ void flush_weight_grad(float *scratch /* >= nwarps*CO*CI floats of shared memory */, const float (&aw)[CO], bool ok, float *part, int cta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    constexpr int CW = CO * CI;
    __syncthreads();
    if (ok) {
#pragma unroll
        for (int co = 0; co < CO; ++co) scratch[(size_t)warp * CW + co * CI + lane] = aw[co];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CW; i += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nwarps; ++w) t += scratch[(size_t)w * CW + i];
        part[(size_t)cta * CW + i] = t;
    }
}

// STAGE 0 ("A"): layer 2: dy2 = BN-backward of the max-routed gradient; g1 = (dy2 W2) * [z1 > 0] -> sc.g1, sums of layer 1, dW2.
// STAGE 1 ("B"): layer 1: dy1 from g1; g0 = (dy1 W1) * [z0 > 0] -> sc.g0, sums of layer 0, dW1.
// STAGE 2 ("C"): layer 0: dy0 from g0; dU[point] += dy0 (atomics, as group_points_grad), dWx.
// Warp-autonomous 16-row tiles like the forward pass.  Weight rows / columns are (re)loaded from global memory (L1-resident,
// <= 8 KB) at the start of the phase that uses them, so that only the weight-gradient accumulators live in registers for the
// whole kernel.
constexpr int kBwdThreads = 128;
template <int C0, int C1, int C2, int STAGE>
__global__ void __launch_bounds__(kBwdThreads, 3) sa_bwd_kernel(const SaLevelP p) {
    static_assert(C0 <= C1 && C1 <= 32 && C2 % 32 == 0 && C2 <= 64, "fused set abstraction: narrow levels only");
    constexpr int NW = kBwdThreads / 32;
    constexpr int NS2 = C2 / 32;
    constexpr int WF = 16 + 64 + kTileRows * (C0 + 2 * C1 + C2);  // floats of shared memory per warp
    constexpr int CWMAX = C2 * C1;
    constexpr int SF = (NW * WF > NW * CWMAX ? NW * WF : NW * CWMAX);
    constexpr int WT = C0 * C1 + C1 * C2;  // transposed weights of this CTA's scale
    extern __shared__ __align__(16) float smem_dyn[];
    float *w1T = smem_dyn, *w2T = smem_dyn + C0 * C1;
    float *smem = smem_dyn + WT;  // [SF] tiles / weight-gradient scratch + red
    float *cloud = smem + SF + NW * 3 * 32;  // [N*3]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ws = smem + warp * WF;
    int *src_s = reinterpret_cast<int *>(ws);  // [16]
    float *rel_s = ws + 16;                    // [16][4]
    float *z0 = ws + 80;                       // [16][C0]
    float *y1 = z0 + kTileRows * C0;           // [16][C1]  (stage 1: y0 parked here)
    float *z1 = y1 + kTileRows * C1;           // [16][C1]
    float *d2 = z1 + kTileRows * C1;           // [16][C2]  (stage 1: dy1 [16][C1])
    float *red = smem + SF;
    const BwdCtx bc = bwd_ctx(p);
    const SaScale &sc = p.sc[bc.s];
    const int ns = sc.ns, tiles_per_group = ns / kTileRows;
    const int n_tiles = p.B * p.M * tiles_per_group;
    const int per = (n_tiles + bc.G - 1) / bc.G;
    const int t_begin = min(n_tiles, bc.cta * per), t_end = min(n_tiles, t_begin + per);
    const double invP = 1.0 / ((double)p.B * p.M * ns);
    const bool ok0 = lane < C0, ok1 = lane < C1;
    float wx[3] = {0.f, 0.f, 0.f};
    if (ok0) { wx[0] = sc.w0[(size_t)lane * sc.ldw0]; wx[1] = sc.w0[(size_t)lane * sc.ldw0 + 1]; wx[2] = sc.w0[(size_t)lane * sc.ldw0 + 2]; }
    const LaneBn b0 = lane_bn(sc.bn0, lane, ok0);
    const LaneBn b1 = lane_bn(sc.bn1, lane, ok1);
    if (STAGE <= 1) {
        stage_weight_t<C1, C0>(w1T, sc.w1);
        if (STAGE == 0) stage_weight_t<C2, C1>(w2T, sc.w2);
        __syncthreads();
    }

    if (STAGE == 2) {
        // ---------------- layer 0: BatchNorm backward, scatter to the points, dWx
        const float mg = ok0 ? (float)(sc.ws0[lane] * invP) : 0.f, mgx = ok0 ? (float)(sc.ws0[C0 + lane] * invP) : 0.f;
        float aw[1][3] = {{0.f, 0.f, 0.f}};
        for (int seg = t_begin; seg < t_end;) {  // the CTA's tiles, one instance at a time (its cloud staged in shared memory)
        const int b = seg / (p.M * tiles_per_group);
        const int seg_end = min(t_end, (b + 1) * p.M * tiles_per_group);
        stage_cloud(p, b, cloud);
        for (int tile = seg + warp; tile < seg_end; tile += NW) {
            const int bj = tile / tiles_per_group, l0 = (tile - bj * tiles_per_group) * kTileRows;
            const size_t row0 = (size_t)bj * ns + l0;
            __syncwarp();
            if (lane < kTileRows) d2[lane] = __int_as_float(sc.idx[row0 + lane]);
            tile_geometry(p, cloud, b, bj, reinterpret_cast<const int *>(d2), src_s, rel_s);
            float y[kTileRows], g[kTileRows];
            tile_y0(p, bc.s * C0 + lane, ok0, src_s, rel_s, wx, y);
#pragma unroll
            for (int r = 0; r < kTileRows; ++r) g[r] = ok0 ? __ldg(sc.g0 + (row0 + r) * C0 + lane) : 0.f;
            if (ok0) {
#pragma unroll
                for (int r = 0; r < kTileRows; ++r) {
                    const float xh = __fmul_rn(__fsub_rn(y[r], b0.m), b0.s);
                    const float dy = b0.g * b0.s * (g[r] - mg - xh * mgx);
                    aw[0][0] += dy * rel_s[r * 4 + 0]; aw[0][1] += dy * rel_s[r * 4 + 1]; aw[0][2] += dy * rel_s[r * 4 + 2];
                    if (p.dU) atomicAdd(p.dU + (size_t)src_s[r] * p.ldu + bc.s * C0 + lane, dy);
                }
            }
        }
        seg = seg_end;
        }
        // dWx[c][d] -> dw[c*ld + d]: per-CTA partial laid out [d][c] (quantity d, channel c), finished by the scale's last CTA
        constexpr int CW = 3 * 32;
        float *part2 = sc.part_w + (size_t)kMaxPartialRows * CW;
        flush_sums<3, 1>(red, aw, 32, sc.part_w, bc.cta, bc.G);  // part_w[(d*G + cta)*32 + c]
        if (!ticket_reduce<3>(sc.part_w, bc.cta, bc.G, 32, sc.tickets_w, part2)) return;
        for (int i = threadIdx.x; i < 3 * C0; i += blockDim.x) {
            const int c = i / 3, d = i - c * 3;
            sc.dw[(size_t)c * sc.ld_dw + d] = (float)ticket_total(part2, bc.G, 32, d, c);
        }
        return;
    }

    if (STAGE == 1) {
        // ---------------- layer 1: dy1 from the stored g1; dz0 = dy1 W1; g0; sums of layer 0; dW1
        const float mg1 = ok1 ? (float)(sc.ws1[lane] * invP) : 0.f, mgx1 = ok1 ? (float)(sc.ws1[C1 + lane] * invP) : 0.f;
        float acc[1][3] = {{0.f, 0.f, 0.f}};
        float aw[C1];  // dW1[c1][c0] for the lane's c0
#pragma unroll
        for (int k = 0; k < C1; ++k) aw[k] = 0.f;
        float *dy1 = d2;   // [16][C1]
        float *y0s = y1;   // [16][C0]
        for (int seg = t_begin; seg < t_end;) {  // the CTA's tiles, one instance at a time (its cloud staged in shared memory)
        const int b = seg / (p.M * tiles_per_group);
        const int seg_end = min(t_end, (b + 1) * p.M * tiles_per_group);
        stage_cloud(p, b, cloud);
        for (int tile = seg + warp; tile < seg_end; tile += NW) {
            const int bj = tile / tiles_per_group, l0 = (tile - bj * tiles_per_group) * kTileRows;
            const size_t row0 = (size_t)bj * ns + l0;
            __syncwarp();
            if (lane < kTileRows) z1[lane] = __int_as_float(sc.idx[row0 + lane]);
            tile_geometry(p, cloud, b, bj, reinterpret_cast<const int *>(z1), src_s, rel_s);
            {
                float y[kTileRows];
                tile_y0(p, bc.s * C0 + lane, ok0, src_s, rel_s, wx, y);
                if (ok0) {
#pragma unroll
                    for (int r = 0; r < kTileRows; ++r) {
                        y0s[r * C0 + lane] = y[r];
                        z0[r * C0 + lane] = fmaxf(bn_apply(y[r], b0.m, b0.s, b0.g, b0.b), 0.f);
                    }
                }
            }
            __syncwarp();
            {
                float w1[C0];  // row of W1 for the lane's output channel c1
                load_row_t<C0>(w1, w1T, C1, lane, ok1);
                float g[kTileRows];
#pragma unroll
                for (int r = 0; r < kTileRows; ++r) g[r] = ok1 ? __ldg(sc.g1 + (row0 + r) * C1 + lane) : 0.f;
                if (ok1) {
#pragma unroll 4
                    for (int r = 0; r < kTileRows; ++r) {
                        const float y = row_dot<C0>(z0 + r * C0, w1);
                        const float xh = __fmul_rn(__fsub_rn(y, b1.m), b1.s);
                        dy1[r * C1 + lane] = b1.g * b1.s * (g[r] - mg1 - xh * mgx1);
                    }
                }
            }
            __syncwarp();
            {
                float w1c[C1];  // column of W1 for the lane's input channel c0: dz0[c0] = sum_c1 dy1[c1] * W1[c1][c0]
                load_row<C1>(w1c, sc.w1 + lane, C0, ok0);
                if (ok0) {
#pragma unroll 2
                    for (int r = 0; r < kTileRows; ++r) {
                        const float dz0 = row_dot<C1>(dy1 + r * C1, w1c);
                        const float z = z0[r * C0 + lane];
                        const float g = z > 0.f ? dz0 : 0.f;
                        sc.g0[(row0 + r) * C0 + lane] = g;
                        const float xh = __fmul_rn(__fsub_rn(y0s[r * C0 + lane], b0.m), b0.s);
                        acc[0][0] += g; acc[0][1] += g * xh;
#pragma unroll
                        for (int k = 0; k < C1; k += 4) {  // dW1[:, c0] += dy1[r, :] * z0[r, c0]
                            const float4 v = *reinterpret_cast<const float4 *>(dy1 + r * C1 + k);
                            aw[k] = __fmaf_rn(v.x, z, aw[k]); aw[k + 1] = __fmaf_rn(v.y, z, aw[k + 1]);
                            aw[k + 2] = __fmaf_rn(v.z, z, aw[k + 2]); aw[k + 3] = __fmaf_rn(v.w, z, aw[k + 3]);
                        }
                    }
                }
            }
        }
        seg = seg_end;
        }
        flush_sums<3, 1>(red, acc, C0, sc.part, bc.cta, bc.G);
        ticket_finish<3>(sc.fin, sc.part, bc.G, C0, bc.cta);
        flush_weight_grad<C1, C0>(smem, aw, ok0, sc.part_w, bc.cta);
        return;
    }

    // ---------------- STAGE 0: layer 2
    {
        float acc[1][3] = {{0.f, 0.f, 0.f}};
        float aw[C2];  // dW2[c2][c1] for the lane's c1
#pragma unroll
        for (int k = 0; k < C2; ++k) aw[k] = 0.f;
        for (int seg = t_begin; seg < t_end;) {  // the CTA's tiles, one instance at a time (its cloud staged in shared memory)
        const int b = seg / (p.M * tiles_per_group);
        const int seg_end = min(t_end, (b + 1) * p.M * tiles_per_group);
        stage_cloud(p, b, cloud);
        for (int tile = seg + warp; tile < seg_end; tile += NW) {
            const int bj = tile / tiles_per_group, l0 = (tile - bj * tiles_per_group) * kTileRows;
            const size_t row0 = (size_t)bj * ns + l0;
            __syncwarp();
            if (lane < kTileRows) z1[lane] = __int_as_float(sc.idx[row0 + lane]);
            tile_geometry(p, cloud, b, bj, reinterpret_cast<const int *>(z1), src_s, rel_s);
            {
                float y[kTileRows];
                tile_y0(p, bc.s * C0 + lane, ok0, src_s, rel_s, wx, y);
                if (ok0) {
#pragma unroll
                    for (int r = 0; r < kTileRows; ++r) z0[r * C0 + lane] = fmaxf(bn_apply(y[r], b0.m, b0.s, b0.g, b0.b), 0.f);
                }
            }
            __syncwarp();
            {
                float w1[C0];
                load_row_t<C0>(w1, w1T, C1, lane, ok1);
                if (ok1) {
#pragma unroll 4
                    for (int r = 0; r < kTileRows; ++r) {
                        const float y = row_dot<C0>(z0 + r * C0, w1);
                        y1[r * C1 + lane] = y;
                        z1[r * C1 + lane] = fmaxf(bn_apply(y, b1.m, b1.s, b1.g, b1.b), 0.f);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < NS2; ++q) {
                const int c2 = q * 32 + lane;
                float w2[C1];  // row of W2 for the lane's layer-2 channel
                load_row_t<C1>(w2, w2T, C2, c2, true);
                const LaneBn b2 = lane_bn(sc.bn2, c2, true);
                const float mg2 = (float)(sc.ws2[c2] * invP), mgx2 = (float)(sc.ws2[C2 + c2] * invP);
                const int lsel = (int)sc.asel[(size_t)bj * C2 + c2] - l0;  // row of this tile that holds the selected neighbour (if any)
                const float dzv = __ldg(sc.dz + (size_t)bj * sc.ld_dz + sc.off_dz + c2);
#pragma unroll 4
                for (int r = 0; r < kTileRows; ++r) {
                    const float y = row_dot<C1>(z1 + r * C1, w2);
                    const float xh = __fmul_rn(__fsub_rn(y, b2.m), b2.s);
                    float gg = 0.f;
                    if (r == lsel && __fmaf_rn(xh, b2.g, b2.b) > 0.f) gg = dzv;
                    d2[r * C2 + c2] = b2.g * b2.s * (gg - mg2 - xh * mgx2);
                }
            }
            __syncwarp();
            {
                // dz1[c1] = sum_c2 dy2[c2] * W2[c2][c1] for the lane's layer-1 channel c1, 32 layer-2 channels at a time (one 32-entry
                // block of W2's column c1 in registers), together with dW2[c2][c1] += dy2[r][c2] * z1[r][c1]
                float *dz1 = z0;  // [16][C1]: the z0 tile is dead after the layer-1 recompute (C0 == C1 on the supported levels)
                static_assert(C0 == C1, "dz1 partial sums reuse the z0 tile");
#pragma unroll
                for (int hq = 0; hq < NS2; ++hq) {
                    float w2c[32];
                    load_row<32>(w2c, sc.w2 + (size_t)(hq * 32) * C1 + lane, C1, ok1);
                    if (ok1) {
#pragma unroll 2
                        for (int r = 0; r < kTileRows; ++r) {
                            const float part = row_dot<32>(d2 + r * C2 + hq * 32, w2c);
                            dz1[r * C1 + lane] = hq == 0 ? part : dz1[r * C1 + lane] + part;
                            const float z = z1[r * C1 + lane];
#pragma unroll
                            for (int k = 0; k < 32; k += 4) {
                                const float4 v = *reinterpret_cast<const float4 *>(d2 + r * C2 + hq * 32 + k);
                                aw[hq * 32 + k] = __fmaf_rn(v.x, z, aw[hq * 32 + k]); aw[hq * 32 + k + 1] = __fmaf_rn(v.y, z, aw[hq * 32 + k + 1]);
                                aw[hq * 32 + k + 2] = __fmaf_rn(v.z, z, aw[hq * 32 + k + 2]); aw[hq * 32 + k + 3] = __fmaf_rn(v.w, z, aw[hq * 32 + k + 3]);
                            }
                        }
                    }
                }
                if (ok1) {
#pragma unroll
                    for (int r = 0; r < kTileRows; ++r) {
                        const float g = z1[r * C1 + lane] > 0.f ? dz1[r * C1 + lane] : 0.f;
                        sc.g1[(row0 + r) * C1 + lane] = g;
                        const float xh = __fmul_rn(__fsub_rn(y1[r * C1 + lane], b1.m), b1.s);
                        acc[0][0] += g; acc[0][1] += g * xh;
                    }
                }
            }
        }
        seg = seg_end;
        }
        flush_sums<3, 1>(red, acc, C1, sc.part, bc.cta, bc.G);
        ticket_finish<3>(sc.fin, sc.part, bc.G, C1, bc.cta);
        flush_weight_grad<C2, C1>(smem, aw, ok1, sc.part_w, bc.cta);
    }
}

// ------------------------------------------------------------------------------------------------ pointwise GEMMs of layer 0
// u[r][s*C0 + c] = sum_k F[r][k] * w0_s[c][3 + k]   (r over the B*N points): the feature part of layer 0 for both scales
template <int K, int C0>
__global__ void __launch_bounds__(kThreads, 1) sa_u_kernel(int R, const float *__restrict__ F, const float *w0a, const float *w0b, int ldw0, float *u) {
    __shared__ __align__(16) float rows[kWarps][4][K];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CO = 2 * C0;          // 32 or 64
    constexpr int NSL = (CO + 31) / 32;
    float w[NSL][K];
#pragma unroll
    for (int q = 0; q < NSL; ++q) {
        const int j = q * 32 + lane;
        const float *src = j < C0 ? w0a + (size_t)j * ldw0 + 3 : w0b + (size_t)(j - C0) * ldw0 + 3;
#pragma unroll
        for (int k = 0; k < K; ++k) w[q][k] = j < CO ? src[k] : 0.f;
    }
    for (int r0 = (blockIdx.x * kWarps + warp) * 4; r0 < R; r0 += gridDim.x * kWarps * 4) {
        __syncwarp();
        for (int i = lane; i < 4 * K; i += 32) {
            const int rr = i / K, k = i - rr * K;
            rows[warp][rr][k] = (r0 + rr < R) ? F[(size_t)(r0 + rr) * K + k] : 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            if (r0 + rr < R) {
#pragma unroll
                for (int q = 0; q < NSL; ++q) {
                    const int j = q * 32 + lane;
                    const float y = row_dot<K>(rows[warp][rr], w[q]);
                    if (j < CO) u[(size_t)(r0 + rr) * CO + j] = y;
                }
            }
        }
    }
}
// backward of the above: dF[r][k] = sum_j dU[r][j] * Wf[j][k];  dWf[j][k] = sum_r dU[r][j] * F[r][k]  -> dw0_s[c][3 + k].
// lane = input channel k; the K channels are handled 32 at a time (one sweep over the rows per 32-channel slice) so that only
// one column block of Wf and one block of accumulators live in registers.
template <int K, int C0>
__global__ void __launch_bounds__(kThreads, 1) sa_u_bwd_kernel(int R, const float *__restrict__ F, const float *__restrict__ dU, const float *w0a,
                                                                const float *w0b, int ldw0, float *dF, float *part_w) {
    extern __shared__ __align__(16) float ub_smem[];
    constexpr int CO = 2 * C0;
    constexpr int NK = K / 32;
    constexpr int CW = CO * K;
    float *du_s = ub_smem;                       // [kWarps][4][CO]
    float *scratch = ub_smem + kWarps * 4 * CO;  // [kWarps][CO][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *mine = du_s + warp * 4 * CO;
    for (int q = 0; q < NK; ++q) {
        const int k = q * 32 + lane;
        float wc[CO];  // column k of Wf
#pragma unroll
        for (int j = 0; j < CO; ++j) wc[j] = (j < C0 ? w0a[(size_t)j * ldw0 + 3 + k] : w0b[(size_t)(j - C0) * ldw0 + 3 + k]);
        float aw[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) aw[j] = 0.f;
        for (int r0 = (blockIdx.x * kWarps + warp) * 4; r0 < R; r0 += gridDim.x * kWarps * 4) {
            __syncwarp();
            for (int i = lane; i < 4 * CO; i += 32) {
                const int rr = i / CO, j = i - rr * CO;
                mine[rr * CO + j] = (r0 + rr < R) ? dU[(size_t)(r0 + rr) * CO + j] : 0.f;
            }
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                if (r0 + rr < R) {
                    if (dF) dF[(size_t)(r0 + rr) * K + k] = row_dot<CO>(mine + rr * CO, wc);
                    const float f = F[(size_t)(r0 + rr) * K + k];
#pragma unroll
                    for (int j = 0; j < CO; j += 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(mine + rr * CO + j);
                        aw[j] = __fmaf_rn(v.x, f, aw[j]); aw[j + 1] = __fmaf_rn(v.y, f, aw[j + 1]);
                        aw[j + 2] = __fmaf_rn(v.z, f, aw[j + 2]); aw[j + 3] = __fmaf_rn(v.w, f, aw[j + 3]);
                    }
                }
            }
        }
        // this CTA's partial of dWf[:, q*32 .. q*32+31]: fixed-order sum over the warps
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CO; ++j) scratch[((size_t)warp * CO + j) * 32 + lane] = aw[j];
        __syncthreads();
        for (int i = threadIdx.x; i < CO * 32; i += kThreads) {
            const int j = i / 32, kk = i - j * 32;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += scratch[((size_t)w * CO + j) * 32 + kk];
            part_w[(size_t)blockIdx.x * CW + j * K + q * 32 + kk] = t;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host
static bool fill_scale(const istnet_sa_scale &h, int C0, int C1, int C2, int nacc, int Cstat, SaScale &d) {
    d = SaScale{};
    d.radius = h.radius; d.ns = h.nsample; d.idx = h.idx;
    d.w0 = h.w0; d.ldw0 = h.ldw0; d.w1 = h.w1; d.w2 = h.w2;
    d.bn0 = SaBn{h.bn_mean[0], h.bn_invstd[0], h.bn_gamma[0], h.bn_beta[0]};
    d.bn1 = SaBn{h.bn_mean[1], h.bn_invstd[1], h.bn_gamma[1], h.bn_beta[1]};
    d.bn2 = SaBn{h.bn_mean[2], h.bn_invstd[2], h.bn_gamma[2], h.bn_beta[2]};
    d.part = h.part;
    if (!make_fin(h.fin, h.part, nacc, Cstat, d.fin)) return false;
    d.ysel = h.ysel; d.asel = h.asel;
    d.dz = h.dz; d.ld_dz = h.ld_dz; d.off_dz = h.off_dz;
    d.ws2 = h.ws2; d.ws1 = h.ws1; d.ws0 = h.ws0;
    d.g1 = h.g1; d.g0 = h.g0;
    d.part_w = h.part_w; d.tickets_w = h.tickets_w; d.dw = h.dw; d.ld_dw = h.ld_dw;
    (void)C0; (void)C1; (void)C2;
    return h.nsample == 16 || h.nsample == 32;
}
// the kernels assume scale 0 = 16 neighbours, scale 1 = 32 (PointNet2MSG, modules.py:253,266)
static bool scales_ok(const SaLevelP &p) { return p.sc[0].ns == 16 && p.sc[1].ns == 32; }
static bool fill_level(int B, int N, int M, const float *xyz, const float *new_xyz, const float *u, int ldu, float *dU, SaLevelP &p) {
    if (B <= 0 || N <= 0 || M <= 0 || !xyz || !new_xyz) return false;
    if ((long long)B * M * kMaxNs > 0x7fffffffLL || (long long)B * N > 0x7fffffffLL / 4 || N > 8192) return false;  // cloud in shared memory: <= 96 KB
    p = SaLevelP{};
    p.B = B; p.N = N; p.M = M; p.xyz = xyz; p.new_xyz = new_xyz; p.u = u; p.ldu = ldu; p.dU = dU;
    return true;
}
// CTAs of a backward launch shared between the two scales in proportion to their rows (16 : 32 neighbours)
static void split_ctas(SaLevelP &p, int total, int rows_per_unit) {
    const int w0 = p.sc[0].ns, w1 = p.sc[1].ns;
    int g0 = total * w0 / (w0 + w1);
    if (g0 < 1) g0 = 1;
    int g1 = total - g0;
    if (g1 < 1) g1 = 1;
    const long long u0 = (long long)p.B * p.M * w0 / rows_per_unit, u1 = (long long)p.B * p.M * w1 / rows_per_unit;
    if (g0 > u0) g0 = (int)(u0 < 1 ? 1 : u0);
    if (g1 > u1) g1 = (int)(u1 < 1 ? 1 : u1);
    if (g0 > kMaxPartialRows) g0 = kMaxPartialRows;
    if (g1 > kMaxPartialRows) g1 = kMaxPartialRows;
    p.sc[0].cta0 = 0; p.sc[0].ncta = g0;
    p.sc[1].cta0 = g0; p.sc[1].ncta = g1;
}

// CTAs per SM of the persistent grids (tuning knobs, read once): ISTNET_SA_FWD_CTAS (default 2), ISTNET_SA_BWD_CTAS (default 3)
static int env_ctas(const char *name, int dflt) {
    const char *e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}
static int fwd_ctas_per_sm() { static const int v = env_ctas("ISTNET_SA_FWD_CTAS", 2); return v; }
static int bwd_ctas_per_sm() { static const int v = env_ctas("ISTNET_SA_BWD_CTAS", 3); return v; }

template <int C0, int C1, int C2>
static int launch_fwd(SaLevelP &p, int pass, int query, cudaStream_t st) {
    // warp-autonomous forward: one grid for both scales, every CTA's partial row belongs to both reductions
    int grid = kNumSMs * fwd_ctas_per_sm();
    const int units = (p.B * p.M + 7) / 8;  // at least one centroid per warp
    if (grid > units) grid = units;
    if (grid > kMaxPartialRows) grid = kMaxPartialRows;
    constexpr int NW = kFwdThreads / 32, NS2 = C2 / 32;
    const size_t smem = sizeof(float) * (2 * (C0 * C1 + C1 * C2) + NW * (48 + 16 + 64 + kTileRows * C0 + kTileRows * C1) + NW * 2 * NS2 * 32 + ((p.N * 3 + 3) & ~3));
#define SA_FWD(PASS, Q)                                                                                                          \
    do {                                                                                                                         \
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_fwd_kernel<C0, C1, C2, PASS, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_fwd_kernel<C0, C1, C2, PASS, Q><<<grid, kFwdThreads, smem, st>>>(p);                                                   \
    } while (0)
    if (pass == 0 && query) SA_FWD(0, true);
    else if (pass == 0) SA_FWD(0, false);
    else if (pass == 1 && !query) SA_FWD(1, false);
    else if (pass == 2 && query) SA_FWD(2, true);
    else if (pass == 2) SA_FWD(2, false);
    else return ISTNET_ERR_BAD_ARG;
#undef SA_FWD
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
template <int C0, int C1, int C2>
static int launch_bwd(SaLevelP &p, int stage, cudaStream_t st) {
    if (stage < 0) {
        split_ctas(p, kNumSMs * 2, 8);
        sa_bwd_pre_kernel<C2><<<p.sc[0].ncta + p.sc[1].ncta, 256, 0, st>>>(p);
        ISTNET_LAUNCH_CHECK();
        return ISTNET_OK;
    }
    split_ctas(p, kNumSMs * bwd_ctas_per_sm(), kTileRows * (kBwdThreads / 32));
    const int grid = p.sc[0].ncta + p.sc[1].ncta;
    constexpr int NW = kBwdThreads / 32;
    constexpr int WF = 16 + 64 + kTileRows * (C0 + 2 * C1 + C2), CWMAX = C2 * C1;
    constexpr int SF = (NW * WF > NW * CWMAX ? NW * WF : NW * CWMAX);
    const size_t smem = sizeof(float) * (C0 * C1 + C1 * C2 + SF + NW * 3 * 32 + ((p.N * 3 + 3) & ~3));
#define SA_BWD(STAGE)                                                                                                          \
    do {                                                                                                                       \
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_bwd_kernel<C0, C1, C2, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_bwd_kernel<C0, C1, C2, STAGE><<<grid, kBwdThreads, smem, st>>>(p);                                                   \
    } while (0)
    if (stage == 0) SA_BWD(0);
    else if (stage == 1) SA_BWD(1);
    else if (stage == 2) SA_BWD(2);
    else return ISTNET_ERR_BAD_ARG;
#undef SA_BWD
    ISTNET_LAUNCH_CHECK();
    if (stage <= 1) {  // weight gradient of the stage: fixed-order sum of the per-CTA partial rows (channel-parallel launch)
        const int CW = stage == 0 ? C2 * C1 : C1 * C0;
        for (int s = 0; s < 2; ++s) {
            FinP f{};
            f.kind = ISTNET_FIN_COLSUM;
            f.sum_f32 = p.sc[s].dw;
            int e = istnet_fin_finalize_launch(p.sc[s].part_w, p.sc[s].ncta, CW, 1, f, st);
            if (e) return e;
        }
    }
    return ISTNET_OK;
}

extern "C" int istnet_sa_level_supported(int C0, int C1, int C2) {
    return (C0 == 16 && C1 == 16 && C2 == 32) || (C0 == 32 && C1 == 32 && C2 == 64);
}

extern "C" int istnet_sa_level_forward(int B, int N, int M, int C0, int C1, int C2, const float *xyz, const float *new_xyz, const float *u, int ldu,
                                       const istnet_sa_scale *scales, int pass, int query, void *stream) {
    SaLevelP p;
    if (!scales || !fill_level(B, N, M, xyz, new_xyz, u, ldu, nullptr, p)) return ISTNET_ERR_BAD_ARG;
    const int Cstat = pass == 0 ? C0 : (pass == 1 ? C1 : C2);
    for (int s = 0; s < 2; ++s) {
        if (!fill_scale(scales[s], C0, C1, C2, 2, Cstat, p.sc[s])) return ISTNET_ERR_BAD_ARG;
        if (!p.sc[s].idx || !p.sc[s].w0 || (pass >= 1 && (!p.sc[s].w1 || !p.sc[s].bn0.mean)) ||
            (pass >= 2 && (!p.sc[s].w2 || !p.sc[s].bn1.mean || !p.sc[s].bn2.gamma || !p.sc[s].ysel || !p.sc[s].asel)))
            return ISTNET_ERR_BAD_ARG;
    }
    if ((u && ldu < 2 * C0) || !scales_ok(p)) return ISTNET_ERR_BAD_ARG;
    if (C0 == 16 && C1 == 16 && C2 == 32) return launch_fwd<16, 16, 32>(p, pass, query, (cudaStream_t)stream);
    if (C0 == 32 && C1 == 32 && C2 == 64) return launch_fwd<32, 32, 64>(p, pass, query, (cudaStream_t)stream);
    return ISTNET_ERR_UNSUPPORTED;
}

extern "C" int istnet_sa_level_final(int B, int M, int C2, const istnet_sa_scale *scales, float *out, int ld_out, void *stream) {
    if (!scales || B <= 0 || M <= 0 || !out || ld_out < 2 * C2) return ISTNET_ERR_BAD_ARG;
    SaScale s0, s1;
    if (!fill_scale(scales[0], 0, 0, C2, 2, C2, s0) || !fill_scale(scales[1], 0, 0, C2, 2, C2, s1)) return ISTNET_ERR_BAD_ARG;
    if (!s0.ysel || !s1.ysel || !s0.bn2.mean || !s1.bn2.mean) return ISTNET_ERR_BAD_ARG;
    const int total = B * M * 2 * C2;
    int grid = (total + kThreads - 1) / kThreads;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    if (C2 == 32) sa_final_kernel<32><<<grid, kThreads, 0, (cudaStream_t)stream>>>(B * M, s0, s1, out, ld_out);
    else if (C2 == 64) sa_final_kernel<64><<<grid, kThreads, 0, (cudaStream_t)stream>>>(B * M, s0, s1, out, ld_out);
    else return ISTNET_ERR_UNSUPPORTED;
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_sa_level_backward(int B, int N, int M, int C0, int C1, int C2, const float *xyz, const float *new_xyz, const float *u, int ldu,
                                        float *dU, const istnet_sa_scale *scales, int stage, void *stream) {
    SaLevelP p;
    if (!scales || !fill_level(B, N, M, xyz, new_xyz, u, ldu, dU, p)) return ISTNET_ERR_BAD_ARG;
    const int Cstat = stage < 0 ? C2 : (stage == 0 ? C1 : C0);
    for (int s = 0; s < 2; ++s) {
        if (!fill_scale(scales[s], C0, C1, C2, 3, Cstat, p.sc[s])) return ISTNET_ERR_BAD_ARG;
        const SaScale &d = p.sc[s];
        if (!d.idx || !d.w0 || !d.w1 || !d.w2 || !d.bn0.mean || !d.bn1.mean || !d.bn2.mean) return ISTNET_ERR_BAD_ARG;
        if (stage < 0 && (!d.dz || !d.ysel || !d.part || d.fin.kind != ISTNET_FIN_BN_BWD)) return ISTNET_ERR_BAD_ARG;
        if (stage == 0 && (!d.dz || !d.asel || !d.ws2 || !d.g1 || !d.part || d.fin.kind != ISTNET_FIN_BN_BWD || !d.part_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
        if (stage == 1 && (!d.ws1 || !d.g1 || !d.g0 || !d.part || d.fin.kind != ISTNET_FIN_BN_BWD || !d.part_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
        if (stage == 2 && (!d.ws0 || !d.g0 || !d.part_w || !d.tickets_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
    }
    if (!scales_ok(p)) return ISTNET_ERR_BAD_ARG;
    if (stage == 2 && dU) ISTNET_CUDA_TRY(cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)B * N * ldu, (cudaStream_t)stream));
    if (C0 == 16 && C1 == 16 && C2 == 32) return launch_bwd<16, 16, 32>(p, stage, (cudaStream_t)stream);
    if (C0 == 32 && C1 == 32 && C2 == 64) return launch_bwd<32, 32, 64>(p, stage, (cudaStream_t)stream);
    return ISTNET_ERR_UNSUPPORTED;
}

extern "C" int istnet_sa_u(int R, int K, int C0, const float *F, const float *w0a, const float *w0b, int ldw0, float *u, void *stream) {
    if (R <= 0 || !F || !w0a || !w0b || !u || ldw0 < 3 + K) return ISTNET_ERR_BAD_ARG;
    if (!(K == 64 && C0 == 32)) return ISTNET_ERR_UNSUPPORTED;
    int grid = (R + kWarps * 4 - 1) / (kWarps * 4);
    if (grid > kNumSMs) grid = kNumSMs;
    sa_u_kernel<64, 32><<<grid, kThreads, 0, (cudaStream_t)stream>>>(R, F, w0a, w0b, ldw0, u);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_sa_u_bwd(int R, int K, int C0, const float *F, const float *dU, const float *w0a, const float *w0b, int ldw0, float *dF,
                               float *part_w, float *dwf, void *stream) {
    if (R <= 0 || !F || !dU || !w0a || !w0b || !part_w || !dwf || ldw0 < 3 + K) return ISTNET_ERR_BAD_ARG;
    if (!(K == 64 && C0 == 32)) return ISTNET_ERR_UNSUPPORTED;
    int grid = (R + kWarps * 4 - 1) / (kWarps * 4);
    if (grid > kNumSMs) grid = kNumSMs;
    const size_t smem = (size_t)(kWarps * 4 * 64 + kWarps * 64 * 32) * sizeof(float);
    ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_u_bwd_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_u_bwd_kernel<64, 32><<<grid, kThreads, smem, (cudaStream_t)stream>>>(R, F, dU, w0a, w0b, ldw0, dF, part_w);
    ISTNET_LAUNCH_CHECK();
    FinP f{};
    f.kind = ISTNET_FIN_COLSUM;
    f.sum_f32 = dwf;  // [2*C0][K]: rows 0..C0-1 = scale 0, rest = scale 1
    return istnet_fin_finalize_launch(part_w, grid, 2 * C0 * K, 1, f, (cudaStream_t)stream);
}
