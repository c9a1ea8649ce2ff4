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
// Mapping: one lane = one output channel, one warp = one neighbourhood (forward) or a block of rows (backward); the weight row
// of the lane lives in registers, the rows of the current tile are broadcast from shared memory (LDS.128, same address for
// all lanes), so BatchNorm statistics and the max over neighbours are lane-local running values: no shuffles, no atomics.
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

struct TileCtx {
    int s;         // scale
    int cta, G;    // this CTA among the scale's CTAs
    int ns, T;     // neighbours per centroid, rows per tile
    int t_begin, t_end;
};
template <int GT>
__device__ __forceinline__ TileCtx tile_ctx(const SaLevelP &p) {
    TileCtx c;
    c.s = ((int)blockIdx.x >= p.sc[1].cta0 && p.sc[1].ncta > 0) ? 1 : 0;
    const SaScale &sc = p.sc[c.s];
    c.cta = (int)blockIdx.x - sc.cta0;
    c.G = sc.ncta;
    c.ns = sc.ns;
    c.T = GT * sc.ns;
    const int n_tiles = p.B * p.M / GT;
    const int per = (n_tiles + c.G - 1) / c.G;
    c.t_begin = min(n_tiles, c.cta * per);
    c.t_end = min(n_tiles, c.t_begin + per);
    return c;
}

// Shared-memory carve-up (floats unless noted); every region 16-byte aligned.
template <int C0, int C1, int C2, int GT>
struct Smem {
    static constexpr int T = GT * kMaxNs;
    float *cloud;   // [N*3]            (query pass only)
    int *idx;       // [T]
    int *src;       // [T]   global point row b*N + idx
    float *rel;     // [T][4]
    float *z0;      // [T][C0]
    float *y1;      // [T][C1]  (backward A/B)
    float *z1;      // [T][C1]
    float *d2;      // [T][C2]  (backward A: dy2;  backward B reuses it for dy1 [T][C1])
    float *red;     // [kWarps][3][64] cross-warp combine
    __device__ Smem(float *base, int N, bool query, bool bwd) {
        float *q = base;
        cloud = q; q += query ? ((N * 3 + 3) & ~3) : 0;
        idx = reinterpret_cast<int *>(q); q += T;
        src = reinterpret_cast<int *>(q); q += T;
        rel = q; q += T * 4;
        z0 = q; q += T * C0;
        z1 = q; q += T * C1;
        y1 = q; q += bwd ? T * C1 : 0;
        d2 = q; q += bwd ? T * C2 : 0;
        red = q;
    }
    static size_t bytes(int N, bool query, bool bwd) {
        size_t f = (query ? ((N * 3 + 3) & ~3) : 0) + 2 * T + 4 * T + (size_t)T * C0 + (size_t)T * C1 + (bwd ? (size_t)T * C1 + (size_t)T * C2 : 0) +
                   kWarps * 3 * 64;
        return f * sizeof(float);
    }
};

// rows of the tile: ball query (or reload of idx), relative coordinates, point rows.  tile = GT consecutive centroids of one instance.
template <int GT, bool QUERY>
__device__ __forceinline__ void tile_rows(const SaLevelP &p, const SaScale &sc, int tile, int &cur_b, float *cloud, int *idx_s, int *src_s, float *rel_s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = sc.ns, T = GT * ns;
    const int bj0 = tile * GT;   // first flattened centroid index b*M + j
    const int b = bj0 / p.M;
    if (QUERY) {
        if (b != cur_b) {  // uniform over the CTA
            __syncthreads();
            const float *srcp = p.xyz + (size_t)b * p.N * 3;
            for (int i = threadIdx.x; i < p.N * 3; i += kThreads) cloud[i] = srcp[i];
            cur_b = b;
        }
        __syncthreads();
        const float r2 = __fmul_rn(sc.radius, sc.radius);  // ball_query_gpu.cu:27
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int g = warp; g < GT; g += kWarps) {
            const float *q = p.new_xyz + (size_t)(bj0 + g) * 3;
            const float cx = q[0], cy = q[1], cz = q[2];
            int32_t *out = sc.idx + (size_t)(bj0 + g) * ns;
            int *outs = idx_s + g * ns;
            int cnt = 0, first = 0;
            for (int base = 0; base < p.N && cnt < ns; base += 32) {
                const int k = base + lane;
                bool hit = false;
                if (k < p.N) {
                    const float d2 = sqdist_ref(__fsub_rn(cx, cloud[k * 3 + 0]), __fsub_rn(cy, cloud[k * 3 + 1]), __fsub_rn(cz, cloud[k * 3 + 2]));
                    hit = d2 < r2;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    if (cnt == 0) first = base + __ffs(mask) - 1;
                    const int pos = cnt + __popc(mask & lt_mask);
                    if (hit && pos < ns) { out[pos] = k; outs[pos] = k; }
                    cnt += __popc(mask);
                }
            }
            if (cnt > ns) cnt = ns;
            // tail: the reference pre-fills the row with the first hit (ball_query_gpu.cu:39-43); no hit => zeros
            for (int l = cnt + lane; l < ns; l += 32) { out[l] = first; outs[l] = first; }
        }
    } else {
        __syncthreads();  // the previous tile's readers are done with the tile buffers
        for (int t = threadIdx.x; t < T; t += kThreads) idx_s[t] = sc.idx[(size_t)bj0 * ns + t];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += kThreads) {
        const int g = t / ns;
        const int src = b * p.N + idx_s[t];
        src_s[t] = src;
        const float *x = p.xyz + (size_t)src * 3, *c = p.new_xyz + (size_t)(bj0 + g) * 3;
        rel_s[t * 4 + 0] = __fsub_rn(x[0], c[0]);
        rel_s[t * 4 + 1] = __fsub_rn(x[1], c[1]);
        rel_s[t * 4 + 2] = __fsub_rn(x[2], c[2]);
    }
    __syncthreads();
}

// y0 of one row for the lane's channel: same FMA chain as sa_gather_l0_kernel (elementwise.cu)
__device__ __forceinline__ float y0_row(const SaLevelP &p, int s_off, int src, const float *rel, const float (&wx)[3]) {
    const float uv = p.u ? __ldg(p.u + (size_t)src * p.ldu + s_off) : 0.f;
    return __fmaf_rn(wx[0], rel[0], __fmaf_rn(wx[1], rel[1], __fmaf_rn(wx[2], rel[2], uv)));
}

struct LaneBn {
    float m, s, g, b;
};
__device__ __forceinline__ LaneBn lane_bn(const SaBn &bn, int c, bool ok) {
    LaneBn r{0.f, 1.f, 1.f, 0.f};
    if (ok && bn.mean) { r.m = bn.mean[c]; r.s = bn.invstd[c]; r.g = bn.gamma[c]; r.b = bn.beta[c]; }
    return r;
}

// cross-warp combine of per-lane partial sums: warp w, quantity a, lane -> red[(w*3 + a)*64 + slot]; then fixed-order sum over
// the warps that own channel c and one partial row per CTA.
template <int NACC>
__device__ __forceinline__ void flush_lane_sums(float *red, const float (&acc)[NACC], int slot, bool ok, int C, int nslices, float *part, int cta, int G) {
    // slot = channel index this lane owns; warps with the same (warp % nslices) own the same channels
    const int warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NACC; ++a) red[(warp * 3 + a) * 64 + (threadIdx.x & 31)] = ok ? acc[a] : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * C; i += kThreads) {
        const int a = i / C, c = i - a * C;
        const int sl = c >> 5, ln = c & 31;
        float t = 0.f;
        for (int w = sl; w < kWarps; w += nslices) t += red[(w * 3 + a) * 64 + ln];
        part[((size_t)a * G + cta) * C + c] = t;
    }
    (void)slot;
}

// ------------------------------------------------------------------------------------------------ forward
// PASS 0: layer-0 statistics;  PASS 1: layer-1 statistics;  PASS 2: layer-2 statistics + selection.  QUERY: run the ball query.
template <int C0, int C1, int C2, int GT, int PASS, bool QUERY>
__global__ void __launch_bounds__(kThreads, 2) sa_fwd_kernel(const SaLevelP p) {
    extern __shared__ __align__(16) float sa_smem[];
    Smem<C0, C1, C2, GT> sm(sa_smem, p.N, QUERY, false);
    const TileCtx tc = tile_ctx<GT>(p);
    const SaScale &sc = p.sc[tc.s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = tc.ns;
    static_assert(C0 <= 32 && C1 <= 32 && C2 % 32 == 0 && C2 <= 64 && GT == kWarps, "fused set abstraction: narrow levels only");
    constexpr int NS2 = C2 / 32;  // channel slices of layer 2
    // lane-resident weights / BatchNorm parameters
    const bool ok0 = lane < C0, ok1 = lane < C1;
    float wx[3] = {0.f, 0.f, 0.f};
    if (ok0) { wx[0] = sc.w0[(size_t)lane * sc.ldw0]; wx[1] = sc.w0[(size_t)lane * sc.ldw0 + 1]; wx[2] = sc.w0[(size_t)lane * sc.ldw0 + 2]; }
    const LaneBn b0 = lane_bn(sc.bn0, lane, ok0 && PASS >= 1);
    float w1[C0];
#pragma unroll
    for (int k = 0; k < C0; ++k) w1[k] = (PASS >= 1 && ok1) ? sc.w1[(size_t)lane * C0 + k] : 0.f;
    const LaneBn b1 = lane_bn(sc.bn1, lane, ok1 && PASS >= 2);
    const int c2 = (warp % NS2) * 32 + lane;  // layer-2 channel of this lane
    float w2[C1];
#pragma unroll
    for (int k = 0; k < C1; ++k) w2[k] = (PASS >= 2) ? sc.w2[(size_t)c2 * C1 + k] : 0.f;
    const bool pick_max = (PASS >= 2 && sc.bn2.gamma) ? (sc.bn2.gamma[c2] >= 0.f) : true;
    float acc[2] = {0.f, 0.f};  // sum, sum of squares of the pass's layer for the lane's channel
    int cur_b = -1;
    for (int tile = tc.t_begin; tile < tc.t_end; ++tile) {
        tile_rows<GT, QUERY>(p, sc, tile, cur_b, sm.cloud, sm.idx, sm.src, sm.rel);
        const int bj0 = tile * GT;
        // ---- layer 0: warp = neighbourhood, lane = channel
        {
            const int r0 = warp * ns;
            if (ok0) {
                for (int l = 0; l < ns; ++l) {
                    const float y = y0_row(p, tc.s * C0 + lane, sm.src[r0 + l], sm.rel + (r0 + l) * 4, wx);
                    if (PASS == 0) { acc[0] += y; acc[1] += y * y; }
                    else sm.z0[(r0 + l) * C0 + lane] = fmaxf(bn_apply(y, b0.m, b0.s, b0.g, b0.b), 0.f);
                }
            }
        }
        if (PASS == 0) continue;
        __syncwarp();  // layer 1 of this warp reads only the rows this warp wrote
        {
            const int r0 = warp * ns;
            if (ok1) {
                for (int l = 0; l < ns; ++l) {
                    const float y = row_dot<C0>(sm.z0 + (r0 + l) * C0, w1);
                    if (PASS == 1) { acc[0] += y; acc[1] += y * y; }
                    else sm.z1[(r0 + l) * C1 + lane] = fmaxf(bn_apply(y, b1.m, b1.s, b1.g, b1.b), 0.f);
                }
            }
        }
        if (PASS == 1) continue;
        __syncthreads();  // with two channel slices a warp reads neighbourhoods written by other warps
        for (int g = warp / NS2; g < GT; g += kWarps / NS2) {
            const int r0 = g * ns;
            float best = 0.f;
            int bi = 0;
            for (int l = 0; l < ns; ++l) {
                const float y = row_dot<C1>(sm.z1 + (r0 + l) * C1, w2);
                acc[0] += y; acc[1] += y * y;
                const bool take = (l == 0) || (pick_max ? (y > best) : (y < best));  // strict: the first extreme wins (max_pool2d)
                if (take) { best = y; bi = l; }
            }
            sc.ysel[(size_t)(bj0 + g) * C2 + c2] = best;
            sc.asel[(size_t)(bj0 + g) * C2 + c2] = (uint8_t)bi;
        }
    }
    if (sc.part == nullptr) return;  // running statistics (eval): nothing to reduce
    constexpr int CS = PASS == 0 ? C0 : (PASS == 1 ? C1 : C2);
    constexpr int NSL = PASS == 2 ? NS2 : 1;
    flush_lane_sums<2>(sm.red, acc, 0, PASS == 0 ? ok0 : (PASS == 1 ? ok1 : true), CS, NSL, sc.part, tc.cta, tc.G);
    ticket_finish<2>(sc.fin, sc.part, tc.G, CS, tc.cta);
}

// out[bj][off + c] = relu(bn2(ysel[bj][c])) for both scales
template <int C2>
__global__ void __launch_bounds__(kThreads) sa_final_kernel(int BM, SaScale s0, SaScale s1, float *out, int ld_out) {
    const int total = BM * 2 * C2;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        const int c = i % (2 * C2), bj = i / (2 * C2);
        const SaScale &sc = c < C2 ? s0 : s1;
        const int cc = c < C2 ? c : c - C2;
        const float y = sc.ysel[(size_t)bj * C2 + cc];
        out[(size_t)bj * ld_out + c] = fmaxf(bn_apply(y, sc.bn2.mean[cc], sc.bn2.invstd[cc], sc.bn2.gamma[cc], sc.bn2.beta[cc]), 0.f);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// pre: BatchNorm-backward sums of layer 2.  The gradient of the max over the neighbours is non-zero on one row per (centroid,
// channel), so sum g and sum g*xhat run over [B*M, C2] values only: g = dz * [bn2(ysel) > 0], xhat = (ysel - mean) * invstd.
template <int C2>
__global__ void __launch_bounds__(kThreads, 2) sa_bwd_pre_kernel(const SaLevelP p) {
    __shared__ float red[kWarps * 3 * 64];
    const int s = ((int)blockIdx.x >= p.sc[1].cta0 && p.sc[1].ncta > 0) ? 1 : 0;
    const SaScale &sc = p.sc[s];
    const int cta = (int)blockIdx.x - sc.cta0, G = sc.ncta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NS2 = C2 / 32;
    const int c = (warp % NS2) * 32 + lane;
    const LaneBn b2 = lane_bn(sc.bn2, c, true);
    float acc[3] = {0.f, 0.f, 0.f};
    const int BM = p.B * p.M;
    for (int bj = cta * (kWarps / NS2) + warp / NS2; bj < BM; bj += G * (kWarps / NS2)) {
        const float y = sc.ysel[(size_t)bj * C2 + c];
        const float xh = __fmul_rn(__fsub_rn(y, b2.m), b2.s);
        const float u = __fmaf_rn(xh, b2.g, b2.b);
        const float g = u > 0.f ? sc.dz[(size_t)bj * sc.ld_dz + sc.off_dz + c] : 0.f;
        acc[0] += g;
        acc[1] += g * xh;
    }
    flush_lane_sums<3>(red, acc, 0, true, C2, NS2, sc.part, cta, G);
    ticket_finish<3>(sc.fin, sc.part, G, C2, cta);
}

// Weight-gradient accumulators of one warp (lane = input channel ci < CI, one register per output channel co < CO) are combined
// across the CTA's warps in a fixed order and written as this CTA's partial row; the scale's last CTA sums the rows (two-level
// ticket reduction) into dw[co*ld + ci].
template <int CO, int CI>
__device__ void flush_weight_grad(float *scratch /* >= kWarps*CO*CI floats of shared memory */, const float (&aw)[CO], bool ok, float *part, unsigned *tickets,
                                  int cta, int G, float *dw, int ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (ok) {
#pragma unroll
        for (int co = 0; co < CO; ++co) scratch[((size_t)warp * CO + co) * CI + lane] = aw[co];
    }
    __syncthreads();
    constexpr int CW = CO * CI;
    for (int i = threadIdx.x; i < CW; i += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += scratch[(size_t)w * CW + i];
        part[(size_t)cta * CW + i] = t;
    }
    float *part2 = part + (size_t)kMaxPartialRows * CW;
    if (!ticket_reduce<1>(part, cta, G, CW, tickets, part2)) return;
    for (int i = threadIdx.x; i < CW; i += kThreads) {
        const int co = i / CI, ci = i - co * CI;
        dw[(size_t)co * ld + ci] = (float)ticket_total(part2, G, CW, 0, i);
    }
}

// STAGE 0 ("A"): layer 2: dy2 = BN-backward of the max-routed gradient; g1 = (dy2 W2) * [z1 > 0] -> sc.g1, sums of layer 1, dW2.
// STAGE 1 ("B"): layer 1: dy1 from g1; g0 = (dy1 W1) * [z0 > 0] -> sc.g0, sums of layer 0, dW1.
// STAGE 2 ("C"): layer 0: dy0 from g0; dU[point] += dy0 (atomics, as group_points_grad), dWx.
// Weight rows / columns are (re)loaded from global memory (L1-resident, <= 8 KB) at the start of the phase that uses them, so that
// only the weight-gradient accumulators live in registers across a tile.
template <int K>
__device__ __forceinline__ void load_row(float (&w)[K], const float *src, int stride, bool ok) {
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = ok ? __ldg(src + (size_t)k * stride) : 0.f;
}
template <int C0, int C1, int C2, int GT, int STAGE>
__global__ void __launch_bounds__(kThreads, (C2 <= 32 ? 2 : 1)) sa_bwd_kernel(const SaLevelP p) {
    extern __shared__ __align__(16) float sa_smem[];
    Smem<C0, C1, C2, GT> sm(sa_smem, p.N, false, true);
    const TileCtx tc = tile_ctx<GT>(p);
    const SaScale &sc = p.sc[tc.s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ns = tc.ns, T = tc.T;
    const double invP = 1.0 / ((double)p.B * p.M * ns);
    static_assert(C0 <= C1 && C1 <= 32 && C2 % 32 == 0 && C2 <= 64 && (GT * 16) % kWarps == 0, "fused set abstraction: narrow levels only");
    constexpr int NS2 = C2 / 32;
    const bool ok0 = lane < C0, ok1 = lane < C1;
    float wx[3] = {0.f, 0.f, 0.f};
    if (ok0) { wx[0] = sc.w0[(size_t)lane * sc.ldw0]; wx[1] = sc.w0[(size_t)lane * sc.ldw0 + 1]; wx[2] = sc.w0[(size_t)lane * sc.ldw0 + 2]; }
    const LaneBn b0 = lane_bn(sc.bn0, lane, ok0);
    const LaneBn b1 = lane_bn(sc.bn1, lane, ok1);
    const int rows_w = T / kWarps;  // rows of the tile owned by this warp in the row-parallel phases
    const int rw0 = warp * rows_w;
    int cur_b = -1;

    if (STAGE == 2) {
        // ---------------- layer 0: BatchNorm backward, scatter to the points, dWx
        const float mg = ok0 ? (float)(sc.ws0[lane] * invP) : 0.f, mgx = ok0 ? (float)(sc.ws0[C0 + lane] * invP) : 0.f;
        float aw[3] = {0.f, 0.f, 0.f};
        for (int tile = tc.t_begin; tile < tc.t_end; ++tile) {
            tile_rows<GT, false>(p, sc, tile, cur_b, sm.cloud, sm.idx, sm.src, sm.rel);
            const size_t row0 = (size_t)tile * T;
            if (ok0) {
                for (int r = rw0; r < rw0 + rows_w; ++r) {
                    const float *rel = sm.rel + r * 4;
                    const int src = sm.src[r];
                    const float y = y0_row(p, tc.s * C0 + lane, src, rel, wx);
                    const float xh = __fmul_rn(__fsub_rn(y, b0.m), b0.s);
                    const float g = sc.g0[(row0 + r) * C0 + lane];
                    const float dy = b0.g * b0.s * (g - mg - xh * mgx);
                    aw[0] += dy * rel[0]; aw[1] += dy * rel[1]; aw[2] += dy * rel[2];
                    if (p.dU) atomicAdd(p.dU + (size_t)src * p.ldu + tc.s * C0 + lane, dy);
                }
            }
        }
        // dWx[c][d]: lane = c, 3 values -> dw[c*ld + d]
        float *scratch = sm.z0;
        __syncthreads();
        if (ok0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) scratch[(warp * 3 + d) * 32 + lane] = aw[d];
        }
        __syncthreads();
        constexpr int CW = 3 * C0;
        for (int i = threadIdx.x; i < CW; i += kThreads) {
            const int c = i / 3, d = i - c * 3;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += scratch[(w * 3 + d) * 32 + c];
            sc.part_w[(size_t)tc.cta * CW + i] = t;
        }
        float *part2 = sc.part_w + (size_t)kMaxPartialRows * CW;
        if (!ticket_reduce<1>(sc.part_w, tc.cta, tc.G, CW, sc.tickets_w, part2)) return;
        for (int i = threadIdx.x; i < CW; i += kThreads) {
            const int c = i / 3, d = i - c * 3;
            sc.dw[(size_t)c * sc.ld_dw + d] = (float)ticket_total(part2, tc.G, CW, 0, i);
        }
        return;
    }

    if (STAGE == 1) {
        // ---------------- layer 1: dy1 from the stored g1; dz0 = dy1 W1; g0; sums of layer 0; dW1
        const float mg1 = ok1 ? (float)(sc.ws1[lane] * invP) : 0.f, mgx1 = ok1 ? (float)(sc.ws1[C1 + lane] * invP) : 0.f;
        float acc[3] = {0.f, 0.f, 0.f};
        float aw[C1];  // dW1[c1][c0] for the lane's c0
#pragma unroll
        for (int k = 0; k < C1; ++k) aw[k] = 0.f;
        float *dy1 = sm.d2;  // [T][C1]
        for (int tile = tc.t_begin; tile < tc.t_end; ++tile) {
            tile_rows<GT, false>(p, sc, tile, cur_b, sm.cloud, sm.idx, sm.src, sm.rel);
            const size_t row0 = (size_t)tile * T;
            if (ok0) {
                for (int r = rw0; r < rw0 + rows_w; ++r) {
                    const float y = y0_row(p, tc.s * C0 + lane, sm.src[r], sm.rel + r * 4, wx);
                    sm.y1[r * C0 + lane] = y;  // y0 parked in the (otherwise unused) y1 buffer: [T][C0] <= [T][C1] floats... see host check C0 <= C1
                    sm.z0[r * C0 + lane] = fmaxf(bn_apply(y, b0.m, b0.s, b0.g, b0.b), 0.f);
                }
            }
            __syncwarp();
            {
                float w1[C0];  // row of W1 for the lane's output channel c1
                load_row<C0>(w1, sc.w1 + (size_t)lane * C0, 1, ok1);
                if (ok1) {
                    for (int r = rw0; r < rw0 + rows_w; ++r) {
                        const float y = row_dot<C0>(sm.z0 + r * C0, w1);
                        const float xh = __fmul_rn(__fsub_rn(y, b1.m), b1.s);
                        const float g = sc.g1[(row0 + r) * C1 + lane];
                        dy1[r * C1 + lane] = b1.g * b1.s * (g - mg1 - xh * mgx1);
                    }
                }
            }
            __syncwarp();
            {
                float w1c[C1];  // column of W1 for the lane's input channel c0: dz0[c0] = sum_c1 dy1[c1] * W1[c1][c0]
                load_row<C1>(w1c, sc.w1 + lane, C0, ok0);
                if (ok0) {
                    for (int r = rw0; r < rw0 + rows_w; ++r) {
                        const float dz0 = row_dot<C1>(dy1 + r * C1, w1c);
                        const float z = sm.z0[r * C0 + lane];
                        const float g = z > 0.f ? dz0 : 0.f;
                        sc.g0[(row0 + r) * C0 + lane] = g;
                        const float xh = __fmul_rn(__fsub_rn(sm.y1[r * C0 + lane], b0.m), b0.s);
                        acc[0] += g; acc[1] += g * xh;
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
        flush_lane_sums<3>(sm.red, acc, 0, ok0, C0, 1, sc.part, tc.cta, tc.G);
        ticket_finish<3>(sc.fin, sc.part, tc.G, C0, tc.cta);
        flush_weight_grad<C1, C0>(sm.z0, aw, ok0, sc.part_w, sc.tickets_w, tc.cta, tc.G, sc.dw, sc.ld_dw);
        return;
    }

    // ---------------- STAGE 0: layer 2
    {
        float acc[3] = {0.f, 0.f, 0.f};
        float aw[C2];  // dW2[c2][c1] for the lane's c1
#pragma unroll
        for (int k = 0; k < C2; ++k) aw[k] = 0.f;
        // phase 3 (dy2) splits the tile by (channel slice, row block): warp -> slice q3, rows [r3, r3 + rows3)
        const int q3 = warp % NS2, rows3 = T / (kWarps / NS2), r3 = (warp / NS2) * rows3;
        const int c3 = q3 * 32 + lane;
        const LaneBn b2 = lane_bn(sc.bn2, c3, true);
        const float mg2 = (float)(sc.ws2[c3] * invP), mgx2 = (float)(sc.ws2[C2 + c3] * invP);
        for (int tile = tc.t_begin; tile < tc.t_end; ++tile) {
            tile_rows<GT, false>(p, sc, tile, cur_b, sm.cloud, sm.idx, sm.src, sm.rel);
            const size_t row0 = (size_t)tile * T;
            const int bj0 = tile * GT;
            if (ok0) {
                for (int r = rw0; r < rw0 + rows_w; ++r) {
                    const float y = y0_row(p, tc.s * C0 + lane, sm.src[r], sm.rel + r * 4, wx);
                    sm.z0[r * C0 + lane] = fmaxf(bn_apply(y, b0.m, b0.s, b0.g, b0.b), 0.f);
                }
            }
            __syncwarp();
            {
                float w1[C0];
                load_row<C0>(w1, sc.w1 + (size_t)lane * C0, 1, ok1);
                if (ok1) {
                    for (int r = rw0; r < rw0 + rows_w; ++r) {
                        const float y = row_dot<C0>(sm.z0 + r * C0, w1);
                        sm.y1[r * C1 + lane] = y;
                        sm.z1[r * C1 + lane] = fmaxf(bn_apply(y, b1.m, b1.s, b1.g, b1.b), 0.f);
                    }
                }
            }
            __syncthreads();
            {
                float w2[C1];  // row of W2 for the lane's layer-2 channel c3
                load_row<C1>(w2, sc.w2 + (size_t)c3 * C1, 1, true);
                for (int r = r3; r < r3 + rows3; ++r) {
                    const int g = r / ns, l = r - g * ns;
                    const float y = row_dot<C1>(sm.z1 + r * C1, w2);
                    const float xh = __fmul_rn(__fsub_rn(y, b2.m), b2.s);
                    float gg = 0.f;
                    if (sc.asel[(size_t)(bj0 + g) * C2 + c3] == l) {
                        const float u = __fmaf_rn(xh, b2.g, b2.b);
                        if (u > 0.f) gg = sc.dz[(size_t)(bj0 + g) * sc.ld_dz + sc.off_dz + c3];
                    }
                    sm.d2[r * C2 + c3] = b2.g * b2.s * (gg - mg2 - xh * mgx2);
                }
            }
            __syncthreads();
            {
                float w2c[C2];  // column of W2 for the lane's layer-1 channel c1: dz1[c1] = sum_c2 dy2[c2] * W2[c2][c1]
                load_row<C2>(w2c, sc.w2 + lane, C1, ok1);
                if (ok1) {
                    for (int r = rw0; r < rw0 + rows_w; ++r) {
                        const float dz1 = row_dot<C2>(sm.d2 + r * C2, w2c);
                        const float z = sm.z1[r * C1 + lane];
                        const float g = z > 0.f ? dz1 : 0.f;
                        sc.g1[(row0 + r) * C1 + lane] = g;
                        const float xh = __fmul_rn(__fsub_rn(sm.y1[r * C1 + lane], b1.m), b1.s);
                        acc[0] += g; acc[1] += g * xh;
#pragma unroll
                        for (int k = 0; k < C2; k += 4) {  // dW2[:, c1] += dy2[r, :] * z1[r, c1]
                            const float4 v = *reinterpret_cast<const float4 *>(sm.d2 + r * C2 + k);
                            aw[k] = __fmaf_rn(v.x, z, aw[k]); aw[k + 1] = __fmaf_rn(v.y, z, aw[k + 1]);
                            aw[k + 2] = __fmaf_rn(v.z, z, aw[k + 2]); aw[k + 3] = __fmaf_rn(v.w, z, aw[k + 3]);
                        }
                    }
                }
            }
        }
        flush_lane_sums<3>(sm.red, acc, 0, ok1, C1, 1, sc.part, tc.cta, tc.G);
        ticket_finish<3>(sc.fin, sc.part, tc.G, C1, tc.cta);
        flush_weight_grad<C2, C1>(sm.z0, aw, ok1, sc.part_w, sc.tickets_w, tc.cta, tc.G, sc.dw, sc.ld_dw);
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
                                                                const float *w0b, int ldw0, float *dF, float *part_w, unsigned *tickets_w, float *dwa,
                                                                float *dwb, int ld_dw) {
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
    float *part2 = part_w + (size_t)kMaxPartialRows * CW;
    if (!ticket_reduce<1>(part_w, (int)blockIdx.x, (int)gridDim.x, CW, tickets_w, part2)) return;
    for (int i = threadIdx.x; i < CW; i += kThreads) {
        const int j = i / K, k = i - j * K;
        float *dst = j < C0 ? dwa + (size_t)j * ld_dw + 3 + k : dwb + (size_t)(j - C0) * ld_dw + 3 + k;
        *dst = (float)ticket_total(part2, (int)gridDim.x, CW, 0, i);
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
static bool fill_level(int B, int N, int M, const float *xyz, const float *new_xyz, const float *u, int ldu, float *dU, SaLevelP &p) {
    if (B <= 0 || N <= 0 || M <= 0 || (M % kWarps) != 0 || !xyz || !new_xyz) return false;
    if ((long long)B * M * kMaxNs > 0x7fffffffLL || (size_t)N * 12 > 96 * 1024) return false;
    p = SaLevelP{};
    p.B = B; p.N = N; p.M = M; p.xyz = xyz; p.new_xyz = new_xyz; p.u = u; p.ldu = ldu; p.dU = dU;
    return true;
}
// CTAs of one launch shared between the two scales in proportion to their rows (16 : 32 neighbours)
static void split_ctas(SaLevelP &p, int total) {
    const int n_tiles = p.B * p.M / kWarps;
    const int w0 = p.sc[0].ns, w1 = p.sc[1].ns;
    int g0 = total * w0 / (w0 + w1);
    if (g0 < 1) g0 = 1;
    int g1 = total - g0;
    if (g0 > n_tiles) g0 = n_tiles;
    if (g1 > n_tiles) g1 = n_tiles;
    if (g0 > kMaxPartialRows) g0 = kMaxPartialRows;
    if (g1 > kMaxPartialRows) g1 = kMaxPartialRows;
    p.sc[0].cta0 = 0; p.sc[0].ncta = g0;
    p.sc[1].cta0 = g0; p.sc[1].ncta = g1;
}

template <int C0, int C1, int C2>
static int launch_fwd(SaLevelP &p, int pass, int query, cudaStream_t st) {
    constexpr int GT = kWarps;
    const size_t smem = Smem<C0, C1, C2, GT>::bytes(p.N, query != 0, false);
    const int per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
    split_ctas(p, kNumSMs * per_sm);
    const int grid = p.sc[0].ncta + p.sc[1].ncta;
#define SA_FWD(PASS, Q)                                                                                                             \
    do {                                                                                                                            \
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_fwd_kernel<C0, C1, C2, GT, PASS, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_fwd_kernel<C0, C1, C2, GT, PASS, Q><<<grid, kThreads, smem, st>>>(p);                                                    \
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
    constexpr int GT = kWarps;
    if (stage < 0) {
        split_ctas(p, kNumSMs * 2);
        sa_bwd_pre_kernel<C2><<<p.sc[0].ncta + p.sc[1].ncta, kThreads, 0, st>>>(p);
        ISTNET_LAUNCH_CHECK();
        return ISTNET_OK;
    }
    const size_t smem = Smem<C0, C1, C2, GT>::bytes(p.N, false, true);
    const int per_sm = (C2 <= 32 && smem * 2 <= 220 * 1024) ? 2 : 1;
    split_ctas(p, kNumSMs * per_sm);
    const int grid = p.sc[0].ncta + p.sc[1].ncta;
#define SA_BWD(STAGE)                                                                                                              \
    do {                                                                                                                           \
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_bwd_kernel<C0, C1, C2, GT, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sa_bwd_kernel<C0, C1, C2, GT, STAGE><<<grid, kThreads, smem, st>>>(p);                                                     \
    } while (0)
    if (stage == 0) SA_BWD(0);
    else if (stage == 1) SA_BWD(1);
    else if (stage == 2) SA_BWD(2);
    else return ISTNET_ERR_BAD_ARG;
#undef SA_BWD
    ISTNET_LAUNCH_CHECK();
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
    if (u && ldu < 2 * C0) return ISTNET_ERR_BAD_ARG;
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
        if (stage == 0 && (!d.dz || !d.asel || !d.ws2 || !d.g1 || !d.part || d.fin.kind != ISTNET_FIN_BN_BWD || !d.part_w || !d.tickets_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
        if (stage == 1 && (!d.ws1 || !d.g1 || !d.g0 || !d.part || d.fin.kind != ISTNET_FIN_BN_BWD || !d.part_w || !d.tickets_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
        if (stage == 2 && (!d.ws0 || !d.g0 || !d.part_w || !d.tickets_w || !d.dw)) return ISTNET_ERR_BAD_ARG;
    }
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
                               float *part_w, unsigned *tickets_w, float *dw0a, float *dw0b, int ld_dw, void *stream) {
    if (R <= 0 || !F || !dU || !w0a || !w0b || !part_w || !tickets_w || !dw0a || !dw0b || ldw0 < 3 + K || ld_dw < 3 + K) return ISTNET_ERR_BAD_ARG;
    if (!(K == 64 && C0 == 32)) return ISTNET_ERR_UNSUPPORTED;
    int grid = (R + kWarps * 4 - 1) / (kWarps * 4);
    if (grid > kNumSMs) grid = kNumSMs;
    const size_t smem = (size_t)(kWarps * 4 * 64 + kWarps * 64 * 32) * sizeof(float);
    ISTNET_CUDA_TRY(cudaFuncSetAttribute(sa_u_bwd_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_u_bwd_kernel<64, 32><<<grid, kThreads, smem, (cudaStream_t)stream>>>(R, F, dU, w0a, w0b, ldw0, dF, part_w, tickets_w, dw0a, dw0b, ld_dw);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
