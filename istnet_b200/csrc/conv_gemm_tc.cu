// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), FP32-accurate.
//
//   out[p, n] = sum_{tap} sum_{c} act[p + shift(tap), c] * wgt[tap][n][c]      (stride 1, "same" padding)
//
// covers every stride-1 convolution of the image branch (modules.py / resnet.py), its data
// gradient (same kernel, flipped+transposed weights), and all 1x1 convolutions / per-point MLP layers (taps = 1).
//
// Precision: the 1e-4 relative target rules out single-pass TF32/BF16.  Each FP32 operand x is carried as `nsplit`
// bf16 planes (p0 = bf16(x), p1 = bf16(x - p0), p2 = ...; residual <= 2^(-9 nsplit) |x|) and a product is evaluated as
// the sum of all plane products a_i*b_j with i + j < nsplit, smallest first, FP32-accumulated in TMEM:
//   nsplit = 2: 3 tcgen05.mma per k-step, ~3e-6 RMS per product;  nsplit = 3: 6 per k-step, ~1e-8 (below FP32 noise).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0 : TMA producer  — per (tap, 64-channel block): nsplit activation boxes (5-D map over C,W,H,B,plane with
//            box 64 x bw x bh x bb x 1 = 128 pixels, coordinates shifted by the tap, out-of-bounds = zero padding)
//            and nsplit weight boxes (64 x BN) into a SWIZZLE_128B ring, completion on mbarriers
//   warp 1 : MMA issuer    — one elected thread, 4 k-steps x 3|6 products per stage, tcgen05.commit frees the stage
//   warps 2-5 : epilogue   — tcgen05.ld 32x32b, + bias, ReLU, FP32 and/or bf16-plane stores (NHWC rows)
#include <stdlib.h>

#include "tc_common.cuh"
#include "ticket.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;      // 2 + 4 epilogue warps: tiles narrow enough for two CTAs per SM
constexpr int kThreadsWide = 320;  // 2 + 8 epilogue warps: one CTA per SM

struct ConvGemmParams {
    int B, H, W;              // output (= input) extent
    int box_w, box_h, box_b;  // pixel tile, box_w*box_h*box_b == 128
    int tiles_w, tiles_h, tiles_b;
    int kh, kw;               // taps (1x1 or 3x3 ...), padding = k/2
    int cin_blocks;           // ceil(Cin / block_k)
    int block_k;              // K elements per pipeline stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)
    int Cout, BN;             // logical output channels, tile width (multiple of 16, <= 256)
    int n_tiles_n;            // ceil(Cout / BN)
    int acc_stride;           // TMEM columns between the two accumulators (BN rounded up to 32)
    int stages;
    const float *bias;        // [Cout] or null; with bias_group > 0: [ceil(P / bias_group)][Cout], row = output pixel / bias_group
    int bias_group;
    int relu;
    float *out_f32;           // [B,H,W,out_cs] or null
    int out_cs;
    __nv_bfloat16 *out_pl;    // operand planes [nsplit_out][B,H,W,split_cs] or null
    long long out_pl_stride;
    int split_cs, nsplit_out;
    int nsplit;               // planes of the input operands
    float *stat_part;         // optional [gridDim.x][2][Cout]: per-CTA column sums / sums of squares of the output (BN statistics)
    const __nv_bfloat16 *mask_hi;  // optional [B,H,W,mask_cs]: output element kept only where mask > 0 (ReLU backward of the layer below)
    int mask_cs;
    const float *stat_y;      // optional [B,H,W,stat_y_cs]: second statistic = sum(out * stat_y) instead of sum(out^2) (BN backward)
    int stat_y_cs;
    uint32_t tmem_cols;
    FinP fin;                 // optional: the last CTA finishes the statistics reduction (ticket.cuh)
    int cluster;              // 1, or 2: CTA pairs work on two row tiles of the same column tile and share the weight tile (TMA multicast)
};

// CL = CTAs per cluster (compile-time: the single-CTA instantiation is exactly the round-1 kernel; a run-time switch on the hot
// single-thread producer / MMA-issue loops cost the 3x3 convolutions 12 %, profiles/r2_tensor_core_kernel_table.txt vs _cluster2)
template <int CL>
__global__ void __launch_bounds__(kThreadsWide, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const ConvGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kBlockK = p.block_k;
    const int kATileBytes = kTileM * kBlockK * 2;
    const int b_tile_bytes = p.BN * kBlockK * 2;
    const int stage_bytes = p.nsplit * (kATileBytes + b_tile_bytes);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + p.stages;
    uint64_t *tmem_full_bar = empty_bar + p.stages;   // [2] accumulator ready for the epilogue
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;     // [2] accumulator drained by the epilogue
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    float *s_stat = reinterpret_cast<float *>(smem + (size_t)p.stages * stage_bytes + 256);  // [4 warps][2][Cout] when stat_part != null

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = p.kh * p.kw * p.cin_blocks;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    // Work items.  cluster == 2: the CTA pair (cluster) takes two consecutive row tiles of ONE column tile per item; each CTA loads
    // half of the weight tile and multicasts it into both CTAs' rings, so the weight operand crosses L2 -> SM once per pair.  A stage
    // may be overwritten only when BOTH consumers have retired it: the MMA commit arrives on the empty barrier of both CTAs (count 2).
    constexpr int cl = CL;
    const uint32_t crank = cl > 1 ? tc::cluster_ctarank() : 0u;
    const int m_groups = (m_tiles + cl - 1) / cl;
    const int n_work = m_groups * p.n_tiles_n;
    const int w_first = (int)blockIdx.x / cl, w_step = (int)gridDim.x / cl;
    // PERSISTENT: this CTA processes tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The shared-memory ring keeps running
    // across tiles and the accumulator is double-buffered in tensor memory, so the epilogue of tile i overlaps the TMA /
    // MMA main loop of tile i+1 and the per-CTA set-up (TMEM allocation, barrier init, descriptor prefetch) is paid once.

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_a); tc::prefetch_tmap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], (uint32_t)cl); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tmem_full_bar[s], 1); tc::mbar_init(&tmem_empty_bar[s], (blockDim.x >> 5) - 2); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_holder, p.tmem_cols);
    if (p.stat_part)
        for (int i = threadIdx.x; i < 8 * p.Cout; i += blockDim.x) s_stat[i] = 0.f;
    tc::tc_fence_before();
    __syncthreads();
    if (cl > 1) tc::cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int pad_h = p.kh / 2, pad_w = p.kw / 2;
            int it = 0;
            for (int wi = w_first; wi < n_work; wi += w_step) {
                int mt = (wi % m_groups) * cl + (int)crank;  // may lie past the last row tile (odd count): every box is then out of bounds = zeros
                const int n0 = (wi / m_groups) * p.BN;
                const int tw = mt % p.tiles_w; mt /= p.tiles_w;
                const int th = mt % p.tiles_h; mt /= p.tiles_h;
                const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = mt * p.box_b;
                // channel block OUTER, filter tap INNER: the nine shifted boxes of one channel slab are fetched back to
                // back, so taps 1..8 hit in L2 (tap-outer order re-streamed the activation tensor from HBM per tap:
                // 6.9 GB instead of 0.5 GB for up_1 at B=32, profiles/r1_conv_up1_taporder_before.txt)
                for (int kb = 0; kb < p.cin_blocks; ++kb) {
                    for (int tap = 0; tap < p.kh * p.kw; ++tap, ++it) {
                        const int r = tap / p.kw, s = tap % p.kw;
                        const int st = it % p.stages;
                        const uint32_t ph = (it / p.stages) & 1;
                        tc::mbar_wait(&empty_bar[st], ph ^ 1);
                        uint8_t *sa = smem + (size_t)st * stage_bytes;
                        tc::mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
                        uint8_t *sb = sa + p.nsplit * kATileBytes;
                        for (int pl = 0; pl < p.nsplit; ++pl) {
                            tc::tma_load_5d(sa + pl * kATileBytes, &tm_a, &full_bar[st], kb * kBlockK, w0 + s - pad_w, h0 + r - pad_h, b0, pl);
                            if (cl > 1)  // this CTA's half of the weight rows, delivered to both CTAs
                                tc::tma_load_4d_multicast(sb + pl * b_tile_bytes + crank * (b_tile_bytes / 2), &tm_b, &full_bar[st], kb * kBlockK,
                                                          n0 + (int)crank * (p.BN / 2), tap, pl, (uint16_t)0x3);
                            else
                                tc::tma_load_4d(sb + pl * b_tile_bytes, &tm_b, &full_bar[st], kb * kBlockK, n0, tap, pl);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 0, 0);
            const uint32_t ltype = kBlockK == 64 ? 2u : 4u;   // SWIZZLE_128B | SWIZZLE_64B
            const uint32_t sbo = 8u * kBlockK * 2u;           // 8 rows of one swizzle atom
            int it = 0, lt = 0;
            for (int wi = w_first; wi < n_work; wi += w_step, ++lt) {
                const int buf = lt & 1;
                const uint32_t use = (uint32_t)(lt >> 1);
                tc::mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
                tc::tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * p.acc_stride);
                for (int kk = 0; kk < num_k; ++kk, ++it) {
                    const int st = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    tc::mbar_wait(&full_bar[st], ph);
                    tc::tc_fence_after();
                    const uint32_t a0 = tc::smem_u32(smem + (size_t)st * stage_bytes);
                    const uint32_t b0 = a0 + p.nsplit * kATileBytes;
                    for (int j = 0; j < kBlockK / 16; ++j) {
                        uint32_t acc = (kk | j) != 0;
                        // all plane products a_i * b_j with i + j < nsplit, smallest magnitude first
                        for (int sum = p.nsplit - 1; sum >= 0; --sum) {
                            for (int ia = sum; ia >= 0; --ia) {
                                const int ib = sum - ia;
                                const uint64_t da = tc::make_desc_swz(a0 + ia * kATileBytes + j * 32, 16, sbo, ltype);
                                const uint64_t db = tc::make_desc_swz(b0 + ib * b_tile_bytes + j * 32, 16, sbo, ltype);
                                tc::umma_bf16(tacc, da, db, idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                    if (cl > 1) tc::umma_commit_multicast(&empty_bar[st], (uint16_t)0x3);  // both producers wait for both consumers
                    else tc::umma_commit(&empty_bar[st]);  // frees the smem stage once these MMAs retire
                }
                tc::umma_commit(&tmem_full_bar[buf]);
            }
        }
    } else {
        // ===================== epilogue (warps 2 .. 2+n_epi-1) =====================
        // A warp may only read the TMEM lane quarter (warp % 4).  With 8 epilogue warps (one CTA per SM) two warps share a
        // quarter and take alternate 32-column chunks: the epilogue is a dependent ALU chain (bias, ReLU, operand split),
        // so its speed is set by how many warps each scheduler can interleave (measured with ONE warp per scheduler:
        // 92k cycles per 128x256 tile against a 34k-cycle main loop at K = 512, profiles/r1_rowsgemm_epilogue_before.txt).
        const int groups = ((int)(blockDim.x >> 5) - 2) >> 2;  // warps per lane quarter: 1 or 2
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int pb = row / (p.box_w * p.box_h);
        const int ph_ = (row / p.box_w) % p.box_h;
        const int pw = row % p.box_w;
        const bool pl_vec8 = p.out_pl && ((p.out_pl_stride & 7) == 0) && ((p.split_cs & 7) == 0) &&
                             ((reinterpret_cast<uintptr_t>(p.out_pl) & 15) == 0);
        const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        int lt = 0;
        for (int wi = w_first; wi < n_work; wi += w_step, ++lt) {
            int mt = (wi % m_groups) * cl + (int)crank;
            const int n0 = (wi / m_groups) * p.BN;
            const int tw = mt % p.tiles_w; mt /= p.tiles_w;
            const int th = mt % p.tiles_h; mt /= p.tiles_h;
            const int b = mt * p.box_b + pb, h = th * p.box_h + ph_, w = tw * p.box_w + pw;
            const bool row_ok = (b < p.B) && (h < p.H) && (w < p.W);
            const size_t pix = ((size_t)b * p.H + h) * p.W + w;
            const int buf = lt & 1;
            tc::mbar_wait(&tmem_full_bar[buf], (uint32_t)(lt >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            for (int c0 = grp * 32; c0 < p.BN; c0 += 32 * groups) {
                const int n = n0 + c0;
                if (n >= p.Cout) break;  // warp-uniform
                uint32_t v[32];
                tc::tmem_ld_32x32(tacc + (uint32_t)c0, v);
                tc::tmem_ld_wait();
                const int valid = min(32, p.Cout - n);
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                if (p.bias != nullptr) {
                    // per-group bias (the estimators' global feature, ist_net.py:172-173: W [f | mean(f)] = W_a f + W_b mean(f), the second
                    // term is constant over the 1024 rows of an instance): row pix / bias_group of a [groups][Cout] table
                    const float *bias = p.bias + ((p.bias_group > 0 && row_ok) ? (pix / (size_t)p.bias_group) * (size_t)p.Cout : 0);
                    if (valid == 32 && bias_vec && (p.bias_group == 0 || (p.Cout & 3) == 0)) {  // 16-byte aligned loads
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + n + i));
                            f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < valid) f[i] += __ldg(bias + n + i);
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                if (p.mask_hi != nullptr && row_ok) {
                    // data gradient of a layer whose input is the ReLU output z of the layer below: g = dx * [z > 0]; the mask is
                    // plane 0 of z as saved by the forward pass.  The statistics path below then yields sum(g) = that layer's
                    // bias gradient and the plane stores its dy operand: no separate reduce / apply passes over dx.
                    const __nv_bfloat16 *mrow = p.mask_hi + pix * p.mask_cs + n;
                    if (valid == 32 && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            const uint4 m = __ldg(reinterpret_cast<const uint4 *>(mrow + i));
                            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {  // bf16 > 0  <=>  sign clear and magnitude bits non-zero
                                const uint32_t lo = mw[j] & 0xffffu, hi = mw[j] >> 16;
                                if (!(lo != 0u && lo < 0x8000u)) f[i + 2 * j] = 0.f;
                                if (!(hi != 0u && hi < 0x8000u)) f[i + 2 * j + 1] = 0.f;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < valid && !(__bfloat162float(mrow[i]) > 0.f)) f[i] = 0.f;
                    }
                }
                if (p.stat_part) {
                    // BatchNorm statistics of this 32-row x 32-column block: butterfly transpose-reduce over the warp
                    // (31 shuffles per quantity), lane l ends with the column (c0 + l) totals and adds them to this lane
                    // QUARTER's shared-memory slot (warps of one quarter own disjoint columns; no atomics: fixed summation
                    // order, bitwise reproducible); the CTA flushes the sum of its four slots once, after its last tile
                    float a[32], q2[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { a[i] = row_ok ? f[i] : 0.f; q2[i] = a[i] * a[i]; }
                    if (p.stat_y != nullptr) {
                        // BatchNorm backward of the layer below (whose pre-BN output is stat_y): with the ReLU mask applied above,
                        // the two statistics are sum(g) and sum(g*y); sum(g*xhat) = invstd * (sum(g*y) - mean * sum(g)) follows
                        // in the finalize kernel, so the separate reduce pass over (dz, y, mask) is not needed
                        const float *yrow = p.stat_y + pix * p.stat_y_cs + n;
                        if (row_ok && valid == 32 && ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0)) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 yv = __ldg(reinterpret_cast<const float4 *>(yrow + i));
                                q2[i] = a[i] * yv.x; q2[i + 1] = a[i + 1] * yv.y; q2[i + 2] = a[i + 2] * yv.z; q2[i + 3] = a[i + 3] * yv.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) q2[i] = (row_ok && i < valid) ? a[i] * __ldg(yrow + i) : 0.f;
                        }
                    }
#pragma unroll
                    for (int half = 16; half >= 1; half >>= 1) {
                        const bool up = (lane & half) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            // keep the half of the columns selected by this lane's bit, send the other half
                            float keep_a = up ? a[i + half] : a[i], send_a = up ? a[i] : a[i + half];
                            float keep_q = up ? q2[i + half] : q2[i], send_q = up ? q2[i] : q2[i + half];
                            a[i] = keep_a + __shfl_xor_sync(0xffffffffu, send_a, half);
                            q2[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, half);
                        }
                    }
                    if (n + lane < p.Cout) {
                        s_stat[(q * 2 + 0) * p.Cout + n + lane] += a[0];
                        s_stat[(q * 2 + 1) * p.Cout + n + lane] += q2[0];
                    }
                }
                if (!row_ok) continue;
                if (p.out_f32) {
                    float *o = p.out_f32 + pix * p.out_cs + n;
                    if (valid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)  // predicated, fully unrolled: a dynamic index would put f[] in local memory
                            if (i < valid) o[i] = f[i];
                    }
                }
                if (p.out_pl) {
                    __nv_bfloat16 *o = p.out_pl + pix * p.split_cs + n;
                    if (valid == 32 && pl_vec8) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            const float g8[8] = {f[i], f[i + 1], f[i + 2], f[i + 3], f[i + 4], f[i + 5], f[i + 6], f[i + 7]};
                            store_planes8(o + i, p.out_pl_stride, p.nsplit_out, g8);
                        }
                    } else if (valid == 32 && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) store_planes4(o + i, p.out_pl_stride, p.nsplit_out, make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < valid) store_planes1(o + i, p.out_pl_stride, p.nsplit_out, f[i]);
                    }
                }
            }
            // this warp is done reading the accumulator: hand it back to the MMA issuer (one arrival per epilogue warp)
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
        }
    }
    __syncthreads();
    if (p.stat_part)
        for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) {
            // layout expected by the finalize kernel: part[(a*G + g)*C + c]
            const int a = i / p.Cout, c = i % p.Cout;
            const float tot = ((s_stat[(0 * 2 + a) * p.Cout + c] + s_stat[(1 * 2 + a) * p.Cout + c]) + s_stat[(2 * 2 + a) * p.Cout + c]) +
                              s_stat[(3 * 2 + a) * p.Cout + c];
            p.stat_part[((size_t)a * gridDim.x + blockIdx.x) * p.Cout + c] = tot;
        }
    if (p.stat_part && p.fin.kind != 0) {
        // the last CTA of the grid turns the partials into the BatchNorm statistics (or the bias gradient) — no finalize launch
        if (p.fin.kind == 2) ticket_finish<1>(p.fin, p.stat_part, (int)gridDim.x, p.Cout);
        else ticket_finish<2>(p.fin, p.stat_part, (int)gridDim.x, p.Cout);
    }
    if (cl > 1) tc::cluster_sync_all();  // no CTA leaves while its peer may still arrive on its barriers
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

}  // namespace

// ------------------------------------------------------------------ host helpers
PFN_tmapEncodeTiled istnet_get_tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
    }
    return fn;
}

int istnet_make_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                          const uint32_t *box, int swizzle_bytes) {
    PFN_tmapEncodeTiled enc = istnet_get_tmap_encoder();
    if (!enc) return ISTNET_ERR_UNSUPPORTED;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ISTNET_OK : ISTNET_ERR_BAD_ARG;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}
// Tile shape: N = 256 keeps the MMA off the shared-memory-bandwidth limit (an M=128 x N=128 x K=16 MMA reads 8 KB in 64
// cycles = the full 128 B/clk); with 3 operand planes a 64-wide K block would leave a single 144 KB stage, so K = 32
// (SWIZZLE_64B rows) is used there: 72 KB stages, 3 in flight.
static int pick_block_k(int cout, int nsplit) {
    static const int force = env_int("ISTNET_BK", 0);  // tuning knobs are read once per process, not per launch
    if (force == 32 || force == 64) return force;
    return nsplit >= 3 ? 32 : 64;
}
static int pick_bn(int cout, int nsplit, int block_k) {
    const int cap = (nsplit >= 3 && block_k == 64) ? 128 : 256;  // keep >= 2 pipeline stages in 227 KB of shared memory
    if (cout > cap) {  // equal-width column tiles: Cout = 384 -> 2 x 192 instead of 256 + a half-empty 256 (25 % of the MMAs on zero padding)
        const int n = (cout + cap - 1) / cap;
        const int bn = ((cout + n - 1) / n + 31) / 32 * 32;
        return bn < cap ? bn : cap;
    }
    if (cout == cap) return cap;
    int bn = (cout + 15) / 16 * 16;
    return bn < 16 ? 16 : bn;
}

extern "C" int istnet_conv_gemm(const void *act_planes, long long act_plane_stride, int B, int H, int W, int Cin, int act_cs,
                                const void *wgt_planes, long long wgt_plane_stride, int Cout, int wgt_cs, int kh, int kw, int nsplit,
                                const float *bias, int relu, float *out_f32, int out_cs, void *out_planes, long long out_plane_stride,
                                int nsplit_out, int split_cs, int box_w, int box_h, float *stat_part, int *grid_out, const void *mask_hi,
                                int mask_cs, const float *stat_y, int stat_y_cs, const istnet_fin *fin, int bias_group, void *stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || bias_group < 0 || (bias_group > 0 && !bias)) return ISTNET_ERR_BAD_ARG;
    if (nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if ((act_cs & 7) || (wgt_cs & 7) || act_cs < Cin || wgt_cs < Cin) return ISTNET_ERR_BAD_ARG;
    if (box_w <= 0 || box_h <= 0 || (kTileM % (box_w * box_h)) != 0 || box_w > 256 || box_h > 256) return ISTNET_ERR_BAD_ARG;
    if ((kh & 1) == 0 || (kw & 1) == 0) return ISTNET_ERR_UNSUPPORTED;
    if (out_planes && ((split_cs & 3) || nsplit_out < 1 || nsplit_out > kMaxPlanes)) return ISTNET_ERR_BAD_ARG;
    ConvGemmParams p{};
    p.B = B; p.H = H; p.W = W;
    p.box_w = box_w; p.box_h = box_h; p.box_b = kTileM / (box_w * box_h);
    p.tiles_w = ceil_div(W, box_w); p.tiles_h = ceil_div(H, box_h); p.tiles_b = ceil_div(B, p.box_b);
    p.kh = kh; p.kw = kw;
    const int kBlockK = pick_block_k(Cout, nsplit);
    const int kATileBytes = kTileM * kBlockK * 2;
    p.block_k = kBlockK;
    p.cin_blocks = ceil_div(Cin, kBlockK);
    p.Cout = Cout; p.BN = pick_bn(Cout, nsplit, kBlockK);
    p.nsplit = nsplit;
    p.bias = bias; p.relu = relu; p.bias_group = bias_group;
    p.stat_part = stat_part;
    p.mask_hi = (const __nv_bfloat16 *)mask_hi; p.mask_cs = mask_cs;
    p.stat_y = stat_y; p.stat_y_cs = stat_y_cs;
    if (stat_y && (!stat_part || stat_y_cs < Cout)) return ISTNET_ERR_BAD_ARG;
    if (mask_hi && mask_cs < Cout) return ISTNET_ERR_BAD_ARG;
    if (fin && fin->kind != ISTNET_FIN_NONE && (!stat_part || stat_y || fin->kind == ISTNET_FIN_BN_BWD)) return ISTNET_ERR_BAD_ARG;
    FinP fin_p;
    if (!make_fin(fin, stat_part, 2, Cout, fin_p)) return ISTNET_ERR_BAD_ARG;
    p.fin = fin_in_kernel(Cout) ? fin_p : FinP{};  // wide reductions: channel-parallel finalize launch below instead of the in-kernel tail
    p.out_f32 = out_f32; p.out_cs = out_cs;
    p.out_pl = (__nv_bfloat16 *)out_planes; p.out_pl_stride = out_plane_stride; p.split_cs = split_cs; p.nsplit_out = nsplit_out;
    p.n_tiles_n = ceil_div(Cout, p.BN);
    p.tmem_cols = 32;
    p.acc_stride = (p.BN + 31) / 32 * 32;
    while ((int)p.tmem_cols < 2 * p.acc_stride) p.tmem_cols *= 2;  // two accumulators (double buffering across tiles)
    const int num_k = kh * kw * p.cin_blocks;
    const int stage_bytes = nsplit * (kATileBytes + p.BN * kBlockK * 2);
    // Shared-memory budget per CTA.  Wide tiles (BN = 256) are MMA-bound with one CTA per SM and a deep ring.  Narrow
    // tiles (BN <= 128: small-channel layers, per-point MLPs) have short K loops and are latency-bound when the TMA ->
    // MMA -> epilogue chain of a single resident CTA is serial, so they get a smaller ring and two CTAs share the SM
    // (measured on up_3, 64->64 3x3 @192^2: 1.52 ms -> 0.88 ms; profiles/r1_conv_small_tiles.txt).
    const long long n_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_b * ceil_div(Cout, p.BN);
    int budget_kb = (p.BN <= 128 && n_tiles >= 2 * kNumSMs) ? 110 : 225;
    static const int budget_override = env_int("ISTNET_CG_SMEM_KB", 0);
    if (budget_override > 0) budget_kb = budget_override;
    // Optional CTA pairs (ISTNET_CG_CLUSTER=2): wide tiles with one CTA per SM, at least two row tiles
    static const int cluster_env = env_int("ISTNET_CG_CLUSTER", 1);
    const long long m_tiles_h = (long long)p.tiles_w * p.tiles_h * p.tiles_b;
    p.cluster = (cluster_env == 2 && budget_kb == 225 && p.BN >= 128 && (p.BN % 16) == 0 && m_tiles_h >= 2) ? 2 : 1;
    const int stat_bytes = stat_part ? 8 * Cout * (int)sizeof(float) : 0;
    int max_stages = (budget_kb * 1024 - 1024 - 256 - stat_bytes) / stage_bytes;
    if (max_stages < 1) max_stages = 1;
    if (max_stages > 8) max_stages = 8;
    // The ring runs across the tiles of a persistent CTA, so its depth is bounded by the k-iterations of ALL the CTA's
    // tiles, not of one: short-K layers (num_k = 1..4: the set-abstraction MLPs and their data gradients) would otherwise
    // get a 1-deep ring and serialise TMA latency -> MMA -> stage release per tile.
    const long long iters_per_cta = (long long)num_k * ceil_div_ll(n_tiles, kNumSMs);
    p.stages = iters_per_cta < max_stages ? (int)iters_per_cta : max_stages;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + stat_bytes;
    if (smem > 227 * 1024) return ISTNET_ERR_UNSUPPORTED;

    // dynamic + static shared memory share the 227 KB per-CTA limit (the ticket tail of the statistics epilogue keeps a few
    // bytes of static shared memory)
    static int static_smem = -1;
    if (static_smem < 0) {
        cudaFuncAttributes fa;
        ISTNET_CUDA_TRY(cudaFuncGetAttributes(&fa, conv_gemm_tc_kernel<1>));
        static_smem = (int)fa.sharedSizeBytes;
    }
    if (smem + (size_t)static_smem > 227 * 1024) return ISTNET_ERR_UNSUPPORTED;
    ISTNET_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - static_smem));
    if (p.cluster == 2)
        ISTNET_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - static_smem));
    int ctas_per_sm = (int)((227 * 1024) / (smem + static_smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 512 / (int)p.tmem_cols) ctas_per_sm = 512 / (int)p.tmem_cols;  // tensor memory: 512 columns per SM
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 2) ctas_per_sm = 2;  // grid <= 2 * 148 = 296: the size callers give the statistics scratch (stat_part)
    if (ctas_per_sm != 1) p.cluster = 1;  // CTA pairs only for the one-CTA-per-SM configuration; decided before the weight map's box is fixed
    CUtensorMap ta, tb;
    {
        uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)nsplit};
        uint64_t str[4] = {(uint64_t)act_cs * 2, (uint64_t)W * act_cs * 2, (uint64_t)H * W * act_cs * 2, (uint64_t)act_plane_stride * 2};
        uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)p.box_b, 1u};
        int e = istnet_make_tmap_bf16(&ta, act_planes, 5, dims, str, box, kBlockK * 2);
        if (e) return e;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw), (uint64_t)nsplit};
        uint64_t str[3] = {(uint64_t)wgt_cs * 2, (uint64_t)Cout * wgt_cs * 2, (uint64_t)wgt_plane_stride * 2};
        uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(p.BN / p.cluster), 1u, 1u};  // cluster: every CTA fetches its half of the rows
        int e = istnet_make_tmap_bf16(&tb, wgt_planes, 4, dims, str, box, kBlockK * 2);
        if (e) return e;
    }
    long long grid_x = (long long)kNumSMs * ctas_per_sm;
    if (grid_x > n_tiles) grid_x = n_tiles;
    if (grid_out) *grid_out = (int)grid_x;
    int threads = ctas_per_sm >= 2 ? kThreads : kThreadsWide;
    static const int threads_override = env_int("ISTNET_CG_THREADS", 0);
    if (threads_override > 0) threads = threads_override;
    if (p.cluster == 2) {
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        static int max_clusters = -1;  // co-resident CTA pairs at the full shared-memory footprint (GPC granularity)
        if (max_clusters < 0) {
            cfg.gridDim = dim3((unsigned)kNumSMs / 2 * 2);
            cudaLaunchConfig_t q = cfg;
            q.dynamicSmemBytes = 227 * 1024 - static_smem;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_tc_kernel<2>, &q) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = kNumSMs / 2 - 2; }
            max_clusters = n;
        }
        const long long n_work = ((m_tiles_h + 1) / 2) * ceil_div(Cout, p.BN);
        long long pairs = max_clusters < kNumSMs / 2 ? max_clusters : kNumSMs / 2;
        if (pairs > n_work) pairs = n_work;
        grid_x = 2 * pairs;
        if (grid_out) *grid_out = (int)grid_x;
        cfg.gridDim = dim3((unsigned)grid_x);
        ISTNET_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<2>, ta, tb, p));
    } else {
        conv_gemm_tc_kernel<1><<<(unsigned)grid_x, threads, smem, (cudaStream_t)stream>>>(ta, tb, p);
    }
    ISTNET_LAUNCH_CHECK();
    if (fin_p.kind != 0 && !fin_in_kernel(Cout))
        return istnet_fin_finalize_launch(stat_part, (int)grid_x, Cout, fin_p.kind == ISTNET_FIN_COLSUM ? 1 : 2, fin_p, (cudaStream_t)stream);
    return ISTNET_OK;
}
