// Weight gradient of the stride-1 "same" convolutions / 1x1 layers on tcgen05 tensor cores, FP32-accurate.
//
//   dW[tap][co][ci] = sum_{pixels p} dY[p][co] * X[p + shift(tap)][ci]
//
// GEMM view: M = output channels (128 per CTA), N = input channels (BN <= 256 per CTA), K = pixels.  Both operands
// are channels-last activations, i.e. "MN-major" for the tensor core: a TMA box of 64 pixels x 64 channels
// (SWIZZLE_128B, 8 KB) is exactly one canonical MN-major swizzle slab (8-pixel groups 1024 B apart), and
// consecutive 64-channel slabs sit 8 KB apart (the descriptor's leading-byte-offset).  The tap shift and the
// zero padding are the TMA coordinates / out-of-bounds fill, exactly as in the forward kernel.
// bf16-plane arithmetic as in conv_gemm_tc.cu: all plane products dY_i * X_j with i + j < nsplit, smallest first.
// Narrow layers (Cout <= 64: up_2 / up_3, ResNet layer1, the fine PointNet++ levels) would leave half of the 128 accumulator rows
// multiplying zero padding.  For them the shift moves from X to dY — sum_p dY[p] X[p + t] = sum_q dY[q - t] X[q], with the same
// zero fill outside the image — so that ONE unshifted X tile serves two taps whose (differently shifted) dY tiles fill rows
// 0..63 and 64..127: 5 instead of 9 CTA columns for a 3x3 filter, every MMA fully used (up_3: 679 -> see profiles/).
// K (pixels) is split across CTAs; each CTA writes its FP32 partial tile and a second tiny kernel reduces the
// partials in a fixed order into the PyTorch weight layout [Cout][Cin][kh][kw] — deterministic, unlike the
// atomics of the reference's custom grads (SURVEY.md §7 hard part 5).
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;

struct WgradParams {
    int B, H, W;
    int pix_tile;             // pixels per pipeline stage (K of the GEMM): 64 or 32
    int box_w, box_h, box_b;  // box_w*box_h*box_b == pix_tile
    int tiles_w, tiles_h, tiles_b, num_pix_tiles;
    int kh, kw;
    int Cout, Cin, BN;        // BN multiple of 64
    int n_ci_tiles;
    int ksplit, stages, nsplit;
    int pair;                 // Cout <= 64 and more than one tap: the CTA's 128 accumulator rows hold TWO filter taps (rows 64.. = the second)
    float *partial;           // [ksplit][taps][Cout][Cin]
    uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kPixTile = p.pix_tile;
    const int kSlabBytes = kPixTile * 64 * 2;    // pix_tile pixels x 64 channels bf16
    const int n_slabs_b = p.BN / 64;
    const int a_bytes = 2 * kSlabBytes;          // 128 output channels = 2 slabs (per plane)
    const int b_bytes = n_slabs_b * kSlabBytes;
    const int stage_bytes = p.nsplit * (a_bytes + b_bytes);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + p.stages;
    uint64_t *tmem_full_bar = empty_bar + p.stages;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the filter tap is the fastest grid index: the kh*kw CTAs that read the same dY / (shifted) X pixel tiles are
    // co-scheduled and share them through L2
    const int taps = p.kh * p.kw;
    const int tap = p.pair ? 2 * blockIdx.x : blockIdx.x;  // pair mode: taps `tap` and `tap + 1` (the latter may not exist)
    const int co0 = (blockIdx.y / p.n_ci_tiles) * kTileM;
    const int ci0 = (blockIdx.y % p.n_ci_tiles) * p.BN;
    const int split = blockIdx.z;
    const int r = tap / p.kw, s = tap % p.kw;
    const int per = (p.num_pix_tiles + p.ksplit - 1) / p.ksplit;
    const int t_begin = split * per;
    const int t_end = min(p.num_pix_tiles, t_begin + per);
    const int num_k = max(0, t_end - t_begin);

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_dy); tc::prefetch_tmap(&tm_x);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_holder, p.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (lane == 0) {
            const int pad_h = p.kh / 2, pad_w = p.kw / 2;
            for (int it = 0; it < num_k; ++it) {
                int t = t_begin + it;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = t * p.box_b;
                const int st = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                tc::mbar_wait(&empty_bar[st], ph ^ 1);
                uint8_t *sa = smem + (size_t)st * stage_bytes;
                tc::mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
                uint8_t *sb = sa + p.nsplit * a_bytes;
                for (int pl = 0; pl < p.nsplit; ++pl) {
                    if (p.pair) {
                        for (int c = 0; c < 2; ++c) {  // slab c = dY shifted by minus tap (tap + c); a tap past the filter reads channels >= Cout: zero fill
                            const int tc_ = tap + c, rc = tc_ / p.kw, sc = tc_ % p.kw;
                            const bool live = tc_ < taps;
                            tc::tma_load_5d(sa + pl * a_bytes + c * kSlabBytes, &tm_dy, &full_bar[st], live ? 0 : p.Cout + 64,
                                            live ? w0 - (sc - pad_w) : w0, live ? h0 - (rc - pad_h) : h0, b0, pl);
                        }
                        for (int c = 0; c < n_slabs_b; ++c)
                            tc::tma_load_5d(sb + pl * b_bytes + c * kSlabBytes, &tm_x, &full_bar[st], ci0 + c * 64, w0, h0, b0, pl);
                    } else {
                        for (int c = 0; c < 2; ++c)
                            tc::tma_load_5d(sa + pl * a_bytes + c * kSlabBytes, &tm_dy, &full_bar[st], co0 + c * 64, w0, h0, b0, pl);
                        for (int c = 0; c < n_slabs_b; ++c)
                            tc::tma_load_5d(sb + pl * b_bytes + c * kSlabBytes, &tm_x, &full_bar[st], ci0 + c * 64, w0 + s - pad_w, h0 + r - pad_h, b0, pl);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // both operands MN-major
            const uint32_t idesc = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 1, 1);
            for (int it = 0; it < num_k; ++it) {
                const int st = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                tc::mbar_wait(&full_bar[st], ph);
                tc::tc_fence_after();
                const uint32_t a0 = tc::smem_u32(smem + (size_t)st * stage_bytes);
                const uint32_t b0 = a0 + p.nsplit * a_bytes;
                for (int j = 0; j < kPixTile / 16; ++j) {  // 16 pixels (k) per MMA = 2 groups of 8 rows x 128 B
                    const uint32_t off = j * 2048;
                    uint32_t acc = (it | j) != 0;
                    for (int sum = p.nsplit - 1; sum >= 0; --sum) {
                        for (int ia = sum; ia >= 0; --ia) {
                            const int ib = sum - ia;
                            const uint64_t da = tc::make_desc_sw128(a0 + ia * a_bytes + off, kSlabBytes, 1024);
                            const uint64_t db = tc::make_desc_sw128(b0 + ib * b_bytes + off, kSlabBytes, 1024);
                            tc::umma_bf16(tmem_base, da, db, idesc, acc);
                            acc = 1;
                        }
                    }
                }
                tc::umma_commit(&empty_bar[st]);
            }
            tc::umma_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                                   // accumulator row
        const int my_tap = p.pair ? tap + (m >> 6) : tap;              // pair mode: rows 64.. belong to the second tap
        const int co = p.pair ? (my_tap < taps ? (m & 63) : p.Cout) : co0 + m;  // co >= Cout: nothing to store
        float *dst = p.partial + (((size_t)split * taps + min(my_tap, taps - 1)) * p.Cout + min(co, p.Cout - 1)) * p.Cin;
        if (num_k > 0) {
            tc::mbar_wait(tmem_full_bar, 0);
            tc::tc_fence_after();
        }
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
            uint32_t v[32];
            if (num_k > 0) {
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tc::tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            if (co >= p.Cout) continue;
            const int ci = ci0 + c0;
            if (ci >= p.Cin) continue;
            const int valid = min(32, p.Cin - ci);
            float *o = dst + ci;
            if (valid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4 *>(o + i) =
                        make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)  // predicated, fully unrolled: a dynamic index would put v[] in local memory
                    if (i < valid) o[i] = __uint_as_float(v[i]);
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// grad_w[co][ci][tap] = sum_split partial[split][tap][co][ci]   (fixed summation order).  L = 1, 8 or 32 lanes share one
// output element (lane q sums splits q, q+L, ...; the lanes are then combined by a fixed xor tree): small weight tensors
// with many splits would otherwise leave one thread walking 100+ dependent-latency loads.
template <int L>
__global__ void wgrad_reduce_kernel(int ksplit, int taps, int Cout, int Cin, const float *__restrict__ partial, float *__restrict__ grad_w) {
    const long long total = (long long)taps * Cout * Cin;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (total + 31) / 32 * 32 * L; t += nthreads) {  // whole warps stay in the loop
        const long long i = t / L;
        const int q = (int)(t % L);
        float acc = 0.f;
        if (i < total) {
            int sp = q;
            for (; sp + 7 * L < ksplit; sp += 8 * L) {  // 8 independent loads in flight, summed in a fixed order
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = partial[(size_t)(sp + u * L) * total + i];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
            for (; sp < ksplit; sp += L) acc += partial[(size_t)sp * total + i];
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (i < total && q == 0) {
            const int ci = (int)(i % Cin);
            const long long rest = i / Cin;
            const int co = (int)(rest % Cout);
            const int tap = (int)(rest / Cout);
            grad_w[((size_t)co * Cin + ci) * taps + tap] = acc;
        }
    }
}

}  // namespace

static int wg_env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}
static int wgrad_pick_pix() {
    static const int f = wg_env_int("ISTNET_WG_PIX", 0);  // knobs are read once per process
    return (f == 32 || f == 64) ? f : 64;
}
static int wgrad_pick_bn(int cin, int nsplit) {
    int bn = (cin + 63) / 64 * 64;
    int cap = nsplit >= 3 ? 128 : 256;  // keep >= 2 pipeline stages
    static const int f = wg_env_int("ISTNET_WG_BN", 0);
    if (f == 64 || f == 128 || f == 256) cap = f;
    return bn > cap ? cap : bn;
}

// Shared-memory budget per CTA: wide tiles run one deep-ring CTA per SM; narrow multi-tap tiles (small channel counts)
// are bound by L2->SM delivery / latency and do better with several shallow CTAs per SM (measured).
static int wgrad_budget_kb(int bn, int taps) {
    static const int f = wg_env_int("ISTNET_WG_SMEM_KB", 0);
    return f > 0 ? f : ((bn <= 128 && taps > 1) ? 60 : 225);
}
static bool wgrad_pair(int cout, int taps) {
    static const int off = wg_env_int("ISTNET_WG_NOPAIR", 0);
    return !off && cout <= 64 && taps > 1;
}
static int wgrad_stage_bytes(int bn, int nsplit, int pix) { return nsplit * (2 + bn / 64) * pix * 64 * 2; }

extern "C" int istnet_wgrad_ksplit(int B, int H, int W, int Cout, int Cin, int kh, int kw, int nsplit) {
    // fill the chip with an integral number of co-resident CTA "waves", at least 4 pixel tiles per CTA
    const int pix = wgrad_pick_pix();
    const int bn = wgrad_pick_bn(Cin, nsplit);
    const int tap_cols = wgrad_pair(Cout, kh * kw) ? (kh * kw + 1) / 2 : kh * kw;
    const int base = ceil_div(Cout, kTileM) * ceil_div(Cin, bn) * tap_cols;
    const long long pix_tiles = ((long long)B * H * W + pix - 1) / pix;
    const int stage_bytes = wgrad_stage_bytes(bn, nsplit, pix);
    int stages = (wgrad_budget_kb(bn, kh * kw) * 1024 - 1280) / stage_bytes;
    if (stages < 1) stages = 1;
    if (stages > 6) stages = 6;
    int per_sm = (227 * 1024) / (stages * stage_bytes + 1280);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int capacity = kNumSMs * per_sm;
    int ks = capacity / base;
    if (ks > pix_tiles / 4) ks = (int)(pix_tiles / 4);
    if (ks < 1) ks = 1;
    static const int cap = wg_env_int("ISTNET_WG_KSCAP", 148);
    if (ks > cap) ks = cap;
    return ks;
}

extern "C" int istnet_conv_wgrad(const void *dy_planes, long long dy_plane_stride, int dy_cs, const void *x_planes, long long x_plane_stride,
                                 int x_cs, int nsplit, int B, int H, int W, int Cout, int Cin, int kh, int kw, float *partial_ws, int ksplit,
                                 float *grad_w, int box_w, int box_h, void *stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || ksplit <= 0) return ISTNET_ERR_BAD_ARG;
    if (nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if ((dy_cs & 7) || (x_cs & 7) || dy_cs < Cout || x_cs < Cin) return ISTNET_ERR_BAD_ARG;
    const int kPixTile = wgrad_pick_pix();
    const int kSlabBytes = kPixTile * 64 * 2;
    if (kPixTile == 32) {  // halve the caller's 64-pixel box
        if (box_h > 1) box_h /= 2; else box_w /= 2;
    }
    if (box_w <= 0 || box_h <= 0 || (kPixTile % (box_w * box_h)) != 0) return ISTNET_ERR_BAD_ARG;
    if ((kh & 1) == 0 || (kw & 1) == 0) return ISTNET_ERR_UNSUPPORTED;
    WgradParams p{};
    p.pix_tile = kPixTile;
    p.B = B; p.H = H; p.W = W;
    p.box_w = box_w; p.box_h = box_h; p.box_b = kPixTile / (box_w * box_h);
    p.tiles_w = ceil_div(W, box_w); p.tiles_h = ceil_div(H, box_h); p.tiles_b = ceil_div(B, p.box_b);
    p.num_pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    p.kh = kh; p.kw = kw; p.Cout = Cout; p.Cin = Cin;
    p.nsplit = nsplit;
    p.BN = wgrad_pick_bn(Cin, nsplit);
    p.n_ci_tiles = ceil_div(Cin, p.BN);
    p.ksplit = ksplit;
    p.pair = wgrad_pair(Cout, kh * kw) ? 1 : 0;
    p.partial = partial_ws;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < p.BN) p.tmem_cols *= 2;
    const int stage_bytes = nsplit * (2 * kSlabBytes + (p.BN / 64) * kSlabBytes);
    int budget_kb = wgrad_budget_kb(p.BN, kh * kw);
    int max_stages = (budget_kb * 1024 - 1024 - 256) / stage_bytes;
    if (max_stages < 1) max_stages = 1;
    if (max_stages > 6) max_stages = 6;
    const int per = ceil_div(p.num_pix_tiles, ksplit);
    p.stages = per < max_stages ? per : max_stages;
    if (p.stages < 1) p.stages = 1;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;

    CUtensorMap t_dy, t_x;
    uint32_t box[5] = {64u, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)p.box_b, 1u};
    {
        uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)nsplit};
        uint64_t str[4] = {(uint64_t)dy_cs * 2, (uint64_t)W * dy_cs * 2, (uint64_t)H * W * dy_cs * 2, (uint64_t)dy_plane_stride * 2};
        int e = istnet_make_tmap_bf16(&t_dy, dy_planes, 5, dims, str, box);
        if (e) return e;
    }
    {
        uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)nsplit};
        uint64_t str[4] = {(uint64_t)x_cs * 2, (uint64_t)W * x_cs * 2, (uint64_t)H * W * x_cs * 2, (uint64_t)x_plane_stride * 2};
        int e = istnet_make_tmap_bf16(&t_x, x_planes, 5, dims, str, box);
        if (e) return e;
    }
    ISTNET_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid(p.pair ? (kh * kw + 1) / 2 : kh * kw, ceil_div(Cout, kTileM) * p.n_ci_tiles, ksplit);
    wgrad_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(t_dy, t_x, p);
    ISTNET_LAUNCH_CHECK();
    const long long total = (long long)kh * kw * Cout * Cin;
    const int L = (ksplit >= 32 && total <= 16384) ? 32 : (ksplit >= 8 && total <= 131072) ? 8 : 1;
    int rgrid = (int)((total * L + 255) / 256);
    if (rgrid > kNumSMs * 8) rgrid = kNumSMs * 8;
    if (L == 32) wgrad_reduce_kernel<32><<<rgrid, 256, 0, (cudaStream_t)stream>>>(ksplit, kh * kw, Cout, Cin, partial_ws, grad_w);
    else if (L == 8) wgrad_reduce_kernel<8><<<rgrid, 256, 0, (cudaStream_t)stream>>>(ksplit, kh * kw, Cout, Cin, partial_ws, grad_w);
    else wgrad_reduce_kernel<1><<<rgrid, 256, 0, (cudaStream_t)stream>>>(ksplit, kh * kw, Cout, Cin, partial_ws, grad_w);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
