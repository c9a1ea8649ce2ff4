// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), FP32-accurate.
//
//   out[p, n] = sum_{tap} sum_{c} act[p + shift(tap), c] * wgt[tap][n][c]      (stride 1, "same" padding)
//
// covers every stride-1 convolution of the image branch (modules.py / resnet.py), its data
// gradient (same kernel, flipped+transposed weights), and all 1x1 convolutions / per-point MLP layers (taps = 1).
//
// Precision: the 1e-4 relative target rules out single-pass TF32/BF16.  Each FP32 operand x is carried as a
// bf16 pair (hi = bf16(x), lo = bf16(x - hi), residual <= 2^-18 |x|) and each product is evaluated as
// lo*hi + hi*lo + hi*hi with FP32 accumulation in TMEM (3 tcgen05.mma kind::f16 per k-step, error ~3e-6 RMS per
// product before averaging over K) — 3 BF16 passes cost half of what 3xTF32 would.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0 : TMA producer  — per (tap, 64-channel block): 2 activation boxes (hi/lo; 4-D map over C,W,H,B with
//            box 64 x bw x bh x bb = 128 pixels, coordinates shifted by the tap, out-of-bounds = zero padding)
//            and 2 weight boxes (hi/lo; 64 x BN) into a SWIZZLE_128B ring, completion on mbarriers
//   warp 1 : MMA issuer    — one elected thread, 4 k-steps x 3 products per stage, tcgen05.commit frees the stage
//   warps 2-5 : epilogue   — tcgen05.ld 32x32b, + bias, ReLU, FP32 and/or split-bf16 stores (NHWC rows)
#include "tc_common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                          // bf16 elements = 128 B = one swizzle row
constexpr int kATileBytes = kTileM * kBlockK * 2;    // 16 KB
constexpr int kThreads = 192;

struct ConvGemmParams {
    int B, H, W;              // output (= input) extent
    int box_w, box_h, box_b;  // pixel tile, box_w*box_h*box_b == 128
    int tiles_w, tiles_h, tiles_b;
    int kh, kw;               // taps (1x1 or 3x3 ...), padding = k/2
    int cin_blocks;           // ceil(Cin / 64)
    int Cout, BN;             // logical output channels, tile width (multiple of 16, <= 256)
    int stages;
    const float *bias;        // [Cout] or null
    int relu;
    float *out_f32;           // [B,H,W,out_cs] or null
    int out_cs;
    __nv_bfloat16 *out_hi, *out_lo;  // [B,H,W,split_cs] or null
    int split_cs;
    uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const ConvGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_tile_bytes = p.BN * kBlockK * 2;
    const int stage_bytes = 2 * kATileBytes + 2 * b_tile_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + p.stages;
    uint64_t *tmem_full_bar = empty_bar + p.stages;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile coordinates
    int mt = blockIdx.x;
    const int tw = mt % p.tiles_w; mt /= p.tiles_w;
    const int th = mt % p.tiles_h; mt /= p.tiles_h;
    const int tb = mt;
    const int w0 = tw * p.box_w, h0 = th * p.box_h, b0 = tb * p.box_b;
    const int n0 = blockIdx.y * p.BN;
    const int num_k = p.kh * p.kw * p.cin_blocks;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo); tc::prefetch_tmap(&tm_b_hi); tc::prefetch_tmap(&tm_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_holder, p.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int pad_h = p.kh / 2, pad_w = p.kw / 2;
            int it = 0;
            for (int tap = 0; tap < p.kh * p.kw; ++tap) {
                const int r = tap / p.kw, s = tap % p.kw;
                for (int kb = 0; kb < p.cin_blocks; ++kb, ++it) {
                    const int st = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    tc::mbar_wait(&empty_bar[st], ph ^ 1);
                    uint8_t *sa = smem + (size_t)st * stage_bytes;
                    tc::mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
                    tc::tma_load_4d(sa, &tm_a_hi, &full_bar[st], kb * kBlockK, w0 + s - pad_w, h0 + r - pad_h, b0);
                    tc::tma_load_4d(sa + kATileBytes, &tm_a_lo, &full_bar[st], kb * kBlockK, w0 + s - pad_w, h0 + r - pad_h, b0);
                    tc::tma_load_3d(sa + 2 * kATileBytes, &tm_b_hi, &full_bar[st], kb * kBlockK, n0, tap);
                    tc::tma_load_3d(sa + 2 * kATileBytes + b_tile_bytes, &tm_b_lo, &full_bar[st], kb * kBlockK, n0, tap);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t id_hh = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 0, 0), id_hl = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 0, 0);
            const uint32_t id_lh = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 0, 0), id_ll = tc::make_idesc_f16(kTileM, p.BN, 1, 1, 0, 0);
            for (int it = 0; it < num_k; ++it) {
                const int st = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                tc::mbar_wait(&full_bar[st], ph);
                tc::tc_fence_after();
                const uint32_t a_hi = tc::smem_u32(smem + (size_t)st * stage_bytes);
                const uint32_t a_lo = a_hi + kATileBytes;
                const uint32_t b_hi = a_hi + 2 * kATileBytes;
                const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
                for (int j = 0; j < kBlockK / 16; ++j) {
                    const uint64_t dah = tc::make_desc_sw128(a_hi + j * 32, 16, 1024);
                    const uint64_t dal = tc::make_desc_sw128(a_lo + j * 32, 16, 1024);
                    const uint64_t dbh = tc::make_desc_sw128(b_hi + j * 32, 16, 1024);
                    const uint64_t dbl = tc::make_desc_sw128(b_lo + j * 32, 16, 1024);
                    tc::umma_bf16(tmem_base, dal, dbh, id_lh, (it | j) != 0);  // small terms first
                    tc::umma_bf16(tmem_base, dah, dbl, id_hl, 1);
                    tc::umma_bf16(tmem_base, dah, dbh, id_hh, 1);
                }
                tc::umma_commit(&empty_bar[st]);  // frees the smem stage once these MMAs retire
            }
            tc::umma_commit(tmem_full_bar);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int pb = row / (p.box_w * p.box_h);
        const int ph_ = (row / p.box_w) % p.box_h;
        const int pw = row % p.box_w;
        const int b = b0 + pb, h = h0 + ph_, w = w0 + pw;
        const bool row_ok = (b < p.B) && (h < p.H) && (w < p.W);
        const size_t pix = ((size_t)b * p.H + h) * p.W + w;
        tc::mbar_wait(tmem_full_bar, 0);
        tc::tc_fence_after();
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tc::tmem_ld_wait();
            if (!row_ok) continue;
            const int n = n0 + c0;
            if (n >= p.Cout) continue;
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float x = __uint_as_float(v[i]);
                if (p.bias != nullptr && n + i < p.Cout) x += __ldg(p.bias + n + i);
                if (p.relu) x = fmaxf(x, 0.f);
                f[i] = x;
            }
            const int valid = min(32, p.Cout - n);
            if (p.out_f32) {
                float *o = p.out_f32 + pix * p.out_cs + n;
                if (valid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
                } else {
                    for (int i = 0; i < valid; ++i) o[i] = f[i];
                }
            }
            if (p.out_hi) {
                __nv_bfloat16 *oh = p.out_hi + pix * p.split_cs + n;
                __nv_bfloat16 *ol = p.out_lo + pix * p.split_cs + n;
                if (valid == 32 && ((reinterpret_cast<uintptr_t>(oh) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint32_t hh[4], ll[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            unsigned short h0_, h1_, l0_, l1_;
                            split_hi_lo(f[i + 2 * k], h0_, l0_);
                            split_hi_lo(f[i + 2 * k + 1], h1_, l1_);
                            hh[k] = (uint32_t)h0_ | ((uint32_t)h1_ << 16);
                            ll[k] = (uint32_t)l0_ | ((uint32_t)l1_ << 16);
                        }
                        *reinterpret_cast<uint4 *>(oh + i) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                        *reinterpret_cast<uint4 *>(ol + i) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                    }
                } else {
                    for (int i = 0; i < valid; ++i) {
                        unsigned short hv, lv;
                        split_hi_lo(f[i], hv, lv);
                        oh[i] = __ushort_as_bfloat16(hv);
                        ol[i] = __ushort_as_bfloat16(lv);
                    }
                }
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
                          const uint32_t *box) {
    PFN_tmapEncodeTiled enc = istnet_get_tmap_encoder();
    if (!enc) return ISTNET_ERR_UNSUPPORTED;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ISTNET_OK : ISTNET_ERR_BAD_ARG;
}

static int pick_bn(int cout) {
    if (cout >= 256) return 256;
    int bn = (cout + 15) / 16 * 16;
    return bn < 16 ? 16 : bn;
}

extern "C" int istnet_conv_gemm(const void *act_hi, const void *act_lo, int B, int H, int W, int Cin, int act_cs, const void *wgt_hi,
                                const void *wgt_lo, int Cout, int wgt_cs, int kh, int kw, const float *bias, int relu, float *out_f32,
                                int out_cs, void *out_hi, void *out_lo, int split_cs, int box_w, int box_h, void *stream) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return ISTNET_ERR_BAD_ARG;
    if ((act_cs & 7) || (wgt_cs & 7) || act_cs < Cin || wgt_cs < Cin) return ISTNET_ERR_BAD_ARG;
    if (box_w <= 0 || box_h <= 0 || (kTileM % (box_w * box_h)) != 0 || box_w > 256 || box_h > 256) return ISTNET_ERR_BAD_ARG;
    if ((kh & 1) == 0 || (kw & 1) == 0) return ISTNET_ERR_UNSUPPORTED;
    if (out_hi && ((split_cs & 7) || !out_lo)) return ISTNET_ERR_BAD_ARG;
    ConvGemmParams p{};
    p.B = B; p.H = H; p.W = W;
    p.box_w = box_w; p.box_h = box_h; p.box_b = kTileM / (box_w * box_h);
    p.tiles_w = ceil_div(W, box_w); p.tiles_h = ceil_div(H, box_h); p.tiles_b = ceil_div(B, p.box_b);
    p.kh = kh; p.kw = kw;
    p.cin_blocks = ceil_div(Cin, kBlockK);
    p.Cout = Cout; p.BN = pick_bn(Cout);
    p.bias = bias; p.relu = relu;
    p.out_f32 = out_f32; p.out_cs = out_cs;
    p.out_hi = (__nv_bfloat16 *)out_hi; p.out_lo = (__nv_bfloat16 *)out_lo; p.split_cs = split_cs;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < p.BN) p.tmem_cols *= 2;
    const int num_k = kh * kw * p.cin_blocks;
    const int stage_bytes = 2 * kATileBytes + 2 * p.BN * kBlockK * 2;
    int max_stages = (225 * 1024 - 1024 - 256) / stage_bytes;
    if (max_stages > 6) max_stages = 6;
    p.stages = num_k < max_stages ? num_k : max_stages;
    if (p.stages < 1) return ISTNET_ERR_UNSUPPORTED;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;

    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)act_cs * 2, (uint64_t)W * act_cs * 2, (uint64_t)H * W * act_cs * 2};
        uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)p.box_b};
        int e = istnet_make_tmap_bf16(&ta_hi, act_hi, 4, dims, str, box);
        if (e) return e;
        e = istnet_make_tmap_bf16(&ta_lo, act_lo, 4, dims, str, box);
        if (e) return e;
    }
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
        uint64_t str[2] = {(uint64_t)wgt_cs * 2, (uint64_t)Cout * wgt_cs * 2};
        uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)p.BN, 1u};
        int e = istnet_make_tmap_bf16(&tb_hi, wgt_hi, 3, dims, str, box);
        if (e) return e;
        e = istnet_make_tmap_bf16(&tb_lo, wgt_lo, 3, dims, str, box);
        if (e) return e;
    }
    ISTNET_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid(p.tiles_w * p.tiles_h * p.tiles_b, ceil_div(Cout, p.BN));
    conv_gemm_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
