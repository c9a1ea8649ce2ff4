// HBM-bound companions of the tensor-core kernels: everything between two contractions of the image branch /
// per-point MLPs is fused into ONE pass per tensor — BatchNorm statistics, BN-apply + residual + ReLU/PReLU +
// Dropout2d scale + split into the bf16 (hi, lo) operand pair, bilinear x2 up-sampling straight into the operand
// pair, BN backward (reduce + apply + split), im2col for the few strided convolutions, the stem max-pool.
// All tensors are channels-last [P pixels][C channels] FP32 (C % 4 == 0, float4 accesses, a warp reads 512
// contiguous bytes); bf16 pairs have a channel stride `cs` (multiple of 8) and a channel offset.
//
// Reference call sites replaced: nn.BatchNorm2d (resnet.py:40-46,129; modules.py:43,65), ReLU/PReLU, Dropout2d
// (modules.py:56,62), nn.Upsample x2 align_corners=True (modules.py:41), MaxPool2d(3,2,1) (resnet.py:131),
// torch.gather of pixel features (ist_net.py:42-45) and their autograd backward formulas.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ticket.cuh"

namespace {

constexpr int kEwThreads = 256;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 bf4_to_f4(const __nv_bfloat16 *p) {
    uint2 r = *reinterpret_cast<const uint2 *>(p);
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
}

struct BnP {           // per-channel BatchNorm parameters (null mean => identity)
    const float *mean, *invstd, *gamma, *beta;
};
__device__ __forceinline__ float4 bn4(float4 y, const BnP &b, int c) {
    if (!b.mean) return y;
    float4 m = ld4(b.mean + c), s = ld4(b.invstd + c), g = ld4(b.gamma + c), be = ld4(b.beta + c);
    // same operation order as ATen's batch_norm transform: (x - mean) * invstd * weight + bias
    return make_float4((y.x - m.x) * s.x * g.x + be.x, (y.y - m.y) * s.y * g.y + be.y, (y.z - m.z) * s.z * g.z + be.z,
                       (y.w - m.w) * s.w * g.w + be.w);
}

// ------------------------------------------------------------------ per-channel reductions
// Each thread owns 4 channels (float4) of rows r = row0 + i*rows_per_iter; partial sums are combined through shared
// memory.  ATOMIC = true: flushed with double atomics into ws[a*C + c] (ws pre-zeroed).  ATOMIC = false: every CTA
// writes its FP32 partial to part[(a*gridDim.x + blockIdx.x)*C + c]; a finalize kernel sums the partials in a fixed
// order — no atomics (hundreds of CTAs hammering a few dozen addresses cost ~25 us per call), no memset, deterministic.
// Cross-thread part of a column reduction for lanes <= kEwThreads: thread (rr, cv) holds the partial sums of 4 channels
// over its rows; the `rows_per_iter` row groups are combined through shared memory in a fixed order.
template <int NACC, bool ATOMIC>
__device__ __forceinline__ void column_flush(float (&acc)[NACC][4], int lanes, int rows_per_iter, bool active, int C, double *ws, float *part) {
    __shared__ float red[kEwThreads * 4];
    for (int a = 0; a < NACC; ++a) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) red[threadIdx.x * 4 + k] = active ? acc[a][k] : 0.f;
        __syncthreads();
        if (threadIdx.x < lanes) {
            for (int k = 0; k < 4; ++k) {
                float s = 0.f;
                for (int q = 0; q < rows_per_iter; ++q) s += red[(q * lanes + threadIdx.x) * 4 + k];
                if (ATOMIC) atomicAdd(ws + (size_t)a * C + threadIdx.x * 4 + k, (double)s);
                else part[((size_t)a * gridDim.x + blockIdx.x) * C + threadIdx.x * 4 + k] = s;
            }
        }
    }
}
template <int NACC, bool ATOMIC, typename F>
__device__ void column_reduce(long long P, int C, double *ws, float *part, F row_fn) {
    const int lanes = C >> 2;  // float4 lanes per row
    float acc[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    if (lanes <= kEwThreads) {
        const int rows_per_iter = kEwThreads / lanes;
        const int cv = threadIdx.x % lanes, rr = threadIdx.x / lanes;
        if (rr < rows_per_iter) {
            for (long long r = (long long)blockIdx.x * rows_per_iter + rr; r < P; r += (long long)gridDim.x * rows_per_iter) row_fn(r, cv * 4, acc);
        }
        column_flush<NACC, ATOMIC>(acc, lanes, rows_per_iter, rr < rows_per_iter, C, ws, part);
    } else {
        for (int cv = threadIdx.x; cv < lanes; cv += kEwThreads) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
            for (long long r = blockIdx.x; r < P; r += gridDim.x) row_fn(r, cv * 4, acc);
            for (int a = 0; a < NACC; ++a)
                for (int k = 0; k < 4; ++k) {
                    if (ATOMIC) atomicAdd(ws + (size_t)a * C + cv * 4 + k, (double)acc[a][k]);
                    else part[((size_t)a * gridDim.x + blockIdx.x) * C + cv * 4 + k] = acc[a][k];
                }
        }
    }
}
// tot[a*C + c] = sum_g part[(a*G + g)*C + c]  in double, fixed order: block = 32 channels x kFinSlices partial-slices.  The
// kernel is pure latency (a few hundred KB out of L2), so every thread issues all its loads of all NACC quantities at once.
constexpr int kFinSlices = 32;
constexpr int kFinThreads = 32 * kFinSlices;
template <int NACC>
__device__ void sum_partials(const float *__restrict__ part, int G, int C, double (&out)[NACC]) {
    __shared__ double sred[NACC][kFinSlices][32];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + l;
    double s[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) s[a] = 0.0;
    if (c < C) {
        for (int g = w; g < G; g += 4 * kFinSlices) {
            float v[NACC][4];
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int u = 0; u < 4; ++u) v[a][u] = (g + u * kFinSlices < G) ? part[((size_t)a * G + g + u * kFinSlices) * C + c] : 0.f;
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int u = 0; u < 4; ++u) s[a] += (double)v[a][u];
        }
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a) sred[a][w][l] = s[a];
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double t = 0.0;
        for (int q = 0; q < kFinSlices; ++q) t += sred[a][q][l];
        out[a] = t;
    }
}

__global__ void __launch_bounds__(kEwThreads) bn_stats_kernel(const float *__restrict__ y, long long P, int C, float *part, FinP fin) {
    column_reduce<2, false>(P, C, nullptr, part, [&](long long r, int c, float (*acc)[4]) {
        float4 v = ld4(y + r * C + c);
        acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
        acc[1][0] += v.x * v.x; acc[1][1] += v.y * v.y; acc[1][2] += v.z * v.z; acc[1][3] += v.w * v.w;
    });
    ticket_finish<2>(fin, part, (int)gridDim.x, C);
}

__global__ void __launch_bounds__(kFinThreads) bn_finalize_kernel(const float *__restrict__ part, int G, long long P, int C, float eps, float momentum,
                                                          float *running_mean, float *running_var, float *mean, float *invstd,
                                                          long long *num_batches_tracked) {
    double t[2];
    sum_partials<2>(part, G, C, t);
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;  // nn.BatchNorm2d's counter, same launch
    if (threadIdx.x >= 32 || c >= C) return;
    double m = t[0] / (double)P;
    double var = t[1] / (double)P - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}
__global__ void __launch_bounds__(kFinThreads) bwd_finalize_kernel(const float *__restrict__ part, int G, int C, double *ws, float *sum0_f32 = nullptr,
                                                           float *sum1_f32 = nullptr) {
    double t[3];
    sum_partials<3>(part, G, C, t);
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x >= 32 || c >= C) return;
    ws[c] = t[0];
    ws[C + c] = t[1];
    ws[2 * C + c] = t[2];
    if (sum0_f32) sum0_f32[c] = (float)t[0];  // FP32 copies in their own tensors: the BatchNorm bias / weight gradients as autograd wants them
    if (sum1_f32) sum1_f32[c] = (float)t[1];
}

// Channel-parallel finish of a reduction described by a FinP (ticket.cuh) from per-CTA partials part[(a*G + g)*C + c]: one CTA per
// 32 channels, 32 partial-slices each — used instead of the in-kernel ticket tail when the reduction is wide (C > kTicketMaxC).
template <int NACC>
__global__ void __launch_bounds__(kFinThreads) fin_finalize_kernel(const float *__restrict__ part, int G, int C, FinP f) {
    __shared__ long long s_nold;
    if (threadIdx.x == 0) s_nold = (f.kind == 1 && f.num_batches_tracked) ? *f.num_batches_tracked : 0;
    double t[NACC];
    sum_partials<NACC>(part, G, C, t);  // contains a __syncthreads: s_nold is visible afterwards
    const long long n_old = s_nold;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x < 32 && c < C) {
        if (f.kind == 1) {
            bn_fin_channel(f, c, t[0], t[NACC >= 2 ? 1 : 0], n_old);
        } else if (f.kind == 2) {
            if (f.sum_f64) f.sum_f64[c] = t[0];
            if (f.sum_f32) f.sum_f32[c] = (float)t[0];
        } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                if (f.sum_f64) f.sum_f64[(size_t)a * C + c] = t[a];
                if (a == 0 && f.sum_f32) f.sum_f32[c] = (float)t[a];
                if (a == 1 && f.sum2_f32) f.sum2_f32[c] = (float)t[a];
            }
        }
    }
    // nn.BatchNorm2d's step counter: incremented by the LAST block, i.e. after every block has read the old value
    if (f.kind == 1 && f.num_batches_tracked) {
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&f.tickets[0], 1u) == gridDim.x - 1) {
            f.tickets[0] = 0u;
            *f.num_batches_tracked = n_old + 1;
        }
    }
}

// Per-thread channel constants: with lanes = C/4 dividing the CTA, a thread keeps the same 4 channels for every row it
// visits, so the BatchNorm parameters live in registers and the row loop contains no division.
struct Chan4 {
    float4 m, s, ga, be;
};
__device__ __forceinline__ Chan4 load_chan(const BnP &b, int c) {
    Chan4 ch;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    ch.m = z; ch.s = z; ch.ga = z; ch.be = z;
    if (b.mean) { ch.m = ld4(b.mean + c); ch.s = ld4(b.invstd + c); ch.ga = ld4(b.gamma + c); ch.be = ld4(b.beta + c); }
    return ch;
}
// ------------------------------------------------------------------ forward: BN + residual + act + noise + split
struct ActFwdP {
    const float *y;         // [P,C]
    BnP bn;
    const float *res;       // optional residual [P,C] (raw, or pre-BN if res_bn.mean != null)
    BnP res_bn;
    int act;                // 0 none, 1 relu, 2 prelu
    const float *prelu_a;   // scalar
    const float *noise;     // optional [B,C] Dropout2d scale
    long long HW;
    float *out_f32;         // optional [P,C]
    __nv_bfloat16 *out_pl;  // optional operand planes [nsplit][P,cs] (+ch_off)
    long long pl_stride;
    int nsplit, cs, ch_off;
};
__device__ __forceinline__ float4 act_fwd4(float4 u, int act, float a) {
    if (act == 1) return make_float4(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f), fmaxf(u.z, 0.f), fmaxf(u.w, 0.f));
    if (act == 2) return make_float4(u.x > 0.f ? u.x : a * u.x, u.y > 0.f ? u.y : a * u.y, u.z > 0.f ? u.z : a * u.z, u.w > 0.f ? u.w : a * u.w);
    return u;
}
__device__ __forceinline__ void act_fwd_store(const ActFwdP &p, long long r, int c, int C, float a, float4 u, bool has_res, float4 rv, bool has_nz,
                                              float4 nz) {
    if (has_res) { u.x += rv.x; u.y += rv.y; u.z += rv.z; u.w += rv.w; }
    float4 z = act_fwd4(u, p.act, a);
    if (has_nz) { z.x *= nz.x; z.y *= nz.y; z.z *= nz.z; z.w *= nz.w; }
    if (p.out_f32) *reinterpret_cast<float4 *>(p.out_f32 + r * C + c) = z;
    if (p.out_pl) store_planes4(p.out_pl + r * p.cs + p.ch_off + c, p.pl_stride, p.nsplit, z);
}
constexpr int kFwdUnroll = 4;
__global__ void __launch_bounds__(kEwThreads, 2) bn_act_split_kernel(long long P, int C, ActFwdP p) {
    const int lanes = C >> 2;
    const float a = (p.act == 2) ? *p.prelu_a : 0.f;
    if (lanes > kEwThreads || kEwThreads % lanes != 0) {  // generic flat (row, lane) loop
        const long long total = P * lanes;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / lanes;
            const int c = (int)(i % lanes) * 4;
            const float4 u = bn4(ld4(p.y + r * C + c), p.bn, c);
            float4 rv = u, nz = u;
            if (p.res) rv = bn4(ld4(p.res + r * C + c), p.res_bn, c);
            if (p.noise) nz = ld4(p.noise + (r / p.HW) * C + c);
            act_fwd_store(p, r, c, C, a, u, p.res != nullptr, rv, p.noise != nullptr, nz);
        }
        return;
    }
    // lanes divides the CTA: a thread keeps its 4 channels, BatchNorm parameters stay in registers, kFwdUnroll rows in flight
    const int rows_per_iter = kEwThreads / lanes;
    const int c = (threadIdx.x % lanes) * 4, rr = threadIdx.x / lanes;
    const Chan4 ch = load_chan(p.bn, c), rch = load_chan(p.res_bn, c);
    const long long stride = (long long)gridDim.x * rows_per_iter;
    for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < P; r0 += stride * kFwdUnroll) {
        float4 yv[kFwdUnroll], rv[kFwdUnroll], nz[kFwdUnroll];
#pragma unroll
        for (int u = 0; u < kFwdUnroll; ++u) {
            const long long r = r0 + u * stride;
            if (r < P) {
                yv[u] = ld4(p.y + r * C + c);
                if (p.res) rv[u] = ld4(p.res + r * C + c);
                if (p.noise) nz[u] = ld4(p.noise + (r / p.HW) * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < kFwdUnroll; ++u) {
            const long long r = r0 + u * stride;
            if (r < P) {
                float4 v = yv[u], q = rv[u];
                // same operation order as ATen's batch_norm transform: (x - mean) * invstd * weight + bias
                if (p.bn.mean)
                    v = make_float4((v.x - ch.m.x) * ch.s.x * ch.ga.x + ch.be.x, (v.y - ch.m.y) * ch.s.y * ch.ga.y + ch.be.y,
                                    (v.z - ch.m.z) * ch.s.z * ch.ga.z + ch.be.z, (v.w - ch.m.w) * ch.s.w * ch.ga.w + ch.be.w);
                if (p.res && p.res_bn.mean)
                    q = make_float4((q.x - rch.m.x) * rch.s.x * rch.ga.x + rch.be.x, (q.y - rch.m.y) * rch.s.y * rch.ga.y + rch.be.y,
                                    (q.z - rch.m.z) * rch.s.z * rch.ga.z + rch.be.z, (q.w - rch.m.w) * rch.s.w * rch.ga.w + rch.be.w);
                act_fwd_store(p, r, c, C, a, v, p.res != nullptr, q, p.noise != nullptr, nz[u]);
            }
        }
    }
}

// ------------------------------------------------------------------ backward: g = (dz1+dz2)*noise*act'(u); BN backward
struct ActBwdP {
    const float *dz, *dz2;  // [P,C] (+ optional second gradient stream)
    const float *y;         // conv output (pre-BN) [P,C]; needed for BN / PReLU
    BnP bn;
    int act;
    const float *prelu_a;
    const __nv_bfloat16 *z_hi;  // ReLU mask source (saved forward output), [P,cs_z]
    int cs_z;
    int batch_stats;            // 1: BN used batch statistics (train) -> full backward; 0: running statistics -> dy = gamma*invstd*g
    const uint8_t *argmax;      // act == 3 (BN + ReLU + max over `ns` consecutive rows): dz is [P/ns][C], routed to the arg-max row
    int ns;
    const float *noise;
    long long HW;
};
// The raw operands of one (row, 4-channel) element of the backward pass.  Loading (bwd_load) is separated from the
// arithmetic (bwd_finish) so that an unrolled row loop issues all its 16-byte loads before the first use.
struct RowIn {
    float4 d, y;
    uint2 zh;
};
__device__ __forceinline__ void bwd_load(const ActBwdP &p, long long r, int c, int C, RowIn &in) {
    if (p.act == 3) {  // gradient of the max over the neighbour axis: only the selected row of each group receives dz
        const unsigned grp = (unsigned)r / (unsigned)p.ns;  // rows < 2^31 (checked by the launcher)
        const int l = (int)((unsigned)r - grp * (unsigned)p.ns);
        const uchar4 am = *reinterpret_cast<const uchar4 *>(p.argmax + (size_t)grp * C + c);
        const float4 dg = ld4(p.dz + (size_t)grp * C + c);
        in.d = make_float4(am.x == l ? dg.x : 0.f, am.y == l ? dg.y : 0.f, am.z == l ? dg.z : 0.f, am.w == l ? dg.w : 0.f);
    } else {
        in.d = ld4(p.dz + r * C + c);
    }
    if (p.dz2) {
        const float4 d2 = ld4(p.dz2 + r * C + c);
        in.d.x += d2.x; in.d.y += d2.y; in.d.z += d2.z; in.d.w += d2.w;
    }
    if (p.noise) {
        const float4 nz = ld4(p.noise + (size_t)((unsigned)r / (unsigned)p.HW) * C + c);
        in.d.x *= nz.x; in.d.y *= nz.y; in.d.z *= nz.z; in.d.w *= nz.w;
    }
    if (p.bn.mean || p.act == 2) in.y = ld4(p.y + r * C + c);
    if (p.act == 1) in.zh = *reinterpret_cast<const uint2 *>(p.z_hi + r * p.cs_z + c);
}
// returns g (gradient w.r.t. u = bn(y) [+res]) and xhat; extra = dz*noise*u*[u<=0] (PReLU slope gradient)
__device__ __forceinline__ void bwd_finish(const ActBwdP &p, const Chan4 &ch, float a, const RowIn &in, float4 &g, float4 &xh, float4 &extra) {
    const float4 d = in.d;
    xh = make_float4(0.f, 0.f, 0.f, 0.f);
    extra = xh;
    float4 u = xh;
    if (p.bn.mean) {
        const float4 yv = in.y;
        xh = make_float4((yv.x - ch.m.x) * ch.s.x, (yv.y - ch.m.y) * ch.s.y, (yv.z - ch.m.z) * ch.s.z, (yv.w - ch.m.w) * ch.s.w);
        if (p.act == 2 || p.act == 3)
            u = make_float4(xh.x * ch.ga.x + ch.be.x, xh.y * ch.ga.y + ch.be.y, xh.z * ch.ga.z + ch.be.z, xh.w * ch.ga.w + ch.be.w);
    } else if (p.act == 2) {
        u = in.y;
    }
    if (p.act == 1) {
        const float4 z = make_float4(__uint_as_float(in.zh.x << 16), __uint_as_float(in.zh.x & 0xffff0000u), __uint_as_float(in.zh.y << 16),
                                     __uint_as_float(in.zh.y & 0xffff0000u));
        g = make_float4(z.x > 0.f ? d.x : 0.f, z.y > 0.f ? d.y : 0.f, z.z > 0.f ? d.z : 0.f, z.w > 0.f ? d.w : 0.f);
    } else if (p.act == 3) {
        g = make_float4(u.x > 0.f ? d.x : 0.f, u.y > 0.f ? d.y : 0.f, u.z > 0.f ? d.z : 0.f, u.w > 0.f ? d.w : 0.f);
    } else if (p.act == 2) {
        g = make_float4(u.x > 0.f ? d.x : a * d.x, u.y > 0.f ? d.y : a * d.y, u.z > 0.f ? d.z : a * d.z, u.w > 0.f ? d.w : a * d.w);
        extra = make_float4(u.x > 0.f ? 0.f : d.x * u.x, u.y > 0.f ? 0.f : d.y * u.y, u.z > 0.f ? 0.f : d.z * u.z, u.w > 0.f ? 0.f : d.w * u.w);
    } else {
        g = d;
    }
}
__device__ __forceinline__ void act_bwd4(const ActBwdP &p, long long r, int c, int C, float a, float4 &g, float4 &xh, float4 &extra) {
    RowIn in;
    bwd_load(p, r, c, C, in);
    bwd_finish(p, load_chan(p.bn, c), a, in, g, xh, extra);
}
constexpr int kBwdUnroll = 4;  // rows in flight per thread: 4 x (2..3) independent 16-byte loads
// ws[0:C] = sum g, ws[C:2C] = sum g*xhat, ws[2C:3C] = PReLU slope partial
__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_reduce_kernel(long long P, int C, ActBwdP p, float *part, FinP fin) {
    const float a = (p.act == 2) ? *p.prelu_a : 0.f;
    const int lanes = C >> 2;
    if (lanes > kEwThreads) {  // very wide rows: generic path
        column_reduce<3, false>(P, C, nullptr, part, [&](long long r, int c, float (*acc)[4]) {
            float4 g, xh, ex;
            act_bwd4(p, r, c, C, a, g, xh, ex);
            acc[0][0] += g.x; acc[0][1] += g.y; acc[0][2] += g.z; acc[0][3] += g.w;
            acc[1][0] += g.x * xh.x; acc[1][1] += g.y * xh.y; acc[1][2] += g.z * xh.z; acc[1][3] += g.w * xh.w;
            acc[2][0] += ex.x; acc[2][1] += ex.y; acc[2][2] += ex.z; acc[2][3] += ex.w;
        });
        ticket_finish<3>(fin, part, (int)gridDim.x, C);
        return;
    }
    float acc[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    const int rows_per_iter = kEwThreads / lanes;
    const int cv = threadIdx.x % lanes, rr = threadIdx.x / lanes;
    if (rr < rows_per_iter) {
        const int c = cv * 4;
        const Chan4 ch = load_chan(p.bn, c);
        const long long stride = (long long)gridDim.x * rows_per_iter;
        for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < P; r0 += stride * kBwdUnroll) {
            RowIn in[kBwdUnroll];
#pragma unroll
            for (int u = 0; u < kBwdUnroll; ++u)
                if (r0 + u * stride < P) bwd_load(p, r0 + u * stride, c, C, in[u]);
#pragma unroll
            for (int u = 0; u < kBwdUnroll; ++u)
                if (r0 + u * stride < P) {
                    float4 g, xh, ex;
                    bwd_finish(p, ch, a, in[u], g, xh, ex);
                    acc[0][0] += g.x; acc[0][1] += g.y; acc[0][2] += g.z; acc[0][3] += g.w;
                    acc[1][0] += g.x * xh.x; acc[1][1] += g.y * xh.y; acc[1][2] += g.z * xh.z; acc[1][3] += g.w * xh.w;
                    acc[2][0] += ex.x; acc[2][1] += ex.y; acc[2][2] += ex.z; acc[2][3] += ex.w;
                }
        }
    }
    column_flush<3, false>(acc, lanes, rows_per_iter, rr < rows_per_iter, C, nullptr, part);
    ticket_finish<3>(fin, part, (int)gridDim.x, C);  // the last CTA writes ws / the BatchNorm parameter gradients: no finalize launch
}
// dy = gamma*invstd*(g - sum_g/P - xhat*sum_gx/P)  (BN)   or   dy = g   (no BN);  dy -> bf16 pair (+ optional FP32 copies)
__device__ __forceinline__ void bwd_apply_store(const ActBwdP &p, const Chan4 &ch, float a, const RowIn &in, const float (&mg)[4],
                                                const float (&mgx)[4], long long r, int c, int C, __nv_bfloat16 *dy_pl, long long pl_stride,
                                                int nsplit, int cs_dy, float *dy_f32, float *g_out) {
    float4 g, xh, ex;
    bwd_finish(p, ch, a, in, g, xh, ex);
    if (g_out) *reinterpret_cast<float4 *>(g_out + r * C + c) = g;
    float4 dy = g;
    if (p.bn.mean) {
        dy.x = ch.ga.x * ch.s.x * (g.x - mg[0] - xh.x * mgx[0]);
        dy.y = ch.ga.y * ch.s.y * (g.y - mg[1] - xh.y * mgx[1]);
        dy.z = ch.ga.z * ch.s.z * (g.z - mg[2] - xh.z * mgx[2]);
        dy.w = ch.ga.w * ch.s.w * (g.w - mg[3] - xh.w * mgx[3]);
    }
    if (dy_pl) store_planes4(dy_pl + r * cs_dy + c, pl_stride, nsplit, dy);
    if (dy_f32) *reinterpret_cast<float4 *>(dy_f32 + r * C + c) = dy;
}
__device__ __forceinline__ void bwd_apply_consts(const ActBwdP &p, const double *__restrict__ ws, long long P, int C, int c, float (&mg)[4],
                                                 float (&mgx)[4]) {
    const double invP = 1.0 / (double)P;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mg[k] = mgx[k] = 0.f;
        if (p.bn.mean && p.batch_stats) { mg[k] = (float)(ws[c + k] * invP); mgx[k] = (float)(ws[C + c + k] * invP); }
    }
}
__global__ void __launch_bounds__(kEwThreads, 2) bn_bwd_apply_kernel(long long P, int C, ActBwdP p, const double *__restrict__ ws,
                                                                     __nv_bfloat16 *dy_pl, long long pl_stride, int nsplit, int cs_dy,
                                                                     float *dy_f32, float *g_out) {
    const int lanes = C >> 2;
    const float a = (p.act == 2) ? *p.prelu_a : 0.f;
    if (lanes > kEwThreads) {  // very wide rows: flat (row, lane) loop
        const long long total = P * lanes;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / lanes;
            const int c = (int)(i % lanes) * 4;
            float mg[4], mgx[4];
            bwd_apply_consts(p, ws, P, C, c, mg, mgx);
            RowIn in;
            bwd_load(p, r, c, C, in);
            bwd_apply_store(p, load_chan(p.bn, c), a, in, mg, mgx, r, c, C, dy_pl, pl_stride, nsplit, cs_dy, dy_f32, g_out);
        }
        return;
    }
    const int rows_per_iter = kEwThreads / lanes;
    const int cv = threadIdx.x % lanes, rr = threadIdx.x / lanes;
    if (rr >= rows_per_iter) return;
    const int c = cv * 4;
    const Chan4 ch = load_chan(p.bn, c);
    float mg[4], mgx[4];
    bwd_apply_consts(p, ws, P, C, c, mg, mgx);
    const long long stride = (long long)gridDim.x * rows_per_iter;
    for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < P; r0 += stride * kBwdUnroll) {
        RowIn in[kBwdUnroll];
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u)
            if (r0 + u * stride < P) bwd_load(p, r0 + u * stride, c, C, in[u]);
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u)
            if (r0 + u * stride < P)
                bwd_apply_store(p, ch, a, in[u], mg, mgx, r0 + u * stride, c, C, dy_pl, pl_stride, nsplit, cs_dy, dy_f32, g_out);
    }
}

// ------------------------------------------------------------------ generic FP32 -> bf16 pair (optionally NCHW -> NHWC), column sums
__global__ void __launch_bounds__(kEwThreads) split_kernel(long long P, int C, const float *__restrict__ x, long long HW, int nchw,
                                                           __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs, int ch_off) {
    const long long total = P * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C;
        const int c = (int)(i % C);
        float v = nchw ? x[((r / HW) * C + c) * HW + (r % HW)] : x[i];
        store_planes1(pl + r * cs + ch_off + c, pl_stride, nsplit, v);
    }
}
// channels-last rows with C % 4 == 0 (every per-point MLP input, feature rows, gradient rows): float4 in, 8-byte plane stores,
// 32-bit index arithmetic
__global__ void __launch_bounds__(kEwThreads) split_rows4_kernel(unsigned P, int C, const float *__restrict__ x, __nv_bfloat16 *pl,
                                                                 long long pl_stride, int nsplit, int cs, int ch_off) {
    const unsigned lanes = C >> 2;
    const unsigned total = P * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / lanes;
        const int c = (int)(i - r * lanes) * 4;
        store_planes4(pl + (size_t)r * cs + ch_off + c, pl_stride, nsplit, ld4(x + (size_t)i * 4));
    }
}
// PyTorch weight [co][ci][kh][kw] -> operand planes [nsplit][tap][rows][cs]:
//   transpose = 0 (forward operand):        rows = co, cols = ci, tap = r*kw + s
//   transpose = 1 (data-gradient operand):  rows = ci, cols = co, tap = flipped (kh-1-r, kw-1-s)
//   im2col   = 1 (strided convs as 1x1 GEMM over patches): one tap, K index = (r*kw + s)*ci_total + ci
__device__ __forceinline__ void prep_weight_range(int co_n, int ci_n, int kh, int kw, const float *__restrict__ w, int transpose, int im2col,
                                                  __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs, long long first, long long step) {
    const long long total = (long long)co_n * ci_n * kh * kw;
    const int taps = kh * kw;
    for (long long i = first; i < total; i += step) {
        // iterate in OUTPUT order so that stores are coalesced
        long long o;
        float v;
        if (im2col) {
            const int K = taps * ci_n;
            if (!transpose) {  // [co][K]
                const int k = (int)(i % K), co = (int)(i / K);
                const int tap = k / ci_n, ci = k % ci_n;
                v = w[((long long)co * ci_n + ci) * taps + tap];
                o = (long long)co * cs + k;
            } else {  // [K][co]
                const int co = (int)(i % co_n), k = (int)(i / co_n);
                const int tap = k / ci_n, ci = k % ci_n;
                v = w[((long long)co * ci_n + ci) * taps + tap];
                o = (long long)k * cs + co;
            }
        } else if (!transpose) {  // [tap][co][ci]
            const int ci = (int)(i % ci_n);
            const long long r = i / ci_n;
            const int co = (int)(r % co_n), tap = (int)(r / co_n);
            v = w[((long long)co * ci_n + ci) * taps + tap];
            o = ((long long)tap * co_n + co) * cs + ci;
        } else {  // [flipped tap][ci][co]
            const int co = (int)(i % co_n);
            const long long r = i / co_n;
            const int ci = (int)(r % ci_n), tap = (int)(r / ci_n);
            v = w[((long long)co * ci_n + ci) * taps + (taps - 1 - tap)];
            o = ((long long)tap * ci_n + ci) * cs + co;
        }
        store_planes1(pl + o, pl_stride, nsplit, v);
    }
}
__device__ __forceinline__ void prep_weight_body(int co_n, int ci_n, int kh, int kw, const float *__restrict__ w, int transpose, int im2col,
                                                 __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs) {
    prep_weight_range(co_n, ci_n, kh, kw, w, transpose, im2col, pl, pl_stride, nsplit, cs, blockIdx.x * (long long)blockDim.x + threadIdx.x,
                      (long long)gridDim.x * blockDim.x);
}
// Every registered weight of the model in ONE launch (istnet_prep_weight_batch): the table lists, per weight, the FP32 source, its
// shape, the forward and (optional) data-gradient operand-plane destinations and the first CTA that works on it; a CTA finds its
// entry by binary search and handles kPrepChunk consecutive elements of it in both layouts.
struct PrepEntry {
    const float *w;
    __nv_bfloat16 *pf, *pt;
    long long stride_f, stride_t;
    int co, ci, kh, kw, im2col, ns_f, cs_f, ns_t, cs_t, block0;
};
constexpr int kPrepChunk = kEwThreads * 8;
__global__ void __launch_bounds__(kEwThreads) prep_weight_batch_kernel(const PrepEntry *__restrict__ table, int n) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {  // last entry with block0 <= blockIdx.x
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PrepEntry e = table[lo];
    const long long base = (long long)((int)blockIdx.x - e.block0) * kPrepChunk;
    const long long total = (long long)e.co * e.ci * e.kh * e.kw;
    const long long end = base + kPrepChunk < total ? base + kPrepChunk : total;
    // prep_weight_range iterates i = first, first + step, ... < total: bound it to this CTA's chunk by handing it the chunk end
    for (long long i = base + threadIdx.x; i < end; i += kEwThreads) {
        prep_weight_range(e.co, e.ci, e.kh, e.kw, e.w, 0, e.im2col, e.pf, e.stride_f, e.ns_f, e.cs_f, i, total);
        if (e.pt) prep_weight_range(e.co, e.ci, e.kh, e.kw, e.w, 1, e.im2col, e.pt, e.stride_t, e.ns_t, e.cs_t, i, total);
    }
}
__global__ void __launch_bounds__(kEwThreads) prep_weight_kernel(int co_n, int ci_n, int kh, int kw, const float *__restrict__ w, int transpose,
                                                                 int im2col, __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs) {
    prep_weight_body(co_n, ci_n, kh, kw, w, transpose, im2col, pl, pl_stride, nsplit, cs);
}
// forward operand and data-gradient operand of one weight in ONE launch (the training step needs both; the second is kept
// on the tape until the backward pass)
__global__ void __launch_bounds__(kEwThreads) prep_weight_pair_kernel(int co_n, int ci_n, int kh, int kw, const float *__restrict__ w, int im2col,
                                                                      __nv_bfloat16 *pl_f, long long stride_f, int nsplit_f, int cs_f,
                                                                      __nv_bfloat16 *pl_t, long long stride_t, int nsplit_t, int cs_t) {
    prep_weight_body(co_n, ci_n, kh, kw, w, 0, im2col, pl_f, stride_f, nsplit_f, cs_f);
    prep_weight_body(co_n, ci_n, kh, kw, w, 1, im2col, pl_t, stride_t, nsplit_t, cs_t);
}
__global__ void __launch_bounds__(kEwThreads) colsum_kernel(const float *__restrict__ x, long long P, int C, double *ws) {
    column_reduce<1, true>(P, C, ws, nullptr, [&](long long r, int c, float (*acc)[4]) {
        float4 v = ld4(x + r * C + c);
        acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
    });
}

// ------------------------------------------------------------------ bilinear x2, align_corners=True (modules.py:41)
__device__ __forceinline__ void up_src(int o, int in, int out, int &i0, int &i1, float &l0, float &l1) {
    // ATen area_pixel_compute_source_index(align_corners=True): src = scale*dst, scale = (in-1)/(out-1)
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float s = scale * (float)o;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}
// same arithmetic with the (host-computed, identical IEEE division) scale hoisted out of the per-element path
__device__ __forceinline__ void up_src_s(int o, int in, float scale, int &i0, int &i1, float &l0, float &l1) {
    const float s = scale * (float)o;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}
// All index arithmetic below is 32-bit (element counts < 2^31, checked by the launchers): a 64-bit div/mod chain per
// element made these passes instruction-bound at a quarter of the HBM rate.
__global__ void __launch_bounds__(kEwThreads) upsample2x_split_kernel(int B, int H, int W, int C, float sh, float sw, const float *__restrict__ x,
                                                                       __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs, float *out_f32) {
    const unsigned lanes = C >> 2, Ho = 2 * H, Wo = 2 * W;
    const unsigned total = (unsigned)B * Ho * Wo * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 4;
        unsigned r = i / lanes;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        int h0, h1, w0, w1;
        float hl0, hl1, wl0, wl1;
        up_src_s(ho, H, sh, h0, h1, hl0, hl1);
        up_src_s(wo, W, sw, w0, w1, wl0, wl1);
        const float *base = x + (size_t)b * H * W * C + c;
        float4 a = ld4(base + ((size_t)h0 * W + w0) * C), bq = ld4(base + ((size_t)h0 * W + w1) * C);
        float4 cq = ld4(base + ((size_t)h1 * W + w0) * C), d = ld4(base + ((size_t)h1 * W + w1) * C);
        float4 o;
        o.x = hl0 * (wl0 * a.x + wl1 * bq.x) + hl1 * (wl0 * cq.x + wl1 * d.x);
        o.y = hl0 * (wl0 * a.y + wl1 * bq.y) + hl1 * (wl0 * cq.y + wl1 * d.y);
        o.z = hl0 * (wl0 * a.z + wl1 * bq.z) + hl1 * (wl0 * cq.z + wl1 * d.z);
        o.w = hl0 * (wl0 * a.w + wl1 * bq.w) + hl1 * (wl0 * cq.w + wl1 * d.w);
        const size_t op = ((size_t)b * Ho + ho) * Wo + wo;
        if (pl) store_planes4(pl + op * cs + c, pl_stride, nsplit, o);
        if (out_f32) *reinterpret_cast<float4 *>(out_f32 + op * C + c) = o;
    }
}
// C % 8 == 0 (every PSPUpsample width): 8 channels per thread — half the index arithmetic per element and one 16-byte store per
// plane (the 4-channel kernel ran at 58 % issue utilisation and 2.8 TB/s, profiles/r2_ncu_families_all.txt).  Same arithmetic order.
__global__ void __launch_bounds__(kEwThreads) upsample2x_split8_kernel(int B, int H, int W, int C, float sh, float sw, const float *__restrict__ x,
                                                                        __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs, float *out_f32) {
    const unsigned lanes = C >> 3, Ho = 2 * H, Wo = 2 * W;
    const unsigned total = (unsigned)B * Ho * Wo * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 8;
        unsigned r = i / lanes;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        int h0, h1, w0, w1;
        float hl0, hl1, wl0, wl1;
        up_src_s(ho, H, sh, h0, h1, hl0, hl1);
        up_src_s(wo, W, sw, w0, w1, wl0, wl1);
        const float *base = x + (size_t)b * H * W * C + c;
        const float *p00 = base + ((size_t)h0 * W + w0) * C, *p01 = base + ((size_t)h0 * W + w1) * C;
        const float *p10 = base + ((size_t)h1 * W + w0) * C, *p11 = base + ((size_t)h1 * W + w1) * C;
        float o[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 a = ld4(p00 + 4 * q), bq = ld4(p01 + 4 * q), cq = ld4(p10 + 4 * q), d = ld4(p11 + 4 * q);
            o[4 * q + 0] = hl0 * (wl0 * a.x + wl1 * bq.x) + hl1 * (wl0 * cq.x + wl1 * d.x);
            o[4 * q + 1] = hl0 * (wl0 * a.y + wl1 * bq.y) + hl1 * (wl0 * cq.y + wl1 * d.y);
            o[4 * q + 2] = hl0 * (wl0 * a.z + wl1 * bq.z) + hl1 * (wl0 * cq.z + wl1 * d.z);
            o[4 * q + 3] = hl0 * (wl0 * a.w + wl1 * bq.w) + hl1 * (wl0 * cq.w + wl1 * d.w);
        }
        const size_t op = ((size_t)b * Ho + ho) * Wo + wo;
        if (pl) store_planes8(pl + op * cs + c, pl_stride, nsplit, o);
        if (out_f32) {
            *reinterpret_cast<float4 *>(out_f32 + op * C + c) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(out_f32 + op * C + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}
// gather form of the adjoint: dx[b,h,w,:] = sum over the output pixels that read (h,w).  The per-axis weights of the 7
// candidate outputs are computed once per element (7 + 7 evaluations instead of 49 x 2), the products in the same order.
__global__ void __launch_bounds__(kEwThreads) upsample2x_bwd_kernel(int B, int H, int W, int C, float sh, float sw, const float *__restrict__ dout,
                                                                     float *dx) {
    const unsigned lanes = C >> 2;
    const int Ho = 2 * H, Wo = 2 * W;
    const unsigned total = (unsigned)B * H * W * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 4;
        unsigned r = i / lanes;
        const int w = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int h = (int)(r % (unsigned)H);
        const int b = (int)(r / (unsigned)H);
        // candidate outputs: src in (h-1, h+1)  =>  o in ((h-1)*(Ho-1)/(H-1), (h+1)*(Ho-1)/(H-1)); a safe window of 7
        float wh[7], ww[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int ho = 2 * h - 3 + k, wo = 2 * w - 3 + k;
            wh[k] = ww[k] = 0.f;
            if (ho >= 0 && ho <= Ho - 1) {
                int h0, h1; float hl0, hl1;
                up_src_s(ho, H, sh, h0, h1, hl0, hl1);
                wh[k] = (h0 == h ? hl0 : 0.f) + (h1 == h ? hl1 : 0.f);
            }
            if (wo >= 0 && wo <= Wo - 1) {
                int w0, w1; float wl0, wl1;
                up_src_s(wo, W, sw, w0, w1, wl0, wl1);
                ww[k] = (w0 == w ? wl0 : 0.f) + (w1 == w ? wl1 : 0.f);
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (wh[k] == 0.f) continue;
            const int ho = 2 * h - 3 + k;
#pragma unroll
            for (int l = 0; l < 7; ++l) {
                if (ww[l] == 0.f) continue;
                const int wo = 2 * w - 3 + l;
                const float4 d = ld4(dout + (((size_t)b * Ho + ho) * Wo + wo) * C + c);
                const float kk = wh[k] * ww[l];
                acc.x += kk * d.x; acc.y += kk * d.y; acc.z += kk * d.z; acc.w += kk * d.w;
            }
        }
        *reinterpret_cast<float4 *>(dx + (size_t)i * 4) = acc;
    }
}

// C % 8 == 0: 8 channels per thread — the 14 per-axis weight evaluations and the index arithmetic are shared by twice the data and
// every tap is two independent 16-byte loads (the 4-channel kernel ran at 1.9 TB/s on the 302 MB gradients of up_1..3).  Same
// arithmetic order per element.
__global__ void __launch_bounds__(kEwThreads) upsample2x_bwd8_kernel(int B, int H, int W, int C, float sh, float sw, const float *__restrict__ dout,
                                                                      float *dx) {
    const unsigned lanes = C >> 3;
    const int Ho = 2 * H, Wo = 2 * W;
    const unsigned total = (unsigned)B * H * W * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 8;
        unsigned r = i / lanes;
        const int w = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int h = (int)(r % (unsigned)H);
        const int b = (int)(r / (unsigned)H);
        float wh[7], ww[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int ho = 2 * h - 3 + k, wo = 2 * w - 3 + k;
            wh[k] = ww[k] = 0.f;
            if (ho >= 0 && ho <= Ho - 1) {
                int h0, h1; float hl0, hl1;
                up_src_s(ho, H, sh, h0, h1, hl0, hl1);
                wh[k] = (h0 == h ? hl0 : 0.f) + (h1 == h ? hl1 : 0.f);
            }
            if (wo >= 0 && wo <= Wo - 1) {
                int w0, w1; float wl0, wl1;
                up_src_s(wo, W, sw, w0, w1, wl0, wl1);
                ww[k] = (w0 == w ? wl0 : 0.f) + (w1 == w ? wl1 : 0.f);
            }
        }
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (wh[k] == 0.f) continue;
            const int ho = 2 * h - 3 + k;
#pragma unroll
            for (int l = 0; l < 7; ++l) {
                if (ww[l] == 0.f) continue;
                const int wo = 2 * w - 3 + l;
                const float *src = dout + (((size_t)b * Ho + ho) * Wo + wo) * C + c;
                const float4 d0 = ld4(src), d1 = ld4(src + 4);
                const float kk = wh[k] * ww[l];
                a0.x += kk * d0.x; a0.y += kk * d0.y; a0.z += kk * d0.z; a0.w += kk * d0.w;
                a1.x += kk * d1.x; a1.y += kk * d1.y; a1.z += kk * d1.z; a1.w += kk * d1.w;
            }
        }
        *reinterpret_cast<float4 *>(dx + (size_t)i * 8) = a0;
        *reinterpret_cast<float4 *>(dx + (size_t)i * 8 + 4) = a1;
    }
}

// ------------------------------------------------------------------ im2col (strided convs) and its adjoint
// out[b,ho,wo,(r*kw+s)*C + c] = x[b, ho*stride+r-pad, wo*stride+s-pad, c]   (zero outside); x NHWC or NCHW
// One CTA per output pixel (grid-stride): a thread keeps the same patch columns k = tid, tid + 256, ... for every pixel, so
// their (tap, channel) decomposition is hoisted out of the pixel loop and the pixel decomposition is uniform per CTA.
constexpr int kIm2colMaxK = 3;  // K <= 768 columns take the hoisted path
__global__ void __launch_bounds__(kEwThreads) im2col_split_kernel(int B, int H, int W, int C, int kh, int kw, int stride, int pad, int Ho,
                                                                   int Wo, const float *__restrict__ x, int nchw, __nv_bfloat16 *pl,
                                                                   long long pl_stride, int nsplit, int cs) {
    const int K = kh * kw * C;
    const unsigned rows = (unsigned)B * Ho * Wo;
    int kc[kIm2colMaxK], kr[kIm2colMaxK], ks[kIm2colMaxK];
#pragma unroll
    for (int q = 0; q < kIm2colMaxK; ++q) {
        const int k = threadIdx.x + q * kEwThreads;
        const int tap = k / C;
        kc[q] = k % C; kr[q] = tap / kw - pad; ks[q] = tap % kw - pad;
    }
    for (unsigned row = blockIdx.x; row < rows; row += gridDim.x) {
        const int wo = (int)(row % (unsigned)Wo);
        const unsigned t = row / (unsigned)Wo;
        const int ho = (int)(t % (unsigned)Ho), b = (int)(t / (unsigned)Ho);
        const int hb = ho * stride, wb = wo * stride;
        __nv_bfloat16 *orow = pl + (size_t)row * cs;
        if (K <= kIm2colMaxK * kEwThreads) {
#pragma unroll
            for (int q = 0; q < kIm2colMaxK; ++q) {
                const int k = threadIdx.x + q * kEwThreads;
                if (k < K) {
                    const int h = hb + kr[q], w = wb + ks[q], c = kc[q];
                    float v = 0.f;
                    if (h >= 0 && h < H && w >= 0 && w < W) v = nchw ? x[(((size_t)b * C + c) * H + h) * W + w] : x[(((size_t)b * H + h) * W + w) * C + c];
                    store_planes1(orow + k, pl_stride, nsplit, v);
                }
            }
        } else {
            for (int k = threadIdx.x; k < K; k += kEwThreads) {
                const int c = k % C, tap = k / C;
                const int h = hb + tap / kw - pad, w = wb + tap % kw - pad;
                float v = 0.f;
                if (h >= 0 && h < H && w >= 0 && w < W) v = nchw ? x[(((size_t)b * C + c) * H + h) * W + w] : x[(((size_t)b * H + h) * W + w) * C + c];
                store_planes1(orow + k, pl_stride, nsplit, v);
            }
        }
    }
}
// Same map with a thread per 8 consecutive patch columns (cs % 8 == 0): the output [rows][cs] is contiguous, so thread i owns the i-th
// 16 bytes of every plane (coalesced 16-byte stores; the per-column kernel above issues one 2-byte store per plane and element and
// leaves 109 of 256 threads idle on the 7x7x3 stem: 401 us at 0.65 TB/s, profiles/r2_ncu_families_all.txt).  The (tap row, tap
// column, channel) decomposition of every column comes from a shared-memory table built once per CTA; channels-last inputs with
// C % 8 == 0 read their 8 columns as two float4.  Padding columns K..cs-1 are written as zeros.
__global__ void __launch_bounds__(kEwThreads) im2col_split8_kernel(int B, int H, int W, int C, int kh, int kw, int stride, int pad, int Ho, int Wo,
                                                                    const float *__restrict__ x, int nchw, __nv_bfloat16 *pl, long long pl_stride,
                                                                    int nsplit, int cs) {
    extern __shared__ int im2col_tab[];  // [cs]: r << 22 | s << 14 | c, or -1 for padding columns
    const int K = kh * kw * C;
    for (int k = threadIdx.x; k < cs; k += blockDim.x) {
        int e = -1;
        if (k < K) {
            const int tap = k / C;
            e = ((tap / kw) << 22) | ((tap % kw) << 14) | (k - tap * C);
        }
        im2col_tab[k] = e;
    }
    __syncthreads();
    const unsigned chunks = (unsigned)cs >> 3;
    const unsigned total = (unsigned)B * Ho * Wo * chunks;
    const bool vec = !nchw && (C & 7) == 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned row = i / chunks;
        const int k0 = (int)(i - row * chunks) * 8;
        const int wo = (int)(row % (unsigned)Wo);
        const unsigned t = row / (unsigned)Wo;
        const int ho = (int)(t % (unsigned)Ho), b = (int)(t / (unsigned)Ho);
        const int hb = ho * stride - pad, wb = wo * stride - pad;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (vec) {
            const int e = im2col_tab[k0];
            const int h = hb + (e >> 22), w = wb + ((e >> 14) & 0xff);
            if (e >= 0 && h >= 0 && h < H && w >= 0 && w < W) {
                const float *src = x + (((size_t)b * H + h) * W + w) * C + (e & 0x3fff);
                const float4 lo = ld4(src), hi = ld4(src + 4);
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = im2col_tab[k0 + j];
                const int h = hb + (e >> 22), w = wb + ((e >> 14) & 0xff), c = e & 0x3fff;
                if (e >= 0 && h >= 0 && h < H && w >= 0 && w < W)
                    v[j] = nchw ? __ldg(x + (((size_t)b * C + c) * H + h) * W + w) : __ldg(x + (((size_t)b * H + h) * W + w) * C + c);
            }
        }
        store_planes8(pl + (size_t)row * cs + k0, pl_stride, nsplit, v);
    }
}
// dx[b,h,w,c] (+)= sum over (r,s) with (h+pad-r) % stride == 0 ... of dcol[b,ho,wo,(r*kw+s)*C+c]
__global__ void __launch_bounds__(kEwThreads) col2im_kernel(int B, int H, int W, int C, int kh, int kw, int stride, int pad, int Ho, int Wo,
                                                            const float *__restrict__ dcol, float *dx, int accumulate) {
    const int K = kh * kw * C;
    const long long total = (long long)B * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        float acc = 0.f;
        for (int rr = 0; rr < kh; ++rr) {
            const int hn = h + pad - rr;
            if (hn < 0 || hn % stride) continue;
            const int ho = hn / stride;
            if (ho >= Ho) continue;
            for (int ss = 0; ss < kw; ++ss) {
                const int wn = w + pad - ss;
                if (wn < 0 || wn % stride) continue;
                const int wo = wn / stride;
                if (wo >= Wo) continue;
                acc += dcol[(((size_t)b * Ho + ho) * Wo + wo) * K + (rr * kw + ss) * C + c];
            }
        }
        dx[i] = accumulate ? dx[i] + acc : acc;
    }
}

// ------------------------------------------------------------------ stem: BN + ReLU + MaxPool(3,2,1) fused
__global__ void __launch_bounds__(kEwThreads) bn_relu_maxpool_kernel(int B, int H, int W, int C, const float *__restrict__ y, BnP bn,
                                                                      float *out_f32, __nv_bfloat16 *pl, long long pl_stride, int nsplit, int cs,
                                                                      uint8_t *argmax) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, lanes = C >> 2;
    const long long total = (long long)B * Ho * Wo * lanes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 4;
        long long r = i / lanes;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int bi[4] = {0, 0, 0, 0};
        for (int rr = 0; rr < 3; ++rr) {
            const int h = 2 * ho - 1 + rr;
            if (h < 0 || h >= H) continue;
            for (int ss = 0; ss < 3; ++ss) {
                const int w = 2 * wo - 1 + ss;
                if (w < 0 || w >= W) continue;
                float4 u = bn4(ld4(y + (((size_t)b * H + h) * W + w) * C + c), bn, c);
                float v[4] = {fmaxf(u.x, 0.f), fmaxf(u.y, 0.f), fmaxf(u.z, 0.f), fmaxf(u.w, 0.f)};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (v[k] > best[k]) { best[k] = v[k]; bi[k] = rr * 3 + ss; }
            }
        }
        const size_t op = ((size_t)b * Ho + ho) * Wo + wo;
        float4 z = make_float4(best[0], best[1], best[2], best[3]);
        if (out_f32) *reinterpret_cast<float4 *>(out_f32 + op * C + c) = z;
        if (pl) store_planes4(pl + op * cs + c, pl_stride, nsplit, z);
        *reinterpret_cast<uchar4 *>(argmax + op * C + c) = make_uchar4((uint8_t)bi[0], (uint8_t)bi[1], (uint8_t)bi[2], (uint8_t)bi[3]);
    }
}
// g[b,h,w,c] = relu'(bn(y)) * sum over pooled windows whose argmax is (h,w) of (dz1+dz2)      (4 channels per thread)
__global__ void __launch_bounds__(kEwThreads) maxpool_relu_bwd_kernel(int B, int H, int W, int C, const float *__restrict__ y, BnP bn,
                                                                       const float *__restrict__ dz, const float *__restrict__ dz2,
                                                                       const uint8_t *__restrict__ argmax, float *g) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const unsigned lanes = C >> 2;
    const unsigned total = (unsigned)B * H * W * lanes;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % lanes) * 4;
        unsigned r = i / lanes;
        const int w = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int h = (int)(r % (unsigned)H);
        const int b = (int)(r / (unsigned)H);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ho = max(0, (h - 1 + 1) / 2); ho <= min(Ho - 1, (h + 1) / 2); ++ho) {
            const int rr = h - (2 * ho - 1);
            if (rr < 0 || rr > 2) continue;
            for (int wo = max(0, w / 2); wo <= min(Wo - 1, (w + 1) / 2); ++wo) {
                const int ss = w - (2 * wo - 1);
                if (ss < 0 || ss > 2) continue;
                const size_t op = (((size_t)b * Ho + ho) * Wo + wo) * C + c;
                const uchar4 am = *reinterpret_cast<const uchar4 *>(argmax + op);
                float4 d = ld4(dz + op);
                if (dz2) { const float4 d2 = ld4(dz2 + op); d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w; }
                const int sel = rr * 3 + ss;
                if (am.x == sel) acc[0] += d.x;
                if (am.y == sel) acc[1] += d.y;
                if (am.z == sel) acc[2] += d.z;
                if (am.w == sel) acc[3] += d.w;
            }
        }
        const size_t ip = (size_t)i * 4;
        const float4 u = bn4(ld4(y + ip), bn, c);
        *reinterpret_cast<float4 *>(g + ip) = make_float4(u.x > 0.f ? acc[0] : 0.f, u.y > 0.f ? acc[1] : 0.f, u.z > 0.f ? acc[2] : 0.f, u.w > 0.f ? acc[3] : 0.f);
    }
}

// ------------------------------------------------------------------ final head: BN + PReLU only at the chosen pixels
// out[b,n,c] = prelu(bn(y[b, choose[b,n], c]))   (ist_net.py:42-45 gather after modules.py:64-66; row layout)
__global__ void __launch_bounds__(kEwThreads) gather_bn_prelu_kernel(int B, long long HW, int C, int N, const float *__restrict__ y,
                                                                      const long long *__restrict__ choose, BnP bn, const float *prelu_a,
                                                                      float *out) {
    const float a = prelu_a ? *prelu_a : 1.f;
    const long long total = (long long)B * N * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bn_ = i / C;
        const int n = (int)(bn_ % N);
        const int b = (int)(bn_ / N);
        const long long px = choose[(long long)b * N + n];
        float yv = y[((long long)b * HW + px) * C + c];
        float u = bn.mean ? (yv - bn.mean[c]) * bn.invstd[c] * bn.gamma[c] + bn.beta[c] : yv;
        out[i] = u > 0.f ? u : a * u;  // rows: out[b][n][c]
    }
}
// g_dense[b, choose[b,n], c] += dout[b,n,c] * prelu'(u);  slope_ws[c] += dout*u*[u<=0]   (g_dense pre-zeroed)
__global__ void __launch_bounds__(kEwThreads) gather_bn_prelu_bwd_kernel(int B, long long HW, int C, int N, const float *__restrict__ y,
                                                                          const long long *__restrict__ choose, BnP bn, const float *prelu_a,
                                                                          const float *__restrict__ dout, float *g_dense, double *slope_ws) {
    const float a = prelu_a ? *prelu_a : 1.f;
    const long long total = (long long)B * N * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bn_ = i / C;
        const int n = (int)(bn_ % N);
        const int b = (int)(bn_ / N);
        const long long px = choose[(long long)b * N + n];
        const long long src = ((long long)b * HW + px) * C + c;
        float yv = y[src];
        float u = bn.mean ? (yv - bn.mean[c]) * bn.invstd[c] * bn.gamma[c] + bn.beta[c] : yv;
        float d = dout[i];
        atomicAdd(g_dense + src, u > 0.f ? d : a * d);
        if (!(u > 0.f) && slope_ws) atomicAdd(slope_ws + c, (double)(d * u));
    }
}

// ------------------------------------------------------------------ set-abstraction layer 0 on the POINTS instead of the grouped rows
// The first shared-MLP layer of a set-abstraction scale (pointnet2_modules.py:60-66 on QueryAndGroup's output,
// pointnet2_utils.py:335-367) is linear in the grouped row [xyz_j - c_i | f_j]:
//     y0[(i,k), :] = Wx (xyz_j - c_i) + Wf f_j ,   j = idx[i,k]
// so Wf f_j is a GEMM over the N points (u = F Wf^T, 16..32x fewer rows than the M*nsample grouped rows) and the grouped
// tensor never exists: this kernel gathers u[j], adds the 3-term FP32 product with the relative coordinates, writes y0 and
// leaves the per-CTA BatchNorm-statistics partials (same layout as the GEMM epilogue: part[(a*G + g)*C0 + c]).
constexpr int kSaUnroll = 4;
__global__ void __launch_bounds__(kEwThreads, 2) sa_gather_l0_kernel(int N, int M, int ns, int C0, long long rows, const float *__restrict__ xyz,
                                                                     const float *__restrict__ new_xyz, const int32_t *__restrict__ idx,
                                                                     const float *__restrict__ u, const float *__restrict__ w0, int ldw,
                                                                     float *__restrict__ y0, float *part, FinP fin) {
    const int lanes = C0 >> 2;  // <= kEwThreads (launcher)
    const int rows_per_iter = kEwThreads / lanes;
    const int cv = threadIdx.x % lanes, rr = threadIdx.x / lanes;
    const int c = cv * 4;
    float acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    if (rr < rows_per_iter) {
        float wx[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int d = 0; d < 3; ++d) wx[k][d] = w0[(size_t)(c + k) * ldw + d];
        const long long stride = (long long)gridDim.x * rows_per_iter;
        for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < rows; r0 += stride * kSaUnroll) {
            float rel[kSaUnroll][3];
            float4 uv[kSaUnroll];
#pragma unroll
            for (int q = 0; q < kSaUnroll; ++q) {
                const long long r = r0 + q * stride;
                if (r < rows) {
                    const unsigned bj = (unsigned)r / (unsigned)ns;  // rows < 2^31 (launcher)
                    const unsigned b = bj / (unsigned)M;
                    const size_t src = (size_t)b * N + idx[r];
#pragma unroll
                    for (int d = 0; d < 3; ++d) rel[q][d] = __fsub_rn(xyz[src * 3 + d], new_xyz[(size_t)bj * 3 + d]);
                    uv[q] = u ? ld4(u + src * C0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int q = 0; q < kSaUnroll; ++q) {
                const long long r = r0 + q * stride;
                if (r < rows) {
                    float v[4] = {uv[q].x, uv[q].y, uv[q].z, uv[q].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = __fmaf_rn(wx[k][0], rel[q][0], __fmaf_rn(wx[k][1], rel[q][1], __fmaf_rn(wx[k][2], rel[q][2], v[k])));
                        acc[0][k] += v[k];
                        acc[1][k] += v[k] * v[k];
                    }
                    *reinterpret_cast<float4 *>(y0 + r * C0 + c) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
    column_flush<2, false>(acc, lanes, rows_per_iter, rr < rows_per_iter, C0, nullptr, part);
    ticket_finish<2>(fin, part, (int)gridDim.x, C0);
}
// Backward of the above for one scale: dU[j, :] += dy0[(i,k), :] (float atomics, as group_points_grad in the reference) and
// the per-CTA partials of dWx[c][d] = sum_rows dy0[row, c] * (xyz_j - c_i)[d]  (part[(d*G + g)*C0 + c], summed in fixed order).
__global__ void __launch_bounds__(kEwThreads, 2) sa_scatter_l0_kernel(int N, int M, int ns, int C0, long long rows, const float *__restrict__ dy0,
                                                                      const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                      const int32_t *__restrict__ idx, float *dU, float *part, FinP fin) {
    const int lanes = C0 >> 2;
    const int rows_per_iter = kEwThreads / lanes;
    const int cv = threadIdx.x % lanes, rr = threadIdx.x / lanes;
    const int c = cv * 4;
    float acc[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    if (rr < rows_per_iter) {
        const long long stride = (long long)gridDim.x * rows_per_iter;
        for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < rows; r0 += stride * kSaUnroll) {
            float rel[kSaUnroll][3];
            float4 dv[kSaUnroll];
            size_t src[kSaUnroll];
#pragma unroll
            for (int q = 0; q < kSaUnroll; ++q) {
                const long long r = r0 + q * stride;
                if (r < rows) {
                    const unsigned bj = (unsigned)r / (unsigned)ns;
                    const unsigned b = bj / (unsigned)M;
                    src[q] = (size_t)b * N + idx[r];
#pragma unroll
                    for (int d = 0; d < 3; ++d) rel[q][d] = __fsub_rn(xyz[src[q] * 3 + d], new_xyz[(size_t)bj * 3 + d]);
                    dv[q] = ld4(dy0 + r * C0 + c);
                }
            }
#pragma unroll
            for (int q = 0; q < kSaUnroll; ++q) {
                const long long r = r0 + q * stride;
                if (r < rows) {
                    const float g[4] = {dv[q].x, dv[q].y, dv[q].z, dv[q].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int d = 0; d < 3; ++d) acc[d][k] += g[k] * rel[q][d];
                    if (dU) atomicAdd(reinterpret_cast<float4 *>(dU + src[q] * C0 + c), dv[q]);  // one 16-byte reduction (red.global.add.v4.f32)
                }
            }
        }
    }
    column_flush<3, false>(acc, lanes, rows_per_iter, rr < rows_per_iter, C0, nullptr, part);
    ticket_finish<3>(fin, part, (int)gridDim.x, C0);
}

// ------------------------------------------------------------------ PSP priors (modules.py:10-34) on their pooled maps
// The four pyramid levels (sizes 1, 2, 3, 6) of one instance are kept as ONE row block [cells = 1+4+9+36][C]:
//   psp_pool      : nn.AdaptiveAvgPool2d(s) for all sizes in one pass over the feature map (ATen bins: [floor(i*H/s), ceil((i+1)*H/s)) )
//   psp_prior     : sum over the sizes of F.interpolate(t_s, (H,W), bilinear, align_corners=False), written once
// and their adjoints.  Replaces 4 adaptive-pool + 3 up-sampling + 3 add launches (each a full pass over a 512- or 1024-channel
// map, at 32-64 CTAs for the pools) and their autograd counterparts.
struct PspSizes {
    int n, cells;
    int s[4], off[4];
};
__device__ __forceinline__ void psp_cell(const PspSizes &ps, int cell, int &q, int &i, int &j) {
    q = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k < ps.n && cell >= ps.off[k]) q = k;
    const int loc = cell - ps.off[q];
    i = loc / ps.s[q];
    j = loc % ps.s[q];
}
__device__ __forceinline__ int bin_lo(int i, int in, int s) { return (i * in) / s; }
__device__ __forceinline__ int bin_hi(int i, int in, int s) { return ((i + 1) * in + s - 1) / s; }
__global__ void __launch_bounds__(kEwThreads) psp_pool_kernel(int B, int H, int W, int C, PspSizes ps, const float *__restrict__ x, float *pooled) {
    const unsigned lanes = C >> 2;
    const unsigned total = (unsigned)B * ps.cells * lanes;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int c = (int)(t % lanes) * 4;
        const unsigned r = t / lanes;
        const int cell = (int)(r % (unsigned)ps.cells), b = (int)(r / (unsigned)ps.cells);
        int q, i, j;
        psp_cell(ps, cell, q, i, j);
        const int h0 = bin_lo(i, H, ps.s[q]), h1 = bin_hi(i, H, ps.s[q]), w0 = bin_lo(j, W, ps.s[q]), w1 = bin_hi(j, W, ps.s[q]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) {
                const float4 v = ld4(x + (((size_t)b * H + h) * W + w) * C + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
        *reinterpret_cast<float4 *>(pooled + ((size_t)b * ps.cells + cell) * C + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
}
// dx[b,h,w,:] = sum over sizes and bins containing (h,w) of dpooled[b,cell,:] / |bin|
__global__ void __launch_bounds__(kEwThreads) psp_pool_bwd_kernel(int B, int H, int W, int C, PspSizes ps, const float *__restrict__ dp, float *dx) {
    const unsigned lanes = C >> 2;
    const unsigned total = (unsigned)B * H * W * lanes;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int c = (int)(t % lanes) * 4;
        unsigned r = t / lanes;
        const int w = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int h = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < ps.n; ++q) {
            const int sz = ps.s[q];
            for (int i = max(0, (h * sz) / H - 1); i <= min(sz - 1, (h * sz) / H + 1); ++i) {
                const int h0 = bin_lo(i, H, sz), h1 = bin_hi(i, H, sz);
                if (h < h0 || h >= h1) continue;
                for (int j = max(0, (w * sz) / W - 1); j <= min(sz - 1, (w * sz) / W + 1); ++j) {
                    const int w0 = bin_lo(j, W, sz), w1 = bin_hi(j, W, sz);
                    if (w < w0 || w >= w1) continue;
                    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
                    const float4 d = ld4(dp + ((size_t)b * ps.cells + ps.off[q] + i * sz + j) * C + c);
                    acc.x += d.x * inv; acc.y += d.y * inv; acc.z += d.z * inv; acc.w += d.w * inv;
                }
            }
        }
        *reinterpret_cast<float4 *>(dx + (size_t)t * 4) = acc;
    }
}
// ATen area_pixel_compute_source_index(align_corners=False): src = max(scale*(dst+0.5)-0.5, 0), scale = in/out
__device__ __forceinline__ void up_src_half(int o, int in, float scale, int &i0, int &i1, float &l0, float &l1) {
    float s = scale * ((float)o + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}
__global__ void __launch_bounds__(kEwThreads) psp_prior_kernel(int B, int H, int W, int C, PspSizes ps, const float *__restrict__ t, float *prior) {
    const unsigned lanes = C >> 2;
    const unsigned total = (unsigned)B * H * W * lanes;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = (int)(e % lanes) * 4;
        unsigned r = e / lanes;
        const int w = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int h = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < ps.n; ++q) {
            const int sz = ps.s[q];
            const float *base = t + ((size_t)b * ps.cells + ps.off[q]) * C + c;
            int h0, h1, w0, w1;
            float hl0, hl1, wl0, wl1;
            up_src_half(h, sz, (float)sz / (float)H, h0, h1, hl0, hl1);
            up_src_half(w, sz, (float)sz / (float)W, w0, w1, wl0, wl1);
            const float4 a = ld4(base + (size_t)(h0 * sz + w0) * C), bq = ld4(base + (size_t)(h0 * sz + w1) * C);
            const float4 cq = ld4(base + (size_t)(h1 * sz + w0) * C), d = ld4(base + (size_t)(h1 * sz + w1) * C);
            acc.x += hl0 * (wl0 * a.x + wl1 * bq.x) + hl1 * (wl0 * cq.x + wl1 * d.x);
            acc.y += hl0 * (wl0 * a.y + wl1 * bq.y) + hl1 * (wl0 * cq.y + wl1 * d.y);
            acc.z += hl0 * (wl0 * a.z + wl1 * bq.z) + hl1 * (wl0 * cq.z + wl1 * d.z);
            acc.w += hl0 * (wl0 * a.w + wl1 * bq.w) + hl1 * (wl0 * cq.w + wl1 * d.w);
        }
        *reinterpret_cast<float4 *>(prior + (size_t)e * 4) = acc;
    }
}
// dt[b,cell,:] = sum over the pixels that interpolate from the cell of weight * g[b,h,w,:]   (gather form, fixed order)
__global__ void __launch_bounds__(kEwThreads) psp_prior_bwd_kernel(int B, int H, int W, int C, PspSizes ps, const float *__restrict__ g, float *dt) {
    const unsigned lanes = C >> 2;
    const unsigned total = (unsigned)B * ps.cells * lanes;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = (int)(e % lanes) * 4;
        const unsigned r = e / lanes;
        const int cell = (int)(r % (unsigned)ps.cells), b = (int)(r / (unsigned)ps.cells);
        int q, i, j;
        psp_cell(ps, cell, q, i, j);
        const int sz = ps.s[q];
        const float sh = (float)sz / (float)H, sw = (float)sz / (float)W;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h = 0; h < H; ++h) {
            int h0, h1; float hl0, hl1;
            up_src_half(h, sz, sh, h0, h1, hl0, hl1);
            const float wh = (h0 == i ? hl0 : 0.f) + (h1 == i ? hl1 : 0.f);
            if (wh == 0.f) continue;
            for (int w = 0; w < W; ++w) {
                int w0, w1; float wl0, wl1;
                up_src_half(w, sz, sw, w0, w1, wl0, wl1);
                const float ww = (w0 == j ? wl0 : 0.f) + (w1 == j ? wl1 : 0.f);
                if (ww == 0.f) continue;
                const float4 d = ld4(g + (((size_t)b * H + h) * W + w) * C + c);
                const float k = wh * ww;
                acc.x += k * d.x; acc.y += k * d.y; acc.z += k * d.z; acc.w += k * d.w;
            }
        }
        *reinterpret_cast<float4 *>(dt + ((size_t)b * ps.cells + cell) * C + c) = acc;
    }
}

// ATen area_pixel_compute_scale(align_corners=True)
inline float up_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
inline int ew_grid(long long total) {
    long long g = (total + kEwThreads - 1) / kEwThreads;
    long long cap = (long long)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
inline int red_grid(long long P, int C) {
    // >= 4 rows per thread, at most 2 CTAs per SM; every CTA writes one partial vector (no atomics)
    int lanes = C / 4;
    int rows_per_iter = lanes <= kEwThreads ? kEwThreads / lanes : 1;
    long long g = (P + (long long)rows_per_iter * 4 - 1) / ((long long)rows_per_iter * 4);
    long long cap = (long long)kNumSMs * 2;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
// grid of a row-streaming kernel whose CTA covers kEwThreads/lanes rows per iteration, `unroll` iterations in flight
inline int row_grid(long long P, int C, int unroll) {
    const int lanes = C / 4;
    if (lanes > kEwThreads) return ew_grid(P * lanes);
    const long long rows = (long long)(kEwThreads / lanes) * unroll;
    long long g = (P + rows - 1) / rows;
    const long long cap = (long long)kNumSMs * 4;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
inline BnP make_bn(const float *mean, const float *invstd, const float *gamma, const float *beta) { return BnP{mean, invstd, gamma, beta}; }

}  // namespace

#define ST ((cudaStream_t)stream)

int istnet_fin_finalize_launch(const float *part, int G, int C, int nacc, const FinP &f, cudaStream_t st) {
    if (f.kind == 0) return ISTNET_OK;
    if (!part || G <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    const int grid = ceil_div(C, 32);
    if (nacc == 1) fin_finalize_kernel<1><<<grid, kFinThreads, 0, st>>>(part, G, C, f);
    else if (nacc == 2) fin_finalize_kernel<2><<<grid, kFinThreads, 0, st>>>(part, G, C, f);
    else if (nacc == 3) fin_finalize_kernel<3><<<grid, kFinThreads, 0, st>>>(part, G, C, f);
    else return ISTNET_ERR_BAD_ARG;
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// partial rows of both ticket levels (ticket.cuh): [nacc][kMaxPartialRows][C] per-CTA partials + [nacc][kMaxTicketGroups][C] group sums
extern "C" int istnet_reduce_ws_floats(long long P, int C, int nacc) { (void)P; return ISTNET_FIN_ROWS * nacc * C; }

extern "C" int istnet_bn_stats(const float *y, long long P, int C, float *part_ws, float eps, float momentum, float *running_mean,
                               float *running_var, float *mean, float *invstd, long long *num_batches_tracked, void *stream) {
    if (P <= 0 || C <= 0 || (C & 3)) return ISTNET_ERR_BAD_ARG;
    const int G = red_grid(P, C);
    bn_stats_kernel<<<G, kEwThreads, 0, ST>>>(y, P, C, part_ws, FinP{});
    ISTNET_LAUNCH_CHECK();
    bn_finalize_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part_ws, G, P, C, eps, momentum, running_mean, running_var, mean, invstd, num_batches_tracked);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_bn_stats_fin(const float *y, long long P, int C, float *part_ws, const istnet_fin *fin, void *stream) {
    if (P <= 0 || C <= 0 || (C & 3) || !fin || fin->kind != ISTNET_FIN_BN_STATS) return ISTNET_ERR_BAD_ARG;
    FinP f;
    if (!make_fin(fin, part_ws, 2, C, f)) return ISTNET_ERR_BAD_ARG;
    const int G = red_grid(P, C);
    bn_stats_kernel<<<G, kEwThreads, 0, ST>>>(y, P, C, part_ws, fin_in_kernel(C) ? f : FinP{});
    ISTNET_LAUNCH_CHECK();
    if (!fin_in_kernel(C)) return istnet_fin_finalize_launch(part_ws, G, C, 2, f, ST);
    return ISTNET_OK;
}

// second stage of a BN-statistics reduction whose per-CTA partials [2][G][C] were produced elsewhere (the conv epilogue)
extern "C" int istnet_bn_finalize(const float *part, int G, long long P, int C, float eps, float momentum, float *running_mean,
                                  float *running_var, float *mean, float *invstd, long long *num_batches_tracked, void *stream) {
    if (G <= 0 || P <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    bn_finalize_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part, G, P, C, eps, momentum, running_mean, running_var, mean, invstd, num_batches_tracked);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_bn_act_split(const float *y, long long P, int C, long long HW, const float *mean, const float *invstd,
                                   const float *gamma, const float *beta, const float *res, const float *res_mean,
                                   const float *res_invstd, const float *res_gamma, const float *res_beta, int act, const float *prelu_a,
                                   const float *noise, float *out_f32, void *out_planes, long long plane_stride, int nsplit, int cs, int ch_off,
                                   void *stream) {
    if (P <= 0 || C <= 0 || (C & 3) || (out_planes && ((cs & 3) || (ch_off & 3) || nsplit < 1 || nsplit > kMaxPlanes))) return ISTNET_ERR_BAD_ARG;
    ActFwdP p{};
    p.y = y; p.bn = make_bn(mean, invstd, gamma, beta);
    p.res = res; p.res_bn = make_bn(res_mean, res_invstd, res_gamma, res_beta);
    p.act = act; p.prelu_a = prelu_a; p.noise = noise; p.HW = HW > 0 ? HW : 1;
    p.out_f32 = out_f32; p.out_pl = (__nv_bfloat16 *)out_planes; p.pl_stride = plane_stride; p.nsplit = nsplit; p.cs = cs; p.ch_off = ch_off;
    bn_act_split_kernel<<<((C / 4) <= kEwThreads && kEwThreads % (C / 4) == 0) ? row_grid(P, C, kFwdUnroll) : ew_grid(P * (C / 4)), kEwThreads, 0, ST>>>(P, C, p);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_bn_act_bwd(const float *dz, const float *dz2, const float *y, long long P, int C, long long HW, const float *mean,
                                 const float *invstd, const float *gamma, const float *beta, int act, const float *prelu_a,
                                 const void *z_hi, int cs_z, const float *noise, int batch_stats, const uint8_t *argmax, int ns,
                                 float *part_ws, double *ws /*3C*/, void *dy_planes, long long plane_stride, int nsplit, int cs_dy,
                                 float *dy_f32, float *g_out, float *sum_g_f32, float *sum_gx_f32, unsigned *tickets, void *stream) {
    if (P <= 0 || C <= 0 || (C & 3)) return ISTNET_ERR_BAD_ARG;
    if (act == 1 && !z_hi) return ISTNET_ERR_BAD_ARG;
    if (act == 3 && (!argmax || ns <= 0 || !mean || dz2)) return ISTNET_ERR_BAD_ARG;
    if (P > 0x7fffffffLL) return ISTNET_ERR_BAD_ARG;  // group / instance indices are derived with 32-bit divisions
    ActBwdP p{};
    p.dz = dz; p.dz2 = dz2; p.y = y; p.bn = make_bn(mean, invstd, gamma, beta);
    p.act = act; p.prelu_a = prelu_a; p.z_hi = (const __nv_bfloat16 *)z_hi; p.cs_z = cs_z; p.noise = noise; p.HW = HW > 0 ? HW : 1;
    p.batch_stats = batch_stats;
    p.argmax = argmax; p.ns = ns;
    const int G = red_grid(P, C);
    FinP f{};
    if (tickets) {  // the reduce kernel's last CTA finishes the sums itself (ticket.cuh)
        istnet_fin h{};
        h.kind = ISTNET_FIN_BN_BWD; h.tickets = tickets; h.sum_f64 = ws; h.sum_f32 = sum_g_f32; h.sum2_f32 = sum_gx_f32;
        if (!make_fin(&h, part_ws, 3, C, f)) return ISTNET_ERR_BAD_ARG;
    }
    bn_bwd_reduce_kernel<<<G, kEwThreads, 0, ST>>>(P, C, p, part_ws, fin_in_kernel(C) ? f : FinP{});
    ISTNET_LAUNCH_CHECK();
    if (tickets && !fin_in_kernel(C)) {
        int e = istnet_fin_finalize_launch(part_ws, G, C, 3, f, ST);
        if (e) return e;
    }
    if (!tickets) {
        bwd_finalize_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part_ws, G, C, ws, sum_g_f32, sum_gx_f32);
        ISTNET_LAUNCH_CHECK();
    }
    bn_bwd_apply_kernel<<<row_grid(P, C, kBwdUnroll), kEwThreads, 0, ST>>>(P, C, p, ws, (__nv_bfloat16 *)dy_planes, plane_stride, nsplit, cs_dy,
                                                                            dy_f32, g_out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// BN-backward sums from the statistics epilogue of the data-gradient GEMM above (conv_gemm stat_y): part = [sum g | sum g*y] per CTA
__global__ void __launch_bounds__(kFinThreads) bwd_finalize_gy_kernel(const float *__restrict__ part, int G, int C, const float *__restrict__ mean,
                                                                      const float *__restrict__ invstd, double *ws, float *sum_g_f32,
                                                                      float *sum_gx_f32) {
    double t[2];
    sum_partials<2>(part, G, C, t);
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x >= 32 || c >= C) return;
    const double sgx = (double)invstd[c] * (t[1] - (double)mean[c] * t[0]);  // sum g*xhat, xhat = (y - mean)*invstd
    ws[c] = t[0];
    ws[C + c] = sgx;
    ws[2 * C + c] = 0.0;
    if (sum_g_f32) sum_g_f32[c] = (float)t[0];
    if (sum_gx_f32) sum_gx_f32[c] = (float)sgx;
}
extern "C" int istnet_bn_bwd_finalize_gy(const float *part, int G, int C, const float *mean, const float *invstd, double *ws, float *sum_g_f32,
                                         float *sum_gx_f32, void *stream) {
    if (!part || G <= 0 || C <= 0 || !mean || !invstd || !ws) return ISTNET_ERR_BAD_ARG;
    bwd_finalize_gy_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part, G, C, mean, invstd, ws, sum_g_f32, sum_gx_f32);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
// apply pass of istnet_bn_act_bwd alone, for sums (ws) produced elsewhere
extern "C" int istnet_bn_bwd_apply(const float *dz, const float *y, long long P, int C, const float *mean, const float *invstd, const float *gamma,
                                   const float *beta, int act, const void *z_hi, int cs_z, const double *ws, void *dy_planes,
                                   long long plane_stride, int nsplit, int cs_dy, float *dy_f32, void *stream) {
    if (P <= 0 || P > 0x7fffffffLL || C <= 0 || (C & 3) || !mean || !ws || (act != 0 && act != 1) || (act == 1 && !z_hi)) return ISTNET_ERR_BAD_ARG;
    ActBwdP p{};
    p.dz = dz; p.y = y; p.bn = make_bn(mean, invstd, gamma, beta);
    p.act = act; p.z_hi = (const __nv_bfloat16 *)z_hi; p.cs_z = cs_z; p.HW = 1; p.batch_stats = 1;
    bn_bwd_apply_kernel<<<row_grid(P, C, kBwdUnroll), kEwThreads, 0, ST>>>(P, C, p, ws, (__nv_bfloat16 *)dy_planes, plane_stride, nsplit, cs_dy, dy_f32,
                                                                            nullptr);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_split(const float *x, long long P, int C, long long HW, int nchw, void *planes, long long plane_stride, int nsplit,
                            int cs, int ch_off, void *stream) {
    if (P <= 0 || C <= 0 || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if (!nchw && (C & 3) == 0 && (cs & 3) == 0 && (ch_off & 3) == 0 && (plane_stride & 3) == 0 && P * (C / 4) <= 0x7fffffffLL &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(planes) & 7) == 0)
        split_rows4_kernel<<<ew_grid(P * (C / 4)), kEwThreads, 0, ST>>>((unsigned)P, C, x, (__nv_bfloat16 *)planes, plane_stride, nsplit, cs, ch_off);
    else
        split_kernel<<<ew_grid(P * C), kEwThreads, 0, ST>>>(P, C, x, HW > 0 ? HW : 1, nchw, (__nv_bfloat16 *)planes, plane_stride, nsplit, cs, ch_off);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_prep_weight(const float *w, int Cout, int Cin, int kh, int kw, int transpose, int im2col, void *planes, long long plane_stride,
                                  int nsplit, int cs, void *stream) {
    if (Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    prep_weight_kernel<<<ew_grid((long long)Cout * Cin * kh * kw), kEwThreads, 0, ST>>>(Cout, Cin, kh, kw, w, transpose, im2col, (__nv_bfloat16 *)planes,
                                                                                       plane_stride, nsplit, cs);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

// column sums of a tensor that only exists as operand planes: part[g*C + c] = sum over CTA g's rows of (p0 + p1 + ...)[r][c]
__global__ void __launch_bounds__(kEwThreads) colsum_planes_kernel(const __nv_bfloat16 *__restrict__ pl, long long pl_stride, int nsplit, long long P,
                                                                   int C, int cs, float *part) {
    const int lanes = C >> 2;
    if (lanes > kEwThreads) {
        column_reduce<1, false>(P, C, nullptr, part, [&](long long r, int c, float (*acc)[4]) {
            for (int i = 0; i < nsplit; ++i) {
                const float4 v = bf4_to_f4(pl + (size_t)i * pl_stride + r * cs + c);
                acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
            }
        });
        return;
    }
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    const int rows_per_iter = kEwThreads / lanes;
    const int c = (threadIdx.x % lanes) * 4, rr = threadIdx.x / lanes;
    if (rr < rows_per_iter) {
        const long long stride = (long long)gridDim.x * rows_per_iter;
        for (long long r0 = (long long)blockIdx.x * rows_per_iter + rr; r0 < P; r0 += stride * 4) {
            uint2 raw[4][kMaxPlanes];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < kMaxPlanes; ++i)
                    if (i < nsplit && r0 + u * stride < P) raw[u][i] = *reinterpret_cast<const uint2 *>(pl + (size_t)i * pl_stride + (r0 + u * stride) * cs + c);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < kMaxPlanes; ++i)
                    if (i < nsplit && r0 + u * stride < P) {
                        acc[0][0] += __uint_as_float(raw[u][i].x << 16); acc[0][1] += __uint_as_float(raw[u][i].x & 0xffff0000u);
                        acc[0][2] += __uint_as_float(raw[u][i].y << 16); acc[0][3] += __uint_as_float(raw[u][i].y & 0xffff0000u);
                    }
        }
    }
    column_flush<1, false>(acc, lanes, rows_per_iter, rr < rows_per_iter, C, nullptr, part);
}
__global__ void __launch_bounds__(kFinThreads) colsum_finalize_kernel(const float *__restrict__ part, int G, int C, double *ws, float *out_f32 = nullptr) {
    double t[1];
    sum_partials<1>(part, G, C, t);
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (threadIdx.x < 32 && c < C) {
        if (ws) ws[c] = t[0];
        if (out_f32) out_f32[c] = (float)t[0];
    }
}
extern "C" int istnet_colsum_finalize(const float *part, int G, int C, double *ws, float *out_f32, void *stream) {
    if (!part || G <= 0 || C <= 0) return ISTNET_ERR_BAD_ARG;
    colsum_finalize_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part, G, C, ws, out_f32);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_colsum_planes(const void *planes, long long plane_stride, int nsplit, long long P, int C, int cs, float *part_ws,
                                    double *ws, void *stream) {
    if (!planes || P <= 0 || C <= 0 || (C & 3) || (cs & 3) || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    const int G = red_grid(P, C);
    colsum_planes_kernel<<<G, kEwThreads, 0, ST>>>((const __nv_bfloat16 *)planes, plane_stride, nsplit, P, C, cs, part_ws);
    ISTNET_LAUNCH_CHECK();
    colsum_finalize_kernel<<<ceil_div(C, 32), kFinThreads, 0, ST>>>(part_ws, G, C, ws);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
__global__ void marker_kernel(unsigned long long *stamps, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[slot] = t;
}
extern "C" int istnet_sa_gather_l0(int B, int N, int M, int ns, int C0, const float *xyz, const float *new_xyz, const int32_t *idx, const float *u,
                                   const float *w0, int ldw, float *y0, float *stat_part, int *grid_out, const istnet_fin *fin, void *stream) {
    const long long rows = (long long)B * M * ns;
    if (B <= 0 || N <= 0 || M <= 0 || ns <= 0 || C0 <= 0 || (C0 & 3) || C0 / 4 > kEwThreads || ldw < 3 || rows > 0x7fffffffLL) return ISTNET_ERR_BAD_ARG;
    const int G = red_grid(rows, C0);  // <= 296: the size callers give the statistics scratch
    FinP f;
    if ((fin && fin->kind != ISTNET_FIN_NONE && fin->kind != ISTNET_FIN_BN_STATS) || !make_fin(fin, stat_part, 2, C0, f)) return ISTNET_ERR_BAD_ARG;
    sa_gather_l0_kernel<<<G, kEwThreads, 0, ST>>>(N, M, ns, C0, rows, xyz, new_xyz, idx, u, w0, ldw, y0, stat_part, fin_in_kernel(C0) ? f : FinP{});
    ISTNET_LAUNCH_CHECK();
    if (!fin_in_kernel(C0)) {
        int e = istnet_fin_finalize_launch(stat_part, G, C0, 2, f, ST);
        if (e) return e;
    }
    if (grid_out) *grid_out = G;
    return ISTNET_OK;
}
extern "C" int istnet_sa_scatter_l0(int B, int N, int M, int ns, int C0, const float *dy0, const float *xyz, const float *new_xyz,
                                    const int32_t *idx, float *dU, float *part_ws, double *ws, unsigned *tickets, void *stream) {
    const long long rows = (long long)B * M * ns;
    if (B <= 0 || N <= 0 || M <= 0 || ns <= 0 || C0 <= 0 || (C0 & 3) || C0 / 4 > kEwThreads || rows > 0x7fffffffLL) return ISTNET_ERR_BAD_ARG;
    if (dU) ISTNET_CUDA_TRY(cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)B * N * C0, ST));
    const int G = red_grid(rows, C0);
    FinP f{};
    if (tickets) {
        istnet_fin h{};
        h.kind = ISTNET_FIN_BN_BWD; h.tickets = tickets; h.sum_f64 = ws;
        if (!make_fin(&h, part_ws, 3, C0, f)) return ISTNET_ERR_BAD_ARG;
    }
    sa_scatter_l0_kernel<<<G, kEwThreads, 0, ST>>>(N, M, ns, C0, rows, dy0, xyz, new_xyz, idx, dU, part_ws, fin_in_kernel(C0) ? f : FinP{});
    ISTNET_LAUNCH_CHECK();
    if (tickets && !fin_in_kernel(C0)) {
        int e = istnet_fin_finalize_launch(part_ws, G, C0, 3, f, ST);
        if (e) return e;
    }
    if (!tickets) {
        bwd_finalize_kernel<<<ceil_div(C0, 32), kFinThreads, 0, ST>>>(part_ws, G, C0, ws);
        ISTNET_LAUNCH_CHECK();
    }
    return ISTNET_OK;
}
static int make_psp_sizes(int s0, int s1, int s2, int s3, PspSizes &ps) {
    const int v[4] = {s0, s1, s2, s3};
    ps.n = 0; ps.cells = 0;
    for (int k = 0; k < 4; ++k) {
        ps.s[k] = 1; ps.off[k] = 0;
    }
    for (int k = 0; k < 4 && v[k] > 0; ++k) {
        ps.s[k] = v[k]; ps.off[k] = ps.cells; ps.cells += v[k] * v[k]; ps.n = k + 1;
    }
    return ps.n;
}
#define PSP_ARGS_OK(B, H, W, C, ps) ((B) > 0 && (H) > 0 && (W) > 0 && (C) > 0 && !((C) & 3) && (ps).n > 0 && \
                                      (long long)(B) * (H) * (W) * ((C) / 4) <= 0x7fffffffLL && (long long)(B) * (ps).cells * ((C) / 4) <= 0x7fffffffLL)
extern "C" int istnet_psp_pool(const float *x, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *pooled, void *stream) {
    PspSizes ps;
    make_psp_sizes(s0, s1, s2, s3, ps);
    if (!PSP_ARGS_OK(B, H, W, C, ps)) return ISTNET_ERR_BAD_ARG;
    psp_pool_kernel<<<ew_grid((long long)B * ps.cells * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, ps, x, pooled);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_psp_pool_bwd(const float *dpooled, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *dx, void *stream) {
    PspSizes ps;
    make_psp_sizes(s0, s1, s2, s3, ps);
    if (!PSP_ARGS_OK(B, H, W, C, ps)) return ISTNET_ERR_BAD_ARG;
    psp_pool_bwd_kernel<<<ew_grid((long long)B * H * W * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, ps, dpooled, dx);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_psp_prior(const float *t, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *prior, void *stream) {
    PspSizes ps;
    make_psp_sizes(s0, s1, s2, s3, ps);
    if (!PSP_ARGS_OK(B, H, W, C, ps)) return ISTNET_ERR_BAD_ARG;
    psp_prior_kernel<<<ew_grid((long long)B * H * W * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, ps, t, prior);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_psp_prior_bwd(const float *g, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *dt, void *stream) {
    PspSizes ps;
    make_psp_sizes(s0, s1, s2, s3, ps);
    if (!PSP_ARGS_OK(B, H, W, C, ps)) return ISTNET_ERR_BAD_ARG;
    psp_prior_bwd_kernel<<<ew_grid((long long)B * ps.cells * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, ps, g, dt);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_marker(unsigned long long *stamps, int slot, void *stream) {
    if (!stamps || slot < 0) return ISTNET_ERR_BAD_ARG;
    marker_kernel<<<1, 1, 0, ST>>>(stamps, slot);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_prep_weight_pair(const float *w, int Cout, int Cin, int kh, int kw, int im2col, void *planes_fwd, long long stride_fwd,
                                       int nsplit_fwd, int cs_fwd, void *planes_bwd, long long stride_bwd, int nsplit_bwd, int cs_bwd, void *stream) {
    if (Cout <= 0 || Cin <= 0 || nsplit_fwd < 1 || nsplit_fwd > kMaxPlanes || nsplit_bwd < 1 || nsplit_bwd > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    prep_weight_pair_kernel<<<ew_grid((long long)Cout * Cin * kh * kw), kEwThreads, 0, ST>>>(Cout, Cin, kh, kw, w, im2col, (__nv_bfloat16 *)planes_fwd,
                                                                                              stride_fwd, nsplit_fwd, cs_fwd, (__nv_bfloat16 *)planes_bwd,
                                                                                              stride_bwd, nsplit_bwd, cs_bwd);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_colsum(const float *x, long long P, int C, double *ws, void *stream) {
    if (P <= 0 || C <= 0 || (C & 3)) return ISTNET_ERR_BAD_ARG;
    ISTNET_CUDA_TRY(cudaMemsetAsync(ws, 0, sizeof(double) * C, ST));
    colsum_kernel<<<red_grid(P, C), kEwThreads, 0, ST>>>(x, P, C, ws);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_upsample2x_split(const float *x, int B, int H, int W, int C, void *planes, long long plane_stride, int nsplit, int cs,
                                       float *out_f32, void *stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if ((long long)B * 4 * H * W * (C / 4) > 0x7fffffffLL) return ISTNET_ERR_UNSUPPORTED;
    if ((C & 7) == 0 && (cs & 7) == 0 && (plane_stride & 7) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0)
        upsample2x_split8_kernel<<<ew_grid((long long)B * 4 * H * W * (C / 8)), kEwThreads, 0, ST>>>(B, H, W, C, up_scale(H, 2 * H), up_scale(W, 2 * W), x,
                                                                                                    (__nv_bfloat16 *)planes, plane_stride, nsplit, cs, out_f32);
    else
        upsample2x_split_kernel<<<ew_grid((long long)B * 4 * H * W * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, up_scale(H, 2 * H), up_scale(W, 2 * W), x,
                                                                                                   (__nv_bfloat16 *)planes, plane_stride, nsplit, cs, out_f32);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_upsample2x_bwd(const float *dout, int B, int H, int W, int C, float *dx, void *stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return ISTNET_ERR_BAD_ARG;
    if ((long long)B * H * W * (C / 4) > 0x7fffffffLL) return ISTNET_ERR_UNSUPPORTED;
    if ((C & 7) == 0)
        upsample2x_bwd8_kernel<<<ew_grid((long long)B * H * W * (C / 8)), kEwThreads, 0, ST>>>(B, H, W, C, up_scale(H, 2 * H), up_scale(W, 2 * W), dout, dx);
    else
        upsample2x_bwd_kernel<<<ew_grid((long long)B * H * W * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, up_scale(H, 2 * H), up_scale(W, 2 * W), dout, dx);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_im2col_split(const float *x, int nchw, int B, int H, int W, int C, int kh, int kw, int stride, int pad, void *planes,
                                   long long plane_stride, int nsplit, int cs, void *stream) {
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    if (B <= 0 || Ho <= 0 || Wo <= 0 || cs < kh * kw * C || nsplit < 1 || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    if ((long long)B * Ho * Wo > 0x7fffffffLL) return ISTNET_ERR_UNSUPPORTED;
    const long long chunks = (long long)B * Ho * Wo * (cs / 8);
    if ((cs & 7) == 0 && (plane_stride & 7) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        cs * 4 <= 48 * 1024 && kh < 256 && kw < 256 && C < (1 << 14) && chunks <= 0x7fffffffLL)
        im2col_split8_kernel<<<ew_grid(chunks), kEwThreads, (size_t)cs * sizeof(int), ST>>>(B, H, W, C, kh, kw, stride, pad, Ho, Wo, x, nchw,
                                                                                          (__nv_bfloat16 *)planes, plane_stride, nsplit, cs);
    else
        im2col_split_kernel<<<ew_grid((long long)B * Ho * Wo * kEwThreads), kEwThreads, 0, ST>>>(B, H, W, C, kh, kw, stride, pad, Ho, Wo, x, nchw,
                                                                                                 (__nv_bfloat16 *)planes, plane_stride, nsplit, cs);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_col2im(const float *dcol, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float *dx, int accumulate,
                             void *stream) {
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    if (B <= 0 || Ho <= 0 || Wo <= 0) return ISTNET_ERR_BAD_ARG;
    col2im_kernel<<<ew_grid((long long)B * H * W * C), kEwThreads, 0, ST>>>(B, H, W, C, kh, kw, stride, pad, Ho, Wo, dcol, dx, accumulate);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_bn_relu_maxpool(const float *y, int B, int H, int W, int C, const float *mean, const float *invstd, const float *gamma,
                                      const float *beta, float *out_f32, void *planes, long long plane_stride, int nsplit, int cs,
                                      uint8_t *argmax, void *stream) {
    if (B <= 0 || (C & 3) || nsplit > kMaxPlanes) return ISTNET_ERR_BAD_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    bn_relu_maxpool_kernel<<<ew_grid((long long)B * Ho * Wo * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, y, make_bn(mean, invstd, gamma, beta),
                                                                                            out_f32, (__nv_bfloat16 *)planes, plane_stride, nsplit, cs, argmax);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_maxpool_relu_bwd(const float *y, int B, int H, int W, int C, const float *mean, const float *invstd, const float *gamma,
                                       const float *beta, const float *dz, const float *dz2, const uint8_t *argmax, float *g, void *stream) {
    if (B <= 0 || (C & 3) || (long long)B * H * W * (C / 4) > 0x7fffffffLL) return ISTNET_ERR_BAD_ARG;
    maxpool_relu_bwd_kernel<<<ew_grid((long long)B * H * W * (C / 4)), kEwThreads, 0, ST>>>(B, H, W, C, y, make_bn(mean, invstd, gamma, beta), dz, dz2,
                                                                                     argmax, g);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_gather_bn_prelu(const float *y, int B, long long HW, int C, int N, const long long *choose, const float *mean,
                                      const float *invstd, const float *gamma, const float *beta, const float *prelu_a, float *out,
                                      void *stream) {
    if (B <= 0 || N <= 0) return ISTNET_ERR_BAD_ARG;
    gather_bn_prelu_kernel<<<ew_grid((long long)B * N * C), kEwThreads, 0, ST>>>(B, HW, C, N, y, choose, make_bn(mean, invstd, gamma, beta),
                                                                                prelu_a, out);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_gather_bn_prelu_bwd(const float *y, int B, long long HW, int C, int N, const long long *choose, const float *mean,
                                          const float *invstd, const float *gamma, const float *beta, const float *prelu_a,
                                          const float *dout, float *g_dense, double *slope_ws, void *stream) {
    if (B <= 0 || N <= 0) return ISTNET_ERR_BAD_ARG;
    ISTNET_CUDA_TRY(cudaMemsetAsync(g_dense, 0, sizeof(float) * (size_t)B * HW * C, ST));
    if (slope_ws) ISTNET_CUDA_TRY(cudaMemsetAsync(slope_ws, 0, sizeof(double) * C, ST));
    gather_bn_prelu_bwd_kernel<<<ew_grid((long long)B * N * C), kEwThreads, 0, ST>>>(B, HW, C, N, y, choose, make_bn(mean, invstd, gamma, beta),
                                                                                    prelu_a, dout, g_dense, slope_ws);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

#ifdef ISTNET_TICKET_DEBUG
// debug build only (-DISTNET_TICKET_DEBUG): %globaltimer stamps of the ticket tail of the last reduction kernel of THIS file
extern "C" int istnet_ticket_debug(unsigned long long *out8) {
    return (int)cudaMemcpyFromSymbol(out8, g_ticket_dbg, sizeof(unsigned long long) * 8);
}
#endif

// table: device array of n entries laid out as `struct PrepEntry` (see istnet_prep_entry in include/istnet_b200.h); total_blocks =
// sum over the entries of ceil(co*ci*kh*kw / ISTNET_PREP_CHUNK)
extern "C" int istnet_prep_weight_batch(const void *table, int n, int total_blocks, void *stream) {
    if (!table || n <= 0 || total_blocks <= 0) return ISTNET_ERR_BAD_ARG;
    static_assert(sizeof(PrepEntry) == sizeof(istnet_prep_entry), "PrepEntry must match the C ABI struct");
    static_assert(kPrepChunk == ISTNET_PREP_CHUNK, "chunk size is part of the ABI (callers compute block0)");
    prep_weight_batch_kernel<<<total_blocks, kEwThreads, 0, ST>>>((const PrepEntry *)table, n);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
