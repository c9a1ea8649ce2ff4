// "Last CTA finishes the job": fixed-order reduction of per-CTA partial sums inside the kernel that produced them.
//
// Round 1 finished every per-channel reduction (train-mode BatchNorm statistics, BatchNorm-backward sums, bias gradients) with
// a separate one-warp-per-channel kernel: ~90 + ~120 launches per training step, each a 5-15 us node on the dependent chain
// conv -> finalize -> BN-apply.  Here every CTA writes its partial vector as before (no floating-point atomics: the summation
// order stays fixed, results are bitwise reproducible), takes a ticket, and the last CTA to arrive sums the partials:
//   level 1: the last CTA of each group of kTicketGroup CTAs sums its group's partials      -> part2[a][group][c]
//   level 2: the last group to finish sums the <= 37 group partials in double              -> caller's epilogue
// so no CTA reads more than 16 + 37 partial vectors (a single CTA summing 296 x 2C floats out of L2 would take as long as
// the launch it replaces).  Tickets are self-resetting: they must be zero on entry and are zero again on exit, so CUDA-graph
// replays and repeated calls need no memset.
#pragma once
#include "common.cuh"

#ifdef ISTNET_TICKET_DEBUG
static __device__ unsigned long long g_ticket_dbg[8];
__device__ __forceinline__ void ticket_stamp(int i) {
    if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_ticket_dbg[i] = t; }
}
#else
__device__ __forceinline__ void ticket_stamp(int) {}
#endif
constexpr int kTicketGroup = 16;
constexpr int kTicketMaxC = 128;  // wider reductions finish with a separate (parallel) finalize launch: the serial tail would cost more than it saves
constexpr int kMaxPartialRows = 592;                                          // every producer launches <= 4 CTAs per SM
constexpr int kMaxTicketGroups = (kMaxPartialRows + kTicketGroup - 1) / kTicketGroup;  // 37

// Reference-side descriptor (include/istnet_b200.h: istnet_fin) as the kernels see it.
struct FinP {
    int kind;            // 0: none, 1: BatchNorm forward statistics, 2: column sums, 3: BatchNorm backward sums
    unsigned *tickets;   // kMaxTicketGroups + 1 counters, zero on entry / exit
    float *part2;        // [nacc][kMaxTicketGroups][C] level-2 scratch
    long long P;         // rows the statistics run over (kind 1)
    float eps;
    const float *momentum;  // DEVICE scalar, read when the kernel runs (BNMomentumScheduler under graph replay); < 0: cumulative average
    float *running_mean, *running_var, *mean, *invstd;
    long long *num_batches_tracked;
    double *sum_f64;     // kind 2: [C]; kind 3: [3C] = sum g | sum g*xhat | PReLU slope
    float *sum_f32;      // kind 2: [C] (nullable); kind 3: sum g (nullable)
    float *sum2_f32;     // kind 3: sum g*xhat (nullable)
};

// All threads of the CTA call this after the CTA's partial row part[(a*G + cta)*C + c] is written (cta = this CTA's index among
// the G CTAs that share the reduction: blockIdx.x for a whole grid, or an offset into it when one launch carries several
// independent reductions).  Returns true (for every thread) in exactly one of the G CTAs — the last one to finish — after
// part2[(a*G2 + g2)*C + c], g2 < G2 = ceil(G / kTicketGroup), holds the group sums of all CTAs.
template <int NACC>
__device__ __noinline__ bool ticket_reduce(const float *part, int cta, int G, int C, unsigned *tickets, float *part2) {
    __shared__ int s_last;
    const int G2 = (G + kTicketGroup - 1) / kTicketGroup;
    const int grp = cta / kTicketGroup;
    const int g0 = grp * kTicketGroup;
    const int gsz = min(kTicketGroup, G - g0);
    ticket_stamp(0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[grp], 1u) == (unsigned)(gsz - 1));
    __syncthreads();
    if (!s_last) return false;
    ticket_stamp(1);
    __threadfence();
    // COMPACT code on purpose (rolled loops, four loads in flight): this tail runs once per launch on an SM whose instruction
    // cache has never seen it, and a fully unrolled version (40-50 KB of straight-line SASS) spent 10-30 us fetching its own
    // instructions (measured with %globaltimer stamps, profiles/r2_ticket_tail.txt).
    for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
        const int a = i / C, c = i - a * C;
        const float *src = part + ((size_t)a * G + g0) * C + c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int g = 0;
#pragma unroll 1
        for (; g + 4 <= gsz; g += 4) {
            const float v0 = __ldcg(src + (size_t)g * C), v1 = __ldcg(src + (size_t)(g + 1) * C);
            const float v2 = __ldcg(src + (size_t)(g + 2) * C), v3 = __ldcg(src + (size_t)(g + 3) * C);
            s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        }
#pragma unroll 1
        for (; g < gsz; ++g) s0 += __ldcg(src + (size_t)g * C);
        part2[((size_t)a * G2 + grp) * C + c] = (s0 + s1) + (s2 + s3);
    }
    ticket_stamp(2);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        tickets[grp] = 0u;
        s_last = (atomicAdd(&tickets[kMaxTicketGroups], 1u) == (unsigned)(G2 - 1));
    }
    __syncthreads();
    if (!s_last) return false;
    ticket_stamp(3);
    __threadfence();
    if (threadIdx.x == 0) tickets[kMaxTicketGroups] = 0u;
    ticket_stamp(4);
    return true;
}
// sum over the group partials (double, fixed order) of quantity a, channel c — for the CTA ticket_reduce returned true in.
// Rolled loop, four loads in flight (see the note on code size in ticket_reduce).
static __device__ __noinline__ double ticket_total(const float *part2, int G, int C, int a, int c) {
    const int G2 = (G + kTicketGroup - 1) / kTicketGroup;
    const float *src = part2 + (size_t)a * G2 * C + c;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int g = 0;
#pragma unroll 1
    for (; g + 4 <= G2; g += 4) {
        const float v0 = __ldcg(src + (size_t)g * C), v1 = __ldcg(src + (size_t)(g + 1) * C);
        const float v2 = __ldcg(src + (size_t)(g + 2) * C), v3 = __ldcg(src + (size_t)(g + 3) * C);
        s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
    }
#pragma unroll 1
    for (; g < G2; ++g) s0 += (double)__ldcg(src + (size_t)g * C);
    return (s0 + s1) + (s2 + s3);
}
template <int NACC>
__device__ __forceinline__ void ticket_totals(const float *part2, int G, int C, int c, double (&t)[NACC]) {
#pragma unroll 1
    for (int a = 0; a < NACC; ++a) t[a] = ticket_total(part2, G, C, a, c);
}

// nn.BatchNorm2d training statistics from the totals (pytorch_utils.py:53-71, resnet.py:129): biased variance for the
// normalisation, unbiased for running_var, momentum read from device memory.
__device__ __forceinline__ void bn_fin_channel(const FinP &f, int c, double sum, double sumsq, long long n_old) {
    const double m = sum / (double)f.P;
    double var = sumsq / (double)f.P - m * m;
    if (var < 0.0) var = 0.0;
    f.mean[c] = (float)m;
    f.invstd[c] = (float)(1.0 / sqrt(var + (double)f.eps));
    if (f.running_mean) {
        double mom = (double)*f.momentum;
        if (mom < 0.0) mom = 1.0 / (double)(n_old + 1);  // momentum=None: cumulative moving average (torch.nn.BatchNorm2d)
        const double unbiased = f.P > 1 ? var * (double)f.P / (double)(f.P - 1) : var;
        f.running_mean[c] = (float)((1.0 - mom) * f.running_mean[c] + mom * m);
        f.running_var[c] = (float)((1.0 - mom) * f.running_var[c] + mom * unbiased);
    }
}

// The complete tail of a producer whose partials are laid out part[(a*G + cta)*C + c].  kind 1 expects a = {sum, sum of
// squares}; kind 2 a = {sum}; kind 3 a = {sum g, sum g*xhat, slope}.  Call with all threads of every CTA.
template <int NACC>
__device__ void ticket_finish(const FinP &f, const float *part, int G, int C, int cta = -1) {
    if (f.kind == 0 || f.tickets == nullptr) return;
    if (!ticket_reduce<NACC>(part, cta < 0 ? (int)blockIdx.x : cta, G, C, f.tickets, f.part2)) return;
    __shared__ long long s_nold;
    if (threadIdx.x == 0) s_nold = (f.kind == 1 && f.num_batches_tracked) ? *f.num_batches_tracked : 0;
    __syncthreads();
    const long long n_old = s_nold;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double t[NACC];
        ticket_totals<NACC>(f.part2, G, C, c, t);
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
    if (f.kind == 1 && f.num_batches_tracked && threadIdx.x == 0) *f.num_batches_tracked = n_old + 1;
    ticket_stamp(5);
}

// host: istnet_fin (C ABI) -> FinP.  `part` is the kernel's partial-sum scratch of nacc * ISTNET_FIN_ROWS * C floats: rows
// [0, kMaxPartialRows) per quantity hold the per-CTA partials, the level-2 group sums live behind them.  Returns false when the
// descriptor is inconsistent.
static inline bool make_fin(const istnet_fin *h, float *part, int nacc, int C, FinP &f) {
    f = FinP{};
    if (h == nullptr || h->kind == ISTNET_FIN_NONE) return true;
    if (h->kind < 1 || h->kind > 3 || !h->tickets || !part) return false;
    if (h->kind == ISTNET_FIN_BN_STATS && (!h->mean || !h->invstd || h->P <= 0 || (h->running_mean && (!h->running_var || !h->momentum)))) return false;
    if (h->kind == ISTNET_FIN_COLSUM && !h->sum_f64 && !h->sum_f32) return false;
    if (h->kind == ISTNET_FIN_BN_BWD && !h->sum_f64) return false;
    f.kind = h->kind; f.tickets = h->tickets; f.part2 = part + (size_t)nacc * kMaxPartialRows * C; f.P = h->P; f.eps = h->eps;
    f.momentum = h->momentum; f.running_mean = h->running_mean; f.running_var = h->running_var; f.mean = h->mean; f.invstd = h->invstd;
    f.num_batches_tracked = h->num_batches_tracked; f.sum_f64 = h->sum_f64; f.sum_f32 = h->sum_f32; f.sum2_f32 = h->sum2_f32;
    return true;
}
// Reductions wider than kTicketMaxC channels are finished by a separate, channel-parallel launch (defined in elementwise.cu).
static inline bool fin_in_kernel(int C) { return C <= kTicketMaxC; }
int istnet_fin_finalize_launch(const float *part, int G, int C, int nacc, const FinP &f, cudaStream_t st);
static_assert(ISTNET_FIN_TICKETS == kMaxTicketGroups + 1, "ticket counters per reduction");
static_assert(ISTNET_FIN_ROWS == kMaxPartialRows + kMaxTicketGroups, "partial rows per quantity");
