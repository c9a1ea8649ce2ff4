// "Last CTA finishes the job": fixed-order reduction of per-CTA partial sums inside the kernel that produced them.
//
// Round 1 finished every per-channel reduction (train-mode BatchNorm statistics, BatchNorm-backward sums, bias gradients) with
// a separate one-warp-per-channel kernel: ~90 + ~120 launches per training step, each a 5-15 us node on the dependent chain
// conv -> finalize -> BN-apply.  Here every CTA writes its partial vector as before (no floating-point atomics: the summation
// order stays fixed, results are bitwise reproducible), takes a ticket, and the last CTA to arrive sums the partials:
//   level 1: the last CTA of each group of kTicketGroup CTAs sums its group's partials      -> part2[a][group][c]
//   level 2: the last group to finish sums the <= 19 group partials in double              -> caller's epilogue
// so no CTA reads more than 16 + 19 partial vectors (a single CTA summing 296 x 2C floats out of L2 would take as long as
// the launch it replaces).  Tickets are self-resetting: they must be zero on entry and are zero again on exit, so CUDA-graph
// replays and repeated calls need no memset.
#pragma once
#include "common.cuh"

constexpr int kTicketGroup = 16;
constexpr int kMaxPartialRows = 296;                                          // every producer launches <= 2 CTAs per SM
constexpr int kMaxTicketGroups = (kMaxPartialRows + kTicketGroup - 1) / kTicketGroup;  // 19

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
__device__ bool ticket_reduce(const float *part, int cta, int G, int C, unsigned *tickets, float *part2) {
    __shared__ int s_last;
    const int G2 = (G + kTicketGroup - 1) / kTicketGroup;
    const int grp = cta / kTicketGroup;
    const int g0 = grp * kTicketGroup;
    const int gsz = min(kTicketGroup, G - g0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[grp], 1u) == (unsigned)(gsz - 1));
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
        const int a = i / C, c = i - a * C;
        float v[kTicketGroup];
#pragma unroll
        for (int g = 0; g < kTicketGroup; ++g) v[g] = (g < gsz) ? __ldcg(part + ((size_t)a * G + g0 + g) * C + c) : 0.f;
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < kTicketGroup; ++g) s += (double)v[g];
        part2[((size_t)a * G2 + grp) * C + c] = (float)s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        tickets[grp] = 0u;
        s_last = (atomicAdd(&tickets[kMaxTicketGroups], 1u) == (unsigned)(G2 - 1));
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x == 0) tickets[kMaxTicketGroups] = 0u;
    return true;
}
// sum over the group partials (double, fixed order) of quantity a, channel c — for the CTA ticket_reduce returned true in
__device__ __forceinline__ double ticket_total(const float *part2, int G, int C, int a, int c) {
    const int G2 = (G + kTicketGroup - 1) / kTicketGroup;
    float v[kMaxTicketGroups];
#pragma unroll
    for (int g = 0; g < kMaxTicketGroups; ++g) v[g] = (g < G2) ? __ldcg(part2 + ((size_t)a * G2 + g) * C + c) : 0.f;
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < kMaxTicketGroups; ++g) s += (double)v[g];
    return s;
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
        if (f.kind == 1) {
            if (NACC >= 2) bn_fin_channel(f, c, ticket_total(f.part2, G, C, 0, c), ticket_total(f.part2, G, C, NACC >= 2 ? 1 : 0, c), n_old);
        } else if (f.kind == 2) {
            const double t = ticket_total(f.part2, G, C, 0, c);
            if (f.sum_f64) f.sum_f64[c] = t;
            if (f.sum_f32) f.sum_f32[c] = (float)t;
        } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                const double t = ticket_total(f.part2, G, C, a, c);
                if (f.sum_f64) f.sum_f64[(size_t)a * C + c] = t;
                if (a == 0 && f.sum_f32) f.sum_f32[c] = (float)t;
                if (a == 1 && f.sum2_f32) f.sum2_f32[c] = (float)t;
            }
        }
    }
    if (f.kind == 1 && f.num_batches_tracked && threadIdx.x == 0) *f.num_batches_tracked = n_old + 1;
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
static_assert(ISTNET_FIN_TICKETS == kMaxTicketGroups + 1, "ticket counters per reduction");
static_assert(ISTNET_FIN_ROWS == kMaxPartialRows + kMaxTicketGroups, "partial rows per quantity");
