// Adam over FLAT parameter / gradient buckets: the optimizer step of the reference's training loop (utils/solver.py:41-46:
// torch.optim.Adam with default betas / eps; utils/solver.py:98-99: loss.backward(); optimizer.step()) as ONE HBM pass per
// bucket that a captured CUDA graph can contain: the learning rate (rewritten every iteration by CyclicLR, solver.py:45-46,88-89)
// and the step count are DEVICE scalars, and the 1/world_size of the data-parallel gradient mean is folded in (the NCCL
// all-reduce sums).  Arithmetic follows torch.optim.Adam (amsgrad=False, maximize=False):
//   g = grad*scale + wd*p;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// 7 x 4 bytes per parameter: 26.3 M parameters = 0.74 GB -> ~0.12 ms at the measured 6.4 TB/s.
#include <math.h>

#include "common.cuh"

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) adam_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                             float *__restrict__ v, long long n4, const float *__restrict__ lr_ptr, double beta1d,
                                                             double beta2d, float eps, float wd, float gscale, const long long *__restrict__ step_ptr) {
    // step_ptr holds the number of steps ALREADY taken (incremented by adam_tick_kernel after all buckets of this step)
    const double t = (double)(*step_ptr + 1);
    const float lr = *lr_ptr;
    const float bc1 = (float)(1.0 - pow(beta1d, t));
    const float bc2s = (float)sqrt(1.0 - pow(beta2d, t));
    const float step_size = lr / bc1;
    // torch: exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2) with the scalars formed in double
    const float beta2 = (float)beta2d, omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4 *>(p)[i];
        const float4 gv = reinterpret_cast<const float4 *>(g)[i];
        float4 mv = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gg = ga[k] * gscale + wd * pa[k];
            ma[k] = ma[k] + (gg - ma[k]) * omb1;
            va[k] = va[k] * beta2 + omb2 * gg * gg;
            const float denom = sqrtf(va[k]) / bc2s + eps;
            pa[k] = pa[k] - step_size * (ma[k] / denom);
        }
        reinterpret_cast<float4 *>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
        reinterpret_cast<float4 *>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        reinterpret_cast<float4 *>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    }
}
__global__ void adam_tick_kernel(long long *step_ptr) { *step_ptr += 1; }
}  // namespace

// One bucket.  n must be a multiple of 4 and the four buffers 16-byte aligned (flat buckets are padded).  `step` counts completed
// steps and is NOT modified here: call istnet_adam_tick once after the last bucket of a step.
extern "C" int istnet_adam_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, const float *lr_dev,
                                double beta1, double beta2, float eps, float weight_decay, float grad_scale, const long long *step_dev,
                                void *stream) {
    if (n <= 0) return ISTNET_OK;
    if ((n & 3) || !param || !grad || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev) return ISTNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return ISTNET_ERR_BAD_ARG;
    const long long n4 = n / 4;
    long long g = (n4 + kThreads - 1) / kThreads;
    const long long cap = (long long)kNumSMs * 8;
    if (g > cap) g = cap;
    adam_flat_kernel<<<(unsigned)g, kThreads, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, lr_dev, beta1, beta2, eps,
                                                                           weight_decay, grad_scale, step_dev);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
extern "C" int istnet_adam_tick(long long *step_dev, void *stream) {
    if (!step_dev) return ISTNET_ERR_BAD_ARG;
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}
