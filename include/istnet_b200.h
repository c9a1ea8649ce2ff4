/*
 * istnet_b200 — C ABI of the B200 (sm_100a) implementation of IST-Net's per-instance hot path.
 *
 * Every entry point takes raw DEVICE pointers, plain sizes and a CUDA stream (as void*, i.e. cudaStream_t),
 * launches asynchronously on that stream and returns 0 on success or a non-zero cudaError_t / negative
 * argument-error code.  The library keeps no global device state, never allocates device memory, never
 * synchronises and never exits the process (the reference's CUDA_CHECK_ERRORS exit(-1)s:
 * model/pointnet2/_ext_src/include/cuda_utils.h:35-44).  All tensors are contiguous; float = FP32,
 * indices = int32, exactly as the reference's CHECK_* macros demand (include/utils.h:10-30).
 *
 * Section 1 mirrors, one for one, the nine functions of the reference pybind module `pointnet2._ext`
 * (model/pointnet2/_ext_src/src/bindings.cpp:11-24).  The reference-side binding a maintainer would add is
 * shown in INTEGRATION.md.  Sections 2+ are the fused entry points used by the istnet_b200 modules.
 */
#ifndef ISTNET_B200_H
#define ISTNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISTNET_OK 0
#define ISTNET_ERR_BAD_ARG (-1)
#define ISTNET_ERR_UNSUPPORTED (-2)

/* Library / build information: returns the compiled arch (100 for sm_100a). */
int istnet_version(void);
/* Human readable text for a status returned by any function below. */
const char *istnet_strerror(int status);

/* ------------------------------------------------------------------------------------------------------
 * 1. The nine `pointnet2._ext` operators
 * ---------------------------------------------------------------------------------------------------- */

/* bindings.cpp:13 furthest_point_sampling  (sampling.cpp:70-91, sampling_gpu.cu:74-234)
 * xyz[b,n,3] -> idx[b,m].  Bit-exact with the reference, including its block-reduction tie-break
 * (bit-reversed thread id, then lowest k).  No scratch buffer is needed (the reference's `temp` lives in
 * registers here). */
int istnet_furthest_point_sampling(int b, int n, int m, const float *xyz, int32_t *idx, void *stream);

/* bindings.cpp:11 gather_points  (sampling.cpp:20-43, sampling_gpu.cu:13-35): out[b,c,m] = points[b,c,idx[b,m]] */
int istnet_gather_points(int b, int c, int n, int m, const float *points, const int32_t *idx, float *out, void *stream);

/* bindings.cpp:12 gather_points_grad (sampling.cpp:45-69, sampling_gpu.cu:39-62). grad_points[b,c,n] must be zeroed by the caller. */
int istnet_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, float *grad_points, void *stream);

/* bindings.cpp:19 ball_query (ball_query.cpp:13-37, ball_query_gpu.cu:14-59).  Centroids first, as in the
 * reference.  idx[b,m,nsample] is fully written (rows without any hit are zero, like the reference's zeros-init). */
int istnet_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int32_t *idx, void *stream);

/* bindings.cpp:21 group_points (group_points.cpp:17-40, group_points_gpu.cu:13-44): out[b,c,m,ns] = points[b,c,idx[b,m,ns]] */
int istnet_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx, float *out, void *stream);

/* bindings.cpp:22 group_points_grad (group_points.cpp:42-65, group_points_gpu.cu:48-80). grad_points[b,c,n] must be zeroed by the caller. */
int istnet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int32_t *idx, float *grad_points, void *stream);

/* bindings.cpp:15 three_nn (interpolate.cpp:19-45, interpolate_gpu.cu:14-73): dist2[b,n,3] ascending, idx[b,n,3] */
int istnet_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx, void *stream);

/* bindings.cpp:16 three_interpolate (interpolate.cpp:47-75, interpolate_gpu.cu:77-117) points[b,c,m] -> out[b,c,n] */
int istnet_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx, const float *weight, float *out, void *stream);

/* bindings.cpp:17 three_interpolate_grad (interpolate.cpp:76-104, interpolate_gpu.cu:121-159). grad_points[b,c,m] must be zeroed by the caller. */
int istnet_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, const float *weight, float *grad_points, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 2. Fused point-cloud entry points (PointNet2MSG, model/modules.py:244-327)
 * ---------------------------------------------------------------------------------------------------- */

/* FPS + gather for ALL set-abstraction levels of one extractor in ONE launch
 * (pointnet2_modules.py:47-58 executed for SA_modules[0..nlevels-1], modules.py:249-299).
 * xyz[b,n,3]; level l samples npoint[l] points out of the previous level's output.
 * idx_out[l]  -> int32 [b,npoint[l]]   (indices into level l's input cloud)
 * xyz_out[l]  -> float [b,npoint[l],3] (the gathered centroids, == gather_points on the transposed cloud)
 * nlevels <= 4. */
int istnet_fps_chain(int b, int n, int nlevels, const int *npoint, const float *xyz, int32_t *const *idx_out, float *const *xyz_out, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 2b. In-kernel completion of per-channel reductions ("last CTA finishes": csrc/ticket.cuh)
 *
 * Kernels that leave per-CTA partial sums (BatchNorm statistics in a GEMM epilogue, BatchNorm-backward sums, bias gradients)
 * finish the reduction themselves when given this descriptor: the last CTA to arrive sums the partials in a fixed order (no
 * floating-point atomics, bitwise reproducible) and applies the epilogue of `kind`.  Replaces the separate finalize launches
 * (nn.BatchNorm2d's statistics kernels in the reference: resnet.py:129, modules.py:43,65, pytorch_utils.py:53-71).
 * `tickets` points at ISTNET_FIN_TICKETS zero-initialised uint32 counters that are zero again when the kernel exits (graph
 * replays need no memset); the partial-sum scratch given to the kernel must hold ISTNET_FIN_ROWS rows per quantity.
 * `momentum` is a DEVICE scalar read when the kernel runs, so a BNMomentumScheduler update (utils/scheduler.py:277-303) takes
 * effect on a replayed CUDA graph; a negative value selects nn.BatchNorm2d(momentum=None)'s cumulative average.
 * ---------------------------------------------------------------------------------------------------- */
#define ISTNET_FIN_NONE 0
#define ISTNET_FIN_BN_STATS 1 /* partials {sum, sum^2} -> mean, invstd, running stats, num_batches_tracked */
#define ISTNET_FIN_COLSUM 2   /* partials {sum} -> sum_f64 / sum_f32 [C] */
#define ISTNET_FIN_BN_BWD 3   /* partials {sum g, sum g*xhat, slope} -> sum_f64 [3C], sum_f32 = sum g, sum2_f32 = sum g*xhat */
#define ISTNET_FIN_TICKETS 38
#define ISTNET_FIN_ROWS (592 + 37)
typedef struct istnet_fin {
    int kind;
    unsigned *tickets;
    long long P;
    float eps;
    const float *momentum;
    float *running_mean, *running_var, *mean, *invstd;
    long long *num_batches_tracked;
    double *sum_f64;
    float *sum_f32, *sum2_f32;
} istnet_fin;

/* ------------------------------------------------------------------------------------------------------
 * 3. Dense contractions on tcgen05 tensor cores (image branch model/resnet.py + model/modules.py:10-81,
 *    per-point MLPs model/ist_net.py:125-332, SharedMLP 1x1 convolutions pytorch_utils.py:25-206)
 *
 * An FP32 tensor x is carried as `nsplit` bf16 "operand planes" stored plane-major (planes[i] is a whole bf16
 * tensor, `plane_stride` ELEMENTS after planes[i-1]): p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1).
 * A product is evaluated as the sum of all plane products a_i*b_j with i + j < nsplit, FP32-accumulated in tensor
 * memory: nsplit = 2 -> 3 MMAs (~3e-6 per product), nsplit = 3 -> 6 MMAs (~1e-8, below FP32 rounding; default).
 * Activations are channels-last: act[b][h][w][c] with a channel stride `*_cs` (elements, multiple of 8).
 * ---------------------------------------------------------------------------------------------------- */

/* Stride-1 "same" convolution / GEMM:  out[b,h,w,n] = bias[n] + sum_{r,s,c} act[b,h+r-kh/2,w+s-kw/2,c] * wgt[r*kw+s][n][c]
 * (replaces nn.Conv2d / nn.Conv1d(k=1) / nn.Linear call sites: resnet.py:34-35, modules.py:17-25,41-44,64-66,
 * ist_net.py:130-160).  wgt planes: bf16 [nsplit][kh*kw][Cout][wgt_cs].  Any of out_f32 / out_planes may be null.
 * (box_w, box_h): pixel tile; box_w*box_h must divide 128 (images: 8x8 -> 2 images per tile; row matrices: 128x1).
 * stat_part (optional, >= 2*ISTNET_FIN_ROWS*Cout floats): the epilogue also accumulates the per-channel sum / sum of squares of the
 * output (train-mode BatchNorm statistics) per CTA; *grid_out receives the number of CTAs G.  With `fin` (nullable; kinds
 * ISTNET_FIN_BN_STATS, ISTNET_FIN_COLSUM) the kernel's last CTA finishes the reduction itself; otherwise finish with
 * istnet_bn_finalize / istnet_colsum_finalize.
 * mask_hi (nullable, bf16 [B,H,W,mask_cs]): the output is zeroed where mask <= 0 before statistics / stores — used by the data
 * gradient of a layer fed by a bias+ReLU layer (nn.Conv1d + nn.ReLU stacks, ist_net.py:130-160): the epilogue then produces that
 * layer's dy operand planes and, through stat_part, its bias gradient (autograd's threshold_backward + sum in the reference).
 * stat_y (nullable, FP32 [B,H,W,stat_y_cs], needs stat_part): the second statistic becomes sum(out * stat_y) instead of sum(out^2) —
 * with mask_hi this is the reduction of the BatchNorm backward of a conv+BN+ReLU layer below (finish with
 * istnet_bn_bwd_finalize_gy + istnet_bn_bwd_apply).
 * bias_group > 0: bias is a table [ceil(P / bias_group)][Cout] and output pixel p receives row p / bias_group — the per-instance form of
 * the estimators' global feature (ist_net.py:172-173,257-258,325-326: conv([f | mean(f).expand]) = W_a f + (W_b mean(f) + b)). */
int istnet_conv_gemm(const void *act_planes, long long act_plane_stride, int B, int H, int W, int Cin, int act_cs,
                     const void *wgt_planes, long long wgt_plane_stride, int Cout, int wgt_cs, int kh, int kw, int nsplit,
                     const float *bias, int relu, float *out_f32, int out_cs, void *out_planes, long long out_plane_stride,
                     int nsplit_out, int split_cs, int box_w, int box_h, float *stat_part, int *grid_out, const void *mask_hi, int mask_cs,
                     const float *stat_y, int stat_y_cs, const istnet_fin *fin, int bias_group, void *stream);

/* Weight gradient of the layer above (cuDNN wgrad in the reference):
 *   grad_w[co][ci][r][s] = sum_{b,h,w} dy[b,h,w,co] * x[b,h+r-kh/2,w+s-kw/2,ci]       (PyTorch weight layout, FP32)
 * The pixel sum is split over `ksplit` CTAs per tile (istnet_wgrad_ksplit gives the recommended value);
 * partial_ws must hold ksplit*kh*kw*Cout*Cin floats.  Deterministic (fixed-order reduction of the partials).
 * (box_w, box_h): pixel tile with box_w*box_h dividing 64 (images: 8x8; row matrices: 64x1). */
int istnet_wgrad_ksplit(int B, int H, int W, int Cout, int Cin, int kh, int kw, int nsplit);
int istnet_conv_wgrad(const void *dy_planes, long long dy_plane_stride, int dy_cs, const void *x_planes, long long x_plane_stride,
                      int x_cs, int nsplit, int B, int H, int W, int Cout, int Cin, int kh, int kw, float *partial_ws, int ksplit,
                      float *grad_w, int box_w, int box_h, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 4. HBM-bound passes between contractions (channels-last FP32 [P pixels][C channels], C % 4 == 0)
 * ---------------------------------------------------------------------------------------------------- */

/* nn.BatchNorm2d training statistics (resnet.py:129, modules.py:43,65, pytorch_utils.py:53-71): per-channel mean and
 * 1/sqrt(biased var + eps) over P rows; running stats updated in place with `momentum` (unbiased variance) when the
 * pointers are non-null.  part_ws: istnet_reduce_ws_floats(P, C, 2) floats of scratch (per-CTA partial sums, summed in a
 * fixed order by a second tiny kernel: no atomics, deterministic). */
int istnet_reduce_ws_floats(long long P, int C, int nacc); /* floats of partial-sum scratch for a per-channel reduction (both ticket levels) */
int istnet_bn_stats(const float *y, long long P, int C, float *part_ws, float eps, float momentum, float *running_mean,
                    float *running_var, float *mean, float *invstd, long long *num_batches_tracked, void *stream);
/* same, one launch: the reduction kernel's last CTA finishes (fin->kind = ISTNET_FIN_BN_STATS; part_ws as above) */
int istnet_bn_stats_fin(const float *y, long long P, int C, float *part_ws, const istnet_fin *fin, void *stream);

/* num_batches_tracked (nullable): nn.BatchNorm2d's int64 step counter, incremented by the same launch */
int istnet_bn_finalize(const float *part, int G, long long P, int C, float eps, float momentum, float *running_mean, float *running_var,
                       float *mean, float *invstd, long long *num_batches_tracked, void *stream);

/* z = noise[b,c] * act( bn(y) + bn_res(res) )  written as FP32 and/or as bf16 operand planes (channel stride cs, offset ch_off).
 * mean==NULL: no BN.  res==NULL: no residual; res_mean==NULL: raw residual.  act: 0 none, 1 ReLU, 2 PReLU(*prelu_a).
 * noise: Dropout2d scale [B][C] or NULL, b = pixel / HW.  (BasicBlock resnet.py:50-66, PSPUpsample modules.py:37-48) */
int istnet_bn_act_split(const float *y, long long P, int C, long long HW, const float *mean, const float *invstd, const float *gamma,
                        const float *beta, const float *res, const float *res_mean, const float *res_invstd, const float *res_gamma,
                        const float *res_beta, int act, const float *prelu_a, const float *noise, float *out_f32, void *out_planes,
                        long long plane_stride, int nsplit, int cs, int ch_off, void *stream);

/* Backward of the unit above (part_ws: istnet_reduce_ws_floats(P, C, 3) floats): g = (dz + dz2) * noise * act'(u);  ws[0:C] = sum g, ws[C:2C] = sum g*xhat, ws[2C:3C] = PReLU
 * slope partials;  dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))  (batch_stats=1; gamma*invstd*g with running
 * statistics, batch_stats=0; g without BN) written as bf16 operand planes and/or FP32;
 * g_out (optional) receives g (the residual branch's gradient).  ReLU masks come from plane 0 (z_hi) of the saved forward output.
 * tickets (nullable, ISTNET_FIN_TICKETS zeroed counters): the reduce kernel's last CTA finishes the sums (two launches instead of three).
 * act = 3: BN + ReLU + max over `ns` consecutive rows (the last SharedMLP layer of a set-abstraction scale): dz is [P/ns][C]
 * and is routed to the arg-max row of each group (argmax from istnet_bn_relu_maxrows). */
int istnet_bn_act_bwd(const float *dz, const float *dz2, const float *y, long long P, int C, long long HW, const float *mean,
                      const float *invstd, const float *gamma, const float *beta, int act, const float *prelu_a, const void *z_hi,
                      int cs_z, const float *noise, int batch_stats, const uint8_t *argmax, int ns, float *part_ws, double *ws,
                      void *dy_planes, long long plane_stride, int nsplit, int cs_dy, float *dy_f32, float *g_out, float *sum_g_f32,
                      float *sum_gx_f32, unsigned *tickets, void *stream);

/* BatchNorm backward of a conv+BN+ReLU layer whose reduction rode in the data-gradient GEMM above it (istnet_conv_gemm with mask_hi and
 * stat_y): finalize_gy turns the per-CTA partials [sum g | sum g*y] into ws = [sum g | sum g*xhat | 0] (+ FP32 copies = the BN bias /
 * weight gradients); bn_bwd_apply is the apply pass of istnet_bn_act_bwd alone (dz = the masked gradient g written by that GEMM). */
int istnet_bn_bwd_finalize_gy(const float *part, int G, int C, const float *mean, const float *invstd, double *ws, float *sum_g_f32,
                              float *sum_gx_f32, void *stream);
int istnet_bn_bwd_apply(const float *dz, const float *y, long long P, int C, const float *mean, const float *invstd, const float *gamma,
                        const float *beta, int act, const void *z_hi, int cs_z, const double *ws, void *dy_planes, long long plane_stride,
                        int nsplit, int cs_dy, float *dy_f32, void *stream);

/* FP32 [P][C] (or NCHW with HW pixels per image when nchw != 0) -> bf16 operand planes [nsplit][P][cs] at channel offset ch_off */
int istnet_split(const float *x, long long P, int C, long long HW, int nchw, void *planes, long long plane_stride, int nsplit, int cs,
                 int ch_off, void *stream);
/* Weight re-layout + split in one pass: PyTorch [Cout][Cin][kh][kw] FP32 -> operand planes
 *   transpose=0: [nsplit][tap][Cout][cs] (forward);  transpose=1: [nsplit][flipped tap][Cin][cs] (data gradient);
 *   im2col=1: the strided-conv form, one tap with K = (r*kw+s)*Cin + ci  ([Cout][cs] or, transposed, [K][cs]). */
int istnet_prep_weight(const float *w, int Cout, int Cin, int kh, int kw, int transpose, int im2col, void *planes, long long plane_stride,
                       int nsplit, int cs, void *stream);
/* First shared-MLP layer of a set-abstraction scale evaluated on the POINTS (pointnet2_modules.py:60-66 applied to
 * QueryAndGroup's [xyz_j - c_i | f_j], pointnet2_utils.py:335-367): y0[(b,i,k), :] = u[b, idx[b,i,k], :] + Wx (xyz_j - c_i),
 * with u = F Wf^T a GEMM over the N points (u may be null: level 1 has no features) and Wx = w0[:, 0:3] (row stride ldw).
 * Replaces group_points + the grouped 1x1 convolution; also leaves the train-mode BatchNorm statistics partials of y0 in
 * stat_part (>= 2*296*C0 floats, layout of istnet_bn_finalize; *grid_out = number of partial rows G). */
int istnet_sa_gather_l0(int B, int N, int M, int ns, int C0, const float *xyz, const float *new_xyz, const int32_t *idx, const float *u,
                        const float *w0, int ldw, float *y0, float *stat_part, int *grid_out, const istnet_fin *fin, void *stream);
/* Backward of istnet_sa_gather_l0: dU[b, idx, :] += dy0[row, :] (zeroed here; float atomics like group_points_grad,
 * group_points_gpu.cu:48-69; may be null) and ws[d*C0 + c] = sum_rows dy0[row, c] * (xyz_j - c_i)[d] (double, fixed summation
 * order; ws holds 3*C0).  part_ws: istnet_reduce_ws_floats(rows, C0, 3) floats.  tickets (nullable): ISTNET_FIN_TICKETS zeroed
 * counters -> the kernel's last CTA finishes the dWx sums (no finalize launch). */
int istnet_sa_scatter_l0(int B, int N, int M, int ns, int C0, const float *dy0, const float *xyz, const float *new_xyz, const int32_t *idx,
                         float *dU, float *part_ws, double *ws, unsigned *tickets, void *stream);
/* ws[c] = sum_p sum_i planes[i][p][c] (double): column sums of a tensor held only as bf16 operand planes.  With the Gram matrix
 * X^T X (istnet_conv_wgrad of the planes against themselves) this gives the train-mode BatchNorm statistics of the head's 1x1
 * convolution (modules.py:64-66: Conv2d(64,128,1) -> BatchNorm2d -> PReLU) without materialising its 192x192x128 output: the
 * head is only read at the `choose`d pixels (ist_net.py:42-45).  part_ws: istnet_reduce_ws_floats(P, C, 1) floats. */
int istnet_colsum_planes(const void *planes, long long plane_stride, int nsplit, long long P, int C, int cs, float *part_ws, double *ws,
                         void *stream);
/* PSP priors (PSPModule, modules.py:10-34) on their pooled maps.  The pyramid levels s0..s3 (0 = unused; the model uses 1,2,3,6) of an
 * instance form one row block [cells = sum s^2][C], level k starting at row sum_{q<k} s_q^2, cell (i,j) at row i*s+j.
 *   psp_pool      pooled[b,cell,:] = nn.AdaptiveAvgPool2d(s)(x)[b,:,i,j]        x: [B,H,W,C] channels-last (modules.py:17)
 *   psp_prior     prior[b,h,w,:]   = sum_s F.interpolate(t_s, (H,W), 'bilinear', align_corners=False)[b,:,h,w]   (modules.py:30)
 * and their adjoints (gather form, fixed summation order).  t is the level map AFTER both 1x1 convolutions (stage conv and the
 * bottleneck's slice for that level): bilinear up-sampling commutes with a 1x1 convolution. */
int istnet_psp_pool(const float *x, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *pooled, void *stream);
int istnet_psp_pool_bwd(const float *dpooled, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *dx, void *stream);
int istnet_psp_prior(const float *t, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *prior, void *stream);
int istnet_psp_prior_bwd(const float *g, int B, int H, int W, int C, int s0, int s1, int s2, int s3, float *dt, void *stream);
/* Timeline marker: one thread stores %globaltimer (ns) into stamps[slot] in stream order.  A node like any other inside a
 * captured CUDA graph, so phase boundaries of the real (graph-replayed, multi-stream) step can be read back (tools/timeline.py);
 * the reference has no counterpart (its solver times whole iterations with time.time(), utils/solver.py:153). */
int istnet_marker(unsigned long long *stamps, int slot, void *stream);
/* istnet_prep_weight for transpose = 0 (planes_fwd) and transpose = 1 (planes_bwd) in one launch */
int istnet_prep_weight_pair(const float *w, int Cout, int Cin, int kh, int kw, int im2col, void *planes_fwd, long long stride_fwd,
                            int nsplit_fwd, int cs_fwd, void *planes_bwd, long long stride_bwd, int nsplit_bwd, int cs_bwd, void *stream);
/* istnet_prep_weight_pair for MANY weights in one launch: `table` is a DEVICE array of n istnet_prep_entry; entry i is handled by the
 * CTAs [block0_i, block0_i + ceil(co*ci*kh*kw / ISTNET_PREP_CHUNK)), block0 ascending from 0; planes_bwd may be null (forward operand
 * only).  The weights of a model change once per optimizer step, so all of them are re-laid in one launch at the start of a step
 * instead of one launch in front of every convolution. */
#define ISTNET_PREP_CHUNK 2048
typedef struct istnet_prep_entry {
    const float *w;
    void *planes_fwd, *planes_bwd;
    long long stride_fwd, stride_bwd;
    int Cout, Cin, kh, kw, im2col, nsplit_fwd, cs_fwd, nsplit_bwd, cs_bwd, block0;
} istnet_prep_entry;
int istnet_prep_weight_batch(const void *table, int n, int total_blocks, void *stream);
/* out[c] = sum_g part[g*C + c]: column sums from the per-CTA partials of a statistics epilogue (first of its two quantities) */
int istnet_colsum_finalize(const float *part, int G, int C, double *ws, float *out_f32, void *stream);
/* ws[c] = sum_p x[p][c] (double; bias gradients) */
int istnet_colsum(const float *x, long long P, int C, double *ws, void *stream);

/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) (modules.py:41) fused with the operand split; and its adjoint */
int istnet_upsample2x_split(const float *x, int B, int H, int W, int C, void *planes, long long plane_stride, int nsplit, int cs,
                            float *out_f32, void *stream);
int istnet_upsample2x_bwd(const float *dout, int B, int H, int W, int C, float *dx, void *stream);

/* im2col for the strided convolutions (conv1 7x7/2 resnet.py:127, layer2.0 3x3/2 + 1x1/2 resnet.py:153-180) and its adjoint */
int istnet_im2col_split(const float *x, int nchw, int B, int H, int W, int C, int kh, int kw, int stride, int pad, void *planes,
                        long long plane_stride, int nsplit, int cs, void *stream);
int istnet_col2im(const float *dcol, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float *dx, int accumulate,
                  void *stream);

/* stem: BN + ReLU + MaxPool2d(3,2,1) in one pass (resnet.py:183-186) and the matching backward to the conv1 output */
int istnet_bn_relu_maxpool(const float *y, int B, int H, int W, int C, const float *mean, const float *invstd, const float *gamma,
                           const float *beta, float *out_f32, void *planes, long long plane_stride, int nsplit, int cs, uint8_t *argmax,
                           void *stream);
int istnet_maxpool_relu_bwd(const float *y, int B, int H, int W, int C, const float *mean, const float *invstd, const float *gamma,
                            const float *beta, const float *dz, const float *dz2, const uint8_t *argmax, float *g, void *stream);

/* head: BN + PReLU evaluated only at the `choose`d pixels (ist_net.py:42-45), output as rows out[b][n][c] */
int istnet_gather_bn_prelu(const float *y, int B, long long HW, int C, int N, const long long *choose, const float *mean,
                           const float *invstd, const float *gamma, const float *beta, const float *prelu_a, float *out, void *stream);
int istnet_gather_bn_prelu_bwd(const float *y, int B, long long HW, int C, int N, const long long *choose, const float *mean,
                               const float *invstd, const float *gamma, const float *beta, const float *prelu_a, const float *dout,
                               float *g_dense, double *slope_ws, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 5. Channels-last PointNet++ companions (rows = points, features[b][point][channel])
 * ---------------------------------------------------------------------------------------------------- */

/* QueryAndGroup (pointnet2_utils.py:317-377) straight into the operand planes of the first shared-MLP GEMM:
 * row (b,j,l) = [xyz[b,idx[b,j,l]] - new_xyz[b,j] (3 channels first) | feats[b,idx[b,j,l],0:C]];  feats may be NULL when C == 0. */
int istnet_group_rows_split(int B, int N, int M, int ns, int C, const float *xyz, const float *new_xyz, const float *feats,
                            const int32_t *idx, void *planes, long long plane_stride, int nsplit, int cs, void *stream);
/* adjoint w.r.t. feats: d_feats[b,idx,c] += d_grouped[row, 3+c]  (d_feats is zeroed here) */
int istnet_group_rows_bwd(int B, int N, int M, int ns, int C, const float *d_grouped, const int32_t *idx, float *d_feats, void *stream);
/* last BN + ReLU of a SharedMLP fused with F.max_pool2d over nsample (pointnet2_modules.py:65-69): y [G*ns][C] -> out[g][out_off+c] */
int istnet_bn_relu_maxrows(const float *y, long long G, int ns, int C, const float *mean, const float *invstd, const float *gamma,
                           const float *beta, float *out, int out_ld, int out_off, uint8_t *argmax, void *stream);
int istnet_maxrows_bwd(const float *y, long long G, int ns, int C, const float *mean, const float *invstd, const float *gamma,
                       const float *beta, const float *dz, int dz_ld, int dz_off, const uint8_t *argmax, float *gsel, void *stream);
/* three_interpolate (interpolate_gpu.cu:77-159) on rows: feats [B][m][C] -> out [B][n][out_ld] at channel out_off; and its adjoint */
int istnet_interp_rows(int B, int m, int n, int C, const float *feats, const int32_t *idx, const float *weight, float *out, int out_ld,
                       int out_off, void *stream);
int istnet_interp_rows_bwd(int B, int m, int n, int C, const float *dout, int d_ld, int d_off, const int32_t *idx, const float *weight,
                           float *d_feats, void *stream);
/* three_nn (bindings.cpp:15) + the interpolation weights of PointnetFPModule.forward (pointnet2_utils.py:142 sqrt; pointnet2_modules.py:
 * 186-188 recip = 1/(dist + 1e-8), weight = recip / sum(recip)) in one launch; dist2 may be null */
int istnet_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx, float *weight,
                            void *stream);
/* operand planes of the first FP-layer GEMM: row (b,j) = [three_interpolate(feats [B][m][C2]) | skip [B][n][C1]] (the torch.cat of
 * pointnet2_modules.py:196-199 happens in the split; C1 may be 0) */
int istnet_interp_concat_split(int B, int m, int n, int C2, int C1, const float *feats, const int32_t *idx, const float *weight,
                               const float *skip, void *planes, long long plane_stride, int nsplit, int cs, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 5b. Fused set-abstraction level (csrc/sa_fused.cu) — PointnetSAModuleMSG.forward (pointnet2_modules.py:29-73) for one level
 *     with its two (radius, nsample) scales in ONE launch per pass: ball query (ball_query_gpu.cu:14-49, bit-exact) + grouping
 *     (pointnet2_utils.py:335-367) + SharedMLP 3 x [1x1 conv, train-mode BatchNorm, ReLU] (pytorch_utils.py:25-206) + max over
 *     nsample (pointnet2_modules.py:66-68), nothing grouped written to memory.  Supported widths: istnet_sa_level_supported.
 *     xyz [B,N,3], new_xyz [B,M,3] (M % 8 == 0), u [B*N][ldu] = F Wf^T of both scales (scale s at column s*C0; null when the level
 *     has no input features), nsample in {16, 32}.
 *     forward pass p (0, 1, 2) ends with the statistics of layer p: scales[s].part (2*ISTNET_FIN_ROWS*C_p floats) + scales[s].fin
 *     (ISTNET_FIN_BN_STATS; null part: running statistics, nothing reduced); bn_* of the layers below must be valid.  query != 0:
 *     the pass runs the ball query and writes scales[s].idx, otherwise it reads it.  Pass 2 also writes ysel / asel [B*M][C2]:
 *     the max (gamma2 >= 0) or min (gamma2 < 0) of the pre-BatchNorm layer-2 output over the neighbours and its first position;
 *     istnet_sa_level_final then gives out[bj][s*C2 + c] = relu(bn2(ysel)) — equal to max_k relu(bn2(y_k)) because relu(bn(.)) is
 *     monotone per channel.
 *     backward stage -1: sums of layer 2 from (dz, ysel) -> fin (ISTNET_FIN_BN_BWD, part 3*ISTNET_FIN_ROWS*C2);  stage 0: needs ws2,
 *     writes g1 [rows][C1], sums of layer 1 -> fin, dW2 -> dw [C2][ld_dw];  stage 1: needs ws1, g1; writes g0, sums of layer 0 -> fin,
 *     dW1 -> dw;  stage 2: needs ws0, g0; dU[point][s*C0 + c] += dy0 (zeroed here; float atomics like group_points_grad,
 *     group_points_gpu.cu:48-69), dWx -> dw[c*ld_dw + 0..2].  part_w: ISTNET_FIN_ROWS * (rows*cols of that weight) floats, tickets_w:
 *     ISTNET_FIN_TICKETS zeroed counters.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct istnet_sa_scale {
    float radius;
    int nsample;
    int32_t *idx;
    const float *w0; int ldw0;
    const float *w1, *w2;
    const float *bn_mean[3], *bn_invstd[3], *bn_gamma[3], *bn_beta[3];
    float *part;
    const istnet_fin *fin;
    float *ysel;
    uint8_t *asel;
    const float *dz; int ld_dz, off_dz;
    const double *ws2, *ws1, *ws0;
    float *g1, *g0;
    float *part_w;
    unsigned *tickets_w;
    float *dw; int ld_dw;
} istnet_sa_scale;
int istnet_sa_level_supported(int C0, int C1, int C2);
int istnet_sa_level_forward(int B, int N, int M, int C0, int C1, int C2, const float *xyz, const float *new_xyz, const float *u, int ldu,
                            const istnet_sa_scale *scales, int pass, int query, void *stream);
int istnet_sa_level_final(int B, int M, int C2, const istnet_sa_scale *scales, float *out, int ld_out, void *stream);
int istnet_sa_level_backward(int B, int N, int M, int C0, int C1, int C2, const float *xyz, const float *new_xyz, const float *u, int ldu,
                             float *dU, const istnet_sa_scale *scales, int stage, void *stream);
/* u[r][s*C0 + c] = sum_k F[r][k] * w0_s[c][3 + k] over the R = B*N points (the feature part of layer 0, both scales), and its
 * backward: dF = dU Wf, dWf -> dwf [2*C0][K] (rows of scale 0 first; fixed-order reduction of the per-CTA partials in part_w:
 * ISTNET_FIN_ROWS*2*C0*K floats) */
int istnet_sa_u(int R, int K, int C0, const float *F, const float *w0a, const float *w0b, int ldw0, float *u, void *stream);
int istnet_sa_u_bwd(int R, int K, int C0, const float *F, const float *dU, const float *w0a, const float *w0b, int ldw0, float *dF,
                    float *part_w, float *dwf, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 5c. Pose-head tail (csrc/heads.cu): AdaptiveAvgPool1d(1) over the points, the rotation / translation / size heads
 *     [Linear 512-512, ReLU, Linear 512-256, ReLU, Linear 256-k] (ist_net.py:228-248,296-316) and Ortho6d2Mat
 *     (utils/rotation_utils.py:4-28), forward and backward, batch rows B <= 64.
 * ---------------------------------------------------------------------------------------------------- */
/* pooled[b][c] = mean_n feat[(b*N + n)*C + c]  and its adjoint d_feat = d_pooled / N broadcast over the points (C % 4 == 0) */
int istnet_rows_mean(int B, int N, int C, const float *feat, float *pooled, void *stream);
int istnet_rows_mean_bwd(int B, int N, int C, const float *dpooled, float *dfeat, void *stream);
/* y_h[b][o] = act(sum_k w_h[o][k] x_h[b][k] + bias_h[o]) for up to 3 heads in one launch (nn.Linear + nn.ReLU); O[h] outputs each */
int istnet_heads_linear(int nheads, int B, int K, const float *const *x, const float *const *w, const float *const *bias, float *const *y,
                        const int *O, int relu, void *stream);
/* backward: g = dy * [y > 0] (relu) or dy;  dw_h = g^T x, db_h = sum_b g, dx_h = g w_h (dx[h] may be null) */
int istnet_heads_linear_bwd(int nheads, int B, int K, const float *const *x, const float *const *w, const float *const *y,
                            const float *const *dy, float *const *dx, float *const *dw, float *const *db, const int *O, int relu,
                            void *stream);
/* out[g][c] = sum over rows r with r / group == g of (sum of the nsplit bf16 planes)[r][c]: per-instance column sums of a gradient held
 * as operand planes [nsplit][P][cs] (gradient of the per-instance bias of istnet_conv_gemm's bias_group form).  C, cs % 8 == 0. */
int istnet_rows_group_sum_planes(const void *planes, long long plane_stride, int nsplit, long long P, int C, int cs, int group, float *out,
                                 void *stream);
int istnet_sum3(long long n, const float *a, const float *b, const float *c, float *out, void *stream);
/* R[b] = [x y z] columns with y = n(r6[b][3:6]), z = n(r6[b][0:3] x y), x = y x z, n(v) = v / max(|v|, 1e-8); and the gradient */
int istnet_ortho6d(int B, const float *r6, float *R, void *stream);
int istnet_ortho6d_bwd(int B, const float *r6, const float *dR, float *dr6, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 6. Optimizer step of the training loop (utils/solver.py:41-46,98-99: torch.optim.Adam + CyclicLR)
 * ---------------------------------------------------------------------------------------------------- */

/* torch.optim.Adam (amsgrad=False) over one FLAT bucket of n parameters (n % 4 == 0, 16-byte aligned buffers):
 *   g = grad*grad_scale + weight_decay*p;  m += (g-m)(1-beta1);  v = beta2 v + (1-beta2) g^2;
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps),  t = *step_dev + 1.
 * lr_dev and step_dev are DEVICE scalars (CyclicLR rewrites the learning rate every iteration, solver.py:88-89), so the step can
 * live inside a captured CUDA graph; grad_scale = 1/world_size turns the all-reduced gradient SUM into the mean.
 * istnet_adam_tick increments *step_dev once per optimizer step (after the last bucket).
 * beta1 / beta2 are doubles: torch evaluates 1-beta and the bias corrections in double before rounding to FP32 (1 - 0.999 = 1e-3 exactly,
 * whereas 1.f - 0.999f is off by 4.7e-5 relative). */
int istnet_adam_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long long n, const float *lr_dev, double beta1,
                     double beta2, float eps, float weight_decay, float grad_scale, const long long *step_dev, void *stream);
int istnet_adam_tick(long long *step_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------
 * 7. Per-instance input preparation (reference provider/dataset.py:186-233, :369-409)
 * ---------------------------------------------------------------------------------------------------- */

/* For B instances cut out of F decoded frames resident on the device: rgb_frames [F][H][W][3] uint8 (RGB order), depth [F][H][W]
 * float32 (hole-filled, sensor units), boxes [B][5] = {frame, rmin, rmax, cmin, cmax} (square windows of get_bbox,
 * utils/data_utils.py:43-71), choose [B][N] indices into the flattened crop (dataset.py:192-200).  Produces what the reference's
 * Dataset returns per instance: rgb_out [B][3][S][S] = Normalize(ToTensor(cv2.resize(crop, (S,S), INTER_LINEAR))) bit-exact with
 * OpenCV's 8-bit bilinear path and torchvision; pts_out [B][N][3] = back-projection (z = depth / norm_scale in FP32, x, y in float64,
 * one rounding; + optional float64 jitter noise [B][N][3]); choose_out [B][N] int64 positions on the SxS map.
 * mean3 / std3 are HOST pointers to three floats.  rgb_out may be null (points only); N may be 0 (image only). */
int istnet_prepare_instances(const unsigned char *rgb_frames, const float *depth, int F, int H, int W, const int *boxes, const int *choose,
                             int B, int N, int S, double fx, double fy, double cx, double cy, float norm_scale, const float *mean3,
                             const float *std3, const double *noise, const double *label_params, float *rgb_out, float *pts_out,
                             float *qo_out, long long *choose_out, void *stream);
/* label_params (nullable, DEVICE, [B][13] doubles = translation[3], |size| + 1e-8, rotation[9] row-major) with qo_out [B][N][3]: the NOCS
 * coordinates of the training labels, qo = (pts - t) / (|size| + 1e-8) @ R evaluated on the float64 points (provider/dataset.py:249).
 *
 * istnet_augment_points: provider/data_augmentation.py:45-130 (bounding-box deformation + rigid perturbation, the two augmentations the
 * default configuration enables) applied in place to pts / qo [B][N][3]; params DEVICE [B][32] floats = R[9], t[3], stretch e[3],
 * nocs_scale_aug, do_bb, d[3], Rm[9], do_rt (flags as 0 / 1).  The 3x3 / 3-vector label updates are the caller's (host) job. */
int istnet_augment_points(int B, int N, const float *params, float *pts, float *qo, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ISTNET_B200_H */
