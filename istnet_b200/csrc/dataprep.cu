// Per-instance input preparation on the device: what provider/dataset.py:186-233 (train) / :369-409 (test)
// does per instance on the host with OpenCV + torchvision + numpy once a frame is decoded and its depth hole-filled — crop the RGB
// frame to the (square) instance window, cv2.resize(..., (S, S), INTER_LINEAR) on uint8, ToTensor + Normalize, back-project the
// chosen pixels, re-map `choose` to the SxS map — as two small launches for a whole batch of instances, bit-exact with the host
// libraries (tests/test_gpu_dataprep.py against golden vectors produced by cv2 / torchvision / numpy themselves).
//
// OpenCV's 8-bit bilinear path (modules/imgproc/src/resize.cpp), restated:
//   position   f = (float)((d + 0.5) * scale - 0.5), scale = 1.0 / (S / (double)src);  s = floor(f);  f -= s
//   horizontal: s < 0 -> (s, f) = (0, 0);  s >= src-1 -> (src-1, 0);     vertical: only the row indices s, s+1 are clamped, f is kept
//   coefficients a = rint((1 - f) * 2048), rint(f * 2048)   (saturate_cast<short> = round half to even)
//   horizontal pass in int:  R = p[s] * a0 + p[s+1] * a1;   vertical pass: (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2
#include <stdint.h>

#include "common.cuh"

namespace {
constexpr int kThreads = 256;

struct Lin {
    int s0, s1, a0, a1;
};
__device__ __forceinline__ Lin lin_coef(int d, int S, int src, bool clamp_fraction) {
    const double scale = 1.0 / ((double)S / (double)src);
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);  // separately rounded: no FMA contraction (OpenCV's baseline code has none)
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_fraction) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    Lin r;
    r.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    r.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    r.s0 = min(max(s, 0), src - 1);
    r.s1 = min(max(s + 1, 0), src - 1);
    return r;
}

// boxes: [B][5] = frame, rmin, rmax, cmin, cmax.  One thread per output pixel (3 channels); grid.y = instance.
__global__ void __launch_bounds__(kThreads) crop_resize_normalize_kernel(const uint8_t *__restrict__ frames, int H, int W, const int *__restrict__ boxes,
                                                                         int S, float m0, float m1, float m2, float s0, float s1, float s2,
                                                                         float *__restrict__ out) {
    const int b = blockIdx.y;
    const int *bx = boxes + b * 5;
    const int frame = bx[0], rmin = bx[1], rmax = bx[2], cmin = bx[3], cmax = bx[4];
    const int ch = rmax - rmin, cw = cmax - cmin;
    const uint8_t *src = frames + ((size_t)frame * H + rmin) * W * 3 + (size_t)cmin * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
        const int dy = i / S, dx = i - dy * S;
        const Lin lx = lin_coef(dx, S, cw, true), ly = lin_coef(dy, S, ch, false);
        const uint8_t *r0 = src + (size_t)ly.s0 * W * 3, *r1 = src + (size_t)ly.s1 * W * 3;
        const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int R0 = (int)r0[lx.s0 * 3 + c] * lx.a0 + (int)r0[lx.s1 * 3 + c] * lx.a1;
            const int R1 = (int)r1[lx.s0 * 3 + c] * lx.a0 + (int)r1[lx.s1 * 3 + c] * lx.a1;
            int v = (((ly.a0 * (R0 >> 4)) >> 16) + ((ly.a1 * (R1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            // ToTensor: float(u8) / 255;  Normalize: (x - mean) / std   (IEEE divisions, as torch)
            const float x = __fdiv_rn((float)v, 255.f);
            out[(((size_t)b * 3 + c) * S + dy) * S + dx] = __fdiv_rn(__fsub_rn(x, mean[c]), stdv[c]);
        }
    }
}

// pts[b][n] = back-projection of crop pixel choose[b][n]; choose_out[b][n] = its position on the SxS map
__global__ void __launch_bounds__(kThreads) back_project_kernel(const float *__restrict__ depth, int H, int W, const int *__restrict__ boxes,
                                                                const int *__restrict__ choose, int N, int S, double fx, double fy, double cx,
                                                                double cy, float norm_scale, const double *__restrict__ noise,
                                                                const double *__restrict__ lab, float *__restrict__ pts, float *__restrict__ qo,
                                                                long long *__restrict__ choose_out) {
    const int b = blockIdx.y;
    const int *bx = boxes + b * 5;
    const int frame = bx[0], rmin = bx[1], rmax = bx[2], cmin = bx[3], cmax = bx[4];
    const int cw = cmax - cmin, crop = rmax - rmin;
    const double ratio = (double)S / (double)crop;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int k = choose[(size_t)b * N + n];
        const int r = rmin + k / cw, c = cmin + k % cw;
        const float z = __fdiv_rn(depth[((size_t)frame * H + r) * W + c], norm_scale);  // depth / norm_scale in FP32 (float32 depth map)
        const double zd = (double)z;
        double x = __ddiv_rn(__dmul_rn((double)c - cx, zd), fx), y = __ddiv_rn(__dmul_rn((double)r - cy, zd), fy), zz = zd;
        float px = (float)x, py = (float)y, pz = (float)zz;
        double dx_ = (double)px, dy_ = (double)py, dz_ = (double)pz;  // the float32 point as float64 (dataset.py:209-210)
        if (noise) {  // pts (float32) + float64 jitter, rounded once
            const double *nz = noise + ((size_t)b * N + n) * 3;
            dx_ = __dadd_rn(dx_, nz[0]); dy_ = __dadd_rn(dy_, nz[1]); dz_ = __dadd_rn(dz_, nz[2]);
            px = (float)dx_; py = (float)dy_; pz = (float)dz_;
        }
        float *o = pts + ((size_t)b * N + n) * 3;
        o[0] = px; o[1] = py; o[2] = pz;
        if (lab && qo) {  // dataset.py:249: qo = (pts - t) / (|size| + 1e-8) @ R on the float64 points; lab[b] = t[3], den, R[9] (row-major)
            const double *L = lab + (size_t)b * 13;
            const double ux = __ddiv_rn(__dsub_rn(dx_, L[0]), L[3]), uy = __ddiv_rn(__dsub_rn(dy_, L[1]), L[3]), uz = __ddiv_rn(__dsub_rn(dz_, L[2]), L[3]);
            float *q = qo + ((size_t)b * N + n) * 3;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                q[j] = (float)__dadd_rn(__dadd_rn(__dmul_rn(ux, L[4 + j]), __dmul_rn(uy, L[7 + j])), __dmul_rn(uz, L[10 + j]));
        }
        // dataset.py:221-226: row / column of the crop pixel use the crop HEIGHT for both (square windows)
        const int col = k % crop, row = k / crop;
        choose_out[(size_t)b * N + n] = (long long)(floor((double)row * ratio) * (double)S + floor((double)col * ratio));
    }
}
// provider/data_augmentation.py:45-130 applied to the points and NOCS coordinates of every instance (the 3x3 / 3-vector label updates
// are done by the caller): par[b] = R[9], t[3], e[3], k, do_bb, d[3], Rm[9], do_rt  (30 floats, stride 32).
//   bounding-box deformation: p <- R ((R^T (p - t)) * e) + t,  q <- (q * e) / k          (e = per-axis stretch, k = nocs_scale_aug)
//   rigid perturbation:       p <- Rm (p + d)
// FP32 like the torch CPU code, multiplies and adds separately rounded.
__global__ void __launch_bounds__(kThreads) augment_points_kernel(int N, const float *__restrict__ par, float *__restrict__ pts, float *__restrict__ qo) {
    const int b = blockIdx.y;
    const float *P = par + (size_t)b * 32;
    const bool do_bb = P[16] != 0.f, do_rt = P[29] != 0.f;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float *p = pts + ((size_t)b * N + n) * 3;
        float x = p[0], y = p[1], z = p[2];
        if (do_bb) {
            const float ax = __fsub_rn(x, P[9]), ay = __fsub_rn(y, P[10]), az = __fsub_rn(z, P[11]);
            float r[3];
#pragma unroll
            for (int j = 0; j < 3; ++j)  // (p - t) @ R
                r[j] = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(ax, P[j]), __fmul_rn(ay, P[3 + j])), __fmul_rn(az, P[6 + j])), P[12 + j]);
            x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], P[0]), __fmul_rn(r[1], P[1])), __fmul_rn(r[2], P[2])), P[9]);   // rep @ R^T + t
            y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], P[3]), __fmul_rn(r[1], P[4])), __fmul_rn(r[2], P[5])), P[10]);
            z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], P[6]), __fmul_rn(r[1], P[7])), __fmul_rn(r[2], P[8])), P[11]);
            if (qo) {
                float *q = qo + ((size_t)b * N + n) * 3;
#pragma unroll
                for (int j = 0; j < 3; ++j) q[j] = __fdiv_rn(__fmul_rn(q[j], P[12 + j]), P[15]);
            }
        }
        if (do_rt) {
            const float ax = __fadd_rn(x, P[17]), ay = __fadd_rn(y, P[18]), az = __fadd_rn(z, P[19]);
            x = __fadd_rn(__fadd_rn(__fmul_rn(ax, P[20]), __fmul_rn(ay, P[21])), __fmul_rn(az, P[22]));  // (p + d) @ Rm^T
            y = __fadd_rn(__fadd_rn(__fmul_rn(ax, P[23]), __fmul_rn(ay, P[24])), __fmul_rn(az, P[25]));
            z = __fadd_rn(__fadd_rn(__fmul_rn(ax, P[26]), __fmul_rn(ay, P[27])), __fmul_rn(az, P[28]));
        }
        p[0] = x; p[1] = y; p[2] = z;
    }
}
}  // namespace

extern "C" int istnet_augment_points(int B, int N, const float *params, float *pts, float *qo, void *stream) {
    if (B <= 0 || N <= 0 || !params || !pts) return ISTNET_ERR_BAD_ARG;
    if (B > 65535) return ISTNET_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((N + kThreads - 1) / kThreads), (unsigned)B);
    augment_points_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(N, params, pts, qo);
    ISTNET_LAUNCH_CHECK();
    return ISTNET_OK;
}

extern "C" int istnet_prepare_instances(const unsigned char *rgb_frames, const float *depth, int F, int H, int W, const int *boxes,
                                        const int *choose, int B, int N, int S, double fx, double fy, double cx, double cy,
                                        float norm_scale, const float *mean3, const float *std3, const double *noise,
                                        const double *label_params, float *rgb_out, float *pts_out, float *qo_out, long long *choose_out,
                                        void *stream) {
    if (F <= 0 || H <= 0 || W <= 0 || B <= 0 || S <= 0 || N < 0 || !boxes || !mean3 || !std3) return ISTNET_ERR_BAD_ARG;
    if ((rgb_out && !rgb_frames) || (N > 0 && (!depth || !choose || !pts_out || !choose_out))) return ISTNET_ERR_BAD_ARG;
    if (B > 65535) return ISTNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (rgb_out) {
        dim3 grid((unsigned)((S * S + kThreads - 1) / kThreads), (unsigned)B);
        crop_resize_normalize_kernel<<<grid, kThreads, 0, st>>>(rgb_frames, H, W, boxes, S, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2],
                                                                 rgb_out);
        ISTNET_LAUNCH_CHECK();
    }
    if (N > 0) {
        dim3 grid((unsigned)((N + kThreads - 1) / kThreads), (unsigned)B);
        back_project_kernel<<<grid, kThreads, 0, st>>>(depth, H, W, boxes, choose, N, S, fx, fy, cx, cy, norm_scale, noise, label_params, pts_out, qo_out,
                                                       choose_out);
        ISTNET_LAUNCH_CHECK();
    }
    return ISTNET_OK;
}
