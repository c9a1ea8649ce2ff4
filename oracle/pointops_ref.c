/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (istnet_b200/).
 *
 * Plain-C CPU restatement of the nine `pointnet2._ext` operators of the reference
 * (/root/reference/model/pointnet2/_ext_src, pybind list at src/bindings.cpp:11-24).
 * The reference has no CPU path of its own ("CPU not supported", e.g. src/sampling.cpp:39), so
 * this file restates what its CUDA kernels compute, including the FMA contraction nvcc applies
 * to the squared-distance expressions and the FPS block-reduction
 * tie-break.  Build with:  gcc -O2 -ffp-contract=off -shared -fPIC  (see oracle/Makefile) so
 * the compiler adds no contraction of its own; every fused multiply-add is an explicit fmaf().
 *
 * Pinned against: (1) the reference's own CUDA extension compiled from /root/reference into
 * oracle/_ref/ and run on the B200 (tests/test_gpu_pointops.py::test_reference_extension_agrees),
 * (2) the golden vectors under tests/golden/ produced with the reference Python modules.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cuda_utils.h:18-24 — opt_n_threads(work) = clamp(2^floor(log2 work), 1, 512). */
int ref_opt_n_threads(int work_size) {
    int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 512) v = 512;
    if (v < 1) v = 1;
    return v;
}

/* Squared distance exactly as the reference's sm_100a SASS evaluates `dx*dx + dy*dy + dz*dz` (nvcc -fmad=true):
 *   FMUL t = dy*dy ; FFMA t = dx*dx + t ; FFMA d = dz*dz + t        (read from oracle/_ref/build/*_gpu.cuda.o with
 * cuobjdump -sass for furthest_point_sampling_kernel, query_ball_point_kernel and three_nn_kernel — the plain multiply is
 * on the MIDDLE term; SURVEY.md Appendix A has the first and second terms swapped). */
static inline float sqdist(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/*
 * sampling_gpu.cu:74-178 (+ sampling.cpp:70-91: idx zero-initialised, temp filled with 1e10).
 * One thread block of S = opt_n_threads(n) threads per batch element. Thread t scans
 * k = t, t+S, ... keeping a strict-> running best (initial best=-1, besti=0); the block then
 * reduces pairs (t, t+s) for s = S/2 .. 1 with `v2 > v1 ? i2 : i1` (sampling_gpu.cu:64-70).
 */
void ref_furthest_point_sampling(int b, int n, int m, const float *xyz, int *idxs) {
    if (m <= 0) return;
    int S = ref_opt_n_threads(n);
    float *temp = (float *)malloc(sizeof(float) * (size_t)n);
    float *dists = (float *)malloc(sizeof(float) * (size_t)S);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)S);
    for (int bi = 0; bi < b; ++bi) {
        const float *d = xyz + (size_t)bi * n * 3;
        int *out = idxs + (size_t)bi * m;
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = d[old * 3 + 0], y1 = d[old * 3 + 1], z1 = d[old * 3 + 2];
            for (int t = 0; t < S; ++t) {
                int besti = 0;
                float best = -1.0f;
                for (int k = t; k < n; k += S) {
                    float dx = d[k * 3 + 0] - x1, dy = d[k * 3 + 1] - y1, dz = d[k * 3 + 2] - z1;
                    float dd = sqdist(dx, dy, dz);
                    float d2 = fminf(dd, temp[k]);
                    temp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int s = S / 2; s >= 1; s >>= 1) {
                for (int t = 0; t < s; ++t) {
                    float v1 = dists[t], v2 = dists[t + s];
                    int i1 = dists_i[t], i2 = dists_i[t + s];
                    dists[t] = fmaxf(v1, v2);
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(temp); free(dists); free(dists_i);
}

/* sampling_gpu.cu:13-25 — out[b,c,j] = points[b,c,idx[b,j]] */
void ref_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* sampling_gpu.cu:39-52 — scatter-add (reference: float atomics; here: j ascending) */
void ref_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] += grad_out[((size_t)i * c + l) * m + j];
}

/* ball_query_gpu.cu:14-49 (+ ball_query.cpp:24-26 zero init). radius arrives as float. */
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
    float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        const float *q = new_xyz + (size_t)bi * m * 3;
        int *o = idx + (size_t)bi * m * nsample;
        for (int j = 0; j < m; ++j) {
            float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float dx = nx - p[k * 3 + 0], dy = ny - p[k * 3 + 1], dz = nz - p[k * 3 + 2];
                float d2 = sqdist(dx, dy, dz);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[j * nsample + l] = k;
                    o[j * nsample + cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* group_points_gpu.cu:13-33 */
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int k = 0; k < nsample; ++k) {
                    int ii = idx[((size_t)bi * npoints + j) * nsample + k];
                    out[(((size_t)bi * c + l) * npoints + j) * nsample + k] = points[((size_t)bi * c + l) * n + ii];
                }
}

/* group_points_gpu.cu:48-69 (atomics in the reference; deterministic j,k-ascending order here) */
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < npoints; ++j)
                for (int k = 0; k < nsample; ++k) {
                    int ii = idx[((size_t)bi * npoints + j) * nsample + k];
                    grad_points[((size_t)bi * c + l) * n + ii] += grad_out[(((size_t)bi * c + l) * npoints + j) * nsample + k];
                }
}

/* interpolate_gpu.cu:14-64 — the cascade compares float d against double best*, values stay FP32 */
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    for (int bi = 0; bi < b; ++bi) {
        const float *u = unknown + (size_t)bi * n * 3;
        const float *kn = known + (size_t)bi * m * 3;
        for (int j = 0; j < n; ++j) {
            float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float dx = ux - kn[k * 3 + 0], dy = uy - kn[k * 3 + 1], dz = uz - kn[k * 3 + 2];
                float d = sqdist(dx, dy, dz);
                if (d < best1) {
                    best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            size_t o = ((size_t)bi * n + j) * 3;
            dist2[o + 0] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
            idx[o + 0] = besti1; idx[o + 1] = besti2; idx[o + 2] = besti3;
        }
    }
}

/* interpolate_gpu.cu:77-106 — `p1*w1 + p2*w2 + p3*w3` contracts to fma(p3,w3, fma(p1,w1, p2*w2)) (same SASS pattern) */
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < n; ++j) {
                size_t o = ((size_t)bi * n + j) * 3;
                const float *p = points + ((size_t)bi * c + l) * m;
                out[((size_t)bi * c + l) * n + j] =
                    fmaf(p[idx[o + 2]], weight[o + 2], fmaf(p[idx[o + 0]], weight[o + 0], p[idx[o + 1]] * weight[o + 1]));
            }
}

/* interpolate_gpu.cu:121-148 (atomics in the reference; deterministic j-ascending here) */
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < n; ++j) {
                size_t o = ((size_t)bi * n + j) * 3;
                float g = grad_out[((size_t)bi * c + l) * n + j];
                float *gp = grad_points + ((size_t)bi * c + l) * m;
                gp[idx[o + 0]] += g * weight[o + 0];
                gp[idx[o + 1]] += g * weight[o + 1];
                gp[idx[o + 2]] += g * weight[o + 2];
            }
}
