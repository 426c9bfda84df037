/*
 * oracle/pointnet2_oracle.c -- CPU restatement of the reference's pointnet2_batch
 * CUDA kernels.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 * Never on the product path.
 *
 * Every function emulates the reference kernel's *thread mapping* where that
 * mapping decides the result (FPS tie-breaking), and the compiled arithmetic
 * (FMA contraction as nvcc 12.9 emits it for sm_100a):
 *
 *     d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy))          (dx = x2 - x1 rounded first)
 *
 * Build: gcc -O2 -ffp-contract=off (so only the explicit fmaf() calls fuse).
 * Parity pin: tests/golden/ holds outputs of the reference's own CUDA kernels
 * (oracle/_ref, built from /root/reference by oracle/build_ref.py) run on a B200;
 * tests/test_oracle_cpu.py checks this file against them bit for bit.
 *
 * Reference (all under /root/reference/pcdet/ops/pointnet2/pointnet2_batch/src):
 *   sampling_gpu.cu:93-98      __update (tree-reduction tie rule)
 *   sampling_gpu.cu:100-216    farthest_point_sampling_kernel
 *   sampling_gpu.cu:750-855    furthest_point_sampling_matrix_kernel (== :262-377)
 *   sampling_gpu.cu:424-541    furthest_point_sampling_with_weighted_dist_kernel
 *   sampling_gpu.cu:901-1022   furthest_point_sampling_weights_kernel
 *   sampling_gpu.cu:15-31,53-70   gather_points(+grad)
 *   ball_query_gpu.cu:75-112, 138-176  ball_query(+dilated)
 *   group_points_gpu.cu:14-31, 53-72   group_points(+grad)
 *   interpolate_gpu.cu:16-59, 84-104, 127-149  three_nn, three_interpolate(+grad)
 *   cuda_utils.h:10-14         opt_n_threads
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define ORC_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORC_CLONES
#endif

#define ORC_API __attribute__((visibility("default")))

/* cuda_utils.h:10-14 -- same double arithmetic, same truncation. */
ORC_API int orc_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

static inline float sqdist(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* The 10-level shared-memory tree of sampling_gpu.cu:146-214 with __update's
 * rule "v2 > v1 ? i2 : i1" (the lower slot wins ties). */
static inline int tree_argmax(float *dists, int *dists_i, int bs) {
    for (int s = bs >> 1; s >= 1; s >>= 1) {
        for (int t = 0; t < s; ++t) {
            const float v1 = dists[t], v2 = dists[t + s];
            const int i1 = dists_i[t], i2 = dists_i[t + s];
            dists[t] = fmaxf(v1, v2);
            dists_i[t] = v2 > v1 ? i2 : i1;
        }
    }
    return dists_i[0];
}

/* mode bits for the generic driver */
#define FPS_XYZ 0      /* distances from coordinates */
#define FPS_MATRIX 1   /* distances from a (N,N) matrix row */

ORC_CLONES
static void fps_one(int n, int m, const float *xyz_or_mat, int from_matrix, const float *weights,
                    float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    int j0 = 1;
    if (weights) {
        j0 = 0; /* weighted variants: round 0 selects argmax(weights) */
    } else {
        idxs[0] = 0;
    }
    for (int j = j0; j < m; ++j) {
        float x1 = 0.f, y1 = 0.f, z1 = 0.f;
        if (!from_matrix) {
            x1 = xyz_or_mat[old * 3 + 0];
            y1 = xyz_or_mat[old * 3 + 1];
            z1 = xyz_or_mat[old * 3 + 2];
        }
        for (int tid = 0; tid < bs; ++tid) {
            int besti = 0;
            float best = -1.f;
            for (int k = tid; k < n; k += bs) {
                float d2;
                if (weights && j == 0) {
                    d2 = weights[k];
                } else {
                    float d;
                    if (from_matrix)
                        d = xyz_or_mat[(size_t)old * n + k];
                    else
                        d = sqdist(x1, y1, z1, xyz_or_mat[k * 3 + 0], xyz_or_mat[k * 3 + 1], xyz_or_mat[k * 3 + 2]);
                    d = fminf(d, temp[k]);
                    temp[k] = d;
                    if (weights) {
                        /* d * max(weights[k], 1e-12): float * double -> double -> float
                         * (sampling_gpu.cu:466, :947) */
                        double w = (double)weights[k];
                        if (!(w > 1e-12)) w = 1e-12; /* max(w, 1e-12) */
                        d2 = (float)((double)d * w);
                    } else {
                        d2 = d;
                    }
                }
                if (d2 > best) {
                    besti = k;
                    best = d2;
                }
            }
            dists[tid] = best;
            dists_i[tid] = besti;
        }
        old = tree_argmax(dists, dists_i, bs);
        idxs[j] = old;
    }
    free(dists);
    free(dists_i);
}

/* dataset (B,N,3), temp (B,N) pre-filled 1e10 by the caller, idxs (B,M) */
ORC_API void orc_fps(int b, int n, int m, const float *xyz, float *temp, int *idxs) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi)
        fps_one(n, m, xyz + (size_t)bi * n * 3, 0, NULL, temp + (size_t)bi * n, idxs + (size_t)bi * m);
}

ORC_API void orc_fps_weights(int b, int n, int m, const float *xyz, const float *weights, float *temp, int *idxs) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi)
        fps_one(n, m, xyz + (size_t)bi * n * 3, 0, weights + (size_t)bi * n, temp + (size_t)bi * n,
                idxs + (size_t)bi * m);
}

ORC_API void orc_fps_matrix(int b, int n, int m, const float *matrix, float *temp, int *idxs) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi)
        fps_one(n, m, matrix + (size_t)bi * n * n, 1, NULL, temp + (size_t)bi * n, idxs + (size_t)bi * m);
}

ORC_API void orc_fps_weighted_matrix(int b, int n, int m, const float *matrix, const float *weights, float *temp,
                                     int *idxs) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi)
        fps_one(n, m, matrix + (size_t)bi * n * n, 1, weights + (size_t)bi * n, temp + (size_t)bi * n,
                idxs + (size_t)bi * m);
}

/* points (B,C,N), idx (B,M) -> out (B,C,M) */
ORC_API void orc_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                out[((size_t)bi * c + ci) * m + p] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]];
}

/* grad_out (B,C,M), idx (B,M) -> grad_points (B,C,N) += (caller zeroes) */
ORC_API void orc_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                    float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]] += grad_out[((size_t)bi * c + ci) * m + p];
}

/* new_xyz (B,M,3), xyz (B,N,3) -> idx_cnt (B,M), idx (B,M,nsample) (caller zeroes both).
 * radius_in < 0 selects the plain ball (ball_query_gpu.cu:75-112); otherwise the
 * annulus rin^2 <= d2 < rout^2 (:138-176). */
ORC_CLONES
static void ball_query_impl(int b, int n, int m, int dilated, float radius_in, float radius_out, int nsample,
                            const float *new_xyz, const float *xyz, int *idx_cnt, int *idx) {
    const float rin2 = radius_in * radius_in;
    const float rout2 = radius_out * radius_out;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int p = 0; p < m; ++p) {
            const float *q = new_xyz + ((size_t)bi * m + p) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int *row = idx + ((size_t)bi * m + p) * nsample;
            const float qx = q[0], qy = q[1], qz = q[2];
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                /* (new - x)^2: same magnitudes as (x - new)^2 */
                const float d2 = sqdist(pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2], qx, qy, qz);
                const int hit = dilated ? (d2 >= rin2 && d2 < rout2) : (d2 < rout2);
                if (hit) row[cnt++] = k;
            }
            idx_cnt[(size_t)bi * m + p] = cnt;
            for (int l = 0; cnt < nsample; ++l, ++cnt) row[cnt] = row[l]; /* cyclic pad */
        }
    }
}

ORC_API void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                            int *idx_cnt, int *idx) {
    ball_query_impl(b, n, m, 0, 0.f, radius, nsample, new_xyz, xyz, idx_cnt, idx);
}

ORC_API void orc_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample,
                                    const float *new_xyz, const float *xyz, int *idx_cnt, int *idx) {
    ball_query_impl(b, n, m, 1, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx);
}

/* points (B,C,N), idx (B,P,S) -> out (B,C,P,S) */
ORC_API void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                              float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            const int *ii = idx + (size_t)bi * npoints * nsample;
            float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
            for (int e = 0; e < npoints * nsample; ++e) dst[e] = src[ii[e]];
        }
}

ORC_API void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                   const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const int *ii = idx + (size_t)bi * npoints * nsample;
            const float *src = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            for (int e = 0; e < npoints * nsample; ++e) dst[ii[e]] += src[e];
        }
}

/* unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3), idx (B,N,3) */
ORC_CLONES
static void three_nn_impl(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sqdist(kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2], ux, uy, uz);
                if (d < best1) {
                    best3 = best2; i3 = i2;
                    best2 = best1; i2 = i1;
                    best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2;
                    best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *o = dist2 + ((size_t)bi * n + p) * 3;
            int *oi = idx + ((size_t)bi * n + p) * 3;
            o[0] = (float)best1; o[1] = (float)best2; o[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

ORC_API void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    three_nn_impl(b, n, m, unknown, known, dist2, idx);
}

/* points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N)
 * compiled form: fmaf(w2,p2, fmaf(w0,p0, w1*p1))   */
ORC_CLONES
static void three_interpolate_impl(int b, int c, int m, int n, const float *points, const int *idx,
                                   const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            float *dst = out + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *ii = idx + ((size_t)bi * n + p) * 3;
                dst[p] = fmaf(w[2], src[ii[2]], fmaf(w[0], src[ii[0]], w[1] * src[ii[1]]));
            }
        }
}

ORC_API void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                   const float *weight, float *out) {
    three_interpolate_impl(b, c, m, n, points, idx, weight, out);
}

/* grad_out (B,C,N), idx/weight (B,N,3) -> grad_points (B,C,M) += */
ORC_API void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                        const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bi * c + ci) * m;
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int *ii = idx + ((size_t)bi * n + p) * 3;
                dst[ii[0]] += g[p] * w[0];
                dst[ii[1]] += g[p] * w[1];
                dst[ii[2]] += g[p] * w[2];
            }
        }
}
