/*
 * oracle/stack_oracle.c -- CPU restatement of the reference's pointnet2_stack kernels that SA layers >= 1 use
 * (SURVEY.md 8 f3).  TEST INFRASTRUCTURE ONLY; never on the product path.
 *
 * Reference (all under /root/reference/pcdet/ops/pointnet2/pointnet2_stack/src):
 *   voxel_query_gpu.cu:10-98     voxel_query_kernel_stack
 *   voxel_query_gpu.cu:125-215   voxel_query_dilated_kernel_stack
 *   group_points_gpu.cu:14-42, 66-95   group_points(_grad)_kernel_stack
 *   sampling_gpu.cu:188-316      stack_farthest_point_sampling_kernel<1024>
 *
 * The voxel queries replace neighbours at random once a centre has more than nsample hits ("reservoir"): the
 * reference draws from cuRAND's default generator, XORWOW, seeded per centre with curand_init(pt_idx, 0, 0).  That
 * generator is a published algorithm (Marsaglia's xorwow; NVIDIA cuRAND 12.9, curand_kernel.h: _curand_init_inplace
 * :800-843 -- with subsequence = offset = 0 the skip-ahead is the identity -- curand() :863-874, _curand_uniform in
 * curand_uniform.h), restated below; the compiled arithmetic of the kernel (FMA contraction of nvcc 12.9 for sm_100a,
 * checked in the SASS of the reference build) is spelled with explicit fmaf().  Pinned by tests/golden/stack_ops.npz
 * (outputs of the reference's own CUDA kernels on a B200, tests/golden/make_golden.py) in tests/test_oracle_cpu.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    uint32_t d, v[5];
} xorwow_t;

static void xorwow_init(xorwow_t *s, uint64_t seed) {
    const uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}

static uint32_t xorwow_next(xorwow_t *s) {
    const uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1];
    s->v[1] = s->v[2];
    s->v[2] = s->v[3];
    s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}

/* _curand_uniform: x * 2^-32 + 2^-33, contracted to one FFMA by nvcc (default -fmad=true) */
static float xorwow_uniform(xorwow_t *s) {
    return fmaf((float)xorwow_next(s), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

/* new_xyz (M,3), xyz (N,3), new_coords (M,4) [b,z,y,x], point_indices (B,R1,R2,R3) -> idx (M,nsample) (zeroed by
 * the caller), cnt_unique (M), idx_cnt (M) [dilated only; may be NULL].  former_radius < 0: plain query. */
ORC_API void orc_voxel_query(int m, int r1, int r2, int r3, int nsample, float former_radius, float radius, int z_range,
                             int y_range, int x_range, int z_stride, int y_stride, int x_stride, const float *new_xyz,
                             const float *xyz, const int *new_coords, const int *point_indices, int *idx_all,
                             int *cnt_unique, int *idx_cnt) {
    const int dilated = former_radius >= 0.f;
    const float radius2 = radius * radius;
    const float former2 = former_radius * former_radius;
#pragma omp parallel for schedule(dynamic, 16)
    for (int pt = 0; pt < m; ++pt) {
        int *idx = idx_all + (size_t)pt * nsample;
        xorwow_t st;
        xorwow_init(&st, (uint64_t)pt);
        const float nx = new_xyz[pt * 3 + 0], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        const int b = new_coords[pt * 4 + 0], cz = new_coords[pt * 4 + 1], cy = new_coords[pt * 4 + 2],
                  cx = new_coords[pt * 4 + 3];
        int cnt = 0, cnt2 = 0, range_point_num = 0;
        for (int dz = -z_range; dz <= z_range; dz += z_stride) {
            const int z = cz + dz;
            if (z < 0 || z >= r1) continue;
            for (int dy = -y_range; dy <= y_range; dy += y_stride) {
                const int y = cy + dy;
                if (y < 0 || y >= r2) continue;
                for (int dx = -x_range; dx <= x_range; dx += x_stride) {
                    const int x = cx + dx;
                    if (x < 0 || x >= r3) continue;
                    const int nb = point_indices[(((size_t)b * r1 + z) * r2 + y) * r3 + x];
                    if (nb < 0) continue;
                    ++range_point_num;
                    const float ex = xyz[nb * 3 + 0] - nx, ey = xyz[nb * 3 + 1] - ny, ez = xyz[nb * 3 + 2] - nz;
                    const float d2 = fmaf(ez, ez, fmaf(ex, ex, ey * ey));
                    if (d2 > radius2 || (dilated && d2 < former2)) continue;
                    ++cnt2;
                    if (cnt < nsample) {
                        if (cnt == 0)
                            for (int l = 0; l < nsample; ++l) idx[l] = nb;
                        idx[cnt] = nb;
                        ++cnt;
                    } else {
                        const float rnd = xorwow_uniform(&st);
                        if (rnd < ((float)nsample / (float)cnt2)) {
                            const int ins = (int)(ceilf(xorwow_uniform(&st) * (float)nsample) - 1.0f);
                            idx[ins] = nb;
                        }
                    }
                }
            }
        }
        cnt_unique[pt] = range_point_num;
        if (idx_cnt) idx_cnt[pt] = cnt;
        if (cnt == 0) idx[0] = -1;
        for (int l = 0; cnt < nsample; ++l, ++cnt) idx[cnt] = idx[l];
    }
}

static void batch_of(int b, const int *cnt, int pt, int *bs_idx) {
    int k = 0, acc = cnt[0];
    for (int q = 1; q < b; ++q) {
        if (pt < acc) break;
        acc += cnt[q];
        k = q;
    }
    *bs_idx = k;
}

/* features (N,C), idx (M,nsample) frame-local indices -> out (M,C,nsample) */
ORC_API void orc_stack_group_points(int b, int m, int c, int nsample, const float *features, const int *features_batch_cnt,
                                    const int *idx, const int *idx_batch_cnt, float *out) {
#pragma omp parallel for schedule(static)
    for (int pt = 0; pt < m; ++pt) {
        int bs;
        batch_of(b, idx_batch_cnt, pt, &bs);
        size_t start = 0;
        for (int k = 0; k < bs; ++k) start += features_batch_cnt[k];
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < nsample; ++s)
                out[((size_t)pt * c + ci) * nsample + s] = features[(start + idx[(size_t)pt * nsample + s]) * c + ci];
    }
}

ORC_API void orc_stack_group_points_grad(int b, int m, int c, int n, int nsample, const float *grad_out, const int *idx,
                                         const int *idx_batch_cnt, const int *features_batch_cnt, float *grad_features) {
    for (int pt = 0; pt < m; ++pt) {
        int bs;
        batch_of(b, idx_batch_cnt, pt, &bs);
        size_t start = 0;
        for (int k = 0; k < bs; ++k) start += features_batch_cnt[k];
        for (int ci = 0; ci < c; ++ci)
            for (int s = 0; s < nsample; ++s)
                grad_features[(start + idx[(size_t)pt * nsample + s]) * c + ci] += grad_out[((size_t)pt * c + ci) * nsample + s];
    }
}

/* dataset (sum N,3), temp (sum N) pre-filled 1e10, xyz_batch_cnt (B), num_sampled_points (B) -> idxs (sum M), global
 * row numbers.  Block size is ALWAYS 1024 here (sampling_gpu.cu:339), whatever the cloud size. */
ORC_API void orc_stack_fps(int b, const float *dataset, float *temp, const int *xyz_batch_cnt, int *idxs,
                           const int *num_sampled_points) {
    const int bs = 1024;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        size_t start = 0, ostart = 0;
        for (int k = 0; k < bi; ++k) {
            start += xyz_batch_cnt[k];
            ostart += num_sampled_points[k];
        }
        const float *pts = dataset + start * 3;
        float *tp = temp + start;
        int *out = idxs + ostart;
        const int n = xyz_batch_cnt[bi], m = num_sampled_points[bi];
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = (int)start; /* written unconditionally (sampling_gpu.cu:229), even when m == 0 */
        for (int j = 1; j < m; ++j) {
            const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.f;
                for (int k = tid; k < n; k += bs) {
                    const float ex = pts[k * 3 + 0] - x1, ey = pts[k * 3 + 1] - y1, ez = pts[k * 3 + 2] - z1;
                    const float d = fmaf(ez, ez, fmaf(ex, ex, ey * ey));
                    const float d2 = fminf(d, tp[k]);
                    tp[k] = d2;
                    if (d2 > best) {
                        besti = k;
                        best = d2;
                    }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs >> 1; s >= 1; s >>= 1)
                for (int t = 0; t < s; ++t) {
                    const float v1 = dists[t], v2 = dists[t + s];
                    const int i1 = dists_i[t], i2 = dists_i[t + s];
                    dists[t] = fmaxf(v1, v2);
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old + (int)start;
        }
        free(dists);
        free(dists_i);
    }
}
