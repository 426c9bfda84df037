/*
 * oracle/iou3d_oracle.c -- CPU restatement of the reference's rotated BEV IoU
 * and NMS.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline).
 *
 * Follows /root/reference/pcdet/ops/iou3d_nms/src:
 *   iou3d_cpu.cpp:30-229        box geometry (identical algorithm to the device
 *                               code in iou3d_nms_kernel.cu:35-234)
 *   iou3d_cpu.cpp:232-252       boxes_iou_bev_cpu double loop
 *   iou3d_nms_kernel.cu:267-311 nms_kernel: 64-wide suppression words
 *   iou3d_nms_kernel.cu:314-325 iou_normal (axis-aligned)
 *   iou3d_nms.cpp:116-131       greedy OR-sweep over the mask
 *
 * Numerics: built with gcc -O2 -ffp-contract=off on x86-64 (no FMA contraction,
 * glibc cosf/sinf/atan2f), i.e. the arithmetic g++ gives iou3d_cpu.cpp.  That
 * is bit-identical to the reference's boxes_iou_bev_cpu (pinned by
 * tests/test_oracle_cpu.py against oracle/_ref and the committed fixture) and
 * equal to the *GPU* kernels only to ~1e-6 (libdevice trig + FMA contraction),
 * so against the GPU this file is a tolerance oracle for IoU values; bit-exact
 * keep-lists are pinned against the reference CUDA extension itself.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static const float kEps = 1e-8f;
static const float kMargin = 1e-2f;

typedef struct { float x, y; } pt2;

static inline float cross2(pt2 a, pt2 b) { return a.x * b.y - a.y * b.x; }

/* (p1 - p0) x (p2 - p0) */
static inline float cross3(pt2 p1, pt2 p2, pt2 p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

static inline float fmin2(float a, float b) { return a > b ? b : a; }
static inline float fmax2(float a, float b) { return a > b ? a : b; }

static inline int bbox_overlap(pt2 p1, pt2 p2, pt2 q1, pt2 q2) {
    return fmin2(p1.x, p2.x) <= fmax2(q1.x, q2.x) && fmin2(q1.x, q2.x) <= fmax2(p1.x, p2.x) &&
           fmin2(p1.y, p2.y) <= fmax2(q1.y, q2.y) && fmin2(q1.y, q2.y) <= fmax2(p1.y, p2.y);
}

static inline int inside_box(const float *box, pt2 p) {
    const float cx = box[0], cy = box[1];
    const float c = cosf(-box[6]), s = sinf(-box[6]);
    const float rx = (p.x - cx) * c + (p.y - cy) * (-s);
    const float ry = (p.x - cx) * s + (p.y - cy) * c;
    return fabsf(rx) < box[3] / 2 + kMargin && fabsf(ry) < box[4] / 2 + kMargin;
}

/* segment p0->p1 against q0->q1 */
static inline int seg_intersect(pt2 p1, pt2 p0, pt2 q1, pt2 q0, pt2 *ans) {
    if (!bbox_overlap(p0, p1, q0, q1)) return 0;
    const float s1 = cross3(q0, p1, p0);
    const float s2 = cross3(p1, q1, p0);
    const float s3 = cross3(p0, q1, q0);
    const float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > kEps) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static inline pt2 rot_about(pt2 c, float ca, float sa, pt2 p) {
    pt2 r;
    r.x = (p.x - c.x) * ca + (p.y - c.y) * (-sa) + c.x;
    r.y = (p.x - c.x) * sa + (p.y - c.y) * ca + c.y;
    return r;
}

static inline int ang_gt(pt2 a, pt2 b, pt2 c) {
    return atan2f(a.y - c.y, a.x - c.x) > atan2f(b.y - c.y, b.x - c.x);
}

static float overlap_area(const float *A, const float *B) {
    const float a_ang = A[6], b_ang = B[6];
    const float ahx = A[3] / 2, bhx = B[3] / 2, ahy = A[4] / 2, bhy = B[4] / 2;
    const float ax1 = A[0] - ahx, ay1 = A[1] - ahy, ax2 = A[0] + ahx, ay2 = A[1] + ahy;
    const float bx1 = B[0] - bhx, by1 = B[1] - bhy, bx2 = B[0] + bhx, by2 = B[1] + bhy;
    const pt2 ca = {A[0], A[1]}, cb = {B[0], B[1]};
    pt2 pa[5] = {{ax1, ay1}, {ax2, ay1}, {ax2, ay2}, {ax1, ay2}};
    pt2 pb[5] = {{bx1, by1}, {bx2, by1}, {bx2, by2}, {bx1, by2}};
    const float cosa = cosf(a_ang), sina = sinf(a_ang);
    const float cosb = cosf(b_ang), sinb = sinf(b_ang);
    for (int k = 0; k < 4; ++k) {
        pa[k] = rot_about(ca, cosa, sina, pa[k]);
        pb[k] = rot_about(cb, cosb, sinb, pb[k]);
    }
    pa[4] = pa[0];
    pb[4] = pb[0];

    pt2 poly[16];
    pt2 ctr = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersect(pa[i + 1], pa[i], pb[j + 1], pb[j], &poly[cnt])) {
                ctr.x = ctr.x + poly[cnt].x;
                ctr.y = ctr.y + poly[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (inside_box(A, pb[k])) {
            ctr.x = ctr.x + pb[k].x;
            ctr.y = ctr.y + pb[k].y;
            poly[cnt++] = pb[k];
        }
        if (inside_box(B, pa[k])) {
            ctr.x = ctr.x + pa[k].x;
            ctr.y = ctr.y + pa[k].y;
            poly[cnt++] = pa[k];
        }
    }
    ctr.x /= cnt;
    ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang_gt(poly[i], poly[i + 1], ctr)) {
                const pt2 t = poly[i];
                poly[i] = poly[i + 1];
                poly[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const pt2 u = {poly[k].x - poly[0].x, poly[k].y - poly[0].y};
        const pt2 v = {poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
        area += cross2(u, v);
    }
    return (float)(fabsf(area) / 2.0);
}

static inline float iou_rot(const float *A, const float *B) {
    const float sa = A[3] * A[4], sb = B[3] * B[4];
    const float so = overlap_area(A, B);
    return so / fmaxf(sa + sb - so, kEps);
}

static inline float iou_axis(const float *a, const float *b) {
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h;
    const float Sa = a[3] * a[4], Sb = b[3] * b[4];
    return inter / fmaxf(Sa + Sb - inter, kEps);
}

/* boxes_a (N,7), boxes_b (M,7) -> out (N,M); threads>1 parallelises rows (the
 * reference loop is serial; the values do not depend on the order). */
ORC_API void orc_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = iou_rot(a + i * 7, b + j * 7);
}

ORC_API void orc_boxes_overlap_bev(int na, const float *a, int nb, const float *b, float *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = overlap_area(a + i * 7, b + j * 7);
}

/* Suppression mask (N, ceil(N/64)) exactly as nms_kernel lays it out, restricted
 * to the words the sweep reads (column block >= row block; diagonal block: bits
 * above the row).  normal != 0 selects the axis-aligned IoU. */
ORC_API void orc_nms_mask(int n, const float *boxes, float thresh, int normal, uint64_t *mask) {
    const int cb = (n + 63) / 64;
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < cb; ++c) {
            uint64_t t = 0;
            if (c >= i / 64) {
                const int j0 = (c == i / 64) ? (i % 64) + 1 : 0;
                const int jn = (n - c * 64) < 64 ? (n - c * 64) : 64;
                for (int j = j0; j < jn; ++j) {
                    const float v = normal ? iou_axis(boxes + i * 7, boxes + (c * 64 + j) * 7)
                                           : iou_rot(boxes + i * 7, boxes + (c * 64 + j) * 7);
                    if (v > thresh) t |= 1ULL << j;
                }
            }
            mask[(size_t)i * cb + c] = t;
        }
    }
}

/* iou3d_nms.cpp:116-131 */
ORC_API int orc_nms_sweep(int n, const uint64_t *mask, int64_t *keep) {
    const int cb = (n + 63) / 64;
    uint64_t *remv = (uint64_t *)calloc(cb > 0 ? cb : 1, sizeof(uint64_t));
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        const int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {
            keep[nk++] = i;
            const uint64_t *p = mask + (size_t)i * cb;
            for (int j = nb; j < cb; ++j) remv[j] |= p[j];
        }
    }
    free(remv);
    return nk;
}

/* boxes already sorted by descending score; keep (N) int64 out; returns num kept */
ORC_API int orc_nms(int n, const float *boxes, float thresh, int normal, int64_t *keep) {
    const int cb = (n + 63) / 64;
    uint64_t *mask = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1) * (cb > 0 ? cb : 1));
    orc_nms_mask(n, boxes, thresh, normal, mask);
    const int nk = orc_nms_sweep(n, mask, keep);
    free(mask);
    return nk;
}

/* Same greedy sweep driven by a dense IoU matrix (N,N) -- lets the tests replay a
 * GPU-computed IoU matrix (bit-exact to the device arithmetic) through the
 * reference's sweep without re-deriving device trig on the CPU. */
ORC_API int orc_nms_from_iou(int n, const float *iou, float thresh, int64_t *keep) {
    unsigned char *dead = (unsigned char *)calloc(n > 0 ? n : 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j)
            if (iou[(size_t)i * n + j] > thresh) dead[j] = 1;
    }
    free(dead);
    return nk;
}
