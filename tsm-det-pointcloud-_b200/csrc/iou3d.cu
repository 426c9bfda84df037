// iou3d.cu -- rotated BEV overlap / IoU and NMS for sm_100a.
//
// Replaces:
//   boxes_overlap_kernel / boxes_iou_bev_kernel  /root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265
//   nms_kernel / nms_normal_kernel               iou3d_nms_kernel.cu:267-311, 328-372
//   host greedy sweep of nms_gpu/nms_normal_gpu  iou3d_nms.cpp:90-136, 139-186
//
// What is kept from the reference is the ARITHMETIC of one box pair (polygon clipping
// by segment intersection + corner containment, angular sort about the centroid,
// shoelace area; EPS 1e-8, MARGIN 1e-2, IoU = s / max(sa + sb - s, EPS), suppress iff
// IoU > thresh) because keep-lists must be bit-exact.  Everything around it is new:
//   * per-box quantities (rotated corners, cos/sin of -heading, margins, area) are
//     computed ONCE per box by a prep kernel instead of once per pair;
//   * a conservative bounding-circle test rejects far pairs to an exact 0 (the
//     reference's result for disjoint boxes) and the survivors of a 64x64 tile are
//     compacted with warp ballots into a shared-memory queue, so the expensive path
//     runs on dense warps;
//   * only tiles on or above the diagonal are evaluated (the sweep never reads the rest);
//   * the greedy sweep runs on the device, batched over frames: no 2 MiB mask copy, no
//     host loop, no cudaMalloc/cudaFree per call.
#include <math.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tsm {

constexpr float kEps = 1e-8f;
constexpr float kMargin = 1e-2f;

struct __align__(16) BoxPrep {
    float cx, cy;        // centre
    float px[4], py[4];  // rotated corners, reference order
    float ci, si;        // cosf(-heading), sinf(-heading)
    float mx, my;        // dx/2 + MARGIN, dy/2 + MARGIN
    float area;          // dx * dy
    float rad;           // conservative bounding radius (fast reject only)
    float x1, y1, x2, y2;  // axis-aligned extents (nms_normal)
};
static_assert(sizeof(BoxPrep) == 80, "BoxPrep is 20 floats");

__host__ __device__ __forceinline__ void prep_box(const float* __restrict__ box, BoxPrep& o) {
    const float bx = box[0], by = box[1], dx = box[3], dy = box[4], ang = box[6];
    const float hx = dx / 2, hy = dy / 2;
    const float x1 = bx - hx, y1 = by - hy;
    const float x2 = bx + hx, y2 = by + hy;
    const float c = cosf(ang), s = sinf(ang);
    const float qx[4] = {x1, x2, x2, x1};
    const float qy[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float nx = (qx[k] - bx) * c + (qy[k] - by) * (-s) + bx;
        const float ny = (qx[k] - bx) * s + (qy[k] - by) * c + by;
        o.px[k] = nx;
        o.py[k] = ny;
    }
    o.cx = bx;
    o.cy = by;
    o.ci = cosf(-ang);
    o.si = sinf(-ang);
    o.mx = dx / 2 + kMargin;
    o.my = dy / 2 + kMargin;
    o.area = dx * dy;
    o.rad = 0.5f * sqrtf(dx * dx + dy * dy) * 1.0001f + 0.03f + 4e-6f * (fabsf(bx) + fabsf(by));
    o.x1 = x1; o.y1 = y1; o.x2 = x2; o.y2 = y2;
}

// true when the two rectangles (grown by the containment margin) certainly do not touch:
// the reference then finds no intersection and no contained corner and returns exactly 0.
__host__ __device__ __forceinline__ bool surely_disjoint(const BoxPrep& a, const BoxPrep& b) {
    const float ddx = a.cx - b.cx, ddy = a.cy - b.cy;
    const float r = a.rad + b.rad;
    return ddx * ddx + ddy * ddy > r * r;  // NaN compares false -> full path
}

// Second, tighter conservative test for pairs whose bounding circles touch: separating axis theorem on
// the margin-grown rectangles (+4 cm slack, far above the f32 error of the reference's corner arithmetic).
// Separated rectangles have no edge intersection and no corner within MARGIN of the other box, so the
// reference returns exactly 0 for them.
__host__ __device__ __forceinline__ bool surely_separated(const BoxPrep& a, const BoxPrep& b) {
    const float dx = b.cx - a.cx, dy = b.cy - a.cy;
    const float c = fabsf(a.ci * b.ci + a.si * b.si), s = fabsf(a.si * b.ci - a.ci * b.si);
    const float slack = 0.04f + 4e-6f * (fabsf(a.cx) + fabsf(a.cy) + fabsf(b.cx) + fabsf(b.cy));
    // axes of a: u = (ci, -si), v = (si, ci)   (ci, si = cos, sin of -heading)
    const float au = fabsf(dx * a.ci - dy * a.si), av = fabsf(dx * a.si + dy * a.ci);
    if (au > a.mx + b.mx * c + b.my * s + slack) return true;
    if (av > a.my + b.mx * s + b.my * c + slack) return true;
    const float bu = fabsf(dx * b.ci - dy * b.si), bv = fabsf(dx * b.si + dy * b.ci);
    if (bu > b.mx + a.mx * c + a.my * s + slack) return true;
    if (bv > b.my + a.mx * s + a.my * c + slack) return true;
    return false;  // NaNs compare false everywhere -> full path
}

__host__ __device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x, float p0y) {
    return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y);
}

__host__ __device__ __forceinline__ bool corner_inside(const BoxPrep& bx, float x, float y) {
    const float rx = (x - bx.cx) * bx.ci + (y - bx.cy) * (-bx.si);
    const float ry = (x - bx.cx) * bx.si + (y - bx.cy) * bx.ci;
    return fabsf(rx) < bx.mx && fabsf(ry) < bx.my;
}

// Intersection area of rotated rectangles A (first argument of the reference's
// box_overlap) and B.
__host__ __device__ inline float overlap_area(const BoxPrep& A, const BoxPrep& B) {
    float vx[16], vy[16];
    int cnt = 0;
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float p0x = A.px[i], p0y = A.py[i];
        const float p1x = A.px[(i + 1) & 3], p1y = A.py[(i + 1) & 3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float q0x = B.px[j], q0y = B.py[j];
            const float q1x = B.px[(j + 1) & 3], q1y = B.py[(j + 1) & 3];
            const bool boxes_touch = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                                     fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
            if (!boxes_touch) continue;
            const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
            const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
            const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
            const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
            if (!(s1 * s2 > 0 && s3 * s4 > 0)) continue;
            const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
            float ix, iy;
            if (fabsf(s5 - s1) > kEps) {
                ix = (s5 * q0x - s1 * q1x) / (s5 - s1);
                iy = (s5 * q0y - s1 * q1y) / (s5 - s1);
            } else {
                const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
                const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
                const float D = a0 * b1 - a1 * b0;
                ix = (b0 * c1 - b1 * c0) / D;
                iy = (a1 * c0 - a0 * c1) / D;
            }
            sx = sx + ix;
            sy = sy + iy;
            vx[cnt] = ix;
            vy[cnt] = iy;
            ++cnt;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (corner_inside(A, B.px[k], B.py[k])) {
            sx = sx + B.px[k];
            sy = sy + B.py[k];
            vx[cnt] = B.px[k];
            vy[cnt] = B.py[k];
            ++cnt;
        }
        if (corner_inside(B, A.px[k], A.py[k])) {
            sx = sx + A.px[k];
            sy = sy + A.py[k];
            vx[cnt] = A.px[k];
            vy[cnt] = A.py[k];
            ++cnt;
        }
    }
    if (cnt < 3) return 0.f;  // fewer than 3 vertices: the shoelace sum below is exactly 0
    sx /= cnt;
    sy /= cnt;
    // angular bubble sort: same comparator and swap sequence as the reference, with each
    // vertex's atan2f evaluated once (it is a pure function of the vertex and the centre)
    float ang[16];
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(vy[k] - sy, vx[k] - sx);
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                float t;
                t = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = t;
                t = vx[i]; vx[i] = vx[i + 1]; vx[i + 1] = t;
                t = vy[i]; vy[i] = vy[i + 1]; vy[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ax = vx[k] - vx[0], ay = vy[k] - vy[0];
        const float bx = vx[k + 1] - vx[0], by = vy[k + 1] - vy[0];
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0;
}

__host__ __device__ __forceinline__ float iou_rotated(const BoxPrep& A, const BoxPrep& B) {
    const float so = overlap_area(A, B);
    return so / fmaxf(A.area + B.area - so, kEps);
}

__device__ __forceinline__ float iou_axis(const BoxPrep& a, const BoxPrep& b) {
    const float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
    const float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h;
    return inter / fmaxf(a.area + b.area - inter, kEps);
}

// ------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128) prep_boxes_kernel(int total, const float* __restrict__ boxes, int stride,
                                                         BoxPrep* __restrict__ out) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= total) return;
    BoxPrep p;
    prep_box(boxes + (size_t)i * stride, p);
    out[i] = p;
}

// Dense (N,M) overlap or IoU matrix.  64x64 tile per CTA, survivors of the fast reject
// are queued and evaluated by dense warps; rejected pairs are written as exact zeros.
template <bool IOU>
__global__ void __launch_bounds__(256)
    pair_matrix_kernel(int na, const BoxPrep* __restrict__ A, int nb, const BoxPrep* __restrict__ B,
                       float* __restrict__ out) {
    __shared__ BoxPrep sa[64], sb[64];
    __shared__ unsigned short queue[4096];
    __shared__ int qn;
    const int tid = threadIdx.x, lane = tid & 31;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int nr = min(64, na - r0), nc = min(64, nb - c0);
    for (int e = tid; e < 64 * 20; e += 256) {
        const int bi = e / 20, w = e - bi * 20;
        if (bi < nr) reinterpret_cast<float*>(&sa[bi])[w] = reinterpret_cast<const float*>(&A[r0 + bi])[w];
        if (bi < nc) reinterpret_cast<float*>(&sb[bi])[w] = reinterpret_cast<const float*>(&B[c0 + bi])[w];
    }
    if (tid == 0) qn = 0;
    __syncthreads();
    for (int p = tid; p < 4096; p += 256) {
        const int r = p >> 6, c = p & 63;
        const bool in = r < nr && c < nc;
        const bool heavy = in && !surely_disjoint(sa[r], sb[c]) && !surely_separated(sa[r], sb[c]);
        if (in && !heavy) out[(size_t)(r0 + r) * nb + c0 + c] = 0.f;
        const unsigned bal = __ballot_sync(FULL, heavy);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&qn, __popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (heavy) queue[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)p;
        }
    }
    __syncthreads();
    const int total = qn;
    for (int q = tid; q < total; q += 256) {
        const int p = queue[q];
        const int r = p >> 6, c = p & 63;
        const float v = IOU ? iou_rotated(sa[r], sb[c]) : overlap_area(sa[r], sb[c]);
        out[(size_t)(r0 + r) * nb + c0 + c] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Spatial-grid candidate generation (rotated NMS): instead of testing all N(N-1)/2 pairs, boxes are
// binned by centre into a uniform grid whose cell is at least the largest bounding-circle diameter, so two
// boxes whose circles touch lie in the same or adjacent cells.  Pairs outside the 3x3 neighbourhood are
// exactly the pairs the bounding-circle reject would have zeroed, so the mask is unchanged.
struct NmsGrid {
    float x0, y0, inv_cell;
    int gx, gy;
    int fallback;  // 1: non-finite boxes or a degenerate extent -> the all-pairs tile kernel handles this frame
};
constexpr int NMS_GMAX = 64;  // grid is at most 64 x 64 cells

__device__ __forceinline__ int grid_cell(const NmsGrid& g, float cx, float cy, int& ix, int& iy) {
    ix = min(g.gx - 1, max(0, (int)((cx - g.x0) * g.inv_cell)));
    iy = min(g.gy - 1, max(0, (int)((cy - g.y0) * g.inv_cell)));
    return iy * g.gx + ix;
}

// one CTA (1024 threads) per frame: extents -> cell size -> counting sort of box indices by cell
__global__ void __launch_bounds__(1024)
    nms_grid_build_kernel(int nmax, const int* __restrict__ counts, const BoxPrep* __restrict__ prep,
                          NmsGrid* __restrict__ grids, int* __restrict__ cell_start, int* __restrict__ sorted) {
    __shared__ int cnt[NMS_GMAX * NMS_GMAX];
    __shared__ float red[5][32];
    __shared__ int bad_s;
    __shared__ NmsGrid g_s;
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const int n = counts ? min(counts[f], nmax) : nmax;
    const BoxPrep* P = prep + (size_t)f * nmax;
    int* cs = cell_start + (size_t)f * (NMS_GMAX * NMS_GMAX + 1);
    int* so = sorted + (size_t)f * nmax;
    if (tid == 0) bad_s = 0;
    __syncthreads();
    float xmin = 3.4e38f, xmax = -3.4e38f, ymin = 3.4e38f, ymax = -3.4e38f, rmax = 0.f;
    int bad = 0;
    for (int i = tid; i < n; i += 1024) {
        const float cx = P[i].cx, cy = P[i].cy, r = P[i].rad;
        if (!(fabsf(cx) < 1e30f) || !(fabsf(cy) < 1e30f) || !(r < 1e30f) || !(r >= 0.f)) bad = 1;
        xmin = fminf(xmin, cx); xmax = fmaxf(xmax, cx);
        ymin = fminf(ymin, cy); ymax = fmaxf(ymax, cy);
        rmax = fmaxf(rmax, r);
    }
    for (int o = 16; o >= 1; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(FULL, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(FULL, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
        rmax = fmaxf(rmax, __shfl_xor_sync(FULL, rmax, o));
    }
    if (bad) atomicOr(&bad_s, 1);
    if (lane == 0) { red[0][warp] = xmin; red[1][warp] = xmax; red[2][warp] = ymin; red[3][warp] = ymax; red[4][warp] = rmax; }
    for (int i = tid; i < NMS_GMAX * NMS_GMAX; i += 1024) cnt[i] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; ++w) {
            red[0][0] = fminf(red[0][0], red[0][w]); red[1][0] = fmaxf(red[1][0], red[1][w]);
            red[2][0] = fminf(red[2][0], red[2][w]); red[3][0] = fmaxf(red[3][0], red[3][w]);
            red[4][0] = fmaxf(red[4][0], red[4][w]);
        }
        NmsGrid g;
        const float ex = red[1][0] - red[0][0], ey = red[3][0] - red[2][0];
        float cell = fmaxf(2.f * red[4][0], fmaxf(ex, ey) / (float)NMS_GMAX) * 1.0001f;
        g.fallback = (bad_s || n <= 0 || !(cell > 0.f) || !(cell < 1e30f)) ? 1 : 0;
        if (g.fallback) cell = 1.f;
        g.x0 = red[0][0];
        g.y0 = red[2][0];
        g.inv_cell = 1.f / cell;
        g.gx = min(NMS_GMAX, max(1, (int)(ex * g.inv_cell) + 1));
        g.gy = min(NMS_GMAX, max(1, (int)(ey * g.inv_cell) + 1));
        if (g.fallback) { g.gx = 1; g.gy = 1; g.x0 = 0.f; g.y0 = 0.f; }
        g_s = g;
        grids[f] = g;
    }
    __syncthreads();
    const NmsGrid g = g_s;
    if (g.fallback) return;
    for (int i = tid; i < n; i += 1024) {
        int ix, iy;
        atomicAdd(&cnt[grid_cell(g, P[i].cx, P[i].cy, ix, iy)], 1);
    }
    __syncthreads();
    // exclusive scan of the 4096 counters: 4 per thread, warp scan, scan of warp totals
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = cnt[tid * 4 + k]; sum += v[k]; }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;  // inclusive
    }
    __syncthreads();
    int base = (warp ? warp_tot[warp - 1] : 0) + inc - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cs[tid * 4 + k] = base;
        cnt[tid * 4 + k] = base;  // becomes the scatter cursor
        base += v[k];
    }
    if (tid == 1023) cs[NMS_GMAX * NMS_GMAX] = base;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        int ix, iy;
        const int c = grid_cell(g, P[i].cx, P[i].cy, ix, iy);
        so[atomicAdd(&cnt[c], 1)] = i;
    }
}

// Persistent warps: warp w takes boxes i = w, w + W, ...; its lanes sweep the candidates j > i of the 3x3
// cell neighbourhood, apply the two conservative rejects, and push survivors into a per-warp queue that is
// drained 32 pairs at a time, so the exact IoU runs with all lanes busy and no CTA-level barrier at all.
__global__ void __launch_bounds__(256)
    nms_grid_pairs_kernel(int nmax, const int* __restrict__ counts, float thresh, const BoxPrep* __restrict__ prep,
                          const NmsGrid* __restrict__ grids, const int* __restrict__ cell_start,
                          const int* __restrict__ sorted, unsigned long long* __restrict__ mask) {
    __shared__ unsigned int wq[8][96];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.y;
    const NmsGrid g = grids[f];
    if (g.fallback) return;
    const int n = counts ? min(counts[f], nmax) : nmax;
    const int cbmax = divup(nmax, 64);
    const BoxPrep* P = prep + (size_t)f * nmax;
    unsigned long long* M = mask + (size_t)f * nmax * cbmax;
    const int* cs = cell_start + (size_t)f * (NMS_GMAX * NMS_GMAX + 1);
    const int* so = sorted + (size_t)f * nmax;
    const bool all_heavy = !(thresh >= 0.f);
    unsigned int* q = wq[warp];
    int qn = 0;  // warp-uniform

    auto eval32 = [&](int count) {  // exact IoU for queue entries [qn - count, qn)
        if (lane < count) {
            const unsigned int e = q[qn - count + lane];
            const int i = (int)(e >> 16), j = (int)(e & 0xffffu);
            const BoxPrep a = P[i], b = P[j];
            if (iou_rotated(a, b) > thresh) atomicOr(&M[(size_t)i * cbmax + (j >> 6)], 1ull << (j & 63));
        }
        __syncwarp();
        qn -= count;
    };

    // Every unordered pair of boxes in the same or adjacent cells is generated exactly once: the warp walks the
    // boxes in CELL order (position k of the sorted list) and pairs box so[k] with (1) the rest of its own cell
    // and the cell to its right -- one contiguous run of the sorted list starting at k + 1 -- and (2) the three
    // cells below it, another contiguous run.  Half the candidates of a full 3x3 sweep, and none is discarded
    // for being on the wrong side of the diagonal.
    const int wstride = gridDim.x * 8;
    for (int k = blockIdx.x * 8 + warp; k < n; k += wstride) {
        const int i = so[k];
        const BoxPrep a = P[i];
        int ix, iy;
        grid_cell(g, a.cx, a.cy, ix, iy);
#pragma unroll 1
        for (int run = 0; run < 2; ++run) {
            int s0, s1;
            if (run == 0) {
                s0 = k + 1;
                s1 = cs[iy * g.gx + min(ix + 1, g.gx - 1) + 1];
            } else {
                if (iy + 1 >= g.gy) break;
                s0 = cs[(iy + 1) * g.gx + max(ix - 1, 0)];
                s1 = cs[(iy + 1) * g.gx + min(ix + 1, g.gx - 1) + 1];
            }
            for (int k0 = s0; k0 < s1; k0 += 32) {
                const int kk = k0 + lane;
                bool heavy = false;
                int j = 0;
                if (kk < s1) {
                    j = so[kk];
                    const BoxPrep* pb = P + j;
                    if (all_heavy) {
                        heavy = true;
                    } else {
                        BoxPrep b;  // only the fields the rejects read
                        b.cx = pb->cx; b.cy = pb->cy; b.rad = pb->rad;
                        b.ci = pb->ci; b.si = pb->si; b.mx = pb->mx; b.my = pb->my;
                        heavy = !surely_disjoint(a, b) && !surely_separated(a, b);
                    }
                }
                const unsigned bal = __ballot_sync(FULL, heavy);
                if (bal) {
                    if (heavy)
                        q[qn + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)min(i, j) << 16) | (unsigned)max(i, j);
                    qn += __popc(bal);
                    __syncwarp();
                    if (qn >= 32) eval32(32);
                }
            }
        }
    }
    if (qn > 0) eval32(qn);
}

// Suppression words for the tiles on/above the diagonal.
//   prep (F, nmax) per-frame prepared boxes in score order, counts (F) or null (= nmax)
//   mask (F, nmax, cbmax) 64-bit words, ZEROED by the launcher; bits are OR-ed in.
// Persistent CTAs (gridDim.x per frame) walk the frame's upper-triangular 64x64 tiles.  Each tile's
// 4096 pairs go through two conservative rejects (bounding circles, then separating axes); survivors
// are appended -- as (row, col) box indices -- to a shared-memory queue that is kept ACROSS tiles and
// drained by all 256 threads whenever it holds >= 2048 pairs, so the ~2000-instruction exact IoU always
// runs on full warps no matter how sparse the overlaps are.
constexpr int NMS_Q = 4096 + 2048;

template <bool NORMAL>
__global__ void __launch_bounds__(256)
    nms_mask_kernel(int nmax, const int* __restrict__ counts, float thresh, const BoxPrep* __restrict__ prep,
                    unsigned long long* __restrict__ mask, const NmsGrid* __restrict__ gate) {
    __shared__ BoxPrep srow[64], scol[64];
    __shared__ unsigned int queue[NMS_Q];
    __shared__ int qn;
    const int tid = threadIdx.x, lane = tid & 31;
    const int f = blockIdx.y;
    if (gate && !gate[f].fallback) return;  // this frame is handled by the spatial-grid kernel
    const int n = counts ? min(counts[f], nmax) : nmax;
    const int cbmax = divup(nmax, 64);
    const int cb = divup(n, 64);
    const long long ntiles = (long long)cb * (cb + 1) / 2;
    const BoxPrep* P = prep + (size_t)f * nmax;
    unsigned long long* M = mask + (size_t)f * nmax * cbmax;
    const bool all_heavy = !(thresh >= 0.f);  // negative/NaN threshold: a zero IoU may still suppress
    if (tid == 0) qn = 0;
    __syncthreads();

    auto drain = [&]() {  // evaluate every queued pair exactly; all threads busy
        const int total = qn;
        for (int q = tid; q < total; q += 256) {
            const unsigned int e = queue[q];
            const int i = (int)(e >> 16), j = (int)(e & 0xffffu);
            const BoxPrep a = P[i], b = P[j];
            if (iou_rotated(a, b) > thresh) atomicOr(&M[(size_t)i * cbmax + (j >> 6)], 1ull << (j & 63));
        }
        __syncthreads();
        if (tid == 0) qn = 0;
        __syncthreads();
    };

    // tile t -> (rb, cbk), rb <= cbk: row rb owns (cb - rb) tiles
    int rb = 0;
    long long row_first = 0;  // linear id of tile (rb, rb)
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        while (t >= row_first + (cb - rb)) {
            row_first += cb - rb;
            ++rb;
        }
        const int cbk = rb + (int)(t - row_first);
        const int r0 = rb * 64, c0 = cbk * 64;
        const int nr = min(64, n - r0), nc = min(64, n - c0);
        for (int e = tid; e < 64 * 20; e += 256) {
            const int bi = e / 20, w = e - bi * 20;
            if (bi < nr) reinterpret_cast<float*>(&srow[bi])[w] = reinterpret_cast<const float*>(&P[r0 + bi])[w];
            if (bi < nc) reinterpret_cast<float*>(&scol[bi])[w] = reinterpret_cast<const float*>(&P[c0 + bi])[w];
        }
        __syncthreads();
        const bool diag = rb == cbk;
        for (int p = tid; p < 4096; p += 256) {
            const int r = p >> 6, c = p & 63;
            const bool in = r < nr && c < nc && (!diag || c > r);
            if (NORMAL) {
                if (in && iou_axis(srow[r], scol[c]) > thresh)
                    atomicOr(&M[(size_t)(r0 + r) * cbmax + cbk], 1ull << c);
            } else {
                bool heavy = in;
                if (in && !all_heavy) heavy = !surely_disjoint(srow[r], scol[c]) && !surely_separated(srow[r], scol[c]);
                const unsigned bal = __ballot_sync(FULL, heavy);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&qn, __popc(bal));
                    base = __shfl_sync(FULL, base, 0);
                    if (heavy) queue[base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)(r0 + r) << 16) | (unsigned)(c0 + c);
                }
            }
        }
        __syncthreads();
        if (!NORMAL && qn >= 2048) drain();  // uniform: qn is stable after the barrier
    }
    if (!NORMAL) drain();
}

// Lazy rotated NMS: mask rows are only ever read for boxes the greedy sweep KEEPS, and only their bits for
// boxes that are still alive.  So instead of building the whole mask and sweeping it afterwards, one CTA per
// frame walks the score-ordered boxes in blocks of 64 and, per block, (1) computes the suppression rows of
// the boxes not yet removed -- spatial-grid candidates, the two conservative rejects, exact IoU only against
// boxes that are themselves still alive -- into shared memory, (2) resolves the block's diagonal serially
// (iou3d_nms.cpp:116-131), (3) ORs the kept rows into the removal words of the later blocks.  With the
// clustered proposals of a detector (each kept box removes ~20 others) that is ~12x fewer exact IoUs than the
// full mask, and the mask never touches HBM.  Keep-lists are identical: a bit that is not computed is a bit the
// reference's sweep never reads or whose target is already removed.
// zero the mask of the frames the all-pairs path will serve (the others never touch theirs)
__global__ void __launch_bounds__(256)
    nms_clear_flagged_kernel(unsigned long long* __restrict__ mask, size_t words_per_frame, const NmsGrid* __restrict__ grids) {
    const int f = blockIdx.y;
    if (!grids[f].fallback) return;
    unsigned long long* M = mask + (size_t)f * words_per_frame;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < words_per_frame; e += (size_t)gridDim.x * 256) M[e] = 0ull;
}

// One CTA per frame.  The kernel is bound by the latency of its exact IoUs (17 % issue slots; the early blocks, where
// nothing has been removed yet, hold most of the work: ~40k warp-cycles per alive row).  Around them:
//  * the frame's grid (cell_start, box indices in cell order as u16) is staged in shared memory once, and the BoxPrep
//    records of the NEXT block's 64 rows arrive by cp.async while the current block runs: a chunk of candidates pays ONE L2
//    round trip (their reject fields) instead of four dependent ones (row record, cell_start, indices, fields);
//  * the candidates of the three cell rows around a box form ONE range that the warp walks 64 at a time, the two loads of
//    a lane issued before either is used;
//  * only the alive rows are zeroed, and the diagonal is resolved over the bits still standing (find-first-set).
// Measured (16 frames x 4096 proposals, IoU 0.01, whole nms_gpu_batch): 0.78 -> 0.70 ms.  Tried and measured slower: dealing
// 64-candidate chunks instead of rows to the warps (0.78 ms), four loads in flight per lane (0.87 ms).
// 512 threads x 109 registers filled an SM's register file, so a frame's NMS held a whole SM for ~0.6 ms; with 256 threads
// other steps' kernels share the SM -- the pipelined step measured 30.8k -> 31.5k frames/s (128 threads: 31.1k).
#ifndef TSM_NMSL_THREADS
#define TSM_NMSL_THREADS 256
#endif
constexpr int NMSL_THREADS = TSM_NMSL_THREADS;
constexpr int NMSL_WARPS = NMSL_THREADS / 32;

// dynamic shared memory of nms_lazy_kernel: remv[cbmax] | rows[64][cbmax] (u64) | two blocks of 64 BoxPrep | cell_start | u16 indices
__host__ __device__ inline size_t nms_lazy_smem_bytes(int nmax) {
    const size_t cbmax = (size_t)(nmax + 63) / 64;
    size_t b = ((65 * cbmax + 1) & ~(size_t)1) * sizeof(unsigned long long) + 2 * 64 * sizeof(BoxPrep) +
               (NMS_GMAX * NMS_GMAX + 1) * sizeof(int) + (size_t)nmax * sizeof(unsigned short);
    return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(NMSL_THREADS)
    nms_lazy_kernel(int nmax, const int* __restrict__ counts, float thresh, const BoxPrep* __restrict__ prep,
                    const NmsGrid* __restrict__ grids, const int* __restrict__ cell_start,
                    const int* __restrict__ sorted, long long* __restrict__ keep, int* __restrict__ num_keep) {
    extern __shared__ __align__(16) unsigned long long lz[];
    __shared__ unsigned int wq[NMSL_WARPS][96];
    __shared__ unsigned long long kept_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const NmsGrid g = grids[f];
    if (g.fallback) return;  // the all-pairs path serves this frame
    const int n = counts ? min(counts[f], nmax) : nmax;
    const int cbmax = divup(nmax, 64);
    const int cb = divup(n, 64);
    const BoxPrep* P = prep + (size_t)f * nmax;
    const int* cs = cell_start + (size_t)f * (NMS_GMAX * NMS_GMAX + 1);
    const int* so = sorted + (size_t)f * nmax;
    long long* K = keep + (size_t)f * nmax;
    unsigned long long* remv = lz;
    unsigned long long* rows = lz + cbmax;
    BoxPrep* const pa = reinterpret_cast<BoxPrep*>(lz + (((size_t)65 * cbmax + 1) & ~(size_t)1));  // [2][64] rows of block b / b + 1 (16-byte aligned)
    int* const cs_s = reinterpret_cast<int*>(pa + 128);                        // [gx * gy + 1]
    unsigned short* const so_s = reinterpret_cast<unsigned short*>(cs_s + NMS_GMAX * NMS_GMAX + 1);  // [n]
    unsigned int* q = wq[warp];
    // 16-byte pieces of the 64 BoxPrep records of block `blk` -> pa[blk & 1] (asynchronous; rows past the frame's
    // allocation are never read)
    auto fetch_rows = [&](int blk) {
        const char* src = reinterpret_cast<const char*>(P + (size_t)blk * 64);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pa + (blk & 1) * 64);
        const int pieces = min(64, nmax - blk * 64) * (int)(sizeof(BoxPrep) / 16);
        for (int e = tid; e < pieces; e += NMSL_THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * e), "l"(src + 16 * e) : "memory");
    };
    if (cb > 0) fetch_rows(0);
    for (int j = tid; j < cb; j += NMSL_THREADS) remv[j] = 0ull;
    for (int j = tid; j <= g.gx * g.gy; j += NMSL_THREADS) cs_s[j] = cs[j];
    for (int j = tid; j < n; j += NMSL_THREADS) so_s[j] = (unsigned short)so[j];
    asm volatile("cp.async.wait_all;" ::: "memory");
    int base = 0;
    __syncthreads();
    for (int b = 0; b < cb; ++b) {
        const int nrows = min(64, n - b * 64);
        unsigned long long alive = ~remv[b];
        if (nrows < 64) alive &= (1ull << nrows) - 1ull;
        if (b + 1 < cb) fetch_rows(b + 1);
        const BoxPrep* const pb_rows = pa + (b & 1) * 64;
        // words b .. cb-1 of the alive rows start from zero
        for (unsigned long long rest = alive; rest; rest &= rest - 1ull) {
            const int r = __ffsll((long long)rest) - 1;
            for (int j = b + tid; j < cb; j += NMSL_THREADS) rows[(size_t)r * cbmax + j] = 0ull;
        }
        __syncthreads();
        // ---- (1) rows of the alive boxes of this block: warp w takes the w-th, (w+8)-th, ... alive row
        {
            int qn = 0;  // warp-uniform
            auto eval32 = [&](int count) {
                if (lane < count) {
                    const unsigned int e = q[qn - count + lane];
                    const int r = (int)(e >> 16), j = (int)(e & 0xffffu);
                    const BoxPrep a = pb_rows[r], bb = P[j];
                    if (iou_rotated(a, bb) > thresh) atomicOr(&rows[(size_t)r * cbmax + (j >> 6)], 1ull << (j & 63));
                }
                __syncwarp();
                qn -= count;
            };
            unsigned long long rest = alive;
            for (int skip = 0; rest; ++skip) {
                const int r = __ffsll((long long)rest) - 1;
                rest &= rest - 1ull;
                if ((skip & (NMSL_WARPS - 1)) != warp) continue;
                const int i = b * 64 + r;
                BoxPrep a;  // only the fields the rejects read
                a.cx = pb_rows[r].cx; a.cy = pb_rows[r].cy; a.rad = pb_rows[r].rad;
                a.ci = pb_rows[r].ci; a.si = pb_rows[r].si; a.mx = pb_rows[r].mx; a.my = pb_rows[r].my;
                int ix, iy;
                grid_cell(g, a.cx, a.cy, ix, iy);
                // the candidates: cells (ix-1 .. ix+1) of the grid rows iy-1 .. iy+1 = three runs of the cell-ordered
                // index list, walked as ONE range
                const int xa = max(ix - 1, 0), xb = min(ix + 1, g.gx - 1);
                int s0[3], len[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int yy = iy + d - 1;
                    const bool in = yy >= 0 && yy < g.gy;
                    s0[d] = in ? cs_s[yy * g.gx + xa] : 0;
                    len[d] = in ? cs_s[yy * g.gx + xb + 1] - s0[d] : 0;
                }
                const int l01 = len[0] + len[1], tot = l01 + len[2];
                for (int t0 = 0; t0 < tot; t0 += 64) {
                    int j[2];
                    bool want[2], heavy[2];
                    BoxPrep c[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int t = t0 + 32 * u + lane;
                        j[u] = 0;
                        want[u] = false;
                        heavy[u] = false;
                        if (t < tot) {
                            const int k = t < len[0] ? s0[0] + t : (t < l01 ? s0[1] + (t - len[0]) : s0[2] + (t - l01));
                            j[u] = so_s[k];
                            // later in score order and not removed yet: the only bits the sweep can still use
                            want[u] = j[u] > i && !((remv[j[u] >> 6] >> (j[u] & 63)) & 1ull);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {  // both of a lane's loads are issued before either is used
                        if (want[u]) {
                            const BoxPrep* pb = P + j[u];
                            c[u].cx = pb->cx; c[u].cy = pb->cy; c[u].rad = pb->rad;
                            c[u].ci = pb->ci; c[u].si = pb->si; c[u].mx = pb->mx; c[u].my = pb->my;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (want[u]) heavy[u] = !surely_disjoint(a, c[u]) && !surely_separated(a, c[u]);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const unsigned bal = __ballot_sync(FULL, heavy[u]);
                        if (bal) {
                            if (heavy[u]) q[qn + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)r << 16) | (unsigned)j[u];
                            qn += __popc(bal);
                            __syncwarp();
                            if (qn >= 32) eval32(32);
                        }
                    }
                }
            }
            if (qn > 0) eval32(qn);
        }
        __syncthreads();
        // ---- (2) the block's diagonal, serially (iou3d_nms.cpp:116-131): the lowest box still standing is kept and
        // its row removes others; rows only hold bits of LATER boxes, so this is the ascending scan
        if (tid == 0) {
            unsigned long long cur = ~alive, kept = 0ull;
            while (~cur) {
                const int r = __ffsll((long long)~cur) - 1;
                kept |= 1ull << r;
                cur |= rows[(size_t)r * cbmax + b] | (1ull << r);
            }
            kept_s = kept;
        }
        __syncthreads();
        const unsigned long long kept = kept_s;
        if (tid < 64 && ((kept >> tid) & 1ull)) K[base + __popcll(kept & ((1ull << tid) - 1ull))] = (long long)(b * 64 + tid);
        base += __popcll(kept);
        // ---- (3) kept rows -> removal words of the later blocks (one thread per word: no atomics)
        for (int j = b + 1 + tid; j < cb; j += NMSL_THREADS) {
            unsigned long long acc = 0ull, kk = kept;
            while (kk) {
                const int r = __ffsll((long long)kk) - 1;
                kk &= kk - 1ull;
                acc |= rows[(size_t)r * cbmax + j];
            }
            remv[j] |= acc;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");  // the next block's rows have landed (this thread's pieces)
        __syncthreads();
    }
    if (tid == 0) num_keep[f] = base;
}

// Greedy sweep (iou3d_nms.cpp:116-131) on the device, one CTA (256 threads) per frame.
//   keep (F, nmax) int64: kept indices (into the score-sorted boxes) in ascending order
//   num_keep (F) int32
// Per 64-box block: one thread resolves the block against its diagonal word with the 64 words preloaded into
// registers (a pure ALU chain, ~12 cycles per box), then all threads OR the rows of the kept boxes into the
// removal words of the later blocks.
__global__ void __launch_bounds__(256)
    nms_sweep_kernel(int nmax, const int* __restrict__ counts, const unsigned long long* __restrict__ mask,
                     long long* __restrict__ keep, int* __restrict__ num_keep, const NmsGrid* __restrict__ gate) {
    extern __shared__ unsigned long long remv[];  // cbmax words
    __shared__ unsigned long long diagw[64];
    __shared__ unsigned long long kept_s;
    const int tid = threadIdx.x;
    const int f = blockIdx.x;
    if (gate && !gate[f].fallback) return;  // nms_lazy_kernel answered this frame
    const int n = counts ? min(counts[f], nmax) : nmax;
    const int cbmax = divup(nmax, 64);
    const int cb = divup(n, 64);
    const unsigned long long* M = mask + (size_t)f * nmax * cbmax;
    long long* K = keep + (size_t)f * nmax;
    for (int j = tid; j < cb; j += 256) remv[j] = 0ull;
    int base = 0;
    __syncthreads();
    for (int b = 0; b < cb; ++b) {
        const int rows = min(64, n - b * 64);
        if (tid < 64) diagw[tid] = tid < rows ? M[(size_t)(b * 64 + tid) * cbmax + b] : ~0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long dw[64];
#pragma unroll
            for (int r = 0; r < 64; ++r) dw[r] = diagw[r];
            unsigned long long cur = remv[b], kept = 0ull;
            if (rows < 64) cur |= ~0ull << rows;  // rows beyond the frame are never kept
#pragma unroll
            for (int r = 0; r < 64; ++r) {
                const bool take = !((cur >> r) & 1ull);
                kept |= take ? (1ull << r) : 0ull;
                cur |= take ? dw[r] : 0ull;
            }
            kept_s = kept;
        }
        __syncthreads();
        const unsigned long long kept = kept_s;
        if (tid < 64 && ((kept >> tid) & 1ull)) {
            const int pos = base + __popcll(kept & ((1ull << tid) - 1ull));
            K[pos] = (long long)(b * 64 + tid);
        }
        base += __popcll(kept);
        // OR the rows of the boxes kept in this block into remv[j], j > b: thread = (column word, 16-row group)
        {
            const int ncol = cb - b - 1;
            const int g = tid >> 6;  // 0..3
            const int jc = tid & 63;
            for (int j0 = 0; j0 < ncol; j0 += 64) {
                const int j = b + 1 + j0 + jc;
                if (j < cb) {
                    unsigned long long acc = 0ull;
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        const int r = g * 16 + rr;
                        if ((kept >> r) & 1ull) acc |= M[(size_t)(b * 64 + r) * cbmax + j];
                    }
                    if (acc) atomicOr(&remv[j], acc);
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) num_keep[f] = base;
}

}  // namespace tsm

// --------------------------------------------------------------------------------- host
namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int nms_batch_impl(bool normal, int frames, int nmax, const float* boxes, int box_stride, const int* counts,
                   float thresh, long long* keep, int* num_keep, cudaStream_t s) {
    if (frames <= 0) return TSM_OK;
    if (nmax <= 0) {
        TSM_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int) * frames, s));
        return TSM_OK;
    }
    if (frames > 65535 || box_stride < 7) return TSM_ERR_INVALID;
    const int cbmax = tsm::divup(nmax, 64);
    const size_t prep_bytes = align_up((size_t)frames * nmax * sizeof(tsm::BoxPrep), 256);
    const size_t mask_bytes = align_up((size_t)frames * nmax * cbmax * sizeof(unsigned long long), 256);
    const size_t grid_bytes = align_up((size_t)frames * sizeof(tsm::NmsGrid), 256);
    const size_t cs_bytes = align_up((size_t)frames * (tsm::NMS_GMAX * tsm::NMS_GMAX + 1) * sizeof(int), 256);
    const size_t so_bytes = align_up((size_t)frames * nmax * sizeof(int), 256);
    void* scratch = nullptr;
    int rc = tsm_scratch_get(0, prep_bytes + mask_bytes + grid_bytes + cs_bytes + so_bytes, s, &scratch);
    if (rc != TSM_OK) return rc;
    char* sp = (char*)scratch;
    tsm::BoxPrep* prep = (tsm::BoxPrep*)sp;
    unsigned long long* mask = (unsigned long long*)(sp + prep_bytes);
    tsm::NmsGrid* grids = (tsm::NmsGrid*)(sp + prep_bytes + mask_bytes);
    int* cell_start = (int*)(sp + prep_bytes + mask_bytes + grid_bytes);
    int* sorted = (int*)(sp + prep_bytes + mask_bytes + grid_bytes + cs_bytes);
    const int total = frames * nmax;
    tsm::prep_boxes_kernel<<<tsm::divup(total, 128), 128, 0, s>>>(total, boxes, box_stride, prep);
    TSM_LAUNCH_CHECK();
    if (nmax > 65535) return TSM_ERR_INVALID;  // queue entries pack (row, col) into 16 + 16 bits
    const long tiles = (long)cbmax * (cbmax + 1) / 2;
    int ctas_per_sm = 2;
    if (const char* e = tsm_knob(KNOB_NMS_CTAS_PER_SM)) ctas_per_sm = atoi(e) > 0 ? atoi(e) : 2;
    int per_frame = (ctas_per_sm * tsm_num_sms() + frames - 1) / frames;  // CTAs per SM in total
    if (per_frame < 1) per_frame = 1;
    // rotated NMS with a sane threshold: spatial-grid candidates; the all-pairs tile kernel then only serves
    // frames the grid builder flagged (non-finite boxes).  A negative/NaN threshold lets IoU == 0 suppress, so
    // every pair matters and only the tile kernel is exact.
    const bool use_grid = !normal && (thresh >= 0.f);
    const size_t lazy_dyn = tsm::nms_lazy_smem_bytes(nmax);
    const char* algo = tsm_knob(KNOB_NMS_ALGO);  // "mask": always build the full mask (tuning / tests)
    const bool lazy = use_grid && lazy_dyn <= 220 * 1024 && !(algo && !strcmp(algo, "mask"));
    if (!lazy) TSM_CUDA_TRY(cudaMemsetAsync(mask, 0, mask_bytes, s));
    if (use_grid) {
        tsm::nms_grid_build_kernel<<<frames, 1024, 0, s>>>(nmax, counts, prep, grids, cell_start, sorted);
        TSM_LAUNCH_CHECK();
        if (lazy) {
            tsm::nms_clear_flagged_kernel<<<dim3(64, (unsigned)frames), 256, 0, s>>>(mask, (size_t)nmax * cbmax, grids);
            TSM_LAUNCH_CHECK();
            if (lazy_dyn > 40 * 1024)
                TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::nms_lazy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lazy_dyn));
            tsm::nms_lazy_kernel<<<frames, tsm::NMSL_THREADS, lazy_dyn, s>>>(nmax, counts, thresh, prep, grids, cell_start,
                                                                            sorted, keep, num_keep);
        } else {
            dim3 pgrid((unsigned)std::min(per_frame, tsm::divup(nmax, 8)), (unsigned)frames);
            tsm::nms_grid_pairs_kernel<<<pgrid, 256, 0, s>>>(nmax, counts, thresh, prep, grids, cell_start, sorted, mask);
        }
        TSM_LAUNCH_CHECK();
    }
    // the all-pairs tile kernel + mask sweep: axis-aligned NMS, odd thresholds, and -- after the lazy kernel -- only
    // the frames the grid builder flagged (non-finite boxes), for which both exit at once otherwise
    dim3 grid((unsigned)std::min<long>(per_frame, tiles), (unsigned)frames);
    if (normal)
        tsm::nms_mask_kernel<true><<<grid, 256, 0, s>>>(nmax, counts, thresh, prep, mask, nullptr);
    else
        tsm::nms_mask_kernel<false><<<grid, 256, 0, s>>>(nmax, counts, thresh, prep, mask, use_grid ? grids : nullptr);
    TSM_LAUNCH_CHECK();
    const size_t dyn = (size_t)cbmax * sizeof(unsigned long long);
    if (dyn > 200 * 1024) return TSM_ERR_INVALID;
    if (dyn > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    tsm::nms_sweep_kernel<<<frames, 256, dyn, s>>>(nmax, counts, mask, keep, num_keep, lazy ? grids : nullptr);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int pair_matrix_impl(bool iou, int na, const float* a, int nb, const float* b, float* out, cudaStream_t s) {
    if (na <= 0 || nb <= 0) return TSM_OK;
    const size_t pa = align_up((size_t)na * sizeof(tsm::BoxPrep), 256);
    const size_t pb = (size_t)nb * sizeof(tsm::BoxPrep);
    void* scratch = nullptr;
    int rc = tsm_scratch_get(0, pa + pb, s, &scratch);
    if (rc != TSM_OK) return rc;
    tsm::BoxPrep* A = (tsm::BoxPrep*)scratch;
    tsm::BoxPrep* B = (tsm::BoxPrep*)((char*)scratch + pa);
    tsm::prep_boxes_kernel<<<tsm::divup(na, 128), 128, 0, s>>>(na, a, 7, A);
    tsm::prep_boxes_kernel<<<tsm::divup(nb, 128), 128, 0, s>>>(nb, b, 7, B);
    TSM_LAUNCH_CHECK();
    dim3 grid((unsigned)tsm::divup(nb, 64), (unsigned)tsm::divup(na, 64));
    if (grid.y > 65535) return TSM_ERR_INVALID;
    if (iou)
        tsm::pair_matrix_kernel<true><<<grid, 256, 0, s>>>(na, A, nb, B, out);
    else
        tsm::pair_matrix_kernel<false><<<grid, 256, 0, s>>>(na, A, nb, B, out);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

}  // namespace

extern "C" {

int tsmdet_boxes_overlap_bev(int num_a, const float* boxes_a, int num_b, const float* boxes_b, float* ans_overlap,
                             void* stream) {
    return pair_matrix_impl(false, num_a, boxes_a, num_b, boxes_b, ans_overlap, (cudaStream_t)stream);
}

int tsmdet_boxes_iou_bev(int num_a, const float* boxes_a, int num_b, const float* boxes_b, float* ans_iou,
                         void* stream) {
    return pair_matrix_impl(true, num_a, boxes_a, num_b, boxes_b, ans_iou, (cudaStream_t)stream);
}

// Device-resident batched NMS: boxes (frames, nmax, box_stride>=7) already sorted by
// descending score per frame; counts (frames) valid boxes per frame or NULL (= nmax).
int tsmdet_nms_batch(int frames, int nmax, const float* boxes, int box_stride, const int* counts, float thresh,
                     long long* keep, int* num_keep, void* stream) {
    return nms_batch_impl(false, frames, nmax, boxes, box_stride, counts, thresh, keep, num_keep,
                          (cudaStream_t)stream);
}

int tsmdet_nms_normal_batch(int frames, int nmax, const float* boxes, int box_stride, const int* counts, float thresh,
                            long long* keep, int* num_keep, void* stream) {
    return nms_batch_impl(true, frames, nmax, boxes, box_stride, counts, thresh, keep, num_keep,
                          (cudaStream_t)stream);
}

// Reference-shaped entry (iou3d_nms.cpp:90-136): boxes (n,7) on the device in score
// order, keep_host (n) int64 on the HOST; returns through *num_out.  Synchronous like the
// reference, but only the keep list (8n bytes) crosses PCIe, not the n*ceil(n/64) mask.
static int nms_single(bool normal, int n, const float* boxes, float thresh, long long* keep_host, int* num_out,
                      cudaStream_t s) {
    *num_out = 0;
    if (n <= 0) return TSM_OK;
    long long* keep_dev = nullptr;
    TSM_CUDA_TRY(cudaMallocAsync(&keep_dev, (size_t)n * sizeof(long long) + 16, s));
    int* num_dev = (int*)(keep_dev + n);
    int rc = nms_batch_impl(normal, 1, n, boxes, 7, nullptr, thresh, keep_dev, num_dev, s);
    if (rc == TSM_OK) {
        cudaError_t e = cudaMemcpyAsync(num_out, num_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e == cudaSuccess && *num_out > 0)
            e = cudaMemcpyAsync(keep_host, keep_dev, (size_t)(*num_out) * sizeof(long long), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        rc = (int)e;
    }
    cudaFreeAsync(keep_dev, s);
    return rc;
}

// Host-side rotated BEV IoU (the reference keeps a CPU variant for its data pipeline):
// boxes_a (N,7), boxes_b (M,7), ans_iou (N,M) all in HOST memory.  Same per-pair arithmetic as
// the device path, evaluated with the host libm.
// ref: iou3d_nms_api.cpp:16 boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252)
int tsmdet_boxes_iou_bev_cpu(int num_a, const float* boxes_a, int num_b, const float* boxes_b, float* ans_iou) {
    if (num_a <= 0 || num_b <= 0) return TSM_OK;
    std::vector<tsm::BoxPrep> A((size_t)num_a), B((size_t)num_b);
    for (int i = 0; i < num_a; ++i) tsm::prep_box(boxes_a + (size_t)i * 7, A[i]);
    for (int j = 0; j < num_b; ++j) tsm::prep_box(boxes_b + (size_t)j * 7, B[j]);
    for (int i = 0; i < num_a; ++i)
        for (int j = 0; j < num_b; ++j)
            ans_iou[(size_t)i * num_b + j] = (tsm::surely_disjoint(A[i], B[j]) || tsm::surely_separated(A[i], B[j]))
                                                 ? 0.f
                                                 : tsm::iou_rotated(A[i], B[j]);
    return TSM_OK;
}

int tsmdet_nms_gpu(int n, const float* boxes, float thresh, long long* keep_host, int* num_out, void* stream) {
    return nms_single(false, n, boxes, thresh, keep_host, num_out, (cudaStream_t)stream);
}

int tsmdet_nms_normal_gpu(int n, const float* boxes, float thresh, long long* keep_host, int* num_out, void* stream) {
    return nms_single(true, n, boxes, thresh, keep_host, num_out, (cudaStream_t)stream);
}

}  // extern "C"
