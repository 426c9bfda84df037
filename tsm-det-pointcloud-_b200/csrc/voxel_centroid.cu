// voxel_centroid.cu -- centroid voxelisation after SA layer 0 and the dense voxel -> index table.
//
// Replaces, for the tail of _VoxelPointnetSAModuleFS(Distillation)Base.forward's layer-0 branch
// (/root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py:1323-1355):
//   get_voxel_indices        pcdet/utils/voxel_aggregation_utils.py:48-83   ((xyz - range_min) / voxel_size).long()
//   flip + cat batch index   pointnet2_modules.py:1330-1337                  -> (N,4) int64 [b,z,y,x]
//   permute / view / cat     :1346-1348                                      -> (N, 4+C) rows [b,x,y,z,features]
//   get_centroid_per_voxel   voxel_aggregation_utils.py:132-161              voxel_idxs.unique(dim=0, return_inverse,
//                                                                            return_counts) + zeros.scatter_add_ + divide
// and generate_voxel2pinds (pcdet/utils/common_utils.py:248-265: a dense (B,Z,Y,X) int32 table, -1 = empty).
//
// The reference runs torch.unique(dim=0) -- a device-wide lexicographic sort of (N,4) int64 rows -- two scatter_add_
// with atomics (so its sums depend on the atomics' order) and ~10 small copies.  Here:
//   kernel 1 (one CTA per frame): voxel coordinates with the reference's own fp32 sub / div / truncation, a 64-bit key
//            (z | y | x | point index) per point, a bitonic sort of the frame's keys IN SHARED MEMORY (frames are
//            independent: the batch index is the major sort key of the reference's unique), segment heads by scan;
//   kernel 2 (frames x tiles of sorted points): global voxel numbers from the per-frame counts, inverse indices,
//            counts, and the centroids -- every voxel's points are added IN ASCENDING POINT ORDER, i.e. exactly the
//            sequence of a sequential scatter_add_ (torch's CPU kernel), so results are deterministic and bit-equal
//            to the reference functions run on the CPU, which is how the oracle is pinned.
// HBM-bound by construction: 12 + 4C bytes read and <= 16 + 4C + 80 bytes written per point.
#include "common.cuh"

namespace tsm {

constexpr int VC_THREADS = 1024;
constexpr int VC_MAX_POINTS = 16384;   // per frame: 8 B * 16384 keys = 128 KB of shared memory
constexpr int VC_ERR_RANGE = 1;        // a voxel coordinate outside [-32768, 32767]
constexpr int VC_ERR_BATCH = 2;        // generic path: rows are not grouped frame after frame

struct VcArgs {
    int b, m, c;                 // frames, points per frame, feature channels (SoA) | row width - 4 (row-major)
    const float* xyz;            // SoA: (B,M,3)
    const float* features;       // SoA: (B,C,M) or null
    const float* rows;           // row-major: (B*M, 4+C) [b,x,y,z,f...] or null
    const long long* vidx_in;    // row-major path: (B*M,4) i64 voxel indices given by the caller, or null
    const long long* weights;    // optional num_points_in_voxel (B*M) i64, or null
    float vs[3], r0[3];          // voxel size, range minimum (x,y,z)
    long long* voxel_idxs;       // out (B*M,4) i64 [b,z,y,x] (SoA path) -- or == vidx_in
    float* centroids;            // out (<=B*M, 4+C)
    long long* cvi;              // out (<=B*M, 4) i64
    long long* counts;           // out (<=B*M) i64
    long long* inverse;          // out (B*M) i64
    int* num_unique;             // out (1) i32
    int* order;                  // scratch (B*M) sorted point index within its frame
    int* seg;                    // scratch (B*M) voxel number within its frame, per sorted position
    int* ucount;                 // scratch (B) unique voxels per frame
    int* err;                    // device error bits (or-ed)
};

__device__ __forceinline__ unsigned long long vc_key(long long z, long long y, long long x, int idx, int* err) {
    if (z < -32768 || z > 32767 || y < -32768 || y > 32767 || x < -32768 || x > 32767) {
        atomicOr(err, VC_ERR_RANGE);
        z = max(-32768LL, min(32767LL, z));
        y = max(-32768LL, min(32767LL, y));
        x = max(-32768LL, min(32767LL, x));
    }
    return ((unsigned long long)(z + 32768) << 48) | ((unsigned long long)(y + 32768) << 32) |
           ((unsigned long long)(x + 32768) << 16) | (unsigned long long)(unsigned)idx;
}

// One CTA per frame: keys -> shared-memory bitonic sort -> segment numbers.
__global__ void __launch_bounds__(VC_THREADS) voxel_sort_kernel(const VcArgs a, const int n2) {
    extern __shared__ unsigned long long keys[];  // n2 (power of two >= m)
    __shared__ int warp_tot[VC_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, m = a.m;
    for (int i = tid; i < n2; i += VC_THREADS) {
        unsigned long long k = ~0ull;
        if (i < m) {
            const long long row = (long long)b * m + i;
            long long vx, vy, vz;
            if (a.vidx_in) {
                if (a.vidx_in[row * 4 + 0] != (long long)b) atomicOr(a.err, VC_ERR_BATCH);
                vz = a.vidx_in[row * 4 + 1];
                vy = a.vidx_in[row * 4 + 2];
                vx = a.vidx_in[row * 4 + 3];
            } else {
                // get_voxel_indices: fp32 subtract, fp32 IEEE divide, truncation toward zero (.long())
                const float* p = a.xyz + row * 3;
                vx = (long long)__fdiv_rn(__fsub_rn(__ldg(p + 0), a.r0[0]), a.vs[0]);
                vy = (long long)__fdiv_rn(__fsub_rn(__ldg(p + 1), a.r0[1]), a.vs[1]);
                vz = (long long)__fdiv_rn(__fsub_rn(__ldg(p + 2), a.r0[2]), a.vs[2]);
                long long* o = a.voxel_idxs + row * 4;
                o[0] = b;
                o[1] = vz;
                o[2] = vy;
                o[3] = vx;
            }
            k = vc_key(vz, vy, vx, i, a.err);
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n2 >> 1); t += VC_THREADS) {
                // the t-th compare-exchange of this stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const unsigned long long x = keys[i], y = keys[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) {
                    keys[i] = y;
                    keys[p] = x;
                }
            }
            __syncthreads();
        }
    }
    // segment heads -> voxel number within the frame (exclusive scan of heads), written per sorted position
    const int per = (n2 + VC_THREADS - 1) / VC_THREADS;
    const int j0 = tid * per;
    int local = 0;
    for (int q = 0; q < per; ++q) {
        const int j = j0 + q;
        if (j < m) local += (j == 0) || ((keys[j] >> 16) != (keys[j - 1] >> 16));
    }
    int incl = local;
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if ((tid & 31) >= d) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int v = tid < VC_THREADS / 32 ? warp_tot[tid] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(FULL, v, d);
            if (tid >= d) v += u;
        }
        if (tid < VC_THREADS / 32) warp_tot[tid] = v;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - local + ((tid >> 5) > 0 ? warp_tot[(tid >> 5) - 1] : 0);  // heads before this thread's range
    for (int q = 0; q < per; ++q) {
        const int j = j0 + q;
        if (j < m) {
            run += (j == 0) || ((keys[j] >> 16) != (keys[j - 1] >> 16));
            a.seg[(long long)b * m + j] = run - 1;
            a.order[(long long)b * m + j] = (int)(keys[j] & 0xffffull);
        }
    }
    if (tid == VC_THREADS - 1) a.ucount[b] = warp_tot[VC_THREADS / 32 - 1];
}

// grid (tiles of VC_TILE sorted positions, frames)
constexpr int VC_TILE = 64;
template <bool ROWMAJOR>
__global__ void __launch_bounds__(256) voxel_centroid_kernel(const VcArgs a) {
    __shared__ int s_base;
    __shared__ int s_cnt[VC_TILE];    // > 0 for segment heads: points in the voxel
    __shared__ int s_u[VC_TILE];      // global voxel number
    const int b = blockIdx.y, tid = threadIdx.x, m = a.m;
    const int j0 = blockIdx.x * VC_TILE;
    if (tid == 0) {
        int base = 0;
        for (int q = 0; q < b; ++q) base += a.ucount[q];
        s_base = base;
        if (b == a.b - 1 && blockIdx.x == 0) *a.num_unique = base + a.ucount[b];
    }
    __syncthreads();
    const int* seg = a.seg + (long long)b * m;
    const int* order = a.order + (long long)b * m;
    const long long* vin = a.vidx_in ? a.vidx_in : a.voxel_idxs;
    if (tid < VC_TILE) {
        const int j = j0 + tid;
        int cnt = 0, u = 0;
        if (j < m) {
            const int s = seg[j];
            u = s_base + s;
            const int p = order[j];
            a.inverse[(long long)b * m + p] = u;
            if (j == 0 || seg[j - 1] != s) {
                int e = j + 1;
                while (e < m && seg[e] == s) ++e;
                cnt = e - j;
                a.counts[u] = cnt;
                const long long* v = vin + ((long long)b * m + p) * 4;
                long long* o = a.cvi + (long long)u * 4;
                o[0] = v[0];
                o[1] = v[1];
                o[2] = v[2];
                o[3] = v[3];
            }
        }
        s_cnt[tid] = cnt;
        s_u[tid] = u;
    }
    __syncthreads();
    // centroids: one (voxel, column) per thread step, columns fastest -> coalesced row-major stores; the points of a
    // voxel are added in ascending point index (the sort key's low bits), as a sequential scatter_add_ does
    const int w = 4 + a.c;
    for (int e = tid; e < VC_TILE * w; e += 256) {
        const int jj = e / w, col = e - jj * w;
        const int cnt = s_cnt[jj];
        if (cnt == 0) continue;
        const int j = j0 + jj;
        float acc = 0.f;
        long long wsum = 0;
        for (int k = 0; k < cnt; ++k) {
            const int p = order[j + k];
            float v;
            if (ROWMAJOR) {
                v = __ldg(a.rows + ((long long)b * m + p) * w + col);
            } else {
                v = col == 0 ? (float)b
                             : (col < 4 ? __ldg(a.xyz + ((long long)b * m + p) * 3 + (col - 1))
                                        : __ldg(a.features + ((long long)b * a.c + (col - 4)) * m + p));
            }
            if (a.weights) {
                const long long wt = a.weights[(long long)b * m + p];
                v = __fmul_rn(v, (float)wt);  // points * num_points_in_voxel.unsqueeze(-1) (:148)
                wsum += wt;
            }
            acc = __fadd_rn(acc, v);
        }
        const float den = a.weights ? (float)wsum : (float)cnt;
        a.centroids[(long long)s_u[jj] * w + col] = __fdiv_rn(acc, den);
    }
}

// dense (B,Z,Y,X) table: out[b,z,y,x] = row number of the voxel in `indices` ((n,4) i32 [b,z,y,x]); `value` < 0 clears
__global__ void __launch_bounds__(256) voxel2pinds_scatter_kernel(const int n, const int* __restrict__ indices, const int nb,
                                                                  const int nz, const int ny, const int nx,
                                                                  int* __restrict__ out, const int clear, int* err) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int4 v = __ldg(reinterpret_cast<const int4*>(indices) + i);
    if (v.x < 0 || v.x >= nb || v.y < 0 || v.y >= nz || v.z < 0 || v.z >= ny || v.w < 0 || v.w >= nx) {
        if (err) atomicOr(err, VC_ERR_RANGE);
        return;
    }
    out[(((long long)v.x * nz + v.y) * ny + v.z) * nx + v.w] = clear ? -1 : i;
}

}  // namespace tsm

static int vc_launch(tsm::VcArgs& a, bool rowmajor, cudaStream_t s) {
    using namespace tsm;
    if (a.b <= 0 || a.m <= 0) {
        if (a.num_unique) TSM_CUDA_TRY(cudaMemsetAsync(a.num_unique, 0, sizeof(int), s));
        return TSM_OK;
    }
    if (a.m > VC_MAX_POINTS || a.c < 0) return TSM_ERR_INVALID;
    void* p = nullptr;
    const size_t n = (size_t)a.b * a.m;
    int rc = tsm_scratch_get(8, (2 * n + a.b + 1) * sizeof(int), s, &p);
    if (rc != TSM_OK) return rc;
    a.order = (int*)p;
    a.seg = a.order + n;
    a.ucount = a.seg + n;
    a.err = a.ucount + a.b;
    TSM_CUDA_TRY(cudaMemsetAsync(a.err, 0, sizeof(int), s));
    int n2 = 32;
    while (n2 < a.m) n2 <<= 1;
    const size_t smem = (size_t)n2 * sizeof(unsigned long long);
    if (smem > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(voxel_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    voxel_sort_kernel<<<a.b, VC_THREADS, smem, s>>>(a, n2);
    TSM_LAUNCH_CHECK();
    dim3 grid((unsigned)divup(a.m, VC_TILE), (unsigned)a.b);
    if (rowmajor)
        voxel_centroid_kernel<true><<<grid, 256, 0, s>>>(a);
    else
        voxel_centroid_kernel<false><<<grid, 256, 0, s>>>(a);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

// The layer-0 tail in one call.  new_xyz (B,M,3), features (B,C,M) | NULL (C = 0); voxel_size / range_min as x,y,z.
// Outputs have capacity B*M rows; the first *num_unique rows of centroids / centroid_voxel_idxs / labels_count are
// valid.  err (1) i32 device word, or-ed: 1 = a voxel coordinate outside [-32768, 32767].  M <= 16384.
extern "C" int tsmdet_voxel_centroids(int b, int m, int c, const float* new_xyz, const float* features, float vx, float vy,
                                      float vz, float x0, float y0, float z0, long long* voxel_idxs, float* centroids,
                                      long long* centroid_voxel_idxs, long long* labels_count, long long* unique_idxs,
                                      int* num_unique, int* err, void* stream) {
    if (!new_xyz || (c > 0 && !features) || !voxel_idxs || !centroids || !centroid_voxel_idxs || !labels_count ||
        !unique_idxs || !num_unique)
        return TSM_ERR_INVALID;
    tsm::VcArgs a = {};
    a.b = b;
    a.m = m;
    a.c = c;
    a.xyz = new_xyz;
    a.features = features;
    a.vs[0] = vx, a.vs[1] = vy, a.vs[2] = vz;
    a.r0[0] = x0, a.r0[1] = y0, a.r0[2] = z0;
    a.voxel_idxs = voxel_idxs;
    a.centroids = centroids;
    a.cvi = centroid_voxel_idxs;
    a.counts = labels_count;
    a.inverse = unique_idxs;
    a.num_unique = num_unique;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = vc_launch(a, false, s);
    if (rc == TSM_OK && err && b > 0 && m > 0) TSM_CUDA_TRY(cudaMemcpyAsync(err, a.err, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return rc;
}

// get_centroid_per_voxel with the reference's own argument layout: points (B*M, 4+F) rows [b,x,y,z,f...],
// voxel_idxs (B*M,4) i64, rows grouped frame after frame with M rows each (what its call sites pass);
// num_points_in_voxel (B*M) i64 | NULL.  err bit 2 = rows not grouped by frame.
extern "C" int tsmdet_centroid_per_voxel(int b, int m, int f, const float* points, const long long* voxel_idxs,
                                         const long long* num_points_in_voxel, float* centroids,
                                         long long* centroid_voxel_idxs, long long* labels_count, long long* unique_idxs,
                                         int* num_unique, int* err, void* stream) {
    if (!points || !voxel_idxs || !centroids || !centroid_voxel_idxs || !labels_count || !unique_idxs || !num_unique)
        return TSM_ERR_INVALID;
    tsm::VcArgs a = {};
    a.b = b;
    a.m = m;
    a.c = f;
    a.rows = points;
    a.vidx_in = voxel_idxs;
    a.weights = num_points_in_voxel;
    a.centroids = centroids;
    a.cvi = centroid_voxel_idxs;
    a.counts = labels_count;
    a.inverse = unique_idxs;
    a.num_unique = num_unique;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = vc_launch(a, true, s);
    if (rc == TSM_OK && err && b > 0 && m > 0) TSM_CUDA_TRY(cudaMemcpyAsync(err, a.err, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return rc;
}

// generate_voxel2pinds: out (B,Z,Y,X) i32.  prev_indices != NULL: `out` still holds the table of `prev_indices`
// (n_prev rows) from the previous call and only those entries are reset -- O(voxels) instead of re-filling the
// whole volume (5.9 GB for 16 KITTI frames at full resolution); otherwise the table is filled with -1 first.
extern "C" int tsmdet_voxel2pinds(int n, const int* indices, int n_prev, const int* prev_indices, int nb, int nz, int ny,
                                  int nx, int* out, int* err, void* stream) {
    if (!out || nb <= 0 || nz <= 0 || ny <= 0 || nx <= 0 || (n > 0 && !indices)) return TSM_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (prev_indices) {
        if (n_prev > 0) {
            tsm::voxel2pinds_scatter_kernel<<<tsm::divup(n_prev, 256), 256, 0, s>>>(n_prev, prev_indices, nb, nz, ny, nx, out, 1, nullptr);
            TSM_LAUNCH_CHECK();
        }
    } else {
        TSM_CUDA_TRY(cudaMemsetAsync(out, 0xff, sizeof(int) * (size_t)nb * nz * ny * nx, s));  // int32 -1
    }
    if (n > 0) {
        tsm::voxel2pinds_scatter_kernel<<<tsm::divup(n, 256), 256, 0, s>>>(n, indices, nb, nz, ny, nx, out, 0, err);
        TSM_LAUNCH_CHECK();
    }
    return TSM_OK;
}
