// group.cu -- index gathers of the set-abstraction stack for sm_100a.
//
// Replaces (bit-exact copies / the same single f32 subtract):
//   gather_points_kernel_fast(+grad)  /root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:15-31, 53-70
//   group_points_kernel_fast(+grad)   group_points_gpu.cu:53-72, 14-31
//   QueryAndGroup(.Dilated).forward's  transpose -> group -> subtract -> group -> cat chain
//       pointnet2_utils.py:515-530, 553-568   (tsmdet_group_concat: one kernel, one pass)
//
// These are HBM-bound copies.  Layout choices: one thread owns 4 consecutive (point,sample)
// slots, reads its 4 indices once (int4) and loops over channels, so every store is a
// coalesced 16-byte vector and the index tensor is read exactly once instead of once per
// channel; the gathers hit L2 (a (C,N) feature slab is <= a few MB).
#include <stdlib.h>

#include "common.cuh"

namespace tsm {

constexpr int GP_THREADS = 256;

// points (B,C,N), idx (B,E) [E = npoints*nsample] -> out (B,C,E)
template <int VEC>
__global__ void __launch_bounds__(GP_THREADS)
    group_points_kernel(int c, int n, int e, const float* __restrict__ points, const int* __restrict__ idx,
                        float* __restrict__ out) {
    const int b = blockIdx.y;
    const int v = blockIdx.x * GP_THREADS + threadIdx.x;  // vector slot
    const int e0 = v * VEC;
    if (e0 >= e) return;
    points += (size_t)b * c * n;
    out += (size_t)b * c * e;
    int id[VEC];
    if constexpr (VEC == 4) {
        const int4 t = *reinterpret_cast<const int4*>(idx + (size_t)b * e + e0);
        id[0] = t.x; id[1] = t.y; id[2] = t.z; id[VEC - 1] = t.w;
    } else {
        id[0] = idx[(size_t)b * e + e0];
    }
#pragma unroll 4
    for (int ci = 0; ci < c; ++ci) {
        const float* src = points + (size_t)ci * n;
        if constexpr (VEC == 4) {
            float4 o;
            o.x = __ldg(src + id[0]); o.y = __ldg(src + id[1]); o.z = __ldg(src + id[2]); o.w = __ldg(src + id[VEC - 1]);
            __stcs(reinterpret_cast<float4*>(out + (size_t)ci * e + e0), o);
        } else {
            out[(size_t)ci * e + e0] = __ldg(src + id[0]);
        }
    }
}

// grad_out (B,C,E), idx (B,E) -> grad_points (B,C,N) += (atomic, like the reference)
__global__ void __launch_bounds__(GP_THREADS)
    group_points_grad_kernel(int c, int n, int e, const float* __restrict__ grad_out, const int* __restrict__ idx,
                             float* __restrict__ grad_points) {
    const int b = blockIdx.y;
    const int ei = blockIdx.x * GP_THREADS + threadIdx.x;
    if (ei >= e) return;
    const int id = idx[(size_t)b * e + ei];
    grad_out += (size_t)b * c * e + ei;
    grad_points += (size_t)b * c * n + id;
    for (int ci = 0; ci < c; ++ci) atomicAdd(grad_points + (size_t)ci * n, grad_out[(size_t)ci * e]);
}

// Fused QueryAndGroup materialisation.
//   xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or null, idx (B,M,S)
//   -> new_features (B, 3*use_xyz + C, M, S)   [xyz offsets first, pointnet2_utils.py:523]
//   -> grouped_xyz  (B, 3, M, S)               [optional]
template <int VEC>
__global__ void __launch_bounds__(GP_THREADS)
    group_concat_kernel(int c, int n, int m, int s, int use_xyz, const float* __restrict__ xyz,
                        const float* __restrict__ new_xyz, const float* __restrict__ features,
                        const int* __restrict__ idx, float* __restrict__ new_features,
                        float* __restrict__ grouped_xyz, int ctot) {
    const int b = blockIdx.y;
    const int e = m * s;
    const int v = blockIdx.x * GP_THREADS + threadIdx.x;
    const int e0 = v * VEC;
    if (e0 >= e) return;
    xyz += (size_t)b * n * 3;
    int id[VEC];
    if constexpr (VEC == 4) {
        const int4 t = *reinterpret_cast<const int4*>(idx + (size_t)b * e + e0);
        id[0] = t.x; id[1] = t.y; id[2] = t.z; id[VEC - 1] = t.w;
    } else {
        id[0] = idx[(size_t)b * e + e0];
    }
    // ctot: channels of new_features per cloud (the feature part may be filled by the staged gather instead)
    float* nf = new_features ? new_features + (size_t)b * ctot * e : nullptr;
    float* gx = grouped_xyz ? grouped_xyz + (size_t)b * 3 * e : nullptr;
    // the VEC slots of one thread share a centre when VEC divides s; handle the general case per slot
    float off[3][VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int p = (e0 + j) / s;
        const float* q = new_xyz + ((size_t)b * m + p) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) off[a][j] = __fsub_rn(__ldg(xyz + (size_t)id[j] * 3 + a), __ldg(q + a));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if constexpr (VEC == 4) {
            const float4 o = make_float4(off[a][0], off[a][1], off[a][2], off[a][VEC - 1]);
            if (gx) __stcs(reinterpret_cast<float4*>(gx + (size_t)a * e + e0), o);
            if (nf && use_xyz) __stcs(reinterpret_cast<float4*>(nf + (size_t)a * e + e0), o);
        } else {
            if (gx) gx[(size_t)a * e + e0] = off[a][0];
            if (nf && use_xyz) nf[(size_t)a * e + e0] = off[a][0];
        }
    }
    if (features && nf) {
        features += (size_t)b * c * n;
        float* dst = nf + (size_t)(use_xyz ? 3 : 0) * e;
#pragma unroll 4
        for (int ci = 0; ci < c; ++ci) {
            const float* src = features + (size_t)ci * n;
            if constexpr (VEC == 4) {
                float4 o;
                o.x = __ldg(src + id[0]); o.y = __ldg(src + id[1]); o.z = __ldg(src + id[2]); o.w = __ldg(src + id[VEC - 1]);
                __stcs(reinterpret_cast<float4*>(dst + (size_t)ci * e + e0), o);
            } else {
                dst[(size_t)ci * e + e0] = __ldg(src + id[0]);
            }
        }
    }
}

// points (B,C,N), idx (B,M) -> out (B,C,M)
__global__ void __launch_bounds__(GP_THREADS)
    gather_points_kernel(int c, int n, int m, const float* __restrict__ points, const int* __restrict__ idx,
                         float* __restrict__ out) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * GP_THREADS + threadIdx.x;
    if (p >= m) return;
    const int id = idx[(size_t)b * m + p];
    points += (size_t)b * c * n + id;
    out += (size_t)b * c * m + p;
    for (int ci = 0; ci < c; ++ci) out[(size_t)ci * m] = __ldg(points + (size_t)ci * n);
}

__global__ void __launch_bounds__(GP_THREADS)
    gather_points_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out, const int* __restrict__ idx,
                              float* __restrict__ grad_points) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * GP_THREADS + threadIdx.x;
    if (p >= m) return;
    const int id = idx[(size_t)b * m + p];
    grad_out += (size_t)b * c * m + p;
    grad_points += (size_t)b * c * n + id;
    for (int ci = 0; ci < c; ++ci) atomicAdd(grad_points + (size_t)ci * n, grad_out[(size_t)ci * m]);
}

// Gather rows of an (B,N,3) array: new_xyz[b,p,:] = xyz[b,idx[b,p],:]  (replaces the
// transpose -> gather_operation -> transpose chain of pointnet2_modules.py:1143, 1212-1215)
__global__ void __launch_bounds__(GP_THREADS)
    gather_xyz_kernel(int n, int m, const float* __restrict__ xyz, const int* __restrict__ idx,
                      float* __restrict__ out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * GP_THREADS + threadIdx.x;
    if (t >= m * 3) return;
    const int p = t / 3, a = t - p * 3;
    out[(size_t)b * m * 3 + t] = __ldg(xyz + ((size_t)b * n + idx[(size_t)b * m + p]) * 3 + a);
}


// Shared-memory staged variant: a CTA copies a slab of CC feature rows, points[b, c0 : c0+CC, :] (one contiguous
// CC*N*4-byte run of the (B,C,N) tensor), into shared memory with ONE bulk copy of the TMA engine, then serves every
// gather of its E-range from shared memory: HBM sees the slab once and the output once, the L2 no scattered
// 4-byte requests at all.  grid = (E chunks, channel chunks, B); out has `out_ctot` channels per cloud and this
// call fills channels [out_c0, out_c0 + C).
__global__ void __launch_bounds__(GP_THREADS)
    group_points_staged_kernel(int c, int n, int e, int cc, int e_per_cta, const float* __restrict__ points,
                               const int* __restrict__ idx, float* __restrict__ out, int out_ctot, int out_c0,
                               int* status) {
    extern __shared__ __align__(128) float slab[];  // [cc][n]
    __shared__ __align__(8) uint64_t full;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * cc;
    const int nc = min(cc, c - c0);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(smem_u32(&full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)nc * (uint32_t)n * 4u;
        mbar_arrive_expect_tx(smem_u32(&full), bytes);
        bulk_g2s(smem_u32(slab), points + ((size_t)b * c + c0) * n, bytes, smem_u32(&full));
    }
    __syncthreads();
    {
        const uint32_t bar = smem_u32(&full);
        if (!mbar_try_wait_cta(bar, 0)) {
            const long long t0 = clock64();
            while (!mbar_try_wait_cta(bar, 0))
                if (clock64() - t0 > 4000000000LL) watchdog_trip(status, TSM_ERR_WATCHDOG);
        }
    }
    const int e_beg = blockIdx.x * e_per_cta;
    const int e_end = min(e, e_beg + e_per_cta);
    const int* __restrict__ ib = idx + (size_t)b * e;
    float* __restrict__ ob = out + ((size_t)b * out_ctot + out_c0 + c0) * e;
    for (int e0 = e_beg + tid * 4; e0 < e_end; e0 += GP_THREADS * 4) {
        const int4 t = *reinterpret_cast<const int4*>(ib + e0);
#pragma unroll 4
        for (int ci = 0; ci < nc; ++ci) {
            const float* row = slab + (size_t)ci * n;
            __stcs(reinterpret_cast<float4*>(ob + (size_t)ci * e + e0), make_float4(row[t.x], row[t.y], row[t.z], row[t.w]));
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Launches the staged kernel when the shapes allow it (returns false otherwise, nothing launched).
static bool launch_group_staged(int b, int c, int n, long e, const float* points, const int* idx, float* out,
                                int out_ctot, int out_c0, cudaStream_t s, int* rc) {
    *rc = TSM_OK;
    if ((n & 3) || (e & 3) || !aligned16(points) || !aligned16(idx) || !aligned16(out) || e < 4096 || b > 65535) return false;
    if ((size_t)n * 4 > 96 * 1024) return false;  // a row must leave room for two CTAs per SM
    size_t slab_cap = 32 * 1024;  // measured: 16..32 KB slabs (5-6 CTAs per SM) beat larger ones
    if (const char* e = tsm_knob(KNOB_GROUP_SLAB_KB)) slab_cap = (size_t)atoi(e) * 1024;
    int cc = 1;
    while (cc * 2 <= c && (size_t)cc * 2 * n * 4 <= slab_cap) cc *= 2;
    const int cchunks = divup(c, cc);
    if (cchunks > 65535) return false;
    const size_t dyn = (size_t)cc * n * 4;
    const int resident = (int)((227 * 1024) / (dyn + 1024));
    int waves = 4;
    if (const char* e = tsm_knob(KNOB_GROUP_WAVES)) waves = atoi(e) > 0 ? atoi(e) : 4;
    long want = (long)tsm_num_sms() * (resident > 8 ? 8 : resident) * waves;  // about two waves
    long echunks = divup((int)want, b * cchunks);
    const long emax = e / 2048 > 0 ? e / 2048 : 1;
    if (echunks > emax) echunks = emax;
    if (echunks < 1) echunks = 1;
    long e_per = ((e + echunks - 1) / echunks + 1023) / 1024 * 1024;  // whole passes of the CTA
    echunks = (e + e_per - 1) / e_per;
    if (dyn > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(group_points_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (err != cudaSuccess) { *rc = (int)err; return true; }
    }
    dim3 grid((unsigned)echunks, (unsigned)cchunks, (unsigned)b);
    group_points_staged_kernel<<<grid, GP_THREADS, dyn, s>>>(c, n, (int)e, cc, (int)e_per, points, idx, out, out_ctot,
                                                             out_c0, tsm_status_word(s));
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) *rc = (int)err;
    return true;
}

}  // namespace tsm

extern "C" {

int tsmdet_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out,
                        void* stream) {
    const long e = (long)npoints * nsample;
    if (b <= 0 || c <= 0 || e <= 0) return TSM_OK;
    if (b > 65535 || e > 0x7fffffffL) return TSM_ERR_INVALID;
    const bool vec = (e % 4 == 0) && tsm::aligned16(idx) && tsm::aligned16(out);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = TSM_OK;
    if (!tsm_knob(KNOB_GROUP_DIRECT) && tsm::launch_group_staged(b, c, n, e, points, idx, out, c, 0, s, &rc)) return rc;
    if (vec) {
        dim3 grid((unsigned)tsm::divup((int)(e / 4), tsm::GP_THREADS), (unsigned)b);
        tsm::group_points_kernel<4><<<grid, tsm::GP_THREADS, 0, s>>>(c, n, (int)e, points, idx, out);
    } else {
        dim3 grid((unsigned)tsm::divup((int)e, tsm::GP_THREADS), (unsigned)b);
        tsm::group_points_kernel<1><<<grid, tsm::GP_THREADS, 0, s>>>(c, n, (int)e, points, idx, out);
    }
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx,
                             float* grad_points, void* stream) {
    const long e = (long)npoints * nsample;
    if (b <= 0 || c <= 0 || e <= 0) return TSM_OK;
    if (b > 65535 || e > 0x7fffffffL) return TSM_ERR_INVALID;
    dim3 grid((unsigned)tsm::divup((int)e, tsm::GP_THREADS), (unsigned)b);
    tsm::group_points_grad_kernel<<<grid, tsm::GP_THREADS, 0, (cudaStream_t)stream>>>(c, n, (int)e, grad_out, idx,
                                                                                      grad_points);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_group_concat(int b, int c, int n, int m, int nsample, int use_xyz, const float* xyz, const float* new_xyz,
                        const float* features, const int* idx, float* new_features, float* grouped_xyz, void* stream) {
    const long e = (long)m * nsample;
    if (b <= 0 || e <= 0) return TSM_OK;
    if (b > 65535 || e > 0x7fffffffL) return TSM_ERR_INVALID;
    const bool vec = (e % 4 == 0) && tsm::aligned16(idx) && (!new_features || tsm::aligned16(new_features)) &&
                     (!grouped_xyz || tsm::aligned16(grouped_xyz));
    cudaStream_t s = (cudaStream_t)stream;
    const int ctot = (use_xyz ? 3 : 0) + (features ? c : 0);
    if (features && new_features && c >= 8 && !tsm_knob(KNOB_GROUP_DIRECT)) {
        // many feature channels: the staged gather fills channels [3*use_xyz, ...) and the fused kernel below only
        // writes the coordinate offsets
        int rc = TSM_OK;
        if (tsm::launch_group_staged(b, c, n, e, features, idx, new_features, ctot, use_xyz ? 3 : 0, s, &rc)) {
            if (rc != TSM_OK) return rc;
            features = nullptr;
            if (!use_xyz && !grouped_xyz) return TSM_OK;
        }
    }
    if (vec) {
        dim3 grid((unsigned)tsm::divup((int)(e / 4), tsm::GP_THREADS), (unsigned)b);
        tsm::group_concat_kernel<4><<<grid, tsm::GP_THREADS, 0, s>>>(c, n, m, nsample, use_xyz, xyz, new_xyz, features,
                                                                     idx, new_features, grouped_xyz, ctot);
    } else {
        dim3 grid((unsigned)tsm::divup((int)e, tsm::GP_THREADS), (unsigned)b);
        tsm::group_concat_kernel<1><<<grid, tsm::GP_THREADS, 0, s>>>(c, n, m, nsample, use_xyz, xyz, new_xyz, features,
                                                                     idx, new_features, grouped_xyz, ctot);
    }
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out,
                         void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0) return TSM_OK;
    if (b > 65535) return TSM_ERR_INVALID;
    dim3 grid((unsigned)tsm::divup(npoints, tsm::GP_THREADS), (unsigned)b);
    tsm::gather_points_kernel<<<grid, tsm::GP_THREADS, 0, (cudaStream_t)stream>>>(c, n, npoints, points, idx, out);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                              float* grad_points, void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0) return TSM_OK;
    if (b > 65535) return TSM_ERR_INVALID;
    dim3 grid((unsigned)tsm::divup(npoints, tsm::GP_THREADS), (unsigned)b);
    tsm::gather_points_grad_kernel<<<grid, tsm::GP_THREADS, 0, (cudaStream_t)stream>>>(c, n, npoints, grad_out, idx,
                                                                                       grad_points);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_gather_xyz(int b, int n, int m, const float* xyz, const int* idx, float* out, void* stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (b > 65535) return TSM_ERR_INVALID;
    dim3 grid((unsigned)tsm::divup(m * 3, tsm::GP_THREADS), (unsigned)b);
    tsm::gather_xyz_kernel<<<grid, tsm::GP_THREADS, 0, (cudaStream_t)stream>>>(n, m, xyz, idx, out);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

}  // extern "C"
