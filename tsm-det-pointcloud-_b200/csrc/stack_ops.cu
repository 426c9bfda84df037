// stack_ops.cu -- the pointnet2_stack ops that SA layers >= 1 and the head's VSA module call:
// voxel query (+ dilated), stacked grouping (+ grad) and stacked furthest point sampling.
//
// Reference (all under /root/reference/pcdet/ops/pointnet2/pointnet2_stack/src):
//   voxel_query_gpu.cu:10-98, 125-215   one THREAD per centre walks the (2r+1)^3 cells around it in the dense
//                                       voxel -> point table, serially (35 937 dependent loads at range 16), keeps the
//                                       first nsample hits and then replaces at random (cuRAND XORWOW, seed = centre)
//   group_points_gpu.cu:14-42, 66-95    one thread per output element, 4-byte gathers, batch found by a scan
//   sampling_gpu.cu:188-316             one 1024-thread CTA per cloud, points + min-distances in GLOBAL memory
//
// Here: one WARP per centre scans 32 cells per step (the x-runs of the table are contiguous -> coalesced), the hits
// of a step are consumed in lane order = the reference's scan order, and lane 0 replays the reference's sequential
// reservoir logic with the SAME generator (cuRAND's device API, curand_init(centre, 0, 0)) -- so indices, counts and
// every random replacement are bit-identical, which is what makes parity definable for this op at all.  Grouping
// transposes (nsample x C) tiles through shared memory (coalesced rows in, coalesced rows out).  Stack FPS keeps a
// cloud's points and min-distances in registers (<= 16 per thread) and reduces with the reference's tie rule for its
// fixed 1024-thread tree (the winner among equal maxima minimises bitrev10(k mod 1024), then k / 1024).
#include <curand_kernel.h>

#include "common.cuh"

namespace tsm {

constexpr int VQ_WARPS = 8;

template <bool DILATED>
__global__ void __launch_bounds__(VQ_WARPS * 32)
    voxel_query_warp_kernel(const int M, const int R1, const int R2, const int R3, const int nsample, const float former_radius,
                            const float radius, const int zr, const int yr, const int xr, const int zs, const int ys,
                            const int xs, const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                            const int* __restrict__ new_coords, const int* __restrict__ point_indices, int* __restrict__ idx,
                            int* __restrict__ cnt_unique, int* __restrict__ idx_cnt) {
    extern __shared__ int s_idx[];  // VQ_WARPS x nsample
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * VQ_WARPS + warp;
    if (pt >= M) return;  // whole warps leave; no block-wide barrier below
    int* const my = s_idx + warp * nsample;
    const float radius2 = __fmul_rn(radius, radius);
    const float former2 = __fmul_rn(former_radius, former_radius);
    const float nx = __ldg(new_xyz + pt * 3 + 0), ny = __ldg(new_xyz + pt * 3 + 1), nz = __ldg(new_xyz + pt * 3 + 2);
    const int b = __ldg(new_coords + pt * 4 + 0), cz = __ldg(new_coords + pt * 4 + 1), cy = __ldg(new_coords + pt * 4 + 2),
              cx = __ldg(new_coords + pt * 4 + 3);
    // cells per axis of "for (d = -r; d <= r; d += stride)"
    const int nzc = 2 * zr / zs + 1, nyc = 2 * yr / ys + 1, nxc = 2 * xr / xs + 1;
    const int plane = nyc * nxc;
    const int total = nzc * plane;
    curandState st;
    if (lane == 0) curand_init((unsigned long long)pt, 0, 0, &st);  // the reference's generator and seeding (:25-26)
    int cnt = 0, cnt2 = 0, range_point_num = 0;  // cnt / cnt2 live in lane 0
    for (int base = 0; base < total; base += 32) {
        const int c = base + lane;
        int nb = -1;
        if (c < total) {
            const int iz = c / plane, rem = c - iz * plane;
            const int iy = rem / nxc, ix = rem - iy * nxc;
            const int z = cz - zr + iz * zs, y = cy - yr + iy * ys, x = cx - xr + ix * xs;
            if (z >= 0 && z < R1 && y >= 0 && y < R2 && x >= 0 && x < R3)
                nb = __ldg(point_indices + (((size_t)b * R1 + z) * R2 + y) * R3 + x);
        }
        const bool nonempty = nb >= 0;
        bool hit = false;
        if (nonempty) {
            const float* p = xyz + (size_t)nb * 3;
            const float d2 = sqdist3(nx, ny, nz, __ldg(p + 0), __ldg(p + 1), __ldg(p + 2));  // fma(dz,dz,fma(dx,dx,dy*dy))
            hit = !(d2 > radius2 || (DILATED && d2 < former2));
        }
        range_point_num += __popc(__ballot_sync(FULL, nonempty));
        unsigned hm = __ballot_sync(FULL, hit);
        while (hm) {  // hits of this step in scan order
            const int l = __ffs(hm) - 1;
            hm &= hm - 1;
            const int v = __shfl_sync(FULL, nb, l);
            if (lane == 0) {
                ++cnt2;
                if (cnt < nsample) {
                    my[cnt++] = v;
                } else {
                    const float rnd = curand_uniform(&st);
                    if (rnd < ((float)nsample / (float)cnt2)) {
                        const int ins = (int)(ceilf(curand_uniform(&st) * (float)nsample) - 1.0f);
                        my[ins] = v;
                    }
                }
            }
        }
    }
    cnt = __shfl_sync(FULL, cnt, 0);
    __syncwarp();
    int* out = idx + (size_t)pt * nsample;
    if (cnt == 0) {
        if (lane == 0) out[0] = -1;  // the rest of the row keeps the caller's zeros, as in the reference
    } else {
        for (int i = lane; i < nsample; i += 32) out[i] = my[i < cnt ? i : i % cnt];  // cyclic pad (:93-95)
    }
    if (lane == 0) {
        cnt_unique[pt] = range_point_num;
        if (DILATED) idx_cnt[pt] = cnt;
    }
}

// which frame does stacked row `pt` belong to, and where do that frame's source rows start
__device__ __forceinline__ void stack_frame_of(const int B, const int* __restrict__ idx_batch_cnt,
                                               const int* __restrict__ features_batch_cnt, const int pt, int* start) {
    int bs = 0, acc = __ldg(idx_batch_cnt);
    for (int k = 1; k < B; ++k) {
        if (pt < acc) break;
        acc += __ldg(idx_batch_cnt + k);
        bs = k;
    }
    int s = 0;
    for (int k = 0; k < bs; ++k) s += __ldg(features_batch_cnt + k);
    *start = s;
}

// one CTA per stacked centre: (nsample x C) rows in -- coalesced along C -- (C x nsample) out -- coalesced along nsample
__global__ void __launch_bounds__(256) stack_group_points_kernel(const int B, const int M, const int C, const int nsample,
                                                                 const float* __restrict__ features,
                                                                 const int* __restrict__ features_batch_cnt,
                                                                 const int* __restrict__ idx, const int* __restrict__ idx_batch_cnt,
                                                                 float* __restrict__ out) {
    extern __shared__ float tile[];  // C x (nsample + 1)
    __shared__ int s_start;
    const int pt = blockIdx.x;
    if (threadIdx.x == 0) stack_frame_of(B, idx_batch_cnt, features_batch_cnt, pt, &s_start);
    __syncthreads();
    const int start = s_start;
    const int ld = nsample + 1;
    for (int e = threadIdx.x; e < nsample * C; e += 256) {
        const int s = e / C, c = e - s * C;
        const int row = start + __ldg(idx + (size_t)pt * nsample + s);
        tile[c * ld + s] = __ldg(features + (size_t)row * C + c);
    }
    __syncthreads();
    float* o = out + (size_t)pt * C * nsample;
    for (int e = threadIdx.x; e < nsample * C; e += 256) {
        const int c = e / nsample, s = e - c * nsample;
        o[e] = tile[c * ld + s];
    }
}

__global__ void __launch_bounds__(256) stack_group_points_grad_kernel(const int B, const int M, const int C, const int nsample,
                                                                      const float* __restrict__ grad_out,
                                                                      const int* __restrict__ idx,
                                                                      const int* __restrict__ idx_batch_cnt,
                                                                      const int* __restrict__ features_batch_cnt,
                                                                      float* __restrict__ grad_features) {
    __shared__ int s_start;
    const int pt = blockIdx.x;
    if (threadIdx.x == 0) stack_frame_of(B, idx_batch_cnt, features_batch_cnt, pt, &s_start);
    __syncthreads();
    const int start = s_start;
    const float* g = grad_out + (size_t)pt * C * nsample;
    for (int e = threadIdx.x; e < nsample * C; e += 256) {
        const int c = e / nsample, s = e - c * nsample;
        const int row = start + __ldg(idx + (size_t)pt * nsample + s);
        atomicAdd(grad_features + (size_t)row * C + c, __ldg(g + e));
    }
}

// Stack FPS: one 1024-thread CTA per cloud.  P > 0: the cloud (n <= 1024 P points) lives in registers; P == 0: in global
// memory, like the reference.  Thread t stands for the reference's thread t (its block size is 1024 whatever n is).
constexpr int SF_T = 1024;
template <int P>
__global__ void __launch_bounds__(SF_T) stack_fps_kernel(const int batch_size, const float* __restrict__ dataset,
                                                         float* __restrict__ temp, const int* __restrict__ xyz_batch_cnt,
                                                         int* __restrict__ idxs, const int* __restrict__ num_sampled_points) {
    __shared__ unsigned long long s_key[SF_T / 32];
    __shared__ float s_old[2][4];
    const int bs_idx = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int start = 0, ostart = 0;
    for (int k = 0; k < bs_idx; ++k) {
        start += __ldg(xyz_batch_cnt + k);
        ostart += __ldg(num_sampled_points + k);
    }
    const int n = __ldg(xyz_batch_cnt + bs_idx), m = __ldg(num_sampled_points + bs_idx);
    const float* pts = dataset + (size_t)start * 3;
    float* tp = temp + start;
    int* out = idxs + ostart;
    constexpr int PP = P > 0 ? P : 1;
    float px[PP], py[PP], pz[PP], td[PP];
    if (P > 0) {
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int k = t + SF_T * i;
            px[i] = py[i] = pz[i] = 0.f;
            td[i] = 0.f;
            if (k < n) {
                px[i] = __ldg(pts + k * 3 + 0);
                py[i] = __ldg(pts + k * 3 + 1);
                pz[i] = __ldg(pts + k * 3 + 2);
                td[i] = tp[k];
            }
        }
    }
    const unsigned rank = 1023u - (__brev((unsigned)t) >> 22);  // larger = earlier in the reference's tree
    if (t == 0) {
        out[0] = start;  // unconditional in the reference (:229)
        s_old[1][0] = __ldg(pts + 0);
        s_old[1][1] = __ldg(pts + 1);
        s_old[1][2] = __ldg(pts + 2);
    }
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        const float x1 = s_old[j & 1][0], y1 = s_old[j & 1][1], z1 = s_old[j & 1][2];
        float best = -1.f;
        int besti = 0;
        if (P > 0) {
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                if (t + SF_T * i < n) {
                    const float d2 = fminf(sqdist3(x1, y1, z1, px[i], py[i], pz[i]), td[i]);
                    td[i] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = i;
                    }
                }
            }
        } else {
            for (int k = t, i = 0; k < n; k += SF_T, ++i) {
                const float d2 = fminf(sqdist3(x1, y1, z1, __ldg(pts + k * 3 + 0), __ldg(pts + k * 3 + 1), __ldg(pts + k * 3 + 2)), tp[k]);
                tp[k] = d2;
                if (d2 > best) {
                    best = d2;
                    besti = i;
                }
            }
        }
        // (ordered distance | tree rank of the thread | slot): one 64-bit max = the reference's tree reduction
        unsigned long long key = ((unsigned long long)f32_ordered(best) << 32) | ((unsigned long long)rank << 22) |
                                 (unsigned long long)(unsigned)besti;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, key, d);
            key = o > key ? o : key;
        }
        if (lane == 0) s_key[warp] = key;
        __syncthreads();
        unsigned long long w = s_key[lane];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, w, d);
            w = o > w ? o : w;
        }
        const int tw = (int)(__brev(1023u - (unsigned)((w >> 22) & 1023u)) >> 22);
        const int iw = (int)(w & 0x3fffffu);
        const int old = tw + SF_T * iw;
        if (t == tw) {
            float ox, oy, oz;
            if (P > 0) {
                ox = px[0], oy = py[0], oz = pz[0];
#pragma unroll
                for (int i = 1; i < PP; ++i)
                    if (i == iw) ox = px[i], oy = py[i], oz = pz[i];
            } else {
                ox = __ldg(pts + old * 3 + 0), oy = __ldg(pts + old * 3 + 1), oz = __ldg(pts + old * 3 + 2);
            }
            s_old[(j + 1) & 1][0] = ox;
            s_old[(j + 1) & 1][1] = oy;
            s_old[(j + 1) & 1][2] = oz;
            out[j] = old + start;
        }
        __syncthreads();
    }
    if (P > 0) {  // temp is an output of the reference kernel too (the final min-distances)
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (t + SF_T * i < n) tp[t + SF_T * i] = td[i];
    }
}

}  // namespace tsm

// ref: pointnet2_stack/src/pointnet2_api.cpp:13-14 voxel_query_wrapper / voxel_query_dilated_wrapper
// (voxel_query.cpp:27-75).  former_radius is ignored by the plain query.  idx (M,nsample) must arrive zeroed.
extern "C" int tsmdet_voxel_query(int m, int r1, int r2, int r3, int nsample, float radius, int z_range, int y_range,
                                  int x_range, const float* new_xyz, const float* xyz, const int* new_coords,
                                  const int* point_indices, int* idx, int* cnt_unique, void* stream) {
    if (m <= 0) return TSM_OK;
    if (nsample <= 0 || nsample > 2048 || z_range < 0 || y_range < 0 || x_range < 0 || !new_xyz || !xyz || !new_coords ||
        !point_indices || !idx || !cnt_unique)
        return TSM_ERR_INVALID;
    const size_t smem = sizeof(int) * tsm::VQ_WARPS * nsample;
    if (smem > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::voxel_query_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tsm::voxel_query_warp_kernel<false><<<tsm::divup(m, tsm::VQ_WARPS), tsm::VQ_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        m, r1, r2, r3, nsample, 0.f, radius, z_range, y_range, x_range, 1, 1, 1, new_xyz, xyz, new_coords, point_indices, idx,
        cnt_unique, nullptr);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

extern "C" int tsmdet_voxel_query_dilated(int m, int r1, int r2, int r3, int nsample, float former_radius, float radius,
                                          int z_range, int y_range, int x_range, int z_stride, int y_stride, int x_stride,
                                          const float* new_xyz, const float* xyz, const int* new_coords,
                                          const int* point_indices, int* idx, int* cnt_unique, int* idx_cnt, void* stream) {
    if (m <= 0) return TSM_OK;
    if (nsample <= 0 || nsample > 2048 || z_range < 0 || y_range < 0 || x_range < 0 || z_stride < 1 || y_stride < 1 ||
        x_stride < 1 || !new_xyz || !xyz || !new_coords || !point_indices || !idx || !cnt_unique || !idx_cnt)
        return TSM_ERR_INVALID;  // (a stride of 0 never terminates in the reference)
    const size_t smem = sizeof(int) * tsm::VQ_WARPS * nsample;
    if (smem > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::voxel_query_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tsm::voxel_query_warp_kernel<true><<<tsm::divup(m, tsm::VQ_WARPS), tsm::VQ_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        m, r1, r2, r3, nsample, former_radius, radius, z_range, y_range, x_range, z_stride, y_stride, x_stride, new_xyz, xyz,
        new_coords, point_indices, idx, cnt_unique, idx_cnt);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

// ref: pointnet2_api.cpp:19-20 group_points(_grad)_wrapper (group_points.cpp; group_points_gpu.cu:14-95).
// features (N,C), idx (M,nsample) frame-local row numbers, *_batch_cnt (B) i32 on the device -> out (M,C,nsample).
extern "C" int tsmdet_stack_group_points(int b, int m, int c, int nsample, const float* features,
                                         const int* features_batch_cnt, const int* idx, const int* idx_batch_cnt, float* out,
                                         void* stream) {
    if (m <= 0 || c <= 0 || nsample <= 0) return TSM_OK;
    if (b <= 0 || !features || !features_batch_cnt || !idx || !idx_batch_cnt || !out) return TSM_ERR_INVALID;
    const size_t smem = sizeof(float) * (size_t)c * (nsample + 1);
    if (smem > 200 * 1024) return TSM_ERR_INVALID;
    if (smem > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::stack_group_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tsm::stack_group_points_kernel<<<m, 256, smem, static_cast<cudaStream_t>(stream)>>>(b, m, c, nsample, features,
                                                                                       features_batch_cnt, idx, idx_batch_cnt, out);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

extern "C" int tsmdet_stack_group_points_grad(int b, int m, int c, int n, int nsample, const float* grad_out, const int* idx,
                                              const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features,
                                              void* stream) {
    if (m <= 0 || c <= 0 || nsample <= 0) return TSM_OK;
    if (b <= 0 || !grad_out || !idx || !idx_batch_cnt || !features_batch_cnt || !grad_features) return TSM_ERR_INVALID;
    tsm::stack_group_points_grad_kernel<<<m, 256, 0, static_cast<cudaStream_t>(stream)>>>(b, m, c, nsample, grad_out, idx,
                                                                                         idx_batch_cnt, features_batch_cnt,
                                                                                         grad_features);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

// ref: pointnet2_api.cpp:17 stack_farthest_point_sampling_wrapper (sampling.cpp; sampling_gpu.cu:188-345).
// xyz (N,3) stacked, temp (N) pre-filled (1e10), xyz_batch_cnt (B) / num_sampled_points (B) i32 on the device,
// idxs (sum M) out: GLOBAL row numbers.  n_total = N (an upper bound of every cloud's size, as in the reference).
extern "C" int tsmdet_stack_farthest_point_sampling(int n_total, int batch_size, const float* xyz, float* temp,
                                                    const int* xyz_batch_cnt, int* idxs, const int* num_sampled_points,
                                                    void* stream) {
    if (batch_size <= 0) return TSM_OK;
    if (n_total <= 0 || !xyz || !temp || !xyz_batch_cnt || !idxs || !num_sampled_points) return TSM_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    using namespace tsm;
    const int per = divup(n_total, SF_T);
    if (per <= 1) stack_fps_kernel<1><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    else if (per <= 2) stack_fps_kernel<2><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    else if (per <= 4) stack_fps_kernel<4><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    else if (per <= 8) stack_fps_kernel<8><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    else if (per <= 16) stack_fps_kernel<16><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    else stack_fps_kernel<0><<<batch_size, SF_T, 0, s>>>(batch_size, xyz, temp, xyz_batch_cnt, idxs, num_sampled_points);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}
