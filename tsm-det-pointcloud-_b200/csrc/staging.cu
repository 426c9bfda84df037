// staging.cu -- input staging for the set-abstraction path.
//
// The reference moves a batch to the GPU as ONE collated array `points (sum N, 1+3+C)` =
// [batch_idx, x, y, z, features...] (pcdet/models/__init__.py:23-34 load_data_to_gpu) and then re-lays it out
// with five separate torch copies before the first SA layer can run:
//   break_up_pc (pointnet2_backbone.py:796-800): xyz = pc[:,1:4].contiguous(); features = pc[:,4:].contiguous()
//   forward     (:814-823): per-frame counts via `(batch_idx == b).sum()` + assert equal, xyz.view(B,-1,3),
//                features.view(B,-1,C).permute(0,2,1).contiguous()
//   SA layer 0  (pointnet2_modules.py:1143): xyz.transpose(1,2).contiguous()  (only to gather the centres)
// Here one kernel reads the collated array once and leaves the layouts the kernels of this library consume
// resident in HBM: xyz (B,N,3) and features (B,C,N), and checks the equal-count assertion on the device
// (rows whose batch index is not their frame are counted into *bad; no host synchronisation).
// HBM-bound: 4(4+C) bytes read + 4(3+C) bytes written per point.
#include "common.cuh"

namespace tsm {

// One CTA stages 256 consecutive points of one frame through shared memory: the collated rows are read as one
// contiguous run of 256*(4+C) floats (fully coalesced, 16-byte vector loads when aligned), and the three outputs
// are written as contiguous runs (xyz: 768 floats; every feature channel: 256 floats).
template <int TILE>
__global__ void __launch_bounds__(TILE) stage_points_kernel(const int n, const int c, const float* __restrict__ points,
                                                            float* __restrict__ xyz, float* __restrict__ features,
                                                            int* __restrict__ bad) {
    extern __shared__ float tile[];  // TILE * (4 + c) floats
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * TILE;
    const int np = min(TILE, n - p0);
    const int w = 4 + c;
    const long long row0 = (long long)b * n + p0;
    const float* src = points + row0 * w;
    const int total = np * w;
    // contiguous run -> shared memory
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int t4 = total >> 2;
        for (int i = threadIdx.x; i < t4; i += TILE) reinterpret_cast<float4*>(tile)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = (t4 << 2) + threadIdx.x; i < total; i += TILE) tile[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < total; i += TILE) tile[i] = __ldg(src + i);
    }
    __syncthreads();
    // batch-index check (the reference's `assert xyz_batch_cnt.min() == xyz_batch_cnt.max()` plus frame order)
    int wrong = 0;
    if (threadIdx.x < np) wrong = tile[threadIdx.x * w] != (float)b;
    wrong = __syncthreads_or(wrong);
    if (wrong && threadIdx.x == 0 && bad) atomicAdd(bad, 1);
    // xyz (B,N,3): 3*np contiguous floats
    float* dx = xyz + row0 * 3;
    for (int i = threadIdx.x; i < 3 * np; i += TILE) {
        const int p = i / 3, k = i - 3 * p;
        dx[i] = tile[p * w + 1 + k];
    }
    // features (B,C,N): per channel np contiguous floats (row stride w in shared memory: w odd or small -> few conflicts)
    if (features) {
        for (int ch = 0; ch < c; ++ch) {
            float* df = features + ((long long)b * c + ch) * n + p0;
            if (threadIdx.x < np) df[threadIdx.x] = tile[threadIdx.x * w + 4 + ch];
        }
    }
}

}  // namespace tsm

// points (b*n, 4+c) f32 collated, frame-major -> xyz (b,n,3), features (b,c,n) [NULL when c == 0],
// bad (1) i32 device counter (NOT cleared here) [may be NULL: no check output].
// ref: pcdet/models/backbones_3d/pointnet2_backbone.py:796-800, 814-823; pointnet2_modules.py:1143
extern "C" int tsmdet_stage_points(int b, int n, int c, const float* points, float* xyz, float* features, int* bad,
                                   void* stream) {
    if (b <= 0 || n <= 0) return TSM_OK;
    if (c < 0 || c > 60 || !points || !xyz || (c > 0 && !features)) return TSM_ERR_INVALID;
    constexpr int TILE = 256;
    dim3 grid((unsigned)tsm::divup(n, TILE), (unsigned)b);
    const size_t smem = (size_t)TILE * (4 + c) * sizeof(float);
    if (smem > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::stage_points_kernel<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tsm::stage_points_kernel<TILE><<<grid, TILE, smem, static_cast<cudaStream_t>(stream)>>>(n, c, points, xyz, features, bad);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}
