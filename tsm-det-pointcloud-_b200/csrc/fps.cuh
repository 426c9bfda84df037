// fps.cuh -- shared between the two furthest-point-sampling kernels (fps.cu: cluster / register-resident brute
// force, any N; fps_bucket.cu: one CTA per cloud, spatially bucketed with exact pruning, N <= 16384).
#pragma once
#include "common.cuh"

namespace tsm {

struct FpsArgs {
    const float* xyz;      // (B,N,3)
    const float* weights;  // (B,N) or nullptr
    float* temp;           // (B,N) or nullptr; in: initial min-dist, out: final min-dist
    int* idxs;             // (B,M)
    int n, m;
    int log2bs;  // log2 of the reference block size (cuda_utils.h:10-14)
    int* status;
    // ---- chaining (optional; see tsmdet_fps_chain): per-cloud tie/value records of this run ...
    int* tie_iter;     // (B) out: first iteration whose maximum was shared by points with different coordinates
    float* vals;       // (B,M) out: the winning min-distance of every iteration (vals[0] = +inf)
    // ... and of the run that produced this cloud as ITS first parent_m picks, in order
    const int* parent_tie;
    const float* parent_vals;
    int parent_m;
};

// (u desc, rank asc) argmax across the warp.  Returns the winning lane; wu / wrk are warp-uniform.
// The common case (a unique maximum) costs one redux + one ballot.
__device__ __forceinline__ int warp_pick(uint32_t u, uint32_t rk, uint32_t& wu, uint32_t& wrk) {
    wu = __reduce_max_sync(FULL, u);
    const unsigned tie = __ballot_sync(FULL, u == wu);
    if (__popc(tie) == 1) {
        const int wl = __ffs(tie) - 1;
        wrk = __shfl_sync(FULL, rk, wl);
        return wl;
    }
    wrk = __reduce_min_sync(FULL, (u == wu) ? rk : 0xffffffffu);
    return __ffs(__ballot_sync(FULL, u == wu && rk == wrk)) - 1;
}

}  // namespace tsm

// fps_bucket.cu.  Returns TSM_ERR_INVALID when the shape is outside what the bucketed kernel holds in one CTA.
bool tsm_fps_bucket_supports(int n, bool weighted);
int tsm_fps_bucket_launch(const tsm::FpsArgs& a, int b, cudaStream_t stream);
// fps_bucket_cluster.cu: the same sampler over a cluster of CTAs, 16384 < N <= 65536.
bool tsm_fps_bucket_cluster_supports(int n, bool weighted);
int tsm_fps_bucket_cluster_launch(const tsm::FpsArgs& a, int b, cudaStream_t stream);
