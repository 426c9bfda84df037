// ball_query.cu -- radius / annulus neighbour query for sm_100a.
//
// Replaces (same idx / idx_cnt, bit for bit):
//   ball_query_kernel_fast          /root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/ball_query_gpu.cu:75-112
//   ball_query_dilated_kernel_fast  ball_query_gpu.cu:138-176
//
// Reference semantics kept: hits are the FIRST `nsample` points in index order with
// d2 < r^2 (dilated: rin^2 <= d2 < rout^2), idx_cnt = min(hits, nsample), the row is
// padded cyclically with its own prefix, and a row without hits is all zeros.
//
// Design: one WARP per centre group (CW centres held in registers), lanes sweep 32
// consecutive points per step, so a hit's position in the row is
// cnt + popc(ballot & lanes_below): ordered compaction without atomics, early exit per
// warp as soon as its centres are full, and coalesced point traffic.  Point tiles are
// staged in shared memory by the TMA engine (cp.async.bulk, double buffered) in their
// native (N,3) layout -- a 3-float stride is conflict-free across the 32 banks.
#include "common.cuh"

namespace tsm {

constexpr int BQ_THREADS = 256;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_CW = 4;          // centres per warp
constexpr int BQ_TILE = 1024;     // points per shared-memory tile (12 KB)

template <bool DILATED>
__global__ void __launch_bounds__(BQ_THREADS)
    ball_query_kernel(int n, int m, float rin2, float rout2, int nsample, const float* __restrict__ new_xyz,
                      const float* __restrict__ xyz, int* __restrict__ idx_cnt, int* __restrict__ idx, int* status) {
    __shared__ __align__(128) float tile[2][BQ_TILE * 3];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    xyz += (size_t)b * n * 3;
    const int c0 = (blockIdx.x * BQ_WARPS + warp) * BQ_CW;  // first centre of this warp

    float cx[BQ_CW], cy[BQ_CW], cz[BQ_CW];
    int cnt[BQ_CW];
#pragma unroll
    for (int c = 0; c < BQ_CW; ++c) {
        const int ci = c0 + c;
        const bool ok = ci < m;
        const float* q = new_xyz + ((size_t)b * m + (ok ? ci : 0)) * 3;
        cx[c] = q[0];
        cy[c] = q[1];
        cz[c] = q[2];
        cnt[c] = ok ? 0 : nsample;  // out-of-range centres are "full" from the start
    }
    int* rows = idx + ((size_t)b * m + c0) * nsample;

    const int ntiles = divup(n, BQ_TILE);
    // TMA bulk copies need 16-byte aligned source and size; otherwise plain loads.
    const bool use_tma = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && ((n & 3) == 0);
    if (use_tma && tid == 0) {
        mbar_init(smem_u32(&full[0]), 1);
        mbar_init(smem_u32(&full[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int start = t * BQ_TILE;
        const int np = min(BQ_TILE, n - start);
        const uint32_t bytes = (uint32_t)np * 12u;
        mbar_arrive_expect_tx(smem_u32(&full[t & 1]), bytes);
        bulk_g2s(smem_u32(&tile[t & 1][0]), xyz + (size_t)start * 3, bytes, smem_u32(&full[t & 1]));
    };
    if (use_tma && tid == 0) issue(0);

    for (int t = 0; t < ntiles; ++t) {
        const int start = t * BQ_TILE;
        const int np = min(BQ_TILE, n - start);
        const float* tp = tile[t & 1];
        if (use_tma) {
            if (tid == 0 && t + 1 < ntiles) issue(t + 1);  // buffer (t+1)&1 was released by the barrier below
            const uint32_t bar = smem_u32(&full[t & 1]);
            const uint32_t ph = (uint32_t)((t >> 1) & 1);
            if (!mbar_try_wait_cta(bar, ph)) {
                const long long t0 = clock64();
                while (!mbar_try_wait_cta(bar, ph))
                    if (clock64() - t0 > 4000000000LL) watchdog_trip(status, TSM_ERR_WATCHDOG);
            }
        } else {
            for (int e = tid; e < np * 3; e += BQ_THREADS) tile[t & 1][e] = xyz[(size_t)start * 3 + e];
            __syncthreads();
        }

        bool active = false;
#pragma unroll
        for (int c = 0; c < BQ_CW; ++c) active |= cnt[c] < nsample;
        if (active) {
            for (int base = 0; base < np; base += 32) {
                const int pl = base + lane;
                const bool valid = pl < np;
                const float x = valid ? tp[pl * 3 + 0] : 0.f;
                const float y = valid ? tp[pl * 3 + 1] : 0.f;
                const float z = valid ? tp[pl * 3 + 2] : 0.f;
#pragma unroll
                for (int c = 0; c < BQ_CW; ++c) {
                    if (cnt[c] < nsample) {  // warp-uniform
                        const float d2 = sqdist3(x, y, z, cx[c], cy[c], cz[c]);
                        const bool hit = valid && (DILATED ? (d2 >= rin2 && d2 < rout2) : (d2 < rout2));
                        const unsigned bal = __ballot_sync(FULL, hit);
                        if (bal) {
                            const int pos = cnt[c] + __popc(bal & ((1u << lane) - 1u));
                            if (hit && pos < nsample) rows[(size_t)c * nsample + pos] = start + pl;
                            cnt[c] += __popc(bal);
                        }
                    }
                }
            }
        }
        // every warp must be done with tile[t&1] before it is refilled (tile t+2)
        bool still = false;
#pragma unroll
        for (int c = 0; c < BQ_CW; ++c) still |= cnt[c] < nsample;
        const int alldone = __syncthreads_and(!still);
        if (alldone) {
            if (use_tma && t + 1 < ntiles) {  // drain the prefetch in flight before the CTA retires
                const uint32_t bar = smem_u32(&full[(t + 1) & 1]);
                const uint32_t ph = (uint32_t)(((t + 1) >> 1) & 1);
                const long long t0 = clock64();
                while (!mbar_try_wait_cta(bar, ph))
                    if (clock64() - t0 > 4000000000LL) watchdog_trip(status, TSM_ERR_WATCHDOG);
            }
            break;
        }
    }

    // ---- counts + cyclic padding (row[p] = row[p mod cnt]); empty rows are zeroed
    __syncwarp();
#pragma unroll
    for (int c = 0; c < BQ_CW; ++c) {
        const int ci = c0 + c;
        if (ci >= m) continue;
        const int k = min(cnt[c], nsample);
        int* row = rows + (size_t)c * nsample;
        if (lane == 0) idx_cnt[(size_t)b * m + ci] = k;
        if (k == 0) {
            for (int p = lane; p < nsample; p += 32) row[p] = 0;
        } else if (k < nsample) {
            for (int p = k + lane; p < nsample; p += 32) row[p] = row[p % k];
        }
    }
}

}  // namespace tsm

static int run_ball_query(bool dilated, int b, int n, int m, float rin, float rout, int nsample, const float* new_xyz,
                          const float* xyz, int* idx_cnt, int* idx, cudaStream_t stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (n < 0 || nsample <= 0 || b > 65535) return TSM_ERR_INVALID;
    const float rin2 = rin * rin;    // f32 products, as ball_query_gpu.cu:91, 154-155
    const float rout2 = rout * rout;
    dim3 grid((unsigned)tsm::divup(m, tsm::BQ_WARPS * tsm::BQ_CW), (unsigned)b);
    int* status = tsm_status_word(stream);
    if (dilated)
        tsm::ball_query_kernel<true><<<grid, tsm::BQ_THREADS, 0, stream>>>(n, m, rin2, rout2, nsample, new_xyz, xyz,
                                                                            idx_cnt, idx, status);
    else
        tsm::ball_query_kernel<false><<<grid, tsm::BQ_THREADS, 0, stream>>>(n, m, rin2, rout2, nsample, new_xyz, xyz,
                                                                             idx_cnt, idx, status);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

extern "C" {

int tsmdet_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                      int* idx_cnt, int* idx, void* stream) {
    return run_ball_query(false, b, n, m, 0.f, radius, nsample, new_xyz, xyz, idx_cnt, idx, (cudaStream_t)stream);
}

int tsmdet_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample,
                              const float* new_xyz, const float* xyz, int* idx_cnt, int* idx, void* stream) {
    return run_ball_query(true, b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx,
                          (cudaStream_t)stream);
}

}  // extern "C"
