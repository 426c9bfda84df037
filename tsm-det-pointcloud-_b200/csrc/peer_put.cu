// peer_put.cu -- the detection gather as ONE kernel of peer-memory stores (NVLink / NVSwitch), no rendezvous.
//
// Replaces the reference's result merge, /root/reference/pcdet/utils/common_utils.py:224-245 (pickle files on a shared
// tmpdir + two barriers), and this repo's NCCL all_gather of the same packed records: an all_gather kernel holds SMs
// until every rank has launched its own, once per 0.6 ms step.  Here every rank owns a receive buffer (world rows) in
// its HBM and has every peer's buffer mapped through CUDA IPC; one launch per step stores this rank's packed records
// into row `rank` of EVERY rank's buffer with 16-byte stores, and the last CTA to finish publishes the step number
// into every rank's flag word for this rank.  Ordering: each CTA fences (system scope) after its data stores and
// before it counts itself done; the last CTA fences again before the flag stores -- so a reader that sees
// flags[r] >= s (and fences) sees rank r's records of step s.  Nothing waits for a peer.
#include "common.cuh"

namespace tsm {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float* rows[kMaxPeers];        // rows[r]  = row `rank` of rank r's receive buffer
    long long* flags[kMaxPeers];   // flags[r] = flag word `rank` of rank r
};

__global__ void __launch_bounds__(256)
    peer_put_kernel(const float* __restrict__ src, const long long numel, const PeerPtrs pp, const int world,
                    int* __restrict__ sync) {
    const long long n4 = numel >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(s4 + i);
        for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(pp.rows[r])[i] = v;
    }
    if (blockIdx.x == 0) {
        for (long long i = (n4 << 2) + threadIdx.x; i < numel; i += blockDim.x) {
            const float v = __ldg(src + i);
            for (int r = 0; r < world; ++r) pp.rows[r][i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&sync[0], 1);
        if (done == (int)gridDim.x - 1) {  // every CTA's stores are fenced: publish the step
            sync[0] = 0;
            const long long tick = (long long)(++sync[1]);
            __threadfence_system();
            for (int r = 0; r < world; ++r) *reinterpret_cast<volatile long long*>(pp.flags[r]) = tick;
            __threadfence_system();
        }
    }
}

// Stream-ordered consumer side: returns once flags[r] >= want for every r < world (one warp, lane r watches rank r).
__global__ void peer_wait_kernel(const long long* __restrict__ flags, const int world, const int* __restrict__ sync,
                                 const long long want_or_neg, int* __restrict__ status) {
    const long long want = want_or_neg >= 0 ? want_or_neg : (long long)*reinterpret_cast<const volatile int*>(sync + 1);
    const int r = threadIdx.x;
    const long long t0 = clock64();
    if (r < world) {
        while (*reinterpret_cast<const volatile long long*>(flags + r) < want) {
            __nanosleep(200);
            if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer died; report instead of hanging the stream
                watchdog_trip(status, TSM_ERR_WATCHDOG);
                break;
            }
        }
    }
    __threadfence_system();
}

}  // namespace tsm

// Peer access from the current device to `peer_device` (kernels of this process storing into memory that lives there).
extern "C" int tsmdet_enable_peer_access(int peer_device) {
    int dev = 0, can = 0;
    TSM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev == peer_device) return TSM_OK;
    TSM_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return TSM_ERR_INVALID;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return TSM_OK;
    }
    return e == cudaSuccess ? TSM_OK : (int)e;
}

extern "C" int tsmdet_peer_wait(const long long* flags, int world, const int* sync2, long long want, void* stream) {
    if (!flags || world < 1 || world > 32 || !sync2) return TSM_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    tsm::peer_wait_kernel<<<1, 32, 0, s>>>(flags, world, sync2, want, tsm_status_word(s));
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

extern "C" int tsmdet_peer_put(const float* src, long long numel, int world, void* const* rows, void* const* flags,
                               int* sync2, void* stream) {
    using namespace tsm;
    if (!src || numel <= 0 || world < 1 || world > kMaxPeers || !rows || !flags || !sync2) return TSM_ERR_INVALID;
    PeerPtrs pp;
    uintptr_t align = reinterpret_cast<uintptr_t>(src);
    for (int r = 0; r < kMaxPeers; ++r) {
        pp.rows[r] = r < world ? static_cast<float*>(rows[r]) : nullptr;
        pp.flags[r] = r < world ? static_cast<long long*>(flags[r]) : nullptr;
        if (r < world) {
            if (!rows[r] || !flags[r]) return TSM_ERR_INVALID;
            align |= reinterpret_cast<uintptr_t>(rows[r]);
        }
    }
    if (align & 15u) return TSM_ERR_INVALID;  // rows and source must allow 16-byte stores
    long long blocks = ((numel >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    peer_put_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, numel, pp, world, sync2);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}
