// peer_put.cu -- the detection gather as ONE kernel of peer-memory stores (NVLink / NVSwitch), no rendezvous,
// with credit-based flow control.
//
// Replaces the reference's result merge, /root/reference/pcdet/utils/common_utils.py:224-245 (pickle files on a shared
// tmpdir + two barriers), and this repo's NCCL all_gather of the same packed records: an all_gather kernel holds SMs
// until every rank has launched its own, once per 0.6 ms step.  Here every rank owns a RING of `slots` receive buffers
// (each `world` rows) in its HBM and has every peer's ring mapped through CUDA IPC.  Step s (1, 2, ...) of a rank:
//
//   1. ack      : "everything before step s has been consumed here" -- the launch is stream-ordered after this rank's
//                 reads of step s-1, so it stores s-1 into word `rank` of every peer's ack array;
//   2. credit   : slot (s-1) % slots of consumer c last held this rank's step s-slots; the stores wait until
//                 acks[c] >= s-slots (acks = this rank's own array, written by the peers in their step 1);
//   3. stores   : this rank's packed records go into row `rank` of slot (s-1) % slots of EVERY rank, 16 bytes at a time;
//   4. publish  : every CTA fences (system scope) after its stores and counts itself done; the last one stores s into
//                 word `rank` of every rank's flag array.
//
// A consumer that sees flags[r] >= s (and fences) sees rank r's records of step s, and they stay intact until the
// consumer's own step s+1 has been launched (its ack) -- a fast rank can run at most slots-1 steps ahead of the slowest
// one, whatever the skew.  A wait that exceeds `timeout_ns` (a dead peer) does NOT trap: it records
// TSMDET_ERR_WATCHDOG in the host-mapped status word, skips the stores / the publish and returns, so the context
// survives and tsmdet_read_status() can tell the host what happened.
#include "common.cuh"

namespace tsm {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float* rows[kMaxPeers];        // rows[r]  = row `rank` of slot 0 of rank r's receive ring
    long long* flags[kMaxPeers];   // flags[r] = flag word `rank` of rank r
    long long* acks[kMaxPeers];    // acks[r]  = ack word `rank` of rank r (this rank as CONSUMER of r's records)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void status_set(int* status, int code) {
    if (status) *reinterpret_cast<volatile int*>(status) = code;
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
    peer_put_kernel(const float* __restrict__ src, const long long numel, const PeerPtrs pp, const int world,
                    const long long* __restrict__ my_acks, int* __restrict__ sync, const long long step,
                    const int slots, const long long slot_stride, const unsigned long long timeout_ns,
                    int* __restrict__ status) {
    // 1. ack (idempotent; one CTA is enough)
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        *reinterpret_cast<volatile long long*>(pp.acks[threadIdx.x]) = step - 1;
    }
    // 2. credit for the slot this step overwrites
    int ok = 1;
    if (threadIdx.x < world && step > slots) {
        const long long need = step - slots;
        const unsigned long long t0 = globaltimer_ns();
        while (*reinterpret_cast<const volatile long long*>(my_acks + threadIdx.x) < need) {
            __nanosleep(100);
            if (globaltimer_ns() - t0 > timeout_ns) {
                ok = 0;
                break;
            }
        }
    }
    ok = __syncthreads_and(ok);
    __threadfence_system();  // acquire side of the acks
    // 3. stores
    if (ok) {
        const long long off = ((step - 1) % slots) * slot_stride;
        const long long n4 = numel >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const float4 v = __ldg(s4 + i);
            for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(pp.rows[r] + off)[i] = v;
        }
        if (blockIdx.x == 0) {
            for (long long i = (n4 << 2) + threadIdx.x; i < numel; i += blockDim.x) {
                const float v = __ldg(src + i);
                for (int r = 0; r < world; ++r) pp.rows[r][off + i] = v;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    // 4. publish (sync[0] = CTAs done, sync[1] = CTAs that timed out; both return to 0)
    if (threadIdx.x == 0) {
        if (!ok) atomicAdd(&sync[1], 1);
        __threadfence();
        const int done = atomicAdd(&sync[0], 1);
        if (done == (int)gridDim.x - 1) {  // every CTA's stores are fenced
            __threadfence();
            const int failed = atomicExch(&sync[1], 0);
            sync[0] = 0;
            if (failed) {
                status_set(status, TSM_ERR_WATCHDOG);
            } else {
                __threadfence_system();
                for (int r = 0; r < world; ++r) *reinterpret_cast<volatile long long*>(pp.flags[r]) = step;
                __threadfence_system();
            }
        }
    }
}

// Stream-ordered consumer side: returns once flags[r] >= want for every r < world (one warp, lane r watches rank r).
// flags[r] >= want + slots would mean rank r has overwritten the slot being waited for: a protocol violation
// (impossible with the credits above), reported as TSM_ERR_INVALID in the status word.
__global__ void peer_wait_kernel(const long long* __restrict__ flags, const int world, const long long want,
                                 const int slots, const unsigned long long timeout_ns, int* __restrict__ status) {
    const int r = threadIdx.x;
    if (r < world) {
        const unsigned long long t0 = globaltimer_ns();
        long long f;
        while ((f = *reinterpret_cast<const volatile long long*>(flags + r)) < want) {
            __nanosleep(200);
            if (globaltimer_ns() - t0 > timeout_ns) {  // a peer died or stalled: report, do not hang or trap
                status_set(status, TSM_ERR_WATCHDOG);
                break;
            }
        }
        if (f >= want + slots) status_set(status, TSM_ERR_INVALID);
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

static unsigned long long timeout_or_default(long long timeout_ns) {
    return timeout_ns > 0 ? (unsigned long long)timeout_ns : 120000000000ull;  // 120 s
}

extern "C" int tsmdet_peer_wait(const long long* flags, int world, long long want, int slots, long long timeout_ns,
                                void* stream) {
    if (!flags || world < 1 || world > tsm::kMaxPeers || want < 0 || slots < 1) return TSM_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    tsm::peer_wait_kernel<<<1, 32, 0, s>>>(flags, world, want, slots, timeout_or_default(timeout_ns),
                                           tsm_status_word(s));
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

extern "C" int tsmdet_peer_put(const float* src, long long numel, int world, void* const* rows, void* const* flags,
                               void* const* peer_acks, const long long* my_acks, int* sync2, long long step, int slots,
                               long long slot_stride, long long timeout_ns, void* stream) {
    using namespace tsm;
    if (!src || numel <= 0 || world < 1 || world > kMaxPeers || !rows || !flags || !peer_acks || !my_acks || !sync2)
        return TSM_ERR_INVALID;
    if (step < 1 || slots < 1 || slot_stride < numel || (slot_stride & 3)) return TSM_ERR_INVALID;
    PeerPtrs pp;
    uintptr_t align = reinterpret_cast<uintptr_t>(src);
    for (int r = 0; r < kMaxPeers; ++r) {
        pp.rows[r] = r < world ? static_cast<float*>(rows[r]) : nullptr;
        pp.flags[r] = r < world ? static_cast<long long*>(flags[r]) : nullptr;
        pp.acks[r] = r < world ? static_cast<long long*>(peer_acks[r]) : nullptr;
        if (r < world) {
            if (!rows[r] || !flags[r] || !peer_acks[r]) return TSM_ERR_INVALID;
            align |= reinterpret_cast<uintptr_t>(rows[r]);
        }
    }
    if (align & 15u) return TSM_ERR_INVALID;  // rows and source must allow 16-byte stores
    // few CTAs: a launch that has to wait for a credit spins, and spinning CTAs hold SM slots that the step's own
    // kernels need (8 x 256 threads are plenty for 295 KB x world destinations)
    long long blocks = ((numel >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 8) blocks = 8;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    peer_put_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, numel, pp, world, my_acks, sync2, step, slots, slot_stride,
                                                     timeout_or_default(timeout_ns), tsm_status_word(s));
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}
