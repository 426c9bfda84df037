// runtime.cu -- process-wide host helpers of the C-ABI library.
#include <mutex>

#include "common.cuh"

static std::mutex g_mu;
static int g_sms[64];
static int* g_status[64];

int tsm_num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        g_sms[dev] = v;
    }
    return g_sms[dev];
}

// One zero-initialised device int per device: kernels with in-kernel waits record a
// watchdog code here before trapping, so the host can tell a timeout from other faults.
int* tsm_status_word(cudaStream_t) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_status[dev]) {
        int* p = nullptr;
        if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(int));
        g_status[dev] = p;
    }
    return g_status[dev];
}

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
    cudaStream_t stream = nullptr;
    int tag = 0;
    unsigned long long tick = 0;
};
// Grow-only device scratch, one buffer per (device, stream, user tag) so that calls on
// different streams never share a buffer; stream-ordered (cudaMallocAsync), so repeated
// NMS calls allocate nothing.
constexpr int kScratchSlots = 256;  // (device, stream, tag) keys: 8 pipeline lanes x 3 streams x 6 tags and room to spare
static Scratch g_scratch[kScratchSlots];
static unsigned long long g_tick = 0;
static std::mutex g_scratch_mu;

int tsm_scratch_get(int tag, size_t bytes, cudaStream_t s, void** out) {
    int dev = 0;
    TSM_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    Scratch* sc = nullptr;
    for (auto& e : g_scratch)
        if (e.p && e.dev == dev && e.stream == s && e.tag == tag) sc = &e;
    if (!sc) {  // empty slot, else evict the least recently used one
        for (auto& e : g_scratch)
            if (!e.p && !sc) sc = &e;
        if (!sc) {
            sc = &g_scratch[0];
            for (auto& e : g_scratch)
                if (e.tick < sc->tick) sc = &e;
            sc->p = nullptr;  // retired (see below), never freed
            sc->cap = 0;
        }
        sc->dev = dev;
        sc->stream = s;
        sc->tag = tag;
    }
    sc->tick = ++g_tick;
    if (sc->cap < bytes) {
        // The outgrown buffer is RETIRED, not freed: a captured CUDA graph may still hold its address.
        sc->p = nullptr;
        sc->cap = 0;
        const size_t want = bytes + bytes / 4;
        TSM_CUDA_TRY(cudaMallocAsync(&sc->p, want, s));
        sc->cap = want;
    }
    *out = sc->p;
    return TSM_OK;
}


extern "C" {

const char* tsmdet_version() { return "tsmdet_b200 0.1 (sm_100a)"; }

// Human-readable text for a status returned by any tsmdet_* entry point.
const char* tsmdet_error_string(int code) {
    if (code == TSM_OK) return "ok";
    if (code == TSM_ERR_INVALID) return "tsmdet: invalid argument or unsupported size";
    if (code == TSM_ERR_WATCHDOG) return "tsmdet: in-kernel wait timed out";
    return cudaGetErrorString((cudaError_t)code);
}

// Reads (and clears) the watchdog status word of the current device; 0 = clean.
int tsmdet_read_status() {
    int* p = tsm_status_word(nullptr);
    if (!p) return 0;
    int v = 0;
    if (cudaMemcpy(&v, p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (v) cudaMemset(p, 0, sizeof(int));
    return v;
}

}  // extern "C"
