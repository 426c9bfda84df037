// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces (same results, bit for bit):
//   farthest_point_sampling_kernel / furthest_point_sampling_kernel
//       /root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:100-216, 588-704
//   furthest_point_sampling_weights_kernel                       sampling_gpu.cu:901-1022
//   furthest_point_sampling_matrix_kernel / _with_dist_kernel    sampling_gpu.cu:750-855, 262-377
//   furthest_point_sampling_with_weighted_dist_kernel            sampling_gpu.cu:424-541
//
// Design (B200-first, not a translation):
//  * One thread-block CLUSTER per cloud (1..16 CTAs).  Every point and its running
//    min-distance live in REGISTERS for the whole kernel (shared memory for the
//    coordinates only when a CTA owns more than 8192 points); HBM is touched once
//    to load the cloud and once per selected index.
//  * Per iteration: P fused distance/min/argmax updates per thread, a two-instruction
//    warp argmax (redux.sync max on the order-preserving key, redux.sync min on the
//    tie-break rank), one CTA barrier, then the per-CTA winners are exchanged with
//    st.async (DSMEM store + remote mbarrier complete_tx) -- no cluster-wide barrier
//    on the critical path.
//  * Tie-breaking reproduces the reference exactly.  The reference's thread `t`
//    scans k = t, t+bs, ... keeping the first strict maximum, and its shared-memory
//    tree keeps the lower slot on ties, which orders threads by the BIT-REVERSAL of
//    t.  So among equal maxima the winner minimises
//        rank(k) = (bitrev_L(k mod bs) << (32-L)) | (k / bs),  bs = 2^L = opt_n_threads(N)
//    and the whole argmax is a max over the 64-bit key (ordered(value), ~rank).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fps.cuh"

namespace tsm {

struct __align__(16) FpsRec {
    uint32_t u, rank;
    float x, y;
    float z;
    uint32_t pad[3];
};
static_assert(sizeof(FpsRec) == 32, "FpsRec must be 32 bytes");

constexpr uint32_t kRecBytes = 20;  // u, rank, x, y (v4) + z (b32)
constexpr int kMaxRecs = 512;       // warps per cluster (cluster size x warps per CTA)

__device__ __forceinline__ void wait_records(uint32_t bar, uint32_t phase, int* status) {
    if (mbar_try_wait_cta(bar, phase)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cta(bar, phase)) {
        if (clock64() - t0 > 4000000000LL) watchdog_trip(status, TSM_ERR_WATCHDOG);
    }
}

// Per-thread argmax over P register slots as a balanced tree (short dependency chains; one warp per
// SM sub-partition has to hide its own latency).  Ties keep the lower slot = lower k, as the
// reference thread's ascending strict-> scan does.
template <int P>
__device__ __forceinline__ void slot_argmax(const float (&v)[P], float& best, int& bp) {
    float val[P];
    int idx[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        val[p] = v[p];
        idx[p] = p;
    }
#pragma unroll
    for (int w = 1; w < P; w <<= 1) {
#pragma unroll
        for (int p = 0; p + w < P; p += 2 * w) {
            const bool take = val[p + w] > val[p];  // strict: the lower slot wins ties
            val[p] = take ? val[p + w] : val[p];
            idx[p] = take ? idx[p + w] : idx[p];
        }
    }
    best = val[0];
    bp = idx[0];
}

// One cluster (1..16 CTAs) per cloud.  No CTA barrier inside the loop: every WARP publishes its
// own candidate to every CTA of the cluster (st.async into recs[parity][warp id in cluster] +
// complete_tx on that CTA's mbarrier), and every warp reduces the R = csize * NW records itself.
template <int T, int P, bool WEIGHTED, bool SMEM, bool CHAIN>
__global__ void __launch_bounds__(T, 1) fps_kernel(const FpsArgs a) {
    constexpr int NW = T / 32;
    __shared__ __align__(8) uint64_t mbar[2];
    // dynamic smem: recs[2][R] (32 B each), then 3 * P * T floats: SoA planes of this CTA's points
    extern __shared__ __align__(16) unsigned char dyn_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = cluster_ctarank();
    const uint32_t csize = cluster_nctarank();
    const int R = (int)csize * NW;  // records per iteration
    FpsRec* recs0 = reinterpret_cast<FpsRec*>(dyn_smem);
    FpsRec* recs1 = recs0 + R;
    float* dyn_xyz = reinterpret_cast<float*>(recs1 + R);
    const int cloud = blockIdx.x / csize;
    const int L = a.log2bs;
    const int bs = 1 << L;
    const int n = a.n, m = a.m;
    const int g = (int)crank * T + tid;  // thread id within the cluster
    const int r = g & (bs - 1);          // reference thread this thread stands in for
    const int q = g >> L;                // which chunk of that thread's strided scan
    const float* __restrict__ xyz = a.xyz + (size_t)cloud * n * 3;
    const uint32_t lowmask = (L == 0) ? 0xffffffffu : ((1u << (32 - L)) - 1u);
    const uint32_t rbase = ((L == 0) ? 0u : __brev((uint32_t)r)) | (uint32_t)(q * P);

    // ---- chained shortcut: this cloud is the first `n` picks (in order) of an FPS run over a superset.
    // FPS over such a prefix set re-selects it in the same order (each pick maximises the same min-distance
    // over a subset that still contains the maximiser) unless a maximum was shared by distinct points, or
    // zero-distance picks put duplicate coordinates into the set.  Both are recorded by the parent run, so
    // the common case costs a few microseconds instead of m latency-bound iterations.
    if (CHAIN && a.parent_tie != nullptr) {
        const bool prefix_ok = a.parent_tie[cloud] >= m && m <= n && n <= a.parent_m &&
                               a.parent_vals[(size_t)cloud * a.parent_m + a.parent_m - 1] > 0.f;
        if (prefix_ok) {  // uniform across the cluster: nobody reaches the cluster barriers below
            for (int j = g; j < m; j += (int)csize * T) {
                a.idxs[(size_t)cloud * m + j] = j;
                if (a.vals) a.vals[(size_t)cloud * m + j] = a.parent_vals[(size_t)cloud * a.parent_m + j];
            }
            if (g == 0 && a.tie_iter) a.tie_iter[cloud] = a.parent_tie[cloud];
            return;
        }
    }
    if (CHAIN && g == 0 && a.tie_iter) a.tie_iter[cloud] = 0x7fffffff;
    if (CHAIN && g == 0 && a.vals && m > 0) a.vals[(size_t)cloud * m] = __int_as_float(0x7f800000);

    if (tid == 0) {
        mbar_init(smem_u32(&mbar[0]), 1);
        mbar_init(smem_u32(&mbar[1]), 1);
        mbar_fence_init_cluster();
    }

    float px[SMEM ? 1 : P], py[SMEM ? 1 : P], pz[SMEM ? 1 : P];
    float md[P];
    float wf[WEIGHTED ? P : 1];
    float* sx = dyn_xyz;
    float* sy = dyn_xyz + P * T;
    float* sz = dyn_xyz + 2 * P * T;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const long long k = (long long)r + ((long long)(q * P + p) << L);
        float x = 0.f, y = 0.f, z = 0.f, d0 = -2.f, w0 = -2.f;
        if (k < n) {
            x = __ldg(xyz + k * 3 + 0);
            y = __ldg(xyz + k * 3 + 1);
            z = __ldg(xyz + k * 3 + 2);
            d0 = a.temp ? a.temp[(size_t)cloud * n + k] : 1e10f;
            if (WEIGHTED) w0 = __ldg(a.weights + (size_t)cloud * n + k);
        }
        sx[p * T + tid] = x;  // every thread only ever reads back its own slots
        sy[p * T + tid] = y;
        sz[p * T + tid] = z;
        if (!SMEM) {
            px[p] = x;
            py[p] = y;
            pz[p] = z;
        }
        md[p] = d0;
        if (WEIGHTED) wf[p] = w0;
    }

    float x1 = __ldg(xyz + 0), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
    int* __restrict__ idxs = a.idxs + (size_t)cloud * m;
    if (!WEIGHTED && g == 0 && m > 0) idxs[0] = 0;

    // destination of this warp's record in CTA `lane` (lanes < csize send): slot = cluster warp id
    const uint32_t my_slot = crank * NW + warp;
    const uint32_t peer = (uint32_t)lane < csize ? (uint32_t)lane : 0u;
    const uint32_t dst_rec0 = mapa_u32(smem_u32(&recs0[my_slot]), peer);
    const uint32_t dst_rec1 = mapa_u32(smem_u32(&recs1[my_slot]), peer);
    const uint32_t dst_bar0 = mapa_u32(smem_u32(&mbar[0]), peer);
    const uint32_t dst_bar1 = mapa_u32(smem_u32(&mbar[1]), peer);

    cluster_sync_all();  // peers' mbarriers are initialised past this point

    // chain bookkeeping of iteration j is carried out one iteration later, inside that iteration's wait window
    bool pend = false, pend_dup = false;
    uint32_t pend_u = 0u, pend_gu = 0u;
    float pend_bx = 0.f, pend_by = 0.f, pend_bz = 0.f, pend_x1 = 0.f, pend_y1 = 0.f, pend_z1 = 0.f;
    int pend_j = 0;
    auto settle = [&]() {
        if (CHAIN && pend) {
            if (a.tie_iter && pend_u == pend_gu && pend_gu != 0u &&
                (pend_dup || pend_bx != pend_x1 || pend_by != pend_y1 || pend_bz != pend_z1))
                atomicMin(a.tie_iter + cloud, pend_j);
            if (g == 0 && a.vals) a.vals[(size_t)cloud * m + pend_j] = __uint_as_float(pend_gu & 0x7fffffffu);
        }
    };

    int it = 0;
    for (int j = WEIGHTED ? 0 : 1; j < m; ++j, ++it) {
        const int par = it & 1;
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&mbar[par]), (uint32_t)R * kRecBytes);

        // ---- per-thread: update the running min-distance of every owned point, then argmax
        float score[P];
        if (WEIGHTED && j == 0) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float v = wf[p] + 0.0f;  // canonicalise -0.0
                score[p] = (md[p] != -2.f && v > -1.f) ? v : -2.f;
            }
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float X = SMEM ? sx[p * T + tid] : px[p];
                const float Y = SMEM ? sy[p * T + tid] : py[p];
                const float Z = SMEM ? sz[p * T + tid] : pz[p];
                const float d = sqdist3(x1, y1, z1, X, Y, Z);
                const float mm = fminf(d, md[p]);  // invalid slots stay at -2
                md[p] = mm;
                score[p] = mm;
                if (WEIGHTED) score[p] = (mm == -2.f) ? -2.f : (float)((double)mm * fmax((double)wf[p], 1e-12));
            }
        }
        float best;
        int bp;
        slot_argmax<P>(score, best, bp);
        // candidate coordinates: read back this thread's own best slot (conflict-free planes)
        const float bx = sx[bp * T + tid], by = sy[bp * T + tid], bz = sz[bp * T + tid];
        const uint32_t u = (best > -1.f) ? f32_ordered(best) : 0u;  // the reference starts from best = -1
        const uint32_t rk = rbase + (uint32_t)bp;

        // ---- warp argmax; the warp's record goes to every CTA of the cluster
        uint32_t wu, wrk;
        const int wl = warp_pick(u, rk, wu, wrk);
        const float wx = __shfl_sync(FULL, bx, wl), wy = __shfl_sync(FULL, by, wl), wz = __shfl_sync(FULL, bz, wl);
        if ((uint32_t)lane < csize) {
            const uint32_t dr = par ? dst_rec1 : dst_rec0, db = par ? dst_bar1 : dst_bar0;
            st_async_v4(dr, db, wu, wrk, __float_as_uint(wx), __float_as_uint(wy));
            st_async_b32(dr + 16, db, __float_as_uint(wz));
        }

        // ---- while the records are in flight (the warp would idle here): does this thread hold a SECOND slot
        // at its own maximum?  (Chain bookkeeping only; treated as a tie whatever its coordinates -- two of one
        // thread's strided points sharing the cloud-wide maximum is rare enough to just run the full sampler.)
        bool mydup = false;
        if (CHAIN && a.tie_iter) {
            float bscan = best;
            asm volatile("" : "+f"(bscan));  // keep the scan below the send
            int eqc = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) eqc += (score[p] == bscan) ? 1 : 0;
            mydup = eqc > 1;
            settle();  // previous iteration's record keeping
        }

        // ---- all R records of this iteration
        wait_records(smem_u32(&mbar[par]), (uint32_t)((it >> 1) & 1), a.status);
        uint32_t cu = 0u, crk = 0xffffffffu;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        const FpsRec* recs = par ? recs1 : recs0;
        if (lane < R) {  // R <= 32 in the common shapes: one record per lane
            const uint4 v = *reinterpret_cast<const uint4*>(&recs[lane]);
            cz = recs[lane].z;
            cu = v.x;
            crk = v.y;
            cx = __uint_as_float(v.z);
            cy = __uint_as_float(v.w);
        }
#pragma unroll 1
        for (int s = lane + 32; s < R; s += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(&recs[s]);
            if (v.x > cu || (v.x == cu && v.y < crk)) {
                cu = v.x;
                crk = v.y;
                cx = __uint_as_float(v.z);
                cy = __uint_as_float(v.w);
                cz = recs[s].z;
            }
        }
        uint32_t gu, grk;
        const int gl = warp_pick(cu, crk, gu, grk);
        x1 = __shfl_sync(FULL, cx, gl);
        y1 = __shfl_sync(FULL, cy, gl);
        z1 = __shfl_sync(FULL, cz, gl);

        if (gu == 0u) {  // no eligible candidate anywhere: the reference yields index 0
            x1 = __ldg(xyz + 0);
            y1 = __ldg(xyz + 1);
            z1 = __ldg(xyz + 2);
        }
        if (g == 0) {
            int k = 0;
            if (gu != 0u) {
                const uint32_t rr = (L == 0) ? 0u : __brev(grk & ~lowmask);
                k = (int)(rr + ((grk & lowmask) << L));
            }
            idxs[j] = k;
        }
        if (CHAIN) {
            // did another point with DIFFERENT coordinates share this iteration's maximum?  Every thread compares
            // its own candidate with the winner (settled in the next wait window; a second maximum inside one
            // thread was found by the scan above).
            pend = true; pend_dup = mydup; pend_u = u; pend_gu = gu; pend_j = j;
            pend_bx = bx; pend_by = by; pend_bz = bz; pend_x1 = x1; pend_y1 = y1; pend_z1 = z1;
        }
    }
    settle();

    if (a.temp) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const long long k = (long long)r + ((long long)(q * P + p) << L);
            if (k < n) a.temp[(size_t)cloud * n + k] = md[p];
        }
    }
    cluster_sync_all();  // no CTA leaves while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------
// Distance-matrix variants ('f-fps'): the distances come from row `old` of a (N,N)
// matrix, so each iteration is one coalesced row read from L2/HBM; min-distances stay
// in the caller's temp (any N).  One CTA per cloud; same (value, rank) argmax.
template <bool WEIGHTED>
__global__ void __launch_bounds__(1024, 1)
    fps_matrix_kernel(const float* __restrict__ matrix, const float* __restrict__ weights, float* __restrict__ temp,
                      int* __restrict__ idxs, int n, int m, int log2bs) {
    constexpr int T = 1024, NW = 32;
    __shared__ FpsRec slots[2][NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const int L = log2bs, bs = 1 << L;
    matrix += (size_t)cloud * n * n;
    temp += (size_t)cloud * n;
    idxs += (size_t)cloud * m;
    if (WEIGHTED) weights += (size_t)cloud * n;
    const uint32_t lowmask = (L == 0) ? 0xffffffffu : ((1u << (32 - L)) - 1u);
    if (!WEIGHTED && tid == 0 && m > 0) idxs[0] = 0;
    int old = 0;
    int it = 0;
    for (int j = WEIGHTED ? 0 : 1; j < m; ++j, ++it) {
        const int par = it & 1;
        uint32_t u = 0u, rk = 0xffffffffu;
        // thread `tid` plays reference threads r = tid, tid+T, ... (only r = tid when bs <= T)
        for (int r = tid; r < bs; r += T) {
            float best = -1.f;
            int bk = 0;
            for (int k = r; k < n; k += bs) {
                float score;
                if (WEIGHTED && j == 0) {
                    score = weights[k] + 0.0f;
                } else {
                    const float d = fminf(matrix[(size_t)old * n + k], temp[k]);
                    temp[k] = d;
                    score = WEIGHTED ? (float)((double)d * fmax((double)weights[k], 1e-12)) : d;
                }
                if (score > best) {
                    best = score;
                    bk = k;
                }
            }
            const uint32_t cu = (best > -1.f) ? f32_ordered(best) : 0u;
            const uint32_t crk = ((L == 0) ? 0u : __brev((uint32_t)r)) | (uint32_t)(bk >> L);
            if (cu > u || (cu == u && crk < rk)) {
                u = cu;
                rk = crk;
            }
        }
        uint32_t wu, wrk;
        if (warp_pick(u, rk, wu, wrk) == lane) {
            slots[par][warp].u = wu;
            slots[par][warp].rank = wrk;
        }
        __syncthreads();
        uint32_t gu, grk;
        warp_pick(slots[par][lane].u, slots[par][lane].rank, gu, grk);
        old = 0;
        if (gu != 0u) {
            const uint32_t rr = (L == 0) ? 0u : __brev(grk & ~lowmask);
            old = (int)(rr + ((grk & lowmask) << L));
        }
        if (tid == 0) idxs[j] = old;
    }
}

// ------------------------------------------------------------------------------------------
template <int T, int P, bool WEIGHTED, bool SMEM>
static int launch_fps(const FpsArgs& a, int b, int csize, cudaStream_t stream, int* max_clusters) {
    // the chain bookkeeping is compiled only into the 128-thread d-fps variants the planner picks for stacked layers
    constexpr bool kChainable = !WEIGHTED && !SMEM && T == 128;
    const bool chain = kChainable && (a.tie_iter != nullptr || a.parent_tie != nullptr);
    auto kern = chain ? fps_kernel<T, P, WEIGHTED, SMEM, kChainable> : fps_kernel<T, P, WEIGHTED, SMEM, false>;
    const size_t dyn = (size_t)3 * P * T * sizeof(float) + (size_t)2 * csize * (T / 32) * sizeof(FpsRec);
    if (dyn > 227 * 1024 - 64) return TSM_ERR_INVALID;
    if (dyn > 40 * 1024) TSM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    if (csize > 8) TSM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * csize));
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters) {  // query only: how many clusters of this shape are co-resident
        cudaError_t e = cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg);
        if (e != cudaSuccess) {
            cudaGetLastError();
            *max_clusters = 0;
        }
        return TSM_OK;
    }
    TSM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    return TSM_OK;
}

template <bool WEIGHTED>
static int dispatch_fps(const FpsArgs& a, int b, int csize, int T, int P, bool smem, cudaStream_t s,
                        int* max_clusters = nullptr) {
#define FPS_CASE(TT, PP) \
    if (T == TT && P == PP && !smem) return launch_fps<TT, PP, WEIGHTED, false>(a, b, csize, s, max_clusters);
#define FPS_CASE_SMEM(TT, PP) \
    if (T == TT && P == PP && smem) return launch_fps<TT, PP, WEIGHTED, true>(a, b, csize, s, max_clusters);
    FPS_CASE(128, 1) FPS_CASE(128, 2) FPS_CASE(128, 4) FPS_CASE(128, 8) FPS_CASE(128, 16) FPS_CASE(128, 32)
    FPS_CASE(256, 1) FPS_CASE(256, 2) FPS_CASE(256, 4) FPS_CASE(256, 8) FPS_CASE(256, 16) FPS_CASE(256, 32)
    FPS_CASE(512, 1) FPS_CASE(512, 2) FPS_CASE(512, 4) FPS_CASE(512, 8) FPS_CASE(512, 16)
    FPS_CASE(1024, 1) FPS_CASE(1024, 2) FPS_CASE(1024, 4) FPS_CASE(1024, 8)
    FPS_CASE_SMEM(1024, 16) FPS_CASE_SMEM(512, 32)
#undef FPS_CASE
#undef FPS_CASE_SMEM
    return TSM_ERR_INVALID;
}

}  // namespace tsm

// cuda_utils.h:10-14, bit for bit (same libm double log on the host).
static int ref_log2_block(int n) {
    const int pow_2 = (int)(log((double)n) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    int l = 0;
    while ((1 << l) < t) ++l;
    return l;
}

static int g_fps_algo = 0;  // 0 auto, 1 cluster kernel, 2 bucketed single-CTA kernel

struct FpsPlan {
    int csize, T, P;
    bool smem;
};

// Pick cluster size / block size / points per thread.  Overridable for tuning with
// TSMDET_FPS_CLUSTER / TSMDET_FPS_THREADS (values outside the valid set are ignored).
// Measured on B200 (profiles/r01_fps_sweep.txt): per-iteration time is a latency chain, so the
// fewest warps that still hold the cloud in registers win -- 128-thread CTAs with up to 16 points
// per thread, and the widest cluster whose B instances are all co-resident (a second wave of
// clusters would double the time).
static bool plan_for(int c, int want_t, int bs, int J, FpsPlan* out) {
    static const int Ts[4] = {128, 256, 512, 1024};
    static const int Ps[6] = {1, 2, 4, 8, 16, 32};
    FpsPlan fallback = {0, 0, 0, false};
    for (int ti = 0; ti < 4; ++ti) {
        const int T = Ts[ti];
        if (want_t && T != want_t) continue;
        const long G = (long)c * T;
        if (G < bs || G / 32 > tsm::kMaxRecs) continue;
        const int need = tsm::divup(J, (int)(G / bs));
        for (int pi = 0; pi < 6; ++pi) {
            const int P = Ps[pi];
            if (P < need) continue;
            bool smem = false;
            if ((long)T * P > 8192) {
                smem = (T == 1024 && P == 16) || (T == 512 && P == 32);
                if (!smem) break;
            }
            const FpsPlan pl = {c, T, P, smem};
            if (P <= 16 && !smem) {  // smallest block that keeps <= 16 points per thread
                *out = pl;
                return true;
            }
            fallback = pl;  // keeps the largest feasible T
            break;
        }
    }
    if (fallback.T) {
        *out = fallback;
        return true;
    }
    return false;
}

static int co_resident_clusters(const FpsPlan& pl, bool weighted) {
    tsm::FpsArgs dummy = {};
    int n = 0;
    if (weighted)
        tsm::dispatch_fps<true>(dummy, 1, pl.csize, pl.T, pl.P, pl.smem, nullptr, &n);
    else
        tsm::dispatch_fps<false>(dummy, 1, pl.csize, pl.T, pl.P, pl.smem, nullptr, &n);
    return n;
}

// occupancy == nullptr: host-only planning (no CUDA calls; used by tsmdet_fps_plan without a GPU)
static bool plan_fps(int b, int n, int log2bs, bool weighted, bool query_occupancy, FpsPlan* out) {
    const int bs = 1 << log2bs;
    const int J = tsm::divup(n, bs);
    const int sms = tsm_num_sms();
    int cmax = 1;
    while (cmax * 2 <= 8 && b * cmax * 2 <= sms) cmax *= 2;
    int want_c = 0, want_t = 0;
    if (const char* e = tsm_knob(KNOB_FPS_CLUSTER)) want_c = atoi(e);
    if (const char* e = tsm_knob(KNOB_FPS_THREADS)) want_t = atoi(e);
    if (want_t != 128 && want_t != 256 && want_t != 512 && want_t != 1024) want_t = 0;
    const bool forced = (want_c == 1 || want_c == 2 || want_c == 4 || want_c == 8 || want_c == 16);
    int order[8], no = 0;
    if (forced) order[no++] = want_c;
    for (int c = cmax; c >= 1; c >>= 1) order[no++] = c;
    for (int c = cmax * 2; c <= 16; c <<= 1) order[no++] = c;  // a portable cluster cannot hold the cloud
    FpsPlan first = {0, 0, 0, false};
    for (int pass = 0; pass < 2; ++pass) {
        const int wt = pass == 0 ? want_t : 0;
        for (int oi = 0; oi < no && oi < 8; ++oi) {
            FpsPlan pl;
            if (!plan_for(order[oi], wt, bs, J, &pl)) continue;
            if (!first.T) first = pl;
            if (!query_occupancy || (forced && oi == 0)) {
                *out = pl;
                return true;
            }
            // single wave: all b clusters resident at once (or nothing smaller can do better)
            const int fit = co_resident_clusters(pl, weighted);
            if (fit >= b || (fit > 0 && pl.csize == 1)) {
                *out = pl;
                return true;
            }
        }
        if (!want_t) break;
    }
    if (first.T) {
        *out = first;
        return true;
    }
    return false;
}

static int run_fps(int b, int n, int m, const float* xyz, const float* weights, float* temp, int* idxs,
                   cudaStream_t stream, int* tie_iter = nullptr, float* vals = nullptr, const int* parent_tie = nullptr,
                   const float* parent_vals = nullptr, int parent_m = 0) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (n <= 0) return TSM_ERR_INVALID;
    tsm::FpsArgs a;
    a.xyz = xyz;
    a.weights = weights;
    a.temp = temp;
    a.idxs = idxs;
    a.n = n;
    a.m = m;
    a.log2bs = ref_log2_block(n);
    a.status = tsm_status_word(stream);
    a.tie_iter = tie_iter;
    a.vals = vals;
    a.parent_tie = parent_tie;
    a.parent_vals = parent_vals;
    a.parent_m = parent_m;
    // Two kernels.  The cluster kernel (this file) has the shorter pick latency when every cloud can have a
    // full 8-CTA cluster to itself; the spatially pruned single-CTA kernel (fps_bucket.cu) needs one SM per
    // cloud and wins whenever clouds outnumber clusters -- large batches, or several batches in flight.
    // auto: by batch size; tsmdet_fps_configure() / TSMDET_FPS_ALGO = cluster | bucket override, and an explicit
    // cluster launch shape (tuning / tests) implies "cluster".
    {
        int algo = g_fps_algo;
        if (const char* e = tsm_knob(KNOB_FPS_ALGO)) algo = !strcmp(e, "cluster") ? 1 : (!strcmp(e, "bucket") ? 2 : algo);
        if (tsm_knob(KNOB_FPS_CLUSTER) || tsm_knob(KNOB_FPS_THREADS)) algo = 1;
        const bool fits = tsm_fps_bucket_supports(n, weights != nullptr);
        const bool crowded = (long)b * 8 > tsm_num_sms();
        if (fits && (algo == 2 || (algo == 0 && crowded && n >= 1024))) return tsm_fps_bucket_launch(a, b, stream);
        // clouds of 16385..65536 points: the pruned sampler over a cluster of CTAs (one SM per ~15000 points; a pick
        // costs the same whatever N is, where the brute-force cluster kernel below slows down with N)
        if (algo != 1 && tsm_fps_bucket_cluster_supports(n, weights != nullptr)) {
            if (tie_iter) TSM_CUDA_TRY(cudaMemsetAsync(tie_iter, 0, sizeof(int) * (size_t)b, stream));  // no chaining facts
            a.tie_iter = nullptr;
            a.vals = nullptr;
            a.parent_tie = nullptr;
            a.parent_vals = nullptr;
            return tsm_fps_bucket_cluster_launch(a, b, stream);
        }
    }
    FpsPlan pl;
    if (!plan_fps(b, n, a.log2bs, weights != nullptr, true, &pl)) return TSM_ERR_INVALID;
    if (weights != nullptr || pl.smem || pl.T != 128) {
        // launch shapes without the chain bookkeeping: tell any follow-up level "tie at iteration 0"
        if (tie_iter) TSM_CUDA_TRY(cudaMemsetAsync(tie_iter, 0, sizeof(int) * (size_t)b, stream));
        a.tie_iter = nullptr;
        a.vals = nullptr;
        a.parent_tie = nullptr;
        a.parent_vals = nullptr;
    }
    if (weights) return tsm::dispatch_fps<true>(a, b, pl.csize, pl.T, pl.P, pl.smem, stream);
    return tsm::dispatch_fps<false>(a, b, pl.csize, pl.T, pl.P, pl.smem, stream);
}

extern "C" {

int tsmdet_fps_plan(int b, int n, int* csize, int* threads, int* pts_per_thread, int* smem_xyz) {
    FpsPlan pl;
    int ndev = 0;
    const bool gpu = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
    if (!gpu) cudaGetLastError();
    if (n <= 0 || !plan_fps(b, n, ref_log2_block(n), false, gpu, &pl)) return TSM_ERR_INVALID;
    if (csize) *csize = pl.csize;
    if (threads) *threads = pl.T;
    if (pts_per_thread) *pts_per_thread = pl.P;
    if (smem_xyz) *smem_xyz = pl.smem ? 1 : 0;
    return TSM_OK;
}

// Process-wide choice of the d-FPS kernel: 0 = auto (by batch size), 1 = cluster kernel, 2 = bucketed
// single-CTA kernel (clouds of <= 16384 points; others keep the cluster kernel).  Results are identical.
int tsmdet_fps_configure(int algo) {
    if (algo < 0 || algo > 2) return TSM_ERR_INVALID;
    g_fps_algo = algo;
    return TSM_OK;
}

int tsmdet_farthest_point_sampling(int b, int n, int m, const float* xyz, float* temp, int* idxs, void* stream) {
    return run_fps(b, n, m, xyz, nullptr, temp, idxs, (cudaStream_t)stream);
}

// FPS that records what a follow-up FPS over its own output needs (tie_iter (B) i32, vals (B,M) f32), and
// that can itself be the follow-up of a recorded run: parent_tie (B) / parent_vals (B,parent_m) describe the
// FPS whose first n picks, in order, ARE this xyz.  Results are identical to tsmdet_farthest_point_sampling
// (the prediction is only used where the parent run proves it exact); temp is not produced on that path.
int tsmdet_fps_chain(int b, int n, int m, const float* xyz, float* temp, int* idxs, int* tie_iter, float* vals,
                     const int* parent_tie, const float* parent_vals, int parent_m, void* stream) {
    if ((parent_tie == nullptr) != (parent_vals == nullptr)) return TSM_ERR_INVALID;
    return run_fps(b, n, m, xyz, nullptr, temp, idxs, (cudaStream_t)stream, tie_iter, vals, parent_tie, parent_vals,
                   parent_m);
}

int tsmdet_furthest_point_sampling_weights(int b, int n, int m, const float* xyz, const float* weights, float* temp,
                                           int* idxs, void* stream) {
    if (!weights) return TSM_ERR_INVALID;
    return run_fps(b, n, m, xyz, weights, temp, idxs, (cudaStream_t)stream);
}

int tsmdet_furthest_point_sampling_matrix(int b, int n, int m, const float* matrix, float* temp, int* idxs,
                                          void* stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (n <= 0 || !temp) return TSM_ERR_INVALID;
    tsm::fps_matrix_kernel<false><<<b, 1024, 0, (cudaStream_t)stream>>>(matrix, nullptr, temp, idxs, n, m,
                                                                        ref_log2_block(n));
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

int tsmdet_furthest_point_sampling_with_weighted_dist(int b, int n, int m, const float* matrix, const float* weights,
                                                      float* temp, int* idxs, void* stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (n <= 0 || !temp || !weights) return TSM_ERR_INVALID;
    tsm::fps_matrix_kernel<true><<<b, 1024, 0, (cudaStream_t)stream>>>(matrix, weights, temp, idxs, n, m,
                                                                       ref_log2_block(n));
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

}  // extern "C"
