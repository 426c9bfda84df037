// fps_bucket_cluster.cu -- the spatially pruned sampler of fps_bucket.cu for clouds that need several SMs
// (16384 < N <= 240000): one thread-block CLUSTER per cloud, every CTA holding a contiguous slice of the
// Morton-sorted cloud in shared memory.
//
// Same indices, bit for bit, as farthest_point_sampling_kernel
//   /root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:100-216.
//
//  * fps_sort_kernel (one CTA per cloud, any N): bounding box -> adaptive Morton grid (15 key bits, handed out
//    greedily to the axis with the largest cell extent) -> counting sort (histogram in shared memory, scan, scatter)
//    of (x, y, z, original index) float4 records into a global scratch array.
//  * fps_bucket_cluster_kernel: CTA r of the cluster loads slice r of that array into the shared-memory planes of
//    fps_bucket.cu (lanes own P spatially compact points, min-distances in registers, one box per 16 lanes) and
//    runs the same per-pick step -- exact box pruning, one CTA barrier, local 32-record reduction.  The CTAs'
//    winners (key, original index, coordinates: 20 bytes) are then exchanged with st.async into every peer's
//    shared memory, completing on the peer's mbarrier (double-buffered by pick parity, no cluster-wide barrier on
//    the critical path), and every warp picks the cluster winner: largest key, smallest reference rank among
//    equals.  A pick costs what a single-CTA pick costs plus one DSMEM flight, whatever N is, instead of growing
//    with N as the brute-force cluster kernel of fps.cu does.  (TSMDET_FPSC_K=1; kept as the A/B reference.)
//  * fps_bucket_cluster_mp_kernel (the default): the same slices, but ROUNDS of several exact picks per exchange --
//    every CTA publishes its top-KX warp candidates and the bound of the ones it did not publish, and every CTA
//    makes the same merged decision (see the comment above the kernel).  About half the time per pick.
#include "fps.cuh"

namespace tsm {

constexpr int FBC_T = 1024;
constexpr int FBC_P = 16;
constexpr int FBC_CAP = FBC_T * FBC_P;  // points per CTA
constexpr int FBC_CELL_BITS = 15;
constexpr uint32_t kXRecBytes = 20;  // key, index, x, y (v4) + z (b32)

struct __align__(8) XRec {  // 24-byte slots: 16 peers x 2 parities have to fit next to 224 KB of points
    uint32_t u, k;
    float x, y;
    float z;
    uint32_t pad;
};
static_assert(sizeof(XRec) == 24, "XRec must be 24 bytes");

__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "r"(a), "r"(b), "r"(rbar)
                 : "memory");
}

__device__ __forceinline__ uint32_t fbc_rank(uint32_t k, int L) {
    return (L == 0) ? k : (__brev(k & ((1u << L) - 1u)) | (k >> L));
}

__device__ __forceinline__ int fbc_slot_index(int s) {  // sorted position within the CTA -> shared-memory element
    constexpr int P = FBC_P;
    const int p = s & (P - 1), l = (s / P) & 31, w = s / (32 * P);
    return w * (32 * P) + (p >> 2) * 128 + l * 4 + (p & 3);
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
    fps_sort_kernel(int n, const float* __restrict__ xyz_all, float4* __restrict__ sorted_all) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t* const hist = reinterpret_cast<uint32_t*>(dyn);  // [1 << FBC_CELL_BITS]
    __shared__ float red[6][32];
    __shared__ uint32_t woff[32];
    constexpr int T = 1024, NW = 32, ncells = 1 << FBC_CELL_BITS, cell_bits = FBC_CELL_BITS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* __restrict__ xyz = xyz_all + (size_t)cloud * n * 3;
    float4* __restrict__ sorted = sorted_all + (size_t)cloud * n;

    const float inf = __int_as_float(0x7f800000);
    float lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
    for (int k = tid; k < n; k += T) {
        const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
        lo0 = fminf(lo0, x); hi0 = fmaxf(hi0, x);
        lo1 = fminf(lo1, y); hi1 = fmaxf(hi1, y);
        lo2 = fminf(lo2, z); hi2 = fmaxf(hi2, z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo0 = fminf(lo0, __shfl_xor_sync(FULL, lo0, o)); hi0 = fmaxf(hi0, __shfl_xor_sync(FULL, hi0, o));
        lo1 = fminf(lo1, __shfl_xor_sync(FULL, lo1, o)); hi1 = fmaxf(hi1, __shfl_xor_sync(FULL, hi1, o));
        lo2 = fminf(lo2, __shfl_xor_sync(FULL, lo2, o)); hi2 = fmaxf(hi2, __shfl_xor_sync(FULL, hi2, o));
    }
    if (lane == 0) {
        red[0][warp] = lo0; red[1][warp] = lo1; red[2][warp] = lo2;
        red[3][warp] = hi0; red[4][warp] = hi1; red[5][warp] = hi2;
    }
    for (int c = tid; c < ncells; c += T) hist[c] = 0u;
    __syncthreads();
    for (int w = 0; w < NW; ++w) {
        lo0 = fminf(lo0, red[0][w]); lo1 = fminf(lo1, red[1][w]); lo2 = fminf(lo2, red[2][w]);
        hi0 = fmaxf(hi0, red[3][w]); hi1 = fmaxf(hi1, red[4][w]); hi2 = fmaxf(hi2, red[5][w]);
    }
    float e0 = hi0 - lo0, e1 = hi1 - lo1, e2 = hi2 - lo2;
    if (!(e0 > 0.f) || !(e0 < 3.0e38f)) e0 = 0.f;
    if (!(e1 > 0.f) || !(e1 < 3.0e38f)) e1 = 0.f;
    if (!(e2 > 0.f) || !(e2 < 3.0e38f)) e2 = 0.f;
    int nb0 = 0, nb1 = 0, nb2 = 0;
    uint32_t order = 0u;  // 2 bits per key bit, most significant key bit first
    {
        float c0 = e0, c1 = e1, c2 = e2;
        for (int i = 0; i < cell_bits; ++i) {
            int ax = 0;
            if (c1 > c0 && c1 >= c2) ax = 1;
            if (c2 > c0 && c2 > c1) ax = 2;
            order |= (uint32_t)ax << (2 * i);
            if (ax == 0) { ++nb0; c0 *= 0.5f; }
            else if (ax == 1) { ++nb1; c1 *= 0.5f; }
            else { ++nb2; c2 *= 0.5f; }
        }
    }
    const float inv0 = e0 > 0.f ? (float)(1 << nb0) / e0 : 0.f;
    const float inv1 = e1 > 0.f ? (float)(1 << nb1) / e1 : 0.f;
    const float inv2 = e2 > 0.f ? (float)(1 << nb2) / e2 : 0.f;
    auto cell_key = [&](float x, float y, float z) -> uint32_t {
        const int q0 = min(max(__float2int_rd((x - lo0) * inv0), 0), (1 << nb0) - 1);
        const int q1 = min(max(__float2int_rd((y - lo1) * inv1), 0), (1 << nb1) - 1);
        const int q2 = min(max(__float2int_rd((z - lo2) * inv2), 0), (1 << nb2) - 1);
        int r0 = nb0, r1 = nb1, r2 = nb2;
        uint32_t key = 0u;
#pragma unroll 1
        for (int i = 0; i < cell_bits; ++i) {
            const uint32_t ax = (order >> (2 * i)) & 3u;
            uint32_t bit;
            if (ax == 0u) bit = (q0 >> --r0) & 1;
            else if (ax == 1u) bit = (q1 >> --r1) & 1;
            else bit = (q2 >> --r2) & 1;
            key = (key << 1) | bit;
        }
        return key;
    };
    for (int k = tid; k < n; k += T)
        atomicAdd(&hist[cell_key(__ldg(xyz + 3 * k), __ldg(xyz + 3 * k + 1), __ldg(xyz + 3 * k + 2))], 1u);
    __syncthreads();
    {
        constexpr int cpw = ncells / NW;
        uint32_t carry = 0u;
        for (int c = warp * cpw + lane; c < (warp + 1) * cpw; c += 32) {
            const uint32_t v = hist[c];
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            hist[c] = carry + inc - v;
            carry += __shfl_sync(FULL, inc, 31);
        }
        if (lane == 0) woff[warp] = carry;
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = woff[lane];
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            woff[lane] = inc - v;
        }
        __syncthreads();
        const uint32_t off = woff[warp];
        for (int c = warp * cpw + lane; c < (warp + 1) * cpw; c += 32) hist[c] += off;
    }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
        const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
        const uint32_t pos = atomicAdd(&hist[cell_key(x, y, z)], 1u);
        sorted[pos] = make_float4(x, y, z, __int_as_float(k));  // order inside a cell is arbitrary: it never decides a result
    }
}

// ------------------------------------------------------------------------------------------------------------
// dynamic smem: planes 12*CAP | original index u16 2*CAP | group boxes 24*(T/16) | recs [2][32] uint2 |
//               exchange records [2][16] XRec (24 B) | first point (16 B)
constexpr int FBC_MAXC = 16;
constexpr int FBC_SMEM = 14 * FBC_CAP + 24 * (FBC_T / 16) + 512 + 2 * FBC_MAXC * 24 + 16;

// BIG (N > 65536): the u16 next to every point is its position in the CTA's slice; the original index is read from
// the sorted array when it is needed (lane ordering and temp in the prologue, shared maxima, the final translation).
template <bool BIG>
__global__ void __launch_bounds__(FBC_T, 1)
    fps_bucket_cluster_kernel(const FpsArgs a, const float4* __restrict__ sorted_all, int per_cta) {
    constexpr int T = FBC_T, P = FBC_P, CAP = FBC_CAP, C4 = P / 4, NW = T / 32;
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ __align__(8) uint64_t mbar[2];
    float* const sx = reinterpret_cast<float*>(dyn);
    float* const sy = sx + CAP;
    float* const sz = sy + CAP;
    uint16_t* const sk = reinterpret_cast<uint16_t*>(sz + CAP);
    float4* const sbox4 = reinterpret_cast<float4*>(sk + CAP);
    float2* const sbox2 = reinterpret_cast<float2*>(sbox4 + T / 16);
    uint2* const recs = reinterpret_cast<uint2*>(sbox2 + T / 16);  // [2][32]
    XRec* const xrec = reinterpret_cast<XRec*>(recs + 64);         // [2][FBC_MAXC]
    float* const first_xyz = reinterpret_cast<float*>(xrec + 2 * FBC_MAXC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = cluster_ctarank();
    const uint32_t csize = cluster_nctarank();
    const int cloud = blockIdx.x / csize;
    const int n = a.n, m = a.m, L = a.log2bs;
    const float* __restrict__ xyz = a.xyz + (size_t)cloud * n * 3;
    const float4* __restrict__ sorted = sorted_all + (size_t)cloud * n;
    int* __restrict__ idxs = a.idxs + (size_t)cloud * m;
    const int lo = min((int)crank * per_cta, n);
    const int nloc = min(per_cta, n - lo);  // this CTA's slice of the sorted cloud: [lo, lo + nloc)
    auto orig = [&](int e) -> uint32_t {  // original index of the point at shared-memory element e
        const uint32_t v = sk[e];
        return BIG ? (uint32_t)__float_as_int(sorted[lo + (int)v].w) : v;
    };

    if (tid == 0) {
        mbar_init(smem_u32(&mbar[0]), 1);
        mbar_init(smem_u32(&mbar[1]), 1);
        mbar_fence_init_cluster();
        first_xyz[0] = __ldg(xyz + 0);
        first_xyz[1] = __ldg(xyz + 1);
        first_xyz[2] = __ldg(xyz + 2);
        if (crank == 0 && m > 0) idxs[0] = 0;
    }
    // ---- load the slice into the lane-major planes
    for (int s = tid; s < CAP; s += T) {
        const int e = fbc_slot_index(s);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < nloc) v = sorted[lo + s];
        sx[e] = v.x; sy[e] = v.y; sz[e] = v.z;
        sk[e] = s < nloc ? (BIG ? (uint16_t)s : (uint16_t)__float_as_int(v.w)) : (uint16_t)0;
    }
    __syncthreads();

    const int base = warp * (32 * P) + lane * 4;
    const int first = (warp * 32 + lane) * P;
    float md[P];
    {
        // each lane's P points in ascending reference rank (see fps_bucket.cu)
        uint32_t rk[P];
#pragma unroll
        for (int p = 0; p < P; ++p)
            rk[p] = (first + p < nloc) ? fbc_rank(orig(base + (p >> 2) * 128 + (p & 3)), L) : (0xffffffe0u + (uint32_t)p);
        int dst[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int before = 0;
#pragma unroll
            for (int q = 0; q < P; ++q) before += (rk[q] < rk[p]) ? 1 : 0;
            dst[p] = base + (before >> 2) * 128 + (before & 3);
        }
        {
            uint16_t v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = sk[base + (p >> 2) * 128 + (p & 3)];
#pragma unroll
            for (int p = 0; p < P; ++p) sk[dst[p]] = v[p];
        }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            float* const plane = sx + pl * CAP;
            float v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = plane[base + (p >> 2) * 128 + (p & 3)];
#pragma unroll
            for (int p = 0; p < P; ++p) plane[dst[p]] = v[p];
        }
    }
    {
        float bx0 = __int_as_float(0x7f800000), by0 = bx0, bz0 = bx0, bx1 = -bx0, by1 = -bx0, bz1 = -bx0;
#pragma unroll
        for (int c = 0; c < C4; ++c) {
            const float4 X = *reinterpret_cast<const float4*>(sx + base + c * 128);
            const float4 Y = *reinterpret_cast<const float4*>(sy + base + c * 128);
            const float4 Z = *reinterpret_cast<const float4*>(sz + base + c * 128);
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = c * 4 + e;
                float d0 = -2.f;
                if (first + p < nloc) {
                    bx0 = fminf(bx0, xs[e]); bx1 = fmaxf(bx1, xs[e]);
                    by0 = fminf(by0, ys[e]); by1 = fmaxf(by1, ys[e]);
                    bz0 = fminf(bz0, zs[e]); bz1 = fmaxf(bz1, zs[e]);
                    d0 = a.temp ? a.temp[(size_t)cloud * n + orig(base + c * 128 + e)] : 1e10f;
                }
                md[p] = d0;
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(FULL, bx1, o));
            by0 = fminf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(FULL, by1, o));
            bz0 = fminf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fmaxf(bz1, __shfl_xor_sync(FULL, bz1, o));
        }
        if ((lane & 15) == 0) {
            sbox4[tid >> 4] = make_float4(bx0, bx1, by0, by1);
            sbox2[tid >> 4] = make_float2(bz0, bz1);
        }
    }
    __syncthreads();

    float lmax = -2.f;
    uint32_t u = 0u, wu = 0u, wpos = (uint32_t)base;
    int lpos = base;

    auto warp_step = [&](float x1, float y1, float z1, const float4 b4, const float2 b2, bool force) {
        const float dx = fmaxf(fmaxf(__fsub_rn(b4.x, x1), __fsub_rn(x1, b4.y)), 0.f);
        const float dy = fmaxf(fmaxf(__fsub_rn(b4.z, y1), __fsub_rn(y1, b4.w)), 0.f);
        const float dz = fmaxf(fmaxf(__fsub_rn(b2.x, z1), __fsub_rn(z1, b2.y)), 0.f);
        const float lb = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const bool hit = !(lb >= lmax) || force;  // NaN bounds count as hits
        if (!__any_sync(FULL, hit)) return;
#pragma unroll
        for (int c = 0; c < C4; ++c) {
            const float4 X = *reinterpret_cast<const float4*>(sx + base + c * 128);
            const float4 Y = *reinterpret_cast<const float4*>(sy + base + c * 128);
            const float4 Z = *reinterpret_cast<const float4*>(sz + base + c * 128);
            md[c * 4 + 0] = fminf(sqdist3(x1, y1, z1, X.x, Y.x, Z.x), md[c * 4 + 0]);
            md[c * 4 + 1] = fminf(sqdist3(x1, y1, z1, X.y, Y.y, Z.y), md[c * 4 + 1]);
            md[c * 4 + 2] = fminf(sqdist3(x1, y1, z1, X.z, Y.z, Z.z), md[c * 4 + 2]);
            md[c * 4 + 3] = fminf(sqdist3(x1, y1, z1, X.w, Y.w, Z.w), md[c * 4 + 3]);
        }
        float t[P];
#pragma unroll
        for (int p = 0; p < P; ++p) t[p] = md[p];
#pragma unroll
        for (int w = 1; w < P; w <<= 1) {
#pragma unroll
            for (int p = 0; p + w < P; p += 2 * w) t[p] = fmaxf(t[p], t[p + w]);
        }
        lmax = t[0];
        u = (lmax > -1.f) ? f32_ordered(lmax) : 0u;  // the reference starts from best = -1
        uint32_t eq = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) eq |= (md[p] == lmax) ? (1u << p) : 0u;
        const int bp = __ffs(eq) - 1;  // lowest slot among equals = smallest reference rank inside the lane
        lpos = base + (bp >> 2) * 128 + (bp & 3);
        wu = __reduce_max_sync(FULL, u);
        const unsigned tie = __ballot_sync(FULL, u == wu);
        wpos = __shfl_sync(FULL, (uint32_t)lpos, __ffs(tie) - 1);
        if (wu != 0u && (tie & (tie - 1u)) != 0u) {  // shared maximum inside the warp: the reference rank decides
            const uint32_t rk = (u == wu) ? fbc_rank(orig(lpos), L) : 0xffffffffu;
            const uint32_t wrk = __reduce_min_sync(FULL, rk);
            wpos = (uint32_t)__shfl_sync(FULL, lpos, __ffs(__ballot_sync(FULL, rk == wrk)) - 1);
        }
    };

    // peers' record slots / barriers (lanes < csize of warp 0 send)
    const uint32_t peer = (uint32_t)lane < csize ? (uint32_t)lane : 0u;
    const uint32_t dst_rec0 = mapa_u32(smem_u32(&xrec[crank]), peer);
    const uint32_t dst_rec1 = mapa_u32(smem_u32(&xrec[FBC_MAXC + crank]), peer);
    if (BIG && crank == 0)  // picks are written as (CTA << 16 | slice position) and translated after the loop
        for (int j = 1 + tid; j < m; j += T) idxs[j] = -1;
    const uint32_t dst_bar0 = mapa_u32(smem_u32(&mbar[0]), peer);
    const uint32_t dst_bar1 = mapa_u32(smem_u32(&mbar[1]), peer);
    cluster_sync_all();  // peers' mbarriers are initialised past this point

    float x1 = first_xyz[0], y1 = first_xyz[1], z1 = first_xyz[2];
    float4 b4 = sbox4[tid >> 4];
    float2 b2 = sbox2[tid >> 4];
    int it = 0;
    for (int j = 1; j < m; ++j, ++it) {
        const int par = it & 1;
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&mbar[par]), csize * kXRecBytes);
        uint2* const rec = recs + par * 32;
        warp_step(x1, y1, z1, b4, b2, j == 1);
        if (lane == 0) rec[warp] = make_uint2(wu, wpos);
        __syncthreads();
        const uint2 r = rec[lane];
        b4 = sbox4[tid >> 4];
        b2 = sbox2[tid >> 4];
        // ---- this CTA's candidate
        const uint32_t lu = __reduce_max_sync(FULL, r.x);
        const unsigned gt = __ballot_sync(FULL, r.x == lu);
        uint32_t cpos = __shfl_sync(FULL, r.y, __ffs(gt) - 1);
        if (lu != 0u && (gt & (gt - 1u)) != 0u) {
            const uint32_t rk = (r.x == lu) ? fbc_rank(orig((int)r.y), L) : 0xffffffffu;
            const uint32_t grk = __reduce_min_sync(FULL, rk);
            cpos = __shfl_sync(FULL, r.y, __ffs(__ballot_sync(FULL, rk == grk)) - 1);
        }
        // ---- exchange: warp 0 posts (key, original index, coordinates) to every CTA of the cluster
        if (warp == 0 && (uint32_t)lane < csize) {
            const uint32_t kk = lu != 0u ? (BIG ? ((crank << 16) | (uint32_t)sk[cpos]) : (uint32_t)sk[cpos]) : 0u;
            const uint32_t dr = par ? dst_rec1 : dst_rec0, db = par ? dst_bar1 : dst_bar0;
            st_async_v2(dr, db, lu, kk);
            st_async_v2(dr + 8, db, __float_as_uint(sx[cpos]), __float_as_uint(sy[cpos]));
            st_async_b32(dr + 16, db, __float_as_uint(sz[cpos]));
        }
        {
            const uint32_t bar = smem_u32(&mbar[par]), ph = (uint32_t)((it >> 1) & 1);
            if (!mbar_try_wait_cluster(bar, ph)) {
                const long long t0 = clock64();
                while (!mbar_try_wait_cluster(bar, ph))
                    if (clock64() - t0 > 4000000000LL) watchdog_trip(a.status, TSM_ERR_WATCHDOG);
            }
        }
        uint32_t cu = 0u, ck = 0xffffffffu;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if ((uint32_t)lane < csize) {
            const XRec* xr = xrec + par * FBC_MAXC + lane;
            const uint2 v0 = *reinterpret_cast<const uint2*>(xr);
            const float2 v1 = *reinterpret_cast<const float2*>(&xr->x);
            cu = v0.x;
            ck = v0.y;
            cx = v1.x;
            cy = v1.y;
            cz = xr->z;
        }
        const uint32_t gu = __reduce_max_sync(FULL, cu);
        const unsigned ct = __ballot_sync(FULL, cu == gu && (uint32_t)lane < csize);
        int gl = __ffs(ct) - 1;
        if (gu != 0u && (ct & (ct - 1u)) != 0u) {
            uint32_t rk = 0xffffffffu;
            if ((ct >> lane) & 1u) {
                const uint32_t kt = BIG ? (uint32_t)__float_as_int(sorted[(int)(ck >> 16) * per_cta + (int)(ck & 0xffffu)].w) : ck;
                rk = fbc_rank(kt, L);
            }
            const uint32_t grk = __reduce_min_sync(FULL, rk);
            gl = __ffs(__ballot_sync(FULL, rk == grk)) - 1;
        }
        if (gu != 0u) {
            x1 = __shfl_sync(FULL, cx, gl);
            y1 = __shfl_sync(FULL, cy, gl);
            z1 = __shfl_sync(FULL, cz, gl);
        } else {  // no eligible candidate anywhere: the reference yields index 0
            x1 = first_xyz[0];
            y1 = first_xyz[1];
            z1 = first_xyz[2];
        }
        const uint32_t gk = __shfl_sync(FULL, ck, gl);
        if (crank == 0 && tid == 0) idxs[j] = gu != 0u ? (int)gk : (BIG ? -1 : 0);
    }
    if (BIG && crank == 0) {
        __syncthreads();  // this CTA wrote every idxs[j]
        for (int j = 1 + tid; j < m; j += T) {
            const int code = idxs[j];
            idxs[j] = code >= 0 ? __float_as_int(sorted[(code >> 16) * per_cta + (code & 0xffff)].w) : 0;
        }
    }
    if (a.temp) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (first + p < nloc) a.temp[(size_t)cloud * n + orig(base + (p >> 2) * 128 + (p & 3))] = md[p];
    }
    cluster_sync_all();  // no CTA leaves while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------------------------
// Rounds of several exact picks per DSMEM exchange (the multi-pick rule of fps_bucket.cu, DESIGN.md 4.1, over a
// cluster).  One pick per exchange costs ~1.03 us whatever N is: a local 32-record reduction, a DSMEM flight, a
// csize-record reduction -- all on the dependent chain of every pick.  Here a ROUND is
//   APPLY     every warp tests the round's picks against its lanes' group boxes (one redux over the per-lane hit
//             masks), applies the ones that may reach it and re-runs its argmax only if its candidate itself was
//             lowered; its record (key, position, runner-up key = bound of its other points, sort key) stays posted;
//   LOCAL     the CTA's leader warp sorts the 32 warp records (one compare per record against 8 LDS.128) and takes the
//             top KX: list[0] is the EXACT local argmax (largest key, smallest reference rank among equals -- on a
//             near tie at the top the list is cut to that one record), `ubound` = the largest key it did not list;
//   EXCHANGE  KX records (key, index, x, y, z, runner-up) + ubound go to every CTA with st.async, completing on the
//             destination's mbarrier (double-buffered by round parity);
//   GLOBAL    every CTA's leader warp sorts the csize * KX (<= 32) records the same way.  Position 0 is the plain
//             argmax (largest key, smallest reference rank among the records of its 32-ulp bucket).  The record q at
//             sorted position k >= 1 is ALSO the pick the one-at-a-time algorithm would make next when
//               (i)   no other listed record shares q's 32-ulp bucket, and q's bucket is above every CTA's ubound
//                     (every unlisted candidate is strictly below q),
//               (ii)  q's key is strictly above the runner-up key of every warp an earlier pick of the round came from,
//               (iii) no earlier pick of the round lowers it: !(sqdist(pick, q) < md[q]), the update's own expression;
//             every other point is below q already and updates only lower min-distances.  The round ends at the first
//             position that fails; all CTAs compute the same decision from the same records.
// Picks, and therefore every index, are bit-identical to fps_bucket_cluster_kernel's (tests: test_fps_gpu.py).
// The round state lives in the tail of the x plane: a slice holds at most 15360 points (launcher), so the last two
// warps' 1024 slots are never used by points.
constexpr int FBM_KA = 8;  // picks per round at most

struct __align__(16) MpRec {  // one exchanged candidate, 32 bytes
    uint32_t key, kk;  // ordered min-distance (0 = none), original index (BIG: CTA << 16 | slice position)
    float x, y;
    float z;
    uint32_t r2, ubound, pad;  // runner-up key of its warp; the sender's largest unlisted key (same in all its records)
};
static_assert(sizeof(MpRec) == 32, "MpRec must be 32 bytes");

template <bool BIG>
__global__ void __launch_bounds__(FBC_T, 1)
    fps_bucket_cluster_mp_kernel(const FpsArgs a, const float4* __restrict__ sorted_all, int per_cta, int kx) {
    constexpr int T = FBC_T, P = FBC_P, CAP = FBC_CAP, C4 = P / 4, NW = T / 32, KA = FBM_KA;
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ __align__(8) uint64_t mbar[2];
    float* const sx = reinterpret_cast<float*>(dyn);
    float* const sy = sx + CAP;
    float* const sz = sy + CAP;
    uint16_t* const sk = reinterpret_cast<uint16_t*>(sz + CAP);
    float4* const sbox4 = reinterpret_cast<float4*>(sk + CAP);
    float2* const sbox2 = reinterpret_cast<float2*>(sbox4 + T / 16);
    float* const first_xyz = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sbox2 + T / 16) + 512 + 2 * FBC_MAXC * 24);
    // round state: slots 15360.. of the x plane (4 KB)
    uint32_t* const rs = reinterpret_cast<uint32_t*>(sx + 15360);
    uint32_t* const rkey = rs;                                   // [32] warp records: key (bit 31 set when eligible)
    uint32_t* const rpos = rs + 32;                              // [32] shared-memory element of the candidate
    uint32_t* const rkey2 = rs + 64;                             // [32] runner-up key of the warp
    uint32_t* const rskey = rs + 96;                             // [32] sort key: (key & ~31) | (31 - warp)
    uint32_t* const inv = rs + 128;                              // [33] leader: record at a sorted position
    uint32_t* const gsk = rs + 168;                              // [34] leader: sort key at a global position
    float4* const cand = reinterpret_cast<float4*>(rs + 208);    // [KA] the round's picks: x, y, z
    uint32_t* const ckey = rs + 240;                             // [KA] their keys
    uint32_t* const cr2 = rs + 248;                              // [KA] the runner-up keys of their warps
    int* const npick = reinterpret_cast<int*>(rs + 256);
    MpRec* const xl = reinterpret_cast<MpRec*>(rs + 264);        // [2][32] exchanged lists, by round parity
    static_assert((264 + 2 * 32 * 8) * 4 <= 4096, "round state must fit the x-plane tail");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = cluster_ctarank();
    const uint32_t csize = cluster_nctarank();
    const int cloud = blockIdx.x / csize;
    const int n = a.n, m = a.m, L = a.log2bs;
    const float* __restrict__ xyz = a.xyz + (size_t)cloud * n * 3;
    const float4* __restrict__ sorted = sorted_all + (size_t)cloud * n;
    int* __restrict__ idxs = a.idxs + (size_t)cloud * m;
    const int lo = min((int)crank * per_cta, n);
    const int nloc = min(per_cta, n - lo);
    auto orig = [&](int e) -> uint32_t {
        const uint32_t v = sk[e];
        return BIG ? (uint32_t)__float_as_int(sorted[lo + (int)v].w) : v;
    };
    auto orig_of_kk = [&](uint32_t kk) -> uint32_t {
        return BIG ? (uint32_t)__float_as_int(sorted[(int)(kk >> 16) * per_cta + (int)(kk & 0xffffu)].w) : kk;
    };

    if (tid == 0) {
        mbar_init(smem_u32(&mbar[0]), 1);
        mbar_init(smem_u32(&mbar[1]), 1);
        mbar_fence_init_cluster();
        first_xyz[0] = __ldg(xyz + 0);
        first_xyz[1] = __ldg(xyz + 1);
        first_xyz[2] = __ldg(xyz + 2);
        if (crank == 0 && m > 0) idxs[0] = 0;
    }
    for (int s = tid; s < CAP; s += T) {
        if (s >= 15360) continue;  // round state
        const int e = fbc_slot_index(s);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < nloc) v = sorted[lo + s];
        sx[e] = v.x; sy[e] = v.y; sz[e] = v.z;
        sk[e] = s < nloc ? (BIG ? (uint16_t)s : (uint16_t)__float_as_int(v.w)) : (uint16_t)0;
    }
    __syncthreads();

    const int base = warp * (32 * P) + lane * 4;
    const int first = (warp * 32 + lane) * P;
    const bool has_pts = warp < 30;  // slots 15360.. hold the round state
    float md[P];
    if (has_pts) {
        uint32_t rk[P];
#pragma unroll
        for (int p = 0; p < P; ++p)
            rk[p] = (first + p < nloc) ? fbc_rank(orig(base + (p >> 2) * 128 + (p & 3)), L) : (0xffffffe0u + (uint32_t)p);
        int dst[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int before = 0;
#pragma unroll
            for (int q = 0; q < P; ++q) before += (rk[q] < rk[p]) ? 1 : 0;
            dst[p] = base + (before >> 2) * 128 + (before & 3);
        }
        {
            uint16_t v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = sk[base + (p >> 2) * 128 + (p & 3)];
#pragma unroll
            for (int p = 0; p < P; ++p) sk[dst[p]] = v[p];
        }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            float* const plane = sx + pl * CAP;
            float v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = plane[base + (p >> 2) * 128 + (p & 3)];
#pragma unroll
            for (int p = 0; p < P; ++p) plane[dst[p]] = v[p];
        }
    }
    {
        float bx0 = __int_as_float(0x7f800000), by0 = bx0, bz0 = bx0, bx1 = -bx0, by1 = -bx0, bz1 = -bx0;
#pragma unroll
        for (int c = 0; c < C4; ++c) {
            float4 X = make_float4(0.f, 0.f, 0.f, 0.f), Y = X, Z = X;
            if (has_pts) {
                X = *reinterpret_cast<const float4*>(sx + base + c * 128);
                Y = *reinterpret_cast<const float4*>(sy + base + c * 128);
                Z = *reinterpret_cast<const float4*>(sz + base + c * 128);
            }
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = c * 4 + e;
                float d0 = -2.f;
                if (has_pts && first + p < nloc) {
                    bx0 = fminf(bx0, xs[e]); bx1 = fmaxf(bx1, xs[e]);
                    by0 = fminf(by0, ys[e]); by1 = fmaxf(by1, ys[e]);
                    bz0 = fminf(bz0, zs[e]); bz1 = fmaxf(bz1, zs[e]);
                    d0 = a.temp ? a.temp[(size_t)cloud * n + orig(base + c * 128 + e)] : 1e10f;
                }
                md[p] = d0;
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(FULL, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(FULL, bx1, o));
            by0 = fminf(by0, __shfl_xor_sync(FULL, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(FULL, by1, o));
            bz0 = fminf(bz0, __shfl_xor_sync(FULL, bz0, o)); bz1 = fmaxf(bz1, __shfl_xor_sync(FULL, bz1, o));
        }
        if ((lane & 15) == 0) {
            sbox4[tid >> 4] = make_float4(bx0, bx1, by0, by1);
            sbox2[tid >> 4] = make_float2(bz0, bz1);
        }
    }
    if (tid < 32) {
        rkey[tid] = 0u;
        rpos[tid] = 0u;
        rkey2[tid] = 0u;
        rskey[tid] = 31u - (uint32_t)tid;
    }
    if (tid < KA) cand[tid] = make_float4(first_xyz[0], first_xyz[1], first_xyz[2], 0.f);
    if (tid == 0) *npick = 1;
    __syncthreads();

    float lmax = -2.f;
    uint32_t u = 0u, wu = 0u, wpos = (uint32_t)base, wu2 = 0u;
    int lpos = base;
    const float4 b4 = sbox4[tid >> 4];
    const float2 b2 = sbox2[tid >> 4];

    auto box_bound = [&](float x1, float y1, float z1) -> float {
        const float dx = fmaxf(fmaxf(__fsub_rn(b4.x, x1), __fsub_rn(x1, b4.y)), 0.f);
        const float dy = fmaxf(fmaxf(__fsub_rn(b4.z, y1), __fsub_rn(y1, b4.w)), 0.f);
        const float dz = fmaxf(fmaxf(__fsub_rn(b2.x, z1), __fsub_rn(z1, b2.y)), 0.f);
        return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    };
    auto update = [&](float x1, float y1, float z1) {
#pragma unroll
        for (int q = 0; q < C4; ++q) {
            const float4 X = *reinterpret_cast<const float4*>(sx + base + q * 128);
            const float4 Y = *reinterpret_cast<const float4*>(sy + base + q * 128);
            const float4 Z = *reinterpret_cast<const float4*>(sz + base + q * 128);
            md[q * 4 + 0] = fminf(sqdist3(x1, y1, z1, X.x, Y.x, Z.x), md[q * 4 + 0]);
            md[q * 4 + 1] = fminf(sqdist3(x1, y1, z1, X.y, Y.y, Z.y), md[q * 4 + 1]);
            md[q * 4 + 2] = fminf(sqdist3(x1, y1, z1, X.z, Y.z, Z.z), md[q * 4 + 2]);
            md[q * 4 + 3] = fminf(sqdist3(x1, y1, z1, X.w, Y.w, Z.w), md[q * 4 + 3]);
        }
    };
    // lane / warp argmax with the runner-up key (fps_bucket.cu, K > 1): leaves (wu, wpos, wu2)
    auto warp_argmax = [&]() {
        float t[P], t2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { t[p] = md[p]; t2[p] = -2.f; }
#pragma unroll
        for (int w = 1; w < P; w <<= 1) {
#pragma unroll
            for (int p = 0; p + w < P; p += 2 * w) {
                t2[p] = fmaxf(fminf(t[p], t[p + w]), fmaxf(t2[p], t2[p + w]));
                t[p] = fmaxf(t[p], t[p + w]);
            }
        }
        lmax = t[0];
        u = (lmax > -1.f) ? f32_ordered(lmax) : 0u;
        uint32_t eq = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) eq |= (md[p] == lmax) ? (1u << p) : 0u;
        const int bp = __ffs(eq) - 1;
        lpos = base + (bp >> 2) * 128 + (bp & 3);
        const bool several = (eq & (eq - 1u)) != 0u;
        wu = __reduce_max_sync(FULL, u);
        const unsigned tie = __ballot_sync(FULL, u == wu);
        wpos = __shfl_sync(FULL, (uint32_t)lpos | (several ? 0x80000000u : 0u), __ffs(tie) - 1);
        const bool lane_tie = (wpos >> 31) != 0u;
        wpos &= 0x7fffffffu;
        const uint32_t u2 = (t2[0] > -1.f) ? f32_ordered(t2[0]) : 0u;
        wu2 = __reduce_max_sync(FULL, (lane == __ffs(tie) - 1) ? u2 : u);
        if (wu != 0u && ((tie & (tie - 1u)) != 0u || lane_tie)) {
            // shared maximum inside the warp: the reference rank decides; exact duplicates of the candidate are zeroed
            // by picking it, so they do not count for the runner-up
            const uint32_t rk = (u == wu) ? fbc_rank(orig(lpos), L) : 0xffffffffu;
            const uint32_t wrk = __reduce_min_sync(FULL, rk);
            const int wl = __ffs(__ballot_sync(FULL, rk == wrk)) - 1;
            wpos = (uint32_t)__shfl_sync(FULL, lpos, wl);
            bool other = false;
            if (u == wu) {
                const float fx = sx[wpos], fy = sy[wpos], fz = sz[wpos];
                uint32_t rest = eq;
                while (rest) {
                    const int p = __ffs(rest) - 1;
                    rest &= rest - 1u;
                    const int e = base + (p >> 2) * 128 + (p & 3);
                    other |= sx[e] != fx || sy[e] != fy || sz[e] != fz;
                }
            }
            other = __any_sync(FULL, other);
            if (!other) {
                float s2 = -2.f;
#pragma unroll
                for (int p = 0; p < P; ++p) s2 = fmaxf(s2, (u == wu && md[p] == lmax) ? -2.f : md[p]);
                wu2 = __reduce_max_sync(FULL, (s2 > -1.f) ? f32_ordered(s2) : 0u);
            }
        }
    };

    const uint32_t nrec = csize * (uint32_t)kx;  // exchanged records per round (<= 32)
    if (BIG && crank == 0)
        for (int j = 1 + tid; j < m; j += T) idxs[j] = -1;
    cluster_sync_all();  // peers' mbarriers are initialised past this point

    constexpr int kBarPost = 1, kBarDone = 2;
    int c = 1;
    int round = 0;
    for (int j = 1; j < m; ++round) {
        const int par = round & 1;
        // ---- APPLY cand[0..c)
        if (has_pts) {
            unsigned hm = 0u;
#pragma unroll 1
            for (int k = 0; k < c; ++k) {
                const float4 pk = cand[k];
                hm |= (!(box_bound(pk.x, pk.y, pk.z) >= lmax) || round == 0) ? (1u << k) : 0u;  // NaN bounds count as hits
            }
            const unsigned wm = __reduce_or_sync(FULL, hm);
            if (wm != 0u) {
                const float cx = sx[wpos & 0x7fffffffu], cy = sy[wpos & 0x7fffffffu], cz = sz[wpos & 0x7fffffffu];
                const float cval = __uint_as_float(wu & 0x7fffffffu);
                bool redo = (wu & 0x80000000u) == 0u;
#pragma unroll 1
                for (int k = __ffs(wm) - 1; (wm >> k) != 0u; ++k) {  // first..last set bit only (fps_bucket.cu)
                    if ((wm >> k) & 1u) {
                        const float4 pk = cand[k];
                        redo |= sqdist3(pk.x, pk.y, pk.z, cx, cy, cz) < cval;
                        update(pk.x, pk.y, pk.z);
                    }
                }
                if (redo) {
                    warp_argmax();
                    if (lane == 0) {
                        rkey[warp] = wu;
                        rpos[warp] = wpos;
                        rkey2[warp] = wu2;
                        rskey[warp] = (wu & ~31u) | (31u - (uint32_t)warp);
                    }
                }
            }
        }
        if (warp != 0) {
            asm volatile("bar.arrive %0, %1;" ::"n"(kBarPost), "n"(T) : "memory");
        } else {
            if (lane == 0) mbar_arrive_expect_tx(smem_u32(&mbar[par]), nrec * 32u);
            asm volatile("bar.sync %0, %1;" ::"n"(kBarPost), "n"(T) : "memory");
            // ---- LOCAL: sorted position of every warp record
            const uint32_t mykey = rkey[lane], mysk = rskey[lane];
            int pos = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 k4 = reinterpret_cast<const uint4*>(rskey)[q];
                pos += (k4.x > mysk ? 1 : 0) + (k4.y > mysk ? 1 : 0) + (k4.z > mysk ? 1 : 0) + (k4.w > mysk ? 1 : 0);
            }
            inv[pos] = (uint32_t)lane;
            __syncwarp();
            // lane k adopts the record at position k (k <= kx: position kx supplies the bound of the unlisted ones)
            uint32_t src = inv[lane], qkey = rkey[src], qsk = rskey[src];
            const uint32_t nsk = __shfl_down_sync(FULL, qsk, 1);
            const uint32_t nkey = __shfl_down_sync(FULL, qkey, 1);
            // exact local argmax on a near tie at the top: largest key, smallest reference rank (then the list is cut)
            const bool top_tie = __shfl_sync(FULL, (qkey != 0u && nkey != 0u && (qsk >> 5) == (nsk >> 5)) ? 1 : 0, 0) != 0;
            int nlist = kx;
            uint32_t ub;
            if (top_tie) {
                const uint32_t bucket = __shfl_sync(FULL, qsk >> 5, 0);
                const bool in = mykey != 0u && (mysk >> 5) == bucket;
                const uint32_t bk = __reduce_max_sync(FULL, in ? mykey : 0u);
                const uint32_t rk = (in && mykey == bk) ? fbc_rank(orig((int)(rpos[lane] & 0x7fffffffu)), L) : 0xffffffffu;
                const uint32_t brk = __reduce_min_sync(FULL, rk);
                const int wl = __ffs(__ballot_sync(FULL, rk == brk)) - 1;
                if (lane == 0) {
                    src = (uint32_t)wl;
                    qkey = rkey[src];
                }
                nlist = 1;
                ub = bk;  // unlisted records reach up to the top key
            } else {
                ub = __shfl_sync(FULL, qkey, kx);  // kx < 32; 0 when fewer eligible warps
            }
            // ---- EXCHANGE: lane = k * csize + peer sends list entry k to CTA `peer`
            {
                const int k = lane / (int)csize, peer = lane - k * (int)csize;
                const uint32_t skey = __shfl_sync(FULL, qkey, k), ssrc = __shfl_sync(FULL, src, k);
                if ((uint32_t)lane < nrec) {
                    const bool live = k < nlist && skey != 0u;
                    const int e = (int)(rpos[ssrc] & 0x7fffffffu);
                    const uint32_t kk = BIG ? ((crank << 16) | (uint32_t)sk[e]) : (uint32_t)sk[e];
                    const uint32_t dr = mapa_u32(smem_u32(&xl[par * 32 + (int)crank * kx + k]), (uint32_t)peer);
                    const uint32_t db = mapa_u32(smem_u32(&mbar[par]), (uint32_t)peer);
                    st_async_v4(dr, db, live ? skey : 0u, kk, __float_as_uint(sx[e]), __float_as_uint(sy[e]));
                    st_async_v4(dr + 16, db, __float_as_uint(sz[e]), rkey2[ssrc], ub, 0u);
                }
            }
            {
                const uint32_t bar = smem_u32(&mbar[par]), ph = (uint32_t)((round >> 1) & 1);
                if (!mbar_try_wait_cluster(bar, ph)) {
                    const long long t0 = clock64();
                    while (!mbar_try_wait_cluster(bar, ph))
                        if (clock64() - t0 > 4000000000LL) watchdog_trip(a.status, TSM_ERR_WATCHDOG);
                }
            }
            // ---- GLOBAL: the same decision in every CTA
            MpRec r;
            r.key = 0u; r.kk = 0u; r.x = 0.f; r.y = 0.f; r.z = 0.f; r.r2 = 0u; r.ubound = 0u; r.pad = 0u;
            if ((uint32_t)lane < nrec) r = xl[par * 32 + lane];
            const uint32_t umax = __reduce_max_sync(FULL, r.ubound);
            const uint32_t gkey = (r.key & ~31u) | (31u - (uint32_t)lane);  // unique; 0-keys sort last
            gsk[lane] = gkey;
            gsk[32] = 0u;
            gsk[33] = 0u;
            __syncwarp();
            int gpos = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 k4 = reinterpret_cast<const uint4*>(gsk)[q];
                gpos += (k4.x > gkey ? 1 : 0) + (k4.y > gkey ? 1 : 0) + (k4.z > gkey ? 1 : 0) + (k4.w > gkey ? 1 : 0);
            }
            __syncwarp();
            inv[gpos] = (uint32_t)lane;
            __syncwarp();
            const int gsrc = (int)inv[lane];  // the record at global position `lane`
            // position 0: plain argmax over its 32-ulp bucket (largest key, smallest reference rank)
            const uint32_t topb = __reduce_max_sync(FULL, r.key) >> 5;
            int win = -1;
            {
                const bool in = r.key != 0u && (r.key >> 5) == topb;
                const uint32_t bk = __reduce_max_sync(FULL, in ? r.key : 0u);
                const unsigned cm = __ballot_sync(FULL, in && r.key == bk);
                if (cm != 0u) {
                    win = __ffs(cm) - 1;
                    if ((cm & (cm - 1u)) != 0u) {
                        const uint32_t rk = ((cm >> lane) & 1u) ? fbc_rank(orig_of_kk(r.kk), L) : 0xffffffffu;
                        const uint32_t brk = __reduce_min_sync(FULL, rk);
                        win = __ffs(__ballot_sync(FULL, rk == brk)) - 1;
                    }
                }
            }
            const unsigned topcnt = __popc(__ballot_sync(FULL, r.key != 0u && (r.key >> 5) == topb));
            // lane k = position k: fetch its record (position 0: the exact winner)
            const int take = lane == 0 ? (win >= 0 ? win : 0) : gsrc;
            const uint32_t qk = __shfl_sync(FULL, r.key, take), qkk = __shfl_sync(FULL, r.kk, take), qr2 = __shfl_sync(FULL, r.r2, take);
            float qx = __shfl_sync(FULL, r.x, take), qy = __shfl_sync(FULL, r.y, take), qz = __shfl_sync(FULL, r.z, take);
            const bool valid0 = win >= 0;
            if (lane == 0 && !valid0) {  // no eligible candidate anywhere: the reference yields index 0
                qx = first_xyz[0];
                qy = first_xyz[1];
                qz = first_xyz[2];
            }
            if (lane < KA) {
                cand[lane] = make_float4(qx, qy, qz, 0.f);
                ckey[lane] = qk;
                cr2[lane] = qr2;
            }
            __syncwarp();
            bool pmoved = false;
            {
                int hi = 1;
#pragma unroll
                for (int t2 = 2; t2 < KA; ++t2) hi += (lane >= t2 * (t2 - 1) / 2) ? 1 : 0;
                const int lo2 = lane - hi * (hi - 1) / 2;
                if (lane < KA * (KA - 1) / 2) {
                    const float4 pl = cand[lo2], ph = cand[hi];
                    pmoved = sqdist3(pl.x, pl.y, pl.z, ph.x, ph.y, ph.z) < __uint_as_float(ckey[hi] & 0x7fffffffu);
                }
            }
            const unsigned pm = __ballot_sync(FULL, pmoved);
            const bool moved = lane < KA && ((pm >> (lane * (lane - 1) / 2)) & ((1u << lane) - 1u)) != 0u;
            uint32_t run2 = 0u;
#pragma unroll
            for (int q = 0; q < KA / 4; ++q) {
                const uint4 v = reinterpret_cast<const uint4*>(cr2)[q];
                run2 = max(run2, 4 * q + 0 < lane ? v.x : 0u);
                run2 = max(run2, 4 * q + 1 < lane ? v.y : 0u);
                run2 = max(run2, 4 * q + 2 < lane ? v.z : 0u);
                run2 = max(run2, 4 * q + 3 < lane ? v.w : 0u);
            }
            // (i): alone in its bucket (sorted neighbours), the top bucket held one record, above every unlisted bound
            const uint32_t myb = gsk[gsrc] >> 5, nb = __shfl_down_sync(FULL, myb, 1), pb = __shfl_up_sync(FULL, myb, 1);
            const bool alone = myb != pb && (lane == 31 || myb != nb) && topcnt == 1u && myb > (umax >> 5);
            const bool ok = lane == 0 || (lane < KA && valid0 && qk != 0u && alone && j + lane < m && qk > run2 &&
                                          qk > 0x80000000u && qk < 0xff800000u && !moved);
            const unsigned okm = __ballot_sync(FULL, ok);
            const int cnt = __ffs(~okm) - 1;
            if (lane == 0) *npick = cnt;
            if (crank == 0 && lane < cnt) idxs[j + lane] = (lane == 0 && !valid0) ? (BIG ? -1 : 0) : (int)qkk;
        }
        asm volatile("bar.sync %0, %1;" ::"n"(kBarDone), "n"(T) : "memory");
        c = *npick;
        j += c;
    }
    if (a.temp != nullptr && c > 1 && has_pts) {
        // temp leaves with the min-distances to picks 0..m-2: the last round's picks but the final one are outstanding
#pragma unroll 1
        for (int k = 0; k + 1 < c; ++k) {
            const float4 pk = cand[k];
            update(pk.x, pk.y, pk.z);
        }
    }
    if (BIG && crank == 0) {
        __syncthreads();
        for (int j = 1 + tid; j < m; j += T) {
            const int code = idxs[j];
            idxs[j] = code >= 0 ? __float_as_int(sorted[(code >> 16) * per_cta + (code & 0xffff)].w) : 0;
        }
    }
    if (a.temp && has_pts) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (first + p < nloc) a.temp[(size_t)cloud * n + orig(base + (p >> 2) * 128 + (p & 3))] = md[p];
    }
    cluster_sync_all();
}

}  // namespace tsm

bool tsm_fps_bucket_cluster_supports(int n, bool weighted) { return !weighted && n > 16384 && n <= 240000; }

int tsm_fps_bucket_cluster_launch(const tsm::FpsArgs& a, int b, cudaStream_t stream) {
    using namespace tsm;
    const int n = a.n;
    if (!tsm_fps_bucket_cluster_supports(n, a.weights != nullptr)) return TSM_ERR_INVALID;
    // slices of at most ~15000 points (whole 512-point warps), so every CTA keeps a little slack
    int csize = (n + 14999) / 15000;
    if (csize < 2) csize = 2;
    int per_cta = ((n + csize - 1) / csize + 511) / 512 * 512;
    if (per_cta > FBC_CAP || csize > FBC_MAXC) return TSM_ERR_INVALID;
    void* scratch = nullptr;
    int rc = tsm_scratch_get(6, (size_t)b * n * sizeof(float4), stream, &scratch);
    if (rc != TSM_OK) return rc;
    float4* sorted = static_cast<float4*>(scratch);
    const size_t sort_dyn = (size_t)(1 << FBC_CELL_BITS) * sizeof(uint32_t);
    TSM_CUDA_TRY(cudaFuncSetAttribute(fps_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_dyn));
    fps_sort_kernel<<<b, 1024, sort_dyn, stream>>>(n, a.xyz, sorted);
    TSM_LAUNCH_CHECK();
    const bool big = n > 65536;
    // rounds of several picks per exchange (TSMDET_FPSC_K=1: one pick per exchange): kx list entries per CTA, <= 32 in all CTAs
    const int kx_max = 32 / csize < FBM_KA ? 32 / csize : FBM_KA;
    int kx = kx_max < 6 ? kx_max : 6;  // measured (B200): 20000 -> 4096 on two CTAs 2.37 ms with 4, 2.25 ms with 6 or 8
    if (const char* e = tsm_knob(KNOB_FPSC_K)) kx = atoi(e) <= 1 ? 0 : (atoi(e) < kx_max ? atoi(e) : kx_max);
    if (per_cta > 15360 || kx * csize > 32) kx = 0;
    using Kern1 = void (*)(const FpsArgs, const float4*, int);
    using KernM = void (*)(const FpsArgs, const float4*, int, int);
    Kern1 kern1 = big ? fps_bucket_cluster_kernel<true> : fps_bucket_cluster_kernel<false>;
    KernM kernm = big ? fps_bucket_cluster_mp_kernel<true> : fps_bucket_cluster_mp_kernel<false>;
    const void* kfn = kx > 0 ? (const void*)kernm : (const void*)kern1;
    TSM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, FBC_SMEM));
    if (csize > 8) TSM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * csize));
    cfg.blockDim = dim3(FBC_T);
    cfg.dynamicSmemBytes = FBC_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (kx > 0) TSM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernm, a, (const float4*)sorted, per_cta, kx));
    else TSM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern1, a, (const float4*)sorted, per_cta));
    return TSM_OK;
}
