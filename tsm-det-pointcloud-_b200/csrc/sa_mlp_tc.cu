// sa_mlp_tc.cu -- fused set-abstraction scale on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Same operator as sa_mlp_fp32.cu (gather -> mask -> [1x1 conv + ReLU]* -> max over nsample, see
// /root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py:1259-1268, 1297-1300),
// with the contraction in bf16 x bf16 -> fp32 on tcgen05.mma (SASS: UTCHMMA), accumulators in
// TMEM, and the max-pool fused into the last epilogue.  The only dense contraction on this
// path, hence the only place tensor cores are used.
//
// One CTA = 128 threads = one 128-row tile of (centre, sample) rows, persistent over tiles:
//   gather   : thread r builds row r of the layer-0 operand in shared memory, straight in the UMMA
//              K-major "no swizzle" core-matrix layout (8 rows x 16 B cores; row stride 16 B, K-chunk
//              stride 128 rows x 16 B) -- features come from a bf16 (B,N,Cp) transpose of the input,
//              so a row is one contiguous run of 16-byte chunks; xyz offsets fill one extra chunk;
//   layer l  : one thread issues K_l/16 tcgen05.mma (M=128, N=N_l) reading A (activations) and
//              B (weights, resident in shared memory for the whole kernel) through smem descriptors,
//              commits to an mbarrier; the 4 warps then read their 32 TMEM lanes (tcgen05.ld
//              32x32b), add bias, ReLU, repack to bf16 and write the next layer's A operand in place;
//   last     : bias + ReLU, then max over the S rows of each centre: values are >= 0, so the float
//              max is an unsigned max on the bit pattern -> ONE redux.sync per channel per warp.
#include "sa_mlp.cuh"
#include "umma.cuh"

namespace tsm {

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 128;
constexpr int TC_MAX_LAYERS = 4;

struct TcPlan {
    int nl;
    int K[TC_MAX_LAYERS];      // padded input channels of layer l (multiple of 16)
    int Npad[TC_MAX_LAYERS];   // padded output channels (multiple of 16, 16..256)
    int w_off[TC_MAX_LAYERS];  // byte offset of layer l's weights in dynamic smem
    int b_off[TC_MAX_LAYERS];  // byte offset of layer l's bias (fp32)
    int a_off;                 // activation operand buffer
    int o_off;                 // pooled output staging (uint32 [Nlast][centres per tile])
    int smem_bytes;
    int cp;                    // feature channels padded to 8 (width of the bf16 transpose)
    int xyz_chunk;             // 16-byte chunk index of [dx,dy,dz,0...] in layer-0 rows, -1 if unused
    int a_bytes, o_bytes;      // per tile group
    int grp_cols;              // TMEM columns of one tile group
    int tmem_cols;             // power of two >= 32 (all groups)
    int d_off;                 // TMEM column of the odd layers' accumulator (even layers use column 0)
    int packed_bytes;          // weights + biases: the first packed_bytes of dynamic smem, same layout in global
};

__device__ __forceinline__ uint32_t instr_desc_bf16(int n) { return instr_desc_bf16_m128(n); }

// features (B,C,N) fp32 -> (B,N,Cp) bf16, zero padded to Cp channels
__global__ void __launch_bounds__(256) transpose_bf16_kernel(int c, int cp, int n, const float* __restrict__ f,
                                                             __nv_bfloat16* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float* src = f + (size_t)b * c * n + i;
    __nv_bfloat16* dst = out + ((size_t)b * n + i) * cp;
    for (int c0 = 0; c0 < cp; c0 += 8) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ca = c0 + 2 * j, cb = ca + 1;
            const float a = ca < c ? __ldg(src + (size_t)ca * n) : 0.f;
            const float bb = cb < c ? __ldg(src + (size_t)cb * n) : 0.f;
            w[j] = pack_bf16(a, bb);
        }
        *reinterpret_cast<uint4*>(dst + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Weights (cout,cin) fp32 + bias -> the kernel's shared-memory image: per layer bf16 [K/8][Npad][8] (UMMA
// K-major core matrices), then the fp32 biases.  Layer 0's input channels are permuted to the operand order
// [features 0..C-1 | pad to cp | dx,dy,dz | pad] (the reference order is [dx,dy,dz, features...],
// pointnet2_utils.py:523).
__global__ void __launch_bounds__(256) pack_weights_kernel(const SaMlpArgs a, const TcPlan pl, unsigned char* __restrict__ packed) {
    const int l = blockIdx.y;
    const int K = pl.K[l], Np = pl.Npad[l];
    const int cin = a.ch[l], cout = a.ch[l + 1];
    __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(packed + pl.w_off[l]);
    const float* __restrict__ W = a.w[l];
    for (int e = blockIdx.x * 256 + threadIdx.x; e < Np * K; e += gridDim.x * 256) {
        const int k = e / Np, n = e - k * Np;  // consecutive threads -> consecutive n: 16-byte-strided writes
        float v = 0.f;
        if (n < cout) {
            int src = -1;
            if (l == 0) {
                if (k < a.c_feat)
                    src = (a.use_xyz ? 3 : 0) + k;
                else if (pl.xyz_chunk >= 0 && k >= pl.xyz_chunk * 8 && k < pl.xyz_chunk * 8 + 3)
                    src = k - pl.xyz_chunk * 8;
            } else if (k < cin) {
                src = k;
            }
            if (src >= 0) v = __ldg(W + (size_t)n * cin + src);
        }
        ws[(size_t)(k >> 3) * (Np * 8) + n * 8 + (k & 7)] = __float2bfloat16_rn(v);
    }
    float* bs = reinterpret_cast<float*>(packed + pl.b_off[l]);
    for (int e = blockIdx.x * 256 + threadIdx.x; e < Np; e += gridDim.x * 256) bs[e] = e < cout ? __ldg(a.bias[l] + e) : 0.f;
}

// GROUPS = 1: the CTA is one 128-thread tile pipeline.  GROUPS = 2 (used when the weights leave room for only one
// CTA per SM): two independent 128-thread pipelines share the resident weights, each with its own operand
// buffer, TMEM columns, mbarrier and named barrier, so one group's gather / epilogue overlaps the other's MMAs.
// SUBWARP: nsample < 32 (several centres per warp: butterfly pooling) -- a separate instantiation so the common
// nsample >= 32 code does not carry its registers.
template <int GROUPS, bool SUBWARP>
__global__ void __launch_bounds__(TC_THREADS * GROUPS, 1)
    sa_mlp_tc_kernel(const SaMlpArgs a, const TcPlan pl, const __nv_bfloat16* __restrict__ featT,
                     const unsigned char* __restrict__ packed, int num_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mma_bars[GROUPS];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_s;

    const int grp = GROUPS == 1 ? 0 : (int)(threadIdx.x >> 7);
    const int tid = threadIdx.x & 127, lane = tid & 31, warp = tid >> 5;  // within the group
    uint64_t& mma_bar = mma_bars[grp];
    auto group_sync = [&]() {
        if (GROUPS == 1) __syncthreads();
        else asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    };
    const int S = a.s, M = a.m;
    const int cpt = TC_ROWS / S;  // whole centres per tile (S divides 128)
    const int log2s = 31 - __clz(S), log2cpt = 31 - __clz(cpt);
    const int nl = pl.nl;

    // ---- one-time setup: packed weights + biases (bf16 UMMA layout, built by pack_weights_kernel) arrive
    // with ONE bulk copy (TMA engine); barrier; TMEM
    if (threadIdx.x == 0) {
        for (int g = 0; g < GROUPS; ++g) mbar_init(smem_u32(&mma_bars[g]), 1);
        mbar_init(smem_u32(&w_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(smem_u32(&w_bar), (uint32_t)pl.packed_bytes);
        bulk_g2s(smem_u32(smem), packed, (uint32_t)pl.packed_bytes, smem_u32(&w_bar));
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const uint32_t bar = smem_u32(&w_bar);
        const long long t0 = clock64();
        while (!mbar_try_wait_cta(bar, 0))
            if (clock64() - t0 > 4000000000LL) watchdog_trip(a.status, TSM_ERR_WATCHDOG);
    }
    const uint32_t tmem_base = tmem_base_s + (uint32_t)(grp * pl.grp_cols);
    const uint32_t a_smem = smem_u32(smem + pl.a_off + grp * pl.a_bytes);
    unsigned char* a_ptr = smem + pl.a_off + grp * pl.a_bytes;
    uint32_t* omax = reinterpret_cast<uint32_t*>(smem + pl.o_off + grp * pl.o_bytes);
    uint32_t phase = 0;
    const int Nlast = pl.Npad[nl - 1];
    const int cout_last = a.ch[nl];

    for (int tile = blockIdx.x * GROUPS + grp; tile < num_tiles; tile += gridDim.x * GROUPS) {
        // ---- gather: row `tid`
        {
            const long long g = (long long)tile * TC_ROWS + tid;
            const bool rv = g < a.total_rows;
            const long long cpi = rv ? g >> log2s : 0;   // S is a power of two; B*M < 2^31 (checked by the launcher)
            const int b = (int)((unsigned)cpi / (unsigned)M);
            bool live = rv;
            int id = 0;
            if (rv) {
                id = __ldg(a.idx + g);
                if (a.idx_cnt && __ldg(a.idx_cnt + cpi) <= 0) live = false;  // empty ball -> zero input row
            }
            const int nchunk0 = pl.K[0] >> 3;
            const int fchunks = pl.cp >> 3;
            const uint4* frow = reinterpret_cast<const uint4*>(featT + ((size_t)b * a.n + id) * pl.cp);
            // feature chunks four at a time: four independent 16-byte loads in flight per thread, then four stores
            for (int kc0 = 0; kc0 < nchunk0; kc0 += 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kc = kc0 + u;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (live && kc < fchunks) v[u] = __ldg(frow + kc);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kc = kc0 + u;
                    if (kc >= nchunk0) break;
                    if (live && kc == pl.xyz_chunk) {
                        const float* p = a.xyz + ((size_t)b * a.n + id) * 3;
                        const float* q = a.new_xyz + (size_t)cpi * 3;
                        const float dx = __fsub_rn(__ldg(p + 0), __ldg(q + 0));
                        const float dy = __fsub_rn(__ldg(p + 1), __ldg(q + 1));
                        const float dz = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
                        v[u].x = pack_bf16(dx, dy);
                        v[u].y = pack_bf16(dz, 0.f);
                    }
                    *reinterpret_cast<uint4*>(a_ptr + (size_t)kc * (TC_ROWS * 16) + tid * 16) = v[u];
                }
            }
            if (S > 32)
                for (int e = tid; e < Nlast * cpt; e += TC_THREADS) omax[e] = 0u;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_sync();

        for (int l = 0; l < nl; ++l) {
            const int K = pl.K[l], Np = pl.Npad[l];
            const uint32_t d_tmem = tmem_base + (uint32_t)((l & 1) ? pl.d_off : 0);
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t w_smem = smem_u32(smem + pl.w_off[l]);
                const uint32_t idesc = instr_desc_bf16(Np);
                const uint32_t a_lbo = TC_ROWS * 16, b_lbo = (uint32_t)Np * 16;
                for (int kk = 0; kk < (K >> 4); ++kk) {
                    const uint64_t ad = smem_desc(a_smem + (uint32_t)kk * 2u * a_lbo, a_lbo, 128);
                    const uint64_t bd = smem_desc(w_smem + (uint32_t)kk * 2u * b_lbo, b_lbo, 128);
                    umma_bf16(d_tmem, ad, bd, idesc, kk > 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&mma_bar));
            }
            {
                const uint32_t bar = smem_u32(&mma_bar);
                if (!mbar_try_wait_cta(bar, phase)) {
                    const long long t0 = clock64();
                    while (!mbar_try_wait_cta(bar, phase))
                        if (clock64() - t0 > 4000000000LL) watchdog_trip(a.status, TSM_ERR_WATCHDOG);
                }
                phase ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            const float* bs = reinterpret_cast<const float*>(smem + pl.b_off[l]);
            const uint32_t t_row = d_tmem + ((uint32_t)(warp * 32) << 16);
            if (l + 1 < nl) {
                // bias + ReLU -> bf16 -> next layer's A operand (written over the consumed one)
                for (int c0 = 0; c0 < Np; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + (uint32_t)c0, v);
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float x0 = fmaxf(__uint_as_float(v[2 * j]) + bs[c0 + 2 * j], 0.f);
                        const float x1 = fmaxf(__uint_as_float(v[2 * j + 1]) + bs[c0 + 2 * j + 1], 0.f);
                        w[j] = pack_bf16(x0, x1);
                    }
                    unsigned char* dst = a_ptr + (size_t)(c0 >> 3) * (TC_ROWS * 16) + tid * 16;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(dst + TC_ROWS * 16) = make_uint4(w[4], w[5], w[6], w[7]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                group_sync();
            } else {
                // bias + ReLU + max over the S rows of each centre
                const int ci = tid / S;  // centre within the tile
                for (int c0 = 0; c0 < Np; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + (uint32_t)c0, v);
                    if constexpr (!SUBWARP) {
                        uint32_t mine = 0u;  // lane j ends up holding channel c0 + j's maximum
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x = fmaxf(__uint_as_float(v[j]) + bs[c0 + j], 0.f);
                            // x >= 0: unsigned order == float order, so one redux.sync is the 32-row max
                            const uint32_t u = __reduce_max_sync(FULL, __float_as_uint(x));
                            mine = lane == j ? u : mine;
                        }
                        if (lane < 16) {
                            if (S == 32)
                                omax[(c0 + lane) * cpt + ci] = mine;
                            else
                                atomicMax(&omax[(c0 + lane) * cpt + ci], mine);
                        }
                    } else {
                        // S < 32: butterfly over the S rows of a centre in which every step also halves the
                        // channels a lane carries (15 shuffles per 16 channels instead of 16 * log2 S)
                        uint32_t u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) u[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) + bs[c0 + j], 0.f));
                        int base = 0, cnt = 16;
#pragma unroll
                        for (int st = 0; st < 4; ++st) {
                            const int d = S >> (st + 1);  // 8, 4, 2, 1 for S = 16; fewer steps for smaller S
                            if (d >= 1) {
                                constexpr int kHalf[4] = {8, 4, 2, 1};
                                const int half = kHalf[st];
                                const bool upper = (lane & d) != 0;
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    if (i < half) {
                                        const uint32_t mine = upper ? u[half + i] : u[i];
                                        const uint32_t give = upper ? u[i] : u[half + i];
                                        u[i] = max(mine, __shfl_xor_sync(FULL, give, d));
                                    }
                                }
                                if (upper) base += half;
                                cnt = half;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < cnt) omax[(c0 + base + i) * cpt + ci] = u[i];  // 16 / S channels per lane
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                group_sync();
                // pooled tile -> out[b, out_c0 + c, p]; consecutive threads write consecutive centres
                const unsigned cbase = (unsigned)tile * (unsigned)cpt;
                const unsigned ctot = (unsigned)(a.total_rows >> log2s);
                for (int e = tid; e < cout_last * cpt; e += TC_THREADS) {
                    const int c = e >> log2cpt, cc = e & (cpt - 1);
                    const unsigned cg = cbase + (unsigned)cc;
                    if (cg >= ctot) continue;
                    const int b2 = (int)(cg / (unsigned)M);
                    const int p2 = (int)(cg - (unsigned)b2 * (unsigned)M);
                    a.out[((size_t)b2 * a.out_ctot + a.out_c0 + c) * M + p2] = __uint_as_float(omax[c * cpt + cc]);
                }
                group_sync();  // omax / A buffer are reused by the next tile
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"((uint32_t)pl.tmem_cols)
                     : "memory");
    }
}

}  // namespace tsm

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

int tsm_sa_mlp_tc(const tsm::SaMlpArgs& a, int b, cudaStream_t stream) {
    using namespace tsm;
    const int S = a.s;
    if (S > TC_ROWS || (TC_ROWS % S) != 0 || (S & (S - 1)) != 0) return TSM_ERR_INVALID;  // whole centres per tile
    if ((long long)b * a.m >= 0x7fffffffLL) return TSM_ERR_INVALID;                         // 32-bit centre arithmetic
    if (a.num_layers < 1 || a.num_layers > TC_MAX_LAYERS) return TSM_ERR_INVALID;
    TcPlan pl;
    pl.nl = a.num_layers;
    pl.cp = round_up(a.c_feat, 8);
    pl.xyz_chunk = a.use_xyz ? (pl.cp >> 3) : -1;
    pl.K[0] = round_up(pl.cp + (a.use_xyz ? 8 : 0), 16);
    int off = 0, kmax = pl.K[0], nmax = 0;
    for (int l = 0; l < pl.nl; ++l) {
        pl.Npad[l] = round_up(a.ch[l + 1], 16);
        if (pl.Npad[l] > 256) return TSM_ERR_INVALID;
        if (l > 0) pl.K[l] = pl.Npad[l - 1];
        kmax = pl.K[l] > kmax ? pl.K[l] : kmax;
        nmax = pl.Npad[l] > nmax ? pl.Npad[l] : nmax;
        pl.w_off[l] = off;
        off += pl.Npad[l] * pl.K[l] * 2;
    }
    if (kmax > 512) return TSM_ERR_INVALID;
    for (int l = 0; l < pl.nl; ++l) {
        pl.b_off[l] = off;
        off += pl.Npad[l] * 4;
    }
    pl.packed_bytes = off;  // multiple of 64
    off = round_up(off, 128);
    pl.a_off = off;
    pl.a_bytes = round_up(TC_ROWS * kmax * 2, 128);
    pl.o_bytes = round_up(pl.Npad[pl.nl - 1] * (TC_ROWS / S) * 4, 16);
    const int fixed = off;
    auto smem_for = [&](int groups) { return fixed + groups * (pl.a_bytes + pl.o_bytes); };
    if (smem_for(1) > 227 * 1024 - 64) return TSM_ERR_INVALID;
    // One accumulator region serves every layer: an epilogue's tcgen05.ld's are complete (wait::ld) before the
    // barrier that precedes the next layer's MMA, so nothing races.  (Alternating regions, d_off = nmax, halved
    // the CTAs per SM the TMEM allows; TSMDET_MLP_TMEM_ALT=1 restores it for experiments.)
    pl.d_off = 0;
    pl.grp_cols = 32;
    while (pl.grp_cols < pl.d_off + nmax) pl.grp_cols <<= 1;
    if (pl.grp_cols > 512) return TSM_ERR_INVALID;
    // two tile groups per CTA when the weights allow only one CTA per SM but a second operand buffer still fits
    int groups = 1;
    if ((227 * 1024) / (smem_for(1) + 2048) < 2 && smem_for(2) <= 227 * 1024 - 64 && 2 * pl.grp_cols <= 512 &&
        !tsm_knob(KNOB_MLP_ONE_GROUP))
        groups = 2;
    pl.tmem_cols = pl.grp_cols * groups;
    pl.o_off = pl.a_off + groups * pl.a_bytes;
    pl.smem_bytes = smem_for(groups);

    // bf16 (B,N,Cp) transpose of the features
    __nv_bfloat16* featT = nullptr;
    if (a.c_feat > 0) {
        void* p = nullptr;
        const size_t bytes = (size_t)b * a.n * pl.cp * sizeof(__nv_bfloat16);
        int rc = tsm_scratch_get(1, bytes, stream, &p);
        if (rc != TSM_OK) return rc;
        featT = (__nv_bfloat16*)p;
        dim3 grid((unsigned)divup(a.n, 256), (unsigned)b);
        transpose_bf16_kernel<<<grid, 256, 0, stream>>>(a.c_feat, pl.cp, a.n, a.features, featT);
        TSM_LAUNCH_CHECK();
    }
    const long long tiles = (a.total_rows + TC_ROWS - 1) / TC_ROWS;
    if (tiles > 0x7fffffffLL) return TSM_ERR_INVALID;
    const bool subwarp = S < 32;
    auto kern = groups == 2 ? (subwarp ? sa_mlp_tc_kernel<2, true> : sa_mlp_tc_kernel<2, false>)
                            : (subwarp ? sa_mlp_tc_kernel<1, true> : sa_mlp_tc_kernel<1, false>);
    TSM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    // resident CTAs per SM: what shared memory, registers (a grid of more CTAs than are resident runs a second,
    // partial wave: measured 0.094 -> 0.121 ms on layer 1) and the 512 TMEM columns allow
    int occ = (227 * 1024) / (pl.smem_bytes + 2048);
    {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e == cudaSuccess && fa.numRegs > 0) {
            const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * TC_THREADS * groups;
            const int occ_regs = 65536 / regs_cta;
            if (occ > occ_regs) occ = occ_regs;
        } else {
            cudaGetLastError();
            if (occ > 4) occ = 4;
        }
    }
    const int tmem_occ = 512 / pl.tmem_cols;
    if (occ > tmem_occ) occ = tmem_occ;
    if (occ < 1) occ = 1;
    if (const char* e = tsm_knob(KNOB_MLP_OCC)) occ = atoi(e) > 0 && atoi(e) < occ ? atoi(e) : occ;
    long long grid = (long long)tsm_num_sms() * occ;
    if (grid * groups > tiles) grid = (tiles + groups - 1) / groups;
    SaMlpArgs args = a;
    args.status = tsm_status_word(stream);
    unsigned char* packed = nullptr;
    {
        void* p = nullptr;
        int rc = tsm_scratch_get(2, (size_t)pl.packed_bytes, stream, &p);
        if (rc != TSM_OK) return rc;
        packed = (unsigned char*)p;
        dim3 pgrid(16, (unsigned)pl.nl);
        pack_weights_kernel<<<pgrid, 256, 0, stream>>>(args, pl, packed);
        TSM_LAUNCH_CHECK();
    }
    kern<<<(unsigned)grid, TC_THREADS * groups, pl.smem_bytes, stream>>>(args, pl, featT, packed, (int)tiles);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}
