// sa_mlp_fp32.cu -- fused set-abstraction scale, fp32 CUDA-core path (parity mode).
//
// Replaces, for one scale of the layer-0 branch of
//   _VoxelPointnetSAModuleFS(Distillation)Base.forward
//   (/root/reference/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py:1259-1268, 1297-1300)
// the chain  grouping_operation x2 -> subtract -> cat -> mask -> [Conv2d 1x1, BN, ReLU] x L -> max_pool2d
// with ONE kernel that never writes a (B,C,npoint,nsample) tensor.  BatchNorm (eval) is folded
// into the conv by the host shim: w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps).
// This file is the fp32 FMA version (max-abs 1e-5 class parity against the eager fp32 stack);
// sa_mlp_tc.cu holds the tcgen05 bf16 version.
//
// Work split: a CTA owns ROWS = 64 (centre, sample) rows; activations live in shared memory as
// [channel][row] planes (conflict-free: a warp reads 32 consecutive rows of one channel), each
// thread owns one row x 4-8 output channels per pass, weights stream through the read-only path.
#include "sa_mlp.cuh"

namespace tsm {

constexpr int MLP_ROWS = 64;
constexpr int MLP_THREADS = 256;  // 64 rows x 4 channel lanes
constexpr int MLP_MAXC = 512;

// dynamic smem: two activation buffers of maxc * MLP_ROWS floats
__global__ void __launch_bounds__(MLP_THREADS) sa_mlp_fp32_kernel(const SaMlpArgs a, int maxc) {
    extern __shared__ float smem[];
    float* bufA = smem;
    float* bufB = smem + (size_t)maxc * MLP_ROWS;
    const int tid = threadIdx.x;
    const int row = tid & (MLP_ROWS - 1);
    const int cl = tid >> 6;  // channel lane 0..3
    const long long r0 = (long long)blockIdx.x * MLP_ROWS;
    const long long gr = r0 + row;
    const bool rv = gr < a.total_rows;
    const int S = a.s, M = a.m;
    long long cp = rv ? gr / S : 0;  // global centre index b*M + p
    const int b = (int)(cp / M);

    // ---- gather the input rows: channel order = [dx,dy,dz, features...] (pointnet2_utils.py:523)
    {
        int id = 0;
        bool live = rv;
        if (rv) {
            // dense (point-wise) mode: no index tensor, S = 1, M = n -> row gr is point gr % n of batch entry b
            id = a.idx ? a.idx[gr] : (int)(cp - (long long)b * M);
            if (a.idx_cnt && a.idx_cnt[cp] <= 0) live = false;  // empty ball: all-zero input (:1265-1267)
        }
        const int c0 = a.ch[0];
        for (int c = cl; c < c0; c += 4) {
            float v = 0.f;
            if (live) {
                if (a.use_xyz && c < 3) {
                    v = __fsub_rn(__ldg(a.xyz + ((size_t)b * a.n + id) * 3 + c), __ldg(a.new_xyz + (size_t)cp * 3 + c));
                } else {
                    const int fc = c - (a.use_xyz ? 3 : 0);
                    v = fc < a.c_feat ? __ldg(a.features + ((size_t)b * a.c_feat + fc) * a.n + id)
                                      : __ldg(a.src1 + ((size_t)b * a.c1 + (fc - a.c_feat)) * a.n + id);
                }
            }
            bufA[(size_t)c * MLP_ROWS + row] = v;
        }
    }
    __syncthreads();

    float* in = bufA;
    float* outb = bufB;
    for (int l = 0; l < a.num_layers; ++l) {
        const int cin = a.ch[l], cout = a.ch[l + 1];
        const float* __restrict__ W = a.w[l];
        const float* __restrict__ Bv = a.bias[l];
        const bool last = (l == a.num_layers - 1);
        // each thread: its row, output channels co = cl*4 + 16*j + {0..3}
        for (int cb = cl * 4; cb < cout; cb += 16) {
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = (cb + j < cout) ? __ldg(Bv + cb + j) : 0.f;
            for (int k = 0; k < cin; ++k) {
                const float x = in[(size_t)k * MLP_ROWS + row];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cb + j < cout) acc[j] = fmaf(__ldg(W + (size_t)(cb + j) * cin + k), x, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (cb + j < cout) outb[(size_t)(cb + j) * MLP_ROWS + row] = fmaxf(acc[j], 0.f);
        }
        __syncthreads();
        if (last) {
            if (MLP_ROWS % S == 0) {
                // whole centres per tile: one thread reduces the S contiguous rows of (channel, centre)
                const int cpt = MLP_ROWS / S;
                for (int e = tid; e < cout * cpt; e += MLP_THREADS) {
                    const int co = e / cpt, ci = e - co * cpt;
                    const long long g2 = r0 + (long long)ci * S;
                    if (g2 >= a.total_rows) continue;
                    const float* src = outb + (size_t)co * MLP_ROWS + ci * S;
                    float mx = src[0];
                    for (int r2 = 1; r2 < S; ++r2) mx = fmaxf(mx, src[r2]);
                    const long long c2 = g2 / S;
                    const int b2 = (int)(c2 / M);
                    const int p2 = (int)(c2 - (long long)b2 * M);
                    a.out[((size_t)b2 * a.out_ctot + a.out_c0 + co) * M + p2] = mx;
                }
            } else {
                // centres straddle tiles: values are >= 0 after ReLU, so an unsigned max on the bit
                // pattern is an exact float max (the launcher zero-fills the destination)
                for (int e = tid; e < cout * MLP_ROWS; e += MLP_THREADS) {
                    const int co = e / MLP_ROWS, r = e - co * MLP_ROWS;
                    const long long g2 = r0 + r;
                    if (g2 >= a.total_rows) continue;
                    const long long c2 = g2 / S;
                    const int b2 = (int)(c2 / M);
                    const int p2 = (int)(c2 - (long long)b2 * M);
                    atomicMax(reinterpret_cast<unsigned int*>(a.out + ((size_t)b2 * a.out_ctot + a.out_c0 + co) * M + p2),
                              __float_as_uint(outb[(size_t)co * MLP_ROWS + r]));
                }
            }
        }
        float* t = in;
        in = outb;
        outb = t;
    }
}

}  // namespace tsm

int tsm_sa_mlp_fp32(const tsm::SaMlpArgs& a, int b, cudaStream_t stream) {
    int maxc = 0;
    for (int l = 0; l <= a.num_layers; ++l) maxc = a.ch[l] > maxc ? a.ch[l] : maxc;
    if (maxc > tsm::MLP_MAXC || maxc <= 0) return TSM_ERR_INVALID;
    const size_t dyn = (size_t)2 * maxc * tsm::MLP_ROWS * sizeof(float);
    if (dyn > 200 * 1024) return TSM_ERR_INVALID;
    if (dyn > 48 * 1024)
        TSM_CUDA_TRY(cudaFuncSetAttribute(tsm::sa_mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    const long long tiles = (a.total_rows + tsm::MLP_ROWS - 1) / tsm::MLP_ROWS;
    if (tiles > 0x7fffffffLL) return TSM_ERR_INVALID;
    // centres may straddle tiles when S does not divide MLP_ROWS: the kernel then combines
    // partial maxima with atomicMax, which needs a zeroed destination
    if (tsm::MLP_ROWS % a.s != 0) {
        for (int bi = 0; bi < b; ++bi)
            TSM_CUDA_TRY(cudaMemsetAsync(a.out + ((size_t)bi * a.out_ctot + a.out_c0) * a.m, 0,
                                         sizeof(float) * (size_t)a.ch[a.num_layers] * a.m, stream));
    }
    tsm::sa_mlp_fp32_kernel<<<(unsigned)tiles, tsm::MLP_THREADS, dyn, stream>>>(a, maxc);
    TSM_LAUNCH_CHECK();
    return TSM_OK;
}

// dense (point-wise) mode of the same kernel: a.idx == nullptr, a.s == 1, a.m == a.n (see tsmdet_pointwise_mlp)
int tsm_sa_mlp_fp32_dense(const tsm::SaMlpArgs& a, int b, cudaStream_t stream) {
    if (a.idx || a.s != 1 || a.m != a.n) return TSM_ERR_INVALID;
    return tsm_sa_mlp_fp32(a, b, stream);
}
