// sa_mlp.cu -- C entry points of the fused shared-MLP kernels (dispatch fp32 / tcgen05, per-call or pre-packed weights).
#include "sa_mlp.cuh"

namespace {

int fill_layers(tsm::SaMlpArgs& a, int num_layers, const int* channels, const float* const* weights,
                const float* const* biases) {
    if (num_layers < 1 || num_layers > 4 || !channels) return TSM_ERR_INVALID;
    a.num_layers = num_layers;
    for (int l = 0; l < 4; ++l) {
        a.w[l] = (weights && l < num_layers) ? weights[l] : nullptr;
        a.bias[l] = (biases && l < num_layers) ? biases[l] : nullptr;
    }
    for (int l = 0; l <= 4; ++l) a.ch[l] = l <= num_layers ? channels[l] : 0;
    return TSM_OK;
}

int sa_args(tsm::SaMlpArgs& a, int b, int n, int m, int nsample, int c_feat, int use_xyz, const float* xyz,
            const float* new_xyz, const float* features, const int* idx, const int* idx_cnt, int num_layers,
            const int* channels, const float* const* weights, const float* const* biases, float* out, int out_ctot,
            int out_c0) {
    if (nsample <= 0 || n <= 0) return TSM_ERR_INVALID;
    if (c_feat < 0 || (c_feat > 0 && !features) || (!use_xyz && c_feat == 0)) return TSM_ERR_INVALID;
    a = tsm::SaMlpArgs{};
    a.xyz = xyz;
    a.new_xyz = new_xyz;
    a.features = features;
    a.idx = idx;
    a.idx_cnt = idx_cnt;
    a.out = out;
    if (int rc = fill_layers(a, num_layers, channels, weights, biases)) return rc;
    if (a.ch[0] != (use_xyz ? 3 : 0) + c_feat) return TSM_ERR_INVALID;
    a.n = n;
    a.m = m;
    a.s = nsample;
    a.c_feat = c_feat;
    a.use_xyz = use_xyz;
    a.out_ctot = out_ctot;
    a.out_c0 = out_c0;
    a.total_rows = (long long)b * m * nsample;
    if (out_c0 < 0 || out_c0 + a.ch[num_layers] > out_ctot) return TSM_ERR_INVALID;
    return TSM_OK;
}

int dense_args(tsm::SaMlpArgs& a, int b, int n, int c0, int c1, const float* src0, const float* src1, int num_layers,
               const int* channels, const float* const* weights, const float* const* biases, float* out, int out_ctot,
               int out_c0) {
    if (c0 <= 0 || c1 < 0 || n <= 0) return TSM_ERR_INVALID;
    a = tsm::SaMlpArgs{};
    a.features = src0;
    a.src1 = c1 > 0 ? src1 : nullptr;
    a.c1 = c1;
    a.out = out;
    if (int rc = fill_layers(a, num_layers, channels, weights, biases)) return rc;
    if (a.ch[0] != c0 + c1) return TSM_ERR_INVALID;
    a.n = n;
    a.m = n;
    a.s = 1;
    a.c_feat = c0;
    a.out_ctot = out_ctot;
    a.out_c0 = out_c0;
    a.total_rows = (long long)b * n;
    if (out_c0 < 0 || out_c0 + a.ch[num_layers] > out_ctot) return TSM_ERR_INVALID;
    return TSM_OK;
}

}  // namespace

extern "C" int tsmdet_sa_mlp_maxpool(int b, int n, int m, int nsample, int c_feat, int use_xyz, const float* xyz,
                                     const float* new_xyz, const float* features, const int* idx, const int* idx_cnt,
                                     int num_layers, const int* channels, const float* const* weights,
                                     const float* const* biases, float* out, int out_ctot, int out_c0, int precision,
                                     void* stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (!weights || !biases) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    if (int rc = sa_args(a, b, n, m, nsample, c_feat, use_xyz, xyz, new_xyz, features, idx, idx_cnt, num_layers, channels,
                         weights, biases, out, out_ctot, out_c0))
        return rc;
    if (precision == 0) return tsm_sa_mlp_fp32(a, b, (cudaStream_t)stream);
    if (precision == 1) {
        // second-generation kernel (transposed last layer, in-register pooling) where it applies; TSMDET_MLP_V1=1
        // keeps every shape on the first-generation kernel (A/B measurements, tests of both)
        if (!tsm_knob(KNOB_MLP_V1)) {
            const int rc = tsm_mlp_tc2(a, b, 0, 2, (cudaStream_t)stream);
            if (rc != TSM_ERR_INVALID) return rc;
        }
        return tsm_sa_mlp_tc(a, b, (cudaStream_t)stream);
    }
    if (precision == 2) return tsm_mlp_tc2(a, b, 0, 4, (cudaStream_t)stream);  // tf32 operands; no other tf32 kernel
    return TSM_ERR_INVALID;
}

// Point-wise shared MLP over dense tensors: out[b, out_c0 + co, i] = MLP(cat(src0[b, :, i], src1[b, :, i])) with
// [1x1 conv (BN folded) + bias + ReLU] x num_layers -- PointnetFPModule.mlp (pointnet2_modules.py:175-176, on the
// concatenation of the interpolated and the skip features, :171) and aggregation_mlp (:1320-1321).
//   src0 (B,c0,n), src1 (B,c1,n) | NULL; channels[0] == c0 + c1; out (B,out_ctot,n).
// precision 0 = fp32 FMA, 1 = bf16 / 2 = tf32 tensor cores (tcgen05, fp32 accumulate).
extern "C" int tsmdet_pointwise_mlp(int b, int n, int c0, int c1, const float* src0, const float* src1, int num_layers,
                                    const int* channels, const float* const* weights, const float* const* biases,
                                    float* out, int out_ctot, int out_c0, int precision, void* stream) {
    if (b <= 0 || n <= 0) return TSM_OK;
    if (!src0 || (c1 > 0 && !src1) || !out || !weights || !biases) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    if (int rc = dense_args(a, b, n, c0, c1, src0, src1, num_layers, channels, weights, biases, out, out_ctot, out_c0)) return rc;
    if (precision == 0) return tsm_sa_mlp_fp32_dense(a, b, (cudaStream_t)stream);
    if (precision == 1 || precision == 2) return tsm_mlp_tc2(a, b, 1, precision == 1 ? 2 : 4, (cudaStream_t)stream);
    return TSM_ERR_INVALID;
}

// The tcgen05 path's weight image for an MLP whose weights do not change between calls (eval): build it once,
// pass it to the *_packed entry points.  dense = 0: fused SA scale (c1 ignored), 1: point-wise MLP (nsample /
// use_xyz ignored).  packed == NULL: only *packed_bytes is set.  TSMDET_ERR_INVALID: the tensor path does not take
// this shape -- use the unpacked entry points (they fall back to the first-generation / fp32 kernels).
extern "C" int tsmdet_mlp_pack_p(int precision, int dense, int nsample, int c_feat, int c1, int use_xyz, int num_layers,
                                 const int* channels, const float* const* weights, const float* const* biases, void* packed,
                                 long long* packed_bytes, void* stream) {
    if (precision != 1 && precision != 2) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    int rc;
    if (dense)
        rc = dense_args(a, 1, 128, c_feat, c1, nullptr, nullptr, num_layers, channels, weights, biases, nullptr, 1 << 30, 0);
    else
        rc = sa_args(a, 1, 128, 128, nsample, c_feat, use_xyz, nullptr, nullptr, (const float*)(c_feat > 0 ? (void*)8 : nullptr),
                     nullptr, nullptr, num_layers, channels, weights, biases, nullptr, 1 << 30, 0);
    if (rc) return rc;
    if (packed && (!weights || !biases)) return TSM_ERR_INVALID;
    return tsm_mlp_tc2_pack(a, dense, precision == 1 ? 2 : 4, static_cast<unsigned char*>(packed), packed_bytes, (cudaStream_t)stream);
}

extern "C" int tsmdet_mlp_pack(int dense, int nsample, int c_feat, int c1, int use_xyz, int num_layers, const int* channels,
                               const float* const* weights, const float* const* biases, void* packed,
                               long long* packed_bytes, void* stream) {
    return tsmdet_mlp_pack_p(1, dense, nsample, c_feat, c1, use_xyz, num_layers, channels, weights, biases, packed, packed_bytes, stream);
}

// features_t (optional): the features as (B,N,round_up(c_feat,8)) bf16 rows -- then `features` may be NULL and no
// transpose runs; out_t (optional): a second output, (B,M,round_up(cout,8)) bf16 rows, i.e. the next SA layer's
// features_t (stacked layers chain through it); out may be NULL when out_t is given.
extern "C" int tsmdet_sa_mlp_maxpool_packed_p(int precision, int b, int n, int m, int nsample, int c_feat, int use_xyz,
                                              const float* xyz, const float* new_xyz, const float* features,
                                              const void* features_t, const int* idx, const int* idx_cnt, int num_layers,
                                              const int* channels, const void* packed, float* out, void* out_t, int out_ctot,
                                              int out_c0, void* stream) {
    if (precision != 1 && precision != 2) return TSM_ERR_INVALID;
    if (b <= 0 || m <= 0) return TSM_OK;
    if (!packed || (!out && !out_t)) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    const float* f = features ? features : (const float*)features_t;  // sa_args only checks for non-null
    if (int rc = sa_args(a, b, n, m, nsample, c_feat, use_xyz, xyz, new_xyz, f, idx, idx_cnt, num_layers, channels,
                         nullptr, nullptr, out, out ? out_ctot : (1 << 30), out ? out_c0 : 0))
        return rc;
    a.features = features;
    a.feat_t = features_t;
    a.out_t = out_t;
    if (c_feat > 0 && !features && !features_t) return TSM_ERR_INVALID;
    return tsm_mlp_tc2(a, b, 0, precision == 1 ? 2 : 4, (cudaStream_t)stream, static_cast<const unsigned char*>(packed));
}

extern "C" int tsmdet_sa_mlp_maxpool_packed(int b, int n, int m, int nsample, int c_feat, int use_xyz, const float* xyz,
                                            const float* new_xyz, const float* features, const void* features_t,
                                            const int* idx, const int* idx_cnt, int num_layers, const int* channels,
                                            const void* packed, float* out, void* out_t, int out_ctot, int out_c0,
                                            void* stream) {
    return tsmdet_sa_mlp_maxpool_packed_p(1, b, n, m, nsample, c_feat, use_xyz, xyz, new_xyz, features, features_t, idx, idx_cnt,
                                          num_layers, channels, packed, out, out_t, out_ctot, out_c0, stream);
}

extern "C" int tsmdet_pointwise_mlp_packed_p(int precision, int b, int n, int c0, int c1, const float* src0, const float* src1,
                                             int num_layers, const int* channels, const void* packed, float* out,
                                             int out_ctot, int out_c0, void* stream) {
    if (precision != 1 && precision != 2) return TSM_ERR_INVALID;
    if (b <= 0 || n <= 0) return TSM_OK;
    if (!src0 || (c1 > 0 && !src1) || !out || !packed) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    if (int rc = dense_args(a, b, n, c0, c1, src0, src1, num_layers, channels, nullptr, nullptr, out, out_ctot, out_c0)) return rc;
    return tsm_mlp_tc2(a, b, 1, precision == 1 ? 2 : 4, (cudaStream_t)stream, static_cast<const unsigned char*>(packed));
}

extern "C" int tsmdet_pointwise_mlp_packed(int b, int n, int c0, int c1, const float* src0, const float* src1,
                                           int num_layers, const int* channels, const void* packed, float* out,
                                           int out_ctot, int out_c0, void* stream) {
    return tsmdet_pointwise_mlp_packed_p(1, b, n, c0, c1, src0, src1, num_layers, channels, packed, out, out_ctot, out_c0, stream);
}
