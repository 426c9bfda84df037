// sa_mlp.cu -- C entry point of the fused set-abstraction scale (dispatch fp32 / tcgen05).
#include "sa_mlp.cuh"

extern "C" int tsmdet_sa_mlp_maxpool(int b, int n, int m, int nsample, int c_feat, int use_xyz, const float* xyz,
                                     const float* new_xyz, const float* features, const int* idx, const int* idx_cnt,
                                     int num_layers, const int* channels, const float* const* weights,
                                     const float* const* biases, float* out, int out_ctot, int out_c0, int precision,
                                     void* stream) {
    if (b <= 0 || m <= 0) return TSM_OK;
    if (num_layers < 1 || num_layers > 4 || nsample <= 0 || n <= 0) return TSM_ERR_INVALID;
    if (c_feat < 0 || (c_feat > 0 && !features) || (!use_xyz && c_feat == 0)) return TSM_ERR_INVALID;
    tsm::SaMlpArgs a;
    a.xyz = xyz;
    a.new_xyz = new_xyz;
    a.features = features;
    a.idx = idx;
    a.idx_cnt = idx_cnt;
    a.out = out;
    a.num_layers = num_layers;
    for (int l = 0; l < 4; ++l) {
        a.w[l] = l < num_layers ? weights[l] : nullptr;
        a.bias[l] = l < num_layers ? biases[l] : nullptr;
    }
    for (int l = 0; l <= 4; ++l) a.ch[l] = l <= num_layers ? channels[l] : 0;
    if (a.ch[0] != (use_xyz ? 3 : 0) + c_feat) return TSM_ERR_INVALID;
    a.n = n;
    a.m = m;
    a.s = nsample;
    a.c_feat = c_feat;
    a.use_xyz = use_xyz;
    a.out_ctot = out_ctot;
    a.out_c0 = out_c0;
    a.total_rows = (long long)b * m * nsample;
    a.status = nullptr;
    if (out_c0 < 0 || out_c0 + a.ch[num_layers] > out_ctot) return TSM_ERR_INVALID;
    if (precision == 0) return tsm_sa_mlp_fp32(a, b, (cudaStream_t)stream);
    if (precision == 1) return tsm_sa_mlp_tc(a, b, (cudaStream_t)stream);
    return TSM_ERR_INVALID;
}
