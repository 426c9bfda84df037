// sa_mlp.cuh -- argument block shared by the fused set-abstraction kernels.
#pragma once
#include "common.cuh"

namespace tsm {

struct SaMlpArgs {
    const float* xyz;       // (B,N,3)
    const float* new_xyz;   // (B,M,3)
    const float* features;  // (B,C,N) or null.  Dense (point-wise) mode: source 0, (B,c_feat,n)
    const float* src1;      // dense mode only: source 1, (B,c1,n), concatenated after source 0 along channels; or null
    const int* idx;         // (B,M,S)
    const int* idx_cnt;     // (B,M) or null (no masking)
    float* out;             // (B,out_ctot,M), or null when only out_t is wanted (tensor path)
    const void* feat_t;     // optional: the features as (B,N,round_up(c_feat,8)) bf16 rows (a previous layer's out_t)
    void* out_t;            // optional second output: (B,M,round_up(cout,8)) bf16 rows, zero padded -- the next layer's feat_t
    const float* w[4];
    const float* bias[4];
    int ch[5];  // ch[0] = input channels (3*use_xyz + C), ch[l+1] = outputs of layer l
    int num_layers;
    int n, m, s, c_feat, use_xyz;
    int c1;                 // channels of src1 (0 outside the dense mode)
    int out_ctot, out_c0;
    long long total_rows;  // B*M*S
    int* status;           // watchdog word (device) or nullptr
};


}  // namespace tsm

int tsm_sa_mlp_fp32(const tsm::SaMlpArgs& a, int b, cudaStream_t stream);
int tsm_sa_mlp_fp32_dense(const tsm::SaMlpArgs& a, int b, cudaStream_t stream);
int tsm_sa_mlp_tc(const tsm::SaMlpArgs& a, int b, cudaStream_t stream);  // TSM_ERR_INVALID if the shape is unsupported
// second-generation tcgen05 kernel (mlp_tc2.cu): SA scales with nsample a power of two in 8..128 (dense == 0),
// point-wise MLPs over dense (B,C,n) inputs (dense != 0); eb = bytes per operand element, 2 = bf16, 4 = tf32.
// TSM_ERR_INVALID if the shape is unsupported (incl. weights that do not fit shared memory in tf32).
int tsm_mlp_tc2(const tsm::SaMlpArgs& a, int b, int dense, int eb, cudaStream_t stream, const unsigned char* prepacked = nullptr);
int tsm_mlp_tc2_pack(const tsm::SaMlpArgs& a, int dense, int eb, unsigned char* packed, long long* bytes, cudaStream_t stream);
