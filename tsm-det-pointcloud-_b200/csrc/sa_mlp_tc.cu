// sa_mlp_tc.cu -- placeholder until the tcgen05 kernel lands in this file.
#include "sa_mlp.cuh"
int tsm_sa_mlp_tc(const tsm::SaMlpArgs&, int, cudaStream_t) { return TSM_ERR_INVALID; }
