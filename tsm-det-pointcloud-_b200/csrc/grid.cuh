// grid.cuh -- the per-cloud uniform grid shared by the grid ball query and the grid 3-NN (built by
// bq_grid_build_kernel in ball_query_grid.cu).
#pragma once
#include "common.cuh"

namespace tsm {

constexpr int kGridCells = 32768;   // shared-memory histogram: 128 KB
constexpr int kMaxCell = 256;       // largest cell population the index fix-up accepts
constexpr int kHdrInts = 16;        // per-cloud header: lo[3], inv[3] (float bits), n[3], ok

struct GridHdr {
    float lo[3];
    float inv[3];
    int n[3];
    int ok;
};

__device__ __forceinline__ int grid_q(float v, float lo, float inv, int n) {
    return min(max(__float2int_rd(__fmul_rn(__fsub_rn(v, lo), inv)), 0), n - 1);
}

}  // namespace tsm

// Builds the grid of every cloud on `stream` (scratch tag `tag`): cells at least 1.01 * rout wide, or -- rout < 0 --
// sized for about two points per cell.  Returns device pointers into the scratch buffer.
int tsm_grid_build(int b, int n, float rout, const float* xyz, cudaStream_t stream, int tag, const int** hdr,
                   const int** cell_start, const float4** sorted);
