// rdp_pfn_host.h -- host-side table of the compiled PFN configurations.
#pragma once
#include "rdp_index_host.h"
#include "rdp_pfn.cuh"

namespace rdp {

struct PfnLaunch {
    int cols, dist, cout, g;
    cudaError_t (*apply)(const PfnArgs &a, int want_arg, int grid, cudaStream_t st);
    cudaError_t (*bwd)(const PfnArgs &a, int grid, cudaStream_t st);
    // pillar table + train-mode feature moments (+ BN epilogue in the last CTA): replaces the generic table kernel
    cudaError_t (*table_stats)(const TableArgs &t, const PfnArgs &a, int64_t pcap, cudaStream_t st);
    cudaError_t (*bn_finalize)(const PfnArgs &a, cudaStream_t st);                       // SyncBatchNorm phase 2
    cudaError_t (*bwd_finalize)(const PfnArgs &a, const double *glob, cudaStream_t st);  // SyncBatchNorm backward phase 2
};

// (id, cols, with_distance, c_out) -- one translation unit each (rdp_pfn_inst.cu, -DRDP_CFG_ID=id).  cols = 1 + raw point
// features (4..8); the feature layout and the USE_* flags of model_cfg are runtime data (the matrix T), not instantiations.
#define RDP_PFN_CONFIGS(X)   \
    X(0, 4, false, 32)  X(1, 4, false, 64)  X(2, 4, false, 128)   \
    X(3, 5, false, 32)  X(4, 5, false, 64)  X(5, 5, false, 128)   \
    X(6, 6, false, 32)  X(7, 6, false, 64)  X(8, 6, false, 128)   \
    X(9, 7, false, 32)  X(10, 7, false, 64) X(11, 7, false, 128)  \
    X(12, 8, false, 32) X(13, 8, false, 64) X(14, 8, false, 128)  \
    X(15, 4, true, 32)  X(16, 4, true, 64)  X(17, 4, true, 128)   \
    X(18, 5, true, 32)  X(19, 5, true, 64)  X(20, 5, true, 128)   \
    X(21, 6, true, 32)  X(22, 6, true, 64)  X(23, 6, true, 128)   \
    X(24, 7, true, 32)  X(25, 7, true, 64)  X(26, 7, true, 128)   \
    X(27, 8, true, 32)  X(28, 8, true, 64)  X(29, 8, true, 128)

#define RDP_DECLARE_CFG(id, cols, dist, cout) const PfnLaunch *rdp_pfn_cfg_##id();
RDP_PFN_CONFIGS(RDP_DECLARE_CFG)
#undef RDP_DECLARE_CFG

}  // namespace rdp
