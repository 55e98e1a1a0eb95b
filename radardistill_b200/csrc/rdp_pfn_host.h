// rdp_pfn_host.h -- host-side table of the compiled PFN configurations.
#pragma once
#include "rdp_pfn.cuh"
#include "rdp_pfn_rows.cuh"
#include "rdp_pfn_bwd.cuh"

namespace rdp {

struct PfnLaunch {
    int cols, layout, dist, cout, cs;
    int stats_partial_doubles, bwd_partial_doubles;
    cudaError_t (*tile)(const PfnArgs &a, int mode, int grid, cudaStream_t st);
    cudaError_t (*rows)(const PfnArgs &a, int want_arg, int grid, cudaStream_t st);   // pfn_rows_kernel (rdp_pfn_rows.cuh)
    cudaError_t (*bwd_stream)(const PfnArgs &a, int grid, cudaStream_t st);            // pfn_bwd_stream_kernel (rdp_pfn_bwd.cuh)
    // one launch each: fixed-order reduction of the per-CTA partials, then (last CTA) the closed-form epilogue
    cudaError_t (*bn_finalize)(const PfnArgs &a, const double *partials, int nblocks, double *totals, int32_t *done, double *bn_state,
                               float *rm, float *rv, double momentum, long long *num_batches_tracked, cudaStream_t st);
    cudaError_t (*bwd_finalize)(const PfnArgs &a, const double *partials, int nblocks, double *totals, int32_t *done,
                                const double *bn_state, int train_bn, float *dW, float *dg, float *db, cudaStream_t st);
};

// (id, cols, layout, with_distance, c_out) -- one translation unit each (rdp_pfn_inst.cu, -DRDP_CFG_ID=id)
#define RDP_PFN_CONFIGS(X)                         \
    X(0, 6, RDP_LAYOUT_SIMPLE2D, false, 32)        \
    X(1, 7, RDP_LAYOUT_SIMPLE2D, false, 32)        \
    X(2, 5, RDP_LAYOUT_DYNPILLAR, false, 64)       \
    X(3, 5, RDP_LAYOUT_DYNPILLAR, true, 32)        \
    X(4, 6, RDP_LAYOUT_SIMPLE2D, true, 32)         \
    X(5, 5, RDP_LAYOUT_SIMPLE2D, false, 32)        \
    X(6, 5, RDP_LAYOUT_DYNPILLAR, false, 32)       \
    X(7, 6, RDP_LAYOUT_SIMPLE2D, false, 64)        \
    X(8, 7, RDP_LAYOUT_SIMPLE2D, false, 64)        \
    X(9, 6, RDP_LAYOUT_DYNPILLAR, false, 64)       \
    X(10, 6, RDP_LAYOUT_DYNPILLAR, false, 32)

#define RDP_DECLARE_CFG(id, cols, layout, dist, cout) const PfnLaunch *rdp_pfn_cfg_##id();
RDP_PFN_CONFIGS(RDP_DECLARE_CFG)
#undef RDP_DECLARE_CFG

}  // namespace rdp
