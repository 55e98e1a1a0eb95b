// rdp_pfn_inst.cu -- one PFN configuration per translation unit (compiled with -DRDP_CFG_ID=<id>).
#include "rdp_pfn_host.h"

#ifndef RDP_CFG_ID
#error "compile with -DRDP_CFG_ID=<id>"
#endif

namespace rdp {

template <int ID> struct CfgOf;
#define RDP_PICK(id, cols, layout, dist, cout) \
    template <> struct CfgOf<id> { using type = PfnCfg<cols, layout, dist, cout>; };
RDP_PFN_CONFIGS(RDP_PICK)
#undef RDP_PICK

using Cfg = CfgOf<RDP_CFG_ID>::type;

template <int MODE>
static cudaError_t launch_tile(const PfnArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(PfnSmem<Cfg, MODE>);
    static bool configured[64] = {false};  // per device; the attribute is sticky, set it once
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(pfn_tile_kernel<Cfg, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    pfn_tile_kernel<Cfg, MODE><<<grid, kPfnThreads, smem, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t tile(const PfnArgs &a, int mode, int grid, cudaStream_t st) {
    if (mode == PFN_MODE_STATS) return launch_tile<PFN_MODE_STATS>(a, grid, st);
    if (mode == PFN_MODE_BWD) return launch_tile<PFN_MODE_BWD>(a, grid, st);
    if (mode == PFN_MODE_APPLY_ARG) return launch_tile<PFN_MODE_APPLY_ARG>(a, grid, st);
    return launch_tile<PFN_MODE_APPLY>(a, grid, st);
}

template <bool ARG>
static cudaError_t launch_rows(const PfnArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(RowsSmem<Cfg, ARG>);
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(pfn_rows_kernel<Cfg, ARG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    pfn_rows_kernel<Cfg, ARG><<<grid, kRowsThreads, smem, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t rows(const PfnArgs &a, int want_arg, int grid, cudaStream_t st) {
    return want_arg ? launch_rows<true>(a, grid, st) : launch_rows<false>(a, grid, st);
}

static cudaError_t bwd_stream(const PfnArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(double) * (kBwdThreads / 32) * Cfg::BWD_DOUBLES;
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(pfn_bwd_stream_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    pfn_bwd_stream_kernel<Cfg><<<grid, kBwdThreads, smem, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t bn_finalize(const PfnArgs &a, const double *partials, int nblocks, double *totals, int32_t *done, double *bn_state,
                               float *rm, float *rv, double momentum, long long *num_batches_tracked, cudaStream_t st) {
    bn_finalize_kernel<Cfg><<<(Cfg::STATS_DOUBLES + 31) / 32, 256, 0, st>>>(a, partials, nblocks, totals, done, bn_state, rm, rv,
                                                                            momentum, num_batches_tracked);
    return cudaGetLastError();
}

constexpr size_t kBwdFinSmem = sizeof(double) * (Cfg::BWD_DOUBLES + Cfg::COUT);

static cudaError_t bwd_finalize(const PfnArgs &a, const double *partials, int nblocks, double *totals, int32_t *done,
                                const double *bn_state, int train_bn, float *dW, float *dg, float *db, cudaStream_t st) {
    bwd_finalize_kernel<Cfg><<<(Cfg::BWD_DOUBLES + 31) / 32, 256, kBwdFinSmem, st>>>(a, partials, nblocks, totals, done, bn_state,
                                                                                     train_bn, dW, dg, db);
    return cudaGetLastError();
}

#define RDP_CAT2(a, b) a##b
#define RDP_CAT(a, b) RDP_CAT2(a, b)
const PfnLaunch *RDP_CAT(rdp_pfn_cfg_, RDP_CFG_ID)() {
    static const PfnLaunch L = {Cfg::COLS, Cfg::LAYOUT, Cfg::DIST ? 1 : 0, Cfg::COUT, Cfg::CS,
                                Cfg::STATS_DOUBLES, Cfg::BWD_DOUBLES, tile, rows, bwd_stream, bn_finalize, bwd_finalize};
    return &L;
}

}  // namespace rdp
