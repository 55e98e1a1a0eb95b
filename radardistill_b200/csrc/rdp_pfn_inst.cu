// rdp_pfn_inst.cu -- one PFN configuration per translation unit (compiled with -DRDP_CFG_ID=<id>).
#include "rdp_pfn_host.h"

#ifndef RDP_CFG_ID
#error "compile with -DRDP_CFG_ID=<id>"
#endif

namespace rdp {

template <int ID> struct CfgOf;
#define RDP_PICK(id, cols, dist, cout) \
    template <> struct CfgOf<id> { using type = PfnCfg<cols, dist, cout>; };
RDP_PFN_CONFIGS(RDP_PICK)
#undef RDP_PICK

using Cfg = CfgOf<RDP_CFG_ID>::type;

// the dynamic shared-memory attribute is sticky per device: set it once
template <class K>
static cudaError_t ensure_smem(K kernel, size_t smem, bool *configured) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    return cudaSuccess;
}

template <bool ARG>
static cudaError_t launch_apply(const PfnArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem<Cfg, 0>);
    static bool configured[64] = {false};
    cudaError_t e = ensure_smem(pfn_apply_kernel<Cfg, ARG>, smem, configured);
    if (e != cudaSuccess) return e;
    return launch_pdl(pfn_apply_kernel<Cfg, ARG>, grid, kPfnThreads, smem, st, a);
}

static cudaError_t apply(const PfnArgs &a, int want_arg, int grid, cudaStream_t st) {
    return want_arg ? launch_apply<true>(a, grid, st) : launch_apply<false>(a, grid, st);
}

static cudaError_t bwd(const PfnArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem<Cfg, 24>);
    static bool configured[64] = {false};
    cudaError_t e = ensure_smem(pfn_bwd_kernel<Cfg>, smem, configured);
    if (e != cudaSuccess) return e;
    return launch_pdl(pfn_bwd_kernel<Cfg>, grid, kPfnThreads, smem, st, a);
}

static cudaError_t table_stats(const TableArgs &t, const PfnArgs &a, int64_t pcap, cudaStream_t st) {
    const int64_t blocks = (pcap + kTableStatsThreads - 1) / kTableStatsThreads, cap = 148 * 8;
    return launch_pdl(pillar_table_stats_kernel<Cfg>, (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), kTableStatsThreads, 0, st, t, a);
}

static cudaError_t bn_finalize(const PfnArgs &a, cudaStream_t st) {
    bn_finalize_kernel<Cfg><<<1, 256, 0, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t bwd_finalize(const PfnArgs &a, const double *glob, cudaStream_t st) {
    const size_t smem = sizeof(double) * Cfg::COUT * Cfg::BWD_PER;
    bwd_finalize_kernel<Cfg><<<1, 256, smem, st>>>(a, glob);
    return cudaGetLastError();
}

#define RDP_CAT2(a, b) a##b
#define RDP_CAT(a, b) RDP_CAT2(a, b)
const PfnLaunch *RDP_CAT(rdp_pfn_cfg_, RDP_CFG_ID)() {
    static const PfnLaunch L = {Cfg::COLS, Cfg::DIST ? 1 : 0, Cfg::COUT, Cfg::G, apply, bwd, table_stats, bn_finalize, bwd_finalize};
    return &L;
}

}  // namespace rdp
