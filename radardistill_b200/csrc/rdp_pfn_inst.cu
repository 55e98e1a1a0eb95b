// rdp_pfn_inst.cu -- one PFN configuration per translation unit (compiled with -DRDP_CFG_ID=<id>).
#include "rdp_pfn_host.h"

#ifndef RDP_CFG_ID
#error "compile with -DRDP_CFG_ID=<id>"
#endif

namespace rdp {

#define RDP_PICK(id, cols, layout, dist, cout)                      \
    template <> struct CfgOf<id> { using type = PfnCfg<cols, layout, dist, cout>; };
template <int ID> struct CfgOf;
RDP_PFN_CONFIGS(RDP_PICK)
#undef RDP_PICK

using Cfg = CfgOf<RDP_CFG_ID>::type;

static cudaError_t fwd(const PfnArgs &a, int mode, int grid, cudaStream_t st) {
    const size_t smem = sizeof(PfnSmem<Cfg>);
    if (mode == PFN_MODE_STATS) {
        cudaError_t e = cudaFuncSetAttribute(pfn_fwd_kernel<Cfg, PFN_MODE_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pfn_fwd_kernel<Cfg, PFN_MODE_STATS><<<grid, kPfnThreads, smem, st>>>(a);
    } else {
        cudaError_t e = cudaFuncSetAttribute(pfn_fwd_kernel<Cfg, PFN_MODE_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pfn_fwd_kernel<Cfg, PFN_MODE_APPLY><<<grid, kPfnThreads, smem, st>>>(a);
    }
    return cudaGetLastError();
}

static cudaError_t bn_finalize(const PfnArgs &a, int nblocks, double *bn_state, float *rm, float *rv, double momentum, cudaStream_t st) {
    bn_finalize_kernel<Cfg><<<1, kPfnThreads, 0, st>>>(a, nblocks, bn_state, rm, rv, momentum);
    return cudaGetLastError();
}

constexpr size_t kBwdSmem = sizeof(double) * (kPfnThreads / 32) * Cfg::COUT * (Cfg::CS + 2);
constexpr size_t kBwdFinSmem = sizeof(double) * (Cfg::COUT * (Cfg::CS + 2) + Cfg::COUT);

static cudaError_t bwd(const PfnArgs &a, int grid, const float *grad, const float *feat, const int32_t *arg, const float *pmean,
                       cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(pfn_bwd_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return e;
    pfn_bwd_kernel<Cfg><<<grid, kPfnThreads, kBwdSmem, st>>>(a, grad, feat, arg, pmean);
    return cudaGetLastError();
}

static cudaError_t bwd_finalize(const PfnArgs &a, int nblocks, const double *bn_state, int train_bn, float *dW, float *dg, float *db,
                                cudaStream_t st) {
    bwd_finalize_kernel<Cfg><<<1, kPfnThreads, kBwdFinSmem, st>>>(a, nblocks, bn_state, train_bn, dW, dg, db);
    return cudaGetLastError();
}

#define RDP_CAT2(a, b) a##b
#define RDP_CAT(a, b) RDP_CAT2(a, b)
const PfnLaunch *RDP_CAT(rdp_pfn_cfg_, RDP_CFG_ID)() {
    static const PfnLaunch L = {Cfg::COLS, Cfg::LAYOUT, Cfg::DIST ? 1 : 0, Cfg::COUT, Cfg::CS,
                                sizeof(PfnSmem<Cfg>), kBwdSmem, kBwdFinSmem,
                                2 * Cfg::COUT + Cfg::CS * (Cfg::CS + 1) / 2 + Cfg::CS, Cfg::COUT * (Cfg::CS + 2),
                                fwd, bn_finalize, bwd, bwd_finalize};
    return &L;
}

}  // namespace rdp
