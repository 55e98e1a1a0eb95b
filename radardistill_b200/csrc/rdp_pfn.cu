// rdp_pfn.cu -- host side of rdp_pfn_fwd / rdp_pfn_bwd: argument marshalling and config dispatch.
#include <cstdlib>
#include <cstring>

#include "rdp_pfn_host.h"

#ifndef RDP_PFN_APPLY_PER_SM
#define RDP_PFN_APPLY_PER_SM 5
#endif

namespace rdp {

static const PfnLaunch *lookup(const rdp_geom_t *g, const rdp_layout_t *l) {
#define RDP_TRY_CFG(id, cols_, layout_, dist_, cout_)                                                          \
    if (g->cols == cols_ && l->layout == layout_ && (l->with_distance != 0) == dist_ && l->c_out == cout_) \
        return rdp_pfn_cfg_##id();
    RDP_PFN_CONFIGS(RDP_TRY_CFG)
#undef RDP_TRY_CFG
    return nullptr;
}

// super-feature -> W column map (see PfnCfg): the layout's concat order with switched-off options mapped to -1.
static int build_kmap(const rdp_geom_t *g, const rdp_layout_t *l, int8_t *kmap, int cs_expected) {
    const int C = g->cols - 1;
    int s = 0, k = 0;
    auto put = [&](bool used) { kmap[s++] = used ? (int8_t)k++ : (int8_t)-1; };
    for (int i = 0; i < kMaxSuper; ++i) kmap[i] = -1;
    if (l->layout == RDP_LAYOUT_SIMPLE2D) {
        for (int i = 0; i < 3; ++i) put(true);                               // f_center
        for (int c = 1; c <= C; ++c) put(l->use_abs || c >= 4);              // points[:, 1:] or points[:, 4:]
        for (int i = 0; i < 3; ++i) put(l->use_cluster != 0);                // f_cluster
        if (l->with_distance) put(true);
        for (int i = 0; i < 3; ++i) put(l->use_relative != 0);               // f_relative
    } else {
        for (int c = 1; c <= C; ++c) put(l->use_abs || c >= 4);
        for (int i = 0; i < 3; ++i) put(true);                               // f_cluster
        for (int i = 0; i < 3; ++i) put(true);                               // f_center
        if (l->with_distance) put(true);
    }
    if (s != cs_expected || k != l->c_in) return RDP_ERR_INVALID_ARG;
    return RDP_OK;
}

static int fill_args(PfnArgs *a, const PfnLaunch *L, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                     const rdp_pfn_params_t *prm, const Workspace &ws, const int32_t *counters) {
    memset(a, 0, sizeof(*a));
    a->grows = ws.grows; a->aux = ws.aux; a->tile_first = ws.tile_first; a->ends = ws.ends; a->counters = counters;
    a->orig2kept = ws.orig2kept;
    a->weight = prm->weight; a->bias = prm->bias; a->gamma = prm->gamma; a->beta = prm->beta;
    a->rmean = prm->running_mean; a->rvar = prm->running_var;
    a->partials = ws.partials;
    a->n0 = n_points;
    a->eps = prm->eps;
    for (int i = 0; i < 3; ++i) { a->lo[i] = geom->lo[i]; a->vsz[i] = geom->vsz[i]; a->off[i] = geom->off[i]; }
    a->c_in = layout->c_in;
    a->use_norm = prm->gamma != nullptr;
    return build_kmap(geom, layout, a->kmap, L->cs);
}

// argmax in the reference's numbering: index of the winning row among the KEPT points (dynamic_pillar_vfe.py:204-206).
__global__ void argpos_to_kept_kernel(const int32_t *__restrict__ argpos, const float *__restrict__ grows, int rs,
                                      const int32_t *__restrict__ orig2kept, const int32_t *__restrict__ counters, long long n0,
                                      int cout, int32_t *__restrict__ out) {
    const long long total = (long long)counters[RDP_CNT_P] * cout;
    const bool none_dropped = ((long long)counters[RDP_CNT_N] == n0);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ap = argpos[e];
        const int row = __float_as_int(grows[((size_t)(ap >= 0 ? ap : ~ap) + 1) * rs + rs - 2]);
        out[e] = none_dropped ? row : orig2kept[row];
    }
}

// scatter_mean's (P, 3) output (:226) for callers that want it: the first three columns of the pillar table
__global__ void table_to_mean_kernel(const float *__restrict__ aux, const int32_t *__restrict__ counters, float *__restrict__ out) {
    const long long total = 3ll * counters[RDP_CNT_P];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        out[e] = aux[(e / 3) * 8 + (e % 3)];
}

static int pfn_grid(int64_t n_points) {
    const int64_t tiles = (n_points + kPfnWin - 1) / kPfnWin;
    return (int)(tiles < kPfnGridCap ? (tiles < 1 ? 1 : tiles) : kPfnGridCap);
}

}  // namespace rdp

using namespace rdp;

extern "C" int rdp_pfn_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                           const rdp_pfn_params_t *prm, void *workspace, size_t workspace_bytes, const int32_t *counters,
                           float *features, int32_t *argpos, float *pillar_mean, double *bn_state, void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!geom || !layout || !prm || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!points || !workspace || !features || !prm->weight) return RDP_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(features) & 15u) || (argpos && (reinterpret_cast<uintptr_t>(argpos) & 15u))) return RDP_ERR_INVALID_ARG;
    const bool use_norm = prm->gamma != nullptr;
    if (use_norm && (!prm->beta || !prm->running_mean || !prm->running_var)) return RDP_ERR_INVALID_ARG;
    const bool train = use_norm && prm->train_bn;
    if (train && !bn_state) return RDP_ERR_INVALID_ARG;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, L, n_points, geom, layout, prm, ws, counters);
    if (rc != RDP_OK) return rc;
    a.features = features;
    a.argpos = argpos;
    if (pillar_mean) table_to_mean_kernel<<<148 * 4, 256, 0, st>>>(ws.aux, counters, pillar_mean);
    const int grid = pfn_grid(n_points);
    if (train) {
        if (L->stats_partial_doubles > ws.partial_doubles_per_block) return RDP_ERR_WORKSPACE;
        RDP_CUDA_OK(L->tile(a, PFN_MODE_STATS, grid, st));
        RDP_CUDA_OK(L->bn_finalize(a, ws.partials, grid, ws.totals, const_cast<int32_t *>(counters) + kCntDoneStats, bn_state,
                                   prm->running_mean, prm->running_var, prm->momentum,
                                   reinterpret_cast<long long *>(prm->num_batches_tracked), st));
        a.bn_state = bn_state;
        a.fold_from_state = 1;
    }
    // forward form: the lane = channel tile stream (default) or the thread = row kernel (RDP_PFN_ROWS=1; bit-identical,
    // measured slower so far -- see DESIGN.md); the row kernel's 256-bit stores need 32-byte aligned outputs
    static const bool use_rows = getenv("RDP_PFN_ROWS") != nullptr;
    const bool aligned32 = !(reinterpret_cast<uintptr_t>(features) & 31u) && !(reinterpret_cast<uintptr_t>(argpos) & 31u);
    if (!use_rows || !aligned32) {
        // the eval kernel is compiled for 5 resident CTAs per SM (no partial-sum slots involved): let it have them
        const int64_t tiles_ = (n_points + kPfnWin - 1) / kPfnWin;
        const int agrid = argpos ? grid : (int)(tiles_ < 148 * RDP_PFN_APPLY_PER_SM ? (tiles_ < 1 ? 1 : tiles_) : 148 * RDP_PFN_APPLY_PER_SM);
        RDP_CUDA_OK(L->tile(a, argpos ? PFN_MODE_APPLY_ARG : PFN_MODE_APPLY, agrid, st));
    } else {
        const int64_t tiles = (n_points + kPfnWin - 1) / kPfnWin;
        const int rgrid = (int)(tiles < kRowsGridCap ? (tiles < 1 ? 1 : tiles) : kRowsGridCap);
        RDP_CUDA_OK(L->rows(a, argpos ? 1 : 0, rgrid, st));
    }
    return RDP_OK;
}

extern "C" int rdp_pfn_bwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                           const rdp_pfn_params_t *prm, void *workspace, size_t workspace_bytes, const int32_t *counters,
                           const float *grad_features, const float *features, const int32_t *argpos, const double *bn_state,
                           float *d_weight, float *d_gamma, float *d_beta, void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!geom || !layout || !prm || !counters || !d_weight || !d_beta || n_points < 0) return RDP_ERR_INVALID_ARG;
    const bool use_norm = prm->gamma != nullptr;
    const bool train = use_norm && prm->train_bn;
    if (n_points == 0) {
        RDP_CUDA_OK(cudaMemsetAsync(d_weight, 0, sizeof(float) * layout->c_out * layout->c_in, st));
        RDP_CUDA_OK(cudaMemsetAsync(d_beta, 0, sizeof(float) * layout->c_out, st));
        if (d_gamma) RDP_CUDA_OK(cudaMemsetAsync(d_gamma, 0, sizeof(float) * layout->c_out, st));
        return RDP_OK;
    }
    (void)features;   // the ReLU mask travels in the sign of argpos; kept in the signature for ABI stability
    if (!points || !workspace || !grad_features || !argpos) return RDP_ERR_INVALID_ARG;
    // both are sources of 16-byte bulk (TMA) copies
    if ((reinterpret_cast<uintptr_t>(grad_features) & 15u) || (reinterpret_cast<uintptr_t>(argpos) & 15u)) return RDP_ERR_INVALID_ARG;
    if (train && !bn_state) return RDP_ERR_INVALID_ARG;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    if (L->bwd_partial_doubles > ws.partial_doubles_per_block) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, L, n_points, geom, layout, prm, ws, counters);
    if (rc != RDP_OK) return rc;
    a.grad = grad_features;
    a.argpos = const_cast<int32_t *>(argpos);
    // backward form: the tile kernel (default) or the pillar-streaming kernel (RDP_BWD_STREAM=1; same results, measured
    // slower so far -- see DESIGN.md)
    static const bool use_stream = getenv("RDP_BWD_STREAM") != nullptr;
    int grid = pfn_grid(n_points);
    if (!use_stream) {
        RDP_CUDA_OK(L->tile(a, PFN_MODE_BWD, grid, st));
    } else {
        grid = kBwdGrid;   // persistent: every warp streams a contiguous pillar range (P is only known on the device)
        RDP_CUDA_OK(L->bwd_stream(a, grid, st));
    }
    RDP_CUDA_OK(L->bwd_finalize(a, ws.partials, grid, ws.totals, const_cast<int32_t *>(counters) + kCntDoneBwd, bn_state,
                                train ? 1 : 0, d_weight, d_gamma, d_beta, st));
    return RDP_OK;
}

extern "C" int rdp_argmax_kept(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, void *workspace,
                               size_t workspace_bytes, const int32_t *counters, const int32_t *argpos, int32_t *argmax_kept,
                               void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!geom || !layout || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!workspace || !argpos || !argmax_kept) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    argpos_to_kept_kernel<<<148 * 8, 256, 0, st>>>(argpos, ws.grows, grouped_row_floats(geom->cols), ws.orig2kept, counters, n_points,
                                                    layout->c_out, argmax_kept);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}
