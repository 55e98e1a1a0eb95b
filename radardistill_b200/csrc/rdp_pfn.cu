// rdp_pfn.cu -- host side of rdp_pfn_fwd / rdp_pfn_bwd: argument marshalling and config dispatch.
#include <cstring>

#include "rdp_pfn_host.h"

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

static int fill_args(PfnArgs *a, const PfnLaunch *L, const float *points, int64_t n_points, const rdp_geom_t *geom,
                     const rdp_layout_t *layout, const rdp_pfn_params_t *prm, const Workspace &ws, const int32_t *counters,
                     const int32_t *coords) {
    memset(a, 0, sizeof(*a));
    a->pts = points;
    a->order = ws.order; a->ends = ws.ends; a->tile_start = ws.tile_start; a->counters = counters; a->coords = coords;
    a->orig2kept = ws.orig2kept; a->kept2orig = ws.kept2orig;
    a->weight = prm->weight; a->bias = prm->bias; a->gamma = prm->gamma; a->beta = prm->beta;
    a->rmean = prm->running_mean; a->rvar = prm->running_var;
    a->partials = ws.partials;
    a->n0 = n_points;
    a->eps = prm->eps;
    for (int i = 0; i < 3; ++i) { a->lo[i] = geom->lo[i]; a->vsz[i] = geom->vsz[i]; a->off[i] = geom->off[i]; }
    a->c_in = layout->c_in;
    a->coord_cols = layout->coord_cols;
    a->use_norm = prm->gamma != nullptr;
    return build_kmap(geom, layout, a->kmap, L->cs);
}

}  // namespace rdp

using namespace rdp;

extern "C" int rdp_pfn_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                              const rdp_pfn_params_t *prm, void *workspace, size_t workspace_bytes, const int32_t *counters,
                              const int32_t *coords, float *features, int32_t *argmax, float *pillar_mean, double *bn_state,
                              void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!geom || !layout || !prm || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!points || !workspace || !coords || !features || !prm->weight) return RDP_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(features) & 15u) || (argmax && (reinterpret_cast<uintptr_t>(argmax) & 15u))) return RDP_ERR_INVALID_ARG;
    const bool use_norm = prm->gamma != nullptr;
    if (use_norm && (!prm->beta || !prm->running_mean || !prm->running_var)) return RDP_ERR_INVALID_ARG;
    const bool train = use_norm && prm->train_bn;
    if (train && !bn_state) return RDP_ERR_INVALID_ARG;
    if (layout->coord_cols != 3 && layout->coord_cols != 4) return RDP_ERR_INVALID_ARG;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, L, points, n_points, geom, layout, prm, ws, counters, coords);
    if (rc != RDP_OK) return rc;
    a.features = features;
    a.argmax = argmax;
    a.pillar_mean = pillar_mean;
    const int tiles = (int)ws.pfn_tiles;
    if (train) {
        if (L->stats_partial_doubles > ws.partial_doubles_per_block) return RDP_ERR_WORKSPACE;
        const int gs = tiles < ws.partial_blocks ? tiles : ws.partial_blocks;
        RDP_CUDA_OK(L->fwd(a, PFN_MODE_STATS, gs, st));
        RDP_CUDA_OK(L->bn_finalize(a, gs, bn_state, prm->running_mean, prm->running_var, prm->momentum, st));
        a.bn_state = bn_state;
        a.fold_from_state = 1;
    }
    const int cap = 148 * 8;
    RDP_CUDA_OK(L->fwd(a, PFN_MODE_APPLY, tiles < cap ? tiles : cap, st));
    return RDP_OK;
}

extern "C" int rdp_pfn_bwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                              const rdp_pfn_params_t *prm, void *workspace, size_t workspace_bytes, const int32_t *counters,
                              const int32_t *coords, const float *grad_features, const float *features, const int32_t *argmax,
                              const float *pillar_mean, const double *bn_state, float *d_weight, float *d_gamma, float *d_beta,
                              int64_t n_pillars_hint, void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!geom || !layout || !prm || !counters || !d_weight || !d_beta || n_points < 0) return RDP_ERR_INVALID_ARG;
    const bool use_norm = prm->gamma != nullptr;
    const bool train = use_norm && prm->train_bn;
    RDP_CUDA_OK(cudaMemsetAsync(d_weight, 0, sizeof(float) * layout->c_out * layout->c_in, st));
    RDP_CUDA_OK(cudaMemsetAsync(d_beta, 0, sizeof(float) * layout->c_out, st));
    if (d_gamma) RDP_CUDA_OK(cudaMemsetAsync(d_gamma, 0, sizeof(float) * layout->c_out, st));
    if (n_points == 0 || n_pillars_hint == 0) return RDP_OK;
    if (!points || !workspace || !coords || !grad_features || !features || !argmax || !pillar_mean) return RDP_ERR_INVALID_ARG;
    if (train && !bn_state) return RDP_ERR_INVALID_ARG;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    if (L->bwd_partial_doubles > ws.partial_doubles_per_block) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, L, points, n_points, geom, layout, prm, ws, counters, coords);
    if (rc != RDP_OK) return rc;
    int64_t want = n_pillars_hint > 0 ? (n_pillars_hint + 7) / 8 : ws.partial_blocks;
    const int grid = (int)(want < ws.partial_blocks ? (want < 1 ? 1 : want) : ws.partial_blocks);
    RDP_CUDA_OK(L->bwd(a, grid, grad_features, features, argmax, pillar_mean, st));
    RDP_CUDA_OK(L->bwd_finalize(a, grid, bn_state, train ? 1 : 0, d_weight, d_gamma, d_beta, st));
    return RDP_OK;
}
