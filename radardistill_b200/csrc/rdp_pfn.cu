// rdp_pfn.cu -- host side of rdp_pfn_fwd / rdp_pfn_bwd / rdp_encode_fwd: argument marshalling and config dispatch.
#include <cstdlib>
#include <cstring>

#include "rdp_pfn_host.h"

namespace rdp {

static const PfnLaunch *lookup(const rdp_geom_t *g, const rdp_layout_t *l) {
#define RDP_TRY_CFG(id, cols_, dist_, cout_) \
    if (g->cols == cols_ && (l->with_distance != 0) == dist_ && l->c_out == cout_) return rdp_pfn_cfg_##id();
    RDP_PFN_CONFIGS(RDP_TRY_CFG)
#undef RDP_TRY_CFG
    return nullptr;
}

// T: the layout's features in the reduced basis g = [dx, dy, dz, raw features 4.., (dist) | cx, cy, cx-mx, cy-my, cz-mz | 1]
// (rdp_pfn.cuh).  Feature order = the reference's concat order (dynamic_pillar_vfe.py:219-237 / :113-121).
static int build_T(const rdp_geom_t *g, const rdp_layout_t *l, float T[kMaxCin][kMaxG + 1]) {
    const int C = g->cols - 1, KIN = C + (l->with_distance ? 1 : 0), G = KIN + 5;
    if (G > kMaxG) return RDP_ERR_UNSUPPORTED;
    for (int j = 0; j < kMaxCin; ++j)
        for (int m = 0; m <= kMaxG; ++m) T[j][m] = 0.0f;
    int j = 0;
    bool overflow = false;
    auto next = [&]() -> float * {
        if (j >= kMaxCin) { overflow = true; return T[kMaxCin - 1]; }
        return T[j++];
    };
    const float lo[3] = {g->lo[0], g->lo[1], g->lo[2]};
    const float off_z = g->off[2];
    auto center = [&]() { for (int a = 0; a < 3; ++a) next()[a] = 1.0f; };                       // f_center = d
    auto cluster = [&]() { for (int a = 0; a < 3; ++a) { float *r = next(); r[a] = 1.0f; r[KIN + 2 + a] = 1.0f; } };   // d + (centre - mean)
    auto points = [&]() {
        if (l->use_abs) {   // x = dx + cx, y = dy + cy, z = dz + z_offset
            float *r = next(); r[0] = 1.0f; r[KIN] = 1.0f;
            r = next(); r[1] = 1.0f; r[KIN + 1] = 1.0f;
            r = next(); r[2] = 1.0f; r[G] = off_z;
        }
        for (int c = 4; c <= C; ++c) next()[c - 1] = 1.0f;
    };
    auto dist = [&]() { if (l->with_distance) next()[C] = 1.0f; };
    auto relative = [&]() {   // xyz - lo = d + (centre - lo)
        float *r = next(); r[0] = 1.0f; r[KIN] = 1.0f; r[G] = -lo[0];
        r = next(); r[1] = 1.0f; r[KIN + 1] = 1.0f; r[G] = -lo[1];
        r = next(); r[2] = 1.0f; r[G] = (float)((double)off_z - (double)lo[2]);
    };
    if (l->layout == RDP_LAYOUT_SIMPLE2D) {
        center();
        points();
        if (l->use_cluster) cluster();
        dist();
        if (l->use_relative) relative();
    } else if (l->layout == RDP_LAYOUT_DYNPILLAR) {
        points();
        cluster();
        center();
        dist();
    } else {
        return RDP_ERR_UNSUPPORTED;
    }
    if (overflow || j != l->c_in) return RDP_ERR_INVALID_ARG;
    return RDP_OK;
}

static int fill_args(PfnArgs *a, const rdp_geom_t *geom, const rdp_layout_t *layout, const rdp_pfn_params_t *prm, const Workspace &ws,
                     int64_t n_points, int32_t *counters) {
    memset(a, 0, sizeof(*a));
    a->grows = ws.grows; a->aux = ws.aux; a->tile_first = ws.tile_first; a->starts = ws.starts; a->counters = counters;
    a->orig2kept = ws.orig2kept;
    a->weight = prm->weight; a->bias = prm->bias; a->gamma = prm->gamma; a->beta = prm->beta;
    a->rmean = prm->running_mean; a->rvar = prm->running_var;
    a->acc_stats = ws.acc_stats; a->acc_bwd = ws.acc_bwd;
    a->num_batches_tracked = reinterpret_cast<long long *>(prm->num_batches_tracked);
    a->n0 = n_points;
    a->eps = prm->eps; a->momentum = prm->momentum;
    a->off_z = geom->off[2];
    a->c_in = layout->c_in;
    a->use_norm = prm->gamma != nullptr;
    a->train_bn = (a->use_norm && prm->train_bn) ? 1 : 0;
    return build_T(geom, layout, a->T);
}

// argmax in the reference's numbering: index of the winning row among the KEPT points (dynamic_pillar_vfe.py:204-206).  A
// clamped (pillar, channel) reports the pillar's lowest kept index (every row ties at 0): thread = pillar scans its rows once.
__global__ void argpos_to_kept_kernel(const int32_t *__restrict__ argpos, const float *__restrict__ grows, int rs,
                                      const int32_t *__restrict__ starts, const int32_t *__restrict__ orig2kept,
                                      const int32_t *__restrict__ counters, long long n0, int cout, int32_t *__restrict__ out) {
    const int P = counters[RDP_CNT_P];
    const bool none_dropped = ((long long)counters[RDP_CNT_N] == n0);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        int lowest = -1;
        for (int c = 0; c < cout; ++c) {
            const int ap = argpos[(size_t)p * cout + c];
            int row;
            if (ap >= 0) {
                row = __float_as_int(grows[((size_t)ap + 1) * rs + rs - 2]);
            } else {
                if (lowest < 0) {
                    lowest = 0x7fffffff;
                    for (int i = starts[p]; i < starts[p + 1]; ++i) lowest = min(lowest, __float_as_int(grows[((size_t)i + 1) * rs + rs - 2]));
                }
                row = lowest;
            }
            out[(size_t)p * cout + c] = none_dropped ? row : orig2kept[row];
        }
    }
}

// scatter_mean's (P, 3) output (:226) for callers that want it: thread = pillar re-sums its rows (fp64: exact, one rounding)
__global__ void pillar_mean_kernel(const float *__restrict__ grows, int rs, const int32_t *__restrict__ starts,
                                   const int32_t *__restrict__ counters, float *__restrict__ out) {
    const int P = counters[RDP_CNT_P];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        const int s = starts[p], e = starts[p + 1];
        for (int i = s; i < e; ++i) {
            const float4 v = *reinterpret_cast<const float4 *>(grows + ((size_t)i + 1) * rs);
            sx += (double)v.y; sy += (double)v.z; sz += (double)v.w;
        }
        mean3(sx, sy, sz, e - s, out + (size_t)p * 3, out + (size_t)p * 3 + 1, out + (size_t)p * 3 + 2);
    }
}

static int grid_for(int64_t n_points, int per_sm) {
    const int64_t tiles = (n_points + kPfnWin - 1) / kPfnWin;
    const int64_t cap = 148 * (int64_t)per_sm;
    return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// PFN forward behind a finished index pass.
// `coords` (may be null) / `coord_cols`: the train-mode statistics kernel rebuilds the pillar table on its way and writes the
// coords when it stands in for the index pass's table kernel (rdp_encode_fwd).
static int pfn_fwd_impl(const PfnLaunch *L, PfnArgs &a, const rdp_pfn_params_t *prm, const rdp_geom_t *geom, const Workspace &ws,
                        int64_t n_points, float *pillar_mean, int32_t *coords, int coord_cols, cudaStream_t st) {
    const bool train = a.train_bn != 0;
    if (pillar_mean)
        pillar_mean_kernel<<<148 * 4, 256, 0, st>>>(ws.grows, grouped_row_floats(geom->cols), ws.starts, a.counters, pillar_mean);
    if (train) {
        if (prm->stats_phase != 2) {
            a.defer_finalize = prm->stats_phase == 1;
            TableArgs t;
            t.grows = ws.grows; t.starts = ws.starts; t.counters = a.counters; t.aux = ws.aux; t.coords = coords;
            t.rs = grouped_row_floats(geom->cols); t.coord_cols = coord_cols; t.g = make_geom_dev(geom);
            RDP_CUDA_OK(L->table_stats(t, a, ws.pcap, st));
        }
        if (prm->stats_phase == 1) return RDP_OK;   // SyncBatchNorm: the caller all-reduces the totals, then phase 2
        if (prm->stats_phase == 2) {
            a.local_stats = prm->local_stats;
            a.n_from_totals = 1;
            RDP_CUDA_OK(L->bn_finalize(a, st));
        }
    }
    static const int per_sm_eval = env_int("RDP_APPLY_PER_SM", 8), per_sm_arg = env_int("RDP_APPLY_ARG_PER_SM", 6);
    const int cpl = L->cout / 32;
    const int per_sm = cpl == 1 ? (a.argpos ? per_sm_arg : per_sm_eval) : (cpl == 2 ? 4 : 2);
    RDP_CUDA_OK(L->apply(a, a.argpos ? 1 : 0, grid_for(n_points, per_sm), st));
    return RDP_OK;
}

}  // namespace rdp

using namespace rdp;

static bool misaligned(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

static int check_fwd_args(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                          const rdp_pfn_params_t *prm, void *workspace, const int32_t *counters, float *features, int32_t *argpos,
                          double *bn_state) {
    if (!geom || !layout || !prm || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!points || !workspace || !features || !prm->weight) return RDP_ERR_INVALID_ARG;
    if (misaligned(features, 16) || (argpos && misaligned(argpos, 16))) return RDP_ERR_INVALID_ARG;
    const bool use_norm = prm->gamma != nullptr;
    if (use_norm && (!prm->beta || !prm->running_mean || !prm->running_var)) return RDP_ERR_INVALID_ARG;
    if (use_norm && prm->train_bn && !bn_state) return RDP_ERR_INVALID_ARG;
    if (geom->nz > 1) return RDP_ERR_UNSUPPORTED;   // voxel grids go through the layer-stack path (rdp_stack_*)
    return RDP_OK;
}

extern "C" int rdp_config_supported(const rdp_geom_t *geom, const rdp_layout_t *layout) {
    if (!geom || !layout) return 0;
    float T[kMaxCin][kMaxG + 1];
    return lookup(geom, layout) != nullptr && geom->nz <= 1 && layout->c_in <= kMaxCin && build_T(geom, layout, T) == RDP_OK;
}

extern "C" int rdp_pfn_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                           const rdp_pfn_params_t *prm, void *workspace, size_t workspace_bytes, const int32_t *counters,
                           float *features, int32_t *argpos, float *pillar_mean, double *bn_state, void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    int rc = check_fwd_args(points, n_points, geom, layout, prm, workspace, counters, features, argpos, bn_state);
    if (rc != RDP_OK || n_points == 0) return rc;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, geom, layout, prm, ws, n_points, const_cast<int32_t *>(counters));
    if (rc != RDP_OK) return rc;
    a.features = features;
    a.argpos = argpos;
    a.bn_state = bn_state;
    return pfn_fwd_impl(L, a, prm, geom, ws, n_points, pillar_mean, nullptr, layout->coord_cols, st);
}

extern "C" int rdp_encode_fwd_frames(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom,
                                     const rdp_layout_t *layout, const rdp_pfn_params_t *prm, void *workspace,
                                     size_t workspace_bytes, int32_t *coords, int32_t *inverse, int32_t *counts, int32_t *counters,
                                     float *features, int32_t *argpos, double *bn_state, int32_t *host_mapped, void *event,
                                     void *stream_v) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (!layout) return RDP_ERR_INVALID_ARG;
    int rc = check_fwd_args(points, n_points, geom, layout, prm, workspace, counters, features, argpos, bn_state);
    if (rc != RDP_OK) return rc;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    PfnArgs a;
    rc = fill_args(&a, geom, layout, prm, ws, n_points, counters);
    if (rc != RDP_OK) return rc;
    a.features = features;
    a.argpos = argpos;
    a.bn_state = bn_state;
    if (prm->stats_phase != 2) {
        // train mode: the statistics kernel builds the pillar table (and the coords) itself
        rc = index_fwd_impl(points, frame_offsets, n_points, geom, layout->coord_cols, workspace, workspace_bytes, coords, inverse,
                            counts, counters, host_mapped, event, st, a.train_bn != 0);
        if (rc != RDP_OK || n_points == 0) return rc;
    }
    if (n_points == 0) return RDP_OK;
    return pfn_fwd_impl(L, a, prm, geom, ws, n_points, nullptr, coords, layout->coord_cols, st);
}

extern "C" int rdp_encode_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                              const rdp_pfn_params_t *params, void *workspace, size_t workspace_bytes, int32_t *coords,
                              int32_t *inverse, int32_t *counts, int32_t *counters, float *features, int32_t *argpos,
                              double *bn_state, int32_t *host_mapped, void *event, void *stream) {
    return rdp_encode_fwd_frames(points, nullptr, n_points, geom, layout, params, workspace, workspace_bytes, coords, inverse, counts,
                                 counters, features, argpos, bn_state, host_mapped, event, stream);
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
    (void)features;   // the ReLU mask travels in argpos (-1); kept in the signature for ABI stability
    (void)points;     // the rows are re-read from the workspace's grouped copy
    if (!workspace || !grad_features || !argpos) return RDP_ERR_INVALID_ARG;
    // both are sources of 16-byte bulk (TMA) copies
    if (misaligned(grad_features, 16) || misaligned(argpos, 16)) return RDP_ERR_INVALID_ARG;
    if (train && !bn_state) return RDP_ERR_INVALID_ARG;
    if (geom->nz > 1) return RDP_ERR_UNSUPPORTED;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.total_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    PfnArgs a;
    rc = fill_args(&a, geom, layout, prm, ws, n_points, const_cast<int32_t *>(counters));
    if (rc != RDP_OK) return rc;
    a.grad = grad_features;
    a.argpos = const_cast<int32_t *>(argpos);
    a.bn_state = const_cast<double *>(bn_state);
    a.d_weight = d_weight; a.d_gamma = d_gamma; a.d_beta = d_beta;
    if (prm->stats_phase == 2) {   // SyncBatchNorm: the sums were left in acc_bwd by phase 1; `global_bwd` = their all-reduce
        if (!prm->global_bwd) return RDP_ERR_INVALID_ARG;
        RDP_CUDA_OK(L->bwd_finalize(a, prm->global_bwd, st));
        return RDP_OK;
    }
    a.defer_finalize = prm->stats_phase == 1;
    static const int per_sm = env_int("RDP_BWD_PER_SM", 5);
    const int cpl = L->cout / 32;
    RDP_CUDA_OK(L->bwd(a, grid_for(n_points, cpl == 1 ? per_sm : (cpl == 2 ? 3 : 2)), st));
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
    argpos_to_kept_kernel<<<148 * 8, 128, 0, st>>>(argpos, ws.grows, grouped_row_floats(geom->cols), ws.starts, ws.orig2kept, counters,
                                                    n_points, layout->c_out, argmax_kept);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" int rdp_stats_buffers(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, size_t *stats_offset,
                                 int64_t *stats_doubles, size_t *bwd_offset, int64_t *bwd_doubles) {
    if (!geom || !layout || !stats_offset || !stats_doubles || !bwd_offset || !bwd_doubles) return RDP_ERR_INVALID_ARG;
    const PfnLaunch *L = lookup(geom, layout);
    if (!L) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    char *const fake = reinterpret_cast<char *>(static_cast<uintptr_t>(1) << 20);   // carve from a fake base to read the offsets off
    int rc = carve_workspace(fake, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    *stats_offset = (size_t)(reinterpret_cast<char *>(ws.acc_stats) - fake);
    *bwd_offset = (size_t)(reinterpret_cast<char *>(ws.acc_bwd) - fake);
    const int G = L->g;
    *stats_doubles = G + G * G + 1;                           // S1 | S2 | number of points
    *bwd_doubles = (int64_t)L->cout * (G + 1);
    return RDP_OK;
}
