// rdp_abi.cu -- status strings, workspace carving and the host-buffer convenience entry point.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rdp_common.cuh"

namespace rdp {

static thread_local char g_last_cuda_error[256] = "";

void set_last_cuda_error(cudaError_t e, const char *where) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), where);
}

bool pdl_enabled() {
    static const bool on = [] { const char *v = getenv("RDP_NO_PDL"); return !(v && v[0] == '1'); }();
    return on;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int carve_workspace(void *base, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, Workspace *ws) {
    (void)layout;
    if (!geom || !ws || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (geom->nx <= 0 || geom->ny <= 0 || geom->batch_size <= 0 || geom->cols < 4) return RDP_ERR_INVALID_ARG;
    const int64_t cells = (int64_t)geom->batch_size * geom->nx * geom->ny * (geom->nz > 1 ? geom->nz : 1);
    if (cells >= (1ll << 31)) return RDP_ERR_KEYSPACE;
    memset(ws, 0, sizeof(*ws));
    const int64_t n = n_points > 0 ? n_points : 1;
    ws->n = n;
    ws->words = (cells + 31) / 32;
    ws->pcap = n < cells ? n : cells;
    ws->index_tiles = (n + kIndexTileRows - 1) / kIndexTileRows;
    ws->pfn_tiles = (n + kPfnWin - 1) / kPfnWin;

    char *p = static_cast<char *>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *r = p ? p + off : nullptr;
        off = align_up(off + bytes, 256);
        return r;
    };
    const size_t words_pad = align_up((size_t)ws->words, 4) + 4;
    ws->zero_begin = take(0);
    ws->scan_state_a = reinterpret_cast<uint64_t *>(take(sizeof(uint64_t) * kScanGrid));
    ws->scan_state_b = reinterpret_cast<uint64_t *>(take(sizeof(uint64_t) * kScanGrid));
    ws->acc_stats = reinterpret_cast<double *>(take(sizeof(double) * kMaxAcc));
    ws->acc_bwd = reinterpret_cast<double *>(take(sizeof(double) * kMaxCout * (kMaxG + 1)));
    ws->bitmap = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * words_pad));
    ws->zero_bytes = off;
    ws->wordrank = reinterpret_cast<uint2 *>(take(sizeof(uint2) * words_pad));
    ws->keys = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(n + 4)));
    ws->ranks = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(n + 4)));
    ws->slots = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(n + 4)));
    ws->tile_keep = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)ws->index_tiles));
    ws->starts = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(ws->pcap + kSliceInts + 8)));
    const size_t pad = kPfnCap + 8;
    ws->grows = reinterpret_cast<float *>(take(sizeof(float) * (size_t)(n + pad + 1) * grouped_row_floats(geom->cols)));
    ws->aux = reinterpret_cast<float *>(take(sizeof(float) * 8 * (size_t)(ws->pcap + kPfnWin + 8)));
    ws->tile_first = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(ws->pfn_tiles + 4)));
    ws->orig2kept = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(n + 4)));
    ws->kept2orig = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)(n + 4)));
    ws->index_bytes = off;
    ws->total_bytes = off;
    return RDP_OK;
}

}  // namespace rdp

using namespace rdp;

extern "C" int rdp_abi_version(void) { return RDP_ABI_VERSION; }

extern "C" const char *rdp_status_string(int status) {
    switch (status) {
        case RDP_OK: return "ok";
        case RDP_ERR_INVALID_ARG: return "invalid argument (null / misaligned pointer or bad size)";
        case RDP_ERR_WORKSPACE: return "workspace too small (see rdp_workspace_bytes)";
        case RDP_ERR_CUDA: return "CUDA runtime error (see rdp_last_cuda_error)";
        case RDP_ERR_UNSUPPORTED: return "unsupported configuration";
        case RDP_ERR_KEYSPACE: return "batch_size*nx*ny exceeds the int32 merged-key space";
        default: return "unknown status";
    }
}

extern "C" const char *rdp_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int rdp_workspace_bytes(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, size_t *bytes) {
    if (!bytes) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(nullptr, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    *bytes = ws.total_bytes;
    return RDP_OK;
}

extern "C" int64_t rdp_bn_state_doubles(const rdp_layout_t *layout) {
    if (!layout) return 0;
    // [mean(cout) | var(cout) | scale(cout) | shift(cout) | n | S1(G) | S2(G*G)] in the reduced basis (G <= kMaxG)
    return 4 * (int64_t)layout->c_out + 1 + kMaxG + kMaxG * kMaxG;
}

extern "C" int rdp_encode_host(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                               const rdp_pfn_params_t *hp, float *features, int32_t *coords, int32_t *inverse,
                               int32_t *counts, int64_t *n_kept, int64_t *n_pillars) {
    if (!geom || !layout || !hp || !n_kept || !n_pillars || n_points < 0) return RDP_ERR_INVALID_ARG;
    *n_kept = *n_pillars = 0;
    if (n_points == 0) return RDP_OK;
    if (!points || !features || !coords || !inverse || !counts || !hp->weight) return RDP_ERR_INVALID_ARG;
    const int cin = layout->c_in, cout = layout->c_out, kc = layout->coord_cols;
    size_t ws_bytes = 0;
    int rc = rdp_workspace_bytes(n_points, geom, layout, &ws_bytes);
    if (rc != RDP_OK) return rc;
    cudaStream_t st;
    RDP_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t n = (size_t)n_points;
    float *d_pts = nullptr, *d_feat = nullptr, *d_par = nullptr;
    int32_t *d_coords = nullptr, *d_inv = nullptr, *d_cnt = nullptr, *d_counters = nullptr;
    void *d_ws = nullptr;
    int32_t h_counters[RDP_NUM_COUNTERS] = {0};
    auto cleanup = [&]() {
        cudaFree(d_pts); cudaFree(d_feat); cudaFree(d_par); cudaFree(d_coords); cudaFree(d_inv); cudaFree(d_cnt);
        cudaFree(d_counters); cudaFree(d_ws); cudaStreamDestroy(st);
    };
#define RDP_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_cuda_error(_e, #expr); cleanup(); return RDP_ERR_CUDA; } } while (0)
    RDP_TRY(cudaMalloc(&d_pts, n * geom->cols * sizeof(float)));
    RDP_TRY(cudaMalloc(&d_feat, n * cout * sizeof(float)));
    RDP_TRY(cudaMalloc(&d_coords, n * kc * sizeof(int32_t)));
    RDP_TRY(cudaMalloc(&d_inv, (n + 4) * sizeof(int32_t)));
    RDP_TRY(cudaMalloc(&d_cnt, (n + 4) * sizeof(int32_t)));
    RDP_TRY(cudaMalloc(&d_counters, sizeof(h_counters)));
    RDP_TRY(cudaMalloc(&d_ws, ws_bytes));
    RDP_TRY(cudaMalloc(&d_par, sizeof(float) * (size_t)(cout * cin + 5 * cout)));
    RDP_TRY(cudaMemcpyAsync(d_pts, points, n * geom->cols * sizeof(float), cudaMemcpyHostToDevice, st));
    rdp_pfn_params_t dp = *hp;
    float *q = d_par;
    auto up = [&](const float *h, size_t cnt) -> float * {
        if (!h) return nullptr;
        float *d = q;
        q += cnt;
        cudaMemcpyAsync(d, h, cnt * sizeof(float), cudaMemcpyHostToDevice, st);
        return d;
    };
    dp.weight = up(hp->weight, (size_t)cout * cin);
    dp.bias = up(hp->bias, cout);
    dp.gamma = up(hp->gamma, cout);
    dp.beta = up(hp->beta, cout);
    dp.running_mean = up(hp->running_mean, cout);
    dp.running_var = up(hp->running_var, cout);
    dp.train_bn = 0;
    dp.num_batches_tracked = nullptr;
    rc = rdp_index_fwd(d_pts, n_points, geom, kc, d_ws, ws_bytes, d_coords, d_inv, d_cnt, d_counters, st);
    if (rc == RDP_OK)
        rc = rdp_pfn_fwd(d_pts, n_points, geom, layout, &dp, d_ws, ws_bytes, d_counters, d_feat, nullptr, nullptr, nullptr, st);
    if (rc != RDP_OK) { cleanup(); return rc; }
    RDP_TRY(cudaMemcpyAsync(h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    RDP_TRY(cudaStreamSynchronize(st));
    const size_t N = (size_t)h_counters[RDP_CNT_N], P = (size_t)h_counters[RDP_CNT_P];
    if (h_counters[RDP_CNT_ERRFLAGS]) { cleanup(); return RDP_ERR_INVALID_ARG; }
    RDP_TRY(cudaMemcpyAsync(features, d_feat, P * cout * sizeof(float), cudaMemcpyDeviceToHost, st));
    RDP_TRY(cudaMemcpyAsync(coords, d_coords, P * kc * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RDP_TRY(cudaMemcpyAsync(counts, d_cnt, P * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RDP_TRY(cudaMemcpyAsync(inverse, d_inv, N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RDP_TRY(cudaStreamSynchronize(st));
#undef RDP_TRY
    *n_kept = (int64_t)N;
    *n_pillars = (int64_t)P;
    cleanup();
    return RDP_OK;
}
