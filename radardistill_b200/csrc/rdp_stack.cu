// rdp_stack.cu -- the pieces of the path outside the single fused PFN layer:
//
//   rdp_decorate          per-point decorated features (N, c_in) in kept-point order -- what the reference concatenates
//                         before its first PFNLayerV2 (dynamic_pillar_vfe.py:214-237 / :105-121, dynamic_voxel_vfe.py:73-91)
//   rdp_segment_max_fwd   scatter_max over the pillars (PFNLayerV2.forward :40) for (N, C) activations of stacked layers
//   rdp_segment_max_bwd   its autograd: route the gradient to the winning row
//   rdp_voxel_mean        DynamicMeanVFE: per-voxel mean of every point column (dynamic_mean_vfe.py:63-65)
//   rdp_prepare_points    device-side input prep: range mask of data_processor.py:80-86 (mask_points_by_range) with stable
//                         compaction, optionally followed by the shuffle of :99-114 as a fixed pseudo-random permutation
//
// Stacked PFN layers (NUM_FILTERS longer than 1) and the voxel encoders run:  index kernels (rdp_index_fwd, 2-D or 3-D
// key) -> rdp_decorate -> per layer {Linear + BatchNorm: plain library GEMM / cuDNN through the host framework;
// rdp_segment_max_fwd; gather + concat for non-last layers}.  All pooling / grouping work reuses the pillar-grouped row
// order and the pillar table of the index pass.
#include "rdp_index_host.h"

namespace rdp {

struct DecorateArgs {
    const float *grows;
    const float *aux;
    const int32_t *counters;
    const int32_t *orig2kept;
    float *out;
    long long n0;
    int rs, cols, c_in;
    int layout, use_abs, use_cluster, use_relative, with_distance;
    float lo_x, lo_y, lo_z;
};

// thread = grouped row: the row, its pillar's table entry, the layout's feature order; written at the row's kept index.
// Arithmetic: f_center = xyz - centre exactly as the reference (:215-217); f_cluster = f_center + (centre - mean), the
// canonical form of xyz - mean used throughout this library; f_relative = xyz - lo (:233-236).
__global__ void __launch_bounds__(256) decorate_kernel(const __grid_constant__ DecorateArgs a) {
    const long long N = a.counters[RDP_CNT_N];
    const bool none_dropped = (N == a.n0);
    const int C = a.cols - 1;   // raw point features incl. xyz
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const float *r = a.grows + ((size_t)i + 1) * a.rs;
        const int orig = __float_as_int(r[a.rs - 2]), gid = __float_as_int(r[a.rs - 1]);
        const float *t = a.aux + (size_t)gid * 8;
        const float4 t0 = __ldg(reinterpret_cast<const float4 *>(t));       // centre xy, (centre - mean) xy
        const float4 t1 = __ldg(reinterpret_cast<const float4 *>(t + 4));   // (centre - mean) z, start, rows, centre z
        const float x = r[1], y = r[2], z = r[3];
        const float dx = __fsub_rn(x, t0.x), dy = __fsub_rn(y, t0.y), dz = __fsub_rn(z, t1.w);
        const float cl[3] = {__fadd_rn(dx, t0.z), __fadd_rn(dy, t0.w), __fadd_rn(dz, t1.x)};
        const long long j = none_dropped ? (long long)orig : (long long)a.orig2kept[orig];
        float *o = a.out + (size_t)j * a.c_in;
        int k = 0;
        auto center = [&]() { o[k++] = dx; o[k++] = dy; o[k++] = dz; };
        auto cluster = [&]() { o[k++] = cl[0]; o[k++] = cl[1]; o[k++] = cl[2]; };
        auto points = [&]() {
            if (a.use_abs) { o[k++] = x; o[k++] = y; o[k++] = z; }
            for (int c = 4; c <= C; ++c) o[k++] = r[c];
        };
        auto dist = [&]() { if (a.with_distance) o[k++] = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x)))); };
        if (a.layout == RDP_LAYOUT_SIMPLE2D) {
            center();
            points();
            if (a.use_cluster) cluster();
            dist();
            if (a.use_relative) { o[k++] = __fsub_rn(x, a.lo_x); o[k++] = __fsub_rn(y, a.lo_y); o[k++] = __fsub_rn(z, a.lo_z); }
        } else {   // DYNPILLAR / DYNVOXEL: [points | f_cluster | f_center | dist]
            points();
            cluster();
            center();
            dist();
        }
    }
}

// warp = pillar, lane = channel (+32, +64, ...): running maximum over the pillar's rows of x (kept-point order, gathered
// through the grouped order); ties -> lowest kept index (torch_scatter's CPU rule, the documented tie rule).
__global__ void __launch_bounds__(256) segment_max_kernel(const float *__restrict__ x, int c, const float *__restrict__ grows, int rs,
                                                          const int32_t *__restrict__ starts, const int32_t *__restrict__ orig2kept,
                                                          const int32_t *__restrict__ counters, long long n0,
                                                          float *__restrict__ out, int32_t *__restrict__ arg) {
    const int P = counters[RDP_CNT_P];
    const bool none_dropped = ((long long)counters[RDP_CNT_N] == n0);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += warps) {
        const int s = starts[p], e = starts[p + 1];
        for (int ch = lane; ch < c; ch += 32) {
            float m = __int_as_float(0xff800000);
            int mj = 0x7fffffff;
            for (int i = s; i < e; ++i) {
                const int orig = __float_as_int(__ldg(grows + ((size_t)i + 1) * rs + rs - 2));
                const int j = none_dropped ? orig : orig2kept[orig];
                const float v = __ldg(x + (size_t)j * c + ch);
                if (v > m || (v == m && j < mj)) { m = v; mj = j; }
            }
            out[(size_t)p * c + ch] = m;
            arg[(size_t)p * c + ch] = mj;
        }
    }
}

__global__ void __launch_bounds__(256) segment_max_bwd_kernel(const float *__restrict__ g, const int32_t *__restrict__ arg, long long pc, int c,
                                                              float *__restrict__ gx) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < pc; e += (long long)gridDim.x * blockDim.x)
        gx[(size_t)arg[e] * c + (int)(e % c)] = g[e];   // one winner per (pillar, channel): plain stores, no atomics
}

// thread = voxel: fp64 sum of every point column over the voxel's rows (exact for any realistic voxel), correctly
// rounded quotient, one rounding to fp32 (the canonical mean of this library; scatter_mean, dynamic_mean_vfe.py:65)
__global__ void __launch_bounds__(256) voxel_mean_kernel(const float *__restrict__ grows, int rs, int cols, const int32_t *__restrict__ starts,
                                                         const int32_t *__restrict__ counters, float *__restrict__ out) {
    const int P = counters[RDP_CNT_P];
    const int C = cols - 1;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const int s = starts[p], e = starts[p + 1];
        const double n = (double)(e - s);
        for (int c = 0; c < C; ++c) {
            double acc = 0.0;
            for (int i = s; i < e; ++i) acc += (double)__ldg(grows + ((size_t)i + 1) * rs + 1 + c);
            out[(size_t)p * C + c] = (float)__ddiv_rn(acc, n);
        }
    }
}

// ----------------------------------------------------------------------------- input prep
// mask_points_by_range (pcdet/utils/common_utils.py:85-88 via data_processor.py:80-86): keep rows with
// lo_x <= x <= hi_x and lo_y <= y <= hi_y (inclusive, as the reference), stable.  Two kernels: per-tile counts, then a
// chained scan + compaction.  With `shuffle_seed != 0` the kept rows are written through a fixed pseudo-random permutation
// (bijective multiplicative hash of the kept index) instead of in order -- the stand-in for np.random.permutation of
// data_processor.py:99-114, which exists to decorrelate the voxel sampling order; every output of this encoder except the
// order of `inverse` is invariant under it.
constexpr int kPrepThreads = 256;

__global__ void __launch_bounds__(kPrepThreads) prep_count_kernel(const float *__restrict__ pts, long long n0, int cols, int xcol, float lo_x,
                                                                  float lo_y, float hi_x, float hi_y, int32_t *__restrict__ tile_keep) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * kPrepThreads + threadIdx.x;
    bool keep = false;
    if (i < n0) {
        const float x = pts[i * cols + xcol], y = pts[i * cols + xcol + 1];
        keep = (x >= lo_x) && (x <= hi_x) && (y >= lo_y) && (y <= hi_y);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) tile_keep[blockIdx.x] = s_cnt;
}

// exclusive scan of the tile counts by one CTA (tiles <= a few 10^4), total -> n_out
__global__ void __launch_bounds__(1024) prep_scan_kernel(int32_t *__restrict__ tile_keep, int tiles, int32_t *__restrict__ n_out) {
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int t = base + (int)threadIdx.x;
        const int v = t < tiles ? tile_keep[t] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const int add = threadIdx.x >= (unsigned)d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            __syncthreads();
        }
        if (t < tiles) tile_keep[t] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

__device__ __forceinline__ unsigned long long perm_index(unsigned long long j, unsigned long long n, unsigned long long seed) {
    // cycle-walking over the next power of two: an odd multiplier + xor-shift is a bijection on [0, 2^k)
    unsigned k = 1;
    while ((1ull << k) < n) ++k;
    const unsigned long long mask = (1ull << k) - 1ull;
    unsigned long long v = j;
    do {
        v = (v * (2ull * seed + 0x9E3779B97F4A7C15ull | 1ull)) & mask;
        v ^= v >> (k / 2 + 1);
        v = (v * 0xD6E8FEB86659FD93ull | 0ull) & mask;
        v = (v + seed) & mask;
    } while (v >= n);
    return v;
}

__global__ void __launch_bounds__(kPrepThreads) prep_compact_kernel(const float *__restrict__ pts, long long n0, int cols, int xcol, float lo_x,
                                                                    float lo_y, float hi_x, float hi_y, const int32_t *__restrict__ tile_base,
                                                                    const int32_t *__restrict__ n_out, unsigned long long shuffle_seed,
                                                                    float *__restrict__ out) {
    __shared__ int s_w[kPrepThreads / 32];
    const long long i = (long long)blockIdx.x * kPrepThreads + threadIdx.x;
    bool keep = false;
    if (i < n0) {
        const float x = pts[i * cols + xcol], y = pts[i * cols + xcol + 1];
        keep = (x >= lo_x) && (x <= hi_x) && (y >= lo_y) && (y <= hi_y);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (keep) {
        unsigned long long j = (unsigned long long)tile_base[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
        if (shuffle_seed) j = perm_index(j, (unsigned long long)*n_out, shuffle_seed);
        for (int c = 0; c < cols; ++c) out[j * cols + c] = pts[i * cols + c];
    }
}

}  // namespace rdp

using namespace rdp;

static int grid_cap(long long work, int threads, int per_sm) {
    const long long blocks = (work + threads - 1) / threads, cap = 148ll * per_sm;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

extern "C" int rdp_decorate(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, void *workspace, size_t workspace_bytes,
                            const int32_t *counters, float *features, void *stream_v) {
    if (!geom || !layout || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!workspace || !features) return RDP_ERR_INVALID_ARG;
    if (layout->layout != RDP_LAYOUT_SIMPLE2D && layout->layout != RDP_LAYOUT_DYNPILLAR && layout->layout != RDP_LAYOUT_DYNVOXEL)
        return RDP_ERR_UNSUPPORTED;
    const int C = geom->cols - 1;
    int c_in;
    if (layout->layout == RDP_LAYOUT_SIMPLE2D)
        c_in = 3 + (layout->use_abs ? C : C - 3) + (layout->use_cluster ? 3 : 0) + (layout->use_relative ? 3 : 0);
    else
        c_in = (layout->use_abs ? C : C - 3) + 6;
    c_in += layout->with_distance ? 1 : 0;
    if (c_in != layout->c_in) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, layout, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    DecorateArgs a;
    a.grows = ws.grows; a.aux = ws.aux; a.counters = counters; a.orig2kept = ws.orig2kept; a.out = features; a.n0 = n_points;
    a.rs = grouped_row_floats(geom->cols); a.cols = geom->cols; a.c_in = c_in;
    a.layout = layout->layout; a.use_abs = layout->use_abs; a.use_cluster = layout->layout == RDP_LAYOUT_SIMPLE2D ? layout->use_cluster : 1;
    a.use_relative = layout->layout == RDP_LAYOUT_SIMPLE2D ? layout->use_relative : 0; a.with_distance = layout->with_distance;
    a.lo_x = geom->lo[0]; a.lo_y = geom->lo[1]; a.lo_z = geom->lo[2];
    decorate_kernel<<<grid_cap(n_points, 256, 8), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(a);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" int rdp_segment_max_fwd(const float *x, int32_t channels, int64_t n_points, const rdp_geom_t *geom, void *workspace,
                                   size_t workspace_bytes, const int32_t *counters, float *out, int32_t *argmax_kept, void *stream_v) {
    if (!geom || !counters || n_points < 0 || channels <= 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!x || !workspace || !out || !argmax_kept) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, nullptr, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    segment_max_kernel<<<grid_cap(ws.pcap * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
        x, channels, ws.grows, grouped_row_floats(geom->cols), ws.starts, ws.orig2kept, counters, n_points, out, argmax_kept);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" int rdp_segment_max_bwd(const float *grad_out, const int32_t *argmax_kept, int64_t n_pillars, int32_t channels, float *grad_x,
                                   void *stream_v) {
    if (n_pillars < 0 || channels <= 0) return RDP_ERR_INVALID_ARG;
    if (n_pillars == 0) return RDP_OK;
    if (!grad_out || !argmax_kept || !grad_x) return RDP_ERR_INVALID_ARG;
    segment_max_bwd_kernel<<<grid_cap(n_pillars * channels, 256, 8), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
        grad_out, argmax_kept, (long long)n_pillars * channels, channels, grad_x);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" int rdp_voxel_mean(int64_t n_points, const rdp_geom_t *geom, void *workspace, size_t workspace_bytes, const int32_t *counters,
                              float *mean, void *stream_v) {
    if (!geom || !counters || n_points < 0) return RDP_ERR_INVALID_ARG;
    if (n_points == 0) return RDP_OK;
    if (!workspace || !mean) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, nullptr, &ws);
    if (rc != RDP_OK) return rc;
    if (ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    voxel_mean_kernel<<<grid_cap(ws.pcap, 256, 8), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
        ws.grows, grouped_row_floats(geom->cols), geom->cols, ws.starts, counters, mean);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" size_t rdp_prepare_scratch_bytes(int64_t n_points) {
    const size_t tiles = (size_t)((n_points > 0 ? n_points : 1) + kPrepThreads - 1) / kPrepThreads;
    return sizeof(int32_t) * (tiles + 8);
}

extern "C" int rdp_prepare_points(const float *points, int64_t n_points, int32_t cols, int32_t x_col, const float *range_xy_lo_hi,
                                  uint64_t shuffle_seed, void *scratch, size_t scratch_bytes, float *out, int32_t *n_out, void *stream_v) {
    if (n_points < 0 || cols < 2 || x_col < 0 || x_col + 1 >= cols || !range_xy_lo_hi || !n_out) return RDP_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream_v);
    if (n_points == 0) {
        RDP_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st));
        return RDP_OK;
    }
    if (!points || !scratch || !out) return RDP_ERR_INVALID_ARG;
    if (n_points >= (1ll << 31) - 8) return RDP_ERR_UNSUPPORTED;
    if (scratch_bytes < rdp_prepare_scratch_bytes(n_points)) return RDP_ERR_WORKSPACE;
    const int tiles = (int)((n_points + kPrepThreads - 1) / kPrepThreads);
    int32_t *tile_keep = static_cast<int32_t *>(scratch);
    const float lx = range_xy_lo_hi[0], ly = range_xy_lo_hi[1], hx = range_xy_lo_hi[2], hy = range_xy_lo_hi[3];
    prep_count_kernel<<<tiles, kPrepThreads, 0, st>>>(points, n_points, cols, x_col, lx, ly, hx, hy, tile_keep);
    prep_scan_kernel<<<1, 1024, 0, st>>>(tile_keep, tiles, n_out);
    prep_compact_kernel<<<tiles, kPrepThreads, 0, st>>>(points, n_points, cols, x_col, lx, ly, hx, hy, tile_keep, n_out, shuffle_seed, out);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}
