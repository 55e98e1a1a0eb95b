// rdp_pfn.cuh -- device code of the fused pillar feature network (shared by the per-config .cu files).
//
// Replaces /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:214-240 and
// PFNLayerV2.forward :35-46: scatter_mean, f_center / f_cluster / f_relative, concat,
// Linear(no bias)+BatchNorm1d+ReLU, scatter_max (+argmax), and the autograd of that chain.
//
// Work decomposition (one CTA = 256 threads, persistent over pillar-aligned tiles of the grouped order):
//   batch    = up to 256 grouped points that form WHOLE pillars (a pillar with more points than
//              that takes the "giant" path: block-wide mean, then chunks with a running max)
//   phase A  thread = point : gather the row, find its pillar in the batch, stash xyz
//   phase B  thread = pillar: fp64 sum of xyz -> mean (one rounding), pillar centre
//   phase C1 thread = point : decorated features -> smem  (same op order/roundings as the reference)
//   phase C2 thread = 4 channels x PT points register tile: x = W f as a k-ascending fmaf chain with the
//            weight rows held in registers, then y = fma(x, scale, shift)       -> smem
//   phase D  thread = (pillar, 4 channels): max over the pillar's rows (+ lowest-index argmax), coalesced store
// Arithmetic is the canonical form of oracle/pillar_oracle.c (ORC_MEAN_F64), so every output is
// bit-identical to the oracle.
#pragma once

#include "rdp_common.cuh"

namespace rdp {

constexpr int kMaxSuper = 24;

struct PfnArgs {
    const float *pts;
    const int32_t *order, *ends, *tile_start, *counters, *coords, *orig2kept, *kept2orig;
    const float *weight, *bias, *gamma, *beta, *rmean, *rvar;
    const double *bn_state;  // train apply: folded scale / shift live here
    float *features;
    int32_t *argmax;
    float *pillar_mean;
    double *partials;
    long long n0;
    double eps;
    float lo[3], vsz[3], off[3];
    int c_in;
    int coord_cols;
    int use_norm;
    int fold_from_state;  // 1: scale/shift from bn_state (train), 0: fold running stats in-kernel (eval)
    int8_t kmap[kMaxSuper];  // super-feature -> layout column of W, or -1 (zero weight)
};

// Compile-time shape of one encoder family.  "Super features" are every decoration the layout could
// use, in the layout's concat order; options switched off in model_cfg get a zero weight column
// (fmaf(0, f, acc) == acc), so one instantiation serves all flag combinations bit-exactly.
template <int COLS_, int LAYOUT_, bool DIST_, int COUT_>
struct PfnCfg {
    static constexpr int COLS = COLS_, LAYOUT = LAYOUT_, COUT = COUT_, C = COLS_ - 1;
    static constexpr bool DIST = DIST_;
    static constexpr int CS = (LAYOUT_ == RDP_LAYOUT_SIMPLE2D) ? (3 + C + 3 + (DIST_ ? 1 : 0) + 3) : (C + 6 + (DIST_ ? 1 : 0));
    static constexpr int CSP4 = (CS + 3) / 4 * 4;
    static constexpr int FSTRIDE = (CSP4 % 16 == 0) ? CSP4 + 4 : CSP4;  // smem row stride of features (conflict free)
    static constexpr int QUADS = COUT / 4;                               // channel quads
    static constexpr int GROUPS = kPfnThreads / QUADS;                   // point groups in the register tiling
    static constexpr int BATCH = (COUT <= 32) ? 256 : 128;               // points per batch
    static constexpr int PT = BATCH / GROUPS;                            // points per thread in phase C2
    static constexpr int ZSTRIDE = COUT + 4;
    static_assert(CS <= kMaxSuper, "too many features");
    static_assert(BATCH % GROUPS == 0 && BATCH <= kPfnThreads, "tiling");
};

template <class Cfg>
struct PfnSmem {
    float z[Cfg::BATCH * Cfg::ZSTRIDE];   // activations of the batch
    float f[Cfg::BATCH * Cfg::FSTRIDE];   // decorated features of the batch
    float xyz[Cfg::BATCH * 3];
    float mean[Cfg::BATCH * 3];
    float cen[Cfg::BATCH * 2];
    int ends[Cfg::BATCH + 1];
    int row[Cfg::BATCH];                  // original row of each point
    int kept[Cfg::BATCH];                 // kept-order index of each point (argmax numbering)
    int lp[Cfg::BATCH];                   // pillar (within batch) of each point
    float scale[Cfg::COUT], shift[Cfg::COUT];
    float part_v[(kPfnThreads / Cfg::QUADS) * Cfg::COUT];  // giant path partials
    int part_i[(kPfnThreads / Cfg::QUADS) * Cfg::COUT];
    float carry_v[Cfg::COUT];
    int carry_i[Cfg::COUT];
    double red[kPfnThreads * 3];
    int misc[4];
};

enum { PFN_MODE_APPLY = 0, PFN_MODE_STATS = 1 };

// ------------------------------------------------------------------------------------------- features
template <class Cfg>
__device__ __forceinline__ void decorate(const float *r, float cenx, float ceny, const float *mean, const PfnArgs &a, float *f) {
    const float x = r[1], y = r[2], z = r[3];
    float cen[3], clu[3];
    cen[0] = __fsub_rn(x, cenx);               // x - (cx*vx + x_off); the bracket is per pillar (phase B)
    cen[1] = __fsub_rn(y, ceny);
    cen[2] = __fsub_rn(z, a.off[2]);           // (:217) z - z_offset
    clu[0] = __fsub_rn(x, mean[0]);            // (:227) xyz - mean[inv]
    clu[1] = __fsub_rn(y, mean[1]);
    clu[2] = __fsub_rn(z, mean[2]);
    int k = 0;
    if (Cfg::LAYOUT == RDP_LAYOUT_SIMPLE2D) {
        f[k++] = cen[0]; f[k++] = cen[1]; f[k++] = cen[2];
#pragma unroll
        for (int c = 1; c <= Cfg::C; ++c) f[k++] = r[c];
        f[k++] = clu[0]; f[k++] = clu[1]; f[k++] = clu[2];
        if (Cfg::DIST) f[k++] = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
        f[k++] = __fsub_rn(x, a.lo[0]); f[k++] = __fsub_rn(y, a.lo[1]); f[k++] = __fsub_rn(z, a.lo[2]);  // (:234)
    } else {
#pragma unroll
        for (int c = 1; c <= Cfg::C; ++c) f[k++] = r[c];
        f[k++] = clu[0]; f[k++] = clu[1]; f[k++] = clu[2];
        f[k++] = cen[0]; f[k++] = cen[1]; f[k++] = cen[2];
        if (Cfg::DIST) f[k++] = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
    }
}

template <class Cfg>
__device__ __forceinline__ void load_row(const float *pts, long long row, float *r) {
    const float *p = pts + row * Cfg::COLS;
    if (Cfg::COLS % 2 == 0) {
#pragma unroll
        for (int c = 0; c < Cfg::COLS; c += 2) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(p + c));
            r[c] = v.x; r[c + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < Cfg::COLS; ++c) r[c] = __ldg(p + c);
    }
}

// BatchNorm folded to y = fma(x, scale, shift) in fp64 with one rounding (oracle: orc_bn_fold).
__device__ __forceinline__ void fold_bn(double gamma, double beta, double mean, double var, double eps, float *scale, float *shift) {
    const double inv_std = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(var, eps)));
    const double s = __dmul_rn(gamma, inv_std);
    *scale = (float)s;
    *shift = (float)__dsub_rn(beta, __dmul_rn(mean, s));
}

// ------------------------------------------------------------------------------------------- forward
template <class Cfg, int MODE>
__global__ void __launch_bounds__(kPfnThreads, 2) pfn_fwd_kernel(const __grid_constant__ PfnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PfnSmem<Cfg> &S = *reinterpret_cast<PfnSmem<Cfg> *>(smem_raw);
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, QUADS = Cfg::QUADS, GROUPS = Cfg::GROUPS, PT = Cfg::PT, BATCH = Cfg::BATCH;
    const int tid = threadIdx.x;
    const long long N = a.counters[RDP_CNT_N];
    const int P = a.counters[RDP_CNT_P];
    const bool none_dropped = (N == a.n0);
    const int ntiles = (int)((N + kPfnTileRows - 1) / kPfnTileRows);
    const bool want_arg = (MODE == PFN_MODE_APPLY) && a.argmax != nullptr;

    // ---- per-CTA constants: BN fold, weight rows of my channel quad in registers
    if (tid < COUT) {
        float sc = 1.0f, sh = a.bias ? a.bias[tid] : 0.0f;
        if (a.use_norm) {
            if (a.fold_from_state) {
                sc = (float)a.bn_state[2 * COUT + tid];
                sh = (float)a.bn_state[3 * COUT + tid];
            } else {
                fold_bn((double)a.gamma[tid], (double)a.beta[tid], (double)a.rmean[tid], (double)a.rvar[tid], a.eps, &sc, &sh);
            }
        }
        S.scale[tid] = sc;
        S.shift[tid] = sh;
    }
    const int quad = tid % QUADS, grp = tid / QUADS;
    float W[4][CS];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int k = a.kmap[s];
            W[j][s] = (k >= 0) ? __ldg(a.weight + (quad * 4 + j) * a.c_in + k) : 0.0f;
        }
    __syncthreads();
    float sc4[4], sh4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sc4[j] = S.scale[quad * 4 + j]; sh4[j] = S.shift[quad * 4 + j]; }

    // ---- STATS accumulators (fp64, live across the whole CTA lifetime)
    double st_x = 0.0, st_x2 = 0.0, st_m[2] = {0.0, 0.0};
    constexpr int NPAIR = CS * (CS + 1) / 2 + CS;  // S2 upper triangle, then S1
    int pa[2] = {0, 0}, pb[2] = {0, 0};
    if (MODE == PFN_MODE_STATS) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int item = tid + u * kPfnThreads;
            if (item < CS * (CS + 1) / 2) {
                int aa = 0, rem = item;
                while (rem >= CS - aa) { rem -= CS - aa; ++aa; }
                pa[u] = aa; pb[u] = aa + rem;
            } else if (item < NPAIR) {
                pa[u] = item - CS * (CS + 1) / 2; pb[u] = -1;
            } else {
                pa[u] = -1;
            }
        }
    }

    // One batch: np points starting at grouped position s; nb pillars starting at p (nb == 0: chunk of giant pillar p).
    auto run_points = [&](int s, int np, int p, int nb) {
        // phase A
        if (tid < np) {
            const int gpos = s + tid;
            const int row = a.order[gpos];
            S.row[tid] = row;
            int lp = 0;
            if (nb > 0) {
                int lo = 0, hi = nb;  // first q with ends[q] > gpos
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (S.ends[mid] > gpos) hi = mid; else lo = mid + 1;
                }
                lp = lo;
            }
            S.lp[tid] = lp;
            const float *r = a.pts + (long long)row * Cfg::COLS;
            S.xyz[tid * 3 + 0] = __ldg(r + 1);
            S.xyz[tid * 3 + 1] = __ldg(r + 2);
            S.xyz[tid * 3 + 2] = __ldg(r + 3);
            if (want_arg) S.kept[tid] = none_dropped ? row : a.orig2kept[row];
        }
        __syncthreads();
        // phase B (whole pillars only; the giant path fills mean/cen slot 0 itself)
        if (tid < nb) {
            const int b0 = (tid == 0 ? s : S.ends[tid - 1]) - s, b1 = S.ends[tid] - s;
            double sx = 0.0, sy = 0.0, sz = 0.0;
            for (int j = b0; j < b1; ++j) { sx += (double)S.xyz[j * 3]; sy += (double)S.xyz[j * 3 + 1]; sz += (double)S.xyz[j * 3 + 2]; }
            const double cnt = (double)(b1 - b0);
            const float mx = (float)__ddiv_rn(sx, cnt), my = (float)__ddiv_rn(sy, cnt), mz = (float)__ddiv_rn(sz, cnt);
            S.mean[tid * 3] = mx; S.mean[tid * 3 + 1] = my; S.mean[tid * 3 + 2] = mz;
            const int32_t *co = a.coords + (size_t)(p + tid) * a.coord_cols + (a.coord_cols - 2);
            const int cy = co[0], cx = co[1];
            // (:215-216) cx.float()*voxel_x + x_offset : separate mul and add roundings
            S.cen[tid * 2] = __fadd_rn(__fmul_rn((float)cx, a.vsz[0]), a.off[0]);
            S.cen[tid * 2 + 1] = __fadd_rn(__fmul_rn((float)cy, a.vsz[1]), a.off[1]);
            if (MODE == PFN_MODE_APPLY && a.pillar_mean) {
                float *pm = a.pillar_mean + (size_t)(p + tid) * 3;
                pm[0] = mx; pm[1] = my; pm[2] = mz;
            }
        }
        __syncthreads();
        // phase C1
        if (tid < np) {
            float r[Cfg::COLS], f[CS];
            load_row<Cfg>(a.pts, S.row[tid], r);
            const int lp = S.lp[tid];
            decorate<Cfg>(r, S.cen[lp * 2], S.cen[lp * 2 + 1], &S.mean[lp * 3], a, f);
            float *dst = &S.f[tid * Cfg::FSTRIDE];
#pragma unroll
            for (int k = 0; k < CS; ++k) dst[k] = f[k];
        }
        __syncthreads();
        // phase C2: rows grp + GROUPS*r, channels quad*4..+3
#pragma unroll
        for (int r = 0; r < PT; ++r) {
            const int j = grp + GROUPS * r;
            if (j < np) {  // warp-uniform up to the last partial group
                float f[Cfg::CSP4];
                const float4 *src = reinterpret_cast<const float4 *>(&S.f[j * Cfg::FSTRIDE]);
#pragma unroll
                for (int k4 = 0; k4 < Cfg::CSP4 / 4; ++k4) {
                    const float4 v = src[k4];
                    f[k4 * 4] = v.x; f[k4 * 4 + 1] = v.y; f[k4 * 4 + 2] = v.z; f[k4 * 4 + 3] = v.w;
                }
                float o[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float acc = 0.0f;
#pragma unroll
                    for (int k = 0; k < CS; ++k) acc = fmaf(W[c][k], f[k], acc);
                    if (MODE == PFN_MODE_STATS) o[c] = acc;
                    else {
                        const float y = fmaf(acc, sc4[c], sh4[c]);
                        o[c] = want_arg ? fmaxf(y, 0.0f) : y;  // eval: ReLU folds into the max with 0
                    }
                }
                *reinterpret_cast<float4 *>(&S.z[j * Cfg::ZSTRIDE + quad * 4]) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
    };

    // STATS: fold the batch in smem into the per-thread fp64 accumulators.
    auto accumulate_stats = [&](int np) {
        {
            const int c = tid % COUT, part = tid / COUT;
            constexpr int PARTS = kPfnThreads / COUT;
            for (int j = part; j < np; j += PARTS) {
                const double v = (double)S.z[j * Cfg::ZSTRIDE + c];
                st_x += v;
                st_x2 = fma(v, v, st_x2);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (pa[u] >= 0 && (u == 0 || NPAIR > kPfnThreads)) {
                // fp64 throughout: tight clusters make var << mean^2 and the backward's (S2 w - mu S1) cancels
                double acc = 0.0;
                if (pb[u] >= 0) {
                    for (int j = 0; j < np; ++j)
                        acc = fma((double)S.f[j * Cfg::FSTRIDE + pa[u]], (double)S.f[j * Cfg::FSTRIDE + pb[u]], acc);
                } else {
                    for (int j = 0; j < np; ++j) acc += (double)S.f[j * Cfg::FSTRIDE + pa[u]];
                }
                st_m[u] += acc;
            }
        }
        __syncthreads();
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p_lo = a.tile_start[tile], p_hi = a.tile_start[tile + 1];
        int p = p_lo;
        while (p < p_hi) {
            const int s = (p == 0) ? 0 : a.ends[p - 1];
            const int e_mine = (tid < BATCH && p + tid < p_hi) ? a.ends[p + tid] : 0x7fffffff;
            const int fits = (e_mine != 0x7fffffff) && (e_mine - s <= BATCH);
            const int nb = __syncthreads_count(fits);
            if (nb > 0) {
                if (tid < nb) S.ends[tid] = e_mine;
                __syncthreads();
                const int np = S.ends[nb - 1] - s;
                run_points(s, np, p, nb);
                if (MODE == PFN_MODE_STATS) {
                    accumulate_stats(np);
                } else {
                    // phase D
                    for (int item = tid; item < nb * QUADS; item += kPfnThreads) {
                        const int q = item / QUADS, qd = item % QUADS;
                        const int b0 = (q == 0 ? s : S.ends[q - 1]) - s, b1 = S.ends[q] - s;
                        if (!want_arg) {
                            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int j = b0; j < b1; ++j) {
                                const float4 v = *reinterpret_cast<const float4 *>(&S.z[j * Cfg::ZSTRIDE + qd * 4]);
                                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                            }
                            *reinterpret_cast<float4 *>(a.features + (size_t)(p + q) * COUT + qd * 4) = m;
                        } else {
                            float m[4] = {-1.f, -1.f, -1.f, -1.f};
                            int mi[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
                            for (int j = b0; j < b1; ++j) {
                                const float4 v4 = *reinterpret_cast<const float4 *>(&S.z[j * Cfg::ZSTRIDE + qd * 4]);
                                const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                                const int kj = S.kept[j];
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    if (v[c] > m[c] || (v[c] == m[c] && kj < mi[c])) { m[c] = v[c]; mi[c] = kj; }
                            }
                            *reinterpret_cast<float4 *>(a.features + (size_t)(p + q) * COUT + qd * 4) = make_float4(m[0], m[1], m[2], m[3]);
                            *reinterpret_cast<int4 *>(a.argmax + (size_t)(p + q) * COUT + qd * 4) = make_int4(mi[0], mi[1], mi[2], mi[3]);
                        }
                    }
                    __syncthreads();
                }
                p += nb;
            } else {
                // ---- giant pillar: more than BATCH points in pillar p
                const int e = a.ends[p];
                double sx = 0.0, sy = 0.0, sz = 0.0;
                for (int g = s + tid; g < e; g += kPfnThreads) {
                    const float *r = a.pts + (long long)a.order[g] * Cfg::COLS;
                    sx += (double)__ldg(r + 1); sy += (double)__ldg(r + 2); sz += (double)__ldg(r + 3);
                }
                S.red[tid * 3] = sx; S.red[tid * 3 + 1] = sy; S.red[tid * 3 + 2] = sz;
                __syncthreads();
                if (tid < 3) {
                    // fp64 adds of fp32 values in this range are exact, so the order is immaterial
                    double t = 0.0;
                    for (int j = 0; j < kPfnThreads; ++j) t += S.red[j * 3 + tid];
                    const float m = (float)__ddiv_rn(t, (double)(e - s));
                    S.mean[tid] = m;
                    if (MODE == PFN_MODE_APPLY && a.pillar_mean) a.pillar_mean[(size_t)p * 3 + tid] = m;
                }
                if (tid == 32) {
                    const int32_t *co = a.coords + (size_t)p * a.coord_cols + (a.coord_cols - 2);
                    S.cen[0] = __fadd_rn(__fmul_rn((float)co[1], a.vsz[0]), a.off[0]);
                    S.cen[1] = __fadd_rn(__fmul_rn((float)co[0], a.vsz[1]), a.off[1]);
                }
                if (tid < COUT) { S.carry_v[tid] = want_arg ? -1.0f : 0.0f; S.carry_i[tid] = 0x7fffffff; }
                __syncthreads();
                for (int cs = s; cs < e; cs += BATCH) {
                    const int np = min(BATCH, e - cs);
                    run_points(cs, np, p, 0);
                    if (MODE == PFN_MODE_STATS) {
                        accumulate_stats(np);
                    } else {
                        // column reduce: thread = (row group g, quad qd)
                        const int qd = tid % QUADS, g = tid / QUADS;
                        float m[4];
                        int mi[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) { m[c] = want_arg ? -1.0f : 0.0f; mi[c] = 0x7fffffff; }
                        for (int j = g; j < np; j += GROUPS) {
                            const float4 v4 = *reinterpret_cast<const float4 *>(&S.z[j * Cfg::ZSTRIDE + qd * 4]);
                            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                            const int kj = want_arg ? S.kept[j] : 0;
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (v[c] > m[c] || (want_arg && v[c] == m[c] && kj < mi[c])) { m[c] = v[c]; mi[c] = kj; }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) { S.part_v[g * COUT + qd * 4 + c] = m[c]; S.part_i[g * COUT + qd * 4 + c] = mi[c]; }
                        __syncthreads();
                        if (tid < COUT) {
                            float bm = S.carry_v[tid];
                            int bi = S.carry_i[tid];
                            for (int gg = 0; gg < GROUPS; ++gg) {
                                const float v = S.part_v[gg * COUT + tid];
                                const int vi = S.part_i[gg * COUT + tid];
                                if (v > bm || (want_arg && v == bm && vi < bi)) { bm = v; bi = vi; }
                            }
                            S.carry_v[tid] = bm; S.carry_i[tid] = bi;
                        }
                        __syncthreads();
                    }
                }
                if (MODE == PFN_MODE_APPLY && tid < COUT) {
                    a.features[(size_t)p * COUT + tid] = S.carry_v[tid];
                    if (want_arg) a.argmax[(size_t)p * COUT + tid] = S.carry_i[tid];
                }
                __syncthreads();
                p += 1;
            }
        }
    }

    if (MODE == PFN_MODE_STATS) {
        // per-CTA partials: [sum x (COUT) | sum x^2 (COUT) | S2 upper + S1 (NPAIR)]
        constexpr int PARTS = kPfnThreads / COUT;
        double *red = S.red;  // kPfnThreads*3 doubles
        red[tid] = st_x; red[kPfnThreads + tid] = st_x2;
        __syncthreads();
        double *out = a.partials + (size_t)blockIdx.x * (2 * COUT + NPAIR);
        if (tid < COUT) {
            double sx = 0.0, sx2 = 0.0;
            for (int q = 0; q < PARTS; ++q) { sx += red[q * COUT + tid]; sx2 += red[kPfnThreads + q * COUT + tid]; }
            out[tid] = sx; out[COUT + tid] = sx2;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int item = tid + u * kPfnThreads;
            if (item < NPAIR) out[2 * COUT + item] = st_m[u];
        }
    }
}

// ------------------------------------------------------------------------------------------- BN finalize (train)
// One CTA.  Reduces the per-CTA partials in a fixed order (deterministic), folds the batch statistics into
// scale/shift, updates the running statistics, and expands the feature moments for the backward.
// bn_state = [mean(COUT) | var(COUT) | scale(COUT) | shift(COUT) | n | S1(CIN) | S2(CIN*CIN)]
template <class Cfg>
__global__ void __launch_bounds__(kPfnThreads) bn_finalize_kernel(const __grid_constant__ PfnArgs a, int nblocks, double *bn_state,
                                                                 float *running_mean, float *running_var, double momentum) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, NPAIR = CS * (CS + 1) / 2 + CS, TOT = 2 * COUT + NPAIR;
    __shared__ double tot[TOT];
    const int tid = threadIdx.x;
    const long long N = a.counters[RDP_CNT_N];
    const int ntiles = (int)((N + kPfnTileRows - 1) / kPfnTileRows);
    const int used = min(nblocks, ntiles);
    for (int e = tid; e < TOT; e += kPfnThreads) {
        double s = 0.0;
        for (int b = 0; b < used; ++b) s += a.partials[(size_t)b * TOT + e];
        tot[e] = s;
    }
    __syncthreads();
    const int cin = a.c_in;
    if (tid < COUT) {
        const double n = (double)N;
        double mean = 0.0, var = 0.0;
        if (N > 0) {
            mean = __ddiv_rn(tot[tid], n);
            var = __dsub_rn(__ddiv_rn(tot[COUT + tid], n), __dmul_rn(mean, mean));
            if (!(var > 0.0)) var = 0.0;
        }
        bn_state[tid] = mean;
        bn_state[COUT + tid] = var;
        float sc, sh;
        fold_bn((double)a.gamma[tid], (double)a.beta[tid], mean, var, a.eps, &sc, &sh);
        bn_state[2 * COUT + tid] = (double)sc;
        bn_state[3 * COUT + tid] = (double)sh;
        if (N > 1) {  // torch raises for N == 1 and leaves the buffers alone for N == 0
            const double m = momentum;
            running_mean[tid] = (float)((1.0 - m) * (double)running_mean[tid] + m * mean);
            running_var[tid] = (float)((1.0 - m) * (double)running_var[tid] + m * var * (n / (n - 1.0)));
        }
    }
    if (tid == 0) bn_state[4 * COUT] = (double)N;
    double *S1 = bn_state + 4 * COUT + 1, *S2 = S1 + cin;
    for (int e = tid; e < NPAIR; e += kPfnThreads) {
        if (e < CS * (CS + 1) / 2) {
            int aa = 0, rem = e;
            while (rem >= CS - aa) { rem -= CS - aa; ++aa; }
            const int ka = a.kmap[aa], kb = a.kmap[aa + rem];
            if (ka >= 0 && kb >= 0) { S2[ka * cin + kb] = tot[2 * COUT + e]; S2[kb * cin + ka] = tot[2 * COUT + e]; }
        } else {
            const int ka = a.kmap[e - CS * (CS + 1) / 2];
            if (ka >= 0) S1[ka] = tot[2 * COUT + e];
        }
    }
}

// ------------------------------------------------------------------------------------------- backward
// warp = pillar, lane = channel (+32).  For every (pillar, channel): route g to the argmax row when the
// output is positive (ReLU'), rebuild that row's features and pre-activation, and accumulate
//   dbeta_c += gy,  G_c += gy * x_lin,  A_ck += gy * f_k.
// per-CTA partials (fp64): [dbeta(COUT) | G(COUT) | A(COUT*CIN)]
template <class Cfg>
__global__ void __launch_bounds__(kPfnThreads) pfn_bwd_kernel(const __grid_constant__ PfnArgs a, const float *__restrict__ grad,
                                                            const float *__restrict__ feat_out, const int32_t *__restrict__ arg,
                                                            const float *__restrict__ pmean) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, CPL = COUT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *red = reinterpret_cast<double *>(smem_raw);  // (kPfnThreads/32) * COUT * (CS+2)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.counters[RDP_CNT_P];
    const bool none_dropped = ((long long)a.counters[RDP_CNT_N] == a.n0);
    float W[CPL][CS];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int k = a.kmap[s];
            W[cc][s] = (k >= 0) ? __ldg(a.weight + (lane + 32 * cc) * a.c_in + k) : 0.0f;
        }
    double dA[CPL][CS], dB[CPL], dG[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
        dB[cc] = dG[cc] = 0.0;
#pragma unroll
        for (int s = 0; s < CS; ++s) dA[cc][s] = 0.0;
    }
    const int nwarps = gridDim.x * (kPfnThreads / 32);
    for (int p = blockIdx.x * (kPfnThreads / 32) + warp; p < P; p += nwarps) {
        const int32_t *co = a.coords + (size_t)p * a.coord_cols + (a.coord_cols - 2);
        const float cenx = __fadd_rn(__fmul_rn((float)co[1], a.vsz[0]), a.off[0]);
        const float ceny = __fadd_rn(__fmul_rn((float)co[0], a.vsz[1]), a.off[1]);
        const float mean[3] = {pmean[(size_t)p * 3], pmean[(size_t)p * 3 + 1], pmean[(size_t)p * 3 + 2]};
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            const size_t o = (size_t)p * COUT + lane + 32 * cc;
            const float out = feat_out[o];
            const float gy = out > 0.0f ? grad[o] : 0.0f;
            const int kj = arg[o];
            const int row = none_dropped ? kj : a.kept2orig[kj];
            float r[Cfg::COLS], f[CS];
            load_row<Cfg>(a.pts, row, r);
            decorate<Cfg>(r, cenx, ceny, mean, a, f);
            float x = 0.0f;
#pragma unroll
            for (int k = 0; k < CS; ++k) x = fmaf(W[cc][k], f[k], x);
            const double g = (double)gy;
            dB[cc] += g;
            dG[cc] = fma(g, (double)x, dG[cc]);
#pragma unroll
            for (int k = 0; k < CS; ++k) dA[cc][k] = fma(g, (double)f[k], dA[cc][k]);
        }
    }
    // reduce the warps of this CTA
    constexpr int PER = CS + 2;
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
        double *dst = red + ((size_t)warp * COUT + lane + 32 * cc) * PER;
        dst[0] = dB[cc]; dst[1] = dG[cc];
#pragma unroll
        for (int k = 0; k < CS; ++k) dst[2 + k] = dA[cc][k];
    }
    __syncthreads();
    double *out = a.partials + (size_t)blockIdx.x * COUT * PER;
    for (int e = tid; e < COUT * PER; e += kPfnThreads) {
        double s = 0.0;
        for (int w = 0; w < kPfnThreads / 32; ++w) s += red[(size_t)w * COUT * PER + e];
        out[e] = s;
    }
}

// One CTA: fixed-order reduction of the partials, then the closed-form BatchNorm backward:
//   dgamma_c = (G_c - mu_c dbeta_c) / sigma_c
//   dW_ck    = (gamma_c/sigma_c) [ A_ck - dbeta_c/N S1_k - dgamma_c/N ((S2 w_c)_k - mu_c S1_k)/sigma_c ]   (train)
//   dW_ck    = (gamma_c/sigma_c) A_ck                                                                    (eval BN)
template <class Cfg>
__global__ void __launch_bounds__(kPfnThreads) bwd_finalize_kernel(const __grid_constant__ PfnArgs a, int nblocks, const double *bn_state,
                                                                  int train_bn, float *d_weight, float *d_gamma, float *d_beta) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, PER = CS + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tot = reinterpret_cast<double *>(smem_raw);  // COUT*PER
    double *dgam = tot + COUT * PER;                       // COUT
    const int tid = threadIdx.x, cin = a.c_in;
    for (int e = tid; e < COUT * PER; e += kPfnThreads) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += a.partials[(size_t)b * COUT * PER + e];
        tot[e] = s;
    }
    __syncthreads();
    const double n = (double)a.counters[RDP_CNT_N];
    auto stat = [&](int c, double *mu, double *inv_std) {
        if (!a.use_norm) { *mu = 0.0; *inv_std = 1.0; return; }
        const double m = train_bn ? bn_state[c] : (double)a.rmean[c];
        const double v = train_bn ? bn_state[COUT + c] : (double)a.rvar[c];
        *mu = m;
        *inv_std = 1.0 / sqrt(v + a.eps);
    };
    if (tid < COUT) {
        double mu, is;
        stat(tid, &mu, &is);
        const double db = tot[tid * PER], G = tot[tid * PER + 1];
        const double dg = (G - mu * db) * is;
        dgam[tid] = dg;
        d_beta[tid] = (float)db;
        if (a.use_norm && d_gamma) d_gamma[tid] = (float)dg;
    }
    __syncthreads();
    const double *S1 = bn_state ? bn_state + 4 * COUT + 1 : nullptr, *S2 = S1 ? S1 + cin : nullptr;
    for (int e = tid; e < COUT * CS; e += kPfnThreads) {
        const int c = e / CS, s = e % CS, k = a.kmap[s];
        if (k < 0) continue;
        double mu, is;
        stat(c, &mu, &is);
        const double gam = a.use_norm ? (double)a.gamma[c] : 1.0;
        double v = tot[c * PER + 2 + s];
        if (a.use_norm && train_bn && n > 0) {
            double s2w = 0.0;
            for (int j = 0; j < cin; ++j) s2w += S2[k * cin + j] * (double)a.weight[c * cin + j];
            const double db = tot[c * PER], dg = dgam[c];
            v = v - db / n * S1[k] - dg / n * (s2w - mu * S1[k]) * is;
        }
        d_weight[c * cin + k] = (float)(gam * is * v);
    }
}

}  // namespace rdp
