// rdp_pfn.cuh -- device code of the fused pillar feature network (shared by the per-config .cu files).
//
// Replaces /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:214-240 and
// PFNLayerV2.forward :35-46 (last layer): scatter_mean, f_center / f_cluster / f_relative, concat,
// Linear(no bias)+BatchNorm1d+ReLU, scatter_max (+argmax), and the autograd of that chain.
//
// Canonical arithmetic ("folded"; oracle: ORC_FOLDED).  Every decorated feature of the reference is
// affine in the row's centre offsets and in per-pillar constants:
//     f_center = d,  xyz = d + centre,  f_cluster = xyz - mean = d + (centre - mean),  f_rel = d + (centre - lo)
// with d = xyz - centre (exactly the reference's f_center, :215-217).  Hence, with T the (c_in x (G+1)) matrix that
// writes the layout's features in the reduced basis g = [d, raw features, (dist) | centre xy, centre - mean xyz] (+1):
//     x_c = W_c . f = wg_c . g + const_c,   [wg_c | const_c] = W_c T   (fp64, one rounding)
//         = v_ic + u_pc,   v = k-ascending fmaf chain over the KIN row inputs,  u = fmaf chain over the 5 pillar constants
// KIN = C (+1) multiplies per row and channel instead of C + 9, all on small, well-conditioned operands; the pillar term
// costs 5 per (pillar, channel).  BatchNorm is folded to y = fma(x, scale, shift), z = max(y, 0).  Because fl(v + u) and
// fma(., scale, shift) are monotone in v, max_i z_i = z(max_i s v_i) with s = sign(scale): the row loop is KIN FMAs and
// one max per channel; the add, BN and ReLU run once per pillar.
// argmax rule (train forward): the row with the largest s*v -- the row that maximises the BatchNorm output -- lowest
// kept-point index on exact ties; a pillar whose maximum the ReLU clamped to 0 reports its lowest kept-point index
// (every row ties at 0; torch_scatter's CPU rule) and gets no gradient.
//
// Input is the pillar-grouped row array written by group_rows_kernel and the pillar table written by the table kernel
// (rdp_table.cuh): rows of one pillar are contiguous, pillars are in key order.  One persistent CTA (128 threads) takes
// tiles b, b + grid, ...; tile t owns the pillars that START in grouped rows [128 t, 128 t + 128) and stages rows
// [128 t - 1, 128 t + 192) plus the table slice of its pillars -- fixed-size, 16-byte aligned windows -- one tile ahead
// with two 1-D TMA bulk copies onto an mbarrier, double buffered.  Per tile a thread = row phase turns the staged rows into
// row records (row inputs, last-row flag, grouped position; in the train forward sorted by original row id inside each
// pillar), then each warp streams a pillar-aligned quarter of the records (lane = channel): no intermediate feature tile,
// two block barriers per tile.  A pillar that runs past the staged window (> 64 rows of overhang) is streamed from global
// memory by all warps.
#pragma once

#include "rdp_table.cuh"

namespace rdp {

struct PfnArgs {
    const float *grows;
    const float *aux;               // per-pillar table [centre xy | centre - mean xyz | first grouped row | rows | centre z]
    const int32_t *tile_first;      // first pillar starting at or after grouped row 128 t
    const int32_t *starts;
    int32_t *counters;
    const int32_t *orig2kept;
    const float *weight, *bias, *gamma, *beta;
    float *rmean, *rvar;
    double *bn_state;               // train: [mean | var | scale | shift | n | S1(G) | S2(G*G)]
    float *features;
    int32_t *argpos;         // winning GROUPED position per (pillar, channel); -1 when the ReLU clamped the pillar's maximum
    const float *grad;       // backward input: upstream gradient (P, Cout)
    double *acc_stats, *acc_bwd;
    const double *local_stats;   // SyncBatchNorm: this rank's moments (acc_stats then holds the all-reduced ones); else null
    float *d_weight, *d_gamma, *d_beta;
    long long *num_batches_tracked;
    long long n0;
    int n_from_totals;       // 1: the number of points the statistics cover is totals[NACC] (SyncBatchNorm: all ranks), else counters[N]
    double eps, momentum;
    float off_z;
    int c_in;
    int use_norm;
    int train_bn;            // scale / shift from bn_state (batch statistics) instead of the running statistics
    int defer_finalize;      // moments pass: leave the totals in acc_stats (SyncBatchNorm all-reduces them first)
    float T[kMaxCin][kMaxG + 1];   // feature j of the layout = sum_m T[j][m] * [g | 1][m]
};

// Compile-time shape of one encoder family: row width, distance feature, output channels.  The layout (Simple2D /
// DynamicPillarVFE order, USE_ABSLOTE_XYZ / USE_CLUSTER_XYZ / USE_RELATIVE_XYZ) only enters through T.
template <int COLS_, bool DIST_, int COUT_>
struct PfnCfg {
    static constexpr int COLS = COLS_, COUT = COUT_;
    static constexpr bool DIST = DIST_;
    using B = RowBasis<COLS_, DIST_>;
    static constexpr int KIN = B::KIN, G = B::G, NACC = B::NACC, RS = B::RS;
    static constexpr int CPL = COUT / 32;                              // channels per lane
    static constexpr int BWD_PER = G + 1;                              // per channel: A(G) | dbeta
    static constexpr int FW = (KIN + 2 + 3) / 4 * 4;                   // floats per row record in shared memory: row inputs | last-row flag | grouped position
    static constexpr uint32_t ROW_BYTES = (kPfnCap + 1) * RS * 4;      // the window plus the row in front of it
    static constexpr uint32_t AUX_BYTES = kPfnWin * 8 * 4;
    static_assert(G <= kMaxG && COUT % 32 == 0 && COUT <= kMaxCout, "shape");
    static_assert(ROW_BYTES % 16 == 0, "TMA sizes");
};

// One staged tile: grouped rows [128 t - 1, 128 t + 192); row w of the window is rows[(w + 1) * RS ...].
// Slot RS-2 of a row holds its original row index, slot RS-1 its pillar id (int bit patterns).
template <class Cfg>
struct PfnStage {
    alignas(32) float aux[kPfnWin * 8];   // table entries of the (at most 128) pillars that start in the window
    alignas(16) float rows[(kPfnCap + 1) * Cfg::RS];
    __device__ __forceinline__ int gid(int w) const { return __float_as_int(rows[(w + 1) * Cfg::RS + Cfg::RS - 1]); }
    __device__ __forceinline__ int orig(int w) const { return __float_as_int(rows[(w + 1) * Cfg::RS + Cfg::RS - 2]); }
    __device__ __forceinline__ const float *row(int w) const { return rows + (w + 1) * Cfg::RS; }
};

// BatchNorm folded to y = fma(x, scale, shift) in fp64 with one rounding (oracle: orc_bn_fold).
__device__ __forceinline__ void fold_bn(double gamma, double beta, double mean, double var, double eps, float *scale, float *shift) {
    const double inv_std = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(var, eps)));
    const double s = __dmul_rn(gamma, inv_std);
    *scale = (float)s;
    *shift = (float)__dsub_rn(beta, __dmul_rn(mean, s));
}

// [wg_c | const_c] = W_c T (+ bias): fp64 products (exact) and sums in feature order, rounded once to fp32.
template <int G>
__device__ __forceinline__ void fold_channel(const PfnArgs &a, int c, float *wg, float *cst) {
    double acc[G + 1];
#pragma unroll
    for (int m = 0; m <= G; ++m) acc[m] = 0.0;
    for (int j = 0; j < a.c_in; ++j) {
        const double w = (double)__ldg(a.weight + c * a.c_in + j);
#pragma unroll
        for (int m = 0; m <= G; ++m) acc[m] = __dadd_rn(acc[m], __dmul_rn(w, (double)a.T[j][m]));
    }
    if (a.bias) acc[G] = __dadd_rn(acc[G], (double)__ldg(a.bias + c));
#pragma unroll
    for (int m = 0; m < G; ++m) wg[m] = (float)acc[m];
    *cst = (float)acc[G];
}

// scale / shift of channel c for the apply kernels
__device__ __forceinline__ void channel_affine(const PfnArgs &a, int cout, int c, float *sc, float *sh) {
    *sc = 1.0f;
    *sh = 0.0f;
    if (a.use_norm) {
        if (a.train_bn) { *sc = (float)a.bn_state[2 * cout + c]; *sh = (float)a.bn_state[3 * cout + c]; }
        else fold_bn((double)a.gamma[c], (double)a.beta[c], (double)a.rmean[c], (double)a.rvar[c], a.eps, sc, sh);
    }
}

// ------------------------------------------------------------------------------------------- tile machinery
template <class Cfg, int PCH>
struct TileSmem {
    PfnStage<Cfg> st[2];
    alignas(8) uint64_t full[2];
    alignas(8) uint64_t pre[2];                    // BWD: arrival of a chunk's (grad, argpos) rows, double buffered
    alignas(16) float pre_grad[2][PCH > 0 ? PCH * Cfg::COUT : 4];
    alignas(16) int pre_arg[2][PCH > 0 ? PCH * Cfg::COUT : 4];
    alignas(16) float frec[kPfnCap * Cfg::FW];     // per staged row: the KIN row inputs + last-row-of-pillar flag (tile_c1)
    int tf[2][2];                                  // per stage: first pillar of the tile, first pillar of the next
    float carry_v[4][Cfg::COUT];                   // big-pillar path: per-warp maxima, combined in warp order
    int carry_p[4][Cfg::COUT], carry_o[4][Cfg::COUT];
};

struct TileBounds {
    int ps, pe;        // pillars [ps, pe) start in this tile's window
    int j0, jstop;     // window rows [j0, jstop) belong to the pillars streamed from the staged rows
    int nb;            // number of those pillars
    bool big;          // the last pillar (pe - 1) runs past the staged rows: rows [big_a0, big_a0 + big_rows) from global memory
    long long big_a0;
    int big_rows;
};

template <class Cfg>
__device__ __forceinline__ TileBounds tile_bounds(const PfnStage<Cfg> &T, int ps, int pe, long long base) {
    TileBounds b;
    b.ps = ps; b.pe = pe;
    b.j0 = __float_as_int(T.aux[5]) - (int)base;                      // first row of the first pillar
    const float *alast = T.aux + (pe - ps - 1) * 8;
    const int last_start = __float_as_int(alast[5]) - (int)base, last_rows = __float_as_int(alast[6]);
    b.big = last_start + last_rows > kPfnCap;
    b.jstop = b.big ? last_start : last_start + last_rows;
    b.nb = b.big ? pe - ps - 1 : pe - ps;
    b.big_a0 = base + last_start;
    b.big_rows = last_rows;
    return b;
}

// ------------------------------------------------------------------------------------------- forward
// Per tile:  C1 (thread = row) turns every staged row into its KIN row inputs + a last-row-of-pillar flag in shared
// memory, so that the redundant per-lane work of the stream is two shared loads per row;  ST (lane = channel) walks the
// rows of a pillar-aligned quarter of the tile per warp in one flat loop: KIN FMAs and a max per row, the pillar epilogue
// (pillar term, BatchNorm, ReLU, one coalesced 128-byte store per 32 channels) when a row carries the flag.
// SORT (train forward): the records of a pillar are written in the order of their ORIGINAL row ids instead of the arrival
// order of the grouping pass, each carrying its grouped position.  The stream then needs no tie handling at all: "first
// strict maximum wins" is the documented tie rule (lowest kept-point index among equal maxima).  The rank of a row inside
// its pillar is a count over the pillar's staged rows -- 1 comparison for 2 of 3 pillars, ~6 on average for the others.
template <class Cfg, bool SORT>
__device__ __forceinline__ void tile_c1(const PfnStage<Cfg> &T, const TileBounds &tb, long long base, float off_z, float *frec) {
    constexpr int COLS = Cfg::COLS, RS = Cfg::RS, KIN = Cfg::KIN, FW = Cfg::FW;
    for (int j = tb.j0 + (int)threadIdx.x; j < tb.jstop; j += kPfnThreads) {
        const float *src = T.row(j);
        float row[RS], rin[FW];
#pragma unroll
        for (int c4 = 0; c4 < RS; c4 += 4) {   // whole row incl. the pillar id in its last slot
            const float4 q = *reinterpret_cast<const float4 *>(src + c4);
            row[c4] = q.x; row[c4 + 1] = q.y; row[c4 + 2] = q.z; row[c4 + 3] = q.w;
        }
        const int gid = __float_as_int(row[RS - 1]);
        const float *ax = &T.aux[(gid - tb.ps) * 8];
        const float2 cen = *reinterpret_cast<const float2 *>(ax);
        row_inputs<COLS, Cfg::DIST>(row, cen.x, cen.y, off_z, rin);
        int slot = j, last;
        if (SORT) {
            const int first = __float_as_int(ax[5]) - (int)base, n = __float_as_int(ax[6]);
            const int o = __float_as_int(row[RS - 2]);
            int idx = 0;
            for (int r = first; r < first + n; ++r) idx += (T.orig(r) < o) ? 1 : 0;   // original row ids are unique
            slot = first + idx;
            last = (idx == n - 1);
        } else {
            last = (j == tb.jstop - 1) || (T.gid(j + 1) != gid);
        }
        rin[KIN] = __int_as_float(last);
        rin[KIN + 1] = __int_as_float((int)base + j);
#pragma unroll
        for (int k = KIN + 2; k < FW; ++k) rin[k] = 0.0f;
        float4 *dst = reinterpret_cast<float4 *>(frec + slot * FW);
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) dst[k4] = make_float4(rin[4 * k4], rin[4 * k4 + 1], rin[4 * k4 + 2], rin[4 * k4 + 3]);
    }
}

template <class Cfg, bool ARG>
__global__ void __launch_bounds__(kPfnThreads, Cfg::CPL == 1 ? (ARG ? 6 : 7) : (Cfg::CPL == 2 ? 4 : 2)) pfn_apply_kernel(const __grid_constant__ PfnArgs a) {
    pdl_wait();
    extern __shared__ __align__(32) unsigned char smem_raw[];
    using Smem = TileSmem<Cfg, 0>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, COLS = Cfg::COLS, RS = Cfg::RS, CPL = Cfg::CPL, KIN = Cfg::KIN, G = Cfg::G, FW = Cfg::FW;
    constexpr int WIN = kPfnWin, NW = kPfnThreads / 32;
    constexpr bool DIST = Cfg::DIST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = a.counters[RDP_CNT_N];
    const int ntiles = (int)((N + WIN - 1) / WIN);
    // Tile order: CTA b takes tiles b, b + grid, b + 2 grid, ... -- at any moment the grid works on one contiguous window
    // of the row array and of the outputs instead of one private stream per CTA.
    const int Gd = (int)gridDim.x;
    const int nk = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + Gd - 1) / Gd : 0;   // tiles of this CTA
    auto tile_of = [&](int k) { return (int)blockIdx.x + k * Gd; };

    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        fence_mbar_init();
    }
    // lane = channel (+ 32 cc): folded weights, sign-flipped so that the running maximum of the flipped v is the row that
    // maximises the BatchNorm output
    float wv[CPL][KIN], wq[CPL][5], cst[CPL], sc[CPL], sh[CPL], sgn[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
        const int ch = lane + 32 * cc;
        float wg[G];
        fold_channel<G>(a, ch, wg, &cst[cc]);
        channel_affine(a, COUT, ch, &sc[cc], &sh[cc]);
        sgn[cc] = sc[cc] < 0.0f ? -1.0f : 1.0f;
#pragma unroll
        for (int k = 0; k < KIN; ++k) wv[cc][k] = sgn[cc] * wg[k];
#pragma unroll
        for (int k = 0; k < 5; ++k) wq[cc][k] = wg[KIN + k];
    }
    __syncthreads();

    // thread 0 only.  (pf, pn) = tile_first[t], tile_first[t + 1]; published to the CTA through S.tf before the arrive.
    auto issue = [&](int t, int s, int pf, int pn) {
        PfnStage<Cfg> &T = S.st[s];
        S.tf[s][0] = pf; S.tf[s][1] = pn;
        mbar_expect_tx(&S.full[s], Cfg::ROW_BYTES + Cfg::AUX_BYTES);
        tma_bulk_g2s(T.rows, a.grows + (size_t)t * WIN * RS, Cfg::ROW_BYTES, &S.full[s]);  // grows row 0 = the row before position 0
        tma_bulk_g2s(T.aux, a.aux + (size_t)pf * 8, Cfg::AUX_BYTES, &S.full[s]);
    };

    // v[cc] = flipped linear value of the row inputs (k-ascending fmaf chain)
    auto chain = [&](const float *rin, float *v) {
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            float acc = __fmul_rn(wv[cc][0], rin[0]);
#pragma unroll
            for (int k = 1; k < KIN; ++k) acc = fmaf(wv[cc][k], rin[k], acc);
            v[cc] = acc;
        }
    };
    auto load_rec = [&](const float *rec, float *rin) {
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
            const float4 q = *reinterpret_cast<const float4 *>(rec + 4 * k4);
            rin[4 * k4] = q.x; rin[4 * k4 + 1] = q.y; rin[4 * k4 + 2] = q.z; rin[4 * k4 + 3] = q.w;
        }
    };
    // big-pillar path: row straight from global memory
    auto row_v_global = [&](const float *src, float cenx, float ceny, float *v) {
        float row[RS], rin[KIN];
#pragma unroll
        for (int c4 = 0; c4 < (COLS + 3) / 4 * 4; c4 += 4) {
            const float4 q = *reinterpret_cast<const float4 *>(src + c4);
            row[c4] = q.x; row[c4 + 1] = q.y; row[c4 + 2] = q.z; row[c4 + 3] = q.w;
        }
        row_inputs<COLS, DIST>(row, cenx, ceny, a.off_z, rin);
        chain(rin, v);
    };
    // pillar epilogue: x = s m + u (one rounding), BatchNorm, ReLU
    auto pillar_z = [&](const float *m, const float4 &a0, float ndz, float *z) {
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            float u = cst[cc];
            u = fmaf(wq[cc][0], a0.x, u);
            u = fmaf(wq[cc][1], a0.y, u);
            u = fmaf(wq[cc][2], a0.z, u);
            u = fmaf(wq[cc][3], a0.w, u);
            u = fmaf(wq[cc][4], ndz, u);
            const float x = fmaf(sgn[cc], m[cc], u);
            z[cc] = fmaxf(fmaf(x, sc[cc], sh[cc]), 0.0f);
        }
    };
    const float NEG_INF = __int_as_float(0xff800000);

    // thread 0 keeps tile_first two tiles ahead in registers so the TMA issue never waits on a global load
    int nfa = 0, nfb = 0;
    if (nk > 0 && tid == 0) {
        issue(tile_of(0), 0, a.tile_first[tile_of(0)], a.tile_first[tile_of(0) + 1]);
        if (nk > 1) { nfa = a.tile_first[tile_of(1)]; nfb = a.tile_first[tile_of(1) + 1]; }
    }
    uint32_t par0 = 0, par1 = 0;

    for (int k = 0; k < nk; ++k) {
        const int t = tile_of(k);
        const int s = k & 1;
        if (tid == 0 && k + 1 < nk) {
            issue(tile_of(k + 1), s ^ 1, nfa, nfb);
            if (k + 2 < nk) { nfa = a.tile_first[tile_of(k + 2)]; nfb = a.tile_first[tile_of(k + 2) + 1]; }
        }
        if (s == 0) { mbar_wait(&S.full[0], par0); par0 ^= 1; } else { mbar_wait(&S.full[1], par1); par1 ^= 1; }
        const PfnStage<Cfg> &T = S.st[s];
        const long long base = (long long)t * WIN;
        const int ps = S.tf[s][0], pe = S.tf[s][1];
        if (pe == ps) { __syncthreads(); continue; }  // a pillar from an earlier tile covers the whole window
        const TileBounds tb = tile_bounds<Cfg>(T, ps, pe, base);
        const int np = tb.jstop - tb.j0;

        if (np > 0) {
            tile_c1<Cfg, ARG>(T, tb, base, a.off_z, S.frec);
            __syncthreads();
            // warp `warp` streams a pillar-aligned quarter of the rows
            auto cut = [&](int q) -> int {
                if (q <= 0) return tb.j0;
                if (q >= NW) return tb.jstop;
                const int w = tb.j0 + (q * np) / NW;
                return __float_as_int(T.aux[(T.gid(w) - ps) * 8 + 5]) - (int)base;
            };
            int w = cut(warp);
            const int wb = cut(warp + 1);
            if (w < wb) {
                const int slot0 = T.gid(w) - ps;
                const float *ax = T.aux + slot0 * 8;
                float *fout = a.features + (size_t)(ps + slot0) * COUT + lane;
                int32_t *aout = ARG ? a.argpos + (size_t)(ps + slot0) * COUT + lane : nullptr;
                float m[CPL];
                int mp[CPL];
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) { m[cc] = NEG_INF; mp[cc] = 0; }
                // one row folded into the running maxima; closes the pillar when the row carries the last-row flag.  ARG: the
                // records of a pillar arrive in original-row order (tile_c1<SORT>), so the first strict maximum is the winner
                // of the documented tie rule and `pos` is its grouped position.
                auto fold = [&](const float *v, int last, int pos) {
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) {
                        if (!ARG) {
                            m[cc] = fmaxf(m[cc], v[cc]);
                        } else if (v[cc] > m[cc]) {
                            m[cc] = v[cc]; mp[cc] = pos;
                        }
                    }
                    if (last) {
                        const float4 a0 = *reinterpret_cast<const float4 *>(ax);   // centre xy, (centre - mean) xy
                        float z[CPL];
                        pillar_z(m, a0, ax[4], z);
#pragma unroll
                        for (int cc = 0; cc < CPL; ++cc) {
                            fout[32 * cc] = z[cc];
                            if (ARG) aout[32 * cc] = (z[cc] > 0.0f) ? mp[cc] : -1;   // -1: ReLU clamped, no gradient
                            m[cc] = NEG_INF;
                        }
                        ax += 8; fout += COUT;
                        if (ARG) aout += COUT;
                    }
                };
                for (; w + 1 < wb; w += 2) {   // two rows in flight: independent load + fmaf chains
                    float r0[FW], r1[FW], v0[CPL], v1[CPL];
                    load_rec(S.frec + w * FW, r0);
                    load_rec(S.frec + (w + 1) * FW, r1);
                    chain(r0, v0);
                    chain(r1, v1);
                    fold(v0, __float_as_int(r0[KIN]), __float_as_int(r0[KIN + 1]));
                    fold(v1, __float_as_int(r1[KIN]), __float_as_int(r1[KIN + 1]));
                }
                if (w < wb) {
                    float r0[FW], v0[CPL];
                    load_rec(S.frec + w * FW, r0);
                    chain(r0, v0);
                    fold(v0, __float_as_int(r0[KIN]), __float_as_int(r0[KIN + 1]));
                }
            }
        }

        if (tb.big) {
            // ---- big pillar pb = grouped rows [a0, a0 + rows): every warp streams a quarter straight from global memory
            const int pb = pe - 1;
            const float *ax = T.aux + (pe - ps - 1) * 8;
            const float4 a0 = *reinterpret_cast<const float4 *>(ax);
            const float ndz = ax[4];
            const int per = (tb.big_rows + NW - 1) / NW;
            const long long ra = tb.big_a0 + (long long)warp * per;
            const long long rb = min(tb.big_a0 + tb.big_rows, ra + per);
            float m[CPL];
            int mp[CPL], mo[CPL];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) { m[cc] = NEG_INF; mp[cc] = 0; mo[cc] = 0x7fffffff; }
            for (long long i = ra; i < rb; ++i) {
                const float *src = a.grows + ((size_t)i + 1) * RS;
                float v0[CPL];
                row_v_global(src, a0.x, a0.y, v0);
                const int o = ARG ? __float_as_int(src[RS - 2]) : 0;
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    if (!ARG) m[cc] = fmaxf(m[cc], v0[cc]);
                    else if (v0[cc] > m[cc] || (v0[cc] == m[cc] && o < mo[cc])) { m[cc] = v0[cc]; mp[cc] = (int)i; mo[cc] = o; }
                }
            }
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                S.carry_v[warp][lane + 32 * cc] = m[cc];
                S.carry_p[warp][lane + 32 * cc] = mp[cc];
                S.carry_o[warp][lane + 32 * cc] = mo[cc];
            }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    const int ch = lane + 32 * cc;
                    for (int w2 = 1; w2 < NW; ++w2) {
                        const float v2 = S.carry_v[w2][ch];
                        const int o2 = S.carry_o[w2][ch];
                        if (v2 > m[cc] || (ARG && v2 == m[cc] && o2 < mo[cc])) { m[cc] = v2; mp[cc] = S.carry_p[w2][ch]; mo[cc] = o2; }
                    }
                }
                float z[CPL];
                pillar_z(m, a0, ndz, z);
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    a.features[(size_t)pb * COUT + lane + 32 * cc] = z[cc];
                    if (ARG) a.argpos[(size_t)pb * COUT + lane + 32 * cc] = (z[cc] > 0.0f) ? mp[cc] : -1;
                }
            }
        }
        __syncthreads();   // every warp is done with stage s (and the row records) before the next tile overwrites them
    }
}

// ------------------------------------------------------------------------------------------- BatchNorm epilogue (train)
// totals = [S1(G) | S2(G x G) | n] of the reduced basis over the n points the statistics cover.  Batch moments of
// x = wg . g + const (oracle: orc_bn_batch_stats_folded):  mean_c = wg_c . S1 / n + const_c,
// var_c = wg_c^T S2 wg_c / n - (wg_c . S1 / n)^2  (biased).  Called by one CTA.
// bn_state = [mean(COUT) | var(COUT) | scale(COUT) | shift(COUT) | n | S1(G) | S2(G*G)]
template <class Cfg>
__device__ __forceinline__ void bn_epilogue(const PfnArgs &a, const double *totals, double *sm /* G + G*G doubles */) {
    constexpr int COUT = Cfg::COUT, G = Cfg::G;
    const int tid = threadIdx.x;
    double *S1 = sm, *S2 = sm + G;
    const long long N = a.n_from_totals ? llrint(__ldcg(totals + Cfg::NACC)) : (long long)a.counters[RDP_CNT_N];
    for (int e = tid; e < G; e += blockDim.x) S1[e] = __ldcg(totals + e);
    for (int e = tid; e < G * G; e += blockDim.x) {   // the totals hold the upper triangle
        const int k = e / G, l = e % G;
        S2[e] = __ldcg(totals + G + (k <= l ? k * G + l : l * G + k));
    }
    __syncthreads();
    double *st = a.bn_state;
    for (int c = tid; c < COUT; c += blockDim.x) {
        float wg[G], cst;
        fold_channel<G>(a, c, wg, &cst);
        double m1 = 0.0, e2 = 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            m1 += (double)wg[k] * S1[k];
            double row = 0.0;
#pragma unroll
            for (int l = 0; l < G; ++l) row += S2[k * G + l] * (double)wg[l];
            e2 += (double)wg[k] * row;
        }
        const double n = (double)N;
        double mean = 0.0, var = 0.0;
        if (N > 0) {
            const double mc = __ddiv_rn(m1, n);
            mean = mc + (double)cst;
            var = __dsub_rn(__ddiv_rn(e2, n), __dmul_rn(mc, mc));
            if (!(var > 0.0)) var = 0.0;
        }
        st[c] = mean;
        st[COUT + c] = var;
        float sc, sh;
        fold_bn((double)a.gamma[c], (double)a.beta[c], mean, var, a.eps, &sc, &sh);
        st[2 * COUT + c] = (double)sc;
        st[3 * COUT + c] = (double)sh;
        if (N > 1) {  // torch raises for N == 1 and leaves the buffers alone for N == 0
            a.rmean[c] = (float)((1.0 - a.momentum) * (double)a.rmean[c] + a.momentum * mean);
            a.rvar[c] = (float)((1.0 - a.momentum) * (double)a.rvar[c] + a.momentum * var * (n / (n - 1.0)));
        }
    }
    if (tid == 0) {
        st[4 * COUT] = (double)N;
        if (a.num_batches_tracked) *a.num_batches_tracked += 1;   // BatchNorm1d counts every train-mode forward (:29)
    }
    // the backward needs THIS rank's S1 / S2 (they differ from `totals` only under SyncBatchNorm)
    if (a.local_stats) {
        __syncthreads();
        for (int e = tid; e < G; e += blockDim.x) S1[e] = a.local_stats[e];
        for (int e = tid; e < G * G; e += blockDim.x) {
            const int k = e / G, l = e % G;
            S2[e] = a.local_stats[G + (k <= l ? k * G + l : l * G + k)];
        }
        __syncthreads();
    }
    for (int e = tid; e < G + G * G; e += blockDim.x) st[4 * COUT + 1 + e] = sm[e];
}

// Pillar table + feature moments in one pass (train-mode BatchNorm; oracle: orc_moments_folded).  The moments
// S1 = sum g, S2 = sum g g^T of the reduced basis g = [row inputs r (KIN) | pillar constants q (5)] over all kept rows split
// by where their factors live:
//   row x row     sum_i r_i r_i^T                     thread-local fp64 accumulators while the thread walks its pillar's rows
//   row x pillar  sum_p (sum_{i in p} r_i) q_p^T      } per pillar, from the row sums the mean needs anyway: a rank-32 update
//   pillar^2      sum_p n_p q_p q_p^T, S1             } per warp batch of 32 pillars,  [rs | n q]^T [q | 1]  (13 x 6), on the fp64
//                                                       tensor pipe (mma.m8n8k4.f64) from per-warp staging in shared memory
// so the separate moments pass over the rows (and its dependent row -> table gathers) is gone.  Products of two fp32
// values are exact in fp64; the row sums are exact; the totals agree with the oracle's sequential fp64 sums to ~1e-13, so
// the folded fp32 scale / shift come out bit-identical in practice (train-mode tolerance: 1e-6).
// Per-CTA reduction, fp64 atomics into the workspace totals; the CTA that finishes last folds them into bn_state.
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <class Cfg>
struct StatMap {
    static constexpr int KIN = Cfg::KIN, G = Cfg::G;
    static constexpr int NA = 16;                             // A operand per pillar: [row sums (KIN) | n q (5) | 0...]
    static constexpr int NB = 8;                              // B operand per pillar: [q (5) | 1 | 0 0]
    static constexpr int NRR = KIN * (KIN + 1) / 2;           // row x row (upper triangle), thread-local when RR
    static_assert(KIN + 5 <= NA, "A operand rows");
    // element (a, b) of the 16 x 8 product -> slot of totals = [S1(G) | S2(G x G, upper triangle)], or -1
    __device__ static int dst(int a, int b) {
        if (a >= KIN + 5 || b > 5) return -1;
        if (b == 5) return a;                                               // S1: row sums | sum n q
        if (a < KIN) return G + a * G + KIN + b;                            // row x pillar
        return (a - KIN) <= b ? G + a * G + KIN + b : -1;                   // pillar^2, upper triangle
    }
    __device__ static int rr_dst(int e) {   // e-th element of the row x row upper triangle, row-major
        int k = 0;
        while (e >= KIN - k) { e -= KIN - k; ++k; }
        return G + k * G + k + e;
    }
};

constexpr int kTableStatsThreads = 128;

template <class Cfg>
__global__ void __launch_bounds__(kTableStatsThreads, 4) pillar_table_stats_kernel(const __grid_constant__ TableArgs t,
                                                                                         const __grid_constant__ PfnArgs a) {
    pdl_wait();
    constexpr int KIN = Cfg::KIN, G = Cfg::G, RS = Cfg::RS, COLS = Cfg::COLS, NACC = Cfg::NACC;
    using SM = StatMap<Cfg>;
    constexpr int NA = SM::NA, NB = SM::NB, NRR = SM::NRR, NW = kTableStatsThreads / 32;
    constexpr int VS = NA + NB + 1;                  // odd stride (in doubles): the lanes' vectors start in different banks
    constexpr int LW = (COLS + 3) / 4 * 4;
    __shared__ double vec[NW][32][VS];
    __shared__ double red[NW][3 * 64 + NRR];
    __shared__ double sm_ep[G + G * G];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = t.counters[RDP_CNT_P];
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;   // C fragments of the two 8 x 8 tiles (rows 0..7 | 8..15 of the A operand)
    double rr[NRR];
#pragma unroll
    for (int e = 0; e < NRR; ++e) rr[e] = 0.0;

    // one row into the thread's sums: rsum += r, rr += r r^T (upper triangle)
    auto add_row = [&](const float *src, float cenx, float ceny, float cenz, double *rsum) {
        float r[LW], g[KIN];
        load_row<LW, RS>(src, r);
        row_inputs<COLS, Cfg::DIST>(r, cenx, ceny, cenz, g);
        double gd[KIN];
#pragma unroll
        for (int k = 0; k < KIN; ++k) { gd[k] = (double)g[k]; rsum[k] += gd[k]; }
        int e = 0;
#pragma unroll
        for (int k = 0; k < KIN; ++k)
#pragma unroll
            for (int l = k; l < KIN; ++l, ++e) rr[e] = fma(gd[k], gd[l], rr[e]);
    };

    const int stride = gridDim.x * kTableStatsThreads;
    for (int pb = blockIdx.x * kTableStatsThreads + (tid & ~31); pb < P; pb += stride) {
        const int p = pb + lane;
        const bool valid = p < P;
        int s = 0, e = 0, key = 0;
        if (valid) {
            s = t.starts[p]; e = t.starts[p + 1];
            key = __float_as_int(__ldg(t.grows + ((size_t)s + 1) * RS));
        }
        float cenx = 0.f, ceny = 0.f, cenz = 0.f;
        int4 c = make_int4(0, 0, 0, 0);
        if (valid) decode_key(t.g, key, &cenx, &ceny, &cenz, &c);
        const bool big = valid && (e - s) > kBigRows;
        double rsum[KIN];
#pragma unroll
        for (int k = 0; k < KIN; ++k) rsum[k] = 0.0;
        if (valid && !big) {
            const float *r = t.grows + ((size_t)s + 1) * RS;
            for (int i = s; i < e; ++i, r += RS) add_row(r, cenx, ceny, cenz, rsum);
        }
        unsigned bigmask = __ballot_sync(0xffffffffu, big);
        while (bigmask) {   // one long pillar: the whole warp walks its rows (any lane may own any row's products)
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const int sb = __shfl_sync(0xffffffffu, s, src), eb = __shfl_sync(0xffffffffu, e, src);
            const float bx = __shfl_sync(0xffffffffu, cenx, src), by = __shfl_sync(0xffffffffu, ceny, src),
                        bz = __shfl_sync(0xffffffffu, cenz, src);
            double part[KIN];
#pragma unroll
            for (int k = 0; k < KIN; ++k) part[k] = 0.0;
            for (int i = sb + lane; i < eb; i += 32) add_row(t.grows + ((size_t)i + 1) * RS, bx, by, bz, part);
#pragma unroll
            for (int k = 0; k < KIN; ++k) {
                double v = part[k];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == src) rsum[k] = v;
            }
        }
        // stage the pillar's operands: A = [row sums | n q | 0], B = [q | 1 | 0 0]   (an invalid lane stages zeros)
        double *ev = vec[warp][lane];
        if (valid) {
            float entry[8];
            pillar_entry(cenx, ceny, cenz, rsum[0], rsum[1], rsum[2], s, e - s, entry);
            if (t.coords) store_coords(t.coords, t.coord_cols, (size_t)p, c);
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(t.aux + (size_t)p * 8), "f"(entry[0]), "f"(entry[1]),
                         "f"(entry[2]), "f"(entry[3]), "f"(entry[4]), "f"(entry[5]), "f"(entry[6]), "f"(entry[7])
                         : "memory");
            const double n = (double)(e - s);
#pragma unroll
            for (int k = 0; k < KIN; ++k) ev[k] = rsum[k];
#pragma unroll
            for (int l = 0; l < 5; ++l) {
                const double q = (double)entry[l];
                ev[KIN + l] = n * q;          // exact: n < 2^31, q has 24 significant bits
                ev[NA + l] = q;
            }
#pragma unroll
            for (int k = KIN + 5; k < NA; ++k) ev[k] = 0.0;
            ev[NA + 5] = 1.0; ev[NA + 6] = 0.0; ev[NA + 7] = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < NA + NB; ++k) ev[k] = 0.0;
        }
        __syncwarp();
        // rank-32 update on the fp64 tensor pipe: k-step j covers pillars 4 j .. 4 j + 3 of the batch.  Fragment layout of
        // mma.m8n8k4.f64: A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4], C[row = lane / 4][col = 2 (lane % 4) + {0, 1}]
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double *w = vec[warp][4 * j + (lane & 3)];
            const double bq = w[NA + (lane >> 2)];
            dmma_m8n8k4(c00, c01, w[lane >> 2], bq);
            dmma_m8n8k4(c10, c11, w[8 + (lane >> 2)], bq);
        }
        __syncwarp();
    }

    // ---- per-CTA sums -> fp64 totals (atomics); the CTA that finishes last runs the BatchNorm epilogue
    {
        const int row = lane >> 2, col = 2 * (lane & 3);
        red[warp][row * 8 + col] = c00; red[warp][row * 8 + col + 1] = c01;
        red[warp][64 + row * 8 + col] = c10; red[warp][64 + row * 8 + col + 1] = c11;
    }
#pragma unroll
    for (int e = 0; e < NRR; ++e) {
        double v = rr[e];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) red[warp][128 + e] = v;
    }
    __syncthreads();
    for (int e = tid; e < 128 + NRR; e += kTableStatsThreads) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) sacc += red[w][e];
        if (sacc == 0.0) continue;
        const int dst = e < 128 ? SM::dst(e >> 3, e & 7) : SM::rr_dst(e - 128);
        if (dst >= 0) atomicAdd(a.acc_stats + dst, sacc);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int32_t *done = a.counters + kCntDoneStats;
        const int ticket = atomicAdd(done, 1);
        s_last = (ticket == (int)gridDim.x - 1);
        if (s_last) *done = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (a.defer_finalize) {   // SyncBatchNorm: the totals (and the point count beside them) are all-reduced over the ranks first
        if (tid == 0) a.acc_stats[NACC] = (double)a.counters[RDP_CNT_N];
        return;
    }
    bn_epilogue<Cfg>(a, a.acc_stats, sm_ep);
    __syncthreads();
    for (int e = tid; e <= NACC; e += kTableStatsThreads) a.acc_stats[e] = 0.0;   // ready for the next launch on this workspace
}

// Stand-alone epilogue: SyncBatchNorm (after the all-reduce of the totals).  One CTA.
template <class Cfg>
__global__ void __launch_bounds__(256) bn_finalize_kernel(const __grid_constant__ PfnArgs a) {
    constexpr int NACC = Cfg::NACC, G = Cfg::G;
    __shared__ double sm[G + G * G];
    bn_epilogue<Cfg>(a, a.acc_stats, sm);
    __syncthreads();
    for (int e = threadIdx.x; e <= NACC; e += blockDim.x) a.acc_stats[e] = 0.0;
}

// ------------------------------------------------------------------------------------------- backward
// Closed-form BatchNorm / linear backward in the reduced basis.  A_c = sum_p gy [g_win | 1] (the last
// entry is dbeta_c).  With x = wg . g + const:
//   dgamma_c = (wg_c . A_c + (const_c - mu_c) dbeta_c) / sigma_c
//   D_c      = (gamma_c/sigma_c) [ A_c - dbeta_c/n [S1 | n] - dgamma_c/(n sigma_c) [S2 wg_c + (const_c - mu_c) S1 | 0] ]   (train)
//   D_c      = (gamma_c/sigma_c) A_c                                                                        (eval BN / no norm: 1)
//   dW_c     = T D_c   (back to the layout's feature order)
// `glob` (SyncBatchNorm) holds the all-reduced A when the statistics span several ranks: dbeta / dgamma in the
// correction terms are the global ones while A, S1, S2 stay this rank's (DDP then averages dW over the ranks).
template <class Cfg>
__device__ __forceinline__ void bwd_epilogue(const PfnArgs &a, const double *totals, const double *glob, double *sm) {
    constexpr int COUT = Cfg::COUT, G = Cfg::G, PER = Cfg::BWD_PER;
    const int tid = threadIdx.x;
    double *D = sm;   // COUT * PER
    const double *st = a.bn_state;
    const bool train = a.use_norm && a.train_bn;
    const double n = train ? st[4 * COUT] : 0.0;
    const double *S1 = train ? st + 4 * COUT + 1 : nullptr, *S2 = train ? S1 + G : nullptr;
    for (int c = tid; c < COUT; c += blockDim.x) {
        float wg[G], cst;
        fold_channel<G>(a, c, wg, &cst);
        double mu = 0.0, is = 1.0, gam = 1.0;
        if (a.use_norm) {
            mu = train ? st[c] : (double)a.rmean[c];
            const double var = train ? st[COUT + c] : (double)a.rvar[c];
            is = 1.0 / sqrt(var + a.eps);
            gam = (double)a.gamma[c];
        }
        double A[PER], Ag[PER];
#pragma unroll
        for (int m = 0; m < PER; ++m) { A[m] = __ldcg(totals + c * PER + m); Ag[m] = glob ? __ldcg(glob + c * PER + m) : A[m]; }
        auto dgamma_of = [&](const double *v) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < G; ++k) s = fma((double)wg[k], v[k], s);
            return (s + ((double)cst - mu) * v[G]) * is;
        };
        const double db = A[G], dg = dgamma_of(A);
        const double dbg = Ag[G], dgg = glob ? dgamma_of(Ag) : dg;
        a.d_beta[c] = (float)db;
        if (a.use_norm && a.d_gamma) a.d_gamma[c] = (float)dg;
        const double ak = gam * is;
#pragma unroll
        for (int m = 0; m < G; ++m) {
            double v = A[m];
            if (train && n > 0.0) {
                double s2w = 0.0;
#pragma unroll
                for (int l = 0; l < G; ++l) s2w += S2[m * G + l] * (double)wg[l];
                v = v - dbg / n * S1[m] - dgg / n * is * (s2w + ((double)cst - mu) * S1[m]);
            }
            D[c * PER + m] = ak * v;
        }
        // the constant direction: sum_i g_x = 0 under batch statistics when they cover exactly the rows summed here
        double vlast = A[G];
        if (train && n > 0.0) {
            const double nloc = (double)a.counters[RDP_CNT_N];
            double w1 = 0.0;
#pragma unroll
            for (int k = 0; k < G; ++k) w1 += (double)wg[k] * S1[k];
            vlast = A[G] - dbg / n * nloc - dgg / n * is * (w1 + ((double)cst - mu) * nloc);
        }
        D[c * PER + G] = ak * vlast;
    }
    __syncthreads();
    for (int e = tid; e < COUT * a.c_in; e += blockDim.x) {
        const int c = e / a.c_in, j = e % a.c_in;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < PER; ++m) s += (double)a.T[j][m] * D[c * PER + m];
        a.d_weight[e] = (float)s;
    }
}

template <class Cfg>
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const __grid_constant__ PfnArgs a, const double *glob) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    bwd_epilogue<Cfg>(a, a.acc_bwd, glob, reinterpret_cast<double *>(smem_raw));
    __syncthreads();
    for (int e = threadIdx.x; e < Cfg::COUT * Cfg::BWD_PER; e += blockDim.x) a.acc_bwd[e] = 0.0;
}

template <class Cfg>
__global__ void __launch_bounds__(kPfnThreads, Cfg::CPL == 1 ? 5 : (Cfg::CPL == 2 ? 3 : 2)) pfn_bwd_kernel(const __grid_constant__ PfnArgs a) {
    pdl_wait();
    extern __shared__ __align__(32) unsigned char smem_raw[];
    constexpr int PCH = 24;   // pillars per prefetch chunk of (grad, argpos) rows (two chunks in flight)
    using Smem = TileSmem<Cfg, PCH>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, COLS = Cfg::COLS, RS = Cfg::RS, CPL = Cfg::CPL, KIN = Cfg::KIN, G = Cfg::G, PER = Cfg::BWD_PER;
    constexpr int WIN = kPfnWin, NW = kPfnThreads / 32;
    constexpr bool DIST = Cfg::DIST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = a.counters[RDP_CNT_N];
    const int ntiles = (int)((N + WIN - 1) / WIN);
    const int Gd = (int)gridDim.x;
    const int nk = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + Gd - 1) / Gd : 0;
    auto tile_of = [&](int k) { return (int)blockIdx.x + k * Gd; };
    static_assert(sizeof(double) * COUT * PER <= sizeof(PfnStage<Cfg>) * 2, "end-of-kernel scratch aliases the stages");

    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        mbar_init(&S.pre[0], 1);
        mbar_init(&S.pre[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    double dA[CPL][PER];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
        for (int m = 0; m < PER; ++m) dA[cc][m] = 0.0;

    auto issue = [&](int t, int s, int pf, int pn) {
        PfnStage<Cfg> &T = S.st[s];
        S.tf[s][0] = pf; S.tf[s][1] = pn;
        mbar_expect_tx(&S.full[s], Cfg::ROW_BYTES + Cfg::AUX_BYTES);
        tma_bulk_g2s(T.rows, a.grows + (size_t)t * WIN * RS, Cfg::ROW_BYTES, &S.full[s]);
        tma_bulk_g2s(T.aux, a.aux + (size_t)pf * 8, Cfg::AUX_BYTES, &S.full[s]);
    };
    // the upstream gradient and argmax rows of pillars [p0, p0 + n) -> smem, asynchronously
    auto prefetch = [&](int b, int p0, int n) {
        const uint32_t bytes = (uint32_t)n * COUT * 4;
        mbar_expect_tx(&S.pre[b], 2 * bytes);
        tma_bulk_g2s(S.pre_grad[b], a.grad + (size_t)p0 * COUT, bytes, &S.pre[b]);
        tma_bulk_g2s(S.pre_arg[b], a.argpos + (size_t)p0 * COUT, bytes, &S.pre[b]);
    };
    // gy * [row inputs | pillar constants | 1] of the winning row into the fp32 tile sums
    auto route_rin = [&](const float *rin, const float4 &a0, float ndz, float gy, float *tA) {
#pragma unroll
        for (int k = 0; k < KIN; ++k) tA[k] = fmaf(gy, rin[k], tA[k]);
        tA[KIN] = fmaf(gy, a0.x, tA[KIN]);
        tA[KIN + 1] = fmaf(gy, a0.y, tA[KIN + 1]);
        tA[KIN + 2] = fmaf(gy, a0.z, tA[KIN + 2]);
        tA[KIN + 3] = fmaf(gy, a0.w, tA[KIN + 3]);
        tA[KIN + 4] = fmaf(gy, ndz, tA[KIN + 4]);
        tA[G] += gy;
    };
    // winner row given by its record in shared memory (tile_c1)
    auto route = [&](const float *rec, const float4 &a0, float ndz, float gy, float *tA) {
        float rin[Cfg::FW];
#pragma unroll
        for (int k4 = 0; k4 < (KIN + 3) / 4; ++k4) {
            const float4 q = *reinterpret_cast<const float4 *>(rec + 4 * k4);
            rin[4 * k4] = q.x; rin[4 * k4 + 1] = q.y; rin[4 * k4 + 2] = q.z; rin[4 * k4 + 3] = q.w;
        }
        route_rin(rin, a0, ndz, gy, tA);
    };
    // big-pillar path: winner row straight from global memory
    auto route_global = [&](const float *src, const float4 &a0, float ndz, float gy, float *tA) {
        float row[RS], rin[KIN];
#pragma unroll
        for (int c4 = 0; c4 < (COLS + 3) / 4 * 4; c4 += 4) {
            const float4 q = *reinterpret_cast<const float4 *>(src + c4);
            row[c4] = q.x; row[c4 + 1] = q.y; row[c4 + 2] = q.z; row[c4 + 3] = q.w;
        }
        row_inputs<COLS, DIST>(row, a0.x, a0.y, a.off_z, rin);
        route_rin(rin, a0, ndz, gy, tA);
    };

    int nfa = 0, nfb = 0;
    if (nk > 0 && tid == 0) {
        issue(tile_of(0), 0, a.tile_first[tile_of(0)], a.tile_first[tile_of(0) + 1]);
        if (nk > 1) { nfa = a.tile_first[tile_of(1)]; nfb = a.tile_first[tile_of(1) + 1]; }
    }
    uint32_t par0 = 0, par1 = 0;
    int cj = 0;             // chunks consumed so far (buffer cj & 1, phase (cj >> 1) & 1)
    bool pending = false;   // the first chunk of the tile being entered was requested during the previous tile

    for (int k = 0; k < nk; ++k) {
        const int t = tile_of(k);
        const int s = k & 1;
        if (tid == 0 && k + 1 < nk) {
            issue(tile_of(k + 1), s ^ 1, nfa, nfb);
            if (k + 2 < nk) { nfa = a.tile_first[tile_of(k + 2)]; nfb = a.tile_first[tile_of(k + 2) + 1]; }
        }
        // first pillars of the NEXT tile (for the cross-tile prefetch below): every thread reads them from global memory here,
        // a tile ahead of their use -- S.tf of the other stage is only ordered with respect to that stage's mbarrier
        int ps1 = 0, pe1 = 0;
        if (k + 1 < nk) { ps1 = __ldg(a.tile_first + tile_of(k + 1)); pe1 = __ldg(a.tile_first + tile_of(k + 1) + 1); }
        if (s == 0) { mbar_wait(&S.full[0], par0); par0 ^= 1; } else { mbar_wait(&S.full[1], par1); par1 ^= 1; }
        const PfnStage<Cfg> &T = S.st[s];
        const long long base = (long long)t * WIN;
        const int ps = S.tf[s][0], pe = S.tf[s][1];
        if (pe == ps) { __syncthreads(); continue; }
        const TileBounds tb = tile_bounds<Cfg>(T, ps, pe, base);
        const int nb = tb.nb;
        if (!pending && nb > 0 && tid == 0) prefetch(cj & 1, ps, min(PCH, nb));
        if (pending && nb == 0) {   // only a big pillar starts here: drain the chunk that was requested for this tile
            mbar_wait(&S.pre[cj & 1], (uint32_t)(cj >> 1) & 1u);
            ++cj;
        }
        pending = false;

        if (nb > 0) {
            tile_c1<Cfg, false>(T, tb, base, a.off_z, S.frec);
            __syncthreads();
            float tA[CPL][PER];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
                for (int m = 0; m < PER; ++m) tA[cc][m] = 0.0f;
            for (int q0 = 0; q0 < nb; q0 += PCH) {
                const int nq = min(PCH, nb - q0);
                if (q0 > 0) __syncthreads();   // every warp is done with the buffer the next request lands in
                // request the chunk after this one -- the rest of this tile, else the head of the next tile -- so that one
                // chunk is always in flight while another is folded in
                if (q0 + PCH < nb) {
                    if (tid == 0) prefetch((cj + 1) & 1, ps + q0 + PCH, min(PCH, nb - q0 - PCH));
                } else if (k + 1 < nk) {
                    if (pe1 > ps1) {
                        pending = true;
                        if (tid == 0) prefetch((cj + 1) & 1, ps1, min(PCH, pe1 - ps1));
                    }
                }
                const int b = cj & 1;
                mbar_wait(&S.pre[b], (uint32_t)(cj >> 1) & 1u);
                ++cj;
                for (int q = warp; q < nq; q += NW) {   // warp = pillar, lane = channel
                    const float *ax = T.aux + (q0 + q) * 8;
                    const float4 a0 = *reinterpret_cast<const float4 *>(ax);
                    const float ndz = ax[4];
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) {
                        const int o = q * COUT + lane + 32 * cc;
                        const int ap = S.pre_arg[b][o];   // -1: the forward marked the pillar as ReLU-clamped (:38)
                        if (ap >= 0) route(S.frec + (ap - (int)base) * Cfg::FW, a0, ndz, S.pre_grad[b][o], tA[cc]);
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
                for (int m = 0; m < PER; ++m) dA[cc][m] += (double)tA[cc][m];
        }

        if (tb.big && warp == 0) {
            // the big pillar's winners may lie outside the staged rows: straight from global memory
            const int pb = pe - 1;
            const float *ax = T.aux + (pe - ps - 1) * 8;
            const float4 a0 = *reinterpret_cast<const float4 *>(ax);
            const float ndz = ax[4];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const size_t o = (size_t)pb * COUT + lane + 32 * cc;
                const int ap = a.argpos[o];
                float tB[PER];
#pragma unroll
                for (int m = 0; m < PER; ++m) tB[m] = 0.0f;
                if (ap >= 0) route_global(a.grows + ((size_t)ap + 1) * RS, a0, ndz, a.grad[o], tB);
#pragma unroll
                for (int m = 0; m < PER; ++m) dA[cc][m] += (double)tB[m];
            }
        }
        __syncthreads();
    }

    // ---- per-CTA sums -> fp64 totals (atomics); the CTA that finishes last runs the closed-form epilogue
    double *dscr = reinterpret_cast<double *>(&S.st[0]);   // [channel][PER], aliases the (now idle) stages
    __syncthreads();
    for (int w2 = 0; w2 < NW; ++w2) {   // the warps add their sums in turn
        if (warp == w2) {
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc)
#pragma unroll
                for (int m = 0; m < PER; ++m) {
                    double *d = dscr + (size_t)(lane + 32 * cc) * PER + m;
                    *d = (w2 == 0 ? 0.0 : *d) + dA[cc][m];
                }
        }
        __syncthreads();
    }
    for (int e = tid; e < COUT * PER; e += kPfnThreads) {
        const double sacc = dscr[e];
        if (sacc != 0.0) atomicAdd(a.acc_bwd + e, sacc);
    }
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int32_t *done = a.counters + kCntDoneBwd;
        const int ticket = atomicAdd(done, 1);
        s_last = (ticket == (int)gridDim.x - 1);
        if (s_last) *done = 0;
    }
    __syncthreads();
    if (!s_last || a.defer_finalize) return;
    __threadfence();
    bwd_epilogue<Cfg>(a, a.acc_bwd, nullptr, dscr);
    __syncthreads();
    for (int e = tid; e < COUT * PER; e += kPfnThreads) a.acc_bwd[e] = 0.0;
}

}  // namespace rdp
