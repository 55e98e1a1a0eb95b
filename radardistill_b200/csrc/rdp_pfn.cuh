// rdp_pfn.cuh -- device code of the fused pillar feature network (shared by the per-config .cu files).
//
// Replaces /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:214-240 and
// PFNLayerV2.forward :35-46: scatter_mean, f_center / f_cluster / f_relative, concat,
// Linear(no bias)+BatchNorm1d+ReLU, scatter_max (+argmax), and the autograd of that chain.
//
// Input is the pillar-grouped row array written by group_rows_kernel and the pillar table written by
// pillar_table_kernel (rdp_index.cu): rows of one pillar are contiguous, pillars are in key order, and the table
// holds every pillar's mean, centre, first row and row count.  One persistent CTA (128 threads) takes tiles
// b, b + grid, b + 2 grid, ...; tile t owns the pillars that START in grouped rows [128 t, 128 t + 128) and stages
// rows [128 t - 1, 128 t + 192) plus the table slice of its pillars -- fixed-size, 16-byte aligned windows -- one
// tile ahead with two 1-D TMA bulk copies onto an mbarrier, double buffered.  A pillar that runs past the staged
// window (> 64 rows of overhang) takes the "big pillar" path straight from global memory.
//
//   C1  decorated features (same op order / roundings as the reference) -> smem; slot / last-row flags (thread = row)
//   ST  APPLY / APPLY_ARG: sub-groups of 16 lanes stream pillar-aligned eighths of the tile's rows, each lane
//       owning two output channels with their weight rows in registers: x = W f as a k-ascending fmaf chain (feature
//       row broadcast from smem), y = fma(x, scale, shift), running max (+ lowest-index argmax as one packed 64-bit
//       key), one coalesced 128 B store when a pillar's last row is in; argpos < 0 marks a ReLU-clamped maximum
//       STATS: the fp64 Gram matrix of [features | 1] only (4x4 register blocks) -- the batch statistics of x = W f and
//              the moments the backward needs both follow from it (bn_finalize_kernel), no linear layer in this pass
//       BWD  : (grad, argpos) rows of 24 pillars at a time through a double buffer requested one chunk ahead; per
//              (pillar, channel) route the gradient to the winning row, accumulate dbeta and A (G = w . A)
// Arithmetic is the canonical form of oracle/pillar_oracle.c (ORC_MEAN_F64): outputs are bit-identical to it.
#pragma once

#include "rdp_common.cuh"

#ifndef RDP_LPR
#define RDP_LPR 16   // lanes that share one feature row in the forward stream (32 / 16 / 8)
#endif

namespace rdp {

constexpr int kMaxSuper = 24;

struct PfnArgs {
    const float *grows;
    const float *aux;               // per-pillar table [mean xyz | centre xy | first grouped row | rows | 0]
    const int32_t *tile_first;      // first pillar starting at or after grouped row 128 t
    const int32_t *ends, *counters, *orig2kept;
    const float *weight, *bias, *gamma, *beta, *rmean, *rvar;
    const double *bn_state;  // train apply: folded scale / shift live here
    float *features;
    int32_t *argpos;         // winning GROUPED position per (pillar, channel), bit-complemented (negative) when the
                             // pillar's maximum was clamped by the ReLU (zero gradient); see argpos_to_kept_kernel
    float *pillar_mean;
    double *partials;
    const float *grad;       // backward input: upstream gradient (P, Cout)
    long long n0;
    double eps;
    float lo[3], vsz[3], off[3];
    int c_in;
    int use_norm;
    int fold_from_state;     // 1: scale/shift from bn_state (train), 0: fold running stats in-kernel (eval)
    int8_t kmap[kMaxSuper];  // super-feature -> layout column of W, or -1 (zero weight)
};

enum { PFN_MODE_APPLY = 0, PFN_MODE_STATS = 1, PFN_MODE_BWD = 2, PFN_MODE_APPLY_ARG = 3 };  // APPLY_ARG also records the argmax

// Compile-time shape of one encoder family.  "Super features" are every decoration the layout could
// use, in the layout's concat order; options switched off in model_cfg get a zero weight column
// (fmaf(0, f, acc) == acc), so one instantiation serves all flag combinations bit-exactly.
template <int COLS_, int LAYOUT_, bool DIST_, int COUT_>
struct PfnCfg {
    static constexpr int COLS = COLS_, LAYOUT = LAYOUT_, COUT = COUT_, C = COLS_ - 1;
    static constexpr bool DIST = DIST_;
    static constexpr int CS = (LAYOUT_ == RDP_LAYOUT_SIMPLE2D) ? (3 + C + 3 + (DIST_ ? 1 : 0) + 3) : (C + 6 + (DIST_ ? 1 : 0));
    static constexpr int T4 = (CS + 1 + 3) / 4;   // 4-wide column blocks of [features | 1]
    static constexpr int FW = 4 * T4;
    static constexpr int FSTRIDE = (FW % 16 == 0) ? FW + 4 : FW;  // smem row stride of features (bank-conflict free)
    static constexpr int QUADS = COUT / 4;                         // channel quads
    static constexpr int GROUPS = kPfnThreads / QUADS;             // row groups in the register tiling
    static constexpr int RPT = (kPfnCap + GROUPS - 1) / GROUPS;    // rows per thread in C2
    static constexpr int ZSTRIDE = COUT + 4;
    static constexpr int NBLK = T4 * (T4 + 1) / 2;                 // upper-triangle 4x4 blocks of the Gram matrix
    static constexpr int RG = kPfnThreads / NBLK;                  // row groups in the Gram phase
    static constexpr int STATS_DOUBLES = 16 * NBLK;   // Gram blocks of [features | 1] (batch moments of x follow from them)
    static constexpr int BWD_PER = CS + 2;
    static constexpr int BWD_DOUBLES = COUT * BWD_PER;
    static constexpr int RS = (COLS + 2 + 3) / 4 * 4;  // floats per grouped row: the row, padding, original row id, pillar id
    static constexpr uint32_t ROW_BYTES = (kPfnCap + 1) * RS * 4;  // the window plus the row in front of it
    static constexpr uint32_t AUX_BYTES = kPfnWin * 8 * 4;
    static_assert(CS <= kMaxSuper, "too many features");
    static_assert(ROW_BYTES % 16 == 0, "TMA sizes");
    static_assert(NBLK <= kPfnThreads && COUT % 32 == 0, "tiling");
};

// One staged tile: grouped rows [128 t - 1, 128 t + 192); row j of the window is rows[(j + 1) * RS ...].
// Slot RS-2 of a row holds its original row index, slot RS-1 its pillar id (int bit patterns).
template <class Cfg>
struct PfnStage {
    alignas(32) float aux[kPfnWin * 8];   // table entries of the (at most 128) pillars that start in the window
    alignas(16) float rows[(kPfnCap + 1) * Cfg::RS];
    __device__ __forceinline__ int gid(int j) const { return __float_as_int(rows[(j + 1) * Cfg::RS + Cfg::RS - 1]); }
    __device__ __forceinline__ int ord(int j) const { return __float_as_int(rows[(j + 1) * Cfg::RS + Cfg::RS - 2]); }
    __device__ __forceinline__ const float *row(int j) const { return rows + (j + 1) * Cfg::RS; }
};

template <class Cfg, int MODE>
struct PfnSmem {
    static constexpr size_t Z_BYTES = 16;
    static constexpr size_t S_BYTES = (MODE == PFN_MODE_STATS) ? sizeof(double) * kPfnThreads * 16 : 0;
    static constexpr size_t B_BYTES = 0;  // BWD: the end-of-kernel scratch aliases f + the prefetch buffers (see bwd_scratch())
    static constexpr size_t SCR = Z_BYTES > S_BYTES ? (Z_BYTES > B_BYTES ? Z_BYTES : B_BYTES) : (S_BYTES > B_BYTES ? S_BYTES : B_BYTES);
    static constexpr int PCH = (MODE == PFN_MODE_BWD) ? 24 : 1;  // pillars per backward prefetch chunk (two chunks in flight)
    PfnStage<Cfg> st[2];
    alignas(8) uint64_t full[2];
    alignas(8) uint64_t pre[2];                    // BWD: arrival of a chunk's (grad, argpos) rows, double buffered
    alignas(16) float pre_grad[2][(MODE == PFN_MODE_BWD) ? PCH * Cfg::COUT : 4];
    alignas(16) int pre_arg[2][(MODE == PFN_MODE_BWD) ? PCH * Cfg::COUT : 4];
    alignas(16) float f[kPfnCap * Cfg::FSTRIDE];   // decorated features of the tile's rows
    alignas(16) unsigned char scr[SCR];            // STATS / BWD: fp64 reduction scratch at kernel end
    int start[kPfnCap + 1];
    int lp[kPfnCap];                               // (pillar slot << 1) | last-row-of-pillar flag
    int kept[kPfnCap];
    float bigaux[8];                               // big-pillar path: table entry of the pillar being streamed
    int tf[2][2];                                  // per stage: first pillar of the tile, first pillar of the next
    float scale[Cfg::COUT], shift[Cfg::COUT];
    float carry_v[Cfg::COUT];
    int carry_k[Cfg::COUT], carry_p[Cfg::COUT];
};

// ------------------------------------------------------------------------------------------- features
template <class Cfg>
__device__ __forceinline__ void decorate(const float *r, float cenx, float ceny, const float *mean, const PfnArgs &a, float *f) {
    const float x = r[1], y = r[2], z = r[3];
    float cen[3], clu[3];
    cen[0] = __fsub_rn(x, cenx);               // x - (cx*vx + x_off); the bracket is per pillar (P2)
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

// BatchNorm folded to y = fma(x, scale, shift) in fp64 with one rounding (oracle: orc_bn_fold).
__device__ __forceinline__ void fold_bn(double gamma, double beta, double mean, double var, double eps, float *scale, float *shift) {
    const double inv_std = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(var, eps)));
    const double s = __dmul_rn(gamma, inv_std);
    *scale = (float)s;
    *shift = (float)__dsub_rn(beta, __dmul_rn(mean, s));
}

// (lo, hi) -> one 64-bit register pair, for lexicographic compares that cost two ISETPs
__device__ __forceinline__ unsigned long long pack64(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }

// ------------------------------------------------------------------------------------------- the tile kernel
template <class Cfg, int MODE>
__global__ void __launch_bounds__(kPfnThreads, (MODE == PFN_MODE_APPLY) ? 5 : 4) pfn_tile_kernel(const __grid_constant__ PfnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = PfnSmem<Cfg, MODE>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, COLS = Cfg::COLS, RS = Cfg::RS, CPL = COUT / 32;
    constexpr int WIN = kPfnWin, CAP = kPfnCap, NT = kPfnThreads, NW = kPfnThreads / 32, INF = 0x7fffffff;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = a.counters[RDP_CNT_N];
    const bool none_dropped = (N == a.n0);
    const int ntiles = (int)((N + WIN - 1) / WIN);
    // Tile order: CTA b takes tiles b, b + grid, b + 2 grid, ... -- at any moment the grid works on one contiguous window
    // of the row array and of the outputs (DRAM pages are used completely) instead of one private stream per CTA.
    const int G = (int)gridDim.x;
    const int nk = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + G - 1) / G : 0;   // tiles of this CTA
    auto tile_of = [&](int k) { return (int)blockIdx.x + k * G; };
    constexpr bool want_arg = (MODE == PFN_MODE_APPLY_ARG);   // compile-time: the eval kernel carries no argmax state
    constexpr bool is_apply = (MODE == PFN_MODE_APPLY) || (MODE == PFN_MODE_APPLY_ARG);
    // fp64 reduction scratch used once at the end of the kernel; in BWD it aliases the (then idle) prefetch + feature buffers
    double *dscr = (MODE == PFN_MODE_BWD) ? reinterpret_cast<double *>(&S.pre_grad[0][0]) : reinterpret_cast<double *>(S.scr);
    static_assert(MODE != PFN_MODE_BWD ||
                  sizeof(double) * (kPfnThreads / 32) * Cfg::BWD_DOUBLES <= 4 * sizeof(float) * Smem::PCH * Cfg::COUT + sizeof(S.f),
                  "backward scratch must fit in pre_grad | pre_arg | f");

    // ---- per-CTA constants
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        mbar_init(&S.pre[0], 1);
        mbar_init(&S.pre[1], 1);
        fence_mbar_init();
    }
    // Forward stream: LPR lanes share one feature row; lane sl of the sub-group owns channels sl + LPR q, whose weight
    // rows live in registers for the whole kernel.  With LPR = 16 a 16-byte feature load feeds two rows' worth of lanes
    // (two addresses per warp instruction), which halves the shared-memory wavefronts per row -- the pipe that bounds the
    // lane-per-channel form (ncu: l1tex lsu wavefronts 80 %).  Backward: lane = channel (+32 cc).
    constexpr int LPR = RDP_LPR, SUB = 32 / LPR, NS = NW * SUB;
    constexpr int WN = (MODE == PFN_MODE_BWD) ? CPL : COUT / LPR;
    const int sl = lane % LPR, sid = warp * SUB + lane / LPR;
    float W[WN][CS], sc[WN], sh[WN];
#pragma unroll
    for (int cc = 0; cc < WN; ++cc) {
        const int ch = (MODE == PFN_MODE_BWD) ? (lane + 32 * cc) : (sl + LPR * cc);
#pragma unroll
        for (int s = 0; s < CS; ++s) {
            const int k = a.kmap[s];
            W[cc][s] = (k >= 0 && is_apply) ? __ldg(a.weight + ch * a.c_in + k) : 0.0f;
        }
        sc[cc] = 1.0f;
        sh[cc] = 0.0f;
        if (is_apply) {
            if (a.bias) sh[cc] = a.bias[ch];
            if (a.use_norm) {
                if (a.fold_from_state) { sc[cc] = (float)a.bn_state[2 * COUT + ch]; sh[cc] = (float)a.bn_state[3 * COUT + ch]; }
                else fold_bn((double)a.gamma[ch], (double)a.beta[ch], (double)a.rmean[ch], (double)a.rvar[ch], a.eps, &sc[cc], &sh[cc]);
            }
        }
    }
    __syncthreads();

    // ---- accumulators that live for the whole CTA
    double st_m[16];
    int gba = 0, gbb = 0;
    const int grg = tid % Cfg::RG, gblk = tid / Cfg::RG;
    if (MODE == PFN_MODE_STATS) {
#pragma unroll
        for (int e = 0; e < 16; ++e) st_m[e] = 0.0;
        int rem = gblk;
        while (gba < Cfg::T4 && rem >= Cfg::T4 - gba) { rem -= Cfg::T4 - gba; ++gba; }
        gbb = gba + rem;  // block (gba, gbb), gba <= gbb, valid when gblk < NBLK
    }
    double dB[CPL], dA[CPL][CS];
    if (MODE == PFN_MODE_BWD) {
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            dB[cc] = 0.0;
#pragma unroll
            for (int k = 0; k < CS; ++k) dA[cc][k] = 0.0;
        }
    }

    // thread 0 only.  (pf, pn) = tile_first[t], tile_first[t + 1]; published to the CTA through S.tf before the arrive.
    auto issue = [&](int t, int s, int pf, int pn) {
        PfnStage<Cfg> &T = S.st[s];
        S.tf[s][0] = pf; S.tf[s][1] = pn;
        mbar_expect_tx(&S.full[s], Cfg::ROW_BYTES + Cfg::AUX_BYTES);
        tma_bulk_g2s(T.rows, a.grows + (size_t)t * WIN * RS, Cfg::ROW_BYTES, &S.full[s]);  // grows row 0 = the row before position 0
        tma_bulk_g2s(T.aux, a.aux + (size_t)pf * 8, Cfg::AUX_BYTES, &S.full[s]);
    };

    // C1 (thread = row): decorated features of rows [rowbase, rowbase + np) of `rows` -> S.f[0..np), the row's
    // (pillar slot << 1 | last-row flag) -> S.lp, pillar start table -> S.start.  `ps >= 0`: rows of a staged tile, the
    // slot is gid - ps and `aux` the tile's table slice; ps < 0: chunk of one big pillar (slot 0, never last, S.bigaux).
    auto c1 = [&](const float *rows, int rowbase, int np, int ps, const float *aux) {
        for (int jj = tid; jj < np; jj += NT) {
            float r[COLS], f[Cfg::FW];
            const float *src = rows + (rowbase + jj) * RS;
#pragma unroll
            for (int c = 0; c < COLS; ++c) r[c] = src[c];
            int slot = 0;
            const float *ax = S.bigaux;
            if (ps >= 0) {
                const int gid = __float_as_int(src[RS - 1]);
                slot = gid - ps;
                ax = aux + slot * 8;
                const int last = (jj == np - 1) || (__float_as_int(src[RS + RS - 1]) != gid);
                S.lp[jj] = (slot << 1) | last;
                if (__float_as_int(src[-1]) != gid) S.start[slot] = jj;
            } else {
                S.lp[jj] = 0;
            }
            if (want_arg) { const int row = __float_as_int(src[RS - 2]); S.kept[jj] = none_dropped ? row : a.orig2kept[row]; }
            decorate<Cfg>(r, ax[3], ax[4], ax, a, f);
            f[CS] = 1.0f;  // ones column: the Gram matrix then carries sum f (S1) as well
#pragma unroll
            for (int k = CS + 1; k < Cfg::FW; ++k) f[k] = 0.0f;
            float4 *dst = reinterpret_cast<float4 *>(&S.f[jj * Cfg::FSTRIDE]);
#pragma unroll
            for (int k4 = 0; k4 < Cfg::T4; ++k4) dst[k4] = make_float4(f[k4 * 4], f[k4 * 4 + 1], f[k4 * 4 + 2], f[k4 * 4 + 3]);
        }
    };

    // x[cc] = W[cc] . f(row j) as a k-ascending fmaf chain
    auto dot_row = [&](int j, float *x) {
        float f[Cfg::FW];
        const float4 *src = reinterpret_cast<const float4 *>(&S.f[j * Cfg::FSTRIDE]);
#pragma unroll
        for (int k4 = 0; k4 < (CS + 3) / 4; ++k4) {
            const float4 v = src[k4];
            f[k4 * 4] = v.x; f[k4 * 4 + 1] = v.y; f[k4 * 4 + 2] = v.z; f[k4 * 4 + 3] = v.w;
        }
#pragma unroll
        for (int cc = 0; cc < WN; ++cc) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < CS; ++k) acc = fmaf(W[cc][k], f[k], acc);
            x[cc] = acc;
        }
    };

    // running max state of the pillar a sub-group is streaming (lane = its channels)
    float m[WN];
    int mk[WN], mp[WN];
    auto reset_max = [&]() {
#pragma unroll
        for (int cc = 0; cc < WN; ++cc) { m[cc] = 0.0f; mk[cc] = INF; mp[cc] = 0; }   // (0, INF) loses to every real row
    };
    float *fout = nullptr;
    int32_t *aout = nullptr;
    // one row: BN (+ReLU) -> running max / lowest-index argmax (or the fp64 statistics); stores the pillar when `meta`
    // carries the last-row flag (pillars close in order: the output row pointer just advances)
    auto fold_row = [&](const float *x, int j, int gb) {
        const int meta = S.lp[j];
        const int kj = want_arg ? S.kept[j] : 0;
#pragma unroll
        for (int cc = 0; cc < WN; ++cc) {
            const float y = fmaf(x[cc], sc[cc], sh[cc]);
            if (!want_arg) {
                m[cc] = fmaxf(m[cc], y);  // ReLU folds into the max with 0
            } else {
                // larger z wins, ties go to the lower kept index: z >= 0, so (float bits of z, ~index) orders correctly as one
                // unsigned 64-bit key (two compares instead of three)
                const float z = fmaxf(y, 0.0f);
                if (pack64(~(uint32_t)kj, __float_as_uint(z)) > pack64(~(uint32_t)mk[cc], __float_as_uint(m[cc]))) {
                    m[cc] = z; mk[cc] = kj; mp[cc] = gb + j;
                }
            }
        }
        if (meta & 1) {
#pragma unroll
            for (int cc = 0; cc < WN; ++cc) {
                fout[LPR * cc] = m[cc];
                if (want_arg) aout[LPR * cc] = (m[cc] > 0.0f) ? mp[cc] : ~mp[cc];   // negative: ReLU clamped the pillar (no gradient)
            }
            fout += COUT;
            if (want_arg) aout += COUT;
            reset_max();
        }
    };

    // STREAM: every sub-group of LPR lanes walks its own pillar-aligned rows [ra, rb) of S.f, two rows in flight
    // (sub-groups of a warp with shorter ranges simply leave the loop earlier).
    auto stream = [&](int ra, int rb, int ps, int gb) {
        if (is_apply) {
            reset_max();
            const int slot0 = (ra < rb) ? (S.lp[ra] >> 1) : 0;  // first pillar this sub-group closes
            fout = a.features + (size_t)(ps + slot0) * COUT + sl;
            aout = want_arg ? a.argpos + (size_t)(ps + slot0) * COUT + sl : nullptr;
        }
        int j = ra;
        for (; j + 1 < rb; j += 2) {  // two rows in flight: two independent fmaf chains per channel
            float x0[WN], x1[WN];
            dot_row(j, x0);
            dot_row(j + 1, x1);
            fold_row(x0, j, gb);
            fold_row(x1, j + 1, gb);
        }
        if (j < rb) {
            float x0[WN];
            dot_row(j, x0);
            fold_row(x0, j, gb);
        }
    };

    // STATS: Gram matrix of [features | 1] over the np rows in S.f, 4x4 register blocks
    auto gram = [&](int np) {
        if (gblk < Cfg::NBLK) {
            for (int j = grg; j < np; j += Cfg::RG) {
                const float4 A = *reinterpret_cast<const float4 *>(&S.f[j * Cfg::FSTRIDE + gba * 4]);
                const float4 B = *reinterpret_cast<const float4 *>(&S.f[j * Cfg::FSTRIDE + gbb * 4]);
                const double av[4] = {(double)A.x, (double)A.y, (double)A.z, (double)A.w};
                const double bv[4] = {(double)B.x, (double)B.y, (double)B.z, (double)B.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j2 = 0; j2 < 4; ++j2) st_m[i * 4 + j2] = fma(av[i], bv[j2], st_m[i * 4 + j2]);
            }
        }
    };

    // BWD: the upstream gradient, forward output and argmax rows of pillars [p0, p0 + n) -> smem, asynchronously
    auto prefetch_bwd = [&](int b, int p0, int n) {
        const uint32_t bytes = (uint32_t)n * COUT * 4;
        mbar_expect_tx(&S.pre[b], 2 * bytes);
        tma_bulk_g2s(S.pre_grad[b], a.grad + (size_t)p0 * COUT, bytes, &S.pre[b]);
        tma_bulk_g2s(S.pre_arg[b], a.argpos + (size_t)p0 * COUT, bytes, &S.pre[b]);
    };

    // thread 0 keeps tile_first two tiles ahead in registers so the TMA issue never waits on a global load
    int nfa = 0, nfb = 0;   // (tile_first[t], tile_first[t + 1]) of the tile after the next
    if (nk > 0 && tid == 0) {
        issue(tile_of(0), 0, a.tile_first[tile_of(0)], a.tile_first[tile_of(0) + 1]);
        if (nk > 1) { nfa = a.tile_first[tile_of(1)]; nfb = a.tile_first[tile_of(1) + 1]; }
    }
    uint32_t par0 = 0, par1 = 0;
    int cj = 0;             // BWD: chunks consumed so far (buffer cj & 1, phase (cj >> 1) & 1)
    bool pending = false;   // BWD: the first chunk of the tile being entered was requested during the previous tile

    for (int k = 0; k < nk; ++k) {
        const int t = tile_of(k);
        const int s = k & 1;
        if (tid == 0 && k + 1 < nk) {
            issue(tile_of(k + 1), s ^ 1, nfa, nfb);
            if (k + 2 < nk) { nfa = a.tile_first[tile_of(k + 2)]; nfb = a.tile_first[tile_of(k + 2) + 1]; }
        }
        if (s == 0) { mbar_wait(&S.full[0], par0); par0 ^= 1; } else { mbar_wait(&S.full[1], par1); par1 ^= 1; }
        PfnStage<Cfg> &T = S.st[s];
        const long long base = (long long)t * WIN;

        // ---- tile bounds straight from the pillar table (built once per forward by pillar_table_kernel)
        const int ps = S.tf[s][0], pe = S.tf[s][1];
        if (pe == ps) { __syncthreads(); continue; }  // a pillar from an earlier tile covers the whole window
        const int j0 = __float_as_int(T.aux[5]) - (int)base;                      // first row of the first pillar
        const float *alast = T.aux + (pe - ps - 1) * 8;
        const int last_start = __float_as_int(alast[5]) - (int)base, last_rows = __float_as_int(alast[6]);
        const bool big = last_start + last_rows > CAP;                             // the last pillar runs past the staged rows
        const int jstop = big ? last_start : last_start + last_rows, np = jstop - j0;
        const int nb = big ? pe - ps - 1 : pe - ps;
        const int gb = (int)base + j0;
        if (MODE == PFN_MODE_BWD) {
            if (!pending && nb > 0 && tid == 0) prefetch_bwd(cj & 1, ps, min(Smem::PCH, nb));
            if (pending && nb == 0) {   // only a big pillar starts here: drain the chunk that was requested for this tile
                mbar_wait(&S.pre[cj & 1], (uint32_t)(cj >> 1) & 1u);
                ++cj;
            }
            pending = false;
        }

        if (np > 0) {
            if (tid == 0) S.start[nb] = np;
            c1(T.rows, j0 + 1, np, ps, T.aux);
            __syncthreads();
            if (MODE == PFN_MODE_STATS) {
                gram(np);   // train-mode batch statistics come from the feature moments alone: no linear layer in this pass
            } else if (MODE != PFN_MODE_BWD) {
                // sub-group `sid` streams a pillar-aligned 1/NS of the rows
                const int r_lo = (sid * np) / NS, r_hi = ((sid + 1) * np) / NS;
                const int ra = (sid == 0) ? 0 : S.start[S.lp[r_lo] >> 1];
                const int rb = (sid == NS - 1) ? np : S.start[S.lp[r_hi] >> 1];
                stream(ra, rb, ps, gb);
            } else {
                // ---- E (backward): warp = pillar, lane = channel
                float tB[CPL], tA[CPL][CS];
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    tB[cc] = 0.0f;
#pragma unroll
                    for (int k = 0; k < CS; ++k) tA[cc][k] = 0.0f;
                }
                for (int q0 = 0; q0 < nb; q0 += Smem::PCH) {
                  const int nq = min(Smem::PCH, nb - q0);
                  if (q0 > 0) __syncthreads();   // every warp is done with the buffer the next request lands in
                  // request the chunk after this one -- the rest of this tile, else the head of the next tile -- so that one
                  // chunk is always in flight while another is folded in (the tile form was bound by exactly this latency)
                  if (q0 + Smem::PCH < nb) {
                      if (tid == 0) prefetch_bwd((cj + 1) & 1, ps + q0 + Smem::PCH, min(Smem::PCH, nb - q0 - Smem::PCH));
                  } else if (k + 1 < nk) {
                      const int ps1 = S.tf[s ^ 1][0], pe1 = S.tf[s ^ 1][1];
                      if (pe1 > ps1) {
                          pending = true;
                          if (tid == 0) prefetch_bwd((cj + 1) & 1, ps1, min(Smem::PCH, pe1 - ps1));
                      }
                  }
                  const int b = cj & 1;
                  mbar_wait(&S.pre[b], (uint32_t)(cj >> 1) & 1u);
                  ++cj;
                  for (int q = warp; q < nq; q += NW) {
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) {
                        const int o = q * COUT + lane + 32 * cc;
                        const int ap = S.pre_arg[b][o];   // negative: the forward marked the pillar as ReLU-clamped (:38)
                        const float gy = ap >= 0 ? S.pre_grad[b][o] : 0.0f;
                        const int jj = (ap >= 0 ? ap : ~ap) - gb;
                        float f[Cfg::FW];
                        const float4 *src = reinterpret_cast<const float4 *>(&S.f[jj * Cfg::FSTRIDE]);
#pragma unroll
                        for (int k4 = 0; k4 < (CS + 3) / 4; ++k4) {
                            const float4 v = src[k4];
                            f[k4 * 4] = v.x; f[k4 * 4 + 1] = v.y; f[k4 * 4 + 2] = v.z; f[k4 * 4 + 3] = v.w;
                        }
                        tB[cc] += gy;   // G = sum gy x = w . A (x is linear in f): folded in bwd_finalize_kernel
#pragma unroll
                        for (int k = 0; k < CS; ++k) tA[cc][k] = fmaf(gy, f[k], tA[cc][k]);
                    }
                  }
                }
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    dB[cc] += (double)tB[cc];
#pragma unroll
                    for (int k = 0; k < CS; ++k) dA[cc][k] += (double)tA[cc][k];
                }
            }
        }

        if (big) {
            // ---- big pillar pb = rows [a0, e): straight from global memory
            __syncthreads();
            const int pb = pe - 1;
            const long long a0 = base + last_start, e = a0 + last_rows;
            if (tid < 8) S.bigaux[tid] = alast[tid];
            if (tid < COUT) { S.carry_v[tid] = want_arg ? -1.0f : 0.0f; S.carry_k[tid] = INF; S.carry_p[tid] = 0; }
            __syncthreads();
            if (MODE == PFN_MODE_BWD) {
                if (warp == 0) {
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) {
                        const size_t o = (size_t)pb * COUT + lane + 32 * cc;
                        const int ap = a.argpos[o];
                        const float gy = ap >= 0 ? a.grad[o] : 0.0f;
                        const float *src = a.grows + ((size_t)(ap >= 0 ? ap : ~ap) + 1) * RS;
                        float r[COLS], f[Cfg::FW];
#pragma unroll
                        for (int c = 0; c < COLS; ++c) r[c] = src[c];
                        decorate<Cfg>(r, S.bigaux[3], S.bigaux[4], S.bigaux, a, f);
                        dB[cc] += (double)gy;
#pragma unroll
                        for (int k = 0; k < CS; ++k) dA[cc][k] = fma((double)gy, (double)f[k], dA[cc][k]);
                    }
                }
            } else {
                float *rows = T.rows;  // this stage's row buffer is free again (the next TMA targets the other stage)
                for (long long cs = a0; cs < e; cs += CAP) {
                    const int npc = (int)min((long long)CAP, e - cs);
                    for (int i = tid; i < npc * RS; i += NT) rows[i] = a.grows[(cs + 1) * RS + i];
                    __syncthreads();
                    c1(rows, 0, npc, -1, nullptr);
                    __syncthreads();
                    if (MODE == PFN_MODE_STATS) gram(npc);
                    else stream((sid * npc) / NS, ((sid + 1) * npc) / NS, pb, (int)cs);
                    if (is_apply) {
                        // merge the sub-groups' running maxima into the carry, in stream order (deterministic)
                        for (int w = 0; w < NS; ++w) {
                            if (sid == w) {
#pragma unroll
                                for (int cc = 0; cc < WN; ++cc) {
                                    const int ch = sl + LPR * cc;
                                    const float bm = S.carry_v[ch];
                                    const int bk = S.carry_k[ch];
                                    if (m[cc] > bm || (want_arg && m[cc] == bm && mk[cc] < bk)) {
                                        S.carry_v[ch] = m[cc]; S.carry_k[ch] = mk[cc]; S.carry_p[ch] = mp[cc];
                                    }
                                }
                            }
                            __syncthreads();
                        }
                    } else {
                        __syncthreads();
                    }
                }
                if (is_apply && tid < COUT) {
                    a.features[(size_t)pb * COUT + tid] = S.carry_v[tid];
                    if (want_arg) a.argpos[(size_t)pb * COUT + tid] = (S.carry_v[tid] > 0.0f) ? S.carry_p[tid] : ~S.carry_p[tid];
                }
                fence_proxy_async();  // generic-proxy writes to the stage buffer before a later TMA reuses it
            }
        }
        __syncthreads();
    }

    // ---- per-CTA partial sums
    if (MODE == PFN_MODE_STATS) {
        // layout: NBLK blocks of 16
        double *out = a.partials + (size_t)blockIdx.x * Cfg::STATS_DOUBLES;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) dscr[tid * 16 + e] = st_m[e];
        __syncthreads();
        for (int e = tid; e < Cfg::NBLK * 16; e += NT) {
            const int blk = e / 16, el = e % 16;
            double sacc = 0.0;
            for (int g = 0; g < Cfg::RG; ++g) sacc += dscr[(blk * Cfg::RG + g) * 16 + el];
            out[e] = sacc;
        }
    }
    if (MODE == PFN_MODE_BWD) {
        // layout: per channel [dbeta | G | A(CS)]
        constexpr int PER = Cfg::BWD_PER;
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            double *dst = dscr + ((size_t)warp * COUT + lane + 32 * cc) * PER;
            dst[0] = dB[cc]; dst[1] = 0.0;
#pragma unroll
            for (int k = 0; k < CS; ++k) dst[2 + k] = dA[cc][k];
        }
        __syncthreads();
        double *out = a.partials + (size_t)blockIdx.x * Cfg::BWD_DOUBLES;
        for (int e = tid; e < COUT * PER; e += NT) {
            double sacc = 0.0;
            for (int w = 0; w < NW; ++w) sacc += dscr[(size_t)w * COUT * PER + e];
            out[e] = sacc;
        }
    }
}

// ------------------------------------------------------------------------------------------- BN finalize (train)
// Fixed-order sum of the per-CTA partial vectors (deterministic): CTA b of the launch owns elements [32 b, 32 b + 32);
// lane = element, warp w adds producer CTAs w, w + 8, ..., the 8 warp sums are combined in warp order.  Returns true in
// the CTA that finishes last (ticket counter `done`, reset for the next launch): `totals` is then complete and that CTA
// goes on to the closed-form epilogue -- one launch instead of a reduction kernel followed by a finalize kernel.
__device__ __forceinline__ bool reduce_partials_last(const double *__restrict__ partials, int nblocks, int el, double *__restrict__ totals,
                                                     int32_t *done) {
    __shared__ double sm[8][32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, e = blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (e < el)
        for (int b = warp; b < nblocks; b += 8) acc += partials[(size_t)b * el + e];
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && e < el) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w][lane];
        totals[e] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(done, 1);
        s_last = (ticket == (int)gridDim.x - 1);
        if (s_last) *done = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Reduces the per-CTA partials in a fixed order (deterministic), folds the batch statistics into scale/shift,
// updates the running statistics, and expands the feature moments for the backward.
// bn_state = [mean(COUT) | var(COUT) | scale(COUT) | shift(COUT) | n | S1(CIN) | S2(CIN*CIN)]
template <class Cfg>
__global__ void __launch_bounds__(256) bn_finalize_kernel(const __grid_constant__ PfnArgs a, const double *partials, int nblocks,
                                                         double *totals, int32_t *done, double *bn_state,
                                                         float *running_mean, float *running_var, double momentum,
                                                         long long *num_batches_tracked) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, TOT = Cfg::STATS_DOUBLES, T4 = Cfg::T4;
    __shared__ double tot[TOT];
    __shared__ double part_m[COUT][8], part_e[COUT][8];
    const int tid = threadIdx.x;
    if (!reduce_partials_last(partials, nblocks, TOT, totals, done)) return;
    const long long N = a.counters[RDP_CNT_N];
    for (int e = tid; e < TOT; e += blockDim.x) tot[e] = totals[e];
    __syncthreads();
    const int cin = a.c_in;
    // Gram entry of super features (fa <= fb): block (fa/4, fb/4), element (fa%4, fb%4); column CS is the ones column
    auto gram_at = [&](int fa, int fb) {
        const int ba = fa / 4, bb = fb / 4;
        const int blk = ba * T4 - ba * (ba - 1) / 2 + (bb - ba);
        return tot[blk * 16 + (fa % 4) * 4 + (fb % 4)];
    };
    // batch moments of x = W f from the feature moments (oracle: orc_bn_batch_stats_moments):
    //   mean_c = w_c . S1 / n,  E[x^2]_c = w_c^T S2 w_c / n,  var_c = E[x^2]_c - mean_c^2  (biased)
    // 8 threads per channel take the rows fa = j, j + 8, ... of the quadratic form; the 8 partial sums are added in order.
    for (int item = tid; item < COUT * 8; item += blockDim.x) {
        const int c = item >> 3, j = item & 7;
        double m = 0.0, e2 = 0.0;
        for (int fa = j; fa < CS; fa += 8) {
            const int ka = a.kmap[fa];
            if (ka < 0) continue;
            const double wa = (double)a.weight[c * cin + ka];
            m += wa * gram_at(fa, CS);
            double row = 0.0;
            for (int fb = 0; fb < CS; ++fb) {
                const int kb = a.kmap[fb];
                if (kb >= 0) row += (fa <= fb ? gram_at(fa, fb) : gram_at(fb, fa)) * (double)a.weight[c * cin + kb];
            }
            e2 += wa * row;
        }
        part_m[c][j] = m;
        part_e[c][j] = e2;
    }
    __syncthreads();
    if (tid < COUT) {
        const double n = (double)N;
        double mean = 0.0, var = 0.0;
        if (N > 0) {
            double m = 0.0, e2 = 0.0;
            for (int j = 0; j < 8; ++j) { m += part_m[tid][j]; e2 += part_e[tid][j]; }
            mean = __ddiv_rn(m, n);
            var = __dsub_rn(__ddiv_rn(e2, n), __dmul_rn(mean, mean));
            if (!(var > 0.0)) var = 0.0;
        }
        bn_state[tid] = mean;
        bn_state[COUT + tid] = var;
        float sc, sh;
        fold_bn((double)a.gamma[tid], (double)a.beta[tid], mean, var, a.eps, &sc, &sh);
        bn_state[2 * COUT + tid] = (double)sc;
        bn_state[3 * COUT + tid] = (double)sh;
        if (N > 1) {  // torch raises for N == 1 and leaves the buffers alone for N == 0
            running_mean[tid] = (float)((1.0 - momentum) * (double)running_mean[tid] + momentum * mean);
            running_var[tid] = (float)((1.0 - momentum) * (double)running_var[tid] + momentum * var * (n / (n - 1.0)));
        }
    }
    if (tid == 0) {
        bn_state[4 * COUT] = (double)N;
        if (num_batches_tracked) *num_batches_tracked += 1;   // BatchNorm1d counts every train-mode forward (:29)
    }
    double *S1 = bn_state + 4 * COUT + 1, *S2 = S1 + cin;
    for (int e = tid; e < CS * CS; e += blockDim.x) {
        const int fa = e / CS, fb = e % CS;
        const int ka = a.kmap[fa], kb = a.kmap[fb];
        if (ka >= 0 && kb >= 0) S2[ka * cin + kb] = fa <= fb ? gram_at(fa, fb) : gram_at(fb, fa);
    }
    for (int fa = tid; fa < CS; fa += blockDim.x)
        if (a.kmap[fa] >= 0) S1[a.kmap[fa]] = gram_at(fa, CS);
}

// One CTA: fixed-order reduction of the backward partials, then the closed-form BatchNorm backward:
//   dgamma_c = (G_c - mu_c dbeta_c) / sigma_c
//   dW_ck    = (gamma_c/sigma_c) [ A_ck - dbeta_c/N S1_k - dgamma_c/N ((S2 w_c)_k - mu_c S1_k)/sigma_c ]   (train)
//   dW_ck    = (gamma_c/sigma_c) A_ck                                                                    (eval BN)
template <class Cfg>
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const __grid_constant__ PfnArgs a, const double *partials, int nblocks,
                                                          double *totals, int32_t *done, const double *bn_state,
                                                          int train_bn, float *d_weight, float *d_gamma, float *d_beta) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, PER = Cfg::BWD_PER;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tot = reinterpret_cast<double *>(smem_raw);  // COUT*PER
    double *dgam = tot + COUT * PER;                       // COUT
    const int tid = threadIdx.x, cin = a.c_in;
    if (!reduce_partials_last(partials, nblocks, COUT * PER, totals, done)) return;
    for (int e = tid; e < COUT * PER; e += blockDim.x) tot[e] = totals[e];
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
        // G_c = sum_p gy x_win = w_c . A_c  (x = W f is linear in the features), in fp64
        double G = 0.0;
        for (int s = 0; s < CS; ++s) {
            const int k = a.kmap[s];
            if (k >= 0) G = fma((double)a.weight[tid * cin + k], tot[tid * PER + 2 + s], G);
        }
        const double db = tot[tid * PER];
        const double dg = (G - mu * db) * is;
        dgam[tid] = dg;
        d_beta[tid] = (float)db;
        if (a.use_norm && d_gamma) d_gamma[tid] = (float)dg;
    }
    __syncthreads();
    const double *S1 = bn_state ? bn_state + 4 * COUT + 1 : nullptr, *S2 = S1 ? S1 + cin : nullptr;
    for (int e = tid; e < COUT * CS; e += blockDim.x) {
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
