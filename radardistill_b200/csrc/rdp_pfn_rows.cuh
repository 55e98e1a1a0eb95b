// rdp_pfn_rows.cuh -- forward PFN kernel, "thread = row" form (eval APPLY and train APPLY_ARG).
//
// Same contract and bit-identical results as pfn_tile_kernel<APPLY / APPLY_ARG> (rdp_pfn.cuh): decorated features
// (dynamic_pillar_vfe.py:214-237), Linear + folded BatchNorm + ReLU and scatter_max (+argmax) of PFNLayerV2.forward
// (:35-46), in the canonical arithmetic of oracle/pillar_oracle.c (k-ascending fmaf chain, y = fma(x, scale, shift)).
//
// Why a second form: with lane = channel the feature rows are broadcast from shared memory to the channel lanes, and
// that kernel is bound by shared-memory wavefronts and instruction issue (ncu: 63 % issue slots, 41 % of them FFMA).
// Here a thread owns whole rows (two of them), so the only operand that has to be fetched per FFMA group is a weight
// quad -- one uniform-address LDS.128 feeds 4 x 2 FFMAs -- and the per-row bookkeeping of the stream disappears:
//
//   A  thread t of a 128-thread CTA takes rows t and t + 128 of a 256-row chunk of the pillar-grouped row array:
//      row (32 B) + its pillar's table entry straight from global memory (coalesced / L1), decorate in registers,
//      x_c = sum_k fmaf(W[c][k], f[k], x_c), y_c = fma(x_c, scale_c, shift_c) for all channels, y -> smem tile
//      X[row][channel]; pillar head / last-row flags by warp ballot.
//   B  segmented max over the rows of each pillar: thread = (row group, channel quad) walks the pillars that START in
//      its 16 rows (LDS.128 + 4 FMNMX per row) and writes each pillar's 128-byte feature row with 16-byte stores
//      (8 lanes = one full line).  A pillar left open at the end of the chunk is carried in shared memory.
//
// A persistent CTA owns a contiguous, pillar-aligned range of rows (the PFN tiles of tile_first) and walks it chunk by
// chunk, so a pillar of any length is handled by the carry; there is no separate big-pillar path.
#pragma once

#include "rdp_pfn.cuh"

namespace rdp {

constexpr int kRowsThreads = 128;
constexpr int kRowsPerThread = 2;
constexpr int kRowsChunk = kRowsThreads * kRowsPerThread;
#ifndef RDP_ROWS_GRID_PER_SM
#define RDP_ROWS_GRID_PER_SM 5
#endif
constexpr int kRowsGridCap = 148 * RDP_ROWS_GRID_PER_SM;

template <class Cfg, bool ARG>
struct RowsSmem {
    static constexpr int WS = (Cfg::CS + 2 + 3) / 4 * 4;  // floats per channel: weights | scale | shift | pad
    static constexpr int XS = Cfg::COUT + 4;              // tile row stride: 16-byte stores / loads stay conflict free
    alignas(16) float w[Cfg::COUT * WS];
    alignas(16) float x[kRowsChunk * XS];
    int gid[kRowsChunk];
    int kept[ARG ? kRowsChunk : 1];
    uint32_t heads[kRowsChunk / 32], lasts[kRowsChunk / 32];
    // running max of a pillar left open at the end of a chunk; double buffered by chunk parity (one group may still be
    // reading the incoming carry while another writes the outgoing one)
    alignas(16) float carry_v[2][Cfg::COUT];
    alignas(16) int carry_k[2][ARG ? Cfg::COUT : 4], carry_p[2][ARG ? Cfg::COUT : 4];
};

template <class Cfg, bool ARG>
__global__ void __launch_bounds__(kRowsThreads, RDP_ROWS_GRID_PER_SM) pfn_rows_kernel(const __grid_constant__ PfnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = RowsSmem<Cfg, ARG>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, COLS = Cfg::COLS, RS = Cfg::RS, WS = Smem::WS, XS = Smem::XS;
    constexpr int NT = kRowsThreads, R = kRowsPerThread, CHUNK = kRowsChunk, WIN = kPfnWin, INF = 0x7fffffff;
    constexpr int QUADS = COUT / 4, GROUPS = NT / QUADS, RPG = CHUNK / GROUPS;  // phase B: rows per (row group)
    static_assert(RPG == 16 || RPG == 32, "row groups must align with the 32-bit flag words");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = a.counters[RDP_CNT_N];
    const int P = a.counters[RDP_CNT_P];
    const bool none_dropped = (N == a.n0);

    // ---- weights, folded BatchNorm scale / shift -> smem  (w[c][0..CS) | scale | shift)
    for (int e = tid; e < COUT * WS; e += NT) {
        const int c = e / WS, s = e % WS;
        float v = 0.0f;
        if (s < CS) {
            const int k = a.kmap[s];
            v = (k >= 0) ? __ldg(a.weight + c * a.c_in + k) : 0.0f;
        } else if (s == CS || s == CS + 1) {
            float sc = 1.0f, sh = a.bias ? a.bias[c] : 0.0f;
            if (a.use_norm) {
                if (a.fold_from_state) { sc = (float)a.bn_state[2 * COUT + c]; sh = (float)a.bn_state[3 * COUT + c]; }
                else fold_bn((double)a.gamma[c], (double)a.beta[c], (double)a.rmean[c], (double)a.rvar[c], a.eps, &sc, &sh);
            }
            v = (s == CS) ? sc : sh;
        }
        S.w[e] = v;
    }

    // ---- this CTA's rows: PFN tiles [t_begin, t_end) -> grouped rows [row_begin, row_end), both pillar boundaries
    const int ntiles = (int)((N + WIN - 1) / WIN);
    const int per = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = min(ntiles, (int)blockIdx.x * per), t_end = min(ntiles, t_begin + per);
    auto tile_row = [&](int t) -> long long {
        if (t >= ntiles) return N;
        const int pf = a.tile_first[t];
        return pf >= P ? N : (long long)__float_as_int(__ldg(a.aux + (size_t)pf * 8 + 5));
    };
    const long long row_begin = (t_begin < t_end) ? tile_row(t_begin) : 0, row_end = (t_begin < t_end) ? tile_row(t_end) : 0;
    __syncthreads();

    int par = 0;
    for (long long c0 = row_begin; c0 < row_end; c0 += CHUNK, par ^= 1) {
        const int nrow = (int)min((long long)CHUNK, row_end - c0);

        // =========================================================================== A: thread = row
        float f[R][Cfg::FW];
        bool valid[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = tid + i * NT;
            const long long g = c0 + r;
            valid[i] = r < nrow;
            bool head = false, last = false;
            if (valid[i]) {
                float row[RS];
                const float4 *src = reinterpret_cast<const float4 *>(a.grows + (size_t)(g + 1) * RS);
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) {
                    const float4 v = src[q];
                    row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
                }
                const int gid = __float_as_int(row[RS - 1]);
                const int gprev = __float_as_int(a.grows[(size_t)g * RS + RS - 1]);       // row -1 is the sentinel (pillar -1)
                const int gnext = (g + 1 < N) ? __float_as_int(a.grows[(size_t)(g + 2) * RS + RS - 1]) : -1;
                head = gid != gprev;
                last = gid != gnext;
                const float4 *ax = reinterpret_cast<const float4 *>(a.aux + (size_t)gid * 8);
                const float4 m4 = ax[0], c4 = ax[1];  // [mean x y z | centre x] [centre y | start | rows | 0]
                const float mean[3] = {m4.x, m4.y, m4.z};
                decorate<Cfg>(row, m4.w, c4.x, mean, a, f[i]);
                S.gid[r] = gid;
                if (ARG) {
                    const int orig = __float_as_int(row[RS - 2]);
                    S.kept[r] = none_dropped ? orig : a.orig2kept[orig];
                }
            } else {
#pragma unroll
                for (int k = 0; k < CS; ++k) f[i][k] = 0.0f;
            }
            const uint32_t hb = __ballot_sync(0xffffffffu, head), lb = __ballot_sync(0xffffffffu, last);
            if (lane == 0) { S.heads[i * (NT / 32) + warp] = hb; S.lasts[i * (NT / 32) + warp] = lb; }
        }

        // x = W f (k-ascending fmaf chain), y = fma(x, scale, shift) [ARG: z = max(y, 0)]  -> X[row][channel]
        auto linear = [&](auto rows_tag) {
            constexpr int RR = decltype(rows_tag)::value;
#pragma unroll
            for (int c4 = 0; c4 < QUADS; ++c4) {
                float y[RR][4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = c4 * 4 + cc;
                    float acc[RR];
#pragma unroll
                    for (int i = 0; i < RR; ++i) acc[i] = 0.0f;
                    float wv[WS];
#pragma unroll
                    for (int k4 = 0; k4 < WS / 4; ++k4) {
                        const float4 w = *reinterpret_cast<const float4 *>(&S.w[c * WS + k4 * 4]);  // uniform address: broadcast
                        wv[4 * k4] = w.x; wv[4 * k4 + 1] = w.y; wv[4 * k4 + 2] = w.z; wv[4 * k4 + 3] = w.w;
                    }
#pragma unroll
                    for (int k = 0; k < CS; ++k)
#pragma unroll
                        for (int i = 0; i < RR; ++i) acc[i] = fmaf(wv[k], f[i][k], acc[i]);
#pragma unroll
                    for (int i = 0; i < RR; ++i) {
                        const float yy = fmaf(acc[i], wv[CS], wv[CS + 1]);
                        y[i][cc] = ARG ? fmaxf(yy, 0.0f) : yy;
                    }
                }
#pragma unroll
                for (int i = 0; i < RR; ++i)
                    *reinterpret_cast<float4 *>(&S.x[(tid + i * NT) * XS + c4 * 4]) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
            }
        };
        if (__any_sync(0xffffffffu, valid[R - 1])) linear(std::integral_constant<int, R>{});
        else if (__any_sync(0xffffffffu, valid[0])) linear(std::integral_constant<int, 1>{});
        __syncthreads();

        // =========================================================================== B: thread = (row group, quad)
        {
            const int g = tid / QUADS, q = tid % QUADS;
            const int gr0 = g * RPG;
            uint32_t own = S.heads[gr0 >> 5] >> (gr0 & 31);
            if (RPG < 32) own &= (1u << RPG) - 1u;
            const bool carry_in = (g == 0) && !(S.heads[0] & 1u);   // the chunk starts inside a pillar: continue it from the carry
            int r = -1;
            if (carry_in) r = 0;
            else if (own) r = gr0 + __ffs(own) - 1;
            if (r >= 0 && r < nrow) {
                float m[4];
                int mk[4], mp[4];
                if (carry_in) {
                    const float4 cv = *reinterpret_cast<const float4 *>(&S.carry_v[par][4 * q]);
                    m[0] = cv.x; m[1] = cv.y; m[2] = cv.z; m[3] = cv.w;
                    if (ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { mk[e] = S.carry_k[par][4 * q + e]; mp[e] = S.carry_p[par][4 * q + e]; }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { m[e] = ARG ? -1.0f : 0.0f; mk[e] = INF; mp[e] = 0; }
                }
                int pid = S.gid[r];
                const int stop = gr0 + RPG;   // heads at or after this row belong to later groups
                uint32_t lw = S.lasts[r >> 5];
                bool open = true;
                for (;;) {
                    const float4 v = *reinterpret_cast<const float4 *>(&S.x[r * XS + 4 * q]);
                    const float vv[4] = {v.x, v.y, v.z, v.w};
                    if (!ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) m[e] = fmaxf(m[e], vv[e]);   // ReLU folds into the max with 0
                    } else {
                        const int kj = S.kept[r], pos = (int)c0 + r;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (vv[e] > m[e] || (vv[e] == m[e] && kj < mk[e])) { m[e] = vv[e]; mk[e] = kj; mp[e] = pos; }
                    }
                    const bool is_last = (lw >> (r & 31)) & 1u;
                    if (is_last) {
                        *reinterpret_cast<float4 *>(a.features + (size_t)pid * COUT + 4 * q) = make_float4(m[0], m[1], m[2], m[3]);
                        if (ARG) *reinterpret_cast<int4 *>(a.argpos + (size_t)pid * COUT + 4 * q) = make_int4(mp[0], mp[1], mp[2], mp[3]);
                        ++pid;
#pragma unroll
                        for (int e = 0; e < 4; ++e) { m[e] = ARG ? -1.0f : 0.0f; mk[e] = INF; mp[e] = 0; }
                        open = false;
                        if (r + 1 >= stop) break;
                    } else {
                        open = true;
                    }
                    if (++r >= nrow) break;
                    if ((r & 31) == 0) lw = S.lasts[r >> 5];
                }
                if (open) {   // the pillar continues in the next chunk of this CTA
                    *reinterpret_cast<float4 *>(&S.carry_v[par ^ 1][4 * q]) = make_float4(m[0], m[1], m[2], m[3]);
                    if (ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { S.carry_k[par ^ 1][4 * q + e] = mk[e]; S.carry_p[par ^ 1][4 * q + e] = mp[e]; }
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace rdp
