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
//   A  thread t of a 128-thread CTA takes rows t and t + 128 of a 256-row window of the pillar-grouped row array.  The
//      window's rows (+ one row either side) and the table entries of the pillars that intersect it (tile_first gives
//      the first one) arrive by two TMA bulk copies issued one window ahead, so no thread ever waits on a global
//      load; decorate in registers,
//      x_c = sum_k fmaf(W[c][k], f[k], x_c), z_c = max(fma(x_c, scale_c, shift_c), 0) for all channels; pillar head /
//      last-row flags by warp ballot.  A row that is a whole pillar (two thirds of the LiDAR pillars) stores its
//      feature row directly (256-bit stores, one full sector each); the other rows go to the smem tile X[row][channel].
//   B  segmented max over the rows of each multi-row pillar: thread = (row group, channel quad) walks the pillars that
//      START in its 16 rows (LDS.128 + 4 FMNMX per row) and writes each pillar's 128-byte feature row with 16-byte
//      stores (8 lanes = one full line).  A pillar left open at the end of the chunk is carried in shared memory.
//
// A persistent CTA owns a contiguous, pillar-aligned range of rows (the PFN tiles of tile_first) and walks it chunk by
// chunk, so a pillar of any length is handled by the carry; there is no separate big-pillar path.
#pragma once

#include "rdp_pfn.cuh"

namespace rdp {

constexpr int kRowsThreads = 128;
constexpr int kRowsPerThread = 2;
constexpr int kRowsChunk = kRowsThreads * kRowsPerThread;   // 256 rows = two PFN tile windows
static_assert(kRowsChunk == 2 * kPfnWin, "tile_first is indexed per 128-row window");
#ifndef RDP_ROWS_GRID_PER_SM
#define RDP_ROWS_GRID_PER_SM 3
#endif
constexpr int kRowsGridCap = 148 * RDP_ROWS_GRID_PER_SM;

template <class Cfg, bool ARG>
struct RowsSmem {
    static constexpr int WS = (Cfg::CS + 2 + 3) / 4 * 4;  // floats per channel: weights | scale | shift | pad
    static constexpr int XS = Cfg::COUT + 4;              // tile row stride: 16-byte stores / loads stay conflict free
    static constexpr int SROWS = kRowsChunk + 2;          // staged rows: the one in front of the window, the window, the one behind
    static constexpr int SAUX = kRowsChunk + 2;           // staged table entries: every pillar that intersects the window
    alignas(128) float rows[2][SROWS * Cfg::RS];
    alignas(128) float aux[2][SAUX * 8];
    alignas(16) float w[Cfg::COUT * WS];
    alignas(16) float x[kRowsChunk * XS];
    alignas(8) uint64_t full[2];
    int pa[2];                                            // first staged pillar of each stage
    int kept[ARG ? kRowsChunk : 1];
    uint32_t heads[kRowsChunk / 32], lasts[kRowsChunk / 32];
    // running max of a pillar left open at the end of a window; double buffered by window parity (one group may still
    // be reading the incoming carry while another writes the outgoing one)
    alignas(16) float carry_v[2][Cfg::COUT];
    alignas(16) int carry_k[2][ARG ? Cfg::COUT : 4], carry_p[2][ARG ? Cfg::COUT : 4];
};

__device__ __forceinline__ void st_global_v8(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

template <class Cfg, bool ARG>
__global__ void __launch_bounds__(kRowsThreads, RDP_ROWS_GRID_PER_SM) pfn_rows_kernel(const __grid_constant__ PfnArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = RowsSmem<Cfg, ARG>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, COLS = Cfg::COLS, RS = Cfg::RS, WS = Smem::WS, XS = Smem::XS;
    constexpr int NT = kRowsThreads, R = kRowsPerThread, CHUNK = kRowsChunk, INF = 0x7fffffff;
    constexpr int QUADS = COUT / 4, GROUPS = NT / QUADS, RPG = CHUNK / GROUPS;  // phase B: rows per (row group)
    static_assert(RPG == 8 || RPG == 16 || RPG == 32, "row groups must align with the 32-bit flag words");
    static_assert(COUT % 8 == 0 && COLS <= RS - 2, "layout");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = a.counters[RDP_CNT_N];
    const int P = a.counters[RDP_CNT_P];
    const bool none_dropped = (N == a.n0);

    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        fence_mbar_init();
    }
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

    // ---- this CTA's rows: windows [u_begin, u_end) own the pillars that START in them -> grouped rows
    //      [row_begin, row_end), both pillar boundaries; they are walked window by window (256-row aligned)
    const int ntiles = (int)((N + kPfnWin - 1) / kPfnWin);
    const int nwin = (int)((N + CHUNK - 1) / CHUNK);
    const int per = (nwin + (int)gridDim.x - 1) / (int)gridDim.x;
    const int u_begin = min(nwin, (int)blockIdx.x * per), u_end = min(nwin, u_begin + per);
    auto tile_row = [&](int t) -> long long {
        if (t >= ntiles) return N;
        const int pf = a.tile_first[t];
        return pf >= P ? N : (long long)__float_as_int(__ldg(a.aux + (size_t)pf * 8 + 5));
    };
    const long long row_begin = (u_begin < u_end) ? tile_row(2 * u_begin) : 0, row_end = (u_begin < u_end) ? tile_row(2 * u_end) : 0;
    const int u0 = (int)(row_begin / CHUNK);                                            // first window holding rows of mine
    const int u1 = (row_end > row_begin) ? (int)((row_end + CHUNK - 1) / CHUNK) : u0;    // one past the last

    // thread 0: TMA of window u into stage s.  pa = first pillar staged = the one covering the window's first row.
    auto issue = [&](int u, int s, int pa) {
        const long long g0 = (long long)u * CHUNK;
        const uint32_t nrows = (uint32_t)min((long long)Smem::SROWS, N + 2 - g0);   // array rows g0 .. (row g0 - 1 is the sentinel / neighbour)
        const uint32_t naux = (uint32_t)min(Smem::SAUX, P - pa);
        S.pa[s] = pa;
        mbar_expect_tx(&S.full[s], nrows * RS * 4 + naux * 32);
        tma_bulk_g2s(S.rows[s], a.grows + (size_t)g0 * RS, nrows * RS * 4, &S.full[s]);
        tma_bulk_g2s(S.aux[s], a.aux + (size_t)pa * 8, naux * 32, &S.full[s]);
    };
    auto first_pillar = [&](int u) { return max(min(a.tile_first[2 * u], P) - 1, 0); };
    int pa_next = 0;
    if (tid == 0 && u0 < u1) {
        issue(u0, 0, first_pillar(u0));
        if (u0 + 1 < u1) pa_next = first_pillar(u0 + 1);
    }
    __syncthreads();

    uint32_t par0 = 0, par1 = 0;
    for (int u = u0; u < u1; ++u) {
        const int s = (u - u0) & 1, par = s;
        if (tid == 0 && u + 1 < u1) {
            issue(u + 1, s ^ 1, pa_next);
            if (u + 2 < u1) pa_next = first_pillar(u + 2);
        }
        if (s == 0) { mbar_wait(&S.full[0], par0); par0 ^= 1; } else { mbar_wait(&S.full[1], par1); par1 ^= 1; }
        const long long c0 = (long long)u * CHUNK;
        const int lo = (int)(max(c0, row_begin) - c0), hi = (int)(min(c0 + CHUNK, row_end) - c0);   // my rows of this window
        const float *srows = S.rows[s];
        const float *saux = S.aux[s];
        const int pa = S.pa[s];

        // =========================================================================== A: thread = row
        float f[R][Cfg::FW];
        bool valid[R], single[R];
        int gidv[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = tid + i * NT;
            valid[i] = r >= lo && r < hi;
            bool head = false, last = false;
            gidv[i] = 0;
            if (valid[i]) {
                float row[RS];
                const float4 *src = reinterpret_cast<const float4 *>(srows + (r + 1) * RS);
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) {
                    const float4 v = src[q];
                    row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
                }
                const int gid = __float_as_int(row[RS - 1]);
                gidv[i] = gid;
                head = gid != __float_as_int(srows[r * RS + RS - 1]);                       // the row in front (sentinel: pillar -1)
                last = (c0 + r + 1 >= N) || gid != __float_as_int(srows[(r + 2) * RS + RS - 1]);
                const float4 *ax = reinterpret_cast<const float4 *>(saux + (gid - pa) * 8);   // [mean x y z | centre x] [centre y | ...]
                const float4 m4 = ax[0];
                const float ceny = saux[(gid - pa) * 8 + 4];
                const float mean[3] = {m4.x, m4.y, m4.z};
                decorate<Cfg>(row, m4.w, ceny, mean, a, f[i]);
                if (ARG) {
                    const int orig = __float_as_int(row[RS - 2]);
                    S.kept[r] = none_dropped ? orig : a.orig2kept[orig];
                }
            } else {
#pragma unroll
                for (int k = 0; k < CS; ++k) f[i][k] = 0.0f;
            }
            single[i] = head && last;   // the row is a whole pillar: its feature row leaves from here
            const uint32_t hb = __ballot_sync(0xffffffffu, head), lb = __ballot_sync(0xffffffffu, last);
            if (lane == 0) { S.heads[i * (NT / 32) + warp] = hb; S.lasts[i * (NT / 32) + warp] = lb; }
        }

        // x = W f (k-ascending fmaf chain), z = max(fma(x, scale, shift), 0): single-row pillars -> global, others -> X
        auto linear = [&](auto rows_tag) {
            constexpr int RR = decltype(rows_tag)::value;
#pragma unroll
            for (int c8 = 0; c8 < COUT / 8; ++c8) {
                float y[RR][8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = c8 * 8 + cc;
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
                    for (int i = 0; i < RR; ++i) y[i][cc] = fmaxf(fmaf(acc[i], wv[CS], wv[CS + 1]), 0.0f);
                }
#pragma unroll
                for (int i = 0; i < RR; ++i) {
                    if (single[i]) {   // the only row of its pillar wins every channel
                        st_global_v8(a.features + (size_t)gidv[i] * COUT + c8 * 8, y[i]);
                        if (ARG) {
                            const int pos = (int)c0 + tid + i * NT;
                            float pv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) pv[e] = __int_as_float(y[i][e] > 0.0f ? pos : ~pos);
                            st_global_v8(reinterpret_cast<float *>(a.argpos) + (size_t)gidv[i] * COUT + c8 * 8, pv);
                        }
                    } else {
                        float4 *dst = reinterpret_cast<float4 *>(&S.x[(tid + i * NT) * XS + c8 * 8]);
                        dst[0] = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
                        dst[1] = make_float4(y[i][4], y[i][5], y[i][6], y[i][7]);
                    }
                }
            }
        };
        if (__any_sync(0xffffffffu, valid[R - 1])) linear(std::integral_constant<int, R>{});
        else if (__any_sync(0xffffffffu, valid[0])) linear(std::integral_constant<int, 1>{});
        __syncthreads();

        // =========================================================================== B: thread = (row group, quad)
        {
            const int g = tid / QUADS, q = tid % QUADS;
            const int gr0 = g * RPG;
            uint32_t hw = S.heads[gr0 >> 5] >> (gr0 & 31), lw = S.lasts[gr0 >> 5] >> (gr0 & 31);
            constexpr uint32_t gmask = RPG < 32 ? ((1u << (RPG & 31)) - 1u) : 0xffffffffu;
            hw &= gmask; lw &= gmask;
            uint32_t own = hw & ~lw;                                 // heads of multi-row pillars that start in my rows
            bool carry_in = (g == 0) && lo == 0 && hi > 0 && !(S.heads[0] & 1u);   // the window starts inside a pillar of mine: continue it
            // rows from r through the first last-row flag at or after r; -1 if the pillar is still open at the end of the window
            auto seg_len = [&](int r) -> int {
                int w = r >> 5;
                uint32_t b = S.lasts[w] >> (r & 31);
                if (b) return __ffs(b);
                int len = 32 - (r & 31);
                for (++w; w < CHUNK / 32; ++w) {
                    b = S.lasts[w];
                    if (b) return len + __ffs(b);
                    len += 32;
                }
                return -1;
            };
            while (carry_in || own) {
                int r;
                float m[4];
                int mk[4], mp[4];
                if (carry_in) {
                    r = 0;
                    const float4 cv = *reinterpret_cast<const float4 *>(&S.carry_v[par][4 * q]);
                    m[0] = cv.x; m[1] = cv.y; m[2] = cv.z; m[3] = cv.w;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { mk[e] = ARG ? S.carry_k[par][4 * q + e] : INF; mp[e] = ARG ? S.carry_p[par][4 * q + e] : 0; }
                    carry_in = false;
                } else {
                    const int bit = __ffs(own) - 1;
                    own &= own - 1;
                    r = gr0 + bit;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { m[e] = ARG ? -1.0f : 0.0f; mk[e] = INF; mp[e] = 0; }
                }
                const int pid = __float_as_int(srows[(r + 1) * RS + RS - 1]);
                int len = seg_len(r);
                const bool open = len < 0;
                if (open) len = hi - r;
                const float *xp = &S.x[r * XS + 4 * q];
                for (int k = 0; k < len; ++k, xp += XS) {
                    const float4 v = *reinterpret_cast<const float4 *>(xp);
                    const float vv[4] = {v.x, v.y, v.z, v.w};
                    if (!ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) m[e] = fmaxf(m[e], vv[e]);
                    } else {
                        const int kj = S.kept[r + k], pos = (int)c0 + r + k;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (vv[e] > m[e] || (vv[e] == m[e] && kj < mk[e])) { m[e] = vv[e]; mk[e] = kj; mp[e] = pos; }
                    }
                }
                if (!open) {
                    *reinterpret_cast<float4 *>(a.features + (size_t)pid * COUT + 4 * q) = make_float4(m[0], m[1], m[2], m[3]);
                    if (ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) mp[e] = (m[e] > 0.0f) ? mp[e] : ~mp[e];   // negative: clamped by the ReLU
                        *reinterpret_cast<int4 *>(a.argpos + (size_t)pid * COUT + 4 * q) = make_int4(mp[0], mp[1], mp[2], mp[3]);
                    }
                } else {   // the pillar continues in the next window of this CTA
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
