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
//      row (32 B) + its pillar's table entry from global memory (coalesced / L1) -- fetched into registers one chunk
//      ahead, so the loads fly under the previous chunk's arithmetic -- decorate in registers,
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
#ifndef RDP_ROWS_R
#define RDP_ROWS_R 2
#endif
constexpr int kRowsPerThread = RDP_ROWS_R;
constexpr int kRowsChunk = kRowsThreads * kRowsPerThread;
#ifndef RDP_ROWS_GRID_PER_SM
#define RDP_ROWS_GRID_PER_SM 5
#endif
constexpr int kRowsGridCap = 148 * RDP_ROWS_GRID_PER_SM;
#ifndef RDP_ROWS_REGPIPE
#define RDP_ROWS_REGPIPE 0   // 1: also hold the next chunk's rows / table entries in registers (costs ~36 registers)
#endif

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

__device__ __forceinline__ void st_global_v8(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

template <class Cfg, bool ARG>
__global__ void __launch_bounds__(kRowsThreads, RDP_ROWS_GRID_PER_SM) pfn_rows_kernel(const __grid_constant__ PfnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = RowsSmem<Cfg, ARG>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, COLS = Cfg::COLS, RS = Cfg::RS, WS = Smem::WS, XS = Smem::XS;
    constexpr int NT = kRowsThreads, R = kRowsPerThread, CHUNK = kRowsChunk, WIN = kPfnWin, INF = 0x7fffffff;
    constexpr int QUADS = COUT / 4, GROUPS = NT / QUADS, RPG = CHUNK / GROUPS;  // phase B: rows per (row group)
    static_assert(RPG == 8 || RPG == 16 || RPG == 32, "row groups must align with the 32-bit flag words");
    static_assert(COUT % 8 == 0 && COLS <= RS - 2, "layout");
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

    // ---- register pipeline: rows of chunk i+1 are requested before chunk i's arithmetic, their table entries (whose
    //      address needs the row's pillar id) before chunk i's phase B
    struct RowIn { float4 v[RS / 4]; int gprev, gnext; };
    struct AuxIn { float4 m4, c4; };   // [mean x y z | centre x] [centre y | first row | rows | 0]
    auto fetch_rows = [&](long long c0, int nrow, RowIn (&o)[R]) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = tid + i * NT;
            if (r < nrow) {
                const long long g = c0 + r;
                const float4 *src = reinterpret_cast<const float4 *>(a.grows + (size_t)(g + 1) * RS);
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) o[i].v[q] = src[q];
                o[i].gprev = __float_as_int(a.grows[(size_t)g * RS + RS - 1]);   // the row in front of row 0 is a sentinel (pillar -1)
                o[i].gnext = (g + 1 < N) ? __float_as_int(a.grows[(size_t)(g + 2) * RS + RS - 1]) : -1;
            }
        }
    };
    auto fetch_aux = [&](int nrow, const RowIn (&in)[R], AuxIn (&o)[R]) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (tid + i * NT < nrow) {
                const float4 *ax = reinterpret_cast<const float4 *>(a.aux + (size_t)__float_as_int(in[i].v[RS / 4 - 1].w) * 8);
                o[i].m4 = ax[0];
                o[i].c4 = ax[1];
            }
        }
    };
    RowIn nx[R];
    AuxIn na[R];
    long long c0 = row_begin;
    int nrow = (int)min((long long)CHUNK, row_end - c0);
    if (RDP_ROWS_REGPIPE && c0 < row_end) {
        fetch_rows(c0, nrow, nx);
        fetch_aux(nrow, nx, na);
    }
    __syncthreads();

    int par = 0;
    while (c0 < row_end) {
        RowIn cur[R];
        AuxIn ca[R];
        const long long c1 = c0 + CHUNK;
        const int nrow1 = (c1 < row_end) ? (int)min((long long)CHUNK, row_end - c1) : 0;
        if (RDP_ROWS_REGPIPE) {
#pragma unroll
            for (int i = 0; i < R; ++i) { cur[i] = nx[i]; ca[i] = na[i]; }
            if (nrow1) fetch_rows(c1, nrow1, nx);
        } else {
            fetch_rows(c0, nrow, cur);
            fetch_aux(nrow, cur, ca);
        }
        // =========================================================================== A: thread = row
        float f[R][Cfg::FW];
        bool valid[R], single[R];
        int gidv[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = tid + i * NT;
            valid[i] = r < nrow;
            bool head = false, last = false;
            gidv[i] = 0;
            if (valid[i]) {
                float row[RS];
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) {
                    row[4 * q] = cur[i].v[q].x; row[4 * q + 1] = cur[i].v[q].y; row[4 * q + 2] = cur[i].v[q].z; row[4 * q + 3] = cur[i].v[q].w;
                }
                const int gid = __float_as_int(row[RS - 1]);
                gidv[i] = gid;
                head = gid != cur[i].gprev;
                last = gid != cur[i].gnext;
                const float mean[3] = {ca[i].m4.x, ca[i].m4.y, ca[i].m4.z};
                decorate<Cfg>(row, ca[i].m4.w, ca[i].c4.x, mean, a, f[i]);
                S.gid[r] = gid;
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
        if (RDP_ROWS_REGPIPE && nrow1) fetch_aux(nrow1, nx, na);
        __syncthreads();

        // =========================================================================== B: thread = (row group, quad)
        {
            const int g = tid / QUADS, q = tid % QUADS;
            const int gr0 = g * RPG;
            uint32_t hw = S.heads[gr0 >> 5] >> (gr0 & 31), lw = S.lasts[gr0 >> 5] >> (gr0 & 31);
            constexpr uint32_t gmask = RPG < 32 ? ((1u << (RPG & 31)) - 1u) : 0xffffffffu;
            hw &= gmask; lw &= gmask;
            uint32_t own = hw & ~lw;                                 // heads of multi-row pillars that start in my rows
            bool carry_in = (g == 0) && !(S.heads[0] & 1u);           // the chunk starts inside a pillar: continue it from the carry
            // rows from r through the first last-row flag at or after r; -1 if the pillar is still open at the end of the chunk
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
                const int pid = S.gid[r];
                int len = seg_len(r);
                const bool open = len < 0;
                if (open) len = nrow - r;
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
                } else {   // the pillar continues in the next chunk of this CTA
                    *reinterpret_cast<float4 *>(&S.carry_v[par ^ 1][4 * q]) = make_float4(m[0], m[1], m[2], m[3]);
                    if (ARG) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { S.carry_k[par ^ 1][4 * q + e] = mk[e]; S.carry_p[par ^ 1][4 * q + e] = mp[e]; }
                    }
                }
            }
        }
        __syncthreads();
        c0 = c1;
        nrow = nrow1;
        par ^= 1;
    }
}

}  // namespace rdp
