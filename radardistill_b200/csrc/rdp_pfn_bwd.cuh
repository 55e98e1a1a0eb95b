// rdp_pfn_bwd.cuh -- backward PFN kernel, streaming form (replaces pfn_tile_kernel<BWD> on the default path).
//
// Autograd of PFNLayerV2.forward (dynamic_pillar_vfe.py:35-46) for the parameters: ScatterMax routes each
// (pillar, channel) gradient to the argmax row, ReLU' masks it, and the linear / BatchNorm backward only need
//     dbeta_c = sum_p gy[p][c]            A_c[k] = sum_p gy[p][c] f[argmax(p, c)][k]
// (G_c = w_c . A_c, the rest is closed form: bwd_finalize_kernel).  So the kernel is a stream over pillars:
//   warp = a contiguous range of pillars, lane = channel (+32 per extra channel block);
//   per pillar: coalesced 128-byte rows of grad / argpos (its sign carries the ReLU mask), the pillar's table entry (warp uniform),
//   each lane's winner row straight from the grouped row array (a pillar has ~2 rows: the 32 lanes hit 1-2 lines),
//   features re-decorated in registers, 14 FMAs into fp32 accumulators that are folded into fp64 every 32 pillars.
// No shared-memory staging, no block barriers in the stream: the tile form spent its time waiting for the
// per-tile (grad, output, argpos) copies (ncu: 254 us for 50 M instructions); here the loads of the next pillars are
// in flight while the current ones are folded in, and occupancy hides the rest.
#pragma once

#include "rdp_pfn.cuh"

namespace rdp {

constexpr int kBwdThreads = 128;
#ifndef RDP_BWD_UNROLL
#define RDP_BWD_UNROLL 4
#endif
#ifndef RDP_BWD_CTAS_PER_SM
#define RDP_BWD_CTAS_PER_SM 4
#endif
constexpr int kBwdGrid = 148 * RDP_BWD_CTAS_PER_SM < kPfnGridCap ? 148 * RDP_BWD_CTAS_PER_SM : kPfnGridCap;  // <= partial-sum slots

template <class Cfg>
__global__ void __launch_bounds__(kBwdThreads, RDP_BWD_CTAS_PER_SM) pfn_bwd_stream_kernel(const __grid_constant__ PfnArgs a) {
    constexpr int COUT = Cfg::COUT, CS = Cfg::CS, RS = Cfg::RS, CPL = COUT / 32, PER = Cfg::BWD_PER;
    constexpr int NW = kBwdThreads / 32, U = RDP_BWD_UNROLL, FOLD = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // fp64 totals of every (warp, channel): [dbeta | 0 | A(CS)], touched only when the fp32 accumulators are folded in
    double *dacc = reinterpret_cast<double *>(smem_raw);   // NW * COUT * PER doubles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.counters[RDP_CNT_P];
    // Grid-stride batches of U pillars: at any moment the whole grid reads one contiguous window of every array (the
    // DRAM pages it opens are used completely), instead of one private stream per warp.  The assignment is fixed, so
    // the sums stay deterministic.
    const int nwarps = (int)gridDim.x * NW, gw = (int)blockIdx.x * NW + warp;
    const int p_begin = gw * U, p_end = P, p_step = nwarps * U;

    float tB[CPL], tA[CPL][CS];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
        tB[cc] = 0.0f;
        double *dst = dacc + ((size_t)warp * COUT + lane + 32 * cc) * PER;
#pragma unroll
        for (int k = 0; k < PER; ++k) dst[k] = 0.0;
#pragma unroll
        for (int k = 0; k < CS; ++k) tA[cc][k] = 0.0f;
    }
    auto fold = [&]() {
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            double *dst = dacc + ((size_t)warp * COUT + lane + 32 * cc) * PER;
            dst[0] += (double)tB[cc]; tB[cc] = 0.0f;
#pragma unroll
            for (int k = 0; k < CS; ++k) { dst[2 + k] += (double)tA[cc][k]; tA[cc][k] = 0.0f; }
        }
    };

    // stage 1 of a batch of U pillars: the coalesced (grad, argpos) rows
    struct S1 { float g[U][CPL]; int pos[U][CPL]; };
    auto load1 = [&](int p0, S1 &o) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = min(p0 + u, p_end - 1);
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const size_t off = (size_t)p * COUT + lane + 32 * cc;
                o.g[u][cc] = __ldg(a.grad + off);
                o.pos[u][cc] = __ldg(a.argpos + off);
            }
        }
    };

    S1 nxt;
    if (p_begin < p_end) load1(p_begin, nxt);
    int since_fold = 0;
    for (int p0 = p_begin; p0 < p_end; p0 += p_step) {
        // ---- stage 2 of this batch: winner rows (address = argpos, requested one iteration ago) + table entries ...
        float gy[U][CPL];
        float4 rowv[U][CPL][RS / 4];
        float4 m4[U];
        float cy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool live = p0 + u < p_end;
            const int p = min(p0 + u, p_end - 1);
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int ap = nxt.pos[u][cc];   // negative: the forward marked the pillar as ReLU-clamped (:38)
                gy[u][cc] = (live && ap >= 0) ? nxt.g[u][cc] : 0.0f;
                const float4 *src = reinterpret_cast<const float4 *>(a.grows + ((size_t)(ap >= 0 ? ap : ~ap) + 1) * RS);
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) rowv[u][cc][q] = __ldg(src + q);
            }
            const float *ax = a.aux + (size_t)p * 8;   // warp uniform: [mean x y z | centre x] [centre y | ...]
            m4[u] = __ldg(reinterpret_cast<const float4 *>(ax));
            cy[u] = __ldg(ax + 4);
        }
        // ---- ... and stage 1 of the next batch goes out before anything of this one is waited for
        if (p0 + p_step < p_end) load1(p0 + p_step, nxt);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float mean[3] = {m4[u].x, m4[u].y, m4[u].z};
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                float row[RS], f[Cfg::FW];
#pragma unroll
                for (int q = 0; q < RS / 4; ++q) {
                    row[4 * q] = rowv[u][cc][q].x; row[4 * q + 1] = rowv[u][cc][q].y;
                    row[4 * q + 2] = rowv[u][cc][q].z; row[4 * q + 3] = rowv[u][cc][q].w;
                }
                decorate<Cfg>(row, m4[u].w, cy[u], mean, a, f);
                const float g = gy[u][cc];
                tB[cc] += g;
#pragma unroll
                for (int k = 0; k < CS; ++k) tA[cc][k] = fmaf(g, f[k], tA[cc][k]);
            }
        }
        since_fold += U;
        if (since_fold >= FOLD) { fold(); since_fold = 0; }
    }
    fold();

    // ---- per-CTA partial sums, layout per channel [dbeta | 0 | A(CS)]  (same as the tile form; reduce_partials_kernel
    //      adds the CTAs in a fixed order)
    __syncthreads();
    double *out = a.partials + (size_t)blockIdx.x * Cfg::BWD_DOUBLES;
    for (int e = tid; e < COUT * PER; e += kBwdThreads) {
        double sacc = 0.0;
        for (int w = 0; w < NW; ++w) sacc += dacc[(size_t)w * COUT * PER + e];
        out[e] = sacc;
    }
}

}  // namespace rdp
