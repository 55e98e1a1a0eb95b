// rdp_table.cuh -- the per-pillar table kernel (K6) and, fused into it for train-mode BatchNorm, the feature moments.
//
// Replaces scatter_mean (/root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:226-227, :105-106), the
// pillar centre of f_center (:214-217, :108-111; voxel form dynamic_voxel_vfe.py:75-79) and the coordinate decode
// (:243-248, :132-138; dynamic_voxel_vfe.py:94-100).
//
// Thread = pillar reads its (contiguous) grouped rows once and writes the table entry
//     [centre x, centre y, (centre - mean) x, y, z, first grouped row, rows, centre z]
// and the pillar's coords.  Mean: fp64 sum of the fp32 coordinates (exact => order independent), correctly rounded
// quotient, one rounding to fp32.  A pillar with more than kBigRows rows is summed by its whole warp (lane-strided rows,
// shuffle reduction -- still exact), so one 3 500-point cell does not serialise 3 500 dependent loads on one thread.
//
// The train-mode feature moments are a separate row-parallel pass (pfn_moments_kernel, rdp_pfn.cuh): fused into this
// thread = pillar kernel their ~70 accumulators per thread cut the occupancy of an already latency-bound kernel to 11 warps
// per SM (178 registers; measured at compile time), which costs more than the second read of the rows.
#pragma once

#include "rdp_common.cuh"

namespace rdp {

constexpr int kBigRows = 32;   // rows above which a pillar is handled by its whole warp

struct GeomDev {
    float lo_x, lo_y, lo_z, vx, vy, vz;
    int nx, ny, nz, batch, cols;   // nz <= 1: pillars (z is neither quantised nor masked); nz > 1: voxels
    float off_x, off_y, off_z;
};

struct TableArgs {
    const float *grows;
    const int32_t *starts;
    int32_t *counters;
    float *aux;
    int32_t *coords;
    int rs, coord_cols;
    GeomDev g;
};

template <int COLS, bool DIST>
struct RowBasis {
    static constexpr int KIN = COLS - 1 + (DIST ? 1 : 0);   // dx, dy, dz, raw features 4.., (dist)
    static constexpr int G = KIN + 5;                       // + centre xy, centre - mean xyz
    static constexpr int NACC = G + G * G;                  // S1 | S2 (full, row-major)
    static constexpr int RS = (COLS + 2 + 3) / 4 * 4;
};

// Row inputs of the folded linear layer: centre offsets (exactly f_center of the reference, :215-217), the raw feature
// columns after xyz, and the distance feature (:230-231) when the layout has one.
template <int COLS, bool DIST>
__device__ __forceinline__ void row_inputs(const float *r /* [b,x,y,z,f...] */, float cenx, float ceny, float cenz, float *out) {
    const float x = r[1], y = r[2], z = r[3];
    out[0] = __fsub_rn(x, cenx);
    out[1] = __fsub_rn(y, ceny);
    out[2] = __fsub_rn(z, cenz);
#pragma unroll
    for (int c = 4; c < COLS; ++c) out[c - 1] = r[c];
    if (DIST) out[COLS - 1] = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
}

template <int RSF>
__device__ __forceinline__ void load_row(const float *src, float *r) {
#pragma unroll
    for (int c4 = 0; c4 < RSF; c4 += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + c4));
        r[c4] = v.x; r[c4 + 1] = v.y; r[c4 + 2] = v.z; r[c4 + 3] = v.w;
    }
}

__device__ __forceinline__ void pillar_table_body(const TableArgs &t) {
    const int lane = threadIdx.x & 31;
    const int P = t.counters[RDP_CNT_P];
    const int rs = t.rs;
    const GeomDev &g = t.g;
    const bool vox = g.nz > 1;
    const int stride = gridDim.x * blockDim.x;
    for (int pb = blockIdx.x * blockDim.x + (threadIdx.x & ~31); pb < P; pb += stride) {
        const int p = pb + lane;
        const bool valid = p < P;
        int s = 0, e = 0;
        if (valid) { s = t.starts[p]; e = t.starts[p + 1]; }
        const bool big = valid && (e - s) > kBigRows;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        float x0 = 0.0f, y0 = 0.0f, z0 = 0.0f, b0 = 0.0f;
        if (valid) {
            const float *r = t.grows + ((size_t)s + 1) * rs;   // rows are 16-byte aligned: [b, x, y, z] is one 128-bit load
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(r));
            b0 = v0.x; x0 = v0.y; y0 = v0.z; z0 = v0.w;
            if (!big) {
                sx = (double)x0; sy = (double)y0; sz = (double)z0;
                r += rs;
                for (int i = s + 1; i < e; ++i, r += rs) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(r));
                    sx += (double)v.y; sy += (double)v.z; sz += (double)v.w;
                }
            }
        }
        unsigned bigmask = __ballot_sync(0xffffffffu, big);
        while (bigmask) {   // whole-warp sum of one long pillar (exact in fp64: any order gives the same bits)
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const int sb = __shfl_sync(0xffffffffu, s, src), eb = __shfl_sync(0xffffffffu, e, src);
            double ax = 0.0, ay = 0.0, az = 0.0;
            for (int i = sb + lane; i < eb; i += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(t.grows + ((size_t)i + 1) * rs));
                ax += (double)v.y; ay += (double)v.z; az += (double)v.w;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, d);
                ay += __shfl_xor_sync(0xffffffffu, ay, d);
                az += __shfl_xor_sync(0xffffffffu, az, d);
            }
            if (lane == src) { sx = ax; sy = ay; sz = az; }
        }
        float cenx = 0.0f, ceny = 0.0f, cenz = g.off_z, ndx = 0.0f, ndy = 0.0f, ndz = 0.0f;
        if (valid) {
            float mx, my, mz;
            mean3(sx, sy, sz, e - s, &mx, &my, &mz);
            // centre of the cell: cx*vx + x_off with separate mul / add roundings (:215-216); the quantisation repeats
            // quantize_mark_kernel's IEEE ops on a row of the pillar, so cx / cy (/ cz) equal the emitted coords
            const float qx = floorf(__fdiv_rn(__fsub_rn(x0, g.lo_x), g.vx)), qy = floorf(__fdiv_rn(__fsub_rn(y0, g.lo_y), g.vy));
            cenx = __fadd_rn(__fmul_rn((float)(int)qx, g.vx), g.off_x);
            ceny = __fadd_rn(__fmul_rn((float)(int)qy, g.vy), g.off_y);
            int cz = 0;
            if (vox) {
                const float qz = floorf(__fdiv_rn(__fsub_rn(z0, g.lo_z), g.vz));
                cz = (int)qz;
                cenz = __fadd_rn(__fmul_rn((float)cz, g.vz), g.off_z);   // dynamic_voxel_vfe.py:79
            }
            ndx = __fsub_rn(cenx, mx); ndy = __fsub_rn(ceny, my); ndz = __fsub_rn(cenz, mz);
            const int bi = __float2int_rz(b0);
            if (!t.coords) {
                // table only (a caller that already has the coords)
            } else if (t.coord_cols == 3) {
                int32_t *o = t.coords + (size_t)p * 3;
                o[0] = bi; o[1] = (int)qy; o[2] = (int)qx;   // [b, y, x]  (:248)
            } else {
                *reinterpret_cast<int4 *>(t.coords + (size_t)p * 4) = make_int4(bi, cz, (int)qy, (int)qx);  // (:138) / [b,z,y,x]
            }
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(t.aux + (size_t)p * 8), "f"(cenx), "f"(ceny), "f"(ndx),
                         "f"(ndy), "f"(ndz), "f"(__int_as_float(s)), "f"(__int_as_float(e - s)), "f"(cenz)
                         : "memory");
        }
    }
}

}  // namespace rdp
