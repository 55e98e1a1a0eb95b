// rdp_table.cuh -- the per-pillar table kernel (K6) and the helpers it shares with the train-mode statistics kernel.
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
// Train-mode BatchNorm: pillar_table_stats_kernel (rdp_pfn.cuh, one instantiation per compiled shape) is this kernel plus
// the feature moments -- row x row products thread-local, the pillar-level products as rank-32 updates on the fp64 tensor
// pipe -- and the BatchNorm epilogue in its last CTA.
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

// Row inputs of the folded linear layer from a grouped row [key, x, y, z, f...]: the centre offsets d = xyz - centre (exactly
// f_center of the reference, :215-217), the raw feature columns after xyz, and the distance feature (:230-231) when the
// layout has one.
template <int COLS, bool DIST>
__device__ __forceinline__ void row_inputs(const float *r, float cenx, float ceny, float cenz, float *out) {
    const float x = r[1], y = r[2], z = r[3];
    out[0] = __fsub_rn(x, cenx);
    out[1] = __fsub_rn(y, ceny);
    out[2] = __fsub_rn(z, cenz);
#pragma unroll
    for (int c = 4; c < COLS; ++c) out[c - 1] = r[c];
    if (DIST) out[COLS - 1] = sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
}

template <int RSF, int ROW_STRIDE = 0>   // RSF floats loaded; rows ROW_STRIDE floats apart (0 = unknown: 16-byte loads only)
__device__ __forceinline__ void load_row(const float *src, float *r) {
    if (RSF == 8 && ROW_STRIDE == 8) {   // a 32-byte grouped row is one aligned sector: one 256-bit load (sm_100 LDG.256) instead of two 128-bit ones
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                     : "l"(src));
        return;
    }
#pragma unroll
    for (int c4 = 0; c4 < RSF; c4 += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + c4));
        r[c4] = v.x; r[c4 + 1] = v.y; r[c4 + 2] = v.z; r[c4 + 3] = v.w;
    }
}

// Pillar centre and coords from the merged key (:243-248, :132-138; voxels dynamic_voxel_vfe.py:94-100).
__device__ __forceinline__ void decode_key(const GeomDev &g, int key, float *cenx, float *ceny, float *cenz, int4 *bzyx) {
    const int nz = g.nz > 1 ? g.nz : 1;
    const int sxy = g.nx * g.ny * nz;
    const int b = key / sxy, rem = key - b * sxy;
    const int cx = rem / (g.ny * nz), rem2 = rem - cx * (g.ny * nz);
    const int cy = rem2 / nz, cz = rem2 - cy * nz;
    *cenx = __fadd_rn(__fmul_rn((float)cx, g.vx), g.off_x);   // cx*vx + x_off with separate mul / add roundings (:215-216)
    *ceny = __fadd_rn(__fmul_rn((float)cy, g.vy), g.off_y);
    *cenz = g.nz > 1 ? __fadd_rn(__fmul_rn((float)cz, g.vz), g.off_z) : g.off_z;   // dynamic_voxel_vfe.py:79
    *bzyx = make_int4(b, cz, cy, cx);
}

// One table entry [centre xy | -mean(d) xyz | start | rows | centre z] from the exact fp64 sums of the pillar's centre offsets d
// d = xyz - centre: -mean(d) = centre - mean(xyz), the pillar part of f_cluster (:226-227).
__device__ __forceinline__ void pillar_entry(float cenx, float ceny, float cenz, double sx, double sy, double sz, int start, int rows,
                                             float *entry) {
    float mx, my, mz;
    mean3(sx, sy, sz, rows, &mx, &my, &mz);
    entry[0] = cenx; entry[1] = ceny;
    entry[2] = -mx; entry[3] = -my; entry[4] = -mz;
    entry[5] = __int_as_float(start); entry[6] = __int_as_float(rows); entry[7] = cenz;
}

// fp64 sums of the centre offsets of grouped rows [s, e) by a whole warp (lane-strided, shuffle reduction; exact, so any order gives the same bits)
__device__ __forceinline__ void warp_sum_rows(const float *grows, int rs, int s, int e, float cenx, float ceny, float cenz, double *sx,
                                              double *sy, double *sz) {
    const int lane = threadIdx.x & 31;
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int i = s + lane; i < e; i += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(grows + ((size_t)i + 1) * rs));
        ax += (double)__fsub_rn(v.y, cenx); ay += (double)__fsub_rn(v.z, ceny); az += (double)__fsub_rn(v.w, cenz);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, d);
        ay += __shfl_xor_sync(0xffffffffu, ay, d);
        az += __shfl_xor_sync(0xffffffffu, az, d);
    }
    *sx = ax; *sy = ay; *sz = az;
}

__device__ __forceinline__ void store_coords(int32_t *coords, int coord_cols, size_t p, const int4 &c) {
    if (coord_cols == 3) {
        int32_t *o = coords + p * 3;
        o[0] = c.x; o[1] = c.z; o[2] = c.w;   // [b, y, x]  (:248)
    } else {
        *reinterpret_cast<int4 *>(coords + p * 4) = c;   // [b, 0, y, x] (:138) / [b, z, y, x]
    }
}

__device__ __forceinline__ void pillar_table_body(const TableArgs &t) {
    const int lane = threadIdx.x & 31;
    const int P = t.counters[RDP_CNT_P];
    const int rs = t.rs;
    const int stride = gridDim.x * blockDim.x;
    for (int pb = blockIdx.x * blockDim.x + (threadIdx.x & ~31); pb < P; pb += stride) {
        const int p = pb + lane;
        const bool valid = p < P;
        int s = 0, e = 0;
        if (valid) { s = t.starts[p]; e = t.starts[p + 1]; }
        const bool big = valid && (e - s) > kBigRows;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        float cenx = 0.f, ceny = 0.f, cenz = 0.f;
        int4 c = make_int4(0, 0, 0, 0);
        if (valid) {
            const float *r = t.grows + ((size_t)s + 1) * rs;   // rows are 16-byte aligned: [key, x, y, z] is one 128-bit load
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(r));
            decode_key(t.g, __float_as_int(v0.x), &cenx, &ceny, &cenz, &c);
            if (!big) {
                // the mean is taken over the centre offsets d = xyz - centre (f_center, :215-217): small, exactly summable
                sx = (double)__fsub_rn(v0.y, cenx); sy = (double)__fsub_rn(v0.z, ceny); sz = (double)__fsub_rn(v0.w, cenz);
                r += rs;
                for (int i = s + 1; i < e; ++i, r += rs) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(r));
                    sx += (double)__fsub_rn(v.y, cenx); sy += (double)__fsub_rn(v.z, ceny); sz += (double)__fsub_rn(v.w, cenz);
                }
            }
        }
        unsigned bigmask = __ballot_sync(0xffffffffu, big);
        while (bigmask) {   // whole-warp sum of one long pillar
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const int sb = __shfl_sync(0xffffffffu, s, src), eb = __shfl_sync(0xffffffffu, e, src);
            const float bx = __shfl_sync(0xffffffffu, cenx, src), by = __shfl_sync(0xffffffffu, ceny, src),
                        bz = __shfl_sync(0xffffffffu, cenz, src);
            double ax, ay, az;
            warp_sum_rows(t.grows, rs, sb, eb, bx, by, bz, &ax, &ay, &az);
            if (lane == src) { sx = ax; sy = ay; sz = az; }
        }
        if (valid) {
            float entry[8];
            pillar_entry(cenx, ceny, cenz, sx, sy, sz, s, e - s, entry);
            if (t.coords) store_coords(t.coords, t.coord_cols, (size_t)p, c);
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(t.aux + (size_t)p * 8), "f"(entry[0]), "f"(entry[1]),
                         "f"(entry[2]), "f"(entry[3]), "f"(entry[4]), "f"(entry[5]), "f"(entry[6]), "f"(entry[7])
                         : "memory");
        }
    }
}

}  // namespace rdp
