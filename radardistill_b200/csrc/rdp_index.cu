// rdp_index.cu -- points -> pillars: quantise, range mask, merged key, sorted unique, inverse, counts,
// coords, and the pillar-grouped point order the PFN kernels consume.
//
// Replaces /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:201-212,243-248
// (identical code at :93-103,:132-138, :262-273, :322-333).
//
// The merged key  b*nx*ny + cx*ny + cy  (:208-210) is a small dense integer, so the "hash" is a
// direct-addressed occupancy bitmap (collision-free open addressing with the identity hash) and
// torch.unique's ascending order falls out of a popcount prefix scan over the bitmap -- a counting
// sort with one-bit counters; no comparison sort, no ordering of floats, fully deterministic.
//
//   K1 quantize_mark   points (TMA bulk tile -> smem) -> key[i], bitmap |= bit(key); counts[] = 0   [HBM: read rows]
//   K2 bitmap_rank     popcount scan of the bitmap -> {word, rank of word} pairs, P; publishes (N, P) to the host
//   K3 rank_count      key[i] -> rank[i] = wordrank.rank + popc(below) ; inverse[j] ; slot[i] = counts[rank]++
//   K4 count_scan      exclusive scan of counts -> pillar start offsets (starts[P] = N), tile_first
//   K5 group_rows      grouped_rows[starts[rank] + slot] = [merged key | xyz | features | original row id | rank]   (no atomics; the order
//                      inside a pillar is the arrival order of K3's atomics -- every consumer is order independent).
//                      The PFN kernels then stream contiguous, pillar-aligned row tiles with TMA.
//   K6 pillar_table    thread = pillar: key decode (centre, coords), fp64 mean of its grouped rows' centre offsets, first row,
//                      row count (rdp_table.cuh).  A train-mode fused forward skips this launch: its statistics kernel
//                      (pillar_table_stats_kernel, rdp_pfn.cuh) writes the same table while it accumulates the moments.
// K2..K6 are launched with programmatic dependent launch (rdp_common.cuh: pdl_wait is their first statement).
// nz > 1 (DynamicVoxelVFE / DynamicMeanVFE, dynamic_voxel_vfe.py:57-71, dynamic_mean_vfe.py:52-60): z is quantised and
// masked too and the key is ((b*nx + cx)*ny + cy)*nz + cz.
#include "rdp_index_host.h"

namespace rdp {


// Frame of input row i when the rows carry no batch column: offsets[b] <= i < offsets[b + 1] (offsets has batch + 1 entries).
__device__ __forceinline__ int frame_of(const int32_t *__restrict__ offsets, int batch, long long i) {
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ----------------------------------------------------------------------------- K1
__global__ void __launch_bounds__(kIndexThreads)
quantize_mark_kernel(const float *__restrict__ pts, long long n0, GeomDev g, uint32_t *__restrict__ bitmap,
                     int32_t *__restrict__ keys, int32_t *__restrict__ tile_keep, int32_t *__restrict__ counters,
                     const int32_t *__restrict__ offsets, int32_t *__restrict__ counts) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_keep;

    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kIndexTileRows;
    const int rows = (int)min((long long)kIndexTileRows, n0 - row0);
    // offsets != null: the rows are (x, y, z, features...) without the batch column and frame b owns rows
    // [offsets[b], offsets[b + 1])  (what the dataset hands over before collate_batch pads the index in, dataset_distill.py:237-244)
    const int in_cols = offsets ? g.cols - 1 : g.cols, xc = offsets ? 0 : 1;
    const int floats = rows * in_cols;
    const uint32_t bulk_bytes = (uint32_t)(floats * 4) & ~15u;
    const float *src = pts + row0 * in_cols;
    __shared__ int s_b0;

    if (tid == 0) {
        s_b0 = offsets ? frame_of(offsets, g.batch, row0) : 0;
        s_keep = 0;
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && bulk_bytes) {
        mbar_expect_tx(&bar, bulk_bytes);
        tma_bulk_g2s(tile, src, bulk_bytes, &bar);  // one 1-D TMA copy of the whole row tile
    }
    for (int f = (int)(bulk_bytes >> 2) + tid; f < floats; f += kIndexThreads) tile[f] = src[f];  // <16 B tail
    // counts[0, P) must be zero for K3's atomics and P <= n0: every tile clears its own stretch while its rows are in flight
    if (row0 + tid * 4 + 3 < n0) {
        *reinterpret_cast<int4 *>(counts + row0 + tid * 4) = make_int4(0, 0, 0, 0);
    } else {
        for (int q = 0; q < 4; ++q) if (row0 + tid * 4 + q < n0) counts[row0 + tid * 4 + q] = 0;
    }
    if (bulk_bytes) mbar_wait(&bar, 0);
    __syncthreads();

    const bool vox = g.nz > 1;
    const int sxy = g.nx * g.ny * (vox ? g.nz : 1);
    int kept = 0;
    bool bad_batch = false;
#pragma unroll
    for (int k = 0; k < kIndexTileRows / kIndexThreads; ++k) {
        const int r = k * kIndexThreads + tid;
        if (r < rows) {
            const float *p = tile + r * in_cols;
            // IEEE fp32 subtract + divide + floor, as torch.floor((xy - lo) / vsz)  (:201-202)
            const float qx = floorf(__fdiv_rn(__fsub_rn(p[xc], g.lo_x), g.vx));
            const float qy = floorf(__fdiv_rn(__fsub_rn(p[xc + 1], g.lo_y), g.vy));
            bool ok = (qx >= 0.0f) && (qx < (float)g.nx) && (qy >= 0.0f) && (qy < (float)g.ny);  // NaN/inf fail
            float qz = 0.0f;
            if (vox) {   // dynamic_voxel_vfe.py:57-58: z is quantised and masked as well
                qz = floorf(__fdiv_rn(__fsub_rn(p[xc + 2], g.lo_z), g.vz));
                ok = ok && (qz >= 0.0f) && (qz < (float)g.nz);
            }
            int b;
            if (offsets) {
                b = s_b0;
                while (b + 1 < g.batch && row0 + r >= (long long)offsets[b + 1]) ++b;
            } else {
                b = __float2int_rz(p[0]);  // .int() truncates
            }
            if (ok && (b < 0 || b >= g.batch)) { ok = false; bad_batch = true; }
            int key = -1;
            if (ok) {
                key = vox ? b * sxy + ((int)qx * g.ny + (int)qy) * g.nz + (int)qz   // dynamic_voxel_vfe.py:63-66
                          : b * sxy + (int)qx * g.ny + (int)qy;                     // (:208-210)
                atomicOr(bitmap + (key >> 5), 1u << (key & 31));
                ++kept;
            }
            keys[row0 + r] = key;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
    if ((tid & 31) == 0 && kept) atomicAdd(&s_keep, kept);
    if (bad_batch) atomicOr(counters + RDP_CNT_ERRFLAGS, 1);
    __syncthreads();
    if (tid == 0) {
        tile_keep[blockIdx.x] = s_keep;
        if (s_keep) atomicAdd(counters + RDP_CNT_N, s_keep);
    }
}

// ----------------------------------------------------------------------------- K2
// Chunked two-sweep scan over the bitmap words: sweep 1 = popcount of my chunk, hand-off, sweep 2 =
// the exclusive rank of every word.
__global__ void __launch_bounds__(kScanThreads)
bitmap_rank_kernel(const uint32_t *__restrict__ bitmap, long long words,
                   uint64_t *__restrict__ state, uint2 *__restrict__ wordrank, int32_t *__restrict__ counters,
                   volatile int32_t *host_mapped) {
    pdl_wait();
    __shared__ int s_scan[9];
    __shared__ uint32_t s_u32;
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = atomicAdd(counters + kCntTicketA, 1);
    __syncthreads();
    const int ticket = s_ticket;
    const long long per = (((words + gridDim.x - 1) / gridDim.x) + 1023) / 1024 * 1024;
    const long long w0 = min(words, per * ticket), w1 = min(words, w0 + per);

    // sweep 1
    uint32_t local = 0;
    for (long long w = w0 + tid * 4; w < w1; w += kScanThreads * 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(bitmap + w);  // words padded to a multiple of 4
        local += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    int tot;
    block_excl_scan_256((int)local, s_scan, &tot);
    uint32_t base = chunk_exclusive_prefix(state, ticket, (uint32_t)tot, &s_u32);
    if (ticket == (int)gridDim.x - 1 && tid == 0) {
        // N (K1) and P are final here: publish them to the host (zero-copy store into pinned memory) before sweep 2,
        // so the host learns the output sizes ~50 us into the forward
        const int P = (int)(base + (uint32_t)tot);
        counters[RDP_CNT_P] = P;
        if (host_mapped) {
            for (int i = 0; i < RDP_NUM_COUNTERS; ++i) host_mapped[i] = (i == RDP_CNT_P) ? P : counters[i];
            __threadfence_system();
        }
    }

    // sweep 2
    for (long long wt = w0; wt < w1; wt += kScanThreads * 4) {
        const long long w = wt + tid * 4;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (w < w1) v = *reinterpret_cast<const uint4 *>(bitmap + w);
        const int c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        int tile_total;
        uint32_t rank = base + (uint32_t)block_excl_scan_256(c, s_scan, &tile_total);
        base += (uint32_t)tile_total;
        if (w < w1) {
            // {word, rank of the word's first pillar} pairs: K3 needs both per point -- one 8-byte gather instead of two
            const uint32_t r1 = rank + __popc(v.x), r2 = r1 + __popc(v.y), r3 = r2 + __popc(v.z);
            uint4 *dst = reinterpret_cast<uint4 *>(wordrank + w);
            dst[0] = make_uint4(v.x, rank, v.y, r1);
            dst[1] = make_uint4(v.z, r2, v.w, r3);
        }
    }
}

// ----------------------------------------------------------------------------- K3
__global__ void __launch_bounds__(kIndexThreads)
rank_count_kernel(const int32_t *__restrict__ keys, int32_t *__restrict__ ranks, long long n0, const uint2 *__restrict__ wordrank,
                  const int32_t *__restrict__ tile_keep, int32_t *__restrict__ inverse, int32_t *__restrict__ counts,
                  const int32_t *__restrict__ counters, int32_t *__restrict__ orig2kept, int32_t *__restrict__ kept2orig,
                  int32_t *__restrict__ slots) {
    pdl_wait();
    __shared__ int s_scan[9];
    __shared__ long long s_base;
    const int tid = threadIdx.x;
    // A CTA covers kRankSub index tiles (8 rows per thread): the kernel is bound by two dependent L2 round trips per row (the
    // {word, rank} gather, then the slot atomic), so the rows in flight per SM set its speed.  All gathers are issued first,
    // then all atomics; the stores follow.  Nothing below synchronises the CTA unless the range mask dropped a row.
    const long long cta0 = (long long)blockIdx.x * kRankSub * kIndexTileRows;
    int k[kRankSub][4], r[kRankSub][4], sl[kRankSub][4];
#pragma unroll
    for (int s = 0; s < kRankSub; ++s) {
        const long long i0 = cta0 + (long long)s * kIndexTileRows + tid * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) k[s][q] = -1;
        if (i0 + 3 < n0) {
            const int4 v = *reinterpret_cast<const int4 *>(keys + i0);
            k[s][0] = v.x; k[s][1] = v.y; k[s][2] = v.z; k[s][3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (i0 + q < n0) k[s][q] = keys[i0 + q];
        }
    }
#pragma unroll
    for (int s = 0; s < kRankSub; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            r[s][q] = -1;
            if (k[s][q] >= 0) {
                const uint2 wr = __ldg(wordrank + (k[s][q] >> 5));
                r[s][q] = (int)(wr.y + __popc(wr.x & ((1u << (k[s][q] & 31)) - 1u)));
            }
        }
    // the value the count had before this row arrived is the row's slot inside its pillar: K5 then needs no atomics
#pragma unroll
    for (int s = 0; s < kRankSub; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sl[s][q] = 0;
            if (r[s][q] >= 0) sl[s][q] = atomicAdd(counts + r[s][q], 1);
        }
    const bool none_dropped = (long long)counters[RDP_CNT_N] == n0;
#pragma unroll
    for (int s = 0; s < kRankSub; ++s) {
        const long long i0 = cta0 + (long long)s * kIndexTileRows + tid * 4;
        if (i0 + 3 < n0) {
            *reinterpret_cast<int4 *>(ranks + i0) = make_int4(r[s][0], r[s][1], r[s][2], r[s][3]);
            *reinterpret_cast<int4 *>(slots + i0) = make_int4(sl[s][0], sl[s][1], sl[s][2], sl[s][3]);
            if (none_dropped) *reinterpret_cast<int4 *>(inverse + i0) = make_int4(r[s][0], r[s][1], r[s][2], r[s][3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < n0) {
                    ranks[i0 + q] = r[s][q];
                    slots[i0 + q] = sl[s][q];
                    if (none_dropped) inverse[i0 + q] = r[s][q];
                }
        }
    }
    if (none_dropped) return;
    // position among the KEPT points (stable compaction of :204-206): kept rows of all earlier index tiles, then a block
    // scan per tile of this CTA
    long long part = 0;
    for (int t = tid; t < (int)blockIdx.x * kRankSub; t += kIndexThreads) part += tile_keep[t];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (tid == 0) s_base = 0;
    __syncthreads();
    if ((tid & 31) == 0 && part) atomicAdd(reinterpret_cast<unsigned long long *>(&s_base), (unsigned long long)part);
    __syncthreads();
    long long base = s_base;
#pragma unroll
    for (int s = 0; s < kRankSub; ++s) {
        const long long i0 = cta0 + (long long)s * kIndexTileRows + tid * 4;
        int c = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) c += r[s][q] >= 0 ? 1 : 0;
        int tot;
        const int excl = block_excl_scan_256(c, s_scan, &tot);
        long long j = base + excl;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (r[s][q] >= 0) {
                orig2kept[i0 + q] = (int)j;
                kept2orig[j] = (int)(i0 + q);
                inverse[j++] = r[s][q];
            }
        base += tot;
    }
}

// ----------------------------------------------------------------------------- K4
__global__ void __launch_bounds__(kScanThreads)
count_scan_kernel(const int32_t *__restrict__ counts, uint64_t *__restrict__ state, int32_t *__restrict__ starts,
                  int32_t *__restrict__ tile_first, int32_t *__restrict__ counters) {
    pdl_wait();
    __shared__ int s_scan[9];
    __shared__ uint32_t s_u32;
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = atomicAdd(counters + kCntTicketB, 1);
    __syncthreads();
    const int ticket = s_ticket;
    const long long P = counters[RDP_CNT_P];
    const long long per = (((P + gridDim.x - 1) / gridDim.x) + 1023) / 1024 * 1024;
    const long long p0 = min(P, per * ticket), p1 = min(P, p0 + per);

    int local = 0;
    for (long long p = p0 + tid; p < p1; p += kScanThreads) local += counts[p];
    int tot;
    block_excl_scan_256(local, s_scan, &tot);
    uint32_t base = chunk_exclusive_prefix(state, ticket, (uint32_t)tot, &s_u32);

    for (long long pt = p0; pt < p1; pt += kScanThreads * 4) {
        const long long p = pt + tid * 4;
        int c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q] = (p + q < p1) ? counts[p + q] : 0;
        int tile_total;
        uint32_t s = base + (uint32_t)block_excl_scan_256(c[0] + c[1] + c[2] + c[3], s_scan, &tile_total);
        base += (uint32_t)tile_total;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (p + q < p1) {
                starts[p + q] = (int)s;  // exclusive start of the pillar in the grouped order
                const uint32_t e = s + (uint32_t)c[q];
                // PFN tile t owns the pillars that start in grouped rows [128 t, 128 t + 128): pillar p+q+1 is the first
                // pillar starting at or after every tile boundary in (s, e]
                for (uint32_t t = s / kPfnWin + 1; t * (uint32_t)kPfnWin <= e; ++t) tile_first[t] = (int)(p + q + 1);
                if (p + q == 0) tile_first[0] = 0;
                if (p + q == P - 1) {
                    starts[P] = (int)e;   // == N
                    if (e % kPfnWin != 0) tile_first[e / kPfnWin + 1] = (int)P;
                }
                s = e;
            }
        }
    }
}

// ----------------------------------------------------------------------------- K5
// Grouped row layout: RS = grouped_row_floats(cols) floats = [row | pad | original row id | pillar id], written with
// 16-byte stores only (the scatter is bound by store requests, not bytes).
template <bool FRAMES>   // FRAMES: rows without the batch column + frame offsets (compiled apart: the padded-row path stays as it was)
__global__ void __launch_bounds__(kIndexThreads)
group_rows_kernel(const float *__restrict__ pts, const int32_t *__restrict__ keys, const int32_t *__restrict__ ranks,
                  const int32_t *__restrict__ slots, long long n0, int cols, const int32_t *__restrict__ starts, float *__restrict__ grows) {
    pdl_wait();
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int rs = grouped_row_floats(cols);
    const long long row0 = (long long)blockIdx.x * kIndexTileRows;
    const int rows = (int)min((long long)kIndexTileRows, n0 - row0);
    const int in_cols = FRAMES ? cols - 1 : cols;
    const int floats = rows * in_cols;
    const uint32_t bulk_bytes = (uint32_t)(floats * 4) & ~15u;
    const float *src = pts + row0 * in_cols;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && bulk_bytes) {
        mbar_expect_tx(&bar, bulk_bytes);
        tma_bulk_g2s(tile, src, bulk_bytes, &bar);
    }
    for (int f = (int)(bulk_bytes >> 2) + tid; f < floats; f += kIndexThreads) tile[f] = src[f];
    if (blockIdx.x == 0 && tid == 0) grows[rs - 1] = __int_as_float(-1);  // sentinel row: "no pillar" before position 0
    // grouped positions (pillar start + slot inside the pillar) are gathered while the tile is in flight
    int pos[kIndexTileRows / kIndexThreads], rk[kIndexTileRows / kIndexThreads], ky[kIndexTileRows / kIndexThreads];
#pragma unroll
    for (int k = 0; k < kIndexTileRows / kIndexThreads; ++k) {
        const int r = k * kIndexThreads + tid;
        rk[k] = (r < rows) ? ranks[row0 + r] : -1;
        ky[k] = (r < rows) ? keys[row0 + r] : -1;
        pos[k] = (rk[k] >= 0) ? __ldg(starts + rk[k]) + slots[row0 + r] : -1;
    }
    if (bulk_bytes) mbar_wait(&bar, 0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIndexTileRows / kIndexThreads; ++k) {
        if (pos[k] < 0) continue;
        const int r = k * kIndexThreads + tid;
        // FRAMES: the input rows have no batch column -- logical column c >= 1 is input column c - 1.  Column 0 of the grouped
        // row carries the pillar's merged key (:208-210) instead of the frame index (= key / (nx ny nz)): the table kernel
        // decodes the pillar centre and the coords from it.
        const float *p = tile + r * in_cols - (FRAMES ? 1 : 0);
        const float c0 = __int_as_float(ky[k]), tail0 = __int_as_float((int)(row0 + r)), tail1 = __int_as_float(rk[k]);
        if (rs == 8) {
            // one 256-bit store (STG.256) = one full 32-byte sector per row: the scatter is bound by store requests
            float v[8];
            v[0] = c0;
#pragma unroll
            for (int c = 1; c < 8; ++c) v[c] = c < cols ? p[c] : (c == 6 ? tail0 : (c == 7 ? tail1 : 0.0f));
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(grows + ((size_t)pos[k] + 1) * 8), "f"(v[0]),
                         "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                         : "memory");
            continue;
        }
        float4 *dst = reinterpret_cast<float4 *>(grows + ((size_t)pos[k] + 1) * rs);
        dst[0] = make_float4(c0, p[1], p[2], p[3]);
        for (int c4 = 4; c4 < rs; c4 += 4) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c4 + i;
                v[i] = c < cols ? p[c] : (c == rs - 2 ? tail0 : (c == rs - 1 ? tail1 : 0.0f));
            }
            dst[c4 >> 2] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

__global__ void publish_counters_kernel(const int32_t *__restrict__ counters, volatile int32_t *host_mapped) {
    if (threadIdx.x < RDP_NUM_COUNTERS) host_mapped[threadIdx.x] = counters[threadIdx.x];
    __threadfence_system();
}

// ----------------------------------------------------------------------------- K6 (rdp_table.cuh), generic form
__global__ void __launch_bounds__(256) pillar_table_kernel(const __grid_constant__ TableArgs t) {
    pdl_wait();
    pillar_table_body(t);
}

// ----------------------------------------------------------------------------- pillar-id lookup
// Dense (B, ny, nx) map cell -> pillar row (or -1): what the first SubMConv2d of SparseEnc needs to build its rule
// book (spconv_backbone_2d.py:262-271) without a hash pass of its own -- the occupancy bitmap + rank prefix already are
// that table.  32 x 32 tiles: keys are x-major (b*nx*ny + cx*ny + cy), the output is y-major, so the tile is transposed
// through shared memory and both sides stay coalesced.
__global__ void __launch_bounds__(256) pillar_lookup_kernel(const uint2 *__restrict__ wordrank,
                                                            int nx, int ny, int32_t *__restrict__ lookup) {
    __shared__ int tile[32][33];
    const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int x = x0 + j, y = y0 + tx;
        int v = -1;
        if (x < nx && y < ny) {
            const long long key = ((long long)b * nx + x) * ny + y;
            const uint2 wr = wordrank[key >> 5];
            const uint32_t bit = 1u << (key & 31);
            if (wr.x & bit) v = (int)(wr.y + __popc(wr.x & (bit - 1u)));
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int y = y0 + j, x = x0 + tx;
        if (x < nx && y < ny) lookup[((size_t)b * ny + y) * nx + x] = tile[tx][j];
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rdp

using namespace rdp;

namespace rdp {

GeomDev make_geom_dev(const rdp_geom_t *geom) {
    GeomDev g;
    g.lo_x = geom->lo[0]; g.lo_y = geom->lo[1]; g.lo_z = geom->lo[2];
    g.vx = geom->vsz[0]; g.vy = geom->vsz[1]; g.vz = geom->vsz[2];
    g.nx = geom->nx; g.ny = geom->ny; g.nz = geom->nz > 1 ? geom->nz : 1;
    g.batch = geom->batch_size; g.cols = geom->cols;
    g.off_x = geom->off[0]; g.off_y = geom->off[1]; g.off_z = geom->off[2];
    return g;
}

int index_fwd_impl(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                   void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse, int32_t *counts, int32_t *counters,
                   int32_t *host_mapped, void *event_v, cudaStream_t stream, bool skip_table) {
    cudaEvent_t event = static_cast<cudaEvent_t>(event_v);
    // N, P and the error flags are final after the bitmap scan (K2): publish them there, so the host learns the output
    // sizes ~50 us into the call and can enqueue whatever follows while the remaining kernels run.
    auto publish = [&]() -> int {
        if (host_mapped) publish_counters_kernel<<<1, 32, 0, stream>>>(counters, host_mapped);
        if (event) RDP_CUDA_OK(cudaEventRecord(event, stream));
        return RDP_OK;
    };
    if (!geom || !counters || n_points < 0 || (coord_cols != 3 && coord_cols != 4)) return RDP_ERR_INVALID_ARG;
    if (geom->nz > 1 && coord_cols != 4) return RDP_ERR_INVALID_ARG;   // voxel coords are [b, z, y, x]
    if (n_points > 0 && (!points || !workspace || !coords || !inverse || !counts)) return RDP_ERR_INVALID_ARG;
    if (!aligned16(points) || !aligned16(coords) || !aligned16(inverse) || !aligned16(counts) || !aligned16(workspace)) return RDP_ERR_INVALID_ARG;
    if (n_points >= (1ll << 31) - 8) return RDP_ERR_UNSUPPORTED;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, nullptr, &ws);
    if (rc != RDP_OK) return rc;
    if (n_points > 0 && ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;

    RDP_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t) * RDP_NUM_COUNTERS, stream));
    if (n_points == 0) return publish();
    RDP_CUDA_OK(cudaMemsetAsync(ws.zero_begin, 0, ws.zero_bytes, stream));

    const GeomDev g = make_geom_dev(geom);
    const size_t smem = (size_t)kIndexTileRows * geom->cols * sizeof(float);
    if (smem > 200 * 1024) return RDP_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        RDP_CUDA_OK(cudaFuncSetAttribute(quantize_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RDP_CUDA_OK(cudaFuncSetAttribute(group_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RDP_CUDA_OK(cudaFuncSetAttribute(group_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int tiles = (int)ws.index_tiles;
    quantize_mark_kernel<<<tiles, kIndexThreads, smem, stream>>>(points, n_points, g, ws.bitmap, ws.keys, ws.tile_keep, counters,
                                                                 frame_offsets, counts);
    RDP_CUDA_OK(launch_pdl(bitmap_rank_kernel, kScanGrid, kScanThreads, 0, stream, ws.bitmap, ws.words, ws.scan_state_a, ws.wordrank, counters,
                           host_mapped));
    if (event) RDP_CUDA_OK(cudaEventRecord(event, stream));
    RDP_CUDA_OK(launch_pdl(rank_count_kernel, (tiles + kRankSub - 1) / kRankSub, kIndexThreads, 0, stream, ws.keys, ws.ranks, n_points, ws.wordrank, ws.tile_keep, inverse,
                           counts, counters, ws.orig2kept, ws.kept2orig, ws.slots));
    RDP_CUDA_OK(launch_pdl(count_scan_kernel, kScanGrid, kScanThreads, 0, stream, counts, ws.scan_state_b, ws.starts, ws.tile_first, counters));
    if (frame_offsets)
        RDP_CUDA_OK(launch_pdl(group_rows_kernel<true>, tiles, kIndexThreads, smem, stream, points, ws.keys, ws.ranks, ws.slots, n_points,
                               geom->cols, ws.starts, ws.grows));
    else
        RDP_CUDA_OK(launch_pdl(group_rows_kernel<false>, tiles, kIndexThreads, smem, stream, points, ws.keys, ws.ranks, ws.slots, n_points,
                               geom->cols, ws.starts, ws.grows));
    if (!skip_table) {   // the fused forward builds the table (and the coords) inside its tile kernels instead
        TableArgs t;
        t.grows = ws.grows; t.starts = ws.starts; t.counters = counters; t.aux = ws.aux; t.coords = coords;
        t.rs = grouped_row_floats(geom->cols); t.coord_cols = coord_cols; t.g = g;
        RDP_CUDA_OK(launch_pdl(pillar_table_kernel, table_grid(ws.pcap), 256, 0, stream, t));
    }
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

}  // namespace rdp

extern "C" int rdp_index_fwd_frames(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom,
                                    int32_t coord_cols, void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse,
                                    int32_t *counts, int32_t *counters, int32_t *host_mapped, void *event_v, void *stream_v) {
    return index_fwd_impl(points, frame_offsets, n_points, geom, coord_cols, workspace, workspace_bytes, coords, inverse, counts,
                          counters, host_mapped, event_v, static_cast<cudaStream_t>(stream_v), false);
}

extern "C" int rdp_index_fwd_publish(const float *points, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                                     void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse,
                                     int32_t *counts, int32_t *counters, int32_t *host_mapped, void *event_v, void *stream_v) {
    return rdp_index_fwd_frames(points, nullptr, n_points, geom, coord_cols, workspace, workspace_bytes, coords, inverse, counts,
                                counters, host_mapped, event_v, stream_v);
}

extern "C" int rdp_index_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                             void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse,
                             int32_t *counts, int32_t *counters, void *stream_v) {
    return rdp_index_fwd_publish(points, n_points, geom, coord_cols, workspace, workspace_bytes, coords, inverse, counts, counters,
                                 nullptr, nullptr, stream_v);
}

extern "C" int rdp_pillar_lookup(int64_t n_points, const rdp_geom_t *geom, void *workspace, size_t workspace_bytes, int32_t *lookup,
                                 void *stream_v) {
    if (!geom || !lookup || n_points < 0) return RDP_ERR_INVALID_ARG;
    Workspace ws;
    int rc = carve_workspace(workspace, n_points, geom, nullptr, &ws);
    if (rc != RDP_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (n_points == 0) {   // rdp_index_fwd did not touch the workspace: every cell is empty
        RDP_CUDA_OK(cudaMemsetAsync(lookup, 0xff, sizeof(int32_t) * (size_t)geom->batch_size * geom->nx * geom->ny, stream));
        return RDP_OK;
    }
    if (!workspace || ws.index_bytes > workspace_bytes) return RDP_ERR_WORKSPACE;
    dim3 grid((geom->nx + 31) / 32, (geom->ny + 31) / 32, geom->batch_size);
    if (grid.y > 65535 || grid.z > 65535) return RDP_ERR_UNSUPPORTED;
    if (geom->nz > 1) return RDP_ERR_UNSUPPORTED;   // the dense lookup is a 2-D (pillar) table
    pillar_lookup_kernel<<<grid, 256, 0, stream>>>(ws.wordrank, geom->nx, geom->ny, lookup);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}

extern "C" int rdp_publish_counters(const int32_t *counters, int32_t *host_mapped, void *stream_v) {
    if (!counters || !host_mapped) return RDP_ERR_INVALID_ARG;
    publish_counters_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_v)>>>(counters, host_mapped);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}
