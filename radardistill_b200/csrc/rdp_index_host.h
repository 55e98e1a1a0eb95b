// rdp_index_host.h -- internal entry point of the index pass, shared by rdp_index.cu and the fused forward (rdp_pfn.cu).
#pragma once
#include "rdp_table.cuh"

namespace rdp {

GeomDev make_geom_dev(const rdp_geom_t *geom);

int index_fwd_impl(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                   void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse, int32_t *counts, int32_t *counters,
                   int32_t *host_mapped, void *event, cudaStream_t stream, bool skip_table);

inline int table_grid(int64_t pcap) {
    const int64_t blocks = (pcap + 255) / 256;
    const int64_t cap = 148 * 8;   // persistent: 8 CTAs of 256 threads per SM
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace rdp
