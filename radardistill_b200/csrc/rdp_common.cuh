// rdp_common.cuh -- shared internals of librdp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rdp.h"

namespace rdp {

// ----------------------------------------------------------------------------- status plumbing
void set_last_cuda_error(cudaError_t e, const char *where);

#define RDP_CUDA_OK(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            ::rdp::set_last_cuda_error(_e, #expr);          \
            return RDP_ERR_CUDA;                            \
        }                                                   \
    } while (0)

// ----------------------------------------------------------------------------- tiling constants
#ifndef RDP_INDEX_THREADS
#define RDP_INDEX_THREADS 256
#endif
constexpr int kIndexThreads = RDP_INDEX_THREADS;
constexpr int kIndexTileRows = 4 * RDP_INDEX_THREADS;  // points per CTA in the quantise / rank / fill kernels
constexpr int kRankSub = 2;            // index tiles per CTA of rank_count_kernel (8 rows per thread)
constexpr int kScanThreads = 256;
constexpr int kScanGrid = 296;        // 2 CTAs per SM: every CTA of a chunked scan is co-resident
constexpr int kPfnThreads = 128;
constexpr int kPfnWin = 128;          // grouped rows per PFN tile window (a tile owns the pillars that START in it)
#ifndef RDP_PFN_CAP
#define RDP_PFN_CAP 192
#endif
constexpr int kPfnCap = RDP_PFN_CAP;          // rows staged per tile: the window + 64 rows of overhang for the last pillar
#ifndef RDP_PFN_GRID_PER_SM
#define RDP_PFN_GRID_PER_SM 4
#endif
constexpr int kPfnGridCap = 148 * RDP_PFN_GRID_PER_SM;  // persistent PFN CTAs of the backward tile kernel
constexpr int kSliceInts = 136;       // slack (ints) behind starts[]: a 128-pillar slice + 1 can be read from any 16-byte aligned address
__host__ __device__ constexpr int grouped_row_floats(int cols) { return (cols + 2 + 3) / 4 * 4; }
constexpr int kMaxCin = 24;
constexpr int kMaxCout = 128;
constexpr int kMaxG = 14;             // reduced basis [row inputs (<= 8) | pillar constants (5)] upper bound (+1 spare)
constexpr int kMaxAcc = kMaxG + kMaxG * kMaxG + 2;   // first + second moments of the reduced basis, the point count

// extra counter slots (after the public ones of rdp.h)
constexpr int kCntTicketA = 4;   // bitmap scan ticket
constexpr int kCntTicketB = 5;   // count scan ticket
constexpr int kCntPfnBlocks = 6; // (reserved)
constexpr int kCntDoneStats = 8; // CTAs of the moments pass that have added their sums (the last one runs the BN epilogue)
constexpr int kCntDoneBwd = 9;   // same for the backward tile kernel
constexpr int kCntStatsTag = 10; // (reserved)

// ----------------------------------------------------------------------------- workspace layout
struct Workspace {
    // zeroed by one memset at the start of rdp_index_fwd: [zero_begin, zero_begin + zero_bytes)
    uint64_t *scan_state_a;  // kScanGrid
    uint64_t *scan_state_b;  // kScanGrid
    double *acc_stats;       // kMaxAcc        : fp64 totals of the train-mode feature moments (atomics; last CTA folds them)
    double *acc_bwd;         // kMaxCout*(kMaxG+1) : fp64 totals of the backward sums
    uint32_t *bitmap;        // words
    // not zeroed
    uint2 *wordrank;         // words : {bitmap word, exclusive pillar rank of the word} -- one 8-byte gather per point
    int32_t *keys;           // n  merged key (-1 = dropped)
    int32_t *ranks;          // n  pillar rank of the row (-1 = dropped)
    int32_t *slots;          // n  position of the row inside its pillar (arrival order of rank_count_kernel's atomics)
    int32_t *tile_keep;      // index tiles
    int32_t *starts;         // pcap + 1  exclusive start of every pillar in the grouped order; starts[P] = N
    float *grows;            // (1 + n + pad) * RS : rows physically grouped by pillar (pillar order == key order);
                             // RS floats per row = [merged key | xyz | features | pad | original row id | pillar id];
                             // row 0 is a sentinel (pillar id -1) in front of grouped position 0
    char *zero_begin;
    size_t zero_bytes;
    float *aux;              // (pcap + pad) * 8 : per-pillar table [centre xy | centre - mean xyz | first grouped row | rows | lowest original row id]
    int32_t *tile_first;     // PFN tiles + 2   : first pillar that starts at or after grouped row 128 t
    int32_t *orig2kept;      // n     (only written / read when the range mask dropped rows)
    int32_t *kept2orig;      // n
    int64_t words, n, pcap, index_tiles, pfn_tiles;
    size_t index_bytes;      // bytes rdp_index_fwd needs
    size_t total_bytes;
};

// Carves the workspace; `base` may be null to only compute total_bytes.
int carve_workspace(void *base, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, Workspace *ws);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// --- mbarrier + 1-D bulk (TMA) copy: global -> shared, completion on an mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes must be a multiple of 16, src and dst 16-byte aligned.  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// (sx, sy, sz) / cnt, each correctly rounded in fp64 and then rounded once to fp32 -- bit-identical to the three IEEE
// divisions of the oracle, but with one reciprocal: y = RN(1/b), q = RN(a y), r = a - b q (exact, FMA), RN(q + r y) is the
// correctly rounded quotient (Markstein); b is a small exact integer here.
__device__ __forceinline__ void mean3(double sx, double sy, double sz, int cnt, float *mx, float *my, float *mz) {
    const double b = (double)cnt, y = __drcp_rn(b);
    const double qx = __dmul_rn(sx, y), qy = __dmul_rn(sy, y), qz = __dmul_rn(sz, y);
    *mx = (float)__fma_rn(__fma_rn(-qx, b, sx), y, qx);
    *my = (float)__fma_rn(__fma_rn(-qy, b, sy), y, qy);
    *mz = (float)__fma_rn(__fma_rn(-qz, b, sz), y, qz);
}

// --- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may be scheduled while its predecessor in the
//     stream drains, so the launch latency and the ramp-up of its CTAs hide behind the predecessor's tail.  pdl_wait() -- the
//     FIRST statement of every such kernel -- blocks until the predecessor grid has completed and its writes are visible, so
//     ordering is exactly that of a plain stream (a no-op when the kernel was launched without the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// --- block-wide exclusive scan of one int per thread (blockDim.x == 256), returns exclusive prefix,
//     *total receives the block sum.  `sm` needs 9 ints.
__device__ __forceinline__ int block_excl_scan_256(int v, int *sm, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 8 ? sm[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < 8) sm[lane] = winc - w;
        if (lane == 7) sm[8] = winc;
    }
    __syncthreads();
    int res = sm[warp] + inc - v;
    *total = sm[8];
    __syncthreads();
    return res;
}

// --- chunked-scan hand-off: CTA with ticket t publishes its aggregate and returns the sum of the
//     aggregates of tickets < t.  Tickets are handed out in start order, so a CTA only ever waits
//     for CTAs that are already running (no residency assumption, no deadlock).
//     state[] must be zero before the launch.  Call from all threads; `sm` needs 1 uint32.
__device__ __forceinline__ uint32_t chunk_exclusive_prefix(uint64_t *state, int ticket, uint32_t aggregate, uint32_t *sm) {
    if (threadIdx.x == 0) {
        __threadfence();
        atomicExch(reinterpret_cast<unsigned long long *>(state + ticket), (1ull << 63) | aggregate);
    }
    if (threadIdx.x < 32) {
        // all look-back loads go out together (one L2 round trip when the predecessors have published); only the entries
        // that were not ready are polled again
        constexpr int Q = (kScanGrid + 31) / 32;
        unsigned long long v[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = (int)threadIdx.x + 32 * q;
            v[q] = (j < ticket) ? *reinterpret_cast<volatile unsigned long long *>(state + j) : (1ull << 63);
        }
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int j = (int)threadIdx.x + 32 * q;
            while (!(v[q] >> 63)) {
                __nanosleep(20);
                v[q] = *reinterpret_cast<volatile unsigned long long *>(state + j);
            }
            sum += static_cast<uint32_t>(v[q]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (threadIdx.x == 0) *sm = sum;
    }
    __syncthreads();
    uint32_t r = *sm;
    __syncthreads();
    return r;
}

#endif  // __CUDACC__

// Launch with the programmatic-stream-serialization attribute (see pdl_wait).  RDP_NO_PDL=1 falls back to plain launches.
bool pdl_enabled();
#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace rdp
