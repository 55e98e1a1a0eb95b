// rdp_allreduce.cu -- the step's one collective as a one-shot all-reduce over NVLink peer memory.
//
// What is reduced: the PFN parameter gradients of the two encoders (1 056 floats) -- DistributedDataParallel's gradient
// averaging (tools/train.py:175-176).  NCCL needs ~33 us for this message on 8 B200s (launch on its own stream, two
// cross-stream event hops, LL protocol; measured, tools/dbg_allreduce.py); for a 4 KB vector the transfer itself is noise,
// so the collective is written as ONE kernel on the caller's stream over peer-mapped buffers (every rank maps every
// rank's staging buffer: torch symmetric memory / CUDA IPC does the mapping, this file does the data path):
//   1. pack: gather the gradient tensors into this rank's staging slot (double buffered by step parity);
//   2. signal: release-store the step number into every peer's signal word for this rank (P2P stores over NVLink);
//   3. wait: acquire-spin on this rank's own signal words until every peer has published this step;
//   4. reduce: every rank reads all `world` staging slots (P2P loads, cache-bypassing) and sums them in rank order -- the
//      same order on every rank, so all ranks hold bit-identical averages.
// Step numbers only grow, so no word is ever reset; a rank can run at most one step ahead of its slowest peer, which is
// what makes two staging slots enough.
#include "rdp_common.cuh"

namespace rdp {

struct GradSegments {
    const float *ptr[RDP_ALLREDUCE_MAX_SEGMENTS];
    int32_t count[RDP_ALLREDUCE_MAX_SEGMENTS];
    int32_t n_segments;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// staging buffer of a rank: [2 slots x n floats | world signal words (uint32), 128-byte aligned]
__global__ void __launch_bounds__(1024) allreduce_oneshot_kernel(const __grid_constant__ GradSegments segs, float *const *peers, int rank, int world,
                                                                 int n, int n_pad, uint32_t step, float scale, float *out) {
    __shared__ int s_timeout;
    const int tid = threadIdx.x;
    if (tid == 0) s_timeout = 0;
    float *mine = peers[rank] + (size_t)(step & 1u) * n;
    int off = 0;
    for (int s = 0; s < segs.n_segments; ++s) {
        for (int i = tid; i < segs.count[s]; i += blockDim.x) mine[off + i] = segs.ptr[s][i];
        off += segs.count[s];
    }
    __syncthreads();
    if (tid < world) {
        __threadfence_system();
        uint32_t *peer_signals = reinterpret_cast<uint32_t *>(peers[tid] + 2 * (size_t)n_pad);
        st_release_sys(peer_signals + rank, step);
        const uint32_t *my_signals = reinterpret_cast<const uint32_t *>(peers[rank] + 2 * (size_t)n_pad);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(my_signals + tid) - step) < 0) {   // wrap-safe "published step >= this step"
            if (clock64() - t0 > 240000000000ll) { s_timeout = 1; break; }   // ~2 min: a peer died -- poison the result instead of hanging forever
        }
    }
    __syncthreads();
    if (s_timeout) {
        for (int i = tid; i < n; i += blockDim.x) out[i] = __int_as_float(0x7fc00000);
        return;
    }
    for (int i = tid; i < n; i += blockDim.x) {
        float acc = 0.0f;
        for (int r = 0; r < world; ++r) acc += ld_relaxed_sys(peers[r] + (size_t)(step & 1u) * n + i);   // fixed order: identical on every rank
        out[i] = acc * scale;
    }
}

}  // namespace rdp

using namespace rdp;

extern "C" size_t rdp_allreduce_staging_bytes(int64_t n_floats, int32_t world) {
    const size_t n_pad = ((size_t)(n_floats > 0 ? n_floats : 1) + 31) / 32 * 32;
    return sizeof(float) * 2 * n_pad + sizeof(uint32_t) * (((size_t)world + 31) / 32 * 32);
}

extern "C" int rdp_allreduce_small(const float *const *segments, const int32_t *counts, int32_t n_segments, float *const *peer_staging,
                                   int32_t rank, int32_t world, uint32_t step, float scale, float *out, void *stream_v) {
    if (!segments || !counts || !peer_staging || !out || n_segments < 1 || n_segments > RDP_ALLREDUCE_MAX_SEGMENTS || world < 1 || rank < 0 ||
        rank >= world || world > 1024 || step == 0)
        return RDP_ERR_INVALID_ARG;
    GradSegments g;
    int64_t n = 0;
    for (int s = 0; s < n_segments; ++s) {
        if (!segments[s] || counts[s] < 0) return RDP_ERR_INVALID_ARG;
        g.ptr[s] = segments[s]; g.count[s] = counts[s];
        n += counts[s];
    }
    g.n_segments = n_segments;
    if (n > (1 << 20)) return RDP_ERR_UNSUPPORTED;   // a latency-optimised path for small vectors; large ones belong to NCCL
    const int n_pad = (int)((n + 31) / 32 * 32);
    allreduce_oneshot_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream_v)>>>(g, peer_staging, rank, world, (int)n, n_pad, step, scale, out);
    RDP_CUDA_OK(cudaGetLastError());
    return RDP_OK;
}
