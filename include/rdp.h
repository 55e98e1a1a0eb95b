/*
 * rdp.h -- C ABI of librdp.so: the B200 (sm_100a) dynamic pillar encoder.
 *
 * Drop-in boundary for RadarDistill's pillar-encoding front end.  Every entry point
 * replaces a stretch of /root/reference/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py
 * (cited per function).  Plain C: raw device pointers, sizes and a CUDA stream handle
 * (passed as void* == cudaStream_t); no torch types.  All functions return an int status
 * (RDP_OK == 0, negative = error), never throw, keep no global state, and are re-entrant
 * per (stream, workspace).  The caller owns every buffer.  Nothing here synchronises the
 * stream; the only host<->device traffic is what the caller does with `counters`.
 *
 * Calling sequence for one forward:
 *     rdp_workspace_bytes -> (allocate ws, outputs at capacity n_points)
 *     rdp_index_fwd       -> coords / inverse / counts, counters[N,P] on the device
 *     rdp_pfn_fwd         -> features (+ argmax, pillar_mean, bn state in train mode)
 *     (caller copies `counters` back once to learn N and P and narrows its views)
 *     rdp_pfn_bwd         -> dW, dgamma, dbeta             (training only)
 */
#ifndef RDP_H_
#define RDP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDP_ABI_VERSION 7

#if defined(__GNUC__)
#define RDP_API __attribute__((visibility("default")))
#else
#define RDP_API
#endif

enum rdp_status {
    RDP_OK = 0,
    RDP_ERR_INVALID_ARG = -1,  /* null pointer, non-positive size, misaligned buffer       */
    RDP_ERR_WORKSPACE = -2,    /* workspace smaller than rdp_workspace_bytes says           */
    RDP_ERR_CUDA = -3,         /* a CUDA runtime call failed (see rdp_last_cuda_error)      */
    RDP_ERR_UNSUPPORTED = -4,  /* configuration outside what the kernels implement          */
    RDP_ERR_KEYSPACE = -5      /* batch_size*nx*ny does not fit a non-negative int32 key    */
};

/* counters[] slots (int32, device memory, RDP_NUM_COUNTERS entries, written by rdp_index_fwd) */
enum rdp_counter {
    RDP_CNT_N = 0,       /* points kept by the range mask              (dynamic_pillar_vfe.py:203-206) */
    RDP_CNT_P = 1,       /* pillars == rows of features / coords       (:212)                          */
    RDP_CNT_ERRFLAGS = 2 /* bit0: a row's batch index was outside [0, batch_size)                      */
};
#define RDP_NUM_COUNTERS 16

#define RDP_LAYOUT_SIMPLE2D 0  /* DynamicPillarVFESimple2D + radar subclasses (:219-237): [center|pts|cluster|dist|rel] */
#define RDP_LAYOUT_DYNPILLAR 1 /* DynamicPillarVFE (:113-121):                            [pts|cluster|center|dist]    */
#define RDP_LAYOUT_DYNVOXEL 2  /* DynamicVoxelVFE (dynamic_voxel_vfe.py:73-91): as DYNPILLAR with a z-aware voxel centre;
                                  rdp_decorate only (geom->nz > 1) -- the fused single-layer kernels are pillar-only     */

/* Grid geometry: the constructor arguments of the reference classes (:73-85, :177-189). */
typedef struct rdp_geom {
    float lo[3];        /* point_cloud_range[0:3]                                            */
    float vsz[3];       /* voxel_size, fp32                                                  */
    float off[3];       /* voxel/2 + lo  (x_offset, y_offset, z_offset), rounded once to fp32 */
    int32_t nx, ny;     /* grid_size[0], grid_size[1]                                        */
    int32_t batch_size; /* frames in this call; rows carry their frame index in column 0     */
    int32_t cols;       /* floats per row = 1 + num_point_features                           */
    int32_t nz;         /* <= 1: pillars (z neither quantised nor masked, dynamic_pillar_vfe.py:201-206);
                           > 1: voxels, grid_size[2] (dynamic_voxel_vfe.py:57-66, dynamic_mean_vfe.py:52-60): z is
                           quantised and masked too, the key gains cz and coords are [b, z, y, x]          */
} rdp_geom_t;

/* Feature layout + PFN shape: model_cfg of the reference classes (:53-62, :150-166). */
typedef struct rdp_layout {
    int32_t layout;        /* RDP_LAYOUT_*                        */
    int32_t use_abs;       /* USE_ABSLOTE_XYZ                     */
    int32_t use_cluster;   /* USE_CLUSTER_XYZ (Simple2D only)     */
    int32_t use_relative;  /* USE_RELATIVE_XYZ (Simple2D only)    */
    int32_t with_distance; /* WITH_DISTANCE                       */
    int32_t c_in, c_out;   /* PFNLayerV2 linear: (c_out, c_in)    */
    int32_t coord_cols;    /* 3: [b,y,x]   4: [b,0,y,x]           */
} rdp_layout_t;

/* Batch-norm / linear parameters of the single PFNLayerV2 (:14-46).  Device pointers. */
typedef struct rdp_pfn_params {
    const float *weight;  /* (c_out, c_in) row-major, linear.weight                          */
    const float *bias;    /* (c_out) linear.bias when USE_NORM is false, else NULL           */
    const float *gamma;   /* (c_out) norm.weight, NULL when USE_NORM is false                */
    const float *beta;    /* (c_out) norm.bias                                               */
    float *running_mean;  /* (c_out) read in eval mode, updated in place in train mode       */
    float *running_var;   /* (c_out)                                                         */
    double eps;           /* 1e-3  (:29)                                                     */
    double momentum;      /* 0.01  (:29)                                                     */
    int32_t train_bn;     /* 1: batch statistics + running-stat update; 0: running stats     */
    int64_t *num_batches_tracked; /* norm.num_batches_tracked (device, int64): += 1 by a train-mode forward over more
                                     than one point, as BatchNorm1d does; may be NULL                               */
    /* SyncBatchNorm (tools/train.py:34,144-145): batch statistics over the points of ALL ranks.  The library issues no
       collective; the caller all-reduces two small fp64 vectors of the workspace (rdp_stats_buffers) between phases:
         forward : phase 1 = index + feature moments, stop;  all-reduce(SUM) the stats vector (keep a copy of the local one);
                   phase 2 = fold the (global) moments into bn_state + the rest of the forward, `local_stats` = the copy
         backward: phase 1 = the backward sums, stop;  all-reduce(SUM) a COPY of the bwd vector;
                   phase 2 = closed-form epilogue with `global_bwd` = that all-reduced copy
       0 = everything in one call (per-rank statistics, the reference default).                                    */
    int32_t stats_phase;
    const double *local_stats;    /* forward phase 2: this rank's moments (device), as left by phase 1              */
    const double *global_bwd;     /* backward phase 2: the all-reduced backward sums (device)                        */
} rdp_pfn_params_t;

RDP_API int rdp_abi_version(void);
RDP_API const char *rdp_status_string(int status);
/* Text of the last CUDA error seen by the calling thread inside librdp ("" if none). */
RDP_API const char *rdp_last_cuda_error(void);

/* Bytes of scratch rdp_index_fwd / rdp_pfn_fwd / rdp_pfn_bwd need for up to n_points rows. */
RDP_API int rdp_workspace_bytes(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, size_t *bytes);

/*
 * Quantise -> range mask -> merged key -> sorted unique -> inverse / counts / coords.
 * Replaces dynamic_pillar_vfe.py:201-212 and :243-248 (also :93-103, :132-138).
 *   points   (n_points, cols) fp32, row-major, 16-byte aligned
 *   coords   (cap n_points, coord_cols) int32   rows [0,P) valid, ascending merged key
 *   inverse  (cap n_points) int32               entries [0,N) valid: pillar of the j-th KEPT point
 *   counts   (cap n_points) int32               entries [0,P) valid
 *   counters (RDP_NUM_COUNTERS) int32 device
 */
RDP_API int rdp_index_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                  void *workspace, size_t workspace_bytes,
                  int32_t *coords, int32_t *inverse, int32_t *counts, int32_t *counters, void *stream);

/*
 * rdp_index_fwd that also tells the host N and P as early as they exist: right after the bitmap scan (two kernels
 * into the call) a one-warp kernel copies counters[] to `host_mapped` (pinned, device-visible host memory; NULL to
 * skip) and `event` (a cudaEvent_t passed as void*; NULL to skip) is recorded on `stream`.  A host thread that
 * waits on the event learns the output sizes while the remaining index kernels (and whatever the caller enqueues
 * behind them) are still running -- the reference's size read-backs (:204-206 boolean mask, :212 unique) stall the
 * stream instead.  Everything else is rdp_index_fwd.
 */
RDP_API int rdp_index_fwd_publish(const float *points, int64_t n_points, const rdp_geom_t *geom, int32_t coord_cols,
                                  void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse,
                                  int32_t *counts, int32_t *counters, int32_t *host_mapped, void *event, void *stream);

/*
 * Device-side input preparation (replaces the batch-index padding of collate_batch, pcdet/datasets/dataset_distill.py:237-244,
 * and its 4 bytes per point of upload): `points` holds the frames back to back WITHOUT the batch column, i.e.
 * (n_points, cols - 1) fp32, and frame b owns rows [frame_offsets[b], frame_offsets[b + 1]) (int32, device,
 * batch_size + 1 entries, ascending, frame_offsets[0] == 0, frame_offsets[batch_size] == n_points).  geom->cols still
 * counts the batch column: every output (and the workspace's grouped rows) is identical to what rdp_index_fwd_publish
 * produces for the padded rows.  frame_offsets == NULL is rdp_index_fwd_publish.
 */
RDP_API int rdp_index_fwd_frames(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom,
                                 int32_t coord_cols, void *workspace, size_t workspace_bytes, int32_t *coords, int32_t *inverse,
                                 int32_t *counts, int32_t *counters, int32_t *host_mapped, void *event, void *stream);

/*
 * scatter_mean -> decorated features -> Linear + BatchNorm1d + ReLU -> scatter_max.
 * Replaces dynamic_pillar_vfe.py:214-240 with PFNLayerV2.forward :35-46 (last layer).
 * Must follow rdp_index_fwd on the same points / workspace / stream.
 *   features    (cap n_points, c_out) fp32      rows [0,P)
 *   argpos      (cap n_points, c_out) int32     winning row of every (pillar, channel) as a position in the
 *                                               workspace's pillar-grouped order (lowest kept index on ties),
 *                                               -1 where the ReLU clamped the maximum to 0 (no gradient flows, :38);
 *                                               input of rdp_pfn_bwd / rdp_argmax_kept.
 *                                               NULL if not wanted.
 *   pillar_mean (cap n_points, 3) fp32          per-pillar xyz mean (scatter_mean, :226); NULL if not wanted
 *   bn_state    (rdp_bn_state_doubles(layout)) fp64: batch mean/var, folded scale/shift and the feature
 *               moments the backward needs; required when train_bn, else may be NULL
 */
RDP_API int rdp_pfn_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                        const rdp_pfn_params_t *params, void *workspace, size_t workspace_bytes,
                        const int32_t *counters, float *features, int32_t *argpos, float *pillar_mean,
                        double *bn_state, void *stream);

RDP_API int64_t rdp_bn_state_doubles(const rdp_layout_t *layout);

/* 1 if the fused kernels are compiled for this (row width, distance feature, c_out, feature layout), else 0: constructors
   validate their configuration with this instead of failing at the first forward. */
RDP_API int rdp_config_supported(const rdp_geom_t *geom, const rdp_layout_t *layout);

/* Byte offsets (inside the workspace) and lengths (in doubles) of the two fp64 vectors a SyncBatchNorm caller all-reduces
   between the phases described at rdp_pfn_params_t.stats_phase: the feature moments (+ the point count in the last slot)
   and the backward sums. */
RDP_API int rdp_stats_buffers(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, size_t *stats_offset,
                              int64_t *stats_doubles, size_t *bwd_offset, int64_t *bwd_doubles);

/*
 * rdp_index_fwd_publish followed by rdp_pfn_fwd in one call (one trip through the host binding per forward).
 * Same buffers and meaning as the two functions; replaces dynamic_pillar_vfe.py:201-249.
 */
RDP_API int rdp_encode_fwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                           const rdp_pfn_params_t *params, void *workspace, size_t workspace_bytes,
                           int32_t *coords, int32_t *inverse, int32_t *counts, int32_t *counters,
                           float *features, int32_t *argpos, double *bn_state,
                           int32_t *host_mapped, void *event, void *stream);
/* The same over frames without a batch column (rdp_index_fwd_frames + rdp_pfn_fwd). */
RDP_API int rdp_encode_fwd_frames(const float *points, const int32_t *frame_offsets, int64_t n_points, const rdp_geom_t *geom,
                                  const rdp_layout_t *layout, const rdp_pfn_params_t *params, void *workspace,
                                  size_t workspace_bytes, int32_t *coords, int32_t *inverse, int32_t *counts,
                                  int32_t *counters, float *features, int32_t *argpos, double *bn_state,
                                  int32_t *host_mapped, void *event, void *stream);

/*
 * Parameter gradients of the PFN (autograd of :35-46): argmax routing, ReLU', BatchNorm backward
 * (batch statistics when params->train_bn, running statistics otherwise), dW = g_x^T f.
 * Points are a non-differentiable leaf in the reference, so no point gradient is produced.
 *   grad_features (P, c_out) fp32 ; argpos / bn_state as written by rdp_pfn_fwd; same workspace.  `features` is not
 *   read (the ReLU mask travels in the sign of argpos) and may be NULL.
 *   d_weight (c_out, c_in), d_gamma (c_out) [NULL without norm], d_beta (c_out) [bias grad without norm]
 */
RDP_API int rdp_pfn_bwd(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                        const rdp_pfn_params_t *params, void *workspace, size_t workspace_bytes,
                        const int32_t *counters, const float *grad_features, const float *features,
                        const int32_t *argpos, const double *bn_state,
                        float *d_weight, float *d_gamma, float *d_beta, void *stream);

/*
 * scatter_max's argmax in the reference's numbering (:40): argmax_kept[p][c] = index, among the points kept by the
 * range mask (:204-206), of the row that attains features[p][c]; ties -> lowest index (torch_scatter CPU rule).
 */
RDP_API int rdp_argmax_kept(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, void *workspace,
                            size_t workspace_bytes, const int32_t *counters, const int32_t *argpos,
                            int32_t *argmax_kept, void *stream);

/*
 * Dense cell -> pillar lookup for the consumer's first sparse convolution (spconv_backbone_2d.py:262-271 builds a
 * SparseConvTensor from features + coords; its SubMConv2d rule book needs "which pillar sits at (b, y, x)"):
 *   lookup (batch_size, ny, nx) int32: row of features / coords at that cell, or -1.
 * Reads the occupancy bitmap + rank prefix rdp_index_fwd left in `workspace` (valid until the workspace is reused).
 */
RDP_API int rdp_pillar_lookup(int64_t n_points, const rdp_geom_t *geom, void *workspace, size_t workspace_bytes,
                              int32_t *lookup, void *stream);

/*
 * The one read-back of the path: N and P (the sizes the reference learns through the boolean-mask index at
 * dynamic_pillar_vfe.py:204-206 and torch.unique at :212).  Copies counters[] to `host_mapped` (pinned, device-visible host memory: cudaHostAlloc / torch pin_memory under UVA)
 * from a one-warp kernel instead of a DMA transfer, so the 64-byte read-back never queues behind bulk copies on the
 * copy engines.  The values are visible on the host once `stream` has been synchronised.
 */
RDP_API int rdp_publish_counters(const int32_t *counters, int32_t *host_mapped, void *stream);

/*
 * Host-buffer convenience (what a non-torch caller binds): uploads `points` (host), runs
 * rdp_index_fwd + rdp_pfn_fwd in eval mode, downloads the results and synchronises -- the whole eval-mode
 * forward of dynamic_pillar_vfe.py:195-252 (and :90-142, :255-373) including load_data_to_gpu
 * (pcdet/models/__init__.py:23-36) on the way in.
 * Host output buffers must hold n_points rows; *n_kept / *n_pillars receive N and P.
 * Parameters in `params` are HOST pointers here.
 */
RDP_API int rdp_encode_host(const float *points, int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout,
                    const rdp_pfn_params_t *params, float *features, int32_t *coords, int32_t *inverse,
                    int32_t *counts, int64_t *n_kept, int64_t *n_pillars);


/* ------------------------------------------------------------------------------------------------------------------
 * Stacked PFN layers and the sibling encoders.  All of them follow rdp_index_fwd (2-D key, or 3-D with
 * geom->nz > 1) on the same workspace / stream and reuse its pillar-grouped rows and pillar table.
 */

/*
 * The decorated per-point features the reference concatenates in front of its first PFNLayerV2, in KEPT-point order:
 *   features (cap n_points, layout->c_in) fp32, rows [0, N)
 * Replaces dynamic_pillar_vfe.py:214-237 (Simple2D), :105-121 (DynamicPillarVFE) and dynamic_voxel_vfe.py:73-91
 * (RDP_LAYOUT_DYNVOXEL: voxel centre incl. z).  layout->c_out is ignored.
 */
RDP_API int rdp_decorate(int64_t n_points, const rdp_geom_t *geom, const rdp_layout_t *layout, void *workspace,
                         size_t workspace_bytes, const int32_t *counters, float *features, void *stream);

/*
 * scatter_max over the pillars for the activations of a stacked layer (PFNLayerV2.forward :40):
 *   x (N, channels) fp32 in kept-point order -> out (cap P, channels) fp32, argmax_kept (cap P, channels) int32 = kept index
 *   of the winning row (lowest index on ties).  rdp_segment_max_bwd is its autograd: grad_x[argmax[p][c]][c] = grad_out[p][c]
 *   into a ZEROED grad_x (N, channels).
 */
RDP_API int rdp_segment_max_fwd(const float *x, int32_t channels, int64_t n_points, const rdp_geom_t *geom, void *workspace,
                                size_t workspace_bytes, const int32_t *counters, float *out, int32_t *argmax_kept, void *stream);
RDP_API int rdp_segment_max_bwd(const float *grad_out, const int32_t *argmax_kept, int64_t n_pillars, int32_t channels,
                                float *grad_x, void *stream);

/* DynamicMeanVFE (dynamic_mean_vfe.py:63-65): mean (cap P, cols - 1) fp32 = per-voxel mean of every point column. */
RDP_API int rdp_voxel_mean(int64_t n_points, const rdp_geom_t *geom, void *workspace, size_t workspace_bytes,
                           const int32_t *counters, float *mean, void *stream);

/*
 * Device-side input preparation: the range mask of the data processor (mask_points_by_range,
 * pcdet/datasets/processor/data_processor.py:80-86: lo <= x, y <= hi, inclusive) as a stable compaction, and, with
 * shuffle_seed != 0, the shuffle of :99-114 as a fixed pseudo-random permutation of the kept rows.
 *   points (n_points, cols) fp32, x in column x_col (y in x_col + 1); range_xy_lo_hi = {lo_x, lo_y, hi_x, hi_y} (host)
 *   out (cap n_points, cols); n_out (device int32) = rows kept; scratch of rdp_prepare_scratch_bytes(n_points) bytes
 */
RDP_API size_t rdp_prepare_scratch_bytes(int64_t n_points);
RDP_API int rdp_prepare_points(const float *points, int64_t n_points, int32_t cols, int32_t x_col, const float *range_xy_lo_hi,
                               uint64_t shuffle_seed, void *scratch, size_t scratch_bytes, float *out, int32_t *n_out, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The step's one collective (DistributedDataParallel's gradient averaging of the PFN parameters, tools/train.py:175-176)
 * as a one-shot all-reduce over NVLink peer memory: ONE kernel on the caller's stream packs this rank's gradient tensors
 * into its peer-mapped staging buffer, signals every peer, waits for every peer and sums all ranks' slots in rank order
 * (bit-identical results on all ranks), out = scale * sum.  The caller provides the mapping:
 *   peer_staging  device array of `world` pointers: every rank's staging buffer of rdp_allreduce_staging_bytes(n, world) bytes
 *                 as mapped into THIS process (CUDA IPC / torch symmetric memory), zero-filled before the first call
 *   step          1, 2, 3, ... -- the same on every rank for the same collective (each call uses the next number)
 * segments / counts: host arrays of n_segments (<= RDP_ALLREDUCE_MAX_SEGMENTS) device pointers and float counts.
 */
#define RDP_ALLREDUCE_MAX_SEGMENTS 16
RDP_API size_t rdp_allreduce_staging_bytes(int64_t n_floats, int32_t world);
RDP_API int rdp_allreduce_small(const float *const *segments, const int32_t *counts, int32_t n_segments, float *const *peer_staging,
                                int32_t rank, int32_t world, uint32_t step, float scale, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RDP_H_ */
