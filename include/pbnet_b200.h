/*
 * pbnet_b200.h — C ABI of libpbnet_b200.so: the B200-native (sm_100a) replacement of PBNet's
 * post-backbone instance-grouping hot path.
 *
 * Every entry point below replaces one interface of the reference (paths under /root/reference):
 *
 *   pb_binary_cluster          <-  void binary_cluster(at::Tensor x17, int, float, bool)
 *                                  lib/PB_lib/src/pbnet/cluster.h:13-18, cluster.cu:16-119,
 *                                  bound to Python at lib/PB_lib/src/PB_lib_api.cpp:7 and called from
 *                                  lib/PB_lib/torch_io/pbnet_ops.py:73-74
 *   pb_binary_cluster_batched  <-  the 18-iteration per-class Python loop around that call,
 *                                  network/PBNet.py:151-179 (one launch for many calls)
 *   pb_create / pb_destroy     <-  class BINARY::Solver ctor / (missing) dtor,
 *                                  lib/PB_lib/src/pbnet/binary.cuh:33-95, binary.cu:19-47
 *   error codes                <-  CUDA_ERR_CHK -> exit(code), lib/PB_lib/src/pbnet/binary.cuh:20-27
 *
 * Plain pointers and sizes only; no torch types.  All arrays are 1-D contiguous.  fp32 / int32.
 * The library never falls back to a CPU path: without a usable CUDA device pb_create fails.
 */
#ifndef PBNET_B200_H
#define PBNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

/* status codes (returned by every int function; text via pb_last_error) */
enum {
    PB_OK = 0,
    PB_ERR_ARG = 1,         /* null pointer, negative size, sum(seg_counts) != n_pts ... */
    PB_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    PB_ERR_SEM_RANGE = 3,   /* a class id outside [2,19] (the reference indexes 18-entry tables with sem-2) */
    PB_ERR_NONFINITE = 4,   /* NaN / Inf coordinate */
    PB_ERR_RANGE = 5,       /* a segment spans more than 16383 grid cells along one axis */
    PB_ERR_MIXED_CLASS = 6, /* a segment mixes classes whose radius table entries differ (undefined in the reference) */
    PB_ERR_CAPACITY = 7,    /* center / clt_sem capacity too small for the clusters found */
    PB_ERR_NOMEM = 8
};

/* where the data pointers of a call live */
enum { PB_MEM_HOST = 0, PB_MEM_DEVICE = 1 };

/* Creates a context bound to CUDA device `device`: owns a stream, a growable device workspace
 * (no per-call cudaMalloc) and pinned scratch.  One context per host thread. */
int pb_create(int device, pb_ctx **out);
void pb_destroy(pb_ctx *ctx);
const char *pb_last_error(const pb_ctx *ctx);

/* number of kernel launches issued by the last pb_binary_cluster* call on this context */
int64_t pb_last_launch_count(const pb_ctx *ctx);

/*
 * One reference call.  Segment b covers points [off_b, off_b + seg_counts[b]).
 *
 * in   x,y,z        shifted coordinates (clustering space)            f32[n_pts]
 *      xo,yo,zo     original coordinates (LP assignment space)        f32[n_pts]
 *      sem          class per point, in [2,19]                        i32[n_pts]
 *      seg_counts   points per segment — ALWAYS a host pointer        i32[n_seg]
 *      radius       per-class radius table (index sem-2)              f32[18]   host pointer
 *      min_pts      per-class HP threshold (index sem-2)              i32[18]   host pointer
 *      para_f       fragment filter factor (reference passes 0.05)
 *      assign_lp    the reference's nv_flag
 * out  cluster_id   -1 or id in [0, n_clusters); ids run across segments  i32[n_pts]
 *      cluster_num  clusters per segment                              i32[n_seg]
 *      degree       raw neighbour count (pbnet_ops returns degree+1)  i32[n_pts]
 *      center       xyz per cluster, shifted space                    f32[3*n_clusters] (capacity center_cap floats)
 *      clt_sem      class per cluster                                 i32[n_clusters]   (capacity clt_sem_cap)
 *      n_clusters_out  total clusters — host pointer
 * mem_kind  PB_MEM_HOST: data pointers are host memory (pinned memory makes the copies asynchronous);
 *           PB_MEM_DEVICE: device pointers on the context's device, results stay on the device.
 * stream    cudaStream_t, or NULL: with PB_MEM_HOST the context's own (non-blocking) stream, with PB_MEM_DEVICE the
 *           legacy default stream (so the call is ordered after the producers of the caller's buffers).  The call
 *           returns after the stream has been synchronised (n_clusters_out is valid on return).
 *           A context owns ONE scratch arena: a call on a different stream than the previous call first waits for
 *           that stream (the asynchronous device-only entry points below may still be running on it).
 */
int pb_binary_cluster(pb_ctx *ctx, const float *x, const float *y, const float *z, const float *xo,
                      const float *yo, const float *zo, const int32_t *sem, const int32_t *seg_counts,
                      int32_t n_seg, int64_t n_pts, const float *radius, const int32_t *min_pts,
                      float para_f, int assign_lp, int32_t *cluster_id, int32_t *cluster_num,
                      int32_t *degree, float *center, int64_t center_cap, int32_t *clt_sem,
                      int64_t clt_sem_cap, int64_t *n_clusters_out, int mem_kind, void *stream);

/*
 * Many reference calls in one launch sequence.  Call c owns call_seg_counts[c] consecutive segments;
 * cluster ids restart at 0 for every call (exactly what n_calls separate pb_binary_cluster calls
 * return); center / clt_sem are the per-call results concatenated in call order;
 * call_clusters[c] (host, may be NULL) receives the cluster count of call c.
 */
int pb_binary_cluster_batched(pb_ctx *ctx, const float *x, const float *y, const float *z,
                              const float *xo, const float *yo, const float *zo, const int32_t *sem,
                              const int32_t *seg_counts, int32_t n_seg, const int32_t *call_seg_counts,
                              int32_t n_calls, int64_t n_pts, const float *radius, const int32_t *min_pts,
                              float para_f, int assign_lp, int32_t *cluster_id, int32_t *cluster_num,
                              int32_t *degree, float *center, int64_t center_cap, int32_t *clt_sem,
                              int64_t clt_sem_cap, int64_t *n_clusters_out, int64_t *call_clusters,
                              int mem_kind, void *stream);

/* Large batched calls are split into chunks of about `points` points (runs of whole calls) that alternate
 * over two internal streams, so the latency-bound tail kernels and the host copies of one chunk overlap the
 * neighbour-count kernel of the next.  0 = automatic (3 M points once a call has >= 6 M), < 0 or huge = off.
 * Results do not depend on the chunking. */
void pb_set_chunk_points(pb_ctx *ctx, int64_t points);

/* Calls of ONE reference call with at most 32 segments whose adjacency bitmaps fit 12 MB (a segment of up to ~10 k points:
 * the per-class calls of pbnet_ops.cluster) run as ONE cooperative launch of the small-call kernel (pb_small.cuh) instead of
 * the cell-grid pipeline.  mode 0 = never, 1 / -1 = whenever eligible (default).  Results are identical either way. */
void pb_set_small_calls(pb_ctx *ctx, int mode);

/* Device self-test of the centre kernel's division (k_centres replays the reference's running mean
 * M += (x - M) / n, lib/PB_lib/src/pbnet/binary_cuda_functions.cu:237-239, with a reciprocal-based
 * correctly rounded quotient): compares it bit for bit with div.rn.f32 on n_samples pseudo-random and
 * adversarial (dividend, count) pairs and returns the number of mismatches (expected 0). */
int pb_selftest_division(pb_ctx *ctx, int64_t n_samples, int64_t seed, int64_t *mismatches);

/* Optional per-stage device timings (ms, CUDA events) of the last call when enabled; stage names via
 * pb_stage_name.  Off by default (events add launch overhead). */
void pb_set_profiling(pb_ctx *ctx, int on);
int pb_stage_count(void);
const char *pb_stage_name(int i);
float pb_stage_ms(const pb_ctx *ctx, int i);
/* counters of the last call: [0] pair tests issued by the degree kernel, [1] sum of degrees,
 * [2] HP count, [3] LP-assignment queries, [4] occupied grid cells, [5] raw clusters before the filter,
 * [6] chunks, [7] 1 if the mixed-class kernels were needed, [8] occupied coarse cells, [9] 1 if the one-launch
 * small-call kernel ran, [10] the share of [0] that was tested one-sided (pairs inside one query group); the other
 * tests of [0] are symmetric: one test settles both directions of a pair */
int64_t pb_counter(const pb_ctx *ctx, int i);

/* ------------------------------------------------------------------------------------------------
 * voxelize / devoxelize (scatter-gather around the sparse-conv backbone).  The reference calls
 * MinkowskiEngine for these (un-vendored): ME.utils.sparse_quantize at
 * datasets/scannetv2/dataset_preprocess.py:269-274,348-353; ME.SparseTensor(...).inverse_mapping at
 * network/PBNet.py:236-247,261-271; the devoxelize gathers X_v[v2p] at network/PBNet.py:130-134,250.
 *
 * pb_voxelize   <-  ME.utils.sparse_quantize(coords, quantization_size, return_index, return_inverse)
 *   coords      [n, stride] fp32 (coord_f64 = 0) or fp64 (= 1), stride 3 (x,y,z) or 4 (batch,x,y,z when
 *               has_batch_col); optional int32 batch[n] overrides the batch column
 *   voxel_size  voxel = floor(coord / voxel_size); <= 0 means coords are already in voxel units (floor only)
 *   vcoords     [V,4] int32 (batch,x,y,z), lexicographic order       index   [V] representative point (smallest index)
 *   inverse     [n] point -> voxel                                    order / vox_start  CSR voxel -> points
 *   cap         capacity (in voxels) of vcoords / index / vox_start-1; device outputs need cap >= n
 * pb_voxel_rows <-  feats[index] (mode 0), SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE (mode 1,
 *               commented-out option at network/PBNet.py:243,268), or the autograd scatter-add of the
 *               devoxelize gather (mode 2): out[v,:] = reduce over the points of voxel v, ascending order
 * pb_devoxelize <-  out[p,:] = vfeat[inverse[p],:]
 * Parity note: MinkowskiEngine is not available here; voxel ORDER is implementation-defined in ME, so
 * results are compared with ME's documented contract modulo a permutation of voxels.
 */
int pb_voxelize(pb_ctx *ctx, const void *coords, int coord_f64, int stride, int has_batch_col, const int32_t *batch,
                int64_t n, double voxel_size, int32_t *vcoords, int64_t *index, int64_t *inverse, int32_t *order,
                int32_t *vox_start, int64_t cap, int64_t *n_voxels_out, int mem_kind, void *stream);
int pb_voxel_rows(pb_ctx *ctx, const float *rows, int64_t n_rows, int C, const int32_t *order, const int32_t *vox_start,
                  int64_t V, int mode, float *out, int mem_kind, void *stream);
int pb_devoxelize(pb_ctx *ctx, const float *vfeat, int64_t V, int C, const int64_t *inverse, int64_t n, float *out,
                  int mem_kind, void *stream);

/* ------------------------------------------------------------------------------------------------
 * proposal-vs-instance IoU and mask labels (training-time scoring; device pointers only, asynchronous on
 * `stream`).  Replace
 *   void get_iou(at::Tensor x5, int nInstance, int nProposal)                    lib/PB_lib/src/iou/get_iou.h:16,
 *        bound at lib/PB_lib/src/PB_lib_api.cpp:8, called from lib/PB_lib/torch_io/pbnet_ops.py:101 (PBNet.py:410)
 *   void cal_iou_and_masklabel(at::Tensor x5, int, int, at::Tensor, at::Tensor, int mode)
 *        lib/PB_lib/src/cal_iou_and_masklabel/cal_iou_and_masklabel.h, bound at PB_lib_api.cpp:9
 * proposals_idx i32[sumNPoint], proposals_offset i32[nProposal+1], instance_labels i64[N] (-100 = ignore),
 * instance_pointnum i32[nInstance]; proposals_iou f32[nProposal*nInstance] (out);
 * mode 0: IoU of the whole proposal, mode 1: of its points with mask_scores_sigmoid > 0.5;
 * mask_label f32[sumNPoint] (in/out, caller pre-fills -1): 1/0 where the best IoU exceeds 0.5.
 */
int pb_get_iou(pb_ctx *ctx, const int32_t *proposals_idx, const int32_t *proposals_offset,
               const int64_t *instance_labels, const int32_t *instance_pointnum, float *proposals_iou,
               int32_t nInstance, int32_t nProposal, void *stream);
int pb_cal_iou_and_masklabel(pb_ctx *ctx, const int32_t *proposals_idx, const int32_t *proposals_offset,
                             const int64_t *instance_labels, const int32_t *instance_pointnum, float *proposals_iou,
                             int32_t nInstance, int32_t nProposal, const float *mask_scores_sigmoid,
                             float *mask_label, int mode, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side front end of the fused class loop (device pointers only).  Replaces, for ALL classes at once, the Python of
 *   network/PBNet.py:151-165   per class: nonzero + sort of the class's points, the `count < count_mean[c]*0.05` skip,
 *                              gathers of xyz / offset / class, the fp32 add `ins_orig + ins_offset`
 *   network/PBNet.py:282-287   get_batch_offset: points per scene copy
 * xyz / offset f32[n_pts][3], sem i64[n_pts] (argmax class, 0..19), batch i32 or i64 [n_pts] (scene copy, 0..copies-1),
 * skip_thresh20 = fp32(count_mean[c]*0.05) (HOST, entries 0/1 unused).  Outputs (device, capacity n_pts): SoA shifted /
 * original coordinates, class and original point index of the kept points in class-major, copy-major, ascending-index
 * order — the layout pb_binary_cluster_batched consumes (one call per kept class, `copies` segments each).  Host outputs:
 * class_keep20[c], seg_counts[(kept classes) x copies], n_kept.  One host synchronisation. */
int pb_group_front(pb_ctx *ctx, const float *xyz, const float *offset, const int64_t *sem, const void *batch, int batch_is64,
                   int64_t n_pts, int32_t copies, const float *skip_thresh20, float *x, float *y, float *z, float *xo, float *yo,
                   float *zo, int32_t *sem32, int64_t *point_index, int32_t *class_keep20, int32_t *seg_counts,
                   int64_t *n_kept_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * "local scene" proposal lists and get_proposal (device pointers only; the device-resident continuation of
 * pb_binary_cluster_batched).  Replace the Python loops of
 *   network/PBNet.py:180-234   per cluster: member list (torch.nonzero over the segment), cdist/topk of the centres,
 *                              members of the para_k nearest clusters appended with weights peak_v[k] for clusters
 *                              larger than count_mean[sem]*0.2; training: torch.mode of the instance labels, -100 skip,
 *                              ground-truth mask
 *   network/PBNet.py:317-346   get_proposal: mask score threshold, renumbering of the non-empty proposals
 *
 * pb_local_scenes_plan   cluster_id i32[n_pts] / cluster_num i32[n_seg] / center f32[3*n_clusters] exactly as written by
 *                        pb_binary_cluster_batched on the device (ids restart in every call); seg_counts, call_seg_counts,
 *                        call_sem (class of every call), big_thresh20 (= fp32(count_mean[c]*0.2)), k_max20 are HOST tables;
 *                        ins_label i64[n_pts] (device, or NULL = inference).  Returns the number of proposals and of list
 *                        entries (one host synchronisation) and keeps the plan in the context workspace.
 * pb_local_scenes_fill   must directly follow _plan: writes prop_offsets i64[P+1], prop_cluster i32[P] (global cluster
 *                        index, optional), prop_index i64[E] (position of the listed point in the n_pts input, or
 *                        point_map[position] when point_map i64[n_pts] is given), prop_dpn f32[E] (weights), prop_gt
 *                        i32[E] (training: 1 / 0 / -1, optional), prop_id i32[E] (proposal of every entry, optional).
 * pb_get_proposal        prop_offsets i64[P+1], point_idx i64[E], mask_score f32[E] -> proposals_idx i64[<=E][2],
 *                        proposals_offset i64[<=P+1], cluster_id_v i64[<=P], proposals_ms f32[<=E] (capacities E / P+1 / P);
 *                        returns the kept entry count and the number of non-empty proposals.
 */
int pb_local_scenes_plan(pb_ctx *ctx, const int32_t *cluster_id, const int32_t *seg_counts, int32_t n_seg,
                         const int32_t *call_seg_counts, const int32_t *call_sem, int32_t n_calls, int64_t n_pts,
                         const int32_t *cluster_num, const float *center, int64_t n_clusters, const float *big_thresh20,
                         const int32_t *k_max20, const int64_t *ins_label, int64_t *n_proposals_out, int64_t *n_entries_out,
                         void *stream);
int pb_local_scenes_fill(pb_ctx *ctx, const int64_t *point_map, int64_t *prop_offsets, int32_t *prop_cluster,
                         int64_t *prop_index, float *prop_dpn, int32_t *prop_gt, int32_t *prop_id, void *stream);
int pb_get_proposal(pb_ctx *ctx, const int64_t *prop_offsets, int64_t n_proposals, const int64_t *point_idx,
                    const float *mask_score, int64_t n_entries, float thd, int64_t *proposals_idx, int64_t *proposals_offset,
                    int64_t *cluster_id_v, float *proposals_ms, int64_t *n_kept_out, int64_t *n_nonempty_out, void *stream);
/* Feature rows of the proposal lists (network/PBNet.py:195,231 torch.cat of gathered rows): out f32[E][C+2] =
 * [ point_feat[index[e]] (C channels) | sem_score[index[e]][prop_sem[prop_id[e]]] | dpn[e] ].  Device pointers, asynchronous. */
int pb_scene_features(pb_ctx *ctx, const float *point_feat, int32_t C, const float *sem_score, int32_t n_cls,
                      const int64_t *index, const int32_t *prop_id, const int32_t *prop_sem, const float *dpn,
                      int64_t n_entries, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation post-processing of the proposals (device pointers only).  Replaces the dense-matrix / CPU code of
 *   eval_map.py:63-121          class of a proposal, fold of the rotated scene copies (index % (point_num/3)), score and
 *                               point-count thresholds, cross IoU (dense fp32 mm), per-point labels, rebuilt clusters
 *   tools/mIOU.py:77-87         non_max_suppression (greedy, CPU numpy)
 *   tools/getins.py:72-98       align_superpoint_label (scipy coo_matrix on the CPU)
 * in   proposals_idx i64[n_entries][2], proposals_offset i64[n_proposals+1], clt_score f32[n_proposals] (get_proposal /
 *      the score head), pred_sem i64[point_num], superpoint i64[point_num/copies] with ids in [0, n_superpoints),
 *      sem_table i64[n_table] (HOST; eval_map.py:32 semantic_label_idx)
 * out  label i32[point_num/copies]: final cluster of every point or -100 (clusters are disjoint after the alignment, so this
 *      IS the reference's `clusters` matrix: clusters[c] = (label == c)); cluster_scores f32, cluster_sem i64,
 *      cluster_proposal i32 (proposal index of every final cluster), capacity `cap` >= number of proposals passing the
 *      thresholds; *n_clusters_out.  Two host synchronisations.
 */
int pb_eval_postprocess(pb_ctx *ctx, const int64_t *proposals_idx, int64_t n_entries, const int64_t *proposals_offset,
                        int64_t n_proposals, const float *clt_score, const int64_t *pred_sem, int64_t point_num, int32_t copies,
                        const int64_t *superpoint, int64_t n_superpoints, const int64_t *sem_table, int32_t n_table,
                        float score_thresh, int32_t npoint_thresh, float nms_thresh, int32_t *label, float *cluster_scores,
                        int64_t *cluster_sem, int32_t *cluster_proposal, int64_t cap, int64_t *n_clusters_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Mesh vertex normals — the fourth op of the reference's extension module.  Replaces
 *   void cal_normal_line(at::Tensor xyz, at::Tensor f_abc, at::Tensor n_xyz, int num_vtx, int num_face)
 *        lib/PB_lib/src/normal/cal_normal.h:10, cal_normal.cu:114-160, bound at lib/PB_lib/src/PB_lib_api.cpp:10,
 *        called from lib/PB_lib/torch_io/pbnet_ops.py:163 (offline preprocessing, datasets/scannetv2/decode_scannet.py)
 * xyz f32[num_vtx][3], face i32[>=num_face][3] (only the first num_face faces take part), normal_xyz f32[num_vtx][3] (out):
 * area-weighted mean of the normals of the faces listing the vertex, summed in ascending face order with the reference's
 * fp32 arithmetic; vertices without faces get (0,0,1).  Host or device pointers; synchronous.
 */
int pb_cal_normal_line(pb_ctx *ctx, const float *xyz, const int32_t *face, float *normal_xyz, int32_t num_vtx, int32_t num_face,
                       int mem_kind, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PBNET_B200_H */
