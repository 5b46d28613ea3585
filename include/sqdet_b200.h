/*
 * sqdet_b200.h -- C ABI of libsqdet_b200.so: SqueezeDet's post-backbone detection path as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference (hazenai/SqueezeDet-PyTorch) has NO FFI layer: its "operator API" for this path
 * is the Python class surface of src/model/squeezedet.py and src/engine/detector.py
 * (SURVEY.md 8b).  Each entry point below therefore names the reference function(s) whose device
 * work it replaces; the Python mirrors in squeezedet-pytorch_b200/ bind them through ctypes and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Calls only ENQUEUE work.
 *   - return 0 on success, <0 for a bad argument (SQD_E_*), >0 for a cudaError_t.  A message
 *     for the last failure on the calling thread is available from sqd_last_error().
 *   - the library never allocates device memory: scratch is caller-provided and sized by the
 *     matching *_workspace_bytes() query.  No global mutable state; safe to call from several
 *     host threads (one per device, as torch's DataParallel does -- parallel_apply.py).
 *   - all tensors are dense, row-major, with the shapes given per function.
 *
 * Anchor index convention (src/model/squeezedet.py:85-87, src/utils/boxes.py:49-67):
 *   a = (y*grid_w + x)*K + k ; per-anchor fields interleaved [cls_0..cls_{C-1} | conf | dx dy dw dh].
 */
#ifndef SQDET_B200_H
#define SQDET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQD_ABI_VERSION 4

#if defined(__GNUC__)
#define SQD_API __attribute__((visibility("default")))
#else
#define SQD_API
#endif

#define SQD_OK 0
#define SQD_E_NULL (-1)      /* required pointer is NULL                     */
#define SQD_E_SHAPE (-2)     /* unsupported / inconsistent shape             */
#define SQD_E_WORKSPACE (-3) /* workspace too small or misaligned            */
#define SQD_E_ALIGN (-4)     /* pointer not aligned as required (16 bytes)   */
#define SQD_E_UNSUPPORTED (-5)
#define SQD_E_DRIVER (-6)    /* CUDA driver entry point unavailable          */

#define SQD_MAX_CLASSES 32
#define SQD_MAX_TOPK 1024
#define SQD_MAX_GT 256

/* feature-map memory layouts accepted by the ConvDet head */
#define SQD_LAYOUT_NCHW 0
#define SQD_LAYOUT_NHWC 1 /* torch channels_last: logical NCHW, physical NHWC */
#define SQD_LAYOUT_SPLIT_NHWC 2 /* [max|x| per (image, 64-channel block): B*(Cin/64) x u32, padded to 256 B][x1 plane][x2 plane],
                                   planes (B,gh,gw,Cin) fp16: output of sqd_convdet_split_features, consumed by the
                                   tcgen05 algorithm only */

/* ConvDet algorithms */
#define SQD_CONV_TCGEN05_F16X3 0 /* production: tcgen05.mma cta_group::2 kind::f16 on CTA pairs (M = 256), two-term fp16
                                    split of power-of-two scaled operands (3 products = fp32-level accuracy),
                                    fp32 TMEM accumulators drained in chunks, each CTA holds half of the weight tile */
#define SQD_CONV_SIMT_FP32 1     /* CUDA-core fp32 FMA implicit GEMM (validation yardstick)            */
#define SQD_CONV_TCGEN05_F16X3_1CTA 2 /* same arithmetic, one CTA per tile (cta_group::1); kept as an on-device
                                    cross-check of the pair kernel                                          */

SQD_API int sqd_abi_version(void);
SQD_API const char *sqd_last_error(void);

/* Developer options: alternative routes through the library (each cross-checked bit for bit against the default route
 * by tests/) and tuning knobs, named like the environment variable that seeds them -- e.g. "SQD_SPLIT_TWO_PASS",
 * "SQD_MATCH_SEQUENTIAL", "SQD_NO_PDL", "SQD_HEAD_ONE_KERNEL".  The table is filled from the environment ONCE, on first use;
 * afterwards it changes only through sqd_set_option().  No reference counterpart (the reference has no such switches);
 * unknown names return SQD_E_UNSUPPORTED, a NULL name SQD_E_NULL. */
SQD_API int sqd_set_option(const char *name, int value);
SQD_API int sqd_get_option(const char *name, int *value);

/* ---------------------------------------------------------------------------------------------
 * a1  ConvDet head: 3x3 conv, pad 1, stride 1, + bias, written directly in the (B, A, C+5)
 *     layout.  Replaces SqueezeDetBase.convdet + permute/contiguous/view,
 *     src/model/squeezedet.py:73-75,83-87 (cuDNN conv + ATen copy in the reference).
 *
 *   d_feat   (B, Cin, gh, gw) fp32 in `layout` (NCHW or NHWC-physical), 16-byte aligned
 *   d_weight (Cout, Cin, 3, 3) fp32 -- the reference's state-dict tensor base.convdet.weight
 *   d_bias   (Cout) fp32
 *   d_pred   (B, gh*gw*K, C+5) fp32 == (B, gh, gw, Cout) ; Cout = K*(C+5)
 *
 * sqd_convdet_pack_weights() derives the kernel-side weight matrix (tap-major K, two-term fp16 split of
 * the power-of-two scaled weights, N padded to a multiple of 16) ONCE per weight update; derived data,
 * never saved.  Cin must be a multiple of 64, Cout <= 128.
 * ------------------------------------------------------------------------------------------- */
SQD_API size_t sqd_convdet_packed_weight_bytes(int cout, int cin);
SQD_API int sqd_convdet_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, void *stream);
/* max|x| per (image, 64-channel block) + two-term fp16 split of a feature map into NHWC planes (the tcgen05 kernel's
 * A operand); one pass over HBM for NCHW input (a thread-block cluster per slab).  sqd_convdet_forward
 * does this internally for NCHW / NHWC input; a producer that can emit the planes itself (or reuses them)
 * passes them with SQD_LAYOUT_SPLIT_NHWC and skips the pass. */
SQD_API size_t sqd_convdet_split_bytes(int batch, int cin, int gh, int gw);
SQD_API int sqd_convdet_split_features(const float *d_feat, int layout, int batch, int cin, int gh, int gw,
                                       void *d_planes, void *stream);
SQD_API size_t sqd_convdet_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout, int algo);
SQD_API int sqd_convdet_forward(const float *d_feat, int layout, const void *d_packed, const float *d_weight,
                        const float *d_bias, int batch, int cin, int gh, int gw, int cout, float *d_pred,
                        void *d_workspace, size_t workspace_bytes, int algo, void *stream);

/* ---------------------------------------------------------------------------------------------
 * a2-a7  PredictionResolver.forward + SqueezeDet.forward scoring, one pass over pred.
 *     Replaces safe_softmax (modules.py:66-68), log_softmax / sigmoid (squeezedet.py:110-114),
 *     deltas_to_boxes + xywh_to_xyxy + clamp (modules.py:17-45), probs*=conf / argmax / max
 *     (squeezedet.py:200-202).
 *
 *   d_pred     (B, A, C+5) fp32
 *   d_anchors  (A, 4) fp32 xywh  (== cfg.anchors.astype(float32), squeezedet.py:106)
 *   outputs, each may be NULL to skip it:
 *     d_class_ids (B, A) int64 | d_scores (B, A) | d_boxes (B, A, 4) xyxy clamped
 *     d_probs (B, A, C) softmax (NOT multiplied by conf) | d_logp (B, A, C) | d_conf (B, A, 1)
 *     d_deltas (B, A, 4)
 * ------------------------------------------------------------------------------------------- */
SQD_API int sqd_decode_scores(const float *d_pred, const float *d_anchors, int batch, int num_anchors, int num_classes,
                      int input_h, int input_w, int64_t *d_class_ids, float *d_scores, float *d_boxes,
                      float *d_probs, float *d_logp, float *d_conf, float *d_deltas, void *stream);

/* ---------------------------------------------------------------------------------------------
 * a8-a9  Detector.filter for a whole batch: top-k by score (ties: lower anchor index first),
 *     per-class greedy NMS with torchvision.ops.nms semantics (float IoU, strict '>' against the
 *     double threshold), strict score > (float)score_thresh.  Replaces src/engine/detector.py:87-122
 *     (torch.argsort + 3 torchvision nms calls + >= 3C+3 host syncs per image).
 *
 *   inputs : d_class_ids (B, A) int64, d_scores (B, A), d_boxes (B, A, 4)   [SqueezeDet.forward dict]
 *   outputs: d_count (B) int32 ; rows [0,count) of (B, top_k[,4]) buffers in the reference's output
 *            order (class ascending, score descending inside a class):
 *            d_out_anchor int32 (kept anchor index), d_out_class int32, d_out_score, d_out_box
 * ------------------------------------------------------------------------------------------- */
SQD_API int sqd_topk_nms(const int64_t *d_class_ids, const float *d_scores, const float *d_boxes, int batch,
                 int num_anchors, int num_classes, int top_k, double nms_thresh, double score_thresh,
                 int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class, float *d_out_score,
                 float *d_out_box, void *stream);

/* Fused a2-a9: pred -> final detections; the dense (B,A) ids/scores/boxes of the SqueezeDet.forward contract
 * are never materialised (reads A*(C+5)*4 bytes per image once).
 *   d_workspace != NULL (sqd_detect_workspace_bytes): two launches -- a streaming scan of pred that appends the
 *     anchors above the score threshold to per-image candidate lists (HBM bound, the whole grid on the whole batch),
 *     then one CTA per image selects / sorts / suppresses / emits.  The form to use for large batches.
 *   d_workspace == NULL: one launch, a thread-block cluster per image does both (no scratch memory).
 * Both forms return identical results. */
SQD_API size_t sqd_detect_workspace_bytes(int batch, int num_anchors);
SQD_API int sqd_detect_from_pred(const float *d_pred, const float *d_anchors, int batch, int num_anchors, int num_classes,
                         int input_h, int input_w, int top_k, double nms_thresh, double score_thresh,
                         int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class, float *d_out_score,
                         float *d_out_box, void *d_workspace, size_t workspace_bytes, void *stream);

/* Fused a1-a9: Fire11 features -> final detections (Detector.detect's device work,
 * src/engine/detector.py:20-31).  Workspace from sqd_head_detect_workspace_bytes().  With the tcgen05 algorithm the
 * GEMM epilogue itself scores the anchors (softmax x sigmoid, argmax) and fills the candidate lists, so the filter
 * that follows reads a few hundred keys per image instead of scanning pred. */
SQD_API size_t sqd_head_detect_status_offset(int batch, int gh, int gw, int cout);  /* byte offset of the int32 pipeline status word inside the fused call's workspace */
SQD_API size_t sqd_head_detect_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int layout, int algo);
SQD_API int sqd_head_detect_fused(const float *d_feat, int layout, const void *d_packed, const float *d_weight,
                          const float *d_bias, const float *d_anchors, int batch, int cin, int gh, int gw,
                          int anchors_per_grid, int num_classes, int input_h, int input_w, int top_k,
                          double nms_thresh, double score_thresh, int32_t *d_count, int32_t *d_out_anchor,
                          int32_t *d_out_class, float *d_out_score, float *d_out_box, void *d_workspace,
                          size_t workspace_bytes, int algo, void *stream);

/* Diagnostic twin of sqd_head_detect_fused (tcgen05 algorithm, NCHW / NHWC input): the same kernel sequence with CUDA
 * events recorded on `stream` between the stages.  Synchronises the stream and returns the stage durations in
 * milliseconds in the HOST array h_stage_ms[3]: [0] split pre-pass, [1] ConvDet GEMM (+ score epilogue), [2] filter.
 * Used by bench.py for the per-kernel roofline figures. */
SQD_API size_t sqd_head_detect_profile_workspace_bytes(int batch, int cin, int gh, int gw, int cout);
SQD_API int sqd_head_detect_profile(const float *d_feat, int layout, const void *d_packed, const float *d_bias,
                                    const float *d_anchors, int batch, int cin, int gh, int gw, int anchors_per_grid,
                                    int num_classes, int input_h, int input_w, int top_k, double nms_thresh,
                                    double score_thresh, int32_t *d_count, int32_t *d_out_anchor, int32_t *d_out_class,
                                    float *d_out_score, float *d_out_box, void *d_workspace, size_t workspace_bytes,
                                    void *stream, float *h_stage_ms);

/* Host-buffer form of sqd_head_detect_fused (what Detector.detect does around the model call: batch to device,
 * results back to numpy, src/engine/detector.py:22,37).  h_* are HOST pointers (page-locked for the copies to be
 * asynchronous); the call enqueues, per group of `chunk_images` images, the H2D copy on `copy_stream` and the kernels
 * on `stream`, so the copy of group g+1 overlaps the compute of group g, then the D2H copies of the five output
 * arrays on `stream`.  The caller synchronises `stream` before reading the outputs.  copy_stream NULL = no overlap;
 * chunk_images <= 0 = one group.  The workspace holds the device staging of features and outputs.
 * flags: 0, or SQD_HOST_NO_STAGING_FENCE when the caller guarantees that nothing enqueued earlier on `stream` still
 * reads THIS workspace (a serving loop that alternates two workspaces and has read call i's results before issuing
 * call i+2): the first H2D copy then starts as soon as copy_stream is free instead of after `stream`'s earlier
 * kernels, which keeps the PCIe link busy across calls. */
#define SQD_HOST_NO_STAGING_FENCE 1
SQD_API size_t sqd_head_detect_host_workspace_bytes(int batch, int cin, int gh, int gw, int cout, int top_k, int layout,
                                                    int algo, int chunk_images);
/* Layout of ONE host block mirroring the device-side result block of sqd_head_detect_host: offsets5 = byte offsets of
 * (count, anchor, class, score, box), *total = its size.  Passing views of such a (pinned) block as the five host output
 * pointers lets the call return its detections with a single device-to-host copy.  Host-side helper, no reference
 * counterpart (the reference copies every kept tensor separately: src/engine/detector.py:37). */
SQD_API int sqd_head_detect_host_result_layout(int batch, int top_k, size_t *offsets5, size_t *total);

SQD_API int sqd_head_detect_host(const float *h_feat, int layout, const void *d_packed, const float *d_weight,
                                 const float *d_bias, const float *d_anchors, int batch, int cin, int gh, int gw,
                                 int anchors_per_grid, int num_classes, int input_h, int input_w, int top_k,
                                 double nms_thresh, double score_thresh, int32_t *h_count, int32_t *h_out_anchor,
                                 int32_t *h_out_class, float *h_out_score, float *h_out_box, void *d_workspace,
                                 size_t workspace_bytes, int algo, int chunk_images, void *stream, void *copy_stream,
                                 int flags);

/* ---------------------------------------------------------------------------------------------
 * a11-a12  compute_deltas: greedy sequential anchor<->ground-truth matching in float64.
 *     Replaces numpy compute_overlaps + compute_deltas, src/utils/boxes.py:70-135.
 *     Tie policy: lowest anchor index among equal IoU / equal distance (== the reference with a
 *     stable argsort; its default unstable sort is implementation defined).
 *
 *   d_gt_boxes (B, gmax, 4) fp32 xyxy, rows >= d_gt_count[b] ignored ; d_gt_count (B) int32
 *   d_anchors64 (A, 4) float64 xywh -- the numpy-generated table (boxes.py:37-67)
 *   outputs: d_anchor_idx (B, gmax) int32 (-1 in unused rows), d_deltas (B, gmax, 4) fp32
 * a13  prepare_annotations: dense target (B, A, C+9) = [mask | box | deltas | one-hot]
 *     Replaces src/datasets/base.py:61-76.  d_gt_classes (B, gmax) int32.
 * ------------------------------------------------------------------------------------------- */
SQD_API int sqd_match_anchors(const float *d_gt_boxes, const int32_t *d_gt_count, int batch, int gmax,
                      const double *d_anchors64, int num_anchors, int32_t *d_anchor_idx, float *d_deltas,
                      void *stream);
SQD_API int sqd_build_targets(const float *d_gt_boxes, const int32_t *d_gt_classes, const int32_t *d_gt_count,
                      const int32_t *d_anchor_idx, const float *d_deltas, int batch, int gmax, int num_anchors,
                      int num_classes, float *d_gt_dense, void *stream);

/* ---------------------------------------------------------------------------------------------
 * a14-a16  Loss.forward and its backward.  Replaces PredictionResolver(log_softmax) + torch
 *     compute_overlaps (modules.py:48-63) + the four loss sums (squeezedet.py:133-174) and the
 *     autograd graph behind them.
 *
 *   d_pred (B, A, C+5), d_gt (B, A, C+9) [mask | gt box xyxy | gt deltas | one-hot], d_anchors (A,4) fp32
 *   weights[4] (HOST pointer) = {class, positive_score, negative_score, bbox}
 *   d_losses (B, 4) = per image {class, positive_score, negative_score, bbox}  (loss = their sum)
 *   d_grad_loss (B, 4) upstream d(objective)/d(d_losses[b][t]) or NULL (= all ones) ;
 *   d_dpred (B, A, C+5) or NULL (forward only)
 * ------------------------------------------------------------------------------------------- */
SQD_API size_t sqd_loss_workspace_bytes(int batch, int num_anchors);
SQD_API int sqd_loss_fwd_bwd(const float *d_pred, const float *d_gt, const float *d_anchors, int batch, int num_anchors,
                     int num_classes, int input_h, int input_w, const float *weights, const float *d_grad_loss,
                     float *d_losses, float *d_dpred, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * 8(f) rank 1  boxes_postprocess for the kept detections (src/utils/boxes.py:138-168), in place on
 *     the (B, top_k, 4) output of the filter: /scale, -padding, +crops, flip, +drifts in the
 *     reference's order.  d_meta (B, 10) fp32, one record per image:
 *     [scale_y, scale_x, pad_top, pad_left, crop_top, crop_left, flip_width (<=0: not flipped),
 *      drift_y, drift_x, 0] ; keys absent from image_meta are passed as identity (1 / 0).
 * ------------------------------------------------------------------------------------------- */
SQD_API int sqd_boxes_postprocess(float *d_boxes, const int32_t *d_count, const float *d_meta, int batch, int top_k,
                          void *stream);

/* 8(f) rank 1, second half: result packing.  One (B, top_k, 6) fp32 array [class, score, x1, y1, x2, y2] (rows >= count:
 * class -1, zeros) so that a batch crosses to the host with ONE copy (+ the counts); d_meta (B,10) as above applies
 * boxes_postprocess on the way (NULL: boxes copied as they are).  Inputs are not modified.
 * Replaces the per-image .cpu() calls and the host boxes_postprocess of src/engine/detector.py:37-40. */
SQD_API int sqd_pack_results(const int32_t *d_count, const int32_t *d_class, const float *d_score, const float *d_box,
                             const float *d_meta, int batch, int top_k, float *d_packed, void *stream);

/* 8(f) rank 3  KITTI result writer (HOST function; src/datasets/kitti.py:78-97).  Formats the packed results of a
 * batch as the evaluator's text lines
 *     "{class} -1 -1 0 {x1:.2f} {y1:.2f} {x2:.2f} {y2:.2f} 0 0 0 0 0 0 0 {score:.3f}\n"
 * into h_out; h_offsets (batch+1 entries, may be NULL) delimits the block of each image (an image that kept nothing
 * has an empty block, like the reference's empty file).  class_names_lower: num_classes C strings.
 * Returns the number of bytes needed (call with cap = 0 to size the buffer) or < 0 on a bad argument. */
SQD_API long long sqd_format_kitti(const float *h_packed, const int32_t *h_count, int batch, int top_k,
                                   const char *const *class_names_lower, int num_classes, char *h_out, size_t cap,
                                   long long *h_offsets);

/* 8(f) rank 4  input pre-processing: whiten, bilinear resize, HWC -> CHW in one pass.
 *     Replaces whiten + resize of src/utils/image.py:9-19,77-88 (numpy + cv2.resize INTER_LINEAR on float32, whose
 *     coordinate / weight arithmetic is reproduced) and the transpose of src/datasets/base.py:33 for eval-mode inputs
 *     (no drift / flip).
 *   d_images (B, src_h, src_w, 3) uint8 (dtype 0) or float32 (dtype 1), RGB interleaved as loaded
 *   mean3 / std3: HOST pointers to the 3 channel means / stds (kitti.py:17-18)
 *   d_out (B, 3, dst_h, dst_w) float32 ; image_meta['scales'] = (dst_h/src_h, dst_w/src_w) is the caller's to record */
SQD_API int sqd_preprocess(const void *d_images, int dtype, int batch, int src_h, int src_w, const float *mean3,
                           const float *std3, int dst_h, int dst_w, float *d_out, void *stream);

/* 8(f) rank 2 (first half)  ConvDet backward: gradient with respect to the Fire11 features and the bias.
 *     Replaces autograd through nn.Conv2d (src/model/squeezedet.py:73-75; cuDNN dgrad in the reference).  The input
 *     gradient is the forward implicit GEMM with roles swapped (3x3 convolution of d_gpred with flipped, transposed
 *     weights), run on the same tcgen05 f16x3 kernel in slabs of 128 feature channels.
 *   d_gpred      (B, gh, gw, Cout) fp32 -- gradient of pred in the library's (B, A, C+5) layout
 *   d_gfeat_nhwc (B, gh, gw, Cin)  fp32 -- channels_last memory of the logical (B, Cin, gh, gw) gradient; Cin % 128 == 0
 *   sqd_convdet_dgrad_pack_weights derives the flipped / transposed weight planes once per weight update.
 *   sqd_convdet_wgrad: d_gweight (Cout, Cin, 3, 3) from the NCHW fp32 features and d_gpred -- an fp32 CUDA-core implicit
 *   GEMM split over the pixel axis with a fixed-order reduction (deterministic); any Cout (80 output channels per launch).
 *   The tensor-core form is sqd_convdet_wgrad_tc below (Cout <= 80, even grid width); this one is the yardstick and the
 *   route for every other shape. */
SQD_API size_t sqd_convdet_dgrad_packed_bytes(int cout, int cin);
SQD_API int sqd_convdet_dgrad_pack_weights(const float *d_weight, int cout, int cin, void *d_packed, void *stream);
SQD_API size_t sqd_convdet_dgrad_workspace_bytes(int batch, int cin, int gh, int gw, int cout);
SQD_API int sqd_convdet_dgrad(const float *d_gpred, const void *d_dgrad_packed, int batch, int cin, int gh, int gw, int cout,
                              float *d_gfeat_nhwc, void *d_workspace, size_t workspace_bytes, void *stream);
SQD_API size_t sqd_convdet_wgrad_workspace_bytes(int batch, int cin, int gh, int gw, int cout);
SQD_API int sqd_convdet_wgrad(const float *d_feat_nchw, const float *d_gpred, int batch, int cin, int gh, int gw, int cout,
                              float *d_gweight, void *d_workspace, size_t workspace_bytes, void *stream);
/* The same weight gradient on tcgen05 / TMEM (f16x3, fp32-grade): pixel-major fp16 planes of the features and of the
 * transposed d_gpred are built in the workspace, the contraction over pixels runs as (128 channels x 80 outputs x 64
 * pixels) MMAs with one TMEM accumulation chunk per image, slices of the batch are reduced in a fixed order.
 * Cin % 64 == 0, Cout <= 80.  sqd_convdet_wgrad_tc_status: 0 if the last call drained cleanly (synchronises). */
SQD_API size_t sqd_convdet_wgrad_tc_workspace_bytes(int batch, int cin, int gh, int gw, int cout);
SQD_API int sqd_convdet_wgrad_tc(const float *d_feat_nchw, const float *d_gpred, int batch, int cin, int gh, int gw, int cout,
                                 float *d_gweight, void *d_workspace, size_t workspace_bytes, void *stream);
SQD_API int sqd_convdet_wgrad_tc_status(const void *d_workspace, void *stream);
SQD_API int sqd_convdet_bias_grad(const float *d_gpred, int batch, int gh, int gw, int cout, float *d_gbias, void *stream);

/* Debug aid: synchronise `stream`, return 0 if the last tcgen05 ConvDet launch on this workspace
 * drained cleanly, else the role (1 TMA producer, 2 MMA issuer, 3 epilogue) whose bounded mbarrier
 * wait timed out.  The kernels never spin forever. */
SQD_API int sqd_convdet_status(const void *d_workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SQDET_B200_H */
