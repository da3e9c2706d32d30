/*
 * b200det.h -- C ABI of libb200det.so: the B200 (sm_100a) implementation of SimpleAICV's
 * dense-detection loss + decode hot path.
 *
 * Nothing like this exists in the reference (the path is pure torch / NumPy there); every
 * entry point names the reference code it replaces (paths relative to the reference repo).
 * The Python classes in `b200det.losses` / `b200det.decode` bind these with ctypes and keep
 * the reference's constructor / __call__ signatures.
 *
 * Conventions
 *  - plain C: POD structs, raw device pointers, sizes; no torch / C++ types.
 *  - the caller owns every buffer (including `workspace`); the library allocates nothing
 *    on the device (the 4 KB peer-exchange buffers and the helper stream / events of
 *    b200det_loss_forward_overlap are created for, and owned by, the caller through explicit
 *    create / destroy calls).  Process-wide mutable state is limited to: a launch counter, the
 *    optional event profiler behind b200det_profile* (off by default; a mutex-guarded record list),
 *    the b200det_select_stamps debug pointer, and per-device "shared-memory limit raised" bits.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *    and never synchronises the host.
 *  - return value: 0 = ok, negative = B200DET_E* argument error, positive = cudaError_t of
 *    the failed launch.  Nothing throws or aborts.
 *  - head outputs are channels-last, one contiguous allocation per pyramid level, exactly
 *    as the reference heads emit them (models/retinanet.py:72-83, models/fcos.py:71-80):
 *      Retina: cls[l] float32 [B,H,W,A,C] probabilities, reg[l] [B,H,W,A,4]
 *      FCOS  : cls[l] float32 [B,H,W,C] probabilities, reg[l] [B,H,W,4], ctr[l] [B,H,W,1]
 *    `reg` may be float32, float16 or bfloat16 (autocast); it is upcast on load.
 *  - "rows": one row = one anchor (Retina) or point (FCOS).  Per image there are
 *    N = sum_l H_l*W_l*per_loc rows.  Per-row device arrays produced/consumed by this
 *    library (labels, matched, keys, classes, targets) are LEVEL-MAJOR: level l occupies
 *    rows [B*off_l, B*off_{l+1}) laid out [B, n_l], off_l = rows of one image before level l
 *    (i.e. the same order as the head tensors themselves, so streaming kernels index them
 *    without division).  b200det_rows_to_image_major() converts to the reference's [B,N].
 */
#ifndef B200DET_H_
#define B200DET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DET_ABI_VERSION 1
#define B200DET_MAX_LEVELS 8
#define B200DET_MAX_PER_LOC 16
#define B200DET_MAX_GT 2048   /* annotation rows per image staged in shared memory */
#define B200DET_MAX_TOPN 2048
#define B200DET_MAX_PEERS 16  /* ranks of one NVLink domain in a peer exchange */

/* error codes (negative) */
#define B200DET_EINVAL (-1)    /* null pointer / bad enum / bad size */
#define B200DET_ERANGE (-2)    /* a size exceeds a compiled limit (levels, GT rows, topn, 2^31 rows) */
#define B200DET_EWORKSPACE (-3) /* workspace too small */
#define B200DET_EALIGN (-4)    /* pointer not aligned as required */

/* dtype of the regression head */
#define B200DET_F32 0
#define B200DET_F16 1
#define B200DET_BF16 2
/* OR into a half-precision `reg_dtype`: exp() of a regression value is ROUNDED to that precision
 * before it is used, as the reference's eager half arithmetic does (torch.exp on a float16 tensor,
 * losses.py:417-426 / :568; np.exp on a float16 array, decode.py:257-268 / :356: the result of exp
 * is float16 and only the following multiply / subtract promote to float32).  Without the flag exp
 * runs on the upcast value -- what the reference computes under CUDA autocast, where exp is on the
 * float32 list (tools/scripts.py:886-893 calls the criterion inside `with autocast()`). */
#define B200DET_REG_EXP_ROUNDED 0x10

/* box loss: RetinaLoss box_loss_type (losses.py:139-148) / FCOSLoss box_loss_iou_type (:443-448) */
#define B200DET_BOX_NONE 0 /* assignment only */
#define B200DET_BOX_SMOOTHL1 1
#define B200DET_BOX_IOU 2
#define B200DET_BOX_GIOU 3
#define B200DET_BOX_DIOU 4
#define B200DET_BOX_CIOU 5
#define B200DET_BOX_EIOU 6

/* b200det_sparse_losses: OR into `is_fcos` when `ctr` holds the centre-ness head's float32 LOGITS
 * (the logits path) instead of probabilities; forward only */
#define B200DET_FCOS_CTR_LOGITS 2

/* DetNMSMethod nms_type (decode.py:28-32) */
#define B200DET_NMS_PYTHON 0
#define B200DET_NMS_DIOU_PYTHON 1
#define B200DET_NMS_TORCH 2
#define B200DET_NMS_NONE 3 /* DETRDecoder(nms_type=None): top-n only (decode.py:373-385, :453) */

/* box source of b200det_select_decode_nms (`is_fcos` argument) */
#define B200DET_DECODE_ANCHORS 0 /* (tx,ty,tw,th) against generated anchors, int32 truncation */
#define B200DET_DECODE_POINTS 1  /* (l,t,r,b) against generated points, int32 truncation */
#define B200DET_DECODE_BOXES 2   /* reg[l] holds float32 x1,y1,x2,y2 rows, used as they are */

/* class-score source of b200det_query_scores */
#define B200DET_SCORES_PROBS 0   /* probabilities as given */
#define B200DET_SCORES_SIGMOID 1 /* sigmoid(float(logits))            (DINODETRDecoder, decode.py:515-516) */
#define B200DET_SCORES_SOFTMAX 2 /* softmax(logits) over the channels (DETRDecoder, decode.py:391) */

/*
 * Pyramid geometry shared by every call.  Replaces what the reference re-derives on the host
 * every step: RetinaAnchors.__call__ (models/anchor.py:18-86), FCOSPositions.__call__ (:94-130)
 * and the per-point mi / stride tables of FCOSLoss (losses.py:623-632).  Anchors / points are
 * generated in registers from this table; they are never materialised in memory.
 */
typedef struct b200det_geometry {
    int32_t n_levels;                 /* 1..B200DET_MAX_LEVELS */
    int32_t batch;                    /* B (images held by this rank) */
    int32_t per_loc;                  /* anchors per location: 9 (Retina), 1 (FCOS) */
    int32_t num_classes;              /* C */
    int32_t height[B200DET_MAX_LEVELS];
    int32_t width[B200DET_MAX_LEVELS];
    float stride[B200DET_MAX_LEVELS];
    /* Retina: float32 base anchors [level][a] = (x1,y1,x2,y2) centred on 0 (anchor.py:35-57) */
    float base_anchors[B200DET_MAX_LEVELS][B200DET_MAX_PER_LOC][4];
    /* FCOS: regress range (mi) per level, and stride*center_sample_radius (losses.py:690) */
    float mi_lo[B200DET_MAX_LEVELS];
    float mi_hi[B200DET_MAX_LEVELS];
    float radius[B200DET_MAX_LEVELS];
} b200det_geometry;

/* ---- library info ------------------------------------------------------------------- */
int b200det_abi_version(void);
const char *b200det_error_string(int code);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
unsigned long long b200det_launch_count(void);
/*
 * Optional per-kernel timing: while enabled, every kernel launched by the library is bracketed
 * by CUDA events on its stream.  b200det_profile(1) clears and starts, b200det_profile(0) pauses (records are kept),
 * b200det_profile(2) resumes without clearing;
 * b200det_profile_read() waits for the recorded events and returns the summed device time and the
 * number of launches of one kernel id (0 focal_loss, 1 assign, 2 sparse_losses, 3 loss_reduce,
 * 4 loss_finish, 5 score_argmax, 6 select_decode_nms, 7 other, 8 head_tail, 9 logits_sweep; see
 * b200det_kernel_name).
 */
int b200det_profile(int enable);
int b200det_profile_read(int kernel_id, double *total_ms, int *launches);
const char *b200det_kernel_name(int kernel_id);
/* rows per image (N) for a geometry; negative on error */
long long b200det_rows_per_image(const b200det_geometry *geo);

/* blocks the calling host thread until `stream` has drained (cudaStreamSynchronize): the decoders'
 * only host wait, right before they hand out the result arrays */
int b200det_stream_synchronize(void *stream);

/* ---- loss --------------------------------------------------------------------------- */
/* bytes of scratch the loss calls of one step share (per-CTA partials + the matched annotation
 * row of every row, 2 bytes each); the same buffer must be passed to every call of the step */
size_t b200det_loss_workspace_bytes(const b200det_geometry *geo);

/*
 * Anchor<->GT IoU assignment of RetinaLoss: a pure ALU scan, no head tensor is touched.
 * Replaces RetinaLoss.get_batch_anchors_annotations (losses.py:322-388) and the assignment use of
 * IoUMethod (losses.py:33-70).
 *   annotations : device float32 [B, max_gt, 5] = x1,y1,x2,y2,class ; rows with class < 0 ignored
 *   iou_neg/pos : best IoU < iou_neg -> background, >= iou_pos -> positive, between -> ignored:
 *                 0.4 / 0.5 for RetinaLoss (losses.py:361-365), 0.35 / 0.35 for RetinaFaceLoss
 *                 (simpleAICV/face_detection/losses.py:255-259), which shares this kernel
 *   labels      : device int32 [B*N] level-major, out: -1 ignore, 0 background, k = class k-1
 *   matched     : device int32 [B*N] level-major or NULL, out: arg-max GT index in the image's
 *                 filtered GT list (first maximum on ties), -1 if the image has no GT
 *   workspace   : receives the per-CTA positive counts and the work queues of positive rows
 *                 (with their matched annotation row) and ignored rows consumed by
 *                 b200det_sparse_losses.  The call enqueues one 8-byte cudaMemsetAsync.
 * Environment: B200DET_ASSIGN_CTAS_PER_SM (default: no cap) limits the kernel's residency; only
 * useful when the caller overlaps it with the HBM-bound sweep on another stream.
 */
int b200det_retina_assign(const b200det_geometry *geo, const float *annotations, int max_gt,
                          float iou_neg, float iou_pos, int32_t *labels, int32_t *matched,
                          void *workspace, size_t workspace_bytes, void *stream);

/*
 * Point<->GT assignment with centre sampling of FCOSLoss.
 * Replaces FCOSLoss.get_batch_position_annotations (losses.py:612-833).
 *   targets : device float32 [B*N, 6] level-major or NULL, out: l,t,r,b,label,centre-ness
 */
int b200det_fcos_assign(const b200det_geometry *geo, const float *annotations, int max_gt,
                        int use_center_sample, int32_t *labels, int32_t *matched, float *targets,
                        void *workspace, size_t workspace_bytes, void *stream);

/*
 * Losses that only involve the positive (and ignored) rows; must follow the assign call on the
 * same workspace.  Replaces RetinaLoss.compute_batch_box_loss / compute_batch_smoothl1_loss
 * (losses.py:263-320) with snap_annotations_to_txtytwth (:390-409) and snap_txtytwth_to_xyxy
 * (:411-429); FCOSLoss.compute_batch_iou_loss (:550-586) and compute_batch_centerness_loss
 * (:588-610); IoUMethod for all five IoU types (:33-123).
 *   reg       : host array of n_levels device pointers (NULL iff box_loss == NONE)
 *   ctr       : FCOS: host array of n_levels device float32 centre-ness pointers
 *   cls       : NULL, or host array of n_levels device float32 cls pointers.  When given, the
 *               kernel also produces the focal-loss CORRECTIONS of the rows that are not plain
 *               background (target class of positives, every class of ignored rows), so that
 *               b200det_focal_loss can run label-free (labels == NULL) and concurrently on
 *               another stream.  alpha / gamma are the focal parameters for those terms.
 *   reg_grad / ctr_grad : NULL, or per-level device float32 buffers shaped like reg[l] / ctr[l],
 *               ZEROED by the caller; the rows of the positives receive
 *               d(sum of box / centre-ness loss)/d(input), NOT yet divided by the positive count
 */
int b200det_sparse_losses(const b200det_geometry *geo, int is_fcos, const float *annotations,
                          int max_gt, const int32_t *labels, const void *const *reg,
                          int reg_dtype, const void *const *ctr, int box_loss, float beta,
                          const void *const *cls, float alpha, float gamma,
                          void *const *reg_grad, void *const *ctr_grad, void *workspace,
                          size_t workspace_bytes, void *stream);

/*
 * Focal loss over the classification head: one streaming pass over cls (4*N*C bytes/image).
 * Replaces torch.cat + clamp (losses.py:183-198 / :488-494) and compute_batch_focal_loss
 * (losses.py:220-261 / :513-548).  Rows with label < 0 contribute nothing.
 *   labels     : level-major labels from the assign call, or NULL = label-free sweep: every
 *                element is summed as background and b200det_sparse_losses (given `cls`)
 *                supplies the corrections; the sweep is then independent of the assignment and
 *                may overlap it on another stream.  Forward only (cls_grad must be NULL).
 *   cls_grad   : NULL (forward only) or host array of n_levels device float32 buffers; receives
 *                d(cls_loss)/d(cls) already multiplied by grad_scale / sums[0]
 *   sums       : device double[4] written by b200det_loss_reduce: {positives, cls, box, ctr};
 *                only read when cls_grad != NULL (element 0 = positive count, possibly
 *                all-reduced across ranks by the caller)
 */
int b200det_focal_loss(const b200det_geometry *geo, const void *const *cls,
                       const int32_t *labels, float alpha, float gamma, void *const *cls_grad,
                       const double *sums, float grad_scale, void *workspace,
                       size_t workspace_bytes, void *stream);

/*
 * Deterministic (fixed-order, fp64) reduction of the block partials.
 *   which : bit 0 = assignment + sparse partials (positives, box, ctr, focal corrections),
 *           bit 1 = focal sweep partials,
 *           bit 2 = the positive count alone (sums[0]; complete once the assignment has run)
 *   sums  : device double[4] {positives, cls_sum, box_sum, ctr_sum}; only selected fields written
 * Replaces the `.sum()` / `positive_anchors_num` bookkeeping of losses.py:231-259, :277-293.
 */
int b200det_loss_reduce(const b200det_geometry *geo, int which, const void *workspace,
                        size_t workspace_bytes, double *sums, void *stream);

/*
 * losses[i] = weights[i] * sums[1+i] / sums[0]  (0 when sums[0] == 0), i = 0..2, float32.
 * Replaces the final divisions and weightings of losses.py:210-211, :259, :293, :318, :501-503.
 */
int b200det_loss_finish(const double *sums, float w_cls, float w_box, float w_ctr,
                        float *losses, void *stream);

/* b200det_loss_reduce(which = 3) + b200det_loss_finish in one launch */
int b200det_loss_reduce_finish(const b200det_geometry *geo, const void *workspace,
                               size_t workspace_bytes, float w_cls, float w_box, float w_ctr,
                               double *sums, float *losses, void *stream);

/* x[l][i] *= *g * (sums ? weight / sums[0] : 1) for n_levels float32 buffers in one launch
 * (autograd backward: upstream scalar, loss weight, positive count); no-op when the factor is 1 */
int b200det_scale_levels(void *const *ptrs, const long long *counts, int n_levels,
                         const float *g_dev, const double *sums, float weight, void *stream);
/* Autograd backward of the box / centre-ness losses without a pass over the whole gradient tensors:
 * b200det_sparse_losses wrote d(loss sum)/d(input) at the rows of the positives only, and the
 * assignment left those rows in the workspace's positive queue, so
 *   reg_grad[row] *= *g_box * w_box / sums[0],  ctr_grad[row] *= *g_ctr * w_ctr / sums[0]
 * for the queued rows is all there is to do (0 without positives, NaN after a failed exchange).
 * `workspace` must be the one the forward used, untouched since.  g_box / g_ctr NULL: head skipped. */
int b200det_scale_pos_rows(const b200det_geometry *geo, const void *workspace, size_t workspace_bytes,
                           void *const *reg_grad, void *const *ctr_grad, const float *g_box,
                           const float *g_ctr, const double *sums, float w_box, float w_ctr,
                           void *stream);
/* x[i] *= *scale for n float32 values unless *scale == 1 (autograd backward helper) */
int b200det_scale_f32(float *x, long long n, const float *scale_dev, void *stream);

/* ---- decode ------------------------------------------------------------------------- */
/* scratch for b200det_select_decode_nms / b200det_decode.  Since ABI-compatible revision r02 the
 * selection keeps its per-image histograms and candidate lists in (distributed) shared memory of a
 * thread-block cluster; the workspace is accepted and not touched (returns a token 256 bytes;
 * 0 = bad geometry / topn).  workspace == NULL is allowed. */
size_t b200det_decode_workspace_bytes(const b200det_geometry *geo, int topn);

/*
 * Per-row class arg-max / score / threshold: one streaming pass over cls.
 * Replaces np.concatenate + np.argmax + score gather (decode.py:208-238 / :300-338, incl. the
 * FCOS sqrt(cls*centerness)) and the `score > min_score_threshold` filter (decode.py:133-138).
 *   ctr     : NULL for Retina; per-level centre-ness pointers for FCOS
 *   keys    : device uint32 [B*N] level-major: 0 for rows at or below the threshold, else an
 *             order-preserving transform of the float32 score bits
 *   classes : device int32 [B*N] level-major arg-max class (first maximum)
 */
int b200det_score_argmax(const b200det_geometry *geo, const void *const *cls,
                         const void *const *ctr, float min_score, uint32_t *keys,
                         int32_t *classes, void *stream);

/*
 * Global top-n per image (histogram select + placement by histogram rank), box decode, NMS,
 * max_object_num cap: ONE launch, a cluster of 1-8 CTAs per image.
 * Replaces DecodeMethod.__call__ (decode.py:121-172), DetNMSMethod.__call__ (:34-104),
 * RetinaDecoder.snap_txtytwth_to_x1y1x2y2 (:251-271) / FCOSDecoder.snap_ltrb_to_x1y1x2y2
 * (:350-364) incl. NumPy's float32 exp and the int32 truncation.
 *   is_fcos       : B200DET_DECODE_*: 0 = anchor (tx,ty,tw,th) decoding, 1 = point (l,t,r,b)
 *                   decoding, 2 = reg[l] already holds float32 boxes (DecodeMethod on its own,
 *                   DETR-style decoders; the geometry is then one level of 1 x N rows)
 *   min_score     : the threshold given to b200det_score_argmax (sizes the selection histogram)
 *   scales/sizes/to_xywh : optional evaluation glue of the reference's test loop, fused into the
 *                   epilogue (tools/scripts.py:742-757): scales = device float32 [B] -> boxes /=
 *                   scale; sizes = device float32 [B,2] (h,w) -> clip x1,y1 >= 0, x2 <= w, y2 <= h
 *                   and, if to_xywh, x2,y2 -> w,h.  NULL = the decoder's plain output.
 *   out           : device float32 [6*B*max_out]: scores [B,max_out] (pad -1), classes
 *                   [B,max_out] (pad -1), boxes [B,max_out,4] (pad 0), back to back
 *   order / keep  : NULL or device int32 [B,topn] (pad -1): image-major row index of the sorted
 *                   top-n, and the positions in that list surviving NMS (full list, no cap)
 *   counts        : NULL or device int32 [B,3]: candidates, selected (<= topn), kept by NMS
 *                   (the NMS count stops at max_out unless `keep` is given)
 */
int b200det_select_decode_nms(const b200det_geometry *geo, const uint32_t *keys,
                              const int32_t *classes, const void *const *reg, int reg_dtype,
                              int is_fcos, float min_score, int topn, int max_out, int nms_type,
                              double nms_threshold, const float *scales, const float *sizes,
                              int to_xywh, float *out, int32_t *order, int32_t *keep,
                              int32_t *counts, void *workspace, size_t workspace_bytes,
                              void *stream);

/*
 * Front end of the query-based decoders and of a stand-alone DecodeMethod: per (image, query) row
 * the class scores, their first-maximum class, the class / score filters and the box transform.
 * Replaces DETRDecoder.__call__ (decode.py:388-431: softmax, argmax, score gather, cxcywh ->
 * xyxy, * [w,h,w,h], `class < num_classes`, `score > min_score_threshold`) and
 * DINODETRDecoder.__call__ (:512-556: the same with sigmoid and no class filter); the sort /
 * top-n / NMS / cap that follow (:433-470, :558-592) are b200det_select_decode_nms with
 * B200DET_DECODE_BOXES on a one-level geometry of 1 x queries rows.
 *   cls          : device [batch, queries, channels]; float32 (F16 / BF16 also for SIGMOID)
 *   mode         : B200DET_SCORES_*.  SOFTMAX reproduces torch's CUDA softmax kernel for rows of
 *                  <= 1024 channels bit for bit (per-lane strided sums, xor-butterfly reduction)
 *   boxes_cxcywh : device float32 [batch, queries, 4] normalised cx,cy,w,h, or NULL (no boxes)
 *   sizes_hw     : device float32 [batch, 2] = scaled (h, w) of every image (with boxes_cxcywh)
 *   num_classes  : rows whose arg-max class is >= num_classes are dropped (DETR's no-object
 *                  channel); pass `channels` to keep every row
 *   keys/classes : device uint32 / int32 [batch*queries], as b200det_score_argmax
 *   boxes_xyxy   : device float32 [batch, queries, 4] out (with boxes_cxcywh), 16-byte aligned
 */
int b200det_query_scores(const void *cls, int cls_dtype, int mode, const float *boxes_cxcywh,
                         const float *sizes_hw, int batch, int queries, int channels,
                         int num_classes, float min_score, uint32_t *keys, int32_t *classes,
                         float *boxes_xyxy, void *stream);

/* ---- fused host entries (one C call per loss forward / per decode) -------------------- */
typedef struct b200det_loss_params {
    int32_t is_fcos;            /* 0 RetinaLoss, 1 FCOSLoss */
    int32_t box_loss;           /* B200DET_BOX_* */
    int32_t reg_dtype;          /* B200DET_F32 / F16 / BF16 */
    int32_t use_center_sample;  /* FCOS */
    float alpha, gamma, beta;
    float w_cls, w_box, w_ctr;  /* cls_loss_weight, box_loss_weight, center_ness_loss_weight */
    float iou_neg, iou_pos;     /* anchor assignment thresholds (Retina 0.4 / 0.5, RetinaFace 0.35 / 0.35) */
} b200det_loss_params;

typedef struct b200det_decode_params {
    int32_t is_fcos, reg_dtype, topn, max_out, nms_type;
    float min_score;
    double nms_threshold;
    const float *scales;  /* device [B] or NULL, see b200det_select_decode_nms */
    const float *sizes;   /* device [B,2] (h,w) or NULL */
    int32_t to_xywh;
    /* float16 regression head (reg_dtype = B200DET_F16 | B200DET_REG_EXP_ROUNDED) only: device
     * uint16[65536] = the float16 bits of np.exp for every float16 input as the HOST's NumPy
     * computes it, or NULL = half(expf(float(x))), NumPy's generic half loop.  np.exp on float16 is
     * CPU-dependent (hosts with AVX512-FP16 take an SVML kernel that differs from the correctly
     * rounded value for 17 % of the inputs); the Python layer passes the table that matches the
     * host it runs on (tools/make_half_exp_table.py). */
    const uint16_t *half_exp_table;
} b200det_decode_params;

/*
 * Whole forward pass of RetinaLoss.forward / FCOSLoss.forward (losses.py:161-218, :462-511) in one
 * call: label-free focal sweep, assignment, sparse losses, reduction and (if `losses` != NULL)
 * normalisation.  Equivalent to calling b200det_focal_loss(labels = NULL), b200det_*_assign,
 * b200det_sparse_losses(cls given), b200det_loss_reduce(3), b200det_loss_finish in that order,
 * with a single memset.  Pass losses = NULL to all-reduce `sums` across ranks first.
 */
int b200det_loss_forward(const b200det_geometry *geo, const b200det_loss_params *params,
                         const float *annotations, int max_gt, const void *const *cls,
                         const void *const *reg, const void *const *ctr, int32_t *labels,
                         void *workspace, size_t workspace_bytes, double *sums, float *losses,
                         void *stream);

/*
 * Training-step forward of the losses in one call (single process group member; with a sharded
 * normaliser use the per-kernel calls and all-reduce `sums` in between): assignment, sparse
 * losses writing d(box)/d(reg) and d(ctr), reduce, label-aware focal sweep writing d(cls) already
 * multiplied by w_cls / positives, reduce, finish.  reg_grad / ctr_grad must be zeroed by the
 * caller (only the positives' rows are written).
 */
int b200det_loss_forward_grad(const b200det_geometry *geo, const b200det_loss_params *params,
                              const float *annotations, int max_gt, const void *const *cls,
                              const void *const *reg, const void *const *ctr, int32_t *labels,
                              void *const *cls_grad, void *const *reg_grad,
                              void *const *ctr_grad, void *workspace, size_t workspace_bytes,
                              double *sums, float *losses, void *stream);

/* b200det_loss_forward_grad with the sparse losses on a caller-owned helper stream beside the
 * gradient-writing sweep (see b200det_loss_forward_overlap for the stream / event contract) */
int b200det_loss_forward_grad_overlap(const b200det_geometry *geo, const b200det_loss_params *params,
                                      const float *annotations, int max_gt, const void *const *cls,
                                      const void *const *reg, const void *const *ctr, int32_t *labels,
                                      void *const *cls_grad, void *const *reg_grad,
                                      void *const *ctr_grad, void *workspace, size_t workspace_bytes,
                                      double *sums, float *losses, void *side_stream, void *ev_fork,
                                      void *ev_join, void *stream);

/*
 * Whole RetinaDecoder.__call__ / FCOSDecoder.__call__ (decode.py:201-249, :293-348) up to the
 * device-side result: b200det_score_argmax followed by b200det_select_decode_nms.
 */
int b200det_decode(const b200det_geometry *geo, const b200det_decode_params *params,
                   const void *const *cls, const void *const *ctr, const void *const *reg,
                   uint32_t *keys, int32_t *classes, float *out, int32_t *order, int32_t *keep,
                   int32_t *counts, void *workspace, size_t workspace_bytes, void *stream);

/*
 * b200det_decode without its sweep: selection, box decode and NMS on `keys` / `classes` that
 * b200det_loss_forward_keys left behind for the same head outputs (decode.py:133-167 on the rows the
 * criterion's sweep scored).  cls (and ctr for FCOS) are only read to VERIFY the hand-over: for every
 * selected row the kernel recomputes the key from the class score it names; if one differs -- the head
 * outputs were modified after the criterion read them -- it sets *stale = 1 and the caller must decode
 * again with b200det_decode.  stale: caller-zeroed int32 the device can write (device or mapped pinned
 * host memory), or NULL (then cls / ctr may be NULL).
 * inputs_complete != 0: the caller guarantees that everything this call reads (keys, classes, cls, ctr,
 * reg, params' device arrays) was complete before the kernel that precedes it on `stream` STARTED --
 * true for handed-over keys, whose predecessor is the criterion's reduction / peer-exchange kernel.
 * The select kernel is launched programmatically and then skips its griddepcontrol.wait, i.e. it runs
 * BESIDE that one-CTA kernel (and its wait for the other ranks) instead of after it.  0 = stream order.
 */
int b200det_decode_from_keys(const b200det_geometry *geo, const b200det_decode_params *params,
                             const void *const *cls, const void *const *ctr, const void *const *reg,
                             const uint32_t *keys, const int32_t *classes, float *out,
                             int32_t *order, int32_t *keep, int32_t *counts, int32_t *stale,
                             int inputs_complete, void *stream);

/*
 * OPTIONAL extension (not in the reference's call structure): loss forward + decode of one
 * evaluation step with a SINGLE sweep over the classification tensors -- the score / arg-max
 * sweep also accumulates the label-free focal sum.  Same results as b200det_loss_forward followed
 * by b200det_decode (any class count: multiples of 4 take the TMA-fed row-group sweep, others the raw-tile
 * sweep).  Pass losses = NULL to all-reduce `sums`.
 */
int b200det_eval_step(const b200det_geometry *geo, const b200det_loss_params *loss_params,
                      const b200det_decode_params *decode_params, const float *annotations,
                      int max_gt, const void *const *cls, const void *const *reg,
                      const void *const *ctr, int32_t *labels, void *loss_workspace,
                      size_t loss_workspace_bytes, double *sums, float *losses, uint32_t *keys,
                      int32_t *classes, float *out, void *decode_workspace,
                      size_t decode_workspace_bytes, void *stream);

/* b200det_eval_step with the assignment and the sparse losses on a caller-owned helper stream beside
 * the sweep and the selection (fork after the memset, join before the reduction; see
 * b200det_loss_forward_overlap for the stream / event contract).  No decode workspace. */
int b200det_eval_step_overlap(const b200det_geometry *geo, const b200det_loss_params *loss_params,
                              const b200det_decode_params *decode_params, const float *annotations,
                              int max_gt, const void *const *cls, const void *const *reg,
                              const void *const *ctr, int32_t *labels, void *loss_workspace,
                              size_t loss_workspace_bytes, double *sums, float *losses,
                              uint32_t *keys, int32_t *classes, float *out, void *side_stream,
                              void *ev_fork, void *ev_join, void *stream);

/* ---- utilities (tests / parity outputs) ---------------------------------------------- */
/* dst[b*N + off_l + i] = src[B*off_l + b*n_l + i] for `width` int32/float32 words per row */
int b200det_rows_to_image_major(const b200det_geometry *geo, const void *src, void *dst,
                                int width, void *stream);
/* materialise the anchors (per_loc > 1, [N,4]) or points ([N,2]) the kernels generate */
int b200det_generate_rows(const b200det_geometry *geo, int is_fcos, float *out, void *stream);
/* y[i] = exp(x[i]) with the NumPy float32 algorithm used by the decoders (test hook) */
int b200det_npexp_f32(const float *x, float *y, long long n, void *stream);
/* profiling hook (tools/prof_select.py): device int64 [B,16] receiving %globaltimer stamps of the
 * select kernel's phases (0 start, 1 histograms, 2 candidate list, 3 order, 4 box decode, 5 NMS,
 * 6 outputs), or NULL = off (default).  Process-wide; not for concurrent use. */
int b200det_select_stamps(long long *stamps);

/* ---- head tail: sigmoid + NCHW -> NHWC in one pass --------------------------------
 * Replaces `x = x.float(); x = self.sigmoid(x)` (models/head.py:46-50 RetinaClsHead.forward,
 * :176-179 FCOSClsRegCntHead.forward) followed by `permute(0, 2, 3, 1).contiguous()`
 * (models/retinanet.py:73-77, models/fcos.py:70-79).
 *   src  [batch, channels, hw]  convolution output, B200DET_F32 / F16 / BF16, contiguous
 *   dst  [batch, hw, channels]  float32 probabilities = 1 / (1 + expf(-float(src))); 16-byte
 *        aligned when channels % 4 == 0
 * batch <= 65535, channels * hw < 2^31. */
int b200det_head_sigmoid_permute(const void *src, int src_dtype, int batch, int channels,
                                 long long hw, float *dst, void *stream);
/* Backward of the above (torch sigmoid_backward + permute + `.float()` backward):
 *   grad_in[b, c, i] = grad_dtype( (grad_out[b, i, c] * (1 - probs[b, i, c])) * probs[b, i, c] ) */
int b200det_head_sigmoid_permute_backward(const float *grad_out, const float *probs, int batch,
                                          int channels, long long hw, void *grad_in,
                                          int grad_dtype, void *stream);

/* ---- classification branch straight from NCHW logits (second half) ------------------
 * One pass over the head's logits [batch, per_loc * num_classes, H, W] (B200DET_F32 / F16 / BF16,
 * contiguous) instead of sigmoid + permute (models/head.py:46-50, models/retinanet.py:73-76) followed
 * by the focal loss (losses.py:220-261) and / or the decoder's arg-max / score / threshold
 * (decode.py:208-238).  No probability tensor is written.
 *   ctr_logits     FCOS centre-ness logits [batch, H*W] float32 per level (score = sqrt(cls * ctr)),
 *                  or NULL
 *   labels         level-major labels from b200det_*_assign; NULL: every row counts as background
 *   loss_workspace NULL: no focal sum; else the label-aware focal sum is accumulated into the
 *                  workspace's sweep accumulators (read it with b200det_loss_reduce, which = 2 or 3)
 *   keys, classes  NULL: no arg-max; else the decoder keys / classes of b200det_score_argmax,
 *                  bit-identical to running it on torch's CUDA sigmoid of the logits
 * kernel id 9 (logits_sweep). */
int b200det_logits_sweep(const b200det_geometry *geo, const void *const *cls_logits, int cls_dtype,
                         const void *const *ctr_logits, const int32_t *labels, float alpha,
                         float gamma, void *loss_workspace, size_t loss_workspace_bytes,
                         float min_score, uint32_t *keys, int32_t *classes, void *stream);
/* b200det_eval_step for heads whose classification tensors are NCHW logits: assignment, box (and
 * centre-ness) loss, ONE sweep over the logits (focal sum + decoder keys), reduce / finish,
 * select + decode + NMS.  reg stays [B, H, W, A, 4] / [B, H, W, 4] as in the reference; FCOS heads
 * (lp->is_fcos, dp->is_fcos) pass their float32 centre-ness logits [B, 1, H, W] as ctr_logits,
 * RetinaNet-style heads pass NULL. */
int b200det_logits_eval_step(const b200det_geometry *geo, const b200det_loss_params *lp,
                             const b200det_decode_params *dp, const float *annotations, int max_gt,
                             const void *const *cls_logits, int cls_dtype, const void *const *reg,
                             const void *const *ctr_logits, int32_t *labels, void *loss_workspace,
                             size_t loss_workspace_bytes,
                             double *sums, float *losses, uint32_t *keys, int32_t *classes,
                             float *out, void *decode_workspace, size_t decode_workspace_bytes,
                             void *stream);

/* ---- IoUMethod as a stand-alone operator ------------------------------------------------------ */
/* IoUMethod.__call__ (simpleAICV/detection/losses.py:28-123): out[i*m + j] = iou_type(boxes1 row
 * i*s1n + j*s1m, boxes2 row i*s2n + j*s2m), float32, the reference's op order.  Element-wise use:
 * m = 1, strides (1, 0); the assignment's [N,1,4] x [1,M,4] broadcast (losses.py:350-353):
 * strides (1, 0) and (0, 1).  Strides count boxes (4 floats).  iou_type: B200DET_BOX_IOU ..
 * B200DET_BOX_EIOU; xywh != 0: boxes are [x_ctr, y_ctr, w, h] (losses.py:44-52).
 * jac1 / jac2 (optional, 16-byte aligned, float32 [n*m, 4]): d out / d (the four inputs of the
 * boxes1 / boxes2 row) with autograd's conventions (0.5 / 0.5 on max / min ties, inclusive clamp
 * masks, CIoU's alpha constant). */
int b200det_iou_method(const float *boxes1, long long s1n, long long s1m, const float *boxes2,
                       long long s2n, long long s2m, long long n, long long m, int iou_type,
                       int xywh, float *out, float *jac1, float *jac2, void *stream);

/* ---- evaluation: the per-batch part of the VOC evaluator --------------------------------------- */
/* compute_ious (tools/scripts.py:487-508): out[i*m + j] = IoU(a[i], b[j]) in float32 with the
 * reference's op order and no clamps (degenerate pairs give NaN / inf like NumPy).  a, b: device
 * float32 [n,4] / [m,4] x1,y1,x2,y2, 16-byte aligned. */
int b200det_pair_ious(const float *a, int n, const float *b, int m, float *out, void *stream);
/* The matching loop of evaluate_voc_detection (tools/scripts.py:626-651) for a batch: detection d of
 * image b (decoder order; class <= -1 = padding) is a true positive at threshold t iff the ground-
 * truth box of its class with the largest IoU (np.argmax: first maximum, NaN counts as maximum) has
 * IoU >= thresholds[t] and was not taken by an earlier detection of that image.
 *   pred_boxes [batch,max_det,4], pred_classes [batch,max_det], gt_boxes [batch,max_gt,4],
 *   gt_classes [batch,max_gt] (<= -1 = padding), thresholds [n_thresholds]: device float32
 *   tp : device uint8 [n_thresholds, batch, max_det] out */
int b200det_voc_match(const float *pred_boxes, const float *pred_classes, int max_det,
                      const float *gt_boxes, const float *gt_classes, int max_gt,
                      const float *thresholds, int n_thresholds, int batch, unsigned char *tp,
                      void *stream);

/* ---- multi-GPU: the loss normaliser exchanged over NVLink peer memory ---------------------- */
/*
 * The path shards by image; the only cross-rank coupling is the whole-batch positive count and the
 * loss sums (losses.py:231-259 on the unsharded batch): 4 doubles.  Instead of reduce -> host call
 * -> NCCL all-reduce -> finish, b200det_loss_reduce_exchange is ONE kernel per rank: it reduces the
 * rank's block partials, stores its 4 doubles into every peer's exchange buffer (peer memory mapped
 * with CUDA IPC, st.release.sys), waits (ld.acquire.sys, bounded) for the peers' stores in its own
 * buffer, adds the contributions in rank order (bit-identical totals on every rank) and writes the
 * normalised losses.  One process per GPU, all ranks on one NVLink / NVSwitch domain.
 *
 * Exchange buffers are the one thing this library allocates (4 KB of device memory each, owned by
 * the caller through these handles):
 *   b200det_peer_buffer_create   cudaMalloc + zero + cudaIpcGetMemHandle (64-byte handle to send to
 *                                the peers, e.g. with torch.distributed.all_gather_object)
 *   b200det_peer_buffer_open     maps a peer's buffer into this process (enables peer access)
 *   b200det_peer_buffer_close / _destroy
 */
int b200det_peer_buffer_create(void **buffer, unsigned char *handle64);
int b200det_peer_buffer_open(const unsigned char *handle64, void **mapped);
int b200det_peer_buffer_close(void *mapped);
int b200det_peer_buffer_destroy(void *buffer);

typedef struct b200det_peer_exchange {
    int32_t rank, world;            /* world <= B200DET_MAX_PEERS */
    uint64_t epoch;                 /* 0 (normal use): the exchange number is kept in device memory (in
                                       the rank's own buffer) and advanced by the kernel, so replays of a
                                       captured CUDA graph advance it like eager calls do; != 0: caller-
                                       numbered exchange, the same value on every rank */
    uint64_t timeout_cycles;        /* wait budget in NANOSECONDS of %globaltimer (the field keeps its
                                       r01 name; 0: 120 s); on expiry the sums and losses become NaN and
                                       *status = 1 -- the kernel never hangs.  After a timeout the ranks'
                                       counters may disagree: re-synchronise (barrier) and recreate the
                                       buffers before the next exchange */
    void *peer[B200DET_MAX_PEERS];  /* peer[r]: rank r's buffer as mapped here; peer[rank]: own */
} b200det_peer_exchange;

/* reduce (which = 3) + exchange + finish; sums = device double[4] GLOBAL totals out; losses = device
 * float[3] or NULL; status = device int32 (0 = ok, 1 = a peer timed out; written every call) or NULL */
int b200det_loss_reduce_exchange(const b200det_geometry *geo, const void *workspace,
                                 size_t workspace_bytes, const b200det_peer_exchange *px,
                                 float w_cls, float w_box, float w_ctr, double *sums, float *losses,
                                 int32_t *status, void *stream);
/* the exchange alone: sums = device double[4] of this rank in, totals over the ranks out (in place);
 * losses as above or NULL.  world <= 32 threads: one warp. */
int b200det_sums_exchange(const b200det_peer_exchange *px, float w_cls, float w_box, float w_ctr,
                          double *sums, float *losses, int32_t *status, void *stream);
/* b200det_loss_forward ending in b200det_loss_reduce_exchange */
int b200det_loss_forward_exchange(const b200det_geometry *geo, const b200det_loss_params *params,
                                  const float *annotations, int max_gt, const void *const *cls,
                                  const void *const *reg, const void *const *ctr, int32_t *labels,
                                  void *workspace, size_t workspace_bytes,
                                  const b200det_peer_exchange *px, double *sums, float *losses,
                                  int32_t *status, void *stream);

/*
 * b200det_loss_forward / b200det_loss_forward_exchange (px != NULL) with the assignment and the
 * sparse losses -- which do not read what the focal sweep writes -- enqueued on `side_stream` BESIDE
 * the HBM-bound sweep: fork (ev_fork) after the memset, join (ev_join) before the reduction, so that
 * everything is ordered on `stream` again when the call returns; capturable in a CUDA graph.  The
 * stream and the two events belong to the caller (one set per criterion object); NULL for any of
 * the three = everything on `stream`.  status may be NULL without px.
 * phase: 0 = the whole call.  A host layer that wants the HBM-bound sweep on the GPU as early as
 * possible splits it: phase 1 = memset + fork + focal sweep only (reads geo, params, cls, workspace;
 * the other pointers may be NULL), then -- after it has prepared the remaining arguments -- phase 2 =
 * assignment, sparse losses, join, reduction with the same geo / cls / workspace / streams / events.
 */
int b200det_loss_forward_overlap(const b200det_geometry *geo, const b200det_loss_params *params,
                                 const float *annotations, int max_gt, const void *const *cls,
                                 const void *const *reg, const void *const *ctr, int32_t *labels,
                                 void *workspace, size_t workspace_bytes,
                                 const b200det_peer_exchange *px, double *sums, float *losses,
                                 int32_t *status, void *side_stream, void *ev_fork, void *ev_join,
                                 void *stream, int phase);
/*
 * b200det_loss_forward_overlap whose sweep ALSO does the decoder's: the fused sweep of b200det_eval_step
 * (focal sum + first-maximum class and score key of every row, thresholded with `min_score` like
 * b200det_score_argmax) instead of the focal-only one.  The reference's evaluation loop calls
 * criterion(outs, annots) and then decoder(outs) on the SAME head outputs (tools/scripts.py:733-740);
 * after this call the decoder only needs b200det_select_decode_nms on `keys` / `classes`, so the
 * classification tensors -- 98 % of the step's HBM traffic -- are read once per step while the caller
 * keeps the reference's two calls.  The host layer (b200det._handoff) decides when that is safe.
 * Same arguments as b200det_loss_forward_overlap plus min_score / keys / classes as
 * b200det_score_argmax.  FCOS needs `ctr` in phase 1 too.
 */
int b200det_loss_forward_keys(const b200det_geometry *geo, const b200det_loss_params *params,
                              const float *annotations, int max_gt, const void *const *cls,
                              const void *const *reg, const void *const *ctr, int32_t *labels,
                              void *workspace, size_t workspace_bytes,
                              const b200det_peer_exchange *px, double *sums, float *losses,
                              int32_t *status, void *side_stream, void *ev_fork, void *ev_join,
                              void *stream, int phase, float min_score, uint32_t *keys,
                              int32_t *classes);
/* caller-owned helper objects for the call above: a non-blocking stream (high_priority != 0: the
 * device's highest priority) and timing-free events, on the current device */
int b200det_stream_create(void **stream, int high_priority);
int b200det_stream_destroy(void *stream);
int b200det_event_create(void **event);
int b200det_event_destroy(void *event);

#ifdef __cplusplus
}
#endif
#endif /* B200DET_H_ */
