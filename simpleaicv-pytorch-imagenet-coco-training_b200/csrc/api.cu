// api.cu -- fused host entry points: one C call per loss forward / decode.
// For small and medium batches the step time is dominated by host work (argument marshalling,
// one FFI call and one stream operation per kernel); these entries enqueue the whole sequence
// with a single call and a single memset.  They only compose the public per-kernel entry points.
#include "common.cuh"

using namespace b200det;

namespace b200det {
int loss_forward_impl(const b200det_geometry *geo, const b200det_loss_params *p,
                      const float *annotations, int max_gt, const void *const *cls,
                      const void *const *reg, const void *const *ctr, int32_t *labels,
                      void *workspace, size_t workspace_bytes, const b200det_peer_exchange *px,
                      double *sums, float *losses, int32_t *status, void *side, void *ev_fork,
                      void *ev_join, void *stream, int phase, float keys_min_score, uint32_t *keys,
                      int32_t *classes);   // exchange.cu
}

extern "C" int b200det_loss_forward(const b200det_geometry *geo, const b200det_loss_params *p,
                                    const float *annotations, int max_gt, const void *const *cls,
                                    const void *const *reg, const void *const *ctr,
                                    int32_t *labels, void *workspace, size_t workspace_bytes,
                                    double *sums, float *losses, void *stream) {
    return loss_forward_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, workspace,
                             workspace_bytes, nullptr, sums, losses, nullptr, nullptr, nullptr,
                             nullptr, stream, 0, 0.f, nullptr, nullptr);
}

static int loss_forward_grad_impl(const b200det_geometry *geo, const b200det_loss_params *p,
                                  const float *annotations, int max_gt, const void *const *cls,
                                  const void *const *reg, const void *const *ctr, int32_t *labels,
                                  void *const *cls_grad, void *const *reg_grad,
                                  void *const *ctr_grad, void *workspace, size_t workspace_bytes,
                                  double *sums, float *losses, void *side, void *ev_fork,
                                  void *ev_join, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!p || !annotations || !cls || !labels || !workspace || !sums || !cls_grad || !losses)
        return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    const bool fork = side && ev_fork && ev_join;
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(workspace);
    cudaError_t e = cudaMemsetAsync(base + ws.off_focal, 0,
                                    ws.off_counters + 2 * sizeof(int) - ws.off_focal, st);
    if (e != cudaSuccess) return (int)e;
    g_skip_memset = true;
    rc = p->is_fcos ? b200det_fcos_assign(geo, annotations, max_gt, p->use_center_sample, labels,
                                          nullptr, nullptr, workspace, workspace_bytes, stream)
                    : b200det_retina_assign(geo, annotations, max_gt, p->iou_neg, p->iou_pos, labels,
                                            nullptr, workspace, workspace_bytes, stream);
    // the positive count must exist before the sweep (it scales the gradient the sweep writes); it
    // is complete as soon as the assignment is, so the sparse losses (box / centre-ness terms and
    // their gradients) can run beside the sweep on the helper stream
    if (!rc) rc = b200det_loss_reduce(geo, 4, workspace, workspace_bytes, sums, stream);
    void *st_sparse = stream;
    if (!rc && fork) {
        if ((e = cudaEventRecord((cudaEvent_t)ev_fork, st)) != cudaSuccess) rc = (int)e;
        if (!rc && (e = cudaStreamWaitEvent((cudaStream_t)side, (cudaEvent_t)ev_fork, 0)) != cudaSuccess)
            rc = (int)e;
        if (!rc) st_sparse = side;
    }
    if (!rc)
        rc = b200det_sparse_losses(geo, p->is_fcos, annotations, max_gt, labels, reg, p->reg_dtype,
                                   ctr, p->box_loss, p->beta, nullptr, p->alpha, p->gamma,
                                   reg_grad, ctr_grad, workspace, workspace_bytes, st_sparse);
    if (!rc)
        rc = b200det_focal_loss(geo, cls, labels, p->alpha, p->gamma, cls_grad, sums, p->w_cls,
                                workspace, workspace_bytes, stream);
    g_skip_memset = false;
    if (st_sparse != stream) {
        e = cudaEventRecord((cudaEvent_t)ev_join, (cudaStream_t)side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, (cudaEvent_t)ev_join, 0);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    if (!rc)
        rc = b200det_loss_reduce_finish(geo, workspace, workspace_bytes, p->w_cls, p->w_box, p->w_ctr,
                                        sums, losses, stream);
    return rc;
}

extern "C" int b200det_loss_forward_grad(const b200det_geometry *geo, const b200det_loss_params *p,
                                         const float *annotations, int max_gt,
                                         const void *const *cls, const void *const *reg,
                                         const void *const *ctr, int32_t *labels,
                                         void *const *cls_grad, void *const *reg_grad,
                                         void *const *ctr_grad, void *workspace,
                                         size_t workspace_bytes, double *sums, float *losses,
                                         void *stream) {
    return loss_forward_grad_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, cls_grad,
                                  reg_grad, ctr_grad, workspace, workspace_bytes, sums, losses, nullptr,
                                  nullptr, nullptr, stream);
}

extern "C" int b200det_loss_forward_grad_overlap(const b200det_geometry *geo,
                                                 const b200det_loss_params *p,
                                                 const float *annotations, int max_gt,
                                                 const void *const *cls, const void *const *reg,
                                                 const void *const *ctr, int32_t *labels,
                                                 void *const *cls_grad, void *const *reg_grad,
                                                 void *const *ctr_grad, void *workspace,
                                                 size_t workspace_bytes, double *sums, float *losses,
                                                 void *side_stream, void *ev_fork, void *ev_join,
                                                 void *stream) {
    return loss_forward_grad_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, cls_grad,
                                  reg_grad, ctr_grad, workspace, workspace_bytes, sums, losses,
                                  side_stream, ev_fork, ev_join, stream);
}

extern "C" int b200det_decode(const b200det_geometry *geo, const b200det_decode_params *p,
                              const void *const *cls, const void *const *ctr,
                              const void *const *reg, uint32_t *keys, int32_t *classes,
                              float *out, int32_t *order, int32_t *keep, int32_t *counts,
                              void *workspace, size_t workspace_bytes, void *stream) {
    if (!p) return B200DET_EINVAL;
    int rc = b200det_score_argmax(geo, cls, p->is_fcos ? ctr : nullptr, p->min_score, keys,
                                  classes, stream);
    if (!rc)
        rc = select_decode_nms_impl(geo, keys, classes, reg, p->reg_dtype, p->is_fcos, p->min_score,
                                    p->topn, p->max_out, p->nms_type, p->nms_threshold, p->scales,
                                    p->sizes, p->to_xywh, out, order, keep, counts,
                                    p->half_exp_table, stream);
    return rc;
}

// The decoder's tail alone, on keys / classes that the criterion's sweep left behind
// (b200det_loss_forward_keys): selection, box decode, NMS.  cls (+ ctr for FCOS) are only used to
// VERIFY the hand-over: the select kernel re-derives the key of every selected row from the class
// score it names and sets *stale (caller-zeroed, may be mapped host memory) on a mismatch.
extern "C" int b200det_decode_from_keys(const b200det_geometry *geo, const b200det_decode_params *p,
                                        const void *const *cls, const void *const *ctr,
                                        const void *const *reg, const uint32_t *keys,
                                        const int32_t *classes, float *out, int32_t *order,
                                        int32_t *keep, int32_t *counts, int32_t *stale,
                                        int inputs_complete, void *stream) {
    if (!p) return B200DET_EINVAL;
    return select_decode_nms_impl(geo, keys, classes, reg, p->reg_dtype, p->is_fcos, p->min_score,
                                  p->topn, p->max_out, p->nms_type, p->nms_threshold, p->scales,
                                  p->sizes, p->to_xywh, out, order, keep, counts, p->half_exp_table,
                                  stream, cls, p->is_fcos ? ctr : nullptr, stale, inputs_complete != 0);
}

// Loss forward + decode of one evaluation step with ONE sweep over the classification tensors:
// the score/arg-max sweep also accumulates the label-free focal sum.  Not part of the reference's
// call structure (criterion and decoder are separate calls there, tools/scripts.py:733-740); an
// optional extension for eval loops that are willing to make one call instead of two.
static int eval_step_impl(const b200det_geometry *geo, const b200det_loss_params *lp,
                          const b200det_decode_params *dp, const float *annotations, int max_gt,
                          const void *const *cls, const void *const *reg, const void *const *ctr,
                          int32_t *labels, void *loss_workspace, size_t loss_workspace_bytes,
                          double *sums, float *losses, uint32_t *keys, int32_t *classes, float *out,
                          void *side, void *ev_fork, void *ev_join, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!lp || !dp || !annotations || !cls || !labels || !loss_workspace || !sums || !keys ||
        !classes || !out)
        return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (loss_workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    const bool fork = side && ev_fork && ev_join;
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(loss_workspace);
    cudaError_t e = cudaMemsetAsync(base + ws.off_focal, 0,
                                    ws.off_counters + 2 * sizeof(int) - ws.off_focal, st);
    if (e != cudaSuccess) return (int)e;
    void *st_sparse = stream;
    if (fork) {
        // assignment + sparse losses beside the sweep AND the selection (neither reads their output)
        if ((e = cudaEventRecord((cudaEvent_t)ev_fork, st)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent((cudaStream_t)side, (cudaEvent_t)ev_fork, 0)) != cudaSuccess)
            return (int)e;
        st_sparse = side;
    }
    g_skip_memset = true;
    rc = score_argmax_impl(geo, cls, dp->is_fcos ? ctr : nullptr, dp->min_score, keys, classes,
                           lp->alpha, lp->gamma, reinterpret_cast<long long *>(base + ws.off_focal),
                           stream);
    if (!rc) {
        rc = lp->is_fcos ? b200det_fcos_assign(geo, annotations, max_gt, lp->use_center_sample,
                                               labels, nullptr, nullptr, loss_workspace,
                                               loss_workspace_bytes, st_sparse)
                         : b200det_retina_assign(geo, annotations, max_gt, lp->iou_neg, lp->iou_pos,
                                                 labels, nullptr, loss_workspace,
                                                 loss_workspace_bytes, st_sparse);
    }
    if (!rc)
        rc = b200det_sparse_losses(geo, lp->is_fcos, annotations, max_gt, labels, reg,
                                   lp->reg_dtype, ctr, lp->box_loss, lp->beta, cls, lp->alpha,
                                   lp->gamma, nullptr, nullptr, loss_workspace,
                                   loss_workspace_bytes, st_sparse);
    g_skip_memset = false;
    if (!rc && fork)   // the selection only needs the sweep's keys: it goes in front of the join
        rc = select_decode_nms_impl(geo, keys, classes, reg, dp->reg_dtype, dp->is_fcos,
                                    dp->min_score, dp->topn, dp->max_out, dp->nms_type,
                                    dp->nms_threshold, dp->scales, dp->sizes, dp->to_xywh, out,
                                    nullptr, nullptr, nullptr, dp->half_exp_table, stream);
    if (fork) {
        e = cudaEventRecord((cudaEvent_t)ev_join, (cudaStream_t)side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, (cudaEvent_t)ev_join, 0);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    if (!rc)   // reduction and normalisation in one launch unless the caller all-reduces in between
        rc = losses ? b200det_loss_reduce_finish(geo, loss_workspace, loss_workspace_bytes, lp->w_cls,
                                                 lp->w_box, lp->w_ctr, sums, losses, stream)
                    : b200det_loss_reduce(geo, 3, loss_workspace, loss_workspace_bytes, sums, stream);
    if (!rc && !fork)
        rc = select_decode_nms_impl(geo, keys, classes, reg, dp->reg_dtype, dp->is_fcos,
                                    dp->min_score, dp->topn, dp->max_out, dp->nms_type,
                                    dp->nms_threshold, dp->scales, dp->sizes, dp->to_xywh, out,
                                    nullptr, nullptr, nullptr, dp->half_exp_table, stream);
    return rc;
}

extern "C" int b200det_eval_step(const b200det_geometry *geo, const b200det_loss_params *lp,
                                 const b200det_decode_params *dp, const float *annotations,
                                 int max_gt, const void *const *cls, const void *const *reg,
                                 const void *const *ctr, int32_t *labels, void *loss_workspace,
                                 size_t loss_workspace_bytes, double *sums, float *losses,
                                 uint32_t *keys, int32_t *classes, float *out,
                                 void *decode_workspace, size_t decode_workspace_bytes,
                                 void *stream) {
    (void)decode_workspace;
    (void)decode_workspace_bytes;
    return eval_step_impl(geo, lp, dp, annotations, max_gt, cls, reg, ctr, labels, loss_workspace,
                          loss_workspace_bytes, sums, losses, keys, classes, out, nullptr, nullptr,
                          nullptr, stream);
}

extern "C" int b200det_eval_step_overlap(const b200det_geometry *geo, const b200det_loss_params *lp,
                                         const b200det_decode_params *dp, const float *annotations,
                                         int max_gt, const void *const *cls, const void *const *reg,
                                         const void *const *ctr, int32_t *labels,
                                         void *loss_workspace, size_t loss_workspace_bytes,
                                         double *sums, float *losses, uint32_t *keys,
                                         int32_t *classes, float *out, void *side_stream,
                                         void *ev_fork, void *ev_join, void *stream) {
    return eval_step_impl(geo, lp, dp, annotations, max_gt, cls, reg, ctr, labels, loss_workspace,
                          loss_workspace_bytes, sums, losses, keys, classes, out, side_stream, ev_fork,
                          ev_join, stream);
}

// cudaStreamSynchronize for the host layer (the decoder's only host wait): avoids building a
// torch.cuda.Stream object per call; ctypes releases the GIL while it blocks.
extern "C" int b200det_stream_synchronize(void *stream) {
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}
