"""b200det.fused -- OPTIONAL extension beyond the reference's call structure.

The reference's evaluation loop makes two calls per batch (tools/scripts.py:733-740):

    loss_value = criterion(outs_tuple, annots)
    scores, classes, boxes = decoder(outs_tuple)

and each of them has to stream the whole classification tensor (4*N*C bytes per image, 98 % of
the path's HBM traffic).  `EvalStep` makes ONE call whose score / arg-max sweep also accumulates
the focal sum, so cls is read once:

    step = fused.EvalStep(criterion, decoder)          # the two drop-in objects, unchanged
    loss_value, (scores, classes, boxes) = step(outs_tuple, annots)

Results are the same as the two separate calls (loss within float rounding of the summation order,
detections bit-identical).  No gradients (evaluation only).
"""
import ctypes

import torch

from . import _lib
from . import geometry as _geom
from .losses import (_DTYPES, _Plan, _decode_reg_mode, _loss_params, _loss_reg_mode, _on_device,
                     _plan_for, _prep_annotations, _prep_f32, _prep_reg, _require_cuda, _side_stream,
                     _sync_sums)

__all__ = ['EvalStep', 'LogitsEvalStep']


class EvalStep:

    def __init__(self, criterion, decoder):
        if criterion._is_fcos != decoder._is_fcos:
            raise ValueError('criterion and decoder belong to different detectors')
        self.criterion = criterion
        self.decoder = decoder

    def __call__(self, preds, annotations, scales=None, sizes=None, to_xywh=False):
        _require_cuda(preds[0][0], 'cls_preds')
        with _on_device(preds[0][0].device):
            return self._call_on(preds, annotations, scales, sizes, to_xywh)

    def _call_on(self, preds, annotations, scales, sizes, to_xywh):
        lib = _lib.load()
        crit, dec = self.criterion, self.decoder
        is_fcos = crit._is_fcos
        cls = _prep_f32([t.detach() for t in preds[0]], 'cls_preds')
        reg, reg_dtype = _prep_reg([t.detach() for t in preds[1]])
        ctr = _prep_f32([t.detach() for t in preds[2]], 'center_preds') if is_fcos else None
        annotations = _prep_annotations(annotations)
        plan = _plan_for(crit, cls)
        device = cls[0].device
        batch, n_rows = plan.batch, plan.n_rows
        m = int(dec.max_object_num)
        dws_bytes = int(lib.b200det_decode_workspace_bytes(plan.geo_ref, int(dec.topn)))
        # scratch = loss workspace | labels | keys | classes | decode workspace
        rows_bytes = (4 * batch * n_rows + 255) & ~255
        scratch = torch.empty(plan.ws_bytes + 3 * rows_bytes + dws_bytes, dtype=torch.uint8,
                              device=device)
        base = scratch.data_ptr()
        labels_ptr = base + plan.ws_bytes
        keys_ptr = labels_ptr + rows_bytes
        classes_ptr = keys_ptr + rows_bytes
        dws_ptr = classes_ptr + rows_bytes
        small = torch.empty(8, dtype=torch.float64, device=device)   # sums | losses
        out = dec._out_buffer(6 * batch * m, device)
        lp = _loss_params(crit, _loss_reg_mode(reg_dtype))
        dp = dec._params
        dp.reg_dtype = _decode_reg_mode(reg_dtype)
        dp.half_exp_table = None
        if reg_dtype == _lib.F16:
            from .decode import _half_exp_table
            table = _half_exp_table(device)
            dp.half_exp_table = table.data_ptr() if table is not None else None
        glue = dec._set_glue(dp, batch, device, scales, sizes, to_xywh)
        sync = crit.sync_normalizer and torch.distributed.is_available() \
            and torch.distributed.is_initialized()
        st = _lib.raw_stream(device)
        sums_ptr = small.data_ptr()
        # assignment + sparse losses on the criterion's helper stream beside the sweep and the selection
        side = _side_stream(crit, device) if not torch.cuda.is_current_stream_capturing() else None
        side_args = (side.stream, side.fork, side.join) if side is not None else (None, None, None)
        _lib.check(
            lib.b200det_eval_step_overlap(plan.geo_ref, ctypes.byref(lp), ctypes.byref(dp),
                                          annotations.data_ptr(), int(annotations.shape[1]),
                                          _lib.ptr_array(cls), _lib.ptr_array(reg),
                                          _lib.ptr_array(ctr), labels_ptr, base, plan.ws_bytes,
                                          sums_ptr, None if sync else sums_ptr + 32, keys_ptr,
                                          classes_ptr, out.data_ptr(), *side_args, st),
            'b200det_eval_step_overlap')
        if sync:
            _sync_sums(crit, small[0:4], crit.process_group, st,
                       finish=(lp.w_cls, lp.w_box, lp.w_ctr, sums_ptr + 32))
        del glue
        crit.last_stats = {'sums': small[0:4]}
        losses = small[4:8].view(torch.float32)
        loss_dict = {'cls_loss': losses[0], 'reg_loss': losses[1]}
        if is_fcos:
            loss_dict['center_ness_loss'] = losses[2]
        return loss_dict, dec._to_host(out, batch, m, device)


class LogitsEvalStep:
    """EvalStep for detectors that hand over the classification head's raw convolution output:

        step = fused.LogitsEvalStep(criterion, decoder)
        loss_value, (scores, classes, boxes) = step([cls_logits, reg_heads], annots)

    cls_logits[l] is [B, A * num_classes, H, W] (NCHW, float32 / float16 / bfloat16), i.e. what
    RetinaClsHead computes BEFORE `x.float(); sigmoid(x)` (models/head.py:46-50) and before the
    permute of models/retinanet.py:73-76; reg_heads[l] stays [B, H, W, A, 4].  One sweep over the
    logits produces the focal sum and the decoder's keys; no probability tensor is written.
    Detections are bit-identical to decoder(sigmoid + permute of the logits); the loss agrees
    within the 1e-5 tolerance.  Evaluation only (no gradients).

    FCOS heads (FCOSLoss + FCOSDecoder): preds = [cls_logits, reg_heads, center_logits] with
    cls_logits[l] [B, num_classes, H, W], reg_heads[l] [B, H, W, 4] as in the reference and
    center_logits[l] [B, 1, H, W] (what FCOSClsRegCntHead computes before its sigmoid,
    models/head.py:176-179); the centre-ness BCE and the decoder's sqrt(cls * centre-ness) use
    torch's CUDA sigmoid of those logits."""

    def __init__(self, criterion, decoder):
        if criterion._is_fcos != decoder._is_fcos:
            raise ValueError('criterion and decoder belong to different detectors')
        self.criterion = criterion
        self.decoder = decoder
        self._plans = {}

    def _plan(self, cls):
        crit = self.criterion
        shape0 = cls[0].shape
        key = (tuple(t.shape[2:4] for t in cls), shape0[0], shape0[1])
        plan = self._plans.get(key)
        if plan is None:
            per_loc = 1 if crit._is_fcos else crit._per_loc
            if shape0[1] % per_loc:
                raise ValueError('logit channels are not a multiple of the anchors per location')
            shapes = [(int(t.shape[2]), int(t.shape[3])) for t in cls]
            batch, num_classes = int(shape0[0]), int(shape0[1]) // per_loc
            geo = crit._geometry(shapes, batch, num_classes)
            plan = _Plan(geo, batch, _geom.rows_per_image(shapes, geo.per_loc))
            self._plans = {key: plan}
        return plan

    def __call__(self, preds, annotations, scales=None, sizes=None, to_xywh=False):
        _require_cuda(preds[0][0], 'cls_preds')
        with _on_device(preds[0][0].device):
            return self._call_on(preds, annotations, scales, sizes, to_xywh)

    def _call_on(self, preds, annotations, scales, sizes, to_xywh):
        lib = _lib.load()
        crit, dec = self.criterion, self.decoder
        cls = [t.detach() for t in preds[0]]
        if not cls or any((not t.is_cuda) or t.dim() != 4 for t in cls):
            raise RuntimeError('cls logits must be CUDA tensors [B, A*C, H, W]')
        dtype = cls[0].dtype
        if dtype not in _DTYPES or any(t.dtype != dtype for t in cls):
            raise RuntimeError('cls logits must share one dtype: float32, float16 or bfloat16')
        cls = [t if t.is_contiguous() else t.contiguous() for t in cls]
        reg, reg_dtype = _prep_reg([t.detach() for t in preds[1]])
        ctr = None
        if crit._is_fcos:
            if len(preds) != 3:
                raise RuntimeError('FCOS heads: preds = [cls_logits, reg_heads, center_logits]')
            ctr = [t.detach().float().contiguous() for t in preds[2]]
            if any((not t.is_cuda) or t.numel() != c.shape[0] * c.shape[2] * c.shape[3]
                   for t, c in zip(ctr, cls)) or len(ctr) != len(cls):
                raise RuntimeError('center logits must be CUDA tensors [B, 1, H, W]')
        annotations = _prep_annotations(annotations)
        plan = self._plan(cls)
        if annotations.shape[0] != plan.batch:
            raise ValueError('annotations and predictions disagree on the batch size')
        device = cls[0].device
        batch, n_rows = plan.batch, plan.n_rows
        m = int(dec.max_object_num)
        dws_bytes = int(lib.b200det_decode_workspace_bytes(plan.geo_ref, int(dec.topn)))
        rows_bytes = (4 * batch * n_rows + 255) & ~255
        scratch = torch.empty(plan.ws_bytes + 3 * rows_bytes + dws_bytes, dtype=torch.uint8,
                              device=device)
        base = scratch.data_ptr()
        labels_ptr = base + plan.ws_bytes
        keys_ptr = labels_ptr + rows_bytes
        classes_ptr = keys_ptr + rows_bytes
        dws_ptr = classes_ptr + rows_bytes
        small = torch.empty(8, dtype=torch.float64, device=device)   # sums | losses
        out = dec._out_buffer(6 * batch * m, device)
        lp = _loss_params(crit, _loss_reg_mode(reg_dtype))
        dp = dec._params
        dp.reg_dtype = _decode_reg_mode(reg_dtype)
        dp.half_exp_table = None
        if reg_dtype == _lib.F16:
            from .decode import _half_exp_table
            table = _half_exp_table(device)
            dp.half_exp_table = table.data_ptr() if table is not None else None
        glue = dec._set_glue(dp, batch, device, scales, sizes, to_xywh)
        sync = crit.sync_normalizer and torch.distributed.is_available() \
            and torch.distributed.is_initialized()
        st = _lib.raw_stream(device)
        sums_ptr = small.data_ptr()
        _lib.check(
            lib.b200det_logits_eval_step(plan.geo_ref, ctypes.byref(lp), ctypes.byref(dp),
                                         annotations.data_ptr(), int(annotations.shape[1]),
                                         _lib.ptr_array(cls), _DTYPES[dtype], _lib.ptr_array(reg),
                                         _lib.ptr_array(ctr), labels_ptr, base, plan.ws_bytes,
                                         sums_ptr,
                                         None if sync else sums_ptr + 32, keys_ptr, classes_ptr,
                                         out.data_ptr(), dws_ptr, dws_bytes, st),
            'b200det_logits_eval_step')
        if sync:
            _sync_sums(crit, small[0:4], crit.process_group, st,
                       finish=(lp.w_cls, lp.w_box, lp.w_ctr, sums_ptr + 32))
        del glue
        crit.last_stats = {'sums': small[0:4]}
        losses = small[4:8].view(torch.float32)
        loss_dict = {'cls_loss': losses[0], 'reg_loss': losses[1]}
        if crit._is_fcos:
            loss_dict['center_ness_loss'] = losses[2]
        return loss_dict, dec._to_host(out, batch, m, device)
