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
detections bit-identical).  Requires num_classes % 4 == 0 and no gradients (evaluation only).
"""
import ctypes

import torch

from . import _lib
from .losses import (_loss_params, _maybe_all_reduce, _plan_for, _prep_annotations, _prep_f32,
                     _prep_reg)

__all__ = ['EvalStep']


class EvalStep:

    def __init__(self, criterion, decoder):
        if criterion._is_fcos != decoder._is_fcos:
            raise ValueError('criterion and decoder belong to different detectors')
        self.criterion = criterion
        self.decoder = decoder

    def __call__(self, preds, annotations, scales=None, sizes=None, to_xywh=False):
        lib = _lib.load()
        crit, dec = self.criterion, self.decoder
        is_fcos = crit._is_fcos
        cls = _prep_f32([t.detach() for t in preds[0]], 'cls_preds')
        reg, reg_dtype = _prep_reg([t.detach() for t in preds[1]])
        ctr = _prep_f32([t.detach() for t in preds[2]], 'center_preds') if is_fcos else None
        annotations = _prep_annotations(annotations)
        plan = _plan_for(crit, cls)
        if int(cls[0].shape[-1]) % 4:
            raise ValueError('EvalStep needs num_classes % 4 == 0; call criterion and decoder '
                             'separately')
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
        lp = _loss_params(crit, reg_dtype)
        dp = dec._params
        dp.reg_dtype = reg_dtype
        glue = dec._set_glue(dp, batch, device, scales, sizes, to_xywh)
        sync = crit.sync_normalizer and torch.distributed.is_available() \
            and torch.distributed.is_initialized()
        st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        sums_ptr = small.data_ptr()
        _lib.check(
            lib.b200det_eval_step(plan.geo_ref, ctypes.byref(lp), ctypes.byref(dp),
                                  annotations.data_ptr(), int(annotations.shape[1]),
                                  _lib.ptr_array(cls), _lib.ptr_array(reg), _lib.ptr_array(ctr),
                                  labels_ptr, base, plan.ws_bytes, sums_ptr,
                                  None if sync else sums_ptr + 32, keys_ptr, classes_ptr,
                                  out.data_ptr(), dws_ptr, dws_bytes, st), 'b200det_eval_step')
        if sync:
            _maybe_all_reduce(small[0:4], True, crit.process_group)
            _lib.check(lib.b200det_loss_finish(sums_ptr, lp.w_cls, lp.w_box, lp.w_ctr,
                                               sums_ptr + 32, st), 'b200det_loss_finish')
        del glue
        crit.last_stats = {'sums': small[0:4]}
        losses = small[4:8].view(torch.float32)
        loss_dict = {'cls_loss': losses[0], 'reg_loss': losses[1]}
        if is_fcos:
            loss_dict['center_ness_loss'] = losses[2]
        return loss_dict, dec._to_host(out, batch, m, device)
