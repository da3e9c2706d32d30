"""b200det.losses -- drop-in RetinaLoss / FCOSLoss backed by libb200det.so (sm_100a CUDA).

Same class names, constructor kwargs, forward(preds, annotations) signature and loss-dict keys
as simpleAICV/detection/losses.py:126-218 (RetinaLoss) and :432-511 (FCOSLoss), so a SimpleAICV
config picks them up with `from b200det import losses` in place of
`from simpleAICV.detection import losses` (3.detection_training/*/train_config.py:37-63).

What runs where: nothing numeric runs in Python or torch.  forward() fills a geometry struct,
collects the per-level device pointers (no torch.cat) and launches
    b200det_retina_assign / b200det_fcos_assign   assignment (pure ALU scan)
    b200det_sparse_losses                         box (+centre-ness) loss of the positives
    b200det_focal_loss                            one streaming pass over cls (+ its gradient)
    b200det_loss_reduce / b200det_loss_finish     deterministic fp64 reduction, normalisation
on the current CUDA stream without any host synchronisation.  There is no CPU path: CPU
tensors raise.

Extra, keyword-only constructor arguments (defaults keep reference behaviour):
    sync_normalizer / process_group : all-reduce {positives, loss sums} over the process
        group so that an image-sharded batch reproduces the single-process full-batch loss
        (SURVEY.md section 8e).  Default False = the reference's per-rank normalisation.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from . import geometry as _geom

__all__ = ['RetinaLoss', 'FCOSLoss']

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f'b200det: {what} must be a CUDA tensor (got device {t.device}); there is no CPU path')


def _prep_f32(levels, what):
    """float32, contiguous, 16-byte aligned per-level tensors (the heads already emit these)."""
    out = []
    for t in levels:
        _require_cuda(t, what)
        if t.dtype != torch.float32:
            t = t.float()
        if not t.is_contiguous():
            t = t.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
        out.append(t)
    return out


def _prep_reg(levels):
    out = []
    dtype = levels[0].dtype
    if dtype not in _DTYPES or any(t.dtype != dtype for t in levels):
        levels = [t.float() for t in levels]
        dtype = torch.float32
    for t in levels:
        _require_cuda(t, 'reg_preds')
        if not t.is_contiguous():
            t = t.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
        out.append(t)
    return out, _DTYPES[dtype]


def _prep_annotations(annotations):
    _require_cuda(annotations, 'annotations')
    if annotations.dim() != 3 or annotations.shape[-1] != 5:
        raise ValueError('annotations must be [B, max_annots, 5] (x1,y1,x2,y2,class; -1 padded)')
    if annotations.dtype != torch.float32:
        annotations = annotations.float()
    if not annotations.is_contiguous():
        annotations = annotations.contiguous()
    if annotations.shape[1] == 0:
        # no annotation rows at all: one invalid row keeps the kernels' G >= 1 contract
        annotations = torch.full((annotations.shape[0], 1, 5), -1., device=annotations.device)
    if annotations.shape[1] > _lib.MAX_GT:
        raise ValueError(f'at most {_lib.MAX_GT} annotation rows per image are supported')
    return annotations


_SIDE_STREAMS = {}
# B200DET_OVERLAP=1 runs assignment + sparse losses on a second (high-priority) stream beside the
# classification sweep.  Measured on B200 this is slower than running them back to back (the
# sweep already saturates HBM and loses more than the overlap gains), so the default is one stream.
_OVERLAP = os.environ.get('B200DET_OVERLAP', '0') == '1'


def _side_stream(device):
    """One extra stream per device for the kernels that can overlap (forked from and joined back
    into the caller's current stream inside every call, so callers see plain stream semantics)."""
    key = (device.type, device.index)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device, priority=-1)
        _SIDE_STREAMS[key] = s
    return s


def _maybe_all_reduce(t, sync, group):
    if sync and torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)


class _DetLossFunction(torch.autograd.Function):
    """One autograd node per loss call.  Gradients are produced by the forward kernels (scaled by
    weight / positives) and only multiplied by the upstream scalars in backward."""

    @staticmethod
    def forward(ctx, owner, annotations, n_levels, *heads):
        lib = _lib.load()
        is_fcos = owner._is_fcos
        cls_in = heads[0:n_levels]
        reg_in = heads[n_levels:2 * n_levels]
        ctr_in = heads[2 * n_levels:3 * n_levels] if is_fcos else ()
        need = ctx.needs_input_grad[3:]
        want_cls = any(need[0:n_levels])
        want_reg = any(need[n_levels:2 * n_levels])
        want_ctr = is_fcos and any(need[2 * n_levels:3 * n_levels])
        want_grad = want_cls or want_reg or want_ctr

        cls = _prep_f32(cls_in, 'cls_preds')
        reg, reg_dtype = _prep_reg(reg_in)
        ctr = _prep_f32(ctr_in, 'center_preds') if is_fcos else None
        annotations = _prep_annotations(annotations)
        device = cls[0].device
        batch = int(cls[0].shape[0])
        if annotations.shape[0] != batch:
            raise ValueError('annotations and predictions disagree on the batch size')
        shapes = _geom.level_shapes(cls)
        num_classes = int(cls[0].shape[-1])
        geo = owner._geometry(shapes, batch, num_classes)
        n_rows = _geom.rows_per_image(shapes, geo.per_loc)
        st = _stream()

        ws_bytes = getattr(geo, '_ws_bytes', None)
        if ws_bytes is None:
            ws_bytes = lib.b200det_loss_workspace_bytes(ctypes.byref(geo))
            geo._ws_bytes = ws_bytes
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        labels = torch.empty(batch * n_rows, dtype=torch.int32, device=device)
        sums = torch.zeros(4, dtype=torch.float64, device=device)
        losses = torch.empty(3, dtype=torch.float32, device=device)

        # the sparse kernel writes gradients of the positive rows only
        reg_grad = [torch.zeros(r.shape, dtype=torch.float32, device=device) for r in reg] \
            if want_grad else None
        ctr_grad = [torch.zeros_like(c) for c in ctr] if (want_grad and is_fcos) else None
        cls_grad = [torch.empty_like(c) for c in cls] if want_grad else None

        sync, group = owner.sync_normalizer, owner.process_group
        alpha, gamma = float(owner.alpha), float(owner.gamma)

        def launch_assign(fix_cls, stream):
            """assignment scan, then the sparse (positive / ignored rows) losses"""
            name = 'fcos_assign' if is_fcos else 'retina_assign'
            with _lib.timed(name):
                if is_fcos:
                    _lib.check(
                        lib.b200det_fcos_assign(ctypes.byref(geo), annotations.data_ptr(),
                                                int(annotations.shape[1]),
                                                int(owner.use_center_sample), labels.data_ptr(),
                                                None, None, ws.data_ptr(), ws_bytes, stream),
                        'b200det_fcos_assign')
                else:
                    _lib.check(
                        lib.b200det_retina_assign(ctypes.byref(geo), annotations.data_ptr(),
                                                  int(annotations.shape[1]), labels.data_ptr(),
                                                  None, ws.data_ptr(), ws_bytes, stream),
                        'b200det_retina_assign')
            with _lib.timed('sparse_losses'):
                _lib.check(
                    lib.b200det_sparse_losses(ctypes.byref(geo), int(is_fcos),
                                              annotations.data_ptr(), int(annotations.shape[1]),
                                              labels.data_ptr(), _lib.ptr_array(reg), reg_dtype,
                                              _lib.ptr_array(ctr), owner._box_code,
                                              float(owner.beta), _lib.ptr_array(fix_cls), alpha,
                                              gamma, _lib.ptr_array(reg_grad),
                                              _lib.ptr_array(ctr_grad), ws.data_ptr(), ws_bytes,
                                              stream), 'b200det_sparse_losses')

        if want_grad:
            # training: label-aware sweep; the focal gradient is written once, already divided
            # by the (global) positive count, so the count must exist before the sweep
            launch_assign(None, st)
            _lib.check(
                lib.b200det_loss_reduce(ctypes.byref(geo), 1, ws.data_ptr(), ws_bytes,
                                        sums.data_ptr(), st), 'b200det_loss_reduce')
            _maybe_all_reduce(sums, sync, group)
            with _lib.timed('focal_loss'):
                _lib.check(
                    lib.b200det_focal_loss(ctypes.byref(geo), _lib.ptr_array(cls),
                                           labels.data_ptr(), alpha, gamma,
                                           _lib.ptr_array(cls_grad), sums.data_ptr(),
                                           float(owner.cls_loss_weight), ws.data_ptr(), ws_bytes,
                                           st), 'b200det_focal_loss')
            focal = torch.zeros(4, dtype=torch.float64, device=device)
            _lib.check(
                lib.b200det_loss_reduce(ctypes.byref(geo), 2, ws.data_ptr(), ws_bytes,
                                        focal.data_ptr(), st), 'b200det_loss_reduce')
            _maybe_all_reduce(focal, sync, group)
            sums = sums + focal
        else:
            # forward only: label-free classification sweep; assignment + sparse losses supply the
            # corrections and are independent of it
            # The sweep is enqueued first: it is the long kernel, so the host-side preparation of
            # the remaining launches overlaps with it.
            cur = torch.cuda.current_stream(device)
            overlap = _OVERLAP
            side = _side_stream(device) if overlap else cur
            if overlap:
                side.wait_stream(cur)
            with _lib.timed('focal_loss'):
                _lib.check(
                    lib.b200det_focal_loss(ctypes.byref(geo), _lib.ptr_array(cls), None, alpha,
                                           gamma, None, None, 0., ws.data_ptr(), ws_bytes, st),
                    'b200det_focal_loss')
            with torch.cuda.stream(side):
                launch_assign(cls, ctypes.c_void_p(side.cuda_stream))
            if overlap:
                cur.wait_stream(side)
            with _lib.timed('loss_reduce'):
                _lib.check(
                    lib.b200det_loss_reduce(ctypes.byref(geo), 3, ws.data_ptr(), ws_bytes,
                                            sums.data_ptr(), st), 'b200det_loss_reduce')
            _maybe_all_reduce(sums, sync, group)
        _lib.check(
            lib.b200det_loss_finish(sums.data_ptr(), float(owner.cls_loss_weight),
                                    float(owner.box_loss_weight),
                                    float(getattr(owner, 'center_ness_loss_weight', 0.)),
                                    losses.data_ptr(), st), 'b200det_loss_finish')

        ctx.n_levels = n_levels
        ctx.is_fcos = is_fcos
        ctx.in_dtypes = [h.dtype for h in heads]
        ctx.in_shapes = [h.shape for h in heads]
        ctx.want = (want_cls, want_reg, want_ctr)
        ctx.weights = (float(owner.box_loss_weight),
                       float(getattr(owner, 'center_ness_loss_weight', 0.)))
        if want_grad:
            ctx.save_for_backward(sums, *cls_grad, *reg_grad, *(ctr_grad or []))
        owner.last_stats = {'sums': sums, 'labels': labels, 'geometry': geo}
        outs = (losses[0], losses[1], losses[2]) if is_fcos else (losses[0], losses[1])
        return outs

    @staticmethod
    def backward(ctx, *grad_out):
        lib = _lib.load()
        n = ctx.n_levels
        saved = ctx.saved_tensors
        sums = saved[0]
        cls_grad = saved[1:1 + n]
        reg_grad = saved[1 + n:1 + 2 * n]
        ctr_grad = saved[1 + 2 * n:1 + 3 * n]
        want_cls, want_reg, want_ctr = ctx.want
        st = _stream()
        npos = sums[0].float()
        inv = torch.where(npos > 0, 1.0 / npos.clamp(min=1.), torch.zeros_like(npos))
        grads = [None] * (3 * n if ctx.is_fcos else 2 * n)
        if want_cls:
            g = grad_out[0].detach().float().contiguous()
            for i in range(n):
                # already scaled by cls_loss_weight / positives; only the upstream scalar is left
                _lib.check(
                    lib.b200det_scale_f32(cls_grad[i].data_ptr(), cls_grad[i].numel(),
                                          g.data_ptr(), st), 'b200det_scale_f32')
                grads[i] = cls_grad[i].view(ctx.in_shapes[i]).to(ctx.in_dtypes[i])
        if want_reg:
            s = grad_out[1].detach().float() * ctx.weights[0] * inv
            for i in range(n):
                grads[n + i] = (reg_grad[i] * s).view(ctx.in_shapes[n + i]).to(ctx.in_dtypes[n + i])
        if want_ctr:
            s = grad_out[2].detach().float() * ctx.weights[1] * inv
            for i in range(n):
                grads[2 * n + i] = (ctr_grad[i] * s).view(ctx.in_shapes[2 * n + i]).to(
                    ctx.in_dtypes[2 * n + i])
        return (None, None, None, *grads)


def _debug_assign(owner, preds, annotations, exact=True):
    """Parity hook (tests / smoke): runs only the assignment kernel and returns the reference's
    intermediate truth in IMAGE-major order: labels [B,N] int32, matched [B,N] int32 (index in
    the image's filtered GT list; -1 = none) and, for FCOS, targets [B,N,6] float32.
    exact=False runs the production scan (no `matched` output: pairs that cannot reach IoU 0.38
    are dropped early) and returns the labels only."""
    lib = _lib.load()
    is_fcos = owner._is_fcos
    cls = _prep_f32(preds[0], 'cls_preds')
    annotations = _prep_annotations(annotations)
    device = cls[0].device
    batch = int(cls[0].shape[0])
    shapes = _geom.level_shapes(cls)
    geo = owner._geometry(shapes, batch, int(cls[0].shape[-1]))
    n_rows = _geom.rows_per_image(shapes, geo.per_loc)
    st = _stream()
    ws_bytes = lib.b200det_loss_workspace_bytes(ctypes.byref(geo))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    labels = torch.empty(batch * n_rows, dtype=torch.int32, device=device)
    matched = torch.empty(batch * n_rows, dtype=torch.int32, device=device) if exact else None
    matched_ptr = matched.data_ptr() if exact else None
    targets = None
    if is_fcos:
        targets = torch.empty(batch * n_rows * 6, dtype=torch.float32, device=device)
        _lib.check(
            lib.b200det_fcos_assign(ctypes.byref(geo), annotations.data_ptr(),
                                    int(annotations.shape[1]), int(owner.use_center_sample),
                                    labels.data_ptr(), matched_ptr, targets.data_ptr(),
                                    ws.data_ptr(), ws_bytes, st),
            'b200det_fcos_assign')
    else:
        _lib.check(
            lib.b200det_retina_assign(ctypes.byref(geo), annotations.data_ptr(),
                                      int(annotations.shape[1]), labels.data_ptr(),
                                      matched_ptr, ws.data_ptr(), ws_bytes, st),
            'b200det_retina_assign')

    def to_image_major(t, width):
        out = torch.empty_like(t)
        _lib.check(
            lib.b200det_rows_to_image_major(ctypes.byref(geo), t.data_ptr(), out.data_ptr(),
                                            width, st), 'b200det_rows_to_image_major')
        return out

    res = {'labels': to_image_major(labels, 1).view(batch, n_rows)}
    if exact:
        res['matched'] = to_image_major(matched, 1).view(batch, n_rows)
    if targets is not None:
        res['targets'] = to_image_major(targets, 6).view(batch, n_rows, 6)
    return res


class RetinaLoss(nn.Module):
    """Drop-in for simpleAICV.detection.losses.RetinaLoss (losses.py:126-429)."""

    _is_fcos = False

    def __init__(self,
                 areas=[[32, 32], [64, 64], [128, 128], [256, 256], [512, 512]],
                 ratios=[0.5, 1, 2],
                 scales=[2**0, 2**(1.0 / 3.0), 2**(2.0 / 3.0)],
                 strides=[8, 16, 32, 64, 128],
                 alpha=0.25,
                 gamma=2,
                 beta=1.0 / 9.0,
                 cls_loss_weight=1.,
                 box_loss_weight=1.,
                 box_loss_type='SmoothL1',
                 *,
                 sync_normalizer=False,
                 process_group=None):
        super(RetinaLoss, self).__init__()
        assert box_loss_type in [
            'SmoothL1',
            'IoU',
            'GIoU',
            'DIoU',
            'CIoU',
            'EIoU',
        ], 'wrong IoU type!'
        self.areas = areas
        self.ratios = ratios
        self.scales = scales
        self.strides = strides
        self.alpha = alpha
        self.gamma = gamma
        self.beta = beta
        self.cls_loss_weight = cls_loss_weight
        self.box_loss_weight = box_loss_weight
        self.box_loss_type = box_loss_type
        self.sync_normalizer = sync_normalizer
        self.process_group = process_group
        self._box_code = _lib.BOX_LOSS_CODES[box_loss_type]
        self._per_loc = len(ratios) * len(scales)
        self._base = _geom.retina_base_anchors(areas, ratios, scales)
        self._geo_cache = {}
        self.last_stats = None

    def _geometry(self, shapes, batch, num_classes):
        key = (tuple(shapes), batch, num_classes)
        geo = self._geo_cache.get(key)
        if geo is None:
            if len(shapes) > len(self.areas):
                raise ValueError('more pyramid levels than anchor areas')
            geo = _geom.make_geometry(shapes, batch, self._per_loc, num_classes, self.strides,
                                      base_anchors=self._base)
            self._geo_cache = {key: geo}
        return geo

    def debug_assign(self, preds, annotations, exact=True):
        return _debug_assign(self, preds, annotations, exact)

    def forward(self, preds, annotations):
        '''
        compute cls loss and reg loss in one batch
        '''
        cls_preds, reg_preds = preds
        n = len(cls_preds)
        assert len(reg_preds) == n
        cls_loss, reg_loss = _DetLossFunction.apply(self, annotations, n, *cls_preds, *reg_preds)
        loss_dict = {
            'cls_loss': cls_loss,
            'reg_loss': reg_loss,
        }
        return loss_dict


class FCOSLoss(nn.Module):
    """Drop-in for simpleAICV.detection.losses.FCOSLoss (losses.py:432-833)."""

    _is_fcos = True

    def __init__(self,
                 strides=[8, 16, 32, 64, 128],
                 mi=[[-1, 64], [64, 128], [128, 256], [256, 512], [512, 100000000]],
                 alpha=0.25,
                 gamma=2.,
                 cls_loss_weight=1.,
                 box_loss_weight=1.,
                 center_ness_loss_weight=1.,
                 box_loss_iou_type='GIoU',
                 center_sample_radius=1.5,
                 use_center_sample=True,
                 *,
                 sync_normalizer=False,
                 process_group=None):
        super(FCOSLoss, self).__init__()
        assert box_loss_iou_type in ['IoU', 'GIoU', 'DIoU', 'CIoU', 'EIoU'], 'wrong IoU type!'
        self.alpha = alpha
        self.gamma = gamma
        self.strides = strides
        self.mi = mi
        self.cls_loss_weight = cls_loss_weight
        self.box_loss_weight = box_loss_weight
        self.center_ness_loss_weight = center_ness_loss_weight
        self.box_loss_iou_type = box_loss_iou_type
        self.center_sample_radius = center_sample_radius
        self.use_center_sample = use_center_sample
        self.sync_normalizer = sync_normalizer
        self.process_group = process_group
        self.beta = 0.
        self._box_code = _lib.BOX_LOSS_CODES[box_loss_iou_type]
        self._geo_cache = {}
        self.last_stats = None

    def _geometry(self, shapes, batch, num_classes):
        key = (tuple(shapes), batch, num_classes)
        geo = self._geo_cache.get(key)
        if geo is None:
            if len(shapes) > len(self.mi):
                raise ValueError('more pyramid levels than mi ranges')
            geo = _geom.make_geometry(shapes, batch, 1, num_classes, self.strides, mi=self.mi,
                                      center_sample_radius=self.center_sample_radius)
            self._geo_cache = {key: geo}
        return geo

    def debug_assign(self, preds, annotations, exact=True):
        return _debug_assign(self, preds, annotations, exact)

    def forward(self, preds, annotations):
        '''
        compute cls loss, reg loss and center-ness loss in one batch
        '''
        cls_preds, reg_preds, center_preds = preds
        n = len(cls_preds)
        assert len(reg_preds) == n and len(center_preds) == n
        cls_loss, reg_loss, center_ness_loss = _DetLossFunction.apply(
            self, annotations, n, *cls_preds, *reg_preds, *center_preds)
        loss_dict = {
            'cls_loss': cls_loss,
            'reg_loss': reg_loss,
            'center_ness_loss': center_ness_loss,
        }
        return loss_dict
