"""b200det.losses -- drop-in RetinaLoss / FCOSLoss backed by libb200det.so (sm_100a CUDA).

Same class names, constructor kwargs, forward(preds, annotations) signature and loss-dict keys
as simpleAICV/detection/losses.py:126-218 (RetinaLoss) and :432-511 (FCOSLoss), so a SimpleAICV
config picks them up with `from b200det import losses` in place of
`from simpleAICV.detection import losses` (3.detection_training/*/train_config.py:37-63).

What runs where: nothing numeric runs in Python or torch.  forward() fills a geometry struct,
collects the per-level device pointers (no torch.cat) and enqueues, on the current CUDA stream and
without any host synchronisation,
  * evaluation / no_grad (tools/scripts.py:733): ONE C call, b200det_loss_forward =
        focal sweep (label-free) -> assignment -> sparse losses -> reduce -> finish;
  * training (inputs require grad, tools/scripts.py:893-945): assignment -> sparse losses (+ box /
    centre-ness gradients) -> reduce -> label-aware focal sweep writing d(cls) in the same pass ->
    reduce -> finish; backward() only multiplies the stored gradients by the upstream scalars.
There is no CPU path: CPU tensors raise.

Extra, keyword-only constructor arguments (defaults keep reference behaviour):
    sync_normalizer / process_group : all-reduce {positives, loss sums} over the process
        group so that an image-sharded batch reproduces the single-process full-batch loss
        (SURVEY.md section 8e).  Default False = the reference's per-rank normalisation.
        True = torch.distributed.all_reduce (NCCL); 'p2p' = the no-grad forward exchanges the 4
        doubles over NVLink peer memory inside its reduction kernel (b200det.peer, csrc/exchange.cu;
        one node, every GPU a peer of every other); the training path and the fused steps
        exchange their sums with the stand-alone b200det_sums_exchange kernel.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _handoff
from . import _lib
from . import geometry as _geom

__all__ = ['IoUMethod', 'RetinaLoss', 'FCOSLoss']

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream(device=None):
    return _lib.raw_stream(device)


class _on_device:
    """`with _on_device(t.device):` makes that GPU current for the C calls inside (kernel launches
    go to the CURRENT device; a process may drive several).  A no-op when it already is."""

    __slots__ = ('index', 'prev')

    def __init__(self, device):
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.index:
            self.prev = cur
            torch.cuda.set_device(self.index)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f'b200det: {what} must be a CUDA tensor (got device {t.device}); there is no CPU path')


def _prep_f32(levels, what):
    """float32, contiguous, 16-byte aligned per-level tensors (the heads already emit these)."""
    out = []
    for t in levels:
        if not t.is_cuda:
            _require_cuda(t, what)
        if t.dtype != torch.float32:
            t = t.float()
        if not t.is_contiguous():
            t = t.contiguous()
        if t.data_ptr() & 15:
            t = t.clone()
        out.append(t)
    return out


def _prep_reg(levels):
    dtype = levels[0].dtype
    if dtype not in _DTYPES or any(t.dtype != dtype for t in levels):
        levels = [t.float() for t in levels]
        dtype = torch.float32
    out = []
    for t in levels:
        if not t.is_cuda:
            _require_cuda(t, 'reg_preds')
        if not t.is_contiguous():
            t = t.contiguous()
        if t.data_ptr() & 15:
            t = t.clone()
        out.append(t)
    return out, _DTYPES[dtype]


def _loss_reg_mode(reg_dtype):
    """reg_dtype code for the LOSS kernels.  A half-precision regression head reaches the criterion
    in two ways in the reference: inside `with autocast()` (tools/scripts.py:886-893), where CUDA
    autocast runs torch.exp in float32 on the upcast value; or as plain half tensors (model.half()),
    where torch.exp rounds its result to half (losses.py:417-426, :568).  Both are reproduced."""
    if reg_dtype != _lib.F32 and not torch.is_autocast_enabled():
        return reg_dtype | _lib.REG_EXP_ROUNDED
    return reg_dtype


def _decode_reg_mode(reg_dtype):
    """reg_dtype code for the DECODERS: they leave torch (`.cpu().numpy()`, decode.py:208-219), so
    np.exp runs on the float16 array whatever the autocast state and rounds to half."""
    return reg_dtype | _lib.REG_EXP_ROUNDED if reg_dtype != _lib.F32 else reg_dtype


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


def _maybe_all_reduce(t, sync, group):
    if sync and torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)


def _sync_sums(owner, t, group, st, finish=None):
    """Sums `t` (device float64[4]) over the ranks in place: torch.distributed.all_reduce (NCCL), or,
    with sync_normalizer='p2p', ONE small kernel that exchanges the 4 doubles over NVLink peer
    memory (b200det_sums_exchange).  finish = (w_cls, w_box, w_ctr, losses_ptr) also normalises."""
    lib = _lib.load()
    if owner.sync_normalizer == 'p2p':
        px = _peer_exchange(owner, t.device)
        w = finish if finish is not None else (0., 0., 0., None)
        _lib.check(lib.b200det_sums_exchange(px.next(), w[0], w[1], w[2], t.data_ptr(), w[3], None,
                                             st), 'b200det_sums_exchange')
        return
    _maybe_all_reduce(t, True, group)
    if finish is not None:
        _lib.check(lib.b200det_loss_finish(t.data_ptr(), finish[0], finish[1], finish[2], finish[3],
                                           st), 'b200det_loss_finish')


def _peer_exchange(owner, device):
    """The owner's PeerExchange (created on first use: one all_gather_object + one barrier)."""
    px = getattr(owner, '_peer', None)
    if px is None:
        from .peer import PeerExchange
        px = PeerExchange(owner.process_group, device)
        owner._peer = px
    return px


class _SideStream:
    """The helper stream + fork / join events of b200det_loss_forward_overlap for one device,
    owned by the criterion object that uses them (the library keeps no streams itself)."""

    def __init__(self, device):
        lib = _lib.load()
        self._lib = lib
        self.stream, self.fork, self.join = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.b200det_stream_create(ctypes.byref(self.stream), _SIDE_HIGH), 'b200det_stream_create')
            _lib.check(lib.b200det_event_create(ctypes.byref(self.fork)), 'b200det_event_create')
            _lib.check(lib.b200det_event_create(ctypes.byref(self.join)), 'b200det_event_create')

    def __del__(self):
        try:
            for ev in (self.fork, self.join):
                if ev:
                    self._lib.b200det_event_destroy(ev)
            if self.stream:
                self._lib.b200det_stream_destroy(self.stream)
        except Exception:   # interpreter shutdown
            pass


_OVERLAP = os.environ.get('B200DET_LOSS_OVERLAP', '1') != '0'
_SIDE_HIGH = int(os.environ.get('B200DET_SIDE_PRIORITY', '1') != '0')


def _side_stream(owner, device):
    """The owner's side stream for `device` (created on first use), or None when switched off."""
    if not _OVERLAP:
        return None
    sides = owner.__dict__.get('_sides')
    if sides is None:
        sides = owner.__dict__['_sides'] = {}
    side = sides.get(device.index)
    if side is None:
        side = sides[device.index] = _SideStream(device)
    return side


class _Stateless:
    """Copy / pickle support: CUDA handles, peer mappings and cached plans are per-process runtime
    state and are rebuilt on first use."""

    def __getstate__(self):
        state = self.__dict__.copy()
        for key in ('_sides', '_peer', '_params_cache'):
            state.pop(key, None)
        state['_plans'] = {}
        state['last_stats'] = None
        return state


def _wants_grad(tensors):
    if not torch.is_grad_enabled():
        return False
    for t in tensors:
        if t.requires_grad:
            return True
    return False


class _Plan:
    """Everything about one (pyramid shapes, batch, classes) combination that does not change
    between calls: geometry struct, workspace size, row count."""

    __slots__ = ('geo', 'geo_ref', 'ws_bytes', 'n_rows', 'batch', 'scratch')

    def __init__(self, geo, batch, n_rows):
        self.geo = geo
        self.geo_ref = ctypes.byref(geo)
        self.batch = batch
        self.n_rows = n_rows
        self.ws_bytes = int(_lib.load().b200det_loss_workspace_bytes(self.geo_ref))
        self.scratch = {}   # (device, raw stream) -> workspace | labels of the no-grad forward

    def eval_scratch(self, device, stream):
        """Workspace + labels of the no-grad forward, kept per (device, stream): every call
        re-initialises what it reads, reuse on one stream is stream-ordered, and calls on different
        streams get different buffers.  (A fresh torch.empty per call cost ~3.5 us of the host time
        between the decoder's sync and the first launch of the next step.)"""
        key = (device.index, stream.value)
        buf = self.scratch.get(key)
        if buf is None:
            if len(self.scratch) >= 4:
                self.scratch.clear()
            buf = self.scratch[key] = torch.empty(self.ws_bytes + 4 * self.batch * self.n_rows,
                                                  dtype=torch.uint8, device=device)
        return buf


def _plan_for(owner, cls):
    shape0 = cls[0].shape
    key = tuple([t.shape for t in cls])
    plan = owner._plans.get(key)
    if plan is None:
        shapes = _geom.level_shapes(cls)
        batch, num_classes = int(shape0[0]), int(shape0[-1])
        geo = owner._geometry(shapes, batch, num_classes)
        plan = _Plan(geo, batch, _geom.rows_per_image(shapes, geo.per_loc))
        owner._plans = {key: plan}
    return plan


def _loss_params(owner, reg_dtype):
    cached = owner.__dict__.get('_params_cache')
    if cached is not None and cached[0] == (reg_dtype, owner.alpha, owner.gamma, owner.beta,
                                            owner.cls_loss_weight, owner.box_loss_weight):
        return cached[1]
    p = _lib.LossParams()
    p.is_fcos = int(owner._is_fcos)
    p.box_loss = owner._box_code
    p.reg_dtype = reg_dtype
    p.use_center_sample = int(getattr(owner, 'use_center_sample', 0))
    p.alpha = float(owner.alpha)
    p.gamma = float(owner.gamma)
    p.beta = float(owner.beta)
    p.w_cls = float(owner.cls_loss_weight)
    p.w_box = float(owner.box_loss_weight)
    p.w_ctr = float(getattr(owner, 'center_ness_loss_weight', 0.))
    p.iou_neg, p.iou_pos = owner._iou_thresholds
    if not hasattr(owner, 'center_ness_loss_weight'):   # FCOS has a third weight: not cached
        owner.__dict__['_params_cache'] = ((reg_dtype, owner.alpha, owner.gamma, owner.beta,
                                            owner.cls_loss_weight, owner.box_loss_weight), p)
    return p


def _handoff_for(owner, cls_in, reg_in, ctr_in, device, st):
    """The decoder that asked for its keys (b200det._handoff): (decoder, keys pointer, classes
    pointer, its score threshold, the tensors the keys will be derived from), or None."""
    shapes = tuple([t.shape for t in cls_in])
    dec = _handoff.offer(owner, device, st, shapes)
    if dec is None:
        return None
    target = dec._handoff_target(shapes, device, st)
    if target is None:
        return None
    tensors = list(cls_in) + (list(ctr_in) if ctr_in is not None else []) + list(reg_in)
    return dec, target[0], target[1], float(dec._params.min_score), tensors


def _forward_eval(owner, annotations, cls_in, reg_in, ctr_in):
    """No-grad forward: one C call (b200det_loss_forward_overlap)."""
    _require_cuda(cls_in[0], 'cls_preds')
    with _on_device(cls_in[0].device):
        return _forward_eval_on(owner, annotations, cls_in, reg_in, ctr_in)


def _forward_eval_fast(fast, owner, annotations, cls_in, reg_in, ctr_in):
    """The common case through csrc/fastpath.cpp: the same checks, pointer marshalling and C-ABI call
    as _forward_eval_on below, done in C++ (the host time in front of the first launch is GPU idle
    time in an eval loop).  Returns the losses tensor, or None when the inputs need the Python path."""
    if not (isinstance(annotations, torch.Tensor) and annotations.dim() == 3):
        return None
    try:
        plan = _plan_for(owner, cls_in)
    except Exception:   # noqa: BLE001 -- unusual inputs: let the Python path report them
        return None
    device = cls_in[0].device
    if device.type != 'cuda' or torch.cuda.is_current_stream_capturing():
        return None
    st = _stream(device)
    scratch = plan.eval_scratch(device, st)
    big = plan.batch * plan.n_rows * int(cls_in[0].shape[-1]) >= (64 << 20)
    side = _side_stream(owner, device) if big else None
    sync = owner.sync_normalizer and torch.distributed.is_available() \
        and torch.distributed.is_initialized()
    p2p = sync and owner.sync_normalizer == 'p2p'
    px = ctypes.addressof(_peer_exchange(owner, device).params) if p2p else 0
    hand = _handoff_for(owner, cls_in, reg_in, ctr_in, device, st) if _handoff.ENABLED else None
    params = (int(owner._is_fcos), owner._box_code, int(getattr(owner, 'use_center_sample', 0)),
              float(owner.alpha), float(owner.gamma), float(owner.beta), float(owner.cls_loss_weight),
              float(owner.box_loss_weight), float(getattr(owner, 'center_ness_loss_weight', 0.)),
              owner._iou_thresholds[0], owner._iou_thresholds[1])
    res = fast.loss_eval(ctypes.addressof(plan.geo), list(cls_in), list(reg_in),
                         list(ctr_in) if ctr_in is not None else None, annotations, params,
                         torch.is_autocast_enabled(), scratch.data_ptr(), plan.ws_bytes,
                         2 if p2p else (1 if sync else 0), px,
                         side.stream.value if side is not None else 0,
                         side.fork.value if side is not None else 0,
                         side.join.value if side is not None else 0, st.value or 0,
                         hand[3] if hand else 0., hand[1] if hand else 0, hand[2] if hand else 0)
    if res is None:
        return None
    if isinstance(res, int):
        _lib.check(res, 'b200det_loss_forward_overlap')
    if hand:
        _handoff.produced(hand[0], device, st, hand[4], hand[3])
    out = res
    if p2p:
        owner.__dict__['last_stats'] = {'sums': out[0:4],
                                        'exchange_status': out[6:7].view(torch.int32)[0:1]}
        return out[4:8].view(torch.float32)
    if sync:
        _maybe_all_reduce(out[0:4], True, owner.process_group)
        sums_ptr = out.data_ptr()
        _lib.check(
            _lib.load().b200det_loss_finish(sums_ptr, params[6], params[7], params[8], sums_ptr + 32,
                                            st), 'b200det_loss_finish')
    owner.__dict__['last_stats'] = {'sums': out[0:4]}
    return out[4:8].view(torch.float32)


def _forward_eval_on(owner, annotations, cls_in, reg_in, ctr_in):
    fast = _lib.fastpath()
    if fast is not None:
        res = _forward_eval_fast(fast, owner, annotations, cls_in, reg_in, ctr_in)
        if res is not None:
            return res
    lib = _lib.load()
    # Phase 1 as early as possible: the sweep only needs the classification tensors, so it is on the
    # GPU while the host still checks / marshals the other arguments (the eval loop calls this right
    # after the decoder's host sync: every microsecond before the first launch is GPU idle time).
    cls = _prep_f32(cls_in, 'cls_preds')
    plan = _plan_for(owner, cls)
    device = cls[0].device
    # scratch = workspace | labels ; out = sums (4 doubles) | losses (3 floats, 8 reserved)
    st = _stream(device)
    if torch.cuda.is_current_stream_capturing():
        scratch = torch.empty(plan.ws_bytes + 4 * plan.batch * plan.n_rows, dtype=torch.uint8,
                              device=device)   # from the graph's private pool
    else:
        scratch = plan.eval_scratch(device, st)
    ws_ptr = scratch.data_ptr()
    # assignment + sparse losses run on the criterion's side stream beside the HBM-bound sweep --
    # when there is a sweep worth hiding behind: below ~50 us of sweep (256 MB of class scores) the
    # call is bound by its host work, and the fork / join (4 more stream operations) and the second
    # C call only add to it
    big = plan.batch * plan.n_rows * int(cls[0].shape[-1]) >= (64 << 20)
    side = _side_stream(owner, device) if big else None
    side_args = (side.stream, side.fork, side.join) if side is not None else (None, None, None)
    params = _loss_params(owner, _lib.F32)
    cls_ptrs = _lib.ptr_array(cls)
    # a decoder waiting for keys of these head outputs: the sweep does its work as well (_handoff.py)
    hand = None
    if _handoff.ENABLED and not torch.cuda.is_current_stream_capturing():
        hand = _handoff_for(owner, cls_in, reg_in, ctr_in, device, st)
    ctr = None
    if hand:
        ctr = _prep_f32(ctr_in, 'center_preds') if ctr_in is not None else None
        forward = lib.b200det_loss_forward_keys
        keys_args = (hand[3], hand[1], hand[2])
    else:
        forward = lib.b200det_loss_forward_overlap
        keys_args = ()
    if big:
        _lib.check(
            forward(plan.geo_ref, ctypes.byref(params), None, 0, cls_ptrs, None,
                    _lib.ptr_array(ctr) if hand else None, None, ws_ptr, plan.ws_bytes, None, None, None,
                    None, *side_args, st, 1, *keys_args), 'b200det_loss_forward_overlap')
    reg, reg_dtype = _prep_reg(reg_in)
    reg_dtype = _loss_reg_mode(reg_dtype)
    if ctr is None:
        ctr = _prep_f32(ctr_in, 'center_preds') if ctr_in is not None else None
    annotations = _prep_annotations(annotations)
    if annotations.shape[0] != plan.batch:
        raise ValueError('annotations and predictions disagree on the batch size')
    out = torch.empty(8, dtype=torch.float64, device=device)
    sums_ptr = out.data_ptr()
    sync = owner.sync_normalizer and torch.distributed.is_available() \
        and torch.distributed.is_initialized()
    if reg_dtype != _lib.F32:
        params = _loss_params(owner, reg_dtype)
    p2p = sync and owner.sync_normalizer == 'p2p'
    px = status = None
    if p2p:
        # reduce + exchange over NVLink peer memory + normalisation in ONE kernel (csrc/exchange.cu)
        px = _peer_exchange(owner, device).next()
        status = out[6:7].view(torch.int32)[0:1]   # written by the kernel on every call
    _lib.check(
        forward(plan.geo_ref, ctypes.byref(params), annotations.data_ptr(),
                int(annotations.shape[1]), cls_ptrs, _lib.ptr_array(reg), _lib.ptr_array(ctr),
                ws_ptr + plan.ws_bytes, ws_ptr, plan.ws_bytes, px, sums_ptr,
                None if (sync and not p2p) else sums_ptr + 32,
                status.data_ptr() if p2p else None, *side_args, st, 2 if big else 0, *keys_args),
        'b200det_loss_forward_overlap')
    if hand:
        _handoff.produced(hand[0], device, st, hand[4], hand[3])
    if p2p:
        owner.__dict__['last_stats'] = {'sums': out[0:4], 'exchange_status': status}
        return out[4:8].view(torch.float32)
    if sync:
        _maybe_all_reduce(out[0:4], True, owner.process_group)
        _lib.check(
            lib.b200det_loss_finish(sums_ptr, params.w_cls, params.w_box, params.w_ctr,
                                    sums_ptr + 32, st), 'b200det_loss_finish')
    owner.__dict__['last_stats'] = {'sums': out[0:4]}   # not through nn.Module.__setattr__ (6 us)
    return out[4:8].view(torch.float32)


class _DetLossFunction(torch.autograd.Function):
    """Training path: one autograd node per loss call.  Gradients are produced by the forward
    kernels (the focal gradient already scaled by weight / positives) and only multiplied by the
    upstream scalars in backward.  backward() may be called once (the stored gradients are
    scaled in place)."""

    @staticmethod
    def forward(ctx, owner, annotations, n_levels, *heads):
        _require_cuda(heads[0], 'cls_preds')
        with _on_device(heads[0].device):
            return _DetLossFunction._forward(ctx, owner, annotations, n_levels, *heads)

    @staticmethod
    def _forward(ctx, owner, annotations, n_levels, *heads):
        lib = _lib.load()
        is_fcos = owner._is_fcos
        cls = _prep_f32(heads[0:n_levels], 'cls_preds')
        reg, reg_dtype = _prep_reg(heads[n_levels:2 * n_levels])
        reg_dtype = _loss_reg_mode(reg_dtype)
        ctr = _prep_f32(heads[2 * n_levels:3 * n_levels], 'center_preds') if is_fcos else None
        need = ctx.needs_input_grad[3:]
        annotations = _prep_annotations(annotations)
        plan = _plan_for(owner, cls)
        if annotations.shape[0] != plan.batch:
            raise ValueError('annotations and predictions disagree on the batch size')
        device = cls[0].device
        geo, ws_bytes = plan.geo_ref, plan.ws_bytes
        st = _stream(device)
        max_gt = int(annotations.shape[1])
        alpha, gamma = float(owner.alpha), float(owner.gamma)
        w_cls = float(owner.cls_loss_weight)
        w_box = float(owner.box_loss_weight)
        w_ctr = float(getattr(owner, 'center_ness_loss_weight', 0.))

        scratch = torch.empty(ws_bytes + 4 * plan.batch * plan.n_rows, dtype=torch.uint8,
                              device=device)           # workspace | labels
        ws_ptr = scratch.data_ptr()
        labels_ptr = ws_ptr + ws_bytes
        out = torch.zeros(12, dtype=torch.float64, device=device)   # sums | focal sums | losses
        sums, focal = out[0:4], out[4:8]
        losses = out[8:12].view(torch.float32)

        def flat_like(levels, dtype, zero):
            """one allocation for all levels' gradients; returns (flat, per-level views)"""
            sizes = [t.numel() for t in levels]
            flat = (torch.zeros if zero else torch.empty)(sum(sizes), dtype=dtype, device=device)
            views = [v.view(t.shape) for v, t in zip(flat.split(sizes), levels)]
            return flat, views

        # the sparse kernel writes the rows of the positives only -> zero-initialised
        _, reg_grad = flat_like(reg, torch.float32, True)
        ctr_grad = flat_like(ctr, torch.float32, True)[1] if is_fcos else None
        _, cls_grad = flat_like(cls, torch.float32, False)

        sync = owner.sync_normalizer and torch.distributed.is_available() \
            and torch.distributed.is_initialized()
        if not sync:
            params = _loss_params(owner, reg_dtype)
            # the sparse losses (box / centre-ness terms + their gradients) beside the sweep
            big = plan.batch * plan.n_rows * int(cls[0].shape[-1]) >= (64 << 20)
            side = _side_stream(owner, device) if big else None
            side_args = (side.stream, side.fork, side.join) if side is not None else (None, None, None)
            _lib.check(
                lib.b200det_loss_forward_grad_overlap(geo, ctypes.byref(params), annotations.data_ptr(),
                                                      max_gt, _lib.ptr_array(cls), _lib.ptr_array(reg),
                                                      _lib.ptr_array(ctr), labels_ptr,
                                                      _lib.ptr_array(cls_grad), _lib.ptr_array(reg_grad),
                                                      _lib.ptr_array(ctr_grad), ws_ptr, ws_bytes,
                                                      sums.data_ptr(), losses.data_ptr(), *side_args, st),
                'b200det_loss_forward_grad_overlap')
        else:
            group = owner.process_group
            if is_fcos:
                _lib.check(
                    lib.b200det_fcos_assign(geo, annotations.data_ptr(), max_gt,
                                            int(owner.use_center_sample), labels_ptr, None, None,
                                            ws_ptr, ws_bytes, st), 'b200det_fcos_assign')
            else:
                _lib.check(
                    lib.b200det_retina_assign(geo, annotations.data_ptr(), max_gt,
                                              owner._iou_thresholds[0], owner._iou_thresholds[1],
                                              labels_ptr, None, ws_ptr, ws_bytes, st),
                    'b200det_retina_assign')
            _lib.check(
                lib.b200det_sparse_losses(geo, int(is_fcos), annotations.data_ptr(), max_gt,
                                          labels_ptr, _lib.ptr_array(reg), reg_dtype,
                                          _lib.ptr_array(ctr), owner._box_code,
                                          float(owner.beta), None, alpha, gamma,
                                          _lib.ptr_array(reg_grad), _lib.ptr_array(ctr_grad),
                                          ws_ptr, ws_bytes, st), 'b200det_sparse_losses')
            # the focal gradient is written once, already divided by the GLOBAL positive count
            _lib.check(lib.b200det_loss_reduce(geo, 1, ws_ptr, ws_bytes, sums.data_ptr(), st),
                       'b200det_loss_reduce')
            _sync_sums(owner, sums, group, st)
            _lib.check(
                lib.b200det_focal_loss(geo, _lib.ptr_array(cls), labels_ptr, alpha, gamma,
                                       _lib.ptr_array(cls_grad), sums.data_ptr(), w_cls, ws_ptr,
                                       ws_bytes, st), 'b200det_focal_loss')
            _lib.check(lib.b200det_loss_reduce(geo, 2, ws_ptr, ws_bytes, focal.data_ptr(), st),
                       'b200det_loss_reduce')
            _sync_sums(owner, focal, group, st)
            sums.add_(focal)
            _lib.check(lib.b200det_loss_finish(sums.data_ptr(), w_cls, w_box, w_ctr,
                                               losses.data_ptr(), st), 'b200det_loss_finish')

        ctx.n_levels = n_levels
        ctx.is_fcos = is_fcos
        ctx.in_dtypes = [h.dtype for h in heads]
        ctx.in_shapes = [h.shape for h in heads]
        ctx.want = (any(need[0:n_levels]), any(need[n_levels:2 * n_levels]),
                    is_fcos and any(need[2 * n_levels:3 * n_levels]))
        ctx.weights = (w_box, w_ctr)
        ctx.plan = plan
        ctx.scratch = scratch   # keeps the positive-row queue alive for backward
        ctx.consumed = False
        ctx.set_materialize_grads(False)   # an unused loss term arrives as None, not as zeros
        ctx.save_for_backward(sums, *cls_grad, *reg_grad, *(ctr_grad or []))
        owner.last_stats = {'sums': sums}
        return (losses[0], losses[1], losses[2]) if is_fcos else (losses[0], losses[1])

    @staticmethod
    def backward(ctx, *grad_out):
        # The stored gradients are scaled IN PLACE by the upstream scalars (no second copy of a
        # cls-sized tensor): a second pass over the same node would scale them twice, or resurrect
        # terms a first pass multiplied by an upstream zero.  Refuse instead of returning wrong
        # gradients (torch does the same for a freed graph); call backward once on the summed loss,
        # as tools/scripts.py:918-945 does, or re-run the forward.
        if ctx.consumed:
            raise RuntimeError(
                'b200det: backward through this loss call a second time is not supported (its '
                'gradients are produced by the forward kernels and scaled in place once); sum the '
                'loss terms and call backward once, or run the forward again')
        ctx.consumed = True
        saved = ctx.saved_tensors
        with _on_device(saved[0].device):
            return _DetLossFunction._backward(ctx, saved, *grad_out)

    @staticmethod
    def _backward(ctx, saved, *grad_out):
        lib = _lib.load()
        n = ctx.n_levels
        sums = saved[0]
        cls_grad = saved[1:1 + n]
        reg_grad = saved[1 + n:1 + 2 * n]
        ctr_grad = saved[1 + 2 * n:1 + 3 * n]
        want_cls, want_reg, want_ctr = ctx.want
        st = _stream(sums.device)
        grads = [None] * (3 * n if ctx.is_fcos else 2 * n)
        # a loss term nobody differentiated (grad None): its head gets no gradient from this node
        # unless another term feeds it -- cls only feeds cls_loss, ctr only centre-ness; reg feeds
        # reg_loss only
        want_cls = want_cls and grad_out[0] is not None
        want_reg = want_reg and grad_out[1] is not None
        want_ctr = want_ctr and len(grad_out) > 2 and grad_out[2] is not None

        def scale(levels, g, with_norm, weight):
            """levels[l] *= g * (weight / positives if with_norm else 1): one launch"""
            counts = (ctypes.c_longlong * n)(*[t.numel() for t in levels])
            g = g.detach().float().contiguous()
            _lib.check(
                lib.b200det_scale_levels(_lib.ptr_array(levels), counts, n, g.data_ptr(),
                                         sums.data_ptr() if with_norm else None, weight, st),
                'b200det_scale_levels')

        def as_input(t, k):
            """gradient in the input's shape / dtype (no-ops for the usual float32 heads)"""
            if t.shape != ctx.in_shapes[k]:
                t = t.view(ctx.in_shapes[k])
            return t if t.dtype == ctx.in_dtypes[k] else t.to(ctx.in_dtypes[k])

        if want_cls:
            # already scaled by cls_loss_weight / positives; only the upstream scalar is left
            scale(cls_grad, grad_out[0], False, 1.0)
            for i in range(n):
                grads[i] = as_input(cls_grad[i], i)
        if want_reg or want_ctr:
            # only the positives' rows hold gradients: scale those rows (the assignment's queue is
            # still in the forward's workspace), not the whole [B*N, 4] tensors
            g_box = grad_out[1].detach().float().contiguous() if want_reg else None
            g_ctr = grad_out[2].detach().float().contiguous() if want_ctr else None
            _lib.check(
                lib.b200det_scale_pos_rows(ctx.plan.geo_ref, ctx.scratch.data_ptr(), ctx.plan.ws_bytes,
                                           _lib.ptr_array(list(reg_grad)) if want_reg else None,
                                           _lib.ptr_array(list(ctr_grad)) if want_ctr else None,
                                           g_box.data_ptr() if want_reg else None,
                                           g_ctr.data_ptr() if want_ctr else None, sums.data_ptr(),
                                           ctx.weights[0], ctx.weights[1], st),
                'b200det_scale_pos_rows')
        if want_reg:
            for i in range(n):
                grads[n + i] = as_input(reg_grad[i], n + i)
        if want_ctr:
            for i in range(n):
                grads[2 * n + i] = as_input(ctr_grad[i], 2 * n + i)
        return (None, None, None, *grads)


def _debug_assign(owner, preds, annotations, exact=True):
    """Parity hook (tests / smoke): runs only the assignment kernel and returns the reference's
    intermediate truth in IMAGE-major order: labels [B,N] int32, matched [B,N] int32 (index in
    the image's filtered GT list; -1 = none) and, for FCOS, targets [B,N,6] float32.
    exact=False runs the production scan (no `matched` output: pairs that cannot reach IoU 0.38
    are dropped early) and returns the labels only."""
    with _on_device(preds[0][0].device):
        return _debug_assign_on(owner, preds, annotations, exact)


def _debug_assign_on(owner, preds, annotations, exact):
    lib = _lib.load()
    is_fcos = owner._is_fcos
    cls = _prep_f32(preds[0], 'cls_preds')
    annotations = _prep_annotations(annotations)
    plan = _plan_for(owner, cls)
    device = cls[0].device
    batch, n_rows, geo, ws_bytes = plan.batch, plan.n_rows, plan.geo_ref, plan.ws_bytes
    st = _stream(device)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    labels = torch.empty(batch * n_rows, dtype=torch.int32, device=device)
    matched = torch.empty(batch * n_rows, dtype=torch.int32, device=device) if exact else None
    matched_ptr = matched.data_ptr() if exact else None
    targets = None
    if is_fcos:
        targets = torch.empty(batch * n_rows * 6, dtype=torch.float32, device=device)
        _lib.check(
            lib.b200det_fcos_assign(geo, annotations.data_ptr(), int(annotations.shape[1]),
                                    int(owner.use_center_sample), labels.data_ptr(), matched_ptr,
                                    targets.data_ptr(), ws.data_ptr(), ws_bytes, st),
            'b200det_fcos_assign')
    else:
        _lib.check(
            lib.b200det_retina_assign(geo, annotations.data_ptr(), int(annotations.shape[1]),
                                      owner._iou_thresholds[0], owner._iou_thresholds[1],
                                      labels.data_ptr(), matched_ptr, ws.data_ptr(), ws_bytes,
                                      st), 'b200det_retina_assign')

    def to_image_major(t, width):
        out = torch.empty_like(t)
        _lib.check(
            lib.b200det_rows_to_image_major(geo, t.data_ptr(), out.data_ptr(), width, st),
            'b200det_rows_to_image_major')
        return out

    res = {'labels': to_image_major(labels, 1).view(batch, n_rows)}
    if exact:
        res['matched'] = to_image_major(matched, 1).view(batch, n_rows)
    if targets is not None:
        res['targets'] = to_image_major(targets, 6).view(batch, n_rows, 6)
    return res


class _IoUFunction(torch.autograd.Function):
    """out[n, m] = iou_type(boxes1 row n*s1[0] + m*s1[1], boxes2 row n*s2[0] + m*s2[1]); the kernel
    also writes the Jacobians w.r.t. the inputs that require grad."""

    @staticmethod
    def forward(ctx, boxes1, boxes2, n, m, s1, s2, code, xywh):
        device = boxes1.device
        out = torch.empty((n, m), dtype=torch.float32, device=device)
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        jac1 = torch.empty((n, m, 4), dtype=torch.float32, device=device) if need1 else None
        jac2 = torch.empty((n, m, 4), dtype=torch.float32, device=device) if need2 else None
        with torch.cuda.device(device):
            _lib.check(
                _lib.load().b200det_iou_method(
                    boxes1.data_ptr(), s1[0], s1[1], boxes2.data_ptr(), s2[0], s2[1], n, m, code,
                    int(xywh), out.data_ptr(), jac1.data_ptr() if need1 else None,
                    jac2.data_ptr() if need2 else None, _stream(device)), 'b200det_iou_method')
        ctx.save_for_backward(*[j for j in (jac1, jac2) if j is not None])
        ctx.need = (need1, need2)
        ctx.shapes = (boxes1.shape, boxes2.shape)
        ctx.bcast = (s1, s2)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        saved = list(ctx.saved_tensors)
        grads = [None, None]
        g = grad_out.float().unsqueeze(-1)
        for k in range(2):
            if not ctx.need[k]:
                continue
            full = saved.pop(0) * g                      # [n, m, 4]
            sn, sm = ctx.bcast[k]
            if sn == 0:
                full = full.sum(0, keepdim=True)
            if sm == 0 and full.shape[1] != 1:
                full = full.sum(1, keepdim=True)
            grads[k] = full.reshape(ctx.shapes[k])
        return (grads[0], grads[1], None, None, None, None, None, None)


class IoUMethod:
    """simpleAICV/detection/losses.py:28-123 on the GPU: IoU / GIoU / DIoU /
    CIoU / EIoU between `boxes1` and `boxes2` ([..., 4], 'xyxy' or 'xywh'), float32 with the
    reference's op order, differentiable w.r.t. both.  Same constructor, call signature and asserts.
    Element-wise inputs and the assignment's [N,1,4] x [1,M,4] broadcast (losses.py:350-353) are
    computed in place; other broadcasts are expanded first.  Like the reference's indexing
    (`[:, 0]`, `torch.cat(dim=1)`), GIoU / CIoU and box_type 'xywh' need 2-D [N, 4] inputs.
    CUDA tensors only; there is no fallback."""

    _CODES = {'IoU': _lib.BOX_IOU, 'GIoU': _lib.BOX_GIOU, 'DIoU': _lib.BOX_DIOU,
              'CIoU': _lib.BOX_CIOU, 'EIoU': _lib.BOX_EIOU}

    def __init__(self):
        pass

    def __call__(self, boxes1, boxes2, iou_type='IoU', box_type='xyxy'):
        assert iou_type in ['IoU', 'GIoU', 'DIoU', 'CIoU', 'EIoU'], 'wrong IoU type!'
        assert box_type in ['xyxy', 'xywh'], 'wrong box_type type!'
        _require_cuda(boxes1, 'boxes1')
        _require_cuda(boxes2, 'boxes2')
        if boxes1.device != boxes2.device:
            raise RuntimeError('boxes1 and boxes2 must be on the same CUDA device')
        if boxes1.shape[-1] != 4 or boxes2.shape[-1] != 4:
            raise RuntimeError('boxes must be [..., 4]')
        if (iou_type in ('GIoU', 'CIoU') or box_type == 'xywh') and \
                (boxes1.dim() != 2 or boxes2.dim() != 2):
            raise RuntimeError(f'{iou_type} / {box_type} need 2-D [N, 4] boxes, as in the reference '
                               '(losses.py:46-52, 99-100, 118-120)')
        boxes1, boxes2 = boxes1.float(), boxes2.float()
        lead1, lead2 = tuple(boxes1.shape[:-1]), tuple(boxes2.shape[:-1])
        out_shape = torch.broadcast_shapes(lead1, lead2)
        if len(lead1) == 2 and len(lead2) == 2 and lead1[1] == 1 and lead2[0] == 1:
            n, m, s1, s2 = lead1[0], lead2[1], (1, 0), (0, 1)        # [N,1,4] x [1,M,4]
        elif len(lead1) == 2 and len(lead2) == 2 and lead1[0] == 1 and lead2[1] == 1:
            n, m, s1, s2 = lead2[0], lead1[1], (0, 1), (1, 0)        # [1,M,4] x [N,1,4]
        else:
            if lead1 != out_shape:
                boxes1 = boxes1.expand(out_shape + (4,))
            if lead2 != out_shape:
                boxes2 = boxes2.expand(out_shape + (4,))
            n, m, s1, s2 = 1, 1, (1, 0), (1, 0)
            for d in out_shape:
                n *= d
        out = _IoUFunction.apply(boxes1.contiguous(), boxes2.contiguous(), n, m, s1, s2,
                                 self._CODES[iou_type], box_type == 'xywh')
        return out.view(out_shape)


class RetinaLoss(_Stateless, nn.Module):
    """Drop-in for simpleAICV.detection.losses.RetinaLoss (losses.py:126-429)."""

    _is_fcos = False
    _iou_thresholds = (0.4, 0.5)   # losses.py:361-365

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
        self._plans = {}
        self.last_stats = None

    def _geometry(self, shapes, batch, num_classes):
        if len(shapes) > len(self.areas):
            raise ValueError('more pyramid levels than anchor areas')
        return _geom.make_geometry(shapes, batch, self._per_loc, num_classes, self.strides,
                                   base_anchors=self._base)

    def debug_assign(self, preds, annotations, exact=True):
        return _debug_assign(self, preds, annotations, exact)

    def forward(self, preds, annotations):
        '''
        compute cls loss and reg loss in one batch
        '''
        cls_preds, reg_preds = preds
        n = len(cls_preds)
        assert len(reg_preds) == n
        if _wants_grad(cls_preds) or _wants_grad(reg_preds):
            cls_loss, reg_loss = _DetLossFunction.apply(self, annotations, n, *cls_preds,
                                                        *reg_preds)
        else:
            losses = _forward_eval(self, annotations, cls_preds, reg_preds, None)
            cls_loss, reg_loss = losses[0], losses[1]
        loss_dict = {
            'cls_loss': cls_loss,
            'reg_loss': reg_loss,
        }
        return loss_dict


class FCOSLoss(_Stateless, nn.Module):
    """Drop-in for simpleAICV.detection.losses.FCOSLoss (losses.py:432-833)."""

    _is_fcos = True
    _iou_thresholds = (0.4, 0.5)   # unused by the point assignment

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
        self._plans = {}
        self.last_stats = None

    def _geometry(self, shapes, batch, num_classes):
        if len(shapes) > len(self.mi):
            raise ValueError('more pyramid levels than mi ranges')
        return _geom.make_geometry(shapes, batch, 1, num_classes, self.strides, mi=self.mi,
                                   center_sample_radius=self.center_sample_radius)

    def debug_assign(self, preds, annotations, exact=True):
        return _debug_assign(self, preds, annotations, exact)

    def forward(self, preds, annotations):
        '''
        compute cls loss, reg loss and center-ness loss in one batch
        '''
        cls_preds, reg_preds, center_preds = preds
        n = len(cls_preds)
        assert len(reg_preds) == n and len(center_preds) == n
        if _wants_grad(cls_preds) or _wants_grad(reg_preds) or _wants_grad(center_preds):
            cls_loss, reg_loss, center_ness_loss = _DetLossFunction.apply(
                self, annotations, n, *cls_preds, *reg_preds, *center_preds)
        else:
            losses = _forward_eval(self, annotations, cls_preds, reg_preds, center_preds)
            cls_loss, reg_loss, center_ness_loss = losses[0], losses[1], losses[2]
        loss_dict = {
            'cls_loss': cls_loss,
            'reg_loss': reg_loss,
            'center_ness_loss': center_ness_loss,
        }
        return loss_dict
