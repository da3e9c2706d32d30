"""b200det.decode -- drop-in RetinaDecoder / FCOSDecoder (and the query-based DETRDecoder /
DINODETRDecoder plus the stand-alone DecodeMethod / DetNMSMethod) backed by libb200det.so
(sm_100a CUDA).

Same class names, constructor kwargs, __call__(preds) signature and return value (a list of
three writable float32 NumPy arrays: scores [B,M] padded with -1, classes [B,M] padded with -1,
boxes [B,M,4] padded with 0) as simpleAICV/detection/decode.py:175-271 (RetinaDecoder) and
:274-364 (FCOSDecoder).  The reference copies every head output to the host and decodes in NumPy;
here the head outputs stay in HBM, one C call (b200det_decode) enqueues two kernels
    b200det_score_argmax        arg-max / score / threshold: one streaming pass over cls
    b200det_select_decode_nms   per-image top-n, box decode, NMS, max_object_num cap
and only the [B, M, 6] result (2.4 KB per image) crosses PCIe.  There is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from . import _handoff
from . import _lib
from . import geometry as _geom
from .losses import _decode_reg_mode, _on_device, _prep_f32, _prep_reg, _require_cuda

_ZERO_COPY = os.environ.get('B200DET_ZERO_COPY', '1') != '0'
# the returned arrays are views of the call's own pinned buffer (no host memcpy); '1' hands out
# pageable copies instead (for callers that keep thousands of results alive)
_RESULT_COPY = os.environ.get('B200DET_RESULT_COPY', '0') != '0'

__all__ = ['RetinaDecoder', 'FCOSDecoder', 'DETRDecoder', 'DINODETRDecoder', 'DecodeMethod',
           'DetNMSMethod']


def _detached(levels):
    """The level tensors without autograd history (the decoders never differentiate); tensors that
    do not require grad are passed through -- `.detach()` costs a microsecond per tensor of the host
    time in front of the decoder's first launch."""
    return [t.detach() if t.requires_grad else t for t in levels]


_HALF_EXP_TABLES = {}


def _half_exp_table(device):
    """Device copy of np.exp over all 65536 float16 inputs AS THIS HOST'S NumPy computes it, or None.
    The reference's decoders run np.exp on the float16 regression array (decode.py:260, :356) and
    NumPy's float16 exp depends on the CPU: with AVX512-FP16 ("AVX512_SPR" in NumPy's dispatch) it is
    an SVML kernel whose result differs from the correctly rounded one for 17 % of the inputs;
    elsewhere it is half(expf(float(x))), which the kernel computes itself (table = None)."""
    key = device.index
    if key not in _HALF_EXP_TABLES:
        table = None
        try:
            try:
                from numpy._core._multiarray_umath import __cpu_features__ as feats
            except ImportError:   # numpy < 2
                from numpy.core._multiarray_umath import __cpu_features__ as feats
            if feats.get('AVX512_SPR'):
                path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data',
                                    'np_exp_f16_avx512spr.npy')
                host = np.load(path).astype(np.int16)
                table = torch.from_numpy(host).to(device)
                # complete before anything uses it: a select kernel on handed-over keys does not wait
                # for what precedes it on the stream (b200det_decode_from_keys: inputs_complete)
                torch.cuda.synchronize(device)
        except Exception:   # no dispatch info: the generic half loop
            table = None
        _HALF_EXP_TABLES[key] = table
    return _HALF_EXP_TABLES[key]


class _DecoderBase:
    _is_fcos = False

    def _init_common(self, max_object_num, min_score_threshold, topn, nms_type, nms_threshold):
        assert nms_type in ['torch_nms', 'python_nms', 'diou_python_nms'], 'wrong nms type!'
        if topn > _lib.MAX_TOPN:
            raise ValueError(f'topn <= {_lib.MAX_TOPN} is supported')
        self.max_object_num = max_object_num
        self.min_score_threshold = min_score_threshold
        self.topn = topn
        self.nms_type = nms_type
        self.nms_threshold = nms_threshold
        self._nms_code = _lib.NMS_CODES[nms_type]
        self._geo_cache = {}
        self._pinned = None
        p = _lib.DecodeParams()
        p.is_fcos = int(self._is_fcos)
        p.reg_dtype = _lib.F32
        p.topn = int(topn)
        p.max_out = int(max_object_num)
        p.nms_type = self._nms_code
        p.min_score = float(np.float32(min_score_threshold))
        p.nms_threshold = float(nms_threshold)
        self._params = p

    def _staging(self, numel):
        if self._pinned is None or self._pinned.numel() != numel:
            self._pinned = torch.empty(numel, dtype=torch.float32, pin_memory=True)
        return self._pinned

    @staticmethod
    def _set_glue(params, batch, device, scales, sizes, to_xywh):
        """Fills the optional evaluation-glue fields of the decode params; returns the small
        device tensors, which must stay alive until the call has been enqueued."""
        glue = []
        params.scales = params.sizes = None
        params.to_xywh = int(bool(to_xywh))
        if scales is not None:
            t = torch.as_tensor(np.asarray(scales, dtype=np.float32).reshape(-1)).to(device)
            if t.numel() != batch:
                raise ValueError('scales must have one entry per image')
            glue.append(t)
            params.scales = t.data_ptr()
        if sizes is not None:
            t = torch.as_tensor(np.asarray(sizes, dtype=np.float32).reshape(-1)).to(device)
            if t.numel() != 2 * batch:
                raise ValueError('sizes must be [B, 2] = (height, width) per image')
            glue.append(t)
            params.sizes = t.data_ptr()
        return glue

    def _out_buffer(self, numel, device):
        """Where the select kernel writes scores | classes | boxes.  Default: straight into the
        cached pinned host buffer (mapped into the device's address space under UVA), so the 24*M
        bytes per image cross PCIe as posted writes while other images are still being processed
        and no separate copy is enqueued.  B200DET_ZERO_COPY=0 uses a device buffer + one D2H."""
        if _ZERO_COPY:
            if _RESULT_COPY:
                return self._staging(numel)
            # a fresh block from torch's caching pinned allocator per call (reused once the
            # previous result has been dropped): the caller owns it through the returned arrays
            return torch.empty(numel, dtype=torch.float32, pin_memory=True)
        return torch.empty(numel, dtype=torch.float32, device=device)

    def _handoff_target(self, shapes, device, stream):
        """Where a criterion's fused sweep may leave this decoder's keys / classes (b200det._handoff):
        the scratch of the plan for these level shapes on (device, stream), or None before the
        decoder's first own call there."""
        plan = self._geo_cache.get(shapes)
        if plan is None or not _ZERO_COPY:
            return None
        scratch = plan[5].get((device.index, stream.value))
        if scratch is None:
            return None
        base = scratch.data_ptr()
        return base, base + 4 * plan[2] * plan[3]

    def _to_host(self, out, batch, m, device):
        """The only device->host traffic: 24*M bytes per image into a cached pinned buffer; the
        caller gets fresh, writable arrays (tools/scripts.py:742-758 mutates them in place)."""
        staging = out
        if out.is_cuda:
            staging = self._staging(out.numel())
            staging.copy_(out, non_blocking=True)
        copy = _RESULT_COPY or staging is self._pinned
        result = None
        if not copy:
            # the views are built while the GPU is still working: nothing but the return is left
            # between the end of the synchronisation and the caller's next launch
            host = staging.numpy()
            result = [host[0:batch * m].reshape(batch, m),
                      host[batch * m:2 * batch * m].reshape(batch, m),
                      host[2 * batch * m:6 * batch * m].reshape(batch, m, 4)]
        _lib.check(_lib.load().b200det_stream_synchronize(_lib.raw_stream(device)),
                   'b200det_stream_synchronize')
        if result is None:
            host = staging.numpy().copy()
            result = [host[0:batch * m].reshape(batch, m),
                      host[batch * m:2 * batch * m].reshape(batch, m),
                      host[2 * batch * m:6 * batch * m].reshape(batch, m, 4)]
        return result

    def _run(self, preds, details=False, scales=None, sizes=None, to_xywh=False):
        _require_cuda(preds[0][0], 'cls_preds')
        with _on_device(preds[0][0].device):
            return self._run_on(preds, details, scales, sizes, to_xywh)

    def _handed_over(self, cls_preds, reg_preds, center_preds, device, st):
        """True when the criterion's sweep already left this call's keys in the scratch
        (b200det._handoff)."""
        if not (_handoff.ENABLED and _ZERO_COPY):
            return False
        tensors = list(cls_preds) + (list(center_preds) if center_preds is not None else []) + \
            list(reg_preds)
        return _handoff.take(self, device, st, tensors, float(self._params.min_score))

    def _wish(self, cls_preds, device, st):
        """After a decode that swept for itself AND succeeded (so the class count / shapes are ones the
        sweeps support -- the fused sweep has the decoder's limits): ask the next criterion call on
        head outputs of these shapes to produce the keys."""
        if _handoff.ENABLED and _ZERO_COPY:
            _handoff.wish(self, device, st, tuple([t.shape for t in cls_preds]),
                          int(cls_preds[0].shape[-1]))

    @staticmethod
    def _stale(out, index):
        """After a hand-over (and the stream sync): did the select kernel find a selected row whose
        key no longer matches the class score it names (head outputs modified behind torch's back)?
        `out` is the call's pinned result block, `index` the position of its flag word."""
        return bool(out.numpy()[index] != 0.0)

    def _run_fast(self, fast, cls_preds, reg_preds, center_preds):
        """The plain decoder(preds) call through csrc/fastpath.cpp (same checks / marshalling / C-ABI
        call as _run_on, in C++); None when the inputs need the Python path."""
        try:
            key = tuple([t.shape for t in cls_preds])
            plan = self._geo_cache.get(key)
            if plan is None:
                return None   # first call with these shapes: the Python path builds the plan
            geo, _, batch, n_rows, ws_bytes, scratch_by_stream = plan
            device = cls_preds[0].device
        except Exception:   # noqa: BLE001
            return None
        if device.type != 'cuda':
            return None
        st = _lib.raw_stream(device)
        rows_bytes = (8 * batch * n_rows + 255) & ~255
        skey = (device.index, st.value)
        scratch = scratch_by_stream.get(skey)
        if scratch is None:
            return None
        table = 0
        if reg_preds[0].dtype == torch.float16:
            t = _half_exp_table(device)
            table = t.data_ptr() if t is not None else 0
        p = self._params
        handed = self._handed_over(cls_preds, reg_preds, center_preds, device, st)
        while True:
            res = fast.decode_run(ctypes.addressof(geo), list(cls_preds), list(reg_preds),
                                  list(center_preds) if center_preds is not None else None,
                                  (p.is_fcos, p.topn, p.max_out, p.nms_type, p.min_score,
                                   p.nms_threshold), scratch.data_ptr(), 4 * batch * n_rows, rows_bytes,
                                  ws_bytes, table, st.value or 0, handed)
            if res is None:
                return None
            if isinstance(res, int):
                _lib.check(res, 'b200det_decode')
            if not handed:
                self._wish(cls_preds, device, st)
            result = self._to_host(res, batch, int(self.max_object_num), device)
            if not handed or not self._stale(res, 6 * batch * int(self.max_object_num)):
                return result
            _handoff.stats['stale'] += 1   # the head outputs changed after the criterion's sweep
            handed = False

    def _run_on(self, preds, details, scales, sizes, to_xywh):
        lib = _lib.load()
        if self._is_fcos:
            cls_preds, reg_preds, center_preds = preds
        else:
            cls_preds, reg_preds = preds
            center_preds = None
        if not details and scales is None and sizes is None and not to_xywh and _ZERO_COPY \
                and not _RESULT_COPY:
            fast = _lib.fastpath()
            if fast is not None:
                res = self._run_fast(fast, cls_preds, reg_preds, center_preds)
                if res is not None:
                    return res
        cls = _prep_f32(_detached(cls_preds), 'cls_preds')
        reg, reg_dtype = _prep_reg(_detached(reg_preds))
        ctr = _prep_f32(_detached(center_preds), 'center_preds') \
            if center_preds is not None else None
        device = cls[0].device
        shape0 = cls[0].shape
        key = tuple([t.shape for t in cls])
        plan = self._geo_cache.get(key)
        if plan is None:
            shapes = _geom.level_shapes(cls)
            geo = self._geometry(shapes, int(shape0[0]), int(shape0[-1]))
            ws_bytes = int(lib.b200det_decode_workspace_bytes(ctypes.byref(geo), int(self.topn)))
            plan = (geo, ctypes.byref(geo), int(shape0[0]),
                    _geom.rows_per_image(shapes, geo.per_loc), ws_bytes, {})
            self._geo_cache = {key: plan}
        _, geo_ref, batch, n_rows, ws_bytes, scratch_by_stream = plan
        m = int(self.max_object_num)

        # scratch = keys | classes (int32 each) | selection workspace ; out = scores|classes|boxes
        # (kept per (device, stream), like the criterion's: every call rewrites all of it)
        rows_bytes = (8 * batch * n_rows + 255) & ~255
        st = _lib.raw_stream(device)
        skey = (device.index, st.value)
        scratch = scratch_by_stream.get(skey)
        if scratch is None:
            if len(scratch_by_stream) >= 4:
                scratch_by_stream.clear()
            scratch = scratch_by_stream[skey] = torch.empty(rows_bytes + ws_bytes, dtype=torch.uint8,
                                                            device=device)
        handed = self._handed_over(cls_preds, reg_preds, center_preds, device, st)
        out = self._out_buffer(6 * batch * m + 4, device)   # + the hand-over's `stale` word
        order = keep = counts = None
        if details:
            order = torch.empty(batch * self.topn, dtype=torch.int32, device=device)
            keep = torch.empty(batch * self.topn, dtype=torch.int32, device=device)
            counts = torch.empty(batch * 3, dtype=torch.int32, device=device)
        params = self._params
        params.reg_dtype = _decode_reg_mode(reg_dtype)
        params.half_exp_table = None
        if reg_dtype == _lib.F16:
            table = _half_exp_table(device)
            params.half_exp_table = table.data_ptr() if table is not None else None
        glue = self._set_glue(params, batch, device, scales, sizes, to_xywh)
        keys_ptr = scratch.data_ptr()
        while True:
            if handed:
                # keys / classes are in the scratch already (b200det_loss_forward_keys): select only,
                # verifying on the device that the selected rows still match the head outputs
                out[6 * batch * m:].zero_()
                _lib.check(
                    lib.b200det_decode_from_keys(geo_ref, ctypes.byref(params), _lib.ptr_array(cls),
                                                 _lib.ptr_array(ctr), _lib.ptr_array(reg), keys_ptr,
                                                 keys_ptr + 4 * batch * n_rows, out.data_ptr(),
                                                 order.data_ptr() if details else None,
                                                 keep.data_ptr() if details else None,
                                                 counts.data_ptr() if details else None,
                                                 out.data_ptr() + 24 * batch * m,
                                                 0 if glue else 1, st),
                    'b200det_decode_from_keys')
            else:
                _lib.check(
                    lib.b200det_decode(geo_ref, ctypes.byref(params), _lib.ptr_array(cls),
                                       _lib.ptr_array(ctr), _lib.ptr_array(reg), keys_ptr,
                                       keys_ptr + 4 * batch * n_rows, out.data_ptr(),
                                       order.data_ptr() if details else None,
                                       keep.data_ptr() if details else None,
                                       counts.data_ptr() if details else None,
                                       keys_ptr + rows_bytes, ws_bytes, st),
                    'b200det_decode')
                self._wish(cls_preds, device, st)
            result = self._to_host(out, batch, m, device)
            if not handed or not self._stale(out, 6 * batch * m):
                break
            _handoff.stats['stale'] += 1   # the head outputs changed after the criterion's sweep
            handed = False
            out = self._out_buffer(6 * batch * m + 4, device)
        del glue
        if not details:
            return result
        info = {
            'order': order.view(batch, self.topn).cpu().numpy(),
            'keep': keep.view(batch, self.topn).cpu().numpy(),
            'counts': counts.view(batch, 3).cpu().numpy(),
        }
        return result, info

    def __call__(self, preds, scales=None, sizes=None, to_xywh=False):
        """decoder(preds) is the reference call.  The optional arguments fuse the evaluation glue
        of the reference's test loop into the kernel epilogue (tools/scripts.py:742-757):
        `scales` [B] -> boxes /= scale; `sizes` [B,2] = (h, w) -> clip to the original image and,
        with to_xywh=True, convert x2,y2 to w,h (COCO json format)."""
        return self._run(preds, details=False, scales=scales, sizes=sizes, to_xywh=to_xywh)

    def decode_with_details(self, preds):
        """Parity hook: also returns per image the sorted top-n row indices ('order', -1 padded),
        the NMS keep positions ('keep', full list) and [candidates, selected, kept] counts."""
        return self._run(preds, details=True)


class RetinaDecoder(_DecoderBase):
    """Drop-in for simpleAICV.detection.decode.RetinaDecoder (decode.py:175-271)."""

    _is_fcos = False

    def __init__(self,
                 areas=[[32, 32], [64, 64], [128, 128], [256, 256], [512, 512]],
                 ratios=[0.5, 1, 2],
                 scales=[2**0, 2**(1.0 / 3.0), 2**(2.0 / 3.0)],
                 strides=[8, 16, 32, 64, 128],
                 max_object_num=100,
                 min_score_threshold=0.05,
                 topn=1000,
                 nms_type='python_nms',
                 nms_threshold=0.5):
        self._init_common(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)
        self.areas = areas
        self.ratios = ratios
        self.scales = scales
        self.strides = strides
        self._per_loc = len(ratios) * len(scales)
        self._base = _geom.retina_base_anchors(areas, ratios, scales)

    def _geometry(self, shapes, batch, num_classes):
        if len(shapes) > len(self.areas):
            raise ValueError('more pyramid levels than anchor areas')
        return _geom.make_geometry(shapes, batch, self._per_loc, num_classes, self.strides,
                                   base_anchors=self._base)


class FCOSDecoder(_DecoderBase):
    """Drop-in for simpleAICV.detection.decode.FCOSDecoder (decode.py:274-364)."""

    _is_fcos = True

    def __init__(self,
                 strides=[8, 16, 32, 64, 128],
                 max_object_num=100,
                 min_score_threshold=0.05,
                 topn=1000,
                 nms_type='python_nms',
                 nms_threshold=0.6):
        self._init_common(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)
        self.strides = strides

    def _geometry(self, shapes, batch, num_classes):
        return _geom.make_geometry(shapes, batch, 1, num_classes, self.strides)


class _FlatDecoder(_DecoderBase):
    """Shared tail of the decoders whose candidates are a flat list of (score, class, box) rows per
    image: b200det_select_decode_nms with B200DET_DECODE_BOXES on a one-level 1 x N geometry."""

    def _init_flat(self, max_object_num, min_score_threshold, topn, nms_type, nms_threshold):
        if nms_type is not None:
            assert nms_type in ['torch_nms', 'python_nms', 'diou_python_nms'], 'wrong nms type!'
        if topn > _lib.MAX_TOPN:
            raise ValueError(f'topn <= {_lib.MAX_TOPN} is supported')
        self.max_object_num = max_object_num
        self.min_score_threshold = min_score_threshold
        self.topn = topn
        self.nms_type = nms_type
        self.nms_threshold = nms_threshold
        self._pinned = None

    @staticmethod
    def _flat_geometry(batch, rows):
        geo = _lib.Geometry()
        geo.n_levels, geo.batch, geo.per_loc, geo.num_classes = 1, int(batch), 1, 1
        geo.height[0], geo.width[0], geo.stride[0] = 1, int(rows), 1.0
        return geo

    def _select(self, keys, classes, boxes, batch, rows, topn, max_out, min_score, details=False):
        """keys uint32 / classes int32 [batch*rows], boxes float32 [batch, rows, 4] on the device."""
        lib = _lib.load()
        device = boxes.device
        geo = self._flat_geometry(batch, rows)
        ws_bytes = int(lib.b200det_decode_workspace_bytes(ctypes.byref(geo), int(topn)))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=device)
        out = self._out_buffer(6 * batch * max_out, device)
        order = keep = None
        if details:
            order = torch.empty(batch * topn, dtype=torch.int32, device=device)
            keep = torch.empty(batch * topn, dtype=torch.int32, device=device)
        with _on_device(device):
            rc = (
            lib.b200det_select_decode_nms(
                ctypes.byref(geo), keys.data_ptr(), classes.data_ptr(), _lib.ptr_array([boxes]),
                _lib.F32, _lib.DECODE_BOXES, float(np.float32(min_score)), int(topn), int(max_out),
                _lib.NMS_CODES[self.nms_type],
                float(self.nms_threshold if self.nms_threshold is not None else 0.5), None, None, 0,
                out.data_ptr(), order.data_ptr() if details else None,
                keep.data_ptr() if details else None, None, ws.data_ptr(), ws_bytes,
                _lib.raw_stream(device)))
        _lib.check(rc, 'b200det_select_decode_nms')
        result = self._to_host(out, batch, max_out, device)
        if not details:
            return result
        return result, {'order': order.view(batch, topn).cpu().numpy(),
                        'keep': keep.view(batch, topn).cpu().numpy()}

    def _query_scores(self, cls, mode, boxes, sizes, num_classes, min_score):
        """One warp per (image, query) row: keys, classes and xyxy boxes (b200det_query_scores)."""
        lib = _lib.load()
        batch, queries, channels = (int(v) for v in cls.shape)
        device = cls.device
        keys = torch.empty(batch * queries, dtype=torch.int32, device=device)
        classes = torch.empty(batch * queries, dtype=torch.int32, device=device)
        xyxy = torch.empty((batch, queries, 4), dtype=torch.float32, device=device) \
            if boxes is not None else None
        dtype = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}[cls.dtype]
        with _on_device(device):
            rc = lib.b200det_query_scores(
                cls.data_ptr(), dtype, mode, boxes.data_ptr() if boxes is not None else None,
                sizes.data_ptr() if sizes is not None else None, batch, queries, channels,
                int(num_classes), float(np.float32(min_score)), keys.data_ptr(), classes.data_ptr(),
                xyxy.data_ptr() if xyxy is not None else None,
                _lib.raw_stream(device))
        _lib.check(rc, 'b200det_query_scores')
        return keys, classes, xyxy


def _device_f32(x, what, device=None):
    """CUDA float32 contiguous tensor from a CUDA tensor or a host array (the reference's DecodeMethod
    / DetNMSMethod take NumPy arrays: those are uploaded).  There is no CPU compute path."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError(f'{what} must be a CUDA tensor (or a NumPy array to upload)')
        return x.detach().to(torch.float32).contiguous()
    if not torch.cuda.is_available():
        raise RuntimeError('b200det needs a CUDA device (B200); there is no CPU fallback')
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(device or 'cuda')


class DetNMSMethod(_FlatDecoder):
    """Drop-in for simpleAICV.detection.decode.DetNMSMethod (decode.py:24-104): keep indices of
    greedy NMS over boxes that are already sorted by score."""

    def __init__(self, nms_type='python_nms', nms_threshold=0.5):
        assert nms_type in ['torch_nms', 'python_nms', 'diou_python_nms'], 'wrong nms type!'
        self._init_flat(1, -np.inf, 1, nms_type, nms_threshold)

    def __call__(self, sorted_bboxes, sorted_scores):
        n = int(sorted_scores.shape[0])
        if n == 0:
            return np.array([], dtype=np.int64 if self.nms_type == 'torch_nms' else np.float64)
        if n > _lib.MAX_TOPN:
            raise ValueError(f'at most {_lib.MAX_TOPN} boxes are supported')
        boxes = _device_f32(sorted_bboxes, 'sorted_bboxes').view(1, n, 4)
        device = boxes.device
        if self.nms_type == 'torch_nms':
            # torchvision orders by the scores it is given
            scores = _device_f32(sorted_scores, 'sorted_scores', device).view(1, n, 1)
        else:
            # python_nms never looks at the scores: it walks the boxes in the order given
            scores = torch.arange(n, 0, -1, dtype=torch.float32, device=device).view(1, n, 1)
        keys, classes, _ = self._query_scores(scores, _lib.SCORES_PROBS, None, None, 1, -np.inf)
        _, info = self._select(keys, classes, boxes, 1, n, n, 1, -np.inf, details=True)
        keep = info['keep'][0]
        keep = info['order'][0][keep[keep >= 0]]
        return keep.astype(np.int64 if self.nms_type == 'torch_nms' else np.int32)


class DecodeMethod(_FlatDecoder):
    """Drop-in for simpleAICV.detection.decode.DecodeMethod (decode.py:107-172): score threshold,
    descending sort, top-n, NMS and the max_object_num cap over pre-decoded boxes."""

    def __init__(self, max_object_num=100, min_score_threshold=0.05, topn=1000,
                 nms_type='python_nms', nms_threshold=0.5):
        assert nms_type in ['torch_nms', 'python_nms', 'diou_python_nms'], 'wrong nms type!'
        self._init_flat(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)

    def __call__(self, cls_scores, cls_classes, pred_bboxes):
        scores = _device_f32(cls_scores, 'cls_scores')
        device = scores.device
        batch, rows = (int(v) for v in scores.shape)
        boxes = _device_f32(pred_bboxes, 'pred_bboxes', device).view(batch, rows, 4)
        if isinstance(cls_classes, torch.Tensor):
            classes = cls_classes.detach().to(device=device, dtype=torch.int32).contiguous()
        else:
            classes = torch.as_tensor(np.ascontiguousarray(cls_classes).astype(np.int32)).to(device)
        keys, _, _ = self._query_scores(scores.view(batch, rows, 1), _lib.SCORES_PROBS, None, None,
                                        1, self.min_score_threshold)
        return self._select(keys, classes.view(-1), boxes, batch, rows, self.topn,
                            int(self.max_object_num), self.min_score_threshold)


class _QueryDecoder(_FlatDecoder):
    _mode = _lib.SCORES_SIGMOID

    def _decode(self, cls_preds, reg_preds, scaled_sizes, num_classes, details=False):
        if not (cls_preds.is_cuda and reg_preds.is_cuda):
            raise RuntimeError('b200det decoders need CUDA tensors; there is no CPU fallback')
        cls = cls_preds.detach().contiguous()
        if self._mode == _lib.SCORES_SOFTMAX and cls.dtype != torch.float32:
            raise RuntimeError('DETRDecoder: float32 class logits are supported')
        reg = reg_preds.detach().to(torch.float32).contiguous()
        batch, queries, channels = (int(v) for v in cls.shape)
        sizes = torch.as_tensor(np.asarray(scaled_sizes, dtype=np.float32).reshape(-1)).to(cls.device)
        if sizes.numel() != 2 * batch:
            raise ValueError('scaled_sizes must hold (height, width) per image')
        keys, classes, boxes = self._query_scores(
            cls, self._mode, reg, sizes, channels if num_classes is None else num_classes,
            self.min_score_threshold)
        return self._select(keys, classes, boxes, batch, queries, self.topn,
                            int(self.max_object_num), self.min_score_threshold, details=details)


class DETRDecoder(_QueryDecoder):
    """Drop-in for simpleAICV.detection.decode.DETRDecoder (decode.py:367-482): softmax over the
    num_classes + 1 channels of the last decoder layer, no-object rows dropped, optional NMS."""

    _mode = _lib.SCORES_SOFTMAX

    def __init__(self, num_classes=80, max_object_num=100, min_score_threshold=0.05, topn=100,
                 nms_type=None, nms_threshold=0.5):
        self.num_classes = num_classes
        self._init_flat(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)

    def __call__(self, preds, scaled_sizes):
        return self._decode(preds[0][-1, :, :, :], preds[1][-1, :, :, :], scaled_sizes,
                            self.num_classes)

    def decode_with_details(self, preds, scaled_sizes):
        return self._decode(preds[0][-1, :, :, :], preds[1][-1, :, :, :], scaled_sizes,
                            self.num_classes, details=True)


class DINODETRDecoder(_QueryDecoder):
    """Drop-in for simpleAICV.detection.decode.DINODETRDecoder (decode.py:485-594): sigmoid class
    scores of preds['pred_logits'], boxes from preds['pred_boxes']."""

    _mode = _lib.SCORES_SIGMOID

    def __init__(self, max_object_num=100, min_score_threshold=0.05, topn=300,
                 nms_type='python_nms', nms_threshold=0.5):
        self._init_flat(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)

    def __call__(self, preds, scaled_sizes):
        return self._decode(preds['pred_logits'], preds['pred_boxes'], scaled_sizes, None)

    def decode_with_details(self, preds, scaled_sizes):
        return self._decode(preds['pred_logits'], preds['pred_boxes'], scaled_sizes, None,
                            details=True)
