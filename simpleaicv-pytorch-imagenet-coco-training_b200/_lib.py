"""ctypes binding of libb200det.so (C ABI declared in include/b200det.h).

There is NO CPU fallback: if the library is missing, or a tensor is not on a CUDA device, the
callers raise.  Build the library with `python -m b200det._build` / `__graft_entry__.build()`.
"""
import ctypes
import os

from . import _build

MAX_LEVELS = 8
MAX_PER_LOC = 16
MAX_GT = 2048
MAX_TOPN = 2048

F32, F16, BF16 = 0, 1, 2
REG_EXP_ROUNDED = 0x10   # see include/b200det.h: eager half arithmetic rounds exp() to half
BOX_NONE, BOX_SMOOTHL1, BOX_IOU, BOX_GIOU, BOX_DIOU, BOX_CIOU, BOX_EIOU = range(7)
NMS_PYTHON, NMS_DIOU_PYTHON, NMS_TORCH, NMS_NONE = range(4)
DECODE_ANCHORS, DECODE_POINTS, DECODE_BOXES = range(3)
SCORES_PROBS, SCORES_SIGMOID, SCORES_SOFTMAX = range(3)

BOX_LOSS_CODES = {
    'SmoothL1': BOX_SMOOTHL1,
    'IoU': BOX_IOU,
    'GIoU': BOX_GIOU,
    'DIoU': BOX_DIOU,
    'CIoU': BOX_CIOU,
    'EIoU': BOX_EIOU,
}
NMS_CODES = {'python_nms': NMS_PYTHON, 'diou_python_nms': NMS_DIOU_PYTHON, 'torch_nms': NMS_TORCH,
             None: NMS_NONE}


class Geometry(ctypes.Structure):
    """Mirror of `b200det_geometry`."""
    _fields_ = [
        ('n_levels', ctypes.c_int32),
        ('batch', ctypes.c_int32),
        ('per_loc', ctypes.c_int32),
        ('num_classes', ctypes.c_int32),
        ('height', ctypes.c_int32 * MAX_LEVELS),
        ('width', ctypes.c_int32 * MAX_LEVELS),
        ('stride', ctypes.c_float * MAX_LEVELS),
        ('base_anchors', ((ctypes.c_float * 4) * MAX_PER_LOC) * MAX_LEVELS),
        ('mi_lo', ctypes.c_float * MAX_LEVELS),
        ('mi_hi', ctypes.c_float * MAX_LEVELS),
        ('radius', ctypes.c_float * MAX_LEVELS),
    ]


class LossParams(ctypes.Structure):
    """Mirror of `b200det_loss_params`."""
    _fields_ = [
        ('is_fcos', ctypes.c_int32),
        ('box_loss', ctypes.c_int32),
        ('reg_dtype', ctypes.c_int32),
        ('use_center_sample', ctypes.c_int32),
        ('alpha', ctypes.c_float),
        ('gamma', ctypes.c_float),
        ('beta', ctypes.c_float),
        ('w_cls', ctypes.c_float),
        ('w_box', ctypes.c_float),
        ('w_ctr', ctypes.c_float),
        ('iou_neg', ctypes.c_float),
        ('iou_pos', ctypes.c_float),
    ]


MAX_PEERS = 16


class PeerExchange(ctypes.Structure):
    """Mirror of `b200det_peer_exchange`."""
    _fields_ = [
        ('rank', ctypes.c_int32),
        ('world', ctypes.c_int32),
        ('epoch', ctypes.c_uint64),
        ('timeout_cycles', ctypes.c_uint64),
        ('peer', ctypes.c_void_p * MAX_PEERS),
    ]


class DecodeParams(ctypes.Structure):
    """Mirror of `b200det_decode_params`."""
    _fields_ = [
        ('is_fcos', ctypes.c_int32),
        ('reg_dtype', ctypes.c_int32),
        ('topn', ctypes.c_int32),
        ('max_out', ctypes.c_int32),
        ('nms_type', ctypes.c_int32),
        ('min_score', ctypes.c_float),
        ('nms_threshold', ctypes.c_double),
        ('scales', ctypes.c_void_p),
        ('sizes', ctypes.c_void_p),
        ('to_xywh', ctypes.c_int32),
        ('half_exp_table', ctypes.c_void_p),
    ]


_vp = ctypes.c_void_p
_vpp = ctypes.POINTER(ctypes.c_void_p)
_geo = ctypes.POINTER(Geometry)

# name -> (restype, argtypes); every symbol include/b200det.h declares
SIGNATURES = {
    'b200det_abi_version': (ctypes.c_int, []),
    'b200det_error_string': (ctypes.c_char_p, [ctypes.c_int]),
    'b200det_launch_count': (ctypes.c_ulonglong, []),
    'b200det_profile': (ctypes.c_int, [ctypes.c_int]),
    'b200det_profile_read': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_int)]),
    'b200det_kernel_name': (ctypes.c_char_p, [ctypes.c_int]),
    'b200det_loss_forward': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vp,
        ctypes.c_size_t, _vp, _vp, _vp
    ]),
    'b200det_loss_forward_grad': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vpp, _vpp,
        _vpp, _vp, ctypes.c_size_t, _vp, _vp, _vp
    ]),
    'b200det_loss_forward_grad_overlap': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vpp, _vpp,
        _vpp, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp
    ]),
    'b200det_scale_levels': (ctypes.c_int, [
        _vpp, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, _vp, _vp, ctypes.c_float, _vp
    ]),
    'b200det_eval_step': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), ctypes.POINTER(DecodeParams), _vp, ctypes.c_int, _vpp,
        _vpp, _vpp, _vp, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp
    ]),
    'b200det_eval_step_overlap': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), ctypes.POINTER(DecodeParams), _vp, ctypes.c_int, _vpp,
        _vpp, _vpp, _vp, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp
    ]),
    'b200det_decode': (ctypes.c_int, [
        _geo, ctypes.POINTER(DecodeParams), _vpp, _vpp, _vpp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_size_t, _vp
    ]),
    'b200det_decode_from_keys': (ctypes.c_int, [
        _geo, ctypes.POINTER(DecodeParams), _vpp, _vpp, _vpp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_int, _vp
    ]),
    'b200det_stream_synchronize': (ctypes.c_int, [_vp]),
    'b200det_rows_per_image': (ctypes.c_longlong, [_geo]),
    'b200det_loss_workspace_bytes': (ctypes.c_size_t, [_geo]),
    'b200det_retina_assign': (ctypes.c_int, [
        _geo, _vp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, ctypes.c_size_t, _vp
    ]),
    'b200det_fcos_assign': (ctypes.c_int, [
        _geo, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp
    ]),
    'b200det_sparse_losses': (ctypes.c_int, [
        _geo, ctypes.c_int, _vp, ctypes.c_int, _vp, _vpp, ctypes.c_int, _vpp, ctypes.c_int,
        ctypes.c_float, _vpp, ctypes.c_float, ctypes.c_float, _vpp, _vpp, _vp, ctypes.c_size_t, _vp
    ]),
    'b200det_focal_loss': (ctypes.c_int, [
        _geo, _vpp, _vp, ctypes.c_float, ctypes.c_float, _vpp, _vp, ctypes.c_float, _vp,
        ctypes.c_size_t, _vp
    ]),
    'b200det_loss_reduce': (ctypes.c_int, [_geo, ctypes.c_int, _vp, ctypes.c_size_t, _vp, _vp]),
    'b200det_loss_reduce_finish': (ctypes.c_int, [
        _geo, _vp, ctypes.c_size_t, ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp
    ]),
    'b200det_loss_finish': (ctypes.c_int,
                            [_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp, _vp]),
    'b200det_scale_f32': (ctypes.c_int, [_vp, ctypes.c_longlong, _vp, _vp]),
    'b200det_scale_pos_rows': (ctypes.c_int, [
        _geo, _vp, ctypes.c_size_t, _vpp, _vpp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp
    ]),
    'b200det_decode_workspace_bytes': (ctypes.c_size_t, [_geo, ctypes.c_int]),
    'b200det_score_argmax': (ctypes.c_int, [_geo, _vpp, _vpp, ctypes.c_float, _vp, _vp, _vp]),
    'b200det_select_decode_nms': (ctypes.c_int, [
        _geo, _vp, _vp, _vpp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
        ctypes.c_int, ctypes.c_int, ctypes.c_double, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp,
        _vp, ctypes.c_size_t, _vp
    ]),
    'b200det_query_scores': (ctypes.c_int, [
        _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, ctypes.c_float, _vp, _vp, _vp, _vp
    ]),
    'b200det_rows_to_image_major': (ctypes.c_int, [_geo, _vp, _vp, ctypes.c_int, _vp]),
    'b200det_generate_rows': (ctypes.c_int, [_geo, ctypes.c_int, _vp, _vp]),
    'b200det_npexp_f32': (ctypes.c_int, [_vp, _vp, ctypes.c_longlong, _vp]),
    'b200det_select_stamps': (ctypes.c_int, [_vp]),
    'b200det_logits_sweep': (ctypes.c_int, [
        _geo, _vpp, ctypes.c_int, _vpp, _vp, ctypes.c_float, ctypes.c_float, _vp, ctypes.c_size_t,
        ctypes.c_float, _vp, _vp, _vp
    ]),
    'b200det_logits_eval_step': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), ctypes.POINTER(DecodeParams), _vp, ctypes.c_int, _vpp,
        ctypes.c_int, _vpp, _vpp, _vp, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_size_t, _vp
    ]),
    'b200det_iou_method': (ctypes.c_int, [
        _vp, ctypes.c_longlong, ctypes.c_longlong, _vp, ctypes.c_longlong, ctypes.c_longlong,
        ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp
    ]),
    'b200det_pair_ious': (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp]),
    'b200det_voc_match': (ctypes.c_int, [
        _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp
    ]),
    'b200det_peer_buffer_create': (ctypes.c_int, [_vpp, ctypes.c_char_p]),
    'b200det_peer_buffer_open': (ctypes.c_int, [ctypes.c_char_p, _vpp]),
    'b200det_peer_buffer_close': (ctypes.c_int, [_vp]),
    'b200det_peer_buffer_destroy': (ctypes.c_int, [_vp]),
    'b200det_sums_exchange': (ctypes.c_int, [
        ctypes.POINTER(PeerExchange), ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp,
        _vp
    ]),
    'b200det_loss_reduce_exchange': (ctypes.c_int, [
        _geo, _vp, ctypes.c_size_t, ctypes.POINTER(PeerExchange), ctypes.c_float, ctypes.c_float,
        ctypes.c_float, _vp, _vp, _vp, _vp
    ]),
    'b200det_loss_forward_exchange': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vp,
        ctypes.c_size_t, ctypes.POINTER(PeerExchange), _vp, _vp, _vp, _vp
    ]),
    'b200det_loss_forward_overlap': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vp,
        ctypes.c_size_t, ctypes.POINTER(PeerExchange), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_int
    ]),
    'b200det_loss_forward_keys': (ctypes.c_int, [
        _geo, ctypes.POINTER(LossParams), _vp, ctypes.c_int, _vpp, _vpp, _vpp, _vp, _vp,
        ctypes.c_size_t, ctypes.POINTER(PeerExchange), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
        ctypes.c_int, ctypes.c_float, _vp, _vp
    ]),
    'b200det_stream_create': (ctypes.c_int, [_vpp, ctypes.c_int]),
    'b200det_stream_destroy': (ctypes.c_int, [_vp]),
    'b200det_event_create': (ctypes.c_int, [_vpp]),
    'b200det_event_destroy': (ctypes.c_int, [_vp]),
    'b200det_head_sigmoid_permute': (ctypes.c_int, [
        _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, _vp, _vp]),
    'b200det_head_sigmoid_permute_backward': (ctypes.c_int, [
        _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, _vp, ctypes.c_int, _vp]),
}

_LIB = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Returns the loaded library; raises RuntimeError (loudly) if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f'{path} is missing: the b200det CUDA library has not been built. Run '
            '`python -m b200det._build` (needs nvcc; targets sm_100a). There is no CPU fallback.')
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.b200det_abi_version() != 1:
        raise RuntimeError('libb200det.so ABI version mismatch; rebuild it')
    _LIB = lib
    return lib


_FAST = None


def fastpath():
    """The optional host-side marshalling fast path (csrc/fastpath.cpp, built by
    _build.build_fastpath) bound to the loaded library's entry points, or None (not built, or
    B200DET_FASTPATH=0).  It is not a compute path: it calls the same C-ABI functions."""
    global _FAST
    if _FAST is None:
        _FAST = False
        if os.environ.get('B200DET_FASTPATH', '1') != '0':
            try:
                from . import _fastpath as mod
                lib = load()
                mod.bind(ctypes.cast(lib.b200det_loss_forward_overlap, ctypes.c_void_p).value,
                         ctypes.cast(lib.b200det_decode, ctypes.c_void_p).value,
                         ctypes.cast(lib.b200det_loss_forward_keys, ctypes.c_void_p).value,
                         ctypes.cast(lib.b200det_decode_from_keys, ctypes.c_void_p).value)
                _FAST = mod
            except Exception:   # noqa: BLE001 -- optional
                _FAST = False
    return _FAST or None


def check(rc, what):
    if rc != 0:
        msg = load().b200det_error_string(rc).decode()
        raise RuntimeError(f'{what} failed: {msg} (code {rc})')


def ptr_array(tensors):
    """ctypes void*[n] of device pointers (None -> NULL array)."""
    if tensors is None:
        return None
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


try:   # raw cudaStream_t of torch's current stream without building a torch.cuda.Stream object
    import torch as _torch
    _RAW_STREAM = _torch._C._cuda_getCurrentRawStream
except (ImportError, AttributeError):   # pragma: no cover
    _RAW_STREAM = None


def raw_stream(device=None):
    """ctypes void* of torch's current stream on `device` (torch.device, index or None = current).
    torch.cuda.current_stream() costs ~7-15 us per call; this is ~0.5 us."""
    import torch
    if device is None:
        index = torch.cuda.current_device()
    elif isinstance(device, int):
        index = device
    else:
        index = device.index if device.index is not None else torch.cuda.current_device()
    if _RAW_STREAM is not None:
        return ctypes.c_void_p(_RAW_STREAM(index))
    return ctypes.c_void_p(torch.cuda.current_stream(index).cuda_stream)


def launch_count():
    return int(load().b200det_launch_count())


def profile_start():
    """Per-kernel CUDA-event timing inside the library (b200det_profile)."""
    load().b200det_profile(1)


def profile_pause():
    load().b200det_profile(0)


def profile_resume():
    load().b200det_profile(2)


def profile_stop():
    """Stops the timing and returns {kernel name: (launches, mean ms)}; synchronises the events."""
    lib = load()
    lib.b200det_profile(0)
    out = {}
    for kid in range(10):
        ms = ctypes.c_double(0.0)
        n = ctypes.c_int(0)
        check(lib.b200det_profile_read(kid, ctypes.byref(ms), ctypes.byref(n)), 'profile_read')
        if n.value:
            out[lib.b200det_kernel_name(kid).decode()] = (n.value, ms.value / n.value)
    return out
