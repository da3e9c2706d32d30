"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/b200det.h declares, validates arguments without touching the device, and the Python
classes keep the reference's constructor / call surface."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import b200det
from b200det import _lib, geometry, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'b200det.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200det_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/b200det.h but not exported'
    assert sorted(_lib.SIGNATURES) == names, 'ctypes binding table and header disagree'
    assert lib.b200det_abi_version() == 1
    assert lib.b200det_error_string(0) == b'ok'
    assert b'workspace' in lib.b200det_error_string(-3)


def test_geometry_struct_matches_header_layout():
    # 4 ints + 8+8 ints + 8 floats + 8*16*4 floats + 3*8 floats
    assert ctypes.sizeof(_lib.Geometry) == 4 * (4 + 16 + 8 + 8 * 16 * 4 + 24)
    shapes = [(p, p) for p in synth.pyramid_sizes(800)]
    base = geometry.retina_base_anchors(synth.AREAS, synth.RATIOS, synth.SCALES)
    geo = geometry.make_geometry(shapes, 16, 9, 80, synth.STRIDES, base_anchors=base)
    lib = _lib.load()
    assert lib.b200det_rows_per_image(ctypes.byref(geo)) == 120087
    assert lib.b200det_loss_workspace_bytes(ctypes.byref(geo)) > 16 * 120087 * 12
    geo = geometry.make_geometry([(p, p) for p in synth.pyramid_sizes(1024)], 32, 1, 365,
                                 synth.STRIDES, mi=synth.MI, center_sample_radius=1.5)
    assert lib.b200det_rows_per_image(ctypes.byref(geo)) == 21824
    assert geo.radius[0] == 12.0 and geo.mi_hi[4] == 1e8


def test_argument_errors_are_returned_not_raised():
    lib = _lib.load()
    geo = geometry.make_geometry([(4, 4)], 1, 1, 8, [8], mi=[[-1, 64]], center_sample_radius=1.5)
    # null pointers / bad sizes are rejected on the host before any CUDA call
    assert lib.b200det_retina_assign(ctypes.byref(geo), None, 4, 0.4, 0.5, None, None, None, 0,
                                     None) == -1
    assert lib.b200det_focal_loss(ctypes.byref(geo), None, None, 0.25, 2.0, None, None, 1.0,
                                  None, 0, None) == -1
    assert lib.b200det_loss_finish(None, 1.0, 1.0, 1.0, None, None) == -1
    bad = _lib.Geometry()
    bad.n_levels = 99
    assert lib.b200det_rows_per_image(ctypes.byref(bad)) == -2
    with pytest.raises(RuntimeError, match='workspace too small'):
        _lib.check(-3, 'x')


def test_base_anchor_table_matches_golden():
    import golden_util as G
    t = G.load('tables.npz')
    G.assert_bit_equal(geometry.retina_base_anchors(synth.AREAS, synth.RATIOS, synth.SCALES),
                       t['base_anchors'], 'host base-anchor table vs reference')


def test_drop_in_surface_matches_reference_signatures():
    """Same names in the module __dict__ (configs do losses.__dict__['RetinaLoss'](**kw)), same
    keyword names and defaults as simpleAICV/detection/losses.py:128-139, :434-445 and
    decode.py:177-187, :276-282."""
    from b200det import losses, decode
    expect = {
        (losses, 'RetinaLoss'): ['areas', 'ratios', 'scales', 'strides', 'alpha', 'gamma', 'beta',
                                 'cls_loss_weight', 'box_loss_weight', 'box_loss_type'],
        (losses, 'FCOSLoss'): ['strides', 'mi', 'alpha', 'gamma', 'cls_loss_weight',
                               'box_loss_weight', 'center_ness_loss_weight', 'box_loss_iou_type',
                               'center_sample_radius', 'use_center_sample'],
        (decode, 'RetinaDecoder'): ['areas', 'ratios', 'scales', 'strides', 'max_object_num',
                                    'min_score_threshold', 'topn', 'nms_type', 'nms_threshold'],
        (decode, 'FCOSDecoder'): ['strides', 'max_object_num', 'min_score_threshold', 'topn',
                                  'nms_type', 'nms_threshold'],
    }
    for (mod, name), params in expect.items():
        cls = mod.__dict__[name]
        sig = inspect.signature(cls.__init__)
        positional = [p.name for p in sig.parameters.values()
                      if p.kind == p.POSITIONAL_OR_KEYWORD and p.name != 'self']
        assert positional == params, name
    d = inspect.signature(decode.FCOSDecoder.__init__).parameters
    assert d['nms_threshold'].default == 0.6 and d['topn'].default == 1000
    r = inspect.signature(losses.RetinaLoss.__init__).parameters
    assert r['beta'].default == 1.0 / 9.0 and r['box_loss_type'].default == 'SmoothL1'
    with pytest.raises(AssertionError):
        losses.RetinaLoss(box_loss_type='L2')
    with pytest.raises(AssertionError):
        decode.RetinaDecoder(nms_type='soft_nms')
    from b200det.face_detection import losses as fl, decode as fd
    fsig = inspect.signature(fl.__dict__['RetinaFaceLoss'].__init__).parameters
    assert [p for p in fsig if p not in ('self', 'sync_normalizer', 'process_group')] == [
        'anchor_sizes', 'strides', 'alpha', 'gamma', 'beta', 'cls_loss_weight', 'box_loss_weight',
        'box_loss_type']
    assert fsig['box_loss_type'].default == 'CIoU'
    dsig = inspect.signature(fd.__dict__['RetinaFaceDecoder'].__init__).parameters
    assert dsig['min_score_threshold'].default == 0.3 and dsig['nms_threshold'].default == 0.3
    isig = inspect.signature(losses.__dict__['IoUMethod'].__call__).parameters    # losses.py:33
    assert list(isig) == ['self', 'boxes1', 'boxes2', 'iou_type', 'box_type']
    assert isig['iou_type'].default == 'IoU' and isig['box_type'].default == 'xyxy'
    import torch
    assert isinstance(losses.RetinaLoss(), torch.nn.Module)      # train script calls .cuda() on it


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'simpleaicv-pytorch-imagenet-coco-training_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('oracle/npexp.c', ''), f'{f} mentions the oracle'


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_LIB', None)
    monkeypatch.setattr(_lib._build, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.load()


def test_synthetic_generators_are_deterministic_and_tie_free():
    a = synth.make_retina_preds(1, 128, 8, seed=5)
    b = synth.make_retina_preds(1, 128, 8, seed=5)
    assert all((x == y).all() for x, y in zip(a[0] + a[1], b[0] + b[1]))
    synth.make_tie_free(a)
    cls = np.concatenate([c[0].reshape(-1, 8).numpy() for c in a[0]])
    s = cls.max(axis=1)
    s = s[s > np.float32(0.05)]
    assert len(np.unique(s)) == len(s)
    ann = synth.make_annotations(4, 10, 128, 8, seed=2, empty_images=(1,))
    assert (ann[1] == -1).all() and (ann[0][:, 4] >= 0).any()
    assert synth.pyramid_sizes(800) == [100, 50, 25, 13, 7]
    assert synth.pyramid_sizes(1024) == [128, 64, 32, 16, 8]


def test_peer_exchange_needs_a_process_group():
    """b200det.peer (NVLink peer-memory exchange of the loss normaliser) refuses to start without
    torch.distributed; the struct mirror has the header's layout."""
    import ctypes
    from b200det import _lib, peer
    with pytest.raises(RuntimeError):
        peer.PeerExchange()
    assert ctypes.sizeof(_lib.PeerExchange) == 4 + 4 + 8 + 8 + 8 * _lib.MAX_PEERS


def test_new_modules_have_no_cpu_path():
    """anchor / evaluation / IoUMethod fail loudly without a CUDA device or with CPU tensors (the
    product never falls back to the oracle or to torch-CPU arithmetic)."""
    import numpy as np
    import torch
    from b200det import anchor, evaluation, losses
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(RuntimeError):
        anchor.RetinaAnchors()([[4, 4], [2, 2], [1, 1], [1, 1], [1, 1]])
    with pytest.raises(RuntimeError):
        anchor.FCOSPositions()([[4, 4]])
    boxes = np.array([[0, 0, 10, 10]], dtype=np.float32)
    with pytest.raises(RuntimeError):
        evaluation.compute_ious(boxes, boxes)
    with pytest.raises(RuntimeError):
        evaluation.voc_map([[boxes, np.zeros(1, np.float32), np.ones(1, np.float32)]],
                           [[boxes, np.zeros(1, np.float32)]], [0.5], 1)
    with pytest.raises(RuntimeError):
        losses.IoUMethod()(torch.from_numpy(boxes), torch.from_numpy(boxes))
