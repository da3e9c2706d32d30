"""The sweep hand-over inside the reference's two-call evaluation step (b200det/_handoff.py,
tools/scripts.py:733-740): criterion(preds, annots) followed by decoder(preds) on the SAME tensors
reads cls once -- and must give exactly what the two independent calls give (which the other GPU
tests pin to the oracle and to the reference's golden vectors), whatever the caller does between
the two calls."""
import numpy as np
import pytest
import torch

from b200det import _handoff, decode, losses, synth
from oracle import det_oracle as O

from test_gpu_parity import LOSS_RTOL, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fresh():
    _handoff.reset()
    was = _handoff.ENABLED
    _handoff.ENABLED = True
    yield
    _handoff.ENABLED = was
    _handoff.reset()


def _separate(crit, dec, preds, ann):
    """the two calls with the hand-over off: every call sweeps for itself"""
    _handoff.ENABLED = False
    try:
        with torch.no_grad():
            loss = {k: float(v) for k, v in crit(preds, ann).items()}
            det = dec(preds)
    finally:
        _handoff.ENABLED = True
    return loss, det


def _same_detections(got, want):
    for g, w, what in zip(got, want, ('scores', 'classes', 'boxes')):
        assert np.array_equal(g, w), what


def _same_loss(got, want):
    for k in want:
        g = float(got[k])
        assert abs(g - want[k]) <= LOSS_RTOL * max(abs(want[k]), 1e-6), (k, g, want[k])


def _retina(batch=3, size=256, seed=5, num_classes=80):
    preds = synth.make_tie_free(synth.make_retina_preds(batch, size, num_classes, seed=seed, sigma=1.5))
    ann = synth.make_annotations(batch, 20, size, num_classes, seed=seed + 1)
    return dev(preds), ann.cuda()


def _fcos(batch=3, size=256, seed=6, num_classes=80):
    preds = synth.make_tie_free(synth.make_fcos_preds(batch, size, num_classes, seed=seed, sigma=1.5))
    ann = synth.make_annotations(batch, 20, size, num_classes, seed=seed + 1)
    return dev(preds), ann.cuda()


@pytest.mark.parametrize('family', ['retina', 'fcos', 'retina21', 'fcos365', 'fcos365_gamma'])
@pytest.mark.parametrize('fast', [True, False])
def test_two_calls_share_one_sweep(family, fast, monkeypatch):
    """class counts that are a multiple of 4 take the TMA-fed row-group sweep, the others (21, and
    Objects365's 365 = BASELINE configs[3]) the raw-tile fused sweep; gamma != 2 the exact-form terms"""
    if not fast:
        from b200det import _lib
        monkeypatch.setattr(_lib, '_FAST', False)
    if family == 'retina':
        preds, ann = _retina()
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    elif family == 'retina21':
        preds, ann = _retina(num_classes=21)
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    elif family.startswith('fcos365'):
        preds, ann = _fcos(batch=2, size=320, num_classes=365)
        crit = losses.FCOSLoss(strides=synth.STRIDES, gamma=1.5 if family.endswith('gamma') else 2.)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    else:
        preds, ann = _fcos()
        crit = losses.FCOSLoss(strides=synth.STRIDES)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    want_loss, want_det = _separate(crit, dec, preds, ann)
    before = dict(_handoff.stats)
    for it in range(4):
        with torch.no_grad():
            loss = crit(preds, ann)
            det = dec(preds)
        _same_loss(loss, want_loss)
        _same_detections(det, want_det)
    # the first decoder call asks, iterations 2..4 are served by the criterion's sweep
    assert _handoff.stats['produced'] - before['produced'] == 3
    assert _handoff.stats['consumed'] - before['consumed'] == 3
    assert _handoff.stats['stale'] == before['stale']


def test_handover_matches_the_oracle():
    preds_h = synth.make_tie_free(synth.make_retina_preds(2, 192, 80, seed=11, sigma=1.5))
    ann_h = synth.make_annotations(2, 12, 192, 80, seed=12)
    preds, ann = dev(preds_h), ann_h.cuda()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    for _ in range(2):
        with torch.no_grad():
            loss = crit(preds, ann)
            det = dec(preds)
    assert _handoff.stats['consumed'] >= 1
    want = O.retina_loss(preds_h, ann_h, **synth.RETINA_KW, box_loss_type='GIoU')
    _same_loss(loss, {k: float(want[k]) for k in ('cls_loss', 'reg_loss')})
    (ws, wc, wb), _ = O.retina_decode(preds_h, **synth.RETINA_KW)
    _same_detections(det, (ws, wc, wb))


def test_handover_matches_the_oracle_odd_class_count():
    """C = 37 (not a multiple of 4): the raw-tile fused sweep, FCOS with centre-ness, against the oracle"""
    preds_h = synth.make_tie_free(synth.make_fcos_preds(2, 192, 37, seed=21, sigma=1.5))
    ann_h = synth.make_annotations(2, 12, 192, 37, seed=22)
    preds, ann = dev(preds_h), ann_h.cuda()
    crit = losses.FCOSLoss(strides=synth.STRIDES)
    dec = decode.FCOSDecoder(strides=synth.STRIDES)
    for _ in range(2):
        with torch.no_grad():
            loss = crit(preds, ann)
            det = dec(preds)
    assert _handoff.stats['consumed'] >= 1
    want = O.fcos_loss(preds_h, ann_h, synth.STRIDES, synth.MI)
    _same_loss(loss, {k: float(want[k]) for k in ('cls_loss', 'reg_loss', 'center_ness_loss')})
    (ws, wc, wb), _ = O.fcos_decode(preds_h, synth.STRIDES)
    _same_detections(det, (ws, wc, wb))


@pytest.mark.parametrize('shape', ['configs3', 'configs4_shard'])
def test_handover_equals_separate_calls_at_baseline_shapes(shape):
    """Size-independent property at the BASELINE sizes (the oracle takes ~1 s per image there): the
    two calls sharing one sweep give the detections of two independent calls bit for bit and the same
    losses.  configs[3] = FCOS, Objects365 shape (365 classes: the ring-fed raw-tile sweep), 1024^2,
    batch 32, <= 200 GT; configs[4] shard = RetinaNet 800^2, 80 classes, 32 images (one of 8 GPUs)."""
    dev_ = torch.device('cuda')
    if shape == 'configs3':
        preds = synth.make_fcos_preds(32, 1024, 365, seed=31, device=dev_)
        ann = synth.make_annotations(32, 200, 1024, 365, seed=32).to(dev_)
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    else:
        preds = synth.make_retina_preds(32, 800, 80, seed=33, device=dev_)
        ann = synth.make_annotations(32, 100, 800, 80, seed=34).to(dev_)
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    want_loss, want_det = _separate(crit, dec, preds, ann)
    before = dict(_handoff.stats)
    for _ in range(3):
        with torch.no_grad():
            loss = crit(preds, ann)
            det = dec(preds)
        _same_loss(loss, want_loss)
        _same_detections(det, want_det)
    assert _handoff.stats['consumed'] - before['consumed'] == 2
    assert _handoff.stats['stale'] == before['stale']
    assert (want_det[0] > 0).sum() == 32 * 100   # every image fills its 100 detections at these shapes


def test_retinaface_pair_hands_over_and_matches_the_reference_golden():
    """SURVEY 8f-2 through the hand-over: RetinaFaceLoss then RetinaFaceDecoder on the same head outputs
    (one class: the raw-tile fused sweep with one lane per row) against the vectors of the reference's own
    face_detection classes."""
    import golden_util as G
    from b200det.face_detection import losses as face_losses, decode as face_decode
    face = G.load('retinaface_small.npz')
    preds_h, ann_h = G.retina_inputs(face)
    preds, ann = dev(preds_h), ann_h.cuda()
    sizes, strides = [[8, 16, 32], [32, 64, 128], [128, 256, 512]], [8, 16, 32]
    crit = face_losses.RetinaFaceLoss(anchor_sizes=sizes, strides=strides, box_loss_type='CIoU')
    dec = face_decode.RetinaFaceDecoder(anchor_sizes=sizes, strides=strides, nms_type='python_nms')
    before = dict(_handoff.stats)
    for _ in range(3):
        with torch.no_grad():
            loss = crit(preds, ann)
            s, c, b = dec(preds)
        want = face['loss_CIoU']
        for got, w in zip((float(loss['cls_loss']), float(loss['reg_loss'])), want):
            assert abs(got - float(w)) <= LOSS_RTOL * max(abs(float(w)), 1e-6)
        G.assert_bit_equal(s, face['dec_python_nms_scores'], 'scores')
        G.assert_bit_equal(c, face['dec_python_nms_classes'], 'classes')
        G.assert_bit_equal(b, face['dec_python_nms_boxes'], 'boxes')
    assert _handoff.stats['consumed'] - before['consumed'] == 2


def test_torch_write_between_the_calls_is_seen():
    """an in-place torch op bumps the version counter: the decoder sweeps for itself"""
    preds, ann = _retina()
    other, _ = _retina(seed=9)
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    with torch.no_grad():
        crit(preds, ann), dec(preds)            # arms the hand-over
        crit(preds, ann)                        # keys of `preds` are waiting
        for t, o in zip(preds[0], other[0]):
            t.copy_(o)
        for t, o in zip(preds[1], other[1]):
            t.copy_(o)
        dropped = _handoff.stats['dropped']
        det = dec(preds)
    assert _handoff.stats['dropped'] == dropped + 1
    _, want = _separate(crit, dec, other, ann)
    _same_detections(det, want)


def test_write_behind_torchs_back_is_caught_on_the_device():
    """`.data.copy_` (like a CUDA-graph replay into static buffers) leaves the version counter
    alone: the select kernel finds rows whose keys no longer match and the decoder starts over"""
    preds, ann = _retina()
    other, _ = _retina(seed=9)
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    with torch.no_grad():
        crit(preds, ann), dec(preds)
        crit(preds, ann)
        for t, o in zip(preds[0], other[0]):
            t.data.copy_(o)
        for t, o in zip(preds[1], other[1]):
            t.data.copy_(o)
        stale = _handoff.stats['stale']
        det = dec(preds)
    assert _handoff.stats['stale'] == stale + 1
    _, want = _separate(crit, dec, other, ann)
    _same_detections(det, want)


def test_other_tensor_objects_are_not_served():
    preds, ann = _retina()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    _, want = _separate(crit, dec, preds, ann)
    with torch.no_grad():
        crit(preds, ann), dec(preds)
        crit(preds, ann)
        consumed = _handoff.stats['consumed']
        clones = [[t.clone() for t in grp] for grp in preds]
        det = dec(clones)
        assert _handoff.stats['consumed'] == consumed
        _same_detections(det, want)
        # and a second decoder call after ONE criterion call sweeps itself again
        crit(preds, ann)
        _same_detections(dec(preds), want)
        assert _handoff.stats['consumed'] == consumed + 1
        _same_detections(dec(preds), want)
        assert _handoff.stats['consumed'] == consumed + 1


def test_criterion_only_loops_stop_producing_keys():
    preds, ann = _retina()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    want, _ = _separate(crit, dec, preds, ann)
    with torch.no_grad():
        dec(preds)                              # leaves a wish
        produced = _handoff.stats['produced']
        for _ in range(6):
            _same_loss(crit(preds, ann), want)
    assert _handoff.stats['produced'] - produced == 2


def test_threshold_and_family_must_match():
    preds, ann = _retina()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    dec2 = decode.RetinaDecoder(**synth.RETINA_KW, min_score_threshold=0.2)
    _, want2 = _separate(crit, dec2, preds, ann)
    with torch.no_grad():
        crit(preds, ann), dec(preds)
        crit(preds, ann)                        # keys thresholded for `dec`
        consumed = _handoff.stats['consumed']
        _same_detections(dec2(preds), want2)    # another decoder: its own sweep
        assert _handoff.stats['consumed'] == consumed


def test_training_and_graph_capture_never_hand_over():
    preds, ann = _retina(batch=2, size=128)
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    with torch.no_grad():
        dec(preds)
    produced = _handoff.stats['produced']
    req = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    sum(crit(req, ann).values()).backward()
    assert _handoff.stats['produced'] == produced
