"""Live pin: the oracle against the UNMODIFIED reference imported from /root/reference, on fresh
seeded inputs (other seeds / sizes than the committed golden vectors, including one full-size
800x800 image of BASELINE configs[0]).  Skipped where the reference checkout is absent (GPU box)."""
import numpy as np
import pytest
import torch

import refload
from b200det import synth
from oracle import det_oracle as O

import golden_util as G

pytestmark = pytest.mark.skipif(not refload.available(), reason='needs /root/reference')


@pytest.fixture(scope='module')
def ref():
    return refload.load()


@pytest.mark.parametrize('box_type', ['SmoothL1', 'GIoU', 'DIoU', 'CIoU', 'EIoU', 'IoU'])
def test_retina_loss_live(ref, box_type):
    L, _, _ = ref
    preds = synth.make_retina_preds(2, 192, 12, seed=21)
    ann = synth.make_annotations(2, 20, 192, 12, seed=22)
    with torch.no_grad():
        want = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)(preds, ann)
        got = O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type=box_type)
    for k in ('cls_loss', 'reg_loss'):
        G.assert_bit_equal(np.float32(got[k].item()), np.float32(want[k].item()), k)


def test_fcos_loss_live(ref):
    L, _, _ = ref
    preds = synth.make_fcos_preds(2, 320, 12, seed=23)
    ann = synth.make_annotations(2, 20, 320, 12, seed=24)
    for kw in (dict(), dict(box_loss_iou_type='CIoU', use_center_sample=False),
               dict(center_sample_radius=2.5, alpha=0.3, gamma=2.0)):
        with torch.no_grad():
            want = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, **kw)(preds, ann)
            got = O.fcos_loss(preds, ann, synth.STRIDES, synth.MI, **kw)
        for k in ('cls_loss', 'reg_loss', 'center_ness_loss'):
            G.assert_bit_equal(np.float32(got[k].item()), np.float32(want[k].item()), k)


def test_retina_decoder_full_size_live(ref):
    """BASELINE configs[0]: one 800x800 image, 80 classes, 120 087 anchors."""
    _, D, _ = ref
    preds = synth.make_tie_free(synth.make_retina_preds(1, 800, 80, seed=0))
    want = D.RetinaDecoder(**synth.RETINA_KW)(preds)
    got, extra = O.retina_decode(preds, **synth.RETINA_KW)
    for a, b, name in zip(got, want, ('scores', 'classes', 'boxes')):
        G.assert_bit_equal(a, b, name)
    assert len(extra['per_image'][0]['order']) == 1000


def test_fcos_decoder_live(ref):
    _, D, _ = ref
    preds = synth.make_tie_free(synth.make_fcos_preds(2, 512, 20, seed=25))
    for nms in ('python_nms', 'diou_python_nms'):
        want = D.FCOSDecoder(strides=synth.STRIDES, nms_type=nms)(preds)
        got, _ = O.fcos_decode(preds, synth.STRIDES, nms_type=nms)
        for a, b, name in zip(got, want, ('scores', 'classes', 'boxes')):
            G.assert_bit_equal(a, b, name)
