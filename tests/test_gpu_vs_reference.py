"""GPU parity against the UNMODIFIED reference itself, run on the same box: baseline/_ref holds the
reference's three files of the path (vendored by baseline/fetch_ref.sh; it travels with the gpurun
snapshot), its losses run on torch-CUDA tensors, its decoders in NumPy on the host -- exactly what
tools/scripts.py does.  Covers what CPU-made golden vectors cannot:

  * the BASELINE.json batch sizes (configs[1..3] at B = 16 / 16 / 32), labels and losses;
  * the reference's real AMP flow: criterion called inside `torch.cuda.amp.autocast()` with a
    float16 regression head (tools/scripts.py:886-893);
  * decoders at the BASELINE shapes.

Skipped when baseline/_ref is absent (a checkout that never ran __graft_entry__.build() where
/root/reference exists)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from b200det import synth, losses, decode  # noqa: E402
from baseline import refarm  # noqa: E402

import golden_util as G  # noqa: E402
from test_gpu_parity import LOSS_RTOL, assert_close, dev, loss_values  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refarm.available(), reason='baseline/_ref not vendored')]


@pytest.fixture(scope='module')
def ref():
    try:
        return refarm.load()
    except RuntimeError as exc:   # e.g. /root/reference's package already imported in this process
        pytest.skip(str(exc))


def ref_retina_labels(crit, preds, ann):
    """labels [B,A] of the reference's own assignment (losses.py:322-388), on the tensors' device."""
    sizes = [[c.shape[2], c.shape[1]] for c in preds[0]]
    anchors = torch.cat([torch.tensor(a).view(-1, 4) for a in crit.anchors(sizes)], dim=0)
    anchors = anchors.to(ann.device).unsqueeze(0).repeat(ann.shape[0], 1, 1)
    return crit.get_batch_anchors_annotations(anchors, ann)[:, :, 4].to(torch.int32)


@pytest.mark.parametrize('box_type', ['GIoU', 'SmoothL1'])
def test_config1_retina_loss_batch16_vs_reference_on_cuda(ref, box_type):
    """BASELINE configs[1] at its real batch: RetinaLoss, 800x800, 80 classes, B = 16, <= 100 GT."""
    L, _ = ref
    preds = dev(synth.make_retina_preds(16, 800, 80, seed=61))
    ann = synth.make_annotations(16, 100, 800, 80, seed=62, empty_images=(5,)).cuda()
    theirs = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type).cuda()
    ours = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
    with torch.no_grad():
        want = theirs(preds, ann)
        got = ours(preds, ann)
    keys = ['cls_loss', 'reg_loss']
    assert_close(loss_values(got, keys), loss_values(want, keys), LOSS_RTOL, f'B=16 {box_type}')
    if box_type == 'GIoU':
        labels = ours.debug_assign(preds, ann)['labels']
        assert torch.equal(labels, ref_retina_labels(theirs, preds, ann)), 'labels, B = 16'


def test_config2_fcos_batch16_vs_reference_on_cuda(ref):
    """BASELINE configs[2] at its real batch: FCOSLoss + FCOSDecoder, 800x800, 80 classes, B = 16."""
    L, D = ref
    cpu = synth.make_tie_free(synth.make_fcos_preds(16, 800, 80, seed=63))
    preds = dev(cpu)
    ann = synth.make_annotations(16, 100, 800, 80, seed=64, empty_images=(0,)).cuda()
    theirs = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI).cuda()
    ours = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with torch.no_grad():
        want = theirs(preds, ann)
        got = ours(preds, ann)
    keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    assert_close(loss_values(got, keys), loss_values(want, keys), LOSS_RTOL, 'FCOS B=16')
    s0, c0, b0 = D.FCOSDecoder(strides=synth.STRIDES)(preds)
    s, c, b = decode.FCOSDecoder(strides=synth.STRIDES)(preds)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(c, c0, 'classes')
    G.assert_bit_equal(b, b0, 'boxes')


def test_config3_fcos_objects365_batch32_vs_reference_on_cuda(ref):
    """BASELINE configs[3] at its real batch: 365 classes, 1024x1024, B = 32, <= 200 GT."""
    L, D = ref
    gen_cpu = synth.make_fcos_preds(32, 1024, 365, seed=65)
    # (the first 4 images also go through the decoders: make their scores tie-free, in place)
    synth.make_tie_free([[t[:4] for t in grp] for grp in gen_cpu])
    preds = dev(gen_cpu)
    ann = synth.make_annotations(32, 200, 1024, 365, seed=66).cuda()
    theirs = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI).cuda()
    ours = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with torch.no_grad():
        want = theirs(preds, ann)
        got = ours(preds, ann)
    keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    assert_close(loss_values(got, keys), loss_values(want, keys), LOSS_RTOL, 'FCOS O365 B=32')
    # decoders on the first 4 images (the reference decoder is NumPy on the host)
    sub = [[t[:4].contiguous() for t in grp] for grp in preds]
    s0, c0, b0 = D.FCOSDecoder(strides=synth.STRIDES)(sub)
    s, c, b = decode.FCOSDecoder(strides=synth.STRIDES)(sub)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(c, c0, 'classes')
    G.assert_bit_equal(b, b0, 'boxes')


def test_config0_retina_decoder_vs_reference(ref):
    """BASELINE configs[0]: RetinaDecoder + NMS on one 800x800 image, against the reference's NumPy."""
    _, D = ref
    preds = dev(synth.make_tie_free(synth.make_retina_preds(2, 800, 80, seed=67)))
    for nms in ('python_nms', 'diou_python_nms', 'torch_nms'):
        s0, c0, b0 = D.RetinaDecoder(**synth.RETINA_KW, nms_type=nms)(preds)
        s, c, b = decode.RetinaDecoder(**synth.RETINA_KW, nms_type=nms)(preds)
        G.assert_bit_equal(s, s0, f'{nms} scores')
        G.assert_bit_equal(c, c0, f'{nms} classes')
        G.assert_bit_equal(b, b0, f'{nms} boxes')


@pytest.mark.parametrize('box_type', ['GIoU', 'SmoothL1', 'CIoU'])
def test_amp_flow_matches_the_reference_under_cuda_autocast(ref, box_type):
    """The reference's AMP training step (tools/scripts.py:886-893): the criterion runs INSIDE
    `with autocast()` on a float16 regression head.  CUDA autocast executes torch.exp in float32,
    so the result equals exp of the upcast value; loss values within 1e-5, gradients within half
    precision (they are float16 tensors)."""
    L, _ = ref
    base = synth.make_retina_preds(4, 256, 8, seed=68)
    ann = synth.make_annotations(4, 16, 256, 8, seed=69).cuda()

    def heads():
        cls = [t.cuda().requires_grad_(True) for t in base[0]]
        reg = [t.cuda().half().requires_grad_(True) for t in base[1]]
        return [cls, reg]

    theirs = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type).cuda()
    ours = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
    pa, pb = heads(), heads()
    with torch.autocast('cuda', dtype=torch.float16):
        want = theirs(pa, ann)
        got = ours(pb, ann)
    keys = ['cls_loss', 'reg_loss']
    assert_close([got[k].float().item() for k in keys], [want[k].float().item() for k in keys],
                 2e-6 if box_type == 'CIoU' else LOSS_RTOL, f'autocast {box_type}')
    sum(want.values()).backward()
    sum(got.values()).backward()
    for i in range(len(pa[0])):
        gw, gg = pa[1][i].grad, pb[1][i].grad
        assert gg.dtype == torch.float16 and gw.dtype == torch.float16
        gw, gg = gw.float().cpu().numpy().astype(np.float64), gg.float().cpu().numpy().astype(np.float64)
        tol = np.maximum(np.abs(gw), 1e-7) * 2.0 ** -9 + 1e-8
        assert (np.abs(gg - gw) <= tol).all(), f'reg grad level {i}'
        cw, cg = pa[0][i].grad.cpu().numpy(), pb[0][i].grad.cpu().numpy()
        np.testing.assert_allclose(cg, cw, rtol=1e-4, atol=1e-7)


def test_eager_half_flow_matches_the_reference_on_cuda(ref):
    """Without autocast a float16 regression head is exponentiated IN half precision by the
    reference (torch.exp on the half tensor, losses.py:417-426)."""
    L, D = ref
    base = synth.make_tie_free(synth.make_retina_preds(3, 256, 8, seed=70))
    ann = synth.make_annotations(3, 16, 256, 8, seed=71).cuda()
    preds = [[t.cuda() for t in base[0]], [t.cuda().half() for t in base[1]]]
    for box_type in ('GIoU', 'EIoU'):
        theirs = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type).cuda()
        ours = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
        with torch.no_grad():
            want = theirs(preds, ann)
            got = ours(preds, ann)
        keys = ['cls_loss', 'reg_loss']
        assert_close([got[k].float().item() for k in keys], [want[k].float().item() for k in keys],
                     LOSS_RTOL, f'eager half {box_type}')
    s0, c0, b0 = D.RetinaDecoder(**synth.RETINA_KW)(preds)
    s, c, b = decode.RetinaDecoder(**synth.RETINA_KW)(preds)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(b, b0, 'boxes')
