"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in classes) against
  (a) the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  (b) the oracle on seeded inputs at the BASELINE.json shapes.
Bars: labels / matched indices / targets / top-n order / keep lists / decoded boxes / scores
bit-exact; loss values within 1e-5 relative; gradients within 1e-4 relative (+1e-7 abs)."""
import numpy as np
import pytest
import torch

from b200det import synth, losses, decode
from oracle import det_oracle as O

import golden_util as G

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5   # north_star: "loss values must agree within 1e-5 relative"
GRAD_RTOL = 1e-4
GRAD_ATOL = 1e-7


def dev(preds):
    return synth.to_device(preds, 'cuda')


def loss_values(d, keys):
    return np.array([d[k].item() for k in keys], dtype=np.float64)


def assert_close(got, want, rtol, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
    assert (err <= rtol).all(), f'{what}: got {got}, want {want}, rel err {err}'


def assert_targets_equal(got, want, what):
    """FCOS targets: l,t,r,b,label bit-exact.  Centre-ness is sqrt(...) (losses.py:822-824): the
    kernel and torch-CUDA use the correctly rounded IEEE sqrt, but torch-CPU's float32 sqrt goes
    through MKL VML and is off by one ulp for ~0.6 % of inputs (measured), so against CPU-made
    truth the centre-ness may differ by exactly one ulp; test_fcos_targets_vs_oracle_on_cuda
    checks it bit-exactly against the same oracle run on torch-CUDA."""
    got = np.asarray(got)
    want = np.asarray(want)
    G.assert_bit_equal(got[..., 0:5], want[..., 0:5], what + ' l,t,r,b,label')
    ulp = np.abs(G.bits(got[..., 5]).astype(np.int64) - G.bits(want[..., 5]).astype(np.int64))
    assert ulp.max() <= 1, f'{what}: centre-ness differs by {ulp.max()} ulp'
    assert (ulp != 0).mean() < 0.02


def check_decode_details(info, extras, topn):
    for b, e in enumerate(extras):
        n = len(e['order'])
        assert info['counts'][b, 1] == n
        assert np.array_equal(info['order'][b, :n], e['order']), f'top-n order, image {b}'
        assert (info['order'][b, n:] == -1).all()
        k = len(e['keep'])
        assert info['counts'][b, 2] == k
        assert np.array_equal(info['keep'][b, :k], e['keep']), f'NMS keep list, image {b}'
        assert (info['keep'][b, k:] == -1).all()


# ------------------------------------------------------------------------------------------
# golden vectors (reference outputs)
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def retina():
    return G.load('retina_small.npz')


@pytest.fixture(scope='module')
def fcos():
    return G.load('fcos_small.npz')


@pytest.mark.parametrize('box_type', ['SmoothL1'] + G.IOU_TYPES)
def test_retina_loss_golden(retina, box_type):
    preds, ann = G.retina_inputs(retina)
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type).cuda()
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
    assert set(d.keys()) == {'cls_loss', 'reg_loss'}
    assert_close(loss_values(d, ['cls_loss', 'reg_loss']), retina[f'loss_{box_type}'], LOSS_RTOL,
                 f'RetinaLoss {box_type}')


def test_retina_assignment_golden(retina):
    preds, ann = G.retina_inputs(retina)
    crit = losses.RetinaLoss(**synth.RETINA_KW)
    got = crit.debug_assign(dev(preds), ann.cuda())
    want = retina['assign_GIoU']
    assert np.array_equal(got['labels'].cpu().numpy(), want[..., 4].astype(np.int32))
    fast = crit.debug_assign(dev(preds), ann.cuda(), exact=False)
    assert np.array_equal(fast['labels'].cpu().numpy(), want[..., 4].astype(np.int32))
    _, _, matched = O.retina_assign(torch.from_numpy(retina['anchors'].copy()), ann, 'GIoU')
    assert np.array_equal(got['matched'].cpu().numpy(), matched.numpy().astype(np.int32))


@pytest.mark.parametrize('box_type', ['SmoothL1', 'GIoU', 'CIoU'])
def test_retina_gradients_golden(retina, box_type):
    preds, ann = G.retina_inputs(retina)
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
    d = crit(p, ann.cuda())
    (d['cls_loss'] + 2.0 * d['reg_loss']).backward()
    for i in range(len(p[0])):
        np.testing.assert_allclose(p[0][i].grad.cpu().numpy(), retina[f'gcls_{box_type}_{i}'],
                                   rtol=GRAD_RTOL, atol=GRAD_ATOL)
        np.testing.assert_allclose(p[1][i].grad.cpu().numpy(), retina[f'greg_{box_type}_{i}'],
                                   rtol=GRAD_RTOL, atol=GRAD_ATOL)


@pytest.mark.parametrize('nms', ['python_nms', 'diou_python_nms', 'torch_nms'])
@pytest.mark.parametrize('tag,kw', [('default', {}), ('small', dict(topn=300, max_object_num=20))])
def test_retina_decoder_golden(retina, nms, tag, kw):
    preds, _ = G.retina_inputs(retina)
    dec = decode.RetinaDecoder(**synth.RETINA_KW, nms_type=nms, **kw)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    assert s.dtype == np.float32 and c.dtype == np.float32 and b.dtype == np.float32
    assert s.flags.writeable and b.flags.writeable
    G.assert_bit_equal(s, retina[f'dec_{nms}_{tag}_scores'], 'scores')
    G.assert_bit_equal(c, retina[f'dec_{nms}_{tag}_classes'], 'classes')
    G.assert_bit_equal(b, retina[f'dec_{nms}_{tag}_boxes'], 'boxes')
    _, extra = O.retina_decode(preds, **synth.RETINA_KW, nms_type=nms, **kw)
    check_decode_details(info, extra['per_image'], dec.topn)
    s2, c2, b2 = dec(dev(preds))   # the plain __call__ path (NMS stops at max_object_num)
    G.assert_bit_equal(s2, s)
    G.assert_bit_equal(c2, c)
    G.assert_bit_equal(b2, b)


@pytest.mark.parametrize('iou_type', G.IOU_TYPES)
def test_fcos_loss_golden(fcos, iou_type):
    preds, ann = G.fcos_inputs(fcos)
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou_type)
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
    assert set(d.keys()) == {'cls_loss', 'reg_loss', 'center_ness_loss'}
    assert_close(loss_values(d, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 fcos[f'loss_{iou_type}'], LOSS_RTOL, f'FCOSLoss {iou_type}')


@pytest.mark.parametrize('tag,center', [('center', True), ('nocenter', False)])
def test_fcos_assignment_golden(fcos, tag, center):
    preds, ann = G.fcos_inputs(fcos)
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, use_center_sample=center)
    got = crit.debug_assign(dev(preds), ann.cuda())
    want = fcos[f'targets_{tag}']
    assert_targets_equal(got['targets'].cpu().numpy(), want[..., 0:6], 'golden targets')
    assert np.array_equal(got['labels'].cpu().numpy(), want[..., 4].astype(np.int32))
    with torch.no_grad():
        ref = O.fcos_loss(preds, ann, synth.STRIDES, synth.MI, use_center_sample=center)
        d = crit(dev(preds), ann.cuda())
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    assert_close(loss_values(d, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 fcos[f'loss_{tag}'], LOSS_RTOL, 'FCOSLoss')


@pytest.mark.parametrize('iou_type', ['GIoU', 'EIoU'])
def test_fcos_gradients_golden(fcos, iou_type):
    preds, ann = G.fcos_inputs(fcos)
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou_type)
    d = crit(p, ann.cuda())
    (d['cls_loss'] + 2.0 * d['reg_loss'] + 3.0 * d['center_ness_loss']).backward()
    for i in range(len(p[0])):
        np.testing.assert_allclose(p[0][i].grad.cpu().numpy(), fcos[f'gcls_{iou_type}_{i}'],
                                   rtol=GRAD_RTOL, atol=GRAD_ATOL)
        np.testing.assert_allclose(p[1][i].grad.cpu().numpy(), fcos[f'greg_{iou_type}_{i}'],
                                   rtol=GRAD_RTOL, atol=GRAD_ATOL)
        np.testing.assert_allclose(p[2][i].grad.cpu().numpy(), fcos[f'gctr_{iou_type}_{i}'],
                                   rtol=GRAD_RTOL, atol=GRAD_ATOL)


@pytest.mark.parametrize('nms', ['python_nms', 'diou_python_nms', 'torch_nms'])
@pytest.mark.parametrize('tag,kw', [('default', {}), ('small', dict(topn=300, max_object_num=20))])
def test_fcos_decoder_golden(fcos, nms, tag, kw):
    preds, _ = G.fcos_inputs(fcos)
    dec = decode.FCOSDecoder(strides=synth.STRIDES, nms_type=nms, **kw)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    G.assert_bit_equal(s, fcos[f'dec_{nms}_{tag}_scores'], 'scores')
    G.assert_bit_equal(c, fcos[f'dec_{nms}_{tag}_classes'], 'classes')
    G.assert_bit_equal(b, fcos[f'dec_{nms}_{tag}_boxes'], 'boxes')
    _, extra = O.fcos_decode(preds, synth.STRIDES, nms_type=nms, **kw)
    check_decode_details(info, extra['per_image'], dec.topn)


def test_fcos_targets_vs_oracle_on_cuda(fcos):
    """The oracle's torch ops run on CUDA tensors (how the reference runs in training): every
    primitive is then a correctly rounded IEEE op and the whole target tensor is bit-exact."""
    preds, ann = G.fcos_inputs(fcos)
    for center in (True, False):
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, use_center_sample=center)
        got = crit.debug_assign(dev(preds), ann.cuda())
        with torch.no_grad():
            ref = O.fcos_loss(dev(preds), ann.cuda(), synth.STRIDES, synth.MI,
                              use_center_sample=center)
        G.assert_bit_equal(got['targets'].cpu().numpy(), ref['targets'].cpu().numpy(), 'targets')
        assert np.array_equal(got['matched'].cpu().numpy(),
                              ref['matched'].cpu().numpy().astype(np.int32))
    preds, ann = G.retina_inputs(G.load('retina_small.npz'))
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    got = crit.debug_assign(dev(preds), ann.cuda())
    with torch.no_grad():
        ref = O.retina_loss(dev(preds), ann.cuda(), **synth.RETINA_KW, box_loss_type='GIoU')
    assert np.array_equal(got['labels'].cpu().numpy(), ref['labels'].cpu().numpy().astype(np.int32))
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].cpu().numpy().astype(np.int32))


# ------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------
def test_npexp_kernel():
    import ctypes
    from b200det import _lib
    lib = _lib.load()
    t = G.load('tables.npz')
    rng = np.random.RandomState(11)
    x = np.concatenate([t['exp_x'], rng.uniform(-110, 90, 1 << 20).astype(np.float32),
                        rng.normal(0, 1, 1 << 20).astype(np.float32)])
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    _lib.check(lib.b200det_npexp_f32(xd.data_ptr(), yd.data_ptr(), xd.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               'npexp')
    G.assert_bit_equal(yd.cpu().numpy(), O.np_exp_f32(x), 'device npexp vs oracle')
    G.assert_bit_equal(yd.cpu().numpy()[:t['exp_y'].size], t['exp_y'], 'device npexp vs np.exp')


@pytest.mark.parametrize('size', [800, 1024])
def test_generated_rows(size):
    import ctypes
    from b200det import _lib, geometry
    lib = _lib.load()
    p = synth.pyramid_sizes(size)
    shapes = [(q, q) for q in p]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    base = geometry.retina_base_anchors(synth.AREAS, synth.RATIOS, synth.SCALES)
    geo = geometry.make_geometry(shapes, 1, 9, 80, synth.STRIDES, base_anchors=base)
    n = geometry.rows_per_image(shapes, 9)
    out = torch.empty(n * 4, dtype=torch.float32, device='cuda')
    _lib.check(lib.b200det_generate_rows(ctypes.byref(geo), 0, out.data_ptr(), st), 'rows')
    want = np.concatenate([a.reshape(-1, 4) for a in
                           O.retina_anchors([[q, q] for q in p], **synth.RETINA_KW)])
    G.assert_bit_equal(out.cpu().numpy().reshape(-1, 4), want, 'anchors')
    geo = geometry.make_geometry(shapes, 1, 1, 80, synth.STRIDES)
    n = geometry.rows_per_image(shapes, 1)
    out = torch.empty(n * 2, dtype=torch.float32, device='cuda')
    _lib.check(lib.b200det_generate_rows(ctypes.byref(geo), 1, out.data_ptr(), st), 'rows')
    want = np.concatenate([a.reshape(-1, 2) for a in
                           O.fcos_positions([[q, q] for q in p], synth.STRIDES)])
    G.assert_bit_equal(out.cpu().numpy().reshape(-1, 2), want, 'positions')


# ------------------------------------------------------------------------------------------
# BASELINE.json shapes vs the oracle (sizes the oracle finishes in seconds)
# ------------------------------------------------------------------------------------------
def test_config1_retina_decode_800():
    """configs[0]: RetinaDecoder + NMS, COCO 80 cls, 800x800, 9 anchors, batch 1."""
    preds = synth.make_tie_free(synth.make_retina_preds(1, 800, 80, seed=0))
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(c, c0, 'classes')
    G.assert_bit_equal(b, b0, 'boxes')
    check_decode_details(info, extra['per_image'], 1000)
    assert info['counts'][0, 0] == int((extra['scores'][0] > np.float32(0.05)).sum())


@pytest.mark.parametrize('box_type', ['SmoothL1', 'GIoU'])
def test_config2_retina_loss_800(box_type):
    """configs[1] at batch 4 (the oracle needs ~1 s per image): assignment + focal + box loss."""
    B, G_ = 4, 100
    preds = synth.make_retina_preds(B, 800, 80, seed=0)
    ann = synth.make_annotations(B, G_, 800, 80, seed=1, empty_images=(2,))
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
        ref = O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type=box_type)
    got = crit.debug_assign(dev(preds), ann.cuda())
    assert np.array_equal(got['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    # the production scan (what forward() runs: early rejection below IoU 0.38) gives the same labels
    fast = crit.debug_assign(dev(preds), ann.cuda(), exact=False)
    assert np.array_equal(fast['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
    assert int(crit.last_stats['sums'][0].item()) == ref['num_pos'] > 0
    assert_close(loss_values(d, ['cls_loss', 'reg_loss']),
                 [ref['cls_loss'].item(), ref['reg_loss'].item()], LOSS_RTOL, 'RetinaLoss')


def test_config3_fcos_800():
    """configs[2] at batch 4: FCOSLoss centre-sampling assignment + FCOSDecoder NMS."""
    B, G_ = 4, 100
    preds = synth.make_tie_free(synth.make_fcos_preds(B, 800, 80, seed=0))
    ann = synth.make_annotations(B, G_, 800, 80, seed=1, empty_images=(1,))
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
        ref = O.fcos_loss(preds, ann, synth.STRIDES, synth.MI)
    got = crit.debug_assign(dev(preds), ann.cuda())
    assert_targets_equal(got['targets'].cpu().numpy(), ref['targets'].numpy(), 'targets')
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    assert_close(loss_values(d, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 [ref['cls_loss'].item(), ref['reg_loss'].item(), ref['center_ness_loss'].item()],
                 LOSS_RTOL, 'FCOSLoss')
    dec = decode.FCOSDecoder(strides=synth.STRIDES)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.fcos_decode(preds, synth.STRIDES)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(c, c0, 'classes')
    G.assert_bit_equal(b, b0, 'boxes')
    check_decode_details(info, extra['per_image'], 1000)


def test_config4_fcos_objects365_1024():
    """configs[3] at batch 2: 365 classes (rows not 16-byte aligned -> scalar path), 1024x1024,
    up to 200 GT per image."""
    B, G_ = 2, 200
    preds = synth.make_tie_free(synth.make_fcos_preds(B, 1024, 365, seed=4))
    ann = synth.make_annotations(B, G_, 1024, 365, seed=5, min_gt=150)
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
        ref = O.fcos_loss(preds, ann, synth.STRIDES, synth.MI)
    got = crit.debug_assign(dev(preds), ann.cuda())
    assert_targets_equal(got['targets'].cpu().numpy(), ref['targets'].numpy(), 'targets')
    assert_close(loss_values(d, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 [ref['cls_loss'].item(), ref['reg_loss'].item(), ref['center_ness_loss'].item()],
                 LOSS_RTOL, 'FCOSLoss')
    dec = decode.FCOSDecoder(strides=synth.STRIDES)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.fcos_decode(preds, synth.STRIDES)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(c, c0, 'classes')
    G.assert_bit_equal(b, b0, 'boxes')
    check_decode_details(info, extra['per_image'], 1000)


# ------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------
def test_no_annotations_anywhere():
    preds = synth.make_retina_preds(2, 128, 8, seed=3)
    ann = torch.full((2, 6, 5), -1.)
    d = losses.RetinaLoss(**synth.RETINA_KW)(dev(preds), ann.cuda())
    assert d['cls_loss'].item() == 0. and d['reg_loss'].item() == 0.   # losses.py:234-235
    got = losses.RetinaLoss(**synth.RETINA_KW).debug_assign(dev(preds), ann.cuda())
    assert (got['labels'] == -1).all() and (got['matched'] == -1).all()  # losses.py:341-345
    fp = synth.make_fcos_preds(2, 128, 8, seed=3)
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    d = crit(dev(fp), ann.cuda())
    assert all(v.item() == 0. for v in d.values())
    got = crit.debug_assign(dev(fp), ann.cuda())
    assert (got['labels'] == 0).all() and (got['targets'] == 0).all()     # losses.py:671-675
    empty = torch.zeros((2, 0, 5))
    d = crit(dev(fp), empty.cuda())
    assert all(v.item() == 0. for v in d.values())


def test_no_candidates_and_few_candidates():
    preds = synth.make_retina_preds(2, 128, 8, seed=3)
    for c in preds[0]:
        c.mul_(0.01)                       # every score far below 0.05
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    s, c, b = dec(dev(preds))
    assert (s == -1).all() and (c == -1).all() and (b == 0).all()
    preds[0][0][1, 3, 4, 2, 5] = 0.9       # exactly one candidate in image 1
    preds[0][2][1, 0, 1, 7, 0] = 0.05      # == threshold: strict '>' keeps it out
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    assert info['counts'][:, 0].tolist() == [0, 1]


def test_heavy_score_ties_are_ordered_by_row():
    """All-equal scores: the reference's argsort order is undefined; ours is ascending row index
    (the oracle's stable sort), and the radix select must cut the tie bucket exactly."""
    preds = synth.make_retina_preds(2, 128, 8, seed=3)
    for c in preds[0]:
        c.fill_(0.02)
        c[..., 3] = 0.3
    preds[0][1][0, 1, 1, 4, 6] = 0.7
    dec = decode.RetinaDecoder(**synth.RETINA_KW, topn=500, max_object_num=50)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW, topn=500, max_object_num=50)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    check_decode_details(info, extra['per_image'], 500)


def test_extreme_regression_values_truncate_like_x86():
    """exp overflow / huge boxes: NumPy's astype(int32) yields INT_MIN, CUDA would saturate."""
    preds = synth.make_tie_free(synth.make_retina_preds(1, 128, 8, seed=5))
    preds[1][0][0, :4, :, :, 2:] = 30.      # exp(30) * anchor >> 2^31
    preds[1][0][0, 4:8, :, :, 2:] = 95.     # exp -> inf
    preds[1][1][0, :, :, :, 0] = -1e9
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW)
    G.assert_bit_equal(b, b0, 'boxes')
    G.assert_bit_equal(s, s0, 'scores')
    check_decode_details(info, extra['per_image'], 1000)


def test_half_precision_regression_head():
    """A half-precision regression head reaches the path in two ways (both reproduced):
    inside `with autocast()` (tools/scripts.py:886-893): CUDA autocast runs torch.exp in float32 on
    the upcast value -> equals the oracle fed the upcast tensors; as plain half tensors: torch.exp /
    np.exp round to half (losses.py:417-426, decode.py:257-268) -> equals the oracle fed the half
    tensors (SmoothL1 has no exp: both coincide)."""
    preds = synth.make_tie_free(synth.make_retina_preds(2, 128, 8, seed=6))
    ann = synth.make_annotations(2, 12, 128, 8, seed=7)
    for dt in (torch.float16, torch.bfloat16):
        half = [preds[0], [r.to(dt) for r in preds[1]]]
        up = [preds[0], [r.float() for r in half[1]]]
        for box in ('SmoothL1', 'GIoU'):
            crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box)
            with torch.no_grad():
                eager = crit(dev(half), ann.cuda())
                with torch.autocast('cuda', dtype=dt):
                    amp = crit(dev(half), ann.cuda())
                ref_half = O.retina_loss(half, ann, **synth.RETINA_KW, box_loss_type=box)
                ref_up = O.retina_loss(up, ann, **synth.RETINA_KW, box_loss_type=box)
            keys = ['cls_loss', 'reg_loss']
            assert_close(loss_values(eager, keys), [ref_half[k].float().item() for k in keys],
                         LOSS_RTOL, f'eager {dt} {box}')
            assert_close(loss_values(amp, keys), [ref_up[k].item() for k in keys], LOSS_RTOL,
                         f'autocast {dt} {box}')
        if dt == torch.float16:   # the reference's decoders cannot take bfloat16 (`.numpy()` raises)
            s, c, b = decode.RetinaDecoder(**synth.RETINA_KW)(dev(half))
            (s0, c0, b0), _ = O.retina_decode(half, **synth.RETINA_KW)
            G.assert_bit_equal(b, b0, 'boxes')
            G.assert_bit_equal(s, s0, 'scores')


def host_numpy_half_exp_is_svml():
    """np.exp on float16 is CPU-dependent (AVX512-FP16 hosts take an SVML kernel); the golden
    vectors were made on such a host."""
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except ImportError:
        from numpy.core._multiarray_umath import __cpu_features__ as feats
    return bool(feats.get('AVX512_SPR'))


def half_grad_close(got, want, what):
    """Gradients w.r.t. a half tensor are half values.  The reference's autograd casts the upstream
    gradient of exp() to float16 BEFORE multiplying by exp(t) (the exp node lives in half), which
    loses bits -- all of them below 6e-8, most of them below the 6e-5 subnormal limit -- while the
    kernel keeps float32 up to the final cast: equal within 2 ulp of half precision plus an absolute
    2e-5 for the subnormal upstream values (measured worst case 7.6e-6)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    tol = np.abs(want) * 2.0 ** -9 + 2e-5
    bad = np.abs(got - want) > tol
    assert not bad.any(), f'{what}: {int(bad.sum())} of {bad.size} differ, max ' \
                          f'{np.abs(got - want).max()}'


@pytest.mark.parametrize('reg', ['f16', 'bf16'])
def test_half_reg_golden_retina(reg):
    """tests/golden/half_nan.npz: the unmodified reference on float16 / bfloat16 regression heads
    without autocast (exp rounded to half)."""
    d = G.load('half_nan.npz')
    preds, ann = G.half_nan_inputs(d, 'r', reg)
    for box in ('SmoothL1', 'GIoU', 'CIoU'):
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box)
        with torch.no_grad():
            got = crit(dev(preds), ann.cuda())
        assert_close(loss_values(got, ['cls_loss', 'reg_loss']), d[f'r_loss_{reg}_{box}'],
                     2e-6 if box == 'CIoU' else LOSS_RTOL, f'{reg} {box}')
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
    out = crit(p, ann.cuda())
    (out['cls_loss'] + 2.0 * out['reg_loss']).backward()
    for i in range(len(p[0])):
        assert p[1][i].grad.dtype == preds[1][i].dtype
        if reg == 'f16':
            half_grad_close(p[1][i].grad.float().cpu().numpy(), d[f'r_greg_{reg}_{i}'], f'greg {i}')
    if reg == 'f16':
        s, c, b = decode.RetinaDecoder(**synth.RETINA_KW)(dev(preds))
        (s0, c0, b0), _ = O.retina_decode(preds, **synth.RETINA_KW)    # np.exp of THIS host
        G.assert_bit_equal(s, s0, 'scores')
        G.assert_bit_equal(b, b0, 'boxes')
        if host_numpy_half_exp_is_svml():   # ... which is the golden's when the CPUs match
            G.assert_bit_equal(s, d['r_dec_f16_scores'], 'scores')
            G.assert_bit_equal(c, d['r_dec_f16_classes'], 'classes')
            G.assert_bit_equal(b, d['r_dec_f16_boxes'], 'boxes')


def test_half_reg_golden_fcos():
    d = G.load('half_nan.npz')
    preds, ann = G.half_nan_inputs(d, 'f', 'f16')
    for iou in ('GIoU', 'DIoU'):
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou)
        with torch.no_grad():
            got = crit(dev(preds), ann.cuda())
        assert_close(loss_values(got, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                     d[f'f_loss_f16_{iou}'], LOSS_RTOL, f'fcos f16 {iou}')
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
    out = crit(p, ann.cuda())
    (out['cls_loss'] + 2.0 * out['reg_loss'] + 3.0 * out['center_ness_loss']).backward()
    for i in range(len(p[0])):
        half_grad_close(p[1][i].grad.float().cpu().numpy(), d[f'f_greg_f16_{i}'], f'greg {i}')
    s, c, b = decode.FCOSDecoder(strides=synth.STRIDES)(dev(preds))
    (s0, c0, b0), _ = O.fcos_decode(preds, synth.STRIDES)
    G.assert_bit_equal(s, s0, 'scores')
    G.assert_bit_equal(b, b0, 'boxes')
    if host_numpy_half_exp_is_svml():
        G.assert_bit_equal(s, d['f_dec_f16_scores'], 'scores')
        G.assert_bit_equal(c, d['f_dec_f16_classes'], 'classes')
        G.assert_bit_equal(b, d['f_dec_f16_boxes'], 'boxes')


def test_decoder_nan_scores_golden():
    """np.argmax stops at the first NaN (decode.py:230-238): a row with a NaN class score (or NaN
    centre-ness) has a NaN score and never passes `score > threshold`, whatever its other classes."""
    d = G.load('half_nan.npz')
    preds, _ = G.half_nan_inputs(d, 'rn')
    for nms in ('python_nms', 'torch_nms'):
        s, c, b = decode.RetinaDecoder(**synth.RETINA_KW, nms_type=nms)(dev(preds))
        G.assert_bit_equal(s, d[f'rn_dec_{nms}_scores'], 'scores')
        G.assert_bit_equal(c, d[f'rn_dec_{nms}_classes'], 'classes')
        G.assert_bit_equal(b, d[f'rn_dec_{nms}_boxes'], 'boxes')
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW)
    (s, c, b), info = decode.RetinaDecoder(**synth.RETINA_KW).decode_with_details(dev(preds))
    check_decode_details(info, extra['per_image'], 1000)
    preds, _ = G.half_nan_inputs(d, 'fn')
    s, c, b = decode.FCOSDecoder(strides=synth.STRIDES)(dev(preds))
    G.assert_bit_equal(s, d['fn_dec_scores'], 'scores')
    G.assert_bit_equal(c, d['fn_dec_classes'], 'classes')
    G.assert_bit_equal(b, d['fn_dec_boxes'], 'boxes')
    # the fused sweep (row-group kernel) and a class count that is not a multiple of 4 (raw-tile kernel)
    from b200det import fused
    preds, ann = G.half_nan_inputs(d, 'rn')
    crit = losses.RetinaLoss(**synth.RETINA_KW)
    _, (s, c, b) = fused.EvalStep(crit, decode.RetinaDecoder(**synth.RETINA_KW))(dev(preds), ann.cuda())
    G.assert_bit_equal(s, d['rn_dec_python_nms_scores'], 'fused scores')
    G.assert_bit_equal(b, d['rn_dec_python_nms_boxes'], 'fused boxes')
    odd = synth.make_tie_free(synth.make_fcos_preds(2, 128, 7, seed=50))
    flat = odd[0][0].view(2, -1, 7)
    best = flat[0].max(dim=1).values.argsort(descending=True)
    flat[0, best[0], 6] = float('nan')
    flat[0, best[1], 0] = float('nan')
    with np.errstate(invalid='ignore'):
        (s0, c0, b0), _ = O.fcos_decode(odd, synth.STRIDES)
    s, c, b = decode.FCOSDecoder(strides=synth.STRIDES)(dev(odd))
    G.assert_bit_equal(s, s0, 'C=7 scores')
    G.assert_bit_equal(b, b0, 'C=7 boxes')


def test_host_fast_path_and_python_path_agree():
    """csrc/fastpath.cpp only replaces the Python argument marshalling: both host paths must enqueue
    the same work (bit-identical losses and detections), and inputs outside the fast path's common
    case (non-contiguous / double tensors, a glue call) must fall through to the Python path."""
    from b200det import _lib
    preds = dev(synth.make_tie_free(synth.make_fcos_preds(2, 256, 8, seed=92)))
    ann = synth.make_annotations(2, 12, 256, 8, seed=93).cuda()
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    dec = decode.FCOSDecoder(strides=synth.STRIDES)

    def run(p):
        with torch.no_grad():
            d = crit(p, ann)
        return [d[k].item() for k in ('cls_loss', 'reg_loss', 'center_ness_loss')], dec(p)

    saved = _lib._FAST
    try:
        _lib._FAST = None
        fast_loss, fast_det = run(preds)          # second call: plans exist, the fast path runs
        fast_loss, fast_det = run(preds)
        had_fast = _lib.fastpath() is not None
        _lib._FAST = False
        slow_loss, slow_det = run(preds)
    finally:
        _lib._FAST = saved
    assert fast_loss == slow_loss
    for a, b in zip(fast_det, slow_det):
        assert np.array_equal(a, b)
    # not the common case: permuted (non-contiguous) class tensors and float64 annotations
    odd = [[t.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3) for t in preds[0]], preds[1], preds[2]]
    assert not odd[0][0].is_contiguous()
    with torch.no_grad():
        d = crit(odd, ann.double())
    assert [d[k].item() for k in ('cls_loss', 'reg_loss', 'center_ness_loss')] == slow_loss
    for a, b in zip(dec(odd), slow_det):
        assert np.array_equal(a, b)
    if not had_fast:
        pytest.skip('host fast path not built on this box (Python path verified)')


def test_second_backward_is_refused_not_wrong():
    """The gradients are produced by the forward kernels and scaled in place by the upstream scalars:
    a second pass over the same autograd node would return wrong values, so it raises (as torch does
    for a freed graph); unused loss terms arrive as None and leave their head's gradient alone."""
    preds = dev(synth.make_retina_preds(2, 128, 8, seed=90))
    ann = synth.make_annotations(2, 12, 128, 8, seed=91).cuda()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    d = crit(p, ann)
    d['cls_loss'].backward(retain_graph=True)
    assert p[0][0].grad is not None and p[1][0].grad is None     # reg untouched by cls_loss alone
    with pytest.raises(RuntimeError, match='second time'):
        d['reg_loss'].backward()
    # the supported pattern (tools/scripts.py:918-945): one backward on the summed loss
    q = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    sum(crit(q, ann).values()).backward()
    assert torch.equal(q[0][0].grad, p[0][0].grad) and q[1][0].grad.abs().sum() > 0


def test_non_default_focal_parameters_and_weights():
    preds = synth.make_retina_preds(2, 128, 8, seed=8)
    ann = synth.make_annotations(2, 12, 128, 8, seed=9)
    kw = dict(alpha=0.4, gamma=1.5, beta=0.3, cls_loss_weight=0.7, box_loss_weight=2.5)
    crit = losses.RetinaLoss(**synth.RETINA_KW, **kw)
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
        ref = O.retina_loss(preds, ann, **synth.RETINA_KW, **kw)
    assert_close(loss_values(d, ['cls_loss', 'reg_loss']),
                 [ref['cls_loss'].item(), ref['reg_loss'].item()], LOSS_RTOL, 'RetinaLoss')


def test_fused_eval_glue_matches_numpy_post_processing():
    """boxes /= scale, clip to the image, xyxy -> xywh
    (tools/scripts.py:742-757) fused into the decoder epilogue; bit-exact vs NumPy on the oracle."""
    preds = synth.make_tie_free(synth.make_retina_preds(3, 128, 8, seed=12))
    scales = np.array([0.5, 1.25, 0.8], dtype=np.float32)
    sizes = np.array([[200., 260.], [90., 100.], [150., 120.]], dtype=np.float32)   # (h, w)
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    s, c, b = dec(dev(preds), scales=scales, sizes=sizes, to_xywh=True)
    (s0, c0, b0), _ = O.retina_decode(preds, **synth.RETINA_KW)
    b0 = b0.copy()
    b0 /= np.expand_dims(np.expand_dims(scales, axis=-1), axis=-1)
    for i in range(3):
        b0[i][:, 0] = np.maximum(b0[i][:, 0], 0)
        b0[i][:, 1] = np.maximum(b0[i][:, 1], 0)
        b0[i][:, 2] = np.minimum(b0[i][:, 2], sizes[i][1])
        b0[i][:, 3] = np.minimum(b0[i][:, 3], sizes[i][0])
        b0[i][:, 2:] -= b0[i][:, :2]
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0, 'rescaled / clipped / xywh boxes')
    s1, c1, b1 = dec(dev(preds), scales=scales)       # rescale only
    (_, _, b2), _ = O.retina_decode(preds, **synth.RETINA_KW)
    G.assert_bit_equal(b1, b2 / scales[:, None, None], 'rescaled boxes')


def test_fused_eval_step_matches_separate_calls():
    """Optional extension b200det.fused.EvalStep: one sweep over cls for loss + decode."""
    from b200det import fused
    preds = synth.make_tie_free(synth.make_retina_preds(3, 256, 8, seed=31))
    ann = synth.make_annotations(3, 20, 256, 8, seed=32, empty_images=(1,))
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    with torch.no_grad():
        ref_loss = O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type='GIoU')
    (s0, c0, b0), _ = O.retina_decode(preds, **synth.RETINA_KW)
    loss, (s, c, b) = fused.EvalStep(crit, dec)(dev(preds), ann.cuda())
    assert_close(loss_values(loss, ['cls_loss', 'reg_loss']),
                 [ref_loss['cls_loss'].item(), ref_loss['reg_loss'].item()], LOSS_RTOL, 'fused loss')
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    fp = synth.make_tie_free(synth.make_fcos_preds(2, 256, 8, seed=33))
    fa = synth.make_annotations(2, 20, 256, 8, seed=34)
    fcrit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    fdec = decode.FCOSDecoder(strides=synth.STRIDES)
    with torch.no_grad():
        ref_loss = O.fcos_loss(fp, fa, synth.STRIDES, synth.MI)
    (s0, c0, b0), _ = O.fcos_decode(fp, synth.STRIDES)
    loss, (s, c, b) = fused.EvalStep(fcrit, fdec)(dev(fp), fa.cuda())
    assert_close(loss_values(loss, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 [ref_loss[k].item() for k in ('cls_loss', 'reg_loss', 'center_ness_loss')],
                 LOSS_RTOL, 'fused FCOS loss')
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(b, b0)


def test_heavy_ties_full_size_image_uses_refinement():
    """120 087 rows with identical scores: the multi-CTA front end finds a crowded cut bin and
    the select kernel falls back to the radix refinement + ordered tie pass."""
    preds = synth.make_retina_preds(1, 800, 80, seed=41)
    for c in preds[0]:
        c.fill_(0.01)
        c[..., 17] = 0.25
    preds[0][0][0, 50, 50, 4, 3] = 0.6
    preds[0][2][0, 3, 3, 1, 9] = 0.6          # two equal winners: lower row first
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **synth.RETINA_KW)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    check_decode_details(info, extra['per_image'], 1000)


def test_generic_anchor_layout_non_square_many_gt():
    """Code paths the COCO configs never touch: 6 anchors per location (generic, not the 9-anchor
    register kernel), non-square feature maps, odd strides / sizes, and 300 annotation rows per
    image (several compaction rounds in the assignment kernel)."""
    kw = dict(areas=[[24, 40], [48, 80], [96, 160]], ratios=[0.5, 1, 2], scales=[1, 1.5],
              strides=[6, 12, 24])
    gen = torch.Generator().manual_seed(5)
    B, C = 2, 12
    shapes = [(21, 37), (11, 19), (6, 10)]
    cls = [torch.sigmoid(torch.randn((B, h, w, 6, C), generator=gen) - 3.0) for h, w in shapes]
    reg = [torch.randn((B, h, w, 6, 4), generator=gen) * 0.2 for h, w in shapes]
    preds = synth.make_tie_free([cls, reg])
    ann = synth.make_annotations(B, 300, 220, C, seed=6, min_gt=280)
    ann[..., 0:4] *= torch.tensor([1.0, 0.55, 1.0, 0.55])      # fit the 220 x 126 "image"
    for box_type in ('SmoothL1', 'EIoU'):
        crit = losses.RetinaLoss(**kw, box_loss_type=box_type)
        with torch.no_grad():
            d = crit(dev(preds), ann.cuda())
            ref = O.retina_loss(preds, ann, **kw, box_loss_type=box_type)
        got = crit.debug_assign(dev(preds), ann.cuda())
        assert np.array_equal(got['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
        assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
        fast = crit.debug_assign(dev(preds), ann.cuda(), exact=False)
        assert np.array_equal(fast['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
        assert ref['num_pos'] > 0
        assert_close(loss_values(d, ['cls_loss', 'reg_loss']),
                     [ref['cls_loss'].item(), ref['reg_loss'].item()], LOSS_RTOL, box_type)
    dec = decode.RetinaDecoder(**kw, topn=400, max_object_num=60)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **kw, topn=400, max_object_num=60)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    check_decode_details(info, extra['per_image'], 400)
    # FCOS on the same odd pyramid, 3 levels, custom ranges and radius
    fkw = dict(strides=[6, 12, 24], mi=[[-1, 40], [40, 90], [90, 100000000]])
    fcls = [torch.sigmoid(torch.randn((B, h, w, C), generator=gen) - 3.0) for h, w in shapes]
    freg = [torch.randn((B, h, w, 4), generator=gen) * 0.5 + 2.5 for h, w in shapes]
    fctr = [torch.sigmoid(torch.randn((B, h, w, 1), generator=gen)) for h, w in shapes]
    fpreds = synth.make_tie_free([fcls, freg, fctr])
    fcrit = losses.FCOSLoss(**fkw, center_sample_radius=2.0, box_loss_iou_type='DIoU')
    with torch.no_grad():
        d = fcrit(dev(fpreds), ann.cuda())
        ref = O.fcos_loss(fpreds, ann, fkw['strides'], fkw['mi'], center_sample_radius=2.0,
                          box_loss_iou_type='DIoU')
    got = fcrit.debug_assign(dev(fpreds), ann.cuda())
    assert_targets_equal(got['targets'].cpu().numpy(), ref['targets'].numpy(), 'targets')
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    assert_close(loss_values(d, ['cls_loss', 'reg_loss', 'center_ness_loss']),
                 [ref[k].item() for k in ('cls_loss', 'reg_loss', 'center_ness_loss')], LOSS_RTOL,
                 'FCOS odd pyramid')
    fdec = decode.FCOSDecoder(strides=fkw['strides'], topn=300, max_object_num=40)
    (s, c, b), info = fdec.decode_with_details(dev(fpreds))
    (s0, c0, b0), extra = O.fcos_decode(fpreds, fkw['strides'], topn=300, max_object_num=40)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(b, b0)
    check_decode_details(info, extra['per_image'], 300)


def test_training_and_eval_loops_like_the_reference_scripts():
    """Stub of the reference's loops with the drop-in classes: a tiny 'model' produces head
    outputs, train steps follow tools/scripts.py:893-945 (criterion -> sum(values) -> backward ->
    optimizer step; loss must go down), the eval step follows :733-758 (criterion, decoder,
    in-place rescale / clip of the returned NumPy arrays)."""
    torch.manual_seed(0)
    B, C, S = 2, 8, 128
    sizes = synth.pyramid_sizes(S)
    feats = [torch.randn(B, 16, p, p, device='cuda') for p in sizes]

    class Heads(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.cls = torch.nn.Conv2d(16, 9 * C, 3, padding=1)
            self.reg = torch.nn.Conv2d(16, 9 * 4, 3, padding=1)
            torch.nn.init.constant_(self.cls.bias, -4.595)

        def forward(self, fs):
            cls, reg = [], []
            for f in fs:
                c = self.cls(f).permute(0, 2, 3, 1).contiguous()
                r = self.reg(f).permute(0, 2, 3, 1).contiguous()
                cls.append(torch.sigmoid(c).view(B, f.shape[2], f.shape[3], 9, C).float())
                reg.append(r.view(B, f.shape[2], f.shape[3], 9, 4))
            return [cls, reg]

    model = Heads().cuda()
    criterion = losses.__dict__['RetinaLoss'](**synth.RETINA_KW).cuda()
    decoder = decode.__dict__['RetinaDecoder'](**synth.RETINA_KW)
    annots = synth.make_annotations(B, 6, S, C, seed=4, min_gt=4).cuda()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    history = []
    for _ in range(12):
        opt.zero_grad()
        loss_value = criterion(model(feats), annots)
        loss = sum(loss_value.values())
        assert not (torch.isnan(loss) or torch.isinf(loss)) and loss != 0.
        loss.backward()
        opt.step()
        history.append(loss.item())
    assert history[-1] < 0.8 * history[0], history
    model.eval()
    with torch.no_grad():
        outs = model(feats)
        loss_value = criterion(outs, annots)
        scores, classes, boxes = decoder(outs)
    assert set(loss_value) == {'cls_loss', 'reg_loss'}
    scales = np.array([0.5, 0.8], dtype=np.float32)
    boxes /= np.expand_dims(np.expand_dims(scales, axis=-1), axis=-1)      # scripts.py:742
    boxes[0][:, 0] = np.maximum(boxes[0][:, 0], 0)                          # scripts.py:749
    assert scores.shape == (B, 100) and boxes.shape == (B, 100, 4)


def test_retinaface_drop_in_golden():
    """RetinaFaceLoss / RetinaFaceDecoder reuse the detection kernels
    (square anchors, 0.35 / 0.35 thresholds, one class); checked against the reference's own
    face_detection classes (golden vectors)."""
    from b200det.face_detection import losses as face_losses, decode as face_decode
    face = G.load('retinaface_small.npz')
    preds, ann = G.retina_inputs(face)
    sizes, strides = [[8, 16, 32], [32, 64, 128], [128, 256, 512]], [8, 16, 32]
    for box_type in ['SmoothL1'] + G.IOU_TYPES:
        crit = face_losses.__dict__['RetinaFaceLoss'](anchor_sizes=sizes, strides=strides,
                                                      box_loss_type=box_type)
        with torch.no_grad():
            d = crit(dev(preds), ann.cuda())
        assert_close(loss_values(d, ['cls_loss', 'reg_loss']), face[f'loss_{box_type}'],
                     LOSS_RTOL, f'RetinaFaceLoss {box_type}')
    crit = face_losses.RetinaFaceLoss(anchor_sizes=sizes, strides=strides)
    got = crit.debug_assign(dev(preds), ann.cuda())
    assert np.array_equal(got['labels'].cpu().numpy(), face['assign'][..., 4].astype(np.int32))
    fast = crit.debug_assign(dev(preds), ann.cuda(), exact=False)
    assert np.array_equal(fast['labels'].cpu().numpy(), face['assign'][..., 4].astype(np.int32))
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
    d = crit(p, ann.cuda())
    (d['cls_loss'] + 2.0 * d['reg_loss']).backward()
    for i in range(3):
        np.testing.assert_allclose(p[0][i].grad.cpu().numpy(), face[f'gcls_{i}'], rtol=GRAD_RTOL,
                                   atol=GRAD_ATOL)
        np.testing.assert_allclose(p[1][i].grad.cpu().numpy(), face[f'greg_{i}'], rtol=GRAD_RTOL,
                                   atol=GRAD_ATOL)
    for nms in ('python_nms', 'diou_python_nms'):
        dec = face_decode.__dict__['RetinaFaceDecoder'](anchor_sizes=sizes, strides=strides,
                                                        nms_type=nms)
        s, c, b = dec(dev(preds))
        G.assert_bit_equal(s, face[f'dec_{nms}_scores'], 'scores')
        G.assert_bit_equal(c, face[f'dec_{nms}_classes'], 'classes')
        G.assert_bit_equal(b, face[f'dec_{nms}_boxes'], 'boxes')


def test_limits_topn_2048_many_outputs_2048_gt_single_level():
    """Compiled limits: topn = 2048 (sort of 4096 entries), max_object_num = 300, 2048 annotation
    rows per image (98 KB of staged GT, 16 compaction rounds), single-level pyramid, batch 1."""
    kw = dict(areas=[[32, 32]], ratios=[0.5, 1, 2], scales=[2**0, 2**(1.0 / 3.0), 2**(2.0 / 3.0)],
              strides=[8])
    gen = torch.Generator().manual_seed(9)
    cls = [torch.sigmoid(torch.randn((1, 40, 40, 9, 4), generator=gen) - 1.0)]
    reg = [torch.randn((1, 40, 40, 9, 4), generator=gen) * 0.2]
    preds = synth.make_tie_free([cls, reg])
    ann = synth.make_annotations(1, 2048, 320, 4, seed=10, min_gt=2000)
    ann[..., 0:4] = ann[..., 0:4].clamp(max=320.)
    crit = losses.RetinaLoss(**kw, box_loss_type='IoU')
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
        ref = O.retina_loss(preds, ann, **kw, box_loss_type='IoU')
    got = crit.debug_assign(dev(preds), ann.cuda())
    assert np.array_equal(got['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    assert_close(loss_values(d, ['cls_loss', 'reg_loss']),
                 [ref['cls_loss'].item(), ref['reg_loss'].item()], LOSS_RTOL, 'RetinaLoss G=2048')
    dec = decode.RetinaDecoder(**kw, topn=2048, max_object_num=300, nms_threshold=0.7)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **kw, topn=2048, max_object_num=300,
                                          nms_threshold=0.7)
    G.assert_bit_equal(s, s0)
    G.assert_bit_equal(c, c0)
    G.assert_bit_equal(b, b0)
    check_decode_details(info, extra['per_image'], 2048)
    assert info['counts'][0, 1] == 2048
    with pytest.raises(ValueError):
        decode.RetinaDecoder(**kw, topn=4096)
    with pytest.raises(ValueError):
        crit(dev(preds), torch.full((1, 3000, 5), -1.).cuda())


def test_nan_and_inf_inputs_reach_the_loss_like_the_reference():
    """The reference's training loop skips steps whose loss is NaN / inf (tools/scripts.py:922-930);
    torch.clamp propagates NaN (losses.py:196), so a NaN probability or a NaN / overflowing box
    regression at a positive row must make the matching loss non-finite here too, while values at
    rows the loss never reads (regression of negatives) must not."""
    from b200det import fused
    clean = synth.make_retina_preds(2, 128, 8, seed=41)
    ann = synth.make_annotations(2, 6, 128, 8, seed=42, min_gt=3)
    kw = dict(**synth.RETINA_KW, box_loss_type='GIoU')
    crit = losses.RetinaLoss(**kw)
    with torch.no_grad():
        ref = O.retina_loss(clean, ann, **kw)
    labels = ref['labels'].numpy()
    n0 = clean[0][0][0].numel() // 8                      # rows of level 0 per image
    pos = int(np.nonzero(labels[1, :n0] > 0)[0][0])
    neg = int(np.nonzero(labels[1, :n0] == 0)[0][0])

    def variant(kind, row, value):
        p = [[t.clone() for t in grp] for grp in clean]
        grp = 0 if kind == 'cls' else 1
        width = 8 if kind == 'cls' else 4
        p[grp][0].view(2, -1, width)[1, row, 1] = value
        return p

    cases = [('cls', neg, float('nan')), ('cls', pos, float('nan')), ('reg', pos, float('nan')),
             ('reg', neg, float('nan')), ('reg', pos, 200.0), ('reg', neg, float('inf'))]
    for kind, row, value in cases:
        p = variant(kind, row, value)
        with torch.no_grad():
            want = O.retina_loss(p, ann, **kw)
            got = crit(dev(p), ann.cuda())
            both, _ = fused.EvalStep(crit, decode.RetinaDecoder(**synth.RETINA_KW))(dev(p), ann.cuda())
        pt = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(p)]
        train = crit(pt, ann.cuda())
        for name, d in (('eval', got), ('fused', both), ('train', train)):
            for k in ('cls_loss', 'reg_loss'):
                w, g = want[k].item(), d[k].item()
                assert np.isfinite(w) == np.isfinite(g), f'{name} {kind} row {row} = {value}: {k} {g} vs {w}'
                if np.isfinite(w):
                    assert_close([g], [w], LOSS_RTOL, f'{name} {k}')
    # FCOS: centre-ness NaN at a positive point
    fclean = synth.make_fcos_preds(2, 128, 8, seed=43)
    fcrit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with torch.no_grad():
        fref = O.fcos_loss(fclean, ann, synth.STRIDES, synth.MI)
    fpos = int(np.nonzero(fref['labels'].numpy()[0, :256] > 0)[0][0])
    fp = [[t.clone() for t in grp] for grp in fclean]
    fp[2][0].view(2, -1)[0, fpos] = float('nan')
    with torch.no_grad():
        want = O.fcos_loss(fp, ann, synth.STRIDES, synth.MI)
        got = fcrit(dev(fp), ann.cuda())
    assert np.isnan(want['center_ness_loss'].item()) and np.isnan(got['center_ness_loss'].item())
    assert_close(loss_values(got, ['cls_loss', 'reg_loss']),
                 [want['cls_loss'].item(), want['reg_loss'].item()], LOSS_RTOL, 'FCOS other losses')


def test_cpu_tensors_are_rejected():
    preds = synth.make_retina_preds(1, 128, 8, seed=8)
    ann = synth.make_annotations(1, 4, 128, 8, seed=9)
    with pytest.raises(RuntimeError, match='no CPU path'):
        losses.RetinaLoss(**synth.RETINA_KW)(preds, ann)
    with pytest.raises(RuntimeError, match='no CPU path'):
        decode.RetinaDecoder(**synth.RETINA_KW)(preds)


# ------------------------------------------------------------------------------------------
# size-independent properties at the benchmark shape
# ------------------------------------------------------------------------------------------
def test_sharding_linearity_full_size():
    """Sums over an image-sharded batch add up to the unsharded sums (what the multi-GPU
    all-reduce relies on); labels are per-image, so shard labels concatenate exactly."""
    B = 8
    preds = synth.make_retina_preds(B, 800, 80, seed=0, device='cuda')
    ann = synth.make_annotations(B, 100, 800, 80, seed=1).cuda()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    with torch.no_grad():
        crit(preds, ann)
        full = crit.last_stats['sums'].cpu().numpy().copy()
        full_labels = crit.debug_assign(preds, ann)['labels'].cpu().numpy()
        parts, part_labels = [], []
        for lo in (0, 3):
            hi = 3 if lo == 0 else B
            shard = [[t[lo:hi].contiguous() for t in grp] for grp in preds]
            crit(shard, ann[lo:hi].contiguous())
            parts.append(crit.last_stats['sums'].cpu().numpy().copy())
            part_labels.append(crit.debug_assign(shard, ann[lo:hi].contiguous())['labels'].cpu().numpy())
    assert parts[0][0] + parts[1][0] == full[0]
    assert np.array_equal(np.concatenate(part_labels, axis=0), full_labels)
    np.testing.assert_allclose(parts[0][1:] + parts[1][1:], full[1:], rtol=1e-6)


def test_decoder_output_invariants_full_size():
    B = 4
    preds = synth.make_retina_preds(B, 800, 80, seed=3, device='cuda')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    (s, c, b), info = dec.decode_with_details(preds)
    assert s.shape == (B, 100) and c.shape == (B, 100) and b.shape == (B, 100, 4)
    for i in range(B):
        n = int((s[i] >= 0).sum())
        assert (np.diff(s[i, :n]) <= 0).all()            # sorted by score
        assert (s[i, :n] > np.float32(0.05)).all()
        assert (s[i, n:] == -1).all() and (c[i, n:] == -1).all() and (b[i, n:] == 0).all()
        assert (b[i, :n] == np.trunc(b[i, :n])).all()     # integer-valued coordinates
        order = info['order'][i]
        assert len(set(order[order >= 0].tolist())) == int((order >= 0).sum())
        keep = info['keep'][i]
        assert (np.diff(keep[keep >= 0]) > 0).all()
    # idempotence: same inputs -> bit-identical outputs
    s2, c2, b2 = dec(preds)
    G.assert_bit_equal(s, s2)
    G.assert_bit_equal(b, b2)


def test_batch_with_more_than_2_31_elements_per_level():
    """Maximum sizes: batch 384 at 800x800 / 80 classes puts 2.76e9 floats (> 2^31) into the first
    pyramid level alone, so element offsets need 64 bits everywhere.  Size-independent checks:
    the unsharded sums equal the sum of two half-batch calls, decoder outputs of images from both
    ends of the batch equal decoding those images alone, and the training path's gradient of the
    last image equals the half-batch gradient rescaled by the positive counts."""
    B, H = 384, 192
    preds = synth.make_retina_preds(B, 800, 80, seed=5, device='cuda')
    ann = synth.make_annotations(B, 100, 800, 80, seed=6).cuda()
    assert preds[0][0].numel() > 2**31
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    halves = [[[t[lo:lo + H] for t in grp] for grp in preds] for lo in (0, H)]   # contiguous views
    with torch.no_grad():
        crit(preds, ann)
        full = crit.last_stats['sums'].cpu().numpy().copy()
        parts = []
        for i, lo in enumerate((0, H)):
            crit(halves[i], ann[lo:lo + H])
            parts.append(crit.last_stats['sums'].cpu().numpy().copy())
    assert parts[0][0] + parts[1][0] == full[0] > 0
    np.testing.assert_allclose(parts[0][1:] + parts[1][1:], full[1:], rtol=1e-6)

    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    s, c, b = dec(preds)
    for img in (0, H - 1, H, B - 1):
        one = [[t[img:img + 1] for t in grp] for grp in preds]
        s1, c1, b1 = dec(one)
        G.assert_bit_equal(s[img], s1[0], f'scores of image {img}')
        G.assert_bit_equal(c[img], c1[0], f'classes of image {img}')
        G.assert_bit_equal(b[img], b1[0], f'boxes of image {img}')

    # training path on the big batch (writes a second 2.76e9-element tensor)
    cls0 = preds[0][0].requires_grad_(True)
    d = crit(preds, ann)
    d['cls_loss'].backward()
    g_last = cls0.grad[B - 1].clone()
    cls0.grad = None
    preds[0][0].requires_grad_(False)
    half_cls0 = halves[1][0][0].detach().requires_grad_(True)
    shard = [[half_cls0] + [t.detach() for t in halves[1][0][1:]], halves[1][1]]
    crit(shard, ann[H:])['cls_loss'].backward()
    scale = parts[1][0] / full[0]
    torch.testing.assert_close(g_last, half_cls0.grad[H - 1] * scale, rtol=1e-5, atol=1e-12)


def test_peer_exchange_two_gpus():
    """The loss normaliser exchanged over NVLink peer memory inside the reduction kernel
    (sync_normalizer='p2p', csrc/exchange.cu) against NCCL and against the oracle's unsharded
    loss: tests/p2p_check.py under torchrun.  Needs >= 2 GPUs (skipped on the 1-GPU box)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port',
                          '29517', os.path.join(root, 'tests', 'p2p_check.py')],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
