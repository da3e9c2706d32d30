"""Randomised GPU parity: seeded random pyramids (1-5 levels, non-square maps, odd strides), anchor
layouts (1-9 per location, incl. 9-anchor layouts that are not the COCO one), class counts (also
not multiples of 4), GT counts, loss types, focal parameters, decoder thresholds / top-n / NMS
types -- CUDA path vs the oracle, same bars as test_gpu_parity.py (assignments, top-n order, keep
lists, boxes bit-exact; losses 1e-5 relative; gradients 1e-4 relative to the largest entry)."""
import math

import numpy as np
import pytest
import torch

from b200det import synth, losses, decode
from oracle import det_oracle as O

import golden_util as G
from test_gpu_parity import (LOSS_RTOL, assert_close, assert_targets_equal, check_decode_details,
                             dev, loss_values)

pytestmark = pytest.mark.gpu

CLASS_COUNTS = [1, 2, 3, 4, 7, 20, 33, 80, 91, 92, 364, 640]
NMS_TYPES = ['python_nms', 'diou_python_nms', 'torch_nms']


def random_pyramid(rng):
    n_levels = int(rng.randint(1, 6))
    h, w = int(rng.randint(3, 41)), int(rng.randint(3, 41))
    s0 = float(rng.choice([4, 6, 8, 10]))
    shapes, strides = [], []
    for l in range(n_levels):
        shapes.append((h, w))
        strides.append(s0 * 2**l)
        h, w = (h + 1) // 2, (w + 1) // 2
    return shapes, strides


def random_annotations(rng, batch, shapes, strides, num_classes):
    width = shapes[0][1] * strides[0]
    height = shapes[0][0] * strides[0]
    max_gt = int(rng.choice([1, 3, 17, 40, 130]))
    empty = tuple(b for b in range(batch) if rng.rand() < 0.2)
    ann = synth.make_annotations(batch, max_gt, max(int(width), 17), num_classes,
                                 seed=int(rng.randint(1 << 30)), empty_images=empty)
    valid = ann[..., 4] >= 0
    ann[..., 1] = torch.where(valid, ann[..., 1] * (height / width), ann[..., 1])
    ann[..., 3] = torch.where(valid, torch.maximum(ann[..., 3] * (height / width),
                                                   ann[..., 1] + 1), ann[..., 3])
    return ann


def grads_close(got, want, what):
    got = got.detach().cpu().numpy().astype(np.float64)
    want = want.detach().cpu().numpy().astype(np.float64)
    scale = max(np.abs(want).max(), 1e-12)
    err = np.abs(got - want).max()
    assert err <= 1e-4 * scale + 1e-9, f'{what}: max abs err {err} at gradient scale {scale}'


@pytest.mark.parametrize('seed', range(32))
def test_fuzz_retina(seed):
    rng = np.random.RandomState(1000 + seed)
    shapes, strides = random_pyramid(rng)
    ratios = list(rng.choice([0.4, 0.5, 1, 2, 3], size=int(rng.randint(1, 4)), replace=False))
    scales = list(rng.choice([1, 1.26, 1.5, 2], size=int(rng.randint(1, 4)), replace=False))
    per_loc = len(ratios) * len(scales)
    areas = [[4 * s, 4 * s] for s in strides]
    kw = dict(areas=areas, ratios=[float(r) for r in ratios], scales=[float(s) for s in scales],
              strides=strides)
    B = int(rng.randint(1, 6))
    C = int(rng.choice(CLASS_COUNTS))
    mean = float(rng.choice([-4.595, -3.0, -1.5]))
    gen = torch.Generator().manual_seed(seed)
    cls = [torch.sigmoid(torch.randn((B, h, w, per_loc, C), generator=gen) + mean)
           for h, w in shapes]
    reg = [torch.randn((B, h, w, per_loc, 4), generator=gen) * float(rng.choice([0.1, 0.5]))
           for h, w in shapes]
    min_score = float(rng.choice([0.01, 0.05, 0.3]))
    preds = synth.make_tie_free([cls, reg], min_score=min_score)
    ann = random_annotations(rng, B, shapes, strides, C)

    box_type = str(rng.choice(['SmoothL1'] + G.IOU_TYPES))
    lkw = dict(alpha=float(rng.choice([0.25, 0.4])), gamma=float(rng.choice([2.0, 1.5])),
               beta=float(rng.choice([1 / 9, 0.5])), cls_loss_weight=float(rng.choice([1., 0.7])),
               box_loss_weight=float(rng.choice([1., 2.])), box_loss_type=box_type)
    crit = losses.RetinaLoss(**kw, **lkw)
    got = crit.debug_assign(dev(preds), ann.cuda())
    fast = crit.debug_assign(dev(preds), ann.cuda(), exact=False)
    p_ref = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    ref = O.retina_loss(p_ref, ann, **kw, **lkw)
    assert np.array_equal(got['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    assert np.array_equal(fast['labels'].cpu().numpy(), ref['labels'].numpy().astype(np.int32))
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
    want = [ref['cls_loss'].item(), ref['reg_loss'].item()]
    if ref['num_pos'] == 0:
        assert loss_values(d, ['cls_loss', 'reg_loss']).tolist() == [0., 0.]
    else:
        assert_close(loss_values(d, ['cls_loss', 'reg_loss']), want, LOSS_RTOL,
                     f'RetinaLoss {box_type} {lkw}')
        # training path: same values, gradients like the oracle's autograd
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
        dt = crit(p, ann.cuda())
        assert_close(loss_values(dt, ['cls_loss', 'reg_loss']), want, LOSS_RTOL, 'training fwd')
        (dt['cls_loss'] + 1.5 * dt['reg_loss']).backward()
        (ref['cls_loss'] + 1.5 * ref['reg_loss']).backward()
        for i in range(len(shapes)):
            grads_close(p[0][i].grad, p_ref[0][i].grad, f'cls grad level {i}')
            grads_close(p[1][i].grad, p_ref[1][i].grad, f'reg grad level {i}')

    dkw = dict(topn=int(rng.choice([5, 50, 300, 1000])),
               max_object_num=int(rng.choice([1, 10, 100, 300])),
               min_score_threshold=min_score, nms_type=str(rng.choice(NMS_TYPES)),
               nms_threshold=float(rng.choice([0.3, 0.5, 0.7])))
    dec = decode.RetinaDecoder(**kw, **dkw)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.retina_decode(preds, **kw, **dkw)
    G.assert_bit_equal(s, s0, f'scores {dkw}')
    G.assert_bit_equal(c, c0, f'classes {dkw}')
    G.assert_bit_equal(b, b0, f'boxes {dkw}')
    check_decode_details(info, extra['per_image'], dkw['topn'])
    if True:   # fused eval step: the row-group sweeps (lane split depends on C) / the raw-tile sweep (C % 4 != 0)
        from b200det import fused
        loss, (s1, c1, b1) = fused.EvalStep(crit, dec)(dev(preds), ann.cuda())
        G.assert_bit_equal(s1, s0, 'fused scores')
        G.assert_bit_equal(c1, c0, 'fused classes')
        G.assert_bit_equal(b1, b0, 'fused boxes')
        if ref['num_pos'] > 0:
            assert_close(loss_values(loss, ['cls_loss', 'reg_loss']), want, LOSS_RTOL, 'fused loss')


@pytest.mark.parametrize('seed', range(32))
def test_fuzz_fcos(seed):
    rng = np.random.RandomState(2000 + seed)
    shapes, strides = random_pyramid(rng)
    n = len(shapes)
    edges = [-1] + [float(strides[l] * 8) for l in range(n - 1)] + [100000000]
    mi = [[edges[l], edges[l + 1]] for l in range(n)]
    B = int(rng.randint(1, 6))
    C = int(rng.choice(CLASS_COUNTS))
    mean = float(rng.choice([-4.595, -3.0, -1.5]))
    gen = torch.Generator().manual_seed(100 + seed)
    cls = [torch.sigmoid(torch.randn((B, h, w, C), generator=gen) + mean) for h, w in shapes]
    reg = [torch.randn((B, h, w, 4), generator=gen) * 0.5 + math.log(strides[l])
           for l, (h, w) in enumerate(shapes)]
    ctr = [torch.sigmoid(torch.randn((B, h, w, 1), generator=gen)) for h, w in shapes]
    min_score = float(rng.choice([0.01, 0.05, 0.3]))
    preds = synth.make_tie_free([cls, reg, ctr], min_score=min_score)
    ann = random_annotations(rng, B, shapes, strides, C)

    iou_type = str(rng.choice(G.IOU_TYPES))
    lkw = dict(center_sample_radius=float(rng.choice([1.5, 1.0, 2.5])),
               use_center_sample=bool(rng.rand() < 0.7),
               alpha=float(rng.choice([0.25, 0.4])), gamma=float(rng.choice([2.0, 1.5])),
               cls_loss_weight=float(rng.choice([1., 0.7])),
               box_loss_weight=float(rng.choice([1., 2.])),
               center_ness_loss_weight=float(rng.choice([1., 0.5])),
               box_loss_iou_type=iou_type)
    crit = losses.FCOSLoss(strides=strides, mi=mi, **lkw)
    got = crit.debug_assign(dev(preds), ann.cuda())
    p_ref = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    ref = O.fcos_loss(p_ref, ann, strides, mi, **lkw)
    assert_targets_equal(got['targets'].cpu().numpy(), ref['targets'].numpy(), 'targets')
    assert np.array_equal(got['matched'].cpu().numpy(), ref['matched'].numpy().astype(np.int32))
    keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    want = [ref[k].item() for k in keys]
    with torch.no_grad():
        d = crit(dev(preds), ann.cuda())
    if ref['num_pos'] == 0:
        assert loss_values(d, keys).tolist() == [0., 0., 0.]
    else:
        assert_close(loss_values(d, keys), want, LOSS_RTOL, f'FCOSLoss {lkw}')
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
        dt = crit(p, ann.cuda())
        assert_close(loss_values(dt, keys), want, LOSS_RTOL, 'training fwd')
        (dt['cls_loss'] + 1.5 * dt['reg_loss'] + 0.5 * dt['center_ness_loss']).backward()
        (ref['cls_loss'] + 1.5 * ref['reg_loss'] + 0.5 * ref['center_ness_loss']).backward()
        for i in range(n):
            grads_close(p[0][i].grad, p_ref[0][i].grad, f'cls grad level {i}')
            grads_close(p[1][i].grad, p_ref[1][i].grad, f'reg grad level {i}')
            grads_close(p[2][i].grad, p_ref[2][i].grad, f'ctr grad level {i}')

    dkw = dict(topn=int(rng.choice([5, 50, 300, 1000])),
               max_object_num=int(rng.choice([1, 10, 100, 300])),
               min_score_threshold=min_score, nms_type=str(rng.choice(NMS_TYPES)),
               nms_threshold=float(rng.choice([0.3, 0.5, 0.7])))
    dec = decode.FCOSDecoder(strides=strides, **dkw)
    (s, c, b), info = dec.decode_with_details(dev(preds))
    (s0, c0, b0), extra = O.fcos_decode(preds, strides, **dkw)
    G.assert_bit_equal(s, s0, f'scores {dkw}')
    G.assert_bit_equal(c, c0, f'classes {dkw}')
    G.assert_bit_equal(b, b0, f'boxes {dkw}')
    check_decode_details(info, extra['per_image'], dkw['topn'])
    if True:   # every class count: C % 4 != 0 takes the raw-tile fused sweep
        from b200det import fused
        loss, (s1, c1, b1) = fused.EvalStep(crit, dec)(dev(preds), ann.cuda())
        G.assert_bit_equal(s1, s0, 'fused scores')
        G.assert_bit_equal(c1, c0, 'fused classes')
        G.assert_bit_equal(b1, b0, 'fused boxes')
        if ref['num_pos'] > 0:
            assert_close(loss_values(loss, keys), want, LOSS_RTOL, 'fused loss')


@pytest.mark.parametrize('kind', ['retina', 'fcos'])
@pytest.mark.parametrize('C', [5, 7, 13, 91, 365])
def test_training_sweep_class_counts_not_multiple_of_4(kind, C, monkeypatch):
    """Class counts that are not a multiple of 4 on levels that hold a multiple of 4 floats: the
    label-aware sweep reads 128-bit units that straddle rows (focal.cu XROW).  Loss values and
    gradients against the oracle's autograd, and against the scalar kernel it replaces; many
    positives and (Retina) ignored rows so that target classes and row boundaries meet inside units."""
    rng = np.random.RandomState(C)
    shapes, strides = [(12, 8), (6, 4), (2, 2)], [8., 16., 32.]
    B = 3
    gen = torch.Generator().manual_seed(C)
    if kind == 'retina':
        kw = dict(areas=[[32, 32], [64, 64], [128, 128]], ratios=[0.5, 1., 2.], scales=[1., 1.5],
                  strides=strides)
        cls = [torch.sigmoid(torch.randn((B, h, w, 6, C), generator=gen) - 2.0) for h, w in shapes]
        reg = [torch.randn((B, h, w, 6, 4), generator=gen) * 0.2 for h, w in shapes]
        preds = [cls, reg]
        crit = losses.RetinaLoss(**kw, box_loss_type='SmoothL1')
        keys = ['cls_loss', 'reg_loss']
    else:
        kw = dict(strides=strides, mi=[[-1, 64], [64, 128], [128, 100000]])
        cls = [torch.sigmoid(torch.randn((B, h, w, C), generator=gen) - 2.0) for h, w in shapes]
        reg = [torch.randn((B, h, w, 4), generator=gen) * 0.3 + math.log(s) for (h, w), s in zip(shapes, strides)]
        ctr = [torch.sigmoid(torch.randn((B, h, w, 1), generator=gen)) for h, w in shapes]
        preds = [cls, reg, ctr]
        crit = losses.FCOSLoss(**kw)
        keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    for t in cls:                     # some probabilities outside the clamp and above the fast range
        flat = t.view(-1)
        idx = torch.from_numpy(rng.randint(0, flat.numel(), size=40))
        flat[idx[:10]] = 0.
        flat[idx[10:20]] = 1.
        flat[idx[20:]] = 0.6
    ann = synth.make_annotations(B, 12, 96, C, seed=C + 1)
    ann[..., 1] *= 64 / 96.
    ann[..., 3] = torch.where(ann[..., 4] >= 0, torch.maximum(ann[..., 3] * (64 / 96.), ann[..., 1] + 1),
                              ann[..., 3])
    p_ref = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    ref = (O.retina_loss(p_ref, ann, **kw, box_loss_type='SmoothL1') if kind == 'retina'
           else O.fcos_loss(p_ref, ann, **kw))
    assert ref['num_pos'] > 0
    sum(ref[k] for k in keys).backward()
    want = [ref[k].item() for k in keys]
    results = []
    for scalar in (False, True):
        if scalar:
            monkeypatch.setenv('B200DET_FOCAL_NO_XROW', '1')
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in dev(preds)]
        d = crit(p, ann.cuda())
        assert_close(loss_values(d, keys), want, LOSS_RTOL, f'{kind} C={C} scalar={scalar}')
        sum(d[k] for k in keys).backward()
        for i in range(len(shapes)):
            grads_close(p[0][i].grad, p_ref[0][i].grad, f'cls grad level {i} scalar={scalar}')
        results.append([t.grad.clone() for t in p[0]])
    for a, b in zip(*results):
        grads_close(a, b, 'XROW vs scalar kernel')


@pytest.mark.parametrize('seed', range(24))
def test_tile_assignment_stress(seed):
    """The production assignment (GT-centric tile kernel: candidate ranges derived from conservative
    bounds) against the exhaustive anchor-centric scan, on inputs built to sit on its edges: GT boxes
    that are jittered copies of actual anchors (IoU around the 0.4 / 0.5 thresholds and the 0.38 floor),
    boxes straddling tile borders (tiles are <= 32 locations wide: the pyramids here are up to 70 wide),
    duplicated / degenerate / huge / sub-pixel boxes, odd strides, 1-12 anchors per location, and more
    annotation rows than one work-item round holds."""
    rng = np.random.RandomState(7000 + seed)
    n_levels = int(rng.randint(1, 5))
    h, w = int(rng.randint(5, 71)), int(rng.randint(5, 71))
    s0 = float(rng.choice([4, 6, 8]))
    shapes, strides = [], []
    for l in range(n_levels):
        shapes.append((h, w))
        strides.append(s0 * 2**l)
        h, w = (h + 1) // 2, (w + 1) // 2
    ratios = [float(r) for r in rng.choice([0.33, 0.5, 1, 2, 3], size=int(rng.randint(1, 5)), replace=False)]
    scales = [float(s) for s in rng.choice([1, 1.26, 1.59], size=int(rng.randint(1, 4)), replace=False)]
    per_loc = len(ratios) * len(scales)
    kw = dict(areas=[[4 * s, 4 * s] for s in strides], ratios=ratios, scales=scales, strides=strides)
    B, C = int(rng.randint(1, 4)), 5
    G_rows = int(rng.choice([8, 40, 150, 300]))
    crit = losses.RetinaLoss(**kw)
    cls = [torch.full((B, hh, ww, per_loc, C), 0.01) for hh, ww in shapes]
    reg = [torch.zeros((B, hh, ww, per_loc, 4)) for hh, ww in shapes]
    from b200det import anchor as A
    tables = A.RetinaAnchors(**kw)([[ww, hh] for hh, ww in shapes])
    flat = np.concatenate([t.reshape(-1, 4) for t in tables], axis=0)
    width, height = shapes[0][1] * strides[0], shapes[0][0] * strides[0]
    ann = np.full((B, G_rows, 5), -1, dtype=np.float32)
    for b in range(B):
        n = int(rng.randint(0, G_rows + 1))
        rows = np.sort(rng.permutation(G_rows)[:n])
        for j, r in enumerate(rows):
            kind = rng.randint(0, 10)
            a = flat[rng.randint(0, flat.shape[0])]
            if kind < 6:      # jittered copy of an anchor: IoU anywhere between 0.2 and 1
                aw, ah = a[2] - a[0], a[3] - a[1]
                jit = rng.uniform(-0.35, 0.35, size=4) * np.array([aw, ah, aw, ah])
                box = a + jit
            elif kind == 6:   # exact anchor (IoU 1) -- and sometimes a duplicate of the previous row
                box = a.copy() if j == 0 or rng.rand() < 0.5 else ann[b, rows[j - 1], :4].copy()
            elif kind == 7:   # huge box
                box = np.array([-5., -7., width + 3., height + 9.])
            elif kind == 8:   # sub-pixel / degenerate
                x, y = rng.uniform(0, width), rng.uniform(0, height)
                box = np.array([x, y, x + rng.choice([0., 0.3, 1.]), y + rng.choice([0., 0.5, 2.])])
            else:             # random box
                x1, y1 = rng.uniform(-10, width), rng.uniform(-10, height)
                box = np.array([x1, y1, x1 + rng.uniform(1, width), y1 + rng.uniform(1, height)])
            ann[b, r, :4] = box.astype(np.float32)
            ann[b, r, 4] = rng.randint(0, C)
    ann_t = torch.from_numpy(ann).cuda()
    preds = dev([cls, reg])
    exact = crit.debug_assign(preds, ann_t, exact=True)['labels']
    fast = crit.debug_assign(preds, ann_t, exact=False)['labels']
    assert torch.equal(exact, fast), f'{int((exact != fast).sum())} labels differ'
    # ... and the queues the sparse kernel consumes agree with the labels: same losses either way
    with torch.no_grad():
        d = crit(preds, ann_t)
    assert np.isfinite(d['cls_loss'].item()) and np.isfinite(d['reg_loss'].item())
