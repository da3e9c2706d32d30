"""GPU parity of the logits path:
the classification branch computed straight from the head's NCHW logits must give
  * detections bit-identical to the probability path (itself bit-exact against the oracle, see
    test_gpu_parity.py) fed with the reference's own ops `sigmoid(x.float()).permute(0,2,3,1)` run by
    torch on the same GPU, and
  * loss values within 1e-5 relative of it (tolerance of north_star; the focal term is evaluated
    from t = e^x with a polynomial instead of from the rounded probability).
"""
import numpy as np
import pytest
import torch

from b200det import synth, losses, decode, fused
from oracle import det_oracle as O

import golden_util as G
from test_gpu_parity import LOSS_RTOL, assert_close, loss_values

pytestmark = pytest.mark.gpu


def make_logits(batch, sizes, per_loc, num_classes, seed, mean=-4.595, sigma=1.0,
                dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    cls, reg = [], []
    for h, w in sizes:
        x = torch.randn((batch, per_loc * num_classes, h, w), generator=gen) * sigma + mean
        cls.append(x.to(dtype).cuda())
        reg.append((torch.randn((batch, h, w, per_loc, 4), generator=gen) * 0.2).cuda())
    return cls, reg


def both_paths(cls_logits, reg, ann, kw, num_classes, loss_kw=None, dec_kw=None):
    crit = losses.RetinaLoss(**kw, **(loss_kw or {}))
    dec = decode.RetinaDecoder(**kw, **(dec_kw or {}))
    probs = [O.head_tail(x, num_classes) for x in cls_logits]      # the reference's ops, on CUDA
    with torch.no_grad():
        want_loss = crit([probs, reg], ann)
        want_det = dec([probs, reg])
        got_loss, got_det = fused.LogitsEvalStep(crit, dec)([cls_logits, reg], ann)
    oracle_leg([p.cpu() for p in probs], [r.cpu() for r in reg], ann.cpu(), kw, loss_kw, dec_kw,
               got_loss, got_det)
    return want_loss, want_det, got_loss, got_det


def oracle_leg(probs, reg, ann, kw, loss_kw, dec_kw, got_loss, got_det):
    """Third leg: the ORACLE (pinned to the unmodified reference) on the same probabilities -- the
    logits path is compared with the reference's arithmetic directly, not only through this
    repo's probability path.  Scores can tie on these inputs, where the reference's unstable
    argsort leaves the order undefined: detections are compared when the oracle's candidate
    scores are unique, the loss always."""
    with torch.no_grad():
        ref = O.retina_loss([probs, reg], ann, **kw, **(loss_kw or {}))
    w = np.array([ref['cls_loss'].item(), ref['reg_loss'].item()])
    g = loss_values(got_loss, ['cls_loss', 'reg_loss'])
    if not np.isfinite(w).all():
        assert (np.isfinite(g) == np.isfinite(w)).all()
    elif (w == 0).all():
        assert (g == 0).all()
    else:
        assert_close(g, w, LOSS_RTOL, 'logits path vs oracle: loss')
    (s0, c0, b0), extra = O.retina_decode([probs, reg], **kw, **(dec_kw or {}))
    thr = (dec_kw or {}).get('min_score_threshold', 0.05)
    for b in range(s0.shape[0]):
        sc = extra['scores'][b]
        cand = sc[sc > np.float32(thr)]
        if np.unique(cand).size != cand.size:
            continue   # ties among the candidates: order undefined in the reference
        G.assert_bit_equal(got_det[0][b], s0[b], f'logits path vs oracle: scores, image {b}')
        G.assert_bit_equal(got_det[1][b], c0[b], f'logits path vs oracle: classes, image {b}')
        G.assert_bit_equal(got_det[2][b], b0[b], f'logits path vs oracle: boxes, image {b}')


def check(want_loss, want_det, got_loss, got_det, what=''):
    for name, a, b in zip(('scores', 'classes', 'boxes'), got_det, want_det):
        G.assert_bit_equal(a, b, f'{what} {name}')
    w = loss_values(want_loss, ['cls_loss', 'reg_loss'])
    g = loss_values(got_loss, ['cls_loss', 'reg_loss'])
    if (w == 0).all():
        assert (g == 0).all()
    else:
        assert_close(g, w, LOSS_RTOL, f'{what} loss')


@pytest.mark.parametrize('box_type', ['SmoothL1', 'GIoU'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
def test_logits_eval_step_matches_probability_path(box_type, dtype):
    C = 8
    sizes = [(p, p) for p in synth.pyramid_sizes(256)]
    cls, reg = make_logits(3, sizes, 9, C, seed=1, dtype=dtype)
    ann = synth.make_annotations(3, 20, 256, C, seed=2, empty_images=(1,)).cuda()
    check(*both_paths(cls, reg, ann, synth.RETINA_KW, C, dict(box_loss_type=box_type)),
          what=f'{box_type} {dtype}')


def test_logits_ties_saturation_and_threshold_edges():
    """Rows whose best classes tie after rounding (duplicated logits, logits one ulp apart,
    saturation at p == 1.0) must pick np.argmax's FIRST maximum of the probabilities; scores at the
    threshold follow the strict float32 comparison."""
    C = 12
    sizes = [(9, 7), (5, 4)]
    kw = dict(areas=[[32, 32], [64, 64]], ratios=[0.5, 1, 2], scales=[1, 1.5], strides=[8, 16])
    cls, reg = make_logits(2, sizes, 6, C, seed=3, mean=-2.0)
    x = cls[0].view(2, 6, C, 9 * 7)
    x[0, 0, 3, 0] = x[0, 0, 7, 0] = 5.0                      # exact duplicate maxima -> class 3
    x[0, 1, 9, 1] = 20.0
    x[0, 1, 2, 1] = 30.0                                      # both saturate to 1.0 -> class 2
    x[0, 2, 4, 2] = 3.0
    x[0, 2, 5, 2] = float(np.nextafter(np.float32(3.0), np.float32(4)))   # 1 ulp apart
    x[1, 3, :, 5] = -1.0                                      # a whole row of equal logits -> class 0
    thr = 0.05
    logit_thr = float(np.log(thr / (1 - thr)))
    for i, d in enumerate((-2e-6, -1e-6, 0.0, 1e-6, 2e-6)):   # probabilities straddling 0.05
        x[1, 4, :, 10 + i] = -20.0
        x[1, 4, 6, 10 + i] = logit_thr + d
    x[1, 5, :, 20] = float('-inf')                            # p == 0 everywhere
    x[1, 5, 1, 21] = float('inf')                             # p == 1
    ann = synth.make_annotations(2, 8, 72, C, seed=4).cuda()
    for nms in ('python_nms', 'torch_nms'):
        res = both_paths(cls, reg, ann, kw, C, None, dict(nms_type=nms, topn=500,
                                                            max_object_num=300))
        check(*res, what=nms)
    # the planted rows end up with the classes of the first maxima
    crit = losses.RetinaLoss(**kw)
    dec = decode.RetinaDecoder(**kw, topn=2000, max_object_num=600, nms_threshold=1.1)
    _, (s, c, _) = fused.LogitsEvalStep(crit, dec)([cls, reg], ann)
    probs = torch.sigmoid(cls[0].float()).view(2, 6, C, 63).permute(0, 3, 1, 2)   # [B, hw, A, C]
    want_cls = probs.argmax(dim=-1).cpu().numpy()
    assert want_cls[0, 0, 0] == 3 and want_cls[0, 1, 1] == 2 and want_cls[1, 5, 3] == 0
    assert (s[0] == 1.0).sum() >= 1 and set(c[0][s[0] == 1.0].tolist()) <= {2.0}


@pytest.mark.parametrize('seed', range(10))
def test_logits_fuzz(seed):
    rng = np.random.RandomState(3000 + seed)
    n_levels = int(rng.randint(1, 5))
    h, w = int(rng.randint(3, 33)), int(rng.randint(3, 33))
    s0 = float(rng.choice([4, 8]))
    sizes, strides = [], []
    for l in range(n_levels):
        sizes.append((h, w))
        strides.append(s0 * 2**l)
        h, w = (h + 1) // 2, (w + 1) // 2
    ratios = [float(r) for r in rng.choice([0.5, 1, 2, 3], size=int(rng.randint(1, 4)), replace=False)]
    scales = [float(s) for s in rng.choice([1, 1.26, 1.6], size=int(rng.randint(1, 4)), replace=False)]
    kw = dict(areas=[[4 * s, 4 * s] for s in strides], ratios=ratios, scales=scales, strides=strides)
    A = len(ratios) * len(scales)
    C = int(rng.choice([1, 3, 4, 7, 20, 80, 91]))
    B = int(rng.randint(1, 5))
    dtype = [torch.float32, torch.float16, torch.bfloat16][seed % 3]
    cls, reg = make_logits(B, sizes, A, C, seed=seed, mean=float(rng.choice([-4.595, -2.0, 0.0])),
                           sigma=float(rng.choice([1.0, 3.0])), dtype=dtype)
    width = sizes[0][1] * strides[0]
    ann = synth.make_annotations(B, int(rng.choice([1, 5, 40])), max(int(width), 17), C,
                                 seed=seed + 7).cuda()
    loss_kw = dict(box_loss_type=str(rng.choice(['SmoothL1', 'CIoU'])),
                   alpha=float(rng.choice([0.25, 0.4])), gamma=float(rng.choice([2.0, 1.5])))
    dec_kw = dict(min_score_threshold=float(rng.choice([0.01, 0.05, 0.3])),
                  topn=int(rng.choice([50, 1000])), max_object_num=int(rng.choice([10, 100])),
                  nms_type=str(rng.choice(['python_nms', 'diou_python_nms', 'torch_nms'])))
    check(*both_paths(cls, reg, ann, kw, C, loss_kw, dec_kw), what=f'seed {seed} C={C} A={A} {dtype}')


def test_logits_full_size_and_nan():
    C, B = 80, 4
    sizes = [(p, p) for p in synth.pyramid_sizes(800)]
    cls, reg = make_logits(B, sizes, 9, C, seed=9)
    ann = synth.make_annotations(B, 100, 800, C, seed=10, empty_images=(2,)).cuda()
    check(*both_paths(cls, reg, ann, synth.RETINA_KW, C, dict(box_loss_type='GIoU')), what='800x800')
    cls[1][0, 5, 3, 3] = float('nan')
    crit = losses.RetinaLoss(**synth.RETINA_KW)
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    loss, _ = fused.LogitsEvalStep(crit, dec)([cls, reg], ann)
    assert np.isnan(loss['cls_loss'].item()) and np.isfinite(loss['reg_loss'].item())


def fcos_logits(batch, sizes, num_classes, seed, dtype=torch.float32, mean=-4.595):
    gen = torch.Generator().manual_seed(seed)
    cls, reg, ctr = [], [], []
    for l, (h, w) in enumerate(sizes):
        cls.append((torch.randn((batch, num_classes, h, w), generator=gen) + mean).to(dtype).cuda())
        reg.append((torch.randn((batch, h, w, 4), generator=gen) * 0.5
                    + float(np.log(8.0 * 2**l))).cuda())
        ctr.append(torch.randn((batch, 1, h, w), generator=gen).cuda())
    return cls, reg, ctr


def fcos_both_paths(cls, reg, ctr, ann, strides, mi, loss_kw=None, dec_kw=None):
    crit = losses.FCOSLoss(strides=strides, mi=mi, **(loss_kw or {}))
    dec = decode.FCOSDecoder(strides=strides, **(dec_kw or {}))
    # the reference's ops on CUDA: sigmoid(x.float()) then permute(0, 2, 3, 1)  (models/fcos.py:70-79)
    probs = [O.head_tail(x) for x in cls]
    cprobs = [O.head_tail(x) for x in ctr]
    with torch.no_grad():
        want_loss = crit([probs, reg, cprobs], ann)
        want_det = dec([probs, reg, cprobs])
        got_loss, got_det = fused.LogitsEvalStep(crit, dec)([cls, reg, ctr], ann)
    # third leg: the oracle on the same probabilities (see oracle_leg)
    cpu = [[t.cpu() for t in grp] for grp in (probs, reg, cprobs)]
    with torch.no_grad():
        ref = O.fcos_loss(cpu, ann.cpu(), strides, mi, **(loss_kw or {}))
    keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    w = np.array([ref[k].item() for k in keys])
    g = loss_values(got_loss, keys)
    if (w == 0).all():
        assert (g == 0).all()
    else:
        assert_close(g, w, LOSS_RTOL, 'FCOS logits path vs oracle: loss')
    (s0, c0, b0), extra = O.fcos_decode(cpu, strides, **(dec_kw or {}))
    thr = (dec_kw or {}).get('min_score_threshold', 0.05)
    for b in range(s0.shape[0]):
        sc = extra['scores'][b]
        cand = sc[sc > np.float32(thr)]
        if np.unique(cand).size != cand.size:
            continue
        # torch-CPU sqrt (MKL) may be 1 ulp off the IEEE value the GPU computes (see
        # test_gpu_parity.assert_targets_equal); scores come from NumPy's sqrt here, which is IEEE
        G.assert_bit_equal(got_det[0][b], s0[b], f'FCOS logits path vs oracle: scores, image {b}')
        G.assert_bit_equal(got_det[1][b], c0[b], f'FCOS logits path vs oracle: classes, image {b}')
        G.assert_bit_equal(got_det[2][b], b0[b], f'FCOS logits path vs oracle: boxes, image {b}')
    return want_loss, want_det, got_loss, got_det


def check_fcos(want_loss, want_det, got_loss, got_det, what=''):
    for name, a, b in zip(('scores', 'classes', 'boxes'), got_det, want_det):
        G.assert_bit_equal(a, b, f'{what} {name}')
    keys = ['cls_loss', 'reg_loss', 'center_ness_loss']
    w, g = loss_values(want_loss, keys), loss_values(got_loss, keys)
    if (w == 0).all():
        assert (g == 0).all()
    else:
        assert_close(g, w, LOSS_RTOL, f'{what} loss')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('iou_type', ['GIoU', 'CIoU'])
def test_fcos_logits_eval_step_matches_probability_path(iou_type, dtype):
    C = 20
    sizes = [(p, p) for p in synth.pyramid_sizes(256)]
    cls, reg, ctr = fcos_logits(3, sizes, C, seed=21, dtype=dtype)
    ann = synth.make_annotations(3, 20, 256, C, seed=22, empty_images=(1,)).cuda()
    check_fcos(*fcos_both_paths(cls, reg, ctr, ann, synth.STRIDES, synth.MI,
                                dict(box_loss_iou_type=iou_type)), what=f'FCOS {iou_type} {dtype}')


@pytest.mark.parametrize('seed', range(6))
def test_fcos_logits_fuzz(seed):
    rng = np.random.RandomState(4000 + seed)
    n_levels = int(rng.randint(1, 6))
    h, w = int(rng.randint(3, 40)), int(rng.randint(3, 40))
    sizes, strides, mi = [], [], []
    s0, lo = 8.0, -1.0
    for l in range(n_levels):
        sizes.append((h, w))
        strides.append(s0 * 2**l)
        hi = 64.0 * 2**l if l < n_levels - 1 else 100000000.0
        mi.append([lo, hi])
        lo = hi
        h, w = (h + 1) // 2, (w + 1) // 2
    C = int(rng.choice([1, 3, 7, 16, 80, 365]))
    B = int(rng.randint(1, 4))
    dtype = [torch.float32, torch.float16, torch.bfloat16][seed % 3]
    cls, reg, ctr = fcos_logits(B, sizes, C, seed=seed, dtype=dtype,
                                mean=float(rng.choice([-4.595, -2.0])))
    ann = synth.make_annotations(B, int(rng.choice([1, 5, 40])),
                                 max(int(sizes[0][1] * strides[0]), 17), C, seed=seed + 3).cuda()
    loss_kw = dict(box_loss_iou_type=str(rng.choice(['IoU', 'GIoU', 'DIoU', 'EIoU'])),
                   use_center_sample=bool(rng.randint(0, 2)),
                   alpha=float(rng.choice([0.25, 0.4])), gamma=float(rng.choice([2.0, 1.5])))
    dec_kw = dict(min_score_threshold=float(rng.choice([0.01, 0.05, 0.3])),
                  topn=int(rng.choice([50, 1000])), max_object_num=int(rng.choice([10, 100])),
                  nms_type=str(rng.choice(['python_nms', 'diou_python_nms', 'torch_nms'])))
    check_fcos(*fcos_both_paths(cls, reg, ctr, ann, strides, mi, loss_kw, dec_kw),
               what=f'FCOS seed {seed} C={C} {dtype}')


def test_fcos_logits_full_size():
    """BASELINE configs[2] shape (FCOS 800x800, 80 classes) at a reduced batch."""
    C, B = 80, 4
    sizes = [(p, p) for p in synth.pyramid_sizes(800)]
    cls, reg, ctr = fcos_logits(B, sizes, C, seed=31)
    ann = synth.make_annotations(B, 100, 800, C, seed=32).cuda()
    check_fcos(*fcos_both_paths(cls, reg, ctr, ann, synth.STRIDES, synth.MI), what='FCOS 800x800')


def test_logits_step_rejects_unsupported_inputs():
    crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    with pytest.raises(ValueError):
        fused.LogitsEvalStep(crit, decode.RetinaDecoder(**synth.RETINA_KW))
    fstep = fused.LogitsEvalStep(crit, decode.FCOSDecoder(strides=synth.STRIDES))
    with pytest.raises(RuntimeError):   # FCOS heads need their centre-ness logits
        fstep([[torch.zeros(1, 80, 4, 4).cuda()], [torch.zeros(1, 4, 4, 4).cuda()]],
              torch.zeros(1, 1, 5).cuda())
    rc = losses.RetinaLoss(**synth.RETINA_KW)
    step = fused.LogitsEvalStep(rc, decode.RetinaDecoder(**synth.RETINA_KW))
    with pytest.raises(RuntimeError):
        step([[torch.zeros(1, 72, 4, 4)], [torch.zeros(1, 4, 4, 9, 4)]], torch.zeros(1, 1, 5))


def test_logits_nan_rows_are_dropped_like_np_argmax():
    """A NaN logit makes sigmoid NaN; np.argmax returns the first NaN (decode.py:230-238), the row's
    score is NaN and fails `score > threshold` -- even when another class of the row is confident."""
    C = 8
    sizes = [(p, p) for p in synth.pyramid_sizes(128)]
    cls, reg = make_logits(2, sizes, 9, C, seed=41, mean=-3.0)
    x = cls[0].view(2, 9, C, -1)                    # [B, A, C, HW]
    best = x[0].max(dim=1).values.flatten().argsort(descending=True)
    hw = x.shape[-1]
    for n, (cpos, val) in enumerate(((0, 6.0), (C - 1, 7.0), (3, 8.0))):
        a_i, p_i = int(best[n]) // hw, int(best[n]) % hw
        x[0, a_i, (cpos + 1) % C, p_i] = val        # a confident class ...
        x[0, a_i, cpos, p_i] = float('nan')         # ... and a NaN beside it
    ann = synth.make_annotations(2, 12, 128, C, seed=42).cuda()
    crit = losses.RetinaLoss(**synth.RETINA_KW)
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    probs = [O.head_tail(t, C) for t in cls]
    _, got = fused.LogitsEvalStep(crit, dec)([cls, reg], ann)
    want = dec([probs, reg])
    (s0, c0, b0), _ = O.retina_decode([[p.cpu() for p in probs], [r.cpu() for r in reg]],
                                      **synth.RETINA_KW)
    for name, a, b, c in zip(('scores', 'classes', 'boxes'), got, want, (s0, c0, b0)):
        G.assert_bit_equal(a, b, f'NaN rows, probability path: {name}')
        G.assert_bit_equal(a[1], c[1], f'NaN rows, oracle (image without ties): {name}')
    assert not np.isnan(got[0]).any() and (got[0][0] < 0.99).all()   # the planted rows are gone
