"""Query-based decoders (DETRDecoder / DINODETRDecoder) and the stand-alone DecodeMethod /
DetNMSMethod (reference simpleAICV/detection/decode.py:24-172, :367-594).

CPU part: the oracle restatement reproduces the UNMODIFIED reference's outputs stored in
tests/golden/queries.npz bit for bit.  GPU part (-m gpu): the CUDA path (b200det_query_scores +
b200det_select_decode_nms through the drop-in classes) against the golden vectors and against the
oracle fed with torch's CUDA softmax / sigmoid -- what a CUDA run of the reference computes."""
import numpy as np
import pytest
import torch

from oracle import det_oracle as O

import golden_util as G

DETR_CASES = [('none', dict(nms_type=None)),
              ('nms', dict(nms_type='python_nms', topn=80, max_object_num=20)),
              ('low', dict(nms_type='diou_python_nms', min_score_threshold=0.3))]
DINO_CASES = [('python_nms', dict(nms_type='python_nms')),
              ('diou_python_nms', dict(nms_type='diou_python_nms')),
              ('torch_nms', dict(nms_type='torch_nms')),
              ('hi', dict(nms_type='python_nms', min_score_threshold=0.5, nms_threshold=0.3,
                          topn=50, max_object_num=40))]
NMS_TYPES = ['python_nms', 'diou_python_nms', 'torch_nms']


@pytest.fixture(scope='module')
def q():
    return G.load('queries.npz')


def golden_triplet(q, prefix):
    return [q[f'{prefix}_scores'], q[f'{prefix}_classes'], q[f'{prefix}_boxes']]


def assert_triplet(got, want, what, score_ulp=0):
    if score_ulp == 0:
        G.assert_bit_equal(got[0], want[0], f'{what}: scores')
    else:
        d = np.abs(G.bits(got[0]).astype(np.int64) - G.bits(want[0]).astype(np.int64))
        assert d.max() <= score_ulp, f'{what}: scores differ by {d.max()} ulp'
    G.assert_bit_equal(got[1], want[1], f'{what}: classes')
    G.assert_bit_equal(got[2], want[2], f'{what}: boxes')


# ---------------------------------------------------------------------------------------
# CPU: oracle pinned to the reference
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag,kw', DETR_CASES)
def test_oracle_detr_decoder(q, tag, kw):
    cls, reg = torch.from_numpy(q['detr_cls'].copy()), torch.from_numpy(q['detr_reg'].copy())
    got, _ = O.query_decode(cls[-1], reg[-1], q['sizes'].tolist(), 'softmax',
                            num_classes=int(q['num_classes']), **kw)
    assert_triplet(got, golden_triplet(q, f'detr_{tag}'), f'DETRDecoder {tag}')


@pytest.mark.parametrize('tag,kw', DINO_CASES)
def test_oracle_dino_decoder(q, tag, kw):
    cls, reg = torch.from_numpy(q['dino_cls'].copy()), torch.from_numpy(q['dino_reg'].copy())
    okw = dict(kw)
    okw.setdefault('topn', 300)                    # DINODETRDecoder's default (decode.py:490)
    got, _ = O.query_decode(cls, reg, q['sizes'].tolist(), 'sigmoid', **okw)
    assert_triplet(got, golden_triplet(q, f'dino_{tag}'), f'DINODETRDecoder {tag}')


@pytest.mark.parametrize('nms', NMS_TYPES)
def test_oracle_decode_method_and_nms(q, nms):
    got, _ = O.select_and_nms(q['dm_scores'], q['dm_classes'], q['dm_boxes'], 100, 0.6, 1000, nms,
                              0.5)
    assert_triplet(got, golden_triplet(q, f'dm_{nms}'), f'DecodeMethod {nms}')
    order = np.argsort(-q['dm_scores'][0], kind='stable')[:700]
    keep = O.nms_keep(q['dm_boxes'][0][order], q['dm_scores'][0][order], nms, 0.4)
    assert np.array_equal(keep, q[f'nms_{nms}_keep'])


# ---------------------------------------------------------------------------------------
# GPU: CUDA path vs golden vectors and vs the oracle on torch-CUDA activations
# ---------------------------------------------------------------------------------------
def cuda_softmax(x):
    return torch.nn.functional.softmax(x.cuda(), dim=2).cpu()


def cuda_sigmoid(x):
    return torch.sigmoid(x.cuda().float()).cpu()


@pytest.mark.gpu
@pytest.mark.parametrize('tag,kw', DETR_CASES)
def test_gpu_detr_decoder(q, tag, kw):
    from b200det import decode
    cls, reg = torch.from_numpy(q['detr_cls'].copy()), torch.from_numpy(q['detr_reg'].copy())
    sizes = q['sizes'].tolist()
    C = int(q['num_classes'])
    dec = decode.DETRDecoder(num_classes=C, **kw)
    got = dec([cls.cuda(), reg.cuda()], sizes)
    # bit-exact against the reference's arithmetic on the GPU (torch's CUDA softmax)
    want, _ = O.query_decode(cls[-1], reg[-1], sizes, 'softmax', num_classes=C,
                             prob_fn=cuda_softmax, **kw)
    assert_triplet(got, want, f'DETRDecoder {tag} vs oracle(cuda softmax)')
    # and within 8 ulp of the scores the reference produced on the CPU (torch's CPU softmax uses a
    # vectorised exp and sums in a different order: 4 ulp measured); classes and boxes identical
    assert_triplet(got, golden_triplet(q, f'detr_{tag}'), f'DETRDecoder {tag} vs golden',
                   score_ulp=8)


@pytest.mark.gpu
@pytest.mark.parametrize('tag,kw', DINO_CASES)
def test_gpu_dino_decoder(q, tag, kw):
    from b200det import decode
    cls, reg = torch.from_numpy(q['dino_cls'].copy()), torch.from_numpy(q['dino_reg'].copy())
    sizes = q['sizes'].tolist()
    dec = decode.DINODETRDecoder(**kw)
    got = dec({'pred_logits': cls.cuda(), 'pred_boxes': reg.cuda()}, sizes)
    okw = dict(kw)
    okw.setdefault('topn', 300)
    want, _ = O.query_decode(cls, reg, sizes, 'sigmoid', prob_fn=cuda_sigmoid, **okw)
    assert_triplet(got, want, f'DINODETRDecoder {tag} vs oracle(cuda sigmoid)')
    assert_triplet(got, golden_triplet(q, f'dino_{tag}'), f'DINODETRDecoder {tag} vs golden',
                   score_ulp=2)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_gpu_dino_decoder_half_logits(q, dtype):
    """`cls_preds.float()` first (decode.py:515): half-precision logits are upcast on load."""
    from b200det import decode
    cls = torch.from_numpy(q['dino_cls'].copy()).to(dtype)
    reg = torch.from_numpy(q['dino_reg'].copy())
    sizes = q['sizes'].tolist()
    got = decode.DINODETRDecoder()({'pred_logits': cls.cuda(), 'pred_boxes': reg.cuda()}, sizes)
    want, _ = O.query_decode(cls, reg, sizes, 'sigmoid', prob_fn=cuda_sigmoid, topn=300,
                             nms_type='python_nms')
    assert_triplet(got, want, f'DINODETRDecoder {dtype}')


@pytest.mark.gpu
@pytest.mark.parametrize('seed', range(6))
def test_gpu_query_decoders_fuzz(seed):
    """Random shapes: channel counts on both sides of the warp width (the softmax kernel's lane
    layout changes at 32), query counts, thresholds, NMS types."""
    from b200det import decode
    rng = np.random.RandomState(500 + seed)
    B, Q = int(rng.randint(1, 5)), int(rng.choice([1, 7, 100, 300, 900]))
    C = int(rng.choice([1, 3, 20, 31, 32, 33, 80, 91, 250]))
    gen = torch.Generator().manual_seed(seed)
    cls = torch.randn((1, B, Q, C + 1), generator=gen) * float(rng.choice([1.0, 3.0]))
    reg = torch.cat([torch.rand((1, B, Q, 2), generator=gen),
                     torch.rand((1, B, Q, 2), generator=gen) * 0.5], dim=-1)
    sizes = [[int(rng.randint(200, 900)), int(rng.randint(200, 900))] for _ in range(B)]
    kw = dict(min_score_threshold=float(rng.choice([0.01, 0.05, 0.2])),
              topn=int(rng.choice([10, 100, 300])), max_object_num=int(rng.choice([5, 100])),
              nms_type=[None, 'python_nms', 'diou_python_nms', 'torch_nms'][seed % 4],
              nms_threshold=float(rng.choice([0.3, 0.5])))
    got = decode.DETRDecoder(num_classes=C, **kw)([cls.cuda(), reg.cuda()], sizes)
    want, _ = O.query_decode(cls[-1], reg[-1], sizes, 'softmax', num_classes=C,
                             prob_fn=cuda_softmax, **kw)
    assert_triplet(got, want, f'DETR fuzz {seed} C={C} Q={Q}')
    if kw['nms_type'] is not None:
        got = decode.DINODETRDecoder(**kw)({'pred_logits': cls[-1].cuda(),
                                            'pred_boxes': reg[-1].cuda()}, sizes)
        want, _ = O.query_decode(cls[-1], reg[-1], sizes, 'sigmoid', prob_fn=cuda_sigmoid, **kw)
        assert_triplet(got, want, f'DINO fuzz {seed} C={C} Q={Q}')


@pytest.mark.gpu
@pytest.mark.parametrize('nms', NMS_TYPES)
def test_gpu_decode_method_and_nms(q, nms):
    """NumPy in, NumPy out like the reference's classes (inputs are uploaded); CUDA tensors work
    too.  5000 candidates per image above a 0.6 threshold -> top-1000 -> NMS -> 100."""
    from b200det import decode
    dm = decode.DecodeMethod(max_object_num=100, min_score_threshold=0.6, topn=1000, nms_type=nms,
                             nms_threshold=0.5)
    got = dm(q['dm_scores'], q['dm_classes'], q['dm_boxes'])
    assert_triplet(got, golden_triplet(q, f'dm_{nms}'), f'DecodeMethod {nms}')
    got = dm(torch.from_numpy(q['dm_scores'].copy()).cuda(),
             torch.from_numpy(q['dm_classes'].copy()).cuda(),
             torch.from_numpy(q['dm_boxes'].copy()).cuda())
    assert_triplet(got, golden_triplet(q, f'dm_{nms}'), f'DecodeMethod {nms} (CUDA tensors)')
    order = np.argsort(-q['dm_scores'][0], kind='stable')[:700]
    keep = decode.DetNMSMethod(nms_type=nms, nms_threshold=0.4)(q['dm_boxes'][0][order],
                                                                  q['dm_scores'][0][order])
    want = q[f'nms_{nms}_keep']
    assert keep.dtype == want.dtype and np.array_equal(keep, want)
    assert decode.DetNMSMethod(nms_type=nms)(np.zeros((0, 4), np.float32),
                                             np.zeros((0,), np.float32)).shape == (0,)


@pytest.mark.gpu
def test_gpu_query_decoder_edges():
    """No candidate at all, every query on the no-object channel, a single query, CPU tensors."""
    from b200det import decode
    cls = torch.full((1, 2, 5, 4), -3.0)
    cls[..., 3] = 9.0                                     # no-object wins everywhere
    reg = torch.rand((1, 2, 5, 4))
    s, c, b = decode.DETRDecoder(num_classes=3)([cls.cuda(), reg.cuda()], [[10, 10], [20, 20]])
    assert (s == -1).all() and (c == -1).all() and (b == 0).all()
    s, c, b = decode.DINODETRDecoder(min_score_threshold=0.999)(
        {'pred_logits': cls[0].cuda(), 'pred_boxes': reg[0].cuda()}, [[10, 10], [20, 20]])
    assert (s[:, :5] > 0.999).all() and (s[:, 5:] == -1).all() and (c[:, :5] == 3).all()
    one = torch.tensor([[[[2.0, 0.5, -1.0]]]])            # [1, 1, 1, 3]
    s, c, b = decode.DETRDecoder(num_classes=2)([one.cuda(), torch.tensor([[[[.5, .5, .2, .4]]]]).cuda()],
                                                [[100, 200]])
    assert c[0, 0] == 0 and s[0, 0] == torch.softmax(one.cuda(), -1)[0, 0, 0, 0].item()
    f = np.float32
    want = np.array([(f(.5) - f(.5) * f(.2)) * f(200), (f(.5) - f(.5) * f(.4)) * f(100),
                     (f(.5) + f(.5) * f(.2)) * f(200), (f(.5) + f(.5) * f(.4)) * f(100)], dtype=f)
    G.assert_bit_equal(b[0, 0], want, 'single query box')
    with pytest.raises(RuntimeError):
        decode.DETRDecoder()([cls, reg], [[10, 10], [20, 20]])


@pytest.mark.gpu
@pytest.mark.parametrize('nms', ['python_nms', 'torch_nms'])
def test_gpu_nms_threshold_band(nms):
    """The select kernel decides most NMS pairs without the IEEE division (a pair is only sent to the
    exact function when inter / union is within 5e-7 of the threshold).  Thresholds placed exactly
    on, one ulp below and one ulp above the float32 IoU of real pairs -- where `iou < thr` flips --
    must give the reference's keep list."""
    from b200det import decode
    rng = np.random.RandomState(7)
    n = 300
    xy = rng.randint(0, 60, size=(n, 2)).astype(np.float32)
    wh = rng.randint(4, 40, size=(n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], axis=1)
    boxes[1] = [0, 0, 2, 2]
    boxes[2] = [0, 0, 2, 1]                       # IoU with box 1 is exactly 0.5
    scores = np.linspace(1.0, 0.1, n).astype(np.float32)
    # float32 IoUs of some overlapping pairs, computed like decode.py:45-76
    thrs = [0.5]
    for i, j in [(0, 5), (3, 9), (10, 11), (20, 40), (7, 8), (1, 2)]:
        a, b = boxes[i], boxes[j]
        iw = max(np.float32(min(a[2], b[2]) - max(a[0], b[0])), np.float32(0))
        ih = max(np.float32(min(a[3], b[3]) - max(a[1], b[1])), np.float32(0))
        inter = np.float32(iw * ih)
        area = lambda t: np.float32((t[2] - t[0]) * (t[3] - t[1]))   # noqa: E731
        union = np.float32(max(np.float32(np.float32(area(a) + area(b)) - inter), np.float32(1e-4)))
        iou = np.float32(inter / union)
        if 0 < iou < 1:
            thrs += [float(iou), float(np.nextafter(iou, np.float32(0))),
                     float(np.nextafter(iou, np.float32(1)))]
    for thr in thrs:
        want = O.nms_keep(boxes, scores, nms, thr)
        got = decode.DetNMSMethod(nms_type=nms, nms_threshold=thr)(boxes, scores)
        assert np.array_equal(got, want), f'{nms} thr={thr!r}'


@pytest.mark.gpu
@pytest.mark.parametrize('nms', NMS_TYPES)
@pytest.mark.parametrize('seed', range(4))
def test_gpu_nms_clustered_boxes(nms, seed):
    """Heavily overlapping candidates (clusters of jittered boxes, integer and fractional
    coordinates): most candidates are suppressed, groups of the select kernel's 8-per-step greedy
    scan contain suppressed members, windows slide, and the max_object_num cut falls inside a
    group.  Full keep lists and capped outputs must equal the reference's scan."""
    from b200det import decode
    rng = np.random.RandomState(900 + seed)
    n = int(rng.choice([37, 300, 1000, 2048]))
    centers = rng.uniform(50, 400, size=(int(rng.choice([1, 3, 12])), 2))
    c = centers[rng.randint(0, len(centers), size=n)]
    wh = rng.uniform(20, 90, size=(n, 2))
    xy = c + rng.normal(0, float(rng.choice([2.0, 10.0, 30.0])), size=(n, 2)) - wh / 2
    boxes = np.concatenate([xy, xy + wh], axis=1).astype(np.float32)
    if seed % 2:
        boxes = np.trunc(boxes)                   # the dense decoders' integer-valued boxes
    scores = np.sort(rng.uniform(0.06, 1.0, size=n).astype(np.float32))[::-1].copy()
    scores = np.unique(scores)[::-1].copy()
    boxes = boxes[:len(scores)]
    n = len(scores)
    thr = float(rng.choice([0.3, 0.5, 0.7]))
    want = O.nms_keep(boxes, scores, nms, thr)
    got = decode.DetNMSMethod(nms_type=nms, nms_threshold=thr)(boxes, scores)
    assert np.array_equal(got, want), f'{nms} n={n} thr={thr}: keep list'
    classes = rng.randint(0, 5, size=n)
    for m in (1, 7, 8, 9, 100):
        dm = decode.DecodeMethod(max_object_num=m, min_score_threshold=0.05, topn=min(n, 2048),
                                 nms_type=nms, nms_threshold=thr)
        got3 = dm(scores[None], classes[None], boxes[None])
        want3, _ = O.select_and_nms(scores[None], classes[None], boxes[None], m, 0.05,
                                    min(n, 2048), nms, thr)
        assert_triplet(got3, want3, f'{nms} n={n} thr={thr} max_object_num={m}')
