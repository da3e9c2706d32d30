"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
synthetic inputs.  Run in the development container only:

    python tests/golden/make_golden.py

Each fixture stores the exact inputs (so no RNG has to be reproduced elsewhere) and what the
reference's own classes returned for them:
  * RetinaLoss / FCOSLoss: loss dict values for every box-loss type, plus the labels the
    reference's assignment method produced (get_batch_anchors_annotations /
    get_batch_position_annotations) and FCOS regression/centre-ness targets;
  * RetinaDecoder / FCOSDecoder: the three returned arrays for python_nms, diou_python_nms and
    torch_nms, default and small (topn, max_object_num) settings;
  * RetinaAnchors / FCOSPositions tables; np.exp samples;
  * DETRDecoder / DINODETRDecoder / DecodeMethod / DetNMSMethod outputs (queries.npz).
numpy / torch / torchvision versions used are recorded in the fixture.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import refload  # noqa: E402
from b200det import synth  # noqa: E402

IOU_TYPES = ['IoU', 'GIoU', 'DIoU', 'CIoU', 'EIoU']


def versions():
    import torchvision
    return np.array([f'numpy {np.__version__}', f'torch {torch.__version__}',
                     f'torchvision {torchvision.__version__}'])


def edge_annotations(ann, size):
    """Adds the edge cases the domain has: an image without GT, duplicated GT boxes (IoU ties ->
    first maximum), a degenerate zero-area box, a box touching the image border, interleaved
    invalid rows (already produced by synth)."""
    ann = ann.clone()
    ann[1, :, :] = -1                       # image 1: no annotations at all
    valid = (ann[0, :, 4] >= 0).nonzero().flatten()
    free = (ann[0, :, 4] < 0).nonzero().flatten()
    if len(valid) and len(free) >= 2:
        ann[0, free[0]] = ann[0, valid[0]]      # exact duplicate, different class
        ann[0, free[0], 4] = (ann[0, valid[0], 4] + 1) % 3
        ann[0, free[1]] = torch.tensor([10., 10., 10., 30., 2.])  # zero width
    ann[2, 0] = torch.tensor([0., 0., float(size), float(size), 1.])  # whole image
    return ann


def make_retina(path):
    L, D, _ = refload.load()
    size, C, B, G = 128, 8, 3, 12
    preds = synth.make_retina_preds(B, size, C, seed=0)
    preds = synth.make_tie_free(preds)
    ann = edge_annotations(synth.make_annotations(B, G, size, C, seed=1), size)
    # plant confident, well-localised predictions so positives exist with sensible box losses
    out = {'versions': versions(), 'size': size, 'annotations': ann.numpy()}
    for i, (c, r) in enumerate(zip(*preds)):
        out[f'cls{i}'] = c.numpy()
        out[f'reg{i}'] = r.numpy()
    for box_type in ['SmoothL1'] + IOU_TYPES:
        crit = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
        with torch.no_grad():
            d = crit(preds, ann)
        out[f'loss_{box_type}'] = np.array([d['cls_loss'].item(), d['reg_loss'].item()],
                                           dtype=np.float32)
        if box_type in ('SmoothL1', 'GIoU'):
            anchors = crit.anchors([[c.shape[2], c.shape[1]] for c in preds[0]])
            flat = torch.cat([torch.tensor(a).view(-1, 4) for a in anchors], dim=0)
            t = crit.get_batch_anchors_annotations(flat.unsqueeze(0).repeat(B, 1, 1), ann)
            out[f'assign_{box_type}'] = t.numpy()
            out['anchors'] = flat.numpy()
    # gradients of the reference (autograd) for SmoothL1 and GIoU
    for box_type in ['SmoothL1', 'GIoU', 'CIoU']:
        crit = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
        d = crit(p, ann)
        (d['cls_loss'] + 2.0 * d['reg_loss']).backward()
        for i in range(len(p[0])):
            out[f'gcls_{box_type}_{i}'] = p[0][i].grad.numpy()
            out[f'greg_{box_type}_{i}'] = p[1][i].grad.numpy()
    for nms in ['python_nms', 'diou_python_nms', 'torch_nms']:
        for tag, kw in [('default', {}), ('small', dict(topn=300, max_object_num=20))]:
            dec = D.RetinaDecoder(**synth.RETINA_KW, nms_type=nms, **kw)
            s, c, b = dec(preds)
            out[f'dec_{nms}_{tag}_scores'] = s
            out[f'dec_{nms}_{tag}_classes'] = c
            out[f'dec_{nms}_{tag}_boxes'] = b
    np.savez_compressed(path, **out)


def make_fcos(path):
    L, D, _ = refload.load()
    size, C, B, G = 256, 8, 3, 12
    preds = synth.make_fcos_preds(B, size, C, seed=2)
    preds = synth.make_tie_free(preds)
    ann = edge_annotations(synth.make_annotations(B, G, size, C, seed=3), size)
    out = {'versions': versions(), 'size': size, 'annotations': ann.numpy()}
    for i, (c, r, t) in enumerate(zip(*preds)):
        out[f'cls{i}'] = c.numpy()
        out[f'reg{i}'] = r.numpy()
        out[f'ctr{i}'] = t.numpy()
    for iou_type in IOU_TYPES:
        crit = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou_type)
        with torch.no_grad():
            d = crit(preds, ann)
        out[f'loss_{iou_type}'] = np.array(
            [d['cls_loss'].item(), d['reg_loss'].item(), d['center_ness_loss'].item()],
            dtype=np.float32)
    for tag, kw in [('center', dict(use_center_sample=True)),
                    ('nocenter', dict(use_center_sample=False))]:
        crit = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, **kw)
        with torch.no_grad():
            d = crit(preds, ann)
            out[f'loss_{tag}'] = np.array(
                [d['cls_loss'].item(), d['reg_loss'].item(), d['center_ness_loss'].item()],
                dtype=np.float32)
            positions = crit.positions([[c.shape[2], c.shape[1]] for c in preds[0]])
            batch_positions = [torch.tensor(p).unsqueeze(0).repeat(B, 1, 1, 1) for p in positions]
            t = crit.get_batch_position_annotations(preds[0], preds[1], preds[2], batch_positions,
                                                    ann, use_center_sample=kw['use_center_sample'])
            out[f'targets_{tag}'] = t[3].numpy()
    for iou_type in ['GIoU', 'EIoU']:
        crit = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou_type)
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
        d = crit(p, ann)
        (d['cls_loss'] + 2.0 * d['reg_loss'] + 3.0 * d['center_ness_loss']).backward()
        for i in range(len(p[0])):
            out[f'gcls_{iou_type}_{i}'] = p[0][i].grad.numpy()
            out[f'greg_{iou_type}_{i}'] = p[1][i].grad.numpy()
            out[f'gctr_{iou_type}_{i}'] = p[2][i].grad.numpy()
    for nms in ['python_nms', 'diou_python_nms', 'torch_nms']:
        for tag, kw in [('default', {}), ('small', dict(topn=300, max_object_num=20))]:
            dec = D.FCOSDecoder(strides=synth.STRIDES, nms_type=nms, **kw)
            s, c, b = dec(preds)
            out[f'dec_{nms}_{tag}_scores'] = s
            out[f'dec_{nms}_{tag}_classes'] = c
            out[f'dec_{nms}_{tag}_boxes'] = b
    np.savez_compressed(path, **out)


def make_tables(path):
    _, _, A = refload.load()
    out = {'versions': versions()}
    anchors = A.RetinaAnchors(**synth.RETINA_KW)
    for size in (800, 1024):
        p = synth.pyramid_sizes(size)
        levels = anchors([[q, q] for q in p])
        flat = np.concatenate([a.reshape(-1, 4) for a in levels], axis=0)
        # full tables are ~2 MB each: store the base anchors, two rows per level and a checksum
        out[f'anchors_{size}_sum'] = flat.astype(np.float64).sum(axis=0)
        out[f'anchors_{size}_head'] = np.stack([a.reshape(-1, 4)[:18] for a in levels])
        out[f'anchors_{size}_tail'] = np.stack([a.reshape(-1, 4)[-9:] for a in levels])
        pos = A.FCOSPositions(strides=synth.STRIDES)([[q, q] for q in p])
        out[f'positions_{size}_sum'] = np.concatenate([q.reshape(-1, 2) for q in pos]).astype(
            np.float64).sum(axis=0)
    out['base_anchors'] = np.stack(
        [anchors.generate_base_anchors(a, anchors.scales, anchors.ratios) for a in anchors.areas])
    # non-square map and odd strides
    odd = A.RetinaAnchors(areas=[[24, 40], [48, 80]], ratios=[0.5, 1, 2], scales=[1, 1.5],
                          strides=[6, 12])
    lv = odd([[7, 5], [4, 3]])
    out['odd_anchors0'] = lv[0]
    out['odd_anchors1'] = lv[1]
    rng = np.random.RandomState(7)
    x = np.concatenate([rng.normal(0, 2, 4096), rng.uniform(-104, 89, 4096),
                        np.array([0., -0., 1e-30, -1e-30, 88.7228, 88.7229, -103.97, -103.98,
                                  np.inf, -np.inf, np.nan])]).astype(np.float32)
    out['exp_x'] = x
    with np.errstate(all='ignore'):
        out['exp_y'] = np.exp(x)
    np.savez_compressed(path, **out)


FACE_SIZES = [[8, 16, 32], [32, 64, 128], [128, 256, 512]]
FACE_STRIDES = [8, 16, 32]


def make_face(path):
    FL, FD = refload.load_face()
    size, B, G = 256, 3, 10
    gen = torch.Generator().manual_seed(11)
    cls, reg = [], []
    for p in (32, 16, 8):
        cls.append(torch.sigmoid(torch.randn((B, p, p, 3, 1), generator=gen) - 1.5))
        reg.append(torch.randn((B, p, p, 3, 4), generator=gen) * 0.2)
    preds = synth.make_tie_free([cls, reg], min_score=0.3)
    ann = edge_annotations(synth.make_annotations(B, G, size, 1, seed=12), size)
    ann[..., 4] = torch.where(ann[..., 4] >= 0, torch.zeros_like(ann[..., 4]), ann[..., 4])
    out = {'versions': versions(), 'size': size, 'annotations': ann.numpy()}
    for i, (c, r) in enumerate(zip(*preds)):
        out[f'cls{i}'] = c.numpy()
        out[f'reg{i}'] = r.numpy()
    for box_type in ['SmoothL1'] + IOU_TYPES:
        crit = FL.RetinaFaceLoss(anchor_sizes=FACE_SIZES, strides=FACE_STRIDES,
                                 box_loss_type=box_type)
        with torch.no_grad():
            d = crit(preds, ann)
        out[f'loss_{box_type}'] = np.array([d['cls_loss'].item(), d['reg_loss'].item()],
                                           dtype=np.float32)
    crit = FL.RetinaFaceLoss(anchor_sizes=FACE_SIZES, strides=FACE_STRIDES)
    anchors = crit.anchors([[c.shape[2], c.shape[1]] for c in preds[0]])
    flat = torch.cat([torch.tensor(a).view(-1, 4) for a in anchors], dim=0)
    out['assign'] = crit.get_batch_anchors_annotations(flat.unsqueeze(0).repeat(B, 1, 1),
                                                       ann).numpy()
    out['anchors'] = flat.numpy()
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]
    d = crit(p, ann)
    (d['cls_loss'] + 2.0 * d['reg_loss']).backward()
    for i in range(3):
        out[f'gcls_{i}'] = p[0][i].grad.numpy()
        out[f'greg_{i}'] = p[1][i].grad.numpy()
    for nms in ['python_nms', 'diou_python_nms']:
        dec = FD.RetinaFaceDecoder(anchor_sizes=FACE_SIZES, strides=FACE_STRIDES, nms_type=nms)
        s_, c_, b_ = dec(preds)
        out[f'dec_{nms}_scores'] = s_
        out[f'dec_{nms}_classes'] = c_
        out[f'dec_{nms}_boxes'] = b_
    np.savez_compressed(path, **out)


def make_queries(path):
    """Query-based decoders + the stand-alone DecodeMethod / DetNMSMethod: the
    reference's DETRDecoder (softmax, no-object channel, with and without NMS), DINODETRDecoder
    (sigmoid) and DecodeMethod / DetNMSMethod on float boxes."""
    _, D, _ = refload.load()
    gen = torch.Generator().manual_seed(21)
    B, Q, C = 3, 300, 12
    detr_cls = torch.randn((2, B, Q, C + 1), generator=gen) * 2.5      # [layers, B, Q, C+1]
    detr_cls[..., C] += 1.0                                              # no-object wins often
    ctr = torch.rand((2, B, Q, 2), generator=gen)
    wh = torch.rand((2, B, Q, 2), generator=gen) * 0.4 + 0.02
    detr_reg = torch.cat([ctr, wh], dim=-1)
    sizes = [[480, 640], [600, 800], [333, 500]]
    out = {'versions': versions(), 'detr_cls': detr_cls.numpy(), 'detr_reg': detr_reg.numpy(),
           'sizes': np.array(sizes, dtype=np.int64), 'num_classes': C}
    for tag, kw in (('none', dict(nms_type=None)), ('nms', dict(nms_type='python_nms', topn=80,
                                                              max_object_num=20)),
                    ('low', dict(nms_type='diou_python_nms', min_score_threshold=0.3))):
        dec = D.DETRDecoder(num_classes=C, **kw)
        s_, c_, b_ = dec([detr_cls, detr_reg], sizes)
        out[f'detr_{tag}_scores'], out[f'detr_{tag}_classes'], out[f'detr_{tag}_boxes'] = s_, c_, b_
    dino_cls = torch.randn((B, 900, C), generator=gen) * 1.5 - 2.0
    dctr = torch.rand((B, 900, 2), generator=gen)
    dwh = torch.rand((B, 900, 2), generator=gen) * 0.5 + 0.02
    dino_reg = torch.cat([dctr, dwh], dim=-1)
    out['dino_cls'], out['dino_reg'] = dino_cls.numpy(), dino_reg.numpy()
    for nms in ('python_nms', 'diou_python_nms', 'torch_nms'):
        dec = D.DINODETRDecoder(nms_type=nms)
        s_, c_, b_ = dec({'pred_logits': dino_cls, 'pred_boxes': dino_reg}, sizes)
        out[f'dino_{nms}_scores'], out[f'dino_{nms}_classes'], out[f'dino_{nms}_boxes'] = s_, c_, b_
    dec = D.DINODETRDecoder(nms_type='python_nms', min_score_threshold=0.5, nms_threshold=0.3,
                            topn=50, max_object_num=40)
    s_, c_, b_ = dec({'pred_logits': dino_cls, 'pred_boxes': dino_reg}, sizes)
    out['dino_hi_scores'], out['dino_hi_classes'], out['dino_hi_boxes'] = s_, c_, b_
    # DecodeMethod / DetNMSMethod on float boxes
    N = 5000
    scores = torch.rand((2, N), generator=gen).numpy().astype(np.float32)
    classes = torch.randint(0, 7, (2, N), generator=gen).numpy()
    xy = torch.rand((2, N, 2), generator=gen) * 400
    bwh = torch.rand((2, N, 2), generator=gen) * 120 + 4
    boxes = torch.cat([xy, xy + bwh], dim=-1).numpy().astype(np.float32)
    out['dm_scores'], out['dm_classes'], out['dm_boxes'] = scores, classes, boxes
    for nms in ('python_nms', 'diou_python_nms', 'torch_nms'):
        dm = D.DecodeMethod(max_object_num=100, min_score_threshold=0.6, topn=1000, nms_type=nms,
                            nms_threshold=0.5)
        s_, c_, b_ = dm(scores, classes, boxes)
        out[f'dm_{nms}_scores'], out[f'dm_{nms}_classes'], out[f'dm_{nms}_boxes'] = s_, c_, b_
        order = np.argsort(-scores[0], kind='stable')[:700]
        keep = D.DetNMSMethod(nms_type=nms, nms_threshold=0.4)(boxes[0][order], scores[0][order])
        out[f'nms_{nms}_keep'] = np.asarray(keep)
    np.savez_compressed(path, **out)


def _reference_voc_pieces():
    """compute_voc_ap / compute_ious and the matching + AP loop of evaluate_voc_detection, taken
    verbatim from the reference's tools/scripts.py (the module itself cannot be imported here:
    it needs pycocotools and thop).  Returns (namespace with the two functions, source of the
    loop dedented to top level)."""
    import ast
    import textwrap
    path = os.path.join(refload.REFERENCE_ROOT, 'tools', 'scripts.py')
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.splitlines()
    ns = {'np': np}
    loop_src = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ('compute_voc_ap', 'compute_ious'):
            exec(compile('\n'.join(lines[node.lineno - 1:node.end_lineno]), path, 'exec'), ns)
        if isinstance(node, ast.FunctionDef) and node.name == 'evaluate_voc_detection':
            body = lines[node.lineno - 1:node.end_lineno]
            start = next(i for i, l in enumerate(body)
                         if 'all_iou_threshold_map = collections.OrderedDict()' in l)
            end = next(i for i, l in enumerate(body)
                       if 'for key, value in all_iou_threshold_map.items()' in l)
            loop_src = textwrap.dedent('\n'.join(body[start:end]))
    assert loop_src is not None and 'compute_ious' in ns and 'compute_voc_ap' in ns
    return ns, loop_src


def make_voc(path):
    """VOC evaluator matching (tools/scripts.py:455-508, 592-684): the reference's own
    source lines run on synthetic detections / ground truth."""
    import collections
    import types
    ns, loop_src = _reference_voc_pieces()
    rng = np.random.RandomState(31)
    n_img, n_cls, M, G = 12, 4, 30, 9
    preds, gts = [], []
    raw = {'pred_scores': [], 'pred_classes': [], 'pred_boxes': [], 'gt_boxes': [], 'gt_classes': []}
    for i in range(n_img):
        ng = int(rng.randint(0, G + 1))
        gxy = rng.uniform(0, 200, size=(ng, 2))
        gwh = rng.uniform(10, 120, size=(ng, 2))
        gt_boxes = np.concatenate([gxy, gxy + gwh], axis=1).astype(np.float32)
        gt_classes = rng.randint(0, n_cls, size=ng).astype(np.float32)
        if i == 3 and ng >= 2:
            gt_boxes[1] = gt_boxes[0]            # duplicated GT: first maximum wins
            gt_classes[1] = gt_classes[0]
        nd = int(rng.randint(0, M + 1))
        # detections: jittered copies of GT boxes (hits), plus random boxes (misses)
        boxes = np.zeros((nd, 4), dtype=np.float32)
        classes = rng.randint(0, n_cls, size=nd).astype(np.float32)
        for d in range(nd):
            if ng and rng.rand() < 0.7:
                g = int(rng.randint(0, ng))
                boxes[d] = gt_boxes[g] + rng.normal(0, 6, size=4).astype(np.float32)
                classes[d] = gt_classes[g]
            else:
                xy = rng.uniform(0, 200, size=2)
                boxes[d] = np.concatenate([xy, xy + rng.uniform(10, 120, size=2)])
        if nd >= 2:
            boxes[1] = boxes[0]                   # two detections on the same GT: the second is a FP
            classes[1] = classes[0]
        scores = np.sort(rng.uniform(0.05, 1, size=nd).astype(np.float32))[::-1].copy()
        preds.append([boxes, classes, scores])
        gts.append([gt_boxes, gt_classes])
        for k, v in zip(raw, (scores, classes, boxes, gt_boxes, gt_classes)):
            raw[k].append(v)
    thresholds = [0.5, 0.75, 0.3]
    config = types.SimpleNamespace(eval_voc_iou_threshold_list=thresholds, num_classes=n_cls)
    env = dict(ns)
    env.update(collections=collections, tqdm=lambda x: x, config=config, preds=preds, gts=gts)
    exec(loop_src, env)
    out = {'versions': versions(), 'thresholds': np.array(thresholds), 'num_classes': n_cls,
           'n_images': n_img}
    for i in range(n_img):
        for k in raw:
            out[f'{k}_{i}'] = raw[k][i]
    for key, value in env['all_iou_threshold_map'].items():
        out['map::' + key] = np.float64(value)
    for key, per_class in env['all_iou_threshold_per_class_ap'].items():
        out['ap::' + key] = np.array([per_class[c] for c in range(n_cls)], dtype=np.float64)
    # compute_ious on its own (incl. a degenerate pair -> NaN) and compute_voc_ap both ways
    a = np.concatenate([raw['gt_boxes'][0], np.array([[5, 5, 5, 5]], dtype=np.float32)])
    b = np.concatenate([raw['pred_boxes'][0][:7], np.array([[5, 5, 5, 5]], dtype=np.float32)])
    out['ious_a'], out['ious_b'] = a, b
    with np.errstate(invalid='ignore', divide='ignore'):
        out['ious'] = ns['compute_ious'](a, b)
    rec = np.array([0.1, 0.1, 0.2, 0.4, 0.4, 0.7])
    prec = np.array([1.0, 0.5, 0.66, 0.75, 0.6, 0.58])
    out['ap_rec'], out['ap_prec'] = rec, prec
    out['ap_07'] = np.float64(ns['compute_voc_ap'](rec, prec, use_07_metric=True))
    out['ap_10'] = np.float64(ns['compute_voc_ap'](rec, prec, use_07_metric=False))
    np.savez_compressed(path, **out)


def iou_method_boxes(seed=5, n=96):
    """Box pairs for the IoUMethod fixtures: overlapping, disjoint, nested, identical (max / min
    ties), shared edges, zero-area and inverted boxes."""
    rng = np.random.RandomState(seed)
    xy = rng.uniform(0, 200, size=(n, 2))
    b1 = np.concatenate([xy, xy + rng.uniform(4, 120, size=(n, 2))], axis=1).astype(np.float32)
    b2 = (b1 + rng.normal(0, 12, size=(n, 4))).astype(np.float32)
    far = rng.uniform(300, 500, size=(n, 2))
    b2[0:8] = np.concatenate([far, far + rng.uniform(4, 60, size=(n, 2))], axis=1)[0:8]   # disjoint
    b2[8:12] = b1[8:12]                                                   # identical
    b2[12:16, 0:2] = b1[12:16, 0:2]                                       # shared corner
    b2[16:20] = b1[16:20] + np.array([10, 10, -10, -10], dtype=np.float32)  # nested
    b2[20, 2] = b2[20, 0]                                                 # zero width
    b1[21, 3] = b1[21, 1]                                                 # zero height
    b2[22] = b2[22][[2, 3, 0, 1]]                                         # inverted
    b1[23] = np.trunc(b1[23]); b2[23] = b1[23] + np.array([0, 0, 5, 0], dtype=np.float32)
    return b1, b2


def make_iou_method(path):
    """IoUMethod (losses.py:28-123): values and autograd gradients w.r.t. both
    inputs from the unmodified reference class, all IoU types x both box formats (2-D), and the
    assignment's [N,1,4] x [1,M,4] broadcast for the types whose indexing allows it."""
    L, _, _ = refload.load()
    fn = L.IoUMethod()
    b1, b2 = iou_method_boxes()
    out = {'versions': versions(), 'b1': b1, 'b2': b2}
    rng = np.random.RandomState(9)
    up = rng.uniform(0.5, 1.5, size=b1.shape[0]).astype(np.float32)
    out['upstream'] = up
    # xywh inputs: the same boxes re-expressed (inverted / degenerate ones give w, h <= 0)
    def to_xywh(b):
        return np.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0],
                         b[:, 3] - b[:, 1]], axis=1).astype(np.float32)
    out['b1_xywh'], out['b2_xywh'] = to_xywh(b1), to_xywh(b2)
    for box_type, (x1, x2) in (('xyxy', (b1, b2)), ('xywh', (out['b1_xywh'], out['b2_xywh']))):
        for iou_type in IOU_TYPES:
            t1 = torch.from_numpy(x1.copy()).requires_grad_(True)
            t2 = torch.from_numpy(x2.copy()).requires_grad_(True)
            v = fn(t1, t2, iou_type=iou_type, box_type=box_type)
            (v * torch.from_numpy(up)).sum().backward()
            out[f'{box_type}_{iou_type}'] = v.detach().numpy()
            out[f'{box_type}_{iou_type}_g1'] = t1.grad.numpy()
            out[f'{box_type}_{iou_type}_g2'] = t2.grad.numpy()
    a, g = torch.from_numpy(b1[:40].copy()), torch.from_numpy(b2[:17].copy())
    for iou_type in ('IoU', 'DIoU', 'EIoU'):
        out[f'bcast_{iou_type}'] = fn(a.unsqueeze(1), g.unsqueeze(0), iou_type=iou_type).numpy()
    a.requires_grad_(True)
    g.requires_grad_(True)
    w = torch.from_numpy(rng.uniform(0.5, 1.5, size=(40, 17)).astype(np.float32))
    (fn(a.unsqueeze(1), g.unsqueeze(0), iou_type='DIoU') * w).sum().backward()
    out['bcast_w'], out['bcast_DIoU_g1'], out['bcast_DIoU_g2'] = w.numpy(), a.grad.numpy(), g.grad.numpy()
    np.savez_compressed(path, **out)


def make_heads(path):
    """Head tail: the reference's RetinaClsHead (models/head.py:15-52) on a seeded
    feature map; the convolution output is captured with a hook, the head's own `.float()` +
    sigmoid produce the probabilities, and the three lines of RetinaNet.forward
    (models/retinanet.py:73-76) permute / view them.  Also the gradient w.r.t. the conv output."""
    refload.load()
    from simpleAICV.detection.models import head as H
    torch.manual_seed(5)
    A, C = 3, 5
    m = H.RetinaClsHead(16, A, C, num_layers=1)
    torch.nn.init.normal_(m.cls_out.weight, std=1.5)      # spread the logits (bias is -4.6)
    captured = {}

    def hook(_mod, _inp, out):
        out.retain_grad()
        captured['x'] = out

    m.cls_out.register_forward_hook(hook)
    feat = torch.randn(2, 16, 7, 9) * 2
    y = m(feat)
    y = y.permute(0, 2, 3, 1).contiguous()
    y = y.view(y.shape[0], y.shape[1], y.shape[2], -1, C)
    wgt = torch.randn(y.shape)
    (y * wgt).sum().backward()
    np.savez_compressed(path, versions=versions(), x=captured['x'].detach().numpy(),
                        y=y.detach().numpy(), wgt=wgt.numpy(), gx=captured['x'].grad.numpy())


def make_half_nan(path):
    """(1) Half-precision regression heads WITHOUT autocast (model.half()-style eager arithmetic): the
    reference then runs torch.exp / np.exp on float16 and rounds the result to half
    (losses.py:417-426, :568; decode.py:257-268, :356).  Loss values, gradients and decoder outputs
    of the unmodified classes for float16 `reg` (and bfloat16 for the losses; the decoders cannot
    take bfloat16: `.numpy()` raises).
    (2) NaN class / centre-ness scores in the decoders: np.argmax stops at the first NaN
    (decode.py:230-238, :319-338), the row's score is NaN and fails `score > threshold`."""
    L, D, _ = refload.load()
    out = {'versions': versions()}
    # ---- Retina ----
    size, C, B, G = 128, 8, 3, 12
    preds = synth.make_tie_free(synth.make_retina_preds(B, size, C, seed=40))
    ann = edge_annotations(synth.make_annotations(B, G, size, C, seed=41), size)
    out['r_annotations'] = ann.numpy()
    for i, (c, r) in enumerate(zip(*preds)):
        out[f'r_cls{i}'] = c.numpy()
        out[f'r_reg{i}_f16'] = r.half().numpy()
        out[f'r_reg{i}_bf16_bits'] = r.bfloat16().view(torch.int16).numpy()
    for tag, cast in (('f16', torch.float16), ('bf16', torch.bfloat16)):
        hp = [preds[0], [r.to(cast) for r in preds[1]]]
        for box_type in ('SmoothL1', 'GIoU', 'CIoU'):
            crit = L.RetinaLoss(**synth.RETINA_KW, box_loss_type=box_type)
            with torch.no_grad():
                d = crit(hp, ann)
            out[f'r_loss_{tag}_{box_type}'] = np.array(
                [d['cls_loss'].float().item(), d['reg_loss'].float().item()], dtype=np.float32)
        crit = L.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        p = [[t.clone().requires_grad_(True) for t in grp] for grp in hp]
        d = crit(p, ann)
        (d['cls_loss'] + 2.0 * d['reg_loss']).backward()
        for i in range(len(p[0])):
            out[f'r_greg_{tag}_{i}'] = p[1][i].grad.float().numpy()
    hp = [preds[0], [r.half() for r in preds[1]]]
    s_, c_, b_ = D.RetinaDecoder(**synth.RETINA_KW)(hp)
    out['r_dec_f16_scores'], out['r_dec_f16_classes'], out['r_dec_f16_boxes'] = s_, c_, b_
    # ---- FCOS ----
    size = 256
    fp = synth.make_tie_free(synth.make_fcos_preds(B, size, C, seed=42))
    fann = edge_annotations(synth.make_annotations(B, G, size, C, seed=43), size)
    out['f_annotations'] = fann.numpy()
    for i, (c, r, t) in enumerate(zip(*fp)):
        out[f'f_cls{i}'] = c.numpy()
        out[f'f_reg{i}_f16'] = r.half().numpy()
        out[f'f_ctr{i}'] = t.numpy()
    hp = [fp[0], [r.half() for r in fp[1]], fp[2]]
    for iou_type in ('GIoU', 'DIoU'):
        crit = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI, box_loss_iou_type=iou_type)
        with torch.no_grad():
            d = crit(hp, fann)
        out[f'f_loss_f16_{iou_type}'] = np.array(
            [d['cls_loss'].float().item(), d['reg_loss'].float().item(),
             d['center_ness_loss'].float().item()], dtype=np.float32)
    crit = L.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    p = [[t.clone().requires_grad_(True) for t in grp] for grp in hp]
    d = crit(p, fann)
    (d['cls_loss'] + 2.0 * d['reg_loss'] + 3.0 * d['center_ness_loss']).backward()
    for i in range(len(p[0])):
        out[f'f_greg_f16_{i}'] = p[1][i].grad.float().numpy()
    s_, c_, b_ = D.FCOSDecoder(strides=synth.STRIDES)(hp)
    out['f_dec_f16_scores'], out['f_dec_f16_classes'], out['f_dec_f16_boxes'] = s_, c_, b_

    # ---- NaN scores in the decoders ----
    nan = float('nan')
    npreds = [[t.clone() for t in grp] for grp in preds]
    c0 = npreds[0][0].view(B, -1, C)
    top = c0[0].max(dim=1).values.argsort(descending=True)
    c0[0, top[0], 0] = nan                 # best row of image 0: NaN in the first class
    c0[0, top[1], C - 1] = nan             # second best: NaN in the last class, real maximum before it
    c0[0, top[2], :] = nan                 # third: all NaN
    c0[1, top[0], 3] = nan
    npreds[0][2].view(B, -1, C)[2, 5, 1] = nan
    for i, c in enumerate(npreds[0]):
        out[f'rn_cls{i}'] = c.numpy()
        out[f'rn_reg{i}'] = npreds[1][i].numpy()
    for nms in ('python_nms', 'torch_nms'):
        s_, c_, b_ = D.RetinaDecoder(**synth.RETINA_KW, nms_type=nms)(npreds)
        out[f'rn_dec_{nms}_scores'], out[f'rn_dec_{nms}_classes'], out[f'rn_dec_{nms}_boxes'] = s_, c_, b_
    fpn = [[t.clone() for t in grp] for grp in fp]
    c0 = fpn[0][0].view(B, -1, C)
    t0 = fpn[2][0].view(B, -1)
    score = (c0.max(dim=2).values * t0).sqrt()
    top = score[0].argsort(descending=True)
    c0[0, top[0], 2] = nan                 # NaN class score
    t0[0, top[1]] = nan                    # NaN centre-ness
    c0[0, top[2], C - 1] = nan
    t0[1, top[0]] = nan
    for i in range(len(fpn[0])):
        out[f'fn_cls{i}'] = fpn[0][i].numpy()
        out[f'fn_reg{i}'] = fpn[1][i].numpy()
        out[f'fn_ctr{i}'] = fpn[2][i].numpy()
    with np.errstate(invalid='ignore'):
        s_, c_, b_ = D.FCOSDecoder(strides=synth.STRIDES)(fpn)
    out['fn_dec_scores'], out['fn_dec_classes'], out['fn_dec_boxes'] = s_, c_, b_
    np.savez_compressed(path, **out)


if __name__ == '__main__':
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'half':
        make_half_nan(os.path.join(HERE, 'half_nan.npz'))
        print('half_nan.npz', os.path.getsize(os.path.join(HERE, 'half_nan.npz')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'heads':
        make_heads(os.path.join(HERE, 'head_tail.npz'))
        print('head_tail.npz', os.path.getsize(os.path.join(HERE, 'head_tail.npz')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'iou':
        make_iou_method(os.path.join(HERE, 'iou_method.npz'))
        print('iou_method.npz', os.path.getsize(os.path.join(HERE, 'iou_method.npz')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'voc':
        make_voc(os.path.join(HERE, 'voc_eval.npz'))
        print('voc_eval.npz', os.path.getsize(os.path.join(HERE, 'voc_eval.npz')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'queries':
        make_queries(os.path.join(HERE, 'queries.npz'))
        print('queries.npz', os.path.getsize(os.path.join(HERE, 'queries.npz')))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'face':
        make_face(os.path.join(HERE, 'retinaface_small.npz'))
        print('retinaface_small.npz', os.path.getsize(os.path.join(HERE, 'retinaface_small.npz')))
        sys.exit(0)
    make_retina(os.path.join(HERE, 'retina_small.npz'))
    make_fcos(os.path.join(HERE, 'fcos_small.npz'))
    make_tables(os.path.join(HERE, 'tables.npz'))
    make_face(os.path.join(HERE, 'retinaface_small.npz'))
    make_heads(os.path.join(HERE, 'head_tail.npz'))
    make_queries(os.path.join(HERE, 'queries.npz'))
    make_voc(os.path.join(HERE, 'voc_eval.npz'))
    make_iou_method(os.path.join(HERE, 'iou_method.npz'))
    make_half_nan(os.path.join(HERE, 'half_nan.npz'))
    for f in ('retina_small.npz', 'fcos_small.npz', 'tables.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)))
