"""oracle/det_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch-CPU for the losses, NumPy for the decoders -- the same third-party
primitives the reference itself runs on) of the reference's dense-detection hot path:

    simpleAICV/detection/models/anchor.py:5-130   RetinaAnchors / FCOSPositions
    simpleAICV/detection/losses.py:28-123         IoUMethod
    simpleAICV/detection/losses.py:126-429        RetinaLoss
    simpleAICV/detection/losses.py:432-833        FCOSLoss
    simpleAICV/detection/decode.py:26-104         DetNMSMethod
    simpleAICV/detection/decode.py:107-172        DecodeMethod
    simpleAICV/detection/decode.py:175-271        RetinaDecoder
    simpleAICV/detection/decode.py:274-364        FCOSDecoder
    simpleAICV/detection/decode.py:367-482        DETRDecoder
    simpleAICV/detection/decode.py:485-594        DINODETRDecoder
    tools/scripts.py:455-508, 592-684             compute_voc_ap / compute_ious / the matching and
                                                  AP loop of evaluate_voc_detection

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module; the product (``b200det``) never does and has no
CPU fallback.

Pinning.  The reference ships no golden vectors or tests for this path,
so the oracle is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py``
imports the unmodified reference from /root/reference (one ``traitlets`` import shim), runs it on
seeded inputs and stores inputs + outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` requires this module to reproduce them bit-for-bit (labels,
matched indices, targets, top-n order, keep lists, decoded boxes AND the float loss values), and
``tests/test_oracle_vs_reference.py`` repeats that live whenever /root/reference is present.
The float32 ``np.exp`` used by the decoders is restated in ``oracle/npexp.c`` and was compared
with NumPy 2.3.5 on all 2^32 inputs (0 mismatches).

Every function returns the reference's result plus the intermediate "truth" the CUDA parity
tests compare against (labels, matched-GT index in the *filtered* GT list, sorted top-n
indices, NMS keep list).
"""
import ctypes
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

IOU_TYPES = ('IoU', 'GIoU', 'DIoU', 'CIoU', 'EIoU')


def _oracle_lib():
    """Loads oracle/_build/liboracle.so (built by `make -C oracle` / __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(path):
            raise RuntimeError(
                'oracle helper not built: run `make -C oracle` (or __graft_entry__.build())')
        lib = ctypes.CDLL(path)
        lib.oracle_npexp_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        lib.oracle_npexp_f32.restype = None
        _LIB = lib
    return _LIB


def np_exp_f32(x):
    """float32 exp with NumPy's SIMD algorithm, independent of the host's dispatch
    (decode.py:260, :356 call np.exp on float32 arrays)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _oracle_lib().oracle_npexp_f32(x.ctypes.data, out.ctypes.data, x.size)
    return out


# ----------------------------------------------------------------------------------------
# anchors / positions (models/anchor.py)
# ----------------------------------------------------------------------------------------
def retina_base_anchors(size, scales, ratios):
    """anchor.py:35-57 -- nine [x1,y1,x2,y2] boxes centred on the origin, ratio-major /
    scale-minor, float32 arithmetic on a float64 sqrt."""
    size = np.asarray(size, dtype=np.float32)
    scales = np.asarray(scales, dtype=np.float32)
    ratios = np.asarray(ratios, dtype=np.float32)
    aspect_rows = []
    for r in ratios:
        for s in scales:
            # np.float32 * python-float -> float32 product (NEP 50), as in anchor.py:40
            aspect_rows.append([s * math.sqrt(r), s * math.sqrt(1 / r)])
    aspects = np.array(aspect_rows, dtype=np.float32)
    wh = size * aspects                                   # anchor.py:48
    base = np.zeros((aspects.shape[0], 4), dtype=np.float32)
    base[:, 2:] += wh
    base[:, 0] -= base[:, 2] / 2
    base[:, 1] -= base[:, 3] / 2
    base[:, 2] /= 2
    base[:, 3] /= 2
    return base


def retina_anchors(feature_sizes, areas, ratios, scales, strides):
    """anchor.py:18-86.  feature_sizes = [[W,H], ...]; returns per level float32 [H,W,9,4]."""
    out = []
    for size, (fw, fh), stride in zip(areas, feature_sizes, strides):
        base = retina_base_anchors(size, scales, ratios)
        stride = np.float32(stride)
        sx = ((np.arange(0, fw) + 0.5) * stride).astype(np.float32)
        sy = ((np.arange(0, fh) + 0.5) * stride).astype(np.float32)
        shifts = np.empty((fh, fw, 1, 4), dtype=np.float32)
        shifts[:, :, 0, 0] = sx[None, :]
        shifts[:, :, 0, 1] = sy[:, None]
        shifts[:, :, 0, 2] = sx[None, :]
        shifts[:, :, 0, 3] = sy[:, None]
        out.append(np.ascontiguousarray(base[None, None, :, :] + shifts, dtype=np.float32))
    return out


def fcos_positions(feature_sizes, strides):
    """anchor.py:94-130.  Returns per level float32 [H,W,2] = (x_center, y_center)."""
    out = []
    for (fw, fh), stride in zip(feature_sizes, strides):
        stride = np.float32(stride)
        sx = ((np.arange(0, fw) + 0.5) * stride).astype(np.float32)
        sy = ((np.arange(0, fh) + 0.5) * stride).astype(np.float32)
        pos = np.empty((fh, fw, 2), dtype=np.float32)
        pos[:, :, 0] = sx[None, :]
        pos[:, :, 1] = sy[:, None]
        out.append(pos)
    return out


def feature_sizes_of(level_tensors):
    """[[W,H], ...] from channels-last head outputs (losses.py:169-171, decode.py:203-205)."""
    return [[int(t.shape[2]), int(t.shape[1])] for t in level_tensors]


# ----------------------------------------------------------------------------------------
# IoU family (losses.py:33-123)
# ----------------------------------------------------------------------------------------
def box_iou(b1, b2, iou_type='IoU'):
    """xyxy boxes, broadcasting over leading dims; float32 op order of losses.py:54-123."""
    assert iou_type in IOU_TYPES
    lt = torch.max(b1[..., 0:2], b2[..., 0:2])
    rb = torch.min(b1[..., 2:4], b2[..., 2:4])
    inter_wh = torch.clamp(rb - lt, min=0)
    inter = inter_wh[..., 0] * inter_wh[..., 1]
    wh1 = torch.clamp(b1[..., 2:4] - b1[..., 0:2], min=0)
    wh2 = torch.clamp(b2[..., 2:4] - b2[..., 0:2], min=0)
    area1 = wh1[..., 0] * wh1[..., 1]
    area2 = wh2[..., 0] * wh2[..., 1]
    union = torch.clamp(area1 + area2 - inter, min=1e-4)
    iou = inter / union
    if iou_type == 'IoU':
        return iou
    enc_lt = torch.min(b1[..., 0:2], b2[..., 0:2])
    enc_rb = torch.max(b1[..., 2:4], b2[..., 2:4])
    enc_wh = torch.clamp(enc_rb - enc_lt, min=0)
    if iou_type == 'GIoU':
        enc = torch.clamp(enc_wh[..., 0] * enc_wh[..., 1], min=1e-4)
        return iou - (enc - union) / enc
    c2 = torch.clamp(enc_wh[..., 0]**2 + enc_wh[..., 1]**2, min=1e-4)
    ctr1 = (b1[..., 2:4] + b1[..., 0:2]) / 2
    ctr2 = (b2[..., 2:4] + b2[..., 0:2]) / 2
    p2 = (ctr1[..., 0] - ctr2[..., 0])**2 + (ctr1[..., 1] - ctr2[..., 1])**2
    if iou_type == 'DIoU':
        return iou - p2 / c2
    if iou_type == 'CIoU':
        v = (4 / math.pi**2) * torch.pow(
            torch.atan(wh2[..., 0] / wh2[..., 1]) - torch.atan(wh1[..., 0] / wh1[..., 1]), 2)
        with torch.no_grad():
            alpha = v / torch.clamp(1 - iou + v, min=1e-4)
        return iou - (p2 / c2 + v * alpha)
    # EIoU
    pw2 = (wh2[..., 0] - wh1[..., 0])**2
    ph2 = (wh2[..., 1] - wh1[..., 1])**2
    cw2 = torch.clamp(enc_wh[..., 0]**2, min=1e-4)
    ch2 = torch.clamp(enc_wh[..., 1]**2, min=1e-4)
    return iou - (p2 / c2 + pw2 / cw2 + ph2 / ch2)


def iou_method(boxes1, boxes2, iou_type='IoU', box_type='xyxy'):
    """IoUMethod.__call__ (losses.py:33-123) incl. the 'xywh' input option (:44-52, 2-D boxes)."""
    assert box_type in ['xyxy', 'xywh']
    if box_type == 'xywh':
        boxes1 = torch.cat([boxes1[..., 0:2] - boxes1[..., 2:4] / 2,
                            boxes1[..., 0:2] + boxes1[..., 2:4] / 2], dim=1)
        boxes2 = torch.cat([boxes2[..., 0:2] - boxes2[..., 2:4] / 2,
                            boxes2[..., 0:2] + boxes2[..., 2:4] / 2], dim=1)
    return box_iou(boxes1, boxes2, iou_type)


# ----------------------------------------------------------------------------------------
# shared focal loss (losses.py:220-261 and :513-548)
# ----------------------------------------------------------------------------------------
def _focal_sum(cls_rows, labels, alpha, gamma):
    """cls_rows [N,C] clamped probabilities, labels [N] float (0 = background, k = class k-1).
    Returns the un-normalised focal sum exactly as the reference accumulates it."""
    num_classes = cls_rows.shape[1]
    onehot = torch.nn.functional.one_hot(labels.long(), num_classes=num_classes + 1)
    onehot = onehot[:, 1:].float()
    is_pos = torch.eq(onehot, 1.)
    alpha_t = torch.ones_like(cls_rows) * alpha
    alpha_t = torch.where(is_pos, alpha_t, 1. - alpha_t)
    pt = torch.where(is_pos, cls_rows, 1. - cls_rows)
    weight = alpha_t * torch.pow((1. - pt), gamma)
    bce = -(onehot * torch.log(cls_rows) + (1. - onehot) * torch.log(1. - cls_rows))
    return (weight * bce).sum()


# ----------------------------------------------------------------------------------------
# RetinaLoss (losses.py:126-429)
# ----------------------------------------------------------------------------------------
def retina_encode(gt_boxes, anchors):
    """losses.py:390-409 -- (tx,ty,tw,th) regression targets, no std scaling."""
    a_wh = anchors[:, 2:] - anchors[:, :2]
    a_ctr = anchors[:, :2] + 0.5 * a_wh
    g_wh = torch.clamp(gt_boxes[:, 2:] - gt_boxes[:, :2], min=1e-4)
    g_ctr = gt_boxes[:, :2] + 0.5 * g_wh
    return torch.cat([(g_ctr - a_ctr) / a_wh, torch.log(g_wh / a_wh)], dim=1)


def retina_decode_boxes(deltas, anchors):
    """losses.py:411-429 -- torch version used inside the loss."""
    a_wh = anchors[:, 2:4] - anchors[:, 0:2]
    a_ctr = anchors[:, 0:2] + 0.5 * a_wh
    wh = torch.exp(deltas[:, 2:4]) * a_wh
    ctr = deltas[:, :2] * a_wh + a_ctr
    return torch.cat([ctr - 0.5 * wh, ctr + 0.5 * wh], dim=1)


def retina_assign(anchors, annotations, box_loss_type='SmoothL1', neg_thr=0.4, pos_thr=0.5):
    """losses.py:322-388 (thresholds 0.4 / 0.5) and face_detection/losses.py:222-292
    (thresholds 0.35 / 0.35).  anchors [A,4] float32 tensor, annotations [B,G,5].

    Returns (targets [B,A,5], labels [B,A] int64, matched [B,A] int64).  `matched` indexes the
    image's FILTERED GT list (rows with class >= 0); -1 for images without GT."""
    num_anchors = anchors.shape[0]
    device = annotations.device
    all_targets, all_labels, all_matched = [], [], []
    for annots in annotations:
        annots = annots[annots[:, 4] >= 0]
        if annots.shape[0] == 0:
            targets = torch.ones([num_anchors, 5], dtype=torch.float32, device=device) * (-1)
            matched = torch.full([num_anchors], -1, dtype=torch.int64, device=device)
        else:
            gt_boxes, gt_cls = annots[:, 0:4], annots[:, 4]
            ious = box_iou(anchors.unsqueeze(1), gt_boxes.unsqueeze(0), 'IoU')
            best_iou, matched = ious.max(axis=1)
            labels = torch.ones_like(best_iou) * -1
            labels[best_iou < neg_thr] = 0
            labels[best_iou >= pos_thr] = gt_cls[matched][best_iou >= pos_thr] + 1
            boxes = gt_boxes[matched]
            if box_loss_type == 'SmoothL1':
                boxes = retina_encode(boxes, anchors)
            targets = torch.cat([boxes, labels.unsqueeze(-1)], dim=1)
        all_targets.append(targets.unsqueeze(0))
        all_labels.append(targets[:, 4].long().unsqueeze(0))
        all_matched.append(matched.unsqueeze(0))
    return torch.cat(all_targets, 0), torch.cat(all_labels, 0), torch.cat(all_matched, 0)


def retina_loss(preds, annotations, areas, ratios, scales, strides, alpha=0.25, gamma=2,
                beta=1.0 / 9.0, cls_loss_weight=1., box_loss_weight=1.,
                box_loss_type='SmoothL1', level_anchors=None, neg_thr=0.4, pos_thr=0.5):
    """losses.py:161-320.  Returns dict with the reference's loss dict plus intermediates:
    'labels' [B,A], 'matched' [B,A], 'num_pos', 'cls_sum', 'reg_sum' (un-normalised)."""
    cls_levels, reg_levels = preds
    batch = annotations.shape[0]
    if level_anchors is None:
        level_anchors = retina_anchors(feature_sizes_of(cls_levels), areas, ratios, scales,
                                       strides)
    device = annotations.device
    anchors = torch.cat([torch.tensor(a).view(-1, 4) for a in level_anchors], dim=0).to(device)
    targets, labels, matched = retina_assign(anchors, annotations, box_loss_type, neg_thr, pos_thr)

    cls = torch.cat([c.view(c.shape[0], -1, c.shape[-1]) for c in cls_levels], dim=1)
    reg = torch.cat([r.view(r.shape[0], -1, r.shape[-1]) for r in reg_levels], dim=1)
    cls = torch.clamp(cls, min=1e-4, max=1. - 1e-4)
    cls = cls.view(-1, cls.shape[-1])
    reg = reg.view(-1, reg.shape[-1])
    batch_anchors = anchors.unsqueeze(0).repeat(batch, 1, 1).view(-1, 4)
    flat = targets.view(-1, 5)

    out = {'labels': labels, 'matched': matched}
    # focal (losses.py:220-261)
    used = flat[:, 4] >= 0
    cls_used, flat_used = cls[used], flat[used]
    num_pos = int((flat_used[:, 4] > 0).sum())
    out['num_pos'] = num_pos
    if num_pos == 0:
        cls_loss = torch.tensor(0.).to(device)
        out['cls_sum'] = torch.tensor(0.)
    else:
        cls_sum = _focal_sum(cls_used, flat_used[:, 4], alpha, gamma)
        out['cls_sum'] = cls_sum.detach()
        cls_loss = cls_sum / num_pos
    # box (losses.py:263-320)
    pos = flat[:, 4] > 0
    reg_pos, anc_pos, flat_pos = reg[pos], batch_anchors[pos], flat[pos]
    if flat_pos.shape[0] == 0:
        reg_loss = torch.tensor(0.).to(device)
        out['reg_sum'] = torch.tensor(0.)
    elif box_loss_type == 'SmoothL1':
        x = torch.abs(reg_pos - flat_pos[:, 0:4])
        per = torch.where(torch.ge(x, beta), x - 0.5 * beta, 0.5 * (x**2) / beta)
        reg_sum = per.sum()
        out['reg_sum'] = reg_sum.detach()
        reg_loss = reg_sum / flat_pos.shape[0]
    else:
        boxes = retina_decode_boxes(reg_pos, anc_pos)
        ious = box_iou(boxes, flat_pos[:, 0:4], box_loss_type)
        reg_sum = (1 - ious).sum()
        out['reg_sum'] = reg_sum.detach()
        reg_loss = reg_sum / flat_pos.shape[0]
    out['cls_loss'] = cls_loss_weight * cls_loss
    out['reg_loss'] = box_loss_weight * reg_loss
    return out


# ----------------------------------------------------------------------------------------
# FCOSLoss (losses.py:432-833)
# ----------------------------------------------------------------------------------------
def fcos_assign(points, point_mi, point_stride, annotations, center_sample_radius=1.5,
                use_center_sample=True):
    """losses.py:663-829.  points [P,2], point_mi [P,2], point_stride [P,1] float32 tensors.

    Returns (targets [B,P,6] = l,t,r,b,label,centerness ; labels [B,P] int64 ;
    matched [B,P] int64 = index in the filtered GT list, -1 for background)."""
    num_points = points.shape[0]
    device = annotations.device
    all_targets, all_matched = [], []
    for annots in annotations:
        annots = annots[annots[:, 4] >= 0]
        targets = torch.zeros([num_points, 6], dtype=torch.float32, device=device)
        matched = torch.full([num_points], -1, dtype=torch.int64, device=device)
        if annots.shape[0] > 0:
            num_gt = annots.shape[0]
            gt = annots[:, 0:4]
            cand = torch.zeros([num_points, num_gt, 4], dtype=torch.float32,
                               device=device) + gt.unsqueeze(0)
            pts = points.unsqueeze(1).repeat(1, num_gt, 1)
            if use_center_sample:
                gt_ctr = (cand[:, :, 2:4] + cand[:, :, 0:2]) / 2
                radius = (point_stride * center_sample_radius).repeat(1, num_gt)
            cand[:, :, 0:2] = pts[:, :, 0:2] - cand[:, :, 0:2]
            cand[:, :, 2:4] = cand[:, :, 2:4] - pts[:, :, 0:2]
            inside = (cand.min(axis=-1, keepdim=True)[0][:, :, 0] > 0).int().unsqueeze(-1)
            cand = cand * inside
            if use_center_sample:
                dist = torch.sqrt((pts[:, :, 0] - gt_ctr[:, :, 0])**2 +
                                  (pts[:, :, 1] - gt_ctr[:, :, 1])**2)
                cand = cand * (dist < radius).int().unsqueeze(-1)
            longest = cand.max(axis=-1, keepdim=True)[0]
            mi = point_mi.unsqueeze(1).repeat(1, num_gt, 1)
            cand = cand * (longest[:, :, 0] > mi[:, :, 0]).int().unsqueeze(-1)
            cand = cand * (longest[:, :, 0] < mi[:, :, 1]).int().unsqueeze(-1)
            is_pos = cand.sum(axis=-1).sum(axis=-1) > 0
            pos_idx = is_pos.nonzero(as_tuple=False).squeeze(dim=-1)
            if len(pos_idx) > 0:
                pos_cand = cand[pos_idx]
                gt_cls = annots[:, 4]
                if num_gt == 1:
                    choice = torch.zeros([pos_cand.shape[0]], dtype=torch.int64, device=device)
                else:
                    gt_wh = gt[:, 2:4] - gt[:, 0:2]
                    gt_area = (gt_wh[:, 0] * gt_wh[:, 1]).unsqueeze(0).repeat(
                        pos_cand.shape[0], 1)
                    big = torch.ones_like(gt_area) * 100000000
                    gt_area = torch.where(torch.eq(pos_cand.sum(axis=2), 0.), big, gt_area)
                    choice = gt_area.min(axis=1)[1]
                rows = torch.arange(pos_cand.shape[0], device=device)
                targets[pos_idx, 0:4] = pos_cand[rows, choice, :]
                targets[pos_idx, 4] = gt_cls[choice] + 1
                l, t = targets[pos_idx, 0:1], targets[pos_idx, 1:2]
                r, b = targets[pos_idx, 2:3], targets[pos_idx, 3:4]
                targets[pos_idx, 5:6] = torch.sqrt(
                    (torch.min(l, r) / torch.max(l, r)) * (torch.min(t, b) / torch.max(t, b)))
                matched[pos_idx] = choice
        all_targets.append(targets.unsqueeze(0))
        all_matched.append(matched.unsqueeze(0))
    targets = torch.cat(all_targets, 0)
    return targets, targets[:, :, 4].long(), torch.cat(all_matched, 0)


def fcos_point_tables(reg_levels, strides, mi):
    """losses.py:623-661 -- per-point position, scale range and stride (first image only)."""
    sizes = feature_sizes_of(reg_levels)
    positions = fcos_positions(sizes, strides)
    pts, pmi, pst = [], [], []
    for pos, rng, stride in zip(positions, mi, strides):
        n = pos.shape[0] * pos.shape[1]
        pts.append(torch.tensor(pos).view(-1, 2))
        pmi.append(torch.zeros(n, 2) + torch.tensor(rng))
        pst.append(torch.zeros(n, 1) + stride)
    return torch.cat(pts, 0), torch.cat(pmi, 0), torch.cat(pst, 0)


def fcos_loss(preds, annotations, strides, mi, alpha=0.25, gamma=2., cls_loss_weight=1.,
              box_loss_weight=1., center_ness_loss_weight=1., box_loss_iou_type='GIoU',
              center_sample_radius=1.5, use_center_sample=True):
    """losses.py:462-610.  Returns the loss dict plus 'labels', 'matched', 'targets' [B,P,6],
    'num_pos' and the three un-normalised sums."""
    cls_levels, reg_levels, ctr_levels = preds
    batch = annotations.shape[0]
    device = annotations.device
    points, point_mi, point_stride = [
        t.to(device) for t in fcos_point_tables(reg_levels, strides, mi)]
    targets, labels, matched = fcos_assign(points, point_mi, point_stride, annotations,
                                           center_sample_radius, use_center_sample)
    cls = torch.cat([c.view(c.shape[0], -1, c.shape[-1]) for c in cls_levels], dim=1)
    reg = torch.cat([r.view(r.shape[0], -1, r.shape[-1]) for r in reg_levels], dim=1)
    ctr = torch.cat([c.view(c.shape[0], -1, c.shape[-1]) for c in ctr_levels], dim=1)
    full = torch.cat([targets, points.unsqueeze(0).repeat(batch, 1, 1)], dim=2)

    cls = cls.view(-1, cls.shape[-1])
    reg = reg.view(-1, 4)
    ctr = ctr.view(-1, 1)
    full = full.view(-1, 8)
    cls = torch.clamp(cls, min=1e-4, max=1. - 1e-4)
    ctr = torch.clamp(ctr, min=1e-4, max=1. - 1e-4)

    pos = full[:, 4] > 0
    num_pos = int(pos.sum())
    out = {'labels': labels, 'matched': matched, 'targets': targets, 'num_pos': num_pos}
    zero = torch.tensor(0.).to(device)
    if num_pos == 0:
        out.update(cls_sum=zero, reg_sum=zero, ctr_sum=zero, cls_loss=cls_loss_weight * zero,
                   reg_loss=box_loss_weight * zero,
                   center_ness_loss=center_ness_loss_weight * zero)
        return out
    cls_sum = _focal_sum(cls, full[:, 4], alpha, gamma)
    # IoU loss (losses.py:550-586)
    dist = torch.exp(reg)[pos]
    tp = full[pos]
    pred_box = torch.cat([tp[:, 6:8] - dist[:, 0:2], tp[:, 6:8] + dist[:, 2:4]], dim=1)
    gt_box = torch.cat([tp[:, 6:8] - tp[:, 0:2], tp[:, 6:8] + tp[:, 2:4]], dim=1)
    ious = box_iou(pred_box, gt_box, box_loss_iou_type)
    reg_sum = ((1 - ious) * tp[:, 5]).sum()
    # centre-ness BCE (losses.py:588-610)
    cp = ctr[pos]
    ct = tp[:, 5:6]
    ctr_sum = (-(ct * torch.log(cp) + (1. - ct) * torch.log(1. - cp))).sum()
    out.update(cls_sum=cls_sum.detach(), reg_sum=reg_sum.detach(), ctr_sum=ctr_sum.detach())
    out['cls_loss'] = cls_loss_weight * (cls_sum / num_pos)
    out['reg_loss'] = box_loss_weight * (reg_sum / num_pos)
    out['center_ness_loss'] = center_ness_loss_weight * (ctr_sum / num_pos)
    return out


# ----------------------------------------------------------------------------------------
# NMS + selection (decode.py:26-172)
# ----------------------------------------------------------------------------------------
def nms_keep(boxes, scores, nms_type='python_nms', nms_threshold=0.5):
    """decode.py:34-104.  boxes [n,4] float32 sorted by score; returns kept positions.
    nms_type None: the query decoders' "no NMS" setting (decode.py:453, :577) keeps everything."""
    if nms_type is None:
        return np.arange(scores.shape[0])
    assert nms_type in ('torch_nms', 'python_nms', 'diou_python_nms')
    if nms_type == 'torch_nms':
        from torchvision.ops import nms
        return nms(torch.tensor(boxes), torch.tensor(scores), nms_threshold).numpy()
    wh = boxes[:, 2:4] - boxes[:, 0:2]
    areas = np.maximum(wh[:, 0] * wh[:, 1], 0)
    alive = np.arange(scores.shape[0], dtype=np.int32)
    keep = []
    while alive.shape[0] > 0:
        k = alive[0]
        keep.append(k)
        alive = alive[1:]
        if alive.shape[0] == 0:
            break
        tl = np.maximum(boxes[k, 0:2], boxes[alive, 0:2])
        br = np.minimum(boxes[k, 2:4], boxes[alive, 2:4])
        inter_wh = np.maximum(br - tl, 0)
        inter = inter_wh[:, 0] * inter_wh[:, 1]
        union = np.maximum(areas[k] + areas[alive] - inter, 1e-4)
        ious = inter / union
        if nms_type == 'diou_python_nms':
            enc_tl = np.minimum(boxes[k, 0:2], boxes[alive, 0:2])
            enc_br = np.maximum(boxes[k, 2:4], boxes[alive, 2:4])
            enc_wh = np.maximum(enc_br - enc_tl, 0)
            c2 = np.maximum((enc_wh**2).sum(axis=1), 1e-4)
            ctr_k = (boxes[k, 2:4] + boxes[k, 0:2]) / 2
            ctr_o = (boxes[alive, 2:4] + boxes[alive, 0:2]) / 2
            p2 = ((ctr_k - ctr_o)**2).sum(axis=1)
            ious = ious - p2 / c2
        alive = alive[np.where(ious < nms_threshold)[0]]
    return np.array(keep)


def select_and_nms(scores, classes, boxes, max_object_num=100, min_score_threshold=0.05,
                   topn=1000, nms_type='python_nms', nms_threshold=0.5):
    """decode.py:121-172 on [B,N] scores / classes and [B,N,4] int32 boxes.

    Returns ([scores, classes, boxes] padded like the reference, extras) where extras holds per
    image 'order' (flat anchor indices of the top-n in sorted order) and 'keep' (positions in
    that list that survive NMS, before the max_object_num cut).  Ties in score are ordered by
    ascending flat index (the reference's argsort order is undefined on ties)."""
    batch = scores.shape[0]
    out_scores = np.ones((batch, max_object_num), dtype=np.float32) * (-1)
    out_classes = np.ones((batch, max_object_num), dtype=np.float32) * (-1)
    out_boxes = np.zeros((batch, max_object_num, 4), dtype=np.float32)
    extras = []
    for i in range(batch):
        cand = np.nonzero(scores[i] > min_score_threshold)[0]
        s = scores[i][cand].astype(np.float32)
        c = classes[i][cand].astype(np.float32)
        b = boxes[i][cand].astype(np.float32)
        info = {'order': np.zeros((0,), np.int64), 'keep': np.zeros((0,), np.int64)}
        if s.shape[0] != 0:
            order = np.argsort(-s, kind='stable')
            if topn < order.shape[0]:
                order = order[0:topn]
            s, c, b = s[order], c[order], b[order]
            keep = nms_keep(b, s, nms_type, nms_threshold)
            info = {'order': cand[order].astype(np.int64), 'keep': np.asarray(keep, np.int64)}
            n = min(max_object_num, keep.shape[0])
            out_scores[i, 0:n] = s[keep][0:n]
            out_classes[i, 0:n] = c[keep][0:n]
            out_boxes[i, 0:n, :] = b[keep][0:n, :]
        extras.append(info)
    return [out_scores, out_classes, out_boxes], extras


def query_decode(cls_logits, reg_preds, scaled_sizes, activation, num_classes=None,
                 max_object_num=100, min_score_threshold=0.05, topn=100, nms_type=None,
                 nms_threshold=0.5, prob_fn=None):
    """DETRDecoder.__call__ (decode.py:388-470, activation='softmax', rows whose arg-max is the
    no-object channel >= num_classes dropped) / DINODETRDecoder.__call__ (:512-592,
    activation='sigmoid', num_classes None).  cls_logits [B,Q,C] and reg_preds [B,Q,4]
    (normalised cx,cy,w,h) are torch tensors; the activation runs in torch on their device like
    the reference's, everything after it in NumPy.  prob_fn overrides the activation (the GPU
    parity tests pass torch's CUDA softmax / sigmoid, which is what a CUDA run of the reference
    computes)."""
    if prob_fn is not None:
        probs = prob_fn(cls_logits)
    elif activation == 'softmax':
        probs = torch.nn.functional.softmax(cls_logits, dim=2)
    else:
        probs = torch.sigmoid(cls_logits.float())
    probs = probs.cpu().detach().numpy()
    reg = reg_preds.cpu().detach().numpy()
    classes = np.argmax(probs, axis=2)
    scores = np.take_along_axis(probs, classes[:, :, None], axis=2)[:, :, 0]
    boxes = []
    for idx in range(reg.shape[0]):
        cx, cy, w, h = reg[idx][:, 0], reg[idx][:, 1], reg[idx][:, 2], reg[idx][:, 3]
        b = np.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], axis=1)
        ih, iw = scaled_sizes[idx][0], scaled_sizes[idx][1]
        boxes.append((b * np.array([[iw, ih, iw, ih]], dtype=np.float32))[None])
    boxes = np.concatenate(boxes, axis=0)
    if num_classes is not None:
        # decode.py:423-431: dropped rows never reach the threshold / sort; a score of -inf keeps
        # them out of select_and_nms in the same way
        scores = np.where(classes < num_classes, scores, -np.inf).astype(np.float32)
    result, extras = select_and_nms(scores, classes, boxes, max_object_num, min_score_threshold,
                                    topn, nms_type, nms_threshold)
    return result, {'per_image': extras, 'scores': scores, 'classes': classes, 'boxes': boxes}


def _to_np_rows(level_tensors, keep_half=False):
    """decode.py:208-219: `.cpu().detach().numpy()` keeps the tensor's dtype, so a float16 head stays
    float16 in NumPy (keep_half).  bfloat16 has no NumPy dtype -- the reference raises there -- and is
    upcast."""
    def one(t):
        t = t.cpu().detach()
        if not (keep_half and t.dtype == torch.float16):
            t = t.float()
        return t.numpy().reshape(t.shape[0], -1, t.shape[-1])
    return np.concatenate([one(t) for t in level_tensors], axis=1)


def _decoder_exp(reg, exp_fn):
    """np.exp as the decoders see it (decode.py:260, :356): float32 arrays go through NumPy's SIMD
    float32 kernel (restated in oracle/npexp.c); float16 arrays through NumPy's half loop (float
    conversion, the C library's expf, rounding to half) -- the result IS float16 and only the
    following multiply / subtract promotes to float32."""
    if reg.dtype == np.float16:
        with np.errstate(over='ignore'):
            return np.exp(reg)
    return exp_fn(reg)


def retina_decode(preds, areas, ratios, scales, strides, max_object_num=100,
                  min_score_threshold=0.05, topn=1000, nms_type='python_nms',
                  nms_threshold=0.5, exp_fn=None, level_anchors=None):
    """decode.py:201-271.  exp_fn defaults to the pinned NumPy-exp restatement."""
    exp_fn = exp_fn or np_exp_f32
    cls_levels, reg_levels = preds
    if level_anchors is None:
        level_anchors = retina_anchors(feature_sizes_of(cls_levels), areas, ratios, scales,
                                       strides)
    cls = _to_np_rows(cls_levels)
    reg = _to_np_rows(reg_levels, keep_half=True)
    anchors = np.concatenate([a.reshape(-1, 4) for a in level_anchors], axis=0)[None]
    classes = np.argmax(cls, axis=2)
    scores = np.take_along_axis(cls, classes[:, :, None], axis=2)[:, :, 0]
    a_wh = anchors[:, :, 2:4] - anchors[:, :, 0:2]
    a_ctr = anchors[:, :, 0:2] + 0.5 * a_wh
    wh = _decoder_exp(reg[:, :, 2:4], exp_fn) * a_wh
    ctr = reg[:, :, :2] * a_wh + a_ctr
    boxes = np.concatenate([ctr - 0.5 * wh, ctr + 0.5 * wh], axis=2)
    with np.errstate(invalid='ignore'):
        boxes = boxes.astype(np.int32)
    result, extras = select_and_nms(scores, classes, boxes, max_object_num,
                                    min_score_threshold, topn, nms_type, nms_threshold)
    return result, {'per_image': extras, 'scores': scores, 'classes': classes, 'boxes': boxes}


def fcos_decode(preds, strides, max_object_num=100, min_score_threshold=0.05, topn=1000,
                nms_type='python_nms', nms_threshold=0.6, exp_fn=None):
    """decode.py:293-364."""
    exp_fn = exp_fn or np_exp_f32
    cls_levels, reg_levels, ctr_levels = preds
    positions = fcos_positions(feature_sizes_of(cls_levels), strides)
    cls = _to_np_rows(cls_levels)
    reg = _to_np_rows(reg_levels, keep_half=True)
    ctr = _to_np_rows(ctr_levels)
    pts = np.concatenate([p.reshape(-1, 2) for p in positions], axis=0)[None]
    classes = np.argmax(cls, axis=2)
    scores = np.take_along_axis(cls, classes[:, :, None], axis=2)[:, :, 0]
    scores = np.sqrt(scores * ctr[:, :, 0])
    dist = _decoder_exp(reg, exp_fn)
    boxes = np.concatenate([pts - dist[:, :, 0:2], pts + dist[:, :, 2:4]], axis=2)
    with np.errstate(invalid='ignore'):
        boxes = boxes.astype(np.int32)
    result, extras = select_and_nms(scores, classes, boxes, max_object_num,
                                    min_score_threshold, topn, nms_type, nms_threshold)
    return result, {'per_image': extras, 'scores': scores, 'classes': classes, 'boxes': boxes}


# ----------------------------------------------------------------------------------------
# RetinaFace (simpleAICV/face_detection/{models/anchor,losses,decode}.py) -- SURVEY section 8f-2
# ----------------------------------------------------------------------------------------
def retinaface_anchors(feature_sizes, anchor_sizes, strides):
    """face_detection/models/anchor.py:15-88: square anchors, one per size and location."""
    out = []
    for sizes, (fw, fh), stride in zip(anchor_sizes, feature_sizes, strides):
        wh = np.array([[s, s] for s in sizes], dtype=np.float32)
        base = np.zeros((len(sizes), 4), dtype=np.float32)
        base[:, 2:] += wh
        base[:, 0] -= base[:, 2] / 2
        base[:, 1] -= base[:, 3] / 2
        base[:, 2] /= 2
        base[:, 3] /= 2
        stride = np.float32(stride)
        sx = ((np.arange(0, fw) + 0.5) * stride).astype(np.float32)
        sy = ((np.arange(0, fh) + 0.5) * stride).astype(np.float32)
        shifts = np.empty((fh, fw, 1, 4), dtype=np.float32)
        shifts[:, :, 0, 0] = sx[None, :]
        shifts[:, :, 0, 1] = sy[:, None]
        shifts[:, :, 0, 2] = sx[None, :]
        shifts[:, :, 0, 3] = sy[:, None]
        out.append(np.ascontiguousarray(base[None, None, :, :] + shifts, dtype=np.float32))
    return out


def retinaface_loss(preds, annotations, anchor_sizes, strides, alpha=0.25, gamma=2,
                    beta=1.0 / 9.0, cls_loss_weight=1., box_loss_weight=1., box_loss_type='CIoU'):
    """face_detection/losses.py:54-292: RetinaLoss's arithmetic with square anchors and the
    0.35 / 0.35 assignment thresholds (:255-259)."""
    anchors = retinaface_anchors(feature_sizes_of(preds[0]), anchor_sizes, strides)
    return retina_loss(preds, annotations, None, None, None, strides, alpha, gamma, beta,
                       cls_loss_weight, box_loss_weight, box_loss_type, level_anchors=anchors,
                       neg_thr=0.35, pos_thr=0.35)


def retinaface_decode(preds, anchor_sizes, strides, max_object_num=100, min_score_threshold=0.3,
                      topn=1000, nms_type='python_nms', nms_threshold=0.3, exp_fn=None):
    """face_detection/decode.py:47-117."""
    anchors = retinaface_anchors(feature_sizes_of(preds[0]), anchor_sizes, strides)
    return retina_decode(preds, None, None, None, strides, max_object_num, min_score_threshold,
                         topn, nms_type, nms_threshold, exp_fn, level_anchors=anchors)


# ----------------------------------------------------------------------------------------
# Head tail (models/head.py:46-50, :176-179; models/retinanet.py:73-76; models/fcos.py:70-79)
# ----------------------------------------------------------------------------------------
def head_tail(x, num_classes=None):
    """`x.float()` -> sigmoid -> permute(0, 2, 3, 1).contiguous() (-> view [B,H,W,A,C] for
    RetinaNet).  Runs on x's device; differentiable through torch autograd."""
    y = torch.sigmoid(x.float())
    y = y.permute(0, 2, 3, 1).contiguous()
    if num_classes is not None:
        y = y.view(y.shape[0], y.shape[1], y.shape[2], -1, num_classes)
    return y


# ----------------------------------------------------------------------------------------
# VOC evaluator (tools/scripts.py:455-508, 592-684)
# ----------------------------------------------------------------------------------------
def compute_ious(a, b):
    """tools/scripts.py:487-508: [N,4] x [M,4] -> [N,M]; no clamps: degenerate pairs give NaN / inf."""
    a = np.expand_dims(a, axis=1)
    b = np.expand_dims(b, axis=0)
    overlap = np.maximum(0.0, np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]))
    overlap = np.prod(overlap, axis=-1)
    area_a = np.prod(a[..., 2:] - a[..., :2], axis=-1)
    area_b = np.prod(b[..., 2:] - b[..., :2], axis=-1)
    with np.errstate(invalid='ignore', divide='ignore'):
        return overlap / (area_a + area_b - overlap)


def compute_voc_ap(recall, precision, use_07_metric=False):
    """tools/scripts.py:455-484."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(recall >= t) == 0 else np.max(precision[recall >= t])
            ap = ap + p / 11.
        return ap
    mrecall = np.concatenate(([0.], recall, [1.]))
    mprecision = np.concatenate(([0.], precision, [0.]))
    for i in range(mprecision.size - 1, 0, -1):
        mprecision[i - 1] = np.maximum(mprecision[i - 1], mprecision[i])
    i = np.where(mrecall[1:] != mrecall[:-1])[0]
    return np.sum((mrecall[i + 1] - mrecall[i]) * mprecision[i + 1])


def voc_match(pred_classes, pred_boxes, gt_boxes, gt_classes, iou_threshold):
    """The per-image matching of tools/scripts.py:626-651 for ONE image and threshold, over all
    classes at once: detection d (in the decoder's order) is a true positive iff the ground-truth
    box of ITS class with the largest IoU (np.argmax: first maximum, NaN counts as maximum) has
    IoU >= threshold and was not taken by an earlier detection.  Returns bool [n_det]."""
    tp = np.zeros((len(pred_classes),), dtype=bool)
    assigned = {}
    for d in range(len(pred_classes)):
        c = pred_classes[d]
        idx = np.nonzero(gt_classes == c)[0]
        if idx.shape[0] == 0:
            continue
        iou = compute_ious(gt_boxes[idx], np.expand_dims(pred_boxes[d], axis=0))
        g = np.argmax(iou, axis=0)
        if iou[g, 0] >= iou_threshold and int(g[0]) not in assigned.setdefault(c, []):
            tp[d] = True
            assigned[c].append(int(g[0]))
    return tp


def voc_map(preds, gts, iou_thresholds, num_classes, tp_flags=None):
    """tools/scripts.py:614-684.  preds[i] = [boxes, classes, scores], gts[i] = [boxes, classes] per
    image.  tp_flags[t][i] (bool per detection) may come from elsewhere (the CUDA matcher); default:
    voc_match.  Returns ({key: mAP}, {key: [AP per class]}) with the reference's keys."""
    maps, per_class = {}, {}
    for t, thr in enumerate(iou_thresholds):
        aps = []
        for c in range(num_classes):
            tps, scores, total_gts = [], [], 0
            for i, (p, g) in enumerate(zip(preds, gts)):
                total_gts += int(np.sum(g[1] == c))
                flags = tp_flags[t][i] if tp_flags is not None else voc_match(p[1], p[0], g[0], g[1], thr)
                sel = p[1] == c
                tps.append(np.asarray(flags)[sel])
                scores.append(p[2][sel])
            tp = np.concatenate(tps).astype(np.float64) if tps else np.zeros((0,))
            sc = np.concatenate(scores).astype(np.float64) if scores else np.zeros((0,))
            fp = 1.0 - tp
            order = np.argsort(-sc)
            fp, tp = np.cumsum(fp[order]), np.cumsum(tp[order])
            with np.errstate(invalid='ignore', divide='ignore'):
                recall = tp / total_gts
            precision = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
            aps.append(compute_voc_ap(recall, precision, use_07_metric=False) * 100)
        m = 0.
        for ap in aps:
            m += float(ap)
        maps[f'IoU={thr:.2f},area=all,maxDets=100,mAP'] = m / num_classes
        per_class[f'IoU={thr:.2f},area=all,maxDets=100,per_class_ap'] = aps
    return maps, per_class
