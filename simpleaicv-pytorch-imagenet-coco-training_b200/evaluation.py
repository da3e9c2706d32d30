"""Per-batch part of the reference's VOC evaluator on the GPU (second half).

Reference it replaces:
    compute_ious(a, b)                                  tools/scripts.py:487-508
    the matching loop of evaluate_voc_detection         tools/scripts.py:626-651
    (per threshold, class, image, detection: arg-max IoU over the ground truth of the detection's
    class, true positive iff IoU >= threshold and that box has not been taken)
The AP integration (compute_voc_ap, tools/scripts.py:455-484) and the per-class cumulative sums
(:653-668) are a few hundred float64 operations per class over the whole test set; they stay host
NumPy here, with the reference's op order so the mAP is bit-identical.

`voc_match` takes what the reference's eval loop already holds per image (`preds[i] = [boxes,
classes, scores]`, `gts[i] = [boxes, classes]`, scripts.py:566-590), pads it to one batch and makes
ONE kernel launch for all images and thresholds.  CUDA only; there is no fallback.
"""
import numpy as np
import torch

from . import _lib


def _device(device):
    if not torch.cuda.is_available():
        raise RuntimeError('b200det.evaluation needs a CUDA device (there is no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def compute_ious(a, b, device=None):
    """tools/scripts.py:487-508 on the GPU: float32 [N,4] x [M,4] -> float32 [N,M], same bits as the
    NumPy expression (no clamps: degenerate pairs give NaN / inf).  Accepts NumPy arrays or tensors;
    returns the type it was given."""
    as_numpy = not torch.is_tensor(a)
    device = _device(device if as_numpy else a.device)
    ta = torch.as_tensor(a, dtype=torch.float32).reshape(-1, 4).to(device).contiguous()
    tb = torch.as_tensor(b, dtype=torch.float32).reshape(-1, 4).to(device).contiguous()
    n, m = ta.shape[0], tb.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=device)
    if n and m:
        with torch.cuda.device(device):
            _lib.check(
                _lib.load().b200det_pair_ious(ta.data_ptr(), n, tb.data_ptr(), m, out.data_ptr(),
                                              _lib.raw_stream(device)), 'b200det_pair_ious')
    return out.cpu().numpy() if as_numpy else out


def _pad(rows, width, cols, fill):
    out = np.full((len(rows), max(width, 1)) + ((cols,) if cols else ()), fill, dtype=np.float32)
    for i, r in enumerate(rows):
        r = np.asarray(r, dtype=np.float32)
        if r.shape[0]:
            out[i, :r.shape[0]] = r.reshape((r.shape[0],) + ((cols,) if cols else ()))
    return out


def voc_match_batch(pred_boxes, pred_classes, gt_boxes, gt_classes, thresholds):
    """Padded form: device float32 pred_boxes [B,M,4], pred_classes [B,M], gt_boxes [B,G,4],
    gt_classes [B,G] (class <= -1 = padding, as the decoder and the collater pad), thresholds
    sequence -> uint8 tensor [T,B,M] of true-positive flags.  One launch, no host sync."""
    if not pred_boxes.is_cuda:
        raise RuntimeError('b200det.evaluation needs CUDA tensors (there is no CPU fallback)')
    device = pred_boxes.device
    pred_boxes = pred_boxes.float().contiguous()
    pred_classes = pred_classes.float().contiguous()
    gt_boxes = gt_boxes.float().contiguous()
    gt_classes = gt_classes.float().contiguous()
    batch, max_det = pred_classes.shape
    max_gt = gt_classes.shape[1]
    thr = torch.as_tensor(np.asarray(thresholds, dtype=np.float32)).to(device)
    tp = torch.empty((thr.shape[0], batch, max_det), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(
            _lib.load().b200det_voc_match(pred_boxes.data_ptr(), pred_classes.data_ptr(), max_det,
                                          gt_boxes.data_ptr(), gt_classes.data_ptr(), max_gt,
                                          thr.data_ptr(), thr.shape[0], batch, tp.data_ptr(),
                                          _lib.raw_stream(device)), 'b200det_voc_match')
    return tp


def voc_match(preds, gts, thresholds, device=None):
    """preds[i] = [boxes [n_i,4], classes [n_i], scores [n_i]] in the decoder's (descending score)
    order, gts[i] = [boxes [g_i,4], classes [g_i]] (tools/scripts.py:566-590).  Returns
    flags[t][i] = bool [n_i]: detection is a true positive at thresholds[t]."""
    device = _device(device)
    if len(preds) == 0:
        return [[] for _ in thresholds]
    max_det = max(len(p[1]) for p in preds)
    max_gt = max(len(g[1]) for g in gts)
    pb = _pad([p[0] for p in preds], max_det, 4, 0.)
    pc = _pad([p[1] for p in preds], max_det, 0, -1.)
    gb = _pad([g[0] for g in gts], max_gt, 4, 0.)
    gc = _pad([g[1] for g in gts], max_gt, 0, -1.)
    tp = voc_match_batch(*(torch.from_numpy(x).to(device) for x in (pb, pc, gb, gc)),
                         thresholds).cpu().numpy().astype(bool)
    return [[tp[t, i, :len(p[1])] for i, p in enumerate(preds)] for t in range(len(thresholds))]


def compute_voc_ap(recall, precision, use_07_metric=False):
    """tools/scripts.py:455-484 (host NumPy, float64)."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(recall >= t) == 0 else np.max(precision[recall >= t])
            ap = ap + p / 11.
        return ap
    mrec = np.concatenate(([0.], recall, [1.]))
    mpre = np.concatenate(([0.], precision, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def voc_map(preds, gts, iou_thresholds, num_classes, device=None):
    """The result dictionaries of evaluate_voc_detection (tools/scripts.py:614-684): matching on the
    GPU (one launch), AP / mAP in float64 NumPy with the reference's op order and keys.  Returns
    (all_iou_threshold_map, all_iou_threshold_per_class_ap)."""
    flags = voc_match(preds, gts, iou_thresholds, device)
    maps, per_class = {}, {}
    for t, thr in enumerate(iou_thresholds):
        aps = []
        for c in range(num_classes):
            tps, scores, total_gts = [np.zeros((0,))], [np.zeros((0,))], 0
            for i, (p, g) in enumerate(zip(preds, gts)):
                total_gts += int(np.sum(np.asarray(g[1]) == c))
                sel = np.asarray(p[1]) == c
                tps.append(flags[t][i][sel].astype(np.float64))
                scores.append(np.asarray(p[2])[sel].astype(np.float64))
            tp, sc = np.concatenate(tps), np.concatenate(scores)
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
