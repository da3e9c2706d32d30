"""Deterministic synthetic head outputs and annotations of the shapes named in BASELINE.json
(SURVEY.md section 8d).  Host-side helper shared by tests, bench.py and the golden-vector script;
contains no detection arithmetic.

Pyramid for a square input S (ResNet strides 8/16/32 then two stride-2 3x3 convs,
reference models/fpn.py:51-58): [ceil(S/8), ceil(S/16), ceil(S/32), p6, p7].
"""
import math

import numpy as np
import torch

STRIDES = [8, 16, 32, 64, 128]
AREAS = [[32, 32], [64, 64], [128, 128], [256, 256], [512, 512]]
RATIOS = [0.5, 1, 2]
SCALES = [2**0, 2**(1.0 / 3.0), 2**(2.0 / 3.0)]
MI = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, 100000000]]

RETINA_KW = dict(areas=AREAS, ratios=RATIOS, scales=SCALES, strides=STRIDES)


def pyramid_sizes(size):
    p3 = math.ceil(size / 8)
    p4 = math.ceil(size / 16)
    p5 = math.ceil(size / 32)
    p6 = (p5 - 1) // 2 + 1
    p7 = (p6 - 1) // 2 + 1
    return [p3, p4, p5, p6, p7]


def make_annotations(batch, max_gt, size, num_classes, seed=1, min_gt=1, empty_images=()):
    """float32 [B, max_gt, 5]: n ~ U{min_gt..max_gt} boxes per image, w,h log-uniform in
    [8, size/2], centre uniform in [0,size), clipped to the image, x2 >= x1+1; the remaining rows
    (and every row of `empty_images`) are -1.  Valid rows are NOT a prefix: they are scattered
    over the row slots, because the reference filters on class >= 0, not on a length."""
    rng = np.random.RandomState(seed)
    ann = np.full((batch, max_gt, 5), -1, dtype=np.float32)
    for b in range(batch):
        if b in empty_images:
            continue
        n = int(rng.randint(min_gt, max_gt + 1))
        w = np.exp(rng.uniform(math.log(8), math.log(size / 2), n))
        h = np.exp(rng.uniform(math.log(8), math.log(size / 2), n))
        cx = rng.uniform(0, size, n)
        cy = rng.uniform(0, size, n)
        x1 = np.clip(cx - w / 2, 0, size - 1)
        y1 = np.clip(cy - h / 2, 0, size - 1)
        x2 = np.clip(cx + w / 2, 0, size)
        y2 = np.clip(cy + h / 2, 0, size)
        x2 = np.maximum(x2, x1 + 1)
        y2 = np.maximum(y2, y1 + 1)
        cls = rng.randint(0, num_classes, n)
        slots = np.sort(rng.permutation(max_gt)[:n])
        ann[b, slots, 0] = x1
        ann[b, slots, 1] = y1
        ann[b, slots, 2] = x2
        ann[b, slots, 3] = y2
        ann[b, slots, 4] = cls
    return torch.from_numpy(ann)


def make_retina_preds(batch, size, num_classes, seed=0, sigma=1.0, per_loc=9, device='cpu',
                      sizes=None):
    """cls: sigmoid(N(-4.595, sigma^2)) float32 [B,H,W,A,C]; reg: N(0, 0.2^2) [B,H,W,A,4]."""
    gen = torch.Generator(device=device).manual_seed(seed)
    cls, reg = [], []
    for p in (sizes or pyramid_sizes(size)):
        c = torch.randn((batch, p, p, per_loc, num_classes), generator=gen, device=device)
        cls.append(torch.sigmoid(c * sigma - 4.595))
        reg.append(torch.randn((batch, p, p, per_loc, 4), generator=gen, device=device) * 0.2)
    return [cls, reg]


def make_fcos_preds(batch, size, num_classes, seed=0, sigma=1.0, device='cpu', sizes=None):
    """cls: sigmoid(N(-4.595, sigma^2)) [B,H,W,C]; reg level l: N(log(8*2^l), 0.5^2) [B,H,W,4];
    centre-ness: sigmoid(N(0,1)) [B,H,W,1]."""
    gen = torch.Generator(device=device).manual_seed(seed)
    cls, reg, ctr = [], [], []
    for l, p in enumerate(sizes or pyramid_sizes(size)):
        c = torch.randn((batch, p, p, num_classes), generator=gen, device=device)
        cls.append(torch.sigmoid(c * sigma - 4.595))
        r = torch.randn((batch, p, p, 4), generator=gen, device=device)
        reg.append(r * 0.5 + math.log(8 * 2**l))
        ctr.append(torch.sigmoid(torch.randn((batch, p, p, 1), generator=gen, device=device)))
    return [cls, reg, ctr]


def make_retina_preds_sharded(first, count, size, num_classes, seed=0, sigma=1.0, per_loc=9,
                              device='cpu'):
    """Images [first, first + count) of ONE global synthetic batch: every image has its own seed, so
    the union over any sharding (1, 2, 4, 8 ranks ...) is the same batch, bit for bit.  Same
    distributions as make_retina_preds."""
    sizes = pyramid_sizes(size)
    cls = [torch.empty((count, p, p, per_loc, num_classes), device=device) for p in sizes]
    reg = [torch.empty((count, p, p, per_loc, 4), device=device) for p in sizes]
    gen = torch.Generator(device=device)
    for i in range(count):
        gen.manual_seed(1000003 * (seed + 1) + first + i)
        for l, p in enumerate(sizes):
            c = torch.randn((p, p, per_loc, num_classes), generator=gen, device=device)
            cls[l][i] = torch.sigmoid(c * sigma - 4.595)
            reg[l][i] = torch.randn((p, p, per_loc, 4), generator=gen, device=device) * 0.2
    return [cls, reg]


def make_annotations_sharded(first, count, max_gt, size, num_classes, seed=1):
    """Rows [first, first + count) of the global batch's annotations (per-image seeds)."""
    return torch.cat([make_annotations(1, max_gt, size, num_classes, seed=1000003 * (seed + 1) + first + i)
                      for i in range(count)], dim=0)


def make_tie_free(preds, min_score=0.05, max_rounds=64):
    """Bumps the arg-max class probability of rows whose FINAL score (max prob, or
    sqrt(max prob * centre-ness) for FCOS) collides with another row of the same image, until
    every score above `min_score` is unique per image.  The reference sorts with an unstable
    argsort (decode.py:142), so only tie-free inputs have a defined top-n order.  In place."""
    cls_levels = preds[0]
    ctr_levels = preds[2] if len(preds) == 3 else None
    batch = cls_levels[0].shape[0]
    num_classes = cls_levels[0].shape[-1]
    flat = [c.view(batch, -1, num_classes) for c in cls_levels]  # views: edits hit the inputs
    for b in range(batch):
        for _ in range(max_rounds):
            cls = np.concatenate([f[b].numpy() for f in flat], axis=0)
            arg = cls.argmax(axis=1)
            score = cls[np.arange(cls.shape[0]), arg]
            if ctr_levels is not None:
                ctr = np.concatenate([c[b].reshape(-1).numpy() for c in ctr_levels], axis=0)
                score = np.sqrt(score * ctr)
            cand = np.nonzero(score > np.float32(min_score))[0]
            order = cand[np.argsort(score[cand], kind='stable')]
            dup = order[1:][score[order[1:]] == score[order[:-1]]]
            if dup.size == 0:
                break
            start = 0
            for f in flat:
                n = f.shape[1]
                sel = dup[(dup >= start) & (dup < start + n)] - start
                if sel.size:
                    rows = torch.from_numpy(sel)
                    cols = torch.from_numpy(arg[sel + start])
                    vals = f[b][rows, cols].numpy()
                    f[b][rows, cols] = torch.from_numpy(
                        np.nextafter(vals, np.float32(2), dtype=np.float32))
                start += n
        else:
            raise RuntimeError('could not make the scores tie-free')
    return preds


def to_device(preds, device):
    return [[t.to(device) for t in group] for group in preds]
