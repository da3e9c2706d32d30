"""Host-side pyramid geometry: the tiny tables the kernels need instead of materialised anchors.

Mirrors the *inputs* of the reference's RetinaAnchors / FCOSPositions
(simpleAICV/detection/models/anchor.py:5-130): the reference rebuilds ~120k anchors in Python
loops on every forward and copies them to the device (losses.py:172-179, decode.py:206);
here only the 9 base anchors per level (45 x 4 floats) are computed on the host, once per
constructor, and the kernels generate `base + (x+0.5)*stride` in registers.
"""
import math

import numpy as np

from . import _lib


def retina_base_anchors(areas, ratios, scales):
    """float32 [n_levels, len(ratios)*len(scales), 4] (x1,y1,x2,y2) centred on the origin.

    Same arithmetic as anchor.py:35-57: aspects are float32(scale) * float64 sqrt(ratio)
    rounded to float32 (NEP-50 weak python scalars), sizes multiply in float32, then the box
    is [-w/2, -h/2, w/2, h/2]; ratio-major / scale-minor order."""
    ratios32 = np.asarray(ratios, dtype=np.float32)
    scales32 = np.asarray(scales, dtype=np.float32)
    aspects = np.empty((len(ratios32) * len(scales32), 2), dtype=np.float32)
    k = 0
    for r in ratios32:
        for s in scales32:
            aspects[k, 0] = np.float32(s * math.sqrt(r))
            aspects[k, 1] = np.float32(s * math.sqrt(1 / r))
            k += 1
    out = np.zeros((len(areas), aspects.shape[0], 4), dtype=np.float32)
    for l, size in enumerate(areas):
        wh = np.asarray(size, dtype=np.float32) * aspects
        half = wh / np.float32(2)
        out[l, :, 0] = np.float32(0) - half[:, 0]
        out[l, :, 1] = np.float32(0) - half[:, 1]
        out[l, :, 2] = half[:, 0]
        out[l, :, 3] = half[:, 1]
    return out


def level_shapes(level_tensors):
    """[(H, W), ...] of channels-last head outputs [B,H,W,...]; the reference reads
    shape[2] (W) and shape[1] (H) the same way (losses.py:169-171, decode.py:203-205)."""
    return [(int(t.shape[1]), int(t.shape[2])) for t in level_tensors]


def make_geometry(shapes, batch, per_loc, num_classes, strides, base_anchors=None, mi=None,
                  center_sample_radius=None):
    """Fills a `_lib.Geometry` (b200det_geometry) for the given per-level (H, W) shapes."""
    n = len(shapes)
    if n < 1 or n > _lib.MAX_LEVELS:
        raise ValueError(f'b200det supports 1..{_lib.MAX_LEVELS} pyramid levels, got {n}')
    if len(strides) < n:
        raise ValueError('fewer strides than pyramid levels')
    if per_loc > _lib.MAX_PER_LOC:
        raise ValueError(f'at most {_lib.MAX_PER_LOC} anchors per location')
    g = _lib.Geometry()
    g.n_levels = n
    g.batch = int(batch)
    g.per_loc = int(per_loc)
    g.num_classes = int(num_classes)
    for l, (h, w) in enumerate(shapes):
        g.height[l] = h
        g.width[l] = w
        g.stride[l] = float(np.float32(strides[l]))
        if base_anchors is not None:
            for a in range(per_loc):
                for k in range(4):
                    g.base_anchors[l][a][k] = float(base_anchors[l, a, k])
        if mi is not None:
            g.mi_lo[l] = float(np.float32(mi[l][0]))
            g.mi_hi[l] = float(np.float32(mi[l][1]))
        if center_sample_radius is not None:
            # per_image_stride * self.center_sample_radius in float32 (losses.py:690)
            g.radius[l] = float(np.float32(strides[l]) * np.float32(center_sample_radius))
    return g


def rows_per_image(shapes, per_loc):
    return sum(h * w * per_loc for h, w in shapes)
