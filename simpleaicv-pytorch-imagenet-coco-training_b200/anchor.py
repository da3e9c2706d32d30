"""b200det.anchor -- RetinaAnchors / FCOSPositions on the GPU.

Same constructors and `__call__(fpn_feature_sizes)` as simpleAICV/detection/models/anchor.py:5-130
(`fpn_feature_sizes[i] = [W, H]`, one NumPy float32 array per level: `[H, W, A, 4]` x1,y1,x2,y2 in
ratio-major / scale-minor order, or `[H, W, 2]` x,y centres).  The losses and decoders of this package
never call them -- their kernels generate `base + (x + 0.5) * stride` in registers (csrc/common.cuh:
anchor_of, point_of) instead of reading 120 087 x 16 bytes per image -- but code that wants the tables
gets them from the very same device functions (`b200det_generate_rows`), bit-identical to the
reference's NumPy loops (tests/golden/tables.npz).  CUDA only; there is no fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import geometry as _geom


def _generate(shapes, per_loc, strides, base, is_fcos, width):
    if not torch.cuda.is_available():
        raise RuntimeError('b200det.anchor needs a CUDA device (there is no CPU fallback)')
    geo = _geom.make_geometry(shapes, 1, per_loc, 1, strides, base_anchors=base)
    n = _geom.rows_per_image(shapes, per_loc)
    out = torch.empty(n * width, dtype=torch.float32, device='cuda')
    _lib.check(
        _lib.load().b200det_generate_rows(ctypes.byref(geo), int(is_fcos), out.data_ptr(),
                                          _lib.raw_stream()), 'b200det_generate_rows')
    flat = out.cpu().numpy()
    levels, o = [], 0
    for h, w in shapes:
        k = h * w * per_loc * width
        shape = (h, w, per_loc, width) if not is_fcos else (h, w, width)
        levels.append(flat[o:o + k].reshape(shape).copy())
        o += k
    return levels


class RetinaAnchors:
    """models/anchor.py:5-86."""

    def __init__(self,
                 areas=[[32, 32], [64, 64], [128, 128], [256, 256], [512, 512]],
                 ratios=[0.5, 1, 2],
                 scales=[2**0, 2**(1.0 / 3.0), 2**(2.0 / 3.0)],
                 strides=[8, 16, 32, 64, 128]):
        self.areas = np.array(areas, dtype=np.float32)
        self.ratios = np.array(ratios, dtype=np.float32)
        self.scales = np.array(scales, dtype=np.float32)
        self.strides = np.array(strides, dtype=np.float32)
        self._base = _geom.retina_base_anchors(areas, ratios, scales)

    def __call__(self, fpn_feature_sizes):
        """generate one image's anchors: [[H, W, A, 4], ...] (anchor.py:18-33)"""
        shapes = [(int(s[1]), int(s[0])) for s in fpn_feature_sizes[:len(self.areas)]]
        return _generate(shapes, self._base.shape[1], self.strides, self._base[:len(shapes)], False, 4)

    def generate_base_anchors(self, area, scales, ratios):
        """[A, 4] anchors of one level centred on the origin (anchor.py:35-57)."""
        return _geom.retina_base_anchors([area], ratios, scales)[0]


class FCOSPositions:
    """models/anchor.py:89-130."""

    def __init__(self, strides=[8, 16, 32, 64, 128]):
        self.strides = np.array(strides, dtype=np.float32)

    def __call__(self, fpn_feature_sizes):
        """generate one image's positions: [[H, W, 2], ...] (anchor.py:94-108)"""
        n = min(len(self.strides), len(fpn_feature_sizes))
        shapes = [(int(s[1]), int(s[0])) for s in fpn_feature_sizes[:n]]
        return _generate(shapes, 1, self.strides, None, True, 2)
