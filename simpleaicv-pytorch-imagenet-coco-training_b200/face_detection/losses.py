"""Drop-in for simpleAICV/face_detection/losses.py:23-325 (RetinaFaceLoss)."""
import numpy as np

from .. import geometry as _geom
from ..losses import RetinaLoss

__all__ = ['RetinaFaceLoss']


def square_base_anchors(anchor_sizes):
    """float32 [levels, sizes, 4]: [-s/2, -s/2, s/2, s/2] per anchor size, as
    RetinaFaceAnchors.generate_base_anchors builds them
    (simpleAICV/face_detection/models/anchor.py:34-55)."""
    sizes = np.asarray(anchor_sizes, dtype=np.float32)
    half = sizes / np.float32(2)
    out = np.zeros(sizes.shape + (4,), dtype=np.float32)
    out[..., 0] = np.float32(0) - half
    out[..., 1] = np.float32(0) - half
    out[..., 2] = half
    out[..., 3] = half
    return out


class RetinaFaceLoss(RetinaLoss):
    """Same constructor and forward(preds, annotations) -> {'cls_loss', 'reg_loss'} as the
    reference.  Differences from RetinaLoss: square anchors (one per size and location) and the
    assignment thresholds of face_detection/losses.py:255-259: IoU < 0.35 background,
    IoU >= 0.35 positive (no ignore band)."""

    _iou_thresholds = (0.35, 0.35)

    def __init__(self,
                 anchor_sizes=[[8, 16, 32], [32, 64, 128], [128, 256, 512]],
                 strides=[8, 16, 32],
                 alpha=0.25,
                 gamma=2,
                 beta=1.0 / 9.0,
                 cls_loss_weight=1.,
                 box_loss_weight=1.,
                 box_loss_type='CIoU',
                 *,
                 sync_normalizer=False,
                 process_group=None):
        super(RetinaFaceLoss, self).__init__(areas=[[s[0], s[0]] for s in anchor_sizes],
                                             ratios=[1],
                                             scales=[1],
                                             strides=strides,
                                             alpha=alpha,
                                             gamma=gamma,
                                             beta=beta,
                                             cls_loss_weight=cls_loss_weight,
                                             box_loss_weight=box_loss_weight,
                                             box_loss_type=box_loss_type,
                                             sync_normalizer=sync_normalizer,
                                             process_group=process_group)
        self.anchor_sizes = anchor_sizes
        self._per_loc = len(anchor_sizes[0])
        if any(len(s) != self._per_loc for s in anchor_sizes):
            raise ValueError('every level needs the same number of anchor sizes')
        self._base = square_base_anchors(anchor_sizes)

    def _geometry(self, shapes, batch, num_classes):
        if len(shapes) > len(self.anchor_sizes):
            raise ValueError('more pyramid levels than anchor size lists')
        return _geom.make_geometry(shapes, batch, self._per_loc, num_classes, self.strides,
                                   base_anchors=self._base)
