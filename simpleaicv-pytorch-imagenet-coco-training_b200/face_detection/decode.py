"""Drop-in for simpleAICV/face_detection/decode.py:24-117 (RetinaFaceDecoder)."""
from .. import geometry as _geom
from ..decode import _DecoderBase
from .losses import square_base_anchors

__all__ = ['RetinaFaceDecoder']


class RetinaFaceDecoder(_DecoderBase):
    """Same constructor and __call__(preds) -> [scores, classes, boxes] as the reference; the
    arithmetic is RetinaDecoder's (decode.py:201-271) on square anchors."""

    _is_fcos = False

    def __init__(self,
                 anchor_sizes=[[8, 16, 32], [32, 64, 128], [128, 256, 512]],
                 strides=[8, 16, 32],
                 max_object_num=100,
                 min_score_threshold=0.3,
                 topn=1000,
                 nms_type='python_nms',
                 nms_threshold=0.3):
        self._init_common(max_object_num, min_score_threshold, topn, nms_type, nms_threshold)
        self.anchor_sizes = anchor_sizes
        self.strides = strides
        self._per_loc = len(anchor_sizes[0])
        if any(len(s) != self._per_loc for s in anchor_sizes):
            raise ValueError('every level needs the same number of anchor sizes')
        self._base = square_base_anchors(anchor_sizes)

    def _geometry(self, shapes, batch, num_classes):
        if len(shapes) > len(self.anchor_sizes):
            raise ValueError('more pyramid levels than anchor size lists')
        return _geom.make_geometry(shapes, batch, self._per_loc, num_classes, self.strides,
                                   base_anchors=self._base)
