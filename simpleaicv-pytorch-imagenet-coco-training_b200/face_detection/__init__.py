"""b200det.face_detection -- drop-in RetinaFaceLoss / RetinaFaceDecoder.

    from b200det.face_detection import losses, decode
    # instead of: from simpleAICV.face_detection import losses, decode
    # (10.face_detection_training/resnet50_retinaface/train_config.py:11-12)

The reference's RetinaFace loss and decoder reuse the detection path's IoUMethod and DecodeMethod
(simpleAICV/face_detection/losses.py:15, decode.py:16); here they reuse the same CUDA kernels with
square anchors and 0.35 / 0.35 assignment thresholds.
"""
from . import losses, decode  # noqa: F401
