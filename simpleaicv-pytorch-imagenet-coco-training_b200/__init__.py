"""b200det -- B200-native (sm_100a) dense-detection target assignment, loss and post-processing.

Drop-in for the `losses` and `decode` modules of zgcr/SimpleAICV-pytorch-ImageNet-COCO-training
(simpleAICV/detection/losses.py:126-833, simpleAICV/detection/decode.py:26-364):

    from b200det import losses, decode          # instead of: from simpleAICV.detection import ...
    criterion = losses.__dict__['RetinaLoss'](**kwargs)
    decoder = decode.__dict__['RetinaDecoder'](**kwargs)

The numerics run in hand-written CUDA kernels (csrc/*.cu) behind a C ABI (include/b200det.h)
loaded with ctypes; PyTorch only provides device memory, streams and torch.distributed.
"""
from . import _build, _lib, geometry  # noqa: F401
from . import losses, decode, fused, heads, evaluation, anchor  # noqa: F401
from .losses import RetinaLoss, FCOSLoss  # noqa: F401
from .decode import RetinaDecoder, FCOSDecoder  # noqa: F401

__version__ = '0.1.0'
