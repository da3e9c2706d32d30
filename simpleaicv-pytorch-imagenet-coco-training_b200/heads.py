"""Tail of the detection heads: sigmoid + NCHW -> NHWC in one kernel.

Reference ops it replaces, in this order:
    x = x.float(); x = self.sigmoid(x)                  models/head.py:46-50 (RetinaClsHead.forward),
                                                        :176-179 (FCOSClsRegCntHead.forward)
    x = x.permute(0, 2, 3, 1).contiguous()              models/retinanet.py:73-74, models/fcos.py:70-79
    x = x.view(B, H, W, -1, num_classes)                models/retinanet.py:75-76 (RetinaNet only)

torch runs that as two full passes over the largest tensor of the detector (16 B per element);
`sigmoid_channels_last` does it in one (8 B per element, 6 B for fp16 / bf16 conv outputs) and
its backward in one more.  The values are bit-identical to running the reference's ops on the GPU.
CUDA tensors only; there is no fallback.
"""
import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream():
    return _lib.raw_stream()


class _SigmoidPermute(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda:
            raise RuntimeError('b200det.heads needs CUDA tensors (there is no CPU fallback)')
        if x.dim() != 4 or x.dtype not in _DTYPES:
            raise RuntimeError('expected a [B, C, H, W] float32 / float16 / bfloat16 tensor')
        x = x.contiguous()
        b, c, h, w = x.shape
        out = torch.empty((b, h, w, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(
                _lib.load().b200det_head_sigmoid_permute(x.data_ptr(), _DTYPES[x.dtype], b, c,
                                                         h * w, out.data_ptr(), _stream()),
                'b200det_head_sigmoid_permute')
        ctx.save_for_backward(out)
        ctx.in_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        b, h, w, c = out.shape
        grad_out = grad_out.contiguous().float()
        grad_in = torch.empty((b, c, h, w), dtype=ctx.in_dtype, device=out.device)
        with torch.cuda.device(out.device):
            _lib.check(
                _lib.load().b200det_head_sigmoid_permute_backward(
                    grad_out.data_ptr(), out.data_ptr(), b, c, h * w, grad_in.data_ptr(),
                    _DTYPES[ctx.in_dtype], _stream()),
                'b200det_head_sigmoid_permute_backward')
        return grad_in


def sigmoid_channels_last(x, num_classes=None):
    """[B, C, H, W] convolution output -> float32 probabilities [B, H, W, C]; with `num_classes`
    the result is viewed as [B, H, W, C // num_classes, num_classes] like RetinaNet.forward does.
    Differentiable (one backward kernel)."""
    y = _SigmoidPermute.apply(x)
    if num_classes is not None:
        y = y.view(y.shape[0], y.shape[1], y.shape[2], -1, num_classes)
    return y


class SigmoidChannelsLast(torch.nn.Module):
    """nn.Module form: put it where the head's `self.sigmoid` was and drop the caller's
    `permute(0, 2, 3, 1).contiguous()`."""

    def __init__(self, num_classes=None):
        super().__init__()
        self.num_classes = num_classes

    def forward(self, x):
        return sigmoid_channels_last(x, self.num_classes)
