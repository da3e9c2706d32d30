"""GPU parity of the head tail: one-kernel sigmoid + NCHW->NHWC against
  (a) the golden vectors made by the reference's RetinaClsHead on the CPU (<= 2 ulp: torch-CPU's
      vectorised exp and CUDA's expf round differently), and
  (b) the reference's own ops (`x.float()`, sigmoid, permute, contiguous, view) executed by torch on
      the same GPU -- the device the reference runs its heads on -- bit for bit, forward and backward.
"""
import numpy as np
import pytest
import torch

from b200det import heads
from oracle import det_oracle as O

import golden_util as G

pytestmark = pytest.mark.gpu

ULP_VS_CPU = 2   # float32 ulps between torch-CPU sigmoid and 1 / (1 + expf(-x)) on the GPU


def ulp_diff(a, b):
    return np.abs(G.bits(np.asarray(a)).astype(np.int64) - G.bits(np.asarray(b)).astype(np.int64))


def test_head_tail_golden():
    g = G.load('head_tail.npz')
    x = torch.from_numpy(g['x']).cuda().requires_grad_(True)
    y = heads.sigmoid_channels_last(x, num_classes=5)
    assert y.shape == g['y'].shape and y.dtype == torch.float32 and y.is_contiguous()
    assert ulp_diff(y.detach().cpu().numpy(), g['y']).max() <= ULP_VS_CPU
    (y * torch.from_numpy(g['wgt']).cuda()).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g['gx'], rtol=2e-6, atol=1e-12)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 720, 25, 25), (3, 365, 13, 7), (1, 80, 100, 100),
                                   (2, 1, 7, 7), (1, 36, 64, 65), (2, 4, 3, 5), (1, 130, 1, 1)])
def test_head_tail_equals_the_reference_ops_on_the_gpu(shape, dtype):
    gen = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=gen) * 4 - 2).to(dtype).cuda()
    x[0, 0, 0, 0] = 30.0          # saturates to 1
    x[-1, -1, -1, -1] = -110.0    # exp overflow -> 0
    xr = x.clone().requires_grad_(True)
    xk = x.clone().requires_grad_(True)
    want = O.head_tail(xr)                       # torch ops on CUDA
    got = heads.sigmoid_channels_last(xk)
    G.assert_bit_equal(got.detach().cpu().numpy(), want.detach().cpu().numpy(), 'probabilities')
    wgt = torch.randn(want.shape, generator=gen).cuda()
    (want * wgt).sum().backward()
    (got * wgt).sum().backward()
    assert xk.grad.dtype == dtype
    G.assert_bit_equal(xk.grad.float().cpu().numpy(), xr.grad.float().cpu().numpy(), 'gradient')


def test_head_tail_feeds_the_loss_and_decoder():
    """RetinaNet-style use: logits -> tail -> RetinaLoss / RetinaDecoder; the loss gradient reaches
    the logits through the tail's backward kernel."""
    from b200det import synth, losses, decode
    gen = torch.Generator().manual_seed(3)
    B, C, A = 2, 8, 9
    sizes = synth.pyramid_sizes(128)
    logits = [(torch.randn((B, A * C, p, p), generator=gen) - 4.0).cuda().requires_grad_(True)
              for p in sizes]
    reg = [(torch.randn((B, p, p, A, 4), generator=gen) * 0.2).cuda() for p in sizes]
    ann = synth.make_annotations(B, 6, 128, C, seed=4, min_gt=3).cuda()
    cls = [heads.sigmoid_channels_last(x, num_classes=C) for x in logits]
    crit = losses.RetinaLoss(**synth.RETINA_KW)
    d = crit([cls, reg], ann)
    (d['cls_loss'] + d['reg_loss']).backward()
    # same through torch's ops for the tail
    logits_t = [x.detach().clone().requires_grad_(True) for x in logits]
    cls_t = [O.head_tail(x, num_classes=C) for x in logits_t]
    dt = crit([cls_t, reg], ann)
    (dt['cls_loss'] + dt['reg_loss']).backward()
    assert d['cls_loss'].item() == dt['cls_loss'].item()
    for a, b in zip(logits, logits_t):
        G.assert_bit_equal(a.grad.cpu().numpy(), b.grad.cpu().numpy(), 'd loss / d logits')
    s, c, bx = decode.RetinaDecoder(**synth.RETINA_KW)([[t.detach() for t in cls], reg])
    s2, c2, bx2 = decode.RetinaDecoder(**synth.RETINA_KW)([[t.detach() for t in cls_t], reg])
    G.assert_bit_equal(s, s2)
    G.assert_bit_equal(bx, bx2)


def test_head_tail_rejects_cpu_and_bad_shapes():
    with pytest.raises(RuntimeError):
        heads.sigmoid_channels_last(torch.zeros(1, 4, 2, 2))
    with pytest.raises(RuntimeError):
        heads.sigmoid_channels_last(torch.zeros(4, 2, 2).cuda())
    empty = heads.sigmoid_channels_last(torch.zeros(0, 4, 2, 2).cuda())
    assert empty.shape == (0, 2, 2, 4)
