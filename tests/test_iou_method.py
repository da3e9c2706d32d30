"""IoUMethod as a stand-alone operator (reference losses.py:28-123).

CPU: the oracle restatement reproduces the unmodified reference class bit for bit, values and
autograd gradients (tests/golden/iou_method.npz).  GPU (-m gpu): b200det.losses.IoUMethod
(b200det_iou_method) gives bit-identical values for every type built from correctly rounded ops
(CIoU within 2e-6 absolute: CUDA atanf vs the host's), gradients within 1e-5 relative + 1e-6."""
import numpy as np
import pytest
import torch

from oracle import det_oracle as O

import golden_util as G

TYPES = G.IOU_TYPES
BOX_TYPES = ['xyxy', 'xywh']


@pytest.fixture(scope='module')
def v():
    return G.load('iou_method.npz')


def pair(v, box_type):
    return (v['b1'], v['b2']) if box_type == 'xyxy' else (v['b1_xywh'], v['b2_xywh'])


@pytest.mark.parametrize('box_type', BOX_TYPES)
@pytest.mark.parametrize('iou_type', TYPES)
def test_oracle_iou_method_golden(v, iou_type, box_type):
    x1, x2 = pair(v, box_type)
    t1 = torch.from_numpy(x1.copy()).requires_grad_(True)
    t2 = torch.from_numpy(x2.copy()).requires_grad_(True)
    out = O.iou_method(t1, t2, iou_type, box_type)
    (out * torch.from_numpy(v['upstream'])).sum().backward()
    G.assert_bit_equal(out.detach().numpy(), v[f'{box_type}_{iou_type}'], 'values')
    G.assert_bit_equal(t1.grad.numpy(), v[f'{box_type}_{iou_type}_g1'], 'grad boxes1')
    G.assert_bit_equal(t2.grad.numpy(), v[f'{box_type}_{iou_type}_g2'], 'grad boxes2')


def test_oracle_iou_method_broadcast_golden(v):
    a, g = torch.from_numpy(v['b1'][:40].copy()), torch.from_numpy(v['b2'][:17].copy())
    for iou_type in ('IoU', 'DIoU', 'EIoU'):
        G.assert_bit_equal(O.iou_method(a.unsqueeze(1), g.unsqueeze(0), iou_type).numpy(),
                           v[f'bcast_{iou_type}'], iou_type)


def close(got, want, rel, abs_):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    bad = ~(np.abs(got - want) <= rel * np.abs(want) + abs_)
    bad &= ~(np.isnan(got) & np.isnan(want))
    assert not bad.any(), f'{int(bad.sum())} of {bad.size} differ, worst {np.abs(got - want).max()}'


@pytest.mark.gpu
@pytest.mark.parametrize('box_type', BOX_TYPES)
@pytest.mark.parametrize('iou_type', TYPES)
def test_gpu_iou_method_golden(v, iou_type, box_type):
    from b200det import losses
    x1, x2 = pair(v, box_type)
    t1 = torch.from_numpy(x1.copy()).cuda().requires_grad_(True)
    t2 = torch.from_numpy(x2.copy()).cuda().requires_grad_(True)
    out = losses.IoUMethod()(t1, t2, iou_type=iou_type, box_type=box_type)
    assert out.shape == (x1.shape[0],) and out.dtype == torch.float32
    (out * torch.from_numpy(v['upstream']).cuda()).sum().backward()
    want = v[f'{box_type}_{iou_type}']
    if iou_type == 'CIoU':
        close(out.detach().cpu().numpy(), want, 0, 2e-6)
    else:
        G.assert_bit_equal(out.detach().cpu().numpy(), want, 'values')
    ok = np.ones(x1.shape[0], dtype=bool)
    if iou_type == 'CIoU':
        # zero-area / inverted boxes: w / h is 0/0 or x/0, the value is still compared above, but
        # reverse-mode autograd and the kernel's forward-mode duals meet inf * 0 in different places
        # (the reference's own gradients are NaN or arbitrary there)
        b1, b2 = v['b1'], v['b2']
        ok = (b1[:, 2] > b1[:, 0]) & (b1[:, 3] > b1[:, 1]) & (b2[:, 2] > b2[:, 0]) & (b2[:, 3] > b2[:, 1])
        assert ok.sum() >= 80
    close(t1.grad.cpu().numpy()[ok], v[f'{box_type}_{iou_type}_g1'][ok], 1e-5, 1e-6)
    close(t2.grad.cpu().numpy()[ok], v[f'{box_type}_{iou_type}_g2'][ok], 1e-5, 1e-6)


@pytest.mark.gpu
def test_gpu_iou_method_broadcast(v):
    from b200det import losses
    fn = losses.IoUMethod()
    a = torch.from_numpy(v['b1'][:40].copy()).cuda()
    g = torch.from_numpy(v['b2'][:17].copy()).cuda()
    for iou_type in ('IoU', 'DIoU', 'EIoU'):
        got = fn(a.unsqueeze(1), g.unsqueeze(0), iou_type=iou_type)
        assert got.shape == (40, 17)
        G.assert_bit_equal(got.cpu().numpy(), v[f'bcast_{iou_type}'], iou_type)
        # the transposed broadcast and a general one (expanded) agree with it
        G.assert_bit_equal(fn(a.unsqueeze(0), g.unsqueeze(1), iou_type=iou_type).cpu().numpy(),
                           v[f'bcast_{iou_type}'].T, iou_type + ' transposed')
        got3 = fn(a.view(2, 20, 1, 4), g.view(1, 1, 17, 4), iou_type=iou_type)
        G.assert_bit_equal(got3.cpu().numpy().reshape(40, 17), v[f'bcast_{iou_type}'], '4-D')
    a.requires_grad_(True)
    g.requires_grad_(True)
    w = torch.from_numpy(v['bcast_w']).cuda()
    (fn(a.unsqueeze(1), g.unsqueeze(0), iou_type='DIoU') * w).sum().backward()
    close(a.grad.cpu().numpy(), v['bcast_DIoU_g1'], 1e-5, 1e-6)
    close(g.grad.cpu().numpy(), v['bcast_DIoU_g2'], 1e-5, 1e-6)


@pytest.mark.gpu
def test_gpu_iou_method_matches_the_assignment_matrix_full_size():
    """The [A,1,4] x [1,G,4] matrix of losses.py:350-353 for one 800x800 image (120 087 anchors x
    57 boxes): bit-identical to the oracle, and its row arg-max is what the assignment uses."""
    from b200det import losses, synth
    anchors = np.concatenate([a.reshape(-1, 4) for a in O.retina_anchors(
        [[p, p] for p in synth.pyramid_sizes(800)], synth.RETINA_KW['areas'],
        synth.RETINA_KW['ratios'], synth.RETINA_KW['scales'], synth.RETINA_KW['strides'])])
    ann = synth.make_annotations(1, 100, 800, 80, seed=3)[0]
    gt = ann[ann[:, 4] >= 0][:, :4]
    ta = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float32))
    want = O.iou_method(ta.unsqueeze(1), gt.unsqueeze(0), 'IoU')
    got = losses.IoUMethod()(ta.cuda().unsqueeze(1), gt.cuda().unsqueeze(0))
    G.assert_bit_equal(got.cpu().numpy(), want.numpy(), 'assignment IoU matrix')


@pytest.mark.gpu
def test_gpu_iou_method_edge_cases():
    from b200det import losses
    fn = losses.IoUMethod()
    empty = torch.zeros((0, 4), device='cuda')
    assert fn(empty, empty, iou_type='GIoU').shape == (0,)
    nan = torch.tensor([[0., 0., float('nan'), 10.]], device='cuda')
    box = torch.tensor([[0., 0., 10., 10.]], device='cuda')
    for t in TYPES:        # NaN coordinates on either side reach the result, like torch's ops
        assert torch.isnan(fn(nan, box, iou_type=t)).all() and torch.isnan(fn(box, nan, iou_type=t)).all()
        want = O.iou_method(nan.cpu(), box.cpu(), t)
        assert torch.isnan(want).all()
    with pytest.raises(AssertionError):
        fn(box, box, iou_type='SIoU')
    with pytest.raises(AssertionError):
        fn(box, box, box_type='cxcywh')
    with pytest.raises(RuntimeError):
        fn(box.cpu(), box)
    with pytest.raises(RuntimeError):
        fn(box.unsqueeze(1), box.unsqueeze(0), iou_type='GIoU')
    half = fn(box.half(), box.half())         # other float dtypes are upcast like `.float()`
    assert half.dtype == torch.float32 and half.item() == 1.0
