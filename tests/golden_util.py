"""Helpers shared by the golden / parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
IOU_TYPES = ['IoU', 'GIoU', 'DIoU', 'CIoU', 'EIoU']


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def retina_inputs(d):
    n = len([k for k in d.files if k.startswith('cls')])
    cls = [torch.from_numpy(d[f'cls{i}'].copy()) for i in range(n)]
    reg = [torch.from_numpy(d[f'reg{i}'].copy()) for i in range(n)]
    return [cls, reg], torch.from_numpy(d['annotations'].copy())


def fcos_inputs(d):
    n = len([k for k in d.files if k.startswith('cls')])
    cls = [torch.from_numpy(d[f'cls{i}'].copy()) for i in range(n)]
    reg = [torch.from_numpy(d[f'reg{i}'].copy()) for i in range(n)]
    ctr = [torch.from_numpy(d[f'ctr{i}'].copy()) for i in range(n)]
    return [cls, reg, ctr], torch.from_numpy(d['annotations'].copy())


def half_nan_inputs(d, which, reg='f16'):
    """Inputs of tests/golden/half_nan.npz.  which: 'r' / 'f' = Retina / FCOS with a half-precision
    regression head (reg = 'f16' | 'bf16' | 'f32': the stored half values, as that dtype or upcast);
    'rn' / 'fn' = float32 heads with NaN scores planted."""
    n = len([k for k in d.files if k.startswith(which + '_cls')])
    cls = [torch.from_numpy(d[f'{which}_cls{i}'].copy()) for i in range(n)]
    if which in ('rn', 'fn'):
        regs = [torch.from_numpy(d[f'{which}_reg{i}'].copy()) for i in range(n)]
    elif reg == 'bf16':
        regs = [torch.from_numpy(d[f'{which}_reg{i}_bf16_bits'].copy()).view(torch.bfloat16)
                for i in range(n)]
    else:
        regs = [torch.from_numpy(d[f'{which}_reg{i}_f16'].copy()) for i in range(n)]
        if reg == 'f32':
            regs = [r.float() for r in regs]
    out = [cls, regs]
    if which in ('f', 'fn'):
        out.append([torch.from_numpy(d[f'{which}_ctr{i}'].copy()) for i in range(n)])
    ann = d[f'{which[0]}_annotations']
    return out, torch.from_numpy(ann.copy())


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=''):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    if a.dtype.kind == 'f' or b.dtype.kind == 'f':
        same = bits(a) == bits(b)
        same |= np.isnan(a) & np.isnan(b)
    else:
        same = a == b
    assert same.all(), f'{what}: {int((~same).sum())} of {same.size} elements differ'
