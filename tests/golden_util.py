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
