"""RetinaAnchors / FCOSPositions as drop-in classes (reference
simpleAICV/detection/models/anchor.py:5-130): tables produced by the kernels' own generator
(b200det_generate_rows) against vectors of the unmodified reference (tests/golden/tables.npz) and the
oracle, bit for bit; same constructor / call signature."""
import inspect

import numpy as np
import pytest

from b200det import synth
from oracle import det_oracle as O

import golden_util as G


def test_anchor_surface_matches_reference_signatures():
    from b200det import anchor
    sig = inspect.signature(anchor.__dict__['RetinaAnchors'].__init__).parameters
    assert list(sig) == ['self', 'areas', 'ratios', 'scales', 'strides']
    assert sig['ratios'].default == [0.5, 1, 2] and sig['strides'].default == [8, 16, 32, 64, 128]
    assert list(inspect.signature(anchor.RetinaAnchors.__call__).parameters) == ['self', 'fpn_feature_sizes']
    psig = inspect.signature(anchor.__dict__['FCOSPositions'].__init__).parameters
    assert list(psig) == ['self', 'strides'] and psig['strides'].default == [8, 16, 32, 64, 128]
    a = anchor.RetinaAnchors(**synth.RETINA_KW)
    t = G.load('tables.npz')
    for l, area in enumerate(a.areas):       # host-side table, no GPU needed
        G.assert_bit_equal(a.generate_base_anchors(area, a.scales, a.ratios), t['base_anchors'][l])


@pytest.mark.gpu
def test_gpu_anchor_tables_golden():
    from b200det import anchor
    t = G.load('tables.npz')
    anchors = anchor.RetinaAnchors(**synth.RETINA_KW)
    for size in (800, 1024):
        p = synth.pyramid_sizes(size)
        levels = anchors([[q, q] for q in p])
        assert [a.shape for a in levels] == [(q, q, 9, 4) for q in p]
        assert all(a.dtype == np.float32 and a.flags.writeable for a in levels)
        flat = np.concatenate([a.reshape(-1, 4) for a in levels], axis=0)
        assert np.array_equal(flat.astype(np.float64).sum(axis=0), t[f'anchors_{size}_sum'])
        G.assert_bit_equal(np.stack([a.reshape(-1, 4)[:18] for a in levels]), t[f'anchors_{size}_head'])
        G.assert_bit_equal(np.stack([a.reshape(-1, 4)[-9:] for a in levels]), t[f'anchors_{size}_tail'])
        want = O.retina_anchors([[q, q] for q in p], **synth.RETINA_KW)
        for a, w in zip(levels, want):
            G.assert_bit_equal(a, w, f'anchors {size}')
        pos = anchor.FCOSPositions(strides=synth.STRIDES)([[q, q] for q in p])
        assert [q_.shape for q_ in pos] == [(q, q, 2) for q in p]
        assert np.array_equal(np.concatenate([q_.reshape(-1, 2) for q_ in pos]).astype(np.float64).sum(axis=0),
                              t[f'positions_{size}_sum'])
        for q_, w in zip(pos, O.fcos_positions([[q, q] for q in p], synth.STRIDES)):
            G.assert_bit_equal(q_, w, f'positions {size}')
    # non-square maps ([W, H] order), odd strides, 6 anchors per location
    odd = anchor.RetinaAnchors(areas=[[24, 40], [48, 80]], ratios=[0.5, 1, 2], scales=[1, 1.5],
                               strides=[6, 12])
    lv = odd([[7, 5], [4, 3]])
    G.assert_bit_equal(lv[0], t['odd_anchors0'])
    G.assert_bit_equal(lv[1], t['odd_anchors1'])
    pos = anchor.FCOSPositions(strides=[6, 12])([[7, 5], [4, 3]])
    for q_, w in zip(pos, O.fcos_positions([[7, 5], [4, 3]], [6, 12])):
        G.assert_bit_equal(q_, w, 'odd positions')
