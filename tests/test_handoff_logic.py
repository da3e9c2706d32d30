"""Host-side logic of the sweep hand-over (b200det/_handoff.py) on CPU tensors: when a criterion is
asked to produce a decoder's keys, and when a decoder may consume them.  (The kernels and the
device-side `stale` check are covered by tests/test_gpu_handoff.py.)"""
import ctypes
import gc

import pytest
import torch

from b200det import _handoff


class _Dec:
    def __init__(self, fcos=False):
        self._is_fcos = fcos


class _Crit:
    def __init__(self, fcos=False):
        self._is_fcos = fcos


DEV = torch.device('cuda', 0)
ST = ctypes.c_void_p(1234)
ST2 = ctypes.c_void_p(99)


@pytest.fixture(autouse=True)
def _fresh():
    was = _handoff.ENABLED
    _handoff.ENABLED = True
    _handoff.reset()
    yield
    _handoff.ENABLED = was
    _handoff.reset()


def _tensors(n=3):
    return [torch.zeros(4, 5) for _ in range(n)]


def _shapes(ts):
    return tuple(t.shape for t in ts)


def test_nothing_is_offered_before_a_decoder_asked():
    ts = _tensors()
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is None


def test_wish_offer_produce_take_roundtrip():
    dec, ts = _Dec(), _tensors()
    assert not _handoff.take(dec, DEV, ST, ts, 0.05)
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is dec
    _handoff.produced(dec, DEV, ST, ts, 0.05)
    assert _handoff.take(dec, DEV, ST, ts, 0.05)
    assert not _handoff.take(dec, DEV, ST, ts, 0.05)   # consumed: a second decoder call sweeps itself


def test_every_class_count_may_wish():
    """r02: class counts that are not a multiple of 4 (Objects365's 365) are covered by the raw-tile
    fused sweep, so they hand over like the others"""
    dec, ts = _Dec(), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 365)
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is dec


@pytest.mark.parametrize('what', ['version', 'object', 'threshold', 'decoder', 'count', 'stream'])
def test_take_refuses_anything_but_the_very_same_inputs(what):
    dec, ts = _Dec(), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    _handoff.offer(_Crit(), DEV, ST, _shapes(ts))
    _handoff.produced(dec, DEV, ST, ts, 0.05)
    who, got, thr, st = dec, list(ts), 0.05, ST
    if what == 'version':
        ts[1].add_(1)                  # an in-place torch op bumps the version counter
    elif what == 'object':
        got[2] = ts[2].clone()
    elif what == 'threshold':
        thr = 0.3
    elif what == 'decoder':
        who = _Dec()
    elif what == 'count':
        got = got[:2]
    elif what == 'stream':
        st = ST2
    assert not _handoff.take(who, DEV, st, got, thr)
    if what != 'stream':               # the record is gone either way
        assert not _handoff.take(dec, DEV, ST, ts, 0.05)


def test_family_and_shapes_must_match():
    dec, ts = _Dec(fcos=True), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    assert _handoff.offer(_Crit(fcos=False), DEV, ST, _shapes(ts)) is None
    assert _handoff.offer(_Crit(fcos=True), DEV, ST, _shapes(_tensors(2))) is None
    assert _handoff.offer(_Crit(fcos=True), DEV, ST, _shapes(ts)) is dec


def test_unconsumed_handovers_disarm_the_wish():
    dec, ts = _Dec(), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    produced = 0
    for _ in range(6):                 # a criterion-only loop
        if _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is not None:
            _handoff.produced(dec, DEV, ST, ts, 0.05)
            produced += 1
    assert produced == 2
    _handoff.take(dec, DEV, ST, ts, 0.05)
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)       # the next decoder miss re-arms it
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is dec


def test_a_dead_decoder_is_not_served():
    dec, ts = _Dec(), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    _handoff.offer(_Crit(), DEV, ST, _shapes(ts))
    _handoff.produced(dec, DEV, ST, ts, 0.05)
    del dec
    gc.collect()
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is None


def test_switch():
    dec, ts = _Dec(), _tensors()
    _handoff.ENABLED = False
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    assert _handoff.offer(_Crit(), DEV, ST, _shapes(ts)) is None


def test_a_pending_handover_does_not_keep_the_head_outputs_alive():
    """the record holds weak references: dropping the last user reference frees the tensors, and the
    dead record matches nothing"""
    import gc
    import weakref
    dec, ts = _Dec(), _tensors()
    _handoff.wish(dec, DEV, ST, _shapes(ts), 80)
    _handoff.offer(_Crit(), DEV, ST, _shapes(ts))
    _handoff.produced(dec, DEV, ST, ts, 0.05)
    probe = weakref.ref(ts[0])
    fresh = _tensors()
    del ts
    gc.collect()
    assert probe() is None
    assert not _handoff.take(dec, DEV, ST, fresh, 0.05)
