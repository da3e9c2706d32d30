"""Sweep hand-over between a criterion call and the decoder call that follows it.

The reference's evaluation loop makes two calls per batch on the SAME head outputs
(tools/scripts.py:733-740):

    loss_value = criterion(outs_tuple, annots)
    scores, classes, boxes = decoder(outs_tuple)

Each of them streams the whole classification tensor -- 98 % of the step's HBM traffic.  The fused
sweep (csrc/decode.cu: fused_rows_kernel) produces the focal sum AND the decoder's per-row key /
class from one read, so when the two calls really do see the same tensors the criterion's sweep can
do the decoder's work as well and `decoder(outs_tuple)` is left with selection + box decode + NMS.
The caller keeps the reference's two calls; nothing is skipped, cls is read once instead of twice.

How it is kept safe (all host-side logic lives here; no numerics):

  * A decoder call that could not use a hand-over and whose own sweep succeeded (so the shapes and the
    class count are ones the sweeps support) leaves a WISH for its (device, stream): "the next
    no-grad criterion call on head outputs of these shapes may produce my keys" (same detector
    family, the decoder's scratch for that stream exists).
  * A no-grad criterion call that finds a matching wish runs b200det_loss_forward_keys, writing keys /
    classes (thresholded with the decoder's min_score_threshold) into the decoder's own scratch, and
    records READY = (the very tensor objects it read, their autograd version counters, the
    threshold).  The record holds WEAK references: it never extends the life of a batch of head
    outputs (a validation loop that ends on a criterion call would otherwise pin gigabytes until the
    next evaluation), and a dead reference can match nothing -- memory recycled for other tensors is
    reached through other tensor objects, which fail the identity test below.
  * The next decoder call consumes READY only if it is called with the same tensor objects (`is`),
    unchanged version counters, on the same device / stream, with the same threshold; then it runs
    b200det_decode_from_keys.  Anything else: the record is dropped and the decoder sweeps itself.
  * In-place writes that bypass torch's version counters (a CUDA-graph replay into static output
    buffers between the two calls, raw-pointer kernels) are caught on the device: the select kernel
    re-derives the key of every selected row from the class score it names and raises a `stale`
    flag on any mismatch; the decoder then decodes again from scratch.
  * A criterion whose keys are not consumed (criterion-only loops) pays ~25 % more for its sweep:
    after two unconsumed hand-overs the wish is dropped; the next decoder miss re-arms it.
  * Never during CUDA-graph capture, never on the gradient path.  B200DET_HANDOFF=0 switches it off.
"""
import os
import weakref

ENABLED = os.environ.get('B200DET_HANDOFF', '1') != '0'

_wishes = {}   # (device index, raw stream) -> _Wish
_ready = {}    # (device index, raw stream) -> _Ready

stats = {'produced': 0, 'consumed': 0, 'stale': 0, 'dropped': 0}


class _Wish:
    __slots__ = ('decoder', 'shapes', 'is_fcos', 'misses')

    def __init__(self, decoder, shapes):
        self.decoder = weakref.ref(decoder)
        self.shapes = shapes
        self.is_fcos = decoder._is_fcos
        self.misses = 0


class _Ready:
    __slots__ = ('decoder', 'tensors', 'versions', 'min_score')


def _key(device, stream):
    return (device.index, stream.value)


def wish(decoder, device, stream, shapes, num_classes):
    """Called by a decoder that had to sweep itself."""
    if not ENABLED:
        return
    key = _key(device, stream)
    old = _wishes.get(key)
    if old is not None and old.decoder() is decoder and old.shapes == shapes:
        return
    if len(_wishes) >= 8:
        _wishes.clear()
    _wishes[key] = _Wish(decoder, shapes)


def offer(owner, device, stream, shapes):
    """Called by a no-grad criterion call: the decoder to produce keys for, or None."""
    if not ENABLED:
        return None
    key = _key(device, stream)
    w = _wishes.get(key)
    if w is None:
        return None
    stale = _ready.pop(key, None)
    if stale is not None:        # the previous hand-over was never consumed
        stats['dropped'] += 1
        w.misses += 1
        if w.misses >= 2:
            del _wishes[key]
            return None
    dec = w.decoder()
    if dec is None or w.is_fcos != owner._is_fcos or w.shapes != shapes:
        return None
    return dec


def produced(decoder, device, stream, tensors, min_score):
    r = _Ready()
    r.decoder = weakref.ref(decoder)
    r.tensors = [weakref.ref(t) for t in tensors]
    r.versions = [t._version for t in tensors]
    r.min_score = min_score
    _ready[_key(device, stream)] = r
    stats['produced'] += 1


def take(decoder, device, stream, tensors, min_score):
    """Called by every decoder call: True when the keys in the decoder's scratch for this stream were
    produced from exactly these tensors.  The record is consumed (or invalidated) either way."""
    r = _ready.pop(_key(device, stream), None)
    if r is None:
        return False
    if r.decoder() is not decoder or r.min_score != min_score or len(r.tensors) != len(tensors):
        stats['dropped'] += 1
        return False
    for a, b, v in zip(r.tensors, tensors, r.versions):
        if a() is not b or b._version != v:
            stats['dropped'] += 1
            return False
    w = _wishes.get(_key(device, stream))
    if w is not None:
        w.misses = 0
    stats['consumed'] += 1
    return True


def reset():
    _wishes.clear()
    _ready.clear()
