"""Loads the UNMODIFIED reference (read-only, /root/reference) for pinning the oracle.

Only usable in the development container; the GPU box has no /root/reference, so nothing marked
`gpu` may import this.  One shim: simpleAICV/detection/losses.py:4 imports an unused symbol from
`traitlets`, which is not installed here.
"""
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'simpleAICV', 'detection'))


def load():
    """Returns (losses_module, decode_module, anchor_module) of the reference."""
    if not available():
        raise RuntimeError('reference checkout not present')
    if 'traitlets' not in sys.modules:
        try:
            import traitlets  # noqa: F401
        except ImportError:
            shim = types.ModuleType('traitlets')
            shim.Instance = object
            sys.modules['traitlets'] = shim
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from simpleAICV.detection import losses as ref_losses
    from simpleAICV.detection import decode as ref_decode
    from simpleAICV.detection.models import anchor as ref_anchor
    return ref_losses, ref_decode, ref_anchor


def load_face():
    """Returns (losses_module, decode_module) of simpleAICV.face_detection in the reference."""
    load()
    from simpleAICV.face_detection import losses as face_losses
    from simpleAICV.face_detection import decode as face_decode
    return face_losses, face_decode
