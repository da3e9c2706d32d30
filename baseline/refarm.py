"""Loads the vendored, UNMODIFIED reference classes from baseline/_ref/ (see fetch_ref.sh).

Used only by bench.py's reference arm / cpu_baseline leg: the reference's own RetinaLoss /
RetinaDecoder / FCOSLoss / FCOSDecoder on CPU tensors.  Never imported by the product package.
"""
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.isfile(os.path.join(REF_DIR, 'simpleAICV', 'detection', 'losses.py'))


def load():
    """Returns (losses_module, decode_module) of the vendored reference, or raises RuntimeError."""
    if not available():
        raise RuntimeError('baseline/_ref is empty: run baseline/fetch_ref.sh where /root/reference exists')
    if 'traitlets' not in sys.modules:
        try:
            import traitlets  # noqa: F401
        except ImportError:   # unused import of losses.py:4
            shim = types.ModuleType('traitlets')
            shim.Instance = object
            sys.modules['traitlets'] = shim
    mod = sys.modules.get('simpleAICV')
    if mod is not None and not str((list(getattr(mod, '__path__', [])) or [''])[0]).startswith(REF_DIR):
        raise RuntimeError('another simpleAICV package is already imported in this process')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from simpleAICV.detection import losses as ref_losses
    from simpleAICV.detection import decode as ref_decode
    return ref_losses, ref_decode
