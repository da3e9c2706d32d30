"""np.exp over ALL 65536 float16 bit patterns, as NumPy computes it on a host whose NumPy takes the
AVX512_SPR half-precision path (Intel SVML `exp` for float16, up to ~1.3 ulp from the correctly
rounded value) -- the hosts B200 boxes have.  The decoders of the reference call np.exp on the
float16 regression array when the head is float16 (decode.py:260, :356); their result therefore
depends on the host CPU.  b200det's decoders reproduce the host's own behaviour: this table when
NumPy reports AVX512_SPR, else half(expf(float(x))) (NumPy's generic half loop).

    python tools/make_half_exp_table.py     # writes <package>/data/np_exp_f16_avx512spr.npy

Must run on an AVX512_SPR host (asserted); NumPy version recorded in the file name's sidecar."""
import json
import os

import numpy as np

try:
    from numpy._core._multiarray_umath import __cpu_features__ as FEATURES
except ImportError:   # numpy < 2
    from numpy.core._multiarray_umath import __cpu_features__ as FEATURES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'simpleaicv-pytorch-imagenet-coco-training_b200', 'data')


def main():
    assert FEATURES.get('AVX512_SPR'), 'this host does not take NumPy\'s AVX512_SPR half path'
    x = np.arange(65536, dtype=np.uint16).view(np.float16)
    with np.errstate(all='ignore'):
        y = np.exp(x)
        generic = np.exp(x.astype(np.float64)).astype(np.float32).astype(np.float16)
    assert y.dtype == np.float16
    np.save(os.path.join(OUT, 'np_exp_f16_avx512spr.npy'), y.view(np.uint16))
    differ = int((y.view(np.uint16) != generic.view(np.uint16)).sum() -
                 (np.isnan(y) & np.isnan(generic) & (y.view(np.uint16) != generic.view(np.uint16))).sum())
    with open(os.path.join(OUT, 'np_exp_f16_avx512spr.json'), 'w') as f:
        json.dump({'numpy': np.__version__, 'entries': 65536,
                   'differ_from_generic_half_loop': differ}, f)
    print('table written;', differ, 'of 65536 entries differ from half(expf(float(x)))')


if __name__ == '__main__':
    main()
