"""Import shim: `import b200det` loads the package whose sources live in
`simpleaicv-pytorch-imagenet-coco-training_b200/` (a directory name Python cannot import directly).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         'simpleaicv-pytorch-imagenet-coco-training_b200')
_spec = _ilu.spec_from_file_location('b200det', _os.path.join(_pkg_dir, '__init__.py'),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules['b200det'] = _mod
_spec.loader.exec_module(_mod)
