"""Builds libb200det.so (hand-written CUDA, sm_100a only) in-tree with nvcc.

    python -m b200det._build          # or __graft_entry__.build()

The library is a plain C-ABI shared object (include/b200det.h); it links only cudart.
assign.cu / decode.cu are compiled with -fmad=false because their float32 arithmetic must be
bit-identical to the reference's op-by-op torch / NumPy arithmetic; focal.cu may contract.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libb200det.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']
SOURCES = {
    'assign.cu': ['-fmad=false'],
    'decode.cu': ['-fmad=false'],
    'focal.cu': [],
    'api.cu': [],
    'heads.cu': [],
    'logits.cu': [],
    'queries.cu': [],
    'exchange.cu': [],
    'evaluate.cu': ['-fmad=false'],
    'iou.cu': ['-fmad=false'],
}


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libb200det.so cannot be built')
    return nvcc


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG_DIR, '..', 'include', 'b200det.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles every CUDA source for sm_100a and links libb200det.so. Returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, 'csrc', '_obj')
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    tune = os.environ.get('B200DET_NVCC_EXTRA', '').split()
    for src, extra in SOURCES.items():
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [nvcc] + ARCH + COMMON + extra + tune + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
            print(' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ['-shared', '-o', LIB_PATH] + objs
    subprocess.run(cmd, check=True)
    return LIB_PATH


def fastpath_path():
    import sysconfig
    return os.path.join(PKG_DIR, '_fastpath' + sysconfig.get_config_var('EXT_SUFFIX'))


def build_fastpath(force=False, verbose=False):
    """Compiles csrc/fastpath.cpp -- the optional host-side marshalling fast path (pybind11 + ATen,
    no CUDA code; it calls the C ABI through function pointers) -- in-tree with g++.  Returns the
    path, or None when it cannot be built here (no compiler / torch headers): the Python layer then
    simply keeps doing the marshalling itself."""
    src = os.path.join(CSRC, 'fastpath.cpp')
    out = fastpath_path()
    hdr = os.path.join(PKG_DIR, '..', 'include', 'b200det.h')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src),
                                                                          os.path.getmtime(hdr)):
        return out
    try:
        import sysconfig
        import torch
        from torch.utils import cpp_extension as ce
        gxx = shutil.which('g++')
        if gxx is None:
            raise RuntimeError('g++ not found')
        tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
        cmd = [gxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-fvisibility=hidden',
               '-DTORCH_EXTENSION_NAME=_fastpath', '-DTORCH_API_INCLUDE_EXTENSION_H',
               f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
        cmd += [f'-I{p}' for p in ce.include_paths()] + [f'-I{sysconfig.get_paths()["include"]}']
        cmd += [src, '-o', out, f'-L{tlib}', '-ltorch_python', '-ltorch', '-ltorch_cpu', '-lc10',
                f'-Wl,-rpath,{tlib}']
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return out
    except Exception as exc:   # noqa: BLE001 -- optional component
        print(f'[b200det] host fast path not built ({exc}); the Python marshalling path is used',
              file=sys.stderr)
        return None


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    print(build_fastpath(force='--force' in sys.argv, verbose='-v' in sys.argv))
