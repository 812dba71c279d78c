"""Compile the reference's own DFA3D CUDA extension, UNMODIFIED, from the sources where they lie under
/root/reference, into oracle/_ref/ (git-ignored, travels to the GPU box).  Test infrastructure only:
it is the GPU-side ground truth the oracle (and the product) are pinned against.

The reference's setup.py cannot be used (forces -std=c++14 and no arch flags,
packages/3D-deformable-attention/DFA3D/setup.py:156-192); the sources compile as-is with
torch.utils.cpp_extension for sm_100a.
"""
from __future__ import annotations

import glob
import os
import sys
from pathlib import Path

REF = Path('/root/reference/packages/3D-deformable-attention/DFA3D/dfa3D/ops/csrc')
OUT = Path(__file__).resolve().parent / '_ref'
NAME = 'dfa3d_ref_ext'


def so_path() -> Path | None:
    c = sorted(OUT.glob(f'{NAME}*.so'))
    return c[0] if c else None


def build(quiet: bool = False) -> Path | None:
    if so_path() is not None:
        return so_path()
    if not REF.exists():
        if not quiet:
            print('reference sources not present; nothing to build')
        return None
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '8')
    from torch.utils import cpp_extension
    OUT.mkdir(exist_ok=True)
    srcs = sorted(glob.glob(str(REF / '*.cpp'))) + sorted(glob.glob(str(REF / 'cuda' / '*.cu'))) + \
        sorted(glob.glob(str(REF / 'cuda' / '*.cpp')))
    cpp_extension.load(
        name=NAME, sources=srcs, extra_include_paths=[str(REF / 'common'), str(REF / 'common' / 'cuda')],
        extra_cflags=['-DWITH_CUDA', '-O2'],
        extra_cuda_cflags=['-DWITH_CUDA', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo'],
        build_directory=str(OUT), with_cuda=True, is_python_module=False, verbose=not quiet)
    return so_path()


def load():
    """Import the built extension as a Python module (needs a CUDA device to be useful)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    p = so_path()
    if p is None:
        return None
    spec = importlib.util.spec_from_file_location(NAME, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(quiet='-q' in sys.argv))
