"""Builds sgcdet_b200/_C/libsgcdet_b200.so in-tree with nvcc for sm_100a (no torch headers needed)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
OUT_DIR = ROOT / '_C'
LIB = OUT_DIR / 'libsgcdet_b200.so'
STAMP = OUT_DIR / 'build.stamp'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xptxas=-v', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; the sgcdet_b200 CUDA library cannot be built')


def sources():
    return sorted(CSRC.glob('*.cu'))


def _digest() -> str:
    h = hashlib.sha256()
    hdr = CSRC.parent.parent / 'include' / 'sgcdet_b200.h'
    for p in sorted(list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh'))) + ([hdr] if hdr.exists() else []):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = OUT_DIR / (src.stem + '.o')
        cmd = [nvcc, *NVCC_FLAGS, '-I', str(CSRC), '-c', str(src), '-o', str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src.name}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src.name}:\n{out}')
    cmd = [nvcc, '-shared', '-o', str(LIB), *map(str, objs), '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    (OUT_DIR / 'build.log').write_text('\n'.join(log))
    STAMP.write_text(dig)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
