"""Build libb200fe.so (all sm_100a kernels + the C ABI) in-tree with nvcc.

    python pets-face-recognition_b200/build.py [--force]

Each .cu is compiled to an object (in parallel, cached by mtime) and linked into
b200/libb200fe.so.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
INCLUDE = HERE.parent / 'include'
OBJ = HERE / 'build'
LIB = HERE / 'b200' / 'libb200fe.so'
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '--use_fast_math',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '-I', str(INCLUDE), '-I', str(CSRC)]


def _stale(out: Path, deps) -> bool:
    return not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = True) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob('*.cuh')) + list(INCLUDE.glob('*.h')) + [Path(__file__)]
    sources = sorted(CSRC.glob('*.cu'))
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + '.o')
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + ['-Xptxas', '-v', '-c', str(src), '-o', str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(compile_one, jobs):
                (OBJ / (src.stem + '.ptxas.log')).write_text(r.stderr)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f'nvcc failed on {src.name}')
                if verbose:
                    warn = [l for l in r.stderr.splitlines() if 'warning' in l.lower() or 'spill' in l.lower() and ' 0 bytes spill' not in l]
                    print(f'[build] {src.name} ok' + (f' ({len(warn)} warnings/spill lines, see build/{src.stem}.ptxas.log)' if warn else ''))
    objs = [OBJ / (s.stem + '.o') for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', str(LIB)] + [str(o) for o in objs] + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
        if verbose:
            print(f'[build] linked {LIB}')
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
