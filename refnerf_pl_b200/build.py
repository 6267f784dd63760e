"""In-tree build of the C-ABI CUDA library (sm_100a only): `python -m refnerf_pl_b200.build`.

Each .cu is compiled to an object with nvcc (cross-compiles without a GPU) and linked into
`refnerf_pl_b200/librefnerf_b200.so`; objects are rebuilt only when a source or header is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, 'librefnerf_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')
SOURCES = ['raymarch.cu', 'raygen.cu', 'pointwise.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'chain_pair.cu', 'chain_x3.cu', 'chain_x3t.cu', 'mlp.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr', '-Wno-deprecated-gpu-targets']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _newest_header():
    t = 0.0
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in os.listdir(d):
            if f.endswith(('.cuh', '.h', '.inc')):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(verbose=False, force=False, extra_flags=()):
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc, *NVCC_FLAGS, *extra_flags, '-c', src, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        outs = list(ex.map(run, jobs))
    if verbose:
        for o in outs:
            if o.strip():
                print(o)
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        run([nvcc, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC'])
    return LIB


if __name__ == '__main__':
    print(build(verbose=True, force='--force' in sys.argv,
                extra_flags=('-Xptxas', '-v') if '--ptxas-v' in sys.argv else ()))
