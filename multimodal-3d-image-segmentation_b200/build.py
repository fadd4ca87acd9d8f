"""Builds libhno_b200.so (sm_100a only) in-tree with nvcc.

Usage:  python multimodal-3d-image-segmentation_b200/build.py [--force] [--verbose]

The shared object has no torch / Python dependency: plain C ABI declared in include/hno_b200.h.
nvcc cross-compiles without a GPU; cudart is linked statically so the library also loads on a
CPU-only box (the plan/table builders work there, kernel launches return an error).
"""
import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libhno_b200.so')

SOURCES = ['api.cu', 'dht_plan.cu', 'dht_kernels.cu', 'dht_mid.cu', 'pwconv_kernels.cu', 'pwconv_bwd_tc.cu', 'stem_kernels.cu', 'head_kernels.cu',
           'modes_kernels.cu', 'modechain_kernels.cu', 'input_kernels.cu', 'tc_stream.cu', 'gemm_tc.cu', 'tc_analysis.cu',
           'mha_kernels.cu', 'dsconv_kernels.cu', 'spectral_core.cu', 'fourier_kernels.cu']

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
              '-I', INCLUDE, '-I', CSRC]


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: the hno_b200 CUDA library cannot be built')
    return nvcc


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ['../../include/hno_b200.h']:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    """Compiles every translation unit (in parallel) and links the shared object. Returns its path."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, 'digest.txt')
    digest = _digest()

    def current():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest

    if not force and current():
        return LIB
    import fcntl
    with open(os.path.join(BUILD, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # one builder at a time (ranks of a torchrun job race here otherwise)
        if not force and current():       # another process built it while this one waited
            return LIB
        return _build_locked(digest, stamp, verbose)


def _build_locked(digest, stamp, verbose):
    nvcc = _nvcc()
    jobs = [(nvcc, s, os.path.join(BUILD, s.replace('.cu', '.o')), verbose) for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        results = list(ex.map(_compile, jobs))
    failed = False
    for src, rc, log in results:
        if verbose or rc != 0:
            sys.stderr.write(f'--- {src} (rc={rc})\n{log}\n')
        failed |= rc != 0
    if failed:
        raise RuntimeError('nvcc failed, see log above')
    tmp = LIB + f'.tmp{os.getpid()}'
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
           '-o', tmp] + [j[2] for j in jobs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    os.replace(tmp, LIB)  # atomic: a concurrent dlopen sees the old or the new library, never a partial file
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
