"""Build recipe of the C-ABI CUDA library (hvrnet_b200/libhvr_b200.so) for sm_100a.

nvcc cross-compiles here without a GPU; the built .so is git-ignored but travels to the
GPU box with the gpurun snapshot.  `python -m hvrnet_b200.build [--force]`.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', '_obj')
LIB = os.path.join(HERE, 'libhvr_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
          '--expt-relaxed-constexpr']
# file -> extra flags.  -fmad=false: the parity contract of RoIAlign / box decoding / IoU is
# "every product and sum rounded once" (oracle/c/hvr_oracle.c is built with -ffp-contract=off).
SOURCES = {
    'elementwise.cu': ['-fmad=false'],
    'roi_align.cu': ['-fmad=false'],
    'nms.cu': ['-fmad=false'],
    'igemm_tc.cu': [],
    'preprocess.cu': ['-fmad=false'],
    'intervideo.cu': ['-fmad=false'],
    'window.cu': ['-fmad=false'],
    'relation.cu': ['-fmad=false'],
}


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + \
        [os.path.join(HERE, '..', 'include', 'hvr_b200.h'), os.path.abspath(__file__)]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src, flags, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + ARCH + COMMON + flags + ['-c', src, '-o', obj]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libhvr_b200.so.  Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    jobs = []
    for f, flags in SOURCES.items():
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + '.o')
        if force or _stale(obj, [src] + deps):
            jobs.append((src, flags))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), 8)) as ex:
            list(ex.map(lambda a: _compile(a[0], a[1], verbose), jobs))
    objs = [os.path.join(OBJ, f[:-3] + '.o') for f in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
