"""Build recipe for the oracle's native pieces (test infrastructure only).

* ``build_c()``   gcc-compiles oracle/c/hvr_oracle.c (the C restatement) into
                  oracle/_build/libhvr_oracle.so, strict IEEE (-ffp-contract=off).
* ``build_ref()`` compiles the REFERENCE's own mmdet/ops/nms/src/nms_cpu.cpp,
                  unmodified and where it lies under /root/reference, into
                  oracle/_ref/ (a torch pybind module).  No reference source is
                  copied into the repo.  Only possible in the build container;
                  the GPU box uses the prebuilt file.
"""
import importlib.util
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
C_SRC = os.path.join(HERE, 'c', 'hvr_oracle.c')
C_OUT = os.path.join(HERE, '_build', 'libhvr_oracle.so')
REF_SRC = '/root/reference/mmdet/ops/nms/src/nms_cpu.cpp'
REF_DIR = os.path.join(HERE, '_ref')
REF_NAME = 'hvr_ref_nms_cpu'


def _stale(out, src):
    return (not os.path.exists(out)) or os.path.getmtime(out) < os.path.getmtime(src)


def build_c(force=False):
    os.makedirs(os.path.dirname(C_OUT), exist_ok=True)
    if force or _stale(C_OUT, C_SRC):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-std=c11',
                               '-o', C_OUT, C_SRC, '-lm'])
    return C_OUT


def ref_so_path():
    p = os.path.join(REF_DIR, REF_NAME + '.so')
    return p if os.path.exists(p) else None


def build_ref(force=False):
    """Returns the path of the built module, or None when the reference tree is absent."""
    if not os.path.exists(REF_SRC):
        return ref_so_path()
    if ref_so_path() and not force:
        return ref_so_path()
    os.makedirs(REF_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name=REF_NAME, sources=[REF_SRC], build_directory=REF_DIR, verbose=False,
         extra_cflags=['-O2', '-w'])
    return ref_so_path()


def load_ref():
    """Import the compiled reference nms_cpu module (None if it was never built)."""
    p = ref_so_path()
    if p is None:
        return None
    import torch  # noqa: F401  (the module links against libtorch)
    spec = importlib.util.spec_from_file_location(REF_NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ROI_SRCS = ['/root/reference/mmdet/ops/roi_align/src/roi_align_cuda.cpp',
            '/root/reference/mmdet/ops/roi_align/src/roi_align_kernel.cu']
ROI_NAMES = {True: 'hvr_ref_roi_align_cuda', False: 'hvr_ref_roi_align_cuda_nofma'}


def ref_roi_so_path(fma=True):
    p = os.path.join(REF_DIR, ROI_NAMES[fma] + '.so')
    return p if os.path.exists(p) else None


def build_ref_roi_align(force=False, which=(True, False)):
    """Compiles the REFERENCE's own RoIAlign CUDA op (roi_align_cuda.cpp + roi_align_kernel.cu, unmodified,
    where they lie) for sm_100a into oracle/_ref/, twice: with nvcc's default floating-point contraction (what
    the reference's setup.py builds) and with -fmad=false (every product and sum rounded: the arithmetic
    contract of the CUDA product and of oracle/c).  oracle/shim/ref_compat.h maps the torch names removed
    since mmdetection v1 (AT_CHECK, THCudaCheck).  nvcc cross-compiles here without a GPU; the modules can only
    RUN on the GPU box, where tests/test_gpu_kernels.py loads them if present.  Returns the built paths."""
    if not all(os.path.exists(s) for s in ROI_SRCS):
        return [ref_roi_so_path(True), ref_roi_so_path(False)]
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    from torch.utils.cpp_extension import load
    shim = os.path.join(HERE, 'shim', 'ref_compat.h')
    for fma in which:
        if ref_roi_so_path(fma) and not force:
            continue
        bd = os.path.join(REF_DIR, 'build_' + ROI_NAMES[fma])
        os.makedirs(bd, exist_ok=True)
        load(name=ROI_NAMES[fma], sources=ROI_SRCS, build_directory=bd, verbose=False,
             extra_cflags=['-O2', '-w', '-include', shim],
             extra_cuda_cflags=['-O2', '-w', '-include', shim, '-gencode', 'arch=compute_100a,code=sm_100a'] +
                               ([] if fma else ['-fmad=false']))
        os.replace(os.path.join(bd, ROI_NAMES[fma] + '.so'), os.path.join(REF_DIR, ROI_NAMES[fma] + '.so'))
        shutil.rmtree(bd, ignore_errors=True)
    return [ref_roi_so_path(True), ref_roi_so_path(False)]


NMS_CUDA_SRCS = ['/root/reference/mmdet/ops/nms/src/nms_cuda.cpp', '/root/reference/mmdet/ops/nms/src/nms_kernel.cu']
NMS_CUDA_NAME = 'hvr_ref_nms_cuda'


def ref_nms_cuda_so_path():
    p = os.path.join(REF_DIR, NMS_CUDA_NAME + '.so')
    return p if os.path.exists(p) else None


def build_ref_nms_cuda(force=False):
    """Compiles the REFERENCE's own NMS CUDA op (nms_cuda.cpp + nms_kernel.cu, unmodified, where they lie) for
    sm_100a into oracle/_ref/.  oracle/shim/ref_compat_nms.h + shim/include/THC/THC.h stand in for the THC
    names torch removed (THCState, THCudaMalloc/Free, THCCeilDiv, THCudaCheck).  Runs only on the GPU box."""
    if not all(os.path.exists(s) for s in NMS_CUDA_SRCS):
        return ref_nms_cuda_so_path()
    if ref_nms_cuda_so_path() and not force:
        return ref_nms_cuda_so_path()
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    from torch.utils.cpp_extension import load
    shim, inc = os.path.join(HERE, 'shim', 'ref_compat_nms.h'), os.path.join(HERE, 'shim', 'include')
    bd = os.path.join(REF_DIR, 'build_' + NMS_CUDA_NAME)
    os.makedirs(bd, exist_ok=True)
    load(name=NMS_CUDA_NAME, sources=NMS_CUDA_SRCS, build_directory=bd, verbose=False,
         extra_cflags=['-O2', '-w', '-include', shim, '-I', inc],
         extra_cuda_cflags=['-O2', '-w', '-include', shim, '-I', inc, '-gencode', 'arch=compute_100a,code=sm_100a'])
    os.replace(os.path.join(bd, NMS_CUDA_NAME + '.so'), os.path.join(REF_DIR, NMS_CUDA_NAME + '.so'))
    shutil.rmtree(bd, ignore_errors=True)
    return ref_nms_cuda_so_path()


def load_ref_nms_cuda():
    """Import the compiled reference NMS CUDA module (None if it was never built).  Needs a CUDA device to run."""
    p = ref_nms_cuda_so_path()
    if p is None:
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(NMS_CUDA_NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_roi_align(fma=True):
    """Import a compiled reference RoIAlign module (None if it was never built).  Needs a CUDA device to run."""
    p = ref_roi_so_path(fma)
    if p is None:
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(ROI_NAMES[fma], p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build_c(force='--force' in sys.argv))
    print(build_ref(force='--force' in sys.argv))
    print(build_ref_roi_align(force='--force' in sys.argv))
    print(build_ref_nms_cuda(force='--force' in sys.argv))


def build_all_ref():
    """Every missing oracle/_ref module, each in its own process, side by side: from an empty oracle/_ref the four
    builds (g++ / nvcc over the torch headers) take about two minutes each.  No-op when they exist or when the
    reference tree is absent."""
    import sys
    todo = []
    if os.path.exists(REF_SRC) and not ref_so_path():
        todo.append('build_ref()')
    if all(os.path.exists(s_) for s_ in ROI_SRCS):
        todo += ['build_ref_roi_align(which=(%s,))' % fma for fma in (True, False) if not ref_roi_so_path(fma)]
    if all(os.path.exists(s_) for s_ in NMS_CUDA_SRCS) and not ref_nms_cuda_so_path():
        todo.append('build_ref_nms_cuda()')
    root = os.path.dirname(HERE)
    procs = [subprocess.Popen([sys.executable, '-c', 'from oracle import build as b; b.%s' % call], cwd=root) for call in todo]
    failed = [call for call, pr in zip(todo, procs) if pr.wait() != 0]
    if failed:
        raise RuntimeError('oracle/_ref: %s failed' % ', '.join(failed))

