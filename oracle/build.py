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


if __name__ == '__main__':
    print(build_c(force='--force' in sys.argv))
    print(build_ref(force='--force' in sys.argv))
