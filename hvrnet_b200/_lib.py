"""ctypes binding of the C-ABI library (include/hvr_b200.h).

The product has no CPU path: if libhvr_b200.so is missing or an entry point fails, an
exception is raised - nothing falls back to torch ops or to the oracle.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libhvr_b200.so')

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t


class HvrIGemm(ctypes.Structure):
    """Mirror of struct HvrIGemm (include/hvr_b200.h)."""
    _fields_ = [
        ('a_hi', c_vp), ('a_lo', c_vp),
        ('a_c', c_int), ('a_w', c_int), ('a_h', c_int), ('a_b', c_int),
        ('a_stride_w', c_i64), ('a_stride_h', c_i64), ('a_stride_b', c_i64),
        ('ntaps', c_int), ('tap_dx', c_int * 9), ('tap_dy', c_int * 9),
        ('out_w', c_int), ('out_h', c_int), ('batch', c_int),
        ('tile_w', c_int), ('tile_h', c_int),
        ('b_hi', c_vp), ('b_lo', c_vp), ('n', c_int), ('ldb', c_i64),
        ('alpha', c_f32), ('bias', c_vp),
        ('res_hi', c_vp), ('res_lo', c_vp), ('ld_res', c_i64),
        ('relu', c_int),
        ('out_hi', c_vp), ('out_lo', c_vp), ('ld_out', c_i64),
        ('out_f32', c_vp), ('ld_f32', c_i64),
        ('outT_hi', c_vp), ('outT_lo', c_vp), ('ld_outT', c_i64),
        ('passes', c_int),
        ('b_stride_batch', c_i64),
        ('a2_hi', c_vp), ('a2_lo', c_vp),
        ('a2_c', c_int), ('a2_w', c_int), ('a2_h', c_int), ('a2_b', c_int),
        ('a2_stride_w', c_i64), ('a2_stride_h', c_i64), ('a2_stride_b', c_i64),
    ]


class HvrRelationWeights(ctypes.Structure):
    """Mirror of struct HvrRelationWeights (include/hvr_b200.h)."""
    _fields_ = [('dim', c_int),
                ('q_hi', c_vp), ('q_lo', c_vp), ('q_bias', c_vp),
                ('k_hi', c_vp), ('k_lo', c_vp), ('k_bias', c_vp),
                ('o_hi', c_vp), ('o_lo', c_vp), ('o_bias', c_vp)]


# name -> (restype, argtypes); every symbol include/hvr_b200.h declares
SIGNATURES = {
    'hvr_strerror': (ctypes.c_char_p, [c_int]),
    'hvr_last_cuda_error': (c_int, []),
    'hvr_abi_version': (c_int, []),
    'hvr_launch_count': (ctypes.c_uint64, []),
    'hvr_split_f32': (c_int, [c_vp, c_sz, c_vp, c_vp, c_vp]),
    'hvr_merge_f32': (c_int, [c_vp, c_vp, c_sz, c_vp, c_vp]),
    'hvr_split_f32_2d': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'hvr_transpose_split': (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_vp]),
    'hvr_nchw_to_nhwc_split': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'hvr_nhwc_split_to_nchw': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'hvr_nchw_to_nhwc_f32': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'hvr_nhwc_to_nchw_f32': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'hvr_igemm': (c_int, [ctypes.POINTER(HvrIGemm), c_vp]),
    'hvr_igemm_check': (c_int, [ctypes.POINTER(HvrIGemm), c_vp]),
    'hvr_debug_force_bn': (c_int, [c_int]),
    'hvr_debug_roi_variant': (c_int, [c_int]),
    'hvr_im2col_stem': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp]),
    'hvr_maxpool3x3s2_split': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp]),
    'hvr_roi_align_fwd': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_int,
                                  c_vp, c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    'hvr_roi_align_fast_workspace_bytes': (c_sz, [c_int, c_int]),
    'hvr_roi_align_fwd_fast': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_int,
                                       c_vp, c_int, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp]),
    'hvr_nms_workspace_bytes': (c_sz, [c_int]),
    'hvr_nms': (c_int, [c_vp, c_int, c_f32, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'hvr_rpn_workspace_bytes': (c_sz, [c_int, c_int, c_int]),
    'hvr_rpn_proposals': (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_int, c_f32, c_f32,
                                  c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'hvr_det_workspace_bytes': (c_sz, [c_int, c_int]),
    'hvr_det_postprocess': (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_f32), c_f32,
                                    c_f32, c_f32, c_int, c_f32, c_f32, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'hvr_det_batched_workspace_bytes': (c_sz, [c_int, c_int, c_int]),
    'hvr_det_postprocess_batched': (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int,
                                            ctypes.POINTER(c_f32), c_f32, c_f32, c_f32, c_int, c_f32, c_f32, c_int,
                                            c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'hvr_det_postprocess_batched_ex': (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int,
                                               ctypes.POINTER(c_f32), c_f32, c_f32, c_f32, c_int, c_f32, c_f32, c_int,
                                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'hvr_preprocess_u8': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_f32),
                                  ctypes.POINTER(c_f32), c_vp, c_vp]),
    'hvr_softmax_rows_split': (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_vp]),
    'hvr_softmax_rows_split_masked': (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int,
                                              c_vp]),
    'hvr_packed_rows': (c_sz, [c_int]),
    'hvr_packed_cols': (c_sz, [c_int]),
    'hvr_pack_conv_bn': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                 c_vp]),
    'hvr_pack_linear': (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'hvr_linear_fwd': (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_i64, c_int, c_f32,
                               c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    'hvr_conv_fwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp,
                             c_int, c_vp, c_vp, c_vp, c_vp]),
    'hvr_relation_workspace_bytes': (c_sz, [c_int, c_int, c_int]),
    'hvr_relation_fwd': (c_int, [ctypes.POINTER(HvrRelationWeights), c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_int,
                                 c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp]),
    'hvr_window_rois': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    'hvr_gather_rows_split': (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_int, c_int, c_i64,
                                      c_int, c_int, c_vp]),
    'hvr_support_index': (c_int, [c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'hvr_video_descriptor_workspace_bytes': (c_sz, [c_int, c_int, c_int]),
    'hvr_video_descriptor': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    'hvr_support_select': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
}

_lib = None


class HvrError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises HvrError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HvrError('%s is missing: run `python -m hvrnet_b200.build` (there is no CPU or torch fallback)'
                           % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        L = lib()
        raise HvrError('%s failed: %s (code %d, cuda error %d)' % (what, L.hvr_strerror(rc).decode(), rc,
                                                                    L.hvr_last_cuda_error()))


def launch_count():
    return int(lib().hvr_launch_count())
