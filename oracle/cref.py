"""ctypes front-end of the C restatement (oracle/c/hvr_oracle.c).  Test infrastructure only."""
import ctypes

import numpy as np
import torch

from . import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build.build_c())
        _lib.oracle_roi_align_fwd.restype = ctypes.c_int
        _lib.oracle_nms.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def roi_align(feat, rois, out_size=7, spatial_scale=1 / 16., sample_num=2, feat_nhwc=False, out_nhwc=False):
    """feat (B,C,H,W) [or (B,H,W,C) if feat_nhwc] fp32 torch CPU tensor; rois (n,5)."""
    f = np.ascontiguousarray(feat.detach().cpu().numpy(), dtype=np.float32)
    r = np.ascontiguousarray(rois.detach().cpu().numpy(), dtype=np.float32)
    if feat_nhwc:
        B, H, W, C = f.shape
    else:
        B, C, H, W = f.shape
    n = r.shape[0]
    shape = (n, out_size, out_size, C) if out_nhwc else (n, C, out_size, out_size)
    out = np.zeros(shape, dtype=np.float32)
    rc = lib().oracle_roi_align_fwd(_fp(f), int(feat_nhwc), _fp(r), n, B, C, H, W, out_size, out_size,
                                    ctypes.c_float(spatial_scale), int(sample_num), _fp(out), int(out_nhwc))
    if rc != 0:
        raise ValueError('oracle_roi_align_fwd: roi batch index out of range')
    return torch.from_numpy(out)


def nms(dets, iou_thr, strict_gt=True, max_keep=0, ascending=True):
    d = np.ascontiguousarray(dets.detach().cpu().numpy(), dtype=np.float32)
    n = d.shape[0]
    keep = np.zeros(max(n, 1), dtype=np.int64)
    k = lib().oracle_nms(_fp(d), n, ctypes.c_float(iou_thr), int(strict_gt), int(max_keep), int(ascending),
                         keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return torch.from_numpy(keep[:k].copy())
