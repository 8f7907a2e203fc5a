"""GPU test-time pipeline: Resize(1000x600, keep ratio) -> Normalize -> Pad(16)
-> CHW in one kernel (hvr_preprocess_u8).  Mirrors the `test_pipeline` of the reference configs
(configs/faster_rcnn_r101_hrnmp_c5.py:193-201) and returns the `img_meta` keys the detectors read."""
import ctypes

import torch

from . import _lib
from ._lib import check


def rescale_size(h, w, scale=(1000, 600)):
    """mmcv.imrescale's target size (keep ratio; long edge <= max(scale), short edge <= min(scale))."""
    max_long, max_short = max(scale), min(scale)
    f = min(max_long / max(h, w), max_short / min(h, w))
    return int(h * float(f) + 0.5), int(w * float(f) + 0.5), f


def preprocess(img, scale=(1000, 600), mean=(103.06, 115.90, 123.15), std=(1.0, 1.0, 1.0), size_divisor=16):
    """img: uint8 HWC BGR CUDA tensor [h,w,3] -> (fp32 [1,3,Hp,Wp], img_meta)."""
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise _lib.HvrError('preprocess takes a uint8 HWC CUDA tensor (no CPU path)')
    img = img.contiguous()
    h, w = int(img.shape[0]), int(img.shape[1])
    nh, nw, f = rescale_size(h, w, scale)
    ph = (nh + size_divisor - 1) // size_divisor * size_divisor
    pw = (nw + size_divisor - 1) // size_divisor * size_divisor
    out = torch.empty((1, 3, ph, pw), dtype=torch.float32, device=img.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    check(_lib.lib().hvr_preprocess_u8(ctypes.c_void_p(img.data_ptr()), h, w, nh, nw, ph, pw, m, s,
                                       ctypes.c_void_p(out.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'hvr_preprocess_u8')
    meta = dict(ori_shape=(h, w, 3), img_shape=(nh, nw, 3), pad_shape=(ph, pw, 3), scale_factor=f, flip=False)
    return out, meta
