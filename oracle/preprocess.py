"""CPU oracle of the test-time image pipeline (TEST INFRASTRUCTURE ONLY) - SURVEY.md 8f N2.

Restates, in numpy integer arithmetic, what the reference's test pipeline does to a frame
(configs/faster_rcnn_r101_hrnmp_c5.py:193-201, mmdet/datasets/pipelines/transforms.py:111-125,
240-322, formating.py:48-56):

  Resize(img_scale=(1000,600), keep_ratio=True)  -> mmcv.imrescale -> cv2.resize(INTER_LINEAR)
  Normalize(mean, std, to_rgb=False)             -> (float32(img) - mean) / std
  Pad(size_divisor=16)                           -> zeros bottom/right
  ImageToTensor                                  -> HWC -> CHW

mmcv (0.2.x, un-vendored) only computes the target size and calls OpenCV; the arithmetic is
OpenCV's 8-bit bilinear resize (modules/imgproc/src/resize.cpp: fixed-point coefficients
cvRound(w * 2048), horizontal pass in int32, vertical pass
(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2).  It is pinned against cv2
itself (opencv-python 4.13 in this image) in tests/test_oracle_golden.py.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def rescale_size(h, w, scale=(1000, 600)):
    """mmcv.imrescale's target size: keep ratio, long edge <= max(scale), short edge <= min(scale)."""
    max_long, max_short = max(scale), min(scale)
    f = min(max_long / max(h, w), max_short / min(h, w))
    return int(h * float(f) + 0.5), int(w * float(f) + 0.5), f


def _coeffs(src, dst, clamp_weight):
    """OpenCV's per-axis source indices and fixed-point weights.  Along x the fraction is forced
    to 0 when the left tap falls outside (resize.cpp: `if (sx < 0) fx = 0, sx = 0`); along y the
    fraction is kept and only the ROW INDICES are clipped (`clip(sy + k, 0, height)`), which
    rounds differently on the first / last output rows when upscaling."""
    scale = 1.0 / (float(dst) / float(src))                    # cv::resize: scale_x = 1. / inv_scale_x
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weight:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)              # cvRound(float): round half to even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)
    s0 = np.clip(s, 0, src - 1)
    s1 = np.clip(s + 1, 0, src - 1)
    return s0, s1, w0, w1


def resize_bilinear_u8(img, new_h, new_w):
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    h, w = img.shape[:2]
    x0, x1, a0, a1 = _coeffs(w, new_w, True)
    y0, y1, b0, b1 = _coeffs(h, new_h, False)
    src = img.astype(np.int64)
    hor = src[:, x0] * a0[None, :, None] + src[:, x1] * a1[None, :, None]          # [h, new_w, c] int
    s0, s1 = hor[y0], hor[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def preprocess(img, scale=(1000, 600), mean=(103.06, 115.90, 123.15), std=(1.0, 1.0, 1.0), size_divisor=16):
    """uint8 HWC BGR frame -> (float32 [3, Hp, Wp], img_meta)."""
    h, w = img.shape[:2]
    nh, nw, f = rescale_size(h, w, scale)
    r = resize_bilinear_u8(img, nh, nw)
    x = (r.astype(np.float32) - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    ph = (nh + size_divisor - 1) // size_divisor * size_divisor
    pw = (nw + size_divisor - 1) // size_divisor * size_divisor
    out = np.zeros((3, ph, pw), dtype=np.float32)
    out[:, :nh, :nw] = x.transpose(2, 0, 1)
    meta = dict(ori_shape=(h, w, 3), img_shape=(nh, nw, 3), pad_shape=(ph, pw, 3), scale_factor=f, flip=False)
    return out, meta
