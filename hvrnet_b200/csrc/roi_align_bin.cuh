// Per-bin core of the RoIAlign forward kernel for sample_num == 2: one work item = one output bin of
// one RoI for one group of 4 channels.  The reference (mmdet/ops/roi_align/src/roi_align_kernel.cu:
// 16-61, :86-112) loads 4 taps for each of the 4 samples of a bin.  The two samples of a bin that share
// their x position (iy = 0, 1) use the same two feature-map columns; when the bin is less than two
// feature pixels high they also fall into the same cell (all 4 taps shared) or into vertically adjacent
// cells (the lower taps of the first are the upper taps of the second).  Those taps are taken from
// registers instead of being loaded again - L1 data-pipe wavefronts, not HBM bytes, are what bounds the
// kernel.  Arithmetic is untouched: the same values enter the same products and sums in the same order,
// so the result is bit-identical to the straightforward evaluation.
//
// Host- and device-compilable (tests/test_host.py builds it with g++ against the C oracle).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define HVR_HD __host__ __device__ __forceinline__
typedef float4 hvr_f4;
#else
#define HVR_HD inline
struct hvr_f4 { float x, y, z, w; };
#endif

// One bilinear sample: BYTE offsets of the taps lt, rt, lb, rb inside the image's NHWC map (channel 0)
// and their weights.  A sample outside [-1, H] x [-1, W] (roi_align_kernel.cu:21-25: contributes 0) has
// o0 = kTapInvalid.
struct Tap {
  uint32_t o0, o1, o2, o3;
  float w1, w2, w3, w4;
};
constexpr uint32_t kTapInvalid = 0xffffffffu;

// roi_align_kernel.cu:16-61 -> offsets and weights instead of values.
HVR_HD Tap make_tap(float y, float x, int H, int W, int C) {
  Tap t;
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
    t.o0 = t.o1 = t.o2 = t.o3 = kTapInvalid;
    t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
    return t;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  const float ly = y - (float)yl, lx = x - (float)xl;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  t.o0 = (uint32_t)((yl * W + xl) * C) * 4u;
  t.o1 = (uint32_t)((yl * W + xh) * C) * 4u;
  t.o2 = (uint32_t)((yh * W + xl) * C) * 4u;
  t.o3 = (uint32_t)((yh * W + xh) * C) * 4u;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

// w1*lt + w2*rt + w3*lb + w4*rb, left to right, every operation rounded (callers are built with
// -fmad=false / -ffp-contract=off).
HVR_HD hvr_f4 bilerp4(const Tap& t, const hvr_f4& lt, const hvr_f4& rt, const hvr_f4& lb, const hvr_f4& rb) {
  hvr_f4 v;
  v.x = ((t.w1 * lt.x + t.w2 * rt.x) + t.w3 * lb.x) + t.w4 * rb.x;
  v.y = ((t.w1 * lt.y + t.w2 * rt.y) + t.w3 * lb.y) + t.w4 * rb.y;
  v.z = ((t.w1 * lt.z + t.w2 * rt.z) + t.w3 * lb.z) + t.w4 * rb.z;
  v.w = ((t.w1 * lt.w + t.w2 * rt.w) + t.w3 * lb.w) + t.w4 * rb.w;
  return v;
}

// The two samples (iy = 0: t0, iy = 1: t1) of one bin at the same x position; both valid.
// ld(byte offset) -> the 4 channels of that pixel.  *loads counts the pixel loads issued (host test).
template <class Load>
HVR_HD void roi_sample_column(const Tap& t0, const Tap& t1, Load ld, hvr_f4& v0, hvr_f4& v1, int* loads) {
  const hvr_f4 a = ld(t0.o0), b = ld(t0.o1), c = ld(t0.o2), d = ld(t0.o3);
  v0 = bilerp4(t0, a, b, c, d);
  if (t1.o0 == t0.o0 && t1.o2 == t0.o2) {            // same cell
    v1 = bilerp4(t1, a, b, c, d);
    if (loads) *loads += 4;
  } else if (t1.o0 == t0.o2) {                       // the cell below: its upper taps are held
    const hvr_f4 e = ld(t1.o2), f = ld(t1.o3);
    v1 = bilerp4(t1, c, d, e, f);
    if (loads) *loads += 6;
  } else {
    const hvr_f4 e = ld(t1.o0), f = ld(t1.o1), g = ld(t1.o2), h = ld(t1.o3);
    v1 = bilerp4(t1, e, f, g, h);
    if (loads) *loads += 8;
  }
}

// One bin whose 4 samples t[iy*2+ix] are all valid: ((((0 + s00) + s01) + s10) + s11) / 4
// (roi_align_kernel.cu:100-112: iy outer, ix inner; the division by 4 is exact as a product).
template <class Load>
HVR_HD hvr_f4 roi_bin_sn2(const Tap* t, Load ld, int* loads) {
  hvr_f4 v00, v01, v10, v11;
  roi_sample_column(t[0], t[2], ld, v00, v10, loads);
  roi_sample_column(t[1], t[3], ld, v01, v11, loads);
  hvr_f4 acc;
  acc.x = (((0.f + v00.x) + v01.x) + v10.x) + v11.x;
  acc.y = (((0.f + v00.y) + v01.y) + v10.y) + v11.y;
  acc.z = (((0.f + v00.z) + v01.z) + v10.z) + v11.z;
  acc.w = (((0.f + v00.w) + v01.w) + v10.w) + v11.w;
  acc.x = acc.x * 0.25f; acc.y = acc.y * 0.25f; acc.z = acc.z * 0.25f; acc.w = acc.w * 0.25f;
  return acc;
}
