// Per-bin core of the RoIAlign forward kernel for sample_num == 2: one work item = one output bin of
// one RoI for one group of 4 channels.  The reference (mmdet/ops/roi_align/src/roi_align_kernel.cu:
// 16-61, :86-112) loads 4 taps for each of the 4 samples of a bin.  When a bin is less than two feature
// pixels wide or high, its samples fall into the same or into adjacent bilinear cells and share taps;
// those are taken from registers instead of being loaded again - L1 data-pipe wavefronts, not HBM
// bytes, are what bounds the kernel (profiles/r01q_roi_align_ncu_summary.txt).  Arithmetic is untouched:
// the same values enter the same products and sums in the same order, so the result is bit-identical to
// the straightforward evaluation.
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

// ---- tap reuse inside a bin ------------------------------------------------------------------
// A tap offset is rowoff(y tap) + coloff(x tap), so the 16 taps of a bin are a grid of (2..4 distinct
// rows) x (2..4 distinct columns).  Along each axis the second sample relates to the first as
//   kSame (0): same cell            -> the axis contributes 2 grid lines
//   kAdj  (1): lo(1) == hi(0)       -> 3 grid lines
//   kFar  (2): anything else        -> 4 grid lines (no reuse, the reference's 16 loads when both are far)
// The relation is found by comparing offsets, so every substituted tap has the address of the tap it
// replaces (clamped border cells included).
enum { kSame = 0, kAdj = 1, kFar = 2 };

// t[iy*2+ix], all four valid.  Returns ry * 3 + rx.
HVR_HD int roi_bin_code(const Tap* t) {
  const int ry = (t[2].o0 == t[0].o0 && t[2].o2 == t[0].o2) ? kSame : (t[2].o0 == t[0].o2 ? kAdj : kFar);
  const int rx = (t[1].o0 == t[0].o0 && t[1].o1 == t[0].o1) ? kSame : (t[1].o0 == t[0].o1 ? kAdj : kFar);
  return ry * 3 + rx;
}

// Grid line of tap `tap` (0 = lo, 1 = hi) of sample `smp` (0, 1) along an axis with relation REL.
template <int REL>
HVR_HD constexpr int grid_line(int smp, int tap) {
  return smp == 0 ? tap : (REL == kSame ? tap : (REL == kAdj ? 1 + tap : 2 + tap));
}

// One bin, relations known at compile time: every grid point is loaded once, when its first sample needs
// it.  ((((0 + s00) + s01) + s10) + s11) * 0.25 (roi_align_kernel.cu:100-112: iy outer, ix inner; the
// division by 4 is exact as a product).  *loads counts the pixel loads issued (host test).
template <int RY, int RX, class Load>
HVR_HD hvr_f4 roi_bin_grid(const Tap* t, Load ld, int* loads) {
  hvr_f4 G[4][4];
  bool have[4][4] = {{false, false, false, false}, {false, false, false, false},
                     {false, false, false, false}, {false, false, false, false}};
  hvr_f4 acc;
  acc.x = acc.y = acc.z = acc.w = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int iy = s >> 1, ix = s & 1;
    const uint32_t off[4] = {t[s].o0, t[s].o1, t[s].o2, t[s].o3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = grid_line<RY>(iy, k >> 1), c = grid_line<RX>(ix, k & 1);
      if (!have[r][c]) {
        G[r][c] = ld(off[k]);
        have[r][c] = true;
        if (loads) ++*loads;
      }
    }
    const int r0 = grid_line<RY>(iy, 0), r1 = grid_line<RY>(iy, 1);
    const int c0 = grid_line<RX>(ix, 0), c1 = grid_line<RX>(ix, 1);
    const hvr_f4 v = bilerp4(t[s], G[r0][c0], G[r0][c1], G[r1][c0], G[r1][c1]);
    acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
  }
  acc.x = acc.x * 0.25f; acc.y = acc.y * 0.25f; acc.z = acc.z * 0.25f; acc.w = acc.w * 0.25f;
  return acc;
}

// code = roi_bin_code(t) (warp-uniform in the kernel: the lanes of a warp work on the same bin).
template <class Load>
HVR_HD hvr_f4 roi_bin_sn2(int code, const Tap* t, Load ld, int* loads) {
  switch (code) {
    case 0: return roi_bin_grid<kSame, kSame>(t, ld, loads);
    case 1: return roi_bin_grid<kSame, kAdj>(t, ld, loads);
    case 2: return roi_bin_grid<kSame, kFar>(t, ld, loads);
    case 3: return roi_bin_grid<kAdj, kSame>(t, ld, loads);
    case 4: return roi_bin_grid<kAdj, kAdj>(t, ld, loads);
    case 5: return roi_bin_grid<kAdj, kFar>(t, ld, loads);
    case 6: return roi_bin_grid<kFar, kSame>(t, ld, loads);
    case 7: return roi_bin_grid<kFar, kAdj>(t, ld, loads);
    default: return roi_bin_grid<kFar, kFar>(t, ld, loads);
  }
}
