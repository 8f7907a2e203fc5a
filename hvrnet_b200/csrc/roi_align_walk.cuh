// Row-walk core of the RoIAlign forward kernel (sample_num == 2): one work item = one output
// row p of one RoI for one group of 4 channels.  The item walks its 2*pw x-samples left to
// right and keeps the two feature-map columns of the current bilinear cell, for the (up to) four
// feature-map rows its two y-samples touch, in registers.  A tap is loaded only when its
// (row, column) is not already held: the reference (mmdet/ops/roi_align/src/
// roi_align_kernel.cu:16-61, :86-112) issues 16 loads per output element whatever the RoI size;
// here a bin narrower than two feature pixels reuses the columns of the previous sample, and
// y-samples that fall into the same cell share their rows.  Arithmetic is untouched - the same
// values enter the same products and sums in the same order - so the result is bit-identical.
//
// Host- and device-compilable (tests/test_host.py builds it with g++ against the C oracle).
#pragma once
#ifdef __CUDACC__
#define HVR_HD __host__ __device__ __forceinline__
typedef float4 hvr_f4;
#else
#define HVR_HD inline
struct hvr_f4 { float x, y, z, w; };
#endif

// One sample position along one axis: the two taps and their weights.  lo < 0: the sample lies
// outside [-1, size] and the whole 2-D sample contributes 0 (roi_align_kernel.cu:21-25).
struct AxisSample {
  int lo, hi;
  float l, h;   // weight of hi / of lo
};

// roi_align_kernel.cu:27-51, one axis at a time (the clamps of y and x are independent).
HVR_HD AxisSample make_axis_sample(float v, int size) {
  AxisSample s;
  if (v < -1.0f || v > (float)size) { s.lo = s.hi = -1; s.l = s.h = 0.f; return s; }
  if (v <= 0) v = 0;
  int lo = (int)v, hi;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
  s.lo = lo; s.hi = hi;
  s.l = v - (float)lo;
  s.h = 1.0f - s.l;
  return s;
}

HVR_HD hvr_f4 f4_zero() { hvr_f4 z; z.x = z.y = z.z = z.w = 0.f; return z; }

// w1*lt + w2*rt + w3*lb + w4*rb, left to right, every operation rounded (callers are built with
// -fmad=false / -ffp-contract=off).
HVR_HD hvr_f4 bilerp4(float w1, float w2, float w3, float w4, const hvr_f4& lt, const hvr_f4& rt,
                      const hvr_f4& lb, const hvr_f4& rb) {
  hvr_f4 v;
  v.x = ((w1 * lt.x + w2 * rt.x) + w3 * lb.x) + w4 * rb.x;
  v.y = ((w1 * lt.y + w2 * rt.y) + w3 * lb.y) + w4 * rb.y;
  v.z = ((w1 * lt.z + w2 * rt.z) + w3 * lb.z) + w4 * rb.z;
  v.w = ((w1 * lt.w + w2 * rt.w) + w3 * lb.w) + w4 * rb.w;
  return v;
}
HVR_HD void f4_acc(hvr_f4& a, const hvr_f4& v) { a.x = a.x + v.x; a.y = a.y + v.y; a.z = a.z + v.z; a.w = a.w + v.w; }

// ys: the 2 y-samples of output row p; xs: the 2*pw x-samples of the RoI.
// ld(row, col) -> the 4 channels of feature pixel (row, col); st(q, value) stores bin (p, q).
// Returns the number of pixel loads issued (used by the host test and the traffic model).
template <class Load, class Store>
HVR_HD int roi_row_walk(const AxisSample* ys, const AxisSample* xs, int pw, Load ld, Store st) {
  const AxisSample y0 = ys[0], y1 = ys[1];
  // row slots: 0 = lo(y0), 1 = hi(y0), 2 = lo(y1), 3 = hi(y1); a slot whose row is already held by an
  // earlier slot copies it instead of loading
  const int r0 = y0.lo, r1 = y0.hi, r2 = y1.lo, r3 = y1.hi;
  const bool own0 = r0 >= 0;
  const bool own1 = r1 >= 0 && r1 != r0;
  const bool own2 = r2 >= 0 && r2 != r0 && r2 != r1;
  const bool own3 = r3 >= 0 && r3 != r0 && r3 != r1 && r3 != r2;
  int loads = 0;

  hvr_f4 L0 = f4_zero(), L1 = L0, L2 = L0, L3 = L0;   // column cl of the four row slots
  hvr_f4 H0 = L0, H1 = L0, H2 = L0, H3 = L0;           // column ch
  int cl = -2, ch = -2;

#define HVR_FETCH_COL(col, A0, A1, A2, A3)                                         \
  do {                                                                             \
    if (own0) { A0 = ld(r0, (col)); ++loads; }                                     \
    if (own1) { A1 = ld(r1, (col)); ++loads; }                                     \
    if (own2) { A2 = ld(r2, (col)); ++loads; }                                     \
    if (own3) { A3 = ld(r3, (col)); ++loads; }                                     \
    if (!own1) A1 = A0;                                                            \
    if (!own2) A2 = (r2 == r0) ? A0 : A1;                                          \
    if (!own3) A3 = (r3 == r2) ? A2 : ((r3 == r1) ? A1 : A0);                      \
  } while (0)

  for (int q = 0; q < pw; ++q) {
    hvr_f4 acc = f4_zero();
    hvr_f4 v10 = f4_zero();   // sample (iy 1, ix 0) waits until both iy = 0 samples are summed
    bool ok10 = false;
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      const AxisSample x = xs[q * 2 + ix];
      if (x.lo >= 0) {
        if (x.lo == cl) {
        } else if (x.lo == ch) {
          L0 = H0; L1 = H1; L2 = H2; L3 = H3;
        } else {
          HVR_FETCH_COL(x.lo, L0, L1, L2, L3);
        }
        cl = x.lo;
        if (x.hi == x.lo) {
          H0 = L0; H1 = L1; H2 = L2; H3 = L3;
        } else if (x.hi != ch) {
          HVR_FETCH_COL(x.hi, H0, H1, H2, H3);
        }
        ch = x.hi;
        if (y0.lo >= 0) {   // sample order iy-outer / ix-inner: (0,0), (0,1), (1,0), (1,1)
          const hvr_f4 v = bilerp4(y0.h * x.h, y0.h * x.l, y0.l * x.h, y0.l * x.l, L0, H0, L1, H1);
          f4_acc(acc, v);
        }
      }
      if (ix == 1 && ok10) f4_acc(acc, v10);
      if (x.lo >= 0 && y1.lo >= 0) {
        const hvr_f4 v = bilerp4(y1.h * x.h, y1.h * x.l, y1.l * x.h, y1.l * x.l, L2, H2, L3, H3);
        if (ix == 0) { v10 = v; ok10 = true; } else { f4_acc(acc, v); }
      }
    }
    acc.x = acc.x / 4.0f; acc.y = acc.y / 4.0f; acc.z = acc.z / 4.0f; acc.w = acc.w / 4.0f;
    st(q, acc);
  }
#undef HVR_FETCH_COL
  return loads;
}
