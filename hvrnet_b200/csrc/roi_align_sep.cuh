// Separable evaluation of the RoIAlign forward for a regular sampling grid (the pipeline's fast variant).
//
// The reference (mmdet/ops/roi_align/src/roi_align_kernel.cu:16-61, :86-112) averages sn x sn bilinear
// samples per bin; sample (iy, ix) of bin (p, q) sits at (y_{p,iy}, x_{q,ix}) - the y position does not
// depend on q / ix and the x position does not depend on p / iy - and a bilinear sample is itself a product
// of a row interpolation and a column interpolation.  Hence, exactly (in real arithmetic),
//
//     out[p][q] = sum_r WY_p[r] * ( sum_c WX_q[c] * f[r][c] )
//
//     WY_p[r] = (1/sn) sum_iy valid(y_{p,iy}) * rowweight_{p,iy}(r)        (<= 2*sn rows per p)
//     WX_q[c] = (1/sn) sum_ix valid(x_{q,ix}) * colweight_{q,ix}(c)        (<= 2*sn columns per q)
//
// with the reference's validity rule (a sample outside [-1, H] x [-1, W] contributes 0; the rule is a
// conjunction of a y test and an x test, so it factors too) and its clamping at the map border.  A thread
// that owns (q, 4 channels) walks the rows of the RoI top to bottom, interpolates each distinct row ONCE
// along x (<= 4 merged column taps) and feeds it to the one or two samples that use it: ~6 pixel loads per
// output vector on the bench distribution instead of 16 (reference) / 11.8 (roi_align_sn2_kernel), and
// ~40 fused multiply-adds instead of 138 separately rounded operations - L1 wavefronts and issue slots are
// what bound this kernel, not HBM.
//
// The summation order differs from the reference's, so the result is NOT bit-identical to the strict
// kernels (roi_align.cu); it agrees with them to a few ulp of the largest term (tests: 1e-5 relative, the
// agreement the reference's own default FMA-contracted build shows against its -fmad=false build).
//
// Host- and device-compilable (tests/test_host.py builds the core with g++ against the C oracle).
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#define HVR_SEP_HD __host__ __device__ __forceinline__
typedef float4 hvr_sf4;
#else
#define HVR_SEP_HD inline
struct hvr_sf4 { float x, y, z, w; };
#endif

// One sample along one axis: indices of the two taps and their weights (already scaled by `scale`).
// i0 < 0: the sample is outside [-1, L] and contributes nothing (roi_align_kernel.cu:21-25).
struct AxisTap {
  int i0, i1;
  float w0, w1;
};

// roi_align_kernel.cu:27-52, one axis: clamp at 0, floor, clamp the upper tap at L-1.
HVR_SEP_HD AxisTap axis_tap(float v, int L, float scale) {
  AxisTap t;
  if (v < -1.0f || v > (float)L) {
    t.i0 = t.i1 = -1;
    t.w0 = t.w1 = 0.f;
    return t;
  }
  if (v <= 0) v = 0;
  int lo = (int)v, hi;
  if (lo >= L - 1) { hi = lo = L - 1; v = (float)lo; } else { hi = lo + 1; }
  const float l = v - (float)lo, h = 1.0f - l;
  t.i0 = lo; t.i1 = hi;
  t.w0 = h * scale; t.w1 = l * scale;
  return t;
}

// Up to 4 merged (index, weight) pairs of the two x samples of one output column.
struct ColTaps {
  int n;
  uint32_t off[4];   // byte offset of the column inside a map row (col * C * 4)
  float w[4];
};

HVR_SEP_HD void col_add(ColTaps& c, int idx, float w, uint32_t pitch) {
  if (idx < 0 || w == 0.f) return;
  const uint32_t o = (uint32_t)idx * pitch;
  if (c.n > 0 && c.off[0] == o) { c.w[0] += w; return; }
  if (c.n > 1 && c.off[1] == o) { c.w[1] += w; return; }
  if (c.n > 2 && c.off[2] == o) { c.w[2] += w; return; }
  if (c.n == 0) { c.off[0] = o; c.w[0] = w; }
  else if (c.n == 1) { c.off[1] = o; c.w[1] = w; }
  else if (c.n == 2) { c.off[2] = o; c.w[2] = w; }
  else { c.off[3] = o; c.w[3] = w; }
  ++c.n;
}

// Column taps of output column q (sample_num == 2): x_{q,ix} = start + q*bin + (ix + 0.5) * bin / 2.
HVR_SEP_HD ColTaps col_taps_sn2(float start, float bin, int q, int W, uint32_t pitch) {
  ColTaps c;
  c.n = 0;
  c.off[0] = c.off[1] = c.off[2] = c.off[3] = 0;
  c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0.f;
  for (int ix = 0; ix < 2; ++ix) {
    const float x = start + (float)q * bin + ((float)ix + 0.5f) * bin / 2.0f;
    const AxisTap t = axis_tap(x, W, 0.5f);
    col_add(c, t.i0, t.w0, pitch);
    col_add(c, t.i1, t.w1, pitch);
  }
  return c;
}

// One y sample: byte offsets of its two map rows (row * W * C * 4) and their weights; o0 = kRowInvalid when
// the sample is outside the map.
struct RowTap {
  uint32_t o0, o1;
  float w0, w1;
};
constexpr uint32_t kRowInvalid = 0xffffffffu;

HVR_SEP_HD RowTap row_tap_sn2(float start, float bin, int s, int H, uint32_t row_pitch) {
  const int p = s >> 1, iy = s & 1;
  const float y = start + (float)p * bin + ((float)iy + 0.5f) * bin / 2.0f;
  const AxisTap t = axis_tap(y, H, 0.5f);
  RowTap r;
  if (t.i0 < 0) {
    r.o0 = r.o1 = kRowInvalid;
    r.w0 = r.w1 = 0.f;
  } else {
    r.o0 = (uint32_t)t.i0 * row_pitch; r.o1 = (uint32_t)t.i1 * row_pitch;
    r.w0 = t.w0; r.w1 = t.w1;
  }
  return r;
}

HVR_SEP_HD hvr_sf4 sf4_zero() {
  hvr_sf4 v;
  v.x = v.y = v.z = v.w = 0.f;
  return v;
}
HVR_SEP_HD void sf4_fma(hvr_sf4& a, float w, const hvr_sf4& v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

// x interpolation of one map row for one output column: sum_c WX[c] * f[row][c]  (ld(byte offset) -> 4 channels)
template <class Load>
HVR_SEP_HD hvr_sf4 row_interp(const ColTaps& c, uint32_t row_off, Load ld, int* loads) {
  hvr_sf4 t = sf4_zero();
  if (c.n > 0) {
    const hvr_sf4 v0 = ld(row_off + c.off[0]);
    if (c.n > 1) {
      const hvr_sf4 v1 = ld(row_off + c.off[1]);
      if (c.n > 2) {
        const hvr_sf4 v2 = ld(row_off + c.off[2]);
        if (c.n > 3) {
          const hvr_sf4 v3 = ld(row_off + c.off[3]);
          t.x = c.w[3] * v3.x; t.y = c.w[3] * v3.y; t.z = c.w[3] * v3.z; t.w = c.w[3] * v3.w;
        }
        sf4_fma(t, c.w[2], v2);
      }
      sf4_fma(t, c.w[1], v1);
    }
    sf4_fma(t, c.w[0], v0);
    if (loads) *loads += c.n;
  }
  return t;
}

// All ph bins of output column q for one 4-channel group: rows[s] (s = p*2 + iy) are the RoI's y samples.
// emit(p, value) receives the ph results top to bottom.  The two most recently interpolated rows are kept
// (rows never decrease along s), so a row shared by consecutive samples / bins is interpolated once.
template <class Load, class Emit>
HVR_SEP_HD void roi_column_sep_sn2(const RowTap* rows, int ph, const ColTaps& c, Load ld, Emit emit, int* loads) {
  uint32_t ca = kRowInvalid, cb = kRowInvalid;
  hvr_sf4 ta = sf4_zero(), tb = sf4_zero();
  for (int p = 0; p < ph; ++p) {
    hvr_sf4 acc = sf4_zero();
    for (int iy = 0; iy < 2; ++iy) {
      const RowTap r = rows[p * 2 + iy];
      if (r.o0 == kRowInvalid) continue;
      // ta <- row o0
      if (r.o0 == cb) { ta = tb; ca = cb; }
      else if (r.o0 != ca) { ta = row_interp(c, r.o0, ld, loads); ca = r.o0; }
      sf4_fma(acc, r.w0, ta);
      if (r.w1 != 0.f) {
        if (r.o1 == ca) { sf4_fma(acc, r.w1, ta); }
        else {
          if (r.o1 != cb) { tb = row_interp(c, r.o1, ld, loads); cb = r.o1; }
          sf4_fma(acc, r.w1, tb);
        }
      }
    }
    emit(p, acc);
  }
}
