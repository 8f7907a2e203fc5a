// Host build of the separable RoIAlign core (hvrnet_b200/csrc/roi_align_sep.cuh), for tests/test_host.py: the
// same axis_tap / col_taps_sn2 / row_tap_sn2 / roi_column_sep_sn2 code roi_align_sep_kernel runs, driven by a
// plain loop over (RoI, output column, 4-channel group), so the fast variant can be checked against the C
// oracle (tolerance 1e-5 relative, stated in the test) without a GPU.  Test infrastructure only.
// g++ -O2 -ffp-contract=off -shared -fPIC   (fmaf is explicit in the header, as in the kernel)
#include <math.h>
#include <stddef.h>

#include "../../hvrnet_b200/csrc/roi_align_sep.cuh"

namespace {
struct Load {
  const char* base;
  hvr_sf4 operator()(uint32_t off) const {
    const float* p = reinterpret_cast<const float*>(base + off);
    hvr_sf4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3];
    return v;
  }
};
}  // namespace

// feat NHWC, out (n, ph, pw, C); geometry expressions as in roi_align.cu (roi_geom + roi_align_sep_kernel).
// Returns the number of 4-channel pixel loads issued (the reference issues 16 per output vector).
extern "C" long long sep_roi_align(const float* feat, const float* rois, int n_rois, int n_imgs, int C, int H, int W,
                                   int ph, int pw, float scale, float* out) {
  long long total = 0;
  RowTap rows[32];
  if (2 * ph > 32) return -1;
  for (int n = 0; n < n_rois; ++n) {
    const float* r = rois + (size_t)n * 5;
    int b = (int)r[0];
    b = b < 0 ? 0 : (b >= n_imgs ? n_imgs - 1 : b);
    const float sw = r[1] * scale, sh = r[2] * scale;
    const float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
    const float rw = fmaxf(ew - sw, 0.0f), rh = fmaxf(eh - sh, 0.0f);
    const float bh = rh / (float)ph, bw = rw / (float)pw;
    for (int s = 0; s < 2 * ph; ++s) rows[s] = row_tap_sn2(sh, bh, s, H, (uint32_t)(W * C) * 4u);
    for (int q = 0; q < pw; ++q) {
      const ColTaps ct = col_taps_sn2(sw, bw, q, W, (uint32_t)C * 4u);
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const Load ld{reinterpret_cast<const char*>(feat + (size_t)b * H * W * C + c4 * 4)};
        int loads = 0;
        float* o = out + ((size_t)n * ph * pw + q) * C + c4 * 4;
        const size_t step = (size_t)pw * C;
        roi_column_sep_sn2(rows, ph, ct, ld,
                           [&](int p, const hvr_sf4& v) {
                             float* d = o + p * step;
                             d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                           },
                           &loads);
        total += loads;
      }
    }
  }
  return total;
}
