// Host build of the RoIAlign row-walk core (hvrnet_b200/csrc/roi_align_walk.cuh), for
// tests/test_host.py: the same make_axis_sample / roi_row_walk code the CUDA kernel runs, driven
// by a plain loop over (RoI, output row, 4-channel group), so the reuse logic can be checked bit
// for bit against the C oracle without a GPU.  Test infrastructure only.
// g++ -O2 -ffp-contract=off -shared -fPIC
#include <stddef.h>
#include <math.h>

#include "../../hvrnet_b200/csrc/roi_align_walk.cuh"

namespace {
struct Load {
  const float* base; int WC, C;
  hvr_f4 operator()(int row, int col) const {
    const float* p = base + (size_t)row * WC + (size_t)col * C;
    hvr_f4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3];
    return v;
  }
};
struct Store {
  float* out; int C;
  void operator()(int q, const hvr_f4& a) const {
    float* p = out + (size_t)q * C;
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;
  }
};
}  // namespace

// feat NHWC, out (n, ph, pw, C); geometry expressions as in roi_align.cu (roi_geom + the walk kernel).
// Returns the total number of 4-channel pixel loads issued (the per-bin kernel issues 16 per output vector).
extern "C" long long walk_roi_align(const float* feat, const float* rois, int n_rois, int n_imgs, int C, int H, int W,
                                    int ph, int pw, float scale, float* out) {
  long long loads = 0;
  AxisSample ys[64], xs[64];
  for (int n = 0; n < n_rois; ++n) {
    const float* r = rois + (size_t)n * 5;
    int b = (int)r[0];
    b = b < 0 ? 0 : (b >= n_imgs ? n_imgs - 1 : b);
    const float sw = r[1] * scale, sh = r[2] * scale;
    const float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
    const float rw = fmaxf(ew - sw, 0.0f), rh = fmaxf(eh - sh, 0.0f);
    const float bh = rh / (float)ph, bw = rw / (float)pw;
    for (int i = 0; i < 2 * ph; ++i)
      ys[i] = make_axis_sample(sh + (float)(i >> 1) * bh + ((float)(i & 1) + 0.5f) * bh / 2.0f, H);
    for (int j = 0; j < 2 * pw; ++j)
      xs[j] = make_axis_sample(sw + (float)(j >> 1) * bw + ((float)(j & 1) + 0.5f) * bw / 2.0f, W);
    const float* fm = feat + (size_t)b * H * W * C;
    for (int p = 0; p < ph; ++p)
      for (int c4 = 0; c4 < C / 4; ++c4) {
        Load ld{fm + c4 * 4, W * C, C};
        Store st{out + (((size_t)n * ph + p) * pw) * C + c4 * 4, C};
        loads += roi_row_walk(ys + 2 * p, xs, pw, ld, st);
      }
  }
  return loads;
}
