// Host build of the RoIAlign per-bin core (hvrnet_b200/csrc/roi_align_bin.cuh), for tests/test_host.py:
// the same make_tap / roi_bin_sn2 code the CUDA kernel runs, driven by a plain loop over (RoI, bin,
// 4-channel group) with the kernel's choice between the branch-free and the per-sample path, so the tap
// reuse can be checked bit for bit against the C oracle without a GPU.  Test infrastructure only.
// g++ -O2 -ffp-contract=off -shared -fPIC
#include <math.h>
#include <stddef.h>

#include "../../hvrnet_b200/csrc/roi_align_bin.cuh"

namespace {
struct Load {
  const char* base;
  hvr_f4 operator()(uint32_t off) const {
    const float* p = reinterpret_cast<const float*>(base + off);
    hvr_f4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3];
    return v;
  }
};
}  // namespace

// feat NHWC, out (n, ph, pw, C); geometry expressions as in roi_align.cu (roi_geom + roi_align_sn2_kernel).
// Returns the number of 4-channel pixel loads issued (the reference issues 16 per output vector).
extern "C" long long bin_roi_align(const float* feat, const float* rois, int n_rois, int n_imgs, int C, int H, int W,
                                   int ph, int pw, float scale, float* out) {
  long long total = 0;
  Tap taps[4 * 64];
  if (ph * pw > 64) return -1;
  for (int n = 0; n < n_rois; ++n) {
    const float* r = rois + (size_t)n * 5;
    int b = (int)r[0];
    b = b < 0 ? 0 : (b >= n_imgs ? n_imgs - 1 : b);
    const float sw = r[1] * scale, sh = r[2] * scale;
    const float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
    const float rw = fmaxf(ew - sw, 0.0f), rh = fmaxf(eh - sh, 0.0f);
    const float bh = rh / (float)ph, bw = rw / (float)pw;
    bool any_invalid = false;
    for (int i = 0; i < ph * pw * 4; ++i) {
      const int ix = i & 1, iy = (i >> 1) & 1, bin = i >> 2;
      const int q = bin % pw, p = bin / pw;
      const float y = sh + (float)p * bh + ((float)iy + 0.5f) * bh / 2.0f;
      const float x = sw + (float)q * bw + ((float)ix + 0.5f) * bw / 2.0f;
      taps[i] = make_tap(y, x, H, W, C);
      any_invalid |= taps[i].o0 == kTapInvalid;
    }
    for (int bin = 0; bin < ph * pw; ++bin)
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const Load ld{reinterpret_cast<const char*>(feat + (size_t)b * H * W * C + c4 * 4)};
        const Tap* tp = taps + bin * 4;
        hvr_f4 acc;
        int loads = 0;
        if (!any_invalid) {
          acc = roi_bin_sn2(tp, ld, &loads);
        } else {
          acc.x = acc.y = acc.z = acc.w = 0.f;
          for (int s = 0; s < 4; ++s) {
            if (tp[s].o0 == kTapInvalid) continue;
            const hvr_f4 v = bilerp4(tp[s], ld(tp[s].o0), ld(tp[s].o1), ld(tp[s].o2), ld(tp[s].o3));
            acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
            loads += 4;
          }
          acc.x = acc.x * 0.25f; acc.y = acc.y * 0.25f; acc.z = acc.z * 0.25f; acc.w = acc.w * 0.25f;
        }
        total += loads;
        float* o = out + ((size_t)n * ph * pw + bin) * C + c4 * 4;
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
      }
  }
  return total;
}
