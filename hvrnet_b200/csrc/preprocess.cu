// Test-time image pipeline on the GPU: uint8 HWC BGR frame -> keep-ratio
// bilinear resize -> mean/std normalise -> zero pad to /16 -> CHW fp32, one kernel, one pass.
// Replaces the CPU DataLoader path mmcv.imrescale (cv2.resize INTER_LINEAR) + mmcv.imnormalize +
// mmcv.impad_to_multiple + ImageToTensor (mmdet/datasets/pipelines/transforms.py:111-125,
// 240-322; formating.py:48-56).  The resize reproduces OpenCV's 8-bit fixed-point arithmetic
// (11-bit coefficients, int32 horizontal pass, two-shift vertical pass) bit for bit; see
// oracle/preprocess.py for the restatement that is pinned against cv2.
#include "common.cuh"

namespace {

struct Axis {
  int s0, s1, w0, w1;
};
// resize.cpp: f = (float)((d + 0.5) * scale - 0.5); s = floor(f); f -= s; x axis clamps the
// fraction, y axis only clips the row indices.
__device__ __forceinline__ Axis axis_coeffs(int d, int src, double scale, bool clamp_weight) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp_weight) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Axis a;
  a.w1 = __float2int_rn(f * 2048.0f);              // cvRound: round half to even
  a.w0 = __float2int_rn((1.0f - f) * 2048.0f);
  a.s0 = min(max(s, 0), src - 1);
  a.s1 = min(max(s + 1, 0), src - 1);
  return a;
}

__global__ void preprocess_kernel(const unsigned char* __restrict__ img, int h, int w, int nh, int nw, int ph, int pw,
                                  double scale_y, double scale_x, float m0, float m1, float m2, float d0, float d1,
                                  float d2, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= pw) return;
  const size_t plane = (size_t)ph * pw;
  float* o = out + (size_t)y * pw + x;
  if (x >= nw || y >= nh) {                         // Pad(size_divisor): zeros AFTER normalisation
    o[0] = 0.f; o[plane] = 0.f; o[2 * plane] = 0.f;
    return;
  }
  const Axis ax = axis_coeffs(x, w, scale_x, true);
  const Axis ay = axis_coeffs(y, h, scale_y, false);
  const unsigned char* r0 = img + (size_t)ay.s0 * w * 3;
  const unsigned char* r1 = img + (size_t)ay.s1 * w * 3;
  const float mean[3] = {m0, m1, m2}, sd[3] = {d0, d1, d2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)r0[ax.s0 * 3 + c] * ax.w0 + (int)r0[ax.s1 * 3 + c] * ax.w1;
    const int h1 = (int)r1[ax.s0 * 3 + c] * ax.w0 + (int)r1[ax.s1 * 3 + c] * ax.w1;
    const int v = (((ay.w0 * (h0 >> 4)) >> 16) + ((ay.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
    const int u = v < 0 ? 0 : (v > 255 ? 255 : v);
    o[c * plane] = __fdiv_rn(__fsub_rn((float)u, mean[c]), sd[c]);
  }
}

}  // namespace

extern "C" int hvr_preprocess_u8(const uint8_t* img, int h, int w, int new_h, int new_w, int pad_h, int pad_w,
                                 const float* mean3_host, const float* std3_host, float* out, void* stream) {
  if (!img || !out || !mean3_host || !std3_host || h < 1 || w < 1 || new_h < 1 || new_w < 1 || pad_h < new_h ||
      pad_w < new_w)
    return HVR_ERR_ARG;
  const double scale_y = 1.0 / ((double)new_h / (double)h);     // cv::resize: scale = 1. / inv_scale
  const double scale_x = 1.0 / ((double)new_w / (double)w);
  dim3 grid(hvr_cdiv(pad_w, 128), pad_h);
  preprocess_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, h, w, new_h, new_w, pad_h, pad_w, scale_y, scale_x, mean3_host[0], mean3_host[1], mean3_host[2],
      std3_host[0], std3_host[1], std3_host[2], out);
  HVR_LAUNCHED();
  return HVR_OK;
}
