/* CPU oracle, C restatement of the two native ops on the HVRNet inference hot path.
 * TEST INFRASTRUCTURE ONLY - never linked or loaded by the product (hvrnet_b200).
 *
 *   oracle_roi_align_fwd : mmdet/ops/roi_align/src/roi_align_kernel.cu:16-118
 *   oracle_nms           : mmdet/ops/nms/src/nms_kernel.cu:14-22,57-63,116-135 (strict >)
 *                          mmdet/ops/nms/src/nms_cpu.cpp:34-57              (>=)
 *
 * Compile with -ffp-contract=off: every product and sum is rounded separately
 * (strict IEEE fp32), which is the arithmetic contract the CUDA kernels follow
 * with __fmul_rn/__fadd_rn.  RoIAlign forward values are pinned on the GPU box against the
 * reference's own CUDA kernel compiled unmodified with -fmad=false (bit for bit;
 * tests/test_gpu_kernels.py::test_roi_align_vs_reference_cuda_op); NMS is pinned by
 * nms_wrapper.py:25-35, by the reference's own nms_cpu.cpp and by its CUDA op, both built into
 * oracle/_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* roi_align_kernel.cu:16-61.  `data` points at one (H,W) plane with element stride
 * `es` (es=1 for NCHW planes, es=C for NHWC). */
static float bilinear(const float *data, int es, int H, int W, float y, float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.0f;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  float ly = y - (float)yl, lx = x - (float)xl;
  float hy = 1.0f - ly, hx = 1.0f - lx;
  float lt = data[(size_t)(yl * W + xl) * es], rt = data[(size_t)(yl * W + xh) * es];
  float lb = data[(size_t)(yh * W + xl) * es], rb = data[(size_t)(yh * W + xh) * es];
  float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * lt + w2 * rt + w3 * lb + w4 * rb;
}

/* feat: NCHW (feat_nhwc=0) or NHWC (feat_nhwc=1); rois (n,5) = [batch,x1,y1,x2,y2];
 * out: (n,C,ph,pw) (out_nhwc=0) or (n,ph,pw,C) (out_nhwc=1). */
int oracle_roi_align_fwd(const float *feat, int feat_nhwc, const float *rois, int n_rois, int n_imgs,
                         int C, int H, int W, int ph, int pw, float scale, int sample_num,
                         float *out, int out_nhwc) {
  for (int n = 0; n < n_rois; ++n) {
    const float *r = rois + (size_t)n * 5;
    int b = (int)r[0];
    if (b < 0 || b >= n_imgs) return -1;
    float sw_ = r[1] * scale, sh_ = r[2] * scale;
    float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
    float rw = fmaxf(ew - sw_, 0.0f), rh = fmaxf(eh - sh_, 0.0f);
    float bh = rh / (float)ph, bw = rw / (float)pw;
    int nh = sample_num > 0 ? sample_num : (int)ceilf(rh / (float)ph);
    int nw = sample_num > 0 ? sample_num : (int)ceilf(rw / (float)pw);
    for (int c = 0; c < C; ++c) {
      const float *plane = feat_nhwc ? feat + (size_t)b * H * W * C + c
                                     : feat + ((size_t)b * C + c) * H * W;
      int es = feat_nhwc ? C : 1;
      for (int p = 0; p < ph; ++p)
        for (int q = 0; q < pw; ++q) {
          float acc = 0.0f;
          for (int iy = 0; iy < nh; ++iy) {
            float y = sh_ + (float)p * bh + ((float)iy + 0.5f) * bh / (float)nh;
            for (int ix = 0; ix < nw; ++ix) {
              float x = sw_ + (float)q * bw + ((float)ix + 0.5f) * bw / (float)nw;
              acc += bilinear(plane, es, H, W, y, x);
            }
          }
          acc /= (float)(nh * nw);
          size_t o = out_nhwc ? (((size_t)n * ph + p) * pw + q) * C + c
                              : (((size_t)n * C + c) * ph + p) * pw + q;
          out[o] = acc;
        }
    }
  }
  return 0;
}

static float iou(const float *a, const float *b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float w = fmaxf(right - left + 1.0f, 0.0f), h = fmaxf(bottom - top + 1.0f, 0.0f);
  float inter = w * h;
  float sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f);
  float sb = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
  return inter / (sa + sb - inter);
}

typedef struct { float key; int idx; } kv_t;
static int cmp_desc(const void *pa, const void *pb) {
  const kv_t *a = (const kv_t *)pa, *b = (const kv_t *)pb;
  if (a->key > b->key) return -1;
  if (a->key < b->key) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

/* dets (n,5); total order (score desc, index asc); keep[] receives the kept original
 * indices in ASCENDING order (ascending=1, the reference's return order) or in score
 * order (ascending=0); stops after max_keep survivors when max_keep > 0. */
int oracle_nms(const float *dets, int n, float thr, int strict_gt, int max_keep, int ascending,
               int64_t *keep) {
  if (n <= 0) return 0;
  kv_t *ord = (kv_t *)malloc(sizeof(kv_t) * (size_t)n);
  uint8_t *rm = (uint8_t *)calloc((size_t)n, 1);
  for (int i = 0; i < n; ++i) { ord[i].key = dets[(size_t)i * 5 + 4]; ord[i].idx = i; }
  qsort(ord, (size_t)n, sizeof(kv_t), cmp_desc);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (rm[i]) continue;
    keep[k++] = ord[i].idx;
    if (max_keep > 0 && k >= max_keep) break;
    const float *a = dets + (size_t)ord[i].idx * 5;
    for (int j = i + 1; j < n; ++j) {
      if (rm[j]) continue;
      float v = iou(a, dets + (size_t)ord[j].idx * 5);
      if (strict_gt ? (v > thr) : (v >= thr)) rm[j] = 1;
    }
  }
  if (ascending) {
    for (int i = 1; i < k; ++i) {          /* insertion sort: k <= a few thousand */
      int64_t v = keep[i]; int j = i - 1;
      while (j >= 0 && keep[j] > v) { keep[j + 1] = keep[j]; --j; }
      keep[j + 1] = v;
    }
  }
  free(ord); free(rm);
  return k;
}
