// Device-resident NMS, RPN proposal generation and detection post-processing.
//
// Replaces (all with host round trips in the reference):
//   nms_cuda            mmdet/ops/nms/src/nms_kernel.cu:24-68 (kernel), :71-136 (host: sort,
//                       blocking D2H of the bit mask, serial CPU scan)
//   get_bboxes_single   mmdet/models/anchor_heads/rpn_head.py:55-104 (per-frame Python loop)
//   grid_anchors        mmdet/core/anchor/anchor_generator.py:66-83
//   delta2bbox          mmdet/core/bbox/transforms.py:34-111
//   get_det_bboxes      mmdet/models/bbox_heads/hrnmp_bbox_head.py:1009-1052
//   multiclass_nms      mmdet/core/post_processing/bbox_nms.py:6-66  (30 NMS calls + syncs)
//
// Semantics kept: IoU with the +1 pixel convention evaluated as interS / (Sa + Sb - interS)
// (nms_kernel.cu:14-22), strict `>` threshold (:61), greedy in score order, kept indices
// returned ascending (:132-135).  Every sort is a total order: key descending, index
// ascending (stable radix sort), which is one of the outcomes of the reference's unstable
// sorts.  This file is compiled with -fmad=false (each product / sum rounded once).
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__device__ __forceinline__ float dev_iou(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(right - left + 1.0f, 0.0f), h = fmaxf(bottom - top + 1.0f, 0.0f);
  const float inter = w * h;
  const float sa = (a.z - a.x + 1.0f) * (a.w - a.y + 1.0f);
  const float sb = (b.z - b.x + 1.0f) * (b.w - b.y + 1.0f);
  return inter / (sa + sb - inter);
}

// ---- pairwise suppression bit mask -------------------------------------------------
// boxes [seg][n_cap] float4 in processing order; mask[seg][n_cap][nw] (nw = ceil(n_cap/64));
// bit j of row i is set when IoU(i, j) > thr (strict) / >= thr.  upper_only: only j > i.
// grid (nw, nw, segments), 64 threads.
__global__ void nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ counts, int n_cap, int nw,
                                float thr, int strict_gt, int upper_only, unsigned long long* __restrict__ mask) {
  const int seg = blockIdx.z;
  const int n = counts ? min(counts[seg], n_cap) : n_cap;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (upper_only && cb < rb) return;
  if (rb * 64 >= n || cb * 64 >= n) return;
  __shared__ float4 cbx[64];
  const float4* bx = boxes + (size_t)seg * n_cap;
  const int cj = cb * 64 + threadIdx.x;
  if (cj < n) cbx[threadIdx.x] = bx[cj];
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  const float4 a = bx[i];
  const int ncol = min(64, n - cb * 64);
  unsigned long long bits = 0;
  int start = 0;
  if (upper_only && rb == cb) start = threadIdx.x + 1;
  for (int j = start; j < ncol; ++j) {
    if (!upper_only && cb * 64 + j == i) continue;
    const float v = dev_iou(a, cbx[j]);
    if (strict_gt ? (v > thr) : (v >= thr)) bits |= 1ULL << j;
  }
  mask[((size_t)seg * n_cap + i) * nw + cb] = bits;
}

// ---- greedy scan over the upper-triangular mask, one CTA per segment -----------------
// keep_sorted[seg][*] receives positions (in processing order) of survivors, in order;
// stops after max_keep survivors.  128 threads.
__global__ void __launch_bounds__(128) nms_scan_kernel(const unsigned long long* __restrict__ mask,
                                                       const int* __restrict__ counts, int n_cap, int nw,
                                                       int max_keep, int* __restrict__ keep_sorted,
                                                       int* __restrict__ n_keep) {
  extern __shared__ unsigned long long removed[];  // [nw]
  __shared__ unsigned long long diag[64];
  __shared__ int klist[64];
  __shared__ int kn, ktotal;
  const int seg = blockIdx.x;
  const int n = counts ? min(counts[seg], n_cap) : n_cap;
  const unsigned long long* mk = mask + (size_t)seg * n_cap * nw;
  int* ks = keep_sorted + (size_t)seg * n_cap;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) removed[w] = 0;
  if (threadIdx.x == 0) ktotal = 0;
  __syncthreads();
  const int nchunks = (n + 63) / 64;
  for (int c = 0; c < nchunks; ++c) {
    const int base = c * 64;
    if (threadIdx.x < 64) {
      const int i = base + threadIdx.x;
      diag[threadIdx.x] = i < n ? mk[(size_t)i * nw + c] : 0ULL;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rm = removed[c];
      int k = 0, tot = ktotal;
      const int lim = min(64, n - base);
      for (int b = 0; b < lim && tot < max_keep; ++b) {
        if (!((rm >> b) & 1ULL)) {
          klist[k++] = base + b;
          ks[tot++] = base + b;
          rm |= diag[b];
        }
      }
      kn = k;
      ktotal = tot;
    }
    __syncthreads();
    if (ktotal >= max_keep) break;
    const int k = kn;
    for (int w = c + 1 + threadIdx.x; w < nw; w += blockDim.x) {
      unsigned long long acc = removed[w];
      for (int t = 0; t < k; ++t) acc |= mk[(size_t)klist[t] * nw + w];
      removed[w] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_keep[seg] = ktotal;
}

// ---- greedy NMS against the kept list, one CTA per segment (RPN path) --------------------
// Box j (score order) survives iff no already-kept box i has IoU(i, j) > thr - the same
// decision the mask + scan pair takes, but only the pairs (candidate, kept) are ever
// evaluated and the walk stops after max_keep survivors: ~n_visited * max_keep IoUs instead of
// n^2 / 2 (300-of-6000 proposals: ~30x less work, no mask in HBM).  1024 threads; per chunk of
// 64 candidates: (1) 16 slices of the kept list tested in parallel per candidate, (2) the 64x64
// in-chunk mask, (3) one thread resolves the chunk with one iteration per survivor.
__global__ void __launch_bounds__(1024) nms_greedy_kernel(const float4* __restrict__ boxes, int n_cap, float thr,
                                                          int strict_gt, int max_keep, int* __restrict__ keep_sorted,
                                                          int* __restrict__ n_keep) {
  extern __shared__ float4 kept[];   // [max_keep]
  __shared__ float4 cand[64];
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long supp;
  __shared__ int nk_s;
  const int seg = blockIdx.x;
  const float4* bx = boxes + (size_t)seg * n_cap;
  int* ks = keep_sorted + (size_t)seg * n_cap;
  const int tid = threadIdx.x;
  const int c = tid & 63, sl = tid >> 6;             // candidate, slice (16 slices)
  if (tid == 0) nk_s = 0;
  __syncthreads();
  for (int base = 0; base < n_cap; base += 64) {
    const int lim = min(64, n_cap - base);
    if (tid < 64) {
      if (tid < lim) cand[tid] = bx[base + tid];
      diag[tid] = 0ULL;
      if (tid == 0) supp = 0ULL;
    }
    __syncthreads();
    const int nk = nk_s;
    if (c < lim) {
      const float4 cb = cand[c];
      // (1) candidate c against the kept list, 16 slices of the list in parallel
      bool sup = false;
      for (int k = sl; k < nk; k += 16) {
        const float v = dev_iou(kept[k], cb);
        if (strict_gt ? (v > thr) : (v >= thr)) { sup = true; break; }
      }
      if (sup) atomicOr(&supp, 1ULL << c);
      // (2) in-chunk pairs (c, j), j > c
      unsigned long long bits = 0;
      for (int j = c + 1 + sl; j < lim; j += 16) {
        const float v = dev_iou(cb, cand[j]);
        if (strict_gt ? (v > thr) : (v >= thr)) bits |= 1ULL << j;
      }
      if (bits) atomicOr(&diag[c], bits);
    }
    __syncthreads();
    if (tid == 0) {
      // (3) serial resolve, one iteration per SURVIVOR of the chunk
      unsigned long long alive = ~supp;
      if (lim < 64) alive &= (1ULL << lim) - 1ULL;
      int k = nk;
      while (alive && k < max_keep) {
        const int b = __ffsll((long long)alive) - 1;
        kept[k] = cand[b];
        ks[k] = base + b;
        ++k;
        alive &= ~(diag[b] | (1ULL << b));
      }
      nk_s = k;
    }
    __syncthreads();
    if (nk_s >= max_keep) break;
  }
  if (tid == 0) n_keep[seg] = nk_s;
}

// ---- hvr_nms helpers ------------------------------------------------------------------
__global__ void nms_prepare_kernel(const float* __restrict__ dets, int n, float* __restrict__ keys,
                                   int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = dets[(size_t)i * 5 + 4];
    idx[i] = i;
  }
}
__global__ void nms_gather_boxes_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n,
                                        float4* __restrict__ boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float* d = dets + (size_t)order[i] * 5;
    boxes[i] = make_float4(d[0], d[1], d[2], d[3]);
  }
}
// kept positions (score order) -> kept original indices, ascending.  Single CTA.
__global__ void __launch_bounds__(1024) nms_finalize_kernel(const int* __restrict__ keep_sorted,
                                                            const int* __restrict__ n_keep_in,
                                                            const int* __restrict__ order, int n,
                                                            unsigned char* __restrict__ flags,
                                                            long long* __restrict__ keep, int* __restrict__ n_keep) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int k = *n_keep_in;
  for (int i = threadIdx.x; i < n; i += blockDim.x) flags[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) flags[order[keep_sorted[i]]] = 1;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int f = (i < n) ? flags[i] : 0;
    int v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const int pos = carry + (wid > 0 ? wsum[wid - 1] : 0) + v - f;
    if (f) keep[pos] = i;
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_keep = k;
}

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Carver {
  uint8_t* p;
  size_t off = 0;
  explicit Carver(void* base) : p(reinterpret_cast<uint8_t*>(base)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(p + off);
    off += align_up(count * sizeof(T));
    return r;
  }
};

size_t sort_temp_bytes(int n) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const float*)nullptr, (float*)nullptr, (const int*)nullptr,
                                            (int*)nullptr, n);
  return b;
}
size_t sort64_temp_bytes(int total) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, total);
  return b;
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" size_t hvr_nms_workspace_bytes(int n) {
  if (n < 1) n = 1;
  const size_t nw = (n + 63) / 64;
  size_t b = 0;
  b += 2 * align_up((size_t)n * 4) * 2;       // keys in/out, idx in/out
  b += align_up((size_t)n * 16);              // boxes
  b += align_up((size_t)n * nw * 8);          // mask
  b += align_up((size_t)n * 4);               // keep_sorted
  b += align_up(4);                           // n_keep tmp
  b += align_up((size_t)n);                   // flags
  b += align_up(sort_temp_bytes(n));
  return b + 256;
}

extern "C" int hvr_nms(const float* dets, int n, float iou_thr, int strict_gt, int64_t* keep, int* n_keep, void* ws,
                       size_t ws_bytes, void* stream) {
  if (n < 0 || !n_keep || (n > 0 && (!dets || !keep))) return HVR_ERR_ARG;
  cudaStream_t st = ST(stream);
  if (n == 0) {  // nms_cuda.cpp:8-11: empty in, empty out
    HVR_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int), st));
    return HVR_OK;
  }
  if (!ws || ws_bytes < hvr_nms_workspace_bytes(n)) return HVR_ERR_WORKSPACE;
  const int nw = (n + 63) / 64;
  Carver cv(ws);
  float* keys_in = cv.take<float>(n);
  float* keys_out = cv.take<float>(n);
  int* idx_in = cv.take<int>(n);
  int* order = cv.take<int>(n);
  float4* boxes = cv.take<float4>(n);
  unsigned long long* mask = cv.take<unsigned long long>((size_t)n * nw);
  int* keep_sorted = cv.take<int>(n);
  int* nk_tmp = cv.take<int>(1);
  unsigned char* flags = cv.take<unsigned char>(n);
  size_t tb = sort_temp_bytes(n);
  void* temp = cv.take<uint8_t>(tb);
  nms_prepare_kernel<<<hvr_cdiv(n, 256), 256, 0, st>>>(dets, n, keys_in, idx_in);
  HVR_LAUNCHED();
  HVR_CUDA(cub::DeviceRadixSort::SortPairsDescending(temp, tb, keys_in, keys_out, idx_in, order, n, 0, 32, st));
  g_hvr_launches.fetch_add(1);
  nms_gather_boxes_kernel<<<hvr_cdiv(n, 256), 256, 0, st>>>(dets, order, n, boxes);
  HVR_LAUNCHED();
  nms_mask_kernel<<<dim3(nw, nw, 1), 64, 0, st>>>(boxes, nullptr, n, nw, iou_thr, strict_gt, 1, mask);
  HVR_LAUNCHED();
  nms_scan_kernel<<<1, 128, nw * sizeof(unsigned long long), st>>>(mask, nullptr, n, nw, n, keep_sorted, nk_tmp);
  HVR_LAUNCHED();
  nms_finalize_kernel<<<1, 1024, 0, st>>>(keep_sorted, nk_tmp, order, n, flags, (long long*)keep, n_keep);
  HVR_LAUNCHED();
  return HVR_OK;
}

// =====================================================================================
// RPN proposals
// =====================================================================================
namespace {

// key = frame << 32 | ~orderable(logit): one ascending stable radix sort orders every frame's
// anchors by (logit descending, anchor index ascending) - frames stay contiguous.
__device__ __forceinline__ unsigned orderable(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__global__ void rpn_keys_kernel(const float* __restrict__ cls, long long ld_cls, int T, int cells, int A,
                                unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  const int n_anc = cells * A;
  const size_t total = (size_t)T * n_anc;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % n_anc);
    const int t = (int)(i / n_anc);
    const int a = r % A, cell = r / A;
    const float v = cls[((size_t)t * cells + cell) * ld_cls + a];
    keys[i] = ((unsigned long long)t << 32) | (unsigned long long)(~orderable(v));
    idx[i] = r;
  }
}

// transforms.py:34-111 for one box (means 0).  Explicitly rounded ops; exp via expf.
__device__ __forceinline__ float4 delta2bbox_one(float4 roi, float dx, float dy, float dw, float dh, float sx,
                                                 float sy, float sw, float sh, float max_ratio, float img_h,
                                                 float img_w) {
  dx = dx * sx + 0.0f; dy = dy * sy + 0.0f; dw = dw * sw + 0.0f; dh = dh * sh + 0.0f;
  dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
  dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
  const float px = (roi.x + roi.z) * 0.5f, py = (roi.y + roi.w) * 0.5f;
  const float pw = roi.z - roi.x + 1.0f, ph = roi.w - roi.y + 1.0f;
  const float gw = pw * expf(dw), gh = ph * expf(dh);
  const float gx = px + pw * dx, gy = py + ph * dy;
  float x1 = gx - gw * 0.5f + 0.5f, y1 = gy - gh * 0.5f + 0.5f;
  float x2 = gx + gw * 0.5f - 0.5f, y2 = gy + gh * 0.5f - 0.5f;
  if (img_w > 0.f) {
    x1 = fminf(fmaxf(x1, 0.f), img_w - 1.f); x2 = fminf(fmaxf(x2, 0.f), img_w - 1.f);
    y1 = fminf(fmaxf(y1, 0.f), img_h - 1.f); y2 = fminf(fmaxf(y2, 0.f), img_h - 1.f);
  }
  return make_float4(x1, y1, x2, y2);
}

__global__ void rpn_decode_kernel(const unsigned long long* __restrict__ keys_sorted,
                                  const int* __restrict__ idx_sorted,
                                  const float* __restrict__ reg, long long ld_reg, int T, int cells, int Wf, int A,
                                  const float* __restrict__ base_anchors, int stride, float img_h, float img_w,
                                  int n_anc, int npre, float max_ratio, float4* __restrict__ boxes,
                                  float* __restrict__ scores) {
  const int total = T * npre;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = i % npre, t = i / npre;
  const int ai = idx_sorted[(size_t)t * n_anc + r];
  const float logit = from_orderable(~(unsigned)(keys_sorted[(size_t)t * n_anc + r] & 0xFFFFFFFFull));
  const int a = ai % A, cell = ai / A;
  const int cx = cell % Wf, cy = cell / Wf;
  const float shx = (float)(cx * stride), shy = (float)(cy * stride);
  const float4 anc = make_float4(base_anchors[a * 4 + 0] + shx, base_anchors[a * 4 + 1] + shy,
                                 base_anchors[a * 4 + 2] + shx, base_anchors[a * 4 + 3] + shy);
  const float* d = reg + ((size_t)t * cells + cell) * ld_reg + a * 4;
  boxes[i] = delta2bbox_one(anc, d[0], d[1], d[2], d[3], 1.f, 1.f, 1.f, 1.f, max_ratio, img_h, img_w);
  scores[i] = 1.0f / (1.0f + expf(-logit));
}

__global__ void rpn_emit_kernel(const int* __restrict__ keep_sorted, const int* __restrict__ n_keep,
                                const float4* __restrict__ boxes, const float* __restrict__ scores,
                                const int* __restrict__ idx_sorted, int n_anc, int npre, int max_num,
                                float* __restrict__ proposals, int* __restrict__ counts, int* __restrict__ top_idx) {
  const int t = blockIdx.x;
  const int k = min(n_keep[t], max_num);
  for (int j = threadIdx.x; j < max_num; j += blockDim.x) {
    float* o = proposals + ((size_t)t * max_num + j) * 5;
    if (j < k) {
      const int pos = keep_sorted[(size_t)t * npre + j];
      const float4 b = boxes[(size_t)t * npre + pos];
      o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
      o[4] = scores[(size_t)t * npre + pos];
      if (top_idx) top_idx[(size_t)t * max_num + j] = idx_sorted[(size_t)t * n_anc + pos];
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
      if (top_idx) top_idx[(size_t)t * max_num + j] = -1;
    }
  }
  if (threadIdx.x == 0) counts[t] = k;
}

}  // namespace

extern "C" size_t hvr_rpn_workspace_bytes(int T, int n_anchors, int nms_pre) {
  const int npre = nms_pre > 0 && nms_pre < n_anchors ? nms_pre : n_anchors;
  const size_t nw = (npre + 63) / 64;
  const size_t tot = (size_t)T * n_anchors;
  size_t b = 0;
  (void)nw;
  b += 2 * align_up(tot * 8) + 2 * align_up(tot * 4);   // keys in/out (u64), idx in/out
  b += align_up((size_t)T * npre * 16);              // boxes
  b += align_up((size_t)T * npre * 4);               // scores
  b += align_up((size_t)T * npre * 4);               // keep_sorted
  b += align_up((size_t)T * 4);                      // n_keep
  b += align_up(sort64_temp_bytes((int)tot));
  return b + 256;
}

extern "C" int hvr_rpn_proposals(const float* cls, int64_t ld_cls, const float* reg, int64_t ld_reg, int T, int H,
                                 int W, int A, const float* base_anchors, int stride, float img_h, float img_w,
                                 int nms_pre, int nms_post, int max_num, float nms_thr, float* proposals, int* counts,
                                 int* top_idx, void* ws, size_t ws_bytes, void* stream) {
  if (!cls || !reg || !base_anchors || !proposals || !counts || T < 1 || H < 1 || W < 1 || A < 1 || max_num < 1)
    return HVR_ERR_ARG;
  const int cells = H * W, n_anc = cells * A;
  if (!ws || ws_bytes < hvr_rpn_workspace_bytes(T, n_anc, nms_pre)) return HVR_ERR_WORKSPACE;
  const int npre = nms_pre > 0 && nms_pre < n_anc ? nms_pre : n_anc;
  const int nw = (npre + 63) / 64;
  const size_t tot = (size_t)T * n_anc;
  cudaStream_t st = ST(stream);
  Carver cv(ws);
  unsigned long long* keys_in = cv.take<unsigned long long>(tot);
  unsigned long long* keys_out = cv.take<unsigned long long>(tot);
  int* idx_in = cv.take<int>(tot);
  int* idx_out = cv.take<int>(tot);
  float4* boxes = cv.take<float4>((size_t)T * npre);
  float* scores = cv.take<float>((size_t)T * npre);
  int* keep_sorted = cv.take<int>((size_t)T * npre);
  int* n_keep = cv.take<int>(T);
  size_t tb = sort64_temp_bytes((int)tot);
  void* temp = cv.take<uint8_t>(tb);
  (void)nw;

  size_t blocks = (tot + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  rpn_keys_kernel<<<(int)blocks, 256, 0, st>>>(cls, ld_cls, T, cells, A, keys_in, idx_in);
  HVR_LAUNCHED();
  int frame_bits = 1;
  while ((1 << frame_bits) < T) ++frame_bits;
  HVR_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys_in, keys_out, idx_in, idx_out, (int)tot, 0,
                                           32 + frame_bits, st));
  g_hvr_launches.fetch_add(1);
  const float max_ratio = fabsf(logf(16.0f / 1000.0f));
  rpn_decode_kernel<<<hvr_cdiv((int64_t)T * npre, 256), 256, 0, st>>>(keys_out, idx_out, reg, ld_reg, T, cells, W, A,
                                                                       base_anchors, stride, img_h, img_w, n_anc,
                                                                       npre, max_ratio, boxes, scores);
  HVR_LAUNCHED();
  int cap = nms_post > 0 ? nms_post : npre;
  if (cap > max_num) cap = max_num;
  if ((size_t)cap * sizeof(float4) > 160 * 1024) return HVR_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    HVR_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  nms_greedy_kernel<<<T, 1024, (size_t)cap * sizeof(float4), st>>>(boxes, npre, nms_thr, 1, cap, keep_sorted, n_keep);
  HVR_LAUNCHED();
  rpn_emit_kernel<<<T, 128, 0, st>>>(keep_sorted, n_keep, boxes, scores, idx_out, n_anc, npre, max_num, proposals,
                                     counts, top_idx);
  HVR_LAUNCHED();
  return HVR_OK;
}

// =====================================================================================
// Detection post-processing (softmax + decode + multiclass NMS + top-k)
// =====================================================================================
namespace {

__global__ void det_decode_kernel(const float* __restrict__ rois, const float* __restrict__ cls, long long ld_cls,
                                  const float* __restrict__ reg, long long ld_reg, int n, int n_cls, float s0,
                                  float s1, float s2, float s3, float img_h, float img_w, float scale_factor,
                                  int rescale, float max_ratio, float* __restrict__ scores,
                                  float4* __restrict__ boxes, const int* __restrict__ n_valid = nullptr,
                                  int rows_per_problem = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_valid != nullptr && (i % rows_per_problem) >= n_valid[i / rows_per_problem]) {
    // row beyond this problem's proposal count (a frame that yielded fewer than max_num proposals):
    // score 0 in every class -> below any score threshold, never a candidate
    for (int k = 0; k < n_cls; ++k) scores[(size_t)i * n_cls + k] = 0.f;
    boxes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float* c = cls + (size_t)i * ld_cls;
  float mx = -INFINITY;
  for (int k = 0; k < n_cls; ++k) mx = fmaxf(mx, c[k]);
  float sum = 0.f;
  for (int k = 0; k < n_cls; ++k) sum += expf(c[k] - mx);
  for (int k = 0; k < n_cls; ++k) scores[(size_t)i * n_cls + k] = expf(c[k] - mx) / sum;
  const float* r = rois + (size_t)i * 5;
  const float* d = reg + (size_t)i * ld_reg;
  float4 b = delta2bbox_one(make_float4(r[1], r[2], r[3], r[4]), d[0], d[1], d[2], d[3], s0, s1, s2, s3, max_ratio,
                            img_h, img_w);
  if (rescale) { b.x = b.x / scale_factor; b.y = b.y / scale_factor; b.z = b.z / scale_factor; b.w = b.w / scale_factor; }
  boxes[i] = b;
}

// One CTA per foreground class: order rois by (score desc, index asc), greedy NMS on the
// shared symmetric mask (cached in smem), flags[c][i] = 1 for survivors.  n <= 2048.
__global__ void __launch_bounds__(256) det_class_nms_kernel(const float* __restrict__ scores, int n, int n_cls, int nw,
                                                            const unsigned long long* __restrict__ mask,
                                                            float score_thr, int* __restrict__ flags) {
  extern __shared__ unsigned long long sm[];
  const int npow = 1 << (32 - __clz(max(n - 1, 1)));
  unsigned long long* smask = sm;                              // [n*nw]
  unsigned long long* skey = sm + (size_t)n * nw;              // [npow]
  unsigned long long* removed = skey + npow;                   // [nw]
  const int c = blockIdx.x + 1;
  // blockIdx.y = problem of a batched call: its own scores / mask / flags
  scores += (size_t)blockIdx.y * n * n_cls;
  mask += (size_t)blockIdx.y * n * nw;
  flags += (size_t)blockIdx.y * (n_cls - 1) * n;
  for (int i = threadIdx.x; i < n * nw; i += blockDim.x) smask[i] = mask[i];
  // composite key: high 32 = ordered score bits, low 32 = ~index  -> descending sort gives
  // score desc, index asc.  Entries with score <= thr get key 0 (sort last, ignored).
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    unsigned long long k = 0;
    if (i < n) {
      const float s = scores[(size_t)i * n_cls + c];
      if (s > score_thr) k = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
    }
    skey[i] = k;
  }
  for (int w = threadIdx.x; w < nw; w += blockDim.x) removed[w] = 0;
  __syncthreads();
  for (int k = 2; k <= npow; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skey[i], b = skey[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { skey[i] = b; skey[ixj] = a; }
        }
      }
      __syncthreads();
    }
  int* fl = flags + (size_t)(c - 1) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) fl[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int r = 0; r < n; ++r) {
      const unsigned long long k = skey[r];
      if (k == 0) break;
      const int i = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFu));
      if ((removed[i >> 6] >> (i & 63)) & 1ULL) continue;
      fl[i] = 1;
      for (int w = 0; w < nw; ++w) removed[w] |= smask[(size_t)i * nw + w];
    }
  }
}

// candidates in concatenation order p = (c-1)*n + i ; key = score (flag set) or -1.
__global__ void det_candidates_kernel(const int* __restrict__ flags, const float* __restrict__ scores, int n,
                                      int n_cls, float* __restrict__ keys, int* __restrict__ pos) {
  const int total = (n_cls - 1) * n;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int c = p / n + 1, i = p % n;
  keys[p] = flags[p] ? scores[(size_t)i * n_cls + c] : -1.0f;
  pos[p] = p;
}

__global__ void __launch_bounds__(256) det_emit_kernel(const int* __restrict__ flags, const int* __restrict__ fscan,
                                                       const float* __restrict__ keys_sorted,
                                                       const int* __restrict__ pos_sorted,
                                                       const float* __restrict__ scores,
                                                       const float4* __restrict__ boxes, int n, int n_cls,
                                                       int max_per_img, float* __restrict__ dets,
                                                       long long* __restrict__ labels, int* __restrict__ n_dets) {
  const int total = (n_cls - 1) * n;
  const int count = fscan[total - 1] + flags[total - 1];
  if (count <= max_per_img) {
    // concatenation order (class ascending, roi ascending) - bbox_nms.py:52-56
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
      if (!flags[p]) continue;
      const int o = fscan[p];
      const int c = p / n + 1, i = p % n;
      const float4 b = boxes[i];
      dets[o * 5 + 0] = b.x; dets[o * 5 + 1] = b.y; dets[o * 5 + 2] = b.z; dets[o * 5 + 3] = b.w;
      dets[o * 5 + 4] = scores[(size_t)i * n_cls + c];
      labels[o] = c - 1;
    }
  } else {
    // top max_per_img by (score desc, position asc) - bbox_nms.py:57-61
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < max_per_img; o += gridDim.x * blockDim.x) {
      const int p = pos_sorted[o];
      const int c = p / n + 1, i = p % n;
      const float4 b = boxes[i];
      dets[o * 5 + 0] = b.x; dets[o * 5 + 1] = b.y; dets[o * 5 + 2] = b.z; dets[o * 5 + 3] = b.w;
      dets[o * 5 + 4] = keys_sorted[o];
      labels[o] = c - 1;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_dets = min(count, max_per_img);
}

// Batched variants (G problems of n rois each, e.g. the key frames of G videos): one launch per
// stage for all problems.  Candidate keys are 64-bit (problem << 32 | ~score bits): ONE ascending
// stable radix sort orders every problem by (score desc, position asc), the same total order as the
// per-problem descending float sort.
__global__ void det_candidates_batched_kernel(const int* __restrict__ flags, const float* __restrict__ scores, int n,
                                              int n_cls, int G, unsigned long long* __restrict__ keys,
                                              int* __restrict__ pos) {
  const int total = (n_cls - 1) * n;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= G * total) return;
  const int g = p / total, q = p % total;
  const int c = q / n + 1, i = q % n;
  const unsigned lo = flags[p] ? ~__float_as_uint(scores[((size_t)g * n + i) * n_cls + c]) : 0xFFFFFFFFu;
  keys[p] = ((unsigned long long)g << 32) | lo;
  pos[p] = q;
}

__global__ void __launch_bounds__(256) det_emit_batched_kernel(const int* __restrict__ flags,
                                                               const int* __restrict__ fscan,
                                                               const int* __restrict__ pos_sorted,
                                                               const float* __restrict__ scores,
                                                               const float4* __restrict__ boxes, int n, int n_cls,
                                                               int max_per_img, float* __restrict__ dets,
                                                               long long* __restrict__ labels,
                                                               int* __restrict__ n_dets,
                                                               int* __restrict__ roi_idx = nullptr) {
  const int total = (n_cls - 1) * n;
  const int g = blockIdx.y;
  if (roi_idx) roi_idx += (size_t)g * max_per_img;
  flags += (size_t)g * total;
  fscan += (size_t)g * total;          // exclusive scan over ALL problems: subtract this problem's base
  pos_sorted += (size_t)g * total;
  scores += (size_t)g * n * n_cls;
  boxes += (size_t)g * n;
  dets += (size_t)g * max_per_img * 5;
  labels += (size_t)g * max_per_img;
  const int base = fscan[0];
  const int count = fscan[total - 1] + flags[total - 1] - base;
  if (count <= max_per_img) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
      if (!flags[p]) continue;
      const int o = fscan[p] - base;
      const int c = p / n + 1, i = p % n;
      const float4 b = boxes[i];
      dets[o * 5 + 0] = b.x; dets[o * 5 + 1] = b.y; dets[o * 5 + 2] = b.z; dets[o * 5 + 3] = b.w;
      dets[o * 5 + 4] = scores[(size_t)i * n_cls + c];
      labels[o] = c - 1;
      if (roi_idx) roi_idx[o] = i;
    }
  } else {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < max_per_img; o += gridDim.x * blockDim.x) {
      const int p = pos_sorted[o];
      const int c = p / n + 1, i = p % n;
      const float4 b = boxes[i];
      dets[o * 5 + 0] = b.x; dets[o * 5 + 1] = b.y; dets[o * 5 + 2] = b.z; dets[o * 5 + 3] = b.w;
      dets[o * 5 + 4] = scores[(size_t)i * n_cls + c];
      labels[o] = c - 1;
      if (roi_idx) roi_idx[o] = i;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) n_dets[g] = min(count, max_per_img);
}

size_t scan_temp_bytes(int n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, n);
  return b;
}

}  // namespace

extern "C" size_t hvr_det_workspace_bytes(int n, int n_cls) {
  if (n < 1) n = 1;
  const size_t nw = (n + 63) / 64;
  const size_t total = (size_t)(n_cls - 1) * n;
  size_t b = 0;
  b += align_up((size_t)n * n_cls * 4);     // scores
  b += align_up((size_t)n * 16);            // boxes
  b += align_up((size_t)n * nw * 8);        // mask
  b += 2 * align_up(total * 4);             // flags, scan
  b += 4 * align_up(total * 4);             // keys in/out, pos in/out
  b += align_up(sort_temp_bytes((int)total));
  b += align_up(scan_temp_bytes((int)total));
  return b + 256;
}

extern "C" int hvr_det_postprocess(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                                   int64_t ld_reg, int n, int n_cls, const float* stds4_host, float img_h,
                                   float img_w, float scale_factor, int rescale, float score_thr, float iou_thr,
                                   int max_per_img, float* dets, int64_t* labels, int* n_dets, void* ws,
                                   size_t ws_bytes, void* stream) {
  if (n < 0 || n_cls < 2 || !dets || !labels || !n_dets || max_per_img < 1 || !stds4_host) return HVR_ERR_ARG;
  cudaStream_t st = ST(stream);
  if (n == 0) {
    HVR_CUDA(cudaMemsetAsync(n_dets, 0, sizeof(int), st));
    return HVR_OK;
  }
  if (!rois || !cls || !reg) return HVR_ERR_ARG;
  if (n > 2048) return HVR_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < hvr_det_workspace_bytes(n, n_cls)) return HVR_ERR_WORKSPACE;
  const int nw = (n + 63) / 64;
  const int total = (n_cls - 1) * n;
  Carver cv(ws);
  float* scores = cv.take<float>((size_t)n * n_cls);
  float4* boxes = cv.take<float4>(n);
  unsigned long long* mask = cv.take<unsigned long long>((size_t)n * nw);
  int* flags = cv.take<int>(total);
  int* fscan = cv.take<int>(total);
  float* keys_in = cv.take<float>(total);
  float* keys_out = cv.take<float>(total);
  int* pos_in = cv.take<int>(total);
  int* pos_out = cv.take<int>(total);
  size_t tb = sort_temp_bytes(total), sb = scan_temp_bytes(total);
  void* temp = cv.take<uint8_t>(tb);
  void* stemp = cv.take<uint8_t>(sb);
  const float max_ratio = fabsf(logf(16.0f / 1000.0f));
  det_decode_kernel<<<hvr_cdiv(n, 128), 128, 0, st>>>(rois, cls, ld_cls, reg, ld_reg, n, n_cls, stds4_host[0],
                                                      stds4_host[1], stds4_host[2], stds4_host[3], img_h, img_w,
                                                      scale_factor, rescale, max_ratio, scores, boxes);
  HVR_LAUNCHED();
  nms_mask_kernel<<<dim3(nw, nw, 1), 64, 0, st>>>(boxes, nullptr, n, nw, iou_thr, 1, 0, mask);
  HVR_LAUNCHED();
  const int npow = 1 << (32 - __builtin_clz(n - 1 > 1 ? n - 1 : 1));
  const size_t smem = ((size_t)n * nw + npow + nw) * 8;
  static bool attr = false;
  if (!attr) {
    HVR_CUDA(cudaFuncSetAttribute(det_class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  if (smem > 220 * 1024) return HVR_ERR_UNSUPPORTED;
  det_class_nms_kernel<<<n_cls - 1, 256, smem, st>>>(scores, n, n_cls, nw, mask, score_thr, flags);
  HVR_LAUNCHED();
  det_candidates_kernel<<<hvr_cdiv(total, 256), 256, 0, st>>>(flags, scores, n, n_cls, keys_in, pos_in);
  HVR_LAUNCHED();
  HVR_CUDA(cub::DeviceScan::ExclusiveSum(stemp, sb, flags, fscan, total, st));
  g_hvr_launches.fetch_add(1);
  HVR_CUDA(cub::DeviceRadixSort::SortPairsDescending(temp, tb, keys_in, keys_out, pos_in, pos_out, total, 0, 32, st));
  g_hvr_launches.fetch_add(1);
  det_emit_kernel<<<hvr_cdiv(total, 256), 256, 0, st>>>(flags, fscan, keys_out, pos_out, scores, boxes, n, n_cls,
                                                        max_per_img, dets, (long long*)labels, n_dets);
  HVR_LAUNCHED();
  return HVR_OK;
}

extern "C" size_t hvr_det_batched_workspace_bytes(int G, int n, int n_cls) {
  if (n < 1) n = 1;
  if (G < 1) G = 1;
  const size_t nw = (n + 63) / 64;
  const size_t total = (size_t)G * (n_cls - 1) * n;
  size_t b = 0;
  b += align_up((size_t)G * n * n_cls * 4);   // scores
  b += align_up((size_t)G * n * 16);          // boxes
  b += align_up((size_t)G * n * nw * 8);      // masks
  b += 2 * align_up(total * 4);               // flags, scan
  b += 2 * align_up(total * 8);               // keys in/out
  b += 2 * align_up(total * 4);               // pos in/out
  b += align_up(sort64_temp_bytes((int)total));
  b += align_up(scan_temp_bytes((int)total));
  return b + 256;
}

// G problems (consecutive blocks of n rows of rois / cls / reg) through one launch per stage.
// Per problem bit-identical to hvr_det_postprocess.  dets [G, max_per_img, 5], labels
// [G, max_per_img], n_dets [G].
extern "C" int hvr_det_postprocess_batched(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                                           int64_t ld_reg, int G, int n, int n_cls, const float* stds4_host,
                                           float img_h, float img_w, float scale_factor, int rescale,
                                           float score_thr, float iou_thr, int max_per_img, float* dets,
                                           int64_t* labels, int* n_dets, void* ws, size_t ws_bytes, void* stream) {
  return hvr_det_postprocess_batched_ex(rois, cls, ld_cls, reg, ld_reg, G, n, n_cls, stds4_host, img_h, img_w,
                                        scale_factor, rescale, score_thr, iou_thr, max_per_img, nullptr, dets, labels,
                                        n_dets, nullptr, ws, ws_bytes, stream);
}

extern "C" int hvr_det_postprocess_batched_ex(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                                              int64_t ld_reg, int G, int n, int n_cls, const float* stds4_host,
                                              float img_h, float img_w, float scale_factor, int rescale,
                                              float score_thr, float iou_thr, int max_per_img, const int* n_valid,
                                              float* dets, int64_t* labels, int* n_dets, int* roi_idx, void* ws,
                                              size_t ws_bytes, void* stream) {
  if (G < 1 || n < 1 || n_cls < 2 || !dets || !labels || !n_dets || max_per_img < 1 || !stds4_host) return HVR_ERR_ARG;
  if (!rois || !cls || !reg) return HVR_ERR_ARG;
  if (n > 2048 || G > 4096) return HVR_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < hvr_det_batched_workspace_bytes(G, n, n_cls)) return HVR_ERR_WORKSPACE;
  cudaStream_t st = ST(stream);
  const int nw = (n + 63) / 64;
  const int per = (n_cls - 1) * n;
  const int total = G * per;
  Carver cv(ws);
  float* scores = cv.take<float>((size_t)G * n * n_cls);
  float4* boxes = cv.take<float4>((size_t)G * n);
  unsigned long long* mask = cv.take<unsigned long long>((size_t)G * n * nw);
  int* flags = cv.take<int>(total);
  int* fscan = cv.take<int>(total);
  unsigned long long* keys_in = cv.take<unsigned long long>(total);
  unsigned long long* keys_out = cv.take<unsigned long long>(total);
  int* pos_in = cv.take<int>(total);
  int* pos_out = cv.take<int>(total);
  size_t tb = sort64_temp_bytes(total), sb = scan_temp_bytes(total);
  void* temp = cv.take<uint8_t>(tb);
  void* stemp = cv.take<uint8_t>(sb);
  const float max_ratio = fabsf(logf(16.0f / 1000.0f));
  det_decode_kernel<<<hvr_cdiv((int64_t)G * n, 128), 128, 0, st>>>(rois, cls, ld_cls, reg, ld_reg, G * n, n_cls,
                                                                  stds4_host[0], stds4_host[1], stds4_host[2],
                                                                  stds4_host[3], img_h, img_w, scale_factor, rescale,
                                                                  max_ratio, scores, boxes, n_valid, n);
  HVR_LAUNCHED();
  nms_mask_kernel<<<dim3(nw, nw, G), 64, 0, st>>>(boxes, nullptr, n, nw, iou_thr, 1, 0, mask);
  HVR_LAUNCHED();
  const int npow = 1 << (32 - __builtin_clz(n - 1 > 1 ? n - 1 : 1));
  const size_t smem = ((size_t)n * nw + npow + nw) * 8;
  static bool attr = false;
  if (!attr) {
    HVR_CUDA(cudaFuncSetAttribute(det_class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  if (smem > 220 * 1024) return HVR_ERR_UNSUPPORTED;
  det_class_nms_kernel<<<dim3(n_cls - 1, G), 256, smem, st>>>(scores, n, n_cls, nw, mask, score_thr, flags);
  HVR_LAUNCHED();
  det_candidates_batched_kernel<<<hvr_cdiv(total, 256), 256, 0, st>>>(flags, scores, n, n_cls, G, keys_in, pos_in);
  HVR_LAUNCHED();
  HVR_CUDA(cub::DeviceScan::ExclusiveSum(stemp, sb, flags, fscan, total, st));
  g_hvr_launches.fetch_add(1);
  int gbits = 0;
  while ((1 << gbits) < G) ++gbits;
  HVR_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys_in, keys_out, pos_in, pos_out, total, 0, 32 + gbits, st));
  g_hvr_launches.fetch_add(1);
  det_emit_batched_kernel<<<dim3(hvr_cdiv(per, 256), G), 256, 0, st>>>(flags, fscan, pos_out, scores, boxes, n, n_cls,
                                                                      max_per_img, dets, (long long*)labels, n_dets,
                                                                      roi_idx);
  HVR_LAUNCHED();
  return HVR_OK;
}
