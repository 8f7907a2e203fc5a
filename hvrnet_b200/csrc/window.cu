// Window bookkeeping on the device: the glue between proposal generation, RoIAlign, the relation head and the
// detection post-processing of a batch of windows, written so that a whole key-frame step is a fixed sequence of
// launches on fixed-size buffers (CUDA-graph capturable) even when frames yield FEWER proposals than max_num.
//
// The reference builds these tensors on the host from the actual per-frame counts (hnmb_rcnn.py:580-599:
// bbox2roi per frame, cur_range from np.sum of the counts, torch.cat of the pooled rows).  Here every frame keeps
// a fixed block of `P` rows (P = max_num); the per-frame counts stay on the device and travel as masks
// (hvr_softmax_rows_split_masked, hvr_det_postprocess_batched_ex), so no launch geometry depends on them.
#include "common.cuh"

namespace {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// One thread per roi row of the batched layout [V, Npad] plus the small per-video outputs.
__global__ void window_rois_kernel(const float* __restrict__ props, const int* __restrict__ counts,
                                   const long long* __restrict__ perm, int V, int T, int P, int key_dim, int Npad,
                                   float* __restrict__ rois, float* __restrict__ rois_key, int* __restrict__ seg_counts,
                                   int n_segs, int* __restrict__ key_counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = V * Npad;
  if (i < total) {
    const int v = i / Npad, r = i - v * Npad;
    float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < T * P) {
      const int t = r / P, j = r - t * P;
      const int slot = perm ? (int)perm[v * T + t] : v * T + t;
      const float* p = props + ((size_t)slot * P + j) * 5;
      o[0] = (float)slot;                       // RoIAlign's batch index: the frame's slot in the C5 buffer
      o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = p[3];
      if (t == key_dim) {
        float* k = rois_key + ((size_t)v * P + j) * 5;   // bbox2roi([proposals of the key frame]): batch index 0
        k[0] = 0.f; k[1] = p[0]; k[2] = p[1]; k[3] = p[2]; k[4] = p[3];
      }
    }
    float* d = rois + (size_t)i * 5;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; d[3] = o[3]; d[4] = o[4];
  }
  if (i < V * T) {
    const int v = i / T, t = i - v * T;
    const int slot = perm ? (int)perm[i] : i;
    const int c = counts[slot];
    seg_counts[(size_t)v * n_segs + t] = c;
    if (t == key_dim) key_counts[v] = c;
  }
  if (i < V * (n_segs - T)) {                    // support blocks: empty until hvr_support_index fills them
    const int v = i / (n_segs - T), k = i - v * (n_segs - T);
    seg_counts[(size_t)v * n_segs + T + k] = 0;
  }
}

// dst[p * dst_rpp + dst_row0 + j, :cols] = src[row(p, j), :cols] for j < n_rows, 16-byte vectors.
// row(p, j) = idx ? idx[p * n_rows + j] : p * src_rpp + src_row0 + j; a negative index writes a zero row.
__global__ void gather_rows_kernel(const uint4* __restrict__ shi, const uint4* __restrict__ slo, long long ld_src,
                                   const int* __restrict__ idx, long long src_rpp, int src_row0,
                                   uint4* __restrict__ dhi, uint4* __restrict__ dlo, long long ld_dst, int n_problems,
                                   int n_rows, long long dst_rpp, int dst_row0, int vecs) {
  const long long total = (long long)n_problems * n_rows * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vecs);
    const long long r = i / vecs;
    const int p = (int)(r / n_rows), j = (int)(r - (long long)p * n_rows);
    const long long srow = idx ? (long long)idx[r] : p * src_rpp + src_row0 + j;
    uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
    if (srow >= 0) {
      h = __ldg(shi + srow * ld_src + c);
      l = __ldg(slo + srow * ld_src + c);
    }
    const long long drow = p * dst_rpp + dst_row0 + j;
    dhi[drow * ld_dst + c] = h;
    dlo[drow * ld_dst + c] = l;
  }
}

// Row indices of the support rows of every local key frame inside the gathered pool, and the key-mask entries of
// the support blocks.  Global key frame g = sel[v][s] lives on rank g / vpr as local key frame g % vpr:
//   idx[v][s * P + j]      = (g / vpr) * rank_stride_rows + (g % vpr) * P + j      (g < 0: -1 = zero rows)
//   seg_counts[v][T + s]   = pool_counts[(g / vpr) * counts_rank_stride + g % vpr]  (g < 0: 0)
__global__ void support_index_kernel(const long long* __restrict__ sel, const int* __restrict__ pool_counts,
                                     long long counts_rank_stride, int vpr, long long rank_stride_rows, int V, int S,
                                     int P, int T, int* __restrict__ idx, int* __restrict__ seg_counts, int n_segs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V * S * P) {
    const int v = i / (S * P), r = i - v * S * P;
    const int s = r / P, j = r - s * P;
    const long long g = sel[v * S + s];
    idx[i] = g < 0 ? -1 : (int)((g / vpr) * rank_stride_rows + (g % vpr) * P + j);
  }
  if (i < V * S) {
    const int v = i / S, s = i - v * S;
    const long long g = sel[i];
    seg_counts[(size_t)v * n_segs + T + s] = g < 0 ? 0 : pool_counts[(g / vpr) * counts_rank_stride + g % vpr];
  }
}

}  // namespace

extern "C" int hvr_window_rois(const float* props, const int* counts, const int64_t* perm, int V, int T, int P,
                               int key_dim, int Npad, float* rois, float* rois_key, int* seg_counts, int n_segs,
                               int* key_counts, void* stream) {
  if (!props || !counts || !rois || !rois_key || !seg_counts || !key_counts) return HVR_ERR_ARG;
  if (V < 1 || T < 1 || P < 1 || key_dim < 0 || key_dim >= T || Npad < T * P || n_segs < T) return HVR_ERR_ARG;
  const int total = V * Npad;
  window_rois_kernel<<<hvr_cdiv(total, 256), 256, 0, ST(stream)>>>(props, counts, (const long long*)perm, V, T, P,
                                                                   key_dim, Npad, rois, rois_key, seg_counts, n_segs,
                                                                   key_counts);
  HVR_LAUNCHED();
  return HVR_OK;
}

extern "C" int hvr_gather_rows_split(const hvr_bf16* src_hi, const hvr_bf16* src_lo, int64_t ld_src, const int* idx,
                                     int64_t src_rows_per_problem, int src_row0, hvr_bf16* dst_hi, hvr_bf16* dst_lo,
                                     int64_t ld_dst, int n_problems, int n_rows, int64_t dst_rows_per_problem,
                                     int dst_row0, int cols, void* stream) {
  if (!src_hi || !src_lo || !dst_hi || !dst_lo) return HVR_ERR_ARG;
  if (n_problems < 0 || n_rows < 0 || cols < 8 || cols % 8 || ld_src % 8 || ld_dst % 8 || cols > ld_src || cols > ld_dst)
    return HVR_ERR_ARG;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (!al16(src_hi) || !al16(src_lo) || !al16(dst_hi) || !al16(dst_lo)) return HVR_ERR_ARG;
  if (n_problems == 0 || n_rows == 0) return HVR_OK;
  const int vecs = cols / 8;
  const long long total = (long long)n_problems * n_rows * vecs;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  gather_rows_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(
      reinterpret_cast<const uint4*>(src_hi), reinterpret_cast<const uint4*>(src_lo), ld_src / 8, idx,
      src_rows_per_problem, src_row0, reinterpret_cast<uint4*>(dst_hi), reinterpret_cast<uint4*>(dst_lo), ld_dst / 8,
      n_problems, n_rows, dst_rows_per_problem, dst_row0, vecs);
  HVR_LAUNCHED();
  return HVR_OK;
}

extern "C" int hvr_support_index(const int64_t* sel, const int* pool_counts, int64_t counts_rank_stride, int vpr,
                                 int64_t rank_stride_rows, int V, int S, int P, int T, int* idx, int* seg_counts,
                                 int n_segs, void* stream) {
  if (!sel || !pool_counts || !idx || !seg_counts || V < 1 || S < 1 || P < 1 || T < 0 || n_segs < T + S || vpr < 1)
    return HVR_ERR_ARG;
  support_index_kernel<<<hvr_cdiv((long long)V * S * P, 256), 256, 0, ST(stream)>>>(
      (const long long*)sel, pool_counts, counts_rank_stride, vpr, rank_stride_rows, V, S, P, T, idx, seg_counts,
      n_segs);
  HVR_LAUNCHED();
  return HVR_OK;
}
