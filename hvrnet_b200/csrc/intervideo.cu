// Video-level similarity for the inter-video stage: the descriptor and the
// similarity of HNMBRCNN.get_triplet_patches (mmdet/models/detectors/hnmb_rcnn.py:76-101) used at
// inference to choose a key frame's support videos - oracle/ref_torch.py::video_descriptor,
// ::select_support_by_similarity.
//
//   descriptor  d_v[c] = max over the T frames of video v of mean over the h*w pixels of the shared
//               head's C5 map (:78-81).  HBM-bound: one read of the maps (T*h*w*C*4 B per video), done
//               as per-frame, per-pixel-chunk partial sums in a fixed order (deterministic, the same on
//               every rank), then a tiny finalisation.
//   selection   w_gj = softmax_j((1/sqrt(C)) d_g . d_j) over the other videos j (:85-88, :94-96); the
//               n_support largest, ties to the lower index.  One CTA per local key frame; latency only.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kPixChunks = 16;

// partial[frame][chunk][c] = sum of c5[frame][p][c] over the chunk's pixels, ascending p
__global__ void __launch_bounds__(256) desc_partial_kernel(const float* __restrict__ c5, int HW, int C,
                                                           float* __restrict__ partial) {
  const int frame = blockIdx.x, chunk = blockIdx.y;
  const int per = (HW + kPixChunks - 1) / kPixChunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  const float* base = c5 + (size_t)frame * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll 4
    for (int p = p0; p < p1; ++p) s = __fadd_rn(s, __ldg(base + (size_t)p * C + c));
    partial[((size_t)frame * kPixChunks + chunk) * C + c] = s;
  }
}

// desc[v][c] = max_t ( (sum_chunks partial[v*T+t][chunk][c]) / HW )
__global__ void __launch_bounds__(256) desc_final_kernel(const float* __restrict__ partial, int T, int HW, int C,
                                                         float* __restrict__ desc) {
  const int v = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m = -FLT_MAX;
    for (int t = 0; t < T; ++t) {
      const float* p = partial + ((size_t)(v * T + t) * kPixChunks) * C + c;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kPixChunks; ++k) s = __fadd_rn(s, p[(size_t)k * C]);
      m = fmaxf(m, __fdiv_rn(s, (float)HW));
    }
    desc[(size_t)v * C + c] = m;
  }
}

// block-wide argmax of (value desc, index asc) over sv/si[0..blockDim)
__device__ __forceinline__ void block_argmax(float* sv, int* si) {
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    __syncthreads();
    if ((int)threadIdx.x < off) {
      const float a = sv[threadIdx.x], b = sv[threadIdx.x + off];
      const int ia = si[threadIdx.x], ib = si[threadIdx.x + off];
      if (b > a || (b == a && ib < ia)) { sv[threadIdx.x] = b; si[threadIdx.x] = ib; }
    }
  }
  __syncthreads();
}

// one CTA per local key frame v (global video g0 + v); dynamic smem: w[G]
__global__ void __launch_bounds__(256) support_select_kernel(const float* __restrict__ desc, int G, int C, int g0,
                                                             int n_support, int64_t* __restrict__ idx,
                                                             float* __restrict__ weights) {
  extern __shared__ float w[];
  __shared__ float sv[256];
  __shared__ int si[256];
  const int v = blockIdx.x, g = g0 + v;
  const float scale = 1.0f / sqrtf((float)C);
  const float* dg = desc + (size_t)g * C;
  for (int j = threadIdx.x; j < G; j += blockDim.x) {
    const float* dj = desc + (size_t)j * C;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = __fmaf_rn(dg[k], dj[k], s);
    w[j] = j == g ? -FLT_MAX : scale * s;
  }
  __syncthreads();
  // softmax over the candidates j != g
  float m = -FLT_MAX;
  for (int j = threadIdx.x; j < G; j += blockDim.x) m = fmaxf(m, w[j]);
  sv[threadIdx.x] = m; si[threadIdx.x] = threadIdx.x;
  block_argmax(sv, si);
  m = sv[0];
  __syncthreads();
  float part = 0.f;
  for (int j = threadIdx.x; j < G; j += blockDim.x) {
    const float e = j == g ? 0.f : expf(w[j] - m);
    w[j] = e;
    part += e;
  }
  __syncthreads();
  sv[threadIdx.x] = part;
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    __syncthreads();
    if ((int)threadIdx.x < off) sv[threadIdx.x] += sv[threadIdx.x + off];
  }
  __syncthreads();
  const float total = sv[0];
  __syncthreads();
  for (int j = threadIdx.x; j < G; j += blockDim.x) {
    w[j] = j == g ? -1.f : w[j] / total;                       // -1: never selected (weights are >= 0)
    if (weights) weights[(size_t)v * G + j] = j == g ? 0.f : w[j];
  }
  __syncthreads();
  for (int r = 0; r < n_support; ++r) {
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < G; j += blockDim.x)
      if (w[j] > bv || (w[j] == bv && j < bi)) { bv = w[j]; bi = j; }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    block_argmax(sv, si);
    if (threadIdx.x == 0) {
      const bool ok = sv[0] >= 0.f;
      idx[(size_t)v * n_support + r] = ok ? si[0] : -1;        // fewer than n_support other videos
      if (ok) w[si[0]] = -1.f;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" size_t hvr_video_descriptor_workspace_bytes(int n_videos, int T, int C) {
  return (size_t)n_videos * T * kPixChunks * C * sizeof(float);
}

extern "C" int hvr_video_descriptor(const float* c5_nhwc, int n_videos, int T, int HW, int C, float* desc, void* ws,
                                    size_t ws_bytes, void* stream) {
  if (n_videos == 0) return HVR_OK;
  if (!c5_nhwc || !desc || n_videos < 0 || T < 1 || HW < 1 || C < 1) return HVR_ERR_ARG;
  if (!ws || ws_bytes < hvr_video_descriptor_workspace_bytes(n_videos, T, C)) return HVR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  desc_partial_kernel<<<dim3(n_videos * T, kPixChunks), 256, 0, st>>>(c5_nhwc, HW, C, (float*)ws);
  HVR_LAUNCHED();
  desc_final_kernel<<<n_videos, 256, 0, st>>>((const float*)ws, T, HW, C, desc);
  HVR_LAUNCHED();
  return HVR_OK;
}

extern "C" int hvr_support_select(const float* desc, int G, int C, int g0, int n_local, int n_support, int64_t* idx,
                                  float* weights, void* stream) {
  if (n_local == 0 || n_support == 0) return HVR_OK;
  if (!desc || !idx || G < 1 || C < 1 || g0 < 0 || n_local < 0 || g0 + n_local > G || n_support < 0) return HVR_ERR_ARG;
  if ((size_t)G * sizeof(float) > 40 * 1024) return HVR_ERR_UNSUPPORTED;     // 10 240 videos
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  support_select_kernel<<<n_local, 256, (size_t)G * sizeof(float), st>>>(desc, G, C, g0, n_support, idx, weights);
  HVR_LAUNCHED();
  return HVR_OK;
}
