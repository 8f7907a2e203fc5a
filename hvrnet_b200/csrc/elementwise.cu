// Bandwidth-bound helpers: split-bf16 conversion, NCHW<->NHWC, stem im2col, max-pool,
// relation-head row softmax.  All HBM-bound: 128-bit accesses, coalesced along C.
#include <string.h>

#include "common.cuh"

thread_local int g_hvr_last_cuda_error = 0;
std::atomic<uint64_t> g_hvr_launches{0};

extern "C" const char* hvr_strerror(int code) {
  switch (code) {
    case HVR_OK: return "ok";
    case HVR_ERR_ARG: return "bad argument";
    case HVR_ERR_CUDA: return "CUDA call failed (see hvr_last_cuda_error)";
    case HVR_ERR_WORKSPACE: return "workspace too small";
    case HVR_ERR_UNSUPPORTED: return "unsupported on this device/driver";
    default: return "unknown error";
  }
}
extern "C" int hvr_last_cuda_error(void) { return g_hvr_last_cuda_error; }
extern "C" int hvr_abi_version(void) { return 3; }   // 3: round-2 entry points (fast RoIAlign, masks, window kernels, composites)
extern "C" uint64_t hvr_launch_count(void) { return g_hvr_launches.load(); }

namespace {

constexpr int EW_THREADS = 256;
inline int ew_grid(size_t n, int per_thread = 1) {
  size_t b = (n + (size_t)EW_THREADS * per_thread - 1) / ((size_t)EW_THREADS * per_thread);
  const size_t cap = 148 * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void split_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) split2(x[i], hi[i], lo[i]);
}
__global__ void merge_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, size_t n,
                             float* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = merge2(hi[i], lo[i]);
}
__global__ void split2d_kernel(const float* __restrict__ x, int rows, int cols, int ld_in,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_out) {
  const size_t total = (size_t)rows * ld_out;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int r = (int)(i / ld_out), c = (int)(i % ld_out);
    const float v = c < cols ? x[(size_t)r * ld_in + c] : 0.f;
    split2(v, hi[i], lo[i]);
  }
}

// [B,C,HW] -> [B,HW,C] through a 32x33 shared tile (coalesced on both sides).
template <bool SPLIT>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int HW, float* __restrict__ of,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* xb = x + (size_t)b * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    t[j][threadIdx.x] = (c < C && p < HW) ? xb[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      const size_t o = ((size_t)b * HW + p) * C + c;
      const float v = t[threadIdx.x][j];
      if (SPLIT) split2(v, hi[o], lo[o]);
      else of[o] = v;
    }
  }
}
// [B,HW,C] -> [B,C,HW]
template <bool SPLIT>
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ xf, const __nv_bfloat16* __restrict__ hi,
                                    const __nv_bfloat16* __restrict__ lo, int C, int HW, float* __restrict__ out) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const size_t o = ((size_t)b * HW + p) * C + c;
      v = SPLIT ? merge2(hi[o], lo[o]) : xf[o];
    }
    t[j][threadIdx.x] = v;
  }
  __syncthreads();
  float* ob = out + (size_t)b * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) ob[(size_t)c * HW + p] = t[threadIdx.x][j];
  }
}

// Stem im2col: out[(b,oy,ox), k] with k = (r*7+s)*3 + c for the 7x7/2 pad-3 conv, 192 columns
// (147 real + zero padding).  One CTA per 32 x 4 tile of output pixels: the 3 x 13 x 69 input patch
// is staged in shared memory with coalesced loads (zeros outside the image = conv padding), then
// each thread emits 8-column groups with 16-byte stores (consecutive threads -> consecutive 16 bytes).
constexpr int IM_TW = 32, IM_TH = 4;
constexpr int IM_PH = 2 * IM_TH + 5, IM_PW = 2 * IM_TW + 5, IM_PS = IM_PW + 3;   // patch rows / cols / row pitch
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ img, int B, int H, int W, int OH,
                                                          int OW, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo) {
  __shared__ float patch[3 * IM_PH * IM_PS + 1];
  __shared__ __align__(16) unsigned short ktab[192];     // column k = (r*7 + s)*3 + c -> offset of (c, r, s) inside the patch; padding -> the zero word
  if (threadIdx.x < 192) {
    const int k = threadIdx.x, c = k % 3, rs = k / 3;
    ktab[k] = k < 147 ? (unsigned short)((c * IM_PH + rs / 7) * IM_PS + rs % 7) : (unsigned short)0xffff;
  }
  if (threadIdx.x == 255) patch[3 * IM_PH * IM_PS] = 0.f;
  const int ox0 = blockIdx.x * IM_TW, oy0 = blockIdx.y * IM_TH, b = blockIdx.z;
  const int ix0 = ox0 * 2 - 3, iy0 = oy0 * 2 - 3;
  for (int i = threadIdx.x; i < 3 * IM_PH * IM_PW; i += blockDim.x) {
    const int px = i % IM_PW, py = (i / IM_PW) % IM_PH, c = i / (IM_PW * IM_PH);
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)b * 3 + c) * H + iy) * W + ix);
    patch[(c * IM_PH + py) * IM_PS + px] = v;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < IM_TW * IM_TH * 24; item += blockDim.x) {
    const int g = item % 24, pl = item / 24;
    const int px = pl % IM_TW, py = pl / IM_TW;
    const int ox = ox0 + px, oy = oy0 + py;
    if (ox >= OW || oy >= OH) continue;
    const int base = (2 * py) * IM_PS + 2 * px;
    uint32_t hp[4], lp[4];
    const uint4 tv = *reinterpret_cast<const uint4*>(ktab + g * 8);
    const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      unsigned short hs[2], ls[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = (int)((tw[j / 2] >> (16 * e)) & 0xffffu);
        const float v = patch[t == 0xffff ? 3 * IM_PH * IM_PS : t + base];
        __nv_bfloat16 h, l;
        split2(v, h, l);
        hs[e] = __bfloat16_as_ushort(h);
        ls[e] = __bfloat16_as_ushort(l);
      }
      hp[j / 2] = (uint32_t)hs[0] | ((uint32_t)hs[1] << 16);
      lp[j / 2] = (uint32_t)ls[0] | ((uint32_t)ls[1] << 16);
    }
    const size_t o = (((size_t)b * OH + oy) * OW + ox) * 192 + (size_t)g * 8;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// Transpose of a split matrix: in [rows, ld_in] (cols valid) -> out [cols, ld_out] with columns
// [rows, ld_out) zero-filled.  64 x 64 tiles through shared memory as packed (hi | lo << 16) words,
// 16-byte loads and stores on both sides.  (X^T operands of the relation head's P.V products.)
__global__ void __launch_bounds__(256) transpose_split_kernel(const __nv_bfloat16* __restrict__ hi,
                                                              const __nv_bfloat16* __restrict__ lo, int rows,
                                                              int cols, long long ld_in,
                                                              __nv_bfloat16* __restrict__ ohi,
                                                              __nv_bfloat16* __restrict__ olo, long long ld_out) {
  __shared__ uint32_t tile[64][65];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, cv = i & 7;
    uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
    if (r0 + r < rows && c0 + cv * 8 < cols) {
      h = *reinterpret_cast<const uint4*>(hi + (size_t)(r0 + r) * ld_in + c0 + cv * 8);
      l = *reinterpret_cast<const uint4*>(lo + (size_t)(r0 + r) * ld_in + c0 + cv * 8);
    }
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tile[cv * 8 + 2 * j][r] = (hw[j] & 0xFFFFu) | (lw[j] << 16);
      tile[cv * 8 + 2 * j + 1][r] = (hw[j] >> 16) | (lw[j] & 0xFFFF0000u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int c = i >> 3, rv = i & 7;
    if (c0 + c >= cols || r0 + rv * 8 >= ld_out) continue;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = tile[c][rv * 8 + j];
    uint4 h, l;
    h.x = (w[0] & 0xFFFFu) | (w[1] << 16); l.x = (w[0] >> 16) | (w[1] & 0xFFFF0000u);
    h.y = (w[2] & 0xFFFFu) | (w[3] << 16); l.y = (w[2] >> 16) | (w[3] & 0xFFFF0000u);
    h.z = (w[4] & 0xFFFFu) | (w[5] << 16); l.z = (w[4] >> 16) | (w[5] & 0xFFFF0000u);
    h.w = (w[6] & 0xFFFFu) | (w[7] << 16); l.w = (w[6] >> 16) | (w[7] & 0xFFFF0000u);
    *reinterpret_cast<uint4*>(ohi + (size_t)(c0 + c) * ld_out + r0 + rv * 8) = h;
    *reinterpret_cast<uint4*>(olo + (size_t)(c0 + c) * ld_out + r0 + rv * 8) = l;
  }
}

// 3x3 stride-2 pad-1 max-pool on split NHWC, 8 channels per thread.
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int B,
                               int H, int W, int C, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo,
                               int OH, int OW) {
  const int cg = C / 8;
  const size_t total = (size_t)B * OH * OW * cg;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int g = (int)(i % cg);
    const size_t pix = i / cg;
    const int ox = (int)(pix % OW);
    const int oy = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((size_t)OW * OH));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 - 1 + r;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * 2 - 1 + s;
        if (ix < 0 || ix >= W) continue;
        const size_t o = (((size_t)b * H + iy) * W + ix) * C + (size_t)g * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + o));
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + o));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          m[2 * q] = fmaxf(m[2 * q], __fadd_rn(bf16bits_to_f32(hw[q] & 0xFFFFu), bf16bits_to_f32(lw[q] & 0xFFFFu)));
          m[2 * q + 1] = fmaxf(m[2 * q + 1], __fadd_rn(bf16bits_to_f32(hw[q] >> 16), bf16bits_to_f32(lw[q] >> 16)));
        }
      }
    }
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat16 h0, l0, h1, l1;
      split2(m[2 * q], h0, l0);
      split2(m[2 * q + 1], h1, l1);
      hp[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lp[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t o = pix * C + (size_t)g * 8;
    *reinterpret_cast<uint4*>(ohi + o) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(olo + o) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// Row softmax, one CTA per row, row cached in registers (cols <= 256*32 = 8192) or re-read.
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
constexpr int SM_THREADS = 256;
constexpr int SM_VEC = 8;   // float4 per thread cached in registers: rows up to 8192 columns in one pass
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red, float* bcast) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < SM_THREADS / 32 ? red[lane] : (is_max ? -INFINITY : 0.f);
    t = is_max ? warp_max(t) : warp_sum(t);
    if (lane == 0) *bcast = t;
  }
  __syncthreads();
  return *bcast;
}
// One CTA per row.  Vector path (ld_s % 4 == 0, ld_p % 4 == 0, 16-byte aligned rows): LDG.128 of
// the logits, row cached in registers, 8-byte packed bf16 stores of hi and lo.
// Optional key mask (ragged proposal sets inside fixed row blocks): the columns are n_segs blocks of `slot`
// keys (the frames of a window, then the support key frames); only the first counts[problem][seg] keys of a
// block are proposals, the rest take no part in the softmax (probability exactly 0).  problem = row /
// rows_per_problem.  counts == nullptr: no mask (the arithmetic below is then untouched).
struct SegMask {
  const int* counts;
  int n_segs, slot, rows_per_problem;
  float inv_slot;
};
constexpr int SM_MAX_SEGS = 64;   // blocks whose limits are staged in shared memory (more: read from global)
// first column past the live keys of the block that holds column c.  lim: the row's per-block limits
// (seg * slot + count) staged in shared memory.  The block index comes from a float multiply instead of an integer
// division: (c + 0.5) / slot is never within 1e-3 of an integer, far outside fp32 rounding for c < 2^20.
__device__ __forceinline__ int seg_limit(const SegMask& m, const int* lim, int row, int c) {
  const int seg = (int)(((float)c + 0.5f) * m.inv_slot);
  if (lim) return lim[seg];
  return seg * m.slot + m.counts[(size_t)(row / m.rows_per_problem) * m.n_segs + seg];
}
template <bool MASKED>
__global__ void __launch_bounds__(SM_THREADS, 5) softmax_rows_kernel(const float* __restrict__ S, int cols, long long ld_s,
                                                                  __nv_bfloat16* __restrict__ phi,
                                                                  __nv_bfloat16* __restrict__ plo, long long ld_p,
                                                                  int vec_ok, const SegMask mask) {
  __shared__ float red[SM_THREADS / 32];
  __shared__ float bcast;
  __shared__ int s_lim[SM_MAX_SEGS];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int* lim = nullptr;
  const bool staged = MASKED && mask.n_segs <= SM_MAX_SEGS;
  auto stage_limits = [&]() {               // per-block limits of this row's window -> shared memory
    if (tid < mask.n_segs) s_lim[tid] = tid * mask.slot + mask.counts[(size_t)(row / mask.rows_per_problem) * mask.n_segs + tid];
    __syncthreads();
    lim = s_lim;
  };
  const float* s = S + (size_t)row * ld_s;
  __nv_bfloat16* ph = phi + (size_t)row * ld_p;
  __nv_bfloat16* pl = plo + (size_t)row * ld_p;
  if (vec_ok && cols <= SM_VEC * SM_THREADS * 4) {
    const int nv = (cols + 3) >> 2;          // float4 groups that contain a valid column
    float4 v[SM_VEC];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) {
      const int g = tid + j * SM_THREADS;
      if (g < nv) {
        v[j] = __ldg(reinterpret_cast<const float4*>(s) + g);
        const int c = g << 2;                // tail group: columns >= cols do not take part
        if (c + 1 >= cols) v[j].y = -INFINITY;
        if (c + 2 >= cols) v[j].z = -INFINITY;
        if (c + 3 >= cols) v[j].w = -INFINITY;
      } else {
        v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
    }
    if (MASKED) {
      // the counts are fetched while the logits above are in flight (the mask does not delay the row's loads)
      if (staged) stage_limits();
#pragma unroll
      for (int j = 0; j < SM_VEC; ++j) {
        const int c = (tid + j * SM_THREADS) << 2;   // slot % 4 == 0 on this path: the group lies inside one block
        if (c < cols) {
          const int end = seg_limit(mask, lim, row, c);
          if (c + 3 >= end) {                // only the groups that straddle / lie behind a block's count
            if (c >= end) v[j].x = -INFINITY;
            if (c + 1 >= end) v[j].y = -INFINITY;
            if (c + 2 >= end) v[j].z = -INFINITY;
            v[j].w = -INFINITY;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) mx = fmaxf(mx, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    mx = block_reduce(mx, true, red, &bcast);
    if (mx == -INFINITY) mx = 0.f;           // every key masked: probabilities 0 (exp(-inf) = 0, inv = 0 below)
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) {
      // same per-element arithmetic and per-thread accumulation order as the scalar path would give
      // for the elements this thread owns; the row sum is a tree over threads in both paths
      v[j].x = expf(v[j].x - mx); v[j].y = expf(v[j].y - mx); v[j].z = expf(v[j].z - mx); v[j].w = expf(v[j].w - mx);
      sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    sum = block_reduce(sum, false, red, &bcast);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    const int npv = (int)(ld_p >> 2);
#pragma unroll
    for (int j = 0; j < SM_VEC; ++j) {
      const int g = tid + j * SM_THREADS;
      if (g < npv) {
        const float4 p = g < nv ? make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat16 h[4], l[4];
        split2(p.x, h[0], l[0]); split2(p.y, h[1], l[1]); split2(p.z, h[2], l[2]); split2(p.w, h[3], l[3]);
        uint2 hv, lv;
        hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        *reinterpret_cast<uint2*>(ph + (g << 2)) = hv;
        *reinterpret_cast<uint2*>(pl + (g << 2)) = lv;
      }
    }
    for (int g = tid + SM_VEC * SM_THREADS; g < npv; g += SM_THREADS) {      // zero padding beyond the cached span
      *reinterpret_cast<uint2*>(ph + (g << 2)) = make_uint2(0, 0);
      *reinterpret_cast<uint2*>(pl + (g << 2)) = make_uint2(0, 0);
    }
    return;
  }
  // scalar path: any alignment, any length (re-reads the row)
  if (staged) stage_limits();
  auto live = [&](int c) { return !MASKED || c < seg_limit(mask, lim, row, c); };
  float mx = -INFINITY;
  for (int c = tid; c < cols; c += SM_THREADS)
    if (live(c)) mx = fmaxf(mx, __ldg(s + c));
  mx = block_reduce(mx, true, red, &bcast);
  if (mx == -INFINITY) mx = 0.f;
  float sum = 0.f;
  for (int c = tid; c < cols; c += SM_THREADS)
    if (live(c)) sum += expf(__ldg(s + c) - mx);
  sum = block_reduce(sum, false, red, &bcast);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  for (int c = tid; c < ld_p; c += SM_THREADS) {
    __nv_bfloat16 h, l;
    split2((c < cols && live(c)) ? expf(__ldg(s + c) - mx) * inv : 0.f, h, l);
    ph[c] = h;
    pl[c] = l;
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int hvr_split_f32(const float* x, size_t n, hvr_bf16* hi, hvr_bf16* lo, void* stream) {
  if (!x || !hi || !lo) return HVR_ERR_ARG;
  if (n == 0) return HVR_OK;
  split_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST(stream)>>>(x, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_merge_f32(const hvr_bf16* hi, const hvr_bf16* lo, size_t n, float* out, void* stream) {
  if (!out || !hi || !lo) return HVR_ERR_ARG;
  if (n == 0) return HVR_OK;
  merge_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST(stream)>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, n,
                                                              out);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_split_f32_2d(const float* x, int rows, int cols, int ld_in, hvr_bf16* hi, hvr_bf16* lo,
                                int ld_out, void* stream) {
  if (!x || !hi || !lo || cols > ld_out || cols > ld_in) return HVR_ERR_ARG;
  if (rows == 0) return HVR_OK;
  split2d_kernel<<<ew_grid((size_t)rows * ld_out, 4), EW_THREADS, 0, ST(stream)>>>(
      x, rows, cols, ld_in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_transpose_split(const hvr_bf16* hi, const hvr_bf16* lo, int rows, int cols, int64_t ld_in,
                                   hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, void* stream) {
  if (!hi || !lo || !out_hi || !out_lo || rows < 1 || cols < 1) return HVR_ERR_ARG;
  if (cols % 8 != 0 || ld_in % 8 != 0 || ld_out % 8 != 0 || ld_in < cols || ld_out < rows) return HVR_ERR_ARG;
  if (((uintptr_t)hi | (uintptr_t)lo | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15u) return HVR_ERR_ARG;
  // the grid covers the pad columns [rows, ld_out) too: they are written as zeros
  dim3 grid(hvr_cdiv(ld_out, 64), hvr_cdiv(cols, 64));
  if (grid.y > 65535) return HVR_ERR_UNSUPPORTED;
  transpose_split_kernel<<<grid, 256, 0, ST(stream)>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, rows, cols,
                                                        ld_in, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_out);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_nchw_to_nhwc_split(const float* x, int B, int C, int H, int W, hvr_bf16* hi, hvr_bf16* lo,
                                      void* stream) {
  if (!x || !hi || !lo || B < 1 || B > 65535) return HVR_ERR_ARG;
  dim3 grid(hvr_cdiv(H * W, 32), hvr_cdiv(C, 32), B), block(32, 8);
  nchw_to_nhwc_kernel<true><<<grid, block, 0, ST(stream)>>>(x, C, H * W, nullptr, (__nv_bfloat16*)hi,
                                                            (__nv_bfloat16*)lo);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_nchw_to_nhwc_f32(const float* x, int B, int C, int H, int W, float* out, void* stream) {
  if (!x || !out || B < 1 || B > 65535) return HVR_ERR_ARG;
  dim3 grid(hvr_cdiv(H * W, 32), hvr_cdiv(C, 32), B), block(32, 8);
  nchw_to_nhwc_kernel<false><<<grid, block, 0, ST(stream)>>>(x, C, H * W, out, nullptr, nullptr);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_nhwc_split_to_nchw(const hvr_bf16* hi, const hvr_bf16* lo, int B, int C, int H, int W,
                                      float* out, void* stream) {
  if (!out || !hi || !lo || B < 1 || B > 65535) return HVR_ERR_ARG;
  dim3 grid(hvr_cdiv(H * W, 32), hvr_cdiv(C, 32), B), block(32, 8);
  nhwc_to_nchw_kernel<true><<<grid, block, 0, ST(stream)>>>(nullptr, (const __nv_bfloat16*)hi,
                                                            (const __nv_bfloat16*)lo, C, H * W, out);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_nhwc_to_nchw_f32(const float* x, int B, int C, int H, int W, float* out, void* stream) {
  if (!out || !x || B < 1 || B > 65535) return HVR_ERR_ARG;
  dim3 grid(hvr_cdiv(H * W, 32), hvr_cdiv(C, 32), B), block(32, 8);
  nhwc_to_nchw_kernel<false><<<grid, block, 0, ST(stream)>>>(x, nullptr, nullptr, C, H * W, out);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_im2col_stem(const float* img, int B, int H, int W, hvr_bf16* hi, hvr_bf16* lo, int out_h,
                               int out_w, void* stream) {
  if (!img || !hi || !lo) return HVR_ERR_ARG;
  if (out_h != (H + 6 - 7) / 2 + 1 || out_w != (W + 6 - 7) / 2 + 1) return HVR_ERR_ARG;
  if (B < 1 || B > 65535) return HVR_ERR_ARG;
  im2col_stem_kernel<<<dim3(hvr_cdiv(out_w, IM_TW), hvr_cdiv(out_h, IM_TH), B), 256, 0, ST(stream)>>>(
      img, B, H, W, out_h, out_w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_maxpool3x3s2_split(const hvr_bf16* hi, const hvr_bf16* lo, int B, int H, int W, int C,
                                      hvr_bf16* ohi, hvr_bf16* olo, int out_h, int out_w, void* stream) {
  if (!hi || !lo || !ohi || !olo || C % 8 != 0) return HVR_ERR_ARG;
  if (out_h != (H + 2 - 3) / 2 + 1 || out_w != (W + 2 - 3) / 2 + 1) return HVR_ERR_ARG;
  maxpool_kernel<<<ew_grid((size_t)B * out_h * out_w * (C / 8)), EW_THREADS, 0, ST(stream)>>>(
      (const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, B, H, W, C, (__nv_bfloat16*)ohi, (__nv_bfloat16*)olo, out_h,
      out_w);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_softmax_rows_split(const float* S, int rows, int cols, int64_t ld_s, hvr_bf16* p_hi,
                                      hvr_bf16* p_lo, int64_t ld_p, void* stream) {
  if (!S || !p_hi || !p_lo || cols < 1 || cols > ld_s || cols > ld_p) return HVR_ERR_ARG;
  if (rows == 0) return HVR_OK;
  const int vec_ok = (ld_s % 4 == 0) && (ld_p % 4 == 0) && ((reinterpret_cast<uintptr_t>(S) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p_hi) & 7u) == 0) && ((reinterpret_cast<uintptr_t>(p_lo) & 7u) == 0);
  softmax_rows_kernel<false><<<rows, SM_THREADS, 0, ST(stream)>>>(S, cols, ld_s, (__nv_bfloat16*)p_hi,
                                                                  (__nv_bfloat16*)p_lo, ld_p, vec_ok,
                                                                  SegMask{nullptr, 0, 1, 1, 1.0f});
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_softmax_rows_split_masked(const float* S, int rows, int cols, int64_t ld_s, hvr_bf16* p_hi,
                                             hvr_bf16* p_lo, int64_t ld_p, const int* seg_counts, int n_segs,
                                             int slot, int rows_per_problem, void* stream) {
  if (!S || !p_hi || !p_lo || cols < 1 || cols > ld_s || cols > ld_p) return HVR_ERR_ARG;
  if (!seg_counts || n_segs < 1 || slot < 1 || rows_per_problem < 1 || (int64_t)n_segs * slot < cols) return HVR_ERR_ARG;
  if (rows == 0) return HVR_OK;
  const int vec_ok = (ld_s % 4 == 0) && (ld_p % 4 == 0) && (slot % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(S) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p_hi) & 7u) == 0) && ((reinterpret_cast<uintptr_t>(p_lo) & 7u) == 0);
  softmax_rows_kernel<true><<<rows, SM_THREADS, 0, ST(stream)>>>(
      S, cols, ld_s, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo, ld_p, vec_ok,
      SegMask{seg_counts, n_segs, slot, rows_per_problem, 1.0f / (float)slot});
  HVR_LAUNCHED();
  return HVR_OK;
}
