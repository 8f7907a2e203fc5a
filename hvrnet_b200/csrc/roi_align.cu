// RoIAlign forward for sm_100a: a coalesced, 128-bit-vectorised HBM gather over an NHWC
// feature map with the bilinear taps of one RoI staged in shared memory.
//
// Replaces ROIAlignForward / bilinear_interpolate (mmdet/ops/roi_align/src/
// roi_align_kernel.cu:16-118) and its launcher (:120-141), which use one thread per output
// element and 16 scalar gathers per output.  Arithmetic contract (bit-exact with
// oracle/c/hvr_oracle.c, which restates the reference in strict IEEE fp32): every product
// and sum is rounded separately (this file is compiled with -fmad=false), sample order
// iy-outer / ix-inner, w1*lt + w2*rt + w3*lb + w4*rb left to right, divide by the sample
// count last.
//
// Layout of the work: one CTA per RoI.
//   phase 1  ph*pw*sn*sn threads compute (tap offsets, weights) of every sample once
//            -> shared memory (the reference recomputes them for each of the C channels)
//   phase 2  NHWC out: thread = (bin, 4-channel group): LDG.128 of contiguous channel rows,
//            1 x STG.128 (+ optional split-bf16 copy feeding fc_new_1).  sample_num == 2 (the
//            pipeline) runs roi_align_sn2_kernel: 8-16 loads per output vector (taps shared by
//            the two y-samples of a bin come from registers), no validity branches, no
//            per-iteration index arithmetic; other sample counts run roi_align_kernel<false>
//            with 4*sn*sn loads.
//            NCHW out (reference layout): thread = (bin, channel), lanes along C so loads
//            stay coalesced; the [C, ph*pw] tile is transposed through shared memory and
//            leaves as one contiguous, fully coalesced block.
#include "common.cuh"
#include "roi_align_bin.cuh"
#include "roi_align_sep.cuh"

#ifndef SN2_MIN_CTAS
#define SN2_MIN_CTAS 4
#endif

namespace {

struct RoiGeom {
  float sw, sh, bw, bh;
  int b;
};
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ r, float scale, int ph, int pw, int n_imgs) {
  RoiGeom g;
  int b = (int)r[0];
  g.b = b < 0 ? 0 : (b >= n_imgs ? n_imgs - 1 : b);
  g.sw = r[1] * scale;
  g.sh = r[2] * scale;
  const float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
  const float rw = fmaxf(ew - g.sw, 0.0f), rh = fmaxf(eh - g.sh, 0.0f);
  g.bh = rh / (float)ph;
  g.bw = rw / (float)pw;
  return g;
}

struct TapLoad {
  const char* base;   // image base + 4-channel group
  __device__ __forceinline__ float4 operator()(uint32_t off) const {
    return __ldg(reinterpret_cast<const float4*>(base + off));
  }
};

__device__ __forceinline__ void store_bin(const float4& acc, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                          __nv_bfloat16* __restrict__ out_lo, size_t o_f32, size_t o_split) {
  if (out) *reinterpret_cast<float4*>(out + o_f32) = acc;
  if (out_hi) {
    __nv_bfloat16 h[4], l[4];
    split2(acc.x, h[0], l[0]); split2(acc.y, h[1], l[1]);
    split2(acc.z, h[2], l[2]); split2(acc.w, h[3], l[3]);
    uint2 hv, lv;
    hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(out_hi + o_split) = hv;
    *reinterpret_cast<uint2*>(out_lo + o_split) = lv;
  }
}

template <bool NCHW_OUT>
__global__ void __launch_bounds__(256) roi_align_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                                        int n_imgs, int C, int H, int W, int ph, int pw, float scale,
                                                        int sn, float* __restrict__ out,
                                                        __nv_bfloat16* __restrict__ out_hi,
                                                        __nv_bfloat16* __restrict__ out_lo, long long ld_split) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int nbins = ph * pw;
  const int ns = sn * sn;
  Tap* taps = reinterpret_cast<Tap*>(smem);                          // [nbins*ns]
  float* tile = reinterpret_cast<float*>(smem + (size_t)nbins * ns * sizeof(Tap));  // NCHW_OUT: [cchunk][nbins]
  const int roi = blockIdx.x;
  const RoiGeom g = roi_geom(rois + (size_t)roi * 5, scale, ph, pw, n_imgs);

  for (int i = threadIdx.x; i < nbins * ns; i += blockDim.x) {
    const int s = i % ns, bin = i / ns;
    const int ix = s % sn, iy = s / sn;
    const int q = bin % pw, p = bin / pw;
    const float y = g.sh + (float)p * g.bh + ((float)iy + 0.5f) * g.bh / (float)sn;
    const float x = g.sw + (float)q * g.bw + ((float)ix + 0.5f) * g.bw / (float)sn;
    taps[i] = make_tap(y, x, H, W, C);
  }
  __syncthreads();
  const float* fm = feat + (size_t)g.b * H * W * C;
  const float cnt = (float)ns;

  if (!NCHW_OUT) {
    const int cg = C >> 2;  // float4 groups
    const int total = nbins * cg;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int c4 = i % cg, bin = i / cg;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const Tap* tp = taps + bin * ns;
      const TapLoad ld{reinterpret_cast<const char*>(fm + c4 * 4)};
      for (int s = 0; s < ns; ++s) {
        const Tap t = tp[s];
        if (t.o0 == kTapInvalid) continue;  // bilinear_interpolate returned 0: acc + 0 == acc
        const float4 v = bilerp4(t, ld(t.o0), ld(t.o1), ld(t.o2), ld(t.o3));
        acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
      }
      acc.x = acc.x / cnt; acc.y = acc.y / cnt; acc.z = acc.z / cnt; acc.w = acc.w / cnt;
      store_bin(acc, out, out_hi, out_lo, ((size_t)roi * nbins + bin) * C + (size_t)c4 * 4,
                (size_t)roi * ld_split + (size_t)bin * C + (size_t)c4 * 4);
    }
  } else {
    // channel chunks of blockDim.x: lanes along C (coalesced 128 B rows), tile transposed in smem
    const int cchunk = blockDim.x;
    for (int cbase = 0; cbase < C; cbase += cchunk) {
      const int c = cbase + threadIdx.x;
      const int cw = min(cchunk, C - cbase);
      if (c < C) {
        for (int bin = 0; bin < nbins; ++bin) {
          float acc = 0.f;
          const Tap* tp = taps + bin * ns;
          for (int s = 0; s < ns; ++s) {
            const Tap t = tp[s];
            if (t.o0 == kTapInvalid) continue;
            const float a = __ldg(fm + (t.o0 >> 2) + c), b = __ldg(fm + (t.o1 >> 2) + c);
            const float cc = __ldg(fm + (t.o2 >> 2) + c), d = __ldg(fm + (t.o3 >> 2) + c);
            acc = acc + (((t.w1 * a + t.w2 * b) + t.w3 * cc) + t.w4 * d);
          }
          tile[threadIdx.x * nbins + bin] = acc / cnt;   // stride nbins (odd for 7x7): conflict-free
        }
      }
      __syncthreads();
      float* o = out + ((size_t)roi * C + cbase) * nbins;   // [cw][nbins] contiguous in NCHW
      for (int i = threadIdx.x; i < cw * nbins; i += blockDim.x) o[i] = tile[i];
      __syncthreads();
    }
  }
}

// sample_num == 2, NHWC out (the pipeline's variant): roi_align_bin.cuh.  One CTA per RoI, thread =
// (bin, 4-channel group) with the channel group fixed per thread (blockDim % (C/4) == 0), so the only
// per-bin address arithmetic is base + byte offset.  Samples outside the map are rare; a RoI that has one
// takes the per-sample path below (CTA-uniform choice), all others run without validity branches.
__global__ void __launch_bounds__(256, SN2_MIN_CTAS) roi_align_sn2_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  extern __shared__ __align__(16) uint8_t smem[];
  Tap* taps = reinterpret_cast<Tap*>(smem);   // [nbins][iy][ix]
  const int nbins = ph * pw;
  const int roi = blockIdx.x;
  const RoiGeom g = roi_geom(rois + (size_t)roi * 5, scale, ph, pw, n_imgs);
  int invalid = 0;
  for (int i = threadIdx.x; i < nbins * 4; i += blockDim.x) {
    const int ix = i & 1, iy = (i >> 1) & 1, bin = i >> 2;
    const int q = bin % pw, p = bin / pw;
    const float y = g.sh + (float)p * g.bh + ((float)iy + 0.5f) * g.bh / 2.0f;
    const float x = g.sw + (float)q * g.bw + ((float)ix + 0.5f) * g.bw / 2.0f;
    const Tap t = make_tap(y, x, H, W, C);
    invalid |= (t.o0 == kTapInvalid);
    taps[i] = t;
  }
  const int any_invalid = __syncthreads_or(invalid);
  const int cg = C >> 2;
  const int c4 = threadIdx.x % cg;
  const int bin_step = blockDim.x / cg;
  const TapLoad ld{reinterpret_cast<const char*>(feat + (size_t)g.b * H * W * C + c4 * 4)};
  size_t o_f32 = ((size_t)roi * nbins + threadIdx.x / cg) * C + c4 * 4;
  size_t o_split = (size_t)roi * ld_split + (size_t)(threadIdx.x / cg) * C + c4 * 4;
  for (int bin = threadIdx.x / cg; bin < nbins; bin += bin_step) {
    const Tap* tp = taps + bin * 4;
    float4 acc;
    if (!any_invalid) {
      acc = roi_bin_sn2(tp, ld, nullptr);
    } else {
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < 4; ++s) {
        const Tap t = tp[s];
        if (t.o0 == kTapInvalid) continue;  // bilinear_interpolate returned 0: acc + 0 == acc
        const float4 v = bilerp4(t, ld(t.o0), ld(t.o1), ld(t.o2), ld(t.o3));
        acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
      }
      acc.x = acc.x * 0.25f; acc.y = acc.y * 0.25f; acc.z = acc.z * 0.25f; acc.w = acc.w * 0.25f;
    }
    store_bin(acc, out, out_hi, out_lo, o_f32, o_split);
    o_f32 += (size_t)bin_step * C;
    o_split += (size_t)bin_step * C;
  }
}

// Fast variant of the pipeline's path (sample_num == 2, NHWC rows out): the separable evaluation of
// roi_align_sep.cuh.  (This first walk - two-row cache with tags - serves C != 256 and the bit-identity tests; the
// pipeline's C == 256 launch runs the row-program walk further down, roi_align_sepp_kernel.)  One CTA per RoI, thread = (output column q, 4-channel group); the 2*ph y samples of the
// RoI are computed once into shared memory, the <= 4 merged column taps of q live in registers; every thread
// walks the RoI's rows top to bottom and writes its ph outputs.  All branches depend on the RoI only (rows)
// or on q only (warp-uniform when C % 128 == 0).  Explicit fmaf: independent of this file's -fmad=false.
// The walk of roi_align_sep.cuh::roi_column_sep_sn2 specialised for the device: the number of merged column taps NC
// (1..4, uniform per warp) is compile-time, so a row interpolation is NC loads + 4*NC multiply-adds of straight-line
// code on pre-added 64-bit column pointers; split stores are packed.  ncu of the first version (profiles/r02_roi_align_summary.txt): the
// kernel is instruction-issue bound - 1867 warp instructions per warp and RoI of which 289 are the multiply-adds -
// not memory bound, so the instruction count is what this version attacks.  Same operations in the same order as
// the generic core (the slab kernel, which still runs the core, is its bit-for-bit twin in the tests).
// VEC = float4 vectors (4 channels each) a thread owns: 1, or 2 (8 adjacent channels - the address arithmetic, the row
// walk's control flow and the table reads are then paid once per 8 channels instead of once per 4).
template <int VEC>
struct Vec { float4 v[VEC]; };

// ES = bytes between the VEC vectors of a thread: 16 (adjacent channels) or C/VEC*4 (lane-interleaved: vector e of lane j
// covers channels [(e*C/(4 VEC) + j) * 4, +4), so that every LDG.128 / STG.64 of a warp touches one contiguous run - with
// adjacent channels each of the two loads of a tap uses half of every 32-byte sector it requests).
template <int NC, int VEC, int ES>
__device__ __forceinline__ Vec<VEC> row_interp_nc(const char* const (&pc)[4], const float (&w)[4], uint32_t row_off) {
  float4 v[NC][VEC];
#pragma unroll
  for (int k = 0; k < NC; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[k][e] = __ldg(reinterpret_cast<const float4*>(pc[k] + row_off + e * ES));
  Vec<VEC> t;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    if (NC == 4) {
      t.v[e].x = w[3] * v[3][e].x; t.v[e].y = w[3] * v[3][e].y; t.v[e].z = w[3] * v[3][e].z; t.v[e].w = w[3] * v[3][e].w;
    } else {
      t.v[e] = make_float4(fmaf(w[NC - 1], v[NC - 1][e].x, 0.f), fmaf(w[NC - 1], v[NC - 1][e].y, 0.f),
                           fmaf(w[NC - 1], v[NC - 1][e].z, 0.f), fmaf(w[NC - 1], v[NC - 1][e].w, 0.f));
    }
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) {
      t.v[e].x = fmaf(w[k], v[k][e].x, t.v[e].x); t.v[e].y = fmaf(w[k], v[k][e].y, t.v[e].y);
      t.v[e].z = fmaf(w[k], v[k][e].z, t.v[e].z); t.v[e].w = fmaf(w[k], v[k][e].w, t.v[e].w);
    }
  }
  return t;
}

template <int VEC>
__device__ __forceinline__ void vec_fma(Vec<VEC>& a, float w, const Vec<VEC>& t) {
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    a.v[e].x = fmaf(w, t.v[e].x, a.v[e].x); a.v[e].y = fmaf(w, t.v[e].y, a.v[e].y);
    a.v[e].z = fmaf(w, t.v[e].z, a.v[e].z); a.v[e].w = fmaf(w, t.v[e].w, a.v[e].w);
  }
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> vec_zero() {
  Vec<VEC> z;
#pragma unroll
  for (int e = 0; e < VEC; ++e) z.v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  return z;
}

// split + store of one output vector: packed conversions (the same round-to-nearest-even as split2)
__device__ __forceinline__ void store_bin_packed(const float4& a, float* __restrict__ o_f32,
                                                 __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo) {
  if (o_f32) *reinterpret_cast<float4*>(o_f32) = a;
  if (o_hi) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(a.x, a.y), h23 = __floats2bfloat162_rn(a.z, a.w);
    const uint32_t b01 = *reinterpret_cast<const uint32_t*>(&h01), b23 = *reinterpret_cast<const uint32_t*>(&h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(__fsub_rn(a.x, __uint_as_float(b01 << 16)),
                                                     __fsub_rn(a.y, __uint_as_float(b01 & 0xFFFF0000u)));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(__fsub_rn(a.z, __uint_as_float(b23 << 16)),
                                                     __fsub_rn(a.w, __uint_as_float(b23 & 0xFFFF0000u)));
    *reinterpret_cast<uint2*>(o_hi) = make_uint2(b01, b23);
    *reinterpret_cast<uint2*>(o_lo) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

template <int NC, int VEC, int ES>
__device__ __forceinline__ void roi_column_walk(const RowTap* __restrict__ rows, int ph, const char* const (&pc)[4],
                                                const float (&w)[4], float* o_f32, __nv_bfloat16* o_hi,
                                                __nv_bfloat16* o_lo, size_t step) {
  uint32_t ca = kRowInvalid, cb = kRowInvalid;
  Vec<VEC> ta = vec_zero<VEC>(), tb = ta, acc = ta;
  // NOT unrolled: four NC variants of a fully unrolled 14-sample walk are ~100 KB of code
#pragma unroll 1
  for (int s = 0; s < 2 * ph; ++s) {
    const RowTap r = rows[s];
    if (r.o0 != kRowInvalid) {
      if (r.o0 == cb) { ta = tb; ca = cb; }
      else if (r.o0 != ca) { ta = row_interp_nc<NC, VEC, ES>(pc, w, r.o0); ca = r.o0; }
      vec_fma<VEC>(acc, r.w0, ta);
      if (r.w1 != 0.f) {
        if (r.o1 != ca) {
          if (r.o1 != cb) { tb = row_interp_nc<NC, VEC, ES>(pc, w, r.o1); cb = r.o1; }
          vec_fma<VEC>(acc, r.w1, tb);
        } else {
          vec_fma<VEC>(acc, r.w1, ta);
        }
      }
    }
    if (s & 1) {                                     // second sample of the bin: emit, next bin
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        store_bin_packed(acc.v[e], o_f32 ? o_f32 + (ES / 4) * e : nullptr, o_hi ? o_hi + (ES / 4) * e : nullptr,
                         o_hi ? o_lo + (ES / 4) * e : nullptr);
      if (o_f32) o_f32 += step;
      if (o_hi) { o_hi += step; o_lo += step; }
      acc = vec_zero<VEC>();
    }
  }
}

// ---- the walk as a per-RoI row program (roi_align_sepp_kernel) ----------------------------------------------------
// ncu of the walk above (profiles/r02q_*): 2245 warp instructions per warp and RoI of which 516 are multiply-adds; a
// quarter are register moves (the ta = tb hand-over of the two-row cache and its compares, kept on the local stack).
// Here the two cached rows live in fixed registers by ROW PARITY (a sample's two rows y, y + 1 always have different
// parity, and the rows a column walk needs never decrease, so a slot is never reloaded with an older row), and which
// sample loads which slot is decided ONCE per RoI by warp 0 (__match_any_sync over the samples' row indices: the first
// sample that uses a row loads it) into a 16-byte step per sample.  The same operations on the same operands in the same
// order as roi_column_walk / roi_align_sep.cuh::roi_column_sep_sn2, hence the same bits.
struct RowStep {
  uint32_t oe, oo;   // byte offsets of the even / odd map row of this sample; low bits of oe: 1 load even, 2 load odd,
  float we, wo;      //   4 the sample's upper row is the odd one, 8 sample inside the map.  we / wo: their weights
};

__device__ __forceinline__ void build_row_program(RowStep* __restrict__ prog, float start, float bin, int ph, int H,
                                                  uint32_t row_pitch) {
  const int s = threadIdx.x;                       // warp 0, all 32 lanes (2 * ph <= 32)
  RowStep st{0u, 0u, 0.f, 0.f};
  uint32_t key_e = 0x80000000u | (uint32_t)s, key_o = key_e, flags = 0;
  bool use_e = false, use_o = false;
  if (s < 2 * ph) {
    const RowTap r = row_tap_sn2(start, bin, s, H, 1u);          // o0 / o1 = row indices
    if (r.o0 != kRowInvalid) {
      const bool odd = r.o0 & 1u, second = r.w1 != 0.f;          // second => o1 == o0 + 1 (axis_tap)
      flags = 8u | (odd ? 4u : 0u);
      if (!odd) { st.oe = r.o0; st.we = r.w0; use_e = true; if (second) { st.oo = r.o1; st.wo = r.w1; use_o = true; } }
      else      { st.oo = r.o0; st.wo = r.w0; use_o = true; if (second) { st.oe = r.o1; st.we = r.w1; use_e = true; } }
      if (use_e) key_e = st.oe;
      if (use_o) key_o = st.oo;
    }
  }
  const uint32_t me = __match_any_sync(0xffffffffu, key_e), mo = __match_any_sync(0xffffffffu, key_o);
  if (use_e && __ffs(me) - 1 == s) flags |= 1u;
  if (use_o && __ffs(mo) - 1 == s) flags |= 2u;
  st.oe = st.oe * row_pitch | flags;                             // row_pitch is a multiple of 16 (C % 4 == 0)
  st.oo = st.oo * row_pitch;
  prog[s] = st;                                                  // s >= 2 * ph: empty steps (the walk reads one step ahead)
  if (s == 0) prog[32] = RowStep{0u, 0u, 0.f, 0.f};
}

// Row interpolation with the column taps addressed from ONE pointer (PAIR = false: the NC merged taps are consecutive map
// columns - always the case for RoIs up to 4 columns per bin) or TWO (PAIR = true, NC = 4: columns a, a+1, b, b+1): the
// column pitch CP = C * 4 bytes is compile-time, so the other taps are immediate offsets of the load (2 or 4 address
// instructions per row instead of 8, and 2 or 4 pointer registers instead of 8).  Same arithmetic as row_interp_nc.
template <int NC, int VEC, int ES, int CP, bool PAIR>
__device__ __forceinline__ Vec<VEC> row_interp_cp(const char* pa, const char* pb, const float (&w)[4], uint32_t row_off) {
  float4 v[NC][VEC];
  const char* ra = pa + row_off;
  const char* rb = PAIR ? pb + row_off : nullptr;
#pragma unroll
  for (int k = 0; k < NC; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      v[k][e] = __ldg(reinterpret_cast<const float4*>((PAIR && k >= 2 ? rb + (k - 2) * CP : ra + k * CP) + e * ES));
  Vec<VEC> t;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    if (NC == 4) {
      t.v[e].x = w[3] * v[3][e].x; t.v[e].y = w[3] * v[3][e].y; t.v[e].z = w[3] * v[3][e].z; t.v[e].w = w[3] * v[3][e].w;
    } else {
      t.v[e] = make_float4(fmaf(w[NC - 1], v[NC - 1][e].x, 0.f), fmaf(w[NC - 1], v[NC - 1][e].y, 0.f),
                           fmaf(w[NC - 1], v[NC - 1][e].z, 0.f), fmaf(w[NC - 1], v[NC - 1][e].w, 0.f));
    }
#pragma unroll
    for (int k = NC - 2; k >= 0; --k) {
      t.v[e].x = fmaf(w[k], v[k][e].x, t.v[e].x); t.v[e].y = fmaf(w[k], v[k][e].y, t.v[e].y);
      t.v[e].z = fmaf(w[k], v[k][e].z, t.v[e].z); t.v[e].w = fmaf(w[k], v[k][e].w, t.v[e].w);
    }
  }
  return t;
}

// OUT: 1 = fp32 rows, 2 = split rows, 3 = both (compile-time: no pointer tests in the per-bin epilogue).
// PF: while a step's rows are in flight, the rows the NEXT step will load are prefetched into L1 - one prefetch per
// column tap and row, lane j touching 32-byte sector j of the column's 1 KB (CP == 1024): the walk is bound by exposed
// L2 latency (ncu: 47 % of the cycles without an eligible warp, 46 % of the stall cycles on the loads' scoreboard).
template <int NC, int VEC, int ES, int CP, bool PAIR, bool PF>
__device__ __forceinline__ void prefetch_row(const char* pa, const char* pb, uint32_t row_off) {
  if (!PF) return;
  const uint32_t lane32 = (threadIdx.x & 31u) * 16u;              // pa already carries lane * 16
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const char* ptr = (PAIR && k >= 2 ? pb + (k - 2) * CP : pa + k * CP) + row_off + lane32;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
  }
}

// Outputs are addressed as kernel-uniform base + ONE 32-bit element offset shared by the fp32 / hi / lo rows (the launcher
// checks that the outputs are below 4 GB), which keeps the walk's live registers under the 80 that three CTAs per SM allow.
template <int NC, int VEC, int ES, int CP, bool PAIR, int OUT, bool PF>
__device__ __forceinline__ void roi_column_walk_prog(const RowStep* __restrict__ prog, int ph, const char* pa, const char* pb,
                                                     const float (&w)[4], float* __restrict__ out,
                                                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                     uint32_t o_f32, uint32_t o_split, uint32_t step) {
  Vec<VEC> te = vec_zero<VEC>(), to = te, acc = te;
#pragma unroll 1
  for (int p = 0; p < ph; ++p) {
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const uint4 st = *reinterpret_cast<const uint4*>(prog + 2 * p + iy);
      if (st.x & 1u) te = row_interp_cp<NC, VEC, ES, CP, PAIR>(pa, pb, w, st.x & ~15u);
      if (st.x & 2u) to = row_interp_cp<NC, VEC, ES, CP, PAIR>(pa, pb, w, st.y);
      if (PF) {                                                                     // prog has 2 * ph + 1 entries
        const uint2 nx = *reinterpret_cast<const uint2*>(prog + 2 * p + iy + 1);
        if (nx.x & 1u) prefetch_row<NC, VEC, ES, CP, PAIR, PF>(pa, pb, nx.x & ~15u);
        if (nx.x & 2u) prefetch_row<NC, VEC, ES, CP, PAIR, PF>(pa, pb, nx.y);
      }
      const float we = __uint_as_float(st.z), wo = __uint_as_float(st.w);
      if (st.x & 8u) {
        if (st.x & 4u) {
          vec_fma<VEC>(acc, wo, to);
          if (we != 0.f) vec_fma<VEC>(acc, we, te);
        } else {
          vec_fma<VEC>(acc, we, te);
          if (wo != 0.f) vec_fma<VEC>(acc, wo, to);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      store_bin_packed(acc.v[e], (OUT & 1) ? out + o_f32 + (ES / 4) * e : nullptr, (OUT & 2) ? out_hi + o_split + (ES / 4) * e : nullptr,
                       (OUT & 2) ? out_lo + o_split + (ES / 4) * e : nullptr);
    o_f32 += step;
    o_split += step;
    acc = vec_zero<VEC>();
  }
}

// rare column-tap sets (a hole between 2 or 3 taps): the two-row-cache walk, out of line
template <int VEC, int ES>
__device__ __noinline__ void sepp_hole_walk(float sh, float bh, int ph, int H, uint32_t row_pitch, const char* base, const ColTaps& ct,
                                            float* of, __nv_bfloat16* oh, __nv_bfloat16* ol, uint32_t step) {
  __shared__ RowTap rows_w[8][32];
  // every thread of a warp is here together (q is warp-uniform) but not every warp of the CTA: no CTA barrier, a private
  // table per warp (<= 7 warps: 224 threads), lane s computes sample s
  RowTap* rows = rows_w[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  if (lane < 2 * ph) rows[lane] = row_tap_sn2(sh, bh, lane, H, row_pitch);
  __syncwarp();
  const char* const pc[4] = {base + ct.off[0], base + ct.off[1], base + ct.off[2], base + ct.off[3]};
  const float w[4] = {ct.w[0], ct.w[1], ct.w[2], ct.w[3]};
  if (ct.n == 2) roi_column_walk<2, VEC, ES>(rows, ph, pc, w, of, oh, ol, step);
  else if (ct.n == 3) roi_column_walk<3, VEC, ES>(rows, ph, pc, w, of, oh, ol, step);
  else roi_column_walk<4, VEC, ES>(rows, ph, pc, w, of, oh, ol, step);
}

template <int VEC, int ES, int OUT, bool PF>
__device__ __forceinline__ void roi_align_sepp_body(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  __shared__ __align__(16) RowStep prog[33];
  const int roi = blockIdx.x;
  const RoiGeom g = roi_geom(rois + (size_t)roi * 5, scale, ph, pw, n_imgs);
  if (threadIdx.x < 32) build_row_program(prog, g.sh, g.bh, ph, H, (uint32_t)(W * C) * 4u);
  const int cg = C / (4 * VEC);
  const int q = threadIdx.x / cg, c4 = (threadIdx.x - q * cg) * (ES == 16 ? VEC : 1);
  const ColTaps ct = col_taps_sn2(g.sw, g.bw, q, W, (uint32_t)C * 4u);
  __syncthreads();
  const char* base = reinterpret_cast<const char*>(feat + (size_t)g.b * H * W * C + c4 * 4);
  const uint32_t step = (uint32_t)(pw * C);
  const uint32_t o_f32 = (uint32_t)((roi * ph * pw + q) * C + c4 * 4);            // < 2^32 bytes: checked by the launcher
  const uint32_t o_split = (uint32_t)(roi * (int)ld_split + q * C + c4 * 4);
  if (ct.n >= 1) {
    constexpr int CP = ES * VEC * 4 / 4;              // column pitch in bytes = C * 4 (lane-interleaved: C = ES * VEC / 4)
    const float w[4] = {ct.w[0], ct.w[1], ct.w[2], ct.w[3]};
    bool consec = true;
    for (int k = 1; k < ct.n; ++k) consec = consec && ct.off[k] == ct.off[0] + (uint32_t)k * CP;
    const char* pa = base + ct.off[0];
    if (consec) {
      switch (ct.n) {
        case 1: roi_column_walk_prog<1, VEC, ES, CP, false, OUT, PF>(prog, ph, pa, nullptr, w, out, out_hi, out_lo, o_f32, o_split, step); break;
        case 2: roi_column_walk_prog<2, VEC, ES, CP, false, OUT, PF>(prog, ph, pa, nullptr, w, out, out_hi, out_lo, o_f32, o_split, step); break;
        case 3: roi_column_walk_prog<3, VEC, ES, CP, false, OUT, PF>(prog, ph, pa, nullptr, w, out, out_hi, out_lo, o_f32, o_split, step); break;
        default: roi_column_walk_prog<4, VEC, ES, CP, false, OUT, PF>(prog, ph, pa, nullptr, w, out, out_hi, out_lo, o_f32, o_split, step); break;
      }
    } else if (ct.n == 4 && ct.off[1] == ct.off[0] + (uint32_t)CP && ct.off[3] == ct.off[2] + (uint32_t)CP) {
      // columns a, a+1, b, b+1 (bins wider than 4 columns; the only way col_taps_sn2 produces four taps with a hole)
      roi_column_walk_prog<4, VEC, ES, CP, true, OUT, PF>(prog, ph, pa, base + ct.off[2], w, out, out_hi, out_lo, o_f32, o_split, step);
    } else {                                          // any other tap set with a hole (a sample exactly on a column): the generic walk
      sepp_hole_walk<VEC, ES>(g.sh, g.bh, ph, H, (uint32_t)(W * C) * 4u, base, ct, (OUT & 1) ? out + o_f32 : nullptr,
                              (OUT & 2) ? out_hi + o_split : nullptr, (OUT & 2) ? out_lo + o_split : nullptr, step);
    }
    return;
  }
  // no column tap at all (every x sample outside the map): the column is zero
  const Vec<VEC> z = vec_zero<VEC>();
  for (int p = 0; p < ph; ++p) {
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      store_bin_packed(z.v[e], (OUT & 1) ? out + o_f32 + p * step + (ES / 4) * e : nullptr,
                       (OUT & 2) ? out_hi + o_split + p * step + (ES / 4) * e : nullptr,
                       (OUT & 2) ? out_lo + o_split + p * step + (ES / 4) * e : nullptr);
  }
}

template <int VEC, int ES = 16>
__device__ __forceinline__ void roi_align_sep_body(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  __shared__ RowTap rows[32];
  const int roi = blockIdx.x;
  const RoiGeom g = roi_geom(rois + (size_t)roi * 5, scale, ph, pw, n_imgs);
  if ((int)threadIdx.x < 2 * ph) rows[threadIdx.x] = row_tap_sn2(g.sh, g.bh, threadIdx.x, H, (uint32_t)(W * C) * 4u);
  const int cg = C / (4 * VEC);                       // channel groups of 4 * VEC channels
  const int q = threadIdx.x / cg, c4 = (threadIdx.x - q * cg) * (ES == 16 ? VEC : 1);   // first float4 of the thread
  const ColTaps ct = col_taps_sn2(g.sw, g.bw, q, W, (uint32_t)C * 4u);
  __syncthreads();
  const char* base = reinterpret_cast<const char*>(feat + (size_t)g.b * H * W * C + c4 * 4);
  const size_t o_f32 = ((size_t)roi * ph * pw + q) * C + c4 * 4;
  const size_t o_split = (size_t)roi * ld_split + (size_t)q * C + c4 * 4;
  const size_t step = (size_t)pw * C;
  if (ct.n >= 1) {
    const char* const pc[4] = {base + ct.off[0], base + ct.off[1], base + ct.off[2], base + ct.off[3]};
    const float w[4] = {ct.w[0], ct.w[1], ct.w[2], ct.w[3]};
    float* of = out ? out + o_f32 : nullptr;
    __nv_bfloat16* oh = out_hi ? out_hi + o_split : nullptr;
    __nv_bfloat16* ol = out_hi ? out_lo + o_split : nullptr;
    switch (ct.n) {
      case 1: roi_column_walk<1, VEC, ES>(rows, ph, pc, w, of, oh, ol, step); break;
      case 2: roi_column_walk<2, VEC, ES>(rows, ph, pc, w, of, oh, ol, step); break;
      case 3: roi_column_walk<3, VEC, ES>(rows, ph, pc, w, of, oh, ol, step); break;
      default: roi_column_walk<4, VEC, ES>(rows, ph, pc, w, of, oh, ol, step); break;
    }
    return;
  }
  // no column tap at all (every x sample outside the map): zeros, through the generic core
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const TapLoad ld{base + ES * e};
    roi_column_sep_sn2(rows, ph, ct, ld,
                       [&](int p, const float4& v) {
                         store_bin(v, out, out_hi, out_lo, o_f32 + (ES / 4) * e + p * step, o_split + (ES / 4) * e + p * step);
                       },
                       nullptr);
  }
}

// MINB = 2: 63 registers, two CTAs (28 warps) per SM.  MINB = 3: capped at 48 registers for three CTAs (42 warps) per SM
// (test hook hvr_debug_roi_variant(3); the kernel is bound by exposed L2 latency once the instruction count is down).
__global__ void __launch_bounds__(448, 2) roi_align_sep_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sep_body<1>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}
// thread = (output column, 8 channels): pw * C/8 <= 224 threads per RoI
__global__ void __launch_bounds__(224, 3) roi_align_sep8_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sep_body<2>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}
// the same with the thread's vectors lane-interleaved (C == 256): vector e = channels [128 e + 4 lane, +4) - every warp-wide
// LDG.128 reads 512 contiguous bytes, every STG.64 writes 256
__global__ void __launch_bounds__(224, 3) roi_align_sep8i_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sep_body<2, 512>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}
// the row-program walk, 8 lane-interleaved channels per thread (C == 256, 2 * ph <= 32).  The pipeline's launch is
// <OUT = 2 (split rows), PF = false, MINB = 4>: 551 us on the bench launch; PF = true measured 710 us, MINB = 3 610 us.
template <int OUT, bool PF, int MINB>
__global__ void __launch_bounds__(224, MINB) roi_align_sepp_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sepp_body<2, 512, OUT, PF>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}
// 16 channels per thread, interleaved (vector e = channels [64 e + 4 lane, +4)): experiments (hvr_debug_roi_variant(12))
__global__ void __launch_bounds__(112, 4) roi_align_sep16i_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sep_body<4, 256>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}
__global__ void __maxnreg__(48) roi_align_sep48_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, int n_imgs, int C, int H, int W, int ph, int pw,
    float scale, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  roi_align_sep_body<1>(feat, rois, n_imgs, C, H, W, ph, pw, scale, out, out_hi, out_lo, ld_split);
}

// Slab variant of the fast path.  ncu of the RoI-per-CTA kernels above (profiles/r01q_*, r02a_*): every variant
// sits at ~7.3 TB/s of L2 -> SM traffic (lts__t_sectors_srcunit_tex_op_read: 7.3 GB per 105-frame launch, the chip's
// L2 throughput cap) - each RoI drags its own patch of the map (~200 KB on the bench distribution) out of L2,
// although the 300 RoIs of a frame overlap 26-fold.  Here the MAP is what a CTA owns: CTA = (frame, 16-channel
// slab); the slab of the whole map (H*W pixels x 64 B = 153 KB for 38 x 63) is copied into shared memory ONCE
// (cp.async) and every RoI of the frame is evaluated from it, so the L2 -> SM traffic of a launch drops to one
// pass over the maps (257 MB) and all taps are LDS.128.  One warp per RoI: lane = (output column q, 4-channel
// group) for 7 x 4 = 28 lanes, lanes 0..2*ph-1 compute the RoI's y samples into a per-warp table first; the
// separable evaluation of roi_align_sep.cuh then walks the rows.  RoIs are claimed dynamically (one shared
// counter per CTA).  Needs the RoIs bucketed by frame: roi_bucket_kernel (one small launch) writes `order` / `start`.
constexpr int SLAB_C = 16;           // channels per slab: 64-byte pixels, 32-byte output sectors per (RoI, bin)
constexpr int SLAB_THREADS = 512;

__global__ void roi_bucket_kernel(const float* __restrict__ rois, int n, int n_imgs, int* __restrict__ start,
                                  int* __restrict__ order) {
  extern __shared__ int sb[];        // [n_imgs] counts, [n_imgs] cursors
  int* cnt = sb;
  int* cur = sb + n_imgs;
  for (int i = threadIdx.x; i < n_imgs; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  auto img_of = [&](int i) {
    const int b = (int)rois[(size_t)i * 5];
    return b < 0 ? 0 : (b >= n_imgs ? n_imgs - 1 : b);
  };
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&cnt[img_of(i)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < n_imgs; ++b) {
      start[b] = acc;
      cur[b] = acc;
      acc += cnt[b];
    }
    start[n_imgs] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&cur[img_of(i)], 1)] = i;
}

struct SlabLoad {
  const char* base;   // slab + 4-channel group
  __device__ __forceinline__ float4 operator()(uint32_t off) const {
    return *reinterpret_cast<const float4*>(base + off);
  }
};

__global__ void __launch_bounds__(SLAB_THREADS, 1) roi_align_slab_kernel(
    const float* __restrict__ feat, const float* __restrict__ rois, const int* __restrict__ start,
    const int* __restrict__ order, int n_imgs, int C, int H, int W, int ph, int pw, float scale,
    float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
    long long ld_split) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int HW = H * W;
  uint8_t* slab = smem;                                                      // [HW][16] fp32
  RowTap* rtab = reinterpret_cast<RowTap*>(smem + (size_t)HW * SLAB_C * 4);  // [warps][32]
  __shared__ int next;
  const int img = blockIdx.y, c0 = blockIdx.x * SLAB_C;
  const int first = start[img], last = start[img + 1];
  if (first == last) return;
  // slab of the whole map: 4 x 16 B per pixel
  const float* src = feat + (size_t)img * HW * C + c0;
  for (int i = threadIdx.x; i < HW * 4; i += SLAB_THREADS) {
    const float* g = src + (size_t)(i >> 2) * C + (i & 3) * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slab + (size_t)i * 16)), "l"(g) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (threadIdx.x == 0) next = first;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2, cg = lane & 3;
  RowTap* rows = rtab + warp * 32;
  const SlabLoad ld{reinterpret_cast<const char*>(slab) + cg * 16};
  const uint32_t row_pitch = (uint32_t)W * SLAB_C * 4u;
  const size_t step = (size_t)pw * C;
  for (;;) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= last) break;
    const int roi = order[r];
    const RoiGeom g = roi_geom(rois + (size_t)roi * 5, scale, ph, pw, n_imgs);
    __syncwarp();                                   // the previous RoI's table reads are done
    if (lane < 2 * ph) rows[lane] = row_tap_sn2(g.sh, g.bh, lane, H, row_pitch);
    __syncwarp();
    if (q < pw) {
      const ColTaps ct = col_taps_sn2(g.sw, g.bw, q, W, (uint32_t)SLAB_C * 4u);
      const size_t o_f32 = ((size_t)roi * ph * pw + q) * C + c0 + cg * 4;
      const size_t o_split = (size_t)roi * ld_split + (size_t)q * C + c0 + cg * 4;
      roi_column_sep_sn2(rows, ph, ct, ld,
                         [&](int p, const float4& v) { store_bin(v, out, out_hi, out_lo, o_f32 + p * step, o_split + p * step); },
                         nullptr);
    }
  }
}

// Generic path (adaptive sample_num == 0, or very large sampling grids): reference-style, one
// thread per output element, NHWC or NCHW output.
__global__ void roi_align_generic_kernel(const float* __restrict__ feat, const float* __restrict__ rois, int n_rois,
                                         int n_imgs, int C, int H, int W, int ph, int pw, float scale, int sn,
                                         float* __restrict__ out, int nchw_out) {
  const size_t total = (size_t)n_rois * C * ph * pw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c, q, p, n;
    if (nchw_out) {
      q = (int)(i % pw); p = (int)((i / pw) % ph); c = (int)((i / ((size_t)pw * ph)) % C);
      n = (int)(i / ((size_t)pw * ph * C));
    } else {
      c = (int)(i % C); q = (int)((i / C) % pw); p = (int)((i / ((size_t)C * pw)) % ph);
      n = (int)(i / ((size_t)C * pw * ph));
    }
    const float* r = rois + (size_t)n * 5;
    const RoiGeom g = roi_geom(r, scale, ph, pw, n_imgs);
    const float ew = (r[3] + 1.0f) * scale, eh = (r[4] + 1.0f) * scale;
    const float rw = fmaxf(ew - g.sw, 0.0f), rh = fmaxf(eh - g.sh, 0.0f);
    const int nh = sn > 0 ? sn : (int)ceilf(rh / (float)ph);
    const int nw = sn > 0 ? sn : (int)ceilf(rw / (float)pw);
    const float* fm = feat + (size_t)g.b * H * W * C;
    float acc = 0.f;
    for (int iy = 0; iy < nh; ++iy) {
      const float y = g.sh + (float)p * g.bh + ((float)iy + 0.5f) * g.bh / (float)nh;
      for (int ix = 0; ix < nw; ++ix) {
        const float x = g.sw + (float)q * g.bw + ((float)ix + 0.5f) * g.bw / (float)nw;
        const Tap t = make_tap(y, x, H, W, C);
        if (t.o0 == kTapInvalid) continue;
        acc = acc + (((t.w1 * fm[(t.o0 >> 2) + c] + t.w2 * fm[(t.o1 >> 2) + c]) + t.w3 * fm[(t.o2 >> 2) + c]) +
                     t.w4 * fm[(t.o3 >> 2) + c]);
      }
    }
    out[i] = acc / (float)(nh * nw);
  }
}

int g_roi_variant = 0;   // test hook (hvr_debug_roi_variant): 0 = heuristic, 1 = always the per-bin kernel
int g_sep_minb = 2;      // test hook (hvr_debug_roi_variant 2 / 3): resident CTAs per SM the fast kernel is built for
bool g_sep_vec8 = true;  // test hook (hvr_debug_roi_variant 7 / 8): 8 channels per thread on (default: 835 vs 915 us on the bench launch) / off
int g_sep_layout = 3;    // test hook (hvr_debug_roi_variant 10 .. 14): channels of a thread adjacent / lane-interleaved / 16 interleaved /
                         // lane-interleaved with the row-program walk (default) / the same + L1 prefetch of the next step's rows /
                         // the same built for 3 instead of 4 resident CTAs per SM
int g_slab = 0;          // test hook (hvr_debug_roi_variant 4 / 5 / 6): 0 = heuristic (= never), 1 = slab kernel whenever it applies, 2 = never

}  // namespace

extern "C" int hvr_debug_roi_variant(int v) {
  if (v == 2 || v == 3) {          // occupancy variant of the fast kernel (experiments)
    g_sep_minb = v;
    return HVR_OK;
  }
  if (v == 7 || v == 8) {
    g_sep_vec8 = v == 7;
    return HVR_OK;
  }
  if (v >= 10 && v <= 15) {
    g_sep_layout = v - 10;
    return HVR_OK;
  }
  if (v >= 4 && v <= 6) {          // 4 = heuristic, 5 = slab kernel whenever it applies, 6 = never the slab kernel
    g_slab = v - 4;
    return HVR_OK;
  }
  if (v != 0 && v != 1) return HVR_ERR_ARG;
  g_roi_variant = v;
  return HVR_OK;
}

extern "C" size_t hvr_roi_align_fast_workspace_bytes(int n_rois, int n_imgs) {
  return ((size_t)(n_rois > 0 ? n_rois : 0) + (size_t)(n_imgs > 0 ? n_imgs : 0) + 1) * sizeof(int) + 256;
}

// Fast arithmetic (roi_align_sep.cuh) where it applies - NHWC rows out, sample_num == 2 - else the strict kernels
// of hvr_roi_align_fwd.  Two launch shapes: the slab kernel (CTA = frame x 16-channel slab, map slab in shared
// memory) when the map slab fits and the launch has at least one CTA per SM; otherwise one CTA of pw * C/4 <= 448
// threads per RoI.
extern "C" int hvr_roi_align_fwd_fast(const float* feat, int feat_nhwc, const float* rois, int n_rois, int n_imgs,
                                      int C, int H, int W, int ph, int pw, float spatial_scale, int sample_num,
                                      float* out, int out_layout, hvr_bf16* out_hi, hvr_bf16* out_lo,
                                      int64_t ld_split, void* ws, size_t ws_bytes, void* stream) {
  const bool fast_ok = feat_nhwc && out_layout == 1 && sample_num == 2 && C >= 4 && C % 4 == 0 && 2 * ph <= 32 &&
                       g_roi_variant == 0;
  const bool sep = fast_ok && pw * (C >> 2) <= 448;
  const size_t slab_smem = (size_t)H * W * SLAB_C * 4 + (SLAB_THREADS / 32) * 32 * sizeof(RowTap);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    HVR_CUDA(cudaGetDevice(&dev));
    HVR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool slab = fast_ok && g_slab != 2 && C % SLAB_C == 0 && pw * 4 <= 32 && slab_smem <= 220 * 1024 && n_imgs <= 65535 &&
                    (size_t)n_imgs * 8 <= 96 * 1024 && ws != nullptr &&
                    ws_bytes >= hvr_roi_align_fast_workspace_bytes(n_rois, n_imgs) &&
                    g_slab == 1;   // measured (profiles/r02c_roi_align_variants.txt): instruction-bound, 1.4x slower than the
                                   // RoI-per-CTA shape at 105 frames -> never chosen by the heuristic, kept behind the hook
  if (!sep && !slab)
    return hvr_roi_align_fwd(feat, feat_nhwc, rois, n_rois, n_imgs, C, H, W, ph, pw, spatial_scale, sample_num, out,
                             out_layout, out_hi, out_lo, ld_split, feat_nhwc ? nullptr : (float*)ws, stream);
  if (n_rois == 0) return HVR_OK;
  if (!feat || !rois || n_rois < 0 || n_imgs < 1 || H < 1 || W < 1 || ph < 1 || pw < 1) return HVR_ERR_ARG;
  if (!out && !out_hi) return HVR_ERR_ARG;
  if ((out_hi == nullptr) != (out_lo == nullptr)) return HVR_ERR_ARG;
  if (out_hi && (ld_split < (int64_t)ph * pw * C || ld_split % 4 != 0)) return HVR_ERR_ARG;
  if ((size_t)H * W * C >= (1u << 30)) return HVR_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (slab) {
    int* start = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(ws) + 15) & ~(uintptr_t)15);
    int* order = start + n_imgs + 1;
    static bool attr = false;
    if (!attr) {
      HVR_CUDA(cudaFuncSetAttribute(roi_align_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      HVR_CUDA(cudaFuncSetAttribute(roi_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    roi_bucket_kernel<<<1, 1024, (size_t)n_imgs * 8, st>>>(rois, n_rois, n_imgs, start, order);
    HVR_LAUNCHED();
    roi_align_slab_kernel<<<dim3(C / SLAB_C, n_imgs), SLAB_THREADS, slab_smem, st>>>(
        feat, rois, start, order, n_imgs, C, H, W, ph, pw, spatial_scale, out, (__nv_bfloat16*)out_hi,
        (__nv_bfloat16*)out_lo, ld_split);
    HVR_LAUNCHED();
    return HVR_OK;
  }
  const int threads = pw * (C >> 2);
  if (g_sep_vec8 && g_sep_layout != 0 && C == 256 && pw <= 7) {
    const size_t out_bytes = (size_t)n_rois * (out_hi ? (size_t)ld_split * 2 : 0) > (size_t)n_rois * ph * pw * C * (out ? 4 : 0)
                                 ? (size_t)n_rois * ld_split * 2 : (size_t)n_rois * ph * pw * C * 4;
    if (g_sep_layout >= 3 && 2 * ph <= 32 && out_bytes < ((size_t)1 << 31)) {
      const int o = (out ? 1 : 0) | (out_hi ? 2 : 0);
      // 4 resident CTAs per SM (72 registers) is the default: 551 us on the bench launch against 609 us with 3 (80 registers)
      // and 630 us with 5 (56 registers, spills)
      auto kern = g_sep_layout == 5 ? (o == 1 ? roi_align_sepp_kernel<1, false, 3> : o == 2 ? roi_align_sepp_kernel<2, false, 3> : roi_align_sepp_kernel<3, false, 3>) :
                  g_sep_layout == 4 ? (o == 1 ? roi_align_sepp_kernel<1, true, 4> : o == 2 ? roi_align_sepp_kernel<2, true, 4> : roi_align_sepp_kernel<3, true, 4>)
                                    : (o == 1 ? roi_align_sepp_kernel<1, false, 4> : o == 2 ? roi_align_sepp_kernel<2, false, 4> : roi_align_sepp_kernel<3, false, 4>);
      kern<<<n_rois, pw * 32, 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out, (__nv_bfloat16*)out_hi,
                                       (__nv_bfloat16*)out_lo, ld_split);
    }
    else if (g_sep_layout == 2)
      roi_align_sep16i_kernel<<<n_rois, pw * 16, 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                          (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
    else
      roi_align_sep8i_kernel<<<n_rois, pw * 32, 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                         (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
    HVR_LAUNCHED();
    return HVR_OK;
  }
  if (g_sep_vec8 && C % 8 == 0 && pw * (C >> 3) <= 224) {
    roi_align_sep8_kernel<<<n_rois, pw * (C >> 3), 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                           (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
    HVR_LAUNCHED();
    return HVR_OK;
  }
  if (g_sep_minb == 3)
    roi_align_sep48_kernel<<<n_rois, threads, 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                       (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
  else
    roi_align_sep_kernel<<<n_rois, threads, 0, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                     (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
  HVR_LAUNCHED();
  return HVR_OK;
}
extern "C" int hvr_roi_align_fwd(const float* feat, int feat_nhwc, const float* rois, int n_rois, int n_imgs, int C,
                                 int H, int W, int ph, int pw, float spatial_scale, int sample_num, float* out,
                                 int out_layout, hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_split, float* ws,
                                 void* stream) {
  if (n_rois == 0) return HVR_OK;   // empty in, empty out (the caller's output tensor has no rows)
  if (!feat || !rois || n_rois < 0 || n_imgs < 1 || C < 1 || H < 1 || W < 1 || ph < 1 || pw < 1) return HVR_ERR_ARG;
  if (!out && !out_hi) return HVR_ERR_ARG;
  if ((out_hi == nullptr) != (out_lo == nullptr)) return HVR_ERR_ARG;
  if (out_layout != 0 && out_layout != 1) return HVR_ERR_ARG;
  if (out_hi && (out_layout != 1 || ld_split < (int64_t)ph * pw * C || ld_split % 4 != 0)) return HVR_ERR_ARG;
  if ((size_t)H * W * C >= (1u << 30)) return HVR_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!feat_nhwc) {
    if (!ws) return HVR_ERR_WORKSPACE;
    int rc = hvr_nchw_to_nhwc_f32(feat, n_imgs, C, H, W, ws, stream);
    if (rc) return rc;
    feat = ws;
  }
  const int nsamp = ph * pw * sample_num * sample_num;
  const bool fast = sample_num > 0 && nsamp <= 2048 && (out_layout == 0 || C % 4 == 0);
  if (!fast) {
    if (out_hi || !out) return HVR_ERR_UNSUPPORTED;
    const size_t total = (size_t)n_rois * C * ph * pw;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    roi_align_generic_kernel<<<(int)blocks, 256, 0, st>>>(feat, rois, n_rois, n_imgs, C, H, W, ph, pw, spatial_scale,
                                                          sample_num, out, out_layout == 0);
    HVR_LAUNCHED();
    return HVR_OK;
  }
  const int cg = C >> 2;
  if (out_layout == 1 && sample_num == 2 && g_roi_variant == 0 && 256 % cg == 0) {
    const size_t smem = (size_t)nsamp * sizeof(Tap);
    static bool attr2 = false;
    if (!attr2) {
      HVR_CUDA(cudaFuncSetAttribute(roi_align_sn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      attr2 = true;
    }
    roi_align_sn2_kernel<<<n_rois, 256, smem, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, out,
                                                    (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
  } else if (out_layout == 1) {
    const size_t smem = (size_t)nsamp * sizeof(Tap);
    static bool attr1 = false;
    if (!attr1) {
      HVR_CUDA(cudaFuncSetAttribute(roi_align_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      attr1 = true;
    }
    roi_align_kernel<false><<<n_rois, 256, smem, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, sample_num,
                                                       out, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_split);
  } else {
    const size_t smem = (size_t)nsamp * sizeof(Tap) + (size_t)256 * ph * pw * sizeof(float);
    if (smem > 200 * 1024) return HVR_ERR_UNSUPPORTED;
    static bool attr0 = false;
    if (!attr0) {
      HVR_CUDA(cudaFuncSetAttribute(roi_align_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr0 = true;
    }
    roi_align_kernel<true><<<n_rois, 256, smem, st>>>(feat, rois, n_imgs, C, H, W, ph, pw, spatial_scale, sample_num,
                                                      out, nullptr, nullptr, 0);
  }
  HVR_LAUNCHED();
  return HVR_OK;
}
