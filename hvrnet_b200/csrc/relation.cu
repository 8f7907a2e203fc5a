// Host-facing composites of the C ABI for non-Python callers: weight packing and one whole relation
// block behind one call.  Everything here is orchestration of the kernels the Python engine launches one by one
// (hvrnet_b200/engine.py pack_* / relation): the same descriptors, the same launch sequence, the same bits.
//
//   hvr_pack_conv_bn   conv (+ frozen BN, + conv bias) -> K-major split weights + fp32 bias   (engine._pack_conv_bn)
//   hvr_pack_linear    nn.Linear / 1x1 conv weights     -> split weights + fp32 bias           (engine.pack_linear)
//   hvr_relation_fwd   q/k projections, QK^T / sqrt(D), row softmax, P.X, output projection + residual + ReLU
//                      (hrnmp_bbox_head.py:216-355 forward_single_selsa; selsa_bbox_head.py:108-201)
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

inline int64_t rup(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// hi = bf16_rn(x), lo = bf16_rn(x - float(hi)): the host twin of split2 (same IEEE operations, same bits)
inline void split_host(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  memcpy(&hi, &h, 2);
  memcpy(&lo, &l, 2);
}

int upload_split(const std::vector<float>& W, int64_t rows_pad, int64_t cols_pad, hvr_bf16* w_hi, hvr_bf16* w_lo,
                 cudaStream_t st) {
  std::vector<uint16_t> hi((size_t)rows_pad * cols_pad), lo((size_t)rows_pad * cols_pad);
  for (size_t i = 0; i < hi.size(); ++i) split_host(W[i], hi[i], lo[i]);
  HVR_CUDA(cudaMemcpyAsync(w_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice, st));
  HVR_CUDA(cudaMemcpyAsync(w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice, st));
  HVR_CUDA(cudaStreamSynchronize(st));          // the staging vectors die with this frame
  return HVR_OK;
}

void fill_linear(HvrIGemm& g, const hvr_bf16* a_hi, const hvr_bf16* a_lo, int M, int K, int64_t lda,
                 const hvr_bf16* b_hi, const hvr_bf16* b_lo, int n, int64_t ldb) {
  memset(&g, 0, sizeof(g));
  g.a_hi = a_hi; g.a_lo = a_lo;
  g.a_c = K; g.a_w = M; g.a_h = 1; g.a_b = 1;
  g.a_stride_w = lda; g.a_stride_h = lda * M; g.a_stride_b = lda * M;
  g.ntaps = 1;
  g.out_w = M; g.out_h = 1; g.batch = 1;
  g.tile_w = 128; g.tile_h = 1;                 // ops.pick_tile for a [M, 1] pixel grid
  g.b_hi = b_hi; g.b_lo = b_lo; g.n = n; g.ldb = ldb;
  g.alpha = 1.0f;
  g.passes = 3;
}

}  // namespace

extern "C" size_t hvr_packed_rows(int n) { return (size_t)rup(n, 64); }
extern "C" size_t hvr_packed_cols(int k) { return (size_t)rup(k, 64); }

extern "C" int hvr_pack_conv_bn(const float* w_host, const float* bn_weight_host, const float* bn_bias_host,
                                const float* bn_mean_host, const float* bn_var_host, float bn_eps,
                                const float* conv_bias_host, int cout, int cin, int kh, int kw, hvr_bf16* w_hi,
                                hvr_bf16* w_lo, float* bias, void* stream) {
  if (!w_host || !w_hi || !w_lo || cout < 1 || cin < 1 || kh < 1 || kw < 1) return HVR_ERR_ARG;
  const bool bn = bn_weight_host != nullptr;
  if (bn && (!bn_bias_host || !bn_mean_host || !bn_var_host)) return HVR_ERR_ARG;
  if ((bn || conv_bias_host) && !bias) return HVR_ERR_ARG;
  const int64_t k = (int64_t)kh * kw * cin, rows_pad = rup(cout, 64), cols_pad = rup(k, 64);
  std::vector<float> W((size_t)rows_pad * cols_pad, 0.f);
  std::vector<float> B((size_t)rows_pad, 0.f);
  for (int o = 0; o < cout; ++o) {
    // frozen BN folded in fp64: scale = gamma / sqrt(var + eps), shift = beta - mean * scale (engine._bn_fold)
    double scale = 1.0, shift = 0.0;
    if (bn) {
      scale = (double)bn_weight_host[o] / std::sqrt((double)bn_var_host[o] + (double)bn_eps);
      shift = (double)bn_bias_host[o] - (double)bn_mean_host[o] * scale;
    }
    // [Cout, Cin, kh, kw] -> K-major [Cout, (r*kw + s)*Cin + c]
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s) {
          const double v = (double)w_host[(((size_t)o * cin + c) * kh + r) * kw + s] * scale;
          W[(size_t)o * cols_pad + ((size_t)r * kw + s) * cin + c] = (float)v;
        }
    double b = bn ? shift : 0.0;
    if (conv_bias_host) b = bn ? shift + (double)conv_bias_host[o] * scale : (double)conv_bias_host[o];
    B[o] = (float)b;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bias) HVR_CUDA(cudaMemcpyAsync(bias, B.data(), B.size() * 4, cudaMemcpyHostToDevice, st));
  return upload_split(W, rows_pad, cols_pad, w_hi, w_lo, st);
}

extern "C" int hvr_pack_linear(const float* w_host, const float* bias_host, int n, int k, const int* col_perm_host,
                               hvr_bf16* w_hi, hvr_bf16* w_lo, float* bias, void* stream) {
  if (!w_host || !w_hi || !w_lo || !bias || n < 1 || k < 1) return HVR_ERR_ARG;
  const int64_t rows_pad = rup(n, 64), cols_pad = rup(k, 64);
  std::vector<float> W((size_t)rows_pad * cols_pad, 0.f);
  std::vector<float> B((size_t)rows_pad, 0.f);
  for (int o = 0; o < n; ++o) {
    for (int c = 0; c < k; ++c) W[(size_t)o * cols_pad + c] = w_host[(size_t)o * k + (col_perm_host ? col_perm_host[c] : c)];
    if (bias_host) B[o] = bias_host[o];
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVR_CUDA(cudaMemcpyAsync(bias, B.data(), B.size() * 4, cudaMemcpyHostToDevice, st));
  return upload_split(W, rows_pad, cols_pad, w_hi, w_lo, st);
}

// ------------------------------------------------------------------------------------------------
// One nn.Linear / nn.Conv2d (+ folded BN, + residual, + ReLU) on packed weights: the descriptor filling of
// engine.lin / engine.conv (tile choice, tap offsets, strided views) behind plain arguments.
// ------------------------------------------------------------------------------------------------
extern "C" int hvr_linear_fwd(const hvr_bf16* x_hi, const hvr_bf16* x_lo, int rows, int k, int64_t ld_x,
                              const hvr_bf16* w_hi, const hvr_bf16* w_lo, const float* bias, int n,
                              const hvr_bf16* res_hi, const hvr_bf16* res_lo, int64_t ld_res, int relu, float alpha,
                              hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, float* out_f32, int64_t ld_f32,
                              void* stream) {
  if (!x_hi || !x_lo || !w_hi || !w_lo || rows < 1 || k < 8 || n < 1) return HVR_ERR_ARG;
  HvrIGemm g;
  fill_linear(g, x_hi, x_lo, rows, k, ld_x, w_hi, w_lo, n, rup(k, 64));
  g.alpha = alpha;
  g.bias = bias;
  g.res_hi = res_hi; g.res_lo = res_lo; g.ld_res = ld_res;
  g.relu = relu;
  g.out_hi = out_hi; g.out_lo = out_lo; g.ld_out = ld_out;
  g.out_f32 = out_f32; g.ld_f32 = ld_f32;
  return hvr_igemm(&g, stream);
}

extern "C" int hvr_conv_fwd(const hvr_bf16* x_hi, const hvr_bf16* x_lo, int batch, int h, int w, int cin,
                            const hvr_bf16* w_hi, const hvr_bf16* w_lo, const float* bias, int cout, int ksize,
                            int dilation, int stride, const hvr_bf16* res_hi, const hvr_bf16* res_lo, int relu,
                            hvr_bf16* out_hi, hvr_bf16* out_lo, float* out_f32, void* stream) {
  if (!x_hi || !x_lo || !w_hi || !w_lo || batch < 1 || h < 1 || w < 1 || cin < 8 || cout < 1) return HVR_ERR_ARG;
  if ((ksize != 1 && ksize != 3) || dilation < 1 || stride < 1) return HVR_ERR_ARG;
  if (stride != 1 && ksize != 1) return HVR_ERR_UNSUPPORTED;   // strided convs on the path are 1x1 (caffe-style bottleneck)
  if (!out_hi && !out_f32) return HVR_ERR_ARG;
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const int64_t npad = rup(cout, 64);                           // physical rows of the packed weights = output row pitch
  HvrIGemm g;
  memset(&g, 0, sizeof(g));
  g.a_hi = x_hi; g.a_lo = x_lo;
  g.a_c = cin; g.a_w = wo; g.a_h = ho; g.a_b = batch;          // strided view of the NHWC input: every stride-th pixel
  g.a_stride_w = (int64_t)cin * stride;
  g.a_stride_h = (int64_t)w * cin * stride;
  g.a_stride_b = (int64_t)h * w * cin;
  if (stride == 1) { g.a_w = w; g.a_h = h; }
  const int r = ksize / 2;
  g.ntaps = ksize * ksize;
  for (int t = 0; t < ksize; ++t)
    for (int s2 = 0; s2 < ksize; ++s2) {                        // (dx, dy), row-major over the filter (engine._taps)
      g.tap_dx[t * ksize + s2] = (s2 - r) * dilation;
      g.tap_dy[t * ksize + s2] = (t - r) * dilation;
    }
  g.out_w = wo; g.out_h = ho; g.batch = batch;
  // ops.pick_tile: the 128-pixel tile shape that wastes the fewest rows
  long long best = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    const long long waste = (long long)((wo + tw - 1) / tw * tw) * ((ho + th - 1) / th * th);
    if (best < 0 || waste < best) { best = waste; g.tile_w = tw; g.tile_h = th; }
  }
  g.b_hi = w_hi; g.b_lo = w_lo; g.n = (int)npad; g.ldb = rup((int64_t)ksize * ksize * cin, 64);
  g.alpha = 1.0f;
  g.bias = bias;
  g.res_hi = res_hi; g.res_lo = res_lo; g.ld_res = npad;
  g.relu = relu;
  g.out_hi = out_hi; g.out_lo = out_lo; g.ld_out = npad;
  g.out_f32 = out_f32; g.ld_f32 = rup(npad, 4);
  g.passes = 3;
  return hvr_igemm(&g, stream);
}

namespace {
struct RelWs {
  hvr_bf16 *q_hi, *q_lo, *k_hi, *k_lo, *p_hi, *p_lo, *xt_hi, *xt_lo, *o_hi, *o_lo;
  float* s;
  int64_t ld_s, ld_p;
  size_t bytes;
};
RelWs carve(void* ws, int n_q, int n_k, int D) {
  RelWs r;
  const int64_t ld_p = rup(n_k, 64), ld_s = rup(n_k, 4);
  size_t off = 0;
  auto take = [&](size_t b) {
    void* p = ws ? (void*)((uint8_t*)ws + off) : nullptr;
    off += (b + 255) / 256 * 256;
    return p;
  };
  r.q_hi = (hvr_bf16*)take((size_t)n_q * D * 2); r.q_lo = (hvr_bf16*)take((size_t)n_q * D * 2);
  r.k_hi = (hvr_bf16*)take((size_t)n_k * D * 2); r.k_lo = (hvr_bf16*)take((size_t)n_k * D * 2);
  r.s = (float*)take((size_t)n_q * ld_s * 4);
  r.p_hi = (hvr_bf16*)take((size_t)n_q * ld_p * 2); r.p_lo = (hvr_bf16*)take((size_t)n_q * ld_p * 2);
  r.xt_hi = (hvr_bf16*)take((size_t)D * ld_p * 2); r.xt_lo = (hvr_bf16*)take((size_t)D * ld_p * 2);
  r.o_hi = (hvr_bf16*)take((size_t)n_q * D * 2); r.o_lo = (hvr_bf16*)take((size_t)n_q * D * 2);
  r.ld_s = ld_s; r.ld_p = ld_p; r.bytes = off + 256;
  return r;
}
}  // namespace

extern "C" size_t hvr_relation_workspace_bytes(int n_q, int n_k, int D) {
  if (n_q < 1 || n_k < 1 || D < 1) return 0;
  return carve(nullptr, n_q, n_k, D).bytes;
}

extern "C" int hvr_relation_fwd(const HvrRelationWeights* w, const hvr_bf16* x_hi, const hvr_bf16* x_lo, int64_t ld_x,
                                int n_k, const hvr_bf16* xq_hi, const hvr_bf16* xq_lo, int64_t ld_xq, int n_q,
                                const hvr_bf16* res_hi, const hvr_bf16* res_lo, int64_t ld_res, int relu,
                                hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, void* ws, size_t ws_bytes,
                                void* stream) {
  if (!w || !x_hi || !x_lo || !out_hi || !out_lo || !w->q_hi || !w->k_hi || !w->o_hi) return HVR_ERR_ARG;
  const int D = w->dim;
  if (D < 8 || D % 8 || n_k < 1 || n_q < 1) return HVR_ERR_ARG;
  if (!xq_hi) { xq_hi = x_hi; xq_lo = x_lo; ld_xq = ld_x; if (n_q != n_k) return HVR_ERR_ARG; }
  if ((res_hi == nullptr) != (res_lo == nullptr)) return HVR_ERR_ARG;
  if (!ws || ws_bytes < hvr_relation_workspace_bytes(n_q, n_k, D)) return HVR_ERR_WORKSPACE;
  void* base = (void*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const RelWs r = carve(base, n_q, n_k, D);
  const int64_t ldw = rup(D, 64);
  HvrIGemm g;
  int rc;
  // Q = xq Wq^T + bq ; K = x Wk^T + bk                                   (hrnmp_bbox_head.py:282-291)
  fill_linear(g, xq_hi, xq_lo, n_q, D, ld_xq, w->q_hi, w->q_lo, D, ldw);
  g.bias = w->q_bias; g.out_hi = r.q_hi; g.out_lo = r.q_lo; g.ld_out = D;
  if ((rc = hvr_igemm(&g, stream))) return rc;
  fill_linear(g, x_hi, x_lo, n_k, D, ld_x, w->k_hi, w->k_lo, D, ldw);
  g.bias = w->k_bias; g.out_hi = r.k_hi; g.out_lo = r.k_lo; g.ld_out = D;
  if ((rc = hvr_igemm(&g, stream))) return rc;
  // S = Q K^T / sqrt(D)                                                   (:293-294)
  fill_linear(g, r.q_hi, r.q_lo, n_q, D, D, r.k_hi, r.k_lo, n_k, D);
  g.alpha = 1.0f / sqrtf((float)D); g.out_f32 = r.s; g.ld_f32 = r.ld_s;
  if ((rc = hvr_igemm(&g, stream))) return rc;
  // P = softmax(S, keys)                                                  (:332)
  if ((rc = hvr_softmax_rows_split(r.s, n_q, n_k, r.ld_s, r.p_hi, r.p_lo, r.ld_p, stream))) return rc;
  // O = P X (values are the un-projected rows, conv_g False): right operand X^T [D, n_k]   (:340-342)
  if ((rc = hvr_transpose_split(x_hi, x_lo, n_k, D, ld_x, r.xt_hi, r.xt_lo, r.ld_p, stream))) return rc;
  fill_linear(g, r.p_hi, r.p_lo, n_q, (int)r.ld_p, r.ld_p, r.xt_hi, r.xt_lo, D, r.ld_p);
  g.out_hi = r.o_hi; g.out_lo = r.o_lo; g.ld_out = D;
  if ((rc = hvr_igemm(&g, stream))) return rc;
  // out = [relu](res + O Wo^T + bo)                                       (:343-350 and the callers' residual adds)
  fill_linear(g, r.o_hi, r.o_lo, n_q, D, D, w->o_hi, w->o_lo, D, ldw);
  g.bias = w->o_bias; g.res_hi = res_hi; g.res_lo = res_lo; g.ld_res = ld_res; g.relu = relu;
  g.out_hi = out_hi; g.out_lo = out_lo; g.ld_out = ld_out;
  return hvr_igemm(&g, stream);
}
