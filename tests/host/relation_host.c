/* A plain-C host of the C ABI (include/hvr_b200.h): what a non-Python maintainer binds.
 * Packs the three linear layers of one relation block with hvr_pack_linear, runs hvr_relation_fwd on random rows -
 * every row a query, then key-only queries with a residual - and checks the result against a double-precision
 * restatement of forward_single_selsa (hrnmp_bbox_head.py:216-355) written out below.  Tolerance: 1e-3 relative
 * (north_star).  Test infrastructure: built and run by tests/test_gpu_kernels.py (compiled only, without a GPU, by
 * tests/test_host.py).
 *   gcc -std=c11 -O2 relation_host.c -I include -I /usr/local/cuda/include -L hvrnet_b200 -lhvr_b200 -lcudart -lm */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hvr_b200.h"

static uint32_t rng_state = 12345u;
static float frand(void) { /* uniform in [-1, 1) */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) / 8388608.0f - 1.0f;
}
static uint16_t bf16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
#define CK(call)                                                            \
  do {                                                                      \
    int rc_ = (call);                                                       \
    if (rc_ != 0) {                                                         \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, hvr_strerror(rc_)); \
      return 2;                                                             \
    }                                                                       \
  } while (0)
#define CU(call)                                                  \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) {                                      \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); \
      return 2;                                                   \
    }                                                             \
  } while (0)

typedef struct { hvr_bf16 *hi, *lo; float* bias; float *w, *b; } Lin;

static int make_linear(Lin* l, int D, float scale) {
  l->w = (float*)malloc(sizeof(float) * D * D);
  l->b = (float*)malloc(sizeof(float) * D);
  for (int i = 0; i < D * D; ++i) l->w[i] = frand() * scale;
  for (int i = 0; i < D; ++i) l->b[i] = frand() * 0.1f;
  size_t elems = hvr_packed_rows(D) * hvr_packed_cols(D);
  CU(cudaMalloc((void**)&l->hi, elems * 2));
  CU(cudaMalloc((void**)&l->lo, elems * 2));
  CU(cudaMalloc((void**)&l->bias, hvr_packed_rows(D) * 4));
  CK(hvr_pack_linear(l->w, l->b, D, D, NULL, l->hi, l->lo, l->bias, NULL));
  return 0;
}

/* y[n_q, D] = [relu](res + (softmax(Q K^T / sqrt(D)) X) Wo^T + bo) in double */
static void reference(const float* x, int n_k, const float* xq, int n_q, const float* res, int relu, int D,
                      const Lin* q, const Lin* k, const Lin* o, double* y) {
  double* Q = (double*)malloc(sizeof(double) * n_q * D);
  double* K = (double*)malloc(sizeof(double) * n_k * D);
  double* P = (double*)malloc(sizeof(double) * n_k);
  double* O = (double*)malloc(sizeof(double) * D);
  for (int i = 0; i < n_q; ++i)
    for (int j = 0; j < D; ++j) {
      double a = q->b[j];
      for (int c = 0; c < D; ++c) a += (double)xq[i * D + c] * q->w[j * D + c];
      Q[i * D + j] = a;
    }
  for (int i = 0; i < n_k; ++i)
    for (int j = 0; j < D; ++j) {
      double a = k->b[j];
      for (int c = 0; c < D; ++c) a += (double)x[i * D + c] * k->w[j * D + c];
      K[i * D + j] = a;
    }
  for (int i = 0; i < n_q; ++i) {
    double mx = -1e300, sum = 0;
    for (int j = 0; j < n_k; ++j) {
      double s = 0;
      for (int c = 0; c < D; ++c) s += Q[i * D + c] * K[j * D + c];
      P[j] = s / sqrt((double)D);
      if (P[j] > mx) mx = P[j];
    }
    for (int j = 0; j < n_k; ++j) { P[j] = exp(P[j] - mx); sum += P[j]; }
    for (int c = 0; c < D; ++c) {
      double a = 0;
      for (int j = 0; j < n_k; ++j) a += P[j] / sum * x[j * D + c];
      O[c] = a;
    }
    for (int j = 0; j < D; ++j) {
      double a = o->b[j] + (res ? res[i * D + j] : 0.0);
      for (int c = 0; c < D; ++c) a += O[c] * o->w[j * D + c];
      y[i * D + j] = (relu && a < 0) ? 0.0 : a;
    }
  }
  free(Q); free(K); free(P); free(O);
}

static int upload_split(const float* x, int n, hvr_bf16** hi, hvr_bf16** lo, float* merged) {
  uint16_t* h = (uint16_t*)malloc(2 * n);
  uint16_t* l = (uint16_t*)malloc(2 * n);
  for (int i = 0; i < n; ++i) {
    h[i] = bf16_rn(x[i]);
    l[i] = bf16_rn(x[i] - bf16_f(h[i]));
    merged[i] = bf16_f(h[i]) + bf16_f(l[i]);           /* what the device actually sees */
  }
  CU(cudaMalloc((void**)hi, 2 * n));
  CU(cudaMalloc((void**)lo, 2 * n));
  CU(cudaMemcpy(*hi, h, 2 * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(*lo, l, 2 * n, cudaMemcpyHostToDevice));
  free(h); free(l);
  return 0;
}

int main(void) {
  const int D = 256, n_k = 333, n_key = 77, key0 = 100;
  printf("hvr_abi_version %d\n", hvr_abi_version());
  Lin q, k, o;
  if (make_linear(&q, D, 0.25f) || make_linear(&k, D, 0.25f) || make_linear(&o, D, 0.06f)) return 2;
  float* x = (float*)malloc(sizeof(float) * n_k * D);
  float* xm = (float*)malloc(sizeof(float) * n_k * D);
  for (int i = 0; i < n_k * D; ++i) x[i] = frand();
  hvr_bf16 *x_hi, *x_lo, *y_hi, *y_lo;
  if (upload_split(x, n_k * D, &x_hi, &x_lo, xm)) return 2;
  CU(cudaMalloc((void**)&y_hi, 2 * n_k * D));
  CU(cudaMalloc((void**)&y_lo, 2 * n_k * D));
  HvrRelationWeights w = {D, q.hi, q.lo, q.bias, k.hi, k.lo, k.bias, o.hi, o.lo, o.bias};
  size_t wsb = hvr_relation_workspace_bytes(n_k, n_k, D);
  void* ws;
  CU(cudaMalloc(&ws, wsb));
  double* ref = (double*)malloc(sizeof(double) * n_k * D);
  uint16_t* yh = (uint16_t*)malloc(2 * n_k * D);
  uint16_t* yl = (uint16_t*)malloc(2 * n_k * D);
  double worst = 0;
  for (int mode = 0; mode < 2; ++mode) {
    /* mode 0: every row a query, residual = the rows, ReLU (stages 1 / 3); mode 1: key rows only (stages 2 / 4) */
    const int n_q = mode ? n_key : n_k;
    const float* xq = mode ? xm + key0 * D : xm;
    CK(hvr_relation_fwd(&w, x_hi, x_lo, D, n_k, mode ? x_hi + key0 * D : NULL, mode ? x_lo + key0 * D : NULL, D, n_q,
                        mode ? x_hi + key0 * D : x_hi, mode ? x_lo + key0 * D : x_lo, D, 1, y_hi, y_lo, D, ws, wsb, NULL));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(yh, y_hi, 2 * n_q * D, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(yl, y_lo, 2 * n_q * D, cudaMemcpyDeviceToHost));
    reference(xm, n_k, xq, n_q, xq, 1, D, &q, &k, &o, ref);
    double mx = 0, err = 0;
    for (int i = 0; i < n_q * D; ++i) {
      const double got = (double)bf16_f(yh[i]) + (double)bf16_f(yl[i]);
      if (fabs(ref[i]) > mx) mx = fabs(ref[i]);
      if (fabs(got - ref[i]) > err) err = fabs(got - ref[i]);
    }
    printf("mode %d: n_q %d n_k %d D %d  max|y| %.4f  max err %.3e  rel %.3e\n", mode, n_q, n_k, D, mx, err, err / mx);
    if (err / mx > worst) worst = err / mx;
  }
  /* argument checks cross the boundary as return codes, never as exits */
  if (hvr_relation_fwd(&w, x_hi, x_lo, D, n_k, NULL, NULL, D, n_k, NULL, NULL, 0, 1, y_hi, y_lo, D, ws, 16, NULL) !=
      HVR_ERR_WORKSPACE) { fprintf(stderr, "workspace check missing\n"); return 1; }
  if (worst > 1e-3) { fprintf(stderr, "relation block differs from the double-precision restatement: %.3e\n", worst); return 1; }
  printf("OK %llu kernels launched\n", (unsigned long long)hvr_launch_count());
  return 0;
}
