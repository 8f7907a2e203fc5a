/* hvr_b200.h - C ABI of the B200-native (sm_100a) HVRNet per-key-frame inference hot path.
 *
 * Drop-in boundary.  The reference (youthHan/HVRNet, an mmdetection-v1 fork) crosses into
 * native code through pybind11 functions that take at::Tensor:
 *     roi_align_cuda.forward   mmdet/ops/roi_align/src/roi_align_cuda.cpp:27-53,82-85
 *     nms_cuda.nms             mmdet/ops/nms/src/nms_cuda.cpp:1-16  (nms_kernel.cu:71-136)
 *     nms_cpu.nms              mmdet/ops/nms/src/nms_cpu.cpp:61-67
 * and through library calls (cuDNN conv / cuBLAS gemm) behind nn.Conv2d / nn.Linear /
 * torch.bmm / torch.mm (mmdet/models/utils/conv_module.py:9-13,
 * mmdet/models/bbox_heads/hrnmp_bbox_head.py:140-186,293,342).  This header is what a
 * replacement binds instead: plain pointers and sizes, no torch types.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - the caller owns every buffer (torch's caching allocator on the Python side);
 *  - `stream` is a cudaStream_t passed as void* (the caller's current stream); no entry
 *    point synchronises the device or the stream unless its comment says so;
 *  - return value: 0 = HVR_OK, negative = error (hvr_strerror); nothing exits the process
 *    (cf. the exit() in roi_align_kernel.cu:269-272);
 *  - "split" tensors are the storage format of every activation that feeds a tensor-core
 *    contraction: a pair of bf16 arrays (hi, lo) with  x ~= float(hi) + float(lo),
 *    hi = bf16_rn(x), lo = bf16_rn(x - float(hi))  (16+ mantissa bits).  Contractions are
 *    evaluated as  hi*hi + hi*lo + lo*hi  on tcgen05 with fp32 accumulation in TMEM.
 *  - feature maps are NHWC inside the library; NCHW only at the reference-facing edge.
 */
#ifndef HVR_B200_H_
#define HVR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVR_OK 0
#define HVR_ERR_ARG (-1)      /* bad argument (shape, alignment, null pointer)          */
#define HVR_ERR_CUDA (-2)     /* a CUDA runtime / driver call failed (hvr_last_cuda_error) */
#define HVR_ERR_WORKSPACE (-3) /* workspace too small                                     */
#define HVR_ERR_UNSUPPORTED (-4)

typedef uint16_t hvr_bf16;    /* raw bfloat16 bits */

const char* hvr_strerror(int code);
/* cudaError_t of the last failing CUDA call made by this library on this thread. */
int hvr_last_cuda_error(void);
/* ABI version, bumped on any signature change. */
int hvr_abi_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t hvr_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Split-bf16 storage
 * ---------------------------------------------------------------------------------- */
int hvr_split_f32(const float* x, size_t n, hvr_bf16* hi, hvr_bf16* lo, void* stream);
int hvr_merge_f32(const hvr_bf16* hi, const hvr_bf16* lo, size_t n, float* out, void* stream);
/* 2-D variant with leading dimensions (elements): x[rows, ld_in] -> hi/lo[rows, ld_out],
 * columns [cols, ld_out) are zero-filled. */
int hvr_split_f32_2d(const float* x, int rows, int cols, int ld_in, hvr_bf16* hi, hvr_bf16* lo,
                     int ld_out, void* stream);
/* Transposed copy of a split matrix: in [rows, ld_in] (cols valid, multiple of 8) -> out
 * [cols, ld_out], columns [rows, ld_out) zero-filled.  The X^T operand of the relation head's
 * P.V product (hrnmp_bbox_head.py:340-342, V = un-projected rows). */
int hvr_transpose_split(const hvr_bf16* hi, const hvr_bf16* lo, int rows, int cols, int64_t ld_in,
                        hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, void* stream);
/* NCHW fp32 (the reference's layout) -> NHWC split, and back. */
int hvr_nchw_to_nhwc_split(const float* x, int B, int C, int H, int W, hvr_bf16* hi, hvr_bf16* lo,
                           void* stream);
int hvr_nhwc_split_to_nchw(const hvr_bf16* hi, const hvr_bf16* lo, int B, int C, int H, int W,
                           float* out, void* stream);
int hvr_nchw_to_nhwc_f32(const float* x, int B, int C, int H, int W, float* out, void* stream);
int hvr_nhwc_to_nchw_f32(const float* x, int B, int C, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Implicit GEMM on tcgen05 (replaces the cuDNN / cuBLAS calls behind nn.Conv2d, nn.Linear,
 * torch.bmm, torch.mm on the path: resnet.py:222-257, res_layer.py:67-74, rpn_head.py:30-35,
 * hrnmp_bbox_head.py:283-294,342-350,827-906).
 *
 *   D[m, n] = alpha * sum_{tap, c} A[pixel(m) + off(tap), c] * Wt[n, tap*C + c]
 *   y       = D + bias[n] + residual[m, n];  y = relu ? max(y,0) : y
 *
 * A is a 4-D strided view (C, W, H, B) of a split NHWC tensor; rows m enumerate the output
 * pixels (b, y, x) of an (out_w, out_h, batch) grid and read A at (x + dx, y + dy); reads
 * outside [0,W)x[0,H) are zero (conv padding).  A plain GEMM is W = M, H = B = 1, 1 tap.
 * Wt is [N, ntaps*C] K-major split (BN already folded in).  C must be a multiple of 8
 * (16-byte rows); K tiles of 64 are zero-filled past C.
 * ---------------------------------------------------------------------------------- */
typedef struct HvrIGemm {
  /* A operand view, element strides (bf16 elements); stride of C is 1 */
  const hvr_bf16* a_hi;
  const hvr_bf16* a_lo;
  int a_c, a_w, a_h, a_b;                 /* sizes of the view                          */
  int64_t a_stride_w, a_stride_h, a_stride_b;
  /* taps */
  int ntaps;
  int tap_dx[9], tap_dy[9];
  /* output pixel grid and M tiling: tile = tile_w x tile_h pixels, tile_w*tile_h == 128 */
  int out_w, out_h, batch;
  int tile_w, tile_h;
  /* B operand: weights [n, ntaps*a_c] K-major, leading dimension ldb (elements) */
  const hvr_bf16* b_hi;
  const hvr_bf16* b_lo;
  int n;
  int64_t ldb;
  /* epilogue */
  float alpha;
  const float* bias;                      /* [n] or NULL                                */
  const hvr_bf16* res_hi;                 /* residual [rows, ld_res] split, or NULL     */
  const hvr_bf16* res_lo;
  int64_t ld_res;
  int relu;
  hvr_bf16* out_hi;                       /* [rows, ld_out] split, or NULL              */
  hvr_bf16* out_lo;
  int64_t ld_out;
  float* out_f32;                         /* [rows, ld_f32] fp32, or NULL               */
  int64_t ld_f32;
  hvr_bf16* outT_hi;                      /* transposed [n, ld_outT] split, or NULL     */
  hvr_bf16* outT_lo;
  int64_t ld_outT;
  int passes;                             /* 3 = hi*hi+hi*lo+lo*hi (default), 1 = hi*hi */
  /* Batched product (torch.bmm over videos, hrnmp_bbox_head.py:293,342): image b of the A view
   * multiplies its own matrix Wt + b * b_stride_batch (elements, multiple of 8); 0 = one shared
   * matrix.  No bias / transposed output in this mode. */
  int64_t b_stride_batch;
  /* Optional second A operand (one more K segment, 1x1 at the output pixel itself):
   *   D[m, n] += sum_c A2[pixel(m), c] * Wt[n, ntaps*a_c + c]
   * the residual branch's 1x1 downsample convolution evaluated inside the block's last
   * convolution (resnet.py:243-255: out = bn3(conv3(o)) + bn_d(conv_d(x))).  a_c must be a
   * multiple of 64; view and strides as for A. */
  const hvr_bf16* a2_hi;
  const hvr_bf16* a2_lo;
  int a2_c, a2_w, a2_h, a2_b;
  int64_t a2_stride_w, a2_stride_h, a2_stride_b;
} HvrIGemm;

/* tcgen05 / TMEM / TMA kernel.  rows = batch*out_h*out_w. */
int hvr_igemm(const HvrIGemm* g, void* stream);
/* Test hook (process-wide, not for production use): low bits = N tile width of hvr_igemm (64, 128, 256: single-CTA
 * kernels; 512 / 640: CTA-pair kernel with 256 / 128-wide tiles; 0 = heuristic), OR-ed with: 1024 per-row instead of
 * TMA epilogue, 2048 / 4096 deep epilogue always / never, bits 13-16 L2 prefetch distance (15 = off), 1 << 17 never
 * the lean variant, 1 << 18 tiles handed out by cluster launch control instead of the static round-robin.  Every
 * combination returns the same bits (tests/test_gpu_kernels.py).  HVR_DEBUG_FLAGS=<int> applies a value at first use. */
int hvr_debug_force_bn(int bn);
/* fp32 SIMT evaluation of the same descriptor (one thread per output, fmaf in k order):
 * the on-device cross-check used by the tests at sizes the CPU oracle cannot reach. */
int hvr_igemm_check(const HvrIGemm* g, void* stream);

/* Stem helpers (resnet.py:522-527): 7x7/2 pad 3 conv as im2col (K = 147 -> 192, zero padded)
 * feeding hvr_igemm, and the 3x3/2 pad 1 max-pool on split NHWC. */
int hvr_im2col_stem(const float* img_nchw, int B, int H, int W, hvr_bf16* hi, hvr_bf16* lo,
                    int out_h, int out_w, void* stream);
int hvr_maxpool3x3s2_split(const hvr_bf16* hi, const hvr_bf16* lo, int B, int H, int W, int C,
                           hvr_bf16* ohi, hvr_bf16* olo, int out_h, int out_w, void* stream);

/* ------------------------------------------------------------------------------------
 * RoIAlign forward.  Replaces roi_align_cuda.forward (roi_align_cuda.cpp:27-53) /
 * ROIAlignForward (roi_align_kernel.cu:63-118): legacy geometry (roi_end = (x2+1)*scale,
 * no half-pixel shift), sample_num^2 bilinear samples per bin, strict IEEE fp32 in the
 * reference's operation order.
 *   feat  : NHWC fp32 [n_imgs, H, W, C] (feat_nhwc=1) or NCHW fp32 (feat_nhwc=0; transposed
 *           into `ws`, which must hold n_imgs*H*W*C floats)
 *   rois  : [n_rois, 5] = (batch_idx, x1, y1, x2, y2)
 *   out   : out_layout 0 -> [n_rois, C, ph, pw] fp32 (the reference's layout)
 *           out_layout 1 -> [n_rois, ph, pw, C] fp32
 *   out_hi/out_lo (optional, out_layout 1 order, row pitch ld_split elements): split copy
 *           feeding fc_new_1 directly.
 * A roi whose batch index is outside [0, n_imgs) is clamped (the reference reads out of bounds).
 * ---------------------------------------------------------------------------------- */
int hvr_roi_align_fwd(const float* feat, int feat_nhwc, const float* rois, int n_rois, int n_imgs,
                      int C, int H, int W, int ph, int pw, float spatial_scale, int sample_num,
                      float* out, int out_layout, hvr_bf16* out_hi, hvr_bf16* out_lo,
                      int64_t ld_split, float* ws, void* stream);
/* Test hook: 1 = always the generic per-bin kernel (16 loads per output vector, as the reference),
 * 0 = heuristic (sample_num 2 + out_layout 1 run roi_align_sn2_kernel, which reuses taps held in registers);
 * 2 / 3 = resident CTAs per SM the fast RoI-per-CTA kernel below is compiled for (experiments);
 * 4 / 5 / 6 = fast path launch shape: heuristic / slab kernel whenever it applies / never the slab kernel;
 * 7 / 8 = fast RoI-per-CTA kernel with 8 / 4 channels per thread;
 * 10 .. 15 = the 8-channel kernel at C == 256: channels of a thread adjacent / lane-interleaved / 16 per thread (all with
 * the two-row-cache walk) / lane-interleaved with the per-RoI row program (13, the default) / 13 + L1 prefetch /
 * 13 built for 3 instead of 4 CTAs per SM.  Every variant of the fast path returns the same bits. */
int hvr_debug_roi_variant(int v);
/* Fast arithmetic for the pipeline (feat_nhwc = 1, out_layout = 1, sample_num = 2): the same average of
 * bilinear samples evaluated separably - every map row of the RoI is interpolated once along x with the
 * merged column weights of an output column, then combined along y - with fused multiply-adds
 * (csrc/roi_align_sep.cuh).  Same sample positions, validity and clamping rules as the strict kernels; the
 * summation order differs, so values agree with hvr_roi_align_fwd to a few ulp of the largest term (1e-5
 * relative, the agreement between the reference's own default build, which nvcc contracts into FMAs, and its
 * -fmad=false build), not bit for bit.  Other argument combinations run the strict kernels.
 * Launch shape: one CTA per RoI (C == 256, ph <= 16, pw <= 7: roi_align_sepp_kernel - thread = output column x 8
 * lane-interleaved channels walking a per-RoI row program; other C: roi_align_sep8_kernel / roi_align_sep_kernel).
 * A second shape, CTA = (frame, 16-channel slab of the whole map staged in shared memory, H*W*64 B <= 220 KB), cuts the
 * L2 -> SM traffic 9-fold but measured slower (instruction bound); it runs only under hvr_debug_roi_variant(5) and is the
 * only user of ws: hvr_roi_align_fast_workspace_bytes(n_rois, n_imgs) bytes (the RoIs bucketed by frame); ws may be NULL. */
size_t hvr_roi_align_fast_workspace_bytes(int n_rois, int n_imgs);
int hvr_roi_align_fwd_fast(const float* feat, int feat_nhwc, const float* rois, int n_rois, int n_imgs,
                           int C, int H, int W, int ph, int pw, float spatial_scale, int sample_num,
                           float* out, int out_layout, hvr_bf16* out_hi, hvr_bf16* out_lo,
                           int64_t ld_split, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * NMS.  Replaces nms_cuda.nms (nms_kernel.cu:71-136, strict `>`; strict_gt=0 gives
 * nms_cpu.cpp:55's `>=`).  Device-resident: no D2H copy, no host scan.
 *   dets    : [n, 5] (x1,y1,x2,y2,score) fp32
 *   keep    : [n] int64, receives kept ORIGINAL indices, ascending (nms_kernel.cu:132-135)
 *   n_keep  : device int32
 *   total order of the internal sort: score descending, index ascending.
 *   ws      : hvr_nms_workspace_bytes(n) bytes
 * ---------------------------------------------------------------------------------- */
size_t hvr_nms_workspace_bytes(int n);
int hvr_nms(const float* dets, int n, float iou_thr, int strict_gt, int64_t* keep, int* n_keep,
            void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * RPN proposal generation for T frames in one call.  Replaces the per-frame Python loop
 * AnchorHead.get_bboxes / RPNHead.get_bboxes_single (anchor_head.py:209-278,
 * rpn_head.py:55-104) incl. AnchorGenerator.grid_anchors (anchor_generator.py:66-83),
 * delta2bbox (transforms.py:34-111) and nms (nms_kernel.cu).
 *   cls     : [T, H*W*A] fp32 logits, order (y, x, a)   (= permute(1,2,0) of the reference)
 *   reg     : [T, H*W*A, 4] fp32 deltas, same order; ld_cls / ld_reg = row pitch in floats
 *             of one (y,x) cell (A resp. 4A when dense) so NHWC conv outputs with padded
 *             channel counts can be passed as they are
 *   base_anchors : [A,4] fp32 (rounded base anchors), stride = anchor stride
 *   img_h/img_w  : clamp shape (img_shape, unpadded)
 *   proposals    : [T, max_num, 5] (x1,y1,x2,y2,score), score-ordered; rows >= count are 0
 *   counts       : [T] int32
 *   top_idx (optional) : [T, max_num] int32 anchor index of every proposal (tests)
 * ---------------------------------------------------------------------------------- */
size_t hvr_rpn_workspace_bytes(int T, int n_anchors, int nms_pre);
int hvr_rpn_proposals(const float* cls, int64_t ld_cls, const float* reg, int64_t ld_reg, int T,
                      int H, int W, int A, const float* base_anchors, int stride, float img_h,
                      float img_w, int nms_pre, int nms_post, int max_num, float nms_thr,
                      float* proposals, int* counts, int* top_idx, void* ws, size_t ws_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * Detection post-processing for one head output.  Replaces BBoxHead.get_det_bboxes
 * (hrnmp_bbox_head.py:1009-1052, bbox_head.py:132-169) + multiclass_nms
 * (core/post_processing/bbox_nms.py:6-66): softmax over classes, class-agnostic
 * delta2bbox (stds given), optional division by scale_factor, per-class NMS (score > thr,
 * IoU > iou_thr), concatenation by class with rows in ascending roi order, top max_per_img
 * by score (score desc, position asc).
 *   rois [n,5] (batch,x1,y1,x2,y2); cls [n, n_cls] (ld_cls); reg [n,4] (ld_reg)
 *   dets [max_per_img, 5], labels [max_per_img] int64 (0-based class), n_dets device int32
 * ---------------------------------------------------------------------------------- */
size_t hvr_det_workspace_bytes(int n, int n_cls);
int hvr_det_postprocess(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                        int64_t ld_reg, int n, int n_cls, const float* stds4_host, float img_h,
                        float img_w, float scale_factor, int rescale, float score_thr,
                        float iou_thr, int max_per_img, float* dets, int64_t* labels, int* n_dets,
                        void* ws, size_t ws_bytes, void* stream);
/* G problems of n rois each (the key frames of G videos; problem g = rows [g*n, (g+1)*n) of rois /
 * cls / reg) through ONE launch per stage; per problem bit-identical to hvr_det_postprocess.
 *   dets [G, max_per_img, 5], labels [G, max_per_img], n_dets [G] */
size_t hvr_det_batched_workspace_bytes(int G, int n, int n_cls);
int hvr_det_postprocess_batched(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                                int64_t ld_reg, int G, int n, int n_cls, const float* stds4_host,
                                float img_h, float img_w, float scale_factor, int rescale,
                                float score_thr, float iou_thr, int max_per_img, float* dets,
                                int64_t* labels, int* n_dets, void* ws, size_t ws_bytes, void* stream);

/* Extended form.  n_valid (optional, device int32 [G]): only the first n_valid[g] rows of problem g are
 * proposals (a frame that yielded fewer than max_num proposals keeps its fixed row block; the rows behind
 * the count score 0 in every class and never become candidates) - results equal those of a call with
 * n = n_valid[g] rows.  roi_idx (optional, device int32 [G, max_per_img]): the row (roi) of every detection,
 * i.e. the index information bbox_nms.py:36-61 carries implicitly (parity reports). */
int hvr_det_postprocess_batched_ex(const float* rois, const float* cls, int64_t ld_cls, const float* reg,
                                   int64_t ld_reg, int G, int n, int n_cls, const float* stds4_host,
                                   float img_h, float img_w, float scale_factor, int rescale,
                                   float score_thr, float iou_thr, int max_per_img, const int* n_valid,
                                   float* dets, int64_t* labels, int* n_dets, int* roi_idx, void* ws,
                                   size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Relation-head row softmax (hrnmp_bbox_head.py:332): P = softmax(S, dim=keys), written
 * as split bf16 for the P.V contraction.  S [rows, ld_s] fp32, P [rows, ld_p].
 * Columns [cols, ld_p) of P are zero-filled.
 * ---------------------------------------------------------------------------------- */
int hvr_softmax_rows_split(const float* S, int rows, int cols, int64_t ld_s, hvr_bf16* p_hi,
                           hvr_bf16* p_lo, int64_t ld_p, void* stream);
/* The same with a key mask for ragged proposal sets kept in fixed row blocks (hnmb_rcnn.py:582-599 builds the
 * key set from the ACTUAL per-frame proposal counts): the columns are n_segs blocks of `slot` keys - the T
 * frames of the window, then the support key frames - and only the first seg_counts[problem][seg] keys of a
 * block are proposals; the others get probability exactly 0 and do not enter the row maximum or sum, so a
 * row equals the softmax over the compacted key set.  problem = row / rows_per_problem (one window per
 * problem); seg_counts is a DEVICE int32 [n_problems, n_segs] array (no host round trip, CUDA-graph safe). */
int hvr_softmax_rows_split_masked(const float* S, int rows, int cols, int64_t ld_s, hvr_bf16* p_hi,
                                  hvr_bf16* p_lo, int64_t ld_p, const int* seg_counts, int n_segs, int slot,
                                  int rows_per_problem, void* stream);

/* ------------------------------------------------------------------------------------
 * Composites for non-Python hosts (csrc/relation.cu).  The Python engine (hvrnet_b200/engine.py)
 * packs weights and sequences the launches itself; a C / C++ host binds these instead and gets the same bits.
 *
 * Weight packing (host pointers in, device buffers out; the calls synchronise `stream` - one-time setup):
 *   hvr_pack_conv_bn : conv weight [cout, cin, kh, kw] (+ frozen BatchNorm: weight / bias / running_mean /
 *       running_var [cout], eps; all four NULL = no BN) (+ conv bias [cout] or NULL) -> folded in fp64, re-laid-out
 *       K-major [cout, (r*kw + s)*cin + c], zero-padded to [hvr_packed_rows(cout), hvr_packed_cols(kh*kw*cin)] and
 *       split (resnet.py:222-257 conv + norm; conv_module.py).  bias: fp32 [hvr_packed_rows(cout)] (NULL allowed
 *       when there is neither BN nor conv bias).
 *   hvr_pack_linear  : nn.Linear weight [n, k] (+ bias or NULL) -> split [hvr_packed_rows(n), hvr_packed_cols(k)],
 *       bias fp32 [hvr_packed_rows(n)]; col_perm (optional, host int32 [k]): packed column c = source column
 *       col_perm[c] (fc_new_1 reads RoI rows in (h, w, C) order: hrnmp_bbox_head.py:827-828).
 * ---------------------------------------------------------------------------------- */
size_t hvr_packed_rows(int n);
size_t hvr_packed_cols(int k);
int hvr_pack_conv_bn(const float* w_host, const float* bn_weight_host, const float* bn_bias_host,
                     const float* bn_mean_host, const float* bn_var_host, float bn_eps,
                     const float* conv_bias_host, int cout, int cin, int kh, int kw, hvr_bf16* w_hi,
                     hvr_bf16* w_lo, float* bias, void* stream);
int hvr_pack_linear(const float* w_host, const float* bias_host, int n, int k, const int* col_perm_host,
                    hvr_bf16* w_hi, hvr_bf16* w_lo, float* bias, void* stream);

/* One layer behind one call (the descriptor filling of hvrnet_b200/engine.py in C++; same launches, same bits):
 *   hvr_linear_fwd : y = alpha * x W^T + bias (+ res) (ReLU); x [rows, k] split (pitch ld_x), W as packed by
 *       hvr_pack_linear (n valid rows); out split (pitch ld_out) and / or fp32 (pitch ld_f32); res optional.
 *       Replaces nn.Linear (hrnmp_bbox_head.py:827-906 fc_new_k / fc_cls / fc_reg, convfc_bbox_head.py:126-167).
 *   hvr_conv_fwd   : NHWC split input [batch, h, w, cin] -> NHWC output [batch, ho, wo, hvr_packed_rows(cout)] for a
 *       ksize x ksize (1 or 3) convolution with `dilation`, padding = dilation * (ksize / 2) and `stride` (1x1 only when
 *       > 1: the caffe-style bottleneck puts the stride on conv1 / downsample), weights + bias as packed by
 *       hvr_pack_conv_bn; optional residual (same shape as the output) and ReLU.  Replaces conv + norm + activation of
 *       resnet.py:222-257 / res_layer.py:67-74 / rpn_head.py:30-35 (ConvModule). */
int hvr_linear_fwd(const hvr_bf16* x_hi, const hvr_bf16* x_lo, int rows, int k, int64_t ld_x,
                   const hvr_bf16* w_hi, const hvr_bf16* w_lo, const float* bias, int n,
                   const hvr_bf16* res_hi, const hvr_bf16* res_lo, int64_t ld_res, int relu, float alpha,
                   hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, float* out_f32, int64_t ld_f32,
                   void* stream);
int hvr_conv_fwd(const hvr_bf16* x_hi, const hvr_bf16* x_lo, int batch, int h, int w, int cin,
                 const hvr_bf16* w_hi, const hvr_bf16* w_lo, const float* bias, int cout, int ksize,
                 int dilation, int stride, const hvr_bf16* res_hi, const hvr_bf16* res_lo, int relu,
                 hvr_bf16* out_hi, hvr_bf16* out_lo, float* out_f32, void* stream);

/* One relation block behind one call.  Replaces forward_single_selsa (hrnmp_bbox_head.py:216-355; SelsaBBoxHead:
 * selsa_bbox_head.py:108-201) with conv_g False / conv_z True (the configs' setting):
 *     Q = xq Wq^T + bq;  K = x Wk^T + bk;  P = softmax_keys(Q K^T / sqrt(D));  O = P x;
 *     out = [relu](res + O Wo^T + bo)
 * x  [n_k, D] split (row pitch ld_x): the key / value rows;  xq [n_q, D] (NULL = x: every row is a query);
 * res [n_q, D] or NULL;  out [n_q, D] (row pitch ld_out).  Weights as packed by hvr_pack_linear with n = k = D
 * (linear_out_k is a 1x1 conv over a [N, D, 1, 1] tensor = a linear layer).  Six hvr_igemm / softmax / transpose
 * launches on `stream`, no host synchronisation; ws: hvr_relation_workspace_bytes(n_q, n_k, D) bytes. */
typedef struct HvrRelationWeights {
  int dim;                                   /* D (1024 in both configs; multiple of 8) */
  const hvr_bf16 *q_hi, *q_lo; const float* q_bias;   /* q_data_fc_k  */
  const hvr_bf16 *k_hi, *k_lo; const float* k_bias;   /* k_data_fc_k  */
  const hvr_bf16 *o_hi, *o_lo; const float* o_bias;   /* linear_out_k */
} HvrRelationWeights;
size_t hvr_relation_workspace_bytes(int n_q, int n_k, int D);
int hvr_relation_fwd(const HvrRelationWeights* w, const hvr_bf16* x_hi, const hvr_bf16* x_lo, int64_t ld_x,
                     int n_k, const hvr_bf16* xq_hi, const hvr_bf16* xq_lo, int64_t ld_xq, int n_q,
                     const hvr_bf16* res_hi, const hvr_bf16* res_lo, int64_t ld_res, int relu,
                     hvr_bf16* out_hi, hvr_bf16* out_lo, int64_t ld_out, void* ws, size_t ws_bytes,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * Window bookkeeping on the device (csrc/window.cu).  Replaces the host-side assembly of a window in
 * HNMBRCNN.simple_test_bboxes / get_roi_feat (hnmb_rcnn.py:580-599: bbox2roi per frame, cur_range from the
 * per-frame counts, torch.cat) for V windows of T frames at once, with every frame keeping a fixed block of P
 * (= max_num) rows so that launch geometry never depends on the counts; the counts travel as device masks.
 *   props  [F, P, 5] / counts [F] : hvr_rpn_proposals outputs of the F >= V*T frames held in the C5 buffer
 *   perm   [V*T] int64 (or NULL = identity): buffer slot of window position (v, t)
 *   rois       [V, Npad, 5]  (slot, x1, y1, x2, y2) for RoIAlign, rows t*P + j; rows >= T*P are zero
 *   rois_key   [V*P, 5]      (0, x1, y1, x2, y2) of the key frames (bbox2roi([props_key]))
 *   seg_counts [V, n_segs] int32: entries [0, T) = proposal count of every window frame (the key mask of the
 *                            relation stages); entries >= T (support blocks) are set to 0 (hvr_support_index fills them)
 *   key_counts [V] int32     proposal count of the key frames (n_valid of the post-processing)
 * ---------------------------------------------------------------------------------- */
int hvr_window_rois(const float* props, const int* counts, const int64_t* perm, int V, int T, int P,
                    int key_dim, int Npad, float* rois, float* rois_key, int* seg_counts, int n_segs,
                    int* key_counts, void* stream);
/* Row gather on split matrices (16-byte vectors; cols % 8 == 0):
 *   dst[p*dst_rows_per_problem + dst_row0 + j, :cols] = src[row(p, j), :cols],  j < n_rows, p < n_problems
 *   row(p, j) = idx ? idx[p*n_rows + j] : p*src_rows_per_problem + src_row0 + j;  a negative index gives zeros.
 * Assembles the key / value sets of the relation stages (hrnmp_bbox_head.py:865-868: key rows of stage 2 put
 * back among the window's rows; :740-752: support rows appended) without host-side torch.cat. */
int hvr_gather_rows_split(const hvr_bf16* src_hi, const hvr_bf16* src_lo, int64_t ld_src, const int* idx,
                          int64_t src_rows_per_problem, int src_row0, hvr_bf16* dst_hi, hvr_bf16* dst_lo,
                          int64_t ld_dst, int n_problems, int n_rows, int64_t dst_rows_per_problem,
                          int dst_row0, int cols, void* stream);
/* Support selection -> row indices into the all-gathered pool and the support blocks of the key mask.  The pool
 * is the receive buffer of the ONE all-gather as it lies: rank r's rows start at row r*rank_stride_rows, its vpr
 * local key frames own P rows each; its per-key-frame proposal counts are int32 at
 * pool_counts[r*counts_rank_stride + local].  sel [V, S] int64 = global key-frame indices (ring order or
 * hvr_support_select; < 0 = absent):
 *   idx[v][s*P + j] = (g/vpr)*rank_stride_rows + (g%vpr)*P + j  (-1 when absent -> zero rows in the gather)
 *   seg_counts[v][T + s] = count of key frame g (0 when absent). */
int hvr_support_index(const int64_t* sel, const int* pool_counts, int64_t counts_rank_stride, int vpr,
                      int64_t rank_stride_rows, int V, int S, int P, int T, int* idx, int* seg_counts,
                      int n_segs, void* stream);

/* ------------------------------------------------------------------------------------
 * Test-time image pipeline (next row N2).  Replaces the CPU DataLoader path Resize(keep_ratio)
 * -> Normalize -> Pad(16) -> ImageToTensor (mmdet/datasets/pipelines/transforms.py:111-125,
 * 240-322; formating.py:48-56), i.e. mmcv.imrescale = cv2.resize(INTER_LINEAR) on uint8
 * (OpenCV's fixed-point arithmetic, reproduced bit for bit), (x - mean) / std, zero padding.
 *   img : uint8 HWC BGR [h, w, 3] (device);  out : fp32 CHW [3, pad_h, pad_w]
 *   new_h/new_w : resized size (mmcv rescale_size rule, computed by the caller)
 * ---------------------------------------------------------------------------------- */
int hvr_preprocess_u8(const uint8_t* img, int h, int w, int new_h, int new_w, int pad_h, int pad_w,
                      const float* mean3_host, const float* std3_host, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Video-level similarity for the inter-video stage (next row N4).  Replaces, for inference, the
 * descriptor and similarity of HNMBRCNN.get_triplet_patches (hnmb_rcnn.py:76-101), which the reference
 * evaluates with adaptive_avg_pool2d / max / torch.mm / softmax / argmax.
 *   hvr_video_descriptor : c5_nhwc fp32 [n_videos*T, HW, C] (the shared head's output) ->
 *           desc [n_videos, C] = max over a video's T frames of the per-frame spatial mean (:78-81).
 *           ws : hvr_video_descriptor_workspace_bytes(n_videos, T, C) bytes.
 *   hvr_support_select   : desc [G, C] of ALL videos; for the n_local videos g0 .. g0+n_local-1:
 *           w = softmax over j != g of (1/sqrt(C)) desc_g . desc_j (:85-88), idx [n_local, n_support]
 *           (int64) = the n_support largest, ties to the lower index, -1 when fewer other videos
 *           exist; weights (optional) [n_local, G], 0 at j == g.
 * ---------------------------------------------------------------------------------- */
size_t hvr_video_descriptor_workspace_bytes(int n_videos, int T, int C);
int hvr_video_descriptor(const float* c5_nhwc, int n_videos, int T, int HW, int C, float* desc, void* ws,
                         size_t ws_bytes, void* stream);
int hvr_support_select(const float* desc, int G, int C, int g0, int n_local, int n_support, int64_t* idx,
                       float* weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVR_B200_H_ */
