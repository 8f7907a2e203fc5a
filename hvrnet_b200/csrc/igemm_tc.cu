// Implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM,
// operands staged by TMA into 128B-swizzled shared memory), with a split-bf16 (hi/lo)
// three-product evaluation that keeps ~fp32 accuracy:  A*B ~= Ah*Bh + Ah*Bl + Al*Bh.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear / torch.bmm / torch.mm on
// the reference hot path (resnet.py:222-257, res_layer.py:67-74, rpn_head.py:30-35,
// hrnmp_bbox_head.py:283-294,342-350,827-906).  See include/hvr_b200.h (HvrIGemm).
//
// Persistent CTAs (one per SM) walk 128 x BN output tiles (BN = 64 / 128 / 256), N-fastest.  6 warps:
//   warp 0    TMA producer (one lane): per K step loads A_hi, A_lo (4-D box = tile_w x tile_h
//             pixels x 64 channels at the tap's offset; out-of-range pixels arrive as zeros =
//             conv padding) and B_hi, B_lo (BN x 64) into a STAGES-deep ring
//   warp 1    TMEM allocator + MMA issuer (one lane): 4 x UMMA_K=16 per 64-wide K step, three
//             products per step, tcgen05.commit frees the stage / signals the epilogue
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> alpha, bias, residual, ReLU ->
//             split-bf16 / fp32 / transposed stores; two TMEM accumulator buffers, so the
//             epilogue of tile i overlaps the main loop of tile i+1
// Launched with programmatic stream serialization (PDL): barrier init / TMEM alloc / descriptor
// prefetch overlap the tail of the previous kernel.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
// Epilogue staging per epilogue warp: residual hi|lo and output hi|lo tiles of 32 rows x 32 bf16
// (64-byte rows, SWIZZLE_64B) = 4 x 2 KB, moved by TMA.
constexpr int EPI_TILE_BYTES = 32 * 64;
constexpr int EPI_WARP_BYTES = 4 * EPI_TILE_BYTES;
constexpr int EPI_SMEM = 4 * EPI_WARP_BYTES;

struct alignas(64) KParams {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  CUtensorMap tmA2_hi, tmA2_lo;                  // second A operand (K segment after the taps), cblocks2 > 0
  CUtensorMap tmO_hi, tmO_lo, tmR_hi, tmR_lo;   // epilogue: split output (TMA store) / residual (TMA load)
  int tma_epi;                                   // bit 0: output through TMA, bit 1: residual through TMA
  int ntaps;
  int tap_dx[9], tap_dy[9];
  int C, cblocks;
  int cblocks2;                                  // 64-wide K steps of the second A operand (0 = none)
  int tile_w, tile_h, tiles_x, tiles_y;
  int out_w, out_h, batch;
  int n, n_tiles;
  int b_batched;                                 // B operand has its own matrix per image (3rd TMA coordinate)
  int passes;
  int pf;                                        // L2 prefetch distance of the A operand in K steps (0 = off)
  int clc;                                       // pair kernel: tiles handed out by cluster launch control (grid = one cluster per tile)
  int relu;
  float alpha;
  const float* bias;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  long long ld_res;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ld_out;
  float* out_f32;
  long long ld_f32;
  __nv_bfloat16* outT_hi;
  __nv_bfloat16* outT_lo;
  long long ld_outT;
};

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 64) ? 4 : (BN == 128 ? 3 : 2);
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_SMEM + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // two accumulator buffers: epilogue(i) overlaps mainloop(i+1)
};

// Epilogue for 32 consecutive columns of one output row.
struct ResRegs {
  uint4 h[4], l[4];
  bool valid;     // registers hold the residual of this chunk (fast path), else epilogue_chunk loads it itself
};
// Issue the residual loads of one 32-column chunk (16-byte vectors along the row).  Called one
// chunk ahead so the L2/HBM latency hides behind the previous chunk's tcgen05.ld + math + stores.
__device__ __forceinline__ void prefetch_res(const KParams& p, long long row, int n_base, bool row_ok, ResRegs& r) {
  r.valid = false;
  if (!p.res_hi || !row_ok || p.n - n_base < 32 || (p.ld_res % 8) != 0) return;
  const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + row * p.ld_res + n_base);
  const uint4* rl = reinterpret_cast<const uint4*>(p.res_lo + row * p.ld_res + n_base);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.h[j] = __ldg(rh + j);
    r.l[j] = __ldg(rl + j);
  }
  r.valid = true;
}

// `stage_out` != nullptr: the split output of this row is written (swizzled) into the warp's
// shared-memory tile instead of global memory; the caller ships the tile with one TMA store.
__device__ __forceinline__ void epilogue_chunk(const KParams& p, const uint32_t (&acc)[32], long long row,
                                               int n_base, bool row_ok, const ResRegs& pre, uint8_t* stage_out,
                                               int lane) {
  const int n_left = p.n - n_base;
  if (stage_out == nullptr && (!row_ok || n_left <= 0)) return;
  float v[32];
  if (p.alpha == 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * p.alpha;
  }
  const bool full = n_left >= 32;
  if (p.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) v[j] += __ldg(p.bias + n_base + j);
    }
  }
  if (p.res_hi) {
    const __nv_bfloat16* rh = p.res_hi + row * p.ld_res + n_base;
    const __nv_bfloat16* rl = p.res_lo + row * p.ld_res + n_base;
    if (pre.valid) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 h = pre.h[j / 8];
        const uint4 l = pre.l[j / 8];
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[j + 2 * q] += __fadd_rn(bf16bits_to_f32(hw[q] & 0xFFFFu), bf16bits_to_f32(lw[q] & 0xFFFFu));
          v[j + 2 * q + 1] += __fadd_rn(bf16bits_to_f32(hw[q] >> 16), bf16bits_to_f32(lw[q] >> 16));
        }
      }
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) v[j] += merge2(rh[j], rl[j]);
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (p.out_f32 && row_ok && n_left > 0) {
    float* o = p.out_f32 + row * p.ld_f32 + n_base;
    if (full && (p.ld_f32 % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) o[j] = v[j];
    }
  }
  if (p.out_hi || p.outT_hi) {
    uint32_t hp[16], lp[16];  // packed pairs
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      // packed split: one cvt.rn.bf16x2 for the two hi parts, one for the two lo parts (same
      // round-to-nearest-even as the scalar split2)
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
      const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(__fsub_rn(v[j], __uint_as_float(hb << 16)),
                                                      __fsub_rn(v[j + 1], __uint_as_float(hb & 0xFFFF0000u)));
      hp[j / 2] = hb;
      lp[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    if (stage_out) {
      // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
      uint8_t* mh = stage_out + lane * 64;
      uint8_t* ml = mh + EPI_TILE_BYTES;
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(mh + ((j ^ sw) << 4)) = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
        *reinterpret_cast<uint4*>(ml + ((j ^ sw) << 4)) = make_uint4(lp[4 * j], lp[4 * j + 1], lp[4 * j + 2], lp[4 * j + 3]);
      }
    } else if (p.out_hi && row_ok && n_left > 0) {
      __nv_bfloat16* oh = p.out_hi + row * p.ld_out + n_base;
      __nv_bfloat16* ol = p.out_lo + row * p.ld_out + n_base;
      if (full && (p.ld_out % 8 == 0)) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<uint4*>(oh + 2 * j) = make_uint4(hp[j], hp[j + 1], hp[j + 2], hp[j + 3]);
          *reinterpret_cast<uint4*>(ol + 2 * j) = make_uint4(lp[j], lp[j + 1], lp[j + 2], lp[j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < n_left) {
            oh[j] = __ushort_as_bfloat16((unsigned short)((hp[j / 2] >> (16 * (j & 1))) & 0xFFFFu));
            ol[j] = __ushort_as_bfloat16((unsigned short)((lp[j / 2] >> (16 * (j & 1))) & 0xFFFFu));
          }
      }
    }
    if (p.outT_hi && row_ok) {
      // transposed: element (n, row); consecutive lanes hold consecutive rows -> coalesced
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_left) {
          const long long o = (long long)(n_base + j) * p.ld_outT + row;
          p.outT_hi[o] = __ushort_as_bfloat16((unsigned short)((hp[j / 2] >> (16 * (j & 1))) & 0xFFFFu));
          p.outT_lo[o] = __ushort_as_bfloat16((unsigned short)((lp[j / 2] >> (16 * (j & 1))) & 0xFFFFu));
        }
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); CTA c processes tiles c, c+grid, ...  Tile index
// runs N-fastest so the CTAs in flight share the same A (activation) tiles through L2 while
// the (small) weight matrix stays L2-resident.
template <int BN>
__global__ void __launch_bounds__(192, 1) igemm_tc_kernel(const __grid_constant__ KParams p) {
  using C_ = Cfg<BN>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment of the tile ring (SWIZZLE_128B atoms are 1024 B)
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* tiles = smem_raw + pad;
  uint8_t* epi_smem = tiles + STAGES * C_::STAGE_BYTES;          // 4 warps x [res_hi|res_lo|out_hi|out_lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_SMEM);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;            // [4] one per epilogue warp (residual TMA loads)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_k = p.ntaps * p.cblocks + p.cblocks2;
  const bool three = p.passes >= 3;
  const int n_tiles = p.n_tiles;
  const int num_tiles = p.tiles_x * p.tiles_y * p.batch * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA_hi);
    tma_prefetch_desc(&p.tmB_hi);
    if (three) {
      tma_prefetch_desc(&p.tmA_lo);
      tma_prefetch_desc(&p.tmB_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);   // one arrival per epilogue warp
    }
    for (int w = 0; w < 4; ++w) mbar_init(&res_bar[w], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, C_::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail;
  // from here on we touch global memory it may have produced.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t stage_tx = three ? (uint32_t)C_::STAGE_BYTES : (uint32_t)(A_BYTES + C_::B_BYTES);
      // L2 prefetch cursor: walks the same (tile, tap, channel block) sequence p.pf K steps ahead of
      // the loads, across tile boundaries, so activations streamed from HBM are already in L2 when
      // the ring has room for them (the ring alone holds too few bytes in flight for HBM latency)
      int q_tile = blockIdx.x, q_tap = 0, q_cb = 0, q_x0 = 0, q_y0 = 0, q_b = 0;
      auto q_decode = [&]() {
        const int m_tile = q_tile / n_tiles;
        q_x0 = (m_tile % p.tiles_x) * p.tile_w;
        q_y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.tile_h;
        q_b = m_tile / (p.tiles_x * p.tiles_y);
      };
      auto q_step = [&]() {
        if (q_tile >= num_tiles) return;
        const int ax = q_x0 + p.tap_dx[q_tap], ay = q_y0 + p.tap_dy[q_tap];
        tma_prefetch_4d(&p.tmA_hi, q_cb * BK, ax, ay, q_b);
        if (three) tma_prefetch_4d(&p.tmA_lo, q_cb * BK, ax, ay, q_b);
        if (++q_cb == p.cblocks) {
          q_cb = 0;
          if (++q_tap == p.ntaps) {
            q_tap = 0;
            q_tile += gridDim.x;
            if (q_tile < num_tiles) q_decode();
          }
        }
      };
      if (p.pf > 0) {
        if (q_tile < num_tiles) q_decode();
        for (int i = 0; i < p.pf; ++i) q_step();
      }
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        const int m_tile = tile / n_tiles;
        const int tx = m_tile % p.tiles_x;
        const int ty = (m_tile / p.tiles_x) % p.tiles_y;
        const int bimg = m_tile / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.tile_w, y0 = ty * p.tile_h, n0 = nt * BN;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int ax = x0 + p.tap_dx[tap];
          const int ay = y0 + p.tap_dy[tap];
          for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            if (p.pf > 0) q_step();
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* st = tiles + s * C_::STAGE_BYTES;
            mbar_expect_tx(&full_bar[s], stage_tx);
            const int kc = cb * BK;
            tma_load_4d(st, &p.tmA_hi, &full_bar[s], kc, ax, ay, bimg);
            tma_load_3d(st + 2 * A_BYTES, &p.tmB_hi, &full_bar[s], tap * p.C + kc, n0, p.b_batched ? bimg : 0);
            if (three) {
              tma_load_4d(st + A_BYTES, &p.tmA_lo, &full_bar[s], kc, ax, ay, bimg);
              tma_load_3d(st + 2 * A_BYTES + C_::B_BYTES, &p.tmB_lo, &full_bar[s], tap * p.C + kc, n0,
                          p.b_batched ? bimg : 0);
            }
          }
        }
        // second A operand: one more K segment at the tile's own pixels (1x1), weights columns
        // [ntaps*C, ntaps*C + C2)
        for (int cb = 0; cb < p.cblocks2; ++cb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)((it / STAGES) & 1);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* st = tiles + s * C_::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], stage_tx);
          const int kc = cb * BK, kb = p.ntaps * p.C + kc;
          tma_load_4d(st, &p.tmA2_hi, &full_bar[s], kc, x0, y0, bimg);
          tma_load_3d(st + 2 * A_BYTES, &p.tmB_hi, &full_bar[s], kb, n0, p.b_batched ? bimg : 0);
          if (three) {
            tma_load_4d(st + A_BYTES, &p.tmA2_lo, &full_bar[s], kc, x0, y0, bimg);
            tma_load_3d(st + 2 * A_BYTES + C_::B_BYTES, &p.tmB_lo, &full_bar[s], kb, n0, p.b_batched ? bimg : 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1;
      const uint32_t aph = (uint32_t)((lt >> 1) & 1);
      mbar_wait(&tmem_empty_bar[ab], aph ^ 1u);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * BN);
      for (int kk = 0; kk < total_k; ++kk, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_u32(tiles + s * C_::STAGE_BYTES);
          const uint64_t a_hi = umma_desc_k_sw128(st);
          const uint64_t a_lo = umma_desc_k_sw128(st + A_BYTES);
          const uint64_t b_hi = umma_desc_k_sw128(st + 2 * A_BYTES);
          const uint64_t b_lo = umma_desc_k_sw128(st + 2 * A_BYTES + C_::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr>>4) field
            const uint64_t ko = (uint64_t)(k * 2);
            tc_mma_f16(tmem_acc, a_hi + ko, b_hi + ko, idesc, (kk > 0 || k > 0) ? 1u : 0u);
            if (three) {
              tc_mma_f16(tmem_acc, a_hi + ko, b_lo + ko, idesc, 1u);
              tc_mma_f16(tmem_acc, a_lo + ko, b_hi + ko, idesc, 1u);
            }
          }
          tc_commit(&empty_bar[s]);                             // frees the stage when the MMAs retire
          if (kk == total_k - 1) tc_commit(&tmem_full_bar[ab]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    uint32_t res_phase = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int nt = tile % n_tiles;
      const int m_tile = tile / n_tiles;
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int bimg = m_tile / (p.tiles_x * p.tiles_y);
      const int px = tx * p.tile_w + m % p.tile_w;
      const int py = ty * p.tile_h + m / p.tile_w;
      const bool row_ok = (px < p.out_w) && (py < p.out_h);
      const long long row = ((long long)bimg * p.out_h + py) * p.out_w + px;
      const int ab = lt & 1;
      const uint32_t aph = (uint32_t)((lt >> 1) & 1);
      // TMA epilogue: this warp's 32 rows are the pixel box (bw x bh) at (ox, oy) of image bimg
      const int ox = tx * p.tile_w + (q * 32) % p.tile_w;
      const int oy = ty * p.tile_h + (q * 32) / p.tile_w;
      const bool t_out = (p.tma_epi & 1) != 0, t_res = (p.tma_epi & 2) != 0;
      uint8_t* ebuf = epi_smem + (warp - 2) * EPI_WARP_BYTES;
      uint64_t* rbar = &res_bar[warp - 2];
      if (t_res && lane == 0) {                          // residual of chunk 0: in flight during the main loop
        mbar_expect_tx(rbar, 2 * EPI_TILE_BYTES);
        tma_load_4d(ebuf, &p.tmR_hi, rbar, nt * BN, ox, oy, bimg);
        tma_load_4d(ebuf + EPI_TILE_BYTES, &p.tmR_lo, rbar, nt * BN, ox, oy, bimg);
      }
      mbar_wait(&tmem_full_bar[ab], aph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int nb = nt * BN + c0;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0), acc);
        ResRegs rr;
        rr.valid = false;
        if (t_res) {
          mbar_wait(rbar, res_phase);
          res_phase ^= 1u;
          const uint8_t* mh = ebuf + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rr.h[j] = *reinterpret_cast<const uint4*>(mh + ((j ^ sw) << 4));
            rr.l[j] = *reinterpret_cast<const uint4*>(mh + EPI_TILE_BYTES + ((j ^ sw) << 4));
          }
          rr.valid = true;
          __syncwarp();
          if (c0 + 32 < BN && lane == 0) {               // next chunk's residual lands while this one is processed
            fence_proxy_async();
            mbar_expect_tx(rbar, 2 * EPI_TILE_BYTES);
            tma_load_4d(ebuf, &p.tmR_hi, rbar, nb + 32, ox, oy, bimg);
            tma_load_4d(ebuf + EPI_TILE_BYTES, &p.tmR_lo, rbar, nb + 32, ox, oy, bimg);
          }
        } else {
          prefetch_res(p, row, nb, row_ok, rr);
        }
        tmem_ld_wait();
        if (t_out) {
          if (lane == 0) bulk_wait_read0();              // the previous chunk's store has drained the tile
          __syncwarp();
        }
        epilogue_chunk(p, acc, row, nb, row_ok, rr, t_out ? ebuf + 2 * EPI_TILE_BYTES : nullptr, lane);
        if (t_out) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.tmO_hi, ebuf + 2 * EPI_TILE_BYTES, nb, ox, oy, bimg);
            tma_store_4d(&p.tmO_lo, ebuf + 3 * EPI_TILE_BYTES, nb, ox, oy, bimg);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[ab]);
    }
    if ((p.tma_epi & 1) && lane == 0) bulk_wait0();      // all TMA stores of this warp have landed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C_::TMEM_COLS);
}

// ------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs owns a 256 x BN tile.  Each
// CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows); the leader's single MMA
// thread issues M=256 instructions that read A and B halves from both CTAs' shared memory and
// write each CTA's 128 accumulator rows into its own TMEM.  Per-SM operand ingest per K step
// drops from 32+64 KB to 32+32 KB (BN=256) - the L2->SM path (~42 B/clk/SM) is what limits
// the 3-product kernel - at unchanged tensor work.
// Barriers: full[s] lives in the leader (both CTAs' TMA loads complete_tx on it; the leader
// arms it with the bytes of both); empty[s] / tmem_full[b] are per CTA, signalled by multicast
// tcgen05.commit; tmem_empty[b] lives in the leader and counts the 8 epilogue warps of the pair.
// ------------------------------------------------------------------------------------
// M tile (128 output rows) of CTA `rank` in pair `pair_idx`.  Shared B: pairs run over the global
// list of M tiles (the odd tail pairs a real tile with an image index >= batch).  Batched B: the two
// CTAs of a pair share the B tile, so pairs never straddle images; the idle half of an image's odd
// last pair gets ty = tiles_y.  Either way the idle CTA's loads are zero-filled and its stores clipped.
struct MTile { int tx, ty, bimg; };
__device__ __forceinline__ MTile pair_m_tile(const KParams& p, int pair_idx, int rank) {
  MTile r;
  const int tpi = p.tiles_x * p.tiles_y;
  int m_in;
  if (p.b_batched) {
    const int ppi = (tpi + 1) >> 1;
    r.bimg = pair_idx / ppi;
    m_in = 2 * (pair_idx % ppi) + rank;
  } else {
    const int m_tile = 2 * pair_idx + rank;
    r.bimg = m_tile / tpi;
    m_in = m_tile % tpi;
  }
  if (m_in >= tpi) {
    r.tx = 0;
    r.ty = p.tiles_y;
  } else {
    r.tx = m_in % p.tiles_x;
    r.ty = m_in / p.tiles_x;
  }
  return r;
}

// DEEP = deep-epilogue variant for layers whose epilogue (not the main loop) sets the pace (short
// K, residual add, split output): 2 ring stages instead of 3; the shared memory that frees goes to
// 3 in-place staging buffers per TMEM lane quarter (32 rows x 64 columns, hi|lo, 128-byte rows,
// SWIZZLE_128B) and the CTA runs EIGHT epilogue warps: the two warps that may read a lane quarter
// (w, w+4) each take 32 of the 64 columns of a chunk.  The residual tile of chunk g+2 is in flight
// (TMA load) while chunk g is computed in place and chunk g-1 drains (TMA store), across tile
// boundaries.
constexpr int DEEP_NBUF = 3;
constexpr int DEEP_TILE_BYTES = 32 * 128;            // 32 rows x 64 bf16
constexpr int DEEP_BUF_BYTES = 2 * DEEP_TILE_BYTES;  // hi | lo
// MODE 0: 3-4 ring stages, per-warp residual + output staging.  MODE 1: deep epilogue (above).
// MODE 2 (lean): as 0 without the residual staging tiles, for layers that have no residual: 16 KB
// less shared memory, which leaves room for a small co-resident CTA of another stream (the RPN
// greedy-NMS branch runs UNDER the C5 convolutions instead of taking their SMs away).
template <int BN, int MODE = 0>
struct Cfg2 {
  static constexpr bool DEEP = MODE == 1;
  static constexpr bool LEAN = MODE == 2;
  static constexpr int BH_BYTES = (BN / 2) * BK * 2;              // this CTA's half of one B tile
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;  // A hi/lo + B-half hi/lo
  static constexpr int STAGES = DEEP ? 2 : ((BN == 256) ? 3 : 4);
  static constexpr int EPI_WARP = DEEP ? DEEP_NBUF * DEEP_BUF_BYTES
                                       : (LEAN ? 2 * EPI_TILE_BYTES : EPI_WARP_BYTES);   // per lane quarter
  static constexpr int EPI_BYTES = 4 * EPI_WARP;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int THREADS = DEEP ? 320 : 192;                // 2 + 8 or 2 + 4 warps
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Deep epilogue: 32 columns (half `half` of a 64-column chunk) of one row, in place in the lane
// quarter's staging tile.  row_hi / row_lo: this lane's 128-byte rows; 16-byte chunk j of row r sits
// at chunk j ^ (r & 7).  b4: the 32 bias values, loaded by the caller ahead of the barrier waits.
__device__ __forceinline__ void deep_half(const KParams& p, const uint32_t (&acc)[32], const float4 (&b4)[8],
                                          uint8_t* row_hi, uint8_t* row_lo, int sw, int half, bool has_res) {
  float v[32];
  if (p.alpha == 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * p.alpha;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[4 * j] += b4[j].x; v[4 * j + 1] += b4[j].y; v[4 * j + 2] += b4[j].z; v[4 * j + 3] += b4[j].w;
  }
  if (has_res) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int ch = ((half * 4 + j / 8) ^ sw) << 4;
      const uint4 h = *reinterpret_cast<const uint4*>(row_hi + ch);
      const uint4 l = *reinterpret_cast<const uint4*>(row_lo + ch);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[j + 2 * q] += __fadd_rn(bf16bits_to_f32(hw[q] & 0xFFFFu), bf16bits_to_f32(lw[q] & 0xFFFFu));
        v[j + 2 * q + 1] += __fadd_rn(bf16bits_to_f32(hw[q] >> 16), bf16bits_to_f32(lw[q] >> 16));
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  uint32_t hp[16], lp[16];
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(__fsub_rn(v[j], __uint_as_float(hb << 16)),
                                                    __fsub_rn(v[j + 1], __uint_as_float(hb & 0xFFFF0000u)));
    hp[j / 2] = hb;
    lp[j / 2] = *reinterpret_cast<const uint32_t*>(&l2);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = ((half * 4 + j) ^ sw) << 4;
    *reinterpret_cast<uint4*>(row_hi + ch) = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
    *reinterpret_cast<uint4*>(row_lo + ch) = make_uint4(lp[4 * j], lp[4 * j + 1], lp[4 * j + 2], lp[4 * j + 3]);
  }
}

template <int BN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MODE == 1 ? 320 : 192, 1)
    igemm_tc2_kernel(const __grid_constant__ KParams p) {
  using C_ = Cfg2<BN, MODE>;
  constexpr bool DEEP = C_::DEEP;
  constexpr bool LEAN = C_::LEAN;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* tiles = smem_raw + pad;
  uint8_t* epi_smem = tiles + STAGES * C_::STAGE_BYTES;          // 4 warps x [res_hi|res_lo|out_hi|out_lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + C_::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  constexpr int NRES = DEEP ? 4 * DEEP_NBUF : 4;     // residual TMA barriers: per lane quarter (x staging buffer)
  uint64_t* res_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + NRES);
  // Tile sequence of this cluster.  Static (p.clc == 0): cluster c of G takes tiles c, c + G, ...  Dynamic (p.clc == 1):
  // the grid has one cluster per tile; a running cluster starts with its own tile and then cancels the launch of
  // pending clusters, taking their tiles (clusterlaunchcontrol).  Response j (the tile after this cluster's j-th one,
  // j = 0, 1, ...) is requested by the leader's producer thread when it starts tile j, lands in slot j % CLC_NS of BOTH
  // CTAs, and is read by every role (producer, MMA, epilogue warps) when it finishes tile j; the slot is reused once
  // every reader of both CTAs has released it on the leader's clc_empty barrier.  A co-running kernel of another stream
  // (proposal / detection branches) that holds an SM for a while then costs that SM's share of the tiles, not a stall of
  // a fixed 1/74 of the tiles.
  constexpr int CLC_NS = 4;
  constexpr uint32_t CLC_READERS = 3 + 2 * (DEEP ? 8 : 4);     // leader: producer + MMA + epilogue warps; peer: producer + epilogue warps
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 1) + 15) & ~(uintptr_t)15);
  uint64_t* clc_full = reinterpret_cast<uint64_t*>(clc_resp + 16 * CLC_NS);
  uint64_t* clc_empty = clc_full + CLC_NS;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int total_k = p.ntaps * p.cblocks + p.cblocks2;
  const int n_tiles = p.n_tiles;
  const int m_pairs = p.b_batched ? ((p.tiles_x * p.tiles_y + 1) >> 1) * p.batch
                                  : (p.tiles_x * p.tiles_y * p.batch + 1) >> 1;
  const int num_tiles = m_pairs * n_tiles;           // pair tiles
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const bool clc = p.clc != 0;
  // tile after this cluster's j-th tile `tile` (-1: none).  Idempotent until seq_release(j).
  auto seq_get = [&](int j, int tile) -> int {
    if (!clc) return tile + num_clusters < num_tiles ? tile + num_clusters : -1;
    mbar_wait(&clc_full[j % CLC_NS], (uint32_t)((j / CLC_NS) & 1));
    const int x = clc_query(clc_resp + 16 * (j % CLC_NS));
    return x < 0 ? -1 : (x >> 1);
  };
  // one thread per reader (after every lane of the warp has read the response)
  auto seq_release = [&](int j) {
    if (clc) {
      fence_proxy_async();                           // this read before the async-proxy write of a later response
      mbar_arrive_cta(&clc_empty[j % CLC_NS], 0);
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA_hi);
    tma_prefetch_desc(&p.tmB_hi);
    tma_prefetch_desc(&p.tmA_lo);
    tma_prefetch_desc(&p.tmB_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], DEEP ? 16 : 8);   // epilogue warps x 2 CTAs (used in the leader only)
    }
    for (int w = 0; w < NRES; ++w) mbar_init(&res_bar[w], 1);
    for (int s = 0; s < CLC_NS; ++s) {
      mbar_init(&clc_full[s], 1);
      mbar_init(&clc_empty[s], CLC_READERS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_ptr_smem, C_::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // barrier inits visible to the peer before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      // L2 prefetch cursor of this CTA's A rows, p.pf K steps ahead of the loads (see igemm_tc_kernel)
      int q_tile = cluster_id, q_tap = 0, q_cb = 0, q_x0 = 0, q_y0 = 0, q_b = 0;
      auto q_decode = [&]() {
        const MTile mt = pair_m_tile(p, q_tile / n_tiles, (int)rank);
        q_x0 = mt.tx * p.tile_w;
        q_y0 = mt.ty * p.tile_h;
        q_b = mt.bimg;
      };
      auto q_step = [&]() {
        if (q_tile >= num_tiles) return;
        const int ax = q_x0 + p.tap_dx[q_tap], ay = q_y0 + p.tap_dy[q_tap];
        tma_prefetch_4d(&p.tmA_hi, q_cb * BK, ax, ay, q_b);
        tma_prefetch_4d(&p.tmA_lo, q_cb * BK, ax, ay, q_b);
        if (++q_cb == p.cblocks) {
          q_cb = 0;
          if (++q_tap == p.ntaps) {
            q_tap = 0;
            q_tile += num_clusters;
            if (q_tile < num_tiles) q_decode();
          }
        }
      };
      if (p.pf > 0) {
        if (q_tile < num_tiles) q_decode();
        for (int i = 0; i < p.pf; ++i) q_step();
      }
      int it = 0, tj = 0;
      for (int tile = cluster_id; tile >= 0; tile = seq_get(tj, tile), seq_release(tj), ++tj) {
        if (clc && leader) {                                  // ask for the tile after this one
          const int s = tj % CLC_NS;
          mbar_wait(&clc_empty[s], (uint32_t)(((tj / CLC_NS) & 1) ^ 1));
          mbar_expect_tx_cta(&clc_full[s], 16u, 0u);
          mbar_expect_tx_cta(&clc_full[s], 16u, 1u);
          clc_try_cancel(clc_resp + 16 * s, &clc_full[s]);
        }
        const int nt = tile % n_tiles;
        const MTile mt = pair_m_tile(p, tile / n_tiles, (int)rank);
        const int bimg = mt.bimg;                             // idle half of an odd pair: TMA zero-fills
        const int bb = p.b_batched ? bimg : 0;
        const int x0 = mt.tx * p.tile_w, y0 = mt.ty * p.tile_h;
        const int n0 = nt * BN + (int)rank * (BN / 2);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int ax = x0 + p.tap_dx[tap];
          const int ay = y0 + p.tap_dy[tap];
          for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            if (p.pf > 0) q_step();
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* st = tiles + s * C_::STAGE_BYTES;
            if (leader) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)C_::STAGE_BYTES);
            const int kc = cb * BK;
            tma2_load_4d(st, &p.tmA_hi, &full_bar[s], kc, ax, ay, bimg);
            tma2_load_3d(st + 2 * A_BYTES, &p.tmB_hi, &full_bar[s], tap * p.C + kc, n0, bb);
            tma2_load_4d(st + A_BYTES, &p.tmA_lo, &full_bar[s], kc, ax, ay, bimg);
            tma2_load_3d(st + 2 * A_BYTES + C_::BH_BYTES, &p.tmB_lo, &full_bar[s], tap * p.C + kc, n0, bb);
          }
        }
        for (int cb = 0; cb < p.cblocks2; ++cb, ++it) {        // second A operand (see igemm_tc_kernel)
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)((it / STAGES) & 1);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* st = tiles + s * C_::STAGE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)C_::STAGE_BYTES);
          const int kc = cb * BK, kb = p.ntaps * p.C + kc;
          tma2_load_4d(st, &p.tmA2_hi, &full_bar[s], kc, x0, y0, bimg);
          tma2_load_3d(st + 2 * A_BYTES, &p.tmB_hi, &full_bar[s], kb, n0, bb);
          tma2_load_4d(st + A_BYTES, &p.tmA2_lo, &full_bar[s], kc, x0, y0, bimg);
          tma2_load_3d(st + 2 * A_BYTES + C_::BH_BYTES, &p.tmB_lo, &full_bar[s], kb, n0, bb);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, 256);
      int it = 0, lt = 0;
      for (int tile = cluster_id; tile >= 0; ++lt) {
        const int ab = lt & 1;
        const uint32_t aph = (uint32_t)((lt >> 1) & 1);
        mbar_wait(&tmem_empty_bar[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * BN);
        for (int kk = 0; kk < total_k; ++kk, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)((it / STAGES) & 1);
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t st = smem_u32(tiles + s * C_::STAGE_BYTES);
            const uint64_t a_hi = umma_desc_k_sw128(st);
            const uint64_t a_lo = umma_desc_k_sw128(st + A_BYTES);
            const uint64_t b_hi = umma_desc_k_sw128(st + 2 * A_BYTES);
            const uint64_t b_lo = umma_desc_k_sw128(st + 2 * A_BYTES + C_::BH_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);
              tc_mma2_f16(tmem_acc, a_hi + ko, b_hi + ko, idesc, (kk > 0 || k > 0) ? 1u : 0u);
              tc_mma2_f16(tmem_acc, a_hi + ko, b_lo + ko, idesc, 1u);
              tc_mma2_f16(tmem_acc, a_lo + ko, b_hi + ko, idesc, 1u);
            }
            tc_commit2(&empty_bar[s]);
            if (kk == total_k - 1) tc_commit2(&tmem_full_bar[ab]);
          }
          __syncwarp();
        }
        tile = seq_get(lt, tile);
        __syncwarp();
        if (lane == 0) seq_release(lt);
      }
    }
  } else {
    if constexpr (DEEP) {
      // ================= deep epilogue (both CTAs, own 128 rows, 8 warps) =================
      // Chunks of 64 columns, numbered g = 0, 1, ... over all tiles of this CTA; chunk g lives in
      // staging buffer g % 3 of its lane quarter.  Steady state for chunk g: its residual tile landed
      // two chunks ago; accumulator + residual -> split output written in place by the quarter's two
      // warps (32 columns each) -> TMA store(g); then, once store(g-1) has been read out of buffer
      // (g+2) % 3, the residual of chunk g+2 is loaded into it.
      const int q = warp & 3;                     // TMEM lane quarter
      const int half = (warp - 2) >> 2;           // which 32 of a chunk's 64 columns
      const bool issuer = half == 0 && lane == 0; // the quarter's TMA thread
      const bool t_res = (p.tma_epi & 2) != 0;
      uint8_t* wbuf = epi_smem + q * C_::EPI_WARP;
      uint64_t* rbar = res_bar + q * DEEP_NBUF;
      const int sw = lane & 7;
      const int qx = (q * 32) % p.tile_w, qy = (q * 32) / p.tile_w;
      // load cursor (issuer only): runs two chunks ahead of the consumer, across tile boundaries
      int l_tile = cluster_id, l_j = 0, l_c = 0, l_g = 0, l_n0 = 0, l_x = 0, l_y = 0, l_b = 0;
      auto load_tile = [&]() {
        const int nt = l_tile % n_tiles;
        const MTile mt = pair_m_tile(p, l_tile / n_tiles, (int)rank);
        l_n0 = nt * BN;
        l_x = mt.tx * p.tile_w + qx;
        l_y = mt.ty * p.tile_h + qy;
        l_b = mt.bimg;
      };
      auto issue_load = [&]() {
        if (l_tile >= 0) {
          uint8_t* dst = wbuf + (l_g % DEEP_NBUF) * DEEP_BUF_BYTES;
          uint64_t* bar = &rbar[l_g % DEEP_NBUF];
          mbar_expect_tx(bar, DEEP_BUF_BYTES);
          tma_load_4d(dst, &p.tmR_hi, bar, l_n0 + l_c, l_x, l_y, l_b);
          tma_load_4d(dst + DEEP_TILE_BYTES, &p.tmR_lo, bar, l_n0 + l_c, l_x, l_y, l_b);
          ++l_g;
          l_c += 64;
          if (l_c >= BN || l_n0 + l_c >= p.n) {
            l_c = 0;
            l_tile = seq_get(l_j, l_tile);             // read ahead of this warp's own release of response l_j
            ++l_j;
            if (l_tile >= 0) load_tile();
          }
        }
      };
      if (t_res && issuer) {
        load_tile();
        issue_load();
        issue_load();
      }
      int g = 0, lt = 0;
      for (int tile = cluster_id; tile >= 0; ++lt) {
        const int nt = tile % n_tiles;
        const MTile mt = pair_m_tile(p, tile / n_tiles, (int)rank);
        const int tx = mt.tx, ty = mt.ty, bimg = mt.bimg;
        const int ox = tx * p.tile_w + qx;
        const int oy = ty * p.tile_h + qy;
        const int ab = lt & 1;
        const uint32_t aph = (uint32_t)((lt >> 1) & 1);
        mbar_wait(&tmem_full_bar[ab], aph);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN && nt * BN + c0 < p.n; c0 += 64, ++g) {
          const int nb = nt * BN + c0;            // first column of the chunk
          const int nh = nb + half * 32;          // first column of this warp's half
          uint8_t* buf = wbuf + (g % DEEP_NBUF) * DEEP_BUF_BYTES;
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0 + half * 32), acc);
          // bias of the 32 columns: requested before the waits so its latency hides behind them
          float4 b4[8];
          if (p.bias && p.n - nh >= 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nh) + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = nh + 4 * j;
              b4[j].x = (p.bias && c < p.n) ? __ldg(p.bias + c) : 0.f;
              b4[j].y = (p.bias && c + 1 < p.n) ? __ldg(p.bias + c + 1) : 0.f;
              b4[j].z = (p.bias && c + 2 < p.n) ? __ldg(p.bias + c + 2) : 0.f;
              b4[j].w = (p.bias && c + 3 < p.n) ? __ldg(p.bias + c + 3) : 0.f;
            }
          }
          if (t_res) mbar_wait(&rbar[g % DEEP_NBUF], (uint32_t)((g / DEEP_NBUF) & 1));
          tmem_ld_wait();
          uint8_t* row_hi = buf + lane * 128;
          deep_half(p, acc, b4, row_hi, row_hi + DEEP_TILE_BYTES, sw, half, t_res);
          fence_proxy_async();
          named_bar_sync(1 + q, 64);              // both halves of the chunk are in the staging tile
          if (issuer) {
            tma_store_4d(&p.tmO_hi, buf, nb, ox, oy, bimg);
            tma_store_4d(&p.tmO_lo, buf + DEEP_TILE_BYTES, nb, ox, oy, bimg);
            bulk_commit();
            bulk_wait_read1();                     // store(g-1) has left buffer (g+2) % 3
            if (t_res) issue_load();               // residual of chunk g+2
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[ab]);
        tile = seq_get(lt, tile);
        __syncwarp();
        if (lane == 0) seq_release(lt);
      }
      if (issuer) bulk_wait0();
    } else {
      // ================= epilogue (both CTAs, own 128 rows) =================
      const int q = warp & 3;
      const int m = q * 32 + lane;
      uint32_t res_phase = 0;
      int lt = 0;
      for (int tile = cluster_id; tile >= 0; ++lt) {
        const int nt = tile % n_tiles;
        const MTile mt = pair_m_tile(p, tile / n_tiles, (int)rank);
        const int tx = mt.tx, ty = mt.ty, bimg = mt.bimg;
        const int px = tx * p.tile_w + m % p.tile_w;
        const int py = ty * p.tile_h + m / p.tile_w;
        const bool row_ok = (px < p.out_w) && (py < p.out_h) && (bimg < p.batch);
        const long long row = ((long long)bimg * p.out_h + py) * p.out_w + px;
        const int ab = lt & 1;
        const uint32_t aph = (uint32_t)((lt >> 1) & 1);
        // TMA epilogue: this warp's 32 rows are the pixel box (bw x bh) at (ox, oy) of image bimg
        const int ox = tx * p.tile_w + (q * 32) % p.tile_w;
        const int oy = ty * p.tile_h + (q * 32) / p.tile_w;
        const bool t_out = (p.tma_epi & 1) != 0, t_res = !LEAN && (p.tma_epi & 2) != 0;
        constexpr int OUT_OFF = LEAN ? 0 : 2 * EPI_TILE_BYTES;   // output staging tiles inside the warp's buffer
        uint8_t* ebuf = epi_smem + (warp - 2) * C_::EPI_WARP;
        uint64_t* rbar = &res_bar[warp - 2];
        if (t_res && lane == 0) {                          // residual of chunk 0: in flight during the main loop
          mbar_expect_tx(rbar, 2 * EPI_TILE_BYTES);
          tma_load_4d(ebuf, &p.tmR_hi, rbar, nt * BN, ox, oy, bimg);
          tma_load_4d(ebuf + EPI_TILE_BYTES, &p.tmR_lo, rbar, nt * BN, ox, oy, bimg);
        }
        mbar_wait(&tmem_full_bar[ab], aph);
        tc_fence_after();
  #pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int nb = nt * BN + c0;
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0), acc);
          ResRegs rr;
          rr.valid = false;
          if (t_res) {
            mbar_wait(rbar, res_phase);
            res_phase ^= 1u;
            const uint8_t* mh = ebuf + lane * 64;
            const int sw = (lane >> 1) & 3;
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              rr.h[j] = *reinterpret_cast<const uint4*>(mh + ((j ^ sw) << 4));
              rr.l[j] = *reinterpret_cast<const uint4*>(mh + EPI_TILE_BYTES + ((j ^ sw) << 4));
            }
            rr.valid = true;
            __syncwarp();
            if (c0 + 32 < BN && lane == 0) {               // next chunk's residual lands while this one is processed
              fence_proxy_async();
              mbar_expect_tx(rbar, 2 * EPI_TILE_BYTES);
              tma_load_4d(ebuf, &p.tmR_hi, rbar, nb + 32, ox, oy, bimg);
              tma_load_4d(ebuf + EPI_TILE_BYTES, &p.tmR_lo, rbar, nb + 32, ox, oy, bimg);
            }
          } else {
            prefetch_res(p, row, nb, row_ok, rr);
          }
          tmem_ld_wait();
          if (t_out) {
            if (lane == 0) bulk_wait_read0();              // the previous chunk's store has drained the tile
            __syncwarp();
          }
          epilogue_chunk(p, acc, row, nb, row_ok, rr, t_out ? ebuf + OUT_OFF : nullptr, lane);
          if (t_out) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&p.tmO_hi, ebuf + OUT_OFF, nb, ox, oy, bimg);
              tma_store_4d(&p.tmO_lo, ebuf + OUT_OFF + EPI_TILE_BYTES, nb, ox, oy, bimg);
              bulk_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[ab]);
        tile = seq_get(lt, tile);
        __syncwarp();
        if (lane == 0) seq_release(lt);
      }
      if ((p.tma_epi & 1) && lane == 0) bulk_wait0();
      }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // neither CTA leaves while the other may still touch it
  if (warp == 1) tmem_dealloc2(tmem_base, C_::TMEM_COLS);
}

// ------------------------------------------------------------------------------------
// fp32 SIMT evaluation of the same descriptor (cross-check).
// ------------------------------------------------------------------------------------
__global__ void igemm_check_kernel(HvrIGemm g, long long rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * g.n) return;
  const int n = (int)(idx % g.n);
  const long long row = idx / g.n;
  const int x = (int)(row % g.out_w);
  const int y = (int)((row / g.out_w) % g.out_h);
  const int b = (int)(row / ((long long)g.out_w * g.out_h));
  const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(g.a_hi);
  const __nv_bfloat16* al = reinterpret_cast<const __nv_bfloat16*>(g.a_lo);
  const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(g.b_hi);
  const __nv_bfloat16* bl = reinterpret_cast<const __nv_bfloat16*>(g.b_lo);
  float acc = 0.f;
  for (int t = 0; t < g.ntaps; ++t) {
    const int ax = x + g.tap_dx[t], ay = y + g.tap_dy[t];
    if (ax < 0 || ax >= g.a_w || ay < 0 || ay >= g.a_h) continue;
    const long long ao = b * g.a_stride_b + ay * g.a_stride_h + ax * g.a_stride_w;
    const long long bo = (long long)b * g.b_stride_batch + (long long)n * g.ldb + (long long)t * g.a_c;
    for (int c = 0; c < g.a_c; ++c) {
      float a = __bfloat162float(ah[ao + c]);
      float w = __bfloat162float(bh[bo + c]);
      if (g.passes >= 3) {
        a = __fadd_rn(a, __bfloat162float(al[ao + c]));
        w = __fadd_rn(w, __bfloat162float(bl[bo + c]));
      }
      acc = fmaf(a, w, acc);
    }
  }
  if (g.a2_hi) {
    const __nv_bfloat16* ah2 = reinterpret_cast<const __nv_bfloat16*>(g.a2_hi);
    const __nv_bfloat16* al2 = reinterpret_cast<const __nv_bfloat16*>(g.a2_lo);
    const long long ao = b * g.a2_stride_b + y * g.a2_stride_h + x * g.a2_stride_w;
    const long long bo = (long long)b * g.b_stride_batch + (long long)n * g.ldb + (long long)g.ntaps * g.a_c;
    for (int c = 0; c < g.a2_c; ++c) {
      float a = __bfloat162float(ah2[ao + c]);
      float w = __bfloat162float(bh[bo + c]);
      if (g.passes >= 3) {
        a = __fadd_rn(a, __bfloat162float(al2[ao + c]));
        w = __fadd_rn(w, __bfloat162float(bl[bo + c]));
      }
      acc = fmaf(a, w, acc);
    }
  }
  float v = acc * g.alpha;
  if (g.bias) v += g.bias[n];
  if (g.res_hi)
    v += merge2(reinterpret_cast<const __nv_bfloat16*>(g.res_hi)[row * g.ld_res + n],
                reinterpret_cast<const __nv_bfloat16*>(g.res_lo)[row * g.ld_res + n]);
  if (g.relu) v = fmaxf(v, 0.f);
  if (g.out_f32) g.out_f32[row * g.ld_f32 + n] = v;
  __nv_bfloat16 h, l;
  split2(v, h, l);
  if (g.out_hi) {
    reinterpret_cast<__nv_bfloat16*>(g.out_hi)[row * g.ld_out + n] = h;
    reinterpret_cast<__nv_bfloat16*>(g.out_lo)[row * g.ld_out + n] = l;
  }
  if (g.outT_hi) {
    reinterpret_cast<__nv_bfloat16*>(g.outT_hi)[(long long)n * g.ld_outT + row] = h;
    reinterpret_cast<__nv_bfloat16*>(g.outT_lo)[(long long)n * g.ld_outT + row] = l;
  }
}

// ------------------------------------------------------------------------------------
// Host: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library links against cudart only), cached by geometry + pointer.
// ------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

using MapKey = std::tuple<const void*, int, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                          uint32_t, uint32_t, uint32_t, int>;
std::map<MapKey, CUtensorMap> g_map_cache;
std::mutex g_map_mutex;

int make_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, int swizzle_bytes = 128) {
  MapKey key{ptr, rank, dims[0], dims[1], rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0,
             strides_bytes[0], rank > 2 ? strides_bytes[1] : 0, rank > 3 ? strides_bytes[2] : 0,
             box[0], box[1], rank > 2 ? box[2] : 0, swizzle_bytes};
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return HVR_OK;
    }
  }
  auto enc = get_encode();
  if (!enc) return HVR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return HVR_ERR_ARG;
  for (int i = 0; i < rank - 1; ++i)
    if (strides_bytes[i] % 16 != 0) return HVR_ERR_ARG;
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_hvr_last_cuda_error = 100000 + (int)r;
    return HVR_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_map_mutex);
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache[key] = *out;
  return HVR_OK;
}

int validate(const HvrIGemm* g) {
  if (!g || !g->a_hi || !g->b_hi) return HVR_ERR_ARG;
  if (g->passes >= 3 && (!g->a_lo || !g->b_lo)) return HVR_ERR_ARG;
  if (g->ntaps < 1 || g->ntaps > 9) return HVR_ERR_ARG;
  if (g->tile_w * g->tile_h != BM || g->tile_w > 256 || g->tile_h > 256) return HVR_ERR_ARG;
  if (g->a_c < 1 || g->n < 1 || g->out_w < 1 || g->out_h < 1 || g->batch < 1) return HVR_ERR_ARG;
  if (g->a_c % 8 != 0 && g->ntaps > 1) return HVR_ERR_ARG;
  if ((g->out_hi == nullptr) != (g->out_lo == nullptr)) return HVR_ERR_ARG;
  if ((g->outT_hi == nullptr) != (g->outT_lo == nullptr)) return HVR_ERR_ARG;
  if ((g->res_hi == nullptr) != (g->res_lo == nullptr)) return HVR_ERR_ARG;
  if (!g->out_hi && !g->out_f32 && !g->outT_hi) return HVR_ERR_ARG;
  if ((g->a2_hi == nullptr) != (g->a2_lo == nullptr)) return HVR_ERR_ARG;
  if (g->a2_hi && (g->a2_c < 8 || g->a2_c % 8 != 0 || g->a_c % 64 != 0 || g->passes < 3)) return HVR_ERR_ARG;
  if (g->b_stride_batch < 0 || (g->b_stride_batch % 8) != 0) return HVR_ERR_ARG;
  if (g->b_stride_batch && (g->bias || g->outT_hi)) return HVR_ERR_ARG;   // per-image B: plain products only
  return HVR_OK;
}

int g_force_bn = 0;   // test hook (hvr_debug_force_bn): 0 = heuristic
int g_deep_mode = 0;  // test hook: 0 = heuristic, 1 = deep epilogue wherever it applies, 2 = never
bool g_lean = true;   // test hook (bit 17 of hvr_debug_force_bn's argument): false = never the lean variant
int g_pf_mode = 0;    // test hook: 0 = heuristic, 1..14 = L2 prefetch distance in K steps, 15 = off
bool g_tma_epilogue = true;   // test hook: bit 10 of hvr_debug_force_bn's argument selects the per-row epilogue
bool g_clc = false;           // test hook: bit 18 of hvr_debug_force_bn's argument: pair kernels take their tiles by cluster launch control

template <int BN>
int launch(const HvrIGemm* g, KParams& kp, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    HVR_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    attr_set = true;
  }
  // B as a 3-D tensor (k, n, image): one shared matrix (third extent 1) or one per image
  const uint64_t ktot = (uint64_t)g->ntaps * g->a_c + (uint64_t)(g->a2_hi ? g->a2_c : 0);
  const uint64_t bdims[3] = {ktot, (uint64_t)g->n, (uint64_t)(g->b_stride_batch ? g->batch : 1)};
  const uint64_t bstr[2] = {(uint64_t)g->ldb * 2, (uint64_t)(g->b_stride_batch ? g->b_stride_batch : g->ldb) * 2};
  const uint32_t bbox[3] = {BK, BN, 1};
  int rc = make_map(&kp.tmB_hi, g->b_hi, 3, bdims, bstr, bbox);
  if (rc) return rc;
  if (g->passes >= 3) {
    rc = make_map(&kp.tmB_lo, g->b_lo, 3, bdims, bstr, bbox);
    if (rc) return rc;
  }
  kp.n_tiles = hvr_cdiv(g->n, BN);
  const long long num_tiles = (long long)kp.tiles_x * kp.tiles_y * kp.batch * kp.n_tiles;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    HVR_CUDA(cudaGetDevice(&dev));
    HVR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_tiles < num_sms ? num_tiles : num_sms));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = Cfg<BN>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  HVR_CUDA(cudaLaunchKernelEx(&cfg, igemm_tc_kernel<BN>, kp));
  HVR_LAUNCHED();
  return HVR_OK;
}

template <int BN, int MODE>
int launch2(const HvrIGemm* g, KParams& kp, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    HVR_CUDA(cudaFuncSetAttribute(igemm_tc2_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg2<BN, MODE>::SMEM));
    attr_set = true;
  }
  const uint64_t ktot = (uint64_t)g->ntaps * g->a_c + (uint64_t)(g->a2_hi ? g->a2_c : 0);
  const uint64_t bdims[3] = {ktot, (uint64_t)g->n, (uint64_t)(g->b_stride_batch ? g->batch : 1)};
  const uint64_t bstr[2] = {(uint64_t)g->ldb * 2, (uint64_t)(g->b_stride_batch ? g->b_stride_batch : g->ldb) * 2};
  const uint32_t bbox[3] = {BK, BN / 2, 1};
  int rc = make_map(&kp.tmB_hi, g->b_hi, 3, bdims, bstr, bbox);
  if (rc) return rc;
  rc = make_map(&kp.tmB_lo, g->b_lo, 3, bdims, bstr, bbox);
  if (rc) return rc;
  kp.n_tiles = hvr_cdiv(g->n, BN);
  const long long tpi = (long long)kp.tiles_x * kp.tiles_y;
  const long long m_pairs = kp.b_batched ? ((tpi + 1) / 2) * kp.batch : (tpi * kp.batch + 1) / 2;
  const long long pair_tiles = m_pairs * kp.n_tiles;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    HVR_CUDA(cudaGetDevice(&dev));
    HVR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  kp.clc = g_clc && pair_tiles > num_sms / 2 ? 1 : 0;       // (a single wave has nothing to hand out)
  if (kp.clc) kp.pf = 0;                                    // the L2 prefetch cursor follows the static sequence
  const long long clusters = kp.clc ? pair_tiles : (pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(Cfg2<BN, MODE>::THREADS);
  cfg.dynamicSmemBytes = Cfg2<BN, MODE>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  HVR_CUDA(cudaLaunchKernelEx(&cfg, igemm_tc2_kernel<BN, MODE>, kp));
  HVR_LAUNCHED();
  return HVR_OK;
}

}  // namespace

extern "C" int hvr_debug_force_bn(int bn);

extern "C" int hvr_igemm(const HvrIGemm* g, void* stream) {
  int rc = validate(g);
  if (rc) return rc;
  static std::once_flag env_once;      // HVR_DEBUG_FLAGS=<int>: hvr_debug_force_bn() at first use (experiments)
  std::call_once(env_once, [] {
    if (const char* e = getenv("HVR_DEBUG_FLAGS")) hvr_debug_force_bn(atoi(e));
  });
  KParams kp;
  memset(&kp, 0, sizeof(kp));
  const uint64_t adims[4] = {(uint64_t)g->a_c, (uint64_t)g->a_w, (uint64_t)g->a_h, (uint64_t)g->a_b};
  const uint64_t astr[3] = {(uint64_t)g->a_stride_w * 2, (uint64_t)g->a_stride_h * 2, (uint64_t)g->a_stride_b * 2};
  const uint32_t abox[4] = {BK, (uint32_t)g->tile_w, (uint32_t)g->tile_h, 1};
  rc = make_map(&kp.tmA_hi, g->a_hi, 4, adims, astr, abox);
  if (rc) return rc;
  if (g->passes >= 3) {
    rc = make_map(&kp.tmA_lo, g->a_lo, 4, adims, astr, abox);
    if (rc) return rc;
  }
  kp.cblocks2 = 0;
  if (g->a2_hi) {   // second A operand: the same output pixel grid seen through its own strided view
    const uint64_t a2dims[4] = {(uint64_t)g->a2_c, (uint64_t)g->a2_w, (uint64_t)g->a2_h, (uint64_t)g->a2_b};
    const uint64_t a2str[3] = {(uint64_t)g->a2_stride_w * 2, (uint64_t)g->a2_stride_h * 2,
                               (uint64_t)g->a2_stride_b * 2};
    rc = make_map(&kp.tmA2_hi, g->a2_hi, 4, a2dims, a2str, abox);
    if (rc) return rc;
    rc = make_map(&kp.tmA2_lo, g->a2_lo, 4, a2dims, a2str, abox);
    if (rc) return rc;
    kp.cblocks2 = hvr_cdiv(g->a2_c, BK);
  }
  kp.ntaps = g->ntaps;
  for (int i = 0; i < 9; ++i) {
    kp.tap_dx[i] = g->tap_dx[i];
    kp.tap_dy[i] = g->tap_dy[i];
  }
  kp.C = g->a_c;
  kp.cblocks = hvr_cdiv(g->a_c, BK);
  kp.tile_w = g->tile_w;
  kp.tile_h = g->tile_h;
  kp.tiles_x = hvr_cdiv(g->out_w, g->tile_w);
  kp.tiles_y = hvr_cdiv(g->out_h, g->tile_h);
  kp.out_w = g->out_w;
  kp.out_h = g->out_h;
  kp.batch = g->batch;
  kp.n = g->n;
  kp.b_batched = g->b_stride_batch != 0;
  kp.passes = g->passes >= 3 ? 3 : 1;
  // L2 prefetch of the A operand: 1x1 layers stream their activations once (from HBM when the tensor
  // exceeds L2); multi-tap layers re-read 8 of 9 taps from L2 anyway
  kp.pf = (g_pf_mode > 0 && g_pf_mode < 15) ? g_pf_mode : 0;   // measured in the pipeline: no gain -> off unless forced
  kp.relu = g->relu;
  kp.alpha = g->alpha;
  kp.bias = g->bias;
  kp.res_hi = reinterpret_cast<const __nv_bfloat16*>(g->res_hi);
  kp.res_lo = reinterpret_cast<const __nv_bfloat16*>(g->res_lo);
  kp.ld_res = g->ld_res;
  kp.out_hi = reinterpret_cast<__nv_bfloat16*>(g->out_hi);
  kp.out_lo = reinterpret_cast<__nv_bfloat16*>(g->out_lo);
  kp.ld_out = g->ld_out;
  kp.out_f32 = g->out_f32;
  kp.ld_f32 = g->ld_f32;
  kp.outT_hi = reinterpret_cast<__nv_bfloat16*>(g->outT_hi);
  kp.outT_lo = reinterpret_cast<__nv_bfloat16*>(g->outT_lo);
  kp.ld_outT = g->ld_outT;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long m_tiles = (long long)kp.tiles_x * kp.tiles_y * kp.batch;
  int bn = g_force_bn;
  // CTA pairs (256-row tiles, B split across the pair) once there is a full machine of pair tiles
  const long long tpi = (long long)kp.tiles_x * kp.tiles_y;
  const long long pair_tiles = (kp.b_batched ? ((tpi + 1) / 2) * kp.batch : (m_tiles + 1) / 2) * hvr_cdiv(g->n, 256);
  const bool pair = g->passes >= 3 && g->n >= 128 && ((bn == 0 && pair_tiles >= 64) || bn == 512 || bn == 640);
  // 128-wide pair tiles when the output is only 128 columns wide (half of a 256-wide tile would be
  // MMA work on zero-filled weight rows)
  const bool pair128 = pair && (bn == 640 || (bn == 0 && g->n <= 128));
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool out_tma = g_tma_epilogue && g->out_hi && g->ld_out % 8 == 0 && al16(g->out_hi) && al16(g->out_lo);
  const bool res_tma = g_tma_epilogue && g->res_hi && g->ld_res % 8 == 0 && al16(g->res_hi) && al16(g->res_lo);
  // Deep epilogue (pair kernel, 2 ring stages): split output only, everything through TMA.  Chosen
  // where the epilogue sets the pace: residual layers, and short-K layers (the main loop of a tile
  // is shorter than a 3-stage pipeline needs anyway).
  const bool deep_ok = pair && !pair128 && out_tma && !g->out_f32 && !g->outT_hi && (!g->res_hi || res_tma);
  const int total_k = g->ntaps * kp.cblocks + kp.cblocks2;
  // Measured (scripts/epilogue_bench.py, profiles/r01h_epilogue_bench.csv, and in the pipeline with
  // cold operands): the deep variant wins for K <= 512 (trunk / C5 conv3: 1.1-1.55x) and for
  // single-wave problems, whose epilogue cannot hide behind a next tile (trunk layer3 conv1 / conv2 at
  // 7 frames: 4-6 %); with K >= 1024 and several tiles per CTA pair the third ring stage is worth more.
  int sms = 0;
  {
    int dev = 0;
    static int cached = 0;
    if (!cached) {
      HVR_CUDA(cudaGetDevice(&dev));
      HVR_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    }
    sms = cached;
  }
  const bool deep = deep_ok && (g_deep_mode == 1 || (g_deep_mode == 0 && (total_k <= 8 || pair_tiles <= sms / 2)));
  // TMA epilogue maps: the [rows, ld] output / residual seen as (C = n, W, H, B) pixel tensors; one
  // box = one epilogue warp's 32 rows x 32 columns (64-byte rows, SWIZZLE_64B) or, for the deep
  // epilogue, 32 rows x 64 columns (128-byte rows, SWIZZLE_128B); out-of-range columns / pixels are
  // clipped (store) or zero-filled (load) by the TMA unit.
  kp.tma_epi = 0;
  {
    const uint32_t bw = g->tile_w < 32 ? (uint32_t)g->tile_w : 32u;
    const uint32_t ebox[4] = {deep ? 64u : 32u, bw, 32 / bw, 1};
    const int esw = deep ? 128 : 64;
    const uint64_t edims[4] = {(uint64_t)g->n, (uint64_t)g->out_w, (uint64_t)g->out_h, (uint64_t)g->batch};
    if (out_tma) {
      const uint64_t es[3] = {(uint64_t)g->ld_out * 2, (uint64_t)g->out_w * g->ld_out * 2,
                              (uint64_t)g->out_h * g->out_w * g->ld_out * 2};
      rc = make_map(&kp.tmO_hi, g->out_hi, 4, edims, es, ebox, esw);
      if (rc) return rc;
      rc = make_map(&kp.tmO_lo, g->out_lo, 4, edims, es, ebox, esw);
      if (rc) return rc;
      kp.tma_epi |= 1;
    }
    if (res_tma) {
      const uint64_t es[3] = {(uint64_t)g->ld_res * 2, (uint64_t)g->out_w * g->ld_res * 2,
                              (uint64_t)g->out_h * g->out_w * g->ld_res * 2};
      rc = make_map(&kp.tmR_hi, g->res_hi, 4, edims, es, ebox, esw);
      if (rc) return rc;
      rc = make_map(&kp.tmR_lo, g->res_lo, 4, edims, es, ebox, esw);
      if (rc) return rc;
      kp.tma_epi |= 2;
    }
  }
  const bool lean = g->res_hi == nullptr && g_lean;       // no residual: the 16 KB leaner variant
  if (pair128) return lean ? launch2<128, 2>(g, kp, st) : launch2<128, 0>(g, kp, st);
  if (pair) return deep ? launch2<256, 1>(g, kp, st) : (lean ? launch2<256, 2>(g, kp, st) : launch2<256, 0>(g, kp, st));
  // Tile width: 256 when the problem still fills the machine (halves the A traffic per FLOP),
  // 64 for narrow outputs or when 128-wide tiles would leave most SMs idle.
  if (bn == 0) {
    if (g->n <= 64) bn = 64;
    else if (g->n % 256 == 0 && m_tiles * (g->n / 256) >= 2 * 148) bn = 256;
    else if (m_tiles * hvr_cdiv(g->n, 128) < 100 && g->n >= 128) bn = 64;
    else bn = 128;
  }
  if (bn == 64) return launch<64>(g, kp, st);
  if (bn == 256) return launch<256>(g, kp, st);
  return launch<128>(g, kp, st);
}

extern "C" int hvr_debug_force_bn(int bn) {
  g_tma_epilogue = (bn & 1024) == 0;
  g_deep_mode = (bn & 2048) ? 1 : ((bn & 4096) ? 2 : 0);   // bit 11: deep epilogue wherever it applies, bit 12: never
  g_pf_mode = (bn >> 13) & 15;                             // bits 13-16: L2 prefetch distance (15 = off, 0 = heuristic)
  g_lean = (bn & (1 << 17)) == 0;
  g_clc = (bn & (1 << 18)) != 0;
  bn &= ~(1024 | 2048 | 4096 | (15 << 13) | (1 << 17) | (1 << 18));
  // 512 = CTA-pair kernel (256-wide tiles), 640 = CTA-pair kernel with 128-wide tiles
  if (bn != 0 && bn != 64 && bn != 128 && bn != 256 && bn != 512 && bn != 640) return HVR_ERR_ARG;
  g_force_bn = bn;
  return HVR_OK;
}

extern "C" int hvr_igemm_check(const HvrIGemm* g, void* stream) {
  int rc = validate(g);
  if (rc) return rc;
  const long long rows = (long long)g->batch * g->out_h * g->out_w;
  const long long total = rows * g->n;
  igemm_check_kernel<<<hvr_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*g, rows);
  HVR_LAUNCHED();
  return HVR_OK;
}
