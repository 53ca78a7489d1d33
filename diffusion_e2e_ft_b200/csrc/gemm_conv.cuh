// tcgen05 GEMM + im2col-free implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )        fp16 operands, fp32 accumulate in TMEM
//
// * A/W tiles are staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage smem ring.
// * LINEAR mode: A is a 3-D tensor (K, rows, batch).
// * CONV mode:   A is the NHWC activation itself, a 4-D tensor (C, W, H, N).  An M-tile is a
//   bw x bh patch of output pixels of one image; k-block (tap, c-block) is fetched with the box
//   (64, bw, bh, 1) at coordinates (c0, w0*s+dx, h0*s+dy, n) — the halo / padding comes from TMA
//   out-of-bounds zero fill, the stride from the tensor map's element strides.  No im2col buffer.
//   Optional second source A2 (same pixel tiling, 1x1) appends k-blocks: this fuses the
//   ResnetBlock2D 1x1 `conv_shortcut` into conv2's accumulation.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16); accumulators are double
//   buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
// * Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
// * Persistent: grid = min(#tiles, #SMs), static round-robin tile schedule (n fastest).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;       // 64 x fp16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;
constexpr int kAccStages = 2;
constexpr int kAccStrideCols = 256;
constexpr int kMaxTaps = 9;

enum EpiAct { ACT_NONE = 0, ACT_SILU = 1, ACT_GEGLU = 2 };

struct GemmParams {
  int M, N, num_k_blocks;
  int batch, m_tiles, n_tiles;
  int a_batched, b_batched;    // operand has a batch dimension (else shared across the batch)
  // ---- conv geometry (conv != 0)
  int conv;
  int Ho, Wo;                  // conv-output grid the M tiles walk over
  int bw, bh, tiles_w, tiles_h;
  int cin_blocks, num_taps, in_stride;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps];
  int k2_blocks;               // trailing k-blocks read from A2 (1x1 shortcut)
  // output pixel mapping: pixel (ho,wo) -> (ho*out_mul+out_oy, wo*out_mul+out_ox) in OHxOW
  int out_mul, out_oy, out_ox, OH, OW;
  // ---- epilogue
  void* out;
  long long ldo, out_batch_stride;
  int out_f32;
  int vec_ok;                  // 16-byte vector stores/loads are aligned (ldo, ld_res % 8 == 0)
  int out_nchw;                // conv only: write fp32/fp16 NCHW (small Cout) instead of NHWC
  const float* bias;           // [N] (or [M] when bias_row)
  int bias_row;
  const float* rowvec;         // per-image vector, indexed [img*ld_rowvec + col]
  long long ld_rowvec;
  int rows_per_img;            // LINEAR mode: img = row / rows_per_img (0 -> unused)
  const void* residual;        // same dtype as out
  long long ld_res, res_batch_stride;
  int act;
  float alpha;                 // scale applied to the accumulator before bias
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - kBarrierBytes;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kTotalBytes = kStages * kStageBytes + kBarrierBytes + 1024;
};

template <typename OutT>
__device__ __forceinline__ void store_chunk8(OutT* dst, const float* v);
template <>
__device__ __forceinline__ void store_chunk8<float>(float* dst, const float* v) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store_chunk8<__half>(__half* dst, const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = u;
}
template <typename OutT>
__device__ __forceinline__ void load_chunk8(const OutT* src, float* v);
template <>
__device__ __forceinline__ void load_chunk8<float>(const float* src, float* v) {
  float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load_chunk8<__half>(const __half* src, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(src);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

template <int BLOCK_N, typename OutT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + kAccStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.batch * p.m_tiles * p.n_tiles;
  const uint32_t a_bytes = p.conv ? (uint32_t)(p.bw * p.bh * kBlockK * 2) : (uint32_t)S::kABytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.k2_blocks) tma_prefetch_desc(&tmA2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int tap_blocks = p.num_k_blocks - p.k2_blocks;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        int rest = tile / p.n_tiles;
        const int m_blk = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int tw = m_blk % p.tiles_w;
          const int r2 = m_blk / p.tiles_w;
          const int th = r2 % p.tiles_h;
          img = r2 / p.tiles_h;
          h0 = th * p.bh;
          w0 = tw * p.bw;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], a_bytes + (uint32_t)S::kBBytes);
          void* sa = smem_a + stage * S::kABytes;
          void* sb = smem_b + stage * S::kBBytes;
          if (p.conv) {
            if (kb < tap_blocks) {
              const int tap = kb / p.cin_blocks;
              const int cb = kb - tap * p.cin_blocks;
              tma_load_4d(&tmA, &full_bar[stage], sa, cb * kBlockK, w0 * p.in_stride + p.tap_dx[tap],
                          h0 * p.in_stride + p.tap_dy[tap], img, kEvictNormal);
            } else {
              tma_load_4d(&tmA2, &full_bar[stage], sa, (kb - tap_blocks) * kBlockK, w0, h0, img,
                          kEvictNormal);
            }
          } else {
            tma_load_3d(&tmA, &full_bar[stage], sa, kb * kBlockK, m_blk * kBlockM, p.a_batched ? b : 0,
                        kEvictNormal);
          }
          tma_load_3d(&tmB, &full_bar[stage], sb, kb * kBlockK, n_blk * BLOCK_N, p.b_batched ? b : 0,
                      kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStrideCols;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_desc_sw128(smem_u32(smem_a + stage * S::kABytes), 16, 1024);
          const uint64_t bdesc = make_desc_sw128(smem_u32(smem_b + stage * S::kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 bytes (16 fp16) along K inside the 128B swizzle row: +2 in 16-byte units
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);           // smem slot reusable once these MMAs retire
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ======================================================================= epilogue (4 warps)
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    const OutT* res = reinterpret_cast<const OutT*>(p.residual);
    constexpr int kOutCols = BLOCK_N;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_tiles;
      int rest = tile / p.n_tiles;
      const int m_blk = rest % p.m_tiles;
      const int b = rest / p.m_tiles;
      // ---- row -> output address
      bool row_ok;
      long long orow;   // linear output row index (pixel index for conv)
      long long opix = 0;
      int img = 0;
      if (p.conv) {
        const int tw = m_blk % p.tiles_w;
        const int r2 = m_blk / p.tiles_w;
        const int th = r2 % p.tiles_h;
        img = r2 / p.tiles_h;
        const int dh = row_in_tile / p.bw;
        const int dw = row_in_tile - dh * p.bw;
        const int ho = th * p.bh + dh, wo = tw * p.bw + dw;
        row_ok = (dh < p.bh) && (ho < p.Ho) && (wo < p.Wo);
        opix = (long long)(ho * p.out_mul + p.out_oy) * p.OW + (wo * p.out_mul + p.out_ox);
        orow = (long long)img * p.OH * p.OW + opix;
      } else {
        const int r = m_blk * kBlockM + row_in_tile;
        row_ok = r < p.M;
        orow = r;
        if (p.rows_per_img) img = r / p.rows_per_img;
      }
      OutT* orow_ptr = out + (long long)b * p.out_batch_stride + orow * p.ldo;
      const OutT* rrow_ptr = res ? res + (long long)b * p.res_batch_stride + orow * p.ld_res : nullptr;
      const float* rv = p.rowvec ? p.rowvec + (long long)img * p.ld_rowvec : nullptr;
      const float bias_r = (p.bias && p.bias_row && row_ok) ? p.bias[orow] : 0.f;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * kAccStrideCols + ((uint32_t)(quad * 32) << 16);

      if (p.act == ACT_GEGLU) {
        // tile columns [0,BN/2) = value, [BN/2,BN) = gate for the same output columns
        constexpr int H = kOutCols / 2;
        const int ncol0 = n_blk * H;
        const int n_out = p.N / 2;
#pragma unroll 1
        for (int c = 0; c < H; c += 16) {
          uint32_t rv_[16], rg_[16];
          tmem_ld_32x16(t_row + c, rv_);
          tmem_ld_32x16(t_row + H + c, rg_);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              const int col = ncol0 + c + j;
              if (col < n_out) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float v = __uint_as_float(rv_[j + e]) + p.bias[n_blk * kOutCols + c + j + e];
                  float g = __uint_as_float(rg_[j + e]) + p.bias[n_blk * kOutCols + H + c + j + e];
                  o[e] = v * gelu_erf_f(g);
                }
                store_chunk8<OutT>(orow_ptr + col, o);
              }
            }
          }
        }
      } else {
        const int ncol0 = n_blk * kOutCols;
#pragma unroll 1
        for (int c = 0; c < kOutCols; c += 32) {
          uint32_t r[32];
          if (kOutCols - c >= 32) {
            tmem_ld_32x32(t_row + c, r);
          } else {
            uint32_t r16[16];
            tmem_ld_32x16(t_row + c, r16);
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = r16[e];
          }
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const int col = ncol0 + c + j;
              if (c + j < kOutCols && col < p.N) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[j + e]) * p.alpha;
                if (p.bias) {
                  if (p.bias_row) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += bias_r;
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) if (col + e < p.N) o[e] += p.bias[col + e];
                  }
                }
                if (rv) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) if (col + e < p.N) o[e] += rv[col + e];
                }
                if (p.out_nchw) {
                  const long long plane = (long long)p.OH * p.OW;
                  for (int e = 0; e < 8 && col + e < p.N; ++e) {
                    float v = o[e];
                    if (p.act == ACT_SILU) v = silu_f(v);
                    out[((long long)img * p.N + col + e) * plane + opix] = (OutT)v;
                  }
                } else if (p.vec_ok && col + 8 <= p.N) {
                  if (rrow_ptr) {
                    float rr[8];
                    load_chunk8<OutT>(rrow_ptr + col, rr);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += rr[e];
                  }
                  if (p.act == ACT_SILU) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = silu_f(o[e]);
                  }
                  store_chunk8<OutT>(orow_ptr + col, o);
                } else {
                  // ragged tail (N not a multiple of 8): scalar path
                  for (int e = 0; e < 8 && col + e < p.N; ++e) {
                    float v = o[e];
                    if (rrow_ptr) v += (float)rrow_ptr[col + e];
                    if (p.act == ACT_SILU) v = silu_f(v);
                    orow_ptr[col + e] = (OutT)v;
                  }
                }
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200
