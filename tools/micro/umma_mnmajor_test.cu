// Micro-test (round 2): MN-major SWIZZLE_128B operands with MORE THAN ONE 64-element atom along MN (N = 128 for B,
// M = 128 for A): which (LBO, SBO) pair does the tcgen05 smem descriptor need when the atoms are separate
// [64 k-rows x 128 B] TMA boxes 8192 B apart?  Needed to feed GEMMs straight from row-major [K][MN] tensors
// (weight gradients dW = dY^T X, attention backward dK = dS^T Q, dV = P^T dO, dQ = dS K) without transposition kernels.
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace b200;
namespace b200 { void set_last_error(const char*, ...) {} }

__device__ __forceinline__ uint64_t desc_lbo_sbo(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return make_desc_sw128(addr, lbo, sbo);
}

// mode 0: B MN-major (N = 128, K = 16), A K-major selector      -> out[k][n] = B[k][n]   (16 x 128)
// mode 1: A MN-major (M = 128, K = 16), B K-major selector (N=16) -> out[k][m] = A[k][m] (16 x 128)
__global__ void test_kernel(int mode, uint32_t lbo, uint32_t sbo, float* out) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* SEL = smem;                 // K-major selector tile: 128 rows x 128 B (16 KB)
  uint8_t* MNT = smem + 16384;         // MN-major operand: 2 atoms x [64 k-rows][128 B] (16 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int t = threadIdx.x;
  for (int i = t; i < 128 * 64; i += blockDim.x) {               // selector[r][k] = (r == k), k < 16
    const int r = i / 64, k = i % 64;
    const int off = r * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
    *reinterpret_cast<__half*>(SEL + off) = __float2half((r == k && k < 16) ? 1.f : 0.f);
  }
  for (int i = t; i < 64 * 128; i += blockDim.x) {                // logical [k][mn] = k * 128 + mn for k < 16, else 0
    const int k = i / 128, mn = i % 128;
    const int atom = mn / 64, c = mn % 64;
    const int off = atom * 8192 + k * 128 + (((c / 8) ^ (k % 8)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__half*>(MNT + off) = __float2half(k < 16 ? (float)(k * 128 + mn) : 0.f);
  }
  fence_proxy_async_smem();
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc(&tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (t == 0) {
    if (mode == 0) {
      const uint64_t adesc = make_desc_sw128(smem_u32(SEL), 16, 1024);
      const uint64_t bdesc = desc_lbo_sbo(smem_u32(MNT), lbo, sbo);
      umma_f16(tmem, adesc, bdesc, make_idesc_f16(128, 128, 0, 1), 0);
    } else {
      const uint64_t adesc = desc_lbo_sbo(smem_u32(MNT), lbo, sbo);
      const uint64_t bdesc = make_desc_sw128(smem_u32(SEL), 16, 1024);
      umma_f16(tmem, adesc, bdesc, make_idesc_f16(128, 16, 1, 0), 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = t >> 5, lane = t & 31;
  if (mode == 0) {
    if (warp == 0) {                              // D rows 0..15 (lanes) x 128 columns
      uint32_t r[32];
      for (int c = 0; c < 128; c += 32) {
        tmem_ld_32x32(tmem + c, r);
        tmem_ld_wait();
        if (lane < 16) for (int j = 0; j < 32; ++j) out[lane * 128 + c + j] = __uint_as_float(r[j]);
      }
    }
  } else {                                        // D[m][n] n < 16: all four warps (lane quadrants), 16 columns
    uint32_t r[16];
    tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[j * 128 + warp * 32 + lane] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main() {
  float* d;
  cudaMalloc(&d, 16 * 128 * 4);
  const int smem = 16384 + 16384 + 1024;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const uint32_t cand[4][2] = {{8192, 1024}, {1024, 8192}, {16, 1024}, {8192, 128}};
  for (int mode = 0; mode < 2; ++mode)
    for (int ci = 0; ci < 4; ++ci) {
      cudaMemset(d, 0, 16 * 128 * 4);
      test_kernel<<<1, 128, smem>>>(mode, cand[ci][0], cand[ci][1], d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(16 * 128);
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, bad_lo = 0;
      for (int k = 0; k < 16; ++k)
        for (int mn = 0; mn < 128; ++mn)
          if (h[k * 128 + mn] != (float)(k * 128 + mn)) { ++bad; if (mn < 64) ++bad_lo; }
      printf("%s MN-major  LBO %5u SBO %5u : %s (%d / 2048 mismatches, %d in the first atom)  [1][0..2]=%g %g %g  [1][64..66]=%g %g %g  [9][64]=%g\n",
             mode == 0 ? "B" : "A", cand[ci][0], cand[ci][1], bad ? "WRONG" : "ok", bad, bad_lo, h[128], h[129], h[130],
             h[128 + 64], h[128 + 65], h[128 + 66], h[9 * 128 + 64]);
    }
  return 0;
}
