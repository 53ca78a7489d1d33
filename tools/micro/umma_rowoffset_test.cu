// Micro-test (round 2): does a SWIZZLE_128B K-major UMMA operand descriptor whose start address is offset by a whole
// number of 128-byte rows (not a multiple of the 1024-byte swizzle atom) read the rows it names, and which value of the
// descriptor's base-offset field (bits 49-51) does it need?  Needed for a halo-resident implicit-GEMM conv (all nine
// taps as shifted views of ONE smem patch).     nvcc -arch=sm_100a -I../../diffusion_e2e_ft_b200/csrc ... && ./a.out
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace b200;
namespace b200 { void set_last_error(const char*, ...) {} }

__global__ void test_kernel(int row_off, int base_off_mode, float* out /* [16][64] */) {
  extern __shared__ __align__(16) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __half* A = reinterpret_cast<__half*>(smem);                 // 128 rows x 64 k (one SW128 atom column), 16 KB
  uint8_t* B = smem + 16384;                                   // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int t = threadIdx.x;
  // A[m][k] = (m == k) for k < 16 (identity on the first 16 columns), swizzled K-major
  for (int i = t; i < 128 * 64; i += blockDim.x) {
    const int m = i / 64, k = i % 64;
    const int off = m * 128 + (((k / 8) ^ (m % 8)) * 16) + (k % 8) * 2;
    *reinterpret_cast<__half*>(smem + off) = __float2half((m == k && k < 16) ? 1.f : 0.f);
  }
  // B[row][ch] = (row % 32) * 64 + ch  (exact in fp16), swizzled with the ABSOLUTE row index (as TMA writes it)
  for (int i = t; i < 256 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2;
    *reinterpret_cast<__half*>(B + off) = __float2half((float)((r % 32) * 64 + c));
  }
  fence_proxy_async_smem();
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc(&tmem_ptr, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (t == 0) {
    const uint32_t b_addr = smem_u32(B) + row_off * 128;
    uint64_t bdesc = make_desc_sw128(b_addr, 16, 1024);
    uint64_t bo = 0;
    if (base_off_mode == 1) bo = (uint64_t)((b_addr >> 7) & 7);
    bdesc |= bo << 49;
    const uint64_t adesc = make_desc_sw128(smem_u32(A), 16, 1024);
    umma_f16(tmem, adesc, bdesc, make_idesc_f16(128, 64, 0, 0), 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (t < 32) {                                  // lanes 0..31 hold D rows 0..31; rows 0..15 = channels 0..15
    uint32_t r[32];
    for (int c = 0; c < 64; c += 32) {
      tmem_ld_32x32(tmem + c, r);
      tmem_ld_wait();
      if (t < 16) for (int j = 0; j < 32; ++j) out[t * 64 + c + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  float* d;
  cudaMalloc(&d, 16 * 64 * 4);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  const int offs[] = {0, 1, 2, 3, 7, 8, 9, 66, 67, 131};
  for (int mode = 0; mode < 2; ++mode)
    for (int oi = 0; oi < 10; ++oi) {
      const int o = offs[oi];
      cudaMemset(d, 0, 16 * 64 * 4);
      test_kernel<<<1, 128, 16384 + 32768 + 1024>>>(o, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d off %d: CUDA error %s\n", mode, o, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(16 * 64);
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;                               // expect D[ch][n] = B[o + n][ch] = ((o + n) % 32) * 64 + ch
      for (int ch = 0; ch < 16; ++ch)
        for (int n = 0; n < 64; ++n)
          if (h[ch * 64 + n] != (float)(((o + n) % 32) * 64 + ch)) ++bad;
      printf("base_offset_mode %d  row_off %3d : %s (%d / 1024 mismatches)  sample D[1][0..3] = %g %g %g %g\n", mode, o,
             bad ? "WRONG" : "ok", bad, h[64], h[65], h[66], h[67]);
    }
  return 0;
}
