// Task losses of the E2E fine-tuning step (forward): scale-and-shift-invariant L1 (depth) and angular (normals).
// Reference: training/util/loss.py:13-47 (ScaleAndShiftInvariantLoss, compute_scale_and_shift_masked) and
// :51-67 (AngularLoss), called at training/train.py:542-556.  HBM-bound masked reductions, double atomics.
#include "common.cuh"
#include "../../include/b200_e2eft.h"

namespace b200 {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ws[b*5 + {0..4}] += (sum m p p, sum m p, sum m, sum m p y, sum m y)
__global__ void ssi_moments_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                   const uint8_t* __restrict__ mask, long long HW, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * HW;
  const float* y = tgt + (long long)b * HW;
  const uint8_t* m = mask + (long long)b * HW;
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    if (m[i]) {
      const double pv = p[i], yv = y[i];
      a00 += pv * pv; a01 += pv; a11 += 1.0; b0 += pv * yv; b1 += yv;
    }
  }
  a00 = warp_sum_d(a00); a01 = warp_sum_d(a01); a11 = warp_sum_d(a11); b0 = warp_sum_d(b0); b1 = warp_sum_d(b1);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ws[b * 5 + 0], a00); atomicAdd(&ws[b * 5 + 1], a01); atomicAdd(&ws[b * 5 + 2], a11);
    atomicAdd(&ws[b * 5 + 3], b0);  atomicAdd(&ws[b * 5 + 4], b1);
  }
}

// ws[B*5] += sum m |s p + t - y| ; ws[B*5+1] += sum m      with (s,t) the per-image least-squares fit
__global__ void ssi_l1_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                              const uint8_t* __restrict__ mask, long long HW, int B, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const double a00 = ws[b * 5 + 0], a01 = ws[b * 5 + 1], a11 = ws[b * 5 + 2], b0 = ws[b * 5 + 3], b1 = ws[b * 5 + 4];
  const double det = a00 * a11 - a01 * a01;
  float s = 0.f, t = 0.f;                                     // loss.py:41-46: only a positive determinant is solved
  if (det > 0) { s = (float)((a11 * b0 - a01 * b1) / det); t = (float)((-a01 * b0 + a00 * b1) / det); }
  const float* p = pred + (long long)b * HW;
  const float* y = tgt + (long long)b * HW;
  const uint8_t* m = mask + (long long)b * HW;
  double acc = 0, cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
    if (m[i]) { acc += fabsf(s * p[i] + t - y[i]); cnt += 1.0; }
  acc = warp_sum_d(acc); cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&ws[B * 5], acc); atomicAdd(&ws[B * 5 + 1], cnt); }
}

__global__ void angular_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                               const uint8_t* __restrict__ mask, long long HW, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * 3 * HW;
  const float* y = tgt + (long long)b * 3 * HW;
  const uint8_t* m = mask + (long long)b * HW;
  double acc = 0, cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
    if (m[i]) {
      float d = p[i] * y[i] + p[HW + i] * y[HW + i] + p[2 * HW + i] * y[2 * HW + i];
      d = fminf(fmaxf(d, -1.0f), 1.0f);
      acc += acosf(d);
      cnt += 1.0;
    }
  acc = warp_sum_d(acc); cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&ws[0], acc); atomicAdd(&ws[1], cnt); }
}

__global__ void mean_finalize_kernel(const double* __restrict__ ws, float* __restrict__ out) {
  out[0] = (float)(ws[0] / ws[1]);                            // empty mask -> nan, like torch's mean of an empty tensor
}

static dim3 loss_grid(long long HW, int B) {
  long long g = (HW + 256 * 8 - 1) / (256 * 8);
  long long cap = (long long)sm_count() * 4 / (B > 0 ? B : 1) + 1;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return dim3((unsigned)g, B);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ssi_loss(const float* pred, const float* target, const unsigned char* mask, int B,
                             long long HW, double* workspace, float* out, void* stream) {
  B200_CHECK_ARG(pred && target && mask && workspace && out && B > 0 && HW > 0, "b200_ssi_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid = loss_grid(HW, B);
  ssi_moments_kernel<<<grid, 256, 0, st>>>(pred, target, mask, HW, workspace);
  ssi_l1_kernel<<<grid, 256, 0, st>>>(pred, target, mask, HW, B, workspace);
  mean_finalize_kernel<<<1, 1, 0, st>>>(workspace + B * 5, out);
  B200_CHECK_LAUNCH("ssi_loss kernels");
  return 0;
}

extern "C" int b200_angular_loss(const float* pred, const float* target, const unsigned char* mask, int B,
                                 long long HW, double* workspace, float* out, void* stream) {
  B200_CHECK_ARG(pred && target && mask && workspace && out && B > 0 && HW > 0, "b200_angular_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  angular_kernel<<<loss_grid(HW, B), 256, 0, st>>>(pred, target, mask, HW, workspace);
  mean_finalize_kernel<<<1, 1, 0, st>>>(workspace, out);
  B200_CHECK_LAUNCH("angular_loss kernels");
  return 0;
}
