// Backward-pass kernels of the fine-tuning step (reference: training/train.py:545-566,
// `accelerator.backward(loss)` through the UNet and the frozen VAE decoder).  Everything here is HBM-bound
// streaming / reduction work; the GEMM-shaped halves of the backward pass (conv dgrad, conv/linear wgrad,
// attention S/dP/dQ/dK/dV products) run on the tcgen05 kernels in gemm_conv.cu with re-packed or transposed
// operands (backward_packing.py / backward.py).
//
//   gather_planar       NHWC -> [C][pixels] transpose with an optional tap shift / stride / nearest-2x source map:
//                       produces the K-major operands of the weight-gradient GEMMs (K = pixels)
//   col_sum             bias gradients
//   gn_mean_rstd        group statistics from the forward's fp64 group sums or per-channel sums
//   gn_bwd_sums/apply   GroupNorm(+SiLU) backward, two streaming passes
//   layer_norm_bwd      one warp per row, d_gamma/d_beta through shared-memory then global atomics
//   softmax_bwd_rows    dS = scale * P o (dP - rowsum(dP o P))
//   act_bwd / geglu_bwd SiLU / exact-GELU / GEGLU derivatives
#include "common.cuh"
#include "../../include/b200_e2eft.h"

namespace b200 {

// ----------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ void bw_load8(const __half* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void bw_load8(const float* p, float* v) {
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void bw_store8(__half* p, const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void bw_store8(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f(float v) { return v; }

__device__ __forceinline__ float silu_grad(float z) {
  const float s = 1.0f / (1.0f + __expf(-z));
  return s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ float gelu_fwd(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ------------------------------------------------------------------------------------------ gather_planar
// out[c][q] (row stride ldo, q = (n*Ho + o)*Wo + p) = x[n][(stride*o + oy) / up][(stride*p + ox) / up][c] (pixel stride ldx),
// zero when the source row/column is outside [0, up*H) x [0, up*W) or q >= P (q runs to Ppad).
// grid (ceil(Ppad/32), ceil(C/32)), block (32, 8); 32x32 tile through shared memory, coalesced both sides.
template <typename T>
__global__ void gather_planar_kernel(const T* __restrict__ x, long long ldx, int H, int W, int C, int Ho, int Wo, int stride,
                                     int up, int oy, int ox, long long P, long long Ppad,
                                     __half* __restrict__ out, long long ldo) {
  __shared__ float tile[32][33];
  const long long q0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const long long q = q0 + i;
    const int c = c0 + tx;
    float v = 0.f;
    if (q < P && c < C) {
      const int p = (int)(q % Wo);
      const long long t = q / Wo;
      const int o = (int)(t % Ho);
      const long long n = t / Ho;
      const int yy = stride * o + oy, xx = stride * p + ox;
      if (yy >= 0 && yy < up * H && xx >= 0 && xx < up * W) {
        const int sy = yy / up, sx = xx / up;
        v = to_f(x[((n * H + sy) * W + sx) * ldx + c]);
      }
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long q = q0 + tx;
    if (c < C && q < Ppad) out[(long long)c * ldo + q] = __float2half_rn(tile[tx][i]);
  }
}

// ------------------------------------------------------------------------------------------------ rowdot
// delta[b][h][t] = sum_{d < 64} a[b][t][h*64 + d] * c[b][t][h*64 + d]   (fp16 in, fp32 accumulate / out).
// One warp per (t, h): 64 elements = one __half2 per lane.
__global__ void rowdot_heads_kernel(const __half* __restrict__ a, long long a_bs, long long a_ls,
                                    const __half* __restrict__ c, long long c_bs, long long c_ls, int L, int heads,
                                    float* __restrict__ out) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= L) return;
  const int lane = threadIdx.x & 31;
  const __half2 x = *reinterpret_cast<const __half2*>(a + (long long)b * a_bs + (long long)t * a_ls + h * 64 + 2 * lane);
  const __half2 y = *reinterpret_cast<const __half2*>(c + (long long)b * c_bs + (long long)t * c_ls + h * 64 + 2 * lane);
  const float2 xf = __half22float2(x), yf = __half22float2(y);
  float s = fmaf(xf.x, yf.x, xf.y * yf.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[((long long)b * heads + h) * L + t] = s;
}

// ------------------------------------------------------------------------------------------------ col_sum
// out[c] += sum over rows of x[row][c];  grid (ceil(C/32), row chunks), block (32, 8).
template <typename T>
__global__ void col_sum_kernel(const T* __restrict__ x, long long rows, int C, long long ld,
                               long long rows_per_cta, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f(x[r * ld + c]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(&out[c], s);
  }
}

// ------------------------------------------------------------------------------------------ GroupNorm bwd
// mr[n][g] = (mean, rstd) from the forward's statistics: fp64 group sums, or per-channel fp32 sums of the one
// or two (channel-concatenated) inputs written by the producing kernels' epilogues.
__global__ void gn_mean_rstd_kernel(const double* __restrict__ sums, const double* __restrict__ cs1, int C1,
                                    const double* __restrict__ cs2, int C2, int NB, int HW, int groups, float eps,
                                    float* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB * groups) return;
  const int n = i / groups, g = i % groups;
  const int C = C1 + C2, cg = C / groups;
  double su = 0.0, sq = 0.0;
  if (sums) {
    su = sums[2 * (long long)i];
    sq = sums[2 * (long long)i + 1];
  } else {
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const double* src = c < C1 ? cs1 + ((long long)n * C1 + c) * 2 : cs2 + ((long long)n * C2 + (c - C1)) * 2;
      su += (double)src[0];
      sq += (double)src[1];
    }
  }
  const double cnt = (double)HW * cg;
  const double mean = su / cnt;
  double var = sq / cnt - mean * mean;
  if (var < 0) var = 0;
  mr[2 * (long long)i] = (float)mean;
  mr[2 * (long long)i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// Pass 1: S[n][c_off + c] += (sum dz, sum dz * xhat) over the pixels, dz = dy * silu'(gamma*xhat + beta).
// x: [NB][HW][Cx] (one of the concatenated inputs, channels c_off.. of the normalised tensor);
// dy: fp16 [NB][HW][Ctot].  grid (chunks, NB), block V*rpb with V = Cx/8 (same thread map as the forward).
template <typename T>
__global__ void gn_bwd_sums_kernel(const T* __restrict__ x, int Cx, int c_off, int Ctot,
                                   const __half* __restrict__ dy, int HW, int groups, int pix_per_cta,
                                   const float* __restrict__ mr, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int silu, float* __restrict__ S) {
  extern __shared__ float sm[];   // [2][Cx]
  const int V = Cx / 8;
  const int n = blockIdx.y;
  const int rpb = blockDim.x / V;
  const int v = threadIdx.x % V;
  const int r = threadIdx.x / V;
  for (int i = threadIdx.x; i < 2 * Cx; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int cpg = Ctot / groups;
  const int c0 = v * 8;
  if (r < rpb) {
    float a[8], b[8], rs[8], ms[8], s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c_off + c0 + e;
      const int g = c / cpg;
      const float mean = mr[((long long)n * groups + g) * 2], rstd = mr[((long long)n * groups + g) * 2 + 1];
      rs[e] = rstd;
      ms[e] = -mean * rstd;
      a[e] = rstd * gamma[c];
      b[e] = beta[c] - mean * a[e];
      s1[e] = s2[e] = 0.f;
    }
    const T* xb = x + (long long)n * HW * Cx + c0;
    const __half* db = dy + (long long)n * HW * Ctot + c_off + c0;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    for (int p = p0 + r; p < p1; p += rpb) {
      float f[8], d[8];
      bw_load8(xb + (long long)p * Cx, f);
      bw_load8(db + (long long)p * Ctot, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = silu ? d[e] * silu_grad(f[e] * a[e] + b[e]) : d[e];
        s1[e] += dz;
        s2[e] += dz * (f[e] * rs[e] + ms[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sm[c0 + e], s1[e]);
      atomicAdd(&sm[Cx + c0 + e], s2[e]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cx; c += blockDim.x) {
    atomicAdd(&S[((long long)n * Ctot + c_off + c) * 2 + 0], sm[c]);
    atomicAdd(&S[((long long)n * Ctot + c_off + c) * 2 + 1], sm[Cx + c]);
  }
}

// Pass 2: dx = rstd * (dz*gamma - A_g - xhat * B_g) (+ add), A_g = mean_g(gamma * dz), B_g = mean_g(gamma * dz * xhat)
// from the complete S of pass 1 (all concatenated inputs accumulated).
template <typename T, typename TO>
__global__ void gn_bwd_apply_kernel(const T* __restrict__ x, int Cx, int c_off, int Ctot,
                                    const __half* __restrict__ dy, int HW, int groups, int pix_per_cta,
                                    const float* __restrict__ mr, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int silu, const float* __restrict__ S,
                                    const TO* add, TO* dx) {
  extern __shared__ float sm[];   // gA[groups], gB[groups]
  const int V = Cx / 8;
  const int n = blockIdx.y;
  const int cpg = Ctot / groups;
  const float inv_m = 1.0f / ((float)HW * (float)cpg);
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float A = 0.f, B = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      A += gamma[c] * S[((long long)n * Ctot + c) * 2 + 0];
      B += gamma[c] * S[((long long)n * Ctot + c) * 2 + 1];
    }
    sm[g] = A * inv_m;
    sm[groups + g] = B * inv_m;
  }
  __syncthreads();
  const int rpb = blockDim.x / V;
  const int v = threadIdx.x % V;
  const int r = threadIdx.x / V;
  if (r >= rpb) return;
  const int c0 = v * 8;
  float a[8], b[8], rs[8], ms[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c_off + c0 + e;
    const int g = c / cpg;
    const float mean = mr[((long long)n * groups + g) * 2], rstd = mr[((long long)n * groups + g) * 2 + 1];
    rs[e] = rstd;
    ms[e] = -mean * rstd;
    a[e] = rstd * gamma[c];
    b[e] = beta[c] - mean * a[e];
    k0[e] = rstd * gamma[c];
    k1[e] = rstd * sm[g];
    k2[e] = rstd * sm[groups + g];
  }
  const T* xb = x + (long long)n * HW * Cx + c0;
  const __half* db = dy + (long long)n * HW * Ctot + c_off + c0;
  TO* ob = dx + (long long)n * HW * Cx + c0;
  const TO* ab = add ? add + (long long)n * HW * Cx + c0 : nullptr;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  for (int p = p0 + r; p < p1; p += rpb) {
    float f[8], d[8], o[8];
    bw_load8(xb + (long long)p * Cx, f);
    bw_load8(db + (long long)p * Ctot, d);
    if (ab) bw_load8(ab + (long long)p * Cx, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = silu ? d[e] * silu_grad(f[e] * a[e] + b[e]) : d[e];
      const float g = dz * k0[e] - k1[e] - (f[e] * rs[e] + ms[e]) * k2[e];
      o[e] = ab ? o[e] + g : g;
    }
    bw_store8(ob + (long long)p * Cx, o);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd
// one warp per row (rows strided over the grid); C <= 2048, C % 8 == 0.
template <typename T, typename TO>
__global__ void layer_norm_bwd_kernel(const T* __restrict__ x, long long rows, int C,
                                      const float* __restrict__ gamma, const __half* __restrict__ dy, float eps,
                                      const TO* add, TO* dx,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];   // dgamma[C], dbeta[C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int V = C / 8;
  constexpr int kMaxV = 8;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    float f[kMaxV][8];
    float s = 0.f;
    const T* xr = x + row * C;
    const __half* dr = dy + row * C;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < V) {
        bw_load8(xr + v * 8, f[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += f[i][e];
      }
    }
    const float mean = bw_warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < V) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = f[i][e] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(bw_warp_sum(q) / C + eps);
    // xhat in place; m1 = mean(dy*gamma), m2 = mean(dy*gamma*xhat)
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < V) {
        float d[8], g[8];
        bw_load8(dr + v * 8, d);
        bw_load8(gamma + v * 8, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[i][e] = (f[i][e] - mean) * rstd;
          const float t = d[e] * g[e];
          m1 += t;
          m2 += t * f[i][e];
        }
      }
    }
    m1 = bw_warp_sum(m1) / C;
    m2 = bw_warp_sum(m2) / C;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < V) {
        float d[8], g[8], o[8];
        bw_load8(dr + v * 8, d);
        bw_load8(gamma + v * 8, g);
        if (add) bw_load8(add + row * C + v * 8, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float gx = rstd * (d[e] * g[e] - m1 - f[i][e] * m2);
          o[e] = add ? o[e] + gx : gx;
          atomicAdd(&sm[v * 8 + e], d[e] * f[i][e]);
          atomicAdd(&sm[C + v * 8 + e], d[e]);
        }
        bw_store8(dx + row * C + v * 8, o);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], sm[i]);
    atomicAdd(&dbeta[i], sm[C + i]);
  }
}

// ------------------------------------------------------------------------------------------- softmax bwd
// one CTA per row: dS[j] = scale * P[j] * (dP[j] - sum_k dP[k] P[k]);  P, dS fp16 (row stride ldp), dP fp32 (ldd).
__global__ void softmax_bwd_rows_kernel(const __half* __restrict__ P, long long ldp, const float* __restrict__ dP,
                                        long long ldd, __half* __restrict__ dS, int cols, float scale) {
  __shared__ float red[32];
  const __half* p = P + (long long)blockIdx.x * ldp;
  const float* d = dP + (long long)blockIdx.x * ldd;
  __half* o = dS + (long long)blockIdx.x * ldp;
  const int tid = threadIdx.x, nw = blockDim.x >> 5;
  float dot = 0.f;
  for (int c = tid; c < cols; c += blockDim.x) dot += __half2float(p[c]) * d[c];
  dot = bw_warp_sum(dot);
  if ((tid & 31) == 0) red[tid >> 5] = dot;
  __syncthreads();
  dot = 0.f;
  for (int i = 0; i < nw; ++i) dot += red[i];
  for (int c = tid; c < cols; c += blockDim.x) o[c] = __float2half_rn(scale * __half2float(p[c]) * (d[c] - dot));
}

// -------------------------------------------------------------------------------------------- activations
// mode 1 SiLU, 3 exact GELU: dx = dy * act'(x)
__global__ void act_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ dy, long long n, int mode,
                               __half* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = __half2float(x[i]), d = __half2float(dy[i]);
    dx[i] = __float2half_rn(d * (mode == 1 ? silu_grad(xv) : gelu_grad(xv)));
  }
}
// GEGLU y = h * gelu(g): dh = dy * gelu(g), dg = dy * h * gelu'(g); h/g/dh/dg rows may be strided (two halves of
// one [rows][2*inner] projection), dy is [rows][inner] contiguous.
__global__ void geglu_bwd_kernel(const __half* __restrict__ h, const __half* __restrict__ g, long long ldhg,
                                 const __half* __restrict__ dy, long long rows, int inner,
                                 __half* __restrict__ dh, __half* __restrict__ dg, long long ldd) {
  const long long n = rows * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / inner;
    const int c = (int)(i % inner);
    const float hv = __half2float(h[r * ldhg + c]), gv = __half2float(g[r * ldhg + c]), d = __half2float(dy[i]);
    dh[r * ldd + c] = __float2half_rn(d * gelu_fwd(gv));
    dg[r * ldd + c] = __float2half_rn(d * hv * gelu_grad(gv));
  }
}


// ------------------------------------------------------------------------------------------- loss backward
// Scale-and-shift-invariant L1 (training/util/loss.py:13-47) differentiated THROUGH the per-image least-squares
// (s, t), as torch.autograd does in the reference.  ws (zeroed double [7*B]): [5b..5b+4] = moments
// (sum m p p, sum m p, sum m, sum m p y, sum m y), [5B+2b..] = (sum m sgn(r), sum m sgn(r) p), r = s p + t - y.
__device__ __forceinline__ double bw_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void ssi_fit(const double* ws, int b, double& s, double& t, double& det) {
  const double a00 = ws[b * 5 + 0], a01 = ws[b * 5 + 1], a11 = ws[b * 5 + 2], b0 = ws[b * 5 + 3], b1 = ws[b * 5 + 4];
  det = a00 * a11 - a01 * a01;
  s = 0.0; t = 0.0;
  if (det > 0) { s = (a11 * b0 - a01 * b1) / det; t = (-a01 * b0 + a00 * b1) / det; }
}
__global__ void ssi_bwd_moments_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
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
  a00 = bw_warp_sum_d(a00); a01 = bw_warp_sum_d(a01); a11 = bw_warp_sum_d(a11); b0 = bw_warp_sum_d(b0); b1 = bw_warp_sum_d(b1);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ws[b * 5 + 0], a00); atomicAdd(&ws[b * 5 + 1], a01); atomicAdd(&ws[b * 5 + 2], a11);
    atomicAdd(&ws[b * 5 + 3], b0);  atomicAdd(&ws[b * 5 + 4], b1);
  }
}
__global__ void ssi_bwd_sign_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                         const uint8_t* __restrict__ mask, long long HW, int B, double* __restrict__ ws) {
  const int b = blockIdx.y;
  double s, t, det;
  ssi_fit(ws, b, s, t, det);
  const float sf = (float)s, tf = (float)t;                   // the forward evaluates the residual in fp32
  const float* p = pred + (long long)b * HW;
  const float* y = tgt + (long long)b * HW;
  const uint8_t* m = mask + (long long)b * HW;
  double g0 = 0, g1 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
    if (m[i]) {
      const float r = sf * p[i] + tf - y[i];
      const double sg = (r > 0.f) - (r < 0.f);
      g0 += sg; g1 += sg * (double)p[i];
    }
  g0 = bw_warp_sum_d(g0); g1 = bw_warp_sum_d(g1);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&ws[5 * B + 2 * b], g0); atomicAdd(&ws[5 * B + 2 * b + 1], g1); }
}
__global__ void ssi_bwd_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                    const uint8_t* __restrict__ mask, long long HW, int B, const double* __restrict__ ws,
                                    const float* __restrict__ gscale, float* __restrict__ dpred) {
  const int b = blockIdx.y;
  double s, t, det;
  ssi_fit(ws, b, s, t, det);
  double N = 0;
  for (int i = 0; i < B; ++i) N += ws[i * 5 + 2];
  const double a01 = ws[b * 5 + 1], a11 = ws[b * 5 + 2], b0 = ws[b * 5 + 3], b1 = ws[b * 5 + 4];
  const double G0 = ws[5 * B + 2 * b] / N, G1 = ws[5 * B + 2 * b + 1] / N;
  const double up = (double)gscale[0];
  const float sf = (float)s, tf = (float)t;
  const float* p = pred + (long long)b * HW;
  const float* y = tgt + (long long)b * HW;
  const uint8_t* m = mask + (long long)b * HW;
  float* o = dpred + (long long)b * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    double g = 0.0;
    if (m[i] && N > 0) {
      const double pv = p[i], yv = y[i];
      const float r = sf * p[i] + tf - y[i];
      g = ((r > 0.f) - (r < 0.f)) * s / N;
      if (det > 0) {
        const double ddet = 2.0 * pv * a11 - 2.0 * a01;
        const double dns = a11 * yv - b1;
        const double dnt = -b0 - a01 * yv + 2.0 * pv * b1;
        g += (G1 * (dns - s * ddet) + G0 * (dnt - t * ddet)) / det;
      }
    }
    o[i] = (float)(up * g);
  }
}

// AngularLoss (loss.py:51-67): mean over the mask of acos(clamp(<p, y>, -1, 1)); ws (zeroed double [1]) = count.
__global__ void mask_count_kernel(const uint8_t* __restrict__ mask, long long n, double* __restrict__ ws) {
  double c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += mask[i] ? 1.0 : 0.0;
  c = bw_warp_sum_d(c);
  if ((threadIdx.x & 31) == 0) atomicAdd(ws, c);
}
__global__ void angular_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                   const uint8_t* __restrict__ mask, long long HW, const double* __restrict__ ws,
                                   const float* __restrict__ gscale, float* __restrict__ dpred) {
  const int b = blockIdx.y;
  const float* p = pred + (long long)b * 3 * HW;
  const float* y = tgt + (long long)b * 3 * HW;
  const uint8_t* m = mask + (long long)b * HW;
  float* o = dpred + (long long)b * 3 * HW;
  const float k = (float)((double)gscale[0] / ws[0]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (m[i]) {
      const float d = p[i] * y[i] + p[HW + i] * y[HW + i] + p[2 * HW + i] * y[2 * HW + i];
      if (d > -1.0f && d < 1.0f) g = -k * rsqrtf(1.0f - d * d);
    }
    o[i] = g * y[i];
    o[HW + i] = g * y[HW + i];
    o[2 * HW + i] = g * y[2 * HW + i];
  }
}

// decode_post training modes (train.py:532-540) backward.  mode 2: est = clamp(mean_c x, -1, 1);
// mode 3: est_c = clamp(x_c / (|x| + 1e-5), -1, 1).
__global__ void decode_post_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, long long HW,
                                       int mode, float* __restrict__ dx) {
  const int n = blockIdx.y;
  const float* xb = x + (long long)n * 3 * HW;
  float* ob = dx + (long long)n * 3 * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const float a = xb[p], b = xb[HW + p], c = xb[2 * HW + p];
    if (mode == 2) {
      const float m = (a + b + c) / 3.0f;
      const float g = (m >= -1.0f && m <= 1.0f) ? dout[(long long)n * HW + p] / 3.0f : 0.f;
      ob[p] = g; ob[HW + p] = g; ob[2 * HW + p] = g;
    } else {
      const float* db = dout + (long long)n * 3 * HW;
      const float nrm = sqrtf(a * a + b * b + c * c);
      const float inv = 1.0f / (nrm + 1e-5f);
      const float u0 = a * inv, u1 = b * inv, u2 = c * inv;
      const float g0 = (u0 >= -1.f && u0 <= 1.f) ? db[p] : 0.f;
      const float g1 = (u1 >= -1.f && u1 <= 1.f) ? db[HW + p] : 0.f;
      const float g2 = (u2 >= -1.f && u2 <= 1.f) ? db[2 * HW + p] : 0.f;
      // u = x / (|x| + eps):  du_c/dx_k = delta_ck * inv - x_c x_k * inv^2 / |x|
      const float dot = g0 * a + g1 * b + g2 * c;
      const float k = nrm > 0.f ? dot * inv * inv / nrm : 0.f;
      ob[p] = g0 * inv - a * k;
      ob[HW + p] = g1 * inv - b * k;
      ob[2 * HW + p] = g2 * inv - c * k;
    }
  }
}

// ------------------------------------------------------------------------------------ nearest-upsample backward
// dx[n,h,w,:] (+ add) = sum of dy[n,oh,ow,:] over the output pixels whose torch-nearest source
// (floor(o * in / out), the map of upsample_nearest_kernel) is (h, w).  fp32 NHWC, C % 4 == 0.
__global__ void upsample_nearest_bwd_kernel(const float* __restrict__ dy, int NB, int H, int W, int C, int OH, int OW,
                                            const float* add, float* dx) {
  const int V = C / 4;
  const long long total = (long long)NB * H * W * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    const long long pix = i / V;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    const int oh0 = (int)(((long long)h * OH + H - 1) / H), ow0 = (int)(((long long)w * OW + W - 1) / W);
    float4 acc = add ? reinterpret_cast<const float4*>(add + pix * C)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oh = oh0; oh < OH && (int)(((long long)oh * H) / OH) == h; ++oh)
      for (int ow = ow0; ow < OW && (int)(((long long)ow * W) / OW) == w; ++ow) {
        const float4 d = reinterpret_cast<const float4*>(dy + (((long long)n * OH + oh) * OW + ow) * C)[v];
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
    reinterpret_cast<float4*>(dx + pix * C)[v] = acc;
  }
}

}  // namespace b200

using namespace b200;

static int bw_gn_block(int C) {
  const int V = C / 8;
  if (V > 1024) return -1;
  int rpb = 256 / V;
  if (rpb < 1) rpb = 1;
  return V * rpb;
}
static int bw_gn_ppc(int NB, int HW, int rows_per_pass) {
  int target = (sm_count() * 8 + NB - 1) / NB;
  int ppc = (HW + target - 1) / target;
  if (ppc < rows_per_pass * 4) ppc = rows_per_pass * 4;
  return ppc;
}
static unsigned bw_grid1d(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

extern "C" int b200_gather_planar(const void* x, int in_f32, long long ldx, int NB, int H, int W, int C, int Ho, int Wo,
                                  int stride, int up, int oy, int ox, void* out, long long ldo, void* stream) {
  B200_CHECK_ARG(x && out && NB > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "b200_gather_planar: bad arguments");
  B200_CHECK_ARG(stride >= 1 && (up == 1 || up == 2) && ldx >= C, "b200_gather_planar: stride=%d up=%d ldx=%lld", stride, up, ldx);
  const long long P = (long long)NB * Ho * Wo;
  B200_CHECK_ARG(ldo >= P, "b200_gather_planar: ldo=%lld < pixels=%lld", ldo, P);
  const long long Ppad = ldo;   // the whole row is written (zeros past P)
  dim3 grid((unsigned)((Ppad + 31) / 32), (unsigned)((C + 31) / 32));
  B200_CHECK_ARG(grid.y <= 65535, "b200_gather_planar: C=%d too large", C);
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32)
    gather_planar_kernel<float><<<grid, block, 0, st>>>((const float*)x, ldx, H, W, C, Ho, Wo, stride, up, oy, ox, P, Ppad,
                                                        (__half*)out, ldo);
  else
    gather_planar_kernel<__half><<<grid, block, 0, st>>>((const __half*)x, ldx, H, W, C, Ho, Wo, stride, up, oy, ox, P, Ppad,
                                                         (__half*)out, ldo);
  B200_CHECK_LAUNCH("gather_planar_kernel");
  return 0;
}

extern "C" int b200_rowdot_heads(const void* a, long long a_bs, long long a_ls, const void* c, long long c_bs,
                                 long long c_ls, int B, int L, int heads, float* out, void* stream) {
  B200_CHECK_ARG(a && c && out && B > 0 && L > 0 && heads > 0, "b200_rowdot_heads: bad arguments");
  B200_CHECK_ARG(a_ls % 2 == 0 && c_ls % 2 == 0 && a_bs % 2 == 0 && c_bs % 2 == 0 &&
                     (((uintptr_t)a | (uintptr_t)c) & 3) == 0, "b200_rowdot_heads: 4-byte alignment");
  dim3 grid((L + 7) / 8, heads, B);
  rowdot_heads_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)a, a_bs, a_ls, (const __half*)c, c_bs, c_ls, L,
                                                              heads, out);
  B200_CHECK_LAUNCH("rowdot_heads_kernel");
  return 0;
}

extern "C" int b200_col_sum(const void* x, int in_f32, long long rows, int C, long long ld, float* out, void* stream) {
  B200_CHECK_ARG(x && out && rows > 0 && C > 0 && ld >= C, "b200_col_sum: bad arguments");
  const int cblocks = (C + 31) / 32;
  long long chunks = ((long long)sm_count() * 8 + cblocks - 1) / cblocks;
  if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  const long long rpc = (rows + chunks - 1) / chunks;
  dim3 grid(cblocks, (unsigned)((rows + rpc - 1) / rpc));
  dim3 block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32)
    col_sum_kernel<float><<<grid, block, 0, st>>>((const float*)x, rows, C, ld, rpc, out);
  else
    col_sum_kernel<__half><<<grid, block, 0, st>>>((const __half*)x, rows, C, ld, rpc, out);
  B200_CHECK_LAUNCH("col_sum_kernel");
  return 0;
}

extern "C" int b200_group_norm_mean_rstd(const double* sums, const double* cs1, int C1, const double* cs2, int C2,
                                         int NB, int HW, int groups, float eps, float* mean_rstd, void* stream) {
  B200_CHECK_ARG(mean_rstd && NB > 0 && HW > 0 && groups > 0 && C1 > 0, "b200_group_norm_mean_rstd: bad arguments");
  B200_CHECK_ARG(sums || cs1, "b200_group_norm_mean_rstd: need group sums or per-channel sums");
  B200_CHECK_ARG((C2 == 0) || sums || cs2, "b200_group_norm_mean_rstd: cs2 missing");
  B200_CHECK_ARG((C1 + C2) % groups == 0, "b200_group_norm_mean_rstd: C=%d groups=%d", C1 + C2, groups);
  const int n = NB * groups;
  gn_mean_rstd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, cs1, C1, cs2, C2, NB, HW, groups, eps,
                                                                         mean_rstd);
  B200_CHECK_LAUNCH("gn_mean_rstd_kernel");
  return 0;
}

static int gn_bwd_check(const char* fn, const void* x, int Cx, int c_off, int Ctot, const void* dy, int NB, int HW,
                        int groups) {
  B200_CHECK_ARG(x && dy && NB > 0 && HW > 0 && Cx > 0, "%s: bad arguments", fn);
  B200_CHECK_ARG(Cx % 8 == 0 && c_off % 8 == 0 && Ctot % 8 == 0 && c_off + Cx <= Ctot,
                 "%s: Cx=%d c_off=%d Ctot=%d must be multiples of 8 with c_off+Cx <= Ctot", fn, Cx, c_off, Ctot);
  B200_CHECK_ARG(groups > 0 && Ctot % groups == 0, "%s: Ctot=%d groups=%d", fn, Ctot, groups);
  B200_CHECK_ARG(bw_gn_block(Cx) > 0 && bw_gn_block(Cx) <= 1024, "%s: Cx=%d unsupported", fn, Cx);
  return 0;
}

extern "C" int b200_group_norm_bwd_sums(const void* x, int in_f32, int Cx, int c_off, int Ctot, const void* dy,
                                        int NB, int HW, int groups, const float* mean_rstd, const float* gamma,
                                        const float* beta, int silu, float* S, void* stream) {
  int r = gn_bwd_check("b200_group_norm_bwd_sums", x, Cx, c_off, Ctot, dy, NB, HW, groups);
  if (r) return r;
  B200_CHECK_ARG(mean_rstd && gamma && beta && S, "b200_group_norm_bwd_sums: null pointer");
  const int T = bw_gn_block(Cx);
  const int ppc = bw_gn_ppc(NB, HW, T / (Cx / 8));
  dim3 grid((HW + ppc - 1) / ppc, NB);
  const size_t smem = 2 * Cx * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32)
    gn_bwd_sums_kernel<float><<<grid, T, smem, st>>>((const float*)x, Cx, c_off, Ctot, (const __half*)dy, HW, groups, ppc,
                                                     mean_rstd, gamma, beta, silu, S);
  else
    gn_bwd_sums_kernel<__half><<<grid, T, smem, st>>>((const __half*)x, Cx, c_off, Ctot, (const __half*)dy, HW, groups,
                                                      ppc, mean_rstd, gamma, beta, silu, S);
  B200_CHECK_LAUNCH("gn_bwd_sums_kernel");
  return 0;
}

extern "C" int b200_group_norm_bwd_apply(const void* x, int in_f32, int Cx, int c_off, int Ctot, const void* dy,
                                         int NB, int HW, int groups, const float* mean_rstd, const float* gamma,
                                         const float* beta, int silu, const float* S, const void* add, void* dx,
                                         int out_f32, void* stream) {
  int r = gn_bwd_check("b200_group_norm_bwd_apply", x, Cx, c_off, Ctot, dy, NB, HW, groups);
  if (r) return r;
  B200_CHECK_ARG(mean_rstd && gamma && beta && S && dx, "b200_group_norm_bwd_apply: null pointer");
  const int T = bw_gn_block(Cx);
  const int ppc = bw_gn_ppc(NB, HW, T / (Cx / 8));
  dim3 grid((HW + ppc - 1) / ppc, NB);
  const size_t smem = 2 * groups * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define B200_GN_BWD(T_, TO_)                                                                                          \
  gn_bwd_apply_kernel<T_, TO_><<<grid, T, smem, st>>>((const T_*)x, Cx, c_off, Ctot, (const __half*)dy, HW, groups,   \
                                                      ppc, mean_rstd, gamma, beta, silu, S, (const TO_*)add, (TO_*)dx)
  if (in_f32 && out_f32) B200_GN_BWD(float, float);
  else if (in_f32) B200_GN_BWD(float, __half);
  else if (out_f32) B200_GN_BWD(__half, float);
  else B200_GN_BWD(__half, __half);
#undef B200_GN_BWD
  B200_CHECK_LAUNCH("gn_bwd_apply_kernel");
  return 0;
}

extern "C" int b200_layer_norm_bwd(const void* x, int in_f32, long long rows, int C, const float* gamma,
                                   const void* dy, float eps, const void* add, void* dx, int out_f32,
                                   float* dgamma, float* dbeta, void* stream) {
  B200_CHECK_ARG(x && dy && dx && gamma && dgamma && dbeta && rows > 0, "b200_layer_norm_bwd: bad arguments");
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "b200_layer_norm_bwd: C=%d must be a multiple of 8 and <= 2048", C);
  const int wpb = 8;
  long long grid = (rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  const size_t smem = 2 * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define B200_LN_BWD(T_, TO_)                                                                                        \
  layer_norm_bwd_kernel<T_, TO_><<<(unsigned)grid, wpb * 32, smem, st>>>((const T_*)x, rows, C, gamma,              \
                                                                         (const __half*)dy, eps, (const TO_*)add,  \
                                                                         (TO_*)dx, dgamma, dbeta)
  if (in_f32 && out_f32) B200_LN_BWD(float, float);
  else if (in_f32) B200_LN_BWD(float, __half);
  else if (out_f32) B200_LN_BWD(__half, float);
  else B200_LN_BWD(__half, __half);
#undef B200_LN_BWD
  B200_CHECK_LAUNCH("layer_norm_bwd_kernel");
  return 0;
}

extern "C" int b200_softmax_bwd_rows(const void* P, long long ldp, const float* dP, long long ldd, void* dS,
                                     long long rows, int cols, float scale, void* stream) {
  B200_CHECK_ARG(P && dP && dS && rows > 0 && cols > 0 && ldp >= cols && ldd >= cols, "b200_softmax_bwd_rows: bad arguments");
  B200_CHECK_ARG(rows <= 2147483647LL, "b200_softmax_bwd_rows: too many rows");
  softmax_bwd_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const __half*)P, ldp, dP, ldd, (__half*)dS,
                                                                            cols, scale);
  B200_CHECK_LAUNCH("softmax_bwd_rows_kernel");
  return 0;
}

extern "C" int b200_act_bwd(const void* x, const void* dy, long long n, int act, void* dx, void* stream) {
  B200_CHECK_ARG(x && dy && dx && n > 0, "b200_act_bwd: bad arguments");
  B200_CHECK_ARG(act == B200_ACT_SILU || act == B200_ACT_GELU, "b200_act_bwd: act=%d (SiLU or GELU)", act);
  act_bwd_kernel<<<bw_grid1d(n, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, (const __half*)dy, n,
                                                                      act == B200_ACT_SILU ? 1 : 3, (__half*)dx);
  B200_CHECK_LAUNCH("act_bwd_kernel");
  return 0;
}

extern "C" int b200_geglu_bwd(const void* h, const void* g, long long ld_hg, const void* dy, long long rows, int inner,
                              void* dh, void* dg, long long ld_d, void* stream) {
  B200_CHECK_ARG(h && g && dy && dh && dg && rows > 0 && inner > 0 && ld_hg >= inner && ld_d >= inner,
                 "b200_geglu_bwd: bad arguments");
  geglu_bwd_kernel<<<bw_grid1d(rows * inner, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)h, (const __half*)g, ld_hg, (const __half*)dy, rows, inner, (__half*)dh, (__half*)dg, ld_d);
  B200_CHECK_LAUNCH("geglu_bwd_kernel");
  return 0;
}

static dim3 bw_loss_grid(long long HW, int B) {
  long long g = (HW + 256 * 8 - 1) / (256 * 8);
  long long cap = (long long)sm_count() * 4 / (B > 0 ? B : 1) + 1;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return dim3((unsigned)g, B);
}

extern "C" int b200_ssi_loss_bwd(const float* pred, const float* target, const unsigned char* mask, int B,
                                 long long HW, double* workspace, const float* grad_out, float* dpred, void* stream) {
  B200_CHECK_ARG(pred && target && mask && workspace && grad_out && dpred && B > 0 && HW > 0,
                 "b200_ssi_loss_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid = bw_loss_grid(HW, B);
  ssi_bwd_moments_kernel<<<grid, 256, 0, st>>>(pred, target, mask, HW, workspace);
  ssi_bwd_sign_sums_kernel<<<grid, 256, 0, st>>>(pred, target, mask, HW, B, workspace);
  ssi_bwd_grad_kernel<<<grid, 256, 0, st>>>(pred, target, mask, HW, B, workspace, grad_out, dpred);
  B200_CHECK_LAUNCH("ssi_loss_bwd kernels");
  return 0;
}

extern "C" int b200_angular_loss_bwd(const float* pred, const float* target, const unsigned char* mask, int B,
                                     long long HW, double* workspace, const float* grad_out, float* dpred,
                                     void* stream) {
  B200_CHECK_ARG(pred && target && mask && workspace && grad_out && dpred && B > 0 && HW > 0,
                 "b200_angular_loss_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  mask_count_kernel<<<bw_grid1d((long long)B * HW, 256), 256, 0, st>>>(mask, (long long)B * HW, workspace);
  angular_bwd_kernel<<<bw_loss_grid(HW, B), 256, 0, st>>>(pred, target, mask, HW, workspace, grad_out, dpred);
  B200_CHECK_LAUNCH("angular_loss_bwd kernels");
  return 0;
}

extern "C" int b200_decode_post_bwd(const float* x, const float* dout, int NB, long long HW, int mode, float* dx,
                                    void* stream) {
  B200_CHECK_ARG(x && dout && dx && NB > 0 && HW > 0 && (mode == 2 || mode == 3), "b200_decode_post_bwd: bad arguments");
  decode_post_bwd_kernel<<<bw_loss_grid(HW, NB), 256, 0, (cudaStream_t)stream>>>(x, dout, HW, mode, dx);
  B200_CHECK_LAUNCH("decode_post_bwd_kernel");
  return 0;
}

extern "C" int b200_upsample_nearest_bwd(const float* dy, int NB, int H, int W, int C, int OH, int OW, const float* add,
                                         float* dx, void* stream) {
  B200_CHECK_ARG(dy && dx && NB > 0 && H > 0 && W > 0 && OH >= H && OW >= W, "b200_upsample_nearest_bwd: bad arguments");
  B200_CHECK_ARG(C % 4 == 0, "b200_upsample_nearest_bwd: C=%d must be a multiple of 4", C);
  const long long total = (long long)NB * H * W * (C / 4);
  upsample_nearest_bwd_kernel<<<bw_grid1d(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, NB, H, W, C, OH, OW, add, dx);
  B200_CHECK_LAUNCH("upsample_nearest_bwd_kernel");
  return 0;
}
