// Small HBM-bound helpers around the tensor-core kernels: patch extraction for the tiny-Cin input
// convolutions, nearest upsample, timestep sinusoid, latent channel mixes, decode post-ops,
// boundary casts.
#include "common.cuh"
#include "../../include/b200_e2eft.h"

namespace b200 {

// out[pixel][tap*C + c] (fp16, row length Kpad, zero padded), x NCHW.  One thread per pixel: it gathers the
// 9*C (<= 72) neighbours (L1/L2-served: neighbouring threads share them) and writes its whole patch row with
// 16-byte stores, so the 64-128 byte rows leave as full sectors.
template <typename T, int KPAD>
__global__ void im2col3x3_kernel(const T* __restrict__ x, int NB, int C, int H, int W,
                                 __half* __restrict__ out) {
  const long long total = (long long)NB * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    constexpr int CC = KPAD == 32 ? 3 : (KPAD == 40 ? 4 : 8);       // compile-time channel count: row[] stays in registers
    __align__(16) __half row[KPAD];
#pragma unroll
    for (int k = 0; k < KPAD; ++k) row[k] = __float2half_rn(0.f);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
      for (int c = 0; c < CC; ++c)
        row[tap * CC + c] = ok ? __float2half_rn((float)x[(((long long)n * CC + c) * H + hh) * W + ww])
                               : __float2half_rn(0.f);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + pix * KPAD);
    const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int v = 0; v < KPAD / 8; ++v) dst[v] = src[v];
  }
}

template <typename T>
__global__ void upsample_nearest_kernel(const T* __restrict__ x, int NB, int H, int W, int C, int OH,
                                        int OW, __half* __restrict__ y) {
  const int V = C / 8;
  const long long total = (long long)NB * OH * OW * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    long long pix = i / V;
    const int ow = (int)(pix % OW);
    const int oh = (int)((pix / OW) % OH);
    const int n = (int)(pix / ((long long)OW * OH));
    // torch nearest: src = floor(dst * in / out)
    const int ih = min((int)(((long long)oh * H) / OH), H - 1);
    const int iw = min((int)(((long long)ow * W) / OW), W - 1);
    const T* src = x + (((long long)n * H + ih) * W + iw) * C + v * 8;
    __half* dst = y + pix * C + v * 8;
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    } else {
      float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = u;
    }
  }
}

// emb[b] = [cos(t*f_0..f_{h-1}), sin(t*f_0..)], f_i = exp(-ln(1e4) * i / half)   (flip_sin_to_cos)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim,
                                          __half* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);
  const float a = t[b] * freq;
  out[(long long)b * dim + j] = __float2half_rn(cosf(a));
  out[(long long)b * dim + half + j] = __float2half_rn(sinf(a));
}

// CLIP text embeddings (transformers modeling_clip CLIPTextEmbeddings): out[b*L + l][c] = tok[ids[b*L + l]][c] + pos[l][c],
// fp32 residual stream out; the tables are fp16 or fp32 (the module's parameter dtype).  One warp-strided pass, 8 B+ per lane.
template <typename T>
__global__ void embed_tokens_kernel(const long long* __restrict__ ids, const T* __restrict__ tok, const T* __restrict__ pos,
                                    long long rows, int L, int C, int vocab, float* __restrict__ out) {
  const long long total = rows * (long long)(C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C / 4);
    const int c = (int)(i - r * (C / 4)) * 4;
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const T* a = tok + id * C + c;
    const T* b = pos + (r % L) * C + c;
    float4 o;
    if constexpr (sizeof(T) == 2) {
      const __half2 a0 = reinterpret_cast<const __half2*>(a)[0], a1 = reinterpret_cast<const __half2*>(a)[1];
      const __half2 b0 = reinterpret_cast<const __half2*>(b)[0], b1 = reinterpret_cast<const __half2*>(b)[1];
      const float2 fa0 = __half22float2(a0), fa1 = __half22float2(a1), fb0 = __half22float2(b0), fb1 = __half22float2(b1);
      o = make_float4(fa0.x + fb0.x, fa0.y + fb0.y, fa1.x + fb1.x, fa1.y + fb1.y);
    } else {
      const float4 fa = *reinterpret_cast<const float4*>(a), fb = *reinterpret_cast<const float4*>(b);
      o = make_float4(fa.x + fb.x, fa.y + fb.y, fa.z + fb.z, fa.w + fb.w);
    }
    *reinterpret_cast<float4*>(out + r * C + c) = o;
  }
}

__global__ void pointwise_nchw_kernel(const float* __restrict__ in1, float a1,
                                      const float* __restrict__ in2, float a2, int in_cstride,
                                      const float* __restrict__ Wm, const float* __restrict__ bias,
                                      int Cin, int Cout, long long HW, float* __restrict__ out) {
  __shared__ float w[64 + 8];
  if (threadIdx.x < Cin * Cout) w[threadIdx.x] = Wm[threadIdx.x];
  if (threadIdx.x < Cout) w[64 + threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (long long)gridDim.x * blockDim.x) {
    float v[8];
    for (int ci = 0; ci < Cin; ++ci) {
      float t = a1 * in1[((long long)n * in_cstride + ci) * HW + p];
      if (in2) t += a2 * in2[((long long)n * in_cstride + ci) * HW + p];
      v[ci] = t;
    }
    for (int co = 0; co < Cout; ++co) {
      float acc = w[64 + co];
      for (int ci = 0; ci < Cin; ++ci) acc += w[co * Cin + ci] * v[ci];
      out[((long long)n * Cout + co) * HW + p] = acc;
    }
  }
}

__global__ void decode_post_kernel(const float* __restrict__ x, long long HW, int mode, float sign,
                                   float* __restrict__ out) {
  const int n = blockIdx.y;
  const float* xb = x + (long long)n * 3 * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (long long)gridDim.x * blockDim.x) {
    const float a = xb[p], b = xb[HW + p], c = xb[2 * HW + p];
    if (mode == 0 || mode == 2) {
      float m = (a + b + c) / 3.0f;
      m = fminf(fmaxf(m, -1.0f), 1.0f);
      out[(long long)n * HW + p] = mode == 0 ? (m + 1.0f) / 2.0f : m;     // mode 2: training (train.py:533-534)
    } else {
      const float inv = sign / (sqrtf(a * a + b * b + c * c) + 1e-5f);
      float* ob = out + (long long)n * 3 * HW;
      float v0 = a * inv, v1 = b * inv, v2 = c * inv;
      if (mode == 3) {                                                     // training: clamp (train.py:539)
        v0 = fminf(fmaxf(v0, -1.f), 1.f); v1 = fminf(fmaxf(v1, -1.f), 1.f); v2 = fminf(fmaxf(v2, -1.f), 1.f);
      }
      ob[p] = v0;
      ob[HW + p] = v1;
      ob[2 * HW + p] = v2;
    }
  }
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(x[i]);
}

// NHWC -> NCHW fp32 through a 32x32 smem transpose tile (coalesced both sides).
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int C, long long HW, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long p = p0 + r;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[r][threadIdx.x] = (float)x[((long long)n * HW + p) * C + c];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long p = p0 + threadIdx.x;
    if (p < HW && c < C) y[((long long)n * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

static int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)sm_count() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_im2col3x3_nchw(const void* x, int x_f32, int NB, int C, int H, int W, void* out,
                                   int Kpad, void* stream) {
  B200_CHECK_ARG(x && out && NB > 0 && C > 0 && H > 0 && W > 0, "b200_im2col3x3_nchw: bad arguments");
  B200_CHECK_ARG(Kpad >= 9 * C && Kpad % 8 == 0, "b200_im2col3x3_nchw: Kpad=%d must be >= 9*C and %%8==0", Kpad);
  B200_CHECK_ARG((Kpad == 32 && C == 3) || (Kpad == 40 && C == 4) || (Kpad == 72 && C == 8),
                 "b200_im2col3x3_nchw: (C=%d, Kpad=%d) unsupported: C must be 3, 4 or 8 with Kpad = round_up(9C, 8)", C, Kpad);
  const long long total = (long long)NB * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total, 128);
#define B200_IM2COL(T, K) im2col3x3_kernel<T, K><<<g, 128, 0, st>>>((const T*)x, NB, C, H, W, (__half*)out)
  if (x_f32) {
    if (Kpad == 32) B200_IM2COL(float, 32); else if (Kpad == 40) B200_IM2COL(float, 40); else B200_IM2COL(float, 72);
  } else {
    if (Kpad == 32) B200_IM2COL(__half, 32); else if (Kpad == 40) B200_IM2COL(__half, 40); else B200_IM2COL(__half, 72);
  }
#undef B200_IM2COL
  B200_CHECK_LAUNCH("im2col3x3_kernel");
  return 0;
}

extern "C" int b200_upsample_nearest_nhwc(const void* x, int in_f32, int NB, int H, int W, int C, int OH,
                                          int OW, void* y, void* stream) {
  B200_CHECK_ARG(x && y && NB > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "b200_upsample_nearest_nhwc: bad arguments");
  B200_CHECK_ARG(C % 8 == 0, "b200_upsample_nearest_nhwc: C=%d must be a multiple of 8", C);
  const long long total = (long long)NB * OH * OW * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32)
    upsample_nearest_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)x, NB, H, W, C, OH, OW, (__half*)y);
  else
    upsample_nearest_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>((const __half*)x, NB, H, W, C, OH, OW, (__half*)y);
  B200_CHECK_LAUNCH("upsample_nearest_kernel");
  return 0;
}

extern "C" int b200_timestep_embedding(const float* t, int B, int dim, void* out, void* stream) {
  B200_CHECK_ARG(t && out && B > 0 && dim > 0 && dim % 2 == 0, "b200_timestep_embedding: bad arguments");
  const int n = B * (dim / 2);
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, B, dim, (__half*)out);
  B200_CHECK_LAUNCH("timestep_embedding_kernel");
  return 0;
}

extern "C" int b200_embed_tokens(const long long* ids, const void* tok, const void* pos, int w_f32, long long rows, int L,
                                 int C, int vocab, float* out, void* stream) {
  B200_CHECK_ARG(ids && tok && pos && out && rows > 0 && L > 0 && vocab > 0, "b200_embed_tokens: bad arguments");
  B200_CHECK_ARG(C > 0 && C % 4 == 0, "b200_embed_tokens: C=%d must be a multiple of 4", C);
  const long long total = rows * (long long)(C / 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_f32)
    embed_tokens_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(ids, (const float*)tok, (const float*)pos, rows, L, C, vocab, out);
  else
    embed_tokens_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(ids, (const __half*)tok, (const __half*)pos, rows, L, C, vocab, out);
  B200_CHECK_LAUNCH("embed_tokens_kernel");
  return 0;
}

extern "C" int b200_pointwise_nchw(const float* in1, float a1, const float* in2, float a2, int in_cstride,
                                   const float* Wm, const float* bias, int NB, int Cin, int Cout,
                                   long long HW, float* out, void* stream) {
  B200_CHECK_ARG(in1 && Wm && out && NB > 0 && HW > 0, "b200_pointwise_nchw: bad arguments");
  B200_CHECK_ARG(Cin >= 1 && Cin <= 8 && Cout >= 1 && Cout <= 8 && in_cstride >= Cin,
                 "b200_pointwise_nchw: Cin=%d Cout=%d must be in [1,8]", Cin, Cout);
  dim3 grid(grid_for(HW, 256), NB);
  pointwise_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in1, a1, in2, a2, in_cstride, Wm, bias, Cin, Cout, HW, out);
  B200_CHECK_LAUNCH("pointwise_nchw_kernel");
  return 0;
}

extern "C" int b200_decode_post(const float* x, int NB, long long HW, int mode, float sign, float* out,
                                void* stream) {
  B200_CHECK_ARG(x && out && NB > 0 && HW > 0 && mode >= 0 && mode <= 3, "b200_decode_post: bad arguments");
  dim3 grid(grid_for(HW, 256), NB);
  decode_post_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, HW, mode, sign, out);
  B200_CHECK_LAUNCH("decode_post_kernel");
  return 0;
}

extern "C" int b200_cast_f32_to_f16(const float* x, void* y, long long n, void* stream) {
  B200_CHECK_ARG(x && y && n > 0, "b200_cast_f32_to_f16: bad arguments");
  cast_f32_f16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)y, n);
  B200_CHECK_LAUNCH("cast_f32_f16_kernel");
  return 0;
}

extern "C" int b200_nhwc_to_nchw_f32(const void* x, int in_f32, int NB, int C, long long HW, float* y,
                                     void* stream) {
  B200_CHECK_ARG(x && y && NB > 0 && C > 0 && HW > 0, "b200_nhwc_to_nchw_f32: bad arguments");
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, NB), block(32, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32)
    nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)x, C, HW, y);
  else
    nhwc_to_nchw_kernel<__half><<<grid, block, 0, st>>>((const __half*)x, C, HW, y);
  B200_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return 0;
}
