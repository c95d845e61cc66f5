// CUDA-core kernels of the Mobile-URSONet forward: stem 3x3/s2, depthwise 3x3 (s1/s2), pointwise
// 1x1 as a SIMT FP32 GEMM (the FP32 path and the debug cross-check of the tcgen05 path), global mean.
// Activations are NHWC (channel contiguous = GEMM K); BN is folded on the host (spef_api.cu).
// Reference semantics: src/modeling/common/pytorch_layers.py:35-98, src/modeling/backbone/mobilenet_v2.py:232-271,
// src/modeling/head/ursonet.py:27-33.
#pragma once
#include "common.cuh"

namespace spef {

// --------------------------------------------------------------------------------------------------
// Stem: ConvBnAct(3->32, k3, s2, p1) (mobilenet_v2.py:252-254).  Reads the reference's own tensor
// contract ([B,3,H,W] f32 NCHW), writes NHWC.  4 threads per pair of adjacent output pixels, 8 output channels each: the
// 45 taps of the pair are loaded up front (memory-level parallelism), shared between the 4 lanes through L1, and
// each 64-byte (bf16) pixel is written by 4 adjacent lanes.  Weights [27][32] + bias in shared memory.
// --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) stem_conv3x3s2_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                             const float* __restrict__ bias, T* __restrict__ out,
                                                             int B, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) float ws[27 * 32];
  __shared__ __align__(16) float bs[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 32) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();

  // one thread = 2 horizontally adjacent output pixels x 8 output channels
  const int Wp = (Wo + 1) >> 1;
  const long long total = (long long)B * Ho * Wp * 4;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= total) return;
  const int cg = (int)(tid & 3);
  long long p = tid >> 2;
  const int oxp = (int)(p % Wp); p /= Wp;
  const int oy = (int)(p % Ho);
  const int b = (int)(p / Ho);
  const int ox = oxp * 2;

  // all 45 taps are loaded up front (branch-free: clamped address, zero mask) so the loads overlap
  float in[3][3][5];
  const float* ib = img + (size_t)b * 3 * H * W;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    const bool yok = (iy >= 0) && (iy < H);
    const int iyc = min(max(iy, 0), H - 1);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ix = ox * 2 - 1 + j;
      const bool ok = yok && (ix >= 0) && (ix < W);
      const int ixc = min(max(ix, 0), W - 1);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        float v = __ldg(ib + ((size_t)ci * H + iyc) * W + ixc);
        // BF16 engine: the image is a tensor-core operand on the default path, so the cross-check rounds it the same way
        if constexpr (sizeof(T) == 2) v = __bfloat162float(__float2bfloat16_rn(v));
        in[ci][ky][j] = ok ? v : 0.f;
      }
    }
  }
  // packed FP32 pairs: 4 FFMA2 per tap and pixel instead of 8 FFMA
  uint64_t acc[2][4];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = f32x2(bs[cg * 8 + 2 * j], bs[cg * 8 + 2 * j + 1]);
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * 3 + ky) * 3 + kx) * 32 + cg * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        const uint64_t wv[4] = {f32x2(w0.x, w0.y), f32x2(w0.z, w0.w), f32x2(w1.x, w1.y), f32x2(w1.z, w1.w)};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float v = in[ci][ky][kx + 2 * t];
          const uint64_t vv = f32x2(v, v);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[t][e] = fma_f32x2(vv, wv[e], acc[t][e]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    if (ox + t < Wo) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f32x2_unpack(acc[t][j], o[2 * j], o[2 * j + 1]);
        o[2 * j] = relu_nan(o[2 * j]);
        o[2 * j + 1] = relu_nan(o[2 * j + 1]);
      }
      Vec8<T>::store(out + (((size_t)b * Ho + oy) * Wo + ox + t) * 32 + cg * 8, o);
    }
  }
}

// --------------------------------------------------------------------------------------------------
// Depthwise ConvBnAct(C->C, k3, p1, stride S, groups=C) (pytorch_layers.py:82-83), NHWC.
// One thread = 8 channels x TX consecutive output pixels of one output row: a (TX-1)*S+3 column
// register window per input row, so every loaded 16-byte vector feeds up to 3 outputs.  Channel
// groups are the fastest thread index => a warp reads/writes contiguous channel runs.
// Weights [9][C] f32 (folded), bias [C].
// --------------------------------------------------------------------------------------------------
template <typename T, int S, int TX>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ out,
                                                        int B, int H, int W, int C, int Ho, int Wo, int relu) {
  constexpr int NCOLS = (TX - 1) * S + 3;
  const int CG = C >> 3;
  const int nstrips = (Wo + TX - 1) / TX;
  const long long total = (long long)B * Ho * nstrips * CG;
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= total) return;
  const int cg = (int)(tid % CG); tid /= CG;
  const int strip = (int)(tid % nstrips); tid /= nstrips;
  const int oy = (int)(tid % Ho);
  const int b = (int)(tid / Ho);
  const int c0 = cg * 8;
  const int ox0 = strip * TX;

  float wr[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k) Vec8<float>::load(w + (size_t)k * C + c0, wr[k]);
  float acc[TX][8];
  {
    float bv[8];
    Vec8<float>::load(bias + c0, bv);
#pragma unroll
    for (int t = 0; t < TX; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = bv[j];
  }

  const T* ib = in + (size_t)b * H * W * C + c0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * S - 1 + ky;
    if (iy < 0 || iy >= H) continue;
    const T* row = ib + (size_t)iy * W * C;
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      const int ix = ox0 * S - 1 + j;
      if (ix < 0 || ix >= W) continue;
      float v[8];
      Vec8<T>::load(row + (size_t)ix * C, v);
#pragma unroll
      for (int t = 0; t < TX; ++t) {
        const int kx = j - t * S;  // compile-time after unrolling
        if (kx >= 0 && kx <= 2) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(v[e], wr[ky * 3 + kx][e], acc[t][e]);
        }
      }
    }
  }
  T* ob = out + (((size_t)b * Ho + oy) * Wo) * C + c0;
#pragma unroll
  for (int t = 0; t < TX; ++t) {
    const int ox = ox0 + t;
    if (ox < Wo) {
      if (relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = relu_nan(acc[t][e]);
      }
      Vec8<T>::store(ob + (size_t)ox * C, acc[t]);
    }
  }
}

// --------------------------------------------------------------------------------------------------
// Pointwise 1x1 conv / Linear as a SIMT GEMM with FP32 FMAs:  D[M,N] = act(A[M,K] * Wt[K,N] + bias) (+ R)
// (pytorch_layers.py:78-79, 85-86, 93-98; mobilenet_v2.py:264; ursonet.py:31-32).  This is the FP32 path
// (tcgen05 has no FP32-input MMA; single-pass TF32 would miss the 1e-4 logits gate) and
// the cross-check of the tcgen05 kernel.  64x64 tile, BK = 8, 256 threads, 4x4 outputs per thread.
// K % 8 == 0 and N % 4 == 0 are guaranteed by the architecture (all channel counts are multiples of 8;
// the concatenated head width is padded to 8).  ldd = row pitch of D in elements.
// --------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) pw_gemm_simt_kernel(const TIn* __restrict__ A, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, const TIn* __restrict__ R,
                                                           TOut* __restrict__ D, int M, int N, int K, int ldd, int relu) {
  constexpr int BM = 64, BN = 64, BK = 8;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (tid < BM) {
      float v[8];
      const int m = m0 + tid;
      if (m < M) {
        Vec8<TIn>::load(A + (size_t)m * K + k0, v);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) As[e][tid] = v[e];
    } else if (tid < BM + 128) {
      const int t = tid - BM;          // 0..127
      const int kk = t >> 4, nn = (t & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + nn < N) v = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kk) * N + n0 + nn));
      *reinterpret_cast<float4*>(&Ws[kk][nn]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = acc[i][j] + bv[j];
      if (relu) o[j] = relu_nan(o[j]);
    }
    if (R != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += to_f32<TIn>(R[(size_t)m * N + n + j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) D[(size_t)m * ldd + n + j] = from_f32<TOut>(o[j]);
  }
}

// --------------------------------------------------------------------------------------------------
// Global mean over H*W (head/ursonet.py:30): in [B,HW,C] -> out [B,C], f32 accumulate, divide by HW.
// --------------------------------------------------------------------------------------------------
// One CTA = one image x 32 channel groups (8 channels each: a warp reads 512 contiguous bytes of a pixel); the eight warps take
// the pixels w, w + 8, ... with all their loads in flight, and warp 0 adds the eight partial sums in a fixed order.  (The first
// version -- one thread per (image, channel group) walking the 96 pixels with one dependent load at a time -- ran at 2.4 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256) global_mean_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int HW, int C) {
  __shared__ float part[8][32][9];   // [pixel part][channel group][8 channels + pad]
  const int CG = C >> 3;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int groups_per_img = (CG + 31) / 32;
  const int b = blockIdx.x / groups_per_img, cg = (blockIdx.x % groups_per_img) * 32 + lane;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cg < CG) {
    const T* p = in + (size_t)b * HW * C + cg * 8;
    int i = w;
    for (; i + 24 < HW; i += 32) {   // four pixels of this warp in flight
      float v0[8], v1[8], v2[8], v3[8];
      Vec8<T>::load(p + (size_t)i * C, v0);
      Vec8<T>::load(p + (size_t)(i + 8) * C, v1);
      Vec8<T>::load(p + (size_t)(i + 16) * C, v2);
      Vec8<T>::load(p + (size_t)(i + 24) * C, v3);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = ((acc[e] + v0[e]) + v1[e]) + v2[e] + v3[e];
    }
    for (; i < HW; i += 8) {
      float v[8];
      Vec8<T>::load(p + (size_t)i * C, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[w][lane][e] = acc[e];
  __syncthreads();
  if (w == 0 && cg < CG) {
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = part[0][lane][e];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += part[k][lane][e];
      s[e] = t / (float)HW;
    }
    Vec8<T>::store(out + (size_t)b * C + cg * 8, s);
  }
}

}  // namespace spef
