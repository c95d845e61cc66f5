// Input side of the path: the transform SPEDataset.__getitem__ applies to every camera frame
// (src/data/utils.py:212-226): Image.convert("RGB") -> transforms.Resize(img_size) -> transforms.ToTensor()
// (src/data/datasets/speed.py:59-62).  On a PIL image, Resize is Pillow's antialiased BILINEAR resample: a separable
// triangle filter whose support grows with the reduction factor, evaluated on 8-bit pixels in 22-bit fixed point,
// horizontal pass first, each pass rounded (+2^21) and clipped to 8 bits; ToTensor is float32(u8) / 255.
// One kernel does both passes for a band of output rows (the horizontally filtered rows stay in shared memory), writes
// the planar [B,3,h,w] tensor the stem reads (uint8, or the float32 of the reference contract) and is bit-exact
// against torchvision + Pillow (tests/golden/resize.npz).  HBM-bound: every frame byte is read once.
#pragma once
#include "common.cuh"
#include <cmath>
#include <vector>

namespace spef {
namespace ingest {

constexpr int kFixBits = 32 - 8 - 2;  // fixed-point fraction bits of the 8-bit resampler
constexpr int KWIN = 12;              // taps of the fixed horizontal window (covers reductions up to 5.5x)

// Host: taps of one axis.  first[o] / count[o] = window of input samples of output o, coef[o * ksize + j] = fixed-point
// weight of sample first[o] + j.  All arithmetic in float64, in the operation order of the resampler being reproduced.
struct AxisTaps {
  int ksize = 0;
  std::vector<int> first, count;
  std::vector<int32_t> coef;
};

inline AxisTaps make_axis_taps(int in_size, int out_size) {
  AxisTaps t;
  const double scale = (double)in_size / out_size;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * fscale;  // triangle filter: support 1, stretched by the reduction factor
  t.ksize = (int)std::ceil(support) * 2 + 1;
  t.first.resize(out_size);
  t.count.resize(out_size);
  t.coef.assign((size_t)out_size * t.ksize, 0);
  const double inv = 1.0 / fscale;
  std::vector<double> w(t.ksize);
  for (int o = 0; o < out_size; ++o) {
    const double center = (o + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    double total = 0.0;
    for (int j = 0; j < n; ++j) {
      double x = (j + lo - center + 0.5) * inv;
      if (x < 0.0) x = -x;
      w[j] = x < 1.0 ? 1.0 - x : 0.0;
      total += w[j];
    }
    for (int j = 0; j < n; ++j) {
      const double v = (total != 0.0) ? w[j] / total : w[j];
      t.coef[(size_t)o * t.ksize + j] = v < 0 ? (int)(-0.5 + v * (1 << kFixBits)) : (int)(0.5 + v * (1 << kFixBits));
    }
    t.first[o] = lo;
    t.count[o] = n;
  }
  return t;
}

struct ResizeParams {
  const uint8_t* src;  // [B, sh, sw, C] (HWC, C = 1 | 3)
  void* dst;           // [B, 3, oh, ow] uint8 or float32
  const int* hfirst;   // [ow]
  const int* hcount;   // [ow]
  const int32_t* hcoef;  // [hks][ow]  (tap-major: lanes read consecutive words); fixed-window plan: [KWIN][ow], see resize_plan
  const int* vfirst;   // [oh]
  const int* vcount;   // [oh]
  const int32_t* vcoef;  // [oh][vks]
  int sh, sw, C, oh, ow, hks, vks;
  int band;            // output rows per CTA
  int max_rows;        // input rows a band can need
  int pitch;           // bytes per filtered row in shared memory (>= ow)
  int out_f32;
};

__device__ __forceinline__ uint8_t clip8(int32_t acc) {
  const int v = acc >> kFixBits;
  return (uint8_t)min(max(v, 0), 255);
}

// KREG > 0: the horizontal taps of a thread's output column live in registers (ksize <= KREG); 0: read per use.
template <int KREG, int C>
__global__ void __launch_bounds__(384) resize_aa_kernel(const ResizeParams p) {
  extern __shared__ uint8_t filtered[];  // [C][max_rows][pitch]
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * p.band, y1 = min(y0 + p.band, p.oh);
  const int r0 = p.vfirst[y0];
  const int r1 = p.vfirst[y1 - 1] + p.vcount[y1 - 1];  // windows move monotonically with the output index
  const uint8_t* img = p.src + (size_t)b * p.sh * p.sw * C;

  // pass 1: rows r0..r1 of the frame, filtered along x
  for (int xx = threadIdx.x; xx < p.ow; xx += blockDim.x) {
    const int x0 = p.hfirst[xx], cnt = p.hcount[xx];
    if (KREG > 0) {
      // fixed window of KREG taps at compile-time offsets (the host shifts the window of the right-most outputs left so that
      // it never leaves the row, and pads the coefficients with zeros): no predicates, no per-tap address arithmetic
      int32_t k[KREG > 0 ? KREG : 1];
#pragma unroll
      for (int j = 0; j < KREG; ++j) k[j] = p.hcoef[(size_t)j * p.ow + xx];
      const uint8_t* px = img + ((size_t)r0 * p.sw + x0) * C;
      const size_t row_bytes = (size_t)p.sw * C;
      uint8_t* dst = filtered + xx;
#pragma unroll 2
      for (int r = r0; r < r1; ++r, px += row_bytes, dst += p.pitch) {
        int32_t acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 1 << (kFixBits - 1);
#pragma unroll
        for (int j = 0; j < KREG; ++j) {
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] += (int32_t)px[j * C + c] * k[j];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) dst[(size_t)c * p.max_rows * p.pitch] = clip8(acc[c]);
      }
    } else {
      for (int r = r0; r < r1; ++r) {
        const uint8_t* px = img + ((size_t)r * p.sw + x0) * C;
        int32_t acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 1 << (kFixBits - 1);
        for (int j = 0; j < cnt; ++j) {
          const int32_t kj = p.hcoef[(size_t)j * p.ow + xx];
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] += (int32_t)px[j * C + c] * kj;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) filtered[((size_t)c * p.max_rows + (r - r0)) * p.pitch + xx] = clip8(acc[c]);
      }
    }
  }
  __syncthreads();

  // pass 2: the band's output rows, filtered along y from shared memory; a grey frame is replicated like convert("RGB")
  const size_t plane = (size_t)p.oh * p.ow;
  for (int idx = threadIdx.x; idx < (y1 - y0) * p.ow; idx += blockDim.x) {
    const int y = y0 + idx / p.ow, xx = idx % p.ow;
    const int rb = p.vfirst[y] - r0, cnt = p.vcount[y];
    const int32_t* kv = p.vcoef + (size_t)y * p.vks;
    int32_t acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kFixBits - 1);
    for (int j = 0; j < cnt; ++j) {
      const int32_t kj = kv[j];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += (int32_t)filtered[((size_t)c * p.max_rows + rb + j) * p.pitch + xx] * kj;
    }
    const size_t o = (size_t)b * 3 * plane + (size_t)y * p.ow + xx;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const uint8_t v = clip8(acc[C == 1 ? 0 : ch]);
      if (p.out_f32) reinterpret_cast<float*>(p.dst)[o + ch * plane] = __fdiv_rn((float)v, 255.0f);  // ToTensor
      else reinterpret_cast<uint8_t*>(p.dst)[o + ch * plane] = v;
    }
  }
}

// Greyscale fast path (SPEED's frames): the frame rows a band needs are staged in shared memory RB rows at a time with
// 16-byte cp.async copies (double buffered), and a thread takes its 12-tap window as four aligned 32-bit words; the 22-bit
// coefficients are split into three byte planes laid out on the same word grid, so that a row costs 4 LDS.32 + 12 DP4A
// (u8 x u8 dot products, exact) instead of 12 byte loads + 12 IMAD -- the first version was bound by byte-load issue (LSU).
// Requirements (host): one channel, fixed-window plan, src_w a multiple of 16, 16-byte aligned frames.
struct GrayPlan {
  const int* hword;        // [ow] index of the first aligned word of the window
  const uint32_t* hplane;  // [3][4][ow] byte planes of the coefficients on the word grid (plane q = bits 8q .. 8q+7)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RB>
__global__ void __launch_bounds__(384) resize_aa_gray_kernel(const ResizeParams p, const GrayPlan g) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int stage_bytes = RB * p.sw + 16;                 // + one word the last window may over-read (zero coefficients)
  uint8_t* stage = gsm;                                   // [2][stage_bytes]
  uint8_t* filtered = gsm + 2 * stage_bytes;              // [max_rows][pitch]
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * p.band, y1 = min(y0 + p.band, p.oh);
  const int r0 = p.vfirst[y0];
  const int r1 = p.vfirst[y1 - 1] + p.vcount[y1 - 1];
  const uint8_t* img = p.src + (size_t)b * p.sh * p.sw;
  const int nchunks = (r1 - r0 + RB - 1) / RB;
  const uint32_t stage_u = (uint32_t)__cvta_generic_to_shared(stage);
  const int vec_per_row = p.sw >> 4;
  auto issue = [&](int c) {
    const int rows = min(RB, r1 - r0 - c * RB);
    const uint8_t* src = img + (size_t)(r0 + c * RB) * p.sw;   // RB consecutive rows are contiguous in the frame
    const uint32_t dst = stage_u + (uint32_t)((c & 1) * stage_bytes);
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) cp_async16(dst + (uint32_t)i * 16u, src + (size_t)i * 16);
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int rows = min(RB, r1 - r0 - c * RB);
    const uint32_t* sw32 = reinterpret_cast<const uint32_t*>(stage + (c & 1) * stage_bytes);
    const int row_words = p.sw >> 2;
    for (int xx = threadIdx.x; xx < p.ow; xx += blockDim.x) {
      const int wi = g.hword[xx];
      uint32_t k0[4], k1[4], k2[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        k0[w] = g.hplane[(size_t)(0 * 4 + w) * p.ow + xx];
        k1[w] = g.hplane[(size_t)(1 * 4 + w) * p.ow + xx];
        k2[w] = g.hplane[(size_t)(2 * 4 + w) * p.ow + xx];
      }
      const uint32_t* px = sw32 + wi;
      uint8_t* dst = filtered + (size_t)(c * RB) * p.pitch + xx;
#pragma unroll 4
      for (int rr = 0; rr < rows; ++rr, px += row_words, dst += p.pitch) {
        const uint32_t w0 = px[0], w1 = px[1], w2 = px[2], w3 = px[3];
        uint32_t s0 = __dp4a(w0, k0[0], 0u), s1 = __dp4a(w0, k1[0], 0u), s2 = __dp4a(w0, k2[0], 0u);
        s0 = __dp4a(w1, k0[1], s0); s1 = __dp4a(w1, k1[1], s1); s2 = __dp4a(w1, k2[1], s2);
        s0 = __dp4a(w2, k0[2], s0); s1 = __dp4a(w2, k1[2], s1); s2 = __dp4a(w2, k2[2], s2);
        s0 = __dp4a(w3, k0[3], s0); s1 = __dp4a(w3, k1[3], s1); s2 = __dp4a(w3, k2[3], s2);
        const int32_t acc = (int32_t)(s0 + (s1 << 8) + (s2 << 16)) + (1 << (kFixBits - 1));
        *dst = clip8(acc);
      }
    }
    __syncthreads();   // the buffer is refilled two chunks later
  }

  // pass 2 (same as resize_aa_kernel): filter along y from shared memory, replicate the grey plane three times
  const size_t plane = (size_t)p.oh * p.ow;
  for (int idx = threadIdx.x; idx < (y1 - y0) * p.ow; idx += blockDim.x) {
    const int y = y0 + idx / p.ow, xx = idx % p.ow;
    const int rb = p.vfirst[y] - r0, cnt = p.vcount[y];
    const int32_t* kv = p.vcoef + (size_t)y * p.vks;
    int32_t acc = 1 << (kFixBits - 1);
    for (int j = 0; j < cnt; ++j) acc += (int32_t)filtered[(size_t)(rb + j) * p.pitch + xx] * kv[j];
    const uint8_t v = clip8(acc);
    const size_t o = (size_t)b * 3 * plane + (size_t)y * p.ow + xx;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (p.out_f32) reinterpret_cast<float*>(p.dst)[o + ch * plane] = __fdiv_rn((float)v, 255.0f);  // ToTensor
      else reinterpret_cast<uint8_t*>(p.dst)[o + ch * plane] = v;
    }
  }
}

}  // namespace ingest
}  // namespace spef
