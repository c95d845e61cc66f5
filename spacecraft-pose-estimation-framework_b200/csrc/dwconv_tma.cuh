// Depthwise ConvBnAct(C->C, k3, p1, stride S, groups=C) for BF16 NHWC activations, TMA-tiled.
// (reference semantics: src/modeling/common/pytorch_layers.py:82-83 with BN folded, ReLU)
//
// Persistent CTAs loop over (channel chunk, image, tile_y, tile_x) tiles.  One thread issues a 4-D TMA box
// {CV*8 channels, TWI cols, THI rows, 1 image} of the input into shared memory (double buffered, mbarrier
// complete_tx); the 1-pixel halo and the image border come for free from TMA out-of-bounds zero fill (box start
// at x0-1, y0-1), so the SM executes no address / bounds logic for loads.  Each thread owns one 8-channel vector
// (fixed for the whole kernel => its 9x8 folded weights live in registers) and computes runs of 4 output pixels
// along x with a register sliding window over the smem tile: (3S+3) x 3 LDS.128 per 4 outputs instead of 36.
// Outputs go straight from registers to global memory: 8 lanes x 16 B = one full 128-byte line per pixel.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_tcgen05.cuh"  // mbarrier / TMA PTX wrappers

namespace spef {
namespace dw {


struct DwParams {
  int B, H, W, C, Ho, Wo;
  int TH, TW;        // output tile (TW multiple of the kernel's TX)
  int THI, TWI;      // input box rows / cols = (T-1)*S + 3
  int tiles_y, tiles_x, nchunks;
  int relu;
  int pdl_early;     // trigger the dependent launch at once (common.cuh)
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// two f32 (packed in a b64) -> bf16x2, low half = first element; the .relu form folds max(x, 0) into the conversion
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}

__host__ __device__ inline int stage_stride(int thi, int twi, int cv) { return ((thi * twi * cv * 16 + 127) / 128) * 128; }
inline size_t smem_bytes(const DwParams& p, int cv) { return 2 * (size_t)stage_stride(p.THI, p.TWI, cv) + 128 + 64; }

// TX: outputs per thread along x (4; 3 for the 12-column maps, where 3 strips x 8 rows would leave a quarter of the threads idle)
template <int S, int CV, int TX = 4>
__global__ void __launch_bounds__(32 * CV, (CV == 4) ? 4 : 2)
dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, const float* __restrict__ bias,
                     bf16* __restrict__ out, const DwParams p) {
  constexpr int NT = 32 * CV;
  constexpr int NCOLS = (TX - 1) * S + 3;
  if (p.pdl_early) pdl_launch_dependents();   // the next kernel of the chain may start its own set-up now (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int tile_bytes = p.THI * p.TWI * CV * 16;               // TMA transaction size (full box, OOB included)
  const int sstride = stage_stride(p.THI, p.TWI, CV);           // stage pitch, 128-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)sstride);

  const int tid = threadIdx.x;
  const int cv = tid % CV;
  const int nxs = p.TW / TX;
  const int tasks = nxs * p.TH;            // (x strip, row) pairs per tile; each is done by CV threads
  const int task0 = tid / CV, task_step = NT / CV;
  const long long spatial_tiles = (long long)p.B * p.tiles_y * p.tiles_x;
  // Channel chunk fixed per CTA (weights stay in registers); CTAs b, b+1, ... b+nchunks-1 walk the same spatial tiles at
  // the same time, so the partially used 128-byte lines of a pixel (CV*16 of C*2 bytes) are consumed out of L2 instead of
  // being fetched from HBM once per chunk.  gridDim.x is a multiple of nchunks (host).
  const int chunk = (int)(blockIdx.x % p.nchunks);
  const long long sp0 = blockIdx.x / p.nchunks;
  const long long sp_step = gridDim.x / p.nchunks;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::mbar_init(tc::smem_u32(&full_bar[0]), 1);
    tc::mbar_init(tc::smem_u32(&full_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();   // everything above is independent of the previous kernel's output (common.cuh)

  auto issue = [&](long long sp, int stage) {
    long long r = sp;
    const int tx = (int)(r % p.tiles_x); r /= p.tiles_x;
    const int ty = (int)(r % p.tiles_y);
    const int b = (int)(r / p.tiles_y);
    const uint32_t bar = tc::smem_u32(&full_bar[stage]);
    tc::mbar_arrive_expect_tx(bar, (uint32_t)tile_bytes);
    tma_load_4d(tc::smem_u32(smem + (size_t)stage * sstride), &tmX, chunk * CV * 8, tx * p.TW * S - 1, ty * p.TH * S - 1, b, bar);
  };

  long long sp = sp0;
  if (tid == 0) {
    if (sp < spatial_tiles) issue(sp, 0);
    if (sp + sp_step < spatial_tiles) issue(sp + sp_step, 1);
  }

  uint64_t wr[9][4];   // folded weights of this thread's 8 channels as packed FP32 pairs (FFMA2 operands)
  uint64_t bv[4];
  const int c0 = chunk * CV * 8 + cv * 8;
  const bool c_ok = c0 < p.C;
  if (c_ok) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float t8[8];
      Vec8<float>::load(w + (size_t)k * p.C + c0, t8);
#pragma unroll
      for (int e = 0; e < 4; ++e) wr[k][e] = f32x2(t8[2 * e], t8[2 * e + 1]);
    }
    float b8[8];
    Vec8<float>::load(bias + c0, b8);
#pragma unroll
    for (int e = 0; e < 4; ++e) bv[e] = f32x2(b8[2 * e], b8[2 * e + 1]);
  }
  // This thread's (strip, row) tasks are the same in every tile: decode them once (ncu: the per-task divisions, 64-bit
  // address arithmetic and generic-space loads of the first version made FFMA2 only 19 % of the instruction stream).
  // S == 1 with 64/96-byte pixels: consecutive lane groups take vertically adjacent rows and the host picks the box width
  // so that the row pitch is 64 resp. 96 (mod 128) bytes => the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank
  // groups.  Otherwise strips along x are adjacent.
  constexpr bool rows_fastest = (S == 1 && CV != 8);
  constexpr int MAX_TASKS = 2;             // tasks per thread and tile (checked by the host: TW / TX * TH <= MAX_TASKS * 32)
  int t_xs[MAX_TASKS], t_ry[MAX_TASKS];
  uint32_t t_off[MAX_TASKS];               // byte offset of the task's first input pixel inside a stage
#pragma unroll
  for (int k = 0; k < MAX_TASKS; ++k) {
    const int task = task0 + k * task_step;
    const int xs = rows_fastest ? task / p.TH : task % nxs;
    const int ry = rows_fastest ? task % p.TH : task / nxs;
    t_xs[k] = (task < tasks) ? xs : -1;
    t_ry[k] = ry;
    t_off[k] = (uint32_t)(((ry * S) * p.TWI + xs * TX * S) * (CV * 16) + cv * 16);
  }
  const uint32_t row_pitch = (uint32_t)(p.TWI * CV * 16);
  const uint32_t smem_u = tc::smem_u32(smem);
  // tile coordinates advance by sp_step tiles per iteration: carry (b, ty, tx) instead of dividing
  int tx, ty, b;
  {
    long long r = sp0;
    tx = (int)(r % p.tiles_x); r /= p.tiles_x;
    ty = (int)(r % p.tiles_y);
    b = (int)(r / p.tiles_y);
  }
  const int step_tx = (int)(sp_step % p.tiles_x);
  const int step_ty = (int)((sp_step / p.tiles_x) % p.tiles_y);
  const int step_b = (int)(sp_step / ((long long)p.tiles_x * p.tiles_y));
  uint32_t phases = 0;  // bit s = parity of stage s
  int stage = 0;
  for (; sp < spatial_tiles; sp += sp_step) {
    tc::mbar_wait(tc::smem_u32(&full_bar[stage]), (phases >> stage) & 1u);
    phases ^= 1u << stage;
    const uint32_t tile_u = smem_u + (uint32_t)stage * (uint32_t)sstride;

    if (c_ok) {
#pragma unroll
      for (int k = 0; k < MAX_TASKS; ++k) {
        if (t_xs[k] < 0) break;
        const int oy = ty * p.TH + t_ry[k];
        const int ox0 = tx * p.TW + t_xs[k] * TX;
        if (oy >= p.Ho || ox0 >= p.Wo) continue;
        uint64_t acc[TX][4];
#pragma unroll
        for (int t = 0; t < TX; ++t)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[t][e] = bv[e];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t row_u = tile_u + t_off[k] + (uint32_t)ky * row_pitch;
#pragma unroll
          for (int j = 0; j < NCOLS; ++j) {
            const uint4 u = tc::lds_u4(row_u + (uint32_t)(j * CV * 16));
            // bf16 pair -> packed f32 pair: low half << 16, high half masked
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
            uint64_t v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = f32x2(__uint_as_float(uw[e] << 16), __uint_as_float(uw[e] & 0xffff0000u));
#pragma unroll
            for (int t = 0; t < TX; ++t) {
              const int kx = j - t * S;
              if (kx >= 0 && kx <= 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][e] = fma_f32x2(v[e], wr[ky * 3 + kx][e], acc[t][e]);
              }
            }
          }
        }
        bf16* op = out + (((size_t)b * p.Ho + oy) * p.Wo + ox0) * p.C + c0;
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          if (ox0 + t < p.Wo) {
            uint4 o;
            if (p.relu) {   // ReLU folded into the conversion (max(x, 0) then round == round then max)
              o.x = cvt_relu_bf16x2(acc[t][0]); o.y = cvt_relu_bf16x2(acc[t][1]);
              o.z = cvt_relu_bf16x2(acc[t][2]); o.w = cvt_relu_bf16x2(acc[t][3]);
            } else {
              o.x = cvt_bf16x2(acc[t][0]); o.y = cvt_bf16x2(acc[t][1]);
              o.z = cvt_bf16x2(acc[t][2]); o.w = cvt_bf16x2(acc[t][3]);
            }
            *reinterpret_cast<uint4*>(op + (size_t)t * p.C) = o;
          }
        }
      }
    }
    __syncthreads();  // everyone is done reading this stage
    if (tid == 0) {
      const long long next = sp + 2 * sp_step;
      if (next < spatial_tiles) issue(next, stage);
    }
    stage ^= 1;
    tx += step_tx;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += step_ty;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
    b += step_b;
  }
}

// 4-D NHWC tensor map {C, W, H, B}, box {cv*8, twi, thi, 1}, no swizzle, zero OOB fill.
inline bool make_tmap_nhwc(tc::EncodeTiledFn fn, CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int cv, int twi, int thi) {
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)(cv * 8), (cuuint32_t)twi, (cuuint32_t)thi, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace dw
}  // namespace spef
