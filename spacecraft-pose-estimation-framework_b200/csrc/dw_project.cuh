// Depthwise 3x3 (stride 1, or 2 for the one stride-2 block that needs it; BN folded, ReLU) fused into the project 1x1 conv of an InvertedResidual block:
//   y = [x +] Wp * relu(dw3x3(h) + bd) + bp        (reference: src/modeling/common/pytorch_layers.py:82-98)
// for the wide blocks (hidden width 576 / 960) whose weights do not fit next to the tiles of the single-kernel block
// (fused_block_t.cuh).  Per-layer kernels move the depthwise output through HBM twice (write + read, 2 x 106 MB at
// 576 ch @ 15x24, batch 256); here it is born in shared memory as the A operand of the project GEMM:
//
//   warp 0        TMA: input box {64 channels, W+2, TH+2} of the hidden tensor per (tile, K chunk); the 1-pixel halo and the
//                 image border are TMA out-of-bounds zero fill (box origin at x = -1, y = y0 - 1)
//   warp 3        TMA: the [N x 64] chunk of the project weights of the same K chunk (L2-resident, 128B-swizzled) into a ring of its
//                 own: with the weights in the A stages (first version) their load could only be issued once the MMA of
//                 ab_stages chunks earlier had completed, and every role ended up waiting for that round trip (ncu: producers
//                 16 probes per chunk on the stage-empty barrier, the MMA warp 54 on stage-full, tensor pipe 6 % busy)
//   warps 4-15    depthwise producers (producer_loop below): G groups on alternate K chunks, thread = (4 channels, 2 output
//                 columns, R output rows)
//   warp 1        tcgen05.mma (M = 128 pixels, N = Cout or two halves of it, K = 16 x 4 per chunk), FP32 accumulators in TMEM
//   warps 16-19   epilogue: tcgen05.ld -> + bias (+ skip input) -> BF16 -> 16-byte global stores
//
// A tile is TH full-width rows of one image (5 x 24 = 120 or 8 x 12 = 96 pixels): rows of a tile are contiguous in NHWC, so
// pixel p of tile t is pixel t * n_px + p of the tensor.  Same rounding points and FP32 operation order as the per-layer
// kernels (depthwise output rounded once to BF16, GEMM K chunks of 64 in order): results are bit-identical to them.
#pragma once
#include "dwconv_tma.cuh"
#include "gemm_tcgen05_v2.cuh"
#include "fused_block_t.cuh"   // reg_inc / reg_dec (setmaxnreg)

namespace spef {
namespace dwp {

constexpr int PROD_WARPS = 12;
constexpr int EPI_WARPS = 4;
constexpr int NT = 128 + 32 * PROD_WARPS + 32 * EPI_WARPS;
constexpr int MAX_IN = 4, MAX_AB = 4, MAX_W = 8, MAX_ACC = 2;
constexpr int A_BYTES = 128 * 128;          // one K chunk of the A operand: 128 pixels x 64 channels
constexpr int WDW_CHUNK_FLOATS = 10 * 64;   // per K chunk: 9 taps + bias, 64 channels each

struct DwpParams {
  int B, H, W, C, N;            // OUTPUT map (= the hidden map when S = 1), hidden channels (GEMM K), project outputs
  int S;                        // depthwise stride 1 | 2 (the hidden map is S*H - (S-1)... x S*W: see the host plan)
  int WB;                       // pixels per row of the input box: W + 2 (S = 1) | 2 W + 1 (S = 2); box rows: TH + 2 | 2 TH + 1
  int TH, tiles_y, n_px;        // tile = TH full-width rows; n_px = TH * W <= 128
  int k_chunks;                 // C / 64
  int pdl_early;                // trigger the dependent launch at once (common.cuh)
  int in_period;                // own items after which a producer group meets the same input stage again: in_stages / gcd(in_stages, G)
  int in_stages, ab_stages, w_stages, acc_stages, acc_stride;   // ab_stages: A-operand stages; w_stages: project-weight chunk stages
  int n_half, nh;               // N = n_half * nh: one MMA per half (nh <= 256, multiple of 16)
  int R, G;                     // producer task: R output rows x 2 columns x 4 channels; G producer groups take alternate K chunks
  int in_bytes, in_stride;      // TMA box bytes (TH+2)(W+2)*128 and the 1024-aligned stage pitch
  const float* wdw;             // depthwise weights + bias per K chunk [k_chunks][10][64]
  const float* bias;            // project bias [N]
  const bf16* residual;         // block input [B,H,W,N] or nullptr
  bf16* out;                    // [B,H,W,N]
};

__host__ __device__ inline int w_stride(const DwpParams& p) { return ((p.N * 128 + 1023) / 1024) * 1024; }
inline size_t smem_bytes(const DwpParams& p) {
  return 1024 + (size_t)p.ab_stages * A_BYTES + (size_t)p.w_stages * w_stride(p) + (size_t)p.in_stages * p.in_stride +
         (size_t)p.k_chunks * WDW_CHUNK_FLOATS * 4 + (size_t)p.N * 4 + 512;
}


struct ProdCtx {
  uint32_t in_u, ab_u, wdw_u, in_full, in_empty, ab_full, ab_empty;
  int num_tiles;
};

// IN_FULL.  A parity wait cannot tell phase n from phase n + 2.  With ONE in_full barrier per input stage, three stages and two producer
// groups, a group waited for item k on a stage whose previous item k - 3 belonged to the OTHER group; TMA loads complete out of order
// (an L2 hit overtakes a DRAM miss), so with fast producers (the small stride-2 tiles at batch 256, hidden tensor larger than the L2)
// the wait could pass before item k - 3 had even landed: the group read a stale stage and released it early (sporadic wrong tiles,
// then a fault; found when the stride-2 plan was added, latent in the 960-channel plan with its three input stages).  Waiting for the
// previous phase first is no fix (when the item has already landed that wait is for the NEXT phase: deadlock, tried).  Each group
// therefore has its OWN in_full barrier per stage -- the TMA warp arms the barrier of the group the item goes to -- so a group sees
// every completion of the barriers it waits on; its phase flips every in_period = in_stages / gcd(in_stages, G) own items (the period
// after which it meets the same stage again).  The A / weight / accumulator rings are safe: their consumer sees every completion, or
// the previous phase was awaited by the writer itself.
// Depthwise producers.  The PROD_WARPS warps form G groups of T = 384 / G threads; group g takes the (tile, K chunk) items
// g, g + G, ... of the CTA's sequence, so consecutive chunks are computed concurrently by different warps and every thread of
// every warp has the same amount of work (first version: 2-row x 4-pixel tasks -- on the 5-row tiles a quarter of the producer
// warps had no task and another quarter half a task, and the busy ones ran at 1 instruction per 10 cycles).
// thread = (4-channel group c4, pair of output columns, R output rows): walks down its R + 2 input rows; per row 4 LDS.64 (a
// half-warp reads one complete 128-byte pixel: conflict-free), bf16 -> f32 by shift / mask, packed FFMA2 into the (at most three)
// output rows the input row feeds; every output accumulates bias, then its taps in row-major order (the order of the per-layer
// kernel: bit-identical sums).  An output row is converted (cvt.rn.relu.bf16x2) and stored as soon as its last input row is
// done -- 8 bytes per pixel into the K-major SWIZZLE_128B A stage -- so only three rows of accumulators are live.
template <int R>
__device__ __forceinline__ void producer_loop(const DwpParams& p, const ProdCtx& c) {
  const int t_all = (int)threadIdx.x - 128;
  const int T = (32 * PROD_WARPS) / p.G;
  const int g = t_all / T, t = t_all - g * T;
  const int lane = threadIdx.x & 31;
  const int c4 = t & 15, task = t >> 4;
  const int nxs = p.W >> 1;
  const int rb = task / nxs, xs = task - rb * nxs;
  const int r0 = rb * R, x0 = xs * 2;
  const uint32_t row_pitch = (uint32_t)(p.WB * 128);
  const uint32_t in_off = (uint32_t)((r0 * p.WB + x0) * 128 + c4 * 8);
  const uint32_t a_sw = (uint32_t)(c4 >> 1), a_lo = (uint32_t)((c4 & 1) * 8);
  const int arow0 = r0 * p.W + x0;
  const uint32_t wdw_u = c.wdw_u + (uint32_t)(c4 * 16);
  int is = g % p.in_stages, as = g % p.ab_stages;
  uint32_t ph_in = 0, ph_ab = (uint32_t)((g / p.ab_stages) & 1);
  int pn = 0;                       // own items since the group's phase of its in_full barriers last flipped (see IN_FULL below)
  const uint32_t in_full_g = c.in_full + 8u * (uint32_t)(g * MAX_IN);
  int kc = g;                       // K chunk of this group's current item (g < k_chunks: checked by the host)
  for (int tile = blockIdx.x; tile < c.num_tiles;) {
    tc::mbar_wait(in_full_g + 8u * (uint32_t)is, ph_in);
    tc::mbar_wait(c.ab_empty + 8u * (uint32_t)as, ph_ab ^ 1);
    const uint32_t wb = wdw_u + (uint32_t)(kc * WDW_CHUNK_FLOATS * 4);
    uint64_t w[9][2], bv[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w0 = tc::lds_f4(wb + (uint32_t)k * 256u);
      w[k][0] = f32x2(w0.x, w0.y); w[k][1] = f32x2(w0.z, w0.w);
    }
    {
      const float4 b0 = tc::lds_f4(wb + 9u * 256u);
      bv[0] = f32x2(b0.x, b0.y); bv[1] = f32x2(b0.z, b0.w);
    }
    const uint32_t tile_u = c.in_u + (uint32_t)is * (uint32_t)p.in_stride + in_off;
    const uint32_t sa = c.ab_u + (uint32_t)as * (uint32_t)A_BYTES;
    uint64_t acc[R][2][2];
#pragma unroll
    for (int i = 0; i < R + 2; ++i) {
      if (i < R) {
#pragma unroll
        for (int o = 0; o < 2; ++o) { acc[i][o][0] = bv[0]; acc[i][o][1] = bv[1]; }
      }
      const uint32_t row_u = tile_u + (uint32_t)i * row_pitch;
      uint64_t v[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t ux, uy;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ux), "=r"(uy) : "r"(row_u + (uint32_t)(j * 128)));
        // bf16 pair -> packed f32 pair: low half << 16, high half masked
        v[j][0] = f32x2(__uint_as_float(ux << 16), __uint_as_float(ux & 0xffff0000u));
        v[j][1] = f32x2(__uint_as_float(uy << 16), __uint_as_float(uy & 0xffff0000u));
      }
      if (i == R + 1) {   // last input row is in registers: hand the input stage back before the remaining FMAs and stores
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(c.in_empty + 8u * (uint32_t)is);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ky = i - r;
        if (ky >= 0 && ky <= 2) {
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              acc[r][o][0] = fma_f32x2(v[o + kx][0], w[ky * 3 + kx][0], acc[r][o][0]);
              acc[r][o][1] = fma_f32x2(v[o + kx][1], w[ky * 3 + kx][1], acc[r][o][1]);
            }
        }
      }
      if (i >= 2) {   // output row i - 2 has all its taps
        const int r = i - 2;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const uint32_t row = (uint32_t)(arow0 + r * p.W + o);
          const uint32_t off = row * 128u + ((a_sw ^ (row & 7u)) << 4) + a_lo;
          asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa + off), "r"(dw::cvt_relu_bf16x2(acc[r][o][0])), "r"(dw::cvt_relu_bf16x2(acc[r][o][1])) : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncwarp();                                                  // every lane has read its input pixels and stored its outputs
    if (lane == 0) tc::mbar_arrive(c.ab_full + 8u * (uint32_t)as);
    is += p.G;
    while (is >= p.in_stages) is -= p.in_stages;
    if (++pn == p.in_period) { pn = 0; ph_in ^= 1; }
    as += p.G;
    while (as >= p.ab_stages) { as -= p.ab_stages; ph_ab ^= 1; }
    kc += p.G;
    while (kc >= p.k_chunks) { kc -= p.k_chunks; tile += gridDim.x; }
  }
}

// Stride-2 producers (the one stride-2 block without a single-kernel plan: 576 channels, 15x24 -> 8x12).  Same task shape and the same
// hand-offs as above; box row 0 is hidden row 2 oy0 - 1 and box column 0 hidden column -1 (TMA zero fill = the conv's padding), so
// output (r, x) reads box rows 2r .. 2r + 2 and box columns 2x .. 2x + 2: a thread walks down 2R + 1 input rows of five pixels; an
// even row feeds two output rows (ky = 0 of row r, ky = 2 of row r - 1), an odd one only ky = 1 of row r.  Bias first, taps in
// row-major order like the per-layer kernel: bit-identical sums.
template <int R>
__device__ __forceinline__ void producer_loop_s2(const DwpParams& p, const ProdCtx& c) {
  const int t_all = (int)threadIdx.x - 128;
  const int T = (32 * PROD_WARPS) / p.G;
  const int g = t_all / T, t = t_all - g * T;
  const int lane = threadIdx.x & 31;
  const int c4 = t & 15, task = t >> 4;
  const int nxs = p.W >> 1;
  const int rb = task / nxs, xs = task - rb * nxs;
  const int r0 = rb * R, x0 = xs * 2;
  const uint32_t row_pitch = (uint32_t)(p.WB * 128);
  const uint32_t in_off = (uint32_t)((2 * r0 * p.WB + 2 * x0) * 128 + c4 * 8);
  const uint32_t a_sw = (uint32_t)(c4 >> 1), a_lo = (uint32_t)((c4 & 1) * 8);
  const int arow0 = r0 * p.W + x0;
  const uint32_t wdw_u = c.wdw_u + (uint32_t)(c4 * 16);
  int is = g % p.in_stages, as = g % p.ab_stages;
  uint32_t ph_in = 0, ph_ab = (uint32_t)((g / p.ab_stages) & 1);
  int pn = 0;                       // own items since the group's phase of its in_full barriers last flipped (see IN_FULL below)
  const uint32_t in_full_g = c.in_full + 8u * (uint32_t)(g * MAX_IN);
  int kc = g;
  for (int tile = blockIdx.x; tile < c.num_tiles;) {
    tc::mbar_wait(in_full_g + 8u * (uint32_t)is, ph_in);
    tc::mbar_wait(c.ab_empty + 8u * (uint32_t)as, ph_ab ^ 1);
    const uint32_t wb = wdw_u + (uint32_t)(kc * WDW_CHUNK_FLOATS * 4);
    uint64_t w[9][2], bv[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w0 = tc::lds_f4(wb + (uint32_t)k * 256u);
      w[k][0] = f32x2(w0.x, w0.y); w[k][1] = f32x2(w0.z, w0.w);
    }
    {
      const float4 b0 = tc::lds_f4(wb + 9u * 256u);
      bv[0] = f32x2(b0.x, b0.y); bv[1] = f32x2(b0.z, b0.w);
    }
    const uint32_t tile_u = c.in_u + (uint32_t)is * (uint32_t)p.in_stride + in_off;
    const uint32_t sa = c.ab_u + (uint32_t)as * (uint32_t)A_BYTES;
    uint64_t acc[R][2][2];
#pragma unroll
    for (int i = 0; i < 2 * R + 1; ++i) {
      if ((i & 1) == 0 && (i >> 1) < R) {
#pragma unroll
        for (int o = 0; o < 2; ++o) { acc[i >> 1][o][0] = bv[0]; acc[i >> 1][o][1] = bv[1]; }
      }
      const uint32_t row_u = tile_u + (uint32_t)i * row_pitch;
      uint64_t v[5][2];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        uint32_t ux, uy;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ux), "=r"(uy) : "r"(row_u + (uint32_t)(j * 128)));
        v[j][0] = f32x2(__uint_as_float(ux << 16), __uint_as_float(ux & 0xffff0000u));
        v[j][1] = f32x2(__uint_as_float(uy << 16), __uint_as_float(uy & 0xffff0000u));
      }
      if (i == 2 * R) {   // last input row is in registers: hand the input stage back before the remaining FMAs and stores
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(c.in_empty + 8u * (uint32_t)is);
      }
      // ky = 2 of the row above first (its last taps), then ky = 0 / 1 of the row this input row starts or continues
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ky = i - 2 * r;
        if (ky >= 0 && ky <= 2) {
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              acc[r][o][0] = fma_f32x2(v[2 * o + kx][0], w[ky * 3 + kx][0], acc[r][o][0]);
              acc[r][o][1] = fma_f32x2(v[2 * o + kx][1], w[ky * 3 + kx][1], acc[r][o][1]);
            }
        }
      }
      if (i >= 2 && (i & 1) == 0) {   // output row i / 2 - 1 has all its taps
        const int r = (i >> 1) - 1;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const uint32_t row = (uint32_t)(arow0 + r * p.W + o);
          const uint32_t off = row * 128u + ((a_sw ^ (row & 7u)) << 4) + a_lo;
          asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa + off), "r"(dw::cvt_relu_bf16x2(acc[r][o][0])), "r"(dw::cvt_relu_bf16x2(acc[r][o][1])) : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(c.ab_full + 8u * (uint32_t)as);
    is += p.G;
    while (is >= p.in_stages) is -= p.in_stages;
    if (++pn == p.in_period) { pn = 0; ph_in ^= 1; }
    as += p.G;
    while (as >= p.ab_stages) { as -= p.ab_stages; ph_ab ^= 1; }
    kc += p.G;
    while (kc >= p.k_chunks) { kc -= p.k_chunks; tile += gridDim.x; }
  }
}

__global__ void __launch_bounds__(NT, 1)
dw_project_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const DwpParams p) {
  if (p.pdl_early) pdl_launch_dependents();   // the next kernel of the chain may start its own set-up now (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int abs_ = A_BYTES, ws_ = w_stride(p);
  uint8_t* ab_s = smem;                                            // [ab_stages][A 16 KB]
  uint8_t* w_s = ab_s + (size_t)p.ab_stages * abs_;                // [w_stages][Wp chunk N x 128 B]
  uint8_t* in_s = w_s + (size_t)p.w_stages * ws_;                  // [in_stages][(TH+2)(W+2) pixels x 128 B]
  float* wdw_s = reinterpret_cast<float*>(in_s + (size_t)p.in_stages * p.in_stride);
  float* bias_s = wdw_s + (size_t)p.k_chunks * WDW_CHUNK_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + p.N);
  uint64_t* in_full = bars;                 // [2][MAX_IN] TMA -> producers, one set per producer group (IN_FULL below)
  uint64_t* in_empty = in_full + 2 * MAX_IN; // [MAX_IN]   producers -> TMA
  uint64_t* ab_full = in_empty + MAX_IN;    // [MAX_AB]   producers (A) -> MMA
  uint64_t* ab_empty = ab_full + MAX_AB;    // [MAX_AB]   MMA -> producers
  uint64_t* w_full = ab_empty + MAX_AB;     // [MAX_W]    TMA (Wp chunk) -> MMA
  uint64_t* w_empty = w_full + MAX_W;       // [MAX_W]    MMA -> weight loader
  uint64_t* acc_full = w_empty + MAX_W;     // [MAX_ACC]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + MAX_ACC; // [MAX_ACC]  epilogue -> MMA
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_empty + MAX_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.B * p.tiles_y;

  for (int i = threadIdx.x; i < p.k_chunks * WDW_CHUNK_FLOATS; i += NT) wdw_s[i] = p.wdw[i];
  for (int i = threadIdx.x; i < p.N; i += NT) bias_s[i] = p.bias[i];
  // rows >= n_px of every A stage are never written by the producers: zero them once (their accumulator rows are not stored)
  {
    const int tail16 = (128 - p.n_px) * 8;   // 16-byte pieces
    for (int s = 0; s < p.ab_stages; ++s)
      for (int i = threadIdx.x; i < tail16; i += NT)
        *reinterpret_cast<uint4*>(ab_s + (size_t)s * abs_ + (size_t)p.n_px * 128 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_IN; ++i) {
      tc::mbar_init(tc::smem_u32(&in_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&in_full[MAX_IN + i]), 1);
      tc::mbar_init(tc::smem_u32(&in_empty[i]), PROD_WARPS / p.G);
    }
    for (int i = 0; i < MAX_AB; ++i) {
      tc::mbar_init(tc::smem_u32(&ab_full[i]), PROD_WARPS / p.G); // one arrive per producer warp of the group that owns the chunk
      tc::mbar_init(tc::smem_u32(&ab_empty[i]), 1);
    }
    for (int i = 0; i < MAX_W; ++i) {
      tc::mbar_init(tc::smem_u32(&w_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&w_empty[i]), 1);
    }
    for (int i = 0; i < MAX_ACC; ++i) {
      tc::mbar_init(tc::smem_u32(&acc_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[i]), EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_ptr_s)), "r"(tc::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();   // everything above is independent of the previous kernel's output (common.cuh)

  // register file per role (launch allocation 20 warps x 96): producers 12 x 120, epilogue 4 x 64, the warpgroup of the single-lane
  // roles 4 x 40 (setmaxnreg is a warpgroup instruction: all four warps execute the same one)
  if (warp < 4) fbt::reg_dec<40>();
  if (warp == 0) {
    // ===================== TMA: hidden-tensor boxes =====================
    if (lane == 0) {
      int is = 0, grp = 0;            // grp: the producer group that takes this item (items alternate between the G groups)
      uint32_t ph = 0;
      int b = (int)blockIdx.x / p.tiles_y, ty = (int)blockIdx.x % p.tiles_y;
      const int db = (int)gridDim.x / p.tiles_y, dty = (int)gridDim.x % p.tiles_y;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          tc::mbar_wait(tc::smem_u32(&in_empty[is]), ph ^ 1);
          const uint32_t fb = tc::smem_u32(&in_full[grp * MAX_IN + is]);
          tc::mbar_arrive_expect_tx(fb, (uint32_t)p.in_bytes);
          dw::tma_load_4d(tc::smem_u32(in_s + (size_t)is * p.in_stride), &tmX, kc * 64, -1, ty * p.TH * p.S - 1, b, fb);
          if (++is == p.in_stages) { is = 0; ph ^= 1; }
          if (++grp == p.G) grp = 0;
        }
        b += db; ty += dty;
        if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== TMA: project-weight chunks =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          tc::mbar_wait(tc::smem_u32(&w_empty[ws]), ph ^ 1);
          const uint32_t fb = tc::smem_u32(&w_full[ws]);
          tc::mbar_arrive_expect_tx(fb, (uint32_t)(p.N * 128));
          const uint32_t dst = tc::smem_u32(w_s + (size_t)ws * ws_);
          for (int h = 0; h < p.n_half; ++h) tc::tma_load_2d(dst + (uint32_t)(h * p.nh * 128), &tmW, kc * 64, h * p.nh, fb);
          if (++ws == p.w_stages) { ws = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, p.nh);
    const uint64_t a_base = tc::make_smem_desc_sw128(tc::smem_u32(ab_s));
    const uint64_t b_base = tc::make_smem_desc_sw128(tc::smem_u32(w_s));
    const uint32_t s_step = (uint32_t)abs_ >> 4, w_step = (uint32_t)ws_ >> 4, h_step = (uint32_t)(p.nh * 128) >> 4;
    int as = 0, ws = 0, acc = 0;
    uint32_t ph = 0, w_ph = 0, acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      tc::mbar_wait(tc::smem_u32(&acc_empty[acc]), acc_ph ^ 1);
      tc::tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        tc::mbar_wait(tc::smem_u32(&w_full[ws]), w_ph);
        tc::mbar_wait(tc::smem_u32(&ab_full[as]), ph);
        tc::tcgen05_fence_after();
        const uint64_t a_desc = a_base + (uint64_t)((uint32_t)as * s_step);
        const uint64_t b_desc = b_base + (uint64_t)((uint32_t)ws * w_step);
        for (int h = 0; h < p.n_half; ++h)
          for (uint32_t k = 0; k < 4; ++k)
            tc::mma_elect_v2(d_tmem + (uint32_t)(h * p.nh), a_desc + (uint64_t)(k * 2u), b_desc + (uint64_t)((uint32_t)h * h_step + k * 2u), idesc,
                             (kc > 0 || k > 0) ? 1u : 0u);
        tc::commit_elect_v2(tc::smem_u32(&ab_empty[as]));
        tc::commit_elect_v2(tc::smem_u32(&w_empty[ws]));
        if (++as == p.ab_stages) { as = 0; ph ^= 1; }
        if (++ws == p.w_stages) { ws = 0; w_ph ^= 1; }
      }
      tc::commit_elect_v2(tc::smem_u32(&acc_full[acc]));
      if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1; }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + PROD_WARPS) {
    // ===================== depthwise producers =====================
    fbt::reg_inc<120>();
    const ProdCtx pc{tc::smem_u32(in_s), tc::smem_u32(ab_s), tc::smem_u32(wdw_s), tc::smem_u32(in_full), tc::smem_u32(in_empty),
                     tc::smem_u32(ab_full), tc::smem_u32(ab_empty), num_tiles};
    if (p.S == 2) {
      switch (p.R) {
        case 4: producer_loop_s2<4>(p, pc); break;
        case 2: producer_loop_s2<2>(p, pc); break;
        default: producer_loop_s2<1>(p, pc); break;
      }
    } else switch (p.R) {
      case 5: producer_loop<5>(p, pc); break;
      case 4: producer_loop<4>(p, pc); break;
      case 3: producer_loop<3>(p, pc); break;
      case 2: producer_loop<2>(p, pc); break;
      default: producer_loop<1>(p, pc); break;
    }
  } else if (warp >= 4 + PROD_WARPS) {
    fbt::reg_dec<64>();
    // ===================== epilogue =====================
    const int qd = warp & 3;                            // TMEM lane quarter this warp may read
    const int row = qd * 32 + lane;
    const bool valid = row < p.n_px;
    const uint32_t bias_u = tc::smem_u32(bias_s);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const size_t pix = (size_t)tile * p.n_px + row;
      bf16* op = p.out + pix * p.N;
      const bf16* rp = p.residual ? p.residual + pix * p.N : nullptr;
      tc::mbar_wait_relaxed(tc::smem_u32(&acc_full[acc]), acc_ph, 64);
      tc::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * p.acc_stride);
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (c0 + 32 >= p.N) {
          // last tcgen05.ld of this warp on the accumulator has completed -> hand the TMEM stage back
          tc::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[acc]));
        }
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b0 = tc::lds_f4(bias_u + (uint32_t)(c0 + g * 8) * 4u), b1 = tc::lds_f4(bias_u + (uint32_t)(c0 + g * 8) * 4u + 16u);
            if (rp) {
              // linear bottleneck with skip connection: y = x + (acc + bias), one rounding (pytorch_layers.py:93-98)
              float f[8] = {__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y, __uint_as_float(v[g * 8 + 2]) + b0.z,
                            __uint_as_float(v[g * 8 + 3]) + b0.w, __uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y,
                            __uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w};
              float r[8];
              Vec8<bf16>::load(rp + c0 + g * 8, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += r[e];
              *reinterpret_cast<uint4*>(op + c0 + g * 8) = Vec8<bf16>::pack(f);
            } else {
              uint4 ov;
              ov.x = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[g * 8 + 0], v[g * 8 + 1]), tc::pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y))));
              ov.y = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[g * 8 + 2], v[g * 8 + 3]), tc::pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w))));
              ov.z = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[g * 8 + 4], v[g * 8 + 5]), tc::pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y))));
              ov.w = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[g * 8 + 6], v[g * 8 + 7]), tc::pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w))));
              *reinterpret_cast<uint4*>(op + c0 + g * 8) = ov;
            }
          }
        }
      }
      if (++acc == p.acc_stages) { acc = 0; acc_ph ^= 1; }
    }
  }

  // ---- teardown ----
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tc::TMEM_COLS) : "memory");
  }
}

}  // namespace dwp
}  // namespace spef
