// Fused InvertedResidual block (reference: src/modeling/common/pytorch_layers.py:65-98):
//     y = [x +] project_1x1( relu( dw3x3_s( relu( expand_1x1(x) ) ) ) )          BN folded, BF16 activations
// in ONE persistent kernel, so the 6x hidden tensor never leaves the SM: HBM traffic per block drops from
// (Cin + 2*Ch*(1 + 1/S^2) + Cout) to (Cin*halo + Cout) elements per pixel.
//
// Per CTA (one per SM), per output tile of TH x TW pixels (<= 128 = one project M tile):
//   TMA      x tile with 1-px halo, 4-D box {64 ch, TWI, THI, 1 image}, SWIZZLE_128B -> directly the K-major UMMA
//            A operand (row = box pixel); image border = TMA out-of-bounds zero fill
//   MMA      expand: D_e[box px (<=256 = 2 M tiles), 64 hidden ch] = X * We_chunk^T       (tcgen05, TMEM)
//   workers  drain  : tcgen05.ld -> +bias -> ReLU -> zero the pixels outside the image (dw zero padding applies to the
//                     *hidden* tensor) -> bf16 -> smem hidden tile Hs (128 B per pixel, 16-byte chunks XOR-swizzled)
//            dw     : 3x3 stride-S window over Hs in FP32 (sliding register window), +bias, ReLU -> bf16 -> smem A2 in
//                     the K-major SWIZZLE_128B UMMA layout (row = output pixel)
//   MMA      project: D_p[128 px, Cout] += A2 * Wp_chunk^T, accumulated in TMEM over the hidden chunks
//   epilogue tcgen05.ld -> +bias (+ residual x) -> bf16 -> global
// Hidden channels are processed in chunks of 64; two worker groups alternate chunks (each with its own TMEM expand
// stage, Hs and A2 buffer), so the drain of chunk c+1 overlaps the depthwise of chunk c and the MMAs of both.
// Weights (We chunk, Wp chunk, dw weights + biases) travel through a ring of smem stages; when the whole block fits
// in the ring it is loaded once per CTA and stays resident.
// Rounding points are identical to the unfused kernels (expand out, dw out and block out rounded to BF16 once each).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_v2.cuh"
#include "dwconv_tma.cuh"

namespace spef {
namespace fb {

constexpr int HC = 64;                       // hidden channels per chunk
constexpr int MAX_NG = 2;                    // worker groups (template parameter NG = 1 | 2)
constexpr int AUX_FLOATS = 11 * HC;          // per chunk: expand bias[64] | dw bias[64] | dw weights[9][64]
constexpr int AUX_BYTES = AUX_FLOATS * 4;    // 2816
constexpr int AUX_STRIDE = 3072;
constexpr int MT_BYTES = 128 * 128;          // one 128-row K-major tile of 64 bf16
constexpr int MAX_W_STAGES = 8;
constexpr int CTRL_WARPS = 8;                // 4 epilogue warps + spare, TMEM allocator, TMA producer, MMA issuer
// Warp roles, lowest warp ids first: NG*GW workers | 4 epilogue | TMEM alloc | TMA producer | project MMA issuer | expand
// MMA issuer.  What the clock64 traces showed about the issuer roles: a single thread retires ~1 instruction per 10 cycles
// (dependent uniform-datapath ops), so every instruction between two tcgen05.mma counts -- descriptors are precomputed, the
// warp stays converged (elect.sync inside the asm instead of an `if (lane == 0)` region, which costs an ELECT / R2UR /
// BRA.U.ANY wrapper per MMA), and expand and project are issued by two different warps so that neither queues behind the
// other's mbarrier waits.
constexpr int MAX_ACC = 3;                   // TMEM expand stages: stage s, M tile mt at column s*128 + mt*64; project columns follow

struct FbParams {
  const bf16* x;       // block input  [B,H,W,Cin]  (residual source)
  bf16* y;             // block output [B,Ho,Wo,Cout]
  const float* aux;    // [n_chunks][AUX_FLOATS]
  const float* bp;     // project bias [cpad]
  int B, H, W, Cin, Ch, Cout, Ho, Wo;
  int TH, TW, THI, TWI, tiles_y, tiles_x;
  int kc_in;           // 64-channel K chunks of the expand GEMM
  int n_chunks;        // hidden chunks
  int cpad;            // project MMA N (Cout rounded up to 16)
  int x_stages, w_stages, resident, proj_stages;
  int n_acc;           // TMEM expand accumulator stages (2 | 3), round-robin over the work items
  int proj_col0, proj_stride;   // TMEM columns of the project accumulator stages
  int residual;
  int has_expand;      // 0: t = 1 block (no expand conv): the TMA box *is* the hidden tile
  int debug_skip;      // debug (timing experiments only, wrong results): bit 0 skip the drain body, bit 1 skip the depthwise body
  long long* trace;    // debug: clock64 timestamps of CTA 0, [work item n < 64][16] (nullptr in production)
};

__host__ __device__ inline int w_stage_bytes(int kc_in, int cpad) { return kc_in * (HC * 128) + cpad * 128 + AUX_STRIDE; }
__host__ __device__ inline int x_stage_bytes(int kc_in) { return kc_in * 2 * MT_BYTES; }
inline size_t smem_bytes(const FbParams& p, int ng) {
  return 1024 + (size_t)p.x_stages * x_stage_bytes(p.kc_in) + (size_t)p.w_stages * w_stage_bytes(p.kc_in, p.cpad) +
         (size_t)ng * 2 * MT_BYTES + (size_t)ng * MT_BYTES + 2048 /*bias*/ + 512 /*barriers*/;
}

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
// tcgen05.mma / tcgen05.commit issued by the elected lane of a converged warp (always the same lane: commit tracks the MMAs
// of the executing thread)
__device__ __forceinline__ void mma_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void group_sync(int g, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(nthreads) : "memory");
}

// exact floor(t / d) for 0 <= t < 2^22 via a float reciprocal and one fix-up step (integer division costs ~150 cycles of
// dependent latency, which matters in the single-thread TMA / MMA roles)
__device__ __forceinline__ int fast_div(int t, int d, float rcp) {
  int q = __float2int_rd(((float)t + 0.5f) * rcp);
  if (q * d > t) --q;
  else if ((q + 1) * d <= t) ++q;
  return q;
}

// Position in this CTA's sequence of (tile, chunk) work items, advanced without divisions.  Stage indices and mbarrier
// parities of every ring the item touches are carried along.
struct WorkIt {
  int n, i, c;        // item, tile iteration, chunk
  int g, kph;         // worker group and parity of the group's item counter
  int kph2;           // parity of (item counter >> 1): barrier parity when the group owns two A2 buffers (buffer = kph)
  int xs, xph;        // x stage / parity
  int ws, wph;        // weight stage / parity
  int ps, pph;        // project accumulator stage / parity
  int as, aph;        // TMEM expand stage / parity
};
__device__ __forceinline__ WorkIt work_begin() { return WorkIt{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; }
template <int NG>
__device__ __forceinline__ void work_next(WorkIt& w, const FbParams& p) {
  const int n_chunks = p.n_chunks, x_stages = p.x_stages, w_stages = p.w_stages, resident = p.resident, proj_stages = p.proj_stages;
  ++w.n;
  if (++w.g == NG) { w.g = 0; w.kph ^= 1; if (w.kph == 0) w.kph2 ^= 1; }
  if (++w.as == p.n_acc) { w.as = 0; w.aph ^= 1; }
  if (++w.c == n_chunks) {
    w.c = 0; ++w.i;
    if (++w.xs == x_stages) { w.xs = 0; w.xph ^= 1; }
    if (++w.ps == proj_stages) { w.ps = 0; w.pph ^= 1; }
  }
  if (resident) { w.ws = w.c; }
  else if (++w.ws == w_stages) { w.ws = 0; w.wph ^= 1; }
}

// S: depthwise stride. NG: worker groups (1 or 2). GW: warps per worker group (4 or 8).
template <int S, int NG, int GW>
__global__ void __launch_bounds__(32 * (CTRL_WARPS + NG * GW), 1)
fused_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                   const __grid_constant__ CUtensorMap tmWp, const FbParams p) {
  constexpr int TX = (S == 1) ? 4 : 2;            // outputs per depthwise item along x
  constexpr int NCOLS = (TX - 1) * S + 3;         // input columns of one item
  constexpr int GT = 32 * GW;                     // threads per worker group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xsb = x_stage_bytes(p.kc_in);
  const int wsb = w_stage_bytes(p.kc_in, p.cpad);
  uint8_t* x_s = smem;
  uint8_t* w_s = x_s + (size_t)p.x_stages * xsb;
  uint8_t* hs_s = w_s + (size_t)p.w_stages * wsb;            // [NG][256 px][128 B]
  uint8_t* a2_s = hs_s + (size_t)NG * 2 * MT_BYTES;          // [NG][128 px][128 B]
  float* bp_s = reinterpret_cast<float*>(a2_s + (size_t)NG * MT_BYTES);   // [<= 512]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bp_s + 512);
  uint64_t* x_full = bars;                        // [2]
  uint64_t* x_empty = x_full + 2;                 // [2]
  uint64_t* w_full = x_empty + 2;                 // [MAX_W_STAGES]
  uint64_t* w_empty = w_full + MAX_W_STAGES;      // [MAX_W_STAGES]
  uint64_t* acc_full = w_empty + MAX_W_STAGES;    // [MAX_ACC]  expand MMA -> drain
  uint64_t* acc_empty = acc_full + MAX_ACC;       // [MAX_ACC]  drain -> expand MMA
  uint64_t* a2_full = acc_empty + MAX_ACC;        // [NG]  dw -> project MMA
  uint64_t* a2_empty = a2_full + MAX_NG;              // [NG]  project MMA -> dw
  uint64_t* proj_full = a2_empty + MAX_NG;            // [2]   project MMA -> epilogue
  uint64_t* proj_empty = proj_full + 2;           // [2]   epilogue -> project MMA
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(proj_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int FIRST_EPI_WARP = NG * GW;
  constexpr int WARP_ALLOC = NG * GW + 4, WARP_TMA = NG * GW + 5, WARP_MMA_P = NG * GW + 6, WARP_MMA = NG * GW + 7;
  const int P_in = p.THI * p.TWI;
  const int n_mt = (P_in + 127) >> 7;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const long long num_tiles = (long long)p.B * tiles_per_img;
  const int my_tiles = (int)((num_tiles - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
  const int total = my_tiles * p.n_chunks;        // (tile, chunk) work items of this CTA
  const int pstride = p.proj_stride;
  const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
#define FB_TRACE(n_, slot_) do { if (tr && (n_) < 64) p.trace[(n_) * 16 + (slot_)] = clock64(); } while (0)

  for (int i = threadIdx.x; i < 512; i += (int)blockDim.x) bp_s[i] = (i < p.cpad) ? p.bp[i] : 0.f;
  if (warp == WARP_TMA && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmWe);
    tc::tma_prefetch_desc(&tmWp);
  }
  if (warp == WARP_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tc::smem_u32(&x_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&x_empty[i]), 1);
      tc::mbar_init(tc::smem_u32(&proj_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&proj_empty[i]), 4);
    }
    for (int i = 0; i < MAX_W_STAGES; ++i) {
      tc::mbar_init(tc::smem_u32(&w_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&w_empty[i]), 1);
    }
    for (int i = 0; i < MAX_ACC; ++i) {
      tc::mbar_init(tc::smem_u32(&acc_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[i]), GW);
    }
    for (int i = 0; i < NG; ++i) {
      tc::mbar_init(tc::smem_u32(&a2_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&a2_empty[i]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_ptr_s)), "r"(tc::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // tile iteration i of this CTA -> (image, tile origin in output pixels); num_tiles < 2^22 (host)
  const float rcp_tpi = 1.0f / (float)tiles_per_img, rcp_tx = 1.0f / (float)p.tiles_x, rcp_twi = 1.0f / (float)p.TWI;
  auto tile_coords = [&](int i, int& b, int& oy0, int& ox0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    b = fast_div(t, tiles_per_img, rcp_tpi);
    const int r = t - b * tiles_per_img;
    const int ty = fast_div(r, p.tiles_x, rcp_tx);
    oy0 = ty * p.TH;
    ox0 = (r - ty * p.tiles_x) * p.TW;
  };

  if (warp == WARP_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t x_tx = (uint32_t)(p.kc_in * P_in * 128);
      const uint32_t w_tx = (uint32_t)((p.has_expand ? p.kc_in * HC * 128 : 0) + p.cpad * 128 + AUX_BYTES);
      for (WorkIt w = work_begin(); w.n < total; work_next<NG>(w, p)) {
        const int i = w.i, c = w.c, xs = w.xs;
        if (c == 0) {
          int b, oy0, ox0;
          tile_coords(i, b, oy0, ox0);
          tc::mbar_wait(tc::smem_u32(&x_empty[xs]), (uint32_t)(w.xph ^ 1));
          const uint32_t fb = tc::smem_u32(&x_full[xs]);
          tc::mbar_arrive_expect_tx(fb, x_tx);
          for (int kc = 0; kc < p.kc_in; ++kc)
            dw::tma_load_4d(tc::smem_u32(x_s + (size_t)xs * xsb + (size_t)kc * 2 * MT_BYTES), &tmX, kc * 64, ox0 * S - 1, oy0 * S - 1, b, fb);
        }
        if (!p.resident || i == 0) {
          const int ws = w.ws;
          if (!p.resident) tc::mbar_wait(tc::smem_u32(&w_empty[ws]), (uint32_t)(w.wph ^ 1));
          const uint32_t fb = tc::smem_u32(&w_full[ws]);
          uint8_t* dst = w_s + (size_t)ws * wsb;
          tc::mbar_arrive_expect_tx(fb, w_tx);
          for (int kc = 0; kc < p.kc_in && p.has_expand; ++kc) tc::tma_load_2d(tc::smem_u32(dst + (size_t)kc * HC * 128), &tmWe, kc * 64, c * HC, fb);
          tc::tma_load_2d(tc::smem_u32(dst + (size_t)p.kc_in * HC * 128), &tmWp, c * HC, 0, fb);
          bulk_load_1d(tc::smem_u32(dst + (size_t)p.kc_in * HC * 128 + (size_t)p.cpad * 128), p.aux + (size_t)c * AUX_FLOATS, AUX_BYTES, fb);
        }
      }
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ===================== expand MMA issuer (converged warp, elected lane issues) =====================
    const uint32_t idesc_e = tc::make_idesc_bf16(128, HC);
    const uint32_t kst_last = (uint32_t)(((p.Cin - (p.kc_in - 1) * 64) + 15) / 16);   // K steps of the last 64-channel chunk
    const uint64_t a_base = tc::make_smem_desc_sw128(tc::smem_u32(x_s));
    const uint64_t b_base = tc::make_smem_desc_sw128(tc::smem_u32(w_s));
    const uint32_t x_step = (uint32_t)xsb >> 4, w_step = (uint32_t)wsb >> 4;
    for (WorkIt w = work_begin(); w.n < total && p.has_expand; work_next<NG>(w, p)) {
      const int n = w.n;
      if (lane == 0) FB_TRACE(n, 0);
      if (w.c == 0) tc::mbar_wait(tc::smem_u32(&x_full[w.xs]), (uint32_t)w.xph);
      tc::mbar_wait(tc::smem_u32(&w_full[w.ws]), (uint32_t)w.wph);
      tc::mbar_wait(tc::smem_u32(&acc_empty[w.as]), (uint32_t)(w.aph ^ 1));
      tc::tcgen05_fence_after();
      if (lane == 0) FB_TRACE(n, 1);
      const uint64_t a0 = a_base + (uint64_t)((uint32_t)w.xs * x_step);
      const uint64_t b0 = b_base + (uint64_t)((uint32_t)w.ws * w_step);
      const uint32_t d0 = tmem_base + (uint32_t)(w.as * 128);
      for (int kc = 0; kc < p.kc_in; ++kc) {
        const uint32_t ksteps = (kc == p.kc_in - 1) ? kst_last : 4u;
        for (uint32_t ks = 0; ks < ksteps; ++ks) {
          const uint64_t a = a0 + (uint64_t)(kc * (2 * MT_BYTES >> 4) + (int)ks * 2);
          const uint64_t b = b0 + (uint64_t)(kc * (HC * 128 >> 4) + (int)ks * 2);
          const uint32_t accum = (kc > 0 || ks > 0) ? 1u : 0u;
          mma_elect(d0, a, b, idesc_e, accum);
          if (n_mt > 1) mma_elect(d0 + 64u, a + (uint64_t)(MT_BYTES >> 4), b, idesc_e, accum);
        }
      }
      if (lane == 0) FB_TRACE(n, 3);
      commit_elect(tc::smem_u32(&acc_full[w.as]));
      if (w.c == p.n_chunks - 1) commit_elect(tc::smem_u32(&x_empty[w.xs]));
      if (lane == 0) FB_TRACE(n, 2);
    }
  } else if (warp == WARP_MMA_P) {
    // ===================== project MMA issuer =====================
    const uint32_t idesc_p = tc::make_idesc_bf16(128, p.cpad);
    const uint64_t a_base = tc::make_smem_desc_sw128(tc::smem_u32(a2_s));
    const uint64_t b_base = tc::make_smem_desc_sw128(tc::smem_u32(w_s + (size_t)p.kc_in * HC * 128));
    const uint32_t w_step = (uint32_t)wsb >> 4;
    const uint32_t kst_tail = (uint32_t)((p.Ch - (p.n_chunks - 1) * HC + 15) / 16);  // never read A2 columns the depthwise did not write
    for (WorkIt w = work_begin(); w.n < total; work_next<NG>(w, p)) {
      const int n = w.n;
      if (w.c == 0) tc::mbar_wait(tc::smem_u32(&proj_empty[w.ps]), (uint32_t)(w.pph ^ 1));
      tc::mbar_wait(tc::smem_u32(&a2_full[w.g]), (uint32_t)w.kph);
      tc::tcgen05_fence_after();
      if (lane == 0) FB_TRACE(n, 4);
      const uint64_t a0 = a_base + (uint64_t)((uint32_t)w.g * (uint32_t)(MT_BYTES >> 4));
      const uint64_t b0 = b_base + (uint64_t)((uint32_t)w.ws * w_step);
      const uint32_t d = tmem_base + (uint32_t)(p.proj_col0 + w.ps * pstride);
      const uint32_t ksteps = (w.c == p.n_chunks - 1) ? kst_tail : 4u;
      for (uint32_t ks = 0; ks < ksteps; ++ks)
        mma_elect(d, a0 + (uint64_t)(ks * 2), b0 + (uint64_t)(ks * 2), idesc_p, (w.c > 0 || ks > 0) ? 1u : 0u);
      commit_elect(tc::smem_u32(&a2_empty[w.g]));
      if (!p.resident) commit_elect(tc::smem_u32(&w_empty[w.ws]));
      if (w.c == p.n_chunks - 1) commit_elect(tc::smem_u32(&proj_full[w.ps]));
      if (lane == 0) FB_TRACE(n, 5);
    }
  } else if (warp >= FIRST_EPI_WARP && warp < FIRST_EPI_WARP + 4) {
    // ===================== epilogue: project accumulator -> +bias (+x) -> bf16 -> global =====================
    const int q = warp & 3;
    const int o = q * 32 + lane;                  // accumulator row = output pixel of the tile
    const int oy_l = o / p.TW, ox_l = o - oy_l * p.TW;
    int ps = 0;
    uint32_t pph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      int b, oy0, ox0;
      tile_coords(i, b, oy0, ox0);
      const int gy = oy0 + oy_l, gx = ox0 + ox_l;
      const bool valid = (o < p.TH * p.TW) && gy < p.Ho && gx < p.Wo;
      const size_t pix = ((size_t)b * p.Ho + gy) * p.Wo + gx;
      bf16* yp = p.y + pix * p.Cout;
      const bf16* rp = p.x + pix * p.Cout;        // residual blocks: S == 1, Cin == Cout, same pixel
      tc::mbar_wait(tc::smem_u32(&proj_full[ps]), pph);
      tc::tcgen05_fence_after();
      if (warp == FIRST_EPI_WARP && lane == 0) FB_TRACE(i * p.n_chunks, 13);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.proj_col0 + ps * pstride);
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c0 + j * 8 < p.Cout) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]) + bp_s[c0 + j * 8 + e];
              if (p.residual) {
                float r[8];
                Vec8<bf16>::load(rp + c0 + j * 8, r);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += r[e];
              }
              Vec8<bf16>::store(yp + c0 + j * 8, f);
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&proj_empty[ps]));
      if (warp == FIRST_EPI_WARP && lane == 0) FB_TRACE(i * p.n_chunks, 14);
      if (++ps == p.proj_stages) { ps = 0; pph ^= 1u; }
    }
  } else if (warp < NG * GW) {
    // ===================== workers: drain (TMEM -> Hs) and depthwise (Hs -> A2) =====================
    const int g = warp / GW;
    const int wg = warp % GW;
    const int q = warp & 3;
    const int tg = (int)threadIdx.x - 32 * g * GW;   // thread index inside the group
    uint8_t* hs = hs_s + (size_t)g * 2 * MT_BYTES;
    uint8_t* a2 = a2_s + (size_t)g * MT_BYTES;
    const uint32_t hs_u0 = tc::smem_u32(hs), a2_u = tc::smem_u32(a2);
    const int nstrips_x = p.TW / TX;
    const int nstrips = nstrips_x * p.TH;
    const float rcp_nsx = 1.0f / (float)nstrips_x;
    WorkIt w = work_begin();
    for (int s = 0; s < g && w.n < total; ++s) work_next<NG>(w, p);
    int cur_i = -1, b = 0, oy0 = 0, ox0 = 0;
    while (w.n < total) {
      const int n = w.n, c = w.c;
      if (w.i != cur_i) { cur_i = w.i; tile_coords(cur_i, b, oy0, ox0); }
      const int n_c = min(HC, p.Ch - c * HC);          // valid hidden channels of this chunk (multiple of 16)
      const int ws = w.ws;
      const uint32_t kph = (uint32_t)w.kph;
      const uint32_t aux_u = tc::smem_u32(w_s + (size_t)ws * wsb + (size_t)p.kc_in * HC * 128 + (size_t)p.cpad * 128);
      if (tg == 0) FB_TRACE(n, 6);
      tc::mbar_wait(tc::smem_u32(&w_full[ws]), (uint32_t)w.wph);
      // t = 1 block (no expand conv): the TMA-loaded x tile *is* the hidden tile (same 128-byte pixel rows, same XOR
      // swizzle, image border already zero from the TMA out-of-bounds fill); one chunk per tile, x stage == item
      const uint32_t hs_u = p.has_expand ? hs_u0 : tc::smem_u32(x_s + (size_t)w.xs * xsb);
      if (!p.has_expand) tc::mbar_wait(tc::smem_u32(&x_full[w.xs]), (uint32_t)w.xph);
      // ---- drain ----
      if (p.has_expand) {
        tc::mbar_wait(tc::smem_u32(&acc_full[w.as]), (uint32_t)w.aph);
        tc::tcgen05_fence_after();
      }
      if (tg == 0) FB_TRACE(n, 7);
      for (int mt = wg >> 2; mt < n_mt && !(p.debug_skip & 1) && p.has_expand; mt += GW / 4) {
        const int pp = mt * 128 + q * 32 + lane;       // box pixel = accumulator row
        const int py = fast_div(pp, p.TWI, rcp_twi), px = pp - py * p.TWI;
        const int gy = oy0 * S - 1 + py, gx = ox0 * S - 1 + px;
        const bool in_img = (pp < P_in) && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(w.as * 128 + mt * 64);
        const uint32_t row_u = hs_u + (uint32_t)pp * 128u;
        const uint32_t sw = (uint32_t)(pp & 7);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h * 32 < n_c) {
            uint32_t v[32];
            tc::tmem_ld_32x32b_x32(t_row + (uint32_t)(h * 32), v);
            tc::tmem_ld_wait();
            if (pp < P_in) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b0 = tc::lds_f4(aux_u + (uint32_t)((h * 32 + j * 8) * 4));
                const float4 b1 = tc::lds_f4(aux_u + (uint32_t)((h * 32 + j * 8 + 4) * 4));
                uint32_t o4[4];
                o4[0] = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[j * 8 + 0], v[j * 8 + 1]), tc::pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y))));
                o4[1] = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[j * 8 + 2], v[j * 8 + 3]), tc::pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w))));
                o4[2] = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[j * 8 + 4], v[j * 8 + 5]), tc::pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y))));
                o4[3] = tc::cvt_bf16x2(tc::add_f32x2(tc::pack_f32x2(v[j * 8 + 6], v[j * 8 + 7]), tc::pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w))));
#pragma unroll
                for (int e = 0; e < 4; ++e) o4[e] = in_img ? tc::relu_bf16x2(o4[e]) : 0u;
                tc::sts_u4(row_u + ((((uint32_t)(h * 4 + j)) ^ sw) << 4), make_uint4(o4[0], o4[1], o4[2], o4[3]));
              }
            }
          }
        }
      }
      if (tg == 0) FB_TRACE(n, 8);
      tc::tcgen05_fence_before();
      __syncwarp();
      if (p.has_expand) {
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[w.as]));   // this warp's tcgen05.ld on the stage are complete
        group_sync(g, GT);                                             // Hs complete
      }
      if (tg == 0) FB_TRACE(n, 9);
      // ---- depthwise 3x3 ----
      tc::mbar_wait(tc::smem_u32(&a2_empty[g]), kph ^ 1u);   // project MMA of the previous chunk of this group has read A2
      if (tg == 0) FB_TRACE(n, 10);
      const int ncv = n_c >> 3;                        // 8, 4 or 2
      const int cv_shift = (ncv == 8) ? 3 : ((ncv == 4) ? 2 : 1);
      const int items = nstrips * ncv;
      const uint32_t wd_u = aux_u + 2u * HC * 4u;
      const uint32_t bd_u = aux_u + HC * 4u;
      for (int it = tg; it < items && !(p.debug_skip & 2); it += GT) {
        const int cv = it & (ncv - 1), strip = it >> cv_shift;
        const int sy = fast_div(strip, nstrips_x, rcp_nsx), sx = strip - sy * nstrips_x;
        const int oxl = sx * TX;
        uint64_t acc[TX][4];
        {
          const float4 b0 = tc::lds_f4(bd_u + (uint32_t)(cv * 32));
          const float4 b1 = tc::lds_f4(bd_u + (uint32_t)(cv * 32 + 16));
#pragma unroll
          for (int t = 0; t < TX; ++t) {
            acc[t][0] = f32x2(b0.x, b0.y); acc[t][1] = f32x2(b0.z, b0.w);
            acc[t][2] = f32x2(b1.x, b1.y); acc[t][3] = f32x2(b1.z, b1.w);
          }
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          uint64_t wr[3][4];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 w0 = tc::lds_f4(wd_u + (uint32_t)(((ky * 3 + kx) * HC + cv * 8) * 4));
            const float4 w1 = tc::lds_f4(wd_u + (uint32_t)(((ky * 3 + kx) * HC + cv * 8 + 4) * 4));
            wr[kx][0] = f32x2(w0.x, w0.y); wr[kx][1] = f32x2(w0.z, w0.w);
            wr[kx][2] = f32x2(w1.x, w1.y); wr[kx][3] = f32x2(w1.z, w1.w);
          }
          const int pin0 = (sy * S + ky) * p.TWI + oxl * S;
#pragma unroll
          for (int j = 0; j < NCOLS; ++j) {
            const int pin = pin0 + j;
            const uint4 u = tc::lds_u4(hs_u + (uint32_t)pin * 128u + ((((uint32_t)cv) ^ ((uint32_t)pin & 7u)) << 4));
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
            uint64_t v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = f32x2(__uint_as_float(uw[e] << 16), __uint_as_float(uw[e] & 0xffff0000u));
#pragma unroll
            for (int t = 0; t < TX; ++t) {
              const int kx = j - t * S;
              if (kx >= 0 && kx <= 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][e] = fma_f32x2(v[e], wr[kx][e], acc[t][e]);
              }
            }
          }
        }
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          const int o = sy * p.TW + oxl + t;
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o4[e] = tc::relu_bf16x2(tc::cvt_bf16x2(acc[t][e]));
          tc::sts_u4(a2_u + (uint32_t)o * 128u + ((((uint32_t)cv) ^ ((uint32_t)o & 7u)) << 4), make_uint4(o4[0], o4[1], o4[2], o4[3]));
        }
      }
      if (tg == 0) FB_TRACE(n, 11);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A2 (generic-proxy writes) -> visible to the tensor core
      group_sync(g, GT);                                             // A2 complete; everyone is done reading Hs
      if (tg == 0) tc::mbar_arrive(tc::smem_u32(&a2_full[g]));
      if (tg == 0 && !p.has_expand) tc::mbar_arrive(tc::smem_u32(&x_empty[w.xs]));   // everyone is done reading the x stage
      if (tg == 0) FB_TRACE(n, 12);
      for (int s = 0; s < NG && w.n < total; ++s) work_next<NG>(w, p);
    }
  }

#undef FB_TRACE
  // ---- teardown ----
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == WARP_ALLOC) {
    tc::tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tc::TMEM_COLS) : "memory");
  }
}

// 4-D NHWC tensor map {C, W, H, B}, box {64 ch, twi, thi, 1}, 128-byte swizzle (UMMA K-major rows), zero OOB fill.
inline bool make_tmap_x(tc::EncodeTiledFn fn, CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int twi, int thi) {
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)twi, (cuuint32_t)thi, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Output tile TH x TW (<= 128 px, TW % TX == 0) whose haloed input box fits two 128-row M tiles; fewest tiles per image wins,
// then the smaller box.
inline void pick_tile(int Ho, int Wo, int S, int* th, int* tw) {
  const int tx = (S == 1) ? 4 : 2;
  long long best_cost = -1;
  for (int TH = 1; TH <= 32; ++TH)
    for (int TW = tx; TW <= 64; TW += tx) {
      const int thi = (TH - 1) * S + 3, twi = (TW - 1) * S + 3;
      if (TH * TW > 128 || thi * twi > 256 || twi > 256 || thi > 256) continue;
      const long long tiles = (long long)cdiv(Ho, TH) * cdiv(Wo, TW);
      const long long cost = tiles * 100000 + thi * twi;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; *th = TH; *tw = TW; }
    }
}

}  // namespace fb
}  // namespace spef
