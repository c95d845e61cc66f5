// Pointwise 1x1 conv / Linear on tcgen05, second-generation pipeline (see gemm_tcgen05.cuh for the operand layouts).
//
// What the clock64 traces of v1 on B200 showed (profiles/r01_gemm_trace.txt):
//   * every mbarrier hand-off between roles costs 300-1000 cycles (MMA commit -> epilogue wake-up ~1000, epilogue
//     release -> MMA wake-up ~700), so with two TMEM accumulator stages the MMA <-> epilogue round trip (~2900 cycles)
//     caps the kernel at two tiles per round trip no matter how fast the epilogue is;
//   * per 128x64 output box the epilogue spends ~150-270 cycles in tcgen05.ld, ~500 in FP32 math + staging and ~850 in
//     coalesced st.global, serialised in the same four warps; a TMA store of the box keeps the TMA unit busy for
//     ~2000 cycles (16 cycles per 128-byte row) and is no faster.
// v2 therefore
//   * uses up to 8 accumulator stages (512 TMEM columns / 64, 128 or 256 columns per stage) so that 4-8 tiles are in
//     flight across the MMA <-> drain hand-off,
//   * splits the epilogue into "drain" warps (tcgen05.ld -> packed f32x2 bias add -> cvt.bf16x2 -> packed bf16x2 ReLU,
//     or the f32 residual add -> padded smem staging; NDG groups of 4 warps take boxes round-robin and hand the TMEM
//     stage back right after their last tcgen05.ld) and "store" warps (staging ring -> 16-byte coalesced st.global,
//     8 lanes = one 128-byte row segment), connected by a ring of staging buffers,
// so the store stream of box i overlaps the drain of box i+1 and the MMAs of the following tiles.
#pragma once
#include "gemm_tcgen05.cuh"

namespace spef {
namespace tc {

constexpr int V2_RING = 4;        // staging buffers between drain and store warps
constexpr int V2_MAX_ACC = 8;     // TMEM accumulator stages

__host__ __device__ inline int acc_stride_cols(int block_n) { return block_n <= 64 ? 64 : (block_n <= 128 ? 128 : 256); }

constexpr int V2_W_RESIDENT_MAX = 80 * 1024;  // weights stay in shared memory for the whole kernel when they fit in this

// bytes of the complete (all n-tiles, all K chunks) weight matrix as 128B-swizzled smem tiles
__host__ __device__ inline int w_region_bytes(int N, int K, int block_n) {
  return ((N + block_n - 1) / block_n) * ((K + BLOCK_K - 1) / BLOCK_K) * block_n * 128;
}
__host__ __device__ inline bool w_is_resident(int N, int K, int block_n) { return w_region_bytes(N, K, block_n) <= V2_W_RESIDENT_MAX; }
// per-stage bytes: activations only when the weights are resident
__host__ __device__ inline int stage_bytes_v2(int N, int K, int block_n) {
  return w_is_resident(N, K, block_n) ? A_STAGE_BYTES : stage_bytes(block_n);
}
// image patch: 3-D map {W, H, 3 * B planes} of the NCHW image
__device__ __forceinline__ void tma_load_3d_img(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// tcgen05.mma / tcgen05.commit issued by the elected lane of a converged warp (always the same lane: commit tracks the MMAs of
// the executing thread)
__device__ __forceinline__ void mma_elect_v2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_elect_v2(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
inline size_t smem_bytes_v2(int block_n, int num_stages, int N, int K) {
  return 1024 + (w_is_resident(N, K, block_n) ? (size_t)w_region_bytes(N, K, block_n) : 0) +
         (size_t)num_stages * stage_bytes_v2(N, K, block_n) + (size_t)V2_RING * STAGING_BYTES + (size_t)bias_floats(N) * 4 + 512;
}
inline int pick_stages_v2(int block_n, int N, int K, size_t smem_limit) {
  int s = MAX_STAGES;
  while (s > 2 && smem_bytes_v2(block_n, s, N, K) > smem_limit) --s;
  return s;
}

__device__ __forceinline__ uint64_t pack_f32x2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// two f32 (packed in a b64) -> bf16x2 (low half = first element)
__device__ __forceinline__ uint32_t cvt_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
  uint32_t r;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));
  return r;
}

// Thread layout: warps 0-3 control (TMA, MMA, TMEM alloc, spare), warps 4 .. 4+4*NDG-1 drain, then NSW store warps.
// PROD = 0: activations arrive by TMA (1x1 conv / Linear).  PROD = 1: four producer warps build the A tile by im2col from
// the NCHW f32 image (stem 3x3 s2 as an implicit GEMM, K = 27 padded to 32); they sit between the control and drain warps.
template <bool OUT_F32, int NDG, int NSW, int PROD>
__global__ void __launch_bounds__(128 + 128 * PROD + 128 * NDG + 32 * NSW, 1)
pw_gemm_tcgen05_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  if (p.pdl_early) pdl_launch_dependents();   // the next kernel of the chain may start its own set-up now (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int block_n = p.block_n;
  const int nstages = p.num_stages;
  const bool w_resident = w_is_resident(p.N, p.K, block_n);
  const int sbytes = stage_bytes_v2(p.N, p.K, block_n);
  const int acc_stride = acc_stride_cols(block_n);
  const int acc_stages = min(V2_MAX_ACC, TMEM_COLS / acc_stride);
  uint8_t* w_region = smem;                                     // resident weights: [n_tile][k_chunk][block_n rows x 128 B]
  uint8_t* stage_base = smem + (w_resident ? w_region_bytes(p.N, p.K, block_n) : 0);
  uint8_t* staging = stage_base + (size_t)nstages * sbytes;
  float* bias_s = reinterpret_cast<float*>(staging + (size_t)V2_RING * STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + bias_floats(p.N));
  uint64_t* full_bar = bars;                                   // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty_bar = full_bar + MAX_STAGES;                 // [MAX_STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;            // [V2_MAX_ACC]  MMA -> drain
  uint64_t* tmem_empty_bar = tmem_full_bar + V2_MAX_ACC;       // [V2_MAX_ACC]  drain -> MMA
  uint64_t* sfull_bar = tmem_empty_bar + V2_MAX_ACC;           // [V2_RING]     drain -> store
  uint64_t* sempty_bar = sfull_bar + V2_RING;                  // [V2_RING]     store -> drain
  uint64_t* w_bar = sempty_bar + V2_RING;                      // [1]           resident weights landed
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(w_bar + 1);
  // stem patch ring (PROD, patch mode): barriers and buffers behind everything else (the host adds the bytes)
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(tmem_ptr_s + 2);      // [PATCH_STAGES]  TMA -> im2col producers
  uint64_t* patch_empty = patch_full + PATCH_STAGES;                        // [PATCH_STAGES]  producers -> TMA
  uint8_t* patch_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(patch_empty + PATCH_STAGES) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + block_n - 1) / block_n;
  const int num_tiles = m_tiles * n_tiles;
  const int k_chunks = (p.K + BLOCK_K - 1) / BLOCK_K;
  constexpr int COLS_PER_BOX = OUT_F32 ? 32 : 64;  // one 128-byte staging row
  constexpr int FIRST_DRAIN_WARP = 4 + 4 * PROD;
  constexpr int FIRST_STORE_WARP = FIRST_DRAIN_WARP + 4 * NDG;

  for (int i = threadIdx.x; i < bias_floats(p.N); i += (int)blockDim.x) bias_s[i] = (i < p.N) ? p.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), PROD ? 128 : 1);   // TMA: one arrive + tx bytes; im2col: every producer thread
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < V2_MAX_ACC; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4 * NDG);   // one arrive per drain warp
    }
    for (int i = 0; i < V2_RING; ++i) {
      mbar_init(smem_u32(&sfull_bar[i]), 4);              // the four warps of the group that filled the slot
      mbar_init(smem_u32(&sempty_bar[i]), NSW);           // one arrive per store warp
    }
    mbar_init(smem_u32(w_bar), 1);
    if (PROD && p.patch_mode) {
      for (int i = 0; i < PATCH_STAGES; ++i) {
        mbar_init(smem_u32(&patch_full[i]), 1);
        mbar_init(smem_u32(&patch_empty[i]), 128);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();   // everything above is independent of the previous kernel's output (common.cuh)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)sbytes;
      int tm = (int)blockIdx.x / n_tiles, tn = (int)blockIdx.x % n_tiles;
      const int w_tile_bytes = block_n * 128;
      if (w_resident) {  // the whole weight matrix is fetched once per CTA and stays in shared memory
        const uint32_t wb = smem_u32(w_bar);
        mbar_arrive_expect_tx(wb, (uint32_t)w_region_bytes(p.N, p.K, block_n));
        for (int j = 0; j < n_tiles; ++j)
          for (int kc = 0; kc < k_chunks; ++kc)
            tma_load_2d(smem_u32(w_region + (size_t)(j * k_chunks + kc) * w_tile_bytes), &tmW, kc * BLOCK_K, j * block_n, wb);
      }
      for (int tile = blockIdx.x; PROD == 0 && tile < num_tiles; tile += gridDim.x) {
        const int m_idx = tm * BLOCK_M, n_idx = tn * block_n;
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, tx_bytes);
          uint8_t* sa = stage_base + (size_t)stage * sbytes;
          tma_load_2d(smem_u32(sa), &tmA, kc * BLOCK_K, m_idx, fb);
          if (!w_resident) tma_load_2d(smem_u32(sa + A_STAGE_BYTES), &tmW, kc * BLOCK_K, n_idx, fb);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (n_tiles == 1) { tm += gridDim.x; } else { const int t2 = tile + (int)gridDim.x; tm = t2 / n_tiles; tn = t2 % n_tiles; }
      }
    }
    if (PROD && p.patch_mode) {   // stem: stage the input patch of every tile of this CTA (ring of PATCH_STAGES)
      // ncu (round 2, wait sites of the stem): this single-thread loop never waited for a free patch slot while the im2col producers
      // polled patch_full 18x per tile and the MMA issuer full_bar 5x -- the patch issue itself (two divisions, five sequential
      // TMA launches with their address arithmetic, ~110 dependent instructions) paced the kernel at ~1150 cycles per 128-pixel
      // tile.  Now the whole warp runs the loop: tile coordinates are carried, lane 0 waits for the slot and arms the barrier, and
      // lane ch issues column chunk ch.
      const int tiles_img = (p.out_h / 2) * p.patch_tiles_x;
      const uint32_t cbytes = (uint32_t)(p.patch_w * 5 * 3 * (p.img_u8 ? 1 : 4));   // one column chunk
      const uint32_t pbytes = cbytes * (uint32_t)p.patch_chunks;
      int ps = 0;
      uint32_t pph = 0;
      int tb = (int)blockIdx.x / tiles_img, ty, tx;
      {
        const int r = (int)blockIdx.x - tb * tiles_img;
        ty = r / p.patch_tiles_x; tx = r - ty * p.patch_tiles_x;
      }
      const int g = (int)gridDim.x, rows_img = p.out_h / 2;
      const int db = g / tiles_img, dr = g - db * tiles_img, dty = dr / p.patch_tiles_x, dtx = dr - dty * p.patch_tiles_x;
      // The innermost TMA coordinate must stay 16-byte aligned (x = -1 faulted with "illegal instruction" while -1 in an outer
      // dimension is fine): the patch starts patch_x0 = 4 (f32) or 16 (u8) pixels left of the tile, pixel column -1 = patch column patch_x0 - 1.
      const uint32_t my_dst = smem_u32(patch_s) + (uint32_t)lane * cbytes;
      const int my_x = lane * p.patch_w - p.patch_x0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t fb = smem_u32(&patch_full[ps]);
        if (lane == 0) {
          mbar_wait(smem_u32(&patch_empty[ps]), pph ^ 1);
          mbar_arrive_expect_tx(fb, pbytes);
        }
        __syncwarp();
        if (lane < p.patch_chunks) tma_load_3d_img(my_dst + (uint32_t)ps * (uint32_t)PATCH_STAGE_BYTES, &tmA, tx * 128 + my_x, ty * 4 - 1, 3 * tb, fb);
        if (++ps == PATCH_STAGES) { ps = 0; pph ^= 1; }
        tx += dtx; ty += dty; tb += db;
        if (tx >= p.patch_tiles_x) { tx -= p.patch_tiles_x; ++ty; }
        if (ty >= rows_img) { ty -= rows_img; ++tb; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop converged and one elected lane issues (elect.sync inside the asm): under `if (lane == 0)` the
    // compiler wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY sequence, and a single thread retires ~1 instruction per
    // 10 cycles, so the issuer -- not the tensor pipe -- paced the K-heavy layers (345 cycles per MMA at 960 -> 320).
    // Descriptors are advanced by adding to precomputed bases.
    {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tn = (int)blockIdx.x % n_tiles;
      const uint64_t a_base = make_smem_desc_sw128(smem_u32(stage_base));
      const uint64_t bs_base = make_smem_desc_sw128(smem_u32(stage_base + A_STAGE_BYTES));   // streamed weights: behind the A tile of the stage
      const uint64_t bw_base = make_smem_desc_sw128(smem_u32(w_region));                      // resident weights
      const uint32_t s_step = (uint32_t)sbytes >> 4, w_step = (uint32_t)(block_n * 128) >> 4;
      const uint32_t kst_last = (uint32_t)((p.K - (k_chunks - 1) * BLOCK_K + 15) / 16);
      if (w_resident) mbar_wait(smem_u32(w_bar), 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
        uint64_t bw = bw_base + (uint64_t)((uint32_t)(tn * k_chunks) * w_step);
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint64_t a_desc = a_base + (uint64_t)((uint32_t)stage * s_step);
          const uint64_t b_desc = w_resident ? bw : bs_base + (uint64_t)((uint32_t)stage * s_step);
          const uint32_t ksteps = (kc == k_chunks - 1) ? kst_last : 4u;
          for (uint32_t k = 0; k < ksteps; ++k)
            mma_elect_v2(d_tmem, a_desc + (uint64_t)(k * 2u), b_desc + (uint64_t)(k * 2u), idesc, (kc > 0 || k > 0) ? 1u : 0u);
          commit_elect_v2(smem_u32(&empty_bar[stage]));
          if (++stage == nstages) { stage = 0; phase ^= 1; }
          bw += w_step;
        }
        commit_elect_v2(smem_u32(&tmem_full_bar[acc]));
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
        if (n_tiles > 1) tn = (tile + (int)gridDim.x) % n_tiles;
      }
    }
    __syncwarp();
  } else if (PROD >= 1 && warp >= 4 && warp < FIRST_DRAIN_WARP) {
    // ===================== im2col producer warps (stem): thread t builds row t of the A tile =====================
    // A row = the 27 taps (ci, ky, kx) of one output pixel as bf16, k = (ci*3 + ky)*3 + kx, zero padded to 32; written in
    // the 128B-swizzled K-major layout the UMMA descriptor expects (16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)).
    // PROD producer groups of 128 threads take the CTA's tiles round-robin (tile k of the CTA -> group k % PROD), so
    // PROD x 27 x 128 image loads are in flight per SM
    const int t = ((int)threadIdx.x - 128) & 127;
    const int pg = ((int)threadIdx.x - 128) >> 7;
    const int H = p.img_h, W = p.img_w, Ho = p.out_h, Wo = p.out_w;
    // uint8 images: a 256-entry table maps a pixel to bf16(float(u8) / 255.0f), exactly what rounding the reference's
    // float tensor gives; the table lives at the end of the bias area (bias_floats() reserves 256 spare floats)
    uint16_t* lut = reinterpret_cast<uint16_t*>(bias_s + bias_floats(p.N) - 128);
    if (p.img_u8) {
      for (int i = (int)threadIdx.x - 128; i < 256; i += 128 * PROD) {
        const bf16 h = __float2bfloat16_rn(__fdiv_rn((float)i, 255.0f));
        lut[i] = *reinterpret_cast<const uint16_t*>(&h);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(128 * PROD) : "memory");  // producer warps only
    }
    // gather: 27 taps of this thread's output pixel as bf16 bit patterns (zero outside the image / beyond M)
    auto gather = [&](int tile, uint16_t (&tap)[27]) {
      const int m = tile * BLOCK_M + t;
      const bool m_ok = m < p.M;
      const int mm = m_ok ? m : 0;
      const int ox = mm % Wo;
      const int oy = (mm / Wo) % Ho;
      const int b = mm / (Wo * Ho);
      const size_t base = (size_t)b * 3 * H * W;
      if (p.img_u8) {
        const uint8_t* ib = reinterpret_cast<const uint8_t*>(p.img) + base;
        uint8_t raw[27];
        bool okm[27];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy * 2 - 1 + ky;
          const bool yok = m_ok && (iy >= 0) && (iy < H);
          const int iyc = min(max(iy, 0), H - 1);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            const bool ok = yok && (ix >= 0) && (ix < W);
            const int ixc = min(max(ix, 0), W - 1);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              raw[(ci * 3 + ky) * 3 + kx] = __ldg(ib + ((size_t)ci * H + iyc) * W + ixc);
              okm[(ci * 3 + ky) * 3 + kx] = ok;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) tap[k] = okm[k] ? lut[raw[k]] : (uint16_t)0;
      } else {
        const float* ib = p.img + base;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy * 2 - 1 + ky;
          const bool yok = m_ok && (iy >= 0) && (iy < H);
          const int iyc = min(max(iy, 0), H - 1);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            const bool ok = yok && (ix >= 0) && (ix < W);
            const int ixc = min(max(ix, 0), W - 1);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              const float v = __ldg(ib + ((size_t)ci * H + iyc) * W + ixc);
              const bf16 h = __float2bfloat16_rn(ok ? v : 0.f);
              tap[(ci * 3 + ky) * 3 + kx] = *reinterpret_cast<const uint16_t*>(&h);
            }
          }
        }
      }
    };
    int stage = pg % nstages;
    uint32_t phase = (uint32_t)((pg / nstages) & 1);
    uint16_t cur[27], nxt[27];
    const int tile0 = (int)blockIdx.x + pg * (int)gridDim.x, tstep = PROD * (int)gridDim.x;
    // patch mode: thread t = output pixel (row t >> 6, column t & 63) of the 2 x 64 patch; tap (ci, ky, kx) sits at
    // patch[ci][2 * row + ky][2 * col + kx] of the TMA-staged input patch (zero outside the image)
    int pst = pg % PATCH_STAGES;
    uint32_t pphase = (uint32_t)((pg / PATCH_STAGES) & 1);
    auto gather_patch = [&](uint16_t (&tap)[27]) {
      mbar_wait(smem_u32(&patch_full[pst]), pphase);
      const int pr = t >> 6, pc = t & 63;
      // taps kx = 0, 1, 2 = patch columns colA, colA + 1, colA + 2 with colA = 2c + patch_x0 - 1 (odd): colA + 1, colA + 2 share a
      // chunk (chunk widths are even) and form an aligned pair; colA may sit in the previous chunk
      const int colA = 2 * pc + p.patch_x0 - 1, chA = colA / p.patch_w, inA = colA - chA * p.patch_w;
      const int chB = (colA + 1) / p.patch_w, inB = colA + 1 - chB * p.patch_w;
      const int cpix = p.patch_w * 15;   // pixels per chunk
      if (p.img_u8) {
        const uint8_t* pb = patch_s + (size_t)pst * PATCH_STAGE_BYTES + (size_t)(2 * pr) * p.patch_w;
        const int oA = chA * cpix + inA, oB = chB * cpix + inB;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint8_t* rowp = pb + (size_t)(ci * 5 + ky) * p.patch_w;
            tap[(ci * 3 + ky) * 3 + 0] = lut[rowp[oA]];
            tap[(ci * 3 + ky) * 3 + 1] = lut[rowp[oB]];
            tap[(ci * 3 + ky) * 3 + 2] = lut[rowp[oB + 1]];
          }
      } else {
        const float* pb = reinterpret_cast<const float*>(patch_s + (size_t)pst * PATCH_STAGE_BYTES) + (size_t)(2 * pr) * p.patch_w;
        const int oA = chA * cpix + inA, oB = chB * cpix + inB;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float* rowp = pb + (size_t)(ci * 5 + ky) * p.patch_w;
            const float v0 = rowp[oA];
            const float2 v12 = *reinterpret_cast<const float2*>(rowp + oB);   // patch columns colA + 1, colA + 2 (8-byte aligned)
            const bf16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v12.x), h2 = __float2bfloat16_rn(v12.y);
            tap[(ci * 3 + ky) * 3 + 0] = *reinterpret_cast<const uint16_t*>(&h0);
            tap[(ci * 3 + ky) * 3 + 1] = *reinterpret_cast<const uint16_t*>(&h1);
            tap[(ci * 3 + ky) * 3 + 2] = *reinterpret_cast<const uint16_t*>(&h2);
          }
      }
      mbar_arrive(smem_u32(&patch_empty[pst]));   // this thread's taps are in registers
      pst += PROD;
      if (pst >= PATCH_STAGES) { pst -= PATCH_STAGES; pphase ^= 1; }
    };
    if (p.patch_mode) {
      // Lean loop for the TMA-staged patch (ncu's source page: 350 instructions per thread and tile in the generic loop below --
      // two divisions by the chunk width, 27 single bf16 conversions, 13 byte-permute packs, 27 register copies).  Everything
      // that depends only on the thread is computed once; adjacent taps are converted and packed by one cvt.rn.bf16x2.f32.
      const int pr = t >> 6, pcx = t & 63;
      const int colA = 2 * pcx + p.patch_x0 - 1, chA = colA / p.patch_w, inA = colA - chA * p.patch_w;
      const int chB = (colA + 1) / p.patch_w, inB = colA + 1 - chB * p.patch_w;
      const int cpix = p.patch_w * 15;   // pixels per chunk
      const int oA = chA * cpix + inA + (2 * pr) * p.patch_w, oB = chB * cpix + inB + (2 * pr) * p.patch_w;
      const uint32_t patch_u = smem_u32(patch_s);
      const uint32_t sa_row = (uint32_t)t * 128u, swz = (uint32_t)t & 7u;
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        mbar_wait(smem_u32(&patch_full[pst]), pphase);
        uint32_t pk[16];
        if (p.img_u8) {
          const uint8_t* pb = patch_s + (size_t)pst * PATCH_STAGE_BYTES;
          uint16_t tap[28];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint8_t* rowp = pb + (ci * 5 + ky) * p.patch_w;
              tap[(ci * 3 + ky) * 3 + 0] = lut[rowp[oA]];
              tap[(ci * 3 + ky) * 3 + 1] = lut[rowp[oB]];
              tap[(ci * 3 + ky) * 3 + 2] = lut[rowp[oB + 1]];
            }
          tap[27] = 0;
#pragma unroll
          for (int k = 0; k < 14; ++k) pk[k] = (uint32_t)tap[2 * k] | ((uint32_t)tap[2 * k + 1] << 16);
        } else {
          const uint32_t pb = patch_u + (uint32_t)pst * (uint32_t)PATCH_STAGE_BYTES;
          float f[28];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t rowu = pb + (uint32_t)((ci * 5 + ky) * p.patch_w) * 4u;
              float v0, v1, v2;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(rowu + (uint32_t)oA * 4u));
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v1), "=f"(v2) : "r"(rowu + (uint32_t)oB * 4u));   // 8-byte aligned
              f[(ci * 3 + ky) * 3 + 0] = v0; f[(ci * 3 + ky) * 3 + 1] = v1; f[(ci * 3 + ky) * 3 + 2] = v2;
            }
          f[27] = 0.f;
#pragma unroll
          for (int k = 0; k < 14; ++k) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[k]) : "f"(f[2 * k + 1]), "f"(f[2 * k]));
        }
        pk[14] = 0u; pk[15] = 0u;
        mbar_arrive(smem_u32(&patch_empty[pst]));   // this thread's taps are in registers
        pst += PROD;
        if (pst >= PATCH_STAGES) { pst -= PATCH_STAGES; pphase ^= 1; }
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t sa = smem_u32(stage_base + (size_t)stage * sbytes) + sa_row;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts_u4(sa + (((uint32_t)c ^ swz) << 4), make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        mbar_arrive(smem_u32(&full_bar[stage]));
        stage += PROD;
        if (stage >= nstages) { stage -= nstages; phase ^= 1; }
      }
    } else {
    if (tile0 < num_tiles) gather(tile0, cur);
    for (int tile = tile0; tile < num_tiles; tile += tstep) {
      // software pipeline: the taps of the next tile are in flight while this one is staged
      const int next = tile + tstep;
      if (p.patch_mode) gather_patch(cur);
      else if (next < num_tiles) gather(next, nxt);
      uint32_t pk[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t lo = (2 * k < 27) ? cur[2 * k] : 0u;
        const uint32_t hi = (2 * k + 1 < 27) ? cur[2 * k + 1] : 0u;
        pk[k] = lo | (hi << 16);
      }
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      const uint32_t sa = smem_u32(stage_base + (size_t)stage * sbytes) + (uint32_t)t * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts_u4(sa + (((uint32_t)c ^ ((uint32_t)t & 7u)) << 4), make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(smem_u32(&full_bar[stage]));
      stage += PROD;
      if (stage >= nstages) { stage -= nstages; phase ^= 1; }
#pragma unroll
      for (int k = 0; k < 27; ++k) cur[k] = nxt[k];
    }
    }
  } else if (warp >= FIRST_DRAIN_WARP && warp < FIRST_STORE_WARP) {
    // ===================== drain warps: TMEM -> registers -> bf16 -> staging ring =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = (warp - FIRST_DRAIN_WARP) >> 2;        // drain group
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t box = 0;                       // global box counter: identical sequence in every drain / store warp
    int tm = (int)blockIdx.x / n_tiles, tn = (int)blockIdx.x % n_tiles;
    if (PROD && NDG == 1 && !OUT_F32 && p.patch_mode && p.N == 32 && n_tiles == 1 && p.residual == nullptr && p.relu) {
      // Stem: one 32-column box per tile.  Bias pairs live in registers, ReLU is folded into the conversion, no box bookkeeping.
      uint64_t bias2[16];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b4 = lds_f4(smem_u32(bias_s) + (uint32_t)g * 16u);
        bias2[2 * g] = pack_f32x2(__float_as_uint(b4.x), __float_as_uint(b4.y));
        bias2[2 * g + 1] = pack_f32x2(__float_as_uint(b4.z), __float_as_uint(b4.w));
      }
      const uint32_t srow = (uint32_t)row * (uint32_t)STAGING_PITCH;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++box) {
        const int slot = (int)(box % V2_RING);
        const uint32_t ring_phase = (box / V2_RING) & 1u;
        mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride), v);
        mbar_wait(smem_u32(&sempty_bar[slot]), ring_phase ^ 1);  // staging slot drained by the store warps
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
        const uint32_t sb = smem_u32(staging + (size_t)slot * STAGING_BYTES) + srow;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint64_t sum = add_f32x2(pack_f32x2(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]), bias2[g * 4 + e]);
            uint32_t lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(sum));
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[e]) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
          }
          sts_u4(sb + (uint32_t)g * 16u, make_uint4(o[0], o[1], o[2], o[3]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sfull_bar[slot]));
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    } else
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = tm * BLOCK_M, n_idx = tn * block_n;
      const int m = m_idx + row;
      const int ncols = min(block_n, p.N - n_idx);
      const int nboxes = (ncols + COLS_PER_BOX - 1) / COLS_PER_BOX;
      // index of the last box of this tile that belongs to this group (-1: none)
      int last_mine = -1;
      for (int b = nboxes - 1; b >= 0; --b) {
        if ((int)((box + (uint32_t)b) % NDG) == grp) { last_mine = b; break; }
      }
      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride);
      if (last_mine < 0) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      }
      for (int b = 0; b < nboxes; ++b, ++box) {
        if ((int)(box % NDG) != grp) continue;
        const int c0 = b * COLS_PER_BOX;
        const int slot = (int)(box % V2_RING);
        const uint32_t ring_phase = (box / V2_RING) & 1u;
        uint32_t v[COLS_PER_BOX];
        {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v0);
          if constexpr (!OUT_F32) {
            if (c0 + 32 < ncols) {  // second 32-column half only when it holds real columns
              uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
              tmem_ld_32x32b_x32(t_row + (uint32_t)(c0 + 32), v1);
            }
          }
        }
        mbar_wait(smem_u32(&sempty_bar[slot]), ring_phase ^ 1);  // staging slot drained by the store warps
        tmem_ld_wait();
        if (b == last_mine) {
          // every tcgen05.ld of this warp on this accumulator has completed -> hand the TMEM stage back now
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
        }
        const uint32_t sb = smem_u32(staging + (size_t)slot * STAGING_BYTES) + (uint32_t)row * (uint32_t)STAGING_PITCH;
        const uint32_t bias_sa = smem_u32(bias_s + n_idx + c0);
        if constexpr (OUT_F32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = lds_f4(bias_sa + (uint32_t)g * 16u);
            uint4 o;
            o.x = __float_as_uint(__uint_as_float(v[g * 4 + 0]) + b4.x);
            o.y = __float_as_uint(__uint_as_float(v[g * 4 + 1]) + b4.y);
            o.z = __float_as_uint(__uint_as_float(v[g * 4 + 2]) + b4.z);
            o.w = __float_as_uint(__uint_as_float(v[g * 4 + 3]) + b4.w);
            sts_u4(sb + (uint32_t)g * 16u, o);
          }
        } else if (p.residual == nullptr && c0 + COLS_PER_BOX <= ncols) {
          // full box without skip connection: no per-piece bounds checks, ReLU folded into the conversion (cvt.rn.relu)
          if (p.relu) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b0 = lds_f4(bias_sa + (uint32_t)g * 32u);
              const float4 b1 = lds_f4(bias_sa + (uint32_t)g * 32u + 16u);
              const uint64_t bb[4] = {pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y)), pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w)),
                                      pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y)), pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w))};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint64_t sum = add_f32x2(pack_f32x2(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]), bb[e]);
                uint32_t lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(sum));
                asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[e]) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
              }
              sts_u4(sb + (uint32_t)g * 16u, make_uint4(o[0], o[1], o[2], o[3]));
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b0 = lds_f4(bias_sa + (uint32_t)g * 32u);
              const float4 b1 = lds_f4(bias_sa + (uint32_t)g * 32u + 16u);
              uint32_t o[4];
              o[0] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 0], v[g * 8 + 1]), pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y))));
              o[1] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 2], v[g * 8 + 3]), pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w))));
              o[2] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 4], v[g * 8 + 5]), pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y))));
              o[3] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 6], v[g * 8 + 7]), pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w))));
              sts_u4(sb + (uint32_t)g * 16u, make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
        } else if (p.residual == nullptr) {
          // packed path: f32x2 bias add, cvt to bf16x2, ReLU on the packed pair (round(max(x,0)) == max(round(x),0))
#pragma unroll
          for (int g = 0; g < 8; ++g) {  // eight 16-byte pieces = 64 bf16 columns
            if (c0 + g * 8 >= ncols) break;
            const float4 b0 = lds_f4(bias_sa + (uint32_t)g * 32u);
            const float4 b1 = lds_f4(bias_sa + (uint32_t)g * 32u + 16u);
            uint32_t o[4];
            o[0] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 0], v[g * 8 + 1]), pack_f32x2(__float_as_uint(b0.x), __float_as_uint(b0.y))));
            o[1] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 2], v[g * 8 + 3]), pack_f32x2(__float_as_uint(b0.z), __float_as_uint(b0.w))));
            o[2] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 4], v[g * 8 + 5]), pack_f32x2(__float_as_uint(b1.x), __float_as_uint(b1.y))));
            o[3] = cvt_bf16x2(add_f32x2(pack_f32x2(v[g * 8 + 6], v[g * 8 + 7]), pack_f32x2(__float_as_uint(b1.z), __float_as_uint(b1.w))));
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = relu_bf16x2(o[e]);
            }
            sts_u4(sb + (uint32_t)g * 16u, make_uint4(o[0], o[1], o[2], o[3]));
          }
        } else {
          // linear bottleneck with skip connection: y = x + (acc + bias), one rounding (pytorch_layers.py:93-98)
          const bf16* rp = p.residual + (size_t)m * p.N + n_idx + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (c0 + g * 8 >= ncols) break;
            const float4 b0 = lds_f4(bias_sa + (uint32_t)g * 32u);
            const float4 b1 = lds_f4(bias_sa + (uint32_t)g * 32u + 16u);
            float f[8] = {__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y,
                          __uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w,
                          __uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y,
                          __uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w};
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = relu_nan(f[e]);
            }
            if (m < p.M && n_idx + c0 + g * 8 < p.N) {
              float r[8];
              Vec8<bf16>::load(rp + g * 8, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += r[e];
            }
            sts_u4(sb + (uint32_t)g * 16u, Vec8<bf16>::pack(f));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sfull_bar[slot]));
      }
      if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
      if (n_tiles == 1) { tm += gridDim.x; } else { const int t2 = tile + (int)gridDim.x; tm = t2 / n_tiles; tn = t2 % n_tiles; }
    }
  } else if (warp >= FIRST_STORE_WARP) {
    // ===================== store warps: staging ring -> coalesced st.global =====================
    constexpr int ELEMS_PER_PIECE = OUT_F32 ? 4 : 8;
    constexpr int ESZ = OUT_F32 ? 4 : 2;
    constexpr int NST = 32 * NSW;
    const int ts = (int)threadIdx.x - 32 * FIRST_STORE_WARP;
    const int pc = ts & 7;            // 16-byte piece inside the 128-byte row segment (fixed per thread)
    const int r0 = ts >> 3;           // first row handled by this thread; then += NST / 8
    uint8_t* outb = reinterpret_cast<uint8_t*>(p.out);
    uint32_t box = 0;
    int tm = (int)blockIdx.x / n_tiles, tn = (int)blockIdx.x % n_tiles;
    if (PROD && p.patch_mode && !OUT_F32 && p.N == 32 && n_tiles == 1) {
      // Stem (32 channels = 64-byte rows): ncu's source page showed this role as the slowest of the kernel -- 338 instructions per
      // tile, 130 of them IMAD (64-bit address arithmetic per store, two divisions per tile) with half of the lanes idle (the
      // generic loop below assumes 128-byte rows).  Here: four 16-byte pieces per row so that every lane stores, per-thread byte
      // offsets computed once, the (image, tile row, tile column) of the tile carried instead of divided.
      const int pc4 = ts & 3, r04 = ts >> 2;
      constexpr int RSTEP = NST / 4, NPASS = BLOCK_M / RSTEP;
      uint32_t goff[NPASS], soff[NPASS];
#pragma unroll
      for (int k = 0; k < NPASS; ++k) {
        const int r = r04 + k * RSTEP;          // row r of the tile = output pixel (2 ty + (r >> 6), 64 tx + (r & 63))
        goff[k] = (uint32_t)((((r >> 6) * p.out_w + (r & 63)) * p.ldd + pc4 * 8) * 2);
        soff[k] = (uint32_t)(r * STAGING_PITCH + pc4 * 16);
      }
      const int TXn = p.patch_tiles_x, TYn = p.out_h / 2, G = (int)gridDim.x;
      int tx = tm % TXn, ty = (tm / TXn) % TYn, b = tm / (TXn * TYn);
      const int d_tx = G % TXn, d_ty = (G / TXn) % TYn, d_b = G / (TXn * TYn);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++box) {
        const int slot = (int)(box % V2_RING);
        const uint32_t ring_phase = (box / V2_RING) & 1u;
        uint8_t* base = outb + ((size_t)(b * p.out_h + 2 * ty) * p.out_w + 64 * tx) * (size_t)(p.ldd * 2);
        mbar_wait(smem_u32(&sfull_bar[slot]), ring_phase);
        const uint32_t sbuf = smem_u32(staging + (size_t)slot * STAGING_BYTES);
        uint4 val[NPASS];
#pragma unroll
        for (int k = 0; k < NPASS; ++k) val[k] = lds_u4(sbuf + soff[k]);
#pragma unroll
        for (int k = 0; k < NPASS; ++k) *reinterpret_cast<uint4*>(base + goff[k]) = val[k];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sempty_bar[slot]));
        tx += d_tx;
        if (tx >= TXn) { tx -= TXn; ++ty; }
        ty += d_ty;
        if (ty >= TYn) { ty -= TYn; ++b; }
        b += d_b;
      }
    } else
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = tm * BLOCK_M, n_idx = tn * block_n;
      const int ncols = min(block_n, p.N - n_idx);
      const int rows_here = min(BLOCK_M, p.M - m_idx);
      for (int c0 = 0; c0 < ncols; c0 += COLS_PER_BOX, ++box) {
        const int slot = (int)(box % V2_RING);
        const uint32_t ring_phase = (box / V2_RING) & 1u;
        const int gcol = n_idx + c0 + pc * ELEMS_PER_PIECE;
        const bool col_ok = gcol < p.N;
        mbar_wait(smem_u32(&sfull_bar[slot]), ring_phase);
        const uint32_t sbuf = smem_u32(staging + (size_t)slot * STAGING_BYTES) + (uint32_t)(pc * 16);
        uint8_t* gp = outb + ((size_t)(m_idx + r0) * p.ldd + gcol) * ESZ;
        const size_t gstep = (size_t)(NST / 8) * p.ldd * ESZ;
        if (PROD && p.patch_mode) {
          // patch tile: row r of the tile = output pixel (2 * ty + (r >> 6), 64 * tx + (r & 63)) of image b
          const int tiles_img = (p.out_h / 2) * p.patch_tiles_x;
          const int b = tm / tiles_img, rr = tm - b * tiles_img;
          const int ty = rr / p.patch_tiles_x, tx = rr - ty * p.patch_tiles_x;
          const size_t pix0 = ((size_t)b * p.out_h + 2 * ty) * p.out_w + 64 * tx;
          if (col_ok) {
#pragma unroll
            for (int k = 0; k < 1024 / NST; ++k) {
              const int r = r0 + k * (NST / 8);
              const uint4 val = lds_u4(sbuf + (uint32_t)(r * STAGING_PITCH));
              *reinterpret_cast<uint4*>(outb + ((pix0 + (size_t)(r >> 6) * p.out_w + (r & 63)) * p.ldd + gcol) * ESZ) = val;
            }
          }
        } else if (col_ok && rows_here == BLOCK_M) {
          // full tile (all but the last M tile): loads first, then stores, 32-bit offsets from one 64-bit base (ncu's source page:
          // the predicated loop below spent 197 instructions per box and thread, most of them address arithmetic)
          const uint32_t gstep32 = (uint32_t)gstep;
          uint4 val[1024 / NST];
#pragma unroll
          for (int k = 0; k < 1024 / NST; ++k) val[k] = lds_u4(sbuf + (uint32_t)((r0 + k * (NST / 8)) * STAGING_PITCH));
#pragma unroll
          for (int k = 0; k < 1024 / NST; ++k) *reinterpret_cast<uint4*>(gp + (uint32_t)k * gstep32) = val[k];
        } else if (col_ok) {
#pragma unroll
          for (int k = 0; k < 1024 / NST; ++k) {
            const int r = r0 + k * (NST / 8);
            if (r < rows_here) {
              const uint4 val = lds_u4(sbuf + (uint32_t)(r * STAGING_PITCH));
              *reinterpret_cast<uint4*>(gp + (size_t)k * gstep) = val;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sempty_bar[slot]));
      }
      if (n_tiles == 1) { tm += gridDim.x; } else { const int t2 = tile + (int)gridDim.x; tm = t2 / n_tiles; tn = t2 % n_tiles; }
    }
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace spef
