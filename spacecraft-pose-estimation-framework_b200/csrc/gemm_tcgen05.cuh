// Pointwise 1x1 conv / Linear on the 5th-gen tensor cores:  D[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ R)
//   A = NHWC activations (bf16, K-major = channel contiguous), W = PyTorch [Cout,Cin] weights (bf16, K-major,
//   no repack), D = bf16 NHWC activations (or f32 logits for the head).
// Persistent, warp-specialised, one CTA per SM:
//   warp 0   TMA producer      cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier complete_tx
//   warp 1   MMA issuer        one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n, K=16),
//                              accumulators in TMEM (2 stages x block_n columns), tcgen05.commit -> mbarriers
//   warp 2   TMEM allocator
//   warps 4.. epilogue (NG groups of 4 warps)  tcgen05.ld 32x32b -> +bias -> ReLU -> +residual -> bf16 -> swizzled smem -> TMA store
// K is tiled in 64-element (128-byte) chunks; K tails (16, 24, 32, 96, 144, 160) and M/N tails rely on TMA
// out-of-bounds zero fill on loads and clipping on stores, so no layer needs padding in HBM.
// Reference semantics: src/modeling/common/pytorch_layers.py:78-79,85-86,93-98; mobilenet_v2.py:264; ursonet.py:31-32.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace spef {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * 128;      // 16 KB
constexpr int STAGING_PITCH = 144;                // copy-out mode: 128-byte box row + 16-byte pad (conflict-free)
constexpr int STAGING_BYTES = BLOCK_M * STAGING_PITCH;  // 18 KB (multiple of 1024); TMA-store mode uses the first 16 KB
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 8;

struct GemmParams {
  const float* bias;   // [>= N rounded up to 64] f32 (zero padded)
  const bf16* residual;  // [M,N] bf16 or nullptr
  int M, N, K;
  int block_n;         // MMA N (multiple of 16, <= 256; multiple of 64 when N > block_n)
  int num_stages;      // smem pipeline depth
  int relu;
  int store_mode;      // 0: padded smem staging + coalesced st.global by the epilogue threads; 1: swizzled staging + TMA store
  void* out;           // D (copy-out mode)
  int ldd;             // row pitch of D in elements
  long long* trace;    // debug: clock64 timestamps of CTA 0 (nullptr in production)
  // im2col producer (stem as an implicit GEMM): image [B,3,H,W] f32 NCHW, output pixels [B,Ho,Wo]
  const float* img;
  int img_h, img_w, out_h, out_w;
  int img_u8;          // 1: img points at uint8 pixels, value = float(u8) / 255.0f (torchvision ToTensor)
  // patch mode (stem): an M tile is a 2-row x 64-column patch of output pixels; its 5 x (2*64+1) x 3 input patch is staged in
  // shared memory by TMA (tmA = 4-D map of the NCHW image, zero fill outside) and the im2col gather reads shared memory
  // The patch is fetched as column chunks of 128 bytes (every TMA box in this library has an inner extent <= 128 B; a 528-byte
  // inner box was accepted by cuTensorMapEncodeTiled but faulted with "illegal instruction" at the first load):
  // chunk c = pixels [c * patch_w, (c + 1) * patch_w) stored as [3 planes][5 rows][patch_w]
  int patch_mode, patch_tiles_x, patch_w, patch_chunks, patch_x0;   // tiles per output row (Wo / 64); patch_w = 32 (f32) | 128 (u8) pixels; 5 | 2 chunks
};
constexpr int PATCH_STAGES = 4;
constexpr int PATCH_STAGE_BYTES = 10240;   // >= 5 chunks * 3 * 5 * 32 * 4 B

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (error surfaces at the next sync) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("spef gemm: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Wait for a role that is not on the critical path (epilogue, TMA producer, MMA issuers with look-ahead): back off with
// nanosleep between polls.  try_wait only suspends for ~60 cycles, so a tight poll loop executes ~9 instructions every
// ~70 cycles per waiting warp -- ncu showed these loops taking a third of all issue slots of the fused kernels, competing
// with the arithmetic warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  int it = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if ((++it & 1023) == 0 && clock64() - t0 > 4000000000LL) {
      printf("spef: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart.
// bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64) |
// [46,48) version=1 (Blackwell) | [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ inline uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- shared-memory plan (host and device agree through these helpers) -------------------------------
__host__ __device__ inline int stage_bytes(int block_n) { return A_STAGE_BYTES + block_n * 128; }
__host__ __device__ inline int bias_floats(int N) { return ((N + 63) / 64) * 64 + 256; }
inline size_t smem_bytes(int block_n, int num_stages, int N, int ng) {
  return 1024 /*align slack*/ + (size_t)num_stages * stage_bytes(block_n) + (size_t)ng * 2 * STAGING_BYTES +
         (size_t)bias_floats(N) * 4 + 256 /*barriers + tmem ptr*/;
}
inline int pick_block_n(int N) {
  if (N <= 256) return ((N + 15) / 16) * 16;
  int best = 256, best_pad = 1 << 30;
  const int cands[3] = {256, 192, 128};
  for (int c : cands) {
    int pad = ((N + c - 1) / c) * c;
    if (pad < best_pad) { best_pad = pad; best = c; }
  }
  return best;
}
inline int pick_stages(int block_n, int N, size_t smem_limit, int ng) {
  int s = MAX_STAGES;
  while (s > 2 && smem_bytes(block_n, s, N, ng) > smem_limit) --s;
  return s;
}

template <bool OUT_F32, int NG>
__global__ void __launch_bounds__(128 + 128 * NG, 1)
pw_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int block_n = p.block_n;
  const int nstages = p.num_stages;
  const int sbytes = stage_bytes(block_n);
  uint8_t* stage_base = smem;
  uint8_t* staging = stage_base + (size_t)nstages * sbytes;  // NG x 2 x 16 KB, 1024-aligned (sbytes % 1024 == 0)
  float* bias_s = reinterpret_cast<float*>(staging + (size_t)NG * 2 * STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + bias_floats(p.N));
  uint64_t* full_bar = bars;                       // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;         // [MAX_STAGES]
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;   // [2]
  uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + block_n - 1) / block_n;
  const int num_tiles = m_tiles * n_tiles;
  const int k_chunks = (p.K + BLOCK_K - 1) / BLOCK_K;

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < bias_floats(p.N); i += (int)blockDim.x) bias_s[i] = (i < p.N) ? p.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4 * NG);  // one arrive per epilogue warp
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)sbytes;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = (tile / n_tiles) * BLOCK_M;
        const int n_idx = (tile % n_tiles) * block_n;
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, tx_bytes);
          uint8_t* sa = stage_base + (size_t)stage * sbytes;
          tma_load_2d(smem_u32(sa), &tmA, kc * BLOCK_K, m_idx, fb);
          tma_load_2d(smem_u32(sa + A_STAGE_BYTES), &tmW, kc * BLOCK_K, n_idx, fb);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        const int lt = (tile - (int)blockIdx.x) / (int)gridDim.x;
        if (p.trace && blockIdx.x == 0 && lt < 256) p.trace[lt * 8 + 0] = clock64();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);  // accumulator stages at columns 0 / 256
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          uint8_t* sa = stage_base + (size_t)stage * sbytes;
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sa));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sa + A_STAGE_BYTES));
          const int k_rem = p.K - kc * BLOCK_K;
          const int ksteps = k_rem >= BLOCK_K ? 4 : (k_rem + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {
            // advancing 16 bf16 (32 B) inside the 128-byte swizzle row = +2 in the (addr >> 4) field
            tcgen05_mma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                             (kc > 0 || k > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[stage]));  // frees the smem stage when these MMAs retire
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(smem_u32(&tmem_full_bar[acc]));  // accumulator complete -> epilogue
        if (p.trace && blockIdx.x == 0 && lt < 256) p.trace[lt * 8 + 1] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: NG groups of 4 warps =====================
    // Each group owns two staging boxes and a named barrier; 128-byte column boxes of the accumulator stream are
    // dealt round-robin to the groups, so TMEM loads, the bias/ReLU/residual math, the smem staging and the TMA
    // stores of different groups overlap (one epilogue warp per SM sub-partition cannot hide its own latencies).
    constexpr int COLS_PER_BOX = OUT_F32 ? 32 : 64;   // one 128-byte staging row
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;                  // epilogue group
    const int row = q * 32 + lane;                    // row of the 128-row tile == TMEM lane
    const bool leader = (q == 0 && lane == 0);
    const uint32_t bar_id = 1 + grp;
    uint8_t* my_staging = staging + (size_t)grp * 2 * STAGING_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    uint32_t box_counter = 0;                         // identical sequence in every group
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = (tile / n_tiles) * BLOCK_M;
      const int n_idx = (tile % n_tiles) * block_n;
      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tcgen05_fence_after();
      const int lt = (tile - (int)blockIdx.x) / (int)gridDim.x;
      const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && lt < 256 && leader && grp == 0;
      if (tr) p.trace[lt * 8 + 2] = clock64();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      const int m = m_idx + row;
      for (int c0 = 0; c0 < block_n && n_idx + c0 < p.N; c0 += COLS_PER_BOX, ++box_counter) {
        if ((int)(box_counter % NG) != grp) continue;
        uint32_t v[COLS_PER_BOX];
        {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v0);
          if constexpr (!OUT_F32) {
            uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
            tmem_ld_32x32b_x32(t_row + (uint32_t)(c0 + 32), v1);
          }
        }
        const bool tma_store = (p.store_mode != 0);
        // TMA mode: staging buffer `buf` is free once the TMA store this group issued two boxes ago has read it.
        // Copy-out mode: it is free because every thread passed the barrier of the previous box after finishing the
        // copy-out of the box before that (program order), so no extra wait is needed.
        if (tma_store && leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        tmem_ld_wait();
        if (tr) p.trace[lt * 8 + 3] = clock64();
        if (tma_store) asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const uint32_t sbuf = smem_u32(my_staging + (size_t)buf * STAGING_BYTES);
        const uint32_t sb = sbuf + (uint32_t)row * (tma_store ? 128u : (uint32_t)STAGING_PITCH);
        const uint32_t swz = tma_store ? (uint32_t)(row & 7) : 0u;
        const uint32_t bias_sa = smem_u32(bias_s + n_idx + c0);
#pragma unroll
        for (int h = 0; h < COLS_PER_BOX / 32; ++h) {
          const int cc = c0 + h * 32;  // column offset inside the tile
          float f[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = lds_f4(bias_sa + (uint32_t)(h * 32 + j4 * 4) * 4u);
            f[j4 * 4 + 0] = __uint_as_float(v[h * 32 + j4 * 4 + 0]) + b4.x;
            f[j4 * 4 + 1] = __uint_as_float(v[h * 32 + j4 * 4 + 1]) + b4.y;
            f[j4 * 4 + 2] = __uint_as_float(v[h * 32 + j4 * 4 + 2]) + b4.z;
            f[j4 * 4 + 3] = __uint_as_float(v[h * 32 + j4 * 4 + 3]) + b4.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if constexpr (!OUT_F32) {
            if (p.residual != nullptr && m < p.M) {
              const bf16* rp = p.residual + (size_t)m * p.N + n_idx + cc;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (n_idx + cc + g * 8 < p.N) {
                  float r[8];
                  Vec8<bf16>::load(rp + g * 8, r);
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[g * 8 + e] += r[e];
                }
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // four 16-byte chunks = 32 bf16 columns
              float t8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) t8[e] = f[g * 8 + e];
              const uint32_t chunk = (uint32_t)(h * 4 + g);  // 0..7 inside the 128-byte row
              sts_u4(sb + ((chunk ^ swz) << 4), Vec8<bf16>::pack(t8));
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {  // eight 16-byte chunks = 32 f32 columns
              uint4 o;
              o.x = __float_as_uint(f[g * 4]); o.y = __float_as_uint(f[g * 4 + 1]);
              o.z = __float_as_uint(f[g * 4 + 2]); o.w = __float_as_uint(f[g * 4 + 3]);
              sts_u4(sb + (((uint32_t)g ^ swz) << 4), o);
            }
          }
        }
        if (tr) p.trace[lt * 8 + 4] = clock64();  // math + STS done
        if (tma_store) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to TMA
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (leader) {
            tma_store_2d(&tmD, sbuf, n_idx + c0, m_idx);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          // coalesced copy-out: 8 consecutive threads write one 128-byte row segment of D
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (tr) p.trace[lt * 8 + 5] = clock64();  // barrier passed
          constexpr int ELEMS_PER_PIECE = OUT_F32 ? 4 : 8;
          constexpr int ESZ = OUT_F32 ? 4 : 2;
          const int tg = (int)threadIdx.x - 128 - grp * 128;  // 0..127 inside the group
          uint8_t* outb = reinterpret_cast<uint8_t*>(p.out);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = tg + k * 128;
            const int r = i >> 3, pc = i & 7;
            const int gm = m_idx + r, gcol = n_idx + c0 + pc * ELEMS_PER_PIECE;
            if (gm < p.M && gcol < p.N) {
              const uint4 val = lds_u4(sbuf + (uint32_t)(r * STAGING_PITCH + pc * 16));
              *reinterpret_cast<uint4*>(outb + ((size_t)gm * p.ldd + gcol) * ESZ) = val;
            }
          }
        }
        if (tr) p.trace[lt * 8 + 6] = clock64();  // copy-out / store issue done
        buf ^= 1;
      }
      // all tcgen05.ld of this warp on this accumulator have completed (wait::ld above): hand TMEM back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      if (tr) p.trace[lt * 8 + 7] = clock64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D row-major [rows, cols] tensor, box = {128 bytes of columns, box_rows}, 128-byte swizzle, zero OOB fill.
inline bool make_tmap_2d(EncodeTiledFn fn, CUtensorMap* out, const void* ptr, bool f32, long long rows, long long cols,
                         long long row_pitch_elems, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const size_t esz = f32 ? 4 : 2;
  const cuuint64_t gstride[1] = {(cuuint64_t)(row_pitch_elems * esz)};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace spef
