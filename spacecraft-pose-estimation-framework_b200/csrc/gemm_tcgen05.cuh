// Shared pieces of the tcgen05 kernels (PTX wrappers, UMMA descriptors, GemmParams, tensor-map encoder); the kernel itself is
// pw_gemm_tcgen05_v2_kernel in gemm_tcgen05_v2.cuh (the first-generation kernel that used to live here was removed in round 2).
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
  int pdl_early;       // trigger the dependent launch at once (common.cuh)
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
