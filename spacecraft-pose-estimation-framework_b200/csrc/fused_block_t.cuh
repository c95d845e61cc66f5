// Fused InvertedResidual block, "channel-lane" variant (second generation of fused_block.cuh; same reference semantics,
// src/modeling/common/pytorch_layers.py:65-98, same rounding points as the per-layer kernels).
//
// What the clock64 traces of the first fused kernel showed: with the hidden tile staged in shared memory (TMEM -> regs ->
// smem -> regs -> smem) the worker warps retire one instruction per ~4 cycles -- every phase is a chain of LDS / tcgen05.ld
// latencies, and two block-wide phases per chunk need two barriers.  This variant removes the staging:
//
//   expand is computed TRANSPOSED:  D_e^T[hidden channel (128 TMEM lanes), box pixel (<= 128 columns)] = We_chunk * X_tile^T
//     (A = We chunk [128 x Cin] K-major, B = the TMA-loaded x tile [pixels x Cin] K-major -- the same smem tiles as before
//     with the operand roles swapped).
//   A worker thread owns ONE hidden channel (its TMEM lane) for the whole chunk: 9 depthwise weights + 2 biases live in
//     registers, it walks down the tile row by row: tcgen05.ld.x16 of one pixel row -> +bias, ReLU, round to BF16 (packed
//     f32x2 / bf16x2 ops) -> 3-row register window -> 3x3 stride-S depthwise in FP32 -> +bias, ReLU, BF16 -> shared memory.
//     No shared-memory loads and no intra-chunk barrier in the loop; image-border zero padding is a per-row / per-column
//     select.
//   The depthwise output is written as the MN-major (pixel-contiguous) A operand of the project GEMM
//     D_p[pixel (128 lanes), Cout] += A2^T[channel, pixel]^T * Wp_chunk^T, so a thread stores runs of adjacent pixels of its
//     channel (8-byte stores) instead of 2-byte scatters.
//   Chunks hold up to 128 channels spread evenly over the four 32-lane TMEM quarters (a warp can only read its own
//     quarter), so all four SM sub-partitions carry the same load even when Ch is not a multiple of 128; the host builds
//     the permuted / zero-padded We', Wp' and aux arrays (a padding lane computes exact zeros end to end).
//   Two worker groups (4 warps each) alternate (tile, chunk) items over three TMEM expand stages, so the expand MMA of an
//     item is issued one and a half items ahead of its consumer.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_v2.cuh"
#include "dwconv_tma.cuh"
#include "fused_block.cuh"

namespace spef {
namespace fbt {

constexpr int CL = 128;                      // channel slots (TMEM lanes) per chunk
constexpr int MAX_NGT = 3;                   // worker groups (template parameter NG = 2 | 3)
constexpr int GWT = 4;                       // warps per group: one per TMEM lane quarter
constexpr int AUX_ROWS = 11;                 // per chunk: expand bias | dw bias | dw weights[9], each [128] f32
constexpr int AUX_BYTES = AUX_ROWS * CL * 4; // 5632
constexpr int AUX_STRIDE = 6144;
constexpr int A2_BYTES = 32768;              // A2^T [128 ch][128 px] bf16, MN-major SWIZZLE_128B
constexpr int A2_SBO = 1024;                 // 8 channels x 128 B
constexpr int A2_LBO = 16384;                // next 64-pixel block
constexpr int EPI_WARPS = 8;                 // two teams of four epilogue warps (one warp per TMEM lane quarter and team)
constexpr int CTRL_WARPS = EPI_WARPS + 4;    // epilogue warps + TMEM alloc, TMA producer, project issuer, expand issuer
constexpr int MAX_W_STAGES = 8;
constexpr int MAX_ACC = 4;
constexpr int N_PFULL = 8;                   // project-accumulator "full" barriers, one per tile modulo 8 (see the kernel)
constexpr int MAX_PROJ = 4;                  // project accumulator stages in TMEM: the project issuer -> epilogue -> project issuer round trip is
                                             // ~2000 cycles of hand-off latency, so two stages capped small tiles at one per ~1000 cycles

struct FbtParams {
  const bf16* x;       // block input  [B,H,W,Cin]  (residual source)
  bf16* y;             // block output [B,Ho,Wo,Cout]
  const float* aux;    // [n_chunks][AUX_ROWS][128]
  const float* bp;     // project bias [cpad]
  int B, H, W, Cin, Cout, Ho, Wo;
  int TH, TW, THI, TWI, tiles_y, tiles_x;
  int n_px;            // expand MMA N: THI*TWI rounded up to 16 (<= 128)
  int kc_in, n_chunks, cpad;
  int x_stages, w_stages, resident, proj_stages;
  int n_acc, acc_stride, proj_col0, proj_stride;
  int residual;
  int a2_bufs;         // A2 buffers per worker group: 2 when shared memory allows (the workers of item k + 1 then never wait for the
                       // project MMA of item k to have read A2: ncu showed 24 polls per item on that barrier), else 1
  // Strip stacking (stack = 4) for a block with <= 32 hidden channels and no expand conv (t = 1): the four TMEM lane quarters
  // hold the SAME channels of four horizontally adjacent strips of a 4x wider tile.  The "expand" GEMM is the identity routed
  // per strip: the K chunks of the operand pair are the strips (A chunk q = identity in the rows of quarter q, B chunk q =
  // x tile of strip q), so the issuer code is unchanged; the project GEMM runs once per strip (K = 32 = two K steps) into
  // its own accumulator.
  // Generalised (stack = 2, blocks with an expand conv and Cin <= 64): the lanes are split into `stack` blocks of 128 / stack channel
  // slots; block s works on strip s with the SAME chunk of <= 128 / stack hidden channels, so a 144- or 192-channel block needs
  // 3 chunk passes per two tiles instead of 4 (lane utilisation 75 % instead of 56 %).  The A operand of strip s is the 128-row
  // window starting at row (stack - 1 - s) * (128 / stack) of [zeros | We chunk | zeros].
  int stack;           // 1 | 2 | 4
  int kst_stack;       // K steps of one strip's expand MMA (stack > 1): ceil(Cin / 16), or 2 for the identity
  int cx;              // channels of the x tensor (TMA map); == Cin unless stacked
  int proj_sub;        // TMEM columns between the per-strip project accumulators (stack = 4)
  int we_bytes;        // bytes of the expand-weight region of a weight stage: kc_in * 16 KB, or the 224-row window matrix (stack = 4):
                       // rows 96..127 of We' hold the identity, and strip q's A operand is the 128-row window starting at row
                       // 96 - 32 q (its quarter q sees the identity, the other quarters zeros) -- 28 KB instead of 4 x 16 KB
  int wz_bytes;        // stack = 2: the window matrices of all chunks share their zero blocks -- ONE resident region
                       // [Z | C0 | Z | C1 | ... | Z] of 64-row blocks (chunk c, strip 0: rows [C_c | Z], strip 1: rows [Z | C_c]) in front
                       // of the weight stages, which then hold only Wp and the aux rows (we_bytes = 0): 16 KB less for three chunks,
                       // which is what lets a second x stage fit (one x stage left the expand issuer waiting ~3000 cycles per tile
                       // for the TMA load of the next tile: clock64 trace of block 3)
  // Stem fused into the t = 1 block (STEM instantiation; reference: mobilenet_v2.py:252-262 -- features[0] ConvBnAct(3 -> 32, k3, s2) followed by
  // the first InvertedResidual): the stem conv IS an expand conv with K = 27 taps (padded to 32) whose x tile is built on the SM.
  // The TMA warp stages the image patch of a tile ({patch_w columns, 2 THI + 1 rows, 3 planes} of the NCHW image, zero fill outside the
  // image = the stem's padding), four im2col warps (the second epilogue team's slot) write the K-major operand rows [27 taps | 0 x 5] of the
  // four strips -- two strips share a 128-byte row (strip parity = which 64-byte half), so an x stage is 2 x n_px x 128 bytes --
  // and the stem output (FP32, ReLU folded into the workers' max) never leaves the SM.
  int stem;            // 1: STEM instantiation
  int img_u8;          // image dtype of the patch (uint8: taps through the 256-entry table bf16(u8 / 255))
  int patch_w;         // patch columns (104 f32 | 128 u8) and the patch column of input column 2 ox0 - 1 ... see the producer
  int patch_x0;        // pixels between the patch origin and input column 2 * ox0 (4 f32 | 16 u8: the innermost TMA coordinate stays 16-byte aligned)
  int patch_stages, patch_stride;
  int pdl_early;       // trigger the dependent launch at once (common.cuh)
  long long* trace;
};

__host__ __device__ inline int w_stage_bytes(int we_bytes, int cpad) { return we_bytes + 2 * cpad * 128 + AUX_STRIDE; }
__host__ __device__ inline int x_stage_bytes(int kc_in, int n_px) { return kc_in * n_px * 128; }
inline size_t smem_bytes(const FbtParams& p, int ng) {
  return 1024 + (size_t)p.x_stages * x_stage_bytes(p.stem ? 2 : p.kc_in, p.n_px) + (size_t)p.patch_stages * p.patch_stride + (size_t)p.wz_bytes +
         (size_t)p.w_stages * w_stage_bytes(p.we_bytes, p.cpad) +
         (size_t)ng * p.a2_bufs * A2_BYTES + 1024 /*bias: cpad <= 128 floats, padded (STEM: + the uint8 table)*/ + 640 /*barriers*/;
}
// Register budget per role (setmaxnreg only moves registers inside the CTA's own launch allocation; every count is a multiple
// of 8 and each role is a whole warpgroup of four warps).  Two worker groups + two epilogue teams + control = 20 warps x 96 at launch:
//   8 * 160 + 8 * 56 + 4 * 40 = 1888 <= 1920.   (A first version that counted on the SM's unallocated registers deadlocked in setmaxnreg.inc.)
template <int NG> struct RegPlan { static constexpr int WORKER = 160, EPI = 56, CTRL = 40, EPI_STEM = 64, PROD = 56; };   // STEM: 8 * 160 + 4 * 64 + 4 * 56 + 4 * 40 = 1920
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld with the landing registers as in/out operands: the compiler cannot move a use of v[] above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}
// mbarrier wait that lets the hardware suspend the warp (try_wait with a suspend-time hint) instead of polling: the nanosleep
// poll loops of the first version were ~25 % of all issued instructions of the kernel (ncu source page) on the same schedulers
// as the arithmetic warps.  Still bounded: a protocol bug must trap, not hang the GPU box.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
#ifndef SPEF_FBT_EPI_SLEEP
#define SPEF_FBT_EPI_SLEEP 200   // ns between probes of an epilogue warp that owns every eighth tile (ROT)
#endif
#ifndef SPEF_FBT_WAIT_MODE
#define SPEF_FBT_WAIT_MODE 0   // build-time experiment knob: 0 lean try_wait spin, 1 try_wait with a 1 ms suspend hint, 2 nanosleep poll
#endif
__device__ __forceinline__ void mbar_wait_hw(uint32_t bar, uint32_t parity) {
#if SPEF_FBT_WAIT_MODE == 1
  if (mbar_try_wait_hint(bar, parity, 1000000u)) return;
  const long long t0 = clock64();
  int it = 0;
  while (!mbar_try_wait_hint(bar, parity, 1000000u)) {
    if ((++it & 255) == 0 && clock64() - t0 > 4000000000LL) { printf("spef: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
#elif SPEF_FBT_WAIT_MODE == 2
  tc::mbar_wait_relaxed(bar, parity, 64);
#else
  // lean spin: try_wait suspends the warp in hardware for a system-dependent time and wakes on completion; bounded so that a
  // protocol bug traps instead of hanging the GPU box
  uint32_t it = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if (++it > (1u << 22)) __trap();   // (no printf here: twelve inlined wait sites with a printf slow path each cost 5 KB of code in the hot loops)
  }
#endif
}
// Wait of a role with slack (an epilogue warp that owns every eighth tile, the TMA producer several tiles ahead): back off between
// probes.  ncu: eight epilogue warps in the lean spin above executed 36 % of ALL instructions of the kernel, on the schedulers the
// depthwise workers need.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t it = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++it > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts_u2(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u1(uint32_t saddr, uint32_t a) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(a) : "memory");
}
// UMMA shared-memory descriptor, MN-major operand, SWIZZLE_128B: 64 MN-elements (128 B) contiguous, 8 K-rows per 1024-byte
// atom; LBO = byte distance between 64-element MN blocks, SBO = byte distance between 8-row K groups
// (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor, A MN-major (bit 15), B K-major
__host__ __device__ inline uint32_t make_idesc_bf16_amn(int m, int n) { return tc::make_idesc_bf16(m, n) | (1u << 15); }

// two f32 (packed in a b64) -> bf16x2 with ReLU folded into the conversion (cvt.rn.relu: max(x, 0) then round; identical to
// rounding first because rounding is monotonic and round(0) = 0); low half = first element
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
// round(relu(a + bias)) to BF16 for a pair of FP32 accumulators; returns the packed bf16x2 (low half = first element)
__device__ __forceinline__ uint32_t bias_relu_bf16x2(uint32_t a0, uint32_t a1, uint64_t bias2) {
  return cvt_relu_bf16x2(tc::add_f32x2(tc::pack_f32x2(a0, a1), bias2));
}

// S: depthwise stride; TH: output rows of a tile (compile time: the row loop is fully unrolled, so the register window
// rotates by renaming and every shared-memory store address is a constant)
// Measured and NOT kept (round 2, B = 256, after the worker / epilogue / issuer diets):
//   * three worker groups on 3 x 6 stride-2 tiles (96 TMEM columns per stage, four stages): block 2 288 -> 327 us -- a third more
//     items, and the per-item cost of the single-warp issuer roles does not shrink with the tile;
//   * row split (two worker warps per TMEM lane quarter and group, each computing half of the tile's output rows, 16 worker warps
//     at 96 registers): block 2 288 -> 366 us, block 4 156 -> 190 us -- the extra halo rows, spills and issue contention cost more
//     than the halved per-item latency gains.  The worker loop is not simply latency-bound per warp: ncu shows each worker warp
//     issuing 26 % of the time with no dominant stall reason (wait 17 %, TMEM / barrier scoreboard 13 %, not selected 12 %).
template <int S, int TH, int NG, bool EXP, bool STEM = false>
__global__ void __launch_bounds__(32 * (CTRL_WARPS + NG * GWT), 1)
fused_block_t_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                     const __grid_constant__ CUtensorMap tmWp, const FbtParams p) {
  constexpr int GW = GWT;
  constexpr int GT = 32 * GW;
  constexpr int TW = (S == 1) ? 12 : 6;           // output columns of a tile
  constexpr int TWI = (TW - 1) * S + 3;           // 14 | 13 input columns: one tcgen05.ld.x16 per pixel row
  if (p.pdl_early) pdl_launch_dependents();   // the next kernel of the chain may start its own set-up now (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xsb = x_stage_bytes(STEM ? 2 : p.kc_in, p.n_px);
  constexpr int TEAMS = STEM ? 1 : 2;             // epilogue teams (STEM: the second team's warps are the im2col producers)
  const int wsb = w_stage_bytes(p.we_bytes, p.cpad);
  uint8_t* x_s = smem;
  uint8_t* patch_s = x_s + (size_t)p.x_stages * xsb;          // STEM: image patch ring, else empty
  uint8_t* wz_s = patch_s + (size_t)p.patch_stages * p.patch_stride;   // shared-zero expand weights (stack = 2), else empty
  uint8_t* w_s = wz_s + (size_t)p.wz_bytes;
  uint8_t* a2_s = w_s + (size_t)p.w_stages * wsb;             // [NG][a2_bufs][A2_BYTES]
  float* bp_s = reinterpret_cast<float*>(a2_s + (size_t)NG * p.a2_bufs * A2_BYTES);   // [256] (cpad <= 128)
  uint64_t* bars = reinterpret_cast<uint64_t*>(bp_s + 256);
  uint64_t* x_full = bars;                        // [4]
  uint64_t* x_empty = x_full + 4;                 // [4]
  uint64_t* w_full = x_empty + 4;                 // [MAX_W_STAGES]
  uint64_t* w_empty = w_full + MAX_W_STAGES;      // [MAX_W_STAGES]
  uint64_t* acc_full = w_empty + MAX_W_STAGES;    // [MAX_ACC]  expand MMA -> workers
  uint64_t* acc_empty = acc_full + MAX_ACC;       // [MAX_ACC]  workers -> expand MMA
  uint64_t* a2_full = acc_empty + MAX_ACC;        // [NG][2]  workers -> project MMA
  uint64_t* a2_empty = a2_full + 2 * MAX_NGT;     // [NG][2]  project MMA -> workers
  // proj_full is indexed by the TILE (i & 7), not by the TMEM stage: a consumer (epilogue warp / team) only sees every 2nd or
  // 8th tile, and a parity wait on a barrier whose other phases belong to other consumers returns on a stale phase (ABA) -- with
  // eight barriers every barrier has one consumer set, which observes each of its phases; 8 >= MAX_PROJ keeps the issuer from
  // completing a barrier twice before it is consumed
  uint64_t* proj_full = a2_empty + 2 * MAX_NGT;   // [N_PFULL]
  uint64_t* proj_empty = proj_full + N_PFULL;     // [MAX_PROJ]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(proj_empty + MAX_PROJ);
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(tmem_ptr_s + 2);   // [4]  STEM: TMA -> im2col producers
  uint64_t* patch_empty = patch_full + 4;                               // [4]  STEM: im2col producers -> TMA

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int FIRST_EPI_WARP = NG * GW;
  constexpr int WARP_ALLOC = NG * GW + EPI_WARPS, WARP_TMA = WARP_ALLOC + 1, WARP_MMA_P = WARP_ALLOC + 2, WARP_MMA = WARP_ALLOC + 3;
  constexpr int THI = (TH - 1) * S + 3;
  constexpr int N_EPI = (TH * TW + 31) / 32;      // TMEM lane quarters that hold output pixels of a tile
  // Tiles of at most 32 pixels (stride 2: 4 x 6) ROTATE through the four lane quarters: tile i of the CTA puts its pixels at
  // accumulator rows 32 (i & 3) .. (the workers shift the A2 columns they write), so that eight epilogue warps work on eight
  // different tiles at once.  Larger tiles keep row = pixel and split their strips (or alternate tiles) between the two teams.
  constexpr bool ROT = (TH * TW <= 32);
  const int P_in = THI * TWI;
  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const long long num_tiles = (long long)p.B * tiles_per_img;
  const int my_tiles = (int)((num_tiles - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
  const int total = my_tiles * p.n_chunks;
#ifdef SPEF_FBT_TRACE_BUILD   // clock64 trace of CTA 0 (tools_dev/build_variant.sh trace -DSPEF_FBT_TRACE_BUILD): compiled out of the product
  const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
#define FBT_TRACE(n_, slot_) do { if (tr && (n_) < 64) p.trace[(n_) * 16 + (slot_)] = clock64(); } while (0)
#else
#define FBT_TRACE(n_, slot_) do { } while (0)
#endif

  // the WorkIt iterator of fused_block.cuh only needs these fields
  fb::FbParams itp;
  itp.n_chunks = p.n_chunks; itp.x_stages = p.x_stages; itp.w_stages = p.w_stages; itp.resident = p.resident;
  itp.proj_stages = p.proj_stages; itp.n_acc = p.n_acc;

  for (int i = threadIdx.x; i < (STEM ? 128 : 256); i += (int)blockDim.x) bp_s[i] = (i < p.cpad) ? p.bp[i] : 0.f;
  if (STEM) {   // uint8 images: pixel -> bf16(float(u8) / 255.0f), what rounding the reference's float tensor gives (the stem kernel's table)
    uint16_t* lut = reinterpret_cast<uint16_t*>(bp_s + 128);
    for (int i = threadIdx.x; i < 256; i += (int)blockDim.x) {
      const bf16 h = __float2bfloat16_rn(__fdiv_rn((float)i, 255.0f));
      lut[i] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
  if (warp == WARP_TMA && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmWe);
    tc::tma_prefetch_desc(&tmWp);
  }
  if (warp == WARP_MMA && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(tc::smem_u32(&x_full[i]), STEM ? 4 : 1);    // STEM: one arrival per im2col warp
      tc::mbar_init(tc::smem_u32(&x_empty[i]), 1);
      tc::mbar_init(tc::smem_u32(&patch_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&patch_empty[i]), 4);
    }
    for (int i = 0; i < N_PFULL; ++i) tc::mbar_init(tc::smem_u32(&proj_full[i]), 1);
    for (int i = 0; i < MAX_PROJ; ++i) tc::mbar_init(tc::smem_u32(&proj_empty[i]), ROT ? 1 : (p.stack > 1 ? TEAMS * N_EPI : N_EPI));
    for (int i = 0; i < MAX_W_STAGES; ++i) {
      tc::mbar_init(tc::smem_u32(&w_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&w_empty[i]), 1);
    }
    for (int i = 0; i < MAX_ACC; ++i) {
      tc::mbar_init(tc::smem_u32(&acc_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[i]), GW);
    }
    for (int i = 0; i < 2 * NG; ++i) {
      tc::mbar_init(tc::smem_u32(&a2_full[i]), GW);   // one arrival per worker warp of the group
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
  pdl_wait();   // everything above is independent of the previous kernel's output (common.cuh)

  const float rcp_tpi = 1.0f / (float)tiles_per_img, rcp_tx = 1.0f / (float)p.tiles_x;
  auto tile_coords = [&](int i, int& b, int& oy0, int& ox0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    b = fb::fast_div(t, tiles_per_img, rcp_tpi);
    const int r = t - b * tiles_per_img;
    const int ty = fb::fast_div(r, p.tiles_x, rcp_tx);
    oy0 = ty * TH;
    ox0 = (r - ty * p.tiles_x) * TW * p.stack;
  };

  // every role re-sizes its register file at the top of its own branch (ptxas budgets registers per setmaxnreg in program order:
  // one common if / else chain of the three setmaxnreg in front of the roles made it allocate the worker loop with the control
  // warps' 56 registers -- 1300 bytes of spills)
  // The three single-warp roles below walk (tile, chunk) with plain nested loops and carried ring counters: the clock64 traces
  // showed the expand issuer pacing the kernel at ~1300 cycles per item once the workers and the epilogue were out of the way --
  // ~220 dependent instructions per item, most of them the generic work-item iterator and division-based tile coordinates.
  if (warp == WARP_TMA) {
    // ===================== TMA producer =====================
    reg_dec<RegPlan<NG>::CTRL>();
    if (lane == 0) {
      const uint32_t x_tx = (uint32_t)(p.kc_in * P_in * 128);
      const uint32_t w_tx = (uint32_t)(p.we_bytes + 2 * p.cpad * 128 + AUX_BYTES);
      int tb, ty, tx;
      {
        const int t0 = (int)blockIdx.x;
        tb = t0 / tiles_per_img;
        const int r = t0 - tb * tiles_per_img;
        ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
      }
      const int gdim = (int)gridDim.x;
      const int db = gdim / tiles_per_img, dr = gdim - db * tiles_per_img, dty = dr / p.tiles_x, dtx = dr - dty * p.tiles_x;
      int xs = 0, ws = 0;
      uint32_t xph = 0, wph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int oy0 = ty * TH, ox0 = tx * TW * p.stack;
        if (STEM) {
          // image patch of the tile: hidden (= stem output) rows oy0 - 1 .. oy0 + TH read input rows 2 oy0 - 3 .. 2 oy0 + 2 TH + 1
          mbar_wait_sleep(tc::smem_u32(&patch_empty[xs]), xph ^ 1u, 100);
          const uint32_t fbar = tc::smem_u32(&patch_full[xs]);
          tc::mbar_arrive_expect_tx(fbar, (uint32_t)(p.patch_w * (2 * THI + 1) * 3 * (p.img_u8 ? 1 : 4)));
          tc::tma_load_3d_img(tc::smem_u32(patch_s + (size_t)xs * p.patch_stride), &tmX, 2 * ox0 - p.patch_x0, 2 * oy0 - 3, 3 * tb, fbar);
        } else {
        mbar_wait_sleep(tc::smem_u32(&x_empty[xs]), xph ^ 1u, 100);
          const uint32_t fbar = tc::smem_u32(&x_full[xs]);
          tc::mbar_arrive_expect_tx(fbar, x_tx);
          for (int kc = 0; kc < p.kc_in; ++kc)
            dw::tma_load_4d(tc::smem_u32(x_s + (size_t)xs * xsb + (size_t)kc * p.n_px * 128), &tmX, p.stack > 1 ? 0 : kc * 64,
                            (ox0 + (p.stack > 1 ? kc * TW : 0)) * S - 1, oy0 * S - 1, tb, fbar);
        }
        if (!p.resident || i == 0) {
          for (int c = 0; c < p.n_chunks; ++c) {
            if (p.resident) ws = c; else mbar_wait_hw(tc::smem_u32(&w_empty[ws]), wph ^ 1u);
            const uint32_t fbar = tc::smem_u32(&w_full[ws]);
            uint8_t* dst = w_s + (size_t)ws * wsb;
            tc::mbar_arrive_expect_tx(fbar, w_tx + ((p.wz_bytes && c == 0) ? (uint32_t)p.wz_bytes : 0u));
            if (p.wz_bytes) {                       // the shared-zero region rides on chunk 0's barrier: 64-row blocks
              if (c == 0) for (int blk = 0; blk * 8192 < p.wz_bytes; ++blk) tc::tma_load_2d(tc::smem_u32(wz_s + (size_t)blk * 8192), &tmWe, 0, blk * 64, fbar);
            }
            else if (p.stack > 1) tc::tma_load_2d(tc::smem_u32(dst), &tmWe, 0, c * (p.we_bytes >> 7), fbar);   // one box: the window matrix of chunk c
            else for (int kc = 0; kc < p.kc_in; ++kc) tc::tma_load_2d(tc::smem_u32(dst + (size_t)kc * CL * 128), &tmWe, kc * 64, c * CL, fbar);
            uint8_t* wp = dst + (size_t)p.we_bytes;
            tc::tma_load_2d(tc::smem_u32(wp), &tmWp, c * CL, 0, fbar);
            tc::tma_load_2d(tc::smem_u32(wp + (size_t)p.cpad * 128), &tmWp, c * CL + 64, 0, fbar);
            fb::bulk_load_1d(tc::smem_u32(wp + (size_t)2 * p.cpad * 128), p.aux + (size_t)c * AUX_ROWS * CL, AUX_BYTES, fbar);
            if (!p.resident && ++ws == p.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
        if (++xs == (STEM ? p.patch_stages : p.x_stages)) { xs = 0; xph ^= 1u; }
        tx += dtx; ty += dty; tb += db;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        if (ty >= p.tiles_y) { ty -= p.tiles_y; ++tb; }
      }
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ===================== expand MMA issuer: D_e^T[128 ch, n_px] = We_chunk[128, Cin] * X[n_px, Cin]^T =====================
    reg_dec<RegPlan<NG>::CTRL>();
    const uint32_t idesc_e = tc::make_idesc_bf16(128, p.n_px);
    const uint32_t kst_last = (uint32_t)(((p.Cin - (p.kc_in - 1) * 64) + 15) / 16);
    const uint64_t a_base = tc::make_smem_desc_sw128(tc::smem_u32(w_s));
    const uint64_t az_base = tc::make_smem_desc_sw128(tc::smem_u32(wz_s));
    const uint64_t b_base = tc::make_smem_desc_sw128(tc::smem_u32(x_s));
    const uint32_t x_step = (uint32_t)xsb >> 4, w_step = (uint32_t)wsb >> 4;
    const uint32_t xk_step = (uint32_t)(p.n_px * 128) >> 4;
    const uint32_t ls_rows16 = (uint32_t)((CL / p.stack) * 128) >> 4;   // one strip's block of channel slots, in 16-byte descriptor units
    int xs = 0, ws = 0, as = 0, n = 0;
    uint32_t xph = 0, wph = 0, aph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      if (lane == 0) FBT_TRACE(n, 0);
      mbar_wait_hw(tc::smem_u32(&x_full[xs]), xph);
      if (lane == 0) FBT_TRACE(n, 3);
      const uint64_t b0 = b_base + (uint64_t)((uint32_t)xs * x_step);
      for (int c = 0; c < p.n_chunks; ++c, ++n) {
        if (p.resident) { ws = c; if (i == 0) mbar_wait_hw(tc::smem_u32(&w_full[ws]), 0u); }   // resident weights: loaded once per CTA
        else mbar_wait_hw(tc::smem_u32(&w_full[ws]), wph);
        mbar_wait_hw(tc::smem_u32(&acc_empty[as]), aph ^ 1u);
        tc::tcgen05_fence_after();
        if (lane == 0) FBT_TRACE(n, 1);
        const uint64_t a0 = p.wz_bytes ? az_base + (uint64_t)((uint32_t)c * (uint32_t)(2 * 8192 >> 4)) : a_base + (uint64_t)((uint32_t)ws * w_step);
        const uint32_t d0 = tmem_base + (uint32_t)(as * p.acc_stride);
        for (int kc = 0; kc < p.kc_in; ++kc) {
          const uint32_t ksteps = (p.stack > 1) ? (uint32_t)p.kst_stack : ((kc == p.kc_in - 1) ? kst_last : 4u);
          // stacked: "K chunk" kc is strip kc; its A operand is the window at row (stack - 1 - kc) * (128 / stack)
          const uint32_t a_off = (p.stack > 1) ? (uint32_t)(p.stack - 1 - kc) * ls_rows16 : (uint32_t)kc * (uint32_t)(CL * 128 >> 4);
          for (uint32_t ks = 0; ks < ksteps; ++ks)
            fb::mma_elect(d0, a0 + (uint64_t)(a_off + ks * 2u),
                          b0 + (uint64_t)((STEM ? ((uint32_t)kc >> 1) * xk_step + ((uint32_t)kc & 1u) * 4u : (uint32_t)kc * xk_step) + ks * 2u), idesc_e,
                          (kc > 0 || ks > 0) ? 1u : 0u);
        }
        fb::commit_elect(tc::smem_u32(&acc_full[as]));
        if (lane == 0) FBT_TRACE(n, 2);
        if (++as == p.n_acc) { as = 0; aph ^= 1u; }
        if (!p.resident && ++ws == p.w_stages) { ws = 0; wph ^= 1u; }
      }
      fb::commit_elect(tc::smem_u32(&x_empty[xs]));
      if (++xs == p.x_stages) { xs = 0; xph ^= 1u; }
    }
  } else if (warp == WARP_MMA_P) {
    // ===================== project MMA issuer: D_p[128 px, cpad] += A2^T[128 ch, 128 px]^T * Wp_chunk[cpad, 128 ch]^T =====================
    reg_dec<RegPlan<NG>::CTRL>();
    const uint32_t idesc_p = make_idesc_bf16_amn(128, p.cpad);
    const uint64_t a_base = make_smem_desc_mn_sw128(tc::smem_u32(a2_s), A2_LBO, A2_SBO);
    const uint64_t b_base = tc::make_smem_desc_sw128(tc::smem_u32(w_s + (size_t)p.we_bytes));
    const uint32_t w_step = (uint32_t)wsb >> 4;
    const uint32_t wp_half = (uint32_t)(p.cpad * 128) >> 4;
    int ps = 0, ws = 0, g = 0, k = 0, n = 0;        // g: worker group of the item, k: items that group has finished before it
    uint32_t pph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait_hw(tc::smem_u32(&proj_empty[ps]), pph ^ 1u);
      const uint32_t d = tmem_base + (uint32_t)(p.proj_col0 + ps * p.proj_stride);
      for (int c = 0; c < p.n_chunks; ++c, ++n) {
        const int a2i = (p.a2_bufs == 2) ? 2 * g + (k & 1) : g;            // A2 buffer / barrier of this item
        const uint32_t a2ph = (uint32_t)((p.a2_bufs == 2) ? (k >> 1) & 1 : k & 1);
        mbar_wait_hw(tc::smem_u32(&a2_full[a2i]), a2ph);
        tc::tcgen05_fence_after();
        if (lane == 0) FBT_TRACE(n, 4);
        if (p.resident) ws = c;
        const uint64_t a0 = a_base + (uint64_t)((uint32_t)a2i * (uint32_t)(A2_BYTES >> 4));
        const uint64_t b0 = b_base + (uint64_t)((uint32_t)ws * w_step);
        if (p.stack == 1) {
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks)   // K = 128 channel slots: 16 per step = two 8-channel groups (2 * SBO)
            fb::mma_elect(d, a0 + (uint64_t)(ks * (uint32_t)(2 * A2_SBO >> 4)), b0 + (uint64_t)((ks >> 2) * wp_half + (ks & 3u) * 2u), idesc_p,
                          (c > 0 || ks > 0) ? 1u : 0u);
        } else {
          // K steps per strip = 8 / stack (a power of two): strip s = channel slots [s * 128 / stack, ...) -> its own accumulator
          const uint32_t kshift = (p.stack == 2) ? 2u : 1u, kmask = (1u << kshift) - 1u;
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks)
            fb::mma_elect(d + (ks >> kshift) * (uint32_t)p.proj_sub, a0 + (uint64_t)(ks * (uint32_t)(2 * A2_SBO >> 4)),
                          b0 + (uint64_t)((ks >> 2) * wp_half + (ks & 3u) * 2u), idesc_p, (c > 0 || (ks & kmask) > 0) ? 1u : 0u);
        }
        fb::commit_elect(tc::smem_u32(&a2_empty[a2i]));
        if (!p.resident) { fb::commit_elect(tc::smem_u32(&w_empty[ws])); if (++ws == p.w_stages) ws = 0; }
        if (lane == 0) FBT_TRACE(n, 5);
        if (++g == NG) { g = 0; ++k; }
      }
      fb::commit_elect(tc::smem_u32(&proj_full[i & (N_PFULL - 1)]));
      if (++ps == p.proj_stages) { ps = 0; pph ^= 1u; }
    }
  } else if (warp >= FIRST_EPI_WARP && warp < FIRST_EPI_WARP + 4 * TEAMS) {
    // ===================== epilogue: project accumulator -> +bias (+x) -> bf16 -> global =====================
    // ncu (source page) and the clock64 traces showed this role, not the workers, pacing the kernel after the worker diet: one
    // warp per lane quarter ran ~150 dependent instructions per tile (strip) at ~8-10 cycles each, and for the 24-pixel stride-2
    // tiles only ONE of the four warps had pixels at all.  Now: eight warps in two teams; 24-pixel tiles rotate through the lane
    // quarters (ROT) so that every warp owns every eighth tile; larger tiles split their strips (stacked) or alternate (one
    // strip) between the teams; warps whose quarter holds no pixels leave at once; tile coordinates are carried, not divided.
    const int q = warp & 3;
    const int team = (warp - FIRST_EPI_WARP) >> 2;
    if constexpr (STEM) {
      // One team for four strips: ncu's wait sites showed it pacing the first version of the fused stem (project issuer 78 % of its time
      // on proj_empty, the workers 52 % of theirs on a2_empty behind it) with the generic loop below at ~520 instructions per tile and
      // warp.  Specialised: Cout = 16 (one tcgen05.ld.x16 per strip), no skip input, exact tiling in x; the four strip loads are
      // software-pipelined through two landing buffers and the accumulator stage is handed back as soon as the last one is in registers.
      reg_dec<RegPlan<NG>::EPI_STEM>();
      if (q < N_EPI) {
        const int o = q * 32 + lane;
        const int oy_l = o / TW, ox_l = o - oy_l * TW;
        const bool lane_ok = o < TH * TW;
        int tb, ty, tx;
        {
          const int t0 = (int)blockIdx.x;
          tb = t0 / tiles_per_img;
          const int r = t0 - tb * tiles_per_img;
          ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
        }
        const int gstep = (int)gridDim.x;
        const int db = gstep / tiles_per_img, dr = gstep - db * tiles_per_img, dty = dr / p.tiles_x, dtx = dr - dty * p.tiles_x;
        const int row_px = p.Wo, img_px = p.Ho * p.Wo;
        const uint32_t bias_u = tc::smem_u32(bp_s);
        int ps = 0;
        auto emit = [&](const uint32_t (&v)[16], bf16* yp, bool valid) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 b0 = tc::lds_f4(bias_u + (uint32_t)j * 32u), b1 = tc::lds_f4(bias_u + (uint32_t)j * 32u + 16u);
              float f[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                            __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                            __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                            __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
              Vec8<bf16>::store(yp + j * 8, f);
            }
          }
        };
        for (int i = 0; i < my_tiles; ++i) {
          const int gy = ty * TH + oy_l;
          const bool valid = lane_ok && gy < p.Ho;
          bf16* yp = p.y + (long long)(tb * img_px + gy * row_px + tx * (TW * 4) + ox_l) * 16;   // strip st: + st * TW pixels
          mbar_wait_hw(tc::smem_u32(&proj_full[i & (N_PFULL - 1)]), (uint32_t)((i >> 3) & 1));
          tc::tcgen05_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.proj_col0 + ps * p.proj_stride);
          uint32_t va[16], vb[16];
          tmem_ld_32x32b_x16(t_row, va);
          tmem_ld_wait16(va);
          tmem_ld_32x32b_x16(t_row + 16u, vb);
          emit(va, yp, valid);
          tmem_ld_wait16(vb);
          tmem_ld_32x32b_x16(t_row + 32u, va);
          emit(vb, yp + TW * 16, valid);
          tmem_ld_wait16(va);
          tmem_ld_32x32b_x16(t_row + 48u, vb);
          emit(va, yp + 2 * TW * 16, valid);
          tmem_ld_wait16(vb);
          tc::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&proj_empty[ps]));
          emit(vb, yp + 3 * TW * 16, valid);
          if (++ps == p.proj_stages) ps = 0;
          tx += dtx; ty += dty; tb += db;
          if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
          if (ty >= p.tiles_y) { ty -= p.tiles_y; ++tb; }
        }
      }
    } else {
    reg_dec<RegPlan<NG>::EPI>();
    if (ROT || q < N_EPI) {
    const int o = ROT ? lane : q * 32 + lane;     // output pixel of the tile held by this lane's accumulator row
    const int oy_l = o / TW, ox_l = o - oy_l * TW;
    const bool lane_ok = o < TH * TW;
    const int ng8 = p.Cout >> 3;                  // 8-channel groups (every Cout of the network is a multiple of 8)
    // tiles of this warp: i = i0, i0 + step, ... (tile i of the CTA is global tile blockIdx.x + i * gridDim.x)
    const int i0 = ROT ? (q + 4 * team) : (p.stack > 1 ? 0 : team);
    const int step = ROT ? 8 : (p.stack > 1 ? 1 : 2);
    int tb, ty, tx;
    {
      const long long t0 = (long long)blockIdx.x + (long long)i0 * gridDim.x;
      tb = (int)(t0 / tiles_per_img);
      const int r = (int)(t0 - (long long)tb * tiles_per_img);
      ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
    }
    const int gstep = step * (int)gridDim.x;
    const int db = gstep / tiles_per_img, dr = gstep - db * tiles_per_img, dty = dr / p.tiles_x, dtx = dr - dty * p.tiles_x;
    const int row_px = p.Wo, img_px = p.Ho * p.Wo;
    // project accumulator stage of tile i: i % proj_stages -- carried
    int ps = i0 % p.proj_stages;
    const int ps_step = step % p.proj_stages;
    for (int i = i0; i < my_tiles; i += step) {
      const int oy0 = ty * TH, ox0 = tx * TW * p.stack;
      const int gy = oy0 + oy_l;
      const bool row_ok = lane_ok && gy < p.Ho;
      const int pix0 = tb * img_px + gy * row_px + ox0 + ox_l;     // strip 0; strip st is st * TW pixels to the right
      if (ROT) mbar_wait_sleep(tc::smem_u32(&proj_full[i & (N_PFULL - 1)]), (uint32_t)((i >> 3) & 1), SPEF_FBT_EPI_SLEEP);
      else mbar_wait_hw(tc::smem_u32(&proj_full[i & (N_PFULL - 1)]), (uint32_t)((i >> 3) & 1));
      tc::tcgen05_fence_after();
      if (warp == FIRST_EPI_WARP && lane == 0) FBT_TRACE(i * p.n_chunks, 13);
      // one strip and Cout <= 32: the accumulator row is a single tcgen05.ld -- hand the TMEM stage back as soon as it is in registers
      const bool early = (p.stack == 1 && p.Cout <= 32);
      const uint32_t t_tile = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.proj_col0 + ps * p.proj_stride);
      for (int st = (p.stack > 1 ? team : 0); st < p.stack; st += (p.stack > 1 ? TEAMS : 1)) {
        const bool valid = row_ok && (ox0 + st * TW + ox_l) < p.Wo;
        const long long off = (long long)(pix0 + st * TW) * p.Cout;
        bf16* yp = p.y + off;
        const bf16* rp = p.x + off;               // residual blocks: S == 1, Cin == Cout, same pixel
        const uint32_t t_row = t_tile + (uint32_t)(st * p.proj_sub);
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          // skip-connection input of this pixel's 16 channels, requested BEFORE the accumulator load: behind the stores of the
          // previous group the compiler may not hoist these loads (x and y could alias), and one exposed L2 / HBM round trip per
          // 8 channels made a 72-pixel strip of a residual block take ~3400 cycles (clock64 trace of block 3)
          uint4 rres[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
          if (p.residual && valid) {
            rres[0] = __ldg(reinterpret_cast<const uint4*>(rp + c0));
            if (c0 + 8 < p.Cout) rres[1] = __ldg(reinterpret_cast<const uint4*>(rp + c0 + 8));
          }
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + (uint32_t)c0, v);
          tc::tmem_ld_wait();
          if (early && c0 + 16 >= p.Cout) {
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&proj_empty[ps]));
          }
          if (valid) {
            const uint32_t bias_u = tc::smem_u32(bp_s + c0);
            const int nj = min(2, ng8 - (c0 >> 3));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < nj) {
                const float4 b0 = tc::lds_f4(bias_u + (uint32_t)j * 32u), b1 = tc::lds_f4(bias_u + (uint32_t)j * 32u + 16u);
                float f[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                              __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                              __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                              __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
                if (p.residual) {
                  float r[8];
                  Vec8<bf16>::unpack(rres[j], r);
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += r[e];
                }
                Vec8<bf16>::store(yp + c0 + j * 8, f);
              }
            }
          }
        }
      }
      if (!early) {
        tc::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&proj_empty[ps]));
      }
      if (warp == FIRST_EPI_WARP && lane == 0) FBT_TRACE(i * p.n_chunks, 14);
      // next tile of this warp
      ps += ps_step;
      if (ps >= p.proj_stages) ps -= p.proj_stages;
      tx += dtx; ty += dty; tb += db;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++tb; }
    }
    }
    }
  } else if (STEM && warp >= FIRST_EPI_WARP + 4 && warp < FIRST_EPI_WARP + 8) {
    // ===================== STEM: im2col producers (four warps): image patch -> K-major expand operand of the four strips =====================
    // Thread t < THI * TWI owns hidden pixel (r, j) = (t / TWI, t % TWI) of EVERY strip's haloed box: operand row t of the strip, 27 taps
    // k = (ci * 3 + ky) * 3 + kx as bf16 + 5 zeros = 64 bytes = four 16-byte chunks at logical chunk (strip & 1) * 4 .. + 3 of row t in the
    // 128-byte-swizzled K chunk (strip >> 1).  Tap (ci, ky, kx) of strip s is patch[ci][2 r + ky][24 s + 2 j + patch_x0 - 3 + kx]: hidden
    // column ox0 + 12 s - 1 + j reads input columns 2 (ox0 + 12 s - 1 + j) - 1 + kx, and the patch starts at input column 2 ox0 - patch_x0.
    // kx = 1, 2 are an aligned pair (patch_x0 is a multiple of 4).  A hidden pixel outside the stem's output gets whatever the taps give
    // (zeros beyond the image): the workers overwrite the halo columns and drop the halo rows, exactly as for an expand conv.
    reg_dec<RegPlan<NG>::PROD>();
    const int t = (int)threadIdx.x - 32 * (FIRST_EPI_WARP + 4);
    const bool act = t < THI * TWI;
    const int tt = act ? t : 0;
    const int r = tt / TWI, j = tt - r * TWI;
    const uint32_t xrow = tc::smem_u32(x_s) + (uint32_t)tt * 128u, swz = ((uint32_t)tt & 7u) << 4;
    const uint32_t kc1 = (uint32_t)(p.n_px * 128);                      // second K chunk of an x stage (strips 2, 3)
    // The patch geometry is a compile-time constant per image dtype (the host encodes the TMA box from the same numbers), so every
    // tap load is [thread base + immediate]: the first version (run-time row / plane pitch, strip loop not unrolled) ran 318
    // instructions per tile and warp and paced the kernel (ncu: the workers 35 % of their time on acc_full behind it).
    auto run = [&](auto u8tag) {
      constexpr bool U8 = decltype(u8tag)::value;
      constexpr int ESZ = U8 ? 1 : 4, PW = U8 ? 128 : 104, PX0 = U8 ? 16 : 4;
      constexpr int PROW = PW * ESZ, PPLANE = PROW * (2 * THI + 1), SSTEP = 2 * TW * ESZ;
      const uint32_t patch_u = tc::smem_u32(patch_s) + (uint32_t)((2 * r) * PROW + (2 * j + PX0 - 3) * ESZ);
      int ps = 0, xs = 0;
      uint32_t pph = 0, xph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait_hw(tc::smem_u32(&patch_full[ps]), pph);
        mbar_wait_hw(tc::smem_u32(&x_empty[xs]), xph ^ 1u);
        if (act) {
          const uint32_t pb = patch_u + (uint32_t)ps * (uint32_t)p.patch_stride;
          const uint32_t xb = xrow + (uint32_t)xs * (uint32_t)xsb;
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            uint32_t pk[14];
            if constexpr (U8) {
              // bf16(u8 / 255.0f) without the table: u8 -> float exactly (0x4B0000xx is 2^23 + u8), times float(1 / 255), rounded to
              // BF16 -- equal to the table entry bf16(float(u8) / 255.0f) for all 256 values (the two float32 values differ for 126 of
              // them, never after the rounding to BF16: enumerated in tests/test_host.py).  The table cost a second shared-memory load
              // per tap with random bank conflicts: the uint8 stem took 297 us against 192 us with float pixels.
              float f[28];
#pragma unroll
              for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  // kx = 0 sits at an odd byte, kx = 1, 2 are an aligned pair (PX0 - 3 is odd, every pitch even): two loads per row of taps
                  uint32_t p0, p12;
                  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p0) : "r"(pb + (uint32_t)(ci * PPLANE + ky * PROW + st * SSTEP)));
                  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(p12) : "r"(pb + (uint32_t)(ci * PPLANE + ky * PROW + st * SSTEP + 1)));
                  // one PRMT builds the float 2^23 + u8 (0x4B0000xx), one FMA takes 2^23 * c off again: fma(2^23 + u8, c, -(2^23 * c)) rounds
                  // u8 * c once (2^23 * c is exact), i.e. it IS the product
                  constexpr float C255 = 1.0f / 255.0f, OFF = -8388608.0f * C255;
                  const uint32_t m[3] = {__byte_perm(p0, 0x4B000000u, 0x7440), __byte_perm(p12, 0x4B000000u, 0x7440), __byte_perm(p12, 0x4B000000u, 0x7441)};
#pragma unroll
                  for (int kx = 0; kx < 3; ++kx) f[(ci * 3 + ky) * 3 + kx] = __fmaf_rn(__uint_as_float(m[kx]), C255, OFF);
                }
              f[27] = 0.f;
#pragma unroll
              for (int k = 0; k < 14; ++k) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[k]) : "f"(f[2 * k + 1]), "f"(f[2 * k]));
            } else {
              float f[28];
#pragma unroll
              for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  float v0, v1, v2;
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(pb + (uint32_t)(ci * PPLANE + ky * PROW + st * SSTEP)));
                  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v1), "=f"(v2) : "r"(pb + (uint32_t)(ci * PPLANE + ky * PROW + st * SSTEP + 4)));   // 8-byte aligned
                  f[(ci * 3 + ky) * 3 + 0] = v0; f[(ci * 3 + ky) * 3 + 1] = v1; f[(ci * 3 + ky) * 3 + 2] = v2;
                }
              f[27] = 0.f;
#pragma unroll
              for (int k = 0; k < 14; ++k) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[k]) : "f"(f[2 * k + 1]), "f"(f[2 * k]));
            }
            // logical chunk (st & 1) * 4 + c of row tt in K chunk (st >> 1): physical chunk = logical ^ (row & 7)
            const uint32_t dst = xb + (uint32_t)(st >> 1) * kc1;
            const uint32_t half = (uint32_t)(st & 1) << 6;
            tc::sts_u4(dst + ((0u | half) ^ swz), make_uint4(pk[0], pk[1], pk[2], pk[3]));
            tc::sts_u4(dst + ((16u | half) ^ swz), make_uint4(pk[4], pk[5], pk[6], pk[7]));
            tc::sts_u4(dst + ((32u | half) ^ swz), make_uint4(pk[8], pk[9], pk[10], pk[11]));
            tc::sts_u4(dst + ((48u | half) ^ swz), make_uint4(pk[12], pk[13], 0u, 0u));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(tc::smem_u32(&x_full[xs]));
          tc::mbar_arrive(tc::smem_u32(&patch_empty[ps]));
        }
        if (++ps == p.patch_stages) { ps = 0; pph ^= 1u; }
        if (++xs == p.x_stages) { xs = 0; xph ^= 1u; }
      }
    };
    if (p.img_u8) run(std::true_type{}); else run(std::false_type{});
  } else if (warp < NG * GW) {
    // ===================== workers: one hidden channel per thread, TMEM -> depthwise -> A2^T =====================
    // The hidden tensor stays FP32 between the expand GEMM and the depthwise taps (it never leaves the SM, so rounding it to BF16
    // would only cost instructions): h = relu(acc + be) is carried as h' = max(acc, -be) = h - be -- ONE FMNMX per hidden element --
    // and the constant is folded into the depthwise bias on the host (bd' = bd + be * sum(w)).  A pixel outside the image must
    // be h = 0 (zero padding of the hidden tensor), i.e. h' = -be.  The t = 1 block has no expand conv: h = x, nothing to do.
    reg_inc<RegPlan<NG>::WORKER>();
    const int g = warp / GW;
    const int q = warp & 3;                         // TMEM lane quarter
    const int slot = q * 32 + lane;                 // channel slot of this thread = TMEM lane = K index of the project GEMM
    const int tg = (int)threadIdx.x - 32 * g * GW;
    const uint32_t sw4 = (uint32_t)(slot & 7) << 4; // 128-byte swizzle: 16-byte chunk index ^= (channel row & 7)
    const uint32_t a2_u0 = tc::smem_u32(a2_s) + (uint32_t)((slot >> 3) * A2_SBO + (slot & 7) * 128);
    fb::WorkIt w = fb::work_begin();
    for (int s = 0; s < g && w.n < total; ++s) fb::work_next<NG>(w, itp);
    int cur_i = 0, cur_c = -1;
    // tile row / column of tile cur_i of this CTA (global tile blockIdx.x + cur_i * gridDim.x), carried: the division-based
    // tile_coords sat at the head of every item's dependent chain (~150 cycles)
    int ty, tx;
    {
      const int r = (int)blockIdx.x % tiles_per_img;
      ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
    }
    const int dr1 = (int)gridDim.x % tiles_per_img, dty1 = dr1 / p.tiles_x, dtx1 = dr1 - dty1 * p.tiles_x;
    float nbe = 0.f, bd = 0.f, wd[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // A row of the hidden tile in registers, columns de-interleaved: a[i] = column 2i, c[i] = column 2i+1.  S == 1: (a[i], c[i]) are
    // horizontally adjacent pixels; S == 2: horizontally adjacent OUTPUTS read (a[2i], a[2i+1]) and (c[2i], c[2i+1]) -- adjacent
    // registers either way (packed FFMA2 operands).
    struct Row { float a[8]; float c[8]; };
    while (w.n < total) {
      const int n = w.n;
      for (; cur_i < w.i; ++cur_i) {               // (at most NG steps)
        tx += dtx1; ty += dty1;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        if (ty >= p.tiles_y) ty -= p.tiles_y;
      }
      const int oy0 = ty * TH, ox0 = tx * TW * p.stack;
      if (tg == 0) FBT_TRACE(n, 6);
      if (w.c != cur_c || !p.resident) {            // per-channel constants of this chunk (one chunk, resident: read once per CTA)
        mbar_wait_hw(tc::smem_u32(&w_full[w.ws]), (uint32_t)w.wph);
        cur_c = w.c;
        const uint32_t aux_u = tc::smem_u32(w_s + (size_t)w.ws * wsb + (size_t)p.we_bytes + (size_t)2 * p.cpad * 128) + (uint32_t)slot * 4u;
        nbe = lds_f32(aux_u);
        bd = lds_f32(aux_u + CL * 4);
#pragma unroll
        for (int k = 0; k < 9; ++k) wd[k] = lds_f32(aux_u + (uint32_t)((2 + k) * CL * 4));
      }
      const float mv = EXP ? nbe : 0.f;             // value of a hidden pixel outside the image
      // be * (sum of the top / bottom tap row's weights) with the sign that takes it out of bd (zero for the t = 1 block: be = 0)
      const float dtop = EXP ? nbe * ((wd[0] + wd[1]) + wd[2]) : 0.f, dbot = EXP ? nbe * ((wd[6] + wd[7]) + wd[8]) : 0.f;
      const int ox0q = ox0 + ((q * p.stack) >> 2) * TW;      // stacked: quarter q works on strip q * stack / 4 of the tile
      const bool left_ok = (ox0q * S - 1) >= 0;
      const bool right_ok = (ox0q * S - 1 + TWI - 1) < p.W;
      const int gy0 = oy0 * S - 1;
      const int as = w.as;
      const int gsel = (p.a2_bufs == 2) ? 2 * w.g + w.kph : w.g;            // A2 buffer / barrier of this item
      const uint32_t kph = (uint32_t)((p.a2_bufs == 2) ? w.kph2 : w.kph);
      // ROT: tile i of the CTA writes its pixels at A2 columns (= accumulator rows) 32 (i & 3) ..: columns 64.. are the second
      // 64-pixel block (LBO), columns 32..63 are 64 bytes into the 128-byte row, i.e. bit 6 of the swizzled chunk offset
      const uint32_t rq = ROT ? (uint32_t)(cur_i & 3) : 0u;
      const uint32_t a2_u = a2_u0 + (uint32_t)gsel * (uint32_t)A2_BYTES + (rq >> 1) * (uint32_t)A2_LBO;
      const uint32_t swz = sw4 ^ ((rq & 1u) << 6);
      mbar_wait_hw(tc::smem_u32(&acc_full[as]), (uint32_t)w.aph);
      tc::tcgen05_fence_after();
      if (tg == 0) FBT_TRACE(n, 7);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.acc_stride);
      // bookkeeping of the next item now, so that it overlaps the arithmetic below
      for (int s = 0; s < NG && w.n < total; ++s) fb::work_next<NG>(w, itp);

      // raw accumulator row -> h' (see above); the two halo columns -> mv when they lie outside the image.  Rows outside the image are
      // NOT patched here (a uniform branch per row made ptxas fill the whole row with mv first: 13 moves per row): their TMEM
      // content is finite (the TMA zero fill gives acc = 0), and emit_row drops their tap row instead.
      auto convert = [&](const uint32_t (&v)[16], Row& h) {
#pragma unroll
        for (int j = 0; j < TWI; ++j) {
          const float x = __uint_as_float(v[j]);
          const float y = EXP ? max_nan(x, nbe) : x;   // (NaN-propagating like torch.relu)
          if (j & 1) h.c[j >> 1] = y; else h.a[j >> 1] = y;
        }
        if (EXP) {                                 // (t = 1: TMA zero fill outside the image is already h = 0)
          if (!left_ok) h.a[0] = mv;
          if (!right_ok) { if ((TWI - 1) & 1) h.c[(TWI - 1) / 2] = mv; else h.a[(TWI - 1) / 2] = mv; }
        }
      };
      // one tap row (ky) into the accumulator pairs; per output the order is kx = 0, 1, 2 as in the per-layer kernel
      auto tap_row = [&](const Row& h, const float w0s, const float w1s, const float w2s, uint64_t (&acc)[TW / 2]) {
        const uint64_t w0 = f32x2(w0s, w0s);
        const uint64_t w2 = f32x2(w2s, w2s);
        const uint64_t w1 = f32x2(w1s, w1s);
#pragma unroll
        for (int i = 0; i < TW / 2; ++i) {          // outputs x = 2i, 2i+1
          if (S == 1) {
            // inputs x+kx: kx=0 -> (2i, 2i+1) = (a[i], c[i]) adjacent; kx=1 -> (2i+1, 2i+2) = (c[i], a[i+1]) scalar; kx=2 -> (a[i+1], c[i+1])
            acc[i] = fma_f32x2(f32x2(h.a[i], h.c[i]), w0, acc[i]);
            float lo, hi;
            f32x2_unpack(acc[i], lo, hi);
            lo = fmaf(h.c[i], w1s, lo);
            hi = fmaf(h.a[i + 1], w1s, hi);
            acc[i] = fma_f32x2(f32x2(h.a[i + 1], h.c[i + 1]), w2, f32x2(lo, hi));
          } else {
            // inputs 2x+kx: kx=0 -> cols (4i, 4i+2) = (a[2i], a[2i+1]); kx=1 -> (c[2i], c[2i+1]); kx=2 -> (a[2i+1], a[2i+2]) scalar
            acc[i] = fma_f32x2(f32x2(h.a[2 * i], h.a[2 * i + 1]), w0, acc[i]);
            acc[i] = fma_f32x2(f32x2(h.c[2 * i], h.c[2 * i + 1]), w1, acc[i]);
            float lo, hi;
            f32x2_unpack(acc[i], lo, hi);
            lo = fmaf(h.a[2 * i + 1], w2s, lo);
            hi = fmaf(h.a[2 * i + 2], w2s, hi);
            acc[i] = f32x2(lo, hi);
          }
        }
      };
      // The tile's outputs of this channel as BF16 pairs (word k = pixels 2k, 2k+1 of the tile in row-major order), stored in
      // whole 16-byte chunks (8 adjacent pixels): the lanes of a warp are 32 channel rows 128 bytes apart whose chunks the swizzle
      // spreads over all banks, so a v4 store is four full 128-byte wavefronts (the 4- and 8-byte stores of the first version
      // replayed 4x: ncu counted 59 % of the kernel's shared wavefronts as bank conflicts).
      constexpr int NWORDS = TH * TW / 2;
      uint32_t pk[NWORDS] = {};
      auto flush = [&](int k0) {                     // words k0 .. k0+3 = pixels 2 k0 .. 2 k0 + 7
        const uint32_t op = (uint32_t)(2 * k0);
        const uint32_t u = (op >> 6) * (uint32_t)A2_LBO + (((op & 63u) >> 3) << 4);
        tc::sts_u4(a2_u + (u ^ swz), make_uint4(pk[k0], pk[k0 + 1], pk[k0 + 2], pk[k0 + 3]));
      };
      // output row y (compile-time) of the tile from three hidden rows
      auto emit_row = [&](int y, const Row& r0, const Row& r1, const Row& r2) {
        // zero padding in y: when the top (only y = 0 of the image's first tile row) or the bottom tap row lies outside the image its
        // hidden pixels are h = 0, i.e. the tap row contributes nothing: its three weights are dropped and be * (their sum), which the
        // host folded into bd, is taken out again (dtop / dbot).  Rows further outside only feed outputs the epilogue discards.
        float by = bd, t0 = wd[0], t1 = wd[1], t2 = wd[2], u0 = wd[6], u1 = wd[7], u2 = wd[8];
        {
          if (y == 0 && gy0 < 0) { t0 = 0.f; t1 = 0.f; t2 = 0.f; by += dtop; }
          if (gy0 + S * y + 2 >= p.H) { u0 = 0.f; u1 = 0.f; u2 = 0.f; by += dbot; }
        }
        uint64_t acc[TW / 2];
#pragma unroll
        for (int i = 0; i < TW / 2; ++i) acc[i] = f32x2(by, by);
        tap_row(r0, t0, t1, t2, acc);
        tap_row(r1, wd[3], wd[4], wd[5], acc);
        tap_row(r2, u0, u1, u2, acc);
#pragma unroll
        for (int i = 0; i < TW / 2; ++i) pk[y * (TW / 2) + i] = cvt_relu_bf16x2(acc[i]);
#pragma unroll
        for (int k0 = 0; k0 + 4 <= NWORDS; k0 += 4)   // chunks completed by this row
          if (k0 + 4 > y * (TW / 2) && k0 + 4 <= (y + 1) * (TW / 2)) flush(k0);
        if (y == TH - 1 && (NWORDS & 3)) {            // tail of the tile (TH * TW not a multiple of 8): 4-byte stores
#pragma unroll
          for (int k = NWORDS & ~3; k < NWORDS; ++k) {
            const uint32_t op = (uint32_t)(2 * k);
            const uint32_t u = (op >> 6) * (uint32_t)A2_LBO + (((op & 63u) >> 3) << 4) + (op & 7u) * 2u;
            sts_u1(a2_u + (u ^ swz), pk[k]);
          }
        }
      };

      // Row pipeline: the TMEM load of row r + 1 is in flight while row r is converted and its output row computed (two 16-register
      // landing buffers; tcgen05.wait::ld waits for every outstanding load, so the next load is issued right after the wait).
      Row win[3];
      uint32_t va[16], vb[16];
      tmem_ld_32x32b_x16(t_row, va);
#pragma unroll
      for (int r = 0; r < THI; ++r) {
        if (r & 1) tmem_ld_wait16(vb); else tmem_ld_wait16(va);
        if (r + 1 < THI) {
          if (r & 1) tmem_ld_32x32b_x16(t_row + (uint32_t)((r + 1) * TWI), va);
          else tmem_ld_32x32b_x16(t_row + (uint32_t)((r + 1) * TWI), vb);
        } else {
          // every tcgen05.ld of this warp on the stage has completed -> hand the TMEM stage back now
          tc::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[as]));
        }
        // window slot of row r.  S == 1: output y reads rows y, y+1, y+2 -> slot r % 3.  S == 2: output y reads rows 2y, 2y+1, 2y+2;
        // odd rows -> slot 1, even rows alternate between slots 0 and 2
        const int sl = (S == 1) ? (r % 3) : ((r & 1) ? 1 : ((r >> 1) & 1) * 2);
        if (r & 1) convert(vb, win[sl]); else convert(va, win[sl]);
        if (r == 2) mbar_wait_hw(tc::smem_u32(&a2_empty[gsel]), kph ^ 1u);   // project MMA of this group's previous use of the A2 buffer
        if (S == 1) {
          if (r >= 2) { const int y = r - 2; emit_row(y, win[y % 3], win[(y + 1) % 3], win[(y + 2) % 3]); }
        } else {
          if (r >= 2 && !(r & 1)) { const int y = (r - 2) >> 1; emit_row(y, win[(y & 1) * 2], win[1], win[((y + 1) & 1) * 2]); }
        }
      }
      if (tg == 0) FBT_TRACE(n, 8);
      // A2 (generic-proxy writes) -> visible to the tensor core; every warp arrives for itself (the group barrier + single arrival
      // of the first version kept each warp ~500 cycles per item waiting for the slowest of the four: clock64 trace)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&a2_full[gsel]));
      if (tg == 0) FBT_TRACE(n, 12);
    }
  }
  else if (warp == WARP_ALLOC) {
    reg_dec<RegPlan<NG>::CTRL>();
  }
#undef FBT_TRACE
  // ---- teardown ----
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == WARP_ALLOC) {
    tc::tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tc::TMEM_COLS) : "memory");
  }
}

}  // namespace fbt
}  // namespace spef
