// Last 1x1 conv of the backbone fused with the global average pool in front of the heads:
//   pooled[b, c] = mean over the H x W pixels of  relu(Wc . x[b, pixel, :] + bias_c)
// (reference: features[18] = ConvBnAct(320 -> 1280, k1) in src/modeling/backbone/mobilenet_v2.py:264, then
//  `x.mean([2, 3])` of the URSONet head, src/modeling/head/ursonet.py).  As two launches the conv writes 63 MB at batch 256 only for
// the pool kernel to read them back; the conv is also one of the two tensor-bound layers of the net.
//
// The GEMM is computed TRANSPOSED, D^T[channel (128 TMEM lanes), pixel (columns)] = W_tile * X^T: the same K-major operands as the
// per-layer kernel with the roles swapped (A = 128 rows of W, B = the pixels of IPT whole images, N = IPT * HW <= 256 columns), so an
// epilogue thread owns ONE output channel and the pool is a sum over its own registers -- no shuffles, no atomics, no shared memory.
// The rounding points of the two-launch path are kept (conv output + bias rounded to BF16, ReLU; pool accumulated in FP32) and so is
// the pool kernel's summation order (eight partial sums over pixels p = w mod 8 in increasing order, added in order, divided by HW:
// global_mean_kernel in kernels_conv.cuh), so the result is bit-identical to the two launches.
//
//   grid = n_ct x cpc CTAs: CTA (ct, j) keeps the K chunks of channel tile ct resident in shared memory (80 KB at K = 320) and walks
//   over the image groups j, j + cpc, ...; the X tiles stream through a TMA ring (they are read by the n_ct CTAs of a group: from L2).
//   warp 0  TMA producer          warp 1  tcgen05.mma issuer (elected lane of the converged warp)      warp 2  TMEM allocation
//   warps 4-11 epilogue (two teams, one image of the pair each): tcgen05.ld of the image's HW columns -> + bias -> BF16 -> ReLU -> FP32
//              partial sums -> mean -> BF16 store
#pragma once
#include "gemm_tcgen05_v2.cuh"

namespace spef {
namespace cpool {

constexpr int EPI_TEAMS = 2;                 // epilogue teams of four warps (one warp per TMEM lane quarter); team t takes the images t, t + 2, ... of an item
constexpr int NT = 128 + 128 * EPI_TEAMS;
constexpr int MAX_X_STAGES = 8;
constexpr int ACC_STAGES = 2, ACC_STRIDE = 256;

struct ConvPoolParams {
  const float* bias;   // [C] f32
  bf16* out;           // pooled [B, C]
  int B, HW, K, C;     // images, pixels per image, input channels, output channels (multiple of 128)
  int ipt;             // images per item: N = ipt * HW columns (multiple of 16, <= 256)
  int n_ct, cpc;       // channel tiles (C / 128), CTAs per channel tile
  int x_stages;
  int relu;
  int pdl_early;       // trigger the dependent launch at once (common.cuh)
};

__host__ __device__ inline int w_bytes(int K) { return ((K + 63) / 64) * 128 * 128; }
__host__ __device__ inline int x_stage_bytes(const ConvPoolParams& p) { return ((p.ipt * p.HW * 128 + 1023) / 1024) * 1024; }
inline size_t smem_bytes(const ConvPoolParams& p) { return 1024 + (size_t)w_bytes(p.K) + (size_t)p.x_stages * x_stage_bytes(p) + 256; }

__global__ void __launch_bounds__(NT, 1)
conv_pool_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const ConvPoolParams p) {
  using namespace tc;
  if (p.pdl_early) pdl_launch_dependents();   // the next kernel of the chain may start its own set-up now (common.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int k_chunks = (p.K + 63) / 64;
  const int xsb = x_stage_bytes(p);
  uint8_t* w_s = smem;
  uint8_t* x_s = smem + w_bytes(p.K);
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_s + (size_t)p.x_stages * xsb);
  uint64_t* x_full = bars;                       // [MAX_X_STAGES]  TMA -> MMA
  uint64_t* x_empty = x_full + MAX_X_STAGES;     // [MAX_X_STAGES]  MMA -> TMA
  uint64_t* acc_full = x_empty + MAX_X_STAGES;   // [ACC_STAGES]    MMA -> epilogue
  uint64_t* acc_empty = acc_full + ACC_STAGES;   // [ACC_STAGES]    epilogue -> MMA
  uint64_t* w_bar = acc_empty + ACC_STAGES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = (int)blockIdx.x % p.n_ct, j0 = (int)blockIdx.x / p.n_ct;
  const int n_items = (p.B + p.ipt - 1) / p.ipt;
  const int n_cols = p.ipt * p.HW;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmX); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.x_stages; ++i) { mbar_init(smem_u32(&x_full[i]), 1); mbar_init(smem_u32(&x_empty[i]), 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4 * EPI_TEAMS); }
    mbar_init(smem_u32(w_bar), 1);
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
      const uint32_t wb = smem_u32(w_bar);
      mbar_arrive_expect_tx(wb, (uint32_t)w_bytes(p.K));
      for (int kc = 0; kc < k_chunks; ++kc) tma_load_2d(smem_u32(w_s + (size_t)kc * 16384), &tmW, kc * 64, ct * 128, wb);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(n_cols * 128);
      for (int it = j0; it < n_items; it += p.cpc) {
        const int row0 = it * n_cols;   // rows past B * HW (odd batch) are zero-filled by TMA
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait_relaxed(smem_u32(&x_empty[stage]), phase ^ 1, 64);
          const uint32_t fb = smem_u32(&x_full[stage]);
          mbar_arrive_expect_tx(fb, tx);
          tma_load_2d(smem_u32(x_s + (size_t)stage * xsb), &tmX, kc * 64, row0, fb);
          if (++stage == p.x_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, n_cols);
    const uint64_t a_base = make_smem_desc_sw128(smem_u32(w_s));
    const uint64_t b_base = make_smem_desc_sw128(smem_u32(x_s));
    const uint32_t x_step = (uint32_t)xsb >> 4;
    const uint32_t kst_last = (uint32_t)((p.K - (k_chunks - 1) * 64 + 15) / 16);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    mbar_wait(smem_u32(w_bar), 0);
    for (int it = j0; it < n_items; it += p.cpc) {
      mbar_wait(smem_u32(&acc_empty[acc]), acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
      for (int kc = 0; kc < k_chunks; ++kc) {
        mbar_wait(smem_u32(&x_full[stage]), phase);
        tcgen05_fence_after();
        const uint64_t a_desc = a_base + (uint64_t)((uint32_t)kc * (16384u >> 4));
        const uint64_t b_desc = b_base + (uint64_t)((uint32_t)stage * x_step);
        const uint32_t ksteps = (kc == k_chunks - 1) ? kst_last : 4u;
        for (uint32_t k = 0; k < ksteps; ++k)
          mma_elect_v2(d_tmem, a_desc + (uint64_t)(k * 2u), b_desc + (uint64_t)(k * 2u), idesc, (kc > 0 || k > 0) ? 1u : 0u);
        commit_elect_v2(smem_u32(&x_empty[stage]));
        if (++stage == p.x_stages) { stage = 0; phase ^= 1; }
      }
      commit_elect_v2(smem_u32(&acc_full[acc]));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: one output channel per thread =====================
    // (first version: one team, a tcgen05.wait::ld after each of the six loads of an item, scalar adds -- ncu: the MMA issuer polled
    //  acc_empty 12x per item, tensor pipe 50 %: the epilogue paced the kernel.  Now two teams take one image of the pair each, the three
    //  loads of an image are in flight together and the bias add / partial sums are packed f32x2 operations.)
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int team = (warp - 4) >> 2;
    const int ch = ct * 128 + q * 32 + lane;
    const float bias = p.bias[ch];
    const uint64_t bias2 = pack_f32x2(__float_as_uint(bias), __float_as_uint(bias));
    const float divisor = (float)p.HW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = j0; it < n_items; it += p.cpc) {
      mbar_wait(smem_u32(&acc_full[acc]), acc_phase);
      tcgen05_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      for (int im = team; im < p.ipt; im += EPI_TEAMS) {
        // partial sums of the pool kernel: part[w] takes the pixels p = w (mod 8) in increasing order; pairs (part[2j], part[2j + 1]) are packed
        uint64_t part2[4] = {0ull, 0ull, 0ull, 0ull};
        auto fold = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            // the per-layer epilogue: acc + bias in FP32, one rounding to BF16, ReLU on the BF16 pair (NaN-propagating)
            uint32_t pk = cvt_bf16x2(add_f32x2(pack_f32x2(v[i], v[i + 1]), bias2));
            if (p.relu) pk = relu_bf16x2(pk);
            part2[(i >> 1) & 3] = add_f32x2(part2[(i >> 1) & 3], pack_f32x2(pk << 16, pk & 0xffff0000u));
          }
        };
        const uint32_t tc0 = t0 + (uint32_t)(im * p.HW);
        int c0 = 0;
        for (; c0 + 96 <= p.HW; c0 += 96) {       // three loads in flight (HW = 96: the whole image)
          uint32_t v0[32], v1[32], v2[32];
          tmem_ld_32x32b_x32(tc0 + (uint32_t)c0, v0);
          tmem_ld_32x32b_x32(tc0 + (uint32_t)(c0 + 32), v1);
          tmem_ld_32x32b_x32(tc0 + (uint32_t)(c0 + 64), v2);
          tmem_ld_wait();
          fold(v0); fold(v1); fold(v2);
        }
        for (; c0 < p.HW; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tc0 + (uint32_t)c0, v);
          tmem_ld_wait();
          fold(v);
        }
        float part[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(part2[k]));
          part[2 * k] = __uint_as_float(lo); part[2 * k + 1] = __uint_as_float(hi);
        }
        float t = part[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += part[k];
        const int b = it * p.ipt + im;
        if (b < p.B) p.out[(size_t)b * p.C + ch] = __float2bfloat16_rn(t / divisor);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[acc]));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace cpool
}  // namespace spef
