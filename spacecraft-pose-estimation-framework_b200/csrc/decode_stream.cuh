// Throughput variant of the orientation decode (softmax + sum_b p_b q_b q_b^T + dominant eigenvector) for large batches
// (BASELINE configs[3], the decode sweep).  Same contract as decode_ori_kernel (kernels_post.cuh); reference semantics:
// src/spe/spe_utils.py:75-76 (softmax), src/spe/classification_utils.py:131-147 (decode).
//
// What decode_ori_kernel left on the table at 10^4..10^5 images (20 % of the HBM roof, profiles/r01_bench_decode_sweep.json):
//   * each lane read one float4 of the AoS bin table per bin: 16 cache lines per warp instruction, four times the logits
//     traffic through L1 -> the table is now SoA (one plane per quaternion component, zero padded), staged in shared memory
//     once per CTA (n <= 2048) or chunk by chunk shared by the CTA's 8 images (larger n), and read with conflict-free
//     16-byte loads that bring the same component of four consecutive bins;
//   * per-lane online softmax made every warp take the rescale branch on nearly every step -> the maximum is now
//     warp-uniform: 32 bins per lane are loaded first (8 independent 16-byte loads in flight per lane), one REDUX gives the
//     warp maximum, the running sums are rescaled at most once per 1024 bins, without divergence;
//   * 14 scalar FMUL/FFMA per bin -> two bins per instruction with packed f32x2 (mul / fma / add);
//   * 110 shuffles per image for the 11 f64 sums -> a transposing reduction (16 f64 shuffles);
//   * an f64 Jacobi solve per image -> f32 Jacobi, then one f64 polish step (re-orthogonalised basis, first-order
//     eigenvector correction: error is second order in the f32 residual), 32 images solved lane-parallel; inv(A) by
//     an f64 L D L^T factorisation.
#pragma once
#include "common.cuh"
#include "kernels_post.cuh"
#include "gemm_tcgen05.cuh"  // mbarrier PTX wrappers

namespace spef {
namespace dstream {

constexpr int TC = 2048;     // bins of the table staged in shared memory at a time
constexpr int SUB = 1024;    // bins per warp step: 8 float4 per lane
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// order-preserving map float <-> int32 (so that one REDUX finds the warp maximum)
__device__ __forceinline__ int f2ord(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// sum over the 32 lanes of 16 doubles per lane with 16 f64 shuffles: every step halves the list a lane keeps.
// Returns the total of entry ((lane >> 1) & 15).
__device__ __forceinline__ double warp_transpose_sum16(double (&x)[16], int lane) {
  bool up = lane & 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double send = up ? x[j] : x[j + 8], keep = up ? x[j + 8] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = lane & 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double send = up ? x[j] : x[j + 4], keep = up ? x[j + 4] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = lane & 4;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double send = up ? x[j] : x[j + 2], keep = up ? x[j + 2] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = lane & 2;
  {
    const double send = up ? x[0] : x[1], keep = up ? x[1] : x[0];
    x[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return x[0] + __shfl_xor_sync(0xffffffffu, x[0], 1);
}

// cyclic Jacobi on a symmetric 4x4 in f32: a -> diagonal (eigenvalues), v -> eigenvectors (columns), both to f32 accuracy
__host__ __device__ __forceinline__ void jacobi4_f32(float (&a)[4][4], float (&v)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? 1.f : 0.f;
  for (int sweep = 0; sweep < 10; ++sweep) {
    float off = 0.f, diag = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      diag = fmaf(a[i][i], a[i][i], diag);
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off = fmaf(a[i][j], a[i][j], off);
    }
    // |off| <= 3e-7 |diag|: the f64 polish below corrects the eigenvector to first order, so what is left is second order (1e-13)
    if (!(off > 1e-13f * diag)) break;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const float apq = a[p][q];
        // a zero pivot gives theta = inf -> t = 0, c = 1, s = 0: the identity rotation, no branch needed
#ifdef __CUDA_ARCH__
        // approximate division / square root (MUFU, ~2 ulp): a rotation only has to be a rotation to f32 accuracy (c, s come from
        // the same t) and annihilate a[p][q] approximately -- the sweeps iterate and the f64 polish removes what is left
        const float theta = __fdividef(a[q][q] - a[p][p], 2.f * apq);
        float rt;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(fmaf(theta, theta, 1.f)));
        const float t = copysignf(__frcp_rn(fabsf(theta) + rt), theta);
        const float c = (apq == 0.f) ? 1.f : rsqrtf(fmaf(t, t, 1.f));
#else
        const float theta = (a[q][q] - a[p][p]) / (2.f * apq);
        const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
        const float c = (apq == 0.f) ? 1.f : 1.f / sqrtf(fmaf(t, t, 1.f));
#endif
        const float s = (apq == 0.f) ? 0.f : t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

// tot[11] = S, a00 a01 a02 a03 a11 a12 a13 a22 a23 a33 (sums over the bins of one image).  Returns false (nothing written)
// when A holds a NaN.  q: unit quaternion, scalar part >= 0; hinv (nullable): inv(A), row major.
__host__ __device__ __forceinline__ bool solve_core(const double (&tot)[11], bool is_logits, float (&qout)[4], float* hinv) {
  double A[4][4];
  A[0][0] = tot[1]; A[0][1] = A[1][0] = tot[2]; A[0][2] = A[2][0] = tot[3]; A[0][3] = A[3][0] = tot[4];
  A[1][1] = tot[5]; A[1][2] = A[2][1] = tot[6]; A[1][3] = A[3][1] = tot[7];
  A[2][2] = tot[8]; A[2][3] = A[3][2] = tot[9]; A[3][3] = tot[10];
  if (is_logits) {  // the reference builds A from the softmax output
    const double inv = 1.0 / tot[0];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) A[i][j] *= inv;
  }
  bool bad = is_logits && isnan(tot[0]);
#pragma unroll
  for (int k = 1; k < 11; ++k) bad = bad || isnan(tot[k]);
  if (bad) return false;
  float a[4][4], v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = (float)A[i][j];
  jacobi4_f32(a, v);
  // move the dominant pair into column 0 (static indexing keeps everything in registers)
  float lam[4] = {a[0][0], a[1][1], a[2][2], a[3][3]};
  int best = 0;
  float best_val = lam[0];
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (lam[k] > best_val) { best_val = lam[k]; best = k; }
#pragma unroll
  for (int c = 1; c < 4; ++c) {
    if (best == c) {
      const float tl = lam[0]; lam[0] = lam[c]; lam[c] = tl;
#pragma unroll
      for (int r = 0; r < 4; ++r) { const float tv = v[r][0]; v[r][0] = v[r][c]; v[r][c] = tv; }
    }
  }
  // f64 polish.  V (f32 Jacobi) is orthonormal to ~1e-7: V' = V (I - E/2), E = V^T V - I, is orthonormal to ~1e-14.
  double V[4][4], E[4][4], W[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) V[r][c] = (double)v[r][c];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      double d = (i == j) ? -1.0 : 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) d = fma(V[r][i], V[r][j], d);
      E[i][j] = E[j][i] = d;
    }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double d = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) d = fma(V[r][k], E[k][c], d);
      W[r][c] = fma(-0.5, d, V[r][c]);
    }
  // first-order correction of the dominant eigenvector in the basis W: q = w0 + sum_j (w_j . A w0) / (lambda_0 - lambda_j) w_j
  double u[4], bj[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) u[r] = A[r][0] * W[0][0] + A[r][1] * W[1][0] + A[r][2] * W[2][0] + A[r][3] * W[3][0];
#pragma unroll
  for (int j = 0; j < 4; ++j) bj[j] = W[0][j] * u[0] + W[1][j] * u[1] + W[2][j] * u[2] + W[3][j] * u[3];
  double q[4] = {W[0][0], W[1][0], W[2][0], W[3][0]};
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    const double den = bj[0] - (double)lam[j];
    const double cj = (den != 0.0) ? bj[j] / den : 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = fma(cj, W[r][j], q[r]);
  }
  const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double sc = ((q[0] < 0.0) ? -1.0 : 1.0) / nrm;
#pragma unroll
  for (int r = 0; r < 4; ++r) qout[r] = (float)(q[r] * sc);
  if (hinv != nullptr) {
    // h_inv = np.linalg.inv(a) (classification_utils.py:142).  A is symmetric positive definite and, for sharp pdfs, badly
    // conditioned (1e8): A = L D L^T in f64 (backward stable without pivoting; cofactors cancel catastrophically here),
    // then inv(A) = M^T D^-1 M with M = inv(L).
    double L[4][4], d[4], id[4], M[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double vj = A[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) vj -= L[j][k] * L[j][k] * d[k];
      d[j] = vj;
      id[j] = 1.0 / vj;
#pragma unroll
      for (int i = j + 1; i < 4; ++i) {
        double t = A[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k] * d[k];
        L[i][j] = t * id[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      M[j][j] = 1.0;
#pragma unroll
      for (int i = j + 1; i < 4; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = j; k < i; ++k) t -= L[i][k] * M[k][j];
        M[i][j] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) {
        double t = 0.0;
#pragma unroll
        for (int k = j; k < 4; ++k) t += M[k][i] * M[k][j] * id[k];
        hinv[i * 4 + j] = hinv[j * 4 + i] = (float)t;
      }
  }
  return true;
}

__device__ __forceinline__ void solve_and_store(const double (&tot)[11], bool is_logits, int img, float* __restrict__ quat_out,
                                                float* __restrict__ hinv_out, uint32_t* __restrict__ flags) {
  float q[4];
  if (!solve_core(tot, is_logits, q, hinv_out != nullptr ? hinv_out + (size_t)img * 16 : nullptr)) {
    // classification_utils.py:134-135 raises; the host turns the flag into that ValueError
    if (flags != nullptr) atomicOr(flags + img, 1u);
    const float qn = __int_as_float(0x7fc00000);
    reinterpret_cast<float4*>(quat_out)[img] = make_float4(qn, qn, qn, qn);
    return;
  }
  reinterpret_cast<float4*>(quat_out)[img] = make_float4(q[0], q[1], q[2], q[3]);
}

// The logits of a warp's images arrive through a per-warp ring of RING 4 KB slots filled by 1-D bulk async copies
// (cp.async.bulk + mbarrier complete_tx, issued by lane 0, RING - 1 steps ahead): the bytes in flight per SM do not depend
// on how many registers a thread can spare, and the loads of the next image overlap the reduction and the eigen-solve.
struct Cursor {   // position in a warp's sequence of 1024-bin steps: (group iteration, table chunk, step)
  int g, c0, s0;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

inline size_t smem_bytes(int nw, int ring) {
  return (size_t)4 * TC * sizeof(float) + (size_t)nw * ring * SUB * sizeof(float) + (size_t)nw * 32 * 11 * sizeof(double) + (size_t)nw * ring * 8;
}

// in: [B][ld] f32 rows, 16-byte aligned, n and ld multiples of 4.  tab: [4][tab_ld] f32 SoA (plane c = component c of the
// scalar-first bins), tab_ld a multiple of SUB, zero padded.  Grid: persistent CTAs of NW warps; warp w of CTA c takes images
// (c + i * gridDim.x) * NW + w, i = 0, 1, ...  PRECISE (used when inv(A) is requested: it amplifies errors of A by the
// condition number, 1e8 for sharp pdfs): f32 partial sums are flushed into the f64 sums every 4 bins instead of every 32.
template <int NW, int RING, int PF, bool AMAX, bool PRECISE, bool LOGITS>
__global__ void __launch_bounds__(NW * 32, 1) decode_ori_stream_kernel(const float* __restrict__ in, int ld, int B, int n,
                                                                       const float* __restrict__ tab, int tab_ld,
                                                                       float* __restrict__ soft_out, float* __restrict__ quat_out,
                                                                       float* __restrict__ hinv_out, int* __restrict__ argmax_out,
                                                                       uint32_t* __restrict__ flags) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  extern __shared__ __align__(128) unsigned char dsm[];
  float* stab = reinterpret_cast<float*>(dsm);                                            // [4][TC]
  float* ring = stab + 4 * TC;                                                            // [NW][RING][SUB]
  double* stash = reinterpret_cast<double*>(ring + (size_t)NW * RING * SUB);              // [NW][32][11]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stash + (size_t)NW * 32 * 11);             // [NW][RING]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = cdiv(B, NW);
  const bool whole = tab_ld <= TC;
  auto stage_table = [&](int c0, int cn) {  // cn: multiple of SUB (the padding of tab is zero)
    for (int i = threadIdx.x; i < cn; i += NW * 32) {   // one float4 per component plane per step
      const int c = i / (cn >> 2), o = i % (cn >> 2);
      reinterpret_cast<float4*>(stab + c * TC)[o] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)c * tab_ld + c0) + o);
    }
  };
  const float* my_ring = ring + (size_t)warp * RING * SUB;
  const uint32_t ring_u = tc::smem_u32(my_ring), bar_u = tc::smem_u32(bars + warp * RING);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) tc::mbar_init(bar_u + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // producer cursor: lane 0 copies the step at `pc` into a ring slot; the L2 cursor `lc` runs PF steps further ahead and only
  // prefetches into L2 (cp.async.bulk.prefetch.L2), so that the ring -- whose depth is bounded by shared memory -- is filled
  // from L2 instead of waiting for HBM
  Cursor pc{(int)blockIdx.x, 0, 0}, lc{(int)blockIdx.x, 0, 0};
  auto step_bytes = [&](const Cursor& c) { return (uint32_t)(min(SUB, min(TC, n - c.c0) - c.s0) * 4); };
  auto step_ptr = [&](const Cursor& c) { return in + ((size_t)c.g * NW + warp) * ld + c.c0 + c.s0; };
  auto ended = [&](const Cursor& c) { return c.g >= groups || (long long)c.g * NW + warp >= B; };
  auto advance = [&](Cursor& c) {
    c.s0 += SUB;
    if (c.s0 >= min(TC, n - c.c0)) {
      c.s0 = 0;
      c.c0 += TC;
      if (c.c0 >= n) { c.c0 = 0; c.g += gridDim.x; }
    }
  };
  auto prefetch_l2 = [&]() {
    if (PF == 0 || ended(lc)) return;
    if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(step_ptr(lc)), "r"(step_bytes(lc)) : "memory");
    advance(lc);
  };
  auto produce = [&](int slot) {
    prefetch_l2();
    if (ended(pc)) return;   // this warp's sequence has ended
    if (lane == 0) {
      const uint32_t bytes = step_bytes(pc);
      tc::mbar_arrive_expect_tx(bar_u + 8u * slot, bytes);
      bulk_load_1d(ring_u + (uint32_t)slot * (SUB * 4), step_ptr(pc), bytes, bar_u + 8u * slot);
    }
    advance(pc);
  };
#pragma unroll 1
  for (int s = 0; s < PF; ++s) {   // lc starts PF steps ahead of pc (the first ring fills come straight from HBM)
    if (!ended(lc)) advance(lc);
  }
#pragma unroll
  for (int s = 0; s < RING - 1; ++s) produce(s);
  int slot = 0, fill_slot = RING - 1;
  uint32_t phases = 0;   // bit s = parity to wait for on slot s

  if (whole) {
    stage_table(0, tab_ld);
    __syncthreads();
  }
  double* my_stash = stash + (size_t)warp * 32 * 11;
  auto solve_batch = [&](int i_first, int count) {
    __syncwarp();
    if (lane < count) {
      const int img = (blockIdx.x + (i_first + lane) * gridDim.x) * NW + warp;
      if (img < B) {
        double tot[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) tot[k] = my_stash[lane * 11 + k];
        solve_and_store(tot, LOGITS, img, quat_out, hinv_out, flags);
      }
    }
    __syncwarp();
  };

  const float pad = LOGITS ? -INFINITY : 0.f;
  int it = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
    const int img = g * NW + warp;
    const bool active = img < B;  // warp-uniform
    float M = -INFINITY;
    int AM = 0;
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int c0 = 0; c0 < n; c0 += TC) {
      const int cn = min(TC, n - c0);
      if (!whole) {
        __syncthreads();
        stage_table(c0, min(TC, tab_ld - c0));
        __syncthreads();
      }
      if (!active) continue;
      for (int s0 = 0; s0 < cn; s0 += SUB) {
        // this step's logits: ring slot -> registers, then the slot is refilled with the step RING - 1 ahead
        produce(fill_slot);
        fill_slot = (fill_slot + 1 == RING) ? 0 : fill_slot + 1;
        tc::mbar_wait(bar_u + 8u * slot, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        float4 z[8];
        const float4* zs = reinterpret_cast<const float4*>(my_ring + (size_t)slot * SUB);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int b = s0 + (j * 32 + lane) * 4;
          z[j] = (b < cn) ? zs[j * 32 + lane] : make_float4(pad, pad, pad, pad);
        }
        __syncwarp();
        slot = (slot + 1 == RING) ? 0 : slot + 1;
        if (LOGITS || AMAX) {
          float lm = fmaxf(fmaxf(z[0].x, z[0].y), fmaxf(z[0].z, z[0].w));
#pragma unroll
          for (int j = 1; j < 8; ++j) lm = fmaxf(fmaxf(lm, fmaxf(z[j].x, z[j].y)), fmaxf(z[j].z, z[j].w));
          const float cm = ord2f(__reduce_max_sync(0xffffffffu, f2ord(lm)));
          if (cm > M) {  // warp-uniform; a later step needs a strictly larger value, so the first maximum wins (np.argmax)
            if (AMAX) {
              int cand = 0x7fffffff;
              if (lm == cm) {
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                  const int b = c0 + s0 + (j * 32 + lane) * 4;
                  if (z[j].w == cm) cand = b + 3;
                  if (z[j].z == cm) cand = b + 2;
                  if (z[j].y == cm) cand = b + 1;
                  if (z[j].x == cm) cand = b;
                }
              }
              AM = __reduce_min_sync(0xffffffffu, cand);
            }
            if (LOGITS) {
              const double sc = (double)ex2_approx((M - cm) * LOG2E);  // M = -inf the first time -> 0
#pragma unroll
              for (int k = 0; k < 11; ++k) acc[k] *= sc;
            }
            M = cm;
          }
        }
        const float mb = -M * LOG2E;
        uint64_t P[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) P[k] = 0ull;
        const uint64_t mb2 = f32x2(mb, mb), l2e2 = f32x2(LOG2E, LOG2E);
        auto slice = [&](int j) {   // 128 bins: 4 per lane, two per packed instruction
          uint64_t W01 = f32x2(z[j].x, z[j].y), W23 = f32x2(z[j].z, z[j].w);
          if (LOGITS) {
            // ex2.approx: ~1e-6 relative on the weights moves the eigenvector by < 1e-3 deg (gate 0.05 deg); ori_soft below uses expf
            float t0, t1, t2, t3;
            f32x2_unpack(fma_f32x2(W01, l2e2, mb2), t0, t1);
            f32x2_unpack(fma_f32x2(W23, l2e2, mb2), t2, t3);
            W01 = f32x2(ex2_approx(t0), ex2_approx(t1));
            W23 = f32x2(ex2_approx(t2), ex2_approx(t3));
          }
          const int o = s0 + (j * 32 + lane) * 4;
          const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(stab + 0 * TC + o);
          const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(stab + 1 * TC + o);
          const ulonglong2 q2 = *reinterpret_cast<const ulonglong2*>(stab + 2 * TC + o);
          const ulonglong2 q3 = *reinterpret_cast<const ulonglong2*>(stab + 3 * TC + o);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t W = h ? W23 : W01;
            const uint64_t Q0 = h ? q0.y : q0.x, Q1 = h ? q1.y : q1.x, Q2 = h ? q2.y : q2.x, Q3 = h ? q3.y : q3.x;
            const uint64_t T0 = mul_f32x2(W, Q0), T1 = mul_f32x2(W, Q1), T2 = mul_f32x2(W, Q2), T3 = mul_f32x2(W, Q3);
            P[0] = add_f32x2(P[0], W);
            P[1] = fma_f32x2(T0, Q0, P[1]); P[2] = fma_f32x2(T0, Q1, P[2]); P[3] = fma_f32x2(T0, Q2, P[3]); P[4] = fma_f32x2(T0, Q3, P[4]);
            P[5] = fma_f32x2(T1, Q1, P[5]); P[6] = fma_f32x2(T1, Q2, P[6]); P[7] = fma_f32x2(T1, Q3, P[7]);
            P[8] = fma_f32x2(T2, Q2, P[8]); P[9] = fma_f32x2(T2, Q3, P[9]); P[10] = fma_f32x2(T3, Q3, P[10]);
          }
          if (PRECISE) {
#pragma unroll
            for (int k = 0; k < 11; ++k) {
              float lo, hi;
              f32x2_unpack(P[k], lo, hi);
              acc[k] += (double)lo + (double)hi;
              P[k] = 0ull;
            }
          }
        };
        if (cn - s0 >= SUB) {   // a full step: straight-line code, the loads and ex2 of one slice overlap the FMAs of the previous one
#pragma unroll
          for (int j = 0; j < 8; ++j) slice(j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (s0 + j * 128 >= cn) break;   // warp-uniform: no lane has bins in this or any later 128-bin slice
            slice(j);
          }
        }
        if (!PRECISE) {
#pragma unroll
          for (int k = 0; k < 11; ++k) {  // 32-bin f32 partial sums into the f64 running sums
            float lo, hi;
            f32x2_unpack(P[k], lo, hi);
            acc[k] += (double)(lo + hi);
          }
        }
      }
    }
    if (active) {
      const double r = warp_transpose_sum16(acc, lane);
      if (soft_out != nullptr && LOGITS) {
        const float sf = (float)__shfl_sync(0xffffffffu, r, 0);
        const float4* row4 = reinterpret_cast<const float4*>(in + (size_t)img * ld);
        float4* o4 = reinterpret_cast<float4*>(soft_out + (size_t)img * n);   // n % 4 == 0, base 16-byte aligned (checked by the host)
        for (int i = lane; i < (n >> 2); i += 32) {
          const float4 x = __ldg(row4 + i);
          o4[i] = make_float4(expf(x.x - M) / sf, expf(x.y - M) / sf, expf(x.z - M) / sf, expf(x.w - M) / sf);
        }
      }
      const int k = (lane >> 1) & 15;
      if ((lane & 1) == 0 && k < 11) my_stash[(it & 31) * 11 + k] = r;
      if (AMAX && lane == 0) argmax_out[img] = AM;
    }
    if ((it & 31) == 31) solve_batch(it - 31, 32);
  }
  if (it & 31) solve_batch(it & ~31, it & 31);
}


// ------------------------------------------------------------------------------------------------------------------
// Small histograms (n <= 512 bins, 8 bins per axis in the sweep): half a warp per image.
// ncu on the kernel above at 512 bins (profiles/r02_ncu_decode512.txt): 740 warp instructions per image of which only 170 are
// the per-bin work (exp + 15 packed FP per bin pair + table loads); the rest is per-image overhead that does not shrink with
// n -- ring / cursor bookkeeping, the rescale + flush into the f64 running sums, the 11-value f64 transposing reduction
// (32 SHFL + 60 FSEL + 27 DADD), padding selects of the half-empty 1024-bin step.  Here
//   * lanes 0-15 take image 2p, lanes 16-31 image 2p + 1: every per-image instruction above serves two images;
//   * an image is ONE step (32 bins per lane): no running maximum / rescale, no f64 running sums -- the per-lane sums of 32
//     products are f32, reduced across the 16 lanes in f32 (15 SHFL; relative error ~1e-7, the eigenvector moves by < 1e-4 deg);
//   * the whole table [4][512] is staged once per CTA; three ring slots of one image pair (4 KB) per warp.
// No argmax / softmax / inv(A) outputs: those calls stay on decode_ori_stream_kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int HN = 512;      // bins per image, at most
constexpr int HNW = 16, HRING = 3;
inline size_t half_smem_bytes() { return (size_t)4 * HN * 4 + (size_t)HNW * HRING * 2 * HN * 4 + (size_t)HNW * 32 * 12 * 4 + (size_t)HNW * HRING * 8; }

// sum over the 16 lanes of a half-warp of 16 floats per lane with 15 shuffles; lane hl ends with the total of entry hl
__device__ __forceinline__ float half_transpose_sum16(float (&x)[16], int hl) {
  bool up = hl & 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = up ? x[j] : x[j + 8], keep = up ? x[j + 8] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = hl & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = up ? x[j] : x[j + 4], keep = up ? x[j + 4] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = hl & 2;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = up ? x[j] : x[j + 2], keep = up ? x[j + 2] : x[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  up = hl & 1;
  const float send = up ? x[0] : x[1], keep = up ? x[1] : x[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

template <bool LOGITS>
__global__ void __launch_bounds__(HNW * 32, 1) decode_ori_half_kernel(const float* __restrict__ in, int ld, int B, int n,
                                                                      const float* __restrict__ tab, int tab_ld,
                                                                      float* __restrict__ quat_out, uint32_t* __restrict__ flags) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  extern __shared__ __align__(128) unsigned char dsm[];
  float* stab = reinterpret_cast<float*>(dsm);                                   // [4][HN]
  float* ring = stab + 4 * HN;                                                   // [HNW][HRING][2][HN]
  float* stash = ring + (size_t)HNW * HRING * 2 * HN;                            // [HNW][32][12]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stash + (size_t)HNW * 32 * 12);   // [HNW][HRING]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
  const int pairs = (B + 1) >> 1;
  const int stride = (int)gridDim.x * HNW;            // pair index step of this warp
  const float* my_ring = ring + (size_t)warp * HRING * 2 * HN;
  const uint32_t ring_u = tc::smem_u32(my_ring), bar_u = tc::smem_u32(bars + warp * HRING);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < HRING; ++s) tc::mbar_init(bar_u + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t row_bytes = (uint32_t)n * 4u;
  int pp = (int)blockIdx.x * HNW + warp;              // producer cursor (pair index)
  auto produce = [&](int slot) {
    if (pp < pairs) {
      if (lane == 0) {
        const bool two = 2 * pp + 1 < B;
        const uint32_t bar = bar_u + 8u * slot, dst = ring_u + (uint32_t)slot * (2 * HN * 4);
        const float* src = in + (size_t)(2 * pp) * ld;
        tc::mbar_arrive_expect_tx(bar, two ? 2u * row_bytes : row_bytes);
        bulk_load_1d(dst, src, row_bytes, bar);
        if (two) bulk_load_1d(dst + HN * 4, src + ld, row_bytes, bar);
      }
      pp += stride;
    }
  };
#pragma unroll
  for (int s = 0; s < HRING - 1; ++s) produce(s);
  int slot = 0, fill_slot = HRING - 1;
  uint32_t phases = 0;

  for (int i = threadIdx.x; i < tab_ld; i += HNW * 32) {   // tab_ld <= 2 * HN floats per plane is checked by the host; zero padded
    const int c = i / (tab_ld >> 2), o = i % (tab_ld >> 2);
    if (o * 4 < HN) reinterpret_cast<float4*>(stab + c * HN)[o] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)c * tab_ld) + o);
  }
  __syncthreads();

  float* my_stash = stash + (size_t)warp * 32 * 12;
  auto solve_batch = [&](int it_first, int n_pairs) {   // images of pairs it_first .. it_first + n_pairs - 1 of this warp, one per lane
    __syncwarp();
    if (lane < 2 * n_pairs) {
      const int img = 2 * (((int)blockIdx.x + (it_first + (lane >> 1)) * (int)gridDim.x) * HNW + warp) + (lane & 1);
      if (img < B) {
        double tot[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) tot[k] = (double)my_stash[lane * 12 + k];
        solve_and_store(tot, LOGITS, img, quat_out, nullptr, flags);
      }
    }
    __syncwarp();
  };

  const float pad = LOGITS ? -INFINITY : 0.f;
  const int n4 = n >> 2;
  int it = 0;
  for (int p = (int)blockIdx.x * HNW + warp; p < pairs; p += stride, ++it) {
    produce(fill_slot);
    fill_slot = (fill_slot + 1 == HRING) ? 0 : fill_slot + 1;
    tc::mbar_wait(bar_u + 8u * slot, (phases >> slot) & 1u);
    phases ^= 1u << slot;
    float4 z[8];
    const float4* zs = reinterpret_cast<const float4*>(my_ring + (size_t)slot * 2 * HN + half * HN);
    const bool have = 2 * p + half < B;       // the odd image of the last pair may not exist: its lanes work on padding
    if (n4 == HN / 4 && have) {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = zs[j * 16 + hl];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = (have && j * 16 + hl < n4) ? zs[j * 16 + hl] : make_float4(pad, pad, pad, pad);
    }
    __syncwarp();
    slot = (slot + 1 == HRING) ? 0 : slot + 1;
    uint64_t mb2 = 0ull;
    const uint64_t l2e2 = f32x2(LOG2E, LOG2E);
    if (LOGITS) {
      float lm = fmaxf(fmaxf(z[0].x, z[0].y), fmaxf(z[0].z, z[0].w));
#pragma unroll
      for (int j = 1; j < 8; ++j) lm = fmaxf(fmaxf(lm, fmaxf(z[j].x, z[j].y)), fmaxf(z[j].z, z[j].w));
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));   // NaN logits: fmaxf drops them, ex2(NaN) keeps them
      const float mb = -lm * LOG2E;
      mb2 = f32x2(mb, mb);
    }
    uint64_t P[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) P[k] = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // 64 bins of this image: 4 per lane, two per packed instruction
      uint64_t W01 = f32x2(z[j].x, z[j].y), W23 = f32x2(z[j].z, z[j].w);
      if (LOGITS) {
        float t0, t1, t2, t3;
        f32x2_unpack(fma_f32x2(W01, l2e2, mb2), t0, t1);
        f32x2_unpack(fma_f32x2(W23, l2e2, mb2), t2, t3);
        W01 = f32x2(ex2_approx(t0), ex2_approx(t1));
        W23 = f32x2(ex2_approx(t2), ex2_approx(t3));
      }
      const int o = (j * 16 + hl) * 4;
      const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(stab + 0 * HN + o);
      const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(stab + 1 * HN + o);
      const ulonglong2 q2 = *reinterpret_cast<const ulonglong2*>(stab + 2 * HN + o);
      const ulonglong2 q3 = *reinterpret_cast<const ulonglong2*>(stab + 3 * HN + o);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t W = h ? W23 : W01;
        const uint64_t Q0 = h ? q0.y : q0.x, Q1 = h ? q1.y : q1.x, Q2 = h ? q2.y : q2.x, Q3 = h ? q3.y : q3.x;
        const uint64_t T0 = mul_f32x2(W, Q0), T1 = mul_f32x2(W, Q1), T2 = mul_f32x2(W, Q2), T3 = mul_f32x2(W, Q3);
        P[0] = add_f32x2(P[0], W);
        P[1] = fma_f32x2(T0, Q0, P[1]); P[2] = fma_f32x2(T0, Q1, P[2]); P[3] = fma_f32x2(T0, Q2, P[3]); P[4] = fma_f32x2(T0, Q3, P[4]);
        P[5] = fma_f32x2(T1, Q1, P[5]); P[6] = fma_f32x2(T1, Q2, P[6]); P[7] = fma_f32x2(T1, Q3, P[7]);
        P[8] = fma_f32x2(T2, Q2, P[8]); P[9] = fma_f32x2(T2, Q3, P[9]); P[10] = fma_f32x2(T3, Q3, P[10]);
      }
    }
    float x[16];
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      float lo, hi;
      f32x2_unpack(P[k], lo, hi);
      x[k] = lo + hi;
    }
#pragma unroll
    for (int k = 11; k < 16; ++k) x[k] = 0.f;
    const float r = half_transpose_sum16(x, hl);
    if (hl < 11) my_stash[(2 * (it & 15) + half) * 12 + hl] = r;
    if ((it & 15) == 15) solve_batch(it - 15, 16);
  }
  if (it & 15) solve_batch(it & ~15, it & 15);
}

}  // namespace dstream
}  // namespace spef
