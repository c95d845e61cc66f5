// Post-processing kernels: softmax + soft-classification decode (orientation: dominant eigenvector of
// sum_b p_b q_b q_b^T; position: pdf-weighted mean), pose error / ESA score partial sums, the temporal
// adaptive pdf filter and quaternion sign continuity.
// Reference semantics: src/spe/spe_utils.py:56-159, src/spe/classification_utils.py:113-166,242-285,
// src/tools/evaluation.py:82-85, src/temporal/pdf_compare.py:94-133, src/temporal/inference.py:136-180.
#pragma once
#include "common.cuh"
#include <math.h>

namespace spef {

// ---- row iteration helper: lane-strided float4 when the row is 16-byte aligned, scalar otherwise -----
template <typename F>
__device__ __forceinline__ void for_each_in_row(const float* __restrict__ row, int n, bool vec_ok, int lane, F&& f) {
  if (vec_ok) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const int n4 = n >> 2;
    int i = lane;
    // four independent 16-byte loads in flight per lane before any of them is consumed (memory-level parallelism)
    for (; i + 96 < n4; i += 128) {
      const float4 v0 = r4[i], v1 = r4[i + 32], v2 = r4[i + 64], v3 = r4[i + 96];
      f(i * 4, v0.x, v0.y, v0.z, v0.w, 4);
      f((i + 32) * 4, v1.x, v1.y, v1.z, v1.w, 4);
      f((i + 64) * 4, v2.x, v2.y, v2.z, v2.w, 4);
      f((i + 96) * 4, v3.x, v3.y, v3.z, v3.w, 4);
    }
    for (; i < n4; i += 32) {
      const float4 v = r4[i];
      f(i * 4, v.x, v.y, v.z, v.w, 4);
    }
    for (int i = (n4 << 2) + lane; i < n; i += 32) f(i, row[i], 0.f, 0.f, 0.f, 1);
  } else {
    for (int i = lane; i < n; i += 32) f(i, row[i], 0.f, 0.f, 0.f, 1);
  }
}

// Cyclic Jacobi eigen-solver for a symmetric 4x4 (double).  a is destroyed (diagonal = eigenvalues),
// v receives the eigenvectors as columns.
__host__ __device__ inline void jacobi4(double (&a)[4][4], double (&v)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      diag += a[i][i] * a[i][i];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    }
    if (!(off > 1e-30 * diag)) break;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // A <- A J  (columns p, q)
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // A <- J^T A  (rows p, q)
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------
// Orientation: softmax (spe_utils.py:75-76) + decode (classification_utils.py:131-147).
// One warp streams G images one after the other; every image is read ONCE (online softmax):
//   each lane keeps its own running maximum m and 11 accumulators S = sum w, A = sum w q q^T (10 unique entries) with
//   w = exp(z - m); when a lane meets a larger value it rescales its accumulators by exp(m_old - m_new).  4-bin f32
//   partial sums (up to 16 bins) are flushed into f64 accumulators.  Lanes are combined with M = max m, scale exp(m - M), f64 shuffles.
//   The first-max argmax (np.argmax tie rule) rides along.
// The 4x4 eigen-problems of the G images are then solved in parallel, one image per lane (cyclic Jacobi in f64, dominant
// eigenvector, renormalise, cast) -- with G = 1 every lane solves the same matrix (small batches: latency matters).
// Optional second pass writes ori_soft = exp(z - M) / S.  The eigenvector sign is unspecified in the reference (LAPACK
// geev); we return scalar part >= 0.  qtab: [n] float4 (scalar-first bins).  ld = row pitch of `in` in floats.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_solve_and_store(const double (&acc)[11], bool is_logits, int img, float* __restrict__ quat_out,
                                                       float* __restrict__ hinv_out, uint32_t* __restrict__ flags) {
  const double S = acc[0];
  double a[4][4], v[4][4];
  a[0][0] = acc[1]; a[0][1] = a[1][0] = acc[2]; a[0][2] = a[2][0] = acc[3]; a[0][3] = a[3][0] = acc[4];
  a[1][1] = acc[5]; a[1][2] = a[2][1] = acc[6]; a[1][3] = a[3][1] = acc[7];
  a[2][2] = acc[8]; a[2][3] = a[3][2] = acc[9]; a[3][3] = acc[10];
  if (is_logits) {  // normalise like the reference (A is built from the softmax output)
    const double inv = 1.0 / S;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i][j] *= inv;
  }
  bool bad = false;
#pragma unroll
  for (int k = 1; k < 11; ++k) bad = bad || isnan(acc[k]);
  bad = bad || (is_logits && isnan(S));
  if (bad) {
    if (flags != nullptr) atomicOr(flags + img, 1u);
    const float qn = __int_as_float(0x7fc00000);
    reinterpret_cast<float4*>(quat_out)[img] = make_float4(qn, qn, qn, qn);
    return;
  }
  jacobi4(a, v);
  int best = 0;
  double best_val = a[0][0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    if (a[k][k] > best_val) { best_val = a[k][k]; best = k; }
  }
  double q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = (best == 0) ? v[k][0] : (best == 1) ? v[k][1] : (best == 2) ? v[k][2] : v[k][3];
  const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double sgn = (q[0] < 0.0) ? -1.0 : 1.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = sgn * q[k] / nrm;
  reinterpret_cast<float4*>(quat_out)[img] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
  if (hinv_out != nullptr) {
    // h_inv = A^-1 = V diag(1/lambda) V^T  (classification_utils.py:142); static indexing keeps a, v in registers
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int r = e >> 2, c = e & 3;
      double h = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) h += v[r][k] * v[c][k] / a[k][k];
      hinv_out[(size_t)img * 16 + e] = (float)h;
    }
  }
}

template <int G>
__global__ void __launch_bounds__(128) decode_ori_kernel(const float* __restrict__ in, int ld, int B, int n, int is_logits,
                                                         const float4* __restrict__ qtab, float* __restrict__ soft_out,
                                                         float* __restrict__ quat_out, float* __restrict__ hinv_out,
                                                         int* __restrict__ argmax_out, uint32_t* __restrict__ flags) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  const int lane = threadIdx.x & 31;
  const int img0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * G;
  if (img0 >= B) return;
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  double keep[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) keep[k] = 0.0;

  for (int g = 0; g < G; ++g) {
    const int img = img0 + g;
    if (img >= B) break;  // warp-uniform
    const float* row = in + (size_t)img * ld;
    float m = -INFINITY;      // running maximum of this lane
    int amax = 0x7fffffff;
    double acc[11];           // S, a00 a01 a02 a03 a11 a12 a13 a22 a23 a33
    float part[11];           // f32 partial sums of up to 16 bins, flushed into acc
    int pending = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) { acc[k] = 0.0; part[k] = 0.f; }
    for_each_in_row(row, n, vec_ok, lane, [&](int i, float z0, float z1, float z2, float z3, int cnt) {
      const float z[4] = {z0, z1, z2, z3};
      float lm = z0;
      int li = i;
      if (cnt == 4) {
        if (z1 > lm) { lm = z1; li = i + 1; }
        if (z2 > lm) { lm = z2; li = i + 2; }
        if (z3 > lm) { lm = z3; li = i + 3; }
      }
      if (lm > m) {  // new running maximum on this lane (first occurrence wins ties: indices increase)
        if (is_logits) {
          const float scf = __expf(m - lm);  // m = -inf the first time -> 0
          const double sc = (double)scf;
#pragma unroll
          for (int k = 0; k < 11; ++k) { acc[k] *= sc; part[k] *= scf; }
        }
        m = lm;
        amax = li;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (e < cnt) {
          // ex2.approx path: relative error ~1e-6 on the weights moves the eigenvector by < 1e-3 deg (gate 0.05 deg);
          // the ori_soft output below uses the accurate expf
          const float w = is_logits ? __expf(z[e] - m) : z[e];
          const float4 q = __ldg(qtab + i + e);
          const float w0 = w * q.x, w1 = w * q.y, w2 = w * q.z, w3 = w * q.w;
          part[0] += w;
          part[1] = fmaf(w0, q.x, part[1]); part[2] = fmaf(w0, q.y, part[2]);
          part[3] = fmaf(w0, q.z, part[3]); part[4] = fmaf(w0, q.w, part[4]);
          part[5] = fmaf(w1, q.y, part[5]); part[6] = fmaf(w1, q.z, part[6]);
          part[7] = fmaf(w1, q.w, part[7]); part[8] = fmaf(w2, q.z, part[8]);
          part[9] = fmaf(w2, q.w, part[9]); part[10] = fmaf(w3, q.w, part[10]);
        }
      }
      if (++pending == 4) {
        pending = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) { acc[k] += (double)part[k]; part[k] = 0.f; }
      }
    });
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] += (double)part[k];
    // combine the lanes: global maximum / argmax, rescale, sum
    float M = m;
    int AM = amax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, M, o);
      const int oi = __shfl_xor_sync(0xffffffffu, AM, o);
      if (om > M || (om == M && oi < AM)) { M = om; AM = oi; }
    }
    if (argmax_out != nullptr && lane == 0) argmax_out[img] = AM;
    if (is_logits) {
      const double sc = (double)__expf(m - M);  // lanes that saw nothing (m = -inf) contribute 0; NaN rows stay NaN
#pragma unroll
      for (int k = 0; k < 11; ++k) acc[k] *= sc;
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = warp_sum(acc[k]);

    if (soft_out != nullptr && is_logits) {
      const float sf = (float)acc[0];
      float* orow = soft_out + (size_t)img * n;
      const bool ovec = vec_ok && ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(soft_out) & 15) == 0);
      if (ovec) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        float4* o4 = reinterpret_cast<float4*>(orow);
        for (int i = lane; i < (n >> 2); i += 32) {
          const float4 x = r4[i];
          o4[i] = make_float4(expf(x.x - M) / sf, expf(x.y - M) / sf, expf(x.z - M) / sf, expf(x.w - M) / sf);
        }
      } else {
        for (int i = lane; i < n; i += 32) orow[i] = expf(row[i] - M) / sf;
      }
    }
    if (G == 1) {
      if (lane == 0) decode_solve_and_store(acc, is_logits != 0, img, quat_out, hinv_out, flags);
    } else if (lane == g) {
#pragma unroll
      for (int k = 0; k < 11; ++k) keep[k] = acc[k];
    }
  }
  if (G > 1) {
    const int img = img0 + lane;
    if (lane < G && img < B) decode_solve_and_store(keep, is_logits != 0, img, quat_out, hinv_out, flags);
  }
}

// --------------------------------------------------------------------------------------------------
// Position: softmax (spe_utils.py:77-79) + weighted mean of the bin centres (classification_utils.py:253-265).
// ptab: [n] float4 (x, y, z, 0).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) decode_pos_kernel(const float* __restrict__ in, int ld, int B, int n, int is_logits,
                                                         const float4* __restrict__ ptab, float* __restrict__ soft_out,
                                                         float* __restrict__ pos_out, uint32_t* __restrict__ flags) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  const int lane = threadIdx.x & 31;
  const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (img >= B) return;
  const float* row = in + (size_t)img * ld;
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  float mx = -INFINITY;
  if (is_logits) {
    for_each_in_row(row, n, vec_ok, lane, [&](int, float a, float b, float c, float d, int cnt) {
      mx = fmaxf(mx, a);
      if (cnt == 4) mx = fmaxf(mx, fmaxf(b, fmaxf(c, d)));
    });
    mx = warp_max(mx);
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for_each_in_row(row, n, vec_ok, lane, [&](int i, float a, float b, float c, float d, int cnt) {
    const float z[4] = {a, b, c, d};
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < cnt) {
        const float w = is_logits ? expf(z[e] - mx) : z[e];
        const float4 x = __ldg(ptab + i + e);
        part[0] += w;
        part[1] = fmaf(w, x.x, part[1]); part[2] = fmaf(w, x.y, part[2]); part[3] = fmaf(w, x.z, part[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += (double)part[k];
  });
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  const double S = acc[0];
  if (soft_out != nullptr && is_logits) {
    const float sf = (float)S;
    float* orow = soft_out + (size_t)img * n;
    for (int i = lane; i < n; i += 32) orow[i] = expf(row[i] - mx) / sf;
  }
  if (lane == 0) {
    uint32_t fl = 0;
    if (S == 0.0) fl |= 2u;
    const float x = (float)(acc[1] / S), y = (float)(acc[2] / S), z = (float)(acc[3] / S);
    if (isnan(x) || isnan(y) || isnan(z)) fl |= 4u;
    pos_out[(size_t)img * 3 + 0] = x;
    pos_out[(size_t)img * 3 + 1] = y;
    pos_out[(size_t)img * 3 + 2] = z;
    if (fl && flags != nullptr) atomicOr(flags + img, fl);
  }
}

// --------------------------------------------------------------------------------------------------
// Score (spe_utils.py:119-157): per image, float32 with NumPy's operation order and no FMA contraction:
//   e_t = sqrt(dx^2 + dy^2 + dz^2); e_tn = e_t / |t|; c = |sum q^ q|; c = min(c, 1); e_q = 2 acos(c).
// Block-reduced in f64 and atomically accumulated into sums[0..7] (see spef_b200.h).  per_image [B,2]
// = (e_q in degrees, e_t) as evaluation.py:82-85.  flags (optional): the decode kernels' per-image guard flags of the same
// batch; images whose orientation / position decode raised a guard (the reference's ValueErrors, classification_utils.py:134,
// 253, 262) are COUNTED in sums[6] / sums[7], so that the fused evaluation route can raise like the reference does.
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ qp, const float* __restrict__ tp,
                                                    const float* __restrict__ qt, const float* __restrict__ tt, int B,
                                                    double* __restrict__ sums, float* __restrict__ per_image,
                                                    const uint32_t* __restrict__ flags) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const float dx = __fsub_rn(tt[i * 3 + 0], tp[i * 3 + 0]);
    const float dy = __fsub_rn(tt[i * 3 + 1], tp[i * 3 + 1]);
    const float dz = __fsub_rn(tt[i * 3 + 2], tp[i * 3 + 2]);
    const float e_t = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float tx = tt[i * 3 + 0], ty = tt[i * 3 + 1], tz = tt[i * 3 + 2];
    const float nt = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz)));
    const float e_tn = __fdiv_rn(e_t, nt);
    float d = __fmul_rn(qp[i * 4 + 0], qt[i * 4 + 0]);
    d = __fadd_rn(d, __fmul_rn(qp[i * 4 + 1], qt[i * 4 + 1]));
    d = __fadd_rn(d, __fmul_rn(qp[i * 4 + 2], qt[i * 4 + 2]));
    d = __fadd_rn(d, __fmul_rn(qp[i * 4 + 3], qt[i * 4 + 3]));
    float c = fabsf(d);
    const bool over = c > 1.01f;
    if (c > 1.f) c = 1.f;
    const float e_q = __fmul_rn(2.f, acosf(c));
    s[0] += (double)e_q;
    s[1] += (double)e_tn;
    s[2] += (double)e_t;
    s[3] += 1.0;
    if (over) s[4] += 1.0;
    if (isnan(e_q) || isnan(e_tn)) s[5] += 1.0;
    if (flags != nullptr) {
      const uint32_t f = flags[i];
      if (f & 1u) s[6] += 1.0;          // SPEF_FLAG_ORI_NAN
      if (f & 6u) s[7] += 1.0;          // SPEF_FLAG_POS_ZERO_SUM | SPEF_FLAG_POS_NAN
    }
    if (per_image != nullptr) {
      per_image[i * 2 + 0] = __fdiv_rn(__fmul_rn(e_q, 180.f), 3.14159265358979323846f);
      per_image[i * 2 + 1] = e_t;
    }
  }
  __shared__ double red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = warp_sum(s[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    if (t != 0.0) atomicAdd(sums + threadIdx.x, t);
  }
}

// --------------------------------------------------------------------------------------------------
// Temporal adaptive pdf filter, TemporalPDF.update_pdf with the 'l2' metric (pdf_compare.py:94-133),
// one CTA per stream.  state [S,n] = previous filtered pdf, has_state [S].
//   cur <- cur / sum(cur);  first frame: state = out = cur, d = 0
//   else d = || cur/sum(cur) - state/sum(state) ||_2 ; w = clip(exp(-alpha d), 0, 1)
//        upd = (w n_coef) cur + (1 - w) state ; upd /= sum(upd) ; state = out = upd
// Array arithmetic is float32 like NumPy's (NEP 50: python-float scalars do not promote f32 arrays);
// reductions are accumulated in f64 and rounded to f32.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256) temporal_filter_kernel(const float* __restrict__ cur, int n, float* __restrict__ state,
                                                              int* __restrict__ has_state, float n_coef, float alpha,
                                                              float* __restrict__ out, float* __restrict__ distance) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  __shared__ double red[8];
  const int s = blockIdx.x;
  const float* c = cur + (size_t)s * n;
  float* st = state + (size_t)s * n;
  float* o = out + (size_t)s * n;
  const int had = has_state[s];  // read before the first barrier: thread 0 rewrites it at the end
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += (double)c[i];
  const float s_cur = (float)block_sum_256(acc, red);
  if (!had) {
    for (int i = threadIdx.x; i < n; i += 256) {
      const float v = __fdiv_rn(c[i], s_cur);
      st[i] = v;
      o[i] = v;
    }
    if (threadIdx.x == 0) {
      distance[s] = 0.f;
      has_state[s] = 1;
    }
    return;
  }
  // compute_distance re-normalises both pdfs (pdf_compare.py:47-48)
  double a1 = 0.0, a2 = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    a1 += (double)__fdiv_rn(c[i], s_cur);
    a2 += (double)st[i];
  }
  const float s1 = (float)block_sum_256(a1, red);
  const float s2 = (float)block_sum_256(a2, red);
  double dd = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float p1 = __fdiv_rn(__fdiv_rn(c[i], s_cur), s1);
    const float p2 = __fdiv_rn(st[i], s2);
    const float df = __fsub_rn(p1, p2);
    dd += (double)__fmul_rn(df, df);
  }
  const float d = sqrtf((float)block_sum_256(dd, red));
  float w = expf(__fmul_rn(-alpha, d));
  w = fminf(fmaxf(w, 0.f), 1.f);
  const float wn = __fmul_rn(w, n_coef);
  const float w1 = __fsub_rn(1.f, w);
  double au = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float u = __fadd_rn(__fmul_rn(wn, __fdiv_rn(c[i], s_cur)), __fmul_rn(w1, st[i]));
    o[i] = u;  // unnormalised, same thread re-reads it below
    au += (double)u;
  }
  const float su = (float)block_sum_256(au, red);
  for (int i = threadIdx.x; i < n; i += 256) {
    const float u = __fdiv_rn(o[i], su);
    o[i] = u;
    st[i] = u;
  }
  if (threadIdx.x == 0) distance[s] = d;
}

// Quaternion sign continuity (inference.py:136-144, 173-180): one thread per stream.
__global__ void quat_continuity_kernel(float* __restrict__ quat, float* __restrict__ prev, int* __restrict__ has_prev, int S) {
  pdl_wait();   // launched with programmatic stream serialization (common.cuh): nothing of the previous kernel is read before this
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float4 q = reinterpret_cast<float4*>(quat)[s];
  if (!has_prev[s]) {
    reinterpret_cast<float4*>(prev)[s] = q;
    has_prev[s] = 1;
    return;
  }
  const float4 p = reinterpret_cast<float4*>(prev)[s];
  float dot = __fmul_rn(p.x, q.x);
  dot = __fadd_rn(dot, __fmul_rn(p.y, q.y));
  dot = __fadd_rn(dot, __fmul_rn(p.z, q.z));
  dot = __fadd_rn(dot, __fmul_rn(p.w, q.w));
  if (dot < 0.f) {
    q = make_float4(-q.x, -q.y, -q.z, -q.w);
    reinterpret_cast<float4*>(quat)[s] = q;
  }
  if (fabsf(dot) > 0.5f) reinterpret_cast<float4*>(prev)[s] = q;
}


// --------------------------------------------------------------------------------------------------
// Soft-classification ENCODE (label side):
//   orientation (classification_utils.py:85-111): k_b = exp(-((2 acos(min(1, |q . h_b|)) / pi)^2) / (2 var)), masked bins = 0,
//   p = k / sum k;  position (:218-240): k_b = exp(-|t - x_b|^2 / (2 var)).  The reference computes both in float64 (the labels
//   come from the dataset JSON as float64) and casts the pdf to float32; so do these kernels.  One CTA per label.
//   tab64: [n][4] doubles (quaternion bins, or x, y, z, 0).  flags: SPEF_FLAG_ENC_NAN when the pdf has a NaN (sum == 0).
// --------------------------------------------------------------------------------------------------
template <bool ORI>
__global__ void __launch_bounds__(256) encode_kernel(const double* __restrict__ label, int B, int n, double inv_2var,
                                                     const double* __restrict__ tab64, const uint8_t* __restrict__ masked,
                                                     float* __restrict__ out, uint32_t* __restrict__ flags) {
  const int b = blockIdx.x;
  if (b >= B) return;
  constexpr int LD = ORI ? 4 : 3;
  double l[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < LD; ++k) l[k] = label[(size_t)b * LD + k];
  float* o = out + (size_t)b * n;
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double h0 = tab64[(size_t)i * 4 + 0], h1 = tab64[(size_t)i * 4 + 1], h2 = tab64[(size_t)i * 4 + 2], h3 = tab64[(size_t)i * 4 + 3];
    double k;
    if (ORI) {
      // np.sum(ori * histogram, axis=1): products then pairwise-free left-to-right sum of 4 terms, no FMA contraction
      const double d = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(l[0], h0), __dmul_rn(l[1], h1)), __dmul_rn(l[2], h2)), __dmul_rn(l[3], h3));
      const double c = (fabs(d) < 1.0 || d != d) ? fabs(d) : 1.0;   // np.minimum propagates NaN (fmin would not)
      const double a = 2.0 * acos(c) / 3.141592653589793;
      k = exp(-(a * a) * inv_2var);
      if (masked != nullptr && masked[i]) k = 0.0;
    } else {
      const double dx = l[0] - h0, dy = l[1] - h1, dz = l[2] - h2;
      k = exp(-__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) * inv_2var);
    }
    o[i] = (float)k;   // staged un-normalised; rescaled below (the float64 value is recomputed there to divide in float64)
    part += k;
  }
  __shared__ double red[8];
  __shared__ double total;
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    total = t;
    if (!(t > 0.0) && flags != nullptr) flags[b] |= 16u;   // SPEF_FLAG_ENC_NAN: 0 / 0 (or a NaN label)
  }
  __syncthreads();
  const double t = total;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double h0 = tab64[(size_t)i * 4 + 0], h1 = tab64[(size_t)i * 4 + 1], h2 = tab64[(size_t)i * 4 + 2], h3 = tab64[(size_t)i * 4 + 3];
    double k;
    if (ORI) {
      const double d = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(l[0], h0), __dmul_rn(l[1], h1)), __dmul_rn(l[2], h2)), __dmul_rn(l[3], h3));
      const double a = 2.0 * acos((fabs(d) < 1.0 || d != d) ? fabs(d) : 1.0) / 3.141592653589793;
      k = exp(-(a * a) * inv_2var);
      if (masked != nullptr && masked[i]) k = 0.0;
    } else {
      const double dx = l[0] - h0, dy = l[1] - h1, dz = l[2] - h2;
      k = exp(-__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) * inv_2var);
    }
    o[i] = (float)(k / t);
  }
}

// --------------------------------------------------------------------------------------------------
// Error statistics of evaluation() on the device (src/tools/evaluation.py:16-32, 95-99):
//   out[0] = mean, out[1] = np.std (population, float64), out[2] = np.median, out[3] = mad = median(|x - median|).
// One CTA.  Medians by an exact 4-pass radix select on the order-preserving integer image of the float32 values (np.median
// of an even count averages the two middle elements; NumPy does that in the array's dtype -- float32 here, like the
// reference's lists of float32 scalars converted by np.median).  scratch: n floats for |x - median|.
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_order_key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// k-th smallest (0-based) of x[0..n): block-wide, every thread returns the value
__device__ float block_select(const float* __restrict__ x, int n, int k, uint32_t* hist /*[256] shared*/, uint32_t* sh /*[2] shared*/) {
  uint32_t prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = f32_order_key(x[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      int d = 0;
      for (; d < 256; ++d) {
        if (acc + hist[d] > (uint32_t)kk) break;
        acc += hist[d];
      }
      sh[0] = (uint32_t)d;
      sh[1] = acc;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= 255u << shift;
    kk -= (int)sh[1];
    __syncthreads();
  }
  return f32_from_key(prefix);
}
__device__ float block_median(const float* __restrict__ x, int n, uint32_t* hist, uint32_t* sh) {
  const float hi = block_select(x, n, n / 2, hist, sh);
  if (n & 1) return hi;
  const float lo = block_select(x, n, n / 2 - 1, hist, sh);
  return __fmul_rn(__fadd_rn(lo, hi), 0.5f);   // np.median: mean of the two middle values, float32
}
__global__ void __launch_bounds__(1024) error_stats_kernel(const float* __restrict__ x, int stride, int n, float* __restrict__ scratch,
                                                           float* __restrict__ packed, double* __restrict__ out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh[2];
  __shared__ double red[32];
  __shared__ double mean_s;
  // pack the strided column, mean and population variance in float64 (np.std of float32 data accumulates in float64? no:
  // np.std on a list of float32 scalars builds a float32 array and reduces with float32 pairwise sums; the host wrapper
  // documents the 1e-6 relative tolerance this implies)
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[(size_t)i * stride];
    packed[i] = v;
    s += (double)v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    mean_s = t / (double)n;
  }
  __syncthreads();
  const double mean = mean_s;
  double v2 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)packed[i] - mean;
    v2 += d * d;
  }
  v2 = warp_sum(v2);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[0] = mean;
    out[1] = sqrt(t / (double)n);
  }
  __syncthreads();
  const float med = block_median(packed, n, hist, sh);
  for (int i = threadIdx.x; i < n; i += blockDim.x) scratch[i] = fabsf(__fsub_rn(packed[i], med));
  __syncthreads();
  const float madv = block_median(scratch, n, hist, sh);
  if (threadIdx.x == 0) {
    out[2] = (double)med;
    out[3] = (double)madv;
  }
}

}  // namespace spef
