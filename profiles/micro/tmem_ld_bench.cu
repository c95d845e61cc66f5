// Micro-benchmark: tcgen05.ld (LDTM) latency / throughput on B200, to size the GEMM epilogue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu && ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: one LDTM.x32 + wait per iteration (latency chain); mode 1: two LDTM.x32 back to back + one wait
__global__ void __launch_bounds__(512, 1) k(int iters, int mode, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t v[32], w[32];
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)((i * 64) & 448);
      ld_x32(base + col, v);
      if (mode == 1) ld_x32(base + col + 32, w);
      ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
      if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= w[j];
      }
    }
    t1 = clock64();
  }
  if (lane == 0 && warp < nwarps) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 16 * sizeof(long long)); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      cudaMemset(out, 0, 148 * 16 * sizeof(long long));
      k<<<148, 512>>>(iters, mode, nw, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[16];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes_per_iter_sm = (double)nw * 32 * 32 * 4 * (mode + 1);
      printf("mode %d warps %2d: %s  %.1f cycles/iter  -> %.1f B/clk/SM TMEM->RF\n", mode, nw, cudaGetErrorString(e),
             (double)mx / iters, bytes_per_iter_sm / ((double)mx / iters));
    }
  return 0;
}
