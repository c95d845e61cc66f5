// Micro-benchmark: pure-write and read+write global bandwidth on B200 (sizes the write-dominated layers).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void write_only(uint4* __restrict__ dst, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = make_uint4(v, v + 1, v + 2, v + 3);
}
__global__ void copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}
// 1 read : 6 writes (like 16->96 pointwise conv)
__global__ void expand6(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_src) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n_src * 6; i += stride) { uint4 v = src[i / 6]; v.x += (uint32_t)i; dst[i] = v; }
}
int main() {
  const size_t bytes = 2ull << 30;
  uint4 *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t n = bytes / 16;
  for (int blocks : {148 * 2, 148 * 8, 148 * 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      float ms;
      cudaEventRecord(e0); write_only<<<blocks, 256>>>(b, n, 7); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("write_only blocks %5d: %.1f GB/s\n", blocks, bytes / ms / 1e6);
      cudaEventRecord(e0); copy<<<blocks, 256>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("copy       blocks %5d: %.1f GB/s (read+write)\n", blocks, 2.0 * bytes / ms / 1e6);
      cudaEventRecord(e0); expand6<<<blocks, 256>>>(a, b, n / 6); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("expand 1:6 blocks %5d: %.1f GB/s (read+write)\n", blocks, (7.0 / 6.0) * (n / 6 * 6 * 16.0) / ms / 1e6);
    }
  }
  float ms;
  cudaEventRecord(e0); cudaMemsetAsync(b, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemset: %.1f GB/s\n", bytes / ms / 1e6);
  return 0;
}
