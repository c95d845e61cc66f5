// Host side of the packed image upload (spef_eval_submit_host with SPEF_HOST_PACK, BF16 engine, float images).
//
// The end-to-end rate of the host-buffer route is the PCIe copy of the float images (1.11 MB each; 98 % of the measured H2D
// roof), while the first thing the stem does with a pixel is `cvt.rn.bf16.f32` (the MMA operand is BF16).  Doing that
// one rounding on the host -- the same round-to-nearest-even, NaN -> quiet NaN, denormals kept -- halves the bytes on the bus and
// changes no result bit: the device widens the BF16 pixels back to float (exact) and the stem's own rounding is then the identity.
// (reference contract: SPETorch.predict moves the float batch itself, src/spe/spe_torch.py:57-61.)
//
// A small persistent thread pool converts one chunk at a time while the DMA engine moves the previous one; conversion streams the
// result past the cache (the next reader is the DMA engine).  Plain C++ compiled by the host compiler (nvcc passes .cpp through).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

namespace spef_host {

// float -> bf16 bits, round to nearest even; identical to the device's cvt.rn.bf16.f32 for every non-NaN input; NaN -> a quiet NaN
static inline uint16_t f2bf(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

static void pack_scalar(const float* src, uint16_t* dst, size_t n) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (size_t i = 0; i < n; ++i) dst[i] = f2bf(s[i]);
}

#if defined(__x86_64__)
#define SPEF_AVX512 __attribute__((target("avx512f,avx512bw")))
SPEF_AVX512 static inline __m512i cvt16(__m512i u) {
  const __m512i hi = _mm512_srli_epi32(u, 16);
  const __m512i r = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(u, _mm512_set1_epi32(0x7fff)), _mm512_and_si512(hi, _mm512_set1_epi32(1))), 16);
  const __mmask16 nan = _mm512_cmpgt_epu32_mask(_mm512_and_si512(u, _mm512_set1_epi32(0x7fffffff)), _mm512_set1_epi32(0x7f800000));
  return _mm512_mask_mov_epi32(r, nan, _mm512_or_si512(hi, _mm512_set1_epi32(0x0040)));
}
template <bool NT>
SPEF_AVX512 static void pack_avx512_t(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  // head: up to a 64-byte aligned destination (streaming stores need it)
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u)) { dst[i] = f2bf(reinterpret_cast<const uint32_t*>(src)[i]); ++i; }
  for (; i + 32 <= n; i += 32) {
    const __m256i lo = _mm512_cvtepi32_epi16(cvt16(_mm512_loadu_si512(src + i))), up = _mm512_cvtepi32_epi16(cvt16(_mm512_loadu_si512(src + i + 16)));
    const __m512i v = _mm512_inserti64x4(_mm512_castsi256_si512(lo), up, 1);
    if (NT) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), v); else _mm512_store_si512(reinterpret_cast<__m512i*>(dst + i), v);
  }
  if (NT) _mm_sfence();
  for (; i < n; ++i) dst[i] = f2bf(reinterpret_cast<const uint32_t*>(src)[i]);
}
SPEF_AVX512 static void pack_avx512(const float* src, uint16_t* dst, size_t n) { pack_avx512_t<true>(src, dst, n); }
SPEF_AVX512 static void pack_avx512_cached(const float* src, uint16_t* dst, size_t n) { pack_avx512_t<false>(src, dst, n); }
#endif

using PackFn = void (*)(const float*, uint16_t*, size_t);
static PackFn pick_pack() {
#if defined(__x86_64__)
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && !getenv("SPEF_PACK_SCALAR"))
    return getenv("SPEF_PACK_CACHED") ? pack_avx512_cached : pack_avx512;   // SPEF_PACK_CACHED (developer): plain stores instead of streaming ones
#endif
  return pack_scalar;
}

class Pool {
 public:
  static Pool& get() { static Pool p; return p; }
  int threads() const { return (int)th_.size() + 1; }
  // blocking: dst[i] = bf16(src[i]) for i < n, split into blocks the workers (and the caller) pull from a shared counter
  void pack(const float* src, uint16_t* dst, size_t n) {
    constexpr size_t BLK = 1u << 16;   // 256 KB of floats per pull
    const size_t nblk = (n + BLK - 1) / BLK;
    if (nblk <= 1 || th_.empty()) { fn_(src, dst, n); return; }
    {
      std::lock_guard<std::mutex> g(m_);
      src_ = src; dst_ = dst; n_ = n; nblk_ = nblk; next_.store(0); finished_ = 0; ++gen_;
    }
    cv_.notify_all();
    work();
    // every worker checks in once per generation, so none of them can still be reading this call's state after it returns
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [&] { return finished_ == (int)th_.size(); });
  }

 private:
  Pool() : fn_(pick_pack()) {
    int cores = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
#endif
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    int t = cores / ranks;
    if (t > 16) t = 16;
    if (const char* e = getenv("SPEF_PACK_THREADS")) t = atoi(e);
    if (t < 1) t = 1;
    for (int i = 1; i < t; ++i) th_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void work() {
    constexpr size_t BLK = 1u << 16;
    for (;;) {
      const size_t b = next_.fetch_add(1);
      if (b >= nblk_) break;
      const size_t o = b * BLK, len = (o + BLK <= n_) ? BLK : n_ - o;
      fn_(src_ + o, dst_ + o, len);
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
      {
        std::lock_guard<std::mutex> g(m_);
        ++finished_;
      }
      done_.notify_one();
    }
  }
  PackFn fn_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  bool stop_ = false;
  unsigned long gen_ = 0;
  int finished_ = 0;
  const float* src_ = nullptr;
  uint16_t* dst_ = nullptr;
  size_t n_ = 0, nblk_ = 0;
  std::atomic<size_t> next_{0};
};

void pack_bf16(const float* src, uint16_t* dst, size_t n) { Pool::get().pack(src, dst, n); }
int pack_threads() { return Pool::get().threads(); }

}  // namespace spef_host
