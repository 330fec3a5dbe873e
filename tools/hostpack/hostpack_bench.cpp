// Host-side transposition rate: byte-per-position strings -> planar units (4 bit planes per 32 positions), all cores.
// Build: g++ -O3 -mavx2 -pthread -o hostpack_bench hostpack_bench.cpp ; run: ./hostpack_bench [GB] [threads]
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static inline void unit_from_32(const uint8_t* p, uint32_t out[4]) {
  const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
  out[0] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7));
  out[1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
  out[2] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
  out[3] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 4));
}
int main(int argc, char** argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 4.0;
  const int nt = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
  const int L = 150, nu = 5;
  const int64_t n = (int64_t)(gb * 1e9 / L);
  std::vector<uint8_t> blob((size_t)n * L + 64);
  std::vector<uint32_t> out((size_t)((n + 31) / 32) * nu * 32 * 4);
  auto fill = [&](int t) { for (int64_t i = n * L * t / nt; i < n * L * (t + 1) / nt; i++) blob[i] = (uint8_t)(0x1C + (i * 2654435761u >> 28)); };
  { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(fill, t); for (auto& x : th) x.join(); }
  memset(out.data(), 0, out.size() * 4);
  for (int rep = 0; rep < 3; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
      const int64_t tiles = (n + 31) / 32;
      for (int64_t tile = tiles * t / nt; tile < tiles * (t + 1) / nt; tile++) {
        uint32_t* dst = out.data() + tile * nu * 32 * 4;
        for (int l = 0; l < 32; l++) {
          const int64_t r = tile * 32 + l;
          if (r >= n) break;
          const uint8_t* row = blob.data() + r * L;
          for (int j = 0; j < nu; j++) {
            uint32_t u[4];
            unit_from_32(row + 32 * j, u);
            if (32 * (j + 1) > L) { const uint32_t keep = (1u << (L - 32 * j)) - 1u; u[0] &= keep; u[1] &= keep; u[2] &= keep; u[3] &= keep; }
            memcpy(dst + ((size_t)j * 32 + l) * 4, u, 16);
          }
        }
      }
    };
    std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work, t); for (auto& x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %.2f GB of strings -> planar in %.1f ms = %.1f GB/s of input\n", nt, n * L / 1e9, s * 1e3, n * L / 1e9 / s);
  }
  return 0;
}
