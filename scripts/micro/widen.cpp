// Host-side fp32 -> fp64 widening throughput (would PCIe carry fp32 distances and the host widen them?).
// g++ -O2 -mavx2 -pthread widen.cpp -o widen && ./widen [threads]
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void widen(const float* s, double* d, size_t n, bool stream) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    __m256 v = _mm256_loadu_ps(s + i);
    __m256d lo = _mm256_cvtps_pd(_mm256_castps256_ps128(v));
    __m256d hi = _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1));
    if (stream) { _mm256_stream_pd(d + i, lo); _mm256_stream_pd(d + i + 4, hi); }
    else { _mm256_storeu_pd(d + i, lo); _mm256_storeu_pd(d + i + 4, hi); }
  }
  for (; i < n; ++i) d[i] = s[i];
  if (stream) _mm_sfence();
}

int main(int argc, char** argv) {
  const size_t n = 12'497'500;
  float* s = static_cast<float*>(aligned_alloc(64, n * 4 + 64));
  double* d = static_cast<double*>(aligned_alloc(64, n * 8 + 64));
  for (size_t i = 0; i < n; ++i) s[i] = static_cast<float>(i % 1000) * 1e-3f;
  memset(d, 0, n * 8);
  for (int threads : {1, 2, 4, 8, 12, 16}) {
    if (argc > 1 && threads > atoi(argv[1])) break;
    for (int stream = 0; stream < 2; ++stream) {
      double best = 1e9;
      for (int rep = 0; rep < 7; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) {
          size_t b = n * t / threads / 8 * 8, e = t + 1 == threads ? n : n * (t + 1) / threads / 8 * 8;
          th.emplace_back(widen, s + b, d + b, e - b, stream != 0);
        }
        for (auto& x : th) x.join();
        best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
      printf("threads %2d stream %d: %.3f ms  (%.1f GB/s of doubles)\n", threads, stream, best, n * 8 / best * 1e-6);
    }
  }
  return 0;
}
