// The PCIe leg of the fast paths.  The tensor-core and FP32 pair kernels produce distances with fp32
// precision (the ratio is an fp32 division), so they STORE fp32: a band is 4 bytes per pair in HBM, crosses
// PCIe as 4 bytes per pair (doubles made the D2H copy the end-to-end bound: cfg2 100 MB = 1.76 ms of a
// 2.77 ms call) and lands in a pinned ring slot.
//
//   frc_next_f32   hands that slot out as it is: no host pass over the data.
//   frc_next       (float64 at the boundary, SURVEY §8b: Go prints the values with %v unchanged) widens it
//                  with widen_band on the job's worker pool into the ONE buffer frc_next hands out (valid
//                  until the next call; at most 2 M values = 16 MB, a larger band is delivered over several
//                  calls).  The buffer stays resident in the cores' caches from call to call, so the widened
//                  doubles cost no DRAM traffic (which the DMA engine needs for the next band); measured
//                  against non-temporal stores into per-band buffers: 1.85 vs 2.2 ms per cfg2 call.
//   exceptions     a distance the exact fix-up passes recompute in fp64 and fp32 cannot carry (underflow)
//                  travels beside the band (wire.cuh: store_fixed) and is patched into the doubles.
// Round 1 stored doubles in HBM and narrowed them in a second kernel (12 B/pair of extra HBM traffic and
// a launch per band); that pass is gone.  The exact path (bit-exact fp64) never takes this route.
#include <immintrin.h>

#include "frc_internal.h"

namespace frc {
namespace {

// kStream: non-temporal stores for a destination larger than the caches (no read-for-ownership, no
// write-back of lines nobody re-reads soon); plain stores when the destination is the one band-sized
// buffer that stays cache-resident from band to band.
template <bool kStream>
__attribute__((target("avx2"))) void widen_avx2(const float* s, double* d, int64_t n) {
  int64_t i = 0;
  for (; i < n && (reinterpret_cast<uintptr_t>(d + i) & 31u); ++i) d[i] = static_cast<double>(s[i]);
  for (; i + 8 <= n; i += 8) {
    const __m256 v = _mm256_loadu_ps(s + i);
    const __m256d lo = _mm256_cvtps_pd(_mm256_castps256_ps128(v)), hi = _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1));
    if (kStream) { _mm256_stream_pd(d + i, lo); _mm256_stream_pd(d + i + 4, hi); }
    else { _mm256_store_pd(d + i, lo); _mm256_store_pd(d + i + 4, hi); }
  }
  for (; i < n; ++i) d[i] = static_cast<double>(s[i]);
  if (kStream) _mm_sfence();
}

void widen_plain(const float* s, double* d, int64_t n) {
  for (int64_t i = 0; i < n; ++i) d[i] = static_cast<double>(s[i]);
}

}  // namespace

void widen_band(const float* src, double* dst, int64_t n, bool stream_stores) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (!avx2) widen_plain(src, dst, n);
  else if (stream_stores) widen_avx2<true>(src, dst, n);
  else widen_avx2<false>(src, dst, n);
}

}  // namespace frc

// Host-only entry for the CPU test-suite (no device needed): the widening routine frc_next runs.
extern "C" void frc_debug_widen(const float* src, double* dst, int64_t n, int stream_stores) {
  frc::widen_band(src, dst, n, stream_stores != 0);
}
