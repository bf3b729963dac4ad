// The PCIe leg of the fast paths.  The tensor-core and FP32 pair kernels produce distances with
// fp32 precision (the ratio is an fp32 division widened to the ABI's double, unweighted_tc.cu
// widen_f32); moving them over PCIe as doubles made the 8 B/pair D2H copy the end-to-end bound
// (cfg2: 100 MB = 1.76 ms of a 2.77 ms call).  So a band crosses the bus as fp32 and the host
// widens it into the pinned double buffer frc_next hands out (SURVEY §8b keeps float64 at the
// boundary: Go prints the values with %v unchanged).
//
//   k_narrow_band   device: double band -> float band (one extra pass over HBM, 12 B/pair, only
//                   when the band goes to the host).  Pairs rewritten by the exact fix-up pass
//                   carry full doubles; they round to fp32 (6e-8 relative, budget 1e-5).  A value
//                   fp32 cannot hold to 2^-23 relative (underflow) is counted in mapped host
//                   memory and the host then fetches that band as doubles instead.
//   widen_band      host: float -> double on the job's worker pool, into the ONE buffer frc_next hands
//                   out (valid until the next call, as the ABI says; at most 2 M values = 16 MB, a
//                   larger band is delivered over several calls).  The buffer stays resident in the
//                   cores' caches from call to call, so the widened doubles cost no DRAM traffic
//                   (which the DMA engine needs for the next band) and the consumer finds them in
//                   cache; measured against non-temporal stores into per-band buffers: 1.85 vs 2.2 ms
//                   per cfg2 call.
// The exact path (bit-exact fp64) never takes this route.
#include <immintrin.h>

#include "frc_internal.h"

namespace frc {
namespace {

__global__ void __launch_bounds__(256) k_narrow_band(const double* __restrict__ in, float* __restrict__ out,
                                                     int64_t n, unsigned long long* __restrict__ n_bad) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
  bool bad = false;
  for (int64_t k = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; k < n; k += stride) {
    if (k + 4 <= n) {
      const double2 a = *reinterpret_cast<const double2*>(in + k);
      const double2 b = *reinterpret_cast<const double2*>(in + k + 2);
      const float4 f = make_float4(static_cast<float>(a.x), static_cast<float>(a.y), static_cast<float>(b.x),
                                   static_cast<float>(b.y));
      bad |= fabs(static_cast<double>(f.x) - a.x) > fabs(a.x) * 0x1p-23 || fabs(static_cast<double>(f.y) - a.y) > fabs(a.y) * 0x1p-23 ||
             fabs(static_cast<double>(f.z) - b.x) > fabs(b.x) * 0x1p-23 || fabs(static_cast<double>(f.w) - b.y) > fabs(b.y) * 0x1p-23;
      *reinterpret_cast<float4*>(out + k) = f;
    } else {
      for (int64_t x = k; x < n; ++x) {
        const double d = in[x];
        const float f = static_cast<float>(d);
        bad |= fabs(static_cast<double>(f) - d) > fabs(d) * 0x1p-23;
        out[x] = f;
      }
    }
  }
  if (bad) atomicAdd_system(n_bad, 1ULL);  // rare: straight into mapped pinned memory
}

// Doubles on the bus (ranks with few host threads): the same rounding, in place, so that the host sees the same
// values whichever format crossed PCIe — the output bytes must not depend on how many ranks share a host
// (identical for 1, 2, 4 and 8 GPUs).  A value fp32 cannot carry stays as it is, as on the fp32 route.
__global__ void __launch_bounds__(256) k_round_band(double* __restrict__ io, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const double d = io[k];
    const float f = static_cast<float>(d);
    const bool bad = fabs(static_cast<double>(f) - d) > fabs(d) * 0x1p-23;
    if (!bad) io[k] = static_cast<double>(f);
  }
}

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

int launch_narrow_band(const double* in, float* out, int64_t n, unsigned long long* n_bad_mapped, int num_sms,
                       cudaStream_t s) {
  if (n <= 0) return 0;
  const int64_t want = (n + 256 * 4 - 1) / (256 * 4);
  const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(num_sms) * 8));
  k_narrow_band<<<grid, 256, 0, s>>>(in, out, n, n_bad_mapped);
  return 1;
}

int launch_round_band(double* io, int64_t n, int num_sms, cudaStream_t s) {
  if (n <= 0) return 0;
  const int64_t want = (n + 255) / 256;
  const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(num_sms) * 8));
  k_round_band<<<grid, 256, 0, s>>>(io, n);
  return 1;
}

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
