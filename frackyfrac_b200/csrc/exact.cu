// Exact pair kernel: fp64, the reference's own arithmetic sequence.
//
// One thread per sample pair walks ALL node ids in ascending order over the
// dense fp64 embedding.  The reference (frcfrc/unifrac.go:144-205) walks the
// two id-sorted sparse lists with a merge-join; visiting the absent nodes as
// well changes nothing because they are skipped by the same membership tests,
// and every product and sum below is a separate IEEE operation (__dmul_rn /
// __dadd_rn: no FMA contraction, matching Go on amd64).  The result is
// bit-identical to the reference for any input, which is how the byte-exact
// testdata/*.want files pass through the GPU.  This is a device kernel, not a
// CPU fallback; it is selected for small problems (or on request) because it
// does O(pairs * nodes) fp64 work without tensor cores.
#include "frc_internal.h"

namespace frc {
namespace {

__device__ __forceinline__ void pair_of(int64_t p, int64_t& i, int64_t& j) {
  // largest i with i(i-1)/2 <= p
  int64_t g = static_cast<int64_t>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(p))) * 0.5);
  while (g * (g - 1) / 2 > p) --g;
  while ((g + 1) * g / 2 <= p) ++g;
  i = g;
  j = p - g * (g - 1) / 2;
}

template <bool kWeighted>
__global__ void k_exact_pairs(const double* __restrict__ E, const double* __restrict__ length,
                              int32_t n_nodes, int64_t ld, int64_t first, int64_t count,
                              double* __restrict__ out) {
  int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  int64_t i, j;
  pair_of(first + t, i, j);
  const double* ea = E + i;  // a = sample i (first of the pair, common.go:25)
  const double* eb = E + j;
  if (kWeighted) {
    double numer = 0.0, denom = 0.0;
    for (int32_t k = 0; k < n_nodes; ++k) {
      const double a = ea[static_cast<int64_t>(k) * ld];
      const double b = eb[static_cast<int64_t>(k) * ld];
      const bool pa = a > 0.0, pb = b > 0.0;
      if (!(pa || pb)) continue;
      const double l = length[k];
      if (pa && pb) {
        numer = __dadd_rn(numer, __dmul_rn(l, fabs(__dsub_rn(a, b))));
        denom = __dadd_rn(denom, __dmul_rn(l, __dadd_rn(a, b)));
      } else {
        const double x = __dmul_rn(l, pa ? a : b);
        numer = __dadd_rn(numer, x);
        denom = __dadd_rn(denom, x);
      }
    }
    out[t] = __ddiv_rn(numer, denom);
  } else {
    double result = 0.0, common = 0.0;
    for (int32_t k = 0; k < n_nodes; ++k) {
      const bool pa = ea[static_cast<int64_t>(k) * ld] > 0.0;
      const bool pb = eb[static_cast<int64_t>(k) * ld] > 0.0;
      if (pa && pb) common = __dadd_rn(common, length[k]);
      else if (pa || pb) result = __dadd_rn(result, length[k]);
    }
    out[t] = __ddiv_rn(result, __dadd_rn(result, common));
  }
}

}  // namespace

int launch_exact_pairs(const double* E, const double* length, int32_t n_nodes, int64_t ld,
                       bool weighted, int64_t first, int64_t count, double* out, cudaStream_t s) {
  if (count <= 0) return 0;
  unsigned grid = static_cast<unsigned>((count + 127) / 128);
  if (weighted)
    k_exact_pairs<true><<<grid, 128, 0, s>>>(E, length, n_nodes, ld, first, count, out);
  else
    k_exact_pairs<false><<<grid, 128, 0, s>>>(E, length, n_nodes, ld, first, count, out);
  return 1;
}

}  // namespace frc
