// Weighted UniFrac pair tiles on the FP32 CUDA cores (replaces
// unifracDistWeighted, frcfrc/unifrac.go:174-205, for 128 x 128 sample pairs
// per CTA).
//
//   d(i,j) = sum_k len_k |a_ik - a_jk|  /  sum_k len_k (a_ik + a_jk)
//
// The denominator is separable (W_i + W_j, computed once per sample in fp64 by
// the embedding stage), so the pair kernel only accumulates the numerator: an
// L1 distance between two sample columns of the operand, stored as tile panels Ap[np/128][kp][128].  With
// non-negative branch lengths the operand is pre-scaled (A = len * a), so the
// inner loop is exactly two FP32 instructions per (pair, node): FADD d = a - b,
// FADD acc = acc + |d| (|.| is a free source modifier).  Negative lengths use
// the FFMA form with len staged next to the tile.
//
// Tiling: CTA = 128 x 128 pairs, 256 threads, each an 8 x 8 register block
// (two 4-wide groups per side, so every shared-memory read is a conflict-free
// LDS.128).  The contraction is staged through shared memory in 32-node slabs
// with a 6-deep ring of TMA bulk copies (one 16 KB copy per operand slab).  Accumulation is two-level (flush the 64
// running sums into 64 totals every 256 nodes) to keep fp32 summation error
// well under the 1e-5 budget for contractions of 10^5..10^6 nodes.
#include "frc_internal.h"
#include "ptx.cuh"
#include "wire.cuh"

namespace frc {
namespace {

constexpr int KT = 32;        // nodes per smem slab
constexpr int STAGES = 6;     // ring depth: 5 slabs (160 KB) in flight per CTA, enough to cover NVLink latency in capacity mode
constexpr int FLUSH = 8;      // slabs between flushes (256 nodes)
constexpr int SLAB_FLOATS = KT * kTile;             // one operand side: 16 KB
constexpr int STAGE_FLOATS = 2 * SLAB_FLOATS + KT;  // + per-node lengths (128 B)
constexpr int SMEM_BYTES = STAGES * STAGE_FLOATS * 4 + STAGES * 8 + 128;

// Panel of sample tile t (PanelMap in frc_internal.h): the one local array, or the place its shard currently
// occupies in this device's HBM.
__device__ __forceinline__ const float* panel_of(const PanelMap& m, int32_t t, int32_t kp) {
  if (m.n_shards == 0) return m.shard[0] + static_cast<int64_t>(t) * kp * kTile;
  const int32_t s = t / m.tiles_per_shard, r = t - s * m.tiles_per_shard;
  return m.shard[s] + static_cast<int64_t>(r) * kp * kTile;
}

// Operand staging: one thread issues TMA bulk copies (cp.async.bulk, 16 KB per operand slab: the tile-panel
// layout makes a slab of 32 nodes x 128 samples contiguous) that complete on an mbarrier per stage.  The
// FP32-issue-bound inner loop spends no slot on copies and the copy engine keeps the whole ring in flight.
template <bool kPrescaled>
__global__ void __launch_bounds__(256, 1)
k_weighted_tiles(const PanelMap A, int64_t ld, int32_t kp, const float* __restrict__ lenf,
                 const double* __restrict__ W, const Tile* __restrict__ tiles, int64_t n_samples,
                 int64_t first, float* __restrict__ out, double flag_below, uint32_t* __restrict__ flagged,
                 unsigned long long* __restrict__ n_flagged) {
  extern __shared__ __align__(128) float smem[];
  const Tile tile = tiles[blockIdx.x];
  const int64_t i0 = static_cast<int64_t>(tile.ti) * kTile;
  const int64_t j0 = static_cast<int64_t>(tile.tj) * kTile;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n_slabs = kp / KT;
  const uint32_t bar0 = ptx::smem_u32(smem + STAGES * STAGE_FLOATS);
  (void)ld;

  // tile-panel layout [kp][128] per sample tile
  const float* const pa = panel_of(A, tile.ti, kp);  // row samples: this device's own shard
  const float* const pb = panel_of(A, tile.tj, kp);  // column samples: own shard or the visiting one
  auto issue = [&](int slab) {  // thread 0 only
    const int stage = slab % STAGES;
    float* sa = smem + stage * STAGE_FLOATS;
    const uint32_t bar = bar0 + 8u * stage;
    constexpr uint32_t kSlabBytes = SLAB_FLOATS * 4;
    ptx::mbar_expect_tx(bar, 2 * kSlabBytes + (kPrescaled ? 0u : KT * 4u));
    ptx::bulk_load_1d(ptx::smem_u32(sa), pa + static_cast<int64_t>(slab) * SLAB_FLOATS, kSlabBytes, bar);
    ptx::bulk_load_1d(ptx::smem_u32(sa + SLAB_FLOATS), pb + static_cast<int64_t>(slab) * SLAB_FLOATS, kSlabBytes, bar);
    if (!kPrescaled) ptx::bulk_load_1d(ptx::smem_u32(sa + 2 * SLAB_FLOATS), lenf + slab * KT, KT * 4u, bar);
  };
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(bar0 + 8u * s, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int p = 0; p < STAGES - 1 && p < n_slabs; ++p) issue(p);

  float acc[8][8], tot[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) { acc[a][b] = 0.f; tot[a][b] = 0.f; }

  for (int slab = 0; slab < n_slabs; ++slab) {
    // the stage that slab + STAGES - 1 goes to was read in the previous iteration (barrier at its end)
    if (tid == 0 && slab + STAGES - 1 < n_slabs) issue(slab + STAGES - 1);
    const int stage = slab % STAGES;
    ptx::mbar_wait(bar0 + 8u * stage, static_cast<uint32_t>((slab / STAGES) & 1));
    const float* sa = smem + stage * STAGE_FLOATS;
    const float* sb = sa + SLAB_FLOATS;
    const float* sl = sb + SLAB_FLOATS;
#pragma unroll 4
    for (int kk = 0; kk < KT; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(sa + kk * kTile + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(sa + kk * kTile + 64 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sb + kk * kTile + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(sb + kk * kTile + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      if (kPrescaled) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) acc[a][b] += fabsf(av[a] - bv[b]);
      } else {
        const float l = sl[kk];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(l, fabsf(av[a] - bv[b]), acc[a][b]);
      }
    }
    if ((slab % FLUSH) == FLUSH - 1) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) { tot[a][b] += acc[a][b]; acc[a][b] = 0.f; }
    }
    __syncthreads();  // every thread is done with this stage: it may be refilled
  }

#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t i = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
    if (i >= n_samples) continue;
    const double wi = W[i];
    float* orow = out + (i * (i - 1) / 2 - first);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int64_t j = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
      if (j < i) {
        const double d = static_cast<double>(tot[a][b] + acc[a][b]) / (wi + W[j]);
        orow[j] = static_cast<float>(d);  // fp32 numerator: fp32 is what the value carries (wire.cu)
        // fp32 operands carry 6e-8 relative error each, so the L1 numerator is off by up to
        // 6e-8 * (W_i + W_j) = 6e-8 / d relative: small distances are recomputed (k_weighted_fixup)
        if (d < flag_below) {
          unsigned long long slot = atomicAdd(n_flagged, 1ULL);
          flagged[slot] = static_cast<uint32_t>(orow + j - out);
        }
      }
    }
  }
}

// Recompute of the flagged pairs from the CSR rows themselves, one CTA per pair at a time:
//   diff[v] = a_i(v) - a_j(v) for every node on a leaf-to-root path of either sample, accumulated as
//   62-bit fixed point with integer atomics (order-independent: deterministic, and identical samples
//   cancel to exactly 0), then numerator = sum_v len_v * |diff[v]| in fp64 over the touched nodes (the
//   second walk takes each node's value with an atomic exchange, which also restores the zeros).
// ws is one zeroed int64 array of n_nodes entries per CTA.
__global__ void __launch_bounds__(256)
k_weighted_fixup(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                 const double* __restrict__ val, const int32_t* __restrict__ parent,
                 const double* __restrict__ length, int32_t n_nodes, const double* __restrict__ total,
                 const double* __restrict__ W, const uint32_t* __restrict__ flagged,
                 const unsigned long long* __restrict__ n_flagged, unsigned long long* __restrict__ count_host,
                 int64_t first, long long* __restrict__ ws, float* __restrict__ out, const Exceptions ex) {
  __shared__ double red[8];
  const unsigned long long n = *n_flagged;
  if (blockIdx.x == 0 && threadIdx.x == 0) *count_host = n;  // mapped pinned memory
  long long* diff = ws + static_cast<int64_t>(blockIdx.x) * n_nodes;
  constexpr double kFix = 1152921504606846976.0;  // 2^60: proportions are <= 1, a node sums <= 1
  for (unsigned long long w = blockIdx.x; w < n; w += gridDim.x) {
    const uint32_t off = flagged[w];
    const int64_t p = first + off;
    int64_t i = static_cast<int64_t>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(p))) * 0.5);
    while (i * (i - 1) / 2 > p) --i;
    while ((i + 1) * i / 2 <= p) ++i;
    const int64_t j = p - i * (i - 1) / 2;
    const int64_t smp[2] = {i, j};
    // without normalisation (-l) raw values are summed: scale them into the fixed-point range
    double inv[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const double t = total ? total[smp[side]] : 0.0;
      inv[side] = total ? (t > 0.0 ? 1.0 / t : 0.0) : 0.0;
    }
    double unscale = 1.0 / kFix;
    if (!total) {  // common power-of-two scale from the two raw sums (exact)
      double big = 0.0;
      for (int side = 0; side < 2; ++side)
        for (int64_t k = row_ptr[smp[side]]; k < row_ptr[smp[side] + 1]; ++k) big += val[k];
      int e;
      frexp(big > 0.0 ? big : 1.0, &e);
      inv[0] = inv[1] = ldexp(1.0, -e);
      unscale = ldexp(1.0, e) / kFix;
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int64_t b = row_ptr[smp[side]], e = row_ptr[smp[side] + 1];
      for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) {
        const long long x = __double2ll_rn(val[k] * inv[side] * kFix);
        const long long add = side == 0 ? x : -x;
        for (int32_t v = col[k]; v >= 0; v = parent[v])
          atomicAdd(reinterpret_cast<unsigned long long*>(diff + v), static_cast<unsigned long long>(add));
      }
    }
    __syncthreads();
    double num = 0.0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int64_t b = row_ptr[smp[side]], e = row_ptr[smp[side] + 1];
      for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x)
        for (int32_t v = col[k]; v >= 0; v = parent[v]) {
          const long long x = static_cast<long long>(atomicExch(reinterpret_cast<unsigned long long*>(diff + v), 0ULL));
          if (x != 0) num += length[v] * (fabs(static_cast<double>(x)) * unscale);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) num += __shfl_xor_sync(0xffffffffu, num, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = num;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int k = 0; k < 8; ++k) tot += red[k];
      store_fixed(out, off, tot / (W[i] + W[j]), first, ex);
    }
    __syncthreads();
  }
}

}  // namespace

int launch_weighted_fixup(const DevCsr& a, const DevTree& t, const double* total, const double* W,
                          const uint32_t* flagged, const unsigned long long* n_flagged,
                          unsigned long long* count_host, int64_t first, long long* ws, int ws_ctas, float* out,
                          const Exceptions& ex, cudaStream_t s) {
  k_weighted_fixup<<<ws_ctas, 256, 0, s>>>(a.row_ptr, a.col, a.val, t.parent, t.length, t.n_nodes, total, W, flagged,
                                           n_flagged, count_host, first, ws, out, ex);
  return 1;
}

void weighted_setup() {
  cudaFuncSetAttribute(k_weighted_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(k_weighted_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

int launch_weighted_tiles(const PanelMap& A, int64_t ld, int32_t kp, const float* lenf, bool prescaled,
                          const double* W, const Tile* tiles, int32_t n_tiles, int64_t n_samples,
                          int64_t first, float* out, double flag_below, uint32_t* flagged,
                          unsigned long long* n_flagged, int num_sms, cudaStream_t s) {
  (void)num_sms;
  if (n_tiles <= 0) return 0;
  if (prescaled)
    k_weighted_tiles<true><<<n_tiles, 256, SMEM_BYTES, s>>>(A, ld, kp, lenf, W, tiles, n_samples, first, out,
                                                            flag_below, flagged, n_flagged);
  else
    k_weighted_tiles<false><<<n_tiles, 256, SMEM_BYTES, s>>>(A, ld, kp, lenf, W, tiles, n_samples, first, out,
                                                             flag_below, flagged, n_flagged);
  return 1;
}

}  // namespace frc
