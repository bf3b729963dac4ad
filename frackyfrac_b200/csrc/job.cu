// Host side of libfrcfrc_cuda: the C ABI of include/frcfrc_cuda.h.
//
// A job is one unifrac() call (frcfrc/unifrac.go:97-124): validate + copy the
// flattened tree and CSR abundances, upload them in one pinned transfer, build
// the branch embedding on the device, then compute the lower triangle in bands
// of consecutive rows.  Bands alternate between two streams, each followed by
// its own pinned D2H copy, and are handed to the caller strictly in flat-index
// order by frc_next (the ordered iter.Seq of unifrac.go:209-228).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/frcfrc_cuda.h"
#include "frc_internal.h"
#include "host_pool.h"

using namespace frc;
using frc_host::Pool;

namespace {

thread_local std::string g_create_error;

// ------------------------------------------------------------------ arenas
struct Arena {
  struct Block { char* p; size_t cap, off; };
  std::vector<Block> blocks;
  bool pinned = false;
  size_t min_block = 0;

  void* alloc(size_t bytes, cudaError_t* err) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    for (auto& b : blocks)
      if (b.cap - b.off >= bytes) { void* r = b.p + b.off; b.off += bytes; return r; }
    size_t cap = std::max(bytes, min_block);
    char* p = nullptr;
    cudaError_t e = pinned ? cudaHostAlloc(reinterpret_cast<void**>(&p), cap, cudaHostAllocDefault)
                           : cudaMalloc(reinterpret_cast<void**>(&p), cap);
    if (e != cudaSuccess && cap > bytes) {  // retry without the rounding-up
      cudaGetLastError();
      cap = bytes;
      e = pinned ? cudaHostAlloc(reinterpret_cast<void**>(&p), cap, cudaHostAllocDefault)
                 : cudaMalloc(reinterpret_cast<void**>(&p), cap);
    }
    if (e != cudaSuccess) { cudaGetLastError(); *err = e; return nullptr; }
    blocks.push_back({p, cap, bytes});
    return p;
  }
  void reset() { for (auto& b : blocks) b.off = 0; }
  void release() {
    for (auto& b : blocks) { if (pinned) cudaFreeHost(b.p); else cudaFree(b.p); }
    blocks.clear();
  }
};

uint16_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>((u >> 16) | ((u & 0xFFFFu) ? 0x40u : 0u));
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
float bf16_to_float(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

constexpr int64_t kWidePiece = 2LL << 20;        // fp32 wire: pairs widened per frc_next call (16 MB of doubles: cache-resident)
constexpr int kSlots = 8;                       // output ring: enough for the compute to run ahead of the D2H
constexpr int kMinSlots = 3;
constexpr int64_t kSlotBytesBudget = 1LL << 30;  // bytes the output ring may take beyond kMinSlots (pinned allocation is slow: ~0.4 s per GB)
constexpr double kFlagBelow = 0.125;            // fast unweighted: recompute d below this exactly
constexpr double kFlagBelowW = 0.03125;         // fast weighted: fp32 error is ~6e-8 / d -> 2e-6 at this d
constexpr int64_t kExactWorkLimit = 1LL << 28;  // AUTO: pairs * nodes at or below this -> exact

}  // namespace

struct frc_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream[3] = {nullptr, nullptr, nullptr};  // two compute streams + one copy (D2H) stream
  Arena dev, pin;
  std::unique_ptr<Pool> pool;
  std::vector<std::vector<int64_t>> stamps;  // per-worker duplicate-detection scratch
  int64_t stamp_epoch = 0;
  bool in_use = false;
  Comm* comm = nullptr;  // NCCL communicator over the ranks of a multi-GPU run (frc_ctx_comm_init)
  // "symmetric heap": buffers every rank allocates identically, so block index + offset name the same
  // buffer on every rank; each block is mapped from the peers with CUDA IPC (peer stores over NVLink)
  Arena sym;
  std::vector<std::vector<char*>> sym_peer;  // [block][rank] mapped base (own rank: the local base)
};

struct Band { int64_t row0, row1, first, count; int32_t tile_off, n_tiles; };

struct Slot {
  double* dev = nullptr;
  double* host = nullptr;
  float* dev32 = nullptr;   // fp32 wire (wire.cu): narrowed copy of the band and its pinned landing buffer
  float* host32 = nullptr;
  unsigned long long* n_bad_host = nullptr;  // values of this band fp32 cannot carry (mapped pinned)
  uint32_t* flagged = nullptr;
  unsigned long long* n_flagged_host = nullptr;
  cudaEvent_t k0 = nullptr, k1 = nullptr, k2 = nullptr, k3 = nullptr, done = nullptr;
  int band = -1;  // index into mine[]
};

struct frc_job {
  frc_ctx* ctx = nullptr;
  bool own_ctx = false;
  frc_opts_t opts{};
  std::string err;
  frc_info_t info{};

  int64_t N = 0, np = 0, nnz = 0;
  int32_t B = 0, kp = 0, nw = 0;
  bool exact = false, weighted = false, prescale = true;

  std::vector<int32_t> level_ptr;  // host
  std::vector<Band> bands;         // whole triangle
  std::vector<int> mine;           // indices into bands
  std::vector<Tile> tiles;         // host copy (uploaded)

  // device
  DevTree dtree;
  DevCsr dcsr;
  char* d_inputs = nullptr;
  size_t input_bytes = 0;
  Tile* d_tiles = nullptr;
  int32_t* d_level_ptr = nullptr;
  unsigned long long* d_flag_counts = nullptr;  // one per band of this rank
  bool fused_embed = true;
  bool zero_copy = false;
  bool wire32 = false;  // bands cross PCIe as fp32 and are widened on the host (wire.cu)
  double* wide = nullptr;  // wire32: the one double buffer frc_next hands out (<= kWidePiece values)
  int64_t piece_first = 0, piece_off = 0, piece_end = 0;  // wire32: progress through the held band
  int64_t ws_slab = 0;        // fast weighted: samples per fp64 embedding slab
  bool peer_push = false;     // sharded unweighted: the bits kernel stores into every rank's bitsT (no NCCL for the bits)
  uint32_t* d_bits2[2] = {nullptr, nullptr};  // double-buffered by run parity (a rank may be one step ahead)
  PeerPtrs peers[2];
  int64_t run_count = 0;
  bool sharded = false;       // embedding built for this rank's sample shard, then all-gathered
  int32_t shard_w0 = 0, shard_nw = 0;  // word columns (32 samples) of the shard
  int tc_ctas = 1;  // CTAs per tensor-core tile group (2 = cta_group::2 pairs)
  bool i8 = false;  // fast unweighted: u8 block-floating-point operands (kind::i8) instead of bf16 hi/lo
  // per operand column (fast unweighted): q0/q1/q2 = bf16 (unused, len_hi, len_lo) or u8 (a, m_hi, m_lo)
  void *d_q0 = nullptr, *d_q1 = nullptr, *d_q2 = nullptr;
  int32_t *d_order = nullptr, *d_col_exp = nullptr;
  uint32_t* d_qam = nullptr;  // u8: a * m per operand column
  bool intacc = false;        // u8: chunk scales within 2^16 -> integer accumulation / integer row sums
  bool bits_feed = false;     // operand tiles expanded inside the pair kernel from bit rows (no operands in HBM)
  uint32_t* d_bitsS = nullptr;
  BitsOperands* bo = nullptr;
  int32_t e_min = 0;          // smallest chunk exponent (integer unit = 2^e_min)
  long long* d_r_int = nullptr;
  long long* d_fix_ws = nullptr;  // fast weighted fix-up: one zeroed int64[B] per SM
  uint8_t* d_need = nullptr;  // per block of 256 samples: which operands this rank's tiles read (world > 1)
  double *d_lenq = nullptr, *d_len_col = nullptr, *d_flag_u = nullptr;
  TcChunks d_chunks;
  float* d_lenf = nullptr;
  double *d_E = nullptr, *d_total = nullptr, *d_W = nullptr, *d_r = nullptr, *d_scratch = nullptr;
  float* d_A = nullptr;
  uint32_t *d_bits = nullptr, *d_node_scratch = nullptr;
  void *d_P = nullptr, *d_Bh = nullptr, *d_Bl = nullptr;
  TcOperands* tc = nullptr;

  Slot slots[kSlots];
  int n_slots = 0;
  size_t next_enqueue = 0, next_deliver = 0;
  int held_slot = -1;  // slot whose buffer the caller currently reads
  cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr, ev_embed0 = nullptr, ev_embed1 = nullptr;
  cudaEvent_t ev_run1 = nullptr, ev_join = nullptr, ev_bits = nullptr, ev_rsum = nullptr;
  bool run_timed = false;
  bool embed_timed = false;
  std::chrono::steady_clock::time_point t_host0;
};

namespace {

#define JOB_CUDA(job, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      (job)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                       \
      cudaGetLastError();                                                                    \
      return _e == cudaErrorMemoryAllocation ? FRC_ERR_OOM : FRC_ERR_CUDA;                   \
    }                                                                                        \
  } while (0)

int fail(frc_job* j, int code, const std::string& msg) { j->err = msg; return code; }

// Which rank computes which band: bands are as equal as whole tile rows allow, and the rest of
// the imbalance is removed by dealing them largest-first to the least loaded rank (deterministic,
// the same on every rank and in frc_plan_bands).
std::vector<int> band_owners(const std::vector<int64_t>& rows, int world) {
  const size_t n = rows.size() > 0 ? rows.size() - 1 : 0;
  std::vector<int> owner(n, 0);
  if (world <= 1) return owner;
  std::vector<size_t> order(n);
  std::vector<int64_t> cnt(n);
  for (size_t k = 0; k < n; ++k) {
    order[k] = k;
    cnt[k] = tri(rows[k + 1]) - (rows[k] >= 2 ? tri(rows[k]) : 0);
  }
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cnt[a] > cnt[b]; });
  std::vector<int64_t> load(world, 0);
  for (size_t k : order) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (load[r] < load[best]) best = r;
    owner[k] = best;
    load[best] += cnt[k];
  }
  return owner;
}

// Row boundaries of the bands of the triangle (multiples of the tile size, first 0, last N).
//   requested > 0 : uniform bands of that many rows (rounded up to whole tiles);
//   otherwise     : bands of (nearly) EQUAL PAIR COUNT, boundaries at N * sqrt(k / n) -- equal work per
//                   launch, so the ranks are balanced and no band is a sliver that cannot fill the GPU.
//                   Distances stay in HBM (no D2H to overlap): one band, or two per rank;
//                   with D2H: about 8 per rank so that copies overlap the kernels.  More bands when one
//                   would exceed 96 MB (streamed) / 1.5 GB (kept in HBM) of distances.
std::vector<int64_t> band_boundaries(int64_t N, int64_t requested, int world, bool d2h) {
  std::vector<int64_t> rows;
  rows.push_back(0);
  if (N <= 0) return rows;
  if (requested > 0) {
    const int64_t step = round_up(requested, kTile);
    for (int64_t r = step; r < N; r += step) rows.push_back(r);
    rows.push_back(N);
    return rows;
  }
  const int64_t tile_rows = (N + kTile - 1) / kTile;
  int64_t n = d2h ? 8LL * world : (world == 1 ? 1 : 2LL * world);
  if (const char* e = getenv("FRC_BANDS")) { if (d2h && atoi(e) > 0) n = static_cast<int64_t>(atoi(e)) * world; }
  const double total_bytes = 8.0 * static_cast<double>(N) * static_cast<double>(N - 1) / 2.0;
  // streamed to the host: bands of <= 96 MB keep the pinned ring small (pinning memory costs ~0.4 s
  // per GB, paid by the first job of a context) and are still several waves of tiles each
  const double band_cap = d2h ? 96.0 * 1024 * 1024 : 1536.0 * 1024 * 1024;
  const int64_t by_size = static_cast<int64_t>(std::ceil(total_bytes / band_cap));
  if (by_size > n) n = round_up(by_size, 2LL * world);
  n = std::max<int64_t>(1, std::min(n, std::max<int64_t>(world, tile_rows / 2)));
  for (int64_t k = 1; k < n; ++k) {
    int64_t r = static_cast<int64_t>(std::llround(static_cast<double>(N) * std::sqrt(static_cast<double>(k) / n) / kTile)) * kTile;
    r = std::max(r, rows.back() + kTile);
    if (r >= N) break;
    rows.push_back(r);
  }
  rows.push_back(N);
  return rows;
}

template <class T>
T* dev_alloc(frc_job* j, size_t n, int* rc) {
  cudaError_t e = cudaSuccess;
  void* p = j->ctx->dev.alloc(n * sizeof(T), &e);
  if (!p) {
    j->err = "device allocation of " + std::to_string(n * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e);
    *rc = FRC_ERR_OOM;
  }
  return static_cast<T*>(p);
}
template <class T>
T* pin_alloc(frc_job* j, size_t n, int* rc) {
  cudaError_t e = cudaSuccess;
  void* p = j->ctx->pin.alloc(n * sizeof(T), &e);
  if (!p) {
    j->err = "pinned host allocation of " + std::to_string(n * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e);
    *rc = FRC_ERR_OOM;
  }
  return static_cast<T*>(p);
}

// Maps the blocks of the symmetric heap that the peers have not seen yet (collective: every rank
// allocates the same sequence, so every rank gains the same blocks at the same time).
int sym_sync(frc_job* j) {
  frc_ctx* c = j->ctx;
  const int world = comm_world(c->comm), rank = comm_rank(c->comm);
  while (c->sym_peer.size() < c->sym.blocks.size()) {
    const size_t b = c->sym_peer.size();
    cudaIpcMemHandle_t mine;
    JOB_CUDA(j, cudaIpcGetMemHandle(&mine, c->sym.blocks[b].p));
    std::vector<cudaIpcMemHandle_t> all(world);
    std::string cerr;
    if (!comm_all_gather_host(c->comm, &mine, sizeof(mine), all.data(), c->stream[0], &cerr))
      return fail(j, FRC_ERR_CUDA, cerr);
    std::vector<char*> bases(world, nullptr);
    for (int r = 0; r < world; ++r) {
      if (r == rank) { bases[r] = c->sym.blocks[b].p; continue; }
      void* p = nullptr;
      JOB_CUDA(j, cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
      bases[r] = static_cast<char*>(p);
    }
    c->sym_peer.push_back(bases);
  }
  return FRC_OK;
}

// Queue the embedding stage on stream 0.
int run_embedding(frc_job* j) {
  frc_ctx* c = j->ctx;
  cudaStream_t s = c->stream[0];
  int launches = 0;
  j->t_host0 = std::chrono::steady_clock::now();
  if (j->d_flag_counts)
    JOB_CUDA(j, cudaMemsetAsync(j->d_flag_counts, 0, sizeof(unsigned long long) * (j->mine.size() + 1), s));
  JOB_CUDA(j, cudaEventRecord(j->ev_embed0, s));
  if (j->exact) {
    launches += launch_embed_f64(j->dtree, j->level_ptr.data(), j->dcsr, j->d_E, j->np, 0, s);
    if (j->opts.normalize != 0) {
      launches += launch_totals_f64(j->d_E, j->B, j->np, j->N, j->d_total, s);
      launches += launch_normalize_f64(j->d_E, j->B, j->np, j->N, j->d_total, s);
    }
    j->info.embed_bytes = 2LL * j->B * j->np * 8 + 12LL * j->nnz;
  } else if (j->weighted) {
    // fast weighted: the fp64 embedding is built one slab of samples at a time (E[B][slab] stays
    // <= 2 GB whatever the sample count) and leaves as the fp32 tile-panel operand + denominators;
    // a rank of a sharded run only builds the slabs of its own sample shard
    const bool norm = j->opts.normalize != 0;
    const int64_t per_rank = j->np / (j->sharded ? j->opts.world : 1);
    const int64_t s_begin = j->sharded ? j->opts.rank * per_rank : 0, s_stop = s_begin + per_rank;
    for (int64_t sb = s_begin; sb < s_stop; sb += j->ws_slab) {
      const int64_t ld = std::min<int64_t>(j->ws_slab, s_stop - sb);
      launches += launch_embed_f64(j->dtree, j->level_ptr.data(), j->dcsr, j->d_E, ld, sb, s);
      if (norm) launches += launch_totals_fast_f64(j->d_E, j->B, ld, j->d_total + sb, j->d_scratch, s);
      launches += launch_weighted_operand_panels(j->d_E, j->dtree.length, j->B, j->kp, ld, sb,
                                                 norm ? j->d_total + sb : nullptr, j->prescale, j->d_A, j->d_W,
                                                 j->d_scratch, s);
    }
    if (j->sharded) {
      void* bufs[3] = {j->d_A, j->d_W, j->d_total};
      const size_t bytes[3] = {static_cast<size_t>(per_rank) * j->kp * sizeof(float),
                               static_cast<size_t>(per_rank) * sizeof(double),
                               static_cast<size_t>(per_rank) * sizeof(double)};
      std::string cerr;
      if (!comm_all_gather_inplace(c->comm, bufs, bytes, norm ? 3 : 2, s, &cerr)) return fail(j, FRC_ERR_CUDA, cerr);
      j->info.gather_bytes = static_cast<int64_t>(bytes[0] + bytes[1] + (norm ? bytes[2] : 0)) * (j->opts.world - 1);
    }
    // write each element once, read it once when folded into its parent (+ CSR)
    j->info.embed_bytes = 2LL * j->B * j->np * 8 + 12LL * j->nnz;
  } else {
    if (j->fused_embed) {
      const int cur = static_cast<int>(j->run_count & 1);
      if (j->peer_push) {
        j->d_bits = j->d_bits2[cur];
        if (j->run_count == 0) {  // nobody may still be reading these buffers for an earlier job
          std::string cerr;
          if (!comm_barrier(c->comm, s, &cerr)) return fail(j, FRC_ERR_CUDA, cerr);
        }
      }
      ++j->run_count;
      static const PeerPtrs no_peers;
      launches += launch_embed_presence_fused(j->dtree, j->d_level_ptr, j->dcsr, j->nw, j->shard_w0, j->shard_nw,
                                              j->kp, j->d_order, j->d_node_scratch, j->d_bits, j->d_bitsS,
                                              j->peer_push ? j->peers[cur] : no_peers, s);
      // the row sums (integer-ALU-bound) and the operand expansion (HBM-bound) both only read the bit
      // columns: they run side by side on the two compute streams unless an all-gather sits between
      cudaStream_t rs = j->sharded ? s : c->stream[1];
      if (!j->sharded) {
        JOB_CUDA(j, cudaEventRecord(j->ev_bits, s));
        JOB_CUDA(j, cudaStreamWaitEvent(rs, j->ev_bits, 0));
      }
      launches += launch_presence_rowsum_t(j->d_bits, j->B, j->nw, j->shard_w0, j->shard_nw, j->kp, j->d_lenq,
                                           j->i8 ? j->d_qam : nullptr, j->d_col_exp, j->d_scratch, j->d_r, j->e_min,
                                           j->intacc ? j->d_r_int : nullptr, rs);
      if (j->sharded) {
        // the one exchange step of the path: presence bit columns + row sums of every rank's
        // sample shard, concatenated over NVLink (2.5 GB at cfg4 instead of 120 GB of operands)
        // (bits-fed kernel: the sample-major bit rows are what is exchanged, same size)
        void* bufs[2] = {j->intacc ? static_cast<void*>(j->d_r_int) : static_cast<void*>(j->d_r),
                         j->bits_feed ? j->d_bitsS : j->d_bits};
        const size_t bytes[2] = {static_cast<size_t>(j->shard_nw) * 32 * sizeof(double),
                                 static_cast<size_t>(j->shard_nw) * j->kp * sizeof(uint32_t)};
        // peer_push: the bit columns already sit in every rank's bitsT (stored there by the embedding
        // kernel itself); the small all-gather of the row sums is also the point where the ranks meet
        std::string cerr;
        if (!comm_all_gather_inplace(c->comm, bufs, bytes, j->peer_push ? 1 : 2, s, &cerr)) return fail(j, FRC_ERR_CUDA, cerr);
        j->info.gather_bytes = static_cast<int64_t>(bytes[0] + bytes[1]) * (j->opts.world - 1);
      }
      if (!j->bits_feed)
        launches += launch_expand_operands_t(j->d_bits, j->nw, j->kp, j->np, j->i8, j->d_q0, j->d_q1, j->d_q2,
                                             j->d_P, j->d_Bh, j->d_Bl, j->d_need, s);
      if (!j->sharded) {
        JOB_CUDA(j, cudaEventRecord(j->ev_rsum, rs));
        JOB_CUDA(j, cudaStreamWaitEvent(s, j->ev_rsum, 0));
      }
    } else {
      launches += launch_embed_bits(j->dtree, j->level_ptr.data(), j->dcsr, j->d_bits, j->nw, s);
      launches += launch_presence_rowsum(j->d_bits, j->B, j->nw, j->d_lenq, j->d_r, j->d_scratch, s);
      launches += launch_expand_operands(j->d_bits, j->B, j->nw, j->kp, j->np,
                                         static_cast<const uint16_t*>(j->d_q1), static_cast<const uint16_t*>(j->d_q2),
                                         static_cast<uint16_t*>(j->d_P), static_cast<uint16_t*>(j->d_Bh),
                                         static_cast<uint16_t*>(j->d_Bl), s);
    }
    // bits written + read once per level pass, three bf16 operands written once (+ CSR cols)
    j->info.embed_bytes = 2LL * j->B * j->nw * 4 + 3LL * j->np * j->kp * (j->i8 ? 1 : 2) + 4LL * j->nnz;
  }
  JOB_CUDA(j, cudaGetLastError());
  JOB_CUDA(j, cudaEventRecord(j->ev_embed1, s));
  JOB_CUDA(j, cudaStreamWaitEvent(c->stream[1], j->ev_embed1, 0));
  j->info.kernel_launches += launches;
  j->embed_timed = false;
  return FRC_OK;
}

// Queue band mine[idx] into its slot.
int enqueue_band(frc_job* j, size_t idx) {
  frc_ctx* c = j->ctx;
  if (getenv("FRC_TRACE"))
    fprintf(stderr, "[host] enqueue band %zu at %.3f ms\n", idx,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - j->t_host0).count());
  const Band& b = j->bands[j->mine[idx]];
  Slot& sl = j->slots[idx % j->n_slots];
  cudaStream_t s = c->stream[idx % 2];
  sl.band = static_cast<int>(idx);
  int launches = 0;
  JOB_CUDA(j, cudaEventRecord(sl.k0, s));
  if (j->exact) {
    launches += launch_exact_pairs(j->d_E, j->dtree.length, j->B, j->np, j->weighted, b.first, b.count,
                                   sl.dev, s);
  } else if (j->weighted) {
    unsigned long long* cnt = j->d_flag_counts + idx;
    launches += launch_weighted_tiles(j->d_A, j->np, j->kp, j->d_lenf, j->prescale, j->d_W,
                                      j->d_tiles + b.tile_off, b.n_tiles, j->N, b.first, sl.dev, kFlagBelowW,
                                      sl.flagged, cnt, s);
    JOB_CUDA(j, cudaEventRecord(sl.k1, s));
    launches += launch_weighted_fixup(j->dcsr, j->dtree, j->opts.normalize ? j->d_total : nullptr, j->d_W,
                                      sl.flagged, cnt, sl.n_flagged_host, b.first, j->d_fix_ws, c->num_sms, sl.dev, s);
  } else {
    // no copy-engine work between the kernels: a memset or an 8-byte D2H would queue behind the
    // previous band's bulk D2H and stall this band (measured).  Counters are per band, zeroed once
    // per run; the fix-up kernel publishes the count through mapped pinned memory.
    unsigned long long* cnt = j->d_flag_counts + idx;
    if (j->bits_feed) {
      launches += launch_unweighted_bits(j->bo, j->d_tiles + b.tile_off, b.n_tiles, j->N, b.first, sl.dev,
                                         sl.flagged, cnt, c->num_sms, s);
      JOB_CUDA(j, cudaEventRecord(sl.k1, s));
      launches += launch_unweighted_fixup_bits(j->bo, sl.flagged, cnt, sl.n_flagged_host, b.first, sl.dev,
                                               c->num_sms, s);
    } else {
      launches += launch_unweighted_tc(j->tc, j->d_r, j->d_tiles + b.tile_off, b.n_tiles, j->N,
                                       b.first, sl.dev, kFlagBelow, sl.flagged, cnt, c->num_sms, j->tc_ctas, s);
      JOB_CUDA(j, cudaEventRecord(sl.k1, s));
      launches += launch_unweighted_fixup(j->tc, sl.flagged, cnt, sl.n_flagged_host, b.first, sl.dev,
                                          c->num_sms, s);
    }
  }
  JOB_CUDA(j, cudaGetLastError());
  if (j->exact) JOB_CUDA(j, cudaEventRecord(sl.k1, s));
  JOB_CUDA(j, cudaEventRecord(sl.k2, s));
  // bulk D2H runs on its own stream so the compute streams never hold copy-engine work
  // (a kernel queued behind a copy in the same stream cannot overlap that copy)
  cudaStream_t cs = c->stream[2];
  if (j->wire32) {
    // the band crosses PCIe as fp32 (wire.cu); frc_next widens it into sl.host
    *sl.n_bad_host = 0;
    launches += launch_narrow_band(sl.dev, sl.dev32, b.count, sl.n_bad_host, c->num_sms, s);
    JOB_CUDA(j, cudaGetLastError());
    JOB_CUDA(j, cudaEventRecord(sl.k3, s));
    JOB_CUDA(j, cudaStreamWaitEvent(cs, sl.k3, 0));
    JOB_CUDA(j, cudaMemcpyAsync(sl.host32, sl.dev32, sizeof(float) * b.count, cudaMemcpyDeviceToHost, cs));
    j->info.d2h_bytes += static_cast<int64_t>(sizeof(float)) * b.count;
    JOB_CUDA(j, cudaEventRecord(sl.done, cs));
  } else if (!(j->opts.flags & FRC_FLAG_NO_D2H) && sl.dev != sl.host) {
    cudaEvent_t ready = sl.k2;
    if (!j->exact) {
      // doubles on the bus carry the values the fp32 route would deliver (wire.cu: same bytes for any rank count)
      launches += launch_round_band(sl.dev, b.count, c->num_sms, s);
      JOB_CUDA(j, cudaGetLastError());
      JOB_CUDA(j, cudaEventRecord(sl.k3, s));
      ready = sl.k3;
    }
    JOB_CUDA(j, cudaStreamWaitEvent(cs, ready, 0));
    JOB_CUDA(j, cudaMemcpyAsync(sl.host, sl.dev, sizeof(double) * b.count, cudaMemcpyDeviceToHost, cs));
    j->info.d2h_bytes += static_cast<int64_t>(sizeof(double)) * b.count;
    JOB_CUDA(j, cudaEventRecord(sl.done, cs));
  } else {
    JOB_CUDA(j, cudaEventRecord(sl.done, s));
  }
  j->info.kernel_launches += launches;
  if (idx + 1 == j->mine.size()) {
    // last band queued: join all streams and stamp the end of the run on stream 0
    JOB_CUDA(j, cudaEventRecord(j->ev_join, c->stream[1]));
    JOB_CUDA(j, cudaStreamWaitEvent(c->stream[0], j->ev_join, 0));
    JOB_CUDA(j, cudaStreamWaitEvent(c->stream[0], sl.done, 0));
    JOB_CUDA(j, cudaEventRecord(j->ev_run1, c->stream[0]));
  }
  return FRC_OK;
}

int start_pairs(frc_job* j) {
  j->next_enqueue = j->next_deliver = 0;
  j->held_slot = -1;
  j->piece_off = j->piece_end = 0;
  j->info.pairs_ms = 0;
  j->info.fixup_ms = 0;
  j->info.run_ms = 0;
  j->run_timed = false;
  j->info.flagged_pairs = 0;
  j->info.d2h_bytes = 0;
  // keep one slot free for the band the caller is still reading
  while (j->next_enqueue < j->mine.size() && j->next_enqueue < static_cast<size_t>(j->n_slots - 1)) {
    int rc = enqueue_band(j, j->next_enqueue);
    if (rc) return rc;
    ++j->next_enqueue;
  }
  return FRC_OK;
}

void destroy_job(frc_job* j) {
  if (!j) return;
  if (j->ctx) {
    cudaSetDevice(j->ctx->device);
    for (auto s : j->ctx->stream) if (s) cudaStreamSynchronize(s);
    cudaGetLastError();
  }
  for (auto& sl : j->slots) {
    if (sl.k0) cudaEventDestroy(sl.k0);
    if (sl.k1) cudaEventDestroy(sl.k1);
    if (sl.k2) cudaEventDestroy(sl.k2);
    if (sl.k3) cudaEventDestroy(sl.k3);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  if (j->ev_h2d0) cudaEventDestroy(j->ev_h2d0);
  if (j->ev_h2d1) cudaEventDestroy(j->ev_h2d1);
  if (j->ev_embed0) cudaEventDestroy(j->ev_embed0);
  if (j->ev_embed1) cudaEventDestroy(j->ev_embed1);
  if (j->ev_run1) cudaEventDestroy(j->ev_run1);
  if (j->ev_join) cudaEventDestroy(j->ev_join);
  if (j->ev_bits) cudaEventDestroy(j->ev_bits);
  if (j->ev_rsum) cudaEventDestroy(j->ev_rsum);
  tc_operands_destroy(j->tc);
  bits_operands_destroy(j->bo);
  if (j->ctx) {
    j->ctx->dev.reset();
    j->ctx->pin.reset();
    j->ctx->sym.reset();
    j->ctx->in_use = false;
    if (j->own_ctx) frc_ctx_destroy(j->ctx);
  }
  delete j;
}

}  // namespace

extern "C" {

int frc_abi_version(void) { return FRC_ABI_VERSION; }

const char* frc_last_error(const frc_job_t* job) { return job ? job->err.c_str() : g_create_error.c_str(); }

int frc_ctx_create(int32_t device, frc_ctx_t** out) {
  if (!out) { g_create_error = "frc_ctx_create: out is NULL"; return FRC_ERR_ARG; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
    return FRC_ERR_CUDA;
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= count) { g_create_error = "device ordinal out of range"; return FRC_ERR_ARG; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return FRC_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "libfrcfrc_cuda is built for sm_100a (B200) only; device " + std::to_string(device) +
                     " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + " (" + prop.name + ")";
    return FRC_ERR_UNSUPPORTED;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return FRC_ERR_CUDA;
  }
  frc_ctx* c = new frc_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->dev.pinned = false; c->dev.min_block = 64u << 20;
  c->sym.pinned = false; c->sym.min_block = 32u << 20;
  c->pin.pinned = true;  c->pin.min_block = 8u << 20;
  {
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    int n = std::max(1, std::min(16, hw)) - 1;
    if (const char* e = getenv("FRC_HOST_THREADS")) n = std::max(1, atoi(e)) - 1;
    c->pool.reset(new Pool(n));
    c->stamps.resize(n + 1);
  }
  for (auto& s : c->stream)
    if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) {
      g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
      frc_ctx_destroy(c);
      return FRC_ERR_CUDA;
    }
  std::string terr;
  if (!tc_setup(&terr)) { g_create_error = terr; frc_ctx_destroy(c); return FRC_ERR_CUDA; }
  weighted_setup();
  if ((e = cudaGetLastError()) != cudaSuccess) {
    g_create_error = std::string("kernel attribute setup: ") + cudaGetErrorString(e);
    frc_ctx_destroy(c);
    return FRC_ERR_CUDA;
  }
  *out = c;
  return FRC_OK;
}

void frc_ctx_destroy(frc_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->pool.reset();
  for (auto s : c->stream) if (s) cudaStreamSynchronize(s);
  for (size_t b = 0; b < c->sym_peer.size(); ++b)
    for (size_t r = 0; r < c->sym_peer[b].size(); ++r)
      if (c->sym_peer[b][r] && c->sym_peer[b][r] != c->sym.blocks[b].p) cudaIpcCloseMemHandle(c->sym_peer[b][r]);
  c->sym_peer.clear();
  c->sym.release();
  comm_destroy(c->comm);
  c->comm = nullptr;
  for (auto s : c->stream) if (s) cudaStreamDestroy(s);
  c->dev.release();
  c->pin.release();
  cudaGetLastError();
  delete c;
}

int frc_comm_unique_id(char* id) {
  if (!id) { g_create_error = "frc_comm_unique_id: id is NULL"; return FRC_ERR_ARG; }
  std::string err;
  if (!comm_unique_id(id, &err)) { g_create_error = err; return FRC_ERR_UNSUPPORTED; }
  return FRC_OK;
}

int frc_ctx_comm_init(frc_ctx_t* ctx, const char* id, int32_t rank, int32_t world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) {
    g_create_error = "frc_ctx_comm_init: bad argument";
    return FRC_ERR_ARG;
  }
  if (ctx->in_use) { g_create_error = "context has a live job"; return FRC_ERR_STATE; }
  cudaSetDevice(ctx->device);
  comm_destroy(ctx->comm);
  ctx->comm = nullptr;
  std::string err;
  ctx->comm = comm_create(id, rank, world, &err);
  if (!ctx->comm) { g_create_error = err; return FRC_ERR_CUDA; }
  return FRC_OK;
}

int frc_create(frc_ctx_t* ctx, const frc_tree_t* tree, const frc_csr_t* abnd, const frc_opts_t* opts,
               frc_job_t** out) {
  g_create_error.clear();
  if (!out) { g_create_error = "frc_create: out is NULL"; return FRC_ERR_ARG; }
  *out = nullptr;
  if (!tree || !abnd || !opts) { g_create_error = "frc_create: NULL argument"; return FRC_ERR_ARG; }
  frc_job* j = new frc_job();
  bool csr_workers_busy = false;  // the table is validated + staged by the pool while this thread handles the tree
  auto bail = [&](int rc) {
    if (csr_workers_busy) { j->ctx->pool->wait(); csr_workers_busy = false; }
    g_create_error = j->err; destroy_job(j); return rc;
  };
  const bool trace = getenv("FRC_TRACE") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[frc_create] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - tp0).count());
    tp0 = now;
  };
#define CREATE_CUDA(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cudaGetLastError();                                                                    \
      return bail(fail(j, _e == cudaErrorMemoryAllocation ? FRC_ERR_OOM : FRC_ERR_CUDA,      \
                       std::string(#expr) + ": " + cudaGetErrorString(_e)));                 \
    }                                                                                        \
  } while (0)
  j->opts = *opts;

  // ---------------------------------------------------------- validate options
  if (opts->mode != FRC_UNWEIGHTED && opts->mode != FRC_WEIGHTED) return bail(fail(j, FRC_ERR_ARG, "bad mode"));
  if (opts->path < FRC_PATH_AUTO || opts->path > FRC_PATH_EXACT) return bail(fail(j, FRC_ERR_ARG, "bad path"));
  if (opts->normalize != 0 && opts->normalize != 1) return bail(fail(j, FRC_ERR_ARG, "bad normalize"));
  if (opts->normalize == 0 && opts->mode != FRC_WEIGHTED)  // frcfrc.go:84-86
    return bail(fail(j, FRC_ERR_ARG, "-l can only be used with weighted unifrac"));
  int world = opts->world <= 0 ? 1 : opts->world;
  int rank = opts->world <= 0 ? 0 : opts->rank;
  if (rank < 0 || rank >= world) return bail(fail(j, FRC_ERR_ARG, "rank outside [0, world)"));
  if (opts->band_rows < 0) return bail(fail(j, FRC_ERR_ARG, "band_rows < 0"));
  j->weighted = opts->mode == FRC_WEIGHTED;

  // ------------------------------------------------------------- validate tree
  const int32_t B = tree->n_nodes;
  if (B < 1 || !tree->parent || !tree->length) return bail(fail(j, FRC_ERR_ARG, "empty tree"));
  if (tree->parent[0] != -1) return bail(fail(j, FRC_ERR_ARG, "parent[0] must be -1 (root has pre-order id 0)"));
  // ------------------------------------------------------------ validate table
  const int64_t N = abnd->n_samples;
  if (N < 0 || (N > 0 && !abnd->row_ptr)) return bail(fail(j, FRC_ERR_ARG, "bad abundance table"));
  if (N > (1LL << 31) - 256) return bail(fail(j, FRC_ERR_UNSUPPORTED, "too many samples"));
  const int64_t nnz = N > 0 ? abnd->row_ptr[N] : 0;
  if (N > 0 && abnd->row_ptr[0] != 0) return bail(fail(j, FRC_ERR_ARG, "row_ptr[0] != 0"));
  if (nnz < 0 || (nnz > 0 && (!abnd->col || !abnd->val))) return bail(fail(j, FRC_ERR_ARG, "bad abundance table"));
  for (int64_t s = 0; s < N; ++s)
    if (abnd->row_ptr[s + 1] < abnd->row_ptr[s] || abnd->row_ptr[s + 1] > nnz)
      return bail(fail(j, FRC_ERR_ARG, "row_ptr is not monotone"));
  // (entries are validated while they are copied into the pinned staging buffer, below)

  // -------------------------------------------------------------------- context
  int rc = FRC_OK;
  if (ctx) {
    if (ctx->in_use) return bail(fail(j, FRC_ERR_STATE, "context already has a live job"));
    j->ctx = ctx;
  } else {
    rc = frc_ctx_create(opts->device, &j->ctx);
    if (rc) { j->err = g_create_error; j->ctx = nullptr; return bail(rc); }
    j->own_ctx = true;
  }
  frc_ctx* c = j->ctx;
  c->in_use = true;
  CREATE_CUDA(cudaSetDevice(c->device));

  // ------------------------------------ table: validate + stage on the worker pool
  // The CSR entries (the bulk of the input) are checked and copied into pinned staging memory by the
  // pool while this thread validates the tree, plans the operand columns and bands, and sets up
  // the device side; the two meet just before the embedding is queued.
  const int64_t n_pairs = N >= 2 ? tri(N) : 0;
  if (opts->path == FRC_PATH_EXACT) j->exact = true;
  else if (opts->path == FRC_PATH_FAST) j->exact = false;
  else j->exact = n_pairs == 0 || (static_cast<double>(n_pairs) * B <= static_cast<double>(kExactWorkLimit));
  const bool need_val = j->exact || j->weighted;
  struct Seg { size_t off, bytes; };
  size_t csr_total = 0;
  auto cseg = [&](size_t bytes) { Seg s{csr_total, bytes}; csr_total += (bytes + 255) & ~size_t(255); return s; };
  const int32_t kp_pad = static_cast<int32_t>(round_up(B, kKBlock));
  const Seg s_rowptr = cseg(sizeof(int64_t) * (N + 1)), s_col = cseg(sizeof(int32_t) * nnz),
            s_val = cseg(need_val ? sizeof(double) * nnz : 0),
            // tree topology, filled by one more pool share (tree_share below)
            s_parent = cseg(sizeof(int32_t) * B), s_len = cseg(sizeof(double) * B),
            s_cptr = cseg(sizeof(int32_t) * (B + 1)), s_cidx = cseg(sizeof(int32_t) * B),
            s_lvl = cseg(sizeof(int32_t) * B), s_lpar = cseg(sizeof(int32_t) * B),
            s_lptr = cseg(sizeof(int32_t) * (static_cast<size_t>(B) + 2)), s_lenf = cseg(sizeof(float) * kp_pad);
  char* stage_csr = pin_alloc<char>(j, csr_total, &rc);
  if (!stage_csr) return bail(rc);
  char* d_csr = dev_alloc<char>(j, csr_total, &rc);
  if (!d_csr) return bail(rc);
  {
    int64_t* rp = reinterpret_cast<int64_t*>(stage_csr + s_rowptr.off);
    if (N > 0) memcpy(rp, abnd->row_ptr, sizeof(int64_t) * (N + 1)); else rp[0] = 0;
  }
  // One task list for the pool: two tree tasks (children lists; pre-order check + level order), then the
  // CSR entries in chunks, all handed out through one counter; the calling thread joins in when its own
  // part is done, so the split adapts to any pool size (4 threads per rank on a 32-core 8-GPU host).
  const int n_workers = c->pool->size() - 1;
  int32_t tree_H = -1;                      // height of the tree, from the levels share
  int32_t bad_parent = 0, bad_order = 0;    // first offending node ids, from the levels share
  bool children_ok = false;
  auto tree_children_share = [&, stage_csr]() {
    const int32_t* parent = tree->parent;
    std::vector<int32_t> child_cnt(B, 0);
    for (int32_t v = 1; v < B; ++v) {
      const int32_t p = parent[v];
      if (p < 0 || p >= v) return;  // (reported by the levels share)
      child_cnt[p]++;
    }
    memcpy(stage_csr + s_parent.off, parent, sizeof(int32_t) * B);
    memcpy(stage_csr + s_len.off, tree->length, sizeof(double) * B);
    int32_t* cptr = reinterpret_cast<int32_t*>(stage_csr + s_cptr.off);
    int32_t* cidx = reinterpret_cast<int32_t*>(stage_csr + s_cidx.off);
    cptr[0] = 0;
    for (int32_t v = 0; v < B; ++v) cptr[v + 1] = cptr[v] + child_cnt[v];
    std::vector<int32_t>& fill = child_cnt;  // reused: next free slot of every node's child list
    for (int32_t v = 0; v < B; ++v) fill[v] = cptr[v];
    for (int32_t v = 1; v < B; ++v) cidx[fill[parent[v]]++] = v;  // ascending id = file order
    float* lf = reinterpret_cast<float*>(stage_csr + s_lenf.off);
    for (int32_t v = 0; v < kp_pad; ++v) lf[v] = v < B ? static_cast<float>(tree->length[v]) : 0.f;
    children_ok = true;
  };
  auto tree_levels_share = [&, stage_csr]() {
    const int32_t* parent = tree->parent;
    // pre-order check: parent[v] must lie on the path root..v-1, i.e. be the node of its depth on the
    // current root-to-(v-1) path.  Branch-free per node (the stack-popping form cost 14 ns per node in
    // mispredictions).
    std::vector<int32_t> depth(B, 0), on_path(B, 0), height(B, 0);  // on_path[d] = node at depth d of the current path
    for (int32_t v = 1; v < B; ++v) {
      const int32_t p = parent[v];
      if (p < 0 || p >= v) { bad_parent = v; return; }
      const int32_t d = depth[p];
      if ((d > depth[v - 1] || on_path[d] != p) && !bad_order) bad_order = v;
      depth[v] = d + 1;
      on_path[d + 1] = v;
    }
    if (bad_order) return;
    for (int32_t v = B - 1; v >= 1; --v) height[parent[v]] = std::max(height[parent[v]], height[v] + 1);
    const int32_t H = height[0];
    j->level_ptr.assign(H + 2, 0);
    for (int32_t v = 0; v < B; ++v) j->level_ptr[height[v] + 1]++;
    for (int32_t h = 0; h <= H; ++h) j->level_ptr[h + 1] += j->level_ptr[h];
    int32_t* lvl = reinterpret_cast<int32_t*>(stage_csr + s_lvl.off);
    std::vector<int32_t>& lfill = depth;  // reused: next free slot of every level
    for (int32_t h = 0; h <= H; ++h) lfill[h] = j->level_ptr[h];
    for (int32_t v = 0; v < B; ++v) lvl[lfill[height[v]]++] = v;
    int32_t* lpar = reinterpret_cast<int32_t*>(stage_csr + s_lpar.off);
    for (int32_t k = 0; k < B; ++k) lpar[k] = lvl[k] ? parent[lvl[k]] : 0;
    memcpy(stage_csr + s_lptr.off, j->level_ptr.data(), sizeof(int32_t) * (H + 2));
    tree_H = H;
  };
  // The entries are cut into many more chunks than workers and handed out through a counter: a worker
  // on a busy or slow core (the slowest static share took 1.5-2.4x the median) just takes fewer.
  const int csr_chunks = nnz >= 65536 ? static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(c->pool->size() * 8LL, nnz / 8192))) : 1;
  std::atomic<int> task_next{0};
  std::vector<std::string> csr_errs(csr_chunks);
  std::vector<double> csr_us(c->pool->size(), 0.0);
  // duplicate detection: stamp[leaf] = epoch-tagged sample id; the arrays persist in the
  // context, so nothing is cleared between jobs
  const int64_t stamp_tag = c->stamp_epoch;
  c->stamp_epoch += N + 1;
  const std::function<void(int)> csr_work = [&, stage_csr](int t) {
    const auto w0 = std::chrono::steady_clock::now();
    int32_t* dcol = reinterpret_cast<int32_t*>(stage_csr + s_col.off);
    double* dval = need_val ? reinterpret_cast<double*>(stage_csr + s_val.off) : nullptr;
    const int TS = csr_chunks;
    // a leaf listed twice in a row only matters when values are used (presence is an OR)
    const bool check_dup = need_val;
    const int32_t* parent = tree->parent;
    auto row_at = [&](int64_t target) {
      return static_cast<int64_t>(std::lower_bound(abnd->row_ptr, abnd->row_ptr + N, target) - abnd->row_ptr);
    };
    std::vector<int64_t>& stamp = c->stamps[t];
    if (check_dup && static_cast<int32_t>(stamp.size()) < B) stamp.assign(B, -1);
    for (int task = task_next.fetch_add(1, std::memory_order_relaxed); task < TS + 2; task = task_next.fetch_add(1, std::memory_order_relaxed)) {
    if (task == 0) { tree_levels_share(); continue; }
    if (task == 1) { tree_children_share(); continue; }
    const int ch = task - 2;
    const int64_t s0 = ch == 0 ? 0 : row_at(nnz * ch / TS), s1 = ch == TS - 1 ? N : row_at(nnz * (ch + 1) / TS);
    bool failed = false;
    for (int64_t s = s0; s < s1 && !failed; ++s) {
      const int64_t b = abnd->row_ptr[s], e = abnd->row_ptr[s + 1];
      const int64_t mark_s = stamp_tag + s;
      for (int64_t k = b; k < e; ++k) {
        const int32_t cc = abnd->col[k];
        const double v = abnd->val[k];
        const char* what = nullptr;
        if (cc < 0 || cc >= B) what = "node id out of range";
        // pre-order ids: the first child of a node is the next id (the numbering itself is being
        // checked by the calling thread at this moment; a bad tree fails the call before this result counts)
        else if (cc + 1 < B && parent[cc + 1] == cc) what = "node is not a leaf";
        else if (!(v > 0) || std::isinf(v)) what = "bad value";
        else if (check_dup && stamp[cc] == mark_s) what = "leaf listed twice";
        if (what) {
          csr_errs[ch] = "sample #" + std::to_string(s + 1) + ": entry " + std::to_string(k - b + 1) + " (node " +
                        std::to_string(cc) + "): " + what;
          failed = true;
          break;
        }
        if (check_dup) stamp[cc] = mark_s;
#if defined(__x86_64__)
        // streaming stores: the staging buffer is read next by the DMA engine, not by a core
        _mm_stream_si32(dcol + k, cc);
        if (dval) _mm_stream_si64(reinterpret_cast<long long*>(dval + k), static_cast<long long>(__builtin_bit_cast(int64_t, v)));
#else
        dcol[k] = cc;
        if (dval) dval[k] = v;
#endif
      }
    }
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
    csr_us[t] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
  };
  if (n_workers > 0) {
    csr_workers_busy = true;
    c->pool->start(n_workers, csr_work);
  }
  mark("table checks, context, CSR workers started");

  // ------------------------------------------------------------- branch lengths (the walk runs in the pool)
  bool neg_len = false, bad_len = false;
  for (int32_t v = 0; v < B; ++v) {
    double l = tree->length[v];
    if (!(l == l) || std::isinf(l)) bad_len = true;
    else if (l < 0) neg_len = true;
  }
  mark("branch lengths checked");
  // --------------------------------------------------------------- choose path
  j->N = N; j->B = B; j->nnz = nnz;
  j->sharded = (opts->flags & FRC_FLAG_SHARD_EMBED) != 0 && world > 1;
  if (j->sharded) {
    if (!ctx || !ctx->comm || comm_world(ctx->comm) != world || comm_rank(ctx->comm) != rank)
      return bail(fail(j, FRC_ERR_STATE, "FRC_FLAG_SHARD_EMBED needs a context whose communicator "
                                         "(frc_ctx_comm_init) matches opts->rank / opts->world"));
  }
  // sharded: every rank owns np / world samples, a whole number of tiles
  j->np = std::max<int64_t>(kTile, round_up(N, j->sharded ? kTile * static_cast<int64_t>(world) : kTile));
  j->kp = static_cast<int32_t>(round_up(B, kKBlock));
  j->nw = static_cast<int32_t>(j->np / 32);
  if (!j->exact && bad_len)
    return bail(fail(j, FRC_ERR_UNSUPPORTED, "non-finite branch length: only the exact path handles it"));
  j->prescale = !neg_len;
  if (j->sharded && j->exact) j->sharded = false;  // (the exact path rebuilds the embedding per rank)
  j->shard_nw = j->sharded ? j->nw / world : j->nw;
  j->shard_w0 = j->sharded ? rank * j->shard_nw : 0;
  j->info.path_taken = j->exact ? FRC_PATH_EXACT : FRC_PATH_FAST;
  j->info.n_pairs_total = n_pairs;
  j->info.n_nodes_padded = j->exact ? B : j->kp;

  // ---------------------------------------------------------------------- bands
  {
    const char* e = getenv("FRC_TC_CTAS");
    // CTA pairs (cta_group::2) by default; FRC_TC_CTAS=1 selects the single-CTA kernel
    j->tc_ctas = (!(e && atoi(e) == 1) && !j->exact && !j->weighted) ? 2 : 1;
  }
  // ------------------------------------- operand columns of the fast unweighted path
  // Column k of the K-major operands holds node col_order[k] (-1: padding).  bf16: identity
  // order, uniform accumulation chunks.  u8: nodes grouped by binade pairs of their length so
  // that one power-of-two scale per chunk leaves every length a 21..24-bit integer a * m
  // (zero-length nodes contribute nothing and get no column).
  std::vector<int32_t> col_order, col_exp, chunk_end, chunk_shift;
  std::vector<double> len_col, chunk_scale;
  if (!j->exact && !j->weighted) {
    { const char* e = getenv("FRC_EMBED_LEVELS"); j->fused_embed = !(e && atoi(e) == 1); }
    if (j->sharded && !j->fused_embed)
      return bail(fail(j, FRC_ERR_UNSUPPORTED, "FRC_EMBED_LEVELS=1 cannot build a sample shard"));
    const char* ek = getenv("FRC_UW_KERNEL");  // "bf16" forces the bf16 hi/lo kernel
    j->i8 = j->tc_ctas == 2 && j->fused_embed && !neg_len && !(opts->flags & FRC_FLAG_UW_BF16) &&
            !(ek && strcmp(ek, "bf16") == 0);
    if (j->i8) {
      constexpr int kGroups = 24, kBlockCols = 128, kMaxChunkBlocks = 128;
      // floor(log2(l)) of a positive finite double straight from its exponent field (denormals: -1023,
      // they all land in the last group)
      auto ilog = [](double l) {
        uint64_t u;
        memcpy(&u, &l, 8);
        return static_cast<int>((u >> 52) & 0x7FF) - 1023;
      };
      int e_max = INT32_MIN;
      for (int32_t v = 0; v < B; ++v)
        if (tree->length[v] > 0) e_max = std::max(e_max, ilog(tree->length[v]));
      int gb = 3;  // binades per group: x = len * 2^-e in [2^20, 2^23), >= 110 candidate factors a
      if (const char* e = getenv("FRC_U8_GROUP_BINADES")) gb = std::max(1, std::min(16, atoi(e)));
      // group of every node, through a table over the exponent distance (an integer division per node
      // and pass was most of this plan's 0.15 ms at 2e4 nodes); 255 = no column (length 0)
      uint8_t group_tab[2112];
      for (int d = 0; d < 2112; ++d) group_tab[d] = static_cast<uint8_t>(std::min(kGroups - 1, d / gb));
      std::vector<uint8_t> grp(B);
      int32_t cnt[kGroups + 1] = {0};
      for (int32_t v = 0; v < B; ++v) {
        const double l = tree->length[v];
        const uint8_t g = l > 0 ? group_tab[e_max - ilog(l)] : static_cast<uint8_t>(kGroups);
        grp[v] = g;
        cnt[g]++;
      }
      // Groups -> accumulation chunks.  Every chunk costs a TMEM drain (>= 2k cycles: 128 KB at
      // 64 B/clk) that only hides behind a long enough MMA run, so a group opens a chunk of its own
      // only when it is large; smaller groups below it join the current chunk at that chunk's
      // (larger) scale.  Their lengths then keep fewer significant bits; k_quantize_lengths sums the
      // absolute error of every column that misses 4e-6 relative, and pairs whose unique length is
      // too small for that sum to be harmless are recomputed exactly (flag_u).
      constexpr int32_t kMinChunkCols = 1024;
      int chunk_of[kGroups], chunk_exp[kGroups], n_ch = 0;
      int32_t ch_cols[kGroups] = {0};
      for (int g = 0; g < kGroups; ++g) {
        if (cnt[g] == 0) { chunk_of[g] = -1; continue; }
        if (n_ch == 0 || cnt[g] >= kMinChunkCols) {
          chunk_exp[n_ch] = e_max - gb * g - 22;  // top of the group: len < 2^(e_max-gb*g+1) -> x < 2^23
          ++n_ch;
        }
        chunk_of[g] = n_ch - 1;
        ch_cols[n_ch - 1] += cnt[g];
      }
      // Chunks are laid out along K from the largest to the smallest (the order is free): the MMA
      // run of a large chunk then covers the ratio epilogue of the previous tile plus the drain of
      // the small chunk before it, and the two TMEM buffers are never both waiting for the epilogue
      // warps (with the small chunk first the third chunk of a tile needed its buffer ~57k cycles
      // after the previous tile ended, exactly when the ratio epilogue released it).
      int ch_order[kGroups];
      for (int c = 0; c < n_ch; ++c) ch_order[c] = c;
      std::stable_sort(ch_order, ch_order + n_ch, [&](int a, int b) { return ch_cols[a] > ch_cols[b]; });
      int32_t ch_off[kGroups + 1] = {0}, ch_at[kGroups], fill[kGroups];
      for (int k = 0; k < n_ch; ++k) {
        ch_at[ch_order[k]] = ch_off[k];
        ch_off[k + 1] = ch_off[k] + static_cast<int32_t>(round_up(ch_cols[ch_order[k]], kBlockCols));
      }
      j->kp = std::max<int32_t>(ch_off[n_ch], kBlockCols);
      col_order.assign(j->kp, -1); col_exp.assign(j->kp, 0); len_col.assign(j->kp, 0.0);
      {  // columns of a chunk: its groups from large to small lengths, node id order inside a group
        int32_t next[kGroups];
        for (int c = 0; c < n_ch; ++c) next[c] = ch_at[c];
        for (int g = 0; g < kGroups; ++g)
          if (chunk_of[g] >= 0) { fill[g] = next[chunk_of[g]]; next[chunk_of[g]] += cnt[g]; }
      }
      for (int32_t v = 0; v < B; ++v) {
        if (grp[v] == kGroups) continue;
        const int32_t k = fill[grp[v]]++;
        col_order[k] = v; len_col[k] = tree->length[v];
      }
      for (int k = 0; k < n_ch; ++k) {
        const int c = ch_order[k];
        for (int32_t q = ch_off[k]; q < ch_off[k + 1]; ++q) col_exp[q] = chunk_exp[c];
        for (int32_t kb = ch_off[k] / kBlockCols; kb < ch_off[k + 1] / kBlockCols;) {
          kb = std::min(ch_off[k + 1] / kBlockCols, kb + kMaxChunkBlocks);  // plane sums stay < 2^31
          chunk_end.push_back(kb);
          chunk_scale.push_back(std::ldexp(1.0, chunk_exp[c]));
        }
      }
      if (chunk_end.empty()) { chunk_end.push_back(1); chunk_scale.push_back(1.0); }
      {  // integer mode when all chunk scales lie within 2^16 (sums then stay below 2^60)
        int lo = INT32_MAX, hi = INT32_MIN;
        for (double sc : chunk_scale) { const int e = std::ilogb(sc); lo = std::min(lo, e); hi = std::max(hi, e); }
        const char* ev = getenv("FRC_U8_ACC");  // "f64" forces the fp64 accumulation path (tests)
        j->intacc = hi - lo <= 16 && !(ev && strcmp(ev, "f64") == 0);
        j->e_min = lo;
        for (double sc : chunk_scale) chunk_shift.push_back(std::ilogb(sc) - lo);
        const char* ef = getenv("FRC_UW_FEED");  // "bits": expand the operand tiles inside the pair kernel
        j->bits_feed = j->intacc && (((opts->flags & FRC_FLAG_UW_BITS) != 0) || (ef && strcmp(ef, "bits") == 0));
      }
    } else {
      col_order.assign(j->kp, -1); len_col.assign(j->kp, 0.0);
      for (int32_t v = 0; v < B; ++v) { col_order[v] = v; len_col[v] = tree->length[v]; }
      const int32_t nkb = j->kp / kKBlock, per = tc_chunk_kblocks();
      for (int32_t kb = 0; kb < nkb;) { kb = std::min(nkb, kb + per); chunk_end.push_back(kb); chunk_scale.push_back(1.0); }
    }
    j->info.n_nodes_padded = j->kp;
    j->info.operand_kind = j->i8 ? (j->bits_feed ? 3 : 2) : 1;
  }
  const std::vector<int64_t> brows = band_boundaries(N, opts->band_rows, world, !(opts->flags & FRC_FLAG_NO_D2H));
  for (size_t kb = 0; kb + 1 < brows.size(); ++kb) {
    Band b;
    b.row0 = brows[kb]; b.row1 = brows[kb + 1];
    b.first = b.row0 >= 2 ? tri(b.row0) : 0;
    b.count = tri(b.row1) - b.first;
    b.tile_off = 0; b.n_tiles = 0;
    if (b.count > 0) j->bands.push_back(b);
  }
  int64_t max_band = 1;
  const std::vector<int> owners = band_owners(brows, world);
  for (size_t k = 0; k < j->bands.size(); ++k)
    if (owners[k] == rank) {
      j->mine.push_back(static_cast<int>(k));
      j->info.n_pairs_mine += j->bands[k].count;
      max_band = std::max(max_band, j->bands[k].count);
    }
  if (max_band >= (1LL << 32)) return bail(fail(j, FRC_ERR_UNSUPPORTED, "band too large; lower band_rows"));
  j->info.n_bands_total = static_cast<int32_t>(j->bands.size());
  j->info.n_bands_mine = static_cast<int32_t>(j->mine.size());
  if (!j->exact)
    for (int k : j->mine) {
      Band& b = j->bands[k];
      b.tile_off = static_cast<int32_t>(j->tiles.size());
      int32_t t0 = static_cast<int32_t>(b.row0 / kTile), t1 = static_cast<int32_t>((b.row1 - 1) / kTile);
      // column-major inside the band: CTAs running together share the few
      // i-tiles of the band and neighbouring j-tiles (L2 reuse of both operands)
      if (j->tc_ctas == 2) {
        // pair tiles: (ti, tj) and (ti, tj+1) with tj even; needed when tj <= ti.  The clusters
        // that run together (74 on a B200) take consecutive list entries, so the list walks
        // compact super-tiles of ~9 tile rows x ~8 tile-pair columns: the wave then shares
        // 9 + 8 operand panels instead of 2 + 37 and its working set fits the 126 MB L2.
        const int32_t rows_blk = std::min<int32_t>(9, t1 - t0 + 1);
        const int32_t cols_blk = std::max<int32_t>(1, 74 / rows_blk);  // in tile pairs
        for (int32_t tb = t0; tb <= t1; tb += rows_blk) {
          const int32_t te = std::min(t1, tb + rows_blk - 1);
          for (int32_t pb = 0; 2 * pb <= te; pb += cols_blk)
            for (int32_t tp = pb; tp < pb + cols_blk && 2 * tp <= te; ++tp)
              for (int32_t ti = std::max(2 * tp, tb); ti <= te; ++ti) j->tiles.push_back({ti, 2 * tp});
        }
      } else {
        for (int32_t tj = 0; tj <= t1; ++tj)
          for (int32_t ti = std::max(tj, t0); ti <= t1; ++ti) j->tiles.push_back({ti, tj});
      }
      b.n_tiles = static_cast<int32_t>(j->tiles.size()) - b.tile_off;
    }

  // which operand rows this rank's tiles read (fast unweighted, world > 1): a tile (ti, tj) reads the
  // Bh / Bl rows of its row samples and the A rows of its column samples (both tiles of a pair)
  std::vector<uint8_t> need_blocks;
  if (!j->exact && !j->weighted && world > 1 && j->fused_embed) {
    need_blocks.assign(static_cast<size_t>((j->np + 255) / 256), 0);
    const int32_t pairw = j->tc_ctas == 2 ? 2 : 1;
    for (const Tile& t : j->tiles) {
      need_blocks[t.ti / 2] |= 2;
      for (int32_t c = 0; c < pairw; ++c)
        if ((static_cast<int64_t>(t.tj) + c) * kTile < j->np) need_blocks[(t.tj + c) / 2] |= 1;
    }
  }
  mark("row_ptr check + bands/tiles");
  // ------------------------------------------------- pack + upload the inputs
  // (the column plan and the tile lists; the CSR block and the tree topology are staged by the pool)
  size_t total = 0;
  auto seg = [&](size_t bytes) { Seg s{total, bytes}; total += (bytes + 255) & ~size_t(255); return s; };
  Seg s_q0 = seg(j->kp), s_hi = seg(sizeof(uint16_t) * j->kp), s_lo = seg(sizeof(uint16_t) * j->kp),
      s_lenq = seg(sizeof(double) * j->kp), s_tiles = seg(sizeof(Tile) * j->tiles.size()),
      s_order = seg(sizeof(int32_t) * col_order.size()), s_cexp = seg(sizeof(int32_t) * col_exp.size()),
      s_lcol = seg(sizeof(double) * len_col.size()), s_cend = seg(sizeof(int32_t) * chunk_end.size()),
      s_cscale = seg(sizeof(double) * chunk_scale.size()), s_need = seg(need_blocks.size()),
      s_cshift = seg(sizeof(int32_t) * chunk_shift.size());
  char* stage = pin_alloc<char>(j, total, &rc);
  if (!stage) return bail(rc);
  j->d_inputs = dev_alloc<char>(j, total, &rc);
  if (!j->d_inputs) return bail(rc);
  j->input_bytes = total + csr_total;
  {
    {
    uint16_t* hi = reinterpret_cast<uint16_t*>(stage + s_hi.off);
    uint16_t* lo = reinterpret_cast<uint16_t*>(stage + s_lo.off);
    double* lq = reinterpret_cast<double*>(stage + s_lenq.off);
    if (!j->i8) {  // (u8: k_quantize_lengths fills q0/q1/q2 and lenq on the device)
      for (int32_t v = 0; v < B; ++v) {
        double l = tree->length[v];
        hi[v] = bf16_rn(static_cast<float>(l));
        double res = l - static_cast<double>(bf16_to_float(hi[v]));
        lo[v] = (l == l && !std::isinf(l)) ? bf16_rn(static_cast<float>(res)) : 0;
        lq[v] = static_cast<double>(bf16_to_float(hi[v])) + static_cast<double>(bf16_to_float(lo[v]));
      }
      for (int32_t v = B; v < j->kp; ++v) { hi[v] = 0; lo[v] = 0; lq[v] = 0.0; }
    }
    if (!col_order.empty()) memcpy(stage + s_order.off, col_order.data(), s_order.bytes);
    if (!col_exp.empty()) memcpy(stage + s_cexp.off, col_exp.data(), s_cexp.bytes);
    if (!len_col.empty()) memcpy(stage + s_lcol.off, len_col.data(), s_lcol.bytes);
    if (!chunk_end.empty()) memcpy(stage + s_cend.off, chunk_end.data(), s_cend.bytes);
    if (!chunk_scale.empty()) memcpy(stage + s_cscale.off, chunk_scale.data(), s_cscale.bytes);
    if (!need_blocks.empty()) memcpy(stage + s_need.off, need_blocks.data(), s_need.bytes);
    if (!chunk_shift.empty()) memcpy(stage + s_cshift.off, chunk_shift.data(), s_cshift.bytes);
    if (!j->tiles.empty()) memcpy(stage + s_tiles.off, j->tiles.data(), sizeof(Tile) * j->tiles.size());
    }
  }
  mark("column plan + tiles staged");
  CREATE_CUDA(cudaEventCreate(&j->ev_h2d0));
  CREATE_CUDA(cudaEventCreate(&j->ev_h2d1));
  CREATE_CUDA(cudaEventCreate(&j->ev_embed0));
  CREATE_CUDA(cudaEventCreate(&j->ev_embed1));
  CREATE_CUDA(cudaEventCreate(&j->ev_run1));
  CREATE_CUDA(cudaEventCreateWithFlags(&j->ev_join, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&j->ev_bits, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&j->ev_rsum, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventRecord(j->ev_h2d0, c->stream[0]));
  CREATE_CUDA(cudaMemcpyAsync(j->d_inputs, stage, total, cudaMemcpyHostToDevice, c->stream[0]));
  j->info.h2d_bytes = static_cast<int64_t>(total + csr_total);
  char* d = j->d_inputs;
  j->dcsr.n_samples = N; j->dcsr.nnz = nnz;
  j->dcsr.row_ptr = reinterpret_cast<int64_t*>(d_csr + s_rowptr.off);
  j->dcsr.col = reinterpret_cast<int32_t*>(d_csr + s_col.off);
  j->dcsr.val = need_val ? reinterpret_cast<double*>(d_csr + s_val.off) : nullptr;
  j->dtree.n_nodes = B; j->dtree.height = 0;  // (height: set when the tree share is joined)
  j->dtree.parent = reinterpret_cast<int32_t*>(d_csr + s_parent.off);
  j->dtree.length = reinterpret_cast<double*>(d_csr + s_len.off);
  j->dtree.child_ptr = reinterpret_cast<int32_t*>(d_csr + s_cptr.off);
  j->dtree.child_idx = reinterpret_cast<int32_t*>(d_csr + s_cidx.off);
  j->dtree.level_nodes = reinterpret_cast<int32_t*>(d_csr + s_lvl.off);
  j->d_q0 = d + s_q0.off;
  j->d_q1 = d + s_hi.off;
  j->d_q2 = d + s_lo.off;
  j->d_lenq = reinterpret_cast<double*>(d + s_lenq.off);
  j->d_order = reinterpret_cast<int32_t*>(d + s_order.off);
  j->d_col_exp = reinterpret_cast<int32_t*>(d + s_cexp.off);
  j->d_len_col = reinterpret_cast<double*>(d + s_lcol.off);
  j->d_chunks.end = reinterpret_cast<int32_t*>(d + s_cend.off);
  j->d_chunks.scale = reinterpret_cast<double*>(d + s_cscale.off);
  j->d_chunks.n = static_cast<int32_t>(chunk_end.size());
  j->d_chunks.shift = chunk_shift.empty() ? nullptr : reinterpret_cast<int32_t*>(d + s_cshift.off);
  j->d_need = need_blocks.empty() ? nullptr : reinterpret_cast<uint8_t*>(d + s_need.off);
  if (j->i8 && !chunk_scale.empty()) {
    const auto mm = std::minmax_element(chunk_scale.begin(), chunk_scale.end());
    j->d_chunks.biased = *mm.second <= *mm.first * 256.0;
  }
  j->d_lenf = reinterpret_cast<float*>(d_csr + s_lenf.off);
  j->d_tiles = reinterpret_cast<Tile*>(d + s_tiles.off);
  j->d_level_ptr = reinterpret_cast<int32_t*>(d_csr + s_lptr.off);
  j->dtree.level_parent = reinterpret_cast<int32_t*>(d_csr + s_lpar.off);

  mark("events + H2D enqueue");
  // ------------------------------------------------------------ device buffers
  const int chunks = weighted_scratch_chunks(B);
  if (j->exact || j->weighted) {
    // exact: the whole fp64 embedding; fast weighted: one slab of samples at a time, <= 2 GB
    const int64_t per_rank = j->np / (j->sharded ? world : 1);
    j->ws_slab = j->exact ? j->np
                          : std::min<int64_t>(per_rank, std::max<int64_t>(kTile, (2LL << 30) / (8LL * B) / kTile * kTile));
    if (const char* e = getenv("FRC_WS_SLAB"))  // tests: force several slabs on a small problem
      if (!j->exact && atoi(e) > 0) j->ws_slab = std::min<int64_t>(per_rank, round_up(atoi(e), kTile));
    if (!(j->d_E = dev_alloc<double>(j, static_cast<size_t>(B) * j->ws_slab, &rc))) return bail(rc);
    if (!(j->d_total = dev_alloc<double>(j, j->np, &rc))) return bail(rc);
    if (!j->exact) {
      if (!(j->d_A = dev_alloc<float>(j, static_cast<size_t>(j->kp) * j->np, &rc))) return bail(rc);
      if (!(j->d_W = dev_alloc<double>(j, j->np, &rc))) return bail(rc);
      if (!(j->d_scratch = dev_alloc<double>(j, static_cast<size_t>(chunks) * j->np, &rc))) return bail(rc);
      if (!(j->d_flag_counts = dev_alloc<unsigned long long>(j, j->mine.size() + 1, &rc))) return bail(rc);
      const size_t ws_words = static_cast<size_t>(c->num_sms) * B;
      if (!(j->d_fix_ws = dev_alloc<long long>(j, ws_words, &rc))) return bail(rc);
      CREATE_CUDA(cudaMemsetAsync(j->d_fix_ws, 0, sizeof(long long) * ws_words, c->stream[0]));
    }
  } else {
    {
      // fused: published columns bitsT[nw][kp] (+ node-indexed scratch for trees beyond shared memory);
      // FRC_EMBED_LEVELS=1: node-major bits[B][nw], one launch per tree level
      size_t words = j->fused_embed ? static_cast<size_t>(j->kp) * j->nw : static_cast<size_t>(B) * j->nw;
      const char* ep = getenv("FRC_PEER_PUSH");  // "0": exchange the bit columns with NCCL instead
      j->peer_push = j->sharded && j->fused_embed && !j->bits_feed && world <= 8 && !(ep && atoi(ep) == 0);
      if (j->peer_push) {
        // two buffers from the symmetric heap (same block + offset on every rank), mapped from the peers
        for (int k = 0; k < 2; ++k) {
          cudaError_t e = cudaSuccess;
          j->d_bits2[k] = static_cast<uint32_t*>(c->sym.alloc(words * sizeof(uint32_t), &e));
          if (!j->d_bits2[k]) return bail(fail(j, FRC_ERR_OOM, std::string("symmetric heap: ") + cudaGetErrorString(e)));
        }
        if ((rc = sym_sync(j)) != FRC_OK) return bail(rc);
        for (int k = 0; k < 2; ++k) {
          size_t blk = 0;
          for (; blk < c->sym.blocks.size(); ++blk) {
            char* base = c->sym.blocks[blk].p;
            if (reinterpret_cast<char*>(j->d_bits2[k]) >= base &&
                reinterpret_cast<char*>(j->d_bits2[k]) < base + c->sym.blocks[blk].cap) break;
          }
          const size_t off = reinterpret_cast<char*>(j->d_bits2[k]) - c->sym.blocks[blk].p;
          j->peers[k].n = world; j->peers[k].self = rank;
          for (int r = 0; r < world; ++r) j->peers[k].p[r] = reinterpret_cast<uint32_t*>(c->sym_peer[blk][r] + off);
        }
        j->d_bits = j->d_bits2[0];
      } else if (!(j->d_bits = dev_alloc<uint32_t>(j, std::max<size_t>(words, 64), &rc))) return bail(rc);
      const size_t ns = j->fused_embed ? static_cast<size_t>(presence_node_scratch_words(B, j->shard_nw)) : 0;
      if (ns && !(j->d_node_scratch = dev_alloc<uint32_t>(j, ns, &rc))) return bail(rc);
    }
    if (j->intacc && !j->bits_feed) {
      // capacity: when the three u8 operand arrays would not fit the free HBM (with room for the
      // output ring), the pair kernel expands its tiles from the bit rows instead (8x less memory,
      // about half the speed: DESIGN.md)
      const double need = 3.0 * static_cast<double>(j->np) * j->kp;
      if (need > 32.0 * (1 << 30)) {  // (cudaMemGetInfo costs milliseconds: only asked when it can matter)
        size_t free_b = 0, total_b = 0;
        CREATE_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (need > 0.80 * static_cast<double>(free_b)) j->bits_feed = true;
      }
      j->info.operand_kind = j->bits_feed ? 3 : 2;
    }
    const size_t opsz = j->bits_feed ? 256 : static_cast<size_t>(j->np) * j->kp * (j->i8 ? 1 : 2);
    if (!(j->d_P = dev_alloc<char>(j, opsz, &rc))) return bail(rc);
    if (!(j->d_Bh = dev_alloc<char>(j, opsz, &rc))) return bail(rc);
    if (!(j->d_Bl = dev_alloc<char>(j, opsz, &rc))) return bail(rc);
    if (j->bits_feed && !(j->d_bitsS = dev_alloc<uint32_t>(j, static_cast<size_t>(j->np) * (j->kp / 32), &rc)))
      return bail(rc);
    if (!(j->d_r = dev_alloc<double>(j, j->np, &rc))) return bail(rc);
    if (!(j->d_scratch = dev_alloc<double>(j, static_cast<size_t>(chunks) * j->np, &rc))) return bail(rc);
    if (!(j->d_flag_counts = dev_alloc<unsigned long long>(j, j->mine.size() + 1, &rc))) return bail(rc);
    std::string terr;
    if (j->i8 && !(j->d_flag_u = dev_alloc<double>(j, 1, &rc))) return bail(rc);
    if (j->i8 && !(j->d_qam = dev_alloc<uint32_t>(j, j->kp, &rc))) return bail(rc);
    if (j->intacc) {
      if (!(j->d_r_int = dev_alloc<long long>(j, j->np, &rc))) return bail(rc);
    }
    if (!j->bits_feed) {
      j->tc = tc_operands_create(j->d_P, j->d_Bh, j->d_Bl, j->np, j->kp, j->i8, j->d_chunks, j->d_len_col,
                                 j->d_flag_u, &terr);
      if (!j->tc) return bail(fail(j, FRC_ERR_CUDA, terr));
    } else {
      j->bo = bits_operands_create(j->d_bitsS, j->np, j->kp, static_cast<const uint8_t*>(j->d_q0),
                                   static_cast<const uint8_t*>(j->d_q1), static_cast<const uint8_t*>(j->d_q2),
                                   j->d_chunks, j->d_len_col, j->d_flag_u, j->d_r_int, std::ldexp(1.0, j->e_min), &terr);
      if (!j->bo) return bail(fail(j, FRC_ERR_CUDA, terr));
    }
    if (j->intacc && j->tc) tc_operands_set_int(j->tc, j->d_r_int, std::ldexp(1.0, j->e_min));
    if (j->i8) {  // a function of the tree only: once per job, not per restart
      j->info.kernel_launches += launch_quantize_lengths(
          j->d_len_col, j->d_col_exp, j->kp, static_cast<uint8_t*>(j->d_q0), static_cast<uint8_t*>(j->d_q1),
          static_cast<uint8_t*>(j->d_q2), j->d_qam, j->d_lenq, j->d_flag_u, c->stream[0]);
      CREATE_CUDA(cudaGetLastError());
    }
  }
  { const char* e = getenv("FRC_ZERO_COPY"); j->zero_copy = e && atoi(e) == 1; }
  {
    // fast-path distances have fp32 precision: send them over PCIe as fp32, widen on the host (wire.cu)
    // The widening is host CPU work the copy engine did for free with doubles on the bus: it pays when
    // this rank has the cores for it.  Measured: 16 threads 1.7 vs 2.8 ms per cfg2 call, 8 threads per
    // rank (2 GPUs) 9.9 vs 7.0e9 pairs/s, but 4 threads per rank (8 GPUs on a 32-core host) 8.1e9 against
    // 1.0e10 with doubles.  FRC_WIRE=f64 / f32 overrides.
    const char* e = getenv("FRC_WIRE");
    const bool want32 = e ? !strcmp(e, "f32") : c->pool->size() >= 8;
    j->wire32 = !j->exact && !(opts->flags & FRC_FLAG_NO_D2H) && !j->zero_copy && want32;
  }
  {
    int64_t want = std::max<int64_t>(kMinSlots, std::min<int64_t>(kSlots, kSlotBytesBudget / (max_band * 8)));
    j->n_slots = static_cast<int>(std::min<int64_t>(want, std::max<int64_t>(2, static_cast<int64_t>(j->mine.size()) + 1)));
  }
  if (j->mine.empty()) j->n_slots = 0;
  if (j->wire32 && j->n_slots > 0 && !(j->wide = pin_alloc<double>(j, std::min(max_band, kWidePiece), &rc))) return bail(rc);
  for (int k = 0; k < j->n_slots; ++k) {
    Slot& sl = j->slots[k];
    if (!(opts->flags & FRC_FLAG_NO_D2H) && !j->wire32 && !(sl.host = pin_alloc<double>(j, max_band, &rc))) return bail(rc);
    if (j->zero_copy && sl.host) sl.dev = sl.host;  // kernels store straight into pinned host memory (UVA)
    else if (!(sl.dev = dev_alloc<double>(j, max_band, &rc))) return bail(rc);
    if (j->wire32) {
      if (!(sl.dev32 = dev_alloc<float>(j, max_band, &rc))) return bail(rc);
      if (!(sl.host32 = pin_alloc<float>(j, max_band, &rc))) return bail(rc);
      if (!(sl.n_bad_host = pin_alloc<unsigned long long>(j, 1, &rc))) return bail(rc);
      *sl.n_bad_host = 0;
    }
    if (!j->exact) {
      if (!(sl.flagged = dev_alloc<uint32_t>(j, max_band, &rc))) return bail(rc);
      if (!(sl.n_flagged_host = pin_alloc<unsigned long long>(j, 1, &rc))) return bail(rc);
      *sl.n_flagged_host = 0;
    }
    cudaError_t e;
    if ((e = cudaEventCreate(&sl.k0)) != cudaSuccess || (e = cudaEventCreate(&sl.k1)) != cudaSuccess ||
        (e = cudaEventCreate(&sl.k2)) != cudaSuccess || (e = cudaEventCreate(&sl.k3)) != cudaSuccess ||
        (e = cudaEventCreate(&sl.done)) != cudaSuccess)
      return bail(fail(j, FRC_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e)));
  }

  mark("device buffers, tensor maps, slots");
  // -------------------------------------------- join the table workers, upload the CSR block
  csr_work(n_workers);  // whatever is left of the task list (everything, without workers)
  if (csr_workers_busy) { c->pool->wait(); csr_workers_busy = false; }
  if (trace) { fprintf(stderr, "[frc_create] CSR share times (us):"); for (double u : csr_us) fprintf(stderr, " %.0f", u); fprintf(stderr, "\n"); }
  if (bad_parent) return bail(fail(j, FRC_ERR_ARG, "parent[" + std::to_string(bad_parent) + "] is not a smaller node id"));
  if (bad_order) return bail(fail(j, FRC_ERR_ARG, "node ids are not a pre-order numbering (node " + std::to_string(bad_order) + ")"));
  for (auto& e : csr_errs)
    if (!e.empty()) return bail(fail(j, FRC_ERR_ARG, e));
  if (tree_H < 0 || !children_ok) return bail(fail(j, FRC_ERR_ARG, "tree topology could not be prepared"));  // (unreachable)
  j->dtree.height = tree_H;
  j->info.tree_height = tree_H;
  CREATE_CUDA(cudaMemcpyAsync(d_csr, stage_csr, csr_total, cudaMemcpyHostToDevice, c->stream[0]));
  CREATE_CUDA(cudaEventRecord(j->ev_h2d1, c->stream[0]));
  mark("CSR workers joined, CSR upload queued");
  // ----------------------------------------------------------------- go
  if ((rc = run_embedding(j)) != FRC_OK) return bail(rc);
  if ((rc = start_pairs(j)) != FRC_OK) return bail(rc);
  mark("enqueue embedding + first bands");
  *out = j;
  return FRC_OK;
#undef CREATE_CUDA
}

// fp32 wire: widens pairs [off, off + n) of the band in slot `sl` into j->wide (wire.cu).  A band larger
// than kWidePiece is handed out in pieces, so that the buffer always fits the cores' caches.
static int widen_piece(frc_job* j, Slot& sl, int64_t off, int64_t n, bool trace) {
  if (*sl.n_bad_host) {
    // some value of this band does not survive fp32 (underflow): fetch the doubles themselves, rounded like
    // the rest wherever fp32 can carry them
    cudaStream_t fs = j->ctx->stream[0];
    j->info.kernel_launches += launch_round_band(sl.dev + off, n, j->ctx->num_sms, fs);
    JOB_CUDA(j, cudaGetLastError());
    JOB_CUDA(j, cudaMemcpyAsync(j->wide, sl.dev + off, sizeof(double) * n, cudaMemcpyDeviceToHost, fs));
    JOB_CUDA(j, cudaStreamSynchronize(fs));
    j->info.d2h_bytes += static_cast<int64_t>(sizeof(double)) * n;
    return FRC_OK;
  }
  const auto t0 = std::chrono::steady_clock::now();
  Pool* pool = j->ctx->pool.get();
  int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(pool->size(), n / 65536)));
  if (const char* e = getenv("FRC_WIDEN_THREADS")) T = std::max(1, std::min(T, atoi(e)));
  const float* src = sl.host32 + off;
  double* dst = j->wide;
  double share_us[64] = {0};
  double* su = trace && T <= 64 ? share_us : nullptr;
  // chunks handed out through a counter (a worker on a busy core takes fewer); the first round is
  // static so that thread t starts where it started last time and finds its destination lines in its L2
  constexpr int64_t kChunk = 16384;
  const int64_t n_chunks = (n + kChunk - 1) / kChunk;
  std::atomic<int64_t> next{0};
  std::atomic<int64_t>* nx = &next;
  pool->run(T, [=](int t) {
    const auto w0 = std::chrono::steady_clock::now();
    for (int64_t ch = t < n_chunks ? t : n_chunks; ch < n_chunks;) {
      const int64_t lo = ch * kChunk, hi = std::min(n, lo + kChunk);
      widen_band(src + lo, dst + lo, hi - lo, /*stream_stores=*/false);
      ch = T + nx->fetch_add(1, std::memory_order_relaxed);
    }
    if (su) su[t] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
  });
  if (su) {
    fprintf(stderr, "[host] widen %lld pairs, shares (us):", static_cast<long long>(n));
    for (int t = 0; t < T; ++t) fprintf(stderr, " %.0f", su[t]);
    fprintf(stderr, "  (whole %.0f)\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  }
  return FRC_OK;
}

int frc_next(frc_job_t* j, const double** data, int64_t* first_index, int64_t* count) {
  if (!j || !data || !first_index || !count) return FRC_ERR_ARG;
  *data = nullptr; *first_index = 0; *count = 0;
  JOB_CUDA(j, cudaSetDevice(j->ctx->device));
  // a band larger than the widening buffer is handed out piece by piece
  if (j->held_slot >= 0 && j->wire32 && j->piece_off < j->piece_end) {
    const int64_t n = std::min(kWidePiece, j->piece_end - j->piece_off);
    int rc = widen_piece(j, j->slots[j->held_slot], j->piece_off, n, getenv("FRC_TRACE") != nullptr);
    if (rc) return rc;
    *data = j->wide;
    *first_index = j->piece_first + j->piece_off;
    *count = n;
    j->piece_off += n;
    return FRC_OK;
  }
  // the band handed out by the previous call is released now
  if (j->held_slot >= 0) {
    j->held_slot = -1;
    if (j->next_enqueue < j->mine.size()) {
      int rc = enqueue_band(j, j->next_enqueue);
      if (rc) return rc;
      ++j->next_enqueue;
    }
  }
  if (j->next_deliver >= j->mine.size()) return FRC_OK;  // end of stream
  if (j->next_deliver >= j->next_enqueue) {               // only when n_slots == 2 and first call
    int rc = enqueue_band(j, j->next_enqueue);
    if (rc) return rc;
    ++j->next_enqueue;
  }
  const size_t idx = j->next_deliver;
  Slot& sl = j->slots[idx % j->n_slots];
  const bool trace_next = getenv("FRC_TRACE") != nullptr;
  const auto tn0 = std::chrono::steady_clock::now();
  JOB_CUDA(j, cudaEventSynchronize(sl.done));
  const auto tn1 = std::chrono::steady_clock::now();
  float ms = 0.f;
  JOB_CUDA(j, cudaEventElapsedTime(&ms, sl.k0, sl.k1));
  j->info.pairs_ms += ms;
  JOB_CUDA(j, cudaEventElapsedTime(&ms, sl.k1, sl.k2));
  j->info.fixup_ms += ms;
  if (j->next_deliver + 1 == j->mine.size() && !j->run_timed) {
    JOB_CUDA(j, cudaEventSynchronize(j->ev_run1));
    JOB_CUDA(j, cudaEventElapsedTime(&ms, j->ev_embed0, j->ev_run1));
    j->info.run_ms = ms;
    j->run_timed = true;
  }
  if (sl.n_flagged_host) j->info.flagged_pairs += static_cast<int64_t>(*sl.n_flagged_host);
  if (trace_next) {
    fprintf(stderr, "[host] deliver band %zu at %.3f ms\n", idx,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - j->t_host0).count());
    float a = 0, b = 0, c2 = 0, d = 0;
    cudaEventElapsedTime(&a, j->ev_embed0, sl.k0);
    cudaEventElapsedTime(&b, j->ev_embed0, sl.k1);
    cudaEventElapsedTime(&c2, j->ev_embed0, sl.k2);
    cudaEventElapsedTime(&d, j->ev_embed0, sl.done);
    fprintf(stderr, "[band %3zu] start %.3f kernel_end %.3f fixup_end %.3f d2h_end %.3f ms (pairs %lld)\n", idx, a, b, c2, d,
            static_cast<long long>(j->bands[j->mine[idx]].count));
  }
  const Band& b = j->bands[j->mine[idx]];
  int64_t deliver_n = b.count;
  if (j->wire32) {
    deliver_n = std::min(kWidePiece, b.count);
    j->piece_first = b.first; j->piece_off = deliver_n; j->piece_end = b.count;
    int rc = widen_piece(j, sl, 0, deliver_n, trace_next);
    if (rc) return rc;
  }
  if (trace_next) {
    const auto tn2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[host] band %zu: waited %.1f us for the copy, bookkeeping + widening %.1f us\n", idx,
            std::chrono::duration<double, std::micro>(tn1 - tn0).count(),
            std::chrono::duration<double, std::micro>(tn2 - tn1).count());
  }
  *data = (j->opts.flags & FRC_FLAG_NO_D2H) ? sl.dev : j->wire32 ? j->wide : sl.host;
  *first_index = b.first;
  *count = deliver_n;
  j->held_slot = static_cast<int>(idx % j->n_slots);
  ++j->next_deliver;
  return FRC_OK;
}

int frc_restart(frc_job_t* j) {
  if (!j) return FRC_ERR_ARG;
  JOB_CUDA(j, cudaSetDevice(j->ctx->device));
  for (auto s : j->ctx->stream) JOB_CUDA(j, cudaStreamSynchronize(s));
  j->info.kernel_launches = 0;
  // stream 0 must not start before stream 1 drained (it did: both are idle)
  int rc = run_embedding(j);
  if (rc) return rc;
  return start_pairs(j);
}

int frc_job_info(const frc_job_t* cj, frc_info_t* info) {
  if (!cj || !info) return FRC_ERR_ARG;
  frc_job* j = const_cast<frc_job*>(cj);
  if (!j->embed_timed && j->ev_embed1) {
    cudaSetDevice(j->ctx->device);
    if (cudaEventSynchronize(j->ev_embed1) == cudaSuccess) {
      float a = 0.f, b = 0.f;
      if (cudaEventElapsedTime(&a, j->ev_h2d0, j->ev_h2d1) == cudaSuccess) j->info.h2d_ms = a;
      if (cudaEventElapsedTime(&b, j->ev_embed0, j->ev_embed1) == cudaSuccess) j->info.embed_ms = b;
      j->embed_timed = true;
    }
    cudaGetLastError();
  }
  *info = j->info;
  return FRC_OK;
}

void frc_destroy(frc_job_t* j) { destroy_job(j); }

int64_t frc_plan_bands(int64_t n_samples, int64_t band_rows, int32_t rank, int32_t world, uint32_t flags,
                       int64_t* first_index, int64_t* count, int64_t cap) {
  if (world <= 0) { world = 1; rank = 0; }
  if (n_samples < 0 || band_rows < 0 || rank < 0 || rank >= world) return -FRC_ERR_ARG;
  const std::vector<int64_t> brows = band_boundaries(n_samples, band_rows, world, !(flags & FRC_FLAG_NO_D2H));
  const std::vector<int> owners = band_owners(brows, world);
  int64_t n = 0;
  for (size_t kb = 0; kb + 1 < brows.size(); ++kb) {
    const int64_t r0 = brows[kb], r1 = brows[kb + 1];
    const int64_t first = r0 >= 2 ? tri(r0) : 0;
    const int64_t cnt = tri(r1) - first;
    if (cnt <= 0) continue;
    if (owners[kb] == rank) {
      if (n < cap && first_index && count) { first_index[n] = first; count[n] = cnt; }
      ++n;
    }
  }
  return n;
}

}  // extern "C"
