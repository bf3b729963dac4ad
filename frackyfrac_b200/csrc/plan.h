// Host-only planning of a job (no CUDA calls, no device memory): the pieces of frc_create that decide
// WHAT runs where.  Kept free of the runtime so that the CPU test-suite can drive them through the
// frc_debug_* entry points at the bottom of plan.cpp (tests/test_plan.py) without a GPU.
//
//   band plan      rows of the lower triangle (common/common.go:21-31) -> contiguous flat ranges, owners
//   column plan    fast unweighted: node -> operand column, u8 block-floating-point chunks
//   tile lists     per band: the 128 x 128 sample tiles (CTA-pair tiles for the tensor-core kernel)
//   tree walk      pre-order check, children lists, level order (what enumerateNodes' numbering implies,
//                  frcfrc/unifrac.go:127-133)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "frc_internal.h"

namespace frc {

struct Band {
  int64_t row0 = 0, row1 = 0;   // rows [row0, row1) of the triangle
  int64_t first = 0, count = 0; // flat index of the first pair, number of pairs
  int32_t tile_off = 0, n_tiles = 0;  // into the owner's tile list
  int owner = 0;                // rank (process) or device that computes it
};

// Row boundaries of the bands (multiples of the tile size, first 0, last N).
//   requested > 0 : uniform bands of that many rows (rounded up to whole tiles);
//   otherwise     : bands of (nearly) EQUAL PAIR COUNT, boundaries at N * sqrt(k / n): about
//                   `per_rank` bands per rank when streamed to the host (copies overlap kernels),
//                   one or two per rank when the distances stay in HBM; more when a band would exceed
//                   96 MB (streamed) / 1.5 GB (kept in HBM).
std::vector<int64_t> band_boundaries(int64_t N, int64_t requested, int world, bool d2h, int per_rank, int value_bytes);
// One band per owner in every round of G consecutive bands, greedily balanced (deterministic; the same in
// every process).
std::vector<int> band_owners(const std::vector<int64_t>& rows, int world);
// Every non-empty band of the triangle, in flat-index order, with its owner.
std::vector<Band> make_bands(int64_t N, int64_t requested, int world, bool d2h, int per_rank, int value_bytes);

// Capacity mode of the fast weighted path (PanelMap in frc_internal.h): the padded sample range [0, np) is cut
// into 2G shards of np / 2G rows; shard s belongs to device s (s < G) or 2G-1-s, which balances the pair counts
// (row r has r columns: the two shards of a device add up to the same total).  Bands never straddle a shard.
std::vector<Band> make_bands_capacity(int64_t N, int64_t np, int G, int64_t requested, bool d2h);

// Capacity mode, one device: its bands (in row order) grouped by row shard, and per group the tile list ordered by
// the COLUMN shard the tiles read (the order the shards visit in), column-major inside a column shard.
struct CapGroupPlan {
  int shard = 0;                 // row shard (also the last column shard its tiles need)
  int64_t first = 0, count = 0;  // flat range of the group's rows
  int32_t tile_off = 0;          // into `tiles`
  std::vector<int32_t> coff;     // [n_shards + 1] tile offset (inside the group's list) of every column shard
};
// `bands`: the device's bands; appends the groups' tiles to `tiles`; band_group[k] = group of bands[k].
std::vector<CapGroupPlan> plan_capacity_groups(const std::vector<Band>& bands, int64_t N, int64_t shard_rows, int n_shards,
                                               std::vector<Tile>& tiles, std::vector<int>& band_group);

// Fast unweighted path: which node sits in which column of the K-major operands and how the
// columns group into TMEM accumulation chunks.
struct ColumnPlan {
  bool i8 = false;       // u8 block floating point (kind::i8) / bf16 hi-lo planes (kind::f16)
  bool intacc = false;   // u8: all chunk scales within 2^16 -> 64-bit integer accumulation
  bool biased = false;   // u8 fp64 accumulation: all scales within 2^8 (single-instruction form)
  int32_t kp = 0;        // operand columns (multiple of 128 for u8, of 64 for bf16)
  int32_t e_min = 0;     // u8: smallest chunk exponent (integer unit = 2^e_min)
  std::vector<int32_t> col_order;  // [kp] node of column k, -1 = padding
  std::vector<int32_t> col_exp;    // [kp] u8: exponent of the column's chunk
  std::vector<double> len_col;     // [kp] true length of the column's node (0 in padding)
  std::vector<int32_t> chunk_end;  // K blocks (128 B of every operand row) where each chunk ends
  std::vector<double> chunk_scale;
  std::vector<int32_t> chunk_shift;  // u8: log2(scale / smallest scale)
};
ColumnPlan plan_columns(const double* length, int32_t n_nodes, bool want_i8, int group_binades, bool force_f64_acc,
                        int bf16_chunk_kblocks);

// Tiles of one band, appended to `tiles`: pair tiles (ti, tj) + (ti, tj + 1), tj even, walked in
// super-tiles of ~9 tile rows x ~8 tile-pair columns when `pair_tiles` (tensor-core kernel), plain
// column-major 128 x 128 tiles otherwise (FP32 tile kernel).
void append_band_tiles(Band& b, bool pair_tiles, std::vector<Tile>& tiles);
// Per block of 256 samples: bit 0 = some tile reads its A rows (column samples), bit 1 = its Bh / Bl rows.
std::vector<uint8_t> operand_need_blocks(const std::vector<Tile>& tiles, int64_t np, bool pair_tiles);

// The tree walk.  Outputs go to caller-provided arrays (pinned staging memory in frc_create).
struct TreeLevels {
  int32_t height = -1;             // levels above the leaves; -1 = failed
  int32_t bad_parent = 0, bad_order = 0;  // first offending node ids
  std::vector<int32_t> level_ptr;  // [height + 2]
};
// Pre-order check (parent[v] must be the node of its depth on the current root-to-(v-1) path) and the
// level order: level_nodes[B] grouped by height (leaves first), level_parent[k] = parent of level_nodes[k].
TreeLevels tree_levels(const int32_t* parent, int32_t n_nodes, int32_t* level_nodes, int32_t* level_parent);
// Children lists in ascending id (= file order).  false when a parent id is out of range.
bool tree_children(const int32_t* parent, int32_t n_nodes, int32_t* child_ptr, int32_t* child_idx);
// post_order[k] = node visited k-th by abundanceToFlatNodes' recursion (children first, in file order,
// then the node: unifrac.go:32-53).
void tree_post_order(const int32_t* parent, int32_t n_nodes, int32_t* post_order);

}  // namespace frc
