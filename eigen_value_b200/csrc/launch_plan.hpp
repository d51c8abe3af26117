// launch_plan.hpp -- host-side, CUDA-free arithmetic of the launch plan: kernel configuration
// tables and the dynamic shared-memory layouts the round kernels expect.  Shared by solver.cu and by
// the CPU emulation harness (tests/cuda_emu), so the emulated launches use the very same layouts.
#pragma once

#include <cstddef>
#include <cstdint>

namespace st {

constexpr int kPlanClusterMaxCtas = 8;                 // == kClusterMaxCtas (kernels_cluster.cuh)
constexpr size_t kPlanClusterSmemBudget = 200 * 1024;  // == kClusterSmemBudget

// Warp count for a CTA that owns `nrows` rows, one warp per row at a time: the count in
// [lo, hi] that leaves the fewest warps idle in the last sweep (e.g. 55 rows: 14 warps x 4
// sweeps = 56 slots instead of 16 x 4 = 64); ties go to the larger count.
inline int
balanced_warps(uint32_t nrows, int lo, int hi)
{
  int best = hi;
  double best_eff = -1.0;
  for (int w = hi; w >= lo; w--) {
    const uint32_t sweeps = (nrows + (uint32_t)w - 1u) / (uint32_t)w;
    const double eff = (double)nrows / ((double)sweeps * w);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = w;
    }
  }
  return best;
}

struct TmaConfig
{
  int id, threads, stages, tile_f;
};
// id 2 is the default; the others are tuning variants reachable through st_options.kernel
inline const TmaConfig kTmaConfigs[] = {
  { 2, 512, 3, 1024 }, { 3, 256, 6, 1024 }, { 4, 256, 3, 2048 }, { 5, 512, 2, 1024 },
  { 6, 256, 4, 1024 }, { 7, 1024, 1, 1024 }, { 8, 512, 1, 2048 }, { 9, 256, 2, 2048 },
};
constexpr size_t kSmemLimit = 227 * 1024 - 1024; // opt-in maximum minus the kernel's static shared

inline size_t
tma_smem_bytes(const TmaConfig& c, uint32_t chunk_cols, uint32_t rows_cap, uint32_t* mbar_offset)
{
  size_t off = (size_t)(c.threads / 32) * c.stages * c.tile_f * sizeof(float);
  off += (size_t)chunk_cols * sizeof(float);
  off += (size_t)rows_cap * sizeof(float);
  off = (off + 15) & ~(size_t)15;
  *mbar_offset = (uint32_t)off;
  return off + (size_t)(c.threads / 32) * c.stages * sizeof(uint64_t);
}

struct ScConfig
{
  int id, max_threads, pf_batches;
};
// the automatic choice takes the first entry that fits (one 4 KB batch prefetched per warp --
// measured best or tied at N = 1024..8192, profiles/r1_sweep_kernels_sc*.txt); the others are
// tuning variants reachable through st_options.kernel
inline const ScConfig kScConfigs[] = {
  { 13, 512, 1 }, { 10, 512, 2 }, { 11, 512, 0 }, { 12, 512, 3 }, { 14, 1024, 1 }, { 15, 1024, 0 },
  { 16, 256, 2 }, { 17, 256, 4 }, { 18, 256, 0 }, { 19, 256, 1 },
  // configuration 13 plus an L2 prefetch of 8 / 16 / 32 KB per warp across the round barrier (explicit only)
  { 21, 512, 1 }, { 22, 512, 1 }, { 23, 512, 1 },
  // ... and / or of the warp's next unit while the current one streams (static scheduling: st_options.sweep bit 1)
  { 24, 512, 1 }, { 25, 512, 1 }, { 26, 512, 1 },
};
inline bool
is_sc_kernel_id(int id)
{
  return (id >= 10 && id < 20) || (id >= 21 && id <= 26);
}

inline size_t
sc_smem_bytes(int threads, int pf_batches, uint32_t cols, uint32_t rows_cap, uint32_t* mbar_offset)
{
  size_t off = (size_t)(threads / 32) * pf_batches * 1024 * sizeof(float);
  off += (size_t)cols * sizeof(float);
  off += (size_t)rows_cap * sizeof(float);
  off = (off + 15) & ~(size_t)15;
  *mbar_offset = (uint32_t)off;
  return off + (size_t)(threads / 32) * sizeof(uint64_t);
}

// Cluster size (1, 2, 4, 8 CTAs): the largest that still leaves every CTA >= 16 rows (one per
// warp) -- the loop is latency-bound, so more SMs means shorter per-warp row chains -- and never
// smaller than what it takes to hold rows + e + both s buffers in each CTA's shared memory.
inline int
cluster_ctas_for(uint32_t dim, size_t* smem_bytes)
{
  int best = 0;
  for (int c = 1; c <= kPlanClusterMaxCtas; c *= 2) {
    const size_t rows_cap = (dim + (uint32_t)c - 1u) / (uint32_t)c;
    const size_t need = sizeof(float) * (rows_cap * dim + 3 * (size_t)dim);
    const bool fits = need <= kPlanClusterSmemBudget;
    const bool useful = c == 1 || dim / (uint32_t)c >= 16u;
    if (fits && (useful || !best)) {
      best = c;
      *smem_bytes = need;
    }
  }
  return best;
}

// Streamed solve (Context::solve_streamed): rows per block, number of blocks and cache slots for a
// dim x dim fp32 host matrix and a device budget in bytes.  block_rows == 0 asks for blocks of about
// `auto_block_bytes`, halved until two of them fit the budget.  Returns false when the budget holds
// fewer than two blocks of a matrix that has several (no double buffering possible).
struct StreamShape
{
  uint32_t block_rows, blocks, slots;
};
inline bool
stream_shape_for(uint32_t dim, size_t budget, uint32_t block_rows, size_t auto_block_bytes, StreamShape* out)
{
  const size_t row_bytes = sizeof(float) * (size_t)dim;
  uint32_t b = block_rows ? (block_rows < dim ? block_rows : dim) : (uint32_t)(auto_block_bytes / row_bytes);
  if (!block_rows) {
    b = b < 1u ? 1u : (b > dim ? dim : b);
    while (b > 1u && budget / (row_bytes * b) < 2)
      b = (b + 1u) / 2u;
  }
  const uint32_t nb = (dim + b - 1u) / b;
  const uint64_t fit = budget / (row_bytes * b);
  if (fit < 2 && nb > 1u)
    return false;
  out->block_rows = b;
  out->blocks = nb;
  out->slots = (uint32_t)(fit < 1 ? 1 : (fit < nb ? fit : nb));
  return true;
}

} // namespace st
