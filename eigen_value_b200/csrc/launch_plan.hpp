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

constexpr size_t kSmemLimit = 227 * 1024 - 1024; // opt-in maximum minus the kernel's static shared

struct ScConfig
{
  int id, max_threads, pf_batches;
};
// Resident-e configurations (st_options.kernel): {512 threads, PF_BATCHES x 4 KB prefetch slot per warp}.  The
// automatic choice is 13 (one batch: measured best or tied at N = 1024..32768); 10 / 12 size the slot to a whole
// row so that rows of N = 1025..3072 stay on chip for the whole solve; 11 (no slots) is the bf16-storage build.
// Everything else round 1 tried lost on hardware and was removed in round 2: 1024- and 256-thread CTAs with 8 / 16
// loads in flight (ids 14-19, profiles/r1_sweep_resident_e_variants.txt), per-warp TMA rings for the whole stream
// (ids 2-9, profiles/r1_sweep_tma_ring_vs_ldg.txt) and L2 prefetch hints across the barrier or one unit ahead
// (ids 21-26, profiles/r2_c1_sweep_l2_prefetch.txt: 1.5 % to 40 % slower at N = 8192 and 32768).
inline const ScConfig kScConfigs[] = {
  { 13, 512, 1 }, { 10, 512, 2 }, { 11, 512, 0 }, { 12, 512, 3 },
};
inline bool
is_sc_kernel_id(int id)
{
  return id >= 10 && id <= 13;
}
constexpr int kGeneralKernelId = 1;
constexpr int kWideKernelId = 2; // unit-scheduled, eigenvector staged one window at a time (automatic for N > 32768)
constexpr int kClusterKernelIdPlan = 20;
// st_options.kernel values the library accepts: 0 automatic, 1 general loop, 2 wide, 10-13 resident-e, 20 on-chip cluster
inline bool
is_known_kernel_id(int id)
{
  return id == 0 || id == kGeneralKernelId || id == kWideKernelId || is_sc_kernel_id(id) || id == kClusterKernelIdPlan;
}

// dynamic shared memory of the resident-e kernel: prefetch slots | e | mbarriers
inline size_t
sc_smem_bytes(int threads, int pf_batches, uint32_t cols, uint32_t* mbar_offset)
{
  size_t off = (size_t)(threads / 32) * pf_batches * 1024 * sizeof(float);
  off += (size_t)cols * sizeof(float);
  off = (off + 15) & ~(size_t)15;
  *mbar_offset = (uint32_t)off;
  return off + (size_t)(threads / 32) * sizeof(uint64_t);
}
// The shared-memory carve-out of an SM comes in steps (..., 164, 196, 228 KB; a CTA can use the step minus 1 KB) and
// what is left of the 256 KB is L1, which the streaming loads pass through.  Measured in round 2
// (profiles/r2_c5_endgame_v2_n1.json, r2_c6_endgame_v3_n1.json): Hilbert 32768 -- 64 KB of slots + 128 KB of e --
// runs a round in 587 us with up to 198.8 KB of shared memory per CTA and in 676 us with 199.8 KB.  Anything
// optional in shared memory has to stay under 195 KB.

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
