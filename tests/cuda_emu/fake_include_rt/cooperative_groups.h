// stand-in for <cooperative_groups.h>: the thread-block-cluster subset round_loop_cluster_kernel uses.
// The emulated "cluster" is the whole grid of one launch.
#pragma once
#include "../cuda_emu.h"

namespace cooperative_groups {

struct cluster_group
{
  unsigned num_blocks() const { return gridDim.x; }
  unsigned block_rank() const { return blockIdx.x; }
  void sync() const
  {
    emu::Cta* c = emu::g_cta;
    emu::Grid* g = c->grid;
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned gen = g->cl_gen.load(std::memory_order_acquire);
      if (g->cl_count.fetch_add(1, std::memory_order_acq_rel) + 1 == g->ctas) {
        g->cl_count.store(0, std::memory_order_relaxed);
        g->cl_gen.store(gen + 1, std::memory_order_release);
      } else {
        while (g->cl_gen.load(std::memory_order_acquire) == gen)
          sched_yield();
      }
    }
    __syncthreads();
  }
  // address of the same dynamic-shared-memory object in CTA `rank` (distributed shared memory)
  template<typename T>
  T* map_shared_rank(T* p, unsigned rank) const
  {
    emu::Cta* c = emu::g_cta;
    const ptrdiff_t off = reinterpret_cast<unsigned char*>(p) - c->smem;
    return reinterpret_cast<T*>(c->grid->cta[rank]->smem + off);
  }
};

inline cluster_group
this_cluster()
{
  return cluster_group{};
}

} // namespace cooperative_groups
