// kernels_sc.cuh -- round loop specialised for N <= 8192 columns ("single chunk").
//
// When the whole eigenvector fits in one shared-memory chunk (<= 32 KB) the loop gets much
// tighter than the general round_loop_kernel:
//
//   * e lives in shared memory for the WHOLE solve.  Every CTA updates its private copy in
//     place from the published row sums (e *= s/m, reference similarity_transform.cpp:260);
//     no global E buffers, no rebuild of the chunk per round.
//   * after the round barrier each thread needs ONE L2 round trip: it loads its <= 16 entries
//     of s (and their circular neighbours), the block reduces max / stop flag, and the same
//     registers update e.  The general kernel needs four dependent L2 trips per round.
//   * the matrix never changes, so before a warp enters the barrier its lane 0 issues ONE
//     TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) of the first
//     PF_BATCHES*4 KB of the row it will process first in the NEXT round into a private
//     shared-memory slot.  The L2->SM pipe keeps moving matrix bytes while the CTA sits in the
//     barrier and the vector tail; the first batches of the next round are then consumed
//     from shared memory.  A prefetch issued for a round that never runs is drained at exit.
//
// Row sums are bit-identical to round_loop_kernel: same lane / accumulator / fold order.
// Read-only form, N % 4 == 0, N <= kChunkCols only.
#pragma once

#include "kernels.cuh"
#include "kernels_tma.cuh"

namespace st {

// one row: the leading `npre` float4 come from the prefetched shared-memory tile, the rest
// from global memory.  npre is a multiple of 256 (one batch = 32 lanes x 8 accumulators) or nv.
// LD = independent 128-bit loads in flight per lane (8 or 16); the accumulator a vector goes to
// is always (index / 32) % 8, so the evaluation order does not depend on LD.
template<int LD>
__device__ __forceinline__ float
row_dot_prefetched(const float4* __restrict__ a, const float4* es, uint32_t nv, int lane,
                   const float4* pf, uint32_t npre)
{
  static_assert(LD % kUnroll == 0, "loads in flight must be a multiple of the accumulator count");
  float acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = 0.f;
  for (uint32_t b0 = 0; b0 < npre; b0 += 32u * kUnroll) {
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      const uint32_t j = b0 + lane + 32u * u;
      if (j < npre)
        acc[u] = dot_acc(pf[j], es[j], acc[u]);
    }
  }
  uint32_t i = npre + (uint32_t)lane;
  for (; i + 32u * (LD - 1) < nv; i += 32u * LD) {
    float4 v[LD];
#pragma unroll
    for (int u = 0; u < LD; u++)
      v[u] = ld_stream(a + i + 32u * u);
#pragma unroll
    for (int u = 0; u < LD; u++)
      acc[u % kUnroll] = dot_acc(v[u], es[i + 32u * u], acc[u % kUnroll]);
  }
#pragma unroll
  for (int u = 0; u < LD; u++) {
    const uint32_t j = i + 32u * u;
    if (j < nv)
      acc[u % kUnroll] = dot_acc(ld_stream(a + j), es[j], acc[u % kUnroll]);
  }
#pragma unroll
  for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
    for (int u = 0; u < s; u++)
      acc[u] += acc[u + s];
  return warp_sum(acc[0]);
}

template<int THREADS, int PF_BATCHES, int LD = kUnroll>
__global__ void __launch_bounds__(THREADS, 1) round_loop_sc_kernel(const RoundParams p)
{
  constexpr int kWarps = THREADS / 32;
  constexpr int kPerThread = kChunkCols / THREADS;   // entries of s each thread reduces
  constexpr uint32_t kPfFloats = PF_BATCHES * 1024u; // prefetch slot per warp

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* pf_all = reinterpret_cast<float*>(smem_raw);        // kWarps x kPfFloats
  float* e_s = pf_all + (size_t)kWarps * kPfFloats;          // N floats, lives across rounds
  float* part_s = e_s + p.chunk_cols;                        // one row sum per owned row
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(smem_raw + p.mbar_offset);
  __shared__ float red_max[32];
  __shared__ int red_ok[32];
  __shared__ float bc_max;
  __shared__ int bc_ok;
  __shared__ int s_abort;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t N = p.N;
  const uint32_t nv = N >> 2;

  const uint32_t rb = (uint32_t)((uint64_t)p.rows * blockIdx.x / gridDim.x);
  const uint32_t re = (uint32_t)((uint64_t)p.rows * (blockIdx.x + 1) / gridDim.x);
  const uint32_t nrows = re - rb;
  const uint32_t cb = (uint32_t)((uint64_t)N * blockIdx.x / gridDim.x);
  const uint32_t ce = (uint32_t)((uint64_t)N * (blockIdx.x + 1) / gridDim.x);
  const uint32_t my_rows = nrows > (uint32_t)warp ? (nrows - warp + kWarps - 1) / kWarps : 0u;

  float* my_pf = pf_all + (size_t)warp * kPfFloats;
  uint64_t* my_bar = mbar_all + warp;
  const uint32_t pf_floats = PF_BATCHES > 0 ? min(kPfFloats, N) : 0u;
  if (PF_BATCHES > 0 && lane == 0)
    mbar_init(my_bar, 1u);
  for (uint32_t c = tid; c < N; c += THREADS)
    e_s[c] = 1.f; // initialise_eigen_vector, reference :267-284
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (blockIdx.x == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();

  uint32_t pf_issued = 0, pf_consumed = 0; // bulk copies issued / waited for by this warp

  for (uint32_t k = 0;; ++k) {
    float* Scur = p.S[k & 1];
    const bool backward = p.sweep && (k & 1);

    // ---- the pass over the matrix ----                                   reference :40 (+ :52)
    bool tma_ok = true;
    for (uint32_t ii = 0; ii < my_rows; ii++) {
      const uint32_t i = warp + ii * kWarps;
      const uint32_t rl = backward ? (nrows - 1u - i) : i;
      const float4* row = reinterpret_cast<const float4*>(p.A + (size_t)(rb + rl) * N);
      uint32_t npre = 0;
      if (PF_BATCHES > 0 && ii == 0 && pf_consumed < pf_issued) {
        tma_ok = mbar_wait(my_bar, pf_consumed & 1u, p.timeout_ns);
        pf_consumed++;
        npre = pf_floats >> 2;
      }
      const float t = row_dot_prefetched<LD>(row, reinterpret_cast<const float4*>(e_s), nv, lane,
                                         reinterpret_cast<const float4*>(my_pf), npre);
      if (lane == 0)
        part_s[rl] = t;
    }
    // keep the L2->SM pipe busy across the barrier: fetch the head of next round's first row
    if (PF_BATCHES > 0 && my_rows > 0u && k + 1u < p.max_iter) {
      __syncwarp();
      if (lane == 0) {
        const uint32_t rl = (p.sweep && ((k + 1u) & 1u)) ? (nrows - 1u - warp) : (uint32_t)warp;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(my_bar, pf_floats * 4u);
        bulk_load(my_pf, p.A + (size_t)(rb + rl) * N, pf_floats * 4u, my_bar);
      }
      pf_issued++;
    }
    if (!tma_ok && lane == 0)
      atomicExch(&p.bar->error, 2u);
    __syncthreads();

    // ---- publish: s[r] = (A.e)[r] / e[r], to every rank when sharded ----
    for (uint32_t r = tid; r < nrows; r += THREADS) {
      const uint32_t gr = p.row0 + rb + r;
      const float s = part_s[r] / e_s[gr];
      if (p.world > 1) {
        for (uint32_t g = 0; g < p.world; g++)
          __stcg(p.peer_S[k & 1][g] + gr, s);
      } else {
        __stcg(Scur + gr, s);
      }
    }

    if (!round_barrier(p, k, &s_abort))
      break;

    // ---- every CTA: one L2 trip for s, then max / circular stop / e update ----  :41-44
    float sv[kPerThread];
    float mx = 0.f; // reference zero-fills the max cell (:169)
    int ok = 1;
#pragma unroll
    for (int j = 0; j < kPerThread; j++) {
      const uint32_t c = tid + j * THREADS;
      sv[j] = 0.f;
      if (c < N) {
        const float self = ld_cg(Scur + c);
        const float next = ld_cg(Scur + (c + 1u == N ? 0u : c + 1u));
        sv[j] = self;
        mx = fmaxf(mx, self);
        ok &= (fabsf(self - next) < p.eps) ? 1 : 0; // strict <, wrap pair included (:413-421)
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      ok &= __shfl_xor_sync(0xffffffffu, ok, o);
    }
    if (lane == 0) {
      red_max[warp] = mx;
      red_ok[warp] = ok;
    }
    __syncthreads();
    if (warp == 0) {
      mx = lane < kWarps ? red_max[lane] : 0.f;
      ok = lane < kWarps ? red_ok[lane] : 1;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        ok &= __shfl_xor_sync(0xffffffffu, ok, o);
      }
      if (lane == 0) {
        bc_max = mx;
        bc_ok = ok;
      }
    }
    __syncthreads();
    const float m_k = bc_max;
    const bool converged = bc_ok != 0;

    // e_{k+1} = e_k * (s_k / m_k), in place in shared memory                       :260
#pragma unroll
    for (int j = 0; j < kPerThread; j++) {
      const uint32_t c = tid + j * THREADS;
      if (c < N)
        e_s[c] = e_s[c] * (sv[j] / m_k);
    }
    if (blockIdx.x == 0 && tid == 0)
      p.round_ts[k + 1] = globaltimer_ns();
    __syncthreads();

    if (converged || k + 1u == p.max_iter) {
      // the eigenvector update of this round still happens before the break (:42-50)
      for (uint32_t c = cb + tid; c < ce; c += THREADS)
        p.out_eigen_vec[c] = e_s[c];
      if (blockIdx.x == 0 && tid == 0) {
        *p.out_eigen_val = ld_cg(Scur);             // :60-65
        p.out_iter[0] = converged ? k : p.max_iter; // :54
        p.out_iter[1] = k + 1u;
      }
      break;
    }
  }
  // a prefetch issued for a round that did not run must land before the CTA exits
  if (PF_BATCHES > 0 && lane == 0)
    for (; pf_consumed < pf_issued; pf_consumed++)
      mbar_wait(my_bar, pf_consumed & 1u, p.timeout_ns);
}

} // namespace st
