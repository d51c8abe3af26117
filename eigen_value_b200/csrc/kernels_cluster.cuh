// kernels_cluster.cuh -- round loop for matrices that fit ON CHIP (N <= 512, one GPU).
//
// Below N ~ 1024 a round of the grid-wide kernels costs ~4 us whatever the matrix size: it is six
// dependent L2 round trips (publish, barrier arrive + poll, row-sum re-read, first matrix loads).
// For N <= 512 the matrix (<= 1 MiB) fits in the shared memory of ONE thread-block cluster, so
// this kernel removes every one of those trips:
//
//   * a single cluster of C = 1, 2, 4 or 8 CTAs; CTA q loads rows [N q / C, N (q+1) / C) of A into
//     its shared memory ONCE and keeps them there for the whole solve;
//   * every CTA holds the full eigenvector e and both parity buffers of the row-sum vector s in
//     its own shared memory; the warp that finishes a row stores s[r] into EVERY CTA's copy
//     through distributed shared memory (cluster.map_shared_rank), so after ONE hardware
//     cluster barrier per round (barrier.cluster, ~0.2 us) each CTA reduces max / stop flag and
//     updates e from its own shared memory.  No global memory traffic inside the loop at all.
//
// Same evaluation order as every other round kernel (float4 j -> lane j % 32, accumulator
// (j / 32) % 8, pairwise fold, xor-shuffle tree), hence the same bits.
// Read-only form, N % 4 == 0, N <= kClusterCols, single GPU.
#pragma once

#include <cooperative_groups.h>

#include "kernels.cuh"

namespace st {

constexpr int kClusterCols = 512;        // largest N kept on chip (8 CTAs x 128 KB)
constexpr int kClusterMaxCtas = 8;       // portable cluster size
constexpr size_t kClusterSmemBudget = 200 * 1024;
constexpr int kRowsAtOnce = 4;           // independent rows a warp keeps in flight

template<int THREADS, int STOP = kStopAbsolute>
__global__ void __launch_bounds__(THREADS, 1) round_loop_cluster_kernel(const RoundParams p)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t C = cluster.num_blocks();
  const uint32_t q = cluster.block_rank();
  constexpr uint32_t kWarps = THREADS / 32;

  const uint32_t N = p.N;
  const uint32_t nv = N >> 2;
  const uint32_t rb = (uint32_t)((uint64_t)N * q / C);
  const uint32_t re = (uint32_t)((uint64_t)N * (q + 1) / C);
  const uint32_t nrows = re - rb;
  const uint32_t rows_cap = (N + C - 1u) / C;

  extern __shared__ __align__(16) float smem[];
  float* A_s = smem;                          // rows_cap x N: this CTA's rows of the matrix
  float* e_s = A_s + (size_t)rows_cap * N;    // N: eigenvector
  float* s_s = e_s + N;                       // 2 x N: row sums, by round parity
  __shared__ float red_max[32];
  __shared__ int red_ok[32];
  __shared__ float bc_max;
  __shared__ int bc_ok;

  const uint32_t tid = threadIdx.x;
  const int lane = (int)(tid & 31u);
  const uint32_t warp = tid >> 5;

  // ---- one-time load of this CTA's rows (the only matrix traffic of the whole solve) ----
  {
    const float4* src = reinterpret_cast<const float4*>(p.A) + (size_t)rb * nv;
    float4* dst = reinterpret_cast<float4*>(A_s);
    const uint32_t total = nrows * nv;
    for (uint32_t i = tid; i < total; i += THREADS)
      dst[i] = ld_stream(src + i);
  }
  for (uint32_t c = tid; c < N; c += THREADS)
    e_s[c] = 1.f; // initialise_eigen_vector, reference :267-284
  if (q == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();
  cluster.sync(); // every CTA of the cluster is running: remote stores may start

  // the other CTAs' copies of s, through distributed shared memory
  float* peer_s[kClusterMaxCtas];
#pragma unroll
  for (uint32_t g = 0; g < (uint32_t)kClusterMaxCtas; g++)
    peer_s[g] = g < C ? cluster.map_shared_rank(s_s, g) : nullptr;

  for (uint32_t k = 0;; ++k) {
    float* Scur = s_s + (size_t)(k & 1u) * N;
    // ---- row pass out of shared memory ----                              reference :40 (+ :52)
    // kRowsAtOnce rows per warp in flight: a row is one long dependent chain (LDS -> 4 FMA ->
    // fold -> 5 shuffles -> divide), so independent rows are interleaved to hide its latency
    const float4* e4 = reinterpret_cast<const float4*>(e_s);
    for (uint32_t i0 = warp; i0 < nrows; i0 += kRowsAtOnce * kWarps) {
      float acc[kRowsAtOnce][kUnroll];
#pragma unroll
      for (int r = 0; r < kRowsAtOnce; r++)
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
          acc[r][u] = 0.f;
#pragma unroll
      for (int u = 0; u < kClusterCols / 128; u++) { // <= 4 vectors per lane and row
        const uint32_t j = lane + 32u * u;
        if (j < nv) {
          const float4 ev = e4[j];
#pragma unroll
          for (int r = 0; r < kRowsAtOnce; r++) {
            const uint32_t i = i0 + (uint32_t)r * kWarps;
            if (i < nrows)
              acc[r][u] = dot_acc(reinterpret_cast<const float4*>(A_s + (size_t)i * N)[j], ev, acc[r][u]);
          }
        }
      }
      float t[kRowsAtOnce];
#pragma unroll
      for (int r = 0; r < kRowsAtOnce; r++) {
#pragma unroll
        for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
          for (int u = 0; u < s; u++)
            acc[r][u] += acc[r][u + s];
        t[r] = acc[r][0];
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int r = 0; r < kRowsAtOnce; r++)
          t[r] += __shfl_xor_sync(0xffffffffu, t[r], o);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kRowsAtOnce; r++) {
          const uint32_t i = i0 + (uint32_t)r * kWarps;
          if (i < nrows) {
            const uint32_t row = rb + i;
            const float sval = t[r] / e_s[row]; // s[r] = (A.e)[r] / e[r]
#pragma unroll
            for (uint32_t g = 0; g < (uint32_t)kClusterMaxCtas; g++)
              if (g < C)
                peer_s[g][(size_t)(k & 1u) * N + row] = sval;
          }
        }
      }
    }
    if (q == 0 && tid == 0)
      stamp_phase(p, k, 0u);
    cluster.sync(); // the one barrier of the round: every CTA's copy of s_k is complete
    if (q == 0 && tid == 0)
      stamp_phase(p, k, 1u);

    // ---- max, circular stop test, eigenvector update: all from shared memory ----  :41-44
    float mx = 0.f; // reference zero-fills the max cell (:169)
    int ok = STOP == kStopRelative ? 0 : 1;
    for (uint32_t c = tid; c < N; c += THREADS) {
      const float self = Scur[c];
      const float next = Scur[c + 1u == N ? 0u : c + 1u];
      mx = fmaxf(mx, self);
      if (STOP == kStopRelative) // `ok` carries the bits of the largest adjacent difference (see diff_bits)
        ok = (int)max((uint32_t)ok, diff_bits(self, next));
      else
        ok &= (fabsf(self - next) < p.eps) ? 1 : 0; // strict <, wrap pair included (:413-421)
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (STOP == kStopRelative)
        ok = (int)max((uint32_t)ok, (uint32_t)__shfl_xor_sync(0xffffffffu, ok, o));
      else
        ok &= __shfl_xor_sync(0xffffffffu, ok, o);
    }
    if (lane == 0) {
      red_max[warp] = mx;
      red_ok[warp] = ok;
    }
    __syncthreads();
    if (warp == 0) {
      mx = (uint32_t)lane < kWarps ? red_max[lane] : 0.f;
      ok = (uint32_t)lane < kWarps ? red_ok[lane] : (STOP == kStopRelative ? 0 : 1);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (STOP == kStopRelative)
          ok = (int)max((uint32_t)ok, (uint32_t)__shfl_xor_sync(0xffffffffu, ok, o));
        else
          ok &= __shfl_xor_sync(0xffffffffu, ok, o);
      }
      if (lane == 0) {
        bc_max = mx;
        bc_ok = STOP == kStopRelative ? (__uint_as_float((uint32_t)ok) < p.eps * mx ? 1 : 0) : ok;
      }
    }
    __syncthreads();
    const float m_k = bc_max;
    const bool converged = bc_ok != 0;
    for (uint32_t c = tid; c < N; c += THREADS)
      e_s[c] = e_s[c] * (Scur[c] / m_k); // e *= s / m                                      :260
    __syncthreads();
    if (q == 0 && tid == 0)
      stamp_round_end(p, k);

    if (converged || k + 1u == p.max_iter) {
      if (q == 0) {
        for (uint32_t c = tid; c < N; c += THREADS)
          p.out_eigen_vec[c] = e_s[c];
        if (tid == 0) {
          *p.out_eigen_val = Scur[0];                 // :60-65
          p.out_iter[0] = converged ? k : p.max_iter; // :54
          p.out_iter[1] = k + 1u;
        }
      }
      break;
    }
  }
  // no CTA may leave while a sibling could still address its shared memory
  cluster.sync();
}

} // namespace st
