// kernels_tma.cuh -- the round loop with the matrix streamed by the TMA engine.
//
// Same algorithm, same barrier, same evaluation order (bit-identical row sums) as
// round_loop_kernel in kernels.cuh; what changes is how the matrix reaches the SM:
//
//   * every warp owns a private ring of STAGES shared-memory tiles (TILE_F floats each) and
//     fills it with 1-D bulk copies  cp.async.bulk.shared::cluster.global.mbarrier::complete_tx
//     (SASS: UBLKCP), one mbarrier per stage, issued by lane 0 right after the warp has
//     consumed the tile that occupied the stage.  No cross-warp synchronisation is needed for
//     the ring, and 16 warps x 3 stages x 4 KB = 192 KB per SM are in flight (vs 64 KB with
//     register-staged 128-bit loads).
//   * the matrix never changes, so the issue side simply runs ahead of the consume side in the
//     (round, chunk, row, segment) order -- ACROSS the round barrier.  While a CTA sits in the
//     grid barrier and the vector tail, its ring is already being filled with the first tiles of
//     the next round, so HBM does not idle between rounds.  Tiles fetched for a round that never
//     happens (convergence) are drained before the CTA exits.
//
// Read-only form, N % 4 == 0 only; other cases use round_loop_kernel.
#pragma once

#include "kernels.cuh"

namespace st {





// Bounded wait: a wrong byte count would otherwise hang the GPU.
__device__ __forceinline__ bool
mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long timeout_ns)
{
  if (mbar_try_wait(bar, parity))
    return true;
  const unsigned long long t0 = globaltimer_ns();
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u && globaltimer_ns() - t0 > timeout_ns)
      return false;
  }
  return true;
}



// Position of one warp in its tile sequence: round k, column chunk c0, ordinal ii of the row
// among the warp's rows, column offset t0 inside the chunk.
struct TileCursor
{
  uint32_t k, c0, ii, t0;
};

template<int THREADS, int STAGES, int TILE_F>
__global__ void __launch_bounds__(THREADS, 1) round_loop_tma_kernel(const RoundParams p)
{
  static_assert(TILE_F % 1024 == 0, "tile must keep the lane/accumulator order of row_dot_readonly");
  constexpr int kWarps = THREADS / 32;
  constexpr int kVecPerLane = TILE_F / 128; // float4 per lane per tile

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* scale_s = ring + (size_t)kWarps * STAGES * TILE_F;
  float* part_s = scale_s + p.chunk_cols;
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
  const uint32_t chunk = p.chunk_cols;

  const uint32_t rb = (uint32_t)((uint64_t)p.rows * blockIdx.x / gridDim.x);
  const uint32_t re = (uint32_t)((uint64_t)p.rows * (blockIdx.x + 1) / gridDim.x);
  const uint32_t nrows = re - rb;
  const uint32_t cb = (uint32_t)((uint64_t)N * blockIdx.x / gridDim.x);
  const uint32_t ce = (uint32_t)((uint64_t)N * (blockIdx.x + 1) / gridDim.x);
  // rows of this CTA handled by this warp: warp, warp + kWarps, ...
  const uint32_t my_rows = nrows > (uint32_t)warp ? (nrows - warp + kWarps - 1) / kWarps : 0u;

  float* my_ring = ring + (size_t)warp * STAGES * TILE_F;
  uint64_t* my_bar = mbar_all + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; s++)
      mbar_init(my_bar + s, 1u);
  fence_mbarrier_init();
  fence_proxy_async();
  __syncthreads();

  if (blockIdx.x == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();
  const unsigned long long pol_keep = l2_policy_evict_last();
  const unsigned long long pol_stream = l2_policy_evict_first();

  // ---- issue side (lane 0 of every warp) ----
  TileCursor pc{ 0u, 0u, 0u, 0u };
  uint32_t issued = 0, consumed = 0;
  auto issue_next = [&]() {
    // never fetch for a round that cannot run
    if (my_rows == 0u || pc.k >= p.max_iter)
      return;
    const uint32_t clen = min(chunk, N - pc.c0);
    const uint32_t len = min((uint32_t)TILE_F, clen - pc.t0);
    const uint32_t i = warp + pc.ii * kWarps;
    const uint32_t rl = (p.sweep && (pc.k & 1u)) ? (nrows - 1u - i) : i;
    const float* src = p.A + (size_t)(rb + rl) * N + pc.c0 + pc.t0;
    const uint32_t stage = issued % STAGES;
    float* dst = my_ring + (size_t)stage * TILE_F;
    fence_proxy_async();
    mbar_arrive_expect_tx(my_bar + stage, len * 4u);
    if (p.keep_rows_pct == 0u)
      bulk_load(dst, src, len * 4u, my_bar + stage);
    else
      bulk_load_hint(dst, src, len * 4u, my_bar + stage,
                     rl * 100u < nrows * p.keep_rows_pct ? pol_keep : pol_stream);
    issued++;
    // advance the cursor: segment -> row -> chunk -> round
    pc.t0 += TILE_F;
    if (pc.t0 >= clen) {
      pc.t0 = 0u;
      if (++pc.ii >= my_rows) {
        pc.ii = 0u;
        pc.c0 += chunk;
        if (pc.c0 >= N) {
          pc.c0 = 0u;
          pc.k++;
        }
      }
    }
  };
  if (lane == 0)
    for (int s = 0; s < STAGES; s++)
      issue_next();

  auto drain = [&]() {
    // tiles fetched ahead of a round that will not run must land before the CTA exits
    if (lane == 0)
      for (; consumed < issued; consumed++)
        mbar_wait(my_bar + consumed % STAGES, (consumed / STAGES) & 1u, p.timeout_ns);
    __syncwarp();
  };

  float m_prev = 1.f;
  for (uint32_t k = 0;; ++k) {
    const float* Sprev = p.S[(k + 1) & 1];
    const float* Eprev = p.E[(k + 1) & 1];
    float* Scur = p.S[k & 1];
    float* Ecur = p.E[k & 1];
    const bool first = (k == 0);

    for (uint32_t c = cb + tid; c < ce; c += THREADS)
      Ecur[c] = first ? 1.f : ld_cg(Eprev + c) * (ld_cg(Sprev + c) / m_prev);
    for (uint32_t r = tid; r < nrows; r += THREADS)
      part_s[r] = 0.f;

    const bool backward = p.sweep && (k & 1);
    bool tma_ok = true;
    for (uint32_t c0 = 0; c0 < N; c0 += chunk) {
      const uint32_t clen = min(chunk, N - c0);
      __syncthreads();
      const uint32_t rot = cb % clen; // de-phase the CTAs on the L2 lines of E and S
      for (uint32_t cc = tid; cc < clen; cc += THREADS) {
        const uint32_t c = cc + rot < clen ? cc + rot : cc + rot - clen;
        scale_s[c] = first ? 1.f : ld_cg(Eprev + c0 + c) * (ld_cg(Sprev + c0 + c) / m_prev);
      }
      __syncthreads();
      for (uint32_t ii = 0; ii < my_rows; ii++) {
        const uint32_t i = warp + ii * kWarps;
        const uint32_t rl = backward ? (nrows - 1u - i) : i;
        float acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
          acc[u] = 0.f;
        for (uint32_t t0 = 0; t0 < clen; t0 += TILE_F) {
          const uint32_t nv = min((uint32_t)TILE_F, clen - t0) >> 2;
          const uint32_t stage = consumed % STAGES;
          tma_ok = mbar_wait(my_bar + stage, (consumed / STAGES) & 1u, p.timeout_ns) && tma_ok;
          const float4* a4 = reinterpret_cast<const float4*>(my_ring + (size_t)stage * TILE_F);
          const float4* e4 = reinterpret_cast<const float4*>(scale_s + t0);
          if (nv == (uint32_t)(TILE_F / 4)) {
#pragma unroll
            for (int u = 0; u < kVecPerLane; u++)
              acc[u % kUnroll] = dot_acc(a4[lane + 32 * u], e4[lane + 32 * u], acc[u % kUnroll]);
          } else {
#pragma unroll
            for (int u = 0; u < kVecPerLane; u++)
              if ((uint32_t)(lane + 32 * u) < nv)
                acc[u % kUnroll] = dot_acc(a4[lane + 32 * u], e4[lane + 32 * u], acc[u % kUnroll]);
          }
          consumed++;
          __syncwarp();
          if (lane == 0)
            issue_next();
        }
#pragma unroll
        for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
          for (int u = 0; u < s; u++)
            acc[u] += acc[u + s];
        const float t = warp_sum(acc[0]);
        if (lane == 0)
          part_s[rl] += t;
      }
    }
    if (!tma_ok && lane == 0)
      atomicExch(&p.bar->error, 2u);
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 0u);
    __syncthreads();

    for (uint32_t r = tid; r < nrows; r += THREADS) {
      const uint32_t gr = p.row0 + rb + r;
      float s = part_s[r];
      if (!first)
        s = s / (ld_cg(Eprev + gr) * (ld_cg(Sprev + gr) / m_prev));
      if (p.world > 1) {
        for (uint32_t g = 0; g < p.world; g++)
          __stcg(p.peer_S[k & 1][g] + gr, s);
      } else {
        __stcg(Scur + gr, s);
      }
    }

    if (!round_barrier(p, k, &s_abort)) {
      drain();
      return;
    }
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 1u);

    float mx = 0.f;
    int ok = 1;
    // every CTA scans the same vector at the same time: start each at its own offset (cb) so
    // they do not queue on the same L2 lines; max / AND are order-independent.  The circular
    // neighbour comes from the next lane by shuffle, as in the reference (:413-417); four
    // batches of loads are in flight before anything depends on them.
    for (uint32_t b0 = 0; b0 < N; b0 += 4u * THREADS) {
      float sf[4], nx[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t c0 = b0 + (uint32_t)j * THREADS + tid;
        const bool active = c0 < N;
        const uint32_t c = c0 + cb < N ? c0 + cb : c0 + cb - N;
        sf[j] = active ? ld_cg(Scur + c) : 0.f;
        const bool edge = active && (lane == 31 || c0 + 1u >= N);
        nx[j] = edge ? ld_cg(Scur + (c + 1u == N ? 0u : c + 1u)) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t c0 = b0 + (uint32_t)j * THREADS + tid;
        if (b0 + (uint32_t)j * THREADS < N) { // warp-uniform
          float next = __shfl_down_sync(0xffffffffu, sf[j], 1);
          if (lane == 31 || c0 + 1u >= N)
            next = nx[j];
          if (c0 < N) {
            mx = fmaxf(mx, sf[j]);
            ok &= (fabsf(sf[j] - next) < p.eps) ? 1 : 0; // strict <, wrap pair included (:413-421)
          }
        }
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

    if (blockIdx.x == 0 && tid == 0)
      stamp_round_end(p, k);

    if (converged || k + 1u == p.max_iter) {
      for (uint32_t c = cb + tid; c < ce; c += THREADS)
        p.out_eigen_vec[c] = ld_cg(Ecur + c) * (ld_cg(Scur + c) / m_k);
      if (blockIdx.x == 0 && tid == 0) {
        *p.out_eigen_val = ld_cg(Scur);
        p.out_iter[0] = converged ? k : p.max_iter;
        p.out_iter[1] = k + 1u;
      }
      drain();
      return;
    }
    m_prev = m_k;
  }
}

} // namespace st
