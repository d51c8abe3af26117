// kernels_wide.cuh -- unit-scheduled round loop for matrices wider than the resident-e kernel takes
// (N > 32768: BASELINE configs 4 and 5, 65536 and 131072 columns).
//
// The eigenvector no longer fits in shared memory, so a round is cut into PHASES: phase w covers the
// 32768-column window [32768 w, 32768 (w + 1)) of every row.  Per phase each CTA
//   1. rebuilds the window of e_k in shared memory (128 KB) from the previous round's vectors,
//      e_k[c] = E_{k-1}[c] * (S_{k-1}[c] / m_{k-1})  (reference similarity_transform.cpp:260; all ones in round 0),
//   2. streams work units exactly like round_loop_sc_kernel does: one unit = one 8192-column chunk of one row,
//      reduced by one warp in the one evaluation order; warp gw of the grid takes unit gw of the phase, the rest are
//      handed out by the monotonic atomic counter, so faster SMs simply take more;
//   3. leaves the chunk sum in `partial`; the warp whose arrival completes a ROW (per-row counter, N / 8192
//      arrivals per round) adds the row's chunk sums left to right and publishes s[r] = t / e_k[r] -- to every
//      rank when sharded.
// CTAs move from phase to phase on their own (the only grid-wide synchronisation is the round barrier), so the
// CTA-wide pause for a window rebuild on one SM is covered by the other SMs' streaming.  The general loop
// (round_loop_kernel) does the same arithmetic with static row blocks per CTA; it stays as the fallback for the
// in-place form, ragged dimensions, bf16 storage and fp64 accumulation.
//
// The full-length eigenvector lives in global memory, double-buffered by round parity like S (slice owners write
// E_k at the start of round k); the round barrier carries the max, so the vector tail is one scan of s for the
// stop test.  Before a warp enters the barrier it prefetches the head of the unit it takes first in the next
// round into its shared-memory slot (cp.async.bulk, as in the resident-e kernel).
//
// Row sums are bit-identical to every other round kernel.  Read-only form, fp32, N % 4 == 0.
#pragma once

#include "kernels_sc.cuh"

namespace st {

constexpr int kWideMaxWindows = 64; // counters reserved per solve: N <= 64 x 32768 columns

template<int MAX_THREADS, int STOP = kStopAbsolute>
__global__ void __launch_bounds__(MAX_THREADS, 1) round_loop_wide_kernel(const RoundParams p)
{
  constexpr int LD = kUnroll;
  constexpr uint32_t kPfFloats = 1024u; // prefetch slot per warp: one 4 KB batch
  const uint32_t THREADS = blockDim.x;
  const uint32_t kWarps = THREADS >> 5;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* pf_all = reinterpret_cast<float*>(smem_raw);        // kWarps x kPfFloats
  float* e_s = pf_all + (size_t)kWarps * kPfFloats;          // the staged window of e_k: <= kResidentCols floats
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(smem_raw + p.mbar_offset);
  __shared__ unsigned int s_cta_max;
  __shared__ float s_m;
  __shared__ int s_abort;

  const uint32_t tid = threadIdx.x;
  const int lane = (int)(tid & 31u);
  const uint32_t warp = tid >> 5;
  const uint32_t N = p.N;
  const uint32_t nv = N >> 2;
  const uint32_t cb = (uint32_t)((uint64_t)N * blockIdx.x / gridDim.x);
  const uint32_t ce = (uint32_t)((uint64_t)N * (blockIdx.x + 1) / gridDim.x);

  const uint32_t R = p.rows;
  const uint32_t nch = (N + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols; // chunks (units) per row
  const uint32_t WIN = p.chunk_cols;               // columns per window: kResidentCols (tests may stage less), a multiple of kChunkCols
  const uint32_t cpw = WIN / (uint32_t)kChunkCols; // chunks per full window
  const uint32_t P = (N + WIN - 1u) / WIN;         // phases (windows) per round
  const uint32_t TW = gridDim.x * kWarps;
  const uint32_t gw = blockIdx.x * kWarps + warp;
  constexpr uint32_t kChunkVec = kChunkCols / 4;

  float* my_pf = pf_all + (size_t)warp * kPfFloats;
  uint64_t* my_bar = mbar_all + warp;
  if (lane == 0)
    mbar_init(my_bar, 1u);
  if (tid == 0)
    s_cta_max = 0u;
  fence_mbarrier_init();
  fence_proxy_async();
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();

  // chunks of window w, and the work unit `cur` (in processing order) of a phase: (local row, global chunk)
  auto chunks_in = [&](uint32_t w) { return min(cpw, nch - w * cpw); };

  uint32_t pf_issued = 0, pf_consumed = 0;
  float m_prev = 1.f;
  for (uint32_t k = 0;; ++k) {
    const uint32_t par = (k + p.flip) & 1u;
    const bool first = k == 0u;
    const float* Sprev = p.S[par ^ 1u];
    const float* Eprev = p.E[par ^ 1u];
    float* Scur = p.S[par];
    float* Ecur = p.E[par];
    float wmax = 0.f;
    const bool backward = p.sweep && (k & 1);
    bool tma_ok = true;

    // E_k for this CTA's slice of the vector (read by every CTA in round k + 1)           :42-43 -> :260, :34 -> :280
    for (uint32_t c = cb + tid; c < ce; c += THREADS)
      Ecur[c] = first ? 1.f : ld_cg(Eprev + c) * (ld_cg(Sprev + c) / m_prev);

    // e_k[r] as the publishing lane needs it (the row's column is usually not in the staged window)
    auto e_at = [&](uint32_t gr) { return first ? 1.f : ld_cg(Eprev + gr) * (ld_cg(Sprev + gr) / m_prev); };
    auto publish = [&](uint32_t rl, float t) {
      const uint32_t gr = p.row0 + rl;
      const float s = t / e_at(gr);
      wmax = fmaxf(wmax, s);
      if (p.world > 1) {
        for (uint32_t g = 0; g < p.world; g++)
          __stcg(p.peer_S[par][g] + gr, s);
      } else {
        __stcg(Scur + gr, s);
      }
    };
    // lane 0: the warp whose arrival completes a row adds its chunk sums left to right
    auto finish_row = [&](uint32_t rl, uint32_t old) {
      if (old % nch != nch - 1u) // the per-row counter is monotonic: + nch per round
        return;
      __threadfence();
      const float* pr = p.partial + (size_t)rl * nch;
      float t = 0.f;
      for (uint32_t c0 = 0; c0 < nch; c0 += 16u) {
        float part[16];
#pragma unroll
        for (uint32_t c = 0; c < 16u; c++)
          part[c] = c0 + c < nch ? ld_cg(pr + c0 + c) : 0.f;
#pragma unroll
        for (uint32_t c = 0; c < 16u; c++)
          if (c0 + c < nch)
            t = (c0 + c == 0u) ? part[c] : t + part[c];
      }
      publish(rl, t);
    };

    for (uint32_t q = 0; q < P; q++) {
      const uint32_t w = backward ? P - 1u - q : q;
      const uint32_t w0 = w * WIN;
      const uint32_t wlen = min(WIN, N - w0);
      const uint32_t cw = chunks_in(w);
      const uint32_t Uw = R * cw;
      const uint32_t D = (p.dynamic && Uw > TW) ? Uw - TW : 0u;
      // CTAs move through the phases on their own, so every window has its own monotonic unit counter (on its own
      // 128-byte line): every warp makes exactly one failing grab per round on it, so round k hands out the values
      // [k (D + TW), k (D + TW) + D) of counter w
      unsigned int* counter = p.phase_counter + 32u * w;
      const uint32_t base = k * (D + TW);

      // ---- stage the window of e_k: every thread's loads in flight before anything depends on them ----
      if (q > 0u)
        __syncthreads(); // every warp is done reading the previous window
      {
        const uint32_t cv = wlen >> 2;
        const uint32_t rotv = (cb >> 2) % cv; // each CTA starts at its own offset: 148 CTAs stay off each other's L2 lines
        const float4* S4 = reinterpret_cast<const float4*>(Sprev + w0);
        const float4* E4 = reinterpret_cast<const float4*>(Eprev + w0);
        float4* e4 = reinterpret_cast<float4*>(e_s);
        for (uint32_t v0 = 0; v0 < cv; v0 += 4u * THREADS) {
          float4 sp[4], ep[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
            const uint32_t vi = idx + rotv < cv ? idx + rotv : idx + rotv - cv;
            const bool active = idx < cv && !first;
            sp[j] = active ? ld_cg(S4 + vi) : make_float4(1.f, 1.f, 1.f, 1.f);
            ep[j] = active ? ld_cg(E4 + vi) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
            if (idx < cv) {
              const uint32_t vi = idx + rotv < cv ? idx + rotv : idx + rotv - cv;
              float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
              if (!first) {
                v.x = ep[j].x * (sp[j].x / m_prev);
                v.y = ep[j].y * (sp[j].y / m_prev);
                v.z = ep[j].z * (sp[j].z / m_prev);
                v.w = ep[j].w * (sp[j].w / m_prev);
              }
              e4[vi] = v;
            }
          }
        }
      }
      __syncthreads();

      // ---- the phase's work units ----                                                     reference :40 (+ :52)
      uint32_t cur = gw;
      bool have = cur < Uw;
      bool first_unit = q == 0u; // the cross-barrier prefetch holds the head of the round's first unit
      uint32_t grabbed = 0, pend_row = 0, pend_old = 0;
      bool pending = false;
      if (D > 0u && lane == 0)
        grabbed = atomicAdd(counter, 1u); // one grab always in flight
      for (;;) {
        if (have) {
          const uint32_t u = backward ? (Uw - 1u - cur) : cur;
          const uint32_t rl = u / cw;
          const uint32_t cl = u - rl * cw;                                 // chunk inside the window
          const uint32_t ch = w * cpw + cl;                                // chunk of the row
          const uint32_t seg_nv = min(kChunkVec, nv - ch * kChunkVec);
          const float4* seg = reinterpret_cast<const float4*>(p.A + (size_t)rl * N) + (size_t)ch * kChunkVec;
          uint32_t npre = 0;
          if (first_unit && pf_consumed < pf_issued) {
            tma_ok = mbar_wait(my_bar, pf_consumed & 1u, p.timeout_ns);
            pf_consumed++;
            npre = min(kPfFloats >> 2, seg_nv);
          }
          const float t = chunk_dot_prefetched<LD, float>(seg, reinterpret_cast<const float4*>(e_s) + cl * kChunkVec, seg_nv,
                                                          lane, reinterpret_cast<const float4*>(my_pf), npre);
          if (lane == 0) {
            if (pending)
              finish_row(pend_row, pend_old); // the atomic issued one unit ago has long returned
            __stcg(p.partial + (size_t)rl * nch + ch, t);
            __threadfence();
            pend_old = atomicAdd(p.row_done + rl, 1u);
            pend_row = rl;
            pending = true;
          }
        }
        first_unit = false;
        if (D > 0u) {
          const uint32_t d = __shfl_sync(0xffffffffu, grabbed, 0) - base;
          if (d >= D)
            break;
          cur = TW + d;
          if (lane == 0)
            grabbed = atomicAdd(counter, 1u);
        } else {
          cur += TW;
          if (cur >= Uw)
            break;
        }
        have = true;
      }
      if (lane == 0 && pending)
        finish_row(pend_row, pend_old);
    }

    // keep the L2->SM pipe busy across the barrier: fetch the head of next round's first unit
    if (k + 1u < p.max_iter) {
      const bool bw = p.sweep && ((k + 1u) & 1u);
      const uint32_t w = bw ? P - 1u : 0u;
      const uint32_t cw = chunks_in(w);
      const uint32_t Uw = R * cw;
      if (gw < Uw) {
        __syncwarp();
        if (lane == 0) {
          const uint32_t u = bw ? (Uw - 1u - gw) : gw;
          const uint32_t rl = u / cw;
          const uint32_t ch = w * cpw + (u - rl * cw);
          const uint32_t seg_nv = min(kChunkVec, nv - ch * kChunkVec);
          const uint32_t bytes = min(kPfFloats >> 2, seg_nv) * 16u;
          fence_proxy_async();
          mbar_arrive_expect_tx(my_bar, bytes);
          bulk_load(my_pf, p.A + (size_t)rl * N + (size_t)ch * kChunkCols, bytes, my_bar);
        }
        pf_issued++;
      }
    }
    if (!tma_ok && lane == 0)
      atomicExch(&p.bar->error, 2u);
    if (lane == 0 && wmax > 0.f)
      atomicMax(&s_cta_max, __float_as_uint(wmax));
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 0u);

    if (!round_barrier(p, k, &s_abort, &s_cta_max, &s_m))
      break;
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 1u);
    const float m_k = s_m; // max over every GPU's rows, carried by the barrier            :41

    // ---- every CTA: circular stop test over the full s ----                              :44
    const float4* S4 = reinterpret_cast<const float4*>(Scur);
    const uint32_t rotv = cb >> 2;
    const float thr = STOP == kStopRelative ? p.eps * m_k : p.eps;
    int ok = 1;
    constexpr int kTailBatch = 8;
    for (uint32_t v0 = 0; v0 < nv; v0 += (uint32_t)kTailBatch * THREADS) {
      float4 t4[kTailBatch];
      float nx[kTailBatch];
#pragma unroll
      for (int j = 0; j < kTailBatch; j++) {
        const uint32_t idx = v0 + tid + (uint32_t)j * THREADS;
        const bool active = idx < nv;
        const uint32_t vi = idx + rotv < nv ? idx + rotv : idx + rotv - nv;
        t4[j] = active ? ld_cg(S4 + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool edge = active && (lane == 31 || idx + 1u >= nv);
        nx[j] = edge ? ld_cg(Scur + (vi + 1u == nv ? 0u : 4u * (vi + 1u))) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kTailBatch; j++) {
        const uint32_t idx = v0 + tid + (uint32_t)j * THREADS;
        if (v0 + (uint32_t)j * THREADS < nv) { // warp-uniform
          float nxt = __shfl_down_sync(0xffffffffu, t4[j].x, 1);
          if (lane == 31 || idx + 1u >= nv)
            nxt = nx[j];
          if (idx < nv) {
            const float4 v = t4[j];
            // strict <, wrap pair included (:413-421); relative test: every pair below eps * m
            ok &= (fabsf(v.x - v.y) < thr) & (fabsf(v.y - v.z) < thr) & (fabsf(v.z - v.w) < thr) & (fabsf(v.w - nxt) < thr);
          }
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0)
      stamp_round_end(p, k);
    const bool converged = __syncthreads_and(ok) != 0;

    if (converged || k + 1u == p.max_iter) {
      // the eigenvector update of this round still happens before the break (:42-50)
      for (uint32_t c = cb + tid; c < ce; c += THREADS)
        p.out_eigen_vec[c] = ld_cg(Ecur + c) * (ld_cg(Scur + c) / m_k);
      if (blockIdx.x == 0 && tid == 0) {
        *p.out_eigen_val = ld_cg(Scur);             // :60-65
        p.out_iter[0] = converged ? k : p.max_iter; // :54
        p.out_iter[1] = k + 1u;
      }
      break;
    }
    m_prev = m_k;
  }
  // a prefetch issued for a round that did not run must land before the CTA exits
  if (lane == 0)
    for (; pf_consumed < pf_issued; pf_consumed++)
      mbar_wait(my_bar, pf_consumed & 1u, p.timeout_ns);
}

} // namespace st
