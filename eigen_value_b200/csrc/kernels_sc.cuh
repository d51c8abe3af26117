// kernels_sc.cuh -- round loop with the eigenvector resident in shared memory (N <= 32768).
//
// When the whole eigenvector fits in shared memory (<= 128 KB) the loop gets much tighter than
// the general, column-chunked round_loop_kernel:
//
//   * e lives in shared memory for the WHOLE solve.  Every CTA updates its private copy in
//     place from the published row sums (e *= s/m, reference similarity_transform.cpp:260);
//     no global E buffers, no rebuild of the chunk per round.
//   * after the round barrier each thread needs ONE L2 round trip (N <= ~10240: its <= 5
//     float4 of s stay in registers between the max / stop reduction and the update of e) or
//     two (larger N: s is re-read), with every load of a batch in flight at once.  The general
//     kernel needs four dependent L2 trips per round.
//   * the work unit is one 8192-column chunk of one row (32 KB), handed out dynamically through
//     one atomic counter once the matrix no longer lives in L2 (N >= 8192), round-robin below;
//     rows of several units are finished, in order, by the warp whose arrival completes them.
//   * the matrix never changes, so before a warp enters the barrier its lane 0 issues ONE
//     TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) of the first
//     PF_BATCHES*4 KB of the unit it will process first in the NEXT round into a private
//     shared-memory slot.  The L2->SM pipe keeps moving matrix bytes while the CTA sits in the
//     barrier and the vector tail; the first batches of the next round are then consumed
//     from shared memory.  A prefetch issued for a round that never runs is drained at exit.
//     When a whole row fits the slot and no warp owns more than one row, the row is fetched
//     once and stays in shared memory for the whole solve.
//   * the CTA size is blockDim.x (a run-time value; every warp is used).
//
// Row sums are bit-identical to round_loop_kernel: same lane / accumulator / fold order.
// Read-only form, N % 4 == 0, N <= kResidentCols only.
#pragma once

#include "kernels.cuh"

namespace st {

constexpr int kResidentCols = 32768; // largest N whose eigenvector is kept in shared memory

// One work unit = one 8192-column chunk of one row (<= 32 KB), reduced by one warp.  The
// leading `npre` float4 come from the prefetched shared-memory tile, the rest from global
// memory; npre is a multiple of 256 (one batch = 32 lanes x 8 accumulators) or the whole unit.
// Evaluation order = round_loop_kernel's: vector j of the chunk goes to lane j % 32,
// accumulator (j / 32) % 8, the accumulators are folded pairwise and the lanes by an
// xor-shuffle tree; the chunk sums of a row are added left to right by whoever finishes the
// row.  LD = independent 128-bit loads in flight per lane (8 or 16); it does not affect the order.
template<int LD, typename ACC = float>
__device__ __forceinline__ float
chunk_dot_prefetched(const float4* __restrict__ a, const float4* es, uint32_t nv, int lane,
                     const float4* pf, uint32_t npre)
{
  static_assert(LD % kUnroll == 0, "loads in flight must be a multiple of the accumulator count");
  ACC acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = ACC(0);
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
  return (float)warp_sum(acc[0]);
}

// VEC = 1: dim % 4 != 0 (or a matrix that is not 16-byte aligned).  Rows then start on 4-byte boundaries, the unit of
// the evaluation order is a single float (element j of a chunk -> lane j % 32, accumulator (j / 32) % 8: the order
// round_loop_kernel<1, ...> and the oracle's SUM_CUDA already use for these dimensions), the matrix is streamed with
// 32 scalar loads in flight per lane and there are no prefetch slots (bulk copies need 16-byte alignment).  Everything
// else -- units, scheduling, publication, barrier, vector tail -- is shared.  (A per-warp two-slot ring of bulk copies
// of the 16-byte-aligned superset of each 1024-float batch, read back from shared memory at any 4-byte offset -- same
// bits, sector-exact traffic -- ran at HALF the rate of the plain 4-byte loads: 89.7 against 43.8 us per round at Hilbert
// 8191, profiles/r2_c25_ragged_ring_vs_scalar_loads.json; removed, commit 22a4650.)
template<int MAX_THREADS, int PF_BATCHES, int STOP = kStopAbsolute, typename T = float, typename ACC = float, int VEC = 4>
__global__ void __launch_bounds__(MAX_THREADS, 1) round_loop_sc_kernel(const RoundParams p)
{
  constexpr int LD = kUnroll; // independent 128-bit loads in flight per lane (16 measured no better, profiles/r1_sweep_resident_e_variants.txt)
  static_assert(sizeof(ACC) == 4 || sizeof(T) == 4, "fp64 accumulation is built for fp32 storage");
  static_assert(VEC == 4 || (VEC == 1 && PF_BATCHES == 0 && sizeof(T) == 4), "scalar units: fp32 storage, no prefetch slots");
  // bf16 storage (p.A points to bf16 data, N % 4 == 0): a work unit is still one 8192-column chunk
  // of one row (16 KB); built without the cross-barrier prefetch
  constexpr bool kBf16 = sizeof(T) == 2;
  static_assert(!kBf16 || PF_BATCHES == 0, "bf16 storage: no prefetch slots");
  // fp8 storage (p.A points to e4m3 codes, p.row_scale to the row scales, N % 4 == 0): a work unit is 8 KB
  constexpr bool kFp8 = sizeof(T) == 1;
  static_assert(!kFp8 || (PF_BATCHES == 0 && sizeof(ACC) == 4 && VEC == 4), "fp8 storage: no prefetch slots, fp32 accumulation");
  const uint32_t THREADS = blockDim.x; // run-time CTA size (a multiple of 32, <= MAX_THREADS)
  const uint32_t kWarps = THREADS >> 5;
  constexpr uint32_t kPfFloats = PF_BATCHES * 1024u; // prefetch slot per warp

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* pf_all = reinterpret_cast<float*>(smem_raw);        // kWarps x kPfFloats
  float* e_s = pf_all + (size_t)kWarps * kPfFloats;          // N floats, lives across rounds
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(smem_raw + p.mbar_offset);
  __shared__ unsigned int s_cta_max; // bits of the largest row sum this CTA published in the current round
  __shared__ float s_m;              // max(0, max_r s[r]) of the round, out of the barrier
  __shared__ int s_abort;

  const uint32_t tid = threadIdx.x;
  const int lane = (int)(tid & 31u);
  const uint32_t warp = tid >> 5;
  const uint32_t N = p.N;
  const uint32_t nv = N >> 2;

  // slice of the N-vector this CTA writes at the end (and its phase offset in the tail scan)
  const uint32_t cb = (uint32_t)((uint64_t)N * blockIdx.x / gridDim.x);
  const uint32_t ce = (uint32_t)((uint64_t)N * (blockIdx.x + 1) / gridDim.x);
  // Scheduling: the work unit is one 8192-column chunk of one row (32 KB; a row of N = 32768
  // is 4 units).  Warp gw of the grid takes unit gw first (static, so its head can be
  // prefetched across the barrier); the remaining U - TW units are handed out through one
  // atomic counter, so SMs that stream faster (L2 die locality) simply take more, and the
  // end-of-round straggle is one 32 KB unit whatever N is.  A unit is always reduced by one
  // warp in one fixed order and the chunk sums of a row are added left to right: WHO takes a
  // unit cannot change a bit.
  const uint32_t R = p.rows;
  const uint32_t nch = (N + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols; // units per row
  const uint32_t U = R * nch;
  const uint32_t TW = gridDim.x * kWarps;
  const uint32_t gw = blockIdx.x * kWarps + warp;
  // (p.dynamic == 0: the remaining units are taken round-robin, gw + TW, gw + 2 TW, ... -- for
  // small, L2-resident matrices the one-address atomic stream costs more than the skew it removes)
  const uint32_t D = (p.dynamic && U > TW) ? U - TW : 0u; // dynamically scheduled units per round
  constexpr uint32_t kChunkVec = kChunkCols / 4;
  float* my_pf = pf_all + (size_t)warp * kPfFloats;
  uint64_t* my_bar = mbar_all + warp;
  if (PF_BATCHES > 0 && lane == 0)
    mbar_init(my_bar, 1u);
  if (tid == 0)
    s_cta_max = 0u;
  for (uint32_t c = tid; c < N; c += THREADS)
    e_s[c] = kFp8 ? kFp8EigenScale : 1.f; // initialise_eigen_vector, reference :267-284 (fp8 storage keeps e * 2^120, see dot_acc_fp8)
  fence_mbarrier_init();
  fence_proxy_async();
  __syncthreads();

  if (blockIdx.x == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();

  uint32_t pf_issued = 0, pf_consumed = 0; // bulk copies issued / waited for by this warp

  // Matrix resident on chip: when no warp has more than one unit (U <= TW) and a whole unit fits
  // the warp's prefetch slot, the "prefetch" is issued once, before the first round, and the
  // unit simply stays in shared memory for the whole solve -- N = 1024 (4 KB rows, 64 CTAs) and
  // N = 2048 (8 KB rows, 128 CTAs) never touch the matrix in L2 again after round 0.
  const bool resident = PF_BATCHES > 0 && U <= TW && min(kChunkVec, nv) <= (kPfFloats >> 2);
  if (resident && gw < U) {
    if (lane == 0) {
      const uint32_t rl = gw / nch;
      const uint32_t ch = gw - rl * nch;
      const uint32_t bytes = min(kChunkVec, nv - ch * kChunkVec) * 16u;
      mbar_arrive_expect_tx(my_bar, bytes);
      bulk_load(my_pf, p.A + (size_t)rl * N + (size_t)ch * kChunkCols, bytes, my_bar);
    }
    pf_issued++;
  }

  for (uint32_t k = 0;; ++k) {
    const uint32_t par = (k + p.flip) & 1u; // buffer set of this round
    float* Scur = p.S[par];
    float wmax = 0.f; // lane 0: largest row sum this warp published this round; the reference zero-fills the max cell (:169)
    const bool backward = !resident && p.sweep && (k & 1); // resident units never change owner

    // ---- the pass over the matrix ----                                   reference :40 (+ :52)
    bool tma_ok = true;
    {
      // publish s[r] = (A.e)[r] / e[r] -- to every rank when sharded
      auto publish = [&](uint32_t rl, float t) {
        const uint32_t gr = p.row0 + rl;
        const float s = kFp8 ? (t * p.row_scale[rl]) / (e_s[gr] * kFp8EigenUnscale) : t / e_s[gr];
        wmax = fmaxf(wmax, s); // find_max, reference :154-227 (fmaxf drops NaNs, like the running max of the old scan)
        if (p.world > 1) {
          for (uint32_t g = 0; g < p.world; g++)
            __stcg(p.peer_S[par][g] + gr, s);
        } else {
          __stcg(Scur + gr, s);
        }
      };
      // lane 0: a row of several units is finished by the warp whose arrival completes it
      auto finish_row = [&](uint32_t rl, uint32_t old) {
        if (old % nch == nch - 1u) { // the per-row counter is monotonic: + nch per round
          __threadfence();
          float part[4];
#pragma unroll
          for (uint32_t c = 0; c < 4u; c++)
            part[c] = c < nch ? ld_cg(p.partial + (size_t)rl * nch + c) : 0.f;
          float t = part[0];
#pragma unroll
          for (uint32_t c = 1; c < 4u; c++)
            if (c < nch)
              t = t + part[c]; // left to right, like round_loop_kernel's part_s[rl] += t
          publish(rl, t);
        }
      };
      // the unit counter is monotonic too: every warp makes exactly one failing grab per round,
      // so round k hands out the values [k * (D + TW), k * (D + TW) + D)
      const uint32_t base = k * (D + TW);
      uint32_t cur = gw;
      bool have = cur < U;
      bool first_unit = true;
      uint32_t grabbed = 0, pend_row = 0, pend_old = 0;
      bool pending = false;
      if (p.dynamic && lane == 0)
        grabbed = atomicAdd(&p.bar->row_counter, 1u); // one grab always in flight
      for (;;) {
        if (have) {
          const uint32_t u = backward ? (U - 1u - cur) : cur;
          const uint32_t rl = u / nch;
          const uint32_t ch = u - rl * nch;
          const uint32_t seg_nv = min(kChunkVec, nv - ch * kChunkVec);
          const float4* seg = reinterpret_cast<const float4*>(p.A + (size_t)rl * N) + ch * kChunkVec;
          uint32_t npre = 0;
          if (PF_BATCHES > 0 && first_unit && pf_consumed < pf_issued) {
            tma_ok = mbar_wait(my_bar, pf_consumed & 1u, p.timeout_ns);
            pf_consumed++;
            npre = min(kPfFloats >> 2, seg_nv);
          } else if (resident && first_unit) {
            npre = seg_nv; // landed before round 0 and never evicted
          }
          float t;
          if (VEC == 1) {
            t = row_dot_readonly<1, false, ACC>(p.A + (size_t)rl * N + (size_t)ch * kChunkCols, e_s + (size_t)ch * kChunkCols,
                                                min((uint32_t)kChunkCols, N - ch * (uint32_t)kChunkCols), lane);
          } else if (kFp8) {
            // one 32-bit word = 4 columns = one float4 of the eigenvector chunk
            const uint32_t* seg8 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const fp8_t*>(p.A) + (size_t)rl * N) +
                                   ch * kChunkVec;
            t = row_dot_fp8<4 * LD>(seg8, reinterpret_cast<const float4*>(e_s) + ch * kChunkVec, seg_nv, lane);
          } else if (kBf16) {
            // one 64-bit word = 4 columns = one float4 of the eigenvector chunk
            const uint2* seg16 = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16_t*>(p.A) + (size_t)rl * N) +
                                 ch * kChunkVec;
            t = row_dot_bf16<2 * LD>(seg16, reinterpret_cast<const float4*>(e_s) + ch * kChunkVec, seg_nv, lane);
          } else {
            t = chunk_dot_prefetched<LD, ACC>(seg, reinterpret_cast<const float4*>(e_s) + ch * kChunkVec, seg_nv, lane,
                                              reinterpret_cast<const float4*>(my_pf), npre);
          }
          if (lane == 0) {
            if (pending)
              finish_row(pend_row, pend_old); // the atomic issued one unit ago has long returned
            pending = false;
            if (nch == 1u) {
              publish(rl, t);
            } else {
              __stcg(p.partial + (size_t)rl * nch + ch, t);
              __threadfence();
              pend_old = atomicAdd(p.row_done + rl, 1u);
              pend_row = rl;
              pending = true;
            }
          }
        }
        first_unit = false;
        if (p.dynamic) {
          const uint32_t d = __shfl_sync(0xffffffffu, grabbed, 0) - base;
          if (d >= D)
            break;
          cur = TW + d;
          if (lane == 0)
            grabbed = atomicAdd(&p.bar->row_counter, 1u);
        } else {
          cur += TW;
          if (cur >= U)
            break;
        }
        have = true;
      }
      if (lane == 0 && pending)
        finish_row(pend_row, pend_old);
    }
    // keep the L2->SM pipe busy across the barrier: fetch the head of next round's first unit
    if (PF_BATCHES > 0 && !resident && gw < U && k + 1u < p.max_iter) {
      __syncwarp();
      if (lane == 0) {
        const uint32_t u = (p.sweep && ((k + 1u) & 1u)) ? (U - 1u - gw) : gw;
        const uint32_t rl = u / nch;
        const uint32_t ch = u - rl * nch;
        const uint32_t seg_nv = min(kChunkVec, nv - ch * kChunkVec);
        const uint32_t bytes = min(kPfFloats >> 2, seg_nv) * 16u;
        fence_proxy_async();
        mbar_arrive_expect_tx(my_bar, bytes);
        bulk_load(my_pf, p.A + (size_t)rl * N + (size_t)ch * kChunkCols, bytes, my_bar);
      }
      pf_issued++;
    }
    if (!tma_ok && lane == 0)
      atomicExch(&p.bar->error, 2u);
    if (lane == 0 && wmax > 0.f)
      atomicMax(&s_cta_max, __float_as_uint(wmax));
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 0u);

    // the barrier's first __syncthreads orders the warps' atomicMax before thread 0 reads the CTA's max
    if (!round_barrier(p, k, &s_abort, &s_cta_max, &s_m))
      break;
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 1u);
    const float m_k = s_m; // max over every GPU's rows, carried by the barrier            :41

    // ---- every CTA: circular stop test over the full s and the e update, ONE pass ----   :42-44
    // 128-bit L2 loads, all loads of a batch (8 per thread) issued before anything depends on them.  The
    // circular neighbour of a vector's last element is the next lane's first element (shuffle); only
    // lane 31 reads it from memory -- the reference does the same with shuffle_down + one global read
    // per sub-group (:413-417).  Every CTA scans the same N floats at the same moment, so each starts at
    // its own offset (cb) to stay off the other CTAs' L2 lines; the stop flag is order-independent.
    const float4* S4 = reinterpret_cast<const float4*>(Scur);
    float4* e4_s = reinterpret_cast<float4*>(e_s);
    const uint32_t rotv = cb >> 2;
    const float thr = STOP == kStopRelative ? p.eps * m_k : p.eps;
    int ok = 1;
    // scalar units (N % 4 != 0) scan s in 128-bit vectors too -- s and e are this kernel's own, 16-byte aligned and
    // padded buffers; only the LAST vector is partial (N - 4 (nvt - 1) valid elements, see below)
    const uint32_t nvt = VEC == 4 ? nv : (N + 3u) >> 2;
    auto vec_index = [&](uint32_t idx) {
      const uint32_t v = idx + rotv;
      return v < nvt ? v : v - nvt;
    };
    constexpr int kTailBatch = 8;
    for (uint32_t v0 = 0; v0 < nvt; v0 += (uint32_t)kTailBatch * THREADS) {
      float4 t4[kTailBatch];
      float nx[kTailBatch];
#pragma unroll
      for (int j = 0; j < kTailBatch; j++) {
        const uint32_t idx = v0 + tid + (uint32_t)j * THREADS;
        const bool active = idx < nvt;
        const uint32_t vi = vec_index(idx);
        t4[j] = active ? ld_cg(S4 + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool edge = active && (lane == 31 || idx + 1u >= nvt);
        nx[j] = edge ? ld_cg(Scur + (vi + 1u == nvt ? 0u : 4u * (vi + 1u))) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kTailBatch; j++) {
        const uint32_t idx = v0 + tid + (uint32_t)j * THREADS;
        if (v0 + (uint32_t)j * THREADS < nvt) { // warp-uniform
          float nxt = __shfl_down_sync(0xffffffffu, t4[j].x, 1);
          if (lane == 31 || idx + 1u >= nvt)
            nxt = nx[j];
          if (idx < nvt) {
            float4 v = t4[j];
            if (VEC == 1 && vec_index(idx) == nvt - 1u) {
              // partial last vector: its circular successor is s[0] (= nxt here: the next vector is vector 0).  The
              // padding takes that value, so the pairs past the last element compare s[0] with itself -- they pass
              // exactly when the wrap pair's s[0] is not NaN -- and the padding of e is updated but never read.
              const uint32_t valid = N - 4u * (nvt - 1u);
              if (valid < 2u)
                v.y = nxt;
              if (valid < 3u)
                v.z = nxt;
              if (valid < 4u)
                v.w = nxt;
            }
            // strict <, wrap pair included (:413-421).  Relative test: every pair below eps * m is the same
            // statement as "the largest difference is below eps * m" (a NaN difference fails either way)
            ok &= (fabsf(v.x - v.y) < thr) & (fabsf(v.y - v.z) < thr) & (fabsf(v.z - v.w) < thr) &
                  (fabsf(v.w - nxt) < thr);
            // e_{k+1} = e_k * (s_k / m_k), in place in shared memory                       :260
            const uint32_t vi = vec_index(idx);
            float4 e = e4_s[vi];
            e.x = e.x * (v.x / m_k);
            e.y = e.y * (v.y / m_k);
            e.z = e.z * (v.z / m_k);
            e.w = e.w * (v.w / m_k);
            e4_s[vi] = e;
          }
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0)
      stamp_round_end(p, k);
    const bool converged = __syncthreads_and(ok) != 0; // also orders the e update before the next pass

    if (converged || k + 1u == p.max_iter) {
      // the eigenvector update of this round still happens before the break (:42-50)
      for (uint32_t c = cb + tid; c < ce; c += THREADS)
        p.out_eigen_vec[c] = kFp8 ? e_s[c] * kFp8EigenUnscale : e_s[c];
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
