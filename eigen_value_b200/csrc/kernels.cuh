// kernels.cuh -- sm_100a device code for the similarity_transform() round loop.
//
// One persistent, cooperatively launched kernel (round_loop_kernel) runs the WHOLE loop of
// reference similarity_transform.cpp:39-53 on the device: per round it makes a single
// 128-bit-vectorised pass over this GPU's rows of the matrix, reduces each row inside one
// warp (shuffle tree, no atomics, deterministic), publishes the row sums (to peer GPUs too
// when the matrix is row-block sharded), crosses ONE grid-wide barrier, and then every CTA
// redundantly reduces the N-length row-sum vector to the max, the circular stop flag and the
// next eigenvector -- the work of the reference's five kernels + three fills + blocking
// host read per round (:40-52).
//
// This file holds what every round kernel shares (parameters, the round barrier with its
// cross-GPU exchange, the canonical row reduction) and the GENERAL loop: any N, both forms,
// column-chunked scale vector.  The specialised loops build on it:
//   kernels_cluster.cuh  N <= 512    matrix resident in the shared memory of one cluster
//   kernels_sc.cuh       N <= 32768  eigenvector resident in shared memory (the default; dim % 4 != 0 on scalar units)
//   kernels_wide.cuh     N >  32768  the same work-unit scheduling, eigenvector staged one 32768-column window at a time
//
// Opt-in variants are template parameters, so the default instantiations stay exactly the measured
// code: STOP (the reference's absolute stop test | relative), T (fp32 | bf16 | fp8 STORAGE of the matrix),
// ACC (fp32 | fp64 accumulators).  None of them changes the evaluation order.
//
// The small standalone kernels at the bottom are the per-kernel entry points mirroring the
// reference's L1 functions (similarity_transform.cpp:77-460) and the input generators
// (utils.cpp:136-154, :124-134).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// Every line of inline PTX sits behind this one header, so that the kernel LOGIC in this directory
// can also be compiled for the host by the CPU emulation harness (tests/cuda_emu), which supplies
// its own implementation of the same functions through ST_PTX_HEADER.
#ifdef ST_PTX_HEADER
#include ST_PTX_HEADER
#else
#include "ptx.cuh"
#endif

namespace st {

constexpr int kMaxWorld = 8;      // ST_MAX_WORLD
constexpr int kChunkCols = 8192;  // columns of one work unit: a row is reduced in chunks of this many columns, added left to right
constexpr int kWindowCols = 32768; // general loop: columns of the scale vector staged in shared memory at a time (4 chunks)
constexpr int kUnroll = 8;        // independent 128-bit loads in flight per lane

enum : int
{
  kFormReadOnly = 0, // s = (A.e)/e
  kFormInPlace = 1   // W <- D^-1 W D, s = rowsum(W)
};

// Stop test of a round (template parameter STOP of the round kernels).
//   absolute: every circular adjacent pair |s[r] - s[r+1]| < eps        -- the reference's test
//             (similarity_transform.cpp:413-421), the default and the only parity-relevant one
//   relative: max_r |s[r] - s[r+1]| < eps * max(0, max_r s[r])          -- scale-free extension
//             (SURVEY 8(f) rank 3): the absolute test can never hold once one ulp of lambda
//             exceeds eps (uniform matrices from N = 16384 up)
enum : int
{
  kStopAbsolute = 0,
  kStopRelative = 1
};

// Relative stop test: the largest adjacent difference is reduced as the BIT PATTERN of a
// non-negative float -- for x >= 0 the unsigned order of the bits is the numeric order, and a
// NaN (0x7fc00000...) sorts above +inf, so the max is NaN-propagating and a NaN anywhere makes
// `dmax < threshold` false, exactly like the per-pair comparison of the absolute test.
__device__ __forceinline__ uint32_t
diff_bits(float a, float b)
{
  return __float_as_uint(fabsf(a - b));
}

// Barrier words live on their own 128-byte lines.
struct alignas(128) BarrierState
{
  unsigned int count; // monotonically increasing arrival counter
  unsigned int pad0[31];
  unsigned int error; // set by any CTA whose wait timed out (1) or whose bulk-copy wait did (2)
  unsigned int pad1[31];
  unsigned int row_counter; // resident-e / wide kernels: dynamic work-unit scheduling (monotonic)
  unsigned int pad2[31];
  // largest row sum this GPU published in a round, by round parity: ((k + 1) << 32) | bits of max(0, max_r s[r]).
  // The round tag makes the word monotonic, so it is never reset: round k + 2 simply wins over round k.
  unsigned long long smax[2];
  unsigned int pad3[28];
};

struct RoundParams
{
  const float* A; // this GPU's rows of the input: rows x N, row-major (never written)
  const float* row_scale; // fp8 storage only: one power-of-two scale per owned row
  float* W;       // in-place form only: working copy, rows x N
  uint32_t N;     // matrix dimension
  uint32_t row0;  // first global row owned by this GPU
  uint32_t rows;  // rows owned by this GPU
  float* S[2];    // full-length row-sum vectors, double-buffered by round parity (local)
  float* E[2];    // full-length eigenvector, double-buffered by round parity (local)
  float eps;
  uint32_t max_iter;
  int sweep;           // 1: alternate the row order every round (L2 reuse of the pass tail)
  int dynamic;         // resident-e kernel: hand out work units through an atomic counter
  uint32_t keep_rows_pct; // share of each CTA's rows loaded L2 evict_last (rest evict_first); 0: no hints
  uint32_t chunk_cols; // columns of the scale vector staged at a time (general loop: <= kWindowCols, a multiple of kChunkCols or N)
  uint32_t mbar_offset; // TMA variant: byte offset of the mbarrier array in dynamic smem
  BarrierState* bar;
  float* partial;         // resident-e kernel: chunk sums of multi-unit rows, rows x units
  unsigned int* row_done; // resident-e kernel: per-row arrival counters (monotonic)
  unsigned int* phase_counter; // wide kernel: one monotonic work-unit counter per 32768-column window, 128 bytes apart
  unsigned long long timeout_ns;
  // row-block sharding (world == 1: unused)
  uint32_t rank, world;
  float* peer_S[2][kMaxWorld];            // S buffers of every rank (own entry == S[b])
  unsigned long long* peer_arrive[kMaxWorld]; // arrival counter of every rank (ExchangeHeader::arrive)
  unsigned int* peer_smax3[kMaxWorld];        // max slots of every rank (ExchangeHeader::smax3)
  unsigned long long arrive_base;         // what this group's counters had reached before this solve
  unsigned long long round_base;          // rounds this group ran before this solve (slot = (round_base + k) % 3)
  uint32_t flip;                          // parity offset of this solve: round k uses buffer set (k + flip) & 1, so that
                                          // a rank that starts the next solve early never touches what a slower peer still reads
  // results
  float* out_eigen_vec;  // N floats (device)
  float* out_eigen_val;  // 1 float  (device)
  uint32_t* out_iter;    // [0] iter_count, [1] passes
  unsigned long long* round_ts; // max_iter + 1 globaltimer stamps
  unsigned long long* phase_ts; // 3 per round, CTA 0: pass done, barrier passed, tail done
  uint32_t ts_rounds;           // rounds that have stamp storage (later rounds are not stamped)
};

// ---------------------------------------------------------------------------------------
// per-round instrumentation (the PTX helpers live in ptx.cuh)
// ---------------------------------------------------------------------------------------

// written by one thread of CTA 0
__device__ __forceinline__ void
stamp_phase(const RoundParams& p, uint32_t k, uint32_t which)
{
  if (k < p.ts_rounds)
    p.phase_ts[3u * k + which] = globaltimer_ns();
}
__device__ __forceinline__ void
stamp_round_end(const RoundParams& p, uint32_t k)
{
  if (k < p.ts_rounds) {
    const unsigned long long t = globaltimer_ns();
    p.round_ts[k + 1u] = t;
    p.phase_ts[3u * k + 2u] = t;
  }
}

// Bounded wait for a bulk copy (cp.async.bulk ... mbarrier::complete_tx): a wrong byte count would
// otherwise hang the GPU.
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

// L2-coherent loads/stores for data other CTAs (or other GPUs) write during the kernel.
__device__ __forceinline__ float
ld_cg(const float* p)
{
  return __ldcg(p);
}
__device__ __forceinline__ float4
ld_cg(const float4* p)
{
  return __ldcg(p);
}

__device__ __forceinline__ float
warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// one row segment, one warp
// ---------------------------------------------------------------------------------------
template<int VEC>
struct Vec;
template<>
struct Vec<4>
{
  using type = float4;
};
template<>
struct Vec<1>
{
  using type = float;
};

// Independent loads a lane keeps in flight in the row reductions: 4 KB per warp either way.
template<int VEC>
__host__ __device__ constexpr int
kLoadsInFlight()
{
  return VEC == 4 ? kUnroll : 4 * kUnroll;
}

__device__ __forceinline__ float
dot_acc(float4 a, float4 e, float acc)
{
  acc = fmaf(a.x, e.x, acc);
  acc = fmaf(a.y, e.y, acc);
  acc = fmaf(a.z, e.z, acc);
  acc = fmaf(a.w, e.w, acc);
  return acc;
}
__device__ __forceinline__ float
dot_acc(float a, float e, float acc)
{
  return fmaf(a, e, acc);
}

// fp64 ACCUMULATION (opt-in): the same lane / accumulator / fold / tree order with
// double accumulators.  The product of two floats is exact in double, so each step rounds once (in
// double) and the row sum is rounded to float once at the end (oracle: ORACLE_SUM_CUDA_F64).
__device__ __forceinline__ double
dot_acc(float4 a, float4 e, double acc)
{
  acc = fma((double)a.x, (double)e.x, acc);
  acc = fma((double)a.y, (double)e.y, acc);
  acc = fma((double)a.z, (double)e.z, acc);
  acc = fma((double)a.w, (double)e.w, acc);
  return acc;
}
__device__ __forceinline__ double
dot_acc(float a, float e, double acc)
{
  return fma((double)a, (double)e, acc);
}
__device__ __forceinline__ double
warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- bf16 STORAGE of the matrix (opt-in, outside reference parity) ----------
// The matrix is held as bfloat16 (2 bytes per element: half the HBM bytes per round); everything
// else -- the eigenvector, the row sums, every accumulation -- stays fp32.  bf16 -> fp32 is exact
// (the 16 bits become the high half of the float), so a solve on bf16 storage returns exactly the
// bits an fp32-storage solve returns on the bf16-rounded matrix, in the fp32 kernels' own order: the unit
// is one 64-bit load = FOUR consecutive elements, folded into its accumulator with four sequential FMAs;
// unit j -> lane j % 32, accumulator (j / 32) % 8 (oracle: ORACLE_SUM_CUDA on oracle.to_bf16(mat)).
// (The first build loaded 128 bits = 8 elements per lane: a lane then needs two float4 of e, 32 bytes apart from
// its neighbour's -- a 2-way bank conflict on every LDS.128, 6.0 TB/s of bf16 bytes.  With 64-bit loads the 32
// lanes read 32 consecutive float4 of e, and 16 loads in flight per lane keep the same 4 KB per warp on its way.)
struct bf16_t
{
  unsigned short bits;
};

// little-endian: the element at the lower address is the low half of the 32-bit word
__device__ __forceinline__ float
bf16_lo(uint32_t w)
{
  return __uint_as_float(w << 16);
}
__device__ __forceinline__ float
bf16_hi(uint32_t w)
{
  return __uint_as_float(w & 0xffff0000u);
}

__device__ __forceinline__ float
dot_acc(uint2 a, float4 e, float acc)
{
  acc = fmaf(bf16_lo(a.x), e.x, acc);
  acc = fmaf(bf16_hi(a.x), e.y, acc);
  acc = fmaf(bf16_lo(a.y), e.z, acc);
  acc = fmaf(bf16_hi(a.y), e.w, acc);
  return acc;
}

// One row segment of bf16 storage, one warp: `nu` units of 4 elements starting at `a`; the matching
// eigenvector entries are es[j] (float4, shared memory).  Same loop shape as row_dot_readonly; LDN = independent
// 64-bit loads in flight per lane.
template<int LDN>
__device__ __forceinline__ float
row_dot_bf16(const uint2* __restrict__ a, const float4* es, uint32_t nu, int lane)
{
  static_assert(LDN % kUnroll == 0, "loads in flight must be a multiple of the accumulator count");
  float acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = 0.f;
  uint32_t i = (uint32_t)lane;
  for (; i + 32u * (LDN - 1) < nu; i += 32u * LDN) {
    uint2 v[LDN];
#pragma unroll
    for (int u = 0; u < LDN; u++)
      v[u] = ld_stream(a + i + 32u * u);
#pragma unroll
    for (int u = 0; u < LDN; u++)
      acc[u % kUnroll] = dot_acc(v[u], es[i + 32u * u], acc[u % kUnroll]);
  }
#pragma unroll
  for (int u = 0; u < LDN; u++) {
    const uint32_t j = i + 32u * u;
    if (j < nu)
      acc[u % kUnroll] = dot_acc(ld_stream(a + j), es[j], acc[u % kUnroll]);
  }
#pragma unroll
  for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
    for (int u = 0; u < s; u++)
      acc[u] += acc[u + s];
  return warp_sum(acc[0]);
}

// ---- fp8 STORAGE of the matrix (opt-in, outside reference parity) -----------
// One byte per element (e4m3: 4 exponent bits, 3 mantissa bits, no infinities) and ONE power-of-two scale per row:
// A[r][c] ~= row_scale[r] * q[r][c], with the scale chosen so that the row's largest magnitude lands in (224, 448]
// (st_convert_f32_to_fp8).  A quarter of the HBM bytes per round; e, s and every accumulation stay fp32.  e4m3 -> fp32
// is exact and the scale is a power of two, so the solve returns exactly the bits of an fp32 solve of the dequantised
// matrix in the fp32 kernels' own order: the unit is one 32-bit word = FOUR consecutive elements, folded into its
// accumulator with four sequential FMAs; unit j -> lane j % 32, accumulator (j / 32) % 8; the row's sum is multiplied
// by its scale once, before the division by e[r] (oracle: ORACLE_SUM_CUDA on oracle.to_fp8_rows(mat)).  A row
// that holds a NaN gets the scale NaN: its sum is NaN in every round, as it would be in fp32.
//
// Why 4-byte loads: with one 128-bit load per lane (16 elements; the first build) a lane needs FOUR float4 of e from
// shared memory per load, 64 bytes apart from its neighbour's -- a 4-way bank conflict on every LDS.128, which capped
// the kernel at 2.0 TB/s of fp8 bytes whatever the decode cost (profiles/r2_c23_storage_fp8_cvt_decode.json,
// r2_c24_storage_fp8_integer_decode.json).  With one word per lane the 32 lanes read 32 consecutive float4 of e
// (conflict-free) and a warp request is still a full 128-byte line; 32 loads in flight per lane keep 4 KB per warp
// on its way, as in every other build.
struct fp8_t
{
  unsigned char bits;
};

// Decode without the conversion pipe: the 7 magnitude bits of a code, moved to bits 26..20 of a float, ARE the code's
// value times 2^-120 -- normal codes land on normal floats with the exponent field e (value 2^(e-127) (1 + m/8)
// instead of 2^(e-7) (1 + m/8)), subnormal codes on subnormal floats (m 2^-129 instead of m 2^-9), and FFMA takes
// subnormal inputs at full speed.  So the kernels keep the eigenvector they multiply with PRE-SCALED by 2^120 (exact),
// and fmaf(raw, e * 2^120, acc) rounds the same real number as fmaf(q, e, acc).  Two or three integer operations per
// element: byte i to the top of the word, arithmetic shift right by 4 (the sign stays in bit 31), one mask.  The code
// 0x7f is therefore the finite value 480, not a NaN: st_convert_f32_to_fp8 never emits it for finite input and marks a
// row that holds a NaN through its scale instead.
constexpr float kFp8EigenScale = 0x1p120f;
constexpr float kFp8EigenUnscale = 0x1p-120f;

__device__ __forceinline__ float
fp8_raw(uint32_t w, int byte)
{
  const int32_t top = (int32_t)(w << (24 - 8 * byte));
  return __uint_as_float((uint32_t)(top >> 4) & 0x87f00000u);
}

// e: the matching four eigenvector entries, pre-scaled by 2^120
__device__ __forceinline__ float
dot_acc_fp8(uint32_t w, float4 e, float acc)
{
  acc = fmaf(fp8_raw(w, 0), e.x, acc);
  acc = fmaf(fp8_raw(w, 1), e.y, acc);
  acc = fmaf(fp8_raw(w, 2), e.z, acc);
  acc = fmaf(fp8_raw(w, 3), e.w, acc);
  return acc;
}

// One row segment of fp8 storage, one warp: `nw` words of four elements starting at `a`; the matching eigenvector
// entries are es[j] (float4, shared memory, pre-scaled).  Same loop shape as row_dot_readonly<1>: LDN 4-byte loads in
// flight per lane, all issued before the first is decoded.  (Software-pipelining the loop over half-batches -- the next
// 16 words in flight while 16 are decoded -- measured SLOWER: 401 against 309 us per round at Hilbert 32768,
// profiles/r2_c26_storage_and_ragged.json against r2_c25_storage_fp8_bf16_word_units.json; reverted.)
template<int LDN>
__device__ __forceinline__ float
row_dot_fp8(const uint32_t* __restrict__ a, const float4* es, uint32_t nw, int lane)
{
  static_assert(LDN % kUnroll == 0, "loads in flight must be a multiple of the accumulator count");
  float acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = 0.f;
  uint32_t i = (uint32_t)lane;
  for (; i + 32u * (LDN - 1) < nw; i += 32u * LDN) {
    uint32_t v[LDN];
#pragma unroll
    for (int u = 0; u < LDN; u++)
      v[u] = ld_stream(a + i + 32u * u);
#pragma unroll
    for (int u = 0; u < LDN; u++)
      acc[u % kUnroll] = dot_acc_fp8(v[u], es[i + 32u * u], acc[u % kUnroll]);
  }
#pragma unroll
  for (int u = 0; u < LDN; u++) {
    const uint32_t j = i + 32u * u;
    if (j < nw)
      acc[u % kUnroll] = dot_acc_fp8(ld_stream(a + j), es[j], acc[u % kUnroll]);
  }
#pragma unroll
  for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
    for (int u = 0; u < s; u++)
      acc[u] += acc[u + s];
  return warp_sum(acc[0]);
}

// Read-only form: sum over one row segment of A[r][c] * e[c]; e staged in shared memory.
// Fixed evaluation order (depends on the segment length only): lane l owns vectors
// l, l+32, ...; vector j of a batch goes to accumulator j; accumulators are folded pairwise,
// then the 32 lanes by an xor-shuffle tree.  All lanes return the sum.
template<int VEC, bool HINT = false, typename ACC = float>
__device__ __forceinline__ float
row_dot_readonly(const float* __restrict__ row, const float* e_s, uint32_t len, int lane,
                 unsigned long long pol = 0ull)
{
  using V = typename Vec<VEC>::type;
  const V* __restrict__ a = reinterpret_cast<const V*>(row);
  const V* es = reinterpret_cast<const V*>(e_s);
  const uint32_t nv = len / VEC;
  // loads in flight per lane: 8 x 16 bytes, or -- scalar units, dim % 4 != 0 -- 32 x 4 bytes: the same 4 KB per warp.
  // Unit j still goes to accumulator (j / 32) % 8 = u % 8, in ascending j, so the count does not touch the order.
  constexpr int LDN = kLoadsInFlight<VEC>();
  ACC acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = ACC(0);
  uint32_t i = (uint32_t)lane;
  for (; i + 32u * (LDN - 1) < nv; i += 32u * LDN) {
    V v[LDN];
#pragma unroll
    for (int u = 0; u < LDN; u++)
      v[u] = HINT ? ld_stream(a + i + 32u * u, pol) : ld_stream(a + i + 32u * u);
#pragma unroll
    for (int u = 0; u < LDN; u++)
      acc[u % kUnroll] = dot_acc(v[u], es[i + 32u * u], acc[u % kUnroll]);
  }
#pragma unroll
  for (int u = 0; u < LDN; u++) {
    const uint32_t j = i + 32u * u;
    if (j < nv)
      acc[u % kUnroll] = dot_acc(HINT ? ld_stream(a + j, pol) : ld_stream(a + j), es[j], acc[u % kUnroll]);
  }
#pragma unroll
  for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
    for (int u = 0; u < s; u++)
      acc[u] += acc[u + s];
  return (float)warp_sum(acc[0]);
}

__device__ __forceinline__ float4
rescale(float4 w, float inv_r, float4 sc)
{
  // W[r][c] *= (1.f / s[r]) * s[c]        reference similarity_transform.cpp:324-325
  w.x *= inv_r * sc.x;
  w.y *= inv_r * sc.y;
  w.z *= inv_r * sc.z;
  w.w *= inv_r * sc.w;
  return w;
}
__device__ __forceinline__ float
rescale(float w, float inv_r, float sc)
{
  return w * (inv_r * sc);
}
__device__ __forceinline__ float
sum_acc(float4 w, float acc)
{
  acc += w.x;
  acc += w.y;
  acc += w.z;
  acc += w.w;
  return acc;
}
__device__ __forceinline__ float
sum_acc(float w, float acc)
{
  return acc + w;
}

// In-place form: (first pass) copy A -> W and sum it, (later passes) rescale W in place by
// (1/s_prev[r]) * s_prev[c] and sum the new values.  Same lane/accumulator order as above.
template<int VEC, bool FIRST>
__device__ __forceinline__ float
row_pass_inplace(const float* __restrict__ src, float* dst, const float* sc_s, float inv_r,
                 uint32_t len, int lane)
{
  using V = typename Vec<VEC>::type;
  const V* a = reinterpret_cast<const V*>(src);
  V* w = reinterpret_cast<V*>(dst);
  const V* ss = reinterpret_cast<const V*>(sc_s);
  const uint32_t nv = len / VEC;
  constexpr int LDN = kLoadsInFlight<VEC>();
  float acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++)
    acc[u] = 0.f;
  uint32_t i = (uint32_t)lane;
  for (; i + 32u * (LDN - 1) < nv; i += 32u * LDN) {
    V v[LDN];
#pragma unroll
    for (int u = 0; u < LDN; u++)
      v[u] = FIRST ? ld_stream(a + i + 32u * u) : ld_cg(a + i + 32u * u);
#pragma unroll
    for (int u = 0; u < LDN; u++) {
      if (!FIRST)
        v[u] = rescale(v[u], inv_r, ss[i + 32u * u]);
      __stcg(w + i + 32u * u, v[u]);
      acc[u % kUnroll] = sum_acc(v[u], acc[u % kUnroll]);
    }
  }
#pragma unroll
  for (int u = 0; u < LDN; u++) {
    const uint32_t j = i + 32u * u;
    if (j < nv) {
      V v = FIRST ? ld_stream(a + j) : ld_cg(a + j);
      if (!FIRST)
        v = rescale(v, inv_r, ss[j]);
      __stcg(w + j, v);
      acc[u % kUnroll] = sum_acc(v, acc[u % kUnroll]);
    }
  }
#pragma unroll
  for (int s = kUnroll / 2; s >= 1; s >>= 1)
#pragma unroll
    for (int u = 0; u < s; u++)
      acc[u] += acc[u + s];
  return warp_sum(acc[0]);
}

// ---------------------------------------------------------------------------------------
// grid-wide (and, when sharded, cross-GPU) barrier at the end of round k -- which also carries the max
// ---------------------------------------------------------------------------------------
// Every CTA has written its row sums (locally and into every peer's S buffer) and knows the largest one it
// published (*s_cta_max: the bits of a non-negative, non-NaN float, so unsigned order = numeric order; kernels
// that compute the max in their own vector tail pass nullptr and ignore *s_m).
//   one GPU : atomicMax of ((k + 1) << 32 | bits) into smax[parity] -- the round tag makes the word monotonic, so it
//             is never reset -- then one release-add on the monotonic arrival counter; every CTA polls the counter.
//   sharded : a FLAT barrier over the CTAs of all GPUs.  Every CTA folds its max into every GPU's slot of the round
//             and then arrives at every GPU's counter itself (fire-and-forget reductions over NVLink, peers first);
//             ONE release/acquire system fence in between orders the CTA's peer stores of s (ordered before thread
//             0 by the __syncthreads) and the max before the arrivals.  A waiter polls one word in its own memory:
//             no completing CTA, no forwarding hop, no per-peer flag words.  The counters are never reset (a
//             solve starts from the total its group has reached); the max slots rotate by round: the slot of round
//             r is cleared by its GPU's CTA 0 once barrier r + 1 has been passed (every reader of it has arrived
//             there) and is next written in round r + 3 by CTAs that passed barrier r + 2 -- which CTA 0 only
//             arrives at after the clear.
//   result  : *s_m = max(0, max_r s[r]) over all GPUs = the reference's find_max (:154-227), so the vector
//             tail needs a single pass over s (stop test and eigenvector update together).
// S and E are indexed by the parity of k + flip.  All waits are bounded by timeout_ns (checked every 1024 polls)
// so that a missing rank turns into an error code instead of a hung GPU.  Returns false on timeout or when
// another CTA reported one.
// Measured against the protocol it replaced (this GPU's last CTA forwards one flag word per peer, every CTA polls
// G - 1 flags): 89.5 instead of 90.4 us per round at Hilbert 32768 on 8 GPUs, 84.7 instead of 85.8 at Hilbert
// 16384 on 2 (profiles/r2_c21_ab_barrier_8gpu.json, r2_c20_ab_barrier_2gpu.json).

struct SpinClock
{
  unsigned long long t0 = 0ull;
  unsigned int spins = 0u;
  // true once the wait has lasted longer than limit_ns, or another CTA has given up
  __device__ __forceinline__ bool expired(unsigned long long limit_ns, const unsigned int* error_word)
  {
    if ((++spins & 1023u) != 0u)
      return false;
    if (ld_relaxed_gpu(error_word) != 0u)
      return true;
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0ull) {
      t0 = now;
      return false;
    }
    return now - t0 > limit_ns;
  }
};

// Head of a rank's exchange block (mapped into every peer), each word group on a line of its own; the two
// row-sum buffers follow.
struct alignas(128) ExchangeHeader
{
  unsigned long long arrive; // + kArriveUnits per GPU and round, never reset
  unsigned long long pad0[15];
  unsigned int smax3[3];     // slot r % 3 holds round r's max (float bits); cleared two rounds ahead of its next use
  unsigned int pad1[29];
};
// A GPU's CTAs add up to exactly this much per round whatever its grid size is (ranks of tiny problems can run
// different grids), so no rank needs to know another rank's launch shape.
constexpr unsigned long long kArriveUnits = 4096ull;

__device__ __forceinline__ bool
round_barrier(const RoundParams& p, uint32_t k, volatile int* s_abort, unsigned int* s_cta_max = nullptr,
              volatile float* s_m = nullptr)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    int fail = 0;
    // the CTA's warps have folded their maxima into *s_cta_max (shared memory) before the __syncthreads; it is
    // read and cleared here, and written again only after the closing __syncthreads (next round's pass)
    uint32_t cta_max_bits = 0u;
    if (s_cta_max) {
      cta_max_bits = *s_cta_max;
      *s_cta_max = 0u;
    }
    SpinClock clk;
    uint32_t bits;
    if (p.world == 1) {
      const uint32_t par = (k + p.flip) & 1u;
      const unsigned int target = (k + 1u) * gridDim.x;
      atomicMax(&p.bar->smax[par], ((unsigned long long)(k + 1u) << 32) | cta_max_bits);
      red_release_gpu_add(&p.bar->count, 1u);
      while (ld_acquire_gpu(&p.bar->count) < target) {
        if (clk.expired(p.timeout_ns, &p.bar->error)) {
          fail = 1;
          break;
        }
      }
      bits = (uint32_t)(ld_relaxed_gpu(&p.bar->smax[par]) & 0xffffffffull);
    } else {
      const unsigned long long r = p.round_base + k;
      const uint32_t slot = (uint32_t)(r % 3ull);
      if (cta_max_bits != 0u) // the slots start from 0 = the reference's zero-filled max cell (:169)
        for (uint32_t i = 1; i <= p.world; i++) {
          const uint32_t g = p.rank + i < p.world ? p.rank + i : p.rank + i - p.world;
          red_relaxed_sys_max(p.peer_smax3[g] + slot, cta_max_bits);
        }
      fence_acq_rel_sys();
      const unsigned long long inc =
        kArriveUnits * (blockIdx.x + 1u) / gridDim.x - kArriveUnits * blockIdx.x / gridDim.x;
      for (uint32_t i = 1; i <= p.world; i++) {
        const uint32_t g = p.rank + i < p.world ? p.rank + i : p.rank + i - p.world;
        red_relaxed_sys_add(p.peer_arrive[g], inc);
      }
      const unsigned long long want = p.arrive_base + (unsigned long long)(k + 1u) * p.world * kArriveUnits;
      while (ld_acquire_sys(p.peer_arrive[p.rank]) < want) {
        if (clk.expired(p.timeout_ns, &p.bar->error)) {
          fail = 1;
          break;
        }
      }
      bits = ld_relaxed_sys(p.peer_smax3[p.rank] + slot);
      if (blockIdx.x == 0)
        st_relaxed_sys(p.peer_smax3[p.rank] + (slot + 2u) % 3u, 0u); // round r - 1's slot, next used in round r + 2
    }
    if (fail)
      atomicExch(&p.bar->error, 1u);
    if (s_m)
      *s_m = __uint_as_float(bits);
    *s_abort = fail;
  }
  __syncthreads();
  return *s_abort == 0;
}

// ---------------------------------------------------------------------------------------
// the round loop
// ---------------------------------------------------------------------------------------
template<int VEC, int FORM, int MAX_THREADS, int STOP = kStopAbsolute, typename T = float, typename ACC = float>
__global__ void __launch_bounds__(MAX_THREADS, 1) round_loop_kernel(const RoundParams p)
{
  static_assert(sizeof(ACC) == 4 || (sizeof(T) == 4 && FORM == kFormReadOnly), "fp64 accumulation: fp32 storage, read-only form");
  constexpr bool kBf16 = sizeof(T) == 2; // p.A then points to bf16 storage (read-only form, N % 4 == 0)
  static_assert(!kBf16 || (VEC == 4 && FORM == kFormReadOnly), "bf16 storage: read-only form, vector loads");
  constexpr bool kFp8 = sizeof(T) == 1; // p.A points to e4m3 storage, p.row_scale to the row scales (read-only form, N % 4 == 0)
  static_assert(!kFp8 || (VEC == 4 && FORM == kFormReadOnly && sizeof(ACC) == 4), "fp8 storage: read-only form, vector loads, fp32 accumulation");
  const uint32_t THREADS = blockDim.x; // run-time CTA size (multiple of 32, <= MAX_THREADS)
  extern __shared__ __align__(16) float smem[];
  float* scale_s = smem;               // chunk_cols floats: e (read-only) or s_prev (in-place)
  float* part_s = smem + p.chunk_cols; // one partial row sum per owned row
  __shared__ float red_max[32];
  __shared__ int red_ok[32];
  __shared__ float bc_max;
  __shared__ int bc_ok;
  __shared__ int s_abort;

  const uint32_t tid = threadIdx.x;
  const int lane = (int)(tid & 31u);
  const uint32_t warp = tid >> 5;
  const uint32_t kWarps = THREADS >> 5;
  const uint32_t N = p.N;

  // rows of this GPU's block owned by this CTA, and the slice of the N-vector it maintains
  const uint32_t rb = (uint32_t)((uint64_t)p.rows * blockIdx.x / gridDim.x);
  const uint32_t re = (uint32_t)((uint64_t)p.rows * (blockIdx.x + 1) / gridDim.x);
  const uint32_t nrows = re - rb;
  const uint32_t cb = (uint32_t)((uint64_t)N * blockIdx.x / gridDim.x);
  const uint32_t ce = (uint32_t)((uint64_t)N * (blockIdx.x + 1) / gridDim.x);

  if (blockIdx.x == 0 && tid == 0)
    p.round_ts[0] = globaltimer_ns();
  const unsigned long long pol_keep = l2_policy_evict_last();
  const unsigned long long pol_stream = l2_policy_evict_first();

  float m_prev = 1.f;
  for (uint32_t k = 0;; ++k) {
    const uint32_t par = (k + p.flip) & 1u; // buffer set of this round (sharded: see RoundParams::flip)
    const float* Sprev = p.S[par ^ 1u];
    const float* Eprev = p.E[par ^ 1u];
    float* Scur = p.S[par];
    float* Ecur = p.E[par];
    const bool first = (k == 0);

    // e_k = e_{k-1} * (s_{k-1} / m_{k-1})          reference :42-43 -> :260; e_0 = 1 (:34 -> :280)
    for (uint32_t c = cb + tid; c < ce; c += THREADS)
      Ecur[c] = first ? 1.f : ld_cg(Eprev + c) * (ld_cg(Sprev + c) / m_prev);
    for (uint32_t r = tid; r < nrows; r += THREADS)
      part_s[r] = 0.f;

    // ---- the pass over the matrix ----                                   reference :40 (+ :52)
    const bool backward = p.sweep && (k & 1);
    for (uint32_t c0 = 0; c0 < N; c0 += p.chunk_cols) {
      const uint32_t clen = min(p.chunk_cols, N - c0);
      __syncthreads();
      // every CTA rebuilds the same chunk of the scale vector at the same time: each starts at
      // its own offset (de-phased on the L2 lines of E and S); 128-bit loads, four batches in
      // flight before anything depends on them
      if (VEC == 4) {
        const uint32_t cv = clen >> 2;
        const uint32_t rotv = (cb >> 2) % cv;
        const float4* S4 = reinterpret_cast<const float4*>(Sprev + c0);
        const float4* E4 = reinterpret_cast<const float4*>(Eprev + c0);
        float4* sc4 = reinterpret_cast<float4*>(scale_s);
        for (uint32_t v0 = 0; v0 < cv; v0 += 4u * THREADS) {
          float4 sp[4], ep[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
            const uint32_t vi = idx + rotv < cv ? idx + rotv : idx + rotv - cv;
            const bool active = idx < cv && !first;
            sp[j] = active ? ld_cg(S4 + vi) : make_float4(1.f, 1.f, 1.f, 1.f);
            ep[j] = (active && FORM == kFormReadOnly) ? ld_cg(E4 + vi) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
            if (idx < cv) {
              const uint32_t vi = idx + rotv < cv ? idx + rotv : idx + rotv - cv;
              float4 v = sp[j];
              if (first) {
                v = make_float4(1.f, 1.f, 1.f, 1.f);
              } else if (FORM == kFormReadOnly) {
                v.x = ep[j].x * (sp[j].x / m_prev);
                v.y = ep[j].y * (sp[j].y / m_prev);
                v.z = ep[j].z * (sp[j].z / m_prev);
                v.w = ep[j].w * (sp[j].w / m_prev);
              }
              if (kFp8) { // fp8 storage multiplies with e * 2^120 (dot_acc_fp8)
                v.x *= kFp8EigenScale;
                v.y *= kFp8EigenScale;
                v.z *= kFp8EigenScale;
                v.w *= kFp8EigenScale;
              }
              sc4[vi] = v;
            }
          }
        }
      } else {
        const uint32_t rot = cb % clen;
        for (uint32_t cc = tid; cc < clen; cc += THREADS) {
          const uint32_t c = cc + rot < clen ? cc + rot : cc + rot - clen;
          float v = 1.f;
          if (!first) {
            const float sp = ld_cg(Sprev + c0 + c);
            v = (FORM == kFormReadOnly) ? ld_cg(Eprev + c0 + c) * (sp / m_prev) : sp;
          }
          scale_s[c] = v;
        }
      }
      __syncthreads();
      // the staged window holds up to four 8192-column chunks: a warp streams its row through all of them (128 KB
      // of one row in a piece) -- one chunk sum at a time, added left to right, exactly as if each chunk had been staged
      // on its own; only the number of rebuilds and CTA-wide synchronisations per round drops (N / 32768 instead of N / 8192)
      for (uint32_t i = warp; i < nrows; i += kWarps) {
        const uint32_t rl = backward ? (nrows - 1u - i) : i;
        float inv_r = 1.f;
        if (FORM == kFormInPlace && !first)
          inv_r = 1.f / ld_cg(Sprev + p.row0 + rb + rl);
        for (uint32_t s0 = 0; s0 < clen; s0 += (uint32_t)kChunkCols) {
          const uint32_t slen = min((uint32_t)kChunkCols, clen - s0);
          const size_t off = (size_t)(rb + rl) * N + c0 + s0;
          const float* sc = scale_s + s0;
          float t;
          if (kFp8) {
            const uint32_t* seg = reinterpret_cast<const uint32_t*>(reinterpret_cast<const fp8_t*>(p.A) + off);
            t = row_dot_fp8<4 * kUnroll>(seg, reinterpret_cast<const float4*>(sc), slen >> 2, lane);
          } else if (kBf16) {
            const uint2* seg = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16_t*>(p.A) + off);
            t = row_dot_bf16<2 * kUnroll>(seg, reinterpret_cast<const float4*>(sc), slen >> 2, lane);
          } else if (FORM == kFormReadOnly) {
            if (sizeof(ACC) == 8)
              t = row_dot_readonly<VEC, false, ACC>(p.A + off, sc, slen, lane);
            else if (p.keep_rows_pct == 0u)
              t = row_dot_readonly<VEC>(p.A + off, sc, slen, lane);
            else
              t = row_dot_readonly<VEC, true>(p.A + off, sc, slen, lane,
                                              rl * 100u < nrows * p.keep_rows_pct ? pol_keep : pol_stream);
          } else if (first) {
            t = row_pass_inplace<VEC, true>(p.A + off, p.W + off, sc, 1.f, slen, lane);
          } else {
            t = row_pass_inplace<VEC, false>(p.W + off, p.W + off, sc, inv_r, slen, lane);
          }
          if (lane == 0)
            part_s[rl] += t;
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 0u);
    __syncthreads();

    // ---- publish this CTA's row sums (to every rank when sharded) ----
    for (uint32_t r = tid; r < nrows; r += THREADS) {
      const uint32_t gr = p.row0 + rb + r;
      float s = part_s[r];
      if (kFp8)
        s = s * p.row_scale[rb + r];
      if (FORM == kFormReadOnly && !first)
        s = s / (ld_cg(Eprev + gr) * (ld_cg(Sprev + gr) / m_prev));
      if (p.world > 1) {
        for (uint32_t g = 0; g < p.world; g++)
          __stcg(p.peer_S[par][g] + gr, s);
      } else {
        __stcg(Scur + gr, s);
      }
    }

    if (!round_barrier(p, k, &s_abort))
      return;
    if (blockIdx.x == 0 && tid == 0)
      stamp_phase(p, k, 1u);

    // ---- every CTA: max, circular stop test over the full vector ----   reference :41, :44
    float mx = 0.f; // reference zero-fills the max cell (:169)
    int ok = 1;
    uint32_t dbits = 0u; // relative stop test only: bits of the largest adjacent difference
    // every CTA scans the same vector at the same time: start each at its own offset (cb) so
    // they do not queue on the same L2 lines; max / AND are order-independent.  The circular
    // neighbour comes from the next lane by shuffle, as in the reference (:413-417); four
    // batches of loads are in flight before anything depends on them.
    if (VEC == 4) {
      const uint32_t nv = N >> 2;
      const uint32_t rotv = cb >> 2;
      const float4* S4 = reinterpret_cast<const float4*>(Scur);
      for (uint32_t v0 = 0; v0 < nv; v0 += 4u * THREADS) {
        float4 t4[4];
        float nx[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
          const bool active = idx < nv;
          const uint32_t vi = idx + rotv < nv ? idx + rotv : idx + rotv - nv;
          t4[j] = active ? ld_cg(S4 + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
          const bool edge = active && (lane == 31 || idx + 1u >= nv);
          nx[j] = edge ? ld_cg(Scur + (vi + 1u == nv ? 0u : 4u * (vi + 1u))) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t idx = v0 + (uint32_t)j * THREADS + tid;
          if (v0 + (uint32_t)j * THREADS < nv) { // warp-uniform
            float nxt = __shfl_down_sync(0xffffffffu, t4[j].x, 1);
            if (lane == 31 || idx + 1u >= nv)
              nxt = nx[j];
            if (idx < nv) {
              const float4 v = t4[j];
              mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
              if (STOP == kStopRelative) {
                dbits = max(max(dbits, diff_bits(v.x, v.y)),
                            max(max(diff_bits(v.y, v.z), diff_bits(v.z, v.w)), diff_bits(v.w, nxt)));
              } else {
                // strict <, wrap pair included (:413-421)
                ok &= (fabsf(v.x - v.y) < p.eps) & (fabsf(v.y - v.z) < p.eps) & (fabsf(v.z - v.w) < p.eps) &
                      (fabsf(v.w - nxt) < p.eps);
              }
            }
          }
        }
      }
    } else {
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
              if (STOP == kStopRelative)
                dbits = max(dbits, diff_bits(sf[j], next));
              else
                ok &= (fabsf(sf[j] - next) < p.eps) ? 1 : 0; // strict <, wrap pair included (:413-421)
            }
          }
        }
      }
    }
    if (STOP == kStopRelative)
      ok = (int)dbits; // from here on `ok` carries the difference bits; max instead of AND
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

    if (blockIdx.x == 0 && tid == 0)
      stamp_round_end(p, k);

    if (converged || k + 1u == p.max_iter) {
      // the eigenvector update of this round still happens before the break (:42-50)
      for (uint32_t c = cb + tid; c < ce; c += THREADS)
        p.out_eigen_vec[c] = ld_cg(Ecur + c) * (ld_cg(Scur + c) / m_k);
      if (blockIdx.x == 0 && tid == 0) {
        *p.out_eigen_val = ld_cg(Scur);                  // :60-65
        p.out_iter[0] = converged ? k : p.max_iter;      // :54
        p.out_iter[1] = k + 1u;
      }
      return;
    }
    m_prev = m_k;
  }
}

// ---------------------------------------------------------------------------------------
// per-kernel entry points (reference L1 functions) -- small, not on the fused path
// ---------------------------------------------------------------------------------------

// sum_across_rows()  reference similarity_transform.cpp:77-152: one warp per row, same
// evaluation order as the fused kernel.  With e == nullptr it is the plain row sum
// (vec[r] = sum_c mat[r][c]); with e it is one read-only round's row pass for rows
// [row0, row0+rows): vec[row0+r] = (sum_c mat[r][c] * e[c]) / e[row0+r]  -- the unfused
// building block of the collective (NCCL) variant of the sharded loop.
template<int VEC>
__global__ void __launch_bounds__(256) sum_across_rows_kernel(const float* __restrict__ mat,
                                                              const float* __restrict__ e,
                                                              float* __restrict__ vec, uint32_t dim,
                                                              uint32_t row0, uint32_t rows)
{
  extern __shared__ __align__(16) float scale_s[];
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t chunk = min((uint32_t)kChunkCols, dim);
  const bool single = chunk == dim;
  for (uint32_t c0 = 0; c0 < dim; c0 += chunk) {
    const uint32_t clen = min(chunk, dim - c0);
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < clen; c += blockDim.x)
      scale_s[c] = e ? e[c0 + c] : 1.f;
    __syncthreads();
    for (uint32_t r = gw; r < rows; r += warps) {
      const float t = row_dot_readonly<VEC>(mat + (size_t)r * dim + c0, scale_s, clen, lane);
      if (lane == 0) {
        const float acc = (c0 == 0 ? 0.f : vec[row0 + r]) + t;
        const bool last = single || c0 + clen == dim;
        vec[row0 + r] = (last && e) ? acc / e[row0 + r] : acc;
      }
    }
  }
}

// find_max()  reference :154-227: m = max(0, max_r s[r]).  The host zero-fills the cell first, like
// the reference does (:162-170); blocks combine with an integer atomicMax on the bit pattern, which
// orders non-negative floats correctly (the running max starts at +0 and fmaxf drops NaNs).
__global__ void __launch_bounds__(1024) find_max_kernel(const float* __restrict__ vec,
                                                        float* __restrict__ out, uint32_t dim)
{
  __shared__ float red[32];
  float mx = 0.f;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x)
    mx = fmaxf(mx, vec[c]);
  for (int o = 16; o >= 1; o >>= 1)
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o >= 1; o >>= 1)
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0)
      atomicMax(reinterpret_cast<int*>(out), __float_as_int(mx));
  }
}

// compute_eigen_vector()  reference :229-265: e[r] *= s[r] / m
__global__ void
compute_eigen_vector_kernel(const float* __restrict__ vec, const float* __restrict__ mx,
                            float* __restrict__ eigen_vec, uint32_t dim)
{
  const float m = *mx;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += gridDim.x * blockDim.x)
    eigen_vec[r] *= (vec[r] / m);
}

// initialise_eigen_vector()  reference :267-284
__global__ void
fill_kernel(float* __restrict__ v, float value, uint32_t dim)
{
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += gridDim.x * blockDim.x)
    v[r] = value;
}

// stop()  reference :332-460: ret = 1 iff every circular adjacent pair differs by < eps.  The host
// fills the flag with 1 first, like the reference does (:351-359); a block that sees a failing
// pair clears it (the reference combines with an atomic min, :438-446).
__global__ void __launch_bounds__(1024) stop_kernel(const float* __restrict__ vec,
                                                    uint32_t* __restrict__ ret, uint32_t dim, float eps)
{
  int ok = 1;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x) {
    const float self = vec[c];
    const float next = vec[c + 1u == dim ? 0u : c + 1u];
    ok &= (fabsf(self - next) < eps) ? 1 : 0;
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0 && !ok)
    atomicAnd(ret, 0u);
}

__global__ void
fill_u32_kernel(uint32_t* __restrict__ v, uint32_t value, uint32_t n)
{
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    v[r] = value;
}

// Vector tail of one HOST-DRIVEN round (the streamed solve, Context::solve_streamed: the matrix does
// not fit the device, so the round loop cannot live in one kernel).  Two small grid-wide kernels do what
// the fused kernels' tail does, with the same order-independent reductions:
//   tail_scan_kernel    cells[0] <- bits of m = max(0, max_r s[r])                    reference :154-227
//                       cells[1] <- absolute stop: 1 iff every circular pair differs by < eps (:413-421)
//                                   relative stop: bits of the largest circular adjacent difference
//   tail_update_kernel  e[r] *= s[r] / m (:260); block 0 publishes out[0] = converged, out[1] = bits of s[0]
// The host presets cells[0] = 0 and cells[1] = (absolute ? 1 : 0) before the scan.  Non-negative floats
// (and the NaN of a difference) order like their bit patterns, so integer atomicMax combines the blocks.
__global__ void __launch_bounds__(1024) tail_scan_kernel(const float* __restrict__ vec, uint32_t dim, float eps,
                                                         int stop_kind, uint32_t* __restrict__ cells)
{
  __shared__ float red_max[32];
  __shared__ uint32_t red_bits[32];
  float mx = 0.f;
  int ok = 1;
  uint32_t dbits = 0u;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x) {
    const float self = vec[c];
    const float next = vec[c + 1u == dim ? 0u : c + 1u];
    mx = fmaxf(mx, self);
    if (stop_kind == kStopRelative)
      dbits = max(dbits, diff_bits(self, next));
    else
      ok &= (fabsf(self - next) < eps) ? 1 : 0;
  }
  uint32_t word = stop_kind == kStopRelative ? dbits : (uint32_t)(ok ? 0 : 1); // max-combinable: 1 = a pair failed
  for (int o = 16; o >= 1; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    word = max(word, (uint32_t)__shfl_xor_sync(0xffffffffu, word, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red_max[threadIdx.x >> 5] = mx;
    red_bits[threadIdx.x >> 5] = word;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool live = threadIdx.x < (blockDim.x >> 5);
    mx = live ? red_max[threadIdx.x] : 0.f;
    word = live ? red_bits[threadIdx.x] : 0u;
    for (int o = 16; o >= 1; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      word = max(word, (uint32_t)__shfl_xor_sync(0xffffffffu, word, o));
    }
    if (threadIdx.x == 0) {
      atomicMax(reinterpret_cast<int*>(cells), __float_as_int(mx));
      if (stop_kind == kStopRelative)
        atomicMax(reinterpret_cast<int*>(cells + 1), (int)word);
      else if (word)
        atomicAnd(cells + 1, 0u);
    }
  }
}

__global__ void
tail_update_kernel(const float* __restrict__ vec, float* __restrict__ eigen_vec, uint32_t dim, float eps,
                   int stop_kind, const uint32_t* __restrict__ cells, uint32_t* __restrict__ out)
{
  const float m = __uint_as_float(cells[0]);
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < dim; r += gridDim.x * blockDim.x)
    eigen_vec[r] *= (vec[r] / m);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = stop_kind == kStopRelative ? (__uint_as_float(cells[1]) < eps * m ? 1u : 0u) : cells[1];
    out[1] = __float_as_uint(vec[0]);
  }
}

// compute_next_matrix()  reference :286-330: W[r][c] *= (1.f / s[r]) * s[c]
template<int VEC>
__global__ void __launch_bounds__(256) compute_next_matrix_kernel(float* __restrict__ mat,
                                                                  const float* __restrict__ vec,
                                                                  uint32_t dim)
{
  using V = typename Vec<VEC>::type;
  const uint32_t nv = dim / VEC;
  for (uint32_t r = blockIdx.y; r < dim; r += gridDim.y) {
    const float inv_r = 1.f / vec[r];
    V* row = reinterpret_cast<V*>(mat + (size_t)r * dim);
    const V* sc = reinterpret_cast<const V*>(vec);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nv; j += gridDim.x * blockDim.x)
      row[j] = rescale(row[j], inv_r, sc[j]);
  }
}

// ---------------------------------------------------------------------------------------
// input generation
// ---------------------------------------------------------------------------------------

// generate_hilbert_matrix()  reference utils.cpp:136-154: A[r][c] = 1.f / (float)(r + c + 1)
// (IEEE-rounded division; r + c + 1 < 2^24 up to N = 2^23 so the cast is exact.)
__global__ void __launch_bounds__(256) hilbert_kernel(float* __restrict__ out, uint32_t dim,
                                                      uint32_t row0, uint32_t rows)
{
  for (uint32_t i = blockIdx.y; i < rows; i += gridDim.y) {
    float* row = out + (size_t)i * dim;
    const uint32_t base = row0 + i + 1u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x)
      row[c] = __fdiv_rn(1.f, (float)(base + c));
  }
}


__global__ void __launch_bounds__(256) convert_bf16_kernel(const float* __restrict__ src,
                                                           unsigned short* __restrict__ dst, size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7u) == 0);
  size_t done = 0;
  if (vec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = ld_stream(reinterpret_cast<const float4*>(src) + i);
      uint2 o;
      o.x = (uint32_t)f32_to_bf16_rn(v.x) | ((uint32_t)f32_to_bf16_rn(v.y) << 16);
      o.y = (uint32_t)f32_to_bf16_rn(v.z) | ((uint32_t)f32_to_bf16_rn(v.w) << 16);
      reinterpret_cast<uint2*>(dst)[i] = o;
    }
    done = n4 << 2;
  }
  for (size_t i = done + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = f32_to_bf16_rn(src[i]);
}

// fp32 -> fp8 storage, one warp per row (dim % 4 == 0, 16-byte aligned source rows): the row's largest magnitude a gives the
// scale 2^k with a / 2^k in (224, 448] (k clamped to +-118; an all-zero or non-finite row keeps scale 1), every element
// is divided by it -- exact -- and rounded to the nearest e4m3 code, ties to even, saturating.
__device__ __forceinline__ float
fp8_row_scale(float amax)
{
  if (!(amax > 0.f) || !(amax < __uint_as_float(0x7f800000u)))
    return 1.f;
  int e;
  const float m = frexpf(amax, &e); // amax = m * 2^e, m in [0.5, 1)
  int k = (m <= 0.875f) ? e - 9 : e - 8;
  k = max(-118, min(118, k));
  return ldexpf(1.f, k);
}

__global__ void __launch_bounds__(256) convert_fp8_rows_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst,
                                                               float* __restrict__ row_scale, uint32_t rows, uint32_t dim)
{
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nv = dim >> 2;
  for (uint32_t r = gw; r < rows; r += warps) {
    const float4* row = reinterpret_cast<const float4*>(src + (size_t)r * dim);
    float amax = 0.f;
    float has_nan = 0.f;
    for (uint32_t i = lane; i < nv; i += 32u) {
      const float4 v = row[i];
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      if ((v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w))
        has_nan = 1.f;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      has_nan = fmaxf(has_nan, __shfl_xor_sync(0xffffffffu, has_nan, o));
    }
    const float scale = fp8_row_scale(amax);
    const float inv = 1.f / scale; // a power of two: exact
    if (lane == 0)
      row_scale[r] = has_nan != 0.f ? __uint_as_float(0x7fc00000u) : scale; // the codes cannot carry a NaN (fp8_raw): the scale does
    uint32_t* out = reinterpret_cast<uint32_t*>(dst + (size_t)r * dim);
    for (uint32_t i = lane; i < nv; i += 32u) {
      const float4 v = row[i];
      out[i] = (uint32_t)f32x2_to_fp8x2(v.x * inv, v.y * inv) | ((uint32_t)f32x2_to_fp8x2(v.z * inv, v.w * inv) << 16);
    }
  }
}

// Philox-4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): counter (ctr, 0), key = seed.
__device__ __forceinline__ uint4
philox4x32_10(unsigned long long ctr, unsigned long long key)
{
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float
u01_open_closed(uint32_t x)
{
  return (float)((x >> 8) + 1u) * 5.9604644775390625e-08f; // (0, 1]
}

// Seeded uniform (0,1] fill; element idx = r*dim + c takes word idx%4 of Philox block idx/4,
// so the values do not depend on how the rows are sharded.
__global__ void __launch_bounds__(256) uniform_kernel(float* __restrict__ out, uint32_t dim,
                                                      uint32_t row0, uint32_t rows,
                                                      unsigned long long seed)
{
  const unsigned long long first = (unsigned long long)row0 * dim;
  const unsigned long long last = first + (unsigned long long)rows * dim; // exclusive
  const unsigned long long blk0 = first >> 2;
  const unsigned long long nblk = ((last + 3ull) >> 2) - blk0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk;
       b += stride) {
    const unsigned long long blk = blk0 + b;
    const uint4 w = philox4x32_10(blk, seed);
    const unsigned long long base = blk << 2;
    const float v[4] = { u01_open_closed(w.x), u01_open_closed(w.y), u01_open_closed(w.z),
                         u01_open_closed(w.w) };
    if (base >= first && base + 3ull < last && (((base - first) & 3ull) == 0ull)) {
      *reinterpret_cast<float4*>(out + (base - first)) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned long long idx = base + j;
        if (idx >= first && idx < last)
          out[idx - first] = v[j];
      }
    }
  }
}

} // namespace st
