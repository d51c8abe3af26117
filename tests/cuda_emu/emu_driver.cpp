// emu_driver.cpp -- runs the round kernels of eigen_value_b200/csrc on the CPU emulation harness
// (TEST INFRASTRUCTURE; see cuda_emu.h).  Mirrors the part of st::Context::solve that fills
// RoundParams and picks the dynamic shared-memory layout (the layout arithmetic itself is shared:
// launch_plan.hpp), with explicit, small launch shapes, and can run several emulated GPUs at once
// with their exchange blocks wired to each other, like st_shard_link_local does.
//
// Built by tests/cuda_emu/build.py from rewritten copies of the kernel sources (only the
// `extern __shared__` declarations are rewritten).
#include "cuda_emu.h"

#include "kernels.cuh"
#include "kernels_sc.cuh"
#include "kernels_wide.cuh"
#include "kernels_cluster.cuh"
#include "launch_plan.hpp"

#include <string>

using namespace st;

extern "C" {

struct emu_opts
{
  float eps;
  uint32_t max_iter;
  int32_t form;    // 0 read-only, 1 in-place
  int32_t sweep;   // bit 0: alternate the row order
  int32_t dynamic; // -1: automatic (dim >= 8192), 0 static, 1 dynamic work units
  int32_t threads; // CTA size
  int32_t ctas;    // grid size per emulated GPU
  int32_t kernel;  // 1 general, 2 wide, 10-13 resident-e, 20 cluster
  int32_t stop;    // 0 absolute, 1 relative
  int32_t bf16;    // 1: matrix is bf16 storage; 2: fp8 (e4m3) storage with row scales
  int32_t world;   // emulated GPUs (row-block sharded)
  int32_t acc64;   // fp64 accumulation (read-only form, fp32 storage; kernels 1, 10, 12, 13)
  const float* row_scale; // fp8 storage: one power-of-two scale per row of the whole matrix
};

static thread_local std::string g_err;
const char*
emu_last_error()
{
  return g_err.c_str();
}

} // extern "C"

namespace {

struct Rank
{
  RoundParams p{};
  std::vector<float> vecs;    // S0 S1 (world == 1), E0 E1, OUT
  std::vector<float> partial; // chunk sums
  std::vector<unsigned> row_done;
  std::vector<unsigned> phase_counter;
  std::vector<unsigned long long> ts;
  std::vector<unsigned char> work; // in-place working copy
  BarrierState* bar = nullptr;
  unsigned char* xblock = nullptr; // exchange block: ExchangeHeader + 2 x N floats
  float scalars[16] = {};
  unsigned grid = 1, threads = 32;
  size_t smem = 0;
  std::unique_ptr<emu::Grid> launch;
};

template<int STOP>
void (*general_kernel(bool vec4, int form, bool bf16, bool acc64, bool fp8 = false))(const RoundParams)
{
  if (acc64)
    return vec4 ? round_loop_kernel<4, kFormReadOnly, 512, STOP, float, double>
                : round_loop_kernel<1, kFormReadOnly, 512, STOP, float, double>;
  if (fp8)
    return round_loop_kernel<4, kFormReadOnly, 512, STOP, fp8_t>;
  if (bf16)
    return round_loop_kernel<4, kFormReadOnly, 512, STOP, bf16_t>;
  if (vec4)
    return form == kFormInPlace ? round_loop_kernel<4, kFormInPlace, 512, STOP, float>
                                : round_loop_kernel<4, kFormReadOnly, 512, STOP, float>;
  return form == kFormInPlace ? round_loop_kernel<1, kFormInPlace, 512, STOP, float>
                              : round_loop_kernel<1, kFormReadOnly, 512, STOP, float>;
}

template<int STOP>
void (*sc_kernel(int pf, bool bf16, bool acc64, bool fp8 = false))(const RoundParams)
{
  if (acc64) {
    switch (pf) {
      case 1: return round_loop_sc_kernel<512, 1, STOP, float, double>;
      case 2: return round_loop_sc_kernel<512, 2, STOP, float, double>;
      default: return round_loop_sc_kernel<512, 3, STOP, float, double>;
    }
  }
  if (fp8)
    return round_loop_sc_kernel<512, 0, STOP, fp8_t>;
  if (bf16)
    return round_loop_sc_kernel<512, 0, STOP, bf16_t>;
  switch (pf) {
    case 0: return round_loop_sc_kernel<512, 0, STOP, float>;
    case 1: return round_loop_sc_kernel<512, 1, STOP, float>;
    case 2: return round_loop_sc_kernel<512, 2, STOP, float>;
    default: return round_loop_sc_kernel<512, 3, STOP, float>;
  }
}

} // namespace

extern "C" int
emu_solve(const void* mat, uint32_t dim, const emu_opts* o, float* eigen_val, float* eigen_vec,
          uint32_t* iter_count, uint32_t* passes, uint32_t* ranks_agree)
{
  try {
    const uint32_t world = (uint32_t)std::max(1, o->world);
    if (!mat || dim == 0 || world > (uint32_t)kMaxWorld || world > dim || o->max_iter == 0)
      throw std::string("bad argument");
    const int form = o->form ? kFormInPlace : kFormReadOnly;
    const bool fp8 = o->bf16 == 2;
    const bool bf16 = o->bf16 == 1;
    const bool acc64 = o->acc64 != 0;
    const bool vec4 = dim % 4u == 0u;
    if (acc64 && (bf16 || form != kFormReadOnly || !(o->kernel == 1 || o->kernel == 10 || o->kernel == 12 || o->kernel == 13)))
      throw std::string("fp64 accumulation: fp32 storage, read-only form, kernels 1, 10, 12, 13");
    if (bf16 && (dim % 4u != 0u || form != kFormReadOnly))
      throw std::string("bf16 storage needs dim % 4 == 0 and the read-only form");
    if (fp8 && (dim % 4u != 0u || form != kFormReadOnly || acc64 || !o->row_scale || !(o->kernel == 1 || o->kernel == 11)))
      throw std::string("fp8 storage needs dim % 4 == 0, the read-only form, row scales, kernel 1 or 11");
    const size_t elem = fp8 ? 1 : bf16 ? 2 : 4;
    const size_t nvec = ((size_t)dim + 31) & ~(size_t)31;

    std::vector<Rank> R(world);
    for (uint32_t g = 0; g < world; g++) {
      Rank& r = R[g];
      r.xblock = static_cast<unsigned char*>(aligned_alloc(128, sizeof(ExchangeHeader) + 2 * nvec * sizeof(float)));
      memset(r.xblock, 0, sizeof(ExchangeHeader) + 2 * nvec * sizeof(float));
      reinterpret_cast<ExchangeHeader*>(r.xblock)->arrive = 3ull * world * kArriveUnits; // == arrive_base below
      r.bar = static_cast<BarrierState*>(aligned_alloc(128, sizeof(BarrierState)));
      memset(r.bar, 0, sizeof(BarrierState));
    }
    for (uint32_t g = 0; g < world; g++) {
      Rank& r = R[g];
      RoundParams& p = r.p;
      const uint32_t row0 = (uint32_t)((uint64_t)dim * g / world);
      const uint32_t rows = (uint32_t)((uint64_t)dim * (g + 1) / world) - row0;
      p.A = reinterpret_cast<const float*>(static_cast<const unsigned char*>(mat) + (size_t)row0 * dim * elem);
      p.N = dim;
      p.row0 = row0;
      p.rows = rows;
      p.row_scale = fp8 ? o->row_scale + row0 : nullptr;
      r.vecs.assign(5 * nvec, -7.f);
      p.S[0] = r.vecs.data();
      p.S[1] = r.vecs.data() + nvec;
      p.E[0] = r.vecs.data() + 2 * nvec;
      p.E[1] = r.vecs.data() + 3 * nvec;
      p.out_eigen_vec = r.vecs.data() + 4 * nvec;
      p.eps = o->eps;
      p.max_iter = o->max_iter;
      p.sweep = o->sweep & 1;
      p.dynamic = o->dynamic < 0 ? (dim >= (uint32_t)kChunkCols ? 1 : 0) : (o->dynamic ? 1 : 0);
      p.keep_rows_pct = 0;
      p.chunk_cols = std::min<uint32_t>((uint32_t)kWindowCols, dim);
      if (const char* w = getenv("ST_EMU_WINDOW")) // tests: a smaller staged window, so that several windows per row fit an emulated size
        p.chunk_cols = std::min<uint32_t>((uint32_t)std::max(kChunkCols, atoi(w) / kChunkCols * kChunkCols), dim);
      p.bar = r.bar;
      p.timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
      p.rank = g;
      p.world = world;
      if (world > 1) {
        for (uint32_t h = 0; h < world; h++) {
          p.peer_S[0][h] = reinterpret_cast<float*>(R[h].xblock + sizeof(ExchangeHeader));
          p.peer_S[1][h] = reinterpret_cast<float*>(R[h].xblock + sizeof(ExchangeHeader)) + nvec;
          ExchangeHeader* hd = reinterpret_cast<ExchangeHeader*>(R[h].xblock);
          p.peer_arrive[h] = &hd->arrive;
          p.peer_smax3[h] = hd->smax3;
        }
        // as if earlier solves had left the barrier words somewhere (every rank's counter starts from the same total,
        // any round offset must work)
        p.round_base = 5ull + (o->max_iter % 3u);
        p.arrive_base = 3ull * world * kArriveUnits;
        p.S[0] = p.peer_S[0][g];
        p.S[1] = p.peer_S[1][g];
        p.flip = (uint32_t)(o->max_iter & 1u); // either parity offset must work
      }
      p.out_eigen_val = r.scalars;
      p.out_iter = reinterpret_cast<uint32_t*>(r.scalars + 1);
      const uint32_t stamped = std::min<uint32_t>(o->max_iter, 4096u);
      r.ts.assign(4 * ((size_t)stamped + 2), 0);
      p.round_ts = r.ts.data();
      p.phase_ts = r.ts.data() + stamped + 2;
      p.ts_rounds = stamped;
      if (form == kFormInPlace) {
        r.work.resize((size_t)rows * dim * sizeof(float) + 64);
        p.W = reinterpret_cast<float*>(r.work.data() + (64 - reinterpret_cast<uintptr_t>(r.work.data()) % 64) % 64);
      }

      // ---- launch shape and shared-memory layout (launch_plan.hpp, like Context::solve) ----
      r.threads = (unsigned)std::max(32, (o->threads + 31) / 32 * 32);
      const unsigned warps = r.threads / 32;
      unsigned want = (unsigned)std::max(1, o->ctas);
      void (*kernel)(const RoundParams) = nullptr;
      const int kid = o->kernel;
      if (kid == 1) {
        r.grid = std::max(1u, std::min(want, (rows + warps - 1) / warps));
        const uint32_t cap = (rows + r.grid - 1) / r.grid + 1;
        r.smem = sizeof(float) * ((size_t)p.chunk_cols + cap);
        kernel = o->stop ? general_kernel<kStopRelative>(vec4, form, bf16, acc64, fp8) : general_kernel<kStopAbsolute>(vec4, form, bf16, acc64, fp8);
      } else if (kid >= 10 && kid <= 13) {
        const bool scalar_units = !vec4 && kid == 11 && !bf16 && !fp8 && !acc64; // dim % 4 != 0: configuration 11, 4-byte units
        if ((!vec4 && !scalar_units) || form != kFormReadOnly || dim > (uint32_t)kResidentCols)
          throw std::string("resident-e kernel needs the read-only form, dim <= 32768 and dim % 4 == 0 (configuration 11: any dim)");
        int pf = -1;
        for (const ScConfig& c : kScConfigs)
          if (c.id == kid)
            pf = c.pf_batches;
        if ((bf16 || fp8) && pf != 0)
          throw std::string("bf16 / fp8 storage is built for configuration 11");
        r.grid = std::max(1u, std::min(want, (rows + warps - 1) / warps));
        const uint32_t cap = (rows + r.grid - 1) / r.grid + 1;
        uint32_t moff = 0;
        (void)cap;
        r.smem = sc_smem_bytes((int)r.threads, pf, dim, &moff);
        if (r.smem > kSmemLimit)
          throw std::string("resident-e configuration does not fit shared memory");
        p.mbar_offset = moff;
        p.chunk_cols = dim;
        const uint32_t units = (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols;
        if (units > 1u) {
          r.partial.assign((size_t)rows * units, -3.f);
          r.row_done.assign(rows, 0u);
          p.partial = r.partial.data();
          p.row_done = r.row_done.data();
        }
        kernel = o->stop ? sc_kernel<kStopRelative>(pf, bf16, acc64, fp8) : sc_kernel<kStopAbsolute>(pf, bf16, acc64, fp8);
        if (scalar_units)
          kernel = o->stop ? round_loop_sc_kernel<512, 0, kStopRelative, float, float, 1>
                           : round_loop_sc_kernel<512, 0, kStopAbsolute, float, float, 1>;
      } else if (kid == 2) {
        // wide kernel: windows of p.chunk_cols columns (kResidentCols; ST_EMU_WINDOW stages less so that several
        // windows fit an emulated size), chunk sums + per-row counters + one unit counter per window
        if (!vec4 || form != kFormReadOnly || bf16 || fp8 || acc64)
          throw std::string("wide kernel: read-only form, fp32, dim % 4 == 0");
        uint32_t window = (uint32_t)kResidentCols;
        if (const char* wv = getenv("ST_EMU_WINDOW"))
          window = (uint32_t)std::max(kChunkCols, atoi(wv) / kChunkCols * kChunkCols);
        window = std::min<uint32_t>(window, (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols * (uint32_t)kChunkCols);
        r.grid = std::max(1u, std::min(want, (rows + warps - 1) / warps));
        uint32_t moff = 0;
        r.smem = sc_smem_bytes((int)r.threads, 1, std::min<uint32_t>(window, dim), &moff);
        if (r.smem > kSmemLimit)
          throw std::string("wide kernel does not fit shared memory");
        p.mbar_offset = moff;
        p.chunk_cols = window;
        const uint32_t units = (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols;
        r.partial.assign((size_t)rows * units, -3.f);
        r.row_done.assign(rows, 0u);
        r.phase_counter.assign(32u * (size_t)kWideMaxWindows, 0u);
        p.partial = r.partial.data();
        p.row_done = r.row_done.data();
        p.phase_counter = r.phase_counter.data();
        kernel = o->stop ? round_loop_wide_kernel<512, kStopRelative> : round_loop_wide_kernel<512, kStopAbsolute>;
      } else if (kid == 20) {
        if (!vec4 || form != kFormReadOnly || bf16 || fp8 || world != 1 || dim > (uint32_t)kClusterCols)
          throw std::string("cluster kernel: one GPU, read-only form, fp32, dim % 4 == 0, dim <= 512");
        size_t smem = 0;
        const int c = cluster_ctas_for(dim, &smem);
        if (!c)
          throw std::string("matrix does not fit the cluster");
        r.grid = (unsigned)c;
        r.threads = 512;
        r.smem = smem;
        kernel = o->stop ? round_loop_cluster_kernel<512, kStopRelative> : round_loop_cluster_kernel<512, kStopAbsolute>;
      } else {
        throw std::string("unknown kernel id");
      }
      r.launch = emu::launch_async<RoundParams>(kernel, r.grid, r.threads, r.smem, p);
    }
    for (Rank& r : R)
      emu::join(*r.launch);

    int rc = 0;
    uint32_t agree = 1;
    for (uint32_t g = 0; g < world; g++) {
      Rank& r = R[g];
      if (r.bar->error != 0u) {
        g_err = r.bar->error == 2u ? "bulk-copy wait timed out" : "round barrier timed out";
        rc = -4;
      }
      if (g > 0) {
        agree &= memcmp(r.scalars, R[0].scalars, 12) == 0;
        agree &= memcmp(r.p.out_eigen_vec, R[0].p.out_eigen_vec, sizeof(float) * dim) == 0;
      }
    }
    if (rc == 0) {
      *eigen_val = R[0].scalars[0];
      memcpy(iter_count, &R[0].scalars[1], 4);
      memcpy(passes, &R[0].scalars[2], 4);
      memcpy(eigen_vec, R[0].p.out_eigen_vec, sizeof(float) * dim);
      if (ranks_agree)
        *ranks_agree = agree;
    }
    for (Rank& r : R) {
      free(r.xblock);
      free(r.bar);
    }
    return rc;
  } catch (const std::string& e) {
    g_err = e;
    return -2;
  }
}

// ---- small kernels -------------------------------------------------------------------------------------
extern "C" int
emu_find_max(const float* vec, uint32_t dim, unsigned ctas, float* out)
{
  struct P { const float* v; float* o; uint32_t n; };
  static float cell;
  cell = 0.f; // the host zero-fills the cell (solver.cu: find_max)
  P p{ vec, &cell, dim };
  auto g = emu::launch_async<P>([](const P q) { find_max_kernel(q.v, q.o, q.n); }, ctas, 1024, 0, p);
  emu::join(*g);
  *out = cell;
  return 0;
}

extern "C" int
emu_stop(const float* vec, uint32_t dim, float eps, unsigned ctas, uint32_t* out)
{
  struct P { const float* v; uint32_t* r; uint32_t n; float eps; };
  static uint32_t flag;
  flag = 1u; // the host fills the flag with 1 (solver.cu: stop)
  P p{ vec, &flag, dim, eps };
  auto g = emu::launch_async<P>([](const P q) { stop_kernel(q.v, q.r, q.n, q.eps); }, ctas, 1024, 0, p);
  emu::join(*g);
  *out = flag;
  return 0;
}

extern "C" int
emu_convert_bf16(const float* src, unsigned short* dst, size_t n, unsigned ctas)
{
  struct P { const float* s; unsigned short* d; size_t n; };
  P p{ src, dst, n };
  auto g = emu::launch_async<P>([](const P q) { convert_bf16_kernel(q.s, q.d, q.n); }, ctas, 256, 0, p);
  emu::join(*g);
  return 0;
}

extern "C" int
emu_convert_fp8(const float* src, unsigned char* dst, float* row_scale, uint32_t rows, uint32_t dim, unsigned ctas)
{
  struct P { const float* s; unsigned char* d; float* sc; uint32_t rows, dim; };
  P p{ src, dst, row_scale, rows, dim };
  auto g = emu::launch_async<P>([](const P q) { convert_fp8_rows_kernel(q.s, q.d, q.sc, q.rows, q.dim); }, ctas, 256, 0, p);
  emu::join(*g);
  return 0;
}

extern "C" int
emu_sum_across_rows(const float* mat, const float* e, float* vec, uint32_t dim, uint32_t row0, uint32_t rows, unsigned ctas)
{
  struct P { const float* m; const float* e; float* v; uint32_t dim, row0, rows; };
  P p{ mat, e, vec, dim, row0, rows };
  const size_t smem = std::min<uint32_t>((uint32_t)kChunkCols, dim) * sizeof(float);
  std::unique_ptr<emu::Grid> g;
  if (dim % 4u == 0u)
    g = emu::launch_async<P>([](const P q) { sum_across_rows_kernel<4>(q.m, q.e, q.v, q.dim, q.row0, q.rows); }, ctas, 256, smem, p);
  else
    g = emu::launch_async<P>([](const P q) { sum_across_rows_kernel<1>(q.m, q.e, q.v, q.dim, q.row0, q.rows); }, ctas, 256, smem, p);
  emu::join(*g);
  return 0;
}
