// cuda_emu.h -- a minimal CPU execution model for the round kernels (TEST INFRASTRUCTURE).
//
// The kernel sources under eigen_value_b200/csrc are plain CUDA C++ on top of ptx.cuh.  This
// header supplies just enough of the CUDA execution model for g++ to compile those sources
// UNCHANGED (apart from the `extern __shared__` declarations, rewritten by build.py) and run them
// on the host with small launch shapes:
//
//   * one OS thread per CTA (so `__shared__` variables become `static thread_local`), every CUDA
//     thread of the CTA a fiber (ucontext) scheduled round-robin inside that OS thread;
//   * __syncthreads / __syncwarp / __shfl_*_sync are rendezvous points between fibers;
//   * global-memory atomics and the scoped loads/stores of ptx.cuh map to GCC __atomic builtins, so
//     CTAs (OS threads) and emulated GPUs (groups of OS threads) really run concurrently and the
//     grid barrier / the cross-GPU flag exchange are exercised as written;
//   * bulk copies (TMA) complete synchronously at issue; mbarriers count completed phases.
//
// What this checks: index arithmetic, reduction order (bit-exact against the oracle), work-unit
// scheduling, stop logic, the barrier protocol's liveness.  What it cannot check: anything about the
// hardware (memory model subtleties, alignment faults, performance).
#pragma once

#include <ucontext.h>
#include <sched.h>
#include <time.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

// ---- ThreadSanitizer mode (tests/cuda_emu/build.py --tsan): a CPU "racecheck" ---------------------------
// Every fiber is announced to TSan as a thread of its own and fiber switches do NOT synchronise, so two
// CUDA threads touching the same shared or global address are a reported race unless the kernel ordered
// them through __syncthreads / __syncwarp / a shuffle / an mbarrier (annotated below as release +
// acquire) or through the atomics of the grid barrier.  __ldcg / __stcg become plain accesses so that the
// row-sum exchange itself is checked.  TSan does not model stand-alone fences, so in this mode relaxed
// atomics are strengthened to acq_rel: what is verified is that a synchronisation chain EXISTS between
// conflicting accesses (barrier placement, double buffering, slot reuse), not the choice of fence.
#if defined(__SANITIZE_THREAD__)
#include <sanitizer/tsan_interface.h>
#define EMU_TSAN 1
#define EMU_NO_TSAN __attribute__((no_sanitize("thread")))
#define EMU_RELEASE(addr) __tsan_release((void*)(addr))
#define EMU_ACQUIRE(addr) __tsan_acquire((void*)(addr))
#define EMU_RELAXED __ATOMIC_ACQ_REL
#define EMU_RELAXED_LOAD __ATOMIC_ACQUIRE
#define EMU_RELAXED_STORE __ATOMIC_RELEASE
#else
#define EMU_TSAN 0
#define EMU_NO_TSAN
#define EMU_RELEASE(addr) ((void)0)
#define EMU_ACQUIRE(addr) ((void)0)
#define EMU_RELAXED __ATOMIC_RELAXED
#define EMU_RELAXED_LOAD __ATOMIC_RELAXED
#define EMU_RELAXED_STORE __ATOMIC_RELAXED
#endif

// ---- language keywords ---------------------------------------------------------------------------
#define __global__ inline
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static thread_local
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types ----------------------------------------------------------------------------------
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(8) uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
struct dim3 { unsigned int x = 1, y = 1, z = 1; dim3() = default; dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }

inline thread_local uint3 threadIdx{ 0, 0, 0 };
inline thread_local uint3 blockIdx{ 0, 0, 0 };
inline thread_local dim3 blockDim;
inline thread_local dim3 gridDim;

// ---- fiber switch -------------------------------------------------------------------------------------------
// Outside ThreadSanitizer builds a fiber switch is a dozen instructions (callee-saved registers + stack
// pointer); ucontext's swapcontext costs two signal-mask system calls per switch, which dominated the
// emulation.  TSan builds keep ucontext (TSan knows it).
#if !EMU_TSAN && defined(__x86_64__)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void** save_sp, void* next_sp);
asm(R"(
.text
.weak emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#else
#define EMU_FAST_SWITCH 0
#endif

namespace emu {

constexpr size_t kStackBytes = 96 * 1024;

struct Grid; // one kernel launch

struct WarpState
{
  uint32_t vals[2][32];
  uint32_t arrived = 0;
  uint32_t gen = 0;
};

struct Cta
{
  Grid* grid = nullptr;
  unsigned index = 0, threads = 0;
  unsigned char* smem = nullptr; // dynamic shared memory
  std::vector<ucontext_t> ctx;
  std::vector<void*> sp;  // fast switch: saved stack pointer per fiber
  void* sched_sp = nullptr;
  std::vector<std::unique_ptr<unsigned char[]>> stacks;
  std::vector<char> done;
  ucontext_t sched;
  unsigned current = 0;
  unsigned live = 0;
  // __syncthreads
  unsigned sync_arrived = 0, sync_gen = 0;
  int sync_and = 1, sync_and_result = 1;
  std::vector<WarpState> warps;
  std::function<void()> body;
  std::vector<void*> tsan_fiber; // TSan mode: one announced fiber per CUDA thread
  void* tsan_sched = nullptr;
};

struct Grid
{
  unsigned ctas = 0, threads = 0;
  size_t smem_bytes = 0;
  std::vector<std::unique_ptr<Cta>> cta;
  std::vector<std::thread> workers;
  // cluster-wide barrier (cooperative_groups::cluster_group::sync)
  std::atomic<unsigned> cl_count{ 0 };
  std::atomic<unsigned> cl_gen{ 0 };
};

inline thread_local Cta* g_cta = nullptr;

EMU_NO_TSAN inline void
to_scheduler(Cta* c)
{
#if EMU_TSAN
  __tsan_switch_to_fiber(c->tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
#if EMU_FAST_SWITCH
  emu_switch(&c->sp[c->current], c->sched_sp);
#else
  swapcontext(&c->ctx[c->current], &c->sched);
#endif
}

EMU_NO_TSAN inline void
yield()
{
  to_scheduler(g_cta);
}

inline unsigned char*
dynamic_smem()
{
  return g_cta->smem;
}

EMU_NO_TSAN inline void
fiber_exit(Cta* c)
{
  c->done[c->current] = 1;
  c->live--;
  EMU_RELEASE(&c->live); // kernel completion: this thread's writes happen-before the host's reads after join
  to_scheduler(c);
}

inline void
fiber_entry()
{
  Cta* c = g_cta;
  EMU_ACQUIRE(&c->threads); // kernel launch: the host's writes happen-before this thread
  c->body();
  fiber_exit(c);
  abort(); // a finished fiber is never resumed
}

EMU_NO_TSAN inline void
run_cta(Cta* c, uint3 block_index, dim3 grid_dim)
{
  g_cta = c;
  blockIdx = block_index;
  blockDim = dim3(c->threads);
  gridDim = grid_dim;
#if !EMU_FAST_SWITCH
  c->ctx.resize(c->threads);
#endif
  c->done.assign(c->threads, 0);
  c->warps.assign((c->threads + 31) / 32, WarpState{});
  c->live = c->threads;
#if EMU_FAST_SWITCH
  c->sp.resize(c->threads);
  for (unsigned t = 0; t < c->threads; t++) {
    c->stacks.emplace_back(new unsigned char[kStackBytes]);
    // initial frame: six callee-saved register slots, then the "return address" emu_switch's ret jumps to;
    // fiber_entry then starts with the stack alignment a call would have left (rsp % 16 == 8)
    uintptr_t top = (reinterpret_cast<uintptr_t>(c->stacks.back().get()) + kStackBytes) & ~(uintptr_t)15;
    void** frame = reinterpret_cast<void**>(top - 64);
    for (int i = 0; i < 6; i++)
      frame[i] = nullptr;
    frame[6] = reinterpret_cast<void*>(&fiber_entry);
    frame[7] = nullptr;
    c->sp[t] = frame;
  }
#else
  for (unsigned t = 0; t < c->threads; t++) {
    c->stacks.emplace_back(new unsigned char[kStackBytes]);
    getcontext(&c->ctx[t]);
    c->ctx[t].uc_stack.ss_sp = c->stacks.back().get();
    c->ctx[t].uc_stack.ss_size = kStackBytes;
    c->ctx[t].uc_link = &c->sched;
    makecontext(&c->ctx[t], (void (*)())fiber_entry, 0);
  }
#endif
#if EMU_TSAN
  c->tsan_sched = __tsan_get_current_fiber();
  c->tsan_fiber.resize(c->threads);
  for (unsigned t = 0; t < c->threads; t++)
    c->tsan_fiber[t] = __tsan_create_fiber(0);
#endif
  EMU_RELEASE(&c->threads);
  while (c->live) {
    for (unsigned t = 0; t < c->threads; t++) {
      if (c->done[t])
        continue;
      c->current = t;
      threadIdx = uint3{ t, 0, 0 };
#if EMU_TSAN
      __tsan_switch_to_fiber(c->tsan_fiber[t], __tsan_switch_to_fiber_no_sync);
#endif
#if EMU_FAST_SWITCH
      emu_switch(&c->sched_sp, c->sp[t]);
#else
      swapcontext(&c->sched, &c->ctx[t]);
#endif
    }
  }
  EMU_ACQUIRE(&c->live);
#if EMU_TSAN
  for (unsigned t = 0; t < c->threads; t++)
    __tsan_destroy_fiber(c->tsan_fiber[t]);
#endif
  g_cta = nullptr;
}

// Launches `kernel(params)` on `ctas` x `threads`; returns the grid (join() waits).
template<typename P>
std::unique_ptr<Grid>
launch_async(void (*kernel)(const P), unsigned ctas, unsigned threads, size_t smem_bytes, const P& params)
{
  auto g = std::make_unique<Grid>();
  g->ctas = ctas;
  g->threads = threads;
  g->smem_bytes = smem_bytes;
  for (unsigned b = 0; b < ctas; b++) {
    auto c = std::make_unique<Cta>();
    c->grid = g.get();
    c->index = b;
    c->threads = threads;
    c->smem = static_cast<unsigned char*>(aligned_alloc(1024, (smem_bytes + 1023) / 1024 * 1024 + 1024));
    memset(c->smem, 0xCD, smem_bytes); // shared memory starts as garbage, like the hardware's
    c->body = [kernel, params] { kernel(params); };
    g->cta.push_back(std::move(c));
  }
  for (unsigned b = 0; b < ctas; b++)
    g->workers.emplace_back(run_cta, g->cta[b].get(), uint3{ b, 0, 0 }, dim3(ctas));
  return g;
}

inline void
join(Grid& g)
{
  for (auto& w : g.workers)
    w.join();
  for (auto& c : g.cta)
    free(c->smem);
}

} // namespace emu

// ---- block / warp synchronisation ---------------------------------------------------------------
// All rendezvous are generation-counted: a fiber that arrives bumps the counter, the last one opens the
// next generation, the others yield until the generation changes.  Fibers of one CTA never run
// concurrently (one OS thread), so no locking is needed.
EMU_NO_TSAN inline int
emu_block_rendezvous(int pred)
{
  emu::Cta* c = emu::g_cta;
  const unsigned g = c->sync_gen;
  EMU_RELEASE(&c->sync_gen); // everything this thread did so far happens-before every departure
  c->sync_and &= (pred != 0);
  if (++c->sync_arrived == c->live) {
    c->sync_and_result = c->sync_and;
    c->sync_and = 1;
    c->sync_arrived = 0;
    c->sync_gen = g + 1;
  } else {
    while (c->sync_gen == g)
      emu::yield();
  }
  EMU_ACQUIRE(&c->sync_gen);
  return c->sync_and_result;
}
inline void
__syncthreads()
{
  emu_block_rendezvous(1);
}
inline int
__syncthreads_and(int pred)
{
  // the result must be read by every thread before the NEXT barrier can overwrite it: a second
  // rendezvous keeps it stable (costly, but this is used by one small kernel only)
  const int r = emu_block_rendezvous(pred);
  emu_block_rendezvous(1);
  return r;
}

EMU_NO_TSAN inline uint32_t
emu_warp_exchange(uint32_t mine, unsigned src_lane)
{
  emu::Cta* c = emu::g_cta;
  emu::WarpState& w = c->warps[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned width = std::min(32u, c->threads - (threadIdx.x & ~31u));
  const uint32_t g = w.gen;
  EMU_RELEASE(&w.gen); // a *_sync shuffle / __syncwarp orders the warp's memory accesses
  w.vals[g & 1u][lane] = mine;
  if (++w.arrived == width) {
    w.arrived = 0;
    w.gen = g + 1;
  } else {
    while (w.gen == g)
      emu::yield();
  }
  EMU_ACQUIRE(&w.gen);
  return src_lane < width ? w.vals[g & 1u][src_lane] : mine;
}
inline void
__syncwarp(unsigned = 0xffffffffu)
{
  emu_warp_exchange(0u, 0u);
}

template<typename T>
inline T
emu_shfl(T v, unsigned src_lane)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only");
  uint32_t bits[2] = { 0u, 0u };
  memcpy(bits, &v, sizeof(T));
  bits[0] = emu_warp_exchange(bits[0], src_lane);
  if (sizeof(T) == 8)
    bits[1] = emu_warp_exchange(bits[1], src_lane); // a 64-bit shuffle is two 32-bit ones, as on the hardware
  T out;
  memcpy(&out, bits, sizeof(T));
  return out;
}
template<typename T>
inline T
__shfl_xor_sync(unsigned, T v, int lane_mask)
{
  return emu_shfl(v, (threadIdx.x & 31u) ^ (unsigned)lane_mask);
}
template<typename T>
inline T
__shfl_down_sync(unsigned, T v, unsigned delta)
{
  const unsigned lane = threadIdx.x & 31u;
  return emu_shfl(v, lane + delta < 32u ? lane + delta : lane); // out of range: own value
}
template<typename T>
inline T
__shfl_sync(unsigned, T v, int src_lane)
{
  return emu_shfl(v, (unsigned)src_lane & 31u);
}

// ---- memory ------------------------------------------------------------------------------------------
#if EMU_TSAN
inline float __ldcg(const float* p) { return *p; }       // plain: TSan checks the exchange is ordered
inline void __stcg(float* p, float v) { *p = v; }
#else
inline float __ldcg(const float* p) { float v; uint32_t b = __atomic_load_n(reinterpret_cast<const uint32_t*>(p), __ATOMIC_RELAXED); memcpy(&v, &b, 4); return v; }
inline void __stcg(float* p, float v) { uint32_t b; memcpy(&b, &v, 4); __atomic_store_n(reinterpret_cast<uint32_t*>(p), b, __ATOMIC_RELAXED); }
#endif
inline float4 __ldcg(const float4* p)
{
  float4 v;
  v.x = __ldcg(&p->x); v.y = __ldcg(&p->y); v.z = __ldcg(&p->z); v.w = __ldcg(&p->w);
  return v;
}
inline void __stcg(float4* p, float4 v) { __stcg(&p->x, v.x); __stcg(&p->y, v.y); __stcg(&p->z, v.z); __stcg(&p->w, v.w); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, EMU_RELAXED); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, EMU_RELAXED); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, EMU_RELAXED); }
inline int atomicMax(int* p, int v)
{
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}

inline unsigned atomicMax(unsigned* p, unsigned v)
{
  unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, EMU_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v)
{
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, EMU_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}

// ---- arithmetic --------------------------------------------------------------------------------------
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max;
using std::min;
inline uint32_t min(uint32_t a, int b) { return a < (uint32_t)b ? a : (uint32_t)b; }
inline uint32_t min(int a, uint32_t b) { return (uint32_t)a < b ? (uint32_t)a : b; }
