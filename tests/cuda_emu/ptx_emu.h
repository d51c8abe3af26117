// ptx_emu.h -- host implementation of eigen_value_b200/csrc/ptx.cuh for the CPU emulation harness
// (TEST INFRASTRUCTURE; selected through -DST_PTX_HEADER).  Same names and signatures; scoped
// loads/stores become GCC atomics, bulk copies complete at issue, an mbarrier is a counter of
// completed phases.
#pragma once

#include "cuda_emu.h"

namespace st {

inline unsigned long long
globaltimer_ns()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

// polling loads give the other OS threads (CTAs, emulated GPUs) a chance to run
inline unsigned int
ld_acquire_gpu(const unsigned int* p)
{
  sched_yield();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
inline unsigned int
ld_relaxed_gpu(const unsigned int* p)
{
  return __atomic_load_n(p, EMU_RELAXED_LOAD);
}
inline unsigned long long
ld_relaxed_gpu(const unsigned long long* p)
{
  return __atomic_load_n(p, EMU_RELAXED_LOAD);
}
inline unsigned long long
ld_acquire_sys(const unsigned long long* p)
{
  sched_yield();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
inline void
red_release_gpu_add(unsigned int* p, unsigned int v)
{
  __atomic_fetch_add(p, v, __ATOMIC_RELEASE);
}
inline void
red_relaxed_sys_add(unsigned long long* p, unsigned long long v)
{
  __atomic_fetch_add(p, v, EMU_RELAXED_STORE);
}
inline void
red_relaxed_sys_max(unsigned int* p, unsigned int v)
{
  unsigned int cur = __atomic_load_n(p, EMU_RELAXED_LOAD);
  while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, true, EMU_RELAXED_STORE, EMU_RELAXED_LOAD))
    ;
}
inline void
fence_acq_rel_sys()
{
  __atomic_thread_fence(__ATOMIC_ACQ_REL);
}
inline unsigned int
ld_relaxed_sys(const unsigned int* p)
{
  return __atomic_load_n(p, EMU_RELAXED_LOAD);
}
inline void
st_relaxed_sys(unsigned int* p, unsigned int v)
{
  __atomic_store_n(p, v, EMU_RELAXED_STORE);
}

inline float4 ld_stream(const float4* p) { return *p; }
inline float ld_stream(const float* p) { return *p; }
inline float4 ld_stream(const float4* p, unsigned long long) { return *p; }
inline float ld_stream(const float* p, unsigned long long) { return *p; }
inline uint4 ld_stream(const uint4* p) { return *p; }
inline uint32_t ld_stream(const uint32_t* p) { return *p; }
inline uint2 ld_stream(const uint2* p) { return *p; }
inline unsigned long long l2_policy_evict_last() { return 0ull; }
inline unsigned long long l2_policy_evict_first() { return 0ull; }

// cvt.rn.bf16.f32: round to nearest even; NaN -> canonical 0x7fff
inline unsigned short
f32_to_bf16_rn(float x)
{
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u)
    return 0x7fffu;
  return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

// e4m3 (fn variant: no infinities, 0x7f / 0xff = NaN): value of one code
inline float
emu_fp8_value(unsigned char c)
{
  const int e = (c >> 3) & 15, m = c & 7;
  float v;
  if (e == 15 && m == 7)
    v = NAN;
  else if (e == 0)
    v = ldexpf((float)m, -9);
  else
    v = ldexpf(1.f + (float)m / 8.f, e - 7);
  return (c & 0x80) ? -v : v;
}
// cvt.rn.satfinite.e4m3x2.f32: nearest code, ties to the even code, saturating at 448; NaN -> 0x7f
inline unsigned char
emu_f32_to_fp8(float x)
{
  if (x != x)
    return 0x7f;
  const unsigned char sign = std::signbit(x) ? 0x80 : 0x00;
  const float a = fabsf(x);
  if (a >= 448.f)
    return sign | 0x7e;
  unsigned char lo = 0, hi = 0x7e; // largest code <= a, by bisection over the monotonic code values
  while (lo < hi) {
    const unsigned char mid = (unsigned char)((lo + hi + 1) / 2);
    if (emu_fp8_value(mid) <= a)
      lo = mid;
    else
      hi = (unsigned char)(mid - 1);
  }
  unsigned char best = lo;
  if (lo < 0x7e) {
    const float dl = a - emu_fp8_value(lo), dh = emu_fp8_value((unsigned char)(lo + 1)) - a;
    if (dh < dl || (dh == dl && (lo & 1)))
      best = (unsigned char)(lo + 1);
  }
  return sign | best;
}
inline unsigned short
f32x2_to_fp8x2(float lo, float hi)
{
  return (unsigned short)(emu_f32_to_fp8(lo) | (emu_f32_to_fp8(hi) << 8));
}

inline uint32_t
smem_u32(const void* p)
{
  return (uint32_t)reinterpret_cast<uintptr_t>(p);
}
// an mbarrier is the number of phases completed so far
EMU_NO_TSAN inline void
mbar_init(uint64_t* bar, uint32_t)
{
  *bar = 0;
}
EMU_NO_TSAN inline void
emu_bump(uint64_t* bar)
{
  *bar += 1;
}
inline void
mbar_arrive_expect_tx(uint64_t*, uint32_t)
{
}
EMU_NO_TSAN inline bool
mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  const bool done = ((uint32_t)(*bar) & 1u) != parity; // the phase of that parity has completed
  if (done)
    EMU_ACQUIRE(bar); // the copied bytes are visible to the waiter
  else
    emu::yield(); // a polling lane must let the lane that issues the copy run (the hardware's independent
                  // thread scheduling guarantees that progress; a cooperative fiber has to hand over)
  return done;
}
inline void
bulk_load(float* dst_smem, const float* src_gmem, uint32_t bytes, uint64_t* bar)
{
  if (bytes % 16u != 0u || (reinterpret_cast<uintptr_t>(dst_smem) & 15u) || (reinterpret_cast<uintptr_t>(src_gmem) & 15u)) {
    fprintf(stderr, "cuda_emu: bulk copy with misaligned address or size (%u bytes)\n", bytes);
    abort(); // the hardware would fault
  }
  memcpy(dst_smem, src_gmem, bytes);
  EMU_RELEASE(bar);
  emu_bump(bar);
}
inline void
fence_mbarrier_init()
{
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
inline void
fence_proxy_async()
{
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
}

} // namespace st
