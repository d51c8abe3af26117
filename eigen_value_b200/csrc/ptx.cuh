// ptx.cuh -- every line of inline PTX the round kernels use, in one place.
//
// kernels*.cuh contain only CUDA C++ on top of these functions.  The CPU emulation harness
// (tests/cuda_emu) compiles the same kernel sources for the host and replaces this header by its
// own implementation of the same functions (ST_PTX_HEADER), so the kernels' index arithmetic,
// reductions, scheduling and barrier protocol can be exercised without a GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

// ---- clocks -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long
globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- scoped loads / stores / reductions for the round barrier and the cross-GPU flags --------------
__device__ __forceinline__ unsigned int
ld_acquire_gpu(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned int
ld_relaxed_gpu(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long
ld_relaxed_gpu(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long
ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void
red_release_gpu_add(unsigned int* p, unsigned int v)
{
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Cross-GPU barrier: fire-and-forget reductions at system scope, onto this GPU's or a peer's memory
// (over NVLink the atomic is performed at the owner's L2, like a local one)
__device__ __forceinline__ void
red_relaxed_sys_add(unsigned long long* p, unsigned long long v)
{
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void
red_relaxed_sys_max(unsigned int* p, unsigned int v)
{
  asm volatile("red.relaxed.sys.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// release/acquire fence at system scope (MEMBAR.ALL.SYS; the sequentially consistent __threadfence_system()
// measured the same, profiles/r2_c21_ab_barrier_8gpu.json)
__device__ __forceinline__ void
fence_acq_rel_sys()
{
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}

__device__ __forceinline__ unsigned int
ld_relaxed_sys(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void
st_relaxed_sys(unsigned int* p, unsigned int v)
{
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- matrix stream ------------------------------------------------------------------------------
// Streaming 128-bit load of matrix data that is never written while the kernel runs: read-only
// path, no L1 allocation (each byte is used exactly once per round).  The two-argument forms carry
// an L2 eviction-priority policy (createpolicy): rows the kernel wants to find in L2 again next
// round are loaded evict_last, the rest evict_first.  The uint4 form reads 8 bf16 elements.
__device__ __forceinline__ float4
ld_stream(const float4* p)
{
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p));
  return v;
}

// Scalar form (dim % 4 != 0): rows start on 4-byte boundaries, so a warp's 128-byte request straddles two lines and
// shares a 32-byte sector with the neighbouring request.  These loads DO allocate in L1: the shared sector is then
// fetched once instead of twice (5 sectors per request otherwise, 25 % more L2->SM traffic).
__device__ __forceinline__ float
ld_stream(const float* p)
{
  float v;
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float4
ld_stream(const float4* p, unsigned long long pol)
{
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "l"(pol));
  return v;
}

__device__ __forceinline__ float
ld_stream(const float* p, unsigned long long pol)
{
  float v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

__device__ __forceinline__ uint4
ld_stream(const uint4* p)
{
  uint4 v;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p));
  return v;
}

// one 64-bit word of bf16 storage (four elements)
__device__ __forceinline__ uint2
ld_stream(const uint2* p)
{
  uint2 v;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// one 32-bit word of fp8 storage (four elements)
__device__ __forceinline__ uint32_t
ld_stream(const uint32_t* p)
{
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ unsigned long long
l2_policy_evict_last()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ unsigned long long
l2_policy_evict_first()
{
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// fp32 -> bf16, round to nearest even, for the bf16-storage solves
__device__ __forceinline__ unsigned short
f32_to_bf16_rn(float x)
{
  unsigned short h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}

// ---- fp8 (e4m3) storage of the matrix ---------------------------------------------------------------
// two floats -> two e4m3 codes, round to nearest even, saturating at +-448 (NaN -> 0x7f); `lo` lands in the low byte
__device__ __forceinline__ unsigned short
f32x2_to_fp8x2(float lo, float hi)
{
  unsigned short r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- TMA bulk copies and their mbarriers ----------------------------------------------------------
__device__ __forceinline__ uint32_t
smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void
mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool
mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}

__device__ __forceinline__ void
bulk_load(float* dst_smem, const float* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void
fence_mbarrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// orders generic-proxy accesses to shared memory before the async proxy's (bulk copy) accesses
__device__ __forceinline__ void
fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace st
