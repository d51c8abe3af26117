// cuda_runtime_emu.h -- the slice of the CUDA runtime API that eigen_value_b200/csrc/solver.cu and abi.cu
// use, implemented on the CPU emulation harness (TEST INFRASTRUCTURE; see cuda_emu.h).
//
// With it the WHOLE library -- C ABI, Context::solve's launch planning, the kernels -- builds into
// tests/cuda_emu/libsimilarity_transform_emu.so and the `-m gpu` test files can be exercised on the CPU
// against a small pretend device (tests/conftest.py, ST_EMULATED_LIB=1).  It is never loaded by the
// product package.
//   * "devices": ST_EMU_DEVICES (default 1) pretend B200s with ST_EMU_SMS (default 4) SMs each;
//   * device memory is host memory, copies are memcpy, streams are in-order because every call is
//     synchronous; events record a host clock;
//   * cooperative / cluster launches run all CTAs concurrently (cuda_emu.h); other launches run their
//     CTAs on a few worker threads, elementwise kernels without fibers;
//   * IPC handles carry the pointer itself (same-process only).
#pragma once

#include "cuda_emu.h"

#include <mutex>

typedef int cudaError_t;
enum : int
{
  cudaSuccess = 0,
  cudaErrorInvalidValue = 1,
  cudaErrorMemoryAllocation = 2,
  cudaErrorInvalidDevice = 101,
  cudaErrorPeerAccessAlreadyEnabled = 704
};
typedef struct CUstream_st* cudaStream_t;
typedef struct CUevent_st* cudaEvent_t;
struct CUstream_st { int device; };
struct CUevent_st { double t_ms; };

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDeviceAttr { cudaDevAttrCooperativeLaunch = 95 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4 };

struct cudaDeviceProp
{
  char name[256];
  int major, minor, multiProcessorCount, l2CacheSize;
  size_t totalGlobalMem;
};
struct cudaFuncAttributes { int numRegs; size_t sharedSizeBytes; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaLaunchAttributeValue { struct { unsigned x, y, z; } clusterDim; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t
{
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes = 0;
  cudaStream_t stream = nullptr;
  cudaLaunchAttribute* attrs = nullptr;
  unsigned numAttrs = 0;
};

namespace emu_rt {

inline int
env_int(const char* name, int fallback)
{
  const char* v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}
inline int devices() { return std::max(1, std::min(8, env_int("ST_EMU_DEVICES", 1))); }
inline int sms() { return std::max(1, std::min(148, env_int("ST_EMU_SMS", 4))); }
inline thread_local int current_device = 0;
inline double
now_ms()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Non-cooperative launch: CTAs are independent, so they are run one after another on a few worker
// threads.  `fibers` == false: the kernel's threads do not communicate either, each is a plain call.
template<typename F>
inline void
launch_grid(dim3 grid, dim3 block, size_t smem, bool fibers, F body)
{
  const unsigned total = grid.x * grid.y * grid.z;
  const unsigned workers = std::max(1u, std::min(total, std::min(8u, std::thread::hardware_concurrency())));
  std::atomic<unsigned> next{ 0 };
  auto work = [&] {
    std::vector<unsigned char> shared(smem + 1024);
    for (;;) {
      const unsigned b = next.fetch_add(1);
      if (b >= total)
        break;
      const uint3 bidx{ b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y) };
      if (!fibers) {
        blockIdx = bidx;
        blockDim = block;
        gridDim = grid;
        for (unsigned t = 0; t < block.x; t++) {
          threadIdx = uint3{ t, 0, 0 };
          body();
        }
      } else {
        emu::Grid g;
        g.ctas = total;
        g.threads = block.x;
        emu::Cta c;
        c.grid = &g;
        c.index = b;
        c.threads = block.x;
        c.smem = shared.data() + (1024 - reinterpret_cast<uintptr_t>(shared.data()) % 1024) % 1024;
        c.body = body;
        emu::run_cta(&c, bidx, grid);
      }
    }
  };
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; w++)
    pool.emplace_back(work);
  work();
  for (auto& t : pool)
    t.join();
}

} // namespace emu_rt

#define ST_LAUNCH(kernel, grid, block, smem, stream, ...)                                                        \
  emu_rt::launch_grid(dim3(grid), dim3(block), (size_t)(smem), true, [=] { kernel(__VA_ARGS__); })
#define ST_LAUNCH_ELEMENTWISE(kernel, grid, block, smem, stream, ...)                                            \
  emu_rt::launch_grid(dim3(grid), dim3(block), (size_t)(smem), false, [=] { kernel(__VA_ARGS__); })

// ---- devices --------------------------------------------------------------------------------------------
inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu_rt::devices(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu_rt::devices()) return cudaErrorInvalidDevice; emu_rt::current_device = d; return cudaSuccess; }
inline cudaError_t
cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "emulated B200 (CPU, %d SMs)", emu_rt::sms());
  p->major = 10;
  p->multiProcessorCount = emu_rt::sms();
  p->l2CacheSize = 126 * 1024 * 1024;
  p->totalGlobalMem = (size_t)emu_rt::env_int("ST_EMU_HBM_MIB", 4096) << 20;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetPCIBusId(char* buf, int len, int) { if (len > 0) buf[0] = 0; return cudaErrorInvalidDevice; } // no PCI device: no CPU binding
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error"; }

// ---- streams / events: every operation is synchronous, so a stream is trivially in order ----------------
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new CUstream_st{ emu_rt::current_device }; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CUevent_st{ 0.0 }; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new CUevent_st{ 0.0 }; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; } // everything already happened
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = emu_rt::now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }

// ---- memory ----------------------------------------------------------------------------------------------
template<typename T>
inline cudaError_t
cudaMalloc(T** p, size_t bytes)
{
  // the pretend device has ST_EMU_HBM_MIB of memory per allocation: larger requests fail like a full GPU would
  if (bytes > ((size_t)emu_rt::env_int("ST_EMU_HBM_MIB", 4096) << 20)) {
    *p = nullptr;
    return cudaErrorMemoryAllocation;
  }
  *p = static_cast<T*>(aligned_alloc(256, (bytes + 255) / 256 * 256 + 256));
  if (!*p)
    return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, bytes); // fresh device memory is garbage
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t
cudaMemGetInfo(size_t* free_b, size_t* total_b)
{
  *free_b = *total_b = (size_t)emu_rt::env_int("ST_EMU_HBM_MIB", 4096) << 20;
  return cudaSuccess;
}
template<typename T>
inline cudaError_t
cudaHostAlloc(T** p, size_t bytes, unsigned)
{
  *p = static_cast<T*>(aligned_alloc(64, (bytes + 63) / 64 * 64));
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
enum { cudaHostRegisterDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
namespace emu_rt {
inline std::mutex& pinned_mutex() { static std::mutex m; return m; }
inline std::vector<std::pair<const char*, size_t>>& pinned_ranges() { static std::vector<std::pair<const char*, size_t>> v; return v; }
} // namespace emu_rt
inline cudaError_t
cudaHostRegister(void* p, size_t n, unsigned)
{
  std::lock_guard<std::mutex> lock(emu_rt::pinned_mutex());
  emu_rt::pinned_ranges().emplace_back(static_cast<const char*>(p), n);
  return cudaSuccess;
}
inline cudaError_t
cudaHostUnregister(void* p)
{
  std::lock_guard<std::mutex> lock(emu_rt::pinned_mutex());
  auto& v = emu_rt::pinned_ranges();
  for (size_t i = 0; i < v.size(); i++)
    if (v[i].first == p) {
      v.erase(v.begin() + i);
      break;
    }
  return cudaSuccess;
}
inline cudaError_t
cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p)
{
  std::lock_guard<std::mutex> lock(emu_rt::pinned_mutex());
  a->type = cudaMemoryTypeUnregistered; // device memory is never asked about
  for (auto& r : emu_rt::pinned_ranges())
    if (static_cast<const char*>(p) >= r.first && static_cast<const char*>(p) < r.first + r.second)
      a->type = cudaMemoryTypeHost;
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t
cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
  memset(h, 0, sizeof *h);
  memcpy(h->reserved, &p, sizeof p);
  return cudaSuccess;
}
inline cudaError_t
cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
  memcpy(p, h.reserved, sizeof *p);
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- kernels ---------------------------------------------------------------------------------------------
// The launch limits the planner has to respect on a B200 are enforced here too: dynamic shared memory
// above 48 KB needs the opt-in attribute and can never exceed 227 KB, a CTA has at most 1024 threads, and a
// cooperative grid must be co-resident (these kernels occupy a whole SM each).
enum : int { cudaErrorCooperativeLaunchTooLarge = 720, cudaErrorInvalidConfiguration = 9 };
namespace emu_rt {
constexpr size_t kDefaultDynamicSmem = 48 * 1024, kMaxDynamicSmem = 227 * 1024;
inline std::mutex& attr_mutex() { static std::mutex m; return m; }
inline std::vector<std::pair<const void*, size_t>>& smem_optin() { static std::vector<std::pair<const void*, size_t>> v; return v; }
inline size_t
allowed_smem(const void* func)
{
  std::lock_guard<std::mutex> lock(attr_mutex());
  for (auto& e : smem_optin())
    if (e.first == func)
      return e.second;
  return kDefaultDynamicSmem;
}
inline cudaError_t
check_launch(const void* func, dim3 grid, dim3 block, size_t smem, bool cooperative)
{
  if (block.x == 0 || block.x > 1024 || grid.x == 0)
    return cudaErrorInvalidConfiguration;
  if (smem > allowed_smem(func))
    return cudaErrorInvalidValue;
  if (cooperative && grid.x * grid.y * grid.z > (unsigned)sms())
    return cudaErrorCooperativeLaunchTooLarge;
  return cudaSuccess;
}
} // namespace emu_rt
inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, const void*) { a->numRegs = 0; a->sharedSizeBytes = 0; return cudaSuccess; }
template<typename K>
inline cudaError_t
cudaFuncSetAttribute(K func, cudaFuncAttribute, int value)
{
  if (value < 0 || (size_t)value > emu_rt::kMaxDynamicSmem)
    return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lock(emu_rt::attr_mutex());
  const void* key = reinterpret_cast<const void*>(func);
  for (auto& e : emu_rt::smem_optin())
    if (e.first == key) {
      e.second = (size_t)value;
      return cudaSuccess;
    }
  emu_rt::smem_optin().emplace_back(key, (size_t)value);
  return cudaSuccess;
}

namespace st { struct RoundParams; }
// every cooperative kernel of this library takes one `const RoundParams` argument
cudaError_t emu_launch_round_kernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem);
inline cudaError_t
cudaLaunchCooperativeKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t)
{
  const cudaError_t e = emu_rt::check_launch(func, grid, block, smem, true);
  return e != cudaSuccess ? e : emu_launch_round_kernel(func, grid, block, args, smem);
}
template<typename P>
inline cudaError_t
cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(const P), P params)
{
  const cudaError_t e = emu_rt::check_launch(reinterpret_cast<const void*>(kernel), cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, false);
  if (e != cudaSuccess || cfg->gridDim.x > 8) // portable cluster size
    return e != cudaSuccess ? e : cudaErrorInvalidConfiguration;
  auto g = emu::launch_async<P>(kernel, cfg->gridDim.x, cfg->blockDim.x, cfg->dynamicSmemBytes, params); // cluster == grid
  emu::join(*g);
  return cudaSuccess;
}
