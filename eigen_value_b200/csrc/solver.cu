// solver.cu -- host side of the B200 similarity_transform(): context, scratch, launches.
//
// Replaces the host loop of reference similarity_transform.cpp:5-75 (8 queue submissions and
// one blocking host read per round) by ONE cooperative launch of st::round_loop_kernel; the
// host only stages buffers, launches, and reads back (lambda, iter_count, timestamps).
#include "similarity_transform.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <stdexcept>
#include <thread>

#include <pthread.h>
#include <sched.h>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "kernels_sc.cuh"
#include "kernels_wide.cuh"
#include "kernels_cluster.cuh"
#include "launch_plan.hpp"

static_assert(sizeof(st::ExchangeHeader) == st::kExchangeHeaderBytes, "exchange block header layout");

// Triple-chevron launches go through two macros so that the host code of this file can also be built
// for the CPU emulation harness (tests/cuda_emu), whose stand-in <cuda_runtime.h> defines them first:
//   ST_LAUNCH              kernels whose threads cooperate (barriers, shuffles, shared memory)
//   ST_LAUNCH_ELEMENTWISE  kernels whose threads are independent of each other
#ifndef ST_LAUNCH
#define ST_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define ST_LAUNCH_ELEMENTWISE(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

namespace st {

static_assert(kPlanClusterMaxCtas == kClusterMaxCtas && kPlanClusterSmemBudget == kClusterSmemBudget,
              "launch_plan.hpp must agree with kernels_cluster.cuh");

// ---- error plumbing -----------------------------------------------------------------------
static thread_local std::string g_last_error;

void
set_last_error(const std::string& msg)
{
  g_last_error = msg;
}
const char*
last_error()
{
  return g_last_error.c_str();
}

struct CudaError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

#define ST_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),          \
               __FILE__, __LINE__);                                                                \
      if (e_ == cudaErrorMemoryAllocation) {                                                       \
        (void)cudaGetLastError(); /* not sticky: the context stays usable */                        \
        throw ::st::OutOfDeviceMemory(buf_);                                                       \
      }                                                                                            \
      throw ::st::CudaError(buf_);                                                                 \
    }                                                                                              \
  } while (0)

// ---- Context --------------------------------------------------------------------------------
// With lazy module loading the first launch of a kernel pays for loading it (milliseconds); the
// reference shows the same first-use cost in its published N=128 row (README.md:70,146).  Touch
// the kernels the automatic choice uses when the handle is created, so that the first
// max_eigen_value() already returns steady-state loop times.
void
Context::preload_kernels()
{
  cudaFuncAttributes attr{};
  const void* kernels[] = {
    (const void*)round_loop_cluster_kernel<512, kStopAbsolute>,
    (const void*)round_loop_sc_kernel<512, 1, kStopAbsolute>,
    (const void*)round_loop_wide_kernel<512, kStopAbsolute>,
    (const void*)round_loop_kernel<4, kFormReadOnly, 512, kStopAbsolute>,
    (const void*)round_loop_kernel<1, kFormReadOnly, 512, kStopAbsolute>,
  };
  for (const void* k : kernels)
    ST_CUDA(cudaFuncGetAttributes(&attr, k));
}

Context::Context(int device)
  : device_(device)
{
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
    throw CudaError("no usable CUDA device");
  ST_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop{};
  ST_CUDA(cudaGetDeviceProperties(&prop, device_));
  if (prop.major < 10)
    throw CudaError(std::string("device is not sm_100-class: ") + prop.name);
  int coop = 0;
  ST_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_));
  if (!coop)
    throw CudaError("device does not support cooperative launch");
  sm_count_ = prop.multiProcessorCount;
  l2_bytes_ = (size_t)prop.l2CacheSize;
  hbm_bytes_ = prop.totalGlobalMem;
  name_ = prop.name;
  ST_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  ST_CUDA(cudaEventCreate(&ev0_));
  ST_CUDA(cudaEventCreate(&ev1_));
  ST_CUDA(cudaEventCreate(&ev_timer_[0]));
  ST_CUDA(cudaEventCreate(&ev_timer_[1]));
  ST_CUDA(cudaMalloc(&d_bar_, sizeof(BarrierState)));
  ST_CUDA(cudaMalloc(&d_scalars_, 64));
  ST_CUDA(cudaHostAlloc(&h_pinned_, 64, cudaHostAllocDefault));
  // Pageable host matrices (what the reference's wrapper passes: a numpy array) are staged by T host threads
  // through pinned double buffers.  Measured on a B200 box, Hilbert 8192 through max_eigen_value
  // (profiles/r2_c1_bench_upload.json): driver staging 24.6 ms, T = 2: 11.9, T = 4: 7.6, T = 8: 7.5, pinned 5.6.
  // ST_UPLOAD_THREADS overrides (0 = leave pageable sources to the driver).
  upload_threads_ = 4;
  if (const char* v = getenv("ST_UPLOAD_THREADS"))
    upload_threads_ = std::max(0, std::min(16, atoi(v)));
  load_local_cpus();
  preload_kernels();
}

// CPUs next to this GPU (sysfs: local_cpulist of its PCI function), intersected with the CPUs this process may
// use.  The staging threads of copy_h2d bind themselves to them, so that on a two-socket box the bounce buffers
// and the memcpy into them stay on the GPU's own NUMA node.  Best effort: an empty set means "do not bind".
void
Context::load_local_cpus()
{
  local_cpus_.clear();
  char bus[32] = {};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device_) != cudaSuccess) {
    (void)cudaGetLastError();
    return;
  }
  for (char* c = bus; *c; ++c)
    *c = (char)tolower((unsigned char)*c);
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f)
    return;
  char line[1024] = {};
  const bool ok = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!ok)
    return;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0)
    return;
  for (const char* q = line; *q && *q != '\n';) { // "0-23,48-71"
    char* end = nullptr;
    const long a = strtol(q, &end, 10);
    if (end == q)
      break;
    long b = a;
    if (*end == '-') {
      q = end + 1;
      b = strtol(q, &end, 10);
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++)
      if (c >= 0 && CPU_ISSET((int)c, &allowed))
        local_cpus_.push_back((int)c);
    q = *end == ',' ? end + 1 : end;
    if (*end != ',' )
      break;
  }
}

void
Context::bind_this_thread_near_device() const
{
  if (local_cpus_.empty())
    return;
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int c : local_cpus_)
    CPU_SET(c, &set);
  (void)pthread_setaffinity_np(pthread_self(), sizeof set, &set);
}

Context::~Context()
{
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  cudaFree(d_vec_);
  cudaFree(d_ts_);
  cudaFree(d_bar_);
  cudaFree(d_scalars_);
  cudaFree(d_mat_);
  cudaFree(d_work_);
  cudaFreeHost(h_pinned_);
  for (cudaEvent_t e : slot_ready_)
    cudaEventDestroy(e);
  for (cudaEvent_t e : slot_free_)
    cudaEventDestroy(e);
  if (copy_stream_)
    cudaStreamDestroy(copy_stream_);
  for (cudaStream_t st : up_streams_)
    cudaStreamDestroy(st);
  for (cudaEvent_t e : up_events_)
    cudaEventDestroy(e);
  cudaFreeHost(bounce_);
  cudaEventDestroy(ev0_);
  cudaEventDestroy(ev1_);
  cudaEventDestroy(ev_timer_[0]);
  cudaEventDestroy(ev_timer_[1]);
  cudaStreamDestroy(stream_);
}

void
Context::activate() const
{
  ST_CUDA(cudaSetDevice(device_));
}

void
Context::timer_start()
{
  activate();
  ST_CUDA(cudaEventRecord(ev_timer_[0], stream_));
}

float
Context::timer_stop()
{
  activate();
  ST_CUDA(cudaEventRecord(ev_timer_[1], stream_));
  ST_CUDA(cudaEventSynchronize(ev_timer_[1]));
  float ms = 0.f;
  ST_CUDA(cudaEventElapsedTime(&ms, ev_timer_[0], ev_timer_[1]));
  return ms;
}

// rounds beyond this are still executed but not time-stamped (bounds the stamp buffers when a
// caller raises max_iter far above the reference's 1000)
constexpr uint32_t kMaxStampedRounds = 1u << 16;

// Scratch a solve needs, as a function of the problem only (never of the kernel the planner ends up choosing):
// the five N-vectors, the stamp buffers, and either the in-place working copy or -- rows wider than one
// 8192-column work unit -- the chunk sums and arrival counters of the unit-scheduled kernels.
Context::ScratchNeed
Context::scratch_need(uint32_t dim, uint32_t rows, const st_options& opt) const
{
  ScratchNeed n;
  n.vec = (dim + 31u) & ~31u;
  n.stamps = std::min<uint32_t>(opt.max_iter, kMaxStampedRounds) + 2u;
  const size_t units = ((size_t)dim + (size_t)kChunkCols - 1u) / (size_t)kChunkCols;
  if (opt.form == ST_FORM_INPLACE) {
    n.work = (size_t)rows * dim;
  } else {
    if (units > 1) // chunk sums, per-row arrival counters and (wide kernel) one work-unit counter per window
      n.work = (size_t)rows * units + rows + 32u * (size_t)kWideMaxWindows;
  }
  return n;
}

bool
Context::prepared(uint32_t dim, uint32_t rows, const st_options& opt) const
{
  const ScratchNeed n = scratch_need(dim, rows, opt);
  return n.vec <= vec_cap_ && n.stamps <= ts_cap_ && n.work <= work_cap_;
}

void
Context::prepare(uint32_t dim, uint32_t rows, const st_options& opt)
{
  activate();
  const ScratchNeed n = scratch_need(dim, rows, opt);
  reserve_vectors(dim, opt.max_iter);
  if (n.work)
    reserve_work(n.work);
}

void
Context::reserve_vectors(uint32_t dim, uint32_t max_iter)
{
  const uint32_t need = (dim + 31u) & ~31u;
  if (need > vec_cap_) {
    cudaFree(d_vec_);
    d_vec_ = nullptr;
    vec_cap_ = 0;
    ST_CUDA(cudaMalloc(&d_vec_, sizeof(float) * 5 * (size_t)need));
    vec_cap_ = need;
  }
  const uint32_t stamped = std::min<uint32_t>(max_iter, kMaxStampedRounds);
  if (stamped + 2 > ts_cap_) {
    cudaFree(d_ts_);
    d_ts_ = nullptr;
    ts_cap_ = 0;
    // round stamps (stamped + 2) followed by three phase stamps per round
    ST_CUDA(cudaMalloc(&d_ts_, sizeof(unsigned long long) * 4 * (size_t)(stamped + 2)));
    ts_cap_ = stamped + 2;
  }
}

void
Context::reserve_matrix(size_t elems)
{
  if (elems > mat_cap_) {
    cudaFree(d_mat_);
    d_mat_ = nullptr;
    mat_cap_ = 0;
    ST_CUDA(cudaMalloc(&d_mat_, sizeof(float) * elems));
    mat_cap_ = elems;
  }
}

void
Context::reserve_work(size_t elems)
{
  if (elems > work_cap_) {
    cudaFree(d_work_);
    d_work_ = nullptr;
    work_cap_ = 0;
    ST_CUDA(cudaMalloc(&d_work_, sizeof(float) * elems));
    work_cap_ = elems;
  }
}

// ---- launch of the round loop ------------------------------------------------------------
// CTA size is a run-time value (a multiple of 32); the template parameter is only the
// __launch_bounds__ ceiling (512 -> 128 registers/thread, 1024 -> 64).
template<int VEC, int FORM, int MAX_THREADS, int STOP>
static void
launch_round_loop(const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  auto kernel = round_loop_kernel<VEC, FORM, MAX_THREADS, STOP>;
  if (smem > 48 * 1024)
    ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = { (void*)&p };
  ST_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, smem,
                                      stream));
}

template<int VEC, int FORM, int STOP>
static void
launch_by_threads(int threads, const RoundParams& p, int grid, size_t smem, cudaStream_t stream)
{
  if (threads <= 256)
    return launch_round_loop<VEC, FORM, 256, STOP>(p, grid, threads, smem, stream);
  if (threads <= 512)
    return launch_round_loop<VEC, FORM, 512, STOP>(p, grid, threads, smem, stream);
  return launch_round_loop<VEC, FORM, 1024, STOP>(p, grid, threads, smem, stream);
}

template<int VEC, int FORM>
static void
launch_by_stop(int stop, int threads, const RoundParams& p, int grid, size_t smem, cudaStream_t stream)
{
  if (stop == kStopRelative)
    return launch_by_threads<VEC, FORM, kStopRelative>(threads, p, grid, smem, stream);
  return launch_by_threads<VEC, FORM, kStopAbsolute>(threads, p, grid, smem, stream);
}

// ---- resident-e variant (N <= 32768): e in smem, fused tail, cross-barrier prefetch ----------
template<int MAX_THREADS, int PF, int STOP = kStopAbsolute>
static void
launch_sc_one(const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  auto kernel = round_loop_sc_kernel<MAX_THREADS, PF, STOP>;
  if (smem > 48 * 1024)
    ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = { (void*)&p };
  ST_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, smem,
                                      stream));
}

static void
launch_sc(int id, int stop, const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  if (stop == kStopRelative) {
    switch (id) {
      case 10: return launch_sc_one<512, 2, kStopRelative>(p, grid, threads, smem, stream);
      case 11: return launch_sc_one<512, 0, kStopRelative>(p, grid, threads, smem, stream);
      case 12: return launch_sc_one<512, 3, kStopRelative>(p, grid, threads, smem, stream);
      case 13: return launch_sc_one<512, 1, kStopRelative>(p, grid, threads, smem, stream);
      default: throw std::invalid_argument("unknown resident-e kernel id");
    }
  }
  switch (id) {
    case 10: return launch_sc_one<512, 2>(p, grid, threads, smem, stream);
    case 11: return launch_sc_one<512, 0>(p, grid, threads, smem, stream);
    case 12: return launch_sc_one<512, 3>(p, grid, threads, smem, stream);
    case 13: return launch_sc_one<512, 1>(p, grid, threads, smem, stream);
    default: throw std::invalid_argument("unknown resident-e kernel id");
  }
}

// ---- bf16 storage (read-only form, N % 4 == 0): resident-e kernel without prefetch slots for
// N <= 32768, the general chunked loop above it --------------------------------------------------
constexpr int kScBf16Id = 11; // {512 threads, no prefetch}: the one resident-e configuration built for bf16

template<int STOP>
static void
launch_sc_bf16(const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  auto kernel = round_loop_sc_kernel<512, 0, STOP, bf16_t>;
  if (smem > 48 * 1024)
    ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = { (void*)&p };
  ST_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, smem,
                                      stream));
}

template<int MAX_THREADS, int STOP>
static void
launch_general_bf16_one(const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  auto kernel = round_loop_kernel<4, kFormReadOnly, MAX_THREADS, STOP, bf16_t>;
  if (smem > 48 * 1024)
    ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = { (void*)&p };
  ST_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, smem,
                                      stream));
}

template<int STOP>
static void
launch_general_bf16(int threads, const RoundParams& p, int grid, size_t smem, cudaStream_t stream)
{
  if (threads <= 512)
    return launch_general_bf16_one<512, STOP>(p, grid, threads, smem, stream);
  return launch_general_bf16_one<1024, STOP>(p, grid, threads, smem, stream);
}

// ---- fp64 accumulation (fp32 storage, read-only form): the automatic resident-e configurations
// (13 / 10 / 12) and the general loop ----------------------------------------------------------------
template<typename K>
static void
launch_cooperative(K kernel, const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  if (smem > 48 * 1024)
    ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = { (void*)&p };
  ST_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(threads), args, smem, stream));
}

template<int STOP>
static void
launch_sc_acc64(int id, const RoundParams& p, int grid, int threads, size_t smem, cudaStream_t stream)
{
  switch (id) {
    case 10: return launch_cooperative(round_loop_sc_kernel<512, 2, STOP, float, double>, p, grid, threads, smem, stream);
    case 12: return launch_cooperative(round_loop_sc_kernel<512, 3, STOP, float, double>, p, grid, threads, smem, stream);
    case 13: return launch_cooperative(round_loop_sc_kernel<512, 1, STOP, float, double>, p, grid, threads, smem, stream);
    default: throw std::invalid_argument("fp64 accumulation is built for resident-e configurations 13, 10 and 12");
  }
}

template<int VEC, int STOP>
static void
launch_general_acc64(int threads, const RoundParams& p, int grid, size_t smem, cudaStream_t stream)
{
  if (threads <= 512)
    return launch_cooperative(round_loop_kernel<VEC, kFormReadOnly, 512, STOP, float, double>, p, grid, threads, smem, stream);
  return launch_cooperative(round_loop_kernel<VEC, kFormReadOnly, 1024, STOP, float, double>, p, grid, threads, smem, stream);
}

// ---- on-chip variant (N <= 512): matrix resident in the shared memory of one cluster -----------
constexpr int kClusterKernelId = 20;
constexpr int kClusterThreads = 512;

template<int STOP>
static void
launch_cluster(const RoundParams& p, int ctas, size_t smem, cudaStream_t stream)
{
  auto kernel = round_loop_cluster_kernel<kClusterThreads, STOP>;
  ST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ST_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
}

static bool
aligned16(const void* p)
{
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

int
Context::solve(const float* d_rows, uint32_t dim, const st_options& opt, Shard* shard,
               float* d_eigen_vec, st_result* res, bool bf16, const float* d_row_scale)
{
  if (!d_rows || dim == 0 || opt.max_iter == 0 || !(opt.eps >= 0.f))
    throw std::invalid_argument("solve: bad argument");
  // the round barrier counts arrivals in 32 bits: (max_iter + 1) * grid must not wrap
  if (opt.max_iter > (1u << 24))
    throw std::invalid_argument("solve: max_iter above 2^24 is not supported");
  if (shard && (shard->dim != dim || !shard->linked))
    throw std::invalid_argument("solve: shard not linked or dimension mismatch");
  const auto host_t0 = std::chrono::steady_clock::now();
  activate();
  const uint32_t row0 = shard ? shard->row0 : 0u;
  const uint32_t rows = shard ? shard->rows : dim;
  // A sharded solve is collective: peers may already spin in the round barrier, and an allocation on a device
  // with peer mappings may wait for them.  Its scratch was reserved before any rank launched (st_shard_create for
  // the default options, st_shard_prepare / upload_rows otherwise); refuse instead of allocating here.
  if (shard && shard->world > 1 && !prepared(dim, rows, opt))
    throw std::invalid_argument("solve: these options need more scratch than the shard was prepared for -- call "
                                "st_shard_prepare on every rank and synchronise the ranks before this solve");
  prepare(dim, rows, opt);

  const int form = opt.form == ST_FORM_INPLACE ? kFormInPlace : kFormReadOnly;
  if (opt.stop != ST_STOP_ABSOLUTE && opt.stop != ST_STOP_RELATIVE)
    throw std::invalid_argument("solve: unknown st_options.stop");
  const int stop = opt.stop == ST_STOP_RELATIVE ? kStopRelative : kStopAbsolute;
  if (opt.accumulate != ST_ACC_F32 && opt.accumulate != ST_ACC_F64)
    throw std::invalid_argument("solve: unknown st_options.accumulate");
  if (!is_known_kernel_id(opt.kernel))
    throw std::invalid_argument("solve: unknown st_options.kernel (0 automatic, 1 general loop, 2 wide, 10-13 resident-e, 20 on-chip cluster)");
  const bool acc64 = opt.accumulate == ST_ACC_F64;
  if (acc64) {
    if (bf16 || form != kFormReadOnly)
      throw std::invalid_argument("solve: fp64 accumulation needs fp32 storage and the read-only form");
    if (opt.kernel != 0 && opt.kernel != 1 && opt.kernel != 10 && opt.kernel != 12 && opt.kernel != 13)
      throw std::invalid_argument("solve: fp64 accumulation is built for kernel 0 (automatic), 1, 10, 12 and 13");
  }
  // fp8 storage: d_rows points to e4m3 codes, d_row_scale to one power-of-two scale per owned row
  const bool fp8 = d_row_scale != nullptr;
  const bool narrow = bf16 || fp8; // storage below fp32: configuration 11 or the general loop
  if (fp8) {
    if (bf16 || acc64 || form != kFormReadOnly || dim % 4u != 0u || (reinterpret_cast<uintptr_t>(d_rows) & 15u) != 0)
      throw std::invalid_argument("solve: fp8 storage needs the read-only form, fp32 accumulation, dim % 4 == 0 and a 16-byte aligned matrix");
    if (opt.kernel != 0 && opt.kernel != 1 && opt.kernel != kScBf16Id)
      throw std::invalid_argument("solve: fp8 storage is built for kernel 0 (automatic), 1 (general loop) and 11 (resident-e)");
    if (opt.kernel == kScBf16Id && dim > (uint32_t)kResidentCols)
      throw std::invalid_argument("solve: resident-e kernel needs dim <= 32768");
  }
  if (bf16) {
    // d_rows points to bfloat16 storage: 64-bit loads of 4 elements, fp32 everywhere else
    if (form != kFormReadOnly || dim % 4u != 0u || (reinterpret_cast<uintptr_t>(d_rows) & 15u) != 0)
      throw std::invalid_argument("solve: bf16 storage needs the read-only form, dim % 4 == 0 and a 16-byte aligned matrix");
    if (opt.kernel != 0 && opt.kernel != 1 && opt.kernel != kScBf16Id)
      throw std::invalid_argument("solve: bf16 storage is built for kernel 0 (automatic), 1 (general loop) and 11 (resident-e)");
    if (opt.kernel == kScBf16Id && dim > (uint32_t)kResidentCols)
      throw std::invalid_argument("solve: resident-e kernel needs dim <= 32768");
  }

  RoundParams p{};
  p.A = d_rows;
  p.row_scale = d_row_scale;
  p.W = nullptr;
  if (form == kFormInPlace)
    p.W = d_work_;
  p.N = dim;
  p.row0 = row0;
  p.rows = rows;
  float* v = d_vec_;
  const size_t cap = vec_cap_;
  p.S[0] = v;
  p.S[1] = v + cap;
  p.E[0] = v + 2 * cap;
  p.E[1] = v + 3 * cap;
  float* d_out_vec = d_eigen_vec ? d_eigen_vec : v + 4 * cap;
  p.eps = opt.eps;
  p.max_iter = opt.max_iter;
  p.sweep = opt.sweep & 1;
  // st_options.sweep bit 1 forces static, bit 2 forces dynamic unit scheduling; default: dynamic
  // once a unit is a full 32 KB chunk (N >= 8192), where the matrix no longer lives in L2
  p.dynamic = (opt.sweep & 4) ? 1 : (opt.sweep & 2) ? 0 : (dim >= (uint32_t)kChunkCols ? 1 : 0);
  p.keep_rows_pct = (uint32_t)std::max(0, std::min(100, opt.l2_keep_pct));
  p.chunk_cols = std::min<uint32_t>((uint32_t)kWindowCols, dim); // general loop: staged window of the scale vector
  p.bar = d_bar_;
  p.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
  p.rank = shard ? shard->rank : 0u;
  p.world = shard ? shard->world : 1u;
  if (shard) {
    for (uint32_t g = 0; g < shard->world; g++) {
      char* base = static_cast<char*>(shard->peer_block[g]);
      p.peer_S[0][g] = reinterpret_cast<float*>(base + shard->s_offset[0]);
      p.peer_S[1][g] = reinterpret_cast<float*>(base + shard->s_offset[1]);
      ExchangeHeader* h = reinterpret_cast<ExchangeHeader*>(base);
      p.peer_arrive[g] = &h->arrive;
      p.peer_smax3[g] = h->smax3;
    }
    p.S[0] = p.peer_S[0][shard->rank];
    p.S[1] = p.peer_S[1][shard->rank];
    // the barrier words are never reset: a solve starts from the totals its group has reached; S buffers start from
    // the parity the last solve did NOT end on
    shard->solves += 1;
    p.arrive_base = shard->arrive_total;
    p.round_base = shard->rounds_total;
    p.flip = shard->flip;
    if (opt.max_iter >= (1u << 24) - 1u)
      throw std::invalid_argument("solve: a sharded solve supports max_iter below 2^24 - 1");
  }
  p.out_eigen_vec = d_out_vec;
  p.out_eigen_val = d_scalars_;
  p.out_iter = reinterpret_cast<uint32_t*>(d_scalars_ + 1);
  p.round_ts = d_ts_;
  p.phase_ts = d_ts_ + ts_cap_;
  // st_options.sweep bit 3 switches the per-round instrumentation off (no %globaltimer reads)
  p.ts_rounds = (opt.sweep & 8) ? 0u : std::min<uint32_t>(opt.max_iter, kMaxStampedRounds);

  // ---- launch plan ------------------------------------------------------------------------
  // one persistent CTA per SM (fewer when there are fewer rows than warps); CTA size = the warp
  // count that divides the CTA's rows best unless st_options.threads pins it.
  // kernel: 0 = automatic (on-chip cluster kernel for N <= 512, resident-e kernel when N <= 32768, else the
  // general chunked loop), 1 = general loop, 10-13 = resident-e configurations, 20 = on-chip cluster kernel.
  const bool vec4 = (dim % 4u == 0u) && aligned16(d_rows) && (!p.W || aligned16(p.W));
  const int pinned = opt.threads > 0 ? std::min(1024, (opt.threads + 31) / 32 * 32) : 0;
  auto shape = [&](int max_threads, bool balance, int* out_grid, int* out_threads, uint32_t* out_cap) {
    const int hi = (pinned ? std::min(pinned, max_threads) : max_threads) / 32;
    int g = opt.ctas > 0 ? opt.ctas : sm_count_;
    const int use = (int)std::max<uint32_t>(1u, (rows + (uint32_t)hi - 1u) / (uint32_t)hi);
    g = std::max(1, std::min(g, std::min(use, sm_count_)));
    const uint32_t per_cta = (rows + (uint32_t)g - 1u) / (uint32_t)g;
    // the resident-e kernel schedules rows dynamically and always wants every warp
    const int lo = (pinned || !balance) ? hi : std::max(1, std::min(hi, max_threads >= 512 ? 12 : 6));
    *out_grid = g;
    *out_threads = 32 * balanced_warps(per_cta, lo, hi);
    *out_cap = per_cta + 1u;
  };

  int threads = 512, grid = sm_count_;
  uint32_t rows_cap = 0;
  size_t smem = 0;
  const ScConfig* sc = nullptr;
  const bool readonly4 = vec4 && form == kFormReadOnly;
  int cluster_ctas = 0;
  if (opt.kernel == kClusterKernelId || (opt.kernel == 0 && readonly4 && !shard && !narrow && !acc64 && dim <= (uint32_t)kClusterCols)) {
    if (!readonly4 || shard || dim > (uint32_t)kClusterCols)
      throw std::invalid_argument("solve: on-chip kernel needs one GPU, the read-only form, dim % 4 == 0, dim <= 512");
    cluster_ctas = cluster_ctas_for(dim, &smem);
    if (!cluster_ctas)
      throw std::invalid_argument("solve: matrix does not fit the cluster's shared memory");
    grid = cluster_ctas;
    threads = kClusterThreads;
  }
  // scalar units (dim % 4 != 0 or a matrix off the 16-byte grid): the resident-e kernel without prefetch slots
  // (configuration 11) streams them with 4-byte loads in the order the general loop uses for these dimensions
  const bool readonly1 = !vec4 && form == kFormReadOnly && !narrow && !acc64;
  if (!cluster_ctas && (is_sc_kernel_id(opt.kernel) ||
                        (opt.kernel == 0 && (readonly4 || readonly1) && dim <= (uint32_t)kResidentCols))) {
    if (!(readonly4 || (readonly1 && (opt.kernel == 0 || opt.kernel == kScBf16Id))) || dim > (uint32_t)kResidentCols)
      throw std::invalid_argument("solve: resident-e kernel needs the read-only form and dim <= 32768; dim % 4 != 0 "
                                  "runs on configuration 11 (fp32 storage and accumulation)");
    // Automatic slot size: one 4 KB batch per warp, unless a whole row fits a 2- or 3-batch slot
    // and no warp would own more than one row -- then the rows stay in shared memory for the
    // whole solve (matrix resident on chip, N <= ~2368 on 148 SMs).
    int want_pf = 1;
    if (opt.kernel == 0 && dim > 1024u && dim <= 3072u && (uint64_t)rows <= (uint64_t)sm_count_ * 16u)
      want_pf = (int)((dim + 1023u) / 1024u);
    if (narrow || readonly1)
      want_pf = 0; // bf16 / fp8 storage and scalar units are built without prefetch slots (configuration 11)
    for (const ScConfig& c : kScConfigs) {
      if (opt.kernel >= 10 && c.id != opt.kernel)
        continue;
      if (opt.kernel == 0 && pinned > c.max_threads)
        continue;
      if ((narrow || readonly1) && c.id != kScBf16Id)
        continue; // bf16 / fp8 storage and scalar units are built for configuration 11 only
      if (opt.kernel == 0 && !narrow && acc64 && c.id == kScBf16Id)
        continue; // fp64 accumulation is built for the configurations with prefetch slots
      if (opt.kernel == 0 && !pinned && (c.pf_batches != want_pf || c.max_threads != 512))
        continue;
      int g, t;
      uint32_t cap, moff = 0;
      shape(c.max_threads, false, &g, &t, &cap);
      const size_t need = sc_smem_bytes(t, c.pf_batches, dim, &moff);
      if (need <= kSmemLimit) {
        sc = &c;
        grid = g;
        threads = t;
        rows_cap = cap;
        smem = need;
        p.mbar_offset = moff;
        p.chunk_cols = dim; // the whole eigenvector is resident
        const uint32_t units = (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols;
        if (units > 1u) {
          // chunk sums + per-row arrival counters of rows that span several work units
          p.partial = d_work_;
          p.row_done = reinterpret_cast<unsigned int*>(d_work_ + (size_t)rows * units);
          ST_CUDA(cudaMemsetAsync(p.row_done, 0, sizeof(unsigned int) * rows, stream_));
        }
        break;
      }
    }
    if (!sc && opt.kernel >= 10)
      throw std::invalid_argument("solve: requested resident-e kernel configuration does not fit");
  }
  // wide kernel (kernel 2): unit-scheduled like the resident-e kernel, the eigenvector staged one 32768-column window at
  // a time.  Automatic above the resident limit for the default form / storage / accumulator; explicit for any dim % 4 == 0.
  bool wide = false;
  if (!sc && !cluster_ctas && (opt.kernel == kWideKernelId || (opt.kernel == 0 && readonly4 && !narrow && !acc64 &&
                                                                dim > (uint32_t)kResidentCols && pinned <= 512))) {
    if (!readonly4 || narrow || acc64)
      throw std::invalid_argument("solve: the wide kernel needs the read-only form, fp32 storage and accumulation, dim % 4 == 0");
    const uint32_t window = std::min<uint32_t>((uint32_t)kResidentCols, (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols * (uint32_t)kChunkCols);
    if ((dim + window - 1u) / window > (uint32_t)kWideMaxWindows)
      throw std::invalid_argument("solve: the wide kernel supports up to 64 windows of 32768 columns");
    uint32_t moff = 0;
    shape(512, false, &grid, &threads, &rows_cap);
    smem = sc_smem_bytes(threads, 1, std::min<uint32_t>(window, dim), &moff);
    p.mbar_offset = moff;
    p.chunk_cols = window;
    const uint32_t units = (dim + (uint32_t)kChunkCols - 1u) / (uint32_t)kChunkCols;
    // a row of ONE unit still goes through the chunk-sum protocol here: reserve for it (small matrices, explicit kernel 2)
    const size_t need_work = (size_t)rows * units + rows + 32u * (size_t)kWideMaxWindows;
    if (need_work > work_cap_) {
      if (shard && shard->world > 1)
        throw std::invalid_argument("solve: the wide kernel's scratch was not prepared for this shard (st_shard_prepare)");
      reserve_work(need_work);
    }
    p.partial = d_work_;
    p.row_done = reinterpret_cast<unsigned int*>(d_work_ + (size_t)rows * units);
    p.phase_counter = p.row_done + rows;
    ST_CUDA(cudaMemsetAsync(p.row_done, 0, sizeof(unsigned int) * ((size_t)rows + 32u * (size_t)kWideMaxWindows), stream_));
    wide = true;
  }
  if (!sc && !cluster_ctas && !wide) {
    shape(pinned > 512 ? 1024 : 512, true, &grid, &threads, &rows_cap);
    smem = sizeof(float) * ((size_t)p.chunk_cols + rows_cap);
  }

  ST_CUDA(cudaMemsetAsync(d_bar_, 0, sizeof(BarrierState), stream_));
  ST_CUDA(cudaMemsetAsync(d_scalars_, 0, 64, stream_));
  ST_CUDA(cudaEventRecord(ev0_, stream_));
  if (acc64 && sc) {
    if (stop == kStopRelative)
      launch_sc_acc64<kStopRelative>(sc->id, p, grid, threads, smem, stream_);
    else
      launch_sc_acc64<kStopAbsolute>(sc->id, p, grid, threads, smem, stream_);
  } else if (acc64) {
    if (cluster_ctas)
      throw std::invalid_argument("solve: fp64 accumulation is not built for this kernel");
    if (vec4 && stop == kStopRelative)
      launch_general_acc64<4, kStopRelative>(threads, p, grid, smem, stream_);
    else if (vec4)
      launch_general_acc64<4, kStopAbsolute>(threads, p, grid, smem, stream_);
    else if (stop == kStopRelative)
      launch_general_acc64<1, kStopRelative>(threads, p, grid, smem, stream_);
    else
      launch_general_acc64<1, kStopAbsolute>(threads, p, grid, smem, stream_);
  } else if (fp8 && sc) {
    if (sc->id != kScBf16Id)
      throw std::invalid_argument("solve: fp8 storage needs resident-e configuration 11");
    if (stop == kStopRelative)
      launch_cooperative(round_loop_sc_kernel<512, 0, kStopRelative, fp8_t>, p, grid, threads, smem, stream_);
    else
      launch_cooperative(round_loop_sc_kernel<512, 0, kStopAbsolute, fp8_t>, p, grid, threads, smem, stream_);
  } else if (fp8) {
    if (threads <= 512 && stop == kStopRelative)
      launch_cooperative(round_loop_kernel<4, kFormReadOnly, 512, kStopRelative, fp8_t>, p, grid, threads, smem, stream_);
    else if (threads <= 512)
      launch_cooperative(round_loop_kernel<4, kFormReadOnly, 512, kStopAbsolute, fp8_t>, p, grid, threads, smem, stream_);
    else if (stop == kStopRelative)
      launch_cooperative(round_loop_kernel<4, kFormReadOnly, 1024, kStopRelative, fp8_t>, p, grid, threads, smem, stream_);
    else
      launch_cooperative(round_loop_kernel<4, kFormReadOnly, 1024, kStopAbsolute, fp8_t>, p, grid, threads, smem, stream_);
  } else if (bf16 && sc) {
    if (sc->id != kScBf16Id)
      throw std::invalid_argument("solve: bf16 storage needs resident-e configuration 11");
    if (stop == kStopRelative)
      launch_sc_bf16<kStopRelative>(p, grid, threads, smem, stream_);
    else
      launch_sc_bf16<kStopAbsolute>(p, grid, threads, smem, stream_);
  } else if (bf16) {
    if (stop == kStopRelative)
      launch_general_bf16<kStopRelative>(threads, p, grid, smem, stream_);
    else
      launch_general_bf16<kStopAbsolute>(threads, p, grid, smem, stream_);
  } else if (cluster_ctas) {
    if (stop == kStopRelative)
      launch_cluster<kStopRelative>(p, cluster_ctas, smem, stream_);
    else
      launch_cluster<kStopAbsolute>(p, cluster_ctas, smem, stream_);
  } else if (sc && readonly1) {
    if (stop == kStopRelative)
      launch_cooperative(round_loop_sc_kernel<512, 0, kStopRelative, float, float, 1>, p, grid, threads, smem, stream_);
    else
      launch_cooperative(round_loop_sc_kernel<512, 0, kStopAbsolute, float, float, 1>, p, grid, threads, smem, stream_);
  } else if (sc) {
    launch_sc(sc->id, stop, p, grid, threads, smem, stream_);
  } else if (wide) {
    if (stop == kStopRelative)
      launch_cooperative(round_loop_wide_kernel<512, kStopRelative>, p, grid, threads, smem, stream_);
    else
      launch_cooperative(round_loop_wide_kernel<512, kStopAbsolute>, p, grid, threads, smem, stream_);
  } else if (vec4) {
    if (form == kFormInPlace)
      launch_by_stop<4, kFormInPlace>(stop, threads, p, grid, smem, stream_);
    else
      launch_by_stop<4, kFormReadOnly>(stop, threads, p, grid, smem, stream_);
  } else {
    if (form == kFormInPlace)
      launch_by_stop<1, kFormInPlace>(stop, threads, p, grid, smem, stream_);
    else
      launch_by_stop<1, kFormReadOnly>(stop, threads, p, grid, smem, stream_);
  }
  ST_CUDA(cudaEventRecord(ev1_, stream_));

  // read back: lambda, iter_count, passes, barrier error word
  float* hp = static_cast<float*>(h_pinned_);
  ST_CUDA(cudaMemcpyAsync(hp, d_scalars_, 16, cudaMemcpyDeviceToHost, stream_));
  ST_CUDA(cudaMemcpyAsync(hp + 4, &d_bar_->error, 4, cudaMemcpyDeviceToHost, stream_));
  ST_CUDA(cudaStreamSynchronize(stream_));
  float loop_ms = 0.f;
  ST_CUDA(cudaEventElapsedTime(&loop_ms, ev0_, ev1_));

  uint32_t words[5];
  memcpy(words, hp, sizeof words);
  if (words[4] != 0u) {
    set_last_error(words[4] == 2u ? "bulk-copy wait timed out inside the round kernel"
                                  : "round barrier timed out (a peer rank did not arrive)");
    if (res)
      res->status = ST_ERR_TIMEOUT;
    return ST_ERR_TIMEOUT;
  }
  const uint32_t passes = words[2];
  if (shard) {
    shard->flip = (shard->flip + passes) & 1u; // every rank saw the same number of rounds
    shard->arrive_total += (uint64_t)passes * shard->world * kArriveUnits;
    shard->rounds_total += passes;
  }
  const uint32_t stamped = std::min<uint32_t>(passes, p.ts_rounds);
  last_ts_.assign((size_t)stamped + 1, 0);
  ST_CUDA(cudaMemcpy(last_ts_.data(), d_ts_, sizeof(uint64_t) * ((size_t)stamped + 1),
                     cudaMemcpyDeviceToHost));
  last_phase_ts_.assign(3 * (size_t)stamped, 0);
  ST_CUDA(cudaMemcpy(last_phase_ts_.data(), d_ts_ + ts_cap_, sizeof(uint64_t) * 3 * (size_t)stamped,
                     cudaMemcpyDeviceToHost));

  if (res) {
    memset(res, 0, sizeof *res);
    memcpy(&res->eigen_val, &words[0], 4);
    res->iter_count = words[1];
    res->passes = passes;
    res->launches = 1;
    res->loop_ms = loop_ms;
    res->grid = (uint32_t)grid;
    res->kernel_id = cluster_ctas ? (uint32_t)kClusterKernelId : sc ? (uint32_t)sc->id : wide ? (uint32_t)kWideKernelId : 1u;
    res->threads = (uint32_t)threads;
    const uint64_t per_pass = (uint64_t)rows * dim * (fp8 ? 1u : bf16 ? 2u : sizeof(float));
    res->bytes_per_round = form == kFormInPlace ? 2 * per_pass : per_pass;
    std::vector<float> dt;
    for (uint32_t k = 0; k < stamped; k++)
      dt.push_back((float)(last_ts_[k + 1] - last_ts_[k]) * 1e-3f);
    if (!dt.empty()) {
      std::sort(dt.begin(), dt.end());
      res->round_us_min = dt.front();
      res->round_us_median = dt[dt.size() / 2];
    }
    res->status = ST_OK;
    const auto host_t1 = std::chrono::steady_clock::now();
    res->total_ms = std::chrono::duration<float, std::milli>(host_t1 - host_t0).count();
  }
  return ST_OK;
}

// Host-matrix solve in two steps, so that a device group can finish every allocation on every GPU before
// any GPU enters the collective round kernel (cudaMalloc / cudaFree on a device with peer mappings may wait
// for its peers, and a peer that already spins in the round barrier would never become idle).
void
Context::upload_rows(const float* h_mat, uint32_t dim, const st_options& opt, Shard* shard)
{
  if (!h_mat || dim == 0)
    throw std::invalid_argument("solve_host: bad argument");
  if (shard && (shard->ctx != this || shard->dim != dim))
    throw std::invalid_argument("solve_host: shard belongs to another context or dimension");
  activate();
  // sharded (device group): h_mat is the WHOLE matrix, this context copies and solves its own row block
  const uint32_t row0 = shard ? shard->row0 : 0u;
  const uint32_t rows = shard ? shard->rows : dim;
  const size_t elems = (size_t)rows * dim;
  reserve_matrix(elems);
  prepare(dim, rows, opt);
  // the caller's matrix is never modified (reference similarity_transform.cpp:14,19 copies it)
  copy_h2d(d_mat_, h_mat + (size_t)row0 * dim, sizeof(float) * elems);
}

// Host -> device copy of a caller-owned matrix on the solver stream.  Pinned / registered memory goes
// straight to the copy engine.  A PAGEABLE source (the reference wrapper's numpy array) is staged by
// the driver through its own bounce buffer on one thread, at a fraction of the PCIe rate; with
// ST_UPLOAD_THREADS=T (opt-in, read when the context is created) T host threads copy 4 MiB chunks into
// their own pinned double buffers and feed the copy engine from there, so staging and DMA overlap and
// the host-side memcpy is spread over T cores.
void
Context::upload(void* d_dst, const void* h_src, size_t bytes)
{
  activate();
  copy_h2d(static_cast<float*>(d_dst), static_cast<const float*>(h_src), bytes);
  ST_CUDA(cudaStreamSynchronize(stream_));
}

void
Context::copy_h2d(float* d_dst, const float* h_src, size_t bytes, cudaStream_t stream)
{
  if (!stream)
    stream = stream_;
  constexpr size_t kChunk = 4ull << 20;
  bool staged = upload_threads_ > 0 && bytes >= 8 * kChunk;
  if (staged) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, h_src) != cudaSuccess) {
      (void)cudaGetLastError();
      staged = false;
    } else {
      staged = attr.type == cudaMemoryTypeUnregistered;
    }
  }
  if (!staged) {
    ST_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
    return;
  }
  const int T = upload_threads_;
  if (!bounce_) {
    ST_CUDA(cudaHostAlloc(&bounce_, 2 * kChunk * (size_t)T, cudaHostAllocDefault));
    for (int t = 0; t < T; t++) {
      cudaStream_t st = nullptr;
      ST_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      up_streams_.push_back(st);
      for (int b = 0; b < 3; b++) { // two buffer events + one "this thread's copies are done"
        cudaEvent_t ev = nullptr;
        ST_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        up_events_.push_back(ev);
      }
    }
    up_used_.assign(2 * (size_t)T, 0);
    cudaEvent_t start = nullptr; // last entry: "the solver stream has reached this upload"
    ST_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    up_events_.push_back(start);
  }
  // the destination may still be read by earlier work on the target stream
  cudaEvent_t up_start = up_events_.back();
  ST_CUDA(cudaEventRecord(up_start, stream));
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  const char* src = reinterpret_cast<const char*>(h_src);
  char* dst = reinterpret_cast<char*>(d_dst);
  std::vector<cudaError_t> errs((size_t)T, cudaSuccess);
  auto work = [&](int t) {
    if (t > 0)
      bind_this_thread_near_device(); // worker threads only: the caller's own thread keeps its affinity
    cudaError_t e = cudaSetDevice(device_);
    cudaStream_t st = up_streams_[t];
    char* buf = static_cast<char*>(bounce_) + 2 * kChunk * (size_t)t;
    if (e == cudaSuccess)
      e = cudaStreamWaitEvent(st, up_start, 0);
    int n = 0;
    for (size_t i = (size_t)t; e == cudaSuccess && i < nchunks; i += (size_t)T, n++) {
      const int b = n & 1;
      const size_t len = std::min(kChunk, bytes - i * kChunk);
      if (n >= 2 || up_used_[2 * t + b]) // also across calls: the streamed solve uploads block after block
        e = cudaEventSynchronize(up_events_[3 * t + b]); // the DMA out of this buffer has finished
      if (e != cudaSuccess)
        break;
      memcpy(buf + kChunk * b, src + i * kChunk, len);
      e = cudaMemcpyAsync(dst + i * kChunk, buf + kChunk * b, len, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess)
        e = cudaEventRecord(up_events_[3 * t + b], st);
      up_used_[2 * t + b] = 1;
    }
    if (e == cudaSuccess)
      e = cudaEventRecord(up_events_[3 * t + 2], st);
    errs[t] = e;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; t++)
    pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool)
    th.join();
  for (int t = 0; t < T; t++)
    ST_CUDA(errs[t]);
  for (int t = 0; t < T; t++)
    ST_CUDA(cudaStreamWaitEvent(stream, up_events_[3 * t + 2], 0));
  staged_bytes_ += bytes;
  // the bounce buffers are reused by the next upload: that one starts by waiting for up_start on its target
  // stream; the callers' streams are ordered after these copies before the buffers can be reused (solve_streamed
  // synchronises per round, solve_host per call)
}

int
Context::solve_uploaded(uint32_t dim, const st_options& opt, float* h_eigen_val, float* h_eigen_vec, st_result* res,
                        Shard* shard)
{
  st_result local{};
  const int rc = solve(d_mat_, dim, opt, shard, nullptr, &local);
  if (rc == ST_OK) {
    if (h_eigen_vec)
      ST_CUDA(cudaMemcpy(h_eigen_vec, d_vec_ + 4 * (size_t)vec_cap_, sizeof(float) * dim,
                         cudaMemcpyDeviceToHost));
    if (h_eigen_val)
      *h_eigen_val = local.eigen_val;
  }
  if (res)
    *res = local;
  return rc;
}

int
Context::solve_host(const float* h_mat, uint32_t dim, const st_options& opt, float* h_eigen_val,
                    float* h_eigen_vec, st_result* res, Shard* shard)
{
  const auto host_t0 = std::chrono::steady_clock::now();
  upload_rows(h_mat, dim, opt, shard);
  st_result local{};
  const int rc = solve_uploaded(dim, opt, h_eigen_val, h_eigen_vec, &local, shard);
  local.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
  if (res)
    *res = local;
  return rc;
}

} // namespace st

// =============================================================================================
// reference-named C++ entry points
// =============================================================================================
using namespace st;

int64_t
similarity_transform(st::Context& q, const float* mat, float* const eigen_val,
                     float* const eigen_vec, const uint dim, const uint /*wg_size*/,
                     uint* const iter_count)
{
  std::lock_guard<std::mutex> lock(q.mutex());
  st_options opt;
  st_default_options(&opt);
  st_result res{};
  const int rc = q.solve_host(mat, dim, opt, eigen_val, eigen_vec, &res);
  if (rc != ST_OK)
    return rc;
  if (iter_count)
    *iter_count = res.iter_count; // exactly 4 bytes
  return (int64_t)res.loop_ms;    // whole milliseconds of the loop, like reference :56-58
}

static int
blocks_for(uint32_t n, int threads, int cap)
{
  return (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + threads - 1) / threads, (uint32_t)cap));
}

static int
launch_row_pass(st::Context& q, const float* d_rows, const float* d_e, float* d_vec, uint32_t dim,
                uint32_t row0, uint32_t rows)
{
  q.activate();
  const bool vec4 = (dim % 4u == 0u) && aligned16(d_rows);
  const uint32_t chunk = std::min<uint32_t>((uint32_t)kChunkCols, dim);
  const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((rows + 7u) / 8u, 8u * q.sm_count()));
  if (vec4)
    ST_LAUNCH(sum_across_rows_kernel<4>, grid, 256, chunk * sizeof(float), q.stream(), d_rows, d_e, d_vec, dim, row0, rows);
  else
    ST_LAUNCH(sum_across_rows_kernel<1>, grid, 256, chunk * sizeof(float), q.stream(), d_rows, d_e, d_vec, dim, row0, rows);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
sum_across_rows(st::Context& q, const float* d_mat, float* d_vec, const uint dim, const uint)
{
  return launch_row_pass(q, d_mat, nullptr, d_vec, dim, 0, dim);
}

int
row_pass_readonly(st::Context& q, const float* d_rows, const float* d_e, float* d_vec, const uint dim,
                  const uint row0, const uint rows)
{
  return launch_row_pass(q, d_rows, d_e, d_vec, dim, row0, rows);
}

int
find_max(st::Context& q, const float* d_vec, float* d_max, const uint dim, const uint)
{
  q.activate();
  ST_CUDA(cudaMemsetAsync(d_max, 0, sizeof(float), q.stream())); // the zero-filled cell, reference :162-170
  ST_LAUNCH(find_max_kernel, blocks_for(dim, 1024, 4 * q.sm_count()), 1024, 0, q.stream(), d_vec, d_max, dim);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
compute_eigen_vector(st::Context& q, const float* d_vec, const float* d_max, float* d_eigen_vec,
                     const uint dim, const uint)
{
  q.activate();
  ST_LAUNCH_ELEMENTWISE(compute_eigen_vector_kernel, blocks_for(dim, 256, 4 * q.sm_count()), 256, 0, q.stream(), d_vec, d_max, d_eigen_vec, dim);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
initialise_eigen_vector(st::Context& q, float* d_eigen_vec, const uint dim)
{
  q.activate();
  ST_LAUNCH_ELEMENTWISE(fill_kernel, blocks_for(dim, 256, 4 * q.sm_count()), 256, 0, q.stream(), d_eigen_vec, 1.f, dim);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
compute_next_matrix(st::Context& q, float* d_mat, const float* d_vec, const uint dim, const uint)
{
  q.activate();
  const bool vec4 = (dim % 4u == 0u) && aligned16(d_mat) && aligned16(d_vec);
  const dim3 grid(blocks_for(vec4 ? dim / 4 : dim, 256, 64),
                  (unsigned)std::min<uint32_t>(dim, 16384u));
  if (vec4)
    ST_LAUNCH_ELEMENTWISE(compute_next_matrix_kernel<4>, grid, 256, 0, q.stream(), d_mat, d_vec, dim);
  else
    ST_LAUNCH_ELEMENTWISE(compute_next_matrix_kernel<1>, grid, 256, 0, q.stream(), d_mat, d_vec, dim);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
stop(st::Context& q, const float* d_vec, uint* d_ret, const uint dim, const uint, float eps)
{
  q.activate();
  ST_LAUNCH_ELEMENTWISE(fill_u32_kernel, 1, 32, 0, q.stream(), d_ret, 1u, 1u); // the flag starts at 1, reference :351-359
  ST_LAUNCH(stop_kernel, blocks_for(dim, 1024, 4 * q.sm_count()), 1024, 0, q.stream(), d_vec, d_ret, dim, eps);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
generate_hilbert_matrix(st::Context& q, float* d_rows, const uint dim, const uint row0, const uint rows_)
{
  q.activate();
  const uint rows = rows_ ? rows_ : dim - row0;
  const dim3 grid(blocks_for(dim, 256, 32), (unsigned)std::min<uint32_t>(rows, 32768u));
  ST_LAUNCH_ELEMENTWISE(hilbert_kernel, grid, 256, 0, q.stream(), d_rows, dim, row0, rows);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
convert_to_bf16(st::Context& q, const float* d_src, uint16_t* d_dst, size_t count)
{
  q.activate();
  const size_t work = (count + 3) / 4;
  const int grid = (int)std::min<size_t>((work + 255) / 256, (size_t)q.sm_count() * 16);
  ST_LAUNCH_ELEMENTWISE(convert_bf16_kernel, std::max(1, grid), 256, 0, q.stream(), d_src, d_dst, count);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
convert_to_fp8(st::Context& q, const float* d_src, uint8_t* d_dst, float* d_row_scale, uint32_t rows, uint32_t dim)
{
  q.activate();
  const int grid = (int)std::min<uint32_t>((rows + 7u) / 8u, (uint32_t)q.sm_count() * 8u);
  ST_LAUNCH(convert_fp8_rows_kernel, std::max(1, grid), 256, 0, q.stream(), d_src, reinterpret_cast<unsigned char*>(d_dst),
            d_row_scale, rows, dim);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

int
generate_uniform_matrix(st::Context& q, float* d_rows, const uint dim, uint64_t seed, const uint row0,
                        const uint rows_)
{
  q.activate();
  const uint rows = rows_ ? rows_ : dim - row0;
  const size_t blocks4 = ((size_t)rows * dim + 3) / 4 + 1;
  const int grid = (int)std::min<size_t>((blocks4 + 255) / 256, (size_t)q.sm_count() * 16);
  ST_LAUNCH_ELEMENTWISE(uniform_kernel, std::max(1, grid), 256, 0, q.stream(), d_rows, dim, row0, rows, seed);
  ST_CUDA(cudaGetLastError());
  return ST_OK;
}

// =============================================================================================
// streamed solve: host matrix larger than the device cache (include/similarity_transform.h)
// =============================================================================================
// The reference copies the whole matrix to the device before its loop (similarity_transform.cpp:14-19);
// when it does not fit, this keeps `slots` row blocks in a direct-mapped device cache (block b <-> slot
// b % slots) and sweeps the blocks in alternating direction, so each round starts on the blocks the last
// one ended on.  Per round: the cached blocks are reduced from HBM in contiguous runs, the others are
// copied on copy_stream_ (each copy waits for the pass that last read its slot) and reduced as they land.
// One round = row passes (sum_across_rows_kernel, the fused kernels' evaluation order) + tail_scan_kernel
// + tail_update_kernel + an 8-byte read-back; the host decides about the next round like the reference's
// host loop does (:44-50).
namespace {
constexpr uint32_t kStreamedKernelId = 30;
constexpr size_t kStreamBlockBytes = 64ull << 20;   // automatic block size
constexpr size_t kStreamHitRunBytes = 1024ull << 20; // cached blocks are reduced in runs of at most this
constexpr size_t kStreamReserveBytes = 1024ull << 20; // left free when the budget is taken from cudaMemGetInfo
} // namespace

int
st::Context::solve_streamed(const float* h_mat, uint32_t dim, const st_options& opt, size_t device_budget,
                            uint32_t block_rows, float* h_eigen_val, float* h_eigen_vec, st_result* res,
                            st_stream_plan* plan)
{
  if (!h_mat || dim == 0 || opt.max_iter == 0 || !(opt.eps >= 0.f))
    throw std::invalid_argument("solve_streamed: bad argument");
  if (opt.form != ST_FORM_READONLY || opt.accumulate != ST_ACC_F32)
    throw std::invalid_argument("solve_streamed: read-only form and fp32 accumulation only");
  if (opt.stop != ST_STOP_ABSOLUTE && opt.stop != ST_STOP_RELATIVE)
    throw std::invalid_argument("solve_streamed: unknown st_options.stop");
  const int stop_kind = opt.stop == ST_STOP_RELATIVE ? kStopRelative : kStopAbsolute;
  const auto host_t0 = std::chrono::steady_clock::now();
  activate();
  const size_t row_bytes = sizeof(float) * (size_t)dim;
  const uint64_t mat_bytes = (uint64_t)row_bytes * dim;
  if (plan)
    memset(plan, 0, sizeof *plan);

  size_t budget = device_budget;
  if (budget == 0) {
    size_t free_b = 0, total_b = 0;
    ST_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = free_b + sizeof(float) * mat_cap_; // the staging buffer is reused for the cache
    budget = avail > kStreamReserveBytes ? avail - kStreamReserveBytes : 0;
    if (mat_bytes <= budget) { // it fits: the fused device-side loop is the better path
      const int rc = solve_host(h_mat, dim, opt, h_eigen_val, h_eigen_vec, res);
      if (plan && rc == ST_OK) {
        plan->block_rows = dim;
        plan->blocks = plan->slots = 1;
        plan->cache_bytes = plan->h2d_bytes_first = plan->h2d_bytes_total = mat_bytes;
      }
      return rc;
    }
  }

  // ---- plan: rows per block, blocks, cache slots (launch_plan.hpp) ----
  StreamShape shape{};
  if (!stream_shape_for(dim, budget, block_rows, kStreamBlockBytes, &shape))
    throw std::invalid_argument("solve_streamed: the device budget holds fewer than two row blocks");
  const uint32_t B = shape.block_rows, nb = shape.blocks, C = shape.slots;
  const size_t block_bytes = row_bytes * B;
  reserve_matrix((size_t)C * B * dim);
  reserve_vectors(dim, opt.max_iter);
  if (!copy_stream_)
    ST_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
  while (slot_ready_.size() < C) {
    cudaEvent_t a = nullptr, b = nullptr;
    ST_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    slot_ready_.push_back(a);
    ST_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    slot_free_.push_back(b);
  }

  float* S = d_vec_;
  float* E = d_vec_ + 2 * (size_t)vec_cap_;
  uint32_t* cells = reinterpret_cast<uint32_t*>(d_scalars_) + 8; // [0] max bits, [1] stop evidence
  uint32_t* d_out = cells + 2;                                   // [0] converged, [1] bits of s[0]
  volatile uint32_t* h_out = reinterpret_cast<volatile uint32_t*>(h_pinned_) + 8;
  const int tail_scan_grid = blocks_for(dim, 1024, 4 * sm_count_);
  const int tail_update_grid = blocks_for(dim, 256, 4 * sm_count_);

  auto block_rows_of = [&](uint32_t b) { return std::min(B, dim - b * B); };
  auto slot_ptr = [&](uint32_t slot) { return d_mat_ + (size_t)slot * B * dim; };

  std::vector<int64_t> slot_block(C, -1);
  std::vector<char> slot_used(C, 0);
  uint64_t h2d_first = 0, h2d_last = 0, h2d_total = 0;
  uint32_t launches = 0, passes = 0, it = opt.max_iter;
  float lambda = 0.f;
  std::vector<float> round_us;

  ST_CUDA(cudaEventRecord(ev0_, stream_));
  initialise_eigen_vector(*this, E, dim); // reference :34 -> :280
  launches++;
  for (uint32_t k = 0; k < opt.max_iter; ++k) {
    const auto r0 = std::chrono::steady_clock::now();
    const bool backward = (k & 1u) != 0;
    uint64_t copied = 0;
    // cached blocks adjacent in the sweep form one run = one launch over rows that are contiguous
    // both in the matrix and in the cache
    bool run_open = false;
    uint32_t run_lo = 0, run_hi = 0;
    auto flush_run = [&] {
      if (!run_open)
        return;
      const uint32_t row0 = run_lo * B;
      const uint32_t rows = std::min(dim, (run_hi + 1) * B) - row0;
      launch_row_pass(*this, slot_ptr(run_lo % C), E, S, dim, row0, rows);
      launches++;
      for (uint32_t b = run_lo; b <= run_hi; b++)
        ST_CUDA(cudaEventRecord(slot_free_[b % C], stream_));
      run_open = false;
    };
    for (uint32_t j = 0; j < nb; j++) {
      const uint32_t b = backward ? nb - 1 - j : j;
      const uint32_t slot = b % C;
      if (slot_block[slot] == (int64_t)b) { // cached
        if (run_open) {
          const bool up = b == run_hi + 1 && b % C != 0;
          const bool down = b + 1 == run_lo && run_lo % C != 0;
          const bool room = (size_t)(run_hi - run_lo + 2) * block_bytes <= kStreamHitRunBytes;
          if ((up || down) && room) {
            run_lo = std::min(run_lo, b);
            run_hi = std::max(run_hi, b);
            continue;
          }
          flush_run();
        }
        run_open = true;
        run_lo = run_hi = b;
        continue;
      }
      flush_run();
      const uint32_t rows = block_rows_of(b);
      if (slot_used[slot]) // the pass that last read this slot must be done before it is overwritten
        ST_CUDA(cudaStreamWaitEvent(copy_stream_, slot_free_[slot], 0));
      copy_h2d(slot_ptr(slot), h_mat + (size_t)b * B * dim, row_bytes * rows, copy_stream_);
      ST_CUDA(cudaEventRecord(slot_ready_[slot], copy_stream_));
      ST_CUDA(cudaStreamWaitEvent(stream_, slot_ready_[slot], 0));
      launch_row_pass(*this, slot_ptr(slot), E, S, dim, b * B, rows);
      launches++;
      ST_CUDA(cudaEventRecord(slot_free_[slot], stream_));
      slot_used[slot] = 1;
      slot_block[slot] = b;
      copied += (uint64_t)row_bytes * rows;
    }
    flush_run();

    // ---- vector tail + the host's decision (reference :41-50) ----
    ST_CUDA(cudaMemsetAsync(cells, 0, sizeof(uint32_t), stream_));
    ST_LAUNCH_ELEMENTWISE(fill_u32_kernel, 1, 32, 0, stream_, cells + 1, stop_kind == kStopRelative ? 0u : 1u, 1u);
    ST_LAUNCH(tail_scan_kernel, tail_scan_grid, 1024, 0, stream_, (const float*)S, dim, opt.eps, stop_kind, cells);
    ST_LAUNCH_ELEMENTWISE(tail_update_kernel, tail_update_grid, 256, 0, stream_, (const float*)S, E, dim, opt.eps,
                          stop_kind, (const uint32_t*)cells, d_out);
    ST_CUDA(cudaGetLastError());
    launches += 3;
    ST_CUDA(cudaMemcpyAsync(const_cast<uint32_t*>(h_out), d_out, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    ST_CUDA(cudaStreamSynchronize(stream_));
    passes++;
    h2d_total += copied;
    if (k == 0)
      h2d_first = copied;
    else
      h2d_last = copied;
    const uint32_t bits = h_out[1];
    memcpy(&lambda, &bits, sizeof lambda); // :60-65
    round_us.push_back(std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - r0).count());
    if (h_out[0] != 0u) {
      it = k; // :54
      break;
    }
  }
  ST_CUDA(cudaEventRecord(ev1_, stream_));
  ST_CUDA(cudaEventSynchronize(ev1_));
  float loop_ms = 0.f;
  ST_CUDA(cudaEventElapsedTime(&loop_ms, ev0_, ev1_));
  if (h_eigen_vec)
    ST_CUDA(cudaMemcpy(h_eigen_vec, E, sizeof(float) * dim, cudaMemcpyDeviceToHost));
  if (h_eigen_val)
    *h_eigen_val = lambda;
  last_ts_.clear();
  last_phase_ts_.clear();

  if (plan) {
    plan->block_rows = B;
    plan->blocks = nb;
    plan->slots = C;
    plan->streamed = 1;
    plan->cache_bytes = (uint64_t)C * block_bytes;
    plan->h2d_bytes_first = h2d_first;
    plan->h2d_bytes_per_round = h2d_last;
    plan->h2d_bytes_total = h2d_total;
  }
  if (res) {
    memset(res, 0, sizeof *res);
    res->eigen_val = lambda;
    res->iter_count = it;
    res->passes = passes;
    res->launches = launches;
    res->loop_ms = loop_ms;
    res->bytes_per_round = mat_bytes;
    res->kernel_id = kStreamedKernelId;
    res->threads = 256;
    std::sort(round_us.begin(), round_us.end());
    res->round_us_min = round_us.front();
    res->round_us_median = round_us[round_us.size() / 2];
    res->status = ST_OK;
    res->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
  }
  return ST_OK;
}
