// stream_probe.cu -- how fast can one B200 stream a large read-only buffer into its SMs?
// Standalone probe (not part of the library):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_probe tools/stream_probe.cu
//   ./stream_probe [GiB]
// Variants: register-staged LDG.128 (what round_loop_kernel does), per-warp TMA bulk rings
// (what round_loop_tma_kernel does) with several tile sizes / depths, with and without the
// LDS consumption of the landed tile, and a CTA-wide ring fed by one producer thread.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e = (x);                                                                           \
    if (e != cudaSuccess) {                                                                        \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                                               \
      exit(1);                                                                                     \
    }                                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ float4 ld_stream(const float4* p)
{
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---- (a) register-staged loads: warp w of the grid streams segments of SEG bytes ---------------
template<int UNROLL>
__global__ void __launch_bounds__(512, 1) ldg_kernel(const float4* __restrict__ a, size_t n4, size_t seg4, float* out)
{
  const int lane = threadIdx.x & 31;
  const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (size_t s0 = gw * seg4; s0 < n4; s0 += nw * seg4) {
    for (size_t i = lane; i < seg4; i += 32 * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        v[u] = ld_stream(a + s0 + i + 32 * u);
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f)
    *out = acc;
}

// ---- (b) per-warp TMA rings -------------------------------------------------------------------
template<int TILE_BYTES, int STAGES, bool CONSUME>
__global__ void __launch_bounds__(512, 1) tma_warp_ring(const char* __restrict__ a, size_t bytes, float* out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp * STAGES * TILE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)W * STAGES * TILE_BYTES) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; s++)
      mbar_init(bar + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const size_t ntiles = bytes / TILE_BYTES;
  const size_t gw = (size_t)blockIdx.x * W + warp, nw = (size_t)gridDim.x * W;
  // warp streams rows of 8 tiles (32 KB @ 4 KB tiles) like the solver: tile index = row*8 + t
  const size_t my = (ntiles / 8 > gw) ? ((ntiles / 8 - gw + nw - 1) / nw) * 8 : 0;
  auto tile_addr = [&](size_t n) { return a + (((n / 8) * nw + gw) * 8 + (n % 8)) * (size_t)TILE_BYTES; };
  size_t issued = 0;
  if (lane == 0)
    for (; issued < STAGES && issued < my; issued++) {
      mbar_expect(bar + issued % STAGES, TILE_BYTES);
      bulk_load(ring + (issued % STAGES) * TILE_BYTES, tile_addr(issued), TILE_BYTES, bar + issued % STAGES);
    }
  float acc = 0.f;
  for (size_t n = 0; n < my; n++) {
    const int s = n % STAGES;
    mbar_wait(bar + s, (n / STAGES) & 1);
    const float4* t4 = reinterpret_cast<const float4*>(ring + s * TILE_BYTES);
    if (CONSUME) {
#pragma unroll
      for (int u = 0; u < TILE_BYTES / 512; u++) {
        float4 v = t4[lane + 32 * u];
        acc += v.x + v.y + v.z + v.w;
      }
    } else {
      acc += t4[lane].x;
    }
    __syncwarp();
    if (lane == 0 && issued < my) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(bar + s, TILE_BYTES);
      bulk_load(ring + s * TILE_BYTES, tile_addr(issued), TILE_BYTES, bar + s);
      issued++;
    }
  }
  if (acc == 123.456f)
    *out = acc;
}

// ---- (c) CTA-wide ring, one producer thread, big tiles -------------------------------------------
template<int TILE_BYTES, int STAGES>
__global__ void __launch_bounds__(544, 1) tma_cta_ring(const char* __restrict__ a, size_t bytes, float* out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * TILE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CW = (blockDim.x >> 5) - 1; // consumer warps
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, CW);
    }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const size_t ntiles = bytes / TILE_BYTES;
  const size_t my = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp == CW) { // producer warp
    if (lane == 0)
      for (size_t n = 0; n < my; n++) {
        const int s = n % STAGES;
        if (n >= STAGES)
          mbar_wait(empty + s, ((n / STAGES) - 1) & 1);
        mbar_expect(full + s, TILE_BYTES);
        bulk_load(smem + (size_t)s * TILE_BYTES, a + (blockIdx.x + n * gridDim.x) * (size_t)TILE_BYTES, TILE_BYTES,
                  full + s);
      }
  } else {
    float acc = 0.f;
    for (size_t n = 0; n < my; n++) {
      const int s = n % STAGES;
      mbar_wait(full + s, (n / STAGES) & 1);
      const float4* t4 = reinterpret_cast<const float4*>(smem + (size_t)s * TILE_BYTES);
      for (int j = warp * 32 + lane; j < TILE_BYTES / 16; j += CW * 32) {
        float4 v = t4[j];
        acc += v.x + v.y + v.z + v.w;
      }
      __syncwarp();
      if (lane == 0)
        mbar_arrive(empty + s);
    }
    if (acc == 123.456f)
      *out = acc;
  }
}

template<typename F>
static void
run(const char* name, size_t bytes, F&& launch)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best)
      best = ms;
  }
  printf("%-44s %8.1f GB/s  (%.3f ms)\n", name, bytes / (best * 1e-3) / 1e9, best);
  fflush(stdout);
}

int
main(int argc, char** argv)
{
  const double gib = argc > 1 ? atof(argv[1]) : 4.0;
  const size_t bytes = (size_t)(gib * (1ull << 30)) & ~(size_t)((1 << 20) - 1);
  char* a;
  float* out;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(a, 0, bytes));
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("# buffer %.2f GiB, %d SMs\n", bytes / double(1 << 30), sms);
  const size_t n4 = bytes / 16;

  run("ldg.128 x8  512thr seg=32KB", bytes, [&] { ldg_kernel<8><<<sms, 512>>>((const float4*)a, n4, 2048, out); });
  run("ldg.128 x8  256thr seg=32KB", bytes, [&] { ldg_kernel<8><<<sms, 256>>>((const float4*)a, n4, 2048, out); });
  run("ldg.128 x8  512thr x2 CTAs/SM", bytes, [&] { ldg_kernel<8><<<2 * sms, 512>>>((const float4*)a, n4, 2048, out); });
  run("ldg.128 x4  512thr seg=32KB", bytes, [&] { ldg_kernel<4><<<sms, 512>>>((const float4*)a, n4, 2048, out); });
  run("ldg.128 x16 512thr seg=32KB", bytes, [&] { ldg_kernel<16><<<sms, 512>>>((const float4*)a, n4, 2048, out); });
  run("ldg.128 x8  512thr seg=128KB", bytes, [&] { ldg_kernel<8><<<sms, 512>>>((const float4*)a, n4, 8192, out); });

#define WARP_RING(T, S, C, THREADS)                                                                \
  {                                                                                                \
    auto k = tma_warp_ring<T, S, C>;                                                               \
    const size_t sm = (size_t)(THREADS / 32) * S * T + (THREADS / 32) * S * 8;                      \
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));             \
    char nm[96];                                                                                   \
    snprintf(nm, sizeof nm, "tma warp-ring tile=%dB x%d %dthr %s", T, S, THREADS, C ? "consume" : "touch");    \
    run(nm, bytes, [&] { k<<<sms, THREADS, sm>>>(a, bytes, out); });                               \
  }
  WARP_RING(4096, 3, true, 512)
  WARP_RING(4096, 3, false, 512)
  WARP_RING(4096, 2, true, 512)
  WARP_RING(4096, 2, false, 512)
  WARP_RING(4096, 1, true, 512)
  WARP_RING(2048, 4, true, 512)
  WARP_RING(2048, 6, true, 512)
  WARP_RING(8192, 1, true, 512)
  WARP_RING(8192, 3, true, 256)
  WARP_RING(4096, 6, true, 256)
  WARP_RING(16384, 3, true, 128)

#define CTA_RING(T, S)                                                                             \
  {                                                                                                \
    auto k = tma_cta_ring<T, S>;                                                                   \
    const size_t sm = (size_t)S * T + 2 * S * 8;                                                   \
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));             \
    char nm[96];                                                                                   \
    snprintf(nm, sizeof nm, "tma cta-ring  tile=%dB x%d 16 consumer warps", T, S);                 \
    run(nm, bytes, [&] { k<<<sms, 544, sm>>>(a, bytes, out); });                                   \
  }
  CTA_RING(32768, 6)
  CTA_RING(32768, 4)
  CTA_RING(16384, 12)
  CTA_RING(16384, 6)
  CTA_RING(8192, 24)
  CTA_RING(65536, 3)
  return 0;
}
