// tsan_main.cpp -- runs the emulated round kernels under ThreadSanitizer (TEST INFRASTRUCTURE).
// Built by tests/cuda_emu/build.py::build_tsan(); see the TSan section of cuda_emu.h for what a clean
// run does and does not prove.  `emu_tsan racy` runs a deliberately racy kernel as a positive control.
#include "cuda_emu.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
struct emu_opts
{
  float eps;
  uint32_t max_iter;
  int32_t form, sweep, dynamic, threads, ctas, kernel, stop, bf16, world, acc64;
  const float* row_scale;
};
int emu_solve(const void* mat, uint32_t dim, const emu_opts* o, float* eigen_val, float* eigen_vec,
              uint32_t* iter_count, uint32_t* passes, uint32_t* ranks_agree);
const char* emu_last_error();
}

static std::vector<float>
uniform_matrix(uint32_t dim, uint32_t seed)
{
  std::vector<float> m((size_t)dim * dim);
  uint32_t x = seed * 2654435761u + 12345u;
  for (float& v : m) {
    x = x * 1664525u + 1013904223u;
    v = 0.25f + (float)(x >> 8) * (1.0f / 16777216.0f);
  }
  return m;
}

static std::vector<unsigned short>
to_bf16(const std::vector<float>& m)
{
  std::vector<unsigned short> out(m.size());
  for (size_t i = 0; i < m.size(); i++) {
    uint32_t u;
    memcpy(&u, &m[i], 4);
    out[i] = (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
  }
  return out;
}

struct Case
{
  const char* name;
  uint32_t dim;
  emu_opts o;
};

// A kernel with a missing __syncthreads: thread t writes slot t, then reads its neighbour's slot.
struct RacyParams
{
  int* out;
};
static void
racy_kernel(const RacyParams p)
{
  __shared__ int slot[64];
  slot[threadIdx.x] = (int)threadIdx.x;
  // __syncthreads() is missing here
  p.out[threadIdx.x] = slot[(threadIdx.x + 1u) % blockDim.x];
}

int
main(int argc, char** argv)
{
  if (argc > 1 && std::string(argv[1]) == "racy") {
    std::vector<int> out(64, 0);
    RacyParams p{ out.data() };
    auto g = emu::launch_async<RacyParams>(racy_kernel, 1, 64, 0, p);
    emu::join(*g);
    printf("racy kernel done\n");
    return 0;
  }
  const emu_opts base{ 1e-3f, 1000u, 0, 1, -1, 64, 3, 1, 0, 0, 1, 0 };
  auto with = [&](int kernel, int threads, int ctas, int world = 1, int form = 0, int stop = 0, int dynamic = -1,
                  int bf16 = 0, uint32_t max_iter = 1000u) {
    emu_opts o = base;
    o.kernel = kernel, o.threads = threads, o.ctas = ctas, o.world = world, o.form = form, o.stop = stop;
    o.dynamic = dynamic, o.bf16 = bf16, o.max_iter = max_iter;
    if (stop)
      o.eps = 1e-6f;
    return o;
  };
  const std::vector<Case> cases = {
    { "general read-only 200, 3 CTAs", 200, with(1, 64, 3) },
    { "general in-place 200, 3 CTAs", 200, with(1, 64, 3, 1, 1) },
    { "general scalar 201, relative stop", 201, with(1, 64, 3, 1, 0, 1) },
    { "resident-e 13 (prefetch), static units, 320", 320, with(13, 64, 3, 1, 0, 0, 0) },
    { "resident-e 13 (prefetch), dynamic units, 320", 320, with(13, 64, 3, 1, 0, 0, 1) },
    { "resident-e 10, resident rows, 64", 64, with(10, 64, 8) },
    { "resident-e 11 bf16, relative stop, 320", 320, with(11, 64, 3, 1, 0, 1, 1, 1) },
    { "resident-e 13, two units per row, 8200 x 2 rounds", 8200, with(13, 64, 4, 1, 0, 0, 1, 0, 2u) },
    { "cluster kernel 128 (DSMEM exchange)", 128, with(20, 512, 8) },
    { "2 GPUs, general 200", 200, with(1, 64, 2, 2) },
    { "3 GPUs, resident-e 13 dynamic 320", 320, with(13, 64, 2, 3, 0, 0, 1) },
    { "resident-e 13, fp64 accumulation, 320", 320, [&] { emu_opts o = with(13, 64, 3); o.acc64 = 1; return o; }() },
    { "resident-e 11 on scalar units, dynamic, 321", 321, with(11, 64, 3, 1, 0, 0, 1) },
    { "2 GPUs, resident-e 11 on scalar units, two units per row, 8195 x 2 rounds", 8195, with(11, 64, 3, 2, 0, 0, 1, 0, 2u) },
    { "wide kernel 2, dynamic units, 320", 320, with(2, 64, 3, 1, 0, 0, 1) },
    { "2 GPUs, wide kernel 2, static units, 200", 200, with(2, 64, 2, 2, 0, 0, 0) },
    { "wide kernel 2, three 8192-column windows (ST_EMU_WINDOW), 16400 x 2 rounds", 16400, [&] { setenv("ST_EMU_WINDOW", "8192", 1); return with(2, 64, 4, 1, 0, 0, 1, 0, 2u); }() },
  };
  int failures = 0;
  for (const Case& c : cases) {
    std::vector<float> m = uniform_matrix(c.dim, c.dim);
    std::vector<unsigned short> m16;
    const void* data = m.data();
    if (c.o.bf16) {
      m16 = to_bf16(m);
      data = m16.data();
    }
    float val = 0.f;
    std::vector<float> vec(c.dim);
    uint32_t it = 0, passes = 0, agree = 1;
    const int rc = emu_solve(data, c.dim, &c.o, &val, vec.data(), &it, &passes, &agree);
    printf("%-68s rc=%d rounds=%u lambda=%.6f agree=%u%s%s\n", c.name, rc, it, val, agree, rc ? " error: " : "",
           rc ? emu_last_error() : "");
    fflush(stdout);
    failures += (rc != 0 || !agree);
  }
  return failures ? 1 : 0;
}
