// emu_rt.cpp -- the one runtime entry point that needs the kernels' parameter type (TEST INFRASTRUCTURE)
#include "cuda_runtime_emu.h"

#include "kernels.cuh"

cudaError_t
emu_launch_round_kernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem)
{
  auto kernel = reinterpret_cast<void (*)(const st::RoundParams)>(const_cast<void*>(func));
  auto g = emu::launch_async<st::RoundParams>(kernel, grid.x, block.x, smem, *static_cast<const st::RoundParams*>(args[0]));
  emu::join(*g);
  return cudaSuccess;
}
