// stand-in for <cuda_runtime.h> when the kernel sources are compiled for the host (tests/cuda_emu)
#pragma once
#include "../cuda_emu.h"
