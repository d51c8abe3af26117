// stand-in for <cuda_runtime.h> when the WHOLE library (solver.cu, abi.cu) is built for the host
#pragma once
#include "../cuda_runtime_emu.h"
