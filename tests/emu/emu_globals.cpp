// TEST INFRASTRUCTURE ONLY -- storage for the emulated builtin index variables.
#include "cuda_runtime.h"
thread_local emuIdx threadIdx, blockIdx, blockDim, gridDim;
thread_local EmuWarpSync *emu_warp = nullptr;
thread_local EmuBlockSync *emu_block = nullptr;
thread_local unsigned char *emu_smem = nullptr;
