// sb_launch.h -- one place where kernels are launched, so launches can be counted (bench.py
// "gpu_launches") and so tests/emu can substitute a serial CPU shim for the <<<>>> syntax when it
// compiles these sources on a GPU-less box (test infrastructure; the product build never defines SB_EMU).
#pragma once
#include <utility>

#include "sb_internal.h"

namespace sb {

#ifdef SB_EMU
template <typename... KArgs, typename... Args>
inline void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args &&...args)
{
    count_launch();
    sb_emu_run(grid, block, [&]() { kern(args...); });
}
// kernels whose lanes exchange values (warp shuffles): the emulation runs the 32 lanes of a warp concurrently
template <typename... KArgs, typename... Args>
inline void launch_lanes(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args &&...args)
{
    count_launch();
    sb_emu_run_lanes(grid, block, [&]() { kern(args...); });
}
#else
template <typename... KArgs, typename... Args>
inline void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    count_launch();
    kern<<<grid, block, smem, s>>>(std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_lanes(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    launch(kern, grid, block, smem, s, std::forward<Args>(args)...);
}
#endif

inline int launch_check(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what, __FILE__, __LINE__);
    return SB_OK;
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace sb
