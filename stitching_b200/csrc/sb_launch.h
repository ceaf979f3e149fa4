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
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    launch(kern, grid, block, smem, s, std::forward<Args>(args)...);
}
// kernels whose lanes exchange values (warp shuffles): the emulation runs the 32 lanes of a warp concurrently
template <typename... KArgs, typename... Args>
inline void launch_lanes(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args &&...args)
{
    count_launch();
    sb_emu_run_lanes(grid, block, [&]() { kern(args...); });
}
// kernels whose threads share memory and meet at block barriers: one host thread per thread of a block
template <typename... KArgs, typename... Args>
inline void launch_block(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args &&...args)
{
    count_launch();
    sb_emu_run_block(grid, block, smem, [&]() { kern(args...); });
}
#else
template <typename... KArgs, typename... Args>
inline void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    count_launch();
    kern<<<grid, block, smem, s>>>(std::forward<Args>(args)...);
}
// Programmatic dependent launch for the kernels of a step (each of them starts with grid_dependency_sync()): the
// next kernel's CTAs may become resident while the last wave of the previous kernel is still running and its launch
// latency overlaps the previous kernel; they then wait (griddepcontrol.wait) until the previous grid has completed
// and its stores are visible.  Stream capture turns these launches into programmatic edges of the step's graph.
// SB_PDL=0 launches them like any other kernel.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    if (!pdl_enabled()) {
        launch(kern, grid, block, smem, s, std::forward<Args>(args)...);
        return;
    }
    count_launch();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
template <typename... KArgs, typename... Args>
inline void launch_lanes(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    launch_pdl(kern, grid, block, smem, s, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_block(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    launch_pdl(kern, grid, block, smem, s, std::forward<Args>(args)...);
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
