// sb_tma.cuh -- the Blackwell/Hopper bulk-tensor copy engine (TMA) for the tile kernels: mbarrier + cp.async.bulk.tensor
// wrappers (inline PTX; SASS: SYNCS.*, UTMALDG) and the host side that encodes tensor maps through the driver entry
// point cuTensorMapEncodeTiled (the library links the CUDA runtime statically and does not link libcuda).
//
// A tile kernel stages rectangular windows of the pyramid levels in shared memory with ONE instruction per window,
// issued by one elected thread: the copy engine does the address arithmetic, the bounds tests (out-of-range elements
// arrive as zeros, which is exactly "weight 0 / not covered" for every buffer staged this way) and the coalescing;
// the threads then read their taps with shared-memory loads at compile-time offsets.
#pragma once
#include <stdint.h>

#ifndef SB_EMU
#include <cuda.h>
#include <cuda_runtime.h>

namespace sb {

typedef CUtensorMap TensorMap;

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// makes the barrier initialisation visible to the async proxy (the copy engine); followed by a __syncthreads
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders earlier generic-proxy accesses of shared memory before later async-proxy writes (buffer reuse)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SB_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SB_MBAR_DONE;\n\t"
        "bra SB_MBAR_WAIT;\n\t"
        "SB_MBAR_DONE:\n\t"
        "}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// window of a 2-D / 3-D tensor -> shared memory; (x, y[, z]) = element coordinates of the window's first element,
// innermost first, may be negative / reach beyond the tensor (zero fill); completion counts bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void *dst, const TensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const TensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_addr(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar))
                 : "memory");
}
// A tensor map that lives in GLOBAL memory and was written by the host (cudaMemcpy): the thread that is going to use it
// makes the tensormap proxy acquire it first (CUDA programming guide, "tensor maps in global memory")
__device__ __forceinline__ void tma_acquire_desc(const TensorMap *map)
{
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
}
#endif  // __CUDACC__

}  // namespace sb
#else   // SB_EMU: the emulation build has no copy engine; the tile kernels are not compiled there
namespace sb {
struct alignas(64) TensorMap {
    unsigned long long opaque[16];
};
}  // namespace sb
#endif

namespace sb {
enum TmaType { TMA_U8 = 0, TMA_U16 = 1, TMA_U32 = 2, TMA_U64 = 3, TMA_F32 = 4 };
// Encodes a tiled tensor map for `base` = [planes][h][pitch_elems] elements (planes == 0: two-dimensional) of which
// [w x h] are valid, window = box_w x box_h (x box_planes).  Returns SB_OK, or an error when the geometry violates the
// copy engine's rules (16-byte aligned base and row pitch, box rows a multiple of 16 bytes, box sides <= 256).
int tensor_map_encode(TensorMap *out, int type, const void *base, long long w, long long h, long long pitch_elems, long long planes,
                      long long plane_elems, int box_w, int box_h, int box_planes);
bool tensor_maps_available();  // false in the emulation build
}  // namespace sb
