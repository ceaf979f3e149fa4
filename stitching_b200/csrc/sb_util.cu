// sb_util.cu -- small utility kernels (L2 flush for benchmarking hygiene).
#include "sb_launch.h"

namespace sb {
namespace {
__global__ void k_flush(uint4 *p, size_t n, unsigned v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_uint4(v, v, v, v);
}
}  // namespace

// overwrite a buffer larger than L2 so that the next kernel starts from a cold cache
int launch_flush_l2(void *buf, size_t bytes, cudaStream_t s)
{
    static unsigned v = 0;
    launch(k_flush, dim3(148 * 8), dim3(256), 0, s, (uint4 *)buf, bytes / sizeof(uint4), ++v);
    return launch_check("k_flush");
}
}  // namespace sb
