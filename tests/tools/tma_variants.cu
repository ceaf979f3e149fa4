#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int V>
__global__ void k_tma(const __grid_constant__ CUtensorMap tm, int x, int y, int *out)
{
    __shared__ alignas(1024) int buf[16][64];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (V == 6) {
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tm), "r"(x), "r"(y) : "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar)) : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(4096) : "memory");
            if (V == 2)
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                             "l"(&tm), "r"(x), "r"(y), "r"(smem_addr(&bar)) : "memory");
            else if (V == 3)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_addr(buf)),
                             "l"(&tm), "r"(x), "r"(y), "r"(smem_addr(&bar)), "l"(0x1000000000000000ull) : "memory");
            else if (V == 4)
                asm volatile("cp.async.bulk.tensor.2d.cta_group::1.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                             "l"(&tm), "r"(x), "r"(y), "r"(smem_addr(&bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                             "l"(&tm), "r"(x), "r"(y), "r"(smem_addr(&bar)) : "memory");
        }
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i / 64][i % 64];
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    cudaFree(0);
    const int W = 1024, H = 721;
    int *d, *o;
    cudaMalloc(&d, W * H * 4);
    cudaMalloc(&o, 4096);
    int *h = (int *)malloc(W * H * 4);
    for (int i = 0; i < W * H; ++i) h[i] = i;
    cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    cuuint64_t dims[2] = {(cuuint64_t)(v == 7 ? 1024 : 933), H};
    cuuint64_t strides[1] = {W * 4};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t es[2] = {1, 1};
    CUtensorMap m;
    CUresult r = ((EncodeFn)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                v == 5 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (v == 5) { box[0] = 32; r = ((EncodeFn)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    printf("variant %d encode %d\n", v, (int)r);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(128);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = v == 1 ? 1 : 0;
    cudaError_t e;
    switch (v) {
    case 2: e = cudaLaunchKernelEx(&cfg, k_tma<2>, m, 3, 2, o); break;
    case 3: e = cudaLaunchKernelEx(&cfg, k_tma<3>, m, 3, 2, o); break;
    case 4: e = cudaLaunchKernelEx(&cfg, k_tma<4>, m, 3, 2, o); break;
    case 6: e = cudaLaunchKernelEx(&cfg, k_tma<6>, m, 3, 2, o); break;
    default: e = cudaLaunchKernelEx(&cfg, k_tma<0>, m, 3, 2, o); break;
    }
    printf("launch: %s; ", cudaGetErrorString(e));
    printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    int got[4];
    cudaMemcpy(got, o, 16, cudaMemcpyDeviceToHost);
    printf("got %d %d %d %d expect %d\n", got[0], got[1], got[2], got[3], 2 * W + 3);
    return 0;
}
