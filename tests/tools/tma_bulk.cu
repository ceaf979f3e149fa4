#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_bulk(const int *src, int *out)
{
    __shared__ alignas(128) int buf[1024];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(buf)), "l"(src), "r"(4096),
                     "r"(smem_addr(&bar))
                     : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}
__global__ void k_tma(const __grid_constant__ CUtensorMap tm, int x, int y, int *out)
{
    __shared__ alignas(128) int buf[16][64];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                     "l"(&tm), "r"(x), "r"(y), "r"(smem_addr(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i / 64][i % 64];
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void dump(const char *n, const CUtensorMap &m) { printf("%s:", n); const unsigned long long *p = (const unsigned long long *)&m; for (int i = 0; i < 16; ++i) printf(" %016llx", p[i]); printf("\n"); }
int main(int argc, char **argv)
{
    cudaFree(0);
    const int W = 960, H = 721;
    int *d, *o;
    cudaMalloc(&d, W * H * 4);
    cudaMalloc(&o, 4096);
    int *h = (int *)malloc(W * H * 4);
    for (int i = 0; i < W * H; ++i) h[i] = i;
    cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice);
    int got[4];
    if (argc > 1 && !strcmp(argv[1], "bulk")) {
        k_bulk<<<1, 128>>>(d, o);
        printf("bulk kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        cudaMemcpy(got, o, 16, cudaMemcpyDeviceToHost);
        printf("got %d %d %d %d\n", got[0], got[1], got[2], got[3]);
        return 0;
    }
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    void *lib = dlopen("libcuda.so.1", RTLD_NOW);
    void *fp2 = lib ? dlsym(lib, "cuTensorMapEncodeTiled") : nullptr;
    printf("entry points: runtime %p dlsym %p\n", fp, fp2);
    cuuint64_t dims[2] = {933, H};
    cuuint64_t strides[1] = {W * 4};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t es[2] = {1, 1};
    CUtensorMap m1, m2;
    memset(&m1, 0, sizeof m1); memset(&m2, 0, sizeof m2);
    CUresult r1 = ((EncodeFn)fp)(&m1, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = fp2 ? ((EncodeFn)fp2)(&m2, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) : (CUresult)-1;
    printf("encode %d %d, device ptr %p\n", (int)r1, (int)r2, (void *)d);
    dump("m1", m1); dump("m2", m2);
    k_tma<<<1, 128>>>(fp2 && argc > 1 && !strcmp(argv[1], "dlsym") ? m2 : m1, 3, 2, o);
    printf("tma kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(got, o, 16, cudaMemcpyDeviceToHost);
    printf("got %d %d %d %d expect %d\n", got[0], got[1], got[2], got[3], 2 * W + 3);
    return 0;
}
