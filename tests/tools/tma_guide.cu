#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SMEM_W = 64, SMEM_H = 16;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SMEM_H][SMEM_W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SMEM_H * SMEM_W; i += blockDim.x) out[i] = smem_buffer[i / SMEM_W][i % SMEM_W];
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    cudaFree(0);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    const int W = 960, H = 721;
    int *d, *o;
    cudaMalloc(&d, W * H * 4);
    cudaMalloc(&o, SMEM_W * SMEM_H * 4);
    int *h = (int *)malloc(W * H * 4);
    for (int i = 0; i < W * H; ++i) h[i] = i;
    cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {933, H};
    cuuint64_t strides[1] = {W * 4};
    cuuint32_t box[2] = {SMEM_W, SMEM_H};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    kernel<<<1, 128>>>(m, 3, 2, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    int got[4];
    cudaMemcpy(got, o, 16, cudaMemcpyDeviceToHost);
    printf("got %d %d %d %d expect %d..\n", got[0], got[1], got[2], got[3], 2 * W + 3);
    return 0;
}
