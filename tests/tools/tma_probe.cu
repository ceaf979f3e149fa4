// tma_probe.cu -- standalone check of the tensor-map shapes the tile kernels use (TEST TOOL, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu -cudart static -ldl -lpthread
// Loads one window per shape at a few origins (inside, negative, beyond the tensor) and compares with a host gather.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const CUtensorMap *map, int x, int y, unsigned bytes, unsigned char *out, int fence)
{
    extern __shared__ __align__(128) unsigned char buf[];  // [window | barrier]: no static shared memory in front of it
    uint64_t &bar = *reinterpret_cast<uint64_t *>(buf + ((bytes + 127u) & ~127u));
    if (threadIdx.x == 0) {
        // a tensor map in GLOBAL memory written by the host: the tensormap proxy has to acquire it first
        if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                         "l"(map), "r"(x), "r"(y), "r"(smem_addr(&bar))
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_addr(buf)),
                         "l"(map), "r"(x), "r"(y), "r"(0), "r"(smem_addr(&bar))
                         : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_addr(&bar)),
        "r"(0)
        : "memory");
    for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

template <int RANK>
__global__ void k_probe_param(const __grid_constant__ CUtensorMap pmap, int x, int y, unsigned bytes, unsigned char *out)
{
    extern __shared__ __align__(128) unsigned char buf[];
    uint64_t &bar = *reinterpret_cast<uint64_t *>(buf + ((bytes + 127u) & ~127u));
    const CUtensorMap *map = &pmap;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(buf)),
                         "l"(map), "r"(x), "r"(y), "r"(smem_addr(&bar))
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_addr(buf)),
                         "l"(map), "r"(x), "r"(y), "r"(0), "r"(smem_addr(&bar))
                         : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_addr(&bar)),
        "r"(0)
        : "memory");
    for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int fence = 1;
int probe(EncodeFn enc, const char *name, CUtensorMapDataType type, int es, int w, int h, int pitch, int planes, int bw, int bh)
{
    const size_t plane = (size_t)h * pitch;
    const size_t total = plane * (planes ? planes : 1) * es;
    std::vector<unsigned char> host(total);
    for (size_t i = 0; i < total; ++i) host[i] = (unsigned char)(i * 2654435761u >> 13);
    unsigned char *dev = nullptr, *out = nullptr;
    CUtensorMap *dmap = nullptr;
    CK(cudaMalloc(&dev, total));
    CK(cudaMemcpy(dev, host.data(), total, cudaMemcpyHostToDevice));
    const int bp = planes ? planes : 1;
    const unsigned bytes = (unsigned)bw * bh * bp * es;
    CK(cudaMalloc(&out, bytes));
    CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)bp};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * es, (cuuint64_t)plane * es};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, type, planes ? 3 : 2, dev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("%-28s encode FAILED %d\n", name, (int)r);
        return 0;
    }
    CK(cudaMemcpy(dmap, &m, sizeof m, cudaMemcpyHostToDevice));
    const int origins[4][2] = {{3, 2}, {-1, -1}, {w - bw / 2, h - bh / 2}, {-bw + 1, 5}};
    int bad_total = 0;
    for (int o = 0; o < 4; ++o) {
        const int x = origins[o][0], y = origins[o][1];
        CK(cudaMemset(out, 0xcd, bytes));
        const unsigned sm = ((bytes + 127u) & ~127u) + 16;
        if (fence == 2) {
            if (planes)
                k_probe_param<3><<<1, 128, sm>>>(m, x, y, bytes, out);
            else
                k_probe_param<2><<<1, 128, sm>>>(m, x, y, bytes, out);
        } else if (planes)
            k_probe<3><<<1, 128, sm>>>(dmap, x, y, bytes, out, fence);
        else
            k_probe<2><<<1, 128, sm>>>(dmap, x, y, bytes, out, fence);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-28s origin (%d,%d): KERNEL FAILED: %s\n", name, x, y, cudaGetErrorString(e));
            return 2;
        }
        std::vector<unsigned char> got(bytes);
        CK(cudaMemcpy(got.data(), out, bytes, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int p = 0; p < bp; ++p)
            for (int j = 0; j < bh; ++j)
                for (int i = 0; i < bw; ++i)
                    for (int b = 0; b < es; ++b) {
                        const int sx = x + i, sy = y + j;
                        unsigned char exp = 0;
                        if (sx >= 0 && sx < w && sy >= 0 && sy < h) exp = host[((size_t)p * plane + (size_t)sy * pitch + sx) * es + b];
                        if (got[(((size_t)p * bh + j) * bw + i) * es + b] != exp) ++bad;
                    }
        bad_total += bad;
        printf("%-28s origin (%4d,%4d): %d bytes differ\n", name, x, y, bad);
    }
    cudaFree(dev);
    cudaFree(out);
    cudaFree(dmap);
    return bad_total ? 3 : 0;
}

int main(int argc, char **argv)
{
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    if (!fp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeFn enc = (EncodeFn)fp;
    int rc = 0;
    fence = argc > 1 ? (!strcmp(argv[1], "nofence") ? 0 : !strcmp(argv[1], "param") ? 2 : 1) : 1;
    printf("mode %d (0 = global map, no fence; 1 = global map + tensormap fence; 2 = __grid_constant__ parameter)\n", fence);
    rc |= probe(enc, "rgbm u32 64x16", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 933, 721, 960, 0, 64, 16);
    rc |= probe(enc, "lanes u64 64x16", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 580, 470, 640, 0, 64, 16);
    rc |= probe(enc, "weights f32 64x16", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 580, 470, 640, 0, 64, 16);
    rc |= probe(enc, "up u64 34x10", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 290, 235, 320, 0, 34, 10);
    rc |= probe(enc, "up u64 36x10", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 290, 235, 320, 0, 36, 10);
    rc |= probe(enc, "up u64 32x10", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 290, 235, 320, 0, 32, 10);
    rc |= probe(enc, "c1 u16 40x10x3", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1152, 192, 1152, 3, 40, 10);
    rc |= probe(enc, "c1 u16 64x10x3", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 1152, 192, 1152, 3, 64, 10);
    printf("probe rc %d\n", rc);
    return rc;
}
