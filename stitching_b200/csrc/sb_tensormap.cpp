// sb_tensormap.cpp -- host side of the TMA staging (sb_tma.cuh): encodes CUtensorMap objects through the driver entry
// point, which the statically linked CUDA runtime hands out (the library does not link libcuda itself).
#include "sb_tma.cuh"

#include "sb_internal.h"

namespace sb {

#ifndef SB_EMU
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
}  // namespace

bool tensor_maps_available() { return encode_fn() != nullptr; }

int tensor_map_encode(TensorMap *out, int type, const void *base, long long w, long long h, long long pitch_elems, long long planes,
                      long long plane_elems, int box_w, int box_h, int box_planes)
{
    static const CUtensorMapDataType types[] = {CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_DATA_TYPE_UINT16, CU_TENSOR_MAP_DATA_TYPE_UINT32,
                                                CU_TENSOR_MAP_DATA_TYPE_UINT64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32};
    static const int sizes[] = {1, 2, 4, 8, 4};
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("tensor maps: cuTensorMapEncodeTiled is not available from this driver");
        return SB_ERR_CUDA;
    }
    const int es = sizes[type];
    const bool ok = ((uintptr_t)base & 15) == 0 && (pitch_elems * es) % 16 == 0 && (box_w * es) % 16 == 0 && box_w >= 1 && box_w <= 256 &&
                    box_h >= 1 && box_h <= 256 && w >= 1 && h >= 1 && (planes == 0 || ((plane_elems * es) % 16 == 0 && box_planes >= 1));
    if (!ok) {
        set_error("tensor maps: geometry outside the copy engine's rules (base %p, %lldx%lld, pitch %lld x %d B, box %dx%d)", base, w, h,
                  pitch_elems, es, box_w, box_h);
        return SB_ERR_INVALID;
    }
    const cuuint32_t rank = planes ? 3 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(planes ? planes : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)(pitch_elems * es), (cuuint64_t)(plane_elems * es)};  // bytes, dimensions 1.. (dimension 0 is dense)
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)(planes ? box_planes : 1)};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, types[type], rank, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (base %p, %lldx%lld, pitch %lld x %d B, box %dx%d)", (int)r, base, w, h, pitch_elems,
                  es, box_w, box_h);
        return SB_ERR_CUDA;
    }
    return SB_OK;
}
#else
bool tensor_maps_available() { return false; }
int tensor_map_encode(TensorMap *, int, const void *, long long, long long, long long, long long, long long, int, int, int)
{
    set_error("tensor maps are not available in the emulation build");
    return SB_ERR_STATE;
}
#endif

}  // namespace sb
