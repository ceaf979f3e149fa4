// sb_pyramid.cu -- Gaussian pyramid construction of the fed images (colours int16 x3, weights float32).
//
// Replaces the pyrDown chain of MultiBandBlender::feed (createLaplacePyr on the int16 image and the
// pyrDown loop over the weight map), reached from stitching/blender.py:41.  The padded copy of the image
// (copyMakeBorder) is never materialised: level 0 is read through index maps (sb_pyramid.cuh).
#include <cstdlib>

#include "sb_launch.h"
#include "sb_pyramid.cuh"

namespace sb {

namespace {

constexpr int PD_BX = 32, PD_BY = 8;

// simple variant: one thread per destination pixel, 25 gathered taps
__global__ void __launch_bounds__(PD_BX *PD_BY) k_pyrdown_gather(const FeedImage *__restrict__ imgs, int first, int l)
{
    const FeedImage &im = imgs[first + blockIdx.z];
    const int sw = im.pw >> l, sh = im.ph >> l;  // source level size
    const int dw = sw >> 1, dh = sh >> 1;
    const int x = blockIdx.x * PD_BX + threadIdx.x;
    const int y = blockIdx.y * PD_BY + threadIdx.y;
    if (x >= dw || y >= dh) return;

    int xi[5], yi[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        xi[k] = reflect101(2 * x + k - 2, sw);
        yi[k] = reflect101(2 * y + k - 2, sh);
    }
    const bool h_simd = x >= 1 && x < pyrdown_hs_end(sw);
    const bool v_simd = x < (dw / 4) * 4;

    int acc[3] = {0, 0, 0};
    float rowf[5];
    const int kw[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        int hs[3] = {0, 0, 0};
        float wv[5];
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            int c[3];
            load_level(im, l, xi[kx], yi[ky], c, wv[kx]);
            hs[0] += kw[kx] * c[0];
            hs[1] += kw[kx] * c[1];
            hs[2] += kw[kx] * c[2];
        }
        acc[0] += kw[ky] * hs[0];
        acc[1] += kw[ky] * hs[1];
        acc[2] += kw[ky] * hs[2];
        rowf[ky] = tap5_h(wv[0], wv[1], wv[2], wv[3], wv[4], h_simd);
    }
    const Level &D = im.lv[l + 1];
    const long long o = (long long)y * D.pitch + x;
    const int out[3] = {(acc[0] + 128) >> 8, (acc[1] + 128) >> 8, (acc[2] + 128) >> 8};
    store_colours(D, o, out);
    D.w[o] = tap5_v(rowf[0], rowf[1], rowf[2], rowf[3], rowf[4], v_simd);
}


}  // namespace

int launch_pyrdown_fast(const PyrDesc *pyr, const FeedImage *imgs_host, int count, int l, int max_w, int max_h,
                        cudaStream_t s);  // sb_pyrdown_fast.cu

int launch_pyrdown(const FeedImage *imgs_dev, const FeedImage *imgs_host, const PyrDesc *pyr, int first, int count, int l,
                   int max_w, int max_h, cudaStream_t s)
{
    // max_w / max_h: largest DESTINATION level size among the images of the batch
    if (count <= 0 || max_w <= 0 || max_h <= 0) return SB_OK;
    // the fast kernel works on byte-fed images (RGBM level 0, lane-pair levels); generic int16 feeds use the gather kernel
    bool packed = true;
    for (int i = first; i < first + count; ++i) packed = packed && imgs_host[i].rgbm != nullptr;
#ifdef SB_EMU
    // the emulation plays a warp's 32 lanes with 32 host threads: correct but slow, so the shuffle kernel runs there only
    // on request (tests/test_host_logic.py sets SB_EMU_LANES around a small case); otherwise the gather kernel
    if (!getenv("SB_EMU_LANES")) packed = false;
#endif
    if (!use_simple_kernels() && packed) return launch_pyrdown_fast(pyr + first, imgs_host + first, count, l, max_w, max_h, s);
    dim3 block(PD_BX, PD_BY), grid(div_up(max_w, PD_BX), div_up(max_h, PD_BY), count);
    launch(k_pyrdown_gather, grid, block, 0, s, imgs_dev, first, l);
    return launch_check("k_pyrdown_gather");
}

}  // namespace sb
