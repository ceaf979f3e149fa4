// sb_pyramid.cu -- Gaussian pyramid construction of the fed images (colours int16 x3, weights float32).
//
// Replaces the pyrDown chain of MultiBandBlender::feed (createLaplacePyr on the int16 image and the
// pyrDown loop over the weight map), reached from stitching/blender.py:41.  The padded copy of the image
// (copyMakeBorder) is never materialised: level 0 is read through index maps (sb_pyramid.cuh).
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
    D.g[o] = (int16_t)((acc[0] + 128) >> 8);
    D.g[D.plane + o] = (int16_t)((acc[1] + 128) >> 8);
    D.g[2 * D.plane + o] = (int16_t)((acc[2] + 128) >> 8);
    D.w[o] = tap5_v(rowf[0], rowf[1], rowf[2], rowf[3], rowf[4], v_simd);
}

#ifndef SB_EMU
// ---------------------------------------------------------------------------------------------------
// Fast variant: one warp walks down a strip of 32 destination columns.  Each lane owns the source pair
// (2x, 2x+1) of its column -- one coalesced load per source row -- and gets the other three taps of the
// horizontal [1 4 6 4 1] from its neighbours with warp shuffles; the five horizontally filtered rows live in
// registers as a sliding window (each destination row consumes two new source rows), so every source pixel is
// read from memory once (plus a 3-row warm-up per WK_ROWS rows and the strip-edge halo lanes).  No shared
// memory, no atomics; the float summation orders are per-lane constants (sb_pyramid.cuh).
// ---------------------------------------------------------------------------------------------------
constexpr int WK_ROWS = 32;   // destination rows per warp
constexpr int WK_WARPS = 4;   // warps per block (consecutive row chunks of the same strip)

struct HRow {
    int r, g, b;
    float w;
};

template <bool LEVEL0>
__global__ void __launch_bounds__(32 * WK_WARPS) k_pyrdown_walk(const FeedImage *__restrict__ imgs, int first, int l)
{
    const FeedImage &im = imgs[first + blockIdx.z];
    const int sw = im.pw >> l, sh = im.ph >> l, dw = sw >> 1, dh = sh >> 1;
    const int lane = threadIdx.x;
    const int x_raw = blockIdx.x * 32 + lane;
    const int y_begin = (blockIdx.y * WK_WARPS + threadIdx.y) * WK_ROWS;
    if ((int)blockIdx.x * 32 >= dw || y_begin >= dh) return;  // warp-uniform
    const int y_end = min(y_begin + WK_ROWS, dh);
    const bool valid = x_raw < dw;
    const int x = valid ? x_raw : dw - 1;
    const bool load_left = lane == 0;                       // taps 2x-2, 2x-1 are not in lane-1
    const bool load_right = lane == 31 || x_raw + 1 >= dw;  // tap 2x+2 is not in lane+1
    const bool h_simd = x >= 1 && x < pyrdown_hs_end(sw);
    const bool v_simd = x < (dw / 4) * 4;

    // source column of every tap (level coordinates), then -- at level 0 -- its column in the fed image
    int c[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) c[k] = reflect101(2 * x + k - 2, sw);
    bool cin[5] = {true, true, true, true, true};
    if (LEVEL0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int ix = c[k] - im.left;
            cin[k] = (unsigned)ix < (unsigned)im.w;
            c[k] = reflect(ix, im.w);
        }
    }
    const Level &S = im.lv[LEVEL0 ? 1 : l];  // source level descriptor (unused at level 0)

    auto hrow = [&](int src_row) -> HRow {
        const int sy = reflect101(src_row, sh);
        HRow o;
        if (LEVEL0) {
            const int iy = sy - im.top;
            const bool rin = (unsigned)iy < (unsigned)im.h;
            const uint32_t *row = im.rgbm + (long long)reflect(iy, im.h) * im.rgbm_pitch;
            // outside the fed image the weight is 0: clear the mask byte (0 * (1/255) == 0 exactly)
            unsigned p2 = __ldg(row + c[2]), p3 = __ldg(row + c[3]);
            if (!(rin && cin[2])) p2 &= 0x00ffffffu;
            if (!(rin && cin[3])) p3 &= 0x00ffffffu;
            unsigned p0 = __shfl_up_sync(0xffffffffu, p2, 1), p1 = __shfl_up_sync(0xffffffffu, p3, 1);
            unsigned p4 = __shfl_down_sync(0xffffffffu, p2, 1);
            if (load_left) {
                p0 = __ldg(row + c[0]);
                p1 = __ldg(row + c[1]);
                if (!(rin && cin[0])) p0 &= 0x00ffffffu;
                if (!(rin && cin[1])) p1 &= 0x00ffffffu;
            }
            if (load_right) {
                p4 = __ldg(row + c[4]);
                if (!(rin && cin[4])) p4 &= 0x00ffffffu;
            }
            o.r = (int)((p0 & 255u) + (p4 & 255u)) + 4 * (int)((p1 & 255u) + (p3 & 255u)) + 6 * (int)(p2 & 255u);
            o.g = (int)(((p0 >> 8) & 255u) + ((p4 >> 8) & 255u)) + 4 * (int)(((p1 >> 8) & 255u) + ((p3 >> 8) & 255u)) +
                  6 * (int)((p2 >> 8) & 255u);
            o.b = (int)(((p0 >> 16) & 255u) + ((p4 >> 16) & 255u)) + 4 * (int)(((p1 >> 16) & 255u) + ((p3 >> 16) & 255u)) +
                  6 * (int)((p2 >> 16) & 255u);
            o.w = tap5_h(fmul((float)(p0 >> 24), SB_INV255), fmul((float)(p1 >> 24), SB_INV255), fmul((float)(p2 >> 24), SB_INV255),
                         fmul((float)(p3 >> 24), SB_INV255), fmul((float)(p4 >> 24), SB_INV255), h_simd);
        } else {
            const long long ro = (long long)sy * S.pitch;
            int hs[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int16_t *row = S.g + ch * S.plane + ro;
                const unsigned pr = *reinterpret_cast<const unsigned *>(row + c[2]);  // (2x, 2x+1): 4-byte aligned
                const unsigned pl = __shfl_up_sync(0xffffffffu, pr, 1), pn = __shfl_down_sync(0xffffffffu, pr, 1);
                int v2 = (short)(pr & 0xffffu), v3 = (short)(pr >> 16);
                int v0 = (short)(pl & 0xffffu), v1 = (short)(pl >> 16), v4 = (short)(pn & 0xffffu);
                if (load_left) {
                    v0 = row[c[0]];
                    v1 = row[c[1]];
                }
                if (load_right) v4 = row[c[4]];
                hs[ch] = v0 + v4 + 4 * (v1 + v3) + 6 * v2;
            }
            o.r = hs[0];
            o.g = hs[1];
            o.b = hs[2];
            const float *wrow = S.w + ro;
            const float2 wp = *reinterpret_cast<const float2 *>(wrow + c[2]);
            float w0 = __shfl_up_sync(0xffffffffu, wp.x, 1), w1 = __shfl_up_sync(0xffffffffu, wp.y, 1);
            float w4 = __shfl_down_sync(0xffffffffu, wp.x, 1);
            if (load_left) {
                w0 = wrow[c[0]];
                w1 = wrow[c[1]];
            }
            if (load_right) w4 = wrow[c[4]];
            o.w = tap5_h(w0, w1, wp.x, wp.y, w4, h_simd);
        }
        return o;
    };

    HRow h0 = hrow(2 * y_begin - 2), h1 = hrow(2 * y_begin - 1), h2 = hrow(2 * y_begin);
    HRow h3 = hrow(2 * y_begin + 1), h4 = hrow(2 * y_begin + 2);
    const Level &D = im.lv[l + 1];
    for (int y = y_begin; y < y_end; ++y) {
        if (y > y_begin) {
            h0 = h2;
            h1 = h3;
            h2 = h4;
            h3 = hrow(2 * y + 1);
            h4 = hrow(2 * y + 2);
        }
        if (valid) {
            const long long o = (long long)y * D.pitch + x;
            D.g[o] = (int16_t)((h0.r + h4.r + 4 * (h1.r + h3.r) + 6 * h2.r + 128) >> 8);
            D.g[D.plane + o] = (int16_t)((h0.g + h4.g + 4 * (h1.g + h3.g) + 6 * h2.g + 128) >> 8);
            D.g[2 * D.plane + o] = (int16_t)((h0.b + h4.b + 4 * (h1.b + h3.b) + 6 * h2.b + 128) >> 8);
            D.w[o] = tap5_v(h0.w, h1.w, h2.w, h3.w, h4.w, v_simd);
        }
    }
}
#endif  // SB_EMU

}  // namespace

int launch_pyrdown(const FeedImage *imgs_dev, const FeedImage *imgs_host, int first, int count, int l, int max_w, int max_h,
                   cudaStream_t s)
{
    // max_w / max_h: largest DESTINATION level size among the images of the batch
    if (count <= 0 || max_w <= 0 || max_h <= 0) return SB_OK;
#ifndef SB_EMU
    // the walk kernel reads level 0 through the packed RGBM layout; generic int16 feeds use the gather kernel
    bool packed = true;
    for (int i = first; i < first + count; ++i) packed = packed && imgs_host[i].rgbm != nullptr;
    if (!use_simple_kernels() && (l > 0 || packed)) {
        dim3 block(32, WK_WARPS), grid(div_up(max_w, 32), div_up(div_up(max_h, WK_ROWS), WK_WARPS), count);
        if (l == 0)
            launch(k_pyrdown_walk<true>, grid, block, 0, s, imgs_dev, first, l);
        else
            launch(k_pyrdown_walk<false>, grid, block, 0, s, imgs_dev, first, l);
        return launch_check("k_pyrdown_walk");
    }
#else
    (void)imgs_host;
#endif
    dim3 block(PD_BX, PD_BY), grid(div_up(max_w, PD_BX), div_up(max_h, PD_BY), count);
    launch(k_pyrdown_gather, grid, block, 0, s, imgs_dev, first, l);
    return launch_check("k_pyrdown_gather");
}

}  // namespace sb
