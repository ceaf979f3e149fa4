// sb_gather.cuh -- the per-pixel forms of the two multiband steps (one destination pixel from plain gathers, either
// level layout): the first correct version of the path.  Used by the simple kernels (SB_KERNELS=simple, generic int16
// feeds, 0-band blends; sb_pyramid.cu, sb_blend.cu) and by the fused pyramid-tail kernel (sb_tail.cu), where the
// levels are so small that one launch for all of them beats a tuned kernel per level.
#pragma once
#include "sb_pyramid.cuh"

namespace sb {

#ifndef SB_WEIGHT_EPS
#define SB_WEIGHT_EPS 1e-5f
#endif

// pyrDown of level l -> l+1 of one fed image at destination pixel (x, y): colours int16 (s + 128) >> 8, weights float32
// in the reference build's summation orders (sb_pyramid.cuh)
__device__ __forceinline__ void pyrdown_pixel(const FeedImage &im, int l, int x, int y)
{
    const int sw = im.pw >> l, sh = im.ph >> l, dw = sw >> 1;
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

__device__ __forceinline__ void store_final(const PanoOut &out, int x, int y, const int v[3], bool on, unsigned mask_value)
{
    if (out.s16) {
        int16_t *d = out.s16 + (long long)y * out.s16_pitch + (long long)x * 3;
        d[0] = (int16_t)(on ? v[0] : 0);
        d[1] = (int16_t)(on ? v[1] : 0);
        d[2] = (int16_t)(on ? v[2] : 0);
    }
    if (out.rgb) {
        uint8_t *d = out.rgb + (long long)y * out.rgb_pitch + (long long)x * 3;
        // convertScaleAbs: min(|v|, 255)
        d[0] = (uint8_t)(on ? min(abs(v[0]), 255) : 0);
        d[1] = (uint8_t)(on ? min(abs(v[1]), 255) : 0);
        d[2] = (uint8_t)(on ? min(abs(v[2]), 255) : 0);
    }
    if (out.mask) out.mask[(long long)y * out.mask_pitch + x] = (uint8_t)mask_value;
}

// accumulate + normalise + collapse of pano pixel (x, y) of level l: loops over all images with a rect test
__device__ __forceinline__ void collapse_pixel(const FeedImage *__restrict__ imgs, int n, const PanoLevel *__restrict__ pano, int l, int nb,
                                               int x, int y, const PanoOut &out)
{
    int acc[3] = {0, 0, 0};
    float wsum = 0.f;
    for (int i = 0; i < n; ++i) {
        const FeedImage &im = imgs[i];
        const int X = x - (im.px >> l), Y = y - (im.py >> l);
        const int w_l = im.pw >> l, h_l = im.ph >> l;
        if ((unsigned)X >= (unsigned)w_l || (unsigned)Y >= (unsigned)h_l) continue;
        int g[3];
        float wt;
        load_level(im, l, X, Y, g, wt);
        if (l < nb) {
            int up[3];
            pyrup_level_at(im.lv[l + 1], w_l >> 1, h_l >> 1, X, Y, up);
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = sat_s16(g[c] - up[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += f2s_wrap(fmul((float)g[c], wt));
        wsum = fadd(wsum, wt);
    }
    const float den = fadd(wsum, SB_WEIGHT_EPS);
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = f2s_wrap(fdiv((float)(short)acc[c], den));
    if (l < nb) {
        const PanoLevel &P = pano[l + 1];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = sat_s16(pyrup_at(P.c + c * P.plane, P.pitch, P.w_px, P.h_px, x, y) + v[c]);
    }
    if (l > 0) {
        const PanoLevel &P = pano[l];
        const long long o = (long long)y * P.pitch + x;
        P.c[o] = (int16_t)v[0];
        P.c[P.plane + o] = (int16_t)v[1];
        P.c[2 * P.plane + o] = (int16_t)v[2];
    } else {
        const bool on = wsum > SB_WEIGHT_EPS;
        store_final(out, x, y, v, on, on ? 255u : 0u);
    }
}

}  // namespace sb
