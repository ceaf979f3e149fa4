// sb_collapse_fast.cu -- instruction-lean version of the per-level multiband kernel (levels below the top).
//
// Same arithmetic as k_collapse_gather (sb_blend.cu; see there for the reference call chain
// stitching/blender.py:41,46 -> MultiBandBlender::feed / ::blend), reorganised for the B200 SM, where the
// profile showed the kernel to be issue-bound rather than HBM-bound:
//   * one thread per 2x2 quad: the four pixels share the 3x3 neighbourhood of the coarser level, so each pyrUp
//     (of the image's G_{l+1} and of the collapsed C_{l+1}) is 9 loads per channel per quad, evaluated through
//     shared column sums;
//   * compact 16-byte-aligned descriptors (ColDesc) read with vector loads, 32-bit element offsets;
//   * exact early-outs: a fed image whose four weights in the quad are all zero contributes
//     (short)trunc(L * 0) = 0 and wsum + 0 = wsum, i.e. nothing -- at level 0 this is the whole padding ring
//     (constant-0 border of the weight map) and everything outside the warped footprint (mask byte 0);
//   * level 0 reads the packed RGBM layout; its colours are bytes, so the Laplacian cannot saturate there.
// A block covers a 64x16 tile and first marks which fed images touch it; images are visited in feed order.
#include "sb_launch.h"
#include "sb_pyramid.cuh"

namespace sb {

#ifndef SB_EMU
namespace {

constexpr int CF_BX = 32, CF_BY = 8;
#define SB_WEIGHT_EPS 1e-5f

struct Nbr {  // pyrUp source indices around coarse (ci, cj): element offsets of the three rows, the three columns
    int rp, rc, rn, xp, xc, xn;
};
__device__ __forceinline__ Nbr neighbours(int ci, int cj, int sw, int sh, int pitch)
{
    Nbr q;
    q.xc = ci;
    q.xp = ci > 0 ? ci - 1 : (sw > 1 ? 1 : 0);
    q.xn = ci + 1 < sw ? ci + 1 : sw - 1;
    const int yp = cj > 0 ? cj - 1 : (sh > 1 ? 1 : 0), yn = cj + 1 < sh ? cj + 1 : sh - 1;
    q.rp = yp * pitch;
    q.rc = cj * pitch;
    q.rn = yn * pitch;
    return q;
}
// the four pyrUp values of the quad (u[dy][dx]) from one int16 plane, via vertical column sums:
//   even row: e = a0 + 6 a1 + a2, odd row: o = a1 + a2 (x4 folded into the final shift)
//   (4 s + 32) >> 6 == (s + 8) >> 4 and (16 s + 32) >> 6 == (s + 2) >> 2 for integer s
__device__ __forceinline__ void up_quad(const int16_t *__restrict__ S, const Nbr &q, int u[2][2])
{
    int e[3], o[3];
    const int cols[3] = {q.xp, q.xc, q.xn};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a0 = S[q.rp + cols[k]], a1 = S[q.rc + cols[k]], a2 = S[q.rn + cols[k]];
        e[k] = a0 + a2 + 6 * a1;
        o[k] = a1 + a2;
    }
    u[0][0] = (e[0] + e[2] + 6 * e[1] + 32) >> 6;
    u[0][1] = (e[1] + e[2] + 8) >> 4;
    u[1][0] = (o[0] + o[2] + 6 * o[1] + 8) >> 4;
    u[1][1] = (o[1] + o[2] + 2) >> 2;
}

__device__ __forceinline__ int trunc16(float v) { return (int)(short)__float2int_rz(v); }  // |v| < 2^31 here

template <bool L0>
__global__ void __launch_bounds__(CF_BX *CF_BY)
    k_collapse_fast(const ColDesc *__restrict__ col, int n, PanoLevel up, PanoLevel cur, int lw, int lh, PanoOut out)
{
    __shared__ unsigned char cover[SB_MAX_IMAGES];
    const int tile_x = blockIdx.x * (2 * CF_BX), tile_y = blockIdx.y * (2 * CF_BY);
    for (int i = threadIdx.y * CF_BX + threadIdx.x; i < n; i += CF_BX * CF_BY) {
        const int4 r = __ldg(reinterpret_cast<const int4 *>(&col[i].ox));
        cover[i] = tile_x < r.x + r.z && tile_x + 2 * CF_BX > r.x && tile_y < r.y + r.w && tile_y + 2 * CF_BY > r.y;
    }
    __syncthreads();
    const int x = tile_x + 2 * threadIdx.x, y = tile_y + 2 * threadIdx.y;  // top-left pixel of the quad (even, even)
    if (x >= lw || y >= lh) return;
    if (L0 && (x >= out.w || y >= out.h)) return;

    int acc[2][2][3] = {};
    float wsum[2][2] = {};
    for (int i = 0; i < n; ++i) {
        if (!cover[i]) continue;
        const ColDesc &d = col[i];
        const int4 r = __ldg(reinterpret_cast<const int4 *>(&d.ox));
        const int X = x - r.x, Y = y - r.y;  // the rect origin and size are even: a quad is in or out as a whole
        if ((unsigned)X >= (unsigned)r.z || (unsigned)Y >= (unsigned)r.w) continue;
        const int4 s1 = __ldg(reinterpret_cast<const int4 *>(&d.top));  // top, pitch, plane, upitch
        int g[2][2][3];
        float wt[2][2];
        if (L0) {
            const int4 s0 = __ldg(reinterpret_cast<const int4 *>(&d.rgbm_pitch));  // rgbm_pitch, iw, ih, left
            const int ix = X - s0.w, iy = Y - s1.x;
            if (ix + 1 < 0 || ix >= s0.y || iy + 1 < 0 || iy >= s0.z) continue;  // whole quad in the zero-weight padding
            const uint32_t *base = d.rgbm;
            unsigned p[2][2];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const bool in = (unsigned)(ix + dx) < (unsigned)s0.y && (unsigned)(iy + dy) < (unsigned)s0.z;
                    p[dy][dx] = in ? __ldg(base + (iy + dy) * s0.x + ix + dx) : 0u;
                }
            if (((p[0][0] | p[0][1] | p[1][0] | p[1][1]) >> 24) == 0u) continue;  // all four weights are exactly 0
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    g[dy][dx][0] = p[dy][dx] & 255u;
                    g[dy][dx][1] = (p[dy][dx] >> 8) & 255u;
                    g[dy][dx][2] = (p[dy][dx] >> 16) & 255u;
                    wt[dy][dx] = fmul((float)(p[dy][dx] >> 24), SB_INV255);
                }
        } else {
            const int o0 = Y * s1.y + X;  // X even: the pairs are 4- / 8-byte aligned
            const float2 w0 = __ldg(reinterpret_cast<const float2 *>(d.w + o0));
            const float2 w1 = __ldg(reinterpret_cast<const float2 *>(d.w + o0 + s1.y));
            if (w0.x == 0.f && w0.y == 0.f && w1.x == 0.f && w1.y == 0.f) continue;  // contributes exactly nothing
            wt[0][0] = w0.x; wt[0][1] = w0.y; wt[1][0] = w1.x; wt[1][1] = w1.y;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned a = __ldg(reinterpret_cast<const unsigned *>(d.g + c * s1.z + o0));
                const unsigned b = __ldg(reinterpret_cast<const unsigned *>(d.g + c * s1.z + o0 + s1.y));
                g[0][0][c] = (short)(a & 0xffffu); g[0][1][c] = (short)(a >> 16);
                g[1][0][c] = (short)(b & 0xffffu); g[1][1][c] = (short)(b >> 16);
            }
        }
        const int uplane = __ldg(&d.uplane);
        const Nbr q = neighbours(X >> 1, Y >> 1, r.z >> 1, r.w >> 1, s1.w);
        const int16_t *ug = d.ug;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int u[2][2];
            up_quad(ug + c * uplane, q, u);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    int lap = g[dy][dx][c] - u[dy][dx];
                    if (!L0) lap = sat_s16(lap);  // bytes minus an average of bytes cannot leave int16
                    acc[dy][dx][c] += trunc16(fmul((float)lap, wt[dy][dx]));
                }
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) wsum[dy][dx] = fadd(wsum[dy][dx], wt[dy][dx]);
    }

    // blend step + collapse: v = sat16(pyrUp(C_{l+1}) + (short)trunc(acc / (wsum + eps)))
    int v[2][2][3];
    float den[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) den[dy][dx] = fadd(wsum[dy][dx], SB_WEIGHT_EPS);
    const Nbr q = neighbours(x >> 1, y >> 1, up.w_px, up.h_px, up.pitch);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int u[2][2];
        up_quad(up.c + c * up.plane, q, u);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx)
                v[dy][dx][c] = sat_s16(u[dy][dx] + f2s_wrap(fdiv((float)(short)acc[dy][dx][c], den[dy][dx])));
    }
    if (!L0) {
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int o = (y + dy) * cur.pitch + x;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<unsigned *>(cur.c + c * cur.plane + o) = ((unsigned)v[dy][0][c] & 0xffffu) | ((unsigned)v[dy][1][c] << 16);
        }
        return;
    }
    // level 0: mask, zero outside it, crop to the roi, |v| saturated to uint8 (convertScaleAbs)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        if (y + dy >= out.h) continue;
        const bool on0 = wsum[dy][0] > SB_WEIGHT_EPS, on1 = wsum[dy][1] > SB_WEIGHT_EPS;
        const bool two = x + 1 < out.w;
        if (out.rgb) {
            unsigned b[6];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                b[c] = on0 ? (unsigned)min(abs(v[dy][0][c]), 255) : 0u;
                b[3 + c] = on1 ? (unsigned)min(abs(v[dy][1][c]), 255) : 0u;
            }
            uint8_t *p = out.rgb + (long long)(y + dy) * out.rgb_pitch + 3 * x;  // x even: 2-byte aligned when the pitch is even
            if (two && ((out.rgb_pitch & 1) == 0)) {
                unsigned short *p2 = reinterpret_cast<unsigned short *>(p);
                p2[0] = (unsigned short)(b[0] | (b[1] << 8));
                p2[1] = (unsigned short)(b[2] | (b[3] << 8));
                p2[2] = (unsigned short)(b[4] | (b[5] << 8));
            } else {
                p[0] = (uint8_t)b[0]; p[1] = (uint8_t)b[1]; p[2] = (uint8_t)b[2];
                if (two) { p[3] = (uint8_t)b[3]; p[4] = (uint8_t)b[4]; p[5] = (uint8_t)b[5]; }
            }
        }
        if (out.mask) {
            uint8_t *m = out.mask + (long long)(y + dy) * out.mask_pitch + x;
            if (two && ((out.mask_pitch & 1) == 0)) {
                *reinterpret_cast<unsigned short *>(m) = (unsigned short)((on0 ? 255u : 0u) | (on1 ? 0xff00u : 0u));
            } else {
                m[0] = on0 ? 255 : 0;
                if (two) m[1] = on1 ? 255 : 0;
            }
        }
        if (out.s16) {
            int16_t *d = out.s16 + (long long)(y + dy) * out.s16_pitch + 3 * x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c] = (int16_t)(on0 ? v[dy][0][c] : 0);
                if (two) d[3 + c] = (int16_t)(on1 ? v[dy][1][c] : 0);
            }
        }
    }
}

}  // namespace

int launch_collapse_fast(const ColDesc *col, int n, const PanoLevel &up, const PanoLevel &cur, int l, int lw, int lh, PanoOut out,
                         cudaStream_t s)
{
    const int gw = l == 0 ? out.w : lw, gh = l == 0 ? out.h : lh;
    dim3 block(CF_BX, CF_BY), grid(div_up(gw, 2 * CF_BX), div_up(gh, 2 * CF_BY));
    if (l == 0)
        launch(k_collapse_fast<true>, grid, block, 0, s, col, n, up, cur, lw, lh, out);
    else
        launch(k_collapse_fast<false>, grid, block, 0, s, col, n, up, cur, lw, lh, out);
    return launch_check("k_collapse_fast");
}
#else
int launch_collapse_fast(const ColDesc *, int, const PanoLevel &, const PanoLevel &, int, int, int, PanoOut, cudaStream_t)
{
    return SB_ERR_INVALID;  // never called in the emulation build
}
#endif

}  // namespace sb
