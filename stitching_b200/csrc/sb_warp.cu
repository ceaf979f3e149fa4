// sb_warp.cu -- fused backward-map warp: image (fixed-point bilinear, BORDER_REFLECT) and validity mask
// (nearest, BORDER_CONSTANT) in one pass, no float maps in memory.
//
// Replaces, per output pixel, what cv.PyRotationWarper.warp does in three passes
// (stitching/warper.py:43-52 image, :58-68 mask): buildMaps -> remap(INTER_LINEAR, BORDER_REFLECT) and
// buildMaps -> remap(INTER_NEAREST, BORDER_CONSTANT).
//
// Arithmetic contract (bit-exact with the reference's CPU path):
//   projection  x_ = rowA[v]*colX[u], y_ = rowY[v], z_ = rowA[v]*colZ[u]      (host libm tables)
//               (x,y,z) = k_rinv * (x_,y_,z_)  plain fp32, (a+b)+c, every op rounded, no FMA
//               spherical/cylindrical: z > 0 ? (x/z, y/z) : (-1,-1);  plane: always divide
//   bilinear    sx = cvRound(x*32) (half-even, INT_MIN when unrepresentable), ix = sat16(sx>>5),
//               fx = sx&31; 4 taps with 15-bit weights (32-fy)(32-fx)*32 ..., (sum + 2^14) >> 15
//   mask        255 iff 0 <= sat16(cvRound(x)) < W and 0 <= sat16(cvRound(y)) < H
//
// The job descriptors travel BY VALUE in the kernel parameter block (constant bank, up to SB_WARP_BATCH images
// per launch): the profile of the first version showed 20 of 31 loads per pixel re-reading them from global memory.
#include "sb_device.cuh"
#include "sb_launch.h"

namespace sb {

namespace {

constexpr int WARP_BX = 32, WARP_BY = 8;

struct WarpBatch {
    WarpJob j[SB_WARP_BATCH];
};

__device__ __forceinline__ void project(const WarpJob &j, int u, int v, float &x, float &y)
{
    const float cx = __ldg(j.colX + u), cz = __ldg(j.colZ + u);
    const float ra = __ldg(j.rowA + v), ry = __ldg(j.rowY + v);
    const float x_ = fmul(ra, cx), y_ = ry, z_ = fmul(ra, cz);
    x = fadd(fadd(fmul(j.k[0], x_), fmul(j.k[1], y_)), fmul(j.k[2], z_));
    y = fadd(fadd(fmul(j.k[3], x_), fmul(j.k[4], y_)), fmul(j.k[5], z_));
    const float z = fadd(fadd(fmul(j.k[6], x_), fmul(j.k[7], y_)), fmul(j.k[8], z_));
    if (j.always_divide || z > 0.f) {
        x = fdiv(x, z);
        y = fdiv(y, z);
    } else {
        x = -1.f;
        y = -1.f;
    }
}

__device__ __forceinline__ void store_pixel(const WarpJob &j, int u, int v, unsigned r, unsigned g, unsigned b, unsigned m)
{
    if (j.dst_rgb) {
        uint8_t *d = j.dst_rgb + (long long)v * j.dst_pitch + 3 * u;
        d[0] = (uint8_t)r;
        d[1] = (uint8_t)g;
        d[2] = (uint8_t)b;
    }
    if (j.dst_rgbm) {
        if (j.blend_mask) m = j.blend_mask[(long long)v * j.blend_mask_pitch + u];
        j.dst_rgbm[(unsigned)v * (unsigned)j.rgbm_pitch + (unsigned)u] = r | (g << 8) | (b << 16) | (m << 24);
    }
}

// simple variant: one thread per output pixel, byte gathers
__global__ void __launch_bounds__(WARP_BX *WARP_BY) k_warp_gather(const __grid_constant__ WarpBatch B)
{
    const WarpJob &j = B.j[blockIdx.z];
    const int u = blockIdx.x * WARP_BX + threadIdx.x;
    const int v = blockIdx.y * WARP_BY + threadIdx.y;
    if (u >= j.dw || v >= j.dh) return;

    float x, y;
    project(j, u, v, x, y);

    // validity mask: nearest neighbour into an all-255 source, constant-0 border
    const int nx = sat_s16(cvt_rn_x86(x)), ny = sat_s16(cvt_rn_x86(y));
    const unsigned m = ((unsigned)nx < (unsigned)j.sw && (unsigned)ny < (unsigned)j.sh) ? 255u : 0u;
    if (j.dst_mask) j.dst_mask[(long long)v * j.mask_pitch + u] = (uint8_t)m;
    if (!j.dst_rgb && !j.dst_rgbm) return;

    const int sx = cvt_rn_x86(fmul(x, 32.f)), sy = cvt_rn_x86(fmul(y, 32.f));
    const int ix = sat_s16(sx >> 5), iy = sat_s16(sy >> 5);
    const int fx = sx & 31, fy = sy & 31;
    const int x0 = reflect(ix, j.sw), x1 = reflect(ix + 1, j.sw);
    const int y0 = reflect(iy, j.sh), y1 = reflect(iy + 1, j.sh);
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32;
    const int w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    const uint8_t *r0 = j.src + (long long)y0 * j.spitch, *r1 = j.src + (long long)y1 * j.spitch;
    unsigned out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int acc = __ldg(r0 + x0 * 3 + c) * w00 + __ldg(r0 + x1 * 3 + c) * w01 + __ldg(r1 + x0 * 3 + c) * w10 +
                  __ldg(r1 + x1 * 3 + c) * w11;
        out[c] = (unsigned)sat_u8((acc + (1 << 14)) >> 15);
    }
    store_pixel(j, u, v, out[0], out[1], out[2], m);
}

#ifndef SB_EMU
// Fast variant: the two horizontally adjacent source pixels of a bilinear footprint are 6 contiguous bytes;
// fetch them with one (or two) aligned 8-byte loads instead of six byte loads.  Callers allocate the source
// with SB_SRC_PAD spare bytes so that the second aligned word may straddle the end of the image.
__device__ __forceinline__ void fetch_pair(const uint8_t *__restrict__ src, unsigned row_off, int x0, int x1, unsigned &p0, unsigned &p1)
{
    if (x1 == x0 + 1) {
        const unsigned long long addr = (unsigned long long)src + row_off + 3u * (unsigned)x0;
        const unsigned o = (unsigned)addr & 7u;
        const uint2 *q = reinterpret_cast<const uint2 *>(addr - o);
        const uint2 lo = __ldg(q);
        // bytes o .. o+5 of the 16-byte window (lo, hi): shift the 32-bit words with funnel shifts
        const unsigned sh = 8u * (o & 3u);
        unsigned w0 = lo.x, w1 = lo.y, w2 = 0u;
        if (o > 2) {
            const uint2 hi = __ldg(q + 1);
            w2 = hi.x;
            if (o >= 4) {
                w0 = lo.y;
                w1 = hi.x;
                w2 = hi.y;
            }
        }
        const unsigned a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh);  // bytes 0-3 and 4-7 from the start
        p0 = a & 0xffffffu;
        p1 = __funnelshift_r(a, b, 24) & 0xffffffu;
    } else {  // the footprint straddles a reflected border
        const uint8_t *a = src + row_off + 3 * x0, *b = src + row_off + 3 * x1;
        p0 = (unsigned)__ldg(a) | ((unsigned)__ldg(a + 1) << 8) | ((unsigned)__ldg(a + 2) << 16);
        p1 = (unsigned)__ldg(b) | ((unsigned)__ldg(b + 1) << 8) | ((unsigned)__ldg(b + 2) << 16);
    }
}

// RGBM_ONLY: the compositor's case -- packed output only and source sides <= 32767, where int16 saturation can
// neither move a coordinate across the inside test nor touch a footprint that lies inside the image.
template <bool RGBM_ONLY>
__global__ void __launch_bounds__(WARP_BX *WARP_BY) k_warp_wide(const __grid_constant__ WarpBatch B)
{
    const WarpJob &j = B.j[blockIdx.z];
    const int u = blockIdx.x * WARP_BX + threadIdx.x;
    const int v = blockIdx.y * WARP_BY + threadIdx.y;
    if (u >= j.dw || v >= j.dh) return;
    const int sw = j.sw, sh = j.sh;

    float x, y;
    project(j, u, v, x, y);
    unsigned m;
    if (RGBM_ONLY) {
        m = ((unsigned)cvt_rn_x86(x) < (unsigned)sw && (unsigned)cvt_rn_x86(y) < (unsigned)sh) ? 255u : 0u;
    } else {
        const int nx = sat_s16(cvt_rn_x86(x)), ny = sat_s16(cvt_rn_x86(y));
        m = ((unsigned)nx < (unsigned)sw && (unsigned)ny < (unsigned)sh) ? 255u : 0u;
        if (j.dst_mask) j.dst_mask[(long long)v * j.mask_pitch + u] = (uint8_t)m;
        if (!j.dst_rgb && !j.dst_rgbm) return;
    }

    const int sx = cvt_rn_x86(fmul(x, 32.f)), sy = cvt_rn_x86(fmul(y, 32.f));
    const int fx = sx & 31, fy = sy & 31;
    int ix = sx >> 5, iy = sy >> 5;
    if (!RGBM_ONLY) {
        ix = sat_s16(ix);
        iy = sat_s16(iy);
    }
    unsigned a0, a1, b0, b1;
    const unsigned pitch = (unsigned)j.spitch;  // coordinates are int16-saturated: offsets stay below 2^32
    if ((unsigned)ix < (unsigned)(sw - 1) && (unsigned)iy < (unsigned)(sh - 1)) {
        // the 2x2 footprint lies inside the image: no border rule applies, the pairs are adjacent
        const unsigned off = (unsigned)iy * pitch;
        fetch_pair(j.src, off, ix, ix + 1, a0, a1);
        fetch_pair(j.src, off + pitch, ix, ix + 1, b0, b1);
    } else {
        ix = sat_s16(ix);
        iy = sat_s16(iy);
        const int x0 = reflect(ix, sw), x1 = reflect(ix + 1, sw);
        const int y0 = reflect(iy, sh), y1 = reflect(iy + 1, sh);
        fetch_pair(j.src, (unsigned)y0 * pitch, x0, x1, a0, a1);
        fetch_pair(j.src, (unsigned)y1 * pitch, x0, x1, b0, b1);
    }
    // (sum_k w_k p_k + 2^14) >> 15 with w = 32 (32-fy|fy)(32-fx|fx), evaluated as two exact lerps:
    //   h = (32-fx) a + fx b  (<= 8160),  s = (32-fy) h0 + fy h1  (<= 261120),  out = (s + 512) >> 10
    // red and blue share a register as two 16-bit lanes through the horizontal lerp
    const unsigned M = 0x00ff00ffu;
    const unsigned gx = (unsigned)fx, hx = 32u - gx;
    const unsigned h0rb = (a0 & M) * hx + (a1 & M) * gx, h1rb = (b0 & M) * hx + (b1 & M) * gx;
    const unsigned h0g = ((a0 >> 8) & 255u) * hx + ((a1 >> 8) & 255u) * gx, h1g = ((b0 >> 8) & 255u) * hx + ((b1 >> 8) & 255u) * gx;
    const unsigned gy = (unsigned)fy, hy = 32u - gy;
    const unsigned r = ((h0rb & 0xffffu) * hy + (h1rb & 0xffffu) * gy + 512u) >> 10;
    const unsigned b = ((h0rb >> 16) * hy + (h1rb >> 16) * gy + 512u) >> 10;
    const unsigned g = (h0g * hy + h1g * gy + 512u) >> 10;
    if (RGBM_ONLY) {
        if (j.blend_mask) m = j.blend_mask[(unsigned)v * (unsigned)j.blend_mask_pitch + (unsigned)u];
        j.dst_rgbm[(unsigned)v * (unsigned)j.rgbm_pitch + (unsigned)u] = r | (g << 8) | (b << 16) | (m << 24);
    } else {
        store_pixel(j, u, v, r, g, b, m);
    }
}
#endif  // SB_EMU

__global__ void k_pack_rgbm(const uint8_t *__restrict__ rgb, long long rgb_pitch, const uint8_t *__restrict__ mask,
                            long long mask_pitch, uint32_t *__restrict__ dst, long long dst_pitch, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = rgb + (long long)y * rgb_pitch + (long long)x * 3;
    const unsigned m = mask[(long long)y * mask_pitch + x];
    dst[(long long)y * dst_pitch + x] = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | (m << 24);
}

}  // namespace

int launch_warp(const WarpJob *jobs_host, int n_jobs, cudaStream_t s)
{
    for (int first = 0; first < n_jobs; first += SB_WARP_BATCH) {
        const int cnt = n_jobs - first < SB_WARP_BATCH ? n_jobs - first : SB_WARP_BATCH;
        WarpBatch B;
        int max_w = 0, max_h = 0;
        for (int i = 0; i < cnt; ++i) {
            B.j[i] = jobs_host[first + i];
            max_w = max_w > B.j[i].dw ? max_w : B.j[i].dw;
            max_h = max_h > B.j[i].dh ? max_h : B.j[i].dh;
        }
        for (int i = cnt; i < SB_WARP_BATCH; ++i) B.j[i] = B.j[0];
        if (max_w <= 0 || max_h <= 0) continue;
        dim3 block(WARP_BX, WARP_BY), grid(div_up(max_w, WARP_BX), div_up(max_h, WARP_BY), cnt);
#ifndef SB_EMU
        if (!use_simple_kernels()) {
            bool rgbm_only = true;
            for (int i = 0; i < cnt; ++i)
                rgbm_only = rgbm_only && B.j[i].dst_rgbm && !B.j[i].dst_rgb && !B.j[i].dst_mask && B.j[i].sw <= 32767 && B.j[i].sh <= 32767;
            if (rgbm_only)
                launch(k_warp_wide<true>, grid, block, 0, s, B);
            else
                launch(k_warp_wide<false>, grid, block, 0, s, B);
            SB_TRY(launch_check("k_warp_wide"));
            continue;
        }
#endif
        launch(k_warp_gather, grid, block, 0, s, B);
        SB_TRY(launch_check("k_warp_gather"));
    }
    return SB_OK;
}

int launch_pack_rgbm(const uint8_t *rgb, long long rgb_pitch, const uint8_t *mask, long long mask_pitch, uint32_t *dst,
                     long long dst_pitch, int w, int h, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(w, 32), div_up(h, 8));
    launch(k_pack_rgbm, grid, block, 0, s, rgb, rgb_pitch, mask, mask_pitch, dst, dst_pitch, w, h);
    return launch_check("k_pack_rgbm");
}

}  // namespace sb
