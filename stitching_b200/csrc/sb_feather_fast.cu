// sb_feather_fast.cu -- feather weights: w = min(L1 distance to the nearest zero mask pixel * sharpness, 1).
//
// Replaces createWeightMap of cv.detail_FeatherBlender (stitching/blender.py:34-36, :41):
// distanceTransform(mask, DIST_L1, 3) + multiply + threshold(TRUNC 1).  The exact city-block distance is separable:
// d(x,y) = min_y' ( r(x,y') + |y - y'| ) with r = distance along the row to the nearest zero of that row.
//   rows:    one warp per row; per 32-pixel chunk a ballot gives the zero positions, clz / ffs the nearest zero to the
//            left / right inside the chunk, a carried index the nearest one in earlier chunks.  Two coalesced sweeps.
//   columns: one thread per column (coalesced across the warp), a downward and an upward min-plus sweep; the upward
//            sweep writes the weight.  All images of the blend in one launch each.
// "No zero anywhere" stays at DT_INF and becomes weight 1, as with OpenCV (the image border is not a zero).
#include "sb_launch.h"
#include "sb_pyramid.cuh"

namespace sb {

namespace {

#define DT_INF (1 << 29)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned mask_at(const FeedImage &im, int x, int y)
{
    if (im.rgbm) return __ldg(im.rgbm + (long long)y * im.rgbm_pitch + x) >> 24;
    return im.mask[(long long)y * im.mask_pitch + x];
}

__global__ void __launch_bounds__(256) k_dt_rows_warp(const FeedImage *__restrict__ imgs)
{
    grid_dependency_sync();
    const FeedImage &im = imgs[blockIdx.y];
    const int y = blockIdx.x * 8 + threadIdx.y, lane = threadIdx.x, w = im.w;
    if (y >= im.h) return;  // warp-uniform
    int *d = (int *)im.fw + (long long)y * w;
    int last = -(1 << 30);  // column of the nearest zero seen so far
    for (int b = 0; b < w; b += 32) {
        const int x = b + lane;
        const bool inb = x < w;
        const unsigned zeros = __ballot_sync(FULL, inb && mask_at(im, inb ? x : 0, y) == 0u);
        const unsigned le = zeros & (FULL >> (31 - lane));
        const int lz = le ? b + 31 - __clz(le) : last;
        if (inb) d[x] = min(x - lz, DT_INF);
        if (zeros) last = b + 31 - __clz(zeros);
    }
    int next = 1 << 30;
    for (int b = ((w - 1) / 32) * 32; b >= 0; b -= 32) {
        const int x = b + lane;
        const bool inb = x < w;
        const int dl = inb ? d[x] : DT_INF;
        const unsigned zeros = __ballot_sync(FULL, inb && dl == 0);
        const unsigned ge = zeros & (FULL << lane);
        const int nz = ge ? b + __ffs(ge) - 1 : next;
        if (inb) d[x] = min(dl, min(nz - x, DT_INF));
        if (zeros) next = b + __ffs(zeros) - 1;
    }
}

__global__ void __launch_bounds__(128) k_dt_cols_batched(const FeedImage *__restrict__ imgs, float sharpness)
{
    const FeedImage &im = imgs[blockIdx.y];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, w = im.w, h = im.h;
    if (x >= w) return;
    int *d = (int *)im.fw + x;
    int run = DT_INF;
#pragma unroll 8
    for (int y = 0; y < h; ++y) {
        run = min(d[(long long)y * w], min(run + 1, DT_INF));
        d[(long long)y * w] = run;
    }
    run = DT_INF;
    float *f = (float *)im.fw + x;
#pragma unroll 8
    for (int y = h - 1; y >= 0; --y) {
        run = min(d[(long long)y * w], min(run + 1, DT_INF));
        const float dist = run >= DT_INF ? 3.402823466e+38f : (float)run;
        f[(long long)y * w] = fminf(fmul(dist, sharpness), 1.f);
    }
}

}  // namespace

int launch_feather_weights_fast(const FeedImage *imgs_dev, const FeedImage *imgs_host, int n, float sharpness, cudaStream_t s)
{
    int mw = 0, mh = 0;
    for (int i = 0; i < n; ++i) {
        mw = mw > imgs_host[i].w ? mw : imgs_host[i].w;
        mh = mh > imgs_host[i].h ? mh : imgs_host[i].h;
    }
    if (n <= 0 || mw <= 0 || mh <= 0) return SB_OK;
    launch_lanes(k_dt_rows_warp, dim3(div_up(mh, 8), n), dim3(32, 8), 0, s, imgs_dev);
    launch(k_dt_cols_batched, dim3(div_up(mw, 128), n), dim3(128), 0, s, imgs_dev, sharpness);
    return launch_check("k_dt_*");
}

}  // namespace sb
