// sb_shard.cpp -- geometry and launch sequencing of the multi-GPU composite (no reference counterpart: the
// reference is single-process; this is SURVEY.md 8(e)).
//
// Images are dealt to ranks in contiguous blocks (feed order is rank order); the padded panorama is cut into
// column strips, one per rank, at multiples of 2^nb.  A rank computes warp + pyramids for its own images only.
// Where the padded footprint of its images reaches into another rank's strip it hands that rank a SLAB per level:
// its partial sums (acc int16x3 with wrap-around, wsum float32) over the part of the strip -- plus a 2-pixel
// margin per level, which is what the pyrUp of the next finer level reads -- that its footprint covers.  The owner
// adds the slabs and its own images in rank order inside the ordinary per-level kernel and collapses its strip.
// int16 wrap-around sums are exact under any grouping; the float weight sums are grouped per rank, which is
// bit-identical to the single-GPU order whenever at most two ranks meet in a pixel with at most one image each
// and otherwise differs by float re-association only (tests bound the effect on the final uint8 to +-1).
#include <algorithm>
#include <cstring>

#include "sb_shard.h"

namespace sb {

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}

void ShardPlan::region_x(const BlendPlan &plan, int r, int l, int *a, int *b) const
{
    const int wl = plan.wp >> l;
    *a = std::max(0, (bounds[r] >> l) - 2);
    *b = std::min(wl, (bounds[r + 1] >> l) + 2);
    if (l == 0) {  // nothing finer reads level 0: no margin, and nothing beyond the (even-rounded) roi
        *a = bounds[r];
        *b = std::min(bounds[r + 1], (plan.roi.w + 1) / 2 * 2);
    }
    if (*b < *a) *b = *a;
}

void ShardPlan::slab_geometry(const BlendPlan &plan, int src, int dst, PeerSlab *ps) const
{
    int f0, fc;
    block_of((int)plan.imgs.size(), world, src, &f0, &fc);
    size_t off = 0;
    std::memset(ps->lv, 0, sizeof ps->lv);
    ps->split = 0;
    for (int l = 0; l <= plan.nb; ++l) {
        SlabLevel &L = ps->lv[l];
        if (fc == 0) continue;
        // Bounding box of where the source rank's weights can be non-zero at this level.  W_0 is non-zero only inside
        // the fed image's own rect (the padding is a constant-0 border); every pyrDown grows that support by at most
        // 2 samples per side, i.e. it stays within 3 level-l pixels of the scaled image rect.  Outside it the partial
        // sums are exactly (acc, wsum) = (0, 0) and need not travel.  Clipped to the padded footprint, kept even.
        int fx0 = 1 << 30, fy0 = 1 << 30, fx1 = -1, fy1 = -1;
        const int even = l < plan.nb ? ~1 : ~0;
        for (int i = f0; i < f0 + fc; ++i) {
            const FeedImage &im = plan.imgs[i];
            const int X0 = im.px + im.left, Y0 = im.py + im.top, X1 = X0 + im.w, Y1 = Y0 + im.h;
            const int px0 = im.px >> l, py0 = im.py >> l, px1 = (im.px + im.pw) >> l, py1 = (im.py + im.ph) >> l;
            fx0 = std::min(fx0, std::max(px0, ((X0 >> l) - 3) & even));
            fy0 = std::min(fy0, std::max(py0, ((Y0 >> l) - 3) & even));
            fx1 = std::max(fx1, std::min(px1, ((((X1 - 1) >> l) + 4) + 1) & even));
            fy1 = std::max(fy1, std::min(py1, ((((Y1 - 1) >> l) + 4) + 1) & even));
        }
        int a, b;
        region_x(plan, dst, l, &a, &b);
        const int x0 = std::max(a, fx0), x1 = std::min(b, fx1);
        if (x1 <= x0 || fy1 <= fy0) continue;
        L.x0 = x0;
        L.y0 = fy0;
        L.w = x1 - x0;
        L.h = fy1 - fy0;
        L.pitch = (int)align_up((size_t)L.w, 8);
        L.plane = L.pitch * L.h;
        L.acc_off = off;
        off = align_up(off + (size_t)3 * L.plane * sizeof(int16_t), 256);
        L.w_off = off;
        off = align_up(off + (size_t)L.plane * sizeof(float), 256);
        if (l == 0) ps->split = off;
    }
    ps->bytes = off;
}

int ShardPlan::build(const BlendPlan &plan, int rank_, int world_)
{
    const int n = (int)plan.imgs.size();
    if (plan.kind != SB_BLEND_MULTIBAND || plan.nb < 1) {
        set_error("sharded composite: needs the multiband blender with at least one band (got kind %d, %d bands)", plan.kind, plan.nb);
        return SB_ERR_INVALID;
    }
    if (world_ < 1 || rank_ < 0 || rank_ >= world_ || n < world_) {
        set_error("sharded composite: rank %d of %d with %d images", rank_, world_, n);
        return SB_ERR_INVALID;
    }
    rank = rank_;
    world = world_;
    block_of(n, world, rank, &first, &count);
    // strip boundaries: halfway between the centres of the neighbouring blocks' edge images, snapped to 2^nb
    const int a = 1 << plan.nb;
    bounds.assign(world + 1, 0);
    bounds[world] = plan.wp;
    for (int r = 1; r < world; ++r) {
        int f0, fc, g0, gc;
        block_of(n, world, r - 1, &f0, &fc);
        block_of(n, world, r, &g0, &gc);
        const FeedImage &L = plan.imgs[f0 + fc - 1], &R = plan.imgs[g0];
        const long long mid = ((long long)L.px + L.left + L.w / 2 + R.px + R.left + R.w / 2) / 2;
        int b = (int)((mid + a / 2) / a * a);
        b = std::max(b, bounds[r - 1]);
        b = std::min(b, plan.wp);
        bounds[r] = b;
    }
    for (int r = 1; r <= world; ++r)
        if (bounds[r] < bounds[r - 1]) {
            set_error("sharded composite: image blocks are not ordered left to right in the panorama");
            return SB_ERR_INVALID;
        }
    send.assign(world, PeerSlab());
    recv.assign(world, PeerSlab());
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        slab_geometry(plan, rank, p, &send[p]);
        slab_geometry(plan, p, rank, &recv[p]);
    }
    return SB_OK;
}

void ShardPlan::strip(const BlendPlan &plan, int *lo, int *hi) const
{
    *lo = std::min(bounds[rank], plan.roi.w);
    *hi = std::min(bounds[rank + 1], plan.roi.w);
}

int ShardPlan::allocate(const BlendPlan &plan, cudaStream_t s)
{
    release(s);
    const int n = (int)plan.imgs.size();
    SB_TRY(connect(plan, s));  // direct exchange over NVLink when the ranks can map each other's memory (sb_peer.cpp)
    for (int p = 0; p < world; ++p) {
        if (send[p].bytes) SB_TRY(dev_alloc(&send[p].buf, send[p].bytes, s));
        if (!recv[p].bytes) continue;
        if (connected)
            recv[p].buf = (char *)arena + recv_off[p];
        else
            SB_TRY(dev_alloc(&recv[p].buf, recv[p].bytes, s));
    }
    // item lists per level: slabs of lower ranks, own images, slabs of higher ranks (= feed order)
    std::vector<ColDesc> items;
    size_t offs[SB_MAX_BANDS + 2] = {0};
    for (int l = 0; l <= plan.nb; ++l) {
        offs[l] = items.size();
        auto add_slab = [&](int p) {
            const SlabLevel &L = recv[p].lv[l];
            if (!recv[p].bytes || L.w == 0) return;
            ColDesc d;
            std::memset(&d, 0, sizeof d);
            d.ox = L.x0;
            d.oy = L.y0;
            d.w_l = L.w;
            d.h_l = L.h;
            d.g = (const int16_t *)((const char *)recv[p].buf + L.acc_off);
            d.w = (const float *)((const char *)recv[p].buf + L.w_off);
            d.pitch = L.pitch;
            d.plane = L.plane;
            d.kind = 1;
            items.push_back(d);
        };
        for (int p = 0; p < rank; ++p) add_slab(p);
        for (int i = first; i < first + count; ++i) items.push_back(plan.col_host[(size_t)l * n + i]);
        for (int p = rank + 1; p < world; ++p) add_slab(p);
        n_items[l] = (int)(items.size() - offs[l]);
        if (n_items[l] > SB_MAX_ITEMS) {
            set_error("sharded composite: %d items at level %d exceed SB_MAX_ITEMS", n_items[l], l);
            return SB_ERR_INVALID;
        }
    }
    SB_TRY(dev_alloc(&items_arena_, std::max<size_t>(items.size(), 1) * sizeof(ColDesc), s));
    if (!items.empty()) SB_CUDA(cudaMemcpyAsync(items_arena_, items.data(), items.size() * sizeof(ColDesc), cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaStreamSynchronize(s));  // `items` is a local
    for (int l = 0; l <= plan.nb; ++l) items_dev[l] = (ColDesc *)items_arena_ + offs[l];
    return SB_OK;
}

void ShardPlan::release(cudaStream_t s)
{
    for (auto &p : send) {
        dev_free(p.buf, s);
        p.buf = nullptr;
    }
    for (auto &p : recv) {
        if (!connected) dev_free(p.buf, s);
        p.buf = nullptr;
    }
    if (connected) {
        (void)cudaStreamSynchronize(s);
#ifndef SB_EMU
        for (auto &p : peer_arena)
            if (p) (void)cudaIpcCloseMemHandle(p);
        (void)cudaFree(arena);
#endif
        peer_arena.clear();
        arena = nullptr;
        connected = false;
    }
    dev_free(items_arena_, s);
    items_arena_ = nullptr;
}

int ShardPlan::partial_out(const BlendPlan &plan, cudaStream_t s, int l_lo, int l_hi, bool direct)
{
    direct = direct && connected;
    const int n = (int)plan.imgs.size();
    for (int p = 0; p < world; ++p) {
        if (p == rank || !send[p].bytes) continue;
        for (int l = std::max(l_lo, 0); l <= std::min(l_hi, plan.nb); ++l) {
            const SlabLevel &L = send[p].lv[l];
            if (L.w == 0) continue;
            CollapseArgs A;
            std::memset(&A, 0, sizeof A);
            A.col = plan.col_dev + (size_t)l * n + first;  // this rank's images, contiguous in the level's array
            A.n = count;
            A.rx0 = L.x0;
            A.ry0 = L.y0;
            A.rw = L.w;
            A.rh = L.h;
            A.partial = 1;
            // direct: the slab is stored straight into its place in the owner's arena (peer stores over NVLink)
            char *dst = direct ? peer_arena[p] + peer_slot[p] : (char *)send[p].buf;
            A.slab_acc = (int16_t *)(dst + L.acc_off);
            A.slab_w = (float *)(dst + L.w_off);
            A.slab_pitch = L.pitch;
            A.slab_plane = L.plane;
            SB_TRY(launch_collapse_fast(A, l, plan.nb, s));
        }
    }
    return SB_OK;
}

int ShardPlan::exchange(cudaStream_t s, int part)
{
    std::vector<int> peers;
    std::vector<void *> sp, rp;
    std::vector<size_t> sb_, rb;
    for (int p = 0; p < world; ++p) {
        if (p == rank || (!send[p].bytes && !recv[p].bytes)) continue;
        // byte range of the part: level 0 lies in front
        const size_t s0 = part == 1 ? send[p].split : 0, s1 = part == 0 ? send[p].split : send[p].bytes;
        const size_t r0 = part == 1 ? recv[p].split : 0, r1 = part == 0 ? recv[p].split : recv[p].bytes;
        if (s1 <= s0 && r1 <= r0) continue;
        peers.push_back(p);
        sp.push_back((char *)send[p].buf + s0);
        sb_.push_back(s1 > s0 ? s1 - s0 : 0);
        rp.push_back((char *)recv[p].buf + r0);
        rb.push_back(r1 > r0 ? r1 - r0 : 0);
    }
    if (peers.empty()) return SB_OK;
    return comm_exchange((int)peers.size(), peers.data(), sp.data(), sb_.data(), rp.data(), rb.data(), s);
}

int ShardPlan::finish(const BlendPlan &plan, const PanoOut &out, cudaStream_t s, int l_hi, int l_lo)
{
    int lo, hi;
    strip(plan, &lo, &hi);
    for (int l = std::min(l_hi, plan.nb); l >= std::max(l_lo, 0); --l) {
        int a, b;
        region_x(plan, rank, l, &a, &b);
        CollapseArgs A;
        std::memset(&A, 0, sizeof A);
        A.col = items_dev[l];
        A.n = n_items[l];
        if (l < plan.nb) A.up = plan.pano[l + 1];
        A.cur = plan.pano[l];
        A.rx0 = a;
        A.ry0 = 0;
        A.rw = b - a;
        A.rh = l == 0 ? (plan.roi.h + 1) / 2 * 2 : plan.hp >> l;
        A.out = out;
        A.out_x0 = bounds[rank];
        A.out_lo = lo;
        A.out_hi = hi;
        SB_TRY(launch_collapse_fast(A, l, plan.nb, s));
    }
    return SB_OK;
}

}  // namespace sb
