// sb_shard.h -- one panorama composited by several GPUs (one process per GPU): image blocks per rank, pano column
// strips per rank, exchange of the per-band partial sums where footprints cross a strip boundary.
#pragma once
#include <vector>

#include "sb_plan.h"

namespace sb {

struct SlabLevel {
    int x0, y0, w, h;       // rectangle at this level, pano level coordinates (w == 0: nothing at this level)
    int pitch, plane;       // elements
    size_t acc_off, w_off;  // byte offsets inside the slab buffer
};
struct PeerSlab {
    SlabLevel lv[SB_MAX_BANDS + 1];
    size_t bytes = 0;       // 0: no exchange with this peer in this direction
    size_t split = 0;       // the level-0 part is [0, split), the coarser levels [split, bytes)
    void *buf = nullptr;    // device
};

class ShardPlan {
public:
    int rank = 0, world = 1;
    int first = 0, count = 0;        // this rank's image block [first, first + count)
    std::vector<int> bounds;         // strip boundaries in padded-pano columns, world + 1 entries, multiples of 2^nb
                                     // (feather: in roi columns or rows, see `axis`)
    bool reversed = false;           // feather: the blocks follow each other against the axis (rank 0 owns the LAST strip)
    int axis = 0;                    // 0: column strips (multiband always); 1: row strips (feather, image blocks stacked vertically)
    std::vector<PeerSlab> send, recv;  // indexed by peer rank
    ColDesc *items_dev[SB_MAX_BANDS + 1] = {};
    TileDesc *tile_items_dev[SB_MAX_BANDS + 1] = {};  // the same lists for the tile kernels (null: images do not qualify)
    int n_items[SB_MAX_BANDS + 1] = {};
    // Direct exchange over NVLink (sb_peer.cpp): every rank's receive slabs live in ONE cudaMalloc'ed arena that its
    // neighbours map with CUDA IPC; the partial-sum kernels then store their slabs straight into the owner's arena
    // (peer stores, no copy, no NCCL call on the data path) and the ranks order themselves with flags written and
    // awaited by stream memory operations.  peer_arena[p] == nullptr: not connected (the NCCL exchange is used).
    void *arena = nullptr;                 // recv slabs of all peers (rank order) + the flags
    size_t arena_bytes = 0, flags_off = 0;
    std::vector<size_t> recv_off;          // slab from peer p inside MY arena
    std::vector<char *> peer_arena;        // peer p's arena mapped into this process
    std::vector<size_t> peer_slot;         // offset of MY slab inside peer p's arena
    std::vector<size_t> peer_flags_off;    // offset of peer p's flags inside its arena
    bool connected = false;
    unsigned step = 0;                     // composites issued so far (flag values)

    static void block_of(int n_images, int world, int r, int *first, int *count)
    {
        *first = (int)((long long)r * n_images / world);
        *count = (int)((long long)(r + 1) * n_images / world) - *first;
    }
    // geometry from the plan of ALL images (identical on every rank)
    int build(const BlendPlan &plan, int rank, int world);
    // slab buffers and per-level item lists (own images + the slabs this rank receives), after plan.allocate()
    int allocate(const BlendPlan &plan, cudaStream_t s);
    void release(cudaStream_t s);
    // output columns (axis 0) / rows (axis 1) of this rank: [lo, hi) in pano-roi coordinates (may be empty)
    void strip(const BlendPlan &plan, int *lo, int *hi) const;
    // feather blender (single level): partial sums of the own images over the rectangle each neighbour needs; the own
    // strip from the own images and the neighbours' slabs in rank order (sb_blend.cu k_feather_region)
    int feather_partial_out(const BlendPlan &plan, cudaStream_t s, bool direct = false);
    int feather_finish(const BlendPlan &plan, const PanoOut &out, cudaStream_t s);
    // phase 0: partial sums of the own images over every region a neighbour needs -> send slabs
    // (levels l_lo .. l_hi only: level 0 needs just the first pyrDown, so its slabs -- three quarters of the bytes --
    // can leave while the rest of the pyramid is still being built)
    int partial_out(const BlendPlan &plan, cudaStream_t s, int l_lo = 0, int l_hi = SB_MAX_BANDS, bool direct = false);
    // direct exchange: layout of rank `dst`'s arena (same on every rank), IPC connection, flags
    size_t arena_layout(const BlendPlan &plan, int dst, std::vector<size_t> *off, size_t *flags) const;
    int connect(const BlendPlan &plan, cudaStream_t s);            // collective over the NCCL communicator
    int push(cudaStream_t s, int part);                            // copy-engine copies of the send slabs into the owners' arenas
    bool direct_stores = false;                                    // SB_PEER=direct: the kernels store into the peers' arenas themselves
    int signal_data(cudaStream_t s, int part, unsigned value);     // part 0: level-0 slabs written, 1: the coarser levels
    int wait_data(cudaStream_t s, int part, unsigned value);
    int signal_consumed(cudaStream_t s, unsigned value);           // this rank has read the slabs of step `value`
    int wait_consumed(cudaStream_t s, unsigned value);             // ... before the next ones overwrite them
    // the NCCL exchange of the slabs (grouped send/recv on stream s); part 0: the level-0 part of every slab,
    // part 1: the coarser levels, part -1: everything
    int exchange(cudaStream_t s, int part = -1);
    // phase 1: sum slabs + own images in rank order, normalise, collapse the own strip; `out` is strip-local
    // (levels l_hi down to l_lo)
    int finish(const BlendPlan &plan, const PanoOut &out, cudaStream_t s, int l_hi = SB_MAX_BANDS, int l_lo = 0);

private:
    void region_x(const BlendPlan &plan, int r, int l, int *a, int *b) const;
    void slab_geometry(const BlendPlan &plan, int src, int dst, PeerSlab *ps) const;
    void feather_slab_geometry(const BlendPlan &plan, int src, int dst, PeerSlab *ps) const;
    int build_feather(const BlendPlan &plan);
    void *feather_items_ = nullptr;  // FeatherSlab records of the slabs this rank receives (device)
    int feather_before_ = 0, feather_after_ = 0;
    void *items_arena_ = nullptr;
    void *tile_items_arena_ = nullptr;
};

int comm_allgather_bytes(const void *mine, void *all, size_t bytes_per_rank, cudaStream_t s);  // host buffers
int comm_rank();
int comm_world();
int comm_exchange(int n, const int *peers, void *const *sendp, const size_t *sendb, void *const *recvp, const size_t *recvb,
                  cudaStream_t s);
bool comm_ready();

}  // namespace sb
