"""Multi-GPU sharded composite: host logic and kernel roles without GPUs, and the rendezvous plumbing.

The product sources run through tests/emu (serial CUDA stand-in); the slabs that NCCL would move are copied by the
test through the transport hooks (sb_compositor_shard_phase / _shard_slab).  The NCCL path itself is covered on
real GPUs by tests/test_gpu_sharded.py.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import replay
from conftest import ROOT
from stitching_b200 import Compositor, rigs


def run_sharded(cfg, cams, imgs, world, copy):
    sizes = [(cfg["w"], cfg["h"])] * len(cams)
    ranks = [Compositor(cams, sizes, cfg["warper"], cfg["blender"], cfg["strength"], rank=r, world=world) for r in range(world)]
    assert sum(c.count for c in ranks) == len(cams) and [c.first for c in ranks] == sorted(c.first for c in ranks)
    for c in ranks:
        c.upload(imgs[c.first: c.first + c.count])
        c.shard_phase(0)
    moved = 0
    for src in ranks:
        for dst in ranks:
            if src is dst:
                continue
            sp, sn = src.shard_slab(dst.rank, True)
            dp, dn = dst.shard_slab(src.rank, False)
            assert sn == dn, "both sides must agree on the slab size"
            if sn:
                copy(dp, sp, sn)
                moved += sn
    strips = []
    for c in ranks:
        c.shard_phase(1)
        strips.append(c.download())
    axis = ranks[0].strip_axis
    assert all(c.strip_axis == axis for c in ranks)
    order = sorted(range(world), key=lambda r: ranks[r].strip)  # spatial order of the strips (feather may run against the rank order)
    los = [ranks[r].strip for r in order]
    assert los[0][0] == 0 and all(a[1] == b[0] for a, b in zip(los, los[1:])) and los[-1][1] == ranks[0].roi[3 if axis else 2]
    pano = np.concatenate([strips[r][0] for r in order], axis=0 if axis else 1)
    mask = np.concatenate([strips[r][1] for r in order], axis=0 if axis else 1)
    for c in ranks:
        c.close()
    return pano, mask, moved


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_equals_single(use_emu, oracle, world):
    cfg = rigs.config("cfg2", 16)
    cams = cfg["cameras"]
    imgs = [rigs.noise_image(cfg["h"], cfg["w"], 70 + i) if i % 2 else rigs.synth_image(cfg["h"], cfg["w"], 70 + i) for i in range(len(cams))]
    single = Compositor(cams, [(cfg["w"], cfg["h"])] * len(cams), cfg["warper"], cfg["blender"], cfg["strength"])
    ref_pano, ref_mask = single.composite(imgs)
    assert single.num_bands >= 2
    single.close()
    pano, mask, moved = run_sharded(cfg, cams, imgs, world, lambda d, s, n: C.memmove(d, s, n))
    assert moved > 0
    assert np.array_equal(mask, ref_mask)
    d = np.abs(pano.astype(np.int32) - ref_pano.astype(np.int32))
    # int16 sums are exact under regrouping; the float weight sums are regrouped per rank: at most +-1 in the uint8
    assert d.max() <= 1, int(d.max())
    assert (d != 0).mean() < 1e-3


def test_sharded_cylindrical_many_images(use_emu):
    cfg = rigs.config("cfg3", 20)  # 32 images, cylindrical (BASELINE configs[2] layout: 4 images per rank at world 8)
    cams = cfg["cameras"]
    imgs = [rigs.synth_image(cfg["h"], cfg["w"], 5 + i) for i in range(len(cams))]
    single = Compositor(cams, [(cfg["w"], cfg["h"])] * len(cams), cfg["warper"], cfg["blender"], cfg["strength"])
    ref_pano, ref_mask = single.composite(imgs)
    single.close()
    pano, mask, _ = run_sharded(cfg, cams, imgs, 8, lambda d, s, n: C.memmove(d, s, n))
    assert np.array_equal(mask, ref_mask)
    assert np.abs(pano.astype(np.int32) - ref_pano.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_feather_grid(use_emu, world):
    """BASELINE configs[4]: 4x4 plane grid, feather blender; image blocks are grid rows, so the ranks own ROW strips."""
    cfg = rigs.config("cfg5", 12)
    cams = cfg["cameras"]
    imgs = [rigs.noise_image(cfg["h"], cfg["w"], 30 + i) if i % 3 == 0 else rigs.synth_image(cfg["h"], cfg["w"], 30 + i) for i in range(len(cams))]
    single = Compositor(cams, [(cfg["w"], cfg["h"])] * len(cams), cfg["warper"], cfg["blender"], cfg["strength"])
    ref_pano, ref_mask = single.composite(imgs)
    single.close()
    pano, mask, moved = run_sharded(cfg, cams, imgs, world, lambda d, s, n: C.memmove(d, s, n))
    assert moved > 0
    assert pano.shape == ref_pano.shape
    assert np.array_equal(mask, ref_mask)
    d = np.abs(pano.astype(np.int32) - ref_pano.astype(np.int32))
    assert d.max() <= 1, int(d.max())  # float sums regrouped per rank
    assert (d != 0).mean() < 1e-3


def test_sharded_feather_side_by_side(use_emu):
    """Feather on a single-row rig: the blocks lie side by side and the strips are columns, as for multiband."""
    cfg = dict(rigs.config("cfg2", 16), blender="feather")
    cams = cfg["cameras"]
    imgs = [rigs.synth_image(cfg["h"], cfg["w"], 11 + i) for i in range(len(cams))]
    single = Compositor(cams, [(cfg["w"], cfg["h"])] * len(cams), cfg["warper"], "feather", cfg["strength"])
    ref_pano, ref_mask = single.composite(imgs)
    single.close()
    pano, mask, _ = run_sharded(cfg, cams, imgs, 2, lambda d, s, n: C.memmove(d, s, n))
    assert np.array_equal(mask, ref_mask)
    assert np.abs(pano.astype(np.int32) - ref_pano.astype(np.int32)).max() <= 1


def test_sharded_rejects_what_it_cannot_do(use_emu):
    from stitching_b200 import StitchingError

    cfg = rigs.config("cfg5", 12)  # "no" blender: nothing to exchange, not sharded
    with pytest.raises(StitchingError):
        Compositor(cfg["cameras"], [(cfg["w"], cfg["h"])] * cfg["n"], cfg["warper"], "no", cfg["strength"], rank=0, world=2)
    cfg = rigs.config("cfg2", 16)
    c = Compositor(cfg["cameras"], [(cfg["w"], cfg["h"])] * cfg["n"], cfg["warper"], cfg["blender"], cfg["strength"], rank=1, world=2)
    with pytest.raises(StitchingError):
        c.upload([rigs.noise_image(cfg["h"], cfg["w"], 0)] * cfg["n"])  # only the own block may be uploaded
    c.close()


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the rendezvous used by stitching_b200.dist.init_comm_torchrun: rank 0's 128-byte id reaches every rank
payload = bytes(range(128)) if rank == 0 else None
box = [payload]
dist.broadcast_object_list(box, src=0)
assert box[0] == bytes(range(128))
# every rank derives the same partition and strip geometry from the same rig (host logic, emulation build)
from stitching_b200 import _lib, Compositor, rigs
_lib._lib = _lib.bind(os.path.join(sys.argv[1], "tests", "emu", "libstitch_b200_emu.so"))
cfg = rigs.config("cfg2", 16)
c = Compositor(cfg["cameras"], [(cfg["w"], cfg["h"])] * cfg["n"], cfg["warper"], cfg["blender"], cfg["strength"], rank=rank, world=world)
info = [None] * world
dist.all_gather_object(info, (c.first, c.count, c.strip, [c.shard_slab(p, True)[1] for p in range(world) if p != rank],
                              [c.shard_slab(p, False)[1] for p in range(world) if p != rank]))
if rank == 0:
    assert [i[0] for i in info] == [0, 4] and [i[1] for i in info] == [4, 4]
    assert info[0][2][1] == info[1][2][0]          # strips meet
    assert info[0][3] == info[1][4] and info[1][3] == info[0][4]   # send size on one side == receive size on the other
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_world_size_2_gloo_rendezvous_and_partition(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29613", str(script), ROOT],
        capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
