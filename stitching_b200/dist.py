"""Multi-GPU plumbing: one process per GPU, NCCL communicator for the sharded composite.

The library talks to NCCL itself (sb_comm_* in the C ABI); what it needs from the launcher is a way to hand the
NCCL unique id of rank 0 to the other ranks.  `init_comm` takes any broadcast callable; `init_comm_torchrun` uses a
torch.distributed process group (gloo is enough) when the job was started with torchrun -- plumbing only.
"""
import ctypes as C
import os

from . import _lib

ID_BYTES = 128


def init_comm(rank, world, broadcast_bytes, device=None):
    """broadcast_bytes(payload: bytes | None) -> bytes: returns rank 0's payload on every rank."""
    L = _lib.lib()
    _lib.check(L.sb_init(int(os.environ.get("LOCAL_RANK", rank)) if device is None else int(device)), "sb_init")
    if world == 1:
        return
    buf = (C.c_uint8 * ID_BYTES)()
    if rank == 0:
        _lib.check(L.sb_comm_unique_id(buf), "sb_comm_unique_id")
    payload = broadcast_bytes(bytes(buf) if rank == 0 else None)
    buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(payload)
    _lib.check(L.sb_comm_init(buf, rank, world), "sb_comm_init")


def init_comm_torchrun():
    """Under torchrun: (rank, world) after initialising the library's NCCL communicator."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        init_comm(0, 1, None)
        return 0, 1
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group(backend="gloo")

    def bcast(payload):
        box = [payload]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    init_comm(rank, world, bcast)
    return rank, world


def shutdown():
    _lib.lib().sb_comm_destroy()
