"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The (y, m) grid shards by m rows: blocks of 32 consecutive rows dealt to the ranks in a snake.
Cells are independent, so the only exchange on the table path is one all-gather of the packed
shards, after which every rank un-permutes the rows into its full device table and folds it
locally.  Event sampling shards by Philox counter ranges and needs no collective at all.

torch appears here as plumbing only: tensors alias the library's device buffers through
__cuda_array_interface__, no arithmetic is done by torch.
"""
from __future__ import annotations

import os

import numpy as np


class _DevBuf:
    """Exposes a raw device pointer through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def as_tensor(ptr: int, n: int, device):
    import torch
    return torch.as_tensor(_DevBuf(ptr, n), device=device)


def shard_block(nm: int, world: int) -> int:
    """shard_block() of csrc/upc_lumi.cu: 32 rows when every shard still gets two blocks, else a smaller power of 2."""
    b = 32
    while b > 1 and nm < 2 * b * world:
        b >>= 1
    return b


def shard_of_row(im: int, world: int, blk: int) -> int:
    """shard_of_row() of csrc/upc_lumi.cu: blocks dealt in a snake (cycle 0: shards 0..G-1, cycle 1: G-1..0, ...)."""
    jb = im // blk
    cyc, pos = divmod(jb, world)
    return world - 1 - pos if cyc & 1 else pos


def cyclic_rows(nm: int, rank: int, world: int):
    """m rows owned by `rank` (host-side logic shared with the CPU tests): blocks of shard_block() consecutive rows
    dealt in a snake over the ranks (the cost of a row falls with m; the snake pairs dear blocks with cheap ones);
    ascending."""
    blk = shard_block(nm, world)
    return [im for im in range(nm) if shard_of_row(im, world, blk) == rank]


def rows_per_shard(nm: int, world: int) -> int:
    """Rows of the packed shard buffer = the row count of the largest shard."""
    blk = shard_block(nm, world)
    cyc = blk * world
    return (nm // cyc) * blk + min(nm % cyc, blk)


def unpack_host(gathered: np.ndarray, nm: int, ny: int, world: int) -> np.ndarray:
    """Host restatement of upcgpu_lumi_unpack (for the gloo CPU tests of the layout logic):
    gathered [world][rows_per_shard][ny] -> full [nm][ny]."""
    rps = rows_per_shard(nm, world)
    g = np.asarray(gathered).reshape(world, rps, ny)
    full = np.zeros((nm, ny))
    for r in range(world):
        rows = cyclic_rows(nm, r, world)
        full[rows] = g[r, : len(rows)]
    return full


def init_from_env(backend="nccl"):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            import torch
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def setup_peer_exchange(gpu, rank: int, world: int) -> bool:
    """One-time set-up of the peer-store exchange between the ranks of one node: every rank's full lumi tables are
    mapped into every other rank through CUDA IPC (handles swapped with all_gather_object).  Returns False when a rank
    cannot map its peers (no peer access between the devices): the caller then stays on the NCCL all-gather."""
    import torch.distributed as dist
    mine = gpu.lumi_ipc_export()
    allh = [None] * world
    dist.all_gather_object(allh, mine)
    ok = True
    try:
        gpu.lumi_ipc_import(world, rank, b"".join(allh))
    except Exception:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    return all(flags)


def fill_lumi_peers(gpu, rank: int, world: int, device=None):
    """Table stage with the exchange folded into the cell kernel: every rank fills its m rows and its kernel stores
    each finished cell into the table of every rank (set up by setup_peer_exchange).  A one-element all-reduce on the
    library's stream orders every rank's fold behind every rank's kernel; nothing here waits for the device."""
    import torch
    import torch.distributed as dist
    gpu.fill_lumi_shard_peers()
    tok = getattr(gpu, "_peer_token", None)
    if tok is None:
        tok = (torch.zeros(1, device=device), torch.cuda.ExternalStream(gpu.stream_handle(), device=device))
        gpu._peer_token = tok
    with torch.cuda.stream(tok[1]):
        dist.all_reduce(tok[0])


def fill_lumi_distributed(gpu, rank: int, world: int, device=None):
    """Table stage on `world` GPUs: every rank queues the fill of its m rows (upcgpu_fill_lumi_shard, asynchronous),
    the packed shards are all-gathered with NCCL straight between the library's device buffers, and each rank
    un-permutes them into its full table.  Everything is queued on the library's stream; nothing here waits for the
    device (the fold that follows does, once).  Returns nothing; the table stays on device."""
    gpu.fill_lumi_shard(rank, world)
    if world == 1:
        return
    import torch
    import torch.distributed as dist
    # Zero-copy views of the library's shard / gather buffers.  The library reallocates them when the shard count
    # changes (an interleaved single-GPU fill does that), so the pointers are queried on every call and the views are
    # rebuilt whenever one of them moved.
    kinds = (1, 2) if gpu.P.use_pol else (0,)
    ptrs = []
    for which in kinds:
        sptr, sn = gpu.lumi_shard_buffer(which)
        gptr, gn = gpu.lumi_gather_buffer(which, world)
        ptrs.append((sptr, sn, gptr, gn))
    key = (world, tuple(ptrs), gpu.stream_handle())
    views = getattr(gpu, "_dist_views", None)
    if views is None or views[0] != key:
        pairs = [(as_tensor(sptr, sn, device), as_tensor(gptr, gn, device)) for sptr, sn, gptr, gn in ptrs]
        # the collective is issued on the library's own stream: the un-permute kernel that follows on that stream is
        # ordered behind it without a host synchronisation in between
        ext = torch.cuda.ExternalStream(gpu.stream_handle(), device=device)
        views = (key, pairs, ext)
        gpu._dist_views = views
    with torch.cuda.stream(views[2]):
        for src, dst in views[1]:
            dist.all_gather_into_tensor(dst, src)
    gpu.lumi_unpack(world)
