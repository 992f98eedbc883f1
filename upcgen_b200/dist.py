"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The (y, m) grid shards by m rows, cyclically (rank r owns im = r, r+G, ...).
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


def cyclic_rows(nm: int, rank: int, world: int):
    """m rows owned by `rank` (host-side logic shared with the CPU tests)."""
    return list(range(rank, nm, world))


def rows_per_shard(nm: int, world: int) -> int:
    return (nm + world - 1) // world


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


def fill_lumi_distributed(gpu, rank: int, world: int, device=None):
    """Table stage on `world` GPUs: every rank fills its cyclic m rows (upcgpu_fill_lumi_shard),
    the packed shards are all-gathered with NCCL straight between the library's device buffers, and
    each rank un-permutes them into its full table.  Returns nothing; the table stays on device."""
    gpu.fill_lumi_shard(rank, world)
    if world == 1:
        return
    import torch
    import torch.distributed as dist
    kinds = (1, 2) if gpu.P.use_pol else (0,)
    for which in kinds:
        sptr, sn = gpu.lumi_shard_buffer(which)
        gptr, gn = gpu.lumi_gather_buffer(which, world)
        src = as_tensor(sptr, sn, device)
        dst = as_tensor(gptr, gn, device)
        dist.all_gather_into_tensor(dst, src)
    torch.cuda.current_stream().synchronize()
    gpu.lumi_unpack(world)
