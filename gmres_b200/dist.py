"""torch.distributed plumbing for the multi-GPU (row-slab) path: one process per GPU.

The data path never goes through torch: halo send/recv and the scalar / (m+1)-vector
all-reduces are issued by libkrylov_b200 itself on its NCCL communicator
(kl_comm_init).  torch.distributed is only used to (1) hand the NCCL unique id from
rank 0 to the other ranks and (2) gather results in tests / benchmarks.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import Handle, load_library


def slab_partition(ny: int, rank: int, nranks: int):
    """Lines [j0, j0+ny_local) of a global grid owned by `rank` (kl_partition_rank)."""
    L = load_library()
    j0, nyl = C.c_int(), C.c_int()
    L.kl_partition_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = L.kl_partition_rank(int(ny), int(rank), int(nranks), C.byref(j0), C.byref(nyl))
    if rc != 0:
        raise ValueError(f"kl_partition_rank({ny}, {rank}, {nranks}) -> {rc}")
    return j0.value, nyl.value


def init_handle(device: int | None = None) -> Handle:
    """Create this rank's Handle and, if torch.distributed is initialised with more than
    one rank, its NCCL communicator (unique id broadcast from rank 0)."""
    import torch
    import torch.distributed as dist

    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    h = Handle(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [h.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        h.comm_init(rank, world, ids[0])
    return h


def local_slab(global_vec: np.ndarray, nx: int, ny: int, rank: int, nranks: int) -> np.ndarray:
    j0, nyl = slab_partition(ny, rank, nranks)
    return np.ascontiguousarray(global_vec.reshape(ny, nx)[j0:j0 + nyl].reshape(-1))
