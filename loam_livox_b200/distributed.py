"""Multi-GPU plumbing (one process per GPU, torch.distributed for the rendezvous only).

Scan-parallel mode needs nothing from here (every rank registers its own scans against its own map replica).
Sharded mode: every rank owns the queries whose 8 m cell hashes to it (cell_owner below mirrors the device function in
csrc/knn.cu), builds residual blocks for those only, and the 29 normal-equation sums are all-reduced INSIDE the solver kernel
through peer-mapped staging buffers (CUDA IPC over NVLink).  The only host-side exchange is the one-time all-gather of the
64-byte IPC handles done here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def cell_owner(points_world: np.ndarray, cell_size: float, world: int) -> np.ndarray:
    """Rank owning each world-frame point: hash of floor(p / cell) (float32 arithmetic, like csrc/knn.cu:cell_owner)."""
    inv = np.float32(1.0) / np.float32(cell_size)
    ijk = np.floor(points_world[:, :3].astype(np.float32) * inv).astype(np.int64)
    h = (ijk[:, 0].astype(np.uint32) * np.uint32(73856093)) ^ (ijk[:, 1].astype(np.uint32) * np.uint32(19349663)) ^ (ijk[:, 2].astype(np.uint32) * np.uint32(83492791))
    return (h % np.uint32(world)).astype(np.int32)


def all_gather_handles(local_handle: bytes, world: int, dist) -> bytes:
    """All-gather of fixed-size byte blobs over any torch.distributed backend (gloo on CPU, nccl on GPU)."""
    import torch
    assert len(local_handle) == capi.LL_IPC_HANDLE_BYTES
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(local_handle), dtype=torch.uint8, device=dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().tolist()) for t in out)


def connect(ctx, rank: int, world: int, dist) -> None:
    """Export this context's staging buffer, gather everyone's handle, map the peers (ll_comm_local_handle / ll_comm_connect)."""
    buf = (C.c_ubyte * capi.LL_IPC_HANDLE_BYTES)()
    ctx.check(ctx._lib.ll_comm_local_handle(ctx.h, buf))
    blob = all_gather_handles(bytes(buf), world, dist)
    arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    ctx.check(ctx._lib.ll_comm_connect(ctx.h, rank, world, arr))
