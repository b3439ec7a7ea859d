"""Multi-GPU plumbing (one process per GPU, torch.distributed for the rendezvous only).

Scan-parallel mode needs nothing from here (every rank registers its own scans against its own map replica).
Sharded mode: the map is cut by spatial cell (csrc/shard.cu: Morton-contiguous ranges of cells with equal point counts; every rank indexes its
cells plus the halo of the match gates); a rank owns the queries that fall into its cells (cell_owner below mirrors the device lookup), builds
residual blocks for those only, and the 29 normal-equation sums are all-reduced INSIDE the solver kernel
through peer-mapped staging buffers (CUDA IPC over NVLink).  The only host-side exchange is the one-time all-gather of the
64-byte IPC handles done here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def cell_index(points_world: np.ndarray, origin, cell_size: float, dims) -> np.ndarray:
    """Flat cell index (x fastest) of each world-frame point: clamp(floor((p - origin) * (1 / cell)), 0, dims - 1) per axis in float32, like
    shard_cell_index in csrc/common.cuh (the border cells extend outwards without bound)."""
    inv = np.float32(1.0) / np.float32(cell_size)
    o = np.asarray(origin, np.float32)
    d = np.asarray(dims, np.int64)
    with np.errstate(invalid="ignore"):
        ijk = np.floor((points_world[:, :3].astype(np.float32) - o) * inv)
    ijk = np.clip(np.nan_to_num(ijk, nan=0.0, posinf=1e9, neginf=-1e9), 0, d - 1).astype(np.int64)
    return (ijk[:, 2] * d[1] + ijk[:, 1]) * d[0] + ijk[:, 0]


def plan_shards(points_world: np.ndarray, world: int, cell_size: float):
    """The partition ll_map_build_sharded computes: grid over the bounding box of the finite points, points per cell, Morton-contiguous ranges of
    nearly equal point count (ll_shard_plan, pure host code in the library).  Returns (origin float32[3], dims int[3], owner int32[cells])."""
    p = points_world[:, :3].astype(np.float32)
    p = p[np.isfinite(p).all(1)]
    lo, hi = p.min(0), p.max(0)
    inv = np.float32(1.0) / np.float32(cell_size)
    dims = (np.floor((hi - lo) * inv).astype(np.int64) + 1).astype(np.int32)
    counts = np.bincount(cell_index(p, lo, cell_size, dims), minlength=int(dims.prod())).astype(np.int32)
    owner = np.zeros(int(dims.prod()), np.int32)
    d3 = (C.c_int * 3)(*[int(v) for v in dims])
    st = capi.lib().ll_shard_plan(counts.ctypes.data, d3, int(world), owner.ctypes.data)
    if st != capi.LL_OK:
        raise capi.LoamLivoxError(f"ll_shard_plan failed ({st})")
    return lo, dims, owner


def cell_owner(points_world: np.ndarray, origin, cell_size: float, dims, owner: np.ndarray) -> np.ndarray:
    """Rank owning each world-frame point (what knn_blocks_kernel looks up for a transformed feature)."""
    return owner[cell_index(points_world, origin, cell_size, dims)]


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
