"""One-process-per-GPU plumbing for the plane-slab sharding of a plan (torch.distributed is plumbing only).

The reference's plane loop (plane_slicer_raster_planner.cpp:596-628) carries no state between planes, so a lattice of
P planes is cut into `world` contiguous slabs; rank r plans lattice planes [P r / world, P (r + 1) / world) — the same
split nb2_plan applies for (shard_rank, shard_count) — and the per-slab polylines are concatenated in rank order, which
is plane order.  The only exchange is that gather (NCCL inside libnoether_b200.so on GPUs; this module also carries a
torch.distributed object gather used by the CPU/gloo tests of the host logic).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


def slab_range(num_planes: int, rank: int, world: int) -> tuple[int, int]:
    """Lattice planes [begin, end) of `rank` — identical to the split in nb2_plan (api.cu)."""
    return num_planes * rank // world, num_planes * (rank + 1) // world


@dataclass
class RankInfo:
    rank: int
    local_rank: int
    world: int

    @classmethod
    def from_env(cls) -> "RankInfo":
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
                   int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(info: RankInfo, backend: str):
    """torch.distributed rendezvous from the torchrun environment (MASTER_ADDR / MASTER_PORT)."""
    import torch.distributed as dist
    if info.world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group(backend=backend, rank=info.rank, world_size=info.world)


def max_over_ranks(value: float, world: int, device=None) -> float:
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(payload: bytes | None, src: int, world: int) -> bytes:
    if world <= 1:
        return payload
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


@dataclass
class SlabPaths:
    """Flattened tool paths of one slab (the SoA layout of nb2_result_view)."""
    path_offsets: np.ndarray
    segment_offsets: np.ndarray
    path_plane: np.ndarray
    positions: np.ndarray
    normals: np.ndarray
    edge_lo: np.ndarray
    edge_hi: np.ndarray


def concatenate_slabs(slabs: list[SlabPaths]) -> SlabPaths:
    """Rank-ordered concatenation = plane-ordered tool paths (what nb2_result_gather builds on the root)."""
    po, so = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)]
    s_acc = w_acc = 0
    for s in slabs:
        po.append(s.path_offsets[1:].astype(np.uint64) + s_acc)
        so.append(s.segment_offsets[1:].astype(np.uint64) + w_acc)
        s_acc += int(s.path_offsets[-1])
        w_acc += int(s.segment_offsets[-1])
    cat = lambda name, dt, shape: (np.concatenate([getattr(s, name) for s in slabs]) if slabs
                                   else np.zeros(shape, dt))
    return SlabPaths(np.concatenate(po), np.concatenate(so), cat("path_plane", np.int64, 0),
                     cat("positions", np.float32, (0, 3)), cat("normals", np.float32, (0, 3)),
                     cat("edge_lo", np.uint32, 0), cat("edge_hi", np.uint32, 0))


def gather_slabs(local: SlabPaths, info: RankInfo, root: int = 0) -> SlabPaths | None:
    """Host-side gather over torch.distributed (any backend).  Returns the concatenation on `root`, None elsewhere."""
    if info.world <= 1:
        return local
    import torch.distributed as dist
    out = [None] * info.world if info.rank == root else None
    dist.gather_object(local, out, dst=root)
    if info.rank != root:
        return None
    return concatenate_slabs(out)


def lpt_assign(sizes: list[int], world: int) -> list[list[int]]:
    """Whole-mesh sharding of a batch (partitioning B): longest-processing-time-first on the triangle
    counts — meshes in decreasing size, each to the rank with the least work so far.  Returns, per rank, the indices of
    its meshes in the order it should plan them.  Deterministic (ties: lower mesh index first, lower rank first), so
    every rank computes the same assignment without communicating."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += sizes[i]
    return out
