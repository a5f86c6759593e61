"""Scene sharding across ranks (one process per GPU) and the only collective of the path: the gather
of per-point cluster ids to rank 0.  Scenes are independent units — the reference loops segments
independently (lib/PB_lib/src/pbnet/cluster.cu:57-110) and classes independently
(network/PBNet.py:151) — so there is no data-path collective; NCCL (or gloo in the CPU tests) only
moves results."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .workload import shard_scenes  # noqa: F401  (re-export)


def gather_to_rank0(local: torch.Tensor, pad_value: int = -1):
    """Variable-length gather of a 1-D tensor to rank 0.  Returns the list of per-rank tensors on rank 0,
    None elsewhere.  Works on any backend (nccl: device tensors, gloo: CPU tensors)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    send = torch.full((pad,), pad_value, dtype=local.dtype, device=local.device)
    send[:local.numel()] = local
    recv = [torch.empty(pad, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0)
    if rank != 0:
        return None
    return [r[:s] for r, s in zip(recv, sizes)]


def merge_scene_results(shards, per_rank_call_scene, per_rank_call_points, per_rank_ids):
    """Rank-0 helper: reassembles per-scene cluster-id arrays from the per-rank gathers.
    Returns {scene_index: [ids of call 0, ids of call 1, ...]}."""
    out = {}
    for r, scenes_r in enumerate(shards):
        ids = per_rank_ids[r]
        ids = ids.cpu().numpy() if isinstance(ids, torch.Tensor) else np.asarray(ids)
        o = 0
        for s, npts in zip(per_rank_call_scene[r], per_rank_call_points[r]):
            out.setdefault(int(s), []).append(ids[o:o + int(npts)])
            o += int(npts)
    return out
