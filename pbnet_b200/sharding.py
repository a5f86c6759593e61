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


class Rank0Gather:
    """Variable-length gather of a 1-D tensor to rank 0 with the size exchange and the padded buffers set up
    ONCE (per-step cost = one buffer copy + one ``dist.gather``).  Works on any backend (nccl: device
    tensors, gloo: CPU tensors)."""

    def __init__(self, numel: int, dtype: torch.dtype, device, pad_value: int = -1):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.numel = int(numel)
        if self.world == 1:
            self.sizes = [self.numel]
            return
        n = torch.tensor([self.numel], dtype=torch.int64, device=device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n)
        self.sizes = [int(s.item()) for s in sizes]
        pad = max(self.sizes)
        self.send = torch.full((pad,), pad_value, dtype=dtype, device=device)
        self.recv = [torch.empty(pad, dtype=dtype, device=device) for _ in range(self.world)] if self.rank == 0 else None

    def __call__(self, local: torch.Tensor):
        """Returns the list of per-rank tensors (views into the receive buffers) on rank 0, None elsewhere."""
        if self.world == 1:
            return [local]
        self.send[:self.numel].copy_(local)
        dist.gather(self.send, self.recv, dst=0)
        if self.rank != 0:
            return None
        return [r[:s] for r, s in zip(self.recv, self.sizes)]


def gather_to_rank0(local: torch.Tensor, pad_value: int = -1):
    """One-shot form of Rank0Gather."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    return Rank0Gather(local.numel(), local.dtype, local.device, pad_value)(local)


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
