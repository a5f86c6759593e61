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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int):
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers it allocates
    next (first touch) and its copy submissions are local to the GPU's PCIe root: with one rank per GPU of a two-socket host,
    half of the ranks would otherwise pull their inputs across the socket interconnect.  Returns
    ``(node, n_cpus, previous_affinity)`` or None when the topology is not exposed (single node, container without /sys).
    ``os.sched_setaffinity(0, previous_affinity)`` undoes it."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus or cpus == prev:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus), prev
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


class Rank0Gather:
    """Variable-length gather of a 1-D tensor to rank 0 with the size exchange and the padded buffers set up
    ONCE (per-step cost = one buffer copy + one ``dist.gather``).  Works on any backend (nccl: device
    tensors, gloo: CPU tensors).

    ``narrow_to`` (e.g. ``torch.int16`` for cluster ids, which restart at 0 in every call and stay far below 32768)
    halves the bytes on the wire; the caller promises the values fit.  ``overlap=True`` issues the collective
    asynchronously on a side stream: ``__call__`` only waits until the local tensor has been copied into the send
    buffer (so the producer may overwrite it in the next step) and returns; the transfer itself overlaps the next step's
    kernels.  ``finish()`` waits for the last transfer and returns the per-rank views on rank 0."""

    def __init__(self, numel: int, dtype: torch.dtype, device, pad_value: int = -1, narrow_to: torch.dtype | None = None,
                 overlap: bool = False):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.numel = int(numel)
        self.wire_dtype = narrow_to or dtype
        self.overlap = bool(overlap)
        self._work = None
        self._last = None
        if self.world == 1:
            self.sizes = [self.numel]
            return
        n = torch.tensor([self.numel], dtype=torch.int64, device=device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n)
        self.sizes = [int(s.item()) for s in sizes]
        pad = max(self.sizes)
        self.send = torch.full((pad,), pad_value, dtype=self.wire_dtype, device=device)
        self.recv = [torch.empty(pad, dtype=self.wire_dtype, device=device) for _ in range(self.world)] if self.rank == 0 else None
        # a gather only moves bytes: the collectives see uint8 views (neither NCCL nor gloo has an int16 type)
        self._send_b = self.send.view(torch.uint8)
        self._recv_b = [r.view(torch.uint8) for r in self.recv] if self.rank == 0 else None
        self.cuda = torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device=device) if (self.cuda and self.overlap) else None
        self.copied = torch.cuda.Event() if self.side is not None else None

    def _views(self):
        if self.rank != 0:
            return None
        return [r[:s] for r, s in zip(self.recv, self.sizes)]

    def __call__(self, local: torch.Tensor):
        """Blocking mode: returns the list of per-rank tensors (views into the receive buffers) on rank 0, None elsewhere.
        Overlap mode: starts the gather and returns None; the result of the LAST call comes from ``finish()``."""
        if self.world == 1:
            self._last = [local]
            return self._last
        if self.side is None:
            self.send[:self.numel].copy_(local)
            if self.overlap:   # CPU / gloo: asynchronous work handle
                if self._work is not None:
                    self._work.wait()
                self._work = dist.gather(self._send_b, self._recv_b, dst=0, async_op=True)
                return None
            dist.gather(self._send_b, self._recv_b, dst=0)
            return self._views()
        main = torch.cuda.current_stream(self.send.device)
        self.side.wait_stream(main)                       # the producer of `local` has finished
        with torch.cuda.stream(self.side):
            if self._work is not None:
                self._work.wait()                         # the previous transfer no longer reads `send`
            self.send[:self.numel].copy_(local)           # (narrowing) copy, then the collective behind it
            self.copied.record(self.side)
            self._work = dist.gather(self._send_b, self._recv_b, dst=0, async_op=True)
        main.wait_event(self.copied)                      # `local` may be overwritten from here on
        return None

    def finish(self):
        """Waits for the outstanding transfer (overlap mode) and returns the per-rank views on rank 0."""
        if self.world == 1:
            return self._last
        if self._work is not None:
            if self.side is not None:
                with torch.cuda.stream(self.side):
                    self._work.wait()
                torch.cuda.current_stream(self.send.device).wait_stream(self.side)
            else:
                self._work.wait()
            self._work = None
        return self._views()


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
