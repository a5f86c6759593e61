"""Builds the benchmark / test workload: the per-class grouping calls of many synthetic scenes
(pbnet_b200.scenes) concatenated into ONE batched problem — the layout
``pb_binary_cluster_batched`` consumes (SoA fp32 coordinates, int32 classes, segment and call tables).
"""
from __future__ import annotations

import hashlib
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import scenes


def _scene_calls(args):
    seed, n_points, copies = args
    sc = scenes.make_scene(seed, n_points)
    calls = scenes.class_calls(sc, copies)
    if not calls:
        z3 = np.zeros((0, 3), np.float32)
        return z3, z3, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)
    xs = np.concatenate([c["xyz_shift"] for c in calls])
    xo = np.concatenate([c["xyz_orig"] for c in calls])
    sem = np.concatenate([c["sem"] for c in calls]).astype(np.int32)
    seg = np.concatenate([c["seg_counts"] for c in calls]).astype(np.int32)
    csc = np.array([len(c["seg_counts"]) for c in calls], np.int32)
    return xs, xo, sem, seg, csc


def build(scene_ids, sizes, copies: int = 1, workers: int | None = None, cache_dir: str | None = "/tmp"):
    """scene_ids: indices into the synthetic val set (scene s uses seed 22+s and ``sizes[s]`` points).
    Returns dict(x,y,z,xo,yo,zo f32[N]; sem i32[N]; seg_counts i32[S]; call_seg_counts i32[Ncalls];
    call_scene i32[Ncalls]; call_points i64[Ncalls]; n_points N)."""
    scene_ids = [int(s) for s in scene_ids]
    key = hashlib.sha1(repr((scene_ids, [int(sizes[s]) for s in scene_ids], copies, 3)).encode()).hexdigest()[:16]
    path = os.path.join(cache_dir, f"pbnet_b200_workload_{key}.npz") if cache_dir else None
    if path and os.path.exists(path):
        d = np.load(path)
        return {k: d[k] for k in d.files}
    jobs = [(scenes.BASE_SEED + s, int(sizes[s]), copies) for s in scene_ids]
    if workers is None:
        workers = min(len(jobs), os.cpu_count() or 1, 32)
    if workers > 1:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            parts = list(ex.map(_scene_calls, jobs, chunksize=2))
    else:
        parts = [_scene_calls(j) for j in jobs]
    xs = np.concatenate([p[0] for p in parts])
    xo = np.concatenate([p[1] for p in parts])
    out = dict(
        x=np.ascontiguousarray(xs[:, 0]), y=np.ascontiguousarray(xs[:, 1]), z=np.ascontiguousarray(xs[:, 2]),
        xo=np.ascontiguousarray(xo[:, 0]), yo=np.ascontiguousarray(xo[:, 1]), zo=np.ascontiguousarray(xo[:, 2]),
        sem=np.concatenate([p[2] for p in parts]), seg_counts=np.concatenate([p[3] for p in parts]),
        call_seg_counts=np.concatenate([p[4] for p in parts]),
        call_scene=np.concatenate([np.full(len(p[4]), s, np.int32) for p, s in zip(parts, scene_ids)]),
    )
    seg_off = np.concatenate([[0], np.cumsum(out["call_seg_counts"])])
    segsum = np.concatenate([[0], np.cumsum(out["seg_counts"].astype(np.int64))])
    out["call_points"] = (segsum[seg_off[1:]] - segsum[seg_off[:-1]]).astype(np.int64)
    out["n_points"] = np.int64(len(out["sem"]))
    if path:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, **out)
        os.replace(tmp, path)
    return out


def shard_scenes(sizes, world_size: int):
    """Static size-balanced partition (longest-processing-time first) of scene indices over ranks:
    scenes are independent units (lib/PB_lib/src/pbnet/cluster.cu:57-110 loops segments independently)."""
    order = np.argsort(-np.asarray(sizes), kind="stable")
    load = np.zeros(world_size, np.int64)
    shards = [[] for _ in range(world_size)]
    for s in order:
        r = int(np.argmin(load))
        shards[r].append(int(s))
        load[r] += int(sizes[s])
    return [sorted(s) for s in shards]


def iter_calls(w):
    """Yields (call_index, point slice, segment slice) of a workload built by build()."""
    seg_off = np.concatenate([[0], np.cumsum(w["call_seg_counts"])])
    pt_off = np.concatenate([[0], np.cumsum(w["seg_counts"].astype(np.int64))])
    for c in range(len(w["call_seg_counts"])):
        s0, s1 = int(seg_off[c]), int(seg_off[c + 1])
        yield c, slice(int(pt_off[s0]), int(pt_off[s1])), slice(s0, s1)
