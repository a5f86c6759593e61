"""CPU, world_size 2, gloo: scene sharding + the gather of cluster ids to rank 0 reproduce the
single-process result.  The per-rank compute is the CPU oracle here (the CUDA path needs a GPU); the
sharding / gather code is the one bench.py runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_calls(w):
    from oracle import pb_oracle as po
    from pbnet_b200 import workload
    from tests import helpers as H
    ids = np.empty(int(w["n_points"]), np.int32)
    for c, ps, ss in workload.iter_calls(w):
        xs = np.stack([w["x"][ps], w["y"][ps], w["z"][ps]], 1)
        xo = np.stack([w["xo"][ps], w["yo"][ps], w["zo"][ps]], 1)
        ids[ps] = po.oracle_binary_cluster(xs, xo, w["sem"][ps], w["seg_counts"][ss], H.R18, H.M18)["cluster_id"]
    return ids


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pbnet_b200 import sharding, workload
    shards = workload.shard_scenes(sizes, world)
    w = workload.build(shards[rank], sizes, 1, workers=1, cache_dir=None)
    ids = torch.from_numpy(_run_calls(w))
    got = sharding.gather_to_rank0(ids)
    # the overlapped, int16-on-the-wire form bench.py uses over NCCL: two steps in flight, result of the last one
    g16 = sharding.Rank0Gather(ids.numel(), torch.int32, "cpu", narrow_to=torch.int16, overlap=True)
    g16(torch.zeros_like(ids))
    assert g16(ids) is None
    got16 = g16.finish()
    if rank == 0:
        assert all(a.dtype == torch.int16 and torch.equal(a.to(torch.int32), b) for a, b in zip(got16, got))
    scene = sharding.gather_to_rank0(torch.from_numpy(w["call_scene"].astype(np.int32)))
    pts = sharding.gather_to_rank0(torch.from_numpy(w["call_points"].astype(np.int32)))
    if rank == 0:
        merged = sharding.merge_scene_results(shards, [s.numpy() for s in scene], [p.numpy() for p in pts], got)
        q.put({k: [a.copy() for a in v] for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from pbnet_b200 import scenes, workload
    sizes = np.minimum(scenes.scene_sizes(6), 15000)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = workload.build(range(6), sizes, 1, workers=1, cache_dir=None)
    ids = _run_calls(w)
    o = 0
    per_scene = {}
    for s, n in zip(w["call_scene"], w["call_points"]):
        per_scene.setdefault(int(s), []).append(ids[o:o + int(n)])
        o += int(n)
    assert sorted(merged) == sorted(per_scene)
    for s in per_scene:
        assert len(merged[s]) == len(per_scene[s])
        for a, b in zip(merged[s], per_scene[s]):
            assert np.array_equal(a, b)
