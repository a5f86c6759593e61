import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from pbnet_b200 import scenes
from pbnet_b200.cluster import Context
sc = scenes.make_scene(22, 150000)
r18 = np.full(18, np.float32(0.04), np.float32); m18 = np.full(18, 31, np.int32)
calls = [c for c in scenes.class_calls(sc, 1) if len(c["sem"]) > 12000]
for prof in (False, True):
    ctx = Context(0, profiling=prof)
    for c in calls:
        xs, xo = c["xyz_shift"], c["xyz_orig"]
        cols = [np.ascontiguousarray(xs[:, i]) for i in range(3)] + [np.ascontiguousarray(xo[:, i]) for i in range(3)]
        sem = c["sem"].astype(np.int32)
        for _ in range(3): out = ctx.binary_cluster(*cols, sem, c["seg_counts"], r18, m18)
        t0 = time.perf_counter()
        for _ in range(20): out = ctx.binary_cluster(*cols, sem, c["seg_counts"], r18, m18)
        dt = (time.perf_counter() - t0) / 20 * 1e6
        # device-resident variant
        dcols = [torch.from_numpy(a).cuda() for a in cols]; dsem = torch.from_numpy(sem).cuda()
        for _ in range(3): out = ctx.binary_cluster(*dcols, dsem, c["seg_counts"], r18, m18)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): out = ctx.binary_cluster(*dcols, dsem, c["seg_counts"], r18, m18)
        torch.cuda.synchronize(); dd = (time.perf_counter() - t0) / 20 * 1e6
        st = {k: round(v * 1e3) for k, v in ctx.stage_ms().items()} if prof else {}
        print(f"n={len(sem)} prof={prof} host-arrays {dt:.0f} us  device-arrays {dd:.0f} us launches {ctx.last_launch_count} stages(us) {st} sum {sum(st.values())}"
              + (f" centre halves {ctx.counters()['centre_halves']} replayed {ctx.counters()['centre_halves_replayed']} replay cycles {ctx.counters()['centre_replay_cycles']} gather cycles {ctx.counters()['centre_gather_cycles']} clusters {out['n_clusters']} sizes {np.bincount(out['cluster_id'][out['cluster_id'] >= 0] if not hasattr(out['cluster_id'], 'cpu') else out['cluster_id'].cpu().numpy()[out['cluster_id'].cpu().numpy() >= 0]).tolist()}" if prof else ""))
