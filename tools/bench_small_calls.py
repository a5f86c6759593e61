"""Per-call latency of the drop-in path on small per-class problems (the reference's call pattern)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import pbnet_ops, scenes  # noqa: E402
from pbnet_b200.cluster import Context  # noqa: E402

sc = scenes.make_scene(22, 150000)
calls = scenes.class_calls(sc, 1)
ctx = Context(0)
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)


def bench(f, reps=20):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


for c in calls:
    n = len(c["sem"])
    xs, xo = c["xyz_shift"], c["xyz_orig"]
    cols = [np.ascontiguousarray(xs[:, i]) for i in range(3)] + [np.ascontiguousarray(xo[:, i]) for i in range(3)]
    sem = c["sem"].astype(np.int32)
    d = [torch.from_numpy(a).cuda() for a in cols] + [torch.from_numpy(sem).cuda()]
    t_dev = bench(lambda: ctx.binary_cluster(*d, c["seg_counts"], r18, m18))
    t_host = bench(lambda: ctx.binary_cluster(*cols, sem, c["seg_counts"], r18, m18))
    T = [torch.from_numpy(xs), torch.from_numpy(xo), torch.from_numpy(c["sem"]), torch.from_numpy(c["seg_counts"])]
    t_ops = bench(lambda: pbnet_ops.cluster(T[0], T[1], T[2], T[3], 0.04, 31, 1))
    Tc = [t.cuda() for t in T[:3]] + [T[3]]
    t_ops_dev = bench(lambda: pbnet_ops.cluster(Tc[0], Tc[1], Tc[2], Tc[3], 0.04, 31, 1))
    print(f"class {c['sem_id']:2d} n={n:6d}  C-ABI device {t_dev:7.0f} us  C-ABI host {t_host:7.0f} us  "
          f"pbnet_ops CPU tensors {t_ops:7.0f} us  pbnet_ops CUDA tensors {t_ops_dev:7.0f} us  launches {ctx.last_launch_count}")

# stage breakdown of the largest call
ctx.set_profiling(True)
c = max(calls, key=lambda c: len(c["sem"]))
xs, xo = c["xyz_shift"], c["xyz_orig"]
d = [torch.from_numpy(np.ascontiguousarray(a[:, i])).cuda() for a in (xs, xo) for i in range(3)] + [torch.from_numpy(c["sem"].astype(np.int32)).cuda()]
for _ in range(3):
    ctx.binary_cluster(*d, c["seg_counts"], r18, m18)
print("stages n=%d:" % len(c["sem"]), {k: round(v * 1e3) for k, v in ctx.stage_ms().items()}, "us")
c = min(calls, key=lambda c: abs(len(c["sem"]) - 2300))
xs, xo = c["xyz_shift"], c["xyz_orig"]
d = [torch.from_numpy(np.ascontiguousarray(a[:, i])).cuda() for a in (xs, xo) for i in range(3)] + [torch.from_numpy(c["sem"].astype(np.int32)).cuda()]
for _ in range(3):
    ctx.binary_cluster(*d, c["seg_counts"], r18, m18)
print("stages n=%d:" % len(c["sem"]), {k: round(v * 1e3) for k, v in ctx.stage_ms().items()}, "us")
