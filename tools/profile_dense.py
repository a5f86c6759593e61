"""BASELINE configs[4] (dense single segment) on device-resident inputs, the command wrapped by ncu:
  ncu --set full --clock-control none --import-source on -k regex:k_degree -s 2 -c 1 -o gpurun_out/x python tools/profile_dense.py 0.9"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import scenes  # noqa: E402
from pbnet_b200.cluster import Context  # noqa: E402

f = float(sys.argv[1]) if len(sys.argv) > 1 else 0.9
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
xs, xo, sem = scenes.make_dense_case(7, n, f)
cols = [torch.from_numpy(np.ascontiguousarray(a[:, i])).cuda() for a in (xs, xo) for i in range(3)]
dsem = torch.from_numpy(sem).cuda()
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
ctx = Context(0, profiling=True)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ctx.binary_cluster(*cols, dsem, np.array([n], np.int32), r18, m18)
    torch.cuda.synchronize()
print("ms", (time.perf_counter() - t0) * 1e3, {k: round(v, 3) for k, v in ctx.stage_ms().items()}, ctx.counters())
