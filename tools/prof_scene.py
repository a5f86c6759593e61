"""One scene x 3 rotated copies through grouping.group_instances / grouping.propose (the per-forward unit of PBNet eval):
wall-clock per call and the stage breakdown; wrap in ncu for the launch list."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import grouping, scenes  # noqa: E402
from pbnet_b200.cluster import default_context  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
sizes = scenes.scene_sizes(312)
sc0 = scenes.make_scene(scenes.BASE_SEED, int(sizes[0]))
t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
xyz3 = t_(np.concatenate(scenes.rotate_copies(sc0["xyz_orig"], 3)))
off3 = t_(np.concatenate(scenes.rotate_copies(sc0["offset"], 3)))
sem3 = t_(np.tile(sc0["sem"], 3))
bh3 = t_(np.repeat(np.arange(3), sc0["sem"].shape[0]).astype(np.int32))
gen = torch.Generator(device=dev).manual_seed(22)
feat3 = torch.rand((xyz3.shape[0], 32), device=dev, generator=gen)
sfp3 = torch.softmax(torch.rand((xyz3.shape[0], 20), device=dev, generator=gen), dim=1)
ctx = default_context(0)
for name, f in (("group_instances", lambda: grouping.group_instances(xyz3, off3, sem3, bh3, 0.04, 31, 3)),
                ("propose", lambda: grouping.propose(xyz3, off3, sem3, bh3, feat3, sfp3, 0.04, 31, 3))):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "points", int(xyz3.shape[0]), "ms", [round(t, 3) for t in ts], "launches(last C call)", ctx.last_launch_count)
ctx.set_profiling(True)
grouping.group_instances(xyz3, off3, sem3, bh3, 0.04, 31, 3)
print("stages us:", {k: round(v * 1e3) for k, v in ctx.stage_ms().items()})
