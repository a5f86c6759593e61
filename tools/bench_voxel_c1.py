"""The `voxel` leg of bench.py alone (rows a12-a14 on the first 8 M points of the C1 workload): seconds instead of minutes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pbnet_b200 import scenes, workload  # noqa: E402

dev = torch.device("cuda", 0)
sizes = scenes.scene_sizes()
w = workload.build(range(96), sizes)
d_xo = [torch.from_numpy(w[k]).to(dev) for k in ("xo", "yo", "zo")]
import numpy as np  # noqa: E402

merged = len(sys.argv) > 1 and sys.argv[1] == "merged"   # all scenes in one grid (the round-1 / early round-2 measurement)
sop = None if merged else np.repeat(w["call_scene"], w["call_points"])
print(json.dumps(bench.voxel_bench(d_xo, int(w["n_points"]), dev, sop)))
