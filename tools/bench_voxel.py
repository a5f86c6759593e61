"""Micro-benchmark of the HBM-bound voxel kernels (rows a12-a14): voxelize, devoxelize gather, scatter-add."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import voxel  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
rng = np.random.default_rng(0)
xyz = torch.from_numpy((rng.uniform(0, 8, size=(n, 3)) * [1, 0.75, 0.4]).astype(np.float32)).to(dev)
xyz = (torch.round(xyz / 0.02 * 0.5) * 0.04).contiguous()  # ~2 points per voxel on average
PEAK = 6553.6


def timeit(f, reps=5):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t = timeit(lambda: voxel.voxel_map(xyz, 0.02), 3)
vm = voxel.voxel_map(xyz, 0.02)
print(f"voxelize: {n / t / 1e6:.0f} M points/s ({t * 1e3:.2f} ms, {vm.n_voxels} voxels)")
for C in (76, 32, 20, 3):
    vf = torch.randn((vm.n_voxels, C), device=dev)
    t = timeit(lambda: voxel.devoxelize_raw(vf, vm.inverse))
    gb = (n * C * 4 + vm.n_voxels * C * 4 + n * 8) / 1e9
    print(f"devoxelize C={C}: {t * 1e3:.3f} ms  {gb / t:.0f} GB/s = {gb / t / PEAK:.2f} of measured copy peak")
    g = torch.randn((n, C), device=dev)
    t = timeit(lambda: voxel.voxel_rows(g, vm, "sum"))
    gb = (n * C * 4 + vm.n_voxels * C * 4 + n * 4 + vm.n_voxels * 4) / 1e9
    print(f"scatter-add C={C}: {t * 1e3:.3f} ms  {gb / t:.0f} GB/s = {gb / t / PEAK:.2f}")
