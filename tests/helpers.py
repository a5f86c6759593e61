"""Shared helpers of the parity tests: run the CUDA path / the oracle on the same inputs and compare.
Bar (BASELINE.json north_star): every integer output bit-exact; centres are replayed in the reference's
summation order, so they are compared bit-exact too (stricter than the 1e-5 the north star allows)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
R18 = np.full(18, np.float32(0.04), np.float32)
M18 = np.full(18, 31, np.int32)


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(path):
    d = np.load(path)
    ref = {k[4:]: d[k] for k in d.files if k.startswith("ref_")}
    return d, ref


def is_single_class(sem, seg_counts):
    o = 0
    for c in seg_counts:
        s = sem[o:o + int(c)]
        if len(s) and s.min() != s.max():
            return False
        o += int(c)
    return True


def run_cuda(ctx, xyz_shift, xyz_orig, sem, seg_counts, radius=R18, min_pts=M18, para_f=0.05, nv=True,
             call_seg_counts=None, device=False):
    """Calls the C ABI through pbnet_b200.cluster.Context with host (numpy) or device (torch.cuda) data."""
    import torch
    xs = np.ascontiguousarray(xyz_shift, dtype=np.float32).reshape(-1, 3)
    xo = np.ascontiguousarray(xyz_orig, dtype=np.float32).reshape(-1, 3)
    cols = [np.ascontiguousarray(xs[:, i]) for i in range(3)] + [np.ascontiguousarray(xo[:, i]) for i in range(3)]
    sem32 = np.ascontiguousarray(sem, dtype=np.int32)
    if device:
        dev = torch.device("cuda", ctx.device)
        cols = [torch.from_numpy(c).to(dev) for c in cols]
        sem32 = torch.from_numpy(sem32).to(dev)
    out = ctx.binary_cluster(*cols, sem32, np.asarray(seg_counts, np.int32), radius, min_pts, para_f, nv,
                             call_seg_counts=call_seg_counts)
    res = {}
    for k in ("cluster_id", "cluster_num", "degree", "center", "clt_sem"):
        v = out[k]
        res[k] = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    res["den_queue"] = res["degree"]
    res["n_clusters"] = out["n_clusters"]
    res["call_clusters"] = out["call_clusters"]
    return res


def diff_report(got, want, n_seg=None):
    """Returns a list of human-readable mismatches (empty = identical)."""
    bad = []
    for k in ("den_queue", "cluster_num", "cluster_id", "clt_sem"):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape:
            bad.append(f"{k}: shape {a.shape} != {b.shape}")
        elif not np.array_equal(a, b):
            idx = np.nonzero(a != b)[0]
            bad.append(f"{k}: {len(idx)} of {a.size} differ, first {idx[:5].tolist()} got {a[idx[:5]].tolist()} "
                       f"want {b[idx[:5]].tolist()}")
    a, b = np.asarray(got["center"], np.float32), np.asarray(want["center"], np.float32)
    if a.shape != b.shape:
        bad.append(f"center: shape {a.shape} != {b.shape}")
    elif not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
        bad.append(f"center: max abs diff {np.abs(a - b).max():.3e} (bit-exact replay expected)")
    return bad
