"""Generates tests/golden/*.npz by running the UNMODIFIED compiled reference (oracle/_ref/PB_lib,
built by oracle/build_ref.py from /root/reference) on a B200 and cross-checks the CPU oracle against
it.  Run on the GPU box:

    gpurun -- python tests/golden/make_golden.py            # writes gpurun_out/golden/*.npz + report.json

then copy gpurun_out/golden/*.npz into tests/golden/.  The argument marshalling below restates
lib/PB_lib/torch_io/pbnet_ops.py:14-75 (what the reference wrapper feeds PB_lib.binary_cluster).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import PB_lib  # noqa: E402  (the compiled reference)

from oracle import pb_oracle as po  # noqa: E402
from pbnet_b200 import scenes  # noqa: E402


def ref_binary_cluster(xyz_shift, xyz_orig, sem, seg_counts, radius18, min_pts18, para_f=0.05, nv_flag=True):
    """Feeds the reference extension exactly like pbnet_ops.Cluster.forward does."""
    t = torch.from_numpy
    xs = np.ascontiguousarray(xyz_shift, dtype=np.float32)
    xo = np.ascontiguousarray(xyz_orig, dtype=np.float32)
    x, y, z = (t(xs[:, i].copy()) for i in range(3))
    l1 = torch.abs(x) + torch.abs(y) + torch.abs(z)
    imap = torch.cat([torch.arange(0, int(c)) for c in seg_counts]).type(torch.int32).contiguous()
    ox, oy, oz = (t(xo[:, i].copy()) for i in range(3))
    n = xs.shape[0]
    semt = t(np.ascontiguousarray(sem, dtype=np.int32))
    bp = t(np.ascontiguousarray(seg_counts, dtype=np.int32))
    cid = (torch.ones(n) * -1).type(torch.int32).contiguous()
    cnum = torch.zeros([len(seg_counts)]).type(torch.int32)
    den = torch.zeros(n, dtype=torch.int32)
    center = torch.zeros(n, dtype=torch.float32)
    csem = torch.zeros(n, dtype=torch.int32)
    PB_lib.binary_cluster(x, y, z, l1, imap, ox, oy, oz, semt, bp, t(np.asarray(radius18, np.float32)),
                          t(np.asarray(min_pts18, np.int32)), cid, cnum, den, center, csem,
                          len(seg_counts), float(para_f), bool(nv_flag))
    return dict(cluster_id=cid.numpy().copy(), cluster_num=cnum.numpy().copy(), den_queue=den.numpy().copy(),
                center=center.numpy().copy(), clt_sem=csem.numpy().copy())


def same(a, b):
    bad = [k for k in ("cluster_id", "cluster_num", "den_queue", "clt_sem") if not np.array_equal(a[k], b[k])]
    if a["center"].shape != b["center"].shape or not np.array_equal(a["center"].view(np.uint32),
                                                                      b["center"].view(np.uint32)):
        bad.append("center")
    return bad


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    r18 = np.full(18, np.float32(0.04), np.float32)
    m18 = np.full(18, 31, np.int32)
    report = {"cases": [], "mismatches": []}

    def run_case(name, xs, xo, sem, seg, r=r18, m=m18, nv=True, save=False):
        t0 = time.perf_counter()
        ref = ref_binary_cluster(xs, xo, sem, seg, r, m, 0.05, nv)
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        ora = po.oracle_binary_cluster(xs, xo, sem, seg, r, m, 0.05, nv)
        t_ora = time.perf_counter() - t0
        bad = same(ref, ora)
        report["cases"].append(dict(name=name, n=int(len(sem)), segs=[int(s) for s in seg], K=int(ref["cluster_num"].sum()),
                                    t_ref=t_ref, t_oracle=t_ora, bad=bad))
        if bad:
            report["mismatches"].append(name)
        if save:
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz_shift=np.asarray(xs, np.float32),
                                xyz_orig=np.asarray(xo, np.float32), sem=np.asarray(sem, np.int32),
                                seg_counts=np.asarray(seg, np.int32), radius=r, min_pts=m, nv_flag=np.int32(nv),
                                **{"ref_" + k: v for k, v in ref.items()})
        return ref

    # ---- committed fixtures (small) ----------------------------------------------------------------
    for seed, npts, copies in ((1001, 12000, 1), (1002, 12000, 3), (1003, 30000, 1)):
        sc = scenes.make_scene(seed, npts)
        calls = scenes.class_calls(sc, copies)
        # one fixture per scene: all per-class calls concatenated as separate cases keeps files few
        for c in calls:
            run_case(f"s{seed}_c{c['sem_id']:02d}_b{copies}", c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"],
                     save=True)
    # mixed-class single segment, non-uniform min_pts, and nv_flag=False
    sc = scenes.make_scene(1004, 16000)
    fg = sc["sem"] >= 2
    xs = (sc["xyz_orig"] + sc["offset"])[fg]
    xo = sc["xyz_orig"][fg]
    se = sc["sem"][fg]
    m2 = m18.copy()
    m2[3] = 5
    m2[7] = 100
    half = len(se) // 2
    run_case("mixed_1seg", xs, xo, se, [len(se)], save=True)
    run_case("mixed_2seg_minpts", xs, xo, se, [half, len(se) - half], m=m2, save=True)
    run_case("mixed_novote", xs, xo, se, [0, len(se), 0], nv=False, save=True)
    # tiny / degenerate
    rng = np.random.Generator(np.random.PCG64(7))
    p = rng.normal(0, 0.01, size=(700, 3)).astype(np.float32)
    run_case("blob700_c10", p, p + np.float32(1.0), np.full(700, 10), [700], save=True)
    run_case("blob700_c10_empty_segs", p, p + np.float32(1.0), np.full(700, 10), [0, 300, 0, 400], save=True)
    q = np.repeat(p[:40], 5, axis=0)  # exact duplicates
    run_case("dups200_c17", q, q, np.full(200, 17), [200], save=True)
    run_case("single_point", p[:1], p[:1], np.full(1, 5), [1], save=True)

    # ---- larger sweep, not stored: full 150k scenes, per-class loop ------------------------------------
    for seed in (22, 23, 24):
        sc = scenes.make_scene(seed, 150000)
        for copies in (1, 3):
            for c in scenes.class_calls(sc, copies):
                run_case(f"big{seed}_c{c['sem_id']:02d}_b{copies}", c["xyz_shift"], c["xyz_orig"], c["sem"],
                         c["seg_counts"])
    tot_ref = sum(c["t_ref"] for c in report["cases"] if c["name"].startswith("big22") and c["name"].endswith("b1"))
    npts = sum(c["n"] for c in report["cases"] if c["name"].startswith("big22") and c["name"].endswith("b1"))
    report["ref_scene22_points_per_s"] = npts / tot_ref
    report["n_cases"] = len(report["cases"])
    with open(os.path.join(out_dir, "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("cases", len(report["cases"]), "mismatches", report["mismatches"])
    print("reference (B200, host-driven) scene22 points/s:", report["ref_scene22_points_per_s"])


if __name__ == "__main__":
    main()
