"""-m gpu: BASELINE.json configs 3 and 4 as parity cases against the CPU oracle (the compiled reference
cannot run them: int32 neighbour totals / O(n^2) stages).
C3: 1 M-point synthetic room scene, radius sweep.  C4: dense single-class segment, HP-fraction sweep with
~2 000 neighbours per blob point."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pbnet_b200.cluster import Context
    c = Context(0)
    yield c
    c.close()


def oracle_calls(calls, r18, m18, threads=8):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pb_oracle as po
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda c: po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"],
                                                              r18, m18), calls))


@pytest.mark.parametrize("radius,min_pts", [(0.02, 6), (0.04, 20), (0.06, 31)])
def test_c3_large_scene_radius_sweep(ctx, radius, min_pts):
    from pbnet_b200 import scenes
    sc = scenes.make_scene(3003, 1_000_000, hp_frac=0.25)
    calls = scenes.class_calls(sc, 1)
    r18 = np.full(18, np.float32(radius), np.float32)
    m18 = np.full(18, min_pts, np.int32)
    want = oracle_calls(calls, r18, m18)
    xs = np.concatenate([c["xyz_shift"] for c in calls])
    xo = np.concatenate([c["xyz_orig"] for c in calls])
    sem = np.concatenate([c["sem"] for c in calls])
    seg = np.concatenate([c["seg_counts"] for c in calls])
    got = H.run_cuda(ctx, xs, xo, sem, seg, radius=r18, min_pts=m18, call_seg_counts=np.ones(len(calls), np.int32),
                     device=True)
    o = ko = 0
    for i, (c, w) in enumerate(zip(calls, want)):
        n, k = len(c["sem"]), int(w["cluster_num"].sum())
        part = dict(cluster_id=got["cluster_id"][o:o + n], den_queue=got["den_queue"][o:o + n],
                    cluster_num=got["cluster_num"][i:i + 1], center=got["center"][3 * ko:3 * (ko + k)],
                    clt_sem=got["clt_sem"][ko:ko + k])
        assert H.diff_report(part, w) == [], f"class {c['sem_id']} r={radius}"
        o += n
        ko += k
    assert ko > 0  # the sweep must actually produce clusters


@pytest.mark.parametrize("hp_fraction", [0.1, 0.5, 0.9])
def test_c4_dense_pathological(ctx, hp_fraction):
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    xs, xo, sem = scenes.make_dense_case(4004, 400_000, hp_fraction)
    want = po.oracle_binary_cluster(xs, xo, sem, [len(sem)], H.R18, H.M18, return_stats=True)
    got = H.run_cuda(ctx, xs, xo, sem, [len(sem)], device=True)
    assert H.diff_report(got, want) == []
    assert int(want["den_queue"].max()) >= 1500          # ~2k neighbours per blob point
    assert int(want["cluster_num"][0]) >= 10             # blobs may merge as the room shrinks with the background


def test_full_size_properties(ctx):
    """Size-independent properties at a BASELINE-sized batch (no oracle): symmetric degrees, contiguous ids,
    no unlabelled point in a segment that has clusters, idempotence of a second run."""
    from pbnet_b200 import scenes, workload
    sizes = scenes.scene_sizes(24)
    w = workload.build(range(24), sizes, 1, workers=1, cache_dir=None)  # no fork() once CUDA is initialised
    args = (np.stack([w["x"], w["y"], w["z"]], 1), np.stack([w["xo"], w["yo"], w["zo"]], 1), w["sem"], w["seg_counts"])
    a = H.run_cuda(ctx, *args, call_seg_counts=w["call_seg_counts"], device=True)
    b = H.run_cuda(ctx, *args, call_seg_counts=w["call_seg_counts"], device=True)
    assert H.diff_report(a, b) == []
    assert int(a["den_queue"].astype(np.int64).sum()) % 2 == 0
    o = ko = 0
    for s, n in enumerate(w["seg_counts"]):
        ids, k = a["cluster_id"][o:o + n], int(a["cluster_num"][s])
        if k == 0:
            assert (ids == -1).all()
        else:
            assert ids.min() == 0 and ids.max() == k - 1 and len(np.unique(ids)) == k   # one call = one segment here
        o += n
        ko += k
    assert ko == a["n_clusters"] and len(a["center"]) == 3 * ko


@pytest.mark.gpu
@pytest.mark.parametrize("n_scene,copies", [(20_000, 1), (140_000, 3), (250_000, 3)])
def test_symmetric_degree_kernel_equals_one_sided(n_scene, copies):
    """The symmetric neighbour count (every unordered pair tested once, the candidate's side added with REDUX + RED) and the
    one-sided kernel of round 1 (PB_DEG_SYM=0) give bit-identical outputs: window splitting (small), heavy-first (mid-size)
    and plain grids, with the small-call kernel switched off so that every size takes the cell-grid pipeline."""
    import os

    from pbnet_b200 import scenes
    from pbnet_b200.cluster import Context
    sc = scenes.make_scene(101 + n_scene % 97, n_scene)
    calls = scenes.class_calls(sc, copies)
    xs = np.concatenate([c["xyz_shift"] for c in calls])
    xo = np.concatenate([c["xyz_orig"] for c in calls])
    sem = np.concatenate([c["sem"] for c in calls])
    seg = np.concatenate([c["seg_counts"] for c in calls])
    csc = np.array([len(c["seg_counts"]) for c in calls], np.int32)
    res = {}
    old = {k: os.environ.get(k) for k in ("PB_DEG_SYM", "PB_SMALL", "PB_DEG_TMA", "PB_DEG_SLICES")}
    try:
        os.environ["PB_SMALL"] = "0"
        os.environ["PB_DEG_TMA"] = "0"   # candidates by direct loads
        for sym in ("1", "0"):
            os.environ["PB_DEG_SYM"] = sym
            c = Context(0)
            res[sym] = H.run_cuda(c, xs, xo, sem, seg, call_seg_counts=csc, device=True)
            c.close()
        # the TMA-staged candidate stream (the default), with the problem-size dependent window splitting and with one warp
        # per window (PB_DEG_SLICES=1: the form large problems take)
        os.environ["PB_DEG_SYM"], os.environ["PB_DEG_TMA"] = "1", "1"
        for key, slices in (("tma", None), ("tma1", "1")):
            if slices:
                os.environ["PB_DEG_SLICES"] = slices
            c = Context(0)
            res[key] = H.run_cuda(c, xs, xo, sem, seg, call_seg_counts=csc, device=True)
            c.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert H.diff_report(res["1"], res["0"]) == []
    assert H.diff_report(res["tma"], res["0"]) == []
    assert H.diff_report(res["tma1"], res["0"]) == []
    assert int(res["1"]["degree"].sum()) > 0
