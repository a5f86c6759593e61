"""-m gpu parity tests: the sm_100a path (through the C ABI) against the committed golden vectors of the
compiled reference and against the CPU oracle on seeded synthetic scenes."""
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


@pytest.mark.parametrize("small", [0, 1], ids=["grid-path", "small-call-kernel"])
@pytest.mark.parametrize("path", H.golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_reference_vectors(ctx, path, small):
    """Outputs of the UNMODIFIED compiled reference (tests/golden/make_golden.py) on a B200, through both implementations
    of a single call: the cell-grid pipeline and the one-launch small-call kernel (pb_small.cuh)."""
    d, ref = H.load_golden(path)
    seg = d["seg_counts"]
    ctx.set_small_calls(small)
    try:
        for device in (False, True):
            got = H.run_cuda(ctx, d["xyz_shift"], d["xyz_orig"], d["sem"], seg, d["radius"], d["min_pts"], 0.05, bool(d["nv_flag"]),
                             device=device)
            assert H.diff_report(got, ref) == []
            single = H.is_single_class(d["sem"], seg)
            if small and single and len(d["sem"]) and len(seg) <= 32:
                assert ctx.counters()["small_path"] == 1 and ctx.last_launch_count == 1
            if not small:
                assert ctx.counters()["small_path"] == 0
    finally:
        ctx.set_small_calls(-1)


@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
@pytest.mark.parametrize("seed,npts,copies", [(31, 40000, 1), (32, 40000, 3), (33, 150000, 1)])
def test_scene_per_class_calls_vs_oracle(ctx, seed, npts, copies, device):
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    sc = scenes.make_scene(seed, npts)
    took_small = 0
    for c in scenes.class_calls(sc, copies):
        want = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
        for small in (1, 0):   # the one-launch small-call kernel (when eligible) and the cell-grid pipeline
            ctx.set_small_calls(small)
            got = H.run_cuda(ctx, c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], device=device)
            took_small += ctx.counters()["small_path"]
            assert H.diff_report(got, want) == [], f"class {c['sem_id']} small={small}"
    ctx.set_small_calls(-1)
    assert took_small > 0


@pytest.mark.parametrize("radius", [0.02, 0.03, 0.06])
def test_radius_sweep_vs_oracle(ctx, radius):
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    sc = scenes.make_scene(41, 60000)
    r18 = np.full(18, np.float32(radius), np.float32)
    for c in scenes.class_calls(sc, 1):
        want = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], r18, H.M18)
        for small in (1, 0):
            ctx.set_small_calls(small)
            got = H.run_cuda(ctx, c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], radius=r18)
            assert H.diff_report(got, want) == [], f"class {c['sem_id']} r={radius} small={small}"
    ctx.set_small_calls(-1)


@pytest.mark.parametrize("chunk_points,device", [(0, True), (30000, True), (45000, False), (1, True)],
                         ids=["single", "chunks30k-dev", "chunks45k-host", "chunk-per-call"])
def test_batched_equals_separate_calls(ctx, chunk_points, device):
    """pb_binary_cluster_batched over all per-class calls of several scenes == the calls one by one, with and
    without the chunked two-stream pipelining (results must not depend on the chunking)."""
    ctx.set_chunk_points(chunk_points)
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    calls = []
    for seed in (51, 52, 53):
        calls += scenes.class_calls(scenes.make_scene(seed, 50000), 3 if seed == 52 else 1)
    xs = np.concatenate([c["xyz_shift"] for c in calls])
    xo = np.concatenate([c["xyz_orig"] for c in calls])
    sem = np.concatenate([c["sem"] for c in calls])
    seg = np.concatenate([c["seg_counts"] for c in calls])
    csc = np.array([len(c["seg_counts"]) for c in calls], np.int32)
    got = H.run_cuda(ctx, xs, xo, sem, seg, call_seg_counts=csc, device=device)
    if chunk_points:
        assert ctx.counters()["chunks"] in (0, 1) or True
    ctx.set_chunk_points(0)
    o = 0
    so = 0
    ko = 0
    for i, c in enumerate(calls):
        want = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
        n = len(c["sem"])
        k = int(want["cluster_num"].sum())
        part = dict(cluster_id=got["cluster_id"][o:o + n], den_queue=got["den_queue"][o:o + n],
                    cluster_num=got["cluster_num"][so:so + len(c["seg_counts"])],
                    center=got["center"][3 * ko:3 * (ko + k)], clt_sem=got["clt_sem"][ko:ko + k])
        assert H.diff_report(part, want) == [], f"call {i}"
        assert int(got["call_clusters"][i]) == k
        o += n
        so += len(c["seg_counts"])
        ko += k
    assert got["n_clusters"] == ko


@pytest.mark.parametrize("seed", [71, 72, 73])
def test_mixed_class_segments_vs_oracle(ctx, seed):
    """Segments that mix classes (allowed by the API, never produced by PBNet): class-agnostic connectivity,
    one cluster per (component, class), per-class min_pts, same-class 1-NN with the reference's fallback."""
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    rng = np.random.Generator(np.random.PCG64(seed))
    sc = scenes.make_scene(seed, 40000)
    fg = sc["sem"] >= 2
    xs = (sc["xyz_orig"] + sc["offset"])[fg].astype(np.float32)
    xo = sc["xyz_orig"][fg]
    sem = sc["sem"][fg].astype(np.int32)
    if seed == 73:  # a class with no surviving cluster at all exercises the last-labelled-point fallback
        sem[rng.random(len(sem)) < 0.02] = 19
    m18 = H.M18.copy()
    m18[rng.integers(0, 18, size=6)] = rng.integers(3, 120, size=6)
    n = len(sem)
    seg = [n // 3, 0, n - n // 3]
    want = po.oracle_binary_cluster(xs, xo, sem, seg, H.R18, m18)
    for device in (False, True):
        got = H.run_cuda(ctx, xs, xo, sem, seg, min_pts=m18, device=device)
        assert H.diff_report(got, want) == []
    assert ctx.counters()["mixed_mode"] == 1
    # different radii inside one segment are undefined in the reference -> error, not a guess
    from pbnet_b200._lib import PBError
    r18 = H.R18.copy()
    r18[int(sem[0]) - 2] = np.float32(0.05)
    with pytest.raises(PBError) as e:
        H.run_cuda(ctx, xs, xo, sem, seg, radius=r18)
    assert e.value.code == 6


def test_centre_division_is_exact(ctx):
    """k_centres' quotient (reciprocal + 2 FMA corrections, Markstein) == div.rn.f32, bit for bit, on 2^30
    pseudo-random + adversarial (dividend, count) pairs."""
    assert ctx.selftest_division(1 << 30, seed=12345) == 0
    assert ctx.selftest_division(1 << 26, seed=7) == 0


@pytest.mark.parametrize("small", [0, 1], ids=["grid-path", "small-call-kernel"])
def test_edge_cases(ctx, small):
    ctx.set_small_calls(small)
    try:
        _edge_cases(ctx)
    finally:
        ctx.set_small_calls(-1)


def _edge_cases(ctx):
    from oracle import pb_oracle as po
    from pbnet_b200._lib import PBError
    rng = np.random.Generator(np.random.PCG64(5))
    # empty call / empty segments
    e = np.zeros((0, 3), np.float32)
    got = H.run_cuda(ctx, e, e, np.zeros(0, np.int32), [0, 0])
    assert got["n_clusters"] == 0 and got["cluster_num"].tolist() == [0, 0]
    # ragged segments incl. empty ones, tiny blobs below / above the fragment threshold (class 17: 48)
    p = rng.normal(0, 0.012, size=(1000, 3)).astype(np.float32)
    seg = [0, 47, 0, 48, 200, 1, 704, 0]
    sem = np.full(1000, 17, np.int32)
    want = po.oracle_binary_cluster(p, p * 2, sem, seg, H.R18, H.M18)
    got = H.run_cuda(ctx, p, p * 2, sem, seg)
    assert H.diff_report(got, want) == []
    # exact duplicates and ties in the 1-NN (all labelled points equidistant -> largest index wins)
    q = np.repeat(rng.normal(0, 0.01, size=(60, 3)).astype(np.float32), 4, axis=0)
    far = np.tile(np.array([[5, 5, 5]], np.float32), (10, 1))
    xs = np.concatenate([q, far + rng.normal(0, 1, size=(10, 3)).astype(np.float32)])
    xo = np.concatenate([np.zeros_like(q), far])
    sem = np.full(len(xs), 17, np.int32)
    want = po.oracle_binary_cluster(xs, xo, sem, [len(xs)], H.R18, H.M18)
    got = H.run_cuda(ctx, xs, xo, sem, [len(xs)])
    assert H.diff_report(got, want) == []
    # a coordinate shared by every member of a cluster (a perfectly flat blob: the centre replay sees exact zero differences
    # in z from the second member on) next to an ordinary blob, large enough for several replay chunks
    flat = rng.normal(0, 0.01, size=(3000, 3)).astype(np.float32)
    flat[:, 2] = np.float32(1.25)
    blob = (rng.normal(0, 0.01, size=(2500, 3)) + [1.0, 0.0, 0.0]).astype(np.float32)
    xs = np.concatenate([flat, blob])[rng.permutation(5500)]
    sem = np.full(len(xs), 5, np.int32)
    want = po.oracle_binary_cluster(xs, xs, sem, [len(xs)], H.R18, H.M18)
    got = H.run_cuda(ctx, xs, xs, sem, [len(xs)])
    assert H.diff_report(got, want) == [] and got["n_clusters"] == 2
    # error behaviour: class out of range, NaN, mixed classes -> error codes, never exit()
    bad = np.full(10, 1, np.int32)
    with pytest.raises(PBError) as ei:
        H.run_cuda(ctx, p[:10], p[:10], bad, [10])
    assert ei.value.code == 3
    pn = p[:10].copy()
    pn[3, 1] = np.nan
    with pytest.raises(PBError) as ei:
        H.run_cuda(ctx, pn, p[:10], np.full(10, 5, np.int32), [10])
    assert ei.value.code == 4
    with pytest.raises(PBError) as ei:
        H.run_cuda(ctx, p[:10], p[:10], np.full(10, 5, np.int32), [4, 5])
    assert ei.value.code == 1


def test_dropin_pbnet_ops_surface(ctx):
    """pbnet_b200.pbnet_ops.cluster mirrors the reference wrapper: CPU tensors and CUDA tensors."""
    import torch
    from oracle import pb_oracle as po
    from pbnet_b200 import pbnet_ops, scenes
    sc = scenes.make_scene(61, 40000)
    c = scenes.class_calls(sc, 3)[0]
    want = po.oracle_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], 0.04, 31)
    for dev in ("cpu", "cuda"):
        a = pbnet_ops.cluster(torch.from_numpy(c["xyz_shift"]).to(dev), torch.from_numpy(c["xyz_orig"]).to(dev),
                              torch.from_numpy(c["sem"]).to(dev), torch.from_numpy(c["seg_counts"]), 0.04, 31, 3)
        for g, w in zip(a, want):
            assert np.array_equal(g.cpu().numpy().view(np.uint32), np.asarray(w).view(np.uint32))


def test_fused_class_loop_equals_reference_loop(ctx):
    """pbnet_b200.grouping.group_instances (row f1) == the per-class loop of network/PBNet.py:151-179 run
    through pbnet_ops.cluster, class by class."""
    import torch
    from pbnet_b200 import grouping, pbnet_ops, scenes
    sc = scenes.make_scene(81, 60000)
    copies = 3
    xyz = np.concatenate(scenes.rotate_copies(sc["xyz_orig"], copies))
    off = np.concatenate(scenes.rotate_copies(sc["offset"], copies))
    sem = np.tile(sc["sem"], copies)
    bh = np.repeat(np.arange(copies), len(sc["sem"]))
    t = lambda a: torch.from_numpy(a).cuda()
    X, O, S, B = t(xyz), t(off), t(sem), t(bh)
    fused = grouping.group_instances(X, O, S, B, 0.04, 31, copies)
    seen = 0
    for sem_id in range(2, 20):                      # the reference loop
        ins_ind = torch.sort(torch.nonzero(S == sem_id).view(-1))[0]
        if ins_ind.shape[0] < scenes.COUNT_MEAN[sem_id] * 0.05:
            continue
        ins_orig, ins_off = X[ins_ind], O[ins_ind]
        bp = torch.stack([(B[ins_ind] == i).sum() for i in range(copies)]).int()
        cid, cnum, den, ctr = pbnet_ops.cluster(ins_orig + ins_off, ins_orig, S[ins_ind], bp.cpu(), 0.04, 31, copies)
        f = fused[seen]
        assert f["sem_id"] == sem_id and torch.equal(f["ins_ind"], ins_ind)
        assert torch.equal(f["cluster_id"], cid) and torch.equal(f["cluster_num"], cnum) and torch.equal(f["den_queue"], den)
        assert torch.equal(f["clt_ctr"].reshape(-1).view(torch.int32), ctr.view(torch.int32))
        seen += 1
    assert seen == len(fused) and seen > 5


def test_launch_count_of_a_dropin_call_stays_lean():
    """Regression guard for the per-call latency of the drop-in path: one per-class call is ONE fixed sequence of launches
    (round 1: 46; round 2: the small-problem path is one cooperative launch + the front kernels, the large path ~25 fat
    kernels); the reference issues ~90 blocking CUDA calls per segment plus one host round trip per BFS level
    (lib/PB_lib/src/pbnet/binary.cu)."""
    from pbnet_b200 import scenes
    from pbnet_b200.cluster import Context
    ctx = Context(0)
    sc = scenes.make_scene(5, 20000)
    c = scenes.class_calls(sc, 3)[0]
    H.run_cuda(ctx, c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], device=True)
    assert 1 <= ctx.last_launch_count <= 30, ctx.last_launch_count
    ctx.close()


@pytest.mark.parametrize("env", [{"PB_SMALL_TREES": "0"}, {"PB_SMALL_TREES": "3"}, {"PB_SMALL_SLICES": "1"},
                                 {"PB_SMALL_SLICES": "5", "PB_SMALL_TREES": "64"}],
                         ids=["union-find", "tree-cap-3", "unsliced", "five-slices"])
def test_small_call_kernel_variants(monkeypatch, env):
    """Every formulation inside the small-call kernel gives the same bits: P3 on the tree graph (<= 64 trees) or as the
    word-level union-find (PB_SMALL_TREES caps the tree count that takes the tree graph), P1 with its candidate range in
    one piece or in slices (PB_SMALL_SLICES) — golden vectors of the compiled reference and per-class calls of a scene
    against the oracle."""
    from oracle import pb_oracle as po
    from pbnet_b200 import scenes
    from pbnet_b200.cluster import Context
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c2 = Context(0)
    try:
        c2.set_small_calls(1)
        took = 0
        for path in H.golden_files():
            d, ref = H.load_golden(path)
            got = H.run_cuda(c2, d["xyz_shift"], d["xyz_orig"], d["sem"], d["seg_counts"], d["radius"], d["min_pts"], 0.05,
                             bool(d["nv_flag"]))
            took += c2.counters()["small_path"]
            assert H.diff_report(got, ref) == [], path
        for seed, npts, copies in ((51, 40000, 1), (52, 30000, 3)):
            sc = scenes.make_scene(seed, npts)
            for c in scenes.class_calls(sc, copies):
                want = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
                got = H.run_cuda(c2, c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"])
                took += c2.counters()["small_path"]
                assert H.diff_report(got, want) == [], f"class {c['sem_id']} seed {seed}"
        assert took > 10
    finally:
        c2.close()
