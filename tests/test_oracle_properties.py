"""CPU: size-independent properties of the grouping semantics (SURVEY.md Appendix A invariants), checked on
the oracle; the GPU suite checks the same properties on the CUDA path at full size."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import pb_oracle as po
from pbnet_b200 import scenes
from tests import helpers as H


def check_invariants(out, sem, seg):
    den, cid, cnum = out["den_queue"], out["cluster_id"], out["cluster_num"]
    assert int(den.astype(np.int64).sum()) % 2 == 0          # neighbour relation is symmetric
    K = int(cnum.sum())
    assert len(out["center"]) == 3 * K and len(out["clt_sem"]) == K
    o = 0
    base = 0
    for b, n in enumerate(seg):
        ids = cid[o:o + n]
        k = int(cnum[b])
        if n:
            if k == 0:
                assert (ids == -1).all()
            else:
                assert ids.min() >= base and ids.max() < base + k      # no point left unlabelled
                assert len(np.unique(ids)) == k                        # every id non-empty, contiguous
        base += k
        o += n


@pytest.mark.parametrize("seed", [3, 4])
def test_invariants_on_scene(seed):
    sc = scenes.make_scene(seed, 25000)
    for c in scenes.class_calls(sc, 3):
        out = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
        check_invariants(out, c["sem"], c["seg_counts"])


def test_segment_permutation_invariance():
    """Segments are independent: swapping two segments permutes the per-segment results (ids shift by the
    cluster counts of the segments before)."""
    sc = scenes.make_scene(9, 25000)
    c = scenes.class_calls(sc, 3)[0]
    n = int(c["seg_counts"][0])
    a = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"], c["seg_counts"], H.R18, H.M18)
    perm = np.concatenate([np.arange(n, 2 * n), np.arange(0, n), np.arange(2 * n, 3 * n)])
    b = po.oracle_binary_cluster(c["xyz_shift"][perm], c["xyz_orig"][perm], c["sem"][perm], c["seg_counts"], H.R18, H.M18)
    assert np.array_equal(a["den_queue"][perm], b["den_queue"])
    assert a["cluster_num"][[1, 0, 2]].tolist() == b["cluster_num"].tolist()
    k0, k1 = int(a["cluster_num"][0]), int(a["cluster_num"][1])
    ida = a["cluster_id"]
    exp = np.concatenate([ida[n:2 * n] - k0, ida[:n] + k1, ida[2 * n:]])
    assert np.array_equal(exp, b["cluster_id"])


def test_l1norm_and_index_mapper_are_irrelevant():
    """The reference's l1_norm / index_mapper inputs only steer its slab pruning (binary.cu:49-69): a rigid
    translation changes every l1 norm but not the neighbour relation (as long as fp32 differences are exact)."""
    rng = np.random.Generator(np.random.PCG64(1))
    p = (rng.integers(-400, 400, size=(3000, 3)) * np.float32(1 / 256)).astype(np.float32)  # exactly representable
    sem = np.full(3000, 12, np.int32)
    a = po.oracle_binary_cluster(p, p, sem, [3000], H.R18, H.M18)
    t = np.array([4.0, -2.0, 1.0], np.float32)
    b = po.oracle_binary_cluster(p + t, p + t, sem, [3000], H.R18, H.M18)
    assert np.array_equal(a["den_queue"], b["den_queue"]) and np.array_equal(a["cluster_id"], b["cluster_id"])


def test_fragment_threshold_is_fp32():
    """Class 16 (mean 2120): 2120*0.05f == 106.0 exactly in fp32 — a size-106 cluster survives
    (computing the threshold in double would wrongly drop it)."""
    rng = np.random.Generator(np.random.PCG64(2))
    for size, keep in ((105, 0), (106, 1)):
        p = rng.normal(0, 0.005, size=(size, 3)).astype(np.float32)
        out = po.oracle_binary_cluster(p, p, np.full(size, 16, np.int32), [size], H.R18, H.M18)
        assert int(out["cluster_num"][0]) == keep


def test_argument_errors():
    p = np.zeros((4, 3), np.float32)
    with pytest.raises(ValueError):
        po.oracle_binary_cluster(p, p, np.array([1, 2, 3, 4]), [4], H.R18, H.M18)      # class 1 out of range
    r = H.R18.copy()
    r[3] = 0.05
    with pytest.raises(ValueError):
        po.oracle_binary_cluster(p, p, np.array([2, 5, 2, 2]), [4], r, H.M18)          # non-uniform radius in a segment


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 400), st.sampled_from([0.02, 0.04, 0.08]), st.integers(1, 40))
def test_literal_equals_grid_random(seed, n, radius, min_pts):
    rng = np.random.Generator(np.random.PCG64(seed))
    k = int(rng.integers(1, 4))
    centres = rng.uniform(-0.3, 0.3, size=(k, 3))
    p = (centres[rng.integers(0, k, size=n)] + rng.normal(0, 0.02, size=(n, 3))).astype(np.float32)
    po_ = (p + rng.normal(0, 0.1, size=(n, 3))).astype(np.float32)
    sem = np.full(n, 17, np.int32)
    seg = [n // 3, n - n // 3]
    r18 = np.full(18, np.float32(radius), np.float32)
    m18 = np.full(18, min_pts, np.int32)
    a = po.oracle_binary_cluster(p, po_, sem, seg, r18, m18, mode="grid")
    b = po.oracle_binary_cluster(p, po_, sem, seg, r18, m18, mode="literal")
    assert H.diff_report(a, b) == []
    check_invariants(a, sem, seg)


def test_normals_oracle_geometry():
    """oracle/pb_oracle.c::pb_oracle_normals (restatement of lib/PB_lib/src/normal/cal_normal.cu): a planar fan has the plane's
    normal at every vertex, unreferenced vertices get (0,0,1), flipping the winding flips the normal, and only the first
    num_face faces take part."""
    import numpy as np
    from oracle import pb_oracle as po
    xyz = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.3, 0.2, 5.0]], np.float32)
    face = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    n = po.oracle_normals(xyz, face)
    assert np.array_equal(n, np.tile(np.array([0, 0, 1], np.float32), (5, 1)))
    n = po.oracle_normals(xyz, face[:, ::-1].copy())
    assert np.array_equal(n[:4], np.tile(np.array([0, 0, -1], np.float32), (4, 1))) and np.array_equal(n[4], [0, 0, 1])
    tilted = np.array([[0, 0, 0], [1, 0, 1], [0, 1, 0]], np.float32)          # plane z = x, normal (-1,0,1)/sqrt2
    n = po.oracle_normals(tilted, np.array([[0, 1, 2]], np.int32))
    assert np.allclose(n, np.array([-1, 0, 1]) / np.sqrt(2), atol=1e-7)
    n1 = po.oracle_normals(xyz, face, num_face=1)                             # vertex 3 is only in face 1 -> default
    assert np.array_equal(n1[3], [0, 0, 1]) and np.array_equal(n1[1], [0, 0, 1])
