"""CPU: the local-scene / get_proposal restatement (oracle/scene_oracle.py) against golden vectors produced by the
reference's own source lines (tests/golden/make_golden_scenes.py, network/PBNet.py:180-234, 317-346)."""
import glob
import os

import numpy as np
import pytest

from oracle import scene_oracle as so

SCENES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenes", "*.npz")))


def pseudo_scores(n):
    """Same stand-in mask scores as the generator (tests/golden/make_golden_scenes.py::pseudo_scores)."""
    return ((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(1 << 32)).astype(np.float64).astype(
        np.float32) / np.float32(4294967296.0)


def golden_scores(d):
    lens = d["ref_lens"]
    ms = pseudo_scores(int(lens.sum()))
    if len(lens) > 2:
        o = int(lens[:1].sum())
        ms[o:o + int(lens[1])] = 0.0
    return ms


def call_seg_counts(d):
    return np.full(len(d["call_sem"]), int(d["copies"]), np.int32)


def test_fixtures_present():
    assert len(SCENES) >= 5


@pytest.mark.parametrize("path", SCENES, ids=[os.path.basename(p)[:-4] for p in SCENES])
def test_local_scenes_oracle_equals_reference_lines(path):
    d = np.load(path)
    train = str(d["task"]) != "test"
    out = so.local_scenes(d["cluster_id"], d["cluster_num"], d["center"], d["seg_counts"], call_seg_counts(d), d["call_sem"],
                          ins_label=d["ins_label"] if train else None, k_max=float(d["k_max"]))
    assert np.array_equal(out["lens"], d["ref_lens"])
    assert np.array_equal(d["ins_ind"][out["pos"]], d["ref_idx"])
    assert np.array_equal(out["dpn"].view(np.uint32), d["ref_dpn"].view(np.uint32))
    if train:
        assert np.array_equal(out["gt"], d["ref_gt"].astype(np.int32))
        assert len(out["lens"]) < int(d["cluster_num"].sum())  # the fixtures do exercise the -100 skip


@pytest.mark.parametrize("path", SCENES, ids=[os.path.basename(p)[:-4] for p in SCENES])
def test_get_proposal_oracle_equals_reference_lines(path):
    d = np.load(path)
    ms = golden_scores(d)
    pidx, off, ids, pms = so.get_proposal(d["ref_lens"], d["ref_idx"], ms)
    assert np.array_equal(pidx, d["ref_prop_idx"].astype(np.int64))
    assert np.array_equal(off, d["ref_prop_offset"])
    assert np.array_equal(ids, d["ref_prop_ids"].astype(np.int64))
    assert np.array_equal(pms, ms[ms > np.float32(0.45)])
