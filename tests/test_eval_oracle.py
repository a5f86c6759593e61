"""CPU: the evaluation post-processing restatement (oracle/eval_oracle.py) against golden vectors produced by the
reference's own source lines (tests/golden/make_golden_eval.py: eval_map.py:63-121, tools/mIOU.py:77-87,
tools/getins.py:72-98)."""
import glob
import os

import numpy as np
import pytest

from oracle import eval_oracle as eo

HERE = os.path.dirname(os.path.abspath(__file__))
EVAL = sorted(glob.glob(os.path.join(HERE, "golden", "eval", "*.npz")))


def pseudo(n, mult, mod=None):
    h = (np.arange(n, dtype=np.uint64) * np.uint64(mult)) % np.uint64(1 << 32)
    return h if mod is None else (h % np.uint64(mod))


def eval_inputs(g):
    """Same deterministic inputs as tests/golden/make_golden_eval.py::make_inputs."""
    d = np.load(os.path.join(HERE, "golden", "scenes", str(g["scene"])))
    pidx = d["ref_prop_idx"].astype(np.int64)
    poff = d["ref_prop_offset"].astype(np.int64)
    N = int(d["n_all"])
    n3 = N // 3
    P = len(poff) - 1
    score = (pseudo(P, 2246822519).astype(np.float64) / 4294967296.0 * float(g["score_scale"])).astype(np.float32)
    pred_sem = pseudo(N, 3266489917, 20).astype(np.int64)
    first = np.full(n3, P + 3, np.float64)
    np.minimum.at(first, pidx[:, 1] % n3, pidx[:, 0].astype(np.float64))
    jitter = pseudo(n3, 668265263).astype(np.float64) / 4294967296.0 * 1.5
    order = np.argsort(first + jitter, kind="stable")
    sp = np.empty(n3, np.int64)
    sp[order] = np.arange(n3) // 40
    return dict(proposals_idx=pidx, proposals_offset=poff, clt_score=score, pred_sem=pred_sem, superpoint=sp, point_num=N)


def test_fixtures_present():
    assert len(EVAL) >= 4


@pytest.mark.parametrize("path", EVAL, ids=[os.path.basename(p)[:-4] for p in EVAL])
def test_postprocess_oracle_equals_reference_lines(path):
    g = np.load(path)
    inp = eval_inputs(g)
    out = eo.postprocess(inp["proposals_idx"], inp["proposals_offset"], inp["clt_score"], inp["pred_sem"], inp["superpoint"],
                         inp["point_num"], float(g["nms"]), float(g["score_thr"]), int(g["npoint"]))
    assert len(out["scores"]) == int(g["ref_n"])
    assert np.array_equal(out["label"], g["ref_label"])
    assert np.array_equal(out["scores"].view(np.uint32), g["ref_scores"].view(np.uint32))
    assert np.array_equal(out["sem"], g["ref_sem"])
