"""GPU: evaluation post-processing (C ABI pb_eval_postprocess) against golden vectors produced by the reference's own
source lines (eval_map.py:63-121, tools/mIOU.py:77-87, tools/getins.py:72-98) and against the sparse CPU restatement."""
import os

import numpy as np
import pytest

from tests.test_eval_oracle import EVAL, eval_inputs

pytestmark = pytest.mark.gpu


def _run(inp, nms, sthr, npt):
    import torch
    from pbnet_b200 import evalpost
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = evalpost.postprocess(t(inp["proposals_idx"]), t(inp["proposals_offset"]), t(inp["clt_score"]), t(inp["pred_sem"]),
                               t(inp["superpoint"]), inp["point_num"], nms, sthr, npt)
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("path", EVAL, ids=[os.path.basename(p)[:-4] for p in EVAL])
def test_postprocess_equals_reference_lines(path):
    g = np.load(path)
    inp = eval_inputs(g)
    out = _run(inp, float(g["nms"]), float(g["score_thr"]), int(g["npoint"]))
    assert len(out["scores"]) == int(g["ref_n"])
    assert np.array_equal(out["label"], g["ref_label"])
    assert np.array_equal(out["scores"].view(np.uint32), g["ref_scores"].view(np.uint32))
    assert np.array_equal(out["sem"], g["ref_sem"])


def test_postprocess_threshold_sweep_matches_oracle():
    from oracle import eval_oracle as eo
    g = np.load(EVAL[0])
    inp = eval_inputs(g)
    for nms, sthr, npt in ((0.0, 0.0, 0), (0.3, 0.2, 50), (0.05, 0.5, 1000), (0.9, 0.99, 10), (0.1, 0.07, 10 ** 6)):
        out = _run(inp, nms, sthr, npt)
        want = eo.postprocess(inp["proposals_idx"], inp["proposals_offset"], inp["clt_score"], inp["pred_sem"], inp["superpoint"],
                              inp["point_num"], nms, sthr, npt)
        assert np.array_equal(out["label"], want["label"]), (nms, sthr, npt)
        assert np.array_equal(out["proposal"], want["picked"].astype(np.int32))
        assert np.array_equal(out["scores"], want["scores"]) and np.array_equal(out["sem"], want["sem"])


def test_dense_masks_are_the_reference_cluster_matrix():
    import torch
    from pbnet_b200 import evalpost
    g = np.load(EVAL[1])
    inp = eval_inputs(g)
    out = _run(inp, float(g["nms"]), float(g["score_thr"]), int(g["npoint"]))
    m = evalpost.dense_masks(torch.from_numpy(out["label"]).cuda(), len(out["scores"])).cpu().numpy()
    assert m.shape == (int(g["ref_n"]), inp["point_num"] // 3) and m.sum(0).max() <= 1
    for c in range(m.shape[0]):
        assert np.array_equal(m[c] == 1, g["ref_label"] == c)


def test_postprocess_rejects_bad_superpoints():
    import torch
    from pbnet_b200 import evalpost
    from pbnet_b200._lib import PBError
    g = np.load(EVAL[0])
    inp = eval_inputs(g)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with pytest.raises(PBError):
        evalpost.postprocess(t(inp["proposals_idx"]), t(inp["proposals_offset"]), t(inp["clt_score"]), t(inp["pred_sem"]),
                             t(inp["superpoint"]), inp["point_num"], n_superpoints=5)
