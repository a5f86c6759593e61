"""Generates tests/golden/eval/*.npz: golden vectors for the evaluation post-processing of the proposals
(eval_map.py:63-123, tools/mIOU.py:77-87 non_max_suppression, tools/getins.py:72-98
align_superpoint_label).

Like make_golden_scenes.py this EXECUTES THE REFERENCE'S OWN SOURCE LINES on CPU tensors: the block of
eval_map.py between ``semantic_id = torch.tensor(semantic_label_idx`` and ``# ####full time`` and the two helper
functions are cut out of the reference files by text and run unmodified (``Tensor.cuda`` / ``torch.cuda.current_device``
are patched: there is no GPU in the build container).  Inputs: the proposal lists of the local-scene fixtures
(tests/golden/scenes, three rotated scene copies => the ``% (point_num/3)`` fold matters), pseudo cluster scores,
pseudo classes and a synthetic superpoint partition that mixes neighbouring proposals.

    python tests/golden/make_golden_eval.py        # run in the build container (needs /root/reference)
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _lines(path):
    return open(os.path.join(REF, path), encoding="utf-8").read().split("\n")


def _func(lines, name):
    a = next(i for i, l in enumerate(lines) if l.startswith("def " + name))
    b = next((i for i, l in enumerate(lines) if i > a and l.startswith("def ")), len(lines))
    return "\n".join(lines[a:b])


def load_reference_block():
    ev = _lines("eval_map.py")
    a = next(i for i, l in enumerate(ev) if "semantic_id = torch.tensor(semantic_label_idx" in l)
    b = next(i for i, l in enumerate(ev) if "# ####full time" in l)
    body = textwrap.dedent("\n".join(ev[a:b]))
    src = ("import numpy as np\nfrom scipy.sparse import coo_matrix\n"
           + _func(_lines("tools/mIOU.py"), "non_max_suppression") + "\n"
           + _func(_lines("tools/getins.py"), "align_superpoint_label") + "\n"
           + "def postprocess(pred, pred_sem, batch, cfg, semantic_label_idx, superpoint, point_num, proposals_idx,\n"
             "                proposals_offset, clt_score):\n"
             "    for _once in range(1):\n" + textwrap.indent(body, "        ") + "\n"
             "        return clusters, cluster_scores, cluster_semantic_id\n"
             "    return None\n")
    ns = {"torch": torch}
    exec(compile(src, os.path.join(REF, "eval_map.py"), "exec"), ns)
    return ns


def pseudo(n, mult, mod=None):
    h = (np.arange(n, dtype=np.uint64) * np.uint64(mult)) % np.uint64(1 << 32)
    return h if mod is None else (h % np.uint64(mod))


def make_inputs(scene_npz, score_scale=1.0):
    """Proposal lists from a local-scene fixture + pseudo scores / classes / superpoints (all deterministic)."""
    d = np.load(scene_npz)
    pidx = d["ref_prop_idx"].astype(np.int64)
    poff = d["ref_prop_offset"].astype(np.int64)
    N = int(d["n_all"])
    n3 = N // 3
    P = len(poff) - 1
    score = (pseudo(P, 2246822519).astype(np.float64) / 4294967296.0 * score_scale).astype(np.float32)
    pred_sem = pseudo(N, 3266489917, 20).astype(np.int64)
    # superpoints: sort the folded points by (first proposal containing them + jitter) and cut blocks of 40
    first = np.full(n3, P + 3, np.float64)
    np.minimum.at(first, pidx[:, 1] % n3, pidx[:, 0].astype(np.float64))
    jitter = pseudo(n3, 668265263).astype(np.float64) / 4294967296.0 * 1.5
    order = np.argsort(first + jitter, kind="stable")
    sp = np.empty(n3, np.int64)
    sp[order] = np.arange(n3) // 40
    return dict(proposals_idx=pidx, proposals_offset=poff, clt_score=score, pred_sem=pred_sem, superpoint=sp, point_num=N)


def run_reference(ns, inp, nms=0.10, score_thr=0.07, npoint=101):
    cfg = types.SimpleNamespace(TEST_NMS_THRESH=nms, TEST_SCORE_THRESH=score_thr, TEST_NPOINT_THRESH=npoint)
    sem_idx = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39]  # eval_map.py:32
    out = ns["postprocess"](None, torch.from_numpy(inp["pred_sem"]), None, cfg, sem_idx, torch.from_numpy(inp["superpoint"].copy()),
                            inp["point_num"], torch.from_numpy(inp["proposals_idx"].copy()),
                            torch.from_numpy(inp["proposals_offset"].copy()), torch.from_numpy(inp["clt_score"].copy()))
    return out


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: "cpu"
    ns = load_reference_block()
    out_dir = os.path.join(ROOT, "tests", "golden", "eval")
    os.makedirs(out_dir, exist_ok=True)
    sc = os.path.join(ROOT, "tests", "golden", "scenes")
    cases = [("e1_s2001", "s2001_test_b3.npz", 1.0, 0.10, 0.07, 101), ("e2_s2004", "s2004_test_b3_k2.npz", 1.0, 0.10, 0.07, 101),
             ("e3_s2001_strict", "s2001_test_b3.npz", 1.0, 0.50, 0.30, 400), ("e4_s2004_lowscore", "s2004_test_b3_k2.npz", 0.08, 0.10, 0.07, 101)]
    for name, scene, sscale, nms, sthr, npt in cases:
        inp = make_inputs(os.path.join(sc, scene), sscale)
        res = run_reference(ns, inp, nms, sthr, npt)
        if res is None:
            clusters = np.zeros((0, inp["point_num"] // 3), np.int32)
            scores, sem = np.zeros(0, np.float32), np.zeros(0, np.int64)
        else:
            clusters, scores, sem = (t.numpy() for t in res)
        # clusters are disjoint after the superpoint alignment: store them as one label per point
        assert clusters.sum(0).max() <= 1
        label = np.full(clusters.shape[1], -100, np.int32)
        for c in range(clusters.shape[0]):
            label[clusters[c] == 1] = c
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), scene=np.array(scene), score_scale=np.float64(sscale),
                            nms=np.float64(nms), score_thr=np.float64(sthr), npoint=np.int64(npt),
                            ref_label=label, ref_scores=scores.astype(np.float32), ref_sem=sem.astype(np.int64),
                            ref_n=np.int64(clusters.shape[0]))
        print(name, "proposals", len(inp["proposals_offset"]) - 1, "-> clusters", clusters.shape[0], "labelled points", int((label >= 0).sum()))


if __name__ == "__main__":
    main()
