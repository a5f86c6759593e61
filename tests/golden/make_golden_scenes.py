"""Generates tests/golden/scenes/*.npz: golden vectors for the "local scene" proposal construction and
``get_proposal`` (network/PBNet.py:180-234 and :317-346).

The reference code for this step is plain torch Python inside ``PBNet.forward`` — it cannot be imported here
(MinkowskiEngine, matplotlib … are absent), so this script EXECUTES THE REFERENCE'S OWN SOURCE LINES: it reads
/root/reference/network/PBNet.py, cuts the block between the markers ``# ####cluster center handle`` and
``# ####voxel proposal`` (the body of the per-class loop after ``pbnet_ops.cluster``) and the methods
``get_center_index_sum`` / ``get_proposal`` out of the file text, and runs them unmodified on CPU tensors
(``Tensor.cuda`` is patched to the identity: there is no GPU in the build container).  Nothing is copied into
the repo; only inputs and outputs are stored.  The grouping results that feed the block come from the CPU
oracle (oracle/pb_oracle.c, pinned bit-exact against the compiled reference).

    python tests/golden/make_golden_scenes.py        # run in the build container (needs /root/reference)
"""
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pb_oracle as po  # noqa: E402
from pbnet_b200 import scenes  # noqa: E402

REF = "/root/reference/network/PBNet.py"


def _cut(lines, first_marker, end_marker):
    a = next(i for i, l in enumerate(lines) if first_marker in l)
    b = next(i for i, l in enumerate(lines) if end_marker in l and i > a)
    return lines[a:b]


def load_reference_block():
    lines = open(REF, encoding="utf-8").read().split("\n")
    body = textwrap.dedent("\n".join(_cut(lines, "# ####cluster center handle", "# ####voxel proposal")))
    src = ("def local_scene(self, task, sem_id, cluster_id, cluster_num, clt_ctr, ins_orig, ins_feat, ins_sem_score, ins_ind,\n"
           "                ins_bp_sum, ins_ins_label, list_xyz, list_feat, list_gt_mask, list_ins_idx):\n"
           + textwrap.indent(body, "    "))
    a = next(i for i, l in enumerate(lines) if "def get_center_index_sum" in l)
    b = next(i for i, l in enumerate(lines) if "def get_label_mask" in l)
    src += "\n" + textwrap.dedent("\n".join(lines[a:b]))
    a = next(i for i, l in enumerate(lines) if "def get_proposal" in l)
    b = next(i for i, l in enumerate(lines) if l.startswith("def model_fn"))
    src += "\n" + textwrap.dedent("\n".join(lines[a:b]))
    ns = {"torch": torch}
    exec(compile(src, REF, "exec"), ns)
    return ns


def pseudo_scores(n):
    """Deterministic stand-in for the mask scores (multiplicative hash of the entry index, in [0, 1))."""
    return ((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(1 << 32)).astype(np.float64).astype(
        np.float32) / np.float32(4294967296.0)


class SelfStub:
    """The attributes of PBNet the block reads (network/PBNet.py:33-35)."""

    def __init__(self, ns, cluster_batch, k_max=6.0):
        self.count_mean = torch.tensor([-1., -1., 3917., 12056., 2303., 8331., 3948., 3166., 5629., 11719., 1003.,
                                        3317., 4912., 10221., 3889., 4136., 2120., 945., 3967., 2589.])
        self.K_max = torch.ones(20, dtype=torch.float32) * k_max
        self.cluster_batch = cluster_batch
        self._ns = ns

    def get_center_index_sum(self, clt_num, bs):
        return self._ns["get_center_index_sum"](self, clt_num, bs)


def make_case(ns, seed, n_points, copies, task, k_max=6.0, hp_frac=0.85):
    sc = scenes.make_scene(seed, n_points, hp_frac)
    calls = scenes.class_calls(sc, copies)
    rng = np.random.Generator(np.random.PCG64(seed * 7 + 1))
    n_all = sc["xyz_orig"].shape[0]
    # pseudo instance labels: 1 m grid cell of the original position; 6 % ignored (-100); every fourth cell ignored
    cell = np.floor(sc["xyz_orig"].astype(np.float64)).astype(np.int64)
    lab = cell[:, 0] * 64 + cell[:, 1] * 8 + cell[:, 2]
    lab = np.unique(lab, return_inverse=True)[1].astype(np.int64)
    lab[rng.random(n_all) < 0.06] = -100
    lab[lab % 4 == 3] = -100
    r18 = np.full(18, np.float32(scenes.RADIUS), np.float32)
    m18 = np.full(18, scenes.MIN_PTS, np.int32)
    stub = SelfStub(ns, copies, k_max)
    list_xyz, list_feat, list_gt, list_idx = [], [], [], []
    rec = dict(call_sem=[], call_points=[], seg_counts=[], cluster_id=[], cluster_num=[], center=[], ins_ind=[], ins_label=[],
               xyz_orig=[])
    for c in calls:
        out = po.oracle_binary_cluster(c["xyz_shift"], c["xyz_orig"], c["sem"].astype(np.int32), c["seg_counts"], r18, m18,
                                       0.05, True)
        I = c["xyz_orig"].shape[0]
        # point index in the batched cloud of `copies` copies (copy b holds points b*n_all .. (b+1)*n_all-1)
        ins_ind = np.concatenate([c["index"] + b * n_all for b in range(copies)]).astype(np.int64)
        ins_lab = np.concatenate([lab[c["index"]] for _ in range(copies)])
        feat = torch.zeros(I, 2)           # stand-in for the 32 backbone channels (only gathered, never computed on)
        score = torch.zeros(I)
        bp_sum = torch.from_numpy(np.concatenate([[0], np.cumsum(c["seg_counts"])]).astype(np.int32))
        k = int(out["cluster_num"].sum())
        ns["local_scene"](stub, task, int(c["sem_id"]), torch.from_numpy(out["cluster_id"].copy()),
                          torch.from_numpy(out["cluster_num"].copy()), torch.from_numpy(out["center"][:3 * k].copy()),
                          torch.from_numpy(c["xyz_orig"]), feat, score, torch.from_numpy(ins_ind), bp_sum,
                          torch.from_numpy(ins_lab), list_xyz, list_feat, list_gt, list_idx)
        rec["call_sem"].append(int(c["sem_id"]))
        rec["call_points"].append(I)
        rec["seg_counts"].append(c["seg_counts"])
        rec["cluster_id"].append(out["cluster_id"])
        rec["cluster_num"].append(out["cluster_num"])
        rec["center"].append(out["center"][:3 * k])
        rec["ins_ind"].append(ins_ind)
        rec["ins_label"].append(ins_lab)
        rec["xyz_orig"].append(c["xyz_orig"])
    lens = np.array([len(t) for t in list_idx], np.int64)
    ref_idx = torch.cat(list_idx).numpy().astype(np.int64) if list_idx else np.zeros(0, np.int64)
    ref_dpn = torch.cat([f[:, -1] for f in list_feat]).numpy().astype(np.float32) if list_feat else np.zeros(0, np.float32)
    ref_xyz = torch.cat(list_xyz).numpy().astype(np.float32) if list_xyz else np.zeros((0, 3), np.float32)
    ref_gt = torch.cat(list_gt).numpy().astype(np.int32) if list_gt else np.zeros(0, np.int32)
    d = dict(task=np.array(task), copies=np.int32(copies), k_max=np.float32(k_max), n_all=np.int64(n_all * copies),
             call_sem=np.array(rec["call_sem"], np.int32), call_points=np.array(rec["call_points"], np.int64),
             seg_counts=np.concatenate(rec["seg_counts"]).astype(np.int32),
             cluster_id=np.concatenate(rec["cluster_id"]).astype(np.int32),
             cluster_num=np.concatenate(rec["cluster_num"]).astype(np.int32),
             center=np.concatenate(rec["center"]).astype(np.float32),
             ins_ind=np.concatenate(rec["ins_ind"]).astype(np.int32),
             ins_label=np.concatenate(rec["ins_label"]).astype(np.int64),
             ref_lens=lens, ref_idx=ref_idx.astype(np.int32), ref_dpn=ref_dpn, ref_gt=ref_gt.astype(np.int8))
    # the gathered coordinates are exactly xyz_orig[position of the listed point]: checked here, not stored
    if len(list_idx):
        xo_all = np.concatenate(rec["xyz_orig"])
        order = np.argsort(np.concatenate(rec["ins_ind"]), kind="stable")
        where = order[np.searchsorted(np.concatenate(rec["ins_ind"])[order], ref_idx)]
        assert np.array_equal(xo_all[where], ref_xyz)
    # ---- get_proposal (network/PBNet.py:317-346) on pseudo mask scores (regenerated by the test: pseudo_scores) -----
    if len(list_idx):
        ms = torch.from_numpy(pseudo_scores(int(lens.sum()))).view(-1, 1)
        if len(lens) > 2:  # one proposal empty after thresholding (exercises the "remove null proposals" renumbering)
            o = int(lens[:1].sum())
            ms[o:o + int(lens[1])] = 0.0
        p_idx, p_off, p_ids, p_ms = ns["get_proposal"](None, list_idx, ms)
        assert np.array_equal(p_ms.numpy(), ms.numpy().reshape(-1)[ms.numpy().reshape(-1) > 0.45])
        d.update(ref_prop_idx=p_idx.numpy().astype(np.int32), ref_prop_offset=p_off.numpy().astype(np.int64),
                 ref_prop_ids=p_ids.numpy().astype(np.int32))
    return d


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference block calls .cuda() on tiny index tensors
    ns = load_reference_block()
    out_dir = os.path.join(ROOT, "tests", "golden", "scenes")
    os.makedirs(out_dir, exist_ok=True)
    cases = [("s2001_test_b3", 2001, 60000, 3, "test", 6.0), ("s2002_test_b1", 2002, 90000, 1, "test", 6.0),
             ("s2003_train_b2", 2003, 70000, 2, "train", 6.0), ("s2004_test_b3_k2", 2004, 50000, 3, "test", 2.0),
             ("s2005_train_b1", 2005, 16000, 1, "train", 6.0)]
    for name, seed, n, copies, task, kmax in cases:
        d = make_case(ns, seed, n, copies, task, kmax)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "calls", len(d["call_sem"]), "clusters", int(d["cluster_num"].sum()), "proposals", len(d["ref_lens"]),
              "entries", int(d["ref_lens"].sum()), "big", int((d["ref_dpn"] < 1).sum()))


if __name__ == "__main__":
    main()
