"""GPU: randomised parity of the caller-side rows (local scenes, get_proposal, evaluation post-processing) against their
CPU restatements (oracle/scene_oracle.py, oracle/eval_oracle.py — both pinned against the reference's own source lines).
Covers what the golden scenes cannot: empty segments, segments without clusters, K_max 0 / 1 / 16, every cluster "large",
unclustered points, ignore labels, empty proposals, all-below-threshold scores, score ties, many small superpoints."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a, dt=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dt is not None:
        t = t.to(dt)
    return t.cuda()


def _random_grouping(rng, n_calls, max_seg, max_pts, max_k):
    """A synthetic grouping result in the layout of pb_binary_cluster_batched (ids restart in every call)."""
    seg_counts, call_seg, call_sem, cid, cnum, centres = [], [], [], [], [], []
    for c in range(n_calls):
        ns = int(rng.integers(1, max_seg + 1))
        call_seg.append(ns)
        call_sem.append(int(rng.integers(2, 20)))
        base = 0
        for s in range(ns):
            npts = 0 if rng.random() < 0.15 else int(rng.integers(1, max_pts + 1))
            k = 0 if (npts == 0 or rng.random() < 0.2) else int(rng.integers(1, max_k + 1))
            ids = np.full(npts, -1, np.int32)
            if k:
                ids = rng.integers(-1, k, size=npts).astype(np.int32)
                ids[:min(k, npts)] = np.arange(min(k, npts))      # every cluster owns a point when possible
                k = int(ids.max()) + 1 if npts else 0
                present = np.zeros(k, bool)
                present[ids[ids >= 0]] = True
                remap = np.cumsum(present) - 1                      # compress to the ids that really occur
                ids = np.where(ids >= 0, remap[np.maximum(ids, 0)], -1).astype(np.int32)
                k = int(present.sum())
            seg_counts.append(npts)
            cnum.append(k)
            cid.append(np.where(ids >= 0, ids + base, -1).astype(np.int32))
            centres.append(rng.uniform(-5, 5, size=(k, 3)).astype(np.float32))
            base += k
    return (np.array(seg_counts, np.int32), np.array(call_seg, np.int32), np.array(call_sem, np.int32),
            np.concatenate(cid) if cid else np.zeros(0, np.int32), np.array(cnum, np.int32),
            np.concatenate(centres).reshape(-1) if centres else np.zeros(0, np.float32))


@pytest.mark.parametrize("seed,kmax,train", [(1, 6, False), (2, 0, False), (3, 1, True), (4, 16, False), (5, 3, True), (6, 6, True)])
def test_local_scenes_fuzz(seed, kmax, train):
    import torch
    from oracle import scene_oracle as so
    from pbnet_b200 import grouping
    rng = np.random.Generator(np.random.PCG64(seed))
    seg, calls, csem, cid, cnum, ctr = _random_grouping(rng, n_calls=7, max_seg=3, max_pts=900, max_k=(30 if seed == 4 else 9))
    n = int(seg.sum())
    # tiny class means so that most clusters count as "large" (threshold = count_mean * 0.2)
    cm = np.concatenate([[-1, -1], rng.uniform(20, 200, size=18)]).astype(np.float32)
    lab = rng.integers(0, 6, size=n).astype(np.int64)
    lab[rng.random(n) < 0.3] = -100
    pmap = rng.permutation(4 * n)[:n].astype(np.int64)
    km = np.full(20, kmax, np.int32)
    out = grouping.build_local_scenes(_dev(cid), _dev(cnum), _dev(ctr), seg, calls, csem, point_map=_dev(pmap),
                                      ins_label=_dev(lab) if train else None, k_max=km, count_mean=cm, want_proposal_id=True)
    want = so.local_scenes(cid, cnum, ctr, seg, calls, csem, ins_label=lab if train else None, k_max=float(kmax),
                           count_mean=torch.from_numpy(cm))
    assert np.array_equal(np.diff(out["offsets"].cpu().numpy()), want["lens"])
    assert np.array_equal(out["index"].cpu().numpy(), pmap[want["pos"]])
    assert np.array_equal(out["dpn"].cpu().numpy().view(np.uint32), want["dpn"].view(np.uint32))
    assert np.array_equal(out["cluster"].cpu().numpy(), want["cluster"].astype(np.int32))
    if train:
        assert np.array_equal(out["gt"].cpu().numpy(), want["gt"])
    if kmax > 0 and seed != 2:
        assert (want["dpn"] < 1).any()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_get_proposal_fuzz(seed):
    from oracle import scene_oracle as so
    from pbnet_b200 import grouping
    rng = np.random.Generator(np.random.PCG64(seed))
    P = 40
    lens = rng.integers(0, 300, size=P).astype(np.int64)
    lens[rng.random(P) < 0.2] = 0                              # empty proposals
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    E = int(off[-1])
    idx = rng.integers(0, 10 ** 6, size=E).astype(np.int64)
    ms = rng.random(E).astype(np.float32)
    ms[off[5]:off[9]] = 0.1                                    # proposals 5..8 lose every entry
    ms[rng.random(E) < 0.05] = np.float32(0.45)                # exactly the threshold: not kept (strict >)
    if seed == 13:
        ms[:] = 0.2                                            # nothing survives
    pidx, poff, ids, pms = grouping.get_proposal(_dev(off), _dev(idx), _dev(ms))
    w_idx, w_off, w_ids, w_ms = so.get_proposal(lens, idx, ms)
    assert np.array_equal(pidx.cpu().numpy().reshape(-1, 2), w_idx.reshape(-1, 2))
    assert np.array_equal(poff.cpu().numpy(), w_off)
    assert np.array_equal(ids.cpu().numpy(), w_ids)
    assert np.array_equal(pms.cpu().numpy(), w_ms)


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_eval_postprocess_fuzz(seed):
    from oracle import eval_oracle as eo
    from pbnet_b200 import evalpost
    rng = np.random.Generator(np.random.PCG64(seed))
    n3, copies, P = 4000, 3, 60
    N = n3 * copies
    rows = []
    for p in range(P):
        c = rng.integers(0, n3)
        w = int(rng.integers(20, 700))
        pts = (c + rng.integers(-w, w, size=int(rng.integers(10, 900)))) % n3          # overlapping index windows
        pts = pts + n3 * rng.integers(0, copies, size=pts.shape[0])                   # any of the three copies
        rows.append(np.stack([np.full(pts.shape[0], p), pts], 1))
    pidx = np.concatenate(rows).astype(np.int64)
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    score = rng.random(P).astype(np.float32)
    score[rng.integers(0, P, size=8)] = np.float32(0.5)                               # score ties
    pred_sem = rng.integers(0, 20, size=N).astype(np.int64)
    sp = rng.permutation(np.arange(n3) // int(rng.integers(3, 60))).astype(np.int64) if seed % 2 else (np.arange(n3) // 25).astype(np.int64)
    sp = np.unique(sp, return_inverse=True)[1].astype(np.int64)
    for nms, sthr, npt in ((0.1, 0.07, 101), (0.3, 0.0, 0), (0.0, 0.4, 50)):
        out = evalpost.postprocess(_dev(pidx), _dev(off), _dev(score), _dev(pred_sem), _dev(sp), N, nms, sthr, npt, copies)
        want = eo.postprocess(pidx, off, score, pred_sem, sp, N, nms, sthr, npt, copies)
        assert np.array_equal(out["label"].cpu().numpy(), want["label"]), (seed, nms, sthr, npt)
        assert np.array_equal(out["proposal"].cpu().numpy(), want["picked"].astype(np.int32))
        assert np.array_equal(out["scores"].cpu().numpy(), want["scores"])
        assert np.array_equal(out["sem"].cpu().numpy(), want["sem"])
