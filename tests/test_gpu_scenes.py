"""GPU: local-scene proposal lists and get_proposal (C ABI pb_local_scenes_plan/_fill, pb_get_proposal) against
(1) golden vectors produced by the reference's own source lines (network/PBNet.py:180-234, 317-346) and
(2) the CPU restatement on the grouping output of a larger scene."""
import glob
import os

import numpy as np
import pytest

from tests.test_scene_oracle import SCENES, call_seg_counts, golden_scores

pytestmark = pytest.mark.gpu


def _dev(a, dt=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dt is not None:
        t = t.to(dt)
    return t.cuda()


@pytest.mark.parametrize("path", SCENES, ids=[os.path.basename(p)[:-4] for p in SCENES])
def test_local_scenes_equal_reference_lines(path):
    import torch
    from pbnet_b200 import grouping
    d = np.load(path)
    train = str(d["task"]) != "test"
    km = np.full(20, int(d["k_max"]), np.int32)
    out = grouping.build_local_scenes(_dev(d["cluster_id"]), _dev(d["cluster_num"]), _dev(d["center"]), d["seg_counts"],
                                      call_seg_counts(d), d["call_sem"], point_map=_dev(d["ins_ind"], torch.int64),
                                      ins_label=_dev(d["ins_label"]) if train else None, k_max=km, want_proposal_id=True)
    off = out["offsets"].cpu().numpy()
    assert np.array_equal(np.diff(off), d["ref_lens"])
    assert np.array_equal(out["index"].cpu().numpy(), d["ref_idx"].astype(np.int64))
    assert np.array_equal(out["dpn"].cpu().numpy().view(np.uint32), d["ref_dpn"].view(np.uint32))
    assert np.array_equal(out["proposal"].cpu().numpy(), np.repeat(np.arange(len(d["ref_lens"])), d["ref_lens"]))
    if train:
        assert np.array_equal(out["gt"].cpu().numpy(), d["ref_gt"].astype(np.int32))
    # get_proposal on the same lists
    ms = golden_scores(d)
    pidx, poff, ids, pms = grouping.get_proposal(out["offsets"], out["index"], _dev(ms))
    assert np.array_equal(pidx.cpu().numpy(), d["ref_prop_idx"].astype(np.int64))
    assert np.array_equal(poff.cpu().numpy(), d["ref_prop_offset"])
    assert np.array_equal(ids.cpu().numpy(), d["ref_prop_ids"].astype(np.int64))
    assert np.array_equal(pms.cpu().numpy(), ms[ms > np.float32(0.45)])


def test_local_scenes_after_fused_grouping_match_oracle():
    """group_instances -> build_local_scenes on one 120k-point scene with 3 copies vs the CPU restatement fed with the
    same grouping output."""
    import torch
    from oracle import scene_oracle as so
    from pbnet_b200 import grouping, scenes
    sc = scenes.make_scene(4242, 120_000)
    copies = 3
    xyz = np.concatenate(scenes.rotate_copies(sc["xyz_orig"], copies))
    off = np.concatenate(scenes.rotate_copies(sc["offset"], copies))
    sem = np.tile(sc["sem"], copies)
    bh = np.repeat(np.arange(copies), sc["sem"].shape[0])
    res = grouping.group_instances(_dev(xyz), _dev(off), _dev(sem), _dev(bh.astype(np.int32)), scenes.RADIUS, scenes.MIN_PTS, copies)
    assert len(res) > 5
    cid = torch.cat([r["cluster_id"] for r in res])
    cnum = torch.cat([r["cluster_num"] for r in res])
    ctr = torch.cat([r["clt_ctr"].reshape(-1) for r in res])
    pmap = torch.cat([r["ins_ind"] for r in res])
    seg = np.concatenate([r["seg_counts"] for r in res])
    csem = np.array([r["sem_id"] for r in res], np.int32)
    calls = np.full(len(res), copies, np.int32)
    out = grouping.build_local_scenes(cid, cnum, ctr, seg, calls, csem, point_map=pmap)
    want = so.local_scenes(cid.cpu().numpy(), cnum.cpu().numpy(), ctr.cpu().numpy(), seg, calls, csem)
    assert np.array_equal(np.diff(out["offsets"].cpu().numpy()), want["lens"])
    assert np.array_equal(out["index"].cpu().numpy(), pmap.cpu().numpy()[want["pos"]])
    assert np.array_equal(out["dpn"].cpu().numpy().view(np.uint32), want["dpn"].view(np.uint32))
    assert np.array_equal(out["cluster"].cpu().numpy(), want["cluster"].astype(np.int32))
    assert (want["dpn"] < 1).any()  # the scene does contain local scenes with neighbours


def test_local_scenes_rejects_inconsistent_tables():
    import torch
    from pbnet_b200 import grouping
    from pbnet_b200._lib import PBError
    cid = torch.tensor([0, 0, 1, 5], dtype=torch.int32).cuda()     # id 5 does not exist
    cnum = torch.tensor([2], dtype=torch.int32).cuda()
    ctr = torch.zeros(6, dtype=torch.float32).cuda()
    with pytest.raises(PBError):
        grouping.build_local_scenes(cid, cnum, ctr, [4], [1], [5])


def test_local_scenes_empty():
    import torch
    from pbnet_b200 import grouping
    cid = torch.full((10,), -1, dtype=torch.int32).cuda()
    out = grouping.build_local_scenes(cid, torch.zeros(1, dtype=torch.int32).cuda(), torch.zeros(0, dtype=torch.float32).cuda(),
                                      [10], [1], [4])
    assert out["offsets"].cpu().tolist() == [0] and out["index"].numel() == 0


def test_propose_chain_matches_piecewise_torch():
    """grouping.propose (network/PBNet.py:144-247 on the device) vs the same steps spelled out with torch indexing."""
    import torch
    from pbnet_b200 import grouping, scenes
    sc = scenes.make_scene(777, 80_000)
    copies = 3
    xyz = _dev(np.concatenate(scenes.rotate_copies(sc["xyz_orig"], copies)))
    off = _dev(np.concatenate(scenes.rotate_copies(sc["offset"], copies)))
    sem = _dev(np.tile(sc["sem"], copies))
    bh = _dev(np.repeat(np.arange(copies), sc["sem"].shape[0]).astype(np.int32))
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.rand((xyz.shape[0], 32), device="cuda", generator=g)
    sfp = torch.softmax(torch.rand((xyz.shape[0], 20), device="cuda", generator=g), dim=1)
    out = grouping.propose(xyz, off, sem, bh, feat, sfp, scenes.RADIUS, scenes.MIN_PTS, copies)
    s = out["scenes"]
    idx, pid = s["index"], s["proposal"].to(torch.int64)
    want = torch.cat([feat[idx], sfp[idx, out["proposal_sem"].to(torch.int64)[pid]][:, None], s["dpn"][:, None]], dim=1)
    assert out["features"].shape == (idx.shape[0], 34) and torch.equal(out["features"], want)
    # class of every proposal = class of its member points
    assert torch.equal(sem[idx[s["offsets"][:-1]]].to(torch.int32), out["proposal_sem"])
    # voxelization contract (ME semantics, parity unpinned): floor(xyz/0.02) per proposal, unique rows, inverse map
    q = torch.cat([pid[:, None], torch.floor(xyz[idx] / 0.02).to(torch.int64)], dim=1)
    vm = out["voxel_map"]
    assert torch.equal(out["voxel_coords"].to(torch.int64)[vm.inverse], q)
    assert out["voxel_coords"].shape[0] == torch.unique(q, dim=0).shape[0]
    assert torch.equal(out["voxel_features"], want[vm.index])


def test_front_end_rejects_bad_copy_index_and_skips_small_classes():
    """pb_group_front (device-side class loop): a copy index outside [0, cluster_batch) is an error (the reference asserts
    the per-copy counts add up, network/PBNet.py:282-287); classes below count_mean*0.05 are skipped; no class kept -> None."""
    import torch
    from pbnet_b200 import grouping, scenes
    from pbnet_b200._lib import PBError
    sc = scenes.make_scene(778, 30_000)
    xyz, off, sem = _dev(sc["xyz_orig"]), _dev(sc["offset"]), _dev(sc["sem"])
    bh = torch.zeros(xyz.shape[0], dtype=torch.int32, device="cuda")
    flat = grouping.group_instances_flat(xyz, off, sem, bh, scenes.RADIUS, scenes.MIN_PTS, 1)
    cnt = np.bincount(sc["sem"], minlength=20)
    want = [c for c in range(2, 20) if not (np.float32(cnt[c]) < scenes.COUNT_MEAN[c] * np.float32(0.05))]
    assert flat["classes"].tolist() == want
    assert flat["seg_counts"].reshape(-1).tolist() == [int(cnt[c]) for c in want]
    pi = flat["point_index"].cpu().numpy()
    o = 0
    for c in want:   # ins_ind of every class: ascending point indices of that class
        assert np.array_equal(pi[o:o + cnt[c]], np.nonzero(sc["sem"] == c)[0])
        o += cnt[c]
    bad = bh.clone()
    bad[17] = 1
    with pytest.raises(PBError):
        grouping.group_instances_flat(xyz, off, sem, bad, scenes.RADIUS, scenes.MIN_PTS, 1)
    only_bg = torch.zeros_like(sem)
    assert grouping.group_instances_flat(xyz, off, only_bg, bh, scenes.RADIUS, scenes.MIN_PTS, 1) is None
    assert grouping.group_instances(xyz, off, only_bg, bh, scenes.RADIUS, scenes.MIN_PTS, 1) == []
    # int64 copy indices are accepted as they are
    flat64 = grouping.group_instances_flat(xyz, off, sem, bh.to(torch.int64), scenes.RADIUS, scenes.MIN_PTS, 1)
    assert torch.equal(flat64["cluster_id"], flat["cluster_id"]) and torch.equal(flat64["point_index"], flat["point_index"])
