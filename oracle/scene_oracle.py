"""TEST INFRASTRUCTURE — CPU restatement of the "local scene" proposal construction and of ``get_proposal``
(SURVEY.md §8 row f1).  Only tests/ and __graft_entry__.smoke() may import this module.

Follows /root/reference/network/PBNet.py:
  :180-183   centres per segment, ``get_center_index_sum`` (:296-301) = prefix sum of ``cluster_num``
  :185-202   per segment: ``para_k = min(cluster_num-1, int(K_max[sem]))``, ``peak_v``, ``cdist`` + ``topk`` of the centres
  :204-222   per cluster: member list (ascending), mode label / ``-100`` skip (training), local scene for clusters
             larger than ``count_mean[sem]*0.2`` = own members (weight 1) + members of the para_k nearest
             clusters (weights ``peak_v[k]``)
  :223-233   ground-truth mask of the listed points (training)
  :317-346   ``get_proposal``: threshold the mask scores at 0.45, renumber non-empty proposals

Pinned against the reference's own source lines executed on CPU: tests/golden/make_golden_scenes.py ->
tests/golden/scenes/*.npz (tests/test_scene_oracle.py).
"""
from __future__ import annotations

import numpy as np
import torch

COUNT_MEAN = torch.tensor([-1., -1., 3917., 12056., 2303., 8331., 3948., 3166., 5629., 11719., 1003.,
                           3317., 4912., 10221., 3889., 4136., 2120., 945., 3967., 2589.])  # PBNet.py:33-34


def local_scenes(cluster_id, cluster_num, center, seg_counts, call_seg_counts, call_sem, ins_label=None, k_max=6.0,
                 count_mean=COUNT_MEAN):
    """All calls of one forward concatenated (the layout of pb_binary_cluster_batched): cluster_id i32[n] (ids restart
    at 0 in every call), cluster_num i32[S], center f32[3K], seg_counts i32[S], call_seg_counts i32[C], call_sem i32[C];
    ins_label i64[n] switches the training branch on.  Returns dict(lens i64[P], pos i64[E] (position of every listed
    point in the concatenated input), dpn f32[E], gt i32[E] or None, cluster i64[P] (global cluster index))."""
    cluster_id = np.asarray(cluster_id)
    cluster_num = np.asarray(cluster_num).astype(np.int64)
    center = torch.from_numpy(np.asarray(center, np.float32)).view(-1, 3)
    seg_start = np.concatenate([[0], np.cumsum(np.asarray(seg_counts, np.int64))])
    lens, pos, dpn, gt, which = [], [], [], [], []
    s0 = 0
    k0 = 0
    for c, ns in enumerate(np.asarray(call_seg_counts, np.int64)):
        sem_id = int(call_sem[c])
        cn = cluster_num[s0:s0 + ns]
        ctr_offset = np.concatenate([[0], np.cumsum(cn)])                       # :183 get_center_index_sum
        clt_ctr = center[k0:k0 + int(cn.sum())]
        for bi in range(int(ns)):                                                # :185
            if cn[bi] == 0:
                continue
            p0, p1 = int(seg_start[s0 + bi]), int(seg_start[s0 + bi + 1])
            batch_clt_id = cluster_id[p0:p1]
            para_k = min(int(cn[bi]) - 1, int(k_max))                            # :197
            if para_k > 0:
                peak_v = [0.5 * ((para_k + 1) - p_i) / (para_k + 1) for p_i in range(para_k + 1)]   # :199
                cc = clt_ctr[int(ctr_offset[bi]):int(ctr_offset[bi + 1])]
                knn_idx = torch.cdist(cc, cc).topk(k=int(cn[bi]), dim=1, largest=False)[1].numpy()  # :201-202
            for c_i in range(int(cn[bi])):                                       # :204
                valid = np.nonzero(batch_clt_id == c_i + ctr_offset[bi])[0]
                if ins_label is not None:
                    cur_gt = int(torch.mode(torch.from_numpy(np.asarray(ins_label[p0:p1])[valid]))[0])   # :206
                    if cur_gt == -100:
                        continue
                w = np.ones(valid.shape[0], np.float32)
                if bool(valid.shape[0] > count_mean[sem_id] * 0.2) and para_k > 0:   # :210 (fp32 product)
                    vs, ws = [valid], [w]
                    for k_i in range(para_k):
                        v = np.nonzero(batch_clt_id == knn_idx[c_i, k_i + 1] + ctr_offset[bi])[0]
                        vs.append(v)
                        ws.append((torch.ones(v.shape[0]) * peak_v[k_i]).numpy())     # :219 double -> fp32
                    valid, w = np.concatenate(vs), np.concatenate(ws)
                if ins_label is not None:                                        # :226-230
                    vl = np.asarray(ins_label[p0:p1])[valid]
                    m = (vl == cur_gt).astype(np.int32)
                    m[vl == -100] = -1
                    gt.append(m)
                lens.append(valid.shape[0])
                pos.append(valid + p0)
                dpn.append(w)
                which.append(k0 + int(ctr_offset[bi]) + c_i)
        s0 += int(ns)
        k0 += int(cn.sum())
    cat = lambda l, dt: np.concatenate(l).astype(dt) if l else np.zeros(0, dt)
    return dict(lens=np.asarray(lens, np.int64), pos=cat(pos, np.int64), dpn=cat(dpn, np.float32),
                gt=cat(gt, np.int32) if ins_label is not None else None, cluster=np.asarray(which, np.int64))


def get_proposal(lens, point_idx, mask_score, thd=0.45):
    """PBNet.get_proposal (:317-346).  lens i64[P] / point_idx[E] describe ``list_idx_proposal``; mask_score f32[E].
    Returns (proposals_idx i64[M,2], proposals_offset i64[P'+1], cluster_id_v i64[P'], proposals_ms f32[M])."""
    lens = np.asarray(lens, np.int64)
    pid = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    ms = np.asarray(mask_score, np.float32).reshape(-1)
    keep = np.nonzero(ms > np.float32(thd))[0]          # float32 tensor > python scalar compares in float32
    pidx = np.stack([pid[keep], np.asarray(point_idx, np.int64)[keep]], axis=1)
    ids, cnt = np.unique(pidx[:, 0], return_counts=True)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    if len(ids) and len(ids) != pidx[:, 0].max() + 1:   # "remove null proposals": renumber to 0..P'-1
        pidx[:, 0] = np.searchsorted(ids, pidx[:, 0])
    return pidx, off, ids.astype(np.int64), ms[keep]
