"""TEST INFRASTRUCTURE — CPU restatement of the evaluation post-processing of proposals.
Only tests/ may import this module.

Follows /root/reference:
  eval_map.py:63-66     class of a proposal = class of its first listed point (before the fold), mapped through
                        ``semantic_label_idx`` (:32)
  eval_map.py:68-71     fold the three rotated scene copies: point index % (point_num/3); dense 0/1 proposal masks
  eval_map.py:74-85     score threshold (``>``, fp32), point-count threshold (``>``)
  eval_map.py:87-98     cross IoU = inter / (n_i + n_j - inter) in fp32; tools/mIOU.py:77-87 greedy NMS in
                        descending score order, removing ``iou > threshold``
  eval_map.py:105-110   per-point label = LAST picked cluster containing the point; tools/getins.py:72-98
                        ``align_superpoint_label``: per superpoint the most frequent label (first maximum; unlabelled
                        points vote for the ignore label), broadcast back to the points
  eval_map.py:112-121   clusters rebuilt from the aligned labels, empty ones dropped

Pinned against those source lines executed on CPU: tests/golden/make_golden_eval.py -> tests/golden/eval/*.npz.
Works on sparse lists (no dense nProposal x N matrices).
"""
from __future__ import annotations

import numpy as np

SEMANTIC_LABEL_IDX = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39], np.int64)  # eval_map.py:32


def postprocess(proposals_idx, proposals_offset, clt_score, pred_sem, superpoint, point_num, nms_thresh=0.10,
                score_thresh=0.07, npoint_thresh=101, copies=3):
    """Returns dict(label i32[point_num//copies] (final cluster of every point or -100), scores f32[C], sem i64[C],
    picked i64[C] (proposal index of every final cluster))."""
    pidx = np.asarray(proposals_idx, np.int64)
    poff = np.asarray(proposals_offset, np.int64)
    score = np.asarray(clt_score, np.float32)
    P = len(poff) - 1
    n3 = point_num // copies
    sem = SEMANTIC_LABEL_IDX[np.asarray(pred_sem)[pidx[poff[:-1], 1]]] if P else np.zeros(0, np.int64)
    pts = [np.unique(pidx[poff[p]:poff[p + 1], 1] % n3) for p in range(P)]        # rows of proposals_pred
    npt = np.array([len(s) for s in pts], np.int64)
    keep = np.nonzero(score > np.float32(score_thresh))[0]
    keep = keep[npt[keep] > npoint_thresh]
    empty = dict(label=np.full(n3, -100, np.int32), scores=np.zeros(0, np.float32), sem=np.zeros(0, np.int64),
                 picked=np.zeros(0, np.int64))
    if len(keep) == 0:
        return empty
    # NMS (tools/mIOU.py:77-87); score ties: argsort()[::-1] of a stable sort puts the higher index first
    order = list(keep[np.argsort(score[keep], kind="stable")[::-1]])
    sets = {int(p): set(pts[int(p)].tolist()) for p in keep}
    pick = []
    while order:
        i = int(order.pop(0))
        pick.append(i)
        rest = []
        for j in order:
            inter = np.float32(len(sets[i] & sets[int(j)]))
            iou = inter / (np.float32(npt[i]) + np.float32(npt[j]) - inter)        # fp32, as the torch expression
            if not (iou > np.float32(nms_thresh)):
                rest.append(j)
        order = rest
    label = np.full(n3, -100, np.int64)
    for c, p in enumerate(pick):                                                  # later clusters overwrite (:107-109)
        label[pts[p]] = c
    nlab = len(pick)
    sp = np.asarray(superpoint, np.int64)
    col = np.where(label < 0, nlab, label)
    hist = np.zeros((int(sp.max()) + 1, nlab + 1), np.int64)
    np.add.at(hist, (sp, col), 1)
    sp_label = hist.argmax(1)                                                     # first maximum
    sp_label[sp_label == nlab] = -100
    label = sp_label[sp]
    alive = [c for c in range(nlab) if (label == c).any()]
    remap = np.full(nlab + 1, -100, np.int64)
    remap[alive] = np.arange(len(alive))
    final = np.where(label >= 0, remap[np.maximum(label, 0)], -100).astype(np.int32)
    pick = np.asarray(pick, np.int64)[alive]
    return dict(label=final, scores=score[pick], sem=sem[pick], picked=pick)
