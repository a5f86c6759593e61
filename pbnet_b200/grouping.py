"""Caller-side fusion of PBNet.forward's per-class grouping loop.

The reference loops over the 18 foreground classes in Python (network/PBNet.py:151-179): ``nonzero`` + sort,
a ``.cpu()`` round trip per class, per-batch ``.sum()`` loops (``get_batch_offset`` :282-287) and one
``pbnet_ops.cluster`` call each.  ``group_instances`` takes the same device tensors and issues ONE
``pb_binary_cluster_batched`` launch sequence for all classes; per-class results are bit-identical to the loop.
"""
from __future__ import annotations

import numpy as np
import torch

from .cluster import default_context
from .scenes import COUNT_MEAN


def group_instances(xyz_original: torch.Tensor, offset_pred_p: torch.Tensor, sem_pred_p: torch.Tensor,
                    batch_head_p: torch.Tensor, radius: float, min_pts: int, cluster_batch: int,
                    count_mean=COUNT_MEAN, sem_num: int = 20):
    """All inputs are CUDA tensors of N points: xyz_original [N,3] f32, offset_pred_p [N,3] f32,
    sem_pred_p [N] int64 (argmax class), batch_head_p [N] int (scene copy of each point).

    Returns a list with one dict per class that passes the ``count < count_mean * 0.05`` skip
    (network/PBNet.py:156), in class order:
        sem_id, ins_ind (point indices, ascending), cluster_id, cluster_num [cluster_batch],
        den_queue (= degree + 1, as pbnet_ops.cluster returns it), clt_ctr [K,3]
    """
    dev = xyz_original.device
    assert dev.type == "cuda", "group_instances is the device-resident path; use pbnet_ops.cluster for CPU tensors"
    n = xyz_original.shape[0]
    sem = sem_pred_p.to(torch.int64)
    counts = torch.bincount(sem.clamp(0, sem_num - 1), minlength=sem_num)
    cm = torch.as_tensor(np.asarray(count_mean, np.float32), device=dev)
    keep_cls = (counts.to(torch.float32) >= cm * 0.05)
    keep_cls[:2] = False  # wall / floor are skipped (PBNet.py:151-152)
    sel = keep_cls[sem.clamp(0, sem_num - 1)] & (sem >= 2) & (sem < sem_num)
    idx = torch.nonzero(sel).view(-1)
    if idx.numel() == 0:
        return []
    key = sem[idx] * cluster_batch + batch_head_p[idx].to(torch.int64)
    order = torch.argsort(key, stable=True)          # class-major, batch-major, ascending point index inside
    pidx = idx[order]
    seg_counts = torch.bincount(key[order], minlength=sem_num * cluster_batch).view(sem_num, cluster_batch)
    classes = torch.nonzero(keep_cls).view(-1)
    seg_counts = seg_counts[classes].to(torch.int32).cpu().numpy()            # [n_classes, cluster_batch]
    classes = classes.cpu().numpy()
    orig = xyz_original[pidx].to(torch.float32)
    shifted = orig + offset_pred_p[pidx].to(torch.float32)                     # PBNet.py:165 (fp32 add)
    so, oo = shifted.t().contiguous(), orig.t().contiguous()
    sem32 = sem[pidx].to(torch.int32).contiguous()
    r18 = (torch.ones(18) * radius).to(torch.float32)
    m18 = (torch.ones(18) * min_pts).to(torch.int32)
    ctx = default_context(dev.index)
    out = ctx.binary_cluster(so[0], so[1], so[2], oo[0], oo[1], oo[2], sem32, seg_counts.reshape(-1), r18, m18, 0.05, True,
                             call_seg_counts=np.full(len(classes), cluster_batch, np.int32))
    res = []
    p0 = 0
    k0 = 0
    for ci, c in enumerate(classes):
        npts = int(seg_counts[ci].sum())
        k = int(out["call_clusters"][ci])
        res.append(dict(sem_id=int(c), ins_ind=pidx[p0:p0 + npts], cluster_id=out["cluster_id"][p0:p0 + npts],
                        cluster_num=out["cluster_num"][ci * cluster_batch:(ci + 1) * cluster_batch],
                        den_queue=out["degree"][p0:p0 + npts] + 1,
                        clt_ctr=out["center"][3 * k0:3 * (k0 + k)].view(-1, 3)))
        p0 += npts
        k0 += k
    return res
