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
        sem_id, ins_ind (point indices, ascending), seg_counts (host, points per scene copy), cluster_id,
        cluster_num [cluster_batch], den_queue (= degree + 1, as pbnet_ops.cluster returns it), clt_ctr [K,3]
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
                        den_queue=out["degree"][p0:p0 + npts] + 1, seg_counts=seg_counts[ci].copy(),
                        clt_ctr=out["center"][3 * k0:3 * (k0 + k)].view(-1, 3)))
        p0 += npts
        k0 += k
    return res


# =====================================================================================================
# local scenes + get_proposal (network/PBNet.py:180-234, 317-346) on the device
# =====================================================================================================
import ctypes  # noqa: E402

from ._lib import PBError  # noqa: E402
from .cluster import stream_handle  # noqa: E402

K_MAX = np.full(20, 6, np.int32)  # network/PBNet.py:35  self.K_max = torch.ones(20) * 6


def _big_thresholds(count_mean):
    """``valid_idx.shape[0] > self.count_mean[sem_id] * 0.2`` (network/PBNet.py:210): the product is an fp32 tensor."""
    return (torch.as_tensor(np.asarray(count_mean, np.float32)) * 0.2).numpy().astype(np.float32)


def build_local_scenes(cluster_id: torch.Tensor, cluster_num: torch.Tensor, center: torch.Tensor, seg_counts, call_seg_counts,
                       call_sem, point_map: torch.Tensor | None = None, ins_label: torch.Tensor | None = None,
                       k_max=K_MAX, count_mean=COUNT_MEAN, want_proposal_id: bool = False):
    """Proposal point lists of ALL classes in one launch sequence (replaces the per-cluster Python loops of
    network/PBNet.py:180-234).

    cluster_id i32[n], cluster_num i32[S], center f32[3K]: CUDA tensors as returned by ``Context.binary_cluster`` for a
    batched call (ids restart in every call); seg_counts i32[S], call_seg_counts i32[C], call_sem i32[C]: host tables;
    point_map i64[n] (optional, CUDA): index of every input point in the full cloud (``ins_ind``) — the lists then hold
    those indices (``list_ins_idx``) instead of positions; ins_label i64[n] (optional, CUDA) switches the training
    branch on (mode label, ``-100`` skip, ground-truth mask).

    Returns dict(offsets i64[P+1], index i64[E], dpn f32[E], cluster i32[P], gt i32[E] | None, proposal i32[E] | None).
    Proposal p is ``index[offsets[p]:offsets[p+1]]``: the members of cluster ``cluster[p]`` in ascending order
    (weight 1) followed, for clusters larger than ``count_mean[sem]*0.2``, by the members of its nearest clusters.
    """
    dev = cluster_id.device
    assert dev.type == "cuda", "build_local_scenes is a device-resident op"
    ctx = default_context(dev.index)
    L = ctx._lib
    n = int(cluster_id.shape[0])
    seg = np.ascontiguousarray(np.asarray(seg_counts), dtype=np.int32)
    calls = np.ascontiguousarray(np.asarray(call_seg_counts), dtype=np.int32)
    csem = np.ascontiguousarray(np.asarray(call_sem), dtype=np.int32)
    thr = np.ascontiguousarray(_big_thresholds(count_mean))
    km = np.ascontiguousarray(np.asarray(k_max).astype(np.int32))   # int(self.K_max[sem_id]) truncates
    if thr.shape != (20,) or km.shape != (20,):
        raise ValueError("count_mean / k_max need 20 entries")
    for name, t, dt in (("cluster_id", cluster_id, torch.int32), ("cluster_num", cluster_num, torch.int32),
                        ("center", center, torch.float32), ("point_map", point_map, torch.int64),
                        ("ins_label", ins_label, torch.int64)):
        if t is None:
            continue
        if t.device != dev or t.dtype != dt or not t.is_contiguous():
            raise TypeError(f"{name}: need a contiguous {dt} tensor on {dev}")
    if (point_map is not None and point_map.shape[0] != n) or (ins_label is not None and ins_label.shape[0] != n):
        raise ValueError("point_map / ins_label must have one entry per point")
    K = int(center.numel() // 3)
    st = stream_handle(torch.cuda.current_stream(dev))
    P, E = ctypes.c_int64(0), ctypes.c_int64(0)
    ptr = lambda t: None if t is None else t.data_ptr()
    rc = L.pb_local_scenes_plan(ctx._h, ptr(cluster_id), seg.ctypes.data, int(seg.shape[0]), calls.ctypes.data, csem.ctypes.data,
                                int(calls.shape[0]), n, ptr(cluster_num), ptr(center), K, thr.ctypes.data, km.ctypes.data,
                                ptr(ins_label), ctypes.byref(P), ctypes.byref(E), st)
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    P, E = int(P.value), int(E.value)
    offsets = torch.empty(P + 1, dtype=torch.int64, device=dev)
    cluster = torch.empty(P, dtype=torch.int32, device=dev)
    index = torch.empty(E, dtype=torch.int64, device=dev)
    dpn = torch.empty(E, dtype=torch.float32, device=dev)
    gt = torch.empty(E, dtype=torch.int32, device=dev) if ins_label is not None else None
    pid = torch.empty(E, dtype=torch.int32, device=dev) if want_proposal_id else None
    rc = L.pb_local_scenes_fill(ctx._h, ptr(point_map), ptr(offsets), ptr(cluster), ptr(index), ptr(dpn), ptr(gt), ptr(pid), st)
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    return dict(offsets=offsets, index=index, dpn=dpn, cluster=cluster, gt=gt, proposal=pid)


def get_proposal(offsets: torch.Tensor, index: torch.Tensor, mask_score: torch.Tensor, mask_score_thd: float = 0.45):
    """``PBNet.get_proposal`` (network/PBNet.py:317-346) for proposal lists in CSR form (``build_local_scenes``):
    returns (proposals_idx i64[M,2], proposals_offset i64[P'+1], cluster_id_v i64[P'], proposals_ms f32[M]) — entries
    with ``mask_score > thd``, proposals renumbered 0..P'-1 without the empty ones.  (The reference stores the point
    index in a float32 column before the int64 cast, so it rounds indices above 2^24; this op keeps them exact.)"""
    dev = offsets.device
    assert dev.type == "cuda"
    ctx = default_context(dev.index)
    L = ctx._lib
    ms = mask_score.reshape(-1)
    for name, t, dt in (("offsets", offsets, torch.int64), ("index", index, torch.int64), ("mask_score", ms, torch.float32)):
        if t.device != dev or t.dtype != dt or not t.is_contiguous():
            raise TypeError(f"{name}: need a contiguous {dt} tensor on {dev}")
    P, E = int(offsets.shape[0]) - 1, int(index.shape[0])
    if ms.shape[0] != E:
        raise ValueError("one mask score per list entry")
    pidx = torch.empty((E, 2), dtype=torch.int64, device=dev)
    poff = torch.empty(P + 1, dtype=torch.int64, device=dev)
    ids = torch.empty(max(P, 1), dtype=torch.int64, device=dev)
    pms = torch.empty(E, dtype=torch.float32, device=dev)
    M, P2 = ctypes.c_int64(0), ctypes.c_int64(0)
    st = stream_handle(torch.cuda.current_stream(dev))
    rc = L.pb_get_proposal(ctx._h, offsets.data_ptr(), P, index.data_ptr(), ms.data_ptr(), E, float(np.float32(mask_score_thd)),
                           pidx.data_ptr(), poff.data_ptr(), ids.data_ptr(), pms.data_ptr(), ctypes.byref(M), ctypes.byref(P2), st)
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    M, P2 = int(M.value), int(P2.value)
    return pidx[:M], poff[:P2 + 1], ids[:P2], pms[:M]
