"""Caller-side fusion of PBNet.forward's per-class grouping loop.

The reference loops over the 18 foreground classes in Python (network/PBNet.py:151-179): ``nonzero`` + sort,
a ``.cpu()`` round trip per class, per-batch ``.sum()`` loops (``get_batch_offset`` :282-287) and one
``pbnet_ops.cluster`` call each.  ``group_instances`` takes the same device tensors and issues ONE
``pb_binary_cluster_batched`` launch sequence for all classes; per-class results are bit-identical to the loop.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import PBError
from .cluster import default_context, stream_handle
from .scenes import COUNT_MEAN


def _skip_thresholds(count_mean):
    """``count_mean[sem_id] * 0.05`` of network/PBNet.py:156 — an fp32 tensor product (entries 0/1 are wall / floor)."""
    return (torch.as_tensor(np.asarray(count_mean, np.float32)) * 0.05).numpy().astype(np.float32)


def group_instances_flat(xyz_original: torch.Tensor, offset_pred_p: torch.Tensor, sem_pred_p: torch.Tensor,
                         batch_head_p: torch.Tensor, radius: float, min_pts: int, cluster_batch: int,
                         count_mean=COUNT_MEAN, sem_num: int = 20):
    """The whole class loop of network/PBNet.py:151-179 on the device: ``pb_group_front`` (class histogram, skip rule,
    stable partition into class-major / copy-major order, fp32 ``orig + offset``: four kernels, no torch op) followed by
    ONE ``pb_binary_cluster_batched`` call.  Returns None when no class passes the skip, else a dict of flat results:
        classes i32[C] (host), seg_counts i32[C, cluster_batch] (host), call_clusters i64[C] (host), point_index i64[n]
        (= cat of the reference's ins_ind), cluster_id i32[n], cluster_num i32[C*cluster_batch], degree i32[n],
        center f32[3K], clt_sem i32[K], n_clusters
    ``batch_head_p`` must hold the scene copy of every point in [0, cluster_batch) (the reference slices the class's
    points into consecutive per-copy runs, i.e. it additionally assumes the copies are stored one after the other)."""
    dev = xyz_original.device
    assert dev.type == "cuda", "group_instances is the device-resident path; use pbnet_ops.cluster for CPU tensors"
    if sem_num != 20:
        raise ValueError("the class tables of the path have 20 entries (network/PBNet.py:33-34)")
    n = int(xyz_original.shape[0])
    ctx = default_context(dev.index)
    L = ctx._lib
    xyz = xyz_original if (xyz_original.dtype == torch.float32 and xyz_original.is_contiguous()) else xyz_original.to(torch.float32).contiguous()
    off = offset_pred_p if (offset_pred_p.dtype == torch.float32 and offset_pred_p.is_contiguous()) else offset_pred_p.to(torch.float32).contiguous()
    sem = sem_pred_p if (sem_pred_p.dtype == torch.int64 and sem_pred_p.is_contiguous()) else sem_pred_p.to(torch.int64).contiguous()
    bh = batch_head_p
    if bh.dtype not in (torch.int32, torch.int64) or not bh.is_contiguous():
        bh = bh.to(torch.int64).contiguous()
    if not (xyz.shape == (n, 3) and off.shape == (n, 3) and sem.shape == (n,) and bh.shape == (n,)):
        raise ValueError("xyz / offset [N,3], sem / batch [N] expected")
    for t in (off, sem, bh):
        if t.device != dev:
            raise TypeError("all inputs must live on the same CUDA device")
    buf = torch.empty((6, max(n, 1)), dtype=torch.float32, device=dev)       # x y z xo yo zo of the kept points
    sem32 = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    pidx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    thr = np.ascontiguousarray(_skip_thresholds(count_mean))
    keep20 = np.zeros(20, np.int32)
    segs = np.zeros(20 * cluster_batch, np.int32)
    n_kept = ctypes.c_int64(0)
    st = stream_handle(torch.cuda.current_stream(dev))
    rc = L.pb_group_front(ctx._h, xyz.data_ptr(), off.data_ptr(), sem.data_ptr(), bh.data_ptr(), int(bh.dtype == torch.int64), n,
                          int(cluster_batch), thr.ctypes.data, buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(),
                          buf[3].data_ptr(), buf[4].data_ptr(), buf[5].data_ptr(), sem32.data_ptr(), pidx.data_ptr(),
                          keep20.ctypes.data, segs.ctypes.data, ctypes.byref(n_kept), st)
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    m = int(n_kept.value)
    classes = np.nonzero(keep20)[0].astype(np.int32)
    if m == 0 or len(classes) == 0:
        return None
    seg_counts = segs[:len(classes) * cluster_batch].reshape(len(classes), cluster_batch).copy()
    r18 = np.full(18, np.float32(radius), np.float32)       # torch.ones(18) * radius -> float32 (pbnet_ops.py:33-36)
    m18 = np.full(18, int(min_pts), np.int32)
    out = ctx.binary_cluster(buf[0, :m], buf[1, :m], buf[2, :m], buf[3, :m], buf[4, :m], buf[5, :m], sem32[:m],
                             seg_counts.reshape(-1), r18, m18, 0.05, True,
                             call_seg_counts=np.full(len(classes), cluster_batch, np.int32))
    out.update(classes=classes, seg_counts=seg_counts, point_index=pidx[:m])
    return out


def group_instances(xyz_original: torch.Tensor, offset_pred_p: torch.Tensor, sem_pred_p: torch.Tensor,
                    batch_head_p: torch.Tensor, radius: float, min_pts: int, cluster_batch: int,
                    count_mean=COUNT_MEAN, sem_num: int = 20):
    """All inputs are CUDA tensors of N points: xyz_original [N,3] f32, offset_pred_p [N,3] f32,
    sem_pred_p [N] int64 (argmax class), batch_head_p [N] int (scene copy of each point).

    Returns a list with one dict per class that passes the ``count < count_mean * 0.05`` skip
    (network/PBNet.py:156), in class order:
        sem_id, ins_ind (point indices, ascending), seg_counts (host, points per scene copy), cluster_id,
        cluster_num [cluster_batch], den_queue (= degree + 1, as pbnet_ops.cluster returns it), clt_ctr [K,3]
    (views of the flat result of ``group_instances_flat``)."""
    out = group_instances_flat(xyz_original, offset_pred_p, sem_pred_p, batch_head_p, radius, min_pts, cluster_batch,
                               count_mean, sem_num)
    if out is None:
        return []
    den = out["degree"] + 1
    res = []
    p0 = 0
    k0 = 0
    for ci, c in enumerate(out["classes"]):
        npts = int(out["seg_counts"][ci].sum())
        k = int(out["call_clusters"][ci])
        res.append(dict(sem_id=int(c), ins_ind=out["point_index"][p0:p0 + npts], cluster_id=out["cluster_id"][p0:p0 + npts],
                        cluster_num=out["cluster_num"][ci * cluster_batch:(ci + 1) * cluster_batch],
                        den_queue=den[p0:p0 + npts], seg_counts=out["seg_counts"][ci].copy(),
                        clt_ctr=out["center"][3 * k0:3 * (k0 + k)].view(-1, 3)))
        p0 += npts
        k0 += k
    return res


# =====================================================================================================
# local scenes + get_proposal (network/PBNet.py:180-234, 317-346) on the device
# =====================================================================================================
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


def scene_features(point_feat: torch.Tensor, sem_score: torch.Tensor, index: torch.Tensor, proposal: torch.Tensor,
                   proposal_sem: torch.Tensor, dpn: torch.Tensor) -> torch.Tensor:
    """``list_feat`` of network/PBNet.py:195,231 in one gather: rows ``[point_feat[i] | sem_score[i, class of the proposal] |
    dpn]`` for every list entry (``index``/``proposal``/``dpn`` from ``build_local_scenes(..., want_proposal_id=True)``;
    ``proposal_sem`` i32[P] = class of every proposal)."""
    dev = point_feat.device
    ctx = default_context(dev.index)
    L = ctx._lib
    pf = point_feat.to(torch.float32).contiguous()
    sc = sem_score.to(torch.float32).contiguous()
    ps = proposal_sem.to(device=dev, dtype=torch.int32).contiguous()
    E, C = int(index.shape[0]), int(pf.shape[1])
    out = torch.empty((E, C + 2), dtype=torch.float32, device=dev)
    rc = L.pb_scene_features(ctx._h, pf.data_ptr(), C, sc.data_ptr(), int(sc.shape[1]), index.data_ptr(), proposal.data_ptr(),
                             ps.data_ptr(), dpn.data_ptr(), E, out.data_ptr(), stream_handle(torch.cuda.current_stream(dev)))
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    return out


def propose(xyz_original: torch.Tensor, offset_pred_p: torch.Tensor, sem_pred_p: torch.Tensor, batch_head_p: torch.Tensor,
            point_feat_p: torch.Tensor, sem_score_sfp: torch.Tensor, radius: float, min_pts: int, cluster_batch: int,
            ins_label: torch.Tensor | None = None, voxel_size: float = 0.02, k_max=K_MAX, count_mean=COUNT_MEAN):
    """The whole cluster stage of ``PBNet.forward`` up to the second sparse tensor (network/PBNet.py:144-247) on the device:
    grouping of all classes (one batched call), local-scene lists, feature rows and the voxelization of the proposals.

    Returns dict(groups, scenes, features f32[E,C+2], voxel_features, voxel_coords i32[V,4], voxel_map) where
    ``voxel_map.inverse`` is ``inputs_v2.inverse_mapping`` (:247) and ``scenes['index']`` is ``cat(list_ins_idx)``."""
    from . import voxel
    flat = group_instances_flat(xyz_original, offset_pred_p, sem_pred_p, batch_head_p, radius, min_pts, cluster_batch, count_mean)
    if flat is None:
        return dict(groups=None, scenes=None, features=None, voxel_features=None, voxel_coords=None, voxel_map=None)
    cid, cnum, ctr, pmap = flat["cluster_id"], flat["cluster_num"], flat["center"], flat["point_index"]
    seg = flat["seg_counts"].reshape(-1)
    csem = flat["classes"].astype(np.int32)
    calls = np.full(len(csem), cluster_batch, np.int32)
    lab = ins_label[pmap].contiguous() if ins_label is not None else None   # ins_ins_label = ins_label[ins_ind] (:163)
    sc = build_local_scenes(cid, cnum, ctr, seg, calls, csem, point_map=pmap, ins_label=lab, k_max=k_max,
                            count_mean=count_mean, want_proposal_id=True)
    # class of every proposal = class of its cluster (clt_sem of the batched call, one entry per cluster, call-major)
    prop_sem = flat["clt_sem"][sc["cluster"].to(torch.int64)]
    feats = scene_features(point_feat_p, sem_score_sfp, sc["index"], sc["proposal"], prop_sem, sc["dpn"])
    coords = xyz_original[sc["index"]].to(torch.float32) / voxel_size       # :236 sem_xyz / 0.02
    vm = voxel.voxel_map(coords, None, batch=sc["proposal"])
    vfeat = voxel.voxel_rows(feats, vm, "pick")
    return dict(groups=flat, scenes=sc, proposal_sem=prop_sem, features=feats, voxel_features=vfeat, voxel_coords=vm.vcoords,
                voxel_map=vm)
