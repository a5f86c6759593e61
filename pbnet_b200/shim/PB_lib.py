"""Drop-in module named ``PB_lib``: put ``pbnet_b200/shim`` on ``sys.path`` (or call
``pbnet_b200.install_shim()``) and the reference's unmodified wrapper
``lib/PB_lib/torch_io/pbnet_ops.py`` imports this instead of the reference's pybind11 extension.

Surface = the four ``m.def`` of lib/PB_lib/src/PB_lib_api.cpp:7-10.  ``binary_cluster`` keeps the
20-argument positional signature of lib/PB_lib/src/pbnet/cluster.h:13-18 and the reference's in-place
output convention (cluster.cu:112-118: ``center`` / ``clt_sem`` are resized to 3*K / K).
"""
from __future__ import annotations

import torch

from pbnet_b200.cluster import default_context


def binary_cluster(x, y, z, l1_norm, index_mapper, xo, yo, zo, sem, ins_bp, radius, min_pts, cluster_id,
                   cluster_num, den_queue, center, clt_sem, batch_size, para_f, nv_flag):
    """``l1_norm`` / ``index_mapper`` only steer the reference's slab pruning
    (binary.cu:49-69) and are accepted but unused.  Tensors may be CPU (as the reference requires) or
    all-CUDA (zero-copy fast path)."""
    dev = x.device.index if x.is_cuda else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    ctx = default_context(dev)
    n = x.shape[0]
    cap_c = center if center.shape[0] >= 3 * max(n, 1) else torch.empty(3 * max(n, 1), dtype=torch.float32,
                                                                       device=center.device)
    cap_s = clt_sem if clt_sem.shape[0] >= max(n, 1) else torch.empty(max(n, 1), dtype=torch.int32,
                                                                     device=clt_sem.device)
    segs = ins_bp[:int(batch_size)]
    out = ctx.binary_cluster(x, y, z, xo, yo, zo, sem, segs, radius, min_pts, float(para_f), bool(nv_flag),
                             cluster_id=cluster_id, cluster_num=cluster_num, degree=den_queue, center=cap_c,
                             clt_sem=cap_s)
    k = out["n_clusters"]
    if cap_c is center:
        center.resize_(3 * k)
    else:
        center.resize_(3 * k).copy_(cap_c[:3 * k])
    if cap_s is clt_sem:
        clt_sem.resize_(k)
    else:
        clt_sem.resize_(k).copy_(cap_s[:k])
    return None


def _iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, n_inst, n_prop,
         mask_scores, mask_label, mode):
    from pbnet_b200._lib import PBError
    from pbnet_b200.cluster import stream_handle
    dev = proposals_iou.device
    pb = default_context(dev.index)
    rc = pb._lib.pb_cal_iou_and_masklabel(
        pb._h, proposals_idx.data_ptr(), proposals_offset.data_ptr(), instance_labels.data_ptr(),
        instance_pointnum.data_ptr(), proposals_iou.data_ptr(), int(n_inst), int(n_prop),
        mask_scores.data_ptr() if mask_scores is not None else None,
        mask_label.data_ptr() if mask_label is not None else None, int(mode),
        stream_handle(torch.cuda.current_stream(dev)))
    if rc != 0:
        raise PBError(rc, pb._lib.pb_last_error(pb._h).decode())


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance, nProposal):
    """lib/PB_lib/src/iou/get_iou.cpp:9 — int32 / int32 / int64 / int32 CUDA tensors, float output in place."""
    _iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance, nProposal,
         None, None, 0)


def cal_iou_and_masklabel(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou,
                          nInstance, nProposal, mask_scores_sigmoid, mask_label, mode):
    """lib/PB_lib/src/cal_iou_and_masklabel/cal_iou_and_masklabel.cpp — outputs in place."""
    _iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance, nProposal,
         mask_scores_sigmoid, mask_label, mode)


def cal_normal_line(xyz, face, normal_line, num_vtx, num_face):
    """lib/PB_lib/src/normal/cal_normal.h:10 — CPU (or CUDA) float32 ``xyz[V,3]``, int32 ``face[F,3]``, in-place
    ``normal_line[V,3]``; ``num_face`` faces take part (the reference wrapper passes V, pbnet_ops.py:163)."""
    from pbnet_b200._lib import PBError
    ctx = default_context(xyz.device.index if xyz.is_cuda else (torch.cuda.current_device() if torch.cuda.is_available() else 0))
    if int(num_face) > face.shape[0]:
        raise ValueError(f"num_face={num_face} exceeds the {face.shape[0]} faces given (the reference would read out of bounds)")
    for t, dt in ((xyz, torch.float32), (face, torch.int32), (normal_line, torch.float32)):
        if t.dtype != dt or not t.is_contiguous() or t.is_cuda != xyz.is_cuda:
            raise TypeError("cal_normal_line: contiguous float32 xyz / normal_line and int32 face on one device")
    from pbnet_b200.cluster import stream_handle
    st = stream_handle(torch.cuda.current_stream(xyz.device)) if xyz.is_cuda else None  # ordered after xyz / face's producers
    rc = ctx._lib.pb_cal_normal_line(ctx._h, xyz.data_ptr(), face.data_ptr(), normal_line.data_ptr(), int(num_vtx), int(num_face),
                                     1 if xyz.is_cuda else 0, st)
    if rc != 0:
        raise PBError(rc, ctx._lib.pb_last_error(ctx._h).decode())
    return None
