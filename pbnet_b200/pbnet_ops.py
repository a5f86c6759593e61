"""Host-side mirror of the reference operator ``pbnet_ops.cluster``
(lib/PB_lib/torch_io/pbnet_ops.py:12-82): same name, argument meaning, return tuple and
non-differentiability, on top of the sm_100a library.

    cluster(ins_offseted[I,3], ins_orig[I,3], sem[I], ins_bp[B], radius, min_pts, batch_size)
        -> (cluster_id i32[I], cluster_num i32[B], den_queue+1 i32[I], center f32[3*K])

CPU tensors reproduce the reference call exactly; CUDA tensors skip the host round trip that
network/PBNet.py:165,176-178 forces on the reference.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function

from .cluster import default_context


_TABLES = {}


def _tables(radius, min_pts):
    """``torch.ones(18) * radius`` -> float32, ``torch.ones(18) * min_pts`` -> int32 (pbnet_ops.py:33-36), cached."""
    key = (float(radius), float(min_pts))
    t = _TABLES.get(key)
    if t is None:
        t = _TABLES[key] = ((torch.ones(18) * radius).to(torch.float32).numpy().copy(),
                            (torch.ones(18) * min_pts).to(torch.int32).numpy().copy())
    return t


class Cluster(Function):
    @staticmethod
    def forward(ctx, ins_offseted, ins_orig, sem, ins_bp, radius, min_pts, batch_size):
        dev = ins_offseted.device
        # the reference overwrites batch_size with ins_bp.shape[0] (pbnet_ops.py:43)
        segs = ins_bp.numpy() if (ins_bp.dtype == torch.int32 and ins_bp.device.type == "cpu" and not ins_bp.requires_grad) \
            else ins_bp.detach().to(device="cpu", dtype=torch.int32).numpy()
        radius18, min_pts18 = _tables(radius, min_pts)
        if dev.type != "cuda":
            # CPU tensors (the reference's call pattern, network/PBNet.py:176): SoA split (pbnet_ops.py:16-18, 27-29) through
            # numpy — torch's strided CPU copy takes 0.33 ms for a 27 k x 3 transpose, numpy 0.04 ms
            def soa(t):
                if t.dtype != torch.float32 or t.device.type != "cpu" or t.requires_grad:
                    t = t.detach().to(device="cpu", dtype=torch.float32)
                return np.ascontiguousarray(t.numpy().T)
            so, oo = soa(ins_offseted), soa(ins_orig)
            sem32 = sem.detach().cpu().numpy().astype(np.int32, copy=False)
            if not sem32.flags["C_CONTIGUOUS"]:
                sem32 = np.ascontiguousarray(sem32)
            pb = default_context(torch.cuda.current_device() if torch.cuda.is_available() else 0)
            out = pb.binary_cluster(so[0], so[1], so[2], oo[0], oo[1], oo[2], sem32, segs, radius18, min_pts18, 0.05, True)
            deg = out["degree"]
            deg += 1                                   # den_queue + 1 (pbnet_ops.py:75); the array is ours
            res = (torch.from_numpy(out["cluster_id"]), torch.from_numpy(out["cluster_num"]), torch.from_numpy(deg),
                   torch.from_numpy(out["center"]))
            ctx.mark_non_differentiable(*res)
            return res
        # CUDA tensors: zero-copy, results stay on the device
        so = ins_offseted.to(torch.float32).t().contiguous()
        oo = ins_orig.to(device=dev, dtype=torch.float32).t().contiguous()
        sem32 = sem.to(device=dev, dtype=torch.int32).contiguous()
        pb = default_context(dev.index)
        out = pb.binary_cluster(so[0], so[1], so[2], oo[0], oo[1], oo[2], sem32, segs, radius18, min_pts18,
                                0.05, True)  # para_f, nv_flag: pbnet_ops.py:70-71
        den = out["degree"] + 1
        ctx.mark_non_differentiable(out["cluster_id"], out["cluster_num"], den, out["center"])
        return out["cluster_id"], out["cluster_num"], den, out["center"]

    @staticmethod
    def backward(ctx, *a):
        return None, None, None, None, None, None, None


cluster = Cluster.apply


def _iou_call(proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores, mask_label, mode):
    from .cluster import stream_handle
    from ._lib import PBError
    dev = proposals_idx.device
    if dev.type != "cuda":
        raise TypeError("get_iou / cal_iou_and_masklabel take CUDA tensors (as the reference asserts)")
    pb = default_context(dev.index)
    n_inst = int(instance_pointnum.size(0))
    n_prop = int(proposals_offset.size(0)) - 1
    iou = torch.zeros((n_prop, n_inst), dtype=torch.float32, device=dev)
    idx = proposals_idx.to(torch.int32).contiguous()
    off = proposals_offset.to(torch.int32).contiguous()
    lab = instance_labels.to(torch.int64).contiguous()
    pnum = instance_pointnum.to(torch.int32).contiguous()
    rc = pb._lib.pb_cal_iou_and_masklabel(pb._h, idx.data_ptr(), off.data_ptr(), lab.data_ptr(), pnum.data_ptr(),
                                          iou.data_ptr(), n_inst, n_prop,
                                          mask_scores.data_ptr() if mask_scores is not None else None,
                                          mask_label.data_ptr() if mask_label is not None else None, int(mode),
                                          stream_handle(torch.cuda.current_stream(dev)))
    if rc != 0:
        raise PBError(rc, pb._lib.pb_last_error(pb._h).decode())
    return iou


class GetIoU(Function):
    """Mirror of pbnet_ops.get_iou (lib/PB_lib/torch_io/pbnet_ops.py:85-111; network/PBNet.py:410)."""

    @staticmethod
    def forward(ctx, proposals_idx, proposals_offset, instance_labels, instance_pointnum):
        return _iou_call(proposals_idx, proposals_offset, instance_labels, instance_pointnum, None, None, 0)

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


get_iou = GetIoU.apply


class CalIoUAndMasklabel(Function):
    """Mirror of pbnet_ops.cal_iou_and_masklabel (lib/PB_lib/torch_io/pbnet_ops.py:114-141)."""

    @staticmethod
    def forward(ctx, proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores_sigmoid, mode):
        ms = mask_scores_sigmoid.to(torch.float32).contiguous()
        mask_label = torch.full(ms.shape, -1.0, dtype=torch.float32, device=ms.device)
        iou = _iou_call(proposals_idx, proposals_offset, instance_labels, instance_pointnum, ms, mask_label, mode)
        return iou, mask_label

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None, None


cal_iou_and_masklabel = CalIoUAndMasklabel.apply


class Get_normal_line(Function):
    """Mirror of lib/PB_lib/torch_io/pbnet_ops.py:143-173: numpy ``xyz[V,3]``, ``face[F,3]`` -> float32 tensor ``[V,3]``.
    Like the reference wrapper it passes ``num_face = V`` (:163 uses ``normal_line.shape[0]``), i.e. only the first V faces
    take part; with fewer than V faces the reference reads out of bounds — here that is a ValueError."""

    @staticmethod
    def forward(ctx, xyz, face):
        from .shim import PB_lib
        xyz = torch.from_numpy(xyz).type(torch.float32).contiguous()
        face = torch.from_numpy(face).type(torch.int32).contiguous()
        normal_line = torch.zeros_like(xyz).type(torch.float32).contiguous()
        PB_lib.cal_normal_line(xyz, face, normal_line, xyz.shape[0], normal_line.shape[0])
        return normal_line

    @staticmethod
    def backward(ctx, a=None):
        return None


get_normal_line = Get_normal_line.apply
