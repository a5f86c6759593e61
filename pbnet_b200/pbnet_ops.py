"""Host-side mirror of the reference operator ``pbnet_ops.cluster``
(lib/PB_lib/torch_io/pbnet_ops.py:12-82): same name, argument meaning, return tuple and
non-differentiability, on top of the sm_100a library.

    cluster(ins_offseted[I,3], ins_orig[I,3], sem[I], ins_bp[B], radius, min_pts, batch_size)
        -> (cluster_id i32[I], cluster_num i32[B], den_queue+1 i32[I], center f32[3*K])

CPU tensors reproduce the reference call exactly; CUDA tensors skip the host round trip that
network/PBNet.py:165,176-178 forces on the reference.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .cluster import default_context


class Cluster(Function):
    @staticmethod
    def forward(ctx, ins_offseted, ins_orig, sem, ins_bp, radius, min_pts, batch_size):
        dev = ins_offseted.device
        f32 = dict(dtype=torch.float32)
        # SoA split (pbnet_ops.py:16-18, 27-29): one transpose-copy per coordinate set
        so = ins_offseted.to(**f32).t().contiguous()
        oo = ins_orig.to(device=dev, **f32).t().contiguous()
        sem32 = sem.to(device=dev, dtype=torch.int32).contiguous()
        # the reference overwrites batch_size with ins_bp.shape[0] (pbnet_ops.py:43)
        segs = ins_bp.to(torch.int32).cpu()
        radius18 = (torch.ones(18) * radius).to(torch.float32)  # pbnet_ops.py:33-36
        min_pts18 = (torch.ones(18) * min_pts).to(torch.int32)
        pb = default_context(dev.index if dev.type == "cuda" else
                             (torch.cuda.current_device() if torch.cuda.is_available() else 0))
        out = pb.binary_cluster(so[0], so[1], so[2], oo[0], oo[1], oo[2], sem32, segs, radius18, min_pts18,
                                0.05, True)  # para_f, nv_flag: pbnet_ops.py:70-71
        for t in (out["cluster_id"], out["cluster_num"], out["degree"], out["center"]):
            ctx.mark_non_differentiable(t)
        return out["cluster_id"], out["cluster_num"], out["degree"] + 1, out["center"]

    @staticmethod
    def backward(ctx, *a):
        return None, None, None, None, None, None, None


cluster = Cluster.apply
