"""Evaluation post-processing of the proposals on the device: the replacement of
eval_map.py:63-121 + tools/mIOU.py:77-87 (non_max_suppression) + tools/getins.py:72-98 (align_superpoint_label).

The reference materialises a dense ``nProposal x N`` int matrix, a dense fp32 ``mm`` for the cross IoU, runs the NMS and
the superpoint vote on the CPU (numpy / scipy) and loops over clusters in Python.  ``postprocess`` keeps sparse lists on
the GPU and returns one label per point — the clusters are disjoint after the alignment, so
``clusters[c] = (label == c)`` is exactly the reference's final mask matrix.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import PBError
from .cluster import default_context, stream_handle

SEMANTIC_LABEL_IDX = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39], np.int64)  # eval_map.py:32


def postprocess(proposals_idx: torch.Tensor, proposals_offset: torch.Tensor, clt_score: torch.Tensor, pred_sem: torch.Tensor,
                superpoint: torch.Tensor, point_num: int, nms_thresh: float = 0.10, score_thresh: float = 0.07,
                npoint_thresh: int = 101, copies: int = 3, semantic_label_idx=SEMANTIC_LABEL_IDX, n_superpoints: int | None = None):
    """proposals_idx i64[M,2], proposals_offset i64[P+1] (``get_proposal``), clt_score f32[P], pred_sem i64[point_num],
    superpoint i64[point_num//copies] (compressed ids) — CUDA tensors.  Thresholds default to config_test.py:53-55.
    Returns dict(label i32[point_num//copies], scores f32[C], sem i64[C], proposal i32[C]); ``dense_masks`` rebuilds the
    reference's ``clusters`` matrix."""
    dev = proposals_idx.device
    assert dev.type == "cuda", "postprocess is a device-resident op"
    ctx = default_context(dev.index)
    L = ctx._lib
    pidx = proposals_idx.to(torch.int64).contiguous()
    poff = proposals_offset.to(device=dev, dtype=torch.int64).contiguous()
    score = clt_score.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    psem = pred_sem.to(device=dev, dtype=torch.int64).contiguous()
    sp = superpoint.to(device=dev, dtype=torch.int64).contiguous()
    M, P = int(pidx.shape[0]), int(poff.shape[0]) - 1
    n3 = int(point_num) // int(copies)
    if sp.shape[0] != n3 or psem.shape[0] != point_num or score.shape[0] != P:
        raise ValueError("superpoint needs point_num//copies entries, pred_sem point_num, clt_score one per proposal")
    if n_superpoints is None:
        n_superpoints = int(sp.max().item()) + 1 if n3 else 0
    table = np.ascontiguousarray(np.asarray(semantic_label_idx, np.int64))
    label = torch.empty(n3, dtype=torch.int32, device=dev)
    cap = max(P, 1)
    out_score = torch.empty(cap, dtype=torch.float32, device=dev)
    out_sem = torch.empty(cap, dtype=torch.int64, device=dev)
    out_prop = torch.empty(cap, dtype=torch.int32, device=dev)
    C = ctypes.c_int64(0)
    rc = L.pb_eval_postprocess(ctx._h, pidx.data_ptr(), M, poff.data_ptr(), P, score.data_ptr(), psem.data_ptr(), int(point_num),
                               int(copies), sp.data_ptr(), int(n_superpoints), table.ctypes.data, int(table.shape[0]),
                               float(np.float32(score_thresh)), int(npoint_thresh), float(np.float32(nms_thresh)), label.data_ptr(),
                               out_score.data_ptr(), out_sem.data_ptr(), out_prop.data_ptr(), cap, ctypes.byref(C),
                               stream_handle(torch.cuda.current_stream(dev)))
    if rc != 0:
        raise PBError(rc, L.pb_last_error(ctx._h).decode())
    c = int(C.value)
    return dict(label=label, scores=out_score[:c], sem=out_sem[:c], proposal=out_prop[:c])


def dense_masks(label: torch.Tensor, n_clusters: int) -> torch.Tensor:
    """The reference's final ``clusters`` matrix (eval_map.py:112-118): int32[n_clusters, n_points] one-hot of ``label``."""
    return (label[None, :] == torch.arange(n_clusters, device=label.device, dtype=label.dtype)[:, None]).to(torch.int32)
