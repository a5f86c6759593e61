"""numpy restatement of the reference's proposal IoU / mask-label kernels — TEST INFRASTRUCTURE ONLY.
Follows lib/PB_lib/src/iou/get_iou.cu:12-29 and
lib/PB_lib/src/cal_iou_and_masklabel/cal_iou_and_masklabel.cu:15-90 literally (the arithmetic
``(float)inter / ((float)(total) + 1e-5)`` is an fp64 division rounded once to fp32).  Pinned against the
compiled reference (oracle/_ref) in tests/test_gpu_iou.py."""
import numpy as np


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, mask_scores=None, mode=0):
    n_prop, n_inst = len(proposals_offset) - 1, len(instance_pointnum)
    iou = np.zeros((n_prop, n_inst), np.float32)
    for p in range(n_prop):
        s, e = int(proposals_offset[p]), int(proposals_offset[p + 1])
        idx = np.asarray(proposals_idx[s:e])
        if mode == 1:
            keep = np.asarray(mask_scores[s:e]).reshape(-1) > np.float32(0.5)
            idx = idx[keep]
        total = len(idx)
        labs = np.asarray(instance_labels)[idx].astype(np.int32)
        labs = labs[(labs >= 0) & (labs < n_inst)]
        inter = np.bincount(labs, minlength=n_inst)
        uni = (total + np.asarray(instance_pointnum, np.int64) - inter).astype(np.float32)
        iou[p] = (inter.astype(np.float32).astype(np.float64) / (uni.astype(np.float64) + 1e-5)).astype(np.float32)
    return iou


def mask_label(proposals_idx, proposals_offset, instance_labels, iou, init):
    out = np.array(init, np.float32).copy().reshape(-1)
    for p in range(len(proposals_offset) - 1):
        max_iou, max_ind = np.float32(0.0), 0
        for k in range(iou.shape[1]):
            if iou[p, k] > max_iou:
                max_iou, max_ind = iou[p, k], k
        if max_iou > 0.5:
            s, e = int(proposals_offset[p]), int(proposals_offset[p + 1])
            out[s:e] = (np.asarray(instance_labels)[proposals_idx[s:e]].astype(np.int32) == max_ind).astype(np.float32)
    return out
