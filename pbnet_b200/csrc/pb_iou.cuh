// pb_iou.cuh — proposal-vs-instance IoU and mask labels: the two remaining entry points of
// the reference's PB_lib module that PBNet calls (get_iou: network/PBNet.py:410) or ships
// (cal_iou_and_masklabel).  Reference kernels: lib/PB_lib/src/iou/get_iou.cu:12-29 and
// lib/PB_lib/src/cal_iou_and_masklabel/cal_iou_and_masklabel.cu:15-90 — one thread per (proposal, instance)
// re-scanning the proposal's points, O(nProposal * nInstance * len).  Here: ONE pass over the proposal points
// builds the intersection histogram (atomics into the output matrix reinterpreted as int32), a second pass
// turns counts into IoUs with the reference's exact arithmetic:
//     iou = (float)( (double)(float)inter / ( (double)(float)(p_total + i_total - inter) + 1e-5 ) )
// (the literal 1e-5 is a double, so the reference divides in fp64 and rounds once to fp32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbi {

// K-I1  intersection counts; block per proposal (grid-stride), threads over its points
__global__ void k_iou_count(int nInstance, int nProposal, const int *__restrict__ proposals_idx,
                            const int *__restrict__ proposals_offset, const long long *__restrict__ instance_labels,
                            const float *__restrict__ mask_scores, int mode, int *__restrict__ inter,
                            int *__restrict__ proposal_total) {
    for (int p = blockIdx.x; p < nProposal; p += gridDim.x) {
        int start = proposals_offset[p], end = proposals_offset[p + 1];
        int kept = 0;
        for (int i = start + threadIdx.x; i < end; i += blockDim.x) {
            if (mode == 1 && !(mask_scores[i] > 0.5f)) continue;
            kept++;
            int lab = (int)instance_labels[proposals_idx[i]];  // (int) cast as the reference does
            if (lab >= 0 && lab < nInstance) atomicAdd(inter + (long long)p * nInstance + lab, 1);
        }
        if (mode == 1) {
            for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
            if ((threadIdx.x & 31) == 0 && kept) atomicAdd(proposal_total + p, kept);
        } else if (threadIdx.x == 0) {
            proposal_total[p] = end - start;
        }
    }
}

// K-I2  counts -> IoU (in place: the int32 counts live in the float output matrix)
__global__ void k_iou_finish(int nInstance, long long total, const int *__restrict__ instance_pointnum,
                             const int *__restrict__ proposal_total, float *__restrict__ iou) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int p = (int)(t / nInstance), k = (int)(t - (long long)p * nInstance);
        int x = __float_as_int(iou[t]);
        float uni = (float)(proposal_total[p] + instance_pointnum[k] - x);
        iou[t] = (float)((double)(float)x / ((double)uni + 1e-5));
    }
}

// K-I3  mask labels: instance with the maximum IoU (first maximum, strict '>' scan from 0); if it exceeds 0.5
//       every point of the proposal gets 1 / 0, else the labels keep their initial value (-1 = ignored)
__global__ void k_mask_label(int nInstance, int nProposal, const int *__restrict__ proposals_idx,
                             const int *__restrict__ proposals_offset, const long long *__restrict__ instance_labels,
                             const float *__restrict__ iou, float *__restrict__ mask_label) {
    int lane = threadIdx.x & 31;
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < nProposal; p += warps) {
        float best = 0.f;
        int bi = 0;
        for (int k = lane; k < nInstance; k += 32) {
            float v = iou[(long long)p * nInstance + k];
            if (v > best) best = v, bi = k;  // per lane: first maximum among its (ascending) instances
        }
        for (int o = 16; o; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && ov > 0.f && oi < bi)) best = ov, bi = oi;
        }
        if (!(best > 0.5f)) continue;
        int start = proposals_offset[p], end = proposals_offset[p + 1];
        for (int i = start + lane; i < end; i += 32)
            mask_label[i] = ((int)instance_labels[proposals_idx[i]] == bi) ? 1.f : 0.f;
    }
}

}  // namespace pbi
