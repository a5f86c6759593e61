// pb_eval.cuh — evaluation post-processing of the proposals on the device.
//
// Reference: eval_map.py:63-121 builds a dense nProposal x N int matrix, thresholds it, takes the cross IoU with a
// dense fp32 mm, runs greedy NMS on the CPU (tools/mIOU.py:77-87), paints per-point labels in a Python loop, aligns
// them to superpoints through a scipy coo_matrix on the CPU (tools/getins.py:72-98) and rebuilds the dense masks in
// another Python loop.
//
// Here everything stays sparse and on the device: one radix sort of the folded (proposal, point) pairs, pairwise
// intersection counts from the per-point proposal runs, a single-block greedy NMS, and the superpoint vote as a
// sort + run-length + packed atomicMax (most frequent label, first maximum).  All outputs are integers or copied
// floats; the IoU is the same fp32 expression as the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbe {

constexpr int kMaxValid = 4096;  // proposals that survive the thresholds (the NMS block keeps them in shared state)
enum { kErrProp = 1, kErrPoint = 2, kErrSem = 4, kErrSuper = 8 };
constexpr uint64_t kNoKey = ~0ull;

template <class T>
__device__ __forceinline__ long long upper_bound(const T *__restrict__ a, long long lo, long long hi, T v) {
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// (proposal, folded point) keys; class of every proposal (first listed point BEFORE the fold, eval_map.py:64-66)
__global__ void k_pair_keys(long long M, int P, long long N, int n3, const long long *__restrict__ pidx,
                            const long long *__restrict__ poff, const long long *__restrict__ pred_sem,
                            const long long *__restrict__ sem_table, int n_table, uint64_t *__restrict__ key,
                            long long *__restrict__ prop_sem, int *__restrict__ err) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < P) {
        long long o = poff[e];
        long long v = -1;
        if (o >= 0 && o < M) {
            long long pt = pidx[2 * o + 1];
            if (pt >= 0 && pt < N) {
                long long c = pred_sem[pt];
                if (c >= 0 && c < n_table) v = sem_table[c];
                else atomicOr(err, kErrSem);
            }
        }
        prop_sem[e] = v;
    }
    if (e >= M) return;
    long long p = pidx[2 * e], pt = pidx[2 * e + 1];
    if (p < 0 || p >= P) {
        atomicOr(err, kErrProp);
        key[e] = kNoKey;
        return;
    }
    if (pt < 0 || pt >= N) {
        atomicOr(err, kErrPoint);
        key[e] = kNoKey;
        return;
    }
    key[e] = (uint64_t)p * (uint64_t)n3 + (uint64_t)(pt % n3);  // eval_map.py:68
}

// heads of the sorted keys = distinct (proposal, point) pairs = ones of proposals_pred; point counts per proposal
__global__ void k_pair_heads(long long M, int n3, const uint64_t *__restrict__ skey, int *__restrict__ head,
                             int *__restrict__ npoint) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    uint64_t k = skey[j];
    int h = (k != kNoKey) && (j == 0 || skey[j - 1] != k);
    head[j] = h;
    if (h) atomicAdd(npoint + (int)(k / (uint64_t)n3), 1);
}

__global__ void k_valid(int P, const float *__restrict__ score, float score_thr, const int *__restrict__ npoint, int npoint_thr,
                        int *__restrict__ valid) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) valid[p] = (score[p] > score_thr) && (npoint[p] > npoint_thr);  // eval_map.py:75,81
}

// distinct pairs of valid proposals, keyed by (point, compact proposal index)
__global__ void k_point_keys(long long M, int n3, const uint64_t *__restrict__ skey, const int *__restrict__ head,
                             const int *__restrict__ valid, const int *__restrict__ vid, const int *__restrict__ d_V,
                             uint64_t *__restrict__ key2) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    uint64_t out = kNoKey;
    if (head[j]) {
        uint64_t k = skey[j];
        int p = (int)(k / (uint64_t)n3), pt = (int)(k % (uint64_t)n3);
        if (valid[p]) out = (uint64_t)pt * (uint64_t)(*d_V) + (uint64_t)vid[p];
    }
    key2[j] = out;
}

// every point's run of proposals: intersection counts of all pairs inside the run (symmetric, diagonal = point count)
__global__ void k_intersections(long long M, const uint64_t *__restrict__ s2, const int *__restrict__ d_V, int *__restrict__ inter) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    uint64_t k = s2[j];
    if (k == kNoKey) return;
    const uint64_t V = (uint64_t)(*d_V);
    const uint64_t pt = k / V;
    const int a = (int)(k % V);
    atomicAdd(inter + (long long)a * V + a, 1);
    for (long long q = j + 1; q < M; q++) {
        uint64_t k2 = s2[q];
        if (k2 == kNoKey || k2 / V != pt) break;
        int b = (int)(k2 % V);
        atomicAdd(inter + (long long)a * V + b, 1);
        atomicAdd(inter + (long long)b * V + a, 1);
    }
}

// greedy NMS (tools/mIOU.py:77-87) in one block: descending score (ties: higher index first, as argsort()[::-1] of a
// stable sort), the picked proposal removes every later one with iou > thr; iou = inter / (n_i + n_j - inter) in fp32
__global__ void __launch_bounds__(1024)
k_nms(const int *__restrict__ d_V, const int *__restrict__ vlist, const float *__restrict__ score, const int *__restrict__ inter,
      float thr, int *__restrict__ order, int *__restrict__ pick_rank, int *__restrict__ picked, int *__restrict__ d_C) {
    __shared__ unsigned char dead[kMaxValid];
    __shared__ int s_cur, s_np;
    const int V = *d_V;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        float si = score[vlist[i]];
        int r = 0;
        for (int j = 0; j < V; j++) {
            float sj = score[vlist[j]];
            r += (sj > si) || (sj == si && j > i);
        }
        order[r] = i;
        dead[i] = 0;
        pick_rank[i] = -1;
    }
    if (threadIdx.x == 0) s_np = 0;
    __syncthreads();
    for (int r = 0; r < V; r++) {
        if (threadIdx.x == 0) {
            int i = order[r];
            s_cur = dead[i] ? -1 : i;
            if (!dead[i]) {
                pick_rank[i] = s_np;
                picked[s_np] = vlist[i];
                s_np++;
            }
        }
        __syncthreads();
        const int i = s_cur;
        if (i >= 0) {
            const float ni = (float)inter[(long long)i * V + i];
            for (int q = r + 1 + threadIdx.x; q < V; q += blockDim.x) {
                int j = order[q];
                if (dead[j]) continue;
                float it = (float)inter[(long long)i * V + j];
                float nj = (float)inter[(long long)j * V + j];
                float iou = __fdiv_rn(it, __fsub_rn(__fadd_rn(ni, nj), it));  // eval_map.py:96
                if (iou > thr) dead[j] = 1;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_C = s_np;
}

// per-point label = the LAST picked cluster containing the point (eval_map.py:105-109)
__global__ void k_paint(long long M, const uint64_t *__restrict__ s2, const int *__restrict__ d_V, const int *__restrict__ pick_rank,
                        int *__restrict__ label) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    uint64_t k = s2[j];
    if (k == kNoKey) return;
    const uint64_t V = (uint64_t)(*d_V);
    int c = pick_rank[(int)(k % V)];
    if (c >= 0) atomicMax(label + (int)(k / V), c);
}

// superpoint vote (tools/getins.py:87-93): keys (superpoint, label column); unlabelled points vote for column C
__global__ void k_vote_keys(int n3, const long long *__restrict__ superpoint, int n_sp, const int *__restrict__ label,
                            const int *__restrict__ d_C, uint64_t *__restrict__ key, int *__restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    long long sp = superpoint[i];
    if (sp < 0 || sp >= n_sp) {
        atomicOr(err, kErrSuper);
        key[i] = kNoKey;
        return;
    }
    int C = *d_C;
    int col = label[i] < 0 ? C : label[i];
    key[i] = (uint64_t)sp * (uint64_t)(C + 1) + (uint64_t)col;
}

__global__ void k_vote_count(int n3, const uint64_t *__restrict__ skey, const int *__restrict__ d_C,
                             unsigned long long *__restrict__ best) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n3) return;
    uint64_t k = skey[j];
    if (k == kNoKey || (j > 0 && skey[j - 1] == k)) return;
    int end = (int)upper_bound(skey, (long long)j, (long long)n3, k);
    const uint64_t C1 = (uint64_t)(*d_C + 1);
    unsigned col = (unsigned)(k % C1);
    // np.argmax: largest count, first (lowest) column among equals
    atomicMax(best + (k / C1), ((unsigned long long)(unsigned)(end - j) << 32) | (0xffffffffu - col));
}

// aligned label of every point; which clusters are still alive (eval_map.py:110-118)
__global__ void k_align(int n3, const long long *__restrict__ superpoint, const unsigned long long *__restrict__ best,
                        const int *__restrict__ d_C, int *__restrict__ label, int *__restrict__ alive) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    unsigned long long b = best[superpoint[i]];
    int col = (int)(0xffffffffu - (unsigned)b);
    int l = (col == *d_C) ? -100 : col;
    label[i] = l;
    if (l >= 0) alive[l] = 1;
}

__global__ void k_finish(int n3, const int *__restrict__ d_C, const int *__restrict__ alive, const int *__restrict__ newid,
                         const int *__restrict__ picked, const float *__restrict__ score, const long long *__restrict__ prop_sem,
                         int *__restrict__ label, float *__restrict__ out_score, long long *__restrict__ out_sem,
                         int *__restrict__ out_picked) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *d_C && alive[i]) {
        int p = picked[i], o = newid[i];
        out_score[o] = score[p];
        out_sem[o] = prop_sem[p];
        out_picked[o] = p;
    }
    if (i < n3 && label[i] >= 0) label[i] = newid[label[i]];
}

__global__ void k_valid_list(int P, const int *__restrict__ valid, const int *__restrict__ vid, int *__restrict__ vlist) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P && valid[p]) vlist[vid[p]] = p;
}

__global__ void k_fill_i32(int n, int v, int *__restrict__ a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace pbe
