// pb_scenes.cuh — "local scene" proposal construction and get_proposal on the device.
//
// Reference: network/PBNet.py:180-234 builds, per surviving cluster, the list of its member points and — for
// clusters larger than count_mean[sem]*0.2 — appends the members of its para_k nearest clusters (centre
// distance, torch.cdist + topk) with decreasing weights peak_v[k]; it does so with one torch.nonzero over the
// whole segment per (cluster, neighbour) in a Python loop.  network/PBNet.py:317-346 (get_proposal) thresholds
// the mask scores and renumbers the non-empty proposals, again with Python loops over proposals.
//
// Here: ONE stable radix sort groups the points by global cluster index (member lists in ascending point order),
// one thread per cluster ranks the centres of its segment, and a flat grid-stride kernel writes every list
// entry (binary search over the proposal offsets).  Integer outputs are bit-exact; weights are computed with the
// reference's double-precision expression and rounded to fp32 once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbs {

constexpr int kMaxK = 16;  // upper bound on K_max (the reference uses 6, network/PBNet.py:35)
enum { kErrId = 1, kErrLabel = 2 };

// first index in [lo, hi) with a[i] > v
template <class T>
__device__ __forceinline__ int upper_bound(const T *__restrict__ a, int lo, int hi, T v) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// per segment: copy cluster_num to a scan input (K_s), class, first segment of its call
// per point: global cluster index g = gstart[first segment of the call] + id; size histogram; sort key
__global__ void k_point_keys(int n, int S, const int *__restrict__ seg_start, const int *__restrict__ seg_call_first,
                             const int *__restrict__ gstart, const int *__restrict__ cluster_id, int K,
                             uint32_t *__restrict__ key, uint32_t *__restrict__ val, int *__restrict__ size,
                             int *__restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = upper_bound(seg_start, 0, S + 1, i) - 1;
    int id = cluster_id[i];
    int g = K;
    if (id >= 0) {
        int f = seg_call_first[s];
        g = gstart[f] + id;
        // ids of segment s live in [gstart[s], gstart[s+1])  (ids accumulate over the segments of one call,
        // lib/PB_lib/src/pbnet/cluster.cu:50-51,108)
        if (g < gstart[s] || g >= gstart[s + 1]) {
            atomicOr(err, kErrId);
            g = K;
        } else {
            atomicAdd(size + g, 1);
        }
    }
    key[i] = (uint32_t)g;
    val[i] = (uint32_t)i;
}

__global__ void k_cluster_seg(int S, const int *__restrict__ gstart, int *__restrict__ cseg) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int k = gstart[s]; k < gstart[s + 1]; k++) cseg[k] = s;
}

// training branch: 64-bit keys (g, label) of the clustered points
__global__ void k_label_keys(int n, const uint32_t *__restrict__ gkey, const long long *__restrict__ label, int K,
                             uint64_t *__restrict__ key, int *__restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t g = gkey[i];
    long long l = label[i];
    if (l < -2147483647LL || l > 2147483647LL) {
        atomicOr(err, kErrLabel);
        l = 0;
    }
    uint32_t enc = (uint32_t)(int)l ^ 0x80000000u;  // order-preserving
    key[i] = ((uint64_t)g << 32) | enc;
}

// mode of the labels of every cluster (torch.mode: the SMALLEST of the most frequent values, PBNet.py:206)
__global__ void k_label_mode(int n, const uint64_t *__restrict__ skey, int K, unsigned long long *__restrict__ best) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t k = skey[j];
    uint32_t g = (uint32_t)(k >> 32);
    if ((int)g >= K) return;
    if (j > 0 && skey[j - 1] == k) return;  // not a run head
    int end = upper_bound(skey, j, n, k);
    unsigned long long cand = ((unsigned long long)(unsigned)(end - j) << 32) | (0xffffffffu - (uint32_t)k);
    atomicMax(best + g, cand);
}

__device__ __forceinline__ float peak_weight(int para_k, int k) {  // PBNet.py:199,219: double expression -> fp32
    double v = 0.5 * (double)((para_k + 1) - k) / (double)(para_k + 1);
    return (float)v;
}

// one thread per cluster: local-scene decision, the para_k nearest clusters of the same segment, list length
__global__ void k_cluster_plan(int K, const int *__restrict__ cseg, const int *__restrict__ gstart,
                               const int *__restrict__ seg_sem, const float *__restrict__ center,
                               const int *__restrict__ size, const float *__restrict__ big_thresh20,
                               const int *__restrict__ kmax20, const unsigned long long *__restrict__ best, int train,
                               int *__restrict__ para, int *__restrict__ nb, int *__restrict__ len,
                               int *__restrict__ valid, int *__restrict__ mode) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= K) return;
    int s = cseg[g];
    int g0 = gstart[s], Ks = gstart[s + 1] - g0;
    int sem = seg_sem[s];
    int ok = 1;
    if (train) {
        unsigned long long b = best[g];
        int m = (int)((0xffffffffu - (uint32_t)b) ^ 0x80000000u);
        mode[g] = m;
        ok = (b != 0ull) && (m != -100);   // PBNet.py:207: clusters whose mode label is "ignore" are skipped
    }
    valid[g] = ok;
    int pk = min(Ks - 1, kmax20[sem]);      // PBNet.py:197
    int total = size[g];
    int use = 0;
    if (pk > 0 && (float)size[g] > big_thresh20[sem]) {  // PBNet.py:210 (threshold = fp32(count_mean*0.2))
        use = pk;
        float bd[kMaxK + 1];
        int bi[kMaxK + 1];
        int cnt = 0;
        float cx = center[3 * g], cy = center[3 * g + 1], cz = center[3 * g + 2];
        for (int j = g0; j < g0 + Ks; j++) {
            // torch.cdist on <= 25 centres takes the direct path: sqrt(sum (a-b)^2), accumulated over x, y, z
            float dx = cx - center[3 * j], dy = cy - center[3 * j + 1], dz = cz - center[3 * j + 2];
            float d = sqrtf(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
            // insertion into the (pk+1) smallest by (distance, index): topk(largest=False), ties -> lower index
            int p = cnt;
            while (p > 0 && bd[p - 1] > d) p--;
            if (p <= pk) {
                int last = min(cnt, pk);
                for (int q = last; q > p; q--) bd[q] = bd[q - 1], bi[q] = bi[q - 1];
                bd[p] = d, bi[p] = j;
                if (cnt <= pk) cnt++;
            }
        }
        for (int k = 0; k < pk; k++) {      // rank 0 (the cluster itself) is skipped: knn_idx[c_i, k_i + 1]
            int j = bi[k + 1];
            nb[(long long)g * kMaxK + k] = j;
            total += size[j];
        }
    }
    para[g] = use;
    len[g] = ok ? total : 0;
}

// compaction of the valid clusters into proposals
__global__ void k_proposals(int K, const int *__restrict__ valid, const int *__restrict__ pidx,
                            const int *__restrict__ off_g, const int *__restrict__ d_E, const int *__restrict__ d_P,
                            long long *__restrict__ prop_offsets, int *__restrict__ prop_cluster) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) prop_offsets[*d_P] = *d_E;
    if (g >= K || !valid[g]) return;
    prop_offsets[pidx[g]] = off_g[g];
    prop_cluster[pidx[g]] = g;
}

// every list entry: proposal (binary search), sub-list, member
__global__ void k_fill(const int *__restrict__ d_E, const int *__restrict__ d_P, const long long *__restrict__ prop_offsets,
                       const int *__restrict__ prop_cluster, const int *__restrict__ para, const int *__restrict__ nb,
                       const int *__restrict__ size, const int *__restrict__ mstart, const uint32_t *__restrict__ members,
                       const long long *__restrict__ point_map, const long long *__restrict__ label,
                       const int *__restrict__ mode, long long *__restrict__ out_index, float *__restrict__ out_dpn,
                       int *__restrict__ out_gt, int *__restrict__ out_pid) {
    const int E = *d_E, P = *d_P;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        int p = upper_bound(prop_offsets, 0, P + 1, (long long)e) - 1;
        int g = prop_cluster[p];
        int r = e - (int)prop_offsets[p];
        int src = g, sub = 0;
        const int pk = para[g];
        while (r >= size[src]) {  // walks at most pk sub-lists
            r -= size[src];
            src = nb[(long long)g * kMaxK + sub];
            sub++;
        }
        uint32_t m = members[mstart[src] + r];
        out_index[e] = point_map ? point_map[m] : (long long)m;
        out_dpn[e] = sub == 0 ? 1.0f : peak_weight(pk, sub - 1);
        if (out_pid) out_pid[e] = p;
        if (out_gt) {  // PBNet.py:226-230
            long long l = label[m];
            out_gt[e] = l == -100 ? -1 : (l == (long long)mode[g] ? 1 : 0);
        }
    }
}

// list_feat rows (PBNet.py:195,231): [ point_feat[idx] (C channels) | softmax score of the proposal's class | weight ]
// one warp per entry: the C-channel row is read and written as coalesced 128-B segments
__global__ void k_scene_feat(long long E, int C, int n_cls, const float *__restrict__ feat, const float *__restrict__ score,
                             const long long *__restrict__ index, const int *__restrict__ prop_id,
                             const int *__restrict__ prop_sem, const float *__restrict__ dpn, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < E; e += warps) {
        const long long i = index[e];
        const float *__restrict__ src = feat + i * C;
        float *__restrict__ dst = out + e * (C + 2);
        for (int c = lane; c < C; c += 32) dst[c] = __ldg(src + c);
        if (lane == 0) dst[C] = __ldg(score + i * n_cls + prop_sem[prop_id[e]]);
        if (lane == 1) dst[C + 1] = dpn[e];
    }
}

// ---- get_proposal (PBNet.py:317-346) ------------------------------------------------------------------------
__global__ void k_score_flags(int E, const float *__restrict__ score, float thd, int *__restrict__ flag) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) flag[e] = score[e] > thd ? 1 : 0;
}

// per proposal: kept count (difference of the scanned flags at its bounds); non-empty flag
__global__ void k_prop_counts(int P, int E, const long long *__restrict__ prop_offsets, const int *__restrict__ kpos,
                              const int *__restrict__ d_M, int *__restrict__ nonempty, int *__restrict__ kstart) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    long long a = prop_offsets[p], b = prop_offsets[p + 1];
    int ka = a < E ? kpos[a] : *d_M, kb = b < E ? kpos[b] : *d_M;
    nonempty[p] = kb > ka ? 1 : 0;
    kstart[p] = ka;
}

__global__ void k_prop_write(int P, const int *__restrict__ nonempty, const int *__restrict__ newid,
                             const int *__restrict__ kstart, const int *__restrict__ d_M, const int *__restrict__ d_P2,
                             long long *__restrict__ out_offset, long long *__restrict__ out_ids) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) out_offset[*d_P2] = *d_M;
    if (p >= P || !nonempty[p]) return;
    out_offset[newid[p]] = kstart[p];
    out_ids[newid[p]] = p;
}

__global__ void k_prop_entries(int E, const int *__restrict__ flag, const int *__restrict__ kpos,
                               const long long *__restrict__ prop_offsets, int P, const int *__restrict__ newid,
                               const long long *__restrict__ point_idx, const float *__restrict__ score,
                               long long *__restrict__ out_idx2, float *__restrict__ out_ms) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || !flag[e]) return;
    int p = upper_bound(prop_offsets, 0, P + 1, (long long)e) - 1;
    int k = kpos[e];
    out_idx2[2 * (long long)k] = newid[p];
    out_idx2[2 * (long long)k + 1] = point_idx[e];
    out_ms[k] = score[e];
}

}  // namespace pbs
