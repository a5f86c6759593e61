// pb_front.cuh — device-side front end of the fused class loop (network/PBNet.py:151-179, 282-294).
//
// The reference walks the 18 foreground classes in Python: nonzero + sort per class, the `count < count_mean*0.05` skip,
// per-copy `.sum()` loops (get_batch_offset) and an fp32 `orig + offset` add on the CPU.  Round 1 restated that with ~15
// eager torch ops (argsort, nonzero, bincount, a .cpu() sync).  Here it is four kernels:
//   k_front_keys    key = class * copies + copy (or "dropped") per point + the global key histogram
//   k_sort_pass     ONE stable 9-bit radix pass of the hand-written sort (pb_sort.cuh) -> points in class-major,
//                   copy-major, ascending-index order
//   k_front_tables  class totals, the skip rule, segment tables (one block)
//   k_front_gather  SoA coordinates of the kept points (shifted = fl32(orig + offset)), class, point index
#pragma once
#include "pb_sort.cuh"

namespace pbf {

constexpr int kDropKey = pb::kBins - 1;   // wall / floor / invalid class: sorts behind every kept key
constexpr int kMaxSem = 20;
enum { kErrBatch = 1 };

__global__ void __launch_bounds__(256)
k_front_keys(long long n, const long long *__restrict__ sem, const void *__restrict__ batch, int batch64, int copies,
             uint32_t *__restrict__ key, unsigned *__restrict__ hist, int *err) {
    __shared__ unsigned sh[pb::kBins];
    for (int i = threadIdx.x; i < pb::kBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long c = sem[i];
        long long b = batch64 ? reinterpret_cast<const long long *>(batch)[i] : (long long)reinterpret_cast<const int *>(batch)[i];
        uint32_t k = kDropKey;
        if (b < 0 || b >= copies) atomicOr(err, kErrBatch);   // network/PBNet.py:282-287 asserts the per-copy counts add up
        else if (c >= 2 && c < kMaxSem) k = (uint32_t)(c * copies + b);
        key[i] = k;
        atomicAdd(sh + k, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pb::kBins; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, sh[i]);
}

// tables[]: [0] n_kept  [1] n_classes kept  [2..21] keep flag per class  [32 ..] seg_counts of the kept classes
// (class-major, `copies` entries each)   [32 + 512 ..] source start of every kept segment in the sorted order
// [32 + 1024 ..] output start of every kept segment (+ the total behind the last)   [32 + 1600 ..] its class
constexpr int kTabCnt = 32, kTabSrc = 32 + 512, kTabOut = 32 + 1024, kTabCls = 32 + 1600, kTabSize = 32 + 2200;
__global__ void k_front_tables(const unsigned *__restrict__ hist, int copies, const float *__restrict__ thr20, int *__restrict__ tables) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int src = 0, out = 0, nseg = 0, ncls = 0;
    for (int c = 0; c < kMaxSem; c++) {
        int total = 0;
        for (int b = 0; b < copies; b++) total += (c >= 2) ? (int)hist[c * copies + b] : 0;
        // `ins_ind.shape[0] < self.count_mean[sem_id] * 0.05` -> skip (network/PBNet.py:156): fp32 product, computed on the host
        bool keep = c >= 2 && !((float)total < thr20[c]);
        tables[2 + c] = keep ? 1 : 0;
        for (int b = 0; b < copies; b++) {
            int cnt = (c >= 2) ? (int)hist[c * copies + b] : 0;
            if (keep) {
                tables[kTabCnt + nseg] = cnt;
                tables[kTabSrc + nseg] = src;
                tables[kTabOut + nseg] = out;
                tables[kTabCls + nseg] = c;
                out += cnt;
                nseg++;
            }
            src += cnt;
        }
        ncls += keep ? 1 : 0;
    }
    tables[0] = out;
    tables[1] = ncls;
    tables[kTabOut + nseg] = out;
}

__global__ void __launch_bounds__(256)
k_front_gather(const int *__restrict__ tables, int copies, const uint32_t *__restrict__ sorted_idx,
               const float *__restrict__ xyz, const float *__restrict__ offset, float *__restrict__ x, float *__restrict__ y,
               float *__restrict__ z, float *__restrict__ xo, float *__restrict__ yo, float *__restrict__ zo,
               int *__restrict__ sem32, long long *__restrict__ pidx) {
    const int n_kept = tables[0], nseg = tables[1] * copies;
    const int *src0 = tables + kTabSrc, *out0 = tables + kTabOut, *cls = tables + kTabCls;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_kept; j += gridDim.x * blockDim.x) {
        int lo = 0, hi = nseg - 1;  // last segment whose output start is <= j (empty segments share a start: take the last)
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (out0[mid] <= j) lo = mid;
            else hi = mid - 1;
        }
        uint32_t p = sorted_idx[src0[lo] + (j - out0[lo])];
        const int c = cls[lo];
        float ox = xyz[3 * (size_t)p], oy = xyz[3 * (size_t)p + 1], oz = xyz[3 * (size_t)p + 2];
        xo[j] = ox, yo[j] = oy, zo[j] = oz;
        x[j] = __fadd_rn(ox, offset[3 * (size_t)p]);       // network/PBNet.py:165  ins_orig.cpu() + ins_offset.cpu()
        y[j] = __fadd_rn(oy, offset[3 * (size_t)p + 1]);
        z[j] = __fadd_rn(oz, offset[3 * (size_t)p + 2]);
        sem32[j] = c;
        pidx[j] = (long long)p;
    }
}

}  // namespace pbf
