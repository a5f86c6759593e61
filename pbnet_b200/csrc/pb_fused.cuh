// pb_fused.cuh — the per-point bookkeeping stages of the grouping path as a few fat, segment-tiled kernels
// (round 1 ran them as ~25 tiny launches: the fixed per-call latency of the drop-in path).
//
//   k_prep        validation, segment ids, per-segment bounding boxes            (1 block-level atomic set per tile)
//   k_seg_params  per-segment radius / cell edge / origin + the global cell-extent reduction that trims the sort keys
//   k_keys        segment-local sort keys of both sorts + their per-(segment, pass) digit histograms
//   k_grid_build  after the cell sort: gather into cell order, cell / coarse-cell / row head flags, their THREE
//                 exclusive scans (one chained look-back), and every cell table — one kernel, one pass over the points
//   k_filter_scan fragment filter + compaction scan of the raw clusters + per-segment cluster counts   (one block)
//   k_relabel_scan final ids, LP-query list and labelled list (two scans + both compactions) in one pass
//   k_lab_boxes   both levels of the box hierarchy of the labelled list
// Every body is a __device__ function of a TILE index so that the small-problem kernel (pb_small.cuh) can run the same
// code as phases of ONE cooperative launch.
#pragma once
#include "pb_kernels.cuh"
#include "pb_sort.cuh"

namespace pb {

// Layout of the trimmed cell key: fine bits (3) | coarse x (bx) | coarse y (by) | coarse z (bz), LSB first.
struct KeyLayout {
    int bx, by, bz;   // occupied bits of the coarse cell coordinates (maximum over the segments of the call)
    int bits;         // 3 + bx + by + bz
};
__host__ __device__ inline int bit_width_i(int v) {
    int b = 0;
    while (v > 0) b++, v >>= 1;
    return b < 1 ? 1 : b;
}
__host__ __device__ inline KeyLayout make_key_layout(int ex, int ey, int ez) {  // e* = largest coarse cell index per axis
    KeyLayout k;
    k.bx = bit_width_i(ex), k.by = bit_width_i(ey), k.bz = bit_width_i(ez);
    k.bits = 3 + k.bx + k.by + k.bz;
    return k;
}
// canonical 64-bit cell key of the cell tables (fixed 13-bit fields, segment on top) from the trimmed sort key
__device__ __forceinline__ uint64_t expand_key(uint64_t k, const KeyLayout &L, int seg) {
    uint64_t cx = (k >> 3) & ((1ull << L.bx) - 1ull);
    uint64_t cy = (k >> (3 + L.bx)) & ((1ull << L.by) - 1ull);
    uint64_t cz = (k >> (3 + L.bx + L.by));
    return ((uint64_t)seg << kSegShift) | (cz << (3 + 2 * kCoarseBits)) | (cy << (3 + kCoarseBits)) | (cx << 3) | (k & 7ull);
}

// ------------------------------------------------------------------------------------------------------------------
// K1  prep: one tile = up to kTB*ITEMS points of ONE segment
// ------------------------------------------------------------------------------------------------------------------
template <int ITEMS>
__device__ __forceinline__ void prep_tile(const TileTab &tt, int t, SegArrays sg, const float *__restrict__ x,
                                          const float *__restrict__ y, const float *__restrict__ z,
                                          const float *__restrict__ xo, const float *__restrict__ yo,
                                          const float *__restrict__ zo, const int *__restrict__ sem,
                                          int *__restrict__ seg_of, int *err, unsigned (*sh)[kTB / 32]) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int begin = tt.begin[t], count = tt.count[t], seg = tt.seg[t];
    const int cls0 = __ldg(sem + sg.start[seg]);  // class of the segment = class of its first point
    unsigned mn[6], mx[6];
#pragma unroll
    for (int k = 0; k < 6; k++) mn[k] = 0xffffffffu, mx[k] = 0u;
    int e = 0;
    for (int idx = tid; idx < count; idx += kTB) {
        int i = begin + idx;
        float v[6] = {x[i], y[i], z[i], xo[i], yo[i], zo[i]};
        int c = sem[i];
        seg_of[i] = seg;
        if (c < 2 || c > 19) e |= kErrSem;
        else if (c != cls0) e |= kErrMixed;
        bool fin = true;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            fin &= isfinite(v[k]);
            unsigned en = enc_f(v[k]);
            mn[k] = min(mn[k], en);
            mx[k] = max(mx[k], en);
        }
        if (!fin) e |= kErrNonFinite;
    }
    if (e) atomicOr(err, e);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        mn[k] = __reduce_min_sync(kFull, mn[k]);
        mx[k] = __reduce_max_sync(kFull, mx[k]);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) sh[k][wid] = mn[k], sh[6 + k][wid] = mx[k];
    }
    __syncthreads();
    if (tid < 12) {
        unsigned v = sh[tid][0];
        for (int w2 = 1; w2 < kTB / 32; w2++) v = tid < 6 ? min(v, sh[tid][w2]) : max(v, sh[tid][w2]);
        if (tid < 3) atomicMin(sg.enc_min_s + 3 * seg + tid, v);
        else if (tid < 6) atomicMin(sg.enc_min_o + 3 * seg + (tid - 3), v);
        else if (tid < 9) atomicMax(sg.enc_max_s + 3 * seg + (tid - 6), v);
        else atomicMax(sg.enc_max_o + 3 * seg + (tid - 9), v);
    }
    __syncthreads();
}

template <int ITEMS>
__global__ void __launch_bounds__(kTB)
k_prep(TileTab tt, SegArrays sg, const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
       const float *__restrict__ xo, const float *__restrict__ yo, const float *__restrict__ zo,
       const int *__restrict__ sem, int *__restrict__ seg_of, int *err) {
    __shared__ unsigned sh[12][kTB / 32];
    for (int t = blockIdx.x; t < tt.T; t += gridDim.x) prep_tile<ITEMS>(tt, t, sg, x, y, z, xo, yo, zo, sem, seg_of, err, sh);
}

// ------------------------------------------------------------------------------------------------------------------
// K2  per segment: radius / cell edge / origins (binary_cuda_functions.cu:85: r2 = fl(r*r)); largest coarse cell index
//     per axis over all segments -> d_ext[3] (the trimmed key layout); range check of the cell grid
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void seg_params_one(int s, SegArrays sg, const int *__restrict__ sem,
                                               const float *__restrict__ radius_tab, const int *__restrict__ min_pts_tab,
                                               int *d_ext, int *err) {
    int b = sg.start[s], e = sg.start[s + 1];
    int c = 2;
    if (e > b) c = sem[b];
    if (c < 2 || c > 19) c = 2;  // flagged by k_prep
    sg.cls[s] = c;
    sg.min_pts[s] = min_pts_tab[c - 2];
    float r = radius_tab[c - 2];
    sg.r2[s] = __fmul_rn(r, r);
    // fine cell edge h = r/2 * (1 + 2^-7): h*sqrt(3) < r (one fine cell = clique) and the coarse cell (2h >= r) makes
    // the 3^3 coarse stencil a superset of the r-ball, with margins far above the fp32 rounding of the cell coordinate
    float h = r * 0.5f * (1.0f + 1.0f / 128.0f);
    if (!(h > 0.f)) h = 1e-6f;
    float ih = 1.0f / h;
    sg.inv_h[s] = ih;
    float ext = 0.f;
    for (int k = 0; k < 3; k++) {
        float mn = e > b ? dec_f(sg.enc_min_s[3 * s + k]) : 0.f;
        float mxs = e > b ? dec_f(sg.enc_max_s[3 * s + k]) : 0.f;
        sg.min_s[3 * s + k] = mn;
        float mo = e > b ? dec_f(sg.enc_min_o[3 * s + k]) : 0.f;
        float Mo = e > b ? dec_f(sg.enc_max_o[3 * s + k]) : 0.f;
        sg.min_o[3 * s + k] = mo;
        ext = fmaxf(ext, Mo - mo);
        if (e > b) {
            // the point with the largest coordinate gets the largest cell index (same rounding chain as k_keys, monotone)
            float f = __fmul_rn(__fsub_rn(mxs, mn), ih);
            int cmax = (f >= 0.f && f < 1e9f) ? (int)f : (f >= 1e9f ? 0x7fffffff : 0);
            if (cmax > kCellMax) {
                atomicOr(err, kErrRange);
                cmax = kCellMax;
            }
            atomicMax(d_ext + k, cmax >> 1);
        }
    }
    // LP-assignment sort grid (ordering only, never a correctness filter): 512 cells along the longest axis of the
    // segment's box -> 27 Morton bits
    float g = fmaxf(ext / (float)kMortonMax, 1e-6f);
    sg.inv_g[s] = 1.0f / g;
}

__global__ void k_seg_params(int S, SegArrays sg, const int *__restrict__ sem, const float *__restrict__ radius_tab,
                             const int *__restrict__ min_pts_tab, int *d_ext, int *err) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) seg_params_one(s, sg, sem, radius_tab, min_pts_tab, d_ext, err);
}

// ------------------------------------------------------------------------------------------------------------------
// K3  keys of both sorts + digit histograms of the multi-tile segments.
//     key1 = trimmed cell key (shifted space), key2 = Morton code of the original coordinates (MIXED: class on top)
// ------------------------------------------------------------------------------------------------------------------
struct KeysArgs {
    KeyLayout lay;
    PassPlan plan1, plan2;
    int key64;              // key1 stored as uint64_t (more than 32 occupied bits)
    unsigned *hist;         // [slots][hist_stride]; pass p of sort 1 at p*kBins, of sort 2 at (plan1.npass + p)*kBins
    int hist_stride;
};

template <bool MIXED, int ITEMS>
__device__ __forceinline__ void keys_tile(const TileTab &tt, int t, SegArrays sg, const KeysArgs &ka,
                                          const float *__restrict__ x, const float *__restrict__ y,
                                          const float *__restrict__ z, const float *__restrict__ xo,
                                          const float *__restrict__ yo, const float *__restrict__ zo,
                                          const int *__restrict__ sem, void *__restrict__ key1, uint32_t *__restrict__ key2,
                                          int *err, const float *__restrict__ radius_tab, int *__restrict__ cnt18,
                                          unsigned *shist) {
    const int tid = threadIdx.x;
    const int begin = tt.begin[t], count = tt.count[t], s = tt.seg[t], hslot = tt.hslot[t];
    const int nh = (ka.plan1.npass + ka.plan2.npass) * kBins;
    if (hslot >= 0) {
        for (int i = tid; i < nh; i += kTB) shist[i] = 0u;
        __syncthreads();
    }
    const float ih = sg.inv_h[s], ig = sg.inv_g[s];
    const float m0 = sg.min_s[3 * s], m1 = sg.min_s[3 * s + 1], m2 = sg.min_s[3 * s + 2];
    const float o0 = sg.min_o[3 * s], o1 = sg.min_o[3 * s + 1], o2 = sg.min_o[3 * s + 2];
    for (int idx = tid; idx < count; idx += kTB) {
        int i = begin + idx;
        float fx = __fmul_rn(__fsub_rn(x[i], m0), ih);
        float fy = __fmul_rn(__fsub_rn(y[i], m1), ih);
        float fz = __fmul_rn(__fsub_rn(z[i], m2), ih);
        int cx = (fx >= 0.f && fx < 1e9f) ? (int)fx : 0;
        int cy = (fy >= 0.f && fy < 1e9f) ? (int)fy : 0;
        int cz = (fz >= 0.f && fz < 1e9f) ? (int)fz : 0;
        cx = min(cx, kCellMax), cy = min(cy, kCellMax), cz = min(cz, kCellMax);  // out-of-range segments were flagged by k_seg_params
        uint64_t k1 = ((uint64_t)(cz >> 1) << (3 + ka.lay.bx + ka.lay.by)) | ((uint64_t)(cy >> 1) << (3 + ka.lay.bx)) |
                      ((uint64_t)(cx >> 1) << 3) | (uint64_t)(((cz & 1) << 2) | ((cy & 1) << 1) | (cx & 1));
        float gx = (xo[i] - o0) * ig, gy = (yo[i] - o1) * ig, gz = (zo[i] - o2) * ig;
        uint32_t mx = (uint32_t)min((gx >= 0.f && gx < 1e9f) ? (int)gx : 0, kMortonMax);
        uint32_t my = (uint32_t)min((gy >= 0.f && gy < 1e9f) ? (int)gy : 0, kMortonMax);
        uint32_t mz = (uint32_t)min((gz >= 0.f && gz < 1e9f) ? (int)gz : 0, kMortonMax);
        uint32_t k2 = (uint32_t)(spread3(mx) | (spread3(my) << 1) | (spread3(mz) << 2));
        if (MIXED) {
            // the reference looks the radius up with a sorted-position index (binary_cuda_functions.cu:35,110): only well
            // defined when all classes of a segment share one radius
            int myc = min(max(sem[i], 2), 19);
            if (radius_tab[myc - 2] != radius_tab[sg.cls[s] - 2]) atomicOr(err, kErrRadius);
            atomicAdd(cnt18 + (long long)s * kCls + (myc - 2), 1);
            k2 |= (uint32_t)(myc - 2) << kKey2SegShift;   // class-major inside the segment
        }
        if (ka.key64) reinterpret_cast<uint64_t *>(key1)[i] = k1;
        else reinterpret_cast<uint32_t *>(key1)[i] = (uint32_t)k1;
        key2[i] = k2;
        if (hslot >= 0) {
#pragma unroll
            for (int p = 0; p < kMaxPasses; p++)
                if (p < ka.plan1.npass)
                    atomicAdd(shist + p * kBins + ((unsigned)(k1 >> ka.plan1.shift[p]) & ((1u << ka.plan1.width[p]) - 1u)), 1u);
#pragma unroll
            for (int p = 0; p < kMaxPasses; p++)
                if (p < ka.plan2.npass)
                    atomicAdd(shist + (ka.plan1.npass + p) * kBins + ((k2 >> ka.plan2.shift[p]) & ((1u << ka.plan2.width[p]) - 1u)), 1u);
        }
    }
    if (hslot >= 0) {
        __syncthreads();
        unsigned *h = ka.hist + (size_t)hslot * ka.hist_stride;
        for (int i = tid; i < nh; i += kTB) {
            unsigned v = shist[i];
            if (v) atomicAdd(h + i, v);
        }
        __syncthreads();
    }
}

template <bool MIXED, int ITEMS>
__global__ void __launch_bounds__(kTB)
k_keys(TileTab tt, SegArrays sg, KeysArgs ka, const float *__restrict__ x, const float *__restrict__ y,
       const float *__restrict__ z, const float *__restrict__ xo, const float *__restrict__ yo,
       const float *__restrict__ zo, const int *__restrict__ sem, void *__restrict__ key1, uint32_t *__restrict__ key2,
       int *err, const float *__restrict__ radius_tab, int *__restrict__ cnt18) {
    __shared__ unsigned shist[kMaxPassesGroup * kBins];  // 18 KB
    for (int t = blockIdx.x; t < tt.T; t += gridDim.x)
        keys_tile<MIXED, ITEMS>(tt, t, sg, ka, x, y, z, xo, yo, zo, sem, key1, key2, err, radius_tab, cnt18, shist);
}

// ------------------------------------------------------------------------------------------------------------------
// K4  grid build.  Warp w of a tile owns the contiguous chunk [w*32*ITEMS, (w+1)*32*ITEMS): item k of lane l is chunk
//     element 32k + l (coalesced), head-flag prefixes come from ballots.  The three running ordinals (fine cell, coarse
//     cell, coarse row) are global across segments: tile aggregates are chained with a decoupled look-back on two
//     64-bit words {flag | fine | coarse} and {flag | rows}.
// ------------------------------------------------------------------------------------------------------------------
struct GridOut {
    float4 *pts4;
    float *sx, *sy, *sz;
    int *fcell_of, *row_of, *fcell_start, *fcell_cc, *cc_pstart, *cc_fstart, *parent, *cell_hp, *cell_minhp, *comp_min, *cell_first;
    uint64_t *fcell_key, *cc_key;
    int *d_F, *d_Cc, *d_rows;
    unsigned long long *stateA, *stateB;   // [T] look-back words, zero-initialised
    int *ticket;
};

struct GridSmem {
    int wF[kTB / 32], wC[kTB / 32], wR[kTB / 32];
    int preF, preC, preR;
    int tile;
};

// chained exclusive prefix of one packed 64-bit aggregate (values in the low 62 bits add without carry into the flag)
__device__ __forceinline__ unsigned long long lookback64(unsigned long long *state, int t, unsigned long long agg, int lane) {
    unsigned long long prefix = 0ull;
    if (t == 0) {
        if (lane == 0) st_volatile_u64(state, (2ull << 62) | agg);
        return 0ull;
    }
    if (lane == 0) st_volatile_u64(state + t, (1ull << 62) | agg);
    int j = t - 1;  // lane l inspects tile j - l; tiles before 0 count as a published prefix of 0
    while (true) {
        int idx = j - lane;
        unsigned long long wv = idx >= 0 ? ld_volatile_u64(state + idx) : (2ull << 62);
        unsigned flag = (unsigned)(wv >> 62);
        unsigned not_ready = ~__ballot_sync(kFull, flag != 0u);
        int usable = not_ready ? __ffs(not_ready) - 1 : 32;
        unsigned is_prefix = __ballot_sync(kFull, flag == 2u) & (usable == 32 ? kFull : ((1u << usable) - 1u));
        int stop = is_prefix ? __ffs(is_prefix) - 1 : usable - 1;
        unsigned long long val = (lane <= stop) ? (wv & 0x3fffffffffffffffull) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(kFull, val, o);
        prefix += val;
        if (is_prefix) break;
        j -= usable;
    }
    if (lane == 0) st_volatile_u64(state + t, (2ull << 62) | (prefix + agg));
    return prefix;
}

template <typename KeyT, int ITEMS>
__device__ __forceinline__ void grid_build_tile(const TileTab &tt, int t, int n, SegArrays sg, KeyLayout lay,
                                                const KeyT *__restrict__ skey, const uint32_t *__restrict__ order,
                                                const float *__restrict__ x, const float *__restrict__ y,
                                                const float *__restrict__ z, const GridOut &g, GridSmem &s) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int begin = tt.begin[t], count = tt.count[t], seg = tt.seg[t];
    const bool seg_first_tile = tt.first[t] == t;
    const int seg_end = sg.start[seg + 1];
    const int rowshift = 3 + lay.bx;
    // ---- head flags of my items (bit k of three words) and their exclusive ballot prefixes inside the warp's chunk,
    //      three 10-bit counters per word (a chunk holds at most 512 heads)
    unsigned hf = 0, hc = 0, hr = 0;
    unsigned ex[ITEMS];
    int runF = 0, runC = 0, runR = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        bool f = false, c = false, r = false;
        if (idx < count) {
            int i = begin + idx;
            if (idx == 0 && seg_first_tile) f = c = r = true;  // first point of a segment opens everything
            else {
                KeyT kk = skey[i], p = skey[i - 1];
                f = kk != p;
                c = (kk >> 3) != (p >> 3);
                r = (kk >> rowshift) != (p >> rowshift);
            }
        }
        unsigned mf = __ballot_sync(kFull, f), mc = __ballot_sync(kFull, c), mr = __ballot_sync(kFull, r);
        unsigned lt = (1u << lane) - 1u;
        ex[k] = (unsigned)(runF + __popc(mf & lt)) | ((unsigned)(runC + __popc(mc & lt)) << 10) | ((unsigned)(runR + __popc(mr & lt)) << 20);
        runF += __popc(mf), runC += __popc(mc), runR += __popc(mr);
        hf |= (unsigned)f << k, hc |= (unsigned)c << k, hr |= (unsigned)r << k;
    }
    if (lane == 0) s.wF[w] = runF, s.wC[w] = runC, s.wR[w] = runR;
    __syncthreads();
    if (w == 0) {
        int tF = 0, tC = 0, tR = 0;
#pragma unroll
        for (int ww = 0; ww < kTB / 32; ww++) tF += s.wF[ww], tC += s.wC[ww], tR += s.wR[ww];
        unsigned long long pa = lookback64(g.stateA, t, ((unsigned long long)tF << 31) | (unsigned long long)tC, lane);
        unsigned long long pb = lookback64(g.stateB, t, (unsigned long long)tR, lane);
        if (lane == 0) {
            s.preF = (int)(pa >> 31), s.preC = (int)(pa & 0x7fffffffull), s.preR = (int)pb;
        }
    }
    __syncthreads();
    int baseF = s.preF, baseC = s.preC, baseR = s.preR;
    for (int ww = 0; ww < w; ww++) baseF += s.wF[ww], baseC += s.wC[ww], baseR += s.wR[ww];
    // ---- outputs
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        if (idx >= count) continue;
        int i = begin + idx;
        int f1 = (hf >> k) & 1, c1 = (hc >> k) & 1, r1 = (hr >> k) & 1;
        int f = baseF + (int)(ex[k] & 1023u) + f1 - 1, c = baseC + (int)((ex[k] >> 10) & 1023u) + c1 - 1,
            r = baseR + (int)(ex[k] >> 20) + r1 - 1;
        uint32_t o = order[i];
        float vx = x[o], vy = y[o], vz = z[o];
        g.pts4[i] = make_float4(vx, vy, vz, __int_as_float((int)o));
        g.sx[i] = vx, g.sy[i] = vy, g.sz[i] = vz;
        g.fcell_of[i] = f;
        g.row_of[i] = r;
        if (f1) {
            uint64_t ck = expand_key((uint64_t)skey[i], lay, seg);
            g.fcell_start[f] = i;
            g.fcell_key[f] = ck;
            g.fcell_cc[f] = c;
            g.parent[f] = f;
            g.cell_hp[f] = 0;
            g.cell_minhp[f] = 0x7fffffff;
            g.comp_min[f] = 0x7fffffff;
            g.cell_first[f] = 0x7fffffff;
            if (c1) {
                g.cc_pstart[c] = i;
                g.cc_fstart[c] = f;
                g.cc_key[c] = ck >> 3;
            }
        }
        if (idx == 0 && seg_first_tile) sg.cc_start[seg] = c;
        if (i == seg_end - 1) sg.cc_end[seg] = c + 1;
        if (i == n - 1) {
            g.fcell_start[f + 1] = n;
            g.cc_pstart[c + 1] = n;
            g.cc_fstart[c + 1] = f + 1;
            *g.d_F = f + 1;
            *g.d_Cc = c + 1;
            *g.d_rows = r + 1;
            g.sx[n] = g.sx[n + 1] = g.sy[n] = g.sy[n + 1] = g.sz[n] = g.sz[n + 1] = 0.f;  // pad for 64-bit pair loads
        }
    }
    __syncthreads();
}

#ifndef PB_GRID_MINB
#define PB_GRID_MINB 3
#endif
template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kTB, PB_GRID_MINB)
k_grid_build(TileTab tt, int n, SegArrays sg, KeyLayout lay, const KeyT *__restrict__ skey, const uint32_t *__restrict__ order,
             const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z, GridOut g) {
    __shared__ GridSmem s;
    if (threadIdx.x == 0) s.tile = atomicAdd(g.ticket, 1);
    __syncthreads();
    const int t = s.tile;
    if (t >= tt.T) return;
    grid_build_tile<KeyT, ITEMS>(tt, t, n, sg, lay, skey, order, x, y, z, g, s);
}

// ------------------------------------------------------------------------------------------------------------------
// K15-16  fragment filter (binary.cu:219-268: drop raw cluster g iff float(size) < mean_count*para_f), compaction scan of
//         the kept clusters, per-segment cluster counts / id bases.  ONE block: R (raw clusters) is small.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_excl_scan_any(int v, int *smem, int &total) {  // blockDim.x <= 1024, multiple of 32
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int wv = lane < nw ? smem[lane] : 0, winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) smem[lane] = winc - wv;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = inc - v + smem[wid];
    total = smem[32];
    __syncthreads();
    return res;
}

template <bool MIXED>
__device__ __forceinline__ void filter_scan_block(int n, int S, SegArrays sg, const int *__restrict__ d_R,
                                                  const int *__restrict__ rep, const int *__restrict__ seg_of,
                                                  const int *__restrict__ raw_count, const float *__restrict__ thresh18,
                                                  int *__restrict__ keep, int *__restrict__ kscan, int *d_K,
                                                  const int *__restrict__ sem, const int *__restrict__ seg_call_first,
                                                  const int *__restrict__ gid_at, int *__restrict__ cluster_num_out, int *smem) {
    const int R = *d_R;
    int carry = 0;
    for (int base = 0; base < R; base += blockDim.x) {
        int gi = base + threadIdx.x;
        int k = 0;
        if (gi < R) {
            int u = rep[gi];
            // class of a cluster = class of its highest-index member (binary.cu:245); all members share it
            float t = thresh18[(MIXED ? sem[u] : sg.cls[seg_of[u]]) - 2];
            k = ((float)raw_count[gi] < t) ? 0 : 1;
            keep[gi] = k;
        }
        int total;
        int ex = block_excl_scan_any(k, smem, total);
        if (gi < R) kscan[gi] = carry + ex;
        carry += total;
    }
    const int K = carry;
    if (threadIdx.x == 0) *d_K = K;
    __syncthreads();  // kscan of this block's own writes is visible to the whole block from here on
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        int b = sg.start[s], e = sg.start[s + 1];
        int g0 = b < n ? gid_at[b] : R, g1 = e < n ? gid_at[e] : R;
        int k0 = g0 < R ? kscan[g0] : K, k1 = g1 < R ? kscan[g1] : K;
        sg.k_base[s] = k0;
        sg.cluster_num[s] = k1 - k0;
        cluster_num_out[s] = k1 - k0;
        int f = sg.start[seg_call_first[s]];
        int gf = f < n ? gid_at[f] : R;
        sg.id_base[s] = gf < R ? kscan[gf] : K;
    }
}

template <bool MIXED>
__global__ void __launch_bounds__(1024)
k_filter_scan(int n, int S, SegArrays sg, const int *__restrict__ d_R, const int *__restrict__ rep,
              const int *__restrict__ seg_of, const int *__restrict__ raw_count, const float *__restrict__ thresh18,
              int *__restrict__ keep, int *__restrict__ kscan, int *d_K, const int *__restrict__ sem,
              const int *__restrict__ seg_call_first, const int *__restrict__ gid_at, int *__restrict__ cluster_num_out) {
    __shared__ int smem[33];
    filter_scan_block<MIXED>(n, S, sg, d_R, rep, seg_of, raw_count, thresh18, keep, kscan, d_K, sem, seg_call_first, gid_at,
                             cluster_num_out, smem);
}

// ------------------------------------------------------------------------------------------------------------------
// K17-18  final ids of the HP-stage labels + LP-query list (input order) + labelled list (Morton order) in ONE pass:
//         per tile, point u = begin+idx is relabelled AND sorted position i = begin+idx of the LP-assignment order is
//         classified; both flags are scanned together (one 64-bit look-back word {flag | queries | labelled}) and both
//         compactions are written from registers.
// ------------------------------------------------------------------------------------------------------------------
struct RelabelOut {
    int *cluster_id, *clt_sem, *clt_seg, *qlist, *inv2, *lpos, *seg_lastlab;
    float4 *lab4;
    int *d_Q, *d_L;
    unsigned long long *state;  // [T], zero-initialised
    int *ticket;
};
struct RelabelSmem {
    int wQ[kTB / 32], wL[kTB / 32];
    int preQ, preL;
    int tile;
};

template <bool MIXED, int ITEMS>
__device__ __forceinline__ void relabel_tile(const TileTab &tt, int t, int n, SegArrays sg, const int *__restrict__ raw_label,
                                             const int *__restrict__ keep, const int *__restrict__ kscan, int assign_lp,
                                             const int *__restrict__ rep, const int *__restrict__ sem,
                                             const uint32_t *__restrict__ order2, const float *__restrict__ xo,
                                             const float *__restrict__ yo, const float *__restrict__ zo, const RelabelOut &o,
                                             RelabelSmem &s) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int begin = tt.begin[t], count = tt.count[t], seg = tt.seg[t];
    const int id_base = sg.id_base[seg];
    const bool seg_has_clusters = sg.cluster_num[seg] > 0;
    const int cls = sg.cls[seg];
    unsigned qf = 0, lf = 0;
    int exQ[ITEMS], exL[ITEMS];
    int runQ = 0, runL = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        bool q = false, l = false;
        if (idx < count) {
            int u = begin + idx;
            int gi = raw_label[u];
            int id = -1;
            if (gi >= 0 && keep[gi]) {
                int kk = kscan[gi];
                id = kk - id_base;
                if (rep[gi] == u) {
                    o.clt_sem[kk] = MIXED ? sem[u] : cls;
                    o.clt_seg[kk] = seg;
                }
                if (MIXED) atomicMax(o.seg_lastlab + seg, u);  // fallback target of binary_cuda_functions.cu:287-300
            }
            o.cluster_id[u] = id;
            q = id < 0 && assign_lp && seg_has_clusters;
            if (assign_lp) {
                uint32_t p = order2[u];      // sorted position u of the LP-assignment order holds point p
                int gl = raw_label[p];
                l = gl >= 0 && keep[gl];
                o.inv2[p] = u;
            }
        }
        unsigned mq = __ballot_sync(kFull, q), ml = __ballot_sync(kFull, l);
        unsigned lt = (1u << lane) - 1u;
        exQ[k] = runQ + __popc(mq & lt), exL[k] = runL + __popc(ml & lt);
        runQ += __popc(mq), runL += __popc(ml);
        qf |= (unsigned)q << k, lf |= (unsigned)l << k;
    }
    if (lane == 0) s.wQ[w] = runQ, s.wL[w] = runL;
    __syncthreads();
    if (w == 0) {
        int tQ = 0, tL = 0;
#pragma unroll
        for (int ww = 0; ww < kTB / 32; ww++) tQ += s.wQ[ww], tL += s.wL[ww];
        unsigned long long pa = lookback64(o.state, t, ((unsigned long long)tQ << 31) | (unsigned long long)tL, lane);
        if (lane == 0) {
            s.preQ = (int)(pa >> 31), s.preL = (int)(pa & 0x7fffffffull);
            if (t == tt.T - 1) *o.d_Q = s.preQ + tQ, *o.d_L = s.preL + tL;
        }
    }
    __syncthreads();
    int baseQ = s.preQ, baseL = s.preL;
    for (int ww = 0; ww < w; ww++) baseQ += s.wQ[ww], baseL += s.wL[ww];
    if (assign_lp) {
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            int idx = w * 32 * ITEMS + 32 * k + lane;
            if (idx >= count) continue;
            int u = begin + idx;
            if ((qf >> k) & 1) o.qlist[baseQ + exQ[k]] = u;
            int lp = baseL + exL[k];
            o.lpos[u] = lp;
            if ((lf >> k) & 1) {
                uint32_t p = order2[u];
                o.lab4[lp] = make_float4(xo[p], yo[p], zo[p], __int_as_float((int)p));
            }
        }
    }
    __syncthreads();
}

template <bool MIXED, int ITEMS>
__global__ void __launch_bounds__(kTB)
k_relabel_scan(TileTab tt, int n, SegArrays sg, const int *__restrict__ raw_label, const int *__restrict__ keep,
               const int *__restrict__ kscan, int assign_lp, const int *__restrict__ rep, const int *__restrict__ sem,
               const uint32_t *__restrict__ order2, const float *__restrict__ xo, const float *__restrict__ yo,
               const float *__restrict__ zo, RelabelOut o) {
    __shared__ RelabelSmem s;
    if (threadIdx.x == 0) s.tile = atomicAdd(o.ticket, 1);
    __syncthreads();
    const int t = s.tile;
    if (t >= tt.T) return;
    relabel_tile<MIXED, ITEMS>(tt, t, n, sg, raw_label, keep, kscan, assign_lp, rep, sem, order2, xo, yo, zo, o, s);
}

// ------------------------------------------------------------------------------------------------------------------
// K19  bounding boxes of the labelled list, both levels in one kernel: a block owns 1024 consecutive labelled points =
//      32 level-1 boxes (32 points each, one per warp trip) = one level-2 box
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lab_boxes_block(int g2, int L, const float4 *__restrict__ lab4, float4 *__restrict__ box_lo,
                                                float4 *__restrict__ box_hi, float4 *__restrict__ box2_lo,
                                                float4 *__restrict__ box2_hi, float (*sh)[6]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int G = (L + 31) >> 5;
    for (int b = w; b < 32; b += nw) {
        int gi = g2 * 32 + b;
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        if (gi < G) {
            float4 q = lab4[min(gi * 32 + lane, L - 1)];
            lo[0] = hi[0] = q.x, lo[1] = hi[1] = q.y, lo[2] = hi[2] = q.z;
#pragma unroll
            for (int o = 16; o; o >>= 1)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    lo[k] = fminf(lo[k], __shfl_xor_sync(kFull, lo[k], o));
                    hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFull, hi[k], o));
                }
            if (lane == 0) {
                box_lo[gi] = make_float4(lo[0], lo[1], lo[2], 0.f);
                box_hi[gi] = make_float4(hi[0], hi[1], hi[2], 0.f);
            }
        }
        if (lane == 0)
            for (int k = 0; k < 3; k++) sh[b][k] = lo[k], sh[b][3 + k] = hi[k];
    }
    __syncthreads();
    if (w == 0) {
        float lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) lo[k] = sh[lane][k], hi[k] = sh[lane][3 + k];
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(kFull, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFull, hi[k], o));
            }
        if (lane == 0) {
            box2_lo[g2] = make_float4(lo[0], lo[1], lo[2], 0.f);
            box2_hi[g2] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTB)
k_lab_boxes(const int *__restrict__ d_L, const float4 *__restrict__ lab4, float4 *__restrict__ box_lo,
            float4 *__restrict__ box_hi, float4 *__restrict__ box2_lo, float4 *__restrict__ box2_hi) {
    __shared__ float sh[32][6];
    const int L = *d_L;
    const int G2 = (((L + 31) >> 5) + 31) >> 5;
    for (int g2 = blockIdx.x; g2 < G2; g2 += gridDim.x) lab_boxes_block(g2, L, lab4, box_lo, box_hi, box2_lo, box2_hi, sh);
}

}  // namespace pb
