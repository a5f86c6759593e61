// pb_kernels.cuh — hand-written sm_100a kernels of the instance-grouping path.
//
// Reference semantics: SURVEY.md Appendix A (normative restatement of
// lib/PB_lib/src/pbnet/{cluster.cu,binary.cu,binary_cuda_functions.cu}).  Nothing here is a
// translation of those kernels: the reference materialises neighbour lists and drives a BFS from the
// host; this file works on a sorted cell grid with a cell-level union-find.
//
// Data layout in HBM (N points of all segments concatenated, S segments):
//   pts4[N]      float4 {x,y,z, bits(orig_index | HP<<31)} in (segment, cell) order  — one 16-B
//                broadcast load per candidate in the pair-test loops
//   cell_*[C]    one entry per occupied grid cell (edge h = r/2*(1+2^-7)); cells of a segment are
//                contiguous and sorted (z,y,x)-major so a stencil row is ONE contiguous point range
//   runs[C*25]   the 25 stencil rows of every cell as cell-ordinal ranges
// Grid-cell edge h < r/sqrt(3): any two points of one cell are neighbours, so HP connectivity is a
// union-find over CELLS, not points (k_union), and border LPs only probe cells whose cluster id could
// still raise their maximum (k_label).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int kCellBits = 14;
constexpr int kCellMax = (1 << kCellBits) - 1;
constexpr int kSegShift = 3 * kCellBits;  // 42
constexpr int kRuns = 25;                 // 5 x 5 stencil rows, each up to 5 cells long
constexpr unsigned kFull = 0xffffffffu;
constexpr int kHpBit = 0x80000000;

enum ErrBit { kErrSem = 1, kErrNonFinite = 2, kErrRange = 4, kErrMixed = 8 };

// The reference's square_dist as nvcc 12.9 compiles it for sm_100a (checked in SASS at all three call
// sites, lib/PB_lib/src/pbnet/binary_cuda_functions.cu:85,160,279,305-308):
//   D = fma(dz,dz, fma(dx,dx, fl(dy*dy))).  Intrinsics pin the rounding and forbid re-association.
__device__ __forceinline__ float sqd(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// order-preserving float <-> uint encoding for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned enc_f(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned e) {
    unsigned b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(b);
}

struct SegArrays {
    const int *start;     // [S+1] first point of each segment (host-computed prefix sum)
    unsigned *enc_min_s;  // [3S] encoded min of shifted coords
    unsigned *enc_min_o;  // [3S] encoded min of original coords
    unsigned *enc_max_o;  // [3S]
    int *cls;             // [S] class of the segment (class of its first point)
    int *min_pts;         // [S]
    float *r2;            // [S]
    float *inv_h;         // [S]
    float *min_s;         // [3S]
    float *min_o;         // [3S]
    float *inv_g;         // [S]
    int *cell_start;      // [S+1] first cell ordinal of each segment
    int *lab_start;       // [S+1] first labelled-list position of each segment
    int *id_base;         // [S] global kept-cluster index at which this segment's CALL starts
    int *k_base;          // [S] global kept-cluster index of this segment's first cluster
    int *cluster_num;     // [S]
};

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ uint64_t spread3(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

__device__ __forceinline__ int uf_find(int *parent, int x) {
    // parent pointers only ever move to smaller ordinals -> acyclic; .cg loads bypass the
    // non-coherent L1 so a stale self-pointer cannot livelock the CAS loop in uf_union
    while (true) {
        int p = __ldcg(parent + x);
        if (p == x) return x;
        int gp = __ldcg(parent + p);
        if (gp == p) return p;
        __stcg(parent + x, gp);  // path halving
        x = gp;
    }
}
__device__ __forceinline__ void uf_union(int *parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            int t = a;
            a = b;
            b = t;
        }
        if (atomicCAS(parent + a, a, b) == a) return;  // hook the larger root under the smaller
    }
}

// ------------------------------------------------------------------------------------------------
// K1  per point: segment id, validation, per-segment bounding boxes
// ------------------------------------------------------------------------------------------------
__global__ void k_prep_points(int n, int S, SegArrays sg, const float *__restrict__ x,
                              const float *__restrict__ y, const float *__restrict__ z,
                              const float *__restrict__ xo, const float *__restrict__ yo,
                              const float *__restrict__ zo, const int *__restrict__ sem,
                              int *__restrict__ seg_of, int *err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    int s = 0;
    float vx = 0, vy = 0, vz = 0, ox = 0, oy = 0, oz = 0;
    if (valid) {
        int lo = 0, hi = S - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (__ldg(sg.start + mid) <= i) lo = mid;
            else hi = mid - 1;
        }
        s = lo;
        seg_of[i] = s;
        vx = x[i], vy = y[i], vz = z[i], ox = xo[i], oy = yo[i], oz = zo[i];
        int c = sem[i];
        int e = 0;
        if (c < 2 || c > 19) e |= kErrSem;
        if (!(isfinite(vx) && isfinite(vy) && isfinite(vz) && isfinite(ox) && isfinite(oy) && isfinite(oz)))
            e |= kErrNonFinite;
        if (e) atomicOr(err, e);
    }
    unsigned act = __ballot_sync(kFull, valid);
    if (!valid) return;
    int s0 = __shfl_sync(act, s, __ffs(act) - 1);
    bool uniform = __all_sync(act, s == s0);
    unsigned e0 = enc_f(vx), e1 = enc_f(vy), e2 = enc_f(vz), e3 = enc_f(ox), e4 = enc_f(oy), e5 = enc_f(oz);
    if (uniform) {
        unsigned m0 = __reduce_min_sync(act, e0), m1 = __reduce_min_sync(act, e1), m2 = __reduce_min_sync(act, e2);
        unsigned m3 = __reduce_min_sync(act, e3), m4 = __reduce_min_sync(act, e4), m5 = __reduce_min_sync(act, e5);
        unsigned M3 = __reduce_max_sync(act, e3), M4 = __reduce_max_sync(act, e4), M5 = __reduce_max_sync(act, e5);
        if (lane_id() == __ffs(act) - 1) {
            atomicMin(sg.enc_min_s + 3 * s, m0);
            atomicMin(sg.enc_min_s + 3 * s + 1, m1);
            atomicMin(sg.enc_min_s + 3 * s + 2, m2);
            atomicMin(sg.enc_min_o + 3 * s, m3);
            atomicMin(sg.enc_min_o + 3 * s + 1, m4);
            atomicMin(sg.enc_min_o + 3 * s + 2, m5);
            atomicMax(sg.enc_max_o + 3 * s, M3);
            atomicMax(sg.enc_max_o + 3 * s + 1, M4);
            atomicMax(sg.enc_max_o + 3 * s + 2, M5);
        }
    } else {
        atomicMin(sg.enc_min_s + 3 * s, e0);
        atomicMin(sg.enc_min_s + 3 * s + 1, e1);
        atomicMin(sg.enc_min_s + 3 * s + 2, e2);
        atomicMin(sg.enc_min_o + 3 * s, e3);
        atomicMin(sg.enc_min_o + 3 * s + 1, e4);
        atomicMin(sg.enc_min_o + 3 * s + 2, e5);
        atomicMax(sg.enc_max_o + 3 * s, e3);
        atomicMax(sg.enc_max_o + 3 * s + 1, e4);
        atomicMax(sg.enc_max_o + 3 * s + 2, e5);
    }
}

// K2  per segment: radius / cell edge / origin
__global__ void k_seg_params(int n, int S, SegArrays sg, const int *__restrict__ sem,
                             const float *__restrict__ radius_tab, const int *__restrict__ min_pts_tab) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = sg.start[s], e = sg.start[s + 1];
    int c = 2;
    if (e > b) c = sem[b];
    if (c < 2 || c > 19) c = 2;  // flagged by k_prep_points
    sg.cls[s] = c;
    sg.min_pts[s] = min_pts_tab[c - 2];
    float r = radius_tab[c - 2];
    sg.r2[s] = __fmul_rn(r, r);  // binary_cuda_functions.cu:85  cur_radius * cur_radius
    // cell edge h = r/2 * (1 + 2^-7): h*sqrt(3) < r (one cell = clique) and 2 cells >= r (5^3 stencil
    // is a superset of the r-ball) with margins far above fp32 rounding of the cell coordinate
    float h = r * 0.5f * (1.0f + 1.0f / 128.0f);
    if (!(h > 0.f)) h = 1e-6f;
    sg.inv_h[s] = 1.0f / h;
    float ext = 0.f;
    for (int k = 0; k < 3; k++) {
        float mn = e > b ? dec_f(sg.enc_min_s[3 * s + k]) : 0.f;
        sg.min_s[3 * s + k] = mn;
        float mo = e > b ? dec_f(sg.enc_min_o[3 * s + k]) : 0.f;
        float Mo = e > b ? dec_f(sg.enc_max_o[3 * s + k]) : 0.f;
        sg.min_o[3 * s + k] = mo;
        ext = fmaxf(ext, Mo - mo);
    }
    // LP-assignment sort grid (ordering only, never a correctness filter): 1 cm cells unless the
    // segment is too large for 14 bits per axis
    float g = fmaxf(0.01f, ext / 16000.0f);
    sg.inv_g[s] = 1.0f / g;
}

// K3  per point: 64-bit sort keys.  key1 = seg | cz | cy | cx (shifted space, cell edge h);
//     key2 = seg | morton(original space, cell edge g) for the LP-assignment ordering
__global__ void k_keys(int n, SegArrays sg, const float *__restrict__ x, const float *__restrict__ y,
                       const float *__restrict__ z, const float *__restrict__ xo,
                       const float *__restrict__ yo, const float *__restrict__ zo,
                       const int *__restrict__ sem, const int *__restrict__ seg_of,
                       uint64_t *__restrict__ key1, uint64_t *__restrict__ key2,
                       uint32_t *__restrict__ val, int *err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = seg_of[i];
    if (sem[i] != sg.cls[s]) atomicOr(err, kErrMixed);
    float ih = sg.inv_h[s];
    int c[3];
    float v[3] = {x[i], y[i], z[i]};
    int e = 0;
    for (int k = 0; k < 3; k++) {
        float f = __fmul_rn(__fsub_rn(v[k], sg.min_s[3 * s + k]), ih);
        int q = (f >= 0.f && f < 1e9f) ? (int)f : 0;
        if (q > kCellMax) {
            q = kCellMax;
            e = kErrRange;
        }
        c[k] = q;
    }
    if (e) atomicOr(err, e);
    key1[i] = ((uint64_t)s << kSegShift) | ((uint64_t)c[2] << (2 * kCellBits)) | ((uint64_t)c[1] << kCellBits) |
              (uint64_t)c[0];
    float ig = sg.inv_g[s];
    float o[3] = {xo[i], yo[i], zo[i]};
    uint32_t m[3];
    for (int k = 0; k < 3; k++) {
        float f = (o[k] - sg.min_o[3 * s + k]) * ig;
        int q = (f >= 0.f && f < 1e9f) ? (int)f : 0;
        m[k] = (uint32_t)min(q, kCellMax);
    }
    key2[i] = ((uint64_t)s << kSegShift) | spread3(m[0]) | (spread3(m[1]) << 1) | (spread3(m[2]) << 2);
    val[i] = (uint32_t)i;
}

// K4  after sort 1: gather coordinates into cell order, flag cell heads
__global__ void k_gather_heads(int n, const uint64_t *__restrict__ skey, const uint32_t *__restrict__ order,
                               const float *__restrict__ x, const float *__restrict__ y,
                               const float *__restrict__ z, float4 *__restrict__ pts4, int *__restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = order[i];
    pts4[i] = make_float4(x[o], y[o], z[o], __int_as_float((int)o));
    head[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
}

// K5  cell table from the scanned head flags
__global__ void k_cells(int n, const uint64_t *__restrict__ skey, const int *__restrict__ head,
                        const int *__restrict__ head_excl, int *__restrict__ cell_of,
                        int *__restrict__ cell_start, uint64_t *__restrict__ cell_key,
                        int *__restrict__ parent, int *__restrict__ cell_hp, int *__restrict__ cell_minhp,
                        int *__restrict__ comp_min, int *__restrict__ d_C) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int h = head[i];
    int c = head_excl[i] + h - 1;
    cell_of[i] = c;
    if (h) {
        cell_start[c] = i;
        cell_key[c] = skey[i];
        parent[c] = c;
        cell_hp[c] = 0;
        cell_minhp[c] = 0x7fffffff;
        comp_min[c] = 0x7fffffff;
    }
    if (i == n - 1) {
        cell_start[c + 1] = n;
        *d_C = c + 1;
    }
}

// K6  first cell ordinal of every segment
__global__ void k_seg_cells(int n, int S, SegArrays sg, const int *__restrict__ cell_of,
                            const int *__restrict__ d_C) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > S) return;
    int b = sg.start[s];
    sg.cell_start[s] = (b < n) ? cell_of[b] : *d_C;
}

// K7  stencil rows of every cell: runs[c*25 + (dz+2)*5 + (dy+2)] = [first cell, last cell+1) with
//     cx-2 <= x <= cx+2 in row (cy+dy, cz+dz) — a contiguous range because cells are sorted x-fastest
__global__ void k_runs(SegArrays sg, const uint64_t *__restrict__ cell_key, const int *__restrict__ d_C,
                       int2 *__restrict__ runs) {
    long long total = (long long)(*d_C) * kRuns;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int c = (int)(t / kRuns), k = (int)(t % kRuns);
        uint64_t key = cell_key[c];
        int s = (int)(key >> kSegShift);
        int cx = (int)(key & kCellMax), cy = (int)((key >> kCellBits) & kCellMax),
            cz = (int)((key >> (2 * kCellBits)) & kCellMax);
        int ny = cy + (k % 5) - 2, nz = cz + (k / 5) - 2;
        int2 out = make_int2(0, 0);
        if (ny >= 0 && ny <= kCellMax && nz >= 0 && nz <= kCellMax) {
            uint64_t base = ((uint64_t)s << kSegShift) | ((uint64_t)nz << (2 * kCellBits)) | ((uint64_t)ny << kCellBits);
            uint64_t klo = base | (uint64_t)max(cx - 2, 0);
            uint64_t khi = base | (uint64_t)min(cx + 2, kCellMax);  // inclusive
            int b = sg.cell_start[s], e = sg.cell_start[s + 1];
            int lo = b, hi = e;
            while (lo < hi) {  // first cell with key >= klo
                int mid = (lo + hi) >> 1;
                if (__ldg(cell_key + mid) < klo) lo = mid + 1;
                else hi = mid;
            }
            int first = lo;
            hi = e;
            while (lo < hi) {  // first cell with key > khi
                int mid = (lo + hi) >> 1;
                if (__ldg(cell_key + mid) <= khi) lo = mid + 1;
                else hi = mid;
            }
            out = make_int2(first, lo);
        }
        runs[t] = out;
    }
}

// ------------------------------------------------------------------------------------------------
// K8  degree (pass A) — the dominant kernel.  One warp per 32 consecutive sorted points; the lanes
//     that share a cell walk that cell's 25 stencil rows together, every candidate is one broadcast
//     16-B load tested by all lanes with the exact predicate.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_degree(int n, SegArrays sg, const float4 *__restrict__ pts4, const int *__restrict__ cell_of,
         const int *__restrict__ cell_start, const uint64_t *__restrict__ cell_key,
         const int2 *__restrict__ runs, int *__restrict__ deg_sorted, unsigned long long *__restrict__ n_tests) {
    int i = (blockIdx.x * blockDim.x + threadIdx.x);
    int lane = lane_id();
    bool valid = i < n;
    float4 p = valid ? pts4[i] : make_float4(0, 0, 0, 0);
    int c = valid ? cell_of[i] : -1;
    int deg = 0;
    unsigned long long tests = 0;  // candidate tests issued for this warp's query points (profiling)
    unsigned todo = __ballot_sync(kFull, valid);
    while (todo) {
        int leader = __ffs(todo) - 1;
        int cL = __shfl_sync(kFull, c, leader);
        bool mine = (c == cL);
        unsigned mask = __ballot_sync(kFull, mine);
        todo &= ~mask;
        float r2 = sg.r2[(int)(cell_key[cL] >> kSegShift)];
        int jb = 0, je = 0;
        if (lane < kRuns) {
            int2 rr = runs[(long long)cL * kRuns + lane];
            if (rr.y > rr.x) {
                jb = cell_start[rr.x];
                je = cell_start[rr.y];
            }
        }
        int cnt0 = 0, cnt1 = 0, cnt2 = 0, cnt3 = 0;
        unsigned cand = 0;
#pragma unroll 1
        for (int k = 0; k < kRuns; k++) {
            int b = __shfl_sync(kFull, jb, k), e = __shfl_sync(kFull, je, k);
            cand += (unsigned)(e - b);
            int j = b;
            for (; j + 4 <= e; j += 4) {
                float4 q0 = __ldg(pts4 + j), q1 = __ldg(pts4 + j + 1), q2 = __ldg(pts4 + j + 2), q3 = __ldg(pts4 + j + 3);
                cnt0 += sqd(p.x, p.y, p.z, q0.x, q0.y, q0.z) <= r2;
                cnt1 += sqd(p.x, p.y, p.z, q1.x, q1.y, q1.z) <= r2;
                cnt2 += sqd(p.x, p.y, p.z, q2.x, q2.y, q2.z) <= r2;
                cnt3 += sqd(p.x, p.y, p.z, q3.x, q3.y, q3.z) <= r2;
            }
            for (; j < e; j++) {
                float4 q0 = __ldg(pts4 + j);
                cnt0 += sqd(p.x, p.y, p.z, q0.x, q0.y, q0.z) <= r2;
            }
        }
        if (mine) deg = cnt0 + cnt1 + cnt2 + cnt3 - 1;  // binary_cuda_functions.cu:88  ans - 1
        tests += (unsigned long long)cand * (unsigned)__popc(mask);
    }
    if (n_tests && lane == 0) atomicAdd(n_tests, tests);
    if (valid) deg_sorted[i] = deg;
}

// K9  HP rule + per-cell HP statistics + degree scatter to input order
__global__ void k_hp_cells(int n, SegArrays sg, float4 *__restrict__ pts4, const int *__restrict__ cell_of,
                           const uint64_t *__restrict__ cell_key, const int *__restrict__ deg_sorted,
                           int *__restrict__ degree_out, int *__restrict__ cell_hp, int *__restrict__ cell_minhp,
                           unsigned long long *__restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    unsigned act = __ballot_sync(kFull, valid);
    if (!valid) return;
    int c = cell_of[i];
    int s = (int)(cell_key[c] >> kSegShift);
    int d = deg_sorted[i];
    int *w = reinterpret_cast<int *>(pts4 + i) + 3;
    int orig = *w;
    bool hp = d >= sg.min_pts[s];  // binary_cuda_functions.cu:175-186
    degree_out[orig] = d;
    if (hp) *w = orig | kHpBit;
    unsigned grp = __match_any_sync(act, c);
    unsigned hpm = __ballot_sync(act, hp) & grp;
    int mn = __reduce_min_sync(grp, hp ? orig : 0x7fffffff);
    if (hpm && lane_id() == __ffs(grp) - 1) {
        atomicAdd(cell_hp + c, __popc(hpm));
        atomicMin(cell_minhp + c, mn);
    }
    if (counters) {  // profiling only: [1] sum of degrees, [2] HP count
        unsigned long long dsum = 0, hsum = 0;
        // lanes outside `act` have exited; use a full-mask-free reduction over act via atomics per warp leader
        for (unsigned m = act; m; m &= m - 1) {
            int l = __ffs(m) - 1;
            dsum += (unsigned)__shfl_sync(act, d, l);
            hsum += (unsigned)__shfl_sync(act, (int)hp, l);
        }
        if (lane_id() == __ffs(act) - 1) {
            atomicAdd(counters + 1, dsum);
            atomicAdd(counters + 2, hsum);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K10  HP connectivity (pass B): union-find over cells.  One warp per HP-cell A; for every stencil
//      cell B > A that holds HPs and is not yet in A's set, look for ONE HP pair within r.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_union(SegArrays sg, const float4 *__restrict__ pts4, const int *__restrict__ cell_start,
        const uint64_t *__restrict__ cell_key, const int2 *__restrict__ runs,
        const int *__restrict__ cell_hp, int *parent, const int *__restrict__ d_C) {
    int C = *d_C;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int A = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; A < C; A += warps) {
        if (cell_hp[A] == 0) continue;
        float r2 = sg.r2[(int)(cell_key[A] >> kSegShift)];
        int a0 = cell_start[A], a1 = cell_start[A + 1];
        int2 rr = make_int2(0, 0);
        if (lane < kRuns) rr = runs[(long long)A * kRuns + lane];
        for (int k = 0; k < kRuns; k++) {
            int c0 = __shfl_sync(kFull, rr.x, k), c1 = __shfl_sync(kFull, rr.y, k);
            for (int B = max(c0, A + 1); B < c1; B++) {
                if (cell_hp[B] == 0) continue;
                int same = 0;
                if (lane == 0) same = (uf_find(parent, A) == uf_find(parent, B));
                if (__shfl_sync(kFull, same, 0)) continue;
                int b0 = cell_start[B], b1 = cell_start[B + 1];
                bool found = false;
                for (int ia = a0; ia < a1 && !found; ia += 32) {
                    int i = ia + lane;
                    float4 p = (i < a1) ? pts4[i] : make_float4(0, 0, 0, 0);
                    bool php = (i < a1) && (__float_as_int(p.w) & kHpBit);
                    if (!__any_sync(kFull, php)) continue;
                    for (int j = b0; j < b1; j++) {
                        float4 q = __ldg(pts4 + j);
                        if (!(__float_as_int(q.w) & kHpBit)) continue;
                        bool hit = php && (sqd(p.x, p.y, p.z, q.x, q.y, q.z) <= r2);
                        if (__any_sync(kFull, hit)) {
                            found = true;
                            break;
                        }
                    }
                }
                if (found && lane == 0) uf_union(parent, A, B);
                __syncwarp();
            }
        }
    }
}

// K11  flatten + minimum HP index of every component
__global__ void k_comp_min(const int *__restrict__ d_C, const int *__restrict__ cell_hp, int *parent,
                           const int *__restrict__ cell_minhp, int *comp_min) {
    int C = *d_C;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        if (cell_hp[c] == 0) continue;
        int r = uf_find(parent, c);
        if (r != c) __stcg(parent + c, r);
        atomicMin(comp_min + r, cell_minhp[c]);
    }
}

// K12  flag the minimum-index HP of every component (cluster numbering = rank of that index,
//      binary.cu:161-166: seeds are taken in ascending point order)
__global__ void k_flag_roots(const int *__restrict__ d_C, const int *__restrict__ cell_hp,
                             const int *__restrict__ parent, const int *__restrict__ comp_min,
                             int *__restrict__ flag) {
    int C = *d_C;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x)
        if (cell_hp[c] > 0 && parent[c] == c) flag[comp_min[c]] = 1;
}

// K13  raw cluster id of every HP-cell; representative point of every raw cluster
__global__ void k_cell_gid(const int *__restrict__ d_C, const int *__restrict__ cell_hp,
                           const int *__restrict__ parent, const int *__restrict__ comp_min,
                           const int *__restrict__ gid_at, int *__restrict__ cell_gid, int *__restrict__ rep) {
    int C = *d_C;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        int g = -1;
        if (cell_hp[c] > 0) {
            int r = parent[c];
            int u = comp_min[r];
            g = gid_at[u];
            if (r == c) rep[g] = u;
        }
        cell_gid[c] = g;
    }
}

// ------------------------------------------------------------------------------------------------
// K14  labels (pass C).  HPs take their component's raw id; a border LP takes the MAXIMUM raw id
//      among components owning an HP within r (later BFS overwrites earlier, binary.cu:206-213);
//      LPs with no HP neighbour stay -1.  Cluster sizes (incl. border LPs) are counted here.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_label(int n, SegArrays sg, const float4 *__restrict__ pts4, const int *__restrict__ cell_of,
        const int *__restrict__ cell_start, const uint64_t *__restrict__ cell_key,
        const int2 *__restrict__ runs, const int *__restrict__ cell_hp, const int *__restrict__ cell_gid,
        int *__restrict__ raw_label, int *__restrict__ raw_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = lane_id();
    bool valid = i < n;
    float4 p = valid ? pts4[i] : make_float4(0, 0, 0, 0);
    int c = valid ? cell_of[i] : -1;
    bool hp = valid && (__float_as_int(p.w) & kHpBit);
    int label = -1;
    if (hp) label = cell_gid[c];
    unsigned todo = __ballot_sync(kFull, valid && !hp);
    while (todo) {
        int leader = __ffs(todo) - 1;
        int cL = __shfl_sync(kFull, c, leader);
        bool mine = (c == cL) && valid && !hp;
        unsigned mask = __ballot_sync(kFull, mine);
        todo &= ~mask;
        float r2 = sg.r2[(int)(cell_key[cL] >> kSegShift)];
        int2 rr = make_int2(0, 0);
        if (lane < kRuns) rr = runs[(long long)cL * kRuns + lane];
        int best = -1;
        for (int k = 0; k < kRuns; k++) {
            int c0 = __shfl_sync(kFull, rr.x, k), c1 = __shfl_sync(kFull, rr.y, k);
            for (int B = c0; B < c1; B++) {
                if (cell_hp[B] == 0) continue;
                int gB = cell_gid[B];
                bool pend = mine && best < gB;
                if (!__any_sync(kFull, pend)) continue;
                int b0 = cell_start[B], b1 = cell_start[B + 1];
                for (int j = b0; j < b1; j++) {
                    float4 q = __ldg(pts4 + j);
                    if (!(__float_as_int(q.w) & kHpBit)) continue;
                    if (pend && sqd(p.x, p.y, p.z, q.x, q.y, q.z) <= r2) {
                        best = gB;
                        pend = false;
                    }
                    if (!__any_sync(kFull, pend)) break;
                }
            }
        }
        if (mine) label = best;
    }
    if (valid) raw_label[__float_as_int(p.w) & ~kHpBit] = label;
    // cluster sizes: one atomic per distinct label in the warp
    unsigned act = __ballot_sync(kFull, valid && label >= 0);
    if (valid && label >= 0) {
        unsigned grp = __match_any_sync(act, label);
        if (lane == __ffs(grp) - 1) atomicAdd(raw_count + label, __popc(grp));
    }
}

// K15  fragment filter (binary.cu:219-268): drop raw cluster g iff float(size) < mean_count*para_f
__global__ void k_filter(const int *__restrict__ d_R, SegArrays sg, const int *__restrict__ rep,
                         const int *__restrict__ seg_of, const int *__restrict__ raw_count,
                         const float *__restrict__ thresh18, int *__restrict__ keep) {
    int R = *d_R;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < R; g += gridDim.x * blockDim.x) {
        int s = seg_of[rep[g]];
        float t = thresh18[sg.cls[s] - 2];
        keep[g] = ((float)raw_count[g] < t) ? 0 : 1;
    }
}

__device__ __forceinline__ int kept_before(const int *kscan, const int *d_K, int g, int R) {
    return g < R ? kscan[g] : *d_K;
}

// K16  per segment: cluster count, first kept-cluster index, id base of the segment's call
__global__ void k_seg_clusters(int n, int S, SegArrays sg, const int *__restrict__ seg_call_first,
                               const int *__restrict__ gid_at, const int *__restrict__ d_R,
                               const int *__restrict__ kscan, const int *__restrict__ d_K,
                               int *__restrict__ cluster_num_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int R = *d_R;
    int b = sg.start[s], e = sg.start[s + 1];
    int g0 = b < n ? gid_at[b] : R, g1 = e < n ? gid_at[e] : R;
    int k0 = kept_before(kscan, d_K, g0, R), k1 = kept_before(kscan, d_K, g1, R);
    sg.k_base[s] = k0;
    sg.cluster_num[s] = k1 - k0;
    cluster_num_out[s] = k1 - k0;
    int f = sg.start[seg_call_first[s]];
    int gf = f < n ? gid_at[f] : R;
    sg.id_base[s] = kept_before(kscan, d_K, gf, R);
}

// K17  final ids of HP-stage labels; query flags for LP assignment; per-cluster metadata
__global__ void k_relabel(int n, SegArrays sg, const int *__restrict__ seg_of, const int *__restrict__ raw_label,
                          const int *__restrict__ keep, const int *__restrict__ kscan, int assign_lp,
                          int *__restrict__ cluster_id, int *__restrict__ qflag, int *__restrict__ clt_sem,
                          int *__restrict__ clt_seg, const int *__restrict__ rep) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    int g = raw_label[u];
    int s = seg_of[u];
    int id = -1;
    if (g >= 0 && keep[g]) {
        int kk = kscan[g];
        id = kk - sg.id_base[s];
        if (rep[g] == u) {
            clt_sem[kk] = sg.cls[s];
            clt_seg[kk] = s;
        }
    }
    cluster_id[u] = id;
    qflag[u] = (id < 0 && assign_lp && sg.cluster_num[s] > 0) ? 1 : 0;
}

// K18a  labelled flags in LP-assignment order (order2 = points sorted by segment | morton(original))
__global__ void k_lab_flags(int n, const uint32_t *__restrict__ order2, const int *__restrict__ cluster_id,
                            int *__restrict__ labflag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    labflag[i] = cluster_id[order2[i]] >= 0 ? 1 : 0;
}

// K18b  compaction: query list (input order) and labelled list (order2) as float4 {xo,yo,zo,index}
__global__ void k_compact(int n, const int *__restrict__ qflag, const int *__restrict__ qpos,
                          int *__restrict__ qlist, const uint32_t *__restrict__ order2,
                          const int *__restrict__ labflag, const int *__restrict__ lpos,
                          const float *__restrict__ xo, const float *__restrict__ yo,
                          const float *__restrict__ zo, float4 *__restrict__ lab4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (qflag[i]) qlist[qpos[i]] = i;
    if (labflag[i]) {
        uint32_t o = order2[i];
        lab4[lpos[i]] = make_float4(xo[o], yo[o], zo[o], __int_as_float((int)o));
    }
}

__global__ void k_seg_lab(int n, int S, SegArrays sg, const int *__restrict__ lpos, const int *__restrict__ d_L) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > S) return;
    int b = sg.start[s];
    sg.lab_start[s] = b < n ? lpos[b] : *d_L;
}

// K19  bounding boxes of 32-point groups of the labelled list
__global__ void k_lab_boxes(const int *__restrict__ d_L, const float4 *__restrict__ lab4,
                            float4 *__restrict__ box_lo, float4 *__restrict__ box_hi) {
    int L = *d_L;
    int G = (L + 31) >> 5;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < G; g += warps) {
        int j = g * 32 + lane;
        float4 q = lab4[min(j, L - 1)];
        float lx = q.x, ly = q.y, lz = q.z, hx = q.x, hy = q.y, hz = q.z;
        for (int o = 16; o; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(kFull, lx, o));
            ly = fminf(ly, __shfl_xor_sync(kFull, ly, o));
            lz = fminf(lz, __shfl_xor_sync(kFull, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(kFull, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(kFull, hy, o));
            hz = fmaxf(hz, __shfl_xor_sync(kFull, hz, o));
        }
        if (lane == 0) {
            box_lo[g] = make_float4(lx, ly, lz, 0.f);
            box_hi[g] = make_float4(hx, hy, hz, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K20  LP assignment (binary.cu:270-358, binary_cuda_functions.cu:258-302): exact 1-NN over the
//      labelled points of the same segment in ORIGINAL coordinates, ties -> largest index.
//      One warp per query; branch-and-bound over the 32-point group boxes: sweep 1 finds the group
//      with the smallest lower bound, sweep 2 visits every group whose bound does not exceed the best
//      distance so far.  A group is pruned only if its (conservatively shrunk) bound is STRICTLY
//      greater than the best distance, so equal-distance ties are never lost.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float box_lb(float px, float py, float pz, float4 lo, float4 hi) {
    float dx = fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f);
    float dy = fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f);
    float dz = fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f);
    return (dx * dx + dy * dy + dz * dz) * (1.0f - 1e-5f);
}

__device__ __forceinline__ void nn_scan_group(int g, int l0, int l1, int lane, float px, float py, float pz,
                                              const float4 *__restrict__ lab4, float &bestD, int &bestI) {
    int j = g * 32 + lane;
    bool ok = j >= l0 && j < l1;
    float D = __int_as_float(0x7f800000);
    int idx = -1;
    if (ok) {
        float4 q = __ldg(lab4 + j);
        D = sqd(px, py, pz, q.x, q.y, q.z);
        idx = __float_as_int(q.w);
    }
    unsigned db = __float_as_uint(D);  // D >= 0: bit pattern is order preserving
    unsigned dmin = __reduce_min_sync(kFull, db);
    int imax = __reduce_max_sync(kFull, (db == dmin) ? idx : -1);
    float Dm = __uint_as_float(dmin);
    if (imax >= 0 && (Dm < bestD || (Dm == bestD && imax > bestI))) {
        bestD = Dm;
        bestI = imax;
    }
}

__global__ void __launch_bounds__(256)
k_nn(const int *__restrict__ d_Q, SegArrays sg, const int *__restrict__ qlist, const int *__restrict__ seg_of,
     const float *__restrict__ xo, const float *__restrict__ yo, const float *__restrict__ zo,
     const float4 *__restrict__ lab4, const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi,
     int *cluster_id, unsigned long long *__restrict__ counters) {
    int Q = *d_Q;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < Q; qi += warps) {
        int p = qlist[qi];
        int s = seg_of[p];
        int l0 = sg.lab_start[s], l1 = sg.lab_start[s + 1];
        if (l1 <= l0) continue;
        float px = xo[p], py = yo[p], pz = zo[p];
        int g0 = l0 >> 5, g1 = (l1 - 1) >> 5;
        // sweep 1
        float lbmin = __int_as_float(0x7f800000);
        int gmin = g0;
        for (int gb = g0; gb <= g1; gb += 32) {
            int g = gb + lane;
            float lb = __int_as_float(0x7f800000);
            if (g <= g1) lb = box_lb(px, py, pz, __ldg(box_lo + g), __ldg(box_hi + g));
            unsigned m = __reduce_min_sync(kFull, __float_as_uint(lb));
            if (__uint_as_float(m) < lbmin) {
                lbmin = __uint_as_float(m);
                unsigned who = __ballot_sync(kFull, __float_as_uint(lb) == m);
                gmin = gb + __ffs(who) - 1;
            }
        }
        float bestD = __int_as_float(0x7f800000);
        int bestI = -1;
        nn_scan_group(gmin, l0, l1, lane, px, py, pz, lab4, bestD, bestI);
        // sweep 2
        for (int gb = g0; gb <= g1; gb += 32) {
            int g = gb + lane;
            float lb = __int_as_float(0x7f800000);
            if (g <= g1 && g != gmin) lb = box_lb(px, py, pz, __ldg(box_lo + g), __ldg(box_hi + g));
            unsigned m = __ballot_sync(kFull, lb <= bestD);
            while (m) {
                int gl = __ffs(m) - 1;
                m &= m - 1;
                float lbg = __shfl_sync(kFull, lb, gl);
                if (lbg > bestD) continue;
                nn_scan_group(gb + gl, l0, l1, lane, px, py, pz, lab4, bestD, bestI);
            }
        }
        if (lane == 0 && bestI >= 0) cluster_id[p] = cluster_id[bestI];
    }
    if (counters && blockIdx.x == 0 && threadIdx.x == 0) counters[3] = (unsigned long long)Q;
}

// ------------------------------------------------------------------------------------------------
// K21  cluster centres (binary_cuda_functions.cu:217-246): sequential running mean in ascending
//      point order, M += (x - M) / n with IEEE division — replayed exactly, one warp per cluster.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_centres(const int *__restrict__ d_K, SegArrays sg, const int *__restrict__ clt_seg,
          const int *__restrict__ cluster_id, const float *__restrict__ x, const float *__restrict__ y,
          const float *__restrict__ z, float *__restrict__ center) {
    int K = *d_K;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int kk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; kk < K; kk += warps) {
        int s = clt_seg[kk];
        int local = kk - sg.id_base[s];
        int b = sg.start[s], e = sg.start[s + 1];
        float mx = 0.f, my = 0.f, mz = 0.f;
        int cnt = 0;
        for (int ub = b; ub < e; ub += 32) {
            int u = ub + lane;
            bool hit = (u < e) && (cluster_id[u] == local);
            unsigned m = __ballot_sync(kFull, hit);
            if (!m) continue;
            float vx = 0.f, vy = 0.f, vz = 0.f;
            if (hit) vx = x[u], vy = y[u], vz = z[u];
            while (m) {
                int l = __ffs(m) - 1;
                m &= m - 1;
                float ax = __shfl_sync(kFull, vx, l), ay = __shfl_sync(kFull, vy, l), az = __shfl_sync(kFull, vz, l);
                cnt++;
                float fn = (float)cnt;
                mx = __fadd_rn(mx, __fdiv_rn(__fsub_rn(ax, mx), fn));
                my = __fadd_rn(my, __fdiv_rn(__fsub_rn(ay, my), fn));
                mz = __fadd_rn(mz, __fdiv_rn(__fsub_rn(az, mz), fn));
            }
        }
        if (lane == 0) {
            center[3 * kk] = mx;
            center[3 * kk + 1] = my;
            center[3 * kk + 2] = mz;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 (three kernels: block sums, spine, down-sweep).  n may live on the device.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_excl_scan(int v, int *smem, int &total) {
    int lane = lane_id(), wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < (kScanThreads >> 5)) ? smem[lane] : 0;
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kScanThreads >> 5)) smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = inc - v + smem[wid];
    total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const int *__restrict__ in, int n_host, const int *__restrict__ n_dev, int *__restrict__ block_sums) {
    __shared__ int smem[33];
    int n = n_dev ? *n_dev : n_host;
    int base = blockIdx.x * kScanTile;
    if (base >= n) {
        if (threadIdx.x == 0) block_sums[blockIdx.x] = 0;
        return;
    }
    int sum = 0;
    for (int k = 0; k < kScanItems; k++) {
        int i = base + k * kScanThreads + threadIdx.x;
        if (i < n) sum += in[i];
    }
    int total;
    block_excl_scan(sum, smem, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_spine(int *block_sums, int nb, int *__restrict__ total_out) {
    __shared__ int smem[33];
    int carry = 0;
    for (int base = 0; base < nb; base += kScanThreads) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int total;
        int ex = block_excl_scan(v, smem, total);
        if (i < nb) block_sums[i] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_down(const int *__restrict__ in, int n_host, const int *__restrict__ n_dev,
            const int *__restrict__ block_offs, int *__restrict__ out) {
    __shared__ int smem[33];
    int n = n_dev ? *n_dev : n_host;
    int base = blockIdx.x * kScanTile;
    if (base >= n) return;
    // thread t owns items [base + t*kScanItems, +kScanItems): contiguous per thread
    int v[kScanItems];
    int sum = 0;
    int i0 = base + threadIdx.x * kScanItems;
    for (int k = 0; k < kScanItems; k++) {
        int i = i0 + k;
        v[k] = i < n ? in[i] : 0;
        sum += v[k];
    }
    int total;
    int ex = block_excl_scan(sum, smem, total) + block_offs[blockIdx.x];
    for (int k = 0; k < kScanItems; k++) {
        int i = i0 + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

}  // namespace pb
