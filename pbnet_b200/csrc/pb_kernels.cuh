// pb_kernels.cuh — hand-written sm_100a kernels of the instance-grouping path.
//
// Reference semantics: SURVEY.md Appendix A (normative restatement of
// lib/PB_lib/src/pbnet/{cluster.cu,binary.cu,binary_cuda_functions.cu}).  Nothing here is a
// translation of those kernels: the reference materialises neighbour lists and drives a BFS from the
// host; this file works on a two-level sorted cell grid with a cell-level union-find.
//
// Data layout in HBM (N points of all segments concatenated, S segments):
//   pts4[N]       float4 {x,y,z, bits(orig_index | HP<<31)} sorted by
//                 key = seg | coarse(z,y,x) | fine(z,y,x bit)  — one 16-B broadcast load per candidate
//   fine cells    edge h = r/2*(1+2^-7) < r/sqrt(3): any two points of one fine cell are neighbours,
//                 so HP connectivity is a union-find over FINE CELLS (k_union) and border LPs only
//                 probe cells whose cluster id could still raise their maximum (k_label)
//   coarse cells  2x2x2 fine cells, edge >= r: the 27-cell stencil (9 contiguous x-rows, runs9[])
//                 is a superset of the r-ball; candidates are enumerated per coarse ROW so that a
//                 warp's 128 query points share one candidate stream (k_degree)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int kCellBits = 14;                  // fine cell coordinate bits per axis
constexpr int kCellMax = (1 << kCellBits) - 1;
constexpr int kCoarseBits = kCellBits - 1;     // 13
constexpr int kCoarseMax = (1 << kCoarseBits) - 1;
constexpr int kSegShift = 3 * kCellBits;       // 42: key = seg<<42 | cz'<<29 | cy'<<16 | cx'<<3 | fine bits
constexpr int kRowShift = 3 + kCoarseBits;     // 16: key >> 16 identifies (seg, cz', cy') = one coarse row
#ifndef PB_MORTON_BITS
#define PB_MORTON_BITS 9
#endif
constexpr int kMortonBits = PB_MORTON_BITS;    // LP-assignment order: Morton bits per axis
constexpr int kMortonMax = (1 << kMortonBits) - 1;
constexpr int kKey2SegShift = 3 * kMortonBits; // key2 = seg<<27 | morton27  (mixed: seg<<32 | class<<27 | morton27)
constexpr int kRuns = 9;                       // 3 x 3 coarse stencil rows, each up to 3 coarse cells long
constexpr unsigned kFull = 0xffffffffu;
constexpr int kHpBit = 0x80000000;
#ifndef PB_DEG_MINB
#define PB_DEG_MINB 9
#endif
#ifndef PB_DEG_WINDOW
#define PB_DEG_WINDOW 128
#endif
constexpr int kWindow = PB_DEG_WINDOW;         // query points per warp in k_degree (a multiple of 64: 2 or 3 adjacent pairs per lane)

enum ErrBit { kErrSem = 1, kErrNonFinite = 2, kErrRange = 4, kErrMixed = 8, kErrRadius = 16 };
constexpr int kCls = 18;  // classes 2..19
constexpr int kInf18 = 0x7f7f7f7f;  // 'no HP' marker of the per-class tables (memset 0x7f)

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
    unsigned *enc_max_s;  // [3S]
    unsigned *enc_min_o;  // [3S] encoded min of original coords
    unsigned *enc_max_o;  // [3S]
    int *cls;             // [S] class of the segment (class of its first point)
    int *min_pts;         // [S]
    float *r2;            // [S]
    float *inv_h;         // [S]
    float *min_s;         // [3S]
    float *min_o;         // [3S]
    float *inv_g;         // [S]
    int *cc_start;        // [S] first coarse-cell ordinal of each non-empty segment
    int *cc_end;          // [S] one past its last coarse-cell ordinal
    int *id_base;         // [S] global kept-cluster index at which this segment's CALL starts
    int *k_base;          // [S] global kept-cluster index of this segment's first cluster
    int *cluster_num;     // [S]
};

struct Grid {             // two-level cell table (device pointers)
    const float4 *pts4;
    const float *sx, *sy, *sz; // [N+2] sorted coordinates as SoA (k_degree loads query PAIRS with one 64-bit load)
    const int *fcell_of;      // [N] fine-cell ordinal of every sorted point
    const int *row_of;        // [N] coarse-row ordinal of every sorted point
    const int *fcell_start;   // [F+1] first sorted point of every fine cell
    const uint64_t *fcell_key;// [F] full sort key of the fine cell
    const int *fcell_cc;      // [F] coarse-cell ordinal of the fine cell
    const int *cc_pstart;     // [Cc+1] first sorted point of every coarse cell
    const int *cc_fstart;     // [Cc+1] first fine cell of every coarse cell
    const uint64_t *cc_key;   // [Cc] key >> 3
    const int2 *runs9;        // [Cc*9] coarse-cell ordinal ranges of the 9 stencil rows
    const int *d_F;           // number of fine cells
    const int *d_Cc;          // number of coarse cells
};

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint64_t spread3(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

// fine cell coordinates from a full key
__device__ __forceinline__ void key_fine(uint64_t key, int &cx, int &cy, int &cz) {
    cx = (int)(((key >> 3) & kCoarseMax) << 1) | (int)(key & 1);
    cy = (int)(((key >> (3 + kCoarseBits)) & kCoarseMax) << 1) | (int)((key >> 1) & 1);
    cz = (int)(((key >> (3 + 2 * kCoarseBits)) & kCoarseMax) << 1) | (int)((key >> 2) & 1);
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

// K7  coarse stencil rows: runs9[c*9 + (dz+1)*3 + (dy+1)] = [first coarse cell, last+1) with
//     cx'-1 <= x' <= cx'+1 in coarse row (cy'+dy, cz'+dz) — contiguous because cells sort x-fastest
//     Four (cell, row) searches per thread run in lock step: every probe round issues four independent loads instead of
//     one (the kernel is a chain of ~log2(cells of the segment) dependent L2 loads per search).
constexpr int kRunsIlp = 4;
__global__ void __launch_bounds__(256)
k_runs(SegArrays sg, const uint64_t *__restrict__ cc_key, const int *__restrict__ d_Cc, int2 *__restrict__ runs9) {
    const long long total = (long long)(*d_Cc) * kRuns;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * kRunsIlp) {
        uint64_t klo[kRunsIlp], khi[kRunsIlp];
        int lo[kRunsIlp], hi[kRunsIlp], e[kRunsIlp];
        bool ok[kRunsIlp];
#pragma unroll
        for (int u = 0; u < kRunsIlp; u++) {
            const long long t = t0 + u * stride;
            ok[u] = false, lo[u] = hi[u] = e[u] = 0, klo[u] = khi[u] = 0;
            if (t >= total) continue;
            const int c = (int)(t / kRuns), k = (int)(t % kRuns);
            const uint64_t key = cc_key[c];  // seg<<39 | z'<<26 | y'<<13 | x'
            const int s = (int)(key >> (3 * kCoarseBits));
            const int cx = (int)(key & kCoarseMax), cy = (int)((key >> kCoarseBits) & kCoarseMax),
                      cz = (int)((key >> (2 * kCoarseBits)) & kCoarseMax);
            const int ny = cy + (k % 3) - 1, nz = cz + (k / 3) - 1;
            if (ny >= 0 && ny <= kCoarseMax && nz >= 0 && nz <= kCoarseMax) {
                const uint64_t base = ((uint64_t)s << (3 * kCoarseBits)) | ((uint64_t)nz << (2 * kCoarseBits)) |
                                      ((uint64_t)ny << kCoarseBits);
                klo[u] = base | (uint64_t)max(cx - 1, 0);
                khi[u] = base | (uint64_t)min(cx + 1, kCoarseMax);  // inclusive
                lo[u] = sg.cc_start[s], hi[u] = e[u] = sg.cc_end[s];
                ok[u] = true;
            }
        }
        bool busy = true;
        while (busy) {  // first cell with key >= klo, all searches one probe per round
            busy = false;
#pragma unroll
            for (int u = 0; u < kRunsIlp; u++) {
                if (lo[u] < hi[u]) {
                    const int mid = (lo[u] + hi[u]) >> 1;
                    if (__ldg(cc_key + mid) < klo[u]) lo[u] = mid + 1;
                    else hi[u] = mid;
                    busy = true;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kRunsIlp; u++) {
            const long long t = t0 + u * stride;
            if (t >= total) continue;
            int2 out = make_int2(0, 0);
            if (ok[u]) {
                const int first = lo[u];
                int p = first;
                // at most three cells (x'-1, x', x'+1) of that row follow: a short linear walk instead of a second search
                while (p < e[u] && __ldg(cc_key + p) <= khi[u]) p++;
                out = make_int2(first, p);
            }
            runs9[t] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K8  degree (pass A) — the dominant kernel.
//     One warp owns a window of 128 consecutive sorted points and splits it into groups of points of
//     the same coarse ROW.  A group shares one candidate stream: the 9 stencil rows widened to
//     [cx'min-1, cx'max+1] (any superset of the r-ball is valid, the accept test is the exact
//     predicate).  Every candidate is ONE broadcast 16-B load tested against 4 query points per lane.
//     History (profiles/): v1, one load per 32 tests, was LSU write-back bound at 94 %; register tiling made
//     it issue-bound (8 slots per test); v4 uses Blackwell's packed fp32x2 pipe (FADD2/FMUL2/FFMA2): the
//     two ADJACENT sorted points a lane owns form one 64-bit operand, the candidate coordinate is a
//     broadcast scalar operand -> 3 FADD2 + FMUL2 + 2 FFMA2 + 2 FSETP + 2 IADD3 per TWO tests (5 slots per
//     test).  tools/microbench/pipes.cu (profiles/microbench_pipes_r01_v2.txt, _r02.txt; slowest-warp timing, no
//     loop-invariant operands): the one-sided packed candidate loop tops out at 0.485 warp-tests/clk/SM, the scalar
//     form at 0.332, the symmetric loop (round 2, below) at 0.387 / 0.414 with 2 / 3 query pairs per lane.  A packed
//     instruction occupies two issue slots, so the pair test is ISSUE bound: 6 x 2 + 4 = 16 slots per two tests.
// ------------------------------------------------------------------------------------------------
// cnt += (d <= r2) as FSETP.LE + predicated IADD3 (nvcc emits FSETP.GTU + 2 IADD3 for the C expression)
__device__ __forceinline__ void count_le(int &cnt, float d, float r2) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(d), "f"(r2));
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, IEEE round-to-nearest per element) -----------
// Two query points share one instruction; the candidate coordinate enters as a broadcast scalar operand
// (ptxas folds {c,c} into the .F32 operand form).  Per element this is the same rounding sequence as sqd():
// dx = fl(c - q) = -fl(q - c) exactly, and only squares of the differences are used.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// tests candidate q against the query pairs (P pairs = 2P query points); counts go to cnt[2p], cnt[2p+1]
template <int P>
__device__ __forceinline__ void test_candidate(const unsigned long long (&qx)[P], const unsigned long long (&qy)[P],
                                               const unsigned long long (&qz)[P], float4 q, float r2, int (&cnt)[2 * P]) {
    unsigned long long cx = pack2(q.x, q.x), cy = pack2(q.y, q.y), cz = pack2(q.z, q.z);
#pragma unroll
    for (int p = 0; p < P; p++) {
        unsigned long long dx = sub2(cx, qx[p]), dy = sub2(cy, qy[p]), dz = sub2(cz, qz[p]);
        unsigned long long d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float d0, d1;
        unpack2(d, d0, d1);
        count_le(cnt[2 * p], d0, r2);
        count_le(cnt[2 * p + 1], d1, r2);
    }
}

__device__ __forceinline__ unsigned long long ldg_pair(const float *p) {  // 8-byte aligned pair of floats
    unsigned long long v;
    asm("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// One group of up to 128 query points [g0, g0+total).  Lane l of pair p owns the two ADJACENT sorted points
// base + 64p + 2l and +1 (base = g0 rounded down to even), fetched with one 64-bit load per coordinate: the
// loaded register pair is directly the packed operand of FADD2 (no re-packing moves).  Slots outside the group
// get x = NaN: they can never pass the test (their hits would otherwise be credited to the candidates).
//
// SYMMETRIC counting (round 2).  The neighbour relation is symmetric and the predicate is bit-symmetric
// (fl(a-b) = -fl(b-a), only squares are used), so every unordered pair is tested ONCE: a group streams only the
// candidates that come LATER in the sorted order — the rest of its own coarse row up to cell cx'max+1 and the four
// later stencil rows (dz,dy) = (0,+1), (+1,-1), (+1,0), (+1,+1) — and a hit counts for both sides.  The query's
// side is the lane's register counter as before.  The candidate's side is the growth of the lane's counters over
// one candidate (an IADD3 sum), at most 2P <= 6 per lane: four candidates share one register, one byte each (a
// byte's warp sum is <= 192), ONE REDUX.SUM adds it over the warp and lanes 0..3 add their byte to the candidates'
// degrees with one RED.  Pairs inside the group are tested one-sided (every lane against all points of the
// group), which also supplies the self hit the reference counts and subtracts (binary_cuda_functions.cu:88).
// tools/microbench/pipes.cu (profiles/microbench_pipes_r02.txt): the symmetric candidate loop costs 1.17-1.20x
// the one-sided loop per executed test and resolves two ordered pairs per test.
// candidates are streamed: the line a few batches ahead is pulled into L1 while the current batch is tested (the
// first use of a freshly loaded candidate was the kernel's top stall: 37 % of the candidate loads missed L1)
#ifndef PB_DEG_PREFETCH
#define PB_DEG_PREFETCH 64
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) {
    if (PB_DEG_PREFETCH > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// warp sum without the convergence check the intrinsic carries (all 32 lanes are always here)
__device__ __forceinline__ unsigned warp_sum(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
// RED.ADD predicated on a non-zero addend (no branch)
__device__ __forceinline__ void red_add_nz(int *p, unsigned v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.global.add.u32 [%0], %1;\n\t}" ::"l"(p), "r"(v) : "memory");
}

template <int P>
__device__ __forceinline__ int slot_sum(const int (&cnt)[2 * P]) {
    int t = cnt[0] + cnt[1];
#pragma unroll
    for (int s = 2; s < 2 * P; s++) t += cnt[s];
    return t;
}

template <int P, bool SYM>
__device__ __forceinline__ void degree_group(const Grid &g, int g0, int total, int lane, float r2, int jb, int je,
                                             int *__restrict__ deg_sorted, int slice, int nslice) {
    static_assert(!SYM || 2 * P * 32 <= 255, "a candidate's warp-wide hit count must fit one byte of the packed register");
    const float4 *__restrict__ pts4 = g.pts4;
    const int base = g0 & ~1;
    const float qnan = __int_as_float(0x7fc00000);
    unsigned long long qx[P], qy[P], qz[P];
    int cnt[2 * P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        int i = base + 64 * p + 2 * lane;
        const bool in0 = i >= g0 && i < g0 + total, in1 = i + 1 < g0 + total;  // i + 1 > g0 always
        if (i >= g0 + total) i = base;  // stay inside the arrays
        qx[p] = ldg_pair(g.sx + i), qy[p] = ldg_pair(g.sy + i), qz[p] = ldg_pair(g.sz + i);
        if (SYM) {
            float x0, x1;
            unpack2(qx[p], x0, x1);
            qx[p] = pack2(in0 ? x0 : qnan, in1 ? x1 : qnan);
        }
        cnt[2 * p] = cnt[2 * p + 1] = 0;
    }
    unsigned nprev = 0;                                  // -(sum of the slot counters after the previous candidate)
    const unsigned sh = 8 * (lane & 3), lmask = lane < 4 ? 0xffu : 0u;
    int *const dq = deg_sorted + lane;
    // ranges: SYM: k = 4 the rest of the own row, 5..8 the later rows (both sides counted), then k = 9: the group itself
    //         (one-sided); one-sided build: k = 0..8 the nine stencil rows
#pragma unroll 1
    for (int k = SYM ? 4 : 0; k < (SYM ? kRuns + 1 : kRuns); k++) {
        int b = __shfl_sync(kFull, jb, k), e = __shfl_sync(kFull, je, k);
        // batches of 4 candidates; with nslice > 1 (small problems) the batches are dealt round-robin to the
        // nslice warps that share this window, so one long candidate stream is not one warp's latency
        if (SYM && k < kRuns) {
#pragma unroll 1
            for (int j = b + 4 * slice; j < e; j += 4 * nslice) {
                if (j + 4 <= e) {
                    float4 q0 = __ldg(pts4 + j), q1 = __ldg(pts4 + j + 1), q2 = __ldg(pts4 + j + 2), q3 = __ldg(pts4 + j + 3);
                    prefetch_l1(pts4 + j + PB_DEG_PREFETCH);
                    test_candidate<P>(qx, qy, qz, q0, r2, cnt);
                    const unsigned t0 = (unsigned)slot_sum<P>(cnt);
                    test_candidate<P>(qx, qy, qz, q1, r2, cnt);
                    const unsigned t1 = (unsigned)slot_sum<P>(cnt);
                    test_candidate<P>(qx, qy, qz, q2, r2, cnt);
                    const unsigned t2 = (unsigned)slot_sum<P>(cnt);
                    test_candidate<P>(qx, qy, qz, q3, r2, cnt);
                    const unsigned t3 = (unsigned)slot_sum<P>(cnt);
                    // (t0 - prev) + (t1 - t0) << 8 + (t2 - t1) << 16 + (t3 - t2) << 24, as four multiply-adds (mod 2^32)
                    const unsigned packed = t0 * 0xffffff01u + t1 * 0xffff0100u + t2 * 0xff010000u + t3 * 0x01000000u + nprev;
                    nprev = 0u - t3;
                    red_add_nz(dq + j, (warp_sum(packed) >> sh) & lmask);
                } else {
                    for (int jj = j; jj < e; jj++) {
                        test_candidate<P>(qx, qy, qz, __ldg(pts4 + jj), r2, cnt);
                        const unsigned t = (unsigned)slot_sum<P>(cnt);
                        const unsigned hits = warp_sum(t + nprev);
                        nprev = 0u - t;
                        red_add_nz(deg_sorted + jj, lane == 0 ? hits : 0u);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int j = b + 4 * slice; j < e; j += 4 * nslice) {
                if (j + 4 <= e) {
                    float4 q0 = __ldg(pts4 + j), q1 = __ldg(pts4 + j + 1), q2 = __ldg(pts4 + j + 2), q3 = __ldg(pts4 + j + 3);
                    test_candidate<P>(qx, qy, qz, q0, r2, cnt);
                    test_candidate<P>(qx, qy, qz, q1, r2, cnt);
                    test_candidate<P>(qx, qy, qz, q2, r2, cnt);
                    test_candidate<P>(qx, qy, qz, q3, r2, cnt);
                } else {
                    for (int jj = j; jj < e; jj++) test_candidate<P>(qx, qy, qz, __ldg(pts4 + jj), r2, cnt);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2 * P; s++) {
        int i = base + 64 * (s >> 1) + 2 * lane + (s & 1);
        if (i >= g0 && i < g0 + total) {  // binary_cuda_functions.cu:88  ans - 1 (self)
            const int d = cnt[s] - (slice == 0 ? 1 : 0);
            if (!SYM && nslice == 1) deg_sorted[i] = d;
            else if (d) atomicAdd(deg_sorted + i, d);  // deg_sorted zeroed by the host
        }
    }
}

// ---- round 2, experiment: the candidate stream staged through shared memory by the TMA unit --------------
// Every warp owns a ring of kTmaStages x kTmaBatch candidates (16 B each) and one mbarrier per stage.  Lane 0 issues
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` for the next batches of the group's candidate ranges
// while the warp tests the current one out of shared memory (LDS.128 broadcast instead of an LDG that may miss L1); ranges
// shorter than kTmaMin candidates keep the direct loads.  The default for symmetric counting (PB_DEG_TMA=0: direct loads).
#ifndef PB_TMA_BATCH
#define PB_TMA_BATCH 64
#endif
#ifndef PB_TMA_STAGES
#define PB_TMA_STAGES 3
#endif
#ifndef PB_TMA_MIN
#define PB_TMA_MIN 16
#endif
constexpr int kTmaBatch = PB_TMA_BATCH, kTmaStages = PB_TMA_STAGES, kTmaMin = PB_TMA_MIN;
struct alignas(128) TmaRing {
    float4 buf[kTmaStages][kTmaBatch];
    unsigned long long bar[kTmaStages];
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// symmetric counting of one group, candidates through the warp's TMA ring.  `issued` / `consumed` count the batches this
// warp has sent / tested since the kernel began (stage = count % kTmaStages, parity = (count / kTmaStages) & 1).
template <int P>
__device__ __forceinline__ void degree_group_tma(const Grid &g, int g0, int total, int lane, float r2, int jb, int je,
                                                 int *__restrict__ deg_sorted, TmaRing &ring, unsigned &issued,
                                                 unsigned &consumed, int slice, int nslice) {
    // (slice, nslice): this warp's share of the group's candidate stream on small problems — whole batches of the long ranges
    // and quads of the short ones are dealt round-robin to the nslice warps that share the window
    static_assert(2 * P * 32 <= 255, "a candidate's warp-wide hit count must fit one byte of the packed register");
    const float4 *__restrict__ pts4 = g.pts4;
    const int base = g0 & ~1;
    const float qnan = __int_as_float(0x7fc00000);
    unsigned long long qx[P], qy[P], qz[P];
    int cnt[2 * P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        int i = base + 64 * p + 2 * lane;
        const bool in0 = i >= g0 && i < g0 + total, in1 = i + 1 < g0 + total;
        if (i >= g0 + total) i = base;
        qx[p] = ldg_pair(g.sx + i), qy[p] = ldg_pair(g.sy + i), qz[p] = ldg_pair(g.sz + i);
        float x0, x1;
        unpack2(qx[p], x0, x1);
        qx[p] = pack2(in0 ? x0 : qnan, in1 ? x1 : qnan);
        cnt[2 * p] = cnt[2 * p + 1] = 0;
    }
    unsigned nprev = 0;
    const unsigned sh = 8 * (lane & 3), lmask = lane < 4 ? 0xffu : 0u;
    int *const dq = deg_sorted + lane;
    // four candidates q0..q3 at sorted positions j..j+3: tests, byte-packed credit to the candidates (see degree_group)
    auto quad = [&](float4 q0, float4 q1, float4 q2, float4 q3, int j) {
        test_candidate<P>(qx, qy, qz, q0, r2, cnt);
        const unsigned t0 = (unsigned)slot_sum<P>(cnt);
        test_candidate<P>(qx, qy, qz, q1, r2, cnt);
        const unsigned t1 = (unsigned)slot_sum<P>(cnt);
        test_candidate<P>(qx, qy, qz, q2, r2, cnt);
        const unsigned t2 = (unsigned)slot_sum<P>(cnt);
        test_candidate<P>(qx, qy, qz, q3, r2, cnt);
        const unsigned t3 = (unsigned)slot_sum<P>(cnt);
        const unsigned packed = t0 * 0xffffff01u + t1 * 0xffff0100u + t2 * 0xff010000u + t3 * 0x01000000u + nprev;
        nprev = 0u - t3;
        red_add_nz(dq + j, (warp_sum(packed) >> sh) & lmask);
    };
    auto single = [&](float4 q, int jj) {
        test_candidate<P>(qx, qy, qz, q, r2, cnt);
        const unsigned t = (unsigned)slot_sum<P>(cnt);
        const unsigned hits = warp_sum(t + nprev);
        nprev = 0u - t;
        red_add_nz(deg_sorted + jj, lane == 0 ? hits : 0u);
    };
    // ---- the later candidates (ranges 4..8): long ranges through the ring, short ones directly
    // issue cursor (ik, ib, ie) runs ahead of the test cursor (ck, cb, ce); both skip ranges shorter than kTmaMin
    int ik = 3, ib = 0, ie = 0;
    auto issue = [&]() {   // uniform; sends at most one batch
        while (ik < kRuns && ib >= ie) {
            ik++;
            if (ik < kRuns) {
                ib = __shfl_sync(kFull, jb, ik), ie = __shfl_sync(kFull, je, ik);
                ib = ie - ib < kTmaMin ? ie : ib + kTmaBatch * slice;
            }
        }
        if (ik >= kRuns) return;
        const int nb = min(kTmaBatch, ie - ib);
        const unsigned st = issued % kTmaStages;
        if (lane == 0) tma_load_1d(&ring.buf[st][0], pts4 + ib, (unsigned)nb * 16u, &ring.bar[st]);
        ib += kTmaBatch * nslice;
        issued++;
    };
#pragma unroll 1
    for (int s = 0; s < kTmaStages - 1; s++) issue();
#pragma unroll 1
    for (int k = 4; k < kRuns; k++) {
        const int b = __shfl_sync(kFull, jb, k), e = __shfl_sync(kFull, je, k);
        if (e - b < kTmaMin) {
            for (int j = b + 4 * slice; j < e; j += 4 * nslice) {
                if (j + 4 <= e) quad(__ldg(pts4 + j), __ldg(pts4 + j + 1), __ldg(pts4 + j + 2), __ldg(pts4 + j + 3), j);
                else
                    for (int jj = j; jj < e; jj++) single(__ldg(pts4 + jj), jj);
            }
            continue;
        }
#pragma unroll 1
        for (int j0 = b + kTmaBatch * slice; j0 < e; j0 += kTmaBatch * nslice) {
            const int nb = min(kTmaBatch, e - j0);
            const unsigned st = consumed % kTmaStages;
            issue();                                           // keeps kTmaStages - 1 batches in flight behind this one
            mbar_wait(&ring.bar[st], (consumed / kTmaStages) & 1u);
            const float4 *cb = ring.buf[st];
            int i = 0;
            for (; i + 4 <= nb; i += 4) quad(cb[i], cb[i + 1], cb[i + 2], cb[i + 3], j0 + i);
            for (; i < nb; i++) single(cb[i], j0 + i);
            consumed++;
            __syncwarp();                                      // every lane has read the stage before it is sent again
        }
    }
    // ---- range 9: the group itself, one-sided
    {
        const int b = __shfl_sync(kFull, jb, kRuns), e = __shfl_sync(kFull, je, kRuns);
        for (int j = b + 4 * slice; j < e; j += 4 * nslice) {
            if (j + 4 <= e) {
                float4 q0 = __ldg(pts4 + j), q1 = __ldg(pts4 + j + 1), q2 = __ldg(pts4 + j + 2), q3 = __ldg(pts4 + j + 3);
                test_candidate<P>(qx, qy, qz, q0, r2, cnt);
                test_candidate<P>(qx, qy, qz, q1, r2, cnt);
                test_candidate<P>(qx, qy, qz, q2, r2, cnt);
                test_candidate<P>(qx, qy, qz, q3, r2, cnt);
            } else {
                for (int jj = j; jj < e; jj++) test_candidate<P>(qx, qy, qz, __ldg(pts4 + jj), r2, cnt);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2 * P; s++) {
        int i = base + 64 * (s >> 1) + 2 * lane + (s & 1);
        if (i >= g0 && i < g0 + total) {
            const int d = cnt[s] - (slice == 0 ? 1 : 0);  // binary_cuda_functions.cu:88  ans - 1 (self)
            if (d) atomicAdd(deg_sorted + i, d);
        }
    }
}

// one 128-point window; `warp` = window index, (slice, nslice) = this warp's share of the window (small problems)
// phase: -1 = every window; 0 = only the HEAVY windows (at most two groups: the dense blobs, 128 queries against a long
// candidate stream), 1 = only the others.  The grid runs phase 0 first, so the kernel drains on short windows.
template <bool SYM, bool TMA = false>
__device__ __forceinline__ void degree_window(int n, const SegArrays &sg, const Grid &g, int *__restrict__ deg_sorted,
                                              unsigned long long *__restrict__ n_tests, int warp, int slice, int nslice,
                                              int phase, TmaRing *ring = nullptr) {
    unsigned issued = 0, consumed = 0;   // TMA: batches this warp has sent / tested (one window per warp and launch)
    int lane = lane_id();
    long long base = (long long)warp * kWindow;
    if (base >= n) return;
    int end = (int)min((long long)n, base + kWindow);
    // group heads of the window (a group = points of one coarse row): one coalesced read of row_of, four ballots
    unsigned hm[kWindow / 32];
    int G = 0;
#pragma unroll
    for (int k = 0; k < kWindow / 32; k++) {
        int i = (int)base + 32 * k + lane;
        bool head = (i < end) && (i == (int)base || g.row_of[i] != g.row_of[i - 1]);
        hm[k] = __ballot_sync(kFull, head);
        G += __popc(hm[k]);
    }
    if (phase >= 0 && (G <= 2) != (phase == 0)) return;
    // small problems launch nslice warps per window (gridDim.y).  A window of many small groups is dealt out group by
    // group to teams of `cs` warps (every group's dependent metadata loads then run in parallel on different warps
    // instead of back to back on one); the warps of a team — or, for windows of few large groups, all warps — split the
    // group's candidate stream.
    int cs = nslice, gs = 1;
    if (nslice > 1 && G >= nslice) {
        cs = nslice >= 8 ? 4 : (nslice >= 4 ? 2 : 1);
        gs = nslice / cs;
    }
    const bool by_group = gs > 1;
    const int team = slice / cs, sub = slice % cs;
    if (by_group && team >= gs) return;  // nslice not a multiple of cs: the last warps idle
    unsigned long long tests = 0, intra = 0;
    int gi = 0, pos = (int)base;
    while (pos < end) {  // uniform
        // group end: next head after pos
        int gend = end;
        {
            int rel = pos + 1 - (int)base;
#pragma unroll
            for (int k = kWindow / 32 - 1; k >= 0; k--) {
                int lo = rel - 32 * k;  // first candidate bit inside word k
                unsigned m = hm[k];
                if (lo > 0) m = lo >= 32 ? 0u : (m & (0xffffffffu << lo));
                if (m) gend = (int)base + 32 * k + __ffs(m) - 1;
            }
        }
        const bool mine = !by_group || (gi % gs) == team;
        gi++;
        if (mine) {
            int total = gend - pos;
            int fmin = g.fcell_of[pos], fmax = g.fcell_of[gend - 1];
            int cmin = g.fcell_cc[fmin], cmax = g.fcell_cc[fmax];
            const int seg = (int)(g.fcell_key[fmin] >> kSegShift);
            float r2 = sg.r2[seg];
            int jb = 0, je = 0;
            if (lane < kRuns && (!SYM || lane >= 4)) {
                int c0 = g.runs9[(long long)cmin * kRuns + lane].x, c1 = g.runs9[(long long)cmax * kRuns + lane].y;
                if (c1 > c0) {
                    jb = g.cc_pstart[c0];
                    je = g.cc_pstart[c1];
                }
                if (SYM && lane == 4) jb = gend;  // the own row: only what follows the group (its end cell is never empty)
            }
            unsigned cand = __reduce_add_sync(kFull, (unsigned)(je - jb));
            if (SYM && cand == 0 && total == 1) {  // a lone point with an empty later stencil: its only hit would be itself
                pos = gend;
                continue;
            }
            if (SYM && lane == kRuns) jb = pos, je = gend;  // range 9: the group itself, one-sided
            if (n_tests && sub == 0) {
                tests += (unsigned long long)(cand + (SYM ? (unsigned)total : 0u)) * (unsigned)total;
                if (SYM) intra += (unsigned long long)total * (unsigned)total;  // the one-sided share
            }
            const int sl = sub, ns = cs;
            if (TMA) {
                switch ((total + (pos & 1) + 63) >> 6) {  // query pairs per lane
                    case 1: degree_group_tma<1>(g, pos, total, lane, r2, jb, je, deg_sorted, *ring, issued, consumed, sl, ns); break;
                    case 2: degree_group_tma<2>(g, pos, total, lane, r2, jb, je, deg_sorted, *ring, issued, consumed, sl, ns); break;
                    default: degree_group_tma<3>(g, pos, total, lane, r2, jb, je, deg_sorted, *ring, issued, consumed, sl, ns); break;
                }
            } else
            switch ((total + (pos & 1) + 63) >> 6) {  // query pairs per lane
                case 1: degree_group<1, SYM>(g, pos, total, lane, r2, jb, je, deg_sorted, sl, ns); break;
                case 2: degree_group<2, SYM>(g, pos, total, lane, r2, jb, je, deg_sorted, sl, ns); break;
#if PB_DEG_WINDOW > 192
                case 3: degree_group<3, SYM>(g, pos, total, lane, r2, jb, je, deg_sorted, sl, ns); break;
                default: degree_group<4, SYM>(g, pos, total, lane, r2, jb, je, deg_sorted, sl, ns); break;
#else
                default: degree_group<3, SYM>(g, pos, total, lane, r2, jb, je, deg_sorted, sl, ns); break;
#endif
            }
        }
        pos = gend;
    }
    if (n_tests && lane == 0) {
        atomicAdd(n_tests, tests);
        if (SYM) atomicAdd(n_tests + 3, intra);
    }
}

template <bool SYM, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_degree(int n, SegArrays sg, Grid g, int *__restrict__ deg_sorted, unsigned long long *__restrict__ n_tests, int phased) {
    // phased: the grid holds every window twice, heavy pass first
    const int nb = phased ? gridDim.x >> 1 : gridDim.x;
    const int phase = phased ? (blockIdx.x >= nb ? 1 : 0) : -1;
    const int bx = blockIdx.x - (phase == 1 ? nb : 0);
    degree_window<SYM>(n, sg, g, deg_sorted, n_tests, (bx * blockDim.x + threadIdx.x) >> 5, blockIdx.y, gridDim.y, phase);
}

// the TMA-staged variant of the symmetric counting
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_degree_tma(int n, SegArrays sg, Grid g, int *__restrict__ deg_sorted, unsigned long long *__restrict__ n_tests, int phased) {
    static_assert(PB_DEG_WINDOW <= 192, "degree_group_tma is instantiated for up to three query pairs per lane");
    __shared__ TmaRing rings[4];
    TmaRing &ring = rings[threadIdx.x >> 5];
    if (lane_id() == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; s++) mbar_init(&ring.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nb = phased ? gridDim.x >> 1 : gridDim.x;
    const int phase = phased ? (blockIdx.x >= nb ? 1 : 0) : -1;
    const int bx = blockIdx.x - (phase == 1 ? nb : 0);
    degree_window<true, true>(n, sg, g, deg_sorted, n_tests, (bx * blockDim.x + threadIdx.x) >> 5, blockIdx.y, gridDim.y, phase, &ring);
}

// K9  HP rule + per-cell HP statistics + degree scatter to input order.  A thread owns kHpPer points (32 consecutive
//     points per warp and round): the loads of all rounds are issued before the first round is processed, so the
//     three-deep chain of dependent loads (cell -> key -> segment table) is paid once per kHpPer points.
constexpr int kHpPer = 4;
template <bool MIXED>
__global__ void __launch_bounds__(256)
k_hp_cells(int n, SegArrays sg, float4 *__restrict__ pts4, const int *__restrict__ fcell_of,
           const uint64_t *__restrict__ fcell_key, const int *__restrict__ deg_sorted,
           int *__restrict__ degree_out, int *__restrict__ cell_hp, int *__restrict__ cell_minhp,
           unsigned long long *__restrict__ counters, const int *__restrict__ sem,
           const int *__restrict__ min_pts_tab, int *__restrict__ cell_min18, int *__restrict__ cell_first) {
    const int lane = lane_id();
    const int warp0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 * kHpPer);  // first point of the warp
    int cc[kHpPer], dd[kHpPer], oo[kHpPer], ss[kHpPer];
#pragma unroll
    for (int k = 0; k < kHpPer; k++) {
        const int i = warp0 + 32 * k + lane;
        cc[k] = dd[k] = oo[k] = 0;
        if (i < n) cc[k] = fcell_of[i], dd[k] = deg_sorted[i], oo[k] = reinterpret_cast<const int *>(pts4 + i)[3];
    }
#pragma unroll
    for (int k = 0; k < kHpPer; k++) ss[k] = (warp0 + 32 * k + lane < n) ? (int)(fcell_key[cc[k]] >> kSegShift) : 0;
    unsigned long long dsum = 0, hsum = 0;
#pragma unroll
    for (int k = 0; k < kHpPer; k++) {
        const int i = warp0 + 32 * k + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(kFull, valid);
        if (!act) break;       // uniform
        if (!valid) continue;  // only the last warp: the survivors use `act` as their mask below
        const int c = cc[k], d = dd[k], orig = oo[k];
        bool hp;  // binary_cuda_functions.cu:175-186
        if (!MIXED) {
            hp = d >= sg.min_pts[ss[k]];
        } else {
            int myc = sem[orig] - 2;
            hp = d >= min_pts_tab[myc];
            if (hp) atomicMin(cell_min18 + (long long)c * kCls + myc, orig);
        }
        degree_out[orig] = d;
        if (hp) reinterpret_cast<int *>(pts4 + i)[3] = orig | kHpBit;
        unsigned grp = __match_any_sync(act, c);
        unsigned hpm = __ballot_sync(act, hp) & grp;
        int mn = __reduce_min_sync(grp, hp ? orig : 0x7fffffff);
        int first = __reduce_min_sync(grp, hp ? i : 0x7fffffff);
        if (hpm && lane == __ffs(grp) - 1) {
            // a cell whose points all sit inside this warp round is written with plain stores (the tables start at their
            // neutral values); only the runs that touch the first / last active lane may continue in a neighbouring
            // round and need atomics
            const bool interior = !(grp & 1u) && !(grp & (1u << (31 - __clz(act))));
            if (interior) {
                cell_hp[c] = __popc(hpm);
                cell_minhp[c] = mn;
                cell_first[c] = first;
            } else {
                atomicAdd(cell_hp + c, __popc(hpm));
                atomicMin(cell_minhp + c, mn);
                atomicMin(cell_first + c, first);  // sorted position of the cell's first HP: its representative in k_union
            }
        }
        if (counters) {  // profiling only: [1] sum of degrees, [2] HP count
            dsum += __reduce_add_sync(act, (unsigned)d);
            hsum += __popc(__ballot_sync(act, hp));
        }
    }
    if (counters && lane == 0) {
        atomicAdd(counters + 1, dsum);
        atomicAdd(counters + 2, hsum);
    }
}

// Chebyshev distance (in fine cells) between the cells with keys ka, kb; <= 2 means inside the 5x5x5 stencil
__device__ __forceinline__ int fine_dist(uint64_t ka, uint64_t kb) {
    int ax, ay, az, bx, by, bz;
    key_fine(ka, ax, ay, az);
    key_fine(kb, bx, by, bz);
    return max(max(abs(ax - bx), abs(ay - by)), abs(az - bz));
}
__device__ __forceinline__ bool fine_near(uint64_t ka, uint64_t kb) { return fine_dist(ka, kb) <= 2; }

// cooperative search for ONE HP pair (a in A, b in B) within r; warp-uniform result
__device__ __forceinline__ bool hp_pair_exists(const float4 *__restrict__ pts4, int a0, int a1, int b0, int b1,
                                               float r2, int lane) {
    for (int ia = a0; ia < a1; ia += 32) {
        int i = ia + lane;
        float4 p = pts4[i < a1 ? i : a0];
        bool php = (i < a1) && (__float_as_int(p.w) & kHpBit);
        if (!__any_sync(kFull, php)) continue;
        for (int j = b0; j < b1; j++) {
            float4 q = __ldg(pts4 + j);
            if (!(__float_as_int(q.w) & kHpBit)) continue;
            bool hit = php && (sqd(p.x, p.y, p.z, q.x, q.y, q.z) <= r2);
            if (__any_sync(kFull, hit)) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// K10  HP connectivity (pass B): union-find over FINE cells.  One warp per HP-cell A; the lanes check
//      32 stencil cells at a time (HP-bearing, inside the 5^3 fine stencil, ordinal > A, different
//      root) and only the survivors pay a pair search that stops at the first HP pair within r.
//      Launched twice: touching cells first, then the cells at fine distance 2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_union(SegArrays sg, Grid g, const float4 *__restrict__ pts4, const int *__restrict__ cell_hp, int *parent, int far_pass,
        const int *__restrict__ cell_first) {
    int F = *g.d_F;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int A = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; A < F; A += warps) {
        if (cell_hp[A] == 0) continue;
        uint64_t kA = g.fcell_key[A];
        float r2 = sg.r2[(int)(kA >> kSegShift)];
        int a0 = g.fcell_start[A], a1 = g.fcell_start[A + 1];
        int cA = g.fcell_cc[A];
        int f0 = 0, f1 = 0;
        if (lane < kRuns) {
            int2 rr = g.runs9[(long long)cA * kRuns + lane];
            if (rr.y > rr.x) {
                f0 = g.cc_fstart[rr.x];
                f1 = g.cc_fstart[rr.y];
            }
        }
        int rootA = uf_find(parent, A);
        const float4 pA = __ldg(pts4 + cell_first[A]);  // representative: the cell's first HP
        // only the stencil rows that hold a cell with a larger ordinal (every cell pair is examined from its smaller side)
        for (unsigned rows = __ballot_sync(kFull, f1 > max(f0, A + 1)); rows; rows &= rows - 1) {
            const int k = __ffs(rows) - 1;
            int b = __shfl_sync(kFull, f0, k), e = __shfl_sync(kFull, f1, k);
            for (int fb = max(b, A + 1); fb < e; fb += 32) {
                int B = fb + lane;
                bool cand = false;
                if (B < e && cell_hp[B] > 0) {
                    // pass 0 joins touching cells (almost always an immediate hit); pass 1 then finds most of
                    // the distance-2 cells already in the same set and skips their (expensive) pair search
                    int d = fine_dist(kA, g.fcell_key[B]);
                    if (far_pass ? d == 2 : d <= 1) cand = uf_find(parent, B) != rootA;
                }
                // representative shortcut: every lane tests ONE exact pair (first HP of A, first HP of its cell B); a hit
                // joins the two cells right away (lock-free, lane-parallel), only the misses pay the cooperative search
                bool quick = false;
                if (cand) {
                    float4 q = __ldg(pts4 + cell_first[B]);
                    quick = sqd(pA.x, pA.y, pA.z, q.x, q.y, q.z) <= r2;
                    if (quick) {
                        uf_union(parent, A, B);
                        cand = false;
                    }
                }
                if (__any_sync(kFull, quick)) rootA = uf_find(parent, A);
                unsigned m = __ballot_sync(kFull, cand);
                while (m) {
                    int l = __ffs(m) - 1;
                    m &= m - 1;
                    int Bc = fb + l;
                    int same = 0;
                    if (lane == 0) same = (uf_find(parent, A) == uf_find(parent, Bc));
                    if (__shfl_sync(kFull, same, 0)) continue;
                    if (hp_pair_exists(pts4, a0, a1, g.fcell_start[Bc], g.fcell_start[Bc + 1], r2, lane)) {
                        if (lane == 0) uf_union(parent, A, Bc);
                        __syncwarp();
                        rootA = uf_find(parent, A);
                    }
                }
            }
        }
    }
}

// K11  flatten + minimum HP index of every component
//      (MIXED: one minimum per (component, class): a cluster is a component restricted to one class that
//      owns an HP in it — binary.cu:206-213 labels only visited points of the seed's class)
template <bool MIXED>
__global__ void k_comp_min(const int *__restrict__ d_F, const int *__restrict__ cell_hp, int *parent,
                           const int *__restrict__ cell_minhp, int *comp_min, const int *__restrict__ cell_min18,
                           int *comp_min18) {
    int F = *d_F;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < F; c += gridDim.x * blockDim.x) {
        if (cell_hp[c] == 0) continue;
        int r = uf_find(parent, c);
        if (r != c) __stcg(parent + c, r);
        if (!MIXED) {
            atomicMin(comp_min + r, cell_minhp[c]);
        } else {
            for (int k = 0; k < kCls; k++) {
                int v = cell_min18[(long long)c * kCls + k];
                if (v != kInf18) atomicMin(comp_min18 + (long long)r * kCls + k, v);
            }
        }
    }
}

// K12  flag the minimum-index HP of every component (cluster numbering = rank of that index,
//      binary.cu:161-166: seeds are taken in ascending point order)
template <bool MIXED>
__global__ void k_flag_roots(const int *__restrict__ d_F, const int *__restrict__ cell_hp,
                             const int *__restrict__ parent, const int *__restrict__ comp_min,
                             int *__restrict__ flag, const int *__restrict__ comp_min18) {
    int F = *d_F;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < F; c += gridDim.x * blockDim.x) {
        if (!(cell_hp[c] > 0 && parent[c] == c)) continue;
        if (!MIXED) {
            flag[comp_min[c]] = 1;
        } else {
            for (int k = 0; k < kCls; k++) {
                int v = comp_min18[(long long)c * kCls + k];
                if (v != kInf18) flag[v] = 1;
            }
        }
    }
}

// K13  raw cluster id of every HP-cell; representative point of every raw cluster
template <bool MIXED>
__global__ void k_cell_gid(const int *__restrict__ d_F, const int *__restrict__ cell_hp,
                           const int *__restrict__ parent, const int *__restrict__ comp_min,
                           const int *__restrict__ gid_at, int *__restrict__ cell_gid, int *__restrict__ rep,
                           const int *__restrict__ comp_min18, int *__restrict__ cell_gid18) {
    int F = *d_F;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < F; c += gridDim.x * blockDim.x) {
        if (!MIXED) {
            int gi = -1;
            if (cell_hp[c] > 0) {
                int r = parent[c];
                int u = comp_min[r];
                gi = gid_at[u];
                if (r == c) rep[gi] = u;
            }
            cell_gid[c] = gi;
        } else {
            int r = cell_hp[c] > 0 ? parent[c] : -1;
            for (int k = 0; k < kCls; k++) {
                int gi = -1;
                if (r >= 0) {
                    int u = comp_min18[(long long)r * kCls + k];
                    if (u != kInf18) {
                        gi = gid_at[u];
                        if (r == c) rep[gi] = u;
                    }
                }
                cell_gid18[(long long)c * kCls + k] = gi;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K14  labels (pass C).  HPs take their component's raw id; a border LP takes the MAXIMUM raw id
//      among components owning an HP within r (later BFS overwrites earlier, binary.cu:206-213);
//      LPs with no HP neighbour stay -1.  Cluster sizes (incl. border LPs) are counted here.
// ------------------------------------------------------------------------------------------------
template <bool MIXED>
__global__ void __launch_bounds__(128)
k_label(int n, SegArrays sg, Grid g, const float4 *__restrict__ pts4, const int *__restrict__ cell_hp,
        const int *__restrict__ cell_gid, int *__restrict__ raw_label, int *__restrict__ raw_count,
        const int *__restrict__ sem, const int *__restrict__ cell_gid18, int ppw) {
    // ppw = points per warp: 32 on large problems; small problems spread the points over more warps (the loop below
    // visits the distinct cells of a warp's LPs one after the other, each with its own chain of dependent loads)
    int lane = lane_id();
    int i = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ppw + lane;
    bool valid = lane < ppw && i < n;
    float4 p = pts4[valid ? i : 0];
    int c = valid ? g.fcell_of[i] : -1;
    bool hp = valid && (__float_as_int(p.w) & kHpBit);
    int label = -1;
    int myc = 0;
    if (MIXED && valid) myc = sem[__float_as_int(p.w) & ~kHpBit] - 2;
    if (hp) label = MIXED ? cell_gid18[(long long)c * kCls + myc] : cell_gid[c];
    unsigned todo = __ballot_sync(kFull, valid && !hp);
    // every LP lane fetches the header of its OWN cell up front (key, radius, coarse cell: a chain of dependent loads,
    // now paid once per warp); the loop below broadcasts the leader's copy
    uint64_t kMine = 0;
    float r2Mine = 0.f;
    int ccMine = 0;
    if (valid && !hp) {
        kMine = g.fcell_key[c];
        r2Mine = sg.r2[(int)(kMine >> kSegShift)];
        ccMine = g.fcell_cc[c];
    }
    while (todo) {
        int leader = __ffs(todo) - 1;
        int cL = __shfl_sync(kFull, c, leader);
        bool mine = (c == cL) && valid && !hp;
        unsigned mask = __ballot_sync(kFull, mine);
        todo &= ~mask;
        uint64_t kL = __shfl_sync(kFull, kMine, leader);
        float r2 = __shfl_sync(kFull, r2Mine, leader);
        int cc = __shfl_sync(kFull, ccMine, leader);
        int f0 = 0, f1 = 0;
        if (lane < kRuns) {
            int2 rr = g.runs9[(long long)cc * kRuns + lane];
            if (rr.y > rr.x) {
                f0 = g.cc_fstart[rr.x];
                f1 = g.cc_fstart[rr.y];
            }
        }
        int best = -1;
        for (unsigned rows = __ballot_sync(kFull, f1 > f0); rows; rows &= rows - 1) {   // non-empty stencil rows only
            const int k = __ffs(rows) - 1;
            int b = __shfl_sync(kFull, f0, k), e = __shfl_sync(kFull, f1, k);
            for (int fb = b; fb < e; fb += 32) {
                int B = fb + lane;
                int gB = -1;
                bool okB = B < e && cell_hp[B] > 0 && fine_near(kL, g.fcell_key[B]);
                if (okB) gB = MIXED ? 0x7ffffffe : cell_gid[B];  // MIXED: the id depends on the query's class
                int bestmin = __reduce_min_sync(kFull, mine ? best : 0x7fffffff);  // smallest best among my LPs
                unsigned m = __ballot_sync(kFull, gB > bestmin);
                while (m) {
                    int l = __ffs(m) - 1;
                    m &= m - 1;
                    int gc = __shfl_sync(kFull, gB, l);
                    if (MIXED) gc = mine ? cell_gid18[(long long)(fb + l) * kCls + myc] : -1;
                    bool pend = mine && best < gc;
                    if (!__any_sync(kFull, pend)) continue;
                    int Bc = fb + l;
                    int b0 = g.fcell_start[Bc], b1 = g.fcell_start[Bc + 1];
                    const unsigned pm = __ballot_sync(kFull, pend);
                    if (__popc(pm) <= 4) {
                        // few LPs wait for this cell (the usual case: border LPs are sparse): the LANES take the cell's
                        // points, 32 per trip, and the LPs are served one after the other — a lone LP next to a full
                        // cell costs two trips instead of a serial walk over its ~45 points
                        for (unsigned m2 = pm; m2; m2 &= m2 - 1) {
                            const int l2 = __ffs(m2) - 1;
                            const float lx = __shfl_sync(kFull, p.x, l2), ly = __shfl_sync(kFull, p.y, l2), lz = __shfl_sync(kFull, p.z, l2);
                            bool hit = false;
                            for (int j0 = b0; j0 < b1 && !hit; j0 += 32) {
                                const int j = j0 + lane;
                                bool h = false;
                                if (j < b1) {
                                    float4 q = __ldg(pts4 + j);
                                    h = (__float_as_int(q.w) & kHpBit) && sqd(lx, ly, lz, q.x, q.y, q.z) <= r2;
                                }
                                hit = __any_sync(kFull, h);
                            }
                            if (hit && lane == l2) best = gc, pend = false;
                        }
                        continue;
                    }
                    for (int j = b0; j < b1; j++) {
                        float4 q = __ldg(pts4 + j);
                        if (!(__float_as_int(q.w) & kHpBit)) continue;
                        if (pend && sqd(p.x, p.y, p.z, q.x, q.y, q.z) <= r2) {
                            best = gc;
                            pend = false;
                        }
                        if (!__any_sync(kFull, pend)) break;
                    }
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

// MIXED: first labelled-list position of every (segment, class) block of order2
__global__ void k_seg_lab18(int n, int S, const int *__restrict__ pos18, const int *__restrict__ lpos,
                            const int *__restrict__ d_L, int *__restrict__ lab_start18) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > S * kCls) return;
    int pos = e < S * kCls ? pos18[e] : n;
    lab_start18[e] = pos < n ? lpos[pos] : *d_L;
}

// ------------------------------------------------------------------------------------------------
// K20  LP assignment (binary.cu:270-358, binary_cuda_functions.cu:258-302): exact 1-NN over the
//      labelled points of the same segment in ORIGINAL coordinates, ties -> largest index.
//      One warp per query, branch-and-bound over a two-level box hierarchy of the spatially sorted
//      labelled list.  The first guess is the 32-point group at the query's own rank in that order;
//      afterwards a box is skipped only if its (conservatively shrunk) lower bound is STRICTLY
//      greater than the best distance, so equal-distance ties are never lost.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float box_lb(float px, float py, float pz, float4 lo, float4 hi) {
    float dx = fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f);
    float dy = fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f);
    float dz = fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f);
    return (dx * dx + dy * dy + dz * dz) * (1.0f - 1e-5f);
}

__device__ __forceinline__ void nn_scan_group(int gi, int l0, int l1, int lane, float px, float py, float pz,
                                              const float4 *__restrict__ lab4, float &bestD, int &bestI) {
    int j = gi * 32 + lane;
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

template <bool MIXED>
__global__ void __launch_bounds__(256)
k_nn(int n, const int *__restrict__ d_Q, const int *__restrict__ d_L, SegArrays sg, const int *__restrict__ qlist, const int *__restrict__ seg_of,
     const int *__restrict__ inv2, const int *__restrict__ lpos, const float *__restrict__ xo,
     const float *__restrict__ yo, const float *__restrict__ zo, const float4 *__restrict__ lab4,
     const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi, const float4 *__restrict__ box2_lo,
     const float4 *__restrict__ box2_hi, int *cluster_id, const int *__restrict__ sem,
     const int *__restrict__ lab_start18, const int *__restrict__ seg_lastlab) {
    int Q = *d_Q;
    const int L = *d_L;
    int lane = lane_id();
    int warps = (gridDim.x * blockDim.x) >> 5;
    // a warp takes bs <= 32 consecutive queries at a time: lane l fetches the header of query bs*qb + l (list entry, segment,
    // labelled range, coordinates, own rank — a chain of three dependent loads) for all 32 at once; the queries are then
    // processed one after the other with the header broadcast by shuffles
    // (bs = queries per warp and trip: 32 when there are enough queries to keep every warp busy that way, fewer on small
    // problems, where spreading the queries over the warps matters more than batching the header loads)
    const int bs = max(1, min(32, Q / warps));
    for (int qb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qb * bs < Q; qb += warps) {
      int hp_ = -1, hl0 = 0, hl1 = 0, hrank = 0, hlast = -1;
      float hx = 0.f, hy = 0.f, hz = 0.f;
      if (lane < bs && qb * bs + lane < Q) {
          hp_ = qlist[qb * bs + lane];
          int s = seg_of[hp_];
          if (!MIXED) {
              // labelled range of the segment: lpos (exclusive count of labelled points per sorted position) at its borders
              int b0 = sg.start[s], b1 = sg.start[s + 1];
              hl0 = lpos[b0], hl1 = b1 < n ? lpos[b1] : L;
          } else {  // candidates = labelled points of the query's class (binary_cuda_functions.cu:275)
              long long e = (long long)s * kCls + (sem[hp_] - 2);
              hl0 = lab_start18[e], hl1 = lab_start18[e + 1];
              hlast = seg_lastlab[s];
          }
          hx = xo[hp_], hy = yo[hp_], hz = zo[hp_];
          hrank = lpos[inv2[hp_]];
      }
      const int nq = min(bs, Q - qb * bs);
      for (int j = 0; j < nq; j++) {
        const int p = __shfl_sync(kFull, hp_, j);
        const int l0 = __shfl_sync(kFull, hl0, j), l1 = __shfl_sync(kFull, hl1, j);
        if (l1 <= l0) {
            if (MIXED) {  // no labelled point of this class: the label of the LAST labelled point (:287-300)
                int last = __shfl_sync(kFull, hlast, j);
                if (lane == 0 && last >= 0) cluster_id[p] = cluster_id[last];
            }
            continue;
        }
        const float px = __shfl_sync(kFull, hx, j), py = __shfl_sync(kFull, hy, j), pz = __shfl_sync(kFull, hz, j);
        int g0 = l0 >> 5, g1 = (l1 - 1) >> 5;
        // first guess: the group where the query itself would sit in the sorted labelled list
        int gq = min(max(__shfl_sync(kFull, hrank, j), l0), l1 - 1) >> 5;
        float bestD = __int_as_float(0x7f800000);
        int bestI = -1;
        nn_scan_group(gq, l0, l1, lane, px, py, pz, lab4, bestD, bestI);
        // level 2 sweep (1024 points per box), descending into level 1 only where the bound allows
        for (int hb = g0 >> 5; hb <= (g1 >> 5); hb += 32) {
            int h = hb + lane;
            float lb2 = __int_as_float(0x7f800000);
            if (h <= (g1 >> 5)) lb2 = box_lb(px, py, pz, __ldg(box2_lo + h), __ldg(box2_hi + h));
            unsigned m2 = __ballot_sync(kFull, lb2 <= bestD);
            while (m2) {
                int hl = __ffs(m2) - 1;
                m2 &= m2 - 1;
                if (__shfl_sync(kFull, lb2, hl) > bestD) continue;
                int gi = (hb + hl) * 32 + lane;
                float lb = __int_as_float(0x7f800000);
                if (gi >= g0 && gi <= g1 && gi != gq) lb = box_lb(px, py, pz, __ldg(box_lo + gi), __ldg(box_hi + gi));
                unsigned m = __ballot_sync(kFull, lb <= bestD);
                while (m) {
                    // visit the most promising group first: it tightens the bound for the others
                    unsigned lbmin = __reduce_min_sync(kFull, (m >> lane) & 1 ? __float_as_uint(lb) : 0xffffffffu);
                    if (__uint_as_float(lbmin) > bestD) break;
                    unsigned who = __ballot_sync(kFull, ((m >> lane) & 1) && __float_as_uint(lb) == lbmin);
                    int gl = __ffs(who) - 1;
                    m &= ~(1u << gl);
                    nn_scan_group((hb + hl) * 32 + gl, l0, l1, lane, px, py, pz, lab4, bestD, bestI);
                }
            }
        }
        if (lane == 0 && bestI >= 0) cluster_id[p] = cluster_id[bestI];
      }
    }
}

// ------------------------------------------------------------------------------------------------
// K21  cluster centres (binary_cuda_functions.cu:217-246): sequential running mean in ascending
//      point order, M += (x - M) / n with IEEE fp32 division — replayed exactly.  One warp per cluster
//      (clusters are handed out through an atomic counter, so a warp that drew a small cluster takes the
//      next one): the warp streams its segment, compacts the member coordinates into a shared-memory
//      buffer and lanes 0,1,2 then run the x, y, z recurrences side by side (one instruction stream for
//      all three chains).  The kernel is bound by the latency of that chain for the largest cluster.
//      A warp issues in order, so every instruction of div.rn.f32 (MUFU.RCP, Newton steps, range check:
//      ~240 cycles per member measured) sits on that chain.  div_by_count() keeps only what depends on
//      the running mean on it: with y = RN(1/n) prepared by the other lanes (__frcp_rn),
//          q0 = a*y;  r0 = fma(-n,q0,a);  q1 = fma(r0,y,q0);  r1 = fma(-n,q1,a);  q2 = fma(r1,y,q1)
//      q1 is a faithful quotient, hence r1 is exact and q2 = RN(a/n) (Markstein's theorem; its one
//      exception, an all-ones divisor significand, is n = 2^24-1, which takes the plain division, as do
//      dividends outside [1e-30, 1e30]).  pb_selftest_division compares it with div.rn.f32 on the device.
// ------------------------------------------------------------------------------------------------
constexpr int kCtrWarps = 4;
constexpr int kCtrBuf = 256;  // buffered members per warp (3 floats each)

__device__ __forceinline__ float div_by_count(float a, float fn, float y) {
    float aa = fabsf(a);
    if (!(aa > 1e-30f && aa < 1e30f)) return __fdiv_rn(a, fn);
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-fn, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-fn, q, a);
    return __fmaf_rn(r, y, q);
}

// device self-test: div_by_count vs div.rn.f32 on pseudo-random and adversarial (dividend, count) pairs
__global__ void k_selftest_division(unsigned long long n_samples, unsigned long long seed,
                                    unsigned long long *__restrict__ mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long h = (i + seed) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32; h *= 0x94D049BB133111EBULL; h ^= h >> 29;
        unsigned lo = (unsigned)h, hi = (unsigned)(h >> 32);
        int n = 1 + (int)(hi % ((1u << 24) - 2));            // 1 .. 2^24-2
        if ((i & 7) == 1) n = 1 + (int)(hi % 4096u);         // small counts dominate real clusters
        float fn = (float)n;
        float a;
        unsigned mant = lo & 0x7fffffu, sign = lo & 0x80000000u;
        int e = 127 - 40 + (int)((lo >> 23) % 60u);          // magnitudes 2^-40 .. 2^19
        a = __uint_as_float(sign | ((unsigned)e << 23) | mant);
        if ((i & 3) == 2) {                                  // adversarial: quotient next to a float / a midpoint
            float m = __uint_as_float(((unsigned)(127 - 20 + (int)(hi % 30u)) << 23) | mant);
            a = __fmul_rn(m, fn);
            if (i & 4) a = __uint_as_float(__float_as_uint(a) + ((lo >> 30) & 1u ? 1u : 0xffffffffu));
        }
        float want = __fdiv_rn(a, fn);
        float got = div_by_count(a, fn, __frcp_rn(fn));
        bad += __float_as_uint(want) != __float_as_uint(got);
    }
    if (bad) atomicAdd(mismatches, bad);
}

// Block-level formulation (round 2): a block of 8 warps draws clusters from the ticket counter.  Warps 1-7 scan the
// cluster's segment chunk by chunk and compact the members' coordinates (ascending point order) and the reciprocals of
// their running counts into one half of a double buffer while warp 0 replays the recurrence over the other half — the
// chain never waits for global memory, and the scan of a 27 k-point segment takes 25 trips instead of 215.
constexpr int kCtrGatherWarps = 7;
constexpr int kCtrChunkPts = kCtrGatherWarps * 160;   // points scanned per chunk (5 per gather lane)

struct alignas(16) CtrSmem {
    float buf[2][3][kCtrChunkPts];
    float rcp[2][kCtrChunkPts];
    int wcount[kCtrGatherWarps];
    int m[2];       // members in each half
    int cluster;    // ticket
};

__device__ __forceinline__ void ctr_gather(CtrSmem &sm, int half, int ub, int e, int local, int cnt_before,
                                           const int *__restrict__ cluster_id, const float *__restrict__ x,
                                           const float *__restrict__ y, const float *__restrict__ z) {
    // called by warps 1..7; warp gw owns points [ub + gw*160, +160)
    const int lane = lane_id(), gw = (threadIdx.x >> 5) - 1;
    int id[5];
    float vx[5], vy[5], vz[5];
    unsigned mk[5];
    int total = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        int u = ub + gw * 160 + j * 32 + lane;
        bool ok = u < e;
        id[j] = ok ? cluster_id[u] : -2;
        vx[j] = ok ? x[u] : 0.f, vy[j] = ok ? y[u] : 0.f, vz[j] = ok ? z[u] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        mk[j] = __ballot_sync(kFull, id[j] == local);
        total += __popc(mk[j]);
    }
    if (lane == 0) sm.wcount[gw] = total;
    asm volatile("bar.sync 1, 224;" ::: "memory");   // the seven gather warps only
    int off = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kCtrGatherWarps; k++) {
        int c = sm.wcount[k];
        off += k < gw ? c : 0;
        all += c;
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        if (id[j] == local) {
            int o = off + __popc(mk[j] & ((1u << lane) - 1u));
            sm.buf[half][0][o] = vx[j], sm.buf[half][1][o] = vy[j], sm.buf[half][2][o] = vz[j];
            sm.rcp[half][o] = __frcp_rn((float)(cnt_before + o + 1));
        }
        off += __popc(mk[j]);
    }
    if (gw == 0 && lane == 0) sm.m[half] = all;
    asm volatile("bar.sync 1, 224;" ::: "memory");   // wcount may be rewritten by the next chunk only after everyone read it
}

// lanes 0..2 of the calling warp replay x, y, z side by side over one half of the buffer
// returns how the half was settled (2: one-correction quotient chain, 3: div.rn.f32)
__device__ __forceinline__ int ctr_chain(const CtrSmem &sm, int half, int fill, int &cnt, float &M, int coord) {
    const float *src = sm.buf[half][coord];
    const float *myrcp = sm.rcp[half];
    bool fast = cnt + fill < (1 << 24) - 1;
    if (fast) {
        // The recurrence runs on the ONE-correction quotient q1 (FSUB, FMUL, 2 FFMA, FADD on the chain: 5 dependent
        // operations per member instead of 7); the second Markstein correction q2 = RN(a/n) is computed next to it, off
        // the chain, and compared: if they ever differ (rare: q1 is already the correctly rounded quotient almost always),
        // or the range guard of div_by_count trips, the half is replayed with div.rn.f32.  Results are bit-identical to
        // M += (x - M) / n with IEEE division either way.
        float M0 = M;
        bool odd = false;
        auto step = [&](float v, float y, float fn) {
            float a = __fsub_rn(v, M);
            float q = __fmul_rn(a, y);
            float r = __fmaf_rn(-fn, q, a);
            float q1 = __fmaf_rn(r, y, q);
            M = __fadd_rn(M, q1);
            float r1 = __fmaf_rn(-fn, q1, a);
            float q2 = __fmaf_rn(r1, y, q1);
            float aa = fabsf(a);
            odd |= (q2 != q1) | (!(aa > 1e-30f && aa < 1e30f) && a != 0.f);
        };
        // blocks of four members; the NEXT block's coordinates and reciprocals are fetched (two 16-byte shared loads) while
        // this one runs, so the chain never waits for shared memory (counts stay below 2^24: float increments are exact)
        float fn = (float)cnt;
        int t = 0;
        float4 v4 = *reinterpret_cast<const float4 *>(src), y4 = *reinterpret_cast<const float4 *>(myrcp);
        for (; t + 4 <= fill; t += 4) {
            const int tn = min(t + 4, kCtrChunkPts - 4);
            const float4 nv = *reinterpret_cast<const float4 *>(src + tn), ny = *reinterpret_cast<const float4 *>(myrcp + tn);
            step(v4.x, y4.x, fn + 1.f);
            step(v4.y, y4.y, fn + 2.f);
            step(v4.z, y4.z, fn + 3.f);
            step(v4.w, y4.w, fn + 4.f);
            fn += 4.f;
            v4 = nv, y4 = ny;
        }
        if (t < fill) step(v4.x, y4.x, fn + 1.f);
        if (t + 1 < fill) step(v4.y, y4.y, fn + 2.f);
        if (t + 2 < fill) step(v4.z, y4.z, fn + 3.f);
        if (!odd) {
            cnt += fill;
            return 2;
        }
        M = M0;
    }
    for (int t = 0; t < fill; t++) {
        float v = src[t];
        cnt++;
        M = __fadd_rn(M, __fdiv_rn(__fsub_rn(v, M), (float)cnt));
    }
    return 3;
}

constexpr int kCtrBigSegment = 8192;  // points
// whole block (8 warps = 256 threads) must call
__device__ __forceinline__ void centres_block(int K, const SegArrays &sg, const int *__restrict__ clt_seg,
                                              const int *__restrict__ cluster_id, const float *__restrict__ x,
                                              const float *__restrict__ y, const float *__restrict__ z,
                                              float *__restrict__ center, int *__restrict__ next_cluster, CtrSmem &sm,
                                              float *__restrict__ center_head = nullptr, int n_head = 0,
                                              unsigned long long *__restrict__ stats = nullptr) {
    // center_head: optional second copy of the first n_head centres (the small-call kernel keeps it next to its result
    // block so that one read-back carries everything)
    // stats (profiling runs): [0] (half, coordinate) replays run, [1] of them replayed with div.rn.f32, [2] cycles warp 0
    // spent replaying, [3] cycles warp 1 spent gathering (summed over all clusters)
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    const int coord = lane < 3 ? lane : 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sm.cluster = atomicAdd(next_cluster, 1);
        __syncthreads();
        // the cluster list is walked twice: clusters of LARGE segments first (their serial replay is the kernel's critical path
        // and must not start last), the others afterwards
        const int k2 = sm.cluster;
        if (k2 >= 2 * K) break;
        const int kk = k2 >= K ? k2 - K : k2;
        const int s = clt_seg[kk];
        const int local = kk - sg.id_base[s];
        const int b = sg.start[s], e = sg.start[s + 1];
        if ((e - b >= kCtrBigSegment) != (k2 < K)) continue;  // uniform: every thread reads the same sm.cluster
        const int nchunks = (e - b + kCtrChunkPts - 1) / kCtrChunkPts;
        float M = 0.f;   // lane c < 3 of warp 0 carries coordinate c
        int cnt = 0;     // members replayed so far (warp 0) / gathered so far (warps 1..7)
        if (wid > 0) ctr_gather(sm, 0, b, e, local, 0, cluster_id, x, y, z);
        int gathered = 0;
        for (int c = 0; c < nchunks; c++) {
            __syncthreads();                      // half c&1 is complete
            const int fill = sm.m[c & 1];
            const long long tc0 = stats ? clock64() : 0;
            if (wid == 0) {
                const int tier = ctr_chain(sm, c & 1, fill, cnt, M, coord);
                if (stats && lane < 3 && fill > 0) {
                    atomicAdd(stats, 1ull);
                    if (tier > 2) atomicAdd(stats + 1, 1ull);
                }
                if (stats && lane == 0) atomicAdd(stats + 2, (unsigned long long)(clock64() - tc0));
            } else {
                gathered += fill;
                if (c + 1 < nchunks) ctr_gather(sm, (c + 1) & 1, b + (c + 1) * kCtrChunkPts, e, local, gathered, cluster_id, x, y, z);
                if (stats && threadIdx.x == 32) atomicAdd(stats + 3, (unsigned long long)(clock64() - tc0));
            }
        }
        if (wid == 0 && lane < 3) {
            center[3 * kk + lane] = M;
            if (center_head && kk < n_head) center_head[3 * kk + lane] = M;
        }
    }
}

__global__ void __launch_bounds__(256)
k_centres(const int *__restrict__ d_K, SegArrays sg, const int *__restrict__ clt_seg,
          const int *__restrict__ cluster_id, const float *__restrict__ x, const float *__restrict__ y,
          const float *__restrict__ z, float *__restrict__ center, int *__restrict__ next_cluster,
          unsigned long long *__restrict__ stats) {
    __shared__ CtrSmem sm;
    centres_block(*d_K, sg, clt_seg, cluster_id, x, y, z, center, next_cluster, sm, nullptr, 0, stats);
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 (three kernels: block sums, spine, down-sweep).  n may live on the device.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
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
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        int i = i0 + k;
        v[k] = i < n ? in[i] : 0;
        sum += v[k];
    }
    int total;
    int ex = block_excl_scan(sum, smem, total) + block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        int i = i0 + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

// ------------------------------------------------------------------------------------------------
// single-pass exclusive scan (decoupled look-back): ONE launch instead of three.  Tiles take their index from a ticket
// counter (a tile only ever waits for tiles whose CTAs are already running); a tile publishes {epoch, flag, value} in
// one 64-bit word: flag 1 = tile aggregate, 2 = inclusive prefix.  The epoch (host counter, never 0) makes stale words
// of earlier scans invisible, so the state array is zeroed once per call, not per scan.  The last tile writes the total
// and rearms the ticket counter.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_onepass(const int *__restrict__ in, int n_host, const int *__restrict__ n_dev, int *__restrict__ out,
               int *__restrict__ total_out, unsigned long long *tile_state, int *ticket, unsigned epoch) {
    __shared__ int smem[33];
    __shared__ int s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const int n = n_dev ? *n_dev : n_host;
    const int i0 = tile * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
    // thread t owns kScanItems consecutive items = 16-byte vectors (the arena keeps every array 256-byte aligned)
    const bool full = i0 + kScanItems <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (full) {
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            int4 a = __ldg(reinterpret_cast<const int4 *>(in + i0) + q);
            v[4 * q] = a.x, v[4 * q + 1] = a.y, v[4 * q + 2] = a.z, v[4 * q + 3] = a.w;
        }
#pragma unroll
        for (int k = 0; k < kScanItems; k++) sum += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            int i = i0 + k;
            v[k] = i < n ? in[i] : 0;
            sum += v[k];
        }
    }
    int total;
    int ex = block_excl_scan(sum, smem, total);
    if (threadIdx.x < 32) {  // warp 0: publish the aggregate, look back 32 tiles at a time
        const int lane = threadIdx.x;
        const unsigned long long tag = (unsigned long long)epoch << 34;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(tile_state, tag | (2ull << 32) | (unsigned)total);
        } else {
            if (lane == 0) st_volatile_u64(tile_state + tile, tag | (1ull << 32) | (unsigned)total);
            int j = tile - 1;  // lane l inspects tile j - l; tiles before 0 count as a published prefix of 0
            while (true) {
                int idx = j - lane;
                unsigned long long w = idx >= 0 ? ld_volatile_u64(tile_state + idx) : (tag | (2ull << 32));
                bool ready = (w >> 34) == epoch;
                unsigned not_ready = ~__ballot_sync(kFull, ready);
                int usable = not_ready ? __ffs(not_ready) - 1 : 32;  // lanes [0, usable) hold published words
                unsigned is_prefix = __ballot_sync(kFull, ready && ((w >> 32) & 3ull) == 2ull) & (usable == 32 ? kFull : ((1u << usable) - 1u));
                int stop = is_prefix ? __ffs(is_prefix) - 1 : usable - 1;  // last lane whose value is added
                int val = (lane <= stop) ? (int)(unsigned)w : 0;
                prefix += __reduce_add_sync(kFull, val);
                if (is_prefix) break;
                j -= usable;  // all of [0, usable) were aggregates; retry from the first unpublished tile
            }
            if (lane == 0) st_volatile_u64(tile_state + tile, tag | (2ull << 32) | (unsigned)(prefix + total));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == (int)gridDim.x - 1) {
                *total_out = prefix + total;
                *ticket = 0;
            }
        }
    }
    __syncthreads();
    ex += s_prefix;
    if (full) {
        int o[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; k++) o[k] = ex, ex += v[k];
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++)
            reinterpret_cast<int4 *>(out + i0)[q] = make_int4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            int i = i0 + k;
            if (i < n) out[i] = ex;
            ex += v[k];
        }
    }
}

}  // namespace pb
