// pb_small.cuh — the grouping path for SMALL calls (the reference's own call pattern: one class of one scene,
// hundreds to a few thousand points, lib/PB_lib/torch_io/pbnet_ops.py:14-75) as ONE cooperative launch.
//
// Round 1 ran every call through the cell-grid pipeline: 46 dependent launches, two radix sorts, a dozen scans — 0.7 ms
// of fixed latency for a 6 k-point class.  For a segment this small the O(n^2) formulation is cheaper than building any
// acceleration structure, and it maps onto warp-wide bit operations:
//   P1  adjacency bitmap + degree   one warp tests 8 query points against the whole segment, 32 candidates per
//                                   instruction; a ballot is one 32-bit word of the query's adjacency row
//   P2  parent[u] = smallest-index HP neighbour (first set bit of row & HP mask): a forest that already joins almost
//       every point of a blob
//   P3  components of the HP graph.  The P2 forest has only a handful of trees (its roots are the local minima of the
//       point index: about n / (HP degree + 1) of them), so with <= 64 trees the components are computed on the TREE graph:
//       every HP point ORs the tree numbers of its HP neighbours into one 64-bit word (byte loads, no atomics on the
//       neighbours, no pointer chasing), one 64-bit atomicOr per point builds the 64 x 64 tree adjacency, and every block
//       closes it with Warshall's algorithm in one warp.  More than 64 trees (sparse graphs): union-find at WORD
//       granularity — 32 neighbours' parents are compared with the query's root in one coalesced load; only genuinely
//       foreign neighbours pay a find / union
//   P4  flatten; roots flagged at their (minimum) point index; block-level scan -> raw ids in seed order
//       (binary.cu:161-166)
//   P5  labels: HP -> id of its component; border LP -> maximum id over its HP neighbours (binary.cu:206-213); sizes
//   P6  fragment filter + compaction (the same block routine the large path uses)
//   P7  final ids, labelled-point mask
//   P8  exact 1-NN of the unlabelled points in original coordinates, ties -> largest index
//       (binary_cuda_functions.cu:273-286)
//   P9  centres: exact replay of the running mean (binary_cuda_functions.cu:217-246)
// separated by grid-wide barriers (cooperative groups).  Same predicate, same canonical numbering, same tie rules as the
// large path — the tests run both on the same inputs.
#pragma once
#include <cooperative_groups.h>

#include "pb_fused.cuh"

namespace pbsm {

namespace cg = cooperative_groups;
using pb::kFull;

constexpr int kMaxSeg = 32;       // segments per small call
constexpr int kQB = 8;            // query points per P1 task (the select tree of P1 is written for 8)
constexpr int kThreads = 256;
constexpr int kTreeMax = 64;     // trees the tree-graph formulation of P3 handles
constexpr int kCentreHead = 256;  // clusters whose centres travel with the result block
constexpr int kWPL = 16;           // bitmask words per lane: a segment has at most 32*kWPL words = 16384 points

struct SmallArgs {
    int n, S, assign_lp;
    int start[kMaxSeg + 1];       // first point of every segment
    int mask_off[kMaxSeg + 1];    // first bitmask word of every segment (one bit per point, segments word-aligned)
    long long adj_off[kMaxSeg + 1];  // first adjacency word of every segment (row-major, mask words per row)
    float radius[18];
    int min_pts[18];
    float thresh[18];
    const float *x, *y, *z, *xo, *yo, *zo;
    const int *sem;
    int *cluster_id, *cluster_num, *degree, *clt_sem;
    float *center;
    pb::SegArrays sg;             // start, cls, min_pts, r2, k_base, cluster_num, id_base are used
    const int *call_first;        // [S] zeros (one call)
    unsigned *adj, *hpmask, *labmask;
    int *root0;                   // root of every point in the P2 forest
    unsigned long long *best64;   // P8: per point, min over candidates of (distance bits << 32 | ~index)
    int *seg_of, *parent, *flag, *gid_at, *raw_label, *raw_count, *rep, *keep, *kscan, *clt_seg;
    int deg_slices;               // P1: slices per query group (0: automatic)
    int *deg_acc;                 // P1: partial neighbour counts of the sliced form (zero-initialised)
    int tree_cap;                 // P3 works on the tree graph when the P2 forest has at most this many trees (<= 64; 0: never)
    int *tslot, *troot, *tcnt;    // tree number of every forest root, root point of every tree, number of trees
    unsigned char *tidx;          // tree number of every HP point
    unsigned long long *conn;     // [64] tree adjacency rows (zero-initialised)
    int *lplist, *lpcnt;          // P7/P8: unlabelled points of every segment (unordered), their number (zero-initialised)
    float *center_head;           // centres / classes of the first kSmallCentreHead clusters, next to the result block (or null)
    int *cltsem_head;
    int *scal;                    // [0] err [2] R [3] K [9] centre ticket [16..39] phase time stamps (globaltimer ns, 64-bit)
};

__device__ __forceinline__ int seg_of_point(const SmallArgs &a, int i) {
    int s = 0;
    while (s + 1 < a.S && a.start[s + 1] <= i) s++;
    return s;
}

// exclusive scan of flag[0..n) by ONE block (kThreads threads, 16 items per thread and trip)
__device__ __forceinline__ void block_scan_flags(const int *__restrict__ flag, int n, int *__restrict__ out, int *total_out, int *smem) {
    int carry = 0;
    for (int base = 0; base < n; base += kThreads * 16) {
        int i0 = base + threadIdx.x * 16;
        int v[16], sum = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            v[k] = (i0 + k < n) ? flag[i0 + k] : 0;
            sum += v[k];
        }
        int total;
        int ex = pb::block_excl_scan_any(sum, smem, total) + carry;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
    __syncthreads();
}

__device__ __forceinline__ void stamp(const SmallArgs &a, int k) {  // phase boundaries as seen by thread 0 of the grid
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long *>(a.scal + 16)[k] = t;
    }
}

__global__ void __launch_bounds__(kThreads, 2)
k_small(SmallArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_scan[40];
    __shared__ pb::CtrSmem s_ctr;
    const int lane = threadIdx.x & 31;
    // warp-task index: consecutive tasks go to DIFFERENT blocks, so a short task list still spreads over all SMs
    const int gwarp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int gthread = blockIdx.x * blockDim.x + threadIdx.x, nthread = gridDim.x * blockDim.x;
    const int n = a.n, S = a.S;
    int *d_err = a.scal, *d_R = a.scal + 2, *d_K = a.scal + 3;

    stamp(a, 0);
    // device copies of the (by-value) segment table for the routines shared with the large path
    if (gthread <= S) const_cast<int *>(a.sg.start)[gthread] = a.start[gthread];
    if (gthread < S) const_cast<int *>(a.call_first)[gthread] = 0;
    // ---------------- P1: adjacency rows + degrees + HP mask + validation -------------------------------------------
    {
        // task = 8 query points x a slice of the segment's candidate words.  With fewer query groups than warps in the grid
        // the candidate range of a group is cut into slices (at least four words each) so that the task list fills the grid
        // about once: the partial counts then meet in deg_acc and a short extra phase applies the HP rule.
        int task0[kMaxSeg + 1];  // first task of every segment
        int groups = 0;
        for (int s = 0; s < S; s++) groups += (a.start[s + 1] - a.start[s] + kQB - 1) / kQB;
        const int want = a.deg_slices > 0 ? a.deg_slices : max(1, nwarp / max(groups, 1));
        bool sliced = false;
        int nt = 0;
        for (int s = 0; s < S; s++) {
            const int W = a.mask_off[s + 1] - a.mask_off[s], nsl = max(1, min(want, W >> 2));
            sliced |= nsl > 1;
            task0[s] = nt, nt += ((a.start[s + 1] - a.start[s] + kQB - 1) / kQB) * nsl;
        }
        task0[S] = nt;
        for (int t = gwarp; t < nt; t += nwarp) {
            int s = 0;
            while (task0[s + 1] <= t) s++;
            const int b = a.start[s], e = a.start[s + 1], W = a.mask_off[s + 1] - a.mask_off[s];
            const int nsl = max(1, min(want, W >> 2));
            const int qg = (t - task0[s]) / nsl, sl = (t - task0[s]) % nsl;
            const int w0 = (int)(((long long)W * sl) / nsl), w1 = (int)(((long long)W * (sl + 1)) / nsl);
            const int q0 = b + qg * kQB;
            int cls0 = __ldg(a.sem + b);
            const bool cls_ok = cls0 >= 2 && cls0 <= 19;
            if (!cls_ok) cls0 = 2;
            const float r = a.radius[cls0 - 2];
            const float r2 = __fmul_rn(r, r);                       // binary_cuda_functions.cu:85
            const int minp = a.min_pts[cls0 - 2];
            if (q0 == b && sl == 0 && lane == 0) a.sg.cls[s] = cls0, a.sg.min_pts[s] = minp, a.sg.r2[s] = r2;
            // validation of my query points (every point is a query of exactly one slice-0 task)
            float qx[kQB], qy[kQB], qz[kQB];
            int mycnt = 0;   // lane j < 8: neighbours of query j found in this slice
            const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
            int err = cls_ok ? 0 : pb::kErrSem;
#pragma unroll
            for (int j = 0; j < kQB; j++) {
                int u = min(q0 + j, e - 1);
                qx[j] = __ldg(a.x + u), qy[j] = __ldg(a.y + u), qz[j] = __ldg(a.z + u);
            }
            if (sl == 0 && lane < kQB && q0 + lane < e) {
                int u = q0 + lane, c = a.sem[u];
                a.seg_of[u] = s;
                if (c < 2 || c > 19) err |= pb::kErrSem;
                else if (c != cls0) err |= pb::kErrMixed;
                if (!(isfinite(a.x[u]) && isfinite(a.y[u]) && isfinite(a.z[u]) && isfinite(a.xo[u]) && isfinite(a.yo[u]) && isfinite(a.zo[u])))
                    err |= pb::kErrNonFinite;
            }
            if (err) atomicOr(d_err, err);
            unsigned *row = a.adj + a.adj_off[s] + (long long)(q0 - b) * W;
            // candidates: 32 per word; the coordinates of the next FOUR words are in flight while four are tested (the loop
            // is bound by L2 latency otherwise: one word of tests is ~100 instructions, a load ~600 cycles)
            constexpr int PF = 4;
            float nx[PF], ny[PF], nz[PF];
#pragma unroll
            for (int q = 0; q < PF; q++) {
                int v = b + (w0 + q) * 32 + lane;
                nx[q] = ny[q] = nz[q] = 0.f;
                if (v < e) nx[q] = __ldg(a.x + v), ny[q] = __ldg(a.y + v), nz[q] = __ldg(a.z + v);
            }
            for (int vb0 = w0; vb0 < w1; vb0 += PF) {
                float cx[PF], cy[PF], cz[PF];
#pragma unroll
                for (int q = 0; q < PF; q++) {
                    cx[q] = nx[q], cy[q] = ny[q], cz[q] = nz[q];
                    int v = b + (vb0 + PF + q) * 32 + lane;
                    if (v < e) nx[q] = __ldg(a.x + v), ny[q] = __ldg(a.y + v), nz[q] = __ldg(a.z + v);
                }
#pragma unroll
                for (int q = 0; q < PF; q++) {
                    const int vb = vb0 + q;
                    if (vb >= w1) break;
                    const bool valid = b + vb * 32 + lane < e;
                    unsigned wv[kQB];
#pragma unroll
                    for (int j = 0; j < kQB; j++)
                        wv[j] = __ballot_sync(kFull, valid && pb::sqd(qx[j], qy[j], qz[j], cx[q], cy[q], cz[q]) <= r2);
                    // lane j < 8 keeps the word of query j: a 3-level select on the lane's bits (predicates set once per
                    // task), ONE population count per word for all eight queries (POPC is a quarter-rate instruction)
                    const unsigned s0 = b0 ? wv[1] : wv[0], s1 = b0 ? wv[3] : wv[2], s2 = b0 ? wv[5] : wv[4], s3 = b0 ? wv[7] : wv[6];
                    const unsigned mine = b2 ? (b1 ? s3 : s2) : (b1 ? s1 : s0);
                    mycnt += __popc(mine);
                    if (lane < kQB && q0 + lane < e) row[(long long)lane * W + vb] = mine;
                }
            }
            if (lane < kQB && q0 + lane < e) {
                int u = q0 + lane;
                if (sliced) {
                    if (mycnt) atomicAdd(a.deg_acc + u, mycnt);
                } else {
                    // degree (self excluded by position, binary_cuda_functions.cu:88), HP rule (:175-186)
                    int d = mycnt - 1;
                    a.degree[u] = d;
                    if (d >= minp) atomicOr(a.hpmask + a.mask_off[s] + ((u - b) >> 5), 1u << ((u - b) & 31));
                }
            }
        }
        if (sliced) {   // uniform over the grid
            grid.sync();
            const int words = a.mask_off[S];
            for (int wd = gwarp; wd < words; wd += nwarp) {
                int s = 0;
                while (a.mask_off[s + 1] <= wd) s++;
                const int u = a.start[s] + (wd - a.mask_off[s]) * 32 + lane;
                bool hp = false;
                if (u < a.start[s + 1]) {
                    const int d = __ldcg(a.deg_acc + u) - 1;
                    a.degree[u] = d;
                    hp = d >= a.sg.min_pts[s];
                }
                const unsigned m = __ballot_sync(kFull, hp);
                if (lane == 0) a.hpmask[wd] = m;
            }
        }
    }
    grid.sync();
    stamp(a, 1);
    // ---------------- P2: parent = smallest-index HP neighbour (self for the minimum of its neighbourhood) ----------
    for (int u = gthread; u < n; u += nthread) {
        int s = a.seg_of[u];
        const int b = a.start[s], W = a.mask_off[s + 1] - a.mask_off[s], lu = u - b;
        const unsigned *hm = a.hpmask + a.mask_off[s];
        int p = u;
        if ((hm[lu >> 5] >> (lu & 31)) & 1u) {
            const unsigned *row = a.adj + a.adj_off[s] + (long long)lu * W;
            for (int vb = 0; vb <= (lu >> 5); vb++) {
                unsigned w = row[vb] & hm[vb];
                if (w) {
                    p = b + vb * 32 + __ffs(w) - 1;
                    break;
                }
            }
        }
        a.parent[u] = p;   // p <= u
    }
    grid.sync();
    stamp(a, 2);
    // ---------------- P3: components of the HP graph.
    //   a) flatten the P2 forest (parent[u] = root of its tree): a neighbour whose PARENT equals the query's root is in the
    //      query's component for sure, which is the cheap test of b)
    //   b) union-find at WORD granularity over the HP neighbours with a smaller index (every edge once): the row's words
    //      are fetched lane-parallel up front, the parent vectors of 8 words are in flight together; lanes whose parent
    //      differs look their roots up in parallel (compressing paths), only a genuinely different root costs a union
    for (int u = gthread; u < n; u += nthread) {   // a)
        int s = a.seg_of[u];
        const int lu = u - a.start[s];
        int r = u;
        if ((a.hpmask[a.mask_off[s] + (lu >> 5)] >> (lu & 31)) & 1u) {
            while (true) {
                int p = __ldcg(a.parent + r);
                if (p == r) break;
                r = p;
            }
            if (r == u) {   // a tree root: number it (in arrival order; the canonical root of a component is a minimum over points)
                int slot = atomicAdd(a.tcnt, 1);
                if (slot < kTreeMax) a.troot[slot] = u;
                a.tslot[u] = slot;
            }
        }
        a.root0[u] = r;
    }
    grid.sync();
    const int T = __ldcg(a.tcnt);
    const bool tree_graph = T <= a.tree_cap;   // uniform over the grid
    if (tree_graph) {
        // b') tree number of every HP point
        for (int u = gthread; u < n; u += nthread) {
            int s = a.seg_of[u];
            const int lu = u - a.start[s];
            if ((a.hpmask[a.mask_off[s] + (lu >> 5)] >> (lu & 31)) & 1u) a.tidx[u] = (unsigned char)a.tslot[a.root0[u]];
        }
        grid.sync();
        // c') one warp per HP point: lane l walks the words l, l+32, .. of the point's row (smaller indices only: every edge
        //     once) and ORs the tree numbers of the HP neighbours it finds into a 64-bit set
        for (int u = gwarp; u < n; u += nwarp) {
            int s = a.seg_of[u];
            const int b = a.start[s], W = a.mask_off[s + 1] - a.mask_off[s], lu = u - b;
            const unsigned *hm = a.hpmask + a.mask_off[s];
            if (!((hm[lu >> 5] >> (lu & 31)) & 1u)) continue;
            const unsigned *row = a.adj + a.adj_off[s] + (long long)lu * W;
            const int nwords = (lu >> 5) + 1;
            const unsigned char *tix = a.tidx + b;
            unsigned long long acc = 0ull;
            for (int vb = lane; vb < nwords; vb += 32) {
                unsigned w = row[vb] & hm[vb];
                if (vb == (lu >> 5)) w &= (1u << (lu & 31)) - 1u;
                const unsigned char *tw = tix + vb * 32;
                while (w) {   // two neighbours per trip: the byte loads are independent
                    int k0 = __ffs(w) - 1;
                    w &= w - 1;
                    int k1 = w ? __ffs(w) - 1 : k0;
                    w &= w - 1;
                    acc |= (1ull << tw[k0]) | (1ull << tw[k1]);
                }
            }
            unsigned lo = __reduce_or_sync(kFull, (unsigned)acc), hi = __reduce_or_sync(kFull, (unsigned)(acc >> 32));
            const int mine = a.tidx[u];
            unsigned long long set = (((unsigned long long)hi << 32) | lo) & ~(1ull << mine);
            if (lane == 0 && set) atomicOr(a.conn + mine, set);
        }
    } else {
    for (int u = gthread; u < n; u += nthread) a.parent[u] = a.root0[u];
    grid.sync();
    for (int u = gwarp; u < n; u += nwarp) {       // b)
        int s = a.seg_of[u];
        const int b = a.start[s], W = a.mask_off[s + 1] - a.mask_off[s], lu = u - b;
        const unsigned *hm = a.hpmask + a.mask_off[s];
        if (!((hm[lu >> 5] >> (lu & 31)) & 1u)) continue;
        const unsigned *row = a.adj + a.adj_off[s] + (long long)lu * W;
        const int nwords = (lu >> 5) + 1;   // strictly smaller indices: every edge once
        unsigned wv[kWPL];
#pragma unroll
        for (int i = 0; i < kWPL; i++) {
            int vb = i * 32 + lane;
            wv[i] = vb < nwords ? (row[vb] & hm[vb]) : 0u;
            if (vb == (lu >> 5)) wv[i] &= (1u << (lu & 31)) - 1u;
        }
        int r = pb::uf_find(a.parent, u);
#pragma unroll
        for (int i = 0; i < kWPL; i++) {
            if (i * 32 >= nwords) break;
            for (int k0 = 0; k0 < 32 && i * 32 + k0 < nwords; k0 += 8) {
                unsigned w8[8];
                int pv[8];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    w8[t] = __shfl_sync(kFull, wv[i], k0 + t);
                    pv[t] = ((w8[t] >> lane) & 1u) ? __ldcg(a.parent + b + (i * 32 + k0 + t) * 32 + lane) : r;
                }
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    unsigned f = __ballot_sync(kFull, pv[t] != r);
                    if (f) {
                        const int v = b + (i * 32 + k0 + t) * 32 + lane;
                        bool changed = false;
                        while (f) {
                            int rv = ((f >> lane) & 1u) ? pb::uf_find(a.parent, v) : r;
                            unsigned g = __ballot_sync(kFull, rv != r);
                            if (!g) break;
                            int j = __ffs(g) - 1;
                            if (lane == 0) pb::uf_union(a.parent, u, b + (i * 32 + k0 + t) * 32 + j);
                            __syncwarp();
                            r = pb::uf_find(a.parent, u);
                            changed = true;
                            f = g & ~(1u << j);
                        }
                        if (changed) {   // the batch's remaining parent vectors were fetched for the old root: refresh them
#pragma unroll
                            for (int q = 0; q < 8; q++)
                                if (q > t) pv[q] = ((w8[q] >> lane) & 1u) ? __ldcg(a.parent + b + (i * 32 + k0 + q) * 32 + lane) : r;
                        }
                    }
                }
            }
        }
    }
    }
    grid.sync();
    stamp(a, 3);
    // ---------------- P4a: flatten, flag the roots (root = minimum index of its component) ---------------------------
    if (tree_graph) {
        // every block closes the tree graph on its own (64 rows of 64 bits: lanes hold rows l and l + 32): symmetrise,
        // Warshall in place, then the component's canonical root = the smallest root point among its trees
        __shared__ int s_comp_root[kTreeMax];
        if (threadIdx.x < 32) {
            unsigned long long r0 = lane < T ? __ldcg(a.conn + lane) : 0ull, r1 = lane + 32 < T ? __ldcg(a.conn + lane + 32) : 0ull;
            const int pt0 = lane < T ? __ldcg(a.troot + lane) : 0x7fffffff, pt1 = lane + 32 < T ? __ldcg(a.troot + lane + 32) : 0x7fffffff;
            unsigned long long c0 = 1ull << lane, c1 = 1ull << (lane + 32);
            for (int k = 0; k < T; k++) {   // column k of the transpose: bit (my row) of row k
                unsigned long long rk = __shfl_sync(kFull, k < 32 ? r0 : r1, k & 31);
                if ((rk >> lane) & 1ull) c0 |= 1ull << k;
                if ((rk >> (lane + 32)) & 1ull) c1 |= 1ull << k;
            }
            r0 |= c0, r1 |= c1;
            for (int k = 0; k < T; k++) {
                unsigned long long rk = __shfl_sync(kFull, k < 32 ? r0 : r1, k & 31);
                if ((r0 >> k) & 1ull) r0 |= rk;
                if ((r1 >> k) & 1ull) r1 |= rk;
            }
            int m0 = 0x7fffffff, m1 = 0x7fffffff;
            for (int k = 0; k < T; k++) {
                int pk = __shfl_sync(kFull, k < 32 ? pt0 : pt1, k & 31);
                if ((r0 >> k) & 1ull) m0 = min(m0, pk);
                if ((r1 >> k) & 1ull) m1 = min(m1, pk);
            }
            s_comp_root[lane] = m0, s_comp_root[lane + 32] = m1;
        }
        __syncthreads();
        for (int u = gthread; u < n; u += nthread) {
            int s = a.seg_of[u];
            const int lu = u - a.start[s];
            if ((a.hpmask[a.mask_off[s] + (lu >> 5)] >> (lu & 31)) & 1u) {
                int rt = s_comp_root[a.tidx[u]];
                __stcg(a.parent + u, rt);
                if (rt == u) a.flag[u] = 1;
            }
        }
    } else {
        for (int u = gthread; u < n; u += nthread) {
            int s = a.seg_of[u];
            const int lu = u - a.start[s];
            if ((a.hpmask[a.mask_off[s] + (lu >> 5)] >> (lu & 31)) & 1u) {
                int rt = pb::uf_find(a.parent, u);
                __stcg(a.parent + u, rt);
                if (rt == u) a.flag[u] = 1;
            }
        }
    }
    grid.sync();
    stamp(a, 4);
    // ---------------- P4b: raw ids = rank of the component's minimum HP index (binary.cu:161-166) --------------------
    if (blockIdx.x == 0) {
        block_scan_flags(a.flag, n, a.gid_at, d_R, s_scan);
        for (int u = threadIdx.x; u < n; u += blockDim.x)
            if (a.flag[u]) a.rep[a.gid_at[u]] = u;
    }
    grid.sync();
    stamp(a, 5);
    // ---------------- P5: labels + cluster sizes -----------------------------------------------------------------------
    for (int p = gwarp; p < n; p += nwarp) {
        int s = a.seg_of[p];
        const int b = a.start[s], W = a.mask_off[s + 1] - a.mask_off[s], lp = p - b;
        const unsigned *hm = a.hpmask + a.mask_off[s];
        int label = -1;
        if ((hm[lp >> 5] >> (lp & 31)) & 1u) {
            label = a.gid_at[__ldcg(a.parent + p)];
        } else {  // border LP: the LAST component that reaches it wins = maximum raw id (binary.cu:206-213)
            const unsigned *row = a.adj + a.adj_off[s] + (long long)lp * W;
            unsigned wv[kWPL];
#pragma unroll
            for (int i = 0; i < kWPL; i++) {
                int vb = i * 32 + lane;
                wv[i] = vb < W ? (row[vb] & hm[vb]) : 0u;
            }
            int best = -1;
#pragma unroll
            for (int i = 0; i < kWPL; i++) {
                if (i * 32 >= W) break;
                unsigned nz = __ballot_sync(kFull, wv[i] != 0u);   // words of this chunk that hold an HP neighbour
                while (nz) {
                    int k = __ffs(nz) - 1;
                    nz &= nz - 1;
                    unsigned w = __shfl_sync(kFull, wv[i], k);
                    if ((w >> lane) & 1u) best = max(best, a.gid_at[__ldcg(a.parent + b + (i * 32 + k) * 32 + lane)]);
                }
            }
            label = __reduce_max_sync(kFull, best);
        }
        if (lane == 0) {
            a.raw_label[p] = label;
            if (label >= 0) atomicAdd(a.raw_count + label, 1);
        }
    }
    grid.sync();
    stamp(a, 6);
    // ---------------- P6: fragment filter, compaction, per-segment counts ----------------------------------------------
    if (blockIdx.x == 0)
        pb::filter_scan_block<false>(n, S, a.sg, d_R, a.rep, a.seg_of, a.raw_count, a.thresh, a.keep, a.kscan, d_K, a.sem,
                                     a.call_first, a.gid_at, a.cluster_num, s_scan);
    grid.sync();
    stamp(a, 7);
    // ---------------- P7: final ids of the HP stage, labelled mask ----------------------------------------------------
    {
        const int words = a.mask_off[S];
        for (int wd = gwarp; wd < words; wd += nwarp) {
            int s = 0;
            while (a.mask_off[s + 1] <= wd) s++;
            int u = a.start[s] + (wd - a.mask_off[s]) * 32 + lane;
            bool lab = false;
            if (u < a.start[s + 1]) {
                int gi = a.raw_label[u], id = -1;
                if (gi >= 0 && a.keep[gi]) {
                    int kk = a.kscan[gi];
                    id = kk - a.sg.id_base[s];
                    if (a.rep[gi] == u) {
                        a.clt_sem[kk] = a.sg.cls[s], a.clt_seg[kk] = s;
                        if (a.cltsem_head && kk < kCentreHead) a.cltsem_head[kk] = a.sg.cls[s];
                    }
                    lab = true;
                }
                a.cluster_id[u] = id;
                a.best64[u] = ~0ull;
            }
            unsigned m = __ballot_sync(kFull, lab);
            if (lane == 0) a.labmask[wd] = m;
            // the unlabelled points of a segment that has clusters are the 1-NN queries of P8: appended to the segment's
            // list (the order of the list does not matter: every query is answered on its own)
            unsigned q = __ballot_sync(kFull, u < a.start[s + 1] && !lab);
            if (q && a.assign_lp && a.sg.cluster_num[s] > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(a.lpcnt + s, __popc(q));
                base = __shfl_sync(kFull, base, 0);
                if ((q >> lane) & 1u) a.lplist[a.start[s] + base + __popc(q & ((1u << lane) - 1u))] = u;
            }
        }
    }
    grid.sync();
    stamp(a, 8);
    // ---------------- P8: exact 1-NN of the unlabelled points (original coordinates) ----------------------------------
    // task = (32 queries of the segment's list, one per lane) x (a slice of the segment's candidates, staged through shared
    // memory 32 at a time); a lane walks its candidates in ascending order with '<=' (ties -> largest index, :282); slices
    // merge through atomicMin on (distance bits << 32 | ~index): smallest distance first, then the largest index.  The
    // slices are cut so that the task list fills the grid about once.
    if (a.assign_lp) {
        float4 *stage = reinterpret_cast<float4 *>(&s_ctr.buf[0][0][0]) + (threadIdx.x >> 5) * 32;   // 32 candidates per warp
        int t0[kMaxSeg + 1];
        int nt = 0, groups = 0;
        for (int s = 0; s < S; s++) groups += (__ldcg(a.lpcnt + s) + 31) >> 5;
        const int want = groups > 0 ? max(1, nwarp / groups) : 1;   // slices per query group
        for (int s = 0; s < S; s++) {
            const int W = a.mask_off[s + 1] - a.mask_off[s];
            t0[s] = nt;
            nt += ((__ldcg(a.lpcnt + s) + 31) >> 5) * min(want, (W + 1) >> 1);
        }
        t0[S] = nt;
        for (int t = gwarp; t < nt; t += nwarp) {
            int s = 0;
            while (t0[s + 1] <= t) s++;
            const int b = a.start[s], e = a.start[s + 1], W = a.mask_off[s + 1] - a.mask_off[s];
            const int nsl = min(want, (W + 1) >> 1), nq = __ldcg(a.lpcnt + s);
            const int qg = (t - t0[s]) / nsl, sl = (t - t0[s]) % nsl;
            const unsigned *lm = a.labmask + a.mask_off[s];
            const bool active = qg * 32 + lane < nq;
            const int u = active ? __ldcg(a.lplist + b + qg * 32 + lane) : b;
            const float px = a.xo[u], py = a.yo[u], pz = a.zo[u];
            float bestD = __int_as_float(0x7f800000);
            int bestI = -1;
            const int cw1 = (int)(((long long)W * (sl + 1)) / nsl);
            int cw = (int)(((long long)W * sl) / nsl);
            int vn = min(b + cw * 32 + lane, e - 1);
            float nx = a.xo[vn], ny = a.yo[vn], nz = a.zo[vn];
            for (; cw < cw1; cw++) {
                const unsigned lw = lm[cw];
                __syncwarp();
                stage[lane] = make_float4(nx, ny, nz, 0.f);
                __syncwarp();
                if (cw + 1 < cw1) {
                    vn = min(b + (cw + 1) * 32 + lane, e - 1);
                    nx = a.xo[vn], ny = a.yo[vn], nz = a.zo[vn];
                }
                if (!lw) continue;
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const float4 c = stage[k];
                    float D = pb::sqd(px, py, pz, c.x, c.y, c.z);
                    if (((lw >> k) & 1u) && D <= bestD) bestD = D, bestI = b + cw * 32 + k;
                }
            }
            if (active && bestI >= 0)
                atomicMin(a.best64 + u, ((unsigned long long)__float_as_uint(bestD) << 32) | (unsigned long long)(0xffffffffu - (unsigned)bestI));
        }
    }
    grid.sync();
    for (int u = gthread; u < n; u += nthread) {
        unsigned long long bv = a.best64[u];
        if (a.assign_lp && a.cluster_id[u] < 0 && bv != ~0ull) a.cluster_id[u] = a.cluster_id[0xffffffffu - (unsigned)(bv & 0xffffffffull)];
    }
    grid.sync();
    stamp(a, 9);
    // ---------------- P9: centres --------------------------------------------------------------------------------------
    pb::centres_block(*d_K, a.sg, a.clt_seg, a.cluster_id, a.x, a.y, a.z, a.center, a.scal + 9, s_ctr, a.center_head, kCentreHead);
    stamp(a, 10);   // end of block 0's own centres (the last phase is not followed by a barrier)
}

}  // namespace pbsm
