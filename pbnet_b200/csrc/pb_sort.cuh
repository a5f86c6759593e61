// pb_sort.cuh — hand-written segmented LSD radix sort for sm_100a (replaces thrust/CUB on the grouping path;
// reference: lib/PB_lib/src/pbnet/binary.cu:64-68 sorts every segment with thrust::sort_by_key).
//
// What makes it different from a library sort:
//   * SEGMENTED.  Segments (scene copy x class) are already contiguous in the input, so the segment id never enters the
//     key: every tile of the (host-built) tile table belongs to ONE segment and the decoupled look-back of a pass only
//     chains the tiles of the same segment.  At C1 that removes 12-13 key bits = two 8-bit passes.
//   * TRIMMED KEYS.  The key holds only the occupied bits of the segment-local cell coordinates (device-side extent
//     reduction, pb_fused.cuh), 26 bits at C1 instead of 42 -> 32-bit keys, three 9-bit passes instead of seven 8-bit ones.
//   * The per-(segment, pass) digit histograms are accumulated by the kernel that BUILDS the keys (k_keys), so no
//     separate histogram sweep reads the keys again; single-tile segments need neither histograms nor look-back.
//   * Both sorts of the path (cell order, Morton order of the LP assignment) run in the same launches (blockIdx.y).
// One pass = one launch (or one phase of the small-problem kernel): onesweep-style chained scan, match-based stable
// ranking inside the tile, reordering through shared memory so that the global writes are coalesced digit runs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int kTB = 256;            // threads per block of every tile kernel
constexpr int kRadixBitsMax = 9;
constexpr int kBins = 1 << kRadixBitsMax;  // 512
constexpr int kMaxPasses = 8;       // 64 key bits / 9 (the grouping path needs at most 5 + 4)
constexpr int kMaxPassesGroup = 9;  // passes of BOTH grouping sorts together: 42-bit cell keys (5) + 32-bit Morton/class keys (4)

// Segment-aligned tiles of the point range (host-built, one upload per call)
struct TileTab {
    const int *begin;   // [T] first point of the tile
    const int *count;   // [T] points in the tile, 1..TILE
    const int *seg;     // [T] segment of the tile
    const int *first;   // [T] index of the first tile of the same segment
    const int *hslot;   // [T] histogram slot of the segment (segments of >= 2 tiles) or -1
    const int *srow;    // [T] row of the tile in the look-back state arrays (tiles of multi-tile segments) or -1
    int T;
};

struct PassPlan {       // digit positions of one sort (host- or device-computed from the key width)
    int npass;
    int shift[kMaxPasses];
    int width[kMaxPasses];
};

__host__ __device__ inline PassPlan make_pass_plan(int bits) {
    PassPlan p;
    if (bits < 1) bits = 1;
    p.npass = (bits + kRadixBitsMax - 1) / kRadixBitsMax;
    int base = bits / p.npass, extra = bits % p.npass, s = 0;
    for (int i = 0; i < kMaxPasses; i++) {
        int w = i < p.npass ? base + (i < extra ? 1 : 0) : 0;
        p.shift[i] = s;
        p.width[i] = w;
        s += w;
    }
    return p;
}

template <typename KeyT>
struct SortArgs {
    const KeyT *keys_in;
    KeyT *keys_out;
    const uint32_t *vals_in;   // nullptr: payload = global point index (first pass)
    uint32_t *vals_out;
    const unsigned *hist;      // [slots][hist_stride] digit counts of the multi-tile segments; this sort's pass p at hist_off
    int hist_stride;           // unsigned words per slot
    int hist_off;              // offset of (this sort, this pass) inside a slot
    unsigned *state;           // [rows][kBins] look-back words of this pass (row = tt.srow), zero-initialised: flag<<30 | value
    int *ticket;               // tile ticket of this pass, zero-initialised
    int shift, width;
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan over 2*kTB values (thread t owns values 2t, 2t+1); smem: 33 ints
__device__ __forceinline__ void block_excl_scan2(int v0, int v1, int *smem, int &e0, int &e1) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v = v0 + v1, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (kTB >> 5) ? smem[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kTB >> 5)) smem[lane] = winc - w;
    }
    __syncthreads();
    e0 = inc - v + smem[wid];
    e1 = e0 + v0;
    __syncthreads();
}

template <typename KeyT, int ITEMS>
struct SortSmem {
    unsigned short whist[kTB / 32][kBins];  // per-warp digit counts, then exclusive warp offsets
    int loc[kBins];                         // tile-local start of every digit run
    int dst[kBins];                         // global destination of the run start minus loc
    int scan[40];
    int tile;
    KeyT skey[kTB * ITEMS];
    uint32_t sval[kTB * ITEMS];
};

// One tile of one pass.  `t` = tile index (from the pass's ticket counter: a tile only ever waits for tiles that were
// handed out earlier, i.e. whose blocks are running or finished).
template <typename KeyT, int ITEMS>
__device__ __forceinline__ void sort_tile(const SortArgs<KeyT> &a, const TileTab &tt, const int *__restrict__ seg_start, int t,
                                          SortSmem<KeyT, ITEMS> &s) {
    constexpr int WARPS = kTB / 32;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int begin = tt.begin[t], count = tt.count[t], seg = tt.seg[t], first = tt.first[t], hslot = tt.hslot[t];
    const unsigned mask = (1u << a.width) - 1u;
    // ---- load (coalesced: item k of lane l of warp w is tile element w*32*ITEMS + 32k + l) and zero the warp histograms
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(&s.whist[0][0]);
        for (int i = tid; i < WARPS * kBins / 2; i += kTB) z[i] = 0u;
    }
    KeyT key[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        key[k] = idx < count ? a.keys_in[begin + idx] : (KeyT)0;
    }
    __syncthreads();
    // ---- stable rank inside the warp's chunk: lanes with the same digit form a peer group; its lowest lane bumps the
    //      warp-private counter, the others add their position inside the group
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        bool valid = idx < count;
        unsigned d = valid ? ((unsigned)(key[k] >> a.shift) & mask) : 0xffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned r = 0;
        if (valid) {
            int leader = __ffs(peers) - 1;
            unsigned old = 0;
            if (lane == leader) {
                old = s.whist[w][d];
                s.whist[w][d] = (unsigned short)(old + __popc(peers));
            }
            old = __shfl_sync(peers, old, leader);
            r = old + __popc(peers & ((1u << lane) - 1u));
        }
        rank[k] = (unsigned short)r;
        __syncwarp();
    }
    __syncthreads();
    // ---- per digit (thread t owns digits 2t, 2t+1): counts over the warps -> exclusive warp offsets, tile counts
    const int d0 = 2 * tid, d1 = d0 + 1;
    int c0 = 0, c1 = 0;
#pragma unroll
    for (int ww = 0; ww < WARPS; ww++) {
        int x0 = s.whist[ww][d0], x1 = s.whist[ww][d1];
        s.whist[ww][d0] = (unsigned short)c0;
        s.whist[ww][d1] = (unsigned short)c1;
        c0 += x0;
        c1 += x1;
    }
    const bool multi = hslot >= 0;
    unsigned *st = a.state + (size_t)(multi ? tt.srow[t] : 0) * kBins;  // tiles of one segment own consecutive rows
    if (multi) {  // publish the tile's digit counts right away: aggregate (1) or, for the segment's first tile, prefix (2)
        unsigned f = (t == first ? 2u : 1u) << 30;
        st_volatile_u32(st + d0, f | (unsigned)c0);
        st_volatile_u32(st + d1, f | (unsigned)c1);
    }
    int l0, l1;
    block_excl_scan2(c0, c1, s.scan, l0, l1);
    s.loc[d0] = l0;
    s.loc[d1] = l1;
    int g0 = l0, g1 = l1, p0 = 0, p1 = 0;
    if (multi) {
        const unsigned *h = a.hist + (size_t)hslot * a.hist_stride + a.hist_off;
        block_excl_scan2((int)h[d0], (int)h[d1], s.scan, g0, g1);  // start of every digit inside the segment
        if (t != first) {
            // decoupled look-back over the earlier tiles of this segment, one chain per digit
            bool done0 = false, done1 = false;
            for (int j = t - 1; !(done0 && done1); j--) {
                const unsigned *sj = st - (size_t)(t - j) * kBins;
                if (!done0) {
                    unsigned v;
                    do { v = ld_volatile_u32(sj + d0); } while ((v >> 30) == 0u);
                    p0 += (int)(v & 0x3fffffffu);
                    done0 = (v >> 30) == 2u;
                }
                if (!done1) {
                    unsigned v;
                    do { v = ld_volatile_u32(sj + d1); } while ((v >> 30) == 0u);
                    p1 += (int)(v & 0x3fffffffu);
                    done1 = (v >> 30) == 2u;
                }
            }
            st_volatile_u32(st + d0, (2u << 30) | (unsigned)(p0 + c0));
            st_volatile_u32(st + d1, (2u << 30) | (unsigned)(p1 + c1));
        }
    }
    const int sbase = seg_start[seg];
    s.dst[d0] = sbase + g0 + p0 - l0;
    s.dst[d1] = sbase + g1 + p1 - l1;
    __syncthreads();
    // ---- reorder through shared memory, then write digit runs with consecutive threads on consecutive addresses
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = w * 32 * ITEMS + 32 * k + lane;
        if (idx < count) {
            unsigned d = (unsigned)(key[k] >> a.shift) & mask;
            int lp = s.loc[d] + s.whist[w][d] + rank[k];
            s.skey[lp] = key[k];
            s.sval[lp] = a.vals_in ? a.vals_in[begin + idx] : (uint32_t)(begin + idx);  // payloads are only touched here
        }
    }
    __syncthreads();
    for (int p = tid; p < count; p += kTB) {
        KeyT kk = s.skey[p];
        unsigned d = (unsigned)(kk >> a.shift) & mask;
        int o = s.dst[d] + p;
        a.keys_out[o] = kk;
        a.vals_out[o] = s.sval[p];
    }
    __syncthreads();
}

// One pass of up to two independent sorts over the same tile table (blockIdx.y selects the sort).
#ifndef PB_SORT_MINB
#define PB_SORT_MINB 4   // 64 registers, 4 x 45 KB of shared memory per SM: 2.45 ms per C1 step vs 2.76 ms at 3 (80 registers)
#endif
template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kTB, PB_SORT_MINB)
k_sort_pass(SortArgs<KeyT> a0, SortArgs<KeyT> a1, TileTab tt, const int *__restrict__ seg_start) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem<KeyT, ITEMS> &s = *reinterpret_cast<SortSmem<KeyT, ITEMS> *>(smem_raw);
    const SortArgs<KeyT> &a = blockIdx.y == 0 ? a0 : a1;
    if (threadIdx.x == 0) s.tile = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int t = s.tile;
    if (t >= tt.T) return;
    sort_tile<KeyT, ITEMS>(a, tt, seg_start, t, s);
}

// ------------------------------------------------------------------------------------------------------------------
// Generic use (voxelize, local scenes, evaluation, normals): ONE segment covering all n keys.  The digit histograms of all
// passes come from one sweep over the keys (the grouping path gets them for free from k_keys).
// ------------------------------------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kTB)
k_radix_hist(const KeyT *__restrict__ keys, int n, PassPlan plan, unsigned *__restrict__ hist) {
    __shared__ unsigned sh[kMaxPasses * kBins];
    const int nh = plan.npass * kBins;
    for (int i = threadIdx.x; i < nh; i += kTB) sh[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * kTB + threadIdx.x; i < n; i += (long long)gridDim.x * kTB) {
        KeyT k = keys[i];
#pragma unroll
        for (int p = 0; p < kMaxPasses; p++)
            if (p < plan.npass) atomicAdd(sh + p * kBins + ((unsigned)(k >> plan.shift[p]) & ((1u << plan.width[p]) - 1u)), 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += kTB)
        if (sh[i]) atomicAdd(hist + i, sh[i]);
}

}  // namespace pb
