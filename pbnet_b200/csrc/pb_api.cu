// pb_api.cu — context, workspace and launch sequence behind the C ABI of include/pbnet_b200.h.
//
// Replaces the host side of the reference: binary_cluster's per-segment loop
// (lib/PB_lib/src/pbnet/cluster.cu:57-110) and BINARY::Solver's ~90 synchronous
// cudaMalloc/cudaMemcpy/cudaFree calls per segment plus one host round trip per BFS level
// (lib/PB_lib/src/pbnet/binary.cu).  Here ALL segments of ALL calls go through one fixed sequence
// of ~25 launches per chunk (two chunk streams for large batches); the host reads three integers (the cell extents that
// size the sort keys) in between and synchronises once at the end.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pbnet_b200.h"
#include "pb_kernels.cuh"
#include "pb_fused.cuh"
#include "pb_small.cuh"
#ifndef PB_TMA_CARVEOUT
#define PB_TMA_CARVEOUT 60   // per cent of the SM's 228 KB kept as shared memory for k_degree_tma (nine CTAs x 13.5 KB)
#endif

namespace {

// binary.cu:229 (class mean sizes "from HAIS"); filter threshold = mean_count * para_f in fp32
const float kMeanCount[18] = {3917.0f, 12056.0f, 2303.0f, 8331.0f, 3948.0f, 3166.0f, 5629.0f, 11719.0f, 1003.0f,
                              3317.0f, 4912.0f,  10221.0f, 3889.0f, 4136.0f, 2120.0f, 945.0f,  3967.0f, 2589.0f};

enum Stage {
    ST_H2D = 0, ST_PREP, ST_SORT, ST_GRID, ST_DEGREE, ST_HP, ST_UNION, ST_COMPONENTS, ST_LABEL, ST_FILTER,
    ST_LP_BUILD, ST_LP_NN, ST_CENTRES, ST_D2H, ST_COUNT
};
const char *kStageNames[ST_COUNT] = {"h2d", "prep", "sort", "grid", "degree", "hp_cells", "union", "components",
                                     "label", "filter", "lp_build", "lp_nn", "centres", "d2h"};

struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0;
    bool dry = false;
    template <class T>
    T *get(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        char *p = dry ? nullptr : base + off;
        off += bytes;
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace

struct pb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Arena arena;
    std::string err;
    int64_t launches = 0;
    cudaStream_t last_stream = nullptr;  // stream of the last call that used the arena
    bool profiling = false;
    cudaEvent_t ev[ST_COUNT + 1] = {};
    float stage_ms[ST_COUNT] = {};
    int64_t counters[16] = {};
    // pinned scratch for scalars read back at the end of a call
    int *h_small = nullptr;    // pinned [64]: scalars + phase time stamps of the small-call kernel
    int *h_scalars = nullptr;  // [0] err bits [1] C [2] R [3] K [4] Q [5] L
    unsigned long long *h_counters = nullptr;
    // chunked two-stream pipelining of large batched calls
    long long chunk_points = 0;  // 0 = automatic
    static constexpr int kMaxSlots = 4;  // chunk streams / private work areas in flight
    cudaStream_t aux[kMaxSlots] = {}, auxp[kMaxSlots] = {};
    int slots_host = 2, slots_dev = 2;   // PB_SLOTS_HOST / PB_SLOTS_DEV: work areas (and streams) the chunks of a call rotate over
    int dev_chunks = 2;                  // PB_DEV_CHUNKS: chunks of a device-resident call of >= 6 M points
                                         // (measured at C1 with PB_TIMELINE: the k_degree kernels of concurrent chunks run one after the
                                         //  other — each fills every SM — and the tails of two chunks overlap 1.4x; device data: 31.3 ms
                                         //  for 2 chunks / 2 slots and 3 / 3, 31.7 for 4 / 4, 32.1 for 4 / 2; host data: a third slot lets
                                         //  the last chunk's H2D and k_degree start before the middle chunk's and starves the earlier
                                         //  chunks' tails: 37.2-45.8 ms against 36.8 ms with two slots)
    int prio_mode = -1;  // -1 automatic (host data: prioritised pair), 0 never, 1 always
    bool stagger = false;  // PB_STAGGER=1: serialise k_degree of consecutive chunks (measured: 41.0 ms vs 39.9 ms in lock step at C1)
    int host_split[8] = {150, 350, 500, 0, 0, 0, 0, 0};  // PB_HOST_SPLIT: chunk sizes for host data of >= 6 M points, per mille
                             // (measured at C1, end to end: e2e ~ compute + first H2D + what is left of the last read-back.
                             //  one-sided k_degree: 45.9 ms with equal thirds, 44.1 ms at 15 / 42.5 / 42.5 %, 47-50 ms with 4-5 chunks;
                             //  symmetric k_degree: 38.2 ms at 15 / 42.5 / 42.5, 37.3 ms at 20 / 40 / 40, 38.4-40.2 ms with 4 chunks;
                             //  with the early read-backs (PB_EARLY_D2H) the last chunk may grow: 35.6 ms at 15 / 35 / 50,
                             //  35.9 at 12 / 33 / 55, 37.4 at 20 / 40 / 40, 37.7 without the early read-backs)
    int tile_mode = -1;      // PB_TILES: 1 = tiles of 1024 points, 0 = 4096, -1 = by problem size
    int label_ppw = 0;       // PB_LABEL_PPW: points per warp of k_label (0 = by problem size)
    int deg_smem = 0;        // PB_DEG_SMEM: unused dynamic shared memory requested for k_degree: caps its resident CTAs per SM so that
                             // registers stay free for the other chunk's latency-bound kernels (overlap experiment)
    int deg_slice_mult = 512; // PB_DEG_SLICES: warps per SM that k_degree aims for when it splits the candidate streams of a window over several
                             // warps (measured: one 224 k-point scene 271 us at 16, 211 us at 48, no change up to 512; C4, 1 M points in dense
                             // blobs: k_degree 1.53 ms at 48 (38 % of the SM time idle: whole windows are too coarse a unit), 0.81 at 256,
                             // 0.72 at 512 / 1024; 3.6 M points: 5.13 ms per step at 48, 5.05 at 512; C1 chunks are beyond the range)
    int devox_u = 4;           // PB_DEVOX_U: elements per thread of k_devox
    int small_deg_slices = 0;  // PB_SMALL_SLICES: candidate slices per query group in P1 of the small-call kernel (0: automatic)
    int small_tree_cap = pbsm::kTreeMax;  // PB_SMALL_TREES: P3 of the small-call kernel works on the tree graph up to this many trees (0: union-find always)
    int small_mode = -1;   // PB_SMALL=0: never use the small-call kernel; 1: whenever it is eligible; -1: automatic
    char *h_stage = nullptr;  // pinned staging of the small-call path (inputs in, results out: one copy each way)
    size_t h_stage_cap = 0;
    int coop_blocks_per_sm = 0, sm_count = 0;
    int deg_minb = 9;      // PB_DEG_MINB_SYM: resident CTAs per SM the symmetric k_degree is compiled for (8 = 64 registers, 9 = 56 with spills)
    bool deg_tma = true;   // PB_DEG_TMA=0: the symmetric k_degree reads its candidates with direct loads instead of the TMA-staged ring (A/B runs)
    bool deg_sym = true;   // PB_DEG_SYM=0: one-sided neighbour counting (every ordered pair tested; the round-1 formulation, kept for A/B runs)
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxSlots] = {};
    cudaStream_t copy_st = nullptr;  // host outputs that are final early (the degrees) leave on their own stream
    cudaEvent_t ev_hp[kMaxSlots] = {}, ev_degdone[kMaxSlots] = {};  // per work-area slot
    cudaEvent_t ev_ids[kMaxSlots] = {}, ev_idsdone[kMaxSlots] = {};
    bool deg_phased = true; // PB_DEG_PHASED=0: k_degree always walks the windows in order (1: heavy windows first on mid-size problems)
    int early_d2h = 3;     // PB_EARLY_D2H: bit 0 = degrees leave after k_hp_cells, bit 1 = ids leave before k_centres (host outputs)
    int *h_chunk_scalars = nullptr;
    unsigned long long *h_chunk_counters = nullptr;
    size_t h_chunk_cap = 0;
    std::vector<cudaEvent_t> chunk_ev;
    std::vector<cudaEvent_t> front_ev, deg_ev;  // per chunk: extents delivered / the always-on pair around k_degree
};

namespace {

int fail(pb_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define PB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

inline int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

// Stream of a call.  A NULL stream means "the context's own (non-blocking) stream" for host data; for DEVICE data it
// means the legacy default stream, so that the call is ordered after whatever produced the caller's buffers (the
// context stream is created with cudaStreamNonBlocking and would not wait for them).
inline cudaStream_t pick_stream(pb_ctx *ctx, void *stream_v, bool device_data) {
    if (stream_v) return (cudaStream_t)stream_v;
    return device_data ? cudaStreamLegacy : ctx->stream;
}

struct Scan {  // exclusive scan of int32, n known on host or (upper bound on host, exact on device)
    int *block_sums = nullptr;
    int nb = 0;
};

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int pb_create(int device, pb_ctx **out) {
    if (!out) return PB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return PB_ERR_CUDA;  // no CPU fallback exists
    pb_ctx *ctx = new pb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->h_scalars, 16 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_counters, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_small, 64 * sizeof(int)) != cudaSuccess) {
        delete ctx;
        return PB_ERR_CUDA;
    }
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    {
        // a second pair of chunk streams with different priorities, used when the data comes from the host: blocks of the
        // high-priority chunk's tail kernels are scheduled ahead of the other chunk's pending k_degree blocks, which
        // tightens the copy/compute pipeline (measured at C1: 55.1 -> 52.1 ms end to end with 3 chunks; device-resident
        // calls are 1 ms SLOWER with priorities and keep the plain pair)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int k = 0; k < pb_ctx::kMaxSlots; k++) {
            cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking);
            cudaStreamCreateWithPriority(&ctx->auxp[k], cudaStreamNonBlocking, (k & 1) ? lo : hi);
        }
        const char *sh = getenv("PB_SLOTS_HOST"), *sd = getenv("PB_SLOTS_DEV"), *dc = getenv("PB_DEV_CHUNKS");
        if (sh && atoi(sh) >= 1 && atoi(sh) <= pb_ctx::kMaxSlots) ctx->slots_host = atoi(sh);
        if (sd && atoi(sd) >= 1 && atoi(sd) <= pb_ctx::kMaxSlots) ctx->slots_dev = atoi(sd);
        if (dc && atoi(dc) >= 1 && atoi(dc) <= 8) ctx->dev_chunks = atoi(dc);
        const char *e = getenv("PB_STREAM_PRIO");  // experiments: 0 = never, 1 = always
        ctx->prio_mode = e ? (e[0] == '0' ? 0 : 1) : -1;
        const char *sg = getenv("PB_STAGGER");
        ctx->stagger = sg && sg[0] == '1';
        const char *hs = getenv("PB_HOST_SPLIT");
        if (hs) {
            int k = 0;
            for (const char *q = hs; *q && k < 8;) {
                ctx->host_split[k++] = atoi(q);
                while (*q && *q != ',') q++;
                if (*q == ',') q++;
            }
        }
        const char *tm = getenv("PB_TILES");
        if (tm) ctx->tile_mode = tm[0] == '1' ? 1 : (tm[0] == '0' ? 0 : -1);
        const char *lp = getenv("PB_LABEL_PPW");
        if (lp && atoi(lp) > 0) ctx->label_ppw = atoi(lp);
        const char *dm = getenv("PB_DEG_SMEM");
        if (dm && atoi(dm) > 0) ctx->deg_smem = atoi(dm);
        const char *ds = getenv("PB_DEG_SLICES");
        if (ds && atoi(ds) > 0) ctx->deg_slice_mult = atoi(ds);
        const char *sm = getenv("PB_SMALL");
        ctx->small_mode = sm ? (sm[0] == '0' ? 0 : 1) : -1;
        const char *du = getenv("PB_DEVOX_U");
        if (du) ctx->devox_u = atoi(du);
        const char *ssl = getenv("PB_SMALL_SLICES");
        if (ssl) ctx->small_deg_slices = std::max(0, atoi(ssl));
        const char *stc = getenv("PB_SMALL_TREES");
        if (stc) ctx->small_tree_cap = std::max(0, std::min(atoi(stc), (int)pbsm::kTreeMax));
        const char *dsy = getenv("PB_DEG_SYM");
        if (dsy) ctx->deg_sym = dsy[0] != '0';
        const char *dtm = getenv("PB_DEG_TMA");
        if (dtm) ctx->deg_tma = dtm[0] != '0';
        const char *dp = getenv("PB_DEG_PHASED");
        if (dp) ctx->deg_phased = dp[0] != '0';
        const char *ed = getenv("PB_EARLY_D2H");
        if (ed) ctx->early_d2h = atoi(ed) & 3;
        const char *dmb = getenv("PB_DEG_MINB_SYM");
        if (dmb && atoi(dmb) >= 8 && atoi(dmb) <= 9) ctx->deg_minb = atoi(dmb);
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    for (int k = 0; k < pb_ctx::kMaxSlots; k++) cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking);
    for (int k = 0; k < pb_ctx::kMaxSlots; k++) {
        cudaEventCreateWithFlags(&ctx->ev_hp[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_degdone[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_ids[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_idsdone[k], cudaEventDisableTiming);
    }
    *out = ctx;
    return PB_OK;
}

extern "C" void pb_destroy(pb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    for (auto &ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->chunk_ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->front_ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->deg_ev) cudaEventDestroy(ev);
    if (ctx->h_chunk_scalars) cudaFreeHost(ctx->h_chunk_scalars);
    if (ctx->h_chunk_counters) cudaFreeHost(ctx->h_chunk_counters);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int k = 0; k < pb_ctx::kMaxSlots; k++) {
        if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
        if (ctx->ev_hp[k]) cudaEventDestroy(ctx->ev_hp[k]);
        if (ctx->ev_degdone[k]) cudaEventDestroy(ctx->ev_degdone[k]);
        if (ctx->ev_ids[k]) cudaEventDestroy(ctx->ev_ids[k]);
        if (ctx->ev_idsdone[k]) cudaEventDestroy(ctx->ev_idsdone[k]);
        if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
        if (ctx->auxp[k]) cudaStreamDestroy(ctx->auxp[k]);
    }
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *pb_last_error(const pb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t pb_last_launch_count(const pb_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void pb_set_profiling(pb_ctx *ctx, int on) {
    if (ctx) ctx->profiling = on != 0;
}
extern "C" void pb_set_chunk_points(pb_ctx *ctx, int64_t points) {
    if (ctx) ctx->chunk_points = points;
}
extern "C" void pb_set_small_calls(pb_ctx *ctx, int mode) {
    if (ctx) ctx->small_mode = mode < 0 ? -1 : (mode ? 1 : 0);
}
extern "C" int pb_selftest_division(pb_ctx *ctx, int64_t n_samples, int64_t seed, int64_t *mismatches) {
    if (!ctx || !mismatches || n_samples < 0) return PB_ERR_ARG;
    PB_CUDA(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    PB_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    PB_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
    pb::k_selftest_division<<<148 * 8, 256, 0, ctx->stream>>>((unsigned long long)n_samples, (unsigned long long)seed, d);
    unsigned long long h = 0;
    PB_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    PB_CUDA(cudaFree(d));
    *mismatches = (int64_t)h;
    return PB_OK;
}
extern "C" int pb_stage_count(void) { return ST_COUNT; }
extern "C" const char *pb_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
extern "C" float pb_stage_ms(const pb_ctx *ctx, int i) { return (ctx && i >= 0 && i < ST_COUNT) ? ctx->stage_ms[i] : 0.f; }
extern "C" int64_t pb_counter(const pb_ctx *ctx, int i) { return (ctx && i >= 0 && i < 16) ? ctx->counters[i] : 0; }

// ------------------------------------------------------------------------------------------------
// ====================================================================================================
// grouping path: workspace, launch sequence
// ====================================================================================================
namespace {

#ifndef PB_TILE_ITEMS
#define PB_TILE_ITEMS 16
#endif
constexpr int kTileItems = PB_TILE_ITEMS;          // items per thread of the tile kernels (large path)
constexpr int kTile = pb::kTB * kTileItems;        // 4096 points per tile
constexpr int kTileItemsSmall = 4;                 // below 256 k points per chunk: tiles of 1024 points

// Host-built, segment-aligned tile table of one chunk (pb::TileTab): [begin | count | seg | first | hslot | srow] x T
struct TileHost {
    std::vector<int> buf;
    int T = 0, slots = 0, rows = 0;
};

void build_tiles(const int *start, int S, int tile, TileHost &th) {
    int T = 0;
    for (int s = 0; s < S; s++) T += (start[s + 1] - start[s] + tile - 1) / tile;
    th.T = T, th.slots = 0, th.rows = 0;
    th.buf.assign((size_t)6 * std::max(T, 1), 0);
    int *begin = th.buf.data(), *count = begin + T, *seg = count + T, *first = seg + T, *hslot = first + T, *srow = hslot + T;
    int t = 0;
    for (int s = 0; s < S; s++) {
        int n = start[s + 1] - start[s];
        int nt = (n + tile - 1) / tile;
        int slot = nt >= 2 ? th.slots++ : -1;
        for (int k = 0; k < nt; k++, t++) {
            begin[t] = start[s] + k * tile;
            count[t] = std::min(tile, n - k * tile);
            seg[t] = s;
            first[t] = t - k;
            hslot[t] = slot;
            srow[t] = nt >= 2 ? th.rows++ : -1;
        }
    }
}

struct Work {  // all device buffers of one chunk; laid out by plan() on the arena
    // inputs / outputs (device copies when the caller's data is on the host)
    float *x, *y, *z, *xo, *yo, *zo;
    int *sem;
    int *cluster_id, *cluster_num, *degree;
    // header block (one H2D copy): seg_start[S+1] | call_first[S] | tile table 6T | radius 18 | thresh 18 | min_pts 18
    int *hdr;
    size_t hdr_cap;
    pb::SegArrays sg;
    // per point
    int *seg_of, *fcell_of, *row_of, *deg_sorted, *raw_label, *flag, *gid_at, *lpos, *qlist, *inv2;
    uint64_t *key1[2];
    uint32_t *key2[2], *v1[2], *v2[2];
    float4 *pts4, *lab4, *box_lo, *box_hi, *box2_lo, *box2_hi;
    float *sx, *sy, *sz;
    // per fine cell / coarse cell (upper bound N)
    int *fcell_start, *fcell_cc, *cc_pstart, *cc_fstart, *parent, *cell_hp, *cell_minhp, *comp_min, *cell_gid, *cell_first;
    uint64_t *fcell_key, *cc_key;
    int2 *runs9;
    // per raw cluster (upper bound N)
    int *rep, *raw_count, *keep, *kscan, *clt_seg;
    // mixed-class mode only: per (fine cell, class) and per (segment, class) tables
    int *cell_min18, *comp_min18, *cell_gid18, *cnt18, *pos18, *lab_start18, *seg_lastlab;
    // scalars: [0] err [1] F [2] R [3] K [4] Q [5] L [6] Cc [7] rows [8] mixed scan total [9] centre ticket [10..12] extents
    int *d_scalars;
    unsigned long long *d_counters;
    unsigned long long *scan_state;  // k_scan_onepass: one word per tile + the ticket counter behind them
    size_t scan_tiles;
    unsigned long long *gridA, *gridB, *rel_state;  // look-back words of k_grid_build / k_relabel_scan [T]
    int *tickets;                    // [16] tile tickets: 0..9 sort passes (sort*5 + pass), 10 grid build, 11 relabel
    unsigned *sort_state;            // [(P1+P2)][rows][kBins]
    unsigned *hist;                  // [slots][(P1+P2)*kBins]
    // zero-initialised region A (before the front kernels) and B (sort state + histograms, sized after the key layout is known)
    char *zeroA_begin, *zeroA_end, *ffA_begin, *ffA_end, *zeroB_begin;
    size_t zeroB_cap;
};

// T_max / rows_max / slots_max: bounds of the tile table of any chunk planned on this slot
void plan(Arena &a, Work &w, long long n, int S, int T_max, int rows_max, int slots_max, bool host_io, bool mixed) {
    size_t N = (size_t)n;
    if (mixed) {
        w.cell_min18 = a.get<int>(N * pb::kCls); w.comp_min18 = a.get<int>(N * pb::kCls); w.cell_gid18 = a.get<int>(N * pb::kCls);
        w.pos18 = a.get<int>((size_t)S * pb::kCls + 1); w.lab_start18 = a.get<int>((size_t)S * pb::kCls + 1);
    }
    if (host_io) {
        w.x = a.get<float>(N); w.y = a.get<float>(N); w.z = a.get<float>(N);
        w.xo = a.get<float>(N); w.yo = a.get<float>(N); w.zo = a.get<float>(N);
        w.sem = a.get<int>(N);
        w.cluster_id = a.get<int>(N);
        w.cluster_num = a.get<int>(S);
        w.degree = a.get<int>(N);
    }
    w.hdr_cap = (size_t)2 * S + 1 + (size_t)6 * std::max(T_max, 1) + 54;
    w.hdr = a.get<int>(w.hdr_cap);
    w.clt_seg = a.get<int>(N);
    w.sg.cls = a.get<int>(S); w.sg.min_pts = a.get<int>(S); w.sg.r2 = a.get<float>(S); w.sg.inv_h = a.get<float>(S);
    w.sg.min_s = a.get<float>(3 * S); w.sg.min_o = a.get<float>(3 * S); w.sg.inv_g = a.get<float>(S);
    w.sg.cc_start = a.get<int>(S + 1); w.sg.cc_end = a.get<int>(S + 1); w.sg.id_base = a.get<int>(S);
    w.sg.k_base = a.get<int>(S); w.sg.cluster_num = a.get<int>(S);
    w.seg_of = a.get<int>(N); w.fcell_of = a.get<int>(N); w.row_of = a.get<int>(N); w.deg_sorted = a.get<int>(N);
    w.raw_label = a.get<int>(N); w.gid_at = a.get<int>(N);
    w.lpos = a.get<int>(N); w.qlist = a.get<int>(N); w.inv2 = a.get<int>(N);
    for (int k = 0; k < 2; k++) {
        w.key1[k] = a.get<uint64_t>(N); w.key2[k] = a.get<uint32_t>(N); w.v1[k] = a.get<uint32_t>(N); w.v2[k] = a.get<uint32_t>(N);
    }
    w.pts4 = a.get<float4>(N); w.lab4 = a.get<float4>(N);
    w.sx = a.get<float>(N + 2); w.sy = a.get<float>(N + 2); w.sz = a.get<float>(N + 2);
    w.box_lo = a.get<float4>(N / 32 + 2); w.box_hi = a.get<float4>(N / 32 + 2);
    w.box2_lo = a.get<float4>(N / 1024 + 2); w.box2_hi = a.get<float4>(N / 1024 + 2);
    w.fcell_start = a.get<int>(N + 1); w.fcell_cc = a.get<int>(N); w.cc_pstart = a.get<int>(N + 1); w.cc_fstart = a.get<int>(N + 1);
    w.parent = a.get<int>(N); w.cell_hp = a.get<int>(N); w.cell_minhp = a.get<int>(N); w.cell_first = a.get<int>(N);
    w.comp_min = a.get<int>(N); w.cell_gid = a.get<int>(N); w.fcell_key = a.get<uint64_t>(N); w.cc_key = a.get<uint64_t>(N);
    w.runs9 = a.get<int2>(N * pb::kRuns);
    w.rep = a.get<int>(N); w.keep = a.get<int>(N); w.kscan = a.get<int>(N);
    // ---- 0xff-initialised region: encoded minima (+ mixed: last labelled point per segment)
    w.ffA_begin = reinterpret_cast<char *>(a.get<char>(0));
    w.sg.enc_min_s = a.get<unsigned>(3 * S); w.sg.enc_min_o = a.get<unsigned>(3 * S);
    w.seg_lastlab = a.get<int>(mixed ? S : 1);
    w.ffA_end = reinterpret_cast<char *>(a.get<char>(0));
    // ---- zero-initialised region A
    w.zeroA_begin = reinterpret_cast<char *>(a.get<char>(0));
    w.d_scalars = a.get<int>(32);
    w.d_counters = a.get<unsigned long long>(8);
    w.tickets = a.get<int>(16);
    w.sg.enc_max_s = a.get<unsigned>(3 * S); w.sg.enc_max_o = a.get<unsigned>(3 * S);
    w.flag = a.get<int>(N); w.raw_count = a.get<int>(N);
    w.scan_tiles = std::max(N, (size_t)S * pb::kCls) / pb::kScanTile + 2;
    w.scan_state = a.get<unsigned long long>(w.scan_tiles + 1);
    w.gridA = a.get<unsigned long long>(std::max(T_max, 1)); w.gridB = a.get<unsigned long long>(std::max(T_max, 1));
    w.rel_state = a.get<unsigned long long>(std::max(T_max, 1));
    w.cnt18 = a.get<int>(mixed ? (size_t)S * pb::kCls + 1 : 1);
    w.zeroA_end = reinterpret_cast<char *>(a.get<char>(0));
    // ---- zero-initialised region B: sort look-back state + digit histograms for up to kMaxPassesGroup passes
    w.zeroB_begin = reinterpret_cast<char *>(a.get<char>(0));
    w.zeroB_cap = ((size_t)std::max(rows_max, 1) + (size_t)std::max(slots_max, 1)) * pb::kMaxPassesGroup * pb::kBins * sizeof(unsigned);
    a.get<char>(w.zeroB_cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
static void invalidate_scene_state();  // the arena is about to be reused (pb_local_scenes_plan keeps pointers into it)
namespace {
int ensure_arena(pb_ctx *ctx, size_t need, cudaStream_t st) {
    invalidate_scene_state();
    // The arena is one scratch area reset per call: work still pending on ANOTHER stream (an earlier asynchronous call)
    // must finish before this call overwrites it.  Calls that stay on one stream are ordered by the stream itself.
    if (ctx->last_stream && ctx->last_stream != st) PB_CUDA(cudaStreamSynchronize(ctx->last_stream));
    ctx->last_stream = st;
    if (need > ctx->arena.cap) {
        PB_CUDA(cudaStreamSynchronize(st));
        if (ctx->arena.base) PB_CUDA(cudaFree(ctx->arena.base));
        ctx->arena.base = nullptr;
        ctx->arena.cap = 0;
        size_t want = need + need / 4;
        if (cudaMalloc(&ctx->arena.base, want) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, PB_ERR_NOMEM, "workspace allocation of " + std::to_string(want) + " bytes failed");
        }
        ctx->arena.cap = want;
    }
    ctx->arena.off = 0;
    ctx->arena.dry = false;
    return PB_OK;
}

void launch_scan(cudaStream_t st, const int *in, int n_host, const int *n_dev, int *out, int *total, int *blocks, int64_t &L) {
    int nb = div_up(n_host, pb::kScanTile);
    pb::k_scan_reduce<<<nb, pb::kScanThreads, 0, st>>>(in, n_host, n_dev, blocks);
    pb::k_scan_spine<<<1, pb::kScanThreads, 0, st>>>(blocks, nb, total);
    pb::k_scan_down<<<nb, pb::kScanThreads, 0, st>>>(in, n_host, n_dev, blocks, out);
    L += 3;
}
}  // namespace


// ----------------------------------------------------------------------------------------------------------------
// Generic stable LSD radix sort of n (key[, payload]) pairs over key bits [0, end_bit) with the hand-written tile sort
// (pb_sort.cuh, one segment): what voxelize, the local scenes, the evaluation post-processing and the mesh normals use
// (thrust / CUB in the reference and in round 1).  Inputs are preserved; the result lands in key_out / val_out.
// ----------------------------------------------------------------------------------------------------------------
namespace {

struct RadixTmp {   // scratch of ONE sort, carved from caller-provided device memory
    static size_t bytes(size_t n, size_t key_size) {
        const size_t T = (n + kTile - 1) / kTile + 1;
        auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
        return al(n * key_size) + al(n * 4) * 2 + al((6 * T + 2) * 4) + al((size_t)pb::kMaxPasses * pb::kBins * 4) +
               al((size_t)pb::kMaxPasses * T * pb::kBins * 4) + al(64) + 1024;
    }
};

template <typename KeyT>
int radix_sort(pb_ctx *ctx, void *tmp, const KeyT *key_in, KeyT *key_out, const uint32_t *val_in, uint32_t *val_out, int n,
               int end_bit, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return PB_OK;
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    TileHost th;
    const int start2[2] = {0, n};
    build_tiles(start2, 1, kTile, th);
    const int T = th.T;
    char *p = reinterpret_cast<char *>(tmp);
    KeyT *key_tmp = reinterpret_cast<KeyT *>(p); p += al((size_t)n * sizeof(KeyT));
    uint32_t *val_tmp = reinterpret_cast<uint32_t *>(p); p += al((size_t)n * 4);
    uint32_t *val_dummy = reinterpret_cast<uint32_t *>(p); p += al((size_t)n * 4);
    int *tile_dev = reinterpret_cast<int *>(p); p += al(((size_t)6 * T + 2) * 4);
    char *zero_begin = p;
    unsigned *hist = reinterpret_cast<unsigned *>(p); p += al((size_t)pb::kMaxPasses * pb::kBins * 4);
    const pb::PassPlan plan = pb::make_pass_plan(std::max(1, std::min(end_bit, (int)sizeof(KeyT) * 8)));
    unsigned *state = reinterpret_cast<unsigned *>(p); p += al((size_t)plan.npass * std::max(th.rows, 1) * pb::kBins * 4);
    int *tickets = reinterpret_cast<int *>(p); p += al(64);
    if (!val_out) val_out = val_dummy;   // keys only: the payloads still travel (into scratch)
    std::vector<int> hdr((size_t)6 * T + 2);
    std::memcpy(hdr.data(), th.buf.data(), sizeof(int) * (size_t)6 * T);
    hdr[6 * T] = 0, hdr[6 * T + 1] = n;
    PB_CUDA(cudaMemcpyAsync(tile_dev, hdr.data(), sizeof(int) * hdr.size(), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemsetAsync(zero_begin, 0, p - zero_begin, st));
    pb::TileTab tt;
    tt.begin = tile_dev, tt.count = tile_dev + T, tt.seg = tile_dev + 2 * T, tt.first = tile_dev + 3 * T, tt.hslot = tile_dev + 4 * T;
    tt.srow = tile_dev + 5 * T, tt.T = T;
    const int *seg_start = tile_dev + 6 * T;
    static bool attr_done[2] = {false, false};
    if (!attr_done[sizeof(KeyT) == 8]) {
        cudaFuncSetAttribute(pb::k_sort_pass<KeyT, kTileItems>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<KeyT, kTileItems>));
        attr_done[sizeof(KeyT) == 8] = true;
    }
    if (T > 1) {
        pb::k_radix_hist<KeyT><<<std::min(div_up(n, pb::kTB * 8), 148 * 8), pb::kTB, 0, st>>>(key_in, n, plan, hist);
        if (launches) (*launches)++;
    }
    // ping-pong between the scratch buffer and the output so that the LAST pass writes the output; the input is only read
    const KeyT *kin = key_in;
    const uint32_t *vin = val_in;
    for (int ps = 0; ps < plan.npass; ps++) {
        const bool to_out = ((plan.npass - 1 - ps) & 1) == 0;
        pb::SortArgs<KeyT> a;
        a.keys_in = kin, a.keys_out = to_out ? key_out : key_tmp, a.vals_in = vin, a.vals_out = to_out ? val_out : val_tmp;
        a.hist = hist, a.hist_stride = plan.npass * pb::kBins, a.hist_off = ps * pb::kBins;
        a.state = state + (size_t)ps * std::max(th.rows, 1) * pb::kBins, a.ticket = tickets + ps;
        a.shift = plan.shift[ps], a.width = plan.width[ps];
        pb::k_sort_pass<KeyT, kTileItems><<<dim3(T, 1), pb::kTB, sizeof(pb::SortSmem<KeyT, kTileItems>), st>>>(a, a, tt, seg_start);
        if (launches) (*launches)++;
        kin = a.keys_out, vin = a.vals_out;
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace

namespace {

struct ChunkIO {          // one chunk = a run of consecutive calls; all pointers already offset to the chunk
    const float *x, *y, *z, *xo, *yo, *zo;
    const int *sem;
    int *cluster_id, *cluster_num, *degree;
    int n, S;
    const int *start;       // host [S+1], local to the chunk
    const int *call_first;  // host [S], local to the chunk
    const TileHost *tiles;
    float *center_out;      // device, capacity 3*n floats (chunk-private region)
    int *clt_sem_out;       // device, capacity n
    int *h_scalars;         // pinned [16]
    unsigned long long *h_counters;  // pinned [8]
    cudaEvent_t *ev;        // ST_COUNT+1 events or nullptr
    cudaEvent_t ev_front;   // recorded after the extents are on their way to the host
    cudaEvent_t ev_deg[2];  // always-on pair around k_degree
    cudaEvent_t wait_deg;   // k_degree of the previous chunk (other stream) finished, or nullptr
    int slot;               // work-area slot of the chunk (its early-copy events)
};

struct ChunkDev {  // device views of the header block, valid after enqueue_front
    int *seg_start, *call_first;
    pb::TileTab tt;
    float *radius, *thresh;
    int *min_pts;
};

ChunkDev header_views(const Work &w, int S, int T) {
    ChunkDev d;
    int *p = w.hdr;
    d.seg_start = p, p += S + 1;
    d.call_first = p, p += S;
    const int Tz = std::max(T, 1);
    d.tt.begin = p, d.tt.count = p + Tz, d.tt.seg = p + 2 * Tz, d.tt.first = p + 3 * Tz, d.tt.hslot = p + 4 * Tz, d.tt.srow = p + 5 * Tz;
    d.tt.T = T;
    p += 6 * Tz;
    d.radius = reinterpret_cast<float *>(p), p += 18;
    d.thresh = reinterpret_cast<float *>(p), p += 18;
    d.min_pts = p;
    return d;
}

// Front of a chunk: copies, initialisation, validation + bounding boxes, per-segment parameters, cell extents -> host.
// ITEMS = keys per thread of the tile kernels: 16 (tiles of 4096 points) on large problems, 4 (1024 points) below 256 k
// points, where the per-tile latency chain — not the bandwidth — sets the time of the tile kernels.
template <int ITEMS>
int enqueue_front(pb_ctx *ctx, Work &w, const ChunkIO &io, bool host_io, bool mixed, const float *radius, const int *min_pts,
                  const float *thresh, cudaStream_t st, int64_t &L, std::vector<int> &hdr_host) {
    const int n = io.n, S = io.S, T = io.tiles->T;
    if (io.ev) cudaEventRecord(io.ev[ST_H2D], st);
    if (host_io) {
        size_t fb = sizeof(float) * (size_t)n;
        PB_CUDA(cudaMemcpyAsync(w.x, io.x, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.y, io.y, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.z, io.z, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.xo, io.xo, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.yo, io.yo, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.zo, io.zo, fb, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(w.sem, io.sem, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    // header block: pageable host memory -> staged synchronously by the driver, safe to reuse after return
    const int Tz = std::max(T, 1);
    hdr_host.assign((size_t)2 * S + 1 + (size_t)6 * Tz + 54, 0);
    {
        int *p = hdr_host.data();
        std::memcpy(p, io.start, sizeof(int) * (S + 1)), p += S + 1;
        std::memcpy(p, io.call_first, sizeof(int) * S), p += S;
        if (T > 0) std::memcpy(p, io.tiles->buf.data(), sizeof(int) * (size_t)6 * T);
        p += 6 * Tz;
        std::memcpy(p, radius, 18 * sizeof(float)), p += 18;
        std::memcpy(p, thresh, 18 * sizeof(float)), p += 18;
        std::memcpy(p, min_pts, 18 * sizeof(int));
    }
    PB_CUDA(cudaMemcpyAsync(w.hdr, hdr_host.data(), sizeof(int) * hdr_host.size(), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemsetAsync(w.ffA_begin, 0xff, w.ffA_end - w.ffA_begin, st));
    PB_CUDA(cudaMemsetAsync(w.zeroA_begin, 0, w.zeroA_end - w.zeroA_begin, st));
    if (mixed) {  // per (fine cell, class) tables start at "no HP"
        PB_CUDA(cudaMemsetAsync(w.cell_min18, 0x7f, sizeof(int) * (size_t)n * pb::kCls, st));
        PB_CUDA(cudaMemsetAsync(w.comp_min18, 0x7f, sizeof(int) * (size_t)n * pb::kCls, st));
    }
    ChunkDev d = header_views(w, S, T);
    w.sg.start = d.seg_start;
    const float *dx = host_io ? w.x : io.x, *dy = host_io ? w.y : io.y, *dz = host_io ? w.z : io.z;
    const float *dxo = host_io ? w.xo : io.xo, *dyo = host_io ? w.yo : io.yo, *dzo = host_io ? w.zo : io.zo;
    const int *dsem = host_io ? w.sem : io.sem;
    if (io.ev) cudaEventRecord(io.ev[ST_PREP], st);
    pb::k_prep<ITEMS><<<std::min(Tz, 148 * 16), pb::kTB, 0, st>>>(d.tt, w.sg, dx, dy, dz, dxo, dyo, dzo, dsem, w.seg_of, w.d_scalars);
    pb::k_seg_params<<<div_up(std::max(S, 1), 256), 256, 0, st>>>(S, w.sg, dsem, d.radius, d.min_pts, w.d_scalars + 10, w.d_scalars);
    L += 2;
    PB_CUDA(cudaMemcpyAsync(io.h_scalars + 10, w.d_scalars + 10, sizeof(int) * 3, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaEventRecord(io.ev_front, st));
    return PB_OK;
}

template <typename KeyT, int ITEMS>
void launch_sort_passes(Work &w, const pb::TileTab &tt, const int *seg_start, const pb::PassPlan &p1, const pb::PassPlan &p2,
                        int rows, cudaStream_t st, int64_t &L, int which) {
    // which: 0 = sort 1 only (KeyT = its key type), 1 = sort 2 only, 2 = both (same key type)
    const int stride = (p1.npass + p2.npass) * pb::kBins;
    const size_t state_pass = (size_t)std::max(rows, 1) * pb::kBins;
    const int np = which == 0 ? p1.npass : (which == 1 ? p2.npass : std::max(p1.npass, p2.npass));
    const size_t smem = sizeof(pb::SortSmem<KeyT, ITEMS>);
    for (int p = 0; p < np; p++) {
        pb::SortArgs<KeyT> a[2];
        int na = 0;
        if (which != 1 && p < p1.npass) {
            pb::SortArgs<KeyT> &q = a[na++];
            q.keys_in = reinterpret_cast<const KeyT *>(w.key1[p & 1]), q.keys_out = reinterpret_cast<KeyT *>(w.key1[(p + 1) & 1]);
            q.vals_in = p ? w.v1[p & 1] : nullptr, q.vals_out = w.v1[(p + 1) & 1];
            q.hist = w.hist, q.hist_stride = stride, q.hist_off = p * pb::kBins;
            q.state = w.sort_state + (size_t)p * state_pass, q.ticket = w.tickets + p;
            q.shift = p1.shift[p], q.width = p1.width[p];
        }
        if (which != 0 && p < p2.npass) {
            pb::SortArgs<KeyT> &q = a[na++];
            q.keys_in = reinterpret_cast<const KeyT *>(w.key2[p & 1]), q.keys_out = reinterpret_cast<KeyT *>(w.key2[(p + 1) & 1]);
            q.vals_in = p ? w.v2[p & 1] : nullptr, q.vals_out = w.v2[(p + 1) & 1];
            q.hist = w.hist, q.hist_stride = stride, q.hist_off = (p1.npass + p) * pb::kBins;
            q.state = w.sort_state + (size_t)(p1.npass + p) * state_pass, q.ticket = w.tickets + 5 + p;
            q.shift = p2.shift[p], q.width = p2.width[p];
        }
        if (na == 0) continue;
        if (na == 1) a[1] = a[0];
        pb::k_sort_pass<KeyT, ITEMS><<<dim3(tt.T, na), pb::kTB, smem, st>>>(a[0], a[1], tt, seg_start);
        L++;
    }
}

// Everything after the key layout is known.  No host synchronisation.
template <bool MIXED, int ITEMS>
int enqueue_rest(pb_ctx *ctx, Work &w, const ChunkIO &io, bool host_io, int assign_lp, cudaStream_t st, int64_t &L) {
    const int n = io.n, S = io.S, T = io.tiles->T;
    const bool prof = io.ev != nullptr;
    int stage_ev = ST_SORT;
    auto mark = [&]() {
        if (prof) cudaEventRecord(io.ev[stage_ev], st);
        stage_ev++;
    };
    const int T256 = 256;
    const int gS = div_up(S + 1, T256);
    // grid-stride kernels: one CTA wave on large problems, no more CTAs than warp-sized work items on small ones
    const int gPersist = std::min(148 * 8, std::max(1, div_up(n, 8)));
    ChunkDev d = header_views(w, S, T);
    const float *dx = host_io ? w.x : io.x, *dy = host_io ? w.y : io.y, *dz = host_io ? w.z : io.z;
    const float *dxo = host_io ? w.xo : io.xo, *dyo = host_io ? w.yo : io.yo, *dzo = host_io ? w.zo : io.zo;
    const int *dsem = host_io ? w.sem : io.sem;
    int *d_cluster_id = host_io ? w.cluster_id : io.cluster_id, *d_cluster_num = host_io ? w.cluster_num : io.cluster_num;
    int *d_degree = host_io ? w.degree : io.degree;
    int *d_err = w.d_scalars, *d_F = w.d_scalars + 1, *d_R = w.d_scalars + 2, *d_K = w.d_scalars + 3,
        *d_Q = w.d_scalars + 4, *d_L = w.d_scalars + 5, *d_Cc = w.d_scalars + 6, *d_rows = w.d_scalars + 7;
    unsigned long long *cnt = prof ? w.d_counters : nullptr;

    // ---- key layout from the extents the front delivered
    pb::KeysArgs ka;
    ka.lay = pb::make_key_layout(io.h_scalars[10], io.h_scalars[11], io.h_scalars[12]);
    ka.key64 = ka.lay.bits > 32;
    ka.plan1 = pb::make_pass_plan(ka.lay.bits);
    ka.plan2 = pb::make_pass_plan(assign_lp ? (MIXED ? 32 : 3 * pb::kMortonBits) : 1);
    if (!assign_lp) ka.plan2.npass = 0;
    const int npass_all = ka.plan1.npass + ka.plan2.npass;
    ka.hist_stride = npass_all * pb::kBins;
    const int rows = io.tiles->rows, slots = io.tiles->slots;
    const size_t state_words = (size_t)npass_all * std::max(rows, 1) * pb::kBins;
    w.sort_state = reinterpret_cast<unsigned *>(w.zeroB_begin);
    w.hist = w.sort_state + state_words;
    ka.hist = w.hist;
    const size_t zeroB = (state_words + (size_t)std::max(slots, 1) * ka.hist_stride) * sizeof(unsigned);
    if (zeroB > w.zeroB_cap) return fail(ctx, PB_ERR_ARG, "internal: sort state exceeds its plan");
    if (rows > 0 || slots > 0) PB_CUDA(cudaMemsetAsync(w.zeroB_begin, 0, zeroB, st));

    const int Tz = std::max(T, 1);
    pb::k_keys<MIXED, ITEMS><<<std::min(Tz, 148 * 16), pb::kTB, 0, st>>>(d.tt, w.sg, ka, dx, dy, dz, dxo, dyo, dzo, dsem, w.key1[0],
                                                                             w.key2[0], d_err, d.radius, w.cnt18);
    L++;
    mark();  // SORT
    if (ka.key64) {
        launch_sort_passes<uint64_t, ITEMS>(w, d.tt, d.seg_start, ka.plan1, ka.plan2, rows, st, L, 0);
        if (assign_lp) launch_sort_passes<uint32_t, ITEMS>(w, d.tt, d.seg_start, ka.plan1, ka.plan2, rows, st, L, 1);
    } else {
        launch_sort_passes<uint32_t, ITEMS>(w, d.tt, d.seg_start, ka.plan1, ka.plan2, rows, st, L, assign_lp ? 2 : 0);
    }
    const void *skey = w.key1[ka.plan1.npass & 1];
    const uint32_t *order1 = w.v1[ka.plan1.npass & 1];
    const uint32_t *order2 = w.v2[ka.plan2.npass & 1];

    mark();  // GRID
    pb::GridOut go;
    go.pts4 = w.pts4, go.sx = w.sx, go.sy = w.sy, go.sz = w.sz, go.fcell_of = w.fcell_of, go.row_of = w.row_of;
    go.fcell_start = w.fcell_start, go.fcell_cc = w.fcell_cc, go.cc_pstart = w.cc_pstart, go.cc_fstart = w.cc_fstart;
    go.parent = w.parent, go.cell_hp = w.cell_hp, go.cell_minhp = w.cell_minhp, go.comp_min = w.comp_min, go.cell_first = w.cell_first;
    go.fcell_key = w.fcell_key, go.cc_key = w.cc_key, go.d_F = d_F, go.d_Cc = d_Cc, go.d_rows = d_rows;
    go.stateA = w.gridA, go.stateB = w.gridB, go.ticket = w.tickets + 10;
    if (ka.key64)
        pb::k_grid_build<uint64_t, ITEMS><<<Tz, pb::kTB, 0, st>>>(d.tt, n, w.sg, ka.lay, reinterpret_cast<const uint64_t *>(skey), order1, dx, dy, dz, go);
    else
        pb::k_grid_build<uint32_t, ITEMS><<<Tz, pb::kTB, 0, st>>>(d.tt, n, w.sg, ka.lay, reinterpret_cast<const uint32_t *>(skey), order1, dx, dy, dz, go);
    pb::k_runs<<<gPersist, T256, 0, st>>>(w.sg, w.cc_key, d_Cc, w.runs9);
    L += 2;
    pb::Grid grid;
    grid.pts4 = w.pts4; grid.sx = w.sx; grid.sy = w.sy; grid.sz = w.sz; grid.fcell_of = w.fcell_of; grid.row_of = w.row_of; grid.fcell_start = w.fcell_start;
    grid.fcell_key = w.fcell_key; grid.fcell_cc = w.fcell_cc; grid.cc_pstart = w.cc_pstart; grid.cc_fstart = w.cc_fstart;
    grid.cc_key = w.cc_key; grid.runs9 = w.runs9; grid.d_F = d_F; grid.d_Cc = d_Cc;

    mark();  // DEGREE
    // staggered pipeline: the neighbour-count kernels of consecutive chunks run one after the other, so that the
    // latency-bound tail of chunk i overlaps k_degree of chunk i+1 instead of its own twin
    if (io.wait_deg) PB_CUDA(cudaStreamWaitEvent(st, io.wait_deg, 0));
    PB_CUDA(cudaEventRecord(io.ev_deg[0], st));
    {
        // small problems: several warps share one 128-point window and split its candidate stream
        const int windows = div_up(n, pb::kWindow);
        const int nslice = std::max(1, std::min(32, (148 * ctx->deg_slice_mult) / windows));
        // heavy windows first (two passes over the window list in one grid) on mid-size problems, where the drain of the last
        // dense windows shows (measured, 3.6 M points in two chunks: 5.28 -> 5.12 ms per step; on 14.4 M-point chunks the
        // heavy windows are better left mixed with the light ones: 31.4 -> 32.5 ms)
        const int phased = (nslice == 1 && ctx->deg_phased && windows <= 32768) ? 1 : 0;
        const dim3 g(div_up(n, pb::kWindow * 4) * (phased ? 2 : 1), nslice);
        // symmetric counting: candidates' degrees are accumulated with RED, so the array starts at zero
        if (nslice > 1 || ctx->deg_sym) PB_CUDA(cudaMemsetAsync(w.deg_sorted, 0, sizeof(int) * (size_t)n, st));
        const size_t dsm = nslice == 1 ? ctx->deg_smem : 0;
        if (ctx->deg_sym && ctx->deg_tma && nslice <= 4) {   // measured: the ring pays from ~2.4 M points (3.6 M: 5.10 -> 4.87 ms; 1.26 M sparse points, 7 warps per window: 0.56 -> 0.68 ms; 27 k-point call: 88 -> 110 us)
            static bool carveout_set = false;   // nine resident CTAs need 9 x 13.5 KB of shared memory
            if (!carveout_set) {
                cudaFuncSetAttribute(pb::k_degree_tma<9>, cudaFuncAttributePreferredSharedMemoryCarveout, PB_TMA_CARVEOUT);
                cudaFuncSetAttribute(pb::k_degree_tma<8>, cudaFuncAttributePreferredSharedMemoryCarveout, PB_TMA_CARVEOUT);
                carveout_set = true;
            }
            if (ctx->deg_minb == 9) pb::k_degree_tma<9><<<g, 128, dsm, st>>>(n, w.sg, grid, w.deg_sorted, cnt, phased);
            else pb::k_degree_tma<8><<<g, 128, dsm, st>>>(n, w.sg, grid, w.deg_sorted, cnt, phased);
        } else
        if (!ctx->deg_sym) pb::k_degree<false, PB_DEG_MINB><<<g, 128, dsm, st>>>(n, w.sg, grid, w.deg_sorted, cnt, phased);
        else if (ctx->deg_minb == 9) pb::k_degree<true, 9><<<g, 128, dsm, st>>>(n, w.sg, grid, w.deg_sorted, cnt, phased);
        else pb::k_degree<true, 8><<<g, 128, dsm, st>>>(n, w.sg, grid, w.deg_sorted, cnt, phased);
        PB_CUDA(cudaEventRecord(io.ev_deg[1], st));
        mark();  // HP
        pb::k_hp_cells<MIXED><<<div_up(n, T256 * pb::kHpPer), T256, 0, st>>>(n, w.sg, w.pts4, w.fcell_of, w.fcell_key, w.deg_sorted, d_degree, w.cell_hp,
                                                               w.cell_minhp, cnt, dsem, d.min_pts, w.cell_min18, w.cell_first);
        L++;
        if (host_io && (ctx->early_d2h & 1)) {  // the degrees are final: their read-back overlaps the rest of the chunk instead of trailing it
            PB_CUDA(cudaEventRecord(ctx->ev_hp[io.slot], st));
            PB_CUDA(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_hp[io.slot], 0));
            PB_CUDA(cudaMemcpyAsync(io.degree, w.degree, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->copy_st));
            PB_CUDA(cudaEventRecord(ctx->ev_degdone[io.slot], ctx->copy_st));
        }
    }
    L++;
    mark();  // UNION
    pb::k_union<<<gPersist, T256, 0, st>>>(w.sg, grid, w.pts4, w.cell_hp, w.parent, 0, w.cell_first);
    pb::k_union<<<gPersist, T256, 0, st>>>(w.sg, grid, w.pts4, w.cell_hp, w.parent, 1, w.cell_first);
    L += 2;
    mark();  // COMPONENTS
    unsigned scan_epoch = 0;
    auto scan = [&](const int *in, int n_host, const int *n_dev, int *out, int *total) {
        int nb = std::max(1, div_up(n_host, pb::kScanTile));
        pb::k_scan_onepass<<<nb, pb::kScanThreads, 0, st>>>(in, n_host, n_dev, out, total, w.scan_state,
                                                            reinterpret_cast<int *>(w.scan_state + w.scan_tiles), ++scan_epoch);
        L++;
    };
    pb::k_comp_min<MIXED><<<gPersist, T256, 0, st>>>(d_F, w.cell_hp, w.parent, w.cell_minhp, w.comp_min, w.cell_min18, w.comp_min18);
    pb::k_flag_roots<MIXED><<<gPersist, T256, 0, st>>>(d_F, w.cell_hp, w.parent, w.comp_min, w.flag, w.comp_min18);
    L += 2;
    scan(w.flag, n, nullptr, w.gid_at, d_R);
    pb::k_cell_gid<MIXED><<<gPersist, T256, 0, st>>>(d_F, w.cell_hp, w.parent, w.comp_min, w.gid_at, w.cell_gid, w.rep, w.comp_min18,
                                                    w.cell_gid18);
    L++;
    mark();  // LABEL
    {
        int ppw = n >= (1 << 17) ? 32 : (n >= (1 << 15) ? 16 : 8);   // measured on a 137 k-point scene: 131 / 123 / 117 / 106 us at 4 / 8 / 16 / 32
        if (ctx->label_ppw > 0) ppw = ctx->label_ppw;
        pb::k_label<MIXED><<<div_up(div_up(n, ppw), 4), 128, 0, st>>>(n, w.sg, grid, w.pts4, w.cell_hp, w.cell_gid, w.raw_label,
                                                                      w.raw_count, dsem, w.cell_gid18, ppw);
    }
    L++;
    mark();  // FILTER
    pb::k_filter_scan<MIXED><<<1, 1024, 0, st>>>(n, S, w.sg, d_R, w.rep, w.seg_of, w.raw_count, d.thresh, w.keep, w.kscan, d_K, dsem,
                                                d.call_first, w.gid_at, d_cluster_num);
    L++;
    mark();  // LP_BUILD
    {
        pb::RelabelOut ro;
        ro.cluster_id = d_cluster_id, ro.clt_sem = io.clt_sem_out, ro.clt_seg = w.clt_seg, ro.qlist = w.qlist, ro.inv2 = w.inv2;
        ro.lpos = w.lpos, ro.seg_lastlab = w.seg_lastlab, ro.lab4 = w.lab4, ro.d_Q = d_Q, ro.d_L = d_L, ro.state = w.rel_state;
        ro.ticket = w.tickets + 11;
        pb::k_relabel_scan<MIXED, ITEMS><<<Tz, pb::kTB, 0, st>>>(d.tt, n, w.sg, w.raw_label, w.keep, w.kscan, assign_lp, w.rep, dsem,
                                                                     order2, dxo, dyo, dzo, ro);
        L++;
    }
    if (assign_lp) {
        if (MIXED) {
            scan(w.cnt18, S * pb::kCls, nullptr, w.pos18, w.d_scalars + 8);
            pb::k_seg_lab18<<<div_up((long long)S * pb::kCls + 1, T256), T256, 0, st>>>(n, S, w.pos18, w.lpos, d_L, w.lab_start18);
            L++;
        }
        pb::k_lab_boxes<<<std::min(148 * 8, std::max(1, div_up(n, 1024))), pb::kTB, 0, st>>>(d_L, w.lab4, w.box_lo, w.box_hi, w.box2_lo, w.box2_hi);
        L++;
    }
    mark();  // LP_NN
    if (assign_lp) {
        pb::k_nn<MIXED><<<gPersist, T256, 0, st>>>(n, d_Q, d_L, w.sg, w.qlist, w.seg_of, w.inv2, w.lpos, dxo, dyo, dzo, w.lab4, w.box_lo,
                                                   w.box_hi, w.box2_lo, w.box2_hi, d_cluster_id, dsem, w.lab_start18, w.seg_lastlab);
        L++;
    }
    if (host_io && (ctx->early_d2h & 2)) {  // the ids are final: their read-back overlaps the centre replay
        PB_CUDA(cudaEventRecord(ctx->ev_ids[io.slot], st));
        PB_CUDA(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_ids[io.slot], 0));
        PB_CUDA(cudaMemcpyAsync(io.cluster_id, w.cluster_id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->copy_st));
        PB_CUDA(cudaEventRecord(ctx->ev_idsdone[io.slot], ctx->copy_st));
    }
    mark();  // CENTRES
    pb::k_centres<<<std::min(gPersist, 148 * 4), 256, 0, st>>>(d_K, w.sg, w.clt_seg, d_cluster_id, dx, dy, dz, io.center_out,
                                                            w.d_scalars + 9, cnt ? cnt + 4 : nullptr);
    L++;
    mark();  // D2H
    PB_CUDA(cudaMemcpyAsync(io.h_scalars, w.d_scalars, sizeof(int) * 10, cudaMemcpyDeviceToHost, st));
    if (prof) PB_CUDA(cudaMemcpyAsync(io.h_counters, w.d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
    if (host_io) {
        if (ctx->early_d2h & 2) PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_idsdone[io.slot], 0));  // issued earlier on the copy stream
        else PB_CUDA(cudaMemcpyAsync(io.cluster_id, w.cluster_id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (ctx->early_d2h & 1) PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_degdone[io.slot], 0));
        else PB_CUDA(cudaMemcpyAsync(io.degree, w.degree, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));  // issued after k_hp_cells on the copy stream
        PB_CUDA(cudaMemcpyAsync(io.cluster_num, w.cluster_num, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
    }
    mark();  // end
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------------
// Small calls (one call, few segments, a few thousand points each): ONE cooperative launch of pbsm::k_small, one copy
// in and one copy out through pinned staging when the data lives on the host.  Returns PB_OK, an error, or kNotSmall
// (not eligible / mixed classes found: the caller falls through to the general path).
// ----------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kNotSmall = -1;
constexpr long long kSmallAdjWords = 3ll << 20;   // 12 MB of adjacency bitmap: one segment of ~10 k points (measured on a B200:
                                                  // 0.21 ms at 2.3 k points, 0.36 ms at 6 k vs 0.50 / 0.63 ms through the cell grid; the
                                                  // O(n^2) phases cross over near 12 k points)
constexpr int kSmallCentreHead = pbsm::kCentreHead;             // clusters whose centres travel with the first read-back
constexpr int kSmallHead = 48;                    // ints of scalars in front of the result block (16 scalars + 11 time stamps)
}  // namespace

static int run_small(pb_ctx *ctx, const float *x, const float *y, const float *z, const float *xo, const float *yo,
                     const float *zo, const int32_t *sem, const std::vector<int> &start, int S, int n, const float *radius,
                     const int32_t *min_pts, float para_f, int assign_lp, int32_t *cluster_id, int32_t *cluster_num,
                     int32_t *degree, float *center, int64_t center_cap, int32_t *clt_sem, int64_t clt_sem_cap,
                     int64_t *n_clusters_out, int64_t *call_clusters, bool host_io, cudaStream_t st) {
    if (ctx->small_mode == 0 || S < 1 || S > pbsm::kMaxSeg || n < 1) return kNotSmall;
    pbsm::SmallArgs a;
    std::memset(&a, 0, sizeof(a));
    long long adj_words = 0;
    int mask_words = 0;
    for (int s = 0; s < S; s++) {
        const int ns = start[s + 1] - start[s], W = (ns + 31) / 32;
        if (W > 32 * pbsm::kWPL) return kNotSmall;
        a.start[s] = start[s], a.mask_off[s] = mask_words, a.adj_off[s] = adj_words;
        adj_words += (long long)ns * W;
        mask_words += W;
    }
    a.start[S] = n, a.mask_off[S] = mask_words, a.adj_off[S] = adj_words;
    if (adj_words > kSmallAdjWords) return kNotSmall;
    if (ctx->coop_blocks_per_sm == 0) {
        int coop = 0, occ = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pbsm::k_small, pbsm::kThreads, 0);
        ctx->sm_count = sms;
        ctx->coop_blocks_per_sm = (coop && occ > 0 && sms > 0) ? std::min(occ, 2) : -1;
    }
    if (ctx->coop_blocks_per_sm < 0) return kNotSmall;
    const size_t N = (size_t)n;
    // ---- workspace
    float *din = nullptr, *d_center = nullptr;
    int *d_cltsem = nullptr, *outblk = nullptr, *d_start = nullptr, *d_callfirst = nullptr;
    char *zero_begin = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &ar = pass == 0 ? dry : ctx->arena;
        if (pass == 1) ctx->arena.off = 0, ctx->arena.dry = false;
        if (host_io) din = ar.get<float>(7 * N);  // x y z xo yo zo sem, one H2D copy
        d_center = ar.get<float>(3 * N), d_cltsem = ar.get<int>(N);
        d_start = ar.get<int>(S + 1), d_callfirst = ar.get<int>(S);
        a.sg.cls = ar.get<int>(S), a.sg.min_pts = ar.get<int>(S), a.sg.r2 = ar.get<float>(S), a.sg.k_base = ar.get<int>(S);
        a.sg.cluster_num = ar.get<int>(S), a.sg.id_base = ar.get<int>(S);
        a.seg_of = ar.get<int>(N), a.parent = ar.get<int>(N), a.gid_at = ar.get<int>(N), a.raw_label = ar.get<int>(N);
        a.rep = ar.get<int>(N), a.keep = ar.get<int>(N), a.kscan = ar.get<int>(N), a.clt_seg = ar.get<int>(N);
        a.adj = ar.get<unsigned>((size_t)std::max<long long>(adj_words, 1));
        a.root0 = ar.get<int>(N), a.best64 = ar.get<unsigned long long>(N);
        a.tslot = ar.get<int>(N), a.troot = ar.get<int>(pbsm::kTreeMax), a.lplist = ar.get<int>(N);
        a.tidx = ar.get<unsigned char>(N);
        // zero-initialised region, ending with the scalars; the result block [scalars | cluster_num | cluster_id | degree]
        // starts there (one read-back when the results go to the host)
        zero_begin = reinterpret_cast<char *>(ar.get<char>(0));
        a.hpmask = ar.get<unsigned>(mask_words), a.labmask = ar.get<unsigned>(mask_words);
        a.flag = ar.get<int>(N), a.raw_count = ar.get<int>(N);
        a.deg_acc = ar.get<int>(N);
        a.conn = ar.get<unsigned long long>(pbsm::kTreeMax), a.tcnt = ar.get<int>(4), a.lpcnt = ar.get<int>(pbsm::kMaxSeg);
        // [scalars | cluster_num | cluster_id | degree | centres and classes of the first kSmallCentreHead clusters]
        outblk = ar.get<int>(kSmallHead + pbsm::kMaxSeg + 2 * N + 4 * kSmallCentreHead);
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
        }
    }
    a.scal = outblk;
    const size_t zero_bytes = reinterpret_cast<char *>(outblk + kSmallHead) - zero_begin;
    a.n = n, a.S = S, a.assign_lp = assign_lp;
    a.tree_cap = ctx->small_tree_cap, a.deg_slices = ctx->small_deg_slices;
    for (int i = 0; i < 18; i++) a.radius[i] = radius[i], a.min_pts[i] = min_pts[i], a.thresh[i] = kMeanCount[i] * para_f;
    a.sg.start = d_start, a.call_first = d_callfirst;
    // ---- inputs
    const size_t in_bytes = 7 * N * sizeof(float);
    const size_t out_ints = kSmallHead + pbsm::kMaxSeg + 2 * N;
    const size_t head_bytes = (size_t)kSmallCentreHead * 4 * sizeof(float);
    if (host_io) {
        const size_t need = std::max(in_bytes, out_ints * sizeof(int) + head_bytes);
        if (ctx->h_stage_cap < need) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr, ctx->h_stage_cap = 0;
            PB_CUDA(cudaMallocHost(&ctx->h_stage, need + need / 2));
            ctx->h_stage_cap = need + need / 2;
        }
        float *hs = reinterpret_cast<float *>(ctx->h_stage);
        const float *src[6] = {x, y, z, xo, yo, zo};
        for (int k = 0; k < 6; k++) std::memcpy(hs + k * N, src[k], N * sizeof(float));
        std::memcpy(hs + 6 * N, sem, N * sizeof(int));
        PB_CUDA(cudaMemcpyAsync(din, hs, in_bytes, cudaMemcpyHostToDevice, st));
        a.x = din, a.y = din + N, a.z = din + 2 * N, a.xo = din + 3 * N, a.yo = din + 4 * N, a.zo = din + 5 * N;
        a.sem = reinterpret_cast<const int *>(din + 6 * N);
        a.cluster_num = outblk + kSmallHead, a.cluster_id = outblk + kSmallHead + pbsm::kMaxSeg, a.degree = outblk + kSmallHead + pbsm::kMaxSeg + n;
        a.center = d_center, a.clt_sem = d_cltsem;
        a.center_head = reinterpret_cast<float *>(outblk + out_ints), a.cltsem_head = outblk + out_ints + 3 * kSmallCentreHead;
    } else {
        a.x = x, a.y = y, a.z = z, a.xo = xo, a.yo = yo, a.zo = zo, a.sem = sem;
        a.cluster_num = cluster_num, a.cluster_id = cluster_id, a.degree = degree;
        a.center = d_center, a.clt_sem = d_cltsem;   // copied to the caller's arrays once the cluster count is known
    }
    PB_CUDA(cudaMemsetAsync(zero_begin, 0, zero_bytes, st));
    const int tasks = (n + pbsm::kQB - 1) / pbsm::kQB + S;
    const int max_blocks = ctx->sm_count * ctx->coop_blocks_per_sm;
    const int blocks = std::max(std::min(ctx->sm_count, max_blocks), std::min(max_blocks, div_up(tasks, pbsm::kThreads / 32)));
    void *kargs[] = {&a};
    PB_CUDA(cudaLaunchCooperativeKernel((void *)pbsm::k_small, dim3(blocks), dim3(pbsm::kThreads), kargs, 0, st));
    ctx->launches = 1;
    // ---- results
    int *hres = reinterpret_cast<int *>(ctx->h_stage);
    int scal_local[kSmallHead];
    const int head = std::min(n, kSmallCentreHead);
    if (host_io) {
        // one read-back: the kernel keeps the first clusters' centres and classes right behind the result block
        PB_CUDA(cudaMemcpyAsync(hres, outblk, (out_ints + 4 * (size_t)kSmallCentreHead) * sizeof(int), cudaMemcpyDeviceToHost, st));
    } else {
        PB_CUDA(cudaMemcpyAsync(ctx->h_small, outblk, sizeof(int) * kSmallHead, cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    const int *sc = host_io ? hres : ctx->h_small;
    std::memcpy(scal_local, sc, sizeof(scal_local));
    const int errbits = scal_local[0];
    if (errbits & pb::kErrSem) return fail(ctx, PB_ERR_SEM_RANGE, "class id outside [2,19]");
    if (errbits & pb::kErrNonFinite) return fail(ctx, PB_ERR_NONFINITE, "non-finite coordinate");
    if (errbits & pb::kErrMixed) return kNotSmall;   // segments that mix classes take the general path
    const long long K = scal_local[3];
    *n_clusters_out = K;
    if (host_io) {
        std::memcpy(cluster_num, hres + kSmallHead, sizeof(int) * S);
        std::memcpy(cluster_id, hres + kSmallHead + pbsm::kMaxSeg, sizeof(int) * N);
        std::memcpy(degree, hres + kSmallHead + pbsm::kMaxSeg + N, sizeof(int) * N);
    }
    if (K > 0) {
        if (!center || !clt_sem || 3LL * K > center_cap || K > clt_sem_cap)
            return fail(ctx, PB_ERR_CAPACITY, "center / clt_sem capacity too small for " + std::to_string(K) + " clusters");
        if (host_io && K <= head) {
            std::memcpy(center, hres + out_ints, sizeof(float) * 3 * (size_t)K);
            std::memcpy(clt_sem, hres + out_ints + 3 * kSmallCentreHead, sizeof(int) * (size_t)K);
        } else {
            cudaMemcpyKind kind = host_io ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
            PB_CUDA(cudaMemcpyAsync(center, d_center, sizeof(float) * 3 * (size_t)K, kind, st));
            PB_CUDA(cudaMemcpyAsync(clt_sem, d_cltsem, sizeof(int) * (size_t)K, kind, st));
            PB_CUDA(cudaStreamSynchronize(st));
        }
    }
    if (call_clusters) call_clusters[0] = K;
    for (int i = 0; i < ST_COUNT; i++) ctx->stage_ms[i] = 0.f;
    {   // phase times of the kernel (globaltimer stamps of the grid's first thread at every barrier)
        unsigned long long ts[11];
        std::memcpy(ts, scal_local + 16, sizeof(ts));
        auto ms = [&](int i0, int i1) { return ts[i1] >= ts[i0] ? (float)((double)(ts[i1] - ts[i0]) * 1e-6) : 0.f; };
        ctx->stage_ms[ST_DEGREE] = ms(0, 1), ctx->stage_ms[ST_HP] = ms(1, 2), ctx->stage_ms[ST_UNION] = ms(2, 3);
        ctx->stage_ms[ST_COMPONENTS] = ms(3, 5), ctx->stage_ms[ST_LABEL] = ms(5, 6), ctx->stage_ms[ST_FILTER] = ms(6, 7);
        ctx->stage_ms[ST_LP_BUILD] = ms(7, 8), ctx->stage_ms[ST_LP_NN] = ms(8, 9), ctx->stage_ms[ST_CENTRES] = ms(9, 10);
    }
    for (int i = 0; i < 9; i++) ctx->counters[i] = 0;
    ctx->counters[5] = scal_local[2], ctx->counters[6] = 1;
    ctx->counters[9] = 1;   // small-call path taken
    return PB_OK;
}

static int run_impl(pb_ctx *ctx, const float *x, const float *y, const float *z, const float *xo, const float *yo,
                    const float *zo, const int32_t *sem, const int32_t *seg_counts, int32_t n_seg,
                    const int32_t *call_seg_counts, int32_t n_calls, int64_t n_pts, const float *radius,
                    const int32_t *min_pts, float para_f, int assign_lp, int32_t *cluster_id, int32_t *cluster_num,
                    int32_t *degree, float *center, int64_t center_cap, int32_t *clt_sem, int64_t clt_sem_cap,
                    int64_t *n_clusters_out, int64_t *call_clusters, int mem_kind, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n_seg < 0 || n_pts < 0 || n_calls < 0 || (n_seg > 0 && !seg_counts) || !radius || !min_pts || !n_clusters_out)
        return fail(ctx, PB_ERR_ARG, "null / negative argument");
    if (n_pts >= (int64_t)1 << 30) return fail(ctx, PB_ERR_ARG, "n_pts must be below 2^30");
    if (n_seg >= (1 << 22)) return fail(ctx, PB_ERR_ARG, "n_seg must be below 2^22");
    if (mem_kind != PB_MEM_HOST && mem_kind != PB_MEM_DEVICE) return fail(ctx, PB_ERR_ARG, "bad mem_kind");
    const int S = n_seg;
    const int n = (int)n_pts;
    std::vector<int> start(S + 1, 0);
    for (int s = 0; s < S; s++) {
        if (seg_counts[s] < 0) return fail(ctx, PB_ERR_ARG, "negative segment size");
        long long t = (long long)start[s] + seg_counts[s];
        if (t > n_pts) return fail(ctx, PB_ERR_ARG, "sum(seg_counts) != n_pts");
        start[s + 1] = (int)t;
    }
    if ((S == 0 ? 0 : start[S]) != n) return fail(ctx, PB_ERR_ARG, "sum(seg_counts) != n_pts");
    std::vector<int> call_seg0(n_calls + 1, 0);
    for (int c = 0; c < n_calls; c++) {
        if (!call_seg_counts || call_seg_counts[c] < 0 || call_seg0[c] + call_seg_counts[c] > S)
            return fail(ctx, PB_ERR_ARG, "bad call_seg_counts");
        call_seg0[c + 1] = call_seg0[c] + call_seg_counts[c];
    }
    if (call_seg0[n_calls] != S) return fail(ctx, PB_ERR_ARG, "sum(call_seg_counts) != n_seg");
    *n_clusters_out = 0;
    if (call_clusters)
        for (int c = 0; c < n_calls; c++) call_clusters[c] = 0;
    if (n > 0 && (!x || !y || !z || !xo || !yo || !zo || !sem || !cluster_id || !degree))
        return fail(ctx, PB_ERR_ARG, "null data pointer");
    if (S > 0 && !cluster_num) return fail(ctx, PB_ERR_ARG, "null cluster_num");
    const bool host_io = mem_kind == PB_MEM_HOST;
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, !host_io);
    if (n == 0) {
        if (S > 0) {
            if (host_io) std::memset(cluster_num, 0, sizeof(int) * S);
            else PB_CUDA(cudaMemsetAsync(cluster_num, 0, sizeof(int) * S, st));
            PB_CUDA(cudaStreamSynchronize(st));
        }
        return PB_OK;
    }
    if (n_calls == 1) {
        int rc = run_small(ctx, x, y, z, xo, yo, zo, sem, start, S, n, radius, min_pts, para_f, assign_lp, cluster_id, cluster_num,
                           degree, center, center_cap, clt_sem, clt_sem_cap, n_clusters_out, call_clusters, host_io, st);
        if (rc != kNotSmall) return rc;
    }
    ctx->counters[9] = 0;
    static bool smem_attr_done = false;
    if (!smem_attr_done) {  // the 64-bit-key sort tile needs more than the default 48 KB of dynamic shared memory
        cudaFuncSetAttribute(pb::k_sort_pass<uint64_t, kTileItems>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<uint64_t, kTileItems>));
        cudaFuncSetAttribute(pb::k_sort_pass<uint32_t, kTileItems>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<uint32_t, kTileItems>));
        cudaFuncSetAttribute(pb::k_sort_pass<uint64_t, kTileItemsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<uint64_t, kTileItemsSmall>));
        cudaFuncSetAttribute(pb::k_sort_pass<uint32_t, kTileItemsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<uint32_t, kTileItemsSmall>));
        smem_attr_done = true;
    }

    // ---- chunking: runs of consecutive calls of ~chunk_points points, alternating over two streams so
    //      the latency-bound tail kernels and the host copies of one chunk overlap k_degree of the next
    struct Chunk { int c0, c1, s0, s1, p0, p1; };
    std::vector<Chunk> chunks;
    {
        // automatic: two chunks once a call has >= 6 M points (smaller chunks lose more to launch gaps and
        // kernel tails than the overlap wins: 48 ms at 2 chunks, 51 at 4, 59 at 10 for 28.8 M points)
        // host data: three chunks on the prioritised stream pair (the first H2D copy is the exposed head of the pipeline)
        // (round 2, measured on 3.6 M points = one rank's share of C1 at 8 GPUs: 6.36 ms unchunked, 6.18 ms at 2 chunks,
        // 7.1 / 7.4 ms at 3 / 4 chunks; from host memory 8.7 -> 7.9 ms at 2 chunks)
        const int auto_chunks = n >= 6000000 ? (host_io ? 3 : ctx->dev_chunks) : (n >= 1500000 ? 2 : 1);
        long long target = ctx->chunk_points > 0 ? ctx->chunk_points
                           : (ctx->chunk_points == 0 && auto_chunks > 1 ? ((long long)n + auto_chunks - 1) / auto_chunks : (long long)n + 1);
        int c = 0;
        while (c < n_calls) {
            Chunk ch;
            ch.c0 = c;
            ch.s0 = call_seg0[c];
            ch.p0 = start[ch.s0];
            long long pts = 0;
            // host data, automatic chunking: unequal chunks (PB_HOST_SPLIT, per mille of the call) — a small first chunk
            // shortens the exposed head of the pipeline (its H2D copy), a small last one the exposed read-back
            long long tgt = target;
            if (host_io && ctx->chunk_points == 0 && auto_chunks == 3 && ctx->host_split[0] > 0) {
                const int k = (int)chunks.size();
                tgt = k < 8 && ctx->host_split[k] > 0 ? (long long)n * ctx->host_split[k] / 1000 : (long long)n;
            }
            while (c < n_calls && pts < tgt) {  // whole calls until the chunk reaches the target size
                pts += start[call_seg0[c + 1]] - start[call_seg0[c]];
                c++;
            }
            ch.c1 = c;
            ch.s1 = call_seg0[c];
            ch.p1 = start[ch.s1];
            if (ch.p1 > ch.p0) chunks.push_back(ch);
            else if (ch.s1 > ch.s0 && !chunks.empty()) chunks.back().c1 = ch.c1, chunks.back().s1 = ch.s1;  // empty calls ride along
            else if (ch.s1 > ch.s0) chunks.push_back(ch);
        }
        // leading empty-call chunk followed by real ones: merge forward
        if (chunks.size() > 1 && chunks[0].p1 == chunks[0].p0) {
            chunks[1].c0 = chunks[0].c0, chunks[1].s0 = chunks[0].s0, chunks[1].p0 = chunks[0].p0;
            chunks.erase(chunks.begin());
        }
    }
    const int G = (int)chunks.size();
    const bool multi = G > 1;
    const int slots = multi ? std::min(G, host_io ? ctx->slots_host : ctx->slots_dev) : 1;
    const bool small_tiles = ctx->tile_mode == 1 || (ctx->tile_mode < 0 && n < 262144);
    // per-chunk host tables: local segment starts, first segment of every segment's call, tile table
    std::vector<std::vector<int>> lstart(G), lcall(G);
    std::vector<TileHost> tiles(G);
    int n_max = 0, S_max = 0, T_max = 0, rows_max = 0, hslots_max = 0;
    for (int gi = 0; gi < G; gi++) {
        const Chunk &ch = chunks[gi];
        const int cn = ch.p1 - ch.p0, cS = ch.s1 - ch.s0;
        lstart[gi].assign(cS + 1, 0);
        lcall[gi].assign(std::max(cS, 1), 0);
        for (int s = 0; s <= cS; s++) lstart[gi][s] = start[ch.s0 + s] - ch.p0;
        for (int c = ch.c0; c < ch.c1; c++)
            for (int s = call_seg0[c]; s < call_seg0[c + 1]; s++) lcall[gi][s - ch.s0] = call_seg0[c] - ch.s0;
        build_tiles(lstart[gi].data(), cS, small_tiles ? pb::kTB * kTileItemsSmall : kTile, tiles[gi]);
        n_max = std::max(n_max, cn), S_max = std::max(S_max, cS), T_max = std::max(T_max, tiles[gi].T);
        rows_max = std::max(rows_max, tiles[gi].rows), hslots_max = std::max(hslots_max, tiles[gi].slots);
    }

    // Segments that mix classes (never produced by PBNet) are detected on the device by the first attempt;
    // the call is then repeated with the per-(cell, class) tables of the mixed-class kernels.
    for (int attempt = 0; attempt < 2; attempt++) {
    const bool mixed = attempt == 1;
    ctx->launches = 0;
    // ---- workspace: `slots` private work areas + call-wide cluster metadata ------------------------------
    Work w[pb_ctx::kMaxSlots];
    float *center_all = nullptr;
    int *clt_sem_all = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        if (pass == 1) ctx->arena.off = 0, ctx->arena.dry = false;
        center_all = a.get<float>(3 * (size_t)n);
        clt_sem_all = a.get<int>((size_t)n);
        for (int k = 0; k < slots; k++) {
            std::memset(&w[k], 0, sizeof(Work));
            plan(a, w[k], n_max, S_max, T_max, rows_max, hslots_max, host_io, mixed);
        }
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
        }
    }
    // pinned scalars per chunk
    if ((int)ctx->h_chunk_cap < G) {
        if (ctx->h_chunk_scalars) cudaFreeHost(ctx->h_chunk_scalars);
        if (ctx->h_chunk_counters) cudaFreeHost(ctx->h_chunk_counters);
        ctx->h_chunk_scalars = nullptr, ctx->h_chunk_counters = nullptr, ctx->h_chunk_cap = 0;
        PB_CUDA(cudaMallocHost(&ctx->h_chunk_scalars, sizeof(int) * 16 * (size_t)G));
        PB_CUDA(cudaMallocHost(&ctx->h_chunk_counters, sizeof(unsigned long long) * 8 * (size_t)G));
        ctx->h_chunk_cap = G;
    }
    const bool prof = ctx->profiling;
    if (prof && (int)ctx->chunk_ev.size() < G * (ST_COUNT + 1)) {
        size_t old = ctx->chunk_ev.size();
        ctx->chunk_ev.resize((size_t)G * (ST_COUNT + 1));
        for (size_t i = old; i < ctx->chunk_ev.size(); i++) cudaEventCreate(&ctx->chunk_ev[i]);
    }
    if ((int)ctx->front_ev.size() < G) {
        size_t old = ctx->front_ev.size();
        ctx->front_ev.resize(G);
        ctx->deg_ev.resize((size_t)2 * G);
        for (size_t i = old; i < (size_t)G; i++) {
            cudaEventCreateWithFlags(&ctx->front_ev[i], cudaEventDisableTiming);
            cudaEventCreate(&ctx->deg_ev[2 * i]);
            cudaEventCreate(&ctx->deg_ev[2 * i + 1]);
        }
    }
    float thresh[18];
    for (int i = 0; i < 18; i++) thresh[i] = kMeanCount[i] * para_f;  // fp32 multiply, binary.cu:256

    cudaStream_t cs[pb_ctx::kMaxSlots] = {st, st, st, st};
    if (multi) {
        const bool prio = ctx->prio_mode < 0 ? host_io : ctx->prio_mode == 1;
        PB_CUDA(cudaEventRecord(ctx->ev_fork, st));
        for (int k = 0; k < slots; k++) {
            cs[k] = prio ? ctx->auxp[k] : ctx->aux[k];
            PB_CUDA(cudaStreamWaitEvent(cs[k], ctx->ev_fork, 0));
        }
    }
    std::vector<ChunkIO> ios(G);
    std::vector<int> hdr_host;
    auto make_io = [&](int gi) {
        const Chunk &ch = chunks[gi];
        ChunkIO &io = ios[gi];
        io.n = ch.p1 - ch.p0, io.S = ch.s1 - ch.s0;
        io.x = x + ch.p0, io.y = y + ch.p0, io.z = z + ch.p0, io.xo = xo + ch.p0, io.yo = yo + ch.p0, io.zo = zo + ch.p0;
        io.sem = sem + ch.p0;
        io.cluster_id = cluster_id + ch.p0, io.degree = degree + ch.p0, io.cluster_num = cluster_num + ch.s0;
        io.start = lstart[gi].data(), io.call_first = lcall[gi].data(), io.tiles = &tiles[gi];
        io.center_out = center_all + 3 * (size_t)ch.p0, io.clt_sem_out = clt_sem_all + ch.p0;
        io.h_scalars = ctx->h_chunk_scalars + 16 * gi, io.h_counters = ctx->h_chunk_counters + 8 * gi;
        io.ev = prof ? &ctx->chunk_ev[(size_t)gi * (ST_COUNT + 1)] : nullptr;
        io.ev_front = ctx->front_ev[gi];
        io.ev_deg[0] = ctx->deg_ev[2 * gi], io.ev_deg[1] = ctx->deg_ev[2 * gi + 1];
        io.wait_deg = (ctx->stagger && gi > 0) ? ctx->deg_ev[2 * (gi - 1) + 1] : nullptr;
        io.slot = gi % slots;
    };
    // Fronts run ahead by at most one chunk per stream slot: chunk gi+2 reuses the work area of chunk gi, so its front is
    // enqueued behind the rest of chunk gi (same stream).  The host waits for a front only to read three integers (the cell
    // extents that size the sort keys); the other stream keeps the GPU busy meanwhile.
    int fronts = 0;
    auto front = [&](int gi) -> int {
        make_io(gi);
        return small_tiles ? enqueue_front<kTileItemsSmall>(ctx, w[gi % slots], ios[gi], host_io, mixed, radius, min_pts, thresh, cs[gi % slots], ctx->launches, hdr_host)
                           : enqueue_front<kTileItems>(ctx, w[gi % slots], ios[gi], host_io, mixed, radius, min_pts, thresh, cs[gi % slots], ctx->launches, hdr_host);
    };
    for (; fronts < std::min(G, slots); fronts++) {
        int rc = front(fronts);
        if (rc) return rc;
    }
    for (int gi = 0; gi < G; gi++) {
        PB_CUDA(cudaEventSynchronize(ios[gi].ev_front));
        int rc;
        if (small_tiles)
            rc = mixed ? enqueue_rest<true, kTileItemsSmall>(ctx, w[gi % slots], ios[gi], host_io, assign_lp, cs[gi % slots], ctx->launches)
                       : enqueue_rest<false, kTileItemsSmall>(ctx, w[gi % slots], ios[gi], host_io, assign_lp, cs[gi % slots], ctx->launches);
        else
            rc = mixed ? enqueue_rest<true, kTileItems>(ctx, w[gi % slots], ios[gi], host_io, assign_lp, cs[gi % slots], ctx->launches)
                       : enqueue_rest<false, kTileItems>(ctx, w[gi % slots], ios[gi], host_io, assign_lp, cs[gi % slots], ctx->launches);
        if (rc) return rc;
        if (fronts < G) {  // the next chunk of this slot
            rc = front(fronts++);
            if (rc) return rc;
        }
    }
    if (multi) {
        for (int k = 0; k < slots; k++) {
            PB_CUDA(cudaEventRecord(ctx->ev_join[k], cs[k]));
            PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[k], 0));
        }
    }
    PB_CUDA(cudaStreamSynchronize(st));
    int errbits = 0;
    long long K = 0;
    for (int gi = 0; gi < G; gi++) errbits |= ctx->h_chunk_scalars[16 * gi], K += ctx->h_chunk_scalars[16 * gi + 3];
    if (errbits & pb::kErrSem) return fail(ctx, PB_ERR_SEM_RANGE, "class id outside [2,19]");
    if (errbits & pb::kErrNonFinite) return fail(ctx, PB_ERR_NONFINITE, "non-finite coordinate");
    if (errbits & pb::kErrRange) return fail(ctx, PB_ERR_RANGE, "segment spans more than 16383 grid cells along an axis");
    if (errbits & pb::kErrRadius)
        return fail(ctx, PB_ERR_MIXED_CLASS, "a segment mixes classes with different radii (undefined in the reference)");
    if ((errbits & pb::kErrMixed) && !mixed) continue;  // repeat with the mixed-class kernels
    *n_clusters_out = K;
    if (K > 0) {
        if (!center || !clt_sem || 3LL * K > center_cap || K > clt_sem_cap)
            return fail(ctx, PB_ERR_CAPACITY, "center / clt_sem capacity too small for " + std::to_string(K) + " clusters");
        cudaMemcpyKind kind = host_io ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        long long ko = 0;
        for (int gi = 0; gi < G; gi++) {
            long long kg = ctx->h_chunk_scalars[16 * gi + 3];
            if (kg > 0) {
                PB_CUDA(cudaMemcpyAsync(center + 3 * ko, center_all + 3 * (size_t)chunks[gi].p0, sizeof(float) * 3 * (size_t)kg, kind, st));
                PB_CUDA(cudaMemcpyAsync(clt_sem + ko, clt_sem_all + chunks[gi].p0, sizeof(int) * (size_t)kg, kind, st));
            }
            ko += kg;
        }
        PB_CUDA(cudaStreamSynchronize(st));
    }
    if (call_clusters) {
        std::vector<int> cn(S);
        const int *src = cluster_num;
        if (!host_io) {
            PB_CUDA(cudaMemcpy(cn.data(), cluster_num, sizeof(int) * S, cudaMemcpyDeviceToHost));
            src = cn.data();
        }
        for (int c = 0; c < n_calls; c++) {
            long long t = 0;
            for (int s = call_seg0[c]; s < call_seg0[c + 1]; s++) t += src[s];
            call_clusters[c] = t;
        }
    }
    {   // always-on: time of k_degree (one event pair per chunk); with profiling every stage
        for (int i = 0; i < ST_COUNT; i++) ctx->stage_ms[i] = 0.f;
        // chunks on the two streams run concurrently: report the time during which k_degree was resident at all (the
        // union of the per-chunk intervals on the common event clock), not the sum of overlapping intervals
        std::vector<std::pair<float, float>> iv;
        for (int gi = 0; gi < G; gi++) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, ios[0].ev_deg[0], ios[gi].ev_deg[0]);
            cudaEventElapsedTime(&b, ios[0].ev_deg[0], ios[gi].ev_deg[1]);
            iv.push_back({a, b});
        }
        std::sort(iv.begin(), iv.end());
        float covered = 0.f, cur_a = iv[0].first, cur_b = iv[0].second;
        for (size_t i = 1; i < iv.size(); i++) {
            if (iv[i].first > cur_b) covered += cur_b - cur_a, cur_a = iv[i].first, cur_b = iv[i].second;
            else cur_b = std::max(cur_b, iv[i].second);
        }
        covered += cur_b - cur_a;
        ctx->stage_ms[ST_DEGREE] = covered;
        ctx->counters[6] = G;
    }
    if (prof && getenv("PB_TIMELINE")) {  // diagnostic: when every stage of every chunk started, on the clock of chunk 0's first event
        cudaEvent_t t0 = ctx->chunk_ev[0];
        for (int gi = 0; gi < G; gi++) {
            cudaEvent_t *ev = &ctx->chunk_ev[(size_t)gi * (ST_COUNT + 1)];
            fprintf(stderr, "[pb timeline] chunk %d (%d points, stream %d):", gi, chunks[gi].p1 - chunks[gi].p0, gi % slots);
            for (int i = 0; i <= ST_COUNT; i++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, t0, ev[i]);
                fprintf(stderr, " %s %.2f", i < ST_COUNT ? kStageNames[i] : "end", ms);
            }
            fprintf(stderr, "\n");
        }
    }
    if (prof) {
        const float deg_union = ctx->stage_ms[ST_DEGREE];
        for (int i = 0; i < ST_COUNT; i++) ctx->stage_ms[i] = 0.f;
        for (int i = 0; i < 6; i++) ctx->counters[i] = 0;
        ctx->counters[8] = 0;
        ctx->counters[10] = ctx->counters[11] = ctx->counters[12] = ctx->counters[13] = ctx->counters[14] = 0;
        for (int gi = 0; gi < G; gi++) {
            cudaEvent_t *ev = &ctx->chunk_ev[(size_t)gi * (ST_COUNT + 1)];
            for (int i = 0; i < ST_COUNT; i++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                ctx->stage_ms[i] += ms;
            }
            const unsigned long long *hc = ctx->h_chunk_counters + 8 * gi;
            const int *hs = ctx->h_chunk_scalars + 16 * gi;
            ctx->counters[0] += (int64_t)hc[0];
            ctx->counters[1] += (int64_t)hc[1];
            ctx->counters[2] += (int64_t)hc[2];
            ctx->counters[10] += (int64_t)hc[3];
            ctx->counters[11] += (int64_t)hc[4];
            ctx->counters[12] += (int64_t)hc[5];
            ctx->counters[13] += (int64_t)hc[6];
            ctx->counters[14] += (int64_t)hc[7];
            ctx->counters[3] += hs[4];
            ctx->counters[4] += hs[1];
            ctx->counters[5] += hs[2];
            ctx->counters[8] += hs[6];
        }
        ctx->stage_ms[ST_DEGREE] = deg_union;
    }
    ctx->counters[7] = mixed;
    return PB_OK;
    }  // attempt
    return fail(ctx, PB_ERR_MIXED_CLASS, "unreachable");
}

extern "C" int pb_binary_cluster_batched(pb_ctx *ctx, const float *x, const float *y, const float *z, const float *xo,
                                         const float *yo, const float *zo, const int32_t *sem, const int32_t *seg_counts,
                                         int32_t n_seg, const int32_t *call_seg_counts, int32_t n_calls, int64_t n_pts,
                                         const float *radius, const int32_t *min_pts, float para_f, int assign_lp,
                                         int32_t *cluster_id, int32_t *cluster_num, int32_t *degree, float *center,
                                         int64_t center_cap, int32_t *clt_sem, int64_t clt_sem_cap,
                                         int64_t *n_clusters_out, int64_t *call_clusters, int mem_kind, void *stream) {
    return run_impl(ctx, x, y, z, xo, yo, zo, sem, seg_counts, n_seg, call_seg_counts, n_calls, n_pts, radius, min_pts,
                    para_f, assign_lp, cluster_id, cluster_num, degree, center, center_cap, clt_sem, clt_sem_cap,
                    n_clusters_out, call_clusters, mem_kind, stream);
}

extern "C" int pb_binary_cluster(pb_ctx *ctx, const float *x, const float *y, const float *z, const float *xo,
                                 const float *yo, const float *zo, const int32_t *sem, const int32_t *seg_counts,
                                 int32_t n_seg, int64_t n_pts, const float *radius, const int32_t *min_pts, float para_f,
                                 int assign_lp, int32_t *cluster_id, int32_t *cluster_num, int32_t *degree, float *center,
                                 int64_t center_cap, int32_t *clt_sem, int64_t clt_sem_cap, int64_t *n_clusters_out,
                                 int mem_kind, void *stream) {
    int32_t one = n_seg;
    return run_impl(ctx, x, y, z, xo, yo, zo, sem, seg_counts, n_seg, &one, 1, n_pts, radius, min_pts, para_f, assign_lp,
                    cluster_id, cluster_num, degree, center, center_cap, clt_sem, clt_sem_cap, n_clusters_out, nullptr,
                    mem_kind, stream);
}

// =================================================================================================
// device-side front end of the fused class loop (network/PBNet.py:151-179, 282-294) — see pb_front.cuh.
// Device pointers only; the small class / segment tables come back to the host (one synchronisation).
// =================================================================================================
#include "pb_front.cuh"

extern "C" int pb_group_front(pb_ctx *ctx, const float *xyz, const float *offset, const int64_t *sem, const void *batch,
                              int batch_is64, int64_t n_pts, int32_t copies, const float *skip_thresh20, float *x, float *y,
                              float *z, float *xo, float *yo, float *zo, int32_t *sem32, int64_t *point_index,
                              int32_t *class_keep20, int32_t *seg_counts, int64_t *n_kept_out, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n_pts < 0 || copies < 1 || !skip_thresh20 || !class_keep20 || !seg_counts || !n_kept_out)
        return fail(ctx, PB_ERR_ARG, "null / negative argument");
    if ((int64_t)pbf::kMaxSem * copies >= pbf::kDropKey) return fail(ctx, PB_ERR_ARG, "more than 25 scene copies per call");
    if (n_pts >= ((int64_t)1 << 30)) return fail(ctx, PB_ERR_ARG, "n_pts must be below 2^30");
    *n_kept_out = 0;
    for (int c = 0; c < 20; c++) class_keep20[c] = 0;
    if (n_pts == 0) return PB_OK;
    if (!xyz || !offset || !sem || !batch || !x || !y || !z || !xo || !yo || !zo || !sem32 || !point_index)
        return fail(ctx, PB_ERR_ARG, "null data pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    const int n = (int)n_pts;
    const int start2[2] = {0, n};
    TileHost th;
    build_tiles(start2, 1, kTile, th);
    const int T = th.T;
    uint32_t *key = nullptr, *key_out = nullptr, *order = nullptr;
    unsigned *hist = nullptr, *state = nullptr;
    int *tickets = nullptr, *tabs = nullptr, *tile_dev = nullptr, *seg_start = nullptr, *d_err = nullptr;
    float *d_thr = nullptr;
    char *zero_begin = nullptr, *zero_end = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        key = a.get<uint32_t>(n), key_out = a.get<uint32_t>(n), order = a.get<uint32_t>(n);
        tile_dev = a.get<int>((size_t)6 * T + 2 + 20);
        tabs = a.get<int>(pbf::kTabSize);
        zero_begin = reinterpret_cast<char *>(a.get<char>(0));
        hist = a.get<unsigned>(pb::kBins), state = a.get<unsigned>((size_t)std::max(th.rows, 1) * pb::kBins);
        tickets = a.get<int>(4), d_err = a.get<int>(4);
        zero_end = reinterpret_cast<char *>(a.get<char>(0));
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
        }
    }
    std::vector<int> hdr((size_t)6 * T + 2 + 20);
    std::memcpy(hdr.data(), th.buf.data(), sizeof(int) * (size_t)6 * T);
    hdr[6 * T] = 0, hdr[6 * T + 1] = n;
    std::memcpy(hdr.data() + 6 * T + 2, skip_thresh20, 20 * sizeof(float));
    PB_CUDA(cudaMemcpyAsync(tile_dev, hdr.data(), sizeof(int) * hdr.size(), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemsetAsync(zero_begin, 0, zero_end - zero_begin, st));
    seg_start = tile_dev + 6 * T;
    d_thr = reinterpret_cast<float *>(tile_dev + 6 * T + 2);
    pb::TileTab tt;
    tt.begin = tile_dev, tt.count = tile_dev + T, tt.seg = tile_dev + 2 * T, tt.first = tile_dev + 3 * T, tt.hslot = tile_dev + 4 * T;
    tt.srow = tile_dev + 5 * T, tt.T = T;
    pbf::k_front_keys<<<std::min(div_up(n, 256), 148 * 8), 256, 0, st>>>(n_pts, (const long long *)sem, batch, batch_is64, copies, key, hist, d_err);
    pb::SortArgs<uint32_t> sa;
    sa.keys_in = key, sa.keys_out = key_out, sa.vals_in = nullptr, sa.vals_out = order, sa.hist = hist, sa.hist_stride = pb::kBins;
    sa.hist_off = 0, sa.state = state, sa.ticket = tickets, sa.shift = 0, sa.width = pb::kRadixBitsMax;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(pb::k_sort_pass<uint32_t, kTileItems>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(pb::SortSmem<uint32_t, kTileItems>));
        attr_done = true;
    }
    pb::k_sort_pass<uint32_t, kTileItems><<<dim3(T, 1), pb::kTB, sizeof(pb::SortSmem<uint32_t, kTileItems>), st>>>(sa, sa, tt, seg_start);
    pbf::k_front_tables<<<1, 32, 0, st>>>(hist, copies, d_thr, tabs);
    pbf::k_front_gather<<<std::min(div_up(n, 256), 148 * 8), 256, 0, st>>>(tabs, copies, order, xyz, offset, x, y, z, xo, yo, zo, sem32,
                                                                         (long long *)point_index);
    ctx->launches = 4;
    std::vector<int> h((size_t)pbf::kTabCnt + 512 + 1);
    PB_CUDA(cudaMemcpyAsync(h.data(), tabs, sizeof(int) * (pbf::kTabCnt + 512), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(h.data() + pbf::kTabCnt + 512, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    if (h[pbf::kTabCnt + 512] & pbf::kErrBatch) return fail(ctx, PB_ERR_ARG, "batch index outside [0, cluster_batch)");
    *n_kept_out = h[0];
    for (int c = 0; c < 20; c++) class_keep20[c] = h[2 + c];
    for (int i = 0; i < h[1] * copies; i++) seg_counts[i] = h[pbf::kTabCnt + i];
    return PB_OK;
}

// =================================================================================================
// voxelize / devoxelize — see pb_voxel.cuh
// =================================================================================================
#include "pb_voxel.cuh"


extern "C" int pb_voxelize(pb_ctx *ctx, const void *coords, int coord_f64, int stride, int has_batch_col,
                           const int32_t *batch, int64_t n, double voxel_size, int32_t *vcoords, int64_t *index,
                           int64_t *inverse, int32_t *order, int32_t *vox_start, int64_t cap, int64_t *n_voxels_out,
                           int mem_kind, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n < 0 || !n_voxels_out || (stride != 3 && stride != 4) || (has_batch_col && stride != 4) ||
        (mem_kind != PB_MEM_HOST && mem_kind != PB_MEM_DEVICE))
        return fail(ctx, PB_ERR_ARG, "bad argument");
    if (n >= ((int64_t)1 << 31)) return fail(ctx, PB_ERR_ARG, "n must be below 2^31");
    *n_voxels_out = 0;
    if (n == 0) return PB_OK;
    if (!coords || !vcoords || !index || !inverse || !order || !vox_start || cap < 1) return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, mem_kind == PB_MEM_DEVICE);
    const bool host_io = mem_kind == PB_MEM_HOST;
    const size_t N = (size_t)n, esz = coord_f64 ? 8 : 4;
    int64_t &L = ctx->launches;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        void *d_coords = host_io ? (void *)a.get<char>(N * stride * esz) : nullptr;
        int *d_batch = (host_io && batch) ? a.get<int>(N) : nullptr;
        int4 *d_vcoords = host_io ? a.get<int4>(N) : nullptr;
        long long *d_index = host_io ? a.get<long long>(N) : nullptr;
        long long *d_inverse = host_io ? a.get<long long>(N) : nullptr;
        int *d_order = host_io ? a.get<int>(N) : nullptr;
        int *d_vstart = host_io ? a.get<int>(N + 1) : nullptr;
        int4 *q = a.get<int4>(N);
        uint64_t *key = a.get<uint64_t>(N), *key_alt = a.get<uint64_t>(N);
        uint32_t *val = a.get<uint32_t>(N), *ord = a.get<uint32_t>(N);
        int *head = a.get<int>(N), *ex = a.get<int>(N);
        int *mnmx = a.get<int>(16);
        int *blocks = a.get<int>(N / pb::kScanTile + 2);
        void *sort_tmp = a.get<char>(RadixTmp::bytes(N, sizeof(uint64_t)));
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
            continue;
        }
        const void *c_in = coords;
        const int *b_in = batch;
        if (host_io) {
            PB_CUDA(cudaMemcpyAsync(d_coords, coords, N * stride * esz, cudaMemcpyHostToDevice, st));
            c_in = d_coords;
            if (batch) {
                PB_CUDA(cudaMemcpyAsync(d_batch, batch, N * 4, cudaMemcpyHostToDevice, st));
                b_in = d_batch;
            }
        } else {
            d_vcoords = reinterpret_cast<int4 *>(vcoords);
            d_index = reinterpret_cast<long long *>(index);
            d_inverse = reinterpret_cast<long long *>(inverse);
            d_order = order;
            d_vstart = vox_start;
        }
        if (!host_io && cap < n) return fail(ctx, PB_ERR_CAPACITY, "device outputs need capacity n (the voxel count is not known in advance)");
        const int h_init[16] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000,
                                (int)0x80000000, (int)0x80000000, 0, 0, 0, 0, 0, 0, 0, 0};
        PB_CUDA(cudaMemcpyAsync(mnmx, h_init, sizeof(h_init), cudaMemcpyHostToDevice, st));
        int *d_err = mnmx + 8, *d_V = mnmx + 9;
        const int T = 256, g = div_up(n, T);
        if (coord_f64)
            pbv::k_quantize<double><<<g, T, 0, st>>>((const double *)c_in, stride, has_batch_col, b_in, n, voxel_size, q, mnmx, mnmx + 4);
        else
            pbv::k_quantize<float><<<g, T, 0, st>>>((const float *)c_in, stride, has_batch_col, b_in, n, (float)voxel_size, q, mnmx, mnmx + 4);
        // the key holds only the occupied bits of (batch, x, y, z) relative to the per-call minimum: the host reads the
        // extents (one small synchronisation) to size the radix passes — 26 bits = 3 passes for a ScanNet scene at 2 cm
        PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, mnmx, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        int key_bits = 0;
        for (int k = 0; k < 4; k++) {
            long long ext = (long long)ctx->h_scalars[4 + k] - ctx->h_scalars[k];
            if (ext > 65535) return fail(ctx, PB_ERR_RANGE, "voxel grid spans more than 65535 cells along an axis");
            key_bits += pb::bit_width_i((int)ext);
        }
        pbv::k_vox_keys<<<g, T, 0, st>>>(q, n, mnmx, mnmx + 4, key, val, d_err);
        L += 2;
        {
            int rc = radix_sort<uint64_t>(ctx, sort_tmp, key, key_alt, val, ord, (int)n, key_bits, st, &L);
            if (rc) return rc;
        }
        pbv::k_vox_heads<<<g, T, 0, st>>>(key_alt, n, head);
        L++;
        launch_scan(st, head, (int)n, nullptr, ex, d_V, blocks, L);
        pbv::k_vox_table<<<g, T, 0, st>>>(q, ord, head, ex, n, d_vcoords, d_index, d_inverse, d_order, d_vstart, d_V);
        L++;
        PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, mnmx + 8, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaStreamSynchronize(st));
        if (ctx->h_scalars[0] & 4) return fail(ctx, PB_ERR_RANGE, "voxel grid spans more than 65535 cells along an axis");
        int64_t V = ctx->h_scalars[1];
        *n_voxels_out = V;
        if (host_io) {
            if (V > cap) return fail(ctx, PB_ERR_CAPACITY, "output capacity too small for " + std::to_string(V) + " voxels");
            PB_CUDA(cudaMemcpyAsync(vcoords, d_vcoords, sizeof(int4) * V, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(index, d_index, 8 * V, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(inverse, d_inverse, 8 * N, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(order, d_order, 4 * N, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(vox_start, d_vstart, 4 * (V + 1), cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
        }
    }
    return PB_OK;
}

extern "C" int pb_voxel_rows(pb_ctx *ctx, const float *rows, int64_t n_rows, int C, const int32_t *order,
                             const int32_t *vox_start, int64_t V, int mode, float *out, int mem_kind, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n_rows < 0 || V < 0 || C < 1 || mode < 0 || mode > 2 || (mem_kind != PB_MEM_HOST && mem_kind != PB_MEM_DEVICE))
        return fail(ctx, PB_ERR_ARG, "bad argument");
    if (V == 0) return PB_OK;
    if (!rows || !order || !vox_start || !out) return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, mem_kind == PB_MEM_DEVICE);
    const bool host_io = mem_kind == PB_MEM_HOST;
    const float *d_rows = rows;
    const int *d_order = order, *d_vs = vox_start;
    float *d_out = out;
    if (host_io) {
        size_t need = ((size_t)n_rows * C + (size_t)V * C) * 4 + (size_t)n_rows * 4 + (size_t)(V + 1) * 4 + 4096;
        int rc = ensure_arena(ctx, need, st);
        if (rc) return rc;
        float *r = ctx->arena.get<float>((size_t)n_rows * C);
        float *o = ctx->arena.get<float>((size_t)V * C);
        int *od = ctx->arena.get<int>((size_t)n_rows), *vs = ctx->arena.get<int>((size_t)V + 1);
        PB_CUDA(cudaMemcpyAsync(r, rows, (size_t)n_rows * C * 4, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(od, order, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(vs, vox_start, (size_t)(V + 1) * 4, cudaMemcpyHostToDevice, st));
        d_rows = r, d_out = o, d_order = od, d_vs = vs;
    }
    const int T = 256;
    const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_rows) | reinterpret_cast<uintptr_t>(d_out)) % 16 == 0);
    const unsigned width = (unsigned)(vec4 ? C / 4 : C);
    const unsigned long long total = (unsigned long long)V * width;
    const int g = (int)std::min<unsigned long long>((total + T - 1) / T, 148ULL * 64);
#define PB_ROWS(M)                                                                                                   \
    if (vec4) pbv::k_vox_rows<M, float4><<<g, T, 0, st>>>((const float4 *)d_rows, width, d_order, d_vs, total, (float4 *)d_out); \
    else pbv::k_vox_rows<M, float><<<g, T, 0, st>>>(d_rows, width, d_order, d_vs, total, d_out);
    if (mode == 0) { PB_ROWS(0) } else if (mode == 1) { PB_ROWS(1) } else { PB_ROWS(2) }
#undef PB_ROWS
    ctx->launches = 1;
    if (host_io) PB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)V * C * 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    if (host_io) PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_devoxelize(pb_ctx *ctx, const float *vfeat, int64_t V, int C, const int64_t *inverse, int64_t n,
                             float *out, int mem_kind, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n < 0 || V < 0 || C < 1 || (mem_kind != PB_MEM_HOST && mem_kind != PB_MEM_DEVICE)) return fail(ctx, PB_ERR_ARG, "bad argument");
    if (n == 0) return PB_OK;
    if (!vfeat || !inverse || !out) return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, mem_kind == PB_MEM_DEVICE);
    const bool host_io = mem_kind == PB_MEM_HOST;
    const float *d_v = vfeat;
    const long long *d_inv = reinterpret_cast<const long long *>(inverse);
    float *d_out = out;
    if (host_io) {
        size_t need = ((size_t)V * C + (size_t)n * C) * 4 + (size_t)n * 8 + 4096;
        int rc = ensure_arena(ctx, need, st);
        if (rc) return rc;
        float *v = ctx->arena.get<float>((size_t)V * C), *o = ctx->arena.get<float>((size_t)n * C);
        long long *iv = ctx->arena.get<long long>((size_t)n);
        PB_CUDA(cudaMemcpyAsync(v, vfeat, (size_t)V * C * 4, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(iv, inverse, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        d_v = v, d_out = o, d_inv = iv;
    }
    const int T = 256;
    bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_v) | reinterpret_cast<uintptr_t>(d_out)) % 16 == 0);
    const long long width = vec4 ? C / 4 : C;
    const long long total = n * width;
    const int U = ctx->devox_u;  // elements per thread (PB_DEVOX_U: 2, 4 or 8)
    const int grid = (int)std::min<long long>((total + (long long)U * T - 1) / ((long long)U * T), 148LL * 128 / U);  // grid-stride
#define PB_DEVOX(UU)                                                                                                          \
    if (total < (1LL << 32) - 4096) {                                                                                         \
        if (vec4) pbv::k_devox<float4, uint32_t, UU><<<grid, T, 0, st>>>((const float4 *)d_v, (uint32_t)width, d_inv, (uint32_t)total, (float4 *)d_out); \
        else pbv::k_devox<float, uint32_t, UU><<<grid, T, 0, st>>>(d_v, (uint32_t)width, d_inv, (uint32_t)total, d_out);      \
    } else {                                                                                                                  \
        if (vec4) pbv::k_devox<float4, unsigned long long, UU><<<grid, T, 0, st>>>((const float4 *)d_v, (unsigned long long)width, d_inv, (unsigned long long)total, (float4 *)d_out); \
        else pbv::k_devox<float, unsigned long long, UU><<<grid, T, 0, st>>>(d_v, (unsigned long long)width, d_inv, (unsigned long long)total, d_out); \
    }
    if (U == 2) { PB_DEVOX(2) } else if (U == 8) { PB_DEVOX(8) } else { PB_DEVOX(4) }
#undef PB_DEVOX
    ctx->launches = 1;
    if (host_io) PB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * C * 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    if (host_io) PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

// =================================================================================================
// proposal IoU / mask labels — see pb_iou.cuh.  Device pointers only (the reference
// asserts is_cuda on every argument, lib/PB_lib/torch_io/pbnet_ops.py:91-94,120-124).
// =================================================================================================
#include "pb_iou.cuh"

extern "C" int pb_cal_iou_and_masklabel(pb_ctx *ctx, const int32_t *proposals_idx, const int32_t *proposals_offset,
                                        const int64_t *instance_labels, const int32_t *instance_pointnum,
                                        float *proposals_iou, int32_t nInstance, int32_t nProposal,
                                        const float *mask_scores_sigmoid, float *mask_label, int mode, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (nInstance < 0 || nProposal < 0 || (mode != 0 && mode != 1)) return fail(ctx, PB_ERR_ARG, "bad argument");
    if (nInstance == 0 || nProposal == 0) return PB_OK;
    if (!proposals_idx || !proposals_offset || !instance_labels || !instance_pointnum || !proposals_iou ||
        (mode == 1 && !mask_scores_sigmoid))
        return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    int rc = ensure_arena(ctx, sizeof(int) * (size_t)nProposal + 4096, st);
    if (rc) return rc;
    int *ptotal = ctx->arena.get<int>((size_t)nProposal);
    long long total = (long long)nProposal * nInstance;
    PB_CUDA(cudaMemsetAsync(proposals_iou, 0, sizeof(float) * (size_t)total, st));
    PB_CUDA(cudaMemsetAsync(ptotal, 0, sizeof(int) * (size_t)nProposal, st));
    pbi::k_iou_count<<<std::min(nProposal, 148 * 16), 256, 0, st>>>(nInstance, nProposal, proposals_idx, proposals_offset,
                                                                    (const long long *)instance_labels, mask_scores_sigmoid,
                                                                    mode, reinterpret_cast<int *>(proposals_iou), ptotal);
    pbi::k_iou_finish<<<(int)std::min<long long>((total + 255) / 256, 148 * 32), 256, 0, st>>>(nInstance, total, instance_pointnum,
                                                                                              ptotal, proposals_iou);
    ctx->launches = 2;
    if (mask_label) {
        pbi::k_mask_label<<<std::min((nProposal + 7) / 8, 148 * 8), 256, 0, st>>>(nInstance, nProposal, proposals_idx,
                                                                                   proposals_offset,
                                                                                   (const long long *)instance_labels,
                                                                                   proposals_iou, mask_label);
        ctx->launches = 3;
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_get_iou(pb_ctx *ctx, const int32_t *proposals_idx, const int32_t *proposals_offset,
                          const int64_t *instance_labels, const int32_t *instance_pointnum, float *proposals_iou,
                          int32_t nInstance, int32_t nProposal, void *stream) {
    return pb_cal_iou_and_masklabel(ctx, proposals_idx, proposals_offset, instance_labels, instance_pointnum,
                                    proposals_iou, nInstance, nProposal, nullptr, nullptr, 0, stream);
}

// =================================================================================================
// local scenes + get_proposal (network/PBNet.py:180-234, 317-346) — see pb_scenes.cuh.
// Device pointers only (this is the device-resident continuation of pb_binary_cluster_batched).
// =================================================================================================
#include "pb_scenes.cuh"

struct pb_scene_state {  // device pointers into the context arena, valid from _plan until the next call on the context
    bool ready = false;
    const pb_ctx *owner = nullptr;
    int K = 0, n = 0, train = 0;
    int64_t P = 0, E = 0;
    int *d_scal = nullptr;  // [0] err [1] P [2] E [3] total members
    int *para = nullptr, *nb = nullptr, *size = nullptr, *mstart = nullptr, *mode = nullptr, *prop_cluster = nullptr;
    long long *prop_offsets = nullptr;
    uint32_t *members = nullptr;
    const long long *label = nullptr;
};
static thread_local pb_scene_state g_scene;  // one context per host thread (pb_create)
static void invalidate_scene_state() { g_scene.ready = false; }

extern "C" int pb_local_scenes_plan(pb_ctx *ctx, const int32_t *cluster_id, const int32_t *seg_counts, int32_t n_seg,
                                    const int32_t *call_seg_counts, const int32_t *call_sem, int32_t n_calls, int64_t n_pts,
                                    const int32_t *cluster_num, const float *center, int64_t n_clusters,
                                    const float *big_thresh20, const int32_t *k_max20, const int64_t *ins_label,
                                    int64_t *n_proposals_out, int64_t *n_entries_out, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    g_scene.ready = false;
    if (n_seg < 0 || n_calls < 0 || n_pts < 0 || n_clusters < 0 || !n_proposals_out || !n_entries_out || !big_thresh20 || !k_max20 ||
        (n_seg > 0 && (!seg_counts || !cluster_num)) || (n_calls > 0 && (!call_seg_counts || !call_sem)))
        return fail(ctx, PB_ERR_ARG, "null / negative argument");
    *n_proposals_out = 0, *n_entries_out = 0;
    int kmax_all = 0;
    for (int c = 0; c < 20; c++) {
        if (k_max20[c] < 0 || k_max20[c] > pbs::kMaxK) return fail(ctx, PB_ERR_ARG, "K_max must be in [0, 16]");
        kmax_all = std::max(kmax_all, (int)k_max20[c]);
    }
    if (n_pts * (1 + (int64_t)kmax_all) >= ((int64_t)1 << 31)) return fail(ctx, PB_ERR_ARG, "n_pts * (1 + K_max) must be below 2^31");
    const int S = n_seg, n = (int)n_pts, K = (int)n_clusters;
    std::vector<int> h_start(S + 1, 0), h_first(S, 0), h_sem(S, 2);
    {
        int s = 0;
        for (int c = 0; c < n_calls; c++) {
            if (call_seg_counts[c] < 0 || s + call_seg_counts[c] > S) return fail(ctx, PB_ERR_ARG, "bad call_seg_counts");
            if (call_sem[c] < 0 || call_sem[c] > 19) return fail(ctx, PB_ERR_SEM_RANGE, "call_sem outside [0,19]");
            for (int k = 0; k < call_seg_counts[c]; k++) h_first[s + k] = s, h_sem[s + k] = call_sem[c];
            s += call_seg_counts[c];
        }
        if (s != S) return fail(ctx, PB_ERR_ARG, "sum(call_seg_counts) != n_seg");
        for (int i = 0; i < S; i++) {
            if (seg_counts[i] < 0 || (int64_t)h_start[i] + seg_counts[i] > n_pts) return fail(ctx, PB_ERR_ARG, "sum(seg_counts) != n_pts");
            h_start[i + 1] = h_start[i] + seg_counts[i];
        }
        if ((S ? h_start[S] : 0) != n) return fail(ctx, PB_ERR_ARG, "sum(seg_counts) != n_pts");
    }
    if (K == 0 || n == 0) {
        g_scene = pb_scene_state();
        g_scene.ready = true, g_scene.owner = ctx;
        return PB_OK;
    }
    if (!cluster_id || !center) return fail(ctx, PB_ERR_ARG, "null data pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    const bool train = ins_label != nullptr;
    const size_t N = (size_t)n;
    int64_t &L = ctx->launches;
    pb_scene_state stt;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        int *d_start = a.get<int>(S + 1), *d_first = a.get<int>(S), *d_sem = a.get<int>(S), *gstart = a.get<int>(S + 2);
        float *d_thr = a.get<float>(20);
        int *d_kmax = a.get<int>(20);
        int *d_scal = a.get<int>(16);
        uint32_t *key = a.get<uint32_t>(N), *key_alt = a.get<uint32_t>(N), *val = a.get<uint32_t>(N), *members = a.get<uint32_t>(N);
        int *size = a.get<int>((size_t)K + 1), *mstart = a.get<int>((size_t)K + 2), *cseg = a.get<int>(K);
        int *para = a.get<int>(K), *nb = a.get<int>((size_t)K * pbs::kMaxK), *len = a.get<int>(K), *valid = a.get<int>(K);
        int *mode = a.get<int>(K), *pidx = a.get<int>((size_t)K + 1), *off_g = a.get<int>((size_t)K + 1), *prop_cluster = a.get<int>(K);
        long long *prop_offsets = a.get<long long>((size_t)K + 1);
        unsigned long long *best = a.get<unsigned long long>(K);
        uint64_t *lkey = train ? a.get<uint64_t>(N) : nullptr, *lkey_alt = train ? a.get<uint64_t>(N) : nullptr;
        int *blocks = a.get<int>(std::max(N, (size_t)K) / pb::kScanTile + 2);
        void *sort_tmp = a.get<char>(RadixTmp::bytes(N, train ? sizeof(uint64_t) : sizeof(uint32_t)));
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
            continue;
        }
        const int T = 256;
        PB_CUDA(cudaMemcpyAsync(d_start, h_start.data(), sizeof(int) * (S + 1), cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(d_first, h_first.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(d_sem, h_sem.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(d_thr, big_thresh20, sizeof(float) * 20, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemcpyAsync(d_kmax, k_max20, sizeof(int) * 20, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaMemsetAsync(d_scal, 0, sizeof(int) * 16, st));
        PB_CUDA(cudaMemsetAsync(size, 0, sizeof(int) * ((size_t)K + 1), st));
        PB_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)K, st));
        // global cluster index of the first cluster of every segment; its total must be n_clusters
        launch_scan(st, cluster_num, S, nullptr, gstart, d_scal + 4, blocks, L);
        PB_CUDA(cudaMemcpyAsync(gstart + S, d_scal + 4, sizeof(int), cudaMemcpyDeviceToDevice, st));
        pbs::k_point_keys<<<div_up(n, T), T, 0, st>>>(n, S, d_start, d_first, gstart, cluster_id, K, key, val, size, d_scal);
        pbs::k_cluster_seg<<<div_up(S, T), T, 0, st>>>(S, gstart, cseg);
        int bits = 1;
        while ((1 << bits) <= K) bits++;
        L += 2;
        {
            int rc = radix_sort<uint32_t>(ctx, sort_tmp, key, key_alt, val, members, n, bits, st, &L);
            if (rc) return rc;
        }
        launch_scan(st, size, K + 1, nullptr, mstart, d_scal + 3, blocks, L);
        if (train) {
            pbs::k_label_keys<<<div_up(n, T), T, 0, st>>>(n, key, (const long long *)ins_label, K, lkey, d_scal);
            {
                int rc = radix_sort<uint64_t>(ctx, sort_tmp, lkey, lkey_alt, nullptr, nullptr, n, 32 + bits, st, &L);
                if (rc) return rc;
            }
            pbs::k_label_mode<<<div_up(n, T), T, 0, st>>>(n, lkey_alt, K, best);
            L += 2;
        }
        pbs::k_cluster_plan<<<div_up(K, 128), 128, 0, st>>>(K, cseg, gstart, d_sem, center, size, d_thr, d_kmax, best, train ? 1 : 0,
                                                            para, nb, len, valid, mode);
        L++;
        launch_scan(st, valid, K, nullptr, pidx, d_scal + 1, blocks, L);
        launch_scan(st, len, K, nullptr, off_g, d_scal + 2, blocks, L);
        pbs::k_proposals<<<div_up(K, T), T, 0, st>>>(K, valid, pidx, off_g, d_scal + 2, d_scal + 1, prop_offsets, prop_cluster);
        L++;
        PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_scal, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaStreamSynchronize(st));
        if (ctx->h_scalars[0] & pbs::kErrId) return fail(ctx, PB_ERR_ARG, "a cluster id lies outside the id range of its segment (cluster_num / call tables do not match cluster_id)");
        if (ctx->h_scalars[0] & pbs::kErrLabel) return fail(ctx, PB_ERR_ARG, "instance label outside the int32 range");
        if (ctx->h_scalars[4] != K) return fail(ctx, PB_ERR_ARG, "sum(cluster_num) != n_clusters");
        stt.K = K, stt.n = n, stt.train = train ? 1 : 0, stt.P = ctx->h_scalars[1], stt.E = ctx->h_scalars[2];
        stt.d_scal = d_scal, stt.para = para, stt.nb = nb, stt.size = size, stt.mstart = mstart, stt.mode = mode;
        stt.prop_cluster = prop_cluster, stt.prop_offsets = prop_offsets, stt.members = members;
        stt.label = (const long long *)ins_label;
        stt.ready = true, stt.owner = ctx;
    }
    g_scene = stt;
    *n_proposals_out = stt.P;
    *n_entries_out = stt.E;
    return PB_OK;
}

extern "C" int pb_local_scenes_fill(pb_ctx *ctx, const int64_t *point_map, int64_t *prop_offsets, int32_t *prop_cluster,
                                    int64_t *prop_index, float *prop_dpn, int32_t *prop_gt, int32_t *prop_id, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    const pb_scene_state &s = g_scene;
    if (!s.ready || s.owner != ctx) return fail(ctx, PB_ERR_ARG, "pb_local_scenes_fill must directly follow a successful pb_local_scenes_plan on this thread");
    if (!prop_offsets) return fail(ctx, PB_ERR_ARG, "null prop_offsets");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    if (s.P == 0) {
        PB_CUDA(cudaMemsetAsync(prop_offsets, 0, sizeof(int64_t), st));
        return PB_OK;
    }
    if (!prop_index || !prop_dpn || (prop_gt && !s.train)) return fail(ctx, PB_ERR_ARG, "null output / prop_gt without instance labels");
    PB_CUDA(cudaMemcpyAsync(prop_offsets, s.prop_offsets, sizeof(int64_t) * (size_t)(s.P + 1), cudaMemcpyDeviceToDevice, st));
    if (prop_cluster) PB_CUDA(cudaMemcpyAsync(prop_cluster, s.prop_cluster, sizeof(int) * (size_t)s.P, cudaMemcpyDeviceToDevice, st));
    if (s.E > 0) {
        const int T = 256;
        int grid = (int)std::min<int64_t>((s.E + T - 1) / T, 148 * 16);
        pbs::k_fill<<<grid, T, 0, st>>>(s.d_scal + 2, s.d_scal + 1, s.prop_offsets, s.prop_cluster, s.para, s.nb, s.size, s.mstart,
                                        s.members, (const long long *)point_map, s.label, s.mode, (long long *)prop_index, prop_dpn,
                                        prop_gt, prop_id);
        ctx->launches = 1;
    }
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_get_proposal(pb_ctx *ctx, const int64_t *prop_offsets, int64_t n_proposals, const int64_t *point_idx,
                               const float *mask_score, int64_t n_entries, float thd, int64_t *proposals_idx,
                               int64_t *proposals_offset, int64_t *cluster_id_v, float *proposals_ms, int64_t *n_kept_out,
                               int64_t *n_nonempty_out, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n_proposals < 0 || n_entries < 0 || !n_kept_out || !n_nonempty_out || n_entries >= ((int64_t)1 << 31) ||
        n_proposals >= ((int64_t)1 << 31))
        return fail(ctx, PB_ERR_ARG, "bad argument");
    *n_kept_out = 0, *n_nonempty_out = 0;
    if (!proposals_offset) return fail(ctx, PB_ERR_ARG, "null proposals_offset");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    if (n_proposals == 0 || n_entries == 0) {
        PB_CUDA(cudaMemsetAsync(proposals_offset, 0, sizeof(int64_t), st));
        return PB_OK;
    }
    if (!prop_offsets || !point_idx || !mask_score || !proposals_idx || !cluster_id_v || !proposals_ms)
        return fail(ctx, PB_ERR_ARG, "null pointer");
    const int E = (int)n_entries, P = (int)n_proposals;
    int64_t &L = ctx->launches;
    size_t need = sizeof(int) * ((size_t)2 * E + (size_t)3 * P + (size_t)std::max(E, P) / pb::kScanTile + 64) + 8192;
    int rc = ensure_arena(ctx, need, st);
    if (rc) return rc;
    Arena &a = ctx->arena;
    int *flag = a.get<int>(E), *kpos = a.get<int>(E), *nonempty = a.get<int>(P), *newid = a.get<int>(P), *kstart = a.get<int>(P);
    int *d_scal = a.get<int>(16);
    int *blocks = a.get<int>((size_t)std::max(E, P) / pb::kScanTile + 2);
    const int T = 256;
    pbs::k_score_flags<<<div_up(E, T), T, 0, st>>>(E, mask_score, thd, flag);
    launch_scan(st, flag, E, nullptr, kpos, d_scal, blocks, L);
    pbs::k_prop_counts<<<div_up(P, T), T, 0, st>>>(P, E, (const long long *)prop_offsets, kpos, d_scal, nonempty, kstart);
    launch_scan(st, nonempty, P, nullptr, newid, d_scal + 1, blocks, L);
    pbs::k_prop_write<<<div_up(P, T), T, 0, st>>>(P, nonempty, newid, kstart, d_scal, d_scal + 1, (long long *)proposals_offset,
                                                  (long long *)cluster_id_v);
    pbs::k_prop_entries<<<div_up(E, T), T, 0, st>>>(E, flag, kpos, (const long long *)prop_offsets, P, newid,
                                                    (const long long *)point_idx, mask_score, (long long *)proposals_idx, proposals_ms);
    L += 4;
    PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_scal, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    *n_kept_out = ctx->h_scalars[0];
    *n_nonempty_out = ctx->h_scalars[1];
    return PB_OK;
}

extern "C" int pb_scene_features(pb_ctx *ctx, const float *point_feat, int32_t C, const float *sem_score, int32_t n_cls,
                                 const int64_t *index, const int32_t *prop_id, const int32_t *prop_sem, const float *dpn,
                                 int64_t n_entries, float *out, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (n_entries < 0 || C < 1 || n_cls < 1) return fail(ctx, PB_ERR_ARG, "bad argument");
    if (n_entries == 0) return PB_OK;
    if (!point_feat || !sem_score || !index || !prop_id || !prop_sem || !dpn || !out) return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    const int T = 256;
    int grid = (int)std::min<int64_t>((n_entries * 32 + T - 1) / T, 148 * 32);
    pbs::k_scene_feat<<<grid, T, 0, st>>>(n_entries, C, n_cls, point_feat, sem_score, (const long long *)index, prop_id, prop_sem, dpn, out);
    ctx->launches = 1;
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

// =================================================================================================
// evaluation post-processing of proposals (eval_map.py:63-121, tools/mIOU.py:77-87,
// tools/getins.py:72-98) — see pb_eval.cuh.  Device pointers only.
// =================================================================================================
#include "pb_eval.cuh"

static int bit_width_u64(unsigned long long v) {
    int b = 0;
    while (v) b++, v >>= 1;
    return std::max(b, 1);
}

extern "C" int pb_eval_postprocess(pb_ctx *ctx, const int64_t *proposals_idx, int64_t n_entries, const int64_t *proposals_offset,
                                   int64_t n_proposals, const float *clt_score, const int64_t *pred_sem, int64_t point_num,
                                   int32_t copies, const int64_t *superpoint, int64_t n_superpoints, const int64_t *sem_table,
                                   int32_t n_table, float score_thresh, int32_t npoint_thresh, float nms_thresh, int32_t *label,
                                   float *cluster_scores, int64_t *cluster_sem, int32_t *cluster_proposal, int64_t cap,
                                   int64_t *n_clusters_out, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (!n_clusters_out || n_entries < 0 || n_proposals < 0 || point_num < 0 || copies < 1 || n_superpoints < 0 || n_table < 1 ||
        n_entries >= ((int64_t)1 << 31) || n_proposals >= ((int64_t)1 << 24) || point_num >= ((int64_t)1 << 31) || point_num % copies != 0)
        return fail(ctx, PB_ERR_ARG, "bad argument (sizes must be non-negative, point_num a multiple of copies)");
    *n_clusters_out = 0;
    const int n3 = (int)(point_num / copies), P = (int)n_proposals, nsp = (int)n_superpoints;
    const long long M = n_entries;
    if (n3 == 0) return PB_OK;
    if (!label || !superpoint || !sem_table) return fail(ctx, PB_ERR_ARG, "null pointer");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, true);
    const int T = 256;
    int64_t &L = ctx->launches;
    if (P == 0 || M == 0) {
        pbe::k_fill_i32<<<div_up(n3, T), T, 0, st>>>(n3, -100, label);
        L = 1;
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    }
    if (!proposals_idx || !proposals_offset || !clt_score || !pred_sem || !cluster_scores || !cluster_sem || !cluster_proposal)
        return fail(ctx, PB_ERR_ARG, "null pointer");
    const int vmax = std::min(P, pbe::kMaxValid);
    const size_t Mz = (size_t)std::max<long long>(M, n3);
    // ---- workspace -------------------------------------------------------------------------------------------------
    uint64_t *key = nullptr, *key_alt = nullptr, *key2 = nullptr, *key2_alt = nullptr;
    int *head = nullptr, *npoint = nullptr, *valid = nullptr, *vid = nullptr, *vlist = nullptr, *d_scal = nullptr, *blocks = nullptr;
    int *inter = nullptr, *order = nullptr, *pick_rank = nullptr, *picked = nullptr, *alive = nullptr, *newid = nullptr;
    long long *prop_sem = nullptr, *d_table = nullptr;
    unsigned long long *best = nullptr;
    void *sort_tmp = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        key = a.get<uint64_t>(Mz), key_alt = a.get<uint64_t>(Mz), key2 = a.get<uint64_t>(Mz), key2_alt = a.get<uint64_t>(Mz);
        head = a.get<int>((size_t)M), npoint = a.get<int>(P), valid = a.get<int>(P), vid = a.get<int>((size_t)P + 1), vlist = a.get<int>(P);
        prop_sem = a.get<long long>(P), d_table = a.get<long long>(n_table);
        d_scal = a.get<int>(16), blocks = a.get<int>(Mz / pb::kScanTile + 2);
        inter = a.get<int>((size_t)vmax * vmax), order = a.get<int>(vmax), pick_rank = a.get<int>(vmax), picked = a.get<int>(vmax);
        alive = a.get<int>(vmax), newid = a.get<int>((size_t)vmax + 1);
        best = a.get<unsigned long long>((size_t)std::max(nsp, 1));
        sort_tmp = a.get<char>(RadixTmp::bytes(Mz, sizeof(uint64_t)));
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
        }
    }
    PB_CUDA(cudaMemcpyAsync(d_table, sem_table, sizeof(long long) * n_table, cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemsetAsync(d_scal, 0, sizeof(int) * 16, st));
    PB_CUDA(cudaMemsetAsync(npoint, 0, sizeof(int) * (size_t)P, st));
    int *d_err = d_scal, *d_V = d_scal + 1, *d_C = d_scal + 2, *d_C2 = d_scal + 3;
    // ---- distinct (proposal, folded point) pairs, thresholds -----------------------------------------------------------
    pbe::k_pair_keys<<<div_up(std::max<long long>(M, P), T), T, 0, st>>>(M, P, point_num, n3, (const long long *)proposals_idx,
                                                                          (const long long *)proposals_offset, (const long long *)pred_sem,
                                                                          d_table, n_table, key, prop_sem, d_err);
    int bits1 = bit_width_u64((unsigned long long)P * (unsigned long long)n3 + 1);
    {
        int rc = radix_sort<uint64_t>(ctx, sort_tmp, key, key_alt, nullptr, nullptr, (int)M, bits1, st, nullptr);
        if (rc) return rc;
    }
    pbe::k_pair_heads<<<div_up(M, T), T, 0, st>>>(M, n3, key_alt, head, npoint);
    pbe::k_valid<<<div_up(P, T), T, 0, st>>>(P, clt_score, score_thresh, npoint, npoint_thresh, valid);
    L += 3 + 2 + (bits1 + 7) / 8;
    launch_scan(st, valid, P, nullptr, vid, d_V, blocks, L);
    pbe::k_valid_list<<<div_up(P, T), T, 0, st>>>(P, valid, vid, vlist);
    L++;
    PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_scal, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_scalars[0] & pbe::kErrProp) return fail(ctx, PB_ERR_ARG, "proposals_idx[:,0] outside [0, n_proposals)");
    if (ctx->h_scalars[0] & pbe::kErrPoint) return fail(ctx, PB_ERR_ARG, "proposals_idx[:,1] outside [0, point_num)");
    if (ctx->h_scalars[0] & pbe::kErrSem) return fail(ctx, PB_ERR_SEM_RANGE, "pred_sem outside the class table");
    const int V = ctx->h_scalars[1];
    if (V > pbe::kMaxValid) return fail(ctx, PB_ERR_CAPACITY, "more than 4096 proposals pass the score / point-count thresholds");
    if (V == 0) {  // eval_map.py:88-89, 101-103: nothing to pick
        pbe::k_fill_i32<<<div_up(n3, T), T, 0, st>>>(n3, -100, label);
        L++;
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    }
    if (cap < V) return fail(ctx, PB_ERR_CAPACITY, "output capacity below the number of proposals that pass the thresholds");
    // ---- cross intersections, NMS ------------------------------------------------------------------------------------------
    PB_CUDA(cudaMemsetAsync(inter, 0, sizeof(int) * (size_t)V * V, st));
    pbe::k_point_keys<<<div_up(M, T), T, 0, st>>>(M, n3, key_alt, head, valid, vid, d_V, key2);
    int bits2 = bit_width_u64((unsigned long long)n3 * (unsigned long long)V + 1);
    {
        int rc = radix_sort<uint64_t>(ctx, sort_tmp, key2, key2_alt, nullptr, nullptr, (int)M, bits2, st, nullptr);
        if (rc) return rc;
    }
    pbe::k_intersections<<<div_up(M, T), T, 0, st>>>(M, key2_alt, d_V, inter);
    pbe::k_nms<<<1, 1024, 0, st>>>(d_V, vlist, clt_score, inter, nms_thresh, order, pick_rank, picked, d_C);
    // ---- per-point labels, superpoint vote, rebuilt clusters ---------------------------------------------------------------
    pbe::k_fill_i32<<<div_up(n3, T), T, 0, st>>>(n3, -1, label);
    pbe::k_paint<<<div_up(M, T), T, 0, st>>>(M, key2_alt, d_V, pick_rank, label);
    pbe::k_vote_keys<<<div_up(n3, T), T, 0, st>>>(n3, (const long long *)superpoint, nsp, label, d_C, key, d_err);
    int bits3 = bit_width_u64((unsigned long long)std::max(nsp, 1) * (unsigned long long)(V + 1) + 1);
    {
        int rc = radix_sort<uint64_t>(ctx, sort_tmp, key, key_alt, nullptr, nullptr, n3, bits3, st, nullptr);
        if (rc) return rc;
    }
    PB_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)std::max(nsp, 1), st));
    PB_CUDA(cudaMemsetAsync(alive, 0, sizeof(int) * (size_t)V, st));
    PB_CUDA(cudaMemcpyAsync(ctx->h_scalars + 4, d_scal, sizeof(int), cudaMemcpyDeviceToHost, st));  // superpoint range check
    PB_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_scalars[4] & pbe::kErrSuper) return fail(ctx, PB_ERR_ARG, "superpoint id outside [0, n_superpoints)");
    pbe::k_vote_count<<<div_up(n3, T), T, 0, st>>>(n3, key_alt, d_C, best);
    pbe::k_align<<<div_up(n3, T), T, 0, st>>>(n3, (const long long *)superpoint, best, d_C, label, alive);
    L += 9 + 4 + (bits2 + 7) / 8 + (bits3 + 7) / 8;
    launch_scan(st, alive, V, nullptr, newid, d_C2, blocks, L);
    pbe::k_finish<<<div_up(std::max(n3, V), T), T, 0, st>>>(n3, d_C, alive, newid, picked, clt_score, prop_sem, label, cluster_scores,
                                                            (long long *)cluster_sem, cluster_proposal);
    L++;
    PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_scal, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    *n_clusters_out = ctx->h_scalars[3];
    return PB_OK;
}

// =================================================================================================
// mesh vertex normals (lib/PB_lib/src/normal/cal_normal.cu, PB_lib_api.cpp:10) — see pb_normals.cuh
// =================================================================================================
#include "pb_normals.cuh"

extern "C" int pb_cal_normal_line(pb_ctx *ctx, const float *xyz, const int32_t *face, float *normal_xyz, int32_t num_vtx,
                                  int32_t num_face, int mem_kind, void *stream_v) {
    if (!ctx) return PB_ERR_ARG;
    ctx->err.clear();
    ctx->launches = 0;
    if (num_vtx < 0 || num_face < 0 || (mem_kind != PB_MEM_HOST && mem_kind != PB_MEM_DEVICE)) return fail(ctx, PB_ERR_ARG, "bad argument");
    if (num_vtx == 0) return PB_OK;
    if (!xyz || !normal_xyz || (num_face > 0 && !face)) return fail(ctx, PB_ERR_ARG, "null pointer");
    if ((long long)num_face * 3 >= (1LL << 31)) return fail(ctx, PB_ERR_ARG, "num_face too large");
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream_v, mem_kind == PB_MEM_DEVICE);
    const bool host_io = mem_kind == PB_MEM_HOST;
    const size_t V = (size_t)num_vtx, F = (size_t)num_face, NK = 3 * F;
    float *d_xyz = nullptr, *d_out = nullptr, *fnormal = nullptr, *farea = nullptr;
    int *d_face = nullptr, *d_err = nullptr;
    uint64_t *key = nullptr, *key_alt = nullptr;
    void *sort_tmp = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        Arena dry;
        dry.dry = true;
        Arena &a = pass == 0 ? dry : ctx->arena;
        d_xyz = host_io ? a.get<float>(3 * V) : nullptr;
        d_out = host_io ? a.get<float>(3 * V) : nullptr;
        d_face = host_io ? a.get<int>(std::max<size_t>(NK, 1)) : nullptr;
        fnormal = a.get<float>(std::max<size_t>(NK, 1)), farea = a.get<float>(std::max<size_t>(F, 1));
        key = a.get<uint64_t>(std::max<size_t>(NK, 1)), key_alt = a.get<uint64_t>(std::max<size_t>(NK, 1));
        d_err = a.get<int>(4);
        sort_tmp = a.get<char>(RadixTmp::bytes(std::max<size_t>(NK, 1), sizeof(uint64_t)));
        if (pass == 0) {
            int rc = ensure_arena(ctx, dry.off, st);
            if (rc) return rc;
        }
    }
    const float *x_in = xyz;
    const int *f_in = face;
    float *o = normal_xyz;
    if (host_io) {
        PB_CUDA(cudaMemcpyAsync(d_xyz, xyz, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, st));
        if (NK) PB_CUDA(cudaMemcpyAsync(d_face, face, sizeof(int) * NK, cudaMemcpyHostToDevice, st));
        x_in = d_xyz, f_in = d_face, o = d_out;
    }
    PB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int) * 4, st));
    const int T = 256;
    const uint64_t *skey = key_alt;
    if (num_face > 0) {
        pbn::k_face_normals<<<div_up(num_face, T), T, 0, st>>>(x_in, f_in, num_face, num_vtx, fnormal, farea, key, d_err);
        // key = vertex << 32 | face (all ones for the unused corners of degenerate faces): 32 + bits(num_vtx) key bits
        int rc = radix_sort<uint64_t>(ctx, sort_tmp, key, key_alt, nullptr, nullptr, (int)NK, 32 + pb::bit_width_i(num_vtx), st, &ctx->launches);
        if (rc) return rc;
        ctx->launches += 1;
    }
    pbn::k_vertex_normals<<<div_up(num_vtx, T), T, 0, st>>>(num_vtx, (long long)NK, skey, fnormal, farea, o);
    ctx->launches++;
    PB_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (host_io) PB_CUDA(cudaMemcpyAsync(normal_xyz, d_out, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_scalars[0]) return fail(ctx, PB_ERR_ARG, "face lists a vertex index outside [0, num_vtx)");
    return PB_OK;
}
