/*
 * oracle/pb_oracle.c — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the reference's instance-grouping path
 *   PB_lib.binary_cluster  (lib/PB_lib/src/pbnet/cluster.cu:16-119)
 * step by step, in plain C with exact fp32 arithmetic (explicit fmaf, compiled with
 * -ffp-contract=off).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file; the shipped library never links it.
 *
 * Two entry points with identical outputs:
 *   pb_oracle_literal  — O(n^2), no acceleration structure: the plainest possible transcription
 *                        (materialised neighbour lists, level-synchronous BFS, brute-force 1-NN).
 *   pb_oracle_grid     — same semantics; a uniform grid only ENUMERATES candidates, the accept
 *                        test is the exact predicate, so outputs are bit-identical to _literal.
 *
 * Pinning status: cross-checked against the compiled reference (oracle/_ref, built from
 * /root/reference by oracle/build_ref.py) on the GPU box by tests/golden/make_golden.py (147/147 cases bit for
 * bit, tests/golden/crosscheck_report_r01.json); the committed fixtures under tests/golden/ are re-checked on CPU by
 * tests/test_oracle_golden.py and live against oracle/_ref by tests/test_gpu_shim.py.
 *
 * Reference map (file:line under /root/reference/lib/PB_lib/src/pbnet/):
 *   predicate            binary_cuda_functions.cu:305-308 (+ SASS FMA contraction), :85, :160-161
 *   degree / den_queue   binary.cu:71-103, binary_cuda_functions.cu:29-89
 *   HP rule              binary_cuda_functions.cu:175-186
 *   BFS clustering       binary.cu:154-217, binary_cuda_functions.cu:197-215
 *   fragment filter      binary.cu:219-268, binary_cuda_functions.cu:249-256
 *   LP 1-NN assignment   binary.cu:270-358, binary_cuda_functions.cu:258-302
 *   centres              binary.cu:360-415, binary_cuda_functions.cu:217-246
 *   segment loop         cluster.cu:57-118
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PB_OK 0
#define PB_ERR_ARG 1
#define PB_ERR_NOMEM 2

/* binary.cu:229 */
static const float k_mean_count[18] = {3917.0f, 12056.0f, 2303.0f, 8331.0f, 3948.0f, 3166.0f,
                                       5629.0f, 11719.0f, 1003.0f, 3317.0f, 4912.0f, 10221.0f,
                                       3889.0f, 4136.0f,  2120.0f, 945.0f,  3967.0f, 2589.0f};

/* binary_cuda_functions.cu:305-308 as nvcc 12.9 compiles it for sm_100a.  SASS of all three call
 * sites (k_num_nbs, k_append_neighbours, noise_id_cluster; `cuobjdump -sass oracle/_ref/PB_lib*.so`)
 * reads  FADD dy; FMUL t=dy*dy; FADD dx; FFMA t=dx*dx+t; FADD dz; FFMA D=dz*dz+t  — i.e. the product
 * that stays a separate rounded multiply is the MIDDLE term:
 *     D = fma(dz, dz, fma(dx, dx, fl(dy*dy)))
 * (first GPU cross-check against the compiled reference caught the dx*dx-first guess being wrong on
 * pairs within 1 ulp of r^2). */
static inline float sqd(float x1, float y1, float z1, float x2, float y2, float z2) {
    float dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

typedef struct {
    int64_t pair_tests; /* candidate tests in the degree pass */
    int64_t sum_deg;    /* sum of degrees */
    int64_t n_hp;       /* number of HPs */
    int64_t n_noise;    /* points entering LP assignment */
    int64_t nn_tests;   /* candidate tests in the 1-NN pass */
} pb_oracle_stats;

/* ---------------------------------------------------------------------------------------------
 * shared tail: fragment filter (binary.cu:219-268) on one segment.
 * ids hold raw ids in [start, start+num_cluster) or -1.  Returns number of survivors.
 * ------------------------------------------------------------------------------------------- */
static int filter_segment(int n, const int *sem, int *ids, int num_cluster, int start, float para_f,
                          int *clt_sem_out /* append target */, int *n_clt_sem) {
    if (num_cluster == 0) return 0;
    int *clt_semv = (int *)calloc(num_cluster, sizeof(int));
    int *clt_num = (int *)calloc(num_cluster, sizeof(int));
    int *remap = (int *)malloc(sizeof(int) * num_cluster);
    for (int i = 0; i < n; i++) { /* :241-247 — class = class of the highest-index member */
        int c = ids[i];
        if (c != -1) {
            clt_num[c - start]++;
            clt_semv[c - start] = sem[i];
        }
    }
    int kept = 0;
    for (int i = 0; i < num_cluster; i++) { /* :252-264 */
        int cur_sem = clt_semv[i] - 2;
        float cur_mean_count = k_mean_count[cur_sem] * para_f; /* fp32 multiply */
        if ((float)clt_num[i] < cur_mean_count) {
            remap[i] = -1;
        } else {
            remap[i] = start + kept;
            clt_sem_out[(*n_clt_sem)++] = clt_semv[i];
            kept++;
        }
    }
    for (int i = 0; i < n; i++)
        if (ids[i] != -1) ids[i] = remap[ids[i] - start]; /* net effect of repeated shift_con_clt */
    free(clt_semv);
    free(clt_num);
    free(remap);
    return kept;
}

/* centres (binary_cuda_functions.cu:217-246): sequential running mean, IEEE division */
static void centres_segment(int n, const float *x, const float *y, const float *z, const int *ids,
                            int num_cluster, int start, float *center_out, int *n_center) {
    float *mx = (float *)calloc(num_cluster, sizeof(float));
    float *my = (float *)calloc(num_cluster, sizeof(float));
    float *mz = (float *)calloc(num_cluster, sizeof(float));
    int *cnt = (int *)calloc(num_cluster, sizeof(int));
    for (int i = 0; i < n; i++) {
        int c = ids[i];
        if (c < start || c >= start + num_cluster) continue;
        int k = c - start;
        cnt[k]++;
        float N = (float)cnt[k];
        mx[k] = mx[k] + (x[i] - mx[k]) / N;
        my[k] = my[k] + (y[i] - my[k]) / N;
        mz[k] = mz[k] + (z[i] - mz[k]) / N;
    }
    for (int k = 0; k < num_cluster; k++) {
        center_out[(*n_center)++] = mx[k];
        center_out[(*n_center)++] = my[k];
        center_out[(*n_center)++] = mz[k];
    }
    free(mx);
    free(my);
    free(mz);
    free(cnt);
}

/* =============================================================================================
 * LITERAL variant (O(n^2))
 * =========================================================================================== */
static int literal_segment(int n, const float *x, const float *y, const float *z, const float *xo,
                           const float *yo, const float *zo, const int *sem, const float *radius,
                           const int *min_pts, float para_f, int nv_flag, int *ids, int *den,
                           int acc, int *clt_sem_out, int *n_clt_sem, pb_oracle_stats *st) {
    /* degree + neighbour lists (binary.cu:71-128) */
    int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    int *deg = (int *)malloc(sizeof(int) * n);
    for (int u = 0; u < n; u++) {
        float r = radius[sem[u] - 2];
        float r2 = r * r;
        int ans = 0;
        for (int v = 0; v < n; v++) ans += (sqd(x[u], y[u], z[u], x[v], y[v], z[v]) <= r2);
        deg[u] = ans - 1; /* :88 */
        den[u] = deg[u];
    }
    start[0] = 0;
    for (int u = 0; u < n; u++) start[u + 1] = start[u] + deg[u];
    int *nbr = (int *)malloc(sizeof(int) * (start[n] > 0 ? start[n] : 1));
    for (int u = 0; u < n; u++) {
        float r = radius[sem[u] - 2];
        float r2 = r * r;
        int64_t p = start[u];
        for (int v = 0; v < n; v++)
            if (v != u && sqd(x[u], y[u], z[u], x[v], y[v], z[v]) <= r2) nbr[p++] = v;
    }
    if (st) {
        st->pair_tests += (int64_t)n * n;
        st->sum_deg += start[n];
    }
    /* HP rule (binary_cuda_functions.cu:175-186) */
    int *member = (int *)malloc(sizeof(int) * n);
    for (int u = 0; u < n; u++) {
        member[u] = (deg[u] >= min_pts[sem[u] - 2]) ? 0 : 2;
        if (st && member[u] == 0) st->n_hp++;
    }
    /* identify_clusters + bfs_sem (binary.cu:154-217): level-synchronous boolean frontier */
    unsigned char *visited = (unsigned char *)malloc(n);
    unsigned char *frontier = (unsigned char *)malloc(n);
    unsigned char *next = (unsigned char *)malloc(n);
    int cluster = acc;
    for (int u = 0; u < n; u++) {
        if (ids[u] == -1 && member[u] == 0) {
            memset(visited, 0, n);
            memset(frontier, 0, n);
            frontier[u] = 1;
            int nf = 1;
            while (nf > 0) {
                memcpy(next, frontier, n);
                for (int v = 0; v < n; v++) {
                    if (!frontier[v]) continue;
                    next[v] = 0;
                    visited[v] = 1;
                }
                for (int v = 0; v < n; v++) {
                    if (!frontier[v]) continue;
                    if (member[v] != 0) continue;
                    for (int64_t i = start[v]; i < start[v + 1]; i++)
                        if (!visited[nbr[i]]) next[nbr[i]] = 1;
                }
                nf = 0;
                for (int v = 0; v < n; v++) {
                    frontier[v] = next[v];
                    nf += next[v];
                }
            }
            int sem_cur = sem[u];
            for (int v = 0; v < n; v++) {
                if (visited[v] && sem[v] == sem_cur) {
                    ids[v] = cluster; /* overwrites earlier labels of border LPs */
                    if (member[v] != 0) member[v] = 1;
                }
            }
            cluster++;
        }
    }
    free(visited);
    free(frontier);
    free(next);
    free(nbr);
    free(start);
    free(deg);
    free(member);
    /* filter */
    int K = filter_segment(n, sem, ids, cluster - acc, acc, para_f, clt_sem_out, n_clt_sem);
    /* assigned_LPs (binary.cu:270-358, kernel :258-302) */
    if (nv_flag) {
        int n_lab = 0;
        int *lab = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
        for (int i = 0; i < n; i++)
            if (ids[i] != -1) lab[n_lab++] = i;
        if (n_lab < n) {
            int *newid = (int *)malloc(sizeof(int) * n);
            memcpy(newid, ids, sizeof(int) * n);
            for (int p = 0; p < n; p++) {
                if (ids[p] != -1) continue;
                if (st) st->n_noise++;
                float min_dist = 0.f;
                int min_index = 0, count_i = 0, real = 0;
                for (int i = 0; i < n_lab; i++) {
                    real = lab[i];
                    if (sem[real] != sem[p]) continue;
                    float d = sqd(xo[p], yo[p], zo[p], xo[real], yo[real], zo[real]);
                    if (count_i == 0) min_dist = d;
                    count_i++;
                    if (d <= min_dist) {
                        min_dist = d;
                        min_index = real;
                    }
                }
                if (count_i == 0 && n_lab > 0) min_index = real; /* :287-300: last labelled point */
                newid[p] = ids[min_index]; /* min_index==0 when n_lab==0: ids[0] is -1 then */
            }
            memcpy(ids, newid, sizeof(int) * n);
            free(newid);
        }
        free(lab);
    }
    return K;
}

/* =============================================================================================
 * GRID variant
 * =========================================================================================== */
typedef struct {
    double minx, miny, minz, h;
    int64_t nx, ny, nz;
    int n;        /* points in the grid */
    int *ord;     /* point ids sorted by cell key */
    int C;        /* occupied cells */
    int64_t *ckey;
    int *cstart;  /* C+1 */
    int *cell_of; /* per sorted position: cell ordinal */
} grid_t;

typedef struct {
    int64_t key;
    int id;
} kv_t;

static int kv_cmp(const void *a, const void *b) {
    const kv_t *p = (const kv_t *)a, *q = (const kv_t *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->id < q->id ? -1 : (p->id > q->id);
}

static inline void grid_cell(const grid_t *g, float x, float y, float z, int64_t *cx, int64_t *cy,
                             int64_t *cz) {
    *cx = (int64_t)floor(((double)x - g->minx) / g->h);
    *cy = (int64_t)floor(((double)y - g->miny) / g->h);
    *cz = (int64_t)floor(((double)z - g->minz) / g->h);
}

static int grid_build(grid_t *g, const float *x, const float *y, const float *z, const int *ids,
                      int n, double h) {
    memset(g, 0, sizeof(*g));
    g->h = h;
    g->n = n;
    if (n == 0) return PB_OK;
    double mnx = 1e300, mny = 1e300, mnz = 1e300, mxx = -1e300, mxy = -1e300, mxz = -1e300;
    for (int i = 0; i < n; i++) {
        int p = ids ? ids[i] : i;
        if (x[p] < mnx) mnx = x[p];
        if (y[p] < mny) mny = y[p];
        if (z[p] < mnz) mnz = z[p];
        if (x[p] > mxx) mxx = x[p];
        if (y[p] > mxy) mxy = y[p];
        if (z[p] > mxz) mxz = z[p];
    }
    g->minx = mnx;
    g->miny = mny;
    g->minz = mnz;
    g->nx = (int64_t)floor((mxx - mnx) / h) + 1;
    g->ny = (int64_t)floor((mxy - mny) / h) + 1;
    g->nz = (int64_t)floor((mxz - mnz) / h) + 1;
    if ((double)g->nx * (double)g->ny * (double)g->nz > 9e18) return PB_ERR_ARG;
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * n);
    if (!kv) return PB_ERR_NOMEM;
    for (int i = 0; i < n; i++) {
        int p = ids ? ids[i] : i;
        int64_t cx, cy, cz;
        grid_cell(g, x[p], y[p], z[p], &cx, &cy, &cz);
        kv[i].key = (cz * g->ny + cy) * g->nx + cx;
        kv[i].id = p;
    }
    qsort(kv, n, sizeof(kv_t), kv_cmp);
    g->ord = (int *)malloc(sizeof(int) * n);
    g->cell_of = (int *)malloc(sizeof(int) * n);
    int C = 0;
    for (int i = 0; i < n; i++)
        if (i == 0 || kv[i].key != kv[i - 1].key) C++;
    g->C = C;
    g->ckey = (int64_t *)malloc(sizeof(int64_t) * C);
    g->cstart = (int *)malloc(sizeof(int) * (C + 1));
    int c = -1;
    for (int i = 0; i < n; i++) {
        if (i == 0 || kv[i].key != kv[i - 1].key) {
            c++;
            g->ckey[c] = kv[i].key;
            g->cstart[c] = i;
        }
        g->ord[i] = kv[i].id;
        g->cell_of[i] = c;
    }
    g->cstart[C] = n;
    free(kv);
    return PB_OK;
}

static void grid_free(grid_t *g) {
    free(g->ord);
    free(g->ckey);
    free(g->cstart);
    free(g->cell_of);
    memset(g, 0, sizeof(*g));
}

/* first cell ordinal with key >= k */
static int grid_lower(const grid_t *g, int64_t k) {
    int lo = 0, hi = g->C;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g->ckey[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

/* sorted-position range [*b,*e) of the x-run cells (cx0..cx1, cy, cz); returns 0 if outside grid */
static int grid_run(const grid_t *g, int64_t cx0, int64_t cx1, int64_t cy, int64_t cz, int *b, int *e) {
    if (cy < 0 || cy >= g->ny || cz < 0 || cz >= g->nz) return 0;
    if (cx0 < 0) cx0 = 0;
    if (cx1 >= g->nx) cx1 = g->nx - 1;
    if (cx0 > cx1) return 0;
    int64_t base = (cz * g->ny + cy) * g->nx;
    int c0 = grid_lower(g, base + cx0);
    int c1 = grid_lower(g, base + cx1 + 1);
    if (c0 >= c1) return 0;
    *b = g->cstart[c0];
    *e = g->cstart[c1];
    return 1;
}

/* ---- LP assignment of one range of points (the grid variant's per-query search); queries are independent
 *      (binary_cuda_functions.cu:258-302 runs one thread per noise point), so large segments — the C3 / C4 stress cases of
 *      bench.py — spread them over the host cores with pthreads; results do not depend on it ---- */
typedef struct {
    const grid_t *lgp;
    const float *xo, *yo, *zo;
    const int *sem, *ids;
    int *newid;
    const int *lab;
    int n_lab;
    unsigned lab_classes;
    int last_lab;
    double gh;
    int RING_MAX;
    int n;
    int64_t nnt, nnoise;
    volatile int next; /* block ticket (threads) */
} nn_ctx_t;

static void nn_range(nn_ctx_t *c, int p0, int p1, int64_t *nnt_out, int64_t *nnoise_out) {
    const grid_t lg = *c->lgp;
    const float *xo = c->xo, *yo = c->yo, *zo = c->zo;
    const int *sem = c->sem, *ids = c->ids, *lab = c->lab;
    int *newid = c->newid;
    const int n_lab = c->n_lab, last_lab = c->last_lab, RING_MAX = c->RING_MAX;
    const unsigned lab_classes = c->lab_classes;
    const double gh = c->gh;
    int64_t nnt = 0, nnoise = 0;
    for (int p = p0; p < p1; p++) {
        if (ids[p] != -1) continue;
        nnoise++;
        if (!(lab_classes & (1u << sem[p]))) { /* fallback :287-300 */
            newid[p] = ids[last_lab];
            continue;
        }
        float px = xo[p], py = yo[p], pz = zo[p];
        int64_t cx, cy, cz;
        grid_cell(&lg, px, py, pz, &cx, &cy, &cz);
        float best = 0.f;
        int bestq = -1;
        int done = 0;
        for (int k = 0; k <= RING_MAX && !done; k++) {
            /* scan the shell of Chebyshev radius k, row by row */
            for (int64_t dz = -k; dz <= k; dz++)
                for (int64_t dy = -k; dy <= k; dy++) {
                    int full = (dz == -k || dz == k || dy == -k || dy == k);
                    for (int part = 0; part < (full ? 1 : (k == 0 ? 1 : 2)); part++) {
                        int64_t x0, x1;
                        if (full) {
                            x0 = cx - k;
                            x1 = cx + k;
                        } else {
                            x0 = x1 = (part == 0) ? cx - k : cx + k;
                        }
                        int b, e;
                        if (!grid_run(&lg, x0, x1, cy + dy, cz + dz, &b, &e)) continue;
                        for (int j = b; j < e; j++) {
                            int q = lg.ord[j];
                            if (sem[q] != sem[p]) continue;
                            float d = sqd(px, py, pz, xo[q], yo[q], zo[q]);
                            nnt++;
                            if (bestq < 0 || d < best || (d == best && q > bestq)) {
                                best = d;
                                bestq = q;
                            }
                        }
                    }
                }
            /* every unscanned point is farther than k*gh in true distance */
            double lim = (double)k * gh;
            if (bestq >= 0 && (double)best < lim * lim * (1.0 - 1e-5)) done = 1;
        }
        if (!done) { /* brute force over all labelled points (literal scan) */
            bestq = -1;
            for (int i = 0; i < n_lab; i++) {
                int q = lab[i];
                if (sem[q] != sem[p]) continue;
                float d = sqd(px, py, pz, xo[q], yo[q], zo[q]);
                nnt++;
                if (bestq < 0 || d <= best) {
                    best = d;
                    bestq = q;
                }
            }
        }
        newid[p] = ids[bestq];
    }
    *nnt_out = nnt;
    *nnoise_out = nnoise;
}

#include <pthread.h>
#include <unistd.h>
static void *nn_worker(void *arg) {
    nn_ctx_t *c = (nn_ctx_t *)arg;
    int64_t nnt = 0, nnoise = 0;
    for (;;) {
        int blk = __sync_fetch_and_add(&c->next, 1);
        long p0 = (long)blk * 1024;
        if (p0 >= c->n) break;
        int64_t a = 0, b = 0;
        nn_range(c, (int)p0, (int)(p0 + 1024 < c->n ? p0 + 1024 : c->n), &a, &b);
        nnt += a, nnoise += b;
    }
    __sync_fetch_and_add(&c->nnt, nnt);
    __sync_fetch_and_add(&c->nnoise, nnoise);
    return NULL;
}
static void nn_run(nn_ctx_t *c) {
    long nthreads = c->n > 200000 ? sysconf(_SC_NPROCESSORS_ONLN) : 1;
    if (nthreads > 64) nthreads = 64;
    if (nthreads <= 1) {
        nn_range(c, 0, c->n, &c->nnt, &c->nnoise);
        return;
    }
    pthread_t th[64];
    int started = 0;
    for (long t = 0; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, nn_worker, c) == 0) started++;
    if (started == 0) nn_worker(c);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

static int grid_segment(int n, const float *x, const float *y, const float *z, const float *xo,
                        const float *yo, const float *zo, const int *sem, const float *radius,
                        const int *min_pts, float para_f, int nv_flag, int *ids, int *den, int acc,
                        int *clt_sem_out, int *n_clt_sem, pb_oracle_stats *st) {
    /* the grid needs one cell edge >= every radius in use */
    float rmax = 0.f;
    for (int i = 0; i < n; i++) {
        float r = radius[sem[i] - 2];
        if (r > rmax) rmax = r;
    }
    double h = (double)rmax * 1.001;
    if (!(h > 0)) h = 1e-3;
    grid_t g;
    int rc = grid_build(&g, x, y, z, NULL, n, h);
    if (rc) return -rc;

    /* per cell: the 9 candidate runs in sorted positions */
    int *runs = (int *)malloc(sizeof(int) * 18 * (g.C > 0 ? g.C : 1));
    int *nruns = (int *)malloc(sizeof(int) * (g.C > 0 ? g.C : 1));
    int maxcand = 0;
    for (int c = 0; c < g.C; c++) {
        int64_t k = g.ckey[c];
        int64_t cx = k % g.nx, cy = (k / g.nx) % g.ny, cz = k / (g.nx * g.ny);
        int m = 0, tot = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++) {
                int b, e;
                if (grid_run(&g, cx - 1, cx + 1, cy + dy, cz + dz, &b, &e)) {
                    runs[18 * c + 2 * m] = b;
                    runs[18 * c + 2 * m + 1] = e;
                    tot += e - b;
                    m++;
                }
            }
        nruns[c] = m;
        if (tot > maxcand) maxcand = tot;
    }
    /* coordinates in sorted order for contiguous inner loops */
    float *sx = (float *)malloc(sizeof(float) * (n + 1));
    float *sy = (float *)malloc(sizeof(float) * (n + 1));
    float *sz = (float *)malloc(sizeof(float) * (n + 1));
    for (int i = 0; i < n; i++) {
        sx[i] = x[g.ord[i]];
        sy[i] = y[g.ord[i]];
        sz[i] = z[g.ord[i]];
    }
    /* A1 degree: count of pred-true candidates minus self (binary_cuda_functions.cu:85-88) */
    int64_t sumdeg = 0, tests = 0;
    for (int c = 0; c < g.C; c++) {
        for (int i = g.cstart[c]; i < g.cstart[c + 1]; i++) {
            int u = g.ord[i];
            float r = radius[sem[u] - 2];
            float r2 = r * r;
            float ux = sx[i], uy = sy[i], uz = sz[i];
            int ans = 0;
            for (int m = 0; m < nruns[c]; m++) {
                int b = runs[18 * c + 2 * m], e = runs[18 * c + 2 * m + 1];
                for (int j = b; j < e; j++) ans += (sqd(ux, uy, uz, sx[j], sy[j], sz[j]) <= r2);
                tests += e - b;
            }
            den[u] = ans - 1;
            sumdeg += ans - 1;
        }
    }
    /* A2 HP rule */
    int *member = (int *)malloc(sizeof(int) * (n + 1));
    int64_t nhp = 0;
    for (int u = 0; u < n; u++) {
        member[u] = (den[u] >= min_pts[sem[u] - 2]) ? 0 : 2;
        nhp += member[u] == 0;
    }
    if (st) {
        st->pair_tests += tests;
        st->sum_deg += sumdeg;
        st->n_hp += nhp;
    }
    /* position of each point in sorted order */
    int *pos = (int *)malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; i++) pos[g.ord[i]] = i;
    /* A3 clusters: seeds ascending; BFS expands only out of HPs; every visited same-class vertex is
     * (re)labelled (binary.cu:154-217).  Neighbours are enumerated on the fly with the predicate. */
    unsigned char *state = (unsigned char *)calloc(n + 1, 1); /* bit0 visited, bit1 queued (per BFS) */
    int *queue = (int *)malloc(sizeof(int) * (n + 1));
    int cluster = acc;
    for (int u = 0; u < n; u++) {
        if (!(ids[u] == -1 && member[u] == 0)) continue;
        int qh = 0, qt = 0;
        queue[qt++] = u;
        state[u] = 2;
        while (qh < qt) {
            int v = queue[qh++];
            state[v] |= 1;
            if (member[v] != 0) continue;
            int i = pos[v];
            int c = g.cell_of[i];
            /* the reference looks the radius up with sem indexed by SORTED position
             * (binary_cuda_functions.cu:35,110) — only well defined when every class present in
             * the segment has the same radius, which run() enforces */
            float r = radius[sem[v] - 2];
            float r2 = r * r;
            float vx = sx[i], vy = sy[i], vz = sz[i];
            for (int m = 0; m < nruns[c]; m++) {
                int b = runs[18 * c + 2 * m], e = runs[18 * c + 2 * m + 1];
                for (int j = b; j < e; j++) {
                    if (j == i) continue;
                    if (sqd(vx, vy, vz, sx[j], sy[j], sz[j]) <= r2) {
                        int w = g.ord[j];
                        if (!state[w]) {
                            state[w] = 2;
                            queue[qt++] = w;
                        }
                    }
                }
            }
        }
        int sem_cur = sem[u];
        for (int k = 0; k < qt; k++) {
            int v = queue[k];
            if (sem[v] == sem_cur) {
                ids[v] = cluster;
                if (member[v] != 0) member[v] = 1;
            }
            state[v] = 0;
        }
        cluster++;
    }
    free(state);
    free(queue);
    free(pos);
    free(member);
    free(runs);
    free(nruns);
    free(sx);
    free(sy);
    free(sz);
    grid_free(&g);

    /* A4 filter */
    int K = filter_segment(n, sem, ids, cluster - acc, acc, para_f, clt_sem_out, n_clt_sem);

    /* A5 LP assignment: exact 1-NN in ORIGINAL coordinates, ties -> largest index */
    if (nv_flag) {
        int n_lab = 0;
        int *lab = (int *)malloc(sizeof(int) * (n + 1));
        unsigned lab_classes = 0;
        for (int i = 0; i < n; i++)
            if (ids[i] != -1) {
                lab[n_lab++] = i;
                lab_classes |= 1u << sem[i];
            }
        if (n_lab < n && n_lab > 0) {
            const double gh = 0.06;
            const int RING_MAX = 5;
            grid_t lg;
            rc = grid_build(&lg, xo, yo, zo, lab, n_lab, gh);
            if (rc) {
                free(lab);
                return -rc;
            }
            int *newid = (int *)malloc(sizeof(int) * n);
            memcpy(newid, ids, sizeof(int) * n);
            int last_lab = lab[n_lab - 1];
            int64_t nnt = 0, nnoise = 0;
            nn_ctx_t nc = {&lg, xo, yo, zo, sem, ids, newid, lab, n_lab, lab_classes, last_lab, gh, RING_MAX, n, 0, 0, 0};
            nn_run(&nc);
            nnt = nc.nnt, nnoise = nc.nnoise;
            memcpy(ids, newid, sizeof(int) * n);
            free(newid);
            grid_free(&lg);
            if (st) {
                st->n_noise += nnoise;
                st->nn_tests += nnt;
            }
        }
        free(lab);
    }
    return K;
}

/* =============================================================================================
 * segment loop (cluster.cu:57-118)
 * =========================================================================================== */
static int run(int use_grid, const float *x, const float *y, const float *z, const float *xo,
               const float *yo, const float *zo, const int *sem, const int *seg_counts, int n_seg,
               const float *radius, const int *min_pts, float para_f, int nv_flag, int *cluster_id,
               int *cluster_num, int *den_queue, float *center, int *clt_sem, int *n_clusters_out,
               pb_oracle_stats *st) {
    int64_t total = 0;
    for (int b = 0; b < n_seg; b++) {
        if (seg_counts[b] < 0) return PB_ERR_ARG;
        total += seg_counts[b];
    }
    for (int64_t i = 0; i < total; i++)
        if (sem[i] < 2 || sem[i] > 19) return PB_ERR_ARG;
    {
        int64_t o = 0;
        for (int b = 0; b < n_seg; b++) { /* radius must be uniform over the classes of a segment */
            for (int i = 1; i < seg_counts[b]; i++)
                if (radius[sem[o + i] - 2] != radius[sem[o] - 2]) return PB_ERR_ARG;
            o += seg_counts[b];
        }
    }
    if (st) memset(st, 0, sizeof(*st));
    int start = 0, acc = 0, n_center = 0, n_clt_sem = 0;
    for (int b = 0; b < n_seg; b++) {
        int n = seg_counts[b];
        cluster_num[b] = 0;
        if (n == 0) continue;
        for (int i = 0; i < n; i++) cluster_id[start + i] = -1;
        int K;
        if (use_grid)
            K = grid_segment(n, x + start, y + start, z + start, xo + start, yo + start, zo + start,
                             sem + start, radius, min_pts, para_f, nv_flag, cluster_id + start,
                             den_queue + start, acc, clt_sem, &n_clt_sem, st);
        else
            K = literal_segment(n, x + start, y + start, z + start, xo + start, yo + start,
                                zo + start, sem + start, radius, min_pts, para_f, nv_flag,
                                cluster_id + start, den_queue + start, acc, clt_sem, &n_clt_sem, st);
        if (K < 0) return -K;
        cluster_num[b] = K;
        if (K > 0)
            centres_segment(n, x + start, y + start, z + start, cluster_id + start, K, acc, center,
                            &n_center);
        acc += K;
        start += n;
    }
    *n_clusters_out = acc;
    return PB_OK;
}

int pb_oracle_literal(const float *x, const float *y, const float *z, const float *xo, const float *yo,
                      const float *zo, const int *sem, const int *seg_counts, int n_seg,
                      const float *radius, const int *min_pts, float para_f, int nv_flag,
                      int *cluster_id, int *cluster_num, int *den_queue, float *center, int *clt_sem,
                      int *n_clusters_out, pb_oracle_stats *st) {
    return run(0, x, y, z, xo, yo, zo, sem, seg_counts, n_seg, radius, min_pts, para_f, nv_flag,
               cluster_id, cluster_num, den_queue, center, clt_sem, n_clusters_out, st);
}

int pb_oracle_grid(const float *x, const float *y, const float *z, const float *xo, const float *yo,
                   const float *zo, const int *sem, const int *seg_counts, int n_seg,
                   const float *radius, const int *min_pts, float para_f, int nv_flag, int *cluster_id,
                   int *cluster_num, int *den_queue, float *center, int *clt_sem, int *n_clusters_out,
                   pb_oracle_stats *st) {
    return run(1, x, y, z, xo, yo, zo, sem, seg_counts, n_seg, radius, min_pts, para_f, nv_flag,
               cluster_id, cluster_num, den_queue, center, clt_sem, n_clusters_out, st);
}

/* =================================================================================================
 * Mesh vertex normals — restatement of lib/PB_lib/src/normal/cal_normal.cu (the fourth op of the PB_lib
 * module, bound at PB_lib_api.cpp:10, called from lib/PB_lib/torch_io/pbnet_ops.py:163).
 *
 *   surface_normal_area (:43-76)   n = (B-A) x (C-A); "area" = |n|^2 / 2 (sic); n /= |n|
 *   vertex_normal       (:78-112)  per vertex, over the faces 0..num_face-1 IN ORDER that list it (once per face):
 *                                  s += n_f * area_f, a += area_f; a == 0 -> (0,0,1), else (s/a) / |s/a|
 *
 * Arithmetic as nvcc 12.9 contracts it for sm_100a (SASS of oracle/_ref): cross component = fma(u, v, -(w*t));
 * dot(n,n) = fma(z,z, fma(y,y, x*x)) but linalg_two_norm = sqrt(fma(z,z, fma(x,x, y*y))) (the middle product is the
 * plain multiply, as in the grouping predicate); accumulation s = fma(n, area, s); IEEE sqrt and division.
 * ================================================================================================= */
static inline float norm3(float x, float y, float z) { return sqrtf(fmaf(z, z, fmaf(x, x, y * y))); }

int pb_oracle_normals(const float *xyz, const int *face, int num_vtx, int num_face, float *out) {
    float *fn = (float *)malloc(sizeof(float) * 3 * (size_t)(num_face > 0 ? num_face : 1));
    float *fa = (float *)malloc(sizeof(float) * (size_t)(num_face > 0 ? num_face : 1));
    float *acc = (float *)calloc(4 * (size_t)(num_vtx > 0 ? num_vtx : 1), sizeof(float));
    if (!fn || !fa || !acc) return -1;
    for (int f = 0; f < num_face; f++) {
        int ia = face[3 * f], ib = face[3 * f + 1], ic = face[3 * f + 2];
        if (ia < 0 || ib < 0 || ic < 0 || ia >= num_vtx || ib >= num_vtx || ic >= num_vtx) { free(fn); free(fa); free(acc); return -2; }
        float ax = xyz[3 * ib] - xyz[3 * ia], ay = xyz[3 * ib + 1] - xyz[3 * ia + 1], az = xyz[3 * ib + 2] - xyz[3 * ia + 2];
        float bx = xyz[3 * ic] - xyz[3 * ia], by = xyz[3 * ic + 1] - xyz[3 * ia + 1], bz = xyz[3 * ic + 2] - xyz[3 * ia + 2];
        float nx = fmaf(ay, bz, -(az * by)), ny = fmaf(az, bx, -(ax * bz)), nz = fmaf(ax, by, -(ay * bx));
        fa[f] = fmaf(nz, nz, fmaf(ny, ny, nx * nx)) * 0.5f;
        float l = norm3(nx, ny, nz);
        fn[3 * f] = nx / l, fn[3 * f + 1] = ny / l, fn[3 * f + 2] = nz / l;
        /* faces are visited in ascending order, so per-vertex sums see them in the reference's loop order */
        int v[3] = {ia, ib, ic};
        for (int j = 0; j < 3; j++) {
            if ((j == 1 && v[1] == v[0]) || (j == 2 && (v[2] == v[0] || v[2] == v[1]))) continue; /* once per face */
            float *s = acc + 4 * (size_t)v[j];
            s[0] = fmaf(fn[3 * f], fa[f], s[0]);
            s[1] = fmaf(fn[3 * f + 1], fa[f], s[1]);
            s[2] = fmaf(fn[3 * f + 2], fa[f], s[2]);
            s[3] = s[3] + fa[f];
        }
    }
    for (int u = 0; u < num_vtx; u++) {
        const float *s = acc + 4 * (size_t)u;
        if (s[3] == 0.0f) {
            out[3 * u] = 0.f, out[3 * u + 1] = 0.f, out[3 * u + 2] = 1.f;
        } else {
            float x = s[0] / s[3], y = s[1] / s[3], z = s[2] / s[3];
            float l = norm3(x, y, z);
            out[3 * u] = x / l, out[3 * u + 1] = y / l, out[3 * u + 2] = z / l;
        }
    }
    free(fn); free(fa); free(acc);
    return 0;
}
