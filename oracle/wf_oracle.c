/*
 * oracle/wf_oracle.c — CPU restatement of the reference's weighted Wagner–Fischer path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rna-sequence-diff-patch_b200/ may import, link or
 * execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * the JSON files under tests/golden/, which were produced by importing and running the unmodified reference
 * (/root/reference/StringEditDistance.py, IRMethods.py) with tests/golden/make_golden.py.
 *
 * Citations are /root/reference file:line.  Symbols are 4-bit codes: the index of the letter in
 * IRMethods.py:13 (A G C U Y R W S K M D V H B N = 0..14); code 15 is a "matches only itself"
 * spare used for symbols that the reference never looks up in the table (SED:79-81).
 *
 * Exactness: all arithmetic is IEEE fp64 in the reference's order; compile WITHOUT -ffast-math
 * and with -ffp-contract=off (there are no multiplies in the recurrence, the flag is belt and braces).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    double ins;          /* C['insert']  SED:95,156 */
    double del;          /* C['delete']  SED:97,174 */
    double sub[16][16];  /* C['update'][src][dst] SED:87 ; diagonal never read (SED:79-81) */
} orc_costs;

/* SED:76-89 — substitution cost; equal symbols add (int) 0, i.e. leave the sum unchanged. */
static inline double orc_sub(const orc_costs *C, uint8_t a, uint8_t b) {
    return a == b ? 0.0 : C->sub[a][b];
}

/* SED:133-224 (values only) + IR:439: D[m][n].  Two rolling rows. */
double orc_distance(const uint8_t *a, int m, const uint8_t *b, int n, const orc_costs *C) {
    double *row = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    for (int j = 0; j <= n; ++j) row[j] = (double)j * C->ins;           /* SED:159  j*i_cost */
    for (int i = 1; i <= m; ++i) {
        double diag = row[0];
        row[0] = (double)i * C->del;                                    /* SED:177  i*d_cost */
        const uint8_t ai = a[i - 1];
        for (int j = 1; j <= n; ++j) {
            double c0 = row[j - 1] + C->ins;                            /* SED:95 */
            double c1 = row[j] + C->del;                                /* SED:97 */
            double c2 = diag + orc_sub(C, ai, b[j - 1]);                /* SED:99 */
            diag = row[j];
            double v = c0 < c1 ? c0 : c1;                               /* SED:106-107 min() */
            row[j] = c2 < v ? c2 : v;
        }
    }
    double d = row[n];
    free(row);
    return d;
}

/* Full matrix + 3-bit tie mask (bit0 INS, bit1 DEL, bit2 UPD) — SED:103-124,146-222.
 * D and mask are (m+1)*(n+1), row-major.  Border masks: row 0 -> INS, col 0 -> DEL, (0,0) -> 0. */
void orc_matrix(const uint8_t *a, int m, const uint8_t *b, int n, const orc_costs *C,
                double *D, uint8_t *mask) {
    const size_t W = (size_t)n + 1;
    D[0] = 0.0; mask[0] = 0;
    for (int j = 1; j <= n; ++j) { D[j] = (double)j * C->ins; mask[j] = 1; }
    for (int i = 1; i <= m; ++i) {
        D[i * W] = (double)i * C->del; mask[i * W] = 2;
        for (int j = 1; j <= n; ++j) {
            double c0 = D[i * W + j - 1] + C->ins;
            double c1 = D[(i - 1) * W + j] + C->del;
            double c2 = D[(i - 1) * W + j - 1] + orc_sub(C, a[i - 1], b[j - 1]);
            double v = c0 < c1 ? c0 : c1;
            v = c2 < v ? c2 : v;
            D[i * W + j] = v;
            mask[i * W + j] = (uint8_t)((c0 == v) | ((c1 == v) << 1) | ((c2 == v) << 2)); /* SED:109 */
        }
    }
}

/* Canonical script = create_paths(dp)[0] (SED:228-271 BFS order): min cost, then fewest edges,
 * then first of INS, DEL, UPD.  Forward pass keeps (cost, steps); 2-bit direction per cell.
 * Output: ops[k] in {0 INS, 1 DEL, 2 UPD}, origin -> sink, and the cell (i,j) each op ENTERS
 * (1-based matrix coordinates).  Returns the number of ops, or -1 on allocation failure.
 * generate_es (SED:284-323) then gives source.index = i-1, destination.index = j-1. */
int orc_canonical_script(const uint8_t *a, int m, const uint8_t *b, int n, const orc_costs *C,
                         uint8_t *ops, int32_t *oi, int32_t *oj, double *dist_out) {
    const size_t W = (size_t)n + 1;
    const size_t cells = ((size_t)m + 1) * W;
    uint8_t *dir = (uint8_t *)malloc((cells + 3) / 4);     /* 2 bit / cell */
    double *row = (double *)malloc(sizeof(double) * W);
    int32_t *steps = (int32_t *)malloc(sizeof(int32_t) * W);
    if (!dir || !row || !steps) { free(dir); free(row); free(steps); return -1; }
    memset(dir, 0, (cells + 3) / 4);
#define SETDIR(idx, d) dir[(idx) >> 2] |= (uint8_t)((d) << (((idx) & 3) * 2))
#define GETDIR(idx) ((dir[(idx) >> 2] >> (((idx) & 3) * 2)) & 3)
    for (int j = 0; j <= n; ++j) { row[j] = (double)j * C->ins; steps[j] = j; if (j) SETDIR((size_t)j, 0); }
    for (int i = 1; i <= m; ++i) {
        double diag = row[0]; int32_t sdiag = steps[0];
        row[0] = (double)i * C->del; steps[0] = i; SETDIR((size_t)i * W, 1);
        const uint8_t ai = a[i - 1];
        for (int j = 1; j <= n; ++j) {
            double c0 = row[j - 1] + C->ins;
            double c1 = row[j] + C->del;
            double c2 = diag + orc_sub(C, ai, b[j - 1]);
            int32_t s0 = steps[j - 1] + 1, s1 = steps[j] + 1, s2 = sdiag + 1;
            double v = c0 < c1 ? c0 : c1; v = c2 < v ? c2 : v;
            /* fewest edges among cost-tied predecessors, first in INS,DEL,UPD order */
            int best = -1; int32_t bs = 0;
            if (c0 == v) { best = 0; bs = s0; }
            if (c1 == v && (best < 0 || s1 < bs)) { best = 1; bs = s1; }
            if (c2 == v && (best < 0 || s2 < bs)) { best = 2; bs = s2; }
            diag = row[j]; sdiag = steps[j];
            row[j] = v; steps[j] = bs;
            SETDIR((size_t)i * W + j, best);
        }
    }
    if (dist_out) *dist_out = row[n];
    /* traceback sink -> origin, then reverse */
    int k = 0, i = m, j = n;
    while (i > 0 || j > 0) {
        int d = GETDIR((size_t)i * W + j);
        ops[k] = (uint8_t)d; oi[k] = i; oj[k] = j; ++k;
        if (d == 0) --j; else if (d == 1) --i; else { --i; --j; }
    }
    for (int x = 0, y = k - 1; x < y; ++x, --y) {
        uint8_t t = ops[x]; ops[x] = ops[y]; ops[y] = t;
        int32_t u = oi[x]; oi[x] = oi[y]; oi[y] = u;
        u = oj[x]; oj[x] = oj[y]; oj[y] = u;
    }
    free(dir); free(row); free(steps);
#undef SETDIR
#undef GETDIR
    return k;
}

/* ---- tiny pthread parallel-for (no libgomp dependency): dynamic chunks off an atomic counter ---- */
typedef void (*orc_body)(int64_t idx, void *arg);
typedef struct { orc_body body; void *arg; int64_t n, chunk; atomic_llong next; } orc_pf;
static void *orc_pf_worker(void *p) {
    orc_pf *pf = (orc_pf *)p;
    for (;;) {
        int64_t s = atomic_fetch_add(&pf->next, pf->chunk);
        if (s >= pf->n) break;
        int64_t e = s + pf->chunk < pf->n ? s + pf->chunk : pf->n;
        for (int64_t i = s; i < e; ++i) pf->body(i, pf->arg);
    }
    return NULL;
}
int orc_num_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void orc_parallel_for(int64_t n, int64_t chunk, int nthreads, orc_body body, void *arg) {
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > 256) nthreads = 256;
    orc_pf pf; pf.body = body; pf.arg = arg; pf.n = n; pf.chunk = chunk; atomic_init(&pf.next, 0);
    if (nthreads == 1 || n <= chunk) { orc_pf_worker(&pf); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t) if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
    orc_pf_worker(&pf);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* Batch distance over concatenated code arrays; a_off/b_off have n_pairs+1 entries (symbol offsets).
 * This is the CPU baseline bench.py reports (kind "port"). */
typedef struct { const uint8_t *a, *b; const int64_t *a_off, *b_off; const orc_costs *C; double *out;
                 int64_t max_ops; uint8_t *ops; int32_t *oi, *oj, *n_ops; } orc_batch_arg;
static void orc_dist_body(int64_t p, void *v) {
    orc_batch_arg *g = (orc_batch_arg *)v;
    g->out[p] = orc_distance(g->a + g->a_off[p], (int)(g->a_off[p + 1] - g->a_off[p]),
                             g->b + g->b_off[p], (int)(g->b_off[p + 1] - g->b_off[p]), g->C);
}
void orc_distance_batch(const uint8_t *a, const int64_t *a_off, const uint8_t *b, const int64_t *b_off,
                        int64_t n_pairs, const orc_costs *C, double *out, int nthreads) {
    orc_batch_arg g = {a, b, a_off, b_off, C, out, 0, NULL, NULL, NULL, NULL};
    orc_parallel_for(n_pairs, 8, nthreads, orc_dist_body, &g);
}

/* Batch canonical scripts; per-pair output slots of stride max_ops; n_ops[p] receives the count. */
static void orc_script_body(int64_t p, void *v) {
    orc_batch_arg *g = (orc_batch_arg *)v;
    g->n_ops[p] = orc_canonical_script(g->a + g->a_off[p], (int)(g->a_off[p + 1] - g->a_off[p]),
                                       g->b + g->b_off[p], (int)(g->b_off[p + 1] - g->b_off[p]), g->C,
                                       g->ops + p * g->max_ops, g->oi + p * g->max_ops,
                                       g->oj + p * g->max_ops, g->out + p);
}
void orc_script_batch(const uint8_t *a, const int64_t *a_off, const uint8_t *b, const int64_t *b_off,
                      int64_t n_pairs, const orc_costs *C, int64_t max_ops,
                      uint8_t *ops, int32_t *oi, int32_t *oj, int32_t *n_ops, double *dist, int nthreads) {
    orc_batch_arg g = {a, b, a_off, b_off, C, dist, max_ops, ops, oi, oj, n_ops};
    orc_parallel_for(n_pairs, 1, nthreads, orc_script_body, &g);
}

/* Closed-form patch of a generated script (verified against SED:380-457 in the
 * golden suite): out = dest chars of non-delete ops ++ x[len(src):]; error code 0 / 1 / -1 from
 * comparing x with src = source chars of non-insert ops (SED:389-399).
 * Script given as ops + entered cells (i,j) for strings a (source) / b (destination), with the
 * reference's negative-index wrap (SED:302-323): i==0 -> a[m-1], j==0 -> b[n-1].
 * Returns error code; *out_len receives the patched length (0 when -1). */
int orc_patch_closed(const uint8_t *ops, const int32_t *oi, const int32_t *oj, int n_ops,
                     const uint8_t *a, int m, const uint8_t *b, int n,
                     const uint8_t *x, int xlen, uint8_t *out, int *out_len) {
    int srclen = 0, same = 1;
    for (int k = 0; k < n_ops; ++k) if (ops[k] != 0) {
        uint8_t sc = a[oi[k] > 0 ? oi[k] - 1 : m - 1];
        if (srclen >= xlen || x[srclen] != sc) same = 0;
        ++srclen;
    }
    int code = (same && srclen == xlen) ? 0 : (xlen >= srclen ? 1 : -1);
    if (code < 0) { *out_len = 0; return -1; }
    int o = 0;
    for (int k = 0; k < n_ops; ++k) if (ops[k] != 1) out[o++] = b[oj[k] > 0 ? oj[k] - 1 : n - 1];
    for (int k = srclen; k < xlen; ++k) out[o++] = x[k];
    *out_len = o;
    return code;
}

/* Query vs database: IR:435-440 score = 1/(1+D[m][n]) with str1 = query, str2 = record (IR:470),
 * then performance.py:12-15 top-k = stable sort by score descending (ties keep record order).
 * db_off has n_db+1 symbol offsets.  scores_out (optional) gets all n_db scores.
 * topk_idx/topk_score get k entries (k <= n_db). */
typedef struct { const uint8_t *q; int qlen; const uint8_t *db; const int64_t *db_off; const orc_costs *C; double *sc; } orc_search_arg;
static void orc_search_body(int64_t r, void *v) {
    orc_search_arg *g = (orc_search_arg *)v;
    double d = orc_distance(g->q, g->qlen, g->db + g->db_off[r], (int)(g->db_off[r + 1] - g->db_off[r]), g->C);
    g->sc[r] = 1.0 / (1.0 + d);                                            /* IR:440 */
}
void orc_search_topk(const uint8_t *q, int qlen, const uint8_t *db, const int64_t *db_off, int64_t n_db,
                     const orc_costs *C, int k, int64_t *topk_idx, double *topk_score,
                     double *scores_out, int nthreads) {
    double *sc = scores_out ? scores_out : (double *)malloc(sizeof(double) * (size_t)(n_db > 0 ? n_db : 1));
    orc_search_arg g = {q, qlen, db, db_off, C, sc};
    orc_parallel_for(n_db, 256, nthreads, orc_search_body, &g);
    /* k passes of "best remaining, lowest index on ties" == prefix of a stable descending sort */
    int64_t *taken = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    for (int t = 0; t < k; ++t) {
        int64_t best = -1;
        for (int64_t r = 0; r < n_db; ++r) {
            int used = 0;
            for (int u = 0; u < t; ++u) if (taken[u] == r) { used = 1; break; }
            if (used) continue;
            if (best < 0 || sc[r] > sc[best]) best = r;
        }
        taken[t] = best; topk_idx[t] = best; topk_score[t] = best >= 0 ? sc[best] : 0.0;
    }
    free(taken);
    if (!scores_out) free(sc);
}
