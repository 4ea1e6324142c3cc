/*
 * oracle/lsap_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the rectangular linear-sum-assignment solver the reference calls
 * at losses/WireframeLoss.py:236, models/WireframeHungarianMatcher.py:71 and
 * models/HungarianMatcher.py:127 (`scipy.optimize.linear_sum_assignment`).
 *
 * The solver itself is a third-party dependency that is NOT vendored in /root/reference:
 * scipy (unpinned in requirements.txt; 1.18.1 in this image), compiled module
 * scipy/optimize/_lsap.  Its published algorithm is the "modified Jonker-Volgenant
 * algorithm with no initialization" of D. F. Crouse, "On implementing 2D rectangular
 * assignment algorithms", IEEE TAES 52(4):1679-1696, 2016: for every row in turn, grow a
 * Dijkstra-style shortest augmenting path over reduced costs c[i][j]-u[i]-v[j], update the
 * dual variables, flip the path.
 *
 * Parity is pinned by tests/test_lsap_oracle.py, which runs this file against the installed
 * scipy on >1e5 random, tied, rectangular, constant and infeasible matrices and demands
 * identical (row, col) index arrays.  The details that make the output *bit-identical*
 * (and not merely optimal) are marked [tie] below:
 *   [tie-1] the pool of unscanned columns is seeded in DESCENDING column order;
 *   [tie-2] a scanned column leaves the pool by moving the pool's last entry into its slot;
 *   [tie-3] among equal tentative distances a still-unassigned column wins over the
 *           incumbent (later pool slot wins), otherwise the earliest pool slot is kept;
 *   [tie-4] a tentative distance is only overwritten by a strictly smaller one;
 *   [tie-5] tall matrices (rows > cols) are solved transposed and reported sorted by row.
 * All arithmetic is IEEE double, in the order written (compile with -ffp-contract=off).
 *
 * Return codes: 0 ok, -1 infeasible (scipy: ValueError "cost matrix is infeasible"),
 *               -2 invalid entry, NaN or -inf (scipy: "matrix contains invalid numeric entries").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WFO_OK 0
#define WFO_INFEASIBLE (-1)
#define WFO_INVALID (-2)

typedef struct {
    double *u, *v, *dist;
    int64_t *pred, *col_of_row, *row_of_col, *pool;
    unsigned char *row_seen, *col_seen;
} wfo_ws;

static int ws_alloc(wfo_ws *w, int64_t nr, int64_t nc) {
    w->u = (double *)calloc((size_t)nr, sizeof(double));
    w->v = (double *)calloc((size_t)nc, sizeof(double));
    w->dist = (double *)malloc((size_t)nc * sizeof(double));
    w->pred = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->col_of_row = (int64_t *)malloc((size_t)nr * sizeof(int64_t));
    w->row_of_col = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->pool = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    w->row_seen = (unsigned char *)malloc((size_t)nr);
    w->col_seen = (unsigned char *)malloc((size_t)nc);
    return w->u && w->v && w->dist && w->pred && w->col_of_row && w->row_of_col && w->pool &&
           w->row_seen && w->col_seen;
}

static void ws_free(wfo_ws *w) {
    free(w->u); free(w->v); free(w->dist); free(w->pred); free(w->col_of_row);
    free(w->row_of_col); free(w->pool); free(w->row_seen); free(w->col_seen);
}

/* One Dijkstra sweep from `start_row`; returns the free column reached, or -1. */
static int64_t grow_path(const double *c, int64_t nr, int64_t nc, wfo_ws *w, int64_t start_row,
                         double *reach) {
    int64_t live = nc, sink = -1, row = start_row;
    double frontier = 0.0;
    for (int64_t s = 0; s < nc; ++s) w->pool[s] = nc - 1 - s;             /* [tie-1] */
    memset(w->row_seen, 0, (size_t)nr);
    memset(w->col_seen, 0, (size_t)nc);
    for (int64_t j = 0; j < nc; ++j) w->dist[j] = INFINITY;

    while (sink < 0) {
        int64_t best_slot = -1;
        double best = INFINITY;
        w->row_seen[row] = 1;
        for (int64_t s = 0; s < live; ++s) {
            int64_t j = w->pool[s];
            double cand = frontier + c[row * nc + j] - w->u[row] - w->v[j];
            if (cand < w->dist[j]) {                                       /* [tie-4] */
                w->dist[j] = cand;
                w->pred[j] = row;
            }
            if (w->dist[j] < best || (w->dist[j] == best && w->row_of_col[j] < 0)) { /* [tie-3] */
                best = w->dist[j];
                best_slot = s;
            }
        }
        frontier = best;
        if (frontier == INFINITY) return -1;
        {
            int64_t j = w->pool[best_slot];
            if (w->row_of_col[j] < 0) sink = j; else row = w->row_of_col[j];
            w->col_seen[j] = 1;
            w->pool[best_slot] = w->pool[--live];                          /* [tie-2] */
        }
    }
    *reach = frontier;
    return sink;
}

/*
 * cost: nr x nc row-major doubles.  On success writes min(nr,nc) pairs to rows[]/cols[]
 * (rows ascending), exactly the two arrays scipy returns.
 */
int wfo_lsap_f64(const double *cost, int64_t nr, int64_t nc, int64_t *rows, int64_t *cols) {
    double *tmp = NULL;
    int flipped = 0, rc = WFO_OK;
    wfo_ws w;
    if (nr == 0 || nc == 0) return WFO_OK;
    if (nc < nr) {                                                         /* [tie-5] */
        tmp = (double *)malloc((size_t)(nr * nc) * sizeof(double));
        if (!tmp) return -3;
        for (int64_t i = 0; i < nr; ++i)
            for (int64_t j = 0; j < nc; ++j) tmp[j * nr + i] = cost[i * nc + j];
        { int64_t t = nr; nr = nc; nc = t; }
        cost = tmp;
        flipped = 1;
    }
    for (int64_t k = 0; k < nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) { free(tmp); return WFO_INVALID; }
    if (!ws_alloc(&w, nr, nc)) { ws_free(&w); free(tmp); return -3; }
    for (int64_t i = 0; i < nr; ++i) w.col_of_row[i] = -1;
    for (int64_t j = 0; j < nc; ++j) { w.row_of_col[j] = -1; w.pred[j] = -1; }

    for (int64_t r = 0; r < nr && rc == WFO_OK; ++r) {
        double reach = 0.0;
        int64_t sink = grow_path(cost, nr, nc, &w, r, &reach);
        if (sink < 0) { rc = WFO_INFEASIBLE; break; }
        /* dual update: scanned rows rise, scanned columns fall, by their slack to `reach` */
        w.u[r] += reach;
        for (int64_t i = 0; i < nr; ++i)
            if (w.row_seen[i] && i != r) w.u[i] += reach - w.dist[w.col_of_row[i]];
        for (int64_t j = 0; j < nc; ++j)
            if (w.col_seen[j]) w.v[j] -= reach - w.dist[j];
        /* flip the alternating path back to row r */
        for (int64_t j = sink;;) {
            int64_t i = w.pred[j], prev = w.col_of_row[i];
            w.row_of_col[j] = i;
            w.col_of_row[i] = j;
            j = prev;
            if (i == r) break;
        }
    }
    if (rc == WFO_OK) {
        if (!flipped) {
            for (int64_t i = 0; i < nr; ++i) { rows[i] = i; cols[i] = w.col_of_row[i]; }
        } else {
            /* solved on the transpose: w.col_of_row maps original column -> original row.
               Report sorted by original row (a stable counting pass: rows are distinct). */
            int64_t k = 0;
            for (int64_t orow = 0; orow < nc; ++orow)
                if (w.row_of_col[orow] >= 0) { rows[k] = orow; cols[k] = w.row_of_col[orow]; ++k; }
        }
    }
    ws_free(&w);
    free(tmp);
    return rc;
}

/* float32 entry point: the reference hands scipy a float32 ndarray
   (losses/WireframeLoss.py:235 `.cpu().numpy()`), which scipy widens to double exactly. */
int wfo_lsap_f32(const float *cost, int64_t nr, int64_t nc, int64_t *rows, int64_t *cols) {
    double *d;
    int rc;
    if (nr == 0 || nc == 0) return WFO_OK;
    d = (double *)malloc((size_t)(nr * nc) * sizeof(double));
    if (!d) return -3;
    for (int64_t k = 0; k < nr * nc; ++k) d[k] = (double)cost[k];
    rc = wfo_lsap_f64(d, nr, nc, rows, cols);
    free(d);
    return rc;
}

/*
 * Cost matrix of the loss's inline matcher, losses/WireframeLoss.py:142,211-224:
 *   real columns  j <  count : |p-t|_1 (torch.cdist p=1, terms added in x,y,z order) + |e-1|
 *   dummy columns j >= count : e
 * float32 arithmetic, one rounding per operation, as the reference's ATen CPU ops do.
 * out is V x max(V,count) when count<=V (square V x V), else V x count.
 */
void wfo_loss_cost_f32(const float *pred_v, const float *pred_e, const float *tgt_v, int64_t V,
                       int64_t count, float *out) {
    int64_t ncols = count > V ? count : V;
    for (int64_t i = 0; i < V; ++i) {
        float e = pred_e[i];
        volatile float pen = fabsf(e - 1.0f);
        for (int64_t j = 0; j < ncols; ++j) {
            if (j < count) {
                volatile float d = 0.0f;
                for (int k = 0; k < 3; ++k) {
                    volatile float a = fabsf(pred_v[i * 3 + k] - tgt_v[j * 3 + k]);
                    d = d + a;
                }
                out[i * ncols + j] = d + pen;
            } else {
                out[i * ncols + j] = e;
            }
        }
    }
}
