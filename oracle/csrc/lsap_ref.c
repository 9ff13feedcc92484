/* Oracle (test-only): rectangular linear-sum assignment, CPU, double precision.
 *
 * Restates the algorithm behind scipy.optimize.linear_sum_assignment, which is
 * what the reference calls at model/utils/costTool/hung.py:28.  The arithmetic
 * lives in scipy's compiled _lsap module (third-party, pinned scipy==1.16.3 in
 * /root/reference/requirements.txt:6, not under /root/reference), i.e. the
 * modified Jonker-Volgenant shortest-augmenting-path method of D. F. Crouse,
 * "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016,
 * with no initialisation phase.  Behaviour restated here, including the details
 * that decide ties:
 *   - a tall matrix (rows > cols) is solved on its transpose;
 *   - the unscanned-column list starts in descending column order and a scanned
 *     column is removed by swapping the last list entry into its slot;
 *   - among equal tentative distances the scan keeps the first column in list
 *     order unless a later one is unassigned (an unassigned column ends the search);
 *   - the reduced cost is evaluated as ((minv + c) - u[i]) - v[j].
 * Pinned in tests against the installed scipy on random, gated, tied, integer and
 * rectangular matrices (bit-exact col4row), and against tests/golden/.
 *
 * Returns 0 on success, -1 if the matrix holds NaN/-inf, -2 if infeasible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int64_t shortest_path(int64_t nc, const double *c, double *u, double *v, int64_t *pred,
                             const int64_t *row_of_col, double *dist, int64_t start_row,
                             unsigned char *row_seen, unsigned char *col_seen, int64_t *todo,
                             double *out_min)
{
    double minv = 0.0;
    int64_t n_todo = nc, i = start_row, sink = -1;
    for (int64_t t = 0; t < nc; ++t) {
        todo[t] = nc - 1 - t;
        col_seen[t] = 0;
        dist[t] = INFINITY;
    }
    while (sink < 0) {
        int64_t best_t = -1;
        double best = INFINITY;
        row_seen[i] = 1;
        for (int64_t t = 0; t < n_todo; ++t) {
            int64_t j = todo[t];
            double r = minv + c[i * nc + j] - u[i] - v[j];
            if (r < dist[j]) {
                dist[j] = r;
                pred[j] = i;
            }
            if (dist[j] < best || (dist[j] == best && row_of_col[j] < 0)) {
                best = dist[j];
                best_t = t;
            }
        }
        minv = best;
        if (minv == INFINITY) return -1;
        int64_t j = todo[best_t];
        if (row_of_col[j] < 0) sink = j; else i = row_of_col[j];
        col_seen[j] = 1;
        todo[best_t] = todo[--n_todo];
    }
    *out_min = minv;
    return sink;
}

/* cost: row-major [nr][nc] doubles.  col_of_row: [nr] (-1 = unassigned, tall case);
 * row_of_col: [nc] (-1 = unassigned). */
int oracle_lsap_f64(int64_t nr, int64_t nc, const double *cost, int64_t *col_of_row,
                    int64_t *row_of_col_out)
{
    for (int64_t i = 0; i < nr; ++i) col_of_row[i] = -1;
    for (int64_t j = 0; j < nc; ++j) row_of_col_out[j] = -1;
    if (nr == 0 || nc == 0) return 0;
    for (int64_t k = 0; k < nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) return -1;

    int tall = nr > nc;
    int64_t R = tall ? nc : nr, C = tall ? nr : nc;
    double *work = NULL;
    const double *c = cost;
    if (tall) {
        work = (double *)malloc(sizeof(double) * (size_t)(R * C));
        for (int64_t i = 0; i < nr; ++i)
            for (int64_t j = 0; j < nc; ++j) work[j * C + i] = cost[i * nc + j];
        c = work;
    }
    double *u = (double *)calloc((size_t)R, sizeof(double));
    double *v = (double *)calloc((size_t)C, sizeof(double));
    double *dist = (double *)malloc(sizeof(double) * (size_t)C);
    int64_t *pred = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    int64_t *c4r = (int64_t *)malloc(sizeof(int64_t) * (size_t)R);
    int64_t *r4c = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    int64_t *todo = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
    unsigned char *rs = (unsigned char *)malloc((size_t)R);
    unsigned char *cs = (unsigned char *)malloc((size_t)C);
    for (int64_t i = 0; i < R; ++i) c4r[i] = -1;
    for (int64_t j = 0; j < C; ++j) { r4c[j] = -1; pred[j] = -1; }

    int rc = 0;
    for (int64_t cur = 0; cur < R; ++cur) {
        double minv;
        for (int64_t i = 0; i < R; ++i) rs[i] = 0;
        int64_t sink = shortest_path(C, c, u, v, pred, r4c, dist, cur, rs, cs, todo, &minv);
        if (sink < 0) { rc = -2; break; }
        u[cur] += minv;
        for (int64_t i = 0; i < R; ++i)
            if (rs[i] && i != cur) u[i] += minv - dist[c4r[i]];
        for (int64_t j = 0; j < C; ++j)
            if (cs[j]) v[j] -= minv - dist[j];
        int64_t j = sink;
        for (;;) {
            int64_t i = pred[j];
            r4c[j] = i;
            int64_t prev = c4r[i];
            c4r[i] = j;
            j = prev;
            if (i == cur) break;
        }
    }
    if (rc == 0) {
        if (tall) {           /* solved rows are original columns */
            for (int64_t i = 0; i < R; ++i) { row_of_col_out[i] = c4r[i]; col_of_row[c4r[i]] = i; }
        } else {
            for (int64_t i = 0; i < R; ++i) { col_of_row[i] = c4r[i]; row_of_col_out[c4r[i]] = i; }
        }
    }
    free(work); free(u); free(v); free(dist); free(pred); free(c4r); free(r4c); free(todo);
    free(rs); free(cs);
    return rc;
}
