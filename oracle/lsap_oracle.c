/* TEST INFRASTRUCTURE ONLY — CPU restatement of the linear-sum-assignment routine the
 * reference calls.
 *
 * The reference solves every assignment problem with
 *     scipy.optimize.linear_sum_assignment(A.cpu().numpy(), maximize=True)
 * (pleas/core/solvers.py:29-31; callers: pleas/methods/activation_matching.py:173,
 * pleas/methods/weight_matching.py:78).  SciPy is a third-party dependency that is not
 * vendored under /root/reference (pinned scipy==1.11.4, requirements.txt:160).  Its
 * published algorithm is the shortest-augmenting-path method of D. F. Crouse, "On
 * implementing 2D rectangular assignment algorithms", IEEE T-AES 52(4), 2016: rows are
 * inserted one at a time, each by a Dijkstra-style search over reduced costs with dual
 * variables u (rows) and v (columns), all arithmetic in float64, maximisation by negating
 * the cost matrix.  This file restates that algorithm for square problems, including
 * SciPy's two documented selection rules that decide the answer when optima tie:
 *   - the not-yet-scanned columns are kept in a list filled in REVERSE order and a scanned
 *     column is removed by overwriting it with the list's last element;
 *   - among columns whose tentative distance equals the minimum, an unassigned column
 *     (a sink) is preferred, later list positions winning among sinks, the earliest list
 *     position winning otherwise.
 * It is pinned against SciPy itself on tests/golden/lap_golden.npz (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (cc -O2 -shared -fPIC).  ABI:
 *   int lsap_oracle_solve(int64_t n, const float *cost, int maximize, int64_t *col4row,
 *                         double *objective)
 * returns 0 on success, 1 on an infeasible problem, 2 on NaN/-inf input.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int64_t shortest_path(int64_t n, const double *c, const double *u, const double *v,
                             int64_t *pred, const int64_t *row4col, double *dist, int64_t row,
                             unsigned char *row_seen, unsigned char *col_seen, int64_t *todo,
                             double *out_min)
{
    double min_val = 0.0;
    int64_t n_todo = n;
    for (int64_t t = 0; t < n; ++t) todo[t] = n - t - 1;
    for (int64_t k = 0; k < n; ++k) { row_seen[k] = 0; col_seen[k] = 0; dist[k] = INFINITY; }

    int64_t sink = -1;
    while (sink < 0) {
        int64_t best = -1;
        double lowest = INFINITY;
        row_seen[row] = 1;
        for (int64_t t = 0; t < n_todo; ++t) {
            int64_t j = todo[t];
            double r = min_val + c[row * n + j] - u[row] - v[j];
            if (r < dist[j]) { pred[j] = row; dist[j] = r; }
            if (dist[j] < lowest || (dist[j] == lowest && row4col[j] < 0)) {
                lowest = dist[j];
                best = t;
            }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        int64_t j = todo[best];
        if (row4col[j] < 0) sink = j; else row = row4col[j];
        col_seen[j] = 1;
        todo[best] = todo[--n_todo];
    }
    *out_min = min_val;
    return sink;
}

int lsap_oracle_solve(int64_t n, const float *cost, int maximize, int64_t *col4row_out,
                      double *objective)
{
    if (n <= 0) { if (objective) *objective = 0.0; return 0; }
    double *c = (double *)malloc(sizeof(double) * n * n);
    for (int64_t i = 0; i < n * n; ++i) {
        double x = (double)cost[i];
        c[i] = maximize ? -x : x;
        if (c[i] != c[i] || c[i] == -INFINITY) { free(c); return 2; }
    }
    double *u = (double *)calloc(n, sizeof(double));
    double *v = (double *)calloc(n, sizeof(double));
    double *dist = (double *)malloc(sizeof(double) * n);
    int64_t *pred = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *todo = (int64_t *)malloc(sizeof(int64_t) * n);
    unsigned char *row_seen = (unsigned char *)malloc(n);
    unsigned char *col_seen = (unsigned char *)malloc(n);
    for (int64_t k = 0; k < n; ++k) { pred[k] = -1; col4row[k] = -1; row4col[k] = -1; }

    int status = 0;
    for (int64_t cur = 0; cur < n; ++cur) {
        double min_val;
        int64_t sink = shortest_path(n, c, u, v, pred, row4col, dist, cur, row_seen, col_seen,
                                     todo, &min_val);
        if (sink < 0) { status = 1; break; }
        u[cur] += min_val;
        for (int64_t i = 0; i < n; ++i)
            if (row_seen[i] && i != cur) u[i] += min_val - dist[col4row[i]];
        for (int64_t j = 0; j < n; ++j)
            if (col_seen[j]) v[j] -= min_val - dist[j];
        int64_t j = sink;
        for (;;) {
            int64_t i = pred[j];
            row4col[j] = i;
            int64_t prev = col4row[i];
            col4row[i] = j;
            j = prev;
            if (i == cur) break;
        }
    }
    if (status == 0) {
        double obj = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            col4row_out[i] = col4row[i];
            obj += (double)cost[i * n + col4row[i]];
        }
        if (objective) *objective = obj;
    }
    free(c); free(u); free(v); free(dist); free(pred); free(col4row); free(row4col);
    free(todo); free(row_seen); free(col_seen);
    return status;
}
