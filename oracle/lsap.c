/* CPU oracle: rectangular linear sum assignment -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates the algorithm behind scipy.optimize.linear_sum_assignment (the call the reference makes
 * at models/matcher.py:86; scipy is an un-vendored, unpinned dependency -- 1.18.1 in this image):
 * the shortest-augmenting-path method of Crouse, "On implementing 2D rectangular assignment
 * algorithms", IEEE TAES 2016, with scipy's scan order and tie-breaking.
 * Pinned against the installed scipy by tests/test_lsap_oracle.py (random, tie-heavy, both
 * orientations).  The CUDA kernel in spe_b200/csrc/lsap.cu must return identical indices.
 *
 *   cost : [nr, nc] row-major float32 (cast to double, as scipy does)
 *   out  : row_ind[k], col_ind[k], k = min(nr, nc); returns k, or -1 on infeasible / bad input
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int solve_rows_le_cols(int nr, int nc, const double *c, int64_t *col4row)
{
    double *u = calloc(nr, sizeof(double)), *v = calloc(nc, sizeof(double));
    double *spc = malloc(nc * sizeof(double));
    int *path = malloc(nc * sizeof(int)), *row4col = malloc(nc * sizeof(int));
    int *remaining = malloc(nc * sizeof(int));
    char *SR = malloc(nr), *SC = malloc(nc);
    int ok = 1;
    for (int j = 0; j < nc; ++j) row4col[j] = -1;
    for (int i = 0; i < nr; ++i) col4row[i] = -1;

    for (int cur = 0; cur < nr && ok; ++cur) {
        double min_val = 0.0;
        int i = cur, sink = -1, num_remaining = nc;
        for (int t = 0; t < nc; ++t) { remaining[t] = nc - 1 - t; spc[t] = INFINITY; SC[t] = 0; path[t] = -1; }
        for (int t = 0; t < nr; ++t) SR[t] = 0;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                int j = remaining[it];
                double r = min_val + c[(size_t)i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            min_val = lowest;
            if (min_val == INFINITY) { ok = 0; break; }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (!ok) break;
        u[cur] += min_val;
        for (int t = 0; t < nr; ++t)
            if (SR[t] && t != cur) u[t] += min_val - spc[col4row[t]];
        for (int j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= min_val - spc[j];
        int j = sink;
        for (;;) {
            int r = path[j];
            row4col[j] = r;
            int64_t tmp = col4row[r]; col4row[r] = j; j = (int)tmp;
            if (r == cur) break;
        }
    }
    free(u); free(v); free(spc); free(path); free(row4col); free(remaining); free(SR); free(SC);
    return ok;
}

int spe_oracle_lsap(int nr, int nc, const float *cost, int64_t *row_ind, int64_t *col_ind)
{
    if (nr < 0 || nc < 0) return -1;
    if (nr == 0 || nc == 0) return 0;
    const int transpose = nc < nr;
    const int R = transpose ? nc : nr, C = transpose ? nr : nc;
    double *c = malloc((size_t)R * C * sizeof(double));
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) {
            double x = (double)cost[(size_t)i * nc + j];
            if (isnan(x) || x == -INFINITY) { free(c); return -1; }
            if (transpose) c[(size_t)j * C + i] = x; else c[(size_t)i * C + j] = x;
        }
    int64_t *col4row = malloc(R * sizeof(int64_t));
    int ok = solve_rows_le_cols(R, C, c, col4row);
    if (ok) {
        if (!transpose) {
            for (int i = 0; i < R; ++i) { row_ind[i] = i; col_ind[i] = col4row[i]; }
        } else {
            /* col4row[j] = original row assigned to original column j; emit sorted by original row */
            int64_t *inv = malloc(C * sizeof(int64_t));
            for (int t = 0; t < C; ++t) inv[t] = -1;
            for (int j = 0; j < R; ++j) inv[col4row[j]] = j;
            int k = 0;
            for (int t = 0; t < C; ++t)
                if (inv[t] >= 0) { row_ind[k] = t; col_ind[k] = inv[t]; ++k; }
            free(inv);
        }
    }
    free(c); free(col4row);
    return ok ? R : -1;
}
