/* CPU restatement of scipy.optimize.linear_sum_assignment as used by the reference's
 * hungarian() (adapteacher/modeling/GModule/utils/hungarian.py:34,58-65).
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py) - never linked into the product library.
 *
 * SciPy is a third-party dependency (pinned scipy==1.7.3, requirements.txt:72); its C++ core
 * (rectangular_lsap.cpp, Crouse 2016 shortest augmenting path) is not under /root/reference.
 * This file restates the published algorithm and is checked against the
 * SciPy installed in this image (1.18.1) by tests/test_oracle_lap.py.
 *
 * ttdg_oracle_hungarian(): the full reference wrapper - negate the float32 score, convert to
 * float64, solve min-cost, write a 0/1 matrix.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Solve min-cost assignment for cost[nr][nc] (row-major, fp64).  Returns 0 on success.
 * rows_out/cols_out receive min(nr,nc) pairs sorted by row. */
int ttdg_oracle_lsap(int nr, int nc, const double *cost_in, int64_t *rows_out, int64_t *cols_out)
{
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    double *cost = (double *)malloc(sizeof(double) * (size_t)nr * nc);
    if (transpose) {
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < nc; j++) cost[(size_t)j * nr + i] = cost_in[(size_t)i * nc + j];
        int t = nr; nr = nc; nc = t;
    } else {
        memcpy(cost, cost_in, sizeof(double) * (size_t)nr * nc);
    }
    double *u = (double *)calloc(nr, sizeof(double));
    double *v = (double *)calloc(nc, sizeof(double));
    double *shortest = (double *)malloc(sizeof(double) * nc);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * nc);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * nr);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * nc);
    char *SR = (char *)malloc(nr), *SC = (char *)malloc(nc);
    int64_t *remaining = (int64_t *)malloc(sizeof(int64_t) * nc);
    for (int i = 0; i < nr; i++) col4row[i] = -1;
    for (int j = 0; j < nc; j++) row4col[j] = -1;
    int rc = 0;

    for (int cur = 0; cur < nr; cur++) {
        double minVal = 0;
        int num_remaining = nc;
        for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;   /* reverse fill */
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int j = 0; j < nc; j++) { shortest[j] = INFINITY; path[j] = -1; }
        int64_t sink = -1, i = cur;
        while (sink == -1) {
            int64_t index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < num_remaining; it++) {
                int64_t j = remaining[it];
                double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
                if (r < shortest[j]) { path[j] = i; shortest[j] = r; }
                if (shortest[j] < lowest || (shortest[j] == lowest && row4col[j] == -1)) {
                    lowest = shortest[j]; index = it;
                }
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -1; goto done; }              /* infeasible */
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        u[cur] += minVal;
        for (int r = 0; r < nr; r++)
            if (SR[r] && r != cur) u[r] += minVal - shortest[col4row[r]];
        for (int j = 0; j < nc; j++)
            if (SC[j]) v[j] -= minVal - shortest[j];
        int64_t j = sink;
        while (1) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    if (transpose) {
        /* pairs are (col4row[i], i); sort by first element = SciPy's argsort on the transposed case */
        int n = nr;
        int64_t *order = (int64_t *)malloc(sizeof(int64_t) * n);
        for (int k = 0; k < n; k++) order[k] = k;
        for (int a = 1; a < n; a++) {               /* insertion sort, stable */
            int64_t key = order[a]; int b = a - 1;
            while (b >= 0 && col4row[order[b]] > col4row[key]) { order[b + 1] = order[b]; b--; }
            order[b + 1] = key;
        }
        for (int k = 0; k < n; k++) { rows_out[k] = col4row[order[k]]; cols_out[k] = order[k]; }
        free(order);
    } else {
        for (int k = 0; k < nr; k++) { rows_out[k] = k; cols_out[k] = col4row[k]; }
    }
done:
    free(cost); free(u); free(v); free(shortest); free(path); free(col4row); free(row4col);
    free(SR); free(SC); free(remaining);
    return rc;
}

/* hungarian(s) for one n1 x n2 float32 score matrix: perm[row, col] = 1 on the max-weight assignment. */
int ttdg_oracle_hungarian(int n1, int n2, const float *s, float *perm)
{
    size_t n = (size_t)n1 * n2;
    double *cost = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (size_t k = 0; k < n; k++) cost[k] = (double)(s[k] * -1.0f);
    int m = n1 < n2 ? n1 : n2;
    int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * (m ? m : 1));
    int64_t *cols = (int64_t *)malloc(sizeof(int64_t) * (m ? m : 1));
    memset(perm, 0, sizeof(float) * n);
    int rc = ttdg_oracle_lsap(n1, n2, cost, rows, cols);
    if (rc == 0)
        for (int k = 0; k < m; k++) perm[(size_t)rows[k] * n2 + cols[k]] = 1.0f;
    free(cost); free(rows); free(cols);
    return rc;
}
