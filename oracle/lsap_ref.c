/* CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the rectangular linear-sum-assignment solver that the reference calls
 * at /root/reference/ModelComponents/losses_and_metrics.py:242
 * (scipy.optimize.linear_sum_assignment; third-party, not vendored under /root/reference; the
 * oracle of record is the installed scipy 1.18.1 and tests/test_oracle_lsap.py pins this file
 * against it bit-for-bit, ties and +inf entries included).
 *
 * Algorithm: shortest augmenting path with float64 duals (Crouse 2016), columns scanned through
 * a swap-removed `remaining` list that is filled in REVERSE order; the arg-min prefers, among
 * equal shortest-path costs, a column that is still unassigned, later scan positions winning.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LSAP_OK 0
#define LSAP_INFEASIBLE (-1)
#define LSAP_INVALID (-2)

/* cost: float32 row-major [nr, ld] (first nc columns used).  a/b: output pairs (min(nr,nc) of
 * them), rows ascending like scipy returns them.  Returns number of pairs or a negative status. */
int lsap_ref_f32(const float *cost, int nr, int nc, int ld, int64_t *a, int64_t *b)
{
    if (nr <= 0 || nc <= 0) return 0;
    const int transpose = nc < nr;
    const int R = transpose ? nc : nr;      /* solver rows  */
    const int C = transpose ? nr : nc;      /* solver cols  */
    double *c = (double *)malloc(sizeof(double) * (size_t)R * C);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double x = transpose ? (double)cost[(size_t)j * ld + i] : (double)cost[(size_t)i * ld + j];
            if (x != x || x == -INFINITY) { free(c); return LSAP_INVALID; }
            c[(size_t)i * C + j] = x;
        }
    double *u = (double *)calloc(R, sizeof(double)), *v = (double *)calloc(C, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * C);
    int *path = (int *)malloc(sizeof(int) * C), *col4row = (int *)malloc(sizeof(int) * R);
    int *row4col = (int *)malloc(sizeof(int) * C), *remaining = (int *)malloc(sizeof(int) * C);
    char *SR = (char *)malloc(R), *SC = (char *)malloc(C);
    for (int i = 0; i < R; ++i) col4row[i] = -1;
    for (int j = 0; j < C; ++j) { row4col[j] = -1; path[j] = -1; }
    int status = LSAP_OK;

    for (int cur = 0; cur < R && status == LSAP_OK; ++cur) {
        double minVal = 0.0;
        int i = cur, nrem = C, sink = -1;
        for (int it = 0; it < C; ++it) remaining[it] = C - it - 1;
        memset(SR, 0, R); memset(SC, 0, C);
        for (int j = 0; j < C; ++j) spc[j] = INFINITY;
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int it = 0; it < nrem; ++it) {
                int j = remaining[it];
                double r = minVal + c[(size_t)i * C + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INFINITY) { status = LSAP_INFEASIBLE; break; }
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--nrem];
        }
        if (status != LSAP_OK) break;
        u[cur] += minVal;
        for (int r = 0; r < R; ++r) if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int j = 0; j < C; ++j) if (SC[j]) v[j] -= minVal - spc[j];
        int j = sink;
        for (;;) {
            int r = path[j];
            row4col[j] = r;
            int t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    int n = 0;
    if (status == LSAP_OK) {
        if (transpose) {            /* solver rows are original columns: emit sorted by original row */
            for (int j = 0; j < C; ++j) if (row4col[j] != -1) { a[n] = j; b[n] = row4col[j]; ++n; }
        } else {
            for (int i = 0; i < R; ++i) { a[n] = i; b[n] = col4row[i]; ++n; }
        }
    }
    free(c); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
    return status == LSAP_OK ? n : status;
}

/* The reference's per-batch loop (losses_and_metrics.py:234-245): mask[b, row, col] = 1. */
int lsap_ref_batch_mask(const float *cost, const int32_t *num_objects, int B, int T, int Q, float *mask)
{
    int64_t *a = (int64_t *)malloc(sizeof(int64_t) * (T > Q ? T : Q));
    int64_t *b = (int64_t *)malloc(sizeof(int64_t) * (T > Q ? T : Q));
    int rc = 0;
    memset(mask, 0, sizeof(float) * (size_t)B * T * Q);
    for (int i = 0; i < B; ++i) {
        int n = num_objects[i]; if (n > T) n = T; if (n < 0) n = 0;
        int k = lsap_ref_f32(cost + (size_t)i * T * Q, n, Q, Q, a, b);
        if (k < 0) { rc = k; break; }
        for (int p = 0; p < k; ++p) mask[(size_t)i * T * Q + a[p] * Q + b[p]] = 1.0f;
    }
    free(a); free(b);
    return rc;
}
