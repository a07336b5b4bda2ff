/* CPU oracle (plain C) for the Pasio DP and one flat sliding-window round.
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Restates, in the reference's exact floating-point operation order:
 *   - LogMarginalLikelyhood{Int,Real}AlphaComputer.all_suffixes_self_score
 *       /root/reference/src/pasio/log_marginal_likelyhood.py:105-115, :121-132
 *   - SquareSplitter.split_without_normalizations + collect_split_points
 *       /root/reference/src/pasio/splitters/square_splitter.py:67-109
 *   - one SlidingWindowReducer round with NotConstant/NotZero base reducers, in the
 *     flat formulation of SURVEY.md 7.4
 *       splitters/sliding_window_reducer.py:10-29, constants_reducer.py:5-21,
 *       dto/sliding_window.py:9-15
 * The transcendental values come from host tables built by numpy/scipy (the same
 * calls the reference makes, cached_log.py:9,36), passed in by the caller.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile);
 * -ffp-contract=off keeps mul and add un-fused like numpy's separate ufunc loops.
 * Parity pinning: checked against oracle/pasio_oracle.py and the golden vectors in
 * tests/test_oracle_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Exact DP over N candidates.
 * C[k]  int64 cumulative counts at candidate k (any common offset),
 * L[k]  int64 candidate positions (any common offset),
 * gtab  lgamma table: indexed by (alpha_int + C[j]-C[i]) when alpha_is_int, else the
 *       lgamma(k+alpha) table indexed by C[j]-C[i];  ltab[k] = log(k + beta).
 * The cumulative counts must be expressed relative to the same origin the reference
 * would use (window-local cumsum), because for real alpha (alpha + C[j]) rounds.
 * Returns 0, or -1 when a table is too short. */
int dp_oracle(const int64_t *C, const int64_t *L, int64_t N,
              const double *gtab, int64_t n_gtab, const double *ltab, int64_t n_ltab,
              int alpha_is_int, double alpha, double pen,
              double *P, int64_t *prev)
{
    int64_t alpha_i = alpha_is_int ? (int64_t)alpha : 0;
    P[0] = 0.0;
    prev[0] = 0;
    for (int64_t j = 1; j < N; ++j) {
        double best = 0.0;
        int64_t arg = -1;
        /* (alpha + cumsum[stop]) evaluated once per row, as numpy does (:109 / :126) */
        double shifted_j_real = alpha + (double)C[j];
        int64_t shifted_j_int = alpha_i + C[j];
        for (int64_t i = 0; i < j; ++i) {
            int64_t len = L[j] - L[i];
            int64_t idx;
            double s;
            if (alpha_is_int) {
                idx = shifted_j_int - C[i];
                s = (double)idx;
            } else {
                idx = C[j] - C[i];
                s = shifted_j_real - (double)C[i];
            }
            if (idx < 0 || idx >= n_gtab || len < 0 || len >= n_ltab) return -1;
            double sub = s * ltab[len];
            double self = gtab[idx] - sub;
            double t = self + P[i];
            /* np.argmax: first maximum; a NaN beats everything and the first NaN wins */
            if (arg < 0 || t > best || (t != t && best == best)) { best = t; arg = i; }
        }
        prev[j] = arg;
        P[j] = best + pen;
    }
    return 0;
}

/* One flat sliding-window round.
 * Cg[0..n]     int64 global cumulative counts (Cg[0]=0),
 * cp[0..n]     uint8 change-point flags: cp[p] = counts[p-1] != counts[p] (0 at p=0,n),
 * cand[0..m)   int64 current candidates (positions), cand[0]=0, cand[m-1]=n,
 * constraint   0 none, 1 zeros, 2 constants,
 * keep[0..n]   uint8 out: union of survivors (caller zeroes it); keep[0]=keep[n]=1 set here.
 * cells_out    number of DP (i,j) cells evaluated.
 * Returns 0, -1 table too short, -2 alloc failure. */
int round_oracle(const int64_t *Cg, const uint8_t *cp, int64_t n,
                 const int64_t *cand, int64_t m,
                 int64_t window_size, int64_t window_shift, int constraint,
                 const double *gtab, int64_t n_gtab, const double *ltab, int64_t n_ltab,
                 int alpha_is_int, double alpha, double pen,
                 uint8_t *keep, int64_t *cells_out)
{
    int64_t cap = window_size + 1;
    int64_t *C = (int64_t *)malloc(sizeof(int64_t) * cap);
    int64_t *L = (int64_t *)malloc(sizeof(int64_t) * cap);
    int64_t *prev = (int64_t *)malloc(sizeof(int64_t) * cap);
    double *P = (double *)malloc(sizeof(double) * cap);
    if (!C || !L || !prev || !P) { free(C); free(L); free(prev); free(P); return -2; }
    int64_t cells = 0;
    int rc = 0;
    keep[0] = 1;
    keep[n] = 1;
    for (int64_t st = 0; st < m - 1; st += window_shift) {
        int64_t en = st + window_size + 1;
        if (en > m) en = m;
        int64_t first = cand[st], last = cand[en - 1];
        int64_t k = 0;
        int all_zero = (Cg[last] - Cg[first]) == 0;
        for (int64_t q = st; q < en; ++q) {
            int64_t p = cand[q];
            int take;
            if (q == st || q == en - 1) take = 1;
            else if (constraint == 2) take = cp[p] != 0;
            else if (constraint == 1) take = !all_zero;
            else take = 1;
            if (take) {
                L[k] = p - first;            /* window-local positions  (sliding_window_reducer.py:15) */
                C[k] = Cg[p] - Cg[first];    /* window-local cumsum     (counts[start:stop] slice, :17) */
                ++k;
            }
        }
        rc = dp_oracle(C, L, k, gtab, n_gtab, ltab, n_ltab, alpha_is_int, alpha, pen, P, prev);
        if (rc) break;
        cells += k * (k - 1) / 2;
        int64_t q = k - 1;
        keep[first + L[q]] = 1;
        while (q != 0) { q = prev[q]; keep[first + L[q]] = 1; }
    }
    if (cells_out) *cells_out = cells;
    free(C); free(L); free(prev); free(P);
    return rc;
}

/* The same exact DP with the columns of every row split over OpenMP threads (rows stay sequential).
 * Every cell is computed by the same expressions as dp_oracle; the per-thread (max, first arg-max) of
 * consecutive column chunks are combined in ascending column order with a strict '>', so the result is
 * the first maximum exactly as in the single-threaded loop (NaN: first NaN wins, as np.argmax).
 * Used for the full-size comparisons (N = 200 000: 2e10 cells). */
int dp_oracle_mt(const int64_t *C, const int64_t *L, int64_t N,
                 const double *gtab, int64_t n_gtab, const double *ltab, int64_t n_ltab,
                 int alpha_is_int, double alpha, double pen,
                 double *P, int64_t *prev, int n_threads)
{
#ifndef _OPENMP
    (void)n_threads;
    return dp_oracle(C, L, N, gtab, n_gtab, ltab, n_ltab, alpha_is_int, alpha, pen, P, prev);
#else
    int64_t alpha_i = alpha_is_int ? (int64_t)alpha : 0;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    double tb[256];
    int64_t ta[256];
    int bad = 0;
    P[0] = 0.0;
    prev[0] = 0;
#pragma omp parallel num_threads(n_threads)
    {
        const int T = omp_get_num_threads(), me = omp_get_thread_num();
        for (int64_t j = 1; j < N; ++j) {
            const int64_t chunk = (j + T - 1) / T;
            const int64_t i0 = me * chunk, i1 = (i0 + chunk < j) ? i0 + chunk : j;
            double best = 0.0;
            int64_t arg = -1;
            const double shifted_j_real = alpha + (double)C[j];
            const int64_t shifted_j_int = alpha_i + C[j];
            for (int64_t i = i0; i < i1; ++i) {
                int64_t len = L[j] - L[i];
                int64_t idx;
                double s;
                if (alpha_is_int) {
                    idx = shifted_j_int - C[i];
                    s = (double)idx;
                } else {
                    idx = C[j] - C[i];
                    s = shifted_j_real - (double)C[i];
                }
                if (idx < 0 || idx >= n_gtab || len < 0 || len >= n_ltab) { bad = 1; break; }
                double sub = s * ltab[len];
                double self = gtab[idx] - sub;
                double t = self + P[i];
                if (arg < 0 || t > best || (t != t && best == best)) { best = t; arg = i; }
            }
            tb[me] = best;
            ta[me] = arg;
#pragma omp barrier
            if (me == 0) {
                double b = 0.0;
                int64_t a = -1;
                for (int t = 0; t < T; ++t) {
                    if (ta[t] < 0) continue;
                    if (a < 0 || tb[t] > b || (tb[t] != tb[t] && b == b)) { b = tb[t]; a = ta[t]; }
                }
                prev[j] = a;
                P[j] = b + pen;
            }
#pragma omp barrier
        }
    }
    return bad ? -1 : 0;
#endif
}

/* round_oracle with the windows of the round spread over OpenMP threads (windows are independent: each
 * only ORs its survivors into keep[]; a byte store of 1 by several threads is benign). */
int round_oracle_mt(const int64_t *Cg, const uint8_t *cp, int64_t n,
                    const int64_t *cand, int64_t m,
                    int64_t window_size, int64_t window_shift, int constraint,
                    const double *gtab, int64_t n_gtab, const double *ltab, int64_t n_ltab,
                    int alpha_is_int, double alpha, double pen,
                    uint8_t *keep, int64_t *cells_out, int n_threads)
{
    int64_t cap = window_size + 1;
    int64_t nwin = (m - 1 + window_shift - 1) / window_shift;
    int64_t cells = 0;
    int rc = 0;
    if (n_threads < 1) n_threads = 1;
    keep[0] = 1;
    keep[n] = 1;
#pragma omp parallel num_threads(n_threads) reduction(+ : cells)
    {
        int64_t *C = (int64_t *)malloc(sizeof(int64_t) * cap);
        int64_t *L = (int64_t *)malloc(sizeof(int64_t) * cap);
        int64_t *prev = (int64_t *)malloc(sizeof(int64_t) * cap);
        double *P = (double *)malloc(sizeof(double) * cap);
#pragma omp for schedule(dynamic, 4)
        for (int64_t w = 0; w < nwin; ++w) {
            if (!C || !L || !prev || !P) { rc = -2; continue; }
            int64_t st = w * window_shift;
            int64_t en = st + window_size + 1;
            if (en > m) en = m;
            int64_t first = cand[st], last = cand[en - 1];
            int64_t k = 0;
            int all_zero = (Cg[last] - Cg[first]) == 0;
            for (int64_t q = st; q < en; ++q) {
                int64_t p = cand[q];
                int take;
                if (q == st || q == en - 1) take = 1;
                else if (constraint == 2) take = cp[p] != 0;
                else if (constraint == 1) take = !all_zero;
                else take = 1;
                if (take) {
                    L[k] = p - first;
                    C[k] = Cg[p] - Cg[first];
                    ++k;
                }
            }
            int r = dp_oracle(C, L, k, gtab, n_gtab, ltab, n_ltab, alpha_is_int, alpha, pen, P, prev);
            if (r) { rc = r; continue; }
            cells += k * (k - 1) / 2;
            int64_t q = k - 1;
            keep[first + L[q]] = 1;
            while (q != 0) { q = prev[q]; keep[first + L[q]] = 1; }
        }
        free(C); free(L); free(prev); free(P);
    }
    if (cells_out) *cells_out = cells;
    return rc;
}
