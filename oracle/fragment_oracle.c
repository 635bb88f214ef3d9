/*
 * oracle/fragment_oracle.c -- CPU statement of this repo's OWN fragment-matching spec
 * (tvidz_b200/csrc/fragment.cu header; SURVEY.md Appendix B.4 proposed it).
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference advertises fragment detection
 * (README.md:5) but implements none (inspector/db.py:78-79), so there is no reference
 * behaviour to pin against; the only reference-anchored property is that the zero-offset,
 * zero-tolerance score equals find_duplicates' match_count (db.py:85-89) on tick-exact data,
 * which tests/ check against oracle/match_oracle.
 *
 * Written for clarity, not speed: scores use a plain existence test per query cut.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* ticks = llround(ts * hz), non-finite / out-of-int32 values dropped, sorted, unique */
static int canon_ticks(const double *ts, int64_t n, double hz, int64_t *out)
{
    int m = 0;
    for (int64_t i = 0; i < n; i++) {
        double v = ts[i] * hz;
        if (!isfinite(v) || fabs(v) > 5.0e8) continue;
        out[m++] = llround(v);
    }
    qsort(out, (size_t)m, sizeof(int64_t), cmp_i64);
    int w = 0;
    for (int i = 0; i < m; i++)
        if (w == 0 || out[i] != out[w - 1]) out[w++] = out[i];
    return w;
}

/* exists j with |x - C[j]| <= tol  (C sorted) */
static int near(const int64_t *C, int L, int64_t x, int64_t tol)
{
    int lo = 0, hi = L;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (C[mid] < x - tol) lo = mid + 1; else hi = mid;
    }
    return lo < L && C[lo] <= x + tol;
}

static int score_offset(const int64_t *C, int L, const int64_t *Q, int qn, int64_t d, int64_t tol)
{
    int s = 0;
    for (int i = 0; i < qn; i++) s += near(C, L, Q[i] + d, tol);
    return s;
}

static int better(int s, int64_t d, int bs, int64_t bd)
{
    if (s != bs) return s > bs;
    int64_t ad = d < 0 ? -d : d, ab = bd < 0 ? -bd : bd;
    if (ad != ab) return ad < ab;
    return d < bd;
}

/* For every row: best score and its offset (ticks).  score_out/delta_out: [n_rows]. */
void tvzo_fragment_rows(const double *ts, const int64_t *off, int64_t n_rows, const double *q, int qn,
                        double tick_hz, int tol, int tol_gap, int anchor, int zero_only,
                        int32_t *score_out, int64_t *delta_out, int n_threads)
{
    int64_t *Q = (int64_t *)malloc(sizeof(int64_t) * (size_t)(qn > 0 ? qn : 1));
    int nq = canon_ticks(q, qn, tick_hz, Q);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
#endif
    for (int64_t r = 0; r < n_rows; r++) {
        int64_t len = off[r + 1] - off[r];
        int64_t *C = (int64_t *)malloc(sizeof(int64_t) * (size_t)(len > 0 ? len : 1));
        int L = canon_ticks(ts + off[r], len, tick_hz, C);
        int bs = 0;
        int64_t bd = 0;
        if (nq > 0 && L > 0) {
            if (zero_only) {
                bs = score_offset(C, L, Q, nq, 0, tol);
            } else {
                /* candidates: `anchor` consecutive intervals of query and row agree within tol_gap */
                for (int i = 0; i + anchor < nq; i++)
                    for (int j = 0; j + anchor < L; j++) {
                        int agree = 1;
                        for (int a = 0; a < anchor && agree; a++) {
                            int64_t diff = (C[j + a + 1] - C[j + a]) - (Q[i + a + 1] - Q[i + a]);
                            if (diff > tol_gap || diff < -tol_gap) agree = 0;
                        }
                        if (!agree) continue;
                        int64_t d = C[j] - Q[i];
                        int s = score_offset(C, L, Q, nq, d, tol);
                        if (better(s, d, bs, bd)) { bs = s; bd = d; }
                    }
            }
        }
        score_out[r] = bs;
        delta_out[r] = bd;
        free(C);
    }
    free(Q);
}
