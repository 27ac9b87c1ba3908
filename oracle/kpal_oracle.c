/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of kPAL's hot path.
 *
 * Purpose: (1) a second, independent checker for the CUDA path at sizes the
 * NumPy restatement (oracle/kpal_oracle.py) finishes slowly on, and (2) the
 * CPU baseline that bench.py times on the GPU box's host cores
 * ("cpu_baseline", "--impl reference": kind "port").  It is never linked into
 * or called from the product library.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks every entry point
 * against the golden vectors under tests/golden/ (generated from the
 * unmodified reference, tests/golden/make_golden.py) and against
 * oracle/kpal_oracle.py, which is itself pinned to the reference.
 *
 * file:line citations are relative to the reference tree (LUMC/kPAL).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* kpal/klib.py:43-48 : A/a 0, C/c 1, G/g 2, T/t 3 ; everything else splits
 * (regex [^AaCcGgTt], kpal/klib.py:152). */
static inline int code_of(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/*
 * Count windows that START in [begin, end) of text[0..n).  A window is
 * counted iff all of its k bytes are ACGTacgt -- the rolling form of
 * kpal/klib.py:154-168 (records are joined by any non-ACGT byte, so the
 * no-window-across-records rule of klib.py:154 is the same predicate).
 */
static void count_range(const unsigned char *text, size_t n, size_t begin,
                        size_t end, int k, int64_t *counts, int atomic)
{
    const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    size_t stop = end + (size_t)k - 1;
    if (stop > n) stop = n;
    uint64_t binary = 0;
    size_t run = 0;
    for (size_t p = begin; p < stop; ++p) {
        int c = code_of(text[p]);
        if (c < 0) { run = 0; binary = 0; continue; }
        binary = ((binary << 2) | (uint64_t)c) & mask;      /* klib.py:166 */
        if (++run >= (size_t)k) {                           /* klib.py:168 */
            if (atomic)
                __atomic_fetch_add(&counts[binary], 1, __ATOMIC_RELAXED);
            else
                counts[binary] += 1;
        }
    }
}

/* Single-threaded count over a byte buffer; counts must hold 4^k zeros or
 * running totals (it is accumulated into). */
void oracle_count(const unsigned char *text, size_t n, int k, int64_t *counts)
{
    count_range(text, n, 0, n, k, counts, 0);
}

/* Same result with `threads` host threads (byte ranges; a window belongs to
 * the range its first byte lies in). */
void oracle_count_mt(const unsigned char *text, size_t n, int k,
                     int64_t *counts, int threads)
{
    if (threads <= 1) { oracle_count(text, n, k, counts); return; }
#ifdef _OPENMP
    const size_t number = (size_t)1 << (2 * k);
    const int private_tables = (number * sizeof(int64_t) * (size_t)threads)
                               <= ((size_t)1 << 30);
    int64_t **mine = (int64_t **)calloc((size_t)threads, sizeof(int64_t *));
    #pragma omp parallel num_threads(threads)
    {
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
        size_t b = n / (size_t)nt * (size_t)t;
        size_t e = (t == nt - 1) ? n : n / (size_t)nt * (size_t)(t + 1);
        if (private_tables) {
            mine[t] = (int64_t *)calloc(number, sizeof(int64_t));
            count_range(text, n, b, e, k, mine[t], 0);
            #pragma omp barrier
            #pragma omp for schedule(static)
            for (long i = 0; i < (long)number; ++i) {
                int64_t s = 0;
                for (int u = 0; u < nt; ++u) if (mine[u]) s += mine[u][i];
                counts[i] += s;
            }
            free(mine[t]);
        } else {
            count_range(text, n, b, e, k, counts, 1);
        }
    }
    free(mine);
#else
    oracle_count(text, n, k, counts);
#endif
}

/* kpal/klib.py:394-412 */
static inline uint64_t reverse_complement(uint64_t number, int k)
{
    number = ~number;
    uint64_t result = 0;
    for (int i = 0; i < k; ++i) {
        result = (result << 2) | (number & 3);
        number >>= 2;
    }
    return result;
}

uint64_t oracle_reverse_complement(uint64_t number, int k)
{
    return reverse_complement(number, k);
}

/* kpal/klib.py:285-298, in place. */
void oracle_balance(int64_t *counts, int k)
{
    const uint64_t number = 1ULL << (2 * k);
    for (uint64_t i = 0; i < number; ++i) {
        uint64_t i_rc = reverse_complement(i, k);
        if (i < i_rc) {
            int64_t temp = counts[i];
            counts[i] += counts[i_rc];
            counts[i_rc] += temp;
        } else if (i == i_rc) {
            counts[i] += counts[i];
        }
    }
}

/* metric ids shared with include/kpal_b200.h */
enum { METRIC_MULTISET = 0, METRIC_EUCLIDEAN = 1, METRIC_COSINE = 2 };
enum { PAIRWISE_PROD = 0, PAIRWISE_SUM = 1 };

/*
 * kpal/kdistlib.py:149-161 + kpal/metrics.py:49-86,101-147,159-162 for one
 * pair of (already balanced, if requested) int64 vectors.  Compensated
 * summation (the reference uses NumPy pairwise summation; both are within a
 * few ulp of the exact sum).
 */
double oracle_distance(const int64_t *left, const int64_t *right, size_t n,
                       int do_scale, int down, int metric, int pairwise)
{
    double ls = 1.0, rs = 1.0;
    if (do_scale) {                               /* metrics.py:61-72 */
        int64_t lsum = 0, rsum = 0;
        for (size_t i = 0; i < n; ++i) { lsum += left[i]; rsum += right[i]; }
        if (lsum < rsum) ls = (double)rsum / (double)lsum;
        else             rs = (double)lsum / (double)rsum;
        if (down) {                               /* metrics.py:84-86 */
            double f = ls > rs ? ls : rs;
            ls = ls / f; rs = rs / f;
        }
    }
    double sum = 0.0, comp = 0.0;                 /* Kahan */
    double ll = 0.0, rr = 0.0;
    size_t nz = 0;
    for (size_t i = 0; i < n; ++i) {
        double x = do_scale ? (double)left[i] * ls : (double)left[i];
        double y = do_scale ? (double)right[i] * rs : (double)right[i];
        double term;
        if (metric == METRIC_MULTISET) {
            if (!(x != 0.0 || y != 0.0)) continue;   /* metrics.py:121 (nan counts as set) */
            ++nz;
            if (pairwise == PAIRWISE_PROD)
                term = fabs(x - y) / ((x + 1.0) * (y + 1.0));
            else
                term = fabs(x - y) / (x + y + 1.0);
        } else if (metric == METRIC_EUCLIDEAN) {
            term = (x - y) * (x - y);
        } else {
            term = x * y;
            ll += x * x; rr += y * y;
        }
        double t = term - comp;
        double s = sum + t;
        comp = (s - sum) - t;
        sum = s;
    }
    if (metric == METRIC_MULTISET) return sum / (double)(nz + 1);
    if (metric == METRIC_EUCLIDEAN) return sqrt(sum);
    return sum / (sqrt(ll) * sqrt(rr));
}

/*
 * kpal/kdistlib.py:179-184: out[i*N + j] = d(p_i, p_j) for j < i (strict
 * lower triangle; everything else left untouched).  Profiles are rows of a
 * row-major [N][n] int64 array.  do_balance balances a private copy of each
 * profile once (identical to balancing per pair, kdistlib.py:139-141, since
 * balance is a per-profile operation).
 */
void oracle_distance_matrix(const int64_t *profiles, size_t N, size_t n,
                            int k, int do_balance, int do_scale, int down,
                            int metric, int pairwise, double *out,
                            int threads)
{
    const int64_t *src = profiles;
    int64_t *bal = NULL;
    if (do_balance) {
        bal = (int64_t *)malloc(N * n * sizeof(int64_t));
        memcpy(bal, profiles, N * n * sizeof(int64_t));
        #pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic)
        for (long i = 0; i < (long)N; ++i) oracle_balance(bal + (size_t)i * n, k);
        src = bal;
    }
    const long pairs = (long)(N * (N - 1) / 2);
    #pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 4)
    for (long p = 0; p < pairs; ++p) {
        /* p -> (i, j), j < i */
        long i = (long)((1.0 + sqrt(1.0 + 8.0 * (double)p)) / 2.0);
        while (i * (i - 1) / 2 > p) --i;
        while ((i + 1) * i / 2 <= p) ++i;
        long j = p - i * (i - 1) / 2;
        out[(size_t)i * N + (size_t)j] =
            oracle_distance(src + (size_t)i * n, src + (size_t)j * n, n,
                            do_scale, down, metric, pairwise);
    }
    free(bal);
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
