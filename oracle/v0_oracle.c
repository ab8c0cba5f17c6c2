/*
 * oracle/v0_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * CPU restatement of the reference's V0 linear nearest-neighbour search and of the
 * north-star acceptance rule.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Parity pinning: the reference (sty-hhh/NNS-CUDA) ships no tests and no golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against the reference's own V0
 * code compiled from /root/reference/core.cu:11-54 into oracle/_ref/libv0_ref.so
 * (recipe: oracle/build_ref.sh).  tests/test_oracle.py checks restatement == _ref bit for
 * bit whenever _ref is present, and against tests/golden/ fixtures that were produced by
 * _ref (tests/golden/make_golden.py) everywhere else.
 *
 * Arithmetic: must be compiled with -ffp-contract=off so that sub, mul and add are each
 * rounded to FP32 exactly as the reference's host code does (README.md:20 builds host
 * code without FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Follows core.cu:31-52.  For every query i the index of the first reference point with the
 * smallest squared L2 distance; distance accumulated in FP32 over ascending t starting at 0
 * (core.cu:38-43); strict '>' keeps the first minimum (core.cu:44); start value
 * (INFINITY, 0) (core.cu:34-35), so NaN distances, all-INF distances and n == 0 give 0. */
void oracle_v0_search(int k, int m, int n, const float *s_points, const float *r_points,
                      int *out_idx)
{
    for (int i = 0; i < m; ++i) {
        const float *q = s_points + (size_t)i * (size_t)k;
        float best = INFINITY;
        int best_j = 0;
        for (int j = 0; j < n; ++j) {
            const float *r = r_points + (size_t)j * (size_t)k;
            float acc = 0.0f;
            for (int t = 0; t < k; ++t) {
                const float d = q[t] - r[t];
                acc += d * d;
            }
            if (best > acc) {
                best = acc;
                best_j = j;
            }
        }
        out_idx[i] = best_j;
    }
}

/* Same signature and ownership as the reference callback (core.cu:23-29, 31, 52):
 * *results is malloc'd here and freed by the caller with free(). */
void oracle_v0_cudaCall(int k, int m, int n, float *s_points, float *r_points, int **results)
{
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    oracle_v0_search(k, m, n, s_points, r_points, tmp);
    *results = tmp;
}

/* "V0 OpenMP" of BASELINE.md section 5: the serial V0 above invoked once per query chunk
 * from an OpenMP loop (V0 itself has no pragma).  Indices are identical to
 * the serial run because queries are independent.  Returns the thread count used. */
int oracle_v0_search_omp(int k, int m, int n, const float *s_points, const float *r_points,
                         int *out_idx, int chunk)
{
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    if (chunk <= 0) chunk = 8;
    const int nchunks = (m + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < nchunks; ++c) {
        const int lo = c * chunk;
        const int hi = lo + chunk < m ? lo + chunk : m;
        oracle_v0_search(k, hi - lo, n, s_points + (size_t)lo * (size_t)k, r_points,
                         out_idx + lo);
    }
    return threads;
}

/* K nearest neighbours: an EXTENSION of the reference (which returns one index, core.cu:52), so there is
 * nothing in the reference to pin it against beyond K = 1 (tests check K = 1 == V0 wherever a finite
 * distance exists).  Distances in V0's form and rounding (core.cu:38-43); order = (distance, index)
 * ascending; NaN / +INF distances are never reported; missing entries are index -1, distance +INF. */
void oracle_v0_topk(int k, int m, int n, int K, const float *s_points, const float *r_points,
                    int *out_idx, float *out_dist)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < m; ++i) {
        const float *q = s_points + (size_t)i * (size_t)k;
        int *bi = out_idx + (size_t)i * (size_t)K;
        float *bd = out_dist + (size_t)i * (size_t)K;
        for (int e = 0; e < K; ++e) { bi[e] = -1; bd[e] = INFINITY; }
        for (int j = 0; j < n; ++j) {
            const float *r = r_points + (size_t)j * (size_t)k;
            float acc = 0.0f;
            for (int t = 0; t < k; ++t) {
                const float d = q[t] - r[t];
                acc += d * d;
            }
            if (!(acc < INFINITY) || !(acc < bd[K - 1])) continue; /* ascending j: equal distance, higher index loses */
            int e = K - 1;
            while (e > 0 && bd[e - 1] > acc) { bd[e] = bd[e - 1]; bi[e] = bi[e - 1]; --e; }
            bd[e] = acc;
            bi[e] = j;
        }
    }
}

/* The reference's generator stream (main.cu:10-13): count values of rand() / double(RAND_MAX) from the
 * process-wide libc generator, continuing wherever srand()/rand() left it.  Lets tests reproduce the
 * inputs of the reference's whole shape table (one srand(1000), ten shapes in sequence) quickly. */
void oracle_libc_rand_fill(float *dst, long count)
{
    for (long i = 0; i < count; ++i) dst[i] = (float)(rand() / (double)RAND_MAX);
}

/* torchrun exports OMP_NUM_THREADS=1; the bench sets the team size it reports explicitly */
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double d64(int k, const float *q, const float *r)
{
    double a = 0.0;
    for (int t = 0; t < k; ++t) {
        const double d = (double)q[t] - (double)r[t];
        a += d * d;
    }
    return a;
}

/*
 * North-star acceptance rule (BASELINE.json north_star).
 * For each query, with d64(j) the squared distance in FP64 from the FP32 inputs,
 * d_min = min_j d64(j) and T = { j : d64(j) <= d_min * (1 + rel_tol) }:
 *   (1) the engine answer g must be in T                         -> else violation (bit 0)
 *   (2) if |T| == 1, g must equal that element (implied by 1) and must equal V0's answer v
 *       (v outside T is an "oracle anomaly", bit 2, never expected)
 *   (3) exact ties: there is no j < g with d64(j) == d64(g)      -> else violation (bit 1)
 * Queries whose FP64 distances are all NaN (NaN coordinates) must return 0 like V0.
 *
 * engine_idx and v0_idx are int32[m]; v0_idx may be NULL.  counts[0] = violations of (1),
 * counts[1] = violations of (3), counts[2] = oracle anomalies, counts[3] = answers that
 * differ from V0 but were accepted as near ties, counts[4] = exact matches with V0,
 * counts[5] = out-of-range indices.  Returns the total number of violations.
 */
long oracle_check_tie_rule(int k, int m, int n, const float *s_points, const float *r_points,
                           const int *engine_idx, const int *v0_idx, double rel_tol,
                           long *counts)
{
    long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : c0, c1, c2, c3, c4, c5)
    for (int i = 0; i < m; ++i) {
        const float *q = s_points + (size_t)i * (size_t)k;
        const int g = engine_idx[i];
        if (v0_idx && g == v0_idx[i]) c4++;
        if (n == 0) {
            if (g != 0) c0++;
            continue;
        }
        if (g < 0 || g >= n) {
            c5++;
            c0++;
            continue;
        }
        double dmin = INFINITY;
        for (int j = 0; j < n; ++j) {
            const double d = d64(k, q, r_points + (size_t)j * (size_t)k);
            if (d < dmin) dmin = d;
        }
        if (!(dmin < INFINITY)) { /* every distance is NaN or +INF: V0 answers 0 */
            if (g != 0) c0++;
            continue;
        }
        const double bound = dmin * (1.0 + rel_tol);
        const double dg = d64(k, q, r_points + (size_t)g * (size_t)k);
        if (!(dg <= bound)) {
            c0++;
            continue;
        }
        int lower_tie = 0;
        for (int j = 0; j < g; ++j) {
            if (d64(k, q, r_points + (size_t)j * (size_t)k) == dg) {
                lower_tie = 1;
                break;
            }
        }
        if (lower_tie) c1++;
        if (v0_idx) {
            const int v = v0_idx[i];
            const double dv = (v >= 0 && v < n) ? d64(k, q, r_points + (size_t)v * (size_t)k)
                                                : INFINITY;
            if (!(dv <= bound)) c2++;
            if (g != v) c3++;
        }
    }
    if (counts) {
        counts[0] = c0; counts[1] = c1; counts[2] = c2;
        counts[3] = c3; counts[4] = c4; counts[5] = c5;
    }
    return c0 + c1;
}
