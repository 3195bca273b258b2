/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of Cpeeler.nj_np
 * (dodonaphy/cython/Cpeeler.pyx:15-125) -- same operation order as oracle/nj.py:nj_hard, fast enough
 * for thousands of taxa.  Compile WITHOUT -ffast-math and without FMA contraction:
 *     gcc -O2 -ffp-contract=off -shared -fPIC oracle/c/nj_hard.c -o oracle/_build/liboracle.so
 * pdm: S*S row-major doubles.  peel: (S-1)*3 int64.  blens: 2S-2 doubles.  Returns 0, or -1 on OOM. */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>

int oracle_nj_hard(const double *pdm, int S, int64_t *peel, double *blens) {
    const int N = 2 * S - 1;
    double *d = (double *)malloc((size_t)N * N * sizeof(double));
    double *xsub = (double *)calloc(N, sizeof(double));
    int *pool = (int *)malloc((size_t)N * sizeof(int));
    if (!d || !xsub || !pool) { free(d); free(xsub); free(pool); return -1; }
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j) d[(size_t)i * N + j] = pdm[(size_t)i * S + j];
    for (int i = 0; i < S; ++i) pool[i] = i;
    /* Cpeeler.pyx:46-54 */
    for (int i = 0; i < S; ++i) {
        double acc = 0.0;
        for (int j = 0; j < S; ++j)
            if (j != i) acc += d[(size_t)i * N + j];
        xsub[i] = acc;
    }
    for (int i = 0; i < 2 * S - 2; ++i) blens[i] = 0.0;
    int n_pool = S;
    while (n_pool > 1) {
        /* Cpeeler.pyx:58-66: first strict minimum in pool order */
        double min_q = DBL_MAX;
        int ba = 0, bb = 1;
        const double nm2 = (double)n_pool - 2.0;
        for (int a = 0; a < n_pool - 1; ++a) {
            const int i = pool[a];
            const double xi = xsub[i];
            const double *di = d + (size_t)i * N;
            for (int b = a + 1; b < n_pool; ++b) {
                const int j = pool[b];
                const double v1 = nm2 * di[j];
                const double q = v1 - xi - xsub[j];
                if (q < min_q) { min_q = q; ba = a; bb = b; }
            }
        }
        const int n1 = pool[ba], n2 = pool[bb];
        const int int_i = S - n_pool, parent = int_i + S;
        peel[int_i * 3 + 0] = n1; peel[int_i * 3 + 1] = n2; peel[int_i * 3 + 2] = parent;
        /* remove both from the pool, keeping order */
        int w = 0;
        for (int a = 0; a < n_pool; ++a)
            if (a != ba && a != bb) pool[w++] = pool[a];
        /* Cpeeler.pyx:81-95 */
        const double d12 = d[(size_t)n1 * N + n2];
        double new_x = 0.0;
        for (int a = 0; a < w; ++a) {
            const int nd = pool[a];
            double v1 = 0.0;
            v1 += d[(size_t)nd * N + n1];
            v1 += d[(size_t)nd * N + n2];
            const double dist = 0.5 * (v1 - d12);
            d[(size_t)parent * N + nd] = dist;
            d[(size_t)nd * N + parent] = dist;
            new_x += dist;
            double x = xsub[nd];
            x += dist;
            x -= d[(size_t)n1 * N + nd];
            x -= d[(size_t)n2 * N + nd];
            xsub[nd] = x;
        }
        xsub[parent] = new_x;
        /* Cpeeler.pyx:98-112 */
        if (n_pool > 2) {
            const double v1 = 0.5 * d12;
            const double v4 = 1.0 / (2.0 * ((double)n_pool - 2.0)) * (xsub[n1] - xsub[n2]);
            const double delta_f = v1 + v4;
            blens[n1] = delta_f;
            blens[n2] = d12 - delta_f;
        } else {
            blens[n1] = d12 / 2.0;
            blens[n2] = d12 / 2.0;
        }
        pool[w] = parent;
        n_pool -= 1;
    }
    for (int i = 0; i < 2 * S - 2; ++i)
        if (blens[i] < DBL_EPSILON) blens[i] = DBL_EPSILON; /* Cpeeler.pyx:124 */
    free(d); free(xsub); free(pool);
    return 0;
}
