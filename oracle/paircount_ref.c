/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C brute-force restatement of the arithmetic the reference delegates to
 * scipy's cKDTree.count_neighbors (call site: src/yaw/catalog/trees.py:348-353
 * of the reference; scipy 1.18.1 in this image, unpinned upstream).
 *
 *   hist[k-1] += w1[a]*w2[b]   for   r2[k-1] < d2 <= r2[k],   1 <= k < n_edges
 *   d2 = (dx*dx + dy*dy) + dz*dz    IEEE double, products rounded separately
 *
 * Build with -ffp-contract=off so the compiler cannot fuse the products into
 * FMAs (hard part 2: only this evaluation order reproduces
 * scipy on adversarial on-edge pairs).  Rows of set 1 are
 * interleaved over pthreads (PAIRCOUNT_REF_THREADS, default: online cores);
 * integer results are order independent, weighted sums are reduced per thread
 * and then in thread order (deterministic for a fixed thread count).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static inline int find_bin(const double *r2, int n_edges, double d2)
{
    /* number of edges strictly below d2 == np.searchsorted(r2, d2, "left") */
    int lo = 0, hi = n_edges;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (r2[mid] < d2) lo = mid + 1; else hi = mid;
    }
    return lo;
}

typedef struct {
    const double *xyz1, *w1, *xyz2, *w2, *r2;
    int64_t n1, n2;
    int n_edges, tid, nthreads;
    int64_t *hist_i;
    double *hist_f;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const int nsub = j->n_edges - 1;
    const double *r2 = j->r2;
    const double r2_lo = r2[0], r2_hi = r2[nsub];
    for (int64_t a = j->tid; a < j->n1; a += j->nthreads) {
        const double ax = j->xyz1[3 * a], ay = j->xyz1[3 * a + 1], az = j->xyz1[3 * a + 2];
        const double wa = j->w1 ? j->w1[a] : 1.0;
        for (int64_t b = 0; b < j->n2; ++b) {
            const double dx = ax - j->xyz2[3 * b];
            const double dy = ay - j->xyz2[3 * b + 1];
            const double dz = az - j->xyz2[3 * b + 2];
            const double xx = dx * dx;
            const double yy = dy * dy;
            const double zz = dz * dz;
            const double d2 = (xx + yy) + zz;
            if (d2 > r2_lo && d2 <= r2_hi) {
                const int k = nsub == 1 ? 1 : find_bin(r2, j->n_edges, d2);
                j->hist_i[k - 1] += 1;
                j->hist_f[k - 1] += wa * (j->w2 ? j->w2[b] : 1.0);
            }
        }
    }
    return NULL;
}

void paircount_ref(const double *xyz1, int64_t n1, const double *w1,
                   const double *xyz2, int64_t n2, const double *w2,
                   const double *r2, int n_edges,
                   int64_t *hist_i, double *hist_f)
{
    const int nsub = n_edges - 1;
    int nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    const char *env = getenv("PAIRCOUNT_REF_THREADS");
    if (env) nthreads = atoi(env);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((double)n1 * (double)n2 < 1e6) nthreads = 1;

    int64_t *pi = (int64_t *)calloc((size_t)nthreads * nsub, sizeof(int64_t));
    double *pf = (double *)calloc((size_t)nthreads * nsub, sizeof(double));
    job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        job_t jb = {xyz1, w1, xyz2, w2, r2, n1, n2, n_edges, t, nthreads,
                    pi + (size_t)t * nsub, pf + (size_t)t * nsub};
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);

    for (int k = 0; k < nsub; ++k) {
        hist_i[k] = 0;
        hist_f[k] = 0.0;
        for (int t = 0; t < nthreads; ++t) {
            hist_i[k] += pi[(size_t)t * nsub + k];
            hist_f[k] += pf[(size_t)t * nsub + k];
        }
    }
    free(pi); free(pf); free(jobs); free(th);
}
