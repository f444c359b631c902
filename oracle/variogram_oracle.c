/* TEST INFRASTRUCTURE ONLY -- brute-force all-pairs lag binning in plain C (checker + CPU baseline for the variogram).
 * Same arithmetic as oracle/variogram_oracle.py (restated scikit-gstat 1.0.x, PARITY UNPINNED, see that file):
 * float64 Euclidean distance between float64 coordinates, lag class k <=> edges[k-1] <= d < edges[k], per-class pair
 * count and sum of squared differences (Matheron numerator).  Also returns the largest pair distance. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int xo_variogram_pairs(const double* x, const double* y, const double* v, int64_t n, const double* edges, int n_bins,
                       int64_t* count, double* sumsq, double* dmax_out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double dmax = 0.0;
#pragma omp parallel
    {
        int64_t* c = (int64_t*)calloc(n_bins, sizeof(int64_t));
        double* s = (double*)calloc(n_bins, sizeof(double));
        double dm = 0.0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t j = i + 1; j < n; ++j) {
                const double dx = x[i] - x[j], dy = y[i] - y[j];
                const double d = sqrt(dx * dx + dy * dy);
                if (d > dm) dm = d;
                if (!(d < edges[n_bins - 1])) continue;
                int lo = 0, hi = n_bins - 1; /* first k with d < edges[k] */
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (d < edges[mid]) hi = mid; else lo = mid + 1;
                }
                const double df = v[i] - v[j];
                c[lo] += 1;
                s[lo] += df * df;
            }
        }
#pragma omp critical
        {
            for (int k = 0; k < n_bins; ++k) {
                count[k] += c[k];
                sumsq[k] += s[k];
            }
            if (dm > dmax) dmax = dm;
        }
        free(c);
        free(s);
    }
    *dmax_out = dmax;
    return 0;
}
