/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never on the product path.
 *
 * Exact brute-force restatement of the neighbour search the reference delegates to scikit-learn 1.9.0
 * (third-party, not vendored in the reference; call sites sparseSpatialSampling/s_cube.py:161-163,224,328,372
 * and sparseSpatialSampling/export.py:120,423-441).  sklearn's KD-tree ranks candidates by the reduced
 * distance  sum_j (q_j - x_j)^2  accumulated dimension by dimension in double precision without fused
 * multiply-add (sklearn/metrics/_dist_metrics.pyx euclidean_rdist) and reports sqrt() of it; a brute-force
 * scan over the same values therefore returns the same neighbours whenever there is no exact tie
 * (ties: smaller index first here; unspecified in sklearn).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void s3o_knn(const double* X, int64_t n, int d, const double* Q, int64_t nq, int k, int64_t* idx, double* dist) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        double* bd = dist + qi * k;
        int64_t* bi = idx + qi * k;
        int filled = 0;
        const double* q = Q + qi * d;
        for (int64_t i = 0; i < n; ++i) {
            const double* x = X + i * d;
            double r = 0.0;
            for (int j = 0; j < d; ++j) {
                double t = q[j] - x[j];
                r += t * t;
            }
            if (filled == k && !(r < bd[k - 1])) continue; /* equal to the k-th: later index loses */
            int pos = filled < k ? filled : k - 1;
            while (pos > 0 && bd[pos - 1] > r) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = r;
            bi[pos] = i;
            if (filled < k) ++filled;
        }
        for (int j = 0; j < k; ++j) bd[j] = sqrt(bd[j]);
    }
}
