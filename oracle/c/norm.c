/* norm.c -- restatement of the running normalisation and of the PPO (v1) return scan.  TEST INFRASTRUCTURE (oracle.h).
 *
 * utils/classes.py:626-656:
 *   RunningMeanStd.update(x):  n += 1
 *                              n == 1: mean = x; std = x
 *                              else:   old = mean; mean = old + (x - old) / n; S = S + (x - old) * (x - mean);
 *                                      std = sqrt(S / n)
 *   Normalization.__call__(x, update): if update: update(x);  return (x - mean) / (std + 1e-8)
 * run = double[4][dim] = (n, mean, S, std): std is kept explicitly here, exactly like the reference object. */
#include <math.h>
#include "oracle.h"

void oracle_norm_seq(int64_t rows, int dim, const double *x /* [dim][rows] */, double *y /* [dim][rows] or NULL */,
                     double *run /* [4][dim] */, int update, double eps) {
    for (int f = 0; f < dim; ++f) {
        double n = run[f], mean = run[dim + f], S = run[2 * dim + f], sd = run[3 * dim + f];
        for (int64_t t = 0; t < rows; ++t) {
            const double v = x[(int64_t)f * rows + t];
            if (update) {
                n += 1.0;
                if (n == 1.0) {
                    mean = v;
                    sd = v;
                } else {
                    const double old = mean;
                    mean = old + (v - old) / n;
                    S = S + (v - old) * (v - mean);
                    sd = sqrt(S / n);
                }
            }
            if (y) y[(int64_t)f * rows + t] = (v - mean) / (sd + eps);
        }
        run[f] = n; run[dim + f] = mean; run[2 * dim + f] = S; run[3 * dim + f] = sd;
    }
}

/* algorithm/policy_base/Proximal_Policy_Optimization.py:113-119 (Distributed_PPO.py:58-64): float64 recurrence over a
 * float64 reward column, result stored as float32 (torch.tensor(..., dtype=torch.float32)) */
void oracle_mc_returns(int64_t T, int64_t N, const double *r, const uint8_t *done, double gamma, float *ret) {
    for (int64_t n = 0; n < N; ++n) {
        double acc = 0.0;
        for (int64_t t = T - 1; t >= 0; --t) {
            if (done[t * N + n]) acc = 0.0;
            acc = r[t * N + n] + gamma * acc;
            ret[t * N + n] = (float)acc;
        }
    }
}
