/* gae.c -- restatement of the PPO2 GAE reverse loop.  TEST INFRASTRUCTURE (see oracle.h).
 * algorithm/policy_base/Proximal_Policy_Optimization2.py:91-98:
 *     deltas = r + self.gamma * (1.0 - success) * vs_ - vs                      (torch float32)
 *     for delta, d in reversed: gae = delta + self.gamma * self.lmd * gae * (1.0 - d); adv.insert(0, gae)
 *     v_target = adv + vs
 * Under numpy >= 2 (NEP 50) python floats are weak, so every product/sum is rounded to float32; volatile-free
 * float temporaries + -ffp-contract=off reproduce exactly that. */
#include "oracle.h"

void oracle_gae(int64_t T, int64_t N, const float *r, const float *vs, const float *vs_next, const float *done,
                const float *success, float gamma, float lambda_gamma, float *adv, float *v_target, double *stats) {
    double s1 = 0, s2 = 0;
    for (int64_t n = 0; n < N; ++n) {
        float gae = 0.f;
        for (int64_t t = T - 1; t >= 0; --t) {
            int64_t i = t * N + n;
            float one_m_s = 1.0f - success[i];
            float g = gamma * one_m_s;
            float gv = g * vs_next[i];
            float delta = (r[i] + gv) - vs[i];
            float a = lambda_gamma * gae;
            float b = a * (1.0f - done[i]);
            gae = delta + b;
            adv[i] = gae;
            v_target[i] = gae + vs[i];
            s1 += (double)gae;
            s2 += (double)gae * (double)gae;
        }
    }
    if (stats) { stats[0] += s1; stats[1] += s2; stats[2] += (double)T * (double)N; }
}
