/* cartpole.c -- restatement of the CartPole family, one instance at a time.
 * TEST INFRASTRUCTURE (see oracle.h).
 *   variant 0: environment/CartPole/CartPole.py
 *   variant 1: environment/CartPole/CartPoleAngleOnly.py
 *   variant 2: demonstration/PPO2/PPO2-4-CartPoleAngleOnly/cartpole_angleonly.py
 */
#include <math.h>
#include "oracle.h"
#include "philox.h"

typedef struct { double th, dth, x, dx; } cp_x;

/* CartPole.py:219-238 -- expression order kept (left-to-right products, scalar ** -> pow) */
static cp_x cp_ode(const b200_cartpole_params *p, double force, cp_x s) {
    double S = sin(s.th), C = cos(s.th);
    double ddx = (force + p->m * p->ell * pow(s.dth, 2.0) * S - p->kf * s.dx - 3.0 / 4.0 * p->m * p->g * S * C) /
                 (p->M + p->m - 3.0 / 4.0 * p->m * pow(C, 2.0));
    double ddth = 3.0 / 4.0 / p->m / p->ell * (p->m * p->g * S - p->m * ddx * C);
    cp_x d = {s.dth, ddth, s.dx, ddx};
    return d;
}
static cp_x cp_axpy(cp_x a, cp_x k, double div) { /* a + k / div */
    cp_x r = {a.th + k.th / div, a.dth + k.dth / div, a.x + k.x / div, a.dx + k.dx / div};
    return r;
}
static cp_x cp_scale(double h, cp_x d) { cp_x r = {h * d.th, h * d.dth, h * d.x, h * d.dx}; return r; }

/* CartPole.py:246-251 */
static cp_x cp_rk4(const b200_cartpole_params *p, double force, cp_x xx, double h) {
    cp_x K1 = cp_scale(h, cp_ode(p, force, xx));
    cp_x K2 = cp_scale(h, cp_ode(p, force, cp_axpy(xx, K1, 2.0)));
    cp_x K3 = cp_scale(h, cp_ode(p, force, cp_axpy(xx, K2, 2.0)));
    cp_x K4 = cp_scale(h, cp_ode(p, force, cp_axpy(xx, K3, 1.0)));
    cp_x r;
    r.th = xx.th + (K1.th + 2 * K2.th + 2 * K3.th + K4.th) / 6;
    r.dth = xx.dth + (K1.dth + 2 * K2.dth + 2 * K3.dth + K4.dth) / 6;
    r.x = xx.x + (K1.x + 2 * K2.x + 2 * K3.x + K4.x) / 6;
    r.dx = xx.dx + (K1.dx + 2 * K2.dx + 2 * K3.dx + K4.dx) / 6;
    return r;
}

/* get_state: CartPole.py:145-153, CartPoleAngleOnly.py:249-252, cartpole_angleonly.py:137-143 */
static void cp_observe(const b200_cartpole_params *p, cp_x s, double *o) {
    if (p->variant == 0) {
        o[0] = s.th / p->theta_max * p->static_gain;
        o[1] = s.dth / p->dtheta_max * p->static_gain;
        o[2] = s.x / p->x_max * p->static_gain;
        o[3] = s.dx / p->dx_max * p->static_gain;
    } else {
        o[0] = s.th / p->theta_max * p->static_gain;
        o[1] = s.dth / p->norm_boundless * p->static_gain;
    }
}

static double rad2deg(double r) { return r * 180. / M_PI; } /* utils/functions.py:8-9 */

static void cp_reset_one(const b200_cartpole_params *p, cp_x *s, double *time, uint64_t seed, uint64_t gid, uint32_t ep) {
    orc_rng g;
    orc_rng_init(&g, seed, gid, ep);
    s->th = orc_uniform(&g, p->reset_theta_lo, p->reset_theta_hi); /* CartPole.py:272 */
    s->x = orc_uniform(&g, p->reset_x_lo, p->reset_x_hi);          /* CartPole.py:273 */
    s->dth = 0.; s->dx = 0.; *time = 0.;
}

#define LD(f) io->state[(int64_t)(f) * n + i]

void orc_cartpole_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags,
                           uint64_t seed, int64_t off) {
    const b200_cartpole_params *p = (const b200_cartpole_params *)params;
    const int S = p->variant == 0 ? 4 : 2;
    cp_x s = {LD(0), LD(1), LD(2), LD(3)};
    double time = io->time[i];
    const double force = io->action[i];
    double cur[4], nxt[4];
    cp_observe(p, s, cur); /* self.current_state = self.get_state() */
    int sub = 0;
    if (p->variant == 2) { /* cartpole_angleonly.py:218-229 */
        s = cp_rk4(p, force, s, p->dt);
        time += p->dt;
        sub = 1;
    } else { /* CartPole.py:240-252 */
        double h = p->dt / 10;
        double tt = time + p->dt;
        while (time < tt) {
            s = cp_rk4(p, force, s, h);
            time += h;
            ++sub;
        }
    }
    double eth = 0. - s.th, ex = 0. - s.x;
    int flag = 0, done = 0;
    int angle_out = (s.th > p->theta_term_hi) || (s.th < p->theta_term_lo);
    if (p->variant == 0) { /* CartPole.py:160-185 */
        if (angle_out) { flag = 1; done = 1; }
        if (s.x > p->x_max || s.x < -p->x_max) { flag = 2; done = 1; }
        if (time > p->time_max) { flag = 3; done = 1; }
        if (sqrt(ex * ex + s.dx * s.dx + eth * eth + s.dth * s.dth) < 1e-2) { flag = 4; done = 1; } /* :155-158 */
    } else if (p->variant == 1) { /* CartPoleAngleOnly.py:144-166 */
        if (angle_out) { flag = 1; done = 1; }
        else if (time > p->time_max) { flag = 3; done = 1; }
    } else { /* cartpole_angleonly.py:150-168 */
        if (angle_out) { flag = 1; done = 1; }
        if (time > p->time_max) { flag = 3; done = 1; }
        if (sqrt(eth * eth + s.dth * s.dth) < 1e-2) { flag = 4; done = 1; }
    }
    cp_observe(p, s, nxt);
    double reward;
    if (p->variant == 0) { /* CartPole.py:187-217 */
        double r_x = -fabs(s.x) * 5, r_dx = -fabs(s.dx) * 0.0, r_theta = -fabs(s.th) * 1;
        double r_omega = -fabs(s.dth) * 0.0, r_f = -fabs(force) * 0.01;
        double r_extra = 0.;
        if (flag == 1 || flag == 2) {
            double _n = (p->time_max - time) / p->dt;
            r_extra = _n * (r_x + r_dx + r_theta + r_omega + r_f);
        }
        reward = r_x + r_dx + r_theta + r_omega + r_f + r_extra;
    } else if (p->variant == 1) { /* CartPoleAngleOnly.py:188-208 */
        double ce = fabs(rad2deg(cur[0] / p->static_gain * p->theta_max));
        double ne = fabs(rad2deg(nxt[0] / p->static_gain * p->theta_max));
        double r = ne > ce ? -2 : (ne == ce ? 0 : 2);
        if (ce <= 0.5 && ne <= 0.5) r += 5;
        if (flag == 1) r -= 100;
        else if (flag == 3) r += 500;
        reward = r;
    } else { /* cartpole_angleonly.py:170-195 */
        double r1 = -pow(s.th, 2.0) * 10, r2 = -pow(s.dth, 2.0) * 0.0, r3 = -pow(force, 2.0) * 0.00, r4 = 0.;
        if (flag == 1) {
            double _n = (p->time_max - time) / p->dt;
            r4 = _n * (r1 + r2 + r3);
        }
        reward = r1 + r2 + r3 + r4;
    }
    for (int k = 0; k < S; ++k) {
        if (io->obs) io->obs[(int64_t)k * n + i] = cur[k];
        io->next_obs[(int64_t)k * n + i] = nxt[k];
    }
    io->reward[i] = reward;
    io->done[i] = (uint8_t)done;
    io->flag[i] = flag;
    if (io->substeps) io->substeps[i] = sub;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        uint32_t ep = io->episode[i];
        cp_reset_one(p, &s, &time, seed, (uint64_t)(off + i), ep);
        io->episode[i] = ep + 1u;
        cp_observe(p, s, nxt);
    }
    if (io->reset_obs)
        for (int k = 0; k < S; ++k) io->reset_obs[(int64_t)k * n + i] = nxt[k];
    LD(0) = s.th; LD(1) = s.dth; LD(2) = s.x; LD(3) = s.dx;
    io->time[i] = time;
}

void orc_cartpole_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed,
                            int64_t off, int observe_only) {
    const b200_cartpole_params *p = (const b200_cartpole_params *)params;
    cp_x s = {LD(0), LD(1), LD(2), LD(3)};
    if (!observe_only) {
        double time;
        uint32_t ep = io->episode[i];
        cp_reset_one(p, &s, &time, seed, (uint64_t)(off + i), ep);
        io->episode[i] = ep + 1u;
        LD(0) = s.th; LD(1) = s.dth; LD(2) = s.x; LD(3) = s.dx;
        io->time[i] = time;
    }
    if (io->next_obs) {
        double o[4];
        cp_observe(p, s, o);
        const int S = p->variant == 0 ? 4 : 2;
        for (int k = 0; k < S; ++k) io->next_obs[(int64_t)k * n + i] = o[k];
    }
}
