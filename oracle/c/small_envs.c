/* small_envs.c -- restatements of Flight_Attitude_Simulator, SecondOrderIntegration, BallBalancer1D,
 * TwoLinkManipulator and UGVForward / UGVBidirectional, one instance at a time.
 * TEST INFRASTRUCTURE (see oracle.h).  Expression order follows the cited reference lines; scalar `**` is pow(). */
#include <math.h>
#include "oracle.h"
#include "philox.h"

#define SF(f) io->state[(int64_t)(f) * n + i]
#define OUT(buf, k) (buf)[(int64_t)(k) * n + i]

static void emit(const oracle_io *io, int64_t n, int64_t i, int S, const double *cur, const double *nxt, double reward,
                 int done, int flag) {
    for (int k = 0; k < S; ++k) {
        if (io->obs) OUT(io->obs, k) = cur[k];
        OUT(io->next_obs, k) = nxt[k];
    }
    io->reward[i] = reward;
    io->done[i] = (uint8_t)done;
    io->flag[i] = flag;
}
static void emit_policy(const oracle_io *io, int64_t n, int64_t i, int S, const double *o) {
    if (io->reset_obs) for (int k = 0; k < S; ++k) OUT(io->reset_obs, k) = o[k];
}

/* ============================================================ Flight_Attitude_Simulator */
/* environment/FlightAttitudeSimulator/FlightAttitudeSimulator.py */
typedef b200_fas_params FP;
static void fas_obs(const FP *p, double th, double dth, double *o) { /* :173-185 */
    o[0] = (2 * th - p->max_theta - p->min_theta) / (p->max_theta - p->min_theta) * p->static_gain;
    o[1] = (2 * dth - p->max_omega - p->min_omega) / (p->max_omega - p->min_omega) * p->static_gain;
}
static double fas_dd(const FP *p, double force, double dth) { /* :232-236 */
    return (force * p->L - p->mgd - p->k * dth) / p->denom;
}
static void fas_reset(const FP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    SF(0) = orc_uniform(&g, p->reset_lo, p->reset_hi); /* :270-271 */
    SF(1) = 0.;
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_fas_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const FP *p = (const FP *)params;
    double th = SF(0), dth = SF(1), time = io->time[i], force = io->action[i];
    double cur[2], nxt[2];
    fas_obs(p, th, dth, cur);
    double h = p->dt / 10, tt = time + p->dt; /* :238-252 */
    int sub = 0;
    while (time < tt) {
        double K1a = h * dth, K1b = h * fas_dd(p, force, dth);
        double K2a = h * (dth + K1b / 2), K2b = h * fas_dd(p, force, dth + K1b / 2);
        double K3a = h * (dth + K2b / 2), K3b = h * fas_dd(p, force, dth + K2b / 2);
        double K4a = h * (dth + K3b), K4b = h * fas_dd(p, force, dth + K3b);
        th = th + (K1a + 2 * K2a + 2 * K3a + K4a) / 6;
        dth = dth + (K1b + 2 * K2b + 2 * K3b + K4b) / 6;
        time += h;
        ++sub;
    }
    int flag = 0; /* :193-211 */
    if (th > p->theta_term_hi) flag = 1;
    if (th < p->theta_term_lo) flag = 2;
    if (time > p->time_max) flag = 3;
    int done = flag != 0;
    fas_obs(p, th, dth, nxt);
    double r1 = -pow(th, 2.0) * p->Q, r2 = -pow(dth, 2.0) * p->R, r3 = 0.; /* :217-230 */
    if (flag == 1 || flag == 2) { double _n = (p->time_max - time) / p->dt; r3 = _n * (r1 + r2); }
    emit(io, n, i, 2, cur, nxt, r1 + r2 + r3, done, flag);
    if (io->substeps) io->substeps[i] = sub;
    SF(0) = th; SF(1) = dth; io->time[i] = time;
    if (done && (flags & B200ENV_AUTO_RESET)) { fas_reset(p, io, n, i, seed, off); fas_obs(p, SF(0), SF(1), nxt); }
    emit_policy(io, n, i, 2, nxt);
}
void orc_fas_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const FP *p = (const FP *)params;
    if (!observe_only) fas_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double o[2]; fas_obs(p, SF(0), SF(1), o); for (int k = 0; k < 2; ++k) OUT(io->next_obs, k) = o[k]; }
}

/* ============================================================ FlightAttitudeSimulatorDiscrete */
/* environment/FlightAttitudeSimulator/FlightAttitudeSimulatorDiscrete.py */
typedef b200_fas_discrete_params FDP;
static void fasd_obs(const FDP *p, double th, double dth, double *o) { /* :158-162 */
    o[0] = -th / p->theta_max * p->static_gain;
    o[1] = dth / p->dtheta_max * p->static_gain;
}
static double fasd_f(const FDP *p, double a0, double angle, double dangle) { /* :199-203 */
    return p->a2 * dangle + p->a1 * cos(angle) + a0;
}
static void fasd_reset(const FDP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    SF(0) = orc_uniform(&g, -p->theta_max, p->theta_max); /* :257-260 */
    SF(1) = 0.;
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_fas_discrete_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const FDP *p = (const FDP *)params;
    double th = SF(0), dth = SF(1), time = io->time[i];
    const double a0 = p->L * io->action[i] / p->denom; /* :204 */
    double cur[2], nxt[2];
    fasd_obs(p, th, dth, cur);
    double h = p->dt / 1, t_sim = 0.0; /* :205-219 */
    int sub = 0;
    while (t_sim <= p->dt) {
        double K1 = dth, L1 = fasd_f(p, a0, th, dth);
        double K2 = dth + h * L1 / 2, L2 = fasd_f(p, a0, th + h * K1 / 2, dth + h * L1 / 2);
        double K3 = dth + h * L2 / 2, L3 = fasd_f(p, a0, th + h * K2 / 2, dth + h * L2 / 2);
        double K4 = dth + h * L3, L4 = fasd_f(p, a0, th + h * K3, dth + h * L3);
        th = th + h * (K1 + 2 * K2 + 2 * K3 + K4) / 6;
        dth = dth + h * (L1 + 2 * L2 + 2 * L3 + L4) / 6;
        t_sim = t_sim + h;
        ++sub;
    }
    if (th > p->theta_max) { th = p->theta_max; dth = p->bounce * dth; }   /* :220-222 */
    if (th < -p->theta_max) { th = -p->theta_max; dth = p->bounce * dth; } /* :223-225 */
    time = time + p->dt;
    fasd_obs(p, th, dth, nxt);
    int flag = 0; /* is_Terminal :178-196: first true test returns */
    if (th > p->theta_out || th < -p->theta_out) flag = 1;
    else if (time > p->time_max) flag = 2;
    int done = flag != 0;
    double r1 = -pow(th, 2.0) * p->Q, r2 = -pow(dth, 2.0) * p->R, r3 = 0.; /* :232-244 */
    if (flag == 1 || flag == 2) { double _n = (p->time_max - time) / p->dt; r3 = _n * (r1 + r2); }
    emit(io, n, i, 2, cur, nxt, r1 + r2 + r3, done, flag);
    if (io->substeps) io->substeps[i] = sub;
    SF(0) = th; SF(1) = dth; io->time[i] = time;
    if (done && (flags & B200ENV_AUTO_RESET)) { fasd_reset(p, io, n, i, seed, off); fasd_obs(p, SF(0), SF(1), nxt); }
    emit_policy(io, n, i, 2, nxt);
}
void orc_fas_discrete_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const FDP *p = (const FDP *)params;
    if (!observe_only) fasd_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double o[2]; fasd_obs(p, SF(0), SF(1), o); for (int k = 0; k < 2; ++k) OUT(io->next_obs, k) = o[k]; }
}

/* ============================================================ SecondOrderIntegration */
/* environment/SecondOrderIntegration/SecondOrderIntegration.py */
typedef b200_soi_params SP;
static void soi_obs(const SP *p, const double *s, double *o) { /* :211-219 */
    o[0] = (p->target_x - s[0]) / p->map_x * p->obs_gain;
    o[1] = (p->target_y - s[1]) / p->map_y * p->obs_gain;
    o[2] = -s[2] / p->vmax * p->obs_gain;
    o[3] = -s[3] / p->vmax * p->obs_gain;
}
static void soi_reset(const SP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    SF(0) = orc_uniform(&g, 0 + p->reset_margin, p->map_x - p->reset_margin); /* :329-331 */
    SF(1) = orc_uniform(&g, 0 + p->reset_margin, p->map_y - p->reset_margin);
    SF(2) = 0.; SF(3) = 0.;
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_soi_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const SP *p = (const SP *)params;
    double s[4] = {SF(0), SF(1), SF(2), SF(3)}, time = io->time[i];
    double f[2] = {io->action[i], io->action[n + i]};
    double cur[4], nxt[4];
    soi_obs(p, s, cur);
    double h = p->dt / 1, tt = time + p->dt; /* :298-314 */
    int sub = 0;
    while (time < tt) {
        double K1[4], K2[4], K3[4], K4[4], t[4];
#define SOI_ODE(x, K) { K[0] = h * x[2]; K[1] = h * x[3]; K[2] = h * (f[0] - p->k * x[2]); K[3] = h * (f[1] - p->k * x[3]); }
        SOI_ODE(s, K1);
        for (int k = 0; k < 4; ++k) t[k] = s[k] + K1[k] / 2;
        SOI_ODE(t, K2);
        for (int k = 0; k < 4; ++k) t[k] = s[k] + K2[k] / 2;
        SOI_ODE(t, K3);
        for (int k = 0; k < 4; ++k) t[k] = s[k] + K3[k];
        SOI_ODE(t, K4);
        for (int k = 0; k < 4; ++k) s[k] = s[k] + (K1[k] + 2 * K2[k] + 2 * K3[k] + K4[k]) / 6;
        time += h;
        ++sub;
    }
    double acc[2] = {(f[0] - p->k * s[2]) / p->mass, (f[1] - p->k * s[3]) / p->mass};
    double ex = p->target_x - s[0], ey = p->target_y - s[1];
    double e_pos = sqrt(fma(ey, ey, ex * ex)), e_vel = sqrt(fma(s[3], s[3], s[2] * s[2]));   /* np.linalg.norm: see norm2 in ugvo.c */
    int flag = 0; /* :235-249 */
    if (s[0] > p->map_x + p->admissible_error || s[0] < 0 - p->admissible_error ||
        s[1] > p->map_y + p->admissible_error || s[1] < 0 - p->admissible_error) flag = 1;
    if (time > p->time_max) flag = 2;
    if (p->success_terminal && e_pos <= 0.05 && e_vel < 0.05) flag = 3;
    int done = flag != 0;
    soi_obs(p, s, nxt);
    double a_n = sqrt(fma(acc[1], acc[1], acc[0] * acc[0])); /* :251-284 */
    double u_pos = -e_pos * p->Q_pos, u_vel = -e_vel * p->Q_vel, u_acc = -a_n * p->Q_acc, u_extra = 0.;
    if (flag == 1) { double _n = (p->time_max - time) / p->dt; u_extra = _n * (u_pos + u_vel + u_acc); }
    emit(io, n, i, 4, cur, nxt, u_pos + u_vel + u_acc + u_extra, done, flag);
    if (io->substeps) io->substeps[i] = sub;
    for (int k = 0; k < 4; ++k) SF(k) = s[k];
    io->time[i] = time;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        soi_reset(p, io, n, i, seed, off);
        double r[4] = {SF(0), SF(1), SF(2), SF(3)};
        soi_obs(p, r, nxt);
    }
    emit_policy(io, n, i, 4, nxt);
}
void orc_soi_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const SP *p = (const SP *)params;
    if (!observe_only) soi_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double s[4] = {SF(0), SF(1), SF(2), SF(3)}, o[4]; soi_obs(p, s, o); for (int k = 0; k < 4; ++k) OUT(io->next_obs, k) = o[k]; }
}

/* ============================================================ BallBalancer1D */
/* environment/BallBalancer/BallBalancer1D.py */
typedef b200_ballbalancer_params BP;
static double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
static void bb_obs(const BP *p, double pos, double vel, double th, double *o) { /* :200-211 */
    o[0] = pos / p->L * p->static_gain;
    o[1] = (2 * vel - p->v_max - p->v_min) / (p->v_max - p->v_min) * p->static_gain;
    o[2] = (2 * th - p->theta_max - p->theta_min) / (p->theta_max - p->theta_min) * p->static_gain;
}
static int bb_success(const BP *p, double error, double vel, double th) { /* :213-216 */
    return fabs(error) <= 0.001 && fabs(vel) <= 0.005 && fabs(th) <= p->deg1;
}
static void bb_reset(const BP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    SF(2) = orc_uniform(&g, p->reset_theta_lo, p->reset_theta_hi); /* :294-296 */
    SF(0) = orc_uniform(&g, p->reset_pos_lo, p->reset_pos_hi);
    SF(1) = p->init_vel;
    SF(3) = p->target - SF(0);
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_ballbalancer_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const BP *p = (const BP *)params;
    double pos = SF(0), vel = SF(1), th = SF(2), error = SF(3), time = io->time[i];
    double cur[3], nxt[3];
    bb_obs(p, pos, vel, th, cur);
    double omega = clipd(io->action[i], p->omega_min, p->omega_max); /* :255 */
    double h = p->dt / 10, tt = time + p->dt;
    int sub = 0;
    while (time < tt) { /* :257-270; ode :248-252 = [vel, K sin(theta), omega] */
        double K1[3] = {h * vel, h * (p->K * sin(th)), h * omega};
        double K2[3] = {h * (vel + K1[1] / 2), h * (p->K * sin(th + K1[2] / 2)), h * omega};
        double K3[3] = {h * (vel + K2[1] / 2), h * (p->K * sin(th + K2[2] / 2)), h * omega};
        double K4[3] = {h * (vel + K3[1]), h * (p->K * sin(th + K3[2])), h * omega};
        double n0 = pos + (K1[0] + 2 * K2[0] + 2 * K3[0] + K4[0]) / 6;
        double n1 = vel + (K1[1] + 2 * K2[1] + 2 * K3[1] + K4[1]) / 6;
        double n2 = th + (K1[2] + 2 * K2[2] + 2 * K3[2] + K4[2]) / 6;
        pos = n0;
        vel = clipd(n1, p->v_min, p->v_max);
        th = clipd(n2, p->theta_min, p->theta_max);
        time += h;
        ++sub;
    }
    int flag, done; /* :218-236, first true test returns; is_success sees the previous step's error */
    if (pos < -p->L || pos > p->L) { flag = 1; done = 1; }
    else if (time > p->time_max) { flag = 2; done = 1; }
    else if (bb_success(p, error, vel, th)) { flag = 3; done = 1; }
    else { flag = 0; done = 0; }
    error = p->target - pos; /* :281 */
    bb_obs(p, pos, vel, th, nxt);
    double e = error / p->L * p->static_gain; /* :238-246 */
    double r1 = -pow(e, 2.0) - tanh(100 * e) + 0.5, r2 = 0, r3 = bb_success(p, error, vel, th) ? 1000 : 0;
    emit(io, n, i, 3, cur, nxt, r1 + r2 + r3, done, flag);
    if (io->substeps) io->substeps[i] = sub;
    SF(0) = pos; SF(1) = vel; SF(2) = th; SF(3) = error; io->time[i] = time;
    if (done && (flags & B200ENV_AUTO_RESET)) { bb_reset(p, io, n, i, seed, off); bb_obs(p, SF(0), SF(1), SF(2), nxt); }
    emit_policy(io, n, i, 3, nxt);
}
void orc_ballbalancer_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const BP *p = (const BP *)params;
    if (!observe_only) bb_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double o[3]; bb_obs(p, SF(0), SF(1), SF(2), o); for (int k = 0; k < 3; ++k) OUT(io->next_obs, k) = o[k]; }
}

/* ============================================================ TwoLinkManipulator */
/* environment/RobotManipulator/TwoLinkManipulator.py */
typedef b200_twolink_params TP;
/* np.linalg.solve(a, b) for 2x2: dgesv = LU with partial pivoting + substitution */
static void solve2(const double a[2][2], const double b[2], double x[2]) {
    double r0[2] = {a[0][0], a[0][1]}, r1[2] = {a[1][0], a[1][1]}, b0 = b[0], b1 = b[1];
    if (fabs(r1[0]) > fabs(r0[0])) { double t; t = r0[0]; r0[0] = r1[0]; r1[0] = t; t = r0[1]; r0[1] = r1[1]; r1[1] = t; t = b0; b0 = b1; b1 = t; }
    double l = r1[0] * (1.0 / r0[0]);
    double u11 = r1[1] - l * r0[1];
    double y1 = b1 - l * b0;
    x[1] = y1 / u11;
    x[0] = (b0 - r0[1] * x[1]) / r0[0];
}
static void tlm_ode(const TP *p, const double xx[4], const double tq[2], double d[4]) { /* :226-238 */
    double _theta1 = xx[0], _theta2 = xx[1], _dtheta1 = xx[2], _dtheta2 = xx[3], J = p->J;
    double a[2][2] = {{J * (5 + 3 * cos(_theta2)), J * (1 + 3.0 / 2 * cos(_theta2))}, {J * (1 + 3.0 / 2 * cos(_theta2)), J}};
    double b[2] = {tq[0] + 3.0 / 2 * J * sin(_theta2) * pow(_dtheta2, 2.0) + 3 * J * sin(_theta2) * _dtheta1 * _dtheta2 -
                       p->m * p->g * p->l * (3.0 / 2 * sin(_theta1) + 1.0 / 2 * sin(_theta1 + _theta2)),
                   tq[1] - 3.0 / 2 * J * sin(_theta2) * pow(_dtheta1, 2.0) - 1.0 / 2 * p->m * p->g * p->l * sin(_theta1 + _theta2)};
    double w[2];
    solve2(a, b, w);
    d[0] = _dtheta1; d[1] = _dtheta2; d[2] = w[0]; d[3] = w[1];
}
static void tlm_obs(const oracle_io *io, int64_t n, int64_t i, double *o) { /* :186-192, normalisation is the identity */
    o[0] = SF(4); o[1] = SF(5); o[2] = SF(0); o[3] = SF(1); o[4] = SF(2); o[5] = SF(3);
}
static void tlm_reset(const TP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    double phi = orc_u01(&g) * 2 * M_PI; /* :283-291 */
    double r = orc_uniform(&g, p->r2_lo, p->r2_hi);
    SF(6) = cos(phi) * sqrt(r) + p->base_x;
    SF(7) = sin(phi) * sqrt(r) + p->base_y;
    SF(0) = orc_uniform(&g, -p->theta_max, p->theta_max);
    SF(1) = orc_uniform(&g, -p->theta_max, p->theta_max);
    SF(2) = 0.; SF(3) = 0.;
    SF(4) = SF(6) - p->init_end_x; SF(5) = SF(7) - p->init_end_y; /* :300 */
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_twolink_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const TP *p = (const TP *)params;
    double xx[4] = {SF(0), SF(1), SF(2), SF(3)}, time = io->time[i];
    double tq[2] = {io->action[i], io->action[n + i]};
    double cur[6], nxt[6];
    tlm_obs(io, n, i, cur);
    double K1[4], K2[4], K3[4], K4[4], t[4], d[4]; /* :240-250 */
    tlm_ode(p, xx, tq, d);
    for (int k = 0; k < 4; ++k) { K1[k] = p->dt * d[k]; t[k] = xx[k] + K1[k] / 2; }
    tlm_ode(p, t, tq, d);
    for (int k = 0; k < 4; ++k) { K2[k] = p->dt * d[k]; t[k] = xx[k] + K2[k] / 2; }
    tlm_ode(p, t, tq, d);
    for (int k = 0; k < 4; ++k) { K3[k] = p->dt * d[k]; t[k] = xx[k] + K3[k]; }
    tlm_ode(p, t, tq, d);
    for (int k = 0; k < 4; ++k) { K4[k] = p->dt * d[k]; xx[k] = xx[k] + (K1[k] + 2 * K2[k] + 2 * K3[k] + K4[k]) / 6; }
    time += p->dt;
    double th_sum = 0 + xx[0] + xx[1]; /* sum(self.theta) */
    double midx = p->l * sin(xx[0]) + p->base_x, midy = -p->l * cos(xx[0]) + p->base_y; /* :252-253 */
    double endx = midx + p->l * sin(th_sum), endy = midy + -p->l * cos(th_sum);
    SF(4) = SF(6) - endx; SF(5) = SF(7) - endy; /* :257 */
    if (xx[0] > p->theta_max) xx[0] -= 2 * p->theta_max; else if (xx[0] < -p->theta_max) xx[0] += 2 * p->theta_max; /* :259-272 */
    if (xx[1] > p->theta_max) xx[1] -= 2 * p->theta_max; else if (xx[1] < -p->theta_max) xx[1] += 2 * p->theta_max;
    for (int k = 0; k < 4; ++k) SF(k) = xx[k];
    io->time[i] = time;
    double en = sqrt(fma(SF(5), SF(5), SF(4) * SF(4))), wn = sqrt(fma(xx[3], xx[3], xx[2] * xx[2]));
    int flag, done; /* :199-209 */
    if (time > p->time_max) { flag = 2; done = 1; }
    else if (en <= p->miss && wn <= p->omega_ok) { flag = 3; done = 1; }
    else { flag = 0; done = 0; }
    tlm_obs(io, n, i, nxt);
    double tn = sqrt(fma(tq[1], tq[1], tq[0] * tq[0])); /* :211-224 */
    double reward = -en * p->Q_pos + -wn * p->Q_omega + -tn * p->Q_acc + 0.;
    emit(io, n, i, 6, cur, nxt, reward, done, flag);
    if (done && (flags & B200ENV_AUTO_RESET)) { tlm_reset(p, io, n, i, seed, off); tlm_obs(io, n, i, nxt); }
    emit_policy(io, n, i, 6, nxt);
}
void orc_twolink_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const TP *p = (const TP *)params;
    if (!observe_only) tlm_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double o[6]; tlm_obs(io, n, i, o); for (int k = 0; k < 6; ++k) OUT(io->next_obs, k) = o[k]; }
}

/* ============================================================ UGVForward / UGVBidirectional */
/* environment/UGV/UGVForward.py, UGVBidirectional.py; utils/functions.py:49-60 */
typedef b200_ugv_params UP;
static double vec_rad_oriented(double x1, double y1, double x2, double y2) {
    if (sqrt(fma(y2, y2, x2 * x2)) < 1e-4 || sqrt(fma(y1, y1, x1 * x1)) < 1e-4) return 0;
    double dot = x1 * x2 + y1 * y2, det = x1 * y2 - y1 * x2;
    return atan2(det, dot);
}
static void ugv_err(const UP *p, const double *s, double *e, double *ephi) {
    double dx = p->target_x - s[0], dy = p->target_y - s[1];
    *e = sqrt(fma(dy, dy, dx * dx));
    *ephi = vec_rad_oriented(cos(s[3]), sin(s[3]), dx, dy);
    if (p->bidirectional) { /* UGVBidirectional.py:314-325 */
        double d = cos(s[3]) * dx + sin(s[3]) * dy;
        double sg = d > 0 ? 1. : (d < 0 ? -1. : 0.);
        *e = sg * *e;
        if (*ephi >= M_PI / 2) *ephi = *ephi - M_PI;
        if (*ephi <= -M_PI / 2) *ephi = *ephi + M_PI;
    }
}
static void ugv_obs(const UP *p, const double *s, double *o) { /* :217-227 */
    double e, ephi;
    ugv_err(p, s, &e, &ephi);
    if (p->bidirectional) { o[0] = e / p->e_max * p->static_gain; o[1] = s[2] / p->v_max * p->static_gain; }
    else { o[0] = (2 / p->e_max * e - 1) * p->static_gain; o[1] = (2 / p->v_max * s[2] - 1) * p->static_gain; }
    o[2] = ephi / p->e_phi_max * p->static_gain;
    o[3] = s[4] / p->omega_max * p->static_gain;
}
static void ugv_reset(const UP *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    orc_rng g;
    uint32_t ep = io->episode[i];
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    SF(0) = orc_uniform(&g, p->reset_d0, p->map_x - p->reset_d0); /* :336-338 */
    SF(1) = orc_uniform(&g, p->reset_d0, p->map_y - p->reset_d0);
    SF(3) = orc_uniform(&g, -M_PI, M_PI);
    SF(2) = 0.; SF(4) = 0.;
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}
void orc_ugv_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const UP *p = (const UP *)params;
    double s[5] = {SF(0), SF(1), SF(2), SF(3), SF(4)}, time = io->time[i];
    double al = io->action[i], aa = io->action[n + i];
    double cur[4], nxt[4];
    ugv_obs(p, s, cur);
    double K1[5], K2[5], K3[5], K4[5], t[5];
#define UGV_ODE(x, K) { K[0] = p->dt * (x[2] * cos(x[3])); K[1] = p->dt * (x[2] * sin(x[3])); K[2] = p->dt * (al - p->kf * x[2]); \
                        K[3] = p->dt * x[4]; K[4] = p->dt * (aa - p->kt * x[4]); }
    UGV_ODE(s, K1); /* :294-301 */
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K1[k] / 2;
    UGV_ODE(t, K2);
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K2[k] / 2;
    UGV_ODE(t, K3);
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K3[k];
    UGV_ODE(t, K4);
    for (int k = 0; k < 5; ++k) s[k] = s[k] + (K1[k] + 2 * K2[k] + 2 * K3[k] + K4[k]) / 6;
    if (!p->bidirectional && s[2] < 0.) s[2] = 0.;
    time += p->dt;
    if (s[3] > M_PI) s[3] -= 2 * M_PI;
    if (s[3] < -M_PI) s[3] += 2 * M_PI;
    double e, ephi;
    ugv_err(p, s, &e, &ephi);
    int flag = 0; /* :247-261 */
    if (s[0] > p->map_x || s[0] < 0 || s[1] > p->map_y || s[1] < 0) flag = 1;
    if (time > p->time_max) flag = 2;
    if (fabs(e) <= 0.05 && fabs(s[2]) < 0.01) flag = 3;
    int done = flag != 0;
    ugv_obs(p, s, nxt);
    double u_pos = -fabs(e) * p->Q_pos, u_vel = -fabs(s[2]) * p->Q_vel; /* :263-279 */
    double u_phi = e > 0.1 ? -fabs(ephi) * p->Q_phi : 0.0, u_omega = -fabs(s[4]) * p->Q_omega, u_psi = 0.;
    if (flag == 1) { double _n = (p->time_max - time) / p->dt; u_psi = _n * (u_pos + u_vel + u_phi + u_omega); }
    emit(io, n, i, 4, cur, nxt, u_pos + u_vel + u_phi + u_omega + u_psi, done, flag);
    for (int k = 0; k < 5; ++k) SF(k) = s[k];
    io->time[i] = time;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        ugv_reset(p, io, n, i, seed, off);
        double r[5] = {SF(0), SF(1), SF(2), SF(3), SF(4)};
        ugv_obs(p, r, nxt);
    }
    emit_policy(io, n, i, 4, nxt);
}
void orc_ugv_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const UP *p = (const UP *)params;
    if (!observe_only) ugv_reset(p, io, n, i, seed, off);
    if (io->next_obs) { double s[5] = {SF(0), SF(1), SF(2), SF(3), SF(4)}, o[4]; ugv_obs(p, s, o); for (int k = 0; k < 4; ++k) OUT(io->next_obs, k) = o[k]; }
}
