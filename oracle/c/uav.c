/* uav.c -- restatement of the UavFntsmcParam attitude / position tracking envs, one instance at a time.
 * TEST INFRASTRUCTURE (see oracle.h).  "One env step" is the fused train.py loop body
 *   get_param_from_actor(a) -> generate_action_4_uav() | ref_inner()+att_control() -> step_update()
 * (demonstration/PPO2/PPO2-4-UavFntsmcParamPos/train.py:292-297, ...Att/train.py:265-276).
 *
 * parity note: numpy evaluates the vector pow/tanh of FNTSMC.py with AVX-512 SIMD loops that differ from
 * glibc by <= 1-3 ulp, and np.dot / np.linalg.inv go through OpenBLAS; this restatement uses glibc and
 * plain sequential sums, so it tracks the reference to ~1e-14, not bit-exactly (tests state the bound).
 */
#include <math.h>
#include <string.h>
#include "oracle.h"
#include "philox.h"

typedef b200_uav_params P;

typedef struct {
    double x[12];      /* x y z vx vy vz phi theta psi p q r */
    double time;
} uav_t;

/* uav.py:93-124 */
static void uav_ode(const P *p, const double xx[12], double throttle, const double tq[3], const double dis[3], double d[12]) {
    const double _vx = xx[3], _vy = xx[4], _vz = xx[5], _phi = xx[6], _theta = xx[7], _psi = xx[8];
    const double _p = xx[9], _q = xx[10], _r = xx[11];
    const double *J = p->J;
    double dp = (-p->kr * _p - _q * _r * (J[2] - J[1]) + tq[0]) / J[0];
    double dq = (-p->kr * _q - _p * _r * (J[0] - J[2]) + tq[1]) / J[1];
    double dr = (-p->kr * _r - _p * _q * (J[1] - J[0]) + tq[2]) / J[2];
    double R[3][3] = {{1, tan(_theta) * sin(_phi), tan(_theta) * cos(_phi)},
                      {0, cos(_phi), -sin(_phi)},
                      {0, sin(_phi) / cos(_theta), cos(_phi) / cos(_theta)}};
    double dphi = R[0][0] * _p + R[0][1] * _q + R[0][2] * _r;
    double dtheta = R[1][0] * _p + R[1][1] * _q + R[1][2] * _r;
    double dpsi = R[2][0] * _p + R[2][1] * _q + R[2][2] * _r;
    double dvx = (throttle * (cos(_psi) * sin(_theta) * cos(_phi) + sin(_psi) * sin(_phi)) - p->kt * _vx + dis[0]) / p->m;
    double dvy = (throttle * (sin(_psi) * sin(_theta) * cos(_phi) - cos(_psi) * sin(_phi)) - p->kt * _vy + dis[1]) / p->m;
    double dvz = -p->g + (throttle * cos(_phi) * cos(_theta) - p->kt * _vz + dis[2]) / p->m;
    d[0] = _vx; d[1] = _vy; d[2] = _vz; d[3] = dvx; d[4] = dvy; d[5] = dvz;
    d[6] = dphi; d[7] = dtheta; d[8] = dpsi; d[9] = dp; d[10] = dq; d[11] = dr;
}

/* uav.py:126-148 with n = 1 (as every caller passes) */
static void uav_rk44(const P *p, uav_t *u, const double action[4], const double dis[3], int att_only) {
    const double h = p->dt / 1;
    double K1[12], K2[12], K3[12], K4[12], t[12], d[12];
    uav_ode(p, u->x, action[0], action + 1, dis, d);
    for (int i = 0; i < 12; ++i) { K1[i] = h * d[i]; t[i] = u->x[i] + K1[i] / 2; }
    uav_ode(p, t, action[0], action + 1, dis, d);
    for (int i = 0; i < 12; ++i) { K2[i] = h * d[i]; t[i] = u->x[i] + K2[i] / 2; }
    uav_ode(p, t, action[0], action + 1, dis, d);
    for (int i = 0; i < 12; ++i) { K3[i] = h * d[i]; t[i] = u->x[i] + K3[i]; }
    uav_ode(p, t, action[0], action + 1, dis, d);
    for (int i = 0; i < 12; ++i) { K4[i] = h * d[i]; }
    for (int i = 0; i < 12; ++i) u->x[i] = u->x[i] + (K1[i] + 2 * K2[i] + 2 * K3[i] + K4[i]) / 6;
    if (att_only) for (int i = 0; i < 6; ++i) u->x[i] = 0.;
    u->time += p->dt;
    if (u->x[8] > M_PI) u->x[8] -= 2 * M_PI;
    if (u->x[8] < -M_PI) u->x[8] += 2 * M_PI;
}

/* uav.py:285-300 */
static void uav_f1(const uav_t *u, double f[3][3]) {
    const double phi = u->x[6], theta = u->x[7];
    memset(f, 0, 9 * sizeof(double));
    f[0][0] = 1.;
    f[0][1] = sin(phi) * tan(theta);
    f[0][2] = cos(phi) * tan(theta);
    f[1][1] = cos(phi);
    f[1][2] = -sin(phi);
    f[2][1] = sin(phi) / cos(theta);
    f[2][2] = cos(phi) / cos(theta);
}
static void mat3_vec(const double m[3][3], const double v[3], double o[3]) { /* np.dot(3x3, 3) */
    for (int i = 0; i < 3; ++i) o[i] = m[i][0] * v[0] + m[i][1] * v[1] + m[i][2] * v[2];
}
/* uav.py:302-313 */
static void uav_f2(const P *p, const uav_t *u, double o[3]) {
    const double pp = u->x[9], q = u->x[10], r = u->x[11];
    const double *J = p->J;
    o[0] = (p->kr * pp + q * r * (J[1] - J[2])) / J[0];
    o[1] = (p->kr * q + pp * r * (J[2] - J[0])) / J[1];
    o[2] = (p->kr * r + pp * q * (J[0] - J[1])) / J[2];
}
/* uav.py:340-357: F() . rho2 + f1() . f2() */
static void uav_second_order_att_dynamics(const P *p, const uav_t *u, double o[3]) {
    const double phi = u->x[6], theta = u->x[7];
    double f1[3][3], d1[3], F[3][3], a[3], f2[3], b[3];
    uav_f1(u, f1);
    mat3_vec(f1, u->x + 9, d1); /* dot_rho1 */
    memset(F, 0, sizeof(F));
    F[0][1] = d1[0] * tan(theta) * cos(phi) + d1[1] * sin(phi) / pow(cos(theta), 2.0);
    F[0][2] = -d1[0] * tan(theta) * sin(phi) + d1[1] * cos(phi) / pow(cos(theta), 2.0);
    F[1][1] = -d1[0] * sin(phi);
    F[1][2] = -d1[0] * cos(phi);
    double t1 = d1[0] * cos(phi) * cos(theta) + d1[1] * sin(phi) * sin(theta);
    F[2][1] = t1 / pow(cos(theta), 2.0);
    double t2 = -d1[0] * sin(phi) * cos(theta) + d1[1] * cos(phi) * sin(theta);
    F[2][2] = t2 / pow(cos(theta), 2.0);
    mat3_vec(F, u->x + 9, a);
    uav_f2(p, u, f2);
    mat3_vec(f1, f2, b);
    for (int i = 0; i < 3; ++i) o[i] = a[i] + b[i];
}

/* np.linalg.inv(3x3): LAPACK dgesv(A, I) = LU with partial pivoting (dgetf2 order), then forward/back substitution */
static void inv3(const double a[3][3], double inv[3][3]) {
    double lu[3][3];
    int piv[3] = {0, 1, 2};
    memcpy(lu, a, sizeof(lu));
    for (int k = 0; k < 3; ++k) {
        int pk = k;
        double best = fabs(lu[k][k]);
        for (int i = k + 1; i < 3; ++i)
            if (fabs(lu[i][k]) > best) { best = fabs(lu[i][k]); pk = i; }
        if (pk != k) {
            for (int j = 0; j < 3; ++j) { double t = lu[k][j]; lu[k][j] = lu[pk][j]; lu[pk][j] = t; }
            int t = piv[k]; piv[k] = piv[pk]; piv[pk] = t;
        }
        double r = 1.0 / lu[k][k]; /* dgetf2 scales the column by the reciprocal of the pivot */
        for (int i = k + 1; i < 3; ++i) lu[i][k] *= r;
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) lu[i][j] -= lu[i][k] * lu[k][j];
    }
    for (int c = 0; c < 3; ++c) {
        double y[3];
        for (int i = 0; i < 3; ++i) y[i] = (piv[i] == c) ? 1.0 : 0.0; /* P * e_c */
        for (int i = 1; i < 3; ++i)
            for (int j = 0; j < i; ++j) y[i] -= lu[i][j] * y[j];
        for (int i = 2; i >= 0; --i) {
            for (int j = i + 1; j < 3; ++j) y[i] -= lu[i][j] * y[j];
            y[i] /= lu[i][i];
        }
        for (int i = 0; i < 3; ++i) inv[i][c] = y[i];
    }
}

typedef struct { double k1[3], k2[3], alpha[3], beta[3], gamma[3], lmd[3]; } gains_t;

/* FNTSMC.py:112-137 */
static void fntsmc_att_update(const P *p, const gains_t *g, double s1[3], const double sec[3], const double B[3][3],
                              const double e[3], const double de[3], const double dd_ref[3], double control[3]) {
    double v[3];
    for (int i = 0; i < 3; ++i) {
        double s = 1 * de[i] + g->k1[i] * e[i] + g->gamma[i] * pow(fabs(e[i]), g->alpha[i]) * tanh(5 * e[i]);
        double dot_s1 = pow(fabs(s), g->beta[i]) * tanh(5 * s);
        s1[i] += dot_s1 * p->dt;
        double sigma = s + g->lmd[i] * s1[i];
        double u1 = sec[i] + dd_ref[i] + g->k1[i] * de[i] +
                    g->gamma[i] * g->alpha[i] * pow(fabs(e[i]), g->alpha[i] - 1) * de[i] + g->lmd[i] * dot_s1;
        double u2 = -g->k2[i] * tanh(10 * sigma);
        v[i] = u1 + u2;
    }
    double inv[3][3], o[3];
    inv3(B, inv);
    mat3_vec(inv, v, o);
    for (int i = 0; i < 3; ++i) control[i] = -o[i];
}

/* FNTSMC.py:47-69 (obs = 0) */
static void fntsmc_pos_update(const P *p, const gains_t *g, double sigma_o1[3], const double vel[3], const double e[3],
                              const double de[3], const double dd_ref[3], double control[3]) {
    for (int i = 0; i < 3; ++i) {
        double sigma_o = de[i] + g->k1[i] * e[i] + g->gamma[i] * pow(fabs(e[i]), g->alpha[i]) * tanh(5 * e[i]);
        double dot_sigma_o1 = pow(fabs(sigma_o), g->beta[i]) * tanh(5 * sigma_o);
        sigma_o1[i] += dot_sigma_o1 * p->dt;
        double so = sigma_o + g->lmd[i] * sigma_o1[i];
        double uo1 = p->kt / p->m * vel[i] + dd_ref[i] - g->k1[i] * de[i] -
                     g->gamma[i] * g->alpha[i] * pow(fabs(e[i]), g->alpha[i] - 1) * de[i] - g->lmd[i] * dot_sigma_o1;
        double uo2 = -g->k2[i] * so - 0.;
        control[i] = uo1 + uo2;
    }
}

/* attitude loop shared by both envs: uav_att_ctrl.py:91-108 / uav_pos_ctrl.py:317-337 */
static void att_control(const P *p, const uav_t *u, const gains_t *g, double s1[3], const double ref[3],
                        const double dot_ref[3], double torque[3]) {
    double f1[3][3], d1[3], e[3], de[3], sec[3], B[3][3], zero[3] = {0, 0, 0};
    uav_f1(u, f1);
    mat3_vec(f1, u->x + 9, d1);
    for (int i = 0; i < 3; ++i) { e[i] = u->x[6 + i] - ref[i]; de[i] = d1[i] - dot_ref[i]; }
    uav_second_order_att_dynamics(p, u, sec);
    for (int i = 0; i < 3; ++i) /* att_control_matrix = f1 . h, h = diag(1/J)  uav.py:315-326,359-360 */
        for (int j = 0; j < 3; ++j) B[i][j] = f1[i][j] * (1 / p->J[j]);
    fntsmc_att_update(p, g, s1, sec, B, e, de, zero, torque);
}

/* uav.py:182-219 */
static int uav_terminal_flag(const P *p, const uav_t *u) {
    int flag = 0, out = 0;
    for (int i = 0; i < 3; ++i) if (u->x[i] < p->pos_lo[i] || u->x[i] > p->pos_hi[i]) out = 1;
    if (out) flag = 2;
    out = 0;
    for (int i = 0; i < 3; ++i) if (u->x[6 + i] < p->att_lo[i] || u->x[6 + i] > p->att_hi[i]) out = 1;
    if (out) flag = 3;
    if (u->time > p->t_term) flag = 1;
    return flag;
}

/* ref_cmd.py:4-43 (ref_inner and ref_uav share the formula) */
static void ref_cmd(int dim, double time, const double *A, const double *T, const double *bias, const double *phase,
                    double *r, double *dr, double *ddr) {
    for (int i = 0; i < dim; ++i) {
        double w = 2 * M_PI / T[i];
        r[i] = A[i] * sin(w * time + phase[i]) + bias[i];
        dr[i] = A[i] * w * cos(w * time + phase[i]);
        ddr[i] = -A[i] * (w * w) * sin(w * time + phase[i]);
    }
}

#define SF(f) io->state[(int64_t)(f) * n + i]

/* ------------------------------------------------------------------ attitude env */
enum { A_S1 = 6, A_K1 = 9, A_K2 = 12, A_GAM = 15, A_LMD = 18, A_AMP = 21, A_PER = 24, A_PHS = 27, A_REF = 30, A_DREF = 33 };

static void att_draw_reset(const P *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    uint32_t ep = io->episode[i];
    for (int k = 0; k < 6; ++k) SF(k) = p->init_state[6 + k]; /* reset_uav(): angle0, then pqr <- pos0 (N5) */
    for (int k = 0; k < 3; ++k) {
        SF(A_S1 + k) = 0.;
        SF(A_K1 + k) = p->att_k1[k]; SF(A_K2 + k) = p->att_k2[k];
        SF(A_GAM + k) = p->att_gamma[k]; SF(A_LMD + k) = p->att_lmd[k];
    }
    if (p->random_trajectory) { /* uav_att_ctrl.py:156-161 */
        orc_rng g;
        orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
        for (int k = 0; k < 3; ++k) SF(A_AMP + k) = orc_uniform(&g, 0., p->traj_A_hi[k]);
        for (int k = 0; k < 3; ++k) SF(A_PER + k) = orc_uniform(&g, p->traj_T_lo, p->traj_T_hi);
        for (int k = 0; k < 3; ++k) SF(A_PHS + k) = orc_uniform(&g, 0., p->traj_phase_hi);
    } else {
        for (int k = 0; k < 3; ++k) {
            SF(A_AMP + k) = p->ref_amplitude[k]; SF(A_PER + k) = p->ref_period[k]; SF(A_PHS + k) = p->ref_bias_phase[k];
        }
    }
    if (p->yaw_fixed) { SF(A_AMP + 2) = 0.; SF(A_PHS + 2) = 0.; }
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}

/* uav_att_ctrl_RL.py:59-68 with use_norm = False; uav_dot_att() = T_pqr_2_dot_att . pqr (uav.py:171-177) */
static void att_observe(const uav_t *u, const double ref[3], const double dot_ref[3], double o[6]) {
    double f1[3][3], d1[3];
    uav_f1(u, f1);
    mat3_vec(f1, u->x + 9, d1);
    for (int k = 0; k < 3; ++k) { o[k] = u->x[6 + k] - ref[k]; o[3 + k] = d1[k] - dot_ref[k]; }
}

void orc_uav_att_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags,
                          uint64_t seed, int64_t off) {
    const P *p = (const P *)params;
    uav_t u;
    memset(&u, 0, sizeof(u));
    for (int k = 0; k < 6; ++k) u.x[6 + k] = SF(k);
    u.time = io->time[i];
    gains_t g;
    double s1[3], A[3], T[3], ph[3], zero3[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        s1[k] = SF(A_S1 + k);
        g.k1[k] = SF(A_K1 + k); g.k2[k] = SF(A_K2 + k); g.gamma[k] = SF(A_GAM + k); g.lmd[k] = SF(A_LMD + k);
        g.alpha[k] = p->att_alpha[k]; g.beta[k] = p->att_beta[k];
        A[k] = SF(A_AMP + k); T[k] = SF(A_PER + k); ph[k] = SF(A_PHS + k);
    }
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = io->action[(int64_t)k * n + i];
    /* get_param_from_actor: uav_att_ctrl_RL.py:141-156 */
    for (int k = 0; k < 3; ++k) {
        if (a[k] > 0) g.k1[k] = 10 * a[k];
        if (a[k + 3] > 0) g.k2[k] = a[k + 3] / 10;
    }
    if (a[6] > 0) for (int k = 0; k < 3; ++k) g.gamma[k] = a[6];
    if (a[7] > 0) for (int k = 0; k < 3; ++k) g.lmd[k] = a[7];
    /* ref_inner + att_control(rhod, dot_rhod, None): train.py:267-272 */
    double ref[3], dref[3], ddref[3], torque[3];
    ref_cmd(3, u.time, A, T, zero3, ph, ref, dref, ddref);
    att_control(p, &u, &g, s1, ref, dref, torque);
    /* step_update: uav_att_ctrl_RL.py:129-139 */
    double cur[6], nxt[6];
    att_observe(&u, ref, dref, cur);
    double act4[4] = {p->m * p->g / (cos(u.x[6]) * cos(u.x[7])), torque[0], torque[1], torque[2]}; /* uav_att_ctrl.py:115 */
    uav_rk44(p, &u, act4, zero3, 1);
    int flag = uav_terminal_flag(p, &u);
    int done = (flag == 1 || flag == 3); /* uav_att_ctrl_RL.py:118-127 */
    att_observe(&u, ref, dref, nxt);
    /* get_reward: uav_att_ctrl_RL.py:70-106 */
    double u_att = -(nxt[0] * nxt[0] * p->Q_e[0] + nxt[1] * nxt[1] * p->Q_e[1] + nxt[2] * nxt[2] * p->Q_e[2]);
    double u_pqr = -(nxt[3] * nxt[3] * p->Q_de[0] + nxt[4] * nxt[4] * p->Q_de[1] + nxt[5] * nxt[5] * p->Q_de[2]);
    double u_acc = -(torque[0] * torque[0] * p->R[0] + torque[1] * torque[1] * p->R[1] + torque[2] * torque[2] * p->R[2]);
    double u_extra = 0.;
    if (flag == 3) {
        double _n = (p->time_max - u.time) / p->dt;
        double _u_phi = 0., _u_theta = 0., _u_psi = 0.;
        if (u.x[6] > p->att_zone_max[0] || u.x[6] < p->att_zone_min[0]) _u_phi = -pow(M_PI, 2.0) * p->Q_e[0];
        if (u.x[7] > p->att_zone_max[1] || u.x[7] < p->att_zone_min[1]) _u_theta = -pow(M_PI, 2.0) * p->Q_e[1];
        if (u.x[8] > p->att_zone_max[2] || u.x[8] < p->att_zone_min[2]) _u_theta = -4 * pow(M_PI, 2.0) * p->Q_e[2]; /* sic, N7 */
        u_extra = _n * (_u_phi + _u_theta + _u_psi + u_pqr + u_acc);
    }
    double reward = u_att + u_pqr + u_acc + u_extra;
    for (int k = 0; k < 6; ++k) {
        if (io->obs) io->obs[(int64_t)k * n + i] = cur[k];
        io->next_obs[(int64_t)k * n + i] = nxt[k];
    }
    io->reward[i] = reward; io->done[i] = (uint8_t)done; io->flag[i] = flag;
    for (int k = 0; k < 6; ++k) SF(k) = u.x[6 + k];
    for (int k = 0; k < 3; ++k) {
        SF(A_S1 + k) = s1[k];
        SF(A_K1 + k) = g.k1[k]; SF(A_K2 + k) = g.k2[k]; SF(A_GAM + k) = g.gamma[k]; SF(A_LMD + k) = g.lmd[k];
    }
    io->time[i] = u.time;
    if (done) {
        if (flags & B200ENV_AUTO_RESET) {
            att_draw_reset(p, io, n, i, seed, off);
            uav_t r;
            memset(&r, 0, sizeof(r));
            for (int k = 0; k < 6; ++k) r.x[6 + k] = SF(k);
            att_observe(&r, ref, dref, nxt); /* first obs of the new episode uses the stale ref (reference quirk) */
        } else {
            for (int k = 0; k < 3; ++k) { SF(A_REF + k) = ref[k]; SF(A_DREF + k) = dref[k]; }
        }
    }
    if (io->reset_obs) for (int k = 0; k < 6; ++k) io->reset_obs[(int64_t)k * n + i] = nxt[k];
}

void orc_uav_att_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off,
                           int observe_only) {
    const P *p = (const P *)params;
    if (!observe_only) att_draw_reset(p, io, n, i, seed, off);
    if (io->next_obs) {
        uav_t u;
        memset(&u, 0, sizeof(u));
        for (int k = 0; k < 6; ++k) u.x[6 + k] = SF(k);
        double ref[3], dref[3], o[6];
        for (int k = 0; k < 3; ++k) { ref[k] = SF(A_REF + k); dref[k] = SF(A_DREF + k); }
        att_observe(&u, ref, dref, o);
        for (int k = 0; k < 6; ++k) io->next_obs[(int64_t)k * n + i] = o[k];
    }
}

/* ------------------------------------------------------------------ position env */
enum { P_SIG = 12, P_S1 = 15, P_AREF = 18, P_K1 = 21, P_K2 = 24, P_GAM = 27, P_LMD = 30, P_AMP = 33, P_PER = 37,
       P_PHS = 41, P_PREF = 45, P_DPREF = 48, P_NEXT_PQR0 = 51 /* layout variant 1 */ };

static void pos_draw_reset(const P *p, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    uint32_t ep = io->episode[i];
    for (int k = 0; k < 12; ++k) SF(k) = p->init_state[k];
    for (int k = 0; k < 3; ++k) {
        SF(P_SIG + k) = 0.; SF(P_S1 + k) = 0.;
        SF(P_K1 + k) = p->pos_k1[k]; SF(P_K2 + k) = p->pos_k2[k];
        SF(P_GAM + k) = p->pos_gamma[k]; SF(P_LMD + k) = p->pos_lmd[k];
    }
    orc_rng g;
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    if (p->random_trajectory) { /* uav_pos_ctrl.py:404-408 */
        double a = orc_uniform(&g, 0., p->traj_A_hi[0]);
        double T = orc_uniform(&g, p->traj_T_lo, p->traj_T_hi);
        for (int k = 0; k < 3; ++k) SF(P_AMP + k) = a;
        SF(P_AMP + 3) = 0.;
        for (int k = 0; k < 4; ++k) { SF(P_PER + k) = T * 1.0; SF(P_PHS + k) = p->ref_bias_phase[k]; }
    } else {
        for (int k = 0; k < 4; ++k) {
            SF(P_AMP + k) = p->ref_amplitude[k]; SF(P_PER + k) = p->ref_period[k]; SF(P_PHS + k) = p->ref_bias_phase[k];
        }
    }
    if (p->yaw_fixed) { SF(P_AMP + 3) = 0.; SF(P_PHS + 3) = 0.; }
    if (p->random_pos0) { /* uav_pos_ctrl.py:510-513, set_random_init_pos :457-465, reset_uav_with_param uav.py:252-268 */
        for (int k = 0; k < 3; ++k) {
            double t0 = p->ref_bias_a[k] + SF(P_AMP + k) * sin(SF(P_PHS + k)); /* trajectory[0][k], :386-390 at t = 0 */
            double r = fabs(p->init_pos_r[k]);
            double pos0 = orc_uniform(&g, t0 - r, t0 + r);
            SF(k) = pos0;
            SF(9 + k) = SF(P_NEXT_PQR0 + k); /* pqr0 = init_state[9:12] = previous pos0 (N5) */
            SF(P_NEXT_PQR0 + k) = pos0;      /* init_state = concatenate((pos0, vel0, angle0, pos0)) uav.py:268 */
        }
    }
    /* att_ref is NOT reset by the reference (uav_pos_ctrl.py:488-533) */
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}

static void pos_observe(const uav_t *u, const double pref[3], const double dpref[3], double o[6]) {
    for (int k = 0; k < 3; ++k) { o[k] = u->x[k] - pref[k]; o[3 + k] = u->x[3 + k] - dpref[k]; }
}

void orc_uav_pos_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags,
                          uint64_t seed, int64_t off) {
    const P *p = (const P *)params;
    uav_t u;
    for (int k = 0; k < 12; ++k) u.x[k] = SF(k);
    u.time = io->time[i];
    gains_t gp, ga;
    double sig[3], s1[3], att_ref[3], A[4], T[4], ph[4];
    for (int k = 0; k < 3; ++k) {
        sig[k] = SF(P_SIG + k); s1[k] = SF(P_S1 + k); att_ref[k] = SF(P_AREF + k);
        gp.k1[k] = SF(P_K1 + k); gp.k2[k] = SF(P_K2 + k); gp.gamma[k] = SF(P_GAM + k); gp.lmd[k] = SF(P_LMD + k);
        gp.alpha[k] = p->pos_alpha[k]; gp.beta[k] = p->pos_beta[k];
        ga.k1[k] = p->att_k1[k]; ga.k2[k] = p->att_k2[k]; ga.gamma[k] = p->att_gamma[k]; ga.lmd[k] = p->att_lmd[k];
        ga.alpha[k] = p->att_alpha[k]; ga.beta[k] = p->att_beta[k];
    }
    for (int k = 0; k < 4; ++k) { A[k] = SF(P_AMP + k); T[k] = SF(P_PER + k); ph[k] = SF(P_PHS + k); }
    double a[8], dis[3] = {0, 0, 0};
    for (int k = 0; k < 8; ++k) a[k] = io->action[(int64_t)k * n + i];
    if (io->dis) for (int k = 0; k < 3; ++k) dis[k] = io->dis[(int64_t)k * n + i];
    /* get_param_from_actor: uav_pos_ctrl_RL.py:158-173 */
    for (int k = 0; k < 3; ++k) {
        if (a[k] > 0) gp.k1[k] = a[k];
        if (a[k + 3] > 0) gp.k2[k] = a[k + 3];
    }
    if (a[6] > 0) for (int k = 0; k < 3; ++k) gp.gamma[k] = a[6];
    if (a[7] > 0) for (int k = 0; k < 3; ++k) gp.lmd[k] = a[7];
    /* generate_action_4_uav: uav_pos_ctrl.py:467-486 */
    double ref[4], dref[4], ddref[4];
    ref_cmd(4, u.time, A, T, p->ref_bias_a, ph, ref, dref, ddref);
    /* pos_control: uav_pos_ctrl.py:302-315 */
    double e[3], de[3], ctrl[3];
    for (int k = 0; k < 3; ++k) { e[k] = u.x[k] - ref[k]; de[k] = u.x[3 + k] - dref[k]; }
    fntsmc_pos_update(p, &gp, sig, u.x + 3, e, de, ddref, ctrl);
    /* uo_2_ref_angle_throttle: uav_pos_ctrl.py:339-357 */
    const double phi = u.x[6], theta = u.x[7], psi = u.x[8];
    double uf = (ctrl[2] + p->g) * p->m / (cos(phi) * cos(theta));
    double asin_phi_d = fmin(fmax((ctrl[0] * sin(psi) - ctrl[1] * cos(psi)) * p->m / uf, -1), 1);
    double phi_d = asin(asin_phi_d);
    double asin_theta_d = fmin(fmax((ctrl[0] * cos(psi) + ctrl[1] * sin(psi)) * p->m / (uf * cos(phi_d)), -1), 1);
    double theta_d = asin(asin_theta_d);
    phi_d = fmax(fmin(phi_d, p->att_limit), -p->att_limit);
    theta_d = fmax(fmin(theta_d, p->att_limit), -p->att_limit);
    double dot_phi_d = (phi_d - att_ref[0]) / p->dt;
    double dot_theta_d = (theta_d - att_ref[1]) / p->dt;
    double rho_d[3] = {phi_d, theta_d, ref[3]};
    double dot_rho_d[3] = {dot_phi_d, dot_theta_d, dref[3]};
    for (int k = 0; k < 3; ++k) { /* np.clip + rho_d += dot_rho_d * dt */
        dot_rho_d[k] = fmin(fmax(dot_rho_d[k], -p->dot_att_ref_limit), p->dot_att_ref_limit);
        rho_d[k] += dot_rho_d[k] * p->dt;
    }
    double torque[3];
    for (int k = 0; k < 3; ++k) att_ref[k] = rho_d[k];
    att_control(p, &u, &ga, s1, rho_d, dot_rho_d, torque);
    double act4[4] = {uf, torque[0], torque[1], torque[2]};
    /* step_update: uav_pos_ctrl_RL.py:145-156 */
    double cur[6], nxt[6];
    pos_observe(&u, ref, dref, cur);
    uav_rk44(p, &u, act4, dis, 0);
    int flag = uav_terminal_flag(p, &u);
    int done = flag != 0; /* uav_pos_ctrl_RL.py:132-143 */
    pos_observe(&u, ref, dref, nxt);
    /* get_reward: uav_pos_ctrl_RL.py:82-120 */
    double u_pos = -(nxt[0] * nxt[0] * p->Q_e[0] + nxt[1] * nxt[1] * p->Q_e[1] + nxt[2] * nxt[2] * p->Q_e[2]);
    double u_vel = -(nxt[3] * nxt[3] * p->Q_de[0] + nxt[4] * nxt[4] * p->Q_de[1] + nxt[5] * nxt[5] * p->Q_de[2]);
    double u_acc = -(ctrl[0] * ctrl[0] * p->R[0] + ctrl[1] * ctrl[1] * p->R[1] + ctrl[2] * ctrl[2] * p->R[2]);
    double u_extra = 0.;
    if (flag == 2 || flag == 3) {
        double _n = (p->time_max - u.time) / p->dt;
        u_extra = _n * (u_pos + u_vel + u_acc);
    }
    double reward = u_pos + u_vel + u_acc + u_extra;
    for (int k = 0; k < 6; ++k) {
        if (io->obs) io->obs[(int64_t)k * n + i] = cur[k];
        io->next_obs[(int64_t)k * n + i] = nxt[k];
    }
    io->reward[i] = reward; io->done[i] = (uint8_t)done; io->flag[i] = flag;
    for (int k = 0; k < 12; ++k) SF(k) = u.x[k];
    for (int k = 0; k < 3; ++k) {
        SF(P_SIG + k) = sig[k]; SF(P_S1 + k) = s1[k]; SF(P_AREF + k) = att_ref[k];
        SF(P_K1 + k) = gp.k1[k]; SF(P_K2 + k) = gp.k2[k]; SF(P_GAM + k) = gp.gamma[k]; SF(P_LMD + k) = gp.lmd[k];
    }
    io->time[i] = u.time;
    if (done) {
        if (flags & B200ENV_AUTO_RESET) {
            pos_draw_reset(p, io, n, i, seed, off);
            uav_t r;
            for (int k = 0; k < 12; ++k) r.x[k] = SF(k);
            pos_observe(&r, ref, dref, nxt); /* stale pos_ref, like the reference's reset */
        } else {
            for (int k = 0; k < 3; ++k) { SF(P_PREF + k) = ref[k]; SF(P_DPREF + k) = dref[k]; }
        }
    }
    if (io->reset_obs) for (int k = 0; k < 6; ++k) io->reset_obs[(int64_t)k * n + i] = nxt[k];
}

void orc_uav_pos_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off,
                           int observe_only) {
    const P *p = (const P *)params;
    if (!observe_only) pos_draw_reset(p, io, n, i, seed, off);
    if (io->next_obs) {
        uav_t u;
        for (int k = 0; k < 12; ++k) u.x[k] = SF(k);
        double pref[3], dpref[3], o[6];
        for (int k = 0; k < 3; ++k) { pref[k] = SF(P_PREF + k); dpref[k] = SF(P_DPREF + k); }
        pos_observe(&u, pref, dpref, o);
        for (int k = 0; k < 6; ++k) io->next_obs[(int64_t)k * n + i] = o[k];
    }
}

/* ==================================================================================================================
 * UavRobust: environment/UavRobust/{uav.py:429-560, FNTSMC.py:80-106, uav_pos_ctrl.py:41-76, uav_att_ctrl.py:19-46,
 * UavHoverOuterLoop.py, UavHover.py, UavInnerLoop.py, UavTrackingOuterLoop.py}.  The quadrotor ODE / RK4 / f1 / f2 / F
 * are the same code as UavFntsmcParam's, so the static helpers above are reused through a b200_uav_params shim.
 * ================================================================================================================== */
typedef b200_uavrobust_params RP;
enum { R_S1 = 12, R_AREF = 15, R_DAREF = 18, R_PREF = 21, R_AMP = 24, R_PER = 27, R_PHS = 30 };

static void rp_shim(const RP *r, P *p) {
    memset(p, 0, sizeof(*p));
    p->m = r->m; p->g = r->g; p->kr = r->kr; p->kt = r->kt; p->dt = r->dt; p->time_max = r->time_max;
    for (int k = 0; k < 3; ++k) p->J[k] = r->J[k];
}

static int rob_pos_out(const RP *r, const uav_t *u) { /* uav.py:511-525 */
    int f = 0;
    for (int k = 0; k < 3; ++k) if (u->x[k] < r->pos_zone_min[k] || u->x[k] > r->pos_zone_max[k]) f = 1;
    return f;
}
static int rob_att_out(const RP *r, const uav_t *u) { /* uav.py:527-541 */
    int f = 0;
    for (int k = 0; k < 3; ++k) if (u->x[6 + k] < r->att_zone_min[k] || u->x[6 + k] > r->att_zone_max[k]) f = 1;
    return f;
}
static double sqn(const double *v, int n) { /* np.linalg.norm(v) ** 2 */
    double s = 0;
    for (int k = 0; k < n; ++k) s += v[k] * v[k];
    return pow(sqrt(s), 2.0);
}
static double sqn_tanh10(const double *v, int n) { /* np.linalg.norm(np.tanh(10 * v)) ** 2 */
    double t[6];
    for (int k = 0; k < n; ++k) t[k] = tanh(10 * v[k]);
    return sqn(t, n);
}

static void rob_observe(const RP *r, const P *p, const uav_t *u, const oracle_io *io, int64_t n, int64_t i, double *o) {
    double f1[3][3], d1[3];
    uav_f1(u, f1);
    mat3_vec(f1, u->x + 9, d1); /* dot_rho1 */
    const double g = r->static_gain;
    if (r->variant == 0 || r->variant == 1) { /* UavHoverOuterLoop.py:81-92, UavHover.py:101-114 */
        for (int k = 0; k < 3; ++k) {
            o[k] = (u->x[k] - SF(R_PREF + k)) / r->e_pos_span[k] * g;
            o[3 + k] = 2 * u->x[3 + k] / r->vel_span[k] * g;
        }
        if (r->variant == 1)
            for (int k = 0; k < 3; ++k) {
                o[6 + k] = (u->x[6 + k] - SF(R_AREF + k)) / r->e_att_span[k] * g;
                o[9 + k] = (d1[k] - SF(R_DAREF + k)) / r->e_dot_att_span_neg[k] * g;
            }
    } else if (r->variant == 2) { /* UavInnerLoop.py:88-99 */
        double A[3], T[3], ph[3], z3[3] = {0, 0, 0}, ref[3], dref[3], ddref[3];
        for (int k = 0; k < 3; ++k) { A[k] = SF(R_AMP + k); T[k] = SF(R_PER + k); ph[k] = SF(R_PHS + k); }
        ref_cmd(3, u->time, A, T, z3, ph, ref, dref, ddref);
        for (int k = 0; k < 3; ++k) {
            o[k] = (u->x[6 + k] - ref[k]) / r->e_att_span[k] * g;
            o[3 + k] = (d1[k] - dref[k]) / r->e_dot_att_span_neg[k] * g;
        }
    } else { /* UavTrackingOuterLoop.py:90-101 */
        double A[3], T[3], ph[3], ref[3], dref[3], ddref[3];
        for (int k = 0; k < 3; ++k) { A[k] = SF(R_AMP + k); T[k] = SF(R_PER + k); ph[k] = SF(R_PHS + k); }
        ref_cmd(3, u->time, A, T, r->ref_bias_a, ph, ref, dref, ddref);
        for (int k = 0; k < 3; ++k) {
            o[k] = (u->x[k] - ref[k]) / r->e_pos_span[k] * g;
            o[3 + k] = (u->x[3 + k] - dref[k]) / r->vel_span[k] * g;
        }
    }
}

static void rob_reset(const RP *r, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off) {
    uint32_t ep = io->episode[i];
    orc_rng g;
    orc_rng_init(&g, seed, (uint64_t)(off + i), ep);
    for (int k = 0; k < 3; ++k) { SF(k) = r->pos0[k]; SF(3 + k) = r->vel0[k]; SF(6 + k) = r->angle0[k]; SF(9 + k) = r->pqr0[k]; }
    if (r->variant == 0 || r->variant == 1) { /* generate_random_point(offset = 1.0) */
        for (int k = 0; k < 3; ++k) SF(R_PREF + k) = orc_uniform(&g, r->target_lo[k], r->target_hi[k]);
        if (r->variant == 1) for (int k = 0; k < 3; ++k) SF(R_AREF + k) = 0.; /* UavHover.py:209 */
    } else { /* generate_random_signal UavInnerLoop.py:197-210 / generate_random_trajectory UavTrackingOuterLoop.py:222-250 */
        for (int k = 0; k < 3; ++k) SF(R_AMP + k) = orc_uniform(&g, 0., r->sig_A_hi[k]);
        for (int k = 0; k < 3; ++k) SF(R_PER + k) = orc_uniform(&g, r->sig_T_lo, r->sig_T_hi);
        for (int k = 0; k < 3; ++k) SF(R_PHS + k) = orc_uniform(&g, 0., r->sig_phase_hi);
        if (r->variant == 3) /* set_random_init_pos(trajectory[0], 0.3): trajectory[0] = bias + A sin(phase) */
            for (int k = 0; k < 3; ++k) {
                double t0 = r->ref_bias_a[k] + SF(R_AMP + k) * sin(2 * M_PI / SF(R_PER + k) * 0. + SF(R_PHS + k));
                SF(k) = orc_uniform(&g, t0 - fabs(r->init_pos_r), t0 + fabs(r->init_pos_r));
            }
    }
    /* the FNTSMC integrator s1 (and att_ref / dot_att_ref except where noted) survive the reference's reset() */
    io->time[i] = 0.;
    io->episode[i] = ep + 1u;
}

void orc_uavrobust_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const RP *r = (const RP *)params;
    P p;
    rp_shim(r, &p);
    const int V = r->variant, S = V == 1 ? 12 : 6, AD = V == 1 ? 6 : 3;
    uav_t u;
    for (int k = 0; k < 12; ++k) u.x[k] = SF(k);
    u.time = io->time[i];
    double a[6], dis[3] = {0, 0, 0};
    for (int k = 0; k < AD; ++k) a[k] = io->action[(int64_t)k * n + i];
    if (io->dis) for (int k = 0; k < 3; ++k) dis[k] = io->dis[(int64_t)k * n + i];
    double cur[12], nxt[12];
    rob_observe(r, &p, &u, io, n, i, cur);
    if (V == 2) { /* UavInnerLoop.py:127-142 + uav_att_ctrl.py:33-46 */
        double act4[4] = {0, a[0], a[1], a[2]}, z3[3] = {0, 0, 0};
        for (int k = 0; k < 6; ++k) u.x[k] = SF(k);
        uav_rk44(&p, &u, act4, z3, 1);
    } else {
        /* uo_2_ref_angle_throttle, uav_pos_ctrl.py:67-76 */
        const double phi = u.x[6], theta = u.x[7], psi = u.x[8];
        double uf = (a[2] + p.g) * p.m / (cos(phi) * cos(theta));
        double asin_phi_d = fmin(fmax((a[0] * sin(psi) - a[1] * cos(psi)) * p.m / uf, -1), 1);
        double phi_d = asin(asin_phi_d);
        double asin_theta_d = fmin(fmax((a[0] * cos(psi) + a[1] * sin(psi)) * p.m / (uf * cos(phi_d)), -1), 1);
        double theta_d = asin(asin_theta_d);
        phi_d = fmin(fmax(phi_d, r->att_zone_min[0]), r->att_zone_max[0]); /* np.clip */
        theta_d = fmin(fmax(theta_d, r->att_zone_min[1]), r->att_zone_max[1]);
        double att_new[3] = {phi_d, theta_d, 0.0}, att_ref[3], dot_att_ref[3];
        for (int k = 0; k < 3; ++k) {
            const double old = SF(R_AREF + k);
            double d = (att_new[k] - old) / p.dt;
            d = fmin(fmax(d, r->dot_att_min[k]), r->dot_att_max[k]);
            dot_att_ref[k] = d;
            att_ref[k] = d * p.dt + old;
        }
        double torque[3];
        if (V == 1) {
            for (int k = 0; k < 3; ++k) torque[k] = a[3 + k];
        } else { /* att_control + FNTSMC with saturation, uav_pos_ctrl.py:41-65, FNTSMC.py:80-106 */
            gains_t ga;
            double s1[3];
            for (int k = 0; k < 3; ++k) {
                ga.k1[k] = r->att_k1[k]; ga.k2[k] = r->att_k2[k]; ga.alpha[k] = r->att_alpha[k]; ga.beta[k] = r->att_beta[k];
                ga.gamma[k] = r->att_gamma[k]; ga.lmd[k] = r->att_lmd[k];
                s1[k] = SF(R_S1 + k);
            }
            att_control(&p, &u, &ga, s1, att_ref, dot_att_ref, torque);
            for (int k = 0; k < 3; ++k) {
                torque[k] = fmin(fmax(torque[k], -r->att_saturation[k]), r->att_saturation[k]);
                SF(R_S1 + k) = s1[k];
            }
        }
        for (int k = 0; k < 3; ++k) { SF(R_AREF + k) = att_ref[k]; SF(R_DAREF + k) = dot_att_ref[k]; }
        double act4[4] = {uf, torque[0], torque[1], torque[2]};
        uav_rk44(&p, &u, act4, dis, 0);
    }
    for (int k = 0; k < 12; ++k) SF(k) = u.x[k];
    io->time[i] = u.time;
    /* is_episode_Terminal uav.py:543-560 */
    int flag = 0, done = 0;
    if (u.time > r->t_term) { flag = 1; done = 1; }
    const int pout = rob_pos_out(r, &u), aout = rob_att_out(r, &u);
    if (pout) { flag = 2; done = 1; }
    if (aout) { flag = 3; done = 1; }
    rob_observe(r, &p, &u, io, n, i, nxt);
    /* rewards */
    double reward;
    const double Qx = r->Qx, Qv = r->Qv, R = r->R;
    if (V == 0 || V == 1) { /* UavHoverOuterLoop.py:94-109 / UavHover.py:116-131 */
        double e[3], v[3] = {u.x[3], u.x[4], u.x[5]};
        for (int k = 0; k < 3; ++k) e[k] = u.x[k] - SF(R_PREF + k);
        double r1 = -sqn_tanh10(e, 3) * 0.5 * Qx - sqn(e, 3) * 0.5 * Qx;
        double r2 = -sqn_tanh10(v, 3) * 0.5 * Qx - sqn(v, 3) * 0.5 * Qv;
        double r3 = -sqn(a, AD) * R, r4 = 0;
        if (pout || aout) r4 = -(r->time_max - u.time) / p.dt * (Qx * sqn(e, 3) + Qv * sqn(v, 3) + R * sqn(a, AD));
        reward = r1 + r2 + r3 + r4;
    } else { /* UavInnerLoop.py:101-120 / UavTrackingOuterLoop.py:103-120 */
        double e[3], de[3];
        const double sc_e = V == 2 ? 1.0 : 1.0;
        (void)sc_e;
        /* error / dot_error = un-normalised observation terms */
        double f1[3][3], d1[3];
        uav_f1(&u, f1);
        mat3_vec(f1, u.x + 9, d1);
        double A[3], T[3], ph[3], z3[3] = {0, 0, 0}, ref[3], dref[3], ddref[3];
        for (int k = 0; k < 3; ++k) { A[k] = SF(R_AMP + k); T[k] = SF(R_PER + k); ph[k] = SF(R_PHS + k); }
        ref_cmd(3, u.time, A, T, V == 2 ? z3 : r->ref_bias_a, ph, ref, dref, ddref);
        for (int k = 0; k < 3; ++k) {
            e[k] = (V == 2 ? u.x[6 + k] : u.x[k]) - ref[k];
            de[k] = (V == 2 ? d1[k] : u.x[3 + k]) - dref[k];
        }
        double r1 = -sqn(e, 3) * Qx, r2 = -sqn(de, 3) * Qv;
        r1 -= sqn_tanh10(e, 3) * Qx;
        r2 -= sqn_tanh10(de, 3) * Qv;
        double r3 = -sqn(a, 3) * R, r4 = 0;
        if ((V == 2) ? aout : (pout || aout)) r4 = (r->time_max - u.time) / p.dt * (r1 + r2 + r3);
        reward = r1 + r2 + r3 + r4;
    }
    for (int k = 0; k < S; ++k) {
        if (io->obs) io->obs[(int64_t)k * n + i] = cur[k];
        io->next_obs[(int64_t)k * n + i] = nxt[k];
    }
    io->reward[i] = reward; io->done[i] = (uint8_t)done; io->flag[i] = flag;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        rob_reset(r, io, n, i, seed, off);
        for (int k = 0; k < 12; ++k) u.x[k] = SF(k);
        u.time = 0.;
        rob_observe(r, &p, &u, io, n, i, nxt);
    }
    if (io->reset_obs) for (int k = 0; k < S; ++k) io->reset_obs[(int64_t)k * n + i] = nxt[k];
}

void orc_uavrobust_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const RP *r = (const RP *)params;
    P p;
    rp_shim(r, &p);
    if (!observe_only) rob_reset(r, io, n, i, seed, off);
    if (io->next_obs) {
        uav_t u;
        for (int k = 0; k < 12; ++k) u.x[k] = SF(k);
        u.time = io->time[i];
        double o[12];
        rob_observe(r, &p, &u, io, n, i, o);
        for (int k = 0; k < (r->variant == 1 ? 12 : 6); ++k) io->next_obs[(int64_t)k * n + i] = o[k];
    }
}
