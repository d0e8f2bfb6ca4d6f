/* ugvo.c -- restatement of UGVForwardObstacleAvoidance (fake-laser ray cast against circular obstacles).
 * TEST INFRASTRUCTURE (see oracle.h).
 *   environment/UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py (variant 0) and the PPO2/DPPO2 demo copies
 *   (variant 1); map generation map.py:65-80,120-174 with the bounded, counter-indexed draws documented in b200env.h.
 */
#include <math.h>
#include <string.h>
#include "oracle.h"
#include "philox.h"

typedef b200_ugvo_params P;
#define MAXO B200_UGVO_MAX_OBS
#define SF(f) io->state[(int64_t)(f) * n + i]

typedef struct {
    double x, y, vel, phi, omega, tx, ty;
    int nobs;
    double cx[MAXO], cy[MAXO], r[MAXO];
    double time;
} ugvo_t;

/* np.linalg.norm of a 2-vector = sqrt(x.dot(x)): OpenBLAS' ddot accumulates with FMA, i.e. sqrt(fma(b, b, a * a)) -- pinned by
 * the collision-radius-equality fixture tests/golden/ugvo_edge.npz (plain a * a + b * b flips 2 of 180 decisions) */
static double norm2(double a, double b) { return sqrt(fma(b, b, a * a)); }

/* utils/functions.py:35-46 */
static double cal_vector_rad(double x1, double y1, double x2, double y2) {
    if (norm2(x2, y2) < 1e-4 || norm2(x1, y1) < 1e-4) return 0;
    double c = (x1 * x2 + y1 * y2) / (norm2(x1, y1) * norm2(x2, y2));
    c = fmin(fmax(c, -1), 1);
    return acos(c);
}
/* utils/functions.py:49-60 */
static double cal_vector_rad_oriented(double x1, double y1, double x2, double y2) {
    if (norm2(x2, y2) < 1e-4 || norm2(x1, y1) < 1e-4) return 0;
    return atan2(x1 * y2 - y1 * x2, x1 * x2 + y1 * y2);
}

/* :261-272 */
static int collision_check(const P *p, const ugvo_t *e) {
    for (int k = 0; k < e->nobs; ++k)
        if (norm2(e->x - e->cx[k], e->y - e->cy[k]) <= e->r[k] + p->r_vehicle) return 1;
    return 0;
}

/* get_fake_laser :274-397 */
static void fake_laser(const P *p, const ugvo_t *e, double *laser) {
    const int NR = p->n_rays;
    const double x = e->x, y = e->y, xm = p->map_x, ym = p->map_y;
    if (collision_check(p, e)) {
        for (int k = 0; k < NR; ++k) laser[k] = p->laser_blind;
        return;
    }
    /* np.linspace(phi - R, phi + R, NR): arange * step + start, last element = stop */
    const double start_a = e->phi - p->laser_range, stop_a = e->phi + p->laser_range;
    const double step = (stop_a - start_a) / (NR - 1);
    double ref_dis[MAXO];
    int order[MAXO];
    for (int k = 0; k < e->nobs; ++k) { ref_dis[k] = norm2(x - e->cx[k], y - e->cy[k]); order[k] = k; }
    for (int a = 1; a < e->nobs; ++a) { /* argsort (stable insertion sort; ties have measure zero) */
        int v = order[a], b = a - 1;
        while (b >= 0 && ref_dis[order[b]] > ref_dis[v]) { order[b + 1] = order[b]; --b; }
        order[b + 1] = v;
    }
    const double theta1 = cal_vector_rad(1, 0, xm - x, ym - y);
    const double theta2 = cal_vector_rad(1, 0, 0 - x, ym - y);
    const double theta3 = -cal_vector_rad(1, 0, 0 - x, 0 - y);
    const double theta4 = -cal_vector_rad(1, 0, xm - x, 0 - y);
    for (int ray = 0; ray < NR; ++ray) {
        double phi = (ray == NR - 1) ? stop_a : (double)ray * step + start_a;
        if (phi > M_PI) phi -= 2 * M_PI;
        if (phi < -M_PI) phi += 2 * M_PI;
        const double m = tan(phi);
        const double b = y - m * x;
        const double sq = sqrt(1 + pow(m, 2.0));
        const double cosTheta = fabs(m) / sq, sinTheta = 1 / sq;
        double tx, ty;
        if (theta4 < phi && phi <= theta1) {
            tx = xm; ty = m * xm + b;
            double t = x + p->laser_dis / sq;
            if (t < xm) { tx = t; ty = (m >= 0) ? y + cosTheta * p->laser_dis : y - cosTheta * p->laser_dis; }
        } else if (theta1 < phi && phi <= theta2) {
            if (fabs(m) < 1e8) { tx = (ym - b) / m; ty = ym; } else { tx = x; ty = ym; }
            double t = y + fabs(m) * p->laser_dis / sq;
            if (t < ym) { ty = t; tx = (m >= 0) ? x + p->laser_dis * sinTheta : x - p->laser_dis * sinTheta; }
        } else if (theta3 < phi && phi <= theta4) {
            if (fabs(m) < 1e8) { tx = -b / m; ty = 0; } else { tx = x; ty = 0; }
            double t = y - fabs(m) * p->laser_dis / sq;
            if (t > 0) { ty = t; tx = (m >= 0) ? x - p->laser_dis * sinTheta : x + p->laser_dis * sinTheta; }
        } else {
            tx = 0; ty = b;
            double t = x - p->laser_dis / sq;
            if (t > 0) { tx = t; ty = (m >= 0) ? y - cosTheta * p->laser_dis : y + cosTheta * p->laser_dis; }
        }
        int find = 0;
        double out = 0;
        for (int j = 0; j < e->nobs; ++j) {
            const int idx = order[j];
            const double x0 = e->cx[idx], y0 = e->cy[idx], r0 = e->r[idx];
            if (ref_dis[idx] > p->laser_dis + r0) continue;
            if (fabs(m * x0 - y0 + b) / sq > r0) continue;
            if (cal_vector_rad(tx - x, ty - y, x0 - x, y0 - y) > M_PI / 2) continue;
            const double m2p1 = pow(m, 2.0) + 1;
            const double foot_x = (x0 + m * y0 - m * b) / m2p1;
            const double foot_y = (m * x0 + pow(m, 2.0) * y0 + b) / m2p1;
            const double r_dis = norm2(foot_x - x0, foot_y - y0);
            const double d = tx - x;
            const double sg = d > 0 ? 1. : (d < 0 ? -1. : 0.);
            const double cross = foot_x - sg * sqrt(pow(r0, 2.0) - pow(r_dis, 2.0)) / sqrt(m2p1);
            if (fmin(x, tx) <= cross && cross <= fmax(x, tx)) {
                find = 1;
                const double dis = fabs(cross - x) * sqrt(m2p1);
                out = dis < p->laser_blind ? p->laser_blind : dis;
                break;
            }
        }
        if (!find) {
            const double dis = norm2(x - tx, y - ty);
            if (dis > p->laser_dis) out = p->laser_dis;
            else if (p->laser_blind < dis && dis <= p->laser_dis) out = dis;
            else out = p->laser_blind;
        }
        laser[ray] = out;
    }
}

static void get_errors(const ugvo_t *e, double *err, double *ephi) { /* :503-508 */
    *err = norm2(e->tx - e->x, e->ty - e->y);
    *ephi = cal_vector_rad_oriented(cos(e->phi), sin(e->phi), e->tx - e->x, e->ty - e->y);
}

/* get_state :399-411 */
static void observe(const P *p, const ugvo_t *e, double *o) {
    double err, ephi, laser[B200_UGVO_MAX_RAYS];
    get_errors(e, &err, &ephi);
    fake_laser(p, e, laser);
    o[0] = (2 / p->e_max * err - 1) * p->static_gain;
    o[1] = (2 / p->v_max * e->vel - 1) * p->static_gain;
    o[2] = ephi / p->e_phi_max * p->static_gain;
    o[3] = e->omega / p->omega_max * p->static_gain;
    for (int k = 0; k < p->n_rays; ++k) o[4 + k] = (2 * laser[k] / p->laser_dis - 1) * p->static_gain;
}

static void load(const oracle_io *io, int64_t n, int64_t i, ugvo_t *e) {
    e->x = SF(0); e->y = SF(1); e->vel = SF(2); e->phi = SF(3); e->omega = SF(4); e->tx = SF(5); e->ty = SF(6);
    e->nobs = (int)SF(7);
    for (int k = 0; k < MAXO; ++k) { e->cx[k] = SF(8 + 3 * k); e->cy[k] = SF(9 + 3 * k); e->r[k] = SF(10 + 3 * k); }
    e->time = io->time[i];
}
static void store(const oracle_io *io, int64_t n, int64_t i, const ugvo_t *e, int with_map) {
    SF(0) = e->x; SF(1) = e->y; SF(2) = e->vel; SF(3) = e->phi; SF(4) = e->omega;
    if (with_map) {
        SF(5) = e->tx; SF(6) = e->ty; SF(7) = (double)e->nobs;
        for (int k = 0; k < MAXO; ++k) { SF(8 + 3 * k) = e->cx[k]; SF(9 + 3 * k) = e->cy[k]; SF(10 + 3 * k) = e->r[k]; }
    }
    io->time[i] = e->time;
}

static void draw2(uint64_t seed, uint64_t gid, uint32_t ep, uint32_t block, double *u0, double *u1) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), ep, block}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    orc_philox_block(ctr, key, o);
    *u0 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
    *u1 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
}
/* obstacle candidates: one block, the radius from the 22 low bits draw2 discards (csrc/ugvo.cu draw3) */
static void draw3(uint64_t seed, uint64_t gid, uint32_t ep, uint32_t block, double *u0, double *u1, double *u2) {
    uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), ep, block}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    orc_philox_block(ctr, key, o);
    *u0 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
    *u1 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
    const uint32_t low = ((o[0] & 31u) << 17) | ((o[1] & 63u) << 11) | ((o[2] & 31u) << 6) | (o[3] & 63u);
    *u2 = (double)low * (1.0 / 4194304.0);
}
static double lerp_u(double lo, double hi, double u) { return fma(hi - lo, u, lo); }

/* reset(random=True) :527-557 + Map.generate_circle_obs_training map.py:152-174 */
static void reset_env(const P *p, ugvo_t *e, uint64_t seed, uint64_t gid, uint32_t ep) {
    double u0, u1;
    draw2(seed, gid, ep, 0, &u0, &u1);
    const double lo = p->st_margin, hx = p->map_x - p->st_margin, hy = p->map_y - p->st_margin;
    const double sx = lerp_u(lo, hx, u0), sy = lerp_u(lo, hy, u1);
    double tx = sx, ty = sy;
    for (uint32_t j = 0; j < 64 && norm2(tx - sx, ty - sy) < p->safety_dis_st; ++j) { /* map.py:69-73 */
        draw2(seed, gid, ep, 1 + j, &u0, &u1);
        tx = lerp_u(lo, hx, u0); ty = lerp_u(lo, hy, u1);
    }
    int nobs = 0;
    for (int k = 0; k < p->obs_num && k < MAXO; ++k) {
        int placed = 0;
        for (uint32_t c = 0; c < 2048 && !placed; ++c) {
            const uint32_t blk = 1000u + 2048u * (uint32_t)k + c;
            double ur;
            draw3(seed, gid, ep, blk, &u0, &u1, &ur);
            const double cx = lerp_u(0., p->map_x, u0), cy = lerp_u(0., p->map_y, u1), r = lerp_u(p->r_min, p->r_max, ur);
            int legal = 1; /* map.py:129-139 */
            if (norm2(sx - cx, sy - cy) <= r + p->safety_dis_st) legal = 0;
            if (norm2(tx - cx, ty - cy) <= r + p->safety_dis_st) legal = 0;
            for (int q = 0; q < nobs && legal; ++q)
                if (norm2(e->cx[q] - cx, e->cy[q] - cy) <= e->r[q] + r + p->safety_dis_obs) legal = 0;
            if (legal) { e->cx[nobs] = cx; e->cy[nobs] = cy; e->r[nobs] = r; ++nobs; placed = 1; }
        }
        if (!placed) break;
    }
    for (int k = nobs; k < MAXO; ++k) { e->cx[k] = 0; e->cy[k] = 0; e->r[k] = 0; }
    e->nobs = nobs;
    draw2(seed, gid, ep, 100, &u0, &u1);
    e->x = sx; e->y = sy; e->tx = tx; e->ty = ty;
    e->phi = lerp_u(-M_PI, M_PI, u0);
    e->vel = 0.; e->omega = 0.;
    e->time = 0.;
}

static int is_success(const P *p, const ugvo_t *e, double err) { /* :423-429 / demo copy :421-427 */
    int b1 = fabs(err) <= 0.05, b2 = p->variant == 0 ? (fabs(e->omega) < 0.01) : 1, b3 = fabs(e->vel) < 0.01;
    return b1 && b2 && b3;
}

void orc_ugvo_step_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint32_t flags, uint64_t seed, int64_t off) {
    const P *p = (const P *)params;
    const int S = 4 + p->n_rays;
    ugvo_t e;
    load(io, n, i, &e);
    const double al = io->action[i], aa = io->action[n + i];
    double cur[4 + B200_UGVO_MAX_RAYS], nxt[4 + B200_UGVO_MAX_RAYS];
    observe(p, &e, cur);
    /* rk44 :482-501 / demo copy :488-509 */
    double s[5] = {e.x, e.y, e.vel, e.phi, e.omega}, K1[5], K2[5], K3[5], K4[5], t[5];
#define UGVO_ODE(x, K) { K[0] = p->dt * (x[2] * cos(x[3])); K[1] = p->dt * (x[2] * sin(x[3])); K[2] = p->dt * (al - p->kf * x[2]); \
                         K[3] = p->dt * x[4]; K[4] = p->dt * (aa - p->kt * x[4]); }
    UGVO_ODE(s, K1);
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K1[k] / 2;
    UGVO_ODE(t, K2);
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K2[k] / 2;
    UGVO_ODE(t, K3);
    for (int k = 0; k < 5; ++k) t[k] = s[k] + K3[k];
    UGVO_ODE(t, K4);
    for (int k = 0; k < 5; ++k) s[k] = s[k] + (K1[k] + 2 * K2[k] + 2 * K3[k] + K4[k]) / 6;
    if (p->variant == 0) {
        e.x = s[0]; e.y = s[1]; e.vel = s[2]; e.phi = s[3]; e.omega = s[4];
        if (e.vel < 0.) e.vel = 0.;
    } else if (e.vel < 0.) { /* tests the PRE-update velocity (note N9) */
        e.phi = s[3]; e.omega = s[4]; e.vel = 0.;
    } else {
        e.x = s[0]; e.y = s[1]; e.vel = s[2]; e.phi = s[3]; e.omega = s[4];
    }
    e.time += p->dt;
    if (e.phi > M_PI) e.phi -= 2 * M_PI;
    if (e.phi < -M_PI) e.phi += 2 * M_PI;
    double err, ephi;
    get_errors(&e, &err, &ephi);
    int flag = 0; /* :431-449 */
    if (e.x > p->map_x || e.x < 0 || e.y > p->map_y || e.y < 0) flag = 1;
    if (e.time > p->time_max) flag = 2;
    if (is_success(p, &e, err)) flag = 3;
    if (collision_check(p, &e)) flag = 4;
    int done = flag != 0;
    observe(p, &e, nxt);
    double reward;
    if (p->variant == 0) { /* :451-467 */
        double u_pos = -fabs(err) * p->Q_pos, u_vel = -fabs(e.vel) * p->Q_vel;
        double u_phi = err > 0.1 ? -fabs(ephi) * p->Q_phi : 0.0, u_omega = -fabs(e.omega) * p->Q_omega, u_psi = 0.;
        if (flag == 1) { double _n = (p->time_max - e.time) / p->dt; u_psi = _n * (u_pos + u_vel + u_phi + u_omega); }
        reward = u_pos + u_vel + u_phi + u_omega + u_psi;
    } else { /* demo copy :449-473 */
        double r1 = -1 - fabs(e.omega) * 0.1, r2, r3, r4;
        if (cur[0] > nxt[0] + 1e-3) r2 = 5; else if (1e-3 + cur[0] < nxt[0]) r2 = -5; else r2 = 0;
        if (fabs(cur[1]) > fabs(nxt[1]) + 1e-2) r3 = 2; else if (1e-2 + fabs(cur[1]) < fabs(nxt[1])) r3 = -2; else r3 = 0;
        if (is_success(p, &e, err)) r4 = 500; else if (flag == 4) r4 = -300; else r4 = 0;
        reward = r1 + r2 + r3 + r4;
    }
    for (int k = 0; k < S; ++k) {
        if (io->obs) io->obs[(int64_t)k * n + i] = cur[k];
        io->next_obs[(int64_t)k * n + i] = nxt[k];
    }
    io->reward[i] = reward; io->done[i] = (uint8_t)done; io->flag[i] = flag;
    int with_map = 0;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        uint32_t ep = io->episode[i];
        reset_env(p, &e, seed, (uint64_t)(off + i), ep);
        io->episode[i] = ep + 1u;
        observe(p, &e, nxt);
        with_map = 1;
    }
    if (io->reset_obs) for (int k = 0; k < S; ++k) io->reset_obs[(int64_t)k * n + i] = nxt[k];
    store(io, n, i, &e, with_map);
}

void orc_ugvo_reset_one(const void *params, const oracle_io *io, int64_t n, int64_t i, uint64_t seed, int64_t off, int observe_only) {
    const P *p = (const P *)params;
    ugvo_t e;
    load(io, n, i, &e);
    if (!observe_only) {
        uint32_t ep = io->episode[i];
        reset_env(p, &e, seed, (uint64_t)(off + i), ep);
        io->episode[i] = ep + 1u;
        store(io, n, i, &e, 1);
    }
    if (io->next_obs) {
        double o[4 + B200_UGVO_MAX_RAYS];
        observe(p, &e, o);
        for (int k = 0; k < 4 + p->n_rays; ++k) io->next_obs[(int64_t)k * n + i] = o[k];
    }
}
