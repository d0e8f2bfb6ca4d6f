// ugvo.cu -- K-UGVO: batched UGVForwardObstacleAvoidance step with the 37-ray fake laser (sm_100a).
//
// Replaces environment/UGVForwardObstacleAvoidance/UGVForwardObstacleAvoidance.py:261-557 (collision_check,
// get_fake_laser, get_state, is_Terminal, get_reward, ode, rk44, step_update, reset) and map.py:65-80,120-174 (map
// generation), plus the PPO2/DPPO2 demo copies (variant 1).  97 % of the reference's step time is the ray cast, run
// twice per step (before and after the RK4 update).
//
// ONE WARP = ONE INSTANCE.  The unicycle RK4 is computed redundantly by all lanes (it is tiny and independent of the
// laser), so both poses are known up front and the 2 x 37 rays of the two scans are cast together in 3 passes of 32
// lanes (74/96 lanes busy instead of 74/128 with a pass pair per scan).  Per pose the obstacles are ranked by centre
// distance with a lane-parallel counting sort (lane k = obstacle k, 16 shuffles) and written in rank order to shared
// memory; every ray lane then walks that list and stops at the FIRST obstacle it accepts -- the reference's
// selection rule (note N8: first accepted obstacle in centre-distance order, not the nearest crossing).  Warp votes
// (__any_sync / __all_sync / __ballot_sync) implement the collision test, the early exit and, at reset, the
// lane-parallel rejection sampling of the obstacle map (32 candidates per round, lowest legal index wins, so the
// result is identical to trying the candidates one by one).
#include "common.cuh"

namespace {

typedef b200_ugvo_params P;
constexpr int MAXO = B200_UGVO_MAX_OBS;
constexpr int WARPS = 4; // per block: 4 consecutive instances share the 32-byte sectors of the SoA arrays
constexpr unsigned FULL = 0xffffffffu;

template <typename T> __device__ __forceinline__ T shfl(T v, int src);
template <> __device__ __forceinline__ double shfl<double>(double v, int src) { return __shfl_sync(FULL, v, src); }
template <> __device__ __forceinline__ float shfl<float>(float v, int src) { return __shfl_sync(FULL, v, src); }

template <typename T>
__device__ __forceinline__ T norm2(T a, T b) { return Mth<T>::sqrt(a * a + b * b); }

// utils/functions.py:35-46
template <typename T>
__device__ __forceinline__ T vector_rad(T x1, T y1, T x2, T y2) {
    const T n1 = norm2(x1, y1), n2 = norm2(x2, y2);
    if (n2 < (T)1e-4 || n1 < (T)1e-4) return (T)0;
    const T c = Mth<T>::min(Mth<T>::max((x1 * x2 + y1 * y2) / (n1 * n2), (T)-1), (T)1);
    return Mth<T>::acos(c);
}
// `cal_vector_rad(v1, v2) > pi / 2` without the acos: true iff the clamped cosine is negative (acos is decreasing and
// acos(c) rounds to a double above fl(pi/2) exactly when c < -5e-17; the band (-5e-17, 0) is unreachable in practice)
template <typename T>
__device__ __forceinline__ bool vector_rad_obtuse(T x1, T y1, T x2, T y2) {
    const T n1 = norm2(x1, y1), n2 = norm2(x2, y2);
    if (n2 < (T)1e-4 || n1 < (T)1e-4) return false;
    return (x1 * x2 + y1 * y2) / (n1 * n2) < (T)0;
}
// utils/functions.py:49-60
template <typename T>
__device__ __forceinline__ T vector_rad_oriented(T x1, T y1, T x2, T y2) {
    if (norm2(x2, y2) < (T)1e-4 || norm2(x1, y1) < (T)1e-4) return (T)0;
    return Mth<T>::atan2(x1 * y2 - y1 * x2, x1 * x2 + y1 * y2);
}

template <typename T>
struct Pose {
    T x, y, phi;
    T th1, th2, th3, th4; // bearings of the four map corners (:304-309)
    bool collided;        // collision_check() :261-272 -> every ray returns laserBlind
};

// sorted obstacle list of one pose in shared memory
template <typename T>
struct SortedObs {
    T x0[MAXO], y0[MAXO], r0[MAXO], d[MAXO];
};

// Prepares one pose: collision flag, corner bearings, obstacles ranked by centre distance into `so`.
// lane k < nobs owns obstacle k (cx, cy, r).
template <typename T>
__device__ __forceinline__ void prepare_pose(const P &p, Pose<T> &q, int lane, int nobs, T cx, T cy, T r, SortedObs<T> &so) {
    const bool mine = lane < nobs;
    const T d = mine ? norm2(q.x - cx, q.y - cy) : (T)1e300;
    q.collided = __any_sync(FULL, mine && d <= r + (T)p.r_vehicle);
    // rank = number of obstacles strictly closer (ties: lower index first) -- np.argsort order
    int rank = 0;
#pragma unroll
    for (int j = 0; j < MAXO; ++j) {
        const T dj = shfl<T>(d, j);
        rank += (j < nobs) && (dj < d || (dj == d && j < lane));
    }
    if (mine) { so.x0[rank] = cx; so.y0[rank] = cy; so.r0[rank] = r; so.d[rank] = d; }
    // corner bearings: lanes 0..3 evaluate one acos each
    const T xm = (T)p.map_x, ym = (T)p.map_y;
    const T vx = (lane == 0 || lane == 3) ? xm - q.x : (T)0 - q.x;
    const T vy = (lane == 0 || lane == 1) ? ym - q.y : (T)0 - q.y;
    T th = (T)0;
    if (lane < 4) th = vector_rad<T>((T)1, (T)0, vx, vy);
    q.th1 = shfl<T>(th, 0);
    q.th2 = shfl<T>(th, 1);
    q.th3 = -shfl<T>(th, 2);
    q.th4 = -shfl<T>(th, 3);
    __syncwarp();
}

// One ray of get_fake_laser (:302-395) for pose q; ray index in [0, n_rays).
template <typename T>
__device__ __forceinline__ T cast_ray(const P &p, const Pose<T> &q, int ray, int nobs, const SortedObs<T> &so, bool active) {
    const T LD = (T)p.laser_dis, LB = (T)p.laser_blind;
    const T x = q.x, y = q.y, xm = (T)p.map_x, ym = (T)p.map_y;
    // np.linspace(phi - R, phi + R, n): arange * step + start, last element = stop
    const T a0 = q.phi - (T)p.laser_range, a1 = q.phi + (T)p.laser_range;
    const T step = (a1 - a0) / (T)(p.n_rays - 1);
    T phi = (ray == p.n_rays - 1) ? a1 : (T)ray * step + a0;
    if (phi > (T)M_PI) phi -= (T)(2 * M_PI);
    if (phi < (T)-M_PI) phi += (T)(2 * M_PI);
    const T m = Mth<T>::tan(phi);
    const T b = y - m * x;
    const T m2p1 = m * m + (T)1;
    const T sq = Mth<T>::sqrt(m2p1);
    const T cosT = Mth<T>::abs(m) / sq, sinT = (T)1 / sq;
    T tx, ty;
    if (q.th4 < phi && phi <= q.th1) { // right wall
        tx = xm; ty = m * xm + b;
        const T t = x + LD / sq;
        if (t < xm) { tx = t; ty = (m >= (T)0) ? y + cosT * LD : y - cosT * LD; }
    } else if (q.th1 < phi && phi <= q.th2) { // top wall
        if (Mth<T>::abs(m) < (T)1e8) { tx = (ym - b) / m; ty = ym; } else { tx = x; ty = ym; }
        const T t = y + Mth<T>::abs(m) * LD / sq;
        if (t < ym) { ty = t; tx = (m >= (T)0) ? x + LD * sinT : x - LD * sinT; }
    } else if (q.th3 < phi && phi <= q.th4) { // bottom wall
        if (Mth<T>::abs(m) < (T)1e8) { tx = -b / m; ty = (T)0; } else { tx = x; ty = (T)0; }
        const T t = y - Mth<T>::abs(m) * LD / sq;
        if (t > (T)0) { ty = t; tx = (m >= (T)0) ? x - LD * sinT : x + LD * sinT; }
    } else { // left wall
        tx = (T)0; ty = b;
        const T t = x - LD / sq;
        if (t > (T)0) { tx = t; ty = (m >= (T)0) ? y - cosT * LD : y + cosT * LD; }
    }
    const T dxs = tx - x;
    const T sg = dxs > (T)0 ? (T)1 : (dxs < (T)0 ? (T)-1 : (T)0);
    const T lo = Mth<T>::min(x, tx), hi = Mth<T>::max(x, tx);
    const T rdx = tx - x, rdy = ty - y;
    const bool ray_ok = !(norm2(rdx, rdy) < (T)1e-4); // cal_vector_rad returns 0 for a degenerate ray (never > pi/2)
    bool found = !active;
    T out = (T)0;
    for (int j = 0; j < nobs; ++j) { // obstacles in centre-distance order, first accepted wins (N8)
        if (__all_sync(FULL, found)) break;
        if (!found) {
            const T x0 = so.x0[j], y0 = so.y0[j], r0 = so.r0[j];
            const T dj = so.d[j]; // = |centre - start|, the second norm of cal_vector_rad
            bool rej = dj > LD + r0;                                         // out of range
            rej = rej || (Mth<T>::abs(m * x0 - y0 + b) / sq > r0);           // the line misses the circle
            // cal_vector_rad(ray, centre - start) > pi / 2  <=>  both vectors non-degenerate and their dot product < 0
            // (dividing by the positive norms and taking acos cannot change the sign; see vector_rad_obtuse)
            rej = rej || (ray_ok && !(dj < (T)1e-4) && (rdx * (x0 - x) + rdy * (y0 - y) < (T)0)); // behind the ray
            if (!rej) {
                const T fx = (x0 + m * y0 - m * b) / m2p1;
                const T fy = (m * x0 + m * m * y0 + b) / m2p1;
                const T rd = norm2(fx - x0, fy - y0);
                const T cross = fx - sg * Mth<T>::sqrt(r0 * r0 - rd * rd) / sq; // NaN (no crossing) fails the test below
                if (lo <= cross && cross <= hi) {
                    found = true;
                    const T dis = Mth<T>::abs(cross - x) * sq;
                    out = dis < LB ? LB : dis;
                }
            }
        }
    }
    if (active && !found) { // :382-395
        const T dis = norm2(x - tx, y - ty);
        if (dis > LD) out = LD;
        else if (LB < dis && dis <= LD) out = dis;
        else out = LB;
    } else if (!active) {
        out = (T)0;
    }
    return q.collided ? LB : out;
}

struct Draw2 { double u0, u1; };
__device__ __forceinline__ Draw2 draw2(uint64_t seed, uint64_t gid, uint32_t ep, uint32_t block) {
    Philox g(seed, gid, ep);
    g.c3 = block;
    g.block();
    Draw2 d;
    d.u0 = ((double)(g.r[0] >> 5) * 67108864.0 + (double)(g.r[1] >> 6)) * (1.0 / 9007199254740992.0);
    d.u1 = ((double)(g.r[2] >> 5) * 67108864.0 + (double)(g.r[3] >> 6)) * (1.0 / 9007199254740992.0);
    return d;
}
__device__ __forceinline__ double lerp_u(double lo, double hi, double u) { return ::fma(hi - lo, u, lo); }

// reset(random=True) :527-557 + Map.generate_circle_obs_training (map.py:152-174), cooperative over the warp.
// Outputs (uniform over the warp): start/target/phi0; per-lane obstacle k = lane (cx, cy, r), nobs.
__device__ __forceinline__ void reset_map(const P &p, uint64_t seed, uint64_t gid, uint32_t ep, int lane, double &sx,
                                          double &sy, double &tx, double &ty, double &phi0, double &ocx, double &ocy,
                                          double &orr, int &nobs) {
    const double lo = p.st_margin, hx = p.map_x - p.st_margin, hy = p.map_y - p.st_margin;
    Draw2 d = draw2(seed, gid, ep, 0);
    sx = lerp_u(lo, hx, d.u0);
    sy = lerp_u(lo, hy, d.u1);
    // target: first candidate (blocks 1..64) at least safety_dis_st away; 32 candidates per round
    tx = sx; ty = sy;
    for (int round = 0; round < 2; ++round) {
        d = draw2(seed, gid, ep, 1u + (uint32_t)(round * 32 + lane));
        const double cx = lerp_u(lo, hx, d.u0), cy = lerp_u(lo, hy, d.u1);
        const bool ok = !(sqrt((cx - sx) * (cx - sx) + (cy - sy) * (cy - sy)) < p.safety_dis_st);
        const unsigned m = __ballot_sync(FULL, ok);
        if (m) {
            const int src = __ffs(m) - 1;
            tx = __shfl_sync(FULL, cx, src);
            ty = __shfl_sync(FULL, cy, src);
            break;
        }
        if (round == 1) { // no candidate accepted: the oracle keeps the last one drawn (block 64)
            tx = __shfl_sync(FULL, cx, 31);
            ty = __shfl_sync(FULL, cy, 31);
        }
    }
    ocx = 0.0; ocy = 0.0; orr = 0.0;
    nobs = 0;
    const int want = p.obs_num < MAXO ? p.obs_num : MAXO;
    for (int k = 0; k < want; ++k) {
        bool placed = false;
        for (int round = 0; round < 64 && !placed; ++round) {
            const uint32_t c = (uint32_t)(round * 32 + lane);
            const uint32_t blk = 1000u + 2u * (2048u * (uint32_t)k + c);
            d = draw2(seed, gid, ep, blk);
            const Draw2 dr = draw2(seed, gid, ep, blk + 1u);
            const double cx = lerp_u(0., p.map_x, d.u0), cy = lerp_u(0., p.map_y, d.u1), r = lerp_u(p.r_min, p.r_max, dr.u0);
            bool legal = true; // map.py:129-139
            if (sqrt((sx - cx) * (sx - cx) + (sy - cy) * (sy - cy)) <= r + p.safety_dis_st) legal = false;
            if (sqrt((tx - cx) * (tx - cx) + (ty - cy) * (ty - cy)) <= r + p.safety_dis_st) legal = false;
            for (int q = 0; q < nobs; ++q) {
                const double qx = __shfl_sync(FULL, ocx, q), qy = __shfl_sync(FULL, ocy, q), qr = __shfl_sync(FULL, orr, q);
                if (sqrt((qx - cx) * (qx - cx) + (qy - cy) * (qy - cy)) <= qr + r + p.safety_dis_obs) legal = false;
            }
            const unsigned m = __ballot_sync(FULL, legal);
            if (m) {
                const int src = __ffs(m) - 1;
                const double wx = __shfl_sync(FULL, cx, src), wy = __shfl_sync(FULL, cy, src), wr = __shfl_sync(FULL, r, src);
                if (lane == nobs) { ocx = wx; ocy = wy; orr = wr; }
                ++nobs;
                placed = true;
            }
        }
        if (!placed) break;
    }
    d = draw2(seed, gid, ep, 100);
    phi0 = lerp_u(-M_PI, M_PI, d.u0);
}

enum { F_X = 0, F_Y, F_VEL, F_PHI, F_OMEGA, F_TX, F_TY, F_NOBS, F_OBS };

// shared body of step / reset / observe.  mode 0 = step, 1 = reset (masked), 2 = observe only
template <typename T, bool IO32>
__global__ void __launch_bounds__(WARPS * 32)
ugvo_kernel(const __grid_constant__ P p, const __grid_constant__ b200env_io io, int64_t n, uint32_t flags,
            uint64_t seed, int64_t off, const uint8_t *mask, int mode) {
    __shared__ SortedObs<T> s_obs[WARPS][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * WARPS + w;
    if (i >= n) return; // whole warp exits together
    if (mode == 1 && mask && !mask[i]) return;
    const int NR = p.n_rays;

    // ---- load: lanes 0..7 fetch the scalar fields, lane k the k-th obstacle; scalars are broadcast
    T sc = (T)0;
    if (lane < F_OBS) sc = ld<T>(io.state, n, lane, i);
    T x = shfl<T>(sc, F_X), y = shfl<T>(sc, F_Y), vel = shfl<T>(sc, F_VEL), phi = shfl<T>(sc, F_PHI);
    T omega = shfl<T>(sc, F_OMEGA), tgx = shfl<T>(sc, F_TX), tgy = shfl<T>(sc, F_TY);
    int nobs = (int)shfl<T>(sc, F_NOBS);
    T ocx = (T)0, ocy = (T)0, orr = (T)0;
    if (lane < MAXO) {
        ocx = ld<T>(io.state, n, F_OBS + 3 * lane + 0, i);
        ocy = ld<T>(io.state, n, F_OBS + 3 * lane + 1, i);
        orr = ld<T>(io.state, n, F_OBS + 3 * lane + 2, i);
    }
    double time = io.time[i];
    bool store_map = false;

    auto do_reset = [&]() {
        const uint32_t ep = io.episode[i];
        double sx, sy, ttx, tty, phi0, cx, cy, rr;
        int no;
        reset_map(p, seed, (uint64_t)(off + i), ep, lane, sx, sy, ttx, tty, phi0, cx, cy, rr, no);
        x = (T)sx; y = (T)sy; tgx = (T)ttx; tgy = (T)tty; phi = (T)phi0; vel = (T)0; omega = (T)0;
        ocx = (T)cx; ocy = (T)cy; orr = (T)rr; nobs = no;
        time = 0.0;
        if (lane == 0) io.episode[i] = ep + 1u;
        store_map = true;
    };
    // observation of the current (x, y, vel, phi, omega): get_state :399-411; uses pose slot `slot`
    auto observe_into = [&](void *dst, int slot) {
        Pose<T> q;
        q.x = x; q.y = y; q.phi = phi;
        prepare_pose<T>(p, q, lane, nobs, ocx, ocy, orr, s_obs[w][slot]);
        const T g = (T)p.static_gain;
        for (int pass = 0; pass * 32 < NR; ++pass) {
            const int ray = pass * 32 + lane;
            const T l = cast_ray<T>(p, q, ray < NR ? ray : 0, nobs, s_obs[w][slot], ray < NR);
            if (ray < NR && dst) stio<T, IO32>(dst, n, 4 + ray, i, ((T)2 * l / (T)p.laser_dis - (T)1) * g);
        }
        if (lane == 0 && dst) {
            T s, c;
            Mth<T>::sincos(phi, &s, &c);
            const T e = norm2(tgx - x, tgy - y), ephi = vector_rad_oriented<T>(c, s, tgx - x, tgy - y);
            stio<T, IO32>(dst, n, 0, i, ((T)(2 / p.e_max) * e - (T)1) * g);
            stio<T, IO32>(dst, n, 1, i, ((T)(2 / p.v_max) * vel - (T)1) * g);
            stio<T, IO32>(dst, n, 2, i, ephi / (T)p.e_phi_max * g);
            stio<T, IO32>(dst, n, 3, i, omega / (T)p.omega_max * g);
        }
        __syncwarp();
    };
    auto store_state = [&]() {
        if (lane == 0) {
            st<T>(io.state, n, F_X, i, x); st<T>(io.state, n, F_Y, i, y); st<T>(io.state, n, F_VEL, i, vel);
            st<T>(io.state, n, F_PHI, i, phi); st<T>(io.state, n, F_OMEGA, i, omega);
            io.time[i] = time;
            if (store_map) {
                st<T>(io.state, n, F_TX, i, tgx); st<T>(io.state, n, F_TY, i, tgy); st<T>(io.state, n, F_NOBS, i, (T)nobs);
            }
        }
        if (store_map && lane < MAXO) {
            st<T>(io.state, n, F_OBS + 3 * lane + 0, i, ocx);
            st<T>(io.state, n, F_OBS + 3 * lane + 1, i, ocy);
            st<T>(io.state, n, F_OBS + 3 * lane + 2, i, orr);
        }
    };

    if (mode != 0) { // reset / observe
        if (mode == 1) { do_reset(); store_state(); }
        if (io.next_obs) observe_into(io.next_obs, 0);
        return;
    }

    // ---- step_update :510-520
    T al = (T)0;
    if (lane < 2) al = ldio<T, IO32>(io.action, n, lane, i);
    const T a_lin = shfl<T>(al, 0), a_ang = shfl<T>(al, 1);
    Pose<T> qa, qb;
    qa.x = x; qa.y = y; qa.phi = phi;
    const T cur_vel = vel;
    T cur_e;
    {
        const T e0 = norm2(tgx - x, tgy - y);
        cur_e = ((T)(2 / p.e_max) * e0 - (T)1) * (T)p.static_gain; // current_state[0], used by the demo-copy reward
    }
    if (io.obs && lane == 0) { // kinematic part of current_state
        T s, c;
        Mth<T>::sincos(phi, &s, &c);
        const T g = (T)p.static_gain;
        stio<T, IO32>(io.obs, n, 0, i, cur_e);
        stio<T, IO32>(io.obs, n, 1, i, ((T)(2 / p.v_max) * vel - (T)1) * g);
        stio<T, IO32>(io.obs, n, 2, i, vector_rad_oriented<T>(c, s, tgx - x, tgy - y) / (T)p.e_phi_max * g);
        stio<T, IO32>(io.obs, n, 3, i, omega / (T)p.omega_max * g);
    }
    // rk44 :482-501 / demo copy :488-509 (all lanes, redundantly)
    {
        const T h = (T)p.dt, half = (T)0.5, kf = (T)p.kf, kt = (T)p.kt;
        T s, c;
        Mth<T>::sincos(phi, &s, &c);
        const T k1x = h * (vel * c), k1y = h * (vel * s), k1v = h * (a_lin - kf * vel), k1p = h * omega, k1o = h * (a_ang - kt * omega);
        const T v2 = vel + k1v * half, o2 = omega + k1o * half;
        Mth<T>::sincos(phi + k1p * half, &s, &c);
        const T k2x = h * (v2 * c), k2y = h * (v2 * s), k2v = h * (a_lin - kf * v2), k2p = h * o2, k2o = h * (a_ang - kt * o2);
        const T v3 = vel + k2v * half, o3 = omega + k2o * half;
        Mth<T>::sincos(phi + k2p * half, &s, &c);
        const T k3x = h * (v3 * c), k3y = h * (v3 * s), k3v = h * (a_lin - kf * v3), k3p = h * o3, k3o = h * (a_ang - kt * o3);
        const T v4 = vel + k3v, o4 = omega + k3o;
        Mth<T>::sincos(phi + k3p, &s, &c);
        const T k4x = h * (v4 * c), k4y = h * (v4 * s), k4v = h * (a_lin - kf * v4), k4p = h * o4, k4o = h * (a_ang - kt * o4);
        const T nx = x + (k1x + (T)2 * k2x + (T)2 * k3x + k4x) / (T)6;
        const T ny = y + (k1y + (T)2 * k2y + (T)2 * k3y + k4y) / (T)6;
        const T nv = vel + (k1v + (T)2 * k2v + (T)2 * k3v + k4v) / (T)6;
        const T np_ = phi + (k1p + (T)2 * k2p + (T)2 * k3p + k4p) / (T)6;
        const T no = omega + (k1o + (T)2 * k2o + (T)2 * k3o + k4o) / (T)6;
        if (p.variant == 0) {
            x = nx; y = ny; vel = nv; phi = np_; omega = no;
            if (vel < (T)0) vel = (T)0;
        } else if (vel < (T)0) { // the PRE-update velocity is tested (note N9): pose frozen
            phi = np_; omega = no; vel = (T)0;
        } else {
            x = nx; y = ny; vel = nv; phi = np_; omega = no;
        }
        time += p.dt;
        if (phi > (T)M_PI) phi -= (T)(2 * M_PI);
        if (phi < (T)-M_PI) phi += (T)(2 * M_PI);
    }
    qb.x = x; qb.y = y; qb.phi = phi;
    // ---- is_Terminal :431-449 (uniform over the warp; needs only the new pose and its collision flag)
    const bool scan_a = io.obs != nullptr;
    if (scan_a) prepare_pose<T>(p, qa, lane, nobs, ocx, ocy, orr, s_obs[w][0]);
    prepare_pose<T>(p, qb, lane, nobs, ocx, ocy, orr, s_obs[w][1]);
    T s, c;
    Mth<T>::sincos(phi, &s, &c);
    const T err = norm2(tgx - x, tgy - y), ephi = vector_rad_oriented<T>(c, s, tgx - x, tgy - y);
    const bool succ = Mth<T>::abs(err) <= (T)0.05 && (p.variant != 0 || Mth<T>::abs(omega) < (T)0.01) &&
                      Mth<T>::abs(vel) < (T)0.01;
    int flag = 0;
    if (x > (T)p.map_x || x < (T)0 || y > (T)p.map_y || y < (T)0) flag = 1;
    if (time > p.time_max) flag = 2;
    if (succ) flag = 3;
    if (qb.collided) flag = 4;
    const bool done = flag != 0;
    const bool will_reset = done && (flags & B200ENV_AUTO_RESET);
    void *mirror = (!will_reset && io.reset_obs) ? io.reset_obs : nullptr; // policy-facing obs = next_obs unless reset
    const T g = (T)p.static_gain;
    // ---- both laser scans together: rays 0..NR-1 from pose A (into obs), NR..2NR-1 from pose B (into next_obs)
    {
        const int first = scan_a ? 0 : NR, total = 2 * NR;
        for (int base = first; base < total; base += 32) {
            const int r = base + lane;
            const bool active = r < total;
            const bool is_b = r >= NR;
            const int ray = active ? (is_b ? r - NR : r) : 0;
            // the two poses differ per lane: select the pose data, then one common cast
            Pose<T> q;
            q.x = is_b ? qb.x : qa.x; q.y = is_b ? qb.y : qa.y; q.phi = is_b ? qb.phi : qa.phi;
            q.th1 = is_b ? qb.th1 : qa.th1; q.th2 = is_b ? qb.th2 : qa.th2;
            q.th3 = is_b ? qb.th3 : qa.th3; q.th4 = is_b ? qb.th4 : qa.th4;
            q.collided = is_b ? qb.collided : qa.collided;
            const T l = cast_ray<T>(p, q, ray, nobs, s_obs[w][is_b ? 1 : 0], active);
            if (active) {
                const T v = ((T)2 * l / (T)p.laser_dis - (T)1) * g;
                stio<T, IO32>(is_b ? io.next_obs : io.obs, n, 4 + ray, i, v);
                if (is_b && mirror) stio<T, IO32>(mirror, n, 4 + ray, i, v);
            }
        }
    }
    // ---- get_reward :451-467 / demo copy :449-473
    const T nxt0 = ((T)(2 / p.e_max) * err - (T)1) * g, nxt1 = ((T)(2 / p.v_max) * vel - (T)1) * g;
    T reward;
    if (p.variant == 0) {
        const T u_pos = -Mth<T>::abs(err) * (T)p.Q_pos, u_vel = -Mth<T>::abs(vel) * (T)p.Q_vel;
        const T u_phi = err > (T)0.1 ? -Mth<T>::abs(ephi) * (T)p.Q_phi : (T)0;
        const T u_omega = -Mth<T>::abs(omega) * (T)p.Q_omega;
        T u_psi = (T)0;
        if (flag == 1) u_psi = (T)((p.time_max - time) / p.dt) * (u_pos + u_vel + u_phi + u_omega);
        reward = u_pos + u_vel + u_phi + u_omega + u_psi;
    } else {
        const T cur1 = ((T)(2 / p.v_max) * cur_vel - (T)1) * g;
        const T r1 = (T)-1 - Mth<T>::abs(omega) * (T)0.1;
        const T r2 = cur_e > nxt0 + (T)1e-3 ? (T)5 : ((T)1e-3 + cur_e < nxt0 ? (T)-5 : (T)0);
        const T r3 = Mth<T>::abs(cur1) > Mth<T>::abs(nxt1) + (T)1e-2 ? (T)2
                     : ((T)1e-2 + Mth<T>::abs(cur1) < Mth<T>::abs(nxt1) ? (T)-2 : (T)0);
        const T r4 = succ ? (T)500 : (flag == 4 ? (T)-300 : (T)0);
        reward = r1 + r2 + r3 + r4;
    }
    if (lane == 0) {
        const T o2 = ephi / (T)p.e_phi_max * g, o3 = omega / (T)p.omega_max * g;
        stio<T, IO32>(io.next_obs, n, 0, i, nxt0);
        stio<T, IO32>(io.next_obs, n, 1, i, nxt1);
        stio<T, IO32>(io.next_obs, n, 2, i, o2);
        stio<T, IO32>(io.next_obs, n, 3, i, o3);
        if (mirror) { stio<T, IO32>(mirror, n, 0, i, nxt0); stio<T, IO32>(mirror, n, 1, i, nxt1); stio<T, IO32>(mirror, n, 2, i, o2); stio<T, IO32>(mirror, n, 3, i, o3); }
        stio<T, IO32>(io.reward, n, 0, i, reward);
        io.done[i] = done ? 1 : 0;
        io.flag[i] = flag;
    }
    __syncwarp();
    if (will_reset) {
        do_reset();
        if (io.reset_obs) observe_into(io.reset_obs, 0);
    }
    store_state();
}

int launch(int dtype, int64_t n, const void *params, const b200env_io *io, uint32_t flags, uint64_t seed, int64_t off,
           const uint8_t *mask, int mode, cudaStream_t s) {
    const P &p = *static_cast<const P *>(params);
    if (p.n_rays < 2 || p.n_rays > B200_UGVO_MAX_RAYS || p.obs_num < 0 || p.obs_num > MAXO) return B200ENV_EPARAMS;
    const unsigned grid = (unsigned)((n + WARPS - 1) / WARPS);
    B200_LAUNCH_TIO(ugvo_kernel, grid, WARPS * 32, s, p, *io, n, flags, seed, off, mask, mode);
    return b200_check_launch();
}

} // namespace

int ugvo_dims(int variant, int *sf, int *od, int *ad, int *dd) {
    if (variant != 0 && variant != 1) return B200ENV_EENV;
    if (sf) *sf = B200_UGVO_STATE_FIELDS;
    if (od) *od = 4 + B200_UGVO_MAX_RAYS;
    if (ad) *ad = 2;
    if (dd) *dd = 0;
    return B200ENV_OK;
}
int ugvo_step(int dtype, int64_t n, const void *params, const b200env_io *io, uint32_t flags, uint64_t seed, int64_t off,
              cudaStream_t s) {
    if (!io->state || !io->time || !io->action || !io->next_obs || !io->reward || !io->done || !io->flag) return B200ENV_ENULL;
    if ((flags & B200ENV_AUTO_RESET) && !io->episode) return B200ENV_ENULL;
    return launch(dtype, n, params, io, flags, seed, off, nullptr, 0, s);
}
int ugvo_reset(int dtype, int64_t n, const void *params, const b200env_io *io, const uint8_t *mask, uint64_t seed,
               int64_t off, cudaStream_t s) {
    if (!io->state || !io->time || !io->episode) return B200ENV_ENULL;
    return launch(dtype, n, params, io, 0, seed, off, mask, 1, s);
}
int ugvo_observe(int dtype, int64_t n, const void *params, const b200env_io *io, cudaStream_t s) {
    if (!io->state || !io->time || !io->next_obs) return B200ENV_ENULL;
    return launch(dtype, n, params, io, 0, 0, 0, nullptr, 2, s);
}
