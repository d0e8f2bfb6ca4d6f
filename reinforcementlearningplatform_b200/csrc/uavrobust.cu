// uavrobust.cu -- K-UAVR: the four UavRobust RL wrappers around the quadrotor (sm_100a), one thread per instance.
//   variant 0 uav_hover_outer_loop     environment/UavRobust/UavHoverOuterLoop.py:81-196
//   variant 1 uav_hover                environment/UavRobust/UavHover.py:101-226
//   variant 2 uav_inner_loop           environment/UavRobust/UavInnerLoop.py:88-210
//   variant 3 uav_tracking_outer_loop  environment/UavRobust/UavTrackingOuterLoop.py:90-255
// Shared pieces: quadrotor ODE / RK4 / terminal tests (UavRobust/uav.py:429-560, same maths as UavFntsmcParam),
// throttle / reference-angle mapping (uav_pos_ctrl.py:67-76), inner FNTSMC with torque saturation (FNTSMC.py:80-106).
// The variant is a template parameter, so each wrapper compiles to its own straight-line kernel.
#include "uav_common.cuh"

namespace {
using namespace uavk;
typedef b200_uavrobust_params RP;
enum { R_S1 = 12, R_AREF = 15, R_DAREF = 18, R_PREF = 21, R_AMP = 24, R_PER = 27, R_PHS = 30 };

template <typename T> __device__ __forceinline__ T sq3(const T *v) { // np.linalg.norm(v) ** 2
    const T n = Mth<T>::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    return n * n;
}
template <typename T> __device__ __forceinline__ T sq3_tanh10(const T *v) { // np.linalg.norm(np.tanh(10 * v)) ** 2
    const T t[3] = {Mth<T>::tanh((T)10 * v[0]), Mth<T>::tanh((T)10 * v[1]), Mth<T>::tanh((T)10 * v[2])};
    return sq3<T>(t);
}

// Reciprocals of the observation normalisers, computed once on the host (launcher): the 36 IEEE divisions of the three
// get_state() evaluations per step were 19 % of the kernel's instructions (ncu source page).  a / d is evaluated as
// q = a r, q + (a - q d) r (Markstein: the IEEE quotient except in rare half-way cases, <= 1 ulp).
struct RobDerived {
    double r_pos[3], r_vel[3], r_att[3], r_datt[3];
};
static inline RobDerived rob_derive(const RP &r) {
    RobDerived d;
    for (int k = 0; k < 3; ++k) {
        d.r_pos[k] = 1.0 / r.e_pos_span[k];
        d.r_vel[k] = 1.0 / r.vel_span[k];
        d.r_att[k] = 1.0 / r.e_att_span[k];
        d.r_datt[k] = 1.0 / r.e_dot_att_span_neg[k];
    }
    return d;
}
template <typename T>
__device__ __forceinline__ T qdiv(T a, double d, double rcp) {
    const T q = a * (T)rcp;
    return Mth<T>::fma(Mth<T>::fma(-q, (T)d, a), (T)rcp, q);
}

// get_state of the four wrappers.  x: 12 states, t: trig of the attitude, time: self.time
template <typename T, int V>
__device__ __forceinline__ void observe(const RP &r, const RobDerived &rd, const b200env_io &io, int64_t n, int64_t i,
                                        const T *x, const Trig<T> &t, double time, T *o, T *e_out, T *de_out) {
    const T g = (T)r.static_gain;
    T d1[3];
    d1[0] = x[9] + (t.sphi * t.tth) * x[10] + (t.cphi * t.tth) * x[11]; // dot_rho1 = f1 . pqr
    d1[1] = t.cphi * x[10] - t.sphi * x[11];
    d1[2] = (t.sphi * t.rcth) * x[10] + (t.cphi * t.rcth) * x[11];
    if (V == 0 || V == 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e_out[k] = x[k] - UAV_LDS(io.state, n, R_PREF + k, i);
            de_out[k] = x[3 + k];
            o[k] = qdiv<T>(e_out[k], r.e_pos_span[k], rd.r_pos[k]) * g;
            o[3 + k] = qdiv<T>((T)2 * x[3 + k], r.vel_span[k], rd.r_vel[k]) * g;
        }
        if (V == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                o[6 + k] = qdiv<T>(x[6 + k] - UAV_LDS(io.state, n, R_AREF + k, i), r.e_att_span[k], rd.r_att[k]) * g;
                o[9 + k] = qdiv<T>(d1[k] - UAV_LDS(io.state, n, R_DAREF + k, i), r.e_dot_att_span_neg[k], rd.r_datt[k]) * g; // sic (N10)
            }
        }
    } else {
        T rA[3], rT[3], rB[3], rP[3], refs[3], drefs[3], ddrefs[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            rA[k] = UAV_LDS(io.state, n, R_AMP + k, i); rT[k] = UAV_LDS(io.state, n, R_PER + k, i);
            rB[k] = V == 2 ? (T)0 : (T)r.ref_bias_a[k]; rP[k] = UAV_LDS(io.state, n, R_PHS + k, i);
        }
        ref_channels<T, 3>((T)time, rA, rT, rB, rP, refs, drefs, ddrefs); // equal (period, phase) share the sincos
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T ref = refs[k], dref = drefs[k];
            if (V == 2) {
                e_out[k] = x[6 + k] - ref; de_out[k] = d1[k] - dref;
                o[k] = qdiv<T>(e_out[k], r.e_att_span[k], rd.r_att[k]) * g;
                o[3 + k] = qdiv<T>(de_out[k], r.e_dot_att_span_neg[k], rd.r_datt[k]) * g; // sic (N10)
            } else {
                e_out[k] = x[k] - ref; de_out[k] = x[3 + k] - dref;
                o[k] = qdiv<T>(e_out[k], r.e_pos_span[k], rd.r_pos[k]) * g;
                o[3 + k] = qdiv<T>(de_out[k], r.vel_span[k], rd.r_vel[k]) * g;
            }
        }
    }
}

template <typename T, int V>
__device__ __forceinline__ void reset_state(const RP &r, const b200env_io &io, int64_t n, int64_t i, uint64_t seed,
                                            int64_t off, T *x) {
    const uint32_t ep = io.episode[i];
    Philox rng(seed, (uint64_t)(off + i), ep);
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k] = (T)r.pos0[k]; x[3 + k] = (T)r.vel0[k]; x[6 + k] = (T)r.angle0[k]; x[9 + k] = (T)r.pqr0[k]; }
    if (V == 0 || V == 1) { // generate_random_point(offset = 1.0)
#pragma unroll
        for (int k = 0; k < 3; ++k) UAV_STS(io.state, n, R_PREF + k, i, (T)rng.uniform(r.target_lo[k], r.target_hi[k]));
        if (V == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) UAV_STS(io.state, n, R_AREF + k, i, (T)0); // UavHover.py:209
        }
    } else {
        double A[3], Tp[3], ph[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) A[k] = rng.uniform(0., r.sig_A_hi[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) Tp[k] = rng.uniform(r.sig_T_lo, r.sig_T_hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) ph[k] = rng.uniform(0., r.sig_phase_hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            UAV_STS(io.state, n, R_AMP + k, i, (T)A[k]);
            UAV_STS(io.state, n, R_PER + k, i, (T)Tp[k]);
            UAV_STS(io.state, n, R_PHS + k, i, (T)ph[k]);
        }
        if (V == 3) { // set_random_init_pos(trajectory[0], 0.3)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double t0 = r.ref_bias_a[k] + A[k] * sin(2 * M_PI / Tp[k] * 0. + ph[k]);
                x[k] = (T)rng.uniform(t0 - fabs(r.init_pos_r), t0 + fabs(r.init_pos_r));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) UAV_STS(io.state, n, k, i, x[k]);
    // s1 (and att_ref / dot_att_ref except where noted) survive the reference's reset()
    io.time[i] = 0.0;
    io.episode[i] = ep + 1u;
}

template <typename T, int V, bool IO32>
#ifndef UAVR_MINBLOCKS
#define UAVR_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(B200_BLOCK, UAVR_MINBLOCKS)
uavrobust_step_kernel(const __grid_constant__ RP r, const __grid_constant__ UavDerived dv,
                      const __grid_constant__ RobDerived rd, const __grid_constant__ b200env_io io, int64_t n, uint32_t flags, uint64_t seed, int64_t off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int S = V == 1 ? 12 : 6, AD = V == 1 ? 6 : 3;
    const Consts<T> c(r.m, r.g, r.J, r.kr, r.kt, r.dt, dv);
    T x[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) x[k] = UAV_LDS(io.state, n, k, i);
    double time = io.time[i];
    T a[AD], dis[3] = {(T)0, (T)0, (T)0};
#pragma unroll
    for (int k = 0; k < AD; ++k) a[k] = ldio<T, IO32>(io.action, n, k, i);
    if (V != 2 && io.dis) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dis[k] = ldio<T, IO32>(io.dis, n, k, i);
    }
    T s1[3], aref_old[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s1[k] = (V == 0 || V == 3) ? UAV_LDS(io.state, n, R_S1 + k, i) : (T)0;
        aref_old[k] = (V != 2) ? UAV_LDS(io.state, n, R_AREF + k, i) : (T)0;
    }
    Trig<T> t1;
    t1.eval(x[6], x[7], x[8], V != 2);
    T cur[S], nxt[S], e[3], de[3];
    observe<T, V>(r, rd, io, n, i, x, t1, time, cur, e, de); // current_state = get_state()
    if (io.obs) {
#pragma unroll
        for (int k = 0; k < S; ++k) stio<T, IO32>(io.obs, n, k, i, cur[k]);
    }
    T torque[3], uf = (T)0;
    if (V == 2) { // UavInnerLoop.py:127-133: the action is the torque; throttle 0, attitude only
#pragma unroll
        for (int k = 0; k < 3; ++k) torque[k] = a[k];
        uav_rk44<T, true>(c, x, t1, (T)0, torque, dis);
#pragma unroll
        for (int k = 0; k < 6; ++k) x[k] = (T)0;
    } else {
        // uo_2_ref_angle_throttle, uav_pos_ctrl.py:67-76 (+ np.clip to the attitude zone)
        uf = (a[2] + c.g) * c.m / (t1.cphi * t1.cth);
        const T asin_phi_d = clampc<T>(Mth<T>::div((a[0] * t1.spsi - a[1] * t1.cpsi) * c.m, uf), (T)-1, (T)1);
        T phi_d = Mth<T>::asin(asin_phi_d);
        const T asin_theta_d = clampc<T>(
            Mth<T>::div((a[0] * t1.cpsi + a[1] * t1.spsi) * c.m, uf * cos_of_asin<T>(asin_phi_d, phi_d)), (T)-1, (T)1);
        T theta_d = Mth<T>::asin(asin_theta_d);
        phi_d = clampc<T>(phi_d, (T)r.att_zone_min[0], (T)r.att_zone_max[0]);
        theta_d = clampc<T>(theta_d, (T)r.att_zone_min[1], (T)r.att_zone_max[1]);
        // reference shaping: rate-limited attitude command (UavHoverOuterLoop.py:127-131)
        const T att_new[3] = {phi_d, theta_d, (T)0};
        T att_ref[3], datt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T d = (att_new[k] - aref_old[k]) / c.dt;
            d = clampc<T>(d, (T)r.dot_att_min[k], (T)r.dot_att_max[k]);
            datt[k] = d;
            att_ref[k] = d * c.dt + aref_old[k];
            UAV_STS(io.state, n, R_AREF + k, i, att_ref[k]);
            UAV_STS(io.state, n, R_DAREF + k, i, d);
        }
        if (V == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) torque[k] = a[3 + k];
        } else { // inner FNTSMC with saturation: uav_pos_ctrl.py:41-65, FNTSMC.py:80-106
            T k1[3], k2[3], gam[3], lmd[3], al[3], be[3], d1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                k1[k] = (T)r.att_k1[k]; k2[k] = (T)r.att_k2[k]; gam[k] = (T)r.att_gamma[k]; lmd[k] = (T)r.att_lmd[k];
                al[k] = (T)r.att_alpha[k]; be[k] = (T)r.att_beta[k];
            }
            att_control<T>(c, x, t1, k1, k2, gam, lmd, al, be, s1, att_ref, datt, torque, d1);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                torque[k] = clampc<T>(torque[k], -(T)r.att_saturation[k], (T)r.att_saturation[k]);
                UAV_STS(io.state, n, R_S1 + k, i, s1[k]);
            }
        }
        uav_rk44<T, false>(c, x, t1, uf, torque, dis);
    }
    time += r.dt;
    // is_episode_Terminal, uav.py:543-560
    int flag = 0;
    if (time > r.t_term) flag = 1;
    bool pout = false, aout = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pout = pout || x[k] < (T)r.pos_zone_min[k] || x[k] > (T)r.pos_zone_max[k];
        aout = aout || x[6 + k] < (T)r.att_zone_min[k] || x[6 + k] > (T)r.att_zone_max[k];
    }
    if (pout) flag = 2;
    if (aout) flag = 3;
    const bool done = flag != 0;
    // the observation of the new state needs the freshly written att_ref / dot_att_ref (variant 1): program order
#pragma unroll
    for (int k = 0; k < 12; ++k) UAV_STS(io.state, n, k, i, x[k]);
    io.time[i] = time;
    Trig<T> t2;
    t2.eval(x[6], x[7], x[8], false);
    observe<T, V>(r, rd, io, n, i, x, t2, time, nxt, e, de);
    // rewards
    const T Qx = (T)r.Qx, Qv = (T)r.Qv, R = (T)r.R;
    T an2;
    {
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < AD; ++k) s += a[k] * a[k];
        const T nn = Mth<T>::sqrt(s);
        an2 = nn * nn;
    }
    T reward;
    if (V == 0 || V == 1) { // UavHoverOuterLoop.py:94-109, UavHover.py:116-131
        const T v[3] = {x[3], x[4], x[5]};
        const T r1 = -sq3_tanh10<T>(e) * (T)0.5 * Qx - sq3<T>(e) * (T)0.5 * Qx;
        const T r2 = -sq3_tanh10<T>(v) * (T)0.5 * Qx - sq3<T>(v) * (T)0.5 * Qv;
        const T r3 = -an2 * R;
        T r4 = (T)0;
        if (pout || aout) r4 = -(T)(r.time_max - time) / c.dt * (Qx * sq3<T>(e) + Qv * sq3<T>(v) + R * an2);
        reward = r1 + r2 + r3 + r4;
    } else { // UavInnerLoop.py:101-120, UavTrackingOuterLoop.py:103-120
        T r1 = -sq3<T>(e) * Qx, r2 = -sq3<T>(de) * Qv;
        r1 -= sq3_tanh10<T>(e) * Qx;
        r2 -= sq3_tanh10<T>(de) * Qv;
        const T r3 = -an2 * R;
        T r4 = (T)0;
        if ((V == 2) ? aout : (pout || aout)) r4 = (T)(r.time_max - time) / c.dt * (r1 + r2 + r3);
        reward = r1 + r2 + r3 + r4;
    }
#pragma unroll
    for (int k = 0; k < S; ++k) stio<T, IO32>(io.next_obs, n, k, i, nxt[k]);
    stio<T, IO32>(io.reward, n, 0, i, reward);
    io.done[i] = done ? 1 : 0;
    io.flag[i] = flag;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        T xr[12];
        reset_state<T, V>(r, io, n, i, seed, off, xr);
        Trig<T> tr;
        tr.eval(xr[6], xr[7], xr[8], false);
        observe<T, V>(r, rd, io, n, i, xr, tr, 0.0, nxt, e, de);
    }
    if (io.reset_obs) {
#pragma unroll
        for (int k = 0; k < S; ++k) stio<T, IO32>(io.reset_obs, n, k, i, nxt[k]);
    }
}

template <typename T, int V, bool IO32>
__global__ void __launch_bounds__(B200_BLOCK)
uavrobust_reset_kernel(const __grid_constant__ RP r, const __grid_constant__ RobDerived rd,
                       const __grid_constant__ b200env_io io, int64_t n,
                       const uint8_t *mask, uint64_t seed, int64_t off, int observe_only) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!observe_only && mask && !mask[i]) return;
    constexpr int S = V == 1 ? 12 : 6;
    T x[12];
    if (observe_only) {
#pragma unroll
        for (int k = 0; k < 12; ++k) x[k] = UAV_LDS(io.state, n, k, i);
    } else {
        reset_state<T, V>(r, io, n, i, seed, off, x);
    }
    if (io.next_obs) {
        T o[S], e[3], de[3];
        Trig<T> t;
        t.eval(x[6], x[7], x[8], false);
        observe<T, V>(r, rd, io, n, i, x, t, io.time[i], o, e, de);
#pragma unroll
        for (int k = 0; k < S; ++k) stio<T, IO32>(io.next_obs, n, k, i, o[k]);
    }
}

template <typename T, bool IO32>
int launch_step(int V, const RP &r, const b200env_io &io, int64_t n, uint32_t flags, uint64_t seed, int64_t off, cudaStream_t s) {
    const unsigned g = b200_grid(n);
    const UavDerived dv = uav_derive(r.m, r.J, r.kt);
    const RobDerived rd = rob_derive(r);
    switch (V) {
    case 0: uavrobust_step_kernel<T, 0, IO32><<<g, B200_BLOCK, 0, s>>>(r, dv, rd, io, n, flags, seed, off); break;
    case 1: uavrobust_step_kernel<T, 1, IO32><<<g, B200_BLOCK, 0, s>>>(r, dv, rd, io, n, flags, seed, off); break;
    case 2: uavrobust_step_kernel<T, 2, IO32><<<g, B200_BLOCK, 0, s>>>(r, dv, rd, io, n, flags, seed, off); break;
    case 3: uavrobust_step_kernel<T, 3, IO32><<<g, B200_BLOCK, 0, s>>>(r, dv, rd, io, n, flags, seed, off); break;
    default: return B200ENV_EENV;
    }
    return b200_check_launch();
}
template <typename T, bool IO32>
int launch_reset(int V, const RP &r, const b200env_io &io, int64_t n, const uint8_t *mask, uint64_t seed, int64_t off,
                 int observe_only, cudaStream_t s) {
    const unsigned g = b200_grid(n);
    const RobDerived rd = rob_derive(r);
    switch (V) {
    case 0: uavrobust_reset_kernel<T, 0, IO32><<<g, B200_BLOCK, 0, s>>>(r, rd, io, n, mask, seed, off, observe_only); break;
    case 1: uavrobust_reset_kernel<T, 1, IO32><<<g, B200_BLOCK, 0, s>>>(r, rd, io, n, mask, seed, off, observe_only); break;
    case 2: uavrobust_reset_kernel<T, 2, IO32><<<g, B200_BLOCK, 0, s>>>(r, rd, io, n, mask, seed, off, observe_only); break;
    case 3: uavrobust_reset_kernel<T, 3, IO32><<<g, B200_BLOCK, 0, s>>>(r, rd, io, n, mask, seed, off, observe_only); break;
    default: return B200ENV_EENV;
    }
    return b200_check_launch();
}

} // namespace

int uavrobust_dims(int variant, int *sf, int *od, int *ad, int *dd) {
    if (variant < 0 || variant > 3) return B200ENV_EENV;
    if (sf) *sf = B200_UAVROBUST_STATE_FIELDS;
    if (od) *od = variant == 1 ? 12 : 6;
    if (ad) *ad = variant == 1 ? 6 : 3;
    if (dd) *dd = variant == 2 ? 0 : 3;
    return B200ENV_OK;
}
int uavrobust_step(int dtype, int64_t n, const void *params, const b200env_io *io, uint32_t flags, uint64_t seed,
                   int64_t off, cudaStream_t s) {
    if (!io->state || !io->time || !io->action || !io->next_obs || !io->reward || !io->done || !io->flag) return B200ENV_ENULL;
    if ((flags & B200ENV_AUTO_RESET) && !io->episode) return B200ENV_ENULL;
    const RP &r = *static_cast<const RP *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io)) return launch_step<double, true>(r.variant, r, *io, n, flags, seed, off, s);
    return dtype == B200ENV_F64 ? launch_step<double, false>(r.variant, r, *io, n, flags, seed, off, s) : launch_step<float, false>(r.variant, r, *io, n, flags, seed, off, s);
}
int uavrobust_reset(int dtype, int64_t n, const void *params, const b200env_io *io, const uint8_t *mask, uint64_t seed,
                    int64_t off, cudaStream_t s) {
    if (!io->state || !io->time || !io->episode) return B200ENV_ENULL;
    const RP &r = *static_cast<const RP *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io)) return launch_reset<double, true>(r.variant, r, *io, n, mask, seed, off, 0, s);
    return dtype == B200ENV_F64 ? launch_reset<double, false>(r.variant, r, *io, n, mask, seed, off, 0, s) : launch_reset<float, false>(r.variant, r, *io, n, mask, seed, off, 0, s);
}
int uavrobust_observe(int dtype, int64_t n, const void *params, const b200env_io *io, cudaStream_t s) {
    if (!io->state || !io->time || !io->next_obs) return B200ENV_ENULL;
    const RP &r = *static_cast<const RP *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io)) return launch_reset<double, true>(r.variant, r, *io, n, nullptr, 0, 0, 1, s);
    return dtype == B200ENV_F64 ? launch_reset<double, false>(r.variant, r, *io, n, nullptr, 0, 0, 1, s) : launch_reset<float, false>(r.variant, r, *io, n, nullptr, 0, 0, 1, s);
}
