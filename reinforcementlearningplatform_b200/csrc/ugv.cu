// ugv.cu -- K-UGV: batched UGVForward / UGVBidirectional step (unicycle with drag).
// Replaces environment/UGV/UGVForward.py:217-362 and UGVBidirectional.py:217-367 for n instances.
// eight resident blocks per SM (64 registers): 0.0493 -> 0.0470 ms per 1 M instances, A/B on one box
#ifndef ENV_MINBLOCKS
#define ENV_MINBLOCKS 8
#endif
#include "env_kernel.cuh"

namespace {
// utils/functions.py:49-60 with v1 = (cos phi, sin phi)
template <typename T>
__device__ __forceinline__ T vector_rad_oriented(T x1, T y1, T x2, T y2) {
    if (np_norm2<T>(x2, y2) < (T)1e-4 || np_norm2<T>(x1, y1) < (T)1e-4) return (T)0;
    return Mth<T>::atan2(x1 * y2 - y1 * x2, x1 * x2 + y1 * y2);
}

template <typename T>
struct Ugv {
    typedef b200_ugv_params P;
    static constexpr int SF = B200_UGV_STATE_FIELDS, OD = 4, AD = 2;
    T x, y, vel, phi, omega;
    double time;

    __device__ __forceinline__ void load(const b200env_io &io, int64_t n, int64_t i) {
        x = ld<T>(io.state, n, 0, i); y = ld<T>(io.state, n, 1, i); vel = ld<T>(io.state, n, 2, i);
        phi = ld<T>(io.state, n, 3, i); omega = ld<T>(io.state, n, 4, i);
        time = io.time[i];
    }
    __device__ __forceinline__ void store(const b200env_io &io, int64_t n, int64_t i) const {
        st<T>(io.state, n, 0, i, x); st<T>(io.state, n, 1, i, y); st<T>(io.state, n, 2, i, vel);
        st<T>(io.state, n, 3, i, phi); st<T>(io.state, n, 4, i, omega);
        io.time[i] = time;
    }
    // get_e / get_e_phi: UGVForward.py:315-320, UGVBidirectional.py:314-325
    __device__ __forceinline__ void errors(const P &p, T &e, T &ephi) const {
        T s, c;
        Mth<T>::sincos(phi, &s, &c);
        const T dx = (T)p.target_x - x, dy = (T)p.target_y - y;
        e = np_norm2<T>(dx, dy);
        ephi = vector_rad_oriented<T>(c, s, dx, dy);
        if (p.bidirectional) {
            const T d = c * dx + s * dy;
            const T sg = d > (T)0 ? (T)1 : (d < (T)0 ? (T)-1 : (T)0); // np.sign
            e = sg * e;
            if (ephi >= (T)(M_PI / 2)) ephi = ephi - (T)M_PI;
            if (ephi <= (T)(-M_PI / 2)) ephi = ephi + (T)M_PI;
        }
    }
    // get_state :217-227
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        T e, ephi;
        errors(p, e, ephi);
        const T g = (T)p.static_gain;
        if (p.bidirectional) {
            o[0] = Divisor<T>((T)(p.e_max), Mth<T>::rcp((T)(p.e_max))).div(e) * g;
            o[1] = Divisor<T>((T)(p.v_max), Mth<T>::rcp((T)(p.v_max))).div(vel) * g;
        } else {
            o[0] = ((T)(2 / p.e_max) * e - (T)1) * g;
            o[1] = ((T)(2 / p.v_max) * vel - (T)1) * g;
        }
        o[2] = Divisor<T>((T)(p.e_phi_max), Mth<T>::rcp((T)(p.e_phi_max))).div(ephi) * g;
        o[3] = Divisor<T>((T)(p.omega_max), Mth<T>::rcp((T)(p.omega_max))).div(omega) * g;
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T al = act[0], aa = act[1], kf = (T)p.kf, kt = (T)p.kt;
        const T h = (T)p.dt, half = (T)0.5;
        // rk44 :294-313: one RK4 step of the 5-state ODE :281-292
        T s, c;
        Mth<T>::sincos(phi, &s, &c);
        const T k1x = h * (vel * c), k1y = h * (vel * s), k1v = h * (al - kf * vel), k1p = h * omega, k1o = h * (aa - kt * omega);
        const T v2 = vel + k1v * half, o2 = omega + k1o * half;
        Mth<T>::sincos(phi + k1p * half, &s, &c);
        const T k2x = h * (v2 * c), k2y = h * (v2 * s), k2v = h * (al - kf * v2), k2p = h * o2, k2o = h * (aa - kt * o2);
        const T v3 = vel + k2v * half, o3 = omega + k2o * half;
        Mth<T>::sincos(phi + k2p * half, &s, &c);
        const T k3x = h * (v3 * c), k3y = h * (v3 * s), k3v = h * (al - kf * v3), k3p = h * o3, k3o = h * (aa - kt * o3);
        const T v4 = vel + k3v, o4 = omega + k3o;
        Mth<T>::sincos(phi + k3p, &s, &c);
        const T k4x = h * (v4 * c), k4y = h * (v4 * s), k4v = h * (al - kf * v4), k4p = h * o4, k4o = h * (aa - kt * o4);
        x = x + div6<T>(k1x + (T)2 * k2x + (T)2 * k3x + k4x);
        y = y + div6<T>(k1y + (T)2 * k2y + (T)2 * k3y + k4y);
        vel = vel + div6<T>(k1v + (T)2 * k2v + (T)2 * k3v + k4v);
        phi = phi + div6<T>(k1p + (T)2 * k2p + (T)2 * k3p + k4p);
        omega = omega + div6<T>(k1o + (T)2 * k2o + (T)2 * k3o + k4o);
        if (!p.bidirectional && vel < (T)0) vel = (T)0; // UGVForward.py:303-304
        time += p.dt;
        if (phi > (T)M_PI) phi -= (T)(2 * M_PI);
        if (phi < (T)-M_PI) phi += (T)(2 * M_PI);
        T e, ephi;
        errors(p, e, ephi);
        // is_Terminal :247-261 (all tests run, the last true one wins)
        flag = 0;
        if (x > (T)p.map_x || x < (T)0 || y > (T)p.map_y || y < (T)0) flag = 1;
        if (time > p.time_max) flag = 2;
        if (Mth<T>::abs(e) <= (T)0.05 && Mth<T>::abs(vel) < (T)0.01) flag = 3; // is_success :239-245
        done = flag != 0;
        observe(p, nxt);
        // get_reward :263-279
        const T u_pos = -Mth<T>::abs(e) * (T)p.Q_pos;
        const T u_vel = -Mth<T>::abs(vel) * (T)p.Q_vel;
        const T u_phi = e > (T)0.1 ? -Mth<T>::abs(ephi) * (T)p.Q_phi : (T)0;
        const T u_omega = -Mth<T>::abs(omega) * (T)p.Q_omega;
        T u_psi = (T)0;
        if (flag == 1) u_psi = (T)((p.time_max - time) / p.dt) * (u_pos + u_vel + u_phi + u_omega);
        reward = u_pos + u_vel + u_phi + u_omega + u_psi;
    }
    // reset(random=True) :334-362
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        x = (T)rng.uniform(p.reset_d0, p.map_x - p.reset_d0);
        y = (T)rng.uniform(p.reset_d0, p.map_y - p.reset_d0);
        phi = (T)rng.uniform(-M_PI, M_PI);
        vel = (T)0; omega = (T)0;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(ugv, Ugv, B200_UGV_STATE_FIELDS, 4, 2, 0)
