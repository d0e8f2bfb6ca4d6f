// twolink.cu -- K-TLM: batched TwoLinkManipulator step.
// Replaces environment/RobotManipulator/TwoLinkManipulator.py:186-312 for n instances.  The 2x2 np.linalg.solve of the
// ODE (:237) is an in-register LU with partial pivoting (dgesv order).
// (a 64-register cap = eight resident blocks per SM measured 0.0955 -> 0.0892 ms per 1 M instances, but the step kernel and
// the fused rollout kernel then no longer produce the same low bits of the fp64 state -- ptxas contracts different mul + add
// pairs under the cap -- and tests/test_rollout_gpu.py demands equal bits: not adopted)
#include "env_kernel.cuh"

namespace {
template <typename T>
struct TwoLink {
    typedef b200_twolink_params P;
    static constexpr int SF = B200_TWOLINK_STATE_FIELDS, OD = 6, AD = 2;
    T th1, th2, w1, w2, ex, ey, tx, ty;
    double time;

    __device__ __forceinline__ void load(const b200env_io &io, int64_t n, int64_t i) {
        th1 = ld<T>(io.state, n, 0, i); th2 = ld<T>(io.state, n, 1, i);
        w1 = ld<T>(io.state, n, 2, i); w2 = ld<T>(io.state, n, 3, i);
        ex = ld<T>(io.state, n, 4, i); ey = ld<T>(io.state, n, 5, i);
        tx = ld<T>(io.state, n, 6, i); ty = ld<T>(io.state, n, 7, i);
        time = io.time[i];
    }
    __device__ __forceinline__ void store(const b200env_io &io, int64_t n, int64_t i) const {
        st<T>(io.state, n, 0, i, th1); st<T>(io.state, n, 1, i, th2);
        st<T>(io.state, n, 2, i, w1); st<T>(io.state, n, 3, i, w2);
        st<T>(io.state, n, 4, i, ex); st<T>(io.state, n, 5, i, ey);
        st<T>(io.state, n, 6, i, tx); st<T>(io.state, n, 7, i, ty);
        io.time[i] = time;
    }
    // get_state :186-192: (2 s - (min + max)) / (max - min) * static_gain with range +-static_gain is the identity
    // (2 s / 4 * 2, all exact), so obs = (error, theta, omega)
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        o[0] = ex; o[1] = ey; o[2] = th1; o[3] = th2; o[4] = w1; o[5] = w2;
    }
    // ode :226-238
    __device__ __forceinline__ void ode(const P &p, T t1, T t2, T o1, T o2, T tq0, T tq1, T &d1, T &d2) const {
        const T J = (T)p.J, mgl = (T)(p.m * p.g * p.l);
        T s2, c2;
        Mth<T>::sincos(t2, &s2, &c2);
        const T s1 = Mth<T>::sin_ld(t1), s12 = Mth<T>::sin_ld(t1 + t2);
        const T a00 = J * ((T)5 + (T)3 * c2), a01 = J * ((T)1 + (T)1.5 * c2), a11 = J;
        const T b0 = tq0 + (T)1.5 * J * s2 * (o2 * o2) + (T)3 * J * s2 * o1 * o2 - mgl * ((T)1.5 * s1 + (T)0.5 * s12);
        const T b1 = tq1 - (T)1.5 * J * s2 * (o1 * o1) - (T)0.5 * (T)p.m * (T)p.g * (T)p.l * s12;
        // dgesv on [[a00, a01], [a01, a11]]: pivot on the larger |.| of column 0
        if (Mth<T>::abs(a01) > Mth<T>::abs(a00)) {
            const T l = a00 * Mth<T>::rcp(a01); // rcp-based (<= 1 ulp): the LU solve runs 4 times per step
            const T u11 = a01 - l * a11;
            const T y1 = b0 - l * b1;
            d2 = Mth<T>::div(y1, u11);
            d1 = Mth<T>::div(b1 - a11 * d2, a01);
        } else {
            const T l = a01 * Mth<T>::rcp(a00);
            const T u11 = a11 - l * a01;
            const T y1 = b1 - l * b0;
            d2 = Mth<T>::div(y1, u11);
            d1 = Mth<T>::div(b0 - a01 * d2, a00);
        }
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T tq0 = act[0], tq1 = act[1];
        const T h = (T)p.dt, half = (T)0.5;
        // rk44 :240-250: one RK4 step of size dt
        T a1, b1, a2, b2, a3, b3, a4, b4;
        ode(p, th1, th2, w1, w2, tq0, tq1, a1, b1);
        const T k1t1 = h * w1, k1t2 = h * w2, k1w1 = h * a1, k1w2 = h * b1;
        ode(p, th1 + k1t1 * half, th2 + k1t2 * half, w1 + k1w1 * half, w2 + k1w2 * half, tq0, tq1, a2, b2);
        const T k2t1 = h * (w1 + k1w1 * half), k2t2 = h * (w2 + k1w2 * half), k2w1 = h * a2, k2w2 = h * b2;
        ode(p, th1 + k2t1 * half, th2 + k2t2 * half, w1 + k2w1 * half, w2 + k2w2 * half, tq0, tq1, a3, b3);
        const T k3t1 = h * (w1 + k2w1 * half), k3t2 = h * (w2 + k2w2 * half), k3w1 = h * a3, k3w2 = h * b3;
        ode(p, th1 + k3t1, th2 + k3t2, w1 + k3w1, w2 + k3w2, tq0, tq1, a4, b4);
        const T k4t1 = h * (w1 + k3w1), k4t2 = h * (w2 + k3w2), k4w1 = h * a4, k4w2 = h * b4;
        th1 = th1 + div6<T>(k1t1 + (T)2 * k2t1 + (T)2 * k3t1 + k4t1);
        th2 = th2 + div6<T>(k1t2 + (T)2 * k2t2 + (T)2 * k3t2 + k4t2);
        w1 = w1 + div6<T>(k1w1 + (T)2 * k2w1 + (T)2 * k3w1 + k4w1);
        w2 = w2 + div6<T>(k1w2 + (T)2 * k2w2 + (T)2 * k3w2 + k4w2);
        time += p.dt;
        // forward kinematics :252-257, error BEFORE the angle wrap
        const T l = (T)p.l;
        T sA, cA, sB, cB;
        Mth<T>::sincos(th1, &sA, &cA);
        Mth<T>::sincos(th1 + th2, &sB, &cB);
        const T midx = l * sA + (T)p.base_x, midy = -l * cA + (T)p.base_y;
        const T endx = midx + l * sB, endy = midy + -l * cB;
        ex = tx - endx; ey = ty - endy;
        const T tm = (T)p.theta_max; // wrap :259-272
        if (th1 > tm) th1 -= (T)2 * tm; else if (th1 < -tm) th1 += (T)2 * tm;
        if (th2 > tm) th2 -= (T)2 * tm; else if (th2 < -tm) th2 += (T)2 * tm;
        // is_Terminal :199-209 (returns at the first true test)
        const T en = np_norm2<T>(ex, ey), wn = np_norm2<T>(w1, w2);
        if (time > p.time_max) { flag = 2; done = true; }
        else if (en <= (T)p.miss && wn <= (T)p.omega_ok) { flag = 3; done = true; }
        else { flag = 0; done = false; }
        observe(p, nxt);
        // get_reward :211-224
        const T tn = np_norm2<T>(tq0, tq1);
        reward = -en * (T)p.Q_pos + -wn * (T)p.Q_omega + -tn * (T)p.Q_acc + (T)0;
    }
    // reset(random=True) :282-312
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        const double phi = rng.u01() * 2 * M_PI;
        const double r = rng.uniform(p.r2_lo, p.r2_hi);
        tx = (T)(cos(phi) * sqrt(r) + p.base_x);
        ty = (T)(sin(phi) * sqrt(r) + p.base_y);
        th1 = (T)rng.uniform(-p.theta_max, p.theta_max);
        th2 = (T)rng.uniform(-p.theta_max, p.theta_max);
        w1 = (T)0; w2 = (T)0;
        ex = tx - (T)p.init_end_x; ey = ty - (T)p.init_end_y;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(twolink, TwoLink, B200_TWOLINK_STATE_FIELDS, 6, 2, 0)
