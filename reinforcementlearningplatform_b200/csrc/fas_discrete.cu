// fas_discrete.cu -- K-FASD: batched FlightAttitudeSimulatorDiscrete step.
// Replaces environment/FlightAttitudeSimulator/FlightAttitudeSimulatorDiscrete.py:158-274 (get_state, is_out,
// is_Terminal, step_update with its closure f(), get_reward, reset) for n instances.  The action is the force value
// taken from the env's discrete action_space; quirks kept: the `dis + m dis^2` denominator of a1, two RK4 steps of
// h = dt per control period (`while t_sim <= self.dt`), the -0.8 bounce at +-theta_max.
#include "env_kernel.cuh"

namespace {
template <typename T>
struct FasDiscrete {
    typedef b200_fas_discrete_params P;
    static constexpr int SF = B200_FAS_DISCRETE_STATE_FIELDS, OD = 2, AD = 1;
    T theta, dtheta;
    double time;

    __device__ __forceinline__ void load(const b200env_io &io, int64_t n, int64_t i) {
        theta = ld<T>(io.state, n, 0, i);
        dtheta = ld<T>(io.state, n, 1, i);
        time = io.time[i];
    }
    __device__ __forceinline__ void store(const b200env_io &io, int64_t n, int64_t i) const {
        st<T>(io.state, n, 0, i, theta);
        st<T>(io.state, n, 1, i, dtheta);
        io.time[i] = time;
    }
    // :158-162
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        o[0] = Divisor<T>((T)(p.theta_max), Mth<T>::rcp((T)(p.theta_max))).div(-theta) * (T)p.static_gain;
        o[1] = Divisor<T>((T)(p.dtheta_max), Mth<T>::rcp((T)(p.dtheta_max))).div(dtheta) * (T)p.static_gain;
    }
    static __device__ __forceinline__ T f(const P &p, T a0, T angle, T dangle) { // :199-203
        return (T)p.a2 * dangle + (T)p.a1 * Mth<T>::cos(angle) + a0;
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T a0 = Divisor<T>((T)(p.denom), Mth<T>::rcp((T)(p.denom))).div((T)p.L * act[0]);
        const T h = (T)p.dt, half = (T)0.5;
        // `t_sim = 0; while t_sim <= dt: ...; t_sim += h` with h = dt: trips at t_sim = 0 and t_sim = dt (:205-219)
#pragma unroll 1
        for (int trip = 0; trip < 2; ++trip) {
            const T K1 = dtheta, L1 = f(p, a0, theta, dtheta);
            const T K2 = dtheta + h * L1 * half, L2 = f(p, a0, theta + h * K1 * half, K2);
            const T K3 = dtheta + h * L2 * half, L3 = f(p, a0, theta + h * K2 * half, K3);
            const T K4 = dtheta + h * L3, L4 = f(p, a0, theta + h * K3, K4);
            theta = theta + div6<T>(h * (K1 + (T)2 * K2 + (T)2 * K3 + K4));
            dtheta = dtheta + div6<T>(h * (L1 + (T)2 * L2 + (T)2 * L3 + L4));
        }
        if (theta > (T)p.theta_max) { theta = (T)p.theta_max; dtheta = (T)p.bounce * dtheta; }
        if (theta < -(T)p.theta_max) { theta = -(T)p.theta_max; dtheta = (T)p.bounce * dtheta; }
        time = time + p.dt;
        // is_Terminal :178-196: is_out() first, then the time-out; the success test is commented out in the reference
        flag = 0;
        if (theta > (T)p.theta_out || theta < -(T)p.theta_out) flag = 1;
        else if (time > p.time_max) flag = 2;
        done = flag != 0;
        observe(p, nxt);
        // get_reward :232-244
        const T r1 = -(theta * theta) * (T)p.Q;
        const T r2 = -(dtheta * dtheta) * (T)p.R;
        T r3 = (T)0;
        if (flag != 0) r3 = (T)((p.time_max - time) / p.dt) * (r1 + r2);
        reward = r1 + r2 + r3;
    }
    // reset(random=True) :251-273
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        theta = (T)rng.uniform(-p.theta_max, p.theta_max);
        dtheta = (T)0;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(fas_discrete, FasDiscrete, B200_FAS_DISCRETE_STATE_FIELDS, 2, 1, 0)
