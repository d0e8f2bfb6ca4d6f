// fas.cu -- K-FAS: batched Flight_Attitude_Simulator step.
// Replaces environment/FlightAttitudeSimulator/FlightAttitudeSimulator.py:173-287 (get_state, is_Terminal, get_reward,
// ode, rk44, step_update, reset) for n instances; the PPO2/DPPO2 demo copy differs only in parameters.
#include "env_kernel.cuh"

namespace {
template <typename T>
struct Fas {
    typedef b200_fas_params P;
    static constexpr int SF = B200_FAS_STATE_FIELDS, OD = 2, AD = 1;
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
    // :173-185
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        const Divisor<T> dth((T)(p.max_theta - p.min_theta)), dom((T)(p.max_omega - p.min_omega));
        o[0] = dth.div((T)2 * theta - (T)p.max_theta - (T)p.min_theta) * (T)p.static_gain;
        o[1] = dom.div((T)2 * dtheta - (T)p.max_omega - (T)p.min_omega) * (T)p.static_gain;
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T force = act[0];
        const T FL = force * (T)p.L - (T)p.mgd; // self.force * self.L - self.m * self.g * self.dis
        const T kk = (T)p.k;
        const Divisor<T> den((T)p.denom); // J + m d^2: 40+ quotients per control period
        // rk44 :238-252: `while self.time < tt` with h = dt / 10 (10 or 11 trips, note N1)
        const double h = p.dt / 10.0, tt = time + p.dt;
        const T hT = (T)h, half = (T)0.5;
        while (time < tt) {
            const T a1 = den.div(FL - kk * dtheta);
            const T k1a = hT * dtheta, k1b = hT * a1;
            const T w2 = dtheta + k1b * half;
            const T a2 = den.div(FL - kk * w2);
            const T k2a = hT * w2, k2b = hT * a2;
            const T w3 = dtheta + k2b * half;
            const T a3 = den.div(FL - kk * w3);
            const T k3a = hT * w3, k3b = hT * a3;
            const T w4 = dtheta + k3b;
            const T a4 = den.div(FL - kk * w4);
            const T k4a = hT * w4, k4b = hT * a4;
            theta = theta + div6<T>(k1a + (T)2 * k2a + (T)2 * k3a + k4a);
            dtheta = dtheta + div6<T>(k1b + (T)2 * k2b + (T)2 * k3b + k4b);
            time += h;
        }
        // is_Terminal :193-211 (all tests run, the last true one wins)
        flag = 0;
        if (theta > (T)p.theta_term_hi) flag = 1;
        if (theta < (T)p.theta_term_lo) flag = 2;
        if (time > p.time_max) flag = 3;
        done = flag != 0;
        observe(p, nxt);
        // get_reward :217-230
        const T r1 = -(theta * theta) * (T)p.Q;
        const T r2 = -(dtheta * dtheta) * (T)p.R;
        T r3 = (T)0;
        if (flag == 1 || flag == 2) r3 = (T)((p.time_max - time) / p.dt) * (r1 + r2);
        reward = r1 + r2 + r3;
    }
    // reset(random=True) :264-287
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        theta = (T)rng.uniform(p.reset_lo, p.reset_hi);
        dtheta = (T)0;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(fas, Fas, B200_FAS_STATE_FIELDS, 2, 1, 0)
