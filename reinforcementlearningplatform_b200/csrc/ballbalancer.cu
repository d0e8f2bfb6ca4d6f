// ballbalancer.cu -- K-BB: batched BallBalancer1D step.
// Replaces environment/BallBalancer/BallBalancer1D.py:200-322 for n instances.
#include "env_kernel.cuh"

namespace {
template <typename T>
struct BallBalancer {
    typedef b200_ballbalancer_params P;
    static constexpr int SF = B200_BALLBALANCER_STATE_FIELDS, OD = 3, AD = 1;
    T pos, vel, theta, error;
    double time;

    __device__ __forceinline__ void load(const b200env_io &io, int64_t n, int64_t i) {
        pos = ld<T>(io.state, n, 0, i); vel = ld<T>(io.state, n, 1, i);
        theta = ld<T>(io.state, n, 2, i); error = ld<T>(io.state, n, 3, i);
        time = io.time[i];
    }
    __device__ __forceinline__ void store(const b200env_io &io, int64_t n, int64_t i) const {
        st<T>(io.state, n, 0, i, pos); st<T>(io.state, n, 1, i, vel);
        st<T>(io.state, n, 2, i, theta); st<T>(io.state, n, 3, i, error);
        io.time[i] = time;
    }
    // get_state :200-211
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        const T g = (T)p.static_gain;
        // quotients by step-constant normalisers: reciprocal + Markstein correction (common.cuh Divisor), no IEEE-division slow path
        o[0] = Divisor<T>((T)(p.L), Mth<T>::rcp((T)(p.L))).div(pos) * g;
        o[1] = Divisor<T>((T)(p.v_max - p.v_min), Mth<T>::rcp((T)(p.v_max - p.v_min))).div((T)2 * vel - (T)p.v_max - (T)p.v_min) * g;
        o[2] = Divisor<T>((T)(p.theta_max - p.theta_min), Mth<T>::rcp((T)(p.theta_max - p.theta_min))).div((T)2 * theta - (T)p.theta_max - (T)p.theta_min) * g;
    }
    __device__ __forceinline__ bool success(const P &p) const { // :213-216
        return Mth<T>::abs(error) <= (T)0.001 && Mth<T>::abs(vel) <= (T)0.005 && Mth<T>::abs(theta) <= (T)p.deg1;
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T omega = Mth<T>::min(Mth<T>::max(act[0], (T)p.omega_min), (T)p.omega_max); // np.clip :255
        const T K = (T)p.K;
        const double h = p.dt / 10.0, tt = time + p.dt;
        const T hT = (T)h, half = (T)0.5;
        const T kth = hT * omega; // K_i[2] = h * omega for every stage
        while (time < tt) { // :257-270, sub-step count 10|11 (note N1)
            const T k1p = hT * vel, k1v = hT * (K * Mth<T>::sin(theta));
            const T k2p = hT * (vel + k1v * half), k2v = hT * (K * Mth<T>::sin(theta + kth * half));
            // K3[1] = h K sin(theta + K2[2] / 2) with K2[2] = K1[2] = h omega: the same sine as stage 2, bit for bit
            const T k3p = hT * (vel + k2v * half), k3v = k2v;
            const T k4p = hT * (vel + k3v), k4v = hT * (K * Mth<T>::sin(theta + kth));
            pos = pos + div6<T>(k1p + (T)2 * k2p + (T)2 * k3p + k4p);
            const T nv = vel + div6<T>(k1v + (T)2 * k2v + (T)2 * k3v + k4v);
            const T nt = theta + div6<T>(kth + (T)2 * kth + (T)2 * kth + kth);
            vel = Mth<T>::min(Mth<T>::max(nv, (T)p.v_min), (T)p.v_max);
            theta = Mth<T>::min(Mth<T>::max(nt, (T)p.theta_min), (T)p.theta_max);
            time += h;
        }
        // is_Terminal :218-236 returns at the first true test; is_success() still sees the PREVIOUS step's error
        if (pos < -(T)p.L || pos > (T)p.L) { flag = 1; done = true; }
        else if (time > p.time_max) { flag = 2; done = true; }
        else if (success(p)) { flag = 3; done = true; }
        else { flag = 0; done = false; }
        error = (T)p.target - pos; // :281
        observe(p, nxt);
        // get_reward :238-246
        const T e = Divisor<T>((T)(p.L), Mth<T>::rcp((T)(p.L))).div(error) * (T)p.static_gain;
        const T r1 = -(e * e) - Mth<T>::tanh((T)100 * e) + (T)0.5;
        const T r3 = success(p) ? (T)1000 : (T)0;
        reward = r1 + (T)0 + r3;
    }
    // reset(random=True) :288-322
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        theta = (T)rng.uniform(p.reset_theta_lo, p.reset_theta_hi);
        pos = (T)rng.uniform(p.reset_pos_lo, p.reset_pos_hi);
        vel = (T)p.init_vel;
        error = (T)p.target - pos;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(ballbalancer, BallBalancer, B200_BALLBALANCER_STATE_FIELDS, 3, 1, 0)
