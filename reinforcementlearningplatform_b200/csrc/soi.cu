// soi.cu -- K-SOI: batched SecondOrderIntegration step (2-D point mass with drag).
// Replaces environment/SecondOrderIntegration/SecondOrderIntegration.py:211-352 for n instances; the DPPO2 demo copy
// (obs * static_gain, no success terminal, Q_vel = Q_acc = 0) is a parameter setting.  HBM-bound: ~125 flops per
// ~140 B of traffic.
#include "env_kernel.cuh"

namespace {
template <typename T>
struct Soi {
    typedef b200_soi_params P;
    static constexpr int SF = B200_SOI_STATE_FIELDS, OD = 4, AD = 2;
    T x, y, vx, vy;
    double time;

    __device__ __forceinline__ void load(const b200env_io &io, int64_t n, int64_t i) {
        x = ld<T>(io.state, n, 0, i); y = ld<T>(io.state, n, 1, i);
        vx = ld<T>(io.state, n, 2, i); vy = ld<T>(io.state, n, 3, i);
        time = io.time[i];
    }
    __device__ __forceinline__ void store(const b200env_io &io, int64_t n, int64_t i) const {
        st<T>(io.state, n, 0, i, x); st<T>(io.state, n, 1, i, y);
        st<T>(io.state, n, 2, i, vx); st<T>(io.state, n, 3, i, vy);
        io.time[i] = time;
    }
    // get_state :211-219 (use_norm = True)
    __device__ __forceinline__ void observe(const P &p, T *o) const {
        const T g = (T)p.obs_gain;
        const Divisor<T> dx_((T)p.map_x), dy_((T)p.map_y), dv_((T)p.vmax);
        o[0] = dx_.div((T)p.target_x - x) * g;
        o[1] = dy_.div((T)p.target_y - y) * g;
        o[2] = dv_.div(-vx) * g;
        o[3] = dv_.div(-vy) * g;
    }
    __device__ __forceinline__ void step(const P &p, const T *act, const T *cur, int &flag, bool &done, T &reward, T *nxt) {
        const T fx = act[0], fy = act[1], k = (T)p.k;
        // rk44 :298-314: h = dt / 1 and `while self.time < tt` -> exactly one RK4 step (time + dt == tt)
        const T h = (T)p.dt, half = (T)0.5;
        {
            const T k1x = h * vx, k1y = h * vy, k1u = h * (fx - k * vx), k1v = h * (fy - k * vy);
            const T u2 = vx + k1u * half, v2 = vy + k1v * half;
            const T k2x = h * u2, k2y = h * v2, k2u = h * (fx - k * u2), k2v = h * (fy - k * v2);
            const T u3 = vx + k2u * half, v3 = vy + k2v * half;
            const T k3x = h * u3, k3y = h * v3, k3u = h * (fx - k * u3), k3v = h * (fy - k * v3);
            const T u4 = vx + k3u, v4 = vy + k3v;
            const T k4x = h * u4, k4y = h * v4, k4u = h * (fx - k * u4), k4v = h * (fy - k * v4);
            x = x + div6<T>(k1x + (T)2 * k2x + (T)2 * k3x + k4x);
            y = y + div6<T>(k1y + (T)2 * k2y + (T)2 * k3y + k4y);
            vx = vx + div6<T>(k1u + (T)2 * k2u + (T)2 * k3u + k4u);
            vy = vy + div6<T>(k1v + (T)2 * k2v + (T)2 * k3v + k4v);
            time += p.dt;
        }
        const Divisor<T> dm((T)p.mass);
        const T ax = dm.div(fx - k * vx), ay = dm.div(fy - k * vy); // self.acc :313
        const T ex = (T)p.target_x - x, ey = (T)p.target_y - y;
        const T e_pos = np_norm2<T>(ex, ey);
        const T e_vel = np_norm2<T>(vx, vy);
        // is_Terminal :235-249
        flag = 0;
        const T adm = (T)p.admissible_error;
        if (x > (T)p.map_x + adm || x < (T)0 - adm || y > (T)p.map_y + adm || y < (T)0 - adm) flag = 1;
        if (time > p.time_max) flag = 2;
        if (p.success_terminal && e_pos <= (T)0.05 && e_vel < (T)0.05) flag = 3;
        done = flag != 0;
        observe(p, nxt);
        // get_reward :251-284
        const T acc = np_norm2<T>(ax, ay);
        const T u_pos = -e_pos * (T)p.Q_pos, u_vel = -e_vel * (T)p.Q_vel, u_acc = -acc * (T)p.Q_acc;
        T u_extra = (T)0;
        if (flag == 1) u_extra = (T)((p.time_max - time) / p.dt) * (u_pos + u_vel + u_acc);
        reward = u_pos + u_vel + u_acc + u_extra;
    }
    // reset(random=True) :328-352
    __device__ __forceinline__ void reset(const P &p, Philox &rng) {
        x = (T)rng.uniform(0 + p.reset_margin, p.map_x - p.reset_margin);
        y = (T)rng.uniform(0 + p.reset_margin, p.map_y - p.reset_margin);
        vx = (T)0; vy = (T)0;
        time = 0.0;
    }
};
} // namespace

B200_FAMILY_IMPL(soi, Soi, B200_SOI_STATE_FIELDS, 4, 2, 0)
