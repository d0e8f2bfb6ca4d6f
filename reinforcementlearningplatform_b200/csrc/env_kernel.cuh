// env_kernel.cuh -- generic one-thread-per-instance step / reset / observe kernels for the small env families.
//
// An env family provides a struct E<T> with
//     using P = <params struct>;  static constexpr int SF, OD, AD;       (state fields, obs dim, action dim)
//     load(io, n, i) / store(io, n, i)           SoA state <-> registers (time included)
//     observe(p, o[OD])                          get_state()
//     step(p, act[AD], cur[OD], flag, done, reward)   rk44 + is_Terminal + get_reward (next obs is taken afterwards);
//                                                cur is only valid when io.obs != NULL and must not be read by step()
//     reset(p, rng)                              reset(random=True) with Philox draws in the reference's draw order
// and the kernels below add the common data movement: all loads first, outputs, optional auto-reset.
#pragma once
#include "common.cuh"
#include <type_traits>

// ENV_MINBLOCKS (set by a family file before this header): resident blocks per SM the step and rollout kernels are compiled
// for, i.e. a register cap.  Left undefined, the bounds stay one-argument: `(B200_BLOCK, 1)` is NOT the same thing -- it lifts
// ptxas' default register target (TwoLink: 94 -> 106 registers, 5 -> 4 blocks per SM, 0.0955 -> 0.1036 ms per 1 M instances).
#ifdef ENV_MINBLOCKS
#define ENV_BOUNDS __launch_bounds__(B200_BLOCK, ENV_MINBLOCKS)
#else
#define ENV_BOUNDS __launch_bounds__(B200_BLOCK)
#endif
template <typename T, class E, bool IO32>
__global__ void ENV_BOUNDS
env_step_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ b200env_io io, int64_t n,
                uint32_t flags, uint64_t seed, int64_t off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    E e;
    e.load(io, n, i);
    T act[E::AD];
#pragma unroll
    for (int k = 0; k < E::AD; ++k) act[k] = ldio<T, IO32>(io.action, n, k, i);
    T cur[E::OD] = {}, nxt[E::OD];
    if (io.obs) { // self.current_state = self.get_state(); no family's step() reads it, so it is only evaluated when the
                  // caller wants it stored (pure-observation envs reuse the previous policy observation, vec_env.py)
        e.observe(p, cur);
#pragma unroll
        for (int k = 0; k < E::OD; ++k) stio<T, IO32>(io.obs, n, k, i, cur[k]);
    }
    int flag = 0;
    bool done = false;
    T reward = (T)0;
    e.step(p, act, cur, flag, done, reward, nxt);
#pragma unroll
    for (int k = 0; k < E::OD; ++k) stio<T, IO32>(io.next_obs, n, k, i, nxt[k]);
    stio<T, IO32>(io.reward, n, 0, i, reward);
    io.done[i] = done ? 1 : 0;
    io.flag[i] = flag;
    if (done && (flags & B200ENV_AUTO_RESET)) {
        const uint32_t ep = io.episode[i];
        Philox rng(seed, (uint64_t)(off + i), ep);
        e.reset(p, rng);
        io.episode[i] = ep + 1u;
        e.observe(p, nxt);
    }
    if (io.reset_obs) {
#pragma unroll
        for (int k = 0; k < E::OD; ++k) stio<T, IO32>(io.reset_obs, n, k, i, nxt[k]);
    }
    e.store(io, n, i);
}

template <typename T, class E, bool IO32>
__global__ void __launch_bounds__(B200_BLOCK)
env_reset_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ b200env_io io, int64_t n,
                 const uint8_t *mask, uint64_t seed, int64_t off, int observe_only) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!observe_only && mask && !mask[i]) return;
    E e;
    e.load(io, n, i);
    if (!observe_only) {
        const uint32_t ep = io.episode[i];
        Philox rng(seed, (uint64_t)(off + i), ep);
        e.reset(p, rng);
        io.episode[i] = ep + 1u;
        e.store(io, n, i);
    }
    if (io.next_obs) {
        T o[E::OD];
        e.observe(p, o);
#pragma unroll
        for (int k = 0; k < E::OD; ++k) stio<T, IO32>(io.next_obs, n, k, i, o[k]);
    }
}

// T_steps control periods in ONE launch (b200env_rollout): the instance state stays in registers between steps, so per
// step only the action row is read and the transition row (s, s', r, done, flag) is written -- no state round trip
// through HBM and no per-step launch.  Row t of an array lives `stride` elements after row t-1 (time-major rollout
// buffer, rollout.py).  Same arithmetic as env_step_kernel step by step, including the in-kernel auto-reset.
// (compiled for the same register cap as env_step_kernel: ptxas' scheduling, and with it its choice of mul + add pairs to
// contract, follows the cap, and tests/test_rollout_gpu.py demands the same bits from both kernels)
template <typename T, class E, bool IO32>
__global__ void ENV_BOUNDS
env_rollout_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ b200env_io io,
                   const __grid_constant__ b200env_rollout_spec rs, int64_t n, uint32_t flags, uint64_t seed, int64_t off) {
    typedef typename std::conditional<IO32, float, T>::type TIO;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    E e;
    e.load(io, n, i);
    T nxt[E::OD];
    constexpr int U = 4; // time steps per trip: their action rows are loaded together, ahead of the dependent steps
    for (int64_t t0 = 0; t0 < rs.steps; t0 += U) {
        T acts[U][E::AD];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t t = (t0 + u < rs.steps) ? t0 + u : rs.steps - 1;
            const TIO *act_row = static_cast<const TIO *>(io.action) + t * rs.action_stride;
#pragma unroll
            for (int k = 0; k < E::AD; ++k) acts[u][k] = ldio_idx<T, IO32>(act_row, n, k, i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t t = t0 + u;
            if (t < rs.steps) {
                T cur[E::OD] = {};
                if (io.obs) {
                    e.observe(p, cur);
                    TIO *row = static_cast<TIO *>(io.obs) + t * rs.obs_stride;
#pragma unroll
                    for (int k = 0; k < E::OD; ++k) stio_idx<T, IO32>(row, n, k, i, cur[k]);
                }
                int flag = 0;
                bool done = false;
                T reward = (T)0;
                e.step(p, acts[u], cur, flag, done, reward, nxt);
                {
                    TIO *row = static_cast<TIO *>(io.next_obs) + t * rs.next_obs_stride;
#pragma unroll
                    for (int k = 0; k < E::OD; ++k) stio_idx<T, IO32>(row, n, k, i, nxt[k]);
                }
                stio_idx<T, IO32>(static_cast<TIO *>(io.reward) + t * rs.reward_stride, n, 0, i, reward);
                io.done[t * rs.done_stride + i] = done ? 1 : 0;
                io.flag[t * rs.flag_stride + i] = flag;
                if (done && (flags & B200ENV_AUTO_RESET)) {
                    const uint32_t ep = io.episode[i];
                    Philox rng(seed, (uint64_t)(off + i), ep);
                    e.reset(p, rng);
                    io.episode[i] = ep + 1u;
                    e.observe(p, nxt);
                }
            }
        }
    }
    if (io.reset_obs) {
#pragma unroll
        for (int k = 0; k < E::OD; ++k) stio_idx<T, IO32>(io.reset_obs, n, k, i, nxt[k]);
    }
    e.store(io, n, i);
}

// host-side launchers: EnvT is the family template, instantiated for double and float
template <template <typename> class EnvT>
int env_launch_step(int dtype, int64_t n, const void *params, const b200env_io *io, uint32_t flags, uint64_t seed,
                    int64_t off, cudaStream_t s) {
    typedef typename EnvT<double>::P P;
    if (!io->state || !io->time || !io->action || !io->next_obs || !io->reward || !io->done || !io->flag)
        return B200ENV_ENULL;
    if ((flags & B200ENV_AUTO_RESET) && !io->episode) return B200ENV_ENULL;
    const P &p = *static_cast<const P *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io))
        env_step_kernel<double, EnvT<double>, true><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, flags, seed, off);
    else if (dtype == B200ENV_F64)
        env_step_kernel<double, EnvT<double>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, flags, seed, off);
    else
        env_step_kernel<float, EnvT<float>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, flags, seed, off);
    return b200_check_launch();
}

template <template <typename> class EnvT>
int env_launch_rollout(int dtype, int64_t n, const void *params, const b200env_io *io, const b200env_rollout_spec *rs,
                       uint32_t flags, uint64_t seed, int64_t off, cudaStream_t s) {
    typedef typename EnvT<double>::P P;
    if (!io->state || !io->time || !io->action || !io->next_obs || !io->reward || !io->done || !io->flag)
        return B200ENV_ENULL;
    if ((flags & B200ENV_AUTO_RESET) && !io->episode) return B200ENV_ENULL;
    const P &p = *static_cast<const P *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io))
        env_rollout_kernel<double, EnvT<double>, true><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, *rs, n, flags, seed, off);
    else if (dtype == B200ENV_F64)
        env_rollout_kernel<double, EnvT<double>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, *rs, n, flags, seed, off);
    else
        env_rollout_kernel<float, EnvT<float>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, *rs, n, flags, seed, off);
    return b200_check_launch();
}

template <template <typename> class EnvT>
int env_launch_reset(int dtype, int64_t n, const void *params, const b200env_io *io, const uint8_t *mask,
                     uint64_t seed, int64_t off, cudaStream_t s, int observe_only) {
    typedef typename EnvT<double>::P P;
    if (!io->state || !io->time) return B200ENV_ENULL;
    if (!observe_only && !io->episode) return B200ENV_ENULL;
    if (observe_only && !io->next_obs) return B200ENV_ENULL;
    const P &p = *static_cast<const P *>(params);
    if (dtype == B200ENV_F64 && b200_io32(io))
        env_reset_kernel<double, EnvT<double>, true><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, mask, seed, off, observe_only);
    else if (dtype == B200ENV_F64)
        env_reset_kernel<double, EnvT<double>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, mask, seed, off, observe_only);
    else
        env_reset_kernel<float, EnvT<float>, false><<<b200_grid(n), B200_BLOCK, 0, s>>>(p, *io, n, mask, seed, off, observe_only);
    return b200_check_launch();
}

#define B200_FAMILY_IMPL(name, EnvT, sf, od, ad, dd)                                                                  \
    int name##_dims(int variant, int *psf, int *pod, int *pad, int *pdd) {                                            \
        if (variant != 0) return B200ENV_EENV;                                                                        \
        if (psf) *psf = sf;                                                                                           \
        if (pod) *pod = od;                                                                                           \
        if (pad) *pad = ad;                                                                                           \
        if (pdd) *pdd = dd;                                                                                           \
        return B200ENV_OK;                                                                                            \
    }                                                                                                                 \
    int name##_step(int dtype, int64_t n, const void *params, const b200env_io *io, uint32_t flags, uint64_t seed,    \
                    int64_t off, cudaStream_t s) {                                                                    \
        return env_launch_step<EnvT>(dtype, n, params, io, flags, seed, off, s);                                      \
    }                                                                                                                 \
    int name##_rollout(int dtype, int64_t n, const void *params, const b200env_io *io,                                \
                       const b200env_rollout_spec *rs, uint32_t flags, uint64_t seed, int64_t off, cudaStream_t s) {  \
        return env_launch_rollout<EnvT>(dtype, n, params, io, rs, flags, seed, off, s);                               \
    }                                                                                                                 \
    int name##_reset(int dtype, int64_t n, const void *params, const b200env_io *io, const uint8_t *mask,             \
                     uint64_t seed, int64_t off, cudaStream_t s) {                                                    \
        return env_launch_reset<EnvT>(dtype, n, params, io, mask, seed, off, s, 0);                                   \
    }                                                                                                                 \
    int name##_observe(int dtype, int64_t n, const void *params, const b200env_io *io, cudaStream_t s) {              \
        return env_launch_reset<EnvT>(dtype, n, params, io, nullptr, 0, 0, s, 1);                                     \
    }
