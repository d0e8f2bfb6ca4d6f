/* c_abi_cartpole.c -- libb200env.so from plain C: no Python, no torch, only the CUDA runtime for device memory.
 *
 *   gcc -O2 -std=c99 examples/c_abi_cartpole.c -Iinclude -I/usr/local/cuda/include \
 *       -Lreinforcementlearningplatform_b200 -lb200env -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/c_abi_cartpole
 *   LD_LIBRARY_PATH=reinforcementlearningplatform_b200 /tmp/c_abi_cartpole [n_envs] [steps] [out.bin]
 *
 * Steps n CartPole instances (environment/CartPole/CartPole.py) with the deterministic force
 * a(t, i) = 8 sin(0.37 i + 0.11 t) and the in-kernel auto-reset, then prints the number of finished episodes and the
 * reward sum, and optionally dumps the final SoA state, time and episode counters (tests/test_c_abi_gpu.py replays the
 * same inputs through the C oracle and compares). */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200env.h"

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
    } while (0)

static double deg2rad(double d) { return d * 3.14159265358979323846 / 180.; } /* utils/functions.py:4-5: deg * pi / 180 */

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    const char *dump = argc > 3 ? argv[3] : NULL;

    b200_cartpole_params p;                       /* CartPole.py:26-42, filled like envs/cartpole.py does */
    memset(&p, 0, sizeof p);
    p.M = 1.0; p.m = 0.1; p.g = 9.8; p.ell = 0.2; p.kf = 0.2;
    p.dt = 0.02; p.time_max = 5;
    p.theta_max = deg2rad(45); p.dtheta_max = deg2rad(90); p.x_max = 1.5; p.dx_max = 3;
    p.static_gain = 2.0; p.norm_boundless = 4;
    p.theta_term_hi = p.theta_max + deg2rad(1);
    p.theta_term_lo = -p.dtheta_max - deg2rad(1); /* sic, CartPole.py:167 */
    p.reset_theta_lo = -p.theta_max * 0.5; p.reset_theta_hi = p.theta_max * 0.5;
    p.reset_x_lo = -p.x_max * 0.5; p.reset_x_hi = p.x_max * 0.5;
    p.variant = 0;
    if (b200env_params_bytes(B200ENV_CARTPOLE) != sizeof p) { fprintf(stderr, "ABI mismatch\n"); return 3; }

    int sf, od, ad, dd;
    if (b200env_dims(B200ENV_CARTPOLE, 0, &sf, &od, &ad, &dd)) return 3;

    b200env_io io;
    memset(&io, 0, sizeof io);
    double *action;
    CK(cudaMalloc(&io.state, sizeof(double) * sf * n));
    CK(cudaMalloc((void **)&io.time, sizeof(double) * n));
    CK(cudaMalloc((void **)&io.episode, sizeof(uint32_t) * n));
    CK(cudaMalloc((void **)&action, sizeof(double) * ad * n));
    CK(cudaMalloc(&io.obs, sizeof(double) * od * n));
    CK(cudaMalloc(&io.next_obs, sizeof(double) * od * n));
    CK(cudaMalloc(&io.reset_obs, sizeof(double) * od * n));
    CK(cudaMalloc(&io.reward, sizeof(double) * n));
    CK(cudaMalloc((void **)&io.done, n));
    CK(cudaMalloc((void **)&io.flag, sizeof(int32_t) * n));
    CK(cudaMemset(io.state, 0, sizeof(double) * sf * n));
    CK(cudaMemset(io.time, 0, sizeof(double) * n));
    CK(cudaMemset(io.episode, 0, sizeof(uint32_t) * n));
    io.action = action;
    io.io_dtype = B200ENV_F64;

    const uint64_t seed = 2024;
    int rc = b200env_reset(B200ENV_CARTPOLE, B200ENV_F64, n, &p, sizeof p, &io, NULL, seed, 0, NULL);
    if (rc) { fprintf(stderr, "b200env_reset: %d (cuda %d)\n", rc, b200env_last_cuda_error()); return 4; }

    double *h_a = (double *)malloc(sizeof(double) * n), *h_r = (double *)malloc(sizeof(double) * n);
    unsigned char *h_d = (unsigned char *)malloc(n);
    double reward_sum = 0.0;
    long episodes = 0;
    for (int t = 0; t < steps; ++t) {
        for (long i = 0; i < n; ++i) h_a[i] = 8.0 * sin(0.37 * (double)i + 0.11 * (double)t);
        CK(cudaMemcpy(action, h_a, sizeof(double) * n, cudaMemcpyHostToDevice));
        rc = b200env_step(B200ENV_CARTPOLE, B200ENV_F64, n, &p, sizeof p, &io, B200ENV_AUTO_RESET, seed, 0, NULL);
        if (rc) { fprintf(stderr, "b200env_step: %d (cuda %d)\n", rc, b200env_last_cuda_error()); return 4; }
        CK(cudaMemcpy(h_r, io.reward, sizeof(double) * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_d, io.done, n, cudaMemcpyDeviceToHost));
        for (long i = 0; i < n; ++i) { reward_sum += h_r[i]; episodes += h_d[i]; }
    }
    printf("%s: n=%ld steps=%d episodes=%ld reward_sum=%.17g\n", b200env_version(), n, steps, episodes, reward_sum);
    if (dump) {
        double *h_s = (double *)malloc(sizeof(double) * (sf + 1) * n);
        uint32_t *h_e = (uint32_t *)malloc(sizeof(uint32_t) * n);
        CK(cudaMemcpy(h_s, io.state, sizeof(double) * sf * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_s + (size_t)sf * n, io.time, sizeof(double) * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_e, io.episode, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        FILE *f = fopen(dump, "wb");
        if (!f) return 5;
        fwrite(h_s, sizeof(double), (size_t)(sf + 1) * n, f);
        fwrite(h_e, sizeof(uint32_t), n, f);
        fclose(f);
    }
    return 0;
}
