// policy_common.cuh -- pieces shared by the two K-POLICY kernels (policy.cu: FP32 FMA pipe; policy_tc.cu: tensor cores).
#pragma once
#include "common.cuh"

struct PolicyIO {
    const float *obs;            // [S][n]
    const float *a_min, *a_max;  // [A] device
    const float *noise;          // [A][n] or NULL
    const float *std_vec;        // [A] device per-dimension std (the DPPO2 demos' init_std = range / 6 vector), or NULL
    int out_affine;              // actor out_act == 2: mean = tanh(z) * gain + off, gain = a_max - off, off = (a_min + a_max) / 2
                                 // (demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance/train.py:42-43,66-69)
    float std_;
    uint64_t seed, step;
    int64_t off;
    float *action, *log_prob, *mean, *value;
};

int policy_launch_fp32(int64_t n, const b200_mlp *actor, const b200_mlp *critic, const PolicyIO &io, cudaStream_t s);
int policy_launch_tc(int64_t n, const b200_mlp *actor, const b200_mlp *critic, const PolicyIO &io, cudaStream_t s);
// policy_umma.cu (tcgen05 / TMEM): weights pre-packed into a caller-owned workspace
size_t policy_umma_workspace_bytes(const b200_mlp *actor, const b200_mlp *critic);
int policy_umma_pack(const b200_mlp *actor, const b200_mlp *critic, void *workspace, size_t bytes, cudaStream_t s);
int policy_launch_umma(int64_t n, const b200_mlp *actor, const b200_mlp *critic, const void *workspace, size_t bytes,
                       const PolicyIO &io, cudaStream_t s);
int policy_umma_probe(const float *A, const float *W, float *D, int N, int K, int three_pass, cudaStream_t s);

// Philox key domain of the exploration noise.  The env resets draw from Philox(seed, instance, episode) (common.cuh); with
// equal seeds the reset of (instance i, episode e) and the noise of (instance i, step e) would consume the same 128-bit
// block, so the policy stream flips key bits the env stream never does ("POLI").
#define B200_POLICY_KEY_DOMAIN 0x504F4C49u

// Proximal_Policy_Optimization2.choose_action (:72-75) for instance i given its A means (mean j at src[j * stride]):
// a = clamp(mean + std * eps), log_prob = Normal(mean, std).log_prob(a).  eps: injected noise or Philox + Box-Muller.
// `mean_of(j)` returns the j-th pre-head output of instance i.
template <typename MeanFn>
__device__ __forceinline__ void policy_sample_store_fn(const PolicyIO &a, int64_t n, int64_t i, int A, MeanFn mean_of) {
    Philox rng(a.seed, (uint64_t)(a.off + i), (uint32_t)a.step);
    rng.k1 ^= B200_POLICY_KEY_DOMAIN;
    rng.c3 = (uint32_t)(a.step >> 32) << 8; // high step bits above the block counter
    float lstd = logf(a.std_), var2 = 2.0f * (a.std_ * a.std_), sd = a.std_;
    for (int j = 0; j < A; j += 2) {
        float e0, e1;
        if (a.noise) {
            e0 = __ldcs(a.noise + (int64_t)j * n + i);
            e1 = j + 1 < A ? __ldcs(a.noise + (int64_t)(j + 1) * n + i) : 0.0f;
        } else { // Box-Muller on two 32-bit uniforms; u1 in (0, 1]
            if (rng.have < 2) rng.block();
            const uint32_t r0 = rng.r[4 - rng.have], r1 = rng.r[5 - rng.have];
            rng.have -= 2;
            const float u1 = ((float)(r0 >> 8) + 1.0f) * (1.0f / 16777216.0f);
            const float u2 = (float)(r1 >> 8) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincospif(2.0f * u2, &sn, &cs);
            e0 = rad * cs;
            e1 = rad * sn;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (j + q >= A) break;
            float m = mean_of(j + q);
            const float amax = __ldg(a.a_max + j + q), amin = __ldg(a.a_min + j + q);
            if (a.out_affine) {
                const float off = (amin + amax) / 2.0f;
                m = tanhf(m) * (amax - off) + off;
            }
            if (a.std_vec) {
                sd = __ldg(a.std_vec + j + q);
                lstd = logf(sd);
                var2 = 2.0f * (sd * sd);
            }
            float act = fmaf(sd, q ? e1 : e0, m);                           // dist.sample()
            act = fmaxf(fminf(act, amax), amin);
            const float d = act - m;                                         // Normal.log_prob
            const float lp = -(d * d) / var2 - lstd - 0.91893853320467274178f;
            __stcs(a.action + (int64_t)(j + q) * n + i, act);
            if (a.log_prob) __stcs(a.log_prob + (int64_t)(j + q) * n + i, lp);
            if (a.mean) __stcs(a.mean + (int64_t)(j + q) * n + i, m);
        }
    }
}

__device__ __forceinline__ void policy_sample_store(const PolicyIO &a, int64_t n, int64_t i, int A, const float *src,
                                                    int stride) {
    policy_sample_store_fn(a, n, i, A, [src, stride](int j) { return src[j * stride]; });
}
