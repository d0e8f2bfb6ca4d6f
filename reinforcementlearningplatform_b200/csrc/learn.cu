// learn.cu -- K-LEARN: the PPO2 / DPPO2 network update on device.
//
// Replaces the body of the mini-batch loop of Proximal_Policy_Optimization2.learn (algorithm/policy_base/
// Proximal_Policy_Optimization2.py:102-131) -- actor forward, Normal.log_prob, ratio, clipped surrogate, backward,
// clip_grad_norm_, Adam; critic forward, mse_loss, backward, clip, Adam -- which round 1 left to torch autograd (about
// forty library kernels per mini-batch, 0.24 s per learn() of 4096 x 64 samples against 8 ms for collecting them).
//
// ppo2_grad_kernel: one block works on 64-sample tiles of the mini-batch, two blocks per SM; the SMs' blocks are dealt to
// the two nets (actor / critic) in proportion to their measured cost per tile, so both nets' gradients come out of ONE
// launch.  Everything of a tile stays in shared memory:
//     H_l [64][K_l + 4]    input of layer l (H_0 = the gathered observations, H_l = tanh outputs), sample-major;
//     W_l [N_l][K_l + 4]   the layer's weights as nn.Linear stores them (zero-padded), loaded once per block;
// and the three products of a layer are register-tiled fp32 GEMMs whose operands are both read along their contiguous
// index with 16-byte loads (row pitch = width + 4 floats, so that the rows a warp touches fall into distinct banks):
//     forward   Z[s][n]  = sum_k H_l[s][k] W_l[n][k]          thread: 2 samples x (N/8) outputs
//     dW        dW[n][k] = sum_s dZ_l[s][n] H_l[s][k]         thread: RN x RK entries (4 x 4 for 64 x 64 .. 1 x 1 for 8 x 32:
//                                                             all 256 threads share every layer), added to the block's
//                                                             partial gradient (L2) once per tile by the thread that owns them
//     dH        dH[s][k] = sum_n dZ_l[s][n] W_l[n][k]         thread: 2 samples x (K/32) 4-vectors
// dZ_{l-1} = dH * (1 - H_l^2) overwrites H_l in place (its last reader was dW_l), so no separate gradient buffers exist.
// A block adds up its tiles in program order; ppo2_reduce_kernel (a second, wide launch: one thread per parameter) sums
// the per-block partials in block order: the result does not depend on scheduling (no floating-point atomics anywhere).
// History (profiles/r2/learn.md): 128-sample tiles with the weight gradients in 80 persistent registers and one block per
// SM measured the same at large batches and 25 % slower at 4096 samples (half as many tiles to spread over the SMs).
//
// adam_kernel: clip_grad_norm_ + Adam.step over flat parameter / gradient / moment buffers (one segment per net), the
// global norm recomputed in a fixed order by every block so that no grid-wide synchronisation is needed.
//
// Arithmetic is fp32 on the FMA pipe with CUDA's accurate tanhf / expf (the test compares gradients and the updated
// parameters with torch autograd: <= 1e-6).  Tensor cores are not used here on purpose: a 3xTF32 split would need the
// activations in two operand layouts per layer (K-major for forward / dH, sample-major for dW) and the mini-batches of
// the reference's configurations (4096 .. 16384 samples of a 6-64-64-32-8 net = 0.2 .. 0.9 GFLOP) are launch-latency
// bound either way.
#include "common.cuh"

namespace {

constexpr int LT = 256;     // threads per block
constexpr int TM = 64;      // samples per tile
constexpr int TM_LOG2 = 6;
constexpr int MI = TM / 32; // samples per thread in the sample-by-feature products
constexpr int BLOCKS_PER_SM = 2;
constexpr int LMAX = 4;     // layers per net
constexpr int WMAX = 64;    // widest layer / input
constexpr int GRID_CAP = 148 * BLOCKS_PER_SM;

struct LLayer {
    int K, N;            // padded: K to a multiple of 8, N to 8 / 16 / 32 / 64
    int k_real, n_real;
    int w_off, b_off;    // shared-memory float offsets: weights [N][K + 4], bias [N]
    int h_off;           // shared-memory float offset of the layer's input H_l [TM][K + 4]
    int g_w, g_b;        // offsets of weight / bias inside the net's flat gradient
    int rn, rk;          // register tile of the weight-gradient product (dw_layer)
    const float *w, *b;  // global parameters
};

struct LNet {
    int n_layers, P, out_act;
    int gout_off;        // dZ of the output layer [TM][N_out + 4]
    int smem_floats;
    LLayer L[LMAX];
};

struct LearnArgs {
    LNet net[2];                 // 0 actor, 1 critic
    int net_of_y[2];             // slot -> net: blocks 0 .. nblk[0] - 1 work on slot 0, the next nblk[1] on slot 1
    int nblk[2];
    int S, A;
    int64_t T, N, first, count;
    const float *s, *a, *a_lp, *adv, *v_target;
    const int64_t *index;
    uint64_t perm_key;
    int perm_half_bits;
    float eps_clip, inv_count, entropy_coef;
    float std_;
    const float *std_vec, *a_min, *a_max;
    float *partial[2];           // [gridDim.x][P + 4]
    float *grad[2];
    float *loss_out;             // [2]
    unsigned int *counter;       // [2]
};

// ------------------------------------------------------------------------------------------------ sample permutation
__host__ __device__ __forceinline__ uint32_t feistel_round(uint32_t r, uint32_t k) {
    uint32_t h = (r + k) * 0x9E3779B1u;
    h ^= h >> 15; h *= 0x85EBCA77u;
    h ^= h >> 13; h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}
// bijection of [0, 2^(2 * half_bits)) restricted to [0, B) by cycle walking (B > 2^(2 * half_bits - 2): < 4 walks expected)
__host__ __device__ __forceinline__ int64_t perm_index(uint64_t key, int64_t j, int64_t B, int half_bits) {
    const uint32_t mask = (1u << half_bits) - 1u;
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    uint64_t x = (uint64_t)j;
    do {
        uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const uint32_t f = feistel_round(r, (q & 1 ? k1 : k0) + 0x632BE5ABu * (uint32_t)q) & mask;
            const uint32_t t = l ^ f;
            l = r;
            r = t;
        }
        x = ((uint64_t)l << half_bits) | r;
    } while ((int64_t)x >= B);
    return (int64_t)x;
}
int half_bits_for(int64_t B) {
    int bits = 1;
    while (((int64_t)1 << bits) < B) ++bits;
    return (bits + 1) / 2 < 1 ? 1 : (bits + 1) / 2;
}

// ------------------------------------------------------------------------------------------------ tile GEMMs
// thread coordinates of the sample-by-feature products: sg = 0..31 -> samples sg + 32 i, ng = 0..7 -> features ng + 8 j
template <int NJ>
__device__ __forceinline__ void fwd_layer(const float *__restrict__ H, int ph, const float *__restrict__ W, int pw,
                                          const float *__restrict__ bias, int K, float *__restrict__ out, int po,
                                          bool hidden, int sg, int ng) {
    float acc[MI][NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0f;
    // one base pointer per operand row, hoisted: inside the (unrolled) k loop every load is base + immediate -- the
    // address arithmetic was 9 % of the kernel's instructions (LEA per LDS, profiles/r2/learn.md)
    const float *ha[MI], *wa[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) ha[i] = H + (sg + 32 * i) * ph;
#pragma unroll
    for (int j = 0; j < NJ; ++j) wa[j] = W + (ng + 8 * j) * pw;
#pragma unroll 4
    for (int k0 = 0; k0 < K; k0 += 4) {
        float4 a[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const float4 *>(ha[i] + k0);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(wa[j] + k0);
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                acc[i][j] = fmaf(a[i].x, w.x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, w.y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, w.z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, w.w, acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float b = bias[ng + 8 * j];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const float z = acc[i][j] + b;
            out[(sg + 32 * i) * po + ng + 8 * j] = hidden ? tanhf(z) : z;
        }
    }
}

// dH[s][k] = sum_n G[s][n] W[n][k], then dZ = dH * (1 - H^2) written over H.  k-vectors of 4: v = ng + 8 jv
template <int NV>
__device__ __forceinline__ void bwd_layer(const float *__restrict__ G, int pg, const float *__restrict__ W, int pw, int N,
                                          float *H, int ph, int KG, int sg, int ng) {
    if (ng >= KG) return;
    float4 acc[MI][NV];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int jv = 0; jv < NV; ++jv) acc[i][jv] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *wp = W + 4 * ng;
    const float *ga[MI];
#pragma unroll
    for (int i = 0; i < MI; ++i) ga[i] = G + (sg + 32 * i) * pg;
#pragma unroll 2
    for (int n0 = 0; n0 < N; n0 += 4) {
        float g[MI][4];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const float4 t = *reinterpret_cast<const float4 *>(ga[i] + n0);
            g[i][0] = t.x; g[i][1] = t.y; g[i][2] = t.z; g[i][3] = t.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int jv = 0; jv < NV; ++jv) {
                const float4 w = *reinterpret_cast<const float4 *>(wp + (n0 + q) * pw + 32 * jv);
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    acc[i][jv].x = fmaf(g[i][q], w.x, acc[i][jv].x);
                    acc[i][jv].y = fmaf(g[i][q], w.y, acc[i][jv].y);
                    acc[i][jv].z = fmaf(g[i][q], w.z, acc[i][jv].z);
                    acc[i][jv].w = fmaf(g[i][q], w.w, acc[i][jv].w);
                }
            }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int jv = 0; jv < NV; ++jv) {
            float4 *hp = reinterpret_cast<float4 *>(H + (sg + 32 * i) * ph + 4 * ng + 32 * jv);
            const float4 h = *hp;
            float4 d;
            d.x = acc[i][jv].x * (1.0f - h.x * h.x);
            d.y = acc[i][jv].y * (1.0f - h.y * h.y);
            d.z = acc[i][jv].z * (1.0f - h.z * h.z);
            d.w = acc[i][jv].w * (1.0f - h.w * h.w);
            *hp = d;
        }
}

// dW[RN ng .. ][RK kg .. ] += sum over the tile's samples of dZ[s][n] H[s][k]; db likewise for the kg == 0 threads.
// The RN x RK register tile is chosen per layer (host: fill_net) so that ALL 256 threads share a layer's N x K entries:
// 4 x 4 for 64 x 64, 2 x 4 for 32 x 64, 1 x 2 for 64 x 8, 1 x 1 for 8 x 32 -- with a fixed 4 x 4 tile the small layers
// kept one or two warps busy for 128 iterations while the other warps waited at the barrier (ncu: stall_barrier 1.36 per
// issued instruction, the largest stall of the first version).
template <int RN, int RK>
__device__ __forceinline__ void dw_layer(const float *__restrict__ G, int pg, const float *__restrict__ H, int ph, int ng,
                                         int kg, float (&dw)[16], float (&db)[4]) {
    const float *gp = G + RN * ng, *hp = H + RK * kg;
#pragma unroll 4
    for (int s = 0; s < TM; ++s) {
        float g[RN], h[RK];
        if (RN == 4) { const float4 t = *reinterpret_cast<const float4 *>(gp + s * pg); g[0] = t.x; g[1 % RN] = t.y; g[2 % RN] = t.z; g[3 % RN] = t.w; }
        else if (RN == 2) { const float2 t = *reinterpret_cast<const float2 *>(gp + s * pg); g[0] = t.x; g[1 % RN] = t.y; }
        else g[0] = gp[s * pg];
        if (RK == 4) { const float4 t = *reinterpret_cast<const float4 *>(hp + s * ph); h[0] = t.x; h[1 % RK] = t.y; h[2 % RK] = t.z; h[3 % RK] = t.w; }
        else if (RK == 2) { const float2 t = *reinterpret_cast<const float2 *>(hp + s * ph); h[0] = t.x; h[1 % RK] = t.y; }
        else h[0] = hp[s * ph];
#pragma unroll
        for (int a = 0; a < RN; ++a)
#pragma unroll
            for (int c = 0; c < RK; ++c) dw[a * RK + c] = fmaf(g[a], h[c], dw[a * RK + c]);
        if (kg == 0) {
#pragma unroll
            for (int a = 0; a < RN; ++a) db[a] += g[a];
        }
    }
}

// the thread's RN x RK entries of dW (and RN of db) -> this block's partial gradient, torch parameter layout
// `first`: this is the block's first tile (the partial buffer is not cleared between launches); later tiles add to what
// the SAME thread stored before, in tile order -- no atomics, bit-reproducible.  The accumulators live in the block's
// partial (L2) instead of registers since the tile shrank to 64 samples for two blocks per SM: 80 persistent registers
// per thread did not fit under the 128 of a 2 x 256-thread SM.
template <int RN, int RK>
__device__ __forceinline__ void dw_store(const LLayer &Ly, float *part, int ng, int kg, const float (&dw)[16],
                                         const float (&db)[4], bool first) {
    float old[RN * RK], oldb[RN];
#pragma unroll
    for (int a = 0; a < RN; ++a) {
        const int n = RN * ng + a;
#pragma unroll
        for (int c = 0; c < RK; ++c)
            old[a * RK + c] = (!first && n < Ly.n_real && RK * kg + c < Ly.k_real) ? __ldcg(part + Ly.g_w + n * Ly.k_real + RK * kg + c) : 0.0f;
        oldb[a] = (!first && kg == 0 && n < Ly.n_real) ? __ldcg(part + Ly.g_b + n) : 0.0f;
    }
#pragma unroll
    for (int a = 0; a < RN; ++a) {
        const int n = RN * ng + a;
        if (n < Ly.n_real) {
#pragma unroll
            for (int c = 0; c < RK; ++c)
                if (RK * kg + c < Ly.k_real) __stcg(part + Ly.g_w + n * Ly.k_real + RK * kg + c, old[a * RK + c] + dw[a * RK + c]);
            if (kg == 0) __stcg(part + Ly.g_b + n, oldb[a] + db[a]);
        }
    }
}

// -DLEARN_TRACE (tools/build_variant.sh): thread 0 of block 0 adds up the cycles between the phase barriers of its tiles
// (phase ids: 0 sample indices, 1 gather, 2 + l forward layer l, 6 loss, 7 + 2 l dW_l, 8 + 2 l dH_l + dZ write) ->
// b200_learn_trace_read.  How profiles/r2/learn.md's phase table was measured.  Compiled out of the shipped library.
#ifdef LEARN_TRACE
__device__ long long g_learn_trace[16];
#define LTRACE(id)                                                          \
    do {                                                                    \
        if (blockIdx.x == 0 && tid == 0) {                                  \
            const long long now = clock64();                                \
            g_learn_trace[id] += now - lt_prev;                             \
            lt_prev = now;                                                  \
        }                                                                   \
    } while (0)
#else
#define LTRACE(id) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ the gradient kernel
__global__ void __launch_bounds__(LT, BLOCKS_PER_SM) ppo2_grad_kernel(const __grid_constant__ LearnArgs A) {
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_t[TM], s_i[TM];
    __shared__ float s_act[16 * TM], s_alp[16 * TM], s_tgt[TM];   // the tile's actions, old log-probs [d][s]; adv / v_target
    __shared__ float s_cst[16][4];                                 // per action dimension: var, log std, gain, offset
    __shared__ float s_red[LT / 32];
    const int slot = (int)blockIdx.x >= A.nblk[0] ? 1 : 0;
    const int bx = (int)blockIdx.x - (slot ? A.nblk[0] : 0), nbx = A.nblk[slot];
    const int net_id = A.net_of_y[slot];
    const LNet &net = A.net[net_id];
    const int L = net.n_layers;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sg = (lane & 3) + 4 * warp, ng = lane >> 2;

    // weights and biases -> shared memory, zero-padded
    for (int l = 0; l < L; ++l) {
        const LLayer &Ly = net.L[l];
        const int pw = Ly.K + 4;
        for (int e = tid; e < Ly.N * pw; e += LT) {
            const int n = e / pw, k = e - n * pw;
            sm[Ly.w_off + e] = (n < Ly.n_real && k < Ly.k_real) ? __ldg(Ly.w + n * Ly.k_real + k) : 0.0f;
        }
        for (int n = tid; n < Ly.N; n += LT) sm[Ly.b_off + n] = n < Ly.n_real ? __ldg(Ly.b + n) : 0.0f;
    }

    if (net_id == 0 && tid < 16) {
        const int d = tid;
        const float sd = d < A.A ? (A.std_vec ? __ldg(A.std_vec + d) : A.std_) : 1.0f;
        float gain = 1.0f, off2 = 0.0f;
        if (net.out_act == 2 && d < A.A) {
            const float lo = __ldg(A.a_min + d), hi = __ldg(A.a_max + d);
            off2 = (lo + hi) / 2.0f;
            gain = hi - off2;
        }
        s_cst[d][0] = sd * sd; s_cst[d][1] = logf(sd); s_cst[d][2] = gain; s_cst[d][3] = off2;
    }
    float loss_acc = 0.0f;
    float *part = A.partial[net_id] + (size_t)bx * (size_t)(net.P + 4);

    const int64_t tiles = (A.count + TM - 1) / TM;
    const LLayer &L0 = net.L[0];
    const LLayer &LO = net.L[L - 1];
    float *gout = sm + net.gout_off;
    const int pgo = LO.N + 4;

#ifdef LEARN_TRACE
    long long lt_prev = clock64();
#endif
    for (int64_t tile = bx; tile < tiles; tile += nbx) {
        __syncthreads();   // the previous tile's last phase has finished with H and s_t / s_i
        LTRACE(15);
        if (tid < TM) {
            const int64_t j = tile * TM + tid;
            int t = -1, i = 0;
            if (j < A.count) {
                const int64_t b = A.index ? A.index[A.first + j] : perm_index(A.perm_key, A.first + j, A.T * A.N, A.perm_half_bits);
                t = (int)(b / A.N);
                i = (int)(b - (int64_t)t * A.N);
            }
            s_t[tid] = t;
            s_i[tid] = i;
        }
        __syncthreads();
        LTRACE(0);
        {   // gather: H_0[s][k] (observations) and the loss phase's inputs (actions, old log-probs, adv / v_target).
            // Thread = (sample s, field lane q): its fields are q, q + 4, ...; ALL its loads are issued before the first
            // store to shared memory -- the obvious load-store loop serialised six DRAM round trips per tile (phase
            // trace: 9.5 k of 82 k cycles per tile)
            static_assert(LT % TM == 0 && LT / TM == 4, "4 field lanes per sample");
            float *H0 = sm + L0.h_off;
            const int p0 = L0.K + 4;
            const int s = tid & (TM - 1), q = tid >> TM_LOG2;
            const int t = s_t[s];
            const int64_t i = s_i[s], tt = t < 0 ? 0 : t;
            float vo[WMAX / 4], va[4], vl[4], vt = 0.0f;
#pragma unroll
            for (int j = 0; j < WMAX / 4; ++j) {
                const int k = q + 4 * j;
                vo[j] = (t >= 0 && k < A.S) ? __ldg(A.s + (tt * A.S + k) * A.N + i) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = q + 4 * j;
                const bool on = net_id == 0 && t >= 0 && d < A.A;
                const int64_t off = (tt * A.A + (on ? d : 0)) * A.N + i;
                va[j] = on ? __ldg(A.a + off) : 0.0f;
                vl[j] = on ? __ldg(A.a_lp + off) : 0.0f;
            }
            if (q == 0 && t >= 0) vt = __ldg((net_id == 0 ? A.adv : A.v_target) + tt * A.N + i);
#pragma unroll
            for (int j = 0; j < WMAX / 4; ++j) {
                const int k = q + 4 * j;
                if (k < L0.K) H0[s * p0 + k] = vo[j];
            }
            if (net_id == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = q + 4 * j;
                    if (d < A.A) { s_act[d * TM + s] = va[j]; s_alp[d * TM + s] = vl[j]; }
                }
            }
            if (q == 0) s_tgt[s] = vt;
        }
        __syncthreads();
        LTRACE(1);
        // ---------------------------------------------------------------- forward
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
            if (l < L) {
                const LLayer &Ly = net.L[l];
                const bool last = l == L - 1;
                float *out = last ? gout : sm + net.L[l + 1 < LMAX ? l + 1 : l].h_off;
                const float *H = sm + Ly.h_off, *W = sm + Ly.w_off, *bs = sm + Ly.b_off;
                const int ph = Ly.K + 4, po = Ly.N + 4;
                switch (Ly.N >> 3) {
                case 1: fwd_layer<1>(H, ph, W, ph, bs, Ly.K, out, po, !last, sg, ng); break;
                case 2: fwd_layer<2>(H, ph, W, ph, bs, Ly.K, out, po, !last, sg, ng); break;
                case 4: fwd_layer<4>(H, ph, W, ph, bs, Ly.K, out, po, !last, sg, ng); break;
                default: fwd_layer<8>(H, ph, W, ph, bs, Ly.K, out, po, !last, sg, ng); break;
                }
                __syncthreads();
                LTRACE(2 + l);
            }
        }
        // ---------------------------------------------------------------- loss and its gradient at the net's output
        {   // two threads per sample: thread `half` takes the action dimensions half, half + 2, ...
            const int sx = tid >> 1, half = tid & 1;
            float *z = gout + (sx < TM ? sx : 0) * pgo;
            const bool live = sx < TM && s_t[sx < TM ? sx : 0] >= 0;
            if (sx >= TM) {
                // 2 x TM threads work in this phase
            } else if (net_id == 0) {
                // Normal(mean, std).log_prob(a) summed over dimensions, PPO2.py:106-112
                float lp = 0.0f, lp_old = 0.0f;
                float dm[8];   // d lp / d z_d of this thread's dimensions
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int d = 2 * q + half;
                    dm[q] = 0.0f;
                    if (d < A.A) {
                        const float act = s_act[d * TM + sx], var = s_cst[d][0];
                        lp_old += s_alp[d * TM + sx];
                        float m = z[d], dmdz = 1.0f;
                        if (net.out_act == 1) {
                            dmdz = m > 0.0f ? 1.0f : 0.0f;
                            m = fmaxf(m, 0.0f);
                        } else if (net.out_act == 2) {
                            const float th = tanhf(m), gain = s_cst[d][2];
                            m = th * gain + s_cst[d][3];
                            dmdz = gain * (1.0f - th * th);
                        }
                        const float diff = act - m;
                        lp += -(diff * diff) / (2.0f * var) - s_cst[d][1] - 0.91893853320467274178f;
                        dm[q] = diff / var * dmdz;
                    }
                }
                lp += __shfl_xor_sync(0xffffffffu, lp, 1);
                lp_old += __shfl_xor_sync(0xffffffffu, lp_old, 1);
                const float ratio = expf(lp - lp_old);
                const float adv = s_tgt[sx];
                const float lo = 1.0f - A.eps_clip, hi = 1.0f + A.eps_clip;
                const float surr1 = ratio * adv, surr2 = fminf(fmaxf(ratio, lo), hi) * adv;
                if (live && half == 0) loss_acc += -fminf(surr1, surr2);
                // d(-min(surr1, surr2)) / d ratio: -adv unless the clamp is active and selected
                const bool clamped = ratio < lo || ratio > hi;
                const float dldr = (clamped && !(surr1 < surr2)) ? 0.0f : -adv;
                const float c = live ? dldr * ratio * A.inv_count : 0.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int d = 2 * q + half;
                    if (d < LO.N) z[d] = d < A.A ? c * dm[q] : 0.0f;
                }
            } else if (half == 0) {
                const float e = z[0] - s_tgt[sx];
                if (live) loss_acc += e * e;
                z[0] = live ? 2.0f * e * A.inv_count : 0.0f;
                for (int d = 1; d < LO.N; ++d) z[d] = 0.0f;
            }
        }
        __syncthreads();
        LTRACE(6);
        // ---------------------------------------------------------------- backward
#pragma unroll
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l < L) {
                const LLayer &Ly = net.L[l];
                const float *G = (l == L - 1) ? gout : sm + net.L[l + 1 < LMAX ? l + 1 : l].h_off;
                float *H = sm + Ly.h_off;
                const int pg = Ly.N + 4, ph = Ly.K + 4;
                const int kgs = Ly.K >> 2;
                {
                    const int kt = Ly.K / Ly.rk, nt = Ly.N / Ly.rn;
                    if (tid < nt * kt) {
                        // tn runs fastest: the threads that also sum the bias gradient (tk == 0) are the first nt of
                        // the block, so the other warps skip those adds (they were 5 % of all issued instructions)
                        const int tk = tid / nt, tn = tid - tk * nt;
                        float dw[16], db[4];
#pragma unroll
                        for (int e = 0; e < 16; ++e) dw[e] = 0.0f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) db[e] = 0.0f;
                        const bool first = tile == bx;
                        switch (Ly.rn * 8 + Ly.rk) {
                        case 4 * 8 + 4: dw_layer<4, 4>(G, pg, H, ph, tn, tk, dw, db); dw_store<4, 4>(Ly, part, tn, tk, dw, db, first); break;
                        case 2 * 8 + 4: dw_layer<2, 4>(G, pg, H, ph, tn, tk, dw, db); dw_store<2, 4>(Ly, part, tn, tk, dw, db, first); break;
                        case 1 * 8 + 4: dw_layer<1, 4>(G, pg, H, ph, tn, tk, dw, db); dw_store<1, 4>(Ly, part, tn, tk, dw, db, first); break;
                        case 1 * 8 + 2: dw_layer<1, 2>(G, pg, H, ph, tn, tk, dw, db); dw_store<1, 2>(Ly, part, tn, tk, dw, db, first); break;
                        default: dw_layer<1, 1>(G, pg, H, ph, tn, tk, dw, db); dw_store<1, 1>(Ly, part, tn, tk, dw, db, first); break;
                        }
                    }
                }
                if (l > 0) {
                    __syncthreads();
                    LTRACE(7 + 2 * l);
                    if (kgs == 16) bwd_layer<2>(G, pg, sm + Ly.w_off, ph, Ly.N, H, ph, kgs, sg, ng);
                    else bwd_layer<1>(G, pg, sm + Ly.w_off, ph, Ly.N, H, ph, kgs, sg, ng);
                    __syncthreads();
                    LTRACE(8 + 2 * l);
                }
            }
        }
    }

    // block sum of the loss in a fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) s_red[warp] = loss_acc;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < LT / 32; ++w) t += s_red[w];
        part[net.P] = t;
    }
}

// Sum of the per-block partial gradients in block order: one thread per parameter (consecutive threads read consecutive
// addresses of every partial), the loads of eight partials in flight at a time.  A separate, wide launch: folded into
// the gradient kernel as "the last block to finish reduces" it serialised 64 x 6.9 k dependent L2 reads on one SM and
// cost three times the gradient computation itself (130 of 190 us per 16 k-sample mini-batch).
struct ReduceArgs {
    const float *partial[2];
    float *grad[2];
    float *loss_out;
    int P[2], first_block[2];   // net j owns blocks first_block[j] .. of the launch
    int nets, nblk[2];
    float inv_count;
    // entropy bonus of the fixed-std Gaussian (a constant of the loss value): slot `ent_slot` (-1: none) gets
    // -entropy_coef * sum_d (0.5 + 0.5 log(2 pi) + log std_d), PPO2.py:107,116
    int ent_slot, A;
    float ent_coef, std_;
    const float *std_vec;
};
__global__ void __launch_bounds__(256) ppo2_reduce_kernel(const __grid_constant__ ReduceArgs a) {
    const int j = (a.nets > 1 && (int)blockIdx.x >= a.first_block[1]) ? 1 : 0;
    const int p = ((int)blockIdx.x - a.first_block[j]) * 256 + (int)threadIdx.x;
    const int P = a.P[j];
    if (p > P) return;
    const size_t stride = (size_t)(P + 4);
    const float *base = a.partial[j] + p;
    float t = 0.0f;
    int b = 0;
    const int nblk = a.nblk[j];
    for (; b + 8 <= nblk; b += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldcg(base + (size_t)(b + q) * stride);
#pragma unroll
        for (int q = 0; q < 8; ++q) t += v[q];
    }
    for (; b < nblk; ++b) t += __ldcg(base + (size_t)b * stride);
    if (p < P) {
        a.grad[j][p] = t;
        return;
    }
    float loss = t * a.inv_count;
    if (j == a.ent_slot) {
        float h = 0.0f;
        for (int d = 0; d < a.A; ++d) h += 0.5f + 0.9189385332046727f + logf(a.std_vec ? a.std_vec[d] : a.std_);
        loss -= a.ent_coef * h;
    }
    a.loss_out[j] = loss;
}

// ------------------------------------------------------------------------------------------------ clip + Adam
struct AdamArgs {
    int64_t off[2], len[2];
    float lr[2];
    float *param, *m, *v;
    const float *grad;
    float one_minus_b1, b2, one_minus_b2, eps, bc1, bc2_sqrt, max_norm, grad_scale;
    float *norm_out;
};
constexpr int AT = 1024;

__global__ void __launch_bounds__(AT) adam_kernel(const __grid_constant__ AdamArgs a) {
    __shared__ float s_red[AT / 32];
    __shared__ float s_total;
    const int seg = blockIdx.y, tid = threadIdx.x;
    const int64_t off = a.off[seg], P = a.len[seg];
    const float *g = a.grad + off;
    // total_norm of clip_grad_norm_, every block in the same order
    float ss = 0.0f;
    for (int64_t p = tid; p < P; p += AT) {
        const float x = __ldcg(g + p) * a.grad_scale;
        ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < AT / 32; ++w) t += s_red[w];
        s_total = t;
    }
    __syncthreads();
    const float norm = sqrtf(s_total);
    float coef = a.grad_scale;
    if (a.max_norm > 0.0f) coef *= fminf(a.max_norm / (norm + 1e-6f), 1.0f);
    if (blockIdx.x == 0 && tid == 0 && a.norm_out) a.norm_out[seg] = norm;
    const float step_size = a.lr[seg] / a.bc1;
    float *pp = a.param + off, *pm = a.m + off, *pv = a.v + off;
    for (int64_t p = (int64_t)blockIdx.x * AT + tid; p < P; p += (int64_t)gridDim.x * AT) {
        const float gr = __ldcg(g + p) * coef;
        float m = pm[p], v = pv[p];
        m = m + (gr - m) * a.one_minus_b1;                  // exp_avg.lerp_(grad, 1 - beta1)
        v = v * a.b2 + a.one_minus_b2 * gr * gr;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
        pp[p] = pp[p] - step_size * (m / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
        pm[p] = m;
        pv[p] = v;
    }
}

// ------------------------------------------------------------------------------------------------ host side
int pad_n(int n) { return n <= 8 ? 8 : n <= 16 ? 16 : n <= 32 ? 32 : 64; }

int fill_net(const b200_mlp *m, bool is_actor, LNet *net) {
    *net = LNet{};
    if (m->n_layers < 1 || m->n_layers > LMAX) return B200ENV_ESIZE;
    net->n_layers = m->n_layers;
    net->out_act = m->out_act;
    int off = 0, P = 0;
    for (int l = 0; l < m->n_layers; ++l) {
        const int in = m->dims[l], out = m->dims[l + 1];
        if (in < 1 || out < 1 || in > WMAX || out > WMAX) return B200ENV_ESIZE;
        if (!m->w[l] || !m->b[l]) return B200ENV_ENULL;
        LLayer &Ly = net->L[l];
        const bool last = l + 1 == m->n_layers;
        if (last && out > 16) return B200ENV_ESIZE;
        Ly.k_real = in;
        Ly.n_real = out;
        Ly.K = l == 0 ? (in + 7) / 8 * 8 : net->L[l - 1].N;
        Ly.N = pad_n(out);
        Ly.w = m->w[l];
        Ly.b = m->b[l];
        Ly.g_w = P;
        Ly.g_b = P + in * out;
        P += in * out + out;
        {   // entries per thread e = N K / 256 (>= 1), as RN x RK with RK the widest vector load that fits
            int e = 1;   // the smallest power of two with N K / e <= 256 threads (N is a power of two, K a multiple of 8)
            while (e < 16 && Ly.N * Ly.K > e * LT) e *= 2;
            Ly.rk = e >= 4 ? 4 : e;
            Ly.rn = e / Ly.rk;
        }
        Ly.w_off = off; off += Ly.N * (Ly.K + 4);
        Ly.b_off = off; off += Ly.N;
    }
    for (int l = 0; l < m->n_layers; ++l) {
        net->L[l].h_off = off;
        off += TM * (net->L[l].K + 4);
    }
    net->gout_off = off;
    off += TM * (net->L[m->n_layers - 1].N + 4);
    net->P = P;
    net->smem_floats = off;
    (void)is_actor;
    return B200ENV_OK;
}

int grid_cap() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        v *= BLOCKS_PER_SM;
        sms[dev] = v > GRID_CAP ? GRID_CAP : v;
    }
    return sms[dev];
}

size_t partial_floats(const LNet &n) { return (size_t)GRID_CAP * (size_t)(n.P + 4); }

struct Plan {
    LearnArgs a;
    int nets;
    size_t smem;
};

int make_plan(const b200_ppo2_batch *bt, const b200_mlp *actor, const b200_mlp *critic, float std_, const float *std_vec,
              const float *a_min, const float *a_max, float eps_clip, float entropy_coef, float *grad_actor,
              float *grad_critic, float *loss_out, void *workspace, size_t workspace_bytes, Plan *pl) {
    if (!bt || (!actor && !critic) || !loss_out || !workspace) return B200ENV_ENULL;
    if (!bt->s || (actor && !bt->adv) || (critic && !bt->v_target)) return B200ENV_ENULL;
    if (bt->T <= 0 || bt->N <= 0 || bt->N >= ((int64_t)1 << 31) || bt->T >= ((int64_t)1 << 31)) return B200ENV_ESIZE;
    const int64_t B = bt->T * bt->N;
    if (bt->count <= 0 || bt->first < 0 || (!bt->index && bt->first + bt->count > B)) return B200ENV_ESIZE;
    LearnArgs &a = pl->a;
    a = LearnArgs{};
    int rc;
    pl->nets = 0;
    size_t smem_f = 0, ws = 64;
    if (actor) {
        if (!grad_actor || !bt->a || !bt->a_lp) return B200ENV_ENULL;
        if (!std_vec && !(std_ > 0.0f)) return B200ENV_EPARAMS;
        if (actor->out_act < 0 || actor->out_act > 2) return B200ENV_EPARAMS;
        if (actor->out_act == 2 && (!a_min || !a_max)) return B200ENV_ENULL;
        if ((rc = fill_net(actor, true, &a.net[0]))) return rc;
        a.net_of_y[pl->nets++] = 0;
        smem_f = a.net[0].smem_floats;
        a.partial[0] = reinterpret_cast<float *>(static_cast<char *>(workspace) + ws);
        ws += partial_floats(a.net[0]) * 4;
        a.S = actor->dims[0];
        a.A = actor->dims[actor->n_layers];
    }
    if (critic) {
        if (!grad_critic) return B200ENV_ENULL;
        if (critic->dims[critic->n_layers > 0 && critic->n_layers <= 4 ? critic->n_layers : 0] != 1) return B200ENV_EPARAMS;
        if (actor && actor->dims[0] != critic->dims[0]) return B200ENV_EPARAMS;
        if ((rc = fill_net(critic, false, &a.net[1]))) return rc;
        a.net[1].out_act = 0;
        a.net_of_y[pl->nets++] = 1;
        if ((size_t)a.net[1].smem_floats > smem_f) smem_f = a.net[1].smem_floats;
        a.partial[1] = reinterpret_cast<float *>(static_cast<char *>(workspace) + ws);
        ws += partial_floats(a.net[1]) * 4;
        a.S = critic->dims[0];
    }
    if (ws > workspace_bytes) return B200ENV_ESIZE;
    pl->smem = smem_f * sizeof(float);
    if (pl->smem > 227 * 1024 - 12 * 1024) return B200ENV_ESIZE;   // ~10 KB of static shared memory next to it
    a.T = bt->T; a.N = bt->N; a.first = bt->first; a.count = bt->count;
    a.s = bt->s; a.a = bt->a; a.a_lp = bt->a_lp; a.adv = bt->adv; a.v_target = bt->v_target;
    a.index = bt->index;
    a.perm_key = bt->perm_key;
    a.perm_half_bits = half_bits_for(B);
    a.eps_clip = eps_clip;
    a.inv_count = 1.0f / (float)bt->count;
    a.std_ = std_; a.std_vec = std_vec; a.a_min = a_min; a.a_max = a_max;
    a.entropy_coef = entropy_coef;
    a.grad[0] = grad_actor; a.grad[1] = grad_critic;
    a.loss_out = loss_out;
    a.counter = static_cast<unsigned int *>(workspace);
    return B200ENV_OK;
}

int launch_grad(Plan &pl, cudaStream_t stream) {
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (pl.smem > 48 * 1024 && pl.smem > configured[dev]) {
        if (cudaFuncSetAttribute(ppo2_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess)
            return b200_check_launch();
        configured[dev] = pl.smem;
    }
    const int64_t tiles = (pl.a.count + TM - 1) / TM;
    // one block per SM; the SMs are dealt to the two nets in proportion to their work per sample (3 x sum K N), so that
    // the actor's and the critic's blocks finish together (an even split left the critic's SMs idle 60 % of the time)
    const int cap = grid_cap();
    double work[2] = {0.0, 0.0};
    for (int y = 0; y < pl.nets; ++y) {
        const LNet &n = pl.a.net[pl.a.net_of_y[y]];
        for (int l = 0; l < n.n_layers; ++l) work[y] += (double)n.L[l].K * n.L[l].N;
        work[y] += 2200.0;   // per-tile fixed part (indices, gather, loss, barriers) in units of one K x N product term:
                             // phase trace, ~20 k of an actor tile's 82 k cycles; with 64 the critic's blocks were the tail
    }
    pl.a.nblk[1] = 0;
    if (pl.nets == 2) {
        int nb1 = (int)(cap * work[1] / (work[0] + work[1]) + 0.5);
        nb1 = nb1 < 1 ? 1 : (nb1 > cap - 1 ? cap - 1 : nb1);
        pl.a.nblk[0] = (int)(tiles < cap - nb1 ? tiles : cap - nb1);
        pl.a.nblk[1] = (int)(tiles < nb1 ? tiles : nb1);
    } else {
        pl.a.nblk[0] = (int)(tiles < cap ? tiles : cap);
    }
    ppo2_grad_kernel<<<pl.a.nblk[0] + pl.a.nblk[1], LT, pl.smem, stream>>>(pl.a);
    ReduceArgs r = {};
    int blocks = 0;
    for (int y = 0; y < pl.nets; ++y) {
        const int id = pl.a.net_of_y[y];
        r.partial[y] = pl.a.partial[id];
        r.grad[y] = pl.a.grad[id];
        r.P[y] = pl.a.net[id].P;
        r.first_block[y] = blocks;
        blocks += (pl.a.net[id].P + 1 + 255) / 256;
    }
    // loss_out is indexed by net id (0 actor, 1 critic): with a single net present its slot is selected here
    r.loss_out = pl.a.loss_out + (pl.nets == 1 ? pl.a.net_of_y[0] : 0);
    r.nets = pl.nets;
    r.nblk[0] = pl.a.nblk[0]; r.nblk[1] = pl.a.nblk[1];
    r.inv_count = pl.a.inv_count;
    r.ent_slot = (pl.a.net_of_y[0] == 0 && pl.a.entropy_coef != 0.0f) ? 0 : -1;
    r.A = pl.a.A; r.ent_coef = pl.a.entropy_coef; r.std_ = pl.a.std_; r.std_vec = pl.a.std_vec;
    ppo2_reduce_kernel<<<blocks, 256, 0, stream>>>(r);
    return b200_check_launch();
}

int launch_adam(int n_seg, const int64_t *seg_off, const int64_t *seg_len, const float *lr, float *param, const float *grad,
                float *m, float *v, int64_t step, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                float *norm_out, cudaStream_t stream) {
    if (n_seg < 1 || n_seg > 2 || !seg_off || !seg_len || !lr) return B200ENV_EPARAMS;
    if (!param || !grad || !m || !v) return B200ENV_ENULL;
    if (step < 1) return B200ENV_EPARAMS;
    AdamArgs a = {};
    int64_t longest = 0;
    for (int j = 0; j < n_seg; ++j) {
        if (seg_len[j] <= 0 || seg_off[j] < 0) return B200ENV_ESIZE;
        a.off[j] = seg_off[j]; a.len[j] = seg_len[j]; a.lr[j] = lr[j];
        if (seg_len[j] > longest) longest = seg_len[j];
    }
    a.param = param; a.m = m; a.v = v; a.grad = grad;
    a.one_minus_b1 = (float)(1.0 - (double)beta1);
    a.b2 = beta2;
    a.one_minus_b2 = (float)(1.0 - (double)beta2);
    a.eps = eps;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.max_norm = max_norm;
    a.grad_scale = grad_scale;
    a.norm_out = norm_out;
    int64_t bx = (longest + (int64_t)AT * 4 - 1) / ((int64_t)AT * 4);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    adam_kernel<<<dim3((unsigned)bx, (unsigned)n_seg), AT, 0, stream>>>(a);
    return b200_check_launch();
}

} // namespace

extern "C" B200_API size_t b200_ppo2_workspace_bytes(const b200_mlp *actor, const b200_mlp *critic) {
    size_t ws = 64;
    LNet n;
    if (actor) {
        if (fill_net(actor, true, &n)) return 0;
        ws += partial_floats(n) * 4;
    }
    if (critic) {
        if (fill_net(critic, false, &n)) return 0;
        ws += partial_floats(n) * 4;
    }
    return (actor || critic) ? ws : 0;
}

extern "C" B200_API int b200_ppo2_grad(const b200_ppo2_batch *batch, const b200_mlp *actor, const b200_mlp *critic,
                                       float std_, const float *std_vec, const float *a_min, const float *a_max,
                                       float eps_clip, float entropy_coef, float *grad_actor, float *grad_critic,
                                       float *loss_out, void *workspace, size_t workspace_bytes, void *cuda_stream) {
    Plan pl;
    int rc = make_plan(batch, actor, critic, std_, std_vec, a_min, a_max, eps_clip, entropy_coef, grad_actor, grad_critic,
                       loss_out, workspace, workspace_bytes, &pl);
    if (rc) return rc;
    return launch_grad(pl, (cudaStream_t)cuda_stream);
}

extern "C" B200_API int b200_adam_step(int n_seg, const int64_t *seg_off, const int64_t *seg_len, const float *lr,
                                       float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t step,
                                       float beta1, float beta2, float eps, float max_norm, float grad_scale,
                                       float *grad_norm_out, void *cuda_stream) {
    return launch_adam(n_seg, seg_off, seg_len, lr, param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, eps, max_norm,
                       grad_scale, grad_norm_out, (cudaStream_t)cuda_stream);
}

extern "C" B200_API int b200_ppo2_learn(const b200_ppo2_batch *batch, const b200_mlp *actor, const b200_mlp *critic,
                                        float std_, const float *std_vec, const float *a_min, const float *a_max,
                                        float eps_clip, float entropy_coef, int k_epochs, int64_t mini_batch,
                                        float lr_actor, float lr_critic, float beta1, float beta2, float adam_eps,
                                        float max_norm, int64_t step, float *param, float *grad, float *exp_avg,
                                        float *exp_avg_sq, float *loss_out, void *workspace, size_t workspace_bytes,
                                        void *cuda_stream) {
    if (!batch || !actor || !critic || !param || !grad || !exp_avg || !exp_avg_sq) return B200ENV_ENULL;
    if (batch->index) return B200ENV_EPARAMS;   // the loop draws its own permutation per epoch
    if (k_epochs < 1 || mini_batch < 1 || step < 1) return B200ENV_EPARAMS;
    const int64_t B = batch->T * batch->N;
    LNet na, nc;
    int rc;
    if ((rc = fill_net(actor, true, &na)) || (rc = fill_net(critic, false, &nc))) return rc;
    const int64_t seg_off[2] = {0, na.P}, seg_len[2] = {na.P, nc.P};
    const float lr[2] = {lr_actor, lr_critic};
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    b200_ppo2_batch bt = *batch;
    Plan pl;
    for (int e = 0; e < k_epochs; ++e) {
        bt.perm_key = batch->perm_key + (uint64_t)e;
        for (int64_t first = 0; first < B; first += mini_batch) {
            bt.first = first;
            bt.count = B - first < mini_batch ? B - first : mini_batch;
            if ((rc = make_plan(&bt, actor, critic, std_, std_vec, a_min, a_max, eps_clip, entropy_coef, grad, grad + na.P,
                                loss_out, workspace, workspace_bytes, &pl))) return rc;
            if ((rc = launch_grad(pl, stream))) return rc;
            if ((rc = launch_adam(2, seg_off, seg_len, lr, param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, adam_eps,
                                  max_norm, 1.0f, nullptr, stream))) return rc;
            ++step;
        }
    }
    return B200ENV_OK;
}

extern "C" B200_API int b200_ppo2_permutation(uint64_t perm_key, int64_t B, int64_t first, int64_t count, int64_t *out_host) {
    if (B <= 0 || first < 0 || count < 0 || first + count > B) return B200ENV_ESIZE;
    if (!out_host && count) return B200ENV_ENULL;
    const int hb = half_bits_for(B);
    for (int64_t j = 0; j < count; ++j) out_host[j] = perm_index(perm_key, first + j, B, hb);
    return B200ENV_OK;
}

#ifdef LEARN_TRACE
extern "C" B200_API int b200_learn_trace_read(long long *host) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host, g_learn_trace, sizeof(long long) * 16);
    long long zero[16] = {0};
    cudaMemcpyToSymbol(g_learn_trace, zero, sizeof(zero));
    return 0;
}
#endif
