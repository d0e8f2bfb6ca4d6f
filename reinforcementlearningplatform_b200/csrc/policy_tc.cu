// policy_tc.cu -- K-POLICY on the tensor cores (3xTF32): same contract as policy.cu, ~fp32 accuracy.
//
// Why tensor cores here: the FP32-FMA version of this kernel takes 1.16 ms for 1 M instances of the reference's
// 6-64-64-32-8 actor + 6-64-32-1 critic (profiles/r1: 20.4 k instructions per instance, 57 % of them FFMA, issue 58 %),
// four times the UavFntsmcParam step kernel it feeds -- the measurement BASELINE.json's north_star asks for before
// tensor cores may be used.  The layers are tiny (K, N <= 64) and every layer ends in a tanh over all outputs, so the
// kernel keeps the register-fragment form (warp-level m16n8k8 TF32 MMA, operands from shared memory, accumulators in
// registers, epilogue in place); a TMA/tcgen05/TMEM tile pipeline has nothing to stream at these sizes.
//
// 3xTF32: every fp32 operand is split on the fly into hi = tf32(x) and lo = tf32(x - hi); D += A_lo B_hi + A_hi B_lo +
// A_hi B_hi with fp32 accumulation recovers fp32-level products (the dropped A_lo B_lo term is 2^-22 relative).
// Measured against the reference nets: <= 5e-6 absolute on mean / value (tests/test_policy.py).
//
// Mapping: M = instances, N = layer outputs, K = layer inputs.  A block of 256 threads owns a tile of 256 instances,
// each warp 32 of them (two 16-row M tiles) for ALL layers: its activations live in a private [32][68] shared-memory
// region (row stride 68 floats: the A-fragment loads of a warp hit 32 different banks) that every layer overwrites in
// place after a __syncwarp, so no block-level barrier is needed inside the tile loop.  Weights are staged once per
// (persistent) block in B-fragment order: the pair (W[n][k], W[n][k + 4]) a lane needs for one MMA is one 8-byte load,
// consecutive lanes consecutive addresses.
#include "policy_common.cuh"

namespace {

constexpr int TB = 256;           // threads per block
constexpr int WARPS = TB / 32;
constexpr int LD = 68;            // activation row stride (floats)
constexpr int TC_MAX_LAYERS = 4;
constexpr int TC_MAX_DIM = 64;

struct TcNet {
    int n_layers;
    int dims[TC_MAX_LAYERS + 1];
    int ks[TC_MAX_LAYERS];        // K steps of 8 (inputs padded to a multiple of 8)
    int nt[TC_MAX_LAYERS];        // N tiles of 8 (outputs padded to 8, 16, 32 or 64)
    int w_off[TC_MAX_LAYERS];     // float offsets into the arena: fragment-ordered weights [ks][nt][32 lanes][2]
    int b_off[TC_MAX_LAYERS];     // [nt * 8] biases
    int out_act;
    const float *w[TC_MAX_LAYERS];
    const float *b[TC_MAX_LAYERS];
};

struct TcArgs {
    TcNet actor, critic;
    int has_actor, has_critic;
    int arena_floats;
    PolicyIO io;
};

// round-to-nearest (ties away) to TF32's 10-bit mantissa on the bit pattern: what cvt.rna.tf32.f32 does for finite
// values, in two integer instructions instead of the multi-instruction sequence ptxas emits for the cvt
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2.approx, rcp.approx): 5 instructions, absolute error <= 4e-7 over the
// whole range (the quotient saturates to 0 / 2 for large |x|), against ~25 instructions for the 1-ulp tanhf.  The
// 3xTF32 products carry ~1e-6 themselves, so the accurate version would buy nothing here (policy.cu keeps it).
__device__ __forceinline__ float tanh_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f)); // exp(2x) = 2^(2 x log2 e)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));               // e = inf -> 0, e = 0 -> 1
    return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void stage_tc(const TcNet &nd, float *arena) {
    for (int l = 0; l < nd.n_layers; ++l) {
        const int in = nd.dims[l], out = nd.dims[l + 1], KS = nd.ks[l], NT = nd.nt[l];
        float2 *wf = reinterpret_cast<float2 *>(arena + nd.w_off[l]);
        for (int e = threadIdx.x; e < KS * NT * 32; e += TB) {
            const int lane = e & 31, tile = e >> 5, nt = tile % NT, ks = tile / NT;
            const int g = lane >> 2, t = lane & 3;
            const int nn = nt * 8 + g, k0 = ks * 8 + t, k1 = k0 + 4;
            float2 v;
            v.x = (nn < out && k0 < in) ? __ldg(nd.w[l] + (int64_t)nn * in + k0) : 0.0f; // b0: (k = t,     n = g)
            v.y = (nn < out && k1 < in) ? __ldg(nd.w[l] + (int64_t)nn * in + k1) : 0.0f; // b1: (k = t + 4, n = g)
            wf[e] = v;
        }
        float *bs = arena + nd.b_off[l];
        for (int j = threadIdx.x; j < NT * 8; j += TB) bs[j] = j < out ? __ldg(nd.b[l] + j) : 0.0f;
    }
}

// one layer for the warp's 32 instances, in place in `act` ([32][LD]); NT = output tiles of 8
template <int NT>
__device__ __forceinline__ void layer_tc(const float2 *__restrict__ wf, const float *__restrict__ bs, int KS,
                                         float *__restrict__ act, int actfn) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float d[2][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float b0 = bs[nt * 8 + 2 * t], b1 = bs[nt * 8 + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { d[mt][nt][0] = b0; d[mt][nt][1] = b1; d[mt][nt][2] = b0; d[mt][nt][3] = b1; }
    }
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float *r0 = act + (mt * 16 + g) * LD + ks * 8 + t, *r1 = r0 + 8 * LD;
            split_tf32(r0[0], ahi[mt][0], alo[mt][0]);  // a0: (row g,     col t)
            split_tf32(r1[0], ahi[mt][1], alo[mt][1]);  // a1: (row g + 8, col t)
            split_tf32(r0[4], ahi[mt][2], alo[mt][2]);  // a2: (row g,     col t + 4)
            split_tf32(r1[4], ahi[mt][3], alo[mt][3]);  // a3: (row g + 8, col t + 4)
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float2 w = wf[(ks * NT + nt) * 32 + lane];
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(w.x, bh0, bl0);
            split_tf32(w.y, bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_tf32(d[mt][nt], alo[mt], bh0, bh1);  // small terms first
                mma_tf32(d[mt][nt], ahi[mt], bl0, bl1);
                mma_tf32(d[mt][nt], ahi[mt], bh0, bh1);
            }
        }
    }
    __syncwarp(); // every lane has read its A fragments: the region may be overwritten
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = d[mt][nt][q];
                if (actfn == 2) v[q] = tanh_fast(v[q]);
                else if (actfn == 1) v[q] = fmaxf(v[q], 0.0f);
            }
            float *r0 = act + (mt * 16 + g) * LD + nt * 8 + 2 * t; // c0, c1: (row g, cols 2t, 2t + 1)
            *reinterpret_cast<float2 *>(r0) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2 *>(r0 + 8 * LD) = make_float2(v[2], v[3]); // c2, c3: row g + 8
        }
    __syncwarp();
}

__device__ __forceinline__ void run_tc(const TcNet &nd, const float *arena, float *act) {
    for (int l = 0; l < nd.n_layers; ++l) {
        const int fn = l + 1 < nd.n_layers ? 2 : nd.out_act;
        const float2 *wf = reinterpret_cast<const float2 *>(arena + nd.w_off[l]);
        const float *bs = arena + nd.b_off[l];
        switch (nd.nt[l]) {
        case 8: layer_tc<8>(wf, bs, nd.ks[l], act, fn); break;
        case 4: layer_tc<4>(wf, bs, nd.ks[l], act, fn); break;
        case 2: layer_tc<2>(wf, bs, nd.ks[l], act, fn); break;
        default: layer_tc<1>(wf, bs, nd.ks[l], act, fn); break;
        }
    }
}

// observations of the warp's 32 instances into its region, columns S .. 8 * ks0 - 1 zeroed (they meet zero weights,
// but 0 * garbage must not be NaN)
__device__ __forceinline__ void load_obs(const PolicyIO &io, int64_t n, int64_t i, bool live, int S, int kpad, float *act) {
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < kpad; ++k) act[lane * LD + k] = (live && k < S) ? __ldg(io.obs + (int64_t)k * n + i) : 0.0f;
    __syncwarp();
}

__global__ void __launch_bounds__(TB, 2)
policy_forward_tc_kernel(const __grid_constant__ TcArgs a, int64_t n) {
    extern __shared__ __align__(16) float smem[];
    float *arena = smem;
    float *act = smem + a.arena_floats + (threadIdx.x >> 5) * (32 * LD); // this warp's [32][LD] region
    if (a.has_actor) stage_tc(a.actor, arena);
    if (a.has_critic) stage_tc(a.critic, arena);
    __syncthreads();
    const int S = a.has_actor ? a.actor.dims[0] : a.critic.dims[0];
    const int kpad = (a.has_actor ? a.actor.ks[0] : a.critic.ks[0]) * 8;
    const int lane = threadIdx.x & 31;
    const int64_t tiles = (n + TB - 1) / TB;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t i = tile * TB + threadIdx.x; // lane l of a warp owns row l of the warp's region
        const bool live = i < n;
        if (a.has_actor) {
            load_obs(a.io, n, i, live, S, kpad, act);
            run_tc(a.actor, arena, act);
            if (live) policy_sample_store(a.io, n, i, a.actor.dims[a.actor.n_layers], act + lane * LD, 1);
            __syncwarp();
        }
        if (a.has_critic) {
            load_obs(a.io, n, i, live, S, kpad, act);
            run_tc(a.critic, arena, act);
            if (live) __stcs(a.io.value + i, act[lane * LD]);
            __syncwarp();
        }
    }
}

int fill_tc(const b200_mlp *m, TcNet *nd, int *arena) {
    if (m->n_layers < 1 || m->n_layers > TC_MAX_LAYERS) return B200ENV_ESIZE;
    nd->n_layers = m->n_layers;
    nd->out_act = m->out_act == 2 ? 0 : m->out_act; // 2 (tanh range map) is applied by policy_sample_store
    for (int l = 0; l <= m->n_layers; ++l) {
        if (m->dims[l] < 1 || m->dims[l] > TC_MAX_DIM) return B200ENV_ESIZE; // wider layers: not supported (no fallback)
        nd->dims[l] = m->dims[l];
    }
    for (int l = 0; l < m->n_layers; ++l) {
        if (!m->w[l] || !m->b[l]) return B200ENV_ENULL;
        nd->w[l] = m->w[l];
        nd->b[l] = m->b[l];
        const int o = m->dims[l + 1];
        nd->nt[l] = o <= 8 ? 1 : (o <= 16 ? 2 : (o <= 32 ? 4 : 8));
        // inputs: the previous layer's padded width (its padding columns hold act(0) = 0), or the observation
        nd->ks[l] = l == 0 ? (m->dims[0] + 7) / 8 : nd->nt[l - 1];
        nd->w_off[l] = *arena;
        *arena += nd->ks[l] * nd->nt[l] * 64;
        nd->b_off[l] = *arena;
        *arena += nd->nt[l] * 8;
    }
    return B200ENV_OK;
}

} // namespace

int policy_launch_tc(int64_t n, const b200_mlp *actor, const b200_mlp *critic, const PolicyIO &io, cudaStream_t stream) {
    TcArgs a = {};
    int arena = 0, rc;
    if (actor) {
        if ((rc = fill_tc(actor, &a.actor, &arena))) return rc;
        a.has_actor = 1;
    }
    if (critic) {
        if ((rc = fill_tc(critic, &a.critic, &arena))) return rc;
        a.has_critic = 1;
    }
    a.arena_floats = (arena + 3) / 4 * 4;
    a.io = io;
    const size_t smem = ((size_t)a.arena_floats + (size_t)WARPS * 32 * LD) * sizeof(float);
    if (smem > 227 * 1024) return B200ENV_ESIZE;
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        if (cudaFuncSetAttribute(policy_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return b200_check_launch();
        configured[dev] = smem;
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    const unsigned grid = b200_persistent_grid(n, per_sm, TB);
    policy_forward_tc_kernel<<<grid, TB, smem, stream>>>(a, n);
    return b200_check_launch();
}
