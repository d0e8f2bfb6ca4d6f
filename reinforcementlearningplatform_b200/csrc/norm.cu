// norm.cu -- K-NORM: running mean/std normalisation of observations and rewards (sm_100a), SURVEY 8(f)-1.
//
// Replaces `RunningMeanStd.update` + `Normalization.__call__` of utils/classes.py:626-656, which the train loops apply
// to every reward (`reward_norm(env.reward)`, demonstration/PPO2/PPO2-4-CartPoleAngleOnly/train.py:139,210) and the UAV
// envs to every observation (`env.current_state_norm(env.current_state, update=True)`,
// PPO2-4-UavFntsmcParamPos/train.py:291,308; environment/UavFntsmcParam/uav_pos_ctrl_RL.py:36-37):
//     n += 1
//     n == 1:  mean = x; std = x                                   (sic: the first sample's std is the sample)
//     else:    mean = old + (x - old) / n;  S += (x - old) * (x - mean);  std = sqrt(S / n)
//     y = (x - mean) / (std + 1e-8)
//
// The running state is double run[3][dim] = (n, mean, S); std is derived: n == 1 ? mean : sqrt(S / n), which is what
// the reference holds in every reachable state (n == 0: mean = S = std = 0).
//
// Two entry points:
//  * b200_norm_seq   -- the reference's sample-by-sample recurrence over `rows` samples, one thread per feature,
//                       bit-exact (IEEE add/sub/mul/div/sqrt, no FMA contraction); for single-instance rollouts.
//  * b200_norm_batch_stats + b200_norm_merge_apply -- one step of N instances: the N samples of a step enter the
//    statistics TOGETHER (Chan/Golub/LeVeque pairwise merge of (count, mean, M2)) and all of them are normalised with
//    the merged statistics.  This is a different, explicitly defined semantics (the reference has one instance and
//    never sees a batch); it is chosen so that a batch of ONE sample reproduces the reference recurrence bit for bit:
//        delta = mb - ma;  n = na + nb;  mean = ma + delta * nb / n;  M2 = (Ma + Mb) + delta * (mb - mean) * nb
//    (nb = 1, Mb = 0, mb = x gives exactly the lines above).  Batch statistics of different GPUs are merged in rank
//    order by the same formula (n_batches > 1), so every rank holds identical running statistics.
//
// HBM traffic per element of a [dim][N] batch: one read (statistics) + one read + one write (apply) in the I/O dtype;
// the second read hits the 126 MB L2 when dim * N * sizeof < L2.  Statistics are accumulated in fp64 about a pivot
// (the running mean, or the first sample when there is none) so that the sum of squares does not cancel.
#include "common.cuh"

namespace {

constexpr int NORM_BLOCK = 256;
constexpr int NORM_MAX_PART = 1024; // partial sums per feature kept in scratch

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum of two values (fixed tree for a fixed block size); result valid in thread 0
__device__ __forceinline__ void block_sum2(double &a, double &b) {
    __shared__ double sa[NORM_BLOCK / 32], sb[NORM_BLOCK / 32];
    a = wsum(a);
    b = wsum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads(); // protect sa/sb against a previous use
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < NORM_BLOCK / 32 ? sa[lane] : 0.0;
        b = lane < NORM_BLOCK / 32 ? sb[lane] : 0.0;
        a = wsum(a);
        b = wsum(b);
    }
}

// scratch layout (doubles): part[dim][NORM_MAX_PART][2] then one u32 ticket per feature (as doubles' storage)
template <typename TX>
__global__ void __launch_bounds__(NORM_BLOCK)
norm_batch_stats_kernel(int64_t n, int dim, const TX *__restrict__ x, const double *__restrict__ run,
                        double *__restrict__ batch, double *__restrict__ part, unsigned int *__restrict__ ticket) {
    const int f = blockIdx.y, nb = gridDim.x;
    const TX *xf = x + (int64_t)f * n;
    const double pivot = (run && run[f] > 0.0) ? run[dim + f] : (double)xf[0];
    double s1 = 0.0, s2 = 0.0;
    const int64_t stride = (int64_t)nb * NORM_BLOCK;
    int64_t i = (int64_t)blockIdx.x * NORM_BLOCK + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) { // 4 loads in flight per thread
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (double)__ldg(xf + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double d = v[u] - pivot;
            s1 += d;
            s2 = fma(d, d, s2);
        }
    }
    for (; i < n; i += stride) {
        const double d = (double)__ldg(xf + i) - pivot;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    block_sum2(s1, s2);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        double *p = part + ((int64_t)f * NORM_MAX_PART + blockIdx.x) * 2;
        p[0] = s1;
        p[1] = s2;
        __threadfence();
        last = atomicAdd(ticket + f, 1u) == (unsigned)(nb - 1);
    }
    __syncthreads();
    if (!last) return;
    // the last block of this feature adds the partials in index order (fixed tree) and publishes the batch statistics
    __threadfence();
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nb; k += NORM_BLOCK) {
        const volatile double *p = part + ((int64_t)f * NORM_MAX_PART + k) * 2;
        a += p[0];
        b += p[1];
    }
    block_sum2(a, b);
    if (threadIdx.x == 0) {
        const double cnt = (double)n;
        const double dm = a / cnt;
        batch[f] = cnt;
        // a one-sample batch is the sample itself (pivot + (x - pivot) would round): keeps N = 1 on the reference's bits
        batch[dim + f] = n == 1 ? (double)xf[0] : pivot + dm;
        batch[2 * dim + f] = n == 1 ? 0.0 : fmax(b - a * dm, 0.0);
        ticket[f] = 0u; // ready for the next launch
    }
}

struct Stat { double n, mean, S; };

// Chan merge in the form that degenerates to the reference's Welford update for a one-sample batch
__device__ __forceinline__ Stat merge(Stat a, double nb, double mb, double Mb) {
    if (nb <= 0.0) return a;
    Stat o;
    o.n = __dadd_rn(a.n, nb);
    if (a.n <= 0.0) { // first batch: `self.mean = x` (RunningMeanStd.update, n == 1 branch)
        o.mean = mb;
        o.S = Mb;
        return o;
    }
    const double delta = __dsub_rn(mb, a.mean);
    o.mean = __dadd_rn(a.mean, __ddiv_rn(__dmul_rn(delta, nb), o.n));
    o.S = __dadd_rn(__dadd_rn(a.S, Mb), __dmul_rn(__dmul_rn(delta, __dsub_rn(mb, o.mean)), nb));
    return o;
}

__device__ __forceinline__ double std_of(Stat s) {
    // n == 1: `self.std = x` (= mean); otherwise sqrt(S / n); n == 0: sqrt(0)
    if (s.n == 1.0) return s.mean;
    return s.n > 0.0 ? __dsqrt_rn(__ddiv_rn(s.S, s.n)) : 0.0;
}

template <typename TX>
__global__ void __launch_bounds__(NORM_BLOCK)
norm_merge_apply_kernel(int64_t n, int dim, const TX *x, TX *y, // no __restrict__: y == x is allowed (in-place
                        // normalisation of a rollout row); every element is read and then written by the same thread
                        const double *__restrict__ batch, int n_batches, const double *__restrict__ run_in,
                        double *__restrict__ run_out, int update, double eps) {
    const int f = blockIdx.y;
    Stat s{run_in[f], run_in[dim + f], run_in[2 * dim + f]};
    if (update)
        for (int b = 0; b < n_batches; ++b) {
            const double *bs = batch + (int64_t)b * 3 * dim;
            s = merge(s, bs[f], bs[dim + f], bs[2 * dim + f]);
        }
    if (update && run_out && blockIdx.x == 0 && threadIdx.x == 0) {
        run_out[f] = s.n;
        run_out[dim + f] = s.mean;
        run_out[2 * dim + f] = s.S;
    }
    if (!y) return;
    const double mean = s.mean, den = __dadd_rn(std_of(s), eps);
    const TX *xf = x + (int64_t)f * n;
    TX *yf = y + (int64_t)f * n;
    // float32 I/O: the quotient is rounded to float32 on store, so the reciprocal + one FMA-corrected Newton step of
    // common.cuh (<= 1 ulp of the float64 quotient in rare half-way cases) is invisible; float64 I/O keeps the IEEE
    // division (bit-exact against the reference for one-sample batches).  4 elements per trip: loads issued together.
    const Divisor<double> dv(den);
    const int64_t stride = (int64_t)gridDim.x * NORM_BLOCK;
    int64_t i = (int64_t)blockIdx.x * NORM_BLOCK + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (double)__ldcs(xf + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double d = __dsub_rn(v[u], mean);
            __stcs(yf + i + u * stride, (TX)(sizeof(TX) == 4 ? dv.div(d) : __ddiv_rn(d, den)));
        }
    }
    for (; i < n; i += stride) {
        const double d = __dsub_rn((double)xf[i], mean);
        yf[i] = (TX)(sizeof(TX) == 4 ? dv.div(d) : __ddiv_rn(d, den));
    }
}

// the reference recurrence, sample by sample; x, y are [dim][rows]
template <typename TX>
__global__ void norm_seq_kernel(int64_t rows, int dim, const TX *__restrict__ x, TX *__restrict__ y,
                                double *__restrict__ run, int update, double eps) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= dim) return;
    Stat s{run[f], run[dim + f], run[2 * dim + f]};
    const TX *xf = x + (int64_t)f * rows;
    for (int64_t t = 0; t < rows; ++t) {
        const double v = (double)xf[t];
        if (update) s = merge(s, 1.0, v, 0.0);
        if (y) y[(int64_t)f * rows + t] = (TX)__ddiv_rn(__dsub_rn(v, s.mean), __dadd_rn(std_of(s), eps));
    }
    if (update) {
        run[f] = s.n;
        run[dim + f] = s.mean;
        run[2 * dim + f] = s.S;
    }
}

unsigned stat_blocks(int64_t n, int dim) {
    int64_t want = (n + (int64_t)NORM_BLOCK * 8 - 1) / ((int64_t)NORM_BLOCK * 8);
    int64_t cap = (int64_t)b200_persistent_grid((int64_t)1 << 40, 8, NORM_BLOCK) / (dim > 0 ? dim : 1);
    if (cap < 1) cap = 1;
    if (cap > NORM_MAX_PART) cap = NORM_MAX_PART;
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}

} // namespace

extern "C" B200_API size_t b200_norm_scratch_bytes(int dim) {
    return (size_t)dim * NORM_MAX_PART * 2 * sizeof(double) + (size_t)dim * sizeof(unsigned int);
}

extern "C" B200_API int b200_norm_batch_stats(int dtype, int64_t n, int dim, const void *x, const double *run,
                                              double *batch_stats, void *scratch, void *cuda_stream) {
    if (n <= 0 || dim <= 0 || dim > 65535) return B200ENV_ESIZE;
    if (dtype != B200ENV_F64 && dtype != B200ENV_F32) return B200ENV_EDTYPE;
    if (!x || !batch_stats || !scratch) return B200ENV_ENULL;
    double *part = (double *)scratch;
    unsigned int *ticket = (unsigned int *)(part + (size_t)dim * NORM_MAX_PART * 2);
    const dim3 grid(stat_blocks(n, dim), (unsigned)dim);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (dtype == B200ENV_F64)
        norm_batch_stats_kernel<double><<<grid, NORM_BLOCK, 0, s>>>(n, dim, (const double *)x, run, batch_stats, part, ticket);
    else
        norm_batch_stats_kernel<float><<<grid, NORM_BLOCK, 0, s>>>(n, dim, (const float *)x, run, batch_stats, part, ticket);
    return b200_check_launch();
}

extern "C" B200_API int b200_norm_merge_apply(int dtype, int64_t n, int dim, const void *x, void *y,
                                              const double *batch_stats, int n_batches, const double *run_in,
                                              double *run_out, int update, double eps, void *cuda_stream) {
    if (n <= 0 || dim <= 0 || dim > 65535 || n_batches < 0) return B200ENV_ESIZE;
    if (dtype != B200ENV_F64 && dtype != B200ENV_F32) return B200ENV_EDTYPE;
    if (!run_in || (y && !x) || (update && n_batches > 0 && !batch_stats)) return B200ENV_ENULL;
    const dim3 grid(y ? stat_blocks(n, dim) : 1u, (unsigned)dim);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (dtype == B200ENV_F64)
        norm_merge_apply_kernel<double><<<grid, NORM_BLOCK, 0, s>>>(n, dim, (const double *)x, (double *)y, batch_stats,
                                                                    n_batches, run_in, run_out, update, eps);
    else
        norm_merge_apply_kernel<float><<<grid, NORM_BLOCK, 0, s>>>(n, dim, (const float *)x, (float *)y, batch_stats,
                                                                   n_batches, run_in, run_out, update, eps);
    return b200_check_launch();
}

// Rows of a time-major [rows][n] buffer of ONE scalar feature (the rewards of a rollout) entering the running statistics
// row after row: cum[3][rows] = the running (n, mean, S) after rows 0..t have been merged (each row = n_batches batch
// statistics, e.g. one per rank, in index order), run_out = the state after the last row.  One thread: `rows` dependent
// merges of a few fp64 operations each.
__global__ void norm_rows_prefix_kernel(int rows, int n_batches, const double *__restrict__ batch,
                                        const double *__restrict__ run_in, double *__restrict__ cum,
                                        double *__restrict__ run_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Stat s{run_in[0], run_in[1], run_in[2]};
    for (int t = 0; t < rows; ++t) {
        for (int b = 0; b < n_batches; ++b) {
            const double *bs = batch + (int64_t)b * 3 * rows;
            s = merge(s, bs[t], bs[rows + t], bs[2 * rows + t]);
        }
        cum[t] = s.n;
        cum[rows + t] = s.mean;
        cum[2 * rows + t] = s.S;
    }
    run_out[0] = s.n;
    run_out[1] = s.mean;
    run_out[2] = s.S;
}

extern "C" B200_API int b200_norm_rows_prefix(int rows, const double *batch_stats, int n_batches, const double *run_in,
                                              double *cum, double *run_out, void *cuda_stream) {
    if (rows <= 0 || rows > 65535 || n_batches <= 0) return B200ENV_ESIZE;
    if (!batch_stats || !run_in || !cum || !run_out) return B200ENV_ENULL;
    norm_rows_prefix_kernel<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(rows, n_batches, batch_stats, run_in, cum, run_out);
    return b200_check_launch();
}

extern "C" B200_API int b200_norm_seq(int dtype, int64_t rows, int dim, const void *x, void *y, double *run, int update,
                                      double eps, void *cuda_stream) {
    if (rows <= 0 || dim <= 0) return B200ENV_ESIZE;
    if (dtype != B200ENV_F64 && dtype != B200ENV_F32) return B200ENV_EDTYPE;
    if (!x || !run) return B200ENV_ENULL;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const unsigned grid = (unsigned)((dim + 31) / 32);
    if (dtype == B200ENV_F64)
        norm_seq_kernel<double><<<grid, 32, 0, s>>>(rows, dim, (const double *)x, (double *)y, run, update, eps);
    else
        norm_seq_kernel<float><<<grid, 32, 0, s>>>(rows, dim, (const float *)x, (float *)y, run, update, eps);
    return b200_check_launch();
}
