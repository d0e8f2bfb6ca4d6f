// peak.cu -- roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 and FP32 vector-pipe peaks.
// Independent FMA chains per thread (8 accumulators), enough resident warps to saturate the pipe; timed with CUDA
// events on the given stream.  Diagnostic only -- not on the hot path.
#include "common.cuh"

namespace {
template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3;
    T x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6, x7 = x0 + (T)7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <typename T>
int measure(double *tflops, int iters, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    T *buf = nullptr;
    if (cudaMalloc(&buf, (size_t)blocks * threads * sizeof(T)) != cudaSuccess) return B200ENV_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_chain_kernel<T><<<blocks, threads, 0, s>>>(buf, iters, (T)0.999999, (T)1e-6); // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        fma_chain_kernel<T><<<blocks, threads, 0, s>>>(buf, iters, (T)0.999999, (T)1e-6);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    int rc = b200_check_launch();
    if (rc) return rc;
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return B200ENV_OK;
}
} // namespace

extern "C" B200_API int b200_measure_fma_peak(int dtype, int iters, double *tflops, void *cuda_stream) {
    if (!tflops || iters <= 0) return B200ENV_ENULL;
    return dtype == B200ENV_F64 ? measure<double>(tflops, iters, (cudaStream_t)cuda_stream)
                                : measure<float>(tflops, iters, (cudaStream_t)cuda_stream);
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise evaluation of the fastmath64.cuh functions (accuracy tests only)
namespace {
__global__ void fastmath_eval_kernel(int func, int64_t n, const double *x, double *o0, double *o1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double a = 0, b = 0;
    switch (func) {
    case 0: fm64::sincos(v, &a, &b); break;
    case 1: a = fm64::exp(v); break;
    case 2: a = fm64::log(v); break;
    case 3: a = fm64::tanh(v); break;
    case 4: a = pow_from_log<double>(fm64::log(v), o1[i]); b = o1[i]; break; // exponent passed in o1
    case 5: a = fm64::asin(v); break;
    default: break;
    }
    o0[i] = a;
    if (o1) o1[i] = b;
}
} // namespace

extern "C" B200_API int b200_fastmath_eval(int func, int64_t n, const double *x, double *out0, double *out1,
                                           void *cuda_stream) {
    if (!x || !out0 || n <= 0) return B200ENV_ENULL;
    fastmath_eval_kernel<<<b200_grid(n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(func, n, x, out0, out1);
    return b200_check_launch();
}
