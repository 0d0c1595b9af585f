// FP64 peak micro-benchmark: long independent DFMA chains on every SM.  The roofline denominator
// MEASURED_PEAKS.json lacks (it holds HBM and bf16 only); reported beside the nominal
// 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s.
#include "nb_internal.h"

namespace {
constexpr int CHAINS = 8;
constexpr int ITERS = 4096;
__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = threadIdx.x * 1e-9 + k;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += x[k];
    if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}
// operand-fetch variants: 1 = three distinct register pairs per DFMA (x_k = fma(c_k, d_k, x_k)),
// 2 = one multiplicand shared by consecutive DFMAs (x_k = fma(c, d_k, x_k), the accumulate pattern of the pair term)
template <int VARIANT>
__global__ void __launch_bounds__(256) dfma_regs_kernel(double* out, const double* in) {
    double x[CHAINS], c[CHAINS], d[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) {
        x[k] = in[threadIdx.x + 256 * k];
        c[k] = in[threadIdx.x + 256 * (k + CHAINS)];
        d[k] = in[threadIdx.x + 256 * (k + 2 * CHAINS)];
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) x[k] = fma(VARIANT == 1 ? c[k] : c[0], d[k], x[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += x[k];
    if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" int nb_fp64_peak_variant(int gpu, int variant, double* tflops) {
    if (!tflops || variant < 1 || variant > 2) return NB_ERR_ARG;
    NB_CUDA(cudaSetDevice(gpu));
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, gpu));
    double *out, *in;
    NB_CUDA(cudaMalloc(&out, sizeof(double)));
    NB_CUDA(cudaMalloc(&in, 256 * 3 * CHAINS * sizeof(double)));
    NB_CUDA(cudaMemset(in, 0, 256 * 3 * CHAINS * sizeof(double)));
    cudaEvent_t e0, e1;
    NB_CUDA(cudaEventCreate(&e0));
    NB_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8 * 8;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        NB_CUDA(cudaEventRecord(e0));
        if (variant == 1)
            dfma_regs_kernel<1><<<blocks, 256>>>(out, in);
        else
            dfma_regs_kernel<2><<<blocks, 256>>>(out, in);
        nb::count_launch();
        NB_CUDA(cudaEventRecord(e1));
        NB_CUDA(cudaEventSynchronize(e1));
        float ms;
        NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * CHAINS * (double)ITERS * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(out), cudaFree(in);
    *tflops = best;
    return NB_OK;
}

extern "C" int nb_fp64_peak(int gpu, double* tflops, double* seconds) {
    if (!tflops) return NB_ERR_ARG;
    NB_CUDA(cudaSetDevice(gpu));
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, gpu));
    double* out;
    NB_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    NB_CUDA(cudaEventCreate(&e0));
    NB_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8 * 8;
    double best = 0, best_s = 0;
    for (int rep = 0; rep < 6; rep++) {
        NB_CUDA(cudaEventRecord(e0));
        dfma_kernel<<<blocks, 256>>>(out, 0.999999, 1e-7);
        nb::count_launch();
        NB_CUDA(cudaEventRecord(e1));
        NB_CUDA(cudaEventSynchronize(e1));
        float ms;
        NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * CHAINS * (double)ITERS * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf, best_s = ms * 1e-3;
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(out);
    *tflops = best;
    if (seconds) *seconds = best_s;
    return NB_OK;
}
