// What does a non-FP64 instruction cost in FP64 issue time on B200?  8 independent DFMA chains per thread plus K
// extra instructions of one kind per 8 DFMAs (independent of the chains and of each other).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8, ITERS = 2048;
enum { NONE, SHFL_IDX, SHFL_BFLY, LDS64, LDS128, LDS128_BCAST, STS64, LDS_ADD_STS, IADD, FFMA, MOV64 };

template <int KIND, int K>
__global__ void __launch_bounds__(256) k(double* out, const double* in) {
    __shared__ double sm[256 * 4];
    double x[CH];
    int s[8];
    float f[8];
    double acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = in[threadIdx.x + 256 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = threadIdx.x * 7 + i, f[i] = threadIdx.x + i;
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i;
    __syncthreads();
    const int src = (threadIdx.x + 1) & 31;
    const int lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 128;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) x[i] = fma(x[i], 1.0000001, 0.5);
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (KIND == SHFL_IDX) s[j] = __shfl_sync(0xffffffffu, s[j], src);
            if (KIND == SHFL_BFLY) s[j] = __shfl_xor_sync(0xffffffffu, s[j], 1);
            if (KIND == LDS64) acc[j & 3] += sm[wbase + ((lane + it + j) & 31)];  // includes a DADD: see LDS_ADD_STS baseline
            if (KIND == LDS128) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[wbase + 2 * ((lane + it + j) & 31)]);
                s[j] ^= __double2loint(v.x) ^ __double2loint(v.y);
            }
            if (KIND == LDS128_BCAST) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[wbase + 2 * ((it + j) & 31)]);
                s[j] ^= __double2loint(v.x) ^ __double2loint(v.y);
            }
            if (KIND == STS64) sm[wbase + ((lane + j) & 31) + 32 * (j & 3)] = x[j];
            if (KIND == LDS_ADD_STS) {
                double* p = &sm[wbase + ((lane + it + j) & 31) + 32 * (j & 3)];
                *p = *p + x[j];
            }
            if (KIND == IADD) s[j] = s[j] * 3 + it;
            if (KIND == FFMA) f[j] = fmaf(f[j], 1.0001f, 0.5f);
            if (KIND == MOV64) { double t = x[j]; x[j] = x[(j + 1) & 7]; x[(j + 1) & 7] = t; }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) r += x[i];
#pragma unroll
    for (int i = 0; i < 8; i++) r += s[i] + f[i];
    r += acc[0] + acc[1] + acc[2] + acc[3] + sm[threadIdx.x];
    if (r == 123.456) out[0] = r;
}

template <int KIND, int K>
void run(const char* name, double base) {
    double *out, *in;
    cudaMalloc(&out, 8);
    cudaMalloc(&in, 256 * CH * 8);
    cudaMemset(in, 0, 256 * CH * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int blocks = 148 * 4 * 4;
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<KIND, K><<<blocks, 256>>>(out, in);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double groups = (double)ITERS * 256.0 / 32.0 * blocks;  // groups of 8 DFMAs (+K extras), warp level
    const double cyc = best * 1e-3 * 1.965e9 * 148 * 4 / groups;  // SMSP cycles per group
    printf("%-40s K=%d  %7.3f ms  %6.2f cycles per group of 8 DFMA  -> %5.2f cycles per extra instruction\n", name, K, best, cyc,
           K ? (cyc - base) / K : 0.0);
    cudaFree(out), cudaFree(in);
}

int main() {
    const double base = 16.3;
    run<NONE, 0>("8 DFMA alone", base);
    run<SHFL_IDX, 1>("SHFL.IDX", base);
    run<SHFL_IDX, 2>("SHFL.IDX", base);
    run<SHFL_IDX, 4>("SHFL.IDX", base);
    run<SHFL_IDX, 8>("SHFL.IDX", base);
    run<SHFL_BFLY, 4>("SHFL.BFLY", base);
    run<LDS64, 2>("LDS.64 + DADD", base);
    run<LDS64, 4>("LDS.64 + DADD", base);
    run<LDS128, 2>("LDS.128 (distinct addresses)", base);
    run<LDS128, 4>("LDS.128 (distinct addresses)", base);
    run<LDS128_BCAST, 4>("LDS.128 (broadcast)", base);
    run<STS64, 2>("STS.64", base);
    run<STS64, 4>("STS.64", base);
    run<LDS_ADD_STS, 2>("LDS.64 + DADD + STS.64", base);
    run<LDS_ADD_STS, 4>("LDS.64 + DADD + STS.64", base);
    run<IADD, 8>("IMAD", base);
    run<FFMA, 8>("FFMA", base);
    run<MOV64, 4>("64-bit register swap", base);
    return 0;
}
