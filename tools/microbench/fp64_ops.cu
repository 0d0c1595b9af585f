// FP64 issue-rate microbenchmarks on B200: what does an FP64 instruction cost as a function of its register operands,
// and do SHFL / MUFU.RSQ64H instructions in the same warp take FP64 issue cycles?  (profiles/r02_fp64_ops.md)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8, ITERS = 2048;

__device__ __forceinline__ double rsq(double x) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

template <int V>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int nshfl) {
    double x[CH], c[CH], d[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        x[i] = in[threadIdx.x + 256 * i];
        c[i] = in[threadIdx.x + 256 * (i + CH)];
        d[i] = in[threadIdx.x + 256 * (i + 2 * CH)];
    }
    const int src = (threadIdx.x + 1) & 31;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (V == 0) x[i] = fma(x[i], 1.0000001, 0.5);             // 1 register operand, 2 immediates
            if (V == 1) x[i] = fma(c[i], d[i], x[i]);                 // 3 distinct register pairs
            if (V == 2) x[i] = fma(c[0], d[i], x[i]);                 // shared multiplicand (reuse)
            if (V == 3) x[i] = fma(d[i], d[i], x[i]);                 // 2 distinct (square accumulate)
            if (V == 4) x[i] = x[i] * c[i];                           // DMUL 2 distinct
            if (V == 5) x[i] = x[i] + c[i];                           // DADD 2 distinct
            if (V == 6) x[i] = x[i] * x[i];                           // DMUL 1 distinct
            if (V == 7) x[i] = fma(c[i], d[i], x[i]);                 // + shuffles below
            if (V == 8) x[i] = fma(x[i], 1.0000001, 0.5);             // + shuffles below
            if (V == 9) x[i] = fma(x[i], 1.0000001, 0.5);             // + MUFU below
            if (V == 10) x[i] = fma(c[i], x[i], 0.5);                 // 2 distinct + immediate
        }
        if (V == 7 || V == 8) {
            // nshfl 32-bit shuffles per 8 FP64 instructions on values outside the FP64 chains
            int a = __double2hiint(c[0]), b = __double2loint(c[0]);
            for (int s = 0; s < nshfl; s += 2) {
                a = __shfl_sync(0xffffffffu, a, src);
                b = __shfl_sync(0xffffffffu, b, src);
            }
            c[0] = __hiloint2double(a, b);
        }
        if (V == 9) d[0] = rsq(d[0]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i] + c[i] + d[i];
    if (s == 123.456) out[0] = s;
}

template <int V>
double run(const char* name, int nshfl = 0) {
    double *out, *in;
    cudaMalloc(&out, 8);
    cudaMalloc(&in, 256 * 3 * CH * 8);
    cudaMemset(in, 0, 256 * 3 * CH * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int blocks = 148 * 8 * 4;
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<V><<<blocks, 256>>>(out, in, nshfl);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    // warp-instructions per SM sub-partition per cycle -> cycles per FP64 instruction
    const double inst = (double)CH * ITERS * 256.0 / 32.0 * blocks;  // warp instructions
    const double cyc = best * 1e-3 * 1.965e9 * 148 * 4;             // SMSP cycles available (at 1.965 GHz)
    printf("%-44s %7.3f ms  %.3f cycles per FP64 warp-instruction per SMSP (%.1f%% of 2.0)\n", name, best, cyc / inst, 200.0 * inst / cyc);
    cudaFree(out), cudaFree(in);
    return best;
}

int main() {
    run<0>("DFMA 1 reg + 2 imm");
    run<10>("DFMA 2 distinct regs + imm");
    run<3>("DFMA d,d,x (2 distinct)");
    run<2>("DFMA c0,d,x (shared multiplicand)");
    run<1>("DFMA c,d,x (3 distinct)");
    run<4>("DMUL x,c (2 distinct)");
    run<6>("DMUL x,x (1 distinct)");
    run<5>("DADD x,c (2 distinct)");
    run<8>("DFMA 1 reg + 2 SHFL per 8", 2);
    run<8>("DFMA 1 reg + 4 SHFL per 8", 4);
    run<8>("DFMA 1 reg + 8 SHFL per 8", 8);
    run<7>("DFMA 3 distinct + 2 SHFL per 8", 2);
    run<9>("DFMA 1 reg + 1 MUFU.RSQ64H per 8");
    return 0;
}
