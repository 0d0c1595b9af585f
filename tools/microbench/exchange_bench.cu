// Micro-benchmark of grid-wide exchange schemes for the persistent trajectory kernel (not part of
// the library).  C blocks x 128 threads, one block per SM (cooperative launch).  Every "step" each
// block publishes a 192-byte record and must then obtain all C records in shared memory.
//   scheme 0: per-block flag (16 B apart), every thread p<C polls block p with ld.acquire, then loads its record
//   scheme 1: flags one per 128-B line, relaxed polls + one fence.acq_rel
//   scheme 2: one arrival counter (red.release.add), thread 0 polls, __syncthreads, coalesced load of all records
//   scheme 3: 8 arrival counters on separate lines (block c -> counter c%8), lanes 0..7 of warp 0 poll
//   scheme 4: cooperative_groups grid.sync(), then coalesced load
//   scheme 5: 32-byte records {x,y,z,tag} written/read with 256-bit st/ld, tag checked in the data (no flag hop)
//   scheme 6: like 1 but only warp 0 polls (lane polls 4 producers), then __syncthreads, coalesced load
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o exchange_bench exchange_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int GT = 128, REC = 24;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ double2 ld_strong_d2(const double* p) { double2 v; asm volatile("ld.relaxed.gpu.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_strong_d(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) { asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory"); }
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) { asm volatile("st.relaxed.gpu.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory"); }

template <int SCHEME>
__global__ void __launch_bounds__(GT, 1) bench(double* gbuf, unsigned* flags, int steps, int work, double* sink, long long* cycles) {
    extern __shared__ double smem[];  // [2][C*REC] (scheme 5: C*8*4)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x, C = gridDim.x;
    cg::grid_group grid = cg::this_grid();
    double acc = tid;
    long long t0 = clock64();
    for (int st = 1; st <= steps; st++) {
        // fake compute
        for (int k = 0; k < work; k++) acc = fma(acc, 1.0000001, 1e-9);
        const int par = st & 1;
        double* dst = smem + par * C * REC;
        if (SCHEME == 5) {
            // 8 bodies per block, record = {x,y,z,tag}; lanes 0..1 of each warp publish one body each
            double* g = gbuf + (size_t)par * C * 32;
            if (lane < 2) st256(g + (c * 8 + warp * 2 + lane) * 4, acc, acc + 1, acc + 2, __longlong_as_double((long long)st));
            // each thread fetches bodies tid, tid+128, ... (C*8 bodies) until tagged
            for (int b = tid; b < C * 8; b += GT) {
                double x, y, z, tag;
                do { ld256(g + b * 4, x, y, z, tag); } while (__double_as_longlong(tag) != (long long)st);
                dst[3 * b] = x, dst[3 * b + 1] = y, dst[3 * b + 2] = z;
            }
            __syncthreads();
            acc += dst[(tid * 7) % (C * REC)];
            continue;
        }
        if (SCHEME == 10) {  // flag-in-data, batched polls: thread handles bodies tid + 128k, all loads in flight together
            double* g = gbuf + (size_t)par * C * 32;
            if (lane < 2) st256(g + (c * 8 + warp * 2 + lane) * 4, acc, acc + 1, acc + 2, __longlong_as_double((long long)st));
            const int nb = C * 8;
            double x[8], y[8], z[8], tag[8];
            bool done = false;
            while (!done) {
#pragma unroll
                for (int k = 0; k < 8; k++) { const int b = tid + GT * k; if (b < nb) ld256(g + b * 4, x[k], y[k], z[k], tag[k]); }
                done = true;
#pragma unroll
                for (int k = 0; k < 8; k++) { const int b = tid + GT * k; if (b < nb && __double_as_longlong(tag[k]) != (long long)st) done = false; }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { const int b = tid + GT * k; if (b < nb) { dst[3 * b] = x[k], dst[3 * b + 1] = y[k], dst[3 * b + 2] = z[k]; } }
            __syncthreads();
            acc += dst[(tid * 7) % (C * REC)];
            continue;
        }
        if (SCHEME == 11 || SCHEME == 14) {
            // 11 = PULL: one shared record array of tagged sectors; lane p polls producer p's last sector, then loads its 8
            // 14 = PUSH: every producer writes its 8 tagged sectors into EVERY consumer's private mailbox; consumers poll
            //      only their own mailbox (each sector has one writer and one reader)
            const long long tg = (long long)st;
            if (SCHEME == 11) {
                double* g = gbuf + (size_t)par * C * 32;
                if (tid < 8) st256(g + (c * 8 + tid) * 4, acc, acc + 1, acc + 2, __longlong_as_double(tg));
            } else {
                // mailbox[consumer][parity][producer][8 sectors]; thread t writes consumer t, all 8 sectors of this block
                for (int q = tid; q < C; q += GT) {
                    double* g = gbuf + ((size_t)q * 2 + par) * C * 32 + (size_t)c * 32;
#pragma unroll
                    for (int k = 0; k < 8; k++) st256(g + k * 4, acc, acc + 1, acc + 2, __longlong_as_double(tg));
                }
            }
            if (tid < C) {
                const double* g = SCHEME == 11 ? gbuf + (size_t)par * C * 32 + (size_t)tid * 32
                                               : gbuf + ((size_t)c * 2 + par) * C * 32 + (size_t)tid * 32;
                double x[8], y[8], z[8], tag[8];
                do { asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(tag[7]) : "l"(g + 31) : "memory"); } while (__double_as_longlong(tag[7]) != tg);
#pragma unroll
                for (int k = 0; k < 8; k++) ld256(g + k * 4, x[k], y[k], z[k], tag[k]);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    while (__double_as_longlong(tag[k]) != tg) ld256(g + k * 4, x[k], y[k], z[k], tag[k]);
                    dst[3 * (k * C + tid)] = x[k], dst[3 * (k * C + tid) + 1] = y[k], dst[3 * (k * C + tid) + 2] = z[k];
                }
            }
            __syncthreads();
            acc += dst[(tid * 7) % (C * REC)];
            continue;
        }
        if (SCHEME == 8) {  // data load only, no synchronisation (lower bound of the 24 KB fetch)
            const double* src = gbuf + (size_t)par * C * REC;
            for (int o = 2 * tid; o < C * REC; o += 2 * GT) *reinterpret_cast<double2*>(dst + o) = ld_strong_d2(src + o);
            __syncthreads();
            acc += dst[(tid * 7) % (C * REC)];
            continue;
        }
        if (SCHEME == 9) {  // data load only through one 1-D TMA bulk copy per block
            __shared__ alignas(8) unsigned long long mbar;
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
            if (st == 1 && tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
            __syncthreads();
            if (tid == 0) {
                const unsigned bytes = C * REC * 8;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(gbuf + (size_t)par * C * REC), "r"(bytes), "r"(mb) : "memory");
            }
            const unsigned parity = (st - 1) & 1;
            asm volatile("{\n.reg .pred p;\nW9:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D9;\nbra W9;\nD9:\n}\n" ::"r"(mb), "r"(parity) : "memory");
            acc += dst[(tid * 7) % (C * REC)];
            __syncthreads();
            continue;
        }
        double* rec = gbuf + (size_t)par * C * REC + c * REC;
        if (lane < 6) st_strong_d(rec + warp * 6 + lane, acc + lane);
        if (SCHEME == 7) {  // counter barrier only, no data
            __syncthreads();
            if (tid == 0) {
                red_release_add(flags, 1u);
                while (ld_relaxed(flags) < (unsigned)C * (unsigned)st) {}
                fence_acq_rel();
            }
            __syncthreads();
            continue;
        }
        if (SCHEME == 0 || SCHEME == 1 || SCHEME == 6) {
            const int stride = SCHEME == 0 ? 4 : 32;  // u32 units: 16 B or 128 B apart
            __syncthreads();
            if (tid == 0) st_release(flags + c * stride, (unsigned)st);
            if (SCHEME == 6) {
                if (warp == 0) {
                    for (int p = lane; p < C; p += 32) while (ld_relaxed(flags + p * stride) < (unsigned)st) {}
                    fence_acq_rel();
                }
                __syncthreads();
                const double* src = gbuf + (size_t)par * C * REC;
                for (int o = 2 * tid; o < C * REC; o += 2 * GT) *reinterpret_cast<double2*>(dst + o) = ld_strong_d2(src + o);
            } else if (tid < C) {
                if (SCHEME == 0) { while (ld_acquire(flags + tid * stride) < (unsigned)st) {} }
                else { while (ld_relaxed(flags + tid * stride) < (unsigned)st) {} fence_acq_rel(); }
                const double* src = gbuf + (size_t)par * C * REC + tid * REC;
                double2 r[REC / 2];
#pragma unroll
                for (int k = 0; k < REC / 2; k++) r[k] = ld_strong_d2(src + 2 * k);
#pragma unroll
                for (int k = 0; k < REC / 2; k++) *reinterpret_cast<double2*>(dst + tid * REC + 2 * k) = r[k];
            }
        } else if (SCHEME == 2 || SCHEME == 3) {
            __syncthreads();
            const int NC = SCHEME == 2 ? 1 : 8;
            if (tid == 0) red_release_add(flags + (c % NC) * 32, 1u);
            if (warp == 0) {
                if (lane < NC) {
                    const unsigned members = (C - lane + NC - 1) / NC;  // blocks with c % NC == lane
                    while (ld_relaxed(flags + lane * 32) < members * (unsigned)st) {}
                }
                fence_acq_rel();
            }
            __syncthreads();
            const double* src = gbuf + (size_t)par * C * REC;
            for (int o = 2 * tid; o < C * REC; o += 2 * GT) *reinterpret_cast<double2*>(dst + o) = ld_strong_d2(src + o);
        } else if (SCHEME == 4) {
            __threadfence();
            grid.sync();
            const double* src = gbuf + (size_t)par * C * REC;
            for (int o = 2 * tid; o < C * REC; o += 2 * GT) *reinterpret_cast<double2*>(dst + o) = ld_strong_d2(src + o);
        }
        __syncthreads();
        acc += dst[(tid * 7) % (C * REC)];
    }
    long long t1 = clock64();
    if (tid == 0 && c == 0) cycles[0] = t1 - t0;
    if (acc == 1.2345) sink[0] = acc;
}

// cluster variants: 12 = each block multicasts its 1/CS slice of the record array to the whole cluster (TMA), 13 = cluster.sync only
template <int SCHEME>
__global__ void __launch_bounds__(GT, 1) bench_cluster(double* gbuf, int steps, double* sink, long long* cycles) {
    extern __shared__ double smem[];
    __shared__ alignas(8) unsigned long long mbar;
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x, C = gridDim.x, CS = cl.num_blocks(), r = cl.block_rank();
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    cl.sync();
    double acc = tid;
    long long t0 = clock64();
    const unsigned total = C * REC * 8, slice = total / CS;
    for (int st = 1; st <= steps; st++) {
        const int par = st & 1;
        double* dst = smem + par * C * REC;
        if (SCHEME == 12) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(total) : "memory");
                const unsigned short mask = (unsigned short)((1u << CS) - 1);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"((unsigned)__cvta_generic_to_shared(dst) + r * slice), "l"((const char*)(gbuf + (size_t)par * C * REC) + r * slice),
                               "r"(slice), "r"(mb), "h"(mask) : "memory");
            }
            const unsigned parity = (st - 1) & 1;
            asm volatile("{\n.reg .pred p;\nW%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D%=;\nbra W%=;\nD%=:\n}\n" ::"r"(mb), "r"(parity) : "memory");
            acc += dst[(tid * 7) % (C * REC)];
        }
        cl.sync();
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 1.2345) sink[0] = acc;
}

template <int S>
void run_cluster(int C, int CS, int steps, double* gbuf, double* sink, long long* cyc) {
    size_t smem = 2 * (size_t)C * 32 * sizeof(double);
    cudaFuncSetAttribute(bench_cluster<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench_cluster<S>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C), cfg.blockDim = dim3(GT), cfg.dynamicSmemBytes = smem, cfg.stream = 0;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at, cfg.numAttrs = 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench_cluster<S>, gbuf, steps, sink, cyc);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("cluster scheme %d C=%3d CS=%2d: %.3f us/step (%lld clk/step) %s %s\n", S, C, CS, ms * 1e3 / steps, h / steps,
           e == cudaSuccess ? "" : cudaGetErrorString(e), e2 == cudaSuccess ? "" : cudaGetErrorString(e2));
    fflush(stdout);
}

template <int S>
void run(int C, int steps, int work, double* gbuf, unsigned* flags, double* sink, long long* cyc) {
    size_t smem = 2 * (size_t)C * 32 * sizeof(double);
    cudaFuncSetAttribute(bench<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(flags, 0, 1 << 20);
    cudaMemset(gbuf, 0, 1 << 24);
    void* args[] = {&gbuf, &flags, &steps, &work, &sink, &cyc};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bench<S>, dim3(C), dim3(GT), args, smem, 0);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("scheme %d C=%3d work=%4d: %.3f us/step (%lld clk/step) %s %s\n", S, C, work, ms * 1e3 / steps, h / steps,
           e == cudaSuccess ? "" : cudaGetErrorString(e), e2 == cudaSuccess ? "" : cudaGetErrorString(e2));
    fflush(stdout);
}

int main(int argc, char** argv) {
    int steps = argc > 1 ? atoi(argv[1]) : 20000;
    double *gbuf, *sink;
    unsigned* flags;
    long long* cyc;
    cudaMalloc(&gbuf, 1 << 24), cudaMalloc(&sink, 8), cudaMalloc(&flags, 1 << 20), cudaMalloc(&cyc, 8);
    for (int work : {0}) {
        for (int C : {2, 32, 64, 128}) {
            run<0>(C, steps, work, gbuf, flags, sink, cyc);
            run<1>(C, steps, work, gbuf, flags, sink, cyc);
            run<6>(C, steps, work, gbuf, flags, sink, cyc);
            run<2>(C, steps, work, gbuf, flags, sink, cyc);
            run<3>(C, steps, work, gbuf, flags, sink, cyc);
            run<4>(C, steps, work, gbuf, flags, sink, cyc);
            run<5>(C, steps, work, gbuf, flags, sink, cyc);
            run<10>(C, steps, work, gbuf, flags, sink, cyc);
            run<11>(C, steps, work, gbuf, flags, sink, cyc);
            run<14>(C, steps, work, gbuf, flags, sink, cyc);
            run<7>(C, steps, work, gbuf, flags, sink, cyc);
            run<8>(C, steps, work, gbuf, flags, sink, cyc);
            run<9>(C, steps, work, gbuf, flags, sink, cyc);
        }
    }
    for (int CS : {2, 4, 8, 16}) {
        run_cluster<13>(128, CS, steps, gbuf, sink, cyc);
        run_cluster<12>(128, CS, steps, gbuf, sink, cyc);
    }
    return 0;
}
