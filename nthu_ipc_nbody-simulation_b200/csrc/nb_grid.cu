// Whole-GPU cooperative persistent trajectory kernel: ONE system (or up to 4 systems of the same n
// in lock step) spread over up to 128 SMs, all steps and observers in one launch.  This is the
// low-step-latency path for the b512 / b1024 queries: a single block needs ~133 us per b1024 step
// (nb_traj.cu), the FP64 floor over 128 SMs is ~1.05 us.
//
// Replaces the host-driven per-step launch sequence of the reference (hw5.cu:368-404, 387-403,
// 489-508: 3-4 kernel launches per step, 600 000-800 000 per trajectory).  Arithmetic: nbody.cc:51-89.
//
// Decomposition: block c owns bodies [8c, 8c+8); warp w of the block owns bodies 8c+2w, 8c+2w+1 and
// its 32 lanes split the j range (j = lane, lane+32, ...), two i-bodies per j record so that the
// shared-memory traffic stays at half the FP64 issue rate.  A 5-round xor butterfly combines the
// lanes; lanes 0-5 each integrate one (body, component) and publish the new coordinate.
//
// Exchange (the step's only grid-wide dependency), measured design (tools/microbench/exchange_bench.cu,
// profiles/r01_exchange_microbench.md): on B200 a release/acquire hop costs ~0.4-0.5 us per fence
// (MEMBAR.ALL.GPU) and an atomic-counter grid barrier ~1.5 us, while an un-fenced store -> poll hop is
// ~0.37 us.  So there are no fences, no atomics and no barrier object: every body is published as one
// naturally aligned 32-byte sector {x, y, z, step tag} with a single 256-bit store, into a global
// record array double-buffered by step parity (L2 resident).  A sector is the unit the L2 reads and
// writes, so a reader sees it entirely old or entirely new and the tag validates the data it travels
// with; no ordering between different sectors is assumed anywhere.  Thread p of every block polls the
// tag of producer p's last sector (the block's 8 sectors leave in one store instruction), then the
// block fetches all sectors coalesced with 256-bit loads, re-reading any sector whose tag is still
// old, and scatters x,y,z into shared memory (24 B stride: bank-conflict free for 8 B accesses).
// Reuse of a parity buffer is safe without fences: a block overwrites step s-2's sectors only after
// it has consumed every block's step s-1 sectors, which each block publishes only after its own reads
// of step s-2 have completed (data dependence).  Spins are bounded (clock64) and raise `status`.
//
// Co-residency of all blocks is guaranteed by the cooperative launch (grid <= SM count, 1 block/SM).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "nb_internal.h"
#include "nb_math.cuh"

namespace nb {
namespace {

constexpr int GB = 8;         // bodies per block
constexpr int GT = 128;       // threads per block (4 warps x 2 bodies)
constexpr int REC = GB * 3;   // doubles per block record
constexpr int MAX_T = 4;      // trajectories per launch (shared memory: 56 B x npad each)
constexpr long long SPIN_LIMIT = 4000000000LL;  // ~2 s of SM clocks

__device__ __forceinline__ double ld_strong_d(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// one naturally aligned 32-byte sector per access (SASS LDG/STG.E.ENL2.256): the unit the L2 reads and
// writes, so a sector is observed either entirely old or entirely new
__device__ __forceinline__ void ld_sector(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_sector(double* p, double a, double b, double c, double d) {
    asm volatile("st.relaxed.gpu.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

struct TState {  // per trajectory, per thread (registers after unrolling)
    // uniform
    int n_dev, kind, P, A, DD, step, step_end;
    bool q3_armed, active;
    double min_d2, cost;
    int argmin_step, hit_step, destroyed_step;
    // device bookkeeping (thread k < n_dev)
    int my_dev, my_reach;
    double my_m0;
    // integrator lanes (lane < 6)
    double v, q;
    double fst_next;  // |sin| of the next step, prefetched (the acquire polls invalidate L1 every step)
    int cur;
};

// PROFILE: block 0 thread 0 accumulates clock64 per phase into prof[0..7] (NB_GRID_PROFILE=1, T == 1 only)
template <int MATH, int T, bool PROFILE>
__global__ void __launch_bounds__(GT, 1)
grid_traj_kernel(const TrajDesc* __restrict__ descs, const double* __restrict__ fst, double* __restrict__ gbuf,
                 long long* __restrict__ prof, int* __restrict__ status, int npad) {
    extern __shared__ double smem[];
    __shared__ int s_abort;
    __shared__ double s_stage[REC];  // the block's 8 new positions, staged for the one-instruction publish
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, C = gridDim.x;
    const int n = descs[0].n;
    const int body0 = c * GB + 2 * warp;        // the warp's two bodies: body0, body0 + 1
    const int my_body = body0 + lane / 3;       // integrator lanes 0..5
    const int my_comp = lane % 3;
    const bool integ = lane < 6 && my_body < n;

    // per trajectory: pos[2][3*npad] (body-major x,y,z) then gm[npad]
    auto s_pos = [&](int t, int buf) { return smem + (size_t)t * 7 * npad + (size_t)buf * 3 * npad; };
    auto s_gm = [&](int t) { return smem + (size_t)t * 7 * npad + 6 * npad; };

    TState ts[T];
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;
    auto tick = [&](int phase) {
        if (PROFILE) {
            const long long now = clock64();
            pacc[phase] += now - pt;
            pt = now;
        }
    };
    if (tid == 0) s_abort = 0;
#pragma unroll
    for (int t = 0; t < T; t++) {
        const TrajDesc& d = descs[t];
        TState& s = ts[t];
        s.n_dev = d.n_dev, s.kind = d.kind, s.P = d.planet, s.A = d.asteroid, s.DD = d.destroy_device;
        s.step = d.step_begin, s.step_end = d.step_end;
        s.min_d2 = d.ev->min_d2, s.argmin_step = d.ev->argmin_step, s.hit_step = d.ev->hit_step;
        s.destroyed_step = d.ev->destroyed_step, s.cost = d.ev->cost;
        s.q3_armed = (s.kind == NB_KIND_Q3) && s.DD >= 0 && s.DD < n && d.m[s.DD] != 0.0;
        s.active = !((s.kind >= NB_KIND_Q2) && s.hit_step != -2);
        s.cur = 0;
        s.my_dev = -1, s.my_reach = -2, s.my_m0 = 0.0;
        if (tid < s.n_dev) {
            s.my_dev = d.dev_index[tid];
            s.my_m0 = d.m[s.my_dev];
            s.my_reach = d.ev->reach_step[tid];
        }
        s.fst_next = fst[s.step + 1];
        s.v = s.q = 0.0;
        if (integ) {
            s.v = d.v[my_comp * n + my_body];
            s.q = d.q[my_comp * n + my_body];
        }
        for (int i = tid; i < npad; i += GT) {
            double x = 0, y = 0, z = 0, g = 0;
            if (i < n) {
                x = d.q[i], y = d.q[i + n], z = d.q[i + 2 * n];
                g = d.is_device[i] ? 0.0 : gm_eff(d.m[i], false, 0.0);
            }
#pragma unroll
            for (int b = 0; b < 2; b++) {
                double* p = s_pos(t, b) + 3 * i;
                p[0] = x, p[1] = y, p[2] = z;
            }
            s_gm(t)[i] = g;
        }
    }
    __syncthreads();

    auto observe = [&](TState& s, int t) {
        const double* pos = s_pos(t, s.cur);
        const double* p = pos + 3 * s.P;
        const double* a = pos + 3 * s.A;
        const double px = p[0], py = p[1], pz = p[2];
        const double d2 = dist2_rn(px, py, pz, a[0], a[1], a[2]);
        if (d2 < s.min_d2) {  // hw5.cu:245-247
            s.min_d2 = d2;
            s.argmin_step = s.step;
        }
        if (s.kind == NB_KIND_Q2 && s.my_dev >= 0 && s.my_reach == -2) {  // hw5.cu:265-287
            const double* dv = pos + 3 * s.my_dev;
            const double md = __dmul_rn(MISSILE_STEP, (double)s.step);
            if (dist2_rn(px, py, pz, dv[0], dv[1], dv[2]) < __dmul_rn(md, md)) s.my_reach = s.step;
        }
        if (s.kind >= NB_KIND_Q2) {
            if (d2 < PLANET_RADIUS2) {  // nbody.cc:134, hw5.cu:295-298
                s.hit_step = s.step;
                s.active = false;
            } else if (s.q3_armed && s.destroyed_step == -2) {  // hw5.cu:299-307
                const double* dv = pos + 3 * s.DD;
                const double md = __dmul_rn(MISSILE_STEP, (double)s.step);
                if (dist2_rn(px, py, pz, dv[0], dv[1], dv[2]) < __dmul_rn(md, md)) {
                    s.destroyed_step = s.step;
                    s.cost = __dadd_rn(1e5, __dmul_rn(1e3, __dmul_rn((double)(s.step + 1), DT)));
                }
            }
        }
        if (s.step >= s.step_end) s.active = false;
    };

#pragma unroll
    for (int t = 0; t < T; t++) {
        if (descs[t].ev->steps_done < ts[t].step && ts[t].active) observe(ts[t], t);
        if (ts[t].step >= ts[t].step_end) ts[t].active = false;
    }

    bool any = false;
#pragma unroll
    for (int t = 0; t < T; t++) any |= ts[t].active;
    long long g0 = 0, c0 = 0;
    if (PROFILE) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        c0 = clock64();
    }

    while (any) {
#pragma unroll
        for (int t = 0; t < T; t++) {
            TState& s = ts[t];
            if (!s.active) continue;  // uniform across the grid
            const int st = ++s.step;
            if (PROFILE) pt = clock64();
            // (1) G*m_eff of the devices for this step (nbody.cc:61-64); a destroyed device has mass 0
            if (s.my_dev >= 0) {
                const bool gone = (s.kind == NB_KIND_Q3) && s.my_dev == s.DD && s.destroyed_step != -2;
                s_gm(t)[s.my_dev] = gm_eff(gone ? 0.0 : s.my_m0, true, s.fst_next);
            }
            __syncthreads();
            tick(0);
            s.fst_next = fst[st + 1];
            // (2) forces on the warp's two bodies, j split over the lanes (nbody.cc:56-74)
            const double* cpos = s_pos(t, s.cur);
            const double* cg = s_gm(t);
            const double* pa = cpos + 3 * min(body0, npad - 1);
            const double* pb = cpos + 3 * min(body0 + 1, npad - 1);
            const double xa = pa[0], ya = pa[1], za = pa[2], xb = pb[0], yb = pb[1], zb = pb[2];
            double ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
#pragma unroll 4
            for (int j = lane; j < npad; j += 32) {
                const double jx = cpos[3 * j], jy = cpos[3 * j + 1], jz = cpos[3 * j + 2], jg = cg[j];
                pair<MATH>(xa, ya, za, jx, jy, jz, jg, ax0, ay0, az0);
                pair<MATH>(xb, yb, zb, jx, jy, jz, jg, ax1, ay1, az1);
            }
            tick(1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax0 += __shfl_xor_sync(0xffffffffu, ax0, o);
                ay0 += __shfl_xor_sync(0xffffffffu, ay0, o);
                az0 += __shfl_xor_sync(0xffffffffu, az0, o);
                ax1 += __shfl_xor_sync(0xffffffffu, ax1, o);
                ay1 += __shfl_xor_sync(0xffffffffu, ay1, o);
                az1 += __shfl_xor_sync(0xffffffffu, az1, o);
            }
            // (3) v += a*dt; q += v*dt (nbody.cc:77-88): lane l < 6 owns (body0 + l/3, component l%3)
            if (lane < 6) {
                const double a = lane == 0 ? ax0 : lane == 1 ? ay0 : lane == 2 ? az0 : lane == 3 ? ax1 : lane == 4 ? ay1 : az1;
                if (integ) kick_drift(a, s.v, s.q);
                s_stage[warp * 6 + lane] = s.q;
            }
            __syncthreads();
            // publish: lanes 0..7 of warp 0 write the block's 8 tagged sectors {x, y, z, step} in ONE
            // store instruction; no fence: every sector validates itself (see the header comment)
            double* sec = gbuf + ((size_t)(st & 1) * T + t) * C * GB * 4;
            const long long tag = (long long)st;
            if (tid < GB)
                st_sector(sec + (size_t)(c * GB + tid) * 4, s_stage[3 * tid], s_stage[3 * tid + 1], s_stage[3 * tid + 2],
                          __longlong_as_double(tag));
            tick(2);
            // (4a) wait until every producer's LAST sector carries this step's tag (one 8-byte poll per producer)
            const int nxt = s.cur ^ 1;
            if (tid < C) {
                const double* sentinel = sec + ((size_t)tid * GB + GB - 1) * 4 + 3;
                const long long t0 = clock64();
                while (__double_as_longlong(ld_strong_d(sentinel)) != tag) {
                    if (clock64() - t0 > SPIN_LIMIT) {
                        s_abort = 1;
                        atomicExch(status, 1);
                        break;
                    }
                }
            }
            tick(3);
            __syncthreads();
            if (s_abort) return;
            tick(4);
            // (4b) fetch all sectors, coalesced (consecutive lanes, consecutive sectors); a sector whose tag is
            //      not this step's yet (its store was overtaken by the sentinel's) is simply read again
            {
                double* dst = s_pos(t, nxt);
                const int nsec = C * GB;
#pragma unroll
                for (int k0 = 0; k0 < 8; k0 += 4) {
                    double x[4], y[4], z[4], g[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int b = tid + GT * (k0 + k);
                        if (b < nsec) ld_sector(sec + (size_t)b * 4, x[k], y[k], z[k], g[k]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int b = tid + GT * (k0 + k);
                        if (b < nsec) {
                            const long long t0 = clock64();
                            while (__double_as_longlong(g[k]) != tag) {
                                ld_sector(sec + (size_t)b * 4, x[k], y[k], z[k], g[k]);
                                if (clock64() - t0 > SPIN_LIMIT) {
                                    s_abort = 1;
                                    atomicExch(status, 1);
                                    break;
                                }
                            }
                            dst[3 * b] = x[k], dst[3 * b + 1] = y[k], dst[3 * b + 2] = z[k];
                        }
                    }
                }
            }
            tick(5);
            __syncthreads();
            if (s_abort) return;
            s.cur = nxt;
            // (5) observers of this step (hw5.cu:241-309), evaluated redundantly by every thread
            observe(s, t);
            tick(6);
        }
        any = false;
#pragma unroll
        for (int t = 0; t < T; t++) any |= ts[t].active;
    }

    if (PROFILE && tid == 0 && c == 0) {
        long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        for (int k = 0; k < 7; k++) prof[k] = pacc[k];
        prof[7] = (clock64() - c0) * 1000 / (g1 - g0 > 0 ? g1 - g0 : 1);  // SM MHz over the step loop
    }
    // write back
#pragma unroll
    for (int t = 0; t < T; t++) {
        const TrajDesc& d = descs[t];
        TState& s = ts[t];
        if (integ) {
            d.q[my_comp * n + my_body] = s.q;
            d.v[my_comp * n + my_body] = s.v;
        }
        if (c == 0) {
            if (tid < s.n_dev) d.ev->reach_step[tid] = s.my_reach;
            if (tid == 0) {
                d.ev->min_d2 = s.min_d2;
                d.ev->argmin_step = s.argmin_step;
                d.ev->hit_step = s.hit_step;
                d.ev->destroyed_step = s.destroyed_step;
                d.ev->cost = s.cost;
                d.ev->steps_done = s.step;
                d.ev->n_reach = s.n_dev;
                if (s.kind == NB_KIND_Q3 && s.destroyed_step != -2) d.m[s.DD] = 0.0;  // hw5.cu:306
            }
        }
    }
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int blocks_for(int n) { return (n + GB - 1) / GB; }
int npad_for(int n) { return ((blocks_for(n) * GB + 31) / 32) * 32; }
size_t smem_for(int n, int T) { return (size_t)T * 7 * npad_for(n) * sizeof(double); }

struct WsLayout {
    size_t gbuf_bytes, flags_bytes, total;
};
WsLayout ws_layout(int n, int T) {
    WsLayout w;
    const size_t C = blocks_for(n);
    w.gbuf_bytes = 2 * (size_t)T * C * GB * 4 * sizeof(double);  // [parity][T][C*8] sectors {x, y, z, step tag}
    w.flags_bytes = 0;
    w.total = w.gbuf_bytes + w.flags_bytes + 256;
    return w;
}

template <int MATH, int T>
int launch_t(int n, const TrajDesc* descs, const double* fst, void* ws, cudaStream_t stream) {
    const int C = blocks_for(n), npad = npad_for(n);
    const size_t smem = smem_for(n, T);
    const WsLayout w = ws_layout(n, T);
    double* gbuf = (double*)ws;
    int* status = (int*)((char*)ws + w.gbuf_bytes);
    long long* prof = (long long*)((char*)ws + w.gbuf_bytes + 64);
    static const bool profile = env_int("NB_GRID_PROFILE", 0) != 0;
    auto kern = (T == 1 && profile) ? grid_traj_kernel<MATH, T, (T == 1)> : grid_traj_kernel<MATH, T, false>;
    NB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NB_CUDA(cudaMemsetAsync(ws, 0, w.total, stream));  // tag 0 = no step; also clears status
    int npad_arg = npad;
    void* args[] = {(void*)&descs, (void*)&fst, (void*)&gbuf, (void*)&prof, (void*)&status, (void*)&npad_arg};
    NB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(C), dim3(GT), args, smem, stream));
    count_launch();
    int h_status = 0;
    NB_CUDA(cudaMemcpyAsync(&h_status, status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    NB_CUDA(cudaStreamSynchronize(stream));
    if (T == 1 && profile) {
        long long h[8];
        NB_CUDA(cudaMemcpy(h, prof, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "grid profile (clk, block 0 thread 0): gm+sync %lld | pairs %lld | reduce+integrate+publish %lld | "
                        "sentinel wait %lld | sync %lld | fetch %lld | sync+observe %lld | SM clock %lld MHz\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    if (h_status != 0) {
        set_error_detail("grid trajectory kernel: exchange spin timed out (blocks not co-resident?)");
        return NB_ERR_CUDA;
    }
    return NB_OK;
}

template <int MATH>
int launch_m(int T, int n, const TrajDesc* descs, const double* fst, void* ws, cudaStream_t stream) {
    switch (T) {
        case 1: return launch_t<MATH, 1>(n, descs, fst, ws, stream);
        case 2: return launch_t<MATH, 2>(n, descs, fst, ws, stream);
        case 3: return launch_t<MATH, 3>(n, descs, fst, ws, stream);
        default: return launch_t<MATH, 4>(n, descs, fst, ws, stream);
    }
}

}  // namespace

// the largest group of trajectories one launch takes for this n (shared-memory bound)
static int group_size(int n) {
    int T = MAX_T;
    while (T > 1 && smem_for(n, T) > 220 * 1024) T--;
    return T;
}

bool grid_traj_supported(int gpu, int n, int n_traj) {
    (void)n_traj;
    static const int enabled = env_int("NB_GRID", 1);
    static const int min_n = env_int("NB_GRID_MIN_N", 256);
    if (!enabled || n < min_n || n > NB_MAX_SMALL_N) return false;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, gpu) != cudaSuccess) return false;
    if (!p.cooperativeLaunch) return false;
    if (blocks_for(n) > p.multiProcessorCount) return false;
    if (smem_for(n, 1) > (size_t)p.sharedMemPerBlockOptin) return false;
    return true;
}

size_t grid_traj_workspace_bytes(int n, int n_traj) {
    (void)n_traj;
    return ws_layout(n, MAX_T).total;
}

int launch_grid_traj(int math, int n, int n_traj, const TrajDesc* descs, const double* fst, int gpu, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
    (void)gpu;
    if (ws_bytes < ws_layout(n, MAX_T).total) return NB_ERR_ARG;
    const int G = group_size(n);
    for (int t0 = 0; t0 < n_traj; t0 += G) {  // groups of up to G trajectories run in lock step
        const int T = n_traj - t0 < G ? n_traj - t0 : G;
        // STRICT never comes here: its ascending-j sum is the single-block kernel's (nb_host.cu)
        if (math != NB_MATH_FAST) return NB_ERR_UNSUPPORTED;
        int rc = launch_m<MATH_FAST>(T, n, descs + t0, fst, ws, stream);
        if (rc) return rc;
    }
    return NB_OK;
}

}  // namespace nb
