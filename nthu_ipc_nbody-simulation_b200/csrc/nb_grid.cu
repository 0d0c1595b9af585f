// Whole-GPU cooperative persistent trajectory kernel: ONE system (or up to 4 systems of the same n
// in lock step) spread over up to 128 SMs, all steps and observers in one launch.  This is the
// low-step-latency path for the b512 / b1024 queries: a single block needs ~133 us per b1024 step
// (nb_traj.cu), the FP64 floor over 128 SMs is ~1.05 us.
//
// Replaces the host-driven per-step launch sequence of the reference (hw5.cu:368-404, 387-403,
// 489-508: 3-4 kernel launches per step, 600 000-800 000 per trajectory).  Arithmetic: nbody.cc:51-89.
//
// Block c owns bodies [8c, 8c+8) and is warp-specialised (384 threads):
//   * 8 COMPUTE warps (2 per scheduler): warp w works on the body pair 8c + 2(w&3), +1 against half
//     (w>>2) of every chunk of j records, its lanes taking consecutive slots; two i-bodies share every
//     j record read from shared memory (shared-memory traffic = half the FP64 issue rate).  A
//     transposing xor butterfly (15 instead of 30 64-bit shuffles) leaves the pair's sums in lanes 0
//     and 16, the two j-halves meet in shared memory and threads 0..7 integrate one body each
//     (v += a*dt, q += v*dt) and publish it.
//   * 4 FETCH warps: warp f looks after producer blocks 32f..32f+31 (one per lane) = "chunk" f: waits
//     for their sectors of the current step, copies them into shared memory and completes the chunk's
//     mbarrier.  The compute warps consume chunk after chunk as they arrive, so the all-to-all
//     exchange streams underneath the pair loop instead of in front of it, and with more than one
//     trajectory per launch one system's exchange hides behind another's arithmetic.
//
// Exchange, measured design (tools/microbench/exchange_bench.cu, profiles/r01_exchange_microbench.md):
// on B200 a release/acquire hop costs ~0.4-0.5 us per fence (MEMBAR.ALL.GPU) and an atomic-counter
// grid barrier ~1.5 us, while an un-fenced store -> poll hop is ~0.37 us.  So there are no fences, no
// atomics and no barrier object: every body is published as one naturally aligned 32-byte sector
// {x, y, z, step tag} with a single 256-bit store, into a global record array double-buffered by step
// parity (L2 resident).  A sector is the unit the L2 reads and writes, so a reader sees it entirely
// old or entirely new and the tag validates the data it travels with; no ordering between different
// sectors is assumed anywhere.  A fetch lane polls the tag of its producer's last sector (a block's 8
// sectors leave in one store instruction), then loads the 8 sectors with 256-bit loads, re-reading
// any whose tag is still old.  Reuse of a parity buffer is safe without fences: a block overwrites
// step s-2's sectors only after it has consumed every block's step s-1 sectors, which each block
// publishes only after its own reads of step s-2 have completed (data dependence).  Spins are bounded
// (clock64) and raise `status`.
//
// Gravity devices are kept out of the streamed records' mass column (G*m = 0 there): their pair terms
// are added at the end of the pair loop, when all positions of the previous step are in and the
// observers (hit test, missile reach: hw5.cu:241-309) have decided whether the device still has mass.
//
// Shared memory per trajectory: pos[2][3*nslot] + gm[nslot]; body 8p+k of chunk f sits at slot
// f*256 + k*32 + (p-32f): fetch lanes write consecutive 24-byte slots and compute lanes read them
// (bank-conflict free for 8-byte accesses).  Co-residency of all blocks is guaranteed by the
// cooperative launch (grid <= SM count, 1 block/SM).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "nb_internal.h"
#include "nb_math.cuh"

namespace nb {
namespace {

constexpr int GB = 8;           // bodies per block
constexpr int NCW = 8;          // compute warps: 4 body pairs x 2 j-halves
constexpr int NFW = 4;          // fetch warps = chunks of 32 producer blocks
constexpr int GT = 32 * (NCW + NFW);
constexpr int CHUNK = 32 * GB;  // slots per chunk
constexpr int MAX_T = 4;        // trajectories per launch (shared memory: 56 B x nslot each)
constexpr long long SPIN_LIMIT = 4000000000LL;  // ~2 s of SM clocks

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory"); }

struct TState {  // per trajectory, per COMPUTE thread (registers after unrolling)
    int n_dev, kind, Ps, As, DDs, DD, step, step_begin, step_end;  // Ps/As/DDs: slots of planet/asteroid/device
    bool q3_armed, active, observed;
    double min_d2, cost;
    int argmin_step, hit_step, destroyed_step;
    // device bookkeeping (lane k < n_dev, every compute warp; reach steps are reported by block 0 warp 0)
    int my_dev, my_slot, my_reach;
    double my_m0;
    // integrator threads (tid < 8): the body's velocity and position
    double v[3], q[3];
    double fst_next;  // |sin| of the next step, prefetched one step ahead
};

// PROFILE: block 0 thread 0 accumulates clock64 per phase into prof[0..7] (NB_GRID_PROFILE=1, T == 1 only)
template <int MATH, int T, bool PROFILE>
__global__ void __launch_bounds__(GT, 1)
grid_traj_kernel(const TrajDesc* __restrict__ descs, const double* __restrict__ fst, double* __restrict__ gbuf,
                 long long* __restrict__ prof, int* __restrict__ status, int nchunk, unsigned poll_sleep_ns) {
    extern __shared__ double smem[];
    __shared__ volatile int s_abort;
    __shared__ volatile int s_stop[T];
    __shared__ double s_part[2][NCW][6];  // per warp: partial sums {ax,ay,az} of body A then body B over its j-half
    // chunk f of trajectory t, stage = step parity: full = the fetch warp has filled it (1 arrival), empty = the 8
    // compute warps are done with it.  A stage's barriers complete once per two steps, and the fetch warp waits
    // for "empty" before refilling, so a phase can never be lapped by a waiter (no parity aliasing).
    __shared__ alignas(8) uint64_t s_full[T][NFW][2];
    __shared__ alignas(8) uint64_t s_empty[T][NFW][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, C = gridDim.x;
    const int n = descs[0].n;
    const int nslot = nchunk * CHUNK;
    auto slot_of = [&](int b) {
        const int p = b / GB, k = b % GB;
        return (p >> 5) * CHUNK + k * 32 + (p & 31);
    };
    // per trajectory: pos[2][3*nslot] (slot-major x,y,z) then gm[nslot]
    auto s_pos = [&](int t, int buf) { return smem + (size_t)t * 7 * nslot + (size_t)buf * 3 * nslot; };
    auto s_gm = [&](int t) { return smem + (size_t)t * 7 * nslot + 6 * nslot; };

    if (tid == 0) {
        s_abort = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            // a trajectory that stopped in an earlier launch never steps again: keep its fetch warps out
            s_stop[t] = (descs[t].kind >= NB_KIND_Q2 && descs[t].ev->hit_step != -2) ? 1 : 0;
            for (int f = 0; f < NFW; f++)
                for (int b = 0; b < 2; b++) {
                    mbar_init(&s_full[t][f][b], 1);
                    mbar_init(&s_empty[t][f][b], NCW);
                }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // positions of step_begin into buffer (step_begin & 1); static G*m (0 for devices and padding)
#pragma unroll
    for (int t = 0; t < T; t++) {
        const TrajDesc& d = descs[t];
        double* p0 = s_pos(t, d.step_begin & 1);
        double* p1 = s_pos(t, (d.step_begin & 1) ^ 1);
        for (int sl = tid; sl < nslot; sl += GT) {
            const int f = sl / CHUNK, k = (sl % CHUNK) / 32, lp = sl % 32;
            const int p = 32 * f + lp;
            const int b = p < C ? p * GB + k : n;
            double x = 0, y = 0, z = 0, g = 0;
            if (b < n) {
                x = d.q[b], y = d.q[b + n], z = d.q[b + 2 * n];
                g = d.is_device[b] ? 0.0 : gm_eff(d.m[b], false, 0.0);
            }
            p0[3 * sl] = x, p0[3 * sl + 1] = y, p0[3 * sl + 2] = z;
            p1[3 * sl] = 0.0, p1[3 * sl + 1] = 0.0, p1[3 * sl + 2] = 0.0;
            s_gm(t)[sl] = g;
        }
    }
    __syncthreads();

    if (warp >= NCW) {
        // ------------------------------------------------------------------ FETCH warps
        const int f = warp - NCW;
        const int p = 32 * f + lane;  // producer block this lane looks after
        const bool have = p < C;
        int fstep[T], fend[T];
        bool factive[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            fstep[t] = descs[t].step_begin;
            fend[t] = descs[t].step_end;
            factive[t] = f < nchunk;
        }
        bool any = f < nchunk;
        while (any) {
            any = false;
#pragma unroll
            for (int t = 0; t < T; t++) {
                if (!factive[t]) continue;
                if (fstep[t] >= fend[t]) {
                    factive[t] = false;
                    continue;
                }
                const int st = ++fstep[t];
                const long long tag = (long long)st;
                const double* sec = gbuf + ((size_t)(st & 1) * T + t) * C * GB * 4;
                const double* mine = sec + (size_t)(have ? p : 0) * GB * 4;
                // wait until every producer of the chunk has published step st (or the trajectory stopped)
                bool stopped = false;
                {
                    const long long t0 = clock64();
                    bool ready = !have;
                    for (;;) {
                        if (!ready) ready = __double_as_longlong(ld_strong_d(mine + (GB - 1) * 4 + 3)) == tag;
                        if (__all_sync(0xffffffffu, ready)) break;
                        if (poll_sleep_ns) __nanosleep(poll_sleep_ns);  // leave the issue slots and the L2 to the others
                        if (s_stop[t] || s_abort) {
                            stopped = true;
                            break;
                        }
                        if (clock64() - t0 > SPIN_LIMIT) {
                            s_abort = 1;
                            atomicExch(status, 1);
                            stopped = true;
                            break;
                        }
                    }
                }
                stopped = __any_sync(0xffffffffu, stopped);
                if (stopped) {
                    factive[t] = false;
                    continue;
                }
                {   // the stage's previous contents (step st-2) must have been consumed by the compute warps
                    const int uses = (st - descs[t].step_begin) >> 1;
                    if (uses >= 1) {
                        const long long t0 = clock64();
                        while (!mbar_try_wait(&s_empty[t][f][st & 1], (uint32_t)(uses - 1) & 1u)) {
                            if (s_abort || clock64() - t0 > SPIN_LIMIT) {
                                s_abort = 1;
                                break;
                            }
                        }
                    }
                }
                if (have) {
                    double x[GB], y[GB], z[GB], g[GB];
#pragma unroll
                    for (int k = 0; k < GB; k++) ld_sector(mine + 4 * k, x[k], y[k], z[k], g[k]);
                    double* dst = s_pos(t, st & 1) + 3 * (f * CHUNK + lane);
                    const long long t0 = clock64();
#pragma unroll
                    for (int k = 0; k < GB; k++) {
                        while (__double_as_longlong(g[k]) != tag) {  // overtaken by the sentinel's store: read again
                            ld_sector(mine + 4 * k, x[k], y[k], z[k], g[k]);
                            if (clock64() - t0 > SPIN_LIMIT) {
                                s_abort = 1;
                                atomicExch(status, 1);
                                break;
                            }
                        }
                        double* o = dst + 3 * 32 * k;
                        o[0] = x[k], o[1] = y[k], o[2] = z[k];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_full[t][f][st & 1]);
                any = true;
            }
            if (s_abort) break;
        }
        return;
    }

    // ---------------------------------------------------------------------- COMPUTE warps
    const int pairA = c * GB + 2 * (warp & 3);  // the warp's two bodies: pairA, pairA + 1
    const int half = warp >> 2;
    const int my_body = c * GB + tid;            // integrator threads tid < 8
    const bool integ = tid < GB && my_body < n;
    const int slotA = slot_of(min(pairA, GB * C - 1)), slotB = slot_of(min(pairA + 1, GB * C - 1));

    TState ts[T];
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;
    auto tick = [&](int phase) {
        if (PROFILE) {
            const long long now = clock64();
            pacc[phase] += now - pt;
            pt = now;
        }
    };
#pragma unroll
    for (int t = 0; t < T; t++) {
        const TrajDesc& d = descs[t];
        TState& s = ts[t];
        s.n_dev = d.n_dev, s.kind = d.kind, s.DD = d.destroy_device;
        s.Ps = slot_of(d.planet), s.As = slot_of(d.asteroid);
        s.DDs = (s.DD >= 0 && s.DD < n) ? slot_of(s.DD) : 0;
        s.step = s.step_begin = d.step_begin, s.step_end = d.step_end;
        s.min_d2 = d.ev->min_d2, s.argmin_step = d.ev->argmin_step, s.hit_step = d.ev->hit_step;
        s.destroyed_step = d.ev->destroyed_step, s.cost = d.ev->cost;
        s.q3_armed = (s.kind == NB_KIND_Q3) && s.DD >= 0 && s.DD < n && d.m[s.DD] != 0.0;
        s.active = !((s.kind >= NB_KIND_Q2) && s.hit_step != -2);
        s.observed = d.ev->steps_done >= s.step;
        s.my_dev = -1, s.my_slot = 0, s.my_reach = -2, s.my_m0 = 0.0;
        if (lane < s.n_dev) {
            s.my_dev = d.dev_index[lane];
            s.my_slot = slot_of(s.my_dev);
            s.my_m0 = d.m[s.my_dev];
            s.my_reach = d.ev->reach_step[lane];
        }
        s.fst_next = fst[s.step + 1];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s.v[k] = integ ? d.v[k * n + my_body] : 0.0;
            s.q[k] = integ ? d.q[k * n + my_body] : 0.0;
        }
    }

    // observers of step s.step on its (complete) position buffer; uniform over all compute threads except the
    // per-device reach test
    auto observe = [&](TState& s, int t) {
        const double* pos = s_pos(t, s.step & 1);
        const double* p = pos + 3 * s.Ps;
        const double* a = pos + 3 * s.As;
        const double px = p[0], py = p[1], pz = p[2];
        const double d2 = dist2_rn(px, py, pz, a[0], a[1], a[2]);
        if (d2 < s.min_d2) {  // hw5.cu:245-247
            s.min_d2 = d2;
            s.argmin_step = s.step;
        }
        if (s.kind == NB_KIND_Q2 && s.my_dev >= 0 && s.my_reach == -2) {  // hw5.cu:265-287
            const double* dv = pos + 3 * s.my_slot;
            const double md = __dmul_rn(MISSILE_STEP, (double)s.step);
            if (dist2_rn(px, py, pz, dv[0], dv[1], dv[2]) < __dmul_rn(md, md)) s.my_reach = s.step;
        }
        if (s.kind >= NB_KIND_Q2) {
            if (d2 < PLANET_RADIUS2) {  // nbody.cc:134, hw5.cu:295-298
                s.hit_step = s.step;
                s.active = false;
            } else if (s.q3_armed && s.destroyed_step == -2) {  // hw5.cu:299-307
                const double* dv = pos + 3 * s.DDs;
                const double md = __dmul_rn(MISSILE_STEP, (double)s.step);
                if (dist2_rn(px, py, pz, dv[0], dv[1], dv[2]) < __dmul_rn(md, md)) {
                    s.destroyed_step = s.step;
                    s.cost = __dadd_rn(1e5, __dmul_rn(1e3, __dmul_rn((double)(s.step + 1), DT)));
                }
            }
        }
        s.observed = true;
    };

    bool any = false;
#pragma unroll
    for (int t = 0; t < T; t++) any |= ts[t].active;
    long long g0 = 0, c0 = 0;
    if (PROFILE) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        c0 = clock64();
    }
    int pbuf = 0;  // s_part double buffer

    while (any) {
#pragma unroll
        for (int t = 0; t < T; t++) {
            TState& s = ts[t];
            if (!s.active) continue;  // uniform across the grid
            if (PROFILE) pt = clock64();
            // state: integrator threads hold q, v of step s.step; buffer (s.step & 1) receives the positions of
            // step s.step chunk by chunk (it is complete already when s.step == step_begin)
            const bool last = s.step >= s.step_end;
            const bool wait_chunks = s.step > s.step_begin;
            const uint32_t parity = (uint32_t)((s.step - s.step_begin - 1) >> 1) & 1u;  // fill index of this stage
            const int stage = s.step & 1;
            const double* cpos = s_pos(t, s.step & 1);
            const double* cg = s_gm(t);
            double ax0 = 0, ay0 = 0, az0 = 0, ax1 = 0, ay1 = 0, az1 = 0;
            double xa = 0, ya = 0, za = 0, xb = 0, yb = 0, zb = 0;
            bool aborted = false;
            // (1) forces on the warp's two bodies (nbody.cc:56-74), chunk by chunk as the positions arrive.
            //     The warp's own bodies live in the block's own chunk, so that chunk is awaited first.
            const int fown = c >> 5;
            for (int ff = 0; ff < nchunk; ff++) {
                const int f = ff == 0 ? fown : (ff <= fown ? ff - 1 : ff);
                if (wait_chunks) {
                    const long long t0 = clock64();
                    while (!mbar_try_wait(&s_full[t][f][stage], parity)) {
                        if (s_abort || clock64() - t0 > SPIN_LIMIT) {
                            aborted = true;
                            break;
                        }
                    }
                    if (aborted) break;
                }
                if (ff == 0) {
                    xa = cpos[3 * slotA], ya = cpos[3 * slotA + 1], za = cpos[3 * slotA + 2];
                    xb = cpos[3 * slotB], yb = cpos[3 * slotB + 1], zb = cpos[3 * slotB + 2];
                }
                if (!last) {
                    const int j0 = f * CHUNK + half * (CHUNK / 2) + lane;
#pragma unroll
                    for (int k = 0; k < CHUNK / 64; k++) {
                        const int j = j0 + 32 * k;
                        const double jx = cpos[3 * j], jy = cpos[3 * j + 1], jz = cpos[3 * j + 2], jg = cg[j];
                        pair<MATH>(xa, ya, za, jx, jy, jz, jg, ax0, ay0, az0);
                        pair<MATH>(xb, yb, zb, jx, jy, jz, jg, ax1, ay1, az1);
                    }
                }
            }
            // an abort (exchange timeout) is seen by every compute warp in this same iteration, before compute_bar
            if (aborted) return;
            tick(0);
            // (2) all positions of step s.step are in: observers (hw5.cu:241-309)
            if (!s.observed) observe(s, t);
            if (last) s.active = false;
            if (!s.active) {
                if (tid == 0) s_stop[t] = 1;  // releases the fetch warps of this trajectory
                continue;
            }
            const int st = ++s.step;
            s.observed = false;
            // (3) the gravity devices' pair terms with G*m_eff(st) (nbody.cc:14-16, 61-64); a destroyed device has mass 0
            if (half == 0 && s.my_dev >= 0) {
                const bool gone = (s.kind == NB_KIND_Q3) && s.my_dev == s.DD && s.destroyed_step != -2;
                const double g = gm_eff(gone ? 0.0 : s.my_m0, true, s.fst_next);
                const double* dv = cpos + 3 * s.my_slot;
                pair<MATH>(xa, ya, za, dv[0], dv[1], dv[2], g, ax0, ay0, az0);
                pair<MATH>(xb, yb, zb, dv[0], dv[1], dv[2], g, ax1, ay1, az1);
            }
            s.fst_next = fst[st + 1];
            // this warp is done with the positions of step st-1: hand the stage back to the fetch warps
            __syncwarp();
            if (lane == 0)
                for (int f = 0; f < nchunk; f++) mbar_arrive(&s_empty[t][f][stage]);
            // transposing butterfly: lanes with bit 4 clear keep body A's sums, the others body B's
            {
                const bool hi = lane & 16;
                const double s0 = hi ? ax0 : ax1, s1 = hi ? ay0 : ay1, s2 = hi ? az0 : az1;  // what the partner keeps
                double k0 = hi ? ax1 : ax0, k1 = hi ? ay1 : ay0, k2 = hi ? az1 : az0;        // what this lane keeps
                k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                k2 += __shfl_xor_sync(0xffffffffu, s2, 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    k0 += __shfl_xor_sync(0xffffffffu, k0, o);
                    k1 += __shfl_xor_sync(0xffffffffu, k1, o);
                    k2 += __shfl_xor_sync(0xffffffffu, k2, o);
                }
                if ((lane & 15) == 0) {
                    double* dstp = &s_part[pbuf][warp][hi ? 3 : 0];
                    dstp[0] = k0, dstp[1] = k1, dstp[2] = k2;
                }
            }
            tick(1);
            compute_bar();
            // (4) threads 0..7: a = half0 + half1; v += a*dt; q += v*dt (nbody.cc:77-88); publish the body as one
            //     tagged sector {x, y, z, step}: the 8 sectors of the block leave in one store instruction, unfenced
            if (tid < GB) {
                if (integ) {
                    const double* p0 = &s_part[pbuf][tid >> 1][3 * (tid & 1)];
                    const double* p1 = &s_part[pbuf][4 + (tid >> 1)][3 * (tid & 1)];
#pragma unroll
                    for (int k = 0; k < 3; k++) kick_drift(p0[k] + p1[k], s.v[k], s.q[k]);
                }
                double* sec = gbuf + ((size_t)(st & 1) * T + t) * C * GB * 4;
                st_sector(sec + (size_t)(c * GB + tid) * 4, s.q[0], s.q[1], s.q[2], __longlong_as_double((long long)st));
            }
            pbuf ^= 1;
            tick(2);
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
#pragma unroll
            for (int k = 0; k < 3; k++) {
                d.q[k * n + my_body] = s.q[k];
                d.v[k * n + my_body] = s.v[k];
            }
        }
        if (c == 0 && warp == 0) {
            if (lane < s.n_dev) d.ev->reach_step[lane] = s.my_reach;
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
int nchunk_for(int n) { return (blocks_for(n) + 31) / 32; }
size_t smem_for(int n, int T) { return (size_t)T * 7 * nchunk_for(n) * CHUNK * sizeof(double); }

struct WsLayout {
    size_t gbuf_bytes, total;
};
WsLayout ws_layout(int n, int T) {
    WsLayout w;
    const size_t C = blocks_for(n);
    w.gbuf_bytes = 2 * (size_t)T * C * GB * 4 * sizeof(double);  // [parity][T][C*8] sectors {x, y, z, step tag}
    w.total = w.gbuf_bytes + 256;
    return w;
}

template <int MATH, int T>
int launch_t(int n, const TrajDesc* descs, const double* fst, void* ws, cudaStream_t stream) {
    const int C = blocks_for(n);
    const size_t smem = smem_for(n, T);
    const WsLayout w = ws_layout(n, T);
    double* gbuf = (double*)ws;
    int* status = (int*)((char*)ws + w.gbuf_bytes);
    long long* prof = (long long*)((char*)ws + w.gbuf_bytes + 64);
    static const bool profile = env_int("NB_GRID_PROFILE", 0) != 0;
    auto kern = (T == 1 && profile) ? grid_traj_kernel<MATH, T, (T == 1)> : grid_traj_kernel<MATH, T, false>;
    NB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NB_CUDA(cudaMemsetAsync(ws, 0, w.total, stream));  // tag 0 = no step; also clears status
    int nchunk = nchunk_for(n);
    static unsigned poll_sleep_ns = (unsigned)env_int("NB_GRID_POLL_NS", 0);
    void* args[] = {(void*)&descs, (void*)&fst, (void*)&gbuf, (void*)&prof, (void*)&status, (void*)&nchunk,
                    (void*)&poll_sleep_ns};
    NB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(C), dim3(GT), args, smem, stream));
    count_launch();
    int h_status = 0;
    NB_CUDA(cudaMemcpyAsync(&h_status, status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    NB_CUDA(cudaStreamSynchronize(stream));
    if (T == 1 && profile) {
        long long h[8];
        NB_CUDA(cudaMemcpy(h, prof, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "grid profile (clk, block 0 thread 0): chunk waits + pairs %lld | observers + device pairs + butterfly %lld | "
                        "barrier + integrate + publish %lld | SM clock %lld MHz\n",
                h[0], h[1], h[2], h[7]);
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
    // measured on B200, b1024: 4 systems per launch fit (224 KiB + 1.3 KiB static) but the T = 4 instantiation
    // spills and runs the four-trajectory solve in 3.80 s; groups of 3 + 1 take 3.07 s
    int T = MAX_T;
    while (T > 1 && smem_for(n, T) > 220 * 1024) T--;
    return T;
}

bool grid_traj_supported(int gpu, int n, int n_traj) {
    (void)n_traj;
    static const int enabled = env_int("NB_GRID", 1);
    static const int min_n = env_int("NB_GRID_MIN_N", 128);
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
