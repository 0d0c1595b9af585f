#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the N-body hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's own CPU program)

Metric (BASELINE.json): FP64 pair-interactions per second on the synthetic 65 536-body single system
(SURVEY.md §8d config C5), body-sharded over the N ranks with a per-step in-place NCCL all-gather of
the 32-byte pos4 records (strong scaling: the total work is fixed).  One "step" = one time step of
the whole system = n(n-1) ordered pair interactions.  The b1024 three-query end-to-end seconds —
the other half of the BASELINE metric — is reported in the same line under "b1024".

    value     pairs/s, device-resident state, CUDA events per step, max over ranks
    e2e       the same metric through the run_step operator with HOST (pinned) buffers: every step
              copies q (all bodies) and this rank's v host->device and back
    roofline  the acceleration kernel alone against the B200 FP64 peak (20 flop per pair)
    cpu_baseline  the unmodified reference program (oracle/_ref/nbody = samples/nbody.cc) on one host
              core, full b20 run (N=1, rank 0 only)
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "nthu_ipc_nbody-simulation_b200"
CASES = os.path.join(ROOT, "tests", "golden", "testcases")
FP64_PEAK_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: 148 SM x 64 FMA/clk x 2 x 1.965 GHz
PAIR_FLOPS = 20
METRIC = "fp64_pair_interactions_per_sec"
UNIT = "pairs/s"


def reference_run(case="b20"):
    """One full run of the unmodified reference program; returns (seconds, pair interactions)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nbody")
    kind = "reference"
    if not os.path.exists(exe):  # the reference did not travel: fall back to the oracle port, and say so
        exe = os.path.join(ROOT, "oracle", "_build", "nbody_oracle")
        kind = "port"
    inp = os.path.join(CASES, case + ".in")
    out = "/tmp/bench_ref_%d.out" % os.getpid()
    t0 = time.perf_counter()
    subprocess.check_call([exe, inp, out] + (["200000", "0", "1"] if kind == "port" else []))
    dt = time.perf_counter() - t0
    lines = open(out).read().split("\n")
    gold = open(os.path.join(CASES, case + ".out")).read().split("\n")
    if lines[0] != gold[0] or lines[1] != gold[1]:
        raise RuntimeError("reference output differs from the golden")
    n = int(open(inp).readline().split()[0])
    hit = int(lines[1])
    steps = 200000 + (hit if hit >= 0 else 200000)  # Q1 runs all steps, Q2 stops at the hit (nbody.cc:114-138)
    if kind == "port":
        kats = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_kats.json")))[case]
        for d in kats["devices"]:  # the port also answers query 3 (forked at the reach step)
            end = d["q3_hit_step"] if d["q3_hit_step"] >= 0 else 200000
            if d["reach_step"] >= 0 and d["reach_step"] < hit:
                steps += end - d["reach_step"]
    return dt, steps * n * (n - 1), kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        reference_run()
    tot_t, tot_p, kind = 0.0, 0, "reference"
    for _ in range(args.steps):
        dt, pairs, kind = reference_run()
        tot_t += dt
        tot_p += pairs
    val = tot_p / tot_t
    sample = "full b20.in run of samples/nbody.cc (query 1 + query 2, 338 784 steps x 380 ordered pairs) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "reference golden input b20.in",
        "config": {"workload": "synthetic 65536-body single system (reference arm: bounded CPU sample, see cpu_baseline.sample)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def init_nccl(dist, torch, dev):
    """init_process_group + the first collective with file descriptor 1 pointed at stderr: NCCL prints its version
    banner (NCCL_DEBUG=VERSION on these boxes) to stdout when the communicator is created, and stdout must carry the
    one JSON line only."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled DURING the timed region: NVML from a thread every ~2 ms
    (a timed region can be as short as 10 ms at 8 GPUs), `nvidia-smi -lms 20` as the fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    MASKS = [0x8, 0x40, 0x20, 0x4]  # nvmlClocksThrottleReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, uuid):
        self.proc, self.nvml, self.rows, self.lines, self.source = None, None, [], [], None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml, 2 ms"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 20"
        except Exception:
            pass
        if self.proc:
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                ts = time.perf_counter()
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((ts, sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        sm, mx, pw, reasons = [], None, [], set()
        if self.nvml:
            self._stop.set()
            self.t.join(timeout=1.0)
            mx = self.mx
            for ts, s_, p_, rs in self.rows:
                if ts < t0 or ts > t1:
                    continue
                sm.append(s_)
                pw.append(p_)
                for nme, msk in zip(self.NAMES, self.MASKS):
                    if rs & msk:
                        reasons.add(nme)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            for ts, ln in self.lines:
                if ts < t0 or ts > t1 + 0.1:
                    continue
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx = float(f[1])
                    pw.append(float(f[2]))
                    for nme, val in zip(self.NAMES, f[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(nme)
                except Exception:
                    continue
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"]}
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": self.source}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    nb = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or nb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        init_nccl(dist, torch, dev)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    system = nb.synthetic_system(n, seed=42)
    if args.exchange == "p2p":
        sh = nb.P2PShardedSystem(system, rank=rank, world=world, device=dev)
    else:
        sh = nb.ShardedSystem(system, rank=rank, world=world, device=dev)
    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid

    # ---- device-resident measurement ---------------------------------------------------------------
    sampler = ClockSampler(uuid) if rank == 0 else None
    t_warm = time.perf_counter()
    for _ in range(args.warmup):
        sh.advance(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = nb.kernel_launches()
    t0 = time.perf_counter()
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
        ev[k][0].record()
        sh.advance(1)
        ev[k][1].record()
    barrier()
    t1 = time.perf_counter()
    launches = nb.kernel_launches() - launches0
    clocks = None
    if sampler:
        # nvidia-smi samples every 20 ms: a timed region shorter than ~0.2 s is widened to warm-up + timed steps
        short = (t1 - t0) < 0.2
        clocks = sampler.stop(t_warm if short else t0, t1)
        clocks["window"] = "warmup+timed" if short else "timed"
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms)
    pairs_per_step = n * (n - 1)
    value = pairs_per_step * args.steps / (dev_ms * 1e-3)

    # ---- the acceleration kernel alone (roofline) ---------------------------------------------------
    nb.profile_enable(True)
    for _ in range(min(args.steps, 10)):
        if flush is not None:
            flush.zero_()
        sh.advance(1)
    torch.cuda.synchronize()
    accel_ms, accel_n = nb.profile_read()
    nb.profile_enable(False)
    accel_ms = max_over_ranks(accel_ms / max(accel_n, 1))
    flops_per_launch = PAIR_FLOPS * sh.i_count * (n - 1)
    achieved = flops_per_launch / (accel_ms * 1e-3) / 1e12
    peak_measured = nb.fp64_peak(local)

    # ---- end to end through the host-buffer operator ------------------------------------------------
    qh = torch.from_numpy(system.q.copy()).pin_memory()
    vh = torch.from_numpy(system.v.reshape(3, n)[:, sh.i_begin:sh.i_begin + sh.i_count].copy()).pin_memory()
    e2e_sys = nb.ShardedSystem(system, rank=rank, world=world, device=dev)
    for _ in range(args.warmup):
        h2d, d2h = e2e_sys.step_host(qh, vh)
    barrier()
    te0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        h2d, d2h = e2e_sys.step_host(qh, vh)
    e1.record()
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - te0)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = pairs_per_step * args.steps / max(e2e_ms * 1e-3, e2e_wall if world == 1 else 0.0)
    # the e2e path must produce the same state as the resident path after the same number of steps
    launches_total = nb.kernel_launches()

    # ---- b1024 three-query end to end (the other half of the BASELINE metric) -----------------------
    b1024 = None
    if not args.no_b1024:
        s1024 = nb.read_input(os.path.join(CASES, "b1024.in"))
        barrier()
        tb = time.perf_counter()
        ans, gsecs, pairs = nb.solve_distributed(s1024, rank, world, local)
        barrier()
        wall = max_over_ranks(time.perf_counter() - tb)
        text = nb.format_output(ans.min_dist, ans.hit_time_step, ans.gravity_device_id, ans.missile_cost)
        gold = open(os.path.join(CASES, "b1024.out")).read()
        gl, tl = gold.split("\n"), text.split("\n")
        ok = tl[1] == gl[1] and tl[2] == gl[2] and abs(float(tl[0]) - float(gl[0])) <= 1e-6 * float(gl[0])
        b1024 = {"solve_wall_s": wall, "gpu_s": gsecs, "pairs_per_s_gpu": pairs / gsecs if gsecs else None,
                 "trajectories": ans.n_trajectories, "matches_golden": bool(ok), "line1_byte_identical": tl[0] == gl[0],
                 "note": "in-process solve (contexts already created); as many GPUs as trajectories: Q1, Q2 and one Q3 trajectory per device from step 0, one per rank; fewer: Q1 on rank 0, Q2 -> Q3 candidates forked from Q2 on rank 1 (chain plan, nb_host.cu)"}

    # ---- CPU baseline: the unmodified reference program on this box's host cores --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt, pairs, kind = reference_run()
        cpu = {"value": pairs / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
               "sample": "one full b20.in run of samples/nbody.cc, serial (query 1 + query 2: 338 784 steps x 380 ordered pairs)",
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 65536-body single system, body-sharded, per-step pos4 all-gather" if n == 65536
                       else "synthetic %d-body single system, body-sharded" % n,
                       "n_bodies": n, "seed": 42, "bodies_per_rank": sh.i_count, "pairs_per_step": pairs_per_step,
                       "math": "fast (16 FP64 instr/pair)", "parallelism": "body-sharded x%d" % world,
                       "exchange": ("P2P stores fused into the integrate kernel (peer-mapped buffers, no NCCL on the data path)"
                                    if args.exchange == "p2p" else "in-place NCCL all-gather of pos4 rows") if world > 1 else "none (1 GPU)",
                       "l2": "not flushed" if flush is None else "flushed between timed steps (256 MiB memset, outside the per-step events)",
                       "exchange_bytes_per_rank_per_step": sh.bytes_exchanged_per_step()},
            "frac_of_fp64_peak": value * PAIR_FLOPS / (world * FP64_PEAK_NOMINAL_TFLOPS * 1e12),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": FP64_PEAK_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved / FP64_PEAK_NOMINAL_TFLOPS,
                         "traffic": 2.88e6 if (n == 65536 and world == 1) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture "
                                           "profiles/r01_large_accel_kernel.ncu-rep (n = 65536, 1 GPU)",
                         "kernel": "large_accel_kernel", "kernel_ms": accel_ms,
                         "peak_source": "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_measured_dfma": peak_measured, "frac_of_measured": achieved / peak_measured,
                         "flops_per_pair": PAIR_FLOPS, "fp64_instr_per_pair": 16,
                         "note": "compute-bound path: algorithmic bytes are 104 B x n per step (intensity ~12600 flop/B), "
                                 "so the roofline is the FP64 pipe, not HBM or tensor cores (SURVEY.md 8d)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "gpu_launches_total_process": launches_total,
            "clocks": clocks,
            "cpu_baseline": cpu,
            "b1024": b1024,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ensemble(args):
    """BASELINE config 4: 1024 independent 1024-body systems (member k = b1024.in with velocities scaled by
    1 + 1e-9 k), split over the ranks with no collective; one block per system, one launch."""
    import numpy as np
    import torch
    import torch.distributed as dist

    nb = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dist, torch, dev)
    S_total = args.systems
    base = nb.read_input(os.path.join(CASES, "b1024.in"))
    mine = list(range(rank, S_total, world))
    S, n = len(mine), base.n
    q0 = np.tile(base.q, (S, 1))
    v0 = np.stack([base.v * (1 + 1e-9 * k) for k in mine])
    m = np.tile(base.m, (S, 1))
    isdev = np.tile(base.is_device, (S, 1))
    pl, ast = [base.planet] * S, [base.asteroid] * S

    def once(steps):
        q, v = q0.copy(), v0.copy()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ev, secs = nb.ensemble_run(q, v, m, isdev, pl, ast, kind=nb.KIND_Q2, step_end=steps, gpu=local)
        wall = time.perf_counter() - t0
        t = torch.tensor([secs, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), ev, q

    once(max(args.warmup, 1))
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid) if rank == 0 else None
    ts0 = time.perf_counter()
    secs, wall, ev, q = once(args.steps)
    clocks = sampler.stop(ts0, time.perf_counter()) if sampler else None
    # member 0 of rank 0 is the golden system itself: its state must equal a plain trajectory's
    ok = True
    if rank == 0:
        t = nb.Trajectory(base, nb.KIND_Q2, gpu=local)
        t.run(args.steps)
        ok = bool(np.array_equal(t.state()[0], q[0]))
    pairs = S_total * args.steps * n * (n - 1)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": pairs / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic (b1024.in with scaled velocities)",
            "config": {"workload": "synthetic ensemble of %d independent 1024-body systems" % S_total,
                       "systems_per_rank": S, "parallelism": "ensemble x%d, no collective" % world,
                       "kernel": "traj_kernel (one block per system, whole run in one launch)"},
            "frac_of_fp64_peak": pairs / secs * PAIR_FLOPS / (world * FP64_PEAK_NOMINAL_TFLOPS * 1e12),
            "e2e": {"value": pairs / wall, "unit": UNIT, "h2d_bytes_per_step": int(S * n * 57 / args.steps),
                    "d2h_bytes_per_step": int(S * n * 48 / args.steps), "note": "states uploaded once per launch, not per step"},
            "roofline": {"bound": "fp64", "achieved": pairs / secs * PAIR_FLOPS / 1e12 / world, "peak": FP64_PEAK_NOMINAL_TFLOPS,
                         "unit": "TFLOP/s", "frac": pairs / secs * PAIR_FLOPS / (world * FP64_PEAK_NOMINAL_TFLOPS * 1e12),
                         "traffic": None, "kernel": "traj_kernel", "note": "per GPU; the whole launch is this kernel"},
            "clocks": clocks, "member0_equals_single_trajectory": ok, "gpu_launches": 2}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--no-b1024", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--workload", default="large", choices=["large", "ensemble"])
    ap.add_argument("--exchange", default=os.environ.get("NB_EXCHANGE", "p2p"), choices=["nccl", "p2p"])
    ap.add_argument("--systems", type=int, default=1024)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "ensemble":
        run_ensemble(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
