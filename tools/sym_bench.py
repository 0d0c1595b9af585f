"""Quick A/B of the symmetric stepper against the row kernel on one GPU (n = 65536): per-step and per-kernel times."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
s = nb.synthetic_system(n, seed=42)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def run(name, sh):
    for _ in range(3):
        sh.advance(1)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record(); sh.advance(1); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    nb.profile_enable(True)
    for _ in range(steps):
        flush.zero_(); sh.advance(1)
    torch.cuda.synchronize()
    ams, cnt = nb.profile_read()
    nb.profile_enable(False)
    pairs = n * (n - 1)
    print("%-28s step %.3f ms = %.3e pairs/s = %.1f%% of 37.2 TF | accel kernel %.3f ms (%d launches) = %.1f%%" % (
        name, ms, pairs / ms * 1e3, pairs / ms * 1e3 * 20 / 37.2e12 * 100, ams / max(cnt, 1), cnt,
        pairs / (ams / max(cnt, 1)) * 1e3 * 20 / 37.2e12 * 100), flush=True)
    return sh.positions(), sh.velocities()
q0, v0 = run("row kernel (ShardedSystem)", nb.ShardedSystem(s, device="cuda:0"))
sy = nb.SymShardedSystem(s, device="cuda:0")
q1, v1 = run("symmetric (SymShardedSystem)", sy)
sy.close()
print("max |dv|/|v| sym vs row after %d steps: %.3e ; q max abs diff %.3e (ulp %.3e)" % (
    2 * steps + 3, np.max(np.abs(v1 - v0) / np.abs(v0)), np.abs(q1 - q0).max(), np.spacing(np.abs(q0)).max()))
for w in (2, 8):
    lw = nb.SymLocalWorld(s, w, device="cuda:0")
    t0 = time.perf_counter()
    lw.advance(4)
    q, v = lw.positions(), lw.velocities()
    print("local world %d: 4 steps in %.3f s (serialised on one GPU), status ok" % (w, time.perf_counter() - t0))
    lw.close()
print("fp64 peak dfma: %.2f TF" % nb.fp64_peak(0))
