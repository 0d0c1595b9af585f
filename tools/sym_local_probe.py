"""Per-rank acceleration-kernel time of the symmetric stepper with all ranks emulated on one GPU (no NVLink involved)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
n = 65536
s = nb.synthetic_system(n, seed=42)
for world in (1, 2, 4, 8):
    w = nb.SymLocalWorld(s, world, device="cuda:0")
    w.advance(2)
    torch.cuda.synchronize()
    times = []
    for step in range(3):
        ev = []
        for r in w.ranks:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r.step_phase(1); b.record()
            ev.append((a, b))
        ti = []
        for r in w.ranks:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r.step_phase(2); b.record()
            ti.append((a, b))
        torch.cuda.synchronize()
        times.append(([a.elapsed_time(b) for a, b in ev], [a.elapsed_time(b) for a, b in ti]))
    acc, integ = times[-1]
    print("world %d: accel ms per rank %s | integrate ms per rank %s | ideal accel %.3f" % (
        world, " ".join("%.3f" % x for x in acc), " ".join("%.3f" % x for x in integ), 2.92 / world), flush=True)
    w.close()
