"""b1024 query-1 trajectory through the grid kernel: us/step and (NB_GRID_PROFILE=1) the per-phase clocks."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
case = sys.argv[1] if len(sys.argv) > 1 else "b1024"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
s = nb.read_input(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "testcases", case + ".in"))
t = nb.Trajectory(s, nb.KIND_Q1)
t.run(2000)
t0 = time.perf_counter()
ev = t.run(steps)
dt = time.perf_counter() - t0
print("%s Q1 %d steps: %.3f s = %.3f us/step, min_d2 %r argmin %d" % (case, steps - 2000, dt, dt / (steps - 2000) * 1e6, ev.min_d2, ev.argmin_step), flush=True)
t0 = time.perf_counter()
ans = nb.solve(s, gpus=[0])
print("three-query solve on 1 GPU: wall %.3f s, kernels %.3f s -> %s" % (time.perf_counter() - t0, ans.gpu_seconds,
      nb.format_output(ans.min_dist, ans.hit_time_step, ans.gravity_device_id, ans.missile_cost).replace("\n", " | ")), flush=True)
