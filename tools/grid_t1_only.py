"""One b1024 trajectory (query 1) over the grid kernel for a few thousand steps: the command ncu profiles for the
one-system-per-launch form of the kernel (the >= 4 GPU b1024 path)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
case = sys.argv[1] if len(sys.argv) > 1 else "b1024"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
s = nb.read_input(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "testcases", case + ".in"))
t = nb.Trajectory(s, nb.KIND_Q1)
t.run(2000)
t0 = time.perf_counter()
ev = t.run(steps)
dt = time.perf_counter() - t0
print("%s Q1 %d steps: %.3f s = %.3f us/step" % (case, steps - 2000, dt, dt / (steps - 2000) * 1e6), flush=True)
