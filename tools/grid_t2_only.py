"""Two b1024 trajectories per launch (query 1 + query 2 of the chain plan) for a few chunks: the command ncu profiles for the
two-systems-per-launch form of the grid kernel."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
case = sys.argv[1] if len(sys.argv) > 1 else "b1024"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24576
s = nb.read_input(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "testcases", case + ".in"))
t0 = time.perf_counter()
ans = nb.solve(s, gpus=[0], n_steps=steps)
print("%s solve over %d steps on 1 GPU: wall %.3f s, kernels %.3f s" % (case, steps, time.perf_counter() - t0, ans.gpu_seconds), flush=True)
