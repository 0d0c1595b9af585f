"""One ensemble launch of 148 x 1024-body systems (traj_sym_kernel): the command ncu profiles."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
S, steps = 148, int(sys.argv[1]) if len(sys.argv) > 1 else 40
base = nb.read_input(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "testcases", "b1024.in"))
q = np.tile(base.q, (S, 1)); v = np.stack([base.v * (1 + 1e-9 * k) for k in range(S)])
m = np.tile(base.m, (S, 1)); dev = np.tile(base.is_device, (S, 1))
ev, secs = nb.ensemble_run(q, v, m, dev, [base.planet] * S, [base.asteroid] * S, kind=nb.KIND_Q2, step_end=steps)
print("148 systems x %d steps: %.4f s = %.1f us/step" % (steps, secs, secs / steps * 1e6))
