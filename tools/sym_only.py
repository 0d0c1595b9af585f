"""A few steps of the symmetric stepper alone (n = 65536, one GPU): the command ncu profiles."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("nthu_ipc_nbody-simulation_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s = nb.synthetic_system(n, seed=42)
sy = nb.SymShardedSystem(s, device="cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sy.advance(2)
e0.record(); sy.advance(steps); e1.record()
torch.cuda.synchronize()
print("sym n=%d: %.3f ms/step" % (n, e0.elapsed_time(e1) / steps))
sy.close()
