"""CPU, world_size 2, gloo: the body-sharded driver's partition + per-step all-gather logic, with
the local step replaced by a host function built on the oracle (the CUDA step cannot run here).
The two-rank result must equal the single-rank oracle run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _oracle_local_step(orc, mode):
    """Host twin of nb_large_step: integrates rows [i_begin, i_begin+i_count) against all bodies."""
    from importlib import import_module

    sh = import_module("nthu_ipc_nbody-simulation_b200.sharded")

    def step_fn(step, n, i_begin, i_count, pos4, pos4_out, vel, m0, isdev, scratch):
        p = pos4.numpy()
        q = np.ascontiguousarray(p[:, :3].T).reshape(-1)
        v = np.zeros(3 * n)
        v.reshape(3, n)[:, i_begin:i_begin + i_count] = vel.numpy()
        orc.run_steps(mode, n, q, v, m0.numpy(), isdev.numpy(), step - 1, step, nthreads=1)
        sl = slice(i_begin, i_begin + i_count)
        out = pos4_out.numpy()
        out[sl, :3] = q.reshape(3, n).T[sl]
        out[sl, 3] = sh.gm_eff(m0.numpy(), isdev.numpy(), step + 1)[sl]
        vel.copy_(torch.from_numpy(v.reshape(3, n)[:, sl].copy()))

    return step_fn


def _worker(rank, world, port, n, steps, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from importlib import import_module

    nb = import_module("nthu_ipc_nbody-simulation_b200")
    import oracle_binding as orc

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = nb.synthetic_system(n, seed=9)
        sh = nb.ShardedSystem(s, rank=rank, world=world, local_step=_oracle_local_step(orc, orc.MODE_SQRT3))
        assert (sh.i_begin, sh.i_count) == (rank * n // world, n // world)
        assert sh.bytes_exchanged_per_step() == 32 * (n // world) * (world - 1)
        sh.advance(steps)
        q, v = sh.positions(), sh.velocities()
        if rank == 0:
            ret["q"], ret["v"] = q, v
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_equals_single_rank_oracle():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from importlib import import_module

    nb = import_module("nthu_ipc_nbody-simulation_b200")
    import oracle_binding as orc

    n, steps, world = 64, 5, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, steps, ret), nprocs=world, join=True)
    s = nb.synthetic_system(n, seed=9)
    qo, vo = s.q.copy(), s.v.copy()
    orc.run_steps(orc.MODE_SQRT3, n, qo, vo, s.m, s.is_device, 0, steps, nthreads=1)
    assert np.array_equal(ret["q"], qo) and np.array_equal(ret["v"], vo)


def test_partition_rejects_uneven_split():
    sys.path.insert(0, ROOT)
    from importlib import import_module

    sh = import_module("nthu_ipc_nbody-simulation_b200.sharded")
    assert sh.partition(65536, 8, 3) == (24576, 8192)
    with pytest.raises(ValueError):
        sh.partition(1000, 3, 0)


def test_host_pack_layout():
    sys.path.insert(0, ROOT)
    from importlib import import_module

    nb = import_module("nthu_ipc_nbody-simulation_b200")
    s = nb.synthetic_system(16, seed=1)
    p = nb.sharded.host_pack(s.q, s.m, s.is_device, 7)
    assert p.shape == (16, 4) and np.array_equal(p[:, 0], s.q[:16]) and np.array_equal(p[:, 2], s.q[32:])
    assert p[0, 3] == 6.674e-11 * s.m[0]
    f = nb.sharded.fst(7)
    assert p[15, 3] == 6.674e-11 * (s.m[15] + (0.5 * s.m[15]) * f)
