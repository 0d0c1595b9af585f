"""Body-sharded large-N driver (north-star (d)): rank p integrates bodies
[p*n/P, (p+1)*n/P) against all n bodies with the CUDA kernels of csrc/nb_large.cu, then the new
pos4 records are all-gathered (NCCL over NVLink, in place, on the compute stream) so that every
rank holds all positions for the next step.  Masses are static and |sin| is a per-step scalar every
rank computes itself, so the 32-byte pos4 record {x, y, z, G*m_eff(next step)} is the only thing
that crosses the wire: 32*n/P bytes per rank per step.

The reference has no counterpart (every GPU holds all bodies and no GPU-to-GPU exchange exists,
hw5.cu:343-350); the arithmetic is run_step's (nbody.cc:51-89).

torch is plumbing here (device buffers, streams, torch.distributed); the math is in the C-ABI
library.  `local_step` may be replaced by a host function so that the partition / exchange logic is
testable with gloo on CPU (tests/test_sharded_gloo.py); the default is the CUDA path and raises
without a GPU.
"""
from __future__ import annotations

import numpy as np


def partition(n, world, rank):
    """Contiguous, equal shards (n must divide evenly: the in-place all-gather needs equal counts)."""
    if n % world != 0:
        raise ValueError("n=%d is not divisible by world=%d" % (n, world))
    cnt = n // world
    return rank * cnt, cnt


def synthetic_system(n, seed=42):
    """positions uniform in a cube of side 1e13 m centred at
    (-2.0e20, -2.9e20, 1.8e18), velocities N(0, (1e7 m/s)^2), masses log-uniform in [1e20, 1e30] kg,
    body 0 = planet, body 1 = asteroid, the last 4 bodies are gravity devices."""
    from . import System

    rng = np.random.default_rng(seed)
    centre = np.array([-2.0e20, -2.9e20, 1.8e18])
    q = np.empty(3 * n)
    for c in range(3):
        q[c * n:(c + 1) * n] = centre[c] + (rng.random(n) - 0.5) * 1e13
    v = rng.normal(0.0, 1e7, 3 * n)
    m = 10.0 ** rng.uniform(20.0, 30.0, n)
    dev = np.zeros(n, dtype=np.uint8)
    dev[max(2, n - 4):] = 1
    return System(n, 0, 1, q, v, m, dev)


def fst(step):
    """|sin(step*dt/6000)| exactly as nbody.cc:14-16,63 evaluates it (libm sin through math.sin)."""
    import math

    return abs(math.sin((step * 60.0) / 6000))


def gm_eff(m0, is_device, step):
    """G*m_eff(step) per body: G*(m0 + (0.5*m0)*fst) for devices (nbody.cc:15), G*m0 otherwise."""
    m0 = np.asarray(m0, dtype=np.float64)
    mj = np.where(np.asarray(is_device) != 0, m0 + (0.5 * m0) * fst(step), m0)
    return 6.674e-11 * mj


def host_pack(q_planar, m0, is_device, step_next):
    """Host twin of nb_large_pack: planar q -> pos4[n][4] = {x, y, z, G*m_eff(step_next)}."""
    n = len(m0)
    p = np.empty((n, 4))
    p[:, :3] = np.asarray(q_planar).reshape(3, n).T
    p[:, 3] = gm_eff(m0, is_device, step_next)
    return p


def _cuda_local_step(math):
    import ctypes as C

    import torch

    from . import _check, lib

    L = lib()

    def step_fn(step, n, i_begin, i_count, pos4, pos4_out, vel, m0, isdev, scratch):
        st = torch.cuda.current_stream().cuda_stream
        _check(L.nb_large_step(math, step, n, i_begin, i_count, C.c_void_p(pos4.data_ptr()),
                               C.c_void_p(pos4_out.data_ptr()), C.c_void_p(vel.data_ptr()),
                               C.c_void_p(m0.data_ptr()), C.c_void_p(isdev.data_ptr()),
                               C.c_void_p(scratch.data_ptr()), C.c_void_p(st)))

    return step_fn


class ShardedSystem:
    def __init__(self, system, rank=0, world=1, device=None, math=0, group=None, local_step=None, step0=0):
        import torch

        from . import lib

        self.torch = torch
        self.n, self.rank, self.world, self.group, self.math = system.n, rank, world, group, math
        self.i_begin, self.i_count = partition(system.n, world, rank)
        self.step = step0
        host_mode = local_step is not None
        if not host_mode:
            if not torch.cuda.is_available():
                raise RuntimeError("ShardedSystem needs a CUDA device (no CPU fallback)")
            device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
            local_step = _cuda_local_step(math)
        else:
            device = torch.device("cpu")
        self.device, self.local_step = device, local_step
        n, ib, ic = self.n, self.i_begin, self.i_count
        f64 = torch.float64
        self.m0 = torch.from_numpy(system.m.copy()).to(device)
        self.isdev = torch.from_numpy(system.is_device.copy()).to(device)
        v = system.v.reshape(3, n)[:, ib:ib + ic].copy()
        self.vel = torch.from_numpy(v).to(device).contiguous()
        self.pos4 = [torch.zeros(n, 4, dtype=f64, device=device) for _ in range(2)]
        self.cur = 0
        if host_mode:
            self.scratch = None
            self.pos4[0].copy_(torch.from_numpy(host_pack(system.q, system.m, system.is_device, self.step + 1)))
        else:
            import ctypes as C

            from . import _check

            L = lib()
            self.scratch = torch.empty(int(L.nb_large_scratch_bytes(n, ic)), dtype=torch.uint8, device=device)
            qd = torch.from_numpy(system.q.copy()).to(device)
            st = torch.cuda.current_stream().cuda_stream
            _check(L.nb_large_pack(math, n, C.c_void_p(qd.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                                   C.c_void_p(self.isdev.data_ptr()), self.step + 1,
                                   C.c_void_p(self.pos4[0].data_ptr()), C.c_void_p(st)))
            torch.cuda.current_stream().synchronize()

    def bytes_exchanged_per_step(self):
        return 0 if self.world == 1 else 32 * self.i_count * (self.world - 1)

    def advance(self, steps=1):
        """`steps` time steps; each = local kernels + (world > 1) in-place all-gather of pos4."""
        torch = self.torch
        for _ in range(steps):
            self.step += 1
            src, dst = self.pos4[self.cur], self.pos4[self.cur ^ 1]
            self.local_step(self.step, self.n, self.i_begin, self.i_count, src, dst, self.vel, self.m0, self.isdev,
                            self.scratch)
            if self.world > 1:
                import torch.distributed as dist

                shard = dst[self.i_begin:self.i_begin + self.i_count].view(-1)
                dist.all_gather_into_tensor(dst.view(-1), shard, group=self.group)
            self.cur ^= 1

    def step_host(self, q_host, v_host):
        """The run_step operator with HOST buffers (nbody.cc:51-54 signature, sharded): q_host is a
        pinned planar [3n] tensor holding ALL positions (in: state before the step, out: after),
        v_host a pinned planar [3, i_count] tensor with this rank's velocities (in/out).  Every call
        copies host->device, packs, steps, all-gathers, unpacks and copies device->host."""
        import ctypes as C

        torch = self.torch
        from . import _check, lib

        L = lib()
        if not hasattr(self, "_qd"):
            self._qd = torch.empty(3 * self.n, dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        self._qd.copy_(q_host, non_blocking=True)
        self.vel.copy_(v_host, non_blocking=True)
        self.step += 1
        src, dst = self.pos4[self.cur], self.pos4[self.cur ^ 1]
        _check(L.nb_large_pack(self.math, self.n, C.c_void_p(self._qd.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                               C.c_void_p(self.isdev.data_ptr()), self.step, C.c_void_p(src.data_ptr()), C.c_void_p(st)))
        self.local_step(self.step, self.n, self.i_begin, self.i_count, src, dst, self.vel, self.m0, self.isdev,
                        self.scratch)
        if self.world > 1:
            import torch.distributed as dist

            shard = dst[self.i_begin:self.i_begin + self.i_count].view(-1)
            dist.all_gather_into_tensor(dst.view(-1), shard, group=self.group)
        self.cur ^= 1
        _check(L.nb_large_unpack(self.n, C.c_void_p(dst.data_ptr()), C.c_void_p(self._qd.data_ptr()), C.c_void_p(st)))
        q_host.copy_(self._qd, non_blocking=True)
        v_host.copy_(self.vel, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h2d = q_host.numel() * 8 + v_host.numel() * 8
        return h2d, h2d  # bytes host->device, device->host

    def positions(self):
        """Planar q[3n] of all bodies (host numpy)."""
        p = self.pos4[self.cur].detach().to("cpu").numpy()
        return np.ascontiguousarray(p[:, :3].T).reshape(-1)

    def velocities(self):
        """Planar v[3n] of all bodies (host numpy), gathered from the shards."""
        torch = self.torch
        if self.world == 1:
            return self.vel.detach().to("cpu").numpy().reshape(-1).copy()
        import torch.distributed as dist

        parts = [torch.empty_like(self.vel) for _ in range(self.world)]
        dist.all_gather(parts, self.vel, group=self.group)
        return np.concatenate([p.to("cpu").numpy() for p in parts], axis=1).reshape(-1)
