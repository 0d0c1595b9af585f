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
    """SURVEY.md §8d config C5 through the library's generator (`nb_generate_system`, the one `nbtool gen` uses):
    body 0 = planet, body 1 = asteroid, the last 4 bodies (fewer for tiny n) are gravity devices."""
    from . import generate_system

    return generate_system(n, seed, min(4, max(0, n - 2)))


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


class P2PShardedSystem:
    """Body-sharded driver whose exchange is fused into the integrate kernel (north star (d), "P2P stores"):
    every rank stores its new pos4 rows straight into every rank's buffer over NVLink (peer-mapped CUDA IPC
    pointers) and signals per-parity arrival counters; a device-side wait precedes the next step's reads.
    No NCCL call on the data path: torch.distributed only carries the 64-byte IPC handles once."""

    def __init__(self, system, rank=0, world=1, device=None, math=0, group=None, step0=0):
        import ctypes as C

        import torch

        from . import _check, lib

        if not torch.cuda.is_available():
            raise RuntimeError("P2PShardedSystem needs a CUDA device (no CPU fallback)")
        self.torch, self.C, self.L = torch, C, lib()
        L = self.L
        self.n, self.rank, self.world, self.group, self.math = system.n, rank, world, group, math
        self.i_begin, self.i_count = partition(system.n, world, rank)
        self.step = self.step0 = step0
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        n, ib, ic = self.n, self.i_begin, self.i_count
        self.m0 = torch.from_numpy(system.m.copy()).to(self.device)
        self.isdev = torch.from_numpy(system.is_device.copy()).to(self.device)
        self.vel = torch.from_numpy(system.v.reshape(3, n)[:, ib:ib + ic].copy()).to(self.device).contiguous()
        self.scratch = torch.empty(int(L.nb_large_scratch_bytes(n, ic)), dtype=torch.uint8, device=self.device)
        self.blocks = int(L.nb_large_blocks_per_step(ic))
        # own exchange buffers: pos4[2] and {counter[2], status}
        self.own = []
        self.ctr_bytes = int(L.nb_large_p2p_counter_bytes())
        for nbytes in (32 * n, 32 * n, self.ctr_bytes + 64):
            ptr = C.c_void_p()
            _check(L.nb_dev_alloc(nbytes, C.byref(ptr)))
            self.own.append(ptr.value)
        self.opened = []
        peers = [list(self.own)]
        if world > 1:
            import torch.distributed as dist

            handles = []
            for ptr in self.own:
                h = C.create_string_buffer(64)
                _check(L.nb_ipc_export(C.c_void_p(ptr), h))
                handles.append(h.raw)
            allh = [None] * world
            dist.all_gather_object(allh, handles, group=group)
            peers = []
            for r in range(world):
                if r == rank:
                    peers.append(list(self.own))
                    continue
                ptrs = []
                for raw in allh[r]:
                    ptr = C.c_void_p()
                    _check(L.nb_ipc_open(raw, C.byref(ptr)))
                    ptrs.append(ptr.value)
                    self.opened.append(ptr.value)
                peers.append(ptrs)
        vp = C.c_void_p * world
        self.peer_pos = [vp(*[peers[r][b] for r in range(world)]) for b in range(2)]
        self.peer_ctr = vp(*[peers[r][2] for r in range(world)])
        self.cnt = [0, 0]  # steps executed per parity
        self.cur = 0
        qd = torch.from_numpy(system.q.copy()).to(self.device)
        st = torch.cuda.current_stream().cuda_stream
        _check(L.nb_large_pack(math, n, C.c_void_p(qd.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                               C.c_void_p(self.isdev.data_ptr()), self.step + 1, C.c_void_p(self.own[0]), C.c_void_p(st)))
        torch.cuda.current_stream().synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier(group=group)  # every peer has mapped every buffer before the first remote store

    def bytes_exchanged_per_step(self):
        return 0 if self.world == 1 else 32 * self.i_count * (self.world - 1)

    def _wait_for(self, written_step):
        """Enqueue the device-side wait for the rows of `written_step` (all ranks, all blocks)."""
        from . import _check

        C = self.C
        par = written_step & 1
        target = self.blocks * self.cnt[par]
        st = self.torch.cuda.current_stream().cuda_stream
        _check(self.L.nb_large_wait_p2p(C.c_void_p(self.own[2]), par, self.world, target,
                                        C.c_void_p(self.own[2] + self.ctr_bytes), C.c_void_p(st)))

    def advance(self, steps=1):
        from . import _check

        C, L, torch = self.C, self.L, self.torch
        for _ in range(steps):
            self.step += 1
            # rows of step-1 are awaited INSIDE the acceleration kernel, per source rank (0 = first step: packed locally)
            wait_target = self.blocks * self.cnt[(self.step - 1) & 1] if self.step - 1 > self.step0 else 0
            st = torch.cuda.current_stream().cuda_stream
            _check(L.nb_large_step_p2p(self.math, self.step, self.n, self.i_begin, self.i_count,
                                       C.c_void_p(self.own[self.cur]), self.peer_pos[self.cur ^ 1], self.peer_ctr,
                                       self.world, self.rank, wait_target, C.c_void_p(self.own[2] + self.ctr_bytes),
                                       C.c_void_p(self.vel.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                                       C.c_void_p(self.isdev.data_ptr()), C.c_void_p(self.scratch.data_ptr()), C.c_void_p(st)))
            self.cnt[self.step & 1] += 1
            self.cur ^= 1

    def positions(self):
        """Planar q[3n] of all bodies (host numpy); waits for the last step's rows of every rank."""
        from . import _check

        C, torch = self.C, self.torch
        if self.step > self.step0:
            self._wait_for(self.step)
        st = torch.cuda.current_stream().cuda_stream
        p = np.empty((self.n, 4))
        _check(self.L.nb_dev_copy(p.ctypes.data_as(C.c_void_p), C.c_void_p(self.own[self.cur]), p.nbytes, 1, C.c_void_p(st)))
        status = np.zeros(1, dtype=np.int32)
        _check(self.L.nb_dev_copy(status.ctypes.data_as(C.c_void_p), C.c_void_p(self.own[2] + self.ctr_bytes), 4, 1, C.c_void_p(st)))
        if status[0] != 0:
            raise RuntimeError("P2P exchange: a peer's rows did not arrive (wait timed out)")
        return np.ascontiguousarray(p[:, :3].T).reshape(-1)

    def velocities(self):
        torch = self.torch
        if self.world == 1:
            return self.vel.detach().to("cpu").numpy().reshape(-1).copy()
        import torch.distributed as dist

        parts = [torch.empty_like(self.vel) for _ in range(self.world)]
        dist.all_gather(parts, self.vel, group=self.group)
        return np.concatenate([p.to("cpu").numpy() for p in parts], axis=1).reshape(-1)

    def close(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)  # nobody is still storing into a buffer that is about to go
        for ptr in self.opened:
            self.L.nb_ipc_close(self.C.c_void_p(ptr))
        for ptr in self.own:
            self.L.nb_dev_free(self.C.c_void_p(ptr))
        self.opened, self.own = [], []


class SymShardedSystem:
    """Body-sharded driver on the SYMMETRIC stepper (csrc/nb_sym.cu, FAST math): every unordered pair is evaluated
    once (10 FP64 instructions per ordered pair), rank p takes the block pairs (p, p) .. (p, p + P/2); the partial
    accelerations of a body are stored by the acceleration kernel straight into the memory of the rank that owns it
    (peer-mapped PJ buffers over NVLink), its integrate kernel sums them in a fixed order and stores the new pos4
    row into every rank's buffer.  Arrival counters + in-kernel waits: two launches per step, no NCCL call and no
    host synchronisation on the data path (torch.distributed only carries the 64-byte IPC handles once)."""

    def __init__(self, system, rank=0, world=1, device=None, math=0, group=None, step0=0, _local_world=None):
        import ctypes as C

        import torch

        from . import MATH_FAST, _check, lib

        if not torch.cuda.is_available():
            raise RuntimeError("SymShardedSystem needs a CUDA device (no CPU fallback)")
        if math != MATH_FAST:
            raise ValueError("the symmetric stepper is FAST math only (STRICT keeps the reference's ascending-j sum)")
        self.torch, self.C, self.L = torch, C, lib()
        L = self.L
        self.n, self.rank, self.world, self.group, self.math = system.n, rank, world, group, math
        self.i_begin, self.i_count = partition(system.n, world, rank)
        self.step = self.step0 = step0
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        n, ib, ic = self.n, self.i_begin, self.i_count
        self.h = C.c_void_p()
        _check(L.nb_sym_create(n, world, rank, C.byref(self.h)))
        self.m0 = torch.from_numpy(system.m.copy()).to(self.device)
        self.isdev = torch.from_numpy(system.is_device.copy()).to(self.device)
        self.vel = torch.from_numpy(system.v.reshape(3, n)[:, ib:ib + ic].copy()).to(self.device).contiguous()
        self.ctr_bytes = int(L.nb_sym_counter_bytes())
        self.pj_bytes = int(L.nb_sym_pj_bytes(self.h))
        # own peer-visible buffers: pos4[2], PJ (partial accelerations of my bodies), {counters, status}
        self.own = []
        for nbytes in (32 * n, 32 * n, self.pj_bytes, self.ctr_bytes + 64):
            ptr = C.c_void_p()
            _check(L.nb_dev_alloc(nbytes, C.byref(ptr)))
            self.own.append(ptr.value)
        self.opened = []
        self.cur = 0
        qd = torch.from_numpy(system.q.copy()).to(self.device)
        st = torch.cuda.current_stream().cuda_stream
        _check(L.nb_large_pack(math, n, C.c_void_p(qd.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                               C.c_void_p(self.isdev.data_ptr()), self.step + 1, C.c_void_p(self.own[0]), C.c_void_p(st)))
        torch.cuda.current_stream().synchronize()
        if _local_world is not None:
            return  # SymLocalWorld connects the ranks (all on this GPU)
        peers = [list(self.own)]
        if world > 1:
            import torch.distributed as dist

            handles = []
            for ptr in self.own:
                h = C.create_string_buffer(64)
                _check(L.nb_ipc_export(C.c_void_p(ptr), h))
                handles.append(h.raw)
            allh = [None] * world
            dist.all_gather_object(allh, handles, group=group)
            peers = []
            for r in range(world):
                if r == rank:
                    peers.append(list(self.own))
                    continue
                ptrs = []
                for raw in allh[r]:
                    ptr = C.c_void_p()
                    _check(L.nb_ipc_open(raw, C.byref(ptr)))
                    ptrs.append(ptr.value)
                    self.opened.append(ptr.value)
                peers.append(ptrs)
        self._connect(peers)
        if world > 1:
            import torch.distributed as dist

            dist.barrier(group=group)  # every peer has mapped every buffer before the first remote store

    def _connect(self, peers):
        vp = self.C.c_void_p * self.world
        self.peer_pos = [vp(*[peers[r][b] for r in range(self.world)]) for b in range(2)]
        self.peer_pj = vp(*[peers[r][2] for r in range(self.world)])
        self.peer_ctr = vp(*[peers[r][3] for r in range(self.world)])

    def step_phase(self, phases):
        """phases: 1 = acceleration kernel of the next step, 2 = its integrate kernel (advances the step), 3 = both."""
        from . import _check

        C, L = self.C, self.L
        st = self.torch.cuda.current_stream().cuda_stream
        step = self.step + 1
        _check(L.nb_sym_step_phase(self.h, step, phases, C.c_void_p(self.own[self.cur]), self.peer_pos[self.cur ^ 1],
                                   self.peer_pj, self.peer_ctr, C.c_void_p(self.own[3] + self.ctr_bytes),
                                   C.c_void_p(self.vel.data_ptr()), C.c_void_p(self.m0.data_ptr()),
                                   C.c_void_p(self.isdev.data_ptr()), C.c_void_p(st)))
        if phases & 2:
            self.step = step
            self.cur ^= 1

    def bytes_exchanged_per_step(self):
        """Bytes this rank stores into peers' memory per step: pos4 rows to everybody + partial accelerations."""
        if self.world == 1:
            return 0
        return 32 * self.i_count * (self.world - 1) + int(self.L.nb_sym_remote_partial_bytes(self.h))

    def advance(self, steps=1):
        for _ in range(steps):
            self.step_phase(3)

    def pairs_per_step(self):
        """Ordered pair interactions one step of THIS rank covers (2 x unordered symmetric pairs + the one-sided rows,
        self pairs included)."""
        return int(self.L.nb_sym_pairs(self.h, None, None))

    def step_host(self, q_own_host, v_own_host):
        """The run_step operator with HOST buffers (nbody.cc:51-54 signature, sharded): q_own_host / v_own_host are
        pinned planar [3, n/world] tensors with THIS rank's positions and velocities (in: before the step, out: after).
        Every call copies them host -> device, publishes the rows to every rank (peer stores), runs the step's two
        kernels, extracts the new rows and copies them device -> host.  Returns (h2d bytes, d2h bytes)."""
        from . import _check

        C, L, torch = self.C, self.L, self.torch
        if not hasattr(self, "_q_own"):
            self._q_own = torch.empty(3 * self.i_count, dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        step = self.step + 1
        _check(L.nb_sym_step_host(self.h, step, C.c_void_p(q_own_host.data_ptr()), C.c_void_p(v_own_host.data_ptr()),
                                  C.c_void_p(self._q_own.data_ptr()), C.c_void_p(self.own[self.cur]), self.peer_pos[self.cur],
                                  self.peer_pos[self.cur ^ 1], self.peer_pj, self.peer_ctr,
                                  C.c_void_p(self.own[3] + self.ctr_bytes), C.c_void_p(self.vel.data_ptr()),
                                  C.c_void_p(self.m0.data_ptr()), C.c_void_p(self.isdev.data_ptr()), C.c_void_p(st)))
        self.step = step
        self.cur ^= 1
        nbytes = (q_own_host.numel() + v_own_host.numel()) * 8
        return nbytes, nbytes

    def positions(self):
        """Planar q[3n] of all bodies (host numpy); waits for the last step's rows of every rank."""
        from . import _check

        C, torch = self.C, self.torch
        st = torch.cuda.current_stream().cuda_stream
        if self.step > self.step0 and self.world > 1:
            _check(self.L.nb_sym_wait_positions(self.h, C.c_void_p(self.own[3]), C.c_void_p(self.own[3] + self.ctr_bytes),
                                                C.c_void_p(st)))
        p = np.empty((self.n, 4))
        _check(self.L.nb_dev_copy(p.ctypes.data_as(C.c_void_p), C.c_void_p(self.own[self.cur]), p.nbytes, 1, C.c_void_p(st)))
        status = np.zeros(1, dtype=np.int32)
        _check(self.L.nb_dev_copy(status.ctypes.data_as(C.c_void_p), C.c_void_p(self.own[3] + self.ctr_bytes), 4, 1, C.c_void_p(st)))
        if status[0] != 0:
            raise RuntimeError("symmetric P2P exchange: a peer's rows or partials did not arrive (wait timed out, status %d)" % status[0])
        return np.ascontiguousarray(p[:, :3].T).reshape(-1)

    def velocities(self):
        torch = self.torch
        if self.world == 1:
            return self.vel.detach().to("cpu").numpy().reshape(-1).copy()
        import torch.distributed as dist

        parts = [torch.empty_like(self.vel) for _ in range(self.world)]
        dist.all_gather(parts, self.vel, group=self.group)
        return np.concatenate([p.to("cpu").numpy() for p in parts], axis=1).reshape(-1)

    def close(self):
        if not self.own:
            return
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)  # nobody is still storing into a buffer that is about to go
        for ptr in self.opened:
            self.L.nb_ipc_close(self.C.c_void_p(ptr))
        for ptr in self.own:
            self.L.nb_dev_free(self.C.c_void_p(ptr))
        self.L.nb_sym_destroy(self.h)
        self.opened, self.own = [], []


class SymLocalWorld:
    """`world` ranks of the symmetric stepper on ONE GPU and one stream: every rank's acceleration kernel of a step is
    enqueued before any rank's integrate kernel, so each in-kernel arrival wait is already satisfied when it is
    reached.  Same kernels, same peer stores and counters as one process per GPU (the peers' buffers are ordinary
    device pointers here): the multi-rank path, testable and checkable on a 1-GPU box."""

    def __init__(self, system, world, device=None):
        self.ranks = [SymShardedSystem(system, rank=r, world=world, device=device, _local_world=self) for r in range(world)]
        peers = [list(r.own) for r in self.ranks]
        for r in self.ranks:
            r._connect(peers)
        self.world = world

    def advance(self, steps=1):
        for _ in range(steps):
            for r in self.ranks:
                r.step_phase(1)
            for r in self.ranks:
                r.step_phase(2)

    def positions(self):
        return self.ranks[0].positions()

    def velocities(self):
        return np.concatenate([r.vel.detach().to("cpu").numpy() for r in self.ranks], axis=1).reshape(-1)

    def close(self):
        self.ranks[0].torch.cuda.synchronize()
        for r in self.ranks:
            r.world_for_close, r.world = r.world, 1  # no process group here
            r.close()


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
