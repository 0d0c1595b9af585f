"""CPU tests of the symmetric large-N stepper's host logic (csrc/nb_sym.cu build_plan, no GPU needed):
the static schedule must cover every ordered pair of run_step (nbody.cc:57-60) exactly once, be balanced, and its
PI / PJ bookkeeping must reproduce the plain all-pairs acceleration when executed — on one rank and, through a
world_size-2 gloo exchange of the partial rows, on two."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
F_ONESIDED, F_LOAD, F_FLUSH = 1, 2, 4


@pytest.fixture(autouse=True)
def _row_size(nb):
    global SB
    SB = nb.lib().nb_sym_row_size()


def plans(nb, n, world, blocks):
    return [nb.sym_plan(n, world, r, blocks) for r in range(world)]


@pytest.mark.parametrize("n,world,blocks", [(3000, 1, 7), (4096, 1, 296), (4096, 2, 16), (6144, 3, 5), (6000, 2, 9),
                                            (8192, 4, 3), (8192, 8, 2), (5120, 5, 4)])
def test_plan_covers_every_ordered_pair_once(nb, n, world, blocks):
    cov = np.zeros((n, n), dtype=np.uint8)
    sym_total, one_total = 0, 0
    for r, (segs, bsb, pj_ptr, pj_list, sp, op) in enumerate(plans(nb, n, world, blocks)):
        assert bsb[0] == 0 and bsb[-1] == len(segs) and np.all(np.diff(bsb) >= 0)
        S = n // world
        for row0, rc, j0, j1, fl, slot, pjrow, src in segs:
            assert r * S <= row0 and row0 + rc <= (r + 1) * S, "rows are local"
            assert src * S <= j0 < j1 <= (src + 1) * S, "a j range lies in one rank's shard"
            cov[row0:row0 + rc, j0:j1] += 1
            if fl & F_ONESIDED:
                assert j0 >= row0 and j1 <= row0 + rc
            else:
                assert j1 <= row0 or j0 >= row0 + rc, "a symmetric segment never contains its own row"
                cov[j0:j1, row0:row0 + rc] += 1
                assert pjrow == r * nb.lib().nb_sym_rows(n, world) + (row0 - r * S) // nb.lib().nb_sym_row_stride(n, world)
        sym_total += sp
        one_total += op
    assert cov.min() == 1 and cov.max() == 1
    assert 2 * sym_total + one_total == n * n


def test_plan_runs_and_slots(nb):
    n, world, blocks = 16384, 4, 37
    for r in range(world):
        segs, bsb, pj_ptr, pj_list, sp, op = nb.sym_plan(n, world, r, blocks)
        slots = []
        for b in range(blocks):
            v = segs[bsb[b]:bsb[b + 1]]
            for i, g in enumerate(v):
                first = i == 0 or v[i - 1][0] != g[0]
                last = i + 1 == len(v) or v[i + 1][0] != g[0]
                assert bool(g[4] & F_LOAD) == first and bool(g[4] & F_FLUSH) == last
                if last:
                    slots.append(g[5])
        assert sorted(slots) == list(range(len(slots))), "every row run has its own PI slot"


@pytest.mark.parametrize("n,world,blocks", [(65536, 1, 296), (65536, 2, 296), (65536, 4, 296), (65536, 8, 148), (65536, 4, 148)])
def test_plan_is_balanced(nb, n, world, blocks):
    worst_rank = []
    for r in range(world):
        segs, bsb, *_ = nb.sym_plan(n, world, r, blocks)
        # a row occupies its block's register slots however many bodies it holds: the TIME of a segment goes with its j
        # range, not with its pair count
        cost, pairs = np.zeros(blocks), 0
        stride = nb.lib().nb_sym_row_stride(n, world)
        for b in range(blocks):
            for row0, rc, j0, j1, fl, *_ in segs[bsb[b]:bsb[b + 1]]:
                cost[b] += (j1 - j0) * (0.8 if fl & F_ONESIDED else 1.0)
                pairs += int(rc) * int(j1 - j0)
                assert rc <= stride <= SB
        assert cost.max() <= 1.06 * cost.mean(), (r, cost.max() / cost.mean())
        worst_rank.append(cost.sum())
        # rows of equal length: at most ~one row stride of register slots per row run is padding
        assert pairs >= 0.85 * cost.sum() * stride * 0.9
    assert max(worst_rank) <= 1.03 * min(worst_rank), "ranks carry equal work"


# ---- executing a plan on the host ------------------------------------------------------------------------------------
def pair_block(q, gm, I, J):
    """Accelerations of bodies I from bodies J and of J from I (plain FP64, nbody.cc:65-72 with G*m hoisted)."""
    d = q[J][None, :, :] - q[I][:, None, :]
    r2 = (d * d).sum(-1) + 1e-6
    s = r2 ** -1.5
    ai = (d * (gm[J][None, :] * s)[..., None]).sum(1)
    aj = -(d * (gm[I][:, None] * s)[..., None]).sum(0)
    return ai, aj


def execute_rank(nb, n, world, rank, blocks, q, gm, send_pj):
    """Runs rank's segments; a_i go to PI slots, a_j rows to send_pj(owner, pj_row, j0, j1, values)."""
    segs, bsb, pj_ptr, pj_list, *_ = nb.sym_plan(n, world, rank, blocks)
    pi = {}
    for b in range(blocks):
        acc = None
        for row0, rc, j0, j1, fl, slot, pjrow, src in segs[bsb[b]:bsb[b + 1]]:
            if fl & F_LOAD:
                acc = np.zeros((rc, 3))
            I, J = np.arange(row0, row0 + rc), np.arange(j0, j1)
            ai, aj = pair_block(q, gm, I, J)
            acc += ai
            if not fl & F_ONESIDED:
                send_pj(src, pjrow, j0, j1, aj)
            if fl & F_FLUSH:
                pi[slot] = (row0, acc.copy())
    return pi, pj_ptr, pj_list


def direct(q, gm):
    n = len(gm)
    a = np.zeros((n, 3))
    for i0 in range(0, n, 512):
        I = np.arange(i0, min(n, i0 + 512))
        a[I] = pair_block(q, gm, I, np.arange(n))[0]
    return a


def system(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1e13, 1e13, (n, 3)), 6.674e-11 * 10 ** rng.uniform(20, 30, n)


def test_executed_plan_equals_all_pairs_one_rank(nb):
    n, blocks = 2500, 11
    STRIDE1 = nb.lib().nb_sym_row_stride(n, 1)
    q, gm = system(n, 1)
    pj = {}
    pi, pj_ptr, pj_list = execute_rank(nb, n, 1, 0, blocks, q, gm, lambda o, row, j0, j1, v: pj.__setitem__((row, j0), v))
    a = np.zeros((n, 3))
    for slot in sorted(pi):
        row0, acc = pi[slot]
        a[row0:row0 + len(acc)] += acc
    seen_rows = {}
    for (row, j0), v in pj.items():
        a[j0:j0 + len(v)] += v
        for c in {j0 // STRIDE1, (j0 + len(v) - 1) // STRIDE1}:
            seen_rows.setdefault(c, set()).add(row)
    for c in range(len(pj_ptr) - 1):
        assert sorted(seen_rows.get(c, set())) == list(pj_list[pj_ptr[c]:pj_ptr[c + 1]])
    ref = direct(q, gm)
    assert np.max(np.abs(a - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-12


def _worker(rank, world, port, n, blocks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from importlib import import_module

    nb = import_module("nthu_ipc_nbody-simulation_b200")
    global SB
    SB = nb.lib().nb_sym_row_size()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = n // world
        rows = nb.lib().nb_sym_rows(n, world)
        q, gm = system(n, 2)
        # outgoing partial rows per owner: PJ[rows_global][S][3], exactly the buffer the kernel stores into
        out = [np.zeros((world * rows, S, 3)) for _ in range(world)]
        wrote = [np.zeros((world * rows, S), dtype=bool) for _ in range(world)]

        def send(owner, row, j0, j1, v):
            assert not wrote[owner][row, j0 - owner * S:j1 - owner * S].any(), "a partial row element is written once"
            out[owner][row, j0 - owner * S:j1 - owner * S] = v
            wrote[owner][row, j0 - owner * S:j1 - owner * S] = True

        pi, pj_ptr, pj_list = execute_rank(nb, n, world, rank, blocks, q, gm, send)
        # the reduce-scatter of the accelerations = every rank's stores into the owners' PJ (here: gathered through gloo)
        gathered = [None] * world
        dist.all_gather_object(gathered, (out, wrote))
        pj = np.zeros((world * rows, S, 3))
        have = np.zeros((world * rows, S), dtype=bool)
        for src in range(world):
            o, w = gathered[src]
            assert not (have & w[rank]).any(), "two ranks never write the same partial row element"
            pj[w[rank]] = o[rank][w[rank]]
            have |= w[rank]
        # integrate-side sum: PI slots of the row, then the PJ rows the plan lists, in order
        a = np.zeros((S, 3))
        for slot in sorted(pi):
            row0, acc = pi[slot]
            a[row0 - rank * S:row0 - rank * S + len(acc)] += acc
        for c in range(rows):
            stride = nb.lib().nb_sym_row_stride(n, world)
            lo, hi = c * stride, min(S, (c + 1) * stride)
            listed = list(pj_list[pj_ptr[c]:pj_ptr[c + 1]])
            for row in range(world * rows):
                # a listed PJ row may be written only in part (the shared block pair is cut at shard/2, not at a row
                # boundary): the rest stays zero, as in the zero-filled device buffer
                assert have[row, lo:hi].any() == (row in listed)
            for row in listed:
                a[lo:hi] += pj[row, lo:hi]
        parts = [None] * world
        dist.all_gather_object(parts, a)
        if rank == 0:
            ret["a"] = np.concatenate(parts)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world,blocks", [(4096, 2, 5), (6144, 2, 8), (10000, 2, 7)])
def test_executed_plan_two_ranks_gloo(nb, n, world, blocks):
    port = 29600 + (os.getpid() % 300)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, blocks, ret), nprocs=world, join=True)
        a = ret["a"]
    q, gm = system(n, 2)
    ref = direct(q, gm)
    assert np.max(np.abs(a - ref) / np.abs(ref).max(axis=1, keepdims=True)) < 1e-12
