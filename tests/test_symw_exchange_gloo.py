"""The two NCCL exchanges of a DSOFT_SYM_W plan (`_SymW.exchange_forward / exchange_backward`, loss.py), run for real
over gloo at world sizes 2, 3, 4 and 8: send / receive lists pair up on every rank (no hang), the contested half block
of even world sizes lands where the ownership rule of include/dsoft.h says, and every received block is ADDED to the
right rows.  The layout numbers are what dsoft_plan_symw_info reports for (world, rank, b); the buffers hold values
that encode (sender, primed block, row, column), so the expected sums are computed in closed form."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def plan_ncols(W, r, b):
    """Primed columns rank r computes (dsoft_plan_create: s_ncols of a sym_w plan)."""
    nfull = (W + 1) // 2 if W % 2 else W // 2
    if W % 2:
        return nfull * b
    return (nfull + 1) * b if r < W // 2 else nfull * b + b // 2


def layout(W, r, b, Dz, Dx):
    ncols = plan_ncols(W, r, b)
    Bcol = W * b + 64
    off_colsum = 128
    fwd_numel = off_colsum + 6 * Bcol
    off_r3 = 64
    off_r4 = off_r3 + (ncols - b) * Dz
    off_a3 = off_r4 + (ncols - b) * Dx + 32
    off_a4 = off_a3 + 2 * b * Dz  # two K splits: the received products go to split 0
    numel = off_a4 + 2 * b * Dx
    info = [1, b, Bcol, off_colsum, ncols, off_r3, Dz, off_r4, Dx, off_a3, off_a4, 2]
    return info, fwd_numel, numel


def colsum_value(s, j):  # rank s's column sum of primed column j, quantity q adds 1000 q
    return 10.0 * s + 0.001 * j


def remote_value(s, which, row, col):  # rank s's transposed product, row of the remote buffer
    return (1 + which) * (100.0 * s + 0.01 * row) + 1e-4 * col


def _worker(rank, W, port, b, Dz, Dx, ret, members=None):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=W if members is None else max(members) + 1)
    torch.set_num_threads(1)
    from dinosoft_b200.loss import _SymW

    group = None
    if members is not None:  # the loss lives in a sub-group: ranks and peers are group-relative
        group = dist.new_group(members)
        if rank not in members:
            dist.barrier()
            dist.destroy_process_group()
            return
        rank = members.index(rank)

    info, fwd_numel, numel = layout(W, rank, b, Dz, Dx)
    sw = _SymW(info, W, rank)
    dev = torch.device("cpu")
    # ---- forward: column sums
    fscr = torch.full((fwd_numel,), -7.0)
    cs = sw.colsum(fscr)
    j = torch.arange(sw.Bcol, dtype=torch.float32)
    for q in range(6):
        cs[q] = colsum_value(rank, j) + 1000.0 * q
    order = []
    sw.run(dev, 1, order.append, lambda: (order.append("x"), sw.exchange_forward(fscr, group)))
    assert order == [4, 1, "x", 3, 2]
    # ---- backward: transposed products
    scr = torch.full((numel,), -3.0)
    for which, d in ((0, Dz), (1, Dx)):
        if not d:
            continue
        rem = sw.remote(scr, which)
        rows = torch.arange(rem.shape[0], dtype=torch.float32)[:, None]
        cols = torch.arange(d, dtype=torch.float32)[None, :]
        rem.copy_(remote_value(rank, which, rows, cols))
        sw.own(scr, which).fill_(0.5)
    sw.exchange_backward(scr, group)
    ret[rank] = dict(colsum=sw.colsum(fscr).clone(), own=[sw.own(scr, w).clone() for w in ((0, 1) if Dx else (0,))],
                     guard=(float(fscr[0]), float(scr[0])), sends=list(sw.sends), recvs=list(sw.recvs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W,b,Dx,members", [(2, 8, 4, None), (3, 8, 4, None), (4, 8, 0, None), (8, 4, 4, None),
                                            (2, 8, 4, [0, 2])])
def test_exchanges(W, b, Dx, members):
    Dz = 6
    mgr = mp.Manager()
    ret = mgr.dict()
    nprocs = W if members is None else max(members) + 1
    mp.spawn(_worker, args=(W, _free_port(), b, Dz, Dx, ret, members), nprocs=nprocs, join=True)
    # every unordered pair of row blocks is computed exactly once: rank r owns (r, r + k) for its primed blocks k
    owned = {}
    for r in range(W):
        ncols = plan_ncols(W, r, b)
        for k in range(1, -(-ncols // b)):
            rows = min(b, ncols - k * b)
            owned.setdefault(frozenset((r, (r + k) % W)), []).append((r, k, rows))
    for pair, who in owned.items():
        if len(who) == 1:
            assert who[0][2] == b, (pair, who)
        else:  # the contested block of an even world size: one rank sends all b rows, the other the first b / 2
            assert W % 2 == 0 and sorted(w[2] for w in who) == [b // 2, b], (pair, who)
    assert len(owned) == W * (W - 1) // 2
    for r in range(W):
        o = ret[r]
        assert o["guard"] == (-7.0, -3.0)
        want = torch.stack([colsum_value(r, torch.arange(b, dtype=torch.float32)) + 1000.0 * q for q in range(6)])
        for s in range(W):
            for peer, k, rows in ret[s]["sends"]:
                if peer == r:
                    jj = torch.arange(k * b, k * b + rows, dtype=torch.float32)
                    for q in range(6):
                        want[q, :rows] += colsum_value(s, jj) + 1000.0 * q
        assert torch.allclose(o["colsum"][:, :b], want, rtol=1e-6, atol=1e-4), (W, r)
        for which, d in ((0, Dz), (1, Dx)):
            if not d:
                continue
            want = torch.full((b, d), 0.5)
            cols = torch.arange(d, dtype=torch.float32)[None, :]
            for s in range(W):
                for peer, k, rows in ret[s]["sends"]:
                    if peer == r:
                        rr = torch.arange((k - 1) * b, (k - 1) * b + rows, dtype=torch.float32)[:, None]
                        want[:rows] += remote_value(s, which, rr, cols)
            got = o["own"][which]
            assert torch.allclose(got[:b], want, rtol=1e-6, atol=1e-4), (W, r, which)
            assert torch.all(got[b:] == 0.5)  # split 1 untouched
