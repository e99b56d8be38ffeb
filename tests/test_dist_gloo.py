"""CPU, world_size = 2, gloo: the row-slab decomposition the multi-GPU path uses.

Each rank owns the lines kl_partition_rank gives it, exchanges one halo line per operator
apply (send/recv with the neighbour ranks) and all-reduces every inner product -- the same
message pattern libkrylov_b200 issues over NCCL (kl_core.cu comm_halo_exchange /
comm_allreduce).  The slab solve must reproduce the single-process oracle (cg_omp)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_apply(x_loc, nx, nyl, rank, world):
    """y = A x on a slab with halo lines exchanged with the neighbours (zero at the global boundary)."""
    X = x_loc.reshape(nyl, nx)
    lo = np.zeros(nx)
    hi = np.zeros(nx)
    reqs = []
    t_lo, t_hi = torch.zeros(nx, dtype=torch.float64), torch.zeros(nx, dtype=torch.float64)
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(X[0].copy()), rank - 1))
        reqs.append(dist.irecv(t_lo, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(X[-1].copy()), rank + 1))
        reqs.append(dist.irecv(t_hi, rank + 1))
    for r in reqs:
        r.wait()
    if rank > 0:
        lo = t_lo.numpy()
    if rank < world - 1:
        hi = t_hi.numpy()
    P = np.zeros((nyl + 2, nx + 2))
    P[1:-1, 1:-1] = X
    P[0, 1:-1] = lo
    P[-1, 1:-1] = hi
    # poisson.f90:42 order: ((l + r) + x(idx+n)) + x(idx-n)
    s = ((P[1:-1, :-2] + P[1:-1, 2:]) + P[2:, 1:-1]) + P[:-2, 1:-1]
    return (4.0 * X - s).reshape(-1)


def _allsum(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


def _worker(rank, world, port, ns, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmres_b200.dist import slab_partition, local_slab

    j0, nyl = slab_partition(ns, rank, world)
    # b = A*1 built slab-wise (exercises the halo exchange) must equal the global rhs
    b = _slab_apply(np.ones(ns * nyl), ns, nyl, rank, world)
    # cg_omp (cg.f90:83-152) on the slab
    x = np.zeros_like(b)
    r = b.copy()
    p = r.copy()
    hist = []
    it = -1
    for i in range(1, 5000):
        ax = _slab_apply(p, ns, nyl, rank, world)
        rr = _allsum(r @ r)
        alpha = rr / _allsum(ax @ p)
        x += alpha * p
        r -= alpha * ax
        rn = _allsum(r @ r)
        res = np.sqrt(rn)
        hist.append(res)
        p = r + (rn / rr) * p
        if res < 1e-9:
            it = i
            break
    out[rank] = dict(j0=j0, nyl=nyl, b=b, x=x, it=it, hist=np.array(hist))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ns", [60, 101])
def test_slab_cg_world2_matches_oracle(ko, ns):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ns, out), nprocs=world, join=True)
    A = ko.stvec_fn()
    b = ko.manufactured_rhs(A, ns)
    o = ko.cg_omp(A, b, 1e-9, 5000)
    parts = [out[r] for r in range(world)]
    assert parts[0]["j0"] == 0 and parts[1]["j0"] == parts[0]["nyl"]
    assert sum(p["nyl"] for p in parts) == ns
    assert np.array_equal(np.concatenate([p["b"] for p in parts]), b)
    assert parts[0]["it"] == parts[1]["it"] == o.iter
    assert np.allclose(parts[0]["hist"], o.history, rtol=1e-7)
    x = np.concatenate([p["x"] for p in parts])
    assert np.abs(x - o.x).max() < 1e-10


def test_partition_rule():
    sys.path.insert(0, ROOT)
    from gmres_b200.dist import slab_partition
    for ny, P in ((16384, 8), (10, 4), (7, 7), (4097, 3)):
        parts = [slab_partition(ny, r, P) for r in range(P)]
        assert parts[0][0] == 0
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]
        assert parts[-1][0] + parts[-1][1] == ny
        assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1
    with pytest.raises(ValueError):
        slab_partition(10, 4, 4)
