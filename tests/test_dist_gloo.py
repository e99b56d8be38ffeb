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


# ---- temporally blocked kernels on slabs: L-line halo + redundant recomputation (kl_chain_tma.cuh, MG = true) ----
def _exchange_lines(X, L, rank, world):
    """first / last L lines of the slab <-> neighbours (kl_ops.cu halo_exchange_lines); zeros at the global boundary"""
    nx = X.shape[1]
    t_lo, t_hi = torch.zeros(L, nx, dtype=torch.float64), torch.zeros(L, nx, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(torch.from_numpy(X[:L].copy()), rank - 1), dist.irecv(t_lo, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(torch.from_numpy(X[-L:].copy()), rank + 1), dist.irecv(t_hi, rank + 1)]
    for r in reqs:
        r.wait()
    return t_lo.numpy(), t_hi.numpy()


def _apply_rows(U):
    """5-point operator on the interior lines of U (lines 1..-2), zero Dirichlet left/right; poisson.f90:42 order"""
    P = np.zeros((U.shape[0], U.shape[1] + 2))
    P[:, 1:-1] = U
    s = ((P[1:-1, :-2] + P[1:-1, 2:]) + P[2:, 1:-1]) + P[:-2, 1:-1]
    return 4.0 * U[1:-1] - s


def _slab_cheb_chain(r_loc, nx, nyl, k, params, rank, world):
    """degree-k Chebyshev on a slab in ONE pass: exchange k lines of r, then every level recomputes the
    neighbours' lines it still needs (level l is valid on lines [-k+l, nyl+k-l))."""
    ea, eb = params
    theta, delta = (eb + ea) / 2.0, abs(eb - ea) / 2.0
    sigma = theta / delta
    rho_prev = 1.0 / sigma
    R = r_loc.reshape(nyl, nx)
    lo, hi = _exchange_lines(R, k, rank, world)
    Rext = np.vstack([lo, R, hi])                    # lines -k .. nyl+k-1
    # lines that do not exist (beyond the global boundary) are not unknowns: their values stay zero
    exists = np.ones(nyl + 2 * k, dtype=bool)
    if rank == 0:
        exists[:k] = False
    if rank == world - 1:
        exists[-k:] = False
    d = Rext / theta
    z = d.copy()
    for _ in range(k):
        rho = 1.0 / (2.0 * sigma - rho_prev)
        c1, c2 = rho * rho_prev, 2.0 * rho / delta
        az = np.zeros_like(z)
        az[1:-1] = _apply_rows(z)
        # fma(c1, d, c2*(r - az)): numpy has no fma; emulate with exact-product splitting is overkill here --
        # the oracle comparison below therefore uses a tolerance of a few ulp instead of bit equality
        d = c1 * d + c2 * (Rext - az)
        z = z + d
        z[~exists] = 0.0
        d[~exists] = 0.0
        rho_prev = rho
    return z[k:k + nyl].reshape(-1)


def _worker_chain(rank, world, port, nx, ny, ks, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmres_b200.dist import slab_partition

    j0, nyl = slab_partition(ny, rank, world)
    rng = np.random.default_rng(5)
    rg = rng.standard_normal(nx * ny)                 # same global vector on both ranks
    r_loc = rg.reshape(ny, nx)[j0:j0 + nyl].reshape(-1).copy()
    res = {}
    for k in ks:
        res[k] = _slab_cheb_chain(r_loc, nx, nyl, k, (0.2, 8.2), rank, world)
    out[rank] = dict(j0=j0, nyl=nyl, z=res)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_chebyshev_chain_world2_matches_oracle(ko):
    """The multi-GPU form of the temporally blocked Chebyshev kernel: k-line halo, neighbours' lines recomputed
    redundantly level by level.  Must reproduce the single-process oracle (ko_cheb) on the global grid."""
    ns, world, ks = 48, 2, (1, 2, 4, 6)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_chain, args=(world, _free_port(), ns, ns, ks, out), nprocs=world, join=True)
    rg = np.random.default_rng(5).standard_normal(ns * ns)
    for k in ks:
        z = np.concatenate([out[r]["z"][k] for r in range(world)])
        zo = ko.apply_precond(ko.cheb_fn(k), ko.stvec_fn(), rg, (0.2, 8.2), ns)
        # numpy evaluates c1*d + c2*(...) with two roundings where the oracle and the CUDA kernel use one fma
        assert np.allclose(z, zo, rtol=1e-13, atol=1e-15), (k, np.abs(z - zo).max())


# ---- "producer pushes" halo scheme of the fused multi-GPU CG (kl_cg.cu FCgDirX2 / FCgRUpdate, kCbPush slots) ------
def _worker_push(rank, world, port, ns, out):
    """Mirror of the multi-GPU plain-CG iteration of libkrylov_b200: the operator is applied twice per iteration
    (A p is never stored), the x update is deferred by one iteration, and there is NO halo exchange step: the
    kernel that produces p_new / r_new stores its first and last line into the neighbours' slot
    [parity = iteration & 1][vector][direction]; consumers read parity (iteration - 1) & 1; the all-reduce that
    ends each kernel is the only synchronisation."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmres_b200.dist import slab_partition

    nx = ns
    j0, nyl = slab_partition(ns, rank, world)
    slots = np.zeros((2, 2, 2, nx))        # [parity][0 = r, 1 = p][0 = lo (from rank-1), 1 = hi (from rank+1)]
    R_, P_ = 0, 1

    def push(par, vec, first, last):
        """my first line -> rank-1's hi slot ; my last line -> rank+1's lo slot (comm_push_send / _recv)"""
        reqs = []
        t_lo, t_hi = torch.zeros(nx, dtype=torch.float64), torch.zeros(nx, dtype=torch.float64)
        if rank > 0:
            reqs += [dist.isend(torch.from_numpy(first.copy()), rank - 1), dist.irecv(t_lo, rank - 1)]
        if rank < world - 1:
            reqs += [dist.isend(torch.from_numpy(last.copy()), rank + 1), dist.irecv(t_hi, rank + 1)]
        for q in reqs:
            q.wait()
        if rank > 0:
            slots[par, vec, 0] = t_lo.numpy()
        if rank < world - 1:
            slots[par, vec, 1] = t_hi.numpy()

    def apply_with_halo(U, lo, hi):
        Pd = np.zeros((nyl + 2, nx + 2))
        Pd[1:-1, 1:-1] = U
        if rank > 0:
            Pd[0, 1:-1] = lo
        if rank < world - 1:
            Pd[-1, 1:-1] = hi
        s = ((Pd[1:-1, :-2] + Pd[1:-1, 2:]) + Pd[2:, 1:-1]) + Pd[:-2, 1:-1]
        return 4.0 * U - s

    ones = np.ones((nyl, nx))
    push(0, R_, ones[0], ones[-1])
    b = apply_with_halo(ones, slots[0, R_, 0], slots[0, R_, 1])        # b = A*1 on the slab
    slots[:] = 0.0
    r, p, x = b.copy(), np.zeros_like(b), np.zeros_like(b)
    push(0, R_, r[0], r[-1])                                           # set-up push (comm_push_lines): r0 and p0 = 0
    push(0, P_, p[0], p[-1])
    rr = _allsum(float(np.sum(r * r)))                                 # ... ordered by this all-reduce
    beta, alpha, hist, it_conv = 0.0, 0.0, [], -1
    for it in range(1, 5000):
        pin, pout = (it - 1) & 1, it & 1
        # K1: x += alpha_prev p ; p' = r + beta p ; (A p').p' ; push p' lines
        x += alpha * p
        pn = r + beta * p
        pn_lo = slots[pin, R_, 0] + beta * slots[pin, P_, 0]
        pn_hi = slots[pin, R_, 1] + beta * slots[pin, P_, 1]
        ap = apply_with_halo(pn, pn_lo, pn_hi)
        push(pout, P_, pn[0], pn[-1])
        alpha = rr / _allsum(float(np.sum(ap * pn)))
        # K2: p' and A p' recomputed from the SAME (unchanged) halo lines ; r' = r - alpha A p' ; push r' lines
        ap2 = apply_with_halo(r + beta * p, pn_lo, pn_hi)
        assert np.array_equal(ap2, ap)
        rnew = r - alpha * ap2
        push(pout, R_, rnew[0], rnew[-1])
        rn = _allsum(float(np.sum(rnew * rnew)))
        beta, rr = rn / rr, rn
        r, p = rnew, pn
        hist.append(np.sqrt(rn))
        if hist[-1] < 1e-9:
            it_conv = it
            break
    x += alpha * p                                                     # flush of the deferred update
    out[rank] = dict(j0=j0, nyl=nyl, b=b.reshape(-1), x=x.reshape(-1), it=it_conv, hist=np.array(hist))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pushed_halo_cg_matches_oracle(ko, world):
    ns = 57
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_push, args=(world, _free_port(), ns, out), nprocs=world, join=True)
    A = ko.stvec_fn()
    b = ko.manufactured_rhs(A, ns)
    o = ko.cg_omp(A, b, 1e-9, 5000)
    parts = [out[r] for r in range(world)]
    assert np.array_equal(np.concatenate([p["b"] for p in parts]), b)
    assert all(p["it"] == o.iter for p in parts)
    assert np.allclose(parts[0]["hist"], o.history, rtol=1e-7)
    assert np.abs(np.concatenate([p["x"] for p in parts]) - o.x).max() < 1e-10
