"""Host-side multi-GPU logic on CPU: wavelength / column partitioning and the
lambda-sharded reduce-then-finalise protocol, exercised with world_size-2 gloo
process groups.  The shard back end here is the C oracle standing in for the GPU
(tests only); the protocol code under test is lightweaver_b200.sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightweaver_b200 import sharding, synth
from oracle import oraclelib
from tests.util import gamma_err, rel_err


def test_partition_balanced_properties():
    rng = np.random.default_rng(0)
    for n, s in ((10, 1), (10, 10), (1096, 8), (57, 4), (10056, 8)):
        costs = rng.uniform(1.0, 5.0, n)
        parts = sharding.partition_balanced(costs, s)
        assert parts[0][0] == 0 and parts[-1][1] == n and len(parts) == s
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(hi > lo for lo, hi in parts)
        loads = [costs[lo:hi].sum() for lo, hi in parts]
        if n >= 8 * s:
            assert max(loads) <= 1.25 * costs.sum() / s
    with pytest.raises(ValueError):
        sharding.partition_balanced(np.ones(3), 4)
    assert sharding.partition_columns(4096, 8) == [(i * 512, (i + 1) * 512) for i in range(8)]


def test_lambda_costs_follow_active_lines():
    p = synth.tiny_problem()
    c = sharding.lambda_costs(p)
    assert c.shape == (p.Nspect,)
    line = p.atoms[0].trans[0]
    assert c[(line.Nblue + line.Nred) // 2] > c.min()
    parts = sharding.partition_wavelengths(p, 3)
    assert parts[0][0] == 0 and parts[-1][1] == p.Nspect


class OracleShard(sharding.ShardBackend):
    """CPU stand-in for GpuLambdaShard: partial sums from the C oracle."""

    def __init__(self, problem, laRange):
        self.p = problem
        self.lo, self.hi = laRange
        self.ctx = oraclelib.OracleContext(problem)
        self.prefill = [a.C.copy() for a in problem.active_atoms()]
        n = sum(a.Gamma.size for a in problem.active_atoms())
        n += sum(2 * t.Rij.size for a in problem.atoms for t in a.trans)
        self.accum = torch.zeros(n, dtype=torch.float64)
        self.dj = (0.0, 0)

    def partial_iteration(self, lambdaIterate=False):
        for a in self.p.active_atoms():
            a.Gamma[:] = 0.0
        self.dj = self.ctx.fs_iter(lambdaIterate=lambdaIterate, laStart=self.lo, laEnd=self.hi)
        parts = [a.Gamma.reshape(-1) for a in self.p.active_atoms()]
        parts += [x.reshape(-1) for a in self.p.atoms for t in a.trans for x in (t.Rij, t.Rji)]
        self.accum.copy_(torch.from_numpy(np.concatenate(parts)))

    def accum_tensor(self):
        return self.accum

    def finalise(self):
        buf = self.accum.numpy()
        off = 0
        for a, C in zip(self.p.active_atoms(), self.prefill):
            G = buf[off:off + a.Gamma.size].reshape(a.Gamma.shape) + C
            off += a.Gamma.size
            N = a.Nlevel
            for i in range(N):
                G[:, i, i] = 0.0
                G[:, i, i] = -G[:, :, i].sum(axis=1)
            a.Gamma[:] = G
        for a in self.p.atoms:
            for t in a.trans:
                t.Rij[:] = buf[off:off + t.Rij.size].reshape(t.Rij.shape)
                off += t.Rij.size
                t.Rji[:] = buf[off:off + t.Rji.size].reshape(t.Rji.shape)
                off += t.Rji.size

    def local_dj(self):
        return self.dj

    def stat_eq(self):
        self.ctx.stat_eq()

    def j_tensor(self):
        return torch.from_numpy(self.p.J[0])

    def prd_redistribute(self, maxIter=3, tol=1e-2):
        nl = sum(1 for a in self.p.atoms for t in a.trans if t.rhoPrd is not None)
        return self.ctx.redistribute_prd(maxIter=maxIter, tol=tol, nlines=nl)['nIter']


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    p = synth.tiny_problem()
    ranges = sharding.partition_wavelengths(p, world)
    shard = OracleShard(p, ranges[rank])
    res = []
    for it in range(2):
        dJ, idx = sharding.sharded_gamma_iteration(shard, lambdaIterate=(it == 0))
        shard.stat_eq()
        res.append((dJ, idx))
    lo, hi = ranges[rank]
    J = sharding.gather_rows(torch.from_numpy(p.J[0, lo:hi].copy()), ranges)
    if rank == 0:
        np.savez(out, Gamma=p.atoms[0].Gamma, n=p.atoms[0].n, J=J.numpy(), dJ=np.array(res),
                 R=p.atoms[0].trans[0].Rij)
    dist.barrier()
    dist.destroy_process_group()


def test_lambda_sharded_iteration_over_gloo(tmp_path):
    out = str(tmp_path / 'rank0.npz')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process reference
    q = synth.tiny_problem()
    o = oraclelib.OracleContext(q)
    ref = []
    for it in range(2):
        q.prefill_gamma()
        ref.append(o.fs_iter(lambdaIterate=(it == 0)))
        o.stat_eq()
    assert gamma_err(got['Gamma'], q.atoms[0].Gamma) <= 1e-12
    assert rel_err(got['n'], q.atoms[0].n) <= 1e-10
    assert rel_err(got['J'], q.J[0]) <= 1e-10  # second iteration inherits the re-associated Gamma sums
    assert rel_err(got['R'], q.atoms[0].trans[0].Rij, floor=1e-30) <= 1e-12
    for (dJ, idx), (rdJ, ridx) in zip(got['dJ'], ref):
        assert abs(dJ - rdJ) <= 1e-12 * max(rdJ, 1.0) and int(idx) == ridx


def _prd_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    p = synth.tiny_prd_problem()
    ranges = sharding.partition_wavelengths(p, world)
    shard = OracleShard(p, ranges[rank])
    for it in range(2):
        sharding.sharded_gamma_iteration(shard, lambdaIterate=(it == 0), want_dJ=False)
        n = sharding.sharded_prd_redistribute(shard, ranges, rank, maxIter=2, tol=1e-6)
        assert n == 2
        shard.stat_eq()
    lines = [t for t in p.atoms[0].trans if t.rhoPrd is not None]
    np.savez(out % rank, rho=lines[0].rhoPrd, n=p.atoms[0].n, Rij=lines[0].Rij,
             Jprd=p.J[0, lines[0].Nblue:lines[0].Nred])
    dist.barrier()
    dist.destroy_process_group()


def test_lambda_sharded_prd_over_gloo(tmp_path):
    """PRD after a lambda-sharded iteration: J all-gathered, redistribution replicated; every rank ends
    with the state of the single-process run."""
    out = str(tmp_path / 'rank%d.npz')
    port = _free_port()
    mp.spawn(_prd_worker, args=(2, port, out), nprocs=2, join=True)
    q = synth.tiny_prd_problem()
    o = oraclelib.OracleContext(q)
    for it in range(2):
        q.prefill_gamma()
        o.fs_iter(lambdaIterate=(it == 0))
        o.redistribute_prd(maxIter=2, tol=1e-6, nlines=2)
        o.stat_eq()
    line = [t for t in q.atoms[0].trans if t.rhoPrd is not None][0]
    for rank in range(2):
        got = np.load(out % rank)
        assert rel_err(got['rho'], line.rhoPrd) <= 1e-9
        assert rel_err(got['n'], q.atoms[0].n) <= 1e-9
        assert rel_err(got['Rij'], line.Rij, floor=1e-30) <= 1e-9
        assert rel_err(got['Jprd'], q.J[0, line.Nblue:line.Nred]) <= 1e-9

