"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed
for the plumbing).

The reference parallelises over wavelengths with a work-stealing thread pool and
per-thread Gamma copies that are summed afterwards
(Source/SimdFullIterationTemplates.hpp:675-711, Source/ThreadStorage.cpp:150-166,
:343-396).  Here:

* 1D atmospheres are LAMBDA-SHARDED: every rank sweeps a contiguous,
  cost-balanced wavelength range, leaves its [Gamma | Rij,Rji] partial sums
  un-finalised in one packed device buffer, and ONE all-reduce (sum) of that
  buffer plus a tiny max-reduce for dJ per iteration replaces the host-thread
  reduction; finalise_Gamma and stat_eq then run replicated on every rank.  J
  rows are disjoint per shard and only gathered when asked for.
* 1.5D column stacks are COLUMN-SHARDED: columns never interact, so there is no
  data-path collective at all, only a final gather.

The partitioning functions and the reduce/finalise protocol are backend
agnostic (any object with the small `ShardBackend` surface), which is how the
world_size-2 gloo tests exercise them on CPU.
"""
import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import capi


# ------------------------------------------------------------- partitioning
def lambda_costs(problem, c_ray=1.0, c_line=0.35, c_trans=0.08) -> np.ndarray:
    """Relative cost of each wavelength: per ray a fixed solver cost plus a term
    per active line; per wavelength a term per active transition (epilogue)."""
    L, M = problem.Nspect, problem.Nrays
    nline = np.zeros(L)
    ntrans = np.zeros(L)
    for a in problem.atoms:
        for t in a.trans:
            ntrans[t.Nblue:t.Nred] += 1
            if t.type == capi.LINE:
                nline[t.Nblue:t.Nred] += 1
    return 2 * M * (c_ray + c_line * nline) + 2 * M * c_trans * ntrans


def partition_balanced(costs: np.ndarray, nshards: int) -> List[Tuple[int, int]]:
    """Contiguous ranges [lo, hi) of near-equal total cost (prefix-sum split).
    Every shard is non-empty when len(costs) >= nshards."""
    n = len(costs)
    if nshards < 1:
        raise ValueError('nshards must be >= 1')
    if n < nshards:
        raise ValueError(f'cannot cut {n} items into {nshards} non-empty shards')
    csum = np.concatenate(([0.0], np.cumsum(np.asarray(costs, dtype=np.float64))))
    bounds = [0]
    for s in range(1, nshards):
        target = csum[-1] * s / nshards
        b = int(np.searchsorted(csum, target, side='left'))
        # pick the closer of the two neighbouring cut points
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, n)] - target):
            b -= 1
        b = max(b, bounds[-1] + 1)
        b = min(b, n - (nshards - s))
        bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(nshards)]


def partition_wavelengths(problem, nshards: int) -> List[Tuple[int, int]]:
    return partition_balanced(lambda_costs(problem), nshards)


def partition_columns(ncol: int, nshards: int) -> List[Tuple[int, int]]:
    return partition_balanced(np.ones(ncol), nshards)


# ------------------------------------------------- raw device memory <-> torch
class _CudaArray:
    def __init__(self, ptr, nelem, typestr='<f8'):
        self.__cuda_array_interface__ = {'shape': (int(nelem),), 'typestr': typestr,
                                         'data': (int(ptr), False), 'version': 2}


def device_tensor(ptr, nbytes, device=0):
    """A float64 torch view (no copy) of a device buffer owned by the C library."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, nbytes // 8), device=torch.device('cuda', device))


def copy_device_to_host(host: np.ndarray, ptr, nbytes, device=0):
    import torch
    host[...] = device_tensor(ptr, nbytes, device).cpu().numpy().reshape(host.shape)


def copy_host_to_device(ptr, host: np.ndarray, nbytes, device=0):
    import torch
    device_tensor(ptr, nbytes, device).copy_(torch.from_numpy(np.ascontiguousarray(host).reshape(-1)))


# --------------------------------------------------------------- protocols
class ShardBackend:
    """What a lambda shard must provide.  `lightweaver_b200.context.Context`
    satisfies it on the GPU; the CPU tests use an oracle-backed stand-in."""

    def partial_iteration(self, lambdaIterate: bool):
        """Sweep this shard's wavelengths, leave partial sums un-finalised."""
        raise NotImplementedError

    def accum_tensor(self):
        """torch tensor (device or CPU) viewing the packed partial-sum buffer."""
        raise NotImplementedError

    def finalise(self):
        raise NotImplementedError

    def local_dj(self) -> Tuple[float, int]:
        raise NotImplementedError

    def stat_eq(self):
        raise NotImplementedError

    def j_tensor(self):
        """torch tensor [Nspect, Nspace] viewing this rank's copy of J (1D atmospheres)."""
        raise NotImplementedError

    def prd_redistribute(self, maxIter: int, tol: float):
        """Angle-averaged PRD sub-iterations over the whole spectrum; returns the number taken."""
        raise NotImplementedError


class GpuLambdaShard(ShardBackend):
    def __init__(self, ctx):
        self.ctx = ctx
        ptr, nbytes = ctx.device_buffer(capi.BUF_ACCUM)
        self._accum = device_tensor(ptr, nbytes, ctx.device)

    def partial_iteration(self, lambdaIterate=False, asyncDJ=False):
        """asyncDJ: also reduce dJ over this shard's wavelengths on the device (BUF_DJMAX; no host sync)."""
        self.ctx.fs_iter_device(lambdaIterate=lambdaIterate, deferFinalise=True, want_dJ=False, asyncDJ=asyncDJ)

    def djmax_tensor(self):
        ptr, nbytes = self.ctx.device_buffer(capi.BUF_DJMAX)
        return device_tensor(ptr, nbytes, self.ctx.device)

    def accum_tensor(self):
        return self._accum

    def finalise(self):
        self.ctx.finalise()

    def local_dj(self):
        return self.ctx.dj_max()

    def stat_eq(self):
        self.ctx.stat_eq_device()

    def j_tensor(self):
        ptr, nbytes = self.ctx.device_buffer(capi.BUF_J)
        p = self.ctx.problem
        return device_tensor(ptr, nbytes, self.ctx.device).view(p.Ncol * p.Nspect, p.Nspace)

    def prd_redistribute(self, maxIter=3, tol=1e-2):
        return self.ctx.prd_redistribute_device(maxIter=maxIter, tol=tol)


def reduce_dj(dJ: float, idx: int, group=None, device='cpu'):
    """(max, wavelength index of the max) across ranks: MAX on the value, then
    MIN over the indices of the ranks that hold it."""
    import torch
    import torch.distributed as dist
    v = torch.tensor([dJ], dtype=torch.float64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
    big = np.iinfo(np.int64).max
    i = torch.tensor([idx if dJ == v.item() else big], dtype=torch.int64, device=device)
    dist.all_reduce(i, op=dist.ReduceOp.MIN, group=group)
    return v.item(), int(i.item())


def sharded_gamma_iteration(shard: ShardBackend, lambdaIterate=False, group=None, want_dJ=True):
    """One lambda-sharded Gamma iteration: local sweep -> all-reduce(sum) of the
    packed [Gamma | R] partial sums -> replicated finalise.  Returns (dJMax, idx)
    over all shards when want_dJ."""
    import torch.distributed as dist
    shard.partial_iteration(lambdaIterate)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(shard.accum_tensor(), op=dist.ReduceOp.SUM, group=group)
    shard.finalise()
    if not want_dJ:
        return None
    dJ, idx = shard.local_dj()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = shard.accum_tensor()
        dJ, idx = reduce_dj(dJ, idx, group, device=t.device)
    return dJ, idx


class GraphedIteration:
    """One whole Gamma iteration of a shard -- zero the partial sums, continuum / ray / Gamma kernels
    (forked over the library's side streams), the all-reduce of the packed [Gamma | R] buffer when there
    is more than one rank, finalise, the statistical-equilibrium solve -- captured ONCE into a CUDA graph
    and replayed: a 1D atmosphere is ~15 launches of 10-100 us each, so a wavelength shard on 8 GPUs is
    bound by launch latency and by the host dispatch of the collective, not by its kernels.
    dJ is reduced on the device inside the graph (and max-reduced across ranks): ``dJMax()`` reads it."""

    def __init__(self, shard, group=None, lambdaIterate=False, with_stat_eq=True, warmup=3):
        import torch
        import torch.distributed as dist
        self.shard = shard
        ctx = shard.ctx
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        accum = shard.accum_tensor()
        djmax = shard.djmax_tensor()

        def body():
            shard.partial_iteration(lambdaIterate, asyncDJ=True)
            if multi:
                dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(djmax, op=dist.ReduceOp.MAX, group=group)
            shard.finalise()
            if with_stat_eq:
                ctx.stat_eq_device(wait=False)
        outer = torch.cuda.current_stream()
        # the library launches on the stream it was given: warm up and capture on a side stream of ours
        side = torch.cuda.Stream()
        side.wait_stream(outer)
        with torch.cuda.stream(side):
            ctx.set_stream(side)
            for _ in range(warmup):   # (first calls allocate work lists and opt in to shared memory sizes)
                body()
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                body()
        ctx.set_stream(outer)
        outer.wait_stream(side)

    def replay(self):
        self.graph.replay()

    def dJMax(self):
        """The (all-rank) dJ of the last replay; synchronises."""
        return float(self.shard.djmax_tensor().cpu()[0])


def sharded_prd_redistribute(shard: ShardBackend, ranges: Sequence[Tuple[int, int]], rank: int, maxIter=3,
                             tol=1e-2, group=None):
    """PRD redistribution after a lambda-sharded Gamma iteration (1D atmospheres).  The scattering
    integral of a PRD line needs J over the whole line, and the rates it uses are already identical on
    every rank (they came out of the all-reduce), so: ONE all-gather of the J rows, then every rank
    runs the redistribution and its formal solution over the PRD wavelengths REPLICATED -- they are a
    small subset of the spectrum, and no further exchange is needed.  Afterwards rho, the PRD lines'
    rates and J at the PRD wavelengths are up to date on every rank."""
    J = shard.j_tensor()
    lo, hi = ranges[rank]
    full = gather_rows(J[lo:hi].clone(), ranges, group)
    if full is not J:
        J[:full.shape[0]].copy_(full)
    return shard.prd_redistribute(maxIter, tol)


def gather_rows(local_rows, ranges: Sequence[Tuple[int, int]], group=None):
    """All-gather wavelength rows (e.g. J[lo:hi]) of unequal length from every
    shard; returns the concatenation in shard order as a torch tensor."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows
    width = local_rows.shape[1:]
    maxlen = max(hi - lo for lo, hi in ranges)
    pad = torch.zeros((maxlen,) + tuple(width), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    outs = [torch.empty_like(pad) for _ in ranges]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, ranges)], dim=0)
