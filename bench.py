#!/usr/bin/env python
"""bench.py -- throughput of the Gamma-iteration hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c1] [--columns C]

One "step" = one full Gamma iteration over one batch of synthetic FAL C input:
formal solution + J + Gamma/rate accumulation over every wavelength and ray
(`formal_sol_gamma_matrices`), Gamma finalisation, and the per-depth
statistical-equilibrium solve (`stat_equil`).  Metric (BASELINE.json):
depth*lambda*mu*column ray-depth points per second (both directions counted),
with Gamma-iterations/s alongside.

  value      device-resident inputs, CUDA events on the launch stream, max over ranks
  e2e        the same step through the public API (`Context.formal_sol_gamma_matrices`
             + `stat_equil`) with HOST buffers: per-step H2D of the arrays the host
             mutates between iterations and D2H of everything it reads back
  roofline   the formal-solution launch set of one Gamma iteration (continuum_kernel ->
             ray_kernel<NL> -> gamma_kernel; ray_kernel dominates): algorithmic bytes /
             its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the reference's own multithreaded SIMD CPU path on this box's host cores

Workloads: c2 (default; configs[1]: FAL C 1D, H + Ca II + Mg II + Na I + He I,
~1e4 wavelengths x 10 rays; lambda-sharded with one all-reduce of the packed
[Gamma|R] buffer per step when N > 1), c3 (configs[2]: stack of perturbed FAL C
columns, H + Ca II, 5 rays; column-sharded, no data-path collective), c1
(configs[0]).  `--impl reference` times the reference's CPU implementation
(oracle/_ref, else the C port) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lightweaver_b200 import capi, synth  # noqa: E402

L2_FLUSH_BYTES = 256 << 20


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'MEASURED_PEAKS.json (of measured)'
    except Exception:
        return 6650.0, 'B200_PROFILING.md fallback (of fallback)'


def build_workload(name, columns, rank=0, world=1, for_gpu=True):
    """Returns (problem, description).  For c3 each rank builds only its column shard."""
    if name == 'c1':
        return synth.config_c1(), 'FAL C 1D, H 6-level + Ca II 5+1-level, 5 rays (configs[0])'
    if name == 'c2':
        return synth.config_c2(), ('FAL C 1D, H + Ca II + Mg II + Na I + He I active, 10 rays '
                                   '(configs[1], synthetic atomic data)')
    if name == 'c4':
        return synth.config_c4(nl=2.0), ('FAL C 1D, H + Ca II + Mg II, Mg II h&k in angle-averaged PRD, 5 rays '
                                         '(configs[3]; each step = Gamma iteration + up to 3 PRD sub-iterations on '
                                         'fixed LTE populations -- no stat-eq, the synthetic atom is not meant to be '
                                         'iterated to convergence with PRD; points counted for the Gamma iteration only)')
    if name == 'c5':
        from lightweaver_b200.sharding import partition_columns
        c0, c1 = partition_columns(columns, world)[rank]
        p = synth.config_c5(ncol=columns, col_range=(c0, c1))
        return p, (f'1.5D magnetised stack of {columns} perturbed FAL C columns x 82 depths, Ca II with the 854.2 nm '
                   'line Zeeman-split and polarised, 5 rays (configs[4]; each step = one J-updating full-Stokes '
                   'formal solution of every wavelength, formal_sol_full_stokes(updateJ=True, upOnly=False))')
    if name == 'c3':
        from lightweaver_b200.sharding import partition_columns
        c0, c1 = partition_columns(columns, world)[rank]
        p = synth.config_c3(ncol=columns, col_range=(c0, c1), with_profiles=not for_gpu,
                            alloc_phi=not for_gpu)
        return p, f'1.5D stack of {columns} perturbed FAL C columns x 82 depths, H + Ca II, 5 rays (configs[2])'
    raise SystemExit(f'unknown workload {name}')


def clock_probe_repeats(total_ms, steps):
    """Untimed repeats of the step under which the SM clocks are sampled when the timed region
    (total_ms, already reduced over ranks) is too short for nvidia-smi: 0, or a multiple of 20 worth
    about 1.5 s.  A pure function of rank-independent inputs: every rank of a lambda-sharded run must
    repeat the step -- which contains an all-reduce -- the same number of times."""
    if total_ms >= 1200.0:
        return 0
    per_step_ms = max(total_ms / max(steps, 1), 1e-3)
    return 20 * int(min(max(1500.0 / per_step_ms, 20.0), 20000.0) // 20)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------ reference arm
def time_reference(problem, steps, warmup, budget_s=120.0):
    """The reference's CPU implementation of one step on this box's host cores.
    1D workloads: the reference's own wavelength threading (Nthreads = cores) in
    its fastest usable SIMD scheme.  Column stacks: one single-threaded reference
    Context per column, columns spread over the cores, on a bounded column sample
    scaled linearly.  Returns dict(value, ms_per_step, ...)."""
    from oracle import reflib, oraclelib
    cores = os.cpu_count() or 1
    pts_col = problem.points_per_iter() / problem.Ncol
    if reflib.available():
        schemes = reflib.usable_schemes()
        scheme = 'AVX2FMA' if 'AVX2FMA' in schemes else schemes[-1]
        nprd = sum(1 for a in problem.atoms for t_ in a.trans if t_.rhoPrd is not None)
        if problem.Ncol == 1 and nprd:
            # PRD workload: Gamma iteration + prd_redistribute + stat_eq, timed around the three calls
            ctx = reflib.RefContext(problem, scheme=scheme, Nthreads=cores)

            def one():
                problem.prefill_gamma()
                t0 = time.perf_counter()
                ctx.fs_iter()
                ctx.redistribute_prd(maxIter=3, tol=1e-2, nlines=nprd)
                return time.perf_counter() - t0
            probe = one()
            steps = max(1, min(steps, int(budget_s / max(probe, 1e-6)) - warmup))
            for _ in range(warmup):
                one()
            sec = float(np.median([one() for _ in range(steps)]))
            ctx.close()
            return {'kind': 'reference', 'cores': cores, 'scheme': f'mali_full_precond_{scheme}',
                    'sec_per_step': sec, 'value': pts_col / sec, 'steps': steps,
                    'sample': f'full workload, {steps} timed (Gamma iteration + prd_redistribute(3)), '
                              f'Nthreads={cores}'}
        if problem.Ncol == 1:
            problem.prefill_gamma()
            ctx = reflib.RefContext(problem, scheme=scheme, Nthreads=cores)
            # bound the run: one probe call sizes the number of timed calls
            t0 = time.perf_counter()
            ctx.time_fs_iter(0, 1, True)
            probe = time.perf_counter() - t0
            steps = max(1, min(steps, int(budget_s / max(probe, 1e-6)) - warmup))
            ts = ctx.time_fs_iter(warmup, steps, True)
            ctx.close()
            sec = float(np.median(ts))
            return {'kind': 'reference', 'cores': cores, 'scheme': ctx.scheme_name if False else f'mali_full_precond_{scheme}',
                    'sec_per_step': sec, 'value': pts_col / sec, 'steps': steps,
                    'sample': f'full workload, {steps} timed Gamma iterations + stat_eq, Nthreads={cores}'}
        # column stack: sample of columns, one Context per column, thread pool over columns
        from concurrent.futures import ThreadPoolExecutor
        is_stokes = problem.Quv is not None
        ncs = min(problem.Ncol, (1 if is_stokes else 4) * cores)
        problem.prefill_gamma()
        ctxs = [reflib.RefContext(problem, col=c, scheme=scheme, Nthreads=1) for c in range(ncs)]

        def one(c):
            if is_stokes:
                # the reference's full-Stokes formal solution is single-threaded (FormalStokes.cpp:707-711)
                return c.full_stokes(updateJ=True, upOnly=False)[0]
            return c.time_fs_iter(0, 1, True)[0]
        with ThreadPoolExecutor(cores) as ex:
            t0 = time.perf_counter()
            list(ex.map(one, ctxs))
            probe = time.perf_counter() - t0
            steps = max(1, min(steps, int(budget_s / max(probe, 1e-6))))
            ts = []
            for _ in range(steps):
                t0 = time.perf_counter()
                list(ex.map(one, ctxs))
                ts.append(time.perf_counter() - t0)
        for c in ctxs:
            c.close()
        sec = float(np.median(ts))
        return {'kind': 'reference', 'cores': cores, 'scheme': f'mali_full_precond_{scheme}',
                'sec_per_step': sec * problem.Ncol / ncs, 'value': pts_col * ncs / sec, 'steps': steps,
                'sample': f'{ncs} of {problem.Ncol} columns (one reference Context per column over {cores} '
                          f'threads), scaled linearly'}
    # no compiled reference on this box: the C port, columns / single column over pthreads
    o = oraclelib.OracleContext(problem)
    ncs = min(problem.Ncol, 2 * cores)
    problem.prefill_gamma()
    t0 = time.perf_counter()
    o.fs_iter_columns(0, ncs, withStatEq=True, nthreads=min(cores, ncs))
    sec = time.perf_counter() - t0
    return {'kind': 'port', 'cores': min(cores, ncs), 'scheme': 'lw_oracle.c (scalar)',
            'sec_per_step': sec * problem.Ncol / ncs, 'value': pts_col * ncs / sec, 'steps': 1,
            'sample': f'{ncs} of {problem.Ncol} columns through the scalar C port'}


# ------------------------------------------------------------------------ ours
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from lightweaver_b200.context import Context
    from lightweaver_b200 import sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    problem, desc = build_workload(args.workload, args.columns, rank, world)
    column_sharded = args.workload in ('c3', 'c5')
    stokes = args.workload == 'c5'
    laRange = None
    if world > 1 and not column_sharded:
        laRange = sharding.partition_wavelengths(problem, world)[rank]
    stream = torch.cuda.current_stream()
    ctx = Context(problem, device=local_rank, stream=stream, laRange=laRange, upload=False)
    if stokes:
        ctx.upload(capi.ALL_INPUTS | capi.STOKES)
    elif column_sharded:
        ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
        ctx.update_deps(background=False, profiles_on_device=True)
    else:
        ctx.upload(capi.ALL_INPUTS | (capi.PRD if args.workload == 'c4' else 0))
    ctx.sync()
    with_prd = args.workload == 'c4'
    if with_prd and world > 1:
        raise SystemExit('bench.py: the PRD workload (c4) runs on one GPU')
    shard = sharding.GpuLambdaShard(ctx)
    pts_local, alg_bytes, _ = ctx.work_stats()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def step():
        if stokes:
            capi.check(ctx.lib.lwb200_formal_sol_full_stokes(ctx._h, 1, 0, None, None))
            return
        if world > 1 and not column_sharded:
            sharding.sharded_gamma_iteration(shard, group=None, want_dJ=False)
        else:
            # dJ and the singular-matrix flag come home with the stream (pinned memory): the
            # iteration loop is never stalled by a host round trip; both are checked below
            ctx.fs_iter_device(asyncDJ=True)
        if with_prd:
            ctx.prd_redistribute_device(maxIter=3, tol=1e-2)
        else:
            ctx.stat_eq_device(wait=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = 0
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # kernels launched per step, counted by the library itself (lwb200_work_stats)
    per_step = 0
    if stokes:
        step()
        per_step += ctx.work_stats()[2]
    elif world > 1 and not column_sharded:
        sharding.sharded_gamma_iteration(shard, group=None, want_dJ=False)
        per_step += 1  # the all-reduce
    else:
        ctx.fs_iter_device(asyncDJ=True)
    if not stokes:
        per_step += ctx.work_stats()[2]
    if stokes:
        pass
    elif with_prd:
        ctx.prd_redistribute_device(maxIter=3, tol=1e-2)
    else:
        ctx.stat_eq_device(wait=False)
    if not stokes:
        per_step += ctx.work_stats()[2]
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    for s in range(args.steps):
        flush.zero_()  # untimed L2 flush between steps
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
        kernel_ms.append(ctx.kernel_time_ms())
        launches += per_step
    barrier()
    clk = clocks.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if not stokes and not with_prd:
        ctx.sync()
        ctx.check_singular()   # raises if any timed step met a singular system
        if not (world > 1 and not column_sharded):
            dj_last = ctx.last_dj()[0]
            if not (dj_last == dj_last and dj_last >= 0.0):
                raise SystemExit(f'bench.py: bad dJ {dj_last}')
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    pts = torch.tensor([pts_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pts, op=dist.ReduceOp.SUM)
    # Short timed regions (a 1D atmosphere is < 1 ms per step) end before nvidia-smi has sampled the clocks
    # three times: sample them under the SAME step repeated, untimed, for about a second and a half.  Whether
    # to do so and how many repeats are derived from the all-reduced time, i.e. they are identical on every
    # rank -- the lambda-sharded step contains a collective, so every rank must run it the same number of times.
    nprobe = clock_probe_repeats(t.item(), args.steps)
    if nprobe:
        clocks = ClockSampler(local_rank)
        clocks.start()
        for q in range(nprobe):
            step()
            if q % 20 == 19:
                torch.cuda.synchronize()
        barrier()
        clk = clocks.stop()
        clk['window'] = (f'{nprobe} untimed repeats of the same step right after the timed region '
                         f'(the timed region itself lasts {t.item():.1f} ms)')
    total_ms = t.item()
    pts_total = pts.item()
    ms_per_step = total_ms / args.steps
    value = pts_total / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers
    # per step: the populations (Gamma iteration), then populations + Gamma (stat_equil).  nStar / nTotal /
    # vBroad and the prefill crsw*C are context state: they cross once and again only after update_deps()
    h2d = sum(a.n.nbytes for a in problem.atoms)
    h2d += sum(a.n.nbytes + a.Gamma.nbytes for a in problem.active_atoms())  # stat_equil: n, Gamma
    # lambda-sharded loop: every rank sends the populations, reads back its own rows of J / I and the (replicated)
    # Gamma, rates and populations
    h2d_lambda_sharded = sum(a.n.nbytes for a in problem.atoms)
    d2h = problem.J.nbytes + problem.I.nbytes + sum(a.n.nbytes for a in problem.active_atoms())
    d2h += sum(a.Gamma.nbytes for a in problem.active_atoms())
    d2h += sum(t_.Rij.nbytes + t_.Rji.nbytes for a in problem.atoms for t_ in a.trans)
    if stokes:
        h2d = sum(a.n.nbytes for a in problem.atoms) + problem.J.nbytes
        d2h = problem.I.nbytes + problem.Quv.nbytes + problem.J.nbytes
    e2e = None
    if world == 1 or column_sharded:
        def api_step():
            if stokes:
                ctx.single_stokes_fs(updateJ=True, upOnly=False)
                return
            ctx.formal_sol_gamma_matrices()
            if with_prd:
                ctx.prd_redistribute(maxIter=3, tol=1e-2)
            else:
                ctx.stat_equil()
        for _ in range(2):
            api_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            api_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': pts_total / te.item(), 'unit': 'points/s', 'ms_per_step': te.item() * 1e3,
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'api': ('Context.single_stokes_fs(updateJ=True, upOnly=False) on host numpy buffers' if stokes else
                       'Context.formal_sol_gamma_matrices() + Context.stat_equil() on host numpy buffers')}
    else:
        # lambda-sharded: each rank uploads the (replicated) small inputs, reads back its J rows
        # (the prefill crsw*C, nStar, nTotal, vBroad are context state, as in Context.formal_sol_gamma_matrices)
        problem.prefill_gamma()
        ctx.upload(capi.ITER_INPUTS)
        for _ in range(2):
            ctx.upload(capi.POPS)
            step()
            ctx.download(capi.ITER_OUTPUTS | capi.POPS | capi.OWN_ROWS)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.upload(capi.POPS)
            step()
            ctx.download(capi.ITER_OUTPUTS | capi.POPS | capi.OWN_ROWS)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': pts_total / te.item(), 'unit': 'points/s', 'ms_per_step': te.item() * 1e3,
               'h2d_bytes_per_step': int(h2d_lambda_sharded * world),
               'd2h_bytes_per_step': int(problem.J.nbytes + problem.I.nbytes
                                         + world * (d2h - problem.J.nbytes - problem.I.nbytes)),
               'api': 'per rank: upload(POPS) + sharded Gamma iteration + stat_eq + download of Gamma, rates, populations and its own rows of J, I'}

    if rank != 0:
        ctx.close()
        return None

    peak, peak_src = measured_peaks()
    kms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(args.workload)
            if args.workload == 'c3' and tj.get('c3_per_column'):
                traffic = tj['c3_per_column'] * problem.Ncol
        except Exception:
            traffic = None
    # secondary, fp64-pipe view of the same launch set (F_alg ~ W (110 + 23 tbar) flop,
    # tbar = mean number of active transitions per wavelength), against the DFMA peak measured on
    # this pool's B200 by tools/fp64_peak.cu (profiles/*_fp64_peak.json)
    tbar = sum(t_.Nlambda for a in problem.atoms for t_ in a.trans) / float(problem.Nspect)
    flop_pt = 110.0 + 23.0 * tbar
    fp64_peak = 36.0
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('fp64_peak.json'))
        fp64_peak = float(json.load(open(os.path.join(ROOT, 'profiles', cands[-1])))['fp64_tflops'])
    except Exception:
        pass
    fp64_ach = flop_pt * pts_local / (kms * 1e-3) / 1e12
    line = {
        'metric': 'depth*lambda*mu*column ray-depth points/s, fp64 Gamma iteration (formal solution + Gamma/rates + stat-eq)',
        'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'gamma_iter_per_s': 1e3 / ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {desc}', 'Nspace': problem.Nspace, 'Nspect': problem.Nspect,
                   'Nrays': problem.Nrays, 'Ncolumns': args.columns if column_sharded else 1,
                   'points_per_step': pts_total,
                   'formal_solver': 'piecewise_bezier3_1d',
                   'parallelism': ('1 GPU' if world == 1 else
                                   (f'column-sharded x{world}, no data-path collective' if column_sharded else
                                    f'lambda-sharded x{world}, one all-reduce(sum) of packed [Gamma|R] per step')),
                   'l2': f'{L2_FLUSH_BYTES >> 20} MiB buffer written between timed steps (untimed); per-step CUDA events summed'},
        'e2e': e2e,
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'kernel': 'continuum_kernel + ray_kernel<NL=0..3> + gamma_kernel (one launch set per iteration)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'alg_bytes_per_launch': alg_bytes, 'kernel_ms': kms,
                     'kernel_share_of_step': kms / ms_per_step,
                     'note': 'fp64-pipe / dependent-issue bound, not HBM bound: see DESIGN.md section 5',
                     'fp64': {'alg_flop_per_point': flop_pt, 'achieved': fp64_ach, 'peak': fp64_peak,
                              'unit': 'TFLOP/s', 'frac': fp64_ach / fp64_peak,
                              'peak_source': 'tools/fp64_peak.cu on this pool (profiles/*_fp64_peak.json)'}},
        'clocks': clk,
    }
    if world == 1:
        try:
            ref_problem, _ = build_workload(args.workload, min(args.columns, 64), for_gpu=False)
            cb = time_reference(ref_problem, steps=10, warmup=2, budget_s=25.0)
            line['cpu_baseline'] = {'value': cb['value'], 'unit': 'points/s', 'cores': cb['cores'],
                                    'kind': cb['kind'], 'sample': cb['sample'], 'scheme': cb['scheme']}
        except Exception as e:  # the baseline must never take the bench line down
            line['cpu_baseline'] = {'value': None, 'unit': 'points/s', 'cores': os.cpu_count(), 'kind': 'reference',
                                    'sample': f'failed: {e}'}
    ctx.close()
    return line


def run_reference(args, rank, world):
    if rank != 0:
        return None
    problem, desc = build_workload(args.workload, min(args.columns, 64), for_gpu=False)
    cb = time_reference(problem, steps=args.steps, warmup=args.warmup, budget_s=150.0)
    scale = (args.columns / problem.Ncol) if args.workload == 'c3' else 1.0
    ms = cb['sec_per_step'] * 1e3 * scale
    line = {
        'impl': 'reference',
        'metric': 'depth*lambda*mu*column ray-depth points/s, fp64 Gamma iteration (formal solution + Gamma/rates + stat-eq)',
        'value': cb['value'], 'unit': 'points/s', 'n_gpus': world, 'steps': cb['steps'], 'warmup': args.warmup,
        'ms_per_step': ms, 'gamma_iter_per_s': 1e3 / ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {desc}', 'Nspace': problem.Nspace, 'Nspect': problem.Nspect,
                   'Nrays': problem.Nrays, 'Ncolumns': args.columns if args.workload == 'c3' else 1,
                   'parallelism': f'{cb["cores"]} host threads, {cb["scheme"]}'},
        'cpu_baseline': {'value': cb['value'], 'unit': 'points/s', 'cores': cb['cores'], 'kind': cb['kind'],
                         'sample': cb['sample'], 'scheme': cb['scheme']},
        'e2e': {'value': cb['value'], 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    return line


def main():
    # only the JSON line may reach stdout (NCCL and torch.distributed print banners there)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--columns', type=int, default=None, help='columns of the c3 (4096) / c5 (1024) stacks')
    args = ap.parse_args()
    if args.columns is None:
        args.columns = 1024 if args.workload == 'c5' else 4096

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 back end has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        line = run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()
    return line


if __name__ == '__main__':
    main()
