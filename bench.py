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

Workloads: c3 (DEFAULT headline; configs[2]: stack of 4096 perturbed FAL C columns,
H + Ca II, 5 rays -- the largest single-GPU configuration; column-sharded, no
data-path collective when N > 1), c2 (configs[1]: FAL C 1D, H + Ca II + Mg II +
Na I + He I, ~1e4 wavelengths x 10 rays; lambda-sharded with one all-reduce of the
packed [Gamma|R] buffer per step when N > 1), c1 (configs[0]), c4 (PRD), c5 (full
Stokes).  The default run also measures c2 and reports it in the same JSON line
under "secondary" (so the scaling run exposes both the column-sharded and the
lambda-sharded curve); `--workload X` runs X alone.  `--impl reference` times the
reference's CPU implementation (oracle/_ref, else the C port) on the same workload.

Every run ends with an UNTIMED parity spot check of the very state it timed
("parity_check"): a few columns (c3) / the whole atmosphere (1D workloads) are
stepped once more on the device and compared with the oracle started from the
same device state.
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


def build_workload(name, columns, rank=0, world=1, for_gpu=True, desc_columns=None):
    """Returns (problem, description).  For c3 each rank builds only its column shard.  desc_columns: the
    column count the DESCRIPTION names (the reference arm times a bounded sample of the stack but its
    `config` must read exactly like ours)."""
    dcol = columns if desc_columns is None else desc_columns
    if name == 'c1':
        return synth.config_c1(), 'FAL C 1D, H 6-level + Ca II 5+1-level, 5 rays (configs[0])'
    if name == 'deep':
        return synth.config_c1(ndepth=500), ('FAL C interpolated to 500 depths, H 6-level + Ca II 5+1-level, 5 rays: the '
                                             "reference's own benchmark protocol (lightweaver/benchmark.py:19-45)")
    if name == 'c2':
        return synth.config_c2(), ('FAL C 1D, H + Ca II + Mg II + Na I + He I active, 10 rays '
                                   '(configs[1], synthetic atomic data)')
    if name == 'c4':
        return synth.config_c4(nl=2.0), ('FAL C 1D, H + Ca II + Mg II, Mg II h&k in angle-averaged PRD, 5 rays '
                                         '(configs[3]; each step = Gamma iteration + up to 3 PRD sub-iterations on '
                                         'fixed LTE populations -- no stat-eq, the synthetic atom is not meant to be '
                                         'iterated to convergence with PRD; points counted for the Gamma iteration only)')
    if name == 'c4h':
        p = synth.config_c4(nl=2.0)
        p.configure_hprd()
        return p, ('FAL C 1D, H + Ca II + Mg II, Mg II h&k in HYBRID PRD (configure_hprd_coeffs: rho interpolated per '
                   'ray to the rest frame, JRest scattered by the formal solution), 5 rays (configs[3] as BASELINE words '
                   'it; each step = Gamma iteration + up to 3 PRD sub-iterations, as c4)')
    if name == 'c5':
        from lightweaver_b200.sharding import partition_columns
        c0, c1 = partition_columns(columns, world)[rank]
        # (on the GPU side the profiles -- phi and the six polarised ones -- are made on the device)
        p = synth.config_c5(ncol=columns, col_range=(c0, c1), with_profiles=not for_gpu, alloc_phi=not for_gpu)
        return p, (f'1.5D magnetised stack of {dcol} perturbed FAL C columns x 82 depths, Ca II with the 854.2 nm '
                   'line Zeeman-split and polarised, 5 rays (configs[4]; each step = one J-updating full-Stokes '
                   'formal solution of every wavelength, formal_sol_full_stokes(updateJ=True, upOnly=False))')
    if name == 'c3':
        from lightweaver_b200.sharding import partition_columns
        c0, c1 = partition_columns(columns, world)[rank]
        p = synth.config_c3(ncol=columns, col_range=(c0, c1), with_profiles=not for_gpu,
                            alloc_phi=not for_gpu)
        return p, f'1.5D stack of {dcol} perturbed FAL C columns x 82 depths, H + Ca II, 5 rays (configs[2])'
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
METRIC = 'depth*lambda*mu*column ray-depth points/s, fp64 Gamma iteration (formal solution + Gamma/rates + stat-eq)'
PLUGIN = os.path.join(ROOT, 'lightweaver_b200', 'liblwb200_plugin.so')


def workload_config(workload, desc, problem, columns):
    """The `config` object of the JSON line: a pure function of the workload, identical in both arms."""
    column_stack = workload in ('c3', 'c5')
    ncol = columns if column_stack else 1
    return {'workload': f'{workload}: {desc}', 'Nspace': problem.Nspace, 'Nspect': problem.Nspect,
            'Nrays': problem.Nrays, 'Ncolumns': ncol,
            'points_per_step': float(ncol) * problem.Nspace * problem.Nspect * problem.Nrays * 2,
            'formal_solver': 'piecewise_bezier3_1d',
            'l2': f'{L2_FLUSH_BYTES >> 20} MiB buffer written between timed steps (untimed); per-step CUDA events summed'}


def _packed_views(ctx, problem):
    """torch views (no copies) of the library's packed device buffers of this context."""
    from lightweaver_b200.sharding import device_tensor
    Ncol, L, K, M = problem.Ncol, problem.Nspect, problem.Nspace, problem.Nrays
    nlev = sum(a.Nlevel for a in problem.atoms)
    ngam = sum(a.Nlevel * a.Nlevel for a in problem.atoms if not a.detailedStatic)
    v = {}
    for name, which, shape in (('J', capi.BUF_J, (Ncol, L, K)), ('I', capi.BUF_I, (Ncol, L, M)),
                               ('n', capi.BUF_POPS, (Ncol, nlev, K)), ('Gamma', capi.BUF_GAMMA, (Ncol, max(ngam, 1), K))):
        ptr, nbytes = ctx.device_buffer(which)
        v[name] = device_tensor(ptr, nbytes, ctx.device).view(*shape)
    return v


def parity_spot_check(ctx, problem, workload, columns, col0, step, laRange=None, ranges=None, rank=0, world=1,
                      ncheck=4, tol=1e-9):
    """UNTIMED check of the run that was just timed: read the device state (populations, J) of a few
    columns, run ONE more step on the device, and compare its I, J, Gamma and populations with the
    oracle (oracle/lw_oracle.c, bit-identical to the reference's scalar scheme) started from that same
    state on host-made (Faddeeva) profiles.  Collective-free on column stacks; on a lambda-sharded run
    every rank takes part (the J rows are gathered) and rank 0 compares.  Returns the dict that goes
    into the JSON line as "parity_check"."""
    import torch
    from lightweaver_b200 import sharding
    from oracle import oraclelib
    from tests.util import gamma_err, rel_err
    v = _packed_views(ctx, problem)
    Ncol = problem.Ncol
    cols = sorted(set(int(c) for c in np.linspace(0, Ncol - 1, min(ncheck, Ncol))))
    torch.cuda.synchronize()
    lambda_sharded = laRange is not None and world > 1
    if lambda_sharded:
        lo, hi = laRange
        Jfull = sharding.gather_rows(v['J'][0, lo:hi].clone(), ranges)
        J0 = Jfull.cpu().numpy()[None]
    else:
        J0 = v['J'][cols].cpu().numpy()
    n0 = v['n'][cols].cpu().numpy()
    step()
    torch.cuda.synchronize()
    ctx.check_singular()
    out = {k: v[k][cols].cpu().numpy() for k in ('I', 'J', 'Gamma', 'n')}
    if rank != 0 and lambda_sharded:
        return None
    err = {'I': 0.0, 'J': 0.0, 'Gamma': 0.0, 'n': 0.0}
    for qi, c in enumerate(cols):
        if workload == 'c3':
            q = synth.config_c3(ncol=columns, col_range=(col0 + c, col0 + c + 1))
        else:
            q = build_workload(workload, columns, for_gpu=False)[0]
        q.J[0] = J0[qi]
        lev = gam = 0
        for a in q.atoms:
            a.n[0] = n0[qi, lev:lev + a.Nlevel]
            lev += a.Nlevel
        q.prefill_gamma()
        o = oraclelib.OracleContext(q)
        o.fs_iter(lambdaIterate=False)
        o.stat_eq()
        rows = slice(*laRange) if lambda_sharded else slice(None)
        err['I'] = max(err['I'], rel_err(out['I'][qi][rows], q.I[0][rows]))
        err['J'] = max(err['J'], rel_err(out['J'][qi][rows], q.J[0][rows]))
        lev = 0
        for a in q.atoms:
            err['n'] = max(err['n'], rel_err(out['n'][qi, lev:lev + a.Nlevel], a.n[0]))
            lev += a.Nlevel
            if a.detailedStatic:
                continue
            N2 = a.Nlevel * a.Nlevel
            G = out['Gamma'][qi, gam:gam + N2].reshape(a.Nlevel, a.Nlevel, -1)
            err['Gamma'] = max(err['Gamma'], gamma_err(G, a.Gamma[0]))
            gam += N2
    ok = err['I'] <= tol and err['J'] <= tol and err['Gamma'] <= tol and err['n'] <= 10 * tol
    return {'ok': bool(ok), 'against': 'oracle/lw_oracle.c (bit-identical to the reference scalar scheme) from the '
            'same device state, one untimed step after the timed region',
            'columns_checked': [col0 + c for c in cols], 'max_rel_err': err,
            'tol': {'I': tol, 'J': tol, 'Gamma (norm-wise per depth)': tol, 'n': 10 * tol}}


def stokes_spot_check(ctx, problem, columns, col0, step, ncheck=4, tol=1e-9):
    """parity_spot_check for the full-Stokes workload: J-dagger and the populations of a few columns are read
    from the device, ONE more J-updating full-Stokes pass runs there (on device-made profiles), and its I, Q, U, V
    (where a polarised line is active) and J are compared with the oracle's formal_sol_full_stokes of the same
    columns started from that state on host-made (Faddeeva) profiles."""
    import torch
    from oracle import oraclelib
    from tests.util import rel_err
    v = _packed_views(ctx, problem)
    Ncol = problem.Ncol
    cols = sorted(set(int(c) for c in np.linspace(0, Ncol - 1, min(ncheck, Ncol))))
    torch.cuda.synchronize()
    J0 = v['J'][cols].cpu().numpy()
    n0 = v['n'][cols].cpu().numpy()
    step()
    torch.cuda.synchronize()
    ctx.download(capi.INTENS | capi.JBAR | capi.STOKES)
    err = {'I': 0.0, 'QUV': 0.0, 'J': 0.0}
    for qi, c in enumerate(cols):
        q = synth.config_c5(ncol=columns, col_range=(col0 + c, col0 + c + 1))
        q.J[0] = J0[qi]
        lev = 0
        for a in q.atoms:
            a.n[0] = n0[qi, lev:lev + a.Nlevel]
            lev += a.Nlevel
        oraclelib.OracleContext(q).full_stokes(updateJ=True, upOnly=False)
        pol = np.zeros(q.Nspect, dtype=bool)
        for a in q.atoms:
            for t_ in a.trans:
                if t_.type == capi.LINE and t_.polProfiles is not None:
                    pol[t_.Nblue:t_.Nred] = True
        err['I'] = max(err['I'], rel_err(problem.I[c], q.I[0]))
        err['J'] = max(err['J'], rel_err(problem.J[c], q.J[0]))
        err['QUV'] = max(err['QUV'], float(np.abs(problem.Quv[c][:, pol] - q.Quv[0][:, pol]).max() / np.abs(q.I[0]).max()))
    ok = all(e <= tol for e in err.values())
    return {'ok': bool(ok), 'against': 'oracle/lw_oracle.c formal_sol_full_stokes (bit-identical to the reference) from '
            'the same device state, one untimed pass after the timed region',
            'columns_checked': [col0 + c for c in cols], 'max_rel_err': err,
            'tol': {'I': tol, 'J': tol, 'Q, U, V at polarised wavelengths (relative to max I)': tol}}


def prd_spot_check(ctx, problem, step, tol=1e-9):
    """parity_spot_check for the PRD workload (one 1D atmosphere): J, the populations and rho of the PRD lines are
    read from the device, ONE more step (Gamma iteration + PRD sub-iterations) runs there, and its I, J, rho and
    rates are compared with the oracle's formal solution + redistribute_prd_lines from that state."""
    import torch
    from oracle import oraclelib
    from tests.util import rel_err
    torch.cuda.synchronize()
    ctx.download(capi.JBAR | capi.POPS | capi.PRD)
    q = problem.clone()
    step()
    torch.cuda.synchronize()
    ctx.download(capi.ITER_OUTPUTS | capi.PRD)
    nprd = sum(1 for a in q.atoms for t_ in a.trans if t_.rhoPrd is not None)
    q.prefill_gamma()
    o = oraclelib.OracleContext(q)
    o.fs_iter()
    o.redistribute_prd(maxIter=3, tol=1e-2, nlines=nprd)
    err = {'I': rel_err(problem.I, q.I), 'J': rel_err(problem.J, q.J), 'rho': 0.0, 'R': 0.0}
    for a, b in zip(problem.atoms, q.atoms):
        for ta, tb in zip(a.trans, b.trans):
            if ta.rhoPrd is not None:
                err['rho'] = max(err['rho'], rel_err(ta.rhoPrd, tb.rhoPrd))
            err['R'] = max(err['R'], rel_err(ta.Rij, tb.Rij), rel_err(ta.Rji, tb.Rji))
    ok = all(e <= tol for e in err.values())
    return {'ok': bool(ok), 'against': 'oracle/lw_oracle.c formal solution + redistribute_prd_lines (bit-identical to the '
            'reference) from the same device state, one untimed step after the timed region',
            'max_rel_err': err, 'tol': {k: tol for k in err}}


def plugin_e2e(problem, steps):
    """e2e through the REAL drop-in boundary for a 1D atmosphere: the reference's own compiled core
    (oracle/_ref/liblwref.so = Lightweaver's C++ `formal_sol_gamma_matrices` / `stat_eq` entry points) loads
    liblwb200_plugin.so with ITS plugin manager and calls the scheme `mali_full_precond_B200` on HOST
    buffers -- the same harness call the reference arm times with the AVX2FMA scheme.  The reference
    core is only the caller here; every kernel is ours."""
    from oracle import reflib
    if not (reflib.available() and os.path.exists(PLUGIN)):
        return None
    problem.prefill_gamma()
    r = reflib.RefContext(problem, scheme=PLUGIN, Nthreads=1)
    try:
        ts = r.time_fs_iter(3, steps, True)
    finally:
        r.close()
    return float(np.median(ts))


def bench_workload(args, workload, columns, rank, world, local_rank, with_cpu_baseline=True):
    import torch
    import torch.distributed as dist
    from lightweaver_b200.context import Context
    from lightweaver_b200 import sharding

    dev = torch.device('cuda', local_rank)
    problem, desc = build_workload(workload, columns, rank, world)
    column_sharded = workload in ('c3', 'c5')
    stokes = workload == 'c5'
    with_prd = workload in ('c4', 'c4h')
    col0 = sharding.partition_columns(columns, world)[rank][0] if column_sharded else 0
    laRange = ranges = None
    if world > 1 and not column_sharded:
        ranges = sharding.partition_wavelengths(problem, world)
        laRange = ranges[rank]
    stream = torch.cuda.current_stream()
    ctx = Context(problem, device=local_rank, stream=stream, laRange=laRange, upload=False)
    if stokes:
        ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
        ctx.update_deps(background=False, profiles_on_device=True)
        ctx.compute_polarised_profiles_device()
    elif column_sharded:
        ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
        ctx.update_deps(background=False, profiles_on_device=True)
    else:
        ctx.upload(capi.ALL_INPUTS | (capi.PRD if with_prd else 0))
    ctx.sync()
    if with_prd and world > 1:
        raise SystemExit('bench.py: the PRD workload (c4) runs on one GPU')
    shard = sharding.GpuLambdaShard(ctx)
    pts_local, alg_bytes, _ = ctx.work_stats()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    lambda_sharded = world > 1 and not column_sharded

    def plain_step():
        if stokes:
            capi.check(ctx.lib.lwb200_formal_sol_full_stokes(ctx._h, 1, 0, None, None))
            return
        if lambda_sharded:
            sharding.sharded_gamma_iteration(shard, group=None, want_dJ=False)
        else:
            # dJ and the singular-matrix flag come home with the stream (pinned memory): the
            # iteration loop is never stalled by a host round trip; both are checked below
            ctx.fs_iter_device(asyncDJ=True)
        if with_prd:
            ctx.prd_redistribute_device(maxIter=3, tol=1e-2)
        else:
            ctx.stat_eq_device(wait=False)

    # A 1D atmosphere is ~15 launches of 10-100 us: the whole iteration (with the all-reduce of a wavelength
    # shard) is captured once into a CUDA graph and replayed.  Column stacks (100 ms steps) launch directly.
    graphed = None
    use_graph = not column_sharded and not stokes and not with_prd

    def step():
        if graphed is not None:
            graphed.replay()
        else:
            plain_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = 0
    kernel_ms = []
    for _ in range(max(args.warmup, 3)):
        plain_step()
    barrier()
    if use_graph:
        for _ in range(5):   # (the launch-set time is not available inside a graph: taken from direct launches)
            plain_step()
            kernel_ms.append(ctx.kernel_time_ms())
        barrier()
    if use_graph:
        graphed = sharding.GraphedIteration(shard, group=None)
        for _ in range(3):
            step()
        barrier()
    # kernels launched per step, counted by the library itself (lwb200_work_stats)
    per_step = 0
    if stokes:
        step()
        per_step += ctx.work_stats()[2]
    elif lambda_sharded:
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
    barrier()
    for s in range(args.steps):
        flush.zero_()  # untimed L2 flush between steps
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
        if graphed is None:
            kernel_ms.append(ctx.kernel_time_ms())
        launches += per_step
    barrier()
    clk = clocks.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if not stokes and not with_prd:
        ctx.sync()
        ctx.check_singular()   # raises if any timed step met a singular system
        if not lambda_sharded and graphed is None:
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

    # ---- untimed parity spot check of the state just timed
    parity = None
    if True:
        try:
            if with_prd:
                parity = prd_spot_check(ctx, problem, step)
            elif stokes:
                parity = stokes_spot_check(ctx, problem, columns, col0, step)
            else:
                parity = parity_spot_check(ctx, problem, workload, columns, col0, step, laRange=laRange, ranges=ranges,
                                           rank=rank, world=world)
        except Exception as e:  # reported, never fatal for the measurement itself
            parity = {'ok': False, 'error': f'{type(e).__name__}: {e}'}
        barrier()

    # ---- end to end with HOST buffers (h2d / d2h counted from the arrays that cross)
    # Per call the boundary sends what the reference's caller may have changed since the last call --
    # populations, nStar / nTotal / vBroad -- and brings home everything the caller reads: J, I, Gamma, rates,
    # populations.  The prologue Gamma = crsw*C (LwMiddleLayer.pyx:3198-3203) is made on the device from the
    # resident collisional rates by the ctypes mirror (lwb200_set_collision_prefill); the plugin, which is
    # handed a prefilled Gamma by the reference core, uploads it on every call.
    h2d = sum(2 * a.n.nbytes + a.nStar.nbytes + a.nTotal.nbytes + a.vBroad.nbytes for a in problem.atoms)
    h2d += sum(a.Gamma.nbytes for a in problem.active_atoms())  # the final Gamma stat_equil reads
    gamma_prefill_bytes = sum(a.Gamma.nbytes for a in problem.active_atoms())
    d2h = problem.J.nbytes + problem.I.nbytes + sum(a.n.nbytes for a in problem.active_atoms())
    d2h += sum(a.Gamma.nbytes for a in problem.active_atoms())
    d2h += sum(t_.Rij.nbytes + t_.Rji.nbytes for a in problem.atoms for t_ in a.trans)
    if stokes:
        h2d = sum(a.n.nbytes for a in problem.atoms) + problem.J.nbytes
        d2h = problem.I.nbytes + problem.Quv.nbytes + problem.J.nbytes
    e2e = None
    if world == 1 and not column_sharded and not with_prd:
        # a 1D atmosphere: through the plugin the reference's own core loads (the real drop-in boundary)
        ctx.sync()
        sec = plugin_e2e(problem.clone(), args.steps)
        if sec is not None:
            e2e = {'value': pts_total / sec, 'unit': 'points/s', 'ms_per_step': sec * 1e3,
                   'h2d_bytes_per_step': int(h2d + gamma_prefill_bytes + problem.J.nbytes), 'd2h_bytes_per_step': int(d2h),
                   'api': ('reference core (oracle/_ref/liblwref.so: Lightweaver formal_sol_gamma_matrices + stat_eq) -> '
                           'FsIterationFnsManager -> liblwb200_plugin.so fs_iteration_fns_provider -> b200_fs_iter / '
                           'b200_stat_eq on host buffers; per call: H2D of n, nStar, nTotal, vBroad, crsw*C (+ J, '
                           'background, profiles, atmosphere when their fingerprint changed), D2H of J, I, Gamma, rates, n')}
    if e2e is None and (world == 1 or column_sharded):
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
               'h2d_bytes_per_step': int(h2d) * (world if column_sharded else 1),
               'd2h_bytes_per_step': int(d2h) * (world if column_sharded else 1),
               'api': ('C-ABI (lwb200_upload / lwb200_formal_sol_full_stokes / lwb200_download) on host numpy buffers'
                       if stokes else
                       'C-ABI of include/lwb200.h through its ctypes mirror (Context.formal_sol_gamma_matrices + '
                       'stat_equil): lwb200_upload(POPS|NSTAR) -> lwb200_fs_iter (prefill crsw*C from the device-resident '
                       'collisional rates) -> lwb200_download(J, I, Gamma, rates) -> lwb200_upload(POPS|GAMMA_FINAL) -> '
                       'lwb200_stat_eq -> lwb200_download(POPS), host numpy buffers')}
    elif e2e is None:
        # lambda-sharded: each rank uploads the (replicated) small inputs, reads back its J rows
        problem.prefill_gamma()
        for _ in range(2):
            ctx.upload(capi.ITER_INPUTS)
            step()
            ctx.download(capi.ITER_OUTPUTS | capi.POPS | capi.OWN_ROWS)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.upload(capi.ITER_INPUTS)
            step()
            ctx.download(capi.ITER_OUTPUTS | capi.POPS | capi.OWN_ROWS)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        up = sum(a.n.nbytes + a.nStar.nbytes + a.nTotal.nbytes + a.vBroad.nbytes for a in problem.atoms)
        up += sum(a.Gamma.nbytes for a in problem.active_atoms())
        e2e = {'value': pts_total / te.item(), 'unit': 'points/s', 'ms_per_step': te.item() * 1e3,
               'h2d_bytes_per_step': int(up * world),
               'd2h_bytes_per_step': int(problem.J.nbytes + problem.I.nbytes
                                         + world * (d2h - problem.J.nbytes - problem.I.nbytes)),
               'api': 'C-ABI per rank: lwb200_upload(POPS|NSTAR|GAMMA) + sharded Gamma iteration + stat_eq + download of '
                      'Gamma, rates, populations and its own rows of J, I'}

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
            traffic = tj.get(workload)
            if workload == 'c3' and tj.get('c3_per_column'):
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
        'metric': METRIC,
        'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'gamma_iter_per_s': 1e3 / ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(workload, desc, problem, columns),
        'launch': ('one CUDA graph per step (whole iteration incl. the collective)' if graphed is not None
                   else 'direct kernel launches'),
        'parallelism': ('1 GPU' if world == 1 else
                        (f'column-sharded x{world}, no data-path collective' if column_sharded else
                         f'lambda-sharded x{world}, one all-reduce(sum) of packed [Gamma|R] per step')),
        'e2e': e2e,
        'gpu_launches': launches,
        'parity_check': parity,
        'roofline': {'bound': 'hbm', 'kernel': ('stokes_kernel + ray_smem_kernel<NL=0..3> (one J-updating full-Stokes pass)' if stokes else 'continuum_table_kernel + ray_smem_kernel<NL=0..3> (ray_kernel<..,MULTI> beyond 128 depths) + gamma_tile_kernel / gamma_direct_kernel (one launch set per iteration; dominant: the ray kernel, 58 % of it)'), 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'alg_bytes_per_launch': alg_bytes, 'kernel_ms': kms,
                     'kernel_share_of_step': kms / ms_per_step,
                     'note': 'fp64-pipe / dependent-issue bound, not HBM bound: see DESIGN.md section 5',
                     'fp64': {'alg_flop_per_point': flop_pt, 'achieved': fp64_ach, 'peak': fp64_peak,
                              'unit': 'TFLOP/s', 'frac': fp64_ach / fp64_peak,
                              'peak_source': 'tools/fp64_peak.cu on this pool (profiles/*_fp64_peak.json)'}},
        'clocks': clk,
    }
    if world == 1 and with_cpu_baseline:
        try:
            ref_problem, _ = build_workload(workload, min(columns, 64), for_gpu=False)
            cb = time_reference(ref_problem, steps=10, warmup=2, budget_s=25.0 if workload == args.workload else 12.0)
            line['cpu_baseline'] = {'value': cb['value'], 'unit': 'points/s', 'cores': cb['cores'],
                                    'kind': cb['kind'], 'sample': cb['sample'], 'scheme': cb['scheme']}
        except Exception as e:  # the baseline must never take the bench line down
            line['cpu_baseline'] = {'value': None, 'unit': 'points/s', 'cores': os.cpu_count(), 'kind': 'reference',
                                    'sample': f'failed: {e}'}
    ctx.close()
    return line


def run_ours(args, rank, world, local_rank):
    import torch
    torch.cuda.set_device(local_rank)
    line = bench_workload(args, args.workload, args.columns, rank, world, local_rank)
    if args.secondary and args.workload == 'c3':
        # the lambda-sharded 1D workload (configs[1]) beside the column-sharded headline, same launch
        keep = ('value', 'unit', 'ms_per_step', 'gamma_iter_per_s', 'scaling', 'config', 'parallelism', 'e2e',
                'gpu_launches', 'parity_check', 'roofline', 'cpu_baseline')
        sec = bench_workload(args, 'c2', args.columns, rank, world, local_rank)
        if line is not None and sec is not None:
            line['secondary'] = {'c2_lambda_sharded': {k: sec[k] for k in keep if k in sec}}
        # the other BASELINE configs, so that one default run measures all five: the magnetised full-Stokes stack
        # (configs[4]) and the small 1D cases (configs[0], configs[3]); and the reference's own benchmark shape
        # (lightweaver/benchmark.py:19-45: FAL C at 500 depths)
        # (one GPU only: a secondary that failed on one rank of a multi-rank launch would leave the others
        # waiting in its reductions)
        extra = ([('c5_full_stokes_stack', 'c5', 1024), ('c1', 'c1', 1), ('c4_prd', 'c4', 1), ('c4_hybrid_prd', 'c4h', 1),
                  ('deep_500_depths_reference_benchmark_shape', 'deep', 1)] if world == 1 else [])
        for key, wl, cols in extra:
            try:
                sec = bench_workload(args, wl, cols, rank, world, local_rank, with_cpu_baseline=False)
            except BaseException as e:  # (a secondary must never take the headline down)
                sec = {'error': f'{type(e).__name__}: {e}'} if rank == 0 else None
            if line is not None and sec is not None:
                line['secondary'][key] = {k: sec[k] for k in keep + ('error',) if k in sec}
    return line


def reference_line(args, workload, columns, world, budget_s=None):
    problem, desc = build_workload(workload, min(columns, 64), for_gpu=False, desc_columns=columns)
    if budget_s is None:
        budget_s = 150.0 if workload == args.workload else 40.0
    cb = time_reference(problem, steps=args.steps, warmup=args.warmup, budget_s=budget_s)
    scale = (columns / problem.Ncol) if workload in ('c3', 'c5') else 1.0
    ms = cb['sec_per_step'] * 1e3 * scale
    return {
        'impl': 'reference',
        'metric': METRIC,
        'value': cb['value'], 'unit': 'points/s', 'n_gpus': world, 'steps': cb['steps'], 'warmup': args.warmup,
        'ms_per_step': ms, 'gamma_iter_per_s': 1e3 / ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(workload, desc, problem, columns),
        'parallelism': f'{cb["cores"]} host threads, {cb["scheme"]}',
        'cpu_baseline': {'value': cb['value'], 'unit': 'points/s', 'cores': cb['cores'], 'kind': cb['kind'],
                         'sample': cb['sample'], 'scheme': cb['scheme']},
        'e2e': {'value': cb['value'], 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


def run_reference(args, rank, world):
    if rank != 0:
        return None
    line = reference_line(args, args.workload, args.columns, world)
    if args.secondary and args.workload == 'c3':
        sec = reference_line(args, 'c2', args.columns, world)
        keep = ('value', 'unit', 'ms_per_step', 'gamma_iter_per_s', 'config', 'parallelism', 'cpu_baseline', 'e2e')
        line['secondary'] = {'c2_lambda_sharded': {k: sec[k] for k in keep}}
        if world == 1:
            for key, wl, cols in (('c5_full_stokes_stack', 'c5', 1024), ('c1', 'c1', 1), ('c4_prd', 'c4', 1),
                                  ('c4_hybrid_prd', 'c4h', 1), ('deep_500_depths_reference_benchmark_shape', 'deep', 1)):
                try:
                    sec = reference_line(args, wl, cols, world, budget_s=15.0)
                    line['secondary'][key] = {k: sec[k] for k in keep}
                except Exception as e:
                    line['secondary'][key] = {'error': f'{type(e).__name__}: {e}'}
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
    ap.add_argument('--workload', default='c3', choices=['c1', 'c2', 'c3', 'c4', 'c4h', 'c5', 'deep'])
    ap.add_argument('--columns', type=int, default=None, help='columns of the c3 (4096) / c5 (1024) stacks')
    ap.add_argument('--no-secondary', dest='secondary', action='store_false',
                    help='default run (c3): do not also measure the lambda-sharded c2 workload')
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
