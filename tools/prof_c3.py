"""Small driver for ncu / variant timing: a few device-resident Gamma iterations.

    python tools/prof_c3.py <ncol> <niter> <c3|c2|c1|c5|deep> [check]

Prints the median formal-solution kernel time; with `check`, also the parity of two
iterations of a small problem against the C oracle (development aid)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = sys.argv[3] if len(sys.argv) > 3 else 'c3'
if len(sys.argv) > 4 and sys.argv[4] == 'check':
    from oracle import oraclelib
    from tests.util import compare_problems
    for fs in (2,):
        p = synth.tiny_problem(ncol=2, perturb=True, formal_solver=fs)
        q = p.clone()
        ctx = Context(p, device=0)
        for it in range(2):
            ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
            ctx.stat_equil()
            q.prefill_gamma()
            for c in range(q.Ncol):
                o = oraclelib.OracleContext(q, col=c)
                o.fs_iter(lambdaIterate=(it == 0))
                o.stat_eq()
        print('parity fs', fs, ' '.join('%s=%.1e' % kv for kv in compare_problems(p, q).items()), flush=True)
        ctx.close()
if wl == 'c3':
    p = synth.config_c3(ncol=ncol, with_profiles=False, alloc_phi=False)
    ctx = Context(p, upload=False)
    ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
    ctx.update_deps(background=False, profiles_on_device=True)
elif wl == 'c5':
    p = synth.config_c5(ncol=ncol, with_profiles=False, alloc_phi=False)
    ctx = Context(p, upload=False)
    ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
    ctx.update_deps(background=False, profiles_on_device=True)
    ctx.compute_polarised_profiles_device()
elif wl == 'c1':
    p = synth.config_c1()
    ctx = Context(p)
elif wl == 'deep':
    p = synth.config_c1(ndepth=500)
    ctx = Context(p)
else:
    p = synth.config_c2()
    import os
    if os.environ.get('PROF_SHARDS'):   # time one wavelength shard of an N-way partition
        from lightweaver_b200 import sharding
        ctx = Context(p, laRange=sharding.partition_wavelengths(p, int(os.environ['PROF_SHARDS']))[0])
    else:
        ctx = Context(p)
ts = []
for it in range(nit):
    if wl == 'c5':   # one J-updating full-Stokes formal solution
        capi.check(ctx.lib.lwb200_formal_sol_full_stokes(ctx._h, 1, 0, None, None))
    else:
        ctx.fs_iter_device(want_dJ=False)
        ctx.stat_eq_device()
    ctx.sync()
    ts.append(ctx.kernel_time_ms())
ms = float(np.median(ts[1:] if len(ts) > 1 else ts))
print('%s ncol %d kernel ms %.4f pts/s %.3e' % (wl, ncol, ms, p.points_per_iter() / (ms * 1e-3)), flush=True)
