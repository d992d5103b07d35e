"""Where does the host-side time of one e2e step go? (development aid)"""
import sys, time
import ctypes as C
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
NREP = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if wl == 'c3':
    ncol = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    p = synth.config_c3(ncol=ncol, with_profiles=False, alloc_phi=False)
    ctx = Context(p, upload=False)
    ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
    ctx.update_deps(background=False, profiles_on_device=True)
else:
    p = getattr(synth, 'config_' + wl)()
    ctx = Context(p)
for _ in range(3):
    ctx.formal_sol_gamma_matrices(); ctx.stat_equil()
def T(f, n=NREP):
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    ctx.sync(); return (time.perf_counter() - t0) / n * 1e6
F = capi.FETCH_EARLY | capi.DJ_ASYNC
def fsi(flags): capi.check(ctx.lib.lwb200_fs_iter(ctx._h, flags, None, None))
print('host prefill_gamma             %.0f us' % T(lambda: p.prefill_gamma()))
print('full step                      %.0f us' % T(lambda: (ctx.formal_sol_gamma_matrices(), ctx.stat_equil())))
print('fs_gamma_matrices              %.0f us' % T(ctx.formal_sol_gamma_matrices))
print('stat_equil                     %.0f us' % T(ctx.stat_equil))
print('upload ITER_INPUTS             %.0f us' % T(lambda: ctx.upload(capi.ITER_INPUTS)))
print('fs_iter (no flags) + sync      %.0f us' % T(lambda: (fsi(0), ctx.sync())))
print('fs_iter (early+dj) + sync      %.0f us' % T(lambda: (fsi(F), ctx.sync())))
print('fs_iter(early+dj)+download all %.0f us' % T(lambda: (fsi(F), ctx.download(capi.ITER_OUTPUTS))))
print('fs_iter(dj)+download all       %.0f us' % T(lambda: (fsi(capi.DJ_ASYNC), ctx.download(capi.ITER_OUTPUTS))))
print('fs_iter(early+dj)+download G,R %.0f us' % T(lambda: (fsi(F), ctx.download(capi.GAMMA | capi.RATES))))
print('fs_iter(early+dj)+download J,I %.0f us' % T(lambda: (fsi(F), ctx.download(capi.JBAR | capi.INTENS))))
print('download G,R alone             %.0f us' % T(lambda: ctx.download(capi.GAMMA | capi.RATES)))
print('download J alone               %.0f us' % T(lambda: ctx.download(capi.JBAR)))
print('upload POPS                    %.0f us' % T(lambda: ctx.upload(capi.POPS)))
print('upload GAMMA                   %.0f us' % T(lambda: ctx.upload(capi.GAMMA)))
print('upload NSTAR                   %.0f us' % T(lambda: ctx.upload(capi.NSTAR)))
print('upload GAMMA_FINAL             %.0f us' % T(lambda: ctx.upload(capi.GAMMA_FINAL)))
print('download G alone               %.0f us' % T(lambda: ctx.download(capi.GAMMA)))
print('download R alone               %.0f us' % T(lambda: ctx.download(capi.RATES)))
print('download I alone               %.0f us' % T(lambda: ctx.download(capi.INTENS)))
print('download POPS alone            %.0f us' % T(lambda: ctx.download(capi.POPS)))
