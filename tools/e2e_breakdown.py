"""Where does the host-side time of one e2e step go? (development aid)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
p = synth.config_c2()
ctx = Context(p)
for _ in range(3):
    ctx.formal_sol_gamma_matrices(); ctx.stat_equil()
def T(f, n=50):
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    ctx.sync(); return (time.perf_counter() - t0) / n * 1e6
print('full step           %.0f us' % T(lambda: (ctx.formal_sol_gamma_matrices(), ctx.stat_equil())))
print('fs_gamma_matrices   %.0f us' % T(ctx.formal_sol_gamma_matrices))
print('stat_equil          %.0f us' % T(ctx.stat_equil))
print('prefill_gamma       %.0f us' % T(lambda: p.prefill_gamma(1.0)))
print('upload ITER_INPUTS  %.0f us' % T(lambda: ctx.upload(capi.ITER_INPUTS)))
print('fs_iter_device+dJ   %.0f us' % T(lambda: ctx.fs_iter_device()))
print('fs_iter_device      %.0f us' % T(lambda: ctx.fs_iter_device(want_dJ=False)))
print('download ITER_OUT   %.0f us' % T(lambda: ctx.download(capi.ITER_OUTPUTS)))
print('download J only     %.0f us' % T(lambda: ctx.download(capi.JBAR)))
print('download G,I,R      %.0f us' % T(lambda: ctx.download(capi.GAMMA | capi.INTENS | capi.RATES)))
print('upload POPS|GFINAL  %.0f us' % T(lambda: ctx.upload(capi.POPS | capi.GAMMA_FINAL)))
print('stat_eq_device      %.0f us' % T(ctx.stat_eq_device))
print('download POPS       %.0f us' % T(lambda: ctx.download(capi.POPS)))
