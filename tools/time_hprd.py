"""Device time of one Gamma iteration + PRD redistribution (3 sub-iterations) of config 4 on a column with config 3's
velocity field, angle-averaged vs hybrid PRD (development aid; hybrid wavelengths take the general per-ray kernel)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
for hybrid in (False, True):
    p = synth.config_c4(perturb=True, nl=2.0)
    if hybrid:
        p.configure_hprd()
    ctx = Context(p)
    ctx.upload(capi.PRD)
    ts = []
    for it in range(6):
        ctx.sync(); t0 = time.perf_counter()
        ctx.fs_iter_device(want_dJ=False)
        n = ctx.prd_redistribute_device(maxIter=3, tol=1e-12)
        ctx.sync(); ts.append(time.perf_counter() - t0)
    nh = 0 if p.hprd is None else int((p.hprd.hPrdLaOfLa >= 0).sum())
    print('hybrid' if hybrid else 'angle-averaged', 'L', p.Nspect, 'scattering wavelengths', nh,
          'ms per (Gamma iteration + %d PRD sub-iterations): %.3f' % (n, 1e3 * float(np.median(ts[2:]))), flush=True)
    ctx.close()
