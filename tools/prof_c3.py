"""Small driver for ncu: a few device-resident Gamma iterations of a column stack."""
import sys
sys.path.insert(0, '.')
import torch
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = sys.argv[3] if len(sys.argv) > 3 else 'c3'
if wl == 'c3':
    p = synth.config_c3(ncol=ncol, with_profiles=False, alloc_phi=False)
    ctx = Context(p, upload=False)
    ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
    ctx.update_deps(background=False, profiles_on_device=True)
else:
    p = synth.config_c2()
    ctx = Context(p)
for it in range(nit):
    ctx.fs_iter_device(want_dJ=False)
    ctx.stat_eq_device()
ctx.sync()
print('kernel ms', ctx.kernel_time_ms(), 'pts/s %.3e' % (p.points_per_iter() / (ctx.kernel_time_ms() * 1e-3)))
