"""Small end-to-end exercise of every kernel family for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import sys
sys.path.insert(0, '.')
import numpy as np
from lightweaver_b200 import synth, capi
from lightweaver_b200.context import Context

p = synth.tiny_problem(ncol=2, perturb=True)
ctx = Context(p)
for it in range(2):
    ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
    ctx.stat_equil()
ctx.formal_sol()
ctx.close()
print('scalar ok', flush=True)

p = synth.tiny_prd_problem(ncol=2, perturb=True)
p.configure_hprd()
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.prd_redistribute(maxIter=2, tol=1e-6)
ctx.stat_equil()
ctx.close()
print('hybrid prd ok', flush=True)

p = synth.tiny_prd_problem(ncol=2, perturb=True, ndepth=150)
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.prd_redistribute(maxIter=2, tol=1e-6)
ctx.close()
print('deep prd ok', flush=True)

p = synth.tiny_stokes_problem(ncol=2, perturb=True)
ctx = Context(p)
ctx.compute_polarised_profiles_device()
ctx.formal_sol_gamma_matrices()
J20 = np.zeros((p.Ncol, p.Nspect, p.Nspace))
ctx.single_stokes_fs(updateJ=True, upOnly=False, extraParams={'J20': J20})
ctx.single_stokes_fs(updateJ=False, upOnly=True)
ctx.close()
print('stokes ok', flush=True)

# static atmosphere (ray-independent profiles: the shared-phase path of both ray kernels), 82 and 300 depths
for nd in (None, 300):
    p = synth.tiny_problem(ndepth=nd, nrays=3)
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
    ctx.close()
print('static ok', flush=True)

# the deep general kernel: hybrid PRD at 200 depths, every wavelength on request, more than 32 rays per wavelength
p = synth.tiny_prd_problem(ncol=2, perturb=True, ndepth=200)
p.configure_hprd()
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.prd_redistribute(maxIter=2, tol=1e-6)
ctx.formal_sol_gamma_matrices(extraParams={'generalKernel': True})
ctx.formal_sol()
ctx.close()
p = synth.tiny_problem(nrays=17, ncol=2, perturb=True)
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.close()
print('deep general / many rays ok', flush=True)

# every wavelength through the general kernel at 82 depths (rays split over the warps of a CTA), a 40-level atom
# (population solve with global scratch, general tiles by RED), and 1100 depths (32-warp general kernel)
p = synth.tiny_problem(ncol=1, nrays=3)
ctx = Context(p)
ctx.formal_sol_gamma_matrices(extraParams={'generalKernel': True})
ctx.stat_equil()
ctx.close()
lev = [synth.Level(0.0, 2, 0)] + [synth.Level(55000.0 + 600.0 * i, 2 + 2 * (i % 4), 0) for i in range(38)] + [synth.Level(100000.0, 1, 1)]
big = synth.ModelAtom('Huge', 12.0, 1e-4, lev, [synth.LineSpec(1, 0, 2.0e8, 15, 4.0, 40.0)],
                      [synth.ContSpec(39, i, 4.0e-22, 5, 70.0) for i in range(39)])
p = synth.build_problem([big], nrays=2, perturb=True, ncol=2)
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.stat_equil()
ctx.close()
p = synth.tiny_problem(ndepth=1100, nrays=2, ncol=1)
ctx = Context(p)
ctx.formal_sol_gamma_matrices()
ctx.stat_equil()
ctx.close()
print('ray split / large atom / very deep ok', flush=True)

p = synth.config_c3(ncol=600, with_profiles=False, alloc_phi=False)
ctx = Context(p, upload=False)
ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
ctx.update_deps(background=False, profiles_on_device=True)
ctx.formal_sol_gamma_matrices()
ctx.stat_equil()
ctx.close()
print('column stack ok', flush=True)
