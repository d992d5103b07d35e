"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference C++ (oracle/_ref, scalar scheme `mali_full_precond_scalar`, 1 thread,
g++ -O2 x86-64 without FMA contraction) on the seeded synthetic problems of
lightweaver_b200.synth.  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The .npz files are committed; the GPU box (no /root/reference) checks the CUDA
path and the C oracle against them.  Inputs are regenerated from the seeds at
test time; `input_digest` guards against generator drift.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from lightweaver_b200 import capi, synth  # noqa: E402
from oracle import reflib  # noqa: E402


def input_digest(p):
    h = hashlib.sha256()
    for arr in (p.height, p.temperature, p.wavelength, p.chiBg, p.etaBg, p.scaBg, p.muz, p.wmu):
        h.update(np.ascontiguousarray(arr).tobytes())
    for a in p.atoms:
        h.update(a.nStar.tobytes())
        if a.C is not None:
            h.update(a.C.tobytes())
        for t in a.trans:
            h.update(t.wavelength.tobytes())
            if t.phi is not None:
                h.update(t.phi.tobytes())
                h.update(t.wphi.tobytes())
            if t.alpha is not None:
                h.update(t.alpha.tobytes())
            if t.Qelast is not None:
                h.update(t.Qelast.tobytes())
                h.update(t.aDamp.tobytes())
            if t.polProfiles is not None:
                h.update(t.polProfiles.tobytes())
    return h.hexdigest()


CASES = {
    # name: (builder, kwargs, number of iterations, which J rows to keep (None = all))
    'tiny_bezier3': (synth.tiny_problem, dict(formal_solver=capi.FS_BEZIER3), 3, None),
    'tiny_besser': (synth.tiny_problem, dict(formal_solver=capi.FS_BESSER), 3, None),
    'tiny_linear': (synth.tiny_problem, dict(formal_solver=capi.FS_LINEAR), 3, None),
    'tiny_k40_2col': (synth.tiny_problem, dict(ncol=2, ndepth=40, perturb=True), 2, None),
    'tiny_k100': (synth.tiny_problem, dict(ndepth=100, nrays=2), 2, None),
    'c1_bezier3': (synth.config_c1, dict(), 3, 16),
    # configs[1] at its real size: 5 active atoms, 10 rays, ~1e4 wavelengths, up to 3 overlapping
    # cross-atom lines -- the workload the lambda-sharded numbers are quoted on
    'c2_bezier3': (synth.config_c2, dict(), 2, 16),
}

# angle-averaged PRD: every iteration is formal_sol_gamma_matrices, prd_redistribute(maxIter, tol),
# stat_equil (the order of lightweaver/iterate_ctx.py)
PRD_CASES = {
    'tiny_prd': (synth.tiny_prd_problem, dict(perturb=True), 3, None, dict(maxIter=3, tol=1e-3)),
    'c4_prd': (synth.config_c4, dict(), 2, 16, dict(maxIter=3, tol=1e-2)),
    # hybrid PRD (configure_hprd_coeffs, Prd.cpp:697-946): the same flow on columns with a velocity field,
    # rho interpolated per ray to the rest frame, the redistribution fed by JRest
    'tiny_hprd': (synth.tiny_prd_problem, dict(perturb=True, vscale=6.0), 3, None, dict(maxIter=3, tol=1e-3)),
    'c4_hprd': (synth.config_c4, dict(perturb=True), 2, 16, dict(maxIter=3, tol=1e-2)),
}
HYBRID = ('tiny_hprd', 'c4_hprd')


# full Stokes: two Gamma iterations (scalar), then single_stokes_fs(updateJ=False, upOnly=True), then
# formal_sol_full_stokes(updateJ=True, upOnly=False)
STOKES_CASES = {
    'tiny_stokes': (synth.tiny_stokes_problem, dict(perturb=True), 2, None),
    'c5_stokes_1col': (synth.config_c5, dict(ncol=1), 2, 8),
}


def build_case(name):
    if name in STOKES_CASES:
        fn, kw, niter, jstride = STOKES_CASES[name]
        return fn(**kw), niter, jstride
    if name in PRD_CASES:
        fn, kw, niter, jstride, _ = PRD_CASES[name]
    else:
        fn, kw, niter, jstride = CASES[name]
    kw = dict(kw)
    vscale = kw.pop('vscale', None)
    p = fn(**kw)
    if vscale is not None:
        p.vlosMu *= vscale   # (Doppler shifts of several grid points; the profiles stay those of the generator)
    if name in HYBRID:
        # the tables the oracle and the CUDA path run with come from the product's own host routine
        # (lwb200_configure_hprd, checked against the restatement and the reference in tests/test_oracle.py);
        # the reference builds its own when the goldens are made
        p.configure_hprd()
    return p, niter, jstride


def prd_snapshot(p, res):
    snap = {'prd_nIter': np.array(res['nIter']), 'prd_dRho': np.array(res['dRho']),
            'prd_dJ': np.array(res['dJPrdMax']), 'prd_I': p.I.copy(), 'prd_J': p.J.copy()}
    for ia, a in enumerate(p.atoms):
        for it_, t in enumerate(a.trans):
            if t.rhoPrd is not None:
                snap[f'prd_rho{ia}_{it_}'] = t.rhoPrd.copy()
                snap[f'prd_Rij{ia}_{it_}'] = t.Rij.copy()
                snap[f'prd_Rji{ia}_{it_}'] = t.Rji.copy()
    return snap


def run_reference(p, niter, prd=None, hybrid=False):
    """iterate_ctx_se-style: first iteration pure Lambda, then MALI, stat_eq
    after each (lightweaver/iterate_ctx.py:157-176).  Returns per-iteration
    snapshots."""
    snaps = []
    ctxs = [reflib.RefContext(p, col=c) for c in range(p.Ncol)]
    if hybrid:
        for c in ctxs:
            c.configure_hprd()
    for it in range(niter):
        p.prefill_gamma()
        dJ = [c.fs_iter(lambdaIterate=(it == 0)) for c in ctxs]
        snap = {'dJMax': np.array([d[0] for d in dJ]), 'I': p.I.copy(), 'J': p.J.copy()}
        if hybrid:
            snap['JRest'] = np.stack([c.jrest() for c in ctxs])
        for ia, a in enumerate(p.atoms):
            if not a.detailedStatic:
                snap[f'Gamma{ia}'] = a.Gamma.copy()
            for it_, t in enumerate(a.trans):
                snap[f'Rij{ia}_{it_}'] = t.Rij.copy()
                snap[f'Rji{ia}_{it_}'] = t.Rji.copy()
        if prd is not None:
            assert p.Ncol == 1
            nl = sum(1 for a in p.atoms for t in a.trans if t.rhoPrd is not None)
            snap.update(prd_snapshot(p, ctxs[0].redistribute_prd(nlines=nl, **prd)))
            if hybrid:
                snap['prd_JRest'] = np.stack([c.jrest() for c in ctxs])
        for c in ctxs:
            c.stat_eq()
        for ia, a in enumerate(p.atoms):
            snap[f'n{ia}'] = a.n.copy()
        snaps.append(snap)
    for c in ctxs:
        c.close()
    return snaps


def polarised_mask(p):
    """[Nspect] bool: wavelengths where a polarised line is active (where Quv is meaningful; elsewhere
    the reference reports whatever the last polarised ray left in its scratch)."""
    m = np.zeros(p.Nspect, dtype=bool)
    for a in p.atoms:
        for t in a.trans:
            if t.polProfiles is not None:
                m[t.Nblue:t.Nred] = True
    return m


def run_reference_stokes(p, niter):
    snaps = run_reference(p, niter)
    ctx = reflib.RefContext(p)
    out = {}
    ctx.full_stokes(updateJ=False, upOnly=True)
    out['up_I'], out['up_Quv'] = p.I.copy(), p.Quv.copy()
    dJ, _ = ctx.full_stokes(updateJ=True, upOnly=False)
    out['uj_I'], out['uj_Quv'], out['uj_J'], out['uj_dJ'] = p.I.copy(), p.Quv.copy(), p.J.copy(), np.array(dJ)
    ctx.close()
    return snaps, out


def main():
    only = sys.argv[1:]
    for name in list(CASES) + list(PRD_CASES) + list(STOKES_CASES):
        if only and name not in only:
            continue
        p, niter, jstride = build_case(name)
        digest = input_digest(p)
        stokes = None
        if name in STOKES_CASES:
            snaps, stokes = run_reference_stokes(p, niter)
        else:
            snaps = run_reference(p, niter, PRD_CASES[name][4] if name in PRD_CASES else None, name in HYBRID)
        out = {'input_digest': np.array(digest), 'niter': np.array(niter),
               'jstride': np.array(0 if jstride is None else jstride)}
        for it, s in enumerate(snaps):
            for k, v in s.items():
                if k in ('J', 'prd_J') and jstride is not None:
                    v = v[:, ::jstride]
                out[f'it{it}_{k}'] = v
        if stokes is not None:
            for k, v in stokes.items():
                out['stokes_' + k] = v[:, ::jstride] if (k == 'uj_J' and jstride) else v
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'L', p.Nspect, 'K', p.Nspace, 'size %.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
