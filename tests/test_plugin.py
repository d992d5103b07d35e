"""The Lightweaver plugin shim (liblwb200_plugin.so): the drop-in boundary.

The reference's own compiled core (oracle/_ref, driven by oracle/ref_harness.cpp)
loads the shim through ITS plugin manager
(FsIterationFnsManager::load_fns_from_path, Source/FormalInterface.cpp:62-81) and
calls ITS entry points formal_sol_gamma_matrices / formal_sol / stat_eq
(Source/Lightweaver.hpp:21-32) -- once with the built-in scalar scheme, once with
`mali_full_precond_B200`.  Results must agree to the parity bar."""
import ctypes
import os

import numpy as np
import pytest

from lightweaver_b200 import capi, synth
from oracle import reflib
from tests.util import compare_problems, rel_err

PLUGIN = os.path.join(os.path.dirname(capi.lib_path()), 'liblwb200_plugin.so')
needs_plugin = pytest.mark.skipif(not os.path.exists(PLUGIN), reason='plugin shim not built (needs /root/reference)')


@needs_plugin
def test_plugin_exports_provider_symbols():
    lib = ctypes.CDLL(PLUGIN)
    assert hasattr(lib, 'fs_iteration_fns_provider')
    assert hasattr(lib, 'fs_provider')


@needs_plugin
def test_plugin_carries_no_reference_code():
    """The shim is compiled against the reference headers only: it loads on its own (RTLD_NOW, no
    reference library in the process), exports exactly the two provider symbols and needs no
    library of the reference."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, '-c', f'import ctypes; ctypes.CDLL({PLUGIN!r})'], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(['nm', '-D', '--defined-only', PLUGIN], capture_output=True, text=True)
    if nm.returncode == 0:
        exported = [ln.split()[-1] for ln in nm.stdout.splitlines() if ' T ' in ln]
        assert sorted(exported) == ['fs_iteration_fns_provider', 'fs_provider'], exported
    ldd = subprocess.run(['ldd', PLUGIN], capture_output=True, text=True)
    if ldd.returncode == 0:
        assert 'enkiTS' not in ldd.stdout and 'lwref' not in ldd.stdout and 'oracle' not in ldd.stdout, ldd.stdout


@needs_plugin
@pytest.mark.ref
@pytest.mark.parametrize('toObs', [False, True])
@pytest.mark.parametrize('lowerBc,upperBc', [(capi.BC_THERMALISED, capi.BC_ZERO), (capi.BC_ZERO, capi.BC_THERMALISED)])
def test_own_host_solver_behind_fs_provider_matches_the_reference(toObs, lowerBc, upperBc):
    """fs_provider exports this back end's OWN host Bezier3 solver (the device solver's two-phase
    arithmetic, serial): loaded by the reference's FormalSolverManager and run against the reference's
    piecewise_bezier3_1d on the same rays, optically thin to thick."""
    idx, name = reflib.load_formal_solver(PLUGIN)
    assert name == 'piecewise_bezier3_1d_b200' and idx >= 3
    p = synth.config_c1(nl=0.3)
    rng = np.random.default_rng(7)
    h, T = p.height[0], p.temperature[0]
    for la in range(0, p.Nspect, 23):
        chi = p.chiBg[0, la] * np.exp(rng.normal(0.0, 0.6, p.Nspace)) * 10.0 ** rng.uniform(-1, 3)
        S = synth.planck_nu(p.wavelength[la], T) * np.exp(rng.normal(0.0, 0.4, p.Nspace))
        for mu in (0.11, 0.5, 0.95):
            a = reflib.solve_ray(2, h, T, chi, S, mu, toObs, p.wavelength[la], lowerBc, upperBc)
            b = reflib.solve_ray(idx, h, T, chi, S, mu, toObs, p.wavelength[la], lowerBc, upperBc)
            # (rounding-level agreement; the closed-form Bezier3 coefficients cancel strongly just above
            # dt = 0.05 and amplify the last bit of dt)
            assert rel_err(b[0], a[0]) <= 1e-10, (la, mu)
            assert rel_err(b[1], a[1]) <= 1e-9, (la, mu)
            c = reflib.solve_ray(idx, h, T, chi, S, mu, toObs, p.wavelength[la], lowerBc, upperBc, want_psi=False)
            assert np.array_equal(c[0], b[0])


@needs_plugin
@pytest.mark.ref
def test_reference_plugin_manager_loads_the_scheme():
    import torch
    p = synth.tiny_problem()
    if torch.cuda.is_available():
        r = reflib.RefContext(p, scheme=PLUGIN)
        assert r.scheme_name == 'mali_full_precond_B200'
        r.close()
    else:
        # no GPU: loading works, the first device call raises (no CPU fallback)
        r = reflib.RefContext(p, scheme=PLUGIN)
        assert r.scheme_name == 'mali_full_precond_B200'
        with pytest.raises(RuntimeError):
            r.fs_iter()
        r.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
@pytest.mark.parametrize('solver', [capi.FS_BEZIER3, capi.FS_BESSER, capi.FS_LINEAR])
def test_reference_core_with_b200_scheme_matches_scalar_scheme(solver):
    p = synth.config_c1(formal_solver=solver, nl=0.5)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for it in range(3):
        p.prefill_gamma()
        q.prefill_gamma()
        a = gpu.fs_iter(lambdaIterate=(it == 0))
        b = cpu.fs_iter(lambdaIterate=(it == 0))
        assert abs(a[0] - b[0]) <= 1e-9 * max(b[0], 1.0)
        e = compare_problems(p, q)
        assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['Gamma'] <= 1e-9 and e['R'] <= 1e-9, e
        gpu.stat_eq()
        cpu.stat_eq()
        assert compare_problems(p, q)['n'] <= 1e-8
    # formal_sol (compute_rays path) through the plugin's simple_fs
    p.I[:] = 0.0
    gpu.formal_sol(upOnly=True)
    cpu.formal_sol(upOnly=True)
    assert rel_err(p.I, q.I) <= 1e-9
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_reference_core_hybrid_prd_through_the_b200_scheme():
    """Hybrid PRD: the reference core runs configure_hprd_coeffs itself; the shim flattens ITS tables
    (Spectrum::JCoeffs, hPrdIdxs, Transition::hPrdCoeffs), the device scatters into spect.JRest and
    redistributes in the rest frame.  Against the reference core with its scalar scheme; then the velocity
    field changes, configure_hprd_coeffs runs again and the shim must pick the new tables up."""
    p = synth.tiny_prd_problem(perturb=True)
    p.vlosMu *= 6.0
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    gpu.configure_hprd()
    cpu.configure_hprd()
    for it in range(3):
        if it == 2:
            p.vlosMu *= 1.2
            q.vlosMu *= 1.2
            gpu.configure_hprd()
            cpu.configure_hprd()
        p.prefill_gamma()
        q.prefill_gamma()
        gpu.fs_iter()
        cpu.fs_iter()
        assert rel_err(gpu.jrest(), cpu.jrest()) <= 1e-9 and cpu.jrest().max() > 0.0
        a = gpu.redistribute_prd(maxIter=3, tol=1e-3, nlines=2)
        b = cpu.redistribute_prd(maxIter=3, tol=1e-3, nlines=2)
        assert a['nIter'] == b['nIter']
        n = a['nIter']
        assert rel_err(a['dRho'][:2 * n], b['dRho'][:2 * n]) <= 1e-9
        assert rel_err(a['dJPrdMax'], b['dJPrdMax']) <= 1e-9
        for tp, tq in zip(p.atoms[0].trans, q.atoms[0].trans):
            if tp.rhoPrd is not None:
                assert rel_err(tp.rhoPrd, tq.rhoPrd) <= 1e-9
        assert rel_err(gpu.jrest(), cpu.jrest()) <= 1e-9
        e = compare_problems(p, q)
        assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['R'] <= 1e-9, e
        gpu.stat_eq()
        cpu.stat_eq()
        assert compare_problems(p, q)['n'] <= 1e-8
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_reference_core_redistributes_prd_through_the_b200_scheme():
    """redistribute_prd_lines (Prd.cpp:648-653) dispatches to the scheme's redistribute_prd slot:
    the reference core with our plugin against the reference core with its scalar scheme."""
    p = synth.tiny_prd_problem(perturb=True)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for it in range(2):
        p.prefill_gamma()
        q.prefill_gamma()
        gpu.fs_iter()
        cpu.fs_iter()
        a = gpu.redistribute_prd(maxIter=3, tol=1e-3, nlines=2)
        b = cpu.redistribute_prd(maxIter=3, tol=1e-3, nlines=2)
        assert a['nIter'] == b['nIter']
        n = a['nIter']
        assert rel_err(a['dRho'][:2 * n], b['dRho'][:2 * n]) <= 1e-9
        assert rel_err(a['dJPrdMax'], b['dJPrdMax']) <= 1e-9
        for tp, tq in zip(p.atoms[0].trans, q.atoms[0].trans):
            if tp.rhoPrd is not None:
                assert rel_err(tp.rhoPrd, tq.rhoPrd) <= 1e-9
        e = compare_problems(p, q)
        assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['R'] <= 1e-9, e
        gpu.stat_eq()
        cpu.stat_eq()
        assert compare_problems(p, q)['n'] <= 1e-8
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_reference_core_full_stokes_through_the_b200_scheme():
    """formal_sol_full_stokes dispatches to the scheme's full_stokes_fs slot."""
    from tests.golden.make_golden import polarised_mask
    p = synth.tiny_stokes_problem(perturb=True)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    m = polarised_mask(p)
    for uj, uo in ((False, True), (True, False)):
        a = gpu.full_stokes(updateJ=uj, upOnly=uo)
        b = cpu.full_stokes(updateJ=uj, upOnly=uo)
        assert rel_err(p.I, q.I) <= 1e-9
        assert np.abs(p.Quv[:, :, m] - q.Quv[:, :, m]).max() <= 1e-9 * np.abs(q.I).max()
        if uj:
            assert rel_err(p.J, q.J) <= 1e-9 and abs(a[0] - b[0]) <= 1e-9 * max(b[0], 1.0)
    # the "J20" extra parameter (FormalStokes.cpp:676-681) travels through ExtraParams to the device
    Ja, Jb = np.zeros((p.Nspect, p.Nspace)), np.zeros((p.Nspect, p.Nspace))
    for uj, uo in ((True, False), (True, False), (False, True)):
        gpu.full_stokes(updateJ=uj, upOnly=uo, J20=Ja)
        cpu.full_stokes(updateJ=uj, upOnly=uo, J20=Jb)
        assert rel_err(p.I, q.I) <= 1e-9 and rel_err(p.J, q.J) <= 1e-9
        assert np.abs(p.Quv - q.Quv).max() <= 1e-9 * np.abs(q.I).max()
        assert np.abs(Ja - Jb).max() <= 1e-9 * np.abs(Jb).max() and np.abs(Jb).max() > 0.0
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_reference_core_nr_post_update_through_the_b200_scheme():
    """nr_post_update dispatches to the scheme's slot: Newton-Raphson step with charge conservation."""
    from tests.test_oracle import nr_case
    p = synth.config_c1(nl=0.3)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
        idx, bg, dC, nPrev = nr_case(prob, True, True)
        upd, keep = capi.make_nr_update(idx, bg, dC=dC, nPrev=nPrev, dt=0.05, crswVal=1.0)
        ctx.nr_post_update(upd)
    for a, b in zip(p.atoms, q.atoms):
        assert rel_err(a.n, b.n) <= 1e-7
    assert rel_err(p.ne, q.ne) <= 1e-7
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_reference_core_time_dep_update_through_the_b200_scheme():
    p = synth.tiny_problem(perturb=True)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    nOld = q.atoms[0].n.copy()
    gpu.time_dep_update(0, nOld[0], 0.01)
    cpu.time_dep_update(0, nOld[0], 0.01)
    assert rel_err(p.atoms[0].n, q.atoms[0].n) <= 1e-8
    assert not np.array_equal(q.atoms[0].n, nOld)
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_plugin_sees_in_place_host_mutations():
    """Python mutates buffers in place between calls without telling the plugin
    (update_deps, Ng acceleration): the shim's fingerprints must notice."""
    p = synth.tiny_problem()
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
        ctx.stat_eq()
    # "update_deps": new temperature-dependent background and profiles, populations nudged
    for prob in (p, q):
        prob.chiBg *= 1.07
        prob.etaBg *= 0.93
        t = prob.atoms[0].trans[0]
        t.phi *= 1.01
        t.wphi /= 1.01
        prob.atoms[0].n *= 1.0 + 0.01 * np.linspace(-1, 1, prob.Nspace)
        prob.J *= 1.02
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    e = compare_problems(p, q)
    assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['Gamma'] <= 1e-9, e
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_plugin_singular_matrix_is_a_runtime_error():
    p = synth.tiny_problem()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    p.prefill_gamma()
    gpu.fs_iter()
    p.atoms[0].Gamma[0, 1:, :, 10] = 0.0
    p.atoms[0].n[0, :, 10] = [4.0, 3.0, 2.0, 1.0]  # eliminated row is level 0; rows 1..3 all zero
    with pytest.raises(RuntimeError, match='Singular Matrix'):
        gpu.stat_eq()
    gpu.close()


def _launches():
    return capi.load().lwb200_global_launch_count()


def _collisional_gamma(prob, skew=0.0):
    """Gamma = C (rates skewed away from detailed balance by `skew`) with the diagonal of
    finalise_Gamma: a rate matrix before any radiative term."""
    prob.prefill_gamma()
    for a in prob.atoms:
        if a.detailedStatic:
            continue
        G = a.Gamma
        for i in range(a.Nlevel):
            for j in range(a.Nlevel):
                G[:, i, j] *= 1.0 + skew * (i + 2 * j)
        for i in range(a.Nlevel):
            G[:, i, i] = 0.0
        for i in range(a.Nlevel):
            G[:, i, i] = -G[:, :, i].sum(axis=1)


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_every_plugin_slot_runs_on_the_device():
    """No slot of the scheme has a CPU path behind it: each call of the reference core through the
    plugin launches kernels of liblwb200.so (process-wide launch counter of the C-ABI)."""
    from tests.test_oracle import nr_case
    p = synth.tiny_prd_problem(perturb=True)
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    _collisional_gamma(p)

    def ran(fn, *a, **kw):
        n0 = _launches()
        out = fn(*a, **kw)
        assert _launches() > n0, fn.__name__
        return out
    ran(gpu.stat_eq)                       # before any formal solution: the self-contained device solve
    p.prefill_gamma()
    ran(gpu.fs_iter)                       # fs_iter
    ran(gpu.redistribute_prd, maxIter=2, tol=1e-3, nlines=2)   # redistribute_prd
    ran(gpu.stat_eq)                       # stat_eq
    ran(gpu.formal_sol, upOnly=True)       # simple_fs
    ran(gpu.time_dep_update, 0, p.atoms[0].n[0].copy(), 0.01)   # time_dep_update
    idx, bg, dC, nPrev = nr_case(p, True, True)
    upd, keep = capi.make_nr_update(idx, bg, dC=dC, nPrev=nPrev, dt=0.05, crswVal=1.0)
    ran(gpu.nr_post_update, upd)           # nr_post_update
    gpu.close()
    s = synth.tiny_stokes_problem(perturb=True)
    gs = reflib.RefContext(s, scheme=PLUGIN)
    s.prefill_gamma()
    ran(gs.fs_iter)
    ran(gs.full_stokes, updateJ=True, upOnly=False)   # full_stokes_fs
    gs.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
def test_stat_eq_before_any_formal_solution_matches_the_reference():
    """FsIterationFns::stat_eq on an atom whose Context has not run a formal solution yet (e.g. after the
    escape-probability start): the plugin's self-contained device solve against the reference's stat_eq."""
    p = synth.config_c1(nl=0.3)
    q = p.clone()
    for prob in (p, q):
        _collisional_gamma(prob, skew=0.3)
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    gpu.stat_eq()
    cpu.stat_eq()
    for a, b in zip(p.atoms, q.atoms):
        assert rel_err(a.n, b.n) <= 1e-9
        assert rel_err(b.n, b.nStar) > 1e-3
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
@pytest.mark.parametrize('which', ['both', 'up', 'down'])
def test_zplane_decomposition_through_the_plugin(which):
    """extraParams ZPlaneDecomposition / ZPlaneUp / ZPlaneDown (SimdFullIterationTemplates.hpp:254-281,
    :351-360): the reference core with its scalar scheme against the same call through the B200 scheme,
    for the Gamma iteration and for formal_sol."""
    p = synth.config_c1(nl=0.4)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    shape = (p.Nspect, p.Nrays)
    zs = []
    for ctx in (gpu, cpu):
        up = np.full(shape, -1.0) if which in ('both', 'up') else None
        down = np.full(shape, -1.0) if which in ('both', 'down') else None
        ctx.set_zplane(up, down)
        zs.append((up, down))
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    for a, b in zip(zs[0], zs[1]):
        if a is not None:
            assert (b != -1.0).all() and rel_err(a, b) <= 1e-9
    e = compare_problems(p, q)
    assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['Gamma'] <= 1e-9, e
    for z in zs:
        for a in z:
            if a is not None:
                a[:] = -1.0
    gpu.formal_sol(upOnly=True)
    cpu.formal_sol(upOnly=True)
    for a, b in zip(zs[0], zs[1]):
        if a is not None:
            assert np.array_equal(a == -1.0, b == -1.0)      # upOnly leaves ZPlaneDown untouched
            assert rel_err(a, b) <= 1e-9
    # and switched off again: the arrays are no longer written
    for ctx in (gpu, cpu):
        ctx.set_zplane(None, None)
    for z in zs:
        for a in z:
            if a is not None:
                a[:] = -2.0
    p.prefill_gamma()
    gpu.fs_iter()
    for a in zs[0]:
        if a is not None:
            assert (a == -2.0).all()
    gpu.close()
    cpu.close()


@needs_plugin
@pytest.mark.ref
@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['update_deps', 'by_hand'])
def test_plugin_sees_a_single_depth_change(mode, monkeypatch):
    """A response-function style update: the atmosphere is perturbed at ONE depth and what depends on it
    changes at that depth only -- the background of one wavelength, a profile (with its wphi), J at one
    element.  'update_deps': the temperature changes too (what lw.Context.update_deps follows); the
    atmosphere is hashed over every element and takes background and profiles with it.  'by_hand': only
    the dependent arrays are edited; LWB200_FINGERPRINT=full hashes them over every element as well."""
    if mode == 'by_hand':
        monkeypatch.setenv('LWB200_FINGERPRINT', 'full')
    p = synth.config_c1(nl=0.5)
    q = p.clone()
    gpu = reflib.RefContext(p, scheme=PLUGIN)
    cpu = reflib.RefContext(q, scheme='scalar')
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    k = 37
    for prob in (p, q):
        if mode == 'update_deps':
            prob.temperature[0, k] *= 1.0 + 1e-3
        prob.chiBg[0, prob.Nspect // 3, k] *= 1.5
        prob.etaBg[0, 5, k] *= 0.5
        prob.J[0, 11, k] *= 1.25
        t = prob.atoms[1].trans[0]
        t.phi[0, :, :, :, k] *= 1.1
        t.wphi[0, k] /= 1.1
    for prob, ctx in ((p, gpu), (q, cpu)):
        prob.prefill_gamma()
        ctx.fs_iter()
    e = compare_problems(p, q)
    assert e['I'] <= 1e-9 and e['J'] <= 1e-9 and e['Gamma'] <= 1e-9 and e['R'] <= 1e-9, e
    gpu.close()
    cpu.close()
