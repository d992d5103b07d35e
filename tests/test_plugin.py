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
