"""GPU parity tests (run on the B200 box): the CUDA path, called through the
C-ABI, against (a) the golden vectors produced by the reference's own C++ and
(b) the C oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): per-iteration I, J, Gamma <= 1e-9
relative; populations after stat-eq <= 1e-8 per iteration here (they inherit
the conditioning of the rate matrix), converged populations / spectra <= 1e-6.
"""
import numpy as np
import pytest

from lightweaver_b200 import capi, synth
from lightweaver_b200.context import Context, ExplodingMatrixError
from oracle import oraclelib
from tests.golden.make_golden import CASES, PRD_CASES, STOKES_CASES, build_case, input_digest
from tests.test_oracle import check_prd_snapshot, check_snapshot, check_stokes_snapshot, load_golden
from tests.util import compare_problems, gamma_err, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
TOL_N = 1e-8


def oracle_iter(q, lambdaIterate=False, storeDepth=False, stat_eq=True):
    q.prefill_gamma()
    out = []
    for c in range(q.Ncol):
        o = oraclelib.OracleContext(q, col=c)
        out.append(o.fs_iter(lambdaIterate=lambdaIterate, storeDepth=storeDepth))
        if stat_eq:
            o.stat_eq()
    return out


def assert_close(p, q, tol=TOL, tol_n=TOL_N):
    e = compare_problems(p, q)
    assert e['I'] <= tol and e['J'] <= tol and e['Gamma'] <= tol and e['R'] <= tol, e
    assert e['n'] <= tol_n, e
    return e


@pytest.mark.parametrize('name', list(CASES))
def test_cuda_matches_reference_golden(name):
    p, niter, jstride = build_case(name)
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest'])
    ctx = Context(p)
    for it in range(niter):
        upd = ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        assert abs(upd.dJMax - g[f'it{it}_dJMax'].max()) <= 1e-9 * max(upd.dJMax, 1.0)
        check_snapshot(p, g, it, jstride, TOL)
        ctx.stat_equil()
        for ia, a in enumerate(p.atoms):
            assert rel_err(a.n, g[f'it{it}_n{ia}']) <= TOL_N
    ctx.close()


@pytest.mark.parametrize('solver,tileLen', [(capi.FS_BEZIER3, None), (capi.FS_BESSER, None), (capi.FS_LINEAR, None),
                                            (capi.FS_BEZIER3, 7), (capi.FS_BEZIER3, 32)])
def test_cuda_vs_oracle_multicolumn_c1(solver, tileLen, monkeypatch):
    """Config-3-shaped: perturbed FAL C columns with velocity fields, H + Ca II.  tileLen: wavelengths
    per Gamma tile (the planner picks 1 for a problem this small; larger stacks get up to 32)."""
    if tileLen is not None:
        monkeypatch.setenv('LWB200_TILE_LEN', str(tileLen))
        monkeypatch.setenv('LWB200_GTILE_LEN', str(tileLen))
        monkeypatch.setenv('LWB200_GAMMA_DIRECT', '0')   # the tiled Gamma stage of large launches
        monkeypatch.setenv('LWB200_GAMMA_V1', '1' if tileLen == 32 else '0')   # first / second generation
    p = synth.config_c1(ncol=3, perturb=True, formal_solver=solver, nl=0.4)
    q = p.clone()
    ctx = Context(p)
    for it in range(3):
        upd = ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        ref = oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
        dJ = max(r[0] for r in ref)
        assert abs(upd.dJMax - dJ) <= 1e-9 * max(dJ, 1.0)
    ctx.close()


def test_collisional_prefill_on_device_follows_crsw_and_changed_rates():
    """Gamma = crsw * C (LwMiddleLayer.pyx:3198-3203) is made on the device from the resident collisional
    rates: a crsw other than 1, rates changed on the host (fixCollisionalRates=False sends them again), and
    an uploaded prefill taking over again (lwb200_upload(LWB200_GAMMA))."""
    p = synth.tiny_problem(ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)

    def oracle(crsw):
        q.prefill_gamma(crsw)
        for c in range(q.Ncol):
            o = oraclelib.OracleContext(q, col=c)
            o.fs_iter(lambdaIterate=False)
            o.stat_eq()

    ctx.formal_sol_gamma_matrices(crsw=0.37)
    ctx.stat_equil()
    oracle(0.37)
    assert_close(p, q)
    for a, b in zip(p.active_atoms(), q.active_atoms()):
        a.C *= 1.7
        b.C *= 1.7
    ctx.formal_sol_gamma_matrices(fixCollisionalRates=False, crsw=1.0)
    ctx.stat_equil()
    oracle(1.0)
    assert_close(p, q)
    # the device-resident path with an uploaded prefill (what the plugin and the sharded runs use)
    p.prefill_gamma(0.5)
    ctx.upload(capi.ITER_INPUTS)
    ctx.fs_iter_device()
    ctx.download(capi.ITER_OUTPUTS)
    ctx.stat_equil()
    oracle(0.5)
    assert_close(p, q)
    ctx.close()


def test_dj_index_is_argmax_wavelength():
    p = synth.tiny_problem()
    q = p.clone()
    ctx = Context(p)
    for it in range(3):
        upd = ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        (dJ, idx), = oracle_iter(q)
        assert upd.dJMaxIdx == idx
    ctx.close()


def test_formal_sol_up_only_and_full():
    p = synth.tiny_problem(ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    for upOnly in (True, False):
        p.I[:] = -1.0
        ctx.formal_sol(upOnly=upOnly)
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).formal_sol(upOnly=upOnly)
        assert rel_err(p.I, q.I) <= TOL
    ctx.close()


def test_depth_data_store():
    p = synth.tiny_problem(nrays=2)
    p.alloc_depth_data()
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices(extraParams={'storeDepthData': True})
    oracle_iter(q, storeDepth=True, stat_eq=False)
    assert rel_err(p.depthChi, q.depthChi) <= TOL
    assert rel_err(p.depthEta, q.depthEta) <= TOL
    assert rel_err(p.depthI, q.depthI) <= TOL
    ctx.close()


@pytest.mark.parametrize('ndepth', [3, 4, 5, 31, 32, 33, 64, 65, 96, 97, 128])
def test_depth_counts_across_lane_layouts(ndepth):
    """Every register layout (NCH = 1..4), lane-boundary and end-point cases."""
    p = synth.tiny_problem(ndepth=ndepth, nrays=2)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        oracle_iter(q)
        assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('ndepth,solver', [(129, capi.FS_BEZIER3), (200, capi.FS_BEZIER3), (256, capi.FS_BEZIER3),
                                           (257, capi.FS_BESSER), (500, capi.FS_BEZIER3), (500, capi.FS_LINEAR),
                                           (1000, capi.FS_BEZIER3)])
def test_deep_atmospheres_multi_warp_columns(ndepth, solver):
    """Nspace > 128: several warps per wavelength (4 depths per lane), neighbours and the
    scan carry exchanged between warps; warp-boundary and partially filled last warps."""
    p = synth.tiny_problem(ndepth=ndepth, nrays=2, ncol=2, perturb=True, formal_solver=solver)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    for upOnly in (True, False):
        p.I[:] = -1.0
        ctx.formal_sol(upOnly=upOnly)
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).formal_sol(upOnly=upOnly)
        assert rel_err(p.I, q.I) <= TOL
    ctx.close()


def test_deep_static_atmosphere_shares_directions():
    """Nspace > 128 without velocities: the up/down rays of one mu share the direction-independent
    phase of the solver (identical profiles), also across warps."""
    p = synth.tiny_problem(ndepth=300, nrays=3)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        oracle_iter(q)
        assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('solver', [capi.FS_BEZIER3, capi.FS_BESSER, capi.FS_LINEAR])
def test_deep_atmosphere_general_kernel_on_request(solver):
    """Nspace > 128 through the general multi-warp kernel (fs_long_kernel) for every wavelength: Gamma
    iteration and the plain formal solution, every solver."""
    p = synth.tiny_problem(ndepth=260, nrays=2, ncol=2, perturb=True, formal_solver=solver)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0), extraParams={'generalKernel': True})
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('nrays', [16, 17, 20])
def test_many_rays_per_wavelength(nrays):
    """More than 32 rays (2 Nrays) per wavelength: the end points of the rays are evaluated 32 at a time."""
    p = synth.tiny_problem(nrays=nrays, ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('ndepth,solver', [(1025, capi.FS_BEZIER3), (1500, capi.FS_BESSER), (2100, capi.FS_BEZIER3),
                                           (4096, capi.FS_LINEAR)])
def test_very_deep_atmospheres_general_kernel_with_up_to_32_warps(ndepth, solver):
    """1024 < Nspace <= 4096: the general multi-warp kernel (up to 32 warps per column) is the only formal-solution
    kernel; Gamma iteration, stat-eq and the plain formal solution against the oracle."""
    p = synth.tiny_problem(ndepth=ndepth, nrays=2, ncol=2, perturb=True, formal_solver=solver)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    p.I[:] = -1.0
    ctx.formal_sol(upOnly=False)
    for c in range(q.Ncol):
        oraclelib.OracleContext(q, col=c).formal_sol(upOnly=False)
    assert rel_err(p.I, q.I) <= TOL
    ctx.close()


def test_very_deep_prd_atmosphere():
    """Angle-averaged PRD at 1300 depths: Gamma iteration and the PRD sub-iterations through the general kernel."""
    p = synth.tiny_prd_problem(ncol=2, perturb=True, ndepth=1300)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    upd = ctx.prd_redistribute(maxIter=2, tol=1e-6)
    q.prefill_gamma()
    dRho = []
    for c in range(q.Ncol):
        o = oraclelib.OracleContext(q, col=c)
        o.fs_iter()
        dRho.append(o.redistribute_prd(maxIter=2, tol=1e-6, nlines=2)['dRho'][:4])
    assert np.allclose(np.asarray(upd.dRho), np.max(dRho, axis=0), rtol=TOL, atol=1e-12)
    e = compare_problems(p, q)
    assert e['I'] <= TOL and e['J'] <= TOL and e['R'] <= TOL, e
    ctx.close()


def test_too_many_depths_fails_loudly():
    p = synth.tiny_problem(ndepth=4200, nrays=2, with_profiles=False)
    with pytest.raises(capi.LwB200Error):
        Context(p)


def test_thermalised_upper_and_zero_lower_boundaries():
    p = synth.tiny_problem()
    p.lowerBc, p.upperBc = capi.BC_ZERO, capi.BC_THERMALISED
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    assert_close(p, q)
    ctx.close()


def test_callable_boundaries():
    p = synth.tiny_problem(nrays=3)
    rng = np.random.default_rng(5)
    L, M = p.Nspect, p.Nrays
    p.lowerBc = p.upperBc = capi.BC_CALLABLE
    p.lowerBcData = 1e-8 * rng.random((1, L, M))
    p.upperBcData = 1e-10 * rng.random((1, L, M))
    idx = np.full((M, 2), -1, dtype=np.int32)
    idx[:, 1] = np.arange(M)
    p.lowerBcIdx = idx.copy()
    idx2 = np.full((M, 2), -1, dtype=np.int32)
    idx2[:, 0] = np.arange(M)[::-1]
    p.upperBcIdx = idx2
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    assert_close(p, q)
    ctx.close()


def test_detailed_static_atom_gets_rates_but_no_gamma():
    p = synth.build_problem([synth.h6_atom(0.3), synth.ca2_atom(0.3)], nrays=2, detailed=('Ca',))
    assert p.atoms[1].detailedStatic and p.atoms[1].Gamma is None
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        oracle_iter(q)
        assert_close(p, q)
    ctx.close()


def test_angle_averaged_prd_rho_scales_gij():
    p = synth.tiny_problem()
    rng = np.random.default_rng(11)
    t = p.atoms[0].trans[0]
    t.rhoPrd = np.ascontiguousarray(1.0 + 0.3 * rng.random((1, t.Nlambda, p.Nspace)))
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('name', list(PRD_CASES))
def test_cuda_prd_matches_reference_golden(name):
    """formal_sol_gamma_matrices -> prd_redistribute -> stat_equil against the reference's outputs."""
    p, niter, jstride = build_case(name)
    prd = PRD_CASES[name][4]
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest'])
    ctx = Context(p)
    for it in range(niter):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        check_snapshot(p, g, it, jstride, TOL)
        upd = ctx.prd_redistribute(**prd)
        res = dict(nIter=upd.NprdSubIter, dRho=upd.dRho, dJPrdMax=upd.dJPrdMax)
        check_prd_snapshot(p, g, it, jstride, res, TOL)
        ctx.stat_equil()
        for ia, a in enumerate(p.atoms):
            assert rel_err(a.n, g[f'it{it}_n{ia}']) <= TOL_N
    ctx.close()


@pytest.mark.parametrize('ndepth,tiled', [(None, False), (200, False), (None, True), (200, True)])
def test_cuda_prd_vs_oracle_columns(ndepth, tiled, monkeypatch):
    """Angle-averaged PRD on a perturbed two-column stack (also with several warps per column).
    tiled: the Gamma stage of large launches (shared-memory tiles) instead of the per-wavelength one
    a problem this small gets."""
    if tiled:
        monkeypatch.setenv('LWB200_GAMMA_DIRECT', '0')
        monkeypatch.setenv('LWB200_TILE_LEN', '5')
        monkeypatch.setenv('LWB200_GTILE_LEN', '5')
        monkeypatch.setenv('LWB200_GAMMA_V1', '0')
    p = synth.tiny_prd_problem(ncol=2, perturb=True, ndepth=ndepth)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        upd = ctx.prd_redistribute(maxIter=2, tol=1e-6)
        q.prefill_gamma()
        dRho = []
        for c in range(q.Ncol):
            o = oraclelib.OracleContext(q, col=c)
            o.fs_iter()
            dRho.append(o.redistribute_prd(maxIter=2, tol=1e-6, nlines=2)['dRho'][:4])
        assert upd.NprdSubIter == 2
        assert rel_err(np.asarray(upd.dRho), np.max(dRho, axis=0)) <= TOL
        for tp, tq in zip(p.atoms[0].trans, q.atoms[0].trans):
            if tp.rhoPrd is not None:
                assert rel_err(tp.rhoPrd, tq.rhoPrd) <= TOL
        e = compare_problems(p, q)
        assert e['I'] <= TOL and e['J'] <= TOL and e['R'] <= TOL, e
        ctx.stat_equil()
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).stat_eq()
        assert compare_problems(p, q)['n'] <= TOL_N
    ctx.close()


@pytest.mark.parametrize('vscale,ndepth', [(1.0, None), (8.0, None), (3.0, 200)])
def test_cuda_hybrid_prd_vs_oracle_columns(vscale, ndepth):
    """Hybrid PRD on a three-column stack with different velocity fields: rho interpolated per ray to the
    rest frame, JRest scattered by the formal solution, the redistribution fed by JRest and followed by the
    PRD formal solution over each column's own scattering wavelengths -- against the restatement (itself
    bit-identical to the reference, tests/test_oracle.py).  Then new tables for a changed velocity field
    (lwb200_set_hybrid_prd) and one more iteration."""
    p = synth.tiny_prd_problem(ncol=3, perturb=True, ndepth=ndepth)
    p.vlosMu *= vscale
    p.vlosMu[1] *= -0.5
    p.configure_hprd()
    q = p.clone()
    ctx = Context(p)

    def iterate():
        ctx.formal_sol_gamma_matrices()
        q.prefill_gamma()
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).fs_iter()
        assert rel_err(p.hprd.JRest, q.hprd.JRest) <= TOL and q.hprd.JRest.max() > 0.0
        assert_close(p, q)
        upd = ctx.prd_redistribute(maxIter=2, tol=1e-6)
        dRho = [oraclelib.OracleContext(q, col=c).redistribute_prd(maxIter=2, tol=1e-6, nlines=2)['dRho'][:4]
                for c in range(q.Ncol)]
        assert upd.NprdSubIter == 2
        # (dRho is a difference of nearly equal rho's: an absolute 1e-12 on top of the relative bound)
        assert np.allclose(np.asarray(upd.dRho), np.max(dRho, axis=0), rtol=TOL, atol=1e-12)
        for tp, tq in zip(p.atoms[0].trans, q.atoms[0].trans):
            if tp.rhoPrd is not None:
                assert rel_err(tp.rhoPrd, tq.rhoPrd) <= TOL
        assert rel_err(p.hprd.JRest, q.hprd.JRest) <= TOL
        e = compare_problems(p, q)
        assert e['I'] <= TOL and e['J'] <= TOL and e['R'] <= TOL, e
        ctx.stat_equil()
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).stat_eq()
        assert compare_problems(p, q)['n'] <= TOL_N

    for it in range(2):
        iterate()
    # a changed velocity field: same profiles (they are inputs here), new interpolation tables
    p.vlosMu *= 1.3
    q.vlosMu *= 1.3
    ctx.configure_hprd_coeffs()
    q.hprd = oraclelib.configure_hprd(q)
    q.hprd.JRest[...] = p.hprd.JRest
    iterate()
    # tables that scatter from a wavelength the plan did not route through the general kernel are refused ...
    import copy
    import ctypes as C
    bad = copy.deepcopy(p.hprd)
    far = int(np.flatnonzero(~(p.hprd.hPrdLaOfLa >= 0).any(axis=0))[-1])
    assert all(abs(far - la) > 2 for la in np.flatnonzero((p.hprd.hPrdLaOfLa >= 0).any(axis=0)))
    bad.hPrdLaOfLa[0, far] = 0
    keep = []
    assert ctx.lib.lwb200_set_hybrid_prd(ctx._h, bad.c_struct(keep)) != 0
    assert b'create a new context' in ctx.lib.lwb200_last_error()
    # ... and the mirror's answer to that, a fresh context planned for the problem as it is now, carries on
    ctx._rebuild()
    iterate()
    ctx.close()


@pytest.mark.parametrize('name', list(STOKES_CASES))
def test_cuda_stokes_matches_reference_golden(name):
    """Gamma iterations, then single_stokes_fs (up-going rays) and a J-updating full-Stokes pass,
    against the reference's outputs.  Quv compared (relative to max I) where a polarised line is active."""
    p, niter, jstride = build_case(name)
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest'])
    ctx = Context(p)
    for it in range(niter):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
    ctx.single_stokes_fs(updateJ=False, upOnly=True)
    check_stokes_snapshot(p, g, 'up', jstride, TOL, False)
    upd = ctx.single_stokes_fs(updateJ=True, upOnly=False)
    check_stokes_snapshot(p, g, 'uj', jstride, TOL, False)
    assert abs(upd.dJMax - float(g['stokes_uj_dJ'])) <= TOL * max(upd.dJMax, 1.0)
    ctx.close()


@pytest.mark.parametrize('ndepth', [None, 150])
def test_cuda_stokes_vs_oracle_columns(ndepth):
    """Full Stokes on a perturbed magnetised two-column stack (also deeper than one warp)."""
    from tests.golden.make_golden import polarised_mask
    p = synth.tiny_stokes_problem(ncol=2, perturb=True, ndepth=ndepth)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    m = polarised_mask(p)
    for uj, uo in ((False, True), (True, False), (False, False)):
        upd = ctx.single_stokes_fs(updateJ=uj, upOnly=uo)
        dJ = max(oraclelib.OracleContext(q, col=c).full_stokes(updateJ=uj, upOnly=uo)[0] for c in range(q.Ncol))
        assert rel_err(p.I, q.I) <= TOL
        assert np.abs(p.Quv[:, :, m] - q.Quv[:, :, m]).max() <= TOL * np.abs(q.I).max()
        assert np.all(p.Quv[:, :, ~m] == 0.0)
        if uj:
            assert rel_err(p.J, q.J) <= TOL and abs(upd.dJMax - dJ) <= TOL * max(dJ, 1.0)
    assert np.abs(p.Quv).max() > 1e-3 * p.I.max()
    ctx.close()


def test_device_polarised_profiles_match_host_and_feed_the_stokes_solver():
    """lwb200_compute_polarised_profiles (Transition::compute_polarised_profiles, FormalStokes.cpp:9-117) on a
    perturbed magnetised two-column stack: the Voigt AND Faraday-Voigt functions on the device against the host
    profiles (Faddeeva w(z); themselves checked against the reference's routine in tests/test_oracle.py), then a
    context whose polarised lines carry NO host profiles at all runs the full-Stokes formal solution on the
    device-made ones."""
    p = synth.tiny_stokes_problem(ncol=2, perturb=True)
    q = p.clone()
    host = [(t.phi.copy(), t.wphi.copy(), t.polProfiles.copy()) for a in p.atoms for t in a.trans if t.zeeman is not None]
    for a in p.atoms:
        for t in a.trans:
            if t.zeeman is not None:
                t.phi[...] = 0.0
                t.wphi[...] = 0.0
                t.polProfiles[...] = 0.0
    ctx = Context(p)
    ctx.compute_polarised_profiles_device()
    ctx.download(capi.PROFILE | capi.POLPROF)
    dev = [(t.phi, t.wphi, t.polProfiles) for a in p.atoms for t in a.trans if t.zeeman is not None]
    for (phi0, wphi0, pol0), (phi1, wphi1, pol1) in zip(host, dev):
        scale = np.abs(phi0).max()
        assert np.abs(phi1 - phi0).max() <= 1e-12 * scale
        assert np.abs(pol1 - pol0).max() <= 1e-12 * scale
        assert rel_err(wphi1, wphi0) <= 1e-11
    ctx.close()
    # no host profiles at all: the lines are only flagged polarised
    r = q.clone()
    for a in r.atoms:
        for t in a.trans:
            if t.zeeman is not None:
                t.polProfiles = None
                t.polarised = True
    ctx = Context(r)
    ctx.compute_polarised_profiles_device()
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    assert_close(r, q)
    for uj, uo in ((False, True), (True, False)):
        ctx.single_stokes_fs(updateJ=uj, upOnly=uo)
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).full_stokes(updateJ=uj, upOnly=uo)
        from tests.golden.make_golden import polarised_mask
        m = polarised_mask(q)
        assert rel_err(r.I, q.I) <= TOL
        assert np.abs(r.Quv[:, :, m] - q.Quv[:, :, m]).max() <= TOL * np.abs(q.I).max()
    ctx.close()


def test_cuda_stokes_j20_vs_oracle_columns():
    """The 'J20' extra parameter of the full-Stokes formal solution (FormalStokes.cpp:676-681) on a perturbed
    magnetised two-column stack: two J-updating passes (the second scatters the anisotropy the first one built
    into the I and Q emissivities of every wavelength), a pass that does not update J (sees no anisotropy, as in
    the reference), then the option switched off again."""
    from tests.golden.make_golden import polarised_mask
    p = synth.tiny_stokes_problem(ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    m = polarised_mask(p)
    Ja = np.zeros((p.Ncol, p.Nspect, p.Nspace))
    Jb = np.zeros_like(Ja)
    scale = np.abs(q.I).max()
    for n, (uj, uo) in enumerate(((True, False), (True, False), (False, True))):
        upd = ctx.single_stokes_fs(updateJ=uj, upOnly=uo, extraParams={'J20': Ja})
        dJ = max(oraclelib.OracleContext(q, col=c).full_stokes(updateJ=uj, upOnly=uo, J20=Jb)[0] for c in range(q.Ncol))
        assert rel_err(p.I, q.I) <= TOL
        assert np.abs(p.Quv - q.Quv).max() <= TOL * scale
        assert np.abs(Ja - Jb).max() <= TOL * np.abs(Jb).max() and np.abs(Jb).max() > 0.0
        if uj:
            assert rel_err(p.J, q.J) <= TOL and abs(upd.dJMax - dJ) <= TOL * max(dJ, 1.0)
        if n == 1:
            assert np.abs(p.Quv[:, 0][:, ~m]).max() > 0.0   # polarised by the anisotropy alone
    # without the option again: unpolarised wavelengths take the scalar solver, Quv = 0 there
    ctx.single_stokes_fs(updateJ=False, upOnly=True)
    for c in range(q.Ncol):
        oraclelib.OracleContext(q, col=c).full_stokes(updateJ=False, upOnly=True)
    assert rel_err(p.I, q.I) <= TOL and np.all(p.Quv[:, :, ~m] == 0.0)
    ctx.close()


@pytest.mark.parametrize('timeDep,useDC', [(False, False), (True, True)])
def test_nr_post_update_matches_oracle(timeDep, useDC):
    """Newton-Raphson step with charge conservation (nr_post_update_impl): two atoms coupled through
    the electron density, on a perturbed two-column stack."""
    from tests.test_oracle import nr_case
    p = synth.config_c1(nl=0.3, ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    idx, bg, dC, nPrev = nr_case(q, timeDep, useDC)
    td = {'dt': 0.05, 'nPrev': nPrev} if timeDep else None
    ctx.nr_post_update(idx, dC if dC is not None else [], bg, timeDependentData=td)
    upd, keep = capi.make_nr_update(idx, bg, dC=dC, nPrev=nPrev, dt=0.05, crswVal=1.0)
    for c in range(q.Ncol):
        oraclelib.OracleContext(q, col=c).nr_post_update(upd)
    for a, b in zip(p.atoms, q.atoms):
        assert rel_err(a.n, b.n) <= 1e-7     # (sum Nlevel + 1)^2 Newton system: conditioning ~1e8 x rounding
    assert rel_err(p.ne, q.ne) <= 1e-7
    ctx.close()


def test_nr_post_update_with_more_than_63_coupled_levels():
    """The charge-conserving Newton-Raphson step of four 20-level atoms: 81 unknowns per depth."""
    from tests.test_oracle import nr_case
    p = _many_continua_problem(natoms=4)
    assert sum(a.Nlevel for a in p.atoms) + 1 == 81
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    idx, bg, dC, nPrev = nr_case(q, False, True)
    ctx.nr_post_update(idx, dC, bg)
    upd, keep = capi.make_nr_update(idx, bg, dC=dC, nPrev=None, dt=0.05, crswVal=1.0)
    for c in range(q.Ncol):
        oraclelib.OracleContext(q, col=c).nr_post_update(upd)
    for a, b in zip(p.atoms, q.atoms):
        assert rel_err(a.n, b.n) <= 1e-6
    assert rel_err(p.ne, q.ne) <= 1e-6
    ctx.close()


def test_time_dependent_update_matches_oracle():
    """Backward-Euler population step (time_dependent_update_impl) on a perturbed two-column stack."""
    p = synth.tiny_problem(ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    prev = None
    nOld = q.atoms[0].n.copy()
    for dt in (1e-3, 0.05):
        upd, prev = ctx.time_dep_update(dt, prev)
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).time_dep_update(0, nOld, dt)
        assert rel_err(p.atoms[0].n, q.atoms[0].n) <= TOL_N
        assert upd.updatedPops and len(upd.dPops) == 1
    assert np.array_equal(prev[0], nOld)
    ctx.close()


def test_lambda_shards_sum_to_full_iteration():
    """Two wavelength shards with deferred finalise, summed on the host, equal
    the single-context result (the data path of the NCCL all-reduce)."""
    import ctypes as C
    import torch
    p = synth.tiny_problem()
    full = p.clone()
    cf = Context(full)
    cf.formal_sol_gamma_matrices()
    L = p.Nspect
    split = L // 3
    parts = [p.clone(), p.clone()]
    ctxs = [Context(parts[0], laRange=(0, split)), Context(parts[1], laRange=(split, L))]
    bufs = []
    for c, pp in zip(ctxs, parts):
        pp.prefill_gamma()
        c.upload(capi.ITER_INPUTS)
        c.fs_iter_device(deferFinalise=True, want_dJ=False)
        ptr, nbytes = c.device_buffer(capi.BUF_ACCUM)
        host = np.empty(nbytes // 8)
        torch.cuda.synchronize()
        from lightweaver_b200.sharding import copy_device_to_host
        copy_device_to_host(host, ptr, nbytes)
        bufs.append(host)
    total = bufs[0] + bufs[1]
    from lightweaver_b200.sharding import copy_host_to_device
    ptr, nbytes = ctxs[0].device_buffer(capi.BUF_ACCUM)
    copy_host_to_device(ptr, total, nbytes)
    ctxs[0].finalise()
    ctxs[0].download(capi.GAMMA | capi.RATES)
    for a, b in zip(parts[0].atoms, full.atoms):
        assert gamma_err(a.Gamma, b.Gamma) <= 1e-12
        for t, u in zip(a.trans, b.trans):
            assert rel_err(t.Rij, u.Rij, floor=1e-30) <= 1e-12
    # J rows are disjoint per shard; OWN_ROWS brings home only the rows a shard owns
    parts[0].J[:, split:] = -7.0
    parts[0].I[:, split:] = -7.0
    ctxs[0].download(capi.JBAR | capi.INTENS | capi.OWN_ROWS)
    assert np.all(parts[0].J[:, split:] == -7.0) and np.all(parts[0].I[:, split:] == -7.0)
    cf.download(capi.INTENS)
    assert rel_err(parts[0].I[:, :split], full.I[:, :split]) <= 1e-14
    ctxs[1].download(capi.JBAR)
    assert rel_err(parts[0].J[:, :split], full.J[:, :split]) <= 1e-14
    assert rel_err(parts[1].J[:, split:], full.J[:, split:]) <= 1e-14
    d0, i0 = ctxs[0].dj_max()
    d1, i1 = ctxs[1].dj_max()
    df, idf = cf.dj_max()
    assert max(d0, d1) == df
    for c in ctxs + [cf]:
        c.close()


def test_prd_on_a_wavelength_shard_runs_replicated():
    """lwb200_redistribute_prd on a lambda-sharded context (after the accumulators were summed and the
    J rows gathered, as sharding.sharded_prd_redistribute does) equals the un-sharded call."""
    import torch
    from lightweaver_b200.sharding import GpuLambdaShard
    p = synth.tiny_prd_problem(perturb=True)
    full = p.clone()
    cf = Context(full)
    cf.formal_sol_gamma_matrices()
    cf.prd_redistribute(maxIter=2, tol=1e-6)
    L = p.Nspect
    split = L // 2
    parts = [p.clone(), p.clone()]
    ctxs = [Context(parts[0], laRange=(0, split)), Context(parts[1], laRange=(split, L))]
    shards = []
    for c, pp in zip(ctxs, parts):
        pp.prefill_gamma()
        c.upload(capi.ITER_INPUTS | capi.PRD)
        c.fs_iter_device(deferFinalise=True, want_dJ=False)
        shards.append(GpuLambdaShard(c))
    torch.cuda.synchronize()
    total = shards[0].accum_tensor() + shards[1].accum_tensor()     # the all-reduce
    J = torch.cat([shards[0].j_tensor()[:split], shards[1].j_tensor()[split:]])  # the all-gather
    for sh in shards:
        sh.accum_tensor().copy_(total)
        sh.j_tensor().copy_(J)
        sh.finalise()
        assert sh.prd_redistribute(maxIter=2, tol=1e-6) == 2
    for c, pp in zip(ctxs, parts):
        c.download(capi.PRD | capi.JBAR | capi.RATES | capi.INTENS)
        for a, b in zip(pp.atoms, full.atoms):
            for t, u in zip(a.trans, b.trans):
                if t.rhoPrd is not None:
                    assert rel_err(t.rhoPrd, u.rhoPrd) <= TOL
                    assert rel_err(t.Rij, u.Rij, floor=1e-30) <= TOL
        assert rel_err(pp.J, full.J) <= TOL
    for c in ctxs + [cf]:
        c.close()


def test_singular_matrix_raises():
    p = synth.tiny_problem()
    ctx = Context(p)
    ctx.formal_sol_gamma_matrices()
    p.atoms[0].Gamma[0, 2, :, 10] = 0.0  # an all-zero row that is not the eliminated one
    p.atoms[0].Gamma[0, 1, :, 10] = 0.0
    p.atoms[0].Gamma[0, 3, :, 10] = 0.0
    with pytest.raises(ExplodingMatrixError):
        ctx.stat_equil()
    ctx.close()


def test_nine_level_atom_takes_the_general_lu_path():
    # Nlevel = 9 > 7: the per-depth systems go through the local-memory LU instead of the
    # register-resident one; both follow LuSolve.cpp:8-133
    p = synth.nine_level_problem(ncol=2, perturb=True)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    p.atoms[0].Gamma[0, 2, :, 10] = 0.0
    p.atoms[0].Gamma[0, 1, :, 10] = 0.0
    for j in range(3, 9):
        p.atoms[0].Gamma[0, j, :, 10] = 0.0
    with pytest.raises(ExplodingMatrixError):
        ctx.stat_equil()
    ctx.close()


def test_retired_columns_are_skipped_and_keep_their_values():
    # 1.5D stack with a column mask (lwb200_set_active_columns): the reference iterates one Context
    # per column and stops calling the converged ones; here the kernels skip them
    p = synth.tiny_problem(ncol=3, perturb=True)
    q = p.clone()
    ctx = Context(p)

    def oracle_cols(cols, lambdaIterate):
        kept = q.atoms[0].Gamma.copy()
        q.prefill_gamma()
        for c in range(q.Ncol):
            if c not in cols:
                q.atoms[0].Gamma[c] = kept[c]   # a column nobody iterates keeps its Gamma
        dj = 0.0
        for c in cols:
            o = oraclelib.OracleContext(q, col=c)
            dj = max(dj, o.fs_iter(lambdaIterate=lambdaIterate)[0])
            o.stat_eq()
        return dj

    ctx.formal_sol_gamma_matrices(lambdaIterate=True)
    ctx.stat_equil()
    oracle_cols([0, 1, 2], True)
    assert_close(p, q)

    ctx.set_active_columns([True, False, True])
    a = p.atoms[0]
    frozen = [x[1].copy() for x in (p.J, p.I, a.n, a.Gamma)] + [t.Rij[1].copy() for t in a.trans]
    for it in range(2):
        upd = ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        dj = oracle_cols([0, 2], False)
        assert abs(upd.dJMax - dj) <= 1e-9 * dj
        now = [x[1] for x in (p.J, p.I, a.n)] + [t.Rij[1] for t in a.trans]
        for was, cur in zip(frozen[:3] + frozen[4:], now):
            assert np.array_equal(was, cur)
        assert np.array_equal(frozen[3], a.Gamma[1])
        assert_close(p, q)

    with pytest.raises(capi.LwB200Error):
        ctx.prd_redistribute_device()
    ctx.set_active_columns(None)
    ctx.formal_sol_gamma_matrices()
    ctx.stat_equil()
    oracle_cols([0, 1, 2], False)
    assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('opts', [(2, 3, 5), (3, 2, 5), (0, 0, 0), (4, 5, 12), (1, 1, 0)])
def test_ng_acceleration_on_device_matches_oracle(opts):
    # a prescribed sequence of population vectors (six geometric modes around a fixed point) through
    # the device Ng and through the restatement of Ng.hpp (itself bit-identical to the reference's Ng)
    p = synth.tiny_problem(ncol=2)
    a = p.atoms[0]
    ncol, N, K = a.n.shape
    rng = np.random.default_rng(11)
    fix = np.abs(rng.normal(size=(ncol, N * K))) + 1.0
    V = rng.normal(size=(6, ncol, N * K))
    r = np.array([0.95, 0.9, 0.8, 0.7, 0.5, 0.3])
    nIter = 14
    sols = np.array([fix + sum(0.3 * r[m] ** it * V[m] for m in range(6)) for it in range(nIter + 1)])
    want = [oraclelib.ng_run(*opts, sols[:, c, :]) for c in range(ncol)]
    ctx = Context(p)
    a.n[...] = sols[0].reshape(ncol, N, K)
    ctx.upload(capi.POPS)
    ctx.ng_configure(*opts)
    for it in range(nIter):
        a.n[...] = sols[it + 1].reshape(ncol, N, K)
        ctx.upload(capi.POPS)
        acc, dMax, dIdx = ctx.ng_accelerate_device()
        ctx.download(capi.POPS)
        assert acc == bool(want[0][1][it])
        for c in range(ncol):
            # the device sums the normal equations in another order; their conditioning grows with
            # Norder (measured: 4e-10 at Norder = 4), so the tolerance does too
            assert rel_err(a.n[c].reshape(-1), want[c][0][it]) <= (1e-9 if opts[0] <= 3 else 1e-7)
        dm = [want[c][2][it] for c in range(ncol)]
        cbest = int(np.argmax(dm))
        assert abs(dMax[0] - dm[cbest]) <= 1e-9 * max(dm[cbest], 1e-300)
        assert dIdx[0] == cbest * N * K + want[cbest][3][it]
    with pytest.raises(capi.LwB200Error):
        ctx.ng_configure(3, 1, 0)     # the reference reads outside its history here
    ctx.close()


def test_ng_accelerated_iteration_converges_to_the_same_populations():
    def converge(ng):
        p = synth.tiny_problem()
        ctx = Context(p)
        if ng:
            ctx.ng_configure(2, 3, 5)
        nacc = 0
        for it in range(60):
            ctx.formal_sol_gamma_matrices(lambdaIterate=(it < 2))
            upd = ctx.stat_equil()
            nacc += int(upd.ngAccelerated)
            if it > 3 and max(upd.dPops) < 1e-9:
                break
        ctx.close()
        return p.atoms[0].n.copy(), it, nacc
    plain, itPlain, _ = converge(False)
    acc, itAcc, nacc = converge(True)
    assert nacc > 0
    assert rel_err(acc, plain) <= 1e-6
    assert itAcc <= itPlain + 2    # (acceleration must not cost iterations; the order of the fp64 REDs may move either count by one)


def test_device_profiles_match_host_voigt():
    """lwb200_compute_profiles (device Voigt) vs the Faddeeva-based host profiles."""
    p = synth.config_c1(ncol=2, perturb=True, nl=0.3)
    q = p.clone()
    for a in p.atoms:
        for t in a.trans:
            if t.phi is not None:
                t.phi[:] = 0.0
                t.wphi[:] = 0.0
    ctx = Context(p)
    ctx.update_deps(profiles_on_device=True)
    ctx.download(capi.PROFILE)
    for a, b in zip(p.atoms, q.atoms):
        for t, u in zip(a.trans, b.trans):
            if t.phi is not None:
                assert rel_err(t.phi, u.phi) <= 1e-12, t.name
                assert rel_err(t.wphi, u.wphi) <= 1e-12, t.name
    # and the iteration run on device-made profiles agrees with the oracle on host-made ones
    ctx.formal_sol_gamma_matrices()
    oracle_iter(q, stat_eq=False)
    assert_close(p, q)
    ctx.close()


def test_converged_populations_and_spectrum():
    """Run the Gamma iteration to convergence on both sides (iterate_ctx_se
    semantics, lightweaver/iterate_ctx.py:85-88,157-176)."""
    p = synth.tiny_problem()
    q = p.clone()
    ctx = Context(p)
    o = oraclelib.OracleContext(q)
    for it in range(200):
        upd = ctx.formal_sol_gamma_matrices(lambdaIterate=(it < 3))
        pops = ctx.stat_equil()
        q.prefill_gamma()
        o.fs_iter(lambdaIterate=(it < 3))
        o.stat_eq()
        if it > 3 and upd.dJMax < 1e-5 and max(pops.dPops) < 1e-5:
            break
    assert it < 199, 'did not converge'
    assert rel_err(p.atoms[0].n, q.atoms[0].n) <= 1e-6
    assert rel_err(p.I, q.I) <= 1e-6
    ctx.close()


def test_column_stack_properties_at_scale():
    """Size-independent properties on a larger column stack: identical columns
    give identical results, Gamma columns sum to zero, populations are conserved,
    and sampled columns match the oracle."""
    base = synth.config_c1(ncol=4, perturb=True, nl=0.3)
    reps = 16
    big = synth.config_c1(ncol=4 * reps, perturb=False, nl=0.3, with_profiles=False)

    def tile(dst, src):
        dst[...] = np.tile(src, (reps,) + (1,) * (src.ndim - 1))
    for name in ('height', 'temperature', 'chiBg', 'etaBg', 'scaBg', 'vlosMu'):
        tile(getattr(big, name), getattr(base, name))
    for a, b in zip(big.atoms, base.atoms):
        for name in ('n', 'nStar', 'nTotal', 'vBroad', 'C'):
            tile(getattr(a, name), getattr(b, name))
        for t, u in zip(a.trans, b.trans):
            if t.phi is not None:
                tile(t.phi, u.phi)
                tile(t.wphi, u.wphi)
    ctx = Context(big)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
    q = base.clone()
    for it in range(2):
        oracle_iter(q)
    for a, b in zip(big.atoms, q.atoms):
        n = a.n.reshape(reps, 4, *a.n.shape[1:])
        assert rel_err(n, np.broadcast_to(b.n, n.shape)) <= TOL_N
        assert rel_err(n, np.broadcast_to(n[0], n.shape)) <= 1e-11  # replicas agree (atomics reorder sums)
        G = a.Gamma
        assert np.abs(G.sum(axis=1)).max() <= 1e-9 * np.abs(G).max()
        assert rel_err(a.n.sum(axis=1), a.nTotal) <= 1e-12
    I = big.I.reshape(reps, 4, *big.I.shape[1:])
    assert rel_err(I, np.broadcast_to(q.I, I.shape)) <= TOL
    ctx.close()


@pytest.mark.parametrize('ndepth,general', [(None, False), (None, True), (200, False), (200, True)])
def test_zplane_decomposition_matches_depth_intensities(ndepth, general):
    """ZPlaneUp(la, mu) = I(1) of the up-going ray, ZPlaneDown(la, mu) = I(Nz - 2) of the down-going one
    (SimdFullIterationTemplates.hpp:351-360), for a stack, through the moment pipeline, the general
    per-ray kernel and the multi-warp (deep atmosphere) kernel -- against the oracle's depth data."""
    p = synth.tiny_problem(ncol=2, perturb=True, ndepth=ndepth)
    q = p.clone()
    q.alloc_depth_data()
    up = np.full((p.Ncol, p.Nspect, p.Nrays), -1.0)
    down = np.full((p.Ncol, p.Nspect, p.Nrays), -1.0)
    ctx = Context(p)
    ep = {'ZPlaneDecomposition': True, 'ZPlaneUp': up, 'ZPlaneDown': down}
    if general:
        ep['generalKernel'] = True
    oracle_iter(q, storeDepth=True, stat_eq=False)
    # formal_sol(upOnly) from the same state (same J-dagger): only the up-going plane is written
    ctx.formal_sol(upOnly=True, extraParams=ep)
    assert rel_err(up, q.depthI[:, :, :, 1, 1]) <= TOL and (down == -1.0).all()
    up[:] = -1.0
    ctx.formal_sol_gamma_matrices(extraParams=ep)
    assert rel_err(up, q.depthI[:, :, :, 1, 1]) <= TOL
    assert rel_err(down, q.depthI[:, :, :, 0, p.Nspace - 2]) <= TOL
    assert rel_err(p.I, q.I) <= TOL
    ctx.close()


def test_graphed_iteration_matches_oracle():
    """sharding.GraphedIteration: the whole Gamma iteration + stat-eq captured into a CUDA graph and
    replayed (what bench.py times for 1D atmospheres) gives what the direct launches give."""
    import torch
    from lightweaver_b200 import sharding
    p = synth.config_c1(nl=0.4)
    q = p.clone()
    stream = torch.cuda.current_stream()
    ctx = Context(p, stream=stream)
    shard = sharding.GpuLambdaShard(ctx)
    p.prefill_gamma()
    ctx.upload(capi.GAMMA)
    g = sharding.GraphedIteration(shard)        # (its warm-up iterations run on the uploaded state)
    ctx.upload(capi.POPS | capi.JBAR)           # back to the initial state
    for it in range(3):
        g.replay()
        torch.cuda.synchronize()
        ctx.check_singular()
        (dJ, _), = oracle_iter(q)
        assert abs(g.dJMax() - dJ) <= 1e-9 * max(dJ, 1.0)
        ctx.download(capi.ITER_OUTPUTS | capi.POPS)
        assert_close(p, q)
    ctx.close()


def test_short_wavelength_continuum_boltzmann_factor_underflows_to_zero():
    """A bound-free continuum reaching down to 2 nm: exp(-hc / (k lambda T)) underflows (x < -708) at the cool
    depths; the reference's libm exp() gives 0 there and so must the device's table-free exp."""
    lev = [synth.Level(0.0, 2, 0), synth.Level(60000.0, 6, 0), synth.Level(100000.0, 1, 1)]
    lines = [synth.LineSpec(1, 0, 3.0e8, 15, 5.0, 60.0)]
    cont = [synth.ContSpec(2, 0, 6.0e-22, 12, 2.0), synth.ContSpec(2, 1, 1.2e-21, 8, 100.0)]
    toy = synth.ModelAtom('ToyX', 12.0, 1e-4, lev, lines, cont)
    p = synth.build_problem([toy], ncol=2, nrays=3, perturb=True)
    assert (1.4388e7 / (p.wavelength.min() * p.temperature.min())) > 720.0
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
        oracle_iter(q)
        assert_close(p, q)
    ctx.close()


def test_c3_shaped_stack_sampled_columns_match_oracle():
    """configs[2] as bench.py runs it: a stack of >= 1024 perturbed FAL C columns (H + Ca II, 5 rays)
    with profiles MADE ON THE DEVICE, several 512-column batches, the tiled gamma_kernel and the early
    J / I fetch of the public API -- sampled columns against the oracle on host-made (Faddeeva)
    profiles of the same seeded columns."""
    ncol = 1024
    p = synth.config_c3(ncol=ncol, with_profiles=False, alloc_phi=False)
    ctx = Context(p, upload=False)
    ctx.upload(capi.ALL_INPUTS & ~capi.PROFILE)
    ctx.update_deps(background=False, profiles_on_device=True)
    sample = [0, 1, 255, 511, 512, 513, 777, 1023]
    qs = [synth.config_c3(ncol=ncol, col_range=(c, c + 1)) for c in sample]
    for q, c in zip(qs, sample):
        assert np.array_equal(q.temperature[0], p.temperature[c]) and np.array_equal(q.chiBg[0], p.chiBg[c])
    for it in range(2):
        upd = ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        dJref = 0.0
        for q, c in zip(qs, sample):
            (dJ, _), = oracle_iter(q, lambdaIterate=(it == 0))
            dJref = max(dJref, dJ)
            assert rel_err(p.I[c], q.I[0]) <= TOL and rel_err(p.J[c], q.J[0]) <= TOL, (it, c)
            for a, b in zip(p.atoms, q.atoms):
                assert gamma_err(a.Gamma[c], b.Gamma[0]) <= TOL, (it, c, a.name)
                assert rel_err(a.n[c], b.n[0]) <= TOL_N, (it, c, a.name)
                for t, u in zip(a.trans, b.trans):
                    assert rel_err(t.Rij[c], u.Rij[0], floor=1e-30) <= TOL, (it, c, t.name)
                    assert rel_err(t.Rji[c], u.Rji[0], floor=1e-30) <= TOL, (it, c, t.name)
        assert upd.dJMax >= dJref * (1.0 - 1e-9)   # the stack's dJ is the max over ALL its columns
    # populations conserved and Gamma columns sum to zero in every column of the stack
    for a in p.atoms:
        assert np.abs(a.Gamma.sum(axis=1)).max() <= 1e-9 * np.abs(a.Gamma).max()
        assert rel_err(a.n.sum(axis=1), a.nTotal) <= 1e-12
    ctx.close()


def _overlap_problem(nlines, ndepth=None, prd=None):
    """A toy atom whose `nlines` lines all overlap (same-atom cross moments), plus a
    second atom with an overlapping line (cross-atom case)."""
    lev = [synth.Level(0.0, 2, 0)] + [synth.Level(60000.0 + 18.0 * i, 4 + 2 * i, 0) for i in range(nlines)]
    lev.append(synth.Level(100000.0, 1, 1))
    top = len(lev) - 1
    lines = [synth.LineSpec(i + 1, 0, 2.0e8 / (i + 1), 21, 5.0, 80.0) for i in range(nlines)]
    cont = [synth.ContSpec(top, i, 5.0e-22, 8, 50.0) for i in range(top)]
    a = synth.ModelAtom('Ovl', 12.0, 1e-4, lev, lines, cont)
    lev2 = [synth.Level(0.0, 2, 0), synth.Level(60010.0, 6, 0), synth.Level(90000.0, 1, 1)]
    b = synth.ModelAtom('Oth', 20.0, 3e-5, lev2, [synth.LineSpec(1, 0, 1.0e8, 21, 5.0, 80.0)],
                        [synth.ContSpec(2, 0, 5.0e-22, 8, 50.0), synth.ContSpec(2, 1, 8.0e-22, 8, 80.0)])
    return synth.build_problem([a, b], nrays=3, perturb=True, ncol=2, ndepth=ndepth, prd=prd)


def _many_continua_problem(ndepth=None, natoms=2):
    """Two (or more) 20-level atoms whose 19 bound-free continua each all overlap below 100 nm: 38 active
    transitions at those wavelengths."""
    atoms = []
    for name, mass, ab, e0 in (('Big', 12.0, 1e-4, 60000.0), ('Bag', 24.0, 4e-5, 58000.0), ('Bog', 28.0, 3e-5, 57000.0),
                               ('Bug', 32.0, 2e-5, 59000.0))[:natoms]:
        lev = [synth.Level(0.0, 2, 0)] + [synth.Level(e0 + 1500.0 * i, 2 + 2 * (i % 4), 0) for i in range(18)]
        lev.append(synth.Level(100000.0, 1, 1))
        top = len(lev) - 1
        lines = [synth.LineSpec(1, 0, 2.0e8, 15, 4.0, 40.0), synth.LineSpec(5, 0, 5.0e7, 11, 3.0, 30.0)]
        cont = [synth.ContSpec(top, i, 4.0e-22 * (1 + i % 3), 6, 60.0) for i in range(top)]
        atoms.append(synth.ModelAtom(name, mass, ab, lev, lines, cont))
    return synth.build_problem(atoms, nrays=3, perturb=True, ncol=2, ndepth=ndepth)


def _large_atom_problem(nlev=40, ndepth=None):
    """An atom of `nlev` levels (more than the 32 the per-thread LU of the population solve holds) whose
    bound-free continua all overlap below 100 nm, and a small second atom."""
    lev = [synth.Level(0.0, 2, 0)] + [synth.Level(55000.0 + 600.0 * i, 2 + 2 * (i % 4), 0) for i in range(nlev - 2)]
    lev.append(synth.Level(100000.0, 1, 1))
    top = len(lev) - 1
    lines = [synth.LineSpec(1, 0, 2.0e8, 15, 4.0, 40.0), synth.LineSpec(7, 0, 5.0e7, 11, 3.0, 30.0),
             synth.LineSpec(9, 1, 3.0e7, 11, 3.0, 30.0)]
    cont = [synth.ContSpec(top, i, 4.0e-22 * (1 + i % 3), 5, 70.0) for i in range(top)]
    big = synth.ModelAtom('Huge', 12.0, 1e-4, lev, lines, cont)
    lev2 = [synth.Level(0.0, 2, 0), synth.Level(60010.0, 6, 0), synth.Level(90000.0, 1, 1)]
    small = synth.ModelAtom('Oth', 20.0, 3e-5, lev2, [synth.LineSpec(1, 0, 1.0e8, 21, 5.0, 80.0)],
                            [synth.ContSpec(2, 0, 5.0e-22, 8, 50.0), synth.ContSpec(2, 1, 8.0e-22, 8, 80.0)])
    return synth.build_problem([big, small], nrays=3, perturb=True, ncol=2, ndepth=ndepth)


@pytest.mark.parametrize('nlev,ndepth', [(40, None), (64, None), (40, 150)])
def test_atoms_of_more_than_32_levels(nlev, ndepth):
    """33..64 levels: the population solve keeps its matrices in a global scratch block; the 40+ overlapping
    continua of the far UV neither fit the Gamma stage's entry table nor the general kernel's shared-memory
    tile of partial sums, so those wavelengths accumulate with REDs (fs_long_kernel, at any depth)."""
    p = _large_atom_problem(nlev, ndepth)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    # every wavelength through the general kernel on request, and the plain formal solution
    ctx.formal_sol_gamma_matrices(extraParams={'generalKernel': True})
    oracle_iter(q, stat_eq=False)
    e = compare_problems(p, q)
    assert e['I'] <= TOL and e['J'] <= TOL and e['Gamma'] <= TOL and e['R'] <= TOL, e
    ctx.close()


@pytest.mark.parametrize('ndepth', [None, 160])
def test_more_than_32_active_transitions_at_one_wavelength(ndepth):
    """The Gamma stage of the moment pipeline stages at most 32 active transitions per wavelength; the far
    UV of this atom set has 38, and those wavelengths go through the general kernel instead."""
    p = _many_continua_problem(ndepth)
    nact = np.zeros(p.Nspect, dtype=int)
    for a in p.atoms:
        for t in a.trans:
            nact[t.Nblue:t.Nred] += 1
    assert nact.max() > 32
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    ctx.close()


@pytest.mark.parametrize('ndepth', [None, 200])
def test_prd_line_overlapping_three_other_lines(ndepth):
    """A PRD line inside a blend of four lines: its wavelengths are beyond the moment pipeline, so the Gamma
    iteration AND the formal solution of the PRD sub-iterations take the general kernel there (masked to the
    redistributed wavelengths, PRD rates only)."""
    p = _overlap_problem(4, ndepth=ndepth, prd={'Ovl': [0, 2]})
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        upd = ctx.prd_redistribute(maxIter=2, tol=1e-6)
        q.prefill_gamma()
        dRho = []
        for c in range(q.Ncol):
            o = oraclelib.OracleContext(q, col=c)
            o.fs_iter()
            dRho.append(o.redistribute_prd(maxIter=2, tol=1e-6, nlines=2)['dRho'][:4])
        assert upd.NprdSubIter == 2
        assert np.allclose(np.asarray(upd.dRho), np.max(dRho, axis=0), rtol=TOL, atol=1e-12)
        for tp, tq in zip(p.atoms[0].trans, q.atoms[0].trans):
            if tp.rhoPrd is not None:
                assert rel_err(tp.rhoPrd, tq.rhoPrd) <= TOL
        e = compare_problems(p, q)
        assert e['I'] <= TOL and e['J'] <= TOL and e['R'] <= TOL, e
        ctx.stat_equil()
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).stat_eq()
        assert compare_problems(p, q)['n'] <= TOL_N
    ctx.close()


def test_four_overlapping_lines_in_a_deep_atmosphere():
    """More than three overlapping lines with Nspace > 128: those wavelengths go through the general
    multi-warp kernel beside the moment pipeline; then the plain formal solution over both."""
    p = _overlap_problem(4, ndepth=200)
    q = p.clone()
    ctx = Context(p)
    for it in range(2):
        ctx.formal_sol_gamma_matrices(lambdaIterate=(it == 0))
        ctx.stat_equil()
        oracle_iter(q, lambdaIterate=(it == 0))
        assert_close(p, q)
    for upOnly in (True, False):
        p.I[:] = -1.0
        ctx.formal_sol(upOnly=upOnly)
        for c in range(q.Ncol):
            oraclelib.OracleContext(q, col=c).formal_sol(upOnly=upOnly)
        assert rel_err(p.I, q.I) <= TOL
    ctx.close()


@pytest.mark.parametrize('nlines,gammaStage', [(1, None), (2, None), (3, None), (4, None), (1, 'v2'), (2, 'v2'), (3, 'v2'),
                                               (2, 'v2staged'), (3, 'v1')])
def test_overlapping_lines_moment_and_general_kernels(nlines, gammaStage, monkeypatch):
    """Up to 3 overlapping lines go through the moment kernel, more through the general
    kernel; both must match the oracle, and each other.  gammaStage: force the tiled Gamma stage of
    large launches (first generation / second generation / second generation with several warps per
    CTA and the column staged in shared memory) instead of the per-wavelength one."""
    if gammaStage:
        monkeypatch.setenv('LWB200_GAMMA_DIRECT', '0')
        monkeypatch.setenv('LWB200_GTILE_LEN', '6')
        monkeypatch.setenv('LWB200_TILE_LEN', '6')
        monkeypatch.setenv('LWB200_GAMMA_V1', '1' if gammaStage == 'v1' else '0')
        if gammaStage == 'v2staged':
            monkeypatch.setenv('LWB200_GAMMA_WARPS', '4')
            monkeypatch.setenv('LWB200_GAMMA_STAGE', '1')
    p = _overlap_problem(nlines)
    q = p.clone()
    g = p.clone()
    ctx, ctg = Context(p), Context(g)
    for it in range(2):
        ctx.formal_sol_gamma_matrices()
        ctg.formal_sol_gamma_matrices(extraParams={'generalKernel': True})
        ctx.stat_equil()
        ctg.stat_equil()
        oracle_iter(q)
        assert_close(p, q)
        assert_close(g, q)
    ctx.close()
    ctg.close()
