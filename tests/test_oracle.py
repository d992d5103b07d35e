"""CPU tests (no GPU): the C oracle against the golden vectors produced by the
reference's own C++ (tests/golden/make_golden.py) and, where oracle/_ref exists,
against the reference live; plus the C-ABI surface."""
import os
import re

import numpy as np
import pytest

from lightweaver_b200 import capi, synth
from oracle import oraclelib, reflib
from tests.golden.make_golden import CASES, PRD_CASES, STOKES_CASES, build_case, input_digest, polarised_mask
from tests.util import compare_problems, rel_err

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def check_snapshot(p, g, it, jstride, tol, what=('I', 'J', 'Gamma', 'R')):
    errs = {}
    if 'I' in what:
        errs['I'] = rel_err(p.I, g[f'it{it}_I'])
    if 'J' in what:
        J = p.J if not jstride else p.J[:, ::jstride]
        errs['J'] = rel_err(J, g[f'it{it}_J'])
    for ia, a in enumerate(p.atoms):
        if 'Gamma' in what and not a.detailedStatic:
            G, Gr = a.Gamma, g[f'it{it}_Gamma{ia}']
            d = np.abs(G - Gr).max(axis=(-3, -2)) / np.abs(Gr).max(axis=(-3, -2))
            errs[f'Gamma{ia}'] = float(d.max())
        if 'R' in what:
            for it_, t in enumerate(a.trans):
                errs[f'R{ia}_{it_}'] = max(rel_err(t.Rij, g[f'it{it}_Rij{ia}_{it_}'], floor=1e-30),
                                          rel_err(t.Rji, g[f'it{it}_Rji{ia}_{it_}'], floor=1e-30))
    if f'it{it}_JRest' in g.files and 'J' in what:
        errs['JRest'] = rel_err(p.hprd.JRest, g[f'it{it}_JRest'])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f'iteration {it}: {bad}'
    return errs


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_golden(name):
    """The plain-C restatement reproduces the reference's outputs (pins the oracle)."""
    p, niter, jstride = build_case(name)
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest']), 'synthetic input generator drifted; regenerate goldens'
    for it in range(niter):
        p.prefill_gamma()
        for c in range(p.Ncol):
            o = oraclelib.OracleContext(p, col=c)
            dJ, _ = o.fs_iter(lambdaIterate=(it == 0))
            assert abs(dJ - g[f'it{it}_dJMax'][c]) <= 1e-12 * max(dJ, 1.0)
        check_snapshot(p, g, it, jstride, 1e-12)
        for c in range(p.Ncol):
            oraclelib.OracleContext(p, col=c).stat_eq()
        for ia, a in enumerate(p.atoms):
            assert rel_err(a.n, g[f'it{it}_n{ia}']) <= 1e-11


def check_prd_snapshot(p, g, it, jstride, res, tol):
    """State after prd_redistribute against the reference's (tests/golden, PRD cases)."""
    assert res['nIter'] == int(g[f'it{it}_prd_nIter'])
    n = res['nIter']
    nl = len(g[f'it{it}_prd_dRho']) // max(len(g[f'it{it}_prd_dJ']), 1) if n else 0
    errs = {'dRho': rel_err(np.asarray(res['dRho'][:n * nl]), g[f'it{it}_prd_dRho'][:n * nl], floor=1e-30),
            'dJ': rel_err(np.asarray(res['dJPrdMax'][:n]), g[f'it{it}_prd_dJ'][:n], floor=1e-30),
            'I': rel_err(p.I, g[f'it{it}_prd_I'])}
    J = p.J if not jstride else p.J[:, ::jstride]
    errs['J'] = rel_err(J, g[f'it{it}_prd_J'])
    if f'it{it}_prd_JRest' in g.files:
        errs['JRest'] = rel_err(p.hprd.JRest, g[f'it{it}_prd_JRest'])
    for ia, a in enumerate(p.atoms):
        for it_, t in enumerate(a.trans):
            if t.rhoPrd is not None:
                errs[f'rho{ia}_{it_}'] = rel_err(t.rhoPrd, g[f'it{it}_prd_rho{ia}_{it_}'])
                errs[f'R{ia}_{it_}'] = max(rel_err(t.Rij, g[f'it{it}_prd_Rij{ia}_{it_}'], floor=1e-30),
                                          rel_err(t.Rji, g[f'it{it}_prd_Rji{ia}_{it_}'], floor=1e-30))
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f'iteration {it} (PRD): {bad}'
    return errs


@pytest.mark.parametrize('name', list(PRD_CASES))
def test_oracle_prd_matches_reference_golden(name):
    """redistribute_prd_lines: the C restatement against the reference's own outputs."""
    p, niter, jstride = build_case(name)
    prd = PRD_CASES[name][4]
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest']), 'synthetic input generator drifted; regenerate goldens'
    nl = sum(1 for a in p.atoms for t in a.trans if t.rhoPrd is not None)
    o = oraclelib.OracleContext(p)
    for it in range(niter):
        p.prefill_gamma()
        o.fs_iter(lambdaIterate=(it == 0))
        check_snapshot(p, g, it, jstride, 1e-12)
        res = o.redistribute_prd(nlines=nl, **prd)
        check_prd_snapshot(p, g, it, jstride, res, 1e-12)
        o.stat_eq()
        for ia, a in enumerate(p.atoms):
            assert rel_err(a.n, g[f'it{it}_n{ia}']) <= 1e-11


def check_stokes_snapshot(p, g, tag, jstride, tol, quv_everywhere):
    """I, Quv (and J) after a full-Stokes pass against the reference's.  quv_everywhere: also compare
    Quv at wavelengths without a polarised line (bit-level restatements only: the reference leaves
    stale scratch values there, the device returns 0)."""
    errs = {'I': rel_err(p.I, g[f'stokes_{tag}_I'])}
    m = slice(None) if quv_everywhere else polarised_mask(p)
    scale = np.abs(g[f'stokes_{tag}_I']).max()
    errs['Quv'] = float(np.abs(p.Quv[:, :, m] - g[f'stokes_{tag}_Quv'][:, :, m]).max() / scale)
    if tag == 'uj':
        J = p.J if not jstride else p.J[:, ::jstride]
        errs['J'] = rel_err(J, g['stokes_uj_J'])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f'stokes {tag}: {bad}'
    return errs


@pytest.mark.parametrize('name', list(STOKES_CASES))
def test_oracle_stokes_matches_reference_golden(name):
    """formal_sol_full_stokes: the C restatement against the reference's own outputs."""
    p, niter, jstride = build_case(name)
    g = load_golden(name)
    assert input_digest(p) == str(g['input_digest']), 'synthetic input generator drifted; regenerate goldens'
    o = oraclelib.OracleContext(p)
    for it in range(niter):
        p.prefill_gamma()
        o.fs_iter(lambdaIterate=(it == 0))
        o.stat_eq()
    o.full_stokes(updateJ=False, upOnly=True)
    check_stokes_snapshot(p, g, 'up', jstride, 1e-12, True)
    dJ, _ = o.full_stokes(updateJ=True, upOnly=False)
    check_stokes_snapshot(p, g, 'uj', jstride, 1e-12, True)
    assert abs(dJ - float(g['stokes_uj_dJ'])) <= 1e-12 * max(dJ, 1.0)


@pytest.mark.ref
def test_oracle_stokes_vs_reference_live():
    """Bit-level agreement of the full-Stokes restatement with the compiled reference."""
    p = synth.tiny_stokes_problem(ncol=2, perturb=True, nrays=2)
    q = p.clone()
    p.prefill_gamma()
    q.prefill_gamma()
    for c in range(2):
        r, o = reflib.RefContext(p, col=c), oraclelib.OracleContext(q, col=c)
        r.fs_iter()
        o.fs_iter()
        for uj, uo in ((False, True), (True, False)):
            assert r.full_stokes(updateJ=uj, upOnly=uo)[0] == o.full_stokes(updateJ=uj, upOnly=uo)[0]
            assert np.array_equal(p.I[c], q.I[c]) and np.array_equal(p.Quv[c], q.Quv[c])
            assert np.array_equal(p.J[c], q.J[c])
        r.close()
    assert np.abs(p.Quv).max() > 1e-3 * p.I.max()


def nr_case(pr, timeDep, useDC):
    """A synthetic Newton-Raphson update for every active atom of problem pr."""
    idx = [i for i, a in enumerate(pr.atoms) if not a.detailedStatic]
    bg = np.ascontiguousarray(pr.ne - 0.9 * sum((pr.atoms[i].stages[None, :, None] * pr.atoms[i].n).sum(axis=1) for i in idx))
    dC = [1e-3 * pr.atoms[i].C / pr.ne[:, None, None, :] for i in idx] if useDC else None
    nPrev = ([pr.atoms[i].n * (1.0 + 0.01 * np.sin(np.arange(pr.Nspace)))[None, None, :] for i in idx]
             if timeDep else None)
    return idx, bg, dC, nPrev


@pytest.mark.ref
@pytest.mark.parametrize('timeDep,useDC', [(False, False), (False, True), (True, False), (True, True)])
def test_oracle_nr_post_update_vs_reference(timeDep, useDC):
    """nr_post_update_impl (charge-conserving Newton-Raphson step): restatement against the compiled
    reference, bit level, two atoms coupled through ne."""
    p = synth.config_c1(nl=0.3)
    q = p.clone()
    r, o = reflib.RefContext(p), oraclelib.OracleContext(q)
    for pr, ctx in ((p, r), (q, o)):
        pr.prefill_gamma()
        ctx.fs_iter()
        idx, bg, dC, nPrev = nr_case(pr, timeDep, useDC)
        upd, keep = capi.make_nr_update(idx, bg, dC=dC, nPrev=nPrev, dt=0.05, crswVal=1.0)
        ctx.nr_post_update(upd)
    for a, b in zip(p.atoms, q.atoms):
        assert np.array_equal(a.n, b.n)
    assert np.array_equal(p.ne, q.ne) and np.all(np.isfinite(p.ne))
    r.close()


@pytest.mark.ref
def test_oracle_time_dep_update_vs_reference():
    """time_dependent_update_impl: restatement against the compiled reference (bit level)."""
    p = synth.tiny_problem(perturb=True)
    q = p.clone()
    r, o = reflib.RefContext(p), oraclelib.OracleContext(q)
    p.prefill_gamma()
    q.prefill_gamma()
    r.fs_iter()
    o.fs_iter()
    nOld = p.atoms[0].n.copy()
    for dt in (1e-3, 0.1):
        r.time_dep_update(0, nOld[0], dt)
        o.time_dep_update(0, nOld, dt)
        assert np.array_equal(p.atoms[0].n, q.atoms[0].n)
        assert np.all(np.isfinite(p.atoms[0].n)) and not np.array_equal(p.atoms[0].n, nOld)
    r.close()


@pytest.mark.ref
def test_oracle_prd_vs_reference_live():
    """Bit-level agreement of the PRD restatement with the compiled reference."""
    p = synth.tiny_prd_problem(nrays=2, ndepth=50)
    q = p.clone()
    r, o = reflib.RefContext(p), oraclelib.OracleContext(q)
    for it in range(2):
        p.prefill_gamma()
        q.prefill_gamma()
        r.fs_iter()
        o.fs_iter()
        a, b = r.redistribute_prd(maxIter=4, tol=1e-4, nlines=2), o.redistribute_prd(maxIter=4, tol=1e-4, nlines=2)
        assert a['nIter'] == b['nIter']
        assert np.array_equal(a['dRho'], b['dRho']) and np.array_equal(a['dJPrdMax'], b['dJPrdMax'])
        for ta, tb in zip(p.atoms[0].trans, q.atoms[0].trans):
            if ta.rhoPrd is not None:
                assert np.array_equal(ta.rhoPrd, tb.rhoPrd)
        assert max(compare_problems(q, p).values()) <= 1e-14
        r.stat_eq()
        o.stat_eq()
    r.close()


@pytest.mark.parametrize('includeDetailed', [False, True])
def test_product_hprd_tables_match_the_restatement(includeDetailed):
    """lwb200_configure_hprd (host code of the CUDA library, organised by binary searches) builds the tables
    of the restated configure_hprd_coeffs element for element: a three-column stack with different velocity
    fields, Doppler shifts of several grid points, one atom made detailed-static."""
    p = synth.tiny_prd_problem(nrays=2, ndepth=40, perturb=True, ncol=3)
    p.vlosMu *= 8.0
    p.vlosMu[1] *= -12.0
    p.atoms[0].detailedStatic = includeDetailed
    a = oraclelib.configure_hprd(p, includeDetailed)
    b = p.configure_hprd(includeDetailed)
    assert (a.NprdLa, a.NhPrd) == (b.NprdLa, b.NhPrd) and a.NhPrd > a.NprdLa > 0
    for name in HPRD_TABLES:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert not np.array_equal(a.column(0, p).JCoeffFrac, a.column(1, p).JCoeffFrac)
    c1 = a.column(1, p)
    one = p.column(1)
    assert np.array_equal(c1.JCoeffFrac, oraclelib.configure_hprd(one, includeDetailed).JCoeffFrac)


@pytest.mark.ref
def test_polarised_profile_generator_vs_reference():
    """The host formula the GPU tests check device-made polarised profiles against (synth.polarised_profiles:
    scipy's Faddeeva w(z), a normal Zeeman triplet) is the reference's own Transition::compute_polarised_profiles
    (FormalStokes.cpp:9-117, its vendored Faddeeva package) to rounding: phi, wphi and the six extra profiles."""
    p = synth.tiny_stokes_problem(perturb=True)
    want = [(t.phi.copy(), t.wphi.copy(), t.polProfiles.copy()) for a in p.atoms for t in a.trans if t.zeeman is not None]
    assert want
    for a in p.atoms:
        for t in a.trans:
            if t.zeeman is not None:
                t.phi[...] = 0.0
                t.wphi[...] = 0.0
                t.polProfiles[...] = 0.0
    r = reflib.RefContext(p)
    r.compute_polarised_profiles()
    got = [(t.phi, t.wphi, t.polProfiles) for a in p.atoms for t in a.trans if t.zeeman is not None]
    for (phi0, wphi0, pol0), (phi1, wphi1, pol1) in zip(want, got):
        scale = np.abs(phi0).max()
        assert np.abs(phi1 - phi0).max() <= 1e-13 * scale
        assert np.abs(pol1 - pol0).max() <= 1e-13 * scale and np.abs(pol0[3:]).max() > 1e-3 * scale
        assert rel_err(wphi1, wphi0) <= 1e-12
    r.close()


@pytest.mark.ref
def test_oracle_stokes_j20_vs_reference_live():
    """The 'J20' extra parameter of the full-Stokes formal solution (FormalStokes.cpp:433-437, :469-471, :575-583,
    :642-648): two J-updating passes (the second one scatters the anisotropy the first one built into the I and Q
    emissivities of EVERY wavelength) and a pass that does not update J, restatement against the compiled reference."""
    p = synth.tiny_stokes_problem(perturb=True)
    q = p.clone()
    r, o = reflib.RefContext(p), oraclelib.OracleContext(q)
    for it in range(2):
        p.prefill_gamma()
        q.prefill_gamma()
        r.fs_iter()
        o.fs_iter()
        r.stat_eq()
        o.stat_eq()
    pol = np.zeros(p.Nspect, dtype=bool)
    for a_ in p.atoms:
        for t in a_.trans:
            if t.polProfiles is not None:
                pol[t.Nblue:t.Nred] = True
    Ja, Jb = np.zeros((1, p.Nspect, p.Nspace)), np.zeros((1, p.Nspect, p.Nspace))
    for n, (updateJ, upOnly) in enumerate(((True, False), (True, False), (False, True))):
        a = r.full_stokes(updateJ=updateJ, upOnly=upOnly, J20=Ja[0])
        b = o.full_stokes(updateJ=updateJ, upOnly=upOnly, J20=Jb)
        assert a[0] == b[0]
        assert np.array_equal(Ja, Jb) and np.abs(Ja).max() > 0.0
        assert np.array_equal(p.I, q.I) and np.array_equal(p.J, q.J)
        assert np.array_equal(p.Quv, q.Quv) and (np.abs(p.Quv) > 0).any()
        if n == 1:
            # the anisotropy of the first pass polarises wavelengths no polarised line touches
            assert np.abs(p.Quv[0, 0][~pol]).max() > 0.0
    r.close()


HPRD_TABLES = ('prdLaOfLa', 'hPrdLaOfLa', 'JCoeffOff', 'JCoeffIdx', 'JCoeffFrac', 'lineAtom', 'lineTrans',
               'rhoCoefOff', 'rhoFrac', 'rhoI0')


@pytest.mark.ref
@pytest.mark.parametrize('vscale', [1.0, 8.0])
def test_oracle_hybrid_prd_vs_reference_live(vscale):
    """Hybrid PRD: the restated configure_hprd_coeffs builds the reference's own tables element for element,
    and the formal solution (rho interpolated per ray, JRest scattered), the rest-frame redistribution and
    the PRD formal solution over hPrdIdxs agree with the compiled reference to the bit.  vscale = 8 makes the
    Doppler shifts larger than the core wavelength spacing (several entries per coefficient list)."""
    p = synth.tiny_prd_problem(nrays=2, ndepth=50, perturb=True)
    p.vlosMu *= vscale
    q = p.clone()
    r = reflib.RefContext(p)
    href = r.configure_hprd()
    q.hprd = oraclelib.configure_hprd(q)
    assert (href.NprdLa, href.NhPrd) == (q.hprd.NprdLa, q.hprd.NhPrd)
    for name in HPRD_TABLES:
        assert np.array_equal(getattr(href, name), getattr(q.hprd, name)), name
    assert href.NhPrd >= href.NprdLa > 0 and len(href.JCoeffIdx) > 0
    o = oraclelib.OracleContext(q)
    for it in range(2):
        p.prefill_gamma()
        q.prefill_gamma()
        r.fs_iter()
        o.fs_iter()
        assert np.array_equal(r.jrest(), q.hprd.JRest[0]) and q.hprd.JRest.max() > 0.0
        assert max(compare_problems(q, p).values()) <= 1e-14
        a, b = r.redistribute_prd(maxIter=4, tol=1e-4, nlines=2), o.redistribute_prd(maxIter=4, tol=1e-4, nlines=2)
        assert a['nIter'] == b['nIter']
        assert np.array_equal(a['dRho'], b['dRho']) and np.array_equal(a['dJPrdMax'], b['dJPrdMax'])
        for ta, tb in zip(p.atoms[0].trans, q.atoms[0].trans):
            if ta.rhoPrd is not None:
                assert np.array_equal(ta.rhoPrd, tb.rhoPrd)
        assert np.array_equal(r.jrest(), q.hprd.JRest[0])
        assert max(compare_problems(q, p).values()) <= 1e-14
        r.stat_eq()
        o.stat_eq()
    r.close()


@pytest.mark.ref
@pytest.mark.parametrize('solver', [0, 1, 2])
def test_oracle_solvers_vs_reference_rays(solver):
    """Solver-level known answers: random rays through the reference's own
    LwFsFn solvers vs the restatement, both directions, both boundary types."""
    rng = np.random.default_rng(7 + solver)
    atm = synth.falc_columns(1)
    h, T = atm['height'][0], atm['temperature'][0]
    K = h.shape[0]
    for trial in range(20):
        chi = 10.0**(np.linspace(-9, -1, K) + 0.3 * rng.standard_normal(K))
        S = 1e-8 * (1.0 + np.linspace(0, 3, K) + 0.2 * rng.random(K))
        mu = rng.uniform(0.05, 1.0)
        for toObs in (0, 1):
            for lbc, ubc in ((capi.BC_THERMALISED, capi.BC_ZERO), (capi.BC_ZERO, capi.BC_THERMALISED)):
                Ir, Pr = reflib.solve_ray(solver, h, T, chi, S, mu, toObs, 500.0, lbc, ubc)
                Io, Po = oraclelib.solve_ray(solver, h, T, chi, S, mu, toObs, 500.0, lbc, ubc)
                assert rel_err(Io, Ir) <= 1e-14 and rel_err(Po, Pr) <= 1e-14


@pytest.mark.ref
def test_oracle_lu_vs_reference():
    rng = np.random.default_rng(3)
    for N in (2, 3, 6, 11):
        for trial in range(20):
            A = rng.standard_normal((N, N)) * 10.0**rng.uniform(-6, 6, (N, 1))
            b = rng.standard_normal(N)
            xr = reflib.solve_lin_eq(A, b)
            xo = oraclelib.solve_lin_eq(A, b)
            assert np.array_equal(xr, xo)
    A = np.ones((3, 3))
    A[1] = 0.0
    with pytest.raises(RuntimeError):
        oraclelib.solve_lin_eq(A, np.ones(3))
    with pytest.raises(RuntimeError):
        reflib.solve_lin_eq(A, np.ones(3))


@pytest.mark.ref
def test_oracle_vs_reference_live_full_iteration():
    """Bit-level agreement on a multi-column perturbed problem, incl. depth data."""
    p = synth.tiny_problem(ncol=2, perturb=True, nrays=2)
    p.alloc_depth_data()
    q = p.clone()
    for it in range(2):
        p.prefill_gamma()
        q.prefill_gamma()
        for c in range(2):
            r = reflib.RefContext(p, col=c)
            r.set_depth_fill(True)
            a = r.fs_iter(lambdaIterate=(it == 0))
            r.stat_eq()
            r.close()
            o = oraclelib.OracleContext(q, col=c)
            b = o.fs_iter(lambdaIterate=(it == 0), storeDepth=True, serial_idx=True)
            o.stat_eq()
            assert a == b
        e = compare_problems(q, p)
        assert max(e.values()) <= 1e-14, e
        assert rel_err(q.depthI, p.depthI) <= 1e-14 and rel_err(q.depthChi, p.depthChi) <= 1e-14
        assert rel_err(q.depthEta, p.depthEta) <= 1e-14


@pytest.mark.ref
def test_reference_simd_schemes_load_and_agree_away_from_tail():
    """The reference's SIMD plugins (the timing baseline) load through its own
    plugin manager; they agree with scalar except at the last Nspace % stride
    depths -- which is why scalar is the oracle."""
    schemes = [s for s in reflib.usable_schemes() if s != 'scalar']
    if not schemes:
        pytest.skip('no SIMD scheme usable on this CPU')
    p = synth.tiny_problem()
    q = p.clone()
    reflib.RefContext(p).fs_iter()
    r = reflib.RefContext(q, scheme=schemes[-1])
    assert schemes[-1] in r.scheme_name
    r.fs_iter()
    # the tail mismatch changes chi at the deepest points, which feeds back into J
    # everywhere along up-going rays: only loose agreement can be expected
    assert rel_err(q.J, p.J) <= 1e-4
    assert rel_err(q.J, p.J) > 1e-12, 'SIMD scheme unexpectedly identical to scalar'


def test_profiles_match_reference_voigt():
    """scipy's wofz (used by the synthetic generator) is the Faddeeva package the
    reference vendors: phi/wphi from the reference's compute_phi agree."""
    if not reflib.available():
        pytest.skip('oracle/_ref not built')
    p = synth.tiny_problem(ncol=1, perturb=True)
    q = p.clone()
    for a in q.atoms:
        for t in a.trans:
            if t.phi is not None:
                t.phi[:] = 0.0
                t.wphi[:] = 0.0
    r = reflib.RefContext(q)
    r.compute_profiles()
    for a, b in zip(p.atoms, q.atoms):
        for t, u in zip(a.trans, b.trans):
            if t.phi is not None:
                assert rel_err(t.phi, u.phi) <= 1e-13
                assert rel_err(t.wphi, u.wphi) <= 1e-13


def test_cabi_library_exports_every_declared_symbol():
    """liblwb200.so loads without a GPU and exports what include/lwb200.h declares."""
    hdr = open(os.path.join(ROOT, 'include', 'lwb200.h')).read()
    declared = sorted(set(re.findall(r'\b(lwb200_[a-z_0-9]+)\s*\(', hdr)))
    assert declared == sorted(capi.EXPORTED_SYMBOLS)
    lib = capi.load()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.lwb200_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header():
    """ctypes mirrors of the POD structs have the C sizes (LP64)."""
    import ctypes as C
    assert C.sizeof(capi.LwB200Transition) == 6 * 4 + 5 * 8 + 10 * 8
    assert C.sizeof(capi.LwB200Atom) == 4 * 4 + 8 * 8
    assert C.sizeof(capi.LwB200Problem) == 12 * 4 + 22 * 8
    assert C.sizeof(capi.LwB200HybridPrd) == 4 * 4 + 11 * 8


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a device the product path raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from lightweaver_b200.context import Context
    with pytest.raises(capi.LwB200Error):
        Context(synth.tiny_problem())


@pytest.mark.ref
@pytest.mark.parametrize('opts', [(2, 3, 5), (3, 2, 5), (0, 0, 0), (2, 2, 4), (4, 5, 12), (1, 1, 0)])
def test_oracle_ng_vs_reference(opts):
    """lwo_ng_run against the reference's own Ng object (Ng.hpp) over a prescribed sequence."""
    rng = np.random.default_rng(5)
    n = 60
    fix = np.abs(rng.normal(size=n)) + 1.0
    V = rng.normal(size=(6, n))
    r = np.array([0.95, 0.9, 0.8, 0.7, 0.5, 0.3])
    sols = np.array([fix + sum(0.3 * r[m] ** it * V[m] for m in range(6)) for it in range(15)])
    a = oraclelib.ng_run(*opts, sols)
    b = reflib.ng_run(*opts, sols)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    if opts[0] > 0:
        assert a[1].sum() > 0
