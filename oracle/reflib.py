"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/_ref/liblwref.so, the
UNMODIFIED reference C++ (built by oracle/Makefile) driven through
oracle/ref_harness.cpp on an LwB200Problem."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
_lib = None


def available():
    return os.path.exists(os.path.join(REF_DIR, 'liblwref.so'))


def cpu_flags():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('flags'):
                    return set(line.split(':', 1)[1].split())
    except OSError:
        pass
    return set()


def usable_schemes():
    """Reference iteration schemes this host CPU can run (cf. the reference's
    lightweaver/simd_management.py:27-42)."""
    flags = cpu_flags()
    out = ['scalar']
    if 'sse2' in flags and os.path.exists(os.path.join(REF_DIR, 'SimdImpl_SSE2.so')):
        out.append('SSE2')
    if {'avx2', 'fma'} <= flags and os.path.exists(os.path.join(REF_DIR, 'SimdImpl_AVX2FMA.so')):
        out.append('AVX2FMA')
    if {'avx512f', 'avx512dq'} <= flags and os.path.exists(os.path.join(REF_DIR, 'SimdImpl_AVX512.so')):
        out.append('AVX512')
    return out


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(REF_DIR, 'liblwref.so')
    if not os.path.exists(path):
        raise RuntimeError(f'{path} not built: run `make -C oracle ref` where /root/reference exists')
    lib = C.CDLL(path)
    vp = C.c_void_p
    dp = C.POINTER(C.c_double)
    lib.lwref_last_error.restype = C.c_char_p
    lib.lwref_create.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.lwref_destroy.argtypes = [vp]
    lib.lwref_destroy.restype = None
    lib.lwref_scheme_name.argtypes = [vp]
    lib.lwref_scheme_name.restype = C.c_char_p
    lib.lwref_set_depth_fill.argtypes = [vp, C.c_int]
    lib.lwref_fs_iter.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int64)]
    lib.lwref_formal_sol.argtypes = [vp, C.c_int]
    lib.lwref_stat_eq.argtypes = [vp]
    lib.lwref_full_stokes.argtypes = [vp, C.c_int, C.c_int, dp, C.POINTER(C.c_int64)]
    lib.lwref_full_stokes_j20.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int64)]
    lib.lwref_nr_post_update.argtypes = [vp, vp]
    lib.lwref_time_dep_update.argtypes = [vp, C.c_int, dp, C.c_double]
    lib.lwref_redistribute_prd.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), dp,
                                           C.POINTER(C.c_int), dp, C.POINTER(C.c_int64)]
    lib.lwref_configure_hprd.argtypes = [vp, C.c_int]
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.lwref_hprd_export.argtypes = [vp, C.c_int, i64p, i32p, i32p, i64p, i32p, dp, dp, i32p]
    lib.lwref_get_jrest.argtypes = [vp, dp]
    lib.lwref_compute_profiles.argtypes = [vp]
    lib.lwref_compute_polarised_profiles.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, i32p, dp, dp]
    lib.lwref_time_fs_iter.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    lib.lwref_solve_ray.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int,
                                    C.c_double, C.c_int, C.c_int, dp, dp]
    lib.lwref_solve_lin_eq.argtypes = [C.c_int, dp, dp, C.c_int]
    lib.lwref_set_zplane.argtypes = [vp, dp, dp]
    lib.lwref_load_formal_solver.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_int]
    lib.lwref_ng_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int), dp,
                                  C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(load().lwref_last_error().decode())


class RefContext:
    """The reference's Context over column ``col`` of a Problem.  Results are
    written in place into the Problem's numpy buffers, as the reference does."""

    def __init__(self, problem, col=0, scheme='scalar', Nthreads=1):
        self.lib = load()
        self.problem = problem
        self._cs = problem.c_struct()
        self.h = C.c_void_p()
        _check(self.lib.lwref_create(C.byref(self._cs), col, scheme.encode(), Nthreads, C.byref(self.h)))

    @property
    def scheme_name(self):
        return self.lib.lwref_scheme_name(self.h).decode()

    def fs_iter(self, lambdaIterate=False):
        dJ = C.c_double()
        idx = C.c_int64()
        _check(self.lib.lwref_fs_iter(self.h, int(lambdaIterate), C.byref(dJ), C.byref(idx)))
        return dJ.value, idx.value

    def formal_sol(self, upOnly=True):
        _check(self.lib.lwref_formal_sol(self.h, int(upOnly)))

    def set_zplane(self, up=None, down=None):
        """ZPlaneDecomposition extra parameters for the following fs_iter / formal_sol calls: float64
        arrays [Nspect, Nrays] (or None)."""
        dp = C.POINTER(C.c_double)
        self._zkeep = (up, down)
        _check(self.lib.lwref_set_zplane(self.h, up.ctypes.data_as(dp) if up is not None else dp(),
                                         down.ctypes.data_as(dp) if down is not None else dp()))

    def stat_eq(self):
        _check(self.lib.lwref_stat_eq(self.h))

    def configure_hprd(self, includeDetailed=False):
        """The reference's own configure_hprd_coeffs + update_threads on this column.  Returns its tables as a
        one-column lightweaver_b200.problem.HybridPrd (JRest zeroed; read the reference's with jrest())."""
        import numpy as np
        from lightweaver_b200.problem import HybridPrd
        _check(self.lib.lwref_configure_hprd(self.h, int(includeDetailed)))
        p = self.problem
        K, M, L = p.Nspace, p.Nrays, p.Nspect
        i32p, i64p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        counts = np.zeros(4, dtype=np.int64)
        args0 = [i32p(), i32p(), i64p(), i32p(), dp(), dp(), i32p()]
        _check(self.lib.lwref_hprd_export(self.h, int(includeDetailed), counts.ctypes.data_as(i64p), *args0))
        NprdLa, NhPrd, nnz, nl = (int(x) for x in counts)
        lines = [(ia, it) for pass_ in ((False, True) if includeDetailed else (False,))
                 for ia, a in enumerate(p.atoms) if bool(a.detailedStatic) == pass_
                 for it, t in enumerate(a.trans) if t.rhoPrd is not None]
        assert len(lines) == nl
        tot = sum(p.atoms[a].trans[t].Nlambda * M * 2 * K for a, t in lines)
        prdLa, hPrdLa = np.zeros(L, np.int32), np.zeros(L, np.int32)
        off = np.zeros(NhPrd * M * 2 * K + 1, np.int64)
        cIdx, cFrac = np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1))
        rFrac, rI0 = np.zeros(max(tot, 1)), np.zeros(max(tot, 1), np.int32)
        _check(self.lib.lwref_hprd_export(self.h, int(includeDetailed), counts.ctypes.data_as(i64p),
                                          prdLa.ctypes.data_as(i32p), hPrdLa.ctypes.data_as(i32p),
                                          off.ctypes.data_as(i64p), cIdx.ctypes.data_as(i32p), cFrac.ctypes.data_as(dp),
                                          rFrac.ctypes.data_as(dp), rI0.ctypes.data_as(i32p)))
        self.NprdLa = NprdLa
        ro, o = [], 0
        for a, t in lines:
            ro.append(o)
            o += p.atoms[a].trans[t].Nlambda * M * 2 * K
        return HybridPrd(NprdLa=NprdLa, NhPrd=NhPrd, prdLaOfLa=prdLa, hPrdLaOfLa=hPrdLa.reshape(1, L),
                         JRest=np.zeros((1, NprdLa, K)), JCoeffOff=off, JCoeffIdx=cIdx[:nnz], JCoeffFrac=cFrac[:nnz],
                         lineAtom=np.asarray([a for a, _ in lines], np.int32),
                         lineTrans=np.asarray([t for _, t in lines], np.int32),
                         rhoCoefOff=np.asarray(ro, np.int64), rhoFrac=rFrac[:tot], rhoI0=rI0[:tot])

    def jrest(self):
        import numpy as np
        out = np.zeros((self.NprdLa, self.problem.Nspace))
        _check(self.lib.lwref_get_jrest(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def full_stokes(self, updateJ=False, upOnly=True, J20=None):
        """J20: the 'J20' extra parameter of THIS column, float64 [Nspect, Nspace] (a view is fine as long
        as it is C-contiguous), or None"""
        dJ, idx = C.c_double(0.0), C.c_int64(0)
        if J20 is None:
            _check(self.lib.lwref_full_stokes(self.h, int(updateJ), int(upOnly), C.byref(dJ), C.byref(idx)))
        else:
            assert J20.flags.c_contiguous and J20.shape == (self.problem.Nspect, self.problem.Nspace)
            _check(self.lib.lwref_full_stokes_j20(self.h, int(updateJ), int(upOnly),
                                                  J20.ctypes.data_as(C.POINTER(C.c_double)), C.byref(dJ), C.byref(idx)))
        return dJ.value, idx.value

    def nr_post_update(self, upd):
        """upd: capi.LwB200NrUpdate"""
        _check(self.lib.lwref_nr_post_update(self.h, C.byref(upd)))

    def time_dep_update(self, activeIdx, nOld, dt):
        """nOld: [Nlevel, Nspace] of this context's column"""
        import numpy as np
        nOld = np.ascontiguousarray(nOld, dtype=np.float64)
        _check(self.lib.lwref_time_dep_update(self.h, activeIdx, nOld.ctypes.data_as(C.POINTER(C.c_double)), dt))

    def redistribute_prd(self, maxIter=3, tol=1e-2, includeDetailed=False, nlines=16):
        """-> dict(nIter, dRho [nIter, nlines'], dJPrdMax [nIter])"""
        import numpy as np
        n = C.c_int(0)
        dRho = np.zeros(maxIter * nlines)
        dRhoIdx = np.zeros(maxIter * nlines, dtype=np.int32)
        dJ = np.zeros(maxIter)
        dJIdx = np.zeros(maxIter, dtype=np.int64)
        dp = C.POINTER(C.c_double)
        _check(self.lib.lwref_redistribute_prd(self.h, maxIter, tol, int(includeDetailed), C.byref(n),
                                               dRho.ctypes.data_as(dp), dRhoIdx.ctypes.data_as(C.POINTER(C.c_int)),
                                               dJ.ctypes.data_as(dp), dJIdx.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(nIter=n.value, dRho=dRho, dRhoIdx=dRhoIdx, dJPrdMax=dJ[:n.value], dJPrdMaxIdx=dJIdx[:n.value])

    def compute_profiles(self):
        _check(self.lib.lwref_compute_profiles(self.h))

    def compute_polarised_profiles(self, col=0):
        """The reference's Transition::compute_polarised_profiles for every line of the problem with a Zeeman
        pattern, on column ``col`` (the one this context was made for): rewrites phi, wphi, polProfiles."""
        import numpy as np
        p = self.problem
        dp, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for ia, a in enumerate(p.atoms):
            for it, t in enumerate(a.trans):
                if t.zeeman is None:
                    continue
                al, sh, st = (np.ascontiguousarray(t.zeeman[0], dtype=np.int32), np.ascontiguousarray(t.zeeman[1]),
                              np.ascontiguousarray(t.zeeman[2]))
                arrs = [np.ascontiguousarray(x[col]) for x in (p.B, p.cosGamma, p.cos2chi, p.sin2chi)]
                _check(self.lib.lwref_compute_polarised_profiles(
                    self.h, ia, it, *[x.ctypes.data_as(dp) for x in arrs], len(al), al.ctypes.data_as(i32p),
                    sh.ctypes.data_as(dp), st.ctypes.data_as(dp)))

    def set_depth_fill(self, fill):
        _check(self.lib.lwref_set_depth_fill(self.h, int(fill)))

    def time_fs_iter(self, nWarm=3, nTimed=20, withStatEq=False):
        out = np.zeros(nTimed)
        _check(self.lib.lwref_time_fs_iter(self.h, nWarm, nTimed, int(withStatEq),
                                           out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def close(self):
        if self.h:
            self.lib.lwref_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_ray(solver, height, temperature, chi, S, muz, toObs, wavelength, lowerBc, upperBc,
              want_psi=True):
    lib = load()
    K = len(height)
    dp = C.POINTER(C.c_double)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (height, temperature, chi, S)]
    I = np.zeros(K)
    Psi = np.zeros(K)
    _check(lib.lwref_solve_ray(solver, K, *[a.ctypes.data_as(dp) for a in arrs], float(muz),
                               int(toObs), float(wavelength), lowerBc, upperBc,
                               I.ctypes.data_as(dp), Psi.ctypes.data_as(dp) if want_psi else dp()))
    return I, Psi


def load_formal_solver(path):
    """The reference's FormalSolverManager.load_fs_from_path on a plugin; returns (index to pass to
    solve_ray as `solver`, the solver's name)."""
    lib = load()
    idx = C.c_int(-1)
    name = C.create_string_buffer(128)
    _check(lib.lwref_load_formal_solver(os.fsencode(path), C.byref(idx), name, 128))
    return idx.value, name.value.decode()


def solve_lin_eq(A, b, improve=True):
    lib = load()
    A = np.array(A, dtype=np.float64, order='C')
    b = np.array(b, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    _check(lib.lwref_solve_lin_eq(A.shape[0], A.ctypes.data_as(dp), b.ctypes.data_as(dp), int(improve)))
    return b


def ng_run(Norder, Nperiod, Ndelay, sols):
    """Ng(Norder, Nperiod, Ndelay, sols[0]) then accelerate() + max_change() on sols[1:] (Ng.hpp).
    Returns (solutions after accelerate [nIter, len], accelerated [nIter], dMax, dMaxIdx)."""
    lib = load()
    sols = np.ascontiguousarray(sols, dtype=np.float64)
    nIter, n = sols.shape[0] - 1, sols.shape[1]
    out = np.zeros((nIter, n))
    acc = np.zeros(nIter, dtype=np.int32)
    dMax = np.zeros(nIter)
    dIdx = np.zeros(nIter, dtype=np.int64)
    dp = C.POINTER(C.c_double)
    rc = lib.lwref_ng_run(Norder, Nperiod, Ndelay, n, nIter, sols.ctypes.data_as(dp), out.ctypes.data_as(dp),
                       acc.ctypes.data_as(C.POINTER(C.c_int)), dMax.ctypes.data_as(dp),
                       dIdx.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc:
        raise RuntimeError('Singular Matrix')
    return out, acc, dMax, dIdx
