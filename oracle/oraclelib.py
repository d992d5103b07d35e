"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/liblw_oracle.so (the
plain-C restatement, oracle/lw_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(force=False):
    so = os.path.join(_HERE, 'liblw_oracle.so')
    src = os.path.join(_HERE, 'lw_oracle.c')
    deps = [src, os.path.join(_HERE, 'lw_oracle.h'), os.path.join(_HERE, '..', 'include', 'lwb200.h')]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(['make', '-C', _HERE, 'oracle'], stdout=subprocess.DEVNULL)
    return so


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(_HERE, 'liblw_oracle.so')
    if not os.path.exists(so):
        build()
    lib = C.CDLL(so)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    i64p = C.POINTER(C.c_int64)
    lib.lwo_solve_ray.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, C.c_double,
                                  C.c_int, C.c_int, C.c_double, dp, dp]
    lib.lwo_solve_ray.restype = None
    lib.lwo_fs_iter.argtypes = [vp, C.c_int, C.c_uint, C.c_int, C.c_int, dp, i64p, i64p]
    lib.lwo_formal_sol.argtypes = [vp, C.c_int, C.c_int]
    lib.lwo_stat_eq.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.lwo_solve_lin_eq.argtypes = [C.c_int, dp, dp, C.c_int]
    lib.lwo_ng_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int), dp,
                                  C.POINTER(C.c_int64)]
    lib.lwo_fs_iter_columns.argtypes = [vp, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int]
    lib.lwo_full_stokes.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, i64p]
    lib.lwo_full_stokes_j20.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, i64p]
    lib.lwo_nr_post_update.argtypes = [vp, C.c_int, vp]
    lib.lwo_time_dep_update.argtypes = [vp, C.c_int, C.c_int, dp, C.c_double]
    lib.lwo_redistribute_prd.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), dp,
                                         C.POINTER(C.c_int), dp, i64p]
    lib.lwo_configure_hprd.argtypes = [vp, C.c_int, vp]
    lib.lwo_free_hprd.argtypes = [vp]
    lib.lwo_free_hprd.restype = None
    _lib = lib
    return lib


def configure_hprd(problem, includeDetailed=False):
    """configure_hprd_coeffs (Prd.cpp:697-946) restated, for every column of the problem: returns a
    lightweaver_b200.problem.HybridPrd (assign it to problem.hprd to switch the hybrid scheme on)."""
    from lightweaver_b200 import capi
    from lightweaver_b200.problem import HybridPrd
    lib = load()
    keep, problem.hprd = problem.hprd, None
    try:
        cs = problem.c_struct()
    finally:
        problem.hprd = keep
    h = capi.LwB200HybridPrd()
    rc = lib.lwo_configure_hprd(C.byref(cs), int(includeDetailed), C.byref(h))
    if rc != 0:
        raise RuntimeError('configure_hprd: the problem has no vlosMu')
    try:
        return HybridPrd.from_c(h, problem) if h.Nlines > 0 else None
    finally:
        lib.lwo_free_hprd(C.byref(h))


class OracleContext:
    """Same call surface as reflib.RefContext, results in place in the Problem."""

    def __init__(self, problem, col=0):
        self.lib = load()
        self.problem = problem
        self.col = col
        self._cs = problem.c_struct()

    def fs_iter(self, lambdaIterate=False, storeDepth=False, laStart=0, laEnd=-1, serial_idx=False):
        dJ, idx, idxS = C.c_double(), C.c_int64(), C.c_int64()
        flags = (1 if lambdaIterate else 0) | (2 if storeDepth else 0)
        rc = self.lib.lwo_fs_iter(C.byref(self._cs), self.col, flags, laStart, laEnd,
                                  C.byref(dJ), C.byref(idx), C.byref(idxS))
        assert rc == 0
        return dJ.value, (idxS.value if serial_idx else idx.value)

    def formal_sol(self, upOnly=True):
        assert self.lib.lwo_formal_sol(C.byref(self._cs), self.col, int(upOnly)) == 0

    def stat_eq(self, atom=-1, kStart=-1, kEnd=-1):
        ns = C.c_int(0)
        rc = self.lib.lwo_stat_eq(C.byref(self._cs), self.col, atom, kStart, kEnd, C.byref(ns))
        if rc != 0:
            raise RuntimeError('Singular Matrix')

    def full_stokes(self, updateJ=False, upOnly=True, J20=None):
        """J20: the 'J20' extra parameter for the whole problem, float64 [Ncol, Nspect, Nspace], or None"""
        dJ, idx = C.c_double(0.0), C.c_int64(0)
        if J20 is None:
            assert self.lib.lwo_full_stokes(C.byref(self._cs), self.col, int(updateJ), int(upOnly), C.byref(dJ), C.byref(idx)) == 0
        else:
            assert J20.dtype == np.float64 and J20.flags.c_contiguous
            assert self.lib.lwo_full_stokes_j20(C.byref(self._cs), self.col, int(updateJ), int(upOnly),
                                                J20.ctypes.data_as(C.POINTER(C.c_double)), C.byref(dJ), C.byref(idx)) == 0
        return dJ.value, idx.value

    def nr_post_update(self, upd):
        """upd: capi.LwB200NrUpdate"""
        if self.lib.lwo_nr_post_update(C.byref(self._cs), self.col, C.byref(upd)) != 0:
            raise RuntimeError('Singular Matrix')

    def time_dep_update(self, atom, nOld, dt):
        nOld = np.ascontiguousarray(nOld, dtype=np.float64)
        rc = self.lib.lwo_time_dep_update(C.byref(self._cs), self.col, atom, nOld.ctypes.data_as(C.POINTER(C.c_double)), dt)
        if rc != 0:
            raise RuntimeError('Singular Matrix')

    def redistribute_prd(self, maxIter=3, tol=1e-2, includeDetailed=False, nlines=16):
        """-> dict(nIter, dRho, dRhoIdx, dJPrdMax, dJPrdMaxIdx) like reflib.RefContext.redistribute_prd"""
        n = C.c_int(0)
        dRho = np.zeros(maxIter * nlines)
        dRhoIdx = np.zeros(maxIter * nlines, dtype=np.int32)
        dJ = np.zeros(maxIter)
        dJIdx = np.zeros(maxIter, dtype=np.int64)
        dp = C.POINTER(C.c_double)
        rc = self.lib.lwo_redistribute_prd(C.byref(self._cs), self.col, maxIter, tol, int(includeDetailed),
                                           C.byref(n), dRho.ctypes.data_as(dp),
                                           dRhoIdx.ctypes.data_as(C.POINTER(C.c_int)), dJ.ctypes.data_as(dp),
                                           dJIdx.ctypes.data_as(C.POINTER(C.c_int64)))
        assert rc == 0
        return dict(nIter=n.value, dRho=dRho, dRhoIdx=dRhoIdx, dJPrdMax=dJ[:n.value], dJPrdMaxIdx=dJIdx[:n.value])

    def fs_iter_columns(self, col0, ncol, withStatEq=False, nthreads=0):
        rc = self.lib.lwo_fs_iter_columns(C.byref(self._cs), col0, ncol, 0, int(withStatEq), nthreads)
        assert rc == 0


def solve_ray(solver, height, temperature, chi, S, muz, toObs, wavelength, lowerBc, upperBc,
              bcValue=0.0, want_psi=True):
    lib = load()
    K = len(height)
    dp = C.POINTER(C.c_double)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (height, temperature, chi, S)]
    I, Psi = np.zeros(K), np.zeros(K)
    lib.lwo_solve_ray(solver, K, *[a.ctypes.data_as(dp) for a in arrs], float(muz), int(toObs),
                      float(wavelength), lowerBc, upperBc, float(bcValue), I.ctypes.data_as(dp),
                      Psi.ctypes.data_as(dp) if want_psi else dp())
    return I, Psi


def solve_lin_eq(A, b, improve=True):
    lib = load()
    A = np.array(A, dtype=np.float64, order='C')
    b = np.array(b, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = lib.lwo_solve_lin_eq(A.shape[0], A.ctypes.data_as(dp), b.ctypes.data_as(dp), int(improve))
    if rc:
        raise RuntimeError('Singular Matrix')
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
    rc = lib.lwo_ng_run(Norder, Nperiod, Ndelay, n, nIter, sols.ctypes.data_as(dp), out.ctypes.data_as(dp),
                       acc.ctypes.data_as(C.POINTER(C.c_int)), dMax.ctypes.data_as(dp),
                       dIdx.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc:
        raise RuntimeError('Singular Matrix')
    return out, acc, dMax, dIdx
