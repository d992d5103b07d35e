"""Host-side mirror of ``lw.Context`` for the hot path, over the C-ABI.

Same method names, argument meaning and error behaviour as the reference's
``LwContext`` (``Source/LwMiddleLayer.pyx``): ``formal_sol_gamma_matrices``
(:3152-3210), ``formal_sol`` (:3212-3241), ``stat_equil`` (:3461-3531),
``update_deps`` (:3244-3288).  The numpy buffers of the ``Problem`` are the
source of truth, exactly as the reference's Cython objects own theirs; every
call moves what the reference's caller may have mutated to the device and what
it reads afterwards back.  There is no CPU path:
without the CUDA library or a GPU every call raises.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi
from .problem import Problem


class ExplodingMatrixError(Exception):
    """Same name as the reference's exception for a singular stat-eq system
    (Source/LwMiddleLayer.pyx:3509-3514)."""


@dataclass
class IterationUpdate:
    """Subset of lightweaver.iteration_update.IterationUpdate this path fills
    (iteration_update.py:64-85)."""
    updatedJ: bool = False
    dJMax: float = 0.0
    dJMaxIdx: int = 0
    updatedPops: bool = False
    dPops: List[float] = field(default_factory=list)
    dPopsMaxIdx: List[int] = field(default_factory=list)
    crsw: float = 1.0
    updatedRho: bool = False
    NprdSubIter: int = 0
    dRho: List[float] = field(default_factory=list)
    dRhoMaxIdx: List[int] = field(default_factory=list)
    updatedJPrd: bool = False
    dJPrdMax: List[float] = field(default_factory=list)
    dJPrdMaxIdx: List[int] = field(default_factory=list)
    ngAccelerated: bool = False


class Context:
    """A Lightweaver-style context whose formal solution, Gamma accumulation and
    statistical-equilibrium solve run on one B200.

    ``problem`` may hold one column (a classic 1D Context) or a 1.5D stack.
    ``laRange`` restricts the wavelength sweep to ``[laStart, laEnd)`` (one
    lambda-shard of a multi-GPU run, see ``sharding.py``).
    """

    def __init__(self, problem: Problem, device: int = 0, stream=None, laRange=None,
                 upload: bool = True):
        self.lib = capi.load()
        self.problem = problem
        self.device = device
        self._cs = problem.c_struct()
        self._h = C.c_void_p()
        capi.check(self.lib.lwb200_create(C.byref(self._cs), device, C.byref(self._h)))
        self._stream, self._laRange = stream, laRange
        if stream is not None:
            self.set_stream(stream)
        if laRange is not None:
            self.set_lambda_range(*laRange)
        self.crsw = 1.0
        self._hprd_keep = []
        self._c_resident = False
        if upload:
            self.upload(capi.ALL_INPUTS)
            self.collisions_changed()

    # --------------------------------------------------------------- plumbing
    def close(self):
        if self._h:
            self.lib.lwb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """``stream``: a raw cudaStream_t (int) or a torch.cuda.Stream."""
        ptr = getattr(stream, 'cuda_stream', stream)
        capi.check(self.lib.lwb200_set_stream(self._h, C.c_void_p(int(ptr))))

    def set_lambda_range(self, laStart, laEnd):
        capi.check(self.lib.lwb200_set_lambda_range(self._h, int(laStart), int(laEnd)))

    def ng_configure(self, Norder=0, Nperiod=0, Ndelay=0):
        """Ng acceleration of the populations on the device (lw.Context's ``ngOptions``): the
        populations currently on the device become the first stored solution of every active atom."""
        capi.check(self.lib.lwb200_ng_configure(self._h, int(Norder), int(Nperiod), int(Ndelay)))
        self._ng = True

    def ng_accelerate_device(self):
        """Ng.accelerate + Ng.max_change on the device-resident populations of every active atom.
        Returns (accelerated, dMax per atom, dMaxIdx per atom) over ALL atoms of the problem
        (detailed-static ones report 0)."""
        na = len(self.problem.atoms)
        acc = C.c_int32(0)
        dMax = np.zeros(na)
        dIdx = np.zeros(na, dtype=np.int64)
        rc = self.lib.lwb200_ng_accelerate(self._h, C.byref(acc), capi.dptr(dMax),
                                           dIdx.ctypes.data_as(C.POINTER(C.c_int64)))
        if rc != 0:
            if b'Singular' in self.lib.lwb200_last_error():
                raise ExplodingMatrixError('Singular Matrix')
            capi.check(rc)
        return bool(acc.value), dMax, dIdx

    def set_active_columns(self, active=None):
        """Retire converged columns of a stack: ``active`` is a boolean array [Ncol] (None: all
        columns again).  Retired columns are skipped by the Gamma iteration, the formal solution
        and the population update; their arrays keep their values."""
        if active is None:
            capi.check(self.lib.lwb200_set_active_columns(self._h, None))
            return
        a = np.ascontiguousarray(np.asarray(active).astype(np.uint8))
        if a.shape != (self.problem.Ncol,):
            raise ValueError('active must have one entry per column')
        capi.check(self.lib.lwb200_set_active_columns(self._h, a.ctypes.data_as(C.POINTER(C.c_uint8))))

    def upload(self, mask):
        capi.check(self.lib.lwb200_upload(self._h, mask))

    def configure_hprd_coeffs(self, includeDetailed=False):
        """lw.Context.configure_hprd_coeffs after the velocity field changed: rebuild the hybrid-PRD tables
        and hand them to the device (the context must have been created with tables: call
        problem.configure_hprd() before constructing it)."""
        self.problem.configure_hprd(includeDetailed)
        self._hprd_cs = self.problem.hprd.c_struct(self._hprd_keep)
        if self.lib.lwb200_set_hybrid_prd(self._h, self._hprd_cs) != 0:
            msg = (self.lib.lwb200_last_error() or b'').decode()
            if 'create a new context' not in msg:
                raise capi.LwB200Error(msg)
            # the new velocity field scatters from wavelengths the plan did not route through the general
            # kernel: plan again (what the plugin shim does whenever configure_hprd_coeffs has run)
            self._rebuild()

    def _rebuild(self):
        """A new device context for the problem as it is now (same stream, wavelength range and options);
        every input travels again."""
        self.lib.lwb200_destroy(self._h)
        self._h = C.c_void_p()
        self._cs = self.problem.c_struct()
        capi.check(self.lib.lwb200_create(C.byref(self._cs), self.device, C.byref(self._h)))
        if self._stream is not None:
            self.set_stream(self._stream)
        if self._laRange is not None:
            self.set_lambda_range(*self._laRange)
        self._ng = False
        self._stokes_sent = False
        self._zplane = False
        self.upload(capi.ALL_INPUTS)
        self.collisions_changed()
        if any(t.rhoPrd is not None for a in self.problem.atoms for t in a.trans):
            self.upload(capi.PRD)

    def collisions_changed(self):
        """The collisional rates C of the active atoms are context state on the device (they change
        with the atmosphere, not with the iteration): send them (again).  formal_sol_gamma_matrices
        then makes the prefill Gamma = crsw*C on the device."""
        self.upload(capi.COLLISIONS)
        self._c_resident = True

    def download(self, mask, sync=True):
        capi.check(self.lib.lwb200_download(self._h, mask))
        if sync:
            self.sync()

    def sync(self):
        capi.check(self.lib.lwb200_sync(self._h))

    def device_buffer(self, which):
        """(device pointer, nbytes) of one of the capi.BUF_* buffers."""
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        capi.check(self.lib.lwb200_device_buffer(self._h, which, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def work_stats(self):
        pts, byts, launches = C.c_double(), C.c_double(), C.c_int64()
        capi.check(self.lib.lwb200_work_stats(self._h, C.byref(pts), C.byref(byts), C.byref(launches)))
        return pts.value, byts.value, launches.value

    def kernel_time_ms(self):
        """Device time of the most recent formal-solution kernel (CUDA events)."""
        ms = C.c_double()
        capi.check(self.lib.lwb200_kernel_time(self._h, C.byref(ms)))
        return ms.value

    # ------------------------------------------------- device-resident calls
    def fs_iter_device(self, lambdaIterate=False, storeDepth=False, deferFinalise=False,
                       want_dJ=True, generalKernel=False, fetchEarly=False, asyncDJ=False):
        """One Gamma iteration on device-resident data (no host copies).  asyncDJ: dJ is reduced
        on the device and travels with the stream; read it with last_dj() after a sync()."""
        flags = ((capi.LAMBDA_ITERATE if lambdaIterate else 0) | (capi.STORE_DEPTH if storeDepth else 0)
                 | (capi.DEFER_FINALISE if deferFinalise else 0)
                 | (capi.GENERAL_KERNEL if generalKernel else 0)
                 | (capi.FETCH_EARLY if fetchEarly else 0))
        if asyncDJ:
            capi.check(self.lib.lwb200_fs_iter(self._h, flags | capi.DJ_ASYNC, None, None))
            return None
        if want_dJ:
            dJ, idx = C.c_double(), C.c_int64()
            capi.check(self.lib.lwb200_fs_iter(self._h, flags, C.byref(dJ), C.byref(idx)))
            return dJ.value, idx.value
        capi.check(self.lib.lwb200_fs_iter(self._h, flags, None, None))
        return None

    def finalise(self):
        capi.check(self.lib.lwb200_finalise(self._h))

    def dj_max(self):
        dJ, idx = C.c_double(), C.c_int64()
        capi.check(self.lib.lwb200_dj_max(self._h, C.byref(dJ), C.byref(idx)))
        return dJ.value, idx.value

    def last_dj(self):
        """(dJMax, flat index) of the last asyncDJ iteration; valid after sync()."""
        dJ, idx = C.c_double(), C.c_int64()
        capi.check(self.lib.lwb200_last_dj(self._h, C.byref(dJ), C.byref(idx)))
        return dJ.value, idx.value

    def check_singular(self):
        """Raise ExplodingMatrixError if the last wait=False population update met a singular
        system; valid after sync()."""
        ns = C.c_int32(0)
        if self.lib.lwb200_last_singular(self._h, C.byref(ns)) != 0:
            if ns.value > 0:
                raise ExplodingMatrixError('Singular Matrix')
            capi.check(1)

    def stat_eq_device(self, atom=-1, kStart=-1, kEnd=-1, wait=True):
        """Statistical equilibrium on device-resident data.  wait=False: no host synchronisation;
        call check_singular() after the next sync()."""
        if not wait:
            capi.check(self.lib.lwb200_stat_eq_async(self._h, atom, kStart, kEnd))
            return
        ns = C.c_int32(0)
        rc = self.lib.lwb200_stat_eq(self._h, atom, kStart, kEnd, C.byref(ns))
        if rc != 0:
            if ns.value > 0:
                raise ExplodingMatrixError('Singular Matrix')
            capi.check(rc)

    def prd_redistribute_device(self, maxIter=3, tol=1e-2, includeDetailed=False):
        """PRD sub-iterations on device-resident data (inputs uploaded with capi.PRD)."""
        n = C.c_int32(0)
        capi.check(self.lib.lwb200_redistribute_prd(self._h, maxIter, tol, int(includeDetailed), C.byref(n),
                                                    None, None, None, None))
        return n.value

    def compute_profiles_device(self):
        capi.check(self.lib.lwb200_compute_profiles(self._h))

    def compute_polarised_profiles_device(self):
        """Transition::compute_polarised_profiles (what lw.Context.setup_stokes / update_deps(B=True) run per
        polarised line, FormalStokes.cpp:9-117) on the device: phi, wphi and the six polarised profiles of every
        line with a Zeeman pattern, from the problem's B, cosGamma, cos2chi, sin2chi."""
        p = self.problem
        if p.B is None or p.cosGamma is None or p.cos2chi is None or p.sin2chi is None:
            raise capi.LwB200Error('compute_polarised_profiles: the problem has no magnetic field / projections')
        lines = [(ia, it, t) for ia, a in enumerate(p.atoms) for it, t in enumerate(a.trans) if t.zeeman is not None]
        if not lines:
            return
        arr = (capi.LwB200Zeeman * len(lines))()
        keep = []
        for z, (ia, it, t) in zip(arr, lines):
            al, sh, st = (np.ascontiguousarray(t.zeeman[0], dtype=np.int32),
                          np.ascontiguousarray(t.zeeman[1], dtype=np.float64),
                          np.ascontiguousarray(t.zeeman[2], dtype=np.float64))
            keep += [al, sh, st]
            z.atom, z.trans, z.Ncomponent = ia, it, len(al)
            z.alpha, z.shift, z.strength = capi.iptr(al), capi.dptr(sh), capi.dptr(st)
        capi.check(self.lib.lwb200_compute_polarised_profiles(self._h, arr, len(lines), capi.dptr(p.B),
                                                              capi.dptr(p.cosGamma), capi.dptr(p.cos2chi),
                                                              capi.dptr(p.sin2chi)))
        self._stokes_sent = True

    # ------------------------------------------- the reference's call surface
    def formal_sol_gamma_matrices(self, fixCollisionalRates=True, lambdaIterate=False,
                                  extraParams=None, crsw=None):
        """lw.Context.formal_sol_gamma_matrices: Gamma = crsw*C (the prologue,
        LwMiddleLayer.pyx:3198-3203), then the device Gamma iteration; I, J,
        Gamma and the rates are back in the numpy buffers on return.  The
        collisional rates C are an input of this path (computed by the
        reference's Python layer); fixCollisionalRates=False re-sends them."""
        storeDepth = bool(extraParams and extraParams.get('storeDepthData', False))
        general = bool(extraParams and extraParams.get('generalKernel', False))
        zplane = self._bind_zplane(extraParams)
        if crsw is not None:
            self.crsw = crsw
        # What the reference's caller may have changed since the last call travels on EVERY call, as in the
        # plugin shim (lwb200_plugin.cpp sync_inputs): the populations and nStar / nTotal / vBroad.  The
        # prologue Gamma = crsw*C (LwMiddleLayer.pyx:3198-3203) runs on the device from the resident C:
        # fixCollisionalRates=False is the reference recomputing C first (compute_collisions on the host,
        # outside this path), so C is sent again; otherwise neither the product nor its upload is paid.
        if not fixCollisionalRates or not self._c_resident:
            self.collisions_changed()
        capi.check(self.lib.lwb200_set_collision_prefill(self._h, 1, float(self.crsw)))
        self.upload(capi.POPS | capi.NSTAR)
        # one host synchronisation for the whole call: dJ comes home with the stream (DJ_ASYNC), J and I
        # start travelling as soon as the rays are done (FETCH_EARLY)
        flags = ((capi.LAMBDA_ITERATE if lambdaIterate else 0) | (capi.STORE_DEPTH if storeDepth else 0)
                 | (capi.GENERAL_KERNEL if general else 0) | capi.FETCH_EARLY | capi.DJ_ASYNC)
        capi.check(self.lib.lwb200_fs_iter(self._h, flags, None, None))
        self.download(capi.ITER_OUTPUTS | (capi.DEPTH if storeDepth else 0) | (capi.ZPLANE if zplane else 0)
                      | (capi.PRD if self.problem.hprd is not None else 0))   # (hybrid PRD: JRest)
        dJ, idx = C.c_double(), C.c_int64()
        capi.check(self.lib.lwb200_last_dj(self._h, C.byref(dJ), C.byref(idx)))
        return IterationUpdate(updatedJ=True, dJMax=dJ.value, dJMaxIdx=idx.value % self.problem.Nspect,
                               crsw=self.crsw)

    def _bind_zplane(self, extraParams):
        """The ZPlaneDecomposition extra parameters (SimdFullIterationTemplates.hpp:254-281):
        extraParams = {'ZPlaneDecomposition': True, 'ZPlaneUp': array [Ncol, Nspect, Nrays] and / or
        'ZPlaneDown': ...}; the arrays receive I(1) of the up-going / I(Nz - 2) of the down-going rays."""
        up = down = None
        if extraParams and extraParams.get('ZPlaneDecomposition', False):
            up, down = extraParams.get('ZPlaneUp'), extraParams.get('ZPlaneDown')
            shape = (self.problem.Ncol, self.problem.Nspect, self.problem.Nrays)
            for a in (up, down):
                if a is not None and (a.shape != shape or a.dtype != np.float64 or not a.flags.c_contiguous):
                    raise ValueError(f'ZPlaneUp / ZPlaneDown must be C-contiguous float64 arrays of shape {shape}')
        on = up is not None or down is not None
        if on or getattr(self, '_zplane', False):
            capi.check(self.lib.lwb200_set_zplane(self._h, capi.dptr(up), capi.dptr(down)))
        self._zplane = on
        self._zplane_keep = (up, down)
        return on

    def formal_sol(self, upOnly=True, extraParams=None):
        """lw.Context.formal_sol: intensity only (LwMiddleLayer.pyx:3212-3241)."""
        zplane = self._bind_zplane(extraParams)
        self.upload(capi.POPS | capi.NSTAR | capi.JBAR)
        capi.check(self.lib.lwb200_formal_sol(self._h, int(upOnly)))
        self.download(capi.INTENS | (capi.ZPLANE if zplane else 0))
        return IterationUpdate()

    def single_stokes_fs(self, recompute=False, updateJ=False, upOnly=True, extraParams=None):
        """lw.Context.single_stokes_fs (LwMiddleLayer.pyx:3605-3645): full Stokes formal solution of
        every wavelength from the current populations; I and Quv (and J when updateJ) are back in
        the numpy buffers on return.  The polarised profiles are an input of this path (the
        reference's setup_stokes / compute_polarised_profiles made them); recompute=True sends
        them again."""
        if self.problem.Quv is None:
            raise capi.LwB200Error('single_stokes_fs: the problem has no polarised line')
        # extraParams = {'J20': float64 array [Ncol, Nspect, Nspace]}: the radiation-field anisotropy
        # (FormalStokes.cpp:676-681), read and -- with updateJ -- rewritten in place
        J20 = extraParams.get('J20') if extraParams else None
        if J20 is not None:
            shape = (self.problem.Ncol, self.problem.Nspect, self.problem.Nspace)
            if J20.shape != shape or J20.dtype != np.float64 or not J20.flags.c_contiguous:
                raise ValueError(f'J20 must be a C-contiguous float64 array of shape {shape}')
        if J20 is not None or getattr(self, '_j20', None) is not None:
            capi.check(self.lib.lwb200_set_j20(self._h, capi.dptr(J20)))
        self._j20 = J20
        first = not getattr(self, '_stokes_sent', False)
        self.upload(capi.POPS | capi.JBAR | (capi.STOKES if (first or recompute) else 0))
        self._stokes_sent = True
        dJ, idx = C.c_double(0.0), C.c_int64(0)
        capi.check(self.lib.lwb200_formal_sol_full_stokes(self._h, int(updateJ), int(upOnly), C.byref(dJ),
                                                          C.byref(idx)))
        self.download(capi.INTENS | capi.STOKES | (capi.JBAR if updateJ else 0))
        return IterationUpdate(updatedJ=bool(updateJ), dJMax=dJ.value, dJMaxIdx=idx.value % self.problem.Nspect)

    def prd_redistribute(self, maxIter=3, tol=1e-2, extraParams=None):
        """lw.Context.prd_redistribute (LwMiddleLayer.pyx:3647-3684): update the emission-profile
        ratio rho of every angle-averaged PRD line from the current J, populations and rates, then
        J, I and the PRD lines' rates from the new rho; up to maxIter sub-iterations."""
        include = bool(extraParams and extraParams.get('include_detailed_atoms', False))
        lines = [t for a in self.problem.atoms for t in a.trans
                 if t.rhoPrd is not None and (include or not a.detailedStatic)]
        upd = IterationUpdate()
        if not lines:
            return upd
        self.upload(capi.PRD | capi.POPS | capi.RATES | capi.JBAR)
        n = C.c_int32(0)
        dRho = np.zeros(maxIter * len(lines))
        dRhoIdx = np.zeros(maxIter * len(lines), dtype=np.int32)
        dJ = np.zeros(maxIter)
        dJIdx = np.zeros(maxIter, dtype=np.int64)
        capi.check(self.lib.lwb200_redistribute_prd(
            self._h, maxIter, tol, int(include), C.byref(n), capi.dptr(dRho), capi.iptr(dRhoIdx),
            capi.dptr(dJ), dJIdx.ctypes.data_as(C.POINTER(C.c_int64))))
        self.download(capi.PRD | capi.JBAR | capi.INTENS | capi.RATES)
        k = n.value
        upd.updatedRho = upd.updatedJPrd = True
        upd.NprdSubIter = k
        upd.dRho = list(dRho[:k * len(lines)])
        upd.dRhoMaxIdx = list(dRhoIdx[:k * len(lines)])
        upd.dJPrdMax = list(dJ[:k])
        upd.dJPrdMaxIdx = list(dJIdx[:k])
        return upd

    def stat_equil(self, extraParams=None):
        """lw.Context.stat_equil: per-depth statistical equilibrium from the
        current Gamma (LwMiddleLayer.pyx:3461-3531); raises ExplodingMatrixError
        on a singular system.  Returns the relative population change per atom
        (what rel_diff_ng_accelerate reports without Ng acceleration)."""
        # the update, the Ng step and the change tracking all run on the device-resident populations
        # (rel_diff_ng_accelerate, LwMiddleLayer.pyx:3318-3346); one host synchronisation.  Without
        # ng_configure this is the reference's default Ng(0, 0, 0): change tracking only.
        self.upload(capi.POPS | capi.GAMMA_FINAL)
        if not getattr(self, '_ng', False):
            self.ng_configure(0, 0, 0)      # first stored solution: the populations before this update
        capi.check(self.lib.lwb200_stat_eq_async(self._h, -1, -1, -1))
        acc = C.c_int32(0)
        capi.check(self.lib.lwb200_ng_accelerate(self._h, C.byref(acc), None, None))
        self.download(capi.POPS)   # (synchronises)
        ns = C.c_int32(0)
        if self.lib.lwb200_last_singular(self._h, C.byref(ns)) != 0:
            if ns.value > 0:
                raise ExplodingMatrixError('Singular Matrix')
            capi.check(1)
        na = len(self.problem.atoms)
        dMax = np.zeros(na)
        dIdx = np.zeros(na, dtype=np.int64)
        capi.check(self.lib.lwb200_last_ng(self._h, capi.dptr(dMax), dIdx.ctypes.data_as(C.POINTER(C.c_int64))))
        upd = IterationUpdate(updatedPops=True)
        upd.ngAccelerated = bool(acc.value)
        for q, a in enumerate(self.problem.atoms):
            if not a.detailedStatic:
                upd.dPops.append(float(dMax[q]))
                upd.dPopsMaxIdx.append(int(dIdx[q]))
        return upd

    def time_dep_update(self, dt, prevTimePops=None, extraParams=None):
        """lw.Context.time_dep_update (LwMiddleLayer.pyx:3348-3432) without Ng: one backward-Euler
        step of every active atom's populations from the current Gamma.  Returns (update,
        prevTimePops) like the reference; pass the returned prevTimePops back in on the
        following sub-iterations of the same time step."""
        atoms = self.problem.active_atoms()
        if prevTimePops is None:
            prevTimePops = [np.copy(a.n) for a in atoms]
        before = [a.n.copy() for a in atoms]
        self.upload(capi.GAMMA_FINAL)
        for a, nOld in zip(atoms, prevTimePops):
            idx = self.problem.atoms.index(a)
            ns = C.c_int32(0)
            rc = self.lib.lwb200_time_dep_update(self._h, idx, capi.dptr(np.ascontiguousarray(nOld)), float(dt),
                                                 -1, -1, C.byref(ns))
            if rc != 0:
                if ns.value > 0:
                    raise ExplodingMatrixError('Singular Matrix')
                capi.check(rc)
        self.download(capi.POPS)
        upd = IterationUpdate(updatedPops=True)
        for p, a in zip(before, atoms):
            with np.errstate(divide='ignore', invalid='ignore'):
                d = np.abs(1.0 - p / a.n)
            d = np.where(np.isfinite(d), d, 0.0)
            upd.dPops.append(float(d.max()))
            upd.dPopsMaxIdx.append(int(d.argmax()))
        return upd, prevTimePops

    def nr_post_update(self, atoms, dC, backgroundNe, timeDependentData=None, extraParams=None):
        """lw.Context._nr_post_update_impl (LwMiddleLayer.pyx:3533-3564): one Newton-Raphson step of the
        populations of ``atoms`` (AtomData objects or indices of active atoms) and of the electron
        density (charge conservation).  ``dC``: list of finite-difference dC/dne arrays (one per atom)
        or an empty list; ``timeDependentData``: {'dt': ..., 'nPrev': [...]} or None.  Populations and
        ``problem.ne`` are updated in place."""
        idx = [a if isinstance(a, int) else self.problem.atoms.index(a) for a in atoms]
        nPrev = dt = None
        if timeDependentData is not None:
            dt, nPrev = timeDependentData['dt'], list(timeDependentData['nPrev'])
        upd, keep = capi.make_nr_update(idx, np.ascontiguousarray(backgroundNe, dtype=np.float64),
                                        dC=list(dC) if dC is not None and len(dC) > 0 else None, nPrev=nPrev,
                                        dt=dt or 0.0, crswVal=self.crsw)
        self.upload(capi.POPS | capi.GAMMA_FINAL)
        ns = C.c_int32(0)
        rc = self.lib.lwb200_nr_post_update(self._h, C.byref(upd), -1, -1, C.byref(ns))
        if rc != 0:
            if ns.value > 0:
                raise ExplodingMatrixError('Singular Matrix')
            capi.check(rc)
        self.download(capi.POPS)

    def update_deps(self, temperature=True, ne=True, vturb=True, vlos=True, B=True,
                    background=True, hprd=True, quiet=True, profiles_on_device=False):
        """lw.Context.update_deps (LwMiddleLayer.pyx:3244-3288) as seen from the
        back end: whatever the host recomputed (projections, profiles, LTE
        populations, background) is re-mirrored.  With profiles_on_device the
        Voigt profiles are regenerated on the GPU from aDamp/vBroad/vlosMu
        instead of being uploaded."""
        mask = capi.ATMOS | capi.NSTAR | capi.POPS | capi.COLLISIONS
        if background:
            mask |= capi.BACKGR
        if not profiles_on_device:
            mask |= capi.PROFILE
        self.upload(mask)
        self._c_resident = True
        if profiles_on_device:
            self.upload(capi.ADAMP)
            self.compute_profiles_device()
