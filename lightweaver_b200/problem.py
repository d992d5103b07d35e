"""Host-side data model of the hot path: the numpy buffers a Lightweaver
``Context`` owns for the formal solution / Gamma iteration / stat-eq step, laid
out exactly as the reference's Cython classes allocate them
(``Source/LwMiddleLayer.pyx``: LwAtmosphere :620-715, LwSpectrum :2724-2732,
LwBackground :1571-1597, LwTransition :1772-1825, LwAtom :2346-2424), with one
extension: every per-column array carries a leading ``[Ncol]`` axis so that a
1.5D stack of columns is one object.

``Problem.c_struct()`` marshals the buffers into the ``LwB200Problem`` of
``include/lwb200.h`` (non-owning pointers, as the reference's ``f64_view``
helpers do, ``Source/CmoArrayHelper.pyx:4-15``).
"""
import copy
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi


@dataclass
class TransitionData:
    """One radiative transition (reference: ``struct Transition``,
    Source/LwTransition.hpp:22-91)."""
    type: int
    i: int
    j: int
    Nblue: int
    Nred: int
    lambda0: float
    wavelength: np.ndarray                 # [Nlambda]
    Aji: float = 0.0
    Bji: float = 0.0
    Bij: float = 0.0
    dopplerWidth: float = 1.0
    alpha: Optional[np.ndarray] = None     # [Nlambda] continua
    phi: Optional[np.ndarray] = None       # [Ncol, Nlambda, Nrays, 2, Nspace] lines
    wphi: Optional[np.ndarray] = None      # [Ncol, Nspace]
    rhoPrd: Optional[np.ndarray] = None    # [Ncol, Nlambda, Nspace]
    aDamp: Optional[np.ndarray] = None     # [Ncol, Nspace]
    Qelast: Optional[np.ndarray] = None    # [Ncol, Nspace] PRD lines
    polProfiles: Optional[np.ndarray] = None  # [6, Ncol, Nlambda, Nrays, 2, Nspace] phiQ,U,V psiQ,U,V
    polarised: bool = False                # polarised line whose profiles may live on the device only
    zeeman: Optional[tuple] = None         # (alpha int32 [N], shift [N], strength [N]): ZeemanComponents
    Rij: Optional[np.ndarray] = None       # [Ncol, Nspace]
    Rji: Optional[np.ndarray] = None
    name: str = ''

    @property
    def Nlambda(self):
        return self.Nred - self.Nblue

    def wlambda(self):
        """Wavelength quadrature weights (Transition::wlambda,
        Source/LwTransition.hpp:71-81)."""
        w = self.wavelength
        out = np.empty_like(w)
        out[0] = 0.5 * (w[1] - w[0])
        out[-1] = 0.5 * (w[-1] - w[-2])
        out[1:-1] = 0.5 * (w[2:] - w[:-2])
        return out * self.dopplerWidth


@dataclass
class AtomData:
    """One atom (reference: ``struct Atom``, Source/LwAtom.hpp:42-80)."""
    name: str
    Nlevel: int
    trans: List[TransitionData]
    n: np.ndarray                          # [Ncol, Nlevel, Nspace]
    nStar: np.ndarray
    nTotal: np.ndarray                     # [Ncol, Nspace]
    vBroad: Optional[np.ndarray] = None    # [Ncol, Nspace]
    Gamma: Optional[np.ndarray] = None     # [Ncol, Nlevel, Nlevel, Nspace]
    C: Optional[np.ndarray] = None         # [Ncol, Nlevel, Nlevel, Nspace] collisional rates
    stages: Optional[np.ndarray] = None    # [Nlevel] ionisation stage of every level (float64)
    detailedStatic: bool = False

    @property
    def Ntrans(self):
        return len(self.trans)


@dataclass
class HybridPrd:
    """Hybrid-PRD tables (reference: what ``configure_hprd_coeffs``, Source/Prd.cpp:697-946, leaves in
    ``Spectrum`` and in every PRD line's ``hPrdCoeffs``), flattened as ``LwB200HybridPrd`` lays them out."""
    NprdLa: int
    NhPrd: int
    prdLaOfLa: np.ndarray     # [Nspect] int32
    hPrdLaOfLa: np.ndarray    # [Ncol, Nspect] int32
    JRest: np.ndarray         # [Ncol, NprdLa, Nspace]
    JCoeffOff: np.ndarray     # [Ncol*NhPrd*Nrays*2*Nspace + 1] int64
    JCoeffIdx: np.ndarray     # [nnz] int32
    JCoeffFrac: np.ndarray    # [nnz]
    lineAtom: np.ndarray      # [Nlines] int32
    lineTrans: np.ndarray     # [Nlines] int32
    rhoCoefOff: np.ndarray    # [Nlines] int64
    rhoFrac: np.ndarray       # per line [Ncol, Nlambda, Nrays, 2, Nspace], back to back
    rhoI0: np.ndarray         # int32, same layout

    @staticmethod
    def from_c(h, problem):
        """Copy a C-side LwB200HybridPrd (malloc'ed by a configure routine) into numpy arrays."""
        Ncol, L, K, M = problem.Ncol, problem.Nspect, problem.Nspace, problem.Nrays

        def arr(ptr, n, dtype):
            if n == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
        nRows = Ncol * h.NhPrd * M * 2 * K
        off = arr(h.JCoeffOff, nRows + 1, np.int64)
        nnz = int(off[-1])
        nl = h.Nlines
        lineAtom, lineTrans = arr(h.lineAtom, nl, np.int32), arr(h.lineTrans, nl, np.int32)
        tot = sum(Ncol * problem.atoms[a].trans[t].Nlambda * M * 2 * K for a, t in zip(lineAtom, lineTrans))
        return HybridPrd(NprdLa=h.NprdLa, NhPrd=h.NhPrd, prdLaOfLa=arr(h.prdLaOfLa, L, np.int32),
                         hPrdLaOfLa=arr(h.hPrdLaOfLa, Ncol * L, np.int32).reshape(Ncol, L),
                         JRest=arr(h.JRest, Ncol * h.NprdLa * K, np.float64).reshape(Ncol, h.NprdLa, K),
                         JCoeffOff=off, JCoeffIdx=arr(h.JCoeffIdx, nnz, np.int32),
                         JCoeffFrac=arr(h.JCoeffFrac, nnz, np.float64), lineAtom=lineAtom, lineTrans=lineTrans,
                         rhoCoefOff=arr(h.rhoCoefOff, nl, np.int64), rhoFrac=arr(h.rhoFrac, tot, np.float64),
                         rhoI0=arr(h.rhoI0, tot, np.int32))

    def c_struct(self, keep):
        h = capi.LwB200HybridPrd()
        h.NprdLa, h.NhPrd, h.Nlines = self.NprdLa, self.NhPrd, len(self.lineAtom)
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        h.prdLaOfLa, h.hPrdLaOfLa = capi.iptr(self.prdLaOfLa), capi.iptr(self.hPrdLaOfLa)
        h.JRest = capi.dptr(self.JRest)
        h.JCoeffOff, h.JCoeffIdx, h.JCoeffFrac = lp(self.JCoeffOff), capi.iptr(self.JCoeffIdx), capi.dptr(self.JCoeffFrac)
        h.lineAtom, h.lineTrans = capi.iptr(self.lineAtom), capi.iptr(self.lineTrans)
        h.rhoCoefOff, h.rhoFrac, h.rhoI0 = lp(self.rhoCoefOff), capi.dptr(self.rhoFrac), capi.iptr(self.rhoI0)
        keep.append(h)
        return C.pointer(h)

    def column(self, c, problem):
        """The tables of column ``c`` alone (JRest is a view)."""
        K, M = problem.Nspace, problem.Nrays
        per = self.NhPrd * M * 2 * K
        off = self.JCoeffOff[c * per:(c + 1) * per + 1]
        e0, e1 = int(off[0]), int(off[-1])
        fr, i0, ro, o = [], [], [], 0
        for q, (a, t) in enumerate(zip(self.lineAtom, self.lineTrans)):
            n = problem.atoms[a].trans[t].Nlambda * M * 2 * K
            b = int(self.rhoCoefOff[q]) + c * n
            fr.append(self.rhoFrac[b:b + n])
            i0.append(self.rhoI0[b:b + n])
            ro.append(o)
            o += n
        return HybridPrd(NprdLa=self.NprdLa, NhPrd=self.NhPrd, prdLaOfLa=self.prdLaOfLa,
                         hPrdLaOfLa=np.ascontiguousarray(self.hPrdLaOfLa[c:c + 1]), JRest=self.JRest[c:c + 1],
                         JCoeffOff=np.ascontiguousarray(off - e0), JCoeffIdx=np.ascontiguousarray(self.JCoeffIdx[e0:e1]),
                         JCoeffFrac=np.ascontiguousarray(self.JCoeffFrac[e0:e1]), lineAtom=self.lineAtom,
                         lineTrans=self.lineTrans, rhoCoefOff=np.asarray(ro, dtype=np.int64),
                         rhoFrac=np.concatenate(fr) if fr else np.zeros(0), rhoI0=np.concatenate(i0) if i0 else np.zeros(0, np.int32))


@dataclass
class Problem:
    Nspace: int
    Nrays: int
    height: np.ndarray        # [Ncol, Nspace]
    temperature: np.ndarray
    muz: np.ndarray           # [Nrays]
    wmu: np.ndarray
    wavelength: np.ndarray    # [Nspect]
    chiBg: np.ndarray         # [Ncol, Nspect, Nspace]
    etaBg: np.ndarray
    scaBg: np.ndarray
    atoms: List[AtomData]
    vlosMu: Optional[np.ndarray] = None  # [Ncol, Nrays, Nspace]
    J: Optional[np.ndarray] = None       # [Ncol, Nspect, Nspace]
    I: Optional[np.ndarray] = None       # [Ncol, Nspect, Nrays]
    formalSolver: int = capi.FS_BEZIER3
    lowerBc: int = capi.BC_THERMALISED   # Atmosphere.make_1d defaults, atmosphere.py:923-928
    upperBc: int = capi.BC_ZERO
    lowerBcData: Optional[np.ndarray] = None  # [Ncol, Nspect, Nmu]
    upperBcData: Optional[np.ndarray] = None
    lowerBcIdx: Optional[np.ndarray] = None   # [Nrays, 2] int32
    upperBcIdx: Optional[np.ndarray] = None
    depthChi: Optional[np.ndarray] = None     # [Ncol, Nspect, Nrays, 2, Nspace]
    depthEta: Optional[np.ndarray] = None
    depthI: Optional[np.ndarray] = None
    Quv: Optional[np.ndarray] = None          # [Ncol, 3, Nspect, Nrays]
    ne: Optional[np.ndarray] = None      # [Ncol, Nspace] (inputs of the synthetic generator; not on the path)
    vturb: Optional[np.ndarray] = None
    nHTot: Optional[np.ndarray] = None
    hprd: Optional[HybridPrd] = None     # hybrid-PRD tables (None: every PRD line is angle-averaged)
    B: Optional[np.ndarray] = None       # [Ncol, Nspace] magnetic field strength (atmos.B), Tesla
    cosGamma: Optional[np.ndarray] = None  # [Ncol, Nrays, Nspace] projections of Atmosphere::update_projections
    cos2chi: Optional[np.ndarray] = None
    sin2chi: Optional[np.ndarray] = None
    meta: dict = field(default_factory=dict)
    _keepalive: list = field(default_factory=list, repr=False)

    @property
    def Ncol(self):
        return self.height.shape[0]

    @property
    def Nspect(self):
        return self.wavelength.shape[0]

    def __post_init__(self):
        Ncol, L, K, M = self.Ncol, self.Nspect, self.Nspace, self.Nrays
        if self.J is None:
            self.J = np.zeros((Ncol, L, K))
        if self.I is None:
            self.I = np.zeros((Ncol, L, M))
        for a in self.atoms:
            for t in a.trans:
                if t.Rij is None:
                    t.Rij = np.zeros((Ncol, K))
                if t.Rji is None:
                    t.Rji = np.zeros((Ncol, K))
            if not a.detailedStatic and a.Gamma is None:
                a.Gamma = np.zeros((Ncol, a.Nlevel, a.Nlevel, K))

    # ------------------------------------------------------------------ sizes
    def points_per_iter(self, laStart=0, laEnd=None):
        """Ray-depth points of one Gamma iteration: Nspace*Nspect*Nrays*2 per
        column."""
        laEnd = self.Nspect if laEnd is None else laEnd
        return float(self.Ncol) * self.Nspace * (laEnd - laStart) * self.Nrays * 2

    def alloc_depth_data(self):
        shape = (self.Ncol, self.Nspect, self.Nrays, 2, self.Nspace)
        self.depthChi = np.zeros(shape)
        self.depthEta = np.zeros(shape)
        self.depthI = np.zeros(shape)

    def prefill_gamma(self, crsw=1.0):
        """What LwContext.formal_sol_gamma_matrices does before entering C++
        (Source/LwMiddleLayer.pyx:3198-3203): Gamma = crsw * C."""
        for a in self.atoms:
            if a.detailedStatic:
                continue
            if a.C is None:
                a.Gamma.fill(0.0)
            else:
                np.multiply(a.C, crsw, out=a.Gamma)

    def clone(self):
        """Deep copy of every buffer (for running two implementations on
        identical inputs)."""
        keep, self._keepalive = self._keepalive, []
        try:
            new = copy.deepcopy(self)
        finally:
            self._keepalive = keep
        return new

    def column(self, c):
        """A one-column Problem viewing (not copying) column ``c``."""
        def sl(a):
            return None if a is None else a[c:c + 1]
        atoms = []
        for a in self.atoms:
            trans = [TransitionData(type=t.type, i=t.i, j=t.j, Nblue=t.Nblue, Nred=t.Nred,
                                    lambda0=t.lambda0, wavelength=t.wavelength, Aji=t.Aji,
                                    Bji=t.Bji, Bij=t.Bij, dopplerWidth=t.dopplerWidth,
                                    alpha=t.alpha, phi=sl(t.phi), wphi=sl(t.wphi),
                                    rhoPrd=sl(t.rhoPrd), aDamp=sl(t.aDamp), Qelast=sl(t.Qelast),
                                    polProfiles=(None if t.polProfiles is None
                                                 else np.ascontiguousarray(t.polProfiles[:, c:c + 1])),
                                    polarised=t.polarised, zeeman=t.zeeman,
                                    Rij=sl(t.Rij),
                                    Rji=sl(t.Rji), name=t.name) for t in a.trans]
            atoms.append(AtomData(name=a.name, Nlevel=a.Nlevel, trans=trans, n=sl(a.n),
                                  nStar=sl(a.nStar), nTotal=sl(a.nTotal), vBroad=sl(a.vBroad),
                                  Gamma=sl(a.Gamma), C=sl(a.C), stages=a.stages, detailedStatic=a.detailedStatic))
        return Problem(Nspace=self.Nspace, Nrays=self.Nrays, height=sl(self.height),
                       temperature=sl(self.temperature), muz=self.muz, wmu=self.wmu,
                       wavelength=self.wavelength, chiBg=sl(self.chiBg), etaBg=sl(self.etaBg),
                       scaBg=sl(self.scaBg), atoms=atoms, vlosMu=sl(self.vlosMu), J=sl(self.J),
                       I=sl(self.I), formalSolver=self.formalSolver, lowerBc=self.lowerBc,
                       upperBc=self.upperBc, lowerBcData=sl(self.lowerBcData),
                       upperBcData=sl(self.upperBcData), lowerBcIdx=self.lowerBcIdx,
                       upperBcIdx=self.upperBcIdx, depthChi=sl(self.depthChi),
                       depthEta=sl(self.depthEta), depthI=sl(self.depthI), Quv=sl(self.Quv), ne=sl(self.ne),
                       vturb=sl(self.vturb), nHTot=sl(self.nHTot), meta=dict(self.meta),
                       hprd=None if self.hprd is None else self.hprd.column(c, self),
                       B=sl(self.B), cosGamma=sl(self.cosGamma), cos2chi=sl(self.cos2chi), sin2chi=sl(self.sin2chi))

    # ------------------------------------------------------------ marshalling
    def c_struct(self):
        """Build (and keep alive) the LwB200Problem describing these buffers."""
        keep = []
        p = capi.LwB200Problem()
        p.abiVersion = capi.ABI_VERSION
        p.Ncol, p.Nspace, p.Nrays, p.Nspect = self.Ncol, self.Nspace, self.Nrays, self.Nspect
        p.Natom = len(self.atoms)
        p.formalSolver = self.formalSolver
        p.lowerBc, p.upperBc = self.lowerBc, self.upperBc
        p.NlowerBcMu = 0 if self.lowerBcData is None else self.lowerBcData.shape[-1]
        p.NupperBcMu = 0 if self.upperBcData is None else self.upperBcData.shape[-1]
        d = capi.dptr
        p.height, p.temperature, p.vlosMu = d(self.height), d(self.temperature), d(self.vlosMu)
        p.muz, p.wmu, p.wavelength = d(self.muz), d(self.wmu), d(self.wavelength)
        p.chiBg, p.etaBg, p.scaBg = d(self.chiBg), d(self.etaBg), d(self.scaBg)
        p.lowerBcData, p.upperBcData = d(self.lowerBcData), d(self.upperBcData)
        p.lowerBcIdx, p.upperBcIdx = capi.iptr(self.lowerBcIdx), capi.iptr(self.upperBcIdx)
        p.J, p.I = d(self.J), d(self.I)
        p.depthChi, p.depthEta, p.depthI = d(self.depthChi), d(self.depthEta), d(self.depthI)
        p.Quv = d(self.Quv)
        p.ne = d(self.ne)
        atoms = (capi.LwB200Atom * len(self.atoms))()
        for ia, a in enumerate(self.atoms):
            ca = atoms[ia]
            ca.Nlevel, ca.Ntrans = a.Nlevel, a.Ntrans
            ca.detailedStatic = int(a.detailedStatic)
            trans = (capi.LwB200Transition * max(a.Ntrans, 1))()
            for it, t in enumerate(a.trans):
                ct = trans[it]
                ct.type, ct.i, ct.j = t.type, t.i, t.j
                ct.Nblue, ct.Nred = t.Nblue, t.Nred
                ct.Aji, ct.Bji, ct.Bij = t.Aji, t.Bji, t.Bij
                ct.lambda0, ct.dopplerWidth = t.lambda0, t.dopplerWidth
                ct.wavelength, ct.alpha = d(t.wavelength), d(t.alpha)
                ct.phi, ct.wphi, ct.rhoPrd, ct.aDamp = d(t.phi), d(t.wphi), d(t.rhoPrd), d(t.aDamp)
                ct.Rij, ct.Rji = d(t.Rij), d(t.Rji)
                ct.Qelast = d(t.Qelast)
                ct.polProfiles = d(t.polProfiles)
                ct.polarised = int(t.polarised or t.polProfiles is not None)
            keep.append(trans)
            ca.trans = C.cast(trans, C.POINTER(capi.LwB200Transition))
            ca.n, ca.nStar, ca.nTotal = d(a.n), d(a.nStar), d(a.nTotal)
            ca.vBroad, ca.Gamma = d(a.vBroad), d(a.Gamma)
            ca.C = d(a.C)
            ca.stages = d(a.stages)
        keep.append(atoms)
        p.atoms = C.cast(atoms, C.POINTER(capi.LwB200Atom))
        if self.hprd is not None:
            p.hprd = self.hprd.c_struct(keep)
        self._keepalive.append((p, keep))
        return p

    def configure_hprd(self, includeDetailed=False):
        """lw.Context.configure_hprd_coeffs: build the hybrid-PRD tables for the current velocity field
        (lwb200_configure_hprd, host code of the CUDA library) and switch the hybrid scheme on for this
        problem.  Returns the tables (None when the problem has no PRD line)."""
        lib = capi.load()
        keep, self.hprd = self.hprd, None
        try:
            cs = self.c_struct()
        finally:
            self.hprd = keep
        h = capi.LwB200HybridPrd()
        capi.check(lib.lwb200_configure_hprd(C.byref(cs), int(includeDetailed), C.byref(h)))
        try:
            self.hprd = HybridPrd.from_c(h, self) if h.Nlines > 0 else None
        finally:
            lib.lwb200_free_hprd(C.byref(h))
        return self.hprd

    # ----------------------------------------------------------- conveniences
    def active_atoms(self):
        return [a for a in self.atoms if not a.detailedStatic]

    def alg_bytes_per_iter(self):
        """Algorithmic (compulsory) HBM bytes of one Gamma iteration, every
        array once -- SURVEY.md 8(d):
        8*[P*M*2*K + 3*L*K + L*K + L*K + L*M + sum_a 2*N*K + sum_a 2*N^2*K
           + sum_t 2*K + Nlines*K + 2*K] per column."""
        K, L, M = self.Nspace, self.Nspect, self.Nrays
        P = sum(t.Nlambda for a in self.atoms for t in a.trans if t.type == capi.LINE)
        nlines = sum(1 for a in self.atoms for t in a.trans if t.type == capi.LINE)
        b = P * M * 2 * K + 5 * L * K + L * M + nlines * K + 2 * K
        for a in self.atoms:
            b += 2 * a.Nlevel * K + 2 * a.Ntrans * K
            if not a.detailedStatic:
                b += 2 * a.Nlevel * a.Nlevel * K
        return 8.0 * b * self.Ncol
