"""Synthetic FAL C-shaped inputs for the hot path (BASELINE.json ``configs``).

The reference's model atoms (``lightweaver/rh_atoms.py``) are not part of the
mounted tree and its Python package cannot be imported here, so the benchmark
and parity inputs are generated from first principles: the FAL C 82-point table
(data only, ``data/falc82.npz``), Saha-Boltzmann LTE populations, hydrogenic
cross-sections, RH-style line quadratures, Voigt profiles (scipy's Faddeeva
``wofz``) and a smooth H-minus-like background.  The *shapes* follow the
reference's configs (atoms, level/line/continuum counts, grid merge rule); the
atomic data are approximate literature values.  What matters for the hot path
is the numerical regime: optical depths from 1e-9 to 1e7, overlapping
transitions, LTE-to-NLTE departures -- all present.

Citations (structure only): heights from column mass ``atmosphere.py:1092-1107``;
Gauss-Legendre mu on [0, 1] ``atmosphere.py:1402-1408``; line quadrature
``atomic_model.py:279-341``; grid merge ``atomic_set.py:1048-1082``; LTE pops
``atomic_set.py:19-82``; profiles ``Source/FormalScalar.cpp:28-134``; vBroad
``atomic_model.py:84-86``.
"""
import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import capi
from .problem import AtomData, Problem, TransitionData

# Physical constants: same values as the reference's Source/Constants.hpp:6-47
CLIGHT = 2.99792458e8
HPLANCK = 6.6260755e-34
HC = HPLANCK * CLIGHT
KBOLTZMANN = 1.380658e-23
AMU = 1.6605402e-27
MELECTRON = 9.1093897e-31
NM_TO_M = 1e-9
VMICRO_CHAR = 3.0e3          # lightweaver/constants.py:26
SEED = 20251017


# ----------------------------------------------------------------- model atoms
@dataclass
class Level:
    E_cm: float   # cm^-1
    g: float
    stage: int


@dataclass
class LineSpec:
    j: int
    i: int
    Aji: float
    Nlambda: int
    qCore: float
    qWing: float


@dataclass
class ContSpec:
    j: int
    i: int
    alpha0: float      # m^2 at the edge
    Nlambda: int
    minLambda: float   # nm


@dataclass
class ModelAtom:
    name: str
    mass: float        # amu
    abundance: float   # relative to H
    levels: List[Level]
    lines: List[LineSpec]
    continua: List[ContSpec]

    def E_SI(self):
        return np.array([l.E_cm * 100.0 * HC for l in self.levels])


def h6_atom(nl=1.0):
    """Hydrogen, 5 bound levels + H II: 10 lines, 5 continua."""
    R = 109678.77
    lev = [Level(R * (1.0 - 1.0 / n**2), 2.0 * n * n, 0) for n in range(1, 6)] + [Level(R, 1.0, 1)]
    N = lambda x: int(round(x * nl)) | 1
    lines = [
        LineSpec(1, 0, 4.699e8, N(101), 15.0, 600.0),
        LineSpec(2, 0, 5.575e7, N(51), 10.0, 250.0),
        LineSpec(3, 0, 1.278e7, N(41), 5.0, 100.0),
        LineSpec(4, 0, 4.125e6, N(41), 5.0, 100.0),
        LineSpec(2, 1, 4.410e7, N(71), 3.0, 250.0),
        LineSpec(3, 1, 8.419e6, N(41), 3.0, 250.0),
        LineSpec(4, 1, 2.530e6, N(41), 3.0, 250.0),
        LineSpec(3, 2, 8.986e6, N(31), 2.0, 30.0),
        LineSpec(4, 2, 2.201e6, N(31), 2.0, 30.0),
        LineSpec(4, 3, 2.699e6, N(31), 2.0, 30.0),
    ]
    edges = [1e7 / (R - l.E_cm) for l in lev[:5]]
    continua = [ContSpec(5, n, 7.91e-22 * (n + 1), N(21) - 1, edges[n] / 4.0) for n in range(5)]
    return ModelAtom('H', 1.008, 1.0, lev, lines, continua)


def ca2_atom(nl=1.0):
    """Ca II, 5 levels + Ca III: H, K, infrared triplet, 5 continua."""
    lev = [Level(0.0, 2, 1), Level(13650.19, 4, 1), Level(13710.88, 6, 1),
           Level(25191.51, 2, 1), Level(25414.40, 4, 1), Level(95751.87, 1, 2)]
    N = lambda x: int(round(x * nl)) | 1
    lines = [
        LineSpec(3, 0, 1.40e8, N(101), 10.0, 300.0),   # H 396.8
        LineSpec(4, 0, 1.47e8, N(101), 10.0, 300.0),   # K 393.4
        LineSpec(3, 1, 1.06e7, N(71), 5.0, 150.0),     # 866.2
        LineSpec(4, 1, 1.11e6, N(71), 5.0, 150.0),     # 849.8
        LineSpec(4, 2, 9.90e6, N(71), 5.0, 150.0),     # 854.2
    ]
    a0 = [2.0e-23, 6.0e-22, 6.0e-22, 2.4e-22, 2.4e-22]
    continua = [ContSpec(5, n, a0[n], N(21) - 1, 1e7 / (lev[5].E_cm - lev[n].E_cm) / 2.5)
                for n in range(5)]
    return ModelAtom('Ca', 40.078, 10**(6.36 - 12.0), lev, lines, continua)


def mg2_atom(nl=1.0):
    """Mg II-like, 4 levels + Mg III: h, k and two overlapping subordinate lines."""
    lev = [Level(0.0, 2, 1), Level(35669.31, 2, 1), Level(35760.88, 4, 1),
           Level(71490.5, 10, 1), Level(121267.61, 1, 2)]
    N = lambda x: int(round(x * nl)) | 1
    lines = [
        LineSpec(1, 0, 2.57e8, N(101), 10.0, 400.0),   # h 280.27
        LineSpec(2, 0, 2.60e8, N(101), 10.0, 400.0),   # k 279.55
        LineSpec(3, 1, 4.01e8, N(51), 5.0, 60.0),      # 279.08
        LineSpec(3, 2, 4.79e8, N(51), 5.0, 60.0),      # 279.80
    ]
    a0 = [2.0e-23, 4.0e-23, 4.0e-23, 1.0e-22]
    continua = [ContSpec(4, n, a0[n], N(21) - 1, 1e7 / (lev[4].E_cm - lev[n].E_cm) / 2.5)
                for n in range(4)]
    return ModelAtom('Mg', 24.305, 10**(7.58 - 12.0), lev, lines, continua)


def na1_atom(nl=1.0):
    """Na I-like, 4 levels + Na II: D1, D2 and the 1138/1140 nm pair."""
    lev = [Level(0.0, 2, 0), Level(16956.17, 2, 0), Level(16973.37, 4, 0),
           Level(25739.99, 2, 0), Level(41449.45, 1, 1)]
    N = lambda x: int(round(x * nl)) | 1
    lines = [
        LineSpec(1, 0, 6.14e7, N(81), 8.0, 200.0),
        LineSpec(2, 0, 6.16e7, N(81), 8.0, 200.0),
        LineSpec(3, 1, 8.80e6, N(41), 3.0, 60.0),
        LineSpec(3, 2, 1.76e7, N(41), 3.0, 60.0),
    ]
    a0 = [1.2e-23, 8.0e-22, 8.0e-22, 1.0e-22]
    continua = [ContSpec(4, n, a0[n], N(21) - 1, 1e7 / (lev[4].E_cm - lev[n].E_cm) / 2.5)
                for n in range(4)]
    return ModelAtom('Na', 22.99, 10**(6.33 - 12.0), lev, lines, continua)


def he1_atom(nl=1.0):
    """He I-like, 5 levels + He II: 58.4, 1083, 2058 nm."""
    lev = [Level(0.0, 1, 0), Level(159855.97, 3, 0), Level(166277.44, 1, 0),
           Level(169087.0, 9, 0), Level(171134.90, 3, 0), Level(198310.67, 2, 1)]
    N = lambda x: int(round(x * nl)) | 1
    lines = [
        LineSpec(4, 0, 1.80e9, N(71), 10.0, 250.0),
        LineSpec(3, 1, 1.02e7, N(61), 5.0, 100.0),
        LineSpec(4, 2, 1.97e6, N(41), 3.0, 60.0),
    ]
    a0 = [7.4e-22, 5.5e-22, 9.3e-22, 1.4e-21, 1.3e-21]
    continua = [ContSpec(5, n, a0[n], N(21) - 1, 1e7 / (lev[5].E_cm - lev[n].E_cm) / 3.0)
                for n in range(5)]
    return ModelAtom('He', 4.003, 10**(10.99 - 12.0), lev, lines, continua)


# --------------------------------------------------------------- atmosphere
def load_falc82():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'falc82.npz')
    d = np.load(path)
    return {k: d[k] for k in d.files}


def _smooth_fields(rng, ncol, nfield, K, width=6.0):
    """Low-pass Gaussian random fields along depth, unit max-abs amplitude."""
    x = rng.standard_normal((ncol, nfield, K + 8 * int(width)))
    t = np.arange(-4 * int(width), 4 * int(width) + 1)
    kern = np.exp(-0.5 * (t / width)**2)
    kern /= kern.sum()
    out = np.empty((ncol, nfield, K))
    for c in range(ncol):
        for f in range(nfield):
            out[c, f] = np.convolve(x[c, f], kern, mode='valid')[:K]
    out /= np.max(np.abs(out), axis=-1, keepdims=True)
    return out


def falc_columns(ncol=1, perturb=False, seed=SEED, ndepth=None):
    """FAL C columns [ncol, K]: height from column mass (atmosphere.py:1092-1107),
    SI units.  ``perturb`` applies the seeded smooth perturbations of config 3
    (T*(1+0.05 g1), vz = 5 km/s g2, vturb*(1+0.2 g3), ne*exp(0.1 g4))."""
    f = load_falc82()
    cmass = f['cmass'] * 10.0              # g cm^-2 -> kg m^-2
    temp = f['temp'].copy()
    ne = f['ne'] * 1e6
    vturb = f['vturb'] * 1e3
    nh = f['nh'] * 1e6
    nHTot = nh.sum(axis=0)
    rho = AMU * 1.4271 * nHTot
    height = np.zeros_like(cmass)
    for k in range(1, cmass.shape[0]):
        height[k] = height[k - 1] - 2.0 * (cmass[k] - cmass[k - 1]) / (rho[k - 1] + rho[k])
    if ndepth is not None and ndepth != height.shape[0]:
        # the lw.benchmark() protocol interpolates FAL C to more depth points
        # (lightweaver/benchmark.py:19-45)
        s = np.linspace(0.0, 1.0, height.shape[0])
        sn = np.linspace(0.0, 1.0, ndepth)
        height = np.interp(sn, s, height)
        temp = np.exp(np.interp(sn, s, np.log(temp)))
        ne = np.exp(np.interp(sn, s, np.log(ne)))
        vturb = np.interp(sn, s, vturb)
        nHTot = np.exp(np.interp(sn, s, np.log(nHTot)))
    K = height.shape[0]
    out = {
        'height': np.tile(height, (ncol, 1)), 'temperature': np.tile(temp, (ncol, 1)),
        'ne': np.tile(ne, (ncol, 1)), 'vturb': np.tile(vturb, (ncol, 1)),
        'nHTot': np.tile(nHTot, (ncol, 1)), 'vz': np.zeros((ncol, K)),
    }
    if perturb:
        rng = np.random.default_rng(seed)
        g = _smooth_fields(rng, ncol, 4, K)
        out['temperature'] *= 1.0 + 0.05 * g[:, 0]
        np.maximum(out['temperature'], 2500.0, out=out['temperature'])
        out['vz'] = 5.0e3 * g[:, 1]
        out['vturb'] *= 1.0 + 0.2 * g[:, 2]
        out['ne'] *= np.exp(0.1 * g[:, 3])
    return out


def gauss_legendre_mu(nrays):
    x, w = np.polynomial.legendre.leggauss(nrays)
    return 0.5 + 0.5 * x, 0.5 * w


# ------------------------------------------------------------------ physics
def lte_pops(atom: ModelAtom, temperature, ne, nTotal):
    """Saha-Boltzmann populations [ncol, Nlevel, K] (no Debye shielding)."""
    E = atom.E_SI()
    g = np.array([l.g for l in atom.levels])
    st = np.array([l.stage for l in atom.levels])
    c1 = (HPLANCK / (2.0 * np.pi * MELECTRON)) * (HPLANCK / KBOLTZMANN)
    cNe_T = 0.5 * ne * (c1 / temperature)**1.5
    dE = (E - E[0])[None, :, None] / (KBOLTZMANN * temperature[:, None, :])
    ratio = (g / g[0])[None, :, None] * np.exp(-dE) / cNe_T[:, None, :]**(st - st[0])[None, :, None]
    n0 = nTotal / ratio.sum(axis=1)
    return ratio * n0[:, None, :]


def line_quadrature(lambda0, Nlambda, qCore, qWing):
    """RH-style grid, linear core + exponential wings, in nm
    (q(n) = a (n + exp(b n) - 1); atomic_model.py:279-341)."""
    half = Nlambda // 2 + 1
    beta = 1.0 if qWing <= 2.0 * qCore else qWing / (2.0 * qCore)
    y = beta + np.sqrt(beta**2 + (beta - 1.0) * half + 2.0 - 3.0 * beta)
    b = 2.0 * np.log(y) / (half - 1)
    a = qWing / (half - 2.0 + y**2)
    nn = np.arange(half)
    q = a * (nn + (np.exp(b * nn) - 1.0))
    full = np.concatenate((-q[1:][::-1], q))
    return lambda0 + full * lambda0 * (VMICRO_CHAR / CLIGHT)


def voigt_profile(wavelength, lambda0, aDamp, vBroad, vlosMu):
    """phi[ncol, Nl, Nrays, 2, K] = H(a, v) / (sqrt(pi) vBroad)
    (Transition::compute_phi_la, Source/FormalScalar.cpp:28-51)."""
    from scipy.special import wofz
    vBase = (wavelength - lambda0) * CLIGHT / lambda0              # [Nl]
    sign = np.array([-1.0, 1.0])
    v = (vBase[None, :, None, None, None]
         + sign[None, None, None, :, None] * vlosMu[:, None, :, None, :]) / vBroad[:, None, None, None, :]
    a = np.broadcast_to(aDamp[:, None, None, None, :], v.shape)
    H = wofz(v + 1j * a).real
    return np.ascontiguousarray(H / (np.sqrt(np.pi) * vBroad[:, None, None, None, :]))


QELECTRON = 1.60217733E-19
MELECTRON = 9.1093897E-31


def polarised_profiles(wavelength, lambda0, aDamp, vBroad, vlosMu, B, cosGamma, cos2chi, sin2chi, gEff):
    """Transition::compute_polarised_profiles (Source/FormalStokes.cpp:9-117) for a normal Zeeman
    triplet (alpha = -1, 0, +1, unit strengths, shifts -gEff, 0, +gEff).  Returns
    (phi [ncol, Nl, M, 2, K], pol [6, ncol, Nl, M, 2, K] = phiQ, phiU, phiV, psiQ, psiU, psiV).
    B: [ncol, K] Tesla; cosGamma, cos2chi, sin2chi: [ncol, M, K]."""
    from scipy.special import wofz
    larmor = QELECTRON / (4.0 * np.pi * MELECTRON) * (lambda0 * NM_TO_M)
    vB = larmor * B / vBroad                                          # [ncol, K]
    sv = 1.0 / (np.sqrt(np.pi) * vBroad)
    vBase = (wavelength - lambda0) * CLIGHT / lambda0
    sign = np.array([-1.0, 1.0])
    v = (vBase[None, :, None, None, None]
         + sign[None, None, None, :, None] * vlosMu[:, None, :, None, :]) / vBroad[:, None, None, None, :]
    a = np.broadcast_to(aDamp[:, None, None, None, :], v.shape)
    comp = {}
    for alpha, shift in ((-1, -gEff), (0, 0.0), (1, gEff)):
        w = wofz((v - shift * vB[:, None, None, None, :]) + 1j * a)
        comp[alpha] = (w.real, w.imag)
    phi_sb, psi_sb = comp[-1]
    phi_pi, psi_pi = comp[0]
    phi_sr, psi_sr = comp[1]
    cg = cosGamma[:, None, :, None, :]
    sin2g = 1.0 - cg**2
    c2 = cos2chi[:, None, :, None, :]
    s2 = sin2chi[:, None, :, None, :]
    s = sign[None, None, None, :, None]
    svb = sv[:, None, None, None, :]
    phi_sigma = phi_sr + phi_sb
    phi_delta = 0.5 * phi_pi - 0.25 * phi_sigma
    psi_sigma = psi_sr + psi_sb
    psi_delta = 0.5 * psi_pi - 0.25 * psi_sigma
    phi = (phi_delta * sin2g + 0.5 * phi_sigma) * svb
    pol = np.stack([s * phi_delta * sin2g * c2 * svb, phi_delta * sin2g * s2 * svb + 0.0 * s,
                    s * 0.5 * (phi_sr - phi_sb) * cg * svb,
                    s * psi_delta * sin2g * c2 * svb, psi_delta * sin2g * s2 * svb + 0.0 * s,
                    s * 0.5 * (psi_sr - psi_sb) * cg * svb])
    return np.ascontiguousarray(phi), np.ascontiguousarray(pol)


def planck_nu(wavelength_nm, T):
    """B_nu [J s^-1 m^-2 sr^-1 Hz^-1] (planck_nu, Source/LwMisc.hpp:29-46)."""
    x = HC / (KBOLTZMANN * NM_TO_M) / wavelength_nm / T
    twohnu3_c2 = 2.0 * HC / NM_TO_M**3 / wavelength_nm**3
    with np.errstate(over='ignore'):
        return np.where(x <= 150.0, twohnu3_c2 / np.expm1(np.minimum(x, 150.0)), 0.0)


def background(wavelength, temperature, ne, nHTot, metal_scale=1.0, chunk=64):
    """Smooth H-minus-like continuous absorption + Thomson/Rayleigh scattering.
    Returns chi (absorption + scattering), eta (thermal), sca, each [ncol, L, K]."""
    ncol = temperature.shape[0]
    if ncol > chunk:
        shape = (ncol, wavelength.shape[0], temperature.shape[1])
        chi, eta, sca = np.empty(shape), np.empty(shape), np.empty(shape)
        for c0 in range(0, ncol, chunk):
            sl = slice(c0, min(c0 + chunk, ncol))
            chi[sl], eta[sl], sca[sl] = background(wavelength, temperature[sl], ne[sl], nHTot[sl],
                                                   metal_scale, chunk)
        return chi, eta, sca
    lam = wavelength[None, :, None]
    T = temperature[:, None, :]
    theta = 5040.0 / T
    # H- bound-free/free-free like: peaks near 850 nm, falls in the UV and IR
    shape = (lam / 850.0) * np.exp(-0.5 * ((np.log(lam / 850.0)) / 1.2)**2) + 0.15 * (lam / 1600.0)**2
    kap_hm = 2.5e-48 * theta**2.5 * 10.0**(0.754 * (theta - 0.84)) * shape
    # metal-like photoionisation opacity rising into the UV
    kap_uv = metal_scale * 4.0e-28 * np.exp(-lam / 90.0) * np.exp(-3.0 * (theta - 0.84))
    nH = nHTot[:, None, :]
    chi_abs = nH * ne[:, None, :] * kap_hm + nH * kap_uv
    sca = 6.652e-29 * ne[:, None, :] + nH * 5.8e-32 * (121.6 / np.maximum(lam, 121.6))**4 * 1e3
    eta = chi_abs * planck_nu(lam, T)
    return (np.ascontiguousarray(chi_abs + sca), np.ascontiguousarray(eta),
            np.ascontiguousarray(sca * np.ones_like(chi_abs)))


def collision_matrix(atom: ModelAtom, temperature, ne, nStar, rng):
    """Synthetic detailed-balance collisional rates C[ncol, to, from, K]:
    C(i<-j) = ne q_ij sqrt(5000/T), C(j<-i) = C(i<-j) n*_j/n*_i."""
    N = len(atom.levels)
    ncol, K = temperature.shape
    Cm = np.zeros((ncol, N, N, K))
    for j in range(N):
        for i in range(j):
            q = 10.0**rng.uniform(-15.0, -13.5)
            down = ne * q * np.sqrt(5000.0 / temperature)
            Cm[:, i, j] = down
            Cm[:, j, i] = down * nStar[:, j] / nStar[:, i]
    return Cm


# ------------------------------------------------------------- the generator
def build_problem(atoms: Sequence[ModelAtom], ncol=1, nrays=5, perturb=False, seed=SEED,
                  formal_solver=capi.FS_BEZIER3, ndepth=None, detailed: Sequence[str] = (),
                  with_profiles=True, lambda_reference=500.0, col_range=None,
                  alloc_phi=True, prd=None, polarised=None, Bmax=0.15) -> Problem:
    """Assemble a Problem for ``atoms`` in ``ncol`` FAL C columns.  ``col_range``
    = (c0, c1) keeps only that slice of the ``ncol`` columns (one column shard of
    a multi-GPU run; every rank sees the same seeded stack).  ``with_profiles``
    False leaves phi/wphi zero (to be made on the device); ``alloc_phi`` False
    does not even allocate host phi.  ``prd``: {atom name: [line indices]} treated with
    angle-averaged PRD (rhoPrd = 1 to start with, Qelast = the collisional part of the damping).
    ``polarised``: {atom name: [line indices]} given Zeeman-split polarised profiles in a smooth
    synthetic magnetic field of up to ``Bmax`` Tesla (seeded, different in every column)."""
    rng = np.random.default_rng(seed + 1)
    atm = falc_columns(ncol, perturb=perturb, seed=seed, ndepth=ndepth)
    ncol_full = ncol
    if col_range is not None:
        atm = {k: np.ascontiguousarray(v[col_range[0]:col_range[1]]) for k, v in atm.items()}
        ncol = col_range[1] - col_range[0]
    T, ne, nHTot, vturb = atm['temperature'], atm['ne'], atm['nHTot'], atm['vturb']
    K = T.shape[1]
    muz, wmu = gauss_legendre_mu(nrays)
    vlosMu = np.ascontiguousarray(muz[None, :, None] * atm['vz'][:, None, :])

    if polarised:
        # (drawn for the whole stack, then cut: a column is the same whichever shard holds it)
        prng = np.random.default_rng(seed + 7)
        cut = slice(None) if col_range is None else slice(col_range[0], col_range[1])
        zn = np.linspace(0.0, 1.0, K)[None, :]
        amp = prng.uniform(0.3, 1.0, (ncol_full, 1))[cut]
        Bfield = Bmax * amp * (0.4 + 0.6 * zn)                       # stronger with depth
        gamma0, gamma1 = prng.uniform(0.2, 1.3, (ncol_full, 1))[cut], prng.uniform(0, 6, (ncol_full, 1))[cut]
        gammaB = gamma0 + 0.3 * np.sin(2.0 * np.pi * zn + gamma1)
        chiB = prng.uniform(0.0, np.pi, (ncol_full, 1))[cut] + 0.5 * zn
        cosGamma = np.ascontiguousarray(muz[None, :, None] * np.cos(gammaB)[:, None, :])
        cos2chi = np.ascontiguousarray(np.broadcast_to(np.cos(2.0 * chiB)[:, None, :], cosGamma.shape))
        sin2chi = np.ascontiguousarray(np.broadcast_to(np.sin(2.0 * chiB)[:, None, :], cosGamma.shape))

    # --- wavelength grids (atomic_set.py:1048-1082) ---
    grids, owners = [], []
    for ia, atom in enumerate(atoms):
        E = atom.E_SI()
        for ln in atom.lines:
            lam0 = HC / (E[ln.j] - E[ln.i]) / NM_TO_M
            grids.append(line_quadrature(lam0, ln.Nlambda, ln.qCore, ln.qWing))
            owners.append((ia, 'line', ln, lam0))
        for ct in atom.continua:
            lam0 = HC / (E[ct.j] - E[ct.i]) / NM_TO_M
            grids.append(np.linspace(ct.minLambda, lam0, ct.Nlambda))
            owners.append((ia, 'cont', ct, lam0))
    grid = np.unique(np.concatenate(grids + [np.array([lambda_reference])]))
    L = grid.shape[0]

    chiBg, etaBg, scaBg = background(grid, T, ne, nHTot)

    atom_data = [None] * len(atoms)
    trans_lists = [[] for _ in atoms]
    for (ia, kind, spec, lam0), g in zip(owners, grids):
        atom = atoms[ia]
        Nblue = int(np.searchsorted(grid, g[0]))
        Nred = int(np.searchsorted(grid, g[-1])) + 1
        wl = np.ascontiguousarray(grid[Nblue:Nred])
        if kind == 'line':
            gj, gi = atom.levels[spec.j].g, atom.levels[spec.i].g
            Bji = spec.Aji * (lam0 * NM_TO_M)**3 / (2.0 * HC)
            Bij = gj / gi * Bji
            t = TransitionData(type=capi.LINE, i=spec.i, j=spec.j, Nblue=Nblue, Nred=Nred,
                               lambda0=lam0, wavelength=wl, Aji=spec.Aji, Bji=Bji, Bij=Bij,
                               dopplerWidth=CLIGHT / lam0, name=f'{atom.name} {lam0:.2f}')
        else:
            alpha = spec.alpha0 * (wl / lam0)**3
            alpha[(wl < spec.minLambda) | (wl > lam0)] = 0.0
            t = TransitionData(type=capi.CONTINUUM, i=spec.i, j=spec.j, Nblue=Nblue, Nred=Nred,
                               lambda0=lam0, wavelength=wl, alpha=np.ascontiguousarray(alpha),
                               dopplerWidth=1.0, name=f'{atom.name} bf {lam0:.2f}')
        trans_lists[ia].append(t)

    for ia, atom in enumerate(atoms):
        nTotal = np.ascontiguousarray(atom.abundance * nHTot)
        nStar = np.ascontiguousarray(lte_pops(atom, T, ne, nTotal))
        vBroad = np.ascontiguousarray(np.sqrt(2.0 * KBOLTZMANN * T / (AMU * atom.mass) + vturb**2))
        Cm = collision_matrix(atom, T, ne, nStar, rng)
        for t in trans_lists[ia]:
            if t.type != capi.LINE:
                continue
            gRad = sum(l.Aji for l in atom.lines if l.j in (t.j, t.i))
            gamma = gRad + 1.0e-14 * nHTot * (T / 5000.0)**0.38 + 4.0e-13 * ne
            dnuD = vBroad / (t.lambda0 * NM_TO_M)
            t.aDamp = np.ascontiguousarray(np.clip(gamma / (4.0 * np.pi * dnuD), 1e-4, 1e-1))
            line_idx = [x for x in trans_lists[ia] if x.type == capi.LINE].index(t)
            if prd and line_idx in prd.get(atom.name, ()):
                t.rhoPrd = np.ones((ncol, t.Nlambda, K))
                t.Qelast = np.ascontiguousarray(gamma - gRad)
            t.wphi = np.zeros((ncol, K))
            is_pol = bool(polarised) and line_idx in polarised.get(atom.name, ())
            if is_pol:
                t.polarised = True   # (a normal Zeeman triplet)
                t.zeeman = (np.array([-1, 0, 1], dtype=np.int32), np.array([-1.1, 0.0, 1.1]), np.ones(3))
            if is_pol and with_profiles:
                t.phi, t.polProfiles = polarised_profiles(t.wavelength, t.lambda0, t.aDamp, vBroad, vlosMu,
                                                          Bfield, cosGamma, cos2chi, sin2chi, gEff=1.1)
                wlam = t.wlambda()
                s = np.einsum('clmdk,l,m->ck', t.phi, wlam, 0.5 * wmu)
                t.wphi = np.ascontiguousarray(1.0 / s)
            elif with_profiles:
                t.phi = voigt_profile(t.wavelength, t.lambda0, t.aDamp, vBroad, vlosMu)
                wlam = t.wlambda()
                s = np.einsum('clmdk,l,m->ck', t.phi, wlam, 0.5 * wmu)
                t.wphi = np.ascontiguousarray(1.0 / s)
            elif alloc_phi:
                t.phi = np.zeros((ncol, t.Nlambda, nrays, 2, K))
            else:
                t.wphi = None
        is_detailed = atom.name in detailed
        atom_data[ia] = AtomData(name=atom.name, Nlevel=len(atom.levels), trans=trans_lists[ia],
                                 n=nStar.copy(), nStar=nStar, nTotal=nTotal, vBroad=vBroad,
                                 C=None if is_detailed else Cm, detailedStatic=is_detailed,
                                 stages=np.array([float(lv.stage) for lv in atom.levels]))

    prob = Problem(Nspace=K, Nrays=nrays, height=np.ascontiguousarray(atm['height']),
                   temperature=np.ascontiguousarray(T), muz=muz, wmu=wmu, wavelength=grid,
                   chiBg=chiBg, etaBg=etaBg, scaBg=scaBg, atoms=atom_data, vlosMu=vlosMu,
                   formalSolver=formal_solver, ne=np.ascontiguousarray(ne),
                   vturb=np.ascontiguousarray(vturb), nHTot=np.ascontiguousarray(nHTot))
    if polarised:
        prob.Quv = np.zeros((ncol, 3, L, nrays))
        prob.B = np.ascontiguousarray(Bfield)
        prob.cosGamma, prob.cos2chi, prob.sin2chi = cosGamma, cos2chi, sin2chi
    prob.meta = {'atoms': [a.name for a in atoms], 'perturb': perturb, 'seed': seed}
    prob.prefill_gamma()
    return prob


def config_c1(ncol=1, nrays=5, perturb=False, seed=SEED, nl=1.0, **kw) -> Problem:
    """Config 1: FAL C, H 6-level + Ca II 5+1, 5 rays."""
    return build_problem([h6_atom(nl), ca2_atom(nl)], ncol=ncol, nrays=nrays, perturb=perturb,
                         seed=seed, **kw)


def config_c2(nrays=10, nl=4.8, seed=SEED, **kw) -> Problem:
    """Config 2: FAL C, H + Ca II + Mg II + Na I + He I active (Fe only through
    the background), ~1e4 wavelengths, 10 rays."""
    atoms = [h6_atom(nl), ca2_atom(nl), mg2_atom(nl), na1_atom(nl), he1_atom(nl)]
    return build_problem(atoms, ncol=1, nrays=nrays, seed=seed, **kw)


def config_c4(nrays=5, nl=1.0, seed=SEED, ncol=1, **kw) -> Problem:
    """Config 4: FAL C, H + Ca II + Mg II with the Mg II h and k lines in angle-averaged PRD.  With
    perturb=True the column carries the velocity field of config 3: the case for the hybrid scheme
    (Problem.configure_hprd)."""
    atoms = [h6_atom(nl), ca2_atom(nl), mg2_atom(nl)]
    return build_problem(atoms, ncol=ncol, nrays=nrays, seed=seed, prd={'Mg': [0, 1]}, **kw)


def config_c3(ncol=4096, nrays=5, seed=SEED, **kw) -> Problem:
    """Config 3: stack of perturbed FAL C columns, H + Ca II."""
    return config_c1(ncol=ncol, nrays=nrays, perturb=True, seed=seed, **kw)


def tiny_problem(ncol=1, nrays=3, seed=SEED, ndepth=None, perturb=False, **kw) -> Problem:
    """A small case for quick tests: a 3-level + continuum toy atom."""
    lev = [Level(0.0, 2, 0), Level(60000.0, 6, 0), Level(75000.0, 10, 0), Level(100000.0, 1, 1)]
    lines = [LineSpec(1, 0, 3.0e8, 21, 5.0, 60.0), LineSpec(2, 0, 4.0e7, 15, 3.0, 30.0),
             LineSpec(2, 1, 2.0e7, 15, 3.0, 30.0)]
    cont = [ContSpec(3, 0, 6.0e-22, 8, 40.0), ContSpec(3, 1, 1.2e-21, 8, 100.0),
            ContSpec(3, 2, 2.0e-21, 8, 150.0)]
    toy = ModelAtom('Toy', 12.0, 1e-4, lev, lines, cont)
    return build_problem([toy], ncol=ncol, nrays=nrays, seed=seed, ndepth=ndepth,
                         perturb=perturb, **kw)


def nine_level_problem(ncol=1, nrays=3, seed=SEED, ndepth=None, perturb=False, **kw) -> Problem:
    """An 8-level + continuum toy atom: beyond the register-resident population solver (N <= 7), so the
    general per-depth LU path is exercised."""
    E = [0.0, 60000.0, 70000.0, 75000.0, 80000.0, 83000.0, 86000.0, 88000.0]
    lev = [Level(e, 2.0 * (q + 1), 0) for q, e in enumerate(E)] + [Level(100000.0, 1, 1)]
    lines = [LineSpec(1, 0, 3.0e8, 15, 5.0, 60.0), LineSpec(2, 0, 4.0e7, 11, 3.0, 30.0),
             LineSpec(3, 1, 2.0e7, 11, 3.0, 30.0), LineSpec(4, 1, 1.0e7, 11, 3.0, 30.0),
             LineSpec(5, 2, 8.0e6, 11, 3.0, 30.0), LineSpec(6, 3, 6.0e6, 11, 3.0, 30.0),
             LineSpec(7, 4, 5.0e6, 11, 3.0, 30.0)]
    cont = [ContSpec(8, q, 6.0e-22 * (q + 1), 6, 0.4 * 1e7 / (100000.0 - e)) for q, e in enumerate(E)]
    toy = ModelAtom('Toy9', 12.0, 1e-4, lev, lines, cont)
    return build_problem([toy], ncol=ncol, nrays=nrays, seed=seed, ndepth=ndepth, perturb=perturb, **kw)


def config_c5(ncol=1024, nrays=5, seed=SEED, nl=1.0, **kw) -> Problem:
    """Config 5: magnetised stack of perturbed FAL C columns, Ca II with the 854.2 nm line
    Zeeman-split and polarised (full Stokes)."""
    return build_problem([ca2_atom(nl)], ncol=ncol, nrays=nrays, perturb=True, seed=seed,
                         polarised={'Ca': [4]}, **kw)


def tiny_stokes_problem(ncol=1, nrays=3, **kw) -> Problem:
    """tiny_problem with its 2-1 subordinate line polarised (the two resonance lines stay scalar, so
    polarised and unpolarised wavelengths, overlapping and not, all occur)."""
    return tiny_problem(ncol=ncol, nrays=nrays, polarised={'Toy': [2]}, **kw)


def tiny_prd_problem(ncol=1, nrays=3, **kw) -> Problem:
    """tiny_problem with its two resonance lines in angle-averaged PRD."""
    return tiny_problem(ncol=ncol, nrays=nrays, prd={'Toy': [0, 1]}, **kw)
