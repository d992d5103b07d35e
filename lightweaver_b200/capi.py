"""ctypes binding of the C-ABI declared in ``include/lwb200.h``.

This is the whole Python <-> CUDA boundary: plain structs of pointers and sizes,
int return codes.  The library is built in-tree by ``__graft_entry__.build()``
(``lightweaver_b200/csrc/build.py``) as ``lightweaver_b200/liblwb200.so``; if it
is missing, loading fails loudly -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

ABI_VERSION = 6

LINE, CONTINUUM = 0, 1
BC_UNINITIALISED, BC_ZERO, BC_THERMALISED, BC_PERIODIC, BC_CALLABLE = range(5)
FS_LINEAR, FS_BESSER, FS_BEZIER3 = 0, 1, 2
FS_NAMES = {'piecewise_linear_1d': FS_LINEAR, 'piecewise_besser_1d': FS_BESSER,
            'piecewise_bezier3_1d': FS_BEZIER3}

(ATMOS, BACKGR, POPS, NSTAR, GAMMA, JBAR, PROFILE, INTENS, RATES, DEPTH, ADAMP,
 GAMMA_FINAL, PRD, STOKES, OWN_ROWS, ZPLANE, COLLISIONS, POLPROF) = (1 << i for i in range(18))
ALL_INPUTS = 0x7f
ITER_INPUTS = POPS | NSTAR | GAMMA
ITER_OUTPUTS = GAMMA | JBAR | INTENS | RATES

LAMBDA_ITERATE, STORE_DEPTH, DEFER_FINALISE, GENERAL_KERNEL, FETCH_EARLY, DJ_ASYNC = 1, 2, 4, 8, 16, 32

BUF_ACCUM, BUF_J, BUF_I, BUF_POPS, BUF_GAMMA, BUF_DJ, BUF_DJMAX = range(7)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class LwB200Transition(C.Structure):
    _fields_ = [
        ('type', C.c_int32), ('i', C.c_int32), ('j', C.c_int32),
        ('Nblue', C.c_int32), ('Nred', C.c_int32), ('polarised', C.c_int32),
        ('Aji', C.c_double), ('Bji', C.c_double), ('Bij', C.c_double),
        ('lambda0', C.c_double), ('dopplerWidth', C.c_double),
        ('wavelength', _dp), ('alpha', _dp), ('phi', _dp), ('wphi', _dp),
        ('rhoPrd', _dp), ('aDamp', _dp), ('Rij', _dp), ('Rji', _dp), ('Qelast', _dp),
        ('polProfiles', _dp),
    ]


class LwB200Atom(C.Structure):
    _fields_ = [
        ('Nlevel', C.c_int32), ('Ntrans', C.c_int32),
        ('detailedStatic', C.c_int32), ('reserved', C.c_int32),
        ('trans', C.POINTER(LwB200Transition)),
        ('n', _dp), ('nStar', _dp), ('nTotal', _dp), ('vBroad', _dp), ('Gamma', _dp), ('C', _dp),
        ('stages', _dp),
    ]


_lp = C.POINTER(C.c_int64)


class LwB200HybridPrd(C.Structure):
    _fields_ = [
        ('NprdLa', C.c_int32), ('NhPrd', C.c_int32), ('Nlines', C.c_int32), ('reserved', C.c_int32),
        ('prdLaOfLa', _ip), ('hPrdLaOfLa', _ip), ('JRest', _dp),
        ('JCoeffOff', _lp), ('JCoeffIdx', _ip), ('JCoeffFrac', _dp),
        ('lineAtom', _ip), ('lineTrans', _ip), ('rhoCoefOff', _lp), ('rhoFrac', _dp), ('rhoI0', _ip),
    ]


class LwB200Zeeman(C.Structure):
    _fields_ = [
        ('atom', C.c_int32), ('trans', C.c_int32), ('Ncomponent', C.c_int32), ('reserved', C.c_int32),
        ('alpha', _ip), ('shift', _dp), ('strength', _dp),
    ]


class LwB200Problem(C.Structure):
    _fields_ = [
        ('abiVersion', C.c_int32), ('Ncol', C.c_int32), ('Nspace', C.c_int32),
        ('Nrays', C.c_int32), ('Nspect', C.c_int32), ('Natom', C.c_int32),
        ('formalSolver', C.c_int32), ('lowerBc', C.c_int32), ('upperBc', C.c_int32),
        ('NlowerBcMu', C.c_int32), ('NupperBcMu', C.c_int32), ('reserved', C.c_int32),
        ('height', _dp), ('temperature', _dp), ('vlosMu', _dp), ('muz', _dp), ('wmu', _dp),
        ('wavelength', _dp), ('chiBg', _dp), ('etaBg', _dp), ('scaBg', _dp),
        ('lowerBcData', _dp), ('upperBcData', _dp), ('lowerBcIdx', _ip), ('upperBcIdx', _ip),
        ('J', _dp), ('I', _dp), ('depthChi', _dp), ('depthEta', _dp), ('depthI', _dp),
        ('atoms', C.POINTER(LwB200Atom)), ('Quv', _dp), ('ne', _dp),
        ('hprd', C.POINTER(LwB200HybridPrd)),
    ]


class LwB200NrUpdate(C.Structure):
    _fields_ = [
        ('Natom', C.c_int32), ('timeDependent', C.c_int32), ('atomIdx', _ip),
        ('dC', C.POINTER(_dp)), ('backgroundNe', _dp), ('nPrev', C.POINTER(_dp)),
        ('dt', C.c_double), ('crswVal', C.c_double),
    ]


def make_nr_update(atomIdx, backgroundNe, dC=None, nPrev=None, dt=0.0, crswVal=1.0):
    """(LwB200NrUpdate, keepalive) from numpy arrays: atomIdx list of ints, backgroundNe [Ncol, K],
    dC / nPrev optional lists (one array per atom)."""
    u = LwB200NrUpdate()
    idx = np.ascontiguousarray(atomIdx, dtype=np.int32)
    keep = [idx, backgroundNe]
    u.Natom = len(idx)
    u.atomIdx = iptr(idx)
    u.backgroundNe = dptr(backgroundNe)
    u.timeDependent = int(nPrev is not None)
    u.dt, u.crswVal = float(dt), float(crswVal)
    for name, arrs in (('dC', dC), ('nPrev', nPrev)):
        if arrs is None:
            setattr(u, name, C.POINTER(_dp)())
            continue
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        ptrs = (_dp * len(arrs))(*[dptr(a) for a in arrs])
        keep += [arrs, ptrs]
        setattr(u, name, C.cast(ptrs, C.POINTER(_dp)))
    return u, keep


def dptr(a):
    """Pointer to a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return _dp()
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags['C_CONTIGUOUS'], \
        'expected a C-contiguous float64 ndarray'
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return _ip()
    assert isinstance(a, np.ndarray) and a.dtype == np.int32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_ip)


class LwB200Error(RuntimeError):
    pass


_lib = None


def lib_path():
    # LWB200_LIB: development aid for timing kernel variants built side by side
    # (tools/variants.py); it changes which build of the same C-ABI is loaded, nothing else
    override = os.environ.get('LWB200_LIB')
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'liblwb200.so')


def load():
    """Load liblwb200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise LwB200Error(
            f'{path} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU fallback for the B200 back end)')
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.lwb200_last_error.restype = C.c_char_p
    lib.lwb200_abi_version.restype = C.c_int
    lib.lwb200_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.lwb200_create.argtypes = [C.POINTER(LwB200Problem), C.c_int, C.POINTER(vp)]
    lib.lwb200_destroy.argtypes = [vp]
    lib.lwb200_set_stream.argtypes = [vp, vp]
    lib.lwb200_set_lambda_range.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lwb200_set_active_columns.argtypes = [vp, C.POINTER(C.c_uint8)]
    lib.lwb200_ng_configure.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.lwb200_ng_accelerate.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_ng_clear.argtypes = [vp]
    lib.lwb200_last_ng.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_upload.argtypes = [vp, C.c_uint32]
    lib.lwb200_download.argtypes = [vp, C.c_uint32]
    lib.lwb200_sync.argtypes = [vp]
    lib.lwb200_compute_profiles.argtypes = [vp]
    lib.lwb200_fs_iter.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_finalise.argtypes = [vp]
    lib.lwb200_set_collision_prefill.argtypes = [vp, C.c_int, C.c_double]
    lib.lwb200_set_j20.argtypes = [vp, _dp]
    lib.lwb200_compute_polarised_profiles.argtypes = [vp, C.POINTER(LwB200Zeeman), C.c_int32, _dp, _dp, _dp, _dp]
    lib.lwb200_set_hybrid_prd.argtypes = [vp, C.POINTER(LwB200HybridPrd)]
    lib.lwb200_configure_hprd.argtypes = [C.POINTER(LwB200Problem), C.c_int, C.POINTER(LwB200HybridPrd)]
    lib.lwb200_free_hprd.argtypes = [C.POINTER(LwB200HybridPrd)]
    lib.lwb200_free_hprd.restype = None
    lib.lwb200_dj_max.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_formal_sol.argtypes = [vp, C.c_int]
    lib.lwb200_stat_eq.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.lwb200_redistribute_prd.argtypes = [vp, C.c_int32, C.c_double, C.c_int32, C.POINTER(C.c_int32), _dp,
                                            _ip, _dp, C.POINTER(C.c_int64)]
    lib.lwb200_formal_sol_full_stokes.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_nr_post_update.argtypes = [vp, C.POINTER(LwB200NrUpdate), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.lwb200_stat_eq_async.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.lwb200_last_singular.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.lwb200_last_dj.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.lwb200_time_dep_update.argtypes = [vp, C.c_int32, _dp, C.c_double, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_int32)]
    lib.lwb200_set_zplane.argtypes = [vp, _dp, _dp]
    lib.lwb200_population_solve.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, C.c_double,
                                            C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.lwb200_global_launch_count.restype = C.c_int64
    lib.lwb200_kernel_time.argtypes = [vp, C.POINTER(C.c_double)]
    lib.lwb200_device_buffer.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.lwb200_work_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int64)]
    for name in ('device_count', 'create', 'destroy', 'set_stream', 'set_lambda_range', 'upload',
                 'download', 'sync', 'compute_profiles', 'fs_iter', 'finalise', 'dj_max',
                 'formal_sol', 'stat_eq', 'device_buffer', 'work_stats', 'kernel_time', 'redistribute_prd', 'time_dep_update', 'formal_sol_full_stokes', 'nr_post_update', 'stat_eq_async', 'last_singular', 'last_dj', 'set_zplane', 'population_solve', 'set_collision_prefill', 'set_hybrid_prd', 'configure_hprd', 'set_j20', 'compute_polarised_profiles'):
        getattr(lib, 'lwb200_' + name).restype = C.c_int
    if lib.lwb200_abi_version() != ABI_VERSION:
        raise LwB200Error('liblwb200.so ABI version mismatch; rebuild')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().lwb200_last_error()
        raise LwB200Error(msg.decode() if msg else f'lwb200 call failed ({rc})')


EXPORTED_SYMBOLS = [
    'lwb200_last_error', 'lwb200_abi_version', 'lwb200_device_count', 'lwb200_create',
    'lwb200_destroy', 'lwb200_set_stream', 'lwb200_set_lambda_range', 'lwb200_set_active_columns', 'lwb200_ng_configure', 'lwb200_ng_accelerate', 'lwb200_ng_clear', 'lwb200_last_ng', 'lwb200_upload',
    'lwb200_download', 'lwb200_sync', 'lwb200_compute_profiles', 'lwb200_fs_iter',
    'lwb200_finalise', 'lwb200_dj_max', 'lwb200_formal_sol', 'lwb200_stat_eq',
    'lwb200_device_buffer', 'lwb200_work_stats', 'lwb200_kernel_time', 'lwb200_redistribute_prd',
    'lwb200_time_dep_update', 'lwb200_formal_sol_full_stokes',
    'lwb200_nr_post_update', 'lwb200_stat_eq_async', 'lwb200_last_singular', 'lwb200_last_dj',
    'lwb200_set_zplane', 'lwb200_population_solve', 'lwb200_global_launch_count',
    'lwb200_set_collision_prefill', 'lwb200_set_j20', 'lwb200_compute_polarised_profiles', 'lwb200_set_hybrid_prd', 'lwb200_configure_hprd', 'lwb200_free_hprd',
]
