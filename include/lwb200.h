/*
 * lwb200.h -- the C-ABI of the B200 back end for Lightweaver's formal-solution /
 * Gamma-accumulation / statistical-equilibrium hot path.
 *
 * Plain C: POD structs, raw pointers and sizes, int return codes.  No C++, no
 * torch and no Lightweaver types cross this boundary.  Three things bind it:
 *   - lightweaver_b200/csrc/lwb200_plugin.cpp, the C++ shim compiled against
 *     the reference headers that exports `fs_iteration_fns_provider` /
 *     `fs_provider` (reference: Source/LwFormalInterface.hpp:35-43,110-134,
 *     Source/FormalInterface.cpp:9-28,62-81);
 *   - lightweaver_b200/capi.py (ctypes), the Python host side;
 *   - the oracle (oracle/lw_oracle.c) and the reference harness
 *     (oracle/ref_harness.cpp), which consume the same LwB200Problem so that
 *     all implementations see identical inputs.
 *
 * Array conventions follow the reference (Source/CmoArray.hpp): row-major,
 * C-contiguous fp64, depth index k innermost, k = 0 at the TOP of the
 * atmosphere.  Every per-column array carries a leading [Ncol] dimension; a
 * classic 1D Lightweaver Context is Ncol == 1, a 1.5D stack is Ncol columns
 * that share the atomic models / wavelength grids / quadrature and differ in
 * their atmospheres, populations and line profiles.
 *
 * Units as in the reference: wavelengths nm, heights m, populations m^-3.
 */
#ifndef LWB200_H
#define LWB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LWB200_ABI_VERSION 6

/* TransitionType, Source/LwTransition.hpp:10-14 */
enum { LWB200_LINE = 0, LWB200_CONTINUUM = 1 };
/* RadiationBc, Source/LwAtmosphere.hpp:6-13 */
enum {
    LWB200_BC_UNINITIALISED = 0,
    LWB200_BC_ZERO = 1,
    LWB200_BC_THERMALISED = 2,
    LWB200_BC_PERIODIC = 3,
    LWB200_BC_CALLABLE = 4
};
/* FormalSolverManager order, Source/FormalInterface.cpp:30-37 */
enum { LWB200_FS_LINEAR = 0, LWB200_FS_BESSER = 1, LWB200_FS_BEZIER3 = 2 };

/* One radiative transition of an atom (Source/LwTransition.hpp:22-91). */
typedef struct LwB200Transition {
    int32_t type;          /* LWB200_LINE / LWB200_CONTINUUM */
    int32_t i, j;          /* lower / upper level */
    int32_t Nblue, Nred;   /* active on the global grid for la in [Nblue, Nred) */
    int32_t polarised;     /* != 0: a polarised line whose six extra profiles are only ever made on the device
                              (lwb200_compute_polarised_profiles), so polProfiles may be NULL */
    double Aji, Bji, Bij;  /* lines */
    double lambda0;
    double dopplerWidth;   /* c/lambda0 for lines, 1 for continua (LwMiddleLayer.pyx:1799,1815) */
    const double* wavelength; /* [Nlambda = Nred - Nblue] */
    const double* alpha;      /* [Nlambda] continua, else NULL */
    double* phi;              /* [Ncol][Nlambda][Nrays][2][Nspace] lines, else NULL; may also be NULL for a
                                 line whose profile is only ever made on the device (lwb200_compute_profiles) */
    double* wphi;             /* [Ncol][Nspace] lines (NULL together with phi) */
    double* rhoPrd;           /* [Ncol][Nlambda][Nspace] angle-averaged PRD lines (in; out of
                                 lwb200_redistribute_prd), else NULL */
    const double* aDamp;      /* [Ncol][Nspace] lines; only read by *_compute_profiles */
    double* Rij;              /* [Ncol][Nspace] out */
    double* Rji;              /* [Ncol][Nspace] out */
    const double* Qelast;     /* [Ncol][Nspace] elastic collision rate of a PRD line (Transition::Qelast,
                                 LwTransition.hpp:51); only read by lwb200_redistribute_prd */
    const double* polProfiles; /* polarised lines (Transition::polarised, LwTransition.hpp:44-50), else NULL:
                                 [6][Ncol][Nlambda][Nrays][2][Nspace] = phiQ, phiU, phiV, psiQ, psiU, psiV;
                                 only read by lwb200_formal_sol_full_stokes */
} LwB200Transition;

/* One atom (Source/LwAtom.hpp:42-80).  detailedStatic atoms contribute
 * opacity/emissivity and get rates but no Gamma (ctx.detailedAtoms). */
typedef struct LwB200Atom {
    int32_t Nlevel;          /* 1 .. 64 */
    int32_t Ntrans;
    int32_t detailedStatic;
    int32_t reserved;
    LwB200Transition* trans; /* [Ntrans] */
    double* n;               /* [Ncol][Nlevel][Nspace] in/out */
    const double* nStar;     /* [Ncol][Nlevel][Nspace] */
    const double* nTotal;    /* [Ncol][Nspace] */
    const double* vBroad;    /* [Ncol][Nspace]; only read by *_compute_profiles */
    double* Gamma;           /* [Ncol][Nlevel][Nlevel][Nspace] in (crsw*C prefill) / out; NULL if detailedStatic */
    const double* C;         /* [Ncol][Nlevel][Nlevel][Nspace] collisional rates (Atom::C); only read by
                                lwb200_redistribute_prd and lwb200_nr_post_update, may be NULL otherwise */
    const double* stages;    /* [Nlevel] ionisation stage of every level (Atom::stages); only read by
                                lwb200_nr_post_update, may be NULL otherwise */
} LwB200Atom;

/* Hybrid PRD (Leenaarts et al. 2012): what configure_hprd_coeffs (Source/Prd.cpp:697-946) leaves in
 * Spectrum::{prdActive, la_to_prdLa, hPrdIdxs, JCoeffs, JRest} (Source/LwMisc.hpp:94-104) and in
 * Transition::hPrdCoeffs (Source/LwTransition.hpp:65) of every PRD line, flattened.  With it the emission
 * profile ratio of a PRD line is interpolated per ray to the rest-frame wavelength
 * (Transition::uv, LwTransition.hpp:115-130), the formal solution scatters w_mu/2 * frac * I into the
 * rest-frame mean intensity JRest (SimdFullIterationTemplates.hpp:397-408), and the redistribution reads
 * JRest instead of J (Prd.cpp:384-389, :484-489). */
typedef struct LwB200HybridPrd {
    int32_t NprdLa;            /* wavelengths at which a PRD line is active: rows of JRest */
    int32_t NhPrd;             /* wavelengths that scatter into them (spect.hPrdIdxs.size()); the set depends on the
                                  velocity field, so on the column: the largest count of any column (table stride) */
    int32_t Nlines;            /* PRD lines carrying interpolation coefficients */
    int32_t reserved;
    const int32_t* prdLaOfLa;  /* [Nspect] spect.la_to_prdLa where spect.prdActive, else -1 */
    const int32_t* hPrdLaOfLa; /* [Ncol][Nspect] spect.la_to_hPrdLa where spect.hPrdActive, else -1 */
    double* JRest;             /* [Ncol][NprdLa][Nspace] out of the formal solutions, in of the redistribution */
    const int64_t* JCoeffOff;  /* [Ncol * NhPrd * Nrays * 2 * Nspace + 1] offsets of JCoeffs(hPrdLa, mu, toObs, k) at
                                  ((((col * NhPrd + hPrdLa) * Nrays + mu) * 2 + toObs) * Nspace + k) */
    const int32_t* JCoeffIdx;  /* [JCoeffOff[last]] JInterpCoeffs::idx (row of JRest) */
    const double* JCoeffFrac;  /* [JCoeffOff[last]] JInterpCoeffs::frac */
    const int32_t* lineAtom;   /* [Nlines] index into problem->atoms */
    const int32_t* lineTrans;  /* [Nlines] index into that atom's trans */
    const int64_t* rhoCoefOff; /* [Nlines] element offset of a line's coefficients in rhoFrac / rhoI0 */
    const double* rhoFrac;     /* per line [Ncol][Nlambda][Nrays][2][Nspace] RhoInterpCoeffs::frac */
    const int32_t* rhoI0;      /* same shape: RhoInterpCoeffs::i0 (i1 is always i0 + 1) */
} LwB200HybridPrd;

/* What the hot path reads from / writes to a Context (Source/LwContext.hpp:20-45). */
typedef struct LwB200Problem {
    int32_t abiVersion;    /* LWB200_ABI_VERSION */
    int32_t Ncol;
    int32_t Nspace;        /* 3 .. 4096 (<= 128: one warp per column; <= 1024: up to 8 warps; beyond: the general
                              kernel with up to 32 warps) */
    int32_t Nrays;
    int32_t Nspect;
    int32_t Natom;         /* active + detailed-static atoms */
    int32_t formalSolver;  /* LWB200_FS_* */
    int32_t lowerBc;       /* zLowerBc.type */
    int32_t upperBc;       /* zUpperBc.type */
    int32_t NlowerBcMu;    /* bcData second dim when CALLABLE */
    int32_t NupperBcMu;
    int32_t reserved;
    const double* height;      /* [Ncol][Nspace] */
    const double* temperature; /* [Ncol][Nspace] */
    const double* vlosMu;      /* [Ncol][Nrays][Nspace]; only read by *_compute_profiles */
    const double* muz;         /* [Nrays] */
    const double* wmu;         /* [Nrays] */
    const double* wavelength;  /* [Nspect] */
    const double* chiBg;       /* [Ncol][Nspect][Nspace] background.chi */
    const double* etaBg;       /* [Ncol][Nspect][Nspace] background.eta */
    const double* scaBg;       /* [Ncol][Nspect][Nspace] background.sca */
    const double* lowerBcData; /* [Ncol][Nspect][NlowerBcMu] when CALLABLE (bcData(la, muIdx, 0)) */
    const double* upperBcData; /* [Ncol][Nspect][NupperBcMu] */
    const int32_t* lowerBcIdx; /* [Nrays][2] zLowerBc.idxs(mu, toObs) */
    const int32_t* upperBcIdx; /* [Nrays][2] */
    double* J;                 /* [Ncol][Nspect][Nspace] in (J-dagger) / out */
    double* I;                 /* [Ncol][Nspect][Nrays] out: spect.I(la, mu, 0) */
    double* depthChi;          /* [Ncol][Nspect][Nrays][2][Nspace] or NULL (DepthData) */
    double* depthEta;
    double* depthI;
    LwB200Atom* atoms;         /* [Natom] */
    double* Quv;               /* [Ncol][3][Nspect][Nrays] out of lwb200_formal_sol_full_stokes: spect.Quv(s, la, mu, 0)
                                  (LwMisc.hpp:94), or NULL */
    double* ne;                /* [Ncol][Nspace] electron density (atmos.ne): in/out of lwb200_nr_post_update,
                                  or NULL */
    const LwB200HybridPrd* hprd; /* hybrid PRD tables, or NULL: every PRD line is angle-averaged */
} LwB200Problem;

/* Input/output groups for lwb200_upload / lwb200_download. */
enum {
    LWB200_ATMOS   = 1u << 0, /* height, temperature, vlosMu, BC data */
    LWB200_BACKGR  = 1u << 1, /* chiBg, etaBg, scaBg */
    LWB200_POPS    = 1u << 2, /* n */
    LWB200_NSTAR   = 1u << 3, /* nStar, nTotal, vBroad */
    LWB200_GAMMA   = 1u << 4, /* up: Gamma prefill (crsw*C); down: finalised Gamma */
    LWB200_JBAR    = 1u << 5, /* J */
    LWB200_PROFILE = 1u << 6, /* phi, wphi, rhoPrd, aDamp */
    LWB200_INTENS  = 1u << 7, /* down only: I */
    LWB200_RATES   = 1u << 8, /* down: Rij, Rji; up: host Rij, Rji into the device accumulator rows (only
                                 lwb200_redistribute_prd reads rates on the device) */
    LWB200_DEPTH   = 1u << 9, /* down only: depthChi/Eta/I */
    LWB200_ADAMP   = 1u << 10, /* up only: aDamp alone (profiles are then made by lwb200_compute_profiles) */
    LWB200_GAMMA_FINAL = 1u << 11, /* up only: host Gamma taken as the finalised matrix stat_eq reads */
    LWB200_PRD     = 1u << 12, /* up: rhoPrd, Qelast, C, aDamp (inputs of lwb200_redistribute_prd);
                                  down: rhoPrd */
    LWB200_STOKES  = 1u << 13, /* up: the polarised profiles of every polarised line; down: Quv */
    LWB200_OWN_ROWS = 1u << 14, /* down, with JBAR / INTENS: only the rows of the context's wavelength range
                                   (lwb200_set_lambda_range) -- what a lambda-shard owns */
    LWB200_ZPLANE  = 1u << 15, /* down only: the ZPlaneUp / ZPlaneDown arrays registered with lwb200_set_zplane */
    LWB200_COLLISIONS = 1u << 16, /* up only: C of every active atom, kept on the device so that the prefill
                                   * crsw*C is made there (lwb200_set_collision_prefill) */
    LWB200_POLPROF = 1u << 17,    /* down only: the six polarised profiles of every polarised line with a host array */
    LWB200_ALL_INPUTS  = 0x7fu,
    LWB200_ITER_INPUTS = LWB200_POPS | LWB200_NSTAR | LWB200_GAMMA,
    LWB200_ITER_OUTPUTS = LWB200_GAMMA | LWB200_JBAR | LWB200_INTENS | LWB200_RATES
};

/* Flags for lwb200_fs_iter. */
enum {
    LWB200_LAMBDA_ITERATE = 1u << 0, /* FsMode::PureLambdaIteration */
    LWB200_STORE_DEPTH    = 1u << 1, /* depthData.fill */
    LWB200_DEFER_FINALISE = 1u << 2, /* leave [Gamma|R] partial sums un-finalised (lambda-sharded
                                        ranks all-reduce them, then call lwb200_finalise) */
    LWB200_GENERAL_KERNEL = 1u << 3, /* run every wavelength through the general per-ray accumulation
                                        kernel (normally only wavelengths with > 3 overlapping lines);
                                        a cross-check of the moment pipeline, not a fast path */
    LWB200_DJ_ASYNC       = 1u << 5, /* reduce dJ and send (max, index) to pinned host memory with the stream,
                                        without synchronising; read it with lwb200_last_dj after lwb200_sync */
    LWB200_FETCH_EARLY    = 1u << 4  /* start copying J and I back to the host buffers as soon as the rays
                                        are done, overlapped with the Gamma accumulation; a following
                                        lwb200_download(JBAR | INTENS) then only waits for that copy */
};

/* Device buffers a caller may need to hand to a collective. */
enum {
    LWB200_BUF_ACCUM = 0,  /* packed fp64 [Ncol][ sum_a Nlevel^2*Nspace | 2*sum_t Nspace ] partial Gamma|R */
    LWB200_BUF_J     = 1,  /* [Ncol][Nspect][Nspace] */
    LWB200_BUF_I     = 2,  /* [Ncol][Nspect][Nrays] */
    LWB200_BUF_POPS  = 3,  /* [Ncol][sum_a Nlevel][Nspace] */
    LWB200_BUF_GAMMA = 4,  /* [Ncol][sum_a Nlevel^2][Nspace] finalised */
    LWB200_BUF_DJ    = 5,  /* [Ncol][Nspect] per-wavelength max_k |1 - Jdag/J| */
    LWB200_BUF_DJMAX = 6   /* [1] the reduced dJ of the last lwb200_fs_iter / lwb200_dj_max over this context's
                              wavelength range (what a lambda-sharded run max-reduces across ranks) */
};

typedef struct LwB200Context LwB200Context; /* opaque */

/* Every call returns 0 on success, non-zero on failure; lwb200_last_error()
 * (thread-local) then describes it.  The shim turns failures into
 * std::runtime_error, as the reference expects (LwMiddleLayer.pyx:336-350). */
const char* lwb200_last_error(void);
int lwb200_abi_version(void);
int lwb200_device_count(int* count);
/* Kernels this library has launched in this process so far, over all contexts: lets a host (or a test
 * of the plugin shim, which owns its contexts privately) prove that a call ran on the device. */
int64_t lwb200_global_launch_count(void);

/* Replaces Context::initialise_threads / ThreadData::initialise
 * (Source/ThreadStorage.cpp:480-536): takes the problem description, keeps the
 * host pointers (never frees them), builds the wavelength work plan and
 * allocates the device mirrors.  Nothing is uploaded yet (phi is still zero
 * when the reference fires alloc_global_scratch, LwMiddleLayer.pyx:2973-2975). */
int lwb200_create(const LwB200Problem* problem, int device, LwB200Context** out);
int lwb200_destroy(LwB200Context* ctx);

/* Launch everything on this cudaStream_t (default: the legacy default stream). */
int lwb200_set_stream(LwB200Context* ctx, void* cudaStream);

/* Wavelength partition for 1D lambda-sharding (replaces the enkiTS task set,
 * Source/SimdFullIterationTemplates.hpp:675-698): this context only sweeps
 * la in [laStart, laEnd).  Default [0, Nspect). */
int lwb200_set_lambda_range(LwB200Context* ctx, int32_t laStart, int32_t laEnd);

/* Column mask of a 1.5D stack: columns converge one by one, and a retired column (active[col] == 0)
 * is skipped by every kernel of lwb200_fs_iter / lwb200_formal_sol / lwb200_stat_eq /
 * lwb200_time_dep_update from then on -- its J, I, Gamma, rates and populations stay as they are,
 * and dJ is taken over the active columns only.  The launch grids shrink with the mask, so the tail
 * of a convergence run costs what its remaining columns cost.  (The reference iterates one Context
 * per column and simply stops calling the converged ones: iterate_ctx.py:85-88.)
 * active: [Ncol] bytes, or NULL for "all columns" (the default).  PRD, full-Stokes and
 * Newton-Raphson calls are refused while a mask is set. */
int lwb200_set_active_columns(LwB200Context* ctx, const uint8_t* active);

/* The ZPlaneDecomposition extra parameters of intensity_core_opt
 * (Source/SimdFullIterationTemplates.hpp:254-281, :351-360): when registered, every formal solution
 * (lwb200_fs_iter, lwb200_formal_sol) also records I(1) of each up-going ray into zPlaneUp(la, mu) and
 * I(Nz - 2) of each down-going ray into zPlaneDown(la, mu).  Host arrays [Ncol][Nspect][Nrays], either
 * may be NULL; both NULL switches the recording off again.  They come home with
 * lwb200_download(LWB200_ZPLANE). */
int lwb200_set_zplane(LwB200Context* ctx, double* zPlaneUp, double* zPlaneDown);

/* Host -> device / device -> host copies of the groups in `mask`, using the
 * host pointers registered at create time.  Asynchronous on the context's
 * stream; lwb200_sync waits. */
int lwb200_upload(LwB200Context* ctx, uint32_t mask);
int lwb200_download(LwB200Context* ctx, uint32_t mask);
int lwb200_sync(LwB200Context* ctx);

/* Replaces Transition::compute_phi + compute_wphi (Source/FormalScalar.cpp:28-134)
 * on the device, from aDamp, vBroad, vlosMu already uploaded. */
int lwb200_compute_profiles(LwB200Context* ctx);

/* Replaces formal_sol_gamma_matrices / formal_sol_iteration_matrices_impl
 * (Source/FormalScalar.cpp:678-681, Source/SimdFullIterationTemplates.hpp:588-719)
 * on device-resident data: zeroes rates, sweeps every wavelength, accumulates
 * J, Gamma (onto the uploaded prefill) and Rij/Rji, finalises Gamma.
 * dJMax / dJMaxIdx may be NULL (no device->host sync is then forced). */
int lwb200_fs_iter(LwB200Context* ctx, uint32_t flags, double* dJMax, int64_t* dJMaxIdx);
/* The Zeeman components of one polarised line (ZeemanComponents, Source/LwMisc.hpp:106-111). */
typedef struct LwB200Zeeman {
    int32_t atom;          /* index into problem->atoms */
    int32_t trans;         /* index into that atom's trans */
    int32_t Ncomponent;
    int32_t reserved;
    const int32_t* alpha;  /* [Ncomponent] -1 (sigma blue), 0 (pi), +1 (sigma red) */
    const double* shift;   /* [Ncomponent] in Larmor units */
    const double* strength;/* [Ncomponent] */
} LwB200Zeeman;

/* Replaces Transition::compute_polarised_profiles (Source/FormalStokes.cpp:9-117) on the device for the listed
 * lines: phi, wphi and phiQ, phiU, phiV, psiQ, psiU, psiV from aDamp, vBroad and vlosMu already uploaded, the
 * magnetic field strength B [Ncol][Nspace] (Tesla) and the projections cosGamma, cos2chi, sin2chi
 * [Ncol][Nrays][Nspace] of Atmosphere::update_projections.  The Voigt and Faraday-Voigt functions are the real and
 * imaginary parts of w(v + i a), evaluated as in lwb200_compute_profiles.  The lines must have been declared
 * polarised at lwb200_create (polProfiles or the `polarised` flag).  Download group LWB200_POLPROF brings the
 * six profiles home into polProfiles where that is not NULL. */
int lwb200_compute_polarised_profiles(LwB200Context* ctx, const LwB200Zeeman* lines, int32_t nLines, const double* B,
                                      const double* cosGamma, const double* cos2chi, const double* sin2chi);

/* The "J20" extra parameter of formal_sol_full_stokes (Source/FormalStokes.cpp:676-681): J20 is the caller's
 * [Ncol][Nspect][Nspace] radiation-field anisotropy, or NULL to switch the option off again.  While it is set,
 * lwb200_formal_sol_full_stokes sends every wavelength through the Stokes solver, adds the scattering of the
 * anisotropy it finds in the array to the I and Q emissivities of a J-updating pass (a pass that does not
 * update J sees none, as in the reference, :433-437) and, with updateJ, leaves the new anisotropy
 * sum_rays w_mu (3 mu^2 - 1) / (2 sqrt 2) I + w_mu 3 (mu^2 - 1) / (2 sqrt 2) Q in it (after lwb200_sync). */
int lwb200_set_j20(LwB200Context* ctx, double* J20);

/* Replaces configure_hprd_coeffs (Source/Prd.cpp:697-946) for every column of the problem: from the wavelength
 * grid, the ranges of the PRD lines (lines with rhoPrd; those of detailed-static atoms on request) and
 * vlosMu, fills *out with the tables of LwB200HybridPrd (JRest zeroed).  Host code, no device needed.  The
 * arrays belong to the library until lwb200_free_hprd(out).  out->Nlines == 0: the problem has no PRD line. */
int lwb200_configure_hprd(const LwB200Problem* problem, int includeDetailed, LwB200HybridPrd* out);
void lwb200_free_hprd(LwB200HybridPrd* tables);

/* Replace the hybrid-PRD tables of a context created with LwB200Problem::hprd (configure_hprd_coeffs run
 * again after the velocity field changed, Source/Prd.cpp:697-946).  The tables must name the same PRD lines;
 * the wavelengths that scatter into the PRD grid must stay within two grid points of the set the context was
 * planned for (otherwise: error, create a new context).  JRest travels with LWB200_PRD. */
int lwb200_set_hybrid_prd(LwB200Context* ctx, const LwB200HybridPrd* tables);

/* The prologue of lw.Context.formal_sol_gamma_matrices, Gamma = crsw * C (Source/LwMiddleLayer.pyx:3198-3203),
 * on the device: with enable != 0 the finalisation takes crsw times the collisional rates uploaded with
 * LWB200_COLLISIONS instead of a prefill uploaded with LWB200_GAMMA, so a caller whose C is unchanged
 * (fixCollisionalRates) neither forms the product on the host nor sends it.  Same rounding as the host's.
 * A later upload of LWB200_GAMMA switches back to the uploaded prefill. */
int lwb200_set_collision_prefill(LwB200Context* ctx, int enable, double crsw);
/* Gamma = prefill + partial sums, diagonal = -column sum (finalise_Gamma, :491-508);
 * rates are unpacked.  Only needed after LWB200_DEFER_FINALISE. */
int lwb200_finalise(LwB200Context* ctx);
/* Reduce LWB200_BUF_DJ to (max, wavelength index of the max). */
int lwb200_dj_max(LwB200Context* ctx, double* dJMax, int64_t* dJMaxIdx);

/* Replaces formal_sol / formal_sol_impl (Source/FormalScalar.cpp:691-694,
 * SimdFullIterationTemplates.hpp:721-781): I only, no J / Gamma / rates. */
int lwb200_formal_sol(LwB200Context* ctx, int upOnly);

/* Replaces stat_eq_impl + solve_lin_eq (Source/UpdatePopulations.cpp:7-47,
 * Source/LuSolve.cpp:8-133) for atom `atom` (index into problem->atoms, or -1
 * for every active atom), depths [kStart, kEnd) (both < 0: all).  *nSingular
 * receives the number of (column, depth) systems with an all-zero row; the
 * call then fails like the reference's throw ("Singular Matrix"). */
int lwb200_stat_eq(LwB200Context* ctx, int32_t atom, int32_t kStart, int32_t kEnd, int32_t* nSingular);

/* Replaces formal_sol_full_stokes_impl (Source/FormalStokes.cpp:664-723; FsIterationFns::full_stokes_fs,
 * LwFormalInterface.hpp:117): polarised formal solution of every wavelength -- DELO-Bezier3
 * (piecewise_stokes_bezier3_1d_impl, :166-340) where a polarised line is active, the scalar Bezier3
 * solver elsewhere (as the reference, whatever formalSolver says).  Writes I and Quv, and J / dJ when
 * updateJ; no Gamma, no rates.  As in the reference, a pass that does not update J has no scattering
 * term in its source function (stokes_fs_core fills JDag only under updateJ, FormalStokes.cpp:431-441).
 * Quv at wavelengths without a polarised line is 0 (the reference leaves whatever the last polarised
 * ray left in its scratch there). */
int lwb200_formal_sol_full_stokes(LwB200Context* ctx, int updateJ, int upOnly, double* dJMax, int64_t* dJMaxIdx);

/* Ng acceleration of the populations on the device (Source/Ng.hpp:16-163; the reference runs one Ng
 * object per atom on the host after every population update, LwMiddleLayer.pyx:3318-3346).
 * lwb200_ng_configure = the Ng(Norder, Nperiod, Ndelay, n) constructor on the populations currently
 * on the device, for every active atom (and column); lwb200_ng_accelerate = accelerate(n) followed
 * by max_change(): *accelerated tells whether this call extrapolated, dMax / dMaxIdx [Natom] are the
 * largest relative change between the last two stored solutions of each atom and its flat index
 * (level * Nspace + depth, + column * Nlevel * Nspace in a stack; 0 for detailed-static atoms).
 * Norder <= 4.  Norder = 0 tracks changes only, like the reference's default Ng(0, 0, 0).
 * Configurations with max(Ndelay, Nperiod + 2) < Norder + 2 are refused: the reference indexes
 * its history with a negative row there. */
int lwb200_ng_configure(LwB200Context* ctx, int32_t Norder, int32_t Nperiod, int32_t Ndelay);
int lwb200_ng_accelerate(LwB200Context* ctx, int32_t* accelerated, double* dMax, int64_t* dMaxIdx);
int lwb200_ng_clear(LwB200Context* ctx);
/* lwb200_ng_accelerate(ctx, &accelerated, NULL, NULL) does not synchronise the host: dMax / dMaxIdx are
 * read with lwb200_last_ng after the next lwb200_sync, a singular acceleration system is reported by
 * lwb200_last_singular. */
int lwb200_last_ng(LwB200Context* ctx, double* dMax, int64_t* dMaxIdx);

/* Latency-hiding variants for a host that synchronises once per call sequence (the Python mirror):
 * lwb200_stat_eq_async launches the solve and sends the singular-system count home with the stream;
 * lwb200_last_singular / lwb200_last_dj read those results after the next lwb200_sync
 * (lwb200_last_singular fails with "Singular Matrix" like lwb200_stat_eq; the count is cumulative over
 * the asynchronous updates launched since it was last collected). */
int lwb200_stat_eq_async(LwB200Context* ctx, int32_t atom, int32_t kStart, int32_t kEnd);
int lwb200_last_singular(LwB200Context* ctx, int32_t* nSingular);
int lwb200_last_dj(LwB200Context* ctx, double* dJMax, int64_t* dJMaxIdx);

/* Replaces time_dependent_update_impl (Source/UpdatePopulations.cpp:120-151;
 * FsIterationFns::time_dep_update, LwFormalInterface.hpp:120): backward-Euler population update
 * (1 - Gamma dt) n = nOld of atom `atom`, per depth, through the same LU solve as lwb200_stat_eq.
 * nOld: host [Ncol][Nlevel][Nspace].  Gamma is the finalised matrix on the device (the last
 * lwb200_fs_iter, or an LWB200_GAMMA_FINAL upload). */
int lwb200_time_dep_update(LwB200Context* ctx, int32_t atom, const double* nOld, double dt, int32_t kStart,
                           int32_t kEnd, int32_t* nSingular);

/* stat_eq_impl / time_dependent_update_impl (Source/UpdatePopulations.cpp:7-47, :120-151) for an atom
 * that belongs to no device context: FsIterationFns::stat_eq and ::time_dep_update receive only an
 * Atom*, and a caller may update populations before any formal solution has run on its Context (e.g.
 * after the escape-probability initial solution).  Self-contained: host arrays in, host populations
 * out, synchronous.  Gamma [Ncol][Nlevel][Nlevel][Nspace] finalised, n [Ncol][Nlevel][Nspace] in/out,
 * nTotal [Ncol][Nspace]; nOld [Ncol][Nlevel][Nspace] or NULL (statistical equilibrium), dt used with
 * nOld.  Depths [kStart, kEnd) (both < 0: all). */
int lwb200_population_solve(int device, int32_t Ncol, int32_t Nlevel, int32_t Nspace, const double* Gamma,
                            double* n, const double* nTotal, const double* nOld, double dt, int32_t kStart,
                            int32_t kEnd, int32_t* nSingular);

/* Inputs of lwb200_nr_post_update that are not part of the problem. */
typedef struct LwB200NrUpdate {
    int32_t Natom;              /* atoms taking part in the charge-conservation system */
    int32_t timeDependent;      /* nPrev / dt are valid (NrTimeDependentData, Lightweaver.hpp:12-16) */
    const int32_t* atomIdx;     /* [Natom] indices into problem->atoms (active atoms) */
    const double* const* dC;    /* NULL, or [Natom] pointers to [Ncol][Nlevel][Nlevel][Nspace] finite-difference
                                   dC/dne */
    const double* backgroundNe; /* [Ncol][Nspace] */
    const double* const* nPrev; /* [Natom] pointers to [Ncol][Nlevel][Nspace] (timeDependent) */
    double dt;
    double crswVal;
} LwB200NrUpdate;

/* Replaces nr_post_update_impl (Source/UpdatePopulations.cpp:230-394; FsIterationFns::nr_post_update,
 * LwFormalInterface.hpp:121-125): one Newton-Raphson step of the coupled statistical-equilibrium (or
 * backward-Euler) + charge-conservation system of the listed atoms, per depth: (sum Nlevel + 1)^2
 * unknowns through the same LU solve.  Updates the atoms' populations and problem->ne on the device
 * and on the host (ne is copied back by the call; populations with LWB200_POPS).  Gamma is the
 * finalised matrix on the device. */
int lwb200_nr_post_update(LwB200Context* ctx, const LwB200NrUpdate* upd, int32_t kStart, int32_t kEnd,
                          int32_t* nSingular);

/* Replaces redistribute_prd_lines (Source/Prd.cpp:648-658, PrdTemplates.hpp:164-351;
 * FsIterationFns::redistribute_prd, LwFormalInterface.hpp:118) for angle-averaged PRD lines:
 * up to maxIter sub-iterations of
 *   (1) per PRD line and depth the scattering integral of J against Gouttebroze's GII on the
 *       fixed-step fine grid, rho = 1 + gamma (scatInt / gNorm - Jbar)   (Prd.cpp:9-124, :468-575),
 *   (2) a formal solution over the wavelengths touched by a PRD line that updates J, I and the
 *       rates of the PRD lines only (formal_sol_prd_update_rates, PrdTemplates.hpp:18-155),
 * until the largest relative change of rho falls below tol.  Works on the device-resident state
 * (populations, J, rates of the last lwb200_fs_iter); inputs that only this call reads are
 * uploaded with LWB200_PRD.  includeDetailed: also redistribute the PRD lines of detailed-static
 * atoms (extraParams["include_detailed_atoms"]).
 * Outputs (any may be NULL): *nIter sub-iterations taken; dRho / dRhoIdx [maxIter * NprdLines]
 * in (iteration, line) order; dJPrdMax / dJPrdMaxIdx [maxIter].  Hybrid PRD (hPrdCoeffs, JRest)
 * is not handled here: the plugin shim leaves such Contexts to the reference's own function. */
int lwb200_redistribute_prd(LwB200Context* ctx, int32_t maxIter, double tol, int32_t includeDetailed,
                            int32_t* nIter, double* dRho, int32_t* dRhoIdx, double* dJPrdMax,
                            int64_t* dJPrdMaxIdx);

/* Device time (ms, CUDA events on the context's stream) of the most recent
 * formal-solution kernel launched by lwb200_fs_iter / lwb200_formal_sol: the
 * dominant kernel of the path, for the roofline. */
int lwb200_kernel_time(LwB200Context* ctx, double* ms);

/* Device pointer + byte size of one of the LWB200_BUF_* buffers. */
int lwb200_device_buffer(LwB200Context* ctx, int32_t which, void** ptr, size_t* nbytes);

/* Work accounting for the roofline: ray-depth points and algorithmic bytes of
 * one fs_iter over this context's wavelength range, and how many
 * kernels the last call launched. */
int lwb200_work_stats(LwB200Context* ctx, double* points, double* algBytes, int64_t* lastLaunches);

#ifdef __cplusplus
}
#endif
#endif /* LWB200_H */
