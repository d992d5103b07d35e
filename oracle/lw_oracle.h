/* lw_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's formal-solution / Gamma-iteration /
 * stat-eq hot path, operating on the same LwB200Problem the CUDA library takes.
 * Parity is PINNED: tests/test_oracle_vs_ref.py checks it against the
 * reference's own compiled C++ (oracle/_ref, scalar scheme) and
 * tests/golden/*.npz holds outputs of that reference for use where
 * /root/reference does not exist (the GPU box).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or load this.
 */
#ifndef LW_ORACLE_H
#define LW_ORACLE_H
#include "../include/lwb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One ray through one of the three 1D solvers (FormalScalar.cpp:136-666).
 * Psi may be NULL.  lowerBc/upperBc: LWB200_BC_ZERO or _THERMALISED, or
 * _CALLABLE with the boundary intensity in bcValue. */
void lwo_solve_ray(int solver, int Nspace, const double* height, const double* temperature,
                   const double* chi, const double* S, double muz, int toObs, double wavelength,
                   int lowerBc, int upperBc, double bcValue, double* I, double* Psi);

/* formal_sol_gamma_matrices on column `col` (SimdFullIterationTemplates.hpp:588-637,
 * Nthreads <= 1 branch).  flags: LWB200_LAMBDA_ITERATE | LWB200_STORE_DEPTH.
 * dJMaxIdx is the argmax wavelength (what the reference's threaded branch
 * returns, :688,:703); dJMaxIdxSerial reproduces the single-threaded branch's
 * index (:627, argument order of max_idx). */
int lwo_fs_iter(const LwB200Problem* p, int col, unsigned flags, int laStart, int laEnd,
                double* dJMax, int64_t* dJMaxIdx, int64_t* dJMaxIdxSerial);

/* formal_sol (SimdFullIterationTemplates.hpp:721-737). */
int lwo_formal_sol(const LwB200Problem* p, int col, int upOnly);

/* stat_eq_impl (UpdatePopulations.cpp:7-47); atom < 0: all active atoms.
 * Returns 0, or 1 with *nSingular > 0 when lu_decompose would have thrown. */
int lwo_stat_eq(const LwB200Problem* p, int col, int atom, int kStart, int kEnd, int* nSingular);

/* solve_lin_eq (LuSolve.cpp:103-133); returns 1 for "Singular Matrix". */
int lwo_solve_lin_eq(int N, double* A, double* b, int improve);

/* Gamma iteration (+ optional stat_eq) over columns [col0, col0+ncol) on
 * nthreads OpenMP threads -- the "port" CPU baseline for column stacks. */
int lwo_fs_iter_columns(const LwB200Problem* p, int col0, int ncol, unsigned flags,
                        int withStatEq, int nthreads);

/* formal_sol_full_stokes_impl (FormalStokes.cpp:664-723) on column `col`: DELO-Bezier3 at wavelengths
 * with a polarised line, scalar Bezier3 elsewhere; writes I, Quv (incl. the reference's stale values
 * at unpolarised wavelengths: whatever the last polarised ray left in I(1..3, 0)) and, with updateJ,
 * J and dJ (argmax index). */
int lwo_full_stokes(const LwB200Problem* p, int col, int updateJ, int upOnly, double* dJMax, int64_t* dJMaxIdx);
/* The same with the "J20" extra parameter (FormalStokes.cpp:433-437, :469-471, :485-490, :575-583, :642-648):
 * J20 [Ncol][Nspect][Nspace] is the radiation-field anisotropy of the last J-updating pass on entry (its
 * scattering term enters the I and Q emissivities, every wavelength goes through the Stokes solver) and,
 * with updateJ, the new one on return.  NULL: lwo_full_stokes. */
int lwo_full_stokes_j20(const LwB200Problem* p, int col, int updateJ, int upOnly, double* J20, double* dJMax,
                        int64_t* dJMaxIdx);

/* time_dependent_update_impl (UpdatePopulations.cpp:120-151) of atom `atom` on column `col`;
 * nOld is [Ncol][Nlevel][Nspace].  Returns 1 for "Singular Matrix". */
int lwo_time_dep_update(const LwB200Problem* p, int col, int atom, const double* nOld, double dt);

/* nr_post_update_impl (UpdatePopulations.cpp:230-394) on column `col`: one Newton-Raphson step of the
 * populations of the listed atoms and of p->ne.  Returns 1 for "Singular Matrix". */
int lwo_nr_post_update(const LwB200Problem* p, int col, const LwB200NrUpdate* upd);

/* redistribute_prd_lines for angle-averaged PRD lines (Prd.cpp:9-124, :468-658;
 * PrdTemplates.hpp:18-76, :164-291), Nthreads <= 1 branch, on column `col`.  dRho / dRhoIdx
 * [maxIter * NprdLines] in (iteration, line) order, dJPrdMax / dJPrdMaxIdx [maxIter];
 * dJPrdMaxIdx is the argmax wavelength (the reference's serial branch returns the index quirk
 * described at lwo_fs_iter).  Returns 1 if a PRD line lacks Qelast / C / aDamp / vBroad. */
int lwo_redistribute_prd(const LwB200Problem* p, int col, int maxIter, double tol, int includeDetailed,
                         int* nIter, double* dRho, int* dRhoIdx, double* dJPrdMax, int64_t* dJPrdMaxIdx);

/* configure_hprd_coeffs (Prd.cpp:697-946) for every column of the problem (each column is its own Context
 * in the reference), flattened into LwB200HybridPrd: prdActive / la_to_prdLa, hPrdActive / la_to_hPrdLa,
 * JCoeffs, a zeroed JRest and the hPrdCoeffs of every PRD line.  With problem->hprd pointing at the result,
 * lwo_fs_iter accumulates JRest (SimdFullIterationTemplates.hpp:397-408), uv interpolates rho per ray
 * (LwTransition.hpp:115-130) and lwo_redistribute_prd works in the rest frame over hPrdIdxs.
 * The arrays are malloc'ed: lwo_free_hprd.  Returns 1 without vlosMu. */
int lwo_configure_hprd(const LwB200Problem* p, int includeDetailed, LwB200HybridPrd* out);
void lwo_free_hprd(LwB200HybridPrd* h);

/* Ng acceleration (Ng.hpp:16-163): constructor on sols[0], then accelerate() + max_change() on
 * sols[1..nIter]; out [nIter][len] = the solutions as accelerate() leaves them.  Returns 1 for a
 * singular acceleration system. */
int lwo_ng_run(int Norder, int Nperiod, int Ndelay, int len, int nIter, const double* sols, double* out,
               int* accelerated, double* dMax, int64_t* dMaxIdx);

/* Transition::compute_phi / compute_wphi are NOT restated here (they need
 * Faddeeva); the tests use scipy.special.wofz (the same Faddeeva package). */

#ifdef __cplusplus
}
#endif
#endif
