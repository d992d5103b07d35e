// lwb200_kernels.cuh -- the hot-path kernels.
//
//   fs_kernel<NCH, SOLVER, MODE>   gather + source function + formal solution
//                                  (+ J, Gamma, Rij/Rji for MODE_ITER)
//                                  replaces intensity_core_opt and everything it
//                                  calls (SimdFullIterationTemplates.hpp:59-487)
//   finalise_kernel                Gamma = prefill + partial sums, diagonal
//                                  (finalise_Gamma, :491-508)
//   stat_eq_kernel                 per-depth N x N solve (UpdatePopulations.cpp:7-47,
//                                  LuSolve.cpp:8-133)
//   ratio_kernel, dj_reduce_kernel small helpers
//
// Work decomposition (replaces TaskScheduler/ThreadStorage): the host planner
// cuts the wavelength axis into tiles; one CTA owns (tile, column); its warps
// take wavelengths round-robin; per-transition Gamma/R partial sums live in
// shared memory for the whole tile and are flushed once with fp64 atomics.
#pragma once
#include "lwb200_device.cuh"

namespace lwb200
{
enum { MODE_FS = 0, MODE_ITER = 1 };

struct DevTrans
{
    int type; // 0 line, 1 continuum
    int i, j; // levels within the atom
    int atom;
    int Nblue, Nred;
    int levI, levJ;       // rows in the packed population arrays
    int accIJ, accJI;     // rows in the packed accumulator for Gamma(i,j), Gamma(j,i); -1: none
    int accRij, accRji;   // rows for Rij, Rji
    int tabOff;           // offset into wlambdaTab / alphaTab
    int lineIdx;          // index among lines, -1 for continua
    int contIdx;          // index among continua, -1 for lines
    int detailed;         // belongs to a detailed-static atom
    long long phiOff;     // element offset of phi[col = 0] in the phi pool
    long long phiColStride;
    long long rhoOff;     // -1: no rhoPrd
    long long hprdOff;    // hybrid PRD: element offset of the line's interpolation coefficients (column 0) in the
                          // hprdFrac / hprdI0 pools, [Ncol][Nlambda][M][2][K]; -1: angle-averaged or no PRD
    double Aji_Bji;       // Aji / Bji
    double Bji_Bij;       // Bji / Bij
    double Bij;
    double lambda0;
};

// One active transition at one wavelength.  Self-contained: the pipeline kernels read
// nothing else to know what to do with it (no second indirection through DevTrans).
struct DevEntry
{
    int trans;
    int slot;          // accumulator slot within the wavelength's tile
    int type, i, j;    // 0 line / 1 continuum; levels within the atom
    int levI, levJ;    // rows in the packed population arrays
    int contIdx;       // continuum index (gRatio row), -1 for lines
    int groupEnd;      // one past the last entry of the same atom at this wavelength
    int Nlevel;        // of the atom
    int detailed;      // atom is detailed-static (rates only, no Gamma)
    int atom;
    double al;         // continuum: alpha(lambda); line: 0
    double wlaF;       // continuum: wlambda / lambda * 4 pi / h;  line: wlambda * 4 pi / (h c)
    int nOffI, nOffJ;  // levI * K, levJ * K: element offsets into one column of n
    int gOff;          // contIdx * Ncol * K: element offset of the continuum's gRatio block
    int prd;           // line with rhoPrd (angle-averaged PRD)
    int accIJ, accJI;  // rows of the packed accumulator for Gamma(i,j), Gamma(j,i); -1: none
    int accRij, accRji;
};

// Per (wavelength, overlapping-line slot) descriptor built by the planner (<= 3 slots).
struct LambdaLine
{
    long long phiOff;       // element offset of phi(lt, 0, 0, 0) of column 0 in the phi pool
    long long phiColStride;
    long long rhoOff;       // element offset of rhoPrd(lt, 0) of column 0; -1: none
    long long rhoColStride;
    int trans;              // global transition index
    int levI, levJ;         // rows in the packed population arrays
    int lineIdx;
    int atom, i, j;
    int slot;               // accumulator slot of the line within the wavelength's tile
    double lambda0, Bij, Bji_Bij, Aji_Bji;
    double wlaS;            // wlambda * 4 pi / (h c)  (times wphi(k) gives wla)
    long long polOff;       // polarised line: element offset of phiQ(lt, 0, 0, 0) of column 0 in the pol pool; else -1
    long long polArr;       // stride between the six polarised profile arrays (phiQ, phiU, phiV, psiQ, psiU, psiV)
};

struct DevProblem
{
    int Ncol, K, M, L;
    int NlevTot, GammaTot, AccTot, Natom, NtransTot, Nline, Ncont;
    int lowerBc, upperBc, NlowerBcMu, NupperBcMu;
    int maxNlevel, maxSlots, KP;
    const double *height, *temperature, *muz, *wmu, *wavelength;
    const double *chiBg, *etaBg, *scaBg;
    double *J, *I;
    const double *n, *nStar, *gRatio;
    const double *phi, *wphi, *rhoPrd;
    const double *wlambdaTab, *alphaTab;
    const double *lowerBcData, *upperBcData;
    const int *lowerBcIdx, *upperBcIdx;
    double* accum;
    double* dJ;
    double *depthChi, *depthEta, *depthI;
    const DevTrans* trans;
    const DevEntry* entries;
    const int* laOff;        // [L] first entry of a wavelength's active-transition list
    const int* laCnt;        // [L] its length
    const int* laHasLine;
    const int* tileLambda;   // wavelength indices of every tile, ascending within a tile
    const int* tileLa;       // [Ntile + 1] offsets into tileLambda
    const int* tileSlotOff;  // [Ntile + 1]
    const int* tileSlotTrans;
    const int* atomNlevel;
    const int* atomLevOff;
    const int* atomGammaOff;
    const int* atomDetailed;
    // three-stage pipeline (lwb200_pipeline.cuh); chiC/etaC/mom hold one batch of columns
    double *chiC, *etaC;      // [batch][L][K]
    double* mom;              // [batch][momRows][K]
    const int* momOff;        // [L] first moment row of a wavelength
    const int* laNLines;      // [L] number of overlapping lines
    const LambdaLine* lamLine; // [L][3]
    int momRows;
    const int* phiAsym;       // bit 0 clear: phi(.., dir 0, .) == phi(.., dir 1, .) everywhere; bit 1 clear: phi does
                              // not depend on the ray at all (static atmosphere)
    // full Stokes (lwb200_stokes.cuh)
    // column mask of a stack (lwb200_set_active_columns): retired columns are skipped by every kernel
    const int* colList;             // active column indices, ascending; nullptr: all columns
    const unsigned char* colActive; // [Ncol]; nullptr: all columns
    const double* pol;        // polarised profile pool
    double* Quv;              // [Ncol][3][L][M]
    double* Jdag;             // [Ncol][L][K] copy of J taken before a J-updating Stokes pass
    // ZPlaneDecomposition (lwb200_set_zplane): [Ncol][L][M] each, nullptr: not recorded
    double *zPlaneUp, *zPlaneDown;
    // the "J20" extra parameter of the Stokes pass (lwb200_set_j20): [Ncol][L][K] each, nullptr: off
    double* J20;                // new anisotropy, accumulated by the J-updating pass
    const double* J20dag;       // the one on entry; nullptr in a pass that does not update J (the reference's
                                // J20Dag stays zero there, FormalStokes.cpp:433-437)
    int j20;
    // hybrid PRD (LwB200HybridPrd): nullptr / 0 without it
    const int* hprdLaOfLa;      // [Ncol][L] row of this column's JCoeffs tables, -1: the wavelength does not scatter
    const int* prdLaOfLa;       // [L] row of JRest, -1: no PRD line active
    double* JRest;              // [Ncol][NprdLa][K]
    const long long* JCoeffOff; // CSR over ((((col * NhPrd + hPrdLa) * M + mu) * 2 + toObs) * K + k)
    const int* JCoeffIdx;
    const double* JCoeffFrac;
    const double* hprdFrac;     // interpolation coefficients of every hybrid PRD line (DevTrans::hprdOff)
    const int* hprdI0;
    int NprdLa, NhPrd;
};

// ZPlaneDecomposition of intensity_core_opt (SimdFullIterationTemplates.hpp:351-360): I(1) of an up-going
// ray into zPlaneUp(la, mu), I(Nz - 2) of a down-going one into zPlaneDown(la, mu).  laneGlobal is the
// position of this lane along depth in units of NCH points; ray = (col * L + la) * M + mu.
template <int NCH>
__device__ __forceinline__ void store_zplane(const DevProblem& P, int laneGlobal, const double (&I)[NCH], int dir,
                                             size_t ray)
{
    double* dst = dir == 1 ? P.zPlaneUp : P.zPlaneDown;
    if (!dst)
        return;
    const int k = dir == 1 ? 1 : P.K - 2;
    if (laneGlobal == k / NCH)
    {
        double v = I[0];
#pragma unroll
        for (int j = 1; j < NCH; ++j)
            v = (k % NCH == j) ? I[j] : v;
        dst[ray] = v;
    }
}

// q-th column of the launch: the q-th active column when a mask is set
__device__ __forceinline__ int column_of(const DevProblem& P, int q)
{
    return P.colList ? __ldg(P.colList + q) : q;
}

// U, V for one transition at one (wavelength, ray, depth): Transition::uv
// (LwTransition.hpp:93-144) with gij from Atom::setup_wavelength
// (LwAtom.hpp:82-128), same operation order.
struct UV
{
    double Vij, Vji, Uji;
};

__device__ __forceinline__ UV trans_uv(const DevProblem& P, const DevTrans& t, int col, int lt,
                                       int mu, int dir, int k, double lambda, double expfac)
{
    UV r;
    if (t.type == 0)
    {
        constexpr double hc_4pi = 0.25 * kHC / kPi;
        const double hnu_4pi = hc_4pi * (t.lambda0 / lambda);
        const double p = __ldg(P.phi + t.phiOff + (long long)col * t.phiColStride
                               + ((long long)(lt * P.M + mu) * 2 + dir) * P.K + k);
        double g = t.Bji_Bij;
        if (t.rhoOff >= 0 && t.hprdOff < 0)
            g *= __ldg(P.rhoPrd + t.rhoOff + ((long long)col * (t.Nred - t.Nblue) + lt) * P.K + k);
        r.Vij = hnu_4pi * t.Bij * p;
        r.Vji = g * r.Vij;
        if (t.hprdOff >= 0)
        {
            // hybrid PRD: rho interpolated to the rest-frame wavelength of this ray (LwTransition.hpp:115-130)
            const long long Nl = t.Nred - t.Nblue;
            const long long o = t.hprdOff + ((((long long)col * Nl + lt) * P.M + mu) * 2 + dir) * P.K + k;
            const double frac = __ldg(P.hprdFrac + o);
            const int i0 = __ldg(P.hprdI0 + o);
            const double* rho = P.rhoPrd + t.rhoOff + (long long)col * Nl * P.K + k;
            r.Vji *= (1.0 - frac) * __ldg(rho + (long long)i0 * P.K) + frac * __ldg(rho + (long long)(i0 + 1) * P.K);
        }
        r.Uji = t.Aji_Bji * r.Vji;
    }
    else
    {
        constexpr double twoHc = 2.0 * kHC / (kNmToM * kNmToM * kNmToM);
        const double hcl = twoHc / (lambda * lambda * lambda);
        const double g = __ldg(P.gRatio + ((long long)t.contIdx * P.Ncol + col) * P.K + k) * expfac;
        r.Vij = __ldg(P.alphaTab + t.tabOff + lt);
        r.Vji = g * r.Vij;
        r.Uji = hcl * r.Vji;
    }
    return r;
}

// wla(kr, k) of Atom::setup_wavelength (LwAtom.hpp:100-118)
__device__ __forceinline__ double trans_wla(const DevProblem& P, const DevTrans& t, int col, int lt,
                                            int k, double lambda)
{
    constexpr double pi4_h = 4.0 * kPi / kHPlanck;
    constexpr double hc_4pi = 0.25 * kHC / kPi;
    constexpr double pi4_hc = 1.0 / hc_4pi;
    const double wlam = __ldg(P.wlambdaTab + t.tabOff + lt);
    if (t.type == 0)
        return wlam * __ldg(P.wphi + ((long long)t.lineIdx * P.Ncol + col) * P.K + k) * pi4_hc;
    return (wlam / lambda) * pi4_h;
}

__device__ __forceinline__ void smem_add(double* addr, double v) { atomicAdd(addr, v); }

// ---------------------------------------------------------------------------
// gRatio[cont][col][k] = nStar(i,k) / nStar(j,k): the ray- and wavelength-
// independent factor of gij for continua (LwAtom.hpp:112)
__global__ void ratio_kernel(const DevProblem P, double* gRatio)
{
    const size_t total = (size_t)P.Ncont * P.Ncol * P.K;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = idx % P.K;
        const int col = (idx / P.K) % P.Ncol;
        const int c = idx / ((size_t)P.K * P.Ncol);
        // find the c-th continuum
        int tIdx = -1;
        for (int t = 0; t < P.NtransTot; ++t)
            if (P.trans[t].contIdx == c)
            {
                tIdx = t;
                break;
            }
        const DevTrans& t = P.trans[tIdx];
        const double ni = P.nStar[((size_t)col * P.NlevTot + t.levI) * P.K + k];
        const double nj = P.nStar[((size_t)col * P.NlevTot + t.levJ) * P.K + k];
        gRatio[idx] = ni / nj;
    }
}

// zero_rates + fresh Gamma partial sums of the active columns of a masked stack (the whole
// buffer is a plain memset otherwise)
__global__ void zero_accum_kernel(const DevProblem P, int nActive)
{
    const size_t per = (size_t)P.AccTot * P.K;
    const size_t total = (size_t)nActive * per;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
        P.accum[(size_t)column_of(P, (int)(idx / per)) * per + idx % per] = 0.0;
}

// finalise_Gamma (:491-508): Gamma = prefill (crsw*C) + radiative partial sums,
// then diagonal = -(column sum).  One thread per (column, atom, level i, depth): column i of
// the atom's Gamma at one depth, loads issued together (nothing here aliases).
// `prefill` is the caller's crsw*C (uploaded with LWB200_GAMMA, scale = 1), or C itself kept on the
// device (LWB200_COLLISIONS) with scale = crsw: the product is rounded before the sum, as the host's is.
// Columns [colBase, colBase + ncols): a column stack finalises (and sends home) batch by batch.
__global__ void finalise_kernel(const DevProblem P, const double* __restrict__ prefill, double scale,
                                double* __restrict__ gamma, int colBase, int ncols)
{
    const double* __restrict__ accum = P.accum;
    const int maxN = P.maxNlevel;
    const size_t total = (size_t)ncols * P.Natom * maxN * P.K;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = idx % P.K;
        const int i = (idx / P.K) % maxN;
        const int atom = (idx / ((size_t)P.K * maxN)) % P.Natom;
        const int col = colBase + (int)(idx / ((size_t)P.K * maxN * P.Natom));
        const int N = P.atomNlevel[atom];
        if (P.atomDetailed[atom] || i >= N || (P.colActive && !P.colActive[col]))
            continue;
        const size_t gOff = ((size_t)col * P.GammaTot + P.atomGammaOff[atom]) * P.K + k;
        const size_t aOff = ((size_t)col * P.AccTot + P.atomGammaOff[atom]) * P.K + k;
        double diag = 0.0;
        for (int j = 0; j < N; ++j)
        {
            if (j == i)
                continue;
            // Gamma(j, i): rate from i to j
            const size_t r = (size_t)(j * N + i) * P.K;
            const double v = __dadd_rn(__dmul_rn(prefill[gOff + r], scale), accum[aOff + r]);
            gamma[gOff + r] = v;
            diag += v;
        }
        gamma[gOff + (size_t)(i * N + i) * P.K] = -diag;
    }
}

// ---------------------------------------------------------------------------
// stat_eq_impl + solve_lin_eq (UpdatePopulations.cpp:7-47; LuSolve.cpp:8-133):
// Crout LU with implicit scaled partial pivoting, Tiny = 1e-20 pivots, one
// step of iterative refinement.  One thread per (column, depth) system of one
// atom; MAXN bounds the local arrays.
template <int MAXN>
__device__ bool lu_decompose_dev(int N, double* A, int* index)
{
    constexpr double Tiny = 1e-20;
    double vv[MAXN];
    for (int i = 0; i < N; ++i)
    {
        double big = 0.0;
        for (int j = 0; j < N; ++j)
        {
            const double v = fabs(A[i * N + j]);
            big = (big < v) ? v : big;
        }
        if (big == 0.0)
            return false;
        vv[i] = 1.0 / big;
    }
    for (int j = 0; j < N; ++j)
    {
        for (int i = 0; i < j; ++i)
        {
            double sum = A[i * N + j];
            for (int k = 0; k < i; ++k)
                sum -= A[i * N + k] * A[k * N + j];
            A[i * N + j] = sum;
        }
        int iMax = 0;
        double big = 0.0;
        for (int i = j; i < N; ++i)
        {
            double sum = A[i * N + j];
            for (int k = 0; k < j; ++k)
                sum -= A[i * N + k] * A[k * N + j];
            A[i * N + j] = sum;
            const double cand = vv[i] * fabs(sum);
            if (big < cand)
            {
                big = cand;
                iMax = i;
            }
        }
        if (j != iMax)
        {
            for (int k = 0; k < N; ++k)
            {
                const double temp = A[iMax * N + k];
                A[iMax * N + k] = A[j * N + k];
                A[j * N + k] = temp;
            }
            vv[iMax] = vv[j];
        }
        index[j] = iMax;
        if (A[j * N + j] == 0.0)
            A[j * N + j] = Tiny;
        const double temp = 1.0 / A[j * N + j];
        for (int i = j + 1; i < N; ++i)
            A[i * N + j] *= temp;
    }
    return true;
}

__device__ inline void lu_backsub_dev(int N, const double* A, const int* index, double* b)
{
    int ii = -1;
    for (int i = 0; i < N; ++i)
    {
        const int ip = index[i];
        double sum = b[ip];
        b[ip] = b[i];
        if (ii >= 0)
        {
            for (int j = ii; j < i; ++j)
                sum -= A[i * N + j] * b[j];
        }
        else if (sum != 0.0)
            ii = i;
        b[i] = sum;
    }
    for (int i = N - 1; i >= 0; --i)
    {
        double sum = b[i];
        for (int j = i + 1; j < N; ++j)
            sum -= A[i * N + j] * b[j];
        b[i] = sum / A[i * N + i];
    }
}

// Register-resident forms of the two routines above for small atoms (N <= 7: beyond that the matrix no longer fits the register file): every loop is
// unrolled over the compile-time N and the pivoting row exchanges are selects, so the matrix
// never lives in local memory.  Same operations in the same order (including the reference's
// "all candidates zero => row 0" pivot and the skip-leading-zeros forward substitution).
template <int N>
__device__ __forceinline__ bool lu_decompose_small(double (&A)[N][N], int (&index)[N])
{
    constexpr double Tiny = 1e-20;
    double vv[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
        double big = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j)
        {
            const double v = fabs(A[i][j]);
            big = (big < v) ? v : big;
        }
        if (big == 0.0)
            return false;
        vv[i] = 1.0 / big;
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
    {
#pragma unroll
        for (int i = 0; i < j; ++i)
        {
            double sum = A[i][j];
#pragma unroll
            for (int k = 0; k < i; ++k)
                sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
        }
        int iMax = 0;
        double big = 0.0;
#pragma unroll
        for (int i = j; i < N; ++i)
        {
            double sum = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k)
                sum -= A[i][k] * A[k][j];
            A[i][j] = sum;
            const double cand = vv[i] * fabs(sum);
            if (big < cand)
            {
                big = cand;
                iMax = i;
            }
        }
        // exchange rows j and iMax (iMax is j, a later row, or -- degenerate -- row 0)
#pragma unroll
        for (int k = 0; k < N; ++k)
        {
            const double tj = A[j][k];
            double ti = tj;
#pragma unroll
            for (int r = 0; r < N; ++r)
                if ((r == 0 || r > j) && r != j)
                    ti = (r == iMax) ? A[r][k] : ti;
#pragma unroll
            for (int r = 0; r < N; ++r)
                if ((r == 0 || r > j) && r != j)
                    A[r][k] = (r == iMax) ? tj : A[r][k];
            A[j][k] = ti;
        }
#pragma unroll
        for (int r = 0; r < N; ++r)
            if ((r == 0 || r > j) && r != j)
                vv[r] = (r == iMax) ? vv[j] : vv[r];
        index[j] = iMax;
        if (A[j][j] == 0.0)
            A[j][j] = Tiny;
        const double temp = 1.0 / A[j][j];
#pragma unroll
        for (int i = j + 1; i < N; ++i)
            A[i][j] *= temp;
    }
    return true;
}

template <int N>
__device__ __forceinline__ void lu_backsub_small(const double (&A)[N][N], const int (&index)[N], double (&b)[N])
{
    int ii = -1;
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
        const int ip = index[i];
        const double bi = b[i];
        double sum = bi; // b[ip], then b[ip] = b[i]
#pragma unroll
        for (int r = 0; r < N; ++r)
            if ((r == 0 || r > i) && r != i)
            {
                sum = (r == ip) ? b[r] : sum;
                b[r] = (r == ip) ? bi : b[r];
            }
        if (ii >= 0)
        {
#pragma unroll
            for (int j = 0; j < i; ++j)
                if (j >= ii)
                    sum -= A[i][j] * b[j];
        }
        else if (sum != 0.0)
            ii = i;
        b[i] = sum;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i)
    {
        double sum = b[i];
#pragma unroll
        for (int j = i + 1; j < N; ++j)
            sum -= A[i][j] * b[j];
        b[i] = sum / A[i][i];
    }
}

// One stat-eq / backward-Euler system of an N-level atom in registers; false: singular.
template <int N>
__device__ __noinline__ bool population_solve_small(const double* __restrict__ g, double* __restrict__ nk, size_t K,
                                                    double nTot, const double* __restrict__ nOldK, double dt)
{
    double A[N][N], b[N], bCopy[N], res[N];
    int index[N];
    int iElim = 0;
    double nMax = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
        const double ni = nk[(size_t)i * K];
        if (nMax < ni)
        {
            nMax = ni;
            iElim = i;
        }
    }
    // the system as the reference builds it; evaluated again for the refinement residual
    auto elem = [&](int i, int j) {
        const double gij = g[(size_t)(i * N + j) * K];
        if (nOldK)
            return (i == j) ? 1.0 - gij * dt : -gij * dt;
        return (i == iElim) ? 1.0 : gij;
    };
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
#pragma unroll
        for (int j = 0; j < N; ++j)
            A[i][j] = elem(i, j);
        b[i] = nOldK ? nOldK[(size_t)i * K] : ((i == iElim) ? nTot : 0.0);
        bCopy[i] = b[i];
    }
    if (!lu_decompose_small<N>(A, index))
        return false;
    lu_backsub_small<N>(A, index, b);
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
        double r = bCopy[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
            r -= elem(i, j) * b[j];
        res[i] = r;
    }
    lu_backsub_small<N>(A, index, res);
#pragma unroll
    for (int i = 0; i < N; ++i)
        nk[(size_t)i * K] = b[i] + res[i];
    return true;
}

// nOld != nullptr: time_dependent_update_impl (UpdatePopulations.cpp:120-151) instead -- the
// backward-Euler system (1 - Gamma dt) n = nOld of the same atom, depth by depth, through the
// same solve_lin_eq.  nOld is [Ncol][Nlevel][K] of atom atomSel.
// MAXN = 64 (atoms of 33..64 levels): the matrix and its copy do not live in per-thread local memory (64 KB per
// thread, reserved for every resident thread of the device) but in a global scratch block of the launch's threads.
template <int MAXN>
__global__ void stat_eq_kernel(const DevProblem P, int atomSel, int kStart, int kEnd,
                               const double* __restrict__ gamma, double* __restrict__ n,
                               const double* __restrict__ nTotal, int* __restrict__ nSingular,
                               const double* __restrict__ nOld, double dt, double* __restrict__ scratch = nullptr)
{
    // atomSel >= 0: that atom; atomSel < 0: every active atom in ONE launch (the systems of
    // different atoms are independent; in 1D there are only Nspace of them per atom)
    const int nk = kEnd - kStart;
    const int nAtomLoop = atomSel >= 0 ? 1 : P.Natom;
    const size_t total = (size_t)nAtomLoop * P.Ncol * nk;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = kStart + idx % nk;
        const int col = (idx / nk) % P.Ncol;
        const int atom = atomSel >= 0 ? atomSel : (int)(idx / ((size_t)nk * P.Ncol));
        if (P.atomDetailed[atom] || (P.colActive && !P.colActive[col]))
            continue;
        const int N = P.atomNlevel[atom];
        const size_t gOff = ((size_t)col * P.GammaTot + P.atomGammaOff[atom]) * P.K + k;
        const size_t nOff = ((size_t)col * P.NlevTot + P.atomLevOff[atom]) * P.K + k;
        if (N <= 7)
        {
            const double nTot = nOld ? 0.0 : nTotal[((size_t)col * P.Natom + atom) * P.K + k];
            const double* nOldK = nOld ? nOld + (size_t)col * N * P.K + k : nullptr;
            bool ok = true;
            switch (N)
            {
            case 1: ok = population_solve_small<1>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            case 2: ok = population_solve_small<2>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            case 3: ok = population_solve_small<3>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            case 4: ok = population_solve_small<4>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            case 5: ok = population_solve_small<5>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            case 6: ok = population_solve_small<6>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            default: ok = population_solve_small<7>(gamma + gOff, n + nOff, P.K, nTot, nOldK, dt); break;
            }
            if (!ok)
                atomicAdd_system(nSingular, 1); // may be mapped host memory
            continue;
        }
        constexpr int NLOC = MAXN > 32 ? 1 : MAXN * MAXN;
        double ALoc[NLOC], ACopyLoc[NLOC], b[MAXN], bCopy[MAXN], res[MAXN];
        double* A = ALoc;
        double* ACopy = ACopyLoc;
        if (MAXN > 32)
        {
            A = scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2 * MAXN * MAXN;
            ACopy = A + MAXN * MAXN;
        }
        int index[MAXN];
        int iElim = 0;
        double nMax = 0.0;
        for (int i = 0; i < N; ++i)
        {
            const double ni = n[nOff + (size_t)i * P.K];
            if (nMax < ni)
            {
                nMax = ni;
                iElim = i;
            }
            for (int j = 0; j < N; ++j)
                A[i * N + j] = gamma[gOff + (size_t)(i * N + j) * P.K];
        }
        if (nOld)
        {
            for (int i = 0; i < N; ++i)
            {
                b[i] = nOld[((size_t)col * N + i) * P.K + k];
                for (int j = 0; j < N; ++j)
                    A[i * N + j] = -A[i * N + j] * dt;
                A[i * N + i] = 1.0 - gamma[gOff + (size_t)(i * N + i) * P.K] * dt;
            }
        }
        else
        {
            for (int i = 0; i < N; ++i)
            {
                A[iElim * N + i] = 1.0;
                b[i] = 0.0;
            }
            b[iElim] = nTotal[((size_t)col * P.Natom + atom) * P.K + k];
        }
        for (int i = 0; i < N * N; ++i)
            ACopy[i] = A[i];
        for (int i = 0; i < N; ++i)
            bCopy[i] = b[i];
        if (!lu_decompose_dev<MAXN>(N, A, index))
        {
            atomicAdd_system(nSingular, 1); // may be mapped host memory
            continue;
        }
        lu_backsub_dev(N, A, index, b);
        for (int i = 0; i < N; ++i)
        {
            double r = bCopy[i];
            for (int j = 0; j < N; ++j)
                r -= ACopy[i * N + j] * b[j];
            res[i] = r;
        }
        lu_backsub_dev(N, A, index, res);
        for (int i = 0; i < N; ++i)
            n[nOff + (size_t)i * P.K] = b[i] + res[i];
    }
}

// ---------------------------------------------------------------------------
// nr_post_update_impl with F / Ftd (UpdatePopulations.cpp:230-394): one Newton-Raphson step of the
// coupled rate + charge-conservation system, (sum Nlevel + 1)^2 unknowns per (column, depth).  One
// thread per system; the matrix and its copy (for the refinement step of solve_lin_eq) live in a global
// scratch slice per system, the right-hand side in local memory.
struct NrAtom
{
    int atom;        // index into the problem's atoms
    int N;           // levels
    int levOff;      // row of level 0 in the packed population arrays
    int gammaOff;    // row of Gamma(0, 0) in the packed Gamma arrays
    int transBeg, transEnd;
    long long cOff;  // element offset of C(0, 0, 0) within one column of the C buffer
    long long dcOff; // element offset of dC(0, 0, 0) within one column of the dC buffer, -1: none
    long long prevOff; // element offset of nPrev(0, 0) within one column of the nPrev buffer
    int stageOff;    // offset into the packed stages array
};

template <int MAXN>
__global__ void nr_update_kernel(const DevProblem P, const NrAtom* __restrict__ atoms, int Natom, int Neqn,
                                 int kStart, int kEnd, const double* __restrict__ gamma, double* __restrict__ n,
                                 const double* __restrict__ nTotal, const double* __restrict__ cmat,
                                 long long cColStride, const double* __restrict__ dC, long long dcColStride,
                                 const double* __restrict__ nPrev, long long prevColStride,
                                 const double* __restrict__ stages, const double* __restrict__ bgNe,
                                 double* __restrict__ ne, int timeDep, double dt, double crswVal,
                                 double* __restrict__ scratch, int* __restrict__ nSingular)
{
    const int nk = kEnd - kStart, K = P.K;
    const size_t total = (size_t)P.Ncol * nk;
    const double theta = 1.0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = kStart + idx % nk;
        const int col = idx / nk;
        double* dF = scratch + idx * 2 * (size_t)Neqn * Neqn;
        double* dFCopy = dF + (size_t)Neqn * Neqn;
        double Fg[MAXN], FgCopy[MAXN], res[MAXN];
        int index[MAXN];
        for (int q = 0; q < Neqn * Neqn; ++q)
            dF[q] = 0.0;
        const double nek = ne[(size_t)col * K + k];
        for (int q = 0; q < Neqn; ++q)
            Fg[q] = 0.0;
        Fg[Neqn - 1] = nek;
        int start = 0;
        for (int a = 0; a < Natom; ++a)
        {
            const NrAtom at = atoms[a];
            const int N = at.N;
            const double* G = gamma + ((size_t)col * P.GammaTot + at.gammaOff) * K + k;
            const double* nn = n + ((size_t)col * P.NlevTot + at.levOff) * K + k;
            for (int l = 0; l < N; ++l)
            {
                double f = 0.0;
                if (timeDep)
                {
                    for (int ll = 0; ll < N; ++ll)
                        f += G[(size_t)(l * N + ll) * K] * nn[(size_t)ll * K];
                    f *= theta * dt;
                    f -= nn[(size_t)l * K] - nPrev[(size_t)col * prevColStride + at.prevOff + (size_t)l * K + k];
                }
                else
                {
                    for (int ll = 0; ll < N; ++ll)
                        f -= G[(size_t)(l * N + ll) * K] * nn[(size_t)ll * K];
                }
                Fg[start + l] = f;
            }
            double nTotCur = 0.0, eleContrib = 0.0;
            for (int ll = 0; ll < N; ++ll)
                nTotCur += nn[(size_t)ll * K];
            Fg[start + N - 1] = nTotCur - nTotal[((size_t)col * P.Natom + at.atom) * K + k];
            for (int ll = 0; ll < N; ++ll)
                eleContrib += stages[at.stageOff + ll] * nn[(size_t)ll * K];
            Fg[Neqn - 1] -= eleContrib;
            start += N;
        }
        Fg[Neqn - 1] -= bgNe[(size_t)col * K + k];

        start = 0;
        for (int a = 0; a < Natom; ++a)
        {
            const NrAtom at = atoms[a];
            const int N = at.N;
            const double* G = gamma + ((size_t)col * P.GammaTot + at.gammaOff) * K + k;
            const double* Cm = cmat + (size_t)col * cColStride + at.cOff + k;
            const double* nn = n + ((size_t)col * P.NlevTot + at.levOff) * K + k;
            for (int l = 0; l < N; ++l)
                for (int ll = 0; ll < N; ++ll)
                {
                    double v = -G[(size_t)(l * N + ll) * K];
                    if (timeDep)
                    {
                        v *= -theta * dt;
                        if (l == ll)
                            v -= 1.0;
                    }
                    dF[(start + l) * Neqn + start + ll] = v;
                }
            for (int g = at.transBeg; g < at.transEnd; ++g)
            {
                const DevTrans& t = P.trans[g];
                if (t.type != 0)
                {
                    const double preconRji = G[(size_t)(t.i * N + t.j) * K] - crswVal * Cm[(size_t)(t.i * N + t.j) * K];
                    double entry = -(preconRji / nek) * nn[(size_t)t.j * K];
                    if (timeDep)
                        entry *= -theta * dt;
                    dF[(start + t.i) * Neqn + Neqn - 1] += entry;
                }
            }
            if (at.dcOff >= 0)
            {
                const double* d = dC + (size_t)col * dcColStride + at.dcOff + k;
                for (int i = 0; i < N; ++i)
                {
                    double entry = 0.0;
                    for (int ll = 0; ll < N; ++ll)
                        entry -= d[(size_t)(i * N + ll) * K] * nn[(size_t)ll * K];
                    if (timeDep)
                        entry *= -theta * dt;
                    dF[(start + i) * Neqn + Neqn - 1] += entry;
                }
            }
            for (int q = 0; q < Neqn; ++q)
                dF[(start + N - 1) * Neqn + q] = 0.0;
            for (int ll = 0; ll < N; ++ll)
            {
                dF[(start + N - 1) * Neqn + start + ll] = 1.0;
                dF[(Neqn - 1) * Neqn + start + ll] = -stages[at.stageOff + ll];
            }
            start += N;
        }
        dF[(Neqn - 1) * Neqn + Neqn - 1] = 1.0;
        for (int i = 0; i < Neqn; ++i)
        {
            Fg[i] *= -1.0;
            FgCopy[i] = Fg[i];
        }
        for (int q = 0; q < Neqn * Neqn; ++q)
            dFCopy[q] = dF[q];
        // solve_lin_eq with one refinement step (LuSolve.cpp:103-133)
        if (!lu_decompose_dev<MAXN>(Neqn, dF, index))
        {
            atomicAdd_system(nSingular, 1); // may be mapped host memory
            continue;
        }
        lu_backsub_dev(Neqn, dF, index, Fg);
        for (int i = 0; i < Neqn; ++i)
        {
            double r = FgCopy[i];
            for (int j = 0; j < Neqn; ++j)
                r -= dFCopy[i * Neqn + j] * Fg[j];
            res[i] = r;
        }
        lu_backsub_dev(Neqn, dF, index, res);
        for (int i = 0; i < Neqn; ++i)
            Fg[i] += res[i];
        start = 0;
        for (int a = 0; a < Natom; ++a)
        {
            const NrAtom at = atoms[a];
            double* nn = n + ((size_t)col * P.NlevTot + at.levOff) * K + k;
            for (int ll = 0; ll < at.N; ++ll)
                nn[(size_t)ll * K] += Fg[start + ll];
            start += at.N;
        }
        ne[(size_t)col * K + k] = nek + Fg[Neqn - 1];
    }
}

// (max, first index) over dJ[Ncol][L] restricted to [laLo, laHi): what the
// reference's threaded branch returns (:688, :700-703).  One launch: every CTA reduces a
// grid-strided share into part[], the last one to finish (ticket counter) reduces the parts.
__device__ __forceinline__ void dj_block_reduce(double* sMax, long long* sIdx)
{
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1)
    {
        if (threadIdx.x < s)
        {
            const double o = sMax[threadIdx.x + s];
            const long long oi = sIdx[threadIdx.x + s];
            if (sMax[threadIdx.x] < o || (sMax[threadIdx.x] == o && oi < sIdx[threadIdx.x]))
            {
                sMax[threadIdx.x] = o;
                sIdx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
dj_reduce_kernel(const double* __restrict__ dJ, int Ncol, int L, int laLo, int laHi, double* outMax,
                 long long* outIdx, const unsigned char* __restrict__ laMask, double* partMax,
                 long long* partIdx, unsigned* ticket, const unsigned char* __restrict__ colActive)
{
    __shared__ double sMax[256];
    __shared__ long long sIdx[256];
    __shared__ bool last;
    double best = -1.0;
    long long bestIdx = 0;
    const long long span = laHi - laLo;
    const long long total = (long long)Ncol * span;
    // (col, la) of this thread's elements advance by a fixed stride: no division in the loop
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long dcol = span > 0 ? stride / span : 0, dla = span > 0 ? stride % span : 0;
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long col = span > 0 ? q / span : 0, la = span > 0 ? q % span : 0;
    for (; q < total; q += stride)
    {
        if ((!laMask || laMask[laLo + la]) && (!colActive || colActive[col]))
        {
            const long long idx = col * L + laLo + la;
            const double v = dJ[idx];
            if (best < v)
            {
                best = v;
                bestIdx = idx;
            }
        }
        col += dcol;
        la += dla;
        if (la >= span)
        {
            la -= span;
            ++col;
        }
    }
    sMax[threadIdx.x] = best;
    sIdx[threadIdx.x] = bestIdx;
    dj_block_reduce(sMax, sIdx);
    if (threadIdx.x == 0)
    {
        partMax[blockIdx.x] = sMax[0];
        partIdx[blockIdx.x] = sIdx[0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last)
        return;
    __threadfence();
    const bool have = threadIdx.x < gridDim.x; // gridDim.x <= 256
    sMax[threadIdx.x] = have ? ((volatile double*)partMax)[threadIdx.x] : -1.0;
    sIdx[threadIdx.x] = have ? ((volatile long long*)partIdx)[threadIdx.x] : 0;
    dj_block_reduce(sMax, sIdx);
    if (threadIdx.x == 0)
    {
        *outMax = sMax[0] < 0.0 ? 0.0 : sMax[0];
        *outIdx = sIdx[0];
        *ticket = 0;
    }
}

} // namespace lwb200
