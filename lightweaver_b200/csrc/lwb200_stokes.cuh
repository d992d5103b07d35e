// lwb200_stokes.cuh -- polarised (full Stokes) formal solution on the device.
//
// Replaces stokes_fs_core + piecewise_stokes_bezier3_1d(_impl) (Source/FormalStokes.cpp:166-661)
// at the wavelengths where a polarised line is active; everywhere else the reference falls back
// to the scalar Bezier3 solver (:348-352) and so do we (ray_kernel).  DELO-Bezier3: per depth a
// 4x4 system Md I_k = Ma I_{k-1} + Mb S_{k-1} + Mc S_k + ... built from the propagation matrix
// K = K'/chi_I, its Steffen derivative in optical depth and the Bezier3 coefficients, solved by the
// reference's own solve_lin_eq (Crout, implicit scaled pivoting, one refinement step).
//
// One thread owns one ray (wavelength, mu, direction) of one column and sweeps depth -- the 4x4
// recurrence with a pivoted solve per depth does not map onto the affine warp scan of the scalar
// solver.  Opacities are gathered on the fly from chiC/etaC (continuum_kernel), the line slots of
// the wavelength and the seven profiles of each polarised line; a four-point window (upwind,
// centre, downwind, second downwind) lives in registers / local memory.
#pragma once
#include "lwb200_fsm.cuh"

namespace lwb200
{
struct StokesPoint
{
    double chi[7];
    double S[4];
};

// stokes_K, FormalStokes.cpp:119-143
__device__ __forceinline__ void stokes_K(const StokesPoint& p, double (&Km)[16])
{
#pragma unroll
    for (int q = 0; q < 16; ++q)
        Km[q] = 0.0;
    const double chiI = p.chi[0];
    Km[0 * 4 + 1] = p.chi[1];
    Km[0 * 4 + 2] = p.chi[2];
    Km[0 * 4 + 3] = p.chi[3];
    Km[1 * 4 + 2] = p.chi[6];
    Km[1 * 4 + 3] = p.chi[5];
    Km[2 * 4 + 3] = p.chi[4];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = j + 1; i < 4; ++i)
        {
            Km[j * 4 + i] /= chiI;
            Km[i * 4 + j] = Km[j * 4 + i];
        }
    Km[1 * 4 + 3] *= -1.0;
    Km[2 * 4 + 1] *= -1.0;
    Km[3 * 4 + 2] *= -1.0;
}

// prod(a, b, c): c(j, i) = sum_k a(k, i) b(j, k), FormalStokes.cpp:145-153
__device__ __forceinline__ void prod44(const double (&a)[16], const double (&b)[16], double (&c)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                s += a[k * 4 + i] * b[j * 4 + k];
            c[j * 4 + i] = s;
        }
}

// Bezier::cent_deriv (Bezier.hpp:58-65), the reference's own operation order
__device__ __forceinline__ double cent_deriv_ref(double dsuw, double dsdw, double yuw, double y0, double ydw)
{
    const double S0 = (ydw - y0) / dsdw;
    const double Suw = (y0 - yuw) / dsuw;
    const double P0 = fabs((Suw * dsdw + S0 * dsuw) / (dsdw + dsuw));
    const double aS = fabs(Suw), a0 = fabs(S0), hP = 0.5 * P0;
    double m = (a0 < hP) ? a0 : hP;
    m = (aS < m) ? aS : m;
    return (copysign(1.0, S0) + copysign(1.0, Suw)) * m;
}

// Bezier3_coeffs, Bezier.hpp:81-127
__device__ __forceinline__ void bezier3_coeffs_ref(double dt, double& alpha, double& beta, double& gamma,
                                                   double& delta, double& edt)
{
    const double dt2 = dt * dt, dt3 = dt2 * dt;
    if (dt < 5e-2)
    {
        edt = 1.0 - dt + 0.5 * dt2 - dt3 / 6.0;
        alpha = 0.25 * dt - 0.2 * dt2 + dt3 / 12.0;
        beta = 0.25 * dt - 0.05 * dt2 + dt3 / 120.0;
        gamma = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
        delta = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
    }
    else if (dt > 30.0)
    {
        edt = 0.0;
        alpha = 6.0 / dt3;
        beta = (-6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
        gamma = 3.0 * (2.0 * dt - 6.0) / dt3;
        delta = 3.0 * (6.0 - 4.0 * dt + dt2) / dt3;
    }
    else
    {
        edt = exp(-dt);
        alpha = (6.0 - edt * (6.0 + 6.0 * dt + 3.0 * dt2 + dt3)) / dt3;
        beta = (6.0 * edt - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
        gamma = 3.0 * (2.0 * dt - 6.0 + edt * (6.0 + 4.0 * dt + dt2)) / dt3;
        delta = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edt * (3.0 + dt)) / dt3;
    }
}

// solve_lin_eq (LuSolve.cpp:103-133) for the 4x4 system, with the refinement step
__device__ __forceinline__ void solve4(double (&A)[16], double (&b)[4])
{
    double ACopy[16], bCopy[4], res[4];
    int index[4];
#pragma unroll
    for (int q = 0; q < 16; ++q)
        ACopy[q] = A[q];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        bCopy[q] = b[q];
    lu_decompose_dev<4>(4, A, index);
    lu_backsub_dev(4, A, index, b);
    for (int i = 0; i < 4; ++i)
    {
        double r = bCopy[i];
        for (int j = 0; j < 4; ++j)
            r -= ACopy[i * 4 + j] * b[j];
        res[i] = r;
    }
    lu_backsub_dev(4, A, index, res);
    for (int i = 0; i < 4; ++i)
        b[i] += res[i];
}

// ---------------------------------------------------------------------------
// Second generation of the per-ray sweep.  What ncu showed about the first (kept below as stokes_kernel_v1,
// LWB200_STOKES_V1=1): ~2500 warp-instructions per depth point, a quarter of the issue slots used, most of them
// spent in the 60 fp64 divisions of the twenty Steffen derivatives of a step and in the local-memory traffic
// of the generic pivoted 4x4 solver.  Same recurrence, restructured around what a step really needs:
//   * K = K'/chi_I has six independent elements (zero diagonal, a symmetric and an antisymmetric triple):
//     points carry those six, derivatives are taken of those six (the derivative of -x is minus the
//     derivative of x, bit for bit), the 4x4 forms are expanded where the matrices are assembled;
//   * the upwind slope of a step is the downwind slope of the step before: slopes are carried, and all
//     twenty derivatives of a step share three reciprocals (steffen_r, lwb200_fsm.cuh);
//   * K^2 is a structured product (the zero diagonal skipped); carrying K0^2 into the next step as Ku^2 was
//     measured slower than recomputing it (sixteen more live doubles at 255 registers);
//   * the 4x4 system is solved in registers by straight-line code (solve4_reg).
// Differences from the reference's arithmetic are at rounding level (reciprocal-multiply, the solver).
struct StokesPt
{
    double chiI;  // total opacity
    double S[4];  // eta / chi_I
    double k6[6]; // chi_Q, chi_U, chi_V, rho_Q(chi[4]), rho_U(chi[5]), rho_V(chi[6]) over chi_I
};

// K(6) -> 4x4, row-major (stokes_K, FormalStokes.cpp:119-143)
__device__ __forceinline__ void stokes_expand(const double (&k)[6], double (&Km)[16])
{
    const double a = k[0], b = k[1], c = k[2], d = k[3], e = k[4], f = k[5];
    Km[0] = 0.0;  Km[1] = a;    Km[2] = b;    Km[3] = c;
    Km[4] = a;    Km[5] = 0.0;  Km[6] = f;    Km[7] = -e;
    Km[8] = b;    Km[9] = -f;   Km[10] = 0.0; Km[11] = d;
    Km[12] = c;   Km[13] = e;   Km[14] = -d;  Km[15] = 0.0;
}

// prod(K, K) for a matrix with a zero diagonal: the terms through the diagonal are skipped
__device__ __forceinline__ void stokes_square(const double (&Km)[16], double (&c)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k != i && k != j)
                    s = fma(Km[k * 4 + i], Km[j * 4 + k], s);
            c[j * 4 + i] = s;
        }
}

// Bezier3_coeffs (Bezier.hpp:81-127) with one reciprocal
__device__ __forceinline__ void bezier3_coeffs_r(double dt, double& alpha, double& beta, double& gamma,
                                                 double& delta, double& edt)
{
    const double dt2 = dt * dt, dt3 = dt2 * dt;
    if (dt < 5e-2)
    {
        edt = 1.0 - dt + 0.5 * dt2 - dt3 * (1.0 / 6.0);
        alpha = 0.25 * dt - 0.2 * dt2 + dt3 * (1.0 / 12.0);
        beta = 0.25 * dt - 0.05 * dt2 + dt3 * (1.0 / 120.0);
        gamma = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
        delta = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
        return;
    }
    const double r3 = 1.0 / dt3;
    edt = (dt > 30.0) ? 0.0 : exp(-dt);
    alpha = (6.0 - edt * (6.0 + 6.0 * dt + 3.0 * dt2 + dt3)) * r3;
    beta = (6.0 * edt - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) * r3;
    gamma = 3.0 * (2.0 * dt - 6.0 + edt * (6.0 + 4.0 * dt + dt2)) * r3;
    delta = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edt * (3.0 + dt)) * r3;
}

// The 4x4 system Md x = v.  The reference calls solve_lin_eq (LuSolve.cpp:103-133): Crout elimination with
// implicit scaled pivoting and one refinement step.  Md = 1 + O(K) with |K| < 1 (the polarised opacities
// are fractions of chi_I) is well conditioned, and any stable solver returns the same x to rounding: here
// the adjugate from 2x2 minors (Laplace expansion; straight-line code, no row exchanges) and the same one
// refinement step, x += adj(A) (v - A x) / det.
__device__ __forceinline__ void solve4_reg(const double (&a)[4][4], double (&b)[4])
{
    const double s0 = a[0][0] * a[1][1] - a[1][0] * a[0][1], s1 = a[0][0] * a[1][2] - a[1][0] * a[0][2];
    const double s2 = a[0][0] * a[1][3] - a[1][0] * a[0][3], s3 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    const double s4 = a[0][1] * a[1][3] - a[1][1] * a[0][3], s5 = a[0][2] * a[1][3] - a[1][2] * a[0][3];
    const double c5 = a[2][2] * a[3][3] - a[3][2] * a[2][3], c4 = a[2][1] * a[3][3] - a[3][1] * a[2][3];
    const double c3 = a[2][1] * a[3][2] - a[3][1] * a[2][2], c2 = a[2][0] * a[3][3] - a[3][0] * a[2][3];
    const double c1 = a[2][0] * a[3][2] - a[3][0] * a[2][2], c0 = a[2][0] * a[3][1] - a[3][0] * a[2][1];
    const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const double rd = 1.0 / det;
    double B[4][4];
    B[0][0] = a[1][1] * c5 - a[1][2] * c4 + a[1][3] * c3;
    B[0][1] = -a[0][1] * c5 + a[0][2] * c4 - a[0][3] * c3;
    B[0][2] = a[3][1] * s5 - a[3][2] * s4 + a[3][3] * s3;
    B[0][3] = -a[2][1] * s5 + a[2][2] * s4 - a[2][3] * s3;
    B[1][0] = -a[1][0] * c5 + a[1][2] * c2 - a[1][3] * c1;
    B[1][1] = a[0][0] * c5 - a[0][2] * c2 + a[0][3] * c1;
    B[1][2] = -a[3][0] * s5 + a[3][2] * s2 - a[3][3] * s1;
    B[1][3] = a[2][0] * s5 - a[2][2] * s2 + a[2][3] * s1;
    B[2][0] = a[1][0] * c4 - a[1][1] * c2 + a[1][3] * c0;
    B[2][1] = -a[0][0] * c4 + a[0][1] * c2 - a[0][3] * c0;
    B[2][2] = a[3][0] * s4 - a[3][1] * s2 + a[3][3] * s0;
    B[2][3] = -a[2][0] * s4 + a[2][1] * s2 - a[2][3] * s0;
    B[3][0] = -a[1][0] * c3 + a[1][1] * c1 - a[1][2] * c0;
    B[3][1] = a[0][0] * c3 - a[0][1] * c1 + a[0][2] * c0;
    B[3][2] = -a[3][0] * s3 + a[3][1] * s1 - a[3][2] * s0;
    B[3][3] = a[2][0] * s3 - a[2][1] * s1 + a[2][2] * s0;
    double x[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        x[i] = (B[i][0] * b[0] + B[i][1] * b[1] + B[i][2] * b[2] + B[i][3] * b[3]) * rd;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        r[i] = b[i] - (a[i][0] * x[0] + a[i][1] * x[1] + a[i][2] * x[2] + a[i][3] * x[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        b[i] = x[i] + (B[i][0] * r[0] + B[i][1] * r[1] + B[i][2] * r[2] + B[i][3] * r[3]) * rd;
}

// Grid: (ceil(nPol * 2 M / blockDim), columns of the batch).  upOnly: up-going rays only; updateJ: J (and
// J20) rebuilt with fp64 REDs (J was zeroed and copied to Jdag by the launcher).
#ifndef LWB200_STOKES_MINB
#define LWB200_STOKES_MINB 2
#endif
__global__ void __launch_bounds__(128, LWB200_STOKES_MINB)
stokes_kernel(const DevProblem P, const int* __restrict__ polLam, int nPol, int colBase, int upOnly, int updateJ)
{
    const int K = P.K, M = P.M, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= nPol * 2 * M)
        return;
    const int la = polLam[ray / (2 * M)];
    const int mu = (ray % (2 * M)) >> 1, dir = ray & 1; // dir 1 = toObs
    if (upOnly && dir == 0)
        return;
    const int NL = P.laNLines[la];
    const double lambda = __ldg(P.wavelength + la);
    const double rlambda = 1.0 / lambda;
    constexpr double hc_4pi = 0.25 * kHC / kPi;
    const size_t rowLK = ((size_t)col * L + la) * K, rowB = ((size_t)cb * L + la) * K;
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;
    // J-dagger of the scattering term: the copy taken before this J-updating pass -- or nothing: the
    // reference only fills JDag under updateJ (FormalStokes.cpp:431-441), so a pass that does not
    // update J sees a zero mean intensity in its source function (:593).  Reproduced as is.
    const double* Jsrc = updateJ ? P.Jdag + rowLK : nullptr;
    const double* height = P.height + (size_t)col * K;
    const double muz = __ldg(P.muz + mu);
    const double zmu = 1.0 / muz;
    // the "J20" extra parameter (FormalStokes.cpp:485-486, :575-583): the anisotropy of the last J-updating pass
    // scatters into the I and Q emissivities of every wavelength
    const double inv2root2 = 1.0 / (2.0 * sqrt(2.0));
    const double mu2 = muz * muz;
    const double wJ20_I = inv2root2 * (3.0 * mu2 - 1.0);
    const double wJ20_Q = inv2root2 * 3.0 * (mu2 - 1.0);
    const double* J20src = (P.j20 && P.J20dag) ? P.J20dag + rowLK : nullptr;
    const double* chiCrow = P.chiC + rowB;
    const double* etaCrow = P.etaC + rowB;
    const double* scaRow = P.scaBg + rowLK;

    // line slots of this wavelength (<= 3), constants of Transition::uv (LwTransition.hpp:93-130)
    double vB[3], gS[3], AB[3];
    const double *phiP[3], *polP[3], *rhoP[3], *niP[3], *njP[3];
    long long polArr[3];
#pragma unroll
    for (int l = 0; l < 3; ++l)
    {
        phiP[l] = polP[l] = rhoP[l] = niP[l] = njP[l] = nullptr;
        vB[l] = gS[l] = AB[l] = 0.0;
        polArr[l] = 0;
        if (l < NL)
        {
            const LambdaLine& ll = P.lamLine[(size_t)la * 3 + l];
            vB[l] = hc_4pi * (ll.lambda0 * rlambda) * ll.Bij;
            gS[l] = ll.Bji_Bij;
            AB[l] = ll.Aji_Bji;
            niP[l] = ncol + (size_t)ll.levI * K;
            njP[l] = ncol + (size_t)ll.levJ * K;
            const size_t rayOff = ((size_t)mu * 2 + dir) * K;
            phiP[l] = P.phi + ll.phiOff + (size_t)col * ll.phiColStride + rayOff;
            if (ll.polOff >= 0)
                polP[l] = P.pol + ll.polOff + (size_t)col * ll.phiColStride + rayOff;
            polArr[l] = ll.polArr;
            if (ll.rhoOff >= 0)
                rhoP[l] = P.rhoPrd + ll.rhoOff + (size_t)col * ll.rhoColStride;
        }
    }

    auto eval = [&](int k, StokesPt& pt) {
        double chi[7], eta[4];
        chi[0] = __ldg(chiCrow + k);
        eta[0] = __ldg(etaCrow + k) + (Jsrc ? __ldg(scaRow + k) * Jsrc[k] : 0.0);
#pragma unroll
        for (int q = 1; q < 7; ++q)
            chi[q] = 0.0;
        eta[1] = eta[2] = eta[3] = 0.0;
        if (J20src)
        {
            const double sj = __ldg(scaRow + k) * J20src[k];
            eta[0] += wJ20_I * sj;
            eta[1] = wJ20_Q * sj;
        }
#pragma unroll
        for (int l = 0; l < 3; ++l)
        {
            if (l < NL)
            {
                const double ni = __ldg(niP[l] + k);
                const double nj = __ldg(njP[l] + k);
                const double gk = rhoP[l] ? gS[l] * __ldg(rhoP[l] + k) : gS[l];
                const double cX = vB[l] * (ni - nj * gk);
                const double cE = nj * (AB[l] * (gk * vB[l]));
                const double ph = __ldg(phiP[l] + k);
                chi[0] = fma(cX, ph, chi[0]);
                eta[0] = fma(cE, ph, eta[0]);
                if (polP[l])
                {
                    const double* pp = polP[l] + k;
                    const double pQ = __ldg(pp), pU = __ldg(pp + polArr[l]), pV = __ldg(pp + 2 * polArr[l]);
                    chi[1] = fma(cX, pQ, chi[1]);
                    chi[2] = fma(cX, pU, chi[2]);
                    chi[3] = fma(cX, pV, chi[3]);
                    chi[4] = fma(cX, __ldg(pp + 3 * polArr[l]), chi[4]);
                    chi[5] = fma(cX, __ldg(pp + 4 * polArr[l]), chi[5]);
                    chi[6] = fma(cX, __ldg(pp + 5 * polArr[l]), chi[6]);
                    eta[1] = fma(cE, pQ, eta[1]);
                    eta[2] = fma(cE, pU, eta[2]);
                    eta[3] = fma(cE, pV, eta[3]);
                }
            }
        }
        const double rchi = 1.0 / chi[0];
        pt.chiI = chi[0];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            pt.S[q] = eta[q] * rchi;
#pragma unroll
        for (int q = 0; q < 6; ++q)
            pt.k6[q] = chi[q + 1] * rchi;
    };

    const int toObs = dir;
    int dk = -1, k_start = K - 1, k_end = 0;
    if (!toObs)
    {
        dk = 1;
        k_start = 0;
        k_end = K - 1;
    }
    StokesPt pu, p0, pd;
    eval(k_start, pu);
    eval(k_start + dk, p0);
    eval(k_start + 2 * dk, pd);

    // boundary intensity (piecewise_stokes_bezier3_1d, :354-410): only Stokes I is non-zero
    double Iprev[4] = {0.0, 0.0, 0.0, 0.0};
    {
        const double dtau_b = 0.5 * zmu * (pu.chiI + p0.chiI) * fabs(height[k_start] - height[k_start + dk]);
        const double* Tcol = P.temperature + (size_t)col * K;
        if (toObs)
        {
            if (P.lowerBc == 2)
            {
                const double B0 = planck_nu(Tcol[K - 2], lambda), B1 = planck_nu(Tcol[K - 1], lambda);
                Iprev[0] = B1 - (B0 - B1) / dtau_b;
            }
            else if (P.lowerBc == 4)
                Iprev[0] = P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]];
        }
        else
        {
            if (P.upperBc == 2)
            {
                const double B0 = planck_nu(Tcol[0], lambda), B1 = planck_nu(Tcol[1], lambda);
                Iprev[0] = B0 - (B1 - B0) / dtau_b;
            }
            else if (P.upperBc == 4)
                Iprev[0] = P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
        }
    }
    const double w = 0.5 * __ldg(P.wmu + mu);
    double* Jrow = P.J + rowLK;
    // J20(la, k) += wJ20_I wmu I + wJ20_Q wmu Q  (:642-648; the full quadrature weight, not half of it)
    double* J20row = (P.j20 && updateJ) ? P.J20 + rowLK : nullptr;
    const double wI20 = wJ20_I * __ldg(P.wmu + mu), wQ20 = wJ20_Q * __ldg(P.wmu + mu);
    if (updateJ)
        atomicAdd(Jrow + k_start, w * Iprev[0]);
    if (J20row)
        atomicAdd(J20row + k_start, wI20 * Iprev[0]);

    // set-up at the first interior point (:190-216)
    int k = k_start + dk;
    double ds_uw = fabs(height[k] - height[k - dk]) * zmu;
    double ds_dw = fabs(height[k + dk] - height[k]) * zmu;
    double dx_uw = (p0.chiI - pu.chiI) / ds_uw;
    double dx_c = cent_deriv_ref(ds_uw, ds_dw, pu.chiI, p0.chiI, pd.chiI);
    double c1 = p0.chiI - (ds_uw * (1.0 / 3.0)) * dx_c;
    double c2 = pu.chiI + (ds_uw * (1.0 / 3.0)) * dx_uw;
    double dtau_uw = ds_uw * (p0.chiI + pu.chiI + c1 + c2) * 0.25;
    // slopes over the upwind interval; they are also the one-sided derivatives at the first point
    double slK[6], slS[4], dKu[6], dSu[4];
    {
        const double r = 1.0 / dtau_uw;
#pragma unroll
        for (int q = 0; q < 6; ++q)
            dKu[q] = slK[q] = (p0.k6[q] - pu.k6[q]) * r;
#pragma unroll
        for (int n = 0; n < 4; ++n)
            dSu[n] = slS[n] = (p0.S[n] - pu.S[n]) * r;
    }
    double ds_dw2 = 0.0, dtau_dw = 0.0, dx_dw = 0.0;
#pragma unroll 1
    for (; k != k_end + dk; k += dk)
    {
        double dK0[6], dS0[4], slKd[6], slSd[4];
        StokesPt pd2 = pd;
        if (k == k_end)
        {
            // last point: one-sided derivatives (:303-311)
#pragma unroll
            for (int q = 0; q < 6; ++q)
                dK0[q] = slKd[q] = slK[q];
#pragma unroll
            for (int n = 0; n < 4; ++n)
                dS0[n] = slSd[n] = slS[n];
        }
        else
        {
            if (k_end - k == dk)
                dx_dw = (pd.chiI - p0.chiI) / ds_dw;
            else
            {
                eval(k + 2 * dk, pd2);
                ds_dw2 = fabs(height[k + 2 * dk] - height[k + dk]) * zmu;
                dx_dw = cent_deriv_ref(ds_dw, ds_dw2, p0.chiI, pd.chiI, pd2.chiI);
            }
            c1 = p0.chiI + (ds_dw * (1.0 / 3.0)) * dx_c;
            c2 = pd.chiI - (ds_dw * (1.0 / 3.0)) * dx_dw;
            dtau_dw = ds_dw * (p0.chiI + pd.chiI + c1 + c2) * 0.25;
            // Steffen derivatives of K and S in optical depth (Bezier::cent_deriv): three reciprocals for all ten
            const double rdw = 1.0 / dtau_dw, rs = 1.0 / (dtau_uw + dtau_dw);
            const double wU = dtau_dw * rs, wD = dtau_uw * rs;
#pragma unroll
            for (int q = 0; q < 6; ++q)
            {
                slKd[q] = (pd.k6[q] - p0.k6[q]) * rdw;
                dK0[q] = steffen_r(wU, wD, slK[q], slKd[q]);
            }
#pragma unroll
            for (int n = 0; n < 4; ++n)
            {
                slSd[n] = (pd.S[n] - p0.S[n]) * rdw;
                dS0[n] = steffen_r(wU, wD, slS[n], slSd[n]);
            }
        }
        double Ku[16], K0[16], dKuM[16], dK0M[16], Ku2[16], K02[16];
        stokes_expand(pu.k6, Ku);
        stokes_expand(p0.k6, K0);
        stokes_square(Ku, Ku2);
        stokes_expand(dKu, dKuM);
        stokes_expand(dK0, dK0M);
        stokes_square(K0, K02);
        double alpha, beta, gamma, delta, edt;
        bezier3_coeffs_r(dtau_uw, alpha, beta, gamma, delta, edt);
        const double dt3 = dtau_uw * (1.0 / 3.0);
        double Md[4][4], V0[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int q = i * 4 + j;
                const double idq = (i == j) ? 1.0 : 0.0;
                const double d = dt3 * (Ku2[q] + Ku[q] - dKuM[q]) - Ku[q];
                const double e = dt3 * (K02[q] + K0[q] - dK0M[q]) + K0[q];
                Md[i][j] = idq + beta * K0[q] + delta * e;
                const double Ma = edt * idq - alpha * Ku[q] + gamma * d;
                const double Mb = alpha * idq + gamma * (idq - dt3 * Ku[q]);
                const double Mc = beta * idq + delta * (idq + dt3 * K0[q]);
                v += Ma * Iprev[j] + Mb * pu.S[j] + Mc * p0.S[j];
            }
            v += dt3 * (gamma * dSu[i] - delta * dS0[i]);
            V0[i] = v;
        }
        solve4_reg(Md, V0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            Iprev[i] = V0[i];
        if (updateJ)
            atomicAdd(Jrow + k, w * V0[0]);
        if (J20row)
            atomicAdd(J20row + k, wI20 * V0[0] + wQ20 * V0[1]);
        // shuffle along (:326-338)
        pu = p0;
        p0 = pd;
        pd = pd2;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            dSu[q] = dS0[q];
            slS[q] = slSd[q];
        }
#pragma unroll
        for (int q = 0; q < 6; ++q)
        {
            dKu[q] = dK0[q];
            slK[q] = slKd[q];
        }
        dtau_uw = dtau_dw;
        ds_uw = ds_dw;
        ds_dw = ds_dw2;
        dx_uw = dx_c;
        dx_c = dx_dw;
    }
    if (toObs)
    {
        // spect.I(la, mu, 0), spect.Quv(s, la, mu, 0) = I(s, 0): the emergent up-going ray
        P.I[((size_t)col * L + la) * M + mu] = Iprev[0];
        for (int q = 0; q < 3; ++q)
            P.Quv[(((size_t)col * 3 + q) * L + la) * M + mu] = Iprev[q + 1];
    }
}

// First generation (the reference's own operation order, generic 4x4 helpers): LWB200_STOKES_V1=1.
// Grid: (ceil(nPol * 2 M / blockDim), columns of the batch).  fsMode as ray_kernel: bit 1 = up-going
// rays only, bit 2 = update J (atomically; J was zeroed and copied to Jdag by the launcher).
__global__ void __launch_bounds__(128)
stokes_kernel_v1(const DevProblem P, const int* __restrict__ polLam, int nPol, int colBase, int upOnly, int updateJ)
{
    const int K = P.K, M = P.M, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= nPol * 2 * M)
        return;
    const int la = polLam[ray / (2 * M)];
    const int mu = (ray % (2 * M)) >> 1, dir = ray & 1; // dir 1 = toObs
    if (upOnly && dir == 0)
        return;
    const int NL = P.laNLines[la];
    const double lambda = __ldg(P.wavelength + la);
    const double rlambda = 1.0 / lambda;
    constexpr double hc_4pi = 0.25 * kHC / kPi;
    const size_t rowLK = ((size_t)col * L + la) * K, rowB = ((size_t)cb * L + la) * K;
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;
    // J-dagger of the scattering term: the copy taken before this J-updating pass -- or nothing: the
    // reference only fills JDag under updateJ (FormalStokes.cpp:431-441), so a pass that does not
    // update J sees a zero mean intensity in its source function (:593).  Reproduced as is.
    const double* Jsrc = updateJ ? P.Jdag + rowLK : nullptr;
    const double* height = P.height + (size_t)col * K;
    const double zmu = 1.0 / __ldg(P.muz + mu);
    // the "J20" extra parameter (FormalStokes.cpp:485-486, :575-583): the anisotropy of the last J-updating pass
    // scatters into the I and Q emissivities of every wavelength
    const double inv2root2 = 1.0 / (2.0 * sqrt(2.0));
    const double mu2 = __ldg(P.muz + mu) * __ldg(P.muz + mu);
    const double wJ20_I = inv2root2 * (3.0 * mu2 - 1.0);
    const double wJ20_Q = inv2root2 * 3.0 * (mu2 - 1.0);
    const double* J20src = (P.j20 && P.J20dag) ? P.J20dag + rowLK : nullptr;

    // line slots of this wavelength (<= 3), constants of Transition::uv (LwTransition.hpp:93-130)
    double vB[3], gS[3], AB[3];
    const double *phiP[3], *polP[3], *rhoP[3];
    long long polArr[3];
    int levI[3], levJ[3];
    for (int l = 0; l < 3; ++l)
    {
        phiP[l] = polP[l] = rhoP[l] = nullptr;
        if (l >= NL)
            continue;
        const LambdaLine& ll = P.lamLine[(size_t)la * 3 + l];
        vB[l] = hc_4pi * (ll.lambda0 * rlambda) * ll.Bij;
        gS[l] = ll.Bji_Bij;
        AB[l] = ll.Aji_Bji;
        levI[l] = ll.levI;
        levJ[l] = ll.levJ;
        const size_t rayOff = ((size_t)mu * 2 + dir) * K;
        phiP[l] = P.phi + ll.phiOff + (size_t)col * ll.phiColStride + rayOff;
        if (ll.polOff >= 0)
            polP[l] = P.pol + ll.polOff + (size_t)col * ll.phiColStride + rayOff;
        polArr[l] = ll.polArr;
        if (ll.rhoOff >= 0)
            rhoP[l] = P.rhoPrd + ll.rhoOff + (size_t)col * ll.rhoColStride;
    }

    auto eval = [&](int k, StokesPoint& pt) {
        double chi[7], eta[4];
        chi[0] = __ldg(P.chiC + rowB + k);
        eta[0] = __ldg(P.etaC + rowB + k) + (Jsrc ? __ldg(P.scaBg + rowLK + k) * Jsrc[k] : 0.0);
#pragma unroll
        for (int q = 1; q < 7; ++q)
            chi[q] = 0.0;
        eta[1] = eta[2] = eta[3] = 0.0;
        if (J20src)
        {
            const double sj = __ldg(P.scaBg + rowLK + k) * J20src[k];
            eta[0] += wJ20_I * sj;
            eta[1] = wJ20_Q * sj;
        }
        for (int l = 0; l < NL; ++l)
        {
            const double ni = __ldg(ncol + (size_t)levI[l] * K + k);
            const double nj = __ldg(ncol + (size_t)levJ[l] * K + k);
            const double gk = rhoP[l] ? gS[l] * __ldg(rhoP[l] + k) : gS[l];
            const double cX = vB[l] * (ni - nj * gk);
            const double cE = nj * (AB[l] * (gk * vB[l]));
            const double ph = __ldg(phiP[l] + k);
            chi[0] = fma(cX, ph, chi[0]);
            eta[0] = fma(cE, ph, eta[0]);
            if (polP[l])
            {
                const double* pp = polP[l] + k;
                const double pQ = __ldg(pp), pU = __ldg(pp + polArr[l]), pV = __ldg(pp + 2 * polArr[l]);
                chi[1] = fma(cX, pQ, chi[1]);
                chi[2] = fma(cX, pU, chi[2]);
                chi[3] = fma(cX, pV, chi[3]);
                chi[4] = fma(cX, __ldg(pp + 3 * polArr[l]), chi[4]);
                chi[5] = fma(cX, __ldg(pp + 4 * polArr[l]), chi[5]);
                chi[6] = fma(cX, __ldg(pp + 5 * polArr[l]), chi[6]);
                eta[1] = fma(cE, pQ, eta[1]);
                eta[2] = fma(cE, pU, eta[2]);
                eta[3] = fma(cE, pV, eta[3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 7; ++q)
            pt.chi[q] = chi[q];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            pt.S[q] = eta[q] / chi[0];
    };

    const int toObs = dir;
    int dk = -1, k_start = K - 1, k_end = 0;
    if (!toObs)
    {
        dk = 1;
        k_start = 0;
        k_end = K - 1;
    }
    StokesPoint pu, p0, pd, pd2;
    eval(k_start, pu);
    eval(k_start + dk, p0);
    eval(k_start + 2 * dk, pd);

    // boundary intensity (piecewise_stokes_bezier3_1d, :354-410): only Stokes I is non-zero
    double Iprev[4] = {0.0, 0.0, 0.0, 0.0};
    {
        const double dtau_b = 0.5 * zmu * (pu.chi[0] + p0.chi[0]) * fabs(height[k_start] - height[k_start + dk]);
        const double* Tcol = P.temperature + (size_t)col * K;
        if (toObs)
        {
            if (P.lowerBc == 2)
            {
                const double B0 = planck_nu(Tcol[K - 2], lambda), B1 = planck_nu(Tcol[K - 1], lambda);
                Iprev[0] = B1 - (B0 - B1) / dtau_b;
            }
            else if (P.lowerBc == 4)
                Iprev[0] = P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]];
        }
        else
        {
            if (P.upperBc == 2)
            {
                const double B0 = planck_nu(Tcol[0], lambda), B1 = planck_nu(Tcol[1], lambda);
                Iprev[0] = B0 - (B1 - B0) / dtau_b;
            }
            else if (P.upperBc == 4)
                Iprev[0] = P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
        }
    }
    const double w = 0.5 * __ldg(P.wmu + mu);
    double* Jrow = P.J + rowLK;
    // J20(la, k) += wJ20_I wmu I + wJ20_Q wmu Q  (:642-648; the full quadrature weight, not half of it)
    double* J20row = (P.j20 && updateJ) ? P.J20 + rowLK : nullptr;
    const double wI20 = wJ20_I * __ldg(P.wmu + mu), wQ20 = wJ20_Q * __ldg(P.wmu + mu);
    if (updateJ)
        atomicAdd(Jrow + k_start, w * Iprev[0]);
    if (J20row)
        atomicAdd(J20row + k_start, wI20 * Iprev[0]);

    // set-up at the first interior point (:190-216)
    int k = k_start + dk;
    double ds_uw = fabs(height[k] - height[k - dk]) * zmu;
    double ds_dw = fabs(height[k + dk] - height[k]) * zmu;
    double dx_uw = (p0.chi[0] - pu.chi[0]) / ds_uw;
    double dx_c = cent_deriv_ref(ds_uw, ds_dw, pu.chi[0], p0.chi[0], pd.chi[0]);
    double c1 = p0.chi[0] - (ds_uw / 3.0) * dx_c;
    double c2 = pu.chi[0] + (ds_uw / 3.0) * dx_uw;
    double dtau_uw = ds_uw * (p0.chi[0] + pu.chi[0] + c1 + c2) * 0.25;
    double Ku[16], K0[16], Kd[16], dKu[16], dK0[16], dSu[4], dS0[4];
    stokes_K(pu, Ku);
    stokes_K(p0, K0);
#pragma unroll
    for (int n = 0; n < 4; ++n)
        dSu[n] = (p0.S[n] - pu.S[n]) / dtau_uw;
#pragma unroll
    for (int q = 0; q < 16; ++q)
    {
        dKu[q] = (K0[q] - Ku[q]) / dtau_uw;
        Kd[q] = 0.0;
        dK0[q] = 0.0;
    }
    double ds_dw2 = 0.0, dtau_dw = 0.0, dx_dw = 0.0;
    for (; k != k_end + dk; k += dk)
    {
        if (k == k_end)
        {
#pragma unroll
            for (int n = 0; n < 4; ++n)
                dS0[n] = (p0.S[n] - pu.S[n]) / dtau_uw;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                dK0[q] = (K0[q] - Ku[q]) / dtau_uw;
        }
        else
        {
            if (k_end - k == dk)
                dx_dw = (pd.chi[0] - p0.chi[0]) / ds_dw;
            else
            {
                eval(k + 2 * dk, pd2);
                ds_dw2 = fabs(height[k + 2 * dk] - height[k + dk]) * zmu;
                dx_dw = cent_deriv_ref(ds_dw, ds_dw2, p0.chi[0], pd.chi[0], pd2.chi[0]);
            }
            c1 = p0.chi[0] + (ds_dw / 3.0) * dx_c;
            c2 = pd.chi[0] - (ds_dw / 3.0) * dx_dw;
            dtau_dw = ds_dw * (p0.chi[0] + pd.chi[0] + c1 + c2) * 0.25;
            stokes_K(pd, Kd);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                dK0[q] = cent_deriv_ref(dtau_uw, dtau_dw, Ku[q], K0[q], Kd[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                dS0[q] = cent_deriv_ref(dtau_uw, dtau_dw, pu.S[q], p0.S[q], pd.S[q]);
        }
        double Ku2[16], K02[16], Md[16], V0[4];
        prod44(Ku, Ku, Ku2);
        prod44(K0, K0, K02);
        double alpha, beta, gamma, delta, edt;
        bezier3_coeffs_ref(dtau_uw, alpha, beta, gamma, delta, edt);
        const double dt3 = dtau_uw / 3.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int q = i * 4 + j;
                const double idq = (i == j) ? 1.0 : 0.0;
                const double d = dt3 * (Ku2[q] + Ku[q] - dKu[q]) - Ku[q];
                const double e = dt3 * (K02[q] + K0[q] - dK0[q]) + K0[q];
                Md[q] = idq + beta * K0[q] + delta * e;
                const double Ma = edt * idq - alpha * Ku[q] + gamma * d;
                const double Mb = alpha * idq + gamma * (idq - dt3 * Ku[q]);
                const double Mc = beta * idq + delta * (idq + dt3 * K0[q]);
                v += Ma * Iprev[j] + Mb * pu.S[j] + Mc * p0.S[j];
            }
            v += dt3 * (gamma * dSu[i] - delta * dS0[i]);
            V0[i] = v;
        }
        solve4(Md, V0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            Iprev[i] = V0[i];
        if (updateJ)
            atomicAdd(Jrow + k, w * V0[0]);
        if (J20row)
            atomicAdd(J20row + k, wI20 * V0[0] + wQ20 * V0[1]);
        // shuffle along (:326-338)
        pu = p0;
        p0 = pd;
        pd = pd2;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dSu[q] = dS0[q];
#pragma unroll
        for (int q = 0; q < 16; ++q)
        {
            Ku[q] = K0[q];
            K0[q] = Kd[q];
            dKu[q] = dK0[q];
        }
        dtau_uw = dtau_dw;
        ds_uw = ds_dw;
        ds_dw = ds_dw2;
        dx_uw = dx_c;
        dx_c = dx_dw;
    }
    if (toObs)
    {
        // spect.I(la, mu, 0), spect.Quv(s, la, mu, 0) = I(s, 0): the emergent up-going ray
        P.I[((size_t)col * L + la) * M + mu] = Iprev[0];
        for (int q = 0; q < 3; ++q)
            P.Quv[(((size_t)col * 3 + q) * L + la) * M + mu] = Iprev[q + 1];
    }
}

// J rows of the polarised wavelengths start from zero in a J-updating pass
__global__ void stokes_zero_rows_kernel(const DevProblem P, const int* __restrict__ polLam, int nPol)
{
    const int q = blockIdx.x, col = blockIdx.y;
    if (q >= nPol)
        return;
    const size_t row = ((size_t)col * P.L + polLam[q]) * P.K;
    for (int k = threadIdx.x; k < P.K; k += blockDim.x)
        P.J[row + k] = 0.0;
}

// dJ of the J-updating Stokes pass at the polarised wavelengths (stokes_fs_core, :651-659)
__global__ void stokes_dj_kernel(const DevProblem P, const int* __restrict__ polLam, int nPol)
{
    const int q = blockIdx.x, col = blockIdx.y;
    if (q >= nPol)
        return;
    const int la = polLam[q];
    const size_t row = ((size_t)col * P.L + la) * P.K;
    double dJ = 0.0;
    for (int k = threadIdx.x; k < P.K; k += 32)
    {
        const double d = fabs(1.0 - P.Jdag[row + k] / P.J[row + k]);
        dJ = (d < dJ) ? dJ : d;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const double o = __shfl_xor_sync(kFull, dJ, d);
        dJ = (o < dJ) ? dJ : o;
    }
    if (threadIdx.x == 0)
        P.dJ[(size_t)col * P.L + la] = dJ;
}

} // namespace lwb200
