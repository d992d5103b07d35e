// lwb200_prd.cuh -- angle-averaged partial redistribution on the device.
//
// Replaces the scalar part of redistribute_prd_lines (Source/Prd.cpp:9-124, :468-645;
// PrdTemplates.hpp:164-291): per PRD line, depth and line wavelength, the scattering
// integral of the mean intensity against Gouttebroze's GII on a fixed-step (0.15 Doppler
// widths) fine grid, normalised, giving the emission-profile ratio
//   rho(la, k) = 1 + [n_i/n_j Bij / (Pj + Qj)] (scatInt / gNorm - Jbar).
// The formal solution that follows each redistribution (formal_sol_prd_update_rates,
// PrdTemplates.hpp:18-155) is the regular pipeline restricted to the wavelengths of the PRD
// lines, with gamma_kernel in rates-only mode (lwb200_pipeline.cuh).
//
// One thread per (line wavelength, depth).  The reference caches GII per (depth,
// wavelength, fine point) until the damping changes (~6 MB per line and column); here it is
// recomputed on the fly -- ~90 evaluations of ~40 flops per thread, nothing next to one
// formal solution -- so the redistribution needs no storage and is never stale.
#pragma once
#include "lwb200_kernels.cuh"

namespace lwb200
{
struct DevPrdLine
{
    int trans;              // global transition index
    int atom, j;            // atom and upper level within it
    int levI, levJ;         // rows in the packed population arrays
    int Nblue, Nl;
    int tabOff;             // offset of the line's wavelengths in transWave
    int lineIdx;
    int Nlevel;
    int transBeg, transEnd; // global index range of the atom's transitions
    int qelRow;             // row in the packed Qelast buffer
    long long cOff;         // element offset of the atom's C(0, 0, 0) within one column of the C buffer
    long long rhoOff;       // element offset of rho(col 0, 0, 0) in the rho pool
    double Bij, lambda0;
};

// Prd.cpp:33-36
constexpr double kPrdQWing = 4.0, kPrdQCore = 2.0, kPrdQSpread = 5.0, kPrdDQ = 0.15;

// Prd.cpp:46-49
__device__ __forceinline__ double prd_G_zero(double x) { return 1.0 / (fabs(x) + sqrt(x * x + 1.273239545)); }

// Gouttebroze (1986) fast approximation of GII = PII(q_abs, q_emit) / phi(q_emit), Prd.cpp:51-124
// (resonance case, waveratio = 1)
__device__ double prd_GII(double aDamp, double qEmit, double qAbs)
{
    if (qEmit < 0.0)
    {
        qEmit = -qEmit;
        qAbs = -qAbs;
    }
    double giiCore = 0.0, coreFactor = 0.0;
    if (qEmit < kPrdQWing)
    {
        if ((qAbs < -kPrdQWing) || (qAbs > qEmit + kPrdQSpread))
            return 0.0;
        if (fabs(qAbs) <= qEmit)
            giiCore = prd_G_zero(qEmit);
        else
            giiCore = exp(qEmit * qEmit - qAbs * qAbs) * prd_G_zero(qAbs);
        if (qEmit >= kPrdQCore && qEmit <= kPrdQWing)
        {
            const double phiCore = exp(-(qEmit * qEmit));
            const double phiWing = aDamp / (sqrt(kPi) * (aDamp * aDamp + qEmit * qEmit));
            coreFactor = phiCore / (phiCore + phiWing);
        }
        else
            return giiCore;
    }
    double gii = 0.0;
    if (qEmit >= kPrdQCore)
    {
        if ((qEmit >= kPrdQWing) && (fabs(qAbs - qEmit) > kPrdQSpread))
            return 0.0;
        const double uMin = fabs((qAbs - qEmit) / 2.0);
        double giiWing = 2.0 * (1.0 - 2.0 * uMin * prd_G_zero(uMin)) * exp(-(uMin * uMin)) / (2.0 * sqrt(kPi));
        const double ratio = qAbs / qEmit;
        giiWing *= (2.75 - (2.5 - 0.75 * ratio) * ratio);
        gii = coreFactor * giiCore + (1.0 - coreFactor) * giiWing;
    }
    return gii;
}

// prd_scatter / scattering_int (Prd.cpp:468-575) with total_depop_elastic_scattering_rate
// (Prd.cpp:9-30) folded in.  Grid: (ceil(Nl_max * K / blockDim), Ncol, NprdLines).
__global__ void prd_scatter_kernel(const DevProblem P, const DevPrdLine* __restrict__ lines,
                                   const double* __restrict__ transWave, const double* __restrict__ qelast,
                                   const double* __restrict__ cmat, long long cColStride,
                                   const double* __restrict__ vBroad, const double* __restrict__ aDampBuf,
                                   double* __restrict__ rhoPool)
{
    const DevPrdLine ln = lines[blockIdx.z];
    const int K = P.K, L = P.L, col = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ln.Nl * K)
        return;
    const int la = idx / K, k = idx % K;

    // Pj + Qj: elastic collisions + every collisional and radiative rate out of the upper level
    double PjQj = qelast[((size_t)ln.qelRow * P.Ncol + col) * K + k];
    const double* C = cmat + (size_t)col * cColStride + ln.cOff;
    for (int i = 0; i < ln.Nlevel; ++i)
        PjQj += C[((size_t)i * ln.Nlevel + ln.j) * K + k];
    const double* acc = P.accum + (size_t)col * P.AccTot * K + k;
    for (int g = ln.transBeg; g < ln.transEnd; ++g)
    {
        const DevTrans& t = P.trans[g];
        if (t.j == ln.j)
            PjQj += acc[(size_t)t.accRji * K];
        if (t.i == ln.j)
            PjQj += acc[(size_t)t.accRij * K];
    }
    const double* ncol = P.n + (size_t)col * P.NlevTot * K + k;
    const double gammaPrefactor = ncol[(size_t)ln.levI * K] / ncol[(size_t)ln.levJ * K] * ln.Bij / PjQj;
    const double Jbar = acc[(size_t)P.trans[ln.trans].accRij * K] / ln.Bij;

    const double* w = transWave + ln.tabOff;
    const double vB = vBroad[((size_t)col * P.Natom + ln.atom) * K + k];
    const double aDamp = aDampBuf[((size_t)ln.lineIdx * P.Ncol + col) * K + k];
    // J(la' + Nblue, k) = Jk[la' * K]; the rest-frame mean intensity under hybrid PRD (Prd.cpp:484-499): the
    // wavelengths of a line are consecutive rows of JRest
    const double* Jk = P.JRest ? P.JRest + ((size_t)col * P.NprdLa + P.prdLaOfLa[ln.Nblue]) * K + k
                               : P.J + ((size_t)col * L + ln.Nblue) * K + k;
    auto qWave = [&](int l) { return (w[l] - ln.lambda0) * kCLight / (ln.lambda0 * vB); };

    const double qEmit = qWave(la);
    // scattering_int_range (Prd.cpp:231-259)
    double q0, qN;
    if (fabs(qEmit) < kPrdQCore)
    {
        q0 = -kPrdQWing;
        qN = kPrdQWing;
    }
    else if (fabs(qEmit) < kPrdQWing)
    {
        if (qEmit > 0.0)
        {
            q0 = -kPrdQWing;
            qN = qEmit + kPrdQSpread;
        }
        else
        {
            q0 = qEmit - kPrdQSpread;
            qN = kPrdQWing;
        }
    }
    else
    {
        q0 = qEmit - kPrdQSpread;
        qN = qEmit + kPrdQSpread;
    }
    const int Np = (int)((qN - q0) / kPrdDQ) + 1;

    // optimised_fine_linear_fixed_spacing (Prd.cpp:180-228): upper bound of the first fine
    // point, then a forward walk through the line's own grid
    const int Nt = ln.Nl;
    int it;
    if (q0 <= qWave(0))
        it = 0;
    else if (q0 >= qWave(Nt - 1))
        it = Nt - 1;
    else
    {
        int lo = 0, hi = Nt; // first index with qWave > q0
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (q0 < qWave(mid))
                hi = mid;
            else
                lo = mid + 1;
        }
        it = lo;
    }
    double xNext = it < Nt ? qWave(it) : 0.0;
    double gNorm = 0.0, scatInt = 0.0;
    double qPrime = q0;
    for (int i = 0; i < Np; ++i)
    {
        // the reference's x = xStart + i * xStep, without contraction into an fma
        const double x = __dadd_rn(q0, __dmul_rn((double)i, kPrdDQ));
        while (it < Nt && xNext <= x)
        {
            ++it;
            xNext = it < Nt ? qWave(it) : 0.0;
        }
        double JF;
        if (it == Nt)
            JF = Jk[(size_t)(Nt - 1) * K];
        else if (it == 0)
            JF = Jk[0];
        else
        {
            const double xp = qWave(it - 1);
            const double t = (x - xp) / (xNext - xp);
            JF = __dadd_rn(__dmul_rn(1.0 - t, Jk[(size_t)(it - 1) * K]), __dmul_rn(t, Jk[(size_t)it * K]));
        }
        // trapezoid with end corrections (Press et al. 4.2; Prd.cpp:541-556)
        const int edge = (i == 0 || i == Np - 1) ? 1 : ((i == 1 || i == Np - 2) ? 2 : 0);
        double gii = prd_GII(aDamp, qEmit, qPrime);
        gii = (edge == 0) ? gii * kPrdDQ : gii * (edge == 1 ? 5.0 : 13.0) / 12.0 * kPrdDQ;
        qPrime += kPrdDQ;
        gNorm += gii;
        scatInt = __dadd_rn(scatInt, __dmul_rn(JF, gii));
    }
    rhoPool[ln.rhoOff + ((size_t)col * ln.Nl + la) * K + k] = 1.0 + gammaPrefactor * (scatInt / gNorm - Jbar);
}

// Ng(0, 0, 0)::accelerate + max_change (Ng.hpp:52-62, :137-155): largest relative change of
// rho since the previous redistribution, first index on ties; then prev = cur.  One block per
// PRD line; the maximum is taken over the columns too.
__global__ void prd_change_kernel(const DevPrdLine* __restrict__ lines, int Ncol, int K,
                                  const double* __restrict__ rhoPool, double* __restrict__ prevPool,
                                  double* __restrict__ outMax, int* __restrict__ outIdx)
{
    __shared__ double sMax[1024];   // (launched with up to 1024 threads: one block per line walks Nl * K * Ncol elements,
    __shared__ long long sIdx[1024]; //  and in 1D that walk is the whole kernel)
    const DevPrdLine ln = lines[blockIdx.x];
    const long long per = (long long)ln.Nl * K, total = per * Ncol;
    double best = 0.0;
    long long bestIdx = 0;
    for (long long e = threadIdx.x; e < total; e += blockDim.x)
    {
        const double cur = rhoPool[ln.rhoOff + e], old = prevPool[ln.rhoOff + e];
        prevPool[ln.rhoOff + e] = cur;
        if (cur != 0.0)
        {
            const double change = fabs((cur - old) / cur);
            if (best < change)
            {
                best = change;
                bestIdx = e;
            }
        }
    }
    sMax[threadIdx.x] = best;
    sIdx[threadIdx.x] = bestIdx;
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
    if (threadIdx.x == 0)
    {
        outMax[blockIdx.x] = sMax[0];
        outIdx[blockIdx.x] = (int)((sIdx[0] % per) % ln.Nl); // dMaxIdx % rhoPrd.shape(0), PrdTemplates.hpp:262
    }
}

// zero_rates of the redistributed lines (PrdTemplates.hpp:34-57)
__global__ void prd_zero_rates_kernel(const DevProblem P, const DevPrdLine* __restrict__ lines, int nLines)
{
    const size_t total = (size_t)nLines * P.Ncol * P.K;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = idx % P.K;
        const int col = (idx / P.K) % P.Ncol;
        const int q = idx / ((size_t)P.K * P.Ncol);
        const DevTrans& t = P.trans[lines[q].trans];
        double* acc = P.accum + (size_t)col * P.AccTot * P.K + k;
        acc[(size_t)t.accRij * P.K] = 0.0;
        acc[(size_t)t.accRji * P.K] = 0.0;
    }
}

} // namespace lwb200
