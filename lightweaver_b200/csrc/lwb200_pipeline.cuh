// lwb200_pipeline.cuh -- the production Gamma iteration as a three-stage pipeline.
//
//   continuum_kernel   per (column, wavelength, depth): background + every active
//                      continuum -> chiC, etaC  (the ray-independent part of
//                      chi_eta_aux_accum, SimdFullIterationTemplates.hpp:59-109,
//                      with gij of Atom::setup_wavelength, LwAtom.hpp:82-128)
//   ray_kernel         one warp = one wavelength of one column, lanes over depth:
//                      line opacities, source function, formal solution of the
//                      2*Nrays rays, J, dJ, emergent I, and the RAY MOMENTS
//                      sum_r w {Psi*, p, p I, p Psi*, p p' Psi*} written to HBM
//   gamma_kernel       per (column, wavelength tile), one thread per depth: Gamma
//                      and Rij/Rji from J and the moments
//                      (compute_full_Ieff / compute_full_operator_rates, :192-234)
//
// Why three kernels.  The formal solution is a latency-bound fp64 recurrence that
// runs at 8 warps/SM; every dependent global load issued from inside it (transition
// tables, populations, shared-memory atomics of the rate accumulation) stalls a
// whole wavelength.  Measured on B200 (differential builds, tools/variants.py): in
// the fused kernel the per-wavelength set-up was 14 % and the Gamma/rate epilogue
// 31 % of the time although together they are < 15 % of the flops.  Split off, both
// become plain streaming kernels at high occupancy, and the rate accumulation needs
// no atomics at all inside a CTA (one thread owns one depth).  The price is HBM
// traffic for chiC/etaC and the moments (~1.2x the algorithmic bytes on top), which
// is free here: the path uses < 10 % of the HBM bandwidth.
//
// Moment form (unchanged from round 1): with p = phi for a line and 1 for a
// continuum, Vij = v p, Vji = g v p, Uji = u g v p and chi_atom(m), U_atom(m),
// eta_atom are linear in the p's, so every sum over the rays of one wavelength in
// compute_full_operator_rates reduces to the moments above and the per-transition
// work runs once per wavelength instead of once per ray.
#pragma once
#include "lwb200_fsm.cuh"

namespace lwb200
{
// number of moment rows of a wavelength with NL overlapping lines:
//   Psi*, then per line {p, p I, p Psi*, p^2 Psi*}, then the cross terms p_a p_b Psi* (a < b)
__host__ __device__ constexpr int moment_rows(int NL) { return 1 + 4 * NL + NL * (NL - 1) / 2; }

// ---------------------------------------------------------------------------
// Are the profiles of the two directions of every (line wavelength, mu) bitwise identical?
// (They are whenever vlos = 0 and the profile is the default Voigt; a line model that
// overrides compute_phi may break it, which is why the data is checked, not the flags.)
// The phi pool is a sequence of [2][K] blocks.  Runs after every profile upload / generation.
// Bit 1 of the flag: do the profiles depend on mu?  (Every line's block of the pool is [..][M][2][K]: the M
// pairs of one (column, line wavelength) are consecutive, and are compared with the first of them.)
__global__ void phi_symmetry_kernel(const double* __restrict__ phi, size_t nPairs, int K, int M, int* __restrict__ asym)
{
    bool diff = false, muDiff = false;
    const size_t total = nPairs * (size_t)K;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const size_t pair = idx / K;
        const int k = (int)(idx % K);
        const double a = phi[pair * 2 * K + k], b = phi[(pair * 2 + 1) * K + k];
        const double r0 = phi[(pair / M) * M * 2 * K + k];
        diff |= (__double_as_longlong(a) != __double_as_longlong(b));
        muDiff |= (__double_as_longlong(a) != __double_as_longlong(r0));
    }
    if (__any_sync(kFull, diff) && lane_id() == 0)
        atomicOr(asym, 1);
    if (__any_sync(kFull, muDiff || diff) && lane_id() == 0)
        atomicOr(asym, 2);
}

// ---------------------------------------------------------------------------
// Stage 1: chiC, etaC of the wavelengths of one kind.  Grid (groups of `perBlock` wavelengths of the
// list, columns of the batch), one thread per depth.  Launched on the stream of that kind's ray
// kernel, so the (latency-bound) continuum stage of one kind overlaps the rays of the others.
__global__ void continuum_kernel(const DevProblem P, const int* __restrict__ lamList, int nLam, int perBlock,
                                 int colBase)
{
    const int K = P.K, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int k = threadIdx.x;
    if (k >= K)
        return;
    const double Tk = __ldg(P.temperature + (size_t)col * K + k);
    const double* ncol = P.n + (size_t)col * P.NlevTot * K + k;
    const double* gcol = P.gRatio + (size_t)col * K + k;
    const int qBeg = blockIdx.x * perBlock, qEnd = min(nLam, qBeg + perBlock);
    for (int q = qBeg; q < qEnd; ++q)
    {
        const int la = lamList[q];
        const double lambda = __ldg(P.wavelength + la);
        const double rlambda = 1.0 / lambda;
        constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
        constexpr double twoHc = 2.0 * kHC / (kNmToM * kNmToM * kNmToM);
        const double hc_kl = hc_k * rlambda;
        const double hcl = twoHc * (rlambda * rlambda * rlambda);
        const size_t rowLK = ((size_t)col * L + la) * K + k;
        double chiC = __ldg(P.chiBg + rowLK);
        double etaC = __ldg(P.etaBg + rowLK);
        const double expfac = exp_fast_underflow(-hc_kl / Tk);
        const int eBeg = P.laOff[la], eEnd = eBeg + P.laCnt[la];
        for (int e = eBeg; e < eEnd; ++e)
        {
            const DevEntry& t = P.entries[e];
            if (t.type == 0)
                continue;
            const double al = t.al;
            const double gk = __ldg(gcol + t.gOff) * expfac;
            const double Vji = gk * al;
            const double ni = __ldg(ncol + t.nOffI);
            const double nj = __ldg(ncol + t.nOffJ);
            chiC += ni * al - nj * Vji;
            etaC += nj * (hcl * Vji);
        }
        const size_t o = ((size_t)cb * L + la) * K + k;
        P.chiC[o] = chiC;
        P.etaC[o] = etaC;
    }
}

// ---------------------------------------------------------------------------
// Stage 2: the rays.
#ifndef LWB200_RAY_MINBLOCKS
#define LWB200_RAY_MINBLOCKS 2
#endif

// MULTI = false: one warp per wavelength (Nspace <= 32 * NCH), `perWarp` wavelengths per warp.
// MULTI = true:  one CTA per wavelength, its warps laid over consecutive blocks of 32 * NCH
//                depths (Nspace <= blockDim.x * NCH); neighbours and the scan carry cross
//                warps through DepthComm.
// MUSHARE (Bezier3, NL > 0; chosen by the launcher from the profile flags): the profiles do not depend on
//                the ray (static atmosphere): chi, S and the direction-independent solver phase once per
//                wavelength as at line-free wavelengths, the line moments as products at the end.
template <int NCH, int SOLVER, int NL, bool MULTI, bool MUSHARE = false>
__global__ void __launch_bounds__(MULTI ? 256 : 128, MULTI ? 1 : LWB200_RAY_MINBLOCKS)
ray_kernel(const DevProblem P, const int* __restrict__ lamList, int nLam, int perWarp, int colBase,
           int lambdaIterate, int storeDepth, int fsMode)
{
    // fsMode: 0 Gamma iteration; 1 formal solution only (formal_sol_impl, :721-781: emergent I,
    // nothing else written); bit 1 (2): up-going rays only (upOnly); 4: J, dJ and I but no moments
    // (the unpolarised wavelengths of a J-updating full-Stokes pass)
    // 8: no scattering term in the source function (J-dagger taken as 0): what the reference's
    // full-Stokes pass does when it does not update J (stokes_fs_core only fills JDag under updateJ,
    // FormalStokes.cpp:431-441, :593)
    const bool fsOnly = (fsMode & 1) != 0;
    const bool noMoments = (fsMode & 4) != 0;
    const bool noScatter = (fsMode & 8) != 0;
    const int dirFirst = (fsMode & 2) ? 1 : 0;
    constexpr int NLA = NL > 0 ? NL : 1;
    constexpr int NPAIR = NL > 1 ? NL * (NL - 1) / 2 : 1;
    __shared__ double commBuf[MULTI ? 7 * 8 : 1];
    // chiC, etaC, scaJ and the line coefficients at the depths 0, 1, K - 2, K - 1 (ray_endpoints):
    // per warp when a warp owns a wavelength, once per CTA when the CTA does
    constexpr int NEND = 3 + 2 * (NL > 0 ? NL : 1);
    __shared__ double endBufAll[MULTI ? 1 : 4][4][NEND];
    const int K = P.K, M = P.M, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    // warp index made provably warp-uniform: every loop below stays convergent
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);
    DepthComm<MULTI> cm;
    cm.buf = commBuf;
    cm.warp = MULTI ? warp : 0;
    cm.nwarp = MULTI ? (int)(blockDim.x >> 5) : 1;
    cm.parity = 0;
    const int lane = cm.lane_global(); // position along depth in units of NCH points
    const int first = MULTI ? (int)blockIdx.x : (int)(blockIdx.x * (blockDim.x >> 5) + warp) * perWarp;
    if (MULTI)
        perWarp = 1;
    if (first >= nLam)
        return;

    GeometryR<NCH> g;
    load_geometry_r<NCH>(cm, g, P.height + (size_t)col * K, K);
    double(*endBuf)[NEND] = endBufAll[MULTI ? 0 : (warp & 3)];
    const double dsTop = fabs(__ldg(P.height + (size_t)col * K) - __ldg(P.height + (size_t)col * K + 1));
    const double dsBot = fabs(__ldg(P.height + (size_t)col * K + K - 2) - __ldg(P.height + (size_t)col * K + K - 1));
    const double* Tcol = P.temperature + (size_t)col * K;
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;

    for (int q = first; q < min(first + perWarp, nLam); ++q)
    {
        const int la = lamList[q];
        const double lambda = __ldg(P.wavelength + la);
        const double rlambda = 1.0 / lambda;
        const size_t rowLK = ((size_t)col * L + la) * K;
        const size_t rowB = ((size_t)cb * L + la) * K;
        constexpr double hc_4pi = 0.25 * kHC / kPi;

        double chiC[NCH], etaC[NCH], scaJ[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            const bool v = k < K;
            chiC[j] = v ? __ldg(P.chiC + rowB + k) : 1.0;
            etaC[j] = v ? __ldg(P.etaC + rowB + k) : 0.0;
            const double sca = v ? __ldg(P.scaBg + rowLK + k) : 0.0;
            const double JDag = (v && !noScatter) ? P.J[rowLK + k] : 0.0;
            scaJ[j] = sca * JDag;
        }
        // line slots: chi = chiC + sum_l cX_l phi_l, eta = etaC + sum_l cE_l phi_l
        double cX[NLA][NCH], cE[NLA][NCH];
        const double* ph[NLA];
        const double* phRay[NLA]; // the same profile rows without this lane's depth offset
#pragma unroll
        for (int l = 0; l < NLA; ++l)
        {
            ph[l] = P.phi;
            phRay[l] = P.phi;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                cX[l][j] = cE[l][j] = 0.0;
            if (NL > 0)
            {
                const LambdaLine& ll = P.lamLine[(size_t)la * 3 + l];
                // constants of Transition::uv (LwTransition.hpp:93-130)
                const double vB = hc_4pi * (ll.lambda0 * rlambda) * ll.Bij;
                const double gS = ll.Bji_Bij;
                const double* rho = (ll.rhoOff >= 0) ? P.rhoPrd + ll.rhoOff + (size_t)col * ll.rhoColStride : nullptr;
                phRay[l] = P.phi + ll.phiOff + (size_t)col * ll.phiColStride;
                ph[l] = phRay[l] + lane * NCH;
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = lane * NCH + j;
                    if (k < K)
                    {
                        const double ni = __ldg(ncol + (size_t)ll.levI * K + k);
                        const double nj = __ldg(ncol + (size_t)ll.levJ * K + k);
                        const double gk = rho ? gS * __ldg(rho + k) : gS;
                        cX[l][j] = vB * (ni - nj * gk);
                        cE[l][j] = nj * (ll.Aji_Bji * (gk * vB));
                    }
                }
            }
        }

        // thermalised boundaries: Planck function at the two boundary pairs (once per wavelength)
        double Btop0 = 0.0, Btop1 = 0.0, Bbot0 = 0.0, Bbot1 = 0.0;
        if (P.upperBc == 2)
        {
            Btop0 = planck_nu(__ldg(Tcol + 0), lambda);
            Btop1 = planck_nu(__ldg(Tcol + 1), lambda);
        }
        if (P.lowerBc == 2)
        {
            Bbot0 = planck_nu(__ldg(Tcol + K - 1), lambda);
            Bbot1 = planck_nu(__ldg(Tcol + K - 2), lambda);
        }

        // ---- Bezier3: boundary intensity and last-point coefficients of up to 32 rays at once, lane r taking
        // ray rayBase + r = 2 mu + dir (ray_endpoints, lwb200_fsm.cuh)
        RayEnds endsV{0.0, 0.0, 0.0, 0.0};
        int endsBase = -1;
        auto compute_ends = [&](int rayBase) {
            const int kq[4] = {0, 1, K - 2, K - 1};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane == kq[q] / NCH)
                {
                    const int jq = kq[q] % NCH;
                    endBuf[q][0] = pick<NCH>(chiC, jq);
                    endBuf[q][1] = pick<NCH>(etaC, jq);
                    endBuf[q][2] = pick<NCH>(scaJ, jq);
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        endBuf[q][3 + 2 * l] = pick<NCH>(cX[l], jq);
                        endBuf[q][4 + 2 * l] = pick<NCH>(cE[l], jq);
                    }
                }
            if (MULTI)
                __syncthreads();
            else
                __syncwarp();
            const int rr = rayBase + lane_id();
            if (rr < 2 * M)
            {
                const int mu = rr >> 1, dir = rr & 1;
                const double zmu = 1.0 / __ldg(P.muz + mu);
                double chiK[4], SK[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    double c = endBuf[q][0], e = endBuf[q][1];
                    if (NL > 0)
                    {
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
                        {
                            const double pq = __ldg(phRay[l] + (size_t)rr * K + kq[q]);
                            c = fma(endBuf[q][3 + 2 * l], pq, c);
                            e = fma(endBuf[q][4 + 2 * l], pq, e);
                        }
                    }
                    chiK[q] = c;
                    SK[q] = (e + endBuf[q][2]) / c;
                }
                const int bcType = dir ? P.lowerBc : P.upperBc;
                double bcValue = 0.0;
                if (bcType == 4)
                    bcValue = dir ? P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]]
                                  : P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
                endsV = ray_endpoints(chiK, SK, dsTop, dsBot, zmu, dir, bcType, dir ? Bbot0 : Btop0,
                                      dir ? Bbot1 : Btop1, bcValue);
            }
            if (MULTI)
                __syncthreads();
            else
                __syncwarp();
            endsBase = rayBase;
        };

        // ---- moments over the rays of this wavelength
        //   mJ = sum w I, mP = sum w Psi*, mW[l] = sum w p_l, mA[l] = sum w p_l I,
        //   mB0[l] = sum w Psi* p_l, mB[l] = sum w Psi* p_l^2, mBx[(a,b)] = sum w Psi* p_a p_b (a < b)
        double mJ[NCH], mP[NCH], mW[NLA][NCH], mA[NLA][NCH], mB0[NLA][NCH], mB[NLA][NCH], mBx[NPAIR][NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            mJ[j] = mP[j] = 0.0;
#pragma unroll
            for (int l = 0; l < NLA; ++l)
                mW[l][j] = mA[l][j] = mB0[l][j] = mB[l][j] = 0.0;
#pragma unroll
            for (int pq = 0; pq < NPAIR; ++pq)
                mBx[pq][j] = 0.0;
        }
        double chi[NCH], S[NCH], rchi[NCH], p[NLA][NCH];
        RayPre<NCH> pre;
        // the up and down rays of one mu share opacities, source function and the whole
        // direction-independent phase of the solver when their profiles are identical
        constexpr bool MS = MUSHARE && SOLVER == 2 && NL > 0;
        const bool shareDir = MS || ((SOLVER == 2) && !storeDepth && dirFirst == 0 && ((__ldg(P.phiAsym) & 1) == 0));

        // line-free wavelengths: chi and S are the same for every ray; interpolation data once at mu = 1
        RayPre<NCH> pre1;
        if (NL == 0 || MS)
        {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                double c = chiC[j], e = etaC[j];
                if (MS)
                {
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        p[l][j] = (lane * NCH + j < K) ? __ldg(ph[l] + j) : 0.0;
                        c = fma(cX[l][j], p[l][j], c);
                        e = fma(cE[l][j], p[l][j], e);
                    }
                }
                else
                    p[0][j] = 0.0;
                chi[j] = c;
                rchi[j] = rcp_fast(c);
                S[j] = (e + scaJ[j]) * rchi[j]; // compute_source_fn (:169-179)
            }
            if (SOLVER == 2)
                bezier3_prepare<NCH>(cm, g, chi, S, 1.0, 1.0, pre1);
        }
        // profiles are fetched one ray ahead (NL <= 2; three lines leave no registers for it)
        constexpr bool PREFETCH = (NL == 1 || NL == 2) && !MS;
        double pn[NLA][NCH];
        if (PREFETCH)
        {
#pragma unroll
            for (int l = 0; l < NLA; ++l)
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    pn[l][j] = (lane * NCH + j < K) ? __ldg(ph[l] + (size_t)dirFirst * K + j) : 0.0;
        }

        for (int mu = 0; mu < M; ++mu)
        {
            const double muz = __ldg(P.muz + mu);
            const double zmu = rcp_fast(muz);
            const double w = 0.5 * __ldg(P.wmu + mu);
            if ((NL == 0 || MS) && SOLVER == 2)
            {
                const RayPre<NCH>& q1 = pre1;
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    pre.SN[j] = q1.SN[j];
                    pre.dtf[j] = q1.dtf[j] * zmu;
                    pre.rdtf[j] = q1.rdtf[j] * muz;
                    pre.DSf[j] = q1.DSf[j] * muz;
                }
            }
#pragma unroll
            for (int dir = 0; dir < 2; ++dir)
            {
                if (dir < dirFirst)
                    continue;
                if (NL > 0 && !MS && (dir == 0 || !shareDir))
                {
                    const int row = 2 * mu + dir;
                    if (PREFETCH)
                    {
                        const int nextRow = (shareDir || dirFirst) ? row + 2 : row + 1;
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
                            {
                                p[l][j] = pn[l][j];
                                if (nextRow < 2 * M)
                                    pn[l][j]
                                        = (lane * NCH + j < K) ? __ldg(ph[l] + (size_t)nextRow * K + j) : 0.0;
                            }
                    }
                    else
                    {
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
                                p[l][j] = (lane * NCH + j < K) ? __ldg(ph[l] + (size_t)row * K + j) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = lane * NCH + j;
                        double c = chiC[j], e = etaC[j];
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
                        {
                            c = fma(cX[l][j], p[l][j], c);
                            e = fma(cE[l][j], p[l][j], e);
                        }
                        chi[j] = c;
                        rchi[j] = rcp_fast(c);
                        S[j] = (e + scaJ[j]) * rchi[j]; // compute_source_fn (:169-179)
                        if (storeDepth && k < K)
                        {
                            const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k;
                            P.depthChi[off] = c;
                            P.depthEta[off] = e;
                        }
                    }
                    if (SOLVER == 2)
                        bezier3_prepare<NCH>(cm, g, chi, S, muz, zmu, pre);
                }
                else if (NL == 0 && storeDepth)
                {
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = lane * NCH + j;
                        if (k < K)
                        {
                            const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k;
                            P.depthChi[off] = chiC[j];
                            P.depthEta[off] = etaC[j];
                        }
                    }
                }
                double I[NCH], psi[NCH];
                if (SOLVER == 2)
                {
                    const int ray = 2 * mu + dir;
                    if (endsBase < 0 || ray >= endsBase + 32)
                        compute_ends(ray & ~31);
                    const int rl = ray - endsBase;
                    RayEnds ends;
                    ends.Iupw = __shfl_sync(kFull, endsV.Iupw, rl);
                    ends.aE = __shfl_sync(kFull, endsV.aE, rl);
                    ends.bE = __shfl_sync(kFull, endsV.bE, rl);
                    ends.pE = __shfl_sync(kFull, endsV.pE, rl);
                    if (dir == 0)
                        bezier3_sweep<NCH, true>(cm, g, S, rchi, pre, ends, I, psi);
                    else
                        bezier3_sweep<NCH, false>(cm, g, S, rchi, pre, ends, I, psi);
                }
                else
                {
                    int bcType;
                    double bcB0, bcB1, bcValue = 0.0;
                    if (dir == 1)
                    {
                        bcType = P.lowerBc;
                        bcB0 = Bbot0;
                        bcB1 = Bbot1;
                        if (bcType == 4)
                            bcValue = P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]];
                    }
                    else
                    {
                        bcType = P.upperBc;
                        bcB0 = Btop0;
                        bcB1 = Btop1;
                        if (bcType == 4)
                            bcValue = P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
                    }
                    local_stencil_ray<NCH, SOLVER>(cm, g, chi, S, rchi, muz, dir == 0, bcType, bcB0, bcB1, bcValue, I,
                                                   psi);
                }

                if (lane == 0)
                    P.I[((size_t)col * L + la) * M + mu] = I[0];
                store_zplane<NCH>(P, lane, I, dir, ((size_t)col * L + la) * M + mu);
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = lane * NCH + j;
                    if (storeDepth && k < K)
                        P.depthI[((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k] = I[j];
                    const double wI = w * I[j];
                    const double wP = lambdaIterate ? 0.0 : w * psi[j];
                    mJ[j] += wI;
                    mP[j] += wP;
                    if (NL > 0 && !MS)
                    {
                        double tq[NLA];
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
                        {
                            tq[l] = wP * p[l][j];
                            mW[l][j] = fma(w, p[l][j], mW[l][j]);
                            mA[l][j] = fma(wI, p[l][j], mA[l][j]);
                            mB0[l][j] += tq[l];
                            mB[l][j] = fma(tq[l], p[l][j], mB[l][j]);
                        }
                        if (NL > 1)
                        {
                            int pr = 0;
#pragma unroll
                            for (int a = 0; a < NLA; ++a)
#pragma unroll
                                for (int b = a + 1; b < NLA; ++b)
                                {
                                    mBx[pr][j] = fma(tq[a], p[b][j], mBx[pr][j]);
                                    ++pr;
                                }
                        }
                    }
                }
            }
        }

        if (fsOnly)
            continue;
        // ---- J row, dJ (:477-485) and the moment rows
        double dJ = 0.0;
        double* mom = P.mom + ((size_t)cb * P.momRows + P.momOff[la]) * K;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            if (k < K)
            {
                const double JDag = P.J[rowLK + k];
                P.J[rowLK + k] = mJ[j];
                const double d = fabs(1.0 - JDag / mJ[j]);
                dJ = (d < dJ) ? dJ : d;
                if (noMoments)
                    continue;
                mom[k] = mP[j];
                if (MS)
                {
                    // ray-independent profiles: sum_r w f_r p^a = p^a sum_r w f_r
                    double W0 = 0.0;
                    for (int mu = 0; mu < M; ++mu)
                    {
                        const double w = 0.5 * __ldg(P.wmu + mu);
                        W0 += w;
                        W0 += w;
                    }
                    double tq[NLA];
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        tq[l] = mP[j] * p[l][j];
                        mom[(size_t)(1 + 4 * l) * K + k] = W0 * p[l][j];
                        mom[(size_t)(2 + 4 * l) * K + k] = mJ[j] * p[l][j];
                        mom[(size_t)(3 + 4 * l) * K + k] = tq[l];
                        mom[(size_t)(4 + 4 * l) * K + k] = tq[l] * p[l][j];
                    }
                    if (NL > 1)
                    {
                        int pr = 0;
#pragma unroll
                        for (int a = 0; a < NLA; ++a)
#pragma unroll
                            for (int b = a + 1; b < NLA; ++b)
                            {
                                mom[(size_t)(1 + 4 * NL + pr) * K + k] = tq[a] * p[b][j];
                                ++pr;
                            }
                    }
                }
                else if (NL > 0)
                {
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        mom[(size_t)(1 + 4 * l) * K + k] = mW[l][j];
                        mom[(size_t)(2 + 4 * l) * K + k] = mA[l][j];
                        mom[(size_t)(3 + 4 * l) * K + k] = mB0[l][j];
                        mom[(size_t)(4 + 4 * l) * K + k] = mB[l][j];
                    }
                    if (NL > 1)
                    {
#pragma unroll
                        for (int pr = 0; pr < NPAIR; ++pr)
                            mom[(size_t)(1 + 4 * NL + pr) * K + k] = mBx[pr][j];
                    }
                }
            }
        }
        dJ = cm.max_all(dJ);
        if (lane == 0)
            P.dJ[(size_t)col * L + la] = dJ;
    }
}

// ---------------------------------------------------------------------------
// Stage 3: Gamma and rates from J and the moments.  Grid (tiles, columns of the batch),
// one thread per depth; the tile's per-transition partial sums live in shared memory with
// a single writer per element (no atomics) and are flushed once with fp64 RED.
//
// Profile members of an atom at one wavelength: q = 0 the continua (p = 1), q = l + 1 the
// line in slot l.  M(q, q') = sum_r w Psi* p_q p_q'.
#ifndef LWB200_GAMMA_MINBLOCKS
#define LWB200_GAMMA_MINBLOCKS 4
#endif

// one fp64 RED into a row of a column's packed accumulator (acc points at this thread's depth)
__device__ __forceinline__ void red_row(double* acc, int row, int K, double v)
{
    if (row >= 0 && v != 0.0)
        atomicAdd(acc + (size_t)row * K, v);
}

// DIRECT = false: acc is this thread's column of the tile's shared-memory accumulator, flushed by the
// caller; DIRECT = true: acc is this thread's depth of the column's global accumulator and every
// contribution is a RED (no accumulator in shared memory: CTAs per SM set by registers alone).
template <int NL, bool DIRECT = false>
__device__ __forceinline__ void gamma_lambda(const DevProblem& P, int la, int col, int cb, int k, double Tk,
                                             double W0, double* __restrict__ acc, double* __restrict__ Xs,
                                             double* __restrict__ Us, const bool prdOnly)
{
    constexpr int NLA = NL > 0 ? NL : 1;
    constexpr int NQ = NL + 1;
    const int K = P.K;
    // shared-memory row stride: the depth count itself when one chunk covers the column (no padding
    // rows: shared memory per CTA is what sets this kernel's occupancy), else the chunk width
    const int tid = threadIdx.x, KC = (gridDim.z == 1) ? K : (int)blockDim.x, nthr = KC;
    const double lambda = __ldg(P.wavelength + la);
    const double rlambda = 1.0 / lambda;
    constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
    constexpr double twoHc = 2.0 * kHC / (kNmToM * kNmToM * kNmToM);
    constexpr double hc_4pi = 0.25 * kHC / kPi;
    const double hc_kl = hc_k * rlambda;
    const double hcl = twoHc * (rlambda * rlambda * rlambda);
    const double* ncol = P.n + (size_t)col * P.NlevTot * K + k;
    const double* gcol = P.gRatio + (size_t)col * K + k;

    // moments of this wavelength at depth k (independent loads, issued together)
    const double* mom = P.mom + ((size_t)cb * P.momRows + P.momOff[la]) * K + k;
    const double mJ = P.J[((size_t)col * P.L + la) * K + k];
    double Mq[NQ][NQ], mW[NLA], mA[NLA];
    Mq[0][0] = mom[0];
    {
        int pr = 0;
#pragma unroll
        for (int a = 0; a < NL; ++a)
        {
            mW[a] = mom[(size_t)(1 + 4 * a) * K];
            mA[a] = mom[(size_t)(2 + 4 * a) * K];
            Mq[0][a + 1] = Mq[a + 1][0] = mom[(size_t)(3 + 4 * a) * K];
            Mq[a + 1][a + 1] = mom[(size_t)(4 + 4 * a) * K];
#pragma unroll
            for (int b = a + 1; b < NL; ++b)
            {
                Mq[a + 1][b + 1] = Mq[b + 1][a + 1] = mom[(size_t)(1 + 4 * NL + pr) * K];
                ++pr;
            }
        }
    }
    const double expfac = exp_fast_underflow(-hc_kl / Tk);

    // line slots of this wavelength
    int lsTrans[NLA], lsAtom[NLA], lsI[NLA], lsJ[NLA];
    double lsV[NLA], lsGv[NLA], lsUgv[NLA], lsX[NLA], lsE[NLA], lsWla[NLA];
#pragma unroll
    for (int l = 0; l < NLA; ++l)
    {
        lsTrans[l] = lsAtom[l] = -1;
        lsI[l] = lsJ[l] = 0;
        lsV[l] = lsGv[l] = lsUgv[l] = lsX[l] = lsE[l] = lsWla[l] = 0.0;
        if (NL > 0)
        {
            const LambdaLine& ll = P.lamLine[(size_t)la * 3 + l];
            const double vB = hc_4pi * (ll.lambda0 * rlambda) * ll.Bij;
            const double gS = ll.Bji_Bij;
            const double r = (ll.rhoOff >= 0) ? __ldg(P.rhoPrd + ll.rhoOff + (size_t)col * ll.rhoColStride + k) : 1.0;
            const double gk = (ll.rhoOff >= 0) ? gS * r : gS;
            const double ni = __ldg(ncol + (size_t)ll.levI * K);
            const double nj = __ldg(ncol + (size_t)ll.levJ * K);
            lsTrans[l] = ll.trans;
            lsAtom[l] = ll.atom;
            lsI[l] = ll.i;
            lsJ[l] = ll.j;
            lsV[l] = vB;
            lsGv[l] = (gS * vB) * r;
            lsUgv[l] = (ll.Aji_Bji * (gS * vB)) * r;
            lsX[l] = vB * (ni - nj * gk);
            lsE[l] = nj * (ll.Aji_Bji * (gk * vB));
            lsWla[l] = ll.wlaS * __ldg(P.wphi + ((size_t)ll.lineIdx * P.Ncol + col) * K + k);
        }
    }

    const int eBeg = P.laOff[la], eEnd = eBeg + P.laCnt[la];
    int e0 = eBeg;
    while (e0 < eEnd)
    {
        const DevEntry& first = P.entries[e0];
        const int e1 = first.groupEnd;
        const bool detailed = first.detailed != 0;
        const int N = first.Nlevel;
        const int atom = first.atom;
        double E0 = 0.0;
        if (!detailed && !prdOnly)
        {
            // continuum aggregates per level: chi_atom / U_atom of chi_eta_aux_accum (:59-109)
            for (int m = 0; m < N; ++m)
            {
                Xs[m * nthr + tid] = 0.0;
                Us[m * nthr + tid] = 0.0;
            }
            for (int e = e0; e < e1; ++e)
            {
                const DevEntry& t = P.entries[e];
                if (t.type == 0)
                    continue;
                const double al = t.al;
                const double gk = __ldg(gcol + t.gOff) * expfac;
                const double Vji = gk * al;
                const double Uji = hcl * Vji;
                const double ni = __ldg(ncol + t.nOffI);
                const double nj = __ldg(ncol + t.nOffJ);
                const double x = ni * al - nj * Vji;
                Xs[t.i * nthr + tid] += x;
                Xs[t.j * nthr + tid] -= x;
                Us[t.j * nthr + tid] += Uji;
                E0 += nj * Uji;
            }
        }
        // line members of this atom: per-unit-phi coefficients (0 for other atoms' lines)
        double Xl[NLA], ugvl[NLA], Eq[NQ];
        Eq[0] = E0;
#pragma unroll
        for (int l = 0; l < NLA; ++l)
        {
            Xl[l] = ugvl[l] = 0.0;
            if (NL > 0)
            {
                const bool own = lsAtom[l] == atom;
                Xl[l] = own ? lsX[l] : 0.0;
                ugvl[l] = own ? lsUgv[l] : 0.0;
                Eq[l + (NL > 0 ? 1 : 0)] = own ? lsE[l] : 0.0;
            }
        }
        // EB[q] = sum_q' E_q' M(q, q')
        double EB[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
        {
            double s = 0.0;
#pragma unroll
            for (int q2 = 0; q2 < NQ; ++q2)
                s = fma(Eq[q2], Mq[q][q2], s);
            EB[q] = s;
        }

        for (int e = e0; e < e1; ++e)
        {
            const DevEntry& t = P.entries[e];
            if (prdOnly && !t.prd)
                continue; // formal_sol_prd_update_rates: rates of the PRD lines only (:433-434)
            double v = 0.0, gv = 0.0, ugv = 0.0, Wq = W0, Aq = mJ, EBq = EB[0], wla = 0.0;
            if (t.type != 0)
            {
                const double gk = __ldg(gcol + t.gOff) * expfac;
                v = t.al;
                gv = gk * t.al;
                ugv = hcl * gv;
                wla = t.wlaF;
            }
#pragma unroll
            for (int l = 0; l < NL; ++l)
            {
                if (t.type == 0 && t.trans == lsTrans[l])
                {
                    v = lsV[l];
                    gv = lsGv[l];
                    ugv = lsUgv[l];
                    Wq = mW[l];
                    Aq = mA[l];
                    EBq = EB[l + 1];
                    wla = lsWla[l];
                }
            }
            double* a4 = DIRECT ? acc : acc + t.slot * 4 * KC;
            if (!detailed && !prdOnly)
            {
                // chi_atom(m) = sum_q p_q X_q(m), U_atom(m) = sum_q p_q U_q(m)
                double Xi[NQ], Xj[NQ], Ui[NQ], Uj[NQ];
                Xi[0] = Xs[t.i * nthr + tid];
                Xj[0] = Xs[t.j * nthr + tid];
                Ui[0] = Us[t.i * nthr + tid];
                Uj[0] = Us[t.j * nthr + tid];
#pragma unroll
                for (int l = 0; l < NL; ++l)
                {
                    Xi[l + 1] = t.i == lsI[l] ? Xl[l] : (t.i == lsJ[l] ? -Xl[l] : 0.0);
                    Xj[l + 1] = t.j == lsI[l] ? Xl[l] : (t.j == lsJ[l] ? -Xl[l] : 0.0);
                    Ui[l + 1] = t.i == lsJ[l] ? ugvl[l] : 0.0;
                    Uj[l + 1] = t.j == lsJ[l] ? ugvl[l] : 0.0;
                }
                // sum_r w Psi* chi_atom(a) U_atom(b) = sum_{q,q'} X_q(a) M(q,q') U_q'(b)
                double XUij = 0.0, XUji = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                {
                    double mj = 0.0, mi = 0.0;
#pragma unroll
                    for (int q2 = 0; q2 < NQ; ++q2)
                    {
                        mj = fma(Mq[q][q2], Uj[q2], mj);
                        mi = fma(Mq[q][q2], Ui[q2], mi);
                    }
                    XUij = fma(Xi[q], mj, XUij);
                    XUji = fma(Xj[q], mi, XUji);
                }
                // sum_r w [(Uji + Vji Ieff) - Psi* chi(i) U(j)],  Ieff = I - Psi* eta_atom
                const double gIJ = (ugv * Wq + gv * (Aq - EBq) - XUij) * wla;
                const double gJI = (v * (Aq - EBq) - XUji) * wla;
                if (DIRECT)
                {
                    red_row(a4, t.accIJ, K, gIJ);
                    red_row(a4, t.accJI, K, gJI);
                }
                else
                {
                    a4[0] += gIJ;
                    a4[KC] += gJI;
                }
            }
            const double rIJ = (v * Aq) * wla, rJI = (ugv * Wq + gv * Aq) * wla;
            if (DIRECT)
            {
                red_row(a4, t.accRij, K, rIJ);
                red_row(a4, t.accRji, K, rJI);
            }
            else
            {
                a4[2 * KC] += rIJ;
                a4[3 * KC] += rJI;
            }
        }
        e0 = e1;
    }
}

// laMask / prdOnly: the rates-only pass of the PRD sub-iterations over the masked wavelengths.
__global__ void __launch_bounds__(128, LWB200_GAMMA_MINBLOCKS)
gamma_kernel(const DevProblem P, const int* __restrict__ tileList, int laLo, int laHi, int colBase,
             const unsigned char* __restrict__ laMask, int prdOnly)
{
    extern __shared__ double smem[];
    // depth chunk of this CTA: blockDim.x consecutive depths (one chunk covers Nspace <= 128)
    const int K = P.K, KC = blockDim.x, RS = (gridDim.z == 1) ? K : KC; // RS: shared-memory row stride
    const int tile = tileList[blockIdx.x];
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int kk = threadIdx.x, k = blockIdx.z * KC + kk;
    const int slot0 = P.tileSlotOff[tile];
    const int nslot = P.tileSlotOff[tile + 1] - slot0;
    double* acc = smem;                                   // [nslot][4][RS]  partial sums, one writer each
    double* Xs = smem + (size_t)P.maxSlots * 4 * RS;      // [maxNlevel][RS]
    double* Us = Xs + (size_t)P.maxNlevel * RS;
    if (k < K)
    {
        // Shared memory is private per thread here (its own depth column of every row): the
        // kernel needs no barrier and no shared-memory atomic.  What limits it is latency, i.e.
        // resident CTAs per SM, i.e. shared memory per CTA: only the accumulators live there.
        for (int s = 0; s < nslot; ++s)
        {
            acc[(s * 4 + 0) * RS + kk] = 0.0;
            acc[(s * 4 + 1) * RS + kk] = 0.0;
            acc[(s * 4 + 2) * RS + kk] = 0.0;
            acc[(s * 4 + 3) * RS + kk] = 0.0;
        }
        const double Tk = __ldg(P.temperature + (size_t)col * K + k);
        // W0 = sum_r w over both directions of every mu, in the ray order of ray_kernel
        double W0 = 0.0;
        for (int mu = 0; mu < P.M; ++mu)
        {
            const double w = 0.5 * __ldg(P.wmu + mu);
            W0 += w;
            W0 += w;
        }
        const int tlBeg = P.tileLa[tile], tlEnd = P.tileLa[tile + 1];
        for (int tl = tlBeg; tl < tlEnd; ++tl)
        {
            const int la = P.tileLambda[tl];
            if (la < laLo || la >= laHi || (laMask && !laMask[la]))
                continue;
            switch (P.laNLines[la])
            {
            case 0: gamma_lambda<0>(P, la, col, cb, k, Tk, W0, acc + kk, Xs, Us, prdOnly != 0); break;
            case 1: gamma_lambda<1>(P, la, col, cb, k, Tk, W0, acc + kk, Xs, Us, prdOnly != 0); break;
            case 2: gamma_lambda<2>(P, la, col, cb, k, Tk, W0, acc + kk, Xs, Us, prdOnly != 0); break;
            case 3: gamma_lambda<3>(P, la, col, cb, k, Tk, W0, acc + kk, Xs, Us, prdOnly != 0); break;
            default: break; // > 3 overlapping lines: handled by the general kernel
            }
        }
        // flush: this thread's own column of partial sums (one fp64 RED per element per tile)
        for (int s = 0; s < nslot; ++s)
        {
            const DevTrans& t = P.trans[P.tileSlotTrans[slot0 + s]];
            const int rows[4] = {t.accIJ, t.accJI, t.accRij, t.accRji};
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                const double v = acc[(s * 4 + q) * RS + kk];
                if (rows[q] >= 0 && v != 0.0)
                    atomicAdd(P.accum + ((size_t)col * P.AccTot + rows[q]) * K + k, v);
            }
        }
    }
}

// The Gamma stage of small problems (a 1D atmosphere, a wavelength shard): one CTA per wavelength
// of one kind, contributions go straight to the global accumulator.  No shared-memory accumulator
// means the CTAs per SM are set by the registers of that kind alone -- this stage is bound by the
// latency of its dependent loads, i.e. by resident warps -- and each kind follows its own ray kernel
// on that kernel's stream.  Column stacks keep gamma_kernel: its tiles amortise the REDs.
__host__ __device__ constexpr int gamma_direct_min_blocks(int NL)
{
    return NL == 0 ? 8 : NL == 1 ? 6 : NL == 2 ? 5 : 4;
}

template <int NL>
__global__ void __launch_bounds__(128, gamma_direct_min_blocks(NL))
gamma_direct_kernel(const DevProblem P, const int* __restrict__ lamList, int nLam, int colBase, int prdOnly)
{
    extern __shared__ double smem[];
    const int K = P.K, KC = blockDim.x, RS = (gridDim.z == 1) ? K : KC;
    const int la = lamList[blockIdx.x];
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int kk = threadIdx.x, k = blockIdx.z * KC + kk;
    double* Xs = smem; // [maxNlevel][RS], one column per thread
    double* Us = Xs + (size_t)P.maxNlevel * RS;
    if (k >= K)
        return;
    const double Tk = __ldg(P.temperature + (size_t)col * K + k);
    double W0 = 0.0;
    for (int mu = 0; mu < P.M; ++mu)
    {
        const double w = 0.5 * __ldg(P.wmu + mu);
        W0 += w;
        W0 += w;
    }
    gamma_lambda<NL, true>(P, la, col, cb, k, Tk, W0, P.accum + (size_t)col * P.AccTot * K + k, Xs, Us, prdOnly != 0);
}

} // namespace lwb200
