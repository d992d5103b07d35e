// lwb200_ray2.cuh -- stage 2 of the pipeline for the production case (piecewise Bezier3, one warp per
// wavelength, Nspace <= 128, both directions of every ray): ray_smem_kernel.
//
// What ncu showed on B200 (config 3) about ray_kernel (lwb200_pipeline.cuh), and what this kernel does
// about it:
//   * 255 registers, 8 warps per SM, issue slots 45 % busy.  Everything that lives across rays -- the
//     column's depth geometry (per CTA: its four warps work on wavelengths of one column), the
//     per-wavelength constants chiC, etaC, sigma J-dagger and the line coefficients, AND the moment
//     accumulators (per warp) -- is kept in shared memory, [chunk][lane] order (conflict-free 64-bit
//     accesses), one owner per element: 168 registers, three CTAs per SM.
//   * with both directions' coefficient code inlined the hot loop was ~40 KB of SASS: it did not fit the
//     32 KB instruction cache and `no_instruction` was the second stall reason (1.5 warps per issue).
//     The Bezier3 coefficients are evaluated per INTERVAL by one direction-independent block
//     (interval_coeffs, lwb200_fsm.cuh) -- once per mu when the two directions share their opacities --
//     the per-direction rest is a few fused multiply-adds, and the ray loop is not unrolled.
//   * the register prefetch of the profile rows defeated itself (scoreboards are counters: this ray's
//     move out of the prefetch register also waited for the loads just issued for the next ray, 22 % of all
//     long-scoreboard stalls on one instruction): rows are pulled into the L1 two rays ahead by prefetch
//     hints and read where they are used; the next wavelength's rows go to the L2 the same way.
// Measured: config 3 launch set 14.1 -> 12.9 ms per 512 columns, config 2 0.509 -> 0.492 ms.
#pragma once
#include "lwb200_pipeline.cuh"

namespace lwb200
{
// depth geometry of a column in shared memory; the pointer is this lane's element of chunk 0 of array 0
template <int NCH>
struct GeomS
{
    const double* p;
    int K;
    __device__ __forceinline__ double dsf(int j) const { return p[j * 32]; }                  // |h_k - h_{k+1}|
    __device__ __forceinline__ double rdsf(int j) const { return p[(NCH + j) * 32]; }         // 1 / dsf
    __device__ __forceinline__ double wU(int j) const { return p[(2 * NCH + j) * 32]; }       // dsf / (dsf + dsfP)
    __device__ __forceinline__ double wD(int j) const { return p[(3 * NCH + j) * 32]; }       // dsfP / (dsf + dsfP)
};

// bezier3_prepare (lwb200_fsm.cuh) with the geometry read from shared memory
template <int NCH>
__device__ __forceinline__ void bezier3_prepare_s(const GeomS<NCH>& g, int lane, const double (&chi)[NCH],
                                                  const double (&S)[NCH], double muz, double zmu, RayPre<NCH>& r)
{
    const int K = g.K;
    double chiN[NCH];
    shift_next<NCH>(chi, chiN);
    shift_next<NCH>(S, r.SN);

    double sl[NCH], Df[NCH], DfN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        sl[j] = (chiN[j] - chi[j]) * (g.rdsf(j) * muz);
    const double slUp = __shfl_up_sync(kFull, sl[NCH - 1], 1);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const double slPj = (j == 0) ? slUp : sl[j > 0 ? j - 1 : 0];
        double d = steffen_r(g.wU(j), g.wD(j), slPj, sl[j]);
        d = sel(k == 0, sl[j], d);       // one-sided at the top    (:239 / :288)
        d = sel(k == K - 1, slPj, d);    // one-sided at the bottom
        Df[j] = d;
    }
    shift_next<NCH>(Df, DfN);

    double dtfP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double ds = g.dsf(j) * zmu;
        const double ds3 = ds * (1.0 / 3.0);
        const double cA = fma(ds3, Df[j], chi[j]);
        const double cB = fma(-ds3, DfN[j], chiN[j]);
        const double t1 = chi[j] + chiN[j];
        r.dtf[j] = ds * ((t1 + cA) + cB) * 0.25;
        r.rdtf[j] = rcp_fast(r.dtf[j]);
    }
    shift_prev<NCH>(r.dtf, dtfP);

    double slS[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        slS[j] = (r.SN[j] - S[j]) * r.rdtf[j];
    const double slSUp = __shfl_up_sync(kFull, slS[NCH - 1], 1);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const double slSPj = (j == 0) ? slSUp : slS[j > 0 ? j - 1 : 0];
        const double rs = rcp_fast(r.dtf[j] + dtfP[j]);
        double d = steffen_r(r.dtf[j] * rs, dtfP[j] * rs, slSPj, slS[j]);
        d = sel(k == 0, slS[j], d);      // (:247)
        d = sel(k == K - 1, slSPj, d);
        r.DSf[j] = d;
    }
}

// ---------------------------------------------------------------------------
// ray_smem_kernel: one ray at a time like ray_kernel, but everything that lives across rays -- the depth
// geometry, the per-wavelength constants AND the moment accumulators -- is kept in shared memory with a
// single owner per element (the lane that owns the depth), so the kernel fits MINB = 3 CTAs of four warps
// on an SM instead of two: the recurrence is bound by the latency a scheduler's two warps cannot hide, and
// registers are what stood in the way of a third and fourth.
// Same arithmetic and accumulation order as ray_kernel.
#ifndef LWB200_RAY_WARPS
#define LWB200_RAY_WARPS 4
#endif
template <int NCH, int NL, int MINB>
__global__ void __launch_bounds__(32 * LWB200_RAY_WARPS, MINB)
ray_smem_kernel(const DevProblem P, const int* __restrict__ lamList, int nLam, int perWarp, int colBase,
                int lambdaIterate, int fsMode)
{
    const bool fsOnly = (fsMode & 1) != 0;
    const bool noMoments = (fsMode & 4) != 0;
    const bool noScatter = (fsMode & 8) != 0;
    constexpr int NLA = NL > 0 ? NL : 1;
    constexpr int NPAIR = NL > 1 ? NL * (NL - 1) / 2 : 0;
    constexpr int NCONST = (3 + 2 * NLA) > 6 ? (3 + 2 * NLA) : 6; // (line-free wavelengths park their 6 ray-independent arrays here)
    constexpr int NMOM = 2 + 4 * (NL > 0 ? NL : 0) + NPAIR;       // mJ, mP, per line {mW, mA, mB0, mB}, cross terms
    constexpr int RS = 32 * NCH;
    extern __shared__ double dynS[];
    double* geomS = dynS;                                           // [4][RS], per CTA
    const int K = P.K, M = P.M, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);
    const int lane = lane_id();
    constexpr int NPROW = NL > 0 ? NL : 0;    // this ray's profile rows, parked for the moment phase
    double* warpS = dynS + 4 * RS + (size_t)(warp % LWB200_RAY_WARPS) * (NCONST + NMOM + NPROW + 1) * RS;
    double* cS = warpS + lane;                // constants: this lane's element of chunk 0 of array 0
    double* mS = warpS + NCONST * RS + lane;  // moments
    double* pS = warpS + (NCONST + NMOM) * RS + lane;          // profile rows of the current ray
    double* jS = warpS + (NCONST + NMOM + NPROW) * RS + lane;  // J-dagger of this wavelength

    {
        DepthComm<false> cm{nullptr, 0, 1, 0};
        GeometryR<NCH> g;
        load_geometry_r<NCH>(cm, g, P.height + (size_t)col * K, K);
        if (warp == 0)
        {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                geomS[j * 32 + lane] = g.dsf[j];
                geomS[(NCH + j) * 32 + lane] = g.rdsf[j];
                geomS[(2 * NCH + j) * 32 + lane] = g.dsf[j] * g.rsum[j];
                geomS[(3 * NCH + j) * 32 + lane] = g.dsfP[j] * g.rsum[j];
            }
        }
    }
    __syncthreads();
    const int first = (int)(blockIdx.x * (blockDim.x >> 5) + warp) * perWarp;
    if (first >= nLam)
        return;
    GeomS<NCH> g{geomS + lane, K};
    const double dsTop = fabs(__ldg(P.height + (size_t)col * K) - __ldg(P.height + (size_t)col * K + 1));
    const double dsBot = fabs(__ldg(P.height + (size_t)col * K + K - 2) - __ldg(P.height + (size_t)col * K + K - 1));
    const double* Tcol = P.temperature + (size_t)col * K;
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;
    auto cst = [&](int arr, int j) -> double& { return cS[(arr * NCH + j) * 32]; };
    auto momr = [&](int arr, int j) -> double& { return mS[(arr * NCH + j) * 32]; };
    auto prow = [&](int arr, int j) -> double& { return pS[(arr * NCH + j) * 32]; };

    for (int q = first; q < min(first + perWarp, nLam); ++q)
    {
        const int la = lamList[q];
        const double lambda = __ldg(P.wavelength + la);
        const double rlambda = 1.0 / lambda;
        const size_t rowLK = ((size_t)col * L + la) * K;
        const size_t rowB = ((size_t)cb * L + la) * K;
        constexpr double hc_4pi = 0.25 * kHC / kPi;

        __syncwarp();
        const double* ph[NLA];
        const double* phRay[NLA];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            const bool v = k < K;
            cst(0, j) = v ? __ldg(P.chiC + rowB + k) : 1.0;
            cst(1, j) = v ? __ldg(P.etaC + rowB + k) : 0.0;
            const double sca = v ? __ldg(P.scaBg + rowLK + k) : 0.0;
            const double JDag = (v && !noScatter) ? P.J[rowLK + k] : 0.0;
            cst(2, j) = sca * JDag;
            jS[j * 32] = v ? P.J[rowLK + k] : 0.0;
#pragma unroll
            for (int m = 0; m < NMOM; ++m)
                momr(m, j) = 0.0;
        }
#pragma unroll
        for (int l = 0; l < NLA; ++l)
        {
            ph[l] = phRay[l] = P.phi;
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
                    double x = 0.0, e = 0.0;
                    if (k < K)
                    {
                        const double ni = __ldg(ncol + (size_t)ll.levI * K + k);
                        const double nj = __ldg(ncol + (size_t)ll.levJ * K + k);
                        const double gk = rho ? gS * __ldg(rho + k) : gS;
                        x = vB * (ni - nj * gk);
                        e = nj * (ll.Aji_Bji * (gk * vB));
                    }
                    cst(3 + 2 * l, j) = x;
                    cst(4 + 2 * l, j) = e;
                }
            }
        }
        __syncwarp();

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

        // ---- boundary intensity and last-point coefficients of up to 32 rays at once (ray_endpoints)
        RayEnds endsV{0.0, 0.0, 0.0, 0.0};
        int endsBase = -1;
        auto compute_ends = [&](int rayBase) {
            const int kq[4] = {0, 1, K - 2, K - 1};
            const double* cw = warpS;
            const int rr = rayBase + lane;
            if (rr < 2 * M)
            {
                const int mu = rr >> 1, dir = rr & 1;
                const double zmu = 1.0 / __ldg(P.muz + mu);
                double chiK[4], SK[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                {
                    const int o = (kq[qq] % NCH) * 32 + kq[qq] / NCH; // [chunk][lane] position of depth kq
                    double c = cw[0 * RS + o], e = cw[1 * RS + o];
                    if (NL > 0)
                    {
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
                        {
                            const double pq = __ldg(phRay[l] + (size_t)rr * K + kq[qq]);
                            c = fma(cw[(3 + 2 * l) * RS + o], pq, c);
                            e = fma(cw[(4 + 2 * l) * RS + o], pq, e);
                        }
                    }
                    chiK[qq] = c;
                    SK[qq] = (e + cw[2 * RS + o]) / c;
                }
                const int bcType = dir ? P.lowerBc : P.upperBc;
                double bcValue = 0.0;
                if (bcType == 4)
                    bcValue = dir ? P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]]
                                  : P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
                endsV = ray_endpoints(chiK, SK, dsTop, dsBot, zmu, dir, bcType, dir ? Bbot0 : Btop0,
                                      dir ? Bbot1 : Btop1, bcValue);
            }
            endsBase = rayBase;
        };
        compute_ends(0);
        // the next wavelength's rows of chiC, etaC, sigma and J-dagger: into the L2 / L1 while this one is solved
        if (q + 1 < min(first + perWarp, nLam) && lane * NCH < K)
        {
            const int laN = lamList[q + 1];
            const size_t nLK = ((size_t)col * L + laN) * K + lane * NCH, nB = ((size_t)cb * L + laN) * K + lane * NCH;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.chiC + nB));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.etaC + nB));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.scaBg + nLK));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.J + nLK));
        }

        const int phiFlags = NL == 0 ? 0 : __ldg(P.phiAsym);
        const bool shareDir = (phiFlags & 1) == 0;
        // static atmosphere: the profiles do not depend on the ray at all, a line wavelength is then solved
        // like a line-free one (chi and S once per wavelength) and its line moments are products of the
        // profile with the two scalar moments
        const bool muShare = NL > 0 && phiFlags == 0;

        // line-free wavelengths: chi and S are the same for every ray; interpolation data once at mu = 1,
        // parked in the constants' rows (chiC, etaC, sigma J-dagger are not needed any more)
        if (NL == 0 || muShare)
        {
            double chi[NCH], S[NCH], rchi[NCH];
            RayPre<NCH> pre1;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                double c = cst(0, j), e = cst(1, j);
                if (NL > 0)
                {
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        const double pq = (lane * NCH + j < K) ? __ldg(ph[l] + j) : 0.0;
                        c = fma(cst(3 + 2 * l, j), pq, c);
                        e = fma(cst(4 + 2 * l, j), pq, e);
                        prow(l, j) = pq;
                    }
                }
                chi[j] = c;
                rchi[j] = rcp_fast(c);
                S[j] = (e + cst(2, j)) * rchi[j]; // compute_source_fn (:169-179)
            }
            bezier3_prepare_s<NCH>(g, lane, chi, S, 1.0, 1.0, pre1);
            __syncwarp(); // (compute_ends has read the constants of other lanes)
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                cst(0, j) = S[j];
                cst(1, j) = rchi[j];
                cst(2, j) = pre1.SN[j];
                cst(3, j) = pre1.dtf[j];
                cst(4, j) = pre1.rdtf[j];
                cst(5, j) = pre1.DSf[j];
            }
        }
        // profile rows are pulled into the L1 ahead of their use (prefetch hints, no scoreboard: a register
        // prefetch would make this ray's loads wait for the next ray's, the scoreboards being counters)
        auto prefetch_row = [&](int row) {
            if (NL > 0 && row < 2 * M && lane * NCH < K)
            {
#pragma unroll
                for (int l = 0; l < NLA; ++l)
                {
                    const double* src = ph[l] + (size_t)row * K;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
                    if (NCH > 1)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (NCH - 1)));
                }
            }
        };
        if (!muShare)
        {
            prefetch_row(shareDir ? 2 : 1);
            prefetch_row(shareDir ? 4 : 2);
        }

        // The profile rows travel through registers ONE RAY AHEAD: they are loaded right after the previous
        // ray's opacities are formed (into the registers those just freed -- the rows the moment phase needs
        // are parked in shared memory meanwhile), so a ray never waits for the L1/L2 latency of its own loads.
        double S[NCH], rchi[NCH], SN[NCH], DSf[NCH], p[NLA][NCH];
        auto load_row = [&](int row) {
            if (NL > 0 && row < 2 * M)
            {
#pragma unroll
                for (int l = 0; l < NLA; ++l)
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                        p[l][j] = (lane * NCH + j < K) ? __ldg(ph[l] + (size_t)row * K + j) : 0.0;
            }
        };
        if (!muShare)
            load_row(0);
        RayCoef<NCH> coef;
        const int laneE = (K - 1) / NCH, jE = (K - 1) % NCH; // where the deepest point lives
#pragma unroll 1
        for (int ray = 0; ray < 2 * M; ++ray)
        {
            const int mu = ray >> 1, dir = ray & 1;
            const double w = 0.5 * __ldg(P.wmu + mu);
            if (ray >= endsBase + 32)
                compute_ends(ray & ~31);
            if (dir == 0 || !shareDir)
            {
                // ---- opacity, source function, the direction-independent solver phase and the interval
                // coefficients: once per mu when the two directions share their profiles, else per ray
                const double muz = __ldg(P.muz + mu);
                const double zmu = rcp_fast(muz);
                RayPre<NCH> pre;
                if (NL == 0 || muShare)
                {
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        S[j] = cst(0, j);
                        rchi[j] = cst(1, j);
                        pre.SN[j] = cst(2, j);
                        pre.dtf[j] = cst(3, j) * zmu;
                        pre.rdtf[j] = cst(4, j) * muz;
                        pre.DSf[j] = cst(5, j) * muz;
                    }
                }
                else
                {
                    double chi[NCH];
                    prefetch_row(shareDir ? ray + 6 : ray + 3);
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        double c = cst(0, j), e = cst(1, j);
#pragma unroll
                        for (int l = 0; l < NLA; ++l)
                        {
                            c = fma(cst(3 + 2 * l, j), p[l][j], c);
                            e = fma(cst(4 + 2 * l, j), p[l][j], e);
                            prow(l, j) = p[l][j];
                        }
                        chi[j] = c;
                        rchi[j] = rcp_fast(c);
                        S[j] = (e + cst(2, j)) * rchi[j]; // compute_source_fn (:169-179)
                    }
                    load_row(shareDir ? ray + 2 : ray + 1);
                    bezier3_prepare_s<NCH>(g, lane, chi, S, muz, zmu, pre);
                }
                interval_coeffs<NCH>(pre, coef);
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    SN[j] = pre.SN[j];
                    DSf[j] = pre.DSf[j];
                }
            }
            const int rl = ray - endsBase;
            RayEnds ends;
            ends.Iupw = __shfl_sync(kFull, endsV.Iupw, rl);
            ends.aE = __shfl_sync(kFull, endsV.aE, rl);
            ends.bE = __shfl_sync(kFull, endsV.bE, rl);
            ends.pE = __shfl_sync(kFull, endsV.pE, rl);
            double a[NCH], b[NCH], I[NCH], psi[NCH];
            if (dir == 0)
            {
                ray_down<NCH>(coef, S, DSf, rchi, ends, lane, laneE, jE, a, b, psi);
                affine_scan<NCH, true>(a, b, I);
            }
            else
            {
                ray_up<NCH>(coef, S, SN, DSf, rchi, ends, lane, laneE, jE, a, b, psi);
                affine_scan<NCH, false>(a, b, I);
            }
            if (lane == 0)
                P.I[((size_t)col * L + la) * M + mu] = I[0];
            store_zplane<NCH>(P, lane, I, dir, ((size_t)col * L + la) * M + mu);
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const double wI = w * I[j];
                const double wP = lambdaIterate ? 0.0 : w * psi[j];
                momr(0, j) += wI;
                momr(1, j) += wP;
                if (NL > 0 && !muShare)
                {
                    double tq[NLA], pq[NLA];
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        pq[l] = prow(l, j);
                        tq[l] = wP * pq[l];
                        momr(2 + 4 * l, j) = fma(w, pq[l], momr(2 + 4 * l, j));
                        momr(3 + 4 * l, j) = fma(wI, pq[l], momr(3 + 4 * l, j));
                        momr(4 + 4 * l, j) += tq[l];
                        momr(5 + 4 * l, j) = fma(tq[l], pq[l], momr(5 + 4 * l, j));
                    }
                    if (NL > 1)
                    {
                        int pr = 0;
#pragma unroll
                        for (int a2 = 0; a2 < NLA; ++a2)
#pragma unroll
                            for (int b2 = a2 + 1; b2 < NLA; ++b2)
                            {
                                momr(2 + 4 * NL + pr, j) = fma(tq[a2], pq[b2], momr(2 + 4 * NL + pr, j));
                                ++pr;
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
        double W0 = 0.0; // sum of the ray weights, in ray order
        if (muShare)
            for (int mu = 0; mu < M; ++mu)
            {
                const double w = 0.5 * __ldg(P.wmu + mu);
                W0 += w;
                W0 += w;
            }
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            if (k < K)
            {
                const double mJ = momr(0, j);
                const double JDag = jS[j * 32];
                P.J[rowLK + k] = mJ;
                const double d = fabs(1.0 - JDag / mJ);
                dJ = (d < dJ) ? dJ : d;
                if (noMoments)
                    continue;
                mom[k] = momr(1, j);
                if (NL > 0 && muShare)
                {
                    // ray-independent profiles: sum_r w f_r p^a = p^a sum_r w f_r
                    const double mP = momr(1, j);
                    double tq[NLA], pq[NLA];
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        pq[l] = prow(l, j);
                        tq[l] = mP * pq[l];
                        mom[(size_t)(1 + 4 * l) * K + k] = W0 * pq[l];
                        mom[(size_t)(2 + 4 * l) * K + k] = mJ * pq[l];
                        mom[(size_t)(3 + 4 * l) * K + k] = tq[l];
                        mom[(size_t)(4 + 4 * l) * K + k] = tq[l] * pq[l];
                    }
                    int pr = 0;
#pragma unroll
                    for (int a2 = 0; a2 < NLA; ++a2)
#pragma unroll
                        for (int b2 = a2 + 1; b2 < NLA; ++b2)
                        {
                            mom[(size_t)(1 + 4 * NL + pr) * K + k] = tq[a2] * pq[b2];
                            ++pr;
                        }
                }
                else if (NL > 0)
                {
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        mom[(size_t)(1 + 4 * l) * K + k] = momr(2 + 4 * l, j);
                        mom[(size_t)(2 + 4 * l) * K + k] = momr(3 + 4 * l, j);
                        mom[(size_t)(3 + 4 * l) * K + k] = momr(4 + 4 * l, j);
                        mom[(size_t)(4 + 4 * l) * K + k] = momr(5 + 4 * l, j);
                    }
#pragma unroll
                    for (int pr = 0; pr < NPAIR; ++pr)
                        mom[(size_t)(1 + 4 * NL + pr) * K + k] = momr(2 + 4 * NL + pr, j);
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
        {
            const double o = __shfl_xor_sync(kFull, dJ, d);
            dJ = (o < dJ) ? dJ : o;
        }
        if (lane == 0)
            P.dJ[(size_t)col * L + la] = dJ;
    }
}

template <int NCH, int NL>
constexpr size_t ray_smem_bytes()
{
    constexpr int NLA = NL > 0 ? NL : 1;
    constexpr int NPAIR = NL > 1 ? NL * (NL - 1) / 2 : 0;
    constexpr int NCONST = (3 + 2 * NLA) > 6 ? (3 + 2 * NLA) : 6;
    constexpr int NMOM = 2 + 4 * (NL > 0 ? NL : 0) + NPAIR;
    constexpr int NPROW = NL > 0 ? NL : 0;
    return (size_t)(4 + LWB200_RAY_WARPS * (NCONST + NMOM + NPROW + 1)) * 32 * NCH * sizeof(double);
}

} // namespace lwb200
