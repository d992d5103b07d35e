// lwb200_fslong.cuh -- the general per-ray kernel for deep atmospheres (128 < Nspace <= 1024; with BIG,
// as the only formal-solution kernel, up to 4096): fs_long_kernel.
//
// What fs_kernel (lwb200_kernels.cuh) does for one warp per wavelength -- any number of overlapping lines,
// hybrid PRD (rho interpolated per ray, LwTransition.hpp:115-130; the formal solution scatters into JRest,
// SimdFullIterationTemplates.hpp:397-408), the PRD-rates-only pass of formal_sol_prd_update_rates
// (PrdTemplates.hpp:18-76) and the plain formal solution (formal_sol_impl, :721-781) -- for a column that
// does not fit one warp: a CTA per (wavelength tile, column), its warps laid over consecutive blocks of
// 128 depths (4 per lane), neighbours and the scan carry exchanged through DepthComm exactly as the
// multi-warp ray_kernel does (lwb200_pipeline.cuh).  The wavelengths of the tile are solved one after the
// other by the whole CTA.
//
// Gamma and the rates: every thread owns its depths, and a [transition][4][Nspace] tile of partial sums
// would not fit shared memory at these depths, so each ray's contribution goes straight to the column's
// packed accumulator with one fp64 RED per element (compute_full_operator_rates, :206-234, term for term).
// This is the path of the few wavelengths the moment pipeline does not carry, not the hot one.
#pragma once
#include "lwb200_pipeline.cuh"

namespace lwb200
{
inline size_t fs_long_smem(int maxNlevel, int threads) { return (size_t)2 * maxNlevel * threads * sizeof(double); }

// BIG: up to 32 warps (1024 < Nspace <= 4096) at 64 registers per thread -- the whole iteration of such a column
// runs here (a functional path: the moment pipeline's kernels are sized for <= 8 warps per column).
template <int SOLVER, bool BIG = false>
__global__ void __launch_bounds__(BIG ? 1024 : 256, 1)
fs_long_kernel(const DevProblem P, const int* __restrict__ tileList, int laLo, int laHi, int lambdaIterate,
               int upOnly, int storeDepth, int prdOnly, int fsOnly, const unsigned char* __restrict__ laMask)
{
    constexpr int NCH = 4;
    extern __shared__ double smem[];     // [2][maxNlevel][blockDim.x]: chi_atom / U_atom per level, one column per thread
    __shared__ double commBuf[7 * (BIG ? 32 : 8)];
    __shared__ double endBuf[4][2];      // chi and S of the current ray at depths 0, 1, K - 2, K - 1
    const int K = P.K, M = P.M, L = P.L;
    const int tile = tileList[blockIdx.x];
    const int col = column_of(P, blockIdx.y);
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);
    const int nthr = blockDim.x;
    DepthComm<true> cm;
    cm.buf = commBuf;
    cm.warp = warp;
    cm.nwarp = (int)(blockDim.x >> 5);
    cm.parity = 0;
    const int lane = cm.lane_global(); // position along depth in units of NCH points
    double* Xs = smem + threadIdx.x;                         // chi_atom[level] of this thread's current depth
    double* Us = smem + (size_t)P.maxNlevel * nthr + threadIdx.x;

    GeometryR<NCH> g;
    load_geometry_r<NCH>(cm, g, P.height + (size_t)col * K, K);
    const double* Tcol = P.temperature + (size_t)col * K;
    double rT[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        rT[j] = 1.0 / __ldg(Tcol + (k < K ? k : K - 1));
    }
    const double dsTop = fabs(__ldg(P.height + (size_t)col * K) - __ldg(P.height + (size_t)col * K + 1));
    const double dsBot = fabs(__ldg(P.height + (size_t)col * K + K - 2) - __ldg(P.height + (size_t)col * K + K - 1));
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;
    double* accCol = P.accum + (size_t)col * P.AccTot * K;
    const int kq[4] = {0, 1, K - 2, K - 1};

    const int tlBeg = P.tileLa[tile], tlEnd = P.tileLa[tile + 1];
    for (int tl = tlBeg; tl < tlEnd; ++tl)
    {
        // (everything that decides whether this wavelength is solved is uniform over the CTA: the barriers
        // inside DepthComm are reached by all of its threads or by none)
        const int la = P.tileLambda[tl];
        if (la < laLo || la >= laHi)
            continue;
        const int hPrdLa = P.hprdLaOfLa ? P.hprdLaOfLa[(size_t)col * L + la] : -1;
        if ((prdOnly == 1 && hPrdLa < 0) || (prdOnly == 2 && !laMask[la]))
            continue;
        const double lambda = __ldg(P.wavelength + la);
        const size_t rowLK = ((size_t)col * L + la) * K;
        const int eBeg = P.laOff[la], eEnd = eBeg + P.laCnt[la];
        const bool hasLine = P.laHasLine[la] != 0;

        // --- ray-independent part: background + continua
        double chiC[NCH], etaC[NCH], scaJ[NCH], JDag[NCH], expfac[NCH];
        constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
        const double hc_kl = hc_k / lambda;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            const bool v = k < K;
            chiC[j] = v ? __ldg(P.chiBg + rowLK + k) : 1.0;
            etaC[j] = v ? __ldg(P.etaBg + rowLK + k) : 0.0;
            const double sca = v ? __ldg(P.scaBg + rowLK + k) : 0.0;
            JDag[j] = v ? P.J[rowLK + k] : 0.0;
            scaJ[j] = sca * JDag[j];
            expfac[j] = exp(-hc_kl * rT[j]);
        }
        for (int e = eBeg; e < eEnd; ++e)
        {
            const DevTrans& t = P.trans[P.entries[e].trans];
            if (t.type == 0)
                continue;
            const int lt = la - t.Nblue;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = lane * NCH + j;
                if (k < K)
                {
                    const UV uv = trans_uv(P, t, col, lt, 0, 0, k, lambda, expfac[j]);
                    const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                    const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                    chiC[j] += ni * uv.Vij - nj * uv.Vji;
                    etaC[j] += nj * uv.Uji;
                }
            }
        }
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

        double Jnew[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            Jnew[j] = 0.0;

        for (int mu = 0; mu < M; ++mu)
        {
            const double muz = __ldg(P.muz + mu);
            const double zmu = 1.0 / muz;
            const double halfwmu = 0.5 * __ldg(P.wmu + mu);
            for (int dir = upOnly ? 1 : 0; dir < 2; ++dir)
            {
                // --- opacity, emissivity, source function of this ray
                double chi[NCH], S[NCH], rchi[NCH];
                {
                    double eta[NCH];
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        chi[j] = chiC[j];
                        eta[j] = etaC[j];
                    }
                    if (hasLine)
                    {
                        for (int e = eBeg; e < eEnd; ++e)
                        {
                            const DevTrans& t = P.trans[P.entries[e].trans];
                            if (t.type != 0)
                                continue;
                            const int lt = la - t.Nblue;
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
                            {
                                const int k = lane * NCH + j;
                                if (k < K)
                                {
                                    const UV uv = trans_uv(P, t, col, lt, mu, dir, k, lambda, 0.0);
                                    const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                                    const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                                    chi[j] += ni * uv.Vij - nj * uv.Vji;
                                    eta[j] += nj * uv.Uji;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        rchi[j] = 1.0 / chi[j];
                        S[j] = (eta[j] + scaJ[j]) / chi[j]; // compute_source_fn (:169-179)
                    }
                    if (storeDepth && !fsOnly)
                    {
                        const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K;
#pragma unroll
                        for (int j = 0; j < NCH; ++j)
                        {
                            const int k = lane * NCH + j;
                            if (k < K)
                            {
                                P.depthChi[off + k] = chi[j];
                                P.depthEta[off + k] = eta[j];
                            }
                        }
                    }
                }

                // --- boundary condition + formal solution (:344-349)
                const int bcType = dir ? P.lowerBc : P.upperBc;
                double bcValue = 0.0;
                if (bcType == 4)
                    bcValue = dir ? P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]]
                                  : P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
                double I[NCH], psi[NCH];
                if (SOLVER == 2)
                {
                    RayPre<NCH> pre;
                    bezier3_prepare<NCH>(cm, g, chi, S, muz, zmu, pre);
                    // the two special points of the ray need chi and S at four depths owned by other threads
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (lane == kq[q] / NCH)
                        {
                            endBuf[q][0] = pick<NCH>(chi, kq[q] % NCH);
                            endBuf[q][1] = pick<NCH>(S, kq[q] % NCH);
                        }
                    __syncthreads();
                    double chiK[4], SK[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        chiK[q] = endBuf[q][0];
                        SK[q] = endBuf[q][1];
                    }
                    __syncthreads();
                    const RayEnds ends = ray_endpoints(chiK, SK, dsTop, dsBot, zmu, dir, bcType, dir ? Bbot0 : Btop0,
                                                       dir ? Bbot1 : Btop1, bcValue);
                    if (dir == 0)
                        bezier3_sweep<NCH, true>(cm, g, S, rchi, pre, ends, I, psi);
                    else
                        bezier3_sweep<NCH, false>(cm, g, S, rchi, pre, ends, I, psi);
                }
                else
                    local_stencil_ray<NCH, SOLVER>(cm, g, chi, S, rchi, muz, dir == 0, bcType, dir ? Bbot0 : Btop0,
                                                   dir ? Bbot1 : Btop1, bcValue, I, psi);

                if (lane == 0)
                    P.I[((size_t)col * L + la) * M + mu] = I[0]; // spect.I(la, mu, 0) = I(0)
                store_zplane<NCH>(P, lane, I, dir, ((size_t)col * L + la) * M + mu);
                if (fsOnly)
                    continue;

                if (storeDepth)
                {
                    const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = lane * NCH + j;
                        if (k < K)
                            P.depthI[off + k] = I[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    Jnew[j] += halfwmu * I[j]; // accumulate_J (:181-190)

                if (hPrdLa >= 0)
                {
                    // rest-frame mean intensity of hybrid PRD (:397-408)
                    const size_t row = ((((size_t)col * P.NhPrd + hPrdLa) * M + mu) * 2 + dir) * K;
                    double* JRest = P.JRest + (size_t)col * P.NprdLa * K;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = lane * NCH + j;
                        if (k < K)
                            for (long long e = P.JCoeffOff[row + k]; e < P.JCoeffOff[row + k + 1]; ++e)
                                atomicAdd(JRest + (size_t)P.JCoeffIdx[e] * K + k, halfwmu * P.JCoeffFrac[e] * I[j]);
                    }
                }
                if (lambdaIterate)
                {
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                        psi[j] = 0.0;
                }

                // --- Gamma and rates, atom by atom (:411-467), one of this thread's depths at a time
                int e0 = eBeg;
                while (e0 < eEnd)
                {
                    const int atom = P.trans[P.entries[e0].trans].atom;
                    int e1 = e0 + 1;
                    while (e1 < eEnd && P.trans[P.entries[e1].trans].atom == atom)
                        ++e1;
                    const bool detailed = P.atomDetailed[atom] != 0 || prdOnly; // (prdOnly: no operator, no Gamma)
                    const int N = P.atomNlevel[atom];
#pragma unroll 1
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = lane * NCH + j;
                        if (k >= K)
                            continue;
                        const double Ij = pick<NCH>(I, j), psij = pick<NCH>(psi, j), ef = pick<NCH>(expfac, j);
                        double Ieff = Ij;
                        if (!detailed)
                        {
                            for (int m = 0; m < N; ++m)
                            {
                                Xs[(size_t)m * nthr] = 0.0;
                                Us[(size_t)m * nthr] = 0.0;
                            }
                            double etaA = 0.0;
                            for (int e = e0; e < e1; ++e)
                            {
                                const DevTrans& t = P.trans[P.entries[e].trans];
                                const UV uv = trans_uv(P, t, col, la - t.Nblue, mu, dir, k, lambda, ef);
                                const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                                const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                                const double x = ni * uv.Vij - nj * uv.Vji;
                                Xs[(size_t)t.i * nthr] += x;
                                Xs[(size_t)t.j * nthr] -= x;
                                Us[(size_t)t.j * nthr] += uv.Uji;
                                etaA += nj * uv.Uji;
                            }
                            Ieff = Ij - psij * etaA; // compute_full_Ieff (:192-204)
                        }
                        for (int e = e0; e < e1; ++e)
                        {
                            const DevTrans& t = P.trans[P.entries[e].trans];
                            if (prdOnly && t.rhoOff < 0)
                                continue; // rates of the PRD lines only (:433-434, :455-456)
                            const int lt = la - t.Nblue;
                            const UV uv = trans_uv(P, t, col, lt, mu, dir, k, lambda, ef);
                            const double wlamu = trans_wla(P, t, col, lt, k, lambda) * halfwmu;
                            double* a = accCol + k;
                            if (!detailed)
                            {
                                // compute_full_operator_rates (:218-226)
                                red_row(a, t.accIJ, K,
                                        ((uv.Uji + uv.Vji * Ieff) - (psij * Xs[(size_t)t.i * nthr] * Us[(size_t)t.j * nthr])) * wlamu);
                                red_row(a, t.accJI, K,
                                        ((uv.Vij * Ieff) - (psij * Xs[(size_t)t.j * nthr] * Us[(size_t)t.i * nthr])) * wlamu);
                            }
                            red_row(a, t.accRij, K, Ij * uv.Vij * wlamu);            // Rij (:230)
                            red_row(a, t.accRji, K, (uv.Uji + Ij * uv.Vji) * wlamu); // Rji (:231)
                        }
                    }
                    e0 = e1;
                }
            }
        }

        if (!fsOnly)
        {
            // J row and dJ = max_k |1 - Jdag/J|  (:477-485)
            double dJ = 0.0;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = lane * NCH + j;
                if (k < K)
                {
                    P.J[rowLK + k] = Jnew[j];
                    const double d = fabs(1.0 - JDag[j] / Jnew[j]);
                    dJ = (d < dJ) ? dJ : d;
                }
            }
            dJ = cm.max_all(dJ);
            if (threadIdx.x == 0)
                P.dJ[(size_t)col * L + la] = dJ;
        }
    }
}

} // namespace lwb200
