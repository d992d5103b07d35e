// lwb200_fsgeneral.cuh -- the general per-ray kernel for columns of one warp (Nspace <= 128): fs_kernel.
// Any number of overlapping lines, hybrid PRD (per-ray rho, JRest), the PRD-rates-only passes, the plain
// formal solution (MODE_FS).  The formal solution itself is the pipeline's (lwb200_fsm.cuh: two-phase Bezier3 with
// reciprocals and the table-free exp, local stencils for the linear / BESSER solvers) -- the first version went
// through the reference-order routines of lwb200_device.cuh (IEEE divisions, libm exp) and was the slowest part of
// a hybrid-PRD iteration.
#pragma once
#include "lwb200_fsm.cuh"

namespace lwb200
{
template <int NCH, int SOLVER, int MODE>
__global__ void __launch_bounds__(128)
fs_kernel(const DevProblem P, const int* __restrict__ tileList, int laLo, int laHi, int lambdaIterate,
          int upOnly, int storeDepth, int prdOnly, const unsigned char* __restrict__ laMask)
{
    // prdOnly: the pass of formal_sol_prd_update_rates (PrdTemplates.hpp:18-76): J, I, JRest and the rates of the
    // PRD lines only.  1: hybrid PRD, over the wavelengths that scatter into the PRD grid in THIS column;
    // 2: angle-averaged PRD, over the wavelengths of laMask (those of the redistributed lines).
    extern __shared__ double smem[];
    const int K = P.K, M = P.M, L = P.L, KP = P.KP;
    const int tile = tileList[blockIdx.x];
    const int col = column_of(P, blockIdx.y);
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int lane = lane_id();

    const int slot0 = P.tileSlotOff[tile];
    const int nslot = P.tileSlotOff[tile + 1] - slot0;
    double* acc = smem;                                   // [nslot][4][KP]
    double* scratch = smem + (size_t)P.maxSlots * 4 * KP  // per warp [2][maxNlevel][32]
        + (size_t)warp * 2 * P.maxNlevel * 32;
    double* jbuf = smem + (size_t)P.maxSlots * 4 * KP + (size_t)nwarp * 2 * P.maxNlevel * 32; // [nwarp][KP], split mode

    if (MODE == MODE_ITER)
    {
        for (int idx = threadIdx.x; idx < nslot * 4 * KP; idx += blockDim.x)
            acc[idx] = 0.0;
        __syncthreads();
    }

    Geometry<NCH> g;
    load_geometry<NCH>(g, P.height + (size_t)col * K, K);
    // the formal solvers of the pipeline (lwb200_fsm.cuh), one warp per column
    DepthComm<false> cm{nullptr, 0, 1, 0};
    GeometryR<NCH> gr;
    load_geometry_r<NCH>(cm, gr, P.height + (size_t)col * K, K);
    const double dsTop = fabs(__ldg(P.height + (size_t)col * K) - __ldg(P.height + (size_t)col * K + 1));
    const double dsBot = fabs(__ldg(P.height + (size_t)col * K + K - 2) - __ldg(P.height + (size_t)col * K + K - 1));
    double T[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        T[j] = (k < K) ? __ldg(P.temperature + (size_t)col * K + k) : 1.0;
    }
    const double Ttop0 = __ldg(P.temperature + (size_t)col * K + 0);
    const double Ttop1 = __ldg(P.temperature + (size_t)col * K + 1);
    const double Tbot0 = __ldg(P.temperature + (size_t)col * K + K - 1);
    const double Tbot1 = __ldg(P.temperature + (size_t)col * K + K - 2);

    const int tlBeg = P.tileLa[tile], tlEnd = P.tileLa[tile + 1];
    // A tile with fewer wavelengths than the CTA has warps (one 1D atmosphere: tiles of one wavelength) would leave
    // warps idle behind one warp's serial chain of 2 Nrays rays: the warps then SPLIT THE RAYS of every wavelength
    // of the tile instead (ray r goes to warp r mod nwarp); the partial mean intensities meet in shared memory, the
    // Gamma / rate sums are shared-memory atomics as before.  Every skip below is uniform over the CTA in that mode.
    const bool split = (tlEnd - tlBeg) < nwarp;

    for (int tl = split ? tlBeg : tlBeg + warp; tl < tlEnd; tl += split ? 1 : nwarp)
    {
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

        // --- ray-independent part: background + continua (the reference's
        //     continuaOnly shortcut, :295-307, generalised: continua are
        //     angle-independent at every wavelength)
        double chiC[NCH], etaC[NCH], scaJ[NCH], JDag[NCH], expfac[NCH];
        constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
        const double hc_kl = hc_k / lambda;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = g.k(j);
            const bool v = k < K;
            chiC[j] = v ? __ldg(P.chiBg + rowLK + k) : 1.0;
            etaC[j] = v ? __ldg(P.etaBg + rowLK + k) : 0.0;
            const double sca = v ? __ldg(P.scaBg + rowLK + k) : 0.0;
            JDag[j] = v ? P.J[rowLK + k] : 0.0;
            scaJ[j] = sca * JDag[j];
            expfac[j] = exp(-hc_kl / T[j]);
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
                const int k = g.k(j);
                if (k < K)
                {
                    const UV uv = trans_uv(P, t, col, lt, 0, 0, k, lambda, expfac[j]);
                    const double ni = __ldg(P.n + ((size_t)col * P.NlevTot + t.levI) * K + k);
                    const double nj = __ldg(P.n + ((size_t)col * P.NlevTot + t.levJ) * K + k);
                    chiC[j] += ni * uv.Vij - nj * uv.Vji;
                    etaC[j] += nj * uv.Uji;
                }
            }
        }

        double Jnew[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            Jnew[j] = 0.0;

        for (int mu = 0; mu < M; ++mu)
        {
            const double muz = __ldg(P.muz + mu);
            const double halfwmu = 0.5 * __ldg(P.wmu + mu);
            for (int dir = upOnly ? 1 : 0; dir < 2; ++dir)
            {
                if (split && ((2 * mu + dir) % nwarp) != warp)
                    continue;
                // --- opacity, emissivity, source function for this ray
                double chi[NCH], S[NCH];
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
                                const int k = g.k(j);
                                if (k < K)
                                {
                                    const UV uv = trans_uv(P, t, col, lt, mu, dir, k, lambda, 0.0);
                                    const double ni = __ldg(P.n + ((size_t)col * P.NlevTot + t.levI) * K + k);
                                    const double nj = __ldg(P.n + ((size_t)col * P.NlevTot + t.levJ) * K + k);
                                    chi[j] += ni * uv.Vij - nj * uv.Vji;
                                    eta[j] += nj * uv.Uji;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                        S[j] = (eta[j] + scaJ[j]) / chi[j]; // compute_source_fn (:169-179)
                    if (MODE == MODE_ITER && storeDepth)
                    {
                        const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K;
#pragma unroll
                        for (int j = 0; j < NCH; ++j)
                        {
                            const int k = g.k(j);
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
                double bcB0 = 0.0, bcB1 = 0.0, bcValue = 0.0;
                if (bcType == 2)
                {
                    bcB0 = planck_nu(dir ? Tbot0 : Ttop0, lambda);
                    bcB1 = planck_nu(dir ? Tbot1 : Ttop1, lambda);
                }
                else if (bcType == 4)
                    bcValue = dir ? P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]]
                                  : P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
                double I[NCH], psi[NCH], rchi[NCH];
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    rchi[j] = 1.0 / chi[j];
                if (SOLVER == 2)
                {
                    const double zmu = 1.0 / muz;
                    RayPre<NCH> pre;
                    bezier3_prepare<NCH>(cm, gr, chi, S, muz, zmu, pre);
                    const int kq[4] = {0, 1, K - 2, K - 1};
                    double chiK[4], SK[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        chiK[q] = at_depth<NCH>(chi, kq[q]);
                        SK[q] = at_depth<NCH>(S, kq[q]);
                    }
                    const RayEnds ends = ray_endpoints(chiK, SK, dsTop, dsBot, zmu, dir, bcType, bcB0, bcB1, bcValue);
                    if (dir == 1)
                        bezier3_sweep<NCH, false>(cm, gr, S, rchi, pre, ends, I, psi);
                    else
                        bezier3_sweep<NCH, true>(cm, gr, S, rchi, pre, ends, I, psi);
                }
                else
                    local_stencil_ray<NCH, SOLVER>(cm, gr, chi, S, rchi, muz, dir == 0, bcType, bcB0, bcB1, bcValue, I, psi);

                // spect.I(la, mu, 0) = I(0): the reference stores it after both rays of a mu, the up-going one last
                // (with the rays of a wavelength split over warps the two would race: only that one is stored)
                if (lane == 0 && dir == 1)
                    P.I[((size_t)col * L + la) * M + mu] = I[0];
                store_zplane<NCH>(P, lane, I, dir, ((size_t)col * L + la) * M + mu);

                if (MODE != MODE_ITER)
                    continue;

                if (storeDepth)
                {
                    const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = g.k(j);
                        if (k < K)
                            P.depthI[off + k] = I[j];
                    }
                }

#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    Jnew[j] += halfwmu * I[j]; // accumulate_J (:181-190)

                if (hPrdLa >= 0)
                {
                    // rest-frame mean intensity of hybrid PRD (:397-408): several wavelengths scatter into
                    // one row of JRest, so the sums are fp64 REDs
                    const size_t row = ((((size_t)col * P.NhPrd + hPrdLa) * M + mu) * 2 + dir) * K;
                    double* JRest = P.JRest + (size_t)col * P.NprdLa * K;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = g.k(j);
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

                // --- Gamma and rates, atom by atom (:411-467), one depth chunk
                //     at a time through a per-warp shared scratch that holds
                //     the per-level chi_atom / U_atom of chi_eta_aux_accum (:59-109)
                int e0 = eBeg;
                while (e0 < eEnd)
                {
                    const int atom = P.trans[P.entries[e0].trans].atom;
                    int e1 = e0 + 1;
                    while (e1 < eEnd && P.trans[P.entries[e1].trans].atom == atom)
                        ++e1;
                    const bool detailed = P.atomDetailed[atom] != 0 || prdOnly; // (prdOnly: no operator, no Gamma)
                    const int N = P.atomNlevel[atom];
                    double* Xs = scratch;                      // chi_atom[level][lane]
                    double* Us = scratch + P.maxNlevel * 32;   // U_atom[level][lane]
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
                    {
                        const int k = g.k(j);
                        if (k < K)
                        {
                            double Ieff = I[j];
                            if (!detailed)
                            {
                                for (int m = 0; m < N; ++m)
                                {
                                    Xs[m * 32 + lane] = 0.0;
                                    Us[m * 32 + lane] = 0.0;
                                }
                                double etaA = 0.0;
                                for (int e = e0; e < e1; ++e)
                                {
                                    const DevTrans& t = P.trans[P.entries[e].trans];
                                    const UV uv = trans_uv(P, t, col, la - t.Nblue, mu, dir, k, lambda, expfac[j]);
                                    const double ni = __ldg(P.n + ((size_t)col * P.NlevTot + t.levI) * K + k);
                                    const double nj = __ldg(P.n + ((size_t)col * P.NlevTot + t.levJ) * K + k);
                                    const double x = ni * uv.Vij - nj * uv.Vji;
                                    Xs[t.i * 32 + lane] += x;
                                    Xs[t.j * 32 + lane] -= x;
                                    Us[t.j * 32 + lane] += uv.Uji;
                                    etaA += nj * uv.Uji;
                                }
                                Ieff = I[j] - psi[j] * etaA; // compute_full_Ieff (:192-204)
                            }
                            for (int e = e0; e < e1; ++e)
                            {
                                const DevEntry en = P.entries[e];
                                const DevTrans& t = P.trans[en.trans];
                                if (prdOnly && t.rhoOff < 0)
                                    continue; // rates of the PRD lines only (:433-434, :455-456)
                                const int lt = la - t.Nblue;
                                const UV uv = trans_uv(P, t, col, lt, mu, dir, k, lambda, expfac[j]);
                                const double wlamu = trans_wla(P, t, col, lt, k, lambda) * halfwmu;
                                double* a4 = acc + (size_t)en.slot * 4 * KP + k;
                                if (!detailed)
                                {
                                    // compute_full_operator_rates (:218-226)
                                    double integrand = (uv.Uji + uv.Vji * Ieff)
                                        - (psi[j] * Xs[t.i * 32 + lane] * Us[t.j * 32 + lane]);
                                    smem_add(a4, integrand * wlamu);
                                    integrand = (uv.Vij * Ieff)
                                        - (psi[j] * Xs[t.j * 32 + lane] * Us[t.i * 32 + lane]);
                                    smem_add(a4 + KP, integrand * wlamu);
                                }
                                smem_add(a4 + 2 * KP, I[j] * uv.Vij * wlamu);              // Rij (:230)
                                smem_add(a4 + 3 * KP, (uv.Uji + I[j] * uv.Vji) * wlamu);   // Rji (:231)
                            }
                        }
                    }
                    e0 = e1;
                }
            }
        }

        if (MODE == MODE_ITER && split)
        {
            // the warps' partial sums of J, added in warp order by the first
            __syncthreads();
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = g.k(j);
                if (k < K)
                    jbuf[warp * KP + k] = Jnew[j];
            }
            __syncthreads();
            if (warp == 0)
            {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = g.k(j);
                    double sJ = 0.0;
                    if (k < K)
                        for (int w2 = 0; w2 < nwarp; ++w2)
                            sJ += jbuf[w2 * KP + k];
                    Jnew[j] = sJ;
                }
            }
        }
        if (MODE == MODE_ITER && (!split || warp == 0))
        {
            // J row and dJ = max_k |1 - Jdag/J|  (:477-485)
            double dJ = 0.0;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = g.k(j);
                if (k < K)
                {
                    P.J[rowLK + k] = Jnew[j];
                    const double d = fabs(1.0 - JDag[j] / Jnew[j]);
                    dJ = (d < dJ) ? dJ : d;
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

    if (MODE == MODE_ITER)
    {
        __syncthreads();
        // flush this tile's partial sums (one fp64 RED per element per tile)
        for (int idx = threadIdx.x; idx < nslot * 4 * KP; idx += blockDim.x)
        {
            const int k = idx % KP;
            const int q = (idx / KP) & 3;
            const int s = idx / (4 * KP);
            if (k >= K)
                continue;
            const DevTrans& t = P.trans[P.tileSlotTrans[slot0 + s]];
            const int row = q == 0 ? t.accIJ : q == 1 ? t.accJI : q == 2 ? t.accRij : t.accRji;
            if (row < 0)
                continue;
            const double v = acc[idx];
            if (v != 0.0)
                atomicAdd(P.accum + ((size_t)col * P.AccTot + row) * K + k, v);
        }
    }
}


} // namespace lwb200
