// lwb200_gamma.cuh -- stage 3 of the pipeline, second generation: Gamma and Rij/Rji of every
// active transition from J and the ray moments (compute_full_Ieff / compute_full_operator_rates,
// SimdFullIterationTemplates.hpp:192-234, in the moment form described in lwb200_pipeline.cuh).
//
// One WARP owns a (wavelength tile, column, depth chunk): its 32 lanes are laid over depth, lane l
// holding the NC depths k = chunk + c*32 + l (cyclic, so every global row and every shared-memory row
// is read as 32 consecutive doubles).  What the first generation (gamma_kernel, one thread per depth)
// paid per thread is paid here once per warp and NC depths:
//   * the active-transition list of a wavelength is a table of compact 32-byte GEntry records, brought
//     into shared memory by the warp one wavelength AHEAD of its use (lane e copies entry e); the
//     loops read them as warp-uniform shared-memory loads -- no dependent global loads and no 64-bit
//     address arithmetic per transition;
//   * per-wavelength scalars (hc/(k lambda), 2hc/lambda^3, the line constants) are made by the host
//     planner once (GLam, GLine) instead of by every thread of every column, 1/T once per tile;
//   * the continuum aggregates per level are written "first writer stores" (flags set by the planner),
//     so nothing is zeroed per wavelength;
//   * the NC depths of a lane are independent instruction streams, which is what hides the latency of
//     the shared-memory accumulators at the few resident warps their size allows.
//   * the warps of a CTA work on different tiles of the SAME column, whose populations and continuum
//     ratios n*_i / n*_j are staged once per CTA in shared memory (STAGE): the inner loops then touch
//     shared memory only, with 32-bit addresses (the first generation spent four integer instructions
//     of 64-bit address arithmetic on every global load and missed the L1 four times out of five).
// The tile's per-transition partial sums live in shared memory, one writer per element (the lane that
// owns the depth): no barrier, no shared-memory atomic; one fp64 RED per element per tile at the end.
#pragma once
#include "lwb200_pipeline.cuh"

namespace lwb200
{
// One active transition at one wavelength.
struct GEntry
{
    double al;              // continuum: alpha(lambda); line: 0
    double wlaF;            // continuum: wlambda / lambda * 4 pi / h;  line: unused (GLine::wlaS)
    short accRow;           // first of the transition's 4 accumulator rows within the tile (slot * 4)
    short levI, levJ;       // rows in the packed population arrays
    short cont;             // continuum index (gRatio row); -1 for lines
    unsigned short flags;   // GE_*
    unsigned char type;     // 0 line, 1 continuum
    unsigned char i, j;     // levels within the atom
    unsigned char lslot;    // line: its slot (0..2) at this wavelength
    unsigned char groupLen; // on the first entry of an atom group: number of entries in the group
    unsigned char atomTag;  // small per-wavelength tag of the atom (to match line slots to groups)
};
enum
{
    GE_PRD = 1,        // line with rhoPrd
    GE_DETAILED = 2,   // atom is detailed-static: rates only
    GE_STORE_XI = 4,   // continuum: first writer of Xs[i] / Xs[j] / Us[j] in its group (store, else add)
    GE_STORE_XJ = 8,
    GE_STORE_UJ = 16,
    GE_XI_VALID = 32,  // a continuum of the group wrote Xs[i] / Xs[j] / Us[i] / Us[j] (else that aggregate is 0)
    GE_XJ_VALID = 64,
    GE_UI_VALID = 128,
    GE_UJ_VALID = 256
};
static_assert(sizeof(GEntry) == 32, "GEntry is copied as two 16-byte words");

// Per wavelength of a tile, in tile order.
struct GLam
{
    double hc_kl;   // hc / (k_B lambda)
    double hcl;     // 2 hc / lambda^3
    int la;         // wavelength index
    int eOff;       // first GEntry
    int eCnt;
    int momRow;     // first moment row
    int nLines;     // overlapping lines (0..3)
    int pad;
};

// Per (wavelength, line slot).
struct GLine
{
    double vB;      // hc/4pi * lambda0/lambda * Bij        (Transition::uv, LwTransition.hpp:93-130)
    double gS;      // Bji / Bij
    double AB;      // Aji / Bji
    double wlaS;    // wlambda * 4 pi / (h c)  (times wphi(k) gives wla)
    long long rhoOff;       // element offset of rhoPrd(lt, 0) of column 0; -1: none
    long long rhoColStride;
    int wphiRow;    // line index: wphi row
    short levI, levJ;
    unsigned char i, j, atomTag, pad;
    int pad2;
};

struct GammaPlan
{
    const GEntry* entries;
    const GLam* lam;          // [sum over tiles of their wavelengths], tile order
    const GLine* line;        // [same][3]
    const int* tileLa;        // [Ntile + 1] offsets into lam
    const int* lamOfLa;       // [L] position of a wavelength in `lam` (-1: not a moment wavelength)
    const int* tileSlotOff;   // [Ntile + 1]
    const int4* tileSlotRows; // per slot: rows of the packed accumulator for Gamma(i,j), Gamma(j,i), Rij, Rji (-1: none)
    int maxSlots, maxNlevel;
};

constexpr int kGammaMaxEntries = 32; // active transitions at one wavelength (one per lane of the staging copy)

// shared memory of one warp (its accumulators, the per-level aggregates, the entry staging buffers) and of
// the column staging area the warps of a CTA share (populations and continuum ratios of their column)
inline size_t gamma_warp_smem(int NC, int maxSlots, int maxNlevel)
{
    return ((size_t)maxSlots * 4 + 2 * (size_t)maxNlevel) * 32 * NC * sizeof(double)
           + 2 * kGammaMaxEntries * sizeof(GEntry);
}
inline size_t gamma_stage_smem(int NC, int NlevTot, int Ncont)
{
    return ((size_t)NlevTot + Ncont) * 32 * NC * sizeof(double);
}

// Everything of one wavelength at NCP of the depths of this lane (chunks cBeg .. cBeg + NCP - 1).
// acc / Xs / Us / nS point at this lane's element of row 0 of chunk 0; rows are RS doubles apart, the
// lane's chunks 32 apart within a row.  STAGE: nS holds the column's population rows followed by its
// gRatio rows; otherwise they are read from global memory.
template <int NL, int NCP, int RS, bool STAGE>
__device__ __forceinline__ void gamma_lambda_w(const DevProblem& P, const GLam& gl, const GLine* __restrict__ gline,
                                               const GEntry* __restrict__ ents, int col, int cb, int cBeg, int kLane,
                                               const double* __restrict__ rT, double W0, double* __restrict__ acc,
                                               double* __restrict__ Xs, double* __restrict__ Us,
                                               const double* __restrict__ nS, const bool prdOnly)
{
    constexpr int NLA = NL > 0 ? NL : 1;
    constexpr int NQ = NL + 1;
    const int K = P.K;
    const double* gS = nS + (size_t)P.NlevTot * RS; // staged gRatio rows
    // global fallbacks: this lane's first depth of row 0
    const int kSafe = kLane < K ? kLane : 0;
    const double* nG = P.n + (size_t)col * P.NlevTot * K + kSafe;
    const double* gG = P.gRatio + (size_t)col * K + kSafe;
    const size_t gStride = (size_t)P.Ncol * K;

    int kk[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c)
    {
        const int k = kLane + (cBeg + c) * 32;
        kk[c] = k < K ? k : K - 1; // lanes past the end redo the last depth; their results are never flushed
    }
    // population of packed level `row` / gRatio of continuum `row` at chunk c of this lane
    auto pop = [&](int row, int c) -> double {
        return STAGE ? nS[row * RS + (cBeg + c) * 32] : __ldg(nG + (size_t)row * K + (kk[c] - kSafe));
    };
    auto ratio = [&](int row, int c) -> double {
        return STAGE ? gS[row * RS + (cBeg + c) * 32] : __ldg(gG + (size_t)row * gStride + (kk[c] - kSafe));
    };

    // ---- moments of this wavelength (independent loads, issued together)
    const double* mom = P.mom + ((size_t)cb * P.momRows + gl.momRow) * K;
    const double* Jrow = P.J + ((size_t)col * P.L + gl.la) * K;
    double Mq[NCP][NQ][NQ], mW[NCP][NLA], mA[NCP][NLA], mJ[NCP], expfac[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c)
    {
        const int k = kk[c];
        mJ[c] = Jrow[k];
        Mq[c][0][0] = mom[k];
        mW[c][0] = mA[c][0] = 0.0;
        int pr = 0;
#pragma unroll
        for (int a = 0; a < NL; ++a)
        {
            mW[c][a] = mom[(size_t)(1 + 4 * a) * K + k];
            mA[c][a] = mom[(size_t)(2 + 4 * a) * K + k];
            Mq[c][0][a + 1] = Mq[c][a + 1][0] = mom[(size_t)(3 + 4 * a) * K + k];
            Mq[c][a + 1][a + 1] = mom[(size_t)(4 + 4 * a) * K + k];
#pragma unroll
            for (int b = a + 1; b < NL; ++b)
            {
                Mq[c][a + 1][b + 1] = Mq[c][b + 1][a + 1] = mom[(size_t)(1 + 4 * NL + pr) * K + k];
                ++pr;
            }
        }
        expfac[c] = exp_fast_underflow(-gl.hc_kl * rT[cBeg + c]);
    }
    const double hcl = gl.hcl;

    // ---- line slots of this wavelength
    int lsAtom[NLA], lsI[NLA], lsJ[NLA];
    double lsV[NLA];
    double lsGv[NCP][NLA], lsUgv[NCP][NLA], lsX[NCP][NLA], lsE[NCP][NLA], lsWla[NCP][NLA];
#pragma unroll
    for (int l = 0; l < NLA; ++l)
    {
        lsAtom[l] = -1;
        lsI[l] = lsJ[l] = 0;
        lsV[l] = 0.0;
#pragma unroll
        for (int c = 0; c < NCP; ++c)
            lsGv[c][l] = lsUgv[c][l] = lsX[c][l] = lsE[c][l] = lsWla[c][l] = 0.0;
        if (NL > 0)
        {
            const GLine ll = gline[l];
            lsAtom[l] = ll.atomTag;
            lsI[l] = ll.i;
            lsJ[l] = ll.j;
            lsV[l] = ll.vB;
            const double gSvB = ll.gS * ll.vB;
            const double ABgSvB = ll.AB * gSvB;
            const double* wphi = P.wphi + ((size_t)ll.wphiRow * P.Ncol + col) * K;
#pragma unroll
            for (int c = 0; c < NCP; ++c)
            {
                const int k = kk[c];
                const double r = (ll.rhoOff >= 0) ? __ldg(P.rhoPrd + ll.rhoOff + (size_t)col * ll.rhoColStride + k) : 1.0;
                const double gk = (ll.rhoOff >= 0) ? ll.gS * r : ll.gS;
                const double ni = pop(ll.levI, c);
                const double nj = pop(ll.levJ, c);
                lsGv[c][l] = gSvB * r;
                lsUgv[c][l] = ABgSvB * r;
                lsX[c][l] = ll.vB * (ni - nj * gk);
                lsE[c][l] = nj * (ll.AB * (gk * ll.vB));
                lsWla[c][l] = ll.wlaS * __ldg(wphi + k);
            }
        }
    }

    int e0 = 0;
    while (e0 < gl.eCnt)
    {
        const GEntry first = ents[e0];
        const int e1 = e0 + first.groupLen;
        const bool detailed = (first.flags & GE_DETAILED) != 0;
        const int atom = first.atomTag;
        double E0[NCP];
#pragma unroll
        for (int c = 0; c < NCP; ++c)
            E0[c] = 0.0;
        if (!detailed && !prdOnly)
        {
            // continuum aggregates per level: chi_atom / U_atom of chi_eta_aux_accum (:59-109)
            for (int e = e0; e < e1; ++e)
            {
                const GEntry t = ents[e];
                if (t.type == 0)
                    continue;
                const double al = t.al;
                double* xi = Xs + t.i * RS;
                double* xj = Xs + t.j * RS;
                double* uj = Us + t.j * RS;
                const bool sXi = t.flags & GE_STORE_XI, sXj = t.flags & GE_STORE_XJ, sUj = t.flags & GE_STORE_UJ;
#pragma unroll
                for (int c = 0; c < NCP; ++c)
                {
                    const int o = (cBeg + c) * 32;
                    const double gk = ratio(t.cont, c) * expfac[c];
                    const double Vji = gk * al;
                    const double Uji = hcl * Vji;
                    const double ni = pop(t.levI, c);
                    const double nj = pop(t.levJ, c);
                    const double x = ni * al - nj * Vji;
                    xi[o] = sXi ? x : xi[o] + x;
                    xj[o] = sXj ? -x : xj[o] - x;
                    uj[o] = sUj ? Uji : uj[o] + Uji;
                    E0[c] += nj * Uji;
                }
            }
        }
        // line members of this atom: per-unit-phi coefficients (0 for other atoms' lines)
        bool own[NLA];
#pragma unroll
        for (int l = 0; l < NLA; ++l)
            own[l] = NL > 0 && lsAtom[l] == atom;
        // EB[q] = sum_q' E_q' M(q, q')
        double EB[NCP][NQ];
#pragma unroll
        for (int c = 0; c < NCP; ++c)
        {
            double Eq[NQ];
            Eq[0] = E0[c];
#pragma unroll
            for (int l = 0; l < NL; ++l)
                Eq[l + 1] = own[l] ? lsE[c][l] : 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
            {
                double s = 0.0;
#pragma unroll
                for (int q2 = 0; q2 < NQ; ++q2)
                    s = fma(Eq[q2], Mq[c][q][q2], s);
                EB[c][q] = s;
            }
        }

        for (int e = e0; e < e1; ++e)
        {
            const GEntry t = ents[e];
            if (prdOnly && !(t.flags & GE_PRD))
                continue; // formal_sol_prd_update_rates: rates of the PRD lines only (:433-434)
            const bool isCont = t.type != 0;
            const int ls = t.lslot;
            const int contRow = isCont ? t.cont : 0;
            // which aggregates exist, and how this transition's levels meet the line slots (warp-uniform)
            const bool xiV = t.flags & GE_XI_VALID, xjV = t.flags & GE_XJ_VALID;
            const bool uiV = t.flags & GE_UI_VALID, ujV = t.flags & GE_UJ_VALID;
            double sXi[NLA], sXj[NLA], bUi[NLA], bUj[NLA];
#pragma unroll
            for (int l = 0; l < NLA; ++l)
            {
                const bool o = own[l];
                sXi[l] = !o ? 0.0 : (t.i == lsI[l] ? 1.0 : (t.i == lsJ[l] ? -1.0 : 0.0));
                sXj[l] = !o ? 0.0 : (t.j == lsI[l] ? 1.0 : (t.j == lsJ[l] ? -1.0 : 0.0));
                bUi[l] = (o && t.i == lsJ[l]) ? 1.0 : 0.0;
                bUj[l] = (o && t.j == lsJ[l]) ? 1.0 : 0.0;
            }
            double* a4 = acc + t.accRow * RS;
            const double* xiP = Xs + t.i * RS;
            const double* xjP = Xs + t.j * RS;
            const double* uiP = Us + t.i * RS;
            const double* ujP = Us + t.j * RS;
#pragma unroll
            for (int c = 0; c < NCP; ++c)
            {
                const int o = (cBeg + c) * 32;
                double v, gv, ugv, Wq, Aq, EBq, wla;
                if (isCont)
                {
                    const double gk = ratio(contRow, c) * expfac[c];
                    v = t.al;
                    gv = gk * t.al;
                    ugv = hcl * gv;
                    Wq = W0;
                    Aq = mJ[c];
                    EBq = EB[c][0];
                    wla = t.wlaF;
                }
                else
                {
                    v = lsV[0];
                    gv = lsGv[c][0];
                    ugv = lsUgv[c][0];
                    Wq = mW[c][0];
                    Aq = mA[c][0];
                    EBq = EB[c][NL > 0 ? 1 : 0];
                    wla = lsWla[c][0];
#pragma unroll
                    for (int l = 1; l < NL; ++l)
                        if (ls == l)
                        {
                            v = lsV[l];
                            gv = lsGv[c][l];
                            ugv = lsUgv[c][l];
                            Wq = mW[c][l];
                            Aq = mA[c][l];
                            EBq = EB[c][l + 1];
                            wla = lsWla[c][l];
                        }
                }
                if (!detailed && !prdOnly)
                {
                    // chi_atom(m) = sum_q p_q X_q(m), U_atom(m) = sum_q p_q U_q(m)
                    double Xi[NQ], Xj[NQ], Ui[NQ], Uj[NQ];
                    Xi[0] = xiV ? xiP[o] : 0.0;
                    Xj[0] = xjV ? xjP[o] : 0.0;
                    Ui[0] = uiV ? uiP[o] : 0.0;
                    Uj[0] = ujV ? ujP[o] : 0.0;
#pragma unroll
                    for (int l = 0; l < NL; ++l)
                    {
                        Xi[l + 1] = sXi[l] * lsX[c][l];
                        Xj[l + 1] = sXj[l] * lsX[c][l];
                        Ui[l + 1] = bUi[l] * lsUgv[c][l];
                        Uj[l + 1] = bUj[l] * lsUgv[c][l];
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
                            mj = fma(Mq[c][q][q2], Uj[q2], mj);
                            mi = fma(Mq[c][q][q2], Ui[q2], mi);
                        }
                        XUij = fma(Xi[q], mj, XUij);
                        XUji = fma(Xj[q], mi, XUji);
                    }
                    // sum_r w [(Uji + Vji Ieff) - Psi* chi(i) U(j)],  Ieff = I - Psi* eta_atom
                    a4[o] += (ugv * Wq + gv * (Aq - EBq) - XUij) * wla;
                    a4[RS + o] += (v * (Aq - EBq) - XUji) * wla;
                }
                a4[2 * RS + o] += (v * Aq) * wla;
                a4[3 * RS + o] += (ugv * Wq + gv * Aq) * wla;
            }
        }
        e0 = e1;
    }
}

// laMask / prdOnly: the rates-only pass of the PRD sub-iterations over the masked wavelengths.
// NC: depths per lane (the chunk of a warp is 32 * NC depths; blockIdx.z walks the chunks of a column).
// A CTA is blockDim.x / 32 warps on consecutive tiles of tileList, all in column blockIdx.y.
#ifndef LWB200_GTILE_MINB
#define LWB200_GTILE_MINB 20
#endif
template <int NC, bool STAGE>
#ifndef LWB200_GSTAGE_THREADS
#define LWB200_GSTAGE_THREADS 256
#endif
__global__ void __launch_bounds__(STAGE ? LWB200_GSTAGE_THREADS : 32,
                                  STAGE ? (LWB200_GSTAGE_THREADS == 32 ? 16 : 1) : (NC == 1 ? LWB200_GTILE_MINB : 8))
gamma_tile_kernel(const DevProblem P, const GammaPlan G, const int* __restrict__ tileList, int nTiles, int laLo,
                  int laHi, int colBase, const unsigned char* __restrict__ laMask, int prdOnly, int warpSmemDoubles)
{
    extern __shared__ double smem[];
    constexpr int RS = 32 * NC;
    const int K = P.K;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int nwarp = blockDim.x >> 5;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int kLane = blockIdx.z * RS + lane;

    // ---- the column's populations and continuum ratios, once per CTA
    const int stageRows = STAGE ? P.NlevTot + P.Ncont : 0;
    if (STAGE)
    {
        const double* ncol = P.n + (size_t)col * P.NlevTot * K;
        const double* gcol = P.gRatio + (size_t)col * K;
        const size_t gStride = (size_t)P.Ncol * K;
        const int kChunk = blockIdx.z * RS;
        for (int idx = threadIdx.x; idx < stageRows * RS; idx += blockDim.x)
        {
            const int row = idx / RS, o = idx - row * RS;
            const int k = kChunk + o;
            double v = 0.0;
            if (k < K)
                v = row < P.NlevTot ? __ldg(ncol + (size_t)row * K + k) : __ldg(gcol + (size_t)(row - P.NlevTot) * gStride + k);
            smem[idx] = v;
        }
        __syncthreads();
    }
    const int tileIdx = blockIdx.x * nwarp + warp;
    if (tileIdx >= nTiles)
        return;
    const int tile = tileList[tileIdx];
    const double* nS = smem + lane;
    double* mine = smem + (size_t)stageRows * RS + (size_t)warp * warpSmemDoubles;
    const int slot0 = G.tileSlotOff[tile];
    const int nslot = G.tileSlotOff[tile + 1] - slot0;
    double* acc = mine + lane;                                // [nslot][4][RS]  partial sums, one writer each
    double* Xs = mine + (size_t)G.maxSlots * 4 * RS + lane;   // [maxNlevel][RS]
    double* Us = Xs + (size_t)G.maxNlevel * RS;
    GEntry* stage = reinterpret_cast<GEntry*>(mine + ((size_t)G.maxSlots * 4 + 2 * (size_t)G.maxNlevel) * RS);

    for (int r = 0; r < nslot * 4; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c)
            acc[r * RS + c * 32] = 0.0;
    double rT[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c)
    {
        const int k = kLane + c * 32;
        rT[c] = 1.0 / __ldg(P.temperature + (size_t)col * K + (k < K ? k : K - 1));
    }
    // W0 = sum_r w over both directions of every mu, in the ray order of ray_kernel
    double W0 = 0.0;
    for (int mu = 0; mu < P.M; ++mu)
    {
        const double w = 0.5 * __ldg(P.wmu + mu);
        W0 += w;
        W0 += w;
    }

    const int tlBeg = G.tileLa[tile], tlEnd = G.tileLa[tile + 1];
    // software pipeline over the wavelengths of the tile: header two ahead, entries one ahead
    const uint4* entWords = reinterpret_cast<const uint4*>(G.entries);
    GLam hdr = G.lam[tlBeg];
    GLam hdrNext = G.lam[min(tlBeg + 1, tlEnd - 1)];
    {
        if (lane < hdr.eCnt)
        {
            const uint4* src = entWords + 2 * (size_t)(hdr.eOff + lane);
            uint4* dst = reinterpret_cast<uint4*>(stage + lane);
            dst[0] = src[0];
            dst[1] = src[1];
        }
        __syncwarp();
    }
    int buf = 0;
    for (int tl = tlBeg; tl < tlEnd; ++tl)
    {
        // issue the loads of the next wavelength's entries and of the header after it
        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
        const bool more = tl + 1 < tlEnd;
        if (more && lane < hdrNext.eCnt)
        {
            const uint4* src = entWords + 2 * (size_t)(hdrNext.eOff + lane);
            w0 = src[0];
            w1 = src[1];
        }
        const GLam hdrNext2 = G.lam[min(tl + 2, tlEnd - 1)];
        const int la = hdr.la;
        if (la >= laLo && la < laHi && !(laMask && !laMask[la]))
        {
            const GEntry* ents = stage + buf * kGammaMaxEntries;
            const GLine* gline = G.line + (size_t)tl * 3;
            const bool po = prdOnly != 0;
            switch (hdr.nLines)
            {
            case 0: gamma_lambda_w<0, NC, RS, STAGE>(P, hdr, gline, ents, col, cb, 0, kLane, rT, W0, acc, Xs, Us, nS, po); break;
            case 1: gamma_lambda_w<1, NC, RS, STAGE>(P, hdr, gline, ents, col, cb, 0, kLane, rT, W0, acc, Xs, Us, nS, po); break;
            case 2:
                // two / three overlapping lines: one depth chunk at a time (registers)
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    gamma_lambda_w<2, 1, RS, STAGE>(P, hdr, gline, ents, col, cb, c, kLane, rT, W0, acc, Xs, Us, nS, po);
                break;
            case 3:
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    gamma_lambda_w<3, 1, RS, STAGE>(P, hdr, gline, ents, col, cb, c, kLane, rT, W0, acc, Xs, Us, nS, po);
                break;
            default: break; // > 3 overlapping lines: handled by the general kernel
            }
        }
        // rotate the pipeline: the prefetched entries become current
        buf ^= 1;
        if (more)
        {
            uint4* dst = reinterpret_cast<uint4*>(stage + buf * kGammaMaxEntries + lane);
            dst[0] = w0;
            dst[1] = w1;
        }
        __syncwarp();
        hdr = hdrNext;
        hdrNext = hdrNext2;
    }
    // flush: this lane's own elements of the partial sums (one fp64 RED per element per tile)
    for (int s = 0; s < nslot; ++s)
    {
        const int4 rows4 = G.tileSlotRows[slot0 + s];
        const int rows[4] = {rows4.x, rows4.y, rows4.z, rows4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            if (rows[q] < 0)
                continue;
#pragma unroll
            for (int c = 0; c < NC; ++c)
            {
                const int k = kLane + c * 32;
                const double v = acc[(s * 4 + q) * RS + c * 32];
                if (k < K && v != 0.0)
                    atomicAdd(P.accum + ((size_t)col * P.AccTot + rows[q]) * K + k, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Stage 1 with the planner's tables: chiC, etaC of the wavelengths of one kind (continuum_kernel,
// lwb200_pipeline.cuh, computes the per-wavelength scalars -- two divisions -- and decodes 80-byte
// DevEntry records in every thread; here they come from GLam / GEntry and 1/T is taken once per thread).
// Grid (groups of `perBlock` wavelengths of the list, columns of the batch), one thread per depth.
__global__ void continuum_table_kernel(const DevProblem P, const GammaPlan G, const int* __restrict__ lamList, int nLam,
                                       int perBlock, int colBase)
{
    const int K = P.K, L = P.L;
    const int cb = blockIdx.y, col = column_of(P, colBase + cb);
    const int k = threadIdx.x;
    if (k >= K)
        return;
    const double rT = 1.0 / __ldg(P.temperature + (size_t)col * K + k);
    const double* ncol = P.n + (size_t)col * P.NlevTot * K + k;
    const double* gcol = P.gRatio + (size_t)col * K + k;
    const size_t gStride = (size_t)P.Ncol * K;
    const int qBeg = blockIdx.x * perBlock, qEnd = min(nLam, qBeg + perBlock);
    for (int q = qBeg; q < qEnd; ++q)
    {
        const int la = lamList[q];
        const GLam gl = G.lam[G.lamOfLa[la]];
        const size_t rowLK = ((size_t)col * L + la) * K + k;
        double chiC = __ldg(P.chiBg + rowLK);
        double etaC = __ldg(P.etaBg + rowLK);
        const double expfac = exp_fast_underflow(-gl.hc_kl * rT);
        const GEntry* ents = G.entries + gl.eOff;
        for (int e = 0; e < gl.eCnt; ++e)
        {
            const GEntry t = ents[e];
            if (t.type == 0)
                continue;
            const double gk = __ldg(gcol + (size_t)t.cont * gStride) * expfac;
            const double Vji = gk * t.al;
            const double ni = __ldg(ncol + (size_t)t.levI * K);
            const double nj = __ldg(ncol + (size_t)t.levJ * K);
            chiC += ni * t.al - nj * Vji;
            etaC += nj * (gl.hcl * Vji);
        }
        const size_t o = ((size_t)cb * L + la) * K + k;
        P.chiC[o] = chiC;
        P.etaC[o] = etaC;
    }
}

} // namespace lwb200
