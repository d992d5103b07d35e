// lwb200_api.cu -- implementation of the C-ABI in include/lwb200.h: problem
// marshalling, the wavelength work plan (what TaskScheduler/ThreadStorage did on
// the CPU), device memory, host<->device copies and kernel launches.
#include "../../include/lwb200.h"
#include "lwb200_kernels.cuh"
#include "lwb200_profiles.cuh"
#include "lwb200_pipeline.cuh"
#include "lwb200_gamma.cuh"
#include "lwb200_ray2.cuh"
#include "lwb200_fslong.cuh"
#include "lwb200_fsgeneral.cuh"
#include "lwb200_prd.cuh"
#include "lwb200_stokes.cuh"
#include "lwb200_ng.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <thread>
#include <vector>

using namespace lwb200;

namespace
{
thread_local std::string g_err;
// kernels launched by this library in this process, over all contexts (lwb200_global_launch_count)
std::atomic<long long> g_totalLaunches{0};

// per-call launch counter of a context that also feeds the process-wide count
struct LaunchCounter
{
    int64_t v = 0;
    LaunchCounter& operator+=(int64_t n)
    {
        v += n;
        g_totalLaunches += n;
        return *this;
    }
    LaunchCounter& operator=(int64_t n)
    {
        v = n;
        return *this;
    }
    operator int64_t() const { return v; }
};

int fail(const std::string& msg)
{
    g_err = msg;
    return 1;
}

#define CU(call)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)

struct HostTrans
{
    LwB200Transition t;
    int atom, kr, global;
};

template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count)
    {
        n = count;
        if (count == 0)
            return 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess)
        {
            g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
            return 1;
        }
        return 0;
    }
    int upload(const std::vector<T>& h)
    {
        if (alloc(h.size()))
            return 1;
        if (h.empty())
            return 0;
        cudaError_t e = cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
        {
            g_err = std::string("cudaMemcpy: ") + cudaGetErrorString(e);
            return 1;
        }
        return 0;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
};
} // namespace

struct Pinned
{
    double* p = nullptr;
    size_t n = 0;
    cudaEvent_t ev = nullptr;
    bool inFlight = false;
    void release()
    {
        if (p)
            cudaFreeHost(p);
        if (ev)
            cudaEventDestroy(ev);
        p = nullptr;
        ev = nullptr;
        n = 0;
    }
};

struct Pending
{
    void* dst;
    const void* src;
    size_t bytes;
};

// Host-side pack / scatter of the per-atom arrays of a column stack: a few threads once the
// copy is large enough to be worth them (one 1D atmosphere never is).
template <typename F>
static void host_parallel_for(size_t n, size_t bytes, F body)
{
    unsigned T = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (bytes < ((size_t)4 << 20) || n < 2 * T || T == 1)
    {
        body((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = (n + T - 1) / T;
    for (unsigned t = 1; t < T; ++t)
        if (t * per < n)
            pool.emplace_back([=] { body(t * per, std::min(n, (t + 1) * per)); });
    body((size_t)0, std::min(n, per));
    for (auto& th : pool)
        th.join();
}

// What one pass of the pipeline covers: every wavelength of the context's range (the Gamma
// iteration) or the wavelengths touched by the redistributed PRD lines (rates-only pass).
struct PipelineLists
{
    const int* moment;
    int nMoment;
    const int* gTiles; // tiles of the second-generation Gamma stage
    int nGTiles;
    const int* kindLam[4];
    int nKindLam[4];
    const unsigned char* laMask;
    int prdOnly;
    const int* polLam; // full Stokes pass: wavelengths with a polarised line (stokes_kernel)
    int nPolLam;
    int fullRange;     // the lists cover the whole spectrum whatever the context's wavelength shard
    int directPrdOnly; // hybrid PRD: the pass is the general kernel in PRD-rates-only mode over the direct tiles
    const int* directPrd; // angle-averaged PRD: general-kernel tiles holding redistributed wavelengths (laMask selects them)
    int nDirectPrd;
};

struct LwB200Context
{
    // results of the *_async calls, in pinned host memory: valid after the next lwb200_sync
    struct HostScalars
    {
        double dJ;
        long long dJIdx;
        int nSingular;
    }* hs = nullptr, *hsDev = nullptr;
    bool customLists = false; // launch_fs uses prdPl instead of the context-wide lists (PRD and Stokes passes)
    PipelineLists prdPl{};
    int stokesFsMode = 0; // != 0: the pass is a full-Stokes formal solution (always Bezier3)
    Pinned stN, stNStar, stNTotal, stVBroad, stPrefill, stGamma, stNOut, stGammaOut, stRates;
    std::vector<Pending> pending;
    std::vector<void*> registered;
    cudaEvent_t evK0 = nullptr, evK1 = nullptr;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evRays = nullptr, evCopy = nullptr;
    bool fetchEarly = false, fetched = false, outputsPinned = false;
    bool finaliseEarly = false;    // column stacks: finalise + send Gamma and the rates home batch by batch
    bool finalisedEarly = false, fetchedGR = false;
    bool ioPinned = false; // populations, Gamma, rates, nStar ... are pinned in place: strided DMA, no staging
    cudaStream_t sideStream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evFork = nullptr, evJoin[4] = {nullptr, nullptr, nullptr, nullptr};
    bool kernelTimed = false;
    bool forceDirect = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    LwB200Problem prob{};
    std::vector<LwB200Atom> atoms;
    std::vector<std::vector<LwB200Transition>> atomTrans;
    std::vector<HostTrans> trans; // flattened, global order
    std::vector<DevTrans> devTrans;
    std::vector<int> atomNlevel, atomLevOff, atomGammaOff, atomDetailed;
    std::vector<int> tileLa, tileLambda, tileSlotOff, tileSlotTrans, tileKind;
    // pipeline work lists: wavelengths per kind (0..3 overlapping lines) for ray_kernel, tiles
    // for continuum_kernel / gamma_kernel (moment tiles) and for the general fs_kernel
    std::vector<int> laKind;
    DevBuf<int> dKindLam[4], dListMoment, dListDirect, dListAll, dMomOff, dLaNLines;
    DevBuf<LambdaLine> dLamLine;
    DevBuf<double> chiC, etaC, mom;
    int nKindLam[4] = {0, 0, 0, 0}, nListMoment = 0, nListDirect = 0, nListAll = 0, listLo = -1, listHi = -1;
    int batchCols = 1, momRows = 0;
    bool batchTaper = true;  // LWB200_BATCH_TAPER=0: uniform batches (tuning aid)
    int phiFlagsHost = 3;    // host copy of the profile flags (deep atmospheres only; 3 = assume nothing)
    // angle-averaged PRD (lwb200_redistribute_prd): lines with rhoPrd, active atoms first
    std::vector<DevPrdLine> prdLines;
    std::vector<int> prdLineDetailed;
    DevBuf<DevPrdLine> dPrdLines;
    DevBuf<double> qelast, cmat, rhoPrev, prdMax, nOld;
    // the "J20" extra parameter of the Stokes pass: the caller's array and its device copies
    double* j20Host = nullptr;
    DevBuf<double> dJ20, dJ20dag;
    // hybrid PRD (LwB200HybridPrd): the caller's tables (host pointers) and their device copies
    bool hybrid = false;
    LwB200HybridPrd hprdHost{};
    std::vector<unsigned char> hprdPlanMask; // wavelengths the plan routes through the general kernel for it
    DevBuf<int> dHprdLaOfLa, dPrdLaOfLa, dJCoeffIdx, dHprdI0;
    DevBuf<long long> dJCoeffOff;
    DevBuf<double> dJRest, dJCoeffFrac, dHprdFrac;
    DevBuf<double> collC;          // C of every active atom, packed like `prefill` (LWB200_COLLISIONS)
    bool prefillFromC = false;     // finalise with crswC * collC instead of the uploaded prefill
    double crswC = 1.0;
    DevBuf<double> nrScratch, nrDC, nrPrev, nrStages, nrBgNe, neDev;
    DevBuf<NrAtom> nrAtoms;
    DevBuf<int> prdIdx, dKindLamPrd[4], dListMomentPrd, dListDirectPrd;
    DevBuf<double> statEqScratch;   // stat_eq_kernel<64>: matrices of atoms with more than 32 levels
    bool generalViaLong = false;    // the general kernel's shared-memory tile does not fit: fs_long_kernel (REDs) instead
    DevBuf<unsigned char> dPrdMask;
    std::vector<long long> atomCOff;
    long long cTot = 0;
    int nKindLamPrd[4] = {0, 0, 0, 0}, nListMomentPrd = 0, nListDirectPrd = 0, prdListsFor = -1;
    bool prdUploaded = false;
    // full Stokes (lwb200_formal_sol_full_stokes): pool of the polarised lines' six extra profiles
    std::vector<long long> transPolOff; // per global transition, -1: not polarised
    long long polTot = 0;
    DevBuf<double> pol, Quv, Jdag;
    DevBuf<int> dPolLam, dKindLamUnpol[4];
    int nPolLam = 0, nKindLamUnpol[4] = {0, 0, 0, 0};
    bool stokesLists = false, stokesUploaded = false;
    size_t smemGamma = 0;
    int KC = 0;
    int Ntile = 0;
    int nwarps = 4;
    int fsWarps = 4;  // warps per CTA of the general kernel
    int laLo = 0, laHi = 0;
    int NCH = 0;
    size_t smemBytes = 0;
    LaunchCounter lastLaunches;
    bool nstarUploaded = false;

    DevProblem P{};
    DevBuf<double> height, temperature, vlosMu, muz, wmu, wavelength, chiBg, etaBg, scaBg;
    DevBuf<double> J, I, n, nStar, nTotal, vBroad, gRatio, phi, wphi, rhoPrd, aDamp;
    DevBuf<double> wlambdaTab, alphaTab, transWave, lowerBcData, upperBcData, accum, prefill, gamma, dJ;
    DevBuf<double> depthChi, depthEta, depthI, djOut;
    DevBuf<int> lowerBcIdx, upperBcIdx, dLaOff, dLaCnt, dTileLambda, dLaHasLine, dTileLa, dTileSlotOff, dTileSlotTrans;
    DevBuf<int> dAtomNlevel, dAtomLevOff, dAtomGammaOff, dAtomDetailed, dSingular, dPhiAsym;
    DevBuf<long long> djIdx, djPartIdx;
    DevBuf<double> djPartMax;
    DevBuf<unsigned> djTicket;
    cudaEvent_t evDj = nullptr;
    DevBuf<int> dColList;             // active columns of a masked stack
    DevBuf<unsigned char> dColActive;
    int nActiveCol = -1;              // -1: no mask
    size_t scratchGamma = 0;
    int gammaDirect = -1; // -1: by launch size
    cudaEvent_t evRaysKind[4] = {nullptr, nullptr, nullptr, nullptr};
    // Ng acceleration of the populations (lwb200_ng.cuh)
    DevBuf<double> ngRing, ngMax;
    DevBuf<long long> ngIdx;
    int ngOrder = -1, ngPeriod = 0, ngDelay = 0, ngCount = 0;
    double* ngHostMax = nullptr;   // pinned [Natom]: results of the last lwb200_ng_accelerate
    long long* ngHostIdx = nullptr; // pinned [Natom]
    bool singularPending = false; // an asynchronous population update whose singular count is uncollected
    bool djEarly = false, djDone = false; // dJ reduced on a side stream while Gamma is accumulated
    DevBuf<DevTrans> dTrans;
    DevBuf<DevEntry> dEntries;
    // Gamma stage, second generation (lwb200_gamma.cuh): its own tiling and compact tables
    DevBuf<GEntry> dGEntries;
    DevBuf<GLam> dGLam;
    DevBuf<GLine> dGLine;
    DevBuf<int> dGTileLa, dGTileSlotOff, dGList, dGListPrd, dGLamOfLa;
    DevBuf<int4> dGTileSlotRows;
    std::vector<int> gTileLa, gLamLa; // host copies: tile -> range of gLamLa (wavelength indices in tile order)
    GammaPlan G{};
    int gNC = 1, nGTile = 0, nGList = 0, nGListPrd = 0, gWarps = 1, gStage = 0;
    size_t gSmem = 0;
    bool gammaV1 = false; // LWB200_GAMMA_V1=1: the first-generation gamma_kernel (A/B timing aid)
    bool rayV1 = false;   // LWB200_RAY_V1=1: ray_kernel everywhere (A/B timing aid)
    // ZPlaneDecomposition (lwb200_set_zplane)
    DevBuf<double> zUp, zDown;
    double *zUpHost = nullptr, *zDownHost = nullptr;
    bool zDownWritten = false; // the last formal solution swept the down-going rays too
    std::vector<DevLine> devLines;
    DevBuf<DevLine> dLines;
};

namespace
{
// ---------------------------------------------------------------- planning
int build_plan(LwB200Context* c)
{
    const LwB200Problem& p = c->prob;
    const int K = p.Nspace, L = p.Nspect, M = p.Nrays;
    int levOff = 0, gammaOff = 0, nline = 0, ncont = 0, tabOff = 0;
    long long phiOff = 0, rhoOff = 0;
    for (int a = 0; a < p.Natom; ++a)
    {
        const LwB200Atom& at = c->atoms[a];
        c->atomNlevel.push_back(at.Nlevel);
        c->atomLevOff.push_back(levOff);
        c->atomGammaOff.push_back(gammaOff);
        c->atomDetailed.push_back(at.detailedStatic);
        for (int kr = 0; kr < at.Ntrans; ++kr)
        {
            const LwB200Transition& t = c->atomTrans[a][kr];
            if (t.Nblue < 0 || t.Nred > L || t.Nred - t.Nblue < 2)
                return fail("transition wavelength range invalid");
            if (t.i < 0 || t.j < 0 || t.i >= at.Nlevel || t.j >= at.Nlevel)
                return fail("transition level index out of range");
            HostTrans ht{t, a, kr, (int)c->trans.size()};
            c->trans.push_back(ht);
        }
        levOff += at.Nlevel;
        if (!at.detailedStatic)
            gammaOff += at.Nlevel * at.Nlevel;
    }
    const int NlevTot = levOff, GammaTot = gammaOff, NT = (int)c->trans.size();
    const int AccTot = GammaTot + 2 * NT;
    for (int g = 0; g < NT; ++g)
    {
        const HostTrans& ht = c->trans[g];
        const LwB200Transition& t = ht.t;
        const LwB200Atom& at = c->atoms[ht.atom];
        DevTrans d{};
        d.type = t.type;
        d.i = t.i;
        d.j = t.j;
        d.atom = ht.atom;
        d.Nblue = t.Nblue;
        d.Nred = t.Nred;
        d.levI = c->atomLevOff[ht.atom] + t.i;
        d.levJ = c->atomLevOff[ht.atom] + t.j;
        d.detailed = at.detailedStatic;
        if (at.detailedStatic)
        {
            d.accIJ = d.accJI = -1;
        }
        else
        {
            d.accIJ = c->atomGammaOff[ht.atom] + t.i * at.Nlevel + t.j;
            d.accJI = c->atomGammaOff[ht.atom] + t.j * at.Nlevel + t.i;
        }
        d.accRij = GammaTot + 2 * g;
        d.accRji = GammaTot + 2 * g + 1;
        d.tabOff = tabOff;
        const int Nl = t.Nred - t.Nblue;
        tabOff += Nl;
        d.lineIdx = d.contIdx = -1;
        d.rhoOff = -1;
        d.hprdOff = -1;
        if (p.hprd)
            for (int q = 0; q < p.hprd->Nlines; ++q)
                if (p.hprd->lineAtom[q] == ht.atom && p.hprd->lineTrans[q] == ht.kr)
                {
                    if (t.type != LWB200_LINE || !t.rhoPrd)
                        return fail("hybrid PRD coefficients for a transition that is not a PRD line");
                    d.hprdOff = p.hprd->rhoCoefOff[q];
                }
        if (t.type == LWB200_LINE)
        {
            d.lineIdx = nline++;
            d.phiOff = phiOff;
            d.phiColStride = (long long)Nl * M * 2 * K;
            phiOff += d.phiColStride * p.Ncol;
            if (t.rhoPrd)
            {
                d.rhoOff = rhoOff;
                rhoOff += (long long)p.Ncol * Nl * K;
            }
            d.Aji_Bji = t.Aji / t.Bji;
            d.Bji_Bij = t.Bji / t.Bij;
            d.Bij = t.Bij;
        }
        else
        {
            if (!t.alpha)
                return fail("continuum without alpha");
            d.contIdx = ncont++;
        }
        d.lambda0 = t.lambda0;
        c->devTrans.push_back(d);
    }

    // per-wavelength active lists in the reference's (atom, kr) order
    std::vector<int> laOff(L + 1, 0), laHasLine(L, 0), laNLines(L, 0);
    std::vector<std::vector<int>> active(L);
    for (int g = 0; g < NT; ++g)
        for (int la = c->devTrans[g].Nblue; la < c->devTrans[g].Nred; ++la)
        {
            active[la].push_back(g);
            if (c->devTrans[g].type == 0)
            {
                laHasLine[la] = 1;
                laNLines[la] += 1;
            }
        }
    // wavelengths with more than two overlapping lines go to the general kernel
    // kinds 0..3: that many overlapping lines, moment kernel; 4: more, general kernel
    // (the Gamma stage of the moment pipeline stages at most kGammaMaxEntries active transitions per wavelength:
    // a wavelength with more -- far UV of a large atom set -- takes the general kernel as well)
    auto kind_of = [&](int la) {
        return (laNLines[la] > 3 || (int)active[la].size() > kGammaMaxEntries) ? 4 : laNLines[la];
    };

    // per-transition wavelength tables
    std::vector<double> wlambdaTab(tabOff), alphaTab(tabOff, 0.0);
    for (int g = 0; g < NT; ++g)
    {
        const LwB200Transition& t = c->trans[g].t;
        const int Nl = t.Nred - t.Nblue;
        const int off = c->devTrans[g].tabOff;
        for (int lt = 0; lt < Nl; ++lt)
        {
            // Transition::wlambda, LwTransition.hpp:71-81
            double w;
            if (lt == 0)
                w = 0.5 * (t.wavelength[1] - t.wavelength[0]) * t.dopplerWidth;
            else if (lt == Nl - 1)
                w = 0.5 * (t.wavelength[Nl - 1] - t.wavelength[Nl - 2]) * t.dopplerWidth;
            else
                w = 0.5 * (t.wavelength[lt + 1] - t.wavelength[lt - 1]) * t.dopplerWidth;
            wlambdaTab[off + lt] = w;
            if (t.type == LWB200_CONTINUUM)
                alphaTab[off + lt] = t.alpha[lt];
        }
    }

    // tiles: contiguous runs of wavelengths, homogeneous in "needs the general kernel",
    // whose union of active transitions fits the shared-memory accumulator; sized so that
    // (tiles x columns) fills the GPU several times over
    const int KP = ((K + 31) / 32) * 32;
    int maxNlevel = 1;
    for (int a = 0; a < p.Natom; ++a)
        maxNlevel = std::max(maxNlevel, c->atoms[a].Nlevel);
    // (small problems: the general kernel splits the rays of a wavelength over its warps; 8 warps were measured no
    // faster than 4 -- config 4 with hybrid PRD 1.21 vs 1.14 ms)
    const size_t scratchFs = (size_t)c->fsWarps * 2 * maxNlevel * 32 * sizeof(double);
    const int KC = std::min(KP, 128); // depths per gamma_kernel CTA
    const int RS = K <= 128 ? K : KC; // its shared-memory row stride (no padding rows when one chunk covers K)
    const size_t scratchGamma = (size_t)2 * maxNlevel * RS * sizeof(double);
    const size_t scratchBytes = std::max(scratchFs, scratchGamma);
    const size_t smemLimit = 200 * 1024;
    if (scratchBytes + 4 * KC * sizeof(double) > smemLimit)
        return fail("atom too large for the shared-memory scratch");
    // keep the accumulator <= ~48 KB so that several CTAs share an SM
    // per slot: 4 accumulator rows (+ 3 rows of staged per-depth data in gamma_kernel); keep a
    // CTA under ~48 KB so that several share an SM and the L1 keeps some room
    const size_t accBudget = std::min<size_t>(smemLimit - scratchBytes, 32 * 1024);
    const int slotCap = (int)std::max<size_t>(accBudget / (4 * RS * sizeof(double)), 1);
    if ((long long)std::max(ncont, 1) * p.Ncol * K > 0x7fffffffLL)
        return fail("gRatio block exceeds 2^31 elements");
    const long long targetCtas = 148LL * 16;
    int tileLen = (int)std::max<long long>(1, std::min<long long>(32, ((long long)L * p.Ncol) / targetCtas));
    if (const char* e = std::getenv("LWB200_TILE_LEN")) // tuning aid
        tileLen = std::max(1, std::atoi(e));

    std::vector<DevEntry> entries;
    std::vector<int> laCnt(L, 0);
    c->laKind.resize(L);
    for (int la = 0; la < L; ++la)
        c->laKind[la] = kind_of(la);
    c->hprdPlanMask.assign(L, 0);
    if (p.hprd)
    {
        // Hybrid PRD: the emission profile ratio of a PRD line depends on the ray and the formal solution
        // scatters into JRest, neither of which the moment pipeline carries: every wavelength that scatters
        // into the PRD grid in ANY column (the sets differ with the velocity field), and two neighbours
        // either side (room for a later lwb200_set_hybrid_prd), takes the general per-ray kernel.
        for (int col = 0; col < p.Ncol; ++col)
            for (int la = 0; la < L; ++la)
                if (p.hprd->hPrdLaOfLa[(size_t)col * L + la] >= 0)
                    for (int q = std::max(0, la - 2); q <= std::min(L - 1, la + 2); ++q)
                        c->hprdPlanMask[q] = 1;
        for (int la = 0; la < L; ++la)
            if (c->hprdPlanMask[la])
                c->laKind[la] = 4;
    }
    int maxSlots = 1, maxSlotsGeneral = 1;
    size_t maxTileEntries = 1;
    c->tileLa.push_back(0);
    c->tileSlotOff.push_back(0);
    {
        int pos = 0;
        while (pos < L)
        {
            std::vector<int> slots; // transitions of this tile
            const int start = pos;
            const size_t tileEntry0 = entries.size();
            const bool general = c->laKind[start] == 4;
            while (pos < L && pos - start < tileLen && (c->laKind[pos] == 4) == general)
            {
                std::vector<int> add;
                for (int g : active[pos])
                    if (std::find(slots.begin(), slots.end(), g) == slots.end())
                        add.push_back(g);
                if ((int)(slots.size() + add.size()) > slotCap && pos > start)
                    break;
                // (the [slot][4][depth] tile of fs_kernel; deep atmospheres use fs_long_kernel, which keeps none)
                if (K <= 128 && (int)(slots.size() + add.size()) * 4 * (size_t)KP * sizeof(double) + scratchBytes > smemLimit)
                    c->generalViaLong = true; // (fs_kernel's tile of partial sums does not fit: fs_long_kernel accumulates with REDs)
                slots.insert(slots.end(), add.begin(), add.end());
                ++pos;
            }
            for (int l2 = start; l2 < pos; ++l2)
            {
                laOff[l2] = (int)entries.size();
                laCnt[l2] = (int)active[l2].size();
                const size_t eFirst = entries.size();
                for (int g : active[l2])
                {
                    int sl = (int)(std::find(slots.begin(), slots.end(), g) - slots.begin());
                    const DevTrans& d = c->devTrans[g];
                    const int lt = l2 - d.Nblue;
                    DevEntry en{};
                    en.trans = g;
                    en.slot = sl;
                    en.type = d.type;
                    en.i = d.i;
                    en.j = d.j;
                    en.levI = d.levI;
                    en.levJ = d.levJ;
                    en.contIdx = d.contIdx;
                    en.Nlevel = c->atoms[d.atom].Nlevel;
                    en.detailed = d.detailed;
                    en.atom = d.atom;
                    en.prd = d.rhoOff >= 0 ? 1 : 0;
                    en.accIJ = d.accIJ;
                    en.accJI = d.accJI;
                    en.accRij = d.accRij;
                    en.accRji = d.accRji;
                    en.nOffI = d.levI * K;
                    en.nOffJ = d.levJ * K;
                    en.gOff = d.type == 0 ? 0 : (int)((long long)d.contIdx * p.Ncol * K);
                    constexpr double pi4_h = 4.0 * kPi / kHPlanck;
                    constexpr double pi4_hc = 1.0 / (0.25 * kHC / kPi);
                    if (d.type == 0)
                    {
                        en.al = 0.0;
                        en.wlaF = wlambdaTab[d.tabOff + lt] * pi4_hc;
                    }
                    else
                    {
                        en.al = alphaTab[d.tabOff + lt];
                        en.wlaF = (wlambdaTab[d.tabOff + lt] * (1.0 / p.wavelength[l2])) * pi4_h;
                    }
                    entries.push_back(en);
                }
                // atom groups (entries are in (atom, kr) order)
                for (size_t e = eFirst; e < entries.size();)
                {
                    size_t e1 = e + 1;
                    while (e1 < entries.size()
                           && c->devTrans[entries[e1].trans].atom == c->devTrans[entries[e].trans].atom)
                        ++e1;
                    for (size_t q = e; q < e1; ++q)
                        entries[q].groupEnd = (int)e1;
                    e = e1;
                }
                c->tileLambda.push_back(l2);
            }
            maxTileEntries = std::max(maxTileEntries, entries.size() - tileEntry0);
            c->tileLa.push_back((int)c->tileLambda.size());
            c->tileKind.push_back(general ? 4 : 0);
            for (int g : slots)
                c->tileSlotTrans.push_back(g);
            c->tileSlotOff.push_back((int)c->tileSlotTrans.size());
            (general ? maxSlotsGeneral : maxSlots) = std::max(general ? maxSlotsGeneral : maxSlots, (int)slots.size());
        }
    }
    laOff[L] = (int)entries.size();
    c->Ntile = (int)c->tileLa.size() - 1;
    // (fs_kernel -- the general tiles, or every tile on request -- needs room for the largest tile of all, or leaves
    // the general tiles to fs_long_kernel; both shared-memory kernels lay their scratch behind P.maxSlots slots)
    if (!c->generalViaLong)
        maxSlots = std::max(maxSlots, maxSlotsGeneral);
    c->smemGamma = (size_t)maxSlots * 4 * RS * sizeof(double) + scratchGamma;
    c->smemBytes = (size_t)maxSlots * 4 * KP * sizeof(double) + scratchFs + (size_t)c->fsWarps * KP * sizeof(double);
    c->scratchGamma = scratchGamma;
    if (const char* e = std::getenv("LWB200_GAMMA_DIRECT")) // tuning aid: 0 never, 1 always
        c->gammaDirect = std::atoi(e);
    c->KC = KC;
    // (c->smemGamma is what the first-generation gamma_kernel needs: checked where that kernel is selected)
    c->NCH = (K + 31) / 32;
    if (K > 128)
    {
        // beyond one warp per column: ray_kernel in multi-warp mode (4 depths per lane, up to 8 warps);
        // wavelengths of kind 4 (more than three overlapping lines, hybrid PRD) take fs_long_kernel
        // beyond 1024 depths (up to 4096) the general kernel with up to 32 warps per column does everything
        if (K > 4096)
            return fail("Nspace > 4096 is not supported");
        if (fs_long_smem(maxNlevel, 32 * ((K + 127) / 128)) > smemLimit)
            return fail(K > 1024 ? "Nspace > 1024 with an atom of this many levels is not supported (shared-memory scratch)"
                                 : "atom too large for the shared-memory scratch");
        c->NCH = 4;
    }

    // polarised lines: six extra profile arrays each, [6][Ncol][Nl][M][2][K] as on the host
    c->transPolOff.assign(NT, -1);
    for (int g = 0; g < NT; ++g)
        if (c->devTrans[g].type == 0 && (c->trans[g].t.polProfiles || c->trans[g].t.polarised))
        {
            c->transPolOff[g] = c->polTot;
            c->polTot += 6LL * c->devTrans[g].phiColStride * p.Ncol;
        }
    // per-wavelength line slots and moment rows of the pipeline
    std::vector<LambdaLine> lamLine((size_t)L * 3);
    std::vector<int> momOff(L, 0);
    int momRows = 0;
    for (int la = 0; la < L; ++la)
    {
        momOff[la] = momRows;
        if (c->laKind[la] == 4)
            continue;
        momRows += moment_rows(c->laKind[la]);
        int l = 0;
        for (int g : active[la])
        {
            const DevTrans& d = c->devTrans[g];
            if (d.type != 0)
                continue;
            const int Nl = d.Nred - d.Nblue, lt = la - d.Nblue;
            LambdaLine ll{};
            ll.phiOff = d.phiOff + (long long)lt * M * 2 * K;
            ll.phiColStride = d.phiColStride;
            ll.rhoOff = d.rhoOff >= 0 ? d.rhoOff + (long long)lt * K : -1;
            ll.rhoColStride = (long long)Nl * K;
            ll.trans = g;
            ll.levI = d.levI;
            ll.levJ = d.levJ;
            ll.lineIdx = d.lineIdx;
            ll.polOff = c->transPolOff[g] >= 0 ? c->transPolOff[g] + (long long)lt * M * 2 * K : -1;
            ll.polArr = d.phiColStride * p.Ncol;
            ll.slot = -1;
            for (int e = laOff[la]; e < laOff[la] + laCnt[la]; ++e)
                if (entries[e].trans == g)
                    ll.slot = entries[e].slot;
            ll.atom = d.atom;
            ll.i = d.i;
            ll.j = d.j;
            ll.wlaS = wlambdaTab[d.tabOff + lt] * (1.0 / (0.25 * kHC / kPi));
            ll.lambda0 = d.lambda0;
            ll.Bij = d.Bij;
            ll.Bji_Bij = d.Bji_Bij;
            ll.Aji_Bji = d.Aji_Bji;
            lamLine[(size_t)la * 3 + l++] = ll;
        }
    }
    c->momRows = std::max(momRows, 1);

    // ---- Gamma stage, second generation: tiles of the moment wavelengths (kinds 0..3) and the compact
    // per-wavelength tables its warps read (lwb200_gamma.cuh)
    {
        int NCg = 1; // depths per lane; a warp covers 32 * NCg depths
        if (const char* e = std::getenv("LWB200_GAMMA_NC")) // tuning aid
            NCg = std::max(1, std::min(4, std::atoi(e)));
        c->gNC = NCg;
        const size_t rowBytes = (size_t)32 * NCg * sizeof(double);
        int maxAct = 1;
        for (int la = 0; la < L; ++la)
            if (c->laKind[la] < 4)
                maxAct = std::max(maxAct, (int)active[la].size());
        if (maxAct > kGammaMaxEntries)
            return fail("more than 32 active transitions at one wavelength");
        // accumulator budget per warp: what is left of an SM's shared memory at ~6 resident warps
        const int slotCapG = std::max(maxAct, (int)(((size_t)36 << 10) / (4 * rowBytes)));
        int gTileLen = (int)std::max<long long>(1, std::min<long long>(32, ((long long)L * p.Ncol) / (148LL * 12)));
        if (const char* e = std::getenv("LWB200_GTILE_LEN")) // tuning aid
            gTileLen = std::max(1, std::atoi(e));
        // Measured on B200 under ncu (config 3, 128 columns): first generation 1114 us at 21 % resident warps;
        // second generation with 3 depths per lane 1260 us at 8 % (fewer instructions, 3.4e8 vs 4.3e8, but its
        // shared-memory accumulators leave too few warps to hide their own latency), with ONE depth per lane
        // (a warp per 32 depths, 96 registers, 13 KB of shared memory) 902 us at 22 %.  LWB200_GAMMA_V1=1
        // selects the first generation, LWB200_GAMMA_NC the depths per lane.
        c->gammaV1 = false;
        if (const char* e = std::getenv("LWB200_GAMMA_V1"))
            c->gammaV1 = std::atoi(e) != 0;
        if (c->gammaV1 && c->smemGamma > smemLimit)
            return fail("wavelength tile too large for shared memory (first-generation Gamma stage)");
        if (const char* e = std::getenv("LWB200_RAY_V1"))
            c->rayV1 = std::atoi(e) != 0;
        if (const char* e = std::getenv("LWB200_GAMMA_WARPS"))
            c->gWarps = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("LWB200_GAMMA_STAGE"))
            c->gStage = std::atoi(e);
        std::vector<GEntry> gEntries;
        std::vector<GLam> gLam;
        std::vector<GLine> gLine;
        std::vector<int> gTileSlotOff(1, 0);
        std::vector<int4> gTileSlotRows;
        c->gTileLa.assign(1, 0);
        int gMaxSlots = 1;
        constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
        constexpr double twoHc = 2.0 * kHC / (kNmToM * kNmToM * kNmToM);
        constexpr double hc_4pi = 0.25 * kHC / kPi;
        int pos = 0;
        while (pos < L)
        {
            if (c->laKind[pos] == 4)
            {
                ++pos;
                continue;
            }
            std::vector<int> slots;
            const int start = pos;
            while (pos < L && pos - start < gTileLen && c->laKind[pos] < 4)
            {
                std::vector<int> add;
                for (int g : active[pos])
                    if (std::find(slots.begin(), slots.end(), g) == slots.end())
                        add.push_back(g);
                if ((int)(slots.size() + add.size()) > slotCapG && pos > start)
                    break;
                slots.insert(slots.end(), add.begin(), add.end());
                ++pos;
            }
            for (int la = start; la < pos; ++la)
            {
                const double rlambda = 1.0 / p.wavelength[la];
                GLam gl{};
                gl.hc_kl = hc_k * rlambda;
                gl.hcl = twoHc * (rlambda * rlambda * rlambda);
                gl.la = la;
                gl.eOff = (int)gEntries.size();
                gl.eCnt = (int)active[la].size();
                gl.momRow = momOff[la];
                gl.nLines = laNLines[la];
                // the wavelength's entries, in the order of the pipeline's DevEntry list
                const int eBeg = laOff[la];
                int nl = 0;
                for (int q = 0; q < laCnt[la]; ++q)
                {
                    const DevEntry& en = entries[eBeg + q];
                    const DevTrans& d = c->devTrans[en.trans];
                    GEntry ge{};
                    ge.al = en.al;
                    ge.wlaF = en.wlaF;
                    ge.accRow = (short)(4 * (std::find(slots.begin(), slots.end(), en.trans) - slots.begin()));
                    ge.levI = (short)en.levI;
                    ge.levJ = (short)en.levJ;
                    ge.cont = (short)en.contIdx;
                    ge.flags = (unsigned short)((en.prd ? GE_PRD : 0) | (en.detailed ? GE_DETAILED : 0));
                    ge.type = (unsigned char)en.type;
                    ge.i = (unsigned char)en.i;
                    ge.j = (unsigned char)en.j;
                    ge.lslot = (unsigned char)(d.type == 0 ? nl++ : 0);
                    ge.groupLen = 0;
                    ge.atomTag = (unsigned char)en.atom;
                    gEntries.push_back(ge);
                }
                // atom groups: length on the first entry, first-writer / validity flags of the aggregates
                for (size_t e = gl.eOff; e < gEntries.size();)
                {
                    size_t e1 = e + 1;
                    while (e1 < gEntries.size() && gEntries[e1].atomTag == gEntries[e].atomTag)
                        ++e1;
                    gEntries[e].groupLen = (unsigned char)(e1 - e);
                    unsigned touchedX = 0, touchedU = 0; // level bit masks (Nlevel <= 32)
                    for (size_t q = e; q < e1; ++q)
                    {
                        GEntry& t = gEntries[q];
                        if (t.type == 0)
                            continue;
                        if (!(touchedX & (1u << t.i)))
                            t.flags |= GE_STORE_XI;
                        touchedX |= 1u << t.i;
                        if (!(touchedX & (1u << t.j)))
                            t.flags |= GE_STORE_XJ;
                        touchedX |= 1u << t.j;
                        if (!(touchedU & (1u << t.j)))
                            t.flags |= GE_STORE_UJ;
                        touchedU |= 1u << t.j;
                    }
                    for (size_t q = e; q < e1; ++q)
                    {
                        GEntry& t = gEntries[q];
                        t.flags |= (touchedX & (1u << t.i)) ? GE_XI_VALID : 0;
                        t.flags |= (touchedX & (1u << t.j)) ? GE_XJ_VALID : 0;
                        t.flags |= (touchedU & (1u << t.i)) ? GE_UI_VALID : 0;
                        t.flags |= (touchedU & (1u << t.j)) ? GE_UJ_VALID : 0;
                    }
                    e = e1;
                }
                gLam.push_back(gl);
                c->gLamLa.push_back(la);
                for (int l = 0; l < 3; ++l)
                {
                    GLine gn{};
                    gn.rhoOff = -1;
                    if (l < laNLines[la] && c->laKind[la] < 4)
                    {
                        const LambdaLine& ll = lamLine[(size_t)la * 3 + l];
                        gn.vB = hc_4pi * (ll.lambda0 * rlambda) * ll.Bij;
                        gn.gS = ll.Bji_Bij;
                        gn.AB = ll.Aji_Bji;
                        gn.wlaS = ll.wlaS;
                        gn.rhoOff = ll.rhoOff;
                        gn.rhoColStride = ll.rhoColStride;
                        gn.wphiRow = ll.lineIdx;
                        gn.levI = (short)ll.levI;
                        gn.levJ = (short)ll.levJ;
                        gn.i = (unsigned char)ll.i;
                        gn.j = (unsigned char)ll.j;
                        gn.atomTag = (unsigned char)ll.atom;
                    }
                    gLine.push_back(gn);
                }
            }
            for (int g : slots)
            {
                const DevTrans& d = c->devTrans[g];
                gTileSlotRows.push_back(make_int4(d.accIJ, d.accJI, d.accRij, d.accRji));
            }
            gTileSlotOff.push_back((int)gTileSlotRows.size());
            c->gTileLa.push_back((int)gLam.size());
            gMaxSlots = std::max(gMaxSlots, (int)slots.size());
        }
        if (p.Natom > 255)
            return fail("more than 255 atoms");
        c->nGTile = (int)c->gTileLa.size() - 1;
        c->gSmem = gamma_warp_smem(NCg, gMaxSlots, maxNlevel);
        if (c->gSmem > smemLimit)
            return fail("wavelength tile too large for shared memory (Gamma stage)");
        if (c->dGEntries.upload(gEntries) || c->dGLam.upload(gLam) || c->dGLine.upload(gLine)
            || c->dGTileLa.upload(c->gTileLa) || c->dGTileSlotOff.upload(gTileSlotOff)
            || c->dGTileSlotRows.upload(gTileSlotRows))
            return 1;
        c->G.entries = c->dGEntries.p;
        c->G.lam = c->dGLam.p;
        c->G.line = c->dGLine.p;
        c->G.tileLa = c->dGTileLa.p;
        c->G.tileSlotOff = c->dGTileSlotOff.p;
        c->G.tileSlotRows = c->dGTileSlotRows.p;
        c->G.maxSlots = gMaxSlots;
        c->G.maxNlevel = maxNlevel;
        {
            std::vector<int> lamOfLa(L, -1);
            for (size_t q = 0; q < c->gLamLa.size(); ++q)
                lamOfLa[c->gLamLa[q]] = (int)q;
            if (c->dGLamOfLa.upload(lamOfLa))
                return 1;
            c->G.lamOfLa = c->dGLamOfLa.p;
        }
    }

    // PRD lines (PrdTemplates.hpp:186-211: active atoms first, then detailed ones) and the packed
    // layouts of the inputs only lwb200_redistribute_prd reads
    c->atomCOff.assign(p.Natom, -1);
    for (int a = 0; a < p.Natom; ++a)
        if (c->atoms[a].C)
        {
            c->atomCOff[a] = c->cTot;
            c->cTot += (long long)c->atoms[a].Nlevel * c->atoms[a].Nlevel * K;
        }
    for (int pass = 0; pass < 2; ++pass)
        for (int g = 0; g < NT; ++g)
        {
            const DevTrans& d = c->devTrans[g];
            if (d.type != 0 || d.rhoOff < 0 || (d.detailed != 0) != (pass == 1))
                continue;
            DevPrdLine ln{};
            ln.trans = g;
            ln.atom = d.atom;
            ln.j = d.j;
            ln.levI = d.levI;
            ln.levJ = d.levJ;
            ln.Nblue = d.Nblue;
            ln.Nl = d.Nred - d.Nblue;
            ln.tabOff = d.tabOff;
            ln.lineIdx = d.lineIdx;
            ln.Nlevel = c->atoms[d.atom].Nlevel;
            ln.transBeg = g - c->trans[g].kr;
            ln.transEnd = ln.transBeg + c->atoms[d.atom].Ntrans;
            ln.qelRow = (int)c->prdLines.size();
            ln.cOff = c->atomCOff[d.atom];
            ln.rhoOff = d.rhoOff;
            ln.Bij = d.Bij;
            ln.lambda0 = d.lambda0;
            c->prdLines.push_back(ln);
            c->prdLineDetailed.push_back(d.detailed);
        }
    {
        // columns per pipeline batch: chiC + etaC + moments of one batch stay under ~3 GB
        const size_t perCol = ((size_t)2 * L + c->momRows) * K * sizeof(double);
        const size_t cap = (size_t)3 << 30;
        // 512 columns fill the GPU many times over; a smaller stack (a column shard of a multi-GPU run) is still
        // cut into batches of >= 256 so that the copies of one batch travel under the kernels of the next
        // (measured on 8 GPUs x 512 columns: one batch 11.6 ms device / 43.2 ms e2e, four batches 12.1 / 38.8)
        size_t want = std::min<size_t>(512, std::max<size_t>(256, ((size_t)p.Ncol + 3) / 4));
        if (const char* e = std::getenv("LWB200_BATCH_COLS"))
            want = (size_t)std::max(1, atoi(e));
        c->batchCols = (int)std::max<size_t>(1, std::min<size_t>(p.Ncol, std::min<size_t>(want, cap / perCol)));
        if (const char* e = std::getenv("LWB200_BATCH_TAPER"))
            c->batchTaper = atoi(e) != 0;
    }

    // ---- device allocations
    const size_t ncol = p.Ncol;
    if (c->height.alloc(ncol * K) || c->temperature.alloc(ncol * K) || c->muz.alloc(M) || c->wmu.alloc(M)
        || c->wavelength.alloc(L) || c->chiBg.alloc(ncol * L * K) || c->etaBg.alloc(ncol * L * K)
        || c->scaBg.alloc(ncol * L * K) || c->J.alloc(ncol * L * K) || c->I.alloc(ncol * L * M)
        || c->n.alloc(ncol * NlevTot * K) || c->nStar.alloc(ncol * NlevTot * K)
        || c->nTotal.alloc(ncol * p.Natom * K) || c->vBroad.alloc(ncol * p.Natom * K)
        || c->gRatio.alloc((size_t)std::max(ncont, 1) * ncol * K) || c->phi.alloc((size_t)std::max<long long>(phiOff, 1))
        || c->wphi.alloc((size_t)std::max(nline, 1) * ncol * K) || c->aDamp.alloc((size_t)std::max(nline, 1) * ncol * K)
        || c->rhoPrd.alloc((size_t)std::max<long long>(rhoOff, 1))
        || c->accum.alloc(ncol * AccTot * K) || c->prefill.alloc(ncol * std::max(GammaTot, 1) * K)
        || c->gamma.alloc(ncol * std::max(GammaTot, 1) * K) || c->dJ.alloc(ncol * L) || c->djOut.alloc(1)
        || c->djIdx.alloc(1) || c->djPartIdx.alloc(256) || c->djPartMax.alloc(256) || c->djTicket.alloc(1)
        || c->dSingular.alloc(1) || c->dPhiAsym.alloc(1))
        return 1;
    {
        const int one = 3; // until the profiles have been looked at, assume nothing
        CU(cudaMemcpy(c->dPhiAsym.p, &one, sizeof(int), cudaMemcpyHostToDevice));
    }
    if (p.vlosMu && c->vlosMu.alloc(ncol * M * K))
        return 1;
    if (c->polTot > 0)
    {
        if (!p.Quv)
            return fail("polarised lines without a Quv output array");
        if (c->pol.alloc((size_t)c->polTot) || c->Quv.alloc(ncol * 3 * L * M) || c->Jdag.alloc(ncol * L * K))
            return 1;
        CU(cudaMemset(c->Quv.p, 0, c->Quv.n * sizeof(double)));
    }
    if (!c->prdLines.empty())
    {
        const size_t np = c->prdLines.size();
        if (c->qelast.alloc(np * ncol * K) || c->cmat.alloc((size_t)std::max<long long>(c->cTot, 1) * ncol)
            || c->rhoPrev.alloc(c->rhoPrd.n) || c->prdMax.alloc(np) || c->prdIdx.alloc(np)
            || c->dPrdLines.upload(c->prdLines))
            return 1;
    }
    CU(cudaMemset(c->J.p, 0, c->J.n * sizeof(double)));
    CU(cudaMemset(c->I.p, 0, c->I.n * sizeof(double)));
    CU(cudaMemset(c->accum.p, 0, c->accum.n * sizeof(double)));
    CU(cudaMemset(c->prefill.p, 0, c->prefill.n * sizeof(double)));
    CU(cudaMemset(c->gamma.p, 0, c->gamma.n * sizeof(double)));
    CU(cudaMemset(c->dJ.p, 0, c->dJ.n * sizeof(double)));
    CU(cudaMemset(c->djTicket.p, 0, sizeof(unsigned)));
    CU(cudaMemset(c->rhoPrd.p, 0, c->rhoPrd.n * sizeof(double)));
    if (p.depthChi && p.depthEta && p.depthI)
    {
        const size_t nd = ncol * L * M * 2 * K;
        if (c->depthChi.alloc(nd) || c->depthEta.alloc(nd) || c->depthI.alloc(nd))
            return 1;
    }
    if (p.lowerBc == LWB200_BC_CALLABLE)
    {
        if (!p.lowerBcData || !p.lowerBcIdx || p.NlowerBcMu < 1)
            return fail("CALLABLE lower boundary without data");
        if (c->lowerBcData.alloc(ncol * L * p.NlowerBcMu) || c->lowerBcIdx.alloc(M * 2))
            return 1;
        CU(cudaMemcpy(c->lowerBcIdx.p, p.lowerBcIdx, M * 2 * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (p.upperBc == LWB200_BC_CALLABLE)
    {
        if (!p.upperBcData || !p.upperBcIdx || p.NupperBcMu < 1)
            return fail("CALLABLE upper boundary without data");
        if (c->upperBcData.alloc(ncol * L * p.NupperBcMu) || c->upperBcIdx.alloc(M * 2))
            return 1;
        CU(cudaMemcpy(c->upperBcIdx.p, p.upperBcIdx, M * 2 * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (c->wlambdaTab.upload(wlambdaTab) || c->alphaTab.upload(alphaTab) || c->dTrans.upload(c->devTrans)
        || c->dEntries.upload(entries) || c->dLaOff.upload(laOff) || c->dLaHasLine.upload(laHasLine)
        || c->dLaCnt.upload(laCnt) || c->dTileLambda.upload(c->tileLambda) || c->dTileLa.upload(c->tileLa) || c->dTileSlotOff.upload(c->tileSlotOff)
        || c->dTileSlotTrans.upload(c->tileSlotTrans) || c->dAtomNlevel.upload(c->atomNlevel)
        || c->dAtomLevOff.upload(c->atomLevOff) || c->dAtomGammaOff.upload(c->atomGammaOff)
        || c->dAtomDetailed.upload(c->atomDetailed) || c->dMomOff.upload(momOff) || c->dLaNLines.upload(laNLines)
        || c->dLamLine.upload(lamLine))
        return 1;
    {
        const size_t bc = (size_t)c->batchCols;
        if (c->chiC.alloc(bc * L * K) || c->etaC.alloc(bc * L * K) || c->mom.alloc(bc * c->momRows * K))
            return 1;
    }
    CU(cudaMemcpy(c->muz.p, p.muz, M * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->wmu.p, p.wmu, M * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->wavelength.p, p.wavelength, L * sizeof(double), cudaMemcpyHostToDevice));

    // line descriptors for the device profile kernel
    {
        std::vector<double> transWave(tabOff);
        for (int g = 0; g < NT; ++g)
        {
            const LwB200Transition& t = c->trans[g].t;
            for (int lt = 0; lt < t.Nred - t.Nblue; ++lt)
                transWave[c->devTrans[g].tabOff + lt] = t.wavelength[lt];
        }
        if (c->transWave.upload(transWave))
            return 1;
        for (int g = 0; g < NT; ++g)
        {
            const DevTrans& d = c->devTrans[g];
            if (d.type != 0)
                continue;
            DevLine dl{};
            dl.Nl = d.Nred - d.Nblue;
            dl.tabOff = d.tabOff;
            dl.lineIdx = d.lineIdx;
            dl.atom = d.atom;
            dl.phiOff = d.phiOff;
            dl.phiColStride = d.phiColStride;
            dl.lambda0 = d.lambda0;
            c->devLines.push_back(dl);
        }
        if (c->dLines.upload(c->devLines))
            return 1;
    }

    DevProblem& P = c->P;
    P.Ncol = p.Ncol; P.K = K; P.M = M; P.L = L;
    P.NlevTot = NlevTot; P.GammaTot = GammaTot; P.AccTot = AccTot; P.Natom = p.Natom;
    P.NtransTot = NT; P.Nline = nline; P.Ncont = ncont;
    P.lowerBc = p.lowerBc; P.upperBc = p.upperBc;
    P.NlowerBcMu = p.NlowerBcMu; P.NupperBcMu = p.NupperBcMu;
    P.maxNlevel = maxNlevel; P.maxSlots = maxSlots; P.KP = KP;
    P.height = c->height.p; P.temperature = c->temperature.p; P.muz = c->muz.p; P.wmu = c->wmu.p;
    P.wavelength = c->wavelength.p; P.chiBg = c->chiBg.p; P.etaBg = c->etaBg.p; P.scaBg = c->scaBg.p;
    P.J = c->J.p; P.I = c->I.p; P.n = c->n.p; P.nStar = c->nStar.p; P.gRatio = c->gRatio.p;
    P.phi = c->phi.p; P.wphi = c->wphi.p; P.rhoPrd = c->rhoPrd.p;
    P.wlambdaTab = c->wlambdaTab.p; P.alphaTab = c->alphaTab.p;
    P.lowerBcData = c->lowerBcData.p; P.upperBcData = c->upperBcData.p;
    P.lowerBcIdx = c->lowerBcIdx.p; P.upperBcIdx = c->upperBcIdx.p;
    P.accum = c->accum.p; P.dJ = c->dJ.p;
    P.depthChi = c->depthChi.p; P.depthEta = c->depthEta.p; P.depthI = c->depthI.p;
    P.trans = c->dTrans.p; P.entries = c->dEntries.p; P.laOff = c->dLaOff.p; P.laHasLine = c->dLaHasLine.p;
    P.laCnt = c->dLaCnt.p; P.tileLambda = c->dTileLambda.p;
    P.tileLa = c->dTileLa.p; P.tileSlotOff = c->dTileSlotOff.p; P.tileSlotTrans = c->dTileSlotTrans.p;
    P.atomNlevel = c->dAtomNlevel.p; P.atomLevOff = c->dAtomLevOff.p;
    P.atomGammaOff = c->dAtomGammaOff.p; P.atomDetailed = c->dAtomDetailed.p;
    P.chiC = c->chiC.p; P.etaC = c->etaC.p; P.mom = c->mom.p; P.momOff = c->dMomOff.p;
    P.laNLines = c->dLaNLines.p; P.lamLine = c->dLamLine.p; P.momRows = c->momRows;
    P.phiAsym = c->dPhiAsym.p;
    P.pol = c->pol.p; P.Quv = c->Quv.p; P.Jdag = c->Jdag.p;
    return 0;
}

int refresh_tile_lists(LwB200Context* c)
{
    if (c->listLo == c->laLo && c->listHi == c->laHi)
        return 0;
    std::vector<int> mom, dir, all, kindLam[4];
    for (int t = 0; t < c->Ntile; ++t)
    {
        // tile wavelength lists are ascending: overlap test on first / last
        if (c->tileLambda[c->tileLa[t + 1] - 1] < c->laLo || c->tileLambda[c->tileLa[t]] >= c->laHi)
            continue;
        all.push_back(t);
        (c->tileKind[t] < 4 ? mom : dir).push_back(t);
    }
    for (int la = c->laLo; la < c->laHi; ++la)
        if (c->laKind[la] < 4)
            kindLam[c->laKind[la]].push_back(la);
    c->dListDirect.release();
    c->dListAll.release();
    c->dListMoment.release();
    if (c->dListMoment.upload(mom))
        return 1;
    c->nListMoment = (int)mom.size();
    for (int q = 0; q < 4; ++q)
    {
        c->dKindLam[q].release();
        if (c->dKindLam[q].upload(kindLam[q]))
            return 1;
        c->nKindLam[q] = (int)kindLam[q].size();
    }
    {
        std::vector<int> gl;
        for (int t = 0; t < c->nGTile; ++t)
            if (!(c->gLamLa[c->gTileLa[t + 1] - 1] < c->laLo || c->gLamLa[c->gTileLa[t]] >= c->laHi))
                gl.push_back(t);
        c->dGList.release();
        if (c->dGList.upload(gl))
            return 1;
        c->nGList = (int)gl.size();
    }
    if (c->dListDirect.upload(dir) || c->dListAll.upload(all))
        return 1;
    c->nListDirect = (int)dir.size();
    c->nListAll = (int)all.size();
    c->listLo = c->laLo;
    c->listHi = c->laHi;
    return 0;
}

template <typename Kern>
int set_smem_attr(Kern kern, int device)
{
    // per (kernel, device): opt in to > 48 KB of dynamic shared memory
    static std::set<std::pair<const void*, int>> done;
    const auto key = std::make_pair((const void*)kern, device);
    if (done.count(key))
        return 0;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    done.insert(key);
    return 0;
}

static bool stream_is_capturing_fwd(cudaStream_t s)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

int check_phi_symmetry(LwB200Context* c)
{
    if (c->P.Nline == 0)
        return 0;
    CU(cudaMemsetAsync(c->dPhiAsym.p, 0, sizeof(int), c->stream));
    const size_t nPairs = c->phi.n / ((size_t)2 * c->P.K);
    phi_symmetry_kernel<<<148 * 8, 256, 0, c->stream>>>(c->phi.p, nPairs, c->P.K, c->P.M, c->dPhiAsym.p);
    CU(cudaGetLastError());
    // deep atmospheres pick their ray kernel from the flags on the host (the profiles have just been uploaded or
    // generated: one small synchronous read-back here, never inside an iteration)
    c->phiFlagsHost = 3;
    if (c->prob.Nspace > 128 && !stream_is_capturing_fwd(c->stream))
    {
        CU(cudaMemcpyAsync(&c->phiFlagsHost, c->dPhiAsym.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// columns a launch covers: the active ones of a masked stack
static inline int launch_columns(const LwB200Context* c)
{
    return c->nActiveCol >= 0 ? c->nActiveCol : c->prob.Ncol;
}

// (max, argmax) of dJ over [laLo, laHi) of every column into djOut / djIdx, on stream s
// finalise_Gamma of columns [colBase, colBase + ncols) on the context's stream
static int launch_finalise(LwB200Context* c, int colBase, int ncols)
{
    const size_t total = (size_t)ncols * c->P.Natom * c->P.maxNlevel * c->P.K;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 127) / 128, 148 * 32));
    finalise_kernel<<<grid, 128, 0, c->stream>>>(
        c->P, c->prefillFromC ? c->collC.p : c->prefill.p, c->prefillFromC ? c->crswC : 1.0, c->gamma.p, colBase, ncols);
    CU(cudaGetLastError());
    c->lastLaunches += 1;
    return 0;
}

static int launch_dj_reduce(LwB200Context* c, cudaStream_t s, int laLo, int laHi, const unsigned char* mask)
{
    const long long total = (long long)c->P.Ncol * (laHi - laLo);
    const int grid = (int)std::max<long long>(1, std::min<long long>(256, (total + 1023) / 1024));
    dj_reduce_kernel<<<grid, 256, 0, s>>>(c->dJ.p, c->P.Ncol, c->P.L, laLo, laHi, c->djOut.p, c->djIdx.p, mask,
                                          c->djPartMax.p, c->djPartIdx.p, c->djTicket.p, c->P.colActive);
    CU(cudaGetLastError());
    c->lastLaunches += 1;
    return 0;
}

int ensure_host_scalars(LwB200Context* c)
{
    if (c->hs)
        return 0;
    // mapped: the population kernels count singular systems straight into it (no copy to wait for)
    CU(cudaHostAlloc((void**)&c->hs, sizeof(*c->hs), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void**)&c->hsDev, c->hs, 0));
    c->hs->dJ = 0.0;
    c->hs->dJIdx = 0;
    c->hs->nSingular = 0;
    return 0;
}

int ensure_side_streams(LwB200Context* c)
{
    if (c->evFork)
        return 0;
    int prLow = 0, prHigh = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh)); // numerically lower = more urgent
    for (int q = 0; q < 4; ++q)
    {
        // earlier side streams carry the more expensive kinds: give their CTAs precedence
        const int pr = std::max(prHigh, std::min(prLow, prLow - (3 - q)));
        CU(cudaStreamCreateWithPriority(&c->sideStream[q], cudaStreamNonBlocking, pr));
        CU(cudaEventCreateWithFlags(&c->evJoin[q], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->evRaysKind[q], cudaEventDisableTiming));
    }
    CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->evRays, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evCopy, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evDj, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    return 0;
}

// Three-stage pipeline per batch of columns (lwb200_pipeline.cuh):
//   continuum_kernel -> ray_kernel<NL> for NL = 3, 2, 1, 0 -> gamma_kernel
// The ray kernels of the four kinds run CONCURRENTLY, most expensive per wavelength first, on
// prioritised side streams forked from / joined to the caller's stream, so that they share one
// tail (a 1D atmosphere is only a few waves of warps in total).  fsMode != 0: formal solution
// only (no J, no moments, no Gamma stage).
PipelineLists full_lists(LwB200Context* c)
{
    PipelineLists pl{};
    pl.moment = c->dListMoment.p;
    pl.nMoment = c->nListMoment;
    for (int q = 0; q < 4; ++q)
    {
        pl.kindLam[q] = c->dKindLam[q].p;
        pl.nKindLam[q] = c->nKindLam[q];
    }
    pl.gTiles = c->dGList.p;
    pl.nGTiles = c->nGList;
    pl.laMask = nullptr;
    pl.prdOnly = 0;
    return pl;
}

// Second-generation Gamma stage: one warp per (tile, column, chunk of 32 * gNC depths), the warps of a CTA
// on consecutive tiles of one column whose populations / continuum ratios they stage once in shared
// memory when that leaves room for at least two warps.
template <int NC, bool STAGE>
int launch_gamma_tiles_t(LwB200Context* c, const PipelineLists& pl, int nb, int colBase, int laLo, int laHi,
                         int warps, size_t smem, int warpDoubles)
{
    auto kern = gamma_tile_kernel<NC, STAGE>;
    if (set_smem_attr(kern, c->device))
        return 1;
    const int chunk = 32 * NC;
    const dim3 gg((pl.nGTiles + warps - 1) / warps, nb, (c->P.K + chunk - 1) / chunk);
    kern<<<gg, 32 * warps, smem, c->stream>>>(c->P, c->G, pl.gTiles, pl.nGTiles, laLo, laHi, colBase, pl.laMask,
                                             pl.prdOnly, warpDoubles);
    CU(cudaGetLastError());
    c->lastLaunches += 1;
    return 0;
}

int launch_gamma_tiles(LwB200Context* c, const PipelineLists& pl, int nb, int colBase, int laLo, int laHi)
{
    const int NC = c->gNC;
    const size_t warpBytes = gamma_warp_smem(NC, c->G.maxSlots, c->G.maxNlevel);
    const size_t stageBytes = gamma_stage_smem(NC, c->P.NlevTot, c->P.Ncont);
    const size_t budget = 216 * 1024;
    // Measured on B200 (config 3, 512 columns): single-warp CTAs reading populations / ratios through the
    // L1 13.9 ms per launch set, 5-warp CTAs with the column staged in shared memory 14.7 ms -- the staging
    // costs a resident warp and a barrier.  Both stay selectable (tuning aids).
    const int envWarps = c->gWarps, envStage = c->gStage; // (read from the environment at context creation)
    const bool stage = envStage != 0 && stageBytes + 2 * warpBytes <= budget;
    int warps = (int)((budget - (stage ? stageBytes : 0)) / warpBytes);
    warps = std::max(1, std::min({warps, 8, pl.nGTiles, std::max(1, envWarps)}));
    if (!stage)
        warps = 1; // (the unstaged kernel is compiled for single-warp CTAs)
    const size_t smem = (stage ? stageBytes : 0) + warps * warpBytes;
    const int wd = (int)(warpBytes / sizeof(double));
#define LWB200_GT(NCV)                                                                                           \
    case NCV:                                                                                                    \
        return stage ? launch_gamma_tiles_t<NCV, true>(c, pl, nb, colBase, laLo, laHi, warps, smem, wd)          \
                     : launch_gamma_tiles_t<NCV, false>(c, pl, nb, colBase, laLo, laHi, warps, smem, wd);
    switch (NC)
    {
        LWB200_GT(1)
        LWB200_GT(2)
        LWB200_GT(3)
    default:
        LWB200_GT(4)
    }
#undef LWB200_GT
    return 0;
}

template <int NCH, int SOLVER, bool MULTI>
int launch_pipeline(LwB200Context* c, const PipelineLists& pl, int lambdaIterate, int storeDepth, int fsMode)
{
    if (ensure_side_streams(c))
        return 1;
    if (set_smem_attr(gamma_kernel, c->device))
        return 1;
    const int Ncol = launch_columns(c), KP = c->P.KP, K = c->P.K;
    const int laLo = pl.fullRange ? 0 : c->laLo, laHi = pl.fullRange ? c->prob.Nspect : c->laHi;
    const int threads = MULTI ? 32 * ((K + 32 * NCH - 1) / (32 * NCH)) : c->nwarps * 32;
    // With the early fetch each batch's results travel home under the following batches, and what is exposed
    // is the last batch's copies: the last full batch is cut into halves of halves (512 -> 256, 128, 64, 64).
    const bool taper = c->fetchEarly && c->finaliseEarly && fsMode == 0 && !pl.prdOnly && c->batchTaper;
    for (int colBase = 0, nbNext = 0; colBase < Ncol; colBase += nbNext)
    {
        int nbThis = std::min(c->batchCols, Ncol - colBase);
        if (taper && Ncol - colBase <= c->batchCols && nbThis >= 128)
            nbThis = (nbThis + 1) / 2;
        nbNext = nbThis;
        const int nb = nbThis;
        constexpr int contPerBlock = 4; // wavelengths per continuum_kernel CTA
        // small launches (one small atmosphere, a wavelength shard of a larger one): the Gamma stage
        // goes per kind and straight to the global accumulator (gamma_direct_kernel); measured on B200:
        // 1257 wavelengths 0.123 vs 0.139 ms, 5028: 0.311 vs 0.332 ms, 10056: 0.600 vs 0.538 ms
        const long long lamCols = ((long long)pl.nKindLam[0] + pl.nKindLam[1] + pl.nKindLam[2] + pl.nKindLam[3]) * nb;
        const bool direct = fsMode == 0 && (c->gammaDirect < 0 ? lamCols <= 6000 : c->gammaDirect != 0);
        // a stream that must see this batch's J complete waits for the rays of every kind
        auto wait_rays = [&](cudaStream_t st) -> int {
            if (direct)
            {
                for (int q = 0; q < 4; ++q)
                    if (pl.nKindLam[q] > 0)
                        CU(cudaStreamWaitEvent(st, c->evRaysKind[q], 0));
                return 0;
            }
            CU(cudaEventRecord(c->evRays, c->stream));
            CU(cudaStreamWaitEvent(st, c->evRays, 0));
            return 0;
        };
        int nkinds = 0;
        for (int q = 0; q < 4; ++q)
            nkinds += pl.nKindLam[q] > 0 ? 1 : 0;
        int nside = 0, launched = 0;
        for (int q = 3; q >= 0; --q)
        {
            const int nLam = pl.nKindLam[q];
            if (nLam == 0)
                continue;
            // the last kind runs on the caller's stream itself
            cudaStream_t s = c->stream;
            if (++launched != nkinds)
            {
                if (nside == 0)
                    CU(cudaEventRecord(c->evFork, c->stream));
                s = c->sideStream[nside++];
                CU(cudaStreamWaitEvent(s, c->evFork, 0));
            }
            const long long warps = (long long)nLam * nb;
            const int perWarp = MULTI ? 1 : (int)std::max<long long>(1, std::min<long long>(4, warps / (148LL * 8 * 8)));
            const int nw = c->nwarps;
            dim3 grid(MULTI ? nLam : (nLam + nw * perWarp - 1) / (nw * perWarp), nb);
            const int* list = pl.kindLam[q];
            // (a small launch -- one 1D atmosphere -- is a latency chain of dependent loads per wavelength: one
            // wavelength per CTA there; config 1: 17 us with four)
            const int cpb = (long long)nLam * nb <= 148LL * 64 ? 1 : contPerBlock;
            continuum_table_kernel<<<dim3((nLam + cpb - 1) / cpb, nb), KP, 0, s>>>(c->P, c->G, list, nLam, cpb, colBase);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
            // Bezier3, one warp per wavelength, both directions: the two rays of a mu solved together
            // (ray_smem_kernel evaluates the end points of all rays of a wavelength in one go, lane per ray,
            // before it reuses the rows they are made from: up to 32 rays)
            const bool pairKernel = SOLVER == 2 && !MULTI && !storeDepth && !(fsMode & 2) && !c->rayV1
                && 2 * c->prob.Nrays <= 32;
            if (pairKernel)
            {
                if constexpr (SOLVER == 2 && !MULTI)
                {
#ifndef LWB200_RAY3_MINB
#define LWB200_RAY3_MINB 3
#endif
#define LWB200_RAY3(NLV)                                                                                                   \
    {                                                                                                                      \
        auto kern = ray_smem_kernel<NCH, NLV, LWB200_RAY3_MINB>;                                                          \
        if (set_smem_attr(kern, c->device))                                                                                \
            return 1;                                                                                                      \
        const dim3 g3((nLam + LWB200_RAY_WARPS * perWarp - 1) / (LWB200_RAY_WARPS * perWarp), nb);                        \
        kern<<<g3, 32 * LWB200_RAY_WARPS, ray_smem_bytes<NCH, NLV>(), s>>>(c->P, list, nLam, perWarp, colBase,            \
                                                                           lambdaIterate, fsMode);                        \
    }
                    {
                        switch (q)
                        {
                        case 0: LWB200_RAY3(0) break;
                        case 1: LWB200_RAY3(1) break;
                        case 2: LWB200_RAY3(2) break;
                        default: LWB200_RAY3(3) break;
                        }
                    }
#undef LWB200_RAY3
                }
            }
            else if (MULTI && SOLVER == 2 && q > 0 && c->phiFlagsHost == 0 && !storeDepth && !(fsMode & 2))
            {
                // deep static atmosphere: ray-independent profiles (flags read back after the symmetry check)
                if constexpr (MULTI && SOLVER == 2)
                {
                    switch (q)
                    {
                    case 1:
                        ray_kernel<NCH, SOLVER, 1, MULTI, true><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                        break;
                    case 2:
                        ray_kernel<NCH, SOLVER, 2, MULTI, true><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                        break;
                    default:
                        ray_kernel<NCH, SOLVER, 3, MULTI, true><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                        break;
                    }
                }
            }
            else
            switch (q)
            {
            case 0:
                ray_kernel<NCH, SOLVER, 0, MULTI><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                break;
            case 1:
                ray_kernel<NCH, SOLVER, 1, MULTI><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                break;
            case 2:
                ray_kernel<NCH, SOLVER, 2, MULTI><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                break;
            default:
                ray_kernel<NCH, SOLVER, 3, MULTI><<<grid, threads, 0, s>>>(c->P, list, nLam, perWarp, colBase, lambdaIterate, storeDepth, fsMode);
                break;
            }
            CU(cudaGetLastError());
            c->lastLaunches += 1;
            if (direct)
            {
                // this kind's share of the Gamma stage follows its rays on the same stream
                CU(cudaEventRecord(c->evRaysKind[q], s));
                if (c->scratchGamma > 48 * 1024 // (atoms of more than 36 levels: opt in to more shared memory)
                    && (set_smem_attr(gamma_direct_kernel<0>, c->device) || set_smem_attr(gamma_direct_kernel<1>, c->device)
                        || set_smem_attr(gamma_direct_kernel<2>, c->device) || set_smem_attr(gamma_direct_kernel<3>, c->device)))
                    return 1;
                const dim3 gg(nLam, nb, (K + c->KC - 1) / c->KC);
                switch (q)
                {
                case 0: gamma_direct_kernel<0><<<gg, c->KC, c->scratchGamma, s>>>(c->P, list, nLam, colBase, pl.prdOnly); break;
                case 1: gamma_direct_kernel<1><<<gg, c->KC, c->scratchGamma, s>>>(c->P, list, nLam, colBase, pl.prdOnly); break;
                case 2: gamma_direct_kernel<2><<<gg, c->KC, c->scratchGamma, s>>>(c->P, list, nLam, colBase, pl.prdOnly); break;
                default: gamma_direct_kernel<3><<<gg, c->KC, c->scratchGamma, s>>>(c->P, list, nLam, colBase, pl.prdOnly); break;
                }
                CU(cudaGetLastError());
                c->lastLaunches += 1;
            }
        }
        if (pl.nPolLam > 0)
        {
            // polarised wavelengths: one thread per ray, concurrently with the scalar ray kernels
            const int nRays = pl.nPolLam * 2 * c->prob.Nrays;
            continuum_kernel<<<dim3((pl.nPolLam + contPerBlock - 1) / contPerBlock, nb), KP, 0, c->stream>>>(
                c->P, pl.polLam, pl.nPolLam, contPerBlock, colBase);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
            static const bool stokesV1 = std::getenv("LWB200_STOKES_V1") && std::atoi(std::getenv("LWB200_STOKES_V1")) != 0;
            (stokesV1 ? stokes_kernel_v1 : stokes_kernel)<<<dim3((nRays + 127) / 128, nb), 128, 0, c->stream>>>(
                c->P, pl.polLam, pl.nPolLam, colBase, (fsMode & 2) ? 1 : 0, (fsMode & 4) ? 1 : 0);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
        }
        for (int q = 0; q < nside; ++q)
        {
            CU(cudaEventRecord(c->evJoin[q], c->sideStream[q]));
            CU(cudaStreamWaitEvent(c->stream, c->evJoin[q], 0));
        }
        if (fsMode != 0)
            continue;
        if (c->djEarly && colBase + nb >= Ncol)
        {
            // J is complete: reduce dJ beside the Gamma stage; the result lands in pinned host memory
            // (HostScalars).  With the per-kind Gamma stage every side stream is busy: the copy stream
            // takes it (ahead of the copies).
            cudaStream_t sd = direct ? c->copyStream : c->sideStream[0];
            if (wait_rays(sd))
                return 1;
            if (launch_dj_reduce(c, sd, laLo, laHi, nullptr))
                return 1;
            CU(cudaMemcpyAsync(&c->hs->dJ, c->djOut.p, sizeof(double), cudaMemcpyDeviceToHost, sd));
            CU(cudaMemcpyAsync(&c->hs->dJIdx, c->djIdx.p, sizeof(long long), cudaMemcpyDeviceToHost, sd));
            CU(cudaEventRecord(c->evDj, sd));
            c->djDone = true;
        }
        if (c->fetchEarly && !pl.prdOnly && c->nListDirect == 0)
        {
            // J and I of this batch of columns are final: send them home on the copy stream while
            // Gamma is accumulated (and, in a column stack, while the next batches are computed)
            const LwB200Problem& p = c->prob;
            const size_t perJ = (size_t)p.Nspect * p.Nspace, perI = (size_t)p.Nspect * p.Nrays;
            if (wait_rays(c->copyStream))
                return 1;
            CU(cudaMemcpyAsync(p.J + colBase * perJ, c->J.p + colBase * perJ, nb * perJ * sizeof(double),
                               cudaMemcpyDeviceToHost, c->copyStream));
            CU(cudaMemcpyAsync(p.I + colBase * perI, c->I.p + colBase * perI, nb * perI * sizeof(double),
                               cudaMemcpyDeviceToHost, c->copyStream));
            if (colBase + nb >= Ncol)
            {
                CU(cudaEventRecord(c->evCopy, c->copyStream)); // (recorded again behind Gamma and the rates below)
                c->fetched = true;
            }
        }
        if (!direct && c->gammaV1 && pl.nMoment > 0)
        {
            const int KC = c->KC;
            gamma_kernel<<<dim3(pl.nMoment, nb, (K + KC - 1) / KC), KC, c->smemGamma, c->stream>>>(
                c->P, pl.moment, laLo, laHi, colBase, pl.laMask, pl.prdOnly);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
        }
        else if (!direct && pl.nGTiles > 0)
        {
            if (launch_gamma_tiles(c, pl, nb, colBase, laLo, laHi))
                return 1;
        }
        if (c->finaliseEarly && c->fetchEarly && !pl.prdOnly && c->nListDirect == 0)
        {
            // this batch's Gamma and rates are complete as well: finalise its columns and send them home
            // behind its J and I, under the following batches (only the last batch's copies are exposed)
            const LwB200Problem& p = c->prob;
            const size_t D = sizeof(double), Ks = (size_t)K;
            if (c->P.GammaTot > 0 && launch_finalise(c, colBase, nb))
                return 1;
            CU(cudaEventRecord(c->evRays, c->stream));
            CU(cudaStreamWaitEvent(c->copyStream, c->evRays, 0));
            for (int a = 0; a < p.Natom; ++a)
            {
                if (c->atoms[a].detailedStatic)
                    continue;
                const size_t N2 = (size_t)c->atoms[a].Nlevel * c->atoms[a].Nlevel;
                CU(cudaMemcpy2DAsync(c->atoms[a].Gamma + colBase * N2 * Ks, N2 * Ks * D,
                                     c->gamma.p + ((size_t)colBase * c->P.GammaTot + c->atomGammaOff[a]) * Ks,
                                     (size_t)c->P.GammaTot * Ks * D, N2 * Ks * D, nb, cudaMemcpyDeviceToHost, c->copyStream));
            }
            for (size_t g = 0; g < c->trans.size(); ++g)
            {
                const LwB200Transition& t = c->trans[g].t;
                const DevTrans& d = c->devTrans[g];
                const double* src = c->accum.p + (size_t)colBase * c->P.AccTot * Ks;
                CU(cudaMemcpy2DAsync(t.Rij + colBase * Ks, Ks * D, src + (size_t)d.accRij * Ks, (size_t)c->P.AccTot * Ks * D,
                                     Ks * D, nb, cudaMemcpyDeviceToHost, c->copyStream));
                CU(cudaMemcpy2DAsync(t.Rji + colBase * Ks, Ks * D, src + (size_t)d.accRji * Ks, (size_t)c->P.AccTot * Ks * D,
                                     Ks * D, nb, cudaMemcpyDeviceToHost, c->copyStream));
            }
            if (colBase + nb >= Ncol)
            {
                CU(cudaEventRecord(c->evCopy, c->copyStream));
                c->finalisedEarly = c->fetchedGR = true;
            }
        }
    }
    return 0;
}

// a stream being captured into a CUDA graph (the caller replays whole iterations): no timing events there
static bool stream_is_capturing(cudaStream_t s)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

// the general multi-warp kernel over a list of tiles (fs_long_kernel: any depth up to 4096, partial sums by RED)
template <int SOLVER>
int launch_general_long(LwB200Context* c, const int* tiles, int nTiles, int laLo, int laHi, int lambdaIterate, int upOnly,
                        int storeDepth, int prdOnly, int fsOnly, const unsigned char* mask)
{
    const int K = c->prob.Nspace, threads = 32 * ((K + 127) / 128);
    const bool big = threads > 256;
    auto kern = big ? fs_long_kernel<SOLVER, true> : fs_long_kernel<SOLVER, false>;
    if (set_smem_attr(kern, c->device))
        return 1;
    dim3 grid(nTiles, launch_columns(c));
    kern<<<grid, threads, fs_long_smem(c->P.maxNlevel, threads), c->stream>>>(c->P, tiles, laLo, laHi, lambdaIterate, upOnly,
                                                                             storeDepth, prdOnly, fsOnly, mask);
    CU(cudaGetLastError());
    c->lastLaunches += 1;
    return 0;
}

template <int NCH, int SOLVER, int MODE>
int launch_fs_t(LwB200Context* c, int lambdaIterate, int upOnly, int storeDepth)
{
    if (!c->evK0)
    {
        CU(cudaEventCreate(&c->evK0));
        CU(cudaEventCreate(&c->evK1));
    }
    const bool capturing = stream_is_capturing(c->stream);
    if (!capturing)
        CU(cudaEventRecord(c->evK0, c->stream));
    const int threads = c->fsWarps * 32;
    if (MODE == MODE_ITER && !c->forceDirect)
    {
        // wavelengths with more than three overlapping lines go through the general kernel
        if (launch_pipeline<NCH, SOLVER, false>(c, c->customLists ? c->prdPl : full_lists(c), lambdaIterate, storeDepth,
                                                c->stokesFsMode))
            return 1;
        const bool maskPass = c->customLists && !c->prdPl.directPrdOnly; // angle-averaged PRD sub-iteration
        const int nDirect = maskPass ? c->prdPl.nDirectPrd : c->nListDirect;
        if (nDirect > 0 && c->generalViaLong)
        {
            const bool prdPass = c->customLists;
            if (launch_general_long<SOLVER>(c, maskPass ? c->prdPl.directPrd : c->dListDirect.p, nDirect, prdPass ? 0 : c->laLo,
                                            prdPass ? c->prob.Nspect : c->laHi, lambdaIterate, 0, storeDepth,
                                            maskPass ? 2 : (prdPass ? 1 : 0), 0, maskPass ? c->prdPl.laMask : nullptr))
                return 1;
        }
        else if (nDirect > 0)
        {
            auto kern = fs_kernel<NCH, SOLVER, MODE_ITER>;
            if (set_smem_attr(kern, c->device))
                return 1;
            const bool prdPass = c->customLists; // (PRD sub-iteration: the whole spectrum, PRD rates only)
            dim3 grid(nDirect, launch_columns(c));
            kern<<<grid, threads, c->smemBytes, c->stream>>>(
                c->P, maskPass ? c->prdPl.directPrd : c->dListDirect.p, prdPass ? 0 : c->laLo,
                prdPass ? c->prob.Nspect : c->laHi, lambdaIterate, 0, storeDepth, maskPass ? 2 : (prdPass ? 1 : 0),
                maskPass ? c->prdPl.laMask : nullptr);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
        }
    }
    else if (c->nListAll > 0 && c->generalViaLong)
    {
        if (launch_general_long<SOLVER>(c, c->dListAll.p, c->nListAll, c->laLo, c->laHi, lambdaIterate, upOnly, storeDepth, 0,
                                        MODE == MODE_ITER ? 0 : 1, nullptr))
            return 1;
    }
    else if (c->nListAll > 0)
    {
        auto kern = fs_kernel<NCH, SOLVER, MODE>;
        if (set_smem_attr(kern, c->device))
            return 1;
        dim3 grid(c->nListAll, launch_columns(c));
        kern<<<grid, threads, c->smemBytes, c->stream>>>(c->P, c->dListAll.p, c->laLo, c->laHi, lambdaIterate,
                                                         upOnly, storeDepth, 0, nullptr);
        CU(cudaGetLastError());
        c->lastLaunches += 1;
    }
    if (!capturing)
    {
        CU(cudaEventRecord(c->evK1, c->stream));
        c->kernelTimed = true;
    }
    return 0;
}

// Nspace > 128: the multi-warp ray kernel for everything (Gamma iteration and formal solution)
template <int SOLVER, int MODE>
int launch_fs_long(LwB200Context* c, int lambdaIterate, int upOnly, int storeDepth)
{
    if (!c->evK0)
    {
        CU(cudaEventCreate(&c->evK0));
        CU(cudaEventCreate(&c->evK1));
    }
    const bool capturing = stream_is_capturing(c->stream);
    if (!capturing)
        CU(cudaEventRecord(c->evK0, c->stream));
    const int fsMode = MODE == MODE_ITER ? c->stokesFsMode : (upOnly ? 3 : 1);
    const int K = c->prob.Nspace;
    const bool big = K > 1024; // (the moment pipeline's kernels stop at 8 warps per column)
    if (big && fsMode != 0 && MODE == MODE_ITER)
        return fail("the full-Stokes formal solution is limited to Nspace <= 1024");
    const bool prdPass = MODE == MODE_ITER && c->customLists; // (PRD sub-iteration: the whole spectrum, PRD rates only)
    const bool all = (c->forceDirect && !prdPass) || (big && !prdPass);
    if (!all && !big
        && launch_pipeline<4, SOLVER, true>(c, c->customLists ? c->prdPl : full_lists(c), lambdaIterate, storeDepth, fsMode))
        return 1;
    // the wavelengths the moment pipeline does not carry (more than three overlapping lines, hybrid PRD) --
    // or, on request (LWB200_GENERAL_KERNEL) and beyond 1024 depths, all of them -- go through the general
    // multi-warp kernel
    const bool maskPass = prdPass && !c->prdPl.directPrdOnly;         // (angle-averaged: the masked wavelengths)
    const bool allTiles = all || big;
    const int nTiles = allTiles ? c->nListAll : (maskPass ? c->prdPl.nDirectPrd : c->nListDirect);
    if (nTiles > 0
        && launch_general_long<SOLVER>(c, allTiles ? c->dListAll.p : (maskPass ? c->prdPl.directPrd : c->dListDirect.p), nTiles,
                                       prdPass ? 0 : c->laLo, prdPass ? c->prob.Nspect : c->laHi, lambdaIterate, upOnly,
                                       storeDepth, maskPass ? 2 : (prdPass ? 1 : 0), MODE == MODE_ITER ? 0 : 1,
                                       maskPass ? c->prdPl.laMask : nullptr))
        return 1;
    if (!capturing)
    {
        CU(cudaEventRecord(c->evK1, c->stream));
        c->kernelTimed = true;
    }
    return 0;
}

template <int NCH, int MODE>
int launch_fs_s(LwB200Context* c, int li, int uo, int sd)
{
    switch (c->stokesFsMode ? (int)LWB200_FS_BEZIER3 : c->prob.formalSolver) // (full Stokes is always Bezier3)
    {
#ifndef LWB200_DEV_FAST_BUILD // (variant timing builds instantiate NCH = 3 / bezier3 only)
    case LWB200_FS_LINEAR: return launch_fs_t<NCH, 0, MODE>(c, li, uo, sd);
    case LWB200_FS_BESSER: return launch_fs_t<NCH, 1, MODE>(c, li, uo, sd);
#endif
    case LWB200_FS_BEZIER3: return launch_fs_t<NCH, 2, MODE>(c, li, uo, sd);
    }
    return fail("unknown formal solver");
}

template <int MODE>
int launch_fs(LwB200Context* c, int lambdaIterate, int upOnly, int storeDepth)
{
    if (refresh_tile_lists(c))
        return 1;
    if (c->prob.Nspace > 128)
    {
        switch (c->stokesFsMode ? (int)LWB200_FS_BEZIER3 : c->prob.formalSolver)
        {
#ifndef LWB200_DEV_FAST_BUILD
        case LWB200_FS_LINEAR: return launch_fs_long<0, MODE>(c, lambdaIterate, upOnly, storeDepth);
        case LWB200_FS_BESSER: return launch_fs_long<1, MODE>(c, lambdaIterate, upOnly, storeDepth);
#endif
        case LWB200_FS_BEZIER3: return launch_fs_long<2, MODE>(c, lambdaIterate, upOnly, storeDepth);
        }
        return fail("unknown formal solver");
    }
    switch (c->NCH)
    {
#ifndef LWB200_DEV_FAST_BUILD
    case 1: return launch_fs_s<1, MODE>(c, lambdaIterate, upOnly, storeDepth);
    case 2: return launch_fs_s<2, MODE>(c, lambdaIterate, upOnly, storeDepth);
#endif
    case 3: return launch_fs_s<3, MODE>(c, lambdaIterate, upOnly, storeDepth);
#ifndef LWB200_DEV_FAST_BUILD
    case 4: return launch_fs_s<4, MODE>(c, lambdaIterate, upOnly, storeDepth);
#endif
    }
    return fail("unexpected depth layout");
}

int copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
           cudaMemcpyKind kind, cudaStream_t s)
{
    if (width == 0 || height == 0)
        return 0;
    if (height == 1 || (dpitch == width && spitch == width))
    {
        CU(cudaMemcpyAsync(dst, src, width * height, kind, s));
        return 0;
    }
    CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s));
    return 0;
}

int grid_for(size_t total, int block = 256)
{
    size_t g = (total + block - 1) / block;
    return (int)std::max<size_t>(1, std::min<size_t>(g, 148 * 32));
}
} // namespace

// Hybrid PRD: device copies of the caller's tables (LwB200HybridPrd).  The plan was made for the lines and the
// wavelength set of the tables the context was created with; new tables (a changed velocity field) must name
// the same lines and stay within the planned set.
static int upload_hybrid(LwB200Context* c, const LwB200HybridPrd* h)
{
    const LwB200Problem& p = c->prob;
    const size_t K = p.Nspace, M = p.Nrays, L = p.Nspect, ncol = p.Ncol;
    if (!h || !h->prdLaOfLa || !h->hPrdLaOfLa || !h->JCoeffOff || h->NprdLa < 1 || h->NhPrd < 1)
        return fail("hybrid PRD: incomplete tables");
    for (size_t col = 0; col < ncol; ++col)
        for (size_t la = 0; la < L; ++la)
            if (h->hPrdLaOfLa[col * L + la] >= 0 && !c->hprdPlanMask[la])
                return fail("hybrid PRD: the new tables scatter from wavelengths outside the planned set; "
                            "create a new context");
    int nl = 0;
    size_t rhoTot = 0;
    for (size_t g = 0; g < c->trans.size(); ++g)
        nl += c->devTrans[g].hprdOff >= 0 ? 1 : 0;
    if (nl != h->Nlines)
        return fail("hybrid PRD: the tables name other lines than the context was created with");
    for (int q = 0; q < h->Nlines; ++q)
    {
        bool found = false;
        for (size_t g = 0; g < c->trans.size(); ++g)
            if (c->trans[g].atom == h->lineAtom[q] && c->trans[g].kr == h->lineTrans[q])
            {
                found = c->devTrans[g].hprdOff == h->rhoCoefOff[q];
                rhoTot = std::max<size_t>(rhoTot, (size_t)h->rhoCoefOff[q]
                                                       + ncol * (size_t)(c->devTrans[g].Nred - c->devTrans[g].Nblue) * M * 2 * K);
            }
        if (!found)
            return fail("hybrid PRD: the tables name other lines than the context was created with");
    }
    const size_t nRows = ncol * (size_t)h->NhPrd * M * 2 * K;
    const size_t nnz = (size_t)h->JCoeffOff[nRows];
    DevBuf<int>* ib[] = {&c->dHprdLaOfLa, &c->dPrdLaOfLa, &c->dJCoeffIdx, &c->dHprdI0};
    for (auto* b : ib)
        b->release();
    c->dJCoeffOff.release();
    c->dJCoeffFrac.release();
    c->dHprdFrac.release();
    if (c->dHprdLaOfLa.upload(std::vector<int>(h->hPrdLaOfLa, h->hPrdLaOfLa + ncol * L))
        || c->dPrdLaOfLa.upload(std::vector<int>(h->prdLaOfLa, h->prdLaOfLa + L))
        || c->dJCoeffOff.upload(std::vector<long long>(h->JCoeffOff, h->JCoeffOff + nRows + 1))
        || c->dJCoeffIdx.upload(std::vector<int>(h->JCoeffIdx, h->JCoeffIdx + nnz))
        || c->dJCoeffFrac.upload(std::vector<double>(h->JCoeffFrac, h->JCoeffFrac + nnz))
        || c->dHprdFrac.upload(std::vector<double>(h->rhoFrac, h->rhoFrac + rhoTot))
        || c->dHprdI0.upload(std::vector<int>(h->rhoI0, h->rhoI0 + rhoTot)))
        return 1;
    if (c->dJRest.n != ncol * (size_t)h->NprdLa * K)
    {
        c->dJRest.release();
        if (c->dJRest.alloc(ncol * (size_t)h->NprdLa * K))
            return 1;
    }
    CU(cudaMemset(c->dJRest.p, 0, c->dJRest.n * sizeof(double)));
    c->hprdHost = *h;
    c->prob.hprd = &c->hprdHost;
    c->hybrid = true;
    DevProblem& P = c->P;
    P.hprdLaOfLa = c->dHprdLaOfLa.p; P.prdLaOfLa = c->dPrdLaOfLa.p; P.JRest = c->dJRest.p;
    P.JCoeffOff = c->dJCoeffOff.p; P.JCoeffIdx = c->dJCoeffIdx.p; P.JCoeffFrac = c->dJCoeffFrac.p;
    P.hprdFrac = c->dHprdFrac.p; P.hprdI0 = c->dHprdI0.p;
    P.NprdLa = h->NprdLa; P.NhPrd = h->NhPrd;
    return 0;
}

extern "C"
{
const char* lwb200_last_error(void) { return g_err.c_str(); }
int lwb200_abi_version(void) { return LWB200_ABI_VERSION; }

int64_t lwb200_global_launch_count(void) { return g_totalLaunches.load(); }

int lwb200_device_count(int* count)
{
    CU(cudaGetDeviceCount(count));
    return 0;
}

int lwb200_create(const LwB200Problem* problem, int device, LwB200Context** out)
{
    if (!problem || !out)
        return fail("lwb200_create: null argument");
    if (problem->abiVersion != LWB200_ABI_VERSION)
        return fail("lwb200_create: ABI version mismatch");
    if (problem->Nspace < 3 || problem->Nrays < 1 || problem->Nspect < 1 || problem->Ncol < 1 || problem->Natom < 1)
        return fail("lwb200_create: bad dimensions");
    if (problem->Nspace > 4096)
        return fail("lwb200_create: Nspace > 4096 is not supported");
    if (problem->formalSolver < 0 || problem->formalSolver > 2)
        return fail("lwb200_create: formalSolver must be 0 (linear), 1 (besser) or 2 (bezier3)");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev < 1)
        return fail("lwb200_create: no CUDA device (this back end has no CPU fallback)");
    if (device < 0 || device >= ndev)
        return fail("lwb200_create: device index out of range");
    CU(cudaSetDevice(device));
    auto* c = new LwB200Context();
    c->device = device;
    c->prob = *problem;
    c->atoms.assign(problem->atoms, problem->atoms + problem->Natom);
    c->atomTrans.resize(problem->Natom);
    for (int a = 0; a < problem->Natom; ++a)
    {
        const LwB200Atom& at = problem->atoms[a];
        if (at.Nlevel < 1 || at.Nlevel > 64)
        {
            delete c;
            return fail("lwb200_create: Nlevel must be in [1, 64]");
        }
        c->atomTrans[a].assign(at.trans, at.trans + at.Ntrans);
        c->atoms[a].trans = c->atomTrans[a].data();
    }
    c->prob.atoms = c->atoms.data();
    c->laLo = 0;
    c->laHi = problem->Nspect;
    if (build_plan(c) || (problem->hprd && upload_hybrid(c, problem->hprd)))
    {
        lwb200_destroy(c);
        return 1;
    }
    // pin the two per-iteration outputs J and I in place so that their device->host copies are
    // true async DMA -- both, whatever their size: a copy into pageable memory blocks the host until
    // the stream reaches it, which would serialise the launch of everything behind the early fetch
    // (LWB200_FETCH_EARLY).  Failure to pin is not an error (the early fetch is then skipped).
    {
        const size_t nJ = (size_t)problem->Ncol * problem->Nspect * problem->Nspace * sizeof(double);
        const size_t nI = (size_t)problem->Ncol * problem->Nspect * problem->Nrays * sizeof(double);
        if (nJ > 0 && cudaHostRegister(problem->J, nJ, cudaHostRegisterDefault) == cudaSuccess)
            c->registered.push_back(problem->J);
        if (nI > 0 && cudaHostRegister(problem->I, nI, cudaHostRegisterDefault) == cudaSuccess)
            c->registered.push_back(problem->I);
        c->outputsPinned = c->registered.size() == 2;
        cudaGetLastError();
    }
    // Column stacks: the per-iteration arrays (populations, Gamma, rates, nStar ...) are many per-atom /
    // per-transition host buffers of [Ncol][rows][Nspace]; packing them into pinned staging and scattering
    // them back costs more host time than the copies themselves (measured: 1024 columns, 8 ms for Gamma +
    // rates of which 2 ms on the bus).  Pin them in place instead and let strided (2D) DMA copies do the
    // packing.  Small problems keep the staging path (one copy per group beats dozens of tiny ones).
    {
        const size_t K = problem->Nspace, ncol = problem->Ncol;
        size_t total = 0;
        for (int a = 0; a < problem->Natom; ++a)
            total += ncol * K * sizeof(double) * ((size_t)c->atoms[a].Nlevel * (c->atoms[a].Nlevel + 2) + 2 + 2 * c->atoms[a].Ntrans);
        bool ok = total >= ((size_t)8 << 20) && !std::getenv("LWB200_NO_PIN_IO");
        auto pin = [&](const void* ptr, size_t bytes) {
            if (!ok || !ptr || bytes == 0)
                return;
            if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess)
                c->registered.push_back(const_cast<void*>(ptr));
            else
                ok = false;
        };
        for (int a = 0; a < problem->Natom && ok; ++a)
        {
            const LwB200Atom& at = c->atoms[a];
            const size_t nk = ncol * at.Nlevel * K * sizeof(double);
            pin(at.n, nk);
            pin(at.nStar, nk);
            pin(at.nTotal, ncol * K * sizeof(double));
            pin(at.vBroad, ncol * K * sizeof(double));
            if (!at.detailedStatic)
                pin(at.Gamma, nk * at.Nlevel);
            for (int kr = 0; kr < at.Ntrans && ok; ++kr)
            {
                pin(c->atomTrans[a][kr].Rij, ncol * K * sizeof(double));
                pin(c->atomTrans[a][kr].Rji, ncol * K * sizeof(double));
            }
        }
        c->ioPinned = ok;
        cudaGetLastError();
    }
    *out = c;
    return 0;
}

int lwb200_destroy(LwB200Context* c)
{
    if (!c)
        return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf<double>* dbl[] = {&c->height, &c->temperature, &c->vlosMu, &c->muz, &c->wmu, &c->wavelength,
                             &c->chiBg, &c->etaBg, &c->scaBg, &c->J, &c->I, &c->n, &c->nStar, &c->nTotal,
                             &c->vBroad, &c->gRatio, &c->phi, &c->wphi, &c->rhoPrd, &c->aDamp, &c->wlambdaTab,
                             &c->alphaTab, &c->transWave, &c->lowerBcData, &c->upperBcData, &c->accum,
                             &c->prefill, &c->gamma, &c->dJ, &c->depthChi, &c->depthEta, &c->depthI, &c->djOut, &c->collC};
    for (auto* b : dbl)
        b->release();
    DevBuf<int>* ints[] = {&c->lowerBcIdx, &c->upperBcIdx, &c->dLaOff, &c->dLaCnt, &c->dTileLambda, &c->dLaHasLine, &c->dTileLa,
                           &c->dTileSlotOff, &c->dTileSlotTrans, &c->dAtomNlevel, &c->dAtomLevOff,
                           &c->dAtomGammaOff, &c->dAtomDetailed, &c->dSingular, &c->dPhiAsym};
    for (auto* b : ints)
        b->release();
    c->dJ20.release(); c->dJ20dag.release();
    c->dHprdLaOfLa.release(); c->dPrdLaOfLa.release(); c->dJCoeffIdx.release(); c->dHprdI0.release();
    c->dJCoeffOff.release(); c->dJRest.release(); c->dJCoeffFrac.release(); c->dHprdFrac.release();
    for (void* r : c->registered)
        cudaHostUnregister(r);
    Pinned* pins[] = {&c->stN, &c->stNStar, &c->stNTotal, &c->stVBroad, &c->stPrefill, &c->stGamma,
                      &c->stNOut, &c->stGammaOut, &c->stRates};
    for (auto* q : pins)
        q->release();
    if (c->evFork)
    {
        cudaEventDestroy(c->evFork);
        cudaEventDestroy(c->evRays);
        cudaEventDestroy(c->evCopy);
        cudaEventDestroy(c->evDj);
        cudaStreamDestroy(c->copyStream);
        for (int q = 0; q < 4; ++q)
        {
            cudaEventDestroy(c->evJoin[q]);
            cudaEventDestroy(c->evRaysKind[q]);
            cudaStreamDestroy(c->sideStream[q]);
        }
    }
    if (c->hs)
        cudaFreeHost(c->hs);
    if (c->ngHostMax)
        cudaFreeHost(c->ngHostMax);
    if (c->ngHostIdx)
        cudaFreeHost(c->ngHostIdx);
    if (c->evK0)
        cudaEventDestroy(c->evK0);
    if (c->evK1)
        cudaEventDestroy(c->evK1);
    for (int q = 0; q < 4; ++q)
        c->dKindLam[q].release();
    c->dListMoment.release();
    c->dListMomentPrd.release();
    c->dPrdMask.release();
    c->dPrdLines.release();
    c->qelast.release();
    c->cmat.release();
    c->rhoPrev.release();
    c->nOld.release();
    c->nrScratch.release();
    c->nrDC.release();
    c->nrPrev.release();
    c->nrStages.release();
    c->nrBgNe.release();
    c->neDev.release();
    c->nrAtoms.release();
    c->pol.release();
    c->Quv.release();
    c->Jdag.release();
    c->dPolLam.release();
    for (int q = 0; q < 4; ++q)
        c->dKindLamUnpol[q].release();
    c->prdMax.release();
    c->prdIdx.release();
    for (int q = 0; q < 4; ++q)
        c->dKindLamPrd[q].release();
    c->dMomOff.release();
    c->dLaNLines.release();
    c->dLamLine.release();
    c->chiC.release();
    c->etaC.release();
    c->mom.release();
    c->dListDirect.release();
    c->dListAll.release();
    c->djIdx.release();
    c->ngRing.release();
    c->ngMax.release();
    c->ngIdx.release();
    c->dColList.release();
    c->dColActive.release();
    c->djPartIdx.release();
    c->djPartMax.release();
    c->djTicket.release();
    c->dTrans.release();
    c->dEntries.release();
    c->dLines.release();
    c->zUp.release();
    c->zDown.release();
    c->dGEntries.release();
    c->dGLam.release();
    c->dGLine.release();
    c->dGTileLa.release();
    c->dGTileSlotOff.release();
    c->dGTileSlotRows.release();
    c->dGList.release();
    c->dGListPrd.release();
    c->dGLamOfLa.release();
    delete c;
    return 0;
}

int lwb200_set_stream(LwB200Context* c, void* stream)
{
    c->stream = (cudaStream_t)stream;
    return 0;
}

int lwb200_set_lambda_range(LwB200Context* c, int32_t laStart, int32_t laEnd)
{
    if (laStart < 0 || laEnd > c->prob.Nspect || laStart > laEnd)
        return fail("lwb200_set_lambda_range: bad range");
    c->laLo = laStart;
    c->laHi = laEnd;
    return 0;
}

int lwb200_set_zplane(LwB200Context* c, double* zPlaneUp, double* zPlaneDown)
{
    CU(cudaSetDevice(c->device));
    const size_t count = (size_t)c->prob.Ncol * c->prob.Nspect * c->prob.Nrays;
    // kernels in flight may still write through the previous pointers
    CU(cudaStreamSynchronize(c->stream));
    if (zPlaneUp && !c->zUp.p && c->zUp.alloc(count))
        return 1;
    if (zPlaneDown && !c->zDown.p && c->zDown.alloc(count))
        return 1;
    c->zUpHost = zPlaneUp;
    c->zDownHost = zPlaneDown;
    c->P.zPlaneUp = zPlaneUp ? c->zUp.p : nullptr;
    c->P.zPlaneDown = zPlaneDown ? c->zDown.p : nullptr;
    return 0;
}

int lwb200_set_active_columns(LwB200Context* c, const uint8_t* active)
{
    CU(cudaSetDevice(c->device));
    // the kernels in flight still read the previous list
    CU(cudaStreamSynchronize(c->stream));
    c->dColList.release();
    c->dColActive.release();
    c->P.colList = nullptr;
    c->P.colActive = nullptr;
    c->nActiveCol = -1;
    if (!active)
        return 0;
    std::vector<int> list;
    std::vector<unsigned char> mask(c->prob.Ncol);
    for (int col = 0; col < c->prob.Ncol; ++col)
    {
        mask[col] = active[col] ? 1 : 0;
        if (active[col])
            list.push_back(col);
    }
    if ((int)list.size() == c->prob.Ncol)
        return 0; // nothing retired
    if ((!list.empty() && c->dColList.upload(list)) || c->dColActive.upload(mask))
        return 1;
    c->P.colList = c->dColList.p;
    c->P.colActive = c->dColActive.p;
    c->nActiveCol = (int)list.size();
    return 0;
}

int lwb200_ng_configure(LwB200Context* c, int32_t Norder, int32_t Nperiod, int32_t Ndelay)
{
    CU(cudaSetDevice(c->device));
    if (Norder < 0 || Norder > kNgMaxOrder)
        return fail("lwb200_ng_configure: Norder must be 0.." + std::to_string(kNgMaxOrder));
    if (Norder > 0 && Nperiod < 1)
        return fail("lwb200_ng_configure: Nperiod must be >= 1");
    Ndelay = std::max(Ndelay, Nperiod + 2); // Ng.hpp:34
    if (Norder > 0 && Ndelay < Norder + 2)
        return fail("lwb200_ng_configure: Ndelay < Norder + 2 (the reference reads outside its history there)");
    const int R = Norder + 2;
    c->ngRing.release();
    c->ngMax.release();
    c->ngIdx.release();
    if (c->ngRing.alloc((size_t)R * c->n.n) || c->ngMax.alloc((size_t)c->prob.Natom * (kNgParts + 1)) || c->ngIdx.alloc((size_t)c->prob.Natom * (kNgParts + 1)))
        return 1;
    if (!c->ngHostMax)
    {
        CU(cudaHostAlloc((void**)&c->ngHostMax, std::max(1, c->prob.Natom) * sizeof(double), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&c->ngHostIdx, std::max(1, c->prob.Natom) * sizeof(long long), cudaHostAllocDefault));
    }
    if (ensure_host_scalars(c))
        return 1;
    // the constructor keeps the current populations as the first solution (Ng.hpp:31-41)
    CU(cudaMemsetAsync(c->ngRing.p, 0, c->ngRing.n * sizeof(double), c->stream));
    CU(cudaMemcpyAsync(c->ngRing.p, c->n.p, c->n.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ngOrder = Norder;
    c->ngPeriod = Nperiod;
    c->ngDelay = Ndelay;
    c->ngCount = 1;
    return 0;
}

int lwb200_ng_clear(LwB200Context* c)
{
    CU(cudaSetDevice(c->device));
    if (c->ngOrder < 0)
        return fail("lwb200_ng_clear: lwb200_ng_configure has not been called");
    CU(cudaMemsetAsync(c->ngRing.p, 0, c->ngRing.n * sizeof(double), c->stream));
    c->ngCount = 0;
    return 0;
}

int lwb200_ng_accelerate(LwB200Context* c, int32_t* accelerated, double* dMax, int64_t* dMaxIdx)
{
    CU(cudaSetDevice(c->device));
    if (c->ngOrder < 0)
        return fail("lwb200_ng_accelerate: lwb200_ng_configure has not been called");
    const int R = c->ngOrder + 2, Natom = c->prob.Natom;
    const size_t stride = c->n.n;
    c->lastLaunches = 0;
    // accelerate(): store the new solution (Ng.hpp:62-65)
    CU(cudaMemcpyAsync(c->ngRing.p + (size_t)(c->ngCount % R) * stride, c->n.p, stride * sizeof(double),
                       cudaMemcpyDeviceToDevice, c->stream));
    c->ngCount += 1;
    const bool go = c->ngOrder > 0 && c->ngCount >= c->ngDelay && ((c->ngCount - c->ngDelay) % c->ngPeriod) == 0;
    if (accelerated)
        *accelerated = go ? 1 : 0;
    // dMax == dMaxIdx == NULL: no host synchronisation; the results travel with the stream and are read
    // with lwb200_last_ng after the next lwb200_sync (singular systems: lwb200_last_singular)
    const bool async = !dMax && !dMaxIdx;
    int* dSingular = c->dSingular.p;
    if (async)
    {
        if (!c->singularPending)
            c->hs->nSingular = 0;
        c->singularPending = true;
        dSingular = &c->hsDev->nSingular;
    }
    else
        CU(cudaMemsetAsync(c->dSingular.p, 0, sizeof(int), c->stream));
    if (go)
    {
        ng_accelerate_kernel<<<dim3(Natom, c->prob.Ncol), 256, 0, c->stream>>>(c->P, c->ngRing.p, R, stride, c->ngCount,
                                                                              c->ngOrder, c->n.p, dSingular);
        CU(cudaGetLastError());
        c->lastLaunches += 1;
    }
    if (c->ngCount < 2)
    {
        // nothing to compare yet: zeros, written by the stream like every other result (the pinned host words may
        // still be the target of the previous call's copy, so the host does not touch them)
        CU(cudaMemsetAsync(c->ngMax.p, 0, c->ngMax.n * sizeof(double), c->stream));
        CU(cudaMemsetAsync(c->ngIdx.p, 0, c->ngIdx.n * sizeof(long long), c->stream));
    }
    else
    {
        // max_change(): the last two stored solutions (Ng.hpp:138-156)
        const size_t perAtom = (size_t)c->prob.Ncol * c->P.maxNlevel * c->P.K;
        const int nParts = (int)std::min<size_t>(kNgParts, std::max<size_t>(1, perAtom / 4096));
        ng_max_change_kernel<<<dim3(Natom, nParts), 256, 0, c->stream>>>(
            c->P, c->ngRing.p + (size_t)((c->ngCount - 1) % R) * stride,
            c->ngRing.p + (size_t)((c->ngCount - 2) % R) * stride, c->ngMax.p, c->ngIdx.p);
        CU(cudaGetLastError());
        ng_max_change_final_kernel<<<Natom, kNgParts, 0, c->stream>>>(nParts, c->ngMax.p, c->ngIdx.p);
        CU(cudaGetLastError());
        c->lastLaunches += 2;
    }
    CU(cudaMemcpy2DAsync(c->ngHostMax, sizeof(double), c->ngMax.p, (kNgParts + 1) * sizeof(double), sizeof(double),
                         Natom, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(c->ngHostIdx, sizeof(long long), c->ngIdx.p, (kNgParts + 1) * sizeof(long long),
                         sizeof(long long), Natom, cudaMemcpyDeviceToHost, c->stream));
    if (async)
        return 0;
    int ns = 0;
    CU(cudaMemcpyAsync(&ns, c->dSingular.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int a = 0; a < Natom; ++a)
    {
        if (dMax)
            dMax[a] = c->ngHostMax[a];
        if (dMaxIdx)
            dMaxIdx[a] = c->ngHostIdx[a];
    }
    if (ns > 0)
        return fail("Singular Matrix");
    return 0;
}

int lwb200_last_ng(LwB200Context* c, double* dMax, int64_t* dMaxIdx)
{
    if (c->ngOrder < 0 || !c->ngHostMax)
        return fail("lwb200_last_ng: lwb200_ng_configure has not been called");
    for (int a = 0; a < c->prob.Natom; ++a)
    {
        if (dMax)
            dMax[a] = c->ngHostMax[a];
        if (dMaxIdx)
            dMaxIdx[a] = c->ngHostIdx[a];
    }
    return 0;
}

int lwb200_sync(LwB200Context* c)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    size_t bytes = 0;
    for (const Pending& q : c->pending)
        bytes += q.bytes;
    const Pending* pend = c->pending.data();
    host_parallel_for(c->pending.size(), bytes, [pend](size_t b, size_t e) {
        for (size_t q = b; q < e; ++q)
            std::memcpy(pend[q].dst, pend[q].src, pend[q].bytes);
    });
    c->pending.clear();
    return 0;
}

} // extern "C"

// Per-iteration arrays (populations, Gamma, rates) are many small per-atom /
// per-transition host buffers.  They cross PCIe as ONE packed copy per group
// through pinned staging buffers that mirror the packed device layouts; the
// host-side scatter of downloads runs in lwb200_sync once the stream is done.
static int stage_ready(Pinned& st, size_t count)
{
    if (st.n < count)
    {
        if (st.p)
            cudaFreeHost(st.p);
        st.p = nullptr;
        cudaError_t e = cudaMallocHost((void**)&st.p, count * sizeof(double));
        if (e != cudaSuccess)
            return fail(std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        st.n = count;
    }
    if (!st.ev)
        CU(cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming));
    if (st.inFlight)
    {
        CU(cudaEventSynchronize(st.ev));
        st.inFlight = false;
    }
    return 0;
}

template <typename RowsFn, typename OffFn, typename PtrFn>
static int upload_packed(LwB200Context* c, Pinned& st, double* dev, int totalRows, RowsFn rows, OffFn off,
                         PtrFn ptr)
{
    const size_t K = c->prob.Nspace, ncol = c->prob.Ncol;
    const size_t total = ncol * (size_t)totalRows * K;
    if (c->ioPinned)
    {
        // the host arrays are pinned in place: one strided DMA copy per atom packs them on the way
        for (int a = 0; a < c->prob.Natom; ++a)
        {
            const double* src = ptr(a);
            const size_t r = rows(a);
            if (!src || r == 0)
                continue;
            CU(cudaMemcpy2DAsync(dev + off(a) * K, (size_t)totalRows * K * sizeof(double), src, r * K * sizeof(double),
                                 r * K * sizeof(double), ncol, cudaMemcpyHostToDevice, c->stream));
        }
        return 0;
    }
    if (stage_ready(st, total))
        return 1;
    for (int a = 0; a < c->prob.Natom; ++a)
    {
        const double* src = ptr(a);
        const size_t r = rows(a);
        if (!src || r == 0)
            continue;
        double* dst = st.p + off(a) * K;
        host_parallel_for(ncol, ncol * r * K * sizeof(double), [=](size_t b, size_t e) {
            for (size_t col = b; col < e; ++col)
                std::memcpy(dst + col * totalRows * K, src + col * r * K, r * K * sizeof(double));
        });
    }
    CU(cudaMemcpyAsync(dev, st.p, total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(st.ev, c->stream));
    st.inFlight = true;
    return 0;
}

extern "C"
{
int lwb200_upload(LwB200Context* c, uint32_t mask)
{
    CU(cudaSetDevice(c->device));
    const LwB200Problem& p = c->prob;
    const size_t K = p.Nspace, L = p.Nspect, M = p.Nrays, ncol = p.Ncol;
    const size_t D = sizeof(double);
    cudaStream_t s = c->stream;
    const auto H2D = cudaMemcpyHostToDevice;
    c->lastLaunches = 0;
    if (mask & LWB200_ATMOS)
    {
        CU(cudaMemcpyAsync(c->height.p, p.height, ncol * K * D, H2D, s));
        CU(cudaMemcpyAsync(c->temperature.p, p.temperature, ncol * K * D, H2D, s));
        if (p.vlosMu)
            CU(cudaMemcpyAsync(c->vlosMu.p, p.vlosMu, ncol * M * K * D, H2D, s));
        if (p.lowerBc == LWB200_BC_CALLABLE)
            CU(cudaMemcpyAsync(c->lowerBcData.p, p.lowerBcData, ncol * L * p.NlowerBcMu * D, H2D, s));
        if (p.upperBc == LWB200_BC_CALLABLE)
            CU(cudaMemcpyAsync(c->upperBcData.p, p.upperBcData, ncol * L * p.NupperBcMu * D, H2D, s));
    }
    if (mask & LWB200_BACKGR)
    {
        CU(cudaMemcpyAsync(c->chiBg.p, p.chiBg, ncol * L * K * D, H2D, s));
        CU(cudaMemcpyAsync(c->etaBg.p, p.etaBg, ncol * L * K * D, H2D, s));
        CU(cudaMemcpyAsync(c->scaBg.p, p.scaBg, ncol * L * K * D, H2D, s));
    }
    if (mask & LWB200_JBAR)
        CU(cudaMemcpyAsync(c->J.p, p.J, ncol * L * K * D, H2D, s));
    auto levRows = [&](int a) { return (size_t)c->atoms[a].Nlevel; };
    auto levOff = [&](int a) { return (size_t)c->atomLevOff[a]; };
    auto one = [&](int) { return (size_t)1; };
    auto atomIdx = [&](int a) { return (size_t)a; };
    auto gamRows = [&](int a) {
        return c->atoms[a].detailedStatic ? (size_t)0 : (size_t)c->atoms[a].Nlevel * c->atoms[a].Nlevel;
    };
    auto gamOff = [&](int a) { return (size_t)c->atomGammaOff[a]; };
    if (mask & LWB200_POPS)
        if (upload_packed(c, c->stN, c->n.p, c->P.NlevTot, levRows, levOff,
                          [&](int a) { return (const double*)c->atoms[a].n; }))
            return 1;
    if (mask & LWB200_NSTAR)
    {
        if (upload_packed(c, c->stNStar, c->nStar.p, c->P.NlevTot, levRows, levOff,
                          [&](int a) { return c->atoms[a].nStar; }))
            return 1;
        if (upload_packed(c, c->stNTotal, c->nTotal.p, p.Natom, one, atomIdx,
                          [&](int a) { return c->atoms[a].nTotal; }))
            return 1;
        if (upload_packed(c, c->stVBroad, c->vBroad.p, p.Natom, one, atomIdx,
                          [&](int a) { return c->atoms[a].vBroad; }))
            return 1;
        if (c->P.Ncont > 0)
        {
            const size_t total = (size_t)c->P.Ncont * ncol * K;
            ratio_kernel<<<grid_for(total), 256, 0, s>>>(c->P, c->gRatio.p);
            CU(cudaGetLastError());
            c->lastLaunches += 1;
        }
        c->nstarUploaded = true;
    }
    if ((mask & LWB200_GAMMA) && c->P.GammaTot > 0)
    {
        for (int a = 0; a < p.Natom; ++a)
            if (!c->atoms[a].detailedStatic && !c->atoms[a].Gamma)
                return fail("active atom without Gamma buffer");
        if (upload_packed(c, c->stPrefill, c->prefill.p, c->P.GammaTot, gamRows, gamOff,
                          [&](int a) { return (const double*)c->atoms[a].Gamma; }))
            return 1;
        c->prefillFromC = false; // (an uploaded prefill supersedes lwb200_set_collision_prefill)
    }
    if ((mask & LWB200_COLLISIONS) && c->P.GammaTot > 0)
    {
        // C of every active atom: with lwb200_set_collision_prefill the device makes Gamma's prefill
        // crsw*C itself, and neither the host product nor its upload is paid per iteration
        if (!c->collC.p)
        {
            if (c->collC.alloc(ncol * (size_t)c->P.GammaTot * K))
                return fail("out of device memory (collision rates)");
            CU(cudaMemsetAsync(c->collC.p, 0, c->collC.n * D, s)); // (an atom without C contributes nothing)
        }
        if (upload_packed(c, c->stPrefill, c->collC.p, c->P.GammaTot, gamRows, gamOff,
                          [&](int a) { return c->atoms[a].detailedStatic ? nullptr : (const double*)c->atoms[a].C; }))
            return 1;
    }
    if ((mask & LWB200_GAMMA_FINAL) && c->P.GammaTot > 0)
    {
        if (upload_packed(c, c->stGamma, c->gamma.p, c->P.GammaTot, gamRows, gamOff,
                          [&](int a) { return (const double*)c->atoms[a].Gamma; }))
            return 1;
    }
    if ((mask & LWB200_ADAMP) && !(mask & LWB200_PROFILE))
    {
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const DevTrans& d = c->devTrans[g];
            const LwB200Transition& t = c->trans[g].t;
            if (d.type != 0)
                continue;
            if (!t.aDamp)
                return fail("line without aDamp");
            CU(cudaMemcpyAsync(c->aDamp.p + (size_t)d.lineIdx * ncol * K, t.aDamp, ncol * K * D, H2D, s));
        }
    }
    if (mask & LWB200_PROFILE)
    {
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const DevTrans& d = c->devTrans[g];
            const LwB200Transition& t = c->trans[g].t;
            if (d.type != 0)
                continue;
            if (!t.phi || !t.wphi)
                return fail("LWB200_PROFILE upload of a line without host phi/wphi (use lwb200_compute_profiles)");
            const size_t Nl = d.Nred - d.Nblue;
            CU(cudaMemcpyAsync(c->phi.p + d.phiOff, t.phi, ncol * Nl * M * 2 * K * D, H2D, s));
            CU(cudaMemcpyAsync(c->wphi.p + (size_t)d.lineIdx * ncol * K, t.wphi, ncol * K * D, H2D, s));
            if (t.aDamp)
                CU(cudaMemcpyAsync(c->aDamp.p + (size_t)d.lineIdx * ncol * K, t.aDamp, ncol * K * D, H2D, s));
            if (d.rhoOff >= 0)
                CU(cudaMemcpyAsync(c->rhoPrd.p + d.rhoOff, t.rhoPrd, ncol * Nl * K * D, H2D, s));
        }
        if (check_phi_symmetry(c))
            return 1;
    }
    if (mask & LWB200_RATES)
    {
        // host Rij / Rji -> the rate rows of the accumulator (only lwb200_redistribute_prd reads
        // rates on the device; a caller that changed them on the host sends them back with this)
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const DevTrans& d = c->devTrans[g];
            const LwB200Transition& t = c->trans[g].t;
            if (copy2d(c->accum.p + (size_t)d.accRij * K, (size_t)c->P.AccTot * K * D, t.Rij, K * D, K * D, ncol, H2D, s)
                || copy2d(c->accum.p + (size_t)d.accRji * K, (size_t)c->P.AccTot * K * D, t.Rji, K * D, K * D, ncol, H2D, s))
                return 1;
        }
    }
    if ((mask & LWB200_STOKES) && c->polTot > 0)
    {
        for (size_t g = 0; g < c->trans.size(); ++g)
            if (c->transPolOff[g] >= 0)
            {
                if (!c->trans[g].t.polProfiles)
                    return fail("LWB200_STOKES upload of a polarised line without host profiles "
                                "(use lwb200_compute_polarised_profiles)");
                CU(cudaMemcpyAsync(c->pol.p + c->transPolOff[g], c->trans[g].t.polProfiles,
                                   (size_t)6 * c->devTrans[g].phiColStride * ncol * D, H2D, s));
            }
        c->stokesUploaded = true;
    }
    if ((mask & LWB200_PRD) && !c->prdLines.empty())
    {
        for (size_t q = 0; q < c->prdLines.size(); ++q)
        {
            const DevPrdLine& ln = c->prdLines[q];
            const DevTrans& d = c->devTrans[ln.trans];
            const LwB200Transition& t = c->trans[ln.trans].t;
            if (!t.Qelast || !t.aDamp || !t.rhoPrd)
                return fail("LWB200_PRD upload: PRD line without Qelast / aDamp / rhoPrd");
            CU(cudaMemcpyAsync(c->qelast.p + q * ncol * K, t.Qelast, ncol * K * D, H2D, s));
            CU(cudaMemcpyAsync(c->aDamp.p + (size_t)d.lineIdx * ncol * K, t.aDamp, ncol * K * D, H2D, s));
            CU(cudaMemcpyAsync(c->rhoPrd.p + d.rhoOff, t.rhoPrd, ncol * (size_t)ln.Nl * K * D, H2D, s));
        }
        for (int a = 0; a < p.Natom; ++a)
        {
            if (c->atomCOff[a] < 0)
                continue;
            const size_t n2k = (size_t)c->atoms[a].Nlevel * c->atoms[a].Nlevel * K;
            if (copy2d(c->cmat.p + c->atomCOff[a], (size_t)c->cTot * D, c->atoms[a].C, n2k * D, n2k * D, ncol,
                       H2D, s))
                return 1;
        }
        if (c->hybrid && c->hprdHost.JRest)
            CU(cudaMemcpyAsync(c->dJRest.p, c->hprdHost.JRest, c->dJRest.n * D, H2D, s));
        c->prdUploaded = true;
    }
    return 0;
}

int lwb200_download(LwB200Context* c, uint32_t mask)
{
    CU(cudaSetDevice(c->device));
    const LwB200Problem& p = c->prob;
    const size_t K = p.Nspace, L = p.Nspect, M = p.Nrays, ncol = p.Ncol;
    const size_t D = sizeof(double);
    cudaStream_t s = c->stream;
    const auto D2H = cudaMemcpyDeviceToHost;
    if (c->fetched && (mask & (LWB200_JBAR | LWB200_INTENS)))
    {
        // already on their way (LWB200_FETCH_EARLY): order the stream after that copy
        CU(cudaStreamWaitEvent(s, c->evCopy, 0));
        c->fetched = false;
        mask &= ~(uint32_t)(LWB200_JBAR | LWB200_INTENS);
        if (c->fetchedGR)
            mask &= ~(uint32_t)(LWB200_GAMMA | LWB200_RATES);
        c->fetchedGR = false;
    }
    if ((mask & LWB200_STOKES) && c->polTot > 0)
        CU(cudaMemcpyAsync(p.Quv, c->Quv.p, ncol * 3 * L * M * D, D2H, s));
    if ((mask & LWB200_POLPROF) && c->polTot > 0)
        for (size_t g = 0; g < c->trans.size(); ++g)
            if (c->transPolOff[g] >= 0 && c->trans[g].t.polProfiles)
                CU(cudaMemcpyAsync(const_cast<double*>(c->trans[g].t.polProfiles), c->pol.p + c->transPolOff[g],
                                   (size_t)6 * c->devTrans[g].phiColStride * ncol * D, D2H, s));
    if (mask & LWB200_PRD)
    {
        for (const DevPrdLine& ln : c->prdLines)
        {
            const LwB200Transition& t = c->trans[ln.trans].t;
            CU(cudaMemcpyAsync(t.rhoPrd, c->rhoPrd.p + ln.rhoOff, ncol * (size_t)ln.Nl * K * D, D2H, s));
        }
        if (c->hybrid && c->hprdHost.JRest)
            CU(cudaMemcpyAsync(c->hprdHost.JRest, c->dJRest.p, c->dJRest.n * D, D2H, s));
    }
    // rows [r0, r1) of every column's [L][.] arrays: the context's own wavelength range on request
    const size_t r0 = (mask & LWB200_OWN_ROWS) ? (size_t)c->laLo : 0, r1 = (mask & LWB200_OWN_ROWS) ? (size_t)c->laHi : L;
    if (mask & LWB200_JBAR)
        if (copy2d(p.J + r0 * K, L * K * D, c->J.p + r0 * K, L * K * D, (r1 - r0) * K * D, ncol, D2H, s))
            return 1;
    if (mask & LWB200_INTENS)
        if (copy2d(p.I + r0 * M, L * M * D, c->I.p + r0 * M, L * M * D, (r1 - r0) * M * D, ncol, D2H, s))
            return 1;
    if ((mask & LWB200_POPS) && c->ioPinned)
    {
        const size_t rows = c->P.NlevTot;
        for (int a = 0; a < p.Natom; ++a)
        {
            const size_t N = c->atoms[a].Nlevel;
            CU(cudaMemcpy2DAsync(c->atoms[a].n, N * K * D, c->n.p + (size_t)c->atomLevOff[a] * K, rows * K * D, N * K * D,
                                 ncol, D2H, s));
        }
    }
    else if (mask & LWB200_POPS)
    {
        const size_t rows = c->P.NlevTot;
        if (stage_ready(c->stNOut, ncol * rows * K))
            return 1;
        CU(cudaMemcpyAsync(c->stNOut.p, c->n.p, ncol * rows * K * D, D2H, s));
        for (int a = 0; a < p.Natom; ++a)
        {
            const size_t N = c->atoms[a].Nlevel;
            for (size_t col = 0; col < ncol; ++col)
                c->pending.push_back({c->atoms[a].n + col * N * K,
                                      c->stNOut.p + (col * rows + c->atomLevOff[a]) * K, N * K * D});
        }
    }
    if ((mask & LWB200_GAMMA) && c->P.GammaTot > 0 && c->ioPinned)
    {
        const size_t rows = c->P.GammaTot;
        for (int a = 0; a < p.Natom; ++a)
        {
            if (c->atoms[a].detailedStatic)
                continue;
            const size_t N2 = (size_t)c->atoms[a].Nlevel * c->atoms[a].Nlevel;
            CU(cudaMemcpy2DAsync(c->atoms[a].Gamma, N2 * K * D, c->gamma.p + (size_t)c->atomGammaOff[a] * K, rows * K * D,
                                 N2 * K * D, ncol, D2H, s));
        }
    }
    else if ((mask & LWB200_GAMMA) && c->P.GammaTot > 0)
    {
        const size_t rows = c->P.GammaTot;
        if (stage_ready(c->stGammaOut, ncol * rows * K))
            return 1;
        CU(cudaMemcpyAsync(c->stGammaOut.p, c->gamma.p, ncol * rows * K * D, D2H, s));
        for (int a = 0; a < p.Natom; ++a)
        {
            if (c->atoms[a].detailedStatic)
                continue;
            const size_t N2 = (size_t)c->atoms[a].Nlevel * c->atoms[a].Nlevel;
            for (size_t col = 0; col < ncol; ++col)
                c->pending.push_back({c->atoms[a].Gamma + col * N2 * K,
                                      c->stGammaOut.p + (col * rows + c->atomGammaOff[a]) * K, N2 * K * D});
        }
    }
    if ((mask & LWB200_RATES) && c->ioPinned)
    {
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const LwB200Transition& t = c->trans[g].t;
            const DevTrans& d = c->devTrans[g];
            CU(cudaMemcpy2DAsync(t.Rij, K * D, c->accum.p + (size_t)d.accRij * K, (size_t)c->P.AccTot * K * D, K * D, ncol,
                                 D2H, s));
            CU(cudaMemcpy2DAsync(t.Rji, K * D, c->accum.p + (size_t)d.accRji * K, (size_t)c->P.AccTot * K * D, K * D, ncol,
                                 D2H, s));
        }
    }
    else if (mask & LWB200_RATES)
    {
        // the R rows are the tail of each column's accumulator block
        const size_t rows = 2 * c->trans.size();
        if (stage_ready(c->stRates, ncol * rows * K))
            return 1;
        if (copy2d(c->stRates.p, rows * K * D, c->accum.p + (size_t)c->P.GammaTot * K, (size_t)c->P.AccTot * K * D,
                   rows * K * D, ncol, D2H, s))
            return 1;
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const LwB200Transition& t = c->trans[g].t;
            for (size_t col = 0; col < ncol; ++col)
            {
                c->pending.push_back({t.Rij + col * K, c->stRates.p + (col * rows + 2 * g) * K, K * D});
                c->pending.push_back({t.Rji + col * K, c->stRates.p + (col * rows + 2 * g + 1) * K, K * D});
            }
        }
    }
    if (mask & LWB200_ZPLANE)
    {
        if (c->zUpHost)
            CU(cudaMemcpyAsync(c->zUpHost, c->zUp.p, ncol * L * M * D, D2H, s));
        if (c->zDownHost && c->zDownWritten) // (an up-only formal solution leaves ZPlaneDown as it is)
            CU(cudaMemcpyAsync(c->zDownHost, c->zDown.p, ncol * L * M * D, D2H, s));
    }
    if ((mask & LWB200_DEPTH) && c->depthChi.p)
    {
        const size_t nd = ncol * L * M * 2 * K * D;
        CU(cudaMemcpyAsync(p.depthChi, c->depthChi.p, nd, D2H, s));
        CU(cudaMemcpyAsync(p.depthEta, c->depthEta.p, nd, D2H, s));
        CU(cudaMemcpyAsync(p.depthI, c->depthI.p, nd, D2H, s));
    }
    if (mask & LWB200_PROFILE)
    {
        for (size_t g = 0; g < c->trans.size(); ++g)
        {
            const DevTrans& d = c->devTrans[g];
            const LwB200Transition& t = c->trans[g].t;
            if (d.type != 0 || !t.phi)
                continue;
            const size_t Nl = d.Nred - d.Nblue;
            CU(cudaMemcpyAsync(t.phi, c->phi.p + d.phiOff, ncol * Nl * M * 2 * K * D, D2H, s));
            CU(cudaMemcpyAsync(t.wphi, c->wphi.p + (size_t)d.lineIdx * ncol * K, ncol * K * D, D2H, s));
        }
    }
    return 0;
}

int lwb200_compute_profiles(LwB200Context* c)
{
    CU(cudaSetDevice(c->device));
    if (!c->vlosMu.p)
        return fail("lwb200_compute_profiles: the problem has no vlosMu");
    if (c->devLines.empty())
        return 0;
    c->lastLaunches = 0;
    int64_t nLaunched = 0;
    int rc = launch_profiles(c->P, c->dLines.p, (int)c->devLines.size(), c->devLines.data(), c->transWave.p,
                             c->wlambdaTab.p, c->aDamp.p, c->vBroad.p, c->vlosMu.p, c->phi.p, c->wphi.p,
                             c->stream, &nLaunched);
    c->lastLaunches += nLaunched;
    if (rc)
        return fail(std::string("lwb200_compute_profiles: ") + cudaGetErrorString((cudaError_t)rc));
    return check_phi_symmetry(c);
}

int lwb200_compute_polarised_profiles(LwB200Context* c, const LwB200Zeeman* zl, int32_t nLines, const double* B,
                                      const double* cosGamma, const double* cos2chi, const double* sin2chi)
{
    CU(cudaSetDevice(c->device));
    if (!zl || nLines < 1 || !B || !cosGamma || !cos2chi || !sin2chi)
        return fail("lwb200_compute_polarised_profiles: null argument");
    if (!c->vlosMu.p)
        return fail("lwb200_compute_polarised_profiles: the problem has no vlosMu");
    const LwB200Problem& p = c->prob;
    const size_t K = p.Nspace, M = p.Nrays, ncol = p.Ncol;
    cudaStream_t s = c->stream;
    std::vector<int> alpha;
    std::vector<double> shift, strength;
    std::vector<DevZeeman> dz;
    for (int q = 0; q < nLines; ++q)
    {
        int g = -1;
        for (size_t t = 0; t < c->trans.size(); ++t)
            if (c->trans[t].atom == zl[q].atom && c->trans[t].kr == zl[q].trans)
                g = (int)t;
        if (g < 0 || c->devTrans[g].type != 0 || c->transPolOff[g] < 0)
            return fail("lwb200_compute_polarised_profiles: not a line declared polarised at lwb200_create");
        if (zl[q].Ncomponent < 1 || !zl[q].alpha || !zl[q].shift || !zl[q].strength)
            return fail("lwb200_compute_polarised_profiles: empty Zeeman pattern");
        DevZeeman z{};
        z.line = -1;
        for (size_t l = 0; l < c->devLines.size(); ++l)
            if (c->devLines[l].lineIdx == c->devTrans[g].lineIdx)
                z.line = (int)l;
        z.nComp = zl[q].Ncomponent;
        z.compOff = (int)alpha.size();
        z.polOff = c->transPolOff[g];
        z.polArr = c->devTrans[g].phiColStride * (long long)ncol;
        alpha.insert(alpha.end(), zl[q].alpha, zl[q].alpha + z.nComp);
        shift.insert(shift.end(), zl[q].shift, zl[q].shift + z.nComp);
        strength.insert(strength.end(), zl[q].strength, zl[q].strength + z.nComp);
        dz.push_back(z);
    }
    DevBuf<int> dAlpha;
    DevBuf<double> dShift, dStrength, dB, dAng;
    auto cleanup = [&]() {
        dAlpha.release(); dShift.release(); dStrength.release(); dB.release(); dAng.release();
    };
    if (dAlpha.upload(alpha) || dShift.upload(shift) || dStrength.upload(strength) || dB.alloc(ncol * K)
        || dAng.alloc(3 * ncol * M * K))
    {
        cleanup();
        return 1;
    }
    const size_t na = ncol * M * K;
    cudaError_t e = cudaMemcpyAsync(dB.p, B, ncol * K * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dAng.p, cosGamma, na * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dAng.p + na, cos2chi, na * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dAng.p + 2 * na, sin2chi, na * sizeof(double), cudaMemcpyHostToDevice, s);
    c->lastLaunches = 0;
    for (size_t q = 0; q < dz.size() && e == cudaSuccess; ++q)
    {
        const size_t total = (size_t)c->devLines[dz[q].line].Nl * M * 2 * K * ncol;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 127) / 128, 148 * 64));
        pol_profile_kernel<<<grid, 128, 0, s>>>(c->P, c->dLines.p, dz[q], dAlpha.p, dShift.p, dStrength.p, c->transWave.p,
                                                c->aDamp.p, c->vBroad.p, c->vlosMu.p, dB.p, dAng.p, dAng.p + na,
                                                dAng.p + 2 * na, c->phi.p, c->pol.p);
        e = cudaGetLastError();
        c->lastLaunches += 1;
    }
    if (e == cudaSuccess)
    {
        // wphi of every line from the profiles now on the device (compute_wphi order, FormalScalar.cpp:106-134)
        const size_t total = c->devLines.size() * ncol * K;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 127) / 128, 148 * 32));
        wphi_kernel<<<grid, 128, 0, s>>>(c->P, c->dLines.p, (int)c->devLines.size(), c->wlambdaTab.p, c->phi.p, c->wphi.p);
        e = cudaGetLastError();
        c->lastLaunches += 1;
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s); // (the temporaries go away below)
    cleanup();
    if (e != cudaSuccess)
        return fail(std::string("lwb200_compute_polarised_profiles: ") + cudaGetErrorString(e));
    c->stokesUploaded = true;
    return check_phi_symmetry(c);
}

int lwb200_set_j20(LwB200Context* c, double* J20)
{
    CU(cudaSetDevice(c->device));
    if ((J20 != nullptr) != (c->j20Host != nullptr))
        c->stokesLists = false; // (with J20 every wavelength goes through the Stokes solver)
    c->j20Host = J20;
    if (J20 && !c->dJ20.p)
    {
        const size_t n = (size_t)c->prob.Ncol * c->prob.Nspect * c->prob.Nspace;
        if (c->dJ20.alloc(n) || c->dJ20dag.alloc(n))
            return fail("out of device memory (J20)");
    }
    return 0;
}

int lwb200_set_hybrid_prd(LwB200Context* c, const LwB200HybridPrd* tables)
{
    CU(cudaSetDevice(c->device));
    if (!c->hybrid)
        return fail("lwb200_set_hybrid_prd: the context was created without hybrid PRD tables (LwB200Problem::hprd)");
    CU(cudaStreamSynchronize(c->stream));
    return upload_hybrid(c, tables);
}

int lwb200_set_collision_prefill(LwB200Context* c, int enable, double crsw)
{
    if (enable && !c->collC.p && c->P.GammaTot > 0)
        return fail("lwb200_set_collision_prefill: the collisional rates have not been uploaded (LWB200_COLLISIONS)");
    c->prefillFromC = enable != 0;
    c->crswC = crsw;
    return 0;
}

int lwb200_finalise(LwB200Context* c)
{
    CU(cudaSetDevice(c->device));
    if (c->P.GammaTot > 0)
    {
        const size_t total = (size_t)c->P.Ncol * c->P.Natom * c->P.maxNlevel * c->P.K;
        if (launch_finalise(c, 0, c->P.Ncol))
            return 1;
    }
    return 0;
}


// dJ reduction whose result lands in pinned host memory with the stream (no host synchronisation)
static int dj_max_async(LwB200Context* c)
{
    if (ensure_host_scalars(c))
        return 1;
    if (c->djDone)
    {
        // already reduced beside the Gamma stage
        c->djDone = false;
        CU(cudaStreamWaitEvent(c->stream, c->evDj, 0));
        return 0;
    }
    if (launch_dj_reduce(c, c->stream, c->laLo, c->laHi, nullptr))
        return 1;
    CU(cudaMemcpyAsync(&c->hs->dJ, c->djOut.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&c->hs->dJIdx, c->djIdx.p, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

int lwb200_last_dj(LwB200Context* c, double* dJMax, int64_t* dJMaxIdx)
{
    if (!c->hs)
        return fail("lwb200_last_dj: no asynchronous dJ reduction has been requested");
    if (dJMax)
        *dJMax = c->hs->dJ;
    if (dJMaxIdx)
        *dJMaxIdx = c->hs->dJIdx;
    return 0;
}

int lwb200_last_singular(LwB200Context* c, int32_t* nSingular)
{
    if (!c->hs)
        return fail("lwb200_last_singular: no asynchronous population update has been requested");
    c->singularPending = false;
    if (nSingular)
        *nSingular = c->hs->nSingular;
    if (c->hs->nSingular > 0)
        return fail("Singular Matrix");
    return 0;
}

int lwb200_dj_max(LwB200Context* c, double* dJMax, int64_t* dJMaxIdx)
{
    CU(cudaSetDevice(c->device));
    if (c->djDone)
    {
        c->djDone = false;
        CU(cudaStreamWaitEvent(c->stream, c->evDj, 0));
        CU(cudaStreamSynchronize(c->stream));
        if (dJMax)
            *dJMax = c->hs->dJ;
        if (dJMaxIdx)
            *dJMaxIdx = c->hs->dJIdx;
        return 0;
    }
    if (launch_dj_reduce(c, c->stream, c->laLo, c->laHi, nullptr))
        return 1;
    double m = 0.0;
    long long idx = 0;
    CU(cudaMemcpyAsync(&m, c->djOut.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&idx, c->djIdx.p, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (dJMax)
        *dJMax = m;
    if (dJMaxIdx)
        *dJMaxIdx = idx;
    return 0;
}

int lwb200_fs_iter(LwB200Context* c, uint32_t flags, double* dJMax, int64_t* dJMaxIdx)
{
    CU(cudaSetDevice(c->device));
    if (!c->nstarUploaded)
        return fail("lwb200_fs_iter: inputs have not been uploaded (lwb200_upload)");
    const int storeDepth = (flags & LWB200_STORE_DEPTH) ? 1 : 0;
    c->forceDirect = (flags & LWB200_GENERAL_KERNEL) != 0 || c->prob.Nspace > 1024; // (beyond 1024 depths: always)
    // (a wavelength shard or a masked stack owns only part of J / I: no wholesale early copy there)
    c->fetchEarly = (flags & LWB200_FETCH_EARLY) != 0 && !c->forceDirect && c->outputsPinned && c->nActiveCol < 0
                    && c->laLo == 0 && c->laHi == c->prob.Nspect;
    if (c->fetched)
    {
        // an early copy nobody collected: it must not race with this iteration's writes of J
        CU(cudaStreamWaitEvent(c->stream, c->evCopy, 0));
        c->fetched = false;
    }
    // column stacks with their per-iteration host arrays pinned in place: Gamma is finalised and, with the
    // rates, sent home batch by batch (two or more batches; a single batch has nothing to hide them under)
    c->finaliseEarly = c->fetchEarly && c->ioPinned && !(flags & LWB200_DEFER_FINALISE) && c->prob.Ncol > c->batchCols;
    c->finalisedEarly = c->fetchedGR = false;
    if (storeDepth && !c->depthChi.p)
        return fail("lwb200_fs_iter: STORE_DEPTH without depth arrays in the problem");
    c->lastLaunches = 0;
    const bool wantDj = (flags & LWB200_DJ_ASYNC) || dJMax || dJMaxIdx;
    c->djDone = false;
    c->djEarly = wantDj && !c->forceDirect && c->nListDirect == 0;
    if (c->djEarly && ensure_host_scalars(c))
        return 1;
    // zero_rates + fresh partial sums (:605-612, :643)
    if (c->nActiveCol < 0)
        CU(cudaMemsetAsync(c->accum.p, 0, c->accum.n * sizeof(double), c->stream));
    else if (c->nActiveCol > 0)
    {
        zero_accum_kernel<<<grid_for((size_t)c->nActiveCol * c->P.AccTot * c->P.K), 256, 0, c->stream>>>(c->P, c->nActiveCol);
        CU(cudaGetLastError());
        c->lastLaunches += 1;
    }
    else
        return fail("lwb200_fs_iter: every column is retired");
    c->zDownWritten = true;
    if (c->hybrid) // zero_Gamma_rates_JRest (SimdFullIterationTemplates.hpp:521-586, :602-603)
        CU(cudaMemsetAsync(c->dJRest.p, 0, c->dJRest.n * sizeof(double), c->stream));
    const int rcFs = launch_fs<MODE_ITER>(c, (flags & LWB200_LAMBDA_ITERATE) ? 1 : 0, 0, storeDepth);
    c->djEarly = false;
    if (rcFs)
        return 1;
    if (!(flags & LWB200_DEFER_FINALISE) && !c->finalisedEarly)
        if (lwb200_finalise(c))
            return 1;
    if (flags & LWB200_DJ_ASYNC)
        return dj_max_async(c);
    if (dJMax || dJMaxIdx)
        return lwb200_dj_max(c, dJMax, dJMaxIdx);
    return 0;
}

int lwb200_formal_sol(LwB200Context* c, int upOnly)
{
    CU(cudaSetDevice(c->device));
    if (!c->nstarUploaded)
        return fail("lwb200_formal_sol: inputs have not been uploaded (lwb200_upload)");
    c->lastLaunches = 0;
    c->zDownWritten = !upOnly;
    return launch_fs<MODE_FS>(c, 0, upOnly ? 1 : 0, 0);
}

static int population_update(LwB200Context* c, int32_t atom, int32_t kStart, int32_t kEnd, int32_t* nSingular,
                             const double* nOldHost, double dt, const char* who, bool async = false)
{
    CU(cudaSetDevice(c->device));
    const int K = c->prob.Nspace;
    if (kStart < 0 && kEnd < 0)
    {
        kStart = 0;
        kEnd = K;
    }
    if (kStart < 0 || kEnd > K || kStart >= kEnd)
        return fail(std::string(who) + ": bad depth range");
    if (atom >= c->prob.Natom)
        return fail(std::string(who) + ": atom index out of range");
    c->lastLaunches = 0;
    // asynchronous updates count singular systems cumulatively until lwb200_last_singular collects them
    int* dSingular = c->dSingular.p;
    if (async)
    {
        if (ensure_host_scalars(c))
            return 1;
        if (!c->singularPending)
            c->hs->nSingular = 0;
        dSingular = &c->hsDev->nSingular;
    }
    else
        CU(cudaMemsetAsync(c->dSingular.p, 0, sizeof(int), c->stream));
    c->singularPending = async;
    const double* nOldDev = nullptr;
    if (nOldHost)
    {
        if (atom < 0 || c->atoms[atom].detailedStatic)
            return fail(std::string(who) + ": needs one active atom");
        const size_t count = (size_t)c->prob.Ncol * c->atoms[atom].Nlevel * K;
        if (c->nOld.n < count)
        {
            c->nOld.release();
            if (c->nOld.alloc(count))
                return 1;
        }
        CU(cudaMemcpyAsync(c->nOld.p, nOldHost, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        nOldDev = c->nOld.p;
    }
    {
        int maxN = 1, nActive = 0;
        for (int a = 0; a < c->prob.Natom; ++a)
        {
            if ((atom >= 0 && a != atom) || c->atoms[a].detailedStatic)
                continue;
            maxN = std::max(maxN, c->atoms[a].Nlevel);
            ++nActive;
        }
        if (nActive > 0)
        {
            const size_t total = (size_t)(atom >= 0 ? 1 : c->prob.Natom) * c->prob.Ncol * (kEnd - kStart);
            if (maxN <= 8)
                stat_eq_kernel<8><<<grid_for(total, 64), 64, 0, c->stream>>>(
                    c->P, atom, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, dSingular, nOldDev, dt);
            else if (maxN <= 16)
                stat_eq_kernel<16><<<grid_for(total, 64), 64, 0, c->stream>>>(
                    c->P, atom, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, dSingular, nOldDev, dt);
            else if (maxN <= 32)
                stat_eq_kernel<32><<<grid_for(total, 64), 64, 0, c->stream>>>(
                    c->P, atom, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, dSingular, nOldDev, dt);
            else
            {
                // 33..64 levels: matrices in a global scratch block, one slice per thread of a bounded grid
                const int blocks = std::min(grid_for(total, 64), 128);
                if (!c->statEqScratch.p && c->statEqScratch.alloc((size_t)blocks * 64 * 2 * 64 * 64))
                    return fail("out of device memory (stat-eq scratch)");
                stat_eq_kernel<64><<<blocks, 64, 0, c->stream>>>(c->P, atom, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p,
                                                               dSingular, nOldDev, dt, c->statEqScratch.p);
            }
            CU(cudaGetLastError());
            c->lastLaunches += 1;
        }
    }
    if (async)
        return 0;
    int ns = 0;
    CU(cudaMemcpyAsync(&ns, c->dSingular.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (nSingular)
        *nSingular = ns;
    if (ns > 0)
        return fail("Singular Matrix");
    return 0;
}

int lwb200_stat_eq(LwB200Context* c, int32_t atom, int32_t kStart, int32_t kEnd, int32_t* nSingular)
{
    return population_update(c, atom, kStart, kEnd, nSingular, nullptr, 0.0, "lwb200_stat_eq");
}

int lwb200_stat_eq_async(LwB200Context* c, int32_t atom, int32_t kStart, int32_t kEnd)
{
    return population_update(c, atom, kStart, kEnd, nullptr, nullptr, 0.0, "lwb200_stat_eq_async", true);
}

int lwb200_population_solve(int device, int32_t Ncol, int32_t Nlevel, int32_t Nspace, const double* Gamma,
                            double* n, const double* nTotal, const double* nOld, double dt, int32_t kStart,
                            int32_t kEnd, int32_t* nSingular)
{
    if (nSingular)
        *nSingular = 0;
    if (!Gamma || !n || (!nTotal && !nOld) || Ncol < 1 || Nlevel < 1 || Nlevel > 32 || Nspace < 1)
        return fail("lwb200_population_solve: bad arguments");
    if (kStart < 0 && kEnd < 0)
    {
        kStart = 0;
        kEnd = Nspace;
    }
    if (kStart < 0 || kEnd > Nspace || kStart >= kEnd)
        return fail("lwb200_population_solve: bad depth range");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev < 1)
        return fail("lwb200_population_solve: no CUDA device (this back end has no CPU fallback)");
    if (device < 0 || device >= ndev)
        return fail("lwb200_population_solve: device index out of range");
    CU(cudaSetDevice(device));
    const size_t nk = (size_t)Ncol * Nlevel * Nspace, n2k = nk * Nlevel, D = sizeof(double);
    DevBuf<double> dG, dN, dTot, dOld;
    DevBuf<int> dMeta;
    int rc = 1;
    do
    {
        // one atom: Nlevel, level offset 0, Gamma offset 0, not detailed; then the singular counter
        const std::vector<int> meta = {Nlevel, 0, 0, 0, 0};
        if (dG.alloc(n2k) || dN.alloc(nk) || dTot.alloc((size_t)Ncol * Nspace) || (nOld && dOld.alloc(nk))
            || dMeta.upload(meta))
            break;
        if (cudaMemcpy(dG.p, Gamma, n2k * D, cudaMemcpyHostToDevice) != cudaSuccess
            || cudaMemcpy(dN.p, n, nk * D, cudaMemcpyHostToDevice) != cudaSuccess
            || (nTotal && cudaMemcpy(dTot.p, nTotal, (size_t)Ncol * Nspace * D, cudaMemcpyHostToDevice) != cudaSuccess)
            || (nOld && cudaMemcpy(dOld.p, nOld, nk * D, cudaMemcpyHostToDevice) != cudaSuccess))
        {
            fail("lwb200_population_solve: host to device copy failed");
            break;
        }
        DevProblem P{};
        P.Ncol = Ncol;
        P.K = Nspace;
        P.Natom = 1;
        P.NlevTot = Nlevel;
        P.GammaTot = Nlevel * Nlevel;
        P.atomNlevel = dMeta.p;
        P.atomLevOff = dMeta.p + 1;
        P.atomGammaOff = dMeta.p + 2;
        P.atomDetailed = dMeta.p + 3;
        int* dSing = dMeta.p + 4;
        const size_t total = (size_t)Ncol * (kEnd - kStart);
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 63) / 64, 148 * 32));
        if (Nlevel <= 8)
            stat_eq_kernel<8><<<grid, 64>>>(P, 0, kStart, kEnd, dG.p, dN.p, dTot.p, dSing, nOld ? dOld.p : nullptr, dt);
        else if (Nlevel <= 16)
            stat_eq_kernel<16><<<grid, 64>>>(P, 0, kStart, kEnd, dG.p, dN.p, dTot.p, dSing, nOld ? dOld.p : nullptr, dt);
        else
            stat_eq_kernel<32><<<grid, 64>>>(P, 0, kStart, kEnd, dG.p, dN.p, dTot.p, dSing, nOld ? dOld.p : nullptr, dt);
        g_totalLaunches += 1;
        int ns = 0;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpy(&ns, dSing, sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(n, dN.p, nk * D, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
        {
            fail(std::string("lwb200_population_solve: ") + cudaGetErrorString(e));
            break;
        }
        if (nSingular)
            *nSingular = ns;
        rc = ns > 0 ? fail("Singular Matrix") : 0;
    } while (false);
    dG.release();
    dN.release();
    dTot.release();
    dOld.release();
    dMeta.release();
    return rc;
}

int lwb200_time_dep_update(LwB200Context* c, int32_t atom, const double* nOld, double dt, int32_t kStart,
                           int32_t kEnd, int32_t* nSingular)
{
    if (!nOld)
        return fail("lwb200_time_dep_update: nOld is NULL");
    return population_update(c, atom, kStart, kEnd, nSingular, nOld, dt, "lwb200_time_dep_update");
}

// redistribute_prd_lines_template (PrdTemplates.hpp:164-291) on the device-resident state.
int lwb200_redistribute_prd(LwB200Context* c, int32_t maxIter, double tol, int32_t includeDetailed,
                            int32_t* nIterOut, double* dRho, int32_t* dRhoIdx, double* dJPrdMax,
                            int64_t* dJPrdMaxIdx)
{
    CU(cudaSetDevice(c->device));
    if (c->nActiveCol >= 0)
        return fail("lwb200_redistribute_prd: not available while a column mask is set (lwb200_set_active_columns)");
    if (nIterOut)
        *nIterOut = 0;
    // the lines taking part: active atoms' always, detailed ones on request
    int nLines = 0;
    for (size_t q = 0; q < c->prdLines.size(); ++q)
        if (!c->prdLineDetailed[q] || includeDetailed)
            nLines = (int)q + 1; // (active lines come first, so this is a prefix)
    if (nLines == 0)
        return 0;
    if (!c->nstarUploaded || !c->prdUploaded)
        return fail("lwb200_redistribute_prd: inputs have not been uploaded (lwb200_upload with LWB200_PRD)");
    for (int q = 0; q < nLines; ++q)
        if (c->prdLines[q].cOff < 0)
            return fail("lwb200_redistribute_prd: atom of a PRD line without collisional rates C");
    // On a wavelength shard the redistribution runs REPLICATED over the whole spectrum (it needs all of
    // J -- the caller all-gathers the J rows first, sharding.sharded_prd_redistribute -- and the PRD
    // wavelengths are a small subset): its work lists do not depend on the shard.
    if (refresh_tile_lists(c))
        return 1;
    const LwB200Problem& p = c->prob;
    const int K = p.Nspace, L = p.Nspect;
    cudaStream_t s = c->stream;
    c->forceDirect = false;
    c->fetchEarly = false;

    // wavelengths touched by a redistributed line (:225-240) and the work lists over them
    if (c->hybrid && c->prdListsFor != -2 - nLines)
    {
        // hybrid PRD: the formal solution runs over the wavelengths that scatter into the PRD grid
        // (idxsForFs = spect.hPrdIdxs, PrdTemplates.hpp:233-234), column by column, in the general kernel
        c->dPrdMask.release();
        if (c->dPrdMask.upload(c->hprdPlanMask))
            return 1;
        c->prdListsFor = -2 - nLines;
    }
    else if (!c->hybrid && c->prdListsFor != nLines)
    {
        std::vector<unsigned char> mask(L, 0);
        for (int q = 0; q < nLines; ++q)
            for (int la = c->prdLines[q].Nblue; la < c->prdLines[q].Nblue + c->prdLines[q].Nl; ++la)
                mask[la] = 1;
        // (a redistributed wavelength with more than three overlapping lines is solved by the general kernel,
        // like the rest of its kind: its tiles, masked)
        std::vector<int> kindLam[4], tiles, dtiles;
        for (int la = 0; la < L; ++la)
            if (mask[la] && c->laKind[la] < 4)
                kindLam[c->laKind[la]].push_back(la);
        for (int t = 0; t < c->Ntile; ++t)
        {
            bool any = false;
            for (int q = c->tileLa[t]; q < c->tileLa[t + 1]; ++q)
                any = any || mask[c->tileLambda[q]];
            if (any)
                (c->tileKind[t] < 4 ? tiles : dtiles).push_back(t);
        }
        std::vector<int> gtiles;
        for (int t = 0; t < c->nGTile; ++t)
        {
            bool any = false;
            for (int q = c->gTileLa[t]; q < c->gTileLa[t + 1]; ++q)
                any = any || mask[c->gLamLa[q]];
            if (any)
                gtiles.push_back(t);
        }
        c->dPrdMask.release();
        c->dListMomentPrd.release();
        c->dListDirectPrd.release();
        c->dGListPrd.release();
        if (c->dPrdMask.upload(mask) || c->dListMomentPrd.upload(tiles) || c->dGListPrd.upload(gtiles)
            || c->dListDirectPrd.upload(dtiles))
            return 1;
        c->nListDirectPrd = (int)dtiles.size();
        c->nListMomentPrd = (int)tiles.size();
        c->nGListPrd = (int)gtiles.size();
        for (int q = 0; q < 4; ++q)
        {
            c->dKindLamPrd[q].release();
            if (c->dKindLamPrd[q].upload(kindLam[q]))
                return 1;
            c->nKindLamPrd[q] = (int)kindLam[q].size();
        }
        c->prdListsFor = nLines;
    }
    PipelineLists pl{};
    if (!c->hybrid)
    {
        pl.moment = c->dListMomentPrd.p;
        pl.nMoment = c->nListMomentPrd;
        pl.directPrd = c->dListDirectPrd.p;
        pl.nDirectPrd = c->nListDirectPrd;
        pl.gTiles = c->dGListPrd.p;
        pl.nGTiles = c->nGListPrd;
        for (int q = 0; q < 4; ++q)
        {
            pl.kindLam[q] = c->dKindLamPrd[q].p;
            pl.nKindLam[q] = c->nKindLamPrd[q];
        }
    }
    pl.directPrdOnly = c->hybrid ? 1 : 0;
    pl.laMask = c->dPrdMask.p;
    pl.prdOnly = 1;
    pl.fullRange = 1;

    // Ng(0, 0, 0, rho): the change of the first redistribution is measured against the rho we start from
    CU(cudaMemcpyAsync(c->rhoPrev.p, c->rhoPrd.p, c->rhoPrd.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    int maxNl = 1;
    for (int q = 0; q < nLines; ++q)
        maxNl = std::max(maxNl, c->prdLines[q].Nl);
    std::vector<double> hMax(nLines);
    std::vector<int> hIdx(nLines);
    int iter = 0;
    c->lastLaunches = 0;
    while (iter < maxIter)
    {
        ++iter;
        prd_scatter_kernel<<<dim3((maxNl * K + 127) / 128, p.Ncol, nLines), 128, 0, s>>>(
            c->P, c->dPrdLines.p, c->transWave.p, c->qelast.p, c->cmat.p, c->cTot, c->vBroad.p, c->aDamp.p,
            c->rhoPrd.p);
        CU(cudaGetLastError());
        prd_change_kernel<<<nLines, 1024, 0, s>>>(c->dPrdLines.p, p.Ncol, K, c->rhoPrd.p, c->rhoPrev.p,
                                                 c->prdMax.p, c->prdIdx.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hMax.data(), c->prdMax.p, nLines * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(hIdx.data(), c->prdIdx.p, nLines * sizeof(int), cudaMemcpyDeviceToHost, s));
        // formal_sol_prd_update_rates (PrdTemplates.hpp:18-76)
        prd_zero_rates_kernel<<<grid_for((size_t)nLines * p.Ncol * K), 256, 0, s>>>(c->P, c->dPrdLines.p, nLines);
        CU(cudaGetLastError());
        c->lastLaunches += 3;
        if (c->hybrid)
        {
            // JRest is rebuilt by this pass (PrdTemplates.hpp:57-58); dJ of the wavelengths a column does not
            // visit must not carry over from the Gamma iteration
            CU(cudaMemsetAsync(c->dJRest.p, 0, c->dJRest.n * sizeof(double), s));
            CU(cudaMemsetAsync(c->dJ.p, 0, c->dJ.n * sizeof(double), s));
        }
        c->customLists = true;
        c->prdPl = pl;
        const int rc = launch_fs<MODE_ITER>(c, 0, 0, 0);
        c->customLists = false;
        if (rc)
            return 1;
        if (launch_dj_reduce(c, s, 0, L, c->dPrdMask.p))
            return 1;
        CU(cudaGetLastError());
        c->lastLaunches += 1;
        double m = 0.0;
        long long idx = 0;
        CU(cudaMemcpyAsync(&m, c->djOut.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&idx, c->djIdx.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        double dRhoMax = 0.0;
        for (int q = 0; q < nLines; ++q)
        {
            dRhoMax = std::max(dRhoMax, hMax[q]);
            if (dRho)
                dRho[(size_t)(iter - 1) * nLines + q] = hMax[q];
            if (dRhoIdx)
                dRhoIdx[(size_t)(iter - 1) * nLines + q] = hIdx[q];
        }
        if (dJPrdMax)
            dJPrdMax[iter - 1] = m;
        if (dJPrdMaxIdx)
            dJPrdMaxIdx[iter - 1] = idx % L;
        if (dRhoMax < tol)
            break;
    }
    if (nIterOut)
        *nIterOut = iter;
    return 0;
}


// formal_sol_full_stokes_impl (FormalStokes.cpp:664-723) on the device-resident state.
int lwb200_formal_sol_full_stokes(LwB200Context* c, int updateJ, int upOnly, double* dJMax, int64_t* dJMaxIdx)
{
    CU(cudaSetDevice(c->device));
    if (c->nActiveCol >= 0)
        return fail("lwb200_formal_sol_full_stokes: not available while a column mask is set (lwb200_set_active_columns)");
    if (!c->nstarUploaded)
        return fail("lwb200_formal_sol_full_stokes: inputs have not been uploaded (lwb200_upload)");
    const bool j20 = c->j20Host != nullptr;
    if (c->polTot == 0 && !j20)
        return fail("lwb200_formal_sol_full_stokes: the problem has no polarised line");
    if (c->polTot > 0 && !c->stokesUploaded)
        return fail("lwb200_formal_sol_full_stokes: polarised profiles have not been uploaded (LWB200_STOKES)");
    if (!c->Quv.p)
        return fail("lwb200_formal_sol_full_stokes: the problem has no Quv array");
    if (c->laLo != 0 || c->laHi != c->prob.Nspect)
        return fail("lwb200_formal_sol_full_stokes: not available on a wavelength shard (column-shard instead)");
    if (c->prob.Nspace < 3)
        return fail("lwb200_formal_sol_full_stokes: needs at least three depth points");
    if (refresh_tile_lists(c))
        return 1;
    const LwB200Problem& p = c->prob;
    const int L = p.Nspect;
    if (!c->stokesLists)
    {
        std::vector<int> pol, unpol[4];
        for (int la = 0; la < L; ++la)
        {
            bool isPol = false;
            for (size_t g = 0; g < c->trans.size(); ++g)
                if (c->transPolOff[g] >= 0 && la >= c->devTrans[g].Nblue && la < c->devTrans[g].Nred)
                    isPol = true;
            if (isPol || j20) // (J20: polarisedFrequency = true everywhere, FormalStokes.cpp:490)
            {
                if (c->laKind[la] >= 4)
                    return fail("lwb200_formal_sol_full_stokes: a polarised line overlaps more than two other lines");
                pol.push_back(la);
            }
            else if (c->laKind[la] < 4)
                unpol[c->laKind[la]].push_back(la);
            else
                return fail("lwb200_formal_sol_full_stokes: more than three overlapping lines at one wavelength");
        }
        if (c->dPolLam.upload(pol))
            return 1;
        c->nPolLam = (int)pol.size();
        for (int q = 0; q < 4; ++q)
        {
            if (c->dKindLamUnpol[q].upload(unpol[q]))
                return 1;
            c->nKindLamUnpol[q] = (int)unpol[q].size();
        }
        c->stokesLists = true;
    }
    cudaStream_t s = c->stream;
    c->forceDirect = false;
    c->fetchEarly = false;
    c->lastLaunches = 0;
    CU(cudaMemsetAsync(c->Quv.p, 0, c->Quv.n * sizeof(double), s));
    c->P.j20 = j20 ? 1 : 0;
    c->P.J20 = c->dJ20.p;
    c->P.J20dag = nullptr;
    if (j20)
    {
        // the caller's anisotropy is the J20-dagger of a J-updating pass, which rebuilds it from zero
        CU(cudaMemcpyAsync(c->dJ20.p, c->j20Host, c->dJ20.n * sizeof(double), cudaMemcpyHostToDevice, s));
        if (updateJ)
        {
            CU(cudaMemcpyAsync(c->dJ20dag.p, c->dJ20.p, c->dJ20.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
            CU(cudaMemsetAsync(c->dJ20.p, 0, c->dJ20.n * sizeof(double), s));
            c->P.J20dag = c->dJ20dag.p;
        }
    }
    if (updateJ)
    {
        // the polarised rays add into J atomically: keep J-dagger aside and clear their rows
        CU(cudaMemcpyAsync(c->Jdag.p, c->J.p, c->J.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
        stokes_zero_rows_kernel<<<dim3(c->nPolLam, p.Ncol), 128, 0, s>>>(c->P, c->dPolLam.p, c->nPolLam);
        CU(cudaGetLastError());
        c->lastLaunches += 1;
    }
    PipelineLists pl = full_lists(c);
    for (int q = 0; q < 4; ++q)
    {
        pl.kindLam[q] = c->dKindLamUnpol[q].p;
        pl.nKindLam[q] = c->nKindLamUnpol[q];
    }
    pl.polLam = c->dPolLam.p;
    pl.nPolLam = c->nPolLam;
    pl.fullRange = 1;
    c->customLists = true; // (custom lists; also keeps the general kernel out)
    c->prdPl = pl;
    c->stokesFsMode = (updateJ ? 4 : (1 | 8)) | (upOnly ? 2 : 0);
    const int rc = launch_fs<MODE_ITER>(c, 0, 0, 0);
    c->customLists = false;
    c->stokesFsMode = 0;
    c->P.j20 = 0;
    if (rc)
        return 1;
    if (j20 && updateJ) // (travels with the stream: in the caller's array after the next lwb200_sync)
        CU(cudaMemcpyAsync(c->j20Host, c->dJ20.p, c->dJ20.n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (updateJ)
    {
        stokes_dj_kernel<<<dim3(c->nPolLam, p.Ncol), 32, 0, s>>>(c->P, c->dPolLam.p, c->nPolLam);
        CU(cudaGetLastError());
        c->lastLaunches += 1;
        if (dJMax || dJMaxIdx)
            return lwb200_dj_max(c, dJMax, dJMaxIdx);
    }
    else
    {
        if (dJMax)
            *dJMax = 0.0;
        if (dJMaxIdx)
            *dJMaxIdx = 0;
    }
    return 0;
}


// nr_post_update_impl (UpdatePopulations.cpp:292-394) on the device-resident populations and Gamma.
int lwb200_nr_post_update(LwB200Context* c, const LwB200NrUpdate* u, int32_t kStart, int32_t kEnd,
                          int32_t* nSingular)
{
    CU(cudaSetDevice(c->device));
    if (c->nActiveCol >= 0)
        return fail("lwb200_nr_post_update: not available while a column mask is set (lwb200_set_active_columns)");
    const LwB200Problem& p = c->prob;
    const int K = p.Nspace;
    const size_t ncol = p.Ncol, D = sizeof(double);
    if (!u || u->Natom < 1 || !u->atomIdx || !u->backgroundNe)
        return fail("lwb200_nr_post_update: bad update description");
    if (!p.ne)
        return fail("lwb200_nr_post_update: the problem has no electron density (ne)");
    if (kStart < 0 && kEnd < 0)
    {
        kStart = 0;
        kEnd = K;
    }
    if (kStart < 0 || kEnd > K || kStart >= kEnd)
        return fail("lwb200_nr_post_update: bad depth range");
    cudaStream_t s = c->stream;
    std::vector<NrAtom> atoms;
    std::vector<double> stages;
    long long dcTot = 0, prevTot = 0;
    int Neqn = 1;
    for (int a = 0; a < u->Natom; ++a)
    {
        const int ia = u->atomIdx[a];
        if (ia < 0 || ia >= p.Natom || c->atoms[ia].detailedStatic)
            return fail("lwb200_nr_post_update: atom index out of range or detailed-static");
        const LwB200Atom& at = c->atoms[ia];
        if (!at.C || !at.stages)
            return fail("lwb200_nr_post_update: atom without C or stages");
        NrAtom na{};
        na.atom = ia;
        na.N = at.Nlevel;
        na.levOff = c->atomLevOff[ia];
        na.gammaOff = c->atomGammaOff[ia];
        na.transBeg = 0;
        for (size_t g = 0; g < c->trans.size(); ++g)
            if (c->trans[g].atom == ia)
            {
                na.transBeg = (int)g - c->trans[g].kr;
                break;
            }
        na.transEnd = na.transBeg + at.Ntrans;
        na.cOff = c->atomCOff[ia];
        na.dcOff = (u->dC && u->dC[a]) ? dcTot : -1;
        na.prevOff = prevTot;
        na.stageOff = (int)stages.size();
        dcTot += (long long)at.Nlevel * at.Nlevel * K;
        prevTot += (long long)at.Nlevel * K;
        for (int l = 0; l < at.Nlevel; ++l)
            stages.push_back(at.stages[l]);
        atoms.push_back(na);
        Neqn += at.Nlevel;
    }
    if (Neqn > 160)
        return fail("lwb200_nr_post_update: more than 159 levels in the coupled system");
    if (u->timeDependent && !u->nPrev)
        return fail("lwb200_nr_post_update: timeDependent without nPrev");
    // inputs: C of every atom with one (same packing as the PRD path), dC, nPrev, stages, backgroundNe, ne
    if (c->cmat.n < (size_t)std::max<long long>(c->cTot, 1) * ncol)
    {
        c->cmat.release();
        if (c->cmat.alloc((size_t)std::max<long long>(c->cTot, 1) * ncol))
            return 1;
    }
    auto ensure = [&](DevBuf<double>& b, size_t count) {
        if (b.n >= count)
            return 0;
        b.release();
        return b.alloc(std::max<size_t>(count, 1));
    };
    if (ensure(c->nrDC, (size_t)dcTot * ncol) || ensure(c->nrPrev, (size_t)prevTot * ncol)
        || ensure(c->nrStages, stages.size()) || ensure(c->nrBgNe, ncol * K) || ensure(c->neDev, ncol * K)
        || ensure(c->nrScratch, ncol * (size_t)(kEnd - kStart) * 2 * Neqn * Neqn))
        return 1;
    c->nrAtoms.release();
    if (c->nrAtoms.upload(atoms))
        return 1;
    const auto H2D = cudaMemcpyHostToDevice;
    for (size_t a = 0; a < atoms.size(); ++a)
    {
        const LwB200Atom& at = c->atoms[atoms[a].atom];
        const size_t n2k = (size_t)at.Nlevel * at.Nlevel * K, nk = (size_t)at.Nlevel * K;
        if (copy2d(c->cmat.p + atoms[a].cOff, (size_t)c->cTot * D, at.C, n2k * D, n2k * D, ncol, H2D, s))
            return 1;
        if (atoms[a].dcOff >= 0
            && copy2d(c->nrDC.p + atoms[a].dcOff, (size_t)dcTot * D, u->dC[a], n2k * D, n2k * D, ncol, H2D, s))
            return 1;
        if (u->timeDependent
            && copy2d(c->nrPrev.p + atoms[a].prevOff, (size_t)prevTot * D, u->nPrev[a], nk * D, nk * D, ncol, H2D, s))
            return 1;
    }
    CU(cudaMemcpyAsync(c->nrStages.p, stages.data(), stages.size() * D, H2D, s));
    CU(cudaMemcpyAsync(c->nrBgNe.p, u->backgroundNe, ncol * K * D, H2D, s));
    CU(cudaMemcpyAsync(c->neDev.p, p.ne, ncol * K * D, H2D, s));
    CU(cudaMemsetAsync(c->dSingular.p, 0, sizeof(int), s));
    c->lastLaunches = 0;
    const size_t total = ncol * (size_t)(kEnd - kStart);
    if (Neqn <= 16)
        nr_update_kernel<16><<<grid_for(total, 64), 64, 0, s>>>(
            c->P, c->nrAtoms.p, (int)atoms.size(), Neqn, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, c->cmat.p,
            c->cTot, c->nrDC.p, dcTot, c->nrPrev.p, prevTot, c->nrStages.p, c->nrBgNe.p, c->neDev.p,
            u->timeDependent ? 1 : 0, u->dt, u->crswVal, c->nrScratch.p, c->dSingular.p);
    else if (Neqn <= 32)
        nr_update_kernel<32><<<grid_for(total, 64), 64, 0, s>>>(
            c->P, c->nrAtoms.p, (int)atoms.size(), Neqn, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, c->cmat.p,
            c->cTot, c->nrDC.p, dcTot, c->nrPrev.p, prevTot, c->nrStages.p, c->nrBgNe.p, c->neDev.p,
            u->timeDependent ? 1 : 0, u->dt, u->crswVal, c->nrScratch.p, c->dSingular.p);
    else if (Neqn <= 64)
        nr_update_kernel<64><<<grid_for(total, 64), 64, 0, s>>>(
            c->P, c->nrAtoms.p, (int)atoms.size(), Neqn, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, c->cmat.p,
            c->cTot, c->nrDC.p, dcTot, c->nrPrev.p, prevTot, c->nrStages.p, c->nrBgNe.p, c->neDev.p,
            u->timeDependent ? 1 : 0, u->dt, u->crswVal, c->nrScratch.p, c->dSingular.p);
    else
        nr_update_kernel<160><<<grid_for(total, 64), 64, 0, s>>>(
            c->P, c->nrAtoms.p, (int)atoms.size(), Neqn, kStart, kEnd, c->gamma.p, c->n.p, c->nTotal.p, c->cmat.p,
            c->cTot, c->nrDC.p, dcTot, c->nrPrev.p, prevTot, c->nrStages.p, c->nrBgNe.p, c->neDev.p,
            u->timeDependent ? 1 : 0, u->dt, u->crswVal, c->nrScratch.p, c->dSingular.p);
    CU(cudaGetLastError());
    c->lastLaunches += 1;
    int ns = 0;
    CU(cudaMemcpyAsync(&ns, c->dSingular.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(p.ne, c->neDev.p, ncol * K * D, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (nSingular)
        *nSingular = ns;
    if (ns > 0)
        return fail("Singular Matrix");
    return 0;
}


int lwb200_kernel_time(LwB200Context* c, double* ms)
{
    CU(cudaSetDevice(c->device));
    if (!c->kernelTimed)
        return fail("lwb200_kernel_time: no formal-solution kernel has been launched yet");
    CU(cudaEventSynchronize(c->evK1));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->evK0, c->evK1));
    if (ms)
        *ms = t;
    return 0;
}

int lwb200_device_buffer(LwB200Context* c, int32_t which, void** ptr, size_t* nbytes)
{
    DevBuf<double>* b = nullptr;
    switch (which)
    {
    case LWB200_BUF_ACCUM: b = &c->accum; break;
    case LWB200_BUF_J: b = &c->J; break;
    case LWB200_BUF_I: b = &c->I; break;
    case LWB200_BUF_POPS: b = &c->n; break;
    case LWB200_BUF_GAMMA: b = &c->gamma; break;
    case LWB200_BUF_DJ: b = &c->dJ; break;
    case LWB200_BUF_DJMAX: b = &c->djOut; break;
    default: return fail("lwb200_device_buffer: unknown buffer");
    }
    if (ptr)
        *ptr = b->p;
    if (nbytes)
        *nbytes = b->n * sizeof(double);
    return 0;
}

int lwb200_work_stats(LwB200Context* c, double* points, double* algBytes, int64_t* lastLaunches)
{
    const LwB200Problem& p = c->prob;
    const double K = p.Nspace, M = p.Nrays, ncol = p.Ncol;
    const double Lr = c->laHi - c->laLo;
    if (points)
        *points = ncol * K * Lr * M * 2.0;
    if (algBytes)
    {
        // every array once
        double Pr = 0.0, nlines = 0.0, b = 0.0;
        for (const DevTrans& d : c->devTrans)
        {
            if (d.type != 0)
                continue;
            const int lo = std::max(d.Nblue, c->laLo), hi = std::min(d.Nred, c->laHi);
            if (hi > lo)
            {
                Pr += hi - lo;
                nlines += 1;
            }
        }
        b = Pr * M * 2 * K + 5.0 * Lr * K + Lr * M + nlines * K + 2 * K;
        for (int a = 0; a < p.Natom; ++a)
        {
            const LwB200Atom& at = c->atoms[a];
            b += 2.0 * at.Nlevel * K + 2.0 * at.Ntrans * K;
            if (!at.detailedStatic)
                b += 2.0 * at.Nlevel * at.Nlevel * K;
        }
        *algBytes = 8.0 * b * ncol;
    }
    if (lastLaunches)
        *lastLaunches = c->lastLaunches;
    return 0;
}
} // extern "C"

#include "lwb200_hprd_host.inl"
