// ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference C++ (compiled from /root/reference/Source where
// it lies, see oracle/Makefile) on the flat LwB200Problem that the CUDA library
// consumes, so that the reference, the C restatement (oracle/lw_oracle.c) and
// the GPU path all see bit-identical inputs.  Built into oracle/_ref/
// (git-ignored); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
//
// It fills the reference's POD structs with non-owning views the way the Cython
// middle layer does (reference: Source/LwMiddleLayer.pyx LwAtmosphere.__init__
// :620-715, LwSpectrum :2724-2732, LwBackground :1571-1597, LwTransition
// :1772-1825, LwAtom :2346-2424, LwContext :2946-2973) and then calls the
// reference's own entry points formal_sol_gamma_matrices / formal_sol / stat_eq
// (Source/Lightweaver.hpp:21-32).  Iteration schemes: the built-in scalar one
// (the parity oracle), the reference's SIMD plugins (timing baseline), or any
// plugin path -- which is how tests load OUR plugin through the reference's own
// FsIterationFnsManager::load_fns_from_path (Source/FormalInterface.cpp:62-81).

#include "Lightweaver.hpp"
#include "../include/lwb200.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <deque>
#include <dlfcn.h>
#include <memory>
#include <string>
#include <vector>

namespace
{
thread_local std::string g_err;

struct LwRef
{
    const LwB200Problem* prob = nullptr;
    int col = 0;
    Atmosphere atmos{};
    Spectrum spect{};
    Background background{};
    DepthData depth{};
    std::deque<Atom> atoms;
    std::deque<Transition> trans;
    std::deque<std::vector<int8_t>> activeMasks;
    std::vector<f64> stagesDummy;
    std::vector<f64> vzDummy;
    std::vector<std::vector<f64>> gammaPrefill;
    FormalSolverManager fsManager;
    FsIterationFnsManager iterManager;
    std::unique_ptr<Context> ctx;
    double *zUp = nullptr, *zDown = nullptr; // ZPlaneDecomposition outputs [Nspect][Nrays] (lwref_set_zplane)

    // the ExtraParams a Python caller would pass (dict2ExtraParams, LwMiddleLayer.pyx:358-467)
    ExtraParams params() const
    {
        ExtraParams ep{};
        if (zUp || zDown)
        {
            const i64 L = prob->Nspect, M = prob->Nrays;
            ep.insert("ZPlaneDecomposition", true);
            if (zUp)
                ep.insert("ZPlaneUp", F64View2D(zUp, L, M));
            if (zDown)
                ep.insert("ZPlaneDown", F64View2D(zDown, L, M));
        }
        return ep;
    }
};

std::string harness_dir()
{
    Dl_info info;
    if (dladdr((void*)&harness_dir, &info) && info.dli_fname)
    {
        std::string p(info.dli_fname);
        auto pos = p.find_last_of('/');
        if (pos != std::string::npos)
            return p.substr(0, pos);
    }
    return ".";
}
}

extern "C"
{
struct LwRefHandle;

const char* lwref_last_error() { return g_err.c_str(); }

// scheme: "scalar" | "SSE2" | "AVX2FMA" | "AVX512" | path of an iteration-scheme plugin.
int lwref_create(const LwB200Problem* p, int col, const char* scheme, int Nthreads, LwRefHandle** out)
{
    try
    {
        if (!p || p->abiVersion != LWB200_ABI_VERSION)
            throw std::runtime_error("lwref_create: bad problem / ABI version");
        if (col < 0 || col >= p->Ncol)
            throw std::runtime_error("lwref_create: column out of range");

        auto h = std::make_unique<LwRef>();
        h->prob = p;
        h->col = col;
        const i64 K = p->Nspace;
        const i64 M = p->Nrays;
        const i64 L = p->Nspect;
        const i64 c = col;

        auto& a = h->atmos;
        a.Nspace = K; a.Nrays = M; a.Ndim = 1;
        a.Nx = 0; a.Ny = 0; a.Nz = K; a.Noutgoing = 1;
        a.height = F64View(const_cast<f64*>(p->height) + c * K, K);
        a.z = a.height;
        a.temperature = F64View(const_cast<f64*>(p->temperature) + c * K, K);
        h->vzDummy.assign(K, 0.0);
        a.vz = F64View(h->vzDummy.data(), K);
        if (p->vlosMu)
            a.vlosMu = F64View2D(const_cast<f64*>(p->vlosMu) + c * M * K, M, K);
        a.muz = F64View(const_cast<f64*>(p->muz), M);
        if (p->ne)
            a.ne = F64View(p->ne + c * K, K);
        a.wmu = F64View(const_cast<f64*>(p->wmu), M);

        auto set_bc = [&](AtmosphericBoundaryCondition& bc, int type, int Nmu,
                          const f64* data, const int32_t* idx)
        {
            BcIdxs idxs;
            if (idx)
                idxs = BcIdxs(const_cast<i32*>(idx), M, 2);
            bc = AtmosphericBoundaryCondition((RadiationBc)type, L, Nmu, 1, idxs);
            if (type == CALLABLE)
            {
                if (!data || !idx)
                    throw std::runtime_error("CALLABLE boundary without data/idxs");
                for (i64 la = 0; la < L; ++la)
                    for (int mu = 0; mu < Nmu; ++mu)
                        bc.bcData(la, mu, 0) = data[(c * L + la) * Nmu + mu];
            }
        };
        set_bc(a.zLowerBc, p->lowerBc, p->NlowerBcMu, p->lowerBcData, p->lowerBcIdx);
        set_bc(a.zUpperBc, p->upperBc, p->NupperBcMu, p->upperBcData, p->upperBcIdx);

        auto& s = h->spect;
        s.wavelength = F64View(const_cast<f64*>(p->wavelength), L);
        s.I = F64View3D(p->I + c * L * M, L, M, 1);
        s.J = F64View2D(p->J + c * L * K, L, K);
        if (p->Quv)
        {
            s.Quv = F64View4D(p->Quv + c * 3 * L * M, 3, L, M, 1);
            a.B = F64View(h->vzDummy.data(), K); // formal_sol_full_stokes_impl only checks that B exists
        }

        auto& bg = h->background;
        bg.chi = F64View2D(const_cast<f64*>(p->chiBg) + c * L * K, L, K);
        bg.eta = F64View2D(const_cast<f64*>(p->etaBg) + c * L * K, L, K);
        bg.sca = F64View2D(const_cast<f64*>(p->scaBg) + c * L * K, L, K);

        h->ctx = std::make_unique<Context>();
        auto& ctx = *h->ctx;
        ctx.atmos = &h->atmos;
        ctx.spect = &h->spect;
        ctx.background = &h->background;
        ctx.depthData = nullptr;
        if (p->depthChi && p->depthEta && p->depthI)
        {
            const i64 n = L * M * 2 * K;
            h->depth.fill = false;
            h->depth.chi = F64View4D(p->depthChi + c * n, L, M, 2, K);
            h->depth.eta = F64View4D(p->depthEta + c * n, L, M, 2, K);
            h->depth.I = F64View4D(p->depthI + c * n, L, M, 2, K);
            ctx.depthData = &h->depth;
        }
        ctx.methodScratch = nullptr;

        i64 maxLevel = 1;
        for (int ia = 0; ia < p->Natom; ++ia)
            maxLevel = std::max<i64>(maxLevel, p->atoms[ia].Nlevel);
        h->stagesDummy.assign(maxLevel, 0.0);

        for (int ia = 0; ia < p->Natom; ++ia)
        {
            const LwB200Atom& pa = p->atoms[ia];
            h->atoms.emplace_back();
            Atom& atom = h->atoms.back();
            const i64 N = pa.Nlevel;
            atom.Nlevel = N;
            atom.Ntrans = pa.Ntrans;
            atom.atmos = &h->atmos;
            atom.n = F64View2D(pa.n + c * N * K, N, K);
            atom.nStar = F64View2D(const_cast<f64*>(pa.nStar) + c * N * K, N, K);
            atom.nTotal = F64View(const_cast<f64*>(pa.nTotal) + c * K, K);
            if (pa.vBroad)
                atom.vBroad = F64View(const_cast<f64*>(pa.vBroad) + c * K, K);
            atom.stages = pa.stages ? F64View(const_cast<f64*>(pa.stages), N) : F64View(h->stagesDummy.data(), N);
            if (!pa.detailedStatic)
            {
                if (!pa.Gamma)
                    throw std::runtime_error("active atom without Gamma");
                atom.Gamma = F64View3D(pa.Gamma + c * N * N * K, N, N, K);
            }
            if (pa.C)
                atom.C = F64View3D(const_cast<f64*>(pa.C) + c * N * N * K, N, N, K);
            atom.methodScratch = nullptr;

            for (int kr = 0; kr < pa.Ntrans; ++kr)
            {
                const LwB200Transition& pt = pa.trans[kr];
                h->trans.emplace_back();
                Transition& t = h->trans.back();
                const i64 Nl = pt.Nred - pt.Nblue;
                t.Nblue = pt.Nblue;
                t.Nred = pt.Nred;
                t.type = pt.type == LWB200_LINE ? LINE : CONTINUUM;
                t.i = pt.i;
                t.j = pt.j;
                t.Aji = pt.Aji; t.Bji = pt.Bji; t.Bij = pt.Bij;
                t.lambda0 = pt.lambda0;
                t.dopplerWidth = pt.dopplerWidth;
                t.polarised = false;
                t.wavelength = F64View(const_cast<f64*>(pt.wavelength), Nl);
                if (t.type == LINE)
                {
                    t.phi = F64View4D(pt.phi + c * Nl * M * 2 * K, Nl, M, 2, K);
                    t.wphi = F64View(pt.wphi + c * K, K);
                    if (pt.aDamp)
                        t.aDamp = F64View(const_cast<f64*>(pt.aDamp) + c * K, K);
                    if (pt.rhoPrd)
                        t.rhoPrd = F64View2D(const_cast<f64*>(pt.rhoPrd) + c * Nl * K, Nl, K);
                    if (pt.Qelast)
                        t.Qelast = F64View(const_cast<f64*>(pt.Qelast) + c * K, K);
                    if (pt.polProfiles)
                    {
                        const i64 per = Nl * M * 2 * K, arr = (i64)p->Ncol * per;
                        f64* base = const_cast<f64*>(pt.polProfiles) + c * per;
                        t.polarised = true;
                        t.phiQ = F64View4D(base + 0 * arr, Nl, M, 2, K);
                        t.phiU = F64View4D(base + 1 * arr, Nl, M, 2, K);
                        t.phiV = F64View4D(base + 2 * arr, Nl, M, 2, K);
                        t.psiQ = F64View4D(base + 3 * arr, Nl, M, 2, K);
                        t.psiU = F64View4D(base + 4 * arr, Nl, M, 2, K);
                        t.psiV = F64View4D(base + 5 * arr, Nl, M, 2, K);
                    }
                }
                else
                {
                    t.alpha = F64View(const_cast<f64*>(pt.alpha), Nl);
                }
                h->activeMasks.emplace_back(L, 0);
                auto& mask = h->activeMasks.back();
                for (i64 la = pt.Nblue; la < pt.Nred; ++la)
                    mask[la] = 1;
                t.active = BoolView((bool*)mask.data(), L);
                t.Rij = F64View(pt.Rij + c * K, K);
                t.Rji = F64View(pt.Rji + c * K, K);
                t.methodScratch = nullptr;
                atom.trans.push_back(&t);
            }
            atom.init_scratch(K, pa.detailedStatic != 0, true, true);
            if (pa.detailedStatic)
                ctx.detailedAtoms.push_back(&atom);
            else
                ctx.activeAtoms.push_back(&atom);
        }

        ctx.Nthreads = std::max(Nthreads, 1);
        if (p->formalSolver < 0 || p->formalSolver > 2)
            throw std::runtime_error("formalSolver must be 0 (linear), 1 (besser) or 2 (bezier3)");
        ctx.formalSolver = h->fsManager.formalSolvers[p->formalSolver];

        std::string sch(scheme ? scheme : "scalar");
        if (sch == "scalar")
        {
            ctx.iterFns = h->iterManager.fns[0];
        }
        else
        {
            std::string path = sch;
            if (sch == "SSE2" || sch == "AVX2FMA" || sch == "AVX512")
                path = harness_dir() + "/SimdImpl_" + sch + ".so";
            if (!h->iterManager.load_fns_from_path(path.c_str()))
            {
                const char* why = dlerror();
                throw std::runtime_error("could not load iteration scheme from " + path
                                         + (why ? std::string(": ") + why : std::string()));
            }
            ctx.iterFns = h->iterManager.fns.back();
        }
        ctx.initialise_threads();

        *out = (LwRefHandle*)h.release();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

void lwref_destroy(LwRefHandle* hh)
{
    auto* h = (LwRef*)hh;
    if (!h)
        return;
    if (h->ctx)
        h->ctx->threading.clear(h->ctx.get());
    delete h;
}

const char* lwref_scheme_name(LwRefHandle* hh)
{
    return ((LwRef*)hh)->ctx->iterFns.name;
}

int lwref_set_depth_fill(LwRefHandle* hh, int fill)
{
    auto* h = (LwRef*)hh;
    if (!h->ctx->depthData)
    {
        g_err = "no depthData arrays in the problem";
        return 1;
    }
    h->depth.fill = fill != 0;
    return 0;
}

// One Gamma iteration.  The caller pre-fills Gamma with crsw*C, as
// LwContext.formal_sol_gamma_matrices does (LwMiddleLayer.pyx:3198-3203).
int lwref_fs_iter(LwRefHandle* hh, int lambdaIterate, double* dJMax, int64_t* dJMaxIdx)
{
    auto* h = (LwRef*)hh;
    try
    {
        IterationResult r = formal_sol_gamma_matrices(*h->ctx, lambdaIterate != 0, h->params());
        if (dJMax) *dJMax = r.dJMax;
        if (dJMaxIdx) *dJMaxIdx = r.dJMaxIdx;
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// extraParams {"ZPlaneDecomposition": True, "ZPlaneUp": up, "ZPlaneDown": down} for the following
// lwref_fs_iter / lwref_formal_sol calls (SimdFullIterationTemplates.hpp:254-281); NULL, NULL: none.
int lwref_set_zplane(LwRefHandle* hh, double* up, double* down)
{
    auto* h = (LwRef*)hh;
    h->zUp = up;
    h->zDown = down;
    return 0;
}

int lwref_formal_sol(LwRefHandle* hh, int upOnly)
{
    auto* h = (LwRef*)hh;
    try
    {
        formal_sol(*h->ctx, upOnly != 0, h->params());
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// formal_sol_full_stokes (LwContext.single_stokes_fs, LwMiddleLayer.pyx): polarised formal solution.
int lwref_full_stokes_j20(LwRefHandle* hh, int updateJ, int upOnly, double* J20, double* dJMax, int64_t* dJMaxIdx);
int lwref_full_stokes(LwRefHandle* hh, int updateJ, int upOnly, double* dJMax, int64_t* dJMaxIdx)
{
    return lwref_full_stokes_j20(hh, updateJ, upOnly, nullptr, dJMax, dJMaxIdx);
}

// J20: the "J20" extra parameter of this column, [Nspect][Nspace], or NULL
int lwref_full_stokes_j20(LwRefHandle* hh, int updateJ, int upOnly, double* J20, double* dJMax, int64_t* dJMaxIdx)
{
    auto* h = (LwRef*)hh;
    try
    {
        ExtraParams params{};
        if (J20)
            params.insert("J20", F64View2D(J20, h->prob->Nspect, h->prob->Nspace));
        IterationResult r = formal_sol_full_stokes(*h->ctx, updateJ != 0, upOnly != 0, params);
        if (dJMax) *dJMax = r.dJMax;
        if (dJMaxIdx) *dJMaxIdx = r.dJMaxIdx;
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// nr_post_update (LwContext._nr_post_update_impl, LwMiddleLayer.pyx:3533-3564) on this column.
int lwref_nr_post_update(LwRefHandle* hh, const LwB200NrUpdate* u)
{
    auto* h = (LwRef*)hh;
    try
    {
        const i64 K = h->prob->Nspace, c = h->col;
        std::vector<Atom*> atoms;
        std::vector<F64View3D> dC;
        NrTimeDependentData td{};
        td.dt = u->dt;
        for (int a = 0; a < u->Natom; ++a)
        {
            Atom* atom = &h->atoms.at(u->atomIdx[a]);
            const i64 N = atom->Nlevel;
            atoms.push_back(atom);
            if (u->dC)
                dC.emplace_back(const_cast<f64*>(u->dC[a]) + c * N * N * K, N, N, K);
            if (u->timeDependent)
                td.nPrev.emplace_back(const_cast<f64*>(u->nPrev[a]) + c * N * K, N, K);
        }
        nr_post_update(*h->ctx, &atoms, dC, F64View(const_cast<f64*>(u->backgroundNe) + c * K, K), td, u->crswVal,
                       ExtraParams{}, -1, -1);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// time_dependent_update of active atom number `activeIdx` (LwContext.time_dep_update,
// LwMiddleLayer.pyx:3420-3424).  nOld: [Nlevel][Nspace] of this column.
int lwref_time_dep_update(LwRefHandle* hh, int activeIdx, double* nOld, double dt)
{
    auto* h = (LwRef*)hh;
    try
    {
        Atom* a = h->ctx->activeAtoms.at(activeIdx);
        time_dependent_update(*h->ctx, a, F64View2D(nOld, a->Nlevel, h->atmos.Nspace), dt, ExtraParams{}, -1, -1);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// redistribute_prd_lines (LwContext.prd_redistribute, LwMiddleLayer.pyx:3647-3684).
int lwref_redistribute_prd(LwRefHandle* hh, int maxIter, double tol, int includeDetailed, int* nIter,
                           double* dRho, int* dRhoIdx, double* dJPrdMax, int64_t* dJPrdMaxIdx)
{
    auto* h = (LwRef*)hh;
    try
    {
        ExtraParams params{};
        params.insert("include_detailed_atoms", includeDetailed != 0);
        IterationResult r = redistribute_prd_lines(*h->ctx, maxIter, tol, params);
        if (nIter) *nIter = r.NprdSubIter;
        for (size_t q = 0; q < r.dRho.size(); ++q)
        {
            if (dRho) dRho[q] = r.dRho[q];
            if (dRhoIdx) dRhoIdx[q] = r.dRhoMaxIdx[q];
        }
        for (size_t q = 0; q < r.dJPrdMax.size(); ++q)
        {
            if (dJPrdMax) dJPrdMax[q] = r.dJPrdMax[q];
            if (dJPrdMaxIdx) dJPrdMaxIdx[q] = r.dJPrdMaxIdx[q];
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// configure_hprd_coeffs (LwContext.configure_hprd_coeffs -> update_threads, LwMiddleLayer.pyx): hybrid PRD of
// this handle's column, done by the reference itself.  The tables it builds are exported, flattened exactly
// like LwB200HybridPrd lays out ONE column, so that the oracle's restatement can be compared with them:
//   prdLaOfLa [Nspect], hPrdLaOfLa [Nspect], JCoeffOff [NhPrd*M*2*K + 1], JCoeffIdx / JCoeffFrac [nnz],
//   per PRD line (the order of configure_hprd_coeffs) rhoFrac / rhoI0 [Nl*M*2*K] back to back.
// Every out pointer may be NULL; counts[0..3] = NprdLa, NhPrd, nnz, Nlines.
int lwref_configure_hprd(LwRefHandle* hh, int includeDetailed)
{
    auto* h = (LwRef*)hh;
    try
    {
        configure_hprd_coeffs(*h->ctx, includeDetailed != 0);
        h->ctx->update_threads();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

int lwref_hprd_export(LwRefHandle* hh, int includeDetailed, int64_t* counts, int32_t* prdLaOfLa, int32_t* hPrdLaOfLa,
                      int64_t* JCoeffOff, int32_t* JCoeffIdx, double* JCoeffFrac, double* rhoFrac, int32_t* rhoI0)
{
    auto* h = (LwRef*)hh;
    auto& spect = h->spect;
    const i64 K = h->prob->Nspace, M = h->prob->Nrays, L = h->prob->Nspect;
    if (!spect.JRest)
    {
        g_err = "hybrid PRD has not been configured";
        return 1;
    }
    const i64 NprdLa = spect.JRest.shape(0), NhPrd = (i64)spect.hPrdIdxs.size();
    for (i64 la = 0; la < L; ++la)
    {
        if (prdLaOfLa) prdLaOfLa[la] = spect.prdActive(la) ? spect.la_to_prdLa(la) : -1;
        if (hPrdLaOfLa) hPrdLaOfLa[la] = spect.hPrdActive(la) ? spect.la_to_hPrdLa(la) : -1;
    }
    i64 nnz = 0;
    for (i64 q = 0; q < NhPrd; ++q)
        for (i64 mu = 0; mu < M; ++mu)
            for (int toObs = 0; toObs < 2; ++toObs)
                for (i64 k = 0; k < K; ++k)
                {
                    const auto& v = spect.JCoeffs(q, mu, toObs, k);
                    if (JCoeffOff) JCoeffOff[((q * M + mu) * 2 + toObs) * K + k] = nnz;
                    for (const auto& c : v)
                    {
                        if (JCoeffIdx) JCoeffIdx[nnz] = c.idx;
                        if (JCoeffFrac) JCoeffFrac[nnz] = c.frac;
                        ++nnz;
                    }
                }
    if (JCoeffOff) JCoeffOff[NhPrd * M * 2 * K] = nnz;
    i64 nLines = 0, o = 0;
    auto lines_of = [&](std::vector<Atom*>& atoms) {
        for (auto* a : atoms)
            for (auto* t : a->trans)
                if (t->rhoPrd)
                {
                    ++nLines;
                    const i64 Nl = t->wavelength.shape(0);
                    for (i64 lt = 0; lt < Nl; ++lt)
                        for (i64 mu = 0; mu < M; ++mu)
                            for (int toObs = 0; toObs < 2; ++toObs)
                                for (i64 k = 0; k < K; ++k, ++o)
                                {
                                    const auto& c = t->hPrdCoeffs(lt, mu, toObs, k);
                                    if (rhoFrac) rhoFrac[o] = c.frac;
                                    if (rhoI0) rhoI0[o] = c.i0;
                                    if (c.i1 != c.i0 + 1)
                                        throw std::runtime_error("hPrdCoeffs: i1 != i0 + 1");
                                }
                }
    };
    try
    {
        lines_of(h->ctx->activeAtoms);
        if (includeDetailed)
            lines_of(h->ctx->detailedAtoms);
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
    if (counts)
    {
        counts[0] = NprdLa; counts[1] = NhPrd; counts[2] = nnz; counts[3] = nLines;
    }
    return 0;
}

// the reference's rest-frame mean intensity JRest [NprdLa][Nspace] of this column
int lwref_get_jrest(LwRefHandle* hh, double* out)
{
    auto* h = (LwRef*)hh;
    if (!h->spect.JRest)
    {
        g_err = "hybrid PRD has not been configured";
        return 1;
    }
    std::memcpy(out, h->spect.JRest.dataStore.data(), sizeof(double) * h->spect.JRest.shape(0) * h->spect.JRest.shape(1));
    return 0;
}

// stat_eq over every active atom (LwContext.stat_equil, LwMiddleLayer.pyx:3509-3514).
int lwref_stat_eq(LwRefHandle* hh)
{
    auto* h = (LwRef*)hh;
    try
    {
        for (Atom* a : h->ctx->activeAtoms)
            stat_eq(*h->ctx, a, ExtraParams{}, -1, -1);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// Transition::compute_phi + compute_wphi for every line from aDamp / vBroad /
// vlosMu (LwContext.compute_profiles; Source/FormalScalar.cpp:28-134).
int lwref_compute_profiles(LwRefHandle* hh)
{
    auto* h = (LwRef*)hh;
    try
    {
        auto do_atoms = [&](std::vector<Atom*>& atoms)
        {
            for (Atom* a : atoms)
                for (Transition* t : a->trans)
                {
                    if (t->type != LINE)
                        continue;
                    if (!t->aDamp || !a->vBroad || !h->atmos.vlosMu)
                        throw std::runtime_error("compute_profiles needs aDamp, vBroad and vlosMu");
                    t->compute_phi(h->atmos, t->aDamp, a->vBroad);
                    t->compute_wphi(h->atmos);
                }
        };
        do_atoms(h->ctx->activeAtoms);
        do_atoms(h->ctx->detailedAtoms);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// Transition::compute_polarised_profiles (FormalStokes.cpp:9-117) of one polarised line of this column, by the
// reference itself: B [Nspace], cosGamma / cos2chi / sin2chi [Nrays][Nspace], the Zeeman pattern.  Writes the
// line's phi, wphi and its six polarised profiles (the problem's host arrays).
int lwref_compute_polarised_profiles(LwRefHandle* hh, int atomIdx, int transIdx, const double* B, const double* cosGamma,
                                     const double* cos2chi, const double* sin2chi, int nComp, const int32_t* alpha,
                                     const double* shift, const double* strength)
{
    auto* h = (LwRef*)hh;
    try
    {
        const i64 K = h->prob->Nspace, M = h->prob->Nrays;
        if (atomIdx < 0 || atomIdx >= (int)h->atoms.size())
            throw std::runtime_error("atom index out of range");
        Atom& a = h->atoms[atomIdx];
        if (transIdx < 0 || transIdx >= (int)a.trans.size())
            throw std::runtime_error("transition index out of range");
        Transition* t = a.trans[transIdx];
        if (!t->polarised || !t->aDamp || !a.vBroad || !h->atmos.vlosMu)
            throw std::runtime_error("compute_polarised_profiles needs a polarised line with aDamp, vBroad and vlosMu");
        Atmosphere atm = h->atmos;
        atm.B = F64View(const_cast<f64*>(B), K);
        atm.cosGamma = F64View2D(const_cast<f64*>(cosGamma), M, K);
        atm.cos2chi = F64View2D(const_cast<f64*>(cos2chi), M, K);
        atm.sin2chi = F64View2D(const_cast<f64*>(sin2chi), M, K);
        ZeemanComponents z;
        z.alpha = I32View(const_cast<i32*>(alpha), nComp);
        z.shift = F64View(const_cast<f64*>(shift), nComp);
        z.strength = F64View(const_cast<f64*>(strength), nComp);
        t->compute_polarised_profiles(atm, t->aDamp, a.vBroad, z);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// Timing protocol of lightweaver/benchmark.py:84-89 / BASELINE.md section 3:
// nWarm + nTimed calls, Gamma re-filled from its value at entry before each
// call, steady_clock around the C++ call only.  seconds[] receives nTimed
// samples.  withStatEq != 0 adds stat_eq to the timed region (populations are
// restored afterwards so every call sees the same inputs).
int lwref_time_fs_iter(LwRefHandle* hh, int nWarm, int nTimed, int withStatEq, double* seconds)
{
    auto* h = (LwRef*)hh;
    try
    {
        auto& atoms = h->ctx->activeAtoms;
        std::vector<std::vector<f64>> prefill, pops;
        for (Atom* a : atoms)
        {
            const i64 n3 = a->Gamma.shape(0) * a->Gamma.shape(1) * a->Gamma.shape(2);
            prefill.emplace_back(a->Gamma.data, a->Gamma.data + n3);
            const i64 n2 = a->n.shape(0) * a->n.shape(1);
            pops.emplace_back(a->n.data, a->n.data + n2);
        }
        for (int it = 0; it < nWarm + nTimed; ++it)
        {
            for (size_t ia = 0; ia < atoms.size(); ++ia)
            {
                std::memcpy(atoms[ia]->Gamma.data, prefill[ia].data(), prefill[ia].size() * sizeof(f64));
                if (withStatEq)
                    std::memcpy(atoms[ia]->n.data, pops[ia].data(), pops[ia].size() * sizeof(f64));
            }
            auto t0 = std::chrono::steady_clock::now();
            formal_sol_gamma_matrices(*h->ctx, false, ExtraParams{});
            if (withStatEq)
                for (Atom* a : atoms)
                    stat_eq(*h->ctx, a, ExtraParams{}, -1, -1);
            auto t1 = std::chrono::steady_clock::now();
            if (it >= nWarm)
                seconds[it - nWarm] = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// The reference's formal-solver registry (Source/FormalInterface.cpp:9-37), shared by the calls below.
static FormalSolverManager& solver_manager()
{
    static FormalSolverManager man;
    return man;
}

// FormalSolverManager::load_fs_from_path (Source/FormalInterface.cpp:9-28): dlopen a plugin, look up
// "fs_provider", append its solver.  *index receives its position in the registry, name (if not NULL)
// a copy of FormalSolver::name.
int lwref_load_formal_solver(const char* path, int* index, char* name, int nameLen)
{
    try
    {
        FormalSolverManager& man = solver_manager();
        if (!man.load_fs_from_path(path))
            throw std::runtime_error(std::string("load_fs_from_path failed: ") + path);
        if (index)
            *index = (int)man.formalSolvers.size() - 1;
        if (name && nameLen > 0)
        {
            std::strncpy(name, man.formalSolvers.back().name, nameLen - 1);
            name[nameLen - 1] = 0;
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// One ray through the reference's own 1D solvers (FormalSolverManager order:
// 0 linear, 1 besser, 2 bezier3; >= 3: solvers loaded with lwref_load_formal_solver).
int lwref_solve_ray(int solver, int Nspace, const double* height, const double* temperature,
                    const double* chi, const double* S, double muz, int toObs, double wavelength,
                    int lowerBc, int upperBc, double* I, double* Psi)
{
    try
    {
        FormalSolverManager& man = solver_manager();
        if (solver < 0 || solver >= (int)man.formalSolvers.size())
            throw std::runtime_error("bad solver index");
        Atmosphere atmos{};
        atmos.Nspace = Nspace; atmos.Nrays = 1; atmos.Ndim = 1; atmos.Nz = Nspace;
        atmos.height = F64View(const_cast<f64*>(height), Nspace);
        atmos.temperature = F64View(const_cast<f64*>(temperature), Nspace);
        atmos.muz = F64View(&muz, 1);
        atmos.zLowerBc.type = (RadiationBc)lowerBc;
        atmos.zUpperBc.type = (RadiationBc)upperBc;
        LwInternal::FormalData fd;
        fd.atmos = &atmos;
        fd.chi = F64View(const_cast<f64*>(chi), Nspace);
        fd.S = F64View(const_cast<f64*>(S), Nspace);
        fd.I = F64View(I, Nspace);
        if (Psi)
            fd.Psi = F64View(Psi, Nspace);
        F64View1D wave(&wavelength, 1);
        man.formalSolvers[solver].solver(&fd, 0, 0, toObs != 0, wave);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// The reference's own Ng object (Source/Ng.hpp) driven over a prescribed sequence of solutions:
// constructor on sols[0], then accelerate() + max_change() on sols[1..nIter].
int lwref_ng_run(int Norder, int Nperiod, int Ndelay, int len, int nIter, const double* sols, double* out,
                 int* accelerated, double* dMax, int64_t* dMaxIdx)
{
    try
    {
        std::vector<double> first(sols, sols + len);
        Ng ng(Norder, Nperiod, Ndelay, F64View(first.data(), len));
        for (int it = 0; it < nIter; ++it)
        {
            double* sol = out + (size_t)it * len;
            std::memcpy(sol, sols + (size_t)(it + 1) * len, sizeof(double) * len);
            accelerated[it] = ng.accelerate(F64View(sol, len)) ? 1 : 0;
            const NgChange ch = ng.max_change();
            dMax[it] = ch.dMax;
            dMaxIdx[it] = ch.dMaxIdx;
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}

// solve_lin_eq (Source/LuSolve.cpp:103-133) on a caller-owned N x N system.
int lwref_solve_lin_eq(int N, double* A, double* b, int improve)
{
    try
    {
        solve_lin_eq(F64View2D(A, N, N), F64View(b, N), improve != 0);
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}
}
