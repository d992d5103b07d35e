// lwb200_plugin.cpp -- the Lightweaver plugin shim (layer 1 of the boundary).
//
// Exports the two runtime-loadable provider symbols Lightweaver looks up with
// dlopen/dlsym (Source/LwFormalInterfacePosix.hpp:12-37):
//
//   extern "C" FsIterationFns fs_iteration_fns_provider();   (Source/FormalInterface.cpp:62-81)
//   extern "C" FormalSolver   fs_provider();                 (Source/FormalInterface.cpp:9-28)
//
// and implements `fs_iter`, `simple_fs`, `stat_eq` and the global-scratch hooks
// of FsIterationFns (Source/LwFormalInterface.hpp:110-134) by marshalling the
// reference's Context into the flat LwB200Problem of include/lwb200.h and
// calling the C-ABI of liblwb200.so.  Everything below this file is plain C.
//
// Like the reference's own SIMD plugins (setup.py:255-258) this translation unit
// is COMPILED AGAINST THE REFERENCE HEADERS and linked with the reference core
// (for the slots this back end does not replace: full Stokes, PRD
// redistribution, time-dependent and charge-conservation updates keep the
// core's `*_impl` functions, exactly as SimdImpl_AVX2FMA.cpp:652-656 does).  It
// can therefore only be (re)built where the reference sources exist; no
// reference source is copied into this repository.
//
// Use from Python:
//   lw.LwCompiled.FsIterationSchemes.load_fns_from_path('liblwb200_plugin.so')
//   ctx = lw.Context(..., fsIterScheme='mali_full_precond_B200', Nthreads=1)
#include "Lightweaver.hpp"
#include "lwb200.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
struct Mirror
{
    LwB200Context* dev = nullptr;
    LwB200Problem prob{};
    std::vector<LwB200Atom> atoms;
    std::deque<std::vector<LwB200Transition>> trans;
    std::vector<Atom*> hostAtoms;
    std::vector<double> lowerBc, upperBc;
    std::deque<std::vector<double>> polStage;                    // six polarised profiles per polarised line
    std::vector<std::pair<const Transition*, size_t>> polSrc;   // (line, index into polStage)
    bool uploadedStatic = false;
    bool hasDepth = false;
    uint64_t fpProfiles = 0, fpBackground = 0, fpAtmos = 0, fpJ = 0;
    int solver = -1;
};

std::mutex g_mutex;
std::map<Context*, std::unique_ptr<Mirror>> g_mirrors;
std::map<Atom*, std::pair<Mirror*, int>> g_atoms;

[[noreturn]] void raise(const std::string& what)
{
    const char* e = lwb200_last_error();
    throw std::runtime_error(what + (e && *e ? std::string(": ") + e : std::string()));
}

void check(int rc, const char* what)
{
    if (rc != 0)
        raise(what);
}

// Sampled FNV-1a fingerprint of a host array: Python mutates phi / background /
// atmosphere in place between calls without telling the plugin;
// any update_deps() changes essentially every element, so sampling detects it.
uint64_t fingerprint(uint64_t h, const double* p, size_t n)
{
    if (!p || n == 0)
        return h;
    const size_t stride = n > 8192 ? n / 4096 : 1;
    auto mix = [&](double v)
    {
        uint64_t b;
        std::memcpy(&b, &v, 8);
        h = (h ^ b) * 1099511628211ULL;
    };
    for (size_t i = 0; i < n; i += stride)
        mix(p[i]);
    mix(p[n - 1]);
    return h;
}

int solver_from_name(const char* name)
{
    const std::string s(name ? name : "");
    if (s == "piecewise_linear_1d")
        return LWB200_FS_LINEAR;
    if (s == "piecewise_besser_1d")
        return LWB200_FS_BESSER;
    if (s == "piecewise_bezier3_1d" || s == "piecewise_bezier3_1d_b200")
        return LWB200_FS_BEZIER3;
    throw std::runtime_error("mali_full_precond_B200: formal solver '" + s
                             + "' has no device kernel (1D linear, besser and bezier3 do)");
}

void destroy_mirror(Mirror* m)
{
    for (Atom* a : m->hostAtoms)
        g_atoms.erase(a);
    if (m->dev)
        lwb200_destroy(m->dev);
    m->dev = nullptr;
}

// Fill the flat problem from the reference's Context (field list: SURVEY.md 8b).
void build_mirror(Context& ctx, Mirror& m)
{
    Atmosphere& atmos = *ctx.atmos;
    Spectrum& spect = *ctx.spect;
    Background& bg = *ctx.background;
    if (atmos.Ndim != 1)
        throw std::runtime_error("mali_full_precond_B200 only handles 1D atmospheres");
    if (spect.JRest)
        throw std::runtime_error("mali_full_precond_B200: hybrid PRD (JRest) is not supported yet");
    const int K = atmos.Nspace, M = atmos.Nrays, L = (int)spect.wavelength.shape(0);
    m.solver = solver_from_name(ctx.formalSolver.name);

    LwB200Problem& p = m.prob;
    p = LwB200Problem{};
    p.abiVersion = LWB200_ABI_VERSION;
    p.Ncol = 1;
    p.Nspace = K;
    p.Nrays = M;
    p.Nspect = L;
    p.formalSolver = m.solver;
    p.lowerBc = (int)atmos.zLowerBc.type;
    p.upperBc = (int)atmos.zUpperBc.type;
    p.height = atmos.height.data;
    p.temperature = atmos.temperature.data;
    p.vlosMu = atmos.vlosMu.data;
    p.muz = atmos.muz.data;
    p.wmu = atmos.wmu.data;
    p.wavelength = spect.wavelength.data;
    p.chiBg = bg.chi.data;
    p.etaBg = bg.eta.data;
    p.scaBg = bg.sca.data;
    p.J = spect.J.data;
    p.I = spect.I.data;
    p.Quv = spect.Quv ? spect.Quv.data : nullptr;
    p.ne = atmos.ne ? atmos.ne.data : nullptr;
    if (spect.I.shape(2) != 1)
        throw std::runtime_error("mali_full_precond_B200: spect.I must have one outgoing point per ray");
    auto bind_bc = [&](AtmosphericBoundaryCondition& bc, int& nmu, const double*& data, const int32_t*& idx)
    {
        nmu = 0;
        data = nullptr;
        idx = nullptr;
        if (bc.type == CALLABLE)
        {
            if (bc.bcData.shape(2) != 1)
                throw std::runtime_error("mali_full_precond_B200: boundary data must be 1D");
            nmu = (int)bc.bcData.shape(1);
            data = bc.bcData.data();
            idx = bc.idxs.data;
        }
    };
    bind_bc(atmos.zLowerBc, p.NlowerBcMu, p.lowerBcData, p.lowerBcIdx);
    bind_bc(atmos.zUpperBc, p.NupperBcMu, p.upperBcData, p.upperBcIdx);

    m.hasDepth = false;
    if (ctx.depthData && ctx.depthData->chi && ctx.depthData->eta && ctx.depthData->I)
    {
        p.depthChi = ctx.depthData->chi.data;
        p.depthEta = ctx.depthData->eta.data;
        p.depthI = ctx.depthData->I.data;
        m.hasDepth = true;
    }

    m.atoms.clear();
    m.polStage.clear();
    m.polSrc.clear();
    m.trans.clear();
    m.hostAtoms.clear();
    auto add_atoms = [&](std::vector<Atom*>& list, bool detailed)
    {
        for (Atom* a : list)
        {
            LwB200Atom fa{};
            fa.Nlevel = a->Nlevel;
            fa.Ntrans = a->Ntrans;
            fa.detailedStatic = detailed ? 1 : 0;
            fa.n = a->n.data;
            fa.nStar = a->nStar.data;
            fa.nTotal = a->nTotal.data;
            fa.vBroad = a->vBroad.data;
            fa.Gamma = detailed ? nullptr : a->Gamma.data;
            fa.C = a->C ? a->C.data : nullptr;
            fa.stages = a->stages ? a->stages.data : nullptr;
            m.trans.emplace_back();
            auto& tv = m.trans.back();
            for (Transition* t : a->trans)
            {
                if (t->hPrdCoeffs)
                    throw std::runtime_error("mali_full_precond_B200: hybrid PRD lines are not supported yet");
                LwB200Transition ft{};
                ft.type = t->type == LINE ? LWB200_LINE : LWB200_CONTINUUM;
                ft.i = t->i;
                ft.j = t->j;
                ft.Nblue = t->Nblue;
                ft.Nred = t->Nred;
                ft.Aji = t->Aji;
                ft.Bji = t->Bji;
                ft.Bij = t->Bij;
                ft.lambda0 = t->lambda0;
                ft.dopplerWidth = t->dopplerWidth;
                ft.wavelength = t->wavelength.data;
                ft.alpha = t->alpha.data;
                ft.phi = t->phi.data;
                ft.wphi = t->wphi.data;
                ft.rhoPrd = t->rhoPrd ? t->rhoPrd.data : nullptr;
                ft.aDamp = t->aDamp.data;
                ft.Rij = t->Rij.data;
                ft.Rji = t->Rji.data;
                ft.Qelast = t->Qelast ? t->Qelast.data : nullptr;
                if (t->type == LINE && t->polarised && t->phiQ)
                {
                    // the six extra profiles are separate arrays in the reference: staged contiguously
                    const size_t per = (size_t)(t->Nred - t->Nblue) * M * 2 * K;
                    m.polStage.emplace_back(6 * per);
                    ft.polProfiles = m.polStage.back().data();
                    m.polSrc.push_back({t, m.polStage.size() - 1});
                }
                tv.push_back(ft);
            }
            fa.trans = tv.data();
            m.atoms.push_back(fa);
            m.hostAtoms.push_back(a);
        }
    };
    add_atoms(ctx.activeAtoms, false);
    add_atoms(ctx.detailedAtoms, true);
    p.Natom = (int)m.atoms.size();
    // vector storage is final now: (re)bind the transition arrays
    {
        size_t i = 0;
        for (auto& tv : m.trans)
            m.atoms[i++].trans = tv.data();
    }
    p.atoms = m.atoms.data();

    int device = 0;
    if (const char* env = std::getenv("LWB200_DEVICE"))
        device = std::atoi(env);
    check(lwb200_create(&p, device, &m.dev), "lwb200_create");
    for (size_t i = 0; i < m.hostAtoms.size(); ++i)
        g_atoms[m.hostAtoms[i]] = {&m, (int)i};
    m.uploadedStatic = false;
}

Mirror& mirror_for(Context& ctx)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_mirrors.find(&ctx);
    const bool wantDepth = ctx.depthData && ctx.depthData->fill;
    if (it != g_mirrors.end())
    {
        Mirror& m = *it->second;
        // rebuild when the formal solver changed or depth data appeared
        if (m.dev && m.solver == solver_from_name(ctx.formalSolver.name) && (!wantDepth || m.hasDepth))
            return m;
        destroy_mirror(&m);
        build_mirror(ctx, m);
        return m;
    }
    auto m = std::make_unique<Mirror>();
    build_mirror(ctx, *m);
    Mirror& ref = *m;
    g_mirrors[&ctx] = std::move(m);
    ctx.methodScratch = &ref;
    return ref;
}

// Bring the device mirror up to date with whatever the host changed since the
// last call.  Small per-iteration arrays always travel; the large ones only when
// their fingerprint changed.
void sync_inputs(Context& ctx, Mirror& m, bool withGamma)
{
    const LwB200Problem& p = m.prob;
    const size_t K = p.Nspace, L = p.Nspect, M = p.Nrays;
    uint32_t mask = LWB200_POPS | LWB200_NSTAR | (withGamma ? LWB200_GAMMA : 0);
    uint64_t fa = 1469598103934665603ULL, fb = fa, fpf = fa;
    fa = fingerprint(fa, p.height, K);
    fa = fingerprint(fa, p.temperature, K);
    fa = fingerprint(fa, p.lowerBcData, p.NlowerBcMu ? L * p.NlowerBcMu : 0);
    fa = fingerprint(fa, p.upperBcData, p.NupperBcMu ? L * p.NupperBcMu : 0);
    fb = fingerprint(fb, p.chiBg, L * K);
    fb = fingerprint(fb, p.etaBg, L * K);
    fb = fingerprint(fb, p.scaBg, L * K);
    for (const LwB200Atom& a : m.atoms)
        for (int kr = 0; kr < a.Ntrans; ++kr)
        {
            const LwB200Transition& t = a.trans[kr];
            if (t.type != LWB200_LINE)
                continue;
            const size_t Nl = t.Nred - t.Nblue;
            fpf = fingerprint(fpf, t.phi, Nl * M * 2 * K);
            fpf = fingerprint(fpf, t.wphi, K);
            if (t.rhoPrd)
                fpf = fingerprint(fpf, t.rhoPrd, Nl * K);
        }
    const uint64_t fj = fingerprint(1469598103934665603ULL, p.J, L * K);
    if (!m.uploadedStatic || fa != m.fpAtmos)
        mask |= LWB200_ATMOS;
    if (!m.uploadedStatic || fb != m.fpBackground)
        mask |= LWB200_BACKGR;
    if (!m.uploadedStatic || fpf != m.fpProfiles)
        mask |= LWB200_PROFILE;
    if (!m.uploadedStatic || fj != m.fpJ)
        mask |= LWB200_JBAR;
    check(lwb200_upload(m.dev, mask), "lwb200_upload");
    m.fpAtmos = fa;
    m.fpBackground = fb;
    m.fpProfiles = fpf;
    m.uploadedStatic = true;
    (void)ctx;
}

IterationResult b200_fs_iter(Context& ctx, bool lambdaIterate, ExtraParams params)
{
    Mirror& m = mirror_for(ctx);
    sync_inputs(ctx, m, true);
    // one host synchronisation per call: dJ and J / I travel with the stream
    uint32_t flags = (lambdaIterate ? LWB200_LAMBDA_ITERATE : 0) | LWB200_FETCH_EARLY | LWB200_DJ_ASYNC;
    const bool storeDepth = ctx.depthData && ctx.depthData->fill;
    if (storeDepth)
        flags |= LWB200_STORE_DEPTH;
    if (params.contains("lwb200_general_kernel"))
        flags |= LWB200_GENERAL_KERNEL;
    double dJMax = 0.0;
    int64_t dJIdx = 0;
    check(lwb200_fs_iter(m.dev, flags, nullptr, nullptr), "lwb200_fs_iter");
    check(lwb200_download(m.dev, LWB200_ITER_OUTPUTS | (storeDepth ? LWB200_DEPTH : 0)), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    check(lwb200_last_dj(m.dev, &dJMax, &dJIdx), "lwb200_last_dj");
    m.fpJ = fingerprint(1469598103934665603ULL, m.prob.J, (size_t)m.prob.Nspect * m.prob.Nspace);
    IterationResult result{};
    result.updatedJ = true;
    result.dJMax = dJMax;
    result.dJMaxIdx = (int)dJIdx;
    return result;
}

// FsIterationFns::redistribute_prd (LwFormalInterface.hpp:118; called from
// redistribute_prd_lines, Prd.cpp:648-653): angle-averaged PRD on the device.  The rates of
// the last fs_iter are taken from the host (they may have been touched since).
IterationResult b200_redistribute_prd(Context& ctx, int maxIter, f64 tol, ExtraParams params)
{
    Mirror& m = mirror_for(ctx);
    bool includeDetailed = false;
    if (params.contains("include_detailed_atoms"))
        includeDetailed = params.get_as<bool>("include_detailed_atoms");
    int nLines = 0;
    for (const LwB200Atom& a : m.atoms)
        for (int kr = 0; kr < a.Ntrans; ++kr)
            if (a.trans[kr].rhoPrd && (!a.detailedStatic || includeDetailed))
            {
                if (!a.trans[kr].Qelast || !a.C)
                    return redistribute_prd_lines_scalar(ctx, maxIter, tol, params); // not ours to guess
                ++nLines;
            }
    if (nLines == 0)
        return IterationResult{};
    sync_inputs(ctx, m, false);
    check(lwb200_upload(m.dev, LWB200_PRD | LWB200_RATES), "lwb200_upload");
    std::vector<f64> dRho((size_t)maxIter * nLines, 0.0), dJ(maxIter, 0.0);
    std::vector<int32_t> dRhoIdx((size_t)maxIter * nLines, 0);
    std::vector<int64_t> dJIdx(maxIter, 0);
    int32_t nIter = 0;
    check(lwb200_redistribute_prd(m.dev, maxIter, tol, includeDetailed ? 1 : 0, &nIter, dRho.data(),
                                  dRhoIdx.data(), dJ.data(), dJIdx.data()),
          "lwb200_redistribute_prd");
    check(lwb200_download(m.dev, LWB200_PRD | LWB200_JBAR | LWB200_INTENS | LWB200_RATES), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    m.fpJ = fingerprint(1469598103934665603ULL, m.prob.J, (size_t)m.prob.Nspect * m.prob.Nspace);
    IterationResult result{};
    result.updatedRho = true;
    result.updatedJPrd = true;
    result.NprdSubIter = nIter;
    for (int it = 0; it < nIter; ++it)
    {
        for (int q = 0; q < nLines; ++q)
        {
            result.dRho.push_back(dRho[(size_t)it * nLines + q]);
            result.dRhoMaxIdx.push_back(dRhoIdx[(size_t)it * nLines + q]);
        }
        result.dJPrdMax.push_back(dJ[it]);
        result.dJPrdMaxIdx.push_back((int)dJIdx[it]);
    }
    return result;
}

// FsIterationFns::full_stokes_fs (LwFormalInterface.hpp:117): polarised formal solution.
IterationResult b200_full_stokes_fs(Context& ctx, bool updateJ, bool upOnly, ExtraParams params)
{
    if (params.contains("J20"))
        return formal_sol_full_stokes_impl(ctx, updateJ, upOnly, params); // J20 is not handled on the device
    size_t nPol = 0;
    for (auto* list : {&ctx.activeAtoms, &ctx.detailedAtoms})
        for (Atom* a : *list)
            for (Transition* t : a->trans)
                nPol += (t->type == LINE && t->polarised && t->phiQ) ? 1 : 0;
    if (nPol == 0 || !ctx.spect->Quv)
        return formal_sol_full_stokes_impl(ctx, updateJ, upOnly, params);
    {
        // setup_stokes() allocates the polarised profiles after the Context (and possibly our
        // mirror) was made: a mirror that does not know them is rebuilt
        std::unique_lock<std::mutex> lock(g_mutex);
        auto it = g_mirrors.find(&ctx);
        if (it != g_mirrors.end() && it->second->polSrc.size() != nPol)
        {
            destroy_mirror(it->second.get());
            g_mirrors.erase(it);
            ctx.methodScratch = nullptr;
        }
    }
    Mirror& m = mirror_for(ctx);
    sync_inputs(ctx, m, false);
    // (re)stage the polarised profiles: setup_stokes(recompute) rewrites them in place
    for (auto& src : m.polSrc)
    {
        const Transition* t = src.first;
        const size_t per = (size_t)(t->Nred - t->Nblue) * m.prob.Nrays * 2 * m.prob.Nspace;
        const f64* arrs[6] = {t->phiQ.data, t->phiU.data, t->phiV.data, t->psiQ.data, t->psiU.data, t->psiV.data};
        for (int a = 0; a < 6; ++a)
            std::memcpy(m.polStage[src.second].data() + a * per, arrs[a], per * sizeof(f64));
    }
    check(lwb200_upload(m.dev, LWB200_STOKES | LWB200_PROFILE), "lwb200_upload");
    double dJMax = 0.0;
    int64_t dJIdx = 0;
    check(lwb200_formal_sol_full_stokes(m.dev, updateJ ? 1 : 0, upOnly ? 1 : 0, &dJMax, &dJIdx),
          "lwb200_formal_sol_full_stokes");
    check(lwb200_download(m.dev, LWB200_INTENS | LWB200_STOKES | (updateJ ? LWB200_JBAR : 0)), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    if (updateJ)
        m.fpJ = fingerprint(1469598103934665603ULL, m.prob.J, (size_t)m.prob.Nspect * m.prob.Nspace);
    IterationResult result{};
    result.updatedJ = updateJ;
    if (updateJ)
    {
        result.dJMax = dJMax;
        result.dJMaxIdx = (int)dJIdx;
    }
    return result;
}

IterationResult b200_simple_fs(Context& ctx, bool upOnly, ExtraParams params)
{
    (void)params;
    Mirror& m = mirror_for(ctx);
    sync_inputs(ctx, m, false);
    check(lwb200_formal_sol(m.dev, upOnly ? 1 : 0), "lwb200_formal_sol");
    check(lwb200_download(m.dev, LWB200_INTENS), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    return IterationResult{};
}

void b200_stat_eq(Atom* atom, ExtraParams params, int spaceStart, int spaceEnd)
{
    Mirror* m = nullptr;
    int idx = -1;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_atoms.find(atom);
        if (it != g_atoms.end())
        {
            m = it->second.first;
            idx = it->second.second;
        }
    }
    if (!m || !m->dev)
    {
        // an atom this plugin has never seen (no fs_iter yet on its Context): the core's path
        stat_eq_impl(atom, params, spaceStart, spaceEnd);
        return;
    }
    check(lwb200_upload(m->dev, LWB200_POPS | LWB200_GAMMA_FINAL), "lwb200_upload");
    check(lwb200_stat_eq_async(m->dev, idx, spaceStart, spaceEnd), "lwb200_stat_eq_async");
    check(lwb200_download(m->dev, LWB200_POPS), "lwb200_download");
    check(lwb200_sync(m->dev), "lwb200_sync");
    int32_t nSingular = 0;
    if (lwb200_last_singular(m->dev, &nSingular) != 0)
    {
        if (nSingular > 0)
            throw std::runtime_error("Singular Matrix"); // as lu_decompose does, LuSolve.cpp:22-23
        raise("lwb200_stat_eq_async");
    }
}

// FsIterationFns::time_dep_update (LwFormalInterface.hpp:120): backward-Euler step of one atom.
void b200_time_dep_update(Atom* atom, F64View2D nOld, f64 dt, ExtraParams params, int spaceStart, int spaceEnd)
{
    Mirror* m = nullptr;
    int idx = -1;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_atoms.find(atom);
        if (it != g_atoms.end())
        {
            m = it->second.first;
            idx = it->second.second;
        }
    }
    if (!m || !m->dev)
    {
        time_dependent_update_impl(atom, nOld, dt, params, spaceStart, spaceEnd);
        return;
    }
    check(lwb200_upload(m->dev, LWB200_GAMMA_FINAL), "lwb200_upload");
    int32_t nSingular = 0;
    if (lwb200_time_dep_update(m->dev, idx, nOld.data, dt, spaceStart, spaceEnd, &nSingular) != 0)
    {
        if (nSingular > 0)
            throw std::runtime_error("Singular Matrix");
        raise("lwb200_time_dep_update");
    }
    check(lwb200_download(m->dev, LWB200_POPS), "lwb200_download");
    check(lwb200_sync(m->dev), "lwb200_sync");
}

// FsIterationFns::nr_post_update (LwFormalInterface.hpp:121-125): Newton-Raphson step with charge conservation.
void b200_nr_post_update(Context& ctx, std::vector<Atom*>* atoms, const std::vector<F64View3D>& dC,
                         F64View backgroundNe, const NrTimeDependentData& timeDepData, f64 crswVal,
                         ExtraParams params, int spaceStart, int spaceEnd)
{
    Mirror* m = nullptr;
    std::vector<int32_t> idx;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        for (Atom* a : *atoms)
        {
            auto it = g_atoms.find(a);
            if (it == g_atoms.end() || (m && it->second.first != m))
            {
                m = nullptr;
                break;
            }
            m = it->second.first;
            idx.push_back(it->second.second);
        }
    }
    if (!m || !m->dev || !m->prob.ne || idx.size() != atoms->size())
    {
        nr_post_update_impl(ctx, atoms, dC, backgroundNe, timeDepData, crswVal, params, spaceStart, spaceEnd);
        return;
    }
    std::vector<const double*> dCp, prevP;
    for (const auto& v : dC)
        dCp.push_back(v.data);
    for (const auto& v : timeDepData.nPrev)
        prevP.push_back(v.data);
    LwB200NrUpdate u{};
    u.Natom = (int32_t)idx.size();
    u.timeDependent = timeDepData.nPrev.size() != 0 ? 1 : 0;
    u.atomIdx = idx.data();
    u.dC = dCp.empty() ? nullptr : dCp.data();
    u.backgroundNe = backgroundNe.data;
    u.nPrev = prevP.empty() ? nullptr : prevP.data();
    u.dt = timeDepData.dt;
    u.crswVal = crswVal;
    check(lwb200_upload(m->dev, LWB200_POPS | LWB200_GAMMA_FINAL), "lwb200_upload");
    int32_t nSingular = 0;
    if (lwb200_nr_post_update(m->dev, &u, spaceStart, spaceEnd, &nSingular) != 0)
    {
        if (nSingular > 0)
            throw std::runtime_error("Singular Matrix");
        raise("lwb200_nr_post_update");
    }
    check(lwb200_download(m->dev, LWB200_POPS), "lwb200_download");
    check(lwb200_sync(m->dev), "lwb200_sync");
}

void b200_alloc_global_scratch(Context* ctx)
{
    // Called from ThreadData::initialise (ThreadStorage.cpp:484-493), i.e. BEFORE
    // compute_profiles() in LwContext.__init__: nothing can be uploaded yet, and the
    // formal solver may still change.  The mirror is therefore created lazily on first
    // use; here we only make sure no stale mirror survives an update_threads().
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_mirrors.find(ctx);
    if (it != g_mirrors.end())
    {
        destroy_mirror(it->second.get());
        g_mirrors.erase(it);
    }
    ctx->methodScratch = nullptr;
}

void b200_free_global_scratch(Context* ctx)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_mirrors.find(ctx);
    if (it != g_mirrors.end())
    {
        destroy_mirror(it->second.get());
        g_mirrors.erase(it);
    }
    ctx->methodScratch = nullptr;
}
} // namespace

extern "C"
{
FsIterationFns fs_iteration_fns_provider()
{
    return FsIterationFns{
        1,     // Ndim: 1D atmospheres only
        true,  // dimensionSpecific
        true,  // respectsFormalSolver (linear / besser / bezier3 kernels by name)
        true,  // defaultPerAtomStorage (the delegated core functions use it)
        true,  // defaultWlaGijStorage
        "mali_full_precond_B200",
        b200_fs_iter,
        b200_simple_fs,
        b200_full_stokes_fs,
        b200_redistribute_prd,
        b200_stat_eq,
        b200_time_dep_update,
        b200_nr_post_update,
        nullptr, // alloc_per_atom
        nullptr, // free_per_atom
        nullptr, // alloc_per_trans
        nullptr, // free_per_trans
        b200_alloc_global_scratch,
        b200_free_global_scratch,
        nullptr  // accumulate_over_threads
    };
}

// The per-ray LwFsFn hook cannot feed a GPU (one ray per call on host vectors,
// LwFormalInterface.hpp:33-43); it is exported for API completeness with the
// core's own host solver behind it.  Selecting it (or any of the three 1D
// solver names) together with `mali_full_precond_B200` picks the matching
// device kernel.
FormalSolver fs_provider()
{
    return FormalSolver{LwInternal::piecewise_bezier3_1d, 1, 1, "piecewise_bezier3_1d_b200"};
}
}
