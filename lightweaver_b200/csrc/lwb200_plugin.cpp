// lwb200_plugin.cpp -- the Lightweaver plugin shim (layer 1 of the boundary).
//
// Exports the two runtime-loadable provider symbols Lightweaver looks up with
// dlopen/dlsym (Source/LwFormalInterfacePosix.hpp:12-37):
//
//   extern "C" FsIterationFns fs_iteration_fns_provider();   (Source/FormalInterface.cpp:62-81)
//   extern "C" FormalSolver   fs_provider();                 (Source/FormalInterface.cpp:9-28)
//
// and implements EVERY function slot of FsIterationFns (Source/LwFormalInterface.hpp:110-134) --
// fs_iter, simple_fs, full_stokes_fs, redistribute_prd, stat_eq, time_dep_update, nr_post_update and
// the global-scratch hooks -- by marshalling the reference's Context into the flat LwB200Problem of
// include/lwb200.h and calling the C-ABI of liblwb200.so.  Everything below this file is plain C.
//
// This translation unit is compiled against the reference HEADERS (the provider structs, Context&
// and ExtraParams cross the boundary as C++ types) but links NO reference object code: there is no
// CPU fallback behind any slot.  A set-up the device path does not handle raises std::runtime_error
// with the reason (Lightweaver turns it into a Python exception, LwMiddleLayer.pyx:336-350); the
// formal solver behind fs_provider is this file's own host code.
//
// Use from Python:
//   lw.LwCompiled.FsIterationSchemes.load_fns_from_path('liblwb200_plugin.so')
//   ctx = lw.Context(..., fsIterScheme='mali_full_precond_B200', Nthreads=1)
// Context holds the reference's thread pool by value, so its headers bring in the inline virtual
// interface of the vendored task scheduler (TaskScheduler.h), whose vtables name three out-of-line
// members.  This file never schedules a task: the members are declared hidden and given local bodies
// below, so the plugin carries no undefined reference into (and no code from) the reference's libraries.
#define ENKITS_API __attribute__((visibility("hidden")))
#include "Lightweaver.hpp"
#include "lwb200.h"

#include <algorithm>
#include <cmath>
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

namespace enki
{
void TaskScheduler::TaskComplete(ICompletable*, bool, uint32_t) { std::abort(); }
void TaskScheduler::AddTaskSetToPipeInt(ITaskSet*, uint32_t) { std::abort(); }
void TaskScheduler::AddPinnedTaskInt(IPinnedTask*) { std::abort(); }
} // namespace enki

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
    // hybrid PRD: the reference's tables (Spectrum::JCoeffs, Transition::hPrdCoeffs ...) flattened
    bool hybrid = false;
    LwB200HybridPrd hprd{};
    std::vector<int32_t> hPrdLaOfLa_, prdLaOfLa_, JCoeffIdx_, lineAtom_, lineTrans_, rhoI0_;
    std::vector<int64_t> JCoeffOff_, rhoCoefOff_;
    std::vector<double> JCoeffFrac_, rhoFrac_;
    const void* hprdSig[2] = {nullptr, nullptr};
    uint64_t fpVlos = 0;
    bool uploadedStatic = false;
    bool hasDepth = false;
    bool zplane = false;
    uint64_t fpProfiles = 0, fpBackground = 0, fpAtmos = 0;
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

int device_index()
{
    if (const char* env = std::getenv("LWB200_DEVICE"))
        return std::atoi(env);
    return 0;
}

// ---------------------------------------------------------------------------
// Coherence.  Python mutates phi / background / atmosphere / J in place between calls without telling
// the plugin.  What travels when:
//   * populations, nStar / nTotal / vBroad, the prefill crsw*C and J: on EVERY call (small, or -- J --
//     cheaper to send than to check: it is rewritten by every call anyway);
//   * the atmosphere (height, temperature, ne, vlosMu, vturb, nHTot, boundary data): hashed over EVERY
//     element on every call (a response-function run perturbs ONE depth); when it changed, update_deps()
//     has run, and background and profiles are re-sent with it whatever their own fingerprints say;
//   * every line's wphi, aDamp and rhoPrd: hashed over every element.  wphi(k) is the reference's own
//     normalisation 1 / sum phi(la, mu, dir, k) w, recomputed whenever phi is (FormalScalar.cpp:106-134):
//     a changed profile at any depth shows there;
//   * the two big ones, phi and the background [Nspect][Nspace] arrays: a strided sample (odd stride:
//     it walks through every depth) on every call.  They only change through compute_profiles() /
//     update_background(), i.e. together with what is hashed fully above; LWB200_FINGERPRINT=full hashes
//     them over every element too (for hosts that edit them by hand), at ~1 ms per 5 MB.
// Hash: h = sum over 8 lanes of a rotate-add chain, bijective in every word, so a single changed element
// always changes it.
inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ULL;
    x ^= x >> 29;
    return x;
}

inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

uint64_t hash_block(const double* p, size_t n, size_t stride, uint64_t seed)
{
    uint64_t h[8];
    for (int l = 0; l < 8; ++l)
        h[l] = seed + l * 0x9e3779b97f4a7c15ULL;
    size_t i = 0;
    if (stride == 1)
    {
        for (; i + 8 <= n; i += 8)
        {
            uint64_t w[8];
            std::memcpy(w, p + i, 64);
            for (int l = 0; l < 8; ++l)
                h[l] = rotl64(h[l], 29) + w[l];
        }
    }
    for (size_t q = 0; i < n; i += stride, ++q)
    {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h[q & 7] = rotl64(h[q & 7], 29) + w;
    }
    uint64_t r = seed;
    for (int l = 0; l < 8; ++l)
        r = mix64(r + h[l]);
    return r;
}

bool fingerprint_full()
{
    const char* e = std::getenv("LWB200_FINGERPRINT"); // (read per call: a host may switch it at run time)
    return e && std::strcmp(e, "full") == 0;
}

// every element of the array
uint64_t fingerprint(uint64_t h, const double* p, size_t n)
{
    if (!p || n == 0)
        return h;
    return hash_block(p, n, 1, mix64(h + 0x2545f4914f6cdd1dULL));
}

// a strided sample of a large array (see above), or all of it under LWB200_FINGERPRINT=full
uint64_t fingerprint_sampled(uint64_t h, const double* p, size_t n)
{
    if (!p || n == 0)
        return h;
    if (fingerprint_full() || n <= 8192)
        return fingerprint(h, p, n);
    const size_t stride = (n / 1024) | 1; // (each sample is a cache line of its own: keep them few)
    h = hash_block(p, n, stride, mix64(h + 0x2545f4914f6cdd1dULL));
    return hash_block(p + n - 1, 1, 1, h);
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

    // hybrid PRD (configure_hprd_coeffs has run, Prd.cpp:697-946): flatten the reference's own tables
    m.hybrid = (bool)spect.JRest;
    p.hprd = nullptr;
    if (m.hybrid)
    {
        if (!spect.prdActive || !spect.hPrdActive || !atmos.vlosMu)
            throw std::runtime_error("mali_full_precond_B200: spect.JRest without the tables of configure_hprd_coeffs");
        const size_t NhPrd = spect.hPrdIdxs.size();
        m.prdLaOfLa_.assign(L, -1);
        m.hPrdLaOfLa_.assign(L, -1);
        for (int la = 0; la < L; ++la)
        {
            if (spect.prdActive(la))
                m.prdLaOfLa_[la] = spect.la_to_prdLa(la);
            if (spect.hPrdActive(la))
                m.hPrdLaOfLa_[la] = spect.la_to_hPrdLa(la);
        }
        m.JCoeffOff_.clear();
        m.JCoeffIdx_.clear();
        m.JCoeffFrac_.clear();
        for (size_t q = 0; q < NhPrd; ++q)
            for (int mu = 0; mu < M; ++mu)
                for (int toObs = 0; toObs < 2; ++toObs)
                    for (int k = 0; k < K; ++k)
                    {
                        m.JCoeffOff_.push_back((int64_t)m.JCoeffIdx_.size());
                        for (const auto& c : spect.JCoeffs(q, mu, toObs, k))
                        {
                            m.JCoeffIdx_.push_back(c.idx);
                            m.JCoeffFrac_.push_back(c.frac);
                        }
                    }
        m.JCoeffOff_.push_back((int64_t)m.JCoeffIdx_.size());
        if (m.JCoeffIdx_.empty())
        {
            m.JCoeffIdx_.push_back(0);
            m.JCoeffFrac_.push_back(0.0);
        }
        m.lineAtom_.clear();
        m.lineTrans_.clear();
        m.rhoCoefOff_.clear();
        m.rhoFrac_.clear();
        m.rhoI0_.clear();
        for (size_t ia = 0; ia < m.hostAtoms.size(); ++ia)
            for (size_t kr = 0; kr < m.hostAtoms[ia]->trans.size(); ++kr)
            {
                Transition* t = m.hostAtoms[ia]->trans[kr];
                if (!t->hPrdCoeffs)
                    continue;
                m.lineAtom_.push_back((int32_t)ia);
                m.lineTrans_.push_back((int32_t)kr);
                m.rhoCoefOff_.push_back((int64_t)m.rhoFrac_.size());
                const int Nl = t->Nred - t->Nblue;
                for (int lt = 0; lt < Nl; ++lt)
                    for (int mu = 0; mu < M; ++mu)
                        for (int toObs = 0; toObs < 2; ++toObs)
                            for (int k = 0; k < K; ++k)
                            {
                                const auto& c = t->hPrdCoeffs(lt, mu, toObs, k);
                                m.rhoFrac_.push_back(c.frac);
                                m.rhoI0_.push_back(c.i0);
                            }
            }
        LwB200HybridPrd& h = m.hprd;
        h = LwB200HybridPrd{};
        h.NprdLa = (int32_t)spect.JRest.shape(0);
        h.NhPrd = (int32_t)NhPrd;
        h.Nlines = (int32_t)m.lineAtom_.size();
        h.prdLaOfLa = m.prdLaOfLa_.data();
        h.hPrdLaOfLa = m.hPrdLaOfLa_.data();
        h.JRest = spect.JRest.data();
        h.JCoeffOff = m.JCoeffOff_.data();
        h.JCoeffIdx = m.JCoeffIdx_.data();
        h.JCoeffFrac = m.JCoeffFrac_.data();
        h.lineAtom = m.lineAtom_.data();
        h.lineTrans = m.lineTrans_.data();
        h.rhoCoefOff = m.rhoCoefOff_.data();
        h.rhoFrac = m.rhoFrac_.data();
        h.rhoI0 = m.rhoI0_.data();
        p.hprd = &h;
        m.hprdSig[0] = spect.JRest.data();
        m.hprdSig[1] = spect.JCoeffs.data();
        m.fpVlos = fingerprint(1469598103934665603ULL, atmos.vlosMu.data, (size_t)M * K);
    }

    check(lwb200_create(&p, device_index(), &m.dev), "lwb200_create");
    for (size_t i = 0; i < m.hostAtoms.size(); ++i)
        g_atoms[m.hostAtoms[i]] = {&m, (int)i};
    m.uploadedStatic = false;
    m.zplane = false;
}

Mirror& mirror_for(Context& ctx)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_mirrors.find(&ctx);
    const bool wantDepth = ctx.depthData && ctx.depthData->fill;
    if (it != g_mirrors.end())
    {
        Mirror& m = *it->second;
        // rebuild when the formal solver changed, depth data appeared, or configure_hprd_coeffs ran (again):
        // new tables, possibly another set of scattering wavelengths -- the plan depends on them
        Spectrum& sp = *ctx.spect;
        bool hprdSame = m.hybrid == (bool)sp.JRest;
        if (hprdSame && m.hybrid)
            hprdSame = m.hprdSig[0] == (const void*)sp.JRest.data() && m.hprdSig[1] == (const void*)sp.JCoeffs.data()
                && m.fpVlos == fingerprint(1469598103934665603ULL, ctx.atmos->vlosMu.data,
                                           (size_t)ctx.atmos->Nrays * ctx.atmos->Nspace);
        if (m.dev && m.solver == solver_from_name(ctx.formalSolver.name) && (!wantDepth || m.hasDepth) && hprdSame)
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

// Bring the device mirror up to date with whatever the host changed since the last call (see "Coherence").
void sync_inputs(Context& ctx, Mirror& m, bool withGamma)
{
    const LwB200Problem& p = m.prob;
    const Atmosphere& atmos = *ctx.atmos;
    const size_t K = p.Nspace, L = p.Nspect, M = p.Nrays;
    uint32_t mask = LWB200_POPS | LWB200_NSTAR | LWB200_JBAR | (withGamma ? LWB200_GAMMA : 0);
    uint64_t fa = 1469598103934665603ULL, fb = fa, fpf = fa;
    fa = fingerprint(fa, p.height, K);
    fa = fingerprint(fa, p.temperature, K);
    fa = fingerprint(fa, p.vlosMu, p.vlosMu ? M * K : 0);
    fa = fingerprint(fa, p.ne, p.ne ? K : 0);
    fa = fingerprint(fa, atmos.vturb.data, atmos.vturb ? K : 0);
    fa = fingerprint(fa, atmos.nHTot.data, atmos.nHTot ? K : 0);
    fa = fingerprint(fa, atmos.B.data, atmos.B ? K : 0);
    fa = fingerprint(fa, p.lowerBcData, p.NlowerBcMu ? L * p.NlowerBcMu : 0);
    fa = fingerprint(fa, p.upperBcData, p.NupperBcMu ? L * p.NupperBcMu : 0);
    fb = fingerprint_sampled(fb, p.chiBg, L * K);
    fb = fingerprint_sampled(fb, p.etaBg, L * K);
    fb = fingerprint_sampled(fb, p.scaBg, L * K);
    for (const LwB200Atom& a : m.atoms)
        for (int kr = 0; kr < a.Ntrans; ++kr)
        {
            const LwB200Transition& t = a.trans[kr];
            if (t.type != LWB200_LINE)
                continue;
            const size_t Nl = t.Nred - t.Nblue;
            fpf = fingerprint(fpf, t.wphi, K);
            fpf = fingerprint(fpf, t.aDamp, t.aDamp ? K : 0);
            fpf = fingerprint_sampled(fpf, t.phi, Nl * M * 2 * K);
            if (t.rhoPrd)
                fpf = fingerprint(fpf, t.rhoPrd, Nl * K);
        }
    // a changed atmosphere means update_deps() ran: profiles and background are re-sent with it
    if (!m.uploadedStatic || fa != m.fpAtmos)
        mask |= LWB200_ATMOS | LWB200_BACKGR | LWB200_PROFILE;
    if (fb != m.fpBackground)
        mask |= LWB200_BACKGR;
    if (fpf != m.fpProfiles)
        mask |= LWB200_PROFILE;
    check(lwb200_upload(m.dev, mask), "lwb200_upload");
    m.fpAtmos = fa;
    m.fpBackground = fb;
    m.fpProfiles = fpf;
    m.uploadedStatic = true;
}

// ZPlaneDecomposition (SimdFullIterationTemplates.hpp:254-281): register / unregister the output views.
void bind_zplane(Mirror& m, ExtraParams& params)
{
    double *up = nullptr, *down = nullptr;
    if (params.contains("ZPlaneDecomposition"))
    {
        const size_t L = m.prob.Nspect, M = m.prob.Nrays;
        auto view = [&](const char* key) -> double*
        {
            if (!params.contains(key))
                return nullptr;
            F64View2D v = params.get_as<F64View2D>(key);
            if (!v)
                return nullptr;
            if ((size_t)v.shape(0) != L || (size_t)v.shape(1) != M)
                throw std::runtime_error(std::string("mali_full_precond_B200: ") + key + " must be [Nspect, Nrays]");
            return v.data;
        };
        up = view("ZPlaneUp");
        down = view("ZPlaneDown");
    }
    if (up || down || m.zplane)
        check(lwb200_set_zplane(m.dev, up, down), "lwb200_set_zplane");
    m.zplane = up || down;
}

IterationResult b200_fs_iter(Context& ctx, bool lambdaIterate, ExtraParams params)
{
    Mirror& m = mirror_for(ctx);
    bind_zplane(m, params);
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
    check(lwb200_download(m.dev, LWB200_ITER_OUTPUTS | (storeDepth ? LWB200_DEPTH : 0) | (m.zplane ? LWB200_ZPLANE : 0)
                                     | (m.hybrid ? LWB200_PRD : 0)), // (hybrid PRD: spect.JRest)
          "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    check(lwb200_last_dj(m.dev, &dJMax, &dJIdx), "lwb200_last_dj");
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
                    throw std::runtime_error("mali_full_precond_B200: a PRD line without Qelast, or its atom "
                                             "without collisional rates C, cannot be redistributed");
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
    // the "J20" extra parameter (FormalStokes.cpp:676-681): the anisotropy array, read and rewritten in place
    double* j20 = nullptr;
    if (params.contains("J20"))
    {
        F64View2D v = params.get_as<F64View2D>("J20");
        if (!v || v.shape(0) != ctx.spect->wavelength.shape(0) || v.shape(1) != ctx.atmos->Nspace)
            throw std::runtime_error("mali_full_precond_B200: J20 must be [Nspect, Nspace]");
        j20 = v.data;
    }
    if (!ctx.atmos->B)
        throw std::runtime_error("Magnetic field required"); // as formal_sol_full_stokes_impl, FormalStokes.cpp:670-671
    size_t nPol = 0;
    for (auto* list : {&ctx.activeAtoms, &ctx.detailedAtoms})
        for (Atom* a : *list)
            for (Transition* t : a->trans)
                nPol += (t->type == LINE && t->polarised && t->phiQ) ? 1 : 0;
    if ((nPol == 0 && !j20) || !ctx.spect->Quv)
        throw std::runtime_error("mali_full_precond_B200: full-Stokes formal solution without a polarised line "
                                 "(call setup_stokes first) or without spect.Quv");
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
    check(lwb200_set_j20(m.dev, j20), "lwb200_set_j20");
    double dJMax = 0.0;
    int64_t dJIdx = 0;
    check(lwb200_formal_sol_full_stokes(m.dev, updateJ ? 1 : 0, upOnly ? 1 : 0, &dJMax, &dJIdx),
          "lwb200_formal_sol_full_stokes");
    check(lwb200_download(m.dev, LWB200_INTENS | LWB200_STOKES | (updateJ ? LWB200_JBAR : 0)), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
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
    Mirror& m = mirror_for(ctx);
    bind_zplane(m, params);
    sync_inputs(ctx, m, false);
    check(lwb200_formal_sol(m.dev, upOnly ? 1 : 0), "lwb200_formal_sol");
    check(lwb200_download(m.dev, LWB200_INTENS | (m.zplane ? LWB200_ZPLANE : 0)), "lwb200_download");
    check(lwb200_sync(m.dev), "lwb200_sync");
    return IterationResult{};
}

bool find_atom(Atom* atom, Mirror*& m, int& idx)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_atoms.find(atom);
    if (it == g_atoms.end() || !it->second.first->dev)
        return false;
    m = it->second.first;
    idx = it->second.second;
    return true;
}

// stat_eq / time_dep_update for an atom no device context knows yet (FsIterationFns::stat_eq gets an
// Atom* and nothing else, and a caller may update populations before the first formal solution of its
// Context, e.g. after gamma_matrices_escape_prob): the self-contained device solve.
void population_solve_standalone(Atom* atom, const f64* nOld, f64 dt, int spaceStart, int spaceEnd)
{
    const int N = atom->Nlevel, K = (int)atom->n.shape(1);
    int32_t nSingular = 0;
    if (lwb200_population_solve(device_index(), 1, N, K, atom->Gamma.data, atom->n.data, atom->nTotal.data, nOld, dt,
                                spaceStart, spaceEnd, &nSingular)
        != 0)
    {
        if (nSingular > 0)
            throw std::runtime_error("Singular Matrix"); // as lu_decompose does, LuSolve.cpp:22-23
        raise("lwb200_population_solve");
    }
}

void b200_stat_eq(Atom* atom, ExtraParams params, int spaceStart, int spaceEnd)
{
    (void)params;
    Mirror* m = nullptr;
    int idx = -1;
    if (!find_atom(atom, m, idx))
    {
        population_solve_standalone(atom, nullptr, 0.0, spaceStart, spaceEnd);
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
    (void)params;
    Mirror* m = nullptr;
    int idx = -1;
    if (!find_atom(atom, m, idx))
    {
        population_solve_standalone(atom, nOld.data, dt, spaceStart, spaceEnd);
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
    (void)params;
    Mirror& mr = mirror_for(ctx); // (made here if this Context has not run a formal solution yet)
    Mirror* m = &mr;
    std::vector<int32_t> idx;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        for (Atom* a : *atoms)
        {
            auto it = g_atoms.find(a);
            if (it == g_atoms.end() || it->second.first != m)
                throw std::runtime_error("mali_full_precond_B200: nr_post_update on an atom that is not an active "
                                         "atom of this Context");
            idx.push_back(it->second.second);
        }
    }
    if (!m->prob.ne)
        throw std::runtime_error("mali_full_precond_B200: nr_post_update needs atmos.ne");
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
    check(lwb200_upload(m->dev, LWB200_POPS | LWB200_NSTAR | LWB200_GAMMA_FINAL), "lwb200_upload");
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

// ---------------------------------------------------------------------------
// The host formal solver behind fs_provider: piecewise Bezier3 short characteristics for ONE ray on
// host vectors, in the two-phase form of the device solver (csrc/lwb200_fsm.cuh): first everything
// that does not depend on the direction of the ray, in array-forward orientation -- the Steffen
// derivatives of chi, the Bezier optical depth of every interval, the Steffen derivative of S in
// optical depth -- then the sweep along the ray, whose last point is piecewise linear.
// Arithmetic: Source/FormalScalar.cpp:209-325, :535-600, Bezier.hpp:58-127, LwInternal.hpp:90-110.
inline double steffen_host(double dsUw, double dsDw, double slUw, double sl0)
{
    // slUw: slope over the upwind interval (length dsUw), sl0: over the downwind one (Bezier.hpp:58-65)
    const double P0 = std::fabs((slUw * dsDw + sl0 * dsUw) / (dsDw + dsUw));
    return (std::copysign(1.0, sl0) + std::copysign(1.0, slUw))
           * std::fmin(std::fabs(slUw), std::fmin(std::fabs(sl0), 0.5 * P0));
}

struct HostRayScratch
{
    std::vector<double> ds, dChi, dtau, dS;
    void resize(size_t K)
    {
        if (ds.size() < K)
        {
            ds.resize(K);
            dChi.resize(K);
            dtau.resize(K);
            dS.resize(K);
        }
    }
};

void host_bezier3_ray(int K, const double* height, const double* chi, const double* S, double zmu, bool toObs,
                      double Iupw, double* I, double* Psi)
{
    thread_local HostRayScratch w;
    w.resize((size_t)K);
    double *ds = w.ds.data(), *dChi = w.dChi.data(), *dtau = w.dtau.data(), *dS = w.dS.data();
    // ---- direction-independent phase, forward (k ascending) orientation
    for (int k = 0; k + 1 < K; ++k)
        ds[k] = std::fabs(height[k] - height[k + 1]) * zmu;
    dChi[0] = (chi[1] - chi[0]) / ds[0];
    dChi[K - 1] = (chi[K - 1] - chi[K - 2]) / ds[K - 2];
    for (int k = 1; k + 1 < K; ++k)
        dChi[k] = steffen_host(ds[k - 1], ds[k], (chi[k] - chi[k - 1]) / ds[k - 1], (chi[k + 1] - chi[k]) / ds[k]);
    for (int k = 0; k + 1 < K; ++k)
    {
        const double cA = chi[k] + (ds[k] / 3.0) * dChi[k];
        const double cB = chi[k + 1] - (ds[k] / 3.0) * dChi[k + 1];
        dtau[k] = ds[k] * (chi[k] + chi[k + 1] + cA + cB) * 0.25;
    }
    dS[0] = (S[1] - S[0]) / dtau[0];
    dS[K - 1] = (S[K - 1] - S[K - 2]) / dtau[K - 2];
    for (int k = 1; k + 1 < K; ++k)
        dS[k] = steffen_host(dtau[k - 1], dtau[k], (S[k] - S[k - 1]) / dtau[k - 1], (S[k + 1] - S[k]) / dtau[k]);

    // ---- the sweep
    const int dk = toObs ? -1 : 1, kS = toObs ? K - 1 : 0, kE = toObs ? 0 : K - 1;
    const double sgn = toObs ? -1.0 : 1.0; // derivative along the ray = sgn * forward derivative
    double Iprev = Iupw;
    I[kS] = Iupw;
    if (Psi)
        Psi[kS] = 0.0;
    for (int k = kS + dk; k != kE; k += dk)
    {
        const int ku = k - dk;
        const double dt = dtau[toObs ? k : k - 1];
        const double dt2 = dt * dt, dt3 = dt2 * dt;
        double alpha, beta, gamma, delta, edt;
        if (dt < 5e-2)
        {
            edt = 1.0 - dt + 0.5 * dt2 - dt3 / 6.0;
            alpha = 0.25 * dt - 0.2 * dt2 + dt3 / 12.0;
            beta = 0.25 * dt - 0.05 * dt2 + dt3 / 120.0;
            gamma = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
            delta = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
        }
        else
        {
            edt = dt > 30.0 ? 0.0 : std::exp(-dt);
            alpha = (6.0 - edt * (6.0 + 6.0 * dt + 3.0 * dt2 + dt3)) / dt3;
            beta = (6.0 * edt - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
            gamma = 3.0 * (2.0 * dt - 6.0 + edt * (6.0 + 4.0 * dt + dt2)) / dt3;
            delta = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edt * (3.0 + dt)) / dt3;
        }
        const double Cuw = S[ku] + (dt / 3.0) * (sgn * dS[ku]);
        const double C0 = S[k] - (dt / 3.0) * (sgn * dS[k]);
        Iprev = Iprev * edt + alpha * S[ku] + beta * S[k] + gamma * Cuw + delta * C0;
        I[k] = Iprev;
        if (Psi)
            Psi[k] = beta + delta;
    }
    {
        // last point: piecewise linear through w2()
        const int k = kE, ku = kE - dk;
        const double dt = 0.5 * zmu * (chi[k] + chi[ku]) * std::fabs(height[k] - height[ku]);
        double w0, w1;
        if (dt < 5.0E-4)
        {
            w0 = dt * (1.0 - 0.5 * dt);
            w1 = (dt * dt) * (0.5 - dt * (1.0 / 3.0));
        }
        else if (dt > 50.0)
            w0 = w1 = 1.0;
        else
        {
            const double e = std::exp(-dt);
            w0 = 1.0 - e;
            w1 = w0 - dt * e;
        }
        I[k] = (1.0 - w0) * Iprev + w0 * S[k] - w1 * ((S[k] - S[ku]) / dt);
        if (Psi)
            Psi[k] = w0 - w1 / dt;
    }
    if (Psi)
        for (int k = 0; k < K; ++k)
            Psi[k] /= chi[k];
}

inline double planck_host(double T, double lambda)
{
    namespace C = Constants;
    const double x = (C::HC / (C::KBoltzmann * C::NM_TO_M)) / lambda / T;
    const double pre = (2.0 * C::HC) / (C::NM_TO_M * C::NM_TO_M * C::NM_TO_M) / (lambda * lambda * lambda);
    return x <= 150.0 ? pre / (std::exp(x) - 1.0) : 0.0;
}

// LwFsFn (LwFormalInterface.hpp:33-34): boundary intensity as piecewise_bezier3_1d (FormalScalar.cpp:535-600).
void b200_piecewise_bezier3_1d(LwInternal::FormalData* fd, int la, int mu, bool toObs, const F64View1D& wave)
{
    Atmosphere* atmos = fd->atmos;
    const int K = atmos->Nspace;
    if (K < 3)
        throw std::runtime_error("piecewise_bezier3_1d_b200: needs at least three depth points");
    const double zmu = 1.0 / atmos->muz(mu);
    const double* h = atmos->height.data;
    const double* chi = fd->chi.data;
    const int kS = toObs ? K - 1 : 0, kN = toObs ? K - 2 : 1;
    const double dtauB = 0.5 * zmu * (chi[kS] + chi[kN]) * std::fabs(h[kS] - h[kN]);
    const AtmosphericBoundaryCondition& bc = toObs ? atmos->zLowerBc : atmos->zUpperBc;
    double Iupw = 0.0;
    if (bc.type == THERMALISED)
    {
        const double B0 = planck_host(atmos->temperature(kS), wave(la)), B1 = planck_host(atmos->temperature(kN), wave(la));
        Iupw = B0 - (B1 - B0) / dtauB;
    }
    else if (bc.type == CALLABLE)
    {
        const int muIdx = bc.idxs(mu, int(toObs));
        if (muIdx < 0)
            throw std::runtime_error("piecewise_bezier3_1d_b200: boundary condition index missing for this ray");
        Iupw = bc.bcData(la, muIdx, 0);
    }
    host_bezier3_ray(K, h, chi, fd->S.data, zmu, toObs, Iupw, fd->I.data, fd->Psi ? fd->Psi.data : nullptr);
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
        true,  // defaultPerAtomStorage
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
// LwFormalInterface.hpp:33-43); what it exports is this back end's own host Bezier3 solver, numerically
// the device solver's twin.  Selecting it (or any of the three 1D solver names) together with
// `mali_full_precond_B200` picks the matching device kernel.
FormalSolver fs_provider()
{
    return FormalSolver{b200_piecewise_bezier3_1d, 1, 1, "piecewise_bezier3_1d_b200"};
}
}
