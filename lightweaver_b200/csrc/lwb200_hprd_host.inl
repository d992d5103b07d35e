// lwb200_hprd_host.inl -- host side of hybrid PRD: the tables configure_hprd_coeffs builds
// (Source/Prd.cpp:697-946), made for a whole column stack from the problem's wavelength grid, PRD line
// ranges and velocity field.  Included by lwb200_api.cu (host code only).
//
// Same results as the reference's nested scans, organised by binary searches on the (ascending) wavelength
// grid: for a ray (mu, toObs) at depth k the Doppler factor maps the neighbours of a wavelength to
// [prevLambda, nextLambda]; the wavelengths of the PRD grid inside that window receive the ray's intensity
// with linear weights (JCoeffs), and a PRD line's rho is interpolated at the shifted wavelength (hPrdCoeffs).

namespace
{
struct HprdOwned
{
    std::vector<int32_t> prdLaOfLa, hPrdLaOfLa, JCoeffIdx, lineAtom, lineTrans, rhoI0;
    std::vector<int64_t> JCoeffOff, rhoCoefOff;
    std::vector<double> JCoeffFrac, rhoFrac, JRest;
};
// the arrays handed out by lwb200_configure_hprd, keyed by their prdLaOfLa pointer, until lwb200_free_hprd
std::mutex g_hprdMu;
std::map<const void*, std::unique_ptr<HprdOwned>> g_hprdReg;
}

extern "C"
{
int lwb200_configure_hprd(const LwB200Problem* p, int includeDetailed, LwB200HybridPrd* out)
{
    if (!p || !out)
        return fail("lwb200_configure_hprd: null argument");
    std::memset(out, 0, sizeof(*out));
    if (!p->vlosMu)
        return fail("lwb200_configure_hprd: the problem has no vlosMu");
    const int K = p->Nspace, M = p->Nrays, L = p->Nspect, Ncol = p->Ncol;
    constexpr double cLight = 2.99792458E+08; // Constants.hpp
    const double* wl = p->wavelength;
    auto own = std::make_unique<HprdOwned>();
    // PRD lines: active atoms first, then (on request) the detailed-static ones
    for (int pass = 0; pass < (includeDetailed ? 2 : 1); ++pass)
        for (int a = 0; a < p->Natom; ++a)
        {
            if ((p->atoms[a].detailedStatic != 0) != (pass == 1))
                continue;
            for (int kr = 0; kr < p->atoms[a].Ntrans; ++kr)
                if (p->atoms[a].trans[kr].type == LWB200_LINE && p->atoms[a].trans[kr].rhoPrd)
                {
                    own->lineAtom.push_back(a);
                    own->lineTrans.push_back(kr);
                }
        }
    const int nLines = (int)own->lineAtom.size();
    if (nLines == 0)
        return 0;
    // rows of JRest: wavelengths at which a PRD line is active, and their running count
    own->prdLaOfLa.assign(L, -1);
    std::vector<int> cum(L + 1, 0);
    int NprdLa = 0;
    for (int la = 0; la < L; ++la)
    {
        bool on = false;
        for (int q = 0; q < nLines; ++q)
        {
            const LwB200Transition& t = p->atoms[own->lineAtom[q]].trans[own->lineTrans[q]];
            on = on || (la >= t.Nblue && la < t.Nred);
        }
        if (on)
            own->prdLaOfLa[la] = NprdLa++;
        cum[la + 1] = cum[la] + (on ? 1 : 0);
    }
    auto prdActive = [&](int i) { return own->prdLaOfLa[i] >= 0; };
    // largest index <= hi whose wavelength is <= x (floor at `floorIdx`)
    auto last_le = [&](double x, int hi, int floorIdx) {
        const int i = (int)(std::upper_bound(wl, wl + hi + 1, x) - wl) - 1;
        return std::max(i, floorIdx);
    };
    own->hPrdLaOfLa.assign((size_t)Ncol * L, -1);
    std::vector<int> nh(Ncol, 0);
    int NhPrd = 0;
    for (int col = 0; col < Ncol; ++col)
    {
        const double* v = p->vlosMu + (size_t)col * M * K;
        for (int la = 0; la < L; ++la)
        {
            const int pi = std::max(la - 1, 0), ni = std::min(la + 1, L - 1);
            bool scat = false;
            for (int mu = 0; mu < M && !scat; ++mu)
                for (int dir = 0; dir < 2 && !scat; ++dir)
                    for (int k = 0; k < K && !scat; ++k)
                    {
                        const double fac = 1.0 + v[(size_t)mu * K + k] * (dir ? 1.0 : -1.0) / cLight;
                        const double lo = wl[pi] * fac, hi = wl[ni] * fac;
                        // the reference walks back from la to the first point at or below the window, then
                        // forward up to and including the first point beyond it
                        const int i0 = last_le(lo, la, 0);
                        int i1 = (int)(std::upper_bound(wl + i0, wl + L, hi) - wl); // first point beyond
                        i1 = std::min(i1, L - 1);
                        scat = cum[i1 + 1] - cum[i0] > 0;
                    }
            if (scat)
                own->hPrdLaOfLa[(size_t)col * L + la] = nh[col]++;
        }
        NhPrd = std::max(NhPrd, nh[col]);
    }
    // JCoeffs, CSR
    const size_t nRows = (size_t)Ncol * NhPrd * M * 2 * K;
    own->JCoeffOff.assign(nRows + 1, 0);
    for (int col = 0; col < Ncol; ++col)
    {
        const double* v = p->vlosMu + (size_t)col * M * K;
        for (int h = 0, la = 0; h < NhPrd; ++h)
        {
            // (rows beyond this column's own count stay empty)
            while (la < L && own->hPrdLaOfLa[(size_t)col * L + la] != h)
                ++la;
            for (int mu = 0; mu < M; ++mu)
                for (int dir = 0; dir < 2; ++dir)
                    for (int k = 0; k < K; ++k)
                    {
                        const size_t row = ((((size_t)col * NhPrd + h) * M + mu) * 2 + dir) * K + k;
                        own->JCoeffOff[row] = (int64_t)own->JCoeffIdx.size();
                        if (la >= L)
                            continue;
                        const double fac = 1.0 + v[(size_t)mu * K + k] * (dir ? 1.0 : -1.0) / cLight;
                        const int pi = std::max(la - 1, 0), ni = std::min(la + 1, L - 1);
                        const double lo = wl[pi] * fac, rest = wl[la] * fac, hi = wl[ni] * fac;
                        auto push = [&](int i, double f) {
                            own->JCoeffIdx.push_back(own->prdLaOfLa[i]);
                            own->JCoeffFrac.push_back(f);
                        };
                        bool lower = true, upper = true;
                        if (pi == la)
                        {
                            // first grid point: constant extrapolation below
                            lower = false;
                            for (int i = 0; i < L && wl[i] <= rest && prdActive(i); ++i)
                                push(i, 1.0);
                        }
                        else if (ni == la)
                        {
                            upper = false;
                            for (int i = L - 1; i >= 0 && wl[i] > rest && prdActive(i); --i)
                                push(i, 1.0);
                        }
                        for (int i = last_le(lo, la, 0); i < L && wl[i] <= hi; ++i)
                        {
                            if (!prdActive(i))
                                continue;
                            const double x = wl[i];
                            if (lower && x > lo && x <= rest)
                                push(i, (x - lo) / (rest - lo));
                            else if (upper && x > rest && x < hi)
                                push(i, 1.0 - (x - rest) / (hi - rest));
                        }
                    }
            if (la < L)
                ++la;
        }
    }
    own->JCoeffOff[nRows] = (int64_t)own->JCoeffIdx.size();
    // hPrdCoeffs of every PRD line
    int64_t tot = 0;
    for (int q = 0; q < nLines; ++q)
    {
        const LwB200Transition& t = p->atoms[own->lineAtom[q]].trans[own->lineTrans[q]];
        own->rhoCoefOff.push_back(tot);
        tot += (int64_t)Ncol * (t.Nred - t.Nblue) * M * 2 * K;
    }
    own->rhoFrac.resize(tot);
    own->rhoI0.resize(tot);
    for (int q = 0; q < nLines; ++q)
    {
        const LwB200Transition& t = p->atoms[own->lineAtom[q]].trans[own->lineTrans[q]];
        const int Nl = t.Nred - t.Nblue;
        const double* w = t.wavelength;
        size_t o = (size_t)own->rhoCoefOff[q];
        for (int col = 0; col < Ncol; ++col)
        {
            const double* v = p->vlosMu + (size_t)col * M * K;
            for (int lt = 0; lt < Nl; ++lt)
                for (int mu = 0; mu < M; ++mu)
                    for (int dir = 0; dir < 2; ++dir)
                        for (int k = 0; k < K; ++k, ++o)
                        {
                            const double rest = w[lt] * (1.0 + v[(size_t)mu * K + k] * (dir ? 1.0 : -1.0) / cLight);
                            if (rest <= w[0])
                            {
                                own->rhoFrac[o] = 0.0;
                                own->rhoI0[o] = 0;
                            }
                            else if (rest >= w[Nl - 1])
                            {
                                own->rhoFrac[o] = 1.0;
                                own->rhoI0[o] = Nl - 2;
                            }
                            else
                            {
                                const int i0 = (int)(std::upper_bound(w, w + Nl, rest) - w) - 1;
                                own->rhoFrac[o] = (rest - w[i0]) / (w[i0 + 1] - w[i0]);
                                own->rhoI0[o] = i0;
                            }
                        }
        }
    }
    own->JRest.assign((size_t)Ncol * NprdLa * K, 0.0);
    if (own->JCoeffIdx.empty())
    {
        own->JCoeffIdx.push_back(0);
        own->JCoeffFrac.push_back(0.0);
    }
    out->NprdLa = NprdLa;
    out->NhPrd = NhPrd;
    out->Nlines = nLines;
    out->prdLaOfLa = own->prdLaOfLa.data();
    out->hPrdLaOfLa = own->hPrdLaOfLa.data();
    out->JRest = own->JRest.data();
    out->JCoeffOff = own->JCoeffOff.data();
    out->JCoeffIdx = own->JCoeffIdx.data();
    out->JCoeffFrac = own->JCoeffFrac.data();
    out->lineAtom = own->lineAtom.data();
    out->lineTrans = own->lineTrans.data();
    out->rhoCoefOff = own->rhoCoefOff.data();
    out->rhoFrac = own->rhoFrac.data();
    out->rhoI0 = own->rhoI0.data();
    std::lock_guard<std::mutex> lock(g_hprdMu);
    g_hprdReg[out->prdLaOfLa] = std::move(own);
    return 0;
}

void lwb200_free_hprd(LwB200HybridPrd* h)
{
    if (!h || !h->prdLaOfLa)
        return;
    std::lock_guard<std::mutex> lock(g_hprdMu);
    g_hprdReg.erase(h->prdLaOfLa);
    std::memset(h, 0, sizeof(*h));
}
} // extern "C"
