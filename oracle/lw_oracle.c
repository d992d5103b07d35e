/* lw_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see lw_oracle.h).
 *
 * Scalar CPU restatement of Lightweaver's hot path.  Each function cites the
 * reference lines whose arithmetic (including operation order) it follows, so
 * that agreement with the reference's scalar scheme is at rounding level.
 * All paths below are relative to /root/reference/Source/.
 */
#include "lw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

/* Constants.hpp:6-47 */
#define C_CLIGHT 2.99792458E+08
#define C_HPLANCK 6.6260755E-34
#define C_HC (C_HPLANCK * C_CLIGHT)
#define C_KBOLTZMANN 1.380658E-23
#define C_PI 3.14159265358979323846264338327950288
#define C_NM_TO_M 1.0E-09

static inline double sq(double x) { return x * x; }
static inline double cb(double x) { return x * x * x; }
static inline double dmin(double a, double b) { return b < a ? b : a; } /* std::min */
static inline double dmax(double a, double b) { return a < b ? b : a; } /* std::max */

/* ------------------------------------------------------------------------ */
/* LwInternal.hpp:90-110 */
static void w2(double dtau, double* w)
{
    if (dtau < 5.0E-4)
    {
        w[0] = dtau * (1.0 - 0.5 * dtau);
        w[1] = sq(dtau) * (0.5 - dtau * (1.0 / 3.0));
    }
    else if (dtau > 50.0)
    {
        w[1] = w[0] = 1.0;
    }
    else
    {
        double expdt = exp(-dtau);
        w[0] = 1.0 - expdt;
        w[1] = w[0] - dtau * expdt;
    }
}

/* LwMisc.hpp:29-46 */
static void planck_nu(int n, const double* T, double lambda, double* Bnu)
{
    const double hc_k = C_HC / (C_KBOLTZMANN * C_NM_TO_M);
    const double hc_kla = hc_k / lambda;
    const double twoh_c2 = (2.0 * C_HC) / cb(C_NM_TO_M);
    const double twohnu3_c2 = twoh_c2 / cb(lambda);
    for (int k = 0; k < n; ++k)
    {
        double x = hc_kla / T[k];
        Bnu[k] = (x <= 150.0) ? twohnu3_c2 / (exp(x) - 1.0) : 0.0;
    }
}

/* Bezier.hpp:58-65 (Steffen 1990) */
static double cent_deriv(double dsuw, double dsdw, double yuw, double y0, double ydw)
{
    const double S0 = (ydw - y0) / dsdw;
    const double Suw = (y0 - yuw) / dsuw;
    const double P0 = fabs((Suw * dsdw + S0 * dsuw) / (dsdw + dsuw));
    return (copysign(1.0, S0) + copysign(1.0, Suw)) * dmin(fabs(Suw), dmin(fabs(S0), 0.5 * P0));
}

/* Bezier.hpp:81-127 */
static void bezier3_coeffs(double dt, double* alpha, double* beta, double* gamma, double* delta,
                           double* edt)
{
    double dt2 = sq(dt);
    double dt3 = dt2 * dt;
    if (dt < 5e-2)
    {
        *edt = 1.0 - dt + 0.5 * dt2 - dt3 / 6.0;
        *alpha = 0.25 * dt - 0.2 * dt2 + dt3 / 12.0;
        *beta = 0.25 * dt - 0.05 * dt2 + dt3 / 120.0;
        *gamma = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
        *delta = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
    }
    else if (dt > 30.0)
    {
        *edt = 0.0;
        *alpha = 6.0 / dt3;
        *beta = (-6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
        *gamma = 3.0 * (2.0 * dt - 6.0) / dt3;
        *delta = 3.0 * (6.0 - 4.0 * dt + dt2) / dt3;
    }
    else
    {
        *edt = exp(-dt);
        *alpha = (6.0 - *edt * (6.0 + 6.0 * dt + 3 * dt2 + dt3)) / dt3;
        *beta = (6.0 * *edt - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
        *gamma = 3.0 * (2.0 * dt - 6.0 + *edt * (6.0 + 4.0 * dt + dt2)) / dt3;
        *delta = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * *edt * (3.0 + dt)) / dt3;
    }
}

/* FormalScalar.cpp:136-207 */
static void linear_sweep(int K, const double* h, const double* chi, const double* S, double zmu,
                         int toObs, double Istart, double* I, double* Psi)
{
    int dk = -1, ks = K - 1, ke = 0;
    if (!toObs) { dk = 1; ks = 0; ke = K - 1; }
    double dtau_uw = zmu * (chi[ks] + chi[ks + dk]) * fabs(h[ks] - h[ks + dk]);
    double rcp_uw = 1.0 / dtau_uw;
    double dS_uw = (S[ks] - S[ks + dk]) * rcp_uw;
    double I_upw = Istart;
    I[ks] = I_upw;
    if (Psi) Psi[ks] = 0.0;
    double w[2];
    for (int k = ks + dk; k != ke; k += dk)
    {
        w2(dtau_uw, w);
        double dtau_dw = zmu * (chi[k] + chi[k + dk]) * fabs(h[k] - h[k + dk]);
        double rcp_dw = 1.0 / dtau_dw;
        double dS_dw = (S[k] - S[k + dk]) * rcp_dw;
        I[k] = (1.0 - w[0]) * I_upw + w[0] * S[k] + w[1] * dS_uw;
        if (Psi) Psi[k] = w[0] - w[1] * rcp_uw;
        I_upw = I[k];
        dS_uw = dS_dw;
        dtau_uw = dtau_dw;
        rcp_uw = rcp_dw;
    }
    w2(dtau_uw, w);
    I[ke] = (1.0 - w[0]) * I_upw + w[0] * S[ke] + w[1] * dS_uw;
    if (Psi)
    {
        Psi[ke] = w[0] - w[1] * rcp_uw;
        for (int k = 0; k < K; ++k) Psi[k] /= chi[k];
    }
}

/* FormalScalar.cpp:209-325 */
static void bezier3_sweep(int K, const double* h, const double* chi, const double* S, double zmu,
                          int toObs, double Istart, double* I, double* Psi)
{
    int dk = -1, ks = K - 1, ke = 0;
    if (!toObs) { dk = 1; ks = 0; ke = K - 1; }
    double I_upw = Istart;
    I[ks] = I_upw;
    if (Psi) Psi[ks] = 0.0;

    int k = ks + dk;
    double ds_uw = fabs(h[k] - h[k - dk]) * zmu;
    double ds_dw = fabs(h[k + dk] - h[k]) * zmu;
    double dx_uw = (chi[k] - chi[k - dk]) / ds_uw;
    double dx_c = cent_deriv(ds_uw, ds_dw, chi[k - dk], chi[k], chi[k + dk]);
    double Cuw = chi[k - dk] + (ds_uw / 3.0) * dx_uw;
    double C0 = chi[k] - (ds_uw / 3.0) * dx_c;
    double dtau_uw = ds_uw * (chi[k] + chi[k - dk] + Cuw + C0) * 0.25;
    double dS_uw = (S[k] - S[k - dk]) / dtau_uw;
    double ds_dw2 = 0.0, dtau_dw = 0.0;
    double alpha, beta, gamma, delta, edt;

    for (; k != ke - dk; k += dk)
    {
        ds_dw2 = fabs(h[k + 2 * dk] - h[k + dk]) * zmu;
        double dx_dw = cent_deriv(ds_dw, ds_dw2, chi[k], chi[k + dk], chi[k + 2 * dk]);
        Cuw = chi[k] + (ds_dw / 3.0) * dx_c;
        C0 = chi[k + dk] - (ds_dw / 3.0) * dx_dw;
        dtau_dw = ds_dw * (chi[k] + chi[k + dk] + Cuw + C0) * 0.25;

        bezier3_coeffs(dtau_uw, &alpha, &beta, &gamma, &delta, &edt);
        double dS_c = cent_deriv(dtau_uw, dtau_dw, S[k - dk], S[k], S[k + dk]);
        Cuw = S[k - dk] + (dtau_uw / 3.0) * dS_uw;
        C0 = S[k] - (dtau_uw / 3.0) * dS_c;
        I[k] = I_upw * edt + alpha * S[k - dk] + beta * S[k] + gamma * Cuw + delta * C0;
        if (Psi) Psi[k] = beta + delta;

        I_upw = I[k];
        ds_uw = ds_dw;
        ds_dw = ds_dw2;
        dx_uw = dx_c;
        dx_c = dx_dw;
        dtau_uw = dtau_dw;
        dS_uw = dS_c;
    }
    /* second to last point: one-sided downwind derivative (:286-305) */
    k = ke - dk;
    ds_dw = fabs(h[k + dk] - h[k]) * zmu;
    double dx_dw = (chi[k + dk] - chi[k]) / ds_dw;
    Cuw = chi[k] + (ds_dw / 3.0) * dx_c;
    C0 = chi[k + dk] - (ds_dw / 3.0) * dx_dw;
    dtau_dw = ds_dw * (chi[k] + chi[k + dk] + Cuw + C0) * 0.25;
    bezier3_coeffs(dtau_uw, &alpha, &beta, &gamma, &delta, &edt);
    double dS_c = cent_deriv(dtau_uw, dtau_dw, S[k - dk], S[k], S[k + dk]);
    Cuw = S[k - dk] + dtau_uw / 3.0 * dS_uw;
    C0 = S[k] - dtau_uw / 3.0 * dS_c;
    I[k] = I_upw * edt + alpha * S[k - dk] + beta * S[k] + gamma * Cuw + delta * C0;
    if (Psi) Psi[k] = beta + delta;
    I_upw = I[k];

    /* piecewise linear on the end (:307-323) */
    k = ke;
    dtau_uw = 0.5 * zmu * (chi[k] + chi[k - dk]) * fabs(h[k] - h[k - dk]);
    dS_uw = (S[k] - S[k - dk]) / dtau_uw;
    double w[2];
    w2(dtau_uw, w);
    I[k] = (1.0 - w[0]) * I_upw + w[0] * S[k] - w[1] * dS_uw;
    if (Psi)
    {
        Psi[k] = w[0] - w[1] / dtau_uw;
        for (int kk = 0; kk < K; ++kk) Psi[kk] /= chi[kk];
    }
}

/* FormalScalar.cpp:327-363 */
static double besser_control_point(double hM, double hP, double yM, double yO, double yP)
{
    const double dM = (yO - yM) / hM;
    const double dP = (yP - yO) / hP;
    if (dM * dP <= 0.0)
        return yO;
    double yOp = (hM * dP + hP * dM) / (hM + hP);
    double cM = yO - 0.5 * hM * yOp;
    double cP = yO + 0.5 * hP * yOp;
    double minYMO = yM, maxYMO = yO, minYOP = yO, maxYOP = yP;
    if (dM < 0.0) { minYMO = yO; maxYMO = yM; minYOP = yP; maxYOP = yO; }
    if (cM < minYMO || cM > maxYMO)
        return yM;
    if (cP < minYOP || cP > maxYOP)
    {
        cP = yP;
        yOp = (cP - yO) / (0.5 * hP);
        cM = yO - 0.5 * hM * yOp;
    }
    return cM;
}

/* FormalScalar.cpp:373-393 */
static void besser_coeffs(double t, double* M, double* O, double* Cc, double* edt)
{
    if (t < 0.14)
    {
        *M = (t * (t * (t * (t * (t * (t * ((140.0 - 18.0 * t) * t - 945.0) + 5400.0) - 25200.0) + 90720.0) - 226800.0) + 302400.0)) / 907200.0;
        *O = (t * (t * (t * (t * (t * (t * ((10.0 - t) * t - 90.0) + 720.0) - 5040.0) + 30240.0) - 151200.0) + 604800.0)) / 1814400.0;
        *Cc = (t * (t * (t * (t * (t * (t * ((35.0 - 4.0 * t) * t - 270.0) + 1800.0) - 10080.0) + 45360.0) - 151200.0) + 302400.0)) / 907200.0;
        *edt = 1.0 - t + 0.5 * sq(t) - cb(t) / 6.0 + t * cb(t) / 24.0 - sq(t) * cb(t) / 120.0 + cb(t) * cb(t) / 720.0 - cb(t) * cb(t) * t / 5040.0;
    }
    else
    {
        double t2 = sq(t);
        *edt = exp(-t);
        *M = (2.0 - *edt * (t2 + 2.0 * t + 2.0)) / t2;
        *O = 1.0 - 2.0 * (*edt + t - 1.0) / t2;
        *Cc = 2.0 * (t - 2.0 + *edt * (t + 2.0)) / t2;
    }
}

/* FormalScalar.cpp:395-467 */
static void besser_sweep(int K, const double* h, const double* chi, const double* S, double zmu,
                         int toObs, double Istart, double* I, double* Psi)
{
    int dk = -1, ks = K - 1, ke = 0;
    if (!toObs) { dk = 1; ks = 0; ke = K - 1; }
    double I_upw = Istart;
    I[ks] = I_upw;
    if (Psi) Psi[ks] = 0.0;
    int k = ks + dk;
    for (; k != ke; k += dk)
    {
        double ds_uw = fabs(h[k] - h[k - dk]) * zmu;
        double ds_dw = fabs(h[k + dk] - h[k]) * zmu;
        double chi_uw = chi[k - dk], chiLocal = chi[k], chi_dw = chi[k + dk];
        double chiC = besser_control_point(ds_uw, ds_dw, chi_uw, chiLocal, chi_dw);
        double dtauUw = (1.0 / 3.0) * (chi_uw + chiC + chiLocal) * ds_uw;
        double dtauDw = 0.5 * (chiLocal + chi_dw) * ds_dw;
        double Suw = S[k - dk], SLocal = S[k], Sdw = S[k + dk];
        double SC = besser_control_point(dtauUw, dtauDw, Suw, SLocal, Sdw);
        double M, O, Cc, edt;
        besser_coeffs(dtauUw, &M, &O, &Cc, &edt);
        I[k] = I_upw * edt + M * Suw + O * SLocal + Cc * SC;
        if (Psi) Psi[k] = O + Cc;
        I_upw = I[k];
    }
    k = ke;
    double dtau_uw = 0.5 * zmu * (chi[k] + chi[k - dk]) * fabs(h[k] - h[k - dk]);
    double dS_uw = (S[k] - S[k - dk]) / dtau_uw;
    double w[2];
    w2(dtau_uw, w);
    I[k] = (1.0 - w[0]) * I_upw + w[0] * S[k] - w[1] * dS_uw;
    if (Psi)
    {
        Psi[k] = w[0] - w[1] / dtau_uw;
        for (int kk = 0; kk < K; ++kk) Psi[kk] /= chi[kk];
    }
}

/* Boundary wrappers: FormalScalar.cpp:471-533 (linear), :535-600 (bezier3),
 * :602-666 (besser). */
void lwo_solve_ray(int solver, int K, const double* h, const double* T, const double* chi,
                   const double* S, double muz, int toObs, double wavelength, int lowerBc,
                   int upperBc, double bcValue, double* I, double* Psi)
{
    double zmu = (solver == LWB200_FS_LINEAR) ? 0.5 / muz : 1.0 / muz;
    int dk = -1, ks = K - 1;
    if (!toObs) { dk = 1; ks = 0; }
    double dtau_uw;
    if (solver == LWB200_FS_LINEAR)
        dtau_uw = zmu * (chi[ks] + chi[ks + dk]) * fabs(h[ks] - h[ks + dk]);
    else
        dtau_uw = 0.5 * zmu * (chi[ks] + chi[ks + dk]) * fabs(h[ks] - h[ks + dk]);

    double Iupw = 0.0;
    if (toObs)
    {
        if (lowerBc == LWB200_BC_THERMALISED)
        {
            double Bnu[2];
            planck_nu(2, &T[K - 2], wavelength, Bnu);
            Iupw = Bnu[1] - (Bnu[0] - Bnu[1]) / dtau_uw;
        }
        else if (lowerBc == LWB200_BC_CALLABLE)
            Iupw = bcValue;
    }
    else
    {
        if (upperBc == LWB200_BC_THERMALISED)
        {
            double Bnu[2];
            planck_nu(2, &T[0], wavelength, Bnu);
            Iupw = Bnu[0] - (Bnu[1] - Bnu[0]) / dtau_uw;
        }
        else if (upperBc == LWB200_BC_CALLABLE)
            Iupw = bcValue;
    }
    if (solver == LWB200_FS_LINEAR)
        linear_sweep(K, h, chi, S, zmu, toObs, Iupw, I, Psi);
    else if (solver == LWB200_FS_BESSER)
        besser_sweep(K, h, chi, S, zmu, toObs, Iupw, I, Psi);
    else
        bezier3_sweep(K, h, chi, S, zmu, toObs, Iupw, I, Psi);
}

/* ------------------------------------------------------------------------ */
/* Transition::wlambda, LwTransition.hpp:71-81 */
static double wlambda(const LwB200Transition* t, int lt)
{
    int len = t->Nred - t->Nblue;
    const double* w = t->wavelength;
    if (lt == 0)
        return 0.5 * (w[1] - w[0]) * t->dopplerWidth;
    if (lt == len - 1)
        return 0.5 * (w[len - 1] - w[len - 2]) * t->dopplerWidth;
    return 0.5 * (w[lt + 1] - w[lt - 1]) * t->dopplerWidth;
}

typedef struct
{
    int K;
    double *chiTot, *etaTot, *S, *I, *Psi, *Ieff, *Uji, *Vij, *Vji, *JDag;
    double** gij;   /* [atom][Ntrans*K] */
    double** wla;
    double** aEta;  /* [atom][K] */
    double** aU;    /* [atom][Nlevel*K] */
    double** aChi;
    double* pool;
} Scratch;

static Scratch* scratch_new(const LwB200Problem* p)
{
    Scratch* s = (Scratch*)calloc(1, sizeof(Scratch));
    int K = p->Nspace;
    s->K = K;
    size_t n = 10 * (size_t)K;
    for (int a = 0; a < p->Natom; ++a)
        n += (size_t)K * (2 * p->atoms[a].Ntrans + 1 + 2 * p->atoms[a].Nlevel);
    s->pool = (double*)calloc(n, sizeof(double));
    double* q = s->pool;
    s->chiTot = q; q += K; s->etaTot = q; q += K; s->S = q; q += K; s->I = q; q += K;
    s->Psi = q; q += K; s->Ieff = q; q += K; s->Uji = q; q += K; s->Vij = q; q += K;
    s->Vji = q; q += K; s->JDag = q; q += K;
    s->gij = (double**)calloc(p->Natom, sizeof(double*));
    s->wla = (double**)calloc(p->Natom, sizeof(double*));
    s->aEta = (double**)calloc(p->Natom, sizeof(double*));
    s->aU = (double**)calloc(p->Natom, sizeof(double*));
    s->aChi = (double**)calloc(p->Natom, sizeof(double*));
    for (int a = 0; a < p->Natom; ++a)
    {
        const LwB200Atom* at = &p->atoms[a];
        s->gij[a] = q; q += (size_t)at->Ntrans * K;
        s->wla[a] = q; q += (size_t)at->Ntrans * K;
        s->aEta[a] = q; q += K;
        s->aU[a] = q; q += (size_t)at->Nlevel * K;
        s->aChi[a] = q; q += (size_t)at->Nlevel * K;
    }
    return s;
}

static void scratch_free(Scratch* s)
{
    free(s->gij); free(s->wla); free(s->aEta); free(s->aU); free(s->aChi);
    free(s->pool);
    free(s);
}

static inline int is_active(const LwB200Transition* t, int la) { return la >= t->Nblue && la < t->Nred; }

/* index of line (atom a, transition kr) in the hybrid-PRD tables (Transition::hPrdCoeffs set), or -1 */
static int hprd_line(const LwB200Problem* p, int a, int kr)
{
    const LwB200HybridPrd* h = p->hprd;
    if (!h)
        return -1;
    for (int q = 0; q < h->Nlines; ++q)
        if (h->lineAtom[q] == a && h->lineTrans[q] == kr)
            return q;
    return -1;
}

/* Atom::setup_wavelength, LwAtom.hpp:82-128 */
static void setup_wavelength(const LwB200Problem* p, int col, int a, int la, Scratch* s)
{
    const LwB200Atom* at = &p->atoms[a];
    const int K = p->Nspace;
    const double pi4_h = 4.0 * C_PI / C_HPLANCK;
    const double hc_4pi = 0.25 * C_HC / C_PI;
    const double pi4_hc = 1.0 / hc_4pi;
    const double hc_k = C_HC / (C_KBOLTZMANN * C_NM_TO_M);
    const double* nStar = at->nStar + (size_t)col * at->Nlevel * K;
    const double* T = p->temperature + (size_t)col * K;
    for (int kr = 0; kr < at->Ntrans; ++kr)
    {
        const LwB200Transition* t = &at->trans[kr];
        if (!is_active(t, la))
            continue;
        double* g = s->gij[a] + (size_t)kr * K;
        double* w = s->wla[a] + (size_t)kr * K;
        const int lt = la - t->Nblue;
        const int Nl = t->Nred - t->Nblue;
        const double wlam = wlambda(t, lt);
        if (t->type == LWB200_LINE)
        {
            const double* wphi = t->wphi + (size_t)col * K;
            for (int k = 0; k < K; ++k)
            {
                g[k] = t->Bji / t->Bij;
                w[k] = wlam * wphi[k] * pi4_hc;
            }
            if (t->rhoPrd && hprd_line(p, a, kr) < 0) /* (t.rhoPrd && !t.hPrdCoeffs), LwAtom.hpp:121 */
            {
                const double* rho = t->rhoPrd + ((size_t)col * Nl + lt) * K;
                for (int k = 0; k < K; ++k)
                    g[k] *= rho[k];
            }
        }
        else
        {
            const double hc_kl = hc_k / t->wavelength[lt];
            const double wlambda_lambda = wlam / t->wavelength[lt];
            for (int k = 0; k < K; ++k)
            {
                g[k] = nStar[(size_t)t->i * K + k] / nStar[(size_t)t->j * K + k] * exp(-hc_kl / T[k]);
                w[k] = wlambda_lambda * pi4_h;
            }
        }
    }
}

/* Transition::uv, LwTransition.hpp:93-144 */
static void uv(const LwB200Problem* p, int col, int a, int kr, int la, int mu, int toObs, Scratch* s)
{
    const LwB200Transition* t = &p->atoms[a].trans[kr];
    const int K = p->Nspace, M = p->Nrays;
    const int lt = la - t->Nblue;
    const int Nl = t->Nred - t->Nblue;
    const double* g = s->gij[a] + (size_t)kr * K;
    if (t->type == LWB200_LINE)
    {
        const double hc_4pi = 0.25 * C_HC / C_PI;
        const double hnu_4pi = hc_4pi * (t->lambda0 / t->wavelength[lt]);
        const double* ph = t->phi + ((((size_t)col * Nl + lt) * M + mu) * 2 + toObs) * K;
        for (int k = 0; k < K; ++k)
        {
            s->Vij[k] = hnu_4pi * t->Bij * ph[k];
            s->Vji[k] = g[k] * s->Vij[k];
        }
        /* the HPRD linear interpolation of rho to the rest-frame wavelength of this ray (:115-130) */
        const int hq = hprd_line(p, a, kr);
        if (hq >= 0)
        {
            const LwB200HybridPrd* h = p->hprd;
            const size_t o = (size_t)h->rhoCoefOff[hq] + ((((size_t)col * Nl + lt) * M + mu) * 2 + toObs) * K;
            const double* rho = t->rhoPrd + (size_t)col * Nl * K;
            for (int k = 0; k < K; ++k)
            {
                const double frac = h->rhoFrac[o + k];
                const int i0 = h->rhoI0[o + k], i1 = i0 + 1;
                const double r = (1.0 - frac) * rho[(size_t)i0 * K + k] + frac * rho[(size_t)i1 * K + k];
                s->Vji[k] *= r;
            }
        }
        for (int k = 0; k < K; ++k)
            s->Uji[k] = t->Aji / t->Bji * s->Vji[k];
    }
    else
    {
        const double twoHc = 2.0 * C_HC / cb(C_NM_TO_M);
        const double hcl = twoHc / cb(t->wavelength[lt]);
        const double al = t->alpha[lt];
        for (int k = 0; k < K; ++k)
        {
            s->Vij[k] = al;
            s->Vji[k] = g[k] * s->Vij[k];
            s->Uji[k] = hcl * s->Vji[k];
        }
    }
}

/* gather_opacity_emissivity_opt + chi_eta_aux_accum,
 * SimdFullIterationTemplates.hpp:59-167.  Per-atom aux arrays are zeroed per
 * call instead of using the reference's "first write stores" flags (:124-147);
 * the values read afterwards are identical. */
static void gather(const LwB200Problem* p, int col, int la, int mu, int toObs, Scratch* s)
{
    const int K = p->Nspace;
    for (int a = 0; a < p->Natom; ++a)
    {
        const LwB200Atom* at = &p->atoms[a];
        const double* n = at->n + (size_t)col * at->Nlevel * K;
        double* aChi = s->aChi[a];
        double* aU = s->aU[a];
        double* aEta = s->aEta[a];
        if (!at->detailedStatic)
        {
            memset(aChi, 0, sizeof(double) * at->Nlevel * K);
            memset(aU, 0, sizeof(double) * at->Nlevel * K);
            memset(aEta, 0, sizeof(double) * K);
        }
        for (int kr = 0; kr < at->Ntrans; ++kr)
        {
            const LwB200Transition* t = &at->trans[kr];
            if (!is_active(t, la))
                continue;
            uv(p, col, a, kr, la, mu, toObs, s);
            for (int k = 0; k < K; ++k)
            {
                double chi = n[(size_t)t->i * K + k] * s->Vij[k] - n[(size_t)t->j * K + k] * s->Vji[k];
                double eta = n[(size_t)t->j * K + k] * s->Uji[k];
                if (!at->detailedStatic)
                {
                    aChi[(size_t)t->i * K + k] += chi;
                    aChi[(size_t)t->j * K + k] -= chi;
                    aU[(size_t)t->j * K + k] += s->Uji[k];
                    aEta[k] += eta;
                }
                s->chiTot[k] += chi;
                s->etaTot[k] += eta;
            }
        }
    }
}

/* continua_only, SimdFullIterationTemplates.hpp:30-57 */
static int continua_only(const LwB200Problem* p, int la)
{
    for (int a = 0; a < p->Natom; ++a)
        for (int kr = 0; kr < p->atoms[a].Ntrans; ++kr)
        {
            const LwB200Transition* t = &p->atoms[a].trans[kr];
            if (is_active(t, la) && t->type != LWB200_CONTINUUM)
                return 0;
        }
    return 1;
}

static double bc_value(const LwB200Problem* p, int col, int la, int mu, int toObs)
{
    if (toObs && p->lowerBc == LWB200_BC_CALLABLE)
    {
        int idx = p->lowerBcIdx[mu * 2 + toObs];
        return p->lowerBcData[((size_t)col * p->Nspect + la) * p->NlowerBcMu + idx];
    }
    if (!toObs && p->upperBc == LWB200_BC_CALLABLE)
    {
        int idx = p->upperBcIdx[mu * 2 + toObs];
        return p->upperBcData[((size_t)col * p->Nspect + la) * p->NupperBcMu + idx];
    }
    return 0.0;
}

/* intensity_core_opt, SimdFullIterationTemplates.hpp:238-487.
 * updateRates/computeOperator both on for the Gamma iteration, both off for
 * formal_sol. */
static double intensity_core_mode(const LwB200Problem* p, int col, int la, Scratch* s, int fullIter,
                                  int lambdaIterate, int upOnly, int storeDepth, int prdOnly);

static double intensity_core(const LwB200Problem* p, int col, int la, Scratch* s, int fullIter,
                             int lambdaIterate, int upOnly, int storeDepth)
{
    return intensity_core_mode(p, col, la, s, fullIter, lambdaIterate, upOnly, storeDepth, 0);
}

/* prdOnly: the instantiation <UpdateRates = true, PrdRatesOnly = true, ComputeOperator = false>
 * with FsMode UpdateJ | UpdateRates | PrdOnly that formal_sol_prd_update_rates uses
 * (PrdTemplates.hpp:64-66): J and I as in the full iteration, rates only for transitions
 * with rhoPrd (:433-434, :455-456), no Gamma. */
static double intensity_core_mode(const LwB200Problem* p, int col, int la, Scratch* s, int fullIter,
                                  int lambdaIterate, int upOnly, int storeDepth, int prdOnly)
{
    const int K = p->Nspace, M = p->Nrays, L = p->Nspect;
    const double* h = p->height + (size_t)col * K;
    const double* T = p->temperature + (size_t)col * K;
    double* J = p->J + ((size_t)col * L + la) * K;
    const double* bgChi = p->chiBg + ((size_t)col * L + la) * K;
    const double* bgEta = p->etaBg + ((size_t)col * L + la) * K;
    const double* bgSca = p->scaBg + ((size_t)col * L + la) * K;
    const double wav = p->wavelength[la];

    memcpy(s->JDag, J, sizeof(double) * K);
    if (fullIter)
        memset(J, 0, sizeof(double) * K);

    for (int a = 0; a < p->Natom; ++a)
        setup_wavelength(p, col, a, la, s);

    const int contOnly = continua_only(p, la);
    const int toObsStart = upOnly ? 1 : 0;

    for (int mu = 0; mu < M; ++mu)
    {
        for (int toObs = toObsStart; toObs < 2; ++toObs)
        {
            if (!contOnly || (mu == 0 && toObs == toObsStart))
            {
                memcpy(s->chiTot, bgChi, sizeof(double) * K);
                memcpy(s->etaTot, bgEta, sizeof(double) * K);
                gather(p, col, la, mu, toObs, s);
                for (int k = 0; k < K; ++k)
                    s->S[k] = (s->etaTot[k] + bgSca[k] * s->JDag[k]) / s->chiTot[k];
                if (storeDepth)
                {
                    if (!contOnly)
                    {
                        size_t off = ((((size_t)col * L + la) * M + mu) * 2 + toObs) * K;
                        memcpy(p->depthChi + off, s->chiTot, sizeof(double) * K);
                        memcpy(p->depthEta + off, s->etaTot, sizeof(double) * K);
                    }
                    else
                    {
                        for (int m2 = 0; m2 < M; ++m2)
                            for (int d2 = 0; d2 < 2; ++d2)
                            {
                                size_t off = ((((size_t)col * L + la) * M + m2) * 2 + d2) * K;
                                memcpy(p->depthChi + off, s->chiTot, sizeof(double) * K);
                                memcpy(p->depthEta + off, s->etaTot, sizeof(double) * K);
                            }
                    }
                }
            }

            lwo_solve_ray(p->formalSolver, K, h, T, s->chiTot, s->S, p->muz[mu], toObs, wav,
                          p->lowerBc, p->upperBc, bc_value(p, col, la, mu, toObs), s->I,
                          (fullIter && !prdOnly) ? s->Psi : NULL);
            p->I[((size_t)col * L + la) * M + mu] = s->I[0];

            if (fullIter)
            {
                const double halfwmu = 0.5 * p->wmu[mu];
                for (int k = 0; k < K; ++k)
                    J[k] += halfwmu * s->I[k];
                /* rest-frame mean intensity of hybrid PRD (:397-408) */
                if (p->hprd && p->hprd->JRest)
                {
                    const LwB200HybridPrd* h = p->hprd;
                    const int hPrdLa = h->hPrdLaOfLa[(size_t)col * L + la];
                    if (hPrdLa >= 0)
                    {
                        double* JRest = h->JRest + (size_t)col * h->NprdLa * K;
                        const size_t row = ((((size_t)col * h->NhPrd + hPrdLa) * M + mu) * 2 + toObs) * K;
                        for (int k = 0; k < K; ++k)
                            for (int64_t e = h->JCoeffOff[row + k]; e < h->JCoeffOff[row + k + 1]; ++e)
                                JRest[(size_t)h->JCoeffIdx[e] * K + k] += 0.5 * p->wmu[mu] * h->JCoeffFrac[e] * s->I[k];
                    }
                }

                for (int a = 0; a < p->Natom; ++a)
                {
                    const LwB200Atom* at = &p->atoms[a];
                    const int N = at->Nlevel;
                    if (!at->detailedStatic && !prdOnly)
                    {
                        if (lambdaIterate)
                            memset(s->Psi, 0, sizeof(double) * K);
                        for (int k = 0; k < K; ++k)
                            s->Ieff[k] = s->I[k] - s->Psi[k] * s->aEta[a][k];
                    }
                    double* Gamma = (at->detailedStatic || prdOnly) ? NULL : at->Gamma + (size_t)col * N * N * K;
                    for (int kr = 0; kr < at->Ntrans; ++kr)
                    {
                        const LwB200Transition* t = &at->trans[kr];
                        if (!is_active(t, la))
                            continue;
                        if (prdOnly && !t->rhoPrd)
                            continue;
                        uv(p, col, a, kr, la, mu, toObs, s);
                        const double* wla = s->wla[a] + (size_t)kr * K;
                        double* Rij = t->Rij + (size_t)col * K;
                        double* Rji = t->Rji + (size_t)col * K;
                        /* compute_full_operator_rates, :206-234 */
                        for (int k = 0; k < K; ++k)
                        {
                            const double wlamu = wla[k] * halfwmu;
                            if (Gamma)
                            {
                                double integrand = (s->Uji[k] + s->Vji[k] * s->Ieff[k])
                                    - (s->Psi[k] * s->aChi[a][(size_t)t->i * K + k] * s->aU[a][(size_t)t->j * K + k]);
                                Gamma[((size_t)t->i * N + t->j) * K + k] += integrand * wlamu;
                                integrand = (s->Vij[k] * s->Ieff[k])
                                    - (s->Psi[k] * s->aChi[a][(size_t)t->j * K + k] * s->aU[a][(size_t)t->i * K + k]);
                                Gamma[((size_t)t->j * N + t->i) * K + k] += integrand * wlamu;
                            }
                            Rij[k] += s->I[k] * s->Vij[k] * wlamu;
                            Rji[k] += (s->Uji[k] + s->I[k] * s->Vji[k]) * wlamu;
                        }
                    }
                }
                if (storeDepth)
                {
                    size_t off = ((((size_t)col * L + la) * M + mu) * 2 + toObs) * K;
                    memcpy(p->depthI + off, s->I, sizeof(double) * K);
                }
            }
        }
    }

    double dJMax = 0.0;
    if (fullIter)
        for (int k = 0; k < K; ++k)
        {
            double dJ = fabs(1.0 - s->JDag[k] / J[k]);
            dJMax = dmax(dJ, dJMax);
        }
    return dJMax;
}

/* formal_sol_iteration_matrices_impl, Nthreads <= 1 branch (:597-637), then
 * finalise_Gamma (:491-508).  With a sub-range [laStart, laEnd) the Gamma
 * diagonal is still finalised, which is only meaningful for the full range. */
int lwo_fs_iter(const LwB200Problem* p, int col, unsigned flags, int laStart, int laEnd,
                double* dJMaxOut, int64_t* dJMaxIdx, int64_t* dJMaxIdxSerial)
{
    const int K = p->Nspace;
    if (laStart < 0) laStart = 0;
    if (laEnd < 0 || laEnd > p->Nspect) laEnd = p->Nspect;
    Scratch* s = scratch_new(p);
    for (int a = 0; a < p->Natom; ++a)
        for (int kr = 0; kr < p->atoms[a].Ntrans; ++kr)
        {
            memset(p->atoms[a].trans[kr].Rij + (size_t)col * K, 0, sizeof(double) * K);
            memset(p->atoms[a].trans[kr].Rji + (size_t)col * K, 0, sizeof(double) * K);
        }
    if (p->hprd && p->hprd->JRest) /* zero_Gamma_rates_JRest / :602-603 */
        memset(p->hprd->JRest + (size_t)col * p->hprd->NprdLa * K, 0, sizeof(double) * p->hprd->NprdLa * K);
    double dJMax = 0.0, dJSerial = 0.0;
    int64_t idx = 0, idxSerial = 0;
    const int storeDepth = (flags & LWB200_STORE_DEPTH) && p->depthChi && p->depthEta && p->depthI;
    for (int la = laStart; la < laEnd; ++la)
    {
        double dJ = intensity_core(p, col, la, s, 1, (flags & LWB200_LAMBDA_ITERATE) != 0, 0, storeDepth);
        /* threaded branch: td.dJ = max_idx(td.dJ, dJ, td.dJIdx, la)  (:688) */
        if (dJMax < dJ) { dJMax = dJ; idx = la; }
        /* serial branch: dJMax = max_idx(dJ, dJMax, maxIdx, la)  (:627; Constants.hpp:114-125) */
        if (dJ < dJSerial) { idxSerial = la; } else { dJSerial = dJ; }
    }
    for (int a = 0; a < p->Natom; ++a)
    {
        const LwB200Atom* at = &p->atoms[a];
        if (at->detailedStatic)
            continue;
        const int N = at->Nlevel;
        double* G = at->Gamma + (size_t)col * N * N * K;
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i)
            {
                G[((size_t)i * N + i) * K + k] = 0.0;
                double diag = 0.0;
                for (int j = 0; j < N; ++j)
                    diag += G[((size_t)j * N + i) * K + k];
                G[((size_t)i * N + i) * K + k] = -diag;
            }
    }
    scratch_free(s);
    if (dJMaxOut) *dJMaxOut = dJMax;
    if (dJMaxIdx) *dJMaxIdx = idx;
    if (dJMaxIdxSerial) *dJMaxIdxSerial = idxSerial;
    return 0;
}

/* formal_sol_impl, SimdFullIterationTemplates.hpp:721-737 */
int lwo_formal_sol(const LwB200Problem* p, int col, int upOnly)
{
    Scratch* s = scratch_new(p);
    for (int la = 0; la < p->Nspect; ++la)
        intensity_core(p, col, la, s, 0, 0, upOnly, 0);
    scratch_free(s);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* largest linear system of the restatement (the reference allocates its work arrays to size) */
#define LWO_MAX_SYSTEM 256

/* lu_decompose, LuSolve.cpp:8-70.  Returns 1 for the "Singular Matrix" throw. */
static int lu_decompose(int N, double* A, int* index)
{
    const double Tiny = 1e-20;
    double vv[LWO_MAX_SYSTEM];
    for (int i = 0; i < N; ++i)
    {
        double big = 0.0;
        for (int j = 0; j < N; ++j)
            big = dmax(big, fabs(A[i * N + j]));
        if (big == 0.0)
            return 1;
        vv[i] = 1.0 / big;
    }
    for (int j = 0; j < N; ++j)
    {
        for (int i = 0; i < j; ++i)
        {
            double sum = A[i * N + j];
            for (int k = 0; k < i; ++k)
                sum -= A[i * N + k] * A[k * N + j];
            A[i * N + j] = sum;
        }
        int iMax = 0;
        double big = 0.0;
        for (int i = j; i < N; ++i)
        {
            double sum = A[i * N + j];
            for (int k = 0; k < j; ++k)
                sum -= A[i * N + k] * A[k * N + j];
            A[i * N + j] = sum;
            double cand = vv[i] * fabs(sum);
            if (big < cand) { big = cand; iMax = i; } /* max_idx(big, cand, iMax, i) */
        }
        if (j != iMax)
        {
            for (int k = 0; k < N; ++k)
            {
                double temp = A[iMax * N + k];
                A[iMax * N + k] = A[j * N + k];
                A[j * N + k] = temp;
            }
            vv[iMax] = vv[j];
        }
        index[j] = iMax;
        if (A[j * N + j] == 0.0)
            A[j * N + j] = Tiny;
        double temp = 1.0 / A[j * N + j];
        for (int i = j + 1; i < N; ++i)
            A[i * N + j] *= temp;
    }
    return 0;
}

/* lu_backsub, LuSolve.cpp:72-101 */
static void lu_backsub(int N, const double* A, const int* index, double* b)
{
    int ii = -1;
    for (int i = 0; i < N; ++i)
    {
        int ip = index[i];
        double sum = b[ip];
        b[ip] = b[i];
        if (ii >= 0)
        {
            for (int j = ii; j < i; ++j)
                sum -= A[i * N + j] * b[j];
        }
        else if (sum != 0.0)
            ii = i;
        b[i] = sum;
    }
    for (int i = N - 1; i >= 0; --i)
    {
        double sum = b[i];
        for (int j = i + 1; j < N; ++j)
            sum -= A[i * N + j] * b[j];
        b[i] = sum / A[i * N + i];
    }
}

/* solve_lin_eq, LuSolve.cpp:103-133 */
int lwo_solve_lin_eq(int N, double* A, double* b, int improve)
{
    if (N > LWO_MAX_SYSTEM)
        return 2;
    double ACopySmall[64 * 64], bCopy[LWO_MAX_SYSTEM], residual[LWO_MAX_SYSTEM];
    double* ACopy = N <= 64 ? ACopySmall : (double*)malloc(sizeof(double) * N * N);
    int index[LWO_MAX_SYSTEM];
    if (improve)
    {
        memcpy(ACopy, A, sizeof(double) * N * N);
        memcpy(bCopy, b, sizeof(double) * N);
    }
    if (lu_decompose(N, A, index))
    {
        if (ACopy != ACopySmall)
            free(ACopy);
        return 1;
    }
    lu_backsub(N, A, index, b);
    if (improve)
    {
        memcpy(residual, bCopy, sizeof(double) * N);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                residual[i] -= ACopy[i * N + j] * b[j];
        lu_backsub(N, A, index, residual);
        for (int i = 0; i < N; ++i)
            b[i] += residual[i];
    }
    if (ACopy != ACopySmall)
        free(ACopy);
    return 0;
}

/* Ng acceleration of a sequence of solutions (Ng.hpp:16-163): the Ng(nOrder, nPeriod, nDelay, sol)
 * constructor on sols[0], then accelerate() + max_change() on sols[1..nIter] in turn.
 * out [nIter][len]: the solutions as accelerate() leaves them. */
int lwo_ng_run(int Norder, int Nperiod, int Ndelay, int len, int nIter, const double* sols, double* out,
               int* accelerated, double* dMax, int64_t* dMaxIdx)
{
    if (Norder < 0 || Norder > 8)
        return 2;
    const int R = Norder + 2;
    Ndelay = Ndelay > Nperiod + 2 ? Ndelay : Nperiod + 2;               /* :34 */
    /* the reference indexes previous(count - Norder - 2) at the first acceleration (:80-81): with
     * Ndelay < Norder + 2 that row index is negative (undefined behaviour there); not restated */
    if (Norder > 0 && Ndelay < Norder + 2)
        return 2;
    double* previous = (double*)calloc((size_t)R * len, sizeof(double));
    double* Delta = (double*)malloc((size_t)(Norder + 1) * len * sizeof(double));
    double* weight = (double*)malloc((size_t)len * sizeof(double));
    int count = 0, rc = 0;
    memcpy(previous, sols, sizeof(double) * len);                       /* :37-40 */
    count = 1;
    for (int it = 0; it < nIter && !rc; ++it)
    {
        double* sol = out + (size_t)it * len;
        memcpy(sol, sols + (size_t)(it + 1) * len, sizeof(double) * len);
        /* accelerate, :52-114 */
        memcpy(previous + (size_t)(count % R) * len, sol, sizeof(double) * len);
        count += 1;
        accelerated[it] = 0;
        if (Norder > 0 && count >= Ndelay && ((count - Ndelay) % Nperiod) == 0)
        {
            for (int i = 0; i <= Norder; ++i)
            {
                const double* a = previous + (size_t)((count - i - 1) % R) * len;
                const double* b = previous + (size_t)((count - i - 2) % R) * len;
                for (int k = 0; k < len; ++k)
                    Delta[(size_t)i * len + k] = a[k] - b[k];
            }
            for (int k = 0; k < len; ++k)
                weight[k] = 1.0 / fabs(sol[k]);
            double A[64] = {0.0}, b[8] = {0.0};
            for (int j = 0; j < Norder; ++j)
            {
                for (int k = 0; k < len; ++k)
                    b[j] += weight[k] * Delta[k] * (Delta[k] - Delta[(size_t)(j + 1) * len + k]);
                for (int i = 0; i < Norder; ++i)
                    for (int k = 0; k < len; ++k)
                        A[i * Norder + j] += weight[k] * (Delta[(size_t)(j + 1) * len + k] - Delta[k])
                                             * (Delta[(size_t)(i + 1) * len + k] - Delta[k]);
            }
            if (lwo_solve_lin_eq(Norder, A, b, 1))
            {
                rc = 1;
                break;
            }
            double* p0 = previous + (size_t)((count - 1) % R) * len;
            for (int i = 0; i < Norder; ++i)
            {
                const double* pi = previous + (size_t)((count - i - 2) % R) * len;
                for (int k = 0; k < len; ++k)
                    sol[k] += b[i] * (pi[k] - p0[k]);
            }
            memcpy(p0, sol, sizeof(double) * len);
            accelerated[it] = 1;
        }
        /* max_change, :138-156 */
        dMax[it] = 0.0;
        dMaxIdx[it] = 0;
        if (count >= 2)
        {
            const double* old = previous + (size_t)((count - 2) % R) * len;
            const double* cur = previous + (size_t)((count - 1) % R) * len;
            for (int k = 0; k < len; ++k)
                if (cur[k] != 0.0)
                {
                    const double change = fabs((cur[k] - old[k]) / cur[k]);
                    if (dMax[it] < change)
                    {
                        dMax[it] = change;
                        dMaxIdx[it] = k;
                    }
                }
        }
    }
    free(previous);
    free(Delta);
    free(weight);
    return rc;
}

/* stat_eq_impl, UpdatePopulations.cpp:7-47 */
int lwo_stat_eq(const LwB200Problem* p, int col, int atom, int kStart, int kEnd, int* nSingular)
{
    const int K = p->Nspace;
    if (kStart < 0 && kEnd < 0) { kStart = 0; kEnd = K; }
    int singular = 0;
    for (int a = 0; a < p->Natom; ++a)
    {
        if (atom >= 0 && a != atom)
            continue;
        const LwB200Atom* at = &p->atoms[a];
        if (at->detailedStatic)
            continue;
        const int N = at->Nlevel;
        if (N > 64)
            return 2;
        double* n = at->n + (size_t)col * N * K;
        const double* G = at->Gamma + (size_t)col * N * N * K;
        const double* nTotal = at->nTotal + (size_t)col * K;
        double nk[64], Gam[64 * 64];
        for (int k = kStart; k < kEnd; ++k)
        {
            for (int i = 0; i < N; ++i)
            {
                nk[i] = n[(size_t)i * K + k];
                for (int j = 0; j < N; ++j)
                    Gam[i * N + j] = G[((size_t)i * N + j) * K + k];
            }
            int iEliminate = 0;
            double nMax = 0.0;
            for (int i = 0; i < N; ++i)
                if (nMax < nk[i]) { nMax = nk[i]; iEliminate = i; } /* max_idx(nMax, nk(i), ...) */
            for (int i = 0; i < N; ++i)
            {
                Gam[iEliminate * N + i] = 1.0;
                nk[i] = 0.0;
            }
            nk[iEliminate] = nTotal[k];
            if (lwo_solve_lin_eq(N, Gam, nk, 1))
            {
                singular += 1;
                continue;
            }
            for (int i = 0; i < N; ++i)
                n[(size_t)i * K + k] = nk[i];
        }
    }
    if (nSingular) *nSingular = singular;
    return singular ? 1 : 0;
}

typedef struct
{
    const LwB200Problem* p;
    int colEnd, withStatEq;
    unsigned flags;
    atomic_int next;
    atomic_int rc;
} ColumnJob;

static void* column_worker(void* arg)
{
    ColumnJob* job = (ColumnJob*)arg;
    for (;;)
    {
        int c = atomic_fetch_add(&job->next, 1);
        if (c >= job->colEnd)
            break;
        int rc = lwo_fs_iter(job->p, c, job->flags, 0, -1, NULL, NULL, NULL);
        if (job->withStatEq)
        {
            int ns = 0;
            rc |= lwo_stat_eq(job->p, c, -1, -1, -1, &ns);
        }
        if (rc)
            atomic_fetch_or(&job->rc, rc);
    }
    return NULL;
}

/* Columns are independent: one worker per host thread pulls columns off a
 * shared counter (the user-level ProcessPool/MPI pattern the reference's docs
 * describe for 1.5D, docs/index.rst:37-40). */
int lwo_fs_iter_columns(const LwB200Problem* p, int col0, int ncol, unsigned flags, int withStatEq,
                        int nthreads)
{
    ColumnJob job;
    job.p = p;
    job.colEnd = col0 + ncol;
    job.withStatEq = withStatEq;
    job.flags = flags;
    atomic_init(&job.next, col0);
    atomic_init(&job.rc, 0);
    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > 256)
        nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t)
        pthread_create(&th[t], NULL, column_worker, &job);
    column_worker(&job);
    for (int t = 1; t < nthreads; ++t)
        pthread_join(th[t], NULL);
    return atomic_load(&job.rc);
}

/* time_dependent_update_impl, UpdatePopulations.cpp:120-151.  nOld: [Ncol][Nlevel][K] of the atom. */
int lwo_time_dep_update(const LwB200Problem* p, int col, int atom, const double* nOld, double dt)
{
    const int K = p->Nspace;
    const LwB200Atom* at = &p->atoms[atom];
    const int N = at->Nlevel;
    double* nk = (double*)malloc(sizeof(double) * N);
    double* G = (double*)malloc(sizeof(double) * N * N);
    const double* Gamma = at->Gamma + (size_t)col * N * N * K;
    double* n = at->n + (size_t)col * N * K;
    int rc = 0;
    for (int k = 0; k < K; ++k)
    {
        for (int i = 0; i < N; ++i)
        {
            nk[i] = nOld[((size_t)col * N + i) * K + k];
            for (int j = 0; j < N; ++j)
                G[i * N + j] = -Gamma[((size_t)i * N + j) * K + k] * dt;
            G[i * N + i] = 1.0 - Gamma[((size_t)i * N + i) * K + k] * dt;
        }
        if (lwo_solve_lin_eq(N, G, nk, 1))
        {
            rc = 1;
            break;
        }
        for (int i = 0; i < N; ++i)
            n[(size_t)i * K + k] = nk[i];
    }
    free(nk);
    free(G);
    return rc;
}

/* nr_post_update_impl with F / Ftd, UpdatePopulations.cpp:230-394, on column `col`. */
int lwo_nr_post_update(const LwB200Problem* p, int col, const LwB200NrUpdate* u)
{
    const int K = p->Nspace;
    const int fdCollisionRates = u->dC != NULL;
    const int timeDep = u->timeDependent != 0;
    int Nlevel = 0;
    for (int a = 0; a < u->Natom; ++a)
        Nlevel += p->atoms[u->atomIdx[a]].Nlevel;
    const int Neqn = Nlevel + 1;
    double* dF = (double*)malloc(sizeof(double) * Neqn * Neqn);
    double* Fg = (double*)malloc(sizeof(double) * Neqn);
    double* ne = p->ne + (size_t)col * K;
    const double* bgNe = u->backgroundNe + (size_t)col * K;
    const double theta = 1.0;
    int rc = 0;
    if (!p->ne)
        rc = 1;
    for (int k = 0; k < K && !rc; ++k)
    {
        memset(dF, 0, sizeof(double) * Neqn * Neqn);
        /* F (:230-258) / Ftd (:260-290) */
        memset(Fg, 0, sizeof(double) * Neqn);
        Fg[Neqn - 1] = ne[k];
        int start = 0;
        for (int a = 0; a < u->Natom; ++a)
        {
            const LwB200Atom* at = &p->atoms[u->atomIdx[a]];
            const int N = at->Nlevel;
            const double* G = at->Gamma + (size_t)col * N * N * K;
            const double* n = at->n + (size_t)col * N * K;
            for (int l = 0; l < N; ++l)
            {
                Fg[start + l] = 0.0;
                if (timeDep)
                {
                    for (int ll = 0; ll < N; ++ll)
                        Fg[start + l] += G[((size_t)l * N + ll) * K + k] * n[(size_t)ll * K + k];
                    Fg[start + l] *= theta * u->dt;
                    Fg[start + l] -= n[(size_t)l * K + k] - u->nPrev[a][((size_t)col * N + l) * K + k];
                }
                else
                {
                    for (int ll = 0; ll < N; ++ll)
                        Fg[start + l] -= G[((size_t)l * N + ll) * K + k] * n[(size_t)ll * K + k];
                }
            }
            double nTotCur = 0.0;
            for (int ll = 0; ll < N; ++ll)
                nTotCur += n[(size_t)ll * K + k];
            Fg[start + N - 1] = nTotCur - at->nTotal[(size_t)col * K + k];
            double eleContrib = 0.0;
            for (int ll = 0; ll < N; ++ll)
                eleContrib += at->stages[ll] * n[(size_t)ll * K + k];
            Fg[Neqn - 1] -= eleContrib;
            start += N;
        }
        Fg[Neqn - 1] -= bgNe[k];

        start = 0;
        for (int a = 0; a < u->Natom; ++a)
        {
            const LwB200Atom* at = &p->atoms[u->atomIdx[a]];
            const int N = at->Nlevel;
            const double* G = at->Gamma + (size_t)col * N * N * K;
            const double* Cm = at->C + (size_t)col * N * N * K;
            const double* n = at->n + (size_t)col * N * K;
            for (int l = 0; l < N; ++l)
                for (int ll = 0; ll < N; ++ll)
                    dF[(start + l) * Neqn + start + ll] = -G[((size_t)l * N + ll) * K + k];
            if (timeDep)
            {
                for (int l = 0; l < N; ++l)
                    for (int ll = 0; ll < N; ++ll)
                        dF[(start + l) * Neqn + start + ll] *= -theta * u->dt;
                for (int l = 0; l < N; ++l)
                    dF[(start + l) * Neqn + start + l] -= 1.0;
            }
            for (int tIdx = 0; tIdx < at->Ntrans; ++tIdx)
            {
                const LwB200Transition* t = &at->trans[tIdx];
                if (t->type == LWB200_CONTINUUM)
                {
                    double preconRji = G[((size_t)t->i * N + t->j) * K + k] - u->crswVal * Cm[((size_t)t->i * N + t->j) * K + k];
                    double entry = -(preconRji / ne[k]) * n[(size_t)t->j * K + k];
                    if (timeDep)
                        entry *= -theta * u->dt;
                    dF[(start + t->i) * Neqn + Neqn - 1] += entry;
                }
            }
            if (fdCollisionRates)
            {
                const double* dC = u->dC[a] + (size_t)col * N * N * K;
                for (int i = 0; i < N; ++i)
                {
                    double entry = 0.0;
                    for (int ll = 0; ll < N; ++ll)
                        entry -= dC[((size_t)i * N + ll) * K + k] * n[(size_t)ll * K + k];
                    if (timeDep)
                        entry *= -theta * u->dt;
                    dF[(start + i) * Neqn + Neqn - 1] += entry;
                }
            }
            for (int q = 0; q < Neqn; ++q)
                dF[(start + N - 1) * Neqn + q] = 0.0;
            for (int ll = 0; ll < N; ++ll)
            {
                dF[(start + N - 1) * Neqn + start + ll] = 1.0;
                dF[(Neqn - 1) * Neqn + start + ll] = -at->stages[ll];
            }
            start += N;
        }
        dF[(Neqn - 1) * Neqn + Neqn - 1] = 1.0;
        for (int i = 0; i < Neqn; ++i)
            Fg[i] *= -1.0;
        if (lwo_solve_lin_eq(Neqn, dF, Fg, 1))
        {
            rc = 1;
            break;
        }
        start = 0;
        for (int a = 0; a < u->Natom; ++a)
        {
            const LwB200Atom* at = &p->atoms[u->atomIdx[a]];
            double* n = at->n + (size_t)col * at->Nlevel * K;
            for (int ll = 0; ll < at->Nlevel; ++ll)
                n[(size_t)ll * K + k] += Fg[start + ll];
            start += at->Nlevel;
        }
        ne[k] += Fg[Neqn - 1];
    }
    free(dF);
    free(Fg);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Angle-averaged PRD: redistribute_prd_lines (Prd.cpp:648-658 ->
 * redistribute_prd_lines_template, PrdTemplates.hpp:164-351, Nthreads <= 1 branch). */

/* Prd.cpp:33-36 */
#define PRD_QWING 4.0
#define PRD_QCORE 2.0
#define PRD_QSPREAD 5.0
#define PRD_DQ 0.15

/* Prd.cpp:46-49 */
static double G_zero(double x) { return 1.0 / (fabs(x) + sqrt(sq(x) + 1.273239545)); }

/* Gouttebroze's GII, Prd.cpp:51-124 (waveratio = 1) */
static double GII(double aDamp, double qEmit, double qAbs)
{
    const double waveratio = 1.0;
    if (qEmit < 0.0)
    {
        qEmit = -qEmit;
        qAbs = -qAbs;
    }
    double giiCore = 0.0, coreFactor = 0.0;
    if (qEmit < PRD_QWING)
    {
        if ((qAbs < -PRD_QWING) || (qAbs > qEmit + waveratio * PRD_QSPREAD))
            return 0.0;
        if (fabs(qAbs) <= qEmit)
            giiCore = G_zero(qEmit);
        else
            giiCore = exp(sq(qEmit) - sq(qAbs)) * G_zero(qAbs);
        if (qEmit >= PRD_QCORE && qEmit <= PRD_QWING)
        {
            double phiCore = exp(-sq(qEmit));
            double phiWing = aDamp / (sqrt(C_PI) * (sq(aDamp) + sq(qEmit)));
            coreFactor = phiCore / (phiCore + phiWing);
        }
        else
            return giiCore;
    }
    double gii = 0.0;
    if (qEmit >= PRD_QCORE)
    {
        double aqEmit = waveratio * qEmit;
        if ((qEmit >= PRD_QWING) && (fabs(qAbs - aqEmit) > waveratio * PRD_QSPREAD))
            return 0.0;
        double uMin = fabs((qAbs - aqEmit) / (1.0 + waveratio));
        double giiWing = (1.0 + waveratio) * (1.0 - 2.0 * uMin * G_zero(uMin)) * exp(-sq(uMin))
            / (2.0 * waveratio * sqrt(C_PI));
        double ratio = qAbs / qEmit;
        giiWing *= (2.75 - (2.5 - 0.75 * ratio) * ratio);
        gii = coreFactor * giiCore + (1.0 - coreFactor) * giiWing;
    }
    return gii;
}

/* scattering_int_range, Prd.cpp:231-259 */
static void scattering_int_range(double qEmit, double* q0, double* qN)
{
    if (fabs(qEmit) < PRD_QCORE)
    {
        *q0 = -PRD_QWING;
        *qN = PRD_QWING;
    }
    else if (fabs(qEmit) < PRD_QWING)
    {
        if (qEmit > 0.0)
        {
            *q0 = -PRD_QWING;
            *qN = qEmit + PRD_QSPREAD;
        }
        else
        {
            *q0 = qEmit - PRD_QSPREAD;
            *qN = PRD_QWING;
        }
    }
    else
    {
        *q0 = qEmit - PRD_QSPREAD;
        *qN = qEmit + PRD_QSPREAD;
    }
}

/* optimised_fine_linear_fixed_spacing, Prd.cpp:180-228 */
static void fine_linear_fixed_spacing(int Ntable, const double* xTable, const double* yTable, double xStart,
                                      double xStep, int N, double* y)
{
    if (N < 1)
        return;
    int iter; /* index of the first table entry > x (std::upper_bound) */
    double x = xStart;
    if (x <= xTable[0])
        iter = 0;
    else if (x >= xTable[Ntable - 1])
        iter = Ntable - 1;
    else
    {
        iter = 0;
        while (iter < Ntable && !(x < xTable[iter]))
            ++iter;
    }
    for (int i = 0; i < N; ++i)
    {
        x = xStart + i * xStep;
        while (iter < Ntable && xTable[iter] <= x)
            ++iter;
        if (iter == Ntable)
        {
            y[i] = yTable[Ntable - 1];
            continue;
        }
        else if (iter == 0)
        {
            y[i] = yTable[0];
            continue;
        }
        double xp = xTable[iter - 1], xn = xTable[iter];
        double t = (x - xp) / (xn - xp);
        y[i] = (1.0 - t) * yTable[iter - 1] + t * yTable[iter];
    }
}

#define PRD_MAX_FINE 128 /* max_fine_grid_size() = 87, Prd.cpp:126-129 */

/* total_depop_elastic_scattering_rate (Prd.cpp:9-30) + prd_scatter / scattering_int
 * (Prd.cpp:468-645) for one PRD line of column col.  gII is recomputed on every call: the
 * reference caches it per (depth, wavelength) until aDamp / vBroad change, the values are the same. */
static int prd_scatter_line(const LwB200Problem* p, int col, int a, int kr)
{
    const int K = p->Nspace, L = p->Nspect;
    const LwB200Atom* at = &p->atoms[a];
    const LwB200Transition* t = &at->trans[kr];
    const int N = at->Nlevel, Nl = t->Nred - t->Nblue;
    if (!t->Qelast || !at->C || !t->aDamp || !at->vBroad)
        return 1;
    double* rho = t->rhoPrd + (size_t)col * Nl * K;
    double* Jk = (double*)malloc(sizeof(double) * Nl);
    double* qWave = (double*)malloc(sizeof(double) * Nl);
    double JFine[PRD_MAX_FINE], gII[PRD_MAX_FINE];
    for (size_t q = 0; q < (size_t)Nl * K; ++q)
        rho[q] = 1.0;
    for (int k = 0; k < K; ++k)
    {
        double PjQj = t->Qelast[(size_t)col * K + k];
        for (int i = 0; i < N; ++i)
            PjQj += at->C[(((size_t)col * N + i) * N + t->j) * K + k];
        for (int kr2 = 0; kr2 < at->Ntrans; ++kr2)
        {
            const LwB200Transition* t2 = &at->trans[kr2];
            if (t2->j == t->j)
                PjQj += t2->Rji[(size_t)col * K + k];
            if (t2->i == t->j)
                PjQj += t2->Rij[(size_t)col * K + k];
        }
        const double* n = at->n + (size_t)col * N * K;
        double gammaPrefactor = n[(size_t)t->i * K + k] / n[(size_t)t->j * K + k] * t->Bij / PjQj;
        double Jbar = t->Rij[(size_t)col * K + k] / t->Bij;
        for (int la = 0; la < Nl; ++la)
        {
            /* local mean intensity, in the rest frame if using HPRD (Prd.cpp:484-499) */
            if (p->hprd && p->hprd->JRest)
                Jk[la] = p->hprd->JRest[((size_t)col * p->hprd->NprdLa + p->hprd->prdLaOfLa[la + t->Nblue]) * K + k];
            else
                Jk[la] = p->J[((size_t)col * L + la + t->Nblue) * K + k];
            qWave[la] = (t->wavelength[la] - t->lambda0) * C_CLIGHT / (t->lambda0 * at->vBroad[(size_t)col * K + k]);
        }
        const double aDamp = t->aDamp[(size_t)col * K + k];
        for (int la = 0; la < Nl; ++la)
        {
            const double qEmit = qWave[la];
            double q0, qN;
            scattering_int_range(qEmit, &q0, &qN);
            const int Np = (int)((double)(qN - q0) / PRD_DQ) + 1;
            fine_linear_fixed_spacing(Nl, qWave, Jk, q0, PRD_DQ, Np, JFine);
            double qPrime = q0;
            gII[0] = GII(aDamp, qEmit, qPrime) * 5.0 / 12.0 * PRD_DQ;
            qPrime += PRD_DQ;
            gII[1] = GII(aDamp, qEmit, qPrime) * 13.0 / 12.0 * PRD_DQ;
            for (int laFine = 2; laFine < Np - 2; ++laFine)
            {
                qPrime += PRD_DQ;
                gII[laFine] = GII(aDamp, qEmit, qPrime) * PRD_DQ;
            }
            qPrime += PRD_DQ;
            gII[Np - 2] = GII(aDamp, qEmit, qPrime) * 13.0 / 12.0 * PRD_DQ;
            qPrime += PRD_DQ;
            gII[Np - 1] = GII(aDamp, qEmit, qPrime) * 5.0 / 12.0 * PRD_DQ;
            double gNorm = 0.0, scatInt = 0.0;
            for (int laF = 0; laF < Np; ++laF)
            {
                gNorm += gII[laF];
                scatInt += JFine[laF] * gII[laF];
            }
            rho[(size_t)la * K + k] += gammaPrefactor * (scatInt / gNorm - Jbar);
        }
    }
    free(Jk);
    free(qWave);
    return 0;
}

int lwo_redistribute_prd(const LwB200Problem* p, int col, int maxIter, double tol, int includeDetailed,
                         int* nIterOut, double* dRho, int* dRhoIdx, double* dJPrdMax, int64_t* dJPrdMaxIdx)
{
    const int K = p->Nspace, L = p->Nspect;
    /* the PRD lines, active atoms first then (optionally) detailed ones (PrdTemplates.hpp:186-211) */
    int nLines = 0;
    int (*lines)[2] = (int (*)[2])malloc(sizeof(int[2]) * 256);
    for (int pass = 0; pass < (includeDetailed ? 2 : 1); ++pass)
        for (int a = 0; a < p->Natom; ++a)
        {
            if ((p->atoms[a].detailedStatic != 0) != (pass == 1))
                continue;
            for (int kr = 0; kr < p->atoms[a].Ntrans; ++kr)
                if (p->atoms[a].trans[kr].rhoPrd && nLines < 256)
                {
                    lines[nLines][0] = a;
                    lines[nLines][1] = kr;
                    ++nLines;
                }
        }
    if (nIterOut) *nIterOut = 0;
    if (nLines == 0)
    {
        free(lines);
        return 0;
    }
    /* Ng(0, 0, 0, rho): change tracking only (Ng.hpp:30-40, :52-62, :137-155) */
    double** prev = (double**)malloc(sizeof(double*) * nLines);
    for (int q = 0; q < nLines; ++q)
    {
        const LwB200Transition* t = &p->atoms[lines[q][0]].trans[lines[q][1]];
        const size_t n = (size_t)(t->Nred - t->Nblue) * K;
        prev[q] = (double*)malloc(sizeof(double) * n);
        memcpy(prev[q], t->rhoPrd + (size_t)col * n, sizeof(double) * n);
    }
    /* wavelengths touched by a PRD line (:225-240) */
    char* prdLa = (char*)calloc(L, 1);
    if (p->hprd && p->hprd->NhPrd > 0)
    {
        /* idxsForFs = spect.hPrdIdxs (:233-234) */
        for (int la = 0; la < L; ++la)
            prdLa[la] = p->hprd->hPrdLaOfLa[(size_t)col * L + la] >= 0;
    }
    else
        for (int q = 0; q < nLines; ++q)
        {
            const LwB200Transition* t = &p->atoms[lines[q][0]].trans[lines[q][1]];
            for (int la = t->Nblue; la < t->Nred; ++la)
                prdLa[la] = 1;
        }
    Scratch* s = scratch_new(p);
    int iter = 0, rc = 0;
    while (iter < maxIter)
    {
        ++iter;
        double dRhoMax = 0.0;
        for (int q = 0; q < nLines; ++q)
        {
            const LwB200Transition* t = &p->atoms[lines[q][0]].trans[lines[q][1]];
            const int Nl = t->Nred - t->Nblue;
            if (prd_scatter_line(p, col, lines[q][0], lines[q][1]))
            {
                rc = 1;
                goto done;
            }
            const double* cur = t->rhoPrd + (size_t)col * Nl * K;
            double dMax = 0.0;
            int maxIdx = 0;
            for (size_t e = 0; e < (size_t)Nl * K; ++e)
                if (cur[e] != 0.0)
                {
                    double change = fabs((cur[e] - prev[q][e]) / cur[e]);
                    if (dMax < change)
                    {
                        dMax = change;
                        maxIdx = (int)e;
                    }
                }
            memcpy(prev[q], cur, sizeof(double) * (size_t)Nl * K);
            dRhoMax = dmax(dRhoMax, dMax);
            if (dRho) dRho[(size_t)(iter - 1) * nLines + q] = dMax;
            if (dRhoIdx) dRhoIdx[(size_t)(iter - 1) * nLines + q] = maxIdx % Nl;
        }
        /* formal_sol_prd_update_rates (PrdTemplates.hpp:18-76) */
        for (int q = 0; q < nLines; ++q)
        {
            const LwB200Transition* t = &p->atoms[lines[q][0]].trans[lines[q][1]];
            memset(t->Rij + (size_t)col * K, 0, sizeof(double) * K);
            memset(t->Rji + (size_t)col * K, 0, sizeof(double) * K);
        }
        if (p->hprd && p->hprd->JRest) /* PrdTemplates.hpp:57-58 */
            memset(p->hprd->JRest + (size_t)col * p->hprd->NprdLa * K, 0, sizeof(double) * p->hprd->NprdLa * K);
        double dJMax = 0.0;
        int64_t dJIdx = 0;
        for (int la = 0; la < L; ++la)
        {
            if (!prdLa[la])
                continue;
            double dJ = intensity_core_mode(p, col, la, s, 1, 0, 0, 0, 1);
            if (dJMax < dJ)
            {
                dJMax = dJ;
                dJIdx = la;
            }
        }
        if (dJPrdMax) dJPrdMax[iter - 1] = dJMax;
        if (dJPrdMaxIdx) dJPrdMaxIdx[iter - 1] = dJIdx;
        if (dRhoMax < tol)
            break;
    }
done:
    if (nIterOut) *nIterOut = iter;
    scratch_free(s);
    free(prdLa);
    for (int q = 0; q < nLines; ++q)
        free(prev[q]);
    free(prev);
    free(lines);
    return rc;
}


/* ------------------------------------------------------------------------ */
/* configure_hprd_coeffs, Prd.cpp:697-946, column by column (every column of a stack is its own
 * Context in the reference), flattened into LwB200HybridPrd.  The arrays are malloc'ed; release them
 * with lwo_free_hprd. */
typedef struct { int32_t idx; double frac; } JCoef;
typedef struct { JCoef* v; int64_t n, cap; } JVec;

static void jvec_push(JVec* q, int32_t idx, double frac)
{
    if (q->n == q->cap)
    {
        q->cap = q->cap ? 2 * q->cap : 4;
        q->v = (JCoef*)realloc(q->v, sizeof(JCoef) * q->cap);
    }
    q->v[q->n].idx = idx;
    q->v[q->n].frac = frac;
    q->n += 1;
}

static const double* upper_bound_d(const double* first, const double* last, double value)
{
    /* std::upper_bound: first element greater than value */
    while (first < last)
    {
        const double* mid = first + (last - first) / 2;
        if (value < *mid)
            last = mid;
        else
            first = mid + 1;
    }
    return first;
}

int lwo_configure_hprd(const LwB200Problem* p, int includeDetailed, LwB200HybridPrd* out)
{
    const int K = p->Nspace, M = p->Nrays, Nspect = p->Nspect, Ncol = p->Ncol;
    const double sign[2] = {-1.0, 1.0};
    memset(out, 0, sizeof(*out));
    if (!p->vlosMu)
        return 1;
    /* prdLines: active atoms first, then (optionally) detailed ones (:711-734) */
    int nLines = 0;
    int32_t* lineAtom = (int32_t*)malloc(sizeof(int32_t) * 256);
    int32_t* lineTrans = (int32_t*)malloc(sizeof(int32_t) * 256);
    for (int pass = 0; pass < (includeDetailed ? 2 : 1); ++pass)
        for (int a = 0; a < p->Natom; ++a)
        {
            if ((p->atoms[a].detailedStatic != 0) != (pass == 1))
                continue;
            for (int kr = 0; kr < p->atoms[a].Ntrans; ++kr)
                if (p->atoms[a].trans[kr].rhoPrd && nLines < 256)
                {
                    lineAtom[nLines] = a;
                    lineTrans[nLines] = kr;
                    ++nLines;
                }
        }
    if (nLines == 0)
    {
        free(lineAtom);
        free(lineTrans);
        return 0;
    }
    /* prdActive, la_to_prdLa (:739-757) */
    char* prdActive = (char*)calloc(Nspect, 1);
    int32_t* prdLaOfLa = (int32_t*)malloc(sizeof(int32_t) * Nspect);
    int NprdLa = 0;
    for (int la = 0; la < Nspect; ++la)
    {
        int present = 0;
        for (int q = 0; q < nLines; ++q)
            present = present || is_active(&p->atoms[lineAtom[q]].trans[lineTrans[q]], la);
        prdLaOfLa[la] = -1;
        if (present)
        {
            prdActive[la] = 1;
            prdLaOfLa[la] = NprdLa++;
        }
    }
    const double* wl = p->wavelength;
    int32_t* hPrdLaOfLa = (int32_t*)malloc(sizeof(int32_t) * (size_t)Ncol * Nspect);
    int* NhCol = (int*)calloc(Ncol, sizeof(int));
    int NhPrd = 0;
    for (int col = 0; col < Ncol; ++col)
    {
        const double* vlosMu = p->vlosMu + (size_t)col * M * K;
        for (int la = 0; la < Nspect; ++la)
        {
            /* check_lambda_scatter_into_prd_region (:765-797) */
            int scat = 0;
            for (int mu = 0; mu < M && !scat; ++mu)
                for (int toObs = 0; toObs <= 1 && !scat; ++toObs)
                    for (int k = 0; k < K && !scat; ++k)
                    {
                        const double s = sign[toObs];
                        const double fac = 1.0 + vlosMu[(size_t)mu * K + k] * s / C_CLIGHT;
                        const int prevIndex = la - 1 > 0 ? la - 1 : 0;
                        const int nextIndex = la + 1 < Nspect - 1 ? la + 1 : Nspect - 1;
                        const double prevLambda = wl[prevIndex] * fac;
                        const double nextLambda = wl[nextIndex] * fac;
                        int i = la;
                        for (; wl[i] > prevLambda && i > 0; --i);
                        for (; i < Nspect; ++i)
                        {
                            const double lambdaI = wl[i];
                            if (prdActive[i])
                            {
                                scat = 1;
                                break;
                            }
                            else if (lambdaI > nextLambda)
                                break;
                        }
                    }
            hPrdLaOfLa[(size_t)col * Nspect + la] = scat ? NhCol[col]++ : -1;
        }
        if (NhCol[col] > NhPrd)
            NhPrd = NhCol[col];
    }
    /* JCoeffs (:816-905) */
    const size_t nRows = (size_t)Ncol * NhPrd * M * 2 * K;
    JVec* rows = (JVec*)calloc(nRows ? nRows : 1, sizeof(JVec));
    for (int col = 0; col < Ncol; ++col)
    {
        const double* vlosMu = p->vlosMu + (size_t)col * M * K;
        for (int idx = 0; idx < Nspect; ++idx)
        {
            const int hPrdLa = hPrdLaOfLa[(size_t)col * Nspect + idx];
            if (hPrdLa < 0)
                continue;
            for (int mu = 0; mu < M; ++mu)
                for (int toObs = 0; toObs <= 1; ++toObs)
                    for (int k = 0; k < K; ++k)
                    {
                        JVec* cv = &rows[((((size_t)col * NhPrd + hPrdLa) * M + mu) * 2 + toObs) * K + k];
                        const double s = sign[toObs];
                        const double fac = 1.0 + vlosMu[(size_t)mu * K + k] * s / C_CLIGHT;
                        const int prevIndex = idx - 1 > 0 ? idx - 1 : 0;
                        const int nextIndex = idx + 1 < Nspect - 1 ? idx + 1 : Nspect - 1;
                        const double prevLambda = wl[prevIndex] * fac;
                        const double lambdaRest = wl[idx] * fac;
                        const double nextLambda = wl[nextIndex] * fac;
                        int doLowerHalf = 1, doUpperHalf = 1;
                        if (prevIndex == idx)
                        {
                            doLowerHalf = 0;
                            for (int i = 0; i < Nspect; ++i)
                            {
                                if (wl[i] <= lambdaRest && prdActive[i])
                                    jvec_push(cv, prdLaOfLa[i], 1.0);
                                else
                                    break;
                            }
                        }
                        else if (nextIndex == idx)
                        {
                            doUpperHalf = 0;
                            for (int i = Nspect - 1; i >= 0; --i)
                            {
                                if (wl[i] > lambdaRest && prdActive[i])
                                    jvec_push(cv, prdLaOfLa[i], 1.0);
                                else
                                    break;
                            }
                        }
                        int i = idx;
                        /* (the reference reads wavelength(-1) when the roll-back passes the first point:
                         * `spect.wavelength(i) > prevLambda && i >= 0` tests the array first; guarded here) */
                        for (; i >= 0 && wl[i] > prevLambda; --i);
                        if (i < 0)
                            i = 0;
                        for (; i < Nspect; ++i)
                        {
                            const double lambdaI = wl[i];
                            if (lambdaI > nextLambda)
                                break;
                            if (doLowerHalf && prdActive[i] && lambdaI > prevLambda && lambdaI <= lambdaRest)
                            {
                                const double frac = (lambdaI - prevLambda) / (lambdaRest - prevLambda);
                                jvec_push(cv, prdLaOfLa[i], frac);
                            }
                            else if (doUpperHalf && prdActive[i] && lambdaI > lambdaRest && lambdaI < nextLambda)
                            {
                                const double frac = (lambdaI - lambdaRest) / (nextLambda - lambdaRest);
                                jvec_push(cv, prdLaOfLa[i], 1.0 - frac);
                            }
                        }
                    }
        }
    }
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (nRows + 1));
    int64_t nnz = 0;
    for (size_t r = 0; r < nRows; ++r)
    {
        off[r] = nnz;
        nnz += rows[r].n;
    }
    off[nRows] = nnz;
    int32_t* cIdx = (int32_t*)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    double* cFrac = (double*)malloc(sizeof(double) * (nnz ? nnz : 1));
    for (size_t r = 0; r < nRows; ++r)
    {
        for (int64_t e = 0; e < rows[r].n; ++e)
        {
            cIdx[off[r] + e] = rows[r].v[e].idx;
            cFrac[off[r] + e] = rows[r].v[e].frac;
        }
        free(rows[r].v);
    }
    free(rows);
    /* hPrdCoeffs of every PRD line (:907-945) */
    int64_t* rhoOff = (int64_t*)malloc(sizeof(int64_t) * nLines);
    int64_t tot = 0;
    for (int q = 0; q < nLines; ++q)
    {
        const LwB200Transition* t = &p->atoms[lineAtom[q]].trans[lineTrans[q]];
        rhoOff[q] = tot;
        tot += (int64_t)Ncol * (t->Nred - t->Nblue) * M * 2 * K;
    }
    double* rhoFrac = (double*)malloc(sizeof(double) * tot);
    int32_t* rhoI0 = (int32_t*)malloc(sizeof(int32_t) * tot);
    for (int q = 0; q < nLines; ++q)
    {
        const LwB200Transition* t = &p->atoms[lineAtom[q]].trans[lineTrans[q]];
        const int Nl = t->Nred - t->Nblue;
        const double* w = t->wavelength;
        for (int col = 0; col < Ncol; ++col)
        {
            const double* vlosMu = p->vlosMu + (size_t)col * M * K;
            for (int lt = 0; lt < Nl; ++lt)
                for (int mu = 0; mu < M; ++mu)
                    for (int toObs = 0; toObs <= 1; ++toObs)
                        for (int k = 0; k < K; ++k)
                        {
                            const double s = sign[toObs];
                            const double lambdaRest = w[lt] * (1.0 + vlosMu[(size_t)mu * K + k] * s / C_CLIGHT);
                            const size_t o = (size_t)rhoOff[q] + ((((size_t)col * Nl + lt) * M + mu) * 2 + toObs) * K + k;
                            if (lambdaRest <= w[0])
                            {
                                rhoFrac[o] = 0.0;
                                rhoI0[o] = 0;
                            }
                            else if (lambdaRest >= w[Nl - 1])
                            {
                                rhoFrac[o] = 1.0;
                                rhoI0[o] = Nl - 2;
                            }
                            else
                            {
                                const double* it = upper_bound_d(w, w + Nl, lambdaRest) - 1;
                                rhoFrac[o] = (lambdaRest - *it) / (*(it + 1) - *it);
                                rhoI0[o] = (int32_t)(it - w);
                            }
                        }
        }
    }
    free(prdActive);
    free(NhCol);
    out->NprdLa = NprdLa;
    out->NhPrd = NhPrd;
    out->Nlines = nLines;
    out->prdLaOfLa = prdLaOfLa;
    out->hPrdLaOfLa = hPrdLaOfLa;
    out->JRest = (double*)calloc((size_t)Ncol * NprdLa * K, sizeof(double));
    out->JCoeffOff = off;
    out->JCoeffIdx = cIdx;
    out->JCoeffFrac = cFrac;
    out->lineAtom = lineAtom;
    out->lineTrans = lineTrans;
    out->rhoCoefOff = rhoOff;
    out->rhoFrac = rhoFrac;
    out->rhoI0 = rhoI0;
    return 0;
}

void lwo_free_hprd(LwB200HybridPrd* h)
{
    free((void*)h->prdLaOfLa); free((void*)h->hPrdLaOfLa); free(h->JRest); free((void*)h->JCoeffOff);
    free((void*)h->JCoeffIdx); free((void*)h->JCoeffFrac); free((void*)h->lineAtom); free((void*)h->lineTrans);
    free((void*)h->rhoCoefOff); free((void*)h->rhoFrac); free((void*)h->rhoI0);
    memset(h, 0, sizeof(*h));
}

/* ------------------------------------------------------------------------ */
/* Full Stokes: formal_sol_full_stokes_impl (FormalStokes.cpp:664-723), stokes_fs_core
 * (:418-661), piecewise_stokes_bezier3_1d(_impl) (:166-413). */

/* stokes_K, FormalStokes.cpp:119-143; chi is [7][K] */
static void stokes_K(int k, const double* chi, int K, double chiI, double* Km)
{
    memset(Km, 0, sizeof(double) * 16);
    Km[0 * 4 + 1] = chi[(size_t)1 * K + k];
    Km[0 * 4 + 2] = chi[(size_t)2 * K + k];
    Km[0 * 4 + 3] = chi[(size_t)3 * K + k];
    Km[1 * 4 + 2] = chi[(size_t)6 * K + k];
    Km[1 * 4 + 3] = chi[(size_t)5 * K + k];
    Km[2 * 4 + 3] = chi[(size_t)4 * K + k];
    for (int j = 0; j < 3; ++j)
        for (int i = j + 1; i < 4; ++i)
        {
            Km[j * 4 + i] /= chiI;
            Km[i * 4 + j] = Km[j * 4 + i];
        }
    Km[1 * 4 + 3] *= -1.0;
    Km[2 * 4 + 1] *= -1.0;
    Km[3 * 4 + 2] *= -1.0;
}

/* prod(a, b, c): c(j, i) += a(k, i) * b(j, k), FormalStokes.cpp:145-153 */
static void prod44(const double* a, const double* b, double* c)
{
    memset(c, 0, sizeof(double) * 16);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k)
                c[j * 4 + i] += a[k * 4 + i] * b[j * 4 + k];
}

/* piecewise_stokes_bezier3_1d_impl, FormalStokes.cpp:166-340.  chi [7][K], S [4][K], I [4][K]. */
static void stokes_bezier3_sweep(int Ndep, const double* height, const double* chi, const double* S, double* I,
                                 double zmu, int toObs, const double* Istart)
{
    static const double id[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    const int K = Ndep;
    int dk = -1, k_start = Ndep - 1, k_end = 0;
    if (!toObs)
    {
        dk = 1;
        k_start = 0;
        k_end = Ndep - 1;
    }
    for (int n = 0; n < 4; ++n)
        I[(size_t)n * K + k_start] = Istart[n];

    int k = k_start + dk;
    double ds_uw = fabs(height[k] - height[k - dk]) * zmu;
    double ds_dw = fabs(height[k + dk] - height[k]) * zmu;
    double dx_uw = (chi[k] - chi[k - dk]) / ds_uw;
    double dx_c = cent_deriv(ds_uw, ds_dw, chi[k - dk], chi[k], chi[k + dk]);
    double c1 = chi[k] - (ds_uw / 3.0) * dx_c;
    double c2 = chi[k - dk] + (ds_uw / 3.0) * dx_uw;
    double dtau_uw = ds_uw * (chi[k] + chi[k - dk] + c1 + c2) * 0.25;

    double Ku[16], dKu[16], K0[16], dK0[16], Su[4], dSu[4], S0[4], dS0[4];
    double Kd[16], K02[16], Ku2[16], Ma[16], Mb[16], Mc[16], Md[16], V0[4], Sd[4];
    stokes_K(k_start, chi, K, chi[k_start], Ku);
    stokes_K(k, chi, K, chi[k], K0);
    for (int n = 0; n < 4; ++n)
    {
        Su[n] = S[(size_t)n * K + k_start];
        S0[n] = S[(size_t)n * K + k];
    }
    for (int n = 0; n < 4; ++n)
    {
        dSu[n] = (S0[n] - Su[n]) / dtau_uw;
        for (int m = 0; m < 4; ++m)
            dKu[n * 4 + m] = (K0[n * 4 + m] - Ku[n * 4 + m]) / dtau_uw;
    }
    double ds_dw2 = 0.0, dtau_dw = 0.0, dx_dw = 0.0;
    memset(Kd, 0, sizeof(Kd));
    memset(Sd, 0, sizeof(Sd));
    memset(dK0, 0, sizeof(dK0));
    memset(dS0, 0, sizeof(dS0));
    for (; k != k_end + dk; k += dk)
    {
        if (k == k_end)
        {
            for (int n = 0; n < 4; ++n)
            {
                dS0[n] = (S0[n] - Su[n]) / dtau_uw;
                for (int m = 0; m < 4; ++m)
                    dK0[n * 4 + m] = (K0[n * 4 + m] - Ku[n * 4 + m]) / dtau_uw;
            }
        }
        else
        {
            if (k_end - k == dk)
                dx_dw = (chi[k + dk] - chi[k]) / ds_dw;
            else
            {
                ds_dw2 = fabs(height[k + 2 * dk] - height[k + dk]) * zmu;
                dx_dw = cent_deriv(ds_dw, ds_dw2, chi[k], chi[k + dk], chi[k + 2 * dk]);
            }
            c1 = chi[k] + (ds_dw / 3.0) * dx_c;
            c2 = chi[k + dk] - (ds_dw / 3.0) * dx_dw;
            dtau_dw = ds_dw * (chi[k] + chi[k + dk] + c1 + c2) * 0.25;
            stokes_K(k + dk, chi, K, chi[k + dk], Kd);
            for (int n = 0; n < 4; ++n)
                Sd[n] = S[(size_t)n * K + k + dk];
            for (int q = 0; q < 16; ++q)
                dK0[q] = cent_deriv(dtau_uw, dtau_dw, Ku[q], K0[q], Kd[q]);
            for (int q = 0; q < 4; ++q)
                dS0[q] = cent_deriv(dtau_uw, dtau_dw, Su[q], S0[q], Sd[q]);
        }
        prod44(Ku, Ku, Ku2);
        prod44(K0, K0, K02);
        double alpha, beta, gamma, delta, edt;
        bezier3_coeffs(dtau_uw, &alpha, &beta, &gamma, &delta, &edt);
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i)
            {
                const int q = j * 4 + i;
                double d = dtau_uw / 3.0 * (Ku2[q] + Ku[q] - dKu[q]) - Ku[q];
                double e = dtau_uw / 3.0 * (K02[q] + K0[q] - dK0[q]) + K0[q];
                Md[q] = id[j][i] + beta * K0[q] + delta * e;
                Ma[q] = edt * id[j][i] - alpha * Ku[q] + gamma * d;
                Mb[q] = alpha * id[j][i] + gamma * (id[j][i] - (dtau_uw / 3.0) * Ku[q]);
                Mc[q] = beta * id[j][i] + delta * (id[j][i] + (dtau_uw / 3.0) * K0[q]);
            }
        for (int i = 0; i < 4; ++i)
        {
            V0[i] = 0.0;
            for (int j = 0; j < 4; ++j)
                V0[i] += Ma[i * 4 + j] * I[(size_t)j * K + k - dk] + Mb[i * 4 + j] * Su[j] + Mc[i * 4 + j] * S0[j];
            V0[i] += (dtau_uw / 3.0) * (gamma * dSu[i] - delta * dS0[i]);
        }
        lwo_solve_lin_eq(4, Md, V0, 1);
        for (int i = 0; i < 4; ++i)
            I[(size_t)i * K + k] = V0[i];
        memcpy(Su, S0, sizeof(Su));
        memcpy(S0, Sd, sizeof(S0));
        memcpy(dSu, dS0, sizeof(dSu));
        memcpy(Ku, K0, sizeof(Ku));
        memcpy(K0, Kd, sizeof(K0));
        memcpy(dKu, dK0, sizeof(dKu));
        dtau_uw = dtau_dw;
        ds_uw = ds_dw;
        ds_dw = ds_dw2;
        dx_uw = dx_c;
        dx_c = dx_dw;
    }
}

int lwo_full_stokes(const LwB200Problem* p, int col, int updateJ, int upOnly, double* dJMaxOut, int64_t* dJMaxIdx)
{
    return lwo_full_stokes_j20(p, col, updateJ, upOnly, NULL, dJMaxOut, dJMaxIdx);
}

/* J20all: the "J20" extra parameter (FormalStokes.cpp:676-681), [Ncol][Nspect][Nspace], or NULL. */
int lwo_full_stokes_j20(const LwB200Problem* p, int col, int updateJ, int upOnly, double* J20all, double* dJMaxOut,
                        int64_t* dJMaxIdx)
{
    const int K = p->Nspace, M = p->Nrays, L = p->Nspect;
    if (!p->Quv)
        return 1;
    double* J20 = J20all ? J20all + (size_t)col * L * K : NULL;
    double* J20Dag = (double*)calloc(K, sizeof(double)); /* F64Arr(Nspace): zero until a J-updating wavelength fills it */
    const double inv2root2 = 1.0 / (2.0 * sqrt(2.0));
    const double* h = p->height + (size_t)col * K;
    const double* T = p->temperature + (size_t)col * K;
    Scratch* s = scratch_new(p);
    double* chiTot = (double*)calloc((size_t)7 * K, sizeof(double));
    double* etaTot = (double*)calloc((size_t)4 * K, sizeof(double));
    double* S = (double*)calloc((size_t)4 * K, sizeof(double));
    double* I = (double*)calloc((size_t)4 * K, sizeof(double)); /* persists over rays, as in the reference */
    double dJMax = 0.0;
    int64_t idx = 0;
    for (int la = 0; la < L; ++la)
    {
        double* J = p->J + ((size_t)col * L + la) * K;
        const double* bgChi = p->chiBg + ((size_t)col * L + la) * K;
        const double* bgEta = p->etaBg + ((size_t)col * L + la) * K;
        const double* bgSca = p->scaBg + ((size_t)col * L + la) * K;
        const double wav = p->wavelength[la];
        if (updateJ)
        {
            memcpy(s->JDag, J, sizeof(double) * K);
            memset(J, 0, sizeof(double) * K);
            if (J20)
            {
                memcpy(J20Dag, J20 + (size_t)la * K, sizeof(double) * K);
                memset(J20 + (size_t)la * K, 0, sizeof(double) * K);
            }
        }
        for (int a = 0; a < p->Natom; ++a)
            setup_wavelength(p, col, a, la, s);
        /* (need full integration if we're doing J20, :469-471) */
        const int contOnly = J20 ? 0 : continua_only(p, la);
        const int toObsStart = upOnly ? 1 : 0;
        for (int mu = 0; mu < M; ++mu)
        {
            const double mu2 = sq(p->muz[mu]);
            const double wJ20_I = inv2root2 * (3.0 * mu2 - 1.0);
            const double wJ20_Q = inv2root2 * 3.0 * (mu2 - 1.0);
            for (int toObs = toObsStart; toObs < 2; ++toObs)
            {
                /* NOTE: as in the reference, polarisedFrequency is only (re)determined when the
                 * opacities are gathered; a continua-only wavelength has no polarised line. */
                static int polarisedFrequency;
                if (!contOnly || (mu == 0 && toObs == toObsStart))
                {
                    polarisedFrequency = J20 ? 1 : 0; /* bool polarisedFrequency = J20 || false  (:490) */
                    memset(chiTot, 0, sizeof(double) * 7 * K);
                    memset(etaTot, 0, sizeof(double) * 4 * K);
                    for (int pass = 0; pass < 2; ++pass) /* active atoms, then detailed ones */
                        for (int a = 0; a < p->Natom; ++a)
                        {
                            const LwB200Atom* at = &p->atoms[a];
                            if ((at->detailedStatic != 0) != (pass == 1))
                                continue;
                            const double* n = at->n + (size_t)col * at->Nlevel * K;
                            for (int kr = 0; kr < at->Ntrans; ++kr)
                            {
                                const LwB200Transition* t = &at->trans[kr];
                                if (!is_active(t, la))
                                    continue;
                                uv(p, col, a, kr, la, mu, toObs, s);
                                const int Nl = t->Nred - t->Nblue, lt = la - t->Nblue;
                                const size_t per = (size_t)Nl * M * 2 * K, arr = (size_t)p->Ncol * per;
                                const size_t off = (size_t)col * per + (((size_t)lt * M + mu) * 2 + toObs) * K;
                                for (int k = 0; k < K; ++k)
                                {
                                    double chi = n[(size_t)t->i * K + k] * s->Vij[k] - n[(size_t)t->j * K + k] * s->Vji[k];
                                    double eta = n[(size_t)t->j * K + k] * s->Uji[k];
                                    chiTot[k] += chi;
                                    etaTot[k] += eta;
                                    if (t->type == LWB200_LINE && t->polProfiles)
                                    {
                                        polarisedFrequency = 1;
                                        const double phi = t->phi[off + k];
                                        const double* pol = t->polProfiles + off + k;
                                        double chiNoProfile = chi / phi;
                                        chiTot[(size_t)1 * K + k] += chiNoProfile * pol[0 * arr];
                                        chiTot[(size_t)2 * K + k] += chiNoProfile * pol[1 * arr];
                                        chiTot[(size_t)3 * K + k] += chiNoProfile * pol[2 * arr];
                                        chiTot[(size_t)4 * K + k] += chiNoProfile * pol[3 * arr];
                                        chiTot[(size_t)5 * K + k] += chiNoProfile * pol[4 * arr];
                                        chiTot[(size_t)6 * K + k] += chiNoProfile * pol[5 * arr];
                                        double etaNoProfile = eta / phi;
                                        etaTot[(size_t)1 * K + k] += etaNoProfile * pol[0 * arr];
                                        etaTot[(size_t)2 * K + k] += etaNoProfile * pol[1 * arr];
                                        etaTot[(size_t)3 * K + k] += etaNoProfile * pol[2 * arr];
                                    }
                                }
                            }
                        }
                    if (J20) /* :575-583 */
                        for (int k = 0; k < K; ++k)
                        {
                            etaTot[k] += wJ20_I * bgSca[k] * J20Dag[k];
                            etaTot[(size_t)1 * K + k] += wJ20_Q * bgSca[k] * J20Dag[k];
                        }
                    for (int k = 0; k < K; ++k)
                    {
                        chiTot[k] += bgChi[k];
                        S[k] = (etaTot[k] + bgEta[k] + bgSca[k] * s->JDag[k]) / chiTot[k];
                    }
                    if (polarisedFrequency)
                        for (int n = 1; n < 4; ++n)
                            for (int k = 0; k < K; ++k)
                                S[(size_t)n * K + k] = etaTot[(size_t)n * K + k] / chiTot[k];
                }
                if (!polarisedFrequency)
                {
                    lwo_solve_ray(2, K, h, T, chiTot, S, p->muz[mu], toObs, wav, p->lowerBc, p->upperBc,
                                  bc_value(p, col, la, mu, toObs), I, NULL);
                }
                else
                {
                    /* piecewise_stokes_bezier3_1d, :344-413 */
                    const double zmu = 1.0 / p->muz[mu];
                    int dk = -1, kStart = K - 1;
                    if (!toObs)
                    {
                        dk = 1;
                        kStart = 0;
                    }
                    double dtau_uw = 0.5 * zmu * (chiTot[kStart] + chiTot[kStart + dk]) * fabs(h[kStart] - h[kStart + dk]);
                    double Iupw[4] = {0.0, 0.0, 0.0, 0.0};
                    if (toObs)
                    {
                        if (p->lowerBc == LWB200_BC_THERMALISED)
                        {
                            double Bnu[2];
                            planck_nu(2, &T[K - 2], wav, Bnu);
                            Iupw[0] = Bnu[1] - (Bnu[0] - Bnu[1]) / dtau_uw;
                        }
                        else if (p->lowerBc == LWB200_BC_CALLABLE)
                            Iupw[0] = bc_value(p, col, la, mu, toObs);
                    }
                    else
                    {
                        if (p->upperBc == LWB200_BC_THERMALISED)
                        {
                            double Bnu[2];
                            planck_nu(2, &T[0], wav, Bnu);
                            Iupw[0] = Bnu[0] - (Bnu[1] - Bnu[0]) / dtau_uw;
                        }
                        else if (p->upperBc == LWB200_BC_CALLABLE)
                            Iupw[0] = bc_value(p, col, la, mu, toObs);
                    }
                    stokes_bezier3_sweep(K, h, chiTot, S, I, zmu, toObs, Iupw);
                }
                p->I[((size_t)col * L + la) * M + mu] = I[0];
                for (int q = 0; q < 3; ++q)
                    p->Quv[(((size_t)col * 3 + q) * L + la) * M + mu] = I[(size_t)(q + 1) * K];
                if (updateJ)
                {
                    const double wmu = p->wmu[mu];
                    for (int k = 0; k < K; ++k)
                        J[k] += 0.5 * wmu * I[k];
                    if (J20) /* :642-648 */
                    {
                        const double wmuJ20_I = wJ20_I * wmu;
                        const double wmuJ20_Q = wJ20_Q * wmu;
                        for (int k = 0; k < K; ++k)
                            J20[(size_t)la * K + k] += wmuJ20_I * I[k] + wmuJ20_Q * I[(size_t)1 * K + k];
                    }
                }
            }
        }
        if (updateJ)
        {
            double dJ = 0.0;
            for (int k = 0; k < K; ++k)
                dJ = dmax(fabs(1.0 - s->JDag[k] / J[k]), dJ);
            if (dJMax < dJ)
            {
                dJMax = dJ;
                idx = la;
            }
        }
    }
    free(chiTot);
    free(etaTot);
    free(S);
    free(I);
    free(J20Dag);
    scratch_free(s);
    if (dJMaxOut) *dJMaxOut = updateJ ? dJMax : 0.0;
    if (dJMaxIdx) *dJMaxIdx = updateJ ? idx : 0;
    return 0;
}

