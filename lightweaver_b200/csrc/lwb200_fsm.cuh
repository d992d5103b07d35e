// lwb200_fsm.cuh -- the formal solvers of the production ray kernel
// (lwb200_pipeline.cuh): one ray of one wavelength, lanes over depth.
//
// The formal solver is the same arithmetic as lwb200_device.cuh, restructured
// for the B200 issue/fp64 pipes: one straight-line, branch-free code path for
// both ray directions and all optical-depth regimes (selects instead of
// divergent branches, so the three depth chunks of a lane interleave and no
// shuffle needs re-convergence), divisions by ray-independent geometry replaced
// by precomputed reciprocals, the rest by a Newton reciprocal, and a
// table-free exp.  Differences from the reference are at rounding level.
#pragma once
#include "lwb200_kernels.cuh"

namespace lwb200
{
// 1/x to ~1 ulp: hardware seed (relative error e0 <= 2^-20) + one cubic step,
// r = r0 (1 + e + e^2) with e = 1 - x r0, which leaves e0^3 < 2^-60 in three dependent fused
// multiply-adds (two Newton steps take four).  No IEEE corner cases needed: every argument here is a
// finite positive opacity, path length or optical depth.
__device__ __forceinline__ double rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifdef LWB200_RCP_NEWTON2
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
#else
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
#endif
    return r;
}

// exp(x) for |x| < 700, ~1 ulp: Cody-Waite reduction, degree-13 Taylor on |r| <= ln2/2
__device__ __forceinline__ double exp_fast(double x)
{
    const double magic = 6755399441055744.0; // 1.5 * 2^52
    double t = fma(x, 1.4426950408889634074, magic);
    const int n = __double2loint(t);
    t -= magic;
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
#ifdef LWB200_EXP_HORNER
    double p = 1.6059043836821613e-10;           // 1/13!
    p = fma(p, r, 2.08767569878681e-09);         // 1/12!
    p = fma(p, r, 2.505210838544172e-08);        // 1/11!
    p = fma(p, r, 2.755731922398589e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
    p = fma(p, r, 2.48015873015873e-05);         // 1/8!
    p = fma(p, r, 1.984126984126984e-04);        // 1/7!
    p = fma(p, r, 1.388888888888889e-03);        // 1/6!
    p = fma(p, r, 8.333333333333333e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#else
    // Estrin's scheme: the same degree-13 Taylor polynomial with a dependency depth of 5
    // instead of 14 (the kernel is bound by dependent-issue latency, not by fp64 throughput)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double q01 = fma(r, 1.0, 1.0);
    const double q23 = fma(r, 1.6666666666666666e-01, 0.5);
    const double q45 = fma(r, 8.333333333333333e-03, 4.1666666666666664e-02);
    const double q67 = fma(r, 1.984126984126984e-04, 1.388888888888889e-03);
    const double q89 = fma(r, 2.7557319223985893e-06, 2.48015873015873e-05);
    const double qab = fma(r, 2.505210838544172e-08, 2.755731922398589e-07);
    const double qcd = fma(r, 1.6059043836821613e-10, 2.08767569878681e-09);
    const double s03 = fma(r2, q23, q01);
    const double s47 = fma(r2, q67, q45);
    const double s8b = fma(r2, qab, q89);
    const double t07 = fma(r4, s47, s03);
    const double t8d = fma(r4, qcd, s8b);
    double p = fma(r8, t8d, t07);
#endif
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// exp_fast for arguments that are not bounded below (the Boltzmann factor exp(-hc / (k lambda T)) of a
// bound-free continuum at a short wavelength and a cool depth): the exponent arithmetic of exp_fast
// wraps for x < -708.4, where the reference's libm exp() has long underflowed to (sub)normal nothing.
__device__ __forceinline__ double exp_fast_underflow(double x)
{
    const double e = exp_fast((x < -708.0) ? -708.0 : x);
    return (x < -708.0) ? 0.0 : e;
}

template <int NCH>
struct GeometryR
{
    int K;
    double dsf[NCH];   // |h_k - h_{k+1}|
    double dsfP[NCH];  // |h_{k-1} - h_k|
    double rdsf[NCH];  // 1 / dsf
    double rsum[NCH];  // 1 / (dsf + dsfP)
    double rdsfP0;     // 1 / dsfP[0]
};

template <int NCH, bool MULTI>
__device__ __forceinline__ void load_geometry_r(DepthComm<MULTI>& cm, GeometryR<NCH>& g,
                                                const double* __restrict__ height, int K)
{
    const int lane = cm.lane_global();
    g.K = K;
    double h[NCH], hN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        h[j] = (k < K) ? __ldg(height + k) : 0.0;
    }
    shift_next<NCH>(cm, h, hN);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        g.dsf[j] = (k + 1 < K) ? fabs(h[j] - hN[j]) : 1.0;
    }
    shift_prev<NCH>(cm, g.dsf, g.dsfP);
    if (lane == 0)
        g.dsfP[0] = 1.0;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        g.rdsf[j] = 1.0 / g.dsf[j];
        g.rsum[j] = 1.0 / (g.dsf[j] + g.dsfP[j]);
    }
    g.rdsfP0 = 1.0 / g.dsfP[0];
}

__device__ __forceinline__ double steffen_r(double wUw, double wDw, double Suw, double S0)
{
    // wUw = dsdw/(dsdw+dsuw) multiplies Suw, wDw = dsuw/(dsdw+dsuw) multiplies S0 (Bezier.hpp:58-65):
    //   (sign(S0) + sign(Suw)) * min(|Suw|, |S0|, P0/2)
    // = 0 if the slopes differ in sign, else 2 min(...) with their common sign.  The sign test and
    // the sign transfer are integer operations on the high words (no fp64 issue slots), the
    // minima plain compare/selects (no NaN path: the inputs are finite).
    const double P0h = 0.5 * fabs(fma(Suw, wUw, S0 * wDw));
    const double aU = fabs(Suw), a0 = fabs(S0);
    double m = (a0 < P0h) ? a0 : P0h;
    m = (aU < m) ? aU : m;
    m += m;
    const int h0 = __double2hiint(S0), hU = __double2hiint(Suw);
    const int hm = __double2hiint(m) | (h0 & 0x80000000);
    const double r = __hiloint2double(hm, __double2loint(m));
    return ((h0 ^ hU) < 0) ? 0.0 : r;
}

__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

// optical depth clamped to [0, 700] for exp_fast (plain compare/selects)
__device__ __forceinline__ double clamp_dt(double dt)
{
    dt = (dt < 0.0) ? 0.0 : dt;
    return (dt > 700.0) ? 700.0 : dt;
}

// piecewise_bezier3_1d (FormalScalar.cpp:209-325, :535-600) in two phases.
//
// bezier3_prepare: everything that does not depend on the direction of the ray, in
// array-forward orientation: the Bezier optical depth dtf of every interval (k, k+1)
// (:242-246, :261-263) from the Steffen derivatives of chi (Bezier.hpp:58-65), and the
// Steffen derivative DSf of S with respect to optical depth (:247).  The derivative along
// an up-going ray is exactly the negative (all products commute bitwise), so the up and
// down rays of one mu share this phase whenever their opacities are identical (static
// atmosphere), and at line-free wavelengths it is computed ONCE per wavelength at mu = 1
// and rescaled (dtf ~ 1/mu, DSf ~ mu: every step above is homogeneous in mu).
template <int NCH>
struct RayPre
{
    double SN[NCH];   // S at k + 1
    double dtf[NCH];  // optical depth of the interval (k, k + 1)
    double rdtf[NCH]; // 1 / dtf
    double DSf[NCH];  // dS/dtau at k, forward orientation
};

template <int NCH, bool MULTI>
__device__ __forceinline__ void bezier3_prepare(DepthComm<MULTI>& cm, const GeometryR<NCH>& g,
                                                const double (&chi)[NCH], const double (&S)[NCH], double muz,
                                                double zmu, RayPre<NCH>& r)
{
    const int lane = cm.lane_global();
    const int K = g.K;
    double chiN[NCH];
    shift_next<NCH>(cm, chi, chiN);
    shift_next<NCH>(cm, S, r.SN);

    // chi slopes on forward intervals and Steffen derivatives
    double sl[NCH], Df[NCH], DfN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        sl[j] = (chiN[j] - chi[j]) * (g.rdsf[j] * muz);
    const double slUp = cm.from_prev(sl[NCH - 1]);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const double slPj = (j == 0) ? slUp : sl[j > 0 ? j - 1 : 0];
        double d = steffen_r(g.dsf[j] * g.rsum[j], g.dsfP[j] * g.rsum[j], slPj, sl[j]);
        d = sel(k == 0, sl[j], d);       // one-sided at the top    (:239 / :288)
        d = sel(k == K - 1, slPj, d);    // one-sided at the bottom
        Df[j] = d;
    }
    shift_next<NCH>(cm, Df, DfN);

    double dtfP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double ds = g.dsf[j] * zmu;
        const double ds3 = ds * (1.0 / 3.0);
        const double cA = fma(ds3, Df[j], chi[j]);
        const double cB = fma(-ds3, DfN[j], chiN[j]);
        const double t1 = chi[j] + chiN[j];
        r.dtf[j] = ds * ((t1 + cA) + cB) * 0.25;
        r.rdtf[j] = rcp_fast(r.dtf[j]);
    }
    shift_prev<NCH>(cm, r.dtf, dtfP);

    double slS[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        slS[j] = (r.SN[j] - S[j]) * r.rdtf[j];
    const double slSUp = cm.from_prev(slS[NCH - 1]);
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

// element jx (warp-uniform) of a register array
template <int NCH>
__device__ __forceinline__ double pick(const double (&v)[NCH], int jx)
{
    double x = v[0];
#pragma unroll
    for (int j = 1; j < NCH; ++j)
        x = (jx == j) ? v[j] : x;
    return x;
}

// The two special points of a ray -- the first carries the boundary intensity (:551-597), the last is
// piecewise linear through w2() (:307-321, LwInternal.hpp:90-110) -- depend on chi and S at the two
// depth pairs (0, 1) and (K - 2, K - 1) only.  They are evaluated ONCE PER WAVELENGTH for all its rays
// at once, lane r of the warp taking ray r = 2 mu + dir (ray_endpoints), instead of once per ray by every
// lane: out of the ray loop go two serial chains of ~40 dependent fp64 operations (an exp and two
// reciprocals) that no depth parallelism could hide.
struct RayEnds
{
    double Iupw; // intensity entering at the first point
    double aE;   // last point: I = aE * I_upw + bE, psi = pE
    double bE;
    double pE;
};

// depth k of a register array laid over the warp: valid in every lane
template <int NCH>
__device__ __forceinline__ double at_depth(const double (&v)[NCH], int k)
{
    return __shfl_sync(kFull, pick<NCH>(v, k % NCH), k / NCH);
}

// chiK / SK: opacity and source function of THIS LANE'S RAY at depths 0, 1, K - 2, K - 1;
// dsTop = |h_0 - h_1|, dsBot = |h_{K-2} - h_{K-1}|; dir = 1: up-going (starts at the bottom).
__device__ __forceinline__ RayEnds ray_endpoints(const double (&chiK)[4], const double (&SK)[4], double dsTop,
                                                 double dsBot, double zmu, int dir, int bcType, double bcB0,
                                                 double bcB1, double bcValue)
{
    RayEnds e;
    const bool up = dir != 0;
    // first point and its inner neighbour
    const double chiS = up ? chiK[3] : chiK[0], chiSN = up ? chiK[2] : chiK[1];
    const double dsS = up ? dsBot : dsTop;
    double Iupw = 0.0;
    if (bcType == 2)
    {
        const double dtau_b = 0.5 * zmu * (chiS + chiSN) * dsS;
        Iupw = bcB0 - (bcB1 - bcB0) / dtau_b;
    }
    else if (bcType == 4)
        Iupw = bcValue;
    e.Iupw = Iupw;
    // last point and its upwind neighbour
    const double chiE = up ? chiK[0] : chiK[3], chiUE = up ? chiK[1] : chiK[2];
    const double SE = up ? SK[0] : SK[3], SUE = up ? SK[1] : SK[2];
    const double dsE = up ? dsTop : dsBot;
    const double dt = 0.5 * zmu * (chiE + chiUE) * dsE;
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
        const double ex = exp(-dt);
        w0 = 1.0 - ex;
        w1 = w0 - dt * ex;
    }
    const double dS = (SE - SUE) / dt;
    e.aE = 1.0 - w0;
    e.bE = w0 * SE - w1 * dS;
    e.pE = (w0 - w1 / dt) / chiE;
    return e;
}

// bezier3_coeffs: the direction-dependent phase.  Bezier3 coefficients of every point
// (Bezier.hpp:81-127; both optical-depth regimes through selects, dt > 30 is the closed form
// with edt = 0 term for term), then the two special points of the ray patched in from `ends`
// (this ray's RayEnds, already broadcast).  Outputs the recurrence I_k = a_k I_upwind + b_k and
// psi = Psi* / chi; padding lanes carry the identity.
template <int NCH, bool DOWN, bool MULTI>
__device__ __forceinline__ void bezier3_coeffs(DepthComm<MULTI>& cm, const GeometryR<NCH>& g, const double (&S)[NCH],
                                               const double (&rchi)[NCH], const RayPre<NCH>& r, const RayEnds& ends,
                                               double (&a)[NCH], double (&b)[NCH], double (&psi)[NCH])
{
    const int lane = cm.lane_global();
    const int K = g.K;
    const int kS = DOWN ? 0 : K - 1;
    const int kE = DOWN ? K - 1 : 0;
    // upwind neighbours along the ray
    double SU[NCH], dtU[NCH], rdtU[NCH], DSU[NCH];
    if (DOWN)
    {
        shift_prev<NCH>(cm, S, SU);
        shift_prev<NCH>(cm, r.dtf, dtU);
        shift_prev<NCH>(cm, r.rdtf, rdtU);
        shift_prev<NCH>(cm, r.DSf, DSU);
    }
    else
    {
        shift_next<NCH>(cm, r.DSf, DSU);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            SU[j] = r.SN[j];
            dtU[j] = r.dtf[j];
            rdtU[j] = r.rdtf[j];
            DSU[j] = -DSU[j];
        }
    }

#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double dt = dtU[j], rdt = rdtU[j];
        // (an optical depth beyond 30 takes edt = 0; the clamp only keeps exp_fast in its range)
        const double ex = exp_fast(-((dt > 700.0) ? 700.0 : dt));
        const double dt2 = dt * dt, dt3 = dt2 * dt;
        const double edtC = sel(dt > 30.0, 0.0, ex);
        const double rdt3 = rdt * rdt * rdt;
        const double alphaC = (6.0 - edtC * (6.0 + 6.0 * dt + 3.0 * dt2 + dt3)) * rdt3;
        const double betaC = (6.0 * edtC - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) * rdt3;
        const double gammaC = 3.0 * (2.0 * dt - 6.0 + edtC * (6.0 + 4.0 * dt + dt2)) * rdt3;
        const double deltaC = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edtC * (3.0 + dt)) * rdt3;
        const double edtT = 1.0 - dt + 0.5 * dt2 - dt3 * (1.0 / 6.0);
        const double alphaT = 0.25 * dt - 0.2 * dt2 + dt3 * (1.0 / 12.0);
        const double betaT = 0.25 * dt - 0.05 * dt2 + dt3 * (1.0 / 120.0);
        const double gammaT = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
        const double deltaT = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
        const bool tay = dt < 5e-2;
        const double alpha = sel(tay, alphaT, alphaC), beta = sel(tay, betaT, betaC);
        const double gamma = sel(tay, gammaT, gammaC), delta = sel(tay, deltaT, deltaC);
        const double dt3rd = dt * (1.0 / 3.0);
        const double Cuw = fma(dt3rd, DSU[j], SU[j]);
        const double C0 = DOWN ? fma(-dt3rd, r.DSf[j], S[j]) : fma(dt3rd, r.DSf[j], S[j]);
        a[j] = sel(tay, edtT, edtC);
        b[j] = alpha * SU[j] + beta * S[j] + gamma * Cuw + delta * C0;
        psi[j] = (beta + delta) * rchi[j];
    }

    // ---- the two special points of the ray, and the identity in the padding lanes
    const int jE = kE % NCH, laneE = kE / NCH, jS = kS % NCH, laneS = kS / NCH;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const bool last = lane == laneE && j == jE, first = lane == laneS && j == jS;
        const bool pad = !DOWN && lane * NCH + j >= K; // (DOWN: padding comes after every real point, never used)
        a[j] = first ? 0.0 : (last ? ends.aE : (pad ? 1.0 : a[j]));
        b[j] = first ? ends.Iupw : (last ? ends.bE : (pad ? 0.0 : b[j]));
        psi[j] = first ? 0.0 : (last ? ends.pE : psi[j]);
    }
}

// ---------------------------------------------------------------------------
// The same Bezier3 step split ONCE MORE, by what it depends on.  Alpha .. delta and exp(-dt) are
// functions of the optical depth of an INTERVAL (k, k + 1), whichever way a ray crosses it; the
// recurrence coefficient of a point regroups as
//     b = (alpha + gamma) S_uw + (beta + delta) S_0 + gamma dt/3 S'_uw - delta dt/3 S'_0
// (C_uw = S_uw + dt/3 S'_uw, C_0 = S_0 - dt/3 S'_0 along the ray, Bezier.hpp:81-127 / FormalScalar.cpp:276-281),
// so the expensive part -- the exp, the four cubic-over-cubic coefficients in both regimes -- is
// evaluated per interval in array-forward orientation by ONE piece of code (interval_coeffs), shared
// by the up and the down ray of a mu whenever they share opacities, and what is left per direction is
// a handful of fused multiply-adds on shifted copies (ray_down / ray_up).  Besides the arithmetic saved,
// this keeps the hot loop of the ray kernels small: measured with ncu on B200, the loop with both
// directions' full coefficient code inlined (~40 KB of SASS) did not fit the 32 KB instruction cache
// and stalled on instruction fetch as often as on the fp64 pipe.
template <int NCH>
struct RayCoef
{
    double ed[NCH]; // exp(-dt) of the interval (k, k + 1) (its Taylor / thick-limit forms as Bezier3_coeffs)
    double AG[NCH]; // alpha + gamma
    double GD[NCH]; // gamma dt / 3
    double BD[NCH]; // beta + delta
    double DD[NCH]; // delta dt / 3
};

template <int NCH>
__device__ __forceinline__ void interval_coeffs(const RayPre<NCH>& r, RayCoef<NCH>& c)
{
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double dt = r.dtf[j], rdt = r.rdtf[j];
        // (an optical depth beyond 30 takes edt = 0; the clamp only keeps exp_fast in its range)
        const double ex = exp_fast(-((dt > 700.0) ? 700.0 : dt));
        const double dt2 = dt * dt, dt3 = dt2 * dt;
        const double edtC = sel(dt > 30.0, 0.0, ex);
        const double rdt3 = rdt * rdt * rdt;
        const double alphaC = (6.0 - edtC * (6.0 + 6.0 * dt + 3.0 * dt2 + dt3)) * rdt3;
        const double betaC = (6.0 * edtC - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) * rdt3;
        const double gammaC = 3.0 * (2.0 * dt - 6.0 + edtC * (6.0 + 4.0 * dt + dt2)) * rdt3;
        const double deltaC = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edtC * (3.0 + dt)) * rdt3;
        const double edtT = 1.0 - dt + 0.5 * dt2 - dt3 * (1.0 / 6.0);
        const double alphaT = 0.25 * dt - 0.2 * dt2 + dt3 * (1.0 / 12.0);
        const double betaT = 0.25 * dt - 0.05 * dt2 + dt3 * (1.0 / 120.0);
        const double gammaT = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
        const double deltaT = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
        const bool tay = dt < 5e-2;
        const double alpha = sel(tay, alphaT, alphaC), beta = sel(tay, betaT, betaC);
        const double gamma = sel(tay, gammaT, gammaC), delta = sel(tay, deltaT, deltaC);
        const double dt3rd = dt * (1.0 / 3.0);
        c.ed[j] = sel(tay, edtT, edtC);
        c.AG[j] = alpha + gamma;
        c.GD[j] = gamma * dt3rd;
        c.BD[j] = beta + delta;
        c.DD[j] = delta * dt3rd;
    }
}

// down-going ray (k ascending): point k is reached through the interval (k - 1, k)
template <int NCH>
__device__ __forceinline__ void ray_down(const RayCoef<NCH>& c, const double (&S)[NCH], const double (&DSf)[NCH],
                                         const double (&rchi)[NCH], const RayEnds& ends, int lane, int laneE, int jE,
                                         double (&a)[NCH], double (&b)[NCH], double (&psi)[NCH])
{
    double X[NCH], edP[NCH], XP[NCH], BDP[NCH], DDP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        X[j] = fma(c.AG[j], S[j], c.GD[j] * DSf[j]);
    shift_prev<NCH>(c.ed, edP);
    shift_prev<NCH>(X, XP);
    shift_prev<NCH>(c.BD, BDP);
    shift_prev<NCH>(c.DD, DDP);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        a[j] = edP[j];
        b[j] = fma(-DDP[j], DSf[j], fma(BDP[j], S[j], XP[j]));
        psi[j] = BDP[j] * rchi[j];
    }
    // first point: boundary intensity; last point: piecewise linear (ray_endpoints)
    if (lane == 0)
    {
        a[0] = 0.0;
        b[0] = ends.Iupw;
        psi[0] = 0.0;
    }
    if (lane == laneE)
    {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            if (j == jE)
            {
                a[j] = ends.aE;
                b[j] = ends.bE;
                psi[j] = ends.pE;
            }
    }
}

// up-going ray (k descending): point k is reached through the interval (k, k + 1); derivatives along the
// ray are the negatives of the forward ones
template <int NCH>
__device__ __forceinline__ void ray_up(const RayCoef<NCH>& c, const double (&S)[NCH], const double (&SN)[NCH],
                                       const double (&DSf)[NCH], const double (&rchi)[NCH], const RayEnds& ends,
                                       int lane, int laneE, int jE, double (&a)[NCH], double (&b)[NCH],
                                       double (&psi)[NCH])
{
    double DSfN[NCH];
    shift_next<NCH>(DSf, DSfN);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        a[j] = c.ed[j];
        b[j] = fma(c.DD[j], DSf[j], fma(c.BD[j], S[j], fma(-c.GD[j], DSfN[j], c.AG[j] * SN[j])));
        psi[j] = c.BD[j] * rchi[j];
    }
    // last point (the top): piecewise linear; first point (the bottom): boundary intensity; the padding
    // beyond the bottom comes first in this sweep and must carry the identity
    if (lane == 0)
    {
        a[0] = ends.aE;
        b[0] = ends.bE;
        psi[0] = ends.pE;
    }
    if (lane >= laneE)
    {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const bool first = lane == laneE && j == jE;
            const bool pad = lane > laneE || j > jE;
            if (first || pad)
            {
                a[j] = first ? 0.0 : 1.0;
                b[j] = first ? ends.Iupw : 0.0;
                psi[j] = 0.0;
            }
        }
    }
}

// bezier3_sweep: coefficients + scan (kept for callers that do not pipeline the two).
template <int NCH, bool DOWN, bool MULTI>
__device__ __forceinline__ void bezier3_sweep(DepthComm<MULTI>& cm, const GeometryR<NCH>& g, const double (&S)[NCH],
                                              const double (&rchi)[NCH], const RayPre<NCH>& r, const RayEnds& ends,
                                              double (&I)[NCH], double (&psi)[NCH])
{
    double a[NCH], b[NCH];
    bezier3_coeffs<NCH, DOWN>(cm, g, S, rchi, r, ends, a, b, psi);
    affine_scan<NCH, DOWN>(cm, a, b, I);
}

// linear (SOLVER 0) and besser (SOLVER 1): local stencils only
template <int NCH, int SOLVER, bool MULTI>
__device__ __forceinline__ void local_stencil_ray(DepthComm<MULTI>& cm, const GeometryR<NCH>& g, const double (&chi)[NCH],
                                                  const double (&S)[NCH], const double (&rchi)[NCH],
                                                  double muz, bool down, int bcType, double bcB0,
                                                  double bcB1, double bcValue, double (&I)[NCH],
                                                  double (&psi)[NCH])
{
    const int lane = cm.lane_global();
    const int K = g.K;
    const int ks = down ? 0 : K - 1;
    const int ke = down ? K - 1 : 0;
    double a[NCH], b[NCH];
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(cm, chi, chiN);
    shift_prev<NCH>(cm, chi, chiP);
    shift_next<NCH>(cm, S, SN);
    shift_prev<NCH>(cm, S, SP);
    const double zmu = (SOLVER == 0 ? 0.5 : 1.0) / muz;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const bool isEnd = (k == ke);
        const double chiUw = down ? chiP[j] : chiN[j];
        const double chiDw = down ? chiN[j] : chiP[j];
        const double Suw = down ? SP[j] : SN[j];
        const double Sdw = down ? SN[j] : SP[j];
        const double dsfUw = down ? g.dsfP[j] : g.dsf[j];
        const double dsfDw = down ? g.dsf[j] : g.dsfP[j];
        double aa, bb, pp;
        if (SOLVER == 0 || isEnd)
        {
            const double dt = (SOLVER == 0 ? zmu : 0.5 * zmu) * (chi[j] + chiUw) * dsfUw;
            const double rdt = 1.0 / dt;
            double w0, w1;
            if (dt < 5.0E-4)
            {
                w0 = dt * (1.0 - 0.5 * dt);
                w1 = (dt * dt) * (0.5 - dt * (1.0 / 3.0));
            }
            else if (dt > 50.0)
            {
                w0 = w1 = 1.0;
            }
            else
            {
                const double e = exp(-dt);
                w0 = 1.0 - e;
                w1 = w0 - dt * e;
            }
            aa = 1.0 - w0;
            if (SOLVER == 0)
            {
                bb = w0 * S[j] + w1 * ((Suw - S[j]) * rdt);
                pp = w0 - w1 * rdt;
            }
            else
            {
                bb = w0 * S[j] - w1 * ((S[j] - Suw) / dt);
                pp = w0 - w1 / dt;
            }
        }
        else
        {
            const double ds_uw = dsfUw * zmu, ds_dw = dsfDw * zmu;
            const double chiC = besser_control_point(ds_uw, ds_dw, chiUw, chi[j], chiDw);
            const double dtauUw = (1.0 / 3.0) * (chiUw + chiC + chi[j]) * ds_uw;
            const double dtauDw = 0.5 * (chi[j] + chiDw) * ds_dw;
            const double SC = besser_control_point(dtauUw, dtauDw, Suw, S[j], Sdw);
            const double t = dtauUw;
            double Mc, Oc, Cc, edt;
            if (t < 0.14)
            {
                Mc = (t * (t * (t * (t * (t * (t * ((140.0 - 18.0 * t) * t - 945.0) + 5400.0) - 25200.0) + 90720.0) - 226800.0) + 302400.0)) / 907200.0;
                Oc = (t * (t * (t * (t * (t * (t * ((10.0 - t) * t - 90.0) + 720.0) - 5040.0) + 30240.0) - 151200.0) + 604800.0)) / 1814400.0;
                Cc = (t * (t * (t * (t * (t * (t * ((35.0 - 4.0 * t) * t - 270.0) + 1800.0) - 10080.0) + 45360.0) - 151200.0) + 302400.0)) / 907200.0;
                const double t2 = t * t, t3 = t2 * t;
                edt = 1.0 - t + 0.5 * t2 - t3 / 6.0 + t * t3 / 24.0 - t2 * t3 / 120.0 + t3 * t3 / 720.0 - t3 * t3 * t / 5040.0;
            }
            else
            {
                const double t2 = t * t;
                edt = exp(-t);
                Mc = (2.0 - edt * (t2 + 2.0 * t + 2.0)) / t2;
                Oc = 1.0 - 2.0 * (edt + t - 1.0) / t2;
                Cc = 2.0 * (t - 2.0 + edt * (t + 2.0)) / t2;
            }
            aa = edt;
            bb = Mc * Suw + Oc * S[j] + Cc * SC;
            pp = Oc + Cc;
        }
        if (k == ks)
        {
            double Iupw = 0.0;
            if (bcType == 2)
            {
                const double dtau_b = (SOLVER == 0 ? zmu : 0.5 * zmu) * (chi[j] + chiDw) * dsfDw;
                Iupw = bcB0 - (bcB1 - bcB0) / dtau_b;
            }
            else if (bcType == 4)
                Iupw = bcValue;
            aa = 0.0;
            bb = Iupw;
            pp = 0.0;
        }
        if (k >= K)
        {
            aa = 1.0;
            bb = 0.0;
            pp = 0.0;
        }
        a[j] = aa;
        b[j] = bb;
        psi[j] = pp * rchi[j];
    }
    if (down)
        affine_scan<NCH, true>(cm, a, b, I);
    else
        affine_scan<NCH, false>(cm, a, b, I);
}

} // namespace lwb200
