// lwb200_fsm.cuh -- the production Gamma-iteration kernel ("moments" form).
//
// Same ownership as fs_kernel (one warp = one wavelength of one column, lanes
// over depth), but the Gamma / rate accumulation of
// compute_full_operator_rates (SimdFullIterationTemplates.hpp:206-234) is
// re-associated.  Every quantity that the reference sums over the rays of one
// wavelength is a polynomial in (I_r, Psi*_r, phi_r) with ray-independent
// coefficients:
//
//   Vij = v p,  Vji = g v p,  Uji = u g v p          (p = phi for a line, 1 for a continuum)
//   chi_atom(m) = sum_q p_q X_q(m),  U_atom(m) = sum_q p_q U_q(m),  eta_atom = sum_q p_q E_q
//
// so per ray only the MOMENTS  sum_r w_r {I, Psi*, p, p I, p Psi*, p p' Psi*}
// are accumulated (in registers, private to the lane that owns depth k), and
// the per-transition work (Gamma(i,j), Gamma(j,i), Rij, Rji) runs once per
// wavelength instead of once per ray.  The kernel is specialised on the number
// NL of lines overlapping at a wavelength (0, 1 or 2; the planner cuts tiles so
// that NL is constant inside a tile and routes the rare wavelengths with three
// or more overlapping lines to the general fs_kernel).
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
// 1/x to ~1 ulp: hardware seed + two Newton steps (no IEEE corner cases needed:
// every argument here is a finite positive opacity, path length or optical depth)
__device__ __forceinline__ double rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
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
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// one 64-bit value from the neighbouring lane: lane-1 if `fromPrev`, else lane+1
__device__ __forceinline__ double shfl_neighbour(double v, bool fromPrev)
{
    const int src = lane_id() + (fromPrev ? -1 : 1);
    return __shfl_sync(kFull, v, src & 31);
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

template <int NCH>
__device__ __forceinline__ void load_geometry_r(GeometryR<NCH>& g, const double* __restrict__ height, int K)
{
    const int lane = lane_id();
    g.K = K;
    double h[NCH], hN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        h[j] = (k < K) ? __ldg(height + k) : 0.0;
    }
    shift_next<NCH>(h, hN);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        g.dsf[j] = (k + 1 < K) ? fabs(h[j] - hN[j]) : 1.0;
    }
    shift_prev<NCH>(g.dsf, g.dsfP);
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
    // wUw = dsdw/(dsdw+dsuw) multiplies Suw, wDw = dsuw/(dsdw+dsuw) multiplies S0 (Bezier.hpp:58-65)
    const double P0 = fabs(fma(Suw, wUw, S0 * wDw));
    return (copysign(1.0, S0) + copysign(1.0, Suw)) * fmin(fabs(Suw), fmin(fabs(S0), 0.5 * P0));
}

__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

// One ray of piecewise_bezier3_1d (FormalScalar.cpp:209-325, :535-600), both
// directions and every optical-depth regime through one straight-line path.
// Outputs I and psi = Psi*/chi.
template <int NCH>
__device__ __forceinline__ void bezier3_ray(const GeometryR<NCH>& g, const double (&chi)[NCH],
                                            const double (&S)[NCH], const double (&rchi)[NCH],
                                            double muz, bool down, int bcType, double bcB0,
                                            double bcB1, double bcValue, double (&I)[NCH],
                                            double (&psi)[NCH])
{
    const int lane = lane_id();
    const int K = g.K;
    const int ks = down ? 0 : K - 1;
    const int ke = down ? K - 1 : 0;
    const double zmu = rcp_fast(muz);
    const double sgn = down ? 1.0 : -1.0;
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(chi, chiN);
    shift_prev<NCH>(chi, chiP);
    shift_next<NCH>(S, SN);
    shift_prev<NCH>(S, SP);

    // chi slopes on forward intervals, Steffen derivatives (array-forward orientation; the
    // derivative along an up-going ray is exactly the negative)
    double sl[NCH], Df[NCH], DfN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        sl[j] = (chiN[j] - chi[j]) * (g.rdsf[j] * muz);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const double slPj = (j == 0) ? (chi[0] - chiP[0]) * (g.rdsfP0 * muz) : sl[j > 0 ? j - 1 : 0];
        double d = steffen_r(g.dsf[j] * g.rsum[j], g.dsfP[j] * g.rsum[j], slPj, sl[j]);
        d = sel(k == 0, sl[j], d);       // one-sided at the top    (:239 / :288)
        d = sel(k == K - 1, slPj, d);    // one-sided at the bottom
        Df[j] = d;
    }
    shift_next<NCH>(Df, DfN);

    // Bezier-interpolated optical depth of the forward interval (k, k+1) (:242-246, :261-263)
    double dtf[NCH], dtfP[NCH], rdtf[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double ds = g.dsf[j] * zmu;
        const double ds3 = ds * (1.0 / 3.0);
        const double cA = fma(ds3, Df[j], chi[j]);
        const double cB = fma(-ds3, DfN[j], chiN[j]);
        const double t1 = chi[j] + chiN[j];
        dtf[j] = ds * ((t1 + sel(down, cA, cB)) + sel(down, cB, cA)) * 0.25;
        rdtf[j] = rcp_fast(dtf[j]);
    }
    shift_prev<NCH>(dtf, dtfP);

    // source-function slopes and derivatives with respect to optical depth
    double slS[NCH], DSf[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        slS[j] = (SN[j] - S[j]) * rdtf[j];
    const double rdtfP0 = rcp_fast(dtfP[0]);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const double slSPj = (j == 0) ? (S[0] - SP[0]) * rdtfP0 : slS[j > 0 ? j - 1 : 0];
        const double rs = rcp_fast(dtf[j] + dtfP[j]);
        double d = steffen_r(dtf[j] * rs, dtfP[j] * rs, slSPj, slS[j]);
        d = sel(k == 0, slS[j], d);      // (:247)
        d = sel(k == K - 1, slSPj, d);
        DSf[j] = d;
    }
    // derivative at the upwind point, signed along the ray: one shuffle whose source lane
    // depends on the direction
    const double DSedge = shfl_neighbour(sel(down, DSf[NCH - 1], DSf[0]), down);

    double a[NCH], b[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = lane * NCH + j;
        const bool isEnd = (k == ke);
        const double chiUw = sel(down, chiP[j], chiN[j]);
        const double Suw = sel(down, SP[j], SN[j]);
        const double dsfUw = sel(down, g.dsfP[j], g.dsf[j]);
        // upwind derivative along the ray
        const double DSprev = (j == 0) ? DSedge : DSf[j > 0 ? j - 1 : 0];
        const double DSnext = (j == NCH - 1) ? DSedge : DSf[j < NCH - 1 ? j + 1 : j];
        const double DSuw = sel(down, DSprev, -DSnext);
        // piecewise linear on the end (:307-321), Bezier elsewhere: one optical depth, one exp
        const double dtEnd = 0.5 * zmu * (chi[j] + chiUw) * dsfUw;
        const double dt = sel(isEnd, dtEnd, sel(down, dtfP[j], dtf[j]));
        const double rdt = rcp_fast(dt);
        const double ex = exp_fast(-fmin(fmax(dt, 0.0), 700.0));
        const double dt2 = dt * dt, dt3 = dt2 * dt;

        // Bezier3_coeffs (Bezier.hpp:81-127); dt > 30 is the closed form with edt = 0, term for term
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
        const double Cuw = fma(dt3rd, DSuw, Suw);
        const double C0 = fma(-sgn * dt3rd, DSf[j], S[j]);
        double aa = sel(tay, edtT, edtC);
        double bb = alpha * Suw + beta * S[j] + gamma * Cuw + delta * C0;
        double pp = beta + delta;

        // w2() (LwInternal.hpp:90-110) for the end point
        const bool tayE = dt < 5.0E-4, thickE = dt > 50.0;
        const double w0m = 1.0 - ex;
        const double w0 = sel(tayE, dt * (1.0 - 0.5 * dt), sel(thickE, 1.0, w0m));
        const double w1 = sel(tayE, dt2 * (0.5 - dt * (1.0 / 3.0)), sel(thickE, 1.0, w0m - dt * ex));
        const double dS = (S[j] - Suw) * rdt;
        aa = sel(isEnd, 1.0 - w0, aa);
        bb = sel(isEnd, w0 * S[j] - w1 * dS, bb);
        pp = sel(isEnd, w0 - w1 * rdt, pp);

        // boundary intensity (:551-597)
        double Iupw = bcValue;
        if (bcType == 2)
        {
            const double chiDw = sel(down, chiN[j], chiP[j]);
            const double dsfDw = sel(down, g.dsf[j], g.dsfP[j]);
            const double dtau_b = 0.5 * zmu * (chi[j] + chiDw) * dsfDw;
            Iupw = bcB0 - (bcB1 - bcB0) * rcp_fast(dtau_b);
        }
        const bool isStart = (k == ks);
        aa = sel(isStart, 0.0, aa);
        bb = sel(isStart, Iupw, bb);
        pp = sel(isStart, 0.0, pp);
        const bool valid = k < K;
        a[j] = sel(valid, aa, 1.0);
        b[j] = sel(valid, bb, 0.0);
        psi[j] = sel(valid, pp * rchi[j], 0.0);
    }
    if (down)
        affine_scan<NCH, true>(a, b, I);
    else
        affine_scan<NCH, false>(a, b, I);
}

// linear (SOLVER 0) and besser (SOLVER 1): local stencils only
template <int NCH, int SOLVER>
__device__ __forceinline__ void local_stencil_ray(const GeometryR<NCH>& g, const double (&chi)[NCH],
                                                  const double (&S)[NCH], const double (&rchi)[NCH],
                                                  double muz, bool down, int bcType, double bcB0,
                                                  double bcB1, double bcValue, double (&I)[NCH],
                                                  double (&psi)[NCH])
{
    const int lane = lane_id();
    const int K = g.K;
    const int ks = down ? 0 : K - 1;
    const int ke = down ? K - 1 : 0;
    double a[NCH], b[NCH];
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(chi, chiN);
    shift_prev<NCH>(chi, chiP);
    shift_next<NCH>(S, SN);
    shift_prev<NCH>(S, SP);
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
    __syncwarp();
    if (down)
        affine_scan<NCH, true>(a, b, I);
    else
        affine_scan<NCH, false>(a, b, I);
}

// per-wavelength line slot
struct LineSlot
{
    int trans;   // index into P.trans
    int atom, li, lj;
    double v;    // hnu/4pi * Bij
    double gv;   // g * v   (without rhoPrd)
    double ugv;  // Aji/Bji * g * v
    double wlaS; // wlambda * 4 pi / (h c)   (times wphi(k) gives wla)
    const double* phi;  // phi(lt, 0, 0, 0) of this column
    const double* rho;  // rhoPrd(lt, 0) or nullptr
    const double* wphi; // wphi(0) of this column
};

#ifndef LWB200_FSM_MINBLOCKS
#define LWB200_FSM_MINBLOCKS 2
#endif

template <int NCH, int SOLVER, int NL>
__global__ void __launch_bounds__(128, LWB200_FSM_MINBLOCKS)
fsm_kernel(const DevProblem P, const int* __restrict__ tileList, int laLo, int laHi, int lambdaIterate,
           int storeDepth)
{
    extern __shared__ double smem[];
    constexpr int NLA = NL > 0 ? NL : 1;
    const int K = P.K, M = P.M, L = P.L, KP = P.KP;
    const int tile = tileList[blockIdx.x];
    const int col = blockIdx.y;
    // warp index made provably warp-uniform: every loop below stays convergent
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);
    const int nwarp = blockDim.x >> 5;
    const int lane = lane_id();

    const int slot0 = P.tileSlotOff[tile];
    const int nslot = P.tileSlotOff[tile + 1] - slot0;
    double* acc = smem;
    double* Xs = smem + (size_t)P.maxSlots * 4 * KP + (size_t)warp * 2 * P.maxNlevel * 32;
    double* Us = Xs + P.maxNlevel * 32;

    for (int idx = threadIdx.x; idx < nslot * 4 * KP; idx += blockDim.x)
        acc[idx] = 0.0;
    __syncthreads();

    GeometryR<NCH> g;
    load_geometry_r<NCH>(g, P.height + (size_t)col * K, K);
    const double* Tcol = P.temperature + (size_t)col * K;
    const double* ncol = P.n + (size_t)col * P.NlevTot * K;

    const int tlBeg = P.tileLa[tile], tlEnd = P.tileLa[tile + 1];

    for (int tl = tlBeg + warp; tl < tlEnd; tl += nwarp)
    {
        const int la = P.tileLambda[tl];
        if (la < laLo || la >= laHi)
            continue;
        const double lambda = __ldg(P.wavelength + la);
        const double rlambda = 1.0 / lambda;
        const size_t rowLK = ((size_t)col * L + la) * K;
        const int eBeg = P.laOff[la], eEnd = eBeg + P.laCnt[la];
        constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
        constexpr double twoHc = 2.0 * kHC / (kNmToM * kNmToM * kNmToM);
        constexpr double hc_4pi = 0.25 * kHC / kPi;
        constexpr double pi4_h = 4.0 * kPi / kHPlanck;
        constexpr double pi4_hc = 1.0 / hc_4pi;
        const double hc_kl = hc_k * rlambda;
        const double hcl = twoHc * (rlambda * rlambda * rlambda);

        // ---- ray-independent: background + continua (generalises the reference's
        //      continuaOnly shortcut, :295-307)
        double chiC[NCH], etaC[NCH], scaJ[NCH], expfac[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            const bool v = k < K;
            chiC[j] = v ? __ldg(P.chiBg + rowLK + k) : 1.0;
            etaC[j] = v ? __ldg(P.etaBg + rowLK + k) : 0.0;
            const double sca = v ? __ldg(P.scaBg + rowLK + k) : 0.0;
            const double JDag = v ? P.J[rowLK + k] : 0.0;
            scaJ[j] = sca * JDag;
            const double Tk = v ? __ldg(Tcol + k) : 1.0e4;
            expfac[j] = exp_fast(-hc_kl / Tk);
        }
        LineSlot ls[NLA];
        double cX[NLA][NCH], cE[NLA][NCH];
#pragma unroll
        for (int l = 0; l < NLA; ++l)
        {
            ls[l].trans = -1;
            ls[l].atom = -1;
            ls[l].li = ls[l].lj = 0;
            ls[l].v = ls[l].gv = ls[l].ugv = ls[l].wlaS = 0.0;
            ls[l].phi = P.phi;
            ls[l].rho = nullptr;
            ls[l].wphi = P.wphi;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                cX[l][j] = cE[l][j] = 0.0;
        }
        int nLseen = 0;

        for (int e = eBeg; e < eEnd; ++e)
        {
            const int ti = P.entries[e].trans;
            const DevTrans& t = P.trans[ti];
            const int lt = la - t.Nblue;
            if (t.type == 0)
            {
                if (NL > 0)
                {
                    // line: constants of Transition::uv (LwTransition.hpp:93-130)
                    const double vB = hc_4pi * (t.lambda0 * rlambda) * t.Bij;
                    const double gS = t.Bji_Bij;
                    const double* rho = (t.rhoOff >= 0)
                        ? P.rhoPrd + t.rhoOff + ((size_t)col * (t.Nred - t.Nblue) + lt) * K : nullptr;
                    const int l = nLseen++;
#pragma unroll
                    for (int q = 0; q < NLA; ++q)
                    {
                        if (q == l)
                        {
                            ls[q].trans = ti;
                            ls[q].atom = t.atom;
                            ls[q].li = t.i;
                            ls[q].lj = t.j;
                            ls[q].v = vB;
                            ls[q].gv = gS * vB;
                            ls[q].ugv = t.Aji_Bji * (gS * vB);
                            ls[q].wlaS = __ldg(P.wlambdaTab + t.tabOff + lt) * pi4_hc;
                            ls[q].phi = P.phi + t.phiOff + (size_t)col * t.phiColStride + (size_t)lt * M * 2 * K;
                            ls[q].rho = rho;
                            ls[q].wphi = P.wphi + ((size_t)t.lineIdx * P.Ncol + col) * K;
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
                            {
                                const int k = lane * NCH + j;
                                if (k < K)
                                {
                                    const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                                    const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                                    const double gk = rho ? gS * __ldg(rho + k) : gS;
                                    cX[q][j] = vB * (ni - nj * gk);
                                    cE[q][j] = nj * (t.Aji_Bji * (gk * vB));
                                }
                            }
                        }
                    }
                }
            }
            else
            {
                const double al = __ldg(P.alphaTab + t.tabOff + lt);
                const double* gr = P.gRatio + ((size_t)t.contIdx * P.Ncol + col) * K;
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = lane * NCH + j;
                    if (k < K)
                    {
                        const double gk = __ldg(gr + k) * expfac[j];
                        const double Vji = gk * al;
                        const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                        const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                        chiC[j] += ni * al - nj * Vji;
                        etaC[j] += nj * (hcl * Vji);
                    }
                }
            }
        }

        // thermalised boundaries: Planck function at the two boundary pairs (once per wavelength)
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

        // ---- moments over the rays of this wavelength
        //   mJ = sum w I, mP = sum w Psi*, mW[l] = sum w p_l, mA[l] = sum w p_l I,
        //   mB0[l] = sum w Psi* p_l, mB[l] = sum w Psi* p_l^2, mBx[(a,b)] = sum w Psi* p_a p_b (a < b)
        constexpr int NPAIR = NL > 1 ? NL * (NL - 1) / 2 : 1;
        double mJ[NCH], mP[NCH], mW[NLA][NCH], mA[NLA][NCH], mB0[NLA][NCH], mB[NLA][NCH], mBx[NPAIR][NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            mJ[j] = mP[j] = 0.0;
#pragma unroll
            for (int l = 0; l < NLA; ++l)
                mW[l][j] = mA[l][j] = mB0[l][j] = mB[l][j] = 0.0;
#pragma unroll
            for (int q = 0; q < NPAIR; ++q)
                mBx[q][j] = 0.0;
        }
        double W0 = 0.0;
        double chi[NCH], S[NCH], rchi[NCH];
        const double* ph[NLA];
#pragma unroll
        for (int l = 0; l < NLA; ++l)
            ph[l] = ls[l].phi + lane * NCH;

#ifdef LWB200_EXP_NRAYS
        for (int ray = 0; ray < LWB200_EXP_NRAYS; ++ray)
#else
        for (int ray = 0; ray < 2 * M; ++ray)
#endif
        {
            const int mu = ray >> 1, dir = ray & 1;
            const double muz = __ldg(P.muz + mu);
            const double w = 0.5 * __ldg(P.wmu + mu);
            double p[NLA][NCH];
            if (NL > 0 || ray == 0)
            {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = lane * NCH + j;
                    double c = chiC[j], e = etaC[j];
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        p[l][j] = 0.0;
                        if (NL > 0)
                        {
                            p[l][j] = (k < K) ? __ldg(ph[l] + j) : 0.0;
                            c = fma(cX[l][j], p[l][j], c);
                            e = fma(cE[l][j], p[l][j], e);
                        }
                    }
                    chi[j] = c;
                    rchi[j] = rcp_fast(c);
                    S[j] = (e + scaJ[j]) * rchi[j]; // compute_source_fn (:169-179)
                    if (storeDepth && k < K)
                    {
                        const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k;
                        P.depthChi[off] = c;
                        P.depthEta[off] = e;
                    }
                }
#pragma unroll
                for (int l = 0; l < NLA; ++l)
                    ph[l] += K;
            }
            else if (storeDepth)
            {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                {
                    const int k = lane * NCH + j;
                    if (k < K)
                    {
                        const size_t off = ((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k;
                        P.depthChi[off] = chiC[j];
                        P.depthEta[off] = etaC[j];
                    }
                }
            }
            int bcType;
            double bcB0, bcB1, bcValue = 0.0;
            if (dir == 1)
            {
                bcType = P.lowerBc;
                bcB0 = Bbot0;
                bcB1 = Bbot1;
                if (bcType == 4)
                    bcValue = P.lowerBcData[((size_t)col * L + la) * P.NlowerBcMu + P.lowerBcIdx[mu * 2 + 1]];
            }
            else
            {
                bcType = P.upperBc;
                bcB0 = Btop0;
                bcB1 = Btop1;
                if (bcType == 4)
                    bcValue = P.upperBcData[((size_t)col * L + la) * P.NupperBcMu + P.upperBcIdx[mu * 2 + 0]];
            }
            double I[NCH], psi[NCH];
            if (SOLVER == 2)
                bezier3_ray<NCH>(g, chi, S, rchi, muz, dir == 0, bcType, bcB0, bcB1, bcValue, I, psi);
            else
                local_stencil_ray<NCH, SOLVER>(g, chi, S, rchi, muz, dir == 0, bcType, bcB0, bcB1, bcValue, I, psi);

            if (lane == 0)
                P.I[((size_t)col * L + la) * M + mu] = I[0];
            W0 += w;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = lane * NCH + j;
                if (storeDepth && k < K)
                    P.depthI[((((size_t)col * L + la) * M + mu) * 2 + dir) * K + k] = I[j];
                const double wI = w * I[j];
                const double wP = lambdaIterate ? 0.0 : w * psi[j];
                mJ[j] += wI;
                mP[j] += wP;
                if (NL > 0)
                {
                    double tq[NLA];
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        tq[l] = wP * p[l][j];
                        mW[l][j] = fma(w, p[l][j], mW[l][j]);
                        mA[l][j] = fma(wI, p[l][j], mA[l][j]);
                        mB0[l][j] += tq[l];
                        mB[l][j] = fma(tq[l], p[l][j], mB[l][j]);
                    }
                    if (NL > 1)
                    {
                        int pr = 0;
#pragma unroll
                        for (int a = 0; a < NLA; ++a)
#pragma unroll
                            for (int b = a + 1; b < NLA; ++b)
                            {
                                mBx[pr][j] = fma(tq[a], p[b][j], mBx[pr][j]);
                                ++pr;
                            }
                    }
                }
            }
        }

        // ---- J row and dJ (:477-485)
        double dJ = 0.0;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            const int k = lane * NCH + j;
            if (k < K)
            {
                const double JDag = P.J[rowLK + k];
                P.J[rowLK + k] = mJ[j];
                const double d = fabs(1.0 - JDag / mJ[j]);
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

        // ---- epilogue: Gamma and rates from the moments, atom by atom.
        // Profile members of an atom: q = 0 the continua (p = 1), q = l + 1 line slot l.
        // M(q, q') = sum_r w Psi* p_q p_q'.
        int e0 = eBeg;
#ifdef LWB200_EXP_NOEPI
        e0 = eEnd;
#endif
        while (e0 < eEnd)
        {
            const int atom = P.trans[P.entries[e0].trans].atom;
            int e1 = e0 + 1;
            while (e1 < eEnd && P.trans[P.entries[e1].trans].atom == atom)
                ++e1;
            const bool detailed = P.atomDetailed[atom] != 0;
            const int N = P.atomNlevel[atom];
            bool own[NLA];
#pragma unroll
            for (int l = 0; l < NLA; ++l)
                own[l] = (NL > 0) && ls[l].atom == atom;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
            {
                const int k = lane * NCH + j;
                if (k < K)
                {
                    constexpr int NQ = NL + 1;
                    double Mq[NQ][NQ], Eq[NQ];
                    Mq[0][0] = mP[j];
                    {
                        int pr = 0;
#pragma unroll
                        for (int a = 0; a < NL; ++a)
                        {
                            Mq[0][a + 1] = Mq[a + 1][0] = mB0[a][j];
                            Mq[a + 1][a + 1] = mB[a][j];
#pragma unroll
                            for (int b = a + 1; b < NL; ++b)
                            {
                                Mq[a + 1][b + 1] = Mq[b + 1][a + 1] = mBx[pr][j];
                                ++pr;
                            }
                        }
                    }
                    double E0 = 0.0;
                    if (!detailed)
                    {
                        // continuum aggregates per level: chi_atom / U_atom of chi_eta_aux_accum (:59-109)
                        for (int m = 0; m < N; ++m)
                        {
                            Xs[m * 32 + lane] = 0.0;
                            Us[m * 32 + lane] = 0.0;
                        }
                        for (int e = e0; e < e1; ++e)
                        {
                            const DevTrans& t = P.trans[P.entries[e].trans];
                            if (t.type == 0)
                                continue;
                            const double al = __ldg(P.alphaTab + t.tabOff + (la - t.Nblue));
                            const double gk = __ldg(P.gRatio + ((size_t)t.contIdx * P.Ncol + col) * K + k) * expfac[j];
                            const double Vji = gk * al;
                            const double Uji = hcl * Vji;
                            const double ni = __ldg(ncol + (size_t)t.levI * K + k);
                            const double nj = __ldg(ncol + (size_t)t.levJ * K + k);
                            const double x = ni * al - nj * Vji;
                            Xs[t.i * 32 + lane] += x;
                            Xs[t.j * 32 + lane] -= x;
                            Us[t.j * 32 + lane] += Uji;
                            E0 += nj * Uji;
                        }
                    }
                    // line members of this atom: per-unit-phi coefficients (0 for other atoms' lines)
                    double Xl[NLA], gvl[NLA], ugvl[NLA];
                    Eq[0] = E0;
#pragma unroll
                    for (int l = 0; l < NLA; ++l)
                    {
                        Xl[l] = gvl[l] = ugvl[l] = 0.0;
                        if (NL > 0)
                        {
                            Eq[l + (NL > 0 ? 1 : 0)] = 0.0;
                            if (own[l])
                            {
                                const double r = ls[l].rho ? __ldg(ls[l].rho + k) : 1.0;
                                gvl[l] = ls[l].gv * r;
                                ugvl[l] = ls[l].ugv * r;
                                Xl[l] = cX[l][j];
                                Eq[l + (NL > 0 ? 1 : 0)] = cE[l][j];
                            }
                        }
                    }
                    // EB[q] = sum_q' E_q' M(q, q')
                    double EB[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        double s = 0.0;
#pragma unroll
                        for (int q2 = 0; q2 < NQ; ++q2)
                            s = fma(Eq[q2], Mq[q][q2], s);
                        EB[q] = s;
                    }

                    for (int e = e0; e < e1; ++e)
                    {
                        const DevEntry en = P.entries[e];
                        const DevTrans& t = P.trans[en.trans];
                        const int lt = la - t.Nblue;
                        double v = 0.0, gv = 0.0, ugv = 0.0, Wq = W0, Aq = mJ[j], EBq = EB[0], wla = 0.0;
                        if (t.type != 0)
                        {
                            const double al = __ldg(P.alphaTab + t.tabOff + lt);
                            const double gk = __ldg(P.gRatio + ((size_t)t.contIdx * P.Ncol + col) * K + k) * expfac[j];
                            v = al;
                            gv = gk * al;
                            ugv = hcl * gv;
                            wla = (__ldg(P.wlambdaTab + t.tabOff + lt) * rlambda) * pi4_h;
                        }
#pragma unroll
                        for (int l = 0; l < NL; ++l)
                        {
                            if (t.type == 0 && en.trans == ls[l].trans)
                            {
                                v = ls[l].v;
                                gv = gvl[l];
                                ugv = ugvl[l];
                                Wq = mW[l][j];
                                Aq = mA[l][j];
                                EBq = EB[l + 1];
                                wla = ls[l].wlaS * __ldg(ls[l].wphi + k);
                            }
                        }
                        double* a4 = acc + (size_t)en.slot * 4 * KP + k;
                        if (!detailed)
                        {
                            // chi_atom(m) = sum_q p_q X_q(m), U_atom(m) = sum_q p_q U_q(m)
                            double Xi[NQ], Xj[NQ], Ui[NQ], Uj[NQ];
                            Xi[0] = Xs[t.i * 32 + lane];
                            Xj[0] = Xs[t.j * 32 + lane];
                            Ui[0] = Us[t.i * 32 + lane];
                            Uj[0] = Us[t.j * 32 + lane];
#pragma unroll
                            for (int l = 0; l < NL; ++l)
                            {
                                Xi[l + 1] = t.i == ls[l].li ? Xl[l] : (t.i == ls[l].lj ? -Xl[l] : 0.0);
                                Xj[l + 1] = t.j == ls[l].li ? Xl[l] : (t.j == ls[l].lj ? -Xl[l] : 0.0);
                                Ui[l + 1] = t.i == ls[l].lj ? ugvl[l] : 0.0;
                                Uj[l + 1] = t.j == ls[l].lj ? ugvl[l] : 0.0;
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
                                    mj = fma(Mq[q][q2], Uj[q2], mj);
                                    mi = fma(Mq[q][q2], Ui[q2], mi);
                                }
                                XUij = fma(Xi[q], mj, XUij);
                                XUji = fma(Xj[q], mi, XUji);
                            }
                            // sum_r w [(Uji + Vji Ieff) - Psi* chi(i) U(j)],  Ieff = I - Psi* eta_atom
                            smem_add(a4, (ugv * Wq + gv * (Aq - EBq) - XUij) * wla);
                            smem_add(a4 + KP, (v * (Aq - EBq) - XUji) * wla);
                        }
                        smem_add(a4 + 2 * KP, (v * Aq) * wla);
                        smem_add(a4 + 3 * KP, (ugv * Wq + gv * Aq) * wla);
                    }
                }
            }
            e0 = e1;
        }
    }

    __syncthreads();
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

} // namespace lwb200
