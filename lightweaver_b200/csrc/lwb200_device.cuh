// lwb200_device.cuh -- device-side building blocks of the B200 formal solver.
//
// Execution model: ONE WARP OWNS ONE WAVELENGTH of one
// column and sweeps its 2*Nrays rays one after the other; the 32 lanes are laid
// over DEPTH.  Lane l holds the NCH consecutive depth points k = l*NCH + j
// (j < NCH, "blocked" layout) in registers, so that
//   * every array of the reference (phi, chi/eta background, J, n, Gamma, R --
//     all k-innermost, Source/CmoArray.hpp) is read/written in its native
//     layout with no transposition;
//   * everything that the reference sums over rays and wavelengths at fixed
//     depth (J, Gamma, Rij/Rji) is a private per-lane accumulation -- no
//     cross-lane reduction;
//   * all short-characteristics coefficients (which depend only on chi, S at
//     k-2..k+2) are evaluated depth-parallel, and the only serial part,
//     I_k = a_k I_{k-dk} + b_k, becomes an affine warp scan.
//
// The formulas restate Source/FormalScalar.cpp:136-467, Source/Bezier.hpp:58-127
// and Source/LwInternal.hpp:90-110; operation order is kept where it is free.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lwb200
{
constexpr unsigned kFull = 0xffffffffu;

// Source/Constants.hpp:6-47
constexpr double kCLight = 2.99792458E+08;
constexpr double kHPlanck = 6.6260755E-34;
constexpr double kHC = kHPlanck * kCLight;
constexpr double kKBoltzmann = 1.380658E-23;
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kNmToM = 1.0E-09;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------
// Neighbour access in the blocked layout.  out[j] = value at depth k-1 / k+1.
// Lane boundaries cost one 64-bit shuffle per array, not one per element.
template <int NCH>
__device__ __forceinline__ void shift_prev(const double (&v)[NCH], double (&out)[NCH])
{
    double up = __shfl_up_sync(kFull, v[NCH - 1], 1);
    out[0] = up;
#pragma unroll
    for (int j = 1; j < NCH; ++j)
        out[j] = v[j - 1];
}

template <int NCH>
__device__ __forceinline__ void shift_next(const double (&v)[NCH], double (&out)[NCH])
{
    double dn = __shfl_down_sync(kFull, v[0], 1);
#pragma unroll
    for (int j = 0; j < NCH - 1; ++j)
        out[j] = v[j + 1];
    out[NCH - 1] = dn;
}

// ---------------------------------------------------------------------------
// Affine scan: I_k = a_k * I_{k-dk} + b_k along depth, I "before" the first
// point = 0 (the boundary point carries a = 0, b = I_upw).  DOWN: k ascending
// (toObs == false), else k descending.  Elements past the end of the
// atmosphere must carry the identity (a = 1, b = 0).
template <int NCH, bool DOWN>
__device__ __forceinline__ void affine_scan(const double (&a)[NCH], const double (&b)[NCH],
                                            double (&I)[NCH])
{
    const int lane = lane_id();
    double A, B;
    if (DOWN)
    {
        A = a[0];
        B = b[0];
#pragma unroll
        for (int j = 1; j < NCH; ++j)
        {
            B = fma(a[j], B, b[j]);
            A = a[j] * A;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            double Ap = __shfl_up_sync(kFull, A, d);
            double Bp = __shfl_up_sync(kFull, B, d);
            if (lane >= d)
            {
                B = fma(A, Bp, B);
                A = A * Ap;
            }
        }
        double x = __shfl_up_sync(kFull, B, 1);
        if (lane == 0)
            x = 0.0;
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            x = fma(a[j], x, b[j]);
            I[j] = x;
        }
    }
    else
    {
        A = a[NCH - 1];
        B = b[NCH - 1];
#pragma unroll
        for (int j = NCH - 2; j >= 0; --j)
        {
            B = fma(a[j], B, b[j]);
            A = a[j] * A;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            double Ap = __shfl_down_sync(kFull, A, d);
            double Bp = __shfl_down_sync(kFull, B, d);
            if (lane + d < 32)
            {
                B = fma(A, Bp, B);
                A = A * Ap;
            }
        }
        double x = __shfl_down_sync(kFull, B, 1);
        if (lane == 31)
            x = 0.0;
#pragma unroll
        for (int j = NCH - 1; j >= 0; --j)
        {
            x = fma(a[j], x, b[j]);
            I[j] = x;
        }
    }
}

// ---------------------------------------------------------------------------
// Depth ownership beyond one warp.  MULTI = false: one warp covers the whole column
// (Nspace <= 32 * NCH) and everything below is plain warp shuffles.  MULTI = true: the
// warps of a CTA cover consecutive blocks of 32 * NCH depths of ONE wavelength; a lane's
// depth neighbour may live in the adjacent warp, so edge values and the scan carry cross
// warps through a few doubles of shared memory (double-buffered: one barrier per exchange).
// Every warp of the CTA must make the same sequence of calls.
template <bool MULTI>
struct DepthComm
{
    double* buf;  // MULTI: [2][nwarp] edge exchange, then [2][nwarp][2] scan composites, then [nwarp]
    int warp;     // position of this warp along depth
    int nwarp;
    int parity;

    __device__ __forceinline__ int lane_global() const { return MULTI ? warp * 32 + lane_id() : lane_id(); }

    // `v` of the previous lane along depth (the first lane of the column gets its own value back)
    __device__ __forceinline__ double from_prev(double v)
    {
        double r = __shfl_up_sync(kFull, v, 1);
        if (MULTI)
        {
            const int lane = lane_id();
            if (lane == 31)
                buf[parity * nwarp + warp] = v;
            __syncthreads();
            if (lane == 0 && warp > 0)
                r = buf[parity * nwarp + warp - 1];
            parity ^= 1;
        }
        return r;
    }
    // `v` of the next lane along depth (the last lane gets its own value back)
    __device__ __forceinline__ double from_next(double v)
    {
        double r = __shfl_down_sync(kFull, v, 1);
        if (MULTI)
        {
            const int lane = lane_id();
            if (lane == 0)
                buf[parity * nwarp + warp] = v;
            __syncthreads();
            if (lane == 31 && warp < nwarp - 1)
                r = buf[parity * nwarp + warp + 1];
            parity ^= 1;
        }
        return r;
    }
    // max over all lanes of all warps, valid in every thread
    __device__ __forceinline__ double max_all(double v)
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
        {
            const double o = __shfl_xor_sync(kFull, v, d);
            v = (o < v) ? v : o;
        }
        if (MULTI)
        {
            double* mx = buf + 6 * nwarp;
            if (lane_id() == 0)
                mx[warp] = v;
            __syncthreads();
            for (int w = 0; w < nwarp; ++w)
                v = (mx[w] < v) ? v : mx[w];
            __syncthreads();
        }
        return v;
    }
};

template <int NCH, bool MULTI>
__device__ __forceinline__ void shift_prev(DepthComm<MULTI>& cm, const double (&v)[NCH], double (&out)[NCH])
{
    const double up = cm.from_prev(v[NCH - 1]);
    out[0] = up;
#pragma unroll
    for (int j = 1; j < NCH; ++j)
        out[j] = v[j - 1];
}

template <int NCH, bool MULTI>
__device__ __forceinline__ void shift_next(DepthComm<MULTI>& cm, const double (&v)[NCH], double (&out)[NCH])
{
    const double dn = cm.from_next(v[0]);
#pragma unroll
    for (int j = 0; j < NCH - 1; ++j)
        out[j] = v[j + 1];
    out[NCH - 1] = dn;
}

// affine_scan across the warps of a CTA: in-warp scan as above, then the composite of every
// warp goes through shared memory and each warp folds the ones before it into its carry.
template <int NCH, bool DOWN, bool MULTI>
__device__ __forceinline__ void affine_scan(DepthComm<MULTI>& cm, const double (&a)[NCH], const double (&b)[NCH],
                                            double (&I)[NCH])
{
    if (!MULTI)
    {
        affine_scan<NCH, DOWN>(a, b, I);
        return;
    }
    const int lane = lane_id();
    double A, B;
    if (DOWN)
    {
        A = a[0];
        B = b[0];
#pragma unroll
        for (int j = 1; j < NCH; ++j)
        {
            B = fma(a[j], B, b[j]);
            A = a[j] * A;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const double Ap = __shfl_up_sync(kFull, A, d);
            const double Bp = __shfl_up_sync(kFull, B, d);
            if (lane >= d)
            {
                B = fma(A, Bp, B);
                A = A * Ap;
            }
        }
    }
    else
    {
        A = a[NCH - 1];
        B = b[NCH - 1];
#pragma unroll
        for (int j = NCH - 2; j >= 0; --j)
        {
            B = fma(a[j], B, b[j]);
            A = a[j] * A;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const double Ap = __shfl_down_sync(kFull, A, d);
            const double Bp = __shfl_down_sync(kFull, B, d);
            if (lane + d < 32)
            {
                B = fma(A, Bp, B);
                A = A * Ap;
            }
        }
    }
    // warp composites -> carry into this warp
    double* sc = cm.buf + 2 * cm.nwarp + cm.parity * 2 * cm.nwarp;
    if (lane == (DOWN ? 31 : 0))
    {
        sc[2 * cm.warp] = A;
        sc[2 * cm.warp + 1] = B;
    }
    __syncthreads();
    double c = 0.0;
    if (DOWN)
    {
        for (int w = 0; w < cm.warp; ++w)
            c = fma(sc[2 * w], c, sc[2 * w + 1]);
    }
    else
    {
        for (int w = cm.nwarp - 1; w > cm.warp; --w)
            c = fma(sc[2 * w], c, sc[2 * w + 1]);
    }
    cm.parity ^= 1;
    // intensity entering this lane: composite of the lanes before it applied to the carry
    const double Ap = DOWN ? __shfl_up_sync(kFull, A, 1) : __shfl_down_sync(kFull, A, 1);
    const double Bp = DOWN ? __shfl_up_sync(kFull, B, 1) : __shfl_down_sync(kFull, B, 1);
    double x = (lane == (DOWN ? 0 : 31)) ? c : fma(Ap, c, Bp);
    if (DOWN)
    {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
        {
            x = fma(a[j], x, b[j]);
            I[j] = x;
        }
    }
    else
    {
#pragma unroll
        for (int j = NCH - 1; j >= 0; --j)
        {
            x = fma(a[j], x, b[j]);
            I[j] = x;
        }
    }
}

// ---------------------------------------------------------------------------
// Steffen (1990) derivative from the two adjacent slopes (Bezier.hpp:58-65).
// Written for the array-forward ("DOWN") orientation; the derivative along an
// up-going ray is exactly its negative (all products commute bitwise).
__device__ __forceinline__ double steffen(double dsuw, double dsdw, double Suw, double S0)
{
    const double P0 = fabs((Suw * dsdw + S0 * dsuw) / (dsdw + dsuw));
    return (copysign(1.0, S0) + copysign(1.0, Suw)) * fmin(fabs(Suw), fmin(fabs(S0), 0.5 * P0));
}

// planck_nu for one temperature (LwMisc.hpp:29-46)
__device__ __forceinline__ double planck_nu(double T, double lambda)
{
    constexpr double hc_k = kHC / (kKBoltzmann * kNmToM);
    constexpr double twoh_c2 = (2.0 * kHC) / (kNmToM * kNmToM * kNmToM);
    const double hc_kla = hc_k / lambda;
    const double twohnu3_c2 = twoh_c2 / (lambda * lambda * lambda);
    const double x = hc_kla / T;
    return (x <= 150.0) ? twohnu3_c2 / (exp(x) - 1.0) : 0.0;
}

// Per-warp, per-column depth geometry held in registers.
template <int NCH>
struct Geometry
{
    int K;
    double dsf[NCH];  // |h_k - h_{k+1}|  (forward interval of point k; 0 if k+1 >= K)
    double dsfP[NCH]; // |h_{k-1} - h_k|  (0 if k == 0)
    __device__ __forceinline__ int k(int j) const { return lane_id() * NCH + j; }
};

template <int NCH>
__device__ __forceinline__ void load_geometry(Geometry<NCH>& g, const double* __restrict__ height, int K)
{
    g.K = K;
    double h[NCH], hN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        int k = g.k(j);
        h[j] = (k < K) ? __ldg(height + k) : 0.0;
    }
    shift_next<NCH>(h, hN);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        int k = g.k(j);
        g.dsf[j] = (k + 1 < K) ? fabs(h[j] - hN[j]) : 0.0;
    }
    shift_prev<NCH>(g.dsf, g.dsfP);
    if (lane_id() == 0)
        g.dsfP[0] = 0.0;
}

// Boundary condition description for one ray.
struct RayBc
{
    int type;     // LWB200_BC_* of the boundary the ray starts from
    double B0;    // THERMALISED: Planck function at the boundary point ...
    double B1;    // ... and at its inner neighbour
    double value; // CALLABLE: boundary intensity
};

// ---------------------------------------------------------------------------
// piecewise_bezier3_1d (FormalScalar.cpp:209-325 + boundary :535-600), all
// depth points at once.  On exit a/b are the affine recurrence coefficients and
// psi = Psi* / chi.  WANT_PSI mirrors `computeOperator`.
template <int NCH, bool DOWN, bool WANT_PSI>
__device__ __forceinline__ void bezier3_coefficients(const Geometry<NCH>& g,
                                                     const double (&chi)[NCH],
                                                     const double (&S)[NCH], double zmu,
                                                     const RayBc& bc, double (&a)[NCH],
                                                     double (&b)[NCH], double (&psi)[NCH])
{
    const int K = g.K;
    const int ks = DOWN ? 0 : K - 1;
    const int ke = DOWN ? K - 1 : 0;
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(chi, chiN);
    shift_prev<NCH>(chi, chiP);
    shift_next<NCH>(S, SN);
    shift_prev<NCH>(S, SP);

    double ds[NCH], dsP[NCH], sl[NCH], slP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        ds[j] = g.dsf[j] * zmu;
        dsP[j] = g.dsfP[j] * zmu;
        sl[j] = (chiN[j] - chi[j]) / ds[j]; // garbage where k+1 >= K, never selected
    }
    shift_prev<NCH>(sl, slP);

    double Df[NCH], DfN[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        double d = steffen(dsP[j], ds[j], slP[j], sl[j]);
        d = (k == 0) ? sl[j] : d;      // one-sided at the top   (:239 / :288)
        d = (k == K - 1) ? slP[j] : d; // one-sided at the bottom
        Df[j] = d;
    }
    shift_next<NCH>(Df, DfN);

    // Bezier-interpolated optical depth of the forward interval (k, k+1) (:242-246, :261-263)
    double dtf[NCH], dtfP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const double ds3 = ds[j] / 3.0;
        const double cA = chi[j] + ds3 * Df[j];    // control point at k
        const double cB = chiN[j] - ds3 * DfN[j];  // control point at k+1
        const double t1 = chi[j] + chiN[j];
        dtf[j] = DOWN ? ds[j] * ((t1 + cA) + cB) * 0.25 : ds[j] * ((t1 + cB) + cA) * 0.25;
    }
    shift_prev<NCH>(dtf, dtfP);

    double slS[NCH], slSP[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        slS[j] = (SN[j] - S[j]) / dtf[j];
    shift_prev<NCH>(slS, slSP);

    double DSf[NCH], DSuw[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        double d = steffen(dtfP[j], dtf[j], slSP[j], slS[j]);
        d = (k == 0) ? slS[j] : d;      // (:247)
        d = (k == K - 1) ? slSP[j] : d;
        DSf[j] = d;
    }
    if (DOWN)
        shift_prev<NCH>(DSf, DSuw);
    else
        shift_next<NCH>(DSf, DSuw);

#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        const bool isEnd = (k == ke);
        const double chiUw = DOWN ? chiP[j] : chiN[j];
        const double Suw = DOWN ? SP[j] : SN[j];
        const double dsfUw = DOWN ? g.dsfP[j] : g.dsf[j];
        const double dtBez = DOWN ? dtfP[j] : dtf[j];
        // piecewise linear on the end (:307-321)
        const double dtEnd = 0.5 * zmu * (chi[j] + chiUw) * dsfUw;
        const double dt = isEnd ? dtEnd : dtBez;
        const bool taylor = isEnd ? (dt < 5.0E-4) : (dt < 5e-2);
        const bool thick = isEnd ? (dt > 50.0) : (dt > 30.0);
        double edt = 0.0;
        if (!taylor && !thick)
            edt = exp(-dt);

        double aa, bb, pp;
        if (isEnd)
        {
            double w0, w1; // w2(), LwInternal.hpp:90-110
            if (taylor)
            {
                w0 = dt * (1.0 - 0.5 * dt);
                w1 = (dt * dt) * (0.5 - dt * (1.0 / 3.0));
            }
            else if (thick)
            {
                w0 = w1 = 1.0;
            }
            else
            {
                w0 = 1.0 - edt;
                w1 = w0 - dt * edt;
            }
            const double dS = (S[j] - Suw) / dt;
            aa = 1.0 - w0;
            bb = w0 * S[j] - w1 * dS;
            pp = w0 - w1 / dt;
        }
        else
        {
            const double dt2 = dt * dt;
            const double dt3 = dt2 * dt;
            double alpha, beta, gamma, delta; // Bezier3_coeffs, Bezier.hpp:81-127
            if (taylor)
            {
                edt = 1.0 - dt + 0.5 * dt2 - dt3 / 6.0;
                alpha = 0.25 * dt - 0.2 * dt2 + dt3 / 12.0;
                beta = 0.25 * dt - 0.05 * dt2 + dt3 / 120.0;
                gamma = 0.25 * dt - 0.15 * dt2 + 0.05 * dt3;
                delta = 0.25 * dt - 0.1 * dt2 + 0.025 * dt3;
            }
            else
            {
                // dt > 30 is this branch with edt == 0, term for term
                alpha = (6.0 - edt * (6.0 + 6.0 * dt + 3 * dt2 + dt3)) / dt3;
                beta = (6.0 * edt - 6.0 + 6.0 * dt - 3.0 * dt2 + dt3) / dt3;
                gamma = 3.0 * (2.0 * dt - 6.0 + edt * (6.0 + 4.0 * dt + dt2)) / dt3;
                delta = 3.0 * (6.0 - 4.0 * dt + dt2 - 2.0 * edt * (3.0 + dt)) / dt3;
            }
            const double dt3rd = dt / 3.0;
            // path derivatives: +forward for DOWN rays, -forward for UP rays
            const double Cuw = DOWN ? Suw + dt3rd * DSuw[j] : Suw - dt3rd * DSuw[j];
            const double C0 = DOWN ? S[j] - dt3rd * DSf[j] : S[j] + dt3rd * DSf[j];
            aa = edt;
            bb = alpha * Suw + beta * S[j] + gamma * Cuw + delta * C0;
            pp = beta + delta;
        }
        if (k == ks)
        {
            // boundary intensity (:551-597)
            double Iupw = 0.0;
            const double chiDw = DOWN ? chiN[j] : chiP[j];
            const double dsfDw = DOWN ? g.dsf[j] : g.dsfP[j];
            if (bc.type == 2 /* THERMALISED */)
            {
                const double dtau_b = 0.5 * zmu * (chi[j] + chiDw) * dsfDw;
                Iupw = bc.B0 - (bc.B1 - bc.B0) / dtau_b;
            }
            else if (bc.type == 4 /* CALLABLE */)
            {
                Iupw = bc.value;
            }
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
        if (WANT_PSI)
            psi[j] = pp / chi[j];
    }
}

// ---------------------------------------------------------------------------
// piecewise_linear_1d (FormalScalar.cpp:136-207 + :471-533); zmu = 0.5/mu.
template <int NCH, bool DOWN, bool WANT_PSI>
__device__ __forceinline__ void linear_coefficients(const Geometry<NCH>& g,
                                                    const double (&chi)[NCH],
                                                    const double (&S)[NCH], double zmu,
                                                    const RayBc& bc, double (&a)[NCH],
                                                    double (&b)[NCH], double (&psi)[NCH])
{
    const int K = g.K;
    const int ks = DOWN ? 0 : K - 1;
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(chi, chiN);
    shift_prev<NCH>(chi, chiP);
    shift_next<NCH>(S, SN);
    shift_prev<NCH>(S, SP);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        const double chiUw = DOWN ? chiP[j] : chiN[j];
        const double Suw = DOWN ? SP[j] : SN[j];
        const double dsfUw = DOWN ? g.dsfP[j] : g.dsf[j];
        const double dt = zmu * (chi[j] + chiUw) * dsfUw;
        const double rcp = 1.0 / dt;
        const double dS = (Suw - S[j]) * rcp;
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
        double aa = 1.0 - w0;
        double bb = w0 * S[j] + w1 * dS;
        double pp = w0 - w1 * rcp;
        if (k == ks)
        {
            double Iupw = 0.0;
            const double chiDw = DOWN ? chiN[j] : chiP[j];
            const double dsfDw = DOWN ? g.dsf[j] : g.dsfP[j];
            if (bc.type == 2)
            {
                const double dtau_b = zmu * (chi[j] + chiDw) * dsfDw;
                Iupw = bc.B0 - (bc.B1 - bc.B0) / dtau_b;
            }
            else if (bc.type == 4)
            {
                Iupw = bc.value;
            }
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
        if (WANT_PSI)
            psi[j] = pp / chi[j];
    }
}

// besser_control_point_1d (FormalScalar.cpp:327-363)
__device__ __forceinline__ double besser_control_point(double hM, double hP, double yM, double yO,
                                                       double yP)
{
    const double dM = (yO - yM) / hM;
    const double dP = (yP - yO) / hP;
    if (dM * dP <= 0.0)
        return yO;
    double yOp = (hM * dP + hP * dM) / (hM + hP);
    double cM = yO - 0.5 * hM * yOp;
    double cP = yO + 0.5 * hP * yOp;
    double minYMO = yM, maxYMO = yO, minYOP = yO, maxYOP = yP;
    if (dM < 0.0)
    {
        minYMO = yO;
        maxYMO = yM;
        minYOP = yP;
        maxYOP = yO;
    }
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

// piecewise_besser_1d (FormalScalar.cpp:395-467 + :602-666); zmu = 1/mu.
template <int NCH, bool DOWN, bool WANT_PSI>
__device__ __forceinline__ void besser_coefficients(const Geometry<NCH>& g,
                                                    const double (&chi)[NCH],
                                                    const double (&S)[NCH], double zmu,
                                                    const RayBc& bc, double (&a)[NCH],
                                                    double (&b)[NCH], double (&psi)[NCH])
{
    const int K = g.K;
    const int ks = DOWN ? 0 : K - 1;
    const int ke = DOWN ? K - 1 : 0;
    double chiN[NCH], chiP[NCH], SN[NCH], SP[NCH];
    shift_next<NCH>(chi, chiN);
    shift_prev<NCH>(chi, chiP);
    shift_next<NCH>(S, SN);
    shift_prev<NCH>(S, SP);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
    {
        const int k = g.k(j);
        const bool isEnd = (k == ke);
        const double chiUw = DOWN ? chiP[j] : chiN[j];
        const double chiDw = DOWN ? chiN[j] : chiP[j];
        const double Suw = DOWN ? SP[j] : SN[j];
        const double Sdw = DOWN ? SN[j] : SP[j];
        const double dsfUw = DOWN ? g.dsfP[j] : g.dsf[j];
        const double dsfDw = DOWN ? g.dsf[j] : g.dsfP[j];
        double aa, bb, pp;
        if (isEnd)
        {
            const double dt = 0.5 * zmu * (chi[j] + chiUw) * dsfUw;
            const double dS = (S[j] - Suw) / dt;
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
            bb = w0 * S[j] - w1 * dS;
            pp = w0 - w1 / dt;
        }
        else
        {
            const double ds_uw = dsfUw * zmu;
            const double ds_dw = dsfDw * zmu;
            const double chiC = besser_control_point(ds_uw, ds_dw, chiUw, chi[j], chiDw);
            const double dtauUw = (1.0 / 3.0) * (chiUw + chiC + chi[j]) * ds_uw;
            const double dtauDw = 0.5 * (chi[j] + chiDw) * ds_dw;
            const double SC = besser_control_point(dtauUw, dtauDw, Suw, S[j], Sdw);
            const double t = dtauUw;
            double M, O, C, edt; // besser_coeffs_1d (:373-393)
            if (t < 0.14)
            {
                M = (t * (t * (t * (t * (t * (t * ((140.0 - 18.0 * t) * t - 945.0) + 5400.0) - 25200.0) + 90720.0) - 226800.0) + 302400.0)) / 907200.0;
                O = (t * (t * (t * (t * (t * (t * ((10.0 - t) * t - 90.0) + 720.0) - 5040.0) + 30240.0) - 151200.0) + 604800.0)) / 1814400.0;
                C = (t * (t * (t * (t * (t * (t * ((35.0 - 4.0 * t) * t - 270.0) + 1800.0) - 10080.0) + 45360.0) - 151200.0) + 302400.0)) / 907200.0;
                const double t2 = t * t, t3 = t2 * t;
                edt = 1.0 - t + 0.5 * t2 - t3 / 6.0 + t * t3 / 24.0 - t2 * t3 / 120.0 + t3 * t3 / 720.0 - t3 * t3 * t / 5040.0;
            }
            else
            {
                const double t2 = t * t;
                edt = exp(-t);
                M = (2.0 - edt * (t2 + 2.0 * t + 2.0)) / t2;
                O = 1.0 - 2.0 * (edt + t - 1.0) / t2;
                C = 2.0 * (t - 2.0 + edt * (t + 2.0)) / t2;
            }
            aa = edt;
            bb = M * Suw + O * S[j] + C * SC;
            pp = O + C;
        }
        if (k == ks)
        {
            double Iupw = 0.0;
            if (bc.type == 2)
            {
                const double dtau_b = 0.5 * zmu * (chi[j] + chiDw) * dsfDw;
                Iupw = bc.B0 - (bc.B1 - bc.B0) / dtau_b;
            }
            else if (bc.type == 4)
            {
                Iupw = bc.value;
            }
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
        if (WANT_PSI)
            psi[j] = pp / chi[j];
    }
}

// One ray: coefficients + scan.  SOLVER: 0 linear, 1 besser, 2 bezier3.
template <int NCH, int SOLVER, bool DOWN, bool WANT_PSI>
__device__ __forceinline__ void solve_ray(const Geometry<NCH>& g, const double (&chi)[NCH],
                                          const double (&S)[NCH], double muz, const RayBc& bc,
                                          double (&I)[NCH], double (&psi)[NCH])
{
    double a[NCH], b[NCH];
    if (SOLVER == 0)
        linear_coefficients<NCH, DOWN, WANT_PSI>(g, chi, S, 0.5 / muz, bc, a, b, psi);
    else if (SOLVER == 1)
        besser_coefficients<NCH, DOWN, WANT_PSI>(g, chi, S, 1.0 / muz, bc, a, b, psi);
    else
        bezier3_coefficients<NCH, DOWN, WANT_PSI>(g, chi, S, 1.0 / muz, bc, a, b, psi);
    affine_scan<NCH, DOWN>(a, b, I);
}

} // namespace lwb200
