// lwb200_profiles.cuh -- line absorption profiles on the device.
//
// Replaces Transition::compute_phi / compute_phi_la / compute_wphi
// (Source/FormalScalar.cpp:28-134): phi(lt, mu, dir, k) = H(a_k, v) / (sqrt(pi) vBroad_k),
// v = ((lambda - lambda0) c / lambda0 -/+ vlosMu(mu, k)) / vBroad_k, and
// wphi_k = 1 / sum_{lt,mu,dir} phi * wlambda(lt) * wmu/2.
//
// The reference evaluates the Voigt function through the vendored Faddeeva
// package (Source/Faddeeva.cc, not part of the hot path and not restated).
// Here H(a, v) = Re w(v + i a) is evaluated from first principles:
//   |v| >= 8 : Laplace continued fraction w(z) = (i/sqrt(pi)) / (z - (1/2)/(z - 1/(z - ...)))
//              truncated at a depth chosen from |v| (<= 1e-15 relative);
//   |v| <  8 : midpoint rule for (a/pi) int exp(-t^2) / ((v-t)^2 + a^2) dt on nodes
//              t_m = v + (m + 1/2) h, which keeps every node at least h/2 away from the pole,
//              plus the residue correction 2 Re[exp(-z^2)] / (1 + exp(2 pi a / h))
//              (Matta & Reichel 1971); with h = 1/2 the truncation error is e^{-4 pi^2} ~ 1e-17.
// Both agree with Faddeeva's w(z) to < 3e-14 relative for a in [1e-4, 1] (tools/voigt_check.py).
#pragma once
#include "lwb200_kernels.cuh"

namespace lwb200
{
struct DevLine
{
    int Nl;
    int tabOff;
    int lineIdx;
    int atom;
    long long phiOff;
    long long phiColStride;
    double lambda0;
};

__device__ __forceinline__ double voigt_H(double a, double v)
{
    const double x = fabs(v);
    if (x >= 8.0)
    {
        const int n = x < 10.0 ? 12 : x < 15.0 ? 10 : x < 25.0 ? 8 : x < 50.0 ? 6 : x < 150.0 ? 4 : 3;
        // r = z - (k/2)/r, from the tail inwards, z = x + i a
        double rr = x, ri = a;
        for (int k = n; k >= 1; --k)
        {
            const double c = 0.5 * k;
            const double d = 1.0 / (rr * rr + ri * ri);
            const double qr = c * rr * d;
            const double qi = -c * ri * d;
            rr = x - qr;
            ri = a - qi;
        }
        // Re(i / (sqrt(pi) r)) = ri / (sqrt(pi) |r|^2)
        return 0.56418958354775628695 * ri / (rr * rr + ri * ri);
    }
    constexpr double h = 0.5;
    constexpr double T = 6.3;
    const double mlo = ceil((-T - x) / h - 0.5);
    double acc = 0.0;
    const double a2 = a * a;
#pragma unroll 1
    for (int i = 0; i < 27; ++i)
    {
        const double s = (mlo + i + 0.5) * h;
        const double t = x + s;
        acc += exp(-t * t) * a / (s * s + a2);
    }
    const double corr = 2.0 * exp(a2 - x * x) * cos(2.0 * x * a) / (1.0 + exp(2.0 * kPi * a / h));
    return (h / kPi) * acc + corr;
}

// phi[col][lt][mu][dir][k] for every line; one thread per element, k fastest.
__global__ void phi_kernel(const DevProblem P, const DevLine* __restrict__ lines, int line,
                           const double* __restrict__ transWave, const double* __restrict__ aDamp,
                           const double* __restrict__ vBroad, const double* __restrict__ vlosMu,
                           double* __restrict__ phi)
{
    const DevLine ln = lines[line];
    const size_t perCol = (size_t)ln.Nl * P.M * 2 * P.K;
    const size_t total = perCol * P.Ncol;
    const double sqrtPi = sqrt(kPi);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = idx % P.K;
        size_t r = idx / P.K;
        const int dir = r % 2;
        r /= 2;
        const int mu = r % P.M;
        r /= P.M;
        const int lt = r % ln.Nl;
        const int col = r / ln.Nl;
        const double vBase = (transWave[ln.tabOff + lt] - ln.lambda0) * kCLight / ln.lambda0;
        const double s = dir ? 1.0 : -1.0;
        const double vb = vBroad[((size_t)col * P.Natom + ln.atom) * P.K + k];
        const double vk = (vBase + s * vlosMu[((size_t)col * P.M + mu) * P.K + k]) / vb;
        const double a = aDamp[((size_t)ln.lineIdx * P.Ncol + col) * P.K + k];
        phi[ln.phiOff + idx] = voigt_H(a, vk) / (sqrtPi * vb);
    }
}

// wphi[line][col][k], same summation order as compute_wphi (FormalScalar.cpp:106-134)
__global__ void wphi_kernel(const DevProblem P, const DevLine* __restrict__ lines, int nlines,
                            const double* __restrict__ wlambdaTab, const double* __restrict__ phi,
                            double* __restrict__ wphi)
{
    const size_t total = (size_t)nlines * P.Ncol * P.K;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int k = idx % P.K;
        const int col = (idx / P.K) % P.Ncol;
        const int line = idx / ((size_t)P.K * P.Ncol);
        const DevLine ln = lines[line];
        const double* ph = phi + ln.phiOff + (size_t)col * ln.phiColStride + k;
        double sum = 0.0;
        for (int lt = 0; lt < ln.Nl; ++lt)
        {
            const double wla = wlambdaTab[ln.tabOff + lt];
            for (int mu = 0; mu < P.M; ++mu)
            {
                const double wlamu = wla * 0.5 * P.wmu[mu];
                for (int dir = 0; dir < 2; ++dir)
                    sum += ph[((size_t)(lt * P.M + mu) * 2 + dir) * P.K] * wlamu;
            }
        }
        wphi[((size_t)ln.lineIdx * P.Ncol + col) * P.K + k] = 1.0 / sum;
    }
}

inline int launch_profiles(const DevProblem& P, const DevLine* dLines, int nlines, const DevLine* hLines,
                           const double* transWave, const double* wlambdaTab, const double* aDamp,
                           const double* vBroad, const double* vlosMu, double* phi, double* wphi,
                           cudaStream_t stream, int64_t* launches)
{
    for (int l = 0; l < nlines; ++l)
    {
        const size_t total = (size_t)hLines[l].Nl * P.M * 2 * P.K * P.Ncol;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, 148 * 64));
        phi_kernel<<<grid, 256, 0, stream>>>(P, dLines, l, transWave, aDamp, vBroad, vlosMu, phi);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return (int)e;
        *launches += 1;
    }
    const size_t total = (size_t)nlines * P.Ncol * P.K;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 127) / 128, 148 * 32));
    wphi_kernel<<<grid, 128, 0, stream>>>(P, dLines, nlines, wlambdaTab, phi, wphi);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return (int)e;
    *launches += 1;
    return 0;
}

} // namespace lwb200
