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
// Both agree with Faddeeva's w(z) to < 3e-14 relative for a in [1e-4, 1]
// (tests/test_gpu_parity.py::test_device_profiles_match_host_voigt; the imaginary part:
// ::test_device_polarised_profiles_match_host_and_feed_the_stokes_solver).
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

// w(v + i a) = H + i F: the Voigt and Faraday-Voigt functions, by the two methods of voigt_H carried through for the
// imaginary part: Im[(i / sqrt(pi)) / r] = Re r / (sqrt(pi) |r|^2) for the continued fraction; for the midpoint rule
// Im w = (1/pi) int exp(-t^2) (v - t) / ((v - t)^2 + a^2) dt on the same nodes, with the residue correction
// 2 Im[exp(-z^2)] / (1 + exp(2 pi a / h)).  F is odd in v.
__device__ __forceinline__ void voigt_HF(double a, double v, double& H, double& F)
{
    const double x = fabs(v);
    const double sgn = v < 0.0 ? -1.0 : 1.0;
    if (x >= 8.0)
    {
        const int n = x < 10.0 ? 12 : x < 15.0 ? 10 : x < 25.0 ? 8 : x < 50.0 ? 6 : x < 150.0 ? 4 : 3;
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
        const double d = 0.56418958354775628695 / (rr * rr + ri * ri);
        H = ri * d;
        F = sgn * rr * d;
        return;
    }
    constexpr double h = 0.5;
    constexpr double T = 6.3;
    const double mlo = ceil((-T - x) / h - 0.5);
    double accH = 0.0, accF = 0.0;
    const double a2 = a * a;
#pragma unroll 1
    for (int i = 0; i < 27; ++i)
    {
        const double s = (mlo + i + 0.5) * h;
        const double t = x + s;
        const double e = exp(-t * t) / (s * s + a2);
        accH += e * a;
        accF -= e * s;
    }
    const double ez = 2.0 * exp(a2 - x * x) / (1.0 + exp(2.0 * kPi * a / h));
    double sn, cs;
    sincos(2.0 * x * a, &sn, &cs);
    H = (h / kPi) * accH + ez * cs;
    F = sgn * ((h / kPi) * accF - ez * sn);
}

// One Zeeman pattern on the device
struct DevZeeman
{
    int line;       // index into the DevLine table
    int polIdx;     // which polarised line (for the pol pool offset table)
    int nComp;
    int compOff;    // offset into the packed alpha / shift / strength arrays
    long long polOff;   // element offset of phiQ(col 0) of this line in the pol pool
    long long polArr;   // stride between the six arrays
};

// Transition::compute_polarised_profiles (FormalStokes.cpp:44-108): one thread per (col, lt, mu, dir, k)
__global__ void pol_profile_kernel(const DevProblem P, const DevLine* __restrict__ lines, DevZeeman z,
                                   const int* __restrict__ alpha, const double* __restrict__ shift,
                                   const double* __restrict__ strength, const double* __restrict__ transWave,
                                   const double* __restrict__ aDamp, const double* __restrict__ vBroad,
                                   const double* __restrict__ vlosMu, const double* __restrict__ B,
                                   const double* __restrict__ cosGamma, const double* __restrict__ cos2chi,
                                   const double* __restrict__ sin2chi, double* __restrict__ phi,
                                   double* __restrict__ pol)
{
    const DevLine ln = lines[z.line];
    const size_t total = (size_t)ln.Nl * P.M * 2 * P.K * P.Ncol;
    constexpr double kQElectron = 1.60217733E-19, kMElectron = 9.1093897E-31; // Constants.hpp
    const double larmor = kQElectron / (4.0 * kPi * kMElectron) * (ln.lambda0 * kNmToM);
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
        const double vB = larmor * B[(size_t)col * P.K + k] / vb;
        const double sv = 1.0 / (sqrtPi * vb);
        double phi_sb = 0.0, phi_pi = 0.0, phi_sr = 0.0, psi_sb = 0.0, psi_pi = 0.0, psi_sr = 0.0;
        for (int nz = 0; nz < z.nComp; ++nz)
        {
            double H, F;
            voigt_HF(a, vk - shift[z.compOff + nz] * vB, H, F);
            const double st = strength[z.compOff + nz];
            const int al = alpha[z.compOff + nz];
            if (al == -1)
            {
                phi_sb += st * H;
                psi_sb += st * F;
            }
            else if (al == 0)
            {
                phi_pi += st * H;
                psi_pi += st * F;
            }
            else if (al == 1)
            {
                phi_sr += st * H;
                psi_sr += st * F;
            }
        }
        const size_t ang = ((size_t)col * P.M + mu) * P.K + k;
        const double cos_gamma = cosGamma[ang];
        const double sin2_gamma = 1.0 - cos_gamma * cos_gamma;
        const double cos_2chi = cos2chi[ang], sin_2chi = sin2chi[ang];
        const double phi_sigma = phi_sr + phi_sb;
        const double phi_delta = 0.5 * phi_pi - 0.25 * phi_sigma;
        const double psi_sigma = psi_sr + psi_sb;
        const double psi_delta = 0.5 * psi_pi - 0.25 * psi_sigma;
        phi[ln.phiOff + idx] = (phi_delta * sin2_gamma + 0.5 * phi_sigma) * sv;
        double* pp = pol + z.polOff + idx;
        pp[0] = s * phi_delta * sin2_gamma * cos_2chi * sv;
        pp[z.polArr] = phi_delta * sin2_gamma * sin_2chi * sv;
        pp[2 * z.polArr] = s * 0.5 * (phi_sr - phi_sb) * cos_gamma * sv;
        pp[3 * z.polArr] = s * psi_delta * sin2_gamma * cos_2chi * sv;
        pp[4 * z.polArr] = psi_delta * sin2_gamma * sin_2chi * sv;
        pp[5 * z.polArr] = s * 0.5 * (psi_sr - psi_sb) * cos_gamma * sv;
    }
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
