// lwb200_ng.cuh -- Ng acceleration of the populations on the device (Source/Ng.hpp:16-163).
//
// The reference keeps one Ng object per atom on the host and runs it on atom.n.flatten() after
// every population update (LwMiddleLayer.pyx:3318-3346), which forces n back to the host every
// iteration.  Here the history ring [Norder + 2][whole packed n buffer] lives in HBM; one CTA per
// (atom, column) forms the weighted normal equations of that atom's [Nlevel][Nspace] block
// (block reduction of Norder^2 + Norder sums), solves them with the same solve_lin_eq, and applies
// the extrapolation.  Sums run in a different order than the reference's serial loops: results
// agree to rounding (tests: 1e-10 relative), not bitwise.
#pragma once
#include "lwb200_kernels.cuh"

namespace lwb200
{
constexpr int kNgMaxOrder = 4;

__device__ __forceinline__ double ng_block_sum(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(kFull, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0)
        sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w)
        s += sh[w];
    return s;
}

// ring: [R][ringStride]; the solution just stored sits in row (count - 1) % R.
__global__ void __launch_bounds__(256)
ng_accelerate_kernel(const DevProblem P, double* __restrict__ ring, int R, size_t ringStride, int count, int Norder,
                     double* __restrict__ n, int* __restrict__ nSingular)
{
    __shared__ double sh[8];
    __shared__ double xs[kNgMaxOrder];
    __shared__ int bad;
    const int atom = blockIdx.x, col = blockIdx.y;
    if (P.atomDetailed[atom] || (P.colActive && !P.colActive[col]))
        return;
    const int len = P.atomNlevel[atom] * P.K;
    const size_t base = ((size_t)col * P.NlevTot + P.atomLevOff[atom]) * P.K;
    const double* prev[kNgMaxOrder + 2];
#pragma unroll
    for (int i = 0; i < kNgMaxOrder + 2; ++i)
        prev[i] = ring + (size_t)((count - 1 - min(i, Norder + 1)) % R) * ringStride + base;
    double* sol = n + base;

    // weighted normal equations (:73-98): Delta(i) = prev[i] - prev[i + 1]
    double A[kNgMaxOrder][kNgMaxOrder], b[kNgMaxOrder];
#pragma unroll
    for (int i = 0; i < kNgMaxOrder; ++i)
    {
        b[i] = 0.0;
#pragma unroll
        for (int j = 0; j < kNgMaxOrder; ++j)
            A[i][j] = 0.0;
    }
    for (int k = threadIdx.x; k < len; k += blockDim.x)
    {
        double p[kNgMaxOrder + 2];
#pragma unroll
        for (int i = 0; i < kNgMaxOrder + 2; ++i)
            p[i] = (i <= Norder + 1) ? prev[i][k] : 0.0;
        const double w = 1.0 / fabs(sol[k]);
        const double D0 = p[0] - p[1];
        double dd[kNgMaxOrder];
#pragma unroll
        for (int j = 0; j < kNgMaxOrder; ++j)
            dd[j] = (j < Norder) ? (p[j + 1] - p[j + 2]) - D0 : 0.0; // Delta(j + 1) - Delta(0)
#pragma unroll
        for (int j = 0; j < kNgMaxOrder; ++j)
        {
            b[j] += w * D0 * (-dd[j]);
#pragma unroll
            for (int i = 0; i < kNgMaxOrder; ++i)
                A[i][j] += w * dd[j] * dd[i];
        }
    }
    double Af[kNgMaxOrder * kNgMaxOrder], bf[kNgMaxOrder];
#pragma unroll
    for (int j = 0; j < kNgMaxOrder; ++j)
    {
        const double bj = ng_block_sum(b[j], sh);
        if (j < Norder)
            bf[j] = bj;
#pragma unroll
        for (int i = 0; i < kNgMaxOrder; ++i)
        {
            const double aij = ng_block_sum(A[i][j], sh);
            if (i < Norder && j < Norder)
                Af[i * Norder + j] = aij;
        }
    }
    if (threadIdx.x == 0)
    {
        // solve_lin_eq with its refinement step (LuSolve.cpp:103-133)
        double ACopy[kNgMaxOrder * kNgMaxOrder], bCopy[kNgMaxOrder], res[kNgMaxOrder];
        int index[kNgMaxOrder];
        for (int i = 0; i < Norder * Norder; ++i)
            ACopy[i] = Af[i];
        for (int i = 0; i < Norder; ++i)
            bCopy[i] = bf[i];
        bad = 0;
        if (!lu_decompose_dev<kNgMaxOrder>(Norder, Af, index))
        {
            bad = 1;
            atomicAdd_system(nSingular, 1);
        }
        else
        {
            lu_backsub_dev(Norder, Af, index, bf);
            for (int i = 0; i < Norder; ++i)
            {
                double r = bCopy[i];
                for (int j = 0; j < Norder; ++j)
                    r -= ACopy[i * Norder + j] * bf[j];
                res[i] = r;
            }
            lu_backsub_dev(Norder, Af, index, res);
            for (int i = 0; i < Norder; ++i)
                xs[i] = bf[i] + res[i];
        }
    }
    __syncthreads();
    if (bad)
        return;
    // sol += sum_i x(i) (previous(count - i - 2) - previous(count - 1)); previous(count - 1) = sol (:102-111)
    double* p0 = ring + (size_t)((count - 1) % R) * ringStride + base;
    for (int k = threadIdx.x; k < len; k += blockDim.x)
    {
        double s = sol[k];
        const double c0 = p0[k];
        for (int i = 0; i < Norder; ++i)
            s += xs[i] * (prev[i + 1][k] - c0);
        sol[k] = s;
        p0[k] = s;
    }
}

// Ng::max_change (:138-156) per atom over every active column: max |(cur - old) / cur|, first index.
// Grid (atom, part): every CTA reduces a contiguous share of the atom's [Ncol][Nlevel][K] elements into
// slot 1 + part of that atom's row of outMax / outIdx (rows of kNgParts + 1); ng_max_change_final_kernel
// folds the parts into slot 0.  (One CTA per atom was 4 ms on a 4096-column stack.)
constexpr int kNgParts = 128;

__global__ void __launch_bounds__(256)
ng_max_change_kernel(const DevProblem P, const double* __restrict__ cur, const double* __restrict__ old,
                     double* __restrict__ outMax, long long* __restrict__ outIdx)
{
    __shared__ double sMax[256];
    __shared__ long long sIdx[256];
    const int atom = blockIdx.x;
    const long long len = (long long)P.atomNlevel[atom] * P.K;
    const long long total = len * P.Ncol;
    const long long per = (total + gridDim.y - 1) / gridDim.y;
    const long long qBeg = per * blockIdx.y, qEnd = qBeg + per < total ? qBeg + per : total;
    double best = 0.0;
    long long bestIdx = 0;
    if (!P.atomDetailed[atom])
        for (long long q = qBeg + threadIdx.x; q < qEnd; q += blockDim.x)
        {
            const long long col = q / len, e = q % len;
            if (P.colActive && !P.colActive[col])
                continue;
            const size_t o = ((size_t)col * P.NlevTot + P.atomLevOff[atom]) * P.K + e;
            const double c = cur[o];
            if (c != 0.0)
            {
                const double change = fabs((c - old[o]) / c);
                if (best < change)
                {
                    best = change;
                    bestIdx = q;
                }
            }
        }
    sMax[threadIdx.x] = best;
    sIdx[threadIdx.x] = bestIdx;
    dj_block_reduce(sMax, sIdx);
    if (threadIdx.x == 0)
    {
        outMax[(size_t)atom * (kNgParts + 1) + 1 + blockIdx.y] = sMax[0];
        outIdx[(size_t)atom * (kNgParts + 1) + 1 + blockIdx.y] = sIdx[0];
    }
}

__global__ void __launch_bounds__(kNgParts)
ng_max_change_final_kernel(int nParts, double* __restrict__ outMax, long long* __restrict__ outIdx)
{
    __shared__ double sMax[kNgParts];
    __shared__ long long sIdx[kNgParts];
    const size_t row = (size_t)blockIdx.x * (kNgParts + 1);
    const bool have = (int)threadIdx.x < nParts;
    sMax[threadIdx.x] = have ? outMax[row + 1 + threadIdx.x] : 0.0;
    sIdx[threadIdx.x] = have ? outIdx[row + 1 + threadIdx.x] : 0;
    dj_block_reduce(sMax, sIdx);
    if (threadIdx.x == 0)
    {
        // (no change anywhere: index 0, as a serial first-max scan would report)
        outMax[row] = sMax[0];
        outIdx[row] = sMax[0] > 0.0 ? sIdx[0] : 0;
    }
}

} // namespace lwb200
