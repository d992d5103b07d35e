// Measures the sustained fp64 FMA rate of the device (secondary roofline of DESIGN.md section 5).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
    {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma<<<blocks, threads>>>(out, 1000);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
        cudaEventRecord(e0);
        dfma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fmas = (double)blocks * threads * iters * 8;
    printf("{\"sms\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm_at_1965MHz\": %.1f}\n", sms,
           fmas / (best * 1e-3), 2 * fmas / (best * 1e-3) / 1e12, fmas / (best * 1e-3) / sms / 1.965e9);
    return 0;
}
