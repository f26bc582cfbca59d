/* fp64_peak.cu -- measures the fp64 pipe peak of the GPU it runs on ("the builder must microbenchmark
 * the DFMA peak on the box").  Three streams of independent register-only operations per thread: DFMA, and the
 * DMUL / DADD mix the solver is restricted to (-fmad=false, for bit-exact parity with the reference).
 * Output: one JSON line.  Build: see ddp-generator_b200/Makefile (sm_100a). */
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_peak(double *out, int iters, double seed)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-9;
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __fma_rn(a[i], m, c);          /* 1 instruction, 2 flop */
            if (MODE == 1) a[i] = __dadd_rn(__dmul_rn(a[i], m), c); /* 2 instructions, 2 flop */
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE> static double run(int iters)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    double *d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_peak<MODE><<<blocks, threads>>>(d, iters / 8, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_peak<MODE><<<blocks, threads>>>(d, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    const double ops = (double)blocks * threads * (double)iters * 8.0; /* loop bodies executed by all threads */
    return ops / (best * 1e-3);
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    const double fma = run<0>(1 << 16), muladd = run<1>(1 << 16);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.4g, \"dfma_tflops\": %.3f, \"dmul_dadd_pairs_per_s\": %.4g, "
           "\"dmul_dadd_instr_per_s\": %.4g, \"no_fma_tflops\": %.3f}\n",
           p.name, p.multiProcessorCount, fma, 2 * fma / 1e12, muladd, 2 * muladd, 2 * muladd / 1e12);
    return 0;
}
