// FP64 / FP32 pipe peak of the box's GPU, measured with dependent-chain-free register kernels (SURVEY 8(d):
// "FP64 peak to be measured with a DADD/DMUL microbenchmark on the box").  The narrow-phase kernels are compiled
// with --fmad=false, so their roofline denominator is the DADD/DMUL ISSUE rate (1 instruction = 1 flop), not the
// DFMA flop rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/fp64_peak tools/fp64_peak.cu
//   tools/fp64_peak            -> one JSON line: instructions/s per pipe, per SM and clock
#include <cuda_runtime.h>
#include <stdio.h>

#define ILP 8
#define ITERS 4096

template <int MODE>  // 0: DADD+DMUL alternating (no FMA), 1: DFMA, 2: FADD+FMUL, 3: FFMA
__global__ void __launch_bounds__(256) k_peak(double* out, double a, double b)
{
    double x[ILP];
    float y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a + i + threadIdx.x; y[i] = (float)(a + i + threadIdx.x); }
    const float af = (float)a, bf = (float)b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { x[i] = __dmul_rn(x[i], a); x[i] = __dadd_rn(x[i], b); }
            if (MODE == 1) { x[i] = __fma_rn(x[i], a, b); x[i] = __fma_rn(x[i], b, a); }
            if (MODE == 2) { y[i] = __fmul_rn(y[i], af); y[i] = __fadd_rn(y[i], bf); }
            if (MODE == 3) { y[i] = __fmaf_rn(y[i], af, bf); y[i] = __fmaf_rn(y[i], bf, af); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + (double)y[i];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
static double run(int sms)
{
    double* d;
    cudaMalloc(&d, 8);
    const int blocks = sms * 8;
    k_peak<MODE><<<blocks, 256>>>(d, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_peak<MODE><<<blocks, 256>>>(d, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double inst = (double)blocks * 256 * ITERS * ILP * 2;  // thread-level instructions
        const double rate = inst / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaFree(d);
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    const double d_nofma = run<0>(sms), d_fma = run<1>(sms), f_nofma = run<2>(sms), f_fma = run<3>(sms);
    const double per = 1.0 / (sms * (double)clk * 1e3);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %d, "
           "\"fp64_inst_per_s_nofma\": %.4g, \"fp64_inst_per_s_fma\": %.4g, \"fp32_inst_per_s_nofma\": %.4g, \"fp32_inst_per_s_fma\": %.4g, "
           "\"fp64_lanes_per_sm_clk\": %.2f, \"fp32_lanes_per_sm_clk\": %.2f, "
           "\"fp64_tflops_nofma\": %.3f, \"fp64_tflops_fma\": %.3f}\n",
           p.name, sms, clk / 1000, d_nofma, d_fma, f_nofma, f_fma, d_nofma * per, f_fma * per, d_nofma / 1e12, 2 * d_fma / 1e12);
    return 0;
}
