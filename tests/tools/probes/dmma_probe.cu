// dmma_probe.cu — does the FP64 tensor sub-pipe (mma.sync m8n8k4 f64, "DMMA") run beside the FP64 FMA pipe on sm_100a, and at what rate?
// Answers whether part of the pair arithmetic of the sweep could move to DMMA (profiles/r02_experiments.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// per thread: ND independent DMMA accumulator pairs and NF independent DFMA chains, `iters` rounds
template <int ND, int NF>
__global__ void probe(double* out, int iters, double a, double b) {
    double d[ND > 0 ? 2 * ND : 1], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int k = 0; k < 2 * ND; ++k) d[k] = threadIdx.x + k;
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = threadIdx.x * 0.5 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ND; ++k) dmma(d[2 * k], d[2 * k + 1], a, b);
#pragma unroll
        for (int k = 0; k < NF; ++k) f[k] = fma(f[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 2 * ND; ++k) s += d[k];
#pragma unroll
    for (int k = 0; k < NF; ++k) s += f[k];
    if (s == 12345.678) out[0] = s;
}

template <int ND, int NF>
double run(int blocks, int threads, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<ND, NF><<<blocks, threads>>>(d_out, 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        probe<ND, NF><<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 2, iters = 20000;
    double* d_out; cudaMalloc(&d_out, 8);
    const double warps = (double)blocks * threads / 32;
    const double clk = p.clockRate * 1e3;      // Hz (max boost; informative only)
    auto report = [&](const char* name, double ms, int nd, int nf) {
        const double dm = warps * iters * nd, df = warps * iters * nf;
        const double tf = (dm * 256 * 2 + df * 32 * 2) / (ms * 1e-3) / 1e12;
        const double cyc_per_sm = ms * 1e-3 * clk;
        printf("%-28s %8.3f ms  %6.2f TFLOP/s   per SM and clock: %.2f DMMA + %.2f warp-DFMA   (clock %.0f MHz assumed)\n", name, ms, tf,
               dm / sms / cyc_per_sm, df / sms / cyc_per_sm, clk / 1e6);
    };
    report("DFMA only (8 chains)", run<0, 8>(blocks, threads, iters, d_out), 0, 8);
    report("DMMA only (4 accumulators)", run<4, 0>(blocks, threads, iters, d_out), 4, 0);
    report("DMMA only (8 accumulators)", run<8, 0>(blocks, threads, iters, d_out), 8, 0);
    report("4 DMMA + 8 DFMA per round", run<4, 8>(blocks, threads, iters, d_out), 4, 8);
    report("1 DMMA + 8 DFMA per round", run<1, 8>(blocks, threads, iters, d_out), 1, 8);
    report("4 DMMA + 22 DFMA per round", run<4, 22>(blocks, threads, iters, d_out), 4, 22);
    report("0 DMMA + 22 DFMA per round", run<0, 22>(blocks, threads, iters, d_out), 0, 22);
    return 0;
}
