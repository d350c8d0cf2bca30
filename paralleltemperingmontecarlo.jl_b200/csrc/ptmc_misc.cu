// ptmc_misc.cu — small kernels around the sweep: state initialisation, ladder-axis histogram reduction,
// and the FP64 pipe peak measurement that bench.py uses as its roofline denominator.
#include "sweep_kernel.cuh"

namespace ptmc {

int sweep_max_threads(int G) { return sweep_max_threads_c(G); }

void launch_init(const DevParams& P, i64 n_rep, cudaStream_t st) {
    const int threads = 128;                         // 4 replicas (warps) per CTA
    const unsigned blocks = (unsigned)((n_rep * 32 + threads - 1) / threads);
    if (P.pot_kind == PTMC_POT_EAM) init_kernel<PTMC_POT_EAM><<<blocks, threads, 0, st>>>(P, n_rep);
    else if (P.pot_kind == PTMC_POT_ELJ) init_kernel<PTMC_POT_ELJ><<<blocks, threads, 0, st>>>(P, n_rep);
    else if (P.pot_kind == PTMC_POT_ELJ_B) init_kernel<PTMC_POT_ELJ_B><<<blocks, threads, 0, st>>>(P, n_rep);
    else if (P.pot_kind == PTMC_POT_LUT) init_kernel<PTMC_POT_LUT><<<blocks, threads, 0, st>>>(P, n_rep);
    else init_kernel<PTMC_POT_ELJ_EVEN><<<blocks, threads, 0, st>>>(P, n_rep);
}

// out[k] = sum over ladders of hist[l][k]; exact (u64 counts).
__global__ void reduce_hist_kernel(const u64* __restrict__ hist, u64* __restrict__ out, i64 n_ladders, i64 per) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= per) return;
    u64 s = 0;
    for (i64 l = 0; l < n_ladders; ++l) s += hist[l * per + k];
    out[k] = s;
}
void launch_reduce_hist(const u64* hist, u64* out, i64 n_ladders, i64 per, cudaStream_t st) {
    const int threads = 256;
    reduce_hist_kernel<<<(unsigned)((per + threads - 1) / threads), threads, 0, st>>>(hist, out, n_ladders, per);
}

// update_rdf! (src/Sampling.jl:301-316): one CTA per replica slot, private bins in shared memory.  Distances with the
// reference's distance2 arithmetic from the slot-ordered positions (P.pos is current after every ptmc_run_cycles).
__global__ void rdf_kernel(const DevParams P, u64* rdf, const double delta_r2, const int nrb) {
    extern __shared__ unsigned int rdf_bins[];
    const i64 rep = blockIdx.x;
    const int N = P.n_atoms;
    for (int k = threadIdx.x; k < nrb; k += blockDim.x) rdf_bins[k] = 0u;
    __syncthreads();
    const double* gp = P.pos + rep * 3 * N;
    const double L = P.box[rep * 2], h = P.box[rep * 2 + 1];
    for (int p = threadIdx.x; p < N * N; p += blockDim.x) {
        const int a = p / N, b = p - a * N;
        if (a >= b) continue;
        const double d2 = dist2_exact(P.bc_kind, gp[3 * a], gp[3 * a + 1], gp[3 * a + 2], gp[3 * b], gp[3 * b + 1], gp[3 * b + 2], L, h);
        const double q = floor(xdiv(d2, delta_r2));                  // rdf_index :282
        if (q >= 1.0 && q <= (double)nrb) atomicAdd(&rdf_bins[(int)q - 1], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nrb; k += blockDim.x) if (rdf_bins[k]) rdf[rep * nrb + k] += rdf_bins[k];
}
void launch_rdf(const DevParams& P, u64* rdf, double delta_r2, int n_rdf_bins, i64 n_rep, cudaStream_t st) {
    rdf_kernel<<<(unsigned)n_rep, 256, (size_t)n_rdf_bins * sizeof(unsigned int), st>>>(P, rdf, delta_r2, n_rdf_bins);
}

// Steps of the lookup table's angle index (ptmc_device.cuh: lut_angle_index): thread k bisects for the smallest tangent
// whose index is >= k + 2 (positive doubles order like their bit patterns; index(0) = 1, index(inf) = 91).
__global__ void lut_thresholds_kernel(const DevParams P, double* thr) {
    const int k = threadIdx.x;
    if (k >= PTMC_LUT_THR) return;
    const int target = k + 2;
    double out = __longlong_as_double(0x7ff8000000000000ll);          // NaN: never reached
    if (target <= 91) {
        unsigned long long lo = 0ull, hi = 0x7ff0000000000000ull;      // index(lo) < target <= index(hi)
        while (hi - lo > 1) {
            const unsigned long long mid = lo + (hi - lo) / 2;
            if (lut_angle_index_direct(P, __longlong_as_double((long long)mid)) >= target) hi = mid; else lo = mid;
        }
        out = __longlong_as_double((long long)hi);
    }
    thr[k] = out;
}
void launch_lut_thresholds(const DevParams& P, double* thr, cudaStream_t st) { lut_thresholds_kernel<<<1, 128, 0, st>>>(P, thr); }

// Accuracy probe of the production pair arithmetic: e_new[k], e_old[k] exactly as the sweep kernel accumulates
// them for the pair (d2_new[k], d2_old[k]).
__global__ void pair_probe_kernel(const DevParams P, int n, const double* d2n, const double* d2o, double* en, double* eo) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double a = 0.0, b = 0.0;
    if (P.pot_kind == PTMC_POT_ELJ_EVEN) {
        if (P.n_coeff == 6) elj_even_pair_acc<6>(P, d2n[k], d2o[k], a, b);
        else elj_even_pair_acc<0>(P, d2n[k], d2o[k], a, b);
    } else if (P.pot_kind == PTMC_POT_ELJ) {
        a = elj_odd(P, d2n[k]); b = elj_odd(P, d2o[k]);
    } else {
        double p, r;
        eam_invrexp(P, P.eam_ktab, P.eam_etab, d2n[k], p, r); a = p;      // phi term
        b = sqrt_pos(d2o[k]);                                             // the kernel's square root (of the rho sums)
    }
    en[k] = a; eo[k] = b;
}
int run_pair_probe(const DevParams& P, int n, const double* h_d2n, const double* h_d2o, double* h_en, double* h_eo) {
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * 4 * (size_t)n) != cudaSuccess) return PTMC_ERR_CUDA;
    cudaMemcpy(d, h_d2n, sizeof(double) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, h_d2o, sizeof(double) * n, cudaMemcpyHostToDevice);
    pair_probe_kernel<<<(n + 127) / 128, 128>>>(P, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) {
        cudaMemcpy(h_en, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_eo, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    return e == cudaSuccess ? PTMC_OK : PTMC_ERR_CUDA;
}

// FP64 pipe peak: 8 independent DFMA chains per thread, all SMs, 2 flops per DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int n_iter, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n_iter; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;   // keep the chains alive
}
// Issue-slot probe: K independent integer (LOP3/IADD) instructions interleaved with every DFMA.  If a warp DFMA only
// occupied the FP64 pipe for 2 cycles, up to one other instruction per DFMA would be free; the measured time per DFMA
// as a function of K tells how FP64 and non-FP64 instructions share the issue port (cost model).
template <int K>
__global__ void __launch_bounds__(256) issue_probe_kernel(double* out, int n_iter, double a, double b, unsigned seed) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned u0 = seed + threadIdx.x, u1 = u0 * 3u, u2 = u0 * 5u, u3 = u0 * 7u;
    for (int i = 0; i < n_iter; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#define PTMC_PROBE_INT(n) if (K > n) { u0 = (u0 ^ u1) + 0x9e3779b9u; u1 = (u1 ^ u2) + u0; u2 = (u2 ^ u3) + u1; u3 = (u3 ^ u0) + u2; }
            x0 = fma(x0, a, b); PTMC_PROBE_INT(0) x1 = fma(x1, a, b); PTMC_PROBE_INT(1) x2 = fma(x2, a, b); PTMC_PROBE_INT(2)
            x3 = fma(x3, a, b); PTMC_PROBE_INT(3) x4 = fma(x4, a, b); PTMC_PROBE_INT(0) x5 = fma(x5, a, b); PTMC_PROBE_INT(1)
            x6 = fma(x6, a, b); PTMC_PROBE_INT(2) x7 = fma(x7, a, b); PTMC_PROBE_INT(3)
#undef PTMC_PROBE_INT
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678 || (u0 ^ u1 ^ u2 ^ u3) == 0x12345u) out[0] = s;
}
// returns SM cycles per warp-DFMA per SM sub-partition for K in {0, 1, 2, 4} (8 int instructions per PROBE_INT block)
int run_issue_probe(double sm_mhz, double out_cycles[4]) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return PTMC_ERR_CUDA;
    const int threads = 256, blocks = sms * 8, n_iter = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 4; ++v) {
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            if (v == 0) issue_probe_kernel<0><<<blocks, threads>>>(d, n_iter, 0.999999, 1e-9, 1u);
            if (v == 1) issue_probe_kernel<1><<<blocks, threads>>>(d, n_iter, 0.999999, 1e-9, 1u);
            if (v == 2) issue_probe_kernel<2><<<blocks, threads>>>(d, n_iter, 0.999999, 1e-9, 1u);
            if (v == 3) issue_probe_kernel<4><<<blocks, threads>>>(d, n_iter, 0.999999, 1e-9, 1u);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return PTMC_ERR_CUDA; }
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        // warp-DFMAs per SMSP: blocks*threads/32 warps over sms*4 SMSPs, 64 DFMA per iteration
        const double warp_dfma_per_smsp = (double)blocks * threads / 32.0 / (sms * 4.0) * 64.0 * n_iter;
        out_cycles[v] = best * 1e-3 * sm_mhz * 1e6 / warp_dfma_per_smsp;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return PTMC_OK;
}

double measure_fp64_peak(cudaStream_t st, int iters) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -1;
    const int threads = 256, blocks = sms * 8, n_iter = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(d, 64, 0.999999, 1e-9);   // warm-up
    double best = 0;
    for (int it = 0; it < iters; ++it) {
        cudaEventRecord(e0, st);
        dfma_peak_kernel<<<blocks, threads, 0, st>>>(d, n_iter, 0.999999, 1e-9);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 64.0 * n_iter * (double)threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

}  // namespace ptmc
