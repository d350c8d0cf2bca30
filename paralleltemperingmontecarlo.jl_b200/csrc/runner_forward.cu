// runner_forward.cu — drop-in for the `forward` symbol of librunnerjulia.so, the only FFI the reference has
// (MachineLearningPotential/ForwardPass.jl:60-83; semantics recovered in SURVEY A.2).
//
//   for l = 2..n_layers:  Z = b_l 1^T + W_l A   (W_l column-major n_l x n_{l-1});  A = act_l(Z)
//   eatom[b] = A[0][b] after the last layer;  act code 1 = linear, 2 = tanh.
//
// FP64 throughout.  One thread per atom (batch column); activations ping-pong through a global scratch
// laid out [node][batch] so that loads and stores of a warp are coalesced; the weights are read through
// the read-only path (every thread reads the same address -> broadcast).  The sweep kernels do not call
// this entry point: it exists so that `lib_path()` of the reference can simply point at libptmc_b200.so.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ptmc_b200.h"

namespace {

struct FwdDesc {
    int n_layers, batch, max_nodes;
    int nodes[16];
    int act[16];
};

__global__ void forward_kernel(const double* __restrict__ input, const double* __restrict__ params, double* __restrict__ scratch,
                               double* __restrict__ eatom, const FwdDesc d) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.batch) return;
    const size_t plane = (size_t)d.max_nodes * d.batch;
    double* cur = scratch;
    double* nxt = scratch + plane;
    for (int i = 0; i < d.nodes[0]; ++i) cur[(size_t)i * d.batch + b] = input[(size_t)b * d.nodes[0] + i];
    size_t off = 0;
    for (int l = 1; l < d.n_layers; ++l) {
        const int n_in = d.nodes[l - 1], n_out = d.nodes[l];
        const double* W = params + off;
        const double* bias = W + (size_t)n_out * n_in;
        for (int o = 0; o < n_out; ++o) {
            double z = __ldg(bias + o);
            for (int i = 0; i < n_in; ++i) z = fma(__ldg(W + o + (size_t)i * n_out), cur[(size_t)i * d.batch + b], z);
            nxt[(size_t)o * d.batch + b] = d.act[l] == 2 ? tanh(z) : z;
        }
        off += (size_t)n_out * n_in + n_out;
        double* t = cur; cur = nxt; nxt = t;
    }
    eatom[b] = cur[b];
}

struct Scratch {   // per host thread: forward must be re-entrant (the reference calls it from many tasks)
    double* buf = nullptr;
    size_t cap = 0;
    ~Scratch() { if (buf) cudaFree(buf); }
    double* get(size_t n) {
        if (n > cap) {
            if (buf) cudaFree(buf);
            buf = nullptr; cap = 0;
            if (cudaMalloc(&buf, n * sizeof(double)) != cudaSuccess) return nullptr;
            cap = n;
        }
        return buf;
    }
};
thread_local Scratch g_scratch;

}  // namespace

extern "C" PTMC_API double forward(const double* input, const int32_t* n_in, const int32_t* n_layers, const int32_t* nodes,
                                   const int32_t* batch, const int32_t* n_params, const double* params, const int32_t* act,
                                   double* eatom) {
    const double nan = __builtin_nan("");
    if (!input || !n_in || !n_layers || !nodes || !batch || !n_params || !params || !act || !eatom) return nan;
    FwdDesc d;
    d.n_layers = *n_layers; d.batch = *batch; d.max_nodes = 0;
    if (d.n_layers < 2 || d.n_layers > 16 || d.batch < 1 || nodes[0] != *n_in) return nan;
    size_t need = 0;
    for (int l = 0; l < d.n_layers; ++l) {
        d.nodes[l] = nodes[l]; d.act[l] = act[l];
        if (nodes[l] < 1) return nan;
        if (nodes[l] > d.max_nodes) d.max_nodes = nodes[l];
        if (l) need += (size_t)nodes[l] * nodes[l - 1] + nodes[l];
    }
    if ((size_t)*n_params < need) return nan;
    const size_t n_inp = (size_t)d.nodes[0] * d.batch, plane = (size_t)d.max_nodes * d.batch;
    double* buf = g_scratch.get(n_inp + need + 2 * plane + d.batch);
    if (!buf) return nan;   // no CPU fallback: without a device the call fails
    double* d_in = buf; double* d_par = d_in + n_inp; double* d_scr = d_par + need; double* d_out = d_scr + 2 * plane;
    cudaStream_t st = cudaStreamPerThread;
    if (cudaMemcpyAsync(d_in, input, n_inp * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return nan;
    if (cudaMemcpyAsync(d_par, params, need * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return nan;
    const int threads = 64;
    forward_kernel<<<(d.batch + threads - 1) / threads, threads, 0, st>>>(d_in, d_par, d_scr, d_out, d);
    if (cudaMemcpyAsync(eatom, d_out, d.batch * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return nan;
    if (cudaStreamSynchronize(st) != cudaSuccess) return nan;
    return eatom[d.batch - 1];   // the Fortran function's return register holds the last atom's energy
}
