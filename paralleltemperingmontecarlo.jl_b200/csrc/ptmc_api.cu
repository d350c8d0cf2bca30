// ptmc_api.cu — the C ABI of include/ptmc_b200.h: handle lifetime, state in/out, launches.
// Host side only; all arithmetic of the sweep happens in sweep_kernel.cuh.  No CPU fallback anywhere:
// every entry point either runs on the CUDA device of the handle or fails with PTMC_ERR_CUDA.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cmath>

#include "runner_dev.cuh"
#include "sweep_kernel.cuh"

namespace ptmc {
sweep_fn sweep_table_g2(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads);
sweep_fn sweep_table_g4(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads);
sweep_fn sweep_table_g8(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads);
sweep_fn sweep_table_g16(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads);
sweep_fn sweep_table_g32(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads);
int sweep_max_threads(int G);
void launch_init(const DevParams& P, i64 n_rep, cudaStream_t st);
void launch_reduce_hist(const u64* hist, u64* out, i64 n_ladders, i64 per, cudaStream_t st);
void launch_rdf(const DevParams& P, u64* rdf, double delta_r2, int n_rdf_bins, i64 n_rep, cudaStream_t st);
void launch_lut_thresholds(const DevParams& P, double* thr, cudaStream_t st);
double measure_fp64_peak(cudaStream_t st, int iters);
int run_issue_probe(double sm_mhz, double out_cycles[4]);
int run_pair_probe(const DevParams& P, int n, const double* h_d2n, const double* h_d2o, double* h_en, double* h_eo);
size_t runner_smem_bytes(int N, const RunnerDev& R);
cudaError_t runner_prepare(int N, const RunnerDev& R);
void runner_launch_init(const DevParams& P, const RunnerDev& R, i64 n_rep, cudaStream_t st);
int runner_launch_cycle(const DevParams& P, const RunnerDev& R, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st);
void runner_launch_gather(const DevParams& P, const RunnerDev& R, i64 n_rep, cudaStream_t st);
void runner_launch_export(const DevParams& P, const RunnerDev& R, i64 rep, double* g_out, double* e_out, cudaStream_t st);
size_t runner2_smem_bytes(int N, const Runner2Dev& R);
cudaError_t runner2_prepare(int N, const Runner2Dev& R);
void runner2_launch_init(const DevParams& P, const Runner2Dev& R, i64 n_rep, cudaStream_t st);
int runner2_launch_cycle(const DevParams& P, const Runner2Dev& R, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st);
void runner2_launch_export(const DevParams& P, const Runner2Dev& R, i64 rep, double* g_out, double* e_out, cudaStream_t st);
}  // namespace ptmc

using namespace ptmc;

static thread_local std::string g_create_error;

struct ptmc_handle {
    ptmc_config cfg;
    std::vector<double> betas;
    DevParams P;
    RunnerDev R;
    Runner2Dev R2;
    bool runner = false;             // one of the RuNNer kernels (one CTA per replica, two launches per cycle)
    bool runner2 = false;            // ... the two-species one
    int device = 0;
    int G = 0;
    int threads = 0;
    size_t smem = 0;                 // dynamic shared memory per CTA: smem_k[dbg] for the sweep kernels
    size_t smem_k[2] = {0, 0};
    int geom = 0;
    cudaStream_t stream = nullptr;   // nullptr = legacy default stream
    i64 n_rep = 0, n_steps = 0;
    size_t hist_per_ladder = 0;      // counters per ladder
    // device buffers (owned)
    std::vector<void*> owned;
    u64* d_reduced = nullptr;
    u64* d_rdf = nullptr; double delta_r2 = 0.0;    // radial distribution histogram [L][T][5 n_bin]
    int* d_error = nullptr;
    void* d_slot_draws = nullptr; void* d_ladder_draws = nullptr;
    void* d_trace[4] = {nullptr, nullptr, nullptr, nullptr};
    i64 trace_cap = 0, trace_n = 0;
    bool uploaded = false, hist_ready = false, seeded = false;
    i64 launches = 0;
    std::string err;
};

// ------------------------------------------------------------------------------------------ helpers --
static int fail(ptmc_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
#define CU(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return fail(h, PTMC_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

template <class T>
static int dev_alloc(ptmc_handle* h, T** p, size_t n) {
    void* q = nullptr;
    CU(h, cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    CU(h, cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    h->owned.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
}
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

// SHA-256 of a short message: Julia's Random.seed!(n) hashes the seed's little-endian UInt32 words and
// reads the digest back as four little-endian UInt64.
static void sha256_block(const uint8_t* msg, size_t len, uint8_t out[32]) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    auto rr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
    uint8_t blk[64] = {0};
    memcpy(blk, msg, len);
    blk[len] = 0x80;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) blk[63 - i] = (uint8_t)(bits >> (8 * i));
    uint32_t w[64];
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)blk[4 * i] << 24 | (uint32_t)blk[4 * i + 1] << 16 | (uint32_t)blk[4 * i + 2] << 8 | blk[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t v[8];
    memcpy(v, H, sizeof v);
    for (int i = 0; i < 64; ++i) {
        uint32_t t1 = v[7] + (rr(v[4], 6) ^ rr(v[4], 11) ^ rr(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + K[i] + w[i];
        uint32_t t2 = (rr(v[0], 2) ^ rr(v[0], 13) ^ rr(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
        v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = v[3] + t1; v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t1 + t2;
    }
    for (int i = 0; i < 8; ++i) {
        uint32_t x = H[i] + v[i];
        out[4 * i] = x >> 24; out[4 * i + 1] = x >> 16; out[4 * i + 2] = x >> 8; out[4 * i + 3] = x;
    }
}
static void julia_seed(uint64_t seed, uint64_t s[4]) {
    uint8_t msg[8], dig[32];
    for (int i = 0; i < 8; ++i) msg[i] = (uint8_t)(seed >> (8 * i));
    sha256_block(msg, (seed >> 32) ? 8 : 4, dig);
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | dig[8 * k + i];
        s[k] = v;
    }
}

// Random.seed!(lo + hi * 2^64): the little-endian UInt32 words of the integer, trailing zero words dropped (at least one)
static void julia_seed128(uint64_t lo, uint64_t hi, uint64_t s[4]) {
    uint8_t msg[16], dig[32];
    for (int i = 0; i < 8; ++i) { msg[i] = (uint8_t)(lo >> (8 * i)); msg[8 + i] = (uint8_t)(hi >> (8 * i)); }
    size_t len = 16;
    while (len > 4 && msg[len - 1] == 0 && msg[len - 2] == 0 && msg[len - 3] == 0 && msg[len - 4] == 0) len -= 4;
    sha256_block(msg, len, dig);
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | dig[8 * k + i];
        s[k] = v;
    }
}

static sweep_fn pick_kernel(const ptmc_handle* h, int dbg) {
    const int niter = h->P.n_pad / h->G, nc = h->cfg.n_coeff;
    const int fullw = (h->cfg.n_traj * h->G) % 32 == 0 ? 1 : 0;      // every warp of the ladder fully active
    const int rot = (h->cfg.n_atoms - 1) % h->G == 0 ? 1 : 0;          // the other N - 1 atoms fill the lanes exactly
    switch (h->G) {
        case 2: return sweep_table_g2(h->cfg.pot_kind, h->geom, dbg, niter, nc, fullw, rot, h->threads);
        case 4: return sweep_table_g4(h->cfg.pot_kind, h->geom, dbg, niter, nc, fullw, rot, h->threads);
        case 8: return sweep_table_g8(h->cfg.pot_kind, h->geom, dbg, niter, nc, fullw, rot, h->threads);
        case 16: return sweep_table_g16(h->cfg.pot_kind, h->geom, dbg, niter, nc, fullw, rot, h->threads);
        default: return sweep_table_g32(h->cfg.pot_kind, h->geom, dbg, niter, nc, fullw, rot, h->threads);
    }
}
// words of generator state per slot
static int rng_words(const ptmc_handle* h) { return h->cfg.rng_mode == PTMC_RNG_BATCH8 ? 4 * PTMC_RNG_LANES : 4; }

// RuNNerPotential -> RunnerDev: keep the evaluated symmetry functions only (type 11 / 111), find the distinct eta and
// (lambda, zeta) values of the angular ones, sort those by (lambda, zeta) so that the kernel computes each angular
// factor once, and restrict the first network layer to the evaluated rows.
static int build_runner(ptmc_handle* h, const ptmc_runner_desc* d, std::vector<double>& net_host) {
    RunnerDev& R = h->R;
    memset(&R, 0, sizeof R);
    if (!d->rad_eta || !d->rad_type || !d->rad_gmax_gmin || !d->ang_eta || !d->ang_lambda || !d->ang_zeta || !d->ang_type ||
        !d->ang_gmax_gmin || !d->num_nodes || !d->activation || !d->params)
        return fail(h, PTMC_ERR_ARG, "ptmc_runner_desc has null tables");
    R.n_rad = d->n_rad; R.n_ang = d->n_ang; R.n_sym = d->n_rad + d->n_ang;
    R.r_cut = d->r_cut; R.boundary2 = d->boundary2;
    for (int k = 0; k < d->n_rad; ++k) {
        if (d->rad_type[k] != 11) continue;
        if (R.n_act_rad == RN_MAX_RAD) return fail(h, PTMC_ERR_ARG, "more than %d evaluated radial symmetry functions", RN_MAX_RAD);
        const int q = R.n_act_rad++;
        const double gmax = d->rad_gmax_gmin[2 * k], gmin = d->rad_gmax_gmin[2 * k + 1];
        R.rad_eta[q] = d->rad_eta[k];
        R.rad_norm[q] = 1 / (gmax - gmin);                    // SymmFunc.jl:40-44
        R.rad_off[q] = -gmin * R.rad_norm[q];
        R.act_row[q] = k;
    }
    struct Ang { int e, z, row, init; double tpz, off; };
    std::vector<Ang> ang;
    for (int a = 0; a < d->n_ang; ++a) {
        if (d->ang_type[a] != 111) continue;
        int e = -1, z = -1;
        for (int k = 0; k < R.n_eta; ++k) if (R.eta[k] == d->ang_eta[a]) e = k;
        if (e < 0) {
            if (R.n_eta == RN_MAX_ETA) return fail(h, PTMC_ERR_ARG, "more than %d distinct eta among the angular symmetry functions", RN_MAX_ETA);
            e = R.n_eta++; R.eta[e] = d->ang_eta[a];
        }
        for (int k = 0; k < R.n_lz; ++k) if (R.lz_lambda[k] == d->ang_lambda[a] && R.lz_zeta[k] == d->ang_zeta[a]) z = k;
        if (z < 0) {
            if (R.n_lz == RN_MAX_LZ) return fail(h, PTMC_ERR_ARG, "more than %d distinct (lambda, zeta) among the angular symmetry functions", RN_MAX_LZ);
            z = R.n_lz++; R.lz_lambda[z] = d->ang_lambda[a]; R.lz_zeta[z] = d->ang_zeta[a];
            const double ze = d->ang_zeta[a];
            int zi = (ze >= 1 && ze <= 1024 && ze == std::floor(ze)) ? (int)ze : 0;
            if (zi && (zi & (zi - 1))) zi = 0;                // repeated squaring handles powers of two; others use pow()
            R.lz_zint[z] = zi;
            int nsq = 15;
            if (zi) { nsq = 0; while ((1 << nsq) < zi) ++nsq; }
            if (nsq > 5) nsq = 15;
            R.lz_nsq[z] = nsq;
        }
        const double gmax = d->ang_gmax_gmin[2 * a], gmin = d->ang_gmax_gmin[2 * a + 1];
        const double g_norm = 1 / (gmax - gmin);              // SymmFunc.jl:107-112
        Ang x; x.e = e; x.z = z; x.row = d->n_rad + a; x.off = -gmin * g_norm; x.tpz = std::pow(2.0, 1 - d->ang_zeta[a]) * g_norm;
        x.init = (d->n_rad + a + 1) <= d->n_ang ? 1 : 0;      // total_symm_calc loops (Nrad+1):Nang, SymmFunc.jl:504
        ang.push_back(x);
    }
    if ((int)ang.size() > RN_MAX_ANG) return fail(h, PTMC_ERR_ARG, "more than %d evaluated angular symmetry functions", RN_MAX_ANG);
    std::stable_sort(ang.begin(), ang.end(), [](const Ang& a, const Ang& b) { return a.z < b.z; });
    R.n_act_ang = (int)ang.size();
    R.n_act = R.n_act_rad + R.n_act_ang;
    for (int k = 0; k < RN_MAX_ETA * RN_MAX_LZ; ++k) R.qmap[k] = -1;
    for (int q = 0; q < R.n_act_ang; ++q) {
        R.ang_eta_idx[q] = ang[q].e; R.ang_lz_idx[q] = ang[q].z; R.ang_tpz[q] = ang[q].tpz; R.ang_off[q] = ang[q].off;
        R.ang_init[q] = ang[q].init; R.act_row[R.n_act_rad + q] = ang[q].row;
        const int bit = ang[q].e * RN_MAX_LZ + ang[q].z;
        if (R.present >> bit & 1ull) return fail(h, PTMC_ERR_ARG, "two evaluated angular symmetry functions share (eta, lambda, zeta)");
        R.present |= 1ull << bit;
        R.qmap[bit] = (signed char)q;
    }
    if (R.n_act < 1) return fail(h, PTMC_ERR_ARG, "no symmetry function of type 11 / 111");
    if (d->n_layers != 4 || d->num_nodes[0] != R.n_sym || d->num_nodes[3] != 1 || d->num_nodes[1] < 1 || d->num_nodes[2] < 1 ||
        d->num_nodes[1] > RN_MAX_HID || d->num_nodes[2] > RN_MAX_HID)
        return fail(h, PTMC_ERR_ARG, "the sweep kernel supports networks n_sym-h1-h2-1 with h <= %d (forward() handles any topology)", RN_MAX_HID);
    R.h1 = d->num_nodes[1]; R.h2 = d->num_nodes[2];
    R.act1 = d->activation[1]; R.act2 = d->activation[2]; R.act3 = d->activation[3];
    const size_t need = (size_t)R.h1 * R.n_sym + R.h1 + (size_t)R.h2 * R.h1 + R.h2 + R.h2 + 1;
    if ((size_t)d->n_params < need) return fail(h, PTMC_ERR_ARG, "n_params %d < %zu", d->n_params, need);
    const double* W1 = d->params; const double* b1 = W1 + (size_t)R.h1 * R.n_sym;
    net_host.clear();
    for (int q = 0; q < R.n_act; ++q) for (int o = 0; o < R.h1; ++o) net_host.push_back(W1[o + (size_t)R.act_row[q] * R.h1]);
    net_host.insert(net_host.end(), b1, b1 + (need - (size_t)R.h1 * R.n_sym));     // b1, W2, b2, W3, b3 unchanged
    return 0;
}

// RuNNerPotential2Atom -> Runner2Dev: distinct eta and (lambda, zeta) values of the angular functions, scaling factors per species
// and row (RadialType2a / AngularType3a, SymmFunc.jl:64-83,133-171), the two networks back to back.
static int build_runner2(ptmc_handle* h, const ptmc_runner2_desc* d, std::vector<double>& net_host) {
    Runner2Dev& R = h->R2;
    memset(&R, 0, sizeof R);
    if (!d->rad_eta || !d->ang_eta || !d->ang_lambda || !d->ang_zeta || !d->gmax_gmin_1 || !d->gmax_gmin_2 || !d->num_nodes ||
        !d->activation || !d->params_1 || !d->params_2)
        return fail(h, PTMC_ERR_ARG, "ptmc_runner2_desc has null tables");
    if (d->n_rad < 1 || d->n_rad > RN_MAX_RAD || d->n_ang < 1 || d->n_ang > RN_MAX_ANG)
        return fail(h, PTMC_ERR_ARG, "the two-species sweep handles 1..%d radial and 1..%d angular symmetry functions", RN_MAX_RAD, RN_MAX_ANG);
    if (d->n1 < 1 || d->n2 < 1 || d->n1 + d->n2 != h->cfg.n_atoms) return fail(h, PTMC_ERR_ARG, "n1, n2 >= 1 and n1 + n2 == n_atoms required");
    R.n_rad = d->n_rad; R.n_ang = d->n_ang; R.n_sym = 2 * d->n_rad + 3 * d->n_ang; R.n1 = d->n1;
    R.r_cut = d->r_cut;
    for (int k = 0; k < d->n_rad; ++k) {
        R.rad_eta[k] = d->rad_eta[k];
        for (int q = 0; q < 4; ++q) {
            const double* tab = q < 2 ? d->gmax_gmin_1 : d->gmax_gmin_2;
            const int row = 2 * k + (q & 1);
            const double g_norm = 1 / (tab[2 * row] - tab[2 * row + 1]);
            R.rad_norm[k][q] = g_norm; R.rad_off[k][q] = -tab[2 * row + 1] * g_norm;
        }
    }
    for (int k = 0; k < RN_MAX_ETA * RN_MAX_LZ; ++k) R.qmap[k] = -1;
    for (int a = 0; a < d->n_ang; ++a) {
        int e = -1, z = -1;
        for (int k = 0; k < R.n_eta; ++k) if (R.eta[k] == d->ang_eta[a]) e = k;
        if (e < 0) {
            if (R.n_eta == RN_MAX_ETA) return fail(h, PTMC_ERR_ARG, "more than %d distinct eta among the angular symmetry functions", RN_MAX_ETA);
            e = R.n_eta++; R.eta[e] = d->ang_eta[a];
        }
        for (int k = 0; k < R.n_lz; ++k) if (R.lz_lambda[k] == d->ang_lambda[a] && R.lz_zeta[k] == d->ang_zeta[a]) z = k;
        if (z < 0) {
            if (R.n_lz == RN_MAX_LZ) return fail(h, PTMC_ERR_ARG, "more than %d distinct (lambda, zeta) among the angular symmetry functions", RN_MAX_LZ);
            z = R.n_lz++; R.lz_lambda[z] = d->ang_lambda[a]; R.lz_zeta[z] = d->ang_zeta[a];
            const double ze = d->ang_zeta[a];
            int zi = (ze >= 1 && ze <= 1024 && ze == std::floor(ze)) ? (int)ze : 0;
            if (zi && (zi & (zi - 1))) zi = 0;
            R.lz_zint[z] = zi;
            int nsq = 15;
            if (zi) { nsq = 0; while ((1 << nsq) < zi) ++nsq; }
            if (nsq > 5) nsq = 15;
            R.lz_nsq[z] = nsq;
        }
        R.ang_eta_idx[a] = e; R.ang_lz_idx[a] = z;
        for (int q = 0; q < 6; ++q) {
            const double* tab = q < 3 ? d->gmax_gmin_1 : d->gmax_gmin_2;
            const int row = 2 * d->n_rad + 3 * a + q % 3;
            const double g_norm = 1 / (tab[2 * row] - tab[2 * row + 1]);
            R.ang_off[a][q] = -tab[2 * row + 1] * g_norm;
            R.ang_tpz[a][q] = std::pow(2.0, 1 - d->ang_zeta[a]) * g_norm;
        }
        const int bit = e * RN_MAX_LZ + z;
        if (R.present >> bit & 1ull) return fail(h, PTMC_ERR_ARG, "two angular symmetry functions share (eta, lambda, zeta)");
        R.present |= 1ull << bit;
        R.qmap[bit] = (signed char)a;
    }
    if (d->n_layers != 4 || d->num_nodes[0] != R.n_sym || d->num_nodes[3] != 1 || d->num_nodes[1] < 1 || d->num_nodes[2] < 1 ||
        d->num_nodes[1] > RN_MAX_HID || d->num_nodes[2] > RN_MAX_HID)
        return fail(h, PTMC_ERR_ARG, "the sweep kernel supports networks n_sym-h1-h2-1 with h <= %d (forward() handles any topology)", RN_MAX_HID);
    R.h1 = d->num_nodes[1]; R.h2 = d->num_nodes[2];
    R.act1 = d->activation[1]; R.act2 = d->activation[2]; R.act3 = d->activation[3];
    const size_t need = (size_t)R.h1 * R.n_sym + R.h1 + (size_t)R.h2 * R.h1 + R.h2 + R.h2 + 1;
    if ((size_t)d->n_params < need) return fail(h, PTMC_ERR_ARG, "n_params %d < %zu", d->n_params, need);
    R.net_len = (int)need;
    net_host.assign(d->params_1, d->params_1 + need);
    net_host.insert(net_host.end(), d->params_2, d->params_2 + need);
    return 0;
}

// lanes per replica: enough lanes to keep the per-move serial chain short, few enough that the redundant
// per-group scalar work (RNG, Metropolis) stays small and many replicas are resident.
static int auto_lanes(int n_atoms, int n_traj, bool elj_even_sphere) {
    int g = n_atoms <= 20 ? 4 : (n_atoms <= 72 ? 8 : (n_atoms <= 144 ? 16 : 32));
    // free-space ELJEven ladders of at most 128 threads: 4 lanes per replica, 8 replicas per warp (sweep_kernel_g4s)
    if (elj_even_sphere && n_atoms <= 64 && n_traj * 4 <= 128) g = 4;
    // ... small clusters whose ladder fills whole warps on 2 lanes and whose other atoms split evenly (rotated order): 16 replicas per warp
    if (elj_even_sphere && n_atoms <= 34 && (n_atoms - 1) % 2 == 0 && n_atoms >= 5 && (n_traj * 2) % 32 == 0 && n_traj * 2 <= 128) g = 2;
    while (g > 4 && ((n_traj * g + 31) / 32) * 32 > sweep_max_threads(g)) g /= 2;
    return g;
}

// ------------------------------------------------------------------------------------------ lifetime --
extern "C" PTMC_API int ptmc_abi_version(void) { return PTMC_ABI_VERSION; }

extern "C" PTMC_API const char* ptmc_last_error(const ptmc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" PTMC_API void ptmc_destroy(ptmc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (void* p : h->owned) cudaFree(p);
    if (h->d_slot_draws) cudaFree(h->d_slot_draws);
    if (h->d_ladder_draws) cudaFree(h->d_ladder_draws);
    for (void* p : h->d_trace) if (p) cudaFree(p);
    delete h;
}

extern "C" PTMC_API int ptmc_create(const ptmc_config* c, ptmc_handle** out) {
    if (!c || !out) return fail(nullptr, PTMC_ERR_ARG, "null argument");
    *out = nullptr;
    if (c->abi_version != PTMC_ABI_VERSION) return fail(nullptr, PTMC_ERR_ARG, "abi_version %d != %d", c->abi_version, PTMC_ABI_VERSION);
    if (c->n_ladders < 1 || c->n_traj < 2 || c->n_atoms < 2) return fail(nullptr, PTMC_ERR_ARG, "need n_ladders >= 1, n_traj >= 2, n_atoms >= 2");
    if (!c->betas) return fail(nullptr, PTMC_ERR_ARG, "betas is null");
    const bool nnvt = c->ensemble == PTMC_ENS_NNVT;
    if ((c->pot_kind == PTMC_POT_RUNNER2) != nnvt)
        return fail(nullptr, PTMC_ERR_ARG, "NNVT (atom swaps between two species) goes with RuNNerPotential2Atom and vice versa (src/EnergyEvaluation.jl:1437-1448)");
    if (nnvt && c->n_atom_swaps < 0) return fail(nullptr, PTMC_ERR_ARG, "n_atom_swaps must be >= 0");
    if (c->n_atom_moves < 1 || c->n_atom_moves > c->n_atoms) return fail(nullptr, PTMC_ERR_ARG, "n_atom_moves must be in 1..n_atoms");
    if (c->n_adjust < 1 || c->mc_sample < 1 || c->n_bin < 2) return fail(nullptr, PTMC_ERR_ARG, "n_adjust, mc_sample >= 1 and n_bin >= 2 required");
    if (c->rng_mode < PTMC_RNG_XOSHIRO || c->rng_mode > PTMC_RNG_BATCH8) return fail(nullptr, PTMC_ERR_ARG, "unknown rng_mode %d", c->rng_mode);
    const bool npt = c->ensemble == PTMC_ENS_NPT;
    const bool sphere = c->bc_kind == PTMC_BC_SPHERICAL;
    if (c->ensemble != PTMC_ENS_NVT && !npt && !nnvt) return fail(nullptr, PTMC_ERR_ARG, "unknown ensemble %d", c->ensemble);
    if (c->bc_kind < PTMC_BC_SPHERICAL || c->bc_kind > PTMC_BC_RECTANGULAR) return fail(nullptr, PTMC_ERR_ARG, "unknown boundary condition %d", c->bc_kind);
    if (npt && sphere) return fail(nullptr, PTMC_ERR_ARG, "SphericalBC cannot be used in an NPT ensemble");   // src/Ensembles.jl:189
    if (npt && c->n_volume_moves < 1) return fail(nullptr, PTMC_ERR_ARG, "NPT needs n_volume_moves >= 1");
    if (npt && c->separated_volume && c->bc_kind != PTMC_BC_RHOMBIC && c->bc_kind != PTMC_BC_RECTANGULAR)
        return fail(nullptr, PTMC_ERR_ARG, "separated volume moves need RhombicBC or RectangularBC (scale_xy / scale_z, src/MCMoves.jl:189-222)");
    switch (c->pot_kind) {
        case PTMC_POT_ELJ: case PTMC_POT_ELJ_EVEN:
            if (c->n_coeff < 1 || c->n_coeff > PTMC_MAX_COEFF) return fail(nullptr, PTMC_ERR_ARG, "n_coeff must be in 1..%d", PTMC_MAX_COEFF);
            if (npt && c->pot_kind != PTMC_POT_ELJ_EVEN) return fail(nullptr, PTMC_ERR_ARG, "NPT long-range correction exists for ELJPotentialEven, ELJPotentialB and LookupTablePotential only (src/EnergyEvaluation.jl:968-1009)");
            break;
        case PTMC_POT_ELJ_B:
            if (c->n_coeff < 1 || c->n_coeff > PTMC_MAX_COEFF) return fail(nullptr, PTMC_ERR_ARG, "n_coeff must be in 1..%d", PTMC_MAX_COEFF);
            if (!c->eljb_a || !c->eljb_b) return fail(nullptr, PTMC_ERR_ARG, "pot_kind ELJ_B needs ptmc_config.eljb_a / eljb_b");
            break;
        case PTMC_POT_LUT:
            if (!c->lut_table || c->lut_l_dist < 1 || c->lut_l_angle < 91 || !(c->lut_d_dist > 0) || !(c->lut_d_angle > 0))
                return fail(nullptr, PTMC_ERR_ARG, "pot_kind LUT needs lut_table [l_angle >= 91][l_dist >= 1] and positive grid steps");
            // the angle index reaches round(90 / d_angle + 1) (src/EnergyEvaluation.jl:273-283); the reference throws a BoundsError beyond the table
            if (c->lut_l_angle < (int)std::rint(90.0 / c->lut_d_angle + 1.0))
                return fail(nullptr, PTMC_ERR_ARG, "lookup table has %d angle rows, d_angle = %g needs %d", c->lut_l_angle, c->lut_d_angle, (int)std::rint(90.0 / c->lut_d_angle + 1.0));
            break;
        case PTMC_POT_EAM:
            if (!sphere) return fail(nullptr, PTMC_ERR_ARG, "EmbeddedAtomPotential is built for SphericalBC/NVT (scripts/cu55.jl)");
            if (!(c->eam_n > 0 && c->eam_n <= 24 && c->eam_m > 0 && c->eam_m <= 24)) return fail(nullptr, PTMC_ERR_ARG, "EmbeddedAtomPotential exponents n, m must be in (0, 24] (series of eam_invrexp)");
            break;
        case PTMC_POT_RUNNER:
            if (!c->runner) return fail(nullptr, PTMC_ERR_ARG, "pot_kind RUNNER needs ptmc_config.runner");
            if (!sphere || npt) return fail(nullptr, PTMC_ERR_ARG, "RuNNerPotential is built for SphericalBC/NVT (scripts/cu55.jl)");
            if (c->n_atoms > RN_MAX_ATOMS) return fail(nullptr, PTMC_ERR_ARG, "RuNNerPotential sweep handles up to %d atoms", RN_MAX_ATOMS);
            break;
        case PTMC_POT_RUNNER2:
            if (!c->runner2) return fail(nullptr, PTMC_ERR_ARG, "pot_kind RUNNER2 needs ptmc_config.runner2");
            if (!sphere) return fail(nullptr, PTMC_ERR_ARG, "RuNNerPotential2Atom is built for SphericalBC/NNVT (scripts/brass55.jl)");
            if (c->n_atoms > RN_MAX_ATOMS) return fail(nullptr, PTMC_ERR_ARG, "RuNNerPotential2Atom sweep handles up to %d atoms", RN_MAX_ATOMS);
            break;
        default: return fail(nullptr, PTMC_ERR_ARG, "unknown potential %d", c->pot_kind);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, PTMC_ERR_CUDA, "no CUDA device: libptmc_b200 has no CPU fallback");
    if (c->device < 0 || c->device >= ndev) return fail(nullptr, PTMC_ERR_ARG, "device %d out of range (%d visible)", c->device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) return fail(nullptr, PTMC_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, PTMC_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", c->device, prop.major, prop.minor);

    ptmc_handle* h = new ptmc_handle();
    h->cfg = *c;
    h->betas.assign(c->betas, c->betas + c->n_traj);
    h->cfg.betas = h->betas.data();
    h->cfg.runner = nullptr; h->cfg.runner2 = nullptr;
    h->cfg.eljb_a = nullptr; h->cfg.eljb_b = nullptr;      // copied into DevParams below
    h->cfg.lut_table = nullptr;                              // copied to the device below
    h->runner2 = c->pot_kind == PTMC_POT_RUNNER2;
    h->runner = c->pot_kind == PTMC_POT_RUNNER || h->runner2;
    h->device = c->device;
    h->geom = sphere ? GEOM_SPHERE : (c->bc_kind == PTMC_BC_RHOMBIC ? GEOM_RHOMB : GEOM_PBC);
    h->n_rep = c->n_ladders * c->n_traj;
    h->n_steps = c->n_atom_moves + (npt ? c->n_volume_moves : 0) + (nnvt ? c->n_atom_swaps : 0);     // length(MoveStrategy) src/Ensembles.jl:236-282
    int G = c->lanes_per_replica ? c->lanes_per_replica : auto_lanes(c->n_atoms, c->n_traj, c->pot_kind == PTMC_POT_ELJ_EVEN && sphere && c->n_coeff == 6);
    // the pbc216 class (rhombic NPT, uniform volume moves from power sums, 14 atoms per lane on 16 lanes, 256-thread ladders): two replicas
    // per warp, two ladders per SM (sweep_kernel_p16)
    if (!c->lanes_per_replica && c->pot_kind == PTMC_POT_ELJ_EVEN && c->n_coeff == 6 && c->bc_kind == PTMC_BC_RHOMBIC && npt && !c->separated_volume &&
        (c->n_atoms + 15) / 16 == 14 && c->n_traj * 16 <= 256) G = 16;
    if (h->runner) G = 32;          // the RuNNer kernels use one 256-thread CTA per replica; G is not used
    if (G != 2 && G != 4 && G != 8 && G != 16 && G != 32) { delete h; return fail(nullptr, PTMC_ERR_ARG, "lanes_per_replica must be 0, 2, 4, 8, 16 or 32"); }
    if (G == 2 && !(c->pot_kind == PTMC_POT_ELJ_EVEN && sphere)) { delete h; return fail(nullptr, PTMC_ERR_ARG, "2 lanes per replica are built for free-space ELJPotentialEven ladders only"); }
    h->G = G;
    h->threads = ((c->n_traj * G + 31) / 32) * 32;
    if (!h->runner && h->threads > sweep_max_threads(G)) { int mt = sweep_max_threads(G); delete h; return fail(nullptr, PTMC_ERR_ARG, "n_traj * lanes_per_replica = %d exceeds %d threads per ladder", c->n_traj * G, mt); }

    auto bail = [&](int rc) { g_create_error = h->err; ptmc_destroy(h); return rc; };
    if (cudaSetDevice(h->device) != cudaSuccess) return bail(fail(h, PTMC_ERR_CUDA, "cudaSetDevice failed"));

    DevParams& P = h->P;
    memset(&P, 0, sizeof P);
    P.n_traj = c->n_traj; P.n_atoms = c->n_atoms; P.n_pad = ((c->n_atoms + G - 1) / G) * G;
    P.moved_words = (c->n_atoms + 31) / 32;
    P.n_atom_moves = c->n_atom_moves; P.n_volume_moves = npt ? c->n_volume_moves : 0; P.n_steps = (int)h->n_steps;
    P.ensemble = c->ensemble; P.bc_kind = c->bc_kind; P.pot_kind = c->pot_kind; P.n_coeff = c->n_coeff;
    P.n_adjust = c->n_adjust; P.mc_sample = c->mc_sample; P.n_bin = c->n_bin;
    P.replay = c->rng_mode == PTMC_RNG_REPLAY;
    P.separated = npt && c->separated_volume ? 1 : 0;
    P.min_acc = c->min_acc; P.max_acc = c->max_acc; P.pressure = c->pressure;
    for (int i = 0; i < PTMC_MAX_COEFF; ++i) P.coeff[i] = i < c->n_coeff ? c->coeff[i] : 0.0;
    if (c->pot_kind == PTMC_POT_ELJ_B)
        for (int i = 0; i < c->n_coeff; ++i) { P.eljb_ca[i] = c->coeff[i] * c->eljb_a[i]; P.eljb_cb[i] = c->coeff[i] * c->eljb_b[i]; }
    P.eam_hn = c->eam_n / 2; P.eam_hm = c->eam_m / 2; P.eam_ean = c->eam_ean; P.eam_ecam = c->eam_ecam;

    std::vector<double> net_host;
    if (h->runner) { int rr = h->runner2 ? build_runner2(h, c->runner2, net_host) : build_runner(h, c->runner, net_host); if (rr) return bail(rr); }
    // periodic NPT with uniform volume moves: tie table instead of the N x N matrix (n_atoms <= 256: partner indices are bytes)
    P.use_ties = (!sphere && npt && !P.separated && c->n_atoms <= 256) ? 1 : 0;
    {
        // production free-space ELJEven{6} ladders on 4 lanes with an even number of atoms per lane: blocked atom ownership + LDS.128
        // (BLK in sweep_body) wants the configuration buffers 8 doubles (mod 16) apart
        const int niter = P.n_pad / G;
        const bool rot_kernel = (c->n_atoms - 1) % G == 0 && (c->n_traj * G) % 32 == 0 && niter >= 2;
        P.cfg_mod = (G == 4 && sphere && c->pot_kind == PTMC_POT_ELJ_EVEN && c->n_coeff == 6 && niter % 2 == 0 && !rot_kernel) ? 8 : -1;
    }
    for (int dbg = 0; dbg < 2; ++dbg) {
        // the production ROT kernels (free-space ELJEven / ELJPotentialB, (N - 1) a multiple of G) keep every coordinate array twice: 2 N slots per array
        const bool rot2 = !dbg && sphere && (c->pot_kind == PTMC_POT_ELJ_EVEN || c->pot_kind == PTMC_POT_ELJ_B) && (c->n_atoms - 1) % G == 0;
        h->smem_k[dbg] = smem_layout(c->n_traj, rot2 ? std::max(P.n_pad, 2 * c->n_atoms) : P.n_pad, c->pot_kind, h->threads / 32, !sphere,
                                     !eam_unrolled(P.n_pad / G, dbg != 0), !P.use_ties, G, P.cfg_mod).bytes;
    }
    h->smem = h->runner2 ? runner2_smem_bytes(c->n_atoms, h->R2) : h->runner ? runner_smem_bytes(c->n_atoms, h->R) : std::max(h->smem_k[0], h->smem_k[1]);
    if (h->smem > (size_t)prop.sharedMemPerBlockOptin) return bail(fail(h, PTMC_ERR_ARG, "ladder needs %zu B of shared memory, device offers %zu", h->smem, (size_t)prop.sharedMemPerBlockOptin));

    const size_t R = (size_t)h->n_rep, N = (size_t)c->n_atoms, L = (size_t)c->n_ladders;
    const size_t nb2 = (size_t)c->n_bin + 2;
    h->hist_per_ladder = (size_t)c->n_traj * (npt ? nb2 * nb2 : nb2);
    double* d_betas = nullptr;
    double *en_min, *en_max, *v_min, *v_max;
    int rc = 0;
    if ((rc = dev_alloc(h, &P.pos, R * N * 3)) || (rc = dev_alloc(h, &P.box, R * 2)) || (rc = dev_alloc(h, &P.en_tot, R)) ||
        (rc = dev_alloc(h, &P.comp, c->pot_kind == PTMC_POT_EAM ? R * N * 2 : 1)) ||
        (rc = dev_alloc(h, &P.d2mat, (sphere || P.use_ties) ? 1 : R * N * N)) ||
        (rc = dev_alloc(h, &P.tie_cnt, P.use_ties ? R * N : 1)) || (rc = dev_alloc(h, &P.tie_idx, P.use_ties ? R * N * PTMC_TIE_K : 1)) ||
        (rc = dev_alloc(h, &P.tie_val, P.use_ties ? R * N * PTMC_TIE_K : 1)) ||
        (rc = dev_alloc(h, &P.tan_rho, R)) || (rc = dev_alloc(h, &P.fresh_g, R * ((N + 31) / 32))) || (rc = dev_alloc(h, &P.perm_g, R)) ||
        (rc = dev_alloc(h, &P.moved_g, R * ((N + 31) / 32 + 1))) || (rc = dev_alloc(h, &P.max_displ, R * 4)) ||
        (rc = dev_alloc(h, &P.max_boxlength, R)) || (rc = dev_alloc(h, &P.max_boxheight, R)) || (rc = dev_alloc(h, &P.counters, R * PTMC_NCOUNT)) || (rc = dev_alloc(h, &P.ham, R * 2)) ||
        (rc = dev_alloc(h, &d_betas, (size_t)c->n_traj)) || (rc = dev_alloc(h, &P.slot_rng, R * 4 * PTMC_RNG_LANES)) || (rc = dev_alloc(h, &P.ladder_rng, L * 4)) ||
        (rc = dev_alloc(h, &P.slot_cursor, R)) || (rc = dev_alloc(h, &P.ladder_cursor, L)) ||
        (rc = dev_alloc(h, &en_min, L)) || (rc = dev_alloc(h, &en_max, L)) || (rc = dev_alloc(h, &v_min, L)) || (rc = dev_alloc(h, &v_max, L)) ||
        (rc = dev_alloc(h, &P.hist, L * h->hist_per_ladder)) || (rc = dev_alloc(h, &h->d_reduced, h->hist_per_ladder)) ||
        (rc = dev_alloc(h, &h->d_error, 1)))
        return bail(rc);
    if (h->runner2) {
        Runner2Dev& Rr = h->R2;
        double* d_net = nullptr;
        if ((rc = dev_alloc(h, &Rr.pos_c, R * N * 3)) || (rc = dev_alloc(h, &Rr.en_c, R)) || (rc = dev_alloc(h, &Rr.en_atom_c, R * N)) ||
            (rc = dev_alloc(h, &Rr.g_c, R * N * Rr.n_sym)) || (rc = dev_alloc(h, &Rr.pair_c, R * (N * (N - 1) / 2) * (3 + Rr.n_eta))) ||
            (rc = dev_alloc(h, &d_net, net_host.size())))
            return bail(rc);
        if (cudaMemcpy(d_net, net_host.data(), sizeof(double) * net_host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "copying the networks failed"));
        Rr.net = d_net;
    } else if (h->runner) {
        RunnerDev& Rr = h->R;
        double* d_net = nullptr;
        if ((rc = dev_alloc(h, &Rr.pos_c, R * N * 3)) || (rc = dev_alloc(h, &Rr.en_c, R)) || (rc = dev_alloc(h, &Rr.en_atom_c, R * N)) ||
            (rc = dev_alloc(h, &Rr.g_c, R * N * Rr.n_act)) || (rc = dev_alloc(h, &Rr.pair_c, R * (N * (N - 1) / 2) * (3 + Rr.n_eta))) ||
            (rc = dev_alloc(h, &d_net, net_host.size())))
            return bail(rc);
        if (cudaMemcpy(d_net, net_host.data(), sizeof(double) * net_host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "copying the network failed"));
        Rr.net = d_net;
    }
    if (c->pot_kind == PTMC_POT_EAM) {
        // (r2)^-p tables of eam_invrexp (ptmc_device.cuh)
        std::vector<double> tab(PTMC_EAM_KTAB_DOUBLES + PTMC_EAM_ETAB_DOUBLES, 0.0);
        const long double pn = (long double)P.eam_hn, pm = (long double)P.eam_hm;
        for (int k = 0; k < PTMC_EAM_K; ++k) {
            const double rk = 1.0 / (1.0 + (k + 0.5) / PTMC_EAM_K);
            double* e = tab.data() + (size_t)PTMC_EAM_KSTRIDE * k;
            e[0] = rk; e[1] = (double)powl((long double)rk, pn); e[2] = (double)powl((long double)rk, pm);
        }
        for (int eb = 0; eb < 2048; ++eb) {
            tab[PTMC_EAM_KTAB_DOUBLES + 2 * eb] = (double)powl(2.0L, -pn * (eb - 1023));
            tab[PTMC_EAM_KTAB_DOUBLES + 2 * eb + 1] = (double)powl(2.0L, -pm * (eb - 1023));
        }
        long double cn = 1, cm = 1;
        for (int j = 1; j <= PTMC_EAM_DEG; ++j) {
            cn = cn * (-pn - (j - 1)) / j; cm = cm * (-pm - (j - 1)) / j;      // binom(-p, j)
            P.eam_cn[j - 1] = (double)cn; P.eam_cm[j - 1] = (double)cm;
        }
        double* d_tab = nullptr;
        if ((rc = dev_alloc(h, &d_tab, tab.size()))) return bail(rc);
        if (cudaMemcpy(d_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "copying the EAM power tables failed"));
        P.eam_ktab = d_tab; P.eam_etab = d_tab + PTMC_EAM_KTAB_DOUBLES;
    }
    if (c->pot_kind == PTMC_POT_LUT) {
        double* d_lut = nullptr;
        const size_t nt = (size_t)c->lut_l_dist * c->lut_l_angle;
        if ((rc = dev_alloc(h, &d_lut, nt))) return bail(rc);
        if (cudaMemcpy(d_lut, c->lut_table, sizeof(double) * nt, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "copying the lookup table failed"));
        P.lut_table = d_lut; P.lut_l_dist = c->lut_l_dist;
        P.lut_start = c->lut_start_dist; P.lut_d = c->lut_d_dist; P.lut_dang = c->lut_d_angle; P.lut_c6 = c->lut_c6coeff;
        double* d_thr = nullptr;
        if ((rc = dev_alloc(h, &d_thr, PTMC_LUT_THR + 1))) return bail(rc);
        launch_lut_thresholds(P, d_thr, h->stream);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(fail(h, PTMC_ERR_CUDA, "building the angle-index steps of the lookup table failed"));
        P.lut_thr = d_thr; P.lut_kf = (float)(180.0 / 3.14159265358979323846 / c->lut_d_angle);
    }
    P.betas = d_betas; P.en_min = en_min; P.en_max = en_max; P.v_min = v_min; P.v_max = v_max;
    P.error = h->d_error;
    if (cudaMemcpy(d_betas, h->betas.data(), sizeof(double) * c->n_traj, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(h, PTMC_ERR_CUDA, "copying betas failed"));
    if (h->runner) {
        if ((h->runner2 ? runner2_prepare(c->n_atoms, h->R2) : runner_prepare(c->n_atoms, h->R)) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "RuNNer kernels need %zu B of shared memory: %s", h->smem, cudaGetErrorString(cudaGetLastError())));
    }
    // opt in to large dynamic shared memory for both builds of the kernel
    for (int dbg = 0; dbg < 2 && !h->runner; ++dbg) {
        sweep_fn f = pick_kernel(h, dbg);
        if (!f) return bail(fail(h, PTMC_ERR_ARG, "no kernel for potential %d / boundary %d", c->pot_kind, c->bc_kind));
        if (cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_k[dbg]) != cudaSuccess)
            return bail(fail(h, PTMC_ERR_CUDA, "cudaFuncSetAttribute(%zu B smem) failed: %s", h->smem, cudaGetErrorString(cudaGetLastError())));
        // residency is bounded by shared memory: ask for the largest carve-out instead of the driver's per-kernel guess
        cudaFuncSetAttribute((const void*)f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    *out = h;
    if (c->rng_mode != PTMC_RNG_REPLAY) { int r2 = ptmc_seed_job(h, c->seed, 0); if (r2) { *out = nullptr; return bail(r2); } }
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_set_stream(ptmc_handle* h, void* cuda_stream) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return PTMC_OK;
}

// ------------------------------------------------------------------------------------------ state in --
extern "C" PTMC_API int ptmc_upload_states(ptmc_handle* h, const double* pos, const double* box, const double* max_displ) {
    if (!h || !pos || !box) return h ? fail(h, PTMC_ERR_ARG, "pos and box are required") : PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    const size_t R = (size_t)h->n_rep, N = (size_t)h->cfg.n_atoms;
    const DevParams& P = h->P;
    CU(h, cudaMemcpyAsync(P.pos, pos, sizeof(double) * R * N * 3, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(P.box, box, sizeof(double) * R * 2, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> md(R * 4), mb(R), mh(R);
    for (size_t r = 0; r < R; ++r) {
        for (int k = 0; k < 4; ++k) md[4 * r + k] = max_displ ? max_displ[4 * r + k] : 0.1;             // src/MCStates.jl:104
        mb[r] = h->cfg.bc_kind == PTMC_BC_SPHERICAL ? 30.0 : box[2 * r] * 1.8;                          // src/MCStates.jl:63-75
        mh[r] = h->cfg.bc_kind == PTMC_BC_SPHERICAL ? 30.0 : (h->cfg.bc_kind == PTMC_BC_CUBIC ? box[2 * r] : box[2 * r + 1]) * 1.8;   // :76-87
    }
    CU(h, cudaMemcpyAsync(P.max_boxheight, mh.data(), sizeof(double) * R, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(P.max_displ, md.data(), sizeof(double) * R * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(P.max_boxlength, mb.data(), sizeof(double) * R, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemsetAsync(P.counters, 0, sizeof(i64) * R * PTMC_NCOUNT, h->stream));
    CU(h, cudaMemsetAsync(P.ham, 0, sizeof(double) * R * 2, h->stream));
    CU(h, cudaMemsetAsync(h->d_error, 0, sizeof(int), h->stream));
    {   // slot -> configuration map starts as the identity
        std::vector<int> perm(R);
        for (size_t r = 0; r < R; ++r) perm[r] = (int)(r % (size_t)h->cfg.n_traj);
        CU(h, cudaMemcpyAsync(P.perm_g, perm.data(), sizeof(int) * R, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemsetAsync(P.moved_g, 0, sizeof(int) * R * (P.moved_words + 1), h->stream));
        CU(h, cudaStreamSynchronize(h->stream));   // perm is stack-owned
    }
    if (h->runner2) runner2_launch_init(P, h->R2, h->n_rep, h->stream);
    else if (h->runner) runner_launch_init(P, h->R, h->n_rep, h->stream);
    else launch_init(P, h->n_rep, h->stream);
    h->launches += 1;
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));   // md/mb are stack-owned
    h->uploaded = true;
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_upload_counters(ptmc_handle* h, const int64_t* counters, const double* ham) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    const size_t R = (size_t)h->n_rep;
    if (counters) CU(h, cudaMemcpyAsync(h->P.counters, counters, sizeof(i64) * R * PTMC_NCOUNT, cudaMemcpyHostToDevice, h->stream));
    if (ham) CU(h, cudaMemcpyAsync(h->P.ham, ham, sizeof(double) * R * 2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_reset_counters(ptmc_handle* h) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemsetAsync(h->P.counters, 0, sizeof(i64) * (size_t)h->n_rep * PTMC_NCOUNT, h->stream));
    return PTMC_OK;
}

// ------------------------------------------------------------------------------------ random streams --
extern "C" PTMC_API int ptmc_set_rng_states(ptmc_handle* h, const uint64_t* slot_states, const uint64_t* ladder_states) {
    if (!h || !slot_states || !ladder_states) return h ? fail(h, PTMC_ERR_ARG, "null rng states") : PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpyAsync(h->P.slot_rng, slot_states, sizeof(u64) * (size_t)h->n_rep * rng_words(h), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->P.ladder_rng, ladder_states, sizeof(u64) * (size_t)h->cfg.n_ladders * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->seeded = true;
    return PTMC_OK;
}
extern "C" PTMC_API int ptmc_get_rng_states(ptmc_handle* h, uint64_t* slot_states, uint64_t* ladder_states) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    if (slot_states) CU(h, cudaMemcpyAsync(slot_states, h->P.slot_rng, sizeof(u64) * (size_t)h->n_rep * rng_words(h), cudaMemcpyDeviceToHost, h->stream));
    if (ladder_states) CU(h, cudaMemcpyAsync(ladder_states, h->P.ladder_rng, sizeof(u64) * (size_t)h->cfg.n_ladders * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PTMC_OK;
}
extern "C" PTMC_API int ptmc_seed(ptmc_handle* h, uint64_t seed) {
    if (!h) return PTMC_ERR_ARG;
    const size_t R = (size_t)h->n_rep * (rng_words(h) / 4), L = (size_t)h->cfg.n_ladders;   // R substreams in all
    std::vector<uint64_t> s(R * 4), l(L * 4);
    for (size_t r = 0; r < R; ++r) julia_seed(seed + r, &s[4 * r]);
    for (size_t k = 0; k < L; ++k) julia_seed(seed + R + k, &l[4 * k]);
    return ptmc_set_rng_states(h, s.data(), l.data());
}
extern "C" PTMC_API int ptmc_seed_job(ptmc_handle* h, uint64_t seed, int64_t first_ladder) {
    if (!h) return PTMC_ERR_ARG;
    if (first_ladder < 0) return fail(h, PTMC_ERR_ARG, "first_ladder must be >= 0");
    const size_t W = (size_t)rng_words(h) / 4, T = (size_t)h->cfg.n_traj, L = (size_t)h->cfg.n_ladders;
    std::vector<uint64_t> s(L * T * W * 4), l(L * 4);
    for (size_t a = 0; a < L; ++a) {
        const uint64_t gl = (uint64_t)first_ladder + a;
        for (size_t t = 0; t < T; ++t)
            for (size_t k = 0; k < W; ++k) julia_seed128(seed, (gl * T + t) * W + k + 1, &s[((a * T + t) * W + k) * 4]);
        julia_seed128(seed, (1ull << 63) | gl, &l[4 * a]);
    }
    return ptmc_set_rng_states(h, s.data(), l.data());
}
extern "C" PTMC_API int ptmc_rng_words(const ptmc_handle* h) { return h ? rng_words(h) : 0; }
extern "C" PTMC_API int ptmc_set_replay(ptmc_handle* h, const uint64_t* slot_draws, int64_t slot_len, const uint64_t* ladder_draws, int64_t ladder_len) {
    if (!h || !slot_draws || !ladder_draws || slot_len < 0 || ladder_len < 0) return h ? fail(h, PTMC_ERR_ARG, "bad replay streams") : PTMC_ERR_ARG;
    if (h->cfg.rng_mode != PTMC_RNG_REPLAY) return fail(h, PTMC_ERR_STATE, "ptmc_set_replay needs a handle created with rng_mode = PTMC_RNG_REPLAY");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->d_slot_draws) { cudaFree(h->d_slot_draws); h->d_slot_draws = nullptr; }
    if (h->d_ladder_draws) { cudaFree(h->d_ladder_draws); h->d_ladder_draws = nullptr; }
    const size_t ns = (size_t)h->n_rep * slot_len, nl = (size_t)h->cfg.n_ladders * ladder_len;
    CU(h, cudaMalloc(&h->d_slot_draws, std::max<size_t>(ns, 1) * sizeof(u64)));
    CU(h, cudaMalloc(&h->d_ladder_draws, std::max<size_t>(nl, 1) * sizeof(u64)));
    CU(h, cudaMemcpy(h->d_slot_draws, slot_draws, ns * sizeof(u64), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->d_ladder_draws, ladder_draws, nl * sizeof(u64), cudaMemcpyHostToDevice));
    CU(h, cudaMemset(h->P.slot_cursor, 0, sizeof(i64) * (size_t)h->n_rep));
    CU(h, cudaMemset(h->P.ladder_cursor, 0, sizeof(i64) * (size_t)h->cfg.n_ladders));
    h->P.slot_draws = static_cast<const u64*>(h->d_slot_draws); h->P.slot_len = slot_len;
    h->P.ladder_draws = static_cast<const u64*>(h->d_ladder_draws); h->P.ladder_len = ladder_len;
    h->P.replay = 1;
    h->seeded = true;
    return PTMC_OK;
}

// ---------------------------------------------------------------------------------------- histograms --
extern "C" PTMC_API int ptmc_set_hist_bounds(ptmc_handle* h, const double* en_min, const double* en_max, const double* v_min, const double* v_max, int per_ladder) {
    if (!h || !en_min || !en_max) return h ? fail(h, PTMC_ERR_ARG, "en_min/en_max required") : PTMC_ERR_ARG;
    const bool npt = h->cfg.ensemble == PTMC_ENS_NPT;
    if (npt && (!v_min || !v_max)) return fail(h, PTMC_ERR_ARG, "NPT needs v_min/v_max");
    CU(h, cudaSetDevice(h->device));
    const size_t L = (size_t)h->cfg.n_ladders;
    std::vector<double> a(L), b(L), c(L, 0.0), d(L, 0.0);
    for (size_t l = 0; l < L; ++l) {
        size_t k = per_ladder ? l : 0;
        a[l] = en_min[k]; b[l] = en_max[k];
        if (v_min) c[l] = v_min[k];
        if (v_max) d[l] = v_max[k];
    }
    CU(h, cudaMemcpyAsync(const_cast<double*>(h->P.en_min), a.data(), sizeof(double) * L, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(const_cast<double*>(h->P.en_max), b.data(), sizeof(double) * L, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(const_cast<double*>(h->P.v_min), c.data(), sizeof(double) * L, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(const_cast<double*>(h->P.v_max), d.data(), sizeof(double) * L, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemsetAsync(h->P.hist, 0, sizeof(u64) * L * h->hist_per_ladder, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->hist_ready = true;
    return PTMC_OK;
}

// -------------------------------------------------------------------------------------- the hot path --
extern "C" PTMC_API int ptmc_run_cycles(ptmc_handle* h, int64_t first_idx, int64_t n_cycles, int sample) {
    return ptmc_run_phase(h, first_idx, n_cycles, sample, 0);
}

extern "C" PTMC_API int ptmc_run_phase(ptmc_handle* h, int64_t first_idx, int64_t n_cycles, int sample, int phase) {
    if (!h) return PTMC_ERR_ARG;
    if (phase < 0 || phase > 2) return fail(h, PTMC_ERR_ARG, "phase must be 0 (whole cycle), 1 (moves) or 2 (adaptation + sampling)");
    if (phase && h->runner) return fail(h, PTMC_ERR_ARG, "split cycles are not built for the RuNNer kernels");
    if (phase && h->trace_cap > 0) return fail(h, PTMC_ERR_STATE, "tracing records whole cycles only");
    if (!h->uploaded) return fail(h, PTMC_ERR_STATE, "ptmc_run_cycles before ptmc_upload_states");
    if (!h->seeded) return fail(h, PTMC_ERR_STATE, "no random stream: call ptmc_seed / ptmc_set_rng_states / ptmc_set_replay");
    if (sample && !h->hist_ready) return fail(h, PTMC_ERR_STATE, "sampling before ptmc_set_hist_bounds");
    if (n_cycles < 0 || first_idx < 0) return fail(h, PTMC_ERR_ARG, "negative cycle count / index");
    const bool tracing = h->trace_cap > 0;
    if (tracing && h->trace_n + n_cycles > h->trace_cap) return fail(h, PTMC_ERR_ARG, "trace capacity %lld exceeded", (long long)h->trace_cap);
    CU(h, cudaSetDevice(h->device));
#ifndef PTMC_PROD_TRACE
    if (tracing && h->cfg.rng_mode == PTMC_RNG_BATCH8 && !h->runner)
        return fail(h, PTMC_ERR_STATE, "tracing the production (BATCH8) kernels needs the test build lib/libptmc_b200_trace.so (-DPTMC_PROD_TRACE); "
                                       "this library traces rng_mode XOSHIRO / REPLAY");
#endif
    const int dbg = h->cfg.rng_mode != PTMC_RNG_BATCH8 ? 1 : 0;
    if (h->runner) {
        // one launch pair per cycle: the exchange between replicas happens at the kernel boundary
        DevParams P = h->P;
        P.trace_cap = h->trace_cap; P.trace_base = h->trace_n; P.trace_on = h->trace_cap > 0 ? 1 : 0;
        P.tr_index = static_cast<int32_t*>(h->d_trace[0]); P.tr_de = static_cast<double*>(h->d_trace[1]);
        P.tr_acc = static_cast<uint8_t*>(h->d_trace[2]); P.tr_exc = static_cast<int32_t*>(h->d_trace[3]);
        for (i64 k = 0; k < n_cycles; ++k)
            h->launches += h->runner2 ? runner2_launch_cycle(P, h->R2, h->cfg.n_ladders, first_idx + k, k, sample, dbg != 0, h->stream)
                                      : runner_launch_cycle(P, h->R, h->cfg.n_ladders, first_idx + k, k, sample, dbg != 0, h->stream);
        if (h->runner2) runner_launch_gather_ptr(P, h->R2.pos_c, h->R2.en_c, h->n_rep, h->stream);
        else runner_launch_gather(P, h->R, h->n_rep, h->stream);
        h->launches += 1;
        CU(h, cudaGetLastError());
        if (tracing) h->trace_n += n_cycles;
        return PTMC_OK;
    }
    sweep_fn f = pick_kernel(h, dbg);
    // per-launch acceptance counters are 32-bit: keep moves per launch below 2^31
    const i64 max_chunk = std::max<i64>(1, (i64)0x7fffffff / std::max<i64>(1, h->n_steps));
    i64 done = 0;
    while (done < n_cycles) {
        const i64 k = std::min(max_chunk, n_cycles - done);
        DevParams P = h->P;
        P.phase = phase;
        P.trace_cap = h->trace_cap; P.trace_base = h->trace_n; P.trace_on = h->trace_cap > 0 ? 1 : 0;
        P.tr_index = static_cast<int32_t*>(h->d_trace[0]); P.tr_de = static_cast<double*>(h->d_trace[1]);
        P.tr_acc = static_cast<uint8_t*>(h->d_trace[2]); P.tr_exc = static_cast<int32_t*>(h->d_trace[3]);
        f<<<(unsigned)h->cfg.n_ladders, h->threads, h->smem_k[dbg ? 1 : 0], h->stream>>>(P, first_idx + done, k, sample);
        CU(h, cudaGetLastError());
        h->launches += 1;
        if (tracing) h->trace_n += k;
        done += k;
    }
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_sync(ptmc_handle* h) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    int e = 0;
    CU(h, cudaMemcpyAsync(&e, h->d_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (e == PTMC_ERR_SPHERE) return fail(h, e, "Error: too many moves out of binding sphere");   // src/MCMoves.jl:132
    if (e == PTMC_ERR_REPLAY) return fail(h, e, "replay stream exhausted");
    if (e == PTMC_ERR_TIES) return fail(h, e, "more than %d neighbours of one atom sit on the cut-off sphere of the uploaded configuration", PTMC_TIE_K);
    if (e) return fail(h, e, "device error %d", e);
    return PTMC_OK;
}

// ------------------------------------------------------------------------------------------- trace --
extern "C" PTMC_API int ptmc_trace_enable(ptmc_handle* h, int64_t capacity_cycles) {
    if (!h || capacity_cycles < 0) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    for (void*& p : h->d_trace) { if (p) cudaFree(p); p = nullptr; }
    h->trace_cap = capacity_cycles; h->trace_n = 0;
    if (capacity_cycles > 0) {
        const size_t m = (size_t)h->cfg.n_ladders * capacity_cycles * h->cfg.n_traj * h->n_steps;
        CU(h, cudaMalloc(&h->d_trace[0], m * sizeof(int32_t)));
        CU(h, cudaMalloc(&h->d_trace[1], m * sizeof(double)));
        CU(h, cudaMalloc(&h->d_trace[2], m));
        CU(h, cudaMalloc(&h->d_trace[3], (size_t)h->cfg.n_ladders * capacity_cycles * 3 * sizeof(int32_t)));
        CU(h, cudaMemset(h->d_trace[0], 0, m * sizeof(int32_t)));
        CU(h, cudaMemset(h->d_trace[1], 0, m * sizeof(double)));
        CU(h, cudaMemset(h->d_trace[2], 0, m));
        CU(h, cudaMemset(h->d_trace[3], 0, (size_t)h->cfg.n_ladders * capacity_cycles * 3 * sizeof(int32_t)));
    }
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_trace_download(ptmc_handle* h, int32_t* move_index, double* delta_en, uint8_t* accepted, int32_t* exchange, int64_t* n_cycles_out) {
    if (!h) return PTMC_ERR_ARG;
    TRY(ptmc_sync(h));
    const size_t L = (size_t)h->cfg.n_ladders, per = (size_t)h->cfg.n_traj * h->n_steps;
    const size_t cap = (size_t)h->trace_cap, k = (size_t)h->trace_n;
    if (k > 0) {
        // compact [L][cap][...] -> [L][k][...]
        if (move_index) CU(h, cudaMemcpy2D(move_index, k * per * 4, h->d_trace[0], cap * per * 4, k * per * 4, L, cudaMemcpyDeviceToHost));
        if (delta_en) CU(h, cudaMemcpy2D(delta_en, k * per * 8, h->d_trace[1], cap * per * 8, k * per * 8, L, cudaMemcpyDeviceToHost));
        if (accepted) CU(h, cudaMemcpy2D(accepted, k * per, h->d_trace[2], cap * per, k * per, L, cudaMemcpyDeviceToHost));
        if (exchange) CU(h, cudaMemcpy2D(exchange, k * 12, h->d_trace[3], cap * 12, k * 12, L, cudaMemcpyDeviceToHost));
    }
    if (n_cycles_out) *n_cycles_out = (int64_t)k;
    return PTMC_OK;
}

// ---------------------------------------------------------------------------------------- state out --
extern "C" PTMC_API int ptmc_download_states(ptmc_handle* h, double* pos, double* box, double* en_tot, double* max_displ, int64_t* counters, double* ham) {
    if (!h) return PTMC_ERR_ARG;
    TRY(ptmc_sync(h));
    const size_t R = (size_t)h->n_rep, N = (size_t)h->cfg.n_atoms;
    const DevParams& P = h->P;
    // all copies on the handle's stream (never the legacy default stream: handles on other streams keep running), one wait at the end
    if (pos) CU(h, cudaMemcpyAsync(pos, P.pos, sizeof(double) * R * N * 3, cudaMemcpyDeviceToHost, h->stream));
    if (box) CU(h, cudaMemcpyAsync(box, P.box, sizeof(double) * R * 2, cudaMemcpyDeviceToHost, h->stream));
    if (en_tot) CU(h, cudaMemcpyAsync(en_tot, P.en_tot, sizeof(double) * R, cudaMemcpyDeviceToHost, h->stream));
    if (max_displ) CU(h, cudaMemcpyAsync(max_displ, P.max_displ, sizeof(double) * R * 4, cudaMemcpyDeviceToHost, h->stream));
    if (counters) CU(h, cudaMemcpyAsync(counters, P.counters, sizeof(i64) * R * PTMC_NCOUNT, cudaMemcpyDeviceToHost, h->stream));
    if (ham) CU(h, cudaMemcpyAsync(ham, P.ham, sizeof(double) * R * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_reduced_hist_dev(ptmc_handle* h, void** dev_ptr, int64_t* n_elems) {
    if (!h) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    launch_reduce_hist(h->P.hist, h->d_reduced, h->cfg.n_ladders, (i64)h->hist_per_ladder, h->stream);
    h->launches += 1;
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    if (dev_ptr) *dev_ptr = h->d_reduced;
    if (n_elems) *n_elems = (int64_t)h->hist_per_ladder;
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_download_hist(ptmc_handle* h, double* en_hist, double* ev_hist, int reduce_ladders) {
    if (!h) return PTMC_ERR_ARG;
    TRY(ptmc_sync(h));
    double* dst = h->cfg.ensemble == PTMC_ENS_NPT ? ev_hist : en_hist;
    if (!dst) return PTMC_OK;
    const u64* src = h->P.hist;
    size_t n = (size_t)h->cfg.n_ladders * h->hist_per_ladder;
    if (reduce_ladders) {
        void* p; int64_t m;
        TRY(ptmc_reduced_hist_dev(h, &p, &m));
        src = static_cast<const u64*>(p); n = (size_t)m;
    }
    std::vector<u64> tmp(n);
    CU(h, cudaMemcpyAsync(tmp.data(), src, n * sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) dst[i] = (double)tmp[i];   // Output.en_histogram holds Float64 counts
    return PTMC_OK;
}

// ------------------------------------------------------------------------ configurations in and out --
// everything exc_trajectories! (src/Exchange.jl:115-124) moves between two slots, for SphericalBC handles:
// positions [N][3], box[2], en_tot, and the EAM components [N][2]
extern "C" PTMC_API int64_t ptmc_config_doubles(const ptmc_handle* h) {
    if (!h) return 0;
    const int64_t N = h->cfg.n_atoms;
    return 3 * N + 3 + (h->cfg.pot_kind == PTMC_POT_EAM ? 2 * N : 0);
}
static int config_io(ptmc_handle* h, int64_t ladder, int32_t slot, double* buf, bool out) {
    if (!h || !buf) return PTMC_ERR_ARG;
    if (!h->uploaded) return fail(h, PTMC_ERR_STATE, "no states uploaded");
    if (h->runner || h->cfg.bc_kind != PTMC_BC_SPHERICAL) return fail(h, PTMC_ERR_ARG, "configuration export / import is built for the SphericalBC sweep kernels");
    if (ladder < 0 || ladder >= h->cfg.n_ladders || slot < 0 || slot >= h->cfg.n_traj) return fail(h, PTMC_ERR_ARG, "ladder / slot out of range");
    TRY(ptmc_sync(h));
    const size_t N = (size_t)h->cfg.n_atoms, rep = (size_t)ladder * h->cfg.n_traj + slot;
    const DevParams& P = h->P;
    const cudaMemcpyKind kind = out ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    auto cp = [&](double* dev, double* host, size_t n) { return out ? cudaMemcpy(host, dev, n * sizeof(double), kind) : cudaMemcpy(dev, host, n * sizeof(double), kind); };
    CU(h, cp(P.pos + rep * 3 * N, buf, 3 * N));
    CU(h, cp(P.box + rep * 2, buf + 3 * N, 2));
    CU(h, cp(P.en_tot + rep, buf + 3 * N + 2, 1));
    if (h->cfg.pot_kind == PTMC_POT_EAM) CU(h, cp(P.comp + rep * 2 * N, buf + 3 * N + 3, 2 * N));
    return PTMC_OK;
}
extern "C" PTMC_API int ptmc_export_config(ptmc_handle* h, int64_t ladder, int32_t slot, double* buf) { return config_io(h, ladder, slot, buf, true); }
extern "C" PTMC_API int ptmc_import_config(ptmc_handle* h, int64_t ladder, int32_t slot, const double* buf) {
    return config_io(h, ladder, slot, const_cast<double*>(buf), false);
}

// ----------------------------------------------------------------------------------------------- rdf --
extern "C" PTMC_API int ptmc_rdf_enable(ptmc_handle* h, double delta_r2) {
    if (!h) return PTMC_ERR_ARG;
    if (!(delta_r2 > 0)) return fail(h, PTMC_ERR_ARG, "delta_r2 must be positive");
    CU(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_rep * 5 * (size_t)h->cfg.n_bin;
    if (!h->d_rdf) TRY(dev_alloc(h, &h->d_rdf, n));
    else CU(h, cudaMemsetAsync(h->d_rdf, 0, n * sizeof(u64), h->stream));
    h->delta_r2 = delta_r2;
    return PTMC_OK;
}
extern "C" PTMC_API int ptmc_rdf_sample(ptmc_handle* h) {
    if (!h) return PTMC_ERR_ARG;
    if (!h->d_rdf) return fail(h, PTMC_ERR_STATE, "ptmc_rdf_enable first");
    if (!h->uploaded) return fail(h, PTMC_ERR_STATE, "no states uploaded");
    CU(h, cudaSetDevice(h->device));
    launch_rdf(h->P, h->d_rdf, h->delta_r2, 5 * h->cfg.n_bin, h->n_rep, h->stream);
    h->launches += 1;
    CU(h, cudaGetLastError());
    return PTMC_OK;
}
extern "C" PTMC_API int ptmc_download_rdf(ptmc_handle* h, double* rdf) {
    if (!h || !rdf) return PTMC_ERR_ARG;
    if (!h->d_rdf) return fail(h, PTMC_ERR_STATE, "ptmc_rdf_enable first");
    TRY(ptmc_sync(h));
    const size_t n = (size_t)h->n_rep * 5 * (size_t)h->cfg.n_bin;
    std::vector<u64> tmp(n);
    CU(h, cudaMemcpy(tmp.data(), h->d_rdf, n * sizeof(u64), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) rdf[i] = (double)tmp[i];
    return PTMC_OK;
}

// ------------------------------------------------------------------------------------- introspection --
extern "C" PTMC_API int64_t ptmc_launch_count(const ptmc_handle* h) { return h ? h->launches : 0; }
extern "C" PTMC_API int ptmc_lanes_per_replica(const ptmc_handle* h) { return h ? h->G : 0; }

extern "C" PTMC_API int ptmc_fp64_peak(int device, int iters, double* tflops) {
    if (!tflops) return PTMC_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(nullptr, PTMC_ERR_CUDA, "no such CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, PTMC_ERR_CUDA, "cudaSetDevice failed");
    double v = measure_fp64_peak(nullptr, iters > 0 ? iters : 5);
    if (v <= 0) return fail(nullptr, PTMC_ERR_CUDA, "fp64 peak measurement failed");
    *tflops = v;
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_debug_pair_energy(ptmc_handle* h, int n, const double* d2_new, const double* d2_old, double* e_new, double* e_old) {
    if (!h || n < 1 || !d2_new || !d2_old || !e_new || !e_old) return PTMC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    int rc = run_pair_probe(h->P, n, d2_new, d2_old, e_new, e_old);
    h->launches += 1;
    return rc ? fail(h, rc, "pair probe failed: %s", cudaGetErrorString(cudaGetLastError())) : PTMC_OK;
}

extern "C" PTMC_API int ptmc_debug_runner_state(ptmc_handle* h, int64_t ladder, int32_t slot, double* g_matrix, double* en_atom) {
    if (!h || !h->runner || !g_matrix || !en_atom || ladder < 0 || ladder >= h->cfg.n_ladders || slot < 0 || slot >= h->cfg.n_traj)
        return h ? fail(h, PTMC_ERR_ARG, "ptmc_debug_runner_state: not a RuNNer handle or bad index") : PTMC_ERR_ARG;
    TRY(ptmc_sync(h));
    const size_t N = (size_t)h->cfg.n_atoms, ng = N * (h->runner2 ? h->R2.n_sym : h->R.n_sym);
    double* d = nullptr;
    CU(h, cudaMalloc(&d, sizeof(double) * (ng + N)));
    if (h->runner2) runner2_launch_export(h->P, h->R2, ladder * h->cfg.n_traj + slot, d, d + ng, h->stream);
    else runner_launch_export(h->P, h->R, ladder * h->cfg.n_traj + slot, d, d + ng, h->stream);
    h->launches += 1;
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(g_matrix, d, sizeof(double) * ng, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(en_atom, d + ng, sizeof(double) * N, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, PTMC_ERR_CUDA, "ptmc_debug_runner_state: %s", cudaGetErrorString(e));
    return PTMC_OK;
}

extern "C" PTMC_API int ptmc_issue_probe(int device, double sm_mhz, double* cycles_per_dfma4) {
    if (!cycles_per_dfma4) return PTMC_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(nullptr, PTMC_ERR_CUDA, "no such CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, PTMC_ERR_CUDA, "cudaSetDevice failed");
    return run_issue_probe(sm_mhz, cycles_per_dfma4);
}
