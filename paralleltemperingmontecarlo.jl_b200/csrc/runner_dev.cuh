// runner_dev.cuh — device-side description of a single-species RuNNer potential (RuNNerPotential,
// src/EnergyEvaluation.jl:1182-1218) and of the per-configuration state its sweep keeps in HBM.
#pragma once
#include "ptmc_device.cuh"

namespace ptmc {

constexpr int RN_MAX_RAD = 8;      // evaluated radial functions (type 11)
constexpr int RN_MAX_ANG = 32;     // evaluated angular functions (type 111)
constexpr int RN_MAX_ETA = 8;      // distinct eta among the evaluated angular functions
constexpr int RN_MAX_LZ = 8;       // distinct (lambda, zeta) among them
constexpr int RN_MAX_HID = 32;     // nodes per hidden layer
#ifndef PTMC_RN_THREADS
#define PTMC_RN_THREADS 384
#endif
constexpr int RN_THREADS = PTMC_RN_THREADS;   // one CTA = one replica: 7 warps on the centre-j terms, 5 on the centre-i terms (N = 55)
constexpr int RN_SUB = 4;          // threads sharing one atom j in the angular update
constexpr int RN_MAX_ATOMS = RN_THREADS >= 384 ? (RN_THREADS / 32 - 2) * 32 / RN_SUB : RN_THREADS / RN_SUB;   // split mode: at least two warps stay with the centre-i terms

// Only symmetry functions of type 11 / 111 are ever evaluated (DeltaMatrix.jl:803,840, SymmFunc.jl:261,281); the rows
// of all others stay exactly 0 for ever and contribute exactly 0 to the first network layer, so the device keeps a
// compact g matrix [atom][n_act] of the evaluated rows (radial first) and the first-layer weights of those rows.
struct RunnerDev {
    int n_sym, n_rad, n_ang;
    int n_act_rad, n_act_ang, n_act;
    int n_eta, n_lz;
    int act_row[RN_MAX_RAD + RN_MAX_ANG];       // row of compact function a in the reference's g_matrix
    double rad_eta[RN_MAX_RAD], rad_norm[RN_MAX_RAD], rad_off[RN_MAX_RAD];
    int ang_eta_idx[RN_MAX_ANG], ang_lz_idx[RN_MAX_ANG];
    double ang_tpz[RN_MAX_ANG], ang_off[RN_MAX_ANG];
    int ang_init[RN_MAX_ANG];                   // 0: row is NOT filled at initialisation (SymmFunc.jl:504 loops (Nrad+1):Nang)
    unsigned long long present;                 // bit (e * RN_MAX_LZ + z): an evaluated function has eta e and (lambda, zeta) z
    signed char qmap[RN_MAX_ETA * RN_MAX_LZ];   // (e, z) -> compact angular function index, -1 if none
    double eta[RN_MAX_ETA];
    double lz_lambda[RN_MAX_LZ], lz_zeta[RN_MAX_LZ];
    int lz_zint[RN_MAX_LZ];                     // zeta as a small integer power (1, 2, 4, ...) or 0 -> pow()
    int lz_nsq[RN_MAX_LZ];                      // zeta = 2^nsq: number of squarings; 15 -> pow()
    double r_cut, boundary2;
    int h1, h2, act1, act2, act3;               // network n_act(88) - h1 - h2 - 1, activation codes 1 linear / 2 tanh
    const double* net;                          // W1a [h1 x n_act] col-major, b1, W2 [h2 x h1] col-major, b2, W3 [h2], b3
    // per-configuration state, indexed [ladder][config]; perm_g (DevParams) maps slot -> config
    double* pos_c;      // [N][3]
    double* en_c;       // en_tot
    double* en_atom_c;  // [N]
    double* g_c;        // [N][n_act]
    // pair tables, one entry per unordered pair (upper triangle, NT = N (N-1) / 2 entries, index rn_tri):
    // [pair_c] = d2 | 1/d | f | ef[0] .. ef[n_eta-1], each NT long; staged into shared memory for a whole cycle
    double* pair_c;     // [3 + n_eta][NT]: dist2_mat, 1/sqrt(dist2_mat), f_matrix, exp(-eta d2) * f
};

// Two species (RuNNerPotential2Atom, src/EnergyEvaluation.jl:1287-1314): atoms 0 .. n1-1 are species 1, the rest species 2; every
// symmetry function is evaluated (DeltaMatrix.jl:256-316 has no type filter).  The g matrix keeps the reference's rows:
// radial function k at rows 2k + {species of the neighbour}, angular function q at 2 n_rad + 3q + {number of species-2 atoms among
// the two neighbours} (SymmFunc.jl:511-545).  One network per species.
struct Runner2Dev {
    int n_sym, n_rad, n_ang, n1;
    int n_eta, n_lz;
    double rad_eta[RN_MAX_RAD];
    double rad_norm[RN_MAX_RAD][4], rad_off[RN_MAX_RAD][4];      // [2 * species of the centre + row]  (RadialType2a, SymmFunc.jl:64-83)
    int ang_eta_idx[RN_MAX_ANG], ang_lz_idx[RN_MAX_ANG];
    double ang_tpz[RN_MAX_ANG][6], ang_off[RN_MAX_ANG][6];       // [3 * species of the centre + row]  (AngularType3a, :133-171)
    unsigned long long present;                 // as RunnerDev
    signed char qmap[RN_MAX_ETA * RN_MAX_LZ];
    double eta[RN_MAX_ETA];
    double lz_lambda[RN_MAX_LZ], lz_zeta[RN_MAX_LZ];
    int lz_zint[RN_MAX_LZ], lz_nsq[RN_MAX_LZ];
    double r_cut;
    int h1, h2, act1, act2, act3;
    int net_len;                                // doubles per network: W1 [h1 x n_sym] col-major, b1, W2 [h2 x h1] col-major, b2, W3 [h2], b3
    const double* net;                          // network of species 1, then of species 2
    // per-configuration state, indexed [ladder][config]
    double* pos_c;      // [N][3]
    double* en_c;
    double* en_atom_c;  // [N]
    double* g_c;        // [N][n_sym]
    double* pair_c;     // [3 + n_eta][NT] as RunnerDev; read from HBM / L1 by the kernel (the two g matrices fill shared memory)
};

// end-of-cycle step (exchange / step-size adaptation / sampling) and slot-ordered gather, shared by the one- and two-species sweeps
void runner_launch_cycle_end(const DevParams& P, const double* en_c, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st);
void runner_launch_gather_ptr(const DevParams& P, const double* pos_c, const double* en_c, i64 n_rep, cudaStream_t st);

// The angular symmetry-function set of every RuNNer script of the reference (scripts/cu55.jl:70-97, cu38.jl, brass55.jl,
// test/potentialfile.jl: the 26-row table "V"), as structure only: which (eta index, (lambda, zeta) index) pairs exist
// and how many squarings each (lambda, zeta) needs, with indices in order of first appearance.  The kernels are
// specialised on this structure (all register indices and loop bounds become compile-time); the VALUES of eta, lambda,
// zeta, the scaling and the network stay runtime data.  Any other set runs through the generic instantiation.
constexpr unsigned long long RN_PRESENT_V =
    (1ull << (0 * 8 + 0)) | (1ull << (0 * 8 + 1)) | (1ull << (1 * 8 + 2)) | (1ull << (1 * 8 + 1)) | (1ull << (2 * 8 + 2)) | (1ull << (2 * 8 + 1)) |
    (1ull << (2 * 8 + 3)) | (1ull << (3 * 8 + 0)) | (1ull << (3 * 8 + 1)) | (1ull << (3 * 8 + 4)) | (1ull << (3 * 8 + 5)) | (1ull << (4 * 8 + 2)) |
    (1ull << (4 * 8 + 0)) | (1ull << (4 * 8 + 3)) | (1ull << (4 * 8 + 4)) | (1ull << (4 * 8 + 5)) | (1ull << (4 * 8 + 6)) | (1ull << (5 * 8 + 0)) |
    (1ull << (5 * 8 + 1)) | (1ull << (5 * 8 + 4)) | (1ull << (5 * 8 + 7)) | (1ull << (5 * 8 + 6)) | (1ull << (6 * 8 + 0)) | (1ull << (6 * 8 + 1)) |
    (1ull << (6 * 8 + 4)) | (1ull << (6 * 8 + 7));
// squarings per (lambda, zeta) slot, 4 bits each, slot 0 in the low nibble: zeta = 1, 2, 1, 2, 4, 16, 16, 4
constexpr unsigned RN_NSQ_V = 0u | (1u << 4) | (0u << 8) | (1u << 12) | (2u << 16) | (4u << 20) | (4u << 24) | (2u << 28);

// Evaluation order of the (lambda, zeta) slots of that set: two chains (lambda = +1: slots 0, 3, 7, 6 with zeta 1, 2, 4, 16;
// lambda = -1: slots 2, 1, 4, 5), each slot obtained from its predecessor by squaring - (1 + lambda cos)^zeta costs 2 FMA +
// 8 MUL for all eight slots instead of 8 + 14.  Position q: slot (4 bits), root flag, squarings from the predecessor (4 bits).
constexpr unsigned RN_ZSEQ_V = 0u | (3u << 4) | (7u << 8) | (6u << 12) | (2u << 16) | (1u << 20) | (4u << 24) | (5u << 28);
constexpr unsigned RN_ZROOT_V = (1u << 0) | (1u << 4);
constexpr unsigned RN_ZSQ_V = 0u | (1u << 4) | (1u << 8) | (2u << 12) | (0u << 16) | (1u << 20) | (1u << 24) | (2u << 28);

// index of the unordered pair (a, b), a != b, in the upper-triangle tables
__host__ __device__ __forceinline__ int rn_tri(int a, int b, int N) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return lo * N - (lo * (lo + 1)) / 2 + (hi - lo - 1);
}

// ------------------------------------------------------------------------------------------ helpers --
__device__ __forceinline__ double rn_cutoff(double r, double rc) {       // Cutoff.jl:79-86
    return r < rc ? 0.5 * (cos(3.14159265358979323846 * (r / rc)) + 1.0) : 0.0;
}
// (1 + lambda c)^zeta: repeated squaring for zeta = 2^m (1, 2, 4, 8, 16, ...), pow() otherwise  (SymmFunc.jl:192)
__device__ __forceinline__ double rn_theta(double c, double lambda, double zeta, int zint) {
    double b = fma(lambda, c, 1.0);
    if (zint > 0) {
        for (int m = zint; m > 1; m >>= 1) b *= b;
        return b;
    }
    return pow(b, zeta);
}
// the same with the number of squarings nsq (zeta = 2^nsq, at most 2^5); nsq == 15 -> pow().  In the specialised
// kernels nsq is a compile-time constant after unrolling and the predicates fold away.
__device__ __forceinline__ double rn_theta_n(double c, double lambda, double zeta, int nsq) {
    double b = fma(lambda, c, 1.0);
    if (nsq == 15) return pow(b, zeta);
#pragma unroll
    for (int m = 0; m < 5; ++m) if (m < nsq) b *= b;
    return b;
}
// exp(x) for x in [-700, 0], branch-free: x = k ln2 + r, |r| <= ln2 / 2, Taylor series to degree 13 (remainder 4e-18),
// exponent patched into the high word
__device__ __forceinline__ double rn_exp_neg(double x) {
    x = fmax(x, -700.0);
    const double kf = rint(x * 1.4426950408889634);
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = 1.60590438368216145994e-10;                  // 1/13!
    p = fma(p, r, 2.08767569878680989792e-09);
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + ((int)kf << 20), __double2loint(p));
}
// tanh(z) = sign(z) (1 - t) / (1 + t), t = exp(-2 |z|): branch-free, absolute error ~1e-16 (the library routine
// diverges between its two argument ranges inside a warp: 9 % of the kernel's instructions before)
__device__ __forceinline__ double rn_tanh(double z) {
    const double t = rn_exp_neg(-2.0 * fabs(z));
    const double r = (1.0 - t) * rcp_pos(1.0 + t);
    return copysign(r, z);
}
__device__ __forceinline__ double rn_act(double z, int code) { return code == 2 ? rn_tanh(z) : z; }

// acc[e][z] += ep_new[e] (1 + lambda_z cos_new)^zeta_z - ep_old[e] (1 + lambda_z cos_old)^zeta_z for every evaluated (e, z).
// SPEC (the reference's standard set): the slots are walked along the two squaring chains of RN_ZSEQ_V.
template <bool SPEC, unsigned long long PRESENT, unsigned NSQ, class RD>
__device__ __forceinline__ void rn_angular_acc(const RD& R, unsigned long long present, double cn, double co,
                                               const double (&epn)[RN_MAX_ETA], const double (&epo)[RN_MAX_ETA], double (&acc)[RN_MAX_ETA][RN_MAX_LZ]) {
    if (SPEC) {
        double bn = 0.0, bo = 0.0;
#pragma unroll
        for (int q = 0; q < RN_MAX_LZ; ++q) {
            const int z = (int)((RN_ZSEQ_V >> (4 * q)) & 15u), sq = (int)((RN_ZSQ_V >> (4 * q)) & 15u);
            if ((RN_ZROOT_V >> q) & 1u) { bn = fma(R.lz_lambda[z], cn, 1.0); bo = fma(R.lz_lambda[z], co, 1.0); }
#pragma unroll
            for (int m = 0; m < 4; ++m) if (m < sq) { bn *= bn; bo *= bo; }
#pragma unroll
            for (int e = 0; e < RN_MAX_ETA; ++e)
                if (PRESENT >> (e * RN_MAX_LZ + z) & 1ull) acc[e][z] += epn[e] * bn - epo[e] * bo;
        }
    } else {
#pragma unroll
        for (int z = 0; z < RN_MAX_LZ; ++z) {
            if (z < R.n_lz) {
                const double tn = rn_theta_n(cn, R.lz_lambda[z], R.lz_zeta[z], R.lz_nsq[z]);
                const double to = rn_theta_n(co, R.lz_lambda[z], R.lz_zeta[z], R.lz_nsq[z]);
#pragma unroll
                for (int e = 0; e < RN_MAX_ETA; ++e)
                    if (present >> (e * RN_MAX_LZ + z) & 1ull) acc[e][z] += epn[e] * tn - epo[e] * to;
            }
        }
    }
}

// D (8 x 8, FP64) += A (8 x 4, row) * B (4 x 8, col) on the tensor cores (DMMA; tcgen05 has no f64 kind).  Lane l holds
// A[l / 4][l % 4], B[l % 4][l / 4] and D[l / 4][2 (l % 4) + {0, 1}].
__device__ __forceinline__ void rn_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
}  // namespace ptmc
