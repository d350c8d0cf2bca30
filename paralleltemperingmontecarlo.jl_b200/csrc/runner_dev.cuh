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

}  // namespace ptmc
