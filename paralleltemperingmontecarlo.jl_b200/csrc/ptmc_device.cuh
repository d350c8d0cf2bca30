// ptmc_device.cuh — device-side building blocks of the B200 (sm_100a) sweep kernels.
//
// Data layout.  One CTA = one parallel-tempering ladder; one group of G lanes
// (G = 4/8/16/32, a power of two inside one warp) = one replica slot.  Configuration-side state
// (positions, box, en_tot, r_cut, EAM components) lives in shared memory, indexed through a slot->config
// permutation so that a replica exchange (src/Exchange.jl:115-124) is a swap of two ints.
// Slot-side state (beta, max_displ, counters, ham, RNG) lives in registers, replicated in every lane of
// the group, so every lane takes identical decisions from identical bits.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ptmc_b200.h"

namespace ptmc {

typedef unsigned long long u64;
typedef long long i64;

#define PTMC_TIE_K 32           // tie-table row capacity (fcc / hcp shells on a cut-off sphere: <= 24)
#define PTMC_TIE_REL 1e-7       // relative distance to the cut-off below which a pair keeps its frozen dist2_mat entry
#define PTMC_EAM_DEG 6          // degree of the binomial series of (1 + t)^-p, |t| <= 2^-9
#define PTMC_EAM_K 256          // mantissa intervals
#ifndef PTMC_EAM_KSTRIDE
#define PTMC_EAM_KSTRIDE 6       // doubles per mantissa entry {1/m_k, (1/m_k)^(n/2), (1/m_k)^(m/2), pad}: 48 B spreads LDS.128 over all banks
#endif
// Horner (0) or Estrin (1) form of the two degree-6 series of eam_invrexp.  Estrin is three levels shallower and won in round 1; every one
// of its inner FMAs has two coefficient operands, though, so half of the twelve coefficients must sit in registers (or be re-read: ncu
// counted 6 LDC per pair iteration at the 128-register cap), whereas Horner takes them all as constant-bank operands.  Since the
// divergent branch around the square root went, Horner is ahead: 54.9 -> 53.3 ms per C3 step (profiles/r02_experiments.md).
#ifndef PTMC_EAM_ESTRIN
#define PTMC_EAM_ESTRIN 0
#endif
#define PTMC_EAM_KTAB_DOUBLES (PTMC_EAM_K * PTMC_EAM_KSTRIDE)
#define PTMC_EAM_ETAB_DOUBLES (2048 * 2)  // every biased binary exponent of r2

// Everything the kernels need, passed by value as the kernel parameter (constant bank).
struct DevParams {
    // sizes
    int n_traj, n_atoms, n_pad;          // n_pad = n_atoms rounded up to a multiple of the lanes per replica
    int n_atom_moves, n_volume_moves, n_steps;
    int ensemble, bc_kind, pot_kind, n_coeff;
    int n_adjust, mc_sample, n_bin;
    int replay;                          // 1: consume pre-drawn draws
    int phase;                           // ptmc_run_phase: 0 whole mc_cycle!, 1 mc_step! only, 2 step-size adaptation + sampling only
    double min_acc, max_acc, pressure;
    double coeff[PTMC_MAX_COEFF];
    double eljb_ca[PTMC_MAX_COEFF], eljb_cb[PTMC_MAX_COEFF];   // ELJPotentialB: c_i a_i, c_i b_i (c_i in coeff[])
    // LookupTablePotential (src/EnergyEvaluation.jl:214-223): table [l_angle][l_dist] in HBM (L2 resident, ~1 MB)
    const double* lut_table;
    int lut_l_dist;
    double lut_start, lut_d, lut_dang, lut_c6;
    float lut_kf;             // (float)(180 / pi / d_angle)
    const double* lut_thr;    // [PTMC_LUT_THR] ascending tangents at which the angle index steps up (NaN padded), lut_thresholds_kernel
    double eam_hn, eam_hm;               // n/2, m/2  (r^-n = (r2)^-(n/2))
    double eam_ean, eam_ecam;
    // table-driven (r2)^-p, p = n/2 and m/2 (eam_invrexp): binomial coefficients binom(-p, 1..6) and the tables
    double eam_cn[PTMC_EAM_DEG], eam_cm[PTMC_EAM_DEG];
    const double* eam_ktab;              // [256][PTMC_EAM_KSTRIDE]
    const double* eam_etab;              // [2048][2]
    // config-side state at rest, slot ordered: [L][T]...
    double* pos;        // [L][T][N][3]
    double* box;        // [L][T][2]
    double* en_tot;     // [L][T]
    double* comp;       // [L][T][N][2]  EAM (phi, rho)
    // periodic boundaries: the reference's dist2_mat, kept in HBM per CONFIGURATION (not per slot) so that an
    // exchange never moves it; perm_g is the slot -> configuration map, persisted between launches
    double* d2mat;      // [L][T][N][N]
    double* tan_rho;    // [L][T]  tangent-dependent potentials with separated volume moves: factor of the stale tangents (sweep_kernel.cuh)
    int* fresh_g;       // [L][T][moved_words]  atoms moved since the last xy / z volume trial
    int* perm_g;        // [L][T]
    // per configuration: bit mask of the atoms moved (accepted) since the upload + their count.  Frozen dist2_mat
    // entries can only differ in a DECISION from recomputed distances for pairs of never-moved lattice atoms (exact
    // cut-off ties); once every atom of a configuration has moved, its matrix is no longer read or maintained.
    int* moved_g;       // [L][T][moved_words + 1]  (last int: count)
    // Periodic NPT with uniform volume moves (use_ties): instead of the N x N matrix, the pairs that sit on the cut-off at upload
    // (|d2 / r_cut - 1| < PTMC_TIE_REL, exact distance2 arithmetic) are kept per configuration and per atom with their frozen
    // values; every other pair is recomputed from the positions (its decision cannot differ: it is >= 1e-7 away from the cut-off,
    // rounding moves it by 1e-16 per volume move).  A configuration without such pairs starts thawed.
    int cfg_mod;        // residue (mod 16 doubles) of the stride between configuration buffers in shared memory, -1: the lane count (smem_layout)
    int use_ties;
    int* tie_cnt;             // [L][T][N]
    unsigned char* tie_idx;   // [L][T][N][PTMC_TIE_K]
    double* tie_val;          // [L][T][N][PTMC_TIE_K]  dist2_mat entries, scaled like the reference's at accepted volume moves
    int moved_words;    // ceil(n_atoms / 32)
    // slot-side state
    double* max_displ;  // [L][T][4]
    double* max_boxlength; // [L][T]
    double* max_boxheight; // [L][T]  (separated volume moves)
    int separated;         // NPT.separated_volume
    i64* counters;      // [L][T][PTMC_NCOUNT]
    double* ham;        // [L][T][2]
    const double* betas;// [T]
    u64* slot_rng;      // [L][T][4]
    u64* ladder_rng;    // [L][4]
    // replay
    const u64* slot_draws; i64 slot_len; i64* slot_cursor;       // [L][T][slot_len], [L][T]
    const u64* ladder_draws; i64 ladder_len; i64* ladder_cursor; // [L][ladder_len], [L]
    // histograms
    const double* en_min; const double* en_max; const double* v_min; const double* v_max; // [L]
    u64* hist;          // NVT [L][T][nb+2]; NPT [L][T][nb+2][nb+2]
    // trace (parity mode)
    int32_t* tr_index; double* tr_de; uint8_t* tr_acc; int32_t* tr_exc;
    i64 trace_cap, trace_base;
    int trace_on;       // trace_cap > 0 as a 32-bit flag: one uniform predicate in the move loop of every instantiation
    int* error;         // device error word (most negative ptmc_status wins)
};

// --------------------------------------------------------------------------------- exact arithmetic --
// Decisions and positions must follow the reference bit for bit, so everything that feeds a stored
// position, a comparison or a histogram index is written with non-contracting intrinsics (Julia never
// fuses a*b+c).  Energies may use FMA (parity bar: 1e-10 relative).
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// ---------------------------------------------------------------------------------------------- RNG --
// Julia's default generator: Xoshiro256++; rand() = (u >> 11) * 2^-53; rand(1:n) = Lemire.
struct Xo { u64 s0, s1, s2, s3; };
__device__ __forceinline__ u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ u64 xo_next(Xo& r) {
    u64 res = rotl64(r.s0 + r.s3, 23) + r.s0;
    u64 t = r.s1 << 17;
    r.s2 ^= r.s0; r.s3 ^= r.s1; r.s1 ^= r.s2; r.s0 ^= r.s3; r.s2 ^= t; r.s3 = rotl64(r.s3, 45);
    return res;
}

template <bool DBG>
struct Stream {
    Xo r;
    const u64* draws; i64 pos, len; int exhausted;   // replay (DBG kernels only)
    __device__ __forceinline__ u64 next() {
        if (DBG) {
            if (draws) {
                if (pos >= len) { exhausted = 1; return 0; }
                return draws[pos++];
            }
        }
        return xo_next(r);
    }
    __device__ __forceinline__ double rand() { return (double)(next() >> 11) * 0x1.0p-53; }
    // rand(1:n)
    __device__ __forceinline__ int range(u64 n) {
        u64 x = next();
        u64 hi = __umul64hi(x, n), lo = x * n;
        if (lo < n) {
            u64 t = (0 - n) % n;
            while (lo < t) {
                if (DBG) { if (exhausted) break; }
                x = next(); hi = __umul64hi(x, n); lo = x * n;
            }
        }
        return (int)hi + 1;
    }
};

// PTMC_RNG_BATCH8 (include/ptmc_b200.h): the slot owns 8 substreams; lane g of the group steps substream
// (g & 7) (G = 4: substreams g and g + 4) once per batch and the lanes read position k with a shuffle, so a
// draw costs two SHFL instead of a redundant ~25-instruction generator step in every lane.
template <int G>
struct BatchRng {
    // lane g of the group steps the substreams g, g + G, ... (NS = 8 / G of them for G < 8, one for G >= 8): substream K lives in
    // slot K / G of lane K % G
    static constexpr int NS = G >= 8 ? 1 : 8 / G;
    Xo st_[NS];
    u64 buf[NS];
    __device__ __forceinline__ void load(const u64* st, int g) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            const u64* p = st + 4 * ((G >= 8 ? (g & 7) : g) + m * G);
            st_[m].s0 = p[0]; st_[m].s1 = p[1]; st_[m].s2 = p[2]; st_[m].s3 = p[3];
            buf[m] = 0;
        }
    }
    __device__ __forceinline__ void store(u64* st, int g) const {
        if (G >= 8 && g >= 8) return;
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            u64* p = st + 4 * (g + m * G);
            p[0] = st_[m].s0; p[1] = st_[m].s1; p[2] = st_[m].s2; p[3] = st_[m].s3;
        }
    }
    __device__ __forceinline__ void begin() {
#pragma unroll
        for (int m = 0; m < NS; ++m) buf[m] = xo_next(st_[m]);
    }
    template <int K>
    __device__ __forceinline__ u64 raw(unsigned mask) const {
        if (G >= 8) return __shfl_sync(mask, buf[0], K, G);
        return __shfl_sync(mask, buf[(K / G) % NS], K % G, G);
    }
    template <int K>
    __device__ __forceinline__ double rand(unsigned mask) const { return (double)(raw<K>(mask) >> 11) * 0x1.0p-53; }
    // (rand() - 0.5) * 2^53 = v - 2^52 for the 53-bit integer v = raw >> 11, exactly, without the 64-bit integer -> double conversion
    // (I2F.F64.U64 runs on the XU pipe, and its latency heads the dependency chain of every move): the low 52 bits m go under the
    // exponent of 2^52 (w = 2^52 + m), then w - (bit 52 ? 2^52 : 2^53).  With max_displ * 2^-53 (exact) the displacement
    // (rand() - 0.5) * max_displ of atom_displacement (src/MCMoves.jl:35-37) comes out bit for bit.
    template <int K>
    __device__ __forceinline__ double centered(unsigned mask) const {
        const u64 v = raw<K>(mask) >> 11;
        const unsigned hi = (unsigned)(v >> 32);
        const double w = __hiloint2double((int)(0x43300000u | (hi & 0xfffffu)), (int)(unsigned)v);
        return w - ((hi >> 20) ? 4503599627370496.0 : 9007199254740992.0);
    }
};

// ----------------------------------------------------------------------------------- group reduction --
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(mask, v, m);
    return v;   // xor butterfly: every lane ends with the same bits (a+b == b+a)
}

// --------------------------------------------------------------------------------------- reciprocals --
// 1/x for normal positive x: MUFU.RCP64H seed (~2^-20) + two Newton steps (error ~1 ulp), no special cases.
__device__ __forceinline__ double rcp_pos(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}
// 1/x for normal positive x with ONE third-order (Halley) step: y1 = y0 (1 + e + e^2), e = 1 - x y0.
// The MUFU.RCP64H seed carries >= 20 bits, so e^3 < 2^-60: 3 DFMA instead of 4 (checked by ptmc_debug_pair_energy).
__device__ __forceinline__ double rcp_pos3(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}
// 1/sqrt(x) for normal positive x: MUFU.RSQ64H seed + two Newton steps.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

// -------------------------------------------------------------------------------------- pair energy --
// dimer_energy(::ELJPotentialEven) src/EnergyEvaluation.jl:934-942: sum_i c_i r^-(4+2i) = x^3 * Horner(x),
// x = 1/r2.  Same polynomial, one reciprocal instead of 1+K divisions (differs at the 1e-16 level).
template <int NC>
__device__ __forceinline__ double elj_even(const DevParams& P, double r2) {
    double x = rcp_pos(r2);
    double x3 = x * x * x;
    double p;
    if (NC == 6) {
        p = P.coeff[5];
        p = fma(p, x, P.coeff[4]); p = fma(p, x, P.coeff[3]); p = fma(p, x, P.coeff[2]);
        p = fma(p, x, P.coeff[1]); p = fma(p, x, P.coeff[0]);
    } else {
        p = P.coeff[P.n_coeff - 1];
        for (int k = P.n_coeff - 2; k >= 0; --k) p = fma(p, x, P.coeff[k]);
    }
    return p * x3;
}
// dimer_energy(::ELJPotential) src/EnergyEvaluation.jl:923-932: sum_i c_i r^-(5+i) = y^6 * Horner(y), y = 1/r.
__device__ __forceinline__ double elj_odd(const DevParams& P, double r2) {
    double y = rsqrt_pos(r2);
    double y2 = y * y;
    double y6 = y2 * y2 * y2;
    double p = P.coeff[P.n_coeff - 1];
    for (int k = P.n_coeff - 2; k >= 0; --k) p = fma(p, y, P.coeff[k]);
    return p * y6;
}
// Horner part of ELJPotentialEven for x = 1/r2: returns P(x) with e = x^3 P(x)
template <int NC>
__device__ __forceinline__ double elj_even_poly(const DevParams& P, double x) {
    double p;
    if (NC == 6) {
        p = P.coeff[5];
        p = fma(p, x, P.coeff[4]); p = fma(p, x, P.coeff[3]); p = fma(p, x, P.coeff[2]);
        p = fma(p, x, P.coeff[1]); p = fma(p, x, P.coeff[0]);
    } else {
        p = P.coeff[P.n_coeff - 1];
        for (int k = P.n_coeff - 2; k >= 0; --k) p = fma(p, x, P.coeff[k]);
    }
    return p;
}
// Delta-energy kernel of one pair for ELJPotentialEven (src/EnergyEvaluation.jl:804-812 with :934-942):
// acc_new += e(d2n), acc_old += e(d2o), sharing ONE reciprocal: y = 1/(d2n d2o), 1/d2n = y d2o, 1/d2o = y d2n.
// 34 FP64 instructions per pair including the two squared distances.
// Lanes without a pair pass d2n = d2o = PAIR_OFF: x^3 underflows to exactly 0 and nothing is accumulated.
#define PTMC_PAIR_OFF 1e150
// Coordinate of a "parked" atom: the moved atom itself and the padding atoms of the free-space ELJEven kernels sit at
// (FAR, FAR, FAR) while a row is evaluated, so their pair terms vanish exactly (d2 = 3e150 -> x^3 underflows to 0)
// without any per-pair select.
#define PTMC_FAR 1e75
template <int NC>
__device__ __forceinline__ void elj_even_pair_acc(const DevParams& P, double d2n, double d2o, double& acc_new, double& acc_old) {
    const double y = rcp_pos3(d2n * d2o);
    const double xn = y * d2o, xo = y * d2n;
    const double pn = elj_even_poly<NC>(P, xn), po = elj_even_poly<NC>(P, xo);
    acc_new = fma(pn, xn * xn * xn, acc_new);
    acc_old = fma(po, xo * xo * xo, acc_old);
}

// dimer_energy(::ELJPotentialB) src/EnergyEvaluation.jl:943-958: sum_i c_i r^-k_i (1 + a_i t2 + b_i t4), k = 6, 8, 9, 10, ...;
// t2 = cos 2theta = 2 / (tan^2 + 1) - 1 = 2 dz^2 / r^2 - 1 (the angle comes from the positions like the distance; no tangent
// matrix is stored - where the reference's becomes state of its own, periodic cells with separated volume moves, it is carried
// as one factor per configuration: pair_energy_stale), t4 = 2 t2^2 - 1; below r2 = 5.30 the reference returns 0.1.
template <int NC = 0>
__device__ __forceinline__ double eljb_energy_t2(const DevParams& P, double r2, double y, double t2);
template <int NC = 0>
__device__ __forceinline__ double eljb_energy(const DevParams& P, double r2, double dz2) {
    const double y = rsqrt_pos(r2);
    return eljb_energy_t2<NC>(P, r2, y, fma(2.0 * dz2, y * y, -1.0));
}
template <int NC>
__device__ __forceinline__ double eljb_energy_t2(const DevParams& P, double r2, double y, double t2) {
    const double x = y * y;
    const double t4 = fma(2.0 * t2, t2, -1.0);
    const int n = NC ? NC : P.n_coeff;
    double p = 0.0;
#pragma unroll
    for (int k = n - 1; k >= 1; --k) p = fma(p, y, fma(P.eljb_cb[k], t4, fma(P.eljb_ca[k], t2, P.coeff[k])));
    const double w1 = fma(P.eljb_cb[0], t4, fma(P.eljb_ca[0], t2, P.coeff[0]));
    const double s = (x * x * x) * fma(x, p, w1);
    return r2 >= 5.30 ? s : 0.1;
}

// dimer_energy(::LookupTablePotential) src/EnergyEvaluation.jl:274-295 with get_tan (src/Configurations.jl:177-180) of the
// separation (dx, dy, dz): the table indices are decisions, so the index arithmetic keeps the reference's operation order.
// get_tan src/Configurations.jl:177-180 of a separation
__device__ __forceinline__ double tan_of(double dx, double dy, double dz) { return fabs(xdiv(sqrt(xadd(xmul(dx, dx), xmul(dy, dy))), dz)); }
__device__ __forceinline__ double lut_energy_tan(const DevParams& P, double r2, double t);
__device__ __forceinline__ double lut_energy(const DevParams& P, double r2, double dx, double dy, double dz) {
    return lut_energy_tan(P, r2, tan_of(dx, dy, dz));
}
// angle index as the reference writes it (src/EnergyEvaluation.jl:273-283): round(atan(|tan|) / d_angle * 180 / pi + 1) inside
// (0.00872687153, 114.592845357], 1 below, 91 above
__device__ __forceinline__ int lut_angle_index_direct(const DevParams& P, double t) {
    int ai = 1;
    if (t > 0.00872687153 && t <= 114.592845357) ai = (int)rint(xadd(xdiv(xmul(xdiv(atan(t), P.lut_dang), 180.0), 3.14159265358979323846), 1.0));
    else if (t > 114.592845357) ai = 91;
    return ai;
}
// The index is a non-decreasing step function of the tangent, so the sweep kernels count the steps at or below t instead of
// evaluating an FP64 atan and two divisions per pair: thr[k] = smallest double with index >= k + 2, found by bisection over the
// bit patterns with lut_angle_index_direct itself (lut_thresholds_kernel) -> the same index for every t.  A single-precision
// atan gives the count to +-1; four independent loads around it fix and verify it (a window that does not bracket t - only
// possible for d_angle != 1, where the index jumps to 91 - falls back to the direct formula).
#define PTMC_LUT_THR 127
__device__ __forceinline__ int lut_angle_index(const DevParams& P, double t) {
    const float xf = atanf((float)t) * P.lut_kf;                              // ~ atan(t) / d_angle * 180 / pi
    int lo = (int)rintf(xf) - 1;
    lo = lo < 1 ? 1 : (lo > 90 ? 90 : lo);                                    // window thr[lo-1 .. lo+2]
    const double* w = P.lut_thr + lo - 1;
    const double t0 = __ldg(w), t1 = __ldg(w + 1), t2 = __ldg(w + 2), t3 = __ldg(w + 3);
    int n = lo + (t >= t1 ? 1 : 0) + (t >= t2 ? 1 : 0);
    if (!(t >= t0)) n = (lo == 1) ? 0 : -1;                                   // below the window: fine only if the window starts at 0
    if (t >= t3) n = -1;                                                      // above the window
    if (n < 0) return lut_angle_index_direct(P, t);
    return 1 + n;
}
__device__ __forceinline__ double lut_energy_tan(const DevParams& P, double r2, double t) {
    const int ai = lut_angle_index(P, t);
    const double lo = xadd(P.lut_start, xmul(0.5, P.lut_d)), hi = xadd(P.lut_start, xmul((double)P.lut_l_dist, P.lut_d));
    const double* row = P.lut_table + (size_t)(ai - 1) * P.lut_l_dist;
    if (r2 <= xmul(lo, lo)) return __ldg(row);
    if (r2 <= xmul(hi, hi)) {
        // (the same step-counting for this index was slower: its loads sit in front of the dependent table gather, -17 %)
        int di = (int)rint(xdiv(xsub(sqrt(r2), P.lut_start), P.lut_d));
        di = di < 1 ? 1 : di;              // (the reference indexes table[.., 0] -> BoundsError if the quotient rounds to 0)
        return __ldg(row + di - 1);
    }
    return xdiv(P.lut_c6, xmul(xmul(r2, r2), r2));
}

// NC = 6: the six-coefficient form of every BASELINE script, fully unrolled; NC = 0: runtime n_coeff
template <int POT, int NC = 0>
__device__ __forceinline__ double pair_energy(const DevParams& P, double r2) {
    return POT == PTMC_POT_ELJ ? elj_odd(P, r2) : elj_even<NC>(P, r2);
}
// invrexp src/EnergyEvaluation.jl:1088-1095: (r^-n, r^-m) = ((r2)^-(n/2), (r2)^-(m/2)); r2 > 0.
// Table-driven and branch-free: r2 = 2^e m, m in [1, 2); k = top 8 mantissa bits, t = m / m_k - 1 (one FMA with the
// tabulated reciprocal, |t| <= 2^-9); (r2)^-p = 2^(-p e) (1/m_k)^p (1 + t)^-p with the binomial series to degree 6
// (truncation < 2e-17 for p <= 12): 17 FP64 instructions for both powers, <= 2 ulp (exp(-p log r2): ~80 instructions,
// ~16 ulp).  ktab: [256][PTMC_EAM_KSTRIDE] = {1/m_k, (1/m_k)^(n/2), (1/m_k)^(m/2), pad} (shared memory in the sweep kernel);
// etab: [2048][2] = {2^(-n/2 e), 2^(-m/2 e)}, e = -1023..1024 (global memory: a handful of entries is ever touched).
// r2 = 0 / denormal / inf / NaN are not special-cased (the reference returns (0, 0) for r2 == 0, which needs two atoms
// on the same point).
__device__ __forceinline__ void eam_invrexp(const DevParams& P, const double* ktab, const double* etab, double r2, double& phi, double& rho) {
    const int hi = __double2hiint(r2), lo = __double2loint(r2);
    const int eb = (hi >> 20) & 0x7ff;                                         // biased exponent = e + 1023
    const int k = (hi >> 12) & (PTMC_EAM_K - 1);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 a = *reinterpret_cast<const double2*>(ktab + PTMC_EAM_KSTRIDE * k);   // 1/m_k, (1/m_k)^(n/2)
    const double tm = ktab[PTMC_EAM_KSTRIDE * k + 2];                                    // (1/m_k)^(m/2)
    const double2 ee = __ldg(reinterpret_cast<const double2*>(etab) + eb);
    const double t = fma(m, a.x, -1.0);
#if PTMC_EAM_ESTRIN
    // Estrin: 1 + c1 t + t2 ((c2 + c3 t) + t2 ((c4 + c5 t) + t2 c6)): depth 4 instead of 7
    const double t2 = t * t;
    double pn = fma(P.eam_cn[5], t2, fma(P.eam_cn[4], t, P.eam_cn[3])), pm = fma(P.eam_cm[5], t2, fma(P.eam_cm[4], t, P.eam_cm[3]));
    pn = fma(pn, t2, fma(P.eam_cn[2], t, P.eam_cn[1])); pm = fma(pm, t2, fma(P.eam_cm[2], t, P.eam_cm[1]));
    pn = fma(pn, t2, fma(P.eam_cn[0], t, 1.0)); pm = fma(pm, t2, fma(P.eam_cm[0], t, 1.0));
#else
    double pn = P.eam_cn[5], pm = P.eam_cm[5];
#pragma unroll
    for (int j = 4; j >= 0; --j) { pn = fma(pn, t, P.eam_cn[j]); pm = fma(pm, t, P.eam_cm[j]); }
    pn = fma(pn, t, 1.0); pm = fma(pm, t, 1.0);
#endif
    phi = pn * (a.y * ee.x);
    rho = pm * (tm * ee.y);
}
// sqrt(x) for normal positive x, branch-free: MUFU.RSQ64H seed, coupled Newton iteration on (sqrt, 1/(2 sqrt)); <= 1 ulp
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
#ifdef PTMC_SQRT_TWO_STEPS
    r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
#endif
    // seed 2^-20 -> 2^-39 after the coupled step -> the residual step below squares it again: one coupled step is enough (7 instead
    // of 10 FP64 instructions, depth 5 instead of 7; test_eam_sqrt_accuracy)
    return fma(fma(-g, g, x), h, g);
}

// lrc(::ELJPotentialEven) src/EnergyEvaluation.jl:968-982 (r_cut is the squared cut-off)
__device__ __forceinline__ double lrc_even(const DevParams& P, double r_cut) {
    if (r_cut <= 50.0) return 0.0;
    double rs = sqrt(r_cut);
    double rc3 = r_cut * rs;
    double e = 0.0;
    for (int i = 1; i <= P.n_coeff; ++i) { e += P.coeff[i - 1] / rc3 / (double)(2 * i + 1); rc3 *= r_cut; }
    double nn = (double)((i64)P.n_atoms * P.n_atoms);
    return e * 3.14159265358979323846 * nn / 4.0 / (rs * rs * rs);
}

// lrc(::ELJPotentialB) src/EnergyEvaluation.jl:983-998 (fixed coefficients), lrc(::LookupTablePotential) :999-1009 (1.0 below r_cut = 10)
__device__ __forceinline__ double lrc_eljb(const DevParams& P, double r_cut) {
    if (r_cut <= 16.0) return 0.0;
    const double coeff[4] = {-0.1279111890228638, -1.328138539967966, 12.260941135261255, 41.12212408251662};
    double rs = sqrt(r_cut);
    double rc3 = r_cut * rs;
    double e = 0.0;
    for (int i = 1; i <= 4; ++i) { e += coeff[i - 1] / rc3 / (double)(2 * i + 1); rc3 *= r_cut; }
    double nn = (double)((i64)P.n_atoms * P.n_atoms);
    return e * 3.14159265358979323846 * nn / 4.0 / (rs * rs * rs);
}
__device__ __forceinline__ double lrc_lut(const DevParams& P, double r_cut) {
    if (r_cut <= 10.0) return 1.0;
    double rs = sqrt(r_cut);
    double e = P.lut_c6 / (r_cut * rs) / 3.0;
    double nn = (double)((i64)P.n_atoms * P.n_atoms);
    return e * 3.14159265358979323846 * nn / 4.0 / (rs * rs * rs);
}
template <int POT>
__device__ __forceinline__ double lrc_pot(const DevParams& P, double r_cut) {
    return POT == PTMC_POT_ELJ_B ? lrc_eljb(P, r_cut) : (POT == PTMC_POT_LUT ? lrc_lut(P, r_cut) : lrc_even(P, r_cut));
}
// pair energy of the tangent-dependent potentials from r2 and the (minimum-image) separation; s2 scales the separation's
// squared z component to the scale of r2 (uniform volume trial: r2 = d2 * scale^2, the direction is unchanged)
template <int POT, int NC>
__device__ __forceinline__ double pair_energy_dir(const DevParams& P, double r2, double dx, double dy, double dz, double s2 = 1.0) {
    if (POT == PTMC_POT_ELJ_B) return eljb_energy<NC>(P, r2, dz * dz * s2);
    return lut_energy(P, r2, dx, dy, dz);
}
// ... and with the tangent of the state instead of the pair's own: tan_mat[i, j] = rho * tan(i, j) for a stale pair (separated volume
// moves, sweep_kernel.cuh: sep_trial_energy).  ELJPotentialB needs no tangent: t2 = 2 / (tan^2 + 1) - 1 = 2 dz^2 / (rho^2 (dx^2 + dy^2) + dz^2) - 1.
template <int POT, int NC>
__device__ __forceinline__ double pair_energy_stale(const DevParams& P, double r2, double dx, double dy, double dz, double rho, bool stale) {
    if (POT == PTMC_POT_ELJ_B) {
        const double q = dz * dz, s = fma(dx, dx, dy * dy);
        const double den = fma(stale ? rho * rho : 1.0, s, q);
        return eljb_energy_t2<NC>(P, r2, rsqrt_pos(r2), fma(2.0 * q, rcp_pos3(den), -1.0));
    }
    const double t = tan_of(dx, dy, dz);
    return lut_energy_tan(P, r2, stale ? xmul(t, rho) : t);
}

// ----------------------------------------------------------------------------------------- geometry --
// Periodic box of one configuration with the constants the minimum-image formulas need.
struct Box {
    double L, h;        // box_length, box_height (cubic: h = L)
    double iL, ih;      // reciprocals
    double hy, ihy;     // rhombic: sqrt(3)/2 * L and its reciprocal
    double hL;          // L / 2
    double is3;         // 1/sqrt(3)
};
__device__ __forceinline__ Box make_box(int bc_kind, double L, double h) {
    Box b;
    b.L = L; b.h = (bc_kind == PTMC_BC_CUBIC) ? L : h;
    b.iL = 1.0 / b.L; b.ih = 1.0 / b.h;
    const double s3 = sqrt(3.0);
    b.hy = s3 / 2 * L; b.ihy = 1.0 / b.hy; b.is3 = 1.0 / s3; b.hL = 0.5 * b.L;
    return b;
}
// distance2 with minimum image, src/Configurations.jl:104-109 (cubic), :110-120 (rhombic), :121-126 (rect).
// Divisions are replaced by multiplications with reciprocals: the rounded quotient feeding rint() differs
// only when a pair sits exactly on the half-box plane, where both images are equidistant.
// round(a * b) for the minimum-image quotients of dist2_pbc (|a b| of order 1): fma(a, b, 1.5 * 2^52) - 1.5 * 2^52 on the FP64 pipe
// (ties-to-even like rint for |a b| < 2^51; one rounding instead of two) instead of DMUL + FRND on the XU pipe, whose latency sits
// in the middle of the pair loop's dependency chain: C4 +4.5 % (profiles/r02_experiments.md).  Only where the image of a pair that
// sits exactly on a half-cell plane cannot matter (see dist2_pbc); dist2_pbc_vec keeps rint.
__device__ __forceinline__ double rint_mul(double a, double b) {
#ifdef PTMC_RINT_FRND
    return rint(a * b);
#else
    return __dadd_rn(__fma_rn(a, b, 6755399441055744.0), -6755399441055744.0);
#endif
}
template <bool RHOMBIC>
__device__ __forceinline__ double dist2_pbc(const Box& b, double ax, double ay, double az, double bx, double by, double bz) {
    // The reference's rhombic image (:110-120) is b + nu a1 + nv a2 with a1 = (L, 0), a2 = (L/2, sqrt(3)/2 L) and
    // nv = round(ey / (sqrt(3)/2 L)), nu = round((ex - ey / sqrt(3)) / L) of the ORIGINAL separation e = a - b, i.e. an independent
    // round per lattice coordinate: 17 FP64 instructions per distance (the literal form: 24).
    // Callers: pairs with at least one atom at a random (moved / trial) position, and lattice pairs under the NPT cut-off - a pair on a
    // half-cell plane is at least the inscribed radius = sqrt(r_cut) apart in either image, so it is outside the cut-off or an exact
    // cut-off tie, whose recomputed term the tie table replaces by the frozen one (computed through this same function).
    const double ex = ax - bx, ey = ay - by, ez = az - bz;
    double dx, dy;
    if (RHOMBIC) {
        const double nv = rint_mul(ey, b.ihy);
        const double nu = rint_mul(fma(-b.is3, ey, ex), b.iL);
        dx = fma(-b.hL, nv, fma(-b.L, nu, ex));
        dy = fma(-b.hy, nv, ey);
    } else {
        dx = fma(-b.L, rint_mul(ex, b.iL), ex);
        dy = fma(-b.L, rint_mul(ey, b.iL), ey);
    }
    const double dz = fma(-b.h, rint_mul(ez, b.ih), ez);
    return fma(dz, dz, fma(dy, dy, dx * dx));
}
// The same with the minimum-image separation itself (the tangent-dependent potentials need the direction of the pair)
template <bool RHOMBIC>
__device__ __forceinline__ double dist2_pbc_vec(const Box& b, double ax, double ay, double az, double bx, double by, double bz,
                                                double& dx, double& dy, double& dz) {
    const double ex = ax - bx, ey = ay - by, ez = az - bz;
    if (RHOMBIC) {
        const double nv = rint(ey * b.ihy);
        const double nu = rint(fma(-b.is3, ey, ex) * b.iL);
        dx = fma(-b.hL, nv, fma(-b.L, nu, ex));
        dy = fma(-b.hy, nv, ey);
    } else {
        dx = fma(-b.L, rint(ex * b.iL), ex);
        dy = fma(-b.L, rint(ey * b.iL), ey);
    }
    dz = fma(-b.h, rint(ez * b.ih), ez);
    return fma(dz, dz, fma(dy, dy, dx * dx));
}
// distance2 exactly as the reference writes it (src/Configurations.jl:101-126: true divisions, no FMA, image of
// b built first).  Used where bits matter: the initial dist2_mat of a perfect lattice holds pairs that sit
// EXACTLY on the cut-off (r_cut^2 = L^2/4, 3L^2/16, h^2/4), and d2 <= r_cut must fall the way the reference's does.
__device__ __forceinline__ double dist2_exact(int bc_kind, double ax, double ay, double az, double bx, double by, double bz, double L, double h) {
    double Bx, By, Bz;
    if (bc_kind == PTMC_BC_RHOMBIC) {
        const double s3 = sqrt(3.0);
        const double hy = xmul(xdiv(s3, 2.0), L);
        const double is3 = xdiv(1.0, s3);
        By = xadd(by, xmul(hy, rint(xdiv(xsub(ay, by), hy))));
        Bx = xadd(xadd(xsub(bx, xdiv(by, s3)), xmul(L, rint(xdiv(xsub(xsub(ax, bx), xmul(is3, xsub(ay, by))), L)))), xmul(is3, By));
        Bz = xadd(bz, xmul(h, rint(xdiv(xsub(az, bz), h))));
    } else if (bc_kind == PTMC_BC_SPHERICAL) {
        Bx = bx; By = by; Bz = bz;
    } else {
        const double hz = bc_kind == PTMC_BC_CUBIC ? L : h;
        Bx = xadd(bx, xmul(L, rint(xdiv(xsub(ax, bx), L))));
        By = xadd(by, xmul(L, rint(xdiv(xsub(ay, by), L))));
        Bz = xadd(bz, xmul(hz, rint(xdiv(xsub(az, bz), hz))));
    }
    const double dx = xsub(ax, Bx), dy = xsub(ay, By), dz = xsub(az, Bz);
    return xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
}
// get_volume src/Configurations.jl:246-256
__device__ __forceinline__ double get_volume(int bc_kind, double L, double h) {
    if (bc_kind == PTMC_BC_CUBIC) return xmul(xmul(L, L), L);
    if (bc_kind == PTMC_BC_RHOMBIC) return xdiv(xmul(xmul(xmul(L, L), h), sqrt(3.0)), 2.0);
    return xmul(xmul(L, L), h);
}
// get_r_cut src/Ensembles.jl:118-125,153-155 (squared)
__device__ __forceinline__ double get_r_cut(int bc_kind, double L, double h) {
    if (bc_kind == PTMC_BC_CUBIC) return xdiv(xmul(L, L), 4.0);
    if (bc_kind == PTMC_BC_RHOMBIC) { double a = xdiv(xmul(xmul(L, L), 3.0), 16.0), b = xdiv(xmul(h, h), 4.0); return a < b ? a : b; }
    double a = xdiv(xmul(L, L), 4.0), b = xdiv(xmul(h, h), 4.0);
    return a < b ? a : b;
}

// --------------------------------------------------------------------------------------- Metropolis --
// acc_test! src/MCRun.jl:122-132 with metropolis_condition src/Exchange.jl:53-57: accept iff
// min(1, exp(x)) >= u, u in [0,1).  exp(x) is classified with an FP32 interval first; the FP64 exp is
// evaluated only when u falls inside the interval (probability ~1e-5), so the decision is exactly the FP64 one.
__device__ __forceinline__ bool metropolis_accept(double x, double u) {
    if (!(x < 0.0)) return !(x != x);          // exp(x) >= 1 > u ; NaN -> reject (NaN >= u is false)
    if (x > -30.0) {
        float pf = __expf((float)x);
        float uf = (float)u;
        if (uf < pf * 0.9999f) return true;
        if (uf > pf * 1.0001f) return false;
    } else if (u > 1e-12) {
        return false;                           // exp(-30) < 1e-13
    }
    return exp(x) >= u;
}

}  // namespace ptmc
