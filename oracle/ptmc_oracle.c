/*
 * ptmc_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement, in the reference's own operation order, of the parallel-tempering sweep of
 * ParallelTemperingMonteCarlo.jl (Julia; not runnable here: no julia binary).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the shipped path (libptmc_b200.so) never links or calls it.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: Julia does not contract a*b+c, so neither may this file.
 *
 * Pinning (tests/test_oracle_kat.py): ELJ/ELJEven exact values test/test_potentials.jl:6-15; EAM 3-atom
 * energy :55; seeded (Random.seed!(1234)) atom-move Delta E test/test_states.jl:38,95; distance2
 * Cubic/Rhombic exact values test/runtests.jl:96-100,133-140; get_volume :198-206; get_r_cut :34-47;
 * TempGrid :216-235; rand(1:n) via the (3,6) swap pair of test/test_states.jl:179-186.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#include "../include/ptmc_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* RuNNer potential: oracle/runner_oracle.c */
typedef struct orc_runner orc_runner;
orc_runner* orc_runner_new(const ptmc_runner_desc* d);
void orc_runner_free(orc_runner* r);
int orc_runner_nsym(const orc_runner* r);
double orc_runner_boundary2(const orc_runner* r);
double orc_runner_rcut(const orc_runner* r);
double orc_cutoff(double r_ij, double r_cut);
void orc_runner_fmatrix(const orc_runner* r, const double* d2, int n, double* f);
void orc_runner_init_g(const orc_runner* r, const double* pos, const double* d2, const double* f, int n, double* g);
void orc_runner_forward(const orc_runner* r, const double* g, int n, double* eatom);
/* two species (RuNNerPotential2Atom), oracle/runner_oracle.c */
typedef struct orc_runner2 orc_runner2;
orc_runner2* orc_runner2_new(const ptmc_runner2_desc* d);
void orc_runner2_free(orc_runner2* r);
int orc_runner2_nsym(const orc_runner2* r);
int orc_runner2_n1(const orc_runner2* r);
double orc_runner2_rcut(const orc_runner2* r);
void orc_runner2_init_g(const orc_runner2* r, const double* pos, const double* d2, const double* f, int n, double* g);
void orc_runner2_forward(const orc_runner2* r, const double* g, int n, double* eatom);
void orc_runner2_delta(const orc_runner2* r, const double* pos, const double* trial, const double* d2, const double* nd2,
                       const double* f, const double* nf, int i, int n, double* g);
void orc_runner2_swap(const orc_runner2* r, const double* pos, const double* d2, const double* f, int a1, int a2, int n, double* g);
void orc_runner_delta(const orc_runner* r, const double* pos, const double* newpos, const double* d2, const double* new_d2,
                      const double* f, const double* new_f, int i, int n, double* g);

/* ======================================================================================= SHA-256 == */
/* Needed only for Random.seed!(n): Julia hashes the seed's little-endian UInt32 words. */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

/* SHA-256 of a message shorter than 56 bytes (one block) — all a seed ever needs. */
static void sha256_short(const uint8_t* msg, size_t len, uint8_t out[32]) {
    uint8_t blk[64];
    memset(blk, 0, sizeof blk);
    memcpy(blk, msg, len);
    blk[len] = 0x80;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) blk[63 - i] = (uint8_t)(bits >> (8 * i));
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
}

/* =========================================================================== Julia default RNG == */
/* Random.seed!(n), n < 2^64 (Julia >= 1.7; validated against test/test_states.jl:38): digest of the
 * seed's little-endian UInt32 words (at least one word), read back as four little-endian UInt64. */
void orc_julia_seed(uint64_t seed, uint64_t s[4]) {
    uint8_t msg[8];
    size_t len = 4;
    for (int i = 0; i < 8; ++i) msg[i] = (uint8_t)(seed >> (8 * i));
    if (seed >> 32) len = 8;
    uint8_t dig[32];
    sha256_short(msg, len, dig);
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | dig[8 * k + i];
        s[k] = v;
    }
}

/* Random.seed!(lo + hi * 2^64): little-endian UInt32 words, trailing zero words dropped (make_seed of Julia's Random) */
void orc_julia_seed128(uint64_t lo, uint64_t hi, uint64_t s[4]) {
    uint8_t msg[16];
    for (int i = 0; i < 8; ++i) { msg[i] = (uint8_t)(lo >> (8 * i)); msg[8 + i] = (uint8_t)(hi >> (8 * i)); }
    size_t len = 16;
    while (len > 4 && msg[len - 1] == 0 && msg[len - 2] == 0 && msg[len - 3] == 0 && msg[len - 4] == 0) len -= 4;
    uint8_t dig[32];
    sha256_short(msg, len, dig);
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | dig[8 * k + i];
        s[k] = v;
    }
}

typedef struct orc_rng {
    uint64_t s[4];
    const uint64_t* replay; /* non-NULL: consume pre-drawn raw draws instead of stepping s */
    int64_t pos, len;
    int exhausted;
    /* PTMC_RNG_BATCH8 (include/ptmc_b200.h): PTMC_RNG_LANES substreams stepping together once per batch */
    int batch_mode;
    uint64_t sub[PTMC_RNG_LANES][4];
    uint64_t buf[PTMC_RNG_LANES];
} orc_rng;

static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* One Xoshiro256++ draw. */
uint64_t orc_xoshiro_next(uint64_t s[4]) {
    uint64_t res = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return res;
}

static uint64_t rng_u64(orc_rng* r) {
    if (r->replay) {
        if (r->pos >= r->len) { r->exhausted = 1; return 0; }
        return r->replay[r->pos++];
    }
    return orc_xoshiro_next(r->s);
}
/* rand()::Float64 */
static double rng_rand(orc_rng* r) { return (double)(rng_u64(r) >> 11) * 0x1.0p-53; }
/* rand(1:n): Lemire nearly-divisionless on one UInt64 draw (SamplerRangeNDL). */
static int64_t rng_range(orc_rng* r, uint64_t n) {
    uint64_t x = rng_u64(r);
    __uint128_t m = (__uint128_t)x * n;
    uint64_t l = (uint64_t)m;
    if (l < n) {
        uint64_t t = (0 - n) % n;
        while (l < t && !r->exhausted) {
            x = rng_u64(r);
            m = (__uint128_t)x * n;
            l = (uint64_t)m;
        }
    }
    return (int64_t)(m >> 64) + 1;
}

/* BATCH8 mode: step every substream once; position k of the batch is substream k's output */
static void rng_batch_begin(orc_rng* r) {
    for (int k = 0; k < PTMC_RNG_LANES; ++k) r->buf[k] = orc_xoshiro_next(r->sub[k]);
}
static double rng_batch_rand(const orc_rng* r, int k) { return (double)(r->buf[k] >> 11) * 0x1.0p-53; }

/* exported thin wrappers for the KAT tests */
double orc_rand_from_state(uint64_t s[4]) { return (double)(orc_xoshiro_next(s) >> 11) * 0x1.0p-53; }
int64_t orc_range_from_state(uint64_t s[4], uint64_t n) {
    orc_rng r; memset(&r, 0, sizeof r); memcpy(r.s, s, 32);
    int64_t v = rng_range(&r, n); memcpy(s, r.s, 32); return v;
}

/* ================================================================================== potentials == */
/* dimer_energy(::ELJPotential) src/EnergyEvaluation.jl:923-932 */
double orc_dimer_energy_elj(const double* c, int n, double r2) {
    double r = sqrt(r2);
    double r6inv = 1 / (r2 * r2 * r2);
    double sum1 = 0.0;
    for (int i = 0; i < n; ++i) { sum1 += c[i] * r6inv; r6inv /= r; }
    return sum1;
}
/* dimer_energy(::ELJPotentialEven) src/EnergyEvaluation.jl:934-942 */
double orc_dimer_energy_eljeven(const double* c, int n, double r2) {
    double r6inv = 1 / (r2 * r2 * r2);
    double sum1 = 0.0;
    for (int i = 0; i < n; ++i) { sum1 += c[i] * r6inv; r6inv /= r2; }
    return sum1;
}
/* lrc(::ELJPotentialEven) src/EnergyEvaluation.jl:968-982 */
double orc_lrc_eljeven(const double* c, int n, int n_atoms, double r_cut) {
    if (r_cut <= 50) return 0.0;
    double r_cut_sqrt = pow(r_cut, 0.5);
    double rc3 = r_cut * r_cut_sqrt;
    double e_lrc = 0.0;
    for (int i = 1; i <= n; ++i) { e_lrc += c[i - 1] / rc3 / (2 * i + 1); rc3 *= r_cut; }
    e_lrc *= M_PI * (double)((int64_t)n_atoms * n_atoms) / 4 / (r_cut_sqrt * r_cut_sqrt * r_cut_sqrt);
    return e_lrc;
}
/* EmbeddedAtomPotential(n,m,eps,C,a) src/EnergyEvaluation.jl:1069-1073 */
void orc_eam_ctor(double n, double m, double eps, double C, double a, double* ean, double* ecam) {
    *ean = eps * pow(a, n) / 2;
    *ecam = eps * C * pow(a, m / 2);
}
/* invrexp src/EnergyEvaluation.jl:1088-1095 */
static void invrexp(double r2, double n, double m, double* phi, double* rho) {
    if (r2 != 0.0) {
        double r_term = 1 / sqrt(r2);
        *phi = pow(r_term, n); *rho = pow(r_term, m);
    } else { *phi = 0.0; *rho = 0.0; }
}

/* ==================================================================================== geometry == */
#define SQRT3 (pow(3.0, 0.5)) /* the reference writes 3^0.5 everywhere */

/* distance2(a,b) = (a-b).(a-b) src/Configurations.jl:101 */
static double dist2_free(const double* a, const double* b) {
    double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return x * x + y * y + z * z;
}
/* src/Configurations.jl:104-109 (Cubic), :121-126 (Rectangular: L,L,h) */
static double dist2_rect(const double* a, const double* b, double L, double h) {
    double B[3];
    B[0] = b[0] + L * rint((a[0] - b[0]) / L);
    B[1] = b[1] + L * rint((a[1] - b[1]) / L);
    B[2] = b[2] + h * rint((a[2] - b[2]) / h);
    return dist2_free(a, B);
}
/* src/Configurations.jl:110-120 */
static double dist2_rhombic(const double* a, const double* b, double L, double h) {
    double s3 = SQRT3;
    double B[3];
    B[1] = b[1] + (s3 / 2 * L) * rint((a[1] - b[1]) / (s3 / 2 * L));
    B[0] = b[0] - b[1] / s3 + L * rint(((a[0] - b[0]) - 1 / s3 * (a[1] - b[1])) / L) + 1 / s3 * B[1];
    B[2] = b[2] + h * rint((a[2] - b[2]) / h);
    return dist2_free(a, B);
}
double orc_distance2(int bc_kind, const double* a, const double* b, double L, double h) {
    switch (bc_kind) {
        case PTMC_BC_SPHERICAL: return dist2_free(a, b);
        case PTMC_BC_CUBIC: return dist2_rect(a, b, L, L);
        case PTMC_BC_RECTANGULAR: return dist2_rect(a, b, L, h);
        default: return dist2_rhombic(a, b, L, h);
    }
}
/* get_volume src/Configurations.jl:246-256 */
/* get_tan src/Configurations.jl:177-219: |tan| of the angle between a -> (nearest image of) b and the z axis */
double orc_get_tan(int bc_kind, const double* a, const double* b, double L, double h) {
    double bx = b[0], by = b[1], bz = b[2];
    if (bc_kind == PTMC_BC_CUBIC) {
        bx = b[0] + L * rint((a[0] - b[0]) / L); by = b[1] + L * rint((a[1] - b[1]) / L); bz = b[2] + L * rint((a[2] - b[2]) / L);
    } else if (bc_kind == PTMC_BC_RHOMBIC) {
        by = b[1] + (pow(3, 0.5) / 2 * L) * rint((a[1] - b[1]) / (pow(3, 0.5) / 2 * L));
        bx = b[0] - b[1] / pow(3, 0.5) + L * rint(((a[0] - b[0]) - 1 / pow(3, 0.5) * (a[1] - b[1])) / L) + 1 / pow(3, 0.5) * by;
        bz = b[2] + h * rint((a[2] - b[2]) / h);
    } else if (bc_kind == PTMC_BC_RECTANGULAR) {
        bx = b[0] + L * rint((a[0] - b[0]) / L); by = b[1] + L * rint((a[1] - b[1]) / L); bz = b[2] + h * rint((a[2] - b[2]) / h);
    }
    double dx = a[0] - bx, dy = a[1] - by;
    double t = pow(dx * dx + dy * dy, 0.5) / (a[2] - bz);
    return fabs(t);
}
/* dimer_energy(::ELJPotentialB) src/EnergyEvaluation.jl:943-958 */
double orc_dimer_energy_eljb(const double* ca, const double* cb, const double* cc, int n, double r2, double z_angle) {
    double sum1;
    if (r2 >= 5.30) {
        double r6inv = 1 / (r2 * r2 * r2);
        double t2 = 2 / (z_angle * z_angle + 1) - 1;     /* cos(2 theta) */
        double t4 = 2 * (t2 * t2) - 1;
        sum1 = cc[0] * r6inv * (1 + ca[0] * t2 + cb[0] * t4);
        r6inv /= r2;
        for (int i = 1; i < n; ++i) {
            sum1 += cc[i] * r6inv * (1 + ca[i] * t2 + cb[i] * t4);
            r6inv /= pow(r2, 0.5);
        }
    } else {
        sum1 = 0.1;
    }
    return sum1;
}
/* dimer_energy(::LookupTablePotential) src/EnergyEvaluation.jl:274-295; table[(i_angle - 1) * l_dist + (j_dist - 1)] */
double orc_dimer_energy_lut(const double* table, int l_dist, int l_angle, double start_dist, double d_dist, double d_angle,
                            double c6coeff, double r2, double tan_) {
    (void)l_angle;
    int angle_index = 1;
    if (fabs(tan_) > 0.00872687153 && fabs(tan_) <= 114.592845357)
        angle_index = (int)rint(atan(fabs(tan_)) / d_angle * 180.0 / M_PI + 1.0);      /* round(Int32, .): ties to even */
    else if (fabs(tan_) > 114.592845357)
        angle_index = 91;
    double e;
    const double lo = start_dist + 0.5 * d_dist, hi = start_dist + l_dist * d_dist;
    if (r2 <= lo * lo) {
        e = table[(size_t)(angle_index - 1) * l_dist];
    } else if (r2 <= hi * hi) {
        int dist_index = (int)rint((pow(r2, 0.5) - start_dist) / d_dist);
        e = table[(size_t)(angle_index - 1) * l_dist + (dist_index - 1)];
    } else {
        e = c6coeff / (r2 * r2 * r2);
    }
    return e;
}
/* lrc(::ELJPotentialB) src/EnergyEvaluation.jl:983-998 */
double orc_lrc_eljb(int n_atoms, double r_cut) {
    static const double coeff[4] = {-0.1279111890228638, -1.328138539967966, 12.260941135261255, 41.12212408251662};
    if (r_cut <= 16) return 0.0;
    double r_cut_sqrt = pow(r_cut, 0.5);
    double rc3 = r_cut * r_cut_sqrt;
    double e_lrc = 0.0;
    for (int i = 1; i <= 4; ++i) { e_lrc += coeff[i - 1] / rc3 / (2 * i + 1); rc3 *= r_cut; }
    e_lrc *= M_PI * ((double)n_atoms * n_atoms) / 4 / (r_cut_sqrt * r_cut_sqrt * r_cut_sqrt);
    return e_lrc;
}

/* lrc(::LookupTablePotential) src/EnergyEvaluation.jl:999-1009 (1.0, not 0.0, below r_cut = 10: verbatim) */
double orc_lrc_lut(double c6coeff, int n_atoms, double r_cut) {
    if (r_cut <= 10) return 1.0;
    double r_cut_sqrt = pow(r_cut, 0.5);
    double rc3 = r_cut * r_cut_sqrt;
    double e_lrc = c6coeff / rc3 / 3;
    e_lrc *= M_PI * ((double)n_atoms * n_atoms) / 4 / (r_cut_sqrt * r_cut_sqrt * r_cut_sqrt);
    return e_lrc;
}

double orc_get_volume(int bc_kind, double L, double h) {
    switch (bc_kind) {
        case PTMC_BC_CUBIC: return L * L * L;
        case PTMC_BC_RHOMBIC: return L * L * h * SQRT3 / 2;
        case PTMC_BC_RECTANGULAR: return L * L * h;
        default: return 0.0;
    }
}
/* get_r_cut src/Ensembles.jl:118-125,153-155 (squared cut-off) */
double orc_get_r_cut(int bc_kind, double L, double h) {
    switch (bc_kind) {
        case PTMC_BC_CUBIC: return L * L / 4;
        case PTMC_BC_RHOMBIC: { double a = L * L * 3 / 16, b = h * h / 4; return a < b ? a : b; }
        case PTMC_BC_RECTANGULAR: { double a = L * L / 4, b = h * h / 4; return a < b ? a : b; }
        default: return INFINITY;
    }
}
/* TempGrid{N}(ti,tf) geometric src/InputParams.jl:82-93, kB :19 */
void orc_tempgrid(double ti, double tf, int n, double* t, double* beta) {
    const double kB = 3.16681196E-6;
    for (int i = 1; i <= n; ++i) {
        t[i - 1] = ti * pow(tf / ti, (double)(i - 1) / (double)(n - 1));
        beta[i - 1] = 1.0 / (kB * t[i - 1]);
    }
}

/* ======================================================================================= state == */
typedef struct orc_cfgside { /* everything exc_trajectories! swaps (src/Exchange.jl:115-124) */
    double* pos;      /* [N][3] config.positions */
    double box[2];    /* radius2 | box_length, box_height */
    double* d2;       /* [N][N] dist2_mat */
    double en_tot;
    double r_cut;     /* NPTVariables.r_cut (squared) */
    double* trial_pos;/* NPTVariables.trial_config positions */
    double trial_box[2];
    double* new_d2mat;/* NPTVariables.new_dist2_mat */
    double new_r_cut;
    double* tan;      /* ELJPotentialBVariables.tan_mat [N][N] (src/EnergyEvaluation.jl:187-192) */
    double* comp;     /* EAM component_vector [N][2] */
    double* new_comp; /* EAM new_component_vector */
    /* NNPVariables src/EnergyEvaluation.jl:1229-1236 */
    double *fmat, *new_f, *g, *new_g, *en_atom, *new_en_atom;
} orc_cfgside;

typedef struct orc_state { /* MCState src/MCStates.jl:33-53; slot-side fields stay put at exchange */
    double beta;
    double max_displ[4];
    double max_boxlength;
    double max_boxheight, lh_ratio;   /* src/MCStates.jl:106-107 */
    double ham[2];
    int64_t count[PTMC_NCOUNT];
    double* new_d2;   /* new_dist2_vec */
    double* new_tan;  /* ELJPotentialBVariables.new_tan_vec */
    double new_en;
    orc_rng rng;
    orc_cfgside* c;
} orc_state;

typedef struct orc_handle {
    ptmc_config cfg;
    double* betas;
    int64_t n_rep;
    orc_state* st;       /* [n_ladders][n_traj] */
    orc_cfgside* cs;     /* storage, pointed to by st[].c */
    orc_rng* lrng;       /* [n_ladders] driver-task stream */
    /* histograms */
    double *en_min, *en_max, *v_min, *v_max; /* [n_ladders] */
    double* en_hist;     /* [n_ladders][n_traj][n_bin+2] */
    double* ev_hist;     /* [n_ladders][n_traj][n_bin+2][n_bin+2] (NPT) */
    int hist_ready;
    /* trace */
    int64_t trace_cap, trace_n;
    int32_t* tr_index; double* tr_de; uint8_t* tr_acc; int32_t* tr_exc;
    int error;
    int n_threads;
    orc_runner* runner;
    orc_runner2* runner2;
    double *eljb_a, *eljb_b;   /* copies of cfg.eljb_a / eljb_b */
    double* lut;               /* copy of cfg.lut_table */
} orc_handle;

/* length(MoveStrategy) src/Ensembles.jl:236-282 */
static int n_steps_of(const ptmc_config* c) {
    return c->n_atom_moves + (c->ensemble == PTMC_ENS_NPT ? c->n_volume_moves : 0) + (c->ensemble == PTMC_ENS_NNVT ? c->n_atom_swaps : 0);
}

/* ------------------------------------------------------------------------------ pair energies -- */
static double pair_energy(const ptmc_config* c, double r2) {
    return c->pot_kind == PTMC_POT_ELJ ? orc_dimer_energy_elj(c->coeff, c->n_coeff, r2)
                                       : orc_dimer_energy_eljeven(c->coeff, c->n_coeff, r2);
}
/* dimer_energy_atom src/EnergyEvaluation.jl:313-322 (no cut) / :323-336 (d2 <= r_cut) */
static double dimer_energy_atom(const ptmc_config* c, int i, const double* d2vec, int n, int use_cut, double r_cut) {
    double sum1 = 0.0;
    for (int j = 0; j < i; ++j)
        if (!use_cut || d2vec[j] <= r_cut) sum1 += pair_energy(c, d2vec[j]);
    for (int j = i + 1; j < n; ++j)
        if (!use_cut || d2vec[j] <= r_cut) sum1 += pair_energy(c, d2vec[j]);
    return sum1;
}
/* dimer_energy_config src/EnergyEvaluation.jl:378-396 (NVT), :397-420 (Cubic), :421-445 (Rhombic),
 * :446-474 (Rectangular) */
static double dimer_energy_config(const ptmc_config* c, const double* d2, int n, int use_cut, double r_cut,
                                  const double* box) {
    double energy_tot = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (!use_cut || d2[(size_t)i * n + j] <= r_cut) energy_tot += pair_energy(c, d2[(size_t)i * n + j]);
    if (!use_cut) return energy_tot;
    double l = orc_lrc_eljeven(c->coeff, c->n_coeff, n, r_cut);
    switch (c->bc_kind) {
        case PTMC_BC_CUBIC: return energy_tot + l;
        case PTMC_BC_RHOMBIC: return energy_tot + l * 3 / 4 * box[0] / box[1];
        default: /* rectangular */
            if (box[0] < box[1]) return energy_tot + l * box[0] / box[1];
            return energy_tot + l * (box[1] * box[1]) / (box[0] * box[0]);
    }
}
/* calc_energies_from_components src/EnergyEvaluation.jl:1156-1162 */
static double eam_energy(const ptmc_config* c, const double* comp, int n) {
    double en_val = 0.0;
    for (int i = 0; i < n; ++i) en_val += c->eam_ean * comp[2 * i] - c->eam_ecam * sqrt(comp[2 * i + 1]);
    return en_val;
}

static void d2_matrix(const ptmc_config* c, const double* pos, const double* box, double* d2, int n) {
    /* get_distance2_mat! src/Configurations.jl:146-162 */
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            d2[(size_t)i * n + j] = d2[(size_t)j * n + i] = orc_distance2(c->bc_kind, pos + 3 * i, pos + 3 * j, box[0], box[1]);
    for (int i = 0; i < n; ++i) d2[(size_t)i * n + i] = 0;
}

/* dimer_energy of the potentials that depend on the angle to the z axis (AbstractDimerPotentialB: ELJPotentialB, LookupTablePotential) */
static double pair_energy_b(const ptmc_config* c, double r2, double tan_) {
    if (c->pot_kind == PTMC_POT_LUT)
        return orc_dimer_energy_lut(c->lut_table, c->lut_l_dist, c->lut_l_angle, c->lut_start_dist, c->lut_d_dist, c->lut_d_angle, c->lut_c6coeff, r2, tan_);
    return orc_dimer_energy_eljb(c->eljb_a, c->eljb_b, c->coeff, c->n_coeff, r2, tan_);
}
/* dimer_energy_atom with tangents src/EnergyEvaluation.jl:337-346 (no cut), :347-362 (d2 <= r_cut) */
static double eljb_energy_atom(const ptmc_config* c, int i, const double* d2vec, const double* tanvec, int n, int use_cut, double r_cut) {
    double sum1 = 0.0;
    for (int j = 0; j < i; ++j) if (!use_cut || d2vec[j] <= r_cut) sum1 += pair_energy_b(c, d2vec[j], tanvec[j]);
    for (int j = i + 1; j < n; ++j) if (!use_cut || d2vec[j] <= r_cut) sum1 += pair_energy_b(c, d2vec[j], tanvec[j]);
    return sum1;
}
/* dimer_energy_config for the tangent-dependent potentials: :476-492 (no cut), Cubic :494-516 / :673-696, Rhombic :519-543 /
 * :697-746, Rectangular :590-636 / :747-789.  tan_mat and new_tan_mat are one array in the reference (set_variables :1790,1798),
 * so the "new_tan_mat != tan_mat" branches never differ; uniform volume moves leave the tangents unchanged. */
static double config_energy_b(const ptmc_config* c, const double* d2, const double* tanm, int n, int use_cut, double r_cut, const double* box) {
    double energy_tot = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (!use_cut || d2[(size_t)i * n + j] <= r_cut) energy_tot += pair_energy_b(c, d2[(size_t)i * n + j], tanm[(size_t)i * n + j]);
    if (!use_cut) return energy_tot;
    double l = c->pot_kind == PTMC_POT_LUT ? orc_lrc_lut(c->lut_c6coeff, n, r_cut) : orc_lrc_eljb(n, r_cut);
    switch (c->bc_kind) {
        case PTMC_BC_CUBIC: return energy_tot + l;
        case PTMC_BC_RHOMBIC: return energy_tot + l * 3 / 4 * box[0] / box[1];
        default:
            if (box[0] < box[1]) return energy_tot + l * box[0] / box[1];
            return energy_tot + l * (box[1] * box[1]) / (box[0] * box[0]);
    }
}

/* MCState(temp,beta,config,ensemble,pot) src/MCStates.jl:137-162 */
static void init_cfgside_runner(const orc_runner* R, orc_cfgside* s, int n) {
    /* set_variables(::RuNNerPotential) src/EnergyEvaluation.jl:1812-1822 + initialise_energy :1735-1747 */
    orc_runner_fmatrix(R, s->d2, n, s->fmat);
    orc_runner_init_g(R, s->pos, s->d2, s->fmat, n, s->g);
    memcpy(s->new_g, s->g, sizeof(double) * (size_t)n * orc_runner_nsym(R));
    orc_runner_forward(R, s->g, n, s->en_atom);
    double e = 0.0;
    for (int i = 0; i < n; ++i) e += s->en_atom[i];
    s->en_tot = e;
    memset(s->new_en_atom, 0, sizeof(double) * n);
    memset(s->new_f, 0, sizeof(double) * n);
}

/* set_variables(::RuNNerPotential2Atom) src/EnergyEvaluation.jl:1823-1857 + initialise_energy :1748-1768 */
static void init_cfgside_runner2(const orc_runner2* R, orc_cfgside* s, int n) {
    const size_t ns = (size_t)orc_runner2_nsym(R);
    for (size_t k = 0; k < (size_t)n * n; ++k) s->fmat[k] = orc_cutoff(sqrt(s->d2[k]), orc_runner2_rcut(R));
    orc_runner2_init_g(R, s->pos, s->d2, s->fmat, n, s->g);
    memset(s->new_g, 0, sizeof(double) * (size_t)n * ns);
    orc_runner2_forward(R, s->g, n, s->en_atom);
    double e = 0.0;
    for (int i = 0; i < n; ++i) e += s->en_atom[i];
    s->en_tot = e;
    memset(s->new_en_atom, 0, sizeof(double) * n);
    memset(s->new_f, 0, sizeof(double) * n);
}

static void init_cfgside(const ptmc_config* c, orc_cfgside* s) {
    int n = c->n_atoms;
    d2_matrix(c, s->pos, s->box, s->d2, n);
    int npt = c->ensemble == PTMC_ENS_NPT;
    s->r_cut = npt ? orc_get_r_cut(c->bc_kind, s->box[0], s->box[1]) : 0.0;
    s->new_r_cut = 0.0;
    if (c->pot_kind == PTMC_POT_EAM) {
        /* set_variables(::EmbeddedAtomPotential) src/EnergyEvaluation.jl:1800-1811 */
        for (int i = 0; i < n; ++i) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < n; ++j) {
                double p, r; invrexp(s->d2[(size_t)i * n + j], c->eam_n, c->eam_m, &p, &r);
                a += p; b += r;
            }
            s->comp[2 * i] = a; s->comp[2 * i + 1] = b;
            s->new_comp[2 * i] = 0; s->new_comp[2 * i + 1] = 0;
        }
        s->en_tot = eam_energy(c, s->comp, n); /* initialise_energy :1722-1734 */
    } else if (c->pot_kind == PTMC_POT_RUNNER || c->pot_kind == PTMC_POT_RUNNER2) {
        /* done by the caller (needs the handle's orc_runner / orc_runner2) */
    } else if (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT) {
        /* set_variables(::ELJPotentialB / ::LookupTablePotential): get_tantheta_mat! src/Configurations.jl:229-244; dimer_energy_config :476-492 */
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j)
                s->tan[(size_t)i * n + j] = s->tan[(size_t)j * n + i] = orc_get_tan(c->bc_kind, s->pos + 3 * i, s->pos + 3 * j, s->box[0], s->box[1]);
        for (int i = 0; i < n; ++i) s->tan[(size_t)i * n + i] = 0;
        s->en_tot = config_energy_b(c, s->d2, s->tan, n, npt, s->r_cut, s->box);
    } else {
        s->en_tot = dimer_energy_config(c, s->d2, n, npt, s->r_cut, s->box); /* :1662-1693 */
    }
}

/* ------------------------------------------------------------------------------------- moves -- */
/* atom_displacement(pos, max_displacement, bc) src/MCMoves.jl:33-98 */
static void displace_u(const ptmc_config* c, const double* pos, double dmax, const double* box, const double* u, double* t) {
    double dx = (u[0] - 0.5) * dmax;
    double dy = (u[1] - 0.5) * dmax;
    double dz = (u[2] - 0.5) * dmax;
    t[0] = pos[0] + dx; t[1] = pos[1] + dy; t[2] = pos[2] + dz;
    double L = box[0], h = box[1];
    switch (c->bc_kind) {
        case PTMC_BC_SPHERICAL: break;
        case PTMC_BC_CUBIC: {
            double a = rint(t[0] / L), b = rint(t[1] / L), g = rint(t[2] / L);
            t[0] -= L * a; t[1] -= L * b; t[2] -= L * g;
        } break;
        case PTMC_BC_RECTANGULAR: {
            double a = L * rint(t[0] / L), b = L * rint(t[1] / L), g = h * rint(t[2] / h);
            t[0] -= a; t[1] -= b; t[2] -= g;
        } break;
        default: { /* Rhombic :66-83 */
            double s3 = SQRT3;
            double ry = rint((t[1] - L * s3 / 4) / (L * s3 / 2));
            double sx = L * rint((t[0] - t[1] / s3 - L / 2) / L) + L / 2 * ry;
            double sy = L * s3 / 2 * ry;
            double sz = h * rint((t[2] - h / 2) / h);
            t[0] -= sx; t[1] -= sy; t[2] -= sz;
        }
    }
}

/* sequential-stream form: draw order x, y, z (src/MCMoves.jl:35-37) */
static void displace(const ptmc_config* c, const double* pos, double dmax, const double* box, orc_rng* r, double* t) {
    double u[3];
    u[0] = rng_rand(r); u[1] = rng_rand(r); u[2] = rng_rand(r);
    displace_u(c, pos, dmax, box, u, t);
}
/* BATCH8 form: uniforms from positions k0..k0+2 of the current batch */
static void displace_b(const ptmc_config* c, const double* pos, double dmax, const double* box, const orc_rng* r, int k0, double* t) {
    double u[3];
    u[0] = rng_batch_rand(r, k0); u[1] = rng_batch_rand(r, k0 + 1); u[2] = rng_batch_rand(r, k0 + 2);
    displace_u(c, pos, dmax, box, u, t);
}

/* recentre! src/Configurations.jl:79-85 (mean = sum/n, sequential) */
static void recentre(double* pos, int n) {
    double cx = 0, cy = 0, cz = 0;
    for (int i = 0; i < n; ++i) cx += pos[3 * i];
    for (int i = 0; i < n; ++i) cy += pos[3 * i + 1];
    for (int i = 0; i < n; ++i) cz += pos[3 * i + 2];
    cx /= n; cy /= n; cz /= n;
    for (int i = 0; i < n; ++i) { pos[3 * i] -= cx; pos[3 * i + 1] -= cy; pos[3 * i + 2] -= cz; }
}

/* metropolis_condition src/Exchange.jl:53-57 */
static double metropolis(double delta_energy, double beta) {
    double p = exp(-delta_energy * beta);
    return p > 1 ? 1.0 : p;
}

/* One mc_move! (src/MCRun.jl:142-162).  Returns 0, or a negative ptmc_status. */
static int mc_move(orc_handle* H, orc_state* st, int32_t* tr_i, double* tr_de, uint8_t* tr_a) {
    const ptmc_config* c = &H->cfg;
    orc_cfgside* s = st->c;
    const int n = c->n_atoms;
    const int npt = c->ensemble == PTMC_ENS_NPT;
    const int bm = st->rng.batch_mode;
    int index;
    if (bm) {   /* BATCH8 draw schedule, see ptmc_rng_mode in include/ptmc_b200.h */
        rng_batch_begin(&st->rng);
        index = (int)(((__uint128_t)st->rng.buf[0] * (uint64_t)n_steps_of(c)) >> 64) + 1;
    } else {
        index = (int)rng_range(&st->rng, (uint64_t)n_steps_of(c)); /* :145 */
    }
    double trial[3];
    int accepted = 0;
    if (index <= c->n_atom_moves) {
        const int i = index - 1;
        /* generate_move! -> atom_displacement(mc_state) src/MCMoves.jl:100-113 / :115-145 */
        if (c->bc_kind == PTMC_BC_SPHERICAL) {
            double count = 0.0;
            int redraws = 0;
            if (bm) displace_b(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, 1, trial);
            else displace(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, trial);
            while (trial[0] * trial[0] + trial[1] * trial[1] + trial[2] * trial[2] > s->box[0]) { /* check_boundary */
                count += 1;
                if (count == 50) {
                    recentre(s->pos, n);
                } else {
                    if (bm) {
                        if (++redraws == 1) displace_b(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, 5, trial);
                        else { rng_batch_begin(&st->rng); displace_b(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, 0, trial); }
                    } else {
                        displace(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, trial);
                    }
                    if (count == 100) return PTMC_ERR_SPHERE;
                }
                if (st->rng.exhausted) return PTMC_ERR_REPLAY;
            }
        } else {
            if (bm) displace_b(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, 1, trial);
            else displace(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, trial);
        }
        for (int j = 0; j < n; ++j) st->new_d2[j] = orc_distance2(c->bc_kind, trial, s->pos + 3 * j, s->box[0], s->box[1]);
        st->new_d2[i] = 0.0;
        /* get_energy! src/MCRun.jl:47-87 -> energy_update! */
        if (c->pot_kind == PTMC_POT_RUNNER) {
            /* energy_update!(::RuNNerPotential) src/EnergyEvaluation.jl:1593-1621 */
            const orc_runner* R = H->runner;
            int too_close = 0;
            for (int j = 0; j < n; ++j) if (j != i && st->new_d2[j] < orc_runner_boundary2(R)) too_close = 1;
            if (too_close) {
                st->new_en = 100.0;
            } else {
                /* get_new_state_vars! :1238-1264, calc_new_runner_energy! :1269-1275 */
                for (int j = 0; j < n; ++j) s->new_f[j] = orc_cutoff(sqrt(st->new_d2[j]), orc_runner_rcut(R));
                memcpy(s->new_g, s->g, sizeof(double) * (size_t)n * orc_runner_nsym(R));
                orc_runner_delta(R, s->pos, trial, s->d2, st->new_d2, s->fmat, s->new_f, i, n, s->new_g);
                orc_runner_forward(R, s->new_g, n, s->new_en_atom);
                double e = 0.0;
                for (int j = 0; j < n; ++j) e += s->new_en_atom[j];
                st->new_en = e;
            }
        } else if (c->pot_kind == PTMC_POT_RUNNER2) {
            /* energy_update!(::RuNNerPotential2Atom) src/EnergyEvaluation.jl:1623-1643 (no boundary guard): get_new_state_vars!
             * :1347-1383 (new_g_matrix = g_matrix .+ calc_delta_matrix(zeros)), calc_new_runner_energy! :1418-1433 */
            const orc_runner2* R = H->runner2;
            const size_t ng = (size_t)n * orc_runner2_nsym(R);
            for (int j = 0; j < n; ++j) s->new_f[j] = orc_cutoff(sqrt(st->new_d2[j]), orc_runner2_rcut(R));
            memset(s->new_g, 0, sizeof(double) * ng);
            orc_runner2_delta(R, s->pos, trial, s->d2, st->new_d2, s->fmat, s->new_f, i, n, s->new_g);
            for (size_t q = 0; q < ng; ++q) s->new_g[q] = s->g[q] + s->new_g[q];
            orc_runner2_forward(R, s->new_g, n, s->new_en_atom);
            double e = 0.0;
            for (int j = 0; j < n; ++j) e += s->new_en_atom[j];
            st->new_en = e;
        } else if (c->pot_kind == PTMC_POT_EAM) {
            /* src/EnergyEvaluation.jl:1567-1591 + calc_components :1111-1127 */
            memcpy(s->new_comp, s->comp, sizeof(double) * 2 * n);
            for (int j = 0; j < n; ++j) {
                double pn, rn, po, ro;
                invrexp(st->new_d2[j], c->eam_n, c->eam_m, &pn, &rn);
                invrexp(s->d2[(size_t)i * n + j], c->eam_n, c->eam_m, &po, &ro);
                double t0 = pn - po, t1 = rn - ro;
                s->new_comp[2 * j] += t0; s->new_comp[2 * j + 1] += t1;
                s->new_comp[2 * i] += t0; s->new_comp[2 * i + 1] += t1;
            }
            st->new_en = eam_energy(c, s->new_comp, n);
        } else if (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT) {
            /* energy_update!(::AbstractDimerPotentialB / ::LookupTablePotential) src/EnergyEvaluation.jl:1508-1533, dimer_energy_update! :822-836 */
            for (int j = 0; j < n; ++j) st->new_tan[j] = orc_get_tan(c->bc_kind, trial, s->pos + 3 * j, s->box[0], s->box[1]);
            st->new_tan[i] = 0;
            double delta_en = eljb_energy_atom(c, i, st->new_d2, st->new_tan, n, npt, s->r_cut) -
                              eljb_energy_atom(c, i, s->d2 + (size_t)i * n, s->tan + (size_t)i * n, n, npt, s->r_cut);
            st->new_en = delta_en + s->en_tot;
        } else {
            /* dimer_energy_update! src/EnergyEvaluation.jl:804-821 */
            double delta_en = dimer_energy_atom(c, i, st->new_d2, n, npt, s->r_cut) -
                              dimer_energy_atom(c, i, s->d2 + (size_t)i * n, n, npt, s->r_cut);
            st->new_en = delta_en + s->en_tot;
        }
        /* acc_test! src/MCRun.jl:122-132 */
        double de = st->new_en - s->en_tot;
        double p = metropolis(de, st->beta);
        double u = bm ? rng_batch_rand(&st->rng, 4) : rng_rand(&st->rng);
        if (tr_i) { *tr_i = index; *tr_de = de; }
        if (p >= u) {
            /* swap_atom_config! src/swap_config.jl:40-51 */
            accepted = 1;
            s->pos[3 * i] = trial[0]; s->pos[3 * i + 1] = trial[1]; s->pos[3 * i + 2] = trial[2];
            for (int j = 0; j < n; ++j) { s->d2[(size_t)i * n + j] = st->new_d2[j]; s->d2[(size_t)j * n + i] = st->new_d2[j]; }
            double tmp = s->en_tot; s->en_tot = st->new_en; st->new_en = tmp;
            st->count[PTMC_C_ATOM_TOT] += 1; st->count[PTMC_C_ATOM_ADJ] += 1;
            if (c->pot_kind == PTMC_POT_EAM) { double* t = s->comp; s->comp = s->new_comp; s->new_comp = t; } /* swap_vars! :208-211 */
            if (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT)   /* swap_vars!(::ELJPotentialBVariables / ::LookupTableVariables) src/swap_config.jl:197-206 */
                for (int j = 0; j < n; ++j) { s->tan[(size_t)i * n + j] = st->new_tan[j]; s->tan[(size_t)j * n + i] = st->new_tan[j]; }
            if (c->pot_kind == PTMC_POT_RUNNER || c->pot_kind == PTMC_POT_RUNNER2) {   /* swap_vars!(::NNPVariables / ::NNPVariables2a) src/swap_config.jl:213-235 */
                double* t = s->en_atom; s->en_atom = s->new_en_atom; s->new_en_atom = t;
                t = s->g; s->g = s->new_g; s->new_g = t;
                for (int j = 0; j < n; ++j) { s->fmat[(size_t)i * n + j] = s->new_f[j]; s->fmat[(size_t)j * n + i] = s->new_f[j]; }
            }
        }
    } else if (c->ensemble == PTMC_ENS_NNVT) {
        /* "atomswap": swap_atoms src/MCMoves.jl:363-370 - i1 = rand(1:N1), i2 = rand((N1+1):N); BATCH8: positions 1 and 2 */
        const orc_runner2* R = H->runner2;
        const int n1 = orc_runner2_n1(R), n2 = n - n1;
        int i1, i2;
        if (bm) {
            i1 = (int)(((__uint128_t)st->rng.buf[1] * (uint64_t)n1) >> 64) + 1;
            i2 = n1 + (int)(((__uint128_t)st->rng.buf[2] * (uint64_t)n2) >> 64) + 1;
        } else {
            i1 = (int)rng_range(&st->rng, (uint64_t)n1);
            i2 = n1 + (int)rng_range(&st->rng, (uint64_t)n2);
        }
        const int a1 = i1 - 1, a2 = i2 - 1;
        /* get_energy! src/MCRun.jl:88-110 -> swap_energy_update src/EnergyEvaluation.jl:1437-1448: get_new_state_vars! :1385-1410 */
        const size_t ng = (size_t)n * orc_runner2_nsym(R);
        memset(s->new_g, 0, sizeof(double) * ng);
        orc_runner2_swap(R, s->pos, s->d2, s->fmat, a1, a2, n, s->new_g);
        for (size_t q = 0; q < ng; ++q) s->new_g[q] = s->g[q] + s->new_g[q];
        orc_runner2_forward(R, s->new_g, n, s->new_en_atom);
        double e = 0.0;
        for (int j = 0; j < n; ++j) e += s->new_en_atom[j];
        st->new_en = e;
        /* acc_test! src/MCRun.jl:122-132, metropolis_condition("atomswap") src/Exchange.jl:86-87 */
        double de = st->new_en - s->en_tot;
        double p = metropolis(de, st->beta);
        double u = bm ? rng_batch_rand(&st->rng, 4) : rng_rand(&st->rng);
        if (tr_i) { *tr_i = index + (i1 << 10) + (i2 << 20); *tr_de = de; }
        if (p >= u) {
            /* swap_move_config! src/swap_config.jl:241-276: energies, positions, rows then columns of dist2_mat and f_matrix
             * (diagonals reset), en_atom_vec <-> new_en_atom, g_matrix <-> new_g_matrix; no counter is touched */
            accepted = 1;
            double tmp = s->en_tot; s->en_tot = st->new_en; st->new_en = tmp;
            for (int k = 0; k < 3; ++k) { tmp = s->pos[3 * a1 + k]; s->pos[3 * a1 + k] = s->pos[3 * a2 + k]; s->pos[3 * a2 + k] = tmp; }
            double* mats[2] = {s->d2, s->fmat};
            for (int m = 0; m < 2; ++m) {
                double* M = mats[m];
                for (int k = 0; k < n; ++k) { tmp = M[(size_t)a1 * n + k]; M[(size_t)a1 * n + k] = M[(size_t)a2 * n + k]; M[(size_t)a2 * n + k] = tmp; }
                for (int k = 0; k < n; ++k) { tmp = M[(size_t)k * n + a1]; M[(size_t)k * n + a1] = M[(size_t)k * n + a2]; M[(size_t)k * n + a2] = tmp; }
                M[(size_t)a1 * n + a1] = M[(size_t)a2 * n + a2] = m == 0 ? 0.0 : 1.0;
            }
            double* t = s->en_atom; s->en_atom = s->new_en_atom; s->new_en_atom = t;
            t = s->g; s->g = s->new_g; s->new_g = t;
        }
    } else {
        int xy_or_z = 0;
        if (c->separated_volume) {
            /* volume_change_separated src/MCMoves.jl:300-343: z with 1/6, xy with 2/6, all directions with 3/6; the new
             * distance matrix is recomputed from the trial configuration (get_distance2_mat!) */
            int ra = bm ? (int)(((__uint128_t)st->rng.buf[2] * 6u) >> 64) + 1 : (int)rng_range(&st->rng, 6);
            double uv = bm ? rng_batch_rand(&st->rng, 1) : rng_rand(&st->rng);
            double sxy = 1.0, sz = 1.0;
            const double lh = s->box[0] / s->box[1];
            if (ra == 1) {            /* volume_change_z :256-271 */
                xy_or_z = 2;
                double scale = exp((uv - 0.5) * st->max_displ[3]);
                if (lh <= st->lh_ratio * 1.1 && scale > 1.0) scale = 1 / scale;
                else if (lh >= st->lh_ratio * 0.909 && scale < 1.0) scale = 1 / scale;
                if (s->box[1] >= st->max_boxheight && scale > 1.0) scale = 1 / scale;
                sz = scale;
            } else if (ra <= 3) {     /* volume_change_xy :232-247 */
                xy_or_z = 1;
                double scale = pow(exp((uv - 0.5) * st->max_displ[2]), 1.0 / 2.0);
                if (lh >= st->lh_ratio * 1.1 && scale > 1.0) scale = 1 / scale;
                else if (lh <= st->lh_ratio * 0.909 && scale < 1.0) scale = 1 / scale;
                if (s->box[0] >= st->max_boxlength && scale > 1.0) scale = 1 / scale;
                sxy = scale;
            } else {                  /* volume_change_xyz :169-177 */
                double scale = pow(exp((uv - 0.5) * st->max_displ[1]), 1.0 / 3.0);
                if (s->box[0] >= st->max_boxlength && scale > 1.0) scale = 1.0;
                sxy = scale; sz = scale;
            }
            if (xy_or_z == 0) {
                for (int k = 0; k < 3 * n; ++k) s->trial_pos[k] = sxy * s->pos[k];
                s->trial_box[0] = sxy * s->box[0]; s->trial_box[1] = sxy * s->box[1];
            } else if (xy_or_z == 1) {   /* scale_xy :189-201: p * scale */
                for (int k = 0; k < n; ++k) { s->trial_pos[3 * k] = s->pos[3 * k] * sxy; s->trial_pos[3 * k + 1] = s->pos[3 * k + 1] * sxy; s->trial_pos[3 * k + 2] = s->pos[3 * k + 2]; }
                s->trial_box[0] = s->box[0] * sxy; s->trial_box[1] = s->box[1];
            } else {                     /* scale_z :211-222 */
                for (int k = 0; k < n; ++k) { s->trial_pos[3 * k] = s->pos[3 * k]; s->trial_pos[3 * k + 1] = s->pos[3 * k + 1]; s->trial_pos[3 * k + 2] = s->pos[3 * k + 2] * sz; }
                s->trial_box[0] = s->box[0]; s->trial_box[1] = s->box[1] * sz;
            }
            s->new_r_cut = orc_get_r_cut(c->bc_kind, s->trial_box[0], s->trial_box[1]);
            d2_matrix(c, s->trial_pos, s->trial_box, s->new_d2mat, n);
            if (xy_or_z >= 1 && (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT)) {
                /* src/MCMoves.jl:333-341: get_tantheta_mat!(new_tan_mat, trial_config) - and new_tan_mat IS tan_mat (one array,
                 * set_variables src/EnergyEvaluation.jl:1790,1798): the trial's tangents stay in the state even if the move is rejected */
                for (int i = 0; i < n; ++i)
                    for (int j = i + 1; j < n; ++j)
                        s->tan[(size_t)i * n + j] = s->tan[(size_t)j * n + i] =
                            orc_get_tan(c->bc_kind, s->trial_pos + 3 * i, s->trial_pos + 3 * j, s->trial_box[0], s->trial_box[1]);
                for (int i = 0; i < n; ++i) s->tan[(size_t)i * n + i] = 0;
            }
        } else {
        /* volume_change_uniform src/MCMoves.jl:277-287, volume_change_xyz :169-177 */
        double uv = bm ? rng_batch_rand(&st->rng, 1) : rng_rand(&st->rng);
        double scale = pow(exp((uv - 0.5) * st->max_displ[1]), 1.0 / 3.0);
        if (s->box[0] >= st->max_boxlength && scale > 1.0) scale = 1.0;
        for (int k = 0; k < 3 * n; ++k) s->trial_pos[k] = scale * s->pos[k];
        s->trial_box[0] = scale * s->box[0];
        s->trial_box[1] = scale * s->box[1];
        s->new_r_cut = orc_get_r_cut(c->bc_kind, s->trial_box[0], s->trial_box[1]);
        double sc2 = scale * scale;
        for (size_t k = 0; k < (size_t)n * n; ++k) s->new_d2mat[k] = s->d2[k] * sc2;
        }
        /* get_energy! NPT else-branch src/MCRun.jl:76-85 */
        st->new_en = (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT)
                         ? config_energy_b(c, s->new_d2mat, s->tan, n, 1, s->new_r_cut, s->trial_box)
                         : dimer_energy_config(c, s->new_d2mat, n, 1, s->new_r_cut, s->trial_box);
        /* metropolis_condition NPT src/Exchange.jl:58-71,77-85 */
        double de = st->new_en - s->en_tot;
        double vc = orc_get_volume(c->bc_kind, s->trial_box[0], s->trial_box[1]);
        double vu = orc_get_volume(c->bc_kind, s->box[0], s->box[1]);
        double delta_h = de + c->pressure * (vc - vu);
        double p = exp(-delta_h * st->beta + c->n_atoms * log(vc / vu));
        p = p > 1 ? 1.0 : p;
        double u = bm ? rng_batch_rand(&st->rng, 4) : rng_rand(&st->rng);
        if (tr_i) { *tr_i = index; *tr_de = de; }
        if (p >= u) {
            /* swap_config_v! src/swap_config.jl:113-140 (xy_or_z == 0) */
            accepted = 1;
            memcpy(s->pos, s->trial_pos, sizeof(double) * 3 * n);
            s->box[0] = s->trial_box[0]; s->box[1] = s->trial_box[1];
            memcpy(s->d2, s->new_d2mat, sizeof(double) * (size_t)n * n);
            s->en_tot = st->new_en;
            if (xy_or_z == 0) { st->count[PTMC_C_VOL_TOT] += 1; st->count[PTMC_C_VOL_ADJ] += 1; }
            else if (xy_or_z == 1) { st->count[PTMC_C_VOLXY_TOT] += 1; st->count[PTMC_C_VOLXY_ADJ] += 1; }
            else { st->count[PTMC_C_VOLZ_TOT] += 1; st->count[PTMC_C_VOLZ_ADJ] += 1; }
            s->r_cut = s->new_r_cut;
        }
    }
    if (tr_a) *tr_a = (uint8_t)accepted;
    if (st->rng.exhausted) return PTMC_ERR_REPLAY;
    return 0;
}

/* update_max_stepsize! src/Exchange.jl:250-263 (NVT), :198-249 (NPT, separated_volume == false) */
static void update_max_stepsize(const ptmc_config* c, orc_state* st) {
    double acc_rate = (double)st->count[PTMC_C_ATOM_ADJ] / (double)((int64_t)c->n_adjust * c->n_atom_moves);
    if (acc_rate < c->min_acc) st->max_displ[0] *= 0.9;
    else if (acc_rate > c->max_acc) st->max_displ[0] *= 1.1;
    st->count[PTMC_C_ATOM_ADJ] = 0;
    if (c->ensemble == PTMC_ENS_NPT && !c->separated_volume) {
        acc_rate = (double)st->count[PTMC_C_VOL_ADJ] / (double)((int64_t)c->n_adjust * c->n_volume_moves);
        if (acc_rate < c->min_acc) st->max_displ[1] *= 0.9;
        else if (acc_rate > c->max_acc) st->max_displ[1] *= 1.1;
        st->count[PTMC_C_VOL_ADJ] = 0;
    } else if (c->ensemble == PTMC_ENS_NPT) {
        /* src/Exchange.jl:221-245: n_update * n_volume_moves * 1 / 2, ... * 1 / 3, ... / 6 (integer product, then Float64 division) */
        const double nv = (double)((int64_t)c->n_adjust * c->n_volume_moves);
        acc_rate = (double)st->count[PTMC_C_VOL_ADJ] / (nv / 2);
        if (acc_rate < c->min_acc) st->max_displ[1] *= 0.9;
        else if (acc_rate > c->max_acc) st->max_displ[1] *= 1.1;
        st->count[PTMC_C_VOL_ADJ] = 0;
        acc_rate = (double)st->count[PTMC_C_VOLXY_ADJ] / (nv / 3);
        if (acc_rate < c->min_acc) st->max_displ[2] *= 0.9;
        else if (acc_rate > c->max_acc) st->max_displ[2] *= 1.1;
        st->count[PTMC_C_VOLXY_ADJ] = 0;
        acc_rate = (double)st->count[PTMC_C_VOLZ_ADJ] / (nv / 6);
        if (acc_rate < c->min_acc) st->max_displ[3] *= 0.9;
        else if (acc_rate > c->max_acc) st->max_displ[3] *= 1.1;
        st->count[PTMC_C_VOLZ_ADJ] = 0;
    }
}

/* find_hist_index(hist_index, n_bin) src/Sampling.jl:53-62 ; returns 1-based index */
static int hist_index_f(double hist_index, int n_bin) {
    if (hist_index < 1) return 1;
    if (hist_index > n_bin) return n_bin + 2;
    return (int)floor(hist_index) + 1;
}
static int hist_index_i(int64_t hist_index, int n_bin) {
    if (hist_index < 1) return 1;
    if (hist_index > n_bin) return n_bin + 2;
    return (int)hist_index + 1;
}

/* sampling_step! src/Sampling.jl:330-371 for one ladder */
static void sampling_step(orc_handle* H, int64_t lad) {
    const ptmc_config* c = &H->cfg;
    const int nb = c->n_bin, nt = c->n_traj;
    double en_min = H->en_min[lad], en_max = H->en_max[lad];
    double delta_en = (en_max - en_min) / (nb - 1); /* :150 */
    for (int t = 0; t < nt; ++t) {
        orc_state* st = &H->st[lad * nt + t];
        orc_cfgside* s = st->c;
        if (c->ensemble == PTMC_ENS_NPT) {
            /* update_energy_tot :35-46 */
            double v = orc_get_volume(c->bc_kind, s->box[0], s->box[1]);
            double hval = s->en_tot + c->pressure * v;
            st->ham[0] += hval;
            st->ham[1] += hval * hval;
            /* update_histograms! :267-280 + find_hist_index :72-84 (Cubic), :98-113 (Rect), :115-130 (Rhombic) */
            double delta_v = (H->v_max[lad] - H->v_min[lad]) / nb; /* :183,:211 */
            double vol_h;
            if (c->bc_kind == PTMC_BC_CUBIC) vol_h = s->box[0] * s->box[0] * s->box[0];
            else if (c->bc_kind == PTMC_BC_RECTANGULAR) vol_h = s->box[0] * s->box[0] * s->box[1];
            else vol_h = s->box[0] * s->box[0] * (3 * pow(s->box[1], 0.5)) / 2; /* quirk 5, :121-122 */
            int64_t ie = (int64_t)floor((s->en_tot - en_min) / delta_en) + 1;
            int64_t iv = (int64_t)floor((vol_h - H->v_min[lad]) / delta_v) + 1;
            int he = hist_index_i(ie, nb), hv = hist_index_i(iv, nb);
            H->ev_hist[(((size_t)lad * nt + t) * (nb + 2) + (he - 1)) * (nb + 2) + (hv - 1)] += 1;
        } else {
            st->ham[0] += s->en_tot;                  /* update_energy_tot :29-34 */
            st->ham[1] += (s->en_tot * s->en_tot);
            double hi = (s->en_tot - en_min) / delta_en + 1; /* :63-66 */
            int hidx = hist_index_f(hi, nb);
            H->en_hist[((size_t)lad * nt + t) * (nb + 2) + (hidx - 1)] += 1;
        }
    }
}

/* exc_acceptance src/Exchange.jl:103-108 (Julia's min propagates NaN) */
static double exc_acceptance(double b1, double b2, double e1, double e2) {
    double x = exp((b1 - b2) * (e1 - e2));
    if (isnan(x)) return x;
    return x < 1.0 ? x : 1.0;
}

/* mc_cycle! src/MCRun.jl:192-253 for one ladder.  par != 0: thread over replicas like :176. */
/* phase (ptmc_run_phase): 0 whole cycle, 1 mc_step! only, 2 step-size adaptation + sampling only */
static int mc_cycle(orc_handle* H, int64_t lad, int64_t idx, int sample, int64_t tr_cycle, int par, int phase) {
    const ptmc_config* c = &H->cfg;
    const int nt = c->n_traj, ns = n_steps_of(c);
    int err = 0;
    /* mc_step! :169-182 */
#pragma omp parallel for schedule(static) if (par) reduction(min : err)
    for (int t = 0; t < (phase == 2 ? 0 : nt); ++t) {
        orc_state* st = &H->st[lad * nt + t];
        for (int k = 0; k < ns; ++k) {
            int32_t* ti = NULL; double* td = NULL; uint8_t* ta = NULL;
            if (tr_cycle >= 0) {
                size_t o = (((size_t)lad * H->trace_cap + tr_cycle) * nt + t) * ns + k;
                ti = H->tr_index + o; td = H->tr_de + o; ta = H->tr_acc + o;
            }
            int e = mc_move(H, st, ti, td, ta);
            if (e < err) err = e;
            if (e) break;
        }
    }
    if (err) return err;
    /* :203-205 exchange, driver-task RNG */
    orc_rng* lr = &H->lrng[lad];
    int32_t ex_gate = 0, ex_n = 0, ex_acc = 0;
    if (phase == 0 && rng_rand(lr) < 0.1) {
        ex_gate = 1;
        /* parallel_tempering_exchange! src/Exchange.jl:132-186 */
        int n_exc = (int)rng_range(lr, (uint64_t)(nt - 1));
        ex_n = n_exc;
        orc_state* a = &H->st[lad * nt + n_exc - 1];
        orc_state* b = &H->st[lad * nt + n_exc];
        a->count[PTMC_C_EXC_ATT] += 1; b->count[PTMC_C_EXC_ATT] += 1;
        double ea = a->c->en_tot, eb = b->c->en_tot;
        if (c->ensemble == PTMC_ENS_NPT) {
            ea = ea + c->pressure * orc_get_volume(c->bc_kind, a->c->box[0], a->c->box[1]);
            eb = eb + c->pressure * orc_get_volume(c->bc_kind, b->c->box[0], b->c->box[1]);
        }
        if (exc_acceptance(a->beta, b->beta, ea, eb) > rng_rand(lr)) {
            ex_acc = 1;
            a->count[PTMC_C_EXC_ACC] += 1; b->count[PTMC_C_EXC_ACC] += 1;
            orc_cfgside* t = a->c; a->c = b->c; b->c = t; /* exc_trajectories! :115-124 */
        }
    }
    if (lr->exhausted) return PTMC_ERR_REPLAY;
    if (tr_cycle >= 0) {
        int32_t* e = H->tr_exc + ((size_t)lad * H->trace_cap + tr_cycle) * 3;
        e[0] = ex_gate; e[1] = ex_n; e[2] = ex_acc;
    }
    if (phase == 1) return 0;
    /* :206-212 */
    if (idx % c->n_adjust == 0)
        for (int t = 0; t < nt; ++t) update_max_stepsize(c, &H->st[lad * nt + t]);
    /* :248-250 */
    if (sample && idx % c->mc_sample == 0) sampling_step(H, lad);
    return 0;
}

/* =========================================================================== exported handle API == */
static void* xcalloc(size_t n, size_t sz) { void* p = calloc(n ? n : 1, sz); if (!p) { perror("oracle calloc"); abort(); } return p; }

int orc_seed(orc_handle* H, uint64_t seed);
int orc_seed_job(orc_handle* H, uint64_t seed, int64_t first_ladder);
orc_handle* orc_create(const ptmc_config* cfg) {
    if (cfg->pot_kind == PTMC_POT_RUNNER && (!cfg->runner || cfg->bc_kind != PTMC_BC_SPHERICAL)) return NULL;
    if (cfg->pot_kind == PTMC_POT_ELJ_B && (!cfg->eljb_a || !cfg->eljb_b)) return NULL;
    if ((cfg->pot_kind == PTMC_POT_RUNNER2) != (cfg->ensemble == PTMC_ENS_NNVT)) return NULL;
    if (cfg->pot_kind == PTMC_POT_RUNNER2 && (!cfg->runner2 || cfg->bc_kind != PTMC_BC_SPHERICAL || cfg->runner2->n1 < 1 || cfg->runner2->n2 < 1 ||
                                              cfg->runner2->n1 + cfg->runner2->n2 != cfg->n_atoms || cfg->n_atom_swaps < 0)) return NULL;
    if (cfg->ensemble == PTMC_ENS_NPT && cfg->separated_volume && cfg->bc_kind != PTMC_BC_RHOMBIC && cfg->bc_kind != PTMC_BC_RECTANGULAR) return NULL;
    if (cfg->pot_kind == PTMC_POT_LUT && (!cfg->lut_table || cfg->lut_l_dist < 1 || cfg->lut_l_angle < 91 || !(cfg->lut_d_angle > 0) ||
                                          cfg->lut_l_angle < (int)rint(90.0 / cfg->lut_d_angle + 1.0))) return NULL;
    orc_handle* H = (orc_handle*)xcalloc(1, sizeof *H);
    H->cfg = *cfg;
    if (cfg->pot_kind == PTMC_POT_LUT) {
        const size_t nt_ = (size_t)cfg->lut_l_dist * cfg->lut_l_angle;
        H->lut = (double*)xcalloc(nt_, sizeof(double));
        memcpy(H->lut, cfg->lut_table, sizeof(double) * nt_);
        H->cfg.lut_table = H->lut;
    } else H->cfg.lut_table = NULL;
    if (cfg->pot_kind == PTMC_POT_ELJ_B) {
        H->eljb_a = (double*)xcalloc(cfg->n_coeff, sizeof(double)); H->eljb_b = (double*)xcalloc(cfg->n_coeff, sizeof(double));
        memcpy(H->eljb_a, cfg->eljb_a, sizeof(double) * cfg->n_coeff); memcpy(H->eljb_b, cfg->eljb_b, sizeof(double) * cfg->n_coeff);
        H->cfg.eljb_a = H->eljb_a; H->cfg.eljb_b = H->eljb_b;
    } else { H->cfg.eljb_a = NULL; H->cfg.eljb_b = NULL; }
    if (cfg->pot_kind == PTMC_POT_RUNNER) H->runner = orc_runner_new(cfg->runner);
    if (cfg->pot_kind == PTMC_POT_RUNNER2) H->runner2 = orc_runner2_new(cfg->runner2);
    H->cfg.runner = NULL; H->cfg.runner2 = NULL;
    const int n = cfg->n_atoms, nt = cfg->n_traj;
    H->betas = (double*)xcalloc(nt, sizeof(double));
    memcpy(H->betas, cfg->betas, sizeof(double) * nt);
    H->cfg.betas = H->betas;
    H->n_rep = cfg->n_ladders * nt;
    H->st = (orc_state*)xcalloc(H->n_rep, sizeof(orc_state));
    H->cs = (orc_cfgside*)xcalloc(H->n_rep, sizeof(orc_cfgside));
    H->lrng = (orc_rng*)xcalloc(cfg->n_ladders, sizeof(orc_rng));
    const int npt = cfg->ensemble == PTMC_ENS_NPT;
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_cfgside* s = &H->cs[r];
        s->pos = (double*)xcalloc(3 * n, sizeof(double));
        s->d2 = (double*)xcalloc((size_t)n * n, sizeof(double));
        if (npt) { s->trial_pos = (double*)xcalloc(3 * n, sizeof(double)); s->new_d2mat = (double*)xcalloc((size_t)n * n, sizeof(double)); }
        if (cfg->pot_kind == PTMC_POT_EAM) { s->comp = (double*)xcalloc(2 * n, sizeof(double)); s->new_comp = (double*)xcalloc(2 * n, sizeof(double)); }
        if (cfg->pot_kind == PTMC_POT_ELJ_B || cfg->pot_kind == PTMC_POT_LUT) s->tan = (double*)xcalloc((size_t)n * n, sizeof(double));
        if (cfg->pot_kind == PTMC_POT_RUNNER || cfg->pot_kind == PTMC_POT_RUNNER2) {
            const size_t ns = cfg->pot_kind == PTMC_POT_RUNNER ? (size_t)orc_runner_nsym(H->runner) : (size_t)orc_runner2_nsym(H->runner2);
            s->fmat = (double*)xcalloc((size_t)n * n, 8); s->new_f = (double*)xcalloc(n, 8);
            s->g = (double*)xcalloc(n * ns, 8); s->new_g = (double*)xcalloc(n * ns, 8);
            s->en_atom = (double*)xcalloc(n, 8); s->new_en_atom = (double*)xcalloc(n, 8);
        }
        orc_state* st = &H->st[r];
        st->c = s;
        st->beta = H->betas[r % nt];
        st->new_d2 = (double*)xcalloc(n, sizeof(double));
        st->new_tan = (double*)xcalloc(n, sizeof(double));
    }
    H->en_min = (double*)xcalloc(cfg->n_ladders, sizeof(double));
    H->en_max = (double*)xcalloc(cfg->n_ladders, sizeof(double));
    H->v_min = (double*)xcalloc(cfg->n_ladders, sizeof(double));
    H->v_max = (double*)xcalloc(cfg->n_ladders, sizeof(double));
    size_t nb2 = (size_t)cfg->n_bin + 2;
    if (npt) H->ev_hist = (double*)xcalloc((size_t)H->n_rep * nb2 * nb2, sizeof(double));
    else H->en_hist = (double*)xcalloc((size_t)H->n_rep * nb2, sizeof(double));
    H->n_threads = 1;
    if (cfg->rng_mode != PTMC_RNG_REPLAY) orc_seed_job(H, cfg->seed, 0);   /* like ptmc_create */
    return H;
}

void orc_destroy(orc_handle* H) {
    if (!H) return;
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_cfgside* s = &H->cs[r];
        free(s->pos); free(s->d2); free(s->trial_pos); free(s->new_d2mat); free(s->comp); free(s->new_comp);
        free(s->fmat); free(s->new_f); free(s->g); free(s->new_g); free(s->en_atom); free(s->new_en_atom);
        free(s->tan); free(H->st[r].new_d2); free(H->st[r].new_tan);
    }
    free(H->st); free(H->cs); free(H->lrng); free(H->betas);
    free(H->en_min); free(H->en_max); free(H->v_min); free(H->v_max); free(H->en_hist); free(H->ev_hist);
    free(H->tr_index); free(H->tr_de); free(H->tr_acc); free(H->tr_exc);
    orc_runner_free(H->runner);
    orc_runner2_free(H->runner2);
    free(H->eljb_a); free(H->eljb_b); free(H->lut);
    free(H);
}

void orc_set_threads(orc_handle* H, int n) { H->n_threads = n < 1 ? 1 : n; }

int orc_upload_states(orc_handle* H, const double* pos, const double* box, const double* max_displ) {
    const ptmc_config* c = &H->cfg;
    const int n = c->n_atoms, nt = c->n_traj;
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_state* st = &H->st[r];
        orc_cfgside* s = &H->cs[r];
        st->c = s;
        st->beta = H->betas[r % nt];
        memcpy(s->pos, pos + (size_t)r * 3 * n, sizeof(double) * 3 * n);
        s->box[0] = box[2 * r]; s->box[1] = box[2 * r + 1];
        for (int k = 0; k < 4; ++k) st->max_displ[k] = max_displ ? max_displ[4 * r + k] : 0.1;
        st->max_boxlength = c->bc_kind == PTMC_BC_SPHERICAL ? 30.0 : s->box[0] * 1.8; /* src/MCStates.jl:63-75 */
        st->max_boxheight = c->bc_kind == PTMC_BC_SPHERICAL ? 30.0 : (c->bc_kind == PTMC_BC_CUBIC ? s->box[0] * 1.8 : s->box[1] * 1.8); /* :76-87 */
        st->lh_ratio = st->max_boxlength / st->max_boxheight;
        st->ham[0] = st->ham[1] = 0.0;
        memset(st->count, 0, sizeof st->count);
        st->new_en = 0.0;
    }
#pragma omp parallel for schedule(dynamic) num_threads(H->n_threads)
    for (int64_t r = 0; r < H->n_rep; ++r) {
        init_cfgside(c, &H->cs[r]);
        if (c->pot_kind == PTMC_POT_RUNNER) init_cfgside_runner(H->runner, &H->cs[r], n);
        if (c->pot_kind == PTMC_POT_RUNNER2) init_cfgside_runner2(H->runner2, &H->cs[r], n);
    }
    return 0;
}

int orc_upload_counters(orc_handle* H, const int64_t* counters, const double* ham) {
    for (int64_t r = 0; r < H->n_rep; ++r) {
        if (counters) memcpy(H->st[r].count, counters + r * PTMC_NCOUNT, sizeof(int64_t) * PTMC_NCOUNT);
        if (ham) { H->st[r].ham[0] = ham[2 * r]; H->st[r].ham[1] = ham[2 * r + 1]; }
    }
    return 0;
}
int orc_reset_counters(orc_handle* H) {
    for (int64_t r = 0; r < H->n_rep; ++r) memset(H->st[r].count, 0, sizeof H->st[r].count);
    return 0;
}

/* words of generator state per slot (ptmc_rng_words) */
int orc_rng_words(const orc_handle* H) { return H->cfg.rng_mode == PTMC_RNG_BATCH8 ? 4 * PTMC_RNG_LANES : 4; }

int orc_set_rng_states(orc_handle* H, const uint64_t* slot_states, const uint64_t* ladder_states) {
    const int batch = H->cfg.rng_mode == PTMC_RNG_BATCH8;
    const int w = orc_rng_words(H);
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_rng* g = &H->st[r].rng;
        memset(g, 0, sizeof *g);
        g->batch_mode = batch;
        if (batch) memcpy(g->sub, slot_states + (size_t)w * r, sizeof g->sub);
        else memcpy(g->s, slot_states + 4 * r, 32);
    }
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l) { memset(&H->lrng[l], 0, sizeof(orc_rng)); memcpy(H->lrng[l].s, ladder_states + 4 * l, 32); }
    return 0;
}
int orc_get_rng_states(orc_handle* H, uint64_t* slot_states, uint64_t* ladder_states) {
    const int batch = H->cfg.rng_mode == PTMC_RNG_BATCH8;
    const int w = orc_rng_words(H);
    for (int64_t r = 0; r < H->n_rep; ++r) {
        if (batch) memcpy(slot_states + (size_t)w * r, H->st[r].rng.sub, sizeof H->st[r].rng.sub);
        else memcpy(slot_states + 4 * r, H->st[r].rng.s, 32);
    }
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l) memcpy(ladder_states + 4 * l, H->lrng[l].s, 32);
    return 0;
}
/* seeding rule of ptmc_seed (include/ptmc_b200.h) */
int orc_seed(orc_handle* H, uint64_t seed) {
    const int w = orc_rng_words(H) / 4;
    uint64_t* s = (uint64_t*)xcalloc((size_t)H->n_rep * 4 * w, 8);
    uint64_t* l = (uint64_t*)xcalloc((size_t)H->cfg.n_ladders * 4, 8);
    for (int64_t r = 0; r < H->n_rep * w; ++r) orc_julia_seed(seed + (uint64_t)r, s + 4 * r);
    for (int64_t k = 0; k < H->cfg.n_ladders; ++k) orc_julia_seed(seed + (uint64_t)(H->n_rep * w) + (uint64_t)k, l + 4 * k);
    int rc = orc_set_rng_states(H, s, l);
    free(s); free(l);
    return rc;
}
/* seeding rule of ptmc_seed_job (include/ptmc_b200.h) */
int orc_seed_job(orc_handle* H, uint64_t seed, int64_t first_ladder) {
    const uint64_t W = (uint64_t)orc_rng_words(H) / 4, T = (uint64_t)H->cfg.n_traj, L = (uint64_t)H->cfg.n_ladders;
    uint64_t* s = (uint64_t*)xcalloc((size_t)(L * T * W * 4), 8);
    uint64_t* l = (uint64_t*)xcalloc((size_t)(L * 4), 8);
    for (uint64_t a = 0; a < L; ++a) {
        const uint64_t gl = (uint64_t)first_ladder + a;
        for (uint64_t t = 0; t < T; ++t)
            for (uint64_t k = 0; k < W; ++k) orc_julia_seed128(seed, (gl * T + t) * W + k + 1, s + ((a * T + t) * W + k) * 4);
        orc_julia_seed128(seed, (1ull << 63) | gl, l + 4 * a);
    }
    int rc = orc_set_rng_states(H, s, l);
    free(s); free(l);
    return rc;
}
/* the caller keeps the draw buffers alive */
int orc_set_replay(orc_handle* H, const uint64_t* slot_draws, int64_t slot_len, const uint64_t* ladder_draws, int64_t ladder_len) {
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_rng* g = &H->st[r].rng; memset(g, 0, sizeof *g);
        g->replay = slot_draws + r * slot_len; g->len = slot_len;
    }
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l) {
        orc_rng* g = &H->lrng[l]; memset(g, 0, sizeof *g);
        g->replay = ladder_draws + l * ladder_len; g->len = ladder_len;
    }
    return 0;
}

int orc_set_hist_bounds(orc_handle* H, const double* en_min, const double* en_max, const double* v_min, const double* v_max, int per_ladder) {
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l) {
        int64_t k = per_ladder ? l : 0;
        H->en_min[l] = en_min[k]; H->en_max[l] = en_max[k];
        H->v_min[l] = v_min ? v_min[k] : 0.0; H->v_max[l] = v_max ? v_max[k] : 0.0;
    }
    size_t nb2 = (size_t)H->cfg.n_bin + 2;
    if (H->en_hist) memset(H->en_hist, 0, sizeof(double) * H->n_rep * nb2);
    if (H->ev_hist) memset(H->ev_hist, 0, sizeof(double) * H->n_rep * nb2 * nb2);
    H->hist_ready = 1;
    return 0;
}

int orc_trace_enable(orc_handle* H, int64_t capacity_cycles) {
    free(H->tr_index); free(H->tr_de); free(H->tr_acc); free(H->tr_exc);
    H->tr_index = NULL; H->tr_de = NULL; H->tr_acc = NULL; H->tr_exc = NULL;
    H->trace_cap = capacity_cycles; H->trace_n = 0;
    if (capacity_cycles > 0) {
        size_t m = (size_t)H->cfg.n_ladders * capacity_cycles * H->cfg.n_traj * n_steps_of(&H->cfg);
        H->tr_index = (int32_t*)xcalloc(m, sizeof(int32_t));
        H->tr_de = (double*)xcalloc(m, sizeof(double));
        H->tr_acc = (uint8_t*)xcalloc(m, 1);
        H->tr_exc = (int32_t*)xcalloc((size_t)H->cfg.n_ladders * capacity_cycles * 3, sizeof(int32_t));
    }
    return 0;
}
/* output arrays are [n_ladders][n_cycles_recorded][...] (compacted) */
int orc_trace_download(orc_handle* H, int32_t* move_index, double* delta_en, uint8_t* accepted, int32_t* exchange, int64_t* n_out) {
    const size_t per = (size_t)H->cfg.n_traj * n_steps_of(&H->cfg);
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l) {
        size_t src = (size_t)l * H->trace_cap * per, dst = (size_t)l * H->trace_n * per, cnt = (size_t)H->trace_n * per;
        if (move_index) memcpy(move_index + dst, H->tr_index + src, cnt * sizeof(int32_t));
        if (delta_en) memcpy(delta_en + dst, H->tr_de + src, cnt * sizeof(double));
        if (accepted) memcpy(accepted + dst, H->tr_acc + src, cnt);
        if (exchange) memcpy(exchange + (size_t)l * H->trace_n * 3, H->tr_exc + (size_t)l * H->trace_cap * 3, (size_t)H->trace_n * 3 * sizeof(int32_t));
    }
    if (n_out) *n_out = H->trace_n;
    return 0;
}

int orc_run_phase(orc_handle* H, int64_t first_idx, int64_t n_cycles, int sample, int phase);
int orc_run_cycles(orc_handle* H, int64_t first_idx, int64_t n_cycles, int sample) { return orc_run_phase(H, first_idx, n_cycles, sample, 0); }
int orc_run_phase(orc_handle* H, int64_t first_idx, int64_t n_cycles, int sample, int phase) {
    if (sample && !H->hist_ready) return PTMC_ERR_STATE;
    if (phase < 0 || phase > 2 || (phase && H->trace_cap > 0)) return PTMC_ERR_ARG;
    const int64_t nl = H->cfg.n_ladders;
    int err = 0;
    int64_t tr0 = H->trace_n;
    int tracing = H->trace_cap > 0;
    if (tracing && tr0 + n_cycles > H->trace_cap) return PTMC_ERR_ARG;
    if (nl >= H->n_threads && H->n_threads > 1) {
        /* many ladders: one task per ladder, no barriers (most favourable CPU arrangement) */
#pragma omp parallel for schedule(dynamic) num_threads(H->n_threads) reduction(min : err)
        for (int64_t l = 0; l < nl; ++l) {
            for (int64_t k = 0; k < n_cycles; ++k) {
                int e = mc_cycle(H, l, first_idx + k, sample, tracing ? tr0 + k : -1, 0, phase);
                if (e < err) err = e;
                if (e) break;
            }
        }
    } else {
#ifdef _OPENMP
        omp_set_num_threads(H->n_threads);
#endif
        for (int64_t k = 0; k < n_cycles && !err; ++k)
            for (int64_t l = 0; l < nl && !err; ++l)
                err = mc_cycle(H, l, first_idx + k, sample, tracing ? tr0 + k : -1, H->n_threads > 1, phase);
    }
    if (tracing) H->trace_n += n_cycles;
    H->error = err;
    return err;
}

int orc_download_states(orc_handle* H, double* pos, double* box, double* en_tot, double* max_displ, int64_t* counters, double* ham) {
    const int n = H->cfg.n_atoms;
    for (int64_t r = 0; r < H->n_rep; ++r) {
        orc_state* st = &H->st[r];
        if (pos) memcpy(pos + (size_t)r * 3 * n, st->c->pos, sizeof(double) * 3 * n);
        if (box) { box[2 * r] = st->c->box[0]; box[2 * r + 1] = st->c->box[1]; }
        if (en_tot) en_tot[r] = st->c->en_tot;
        if (max_displ) memcpy(max_displ + 4 * r, st->max_displ, sizeof(double) * 4);
        if (counters) memcpy(counters + r * PTMC_NCOUNT, st->count, sizeof(int64_t) * PTMC_NCOUNT);
        if (ham) { ham[2 * r] = st->ham[0]; ham[2 * r + 1] = st->ham[1]; }
    }
    return 0;
}
/* ptmc_config_doubles / ptmc_export_config / ptmc_import_config: what exc_trajectories! moves (SphericalBC handles) */
int64_t orc_config_doubles(const orc_handle* H) {
    const int64_t n = H->cfg.n_atoms;
    return 3 * n + 3 + (H->cfg.pot_kind == PTMC_POT_EAM ? 2 * n : 0);
}
int orc_export_config(orc_handle* H, int64_t ladder, int32_t slot, double* buf) {
    if (H->cfg.bc_kind != PTMC_BC_SPHERICAL || H->cfg.pot_kind == PTMC_POT_RUNNER || H->cfg.pot_kind == PTMC_POT_RUNNER2) return PTMC_ERR_ARG;
    const int n = H->cfg.n_atoms;
    const orc_cfgside* s = H->st[ladder * H->cfg.n_traj + slot].c;
    memcpy(buf, s->pos, sizeof(double) * 3 * n);
    buf[3 * n] = s->box[0]; buf[3 * n + 1] = s->box[1]; buf[3 * n + 2] = s->en_tot;
    if (H->cfg.pot_kind == PTMC_POT_EAM) memcpy(buf + 3 * n + 3, s->comp, sizeof(double) * 2 * n);
    return 0;
}
int orc_import_config(orc_handle* H, int64_t ladder, int32_t slot, const double* buf) {
    if (H->cfg.bc_kind != PTMC_BC_SPHERICAL || H->cfg.pot_kind == PTMC_POT_RUNNER || H->cfg.pot_kind == PTMC_POT_RUNNER2) return PTMC_ERR_ARG;
    const ptmc_config* c = &H->cfg;
    const int n = c->n_atoms;
    orc_cfgside* s = H->st[ladder * c->n_traj + slot].c;
    memcpy(s->pos, buf, sizeof(double) * 3 * n);
    s->box[0] = buf[3 * n]; s->box[1] = buf[3 * n + 1];
    /* derived matrices are functions of the positions alone in free space (distance2 and get_tan are symmetric bit for bit) */
    d2_matrix(c, s->pos, s->box, s->d2, n);
    if (c->pot_kind == PTMC_POT_ELJ_B || c->pot_kind == PTMC_POT_LUT) {
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j)
                s->tan[(size_t)i * n + j] = s->tan[(size_t)j * n + i] = orc_get_tan(c->bc_kind, s->pos + 3 * i, s->pos + 3 * j, s->box[0], s->box[1]);
        for (int i = 0; i < n; ++i) s->tan[(size_t)i * n + i] = 0;
    }
    s->en_tot = buf[3 * n + 2];
    if (c->pot_kind == PTMC_POT_EAM) memcpy(s->comp, buf + 3 * n + 3, sizeof(double) * 2 * n);
    return 0;
}
int orc_download_hist(orc_handle* H, double* en_hist, double* ev_hist, int reduce_ladders) {
    const size_t nb2 = (size_t)H->cfg.n_bin + 2, nt = H->cfg.n_traj;
    const double* src = H->en_hist ? H->en_hist : H->ev_hist;
    double* dst = H->en_hist ? en_hist : ev_hist;
    if (!dst) return 0;
    const size_t per = nt * (H->en_hist ? nb2 : nb2 * nb2);
    if (!reduce_ladders) { memcpy(dst, src, sizeof(double) * per * H->cfg.n_ladders); return 0; }
    memset(dst, 0, sizeof(double) * per);
    for (int64_t l = 0; l < H->cfg.n_ladders; ++l)
        for (size_t k = 0; k < per; ++k) dst[k] += src[(size_t)l * per + k];
    return 0;
}
/* dist2_mat of one replica (for the Julia-side MCState refill and for tests) */
int orc_get_dist2_mat(orc_handle* H, int64_t replica, double* out) {
    memcpy(out, H->st[replica].c->d2, sizeof(double) * (size_t)H->cfg.n_atoms * H->cfg.n_atoms);
    return 0;
}
/* single-move entry used by the seeded KAT (test/test_states.jl:27-38): performs generate_move! +
 * get_energy! for a *given* atom index with the replica's slot stream and reports new_en - en_tot and
 * the trial position without committing. */
int orc_probe_atom_move(orc_handle* H, int64_t replica, int index1, double* trial_out, double* delta_en) {
    const ptmc_config* c = &H->cfg;
    orc_state* st = &H->st[replica];
    orc_cfgside* s = st->c;
    const int n = c->n_atoms, i = index1 - 1;
    double trial[3];
    displace(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, trial);
    if (c->bc_kind == PTMC_BC_SPHERICAL)
        while (trial[0] * trial[0] + trial[1] * trial[1] + trial[2] * trial[2] > s->box[0])
            displace(c, s->pos + 3 * i, st->max_displ[0], s->box, &st->rng, trial);
    for (int j = 0; j < n; ++j) st->new_d2[j] = orc_distance2(c->bc_kind, trial, s->pos + 3 * j, s->box[0], s->box[1]);
    st->new_d2[i] = 0.0;
    const int npt = c->ensemble == PTMC_ENS_NPT;
    double delta = dimer_energy_atom(c, i, st->new_d2, n, npt, s->r_cut) - dimer_energy_atom(c, i, s->d2 + (size_t)i * n, n, npt, s->r_cut);
    double new_en = delta + s->en_tot;
    memcpy(trial_out, trial, sizeof trial);
    *delta_en = new_en - s->en_tot;
    return 0;
}

/* RuNNer: g_matrix [atom][n_sym] and per-atom energies of one replica (tests) */
int orc_get_runner_state(orc_handle* H, int64_t replica, double* g, double* en_atom) {
    if (!H->runner && !H->runner2) return PTMC_ERR_ARG;
    const size_t ns = H->runner ? (size_t)orc_runner_nsym(H->runner) : (size_t)orc_runner2_nsym(H->runner2), n = (size_t)H->cfg.n_atoms;
    if (g) memcpy(g, H->st[replica].c->g, sizeof(double) * n * ns);
    if (en_atom) memcpy(en_atom, H->st[replica].c->en_atom, sizeof(double) * n);
    return 0;
}
