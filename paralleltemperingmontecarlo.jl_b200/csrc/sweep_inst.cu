// sweep_inst.cu — compiled once per lanes-per-replica value (-DPTMC_G=4|8|16|32); exports the table of
// sweep_kernel instantiations for that G.
#include "sweep_kernel.cuh"

#ifndef PTMC_G
#error "compile with -DPTMC_G=<2|4|8|16|32>"
#endif
#define PTMC_CAT2(a, b) a##b
#define PTMC_CAT(a, b) PTMC_CAT2(a, b)

namespace ptmc {

#if PTMC_G == 2
// Two lanes per replica, 16 replicas per warp: free-space ELJEven ladders whose n_traj x 2 threads fill whole warps (Ne13 at 32
// temperatures: BASELINE C1) and whose other N - 1 atoms split evenly over the two lanes (ROT: no parking, no padding).  The per-move
// work every lane repeats is paid once per 16 moves.  Only this family is instantiated for G = 2.  A 64-thread ladder leaves registers
// to spare; measured on C1 (ms per step of 1.48e9 moves): caps of 80 / 96 / 112 / 128 / 152 registers -> 80.2 / 71.7 / 67.6 / 66.2 / 70.7
// (the 4-lane kernel: 79.1).
#ifndef PTMC_G2_MAXREG
#define PTMC_G2_MAXREG 128
#endif
template <int POT, bool DBG, int NITER, int NC, bool FULLW, bool ROT>
__global__ void __maxnreg__(PTMC_G2_MAXREG)
sweep_kernel_g2(const DevParams P, const i64 first_idx, const i64 n_cycles, const int sample) {
    sweep_body<2, POT, GEOM_SPHERE, DBG, NITER, NC, FULLW, ROT>(P, first_idx, n_cycles, sample);
}
sweep_fn sweep_table_g2(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads) {
    (void)threads;
    if (pot != PTMC_POT_ELJ_EVEN || geom != GEOM_SPHERE) return nullptr;
    if (dbg) return (sweep_fn)sweep_kernel_g2<PTMC_POT_ELJ_EVEN, true, 0, 0, false, false>;
    if (n_coeff == 6 && rot && fullw && niter >= 2) {
        switch (niter - 1) {
#define PTMC_G2R(K) case K: return (sweep_fn)sweep_kernel_g2<PTMC_POT_ELJ_EVEN, false, K, 6, true, true>;
            PTMC_G2R(2) PTMC_G2R(3) PTMC_G2R(4) PTMC_G2R(5) PTMC_G2R(6) PTMC_G2R(7) PTMC_G2R(8) PTMC_G2R(9) PTMC_G2R(10) PTMC_G2R(12) PTMC_G2R(14) PTMC_G2R(16)
#undef PTMC_G2R
            default: break;
        }
    }
    return (sweep_fn)sweep_kernel_g2<PTMC_POT_ELJ_EVEN, false, 0, 0, false, false>;
}
#else

// Generic kernels (runtime pair-loop trip count and coefficient count), one per potential / geometry / debug.
static sweep_fn generic_kernel(int pot, int geom, int dbg) {
    constexpr int G = PTMC_G;
#define PTMC_PICK(POT, GEOM)                                                                   \
    if (pot == POT && geom == GEOM)                                                            \
        return dbg ? (sweep_fn)sweep_kernel<G, POT, GEOM, true, 0, 0> : (sweep_fn)sweep_kernel<G, POT, GEOM, false, 0, 0>;
    PTMC_PICK(PTMC_POT_ELJ_EVEN, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_ELJ_EVEN, GEOM_PBC)
    PTMC_PICK(PTMC_POT_ELJ_EVEN, GEOM_RHOMB)
    PTMC_PICK(PTMC_POT_ELJ, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_ELJ, GEOM_PBC)
    PTMC_PICK(PTMC_POT_ELJ, GEOM_RHOMB)
    PTMC_PICK(PTMC_POT_ELJ_B, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_ELJ_B, GEOM_PBC)
    PTMC_PICK(PTMC_POT_ELJ_B, GEOM_RHOMB)
    PTMC_PICK(PTMC_POT_LUT, GEOM_SPHERE)
    PTMC_PICK(PTMC_POT_LUT, GEOM_PBC)
    PTMC_PICK(PTMC_POT_LUT, GEOM_RHOMB)
    if (pot == PTMC_POT_EAM && geom == GEOM_SPHERE) return dbg ? (sweep_fn)sweep_kernel_eam<G, true, 0> : (sweep_fn)sweep_kernel_eam<G, false, 0>;
#undef PTMC_PICK
    return nullptr;
}

// Specialised production kernels of the headline family (ELJPotentialEven with 6 coefficients, SphericalBC,
// BATCH8 streams): pair loop fully unrolled for ceil(n_atoms / G) = 1..8 (1..16 for G = 4).
// (N - 1) % G == 0 and a fully active warp layout: rotated neighbour order, niter - 1 pair iterations (sweep_kernel.cuh, ROT)
template <int POT>
static sweep_fn rotated_kernel(int niter_rot) {
    constexpr int G = PTMC_G;
    switch (niter_rot) {
#define PTMC_SPR(K) case K: return (sweep_fn)sweep_kernel<G, POT, GEOM_SPHERE, false, K, 6, true, true>;
        PTMC_SPR(1) PTMC_SPR(2) PTMC_SPR(3) PTMC_SPR(4) PTMC_SPR(5) PTMC_SPR(6) PTMC_SPR(7) PTMC_SPR(8)
#undef PTMC_SPR
        default: return nullptr;
    }
}

static sweep_fn special_kernel(int niter, int fullw) {
    constexpr int G = PTMC_G;
    switch (niter) {
#define PTMC_SP(K) case K: return fullw ? (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_SPHERE, false, K, 6, true> \
                                        : (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_SPHERE, false, K, 6, false>;
        PTMC_SP(1) PTMC_SP(2) PTMC_SP(3) PTMC_SP(4) PTMC_SP(5) PTMC_SP(6) PTMC_SP(7) PTMC_SP(8)
#if PTMC_G == 4
        PTMC_SP(9) PTMC_SP(10) PTMC_SP(11) PTMC_SP(12) PTMC_SP(13) PTMC_SP(14) PTMC_SP(15) PTMC_SP(16)
#endif
#undef PTMC_SP
        default: return nullptr;
    }
}

// pot: ptmc_pot_kind (ELJ, ELJ_EVEN, EAM); geom: Geom; dbg: sequential-stream / replay / trace build;
// niter = n_pad / G; n_coeff: number of ELJ coefficients; fullw: n_traj * G is a multiple of 32 (no partly active warp)
sweep_fn PTMC_CAT(sweep_table_g, PTMC_G)(int pot, int geom, int dbg, int niter, int n_coeff, int fullw, int rot, int threads) {
    constexpr int G = PTMC_G;
    (void)threads;
#if PTMC_G == 4
    // ELJEven, six coefficients, sphere, production streams, ladders of at most 128 threads: the 96-register build (sweep_kernel_g4s)
    if (!dbg && pot == PTMC_POT_ELJ_EVEN && geom == GEOM_SPHERE && n_coeff == 6 && threads <= 128) {
        if (rot && fullw && niter >= 2) {
            switch (niter - 1) {
#define PTMC_G4R(K) case K: return (sweep_fn)sweep_kernel_g4s<PTMC_POT_ELJ_EVEN, false, K, 6, true, true>;
                PTMC_G4R(1) PTMC_G4R(2) PTMC_G4R(3) PTMC_G4R(4) PTMC_G4R(5) PTMC_G4R(6) PTMC_G4R(7) PTMC_G4R(8)
#undef PTMC_G4R
                default: break;
            }
        }
        switch (niter) {
#define PTMC_G4N(K) case K: return fullw ? (sweep_fn)sweep_kernel_g4s<PTMC_POT_ELJ_EVEN, false, K, 6, true, false> \
                                         : (sweep_fn)sweep_kernel_g4s<PTMC_POT_ELJ_EVEN, false, K, 6, false, false>;
            PTMC_G4N(1) PTMC_G4N(2) PTMC_G4N(3) PTMC_G4N(4) PTMC_G4N(5) PTMC_G4N(6) PTMC_G4N(7) PTMC_G4N(8)
            PTMC_G4N(9) PTMC_G4N(10) PTMC_G4N(11) PTMC_G4N(12) PTMC_G4N(13) PTMC_G4N(14) PTMC_G4N(15) PTMC_G4N(16)
#undef PTMC_G4N
            default: break;
        }
    }
#endif
    if (rot && fullw && !dbg && geom == GEOM_SPHERE && n_coeff == 6 && niter >= 2) {
        sweep_fn f = pot == PTMC_POT_ELJ_EVEN ? rotated_kernel<PTMC_POT_ELJ_EVEN>(niter - 1) : pot == PTMC_POT_ELJ_B ? rotated_kernel<PTMC_POT_ELJ_B>(niter - 1) : nullptr;
        if (f) return f;
    }
    if (!dbg && pot == PTMC_POT_ELJ_EVEN && geom == GEOM_SPHERE && n_coeff == 6) {
        sweep_fn f = special_kernel(niter, fullw);
        if (f) return f;
    }
    // ELJPotentialB with six coefficients (scripts/ne13_B.jl, ne38_B.jl, ne51_B.jl, ne55.jl): unrolled pair loop and Horner chain
    if (!dbg && pot == PTMC_POT_ELJ_B && geom == GEOM_SPHERE && n_coeff == 6) {
        switch (niter) {
#define PTMC_SPB6(K) case K: return fullw ? (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_B, GEOM_SPHERE, false, K, 6, true> \
                                          : (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_B, GEOM_SPHERE, false, K, 6, false>;
            PTMC_SPB6(1) PTMC_SPB6(2) PTMC_SPB6(3) PTMC_SPB6(4) PTMC_SPB6(5) PTMC_SPB6(6) PTMC_SPB6(7) PTMC_SPB6(8)
#undef PTMC_SPB6
            default: break;
        }
    }
    // EmbeddedAtomPotential: pair loop unrolled, trial components in registers
    if (pot == PTMC_POT_EAM && geom == GEOM_SPHERE && eam_unrolled(niter, dbg != 0)) {
        switch (niter) {
#define PTMC_SPE(K) case K: return fullw ? (sweep_fn)sweep_kernel_eam<G, false, K, true> : (sweep_fn)sweep_kernel_eam<G, false, K, false>;
            PTMC_SPE(1) PTMC_SPE(2) PTMC_SPE(3) PTMC_SPE(4) PTMC_SPE(5) PTMC_SPE(6) PTMC_SPE(7) PTMC_SPE(8)
#undef PTMC_SPE
            default: break;
        }
    }
    // periodic ELJEven with the six-coefficient form: Horner chain unrolled, coefficients in uniform registers; for the wide
    // lane groups that periodic systems use, the thawed pair loop is unrolled as well
#if PTMC_G == 16
    if (!dbg && pot == PTMC_POT_ELJ_EVEN && n_coeff == 6 && geom == GEOM_RHOMB && niter == 14 && threads <= 256)
        return (sweep_fn)sweep_kernel_p16<GEOM_RHOMB, 14>;
#endif
#if PTMC_G >= 16
    if (!dbg && pot == PTMC_POT_ELJ_EVEN && n_coeff == 6 && (geom == GEOM_PBC || geom == GEOM_RHOMB)) {
        switch (niter) {
#define PTMC_SPB(K) case K: return geom == GEOM_PBC ? (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_PBC, false, K, 6> \
                                                    : (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_RHOMB, false, K, 6>;
            PTMC_SPB(1) PTMC_SPB(2) PTMC_SPB(3) PTMC_SPB(4) PTMC_SPB(5) PTMC_SPB(6) PTMC_SPB(7) PTMC_SPB(8)
#undef PTMC_SPB
            default: break;
        }
    }
#endif
    // periodic ELJPotentialB, six coefficients (the scripts/pbc*.jl family at one warp per replica): the same unrolling
#if PTMC_G == 32
    if (!dbg && pot == PTMC_POT_ELJ_B && n_coeff == 6 && (geom == GEOM_PBC || geom == GEOM_RHOMB)) {
        switch (niter) {
#define PTMC_SPBB(K) case K: return geom == GEOM_PBC ? (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_B, GEOM_PBC, false, K, 6> \
                                                     : (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_B, GEOM_RHOMB, false, K, 6>;
            PTMC_SPBB(1) PTMC_SPBB(2) PTMC_SPBB(3) PTMC_SPBB(4) PTMC_SPBB(5) PTMC_SPBB(6) PTMC_SPBB(7) PTMC_SPBB(8)
#undef PTMC_SPBB
            default: break;
        }
    }
#endif
    if (!dbg && pot == PTMC_POT_ELJ_EVEN && n_coeff == 6) {
        if (geom == GEOM_PBC) return (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_PBC, false, 0, 6>;
        if (geom == GEOM_RHOMB) return (sweep_fn)sweep_kernel<G, PTMC_POT_ELJ_EVEN, GEOM_RHOMB, false, 0, 6>;
    }
    return generic_kernel(pot, geom, dbg);
}

#endif   // PTMC_G != 2

}  // namespace ptmc
