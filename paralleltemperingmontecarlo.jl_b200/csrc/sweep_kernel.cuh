// sweep_kernel.cuh — the parallel-tempering sweep: n_cycles x mc_cycle! (src/MCRun.jl:192-253) for one
// ladder per CTA.  Instantiated per (lanes-per-replica G, potential, geometry, debug, unrolling) in sweep_inst.cu.
//
//   mc_step!  (src/MCRun.jl:169-182)  -> every group runs n_steps x mc_move! on its slot, no CTA sync
//   exchange  (src/MCRun.jl:203-205, src/Exchange.jl:132-186) -> lane 0 of the lower slot between two mbarrier phases of the two warps involved
//   update_max_stepsize! (src/Exchange.jl:198-263), sampling_step! (src/Sampling.jl:330-371) -> per group
#pragma once
#include <type_traits>
#include "ptmc_device.cuh"

namespace ptmc {

// Launch bounds per lanes-per-replica value: n_traj * G (rounded up to a warp) threads per ladder.
// G <= 16: 384 threads, two CTAs per SM (80 registers); G = 32: 512 threads (128 registers).
#ifndef PTMC_MAX_THREADS_SMALL
#define PTMC_MAX_THREADS_SMALL 384
#endif
#ifndef PTMC_MIN_BLOCKS_SMALL
#define PTMC_MIN_BLOCKS_SMALL 2
#endif
__host__ __device__ constexpr int sweep_max_threads_c(int G) { return G == 32 ? 512 : PTMC_MAX_THREADS_SMALL; }
__host__ __device__ constexpr int sweep_min_blocks_c(int G) { return G == 32 ? 1 : PTMC_MIN_BLOCKS_SMALL; }
enum Geom { GEOM_SPHERE = 0, GEOM_PBC = 1, GEOM_RHOMB = 2 };   // SphericalBC+NVT | cubic / rectangular box | rhombic box (NVT or NPT at run time)

// shared-memory layout of one CTA (doubles unless noted)
struct SmemLayout {
    int cfg_stride;     // doubles per configuration buffer
    int off_scal;       // per config: en_tot, L, h, r_cut  (4 doubles)
    int off_perm;       // int cfg_of_slot[T]
    int off_exc;        // i64 exc counters [T][2]
    int off_ham;        // ham [T][2] (slot side)
    int off_lrng;       // ladder stream, one private copy per warp: 4 x u64 state + cursor + pad
    int off_moved;      // periodic: int moved[T][words + 1]
    int off_sep;        // periodic: per slot {max_displ[2], max_displ[3], max_boxheight, int n_acc_xy, n_adj_xy, n_acc_z, n_adj_z} (6 doubles)
    int off_tan;        // periodic tangent-dependent potentials: double tan_rho[T], int fresh[T][words] (stale-tangent state, see sep_trial_energy)
    int off_psum;       // periodic NPT with the tie table, ELJPotentialEven: power sums per configuration, double [T][6][lanes] + int valid[T]
    int off_tab;        // EAM: the mantissa table of eam_invrexp (PTMC_EAM_KTAB_DOUBLES, 16-byte aligned)
    size_t bytes;
};
// Power sums (uniform volume moves of ELJPotentialEven with 6 coefficients, periodic NPT with the tie table):
// E = sum_k c_k S_k, S_k = sum over the pairs inside the cut-off of (1 / d2)^(3 + k).  A uniform scaling by s multiplies every
// distance and the cut-off alike, so the trial energy of volume_change_uniform (src/MCMoves.jl:277-287 + dimer_energy_config on
// dist2_mat * s^2, src/MCRun.jl:76-85) is sum_k c_k s^-(6 + 2k) S_k: O(1) instead of N (N - 1) / 2 pair energies, once the S_k are
// carried along by the atom moves.
#define PTMC_PSUM_K 6
// EAM kernels with an unrolled pair loop keep the trial components in registers (eam_scratch = false).
#define PTMC_EAM_MAX_NITER 8
#ifndef PTMC_EAM_UNROLL
#define PTMC_EAM_UNROLL 1
#endif
__host__ __device__ constexpr bool eam_unrolled(int niter, bool dbg) { return PTMC_EAM_UNROLL && !dbg && niter >= 1 && niter <= PTMC_EAM_MAX_NITER; }
__host__ __device__ inline SmemLayout smem_layout(int n_traj, int n_pad, int pot_kind, int n_warps, bool pbc, bool eam_scratch, bool newvec, int lanes,
                                                  int cfg_mod = -1) {
    SmemLayout s;
    const bool dirp = pot_kind == PTMC_POT_ELJ_B || pot_kind == PTMC_POT_LUT;
    // x,y,z (+ phi,rho[,new_phi,new_rho] | + new_dist2_vec where the N x N matrix is kept: periodic without the tie table)
    int arrays = 3 + (pot_kind == PTMC_POT_EAM ? (eam_scratch ? 4 : 2) : 0) + ((pbc && newvec) ? 1 : 0);
    s.cfg_stride = arrays * n_pad;
    // The 32 / G replicas of a warp read the same atom slots of DIFFERENT configuration buffers: a buffer stride of G doubles (mod 16)
    // lets their 8 G-byte runs tile the 32 banks instead of landing on one another (G = 4: 168 doubles = 8 (mod 16) is a 2-way conflict
    // on top of the two wavefronts a 256-byte request needs anyway)
    // (cfg_mod: the residue the host asks for instead - 8 for the kernels whose lanes own a BLOCK of consecutive atoms and read them
    // two at a time with LDS.128: lanes 112 bytes apart, the second replica of a quarter warp half the banks further)
    const int mod = cfg_mod >= 0 ? cfg_mod : lanes;
    if (lanes <= 8) while ((s.cfg_stride & 15) != (mod & 15)) s.cfg_stride += 1;
    s.off_scal = n_traj * s.cfg_stride;
    int off = s.off_scal + 4 * n_traj;
    s.off_exc = off; off += 2 * n_traj;
    s.off_ham = off; off += 2 * n_traj;
    s.off_lrng = off; off += 6 * n_warps;
    s.off_perm = off; off += (n_traj + 1) / 2 + 1;
    s.off_moved = off; off += pbc ? (n_traj * ((n_pad + 31) / 32 + 1) + 1) / 2 + 1 : 0;
    s.off_sep = off; off += pbc ? 6 * n_traj : 0;
    s.off_tan = off; off += (pbc && dirp) ? n_traj + (n_traj * ((n_pad + 31) / 32) + 1) / 2 : 0;
    s.off_psum = off; off += (pbc && !newvec && pot_kind == PTMC_POT_ELJ_EVEN) ? n_traj * PTMC_PSUM_K * lanes + (n_traj + 1) / 2 : 0;
    off += off & 1;
    s.off_tab = off; off += pot_kind == PTMC_POT_EAM ? PTMC_EAM_KTAB_DOUBLES : 0;
    s.bytes = (size_t)off * sizeof(double);
    return s;
}

// Barrier of the two warps that hold the slots of one exchange attempt: an mbarrier in shared memory (64 arrivals), so no
// hardware named barrier is reserved (a data-dependent `bar.sync id` makes every CTA claim all 16 and caps residency at 4 CTAs / SM).
__device__ __forceinline__ void pair_barrier_init(u64* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void pair_barrier(u64* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// Separated volume moves (src/MCMoves.jl:300-343), separate loops so that the uniform-move loops stay as they are.
// Pair energy of the trial configuration (positions scaled by scale_xy / scale_z, box tL x tH) below the new cut-off:
// get_distance2_mat!(new_dist2_mat, trial_config) + dimer_energy_config; exact distance2 arithmetic while the configuration
// still holds never-moved lattice atoms, the reciprocal-multiply minimum image afterwards.  Returns this lane's partial sum.
//
// Tangent-dependent potentials: the reference keeps tan_mat, and its new_tan_mat is the SAME array (set_variables
// src/EnergyEvaluation.jl:1790,1798), so an xy / z trial leaves the TRIAL configuration's tangents in the state whether or not it is
// accepted (src/MCMoves.jl:333-341); later atom moves refresh row and column of the moved atom, the uniform trial reads the matrix.
// A trial tangent is rho = scale_xy / scale_z times the true one, so the whole matrix is carried as one factor per configuration
// plus one bit per atom ("moved since the last xy / z trial"): tan_mat[i, j] = tan(i, j) * (fresh_i | fresh_j ? 1 : rho), equal to
// the stored value up to rounding.  No N x N matrix in HBM, no scattered column writes (profiles/r01_experiments.md).
__device__ __forceinline__ bool tan_fresh(const int* fr, int j) { return (fr[j >> 5] >> (j & 31)) & 1; }
template <int G, int POT, int NC, bool RHOMBIC>
__device__ __forceinline__ double sep_trial_energy(const DevParams& P, const double* X, const double* Y, const double* Z, int N, int g,
                                                  double scale_xy, double scale_z, double tL, double tH, double new_rcut, bool thawed,
                                                  const int* fr = nullptr, double rho = 1.0, bool trial_tan = false) {
    constexpr bool DIRP = POT == PTMC_POT_ELJ_B || POT == PTMC_POT_LUT;
    const Box tbx = make_box(P.bc_kind, tL, tH);
    double acce = 0.0;
    for (int a = 0; a < N - 1; ++a) {
        const double pax = xmul(X[a], scale_xy), pay = xmul(Y[a], scale_xy), paz = xmul(Z[a], scale_z);
        const bool fa = DIRP && !trial_tan && tan_fresh(fr, a);
        for (int j = a + 1 + g; j < N; j += G) {
            const double pbx = xmul(X[j], scale_xy), pby = xmul(Y[j], scale_xy), pbz = xmul(Z[j], scale_z);
            double v0 = 0, v1 = 0, v2 = 0;
            double d2 = (DIRP || thawed) ? dist2_pbc_vec<RHOMBIC>(tbx, pax, pay, paz, pbx, pby, pbz, v0, v1, v2) : 0.0;
            if (!thawed) d2 = dist2_exact(P.bc_kind, pax, pay, paz, pbx, pby, pbz, tL, tH);
            if (DIRP) {
                // xy / z trial: get_tantheta_mat!(new_tan_mat, trial_config); uniform trial: the state's tangents
                const bool stale = !trial_tan && !(fa || tan_fresh(fr, j));
                if (d2 <= new_rcut) acce += pair_energy_stale<POT, NC>(P, d2, v0, v1, v2, rho, stale);
            } else {
                if (d2 <= new_rcut) acce += pair_energy<POT == PTMC_POT_EAM ? PTMC_POT_ELJ_EVEN : POT, NC>(P, d2);
            }
        }
    }
    return acce;
}
// dist2_mat .= new_dist2_mat after an accepted separated move (positions already scaled)
template <int G>
__device__ __forceinline__ void sep_rewrite_matrix(const DevParams& P, const double* X, const double* Y, const double* Z, int N, int g,
                                               double tL, double tH, double* D) {
    for (int a = 0; a < N - 1; ++a)
        for (int j = a + 1 + g; j < N; j += G) {
            const double v = dist2_exact(P.bc_kind, X[a], Y[a], Z[a], X[j], Y[j], Z[j], tL, tH);
            __stcg(D + (size_t)a * N + j, v); __stcg(D + (size_t)j * N + a, v);
        }
}

// power sums: s[k] +/-= x^(3 + k), k = 0..5, under a predicate (the cut-off / self / padding mask of the pair)
template <bool ADD>
__device__ __forceinline__ void psum_acc(double x, bool ok, double (&s)[PTMC_PSUM_K]) {
    // x^3 .. x^8 as a tree (depth 3) rather than a chain of depth 7: the pair loop is latency sensitive.
    // The predicate zeroes x (one select) instead of guarding the six additions: the compiler turned those into a divergent
    // branch per distance (BSSY / BRA / BSYNC: 11 % of the kernel's instructions, ncu) - and a NaN of the self pair never enters.
    // x^5 .. x^8 are leaves of the tree, so their products fuse with the accumulation: 3 DMUL + 2 DADD + 4 DFMA per distance.
    x = ok ? x : 0.0;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    const double m2 = ADD ? x2 : -x2, m3 = ADD ? x3 : -x3, m4 = ADD ? x4 : -x4;
    s[0] += m3;
    s[1] += m4;
    s[2] = fma(m2, x3, s[2]);
    s[3] = fma(m3, x3, s[3]);
    s[4] = fma(m3, x4, s[4]);
    s[5] = fma(m4, x4, s[5]);
    static_assert(PTMC_PSUM_K == 6, "psum_acc is written for six coefficients");
}
// sum_k c_k s_k
__device__ __forceinline__ double psum_energy(const DevParams& P, const double (&s)[PTMC_PSUM_K]) {
    double e = P.coeff[PTMC_PSUM_K - 1] * s[PTMC_PSUM_K - 1];
#pragma unroll
    for (int k = PTMC_PSUM_K - 2; k >= 0; --k) e = fma(P.coeff[k], s[k], e);
    return e;
}

// DBG   : sequential slot stream (Julia-compatible XOSHIRO mode, or REPLAY) + optional trace.  !DBG: BATCH8.
// NITER : compile-time ceil(n_atoms / G) of the pair loop (n_pad = NITER * G), 0 = runtime trip count.
// NC    : compile-time number of ELJ coefficients (6), 0 = runtime.
template <int G, int POT, int GEOM, bool DBG, int NITER, int NC, bool FULLW = false, bool ROT = false>
__device__ __forceinline__ void sweep_body(const DevParams& P, const i64 first_idx, const i64 n_cycles, const int sample) {
    extern __shared__ double sm[];
    constexpr bool EAM = (POT == PTMC_POT_EAM);
    constexpr bool PBC = (GEOM != GEOM_SPHERE);
    constexpr bool rhombic = (GEOM == GEOM_RHOMB);
    constexpr bool BATCH = !DBG;
    constexpr bool DIRP = POT == PTMC_POT_ELJ_B || POT == PTMC_POT_LUT;   // pair energy depends on the direction of the pair
    // ROT (free space, (N - 1) a multiple of G): lane g visits the OTHER atoms i+1+g, i+1+g+G, ... (mod N), so the row takes
    // (N - 1) / G iterations instead of ceil(N / G) (Ne13 on 4 lanes: 3 instead of 4) and needs neither parking nor selects
    constexpr bool PARK = !PBC && POT == PTMC_POT_ELJ_EVEN && !ROT;   // mask by parking atoms far away instead of per-pair selects
    constexpr bool BLK = PARK && G == 4 && !DBG && NITER > 0 && NITER % 2 == 0;   // blocked atom ownership, LDS.128 (see the pair loop)
    constexpr bool PBLK = PBC && POT == PTMC_POT_ELJ_EVEN && G == 16 && !DBG && NITER == 14;   // the same for the pbc216 class (sweep_kernel_p16)
    // ROT kernels: N = NITER G + 1 is a compile-time constant, and every coordinate array is stored twice in a row (atoms 0 .. N-1,
    // then 0 .. N-2 again), so that the rotated walk i+1+g, i+1+g+G, ... needs no wrap-around: all loads of a row are immediate
    // offsets from one address (the index arithmetic was 6 integer instructions per pair iteration)
    const int T = P.n_traj, N = ROT ? NITER * G + 1 : P.n_atoms;
    const int NP = ROT ? 2 * (NITER * G + 1) : (NITER ? NITER * G : P.n_pad);
    const int lad = blockIdx.x;
    const int tid = threadIdx.x;
    const int slot = tid / G, g = tid & (G - 1);
    const bool active = slot < T;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((tid & 31) & ~(G - 1)));
    // hot-path mask: when every group of the warp is active all 32 lanes run the move loop in lock step, and a compile-time
    // full mask spares the MATCH / VOTE convergence sequences the compiler emits around a register mask (FULLW kernels)
    const unsigned hmask = FULLW ? 0xffffffffu : gmask;
    constexpr bool EAM_REG = EAM && NITER > 0 && PTMC_EAM_UNROLL;     // trial (phi, rho) components in registers instead of shared memory
    const bool ties = PBC && P.use_ties != 0;      // periodic NPT, uniform volume moves: tie table instead of dist2_mat
    // ... and with it, for ELJPotentialEven{6}: power sums instead of the O(N^2) trial energy of a volume move
    constexpr bool PSUM = PBC && POT == PTMC_POT_ELJ_EVEN && NC == PTMC_PSUM_K;
    const bool psum = PSUM && ties;
    const SmemLayout lay = smem_layout(T, NP, POT, (int)(blockDim.x >> 5), PBC, !EAM_REG, !ties, G, P.cfg_mod);
    double* scal = sm + lay.off_scal;
    i64* exc_cnt = reinterpret_cast<i64*>(sm + lay.off_exc);
    double* hams = sm + lay.off_ham;
    u64* lrng = reinterpret_cast<u64*>(sm + lay.off_lrng);
    int* perm = reinterpret_cast<int*>(sm + lay.off_perm);
    int* moved = reinterpret_cast<int*>(sm + lay.off_moved);
    double* trho = sm + lay.off_tan;                       // stale-tangent factor per configuration
    int* fresh = reinterpret_cast<int*>(trho + T);          // atoms moved since the last xy / z trial, [T][FW]
    double* psp = sm + lay.off_psum;                                  // power sums per configuration and lane, [T][6][G]
    int* psv = reinterpret_cast<int*>(psp + T * PTMC_PSUM_K * G);     // ... valid in this launch (built at the first volume move)
    const int FW = (NP + 31) / 32;
    const int MW = P.moved_words + 1;            // ints per configuration in moved[]
    const bool npt = PBC && P.ensemble == PTMC_ENS_NPT;
    const i64 rep = (i64)lad * T + slot;      // replica (slot) index in the global arrays
    const double* eam_tab = sm + lay.off_tab;
    if (EAM) for (int k = tid; k < PTMC_EAM_KTAB_DOUBLES; k += blockDim.x) sm[lay.off_tab + k] = P.eam_ktab[k];

    // ---- load: slot s starts with configuration buffer s ----------------------------------------------
    if (active) {
        const int c0 = PBC ? P.perm_g[rep] : slot;   // periodic: the dist2_mat in HBM is indexed by configuration
        double* X = sm + (size_t)c0 * lay.cfg_stride;
        double* Y = X + NP; double* Z = Y + NP;
        const double* gp = P.pos + rep * 3 * N;
        for (int jj = g; jj < NP; jj += G) {
            const int j = (ROT && jj >= N) ? jj - N : jj;       // ROT: second copy
            bool v = j < N;
            const double pad = PARK ? PTMC_FAR : 0.0;
            X[jj] = v ? gp[3 * j] : pad; Y[jj] = v ? gp[3 * j + 1] : pad; Z[jj] = v ? gp[3 * j + 2] : pad;
            if (EAM) {
                Z[NP + j] = v ? P.comp[(rep * N + j) * 2] : 1.0;          // phi
                Z[2 * NP + j] = v ? P.comp[(rep * N + j) * 2 + 1] : 1.0;  // rho
                if (!EAM_REG) { Z[3 * NP + j] = 1.0; Z[4 * NP + j] = 1.0; }
            }
        }
        if (g == 0) {
            double L = P.box[rep * 2], h = P.box[rep * 2 + 1];
            scal[4 * c0] = P.en_tot[rep];
            scal[4 * c0 + 1] = L; scal[4 * c0 + 2] = h;
            scal[4 * c0 + 3] = npt ? get_r_cut(P.bc_kind, L, h) : 0.0;
            perm[slot] = c0;
            if (PSUM && psum) psv[c0] = 0;
            if (PBC) for (int w = 0; w < MW; ++w) moved[c0 * MW + w] = P.moved_g[((i64)lad * T + c0) * MW + w];
            if (PBC && DIRP) {
                const bool st = P.separated != 0;
                trho[c0] = st ? P.tan_rho[(i64)lad * T + c0] : 1.0;
                for (int w = 0; w < FW; ++w) fresh[c0 * FW + w] = (st && w < P.moved_words) ? P.fresh_g[((i64)lad * T + c0) * P.moved_words + w] : 0;
            }
            exc_cnt[2 * slot] = 0; exc_cnt[2 * slot + 1] = 0;
            hams[2 * slot] = P.ham[rep * 2]; hams[2 * slot + 1] = P.ham[rep * 2 + 1];
        }
    }
    // slot-side registers (identical in every lane of the group)
    Stream<DBG> S;      // sequential stream (DBG kernels)
    BatchRng<G> B;      // BATCH8 substreams (production kernels)
    double beta = 0, dmax = 0, vmax = 0, max_boxl = 0;
    double dmax53 = 0;      // max_displ[1] * 2^-53 (BatchRng::centered)
    int n_acc_atom = 0, n_adj_atom = 0, n_acc_vol = 0, n_adj_vol = 0;   // this launch / since last adjustment
    if (active) {
        beta = P.betas[slot];
        dmax = P.max_displ[rep * 4]; dmax53 = dmax * 0x1.0p-53;
        n_adj_atom = (int)P.counters[rep * PTMC_NCOUNT + PTMC_C_ATOM_ADJ];
        if (PBC) {
            vmax = P.max_displ[rep * 4 + 1];
            max_boxl = P.max_boxlength[rep];
            n_adj_vol = (int)P.counters[rep * PTMC_NCOUNT + PTMC_C_VOL_ADJ];
            if (g == 0) {   // separated volume moves: rarely touched slot-side state lives in shared memory
                double* sd = sm + lay.off_sep + 6 * slot; int* si = reinterpret_cast<int*>(sd + 3);
                sd[0] = P.max_displ[rep * 4 + 2]; sd[1] = P.max_displ[rep * 4 + 3]; sd[2] = P.max_boxheight[rep];
                si[0] = 0; si[1] = (int)P.counters[rep * PTMC_NCOUNT + PTMC_C_VOLXY_ADJ];
                si[2] = 0; si[3] = (int)P.counters[rep * PTMC_NCOUNT + PTMC_C_VOLZ_ADJ];
            }
        }
        S.draws = nullptr; S.pos = 0; S.len = 0; S.exhausted = 0;
        if (BATCH) {
            B.load(P.slot_rng + rep * (4 * PTMC_RNG_LANES), g);
        } else {
            S.r.s0 = P.slot_rng[rep * 4]; S.r.s1 = P.slot_rng[rep * 4 + 1];
            S.r.s2 = P.slot_rng[rep * 4 + 2]; S.r.s3 = P.slot_rng[rep * 4 + 3];
            if (P.replay) { S.draws = P.slot_draws + rep * P.slot_len; S.len = P.slot_len; S.pos = P.slot_cursor[rep]; }
        }
    }
    // ladder ("driver task") stream: every warp keeps a private copy in shared memory, stepped by its lane 0
    // in lock-step with the other warps, so cycles without an exchange attempt (90 %) need no CTA barrier
    u64* wl = lrng + 6 * (tid >> 5);
    if ((tid & 31) == 0) {
        wl[0] = P.ladder_rng[lad * 4]; wl[1] = P.ladder_rng[lad * 4 + 1];
        wl[2] = P.ladder_rng[lad * 4 + 2]; wl[3] = P.ladder_rng[lad * 4 + 3];
        wl[4] = (DBG && P.replay) ? (u64)P.ladder_cursor[lad] : 0;
        pair_barrier_init(wl + 5);        // exchange meeting point of this warp and the next one
    }
    // cycle phases instead of 64-bit modulo in the loop
    int adj_phase = (int)(first_idx % P.n_adjust);
    int smp_phase = (int)(first_idx % P.mc_sample);
    __syncthreads();

#ifdef PTMC_PROD_TRACE         // test-only build (lib/libptmc_b200_trace.so): the production instantiations write the per-move trace
    const bool tracing = P.trace_on != 0;      // under a runtime flag; 3 instructions per move (0.8 % on C2), so the shipped library leaves it out
#else
    const bool tracing = DBG && P.trace_on != 0;
#endif
    int err = 0;
    unsigned xpar = 0;      // phase parities of the two exchange mbarriers this warp takes part in (bit 0: with the next warp, bit 1: with the previous one)

    // P.phase splits mc_cycle! at the exchange (a ladder that spans GPUs exchanges on the host, ptmc_run_phase)
    const int phase = P.phase;
    for (i64 cyc = 0; cyc < n_cycles; ++cyc) {
        if (active && phase != 2) {
            // ================================================================== mc_step! for this slot ==
            const int c = perm[slot];
            if (PSUM && psum && ((first_idx + cyc) & 255) == 0) {     // bound the rounding drift of the carried sums: rebuild now and then
                __syncwarp(gmask);
                if (g == 0) psv[c] = 0;
                __syncwarp(gmask);
            }
            double* X = sm + (size_t)c * lay.cfg_stride;
            double* Y = X + NP; double* Z = Y + NP;
            double en = scal[4 * c];
            double boxL = scal[4 * c + 1], boxH = 0.0, rcut = 0.0;
            Box bx;
            if (PBC) { boxH = scal[4 * c + 2]; rcut = scal[4 * c + 3]; bx = make_box(P.bc_kind, boxL, boxH); }
            const double R2 = boxL;   // spherical: box[0] = radius^2

            for (int step = 0; step < P.n_steps; ++step) {
                int index;                                          // rand(1:N) src/MCRun.jl:145
                if (BATCH) { B.begin(); index = (int)__umul64hi(B.template raw<0>(hmask), (u64)P.n_steps) + 1; }
                else index = S.range((u64)P.n_steps);
                double de = 0.0;
                bool acc = false;
                if (!PBC || index <= P.n_atom_moves) {
                    // ------------------------------------------------ atom_displacement src/MCMoves.jl:33-145
                    const int i = index - 1;
                    double xi = X[i], yi = Y[i], zi = Z[i];
                    double tx, ty, tz;
                    if (BATCH) {
                        tx = xadd(xi, xmul(B.template centered<1>(hmask), dmax53));
                        ty = xadd(yi, xmul(B.template centered<2>(hmask), dmax53));
                        tz = xadd(zi, xmul(B.template centered<3>(hmask), dmax53));
                    } else {
                        tx = xadd(xi, xmul(xsub(S.rand(), 0.5), dmax));
                        ty = xadd(yi, xmul(xsub(S.rand(), 0.5), dmax));
                        tz = xadd(zi, xmul(xsub(S.rand(), 0.5), dmax));
                    }
                    if (!PBC) {
                        int count = 0, redraws = 0;
                        while (xadd(xadd(xmul(tx, tx), xmul(ty, ty)), xmul(tz, tz)) > R2) {   // check_boundary
                            ++count;
                            if (count == 50) {
                                // recentre! src/Configurations.jl:79-85: sequential means, then shift
                                __syncwarp(gmask);
                                double cx = 0, cy = 0, cz = 0;
                                for (int j = 0; j < N; ++j) cx = xadd(cx, X[j]);
                                for (int j = 0; j < N; ++j) cy = xadd(cy, Y[j]);
                                for (int j = 0; j < N; ++j) cz = xadd(cz, Z[j]);
                                cx = xdiv(cx, (double)N); cy = xdiv(cy, (double)N); cz = xdiv(cz, (double)N);
                                __syncwarp(gmask);
                                for (int j = g; j < N; j += G) {
                                    const double nx = xsub(X[j], cx), ny = xsub(Y[j], cy), nz = xsub(Z[j], cz);
                                    X[j] = nx; Y[j] = ny; Z[j] = nz;
                                    if (ROT) { X[j + N] = nx; Y[j + N] = ny; Z[j + N] = nz; }
                                }
                                __syncwarp(gmask);
                                xi = X[i]; yi = Y[i]; zi = Z[i];
                            } else {
                                if (BATCH) {
                                    if (++redraws == 1) {
                                        tx = xadd(xi, xmul(B.template centered<5>(gmask), dmax53));
                                        ty = xadd(yi, xmul(B.template centered<6>(gmask), dmax53));
                                        tz = xadd(zi, xmul(B.template centered<7>(gmask), dmax53));
                                    } else {
                                        B.begin();
                                        tx = xadd(xi, xmul(B.template centered<0>(gmask), dmax53));
                                        ty = xadd(yi, xmul(B.template centered<1>(gmask), dmax53));
                                        tz = xadd(zi, xmul(B.template centered<2>(gmask), dmax53));
                                    }
                                } else {
                                    tx = xadd(xi, xmul(xsub(S.rand(), 0.5), dmax));
                                    ty = xadd(yi, xmul(xsub(S.rand(), 0.5), dmax));
                                    tz = xadd(zi, xmul(xsub(S.rand(), 0.5), dmax));
                                }
                                if (count == 100) { err = PTMC_ERR_SPHERE; break; }
                            }
                            if (DBG) { if (S.exhausted) break; }
                        }
                    } else if (P.bc_kind == PTMC_BC_CUBIC) {           // src/MCMoves.jl:50-64
                        tx = xsub(tx, xmul(boxL, rint(xdiv(tx, boxL))));
                        ty = xsub(ty, xmul(boxL, rint(xdiv(ty, boxL))));
                        tz = xsub(tz, xmul(boxL, rint(xdiv(tz, boxL))));
                    } else if (P.bc_kind == PTMC_BC_RECTANGULAR) {     // :85-98
                        tx = xsub(tx, xmul(boxL, rint(xdiv(tx, boxL))));
                        ty = xsub(ty, xmul(boxL, rint(xdiv(ty, boxL))));
                        tz = xsub(tz, xmul(boxH, rint(xdiv(tz, boxH))));
                    } else {                                            // Rhombic :66-83
                        const double s3 = sqrt(3.0);
                        double ry = rint(xdiv(xsub(ty, xdiv(xmul(boxL, s3), 4.0)), xdiv(xmul(boxL, s3), 2.0)));
                        double sx = xadd(xmul(boxL, rint(xdiv(xsub(xsub(tx, xdiv(ty, s3)), xdiv(boxL, 2.0)), boxL))),
                                         xmul(xdiv(boxL, 2.0), ry));
                        double sy = xmul(xdiv(xmul(boxL, s3), 2.0), ry);
                        double sz = xmul(boxH, rint(xdiv(xsub(tz, xdiv(boxH, 2.0)), boxH)));
                        tx = xsub(tx, sx); ty = xsub(ty, sy); tz = xsub(tz, sz);
                    }
                    // -------------------------------- energy_update! src/EnergyEvaluation.jl:1474-1506,1567-1591
                    double new_en;
                    double dS[PTMC_PSUM_K];                 // power sums of the new row minus those of the old row (this lane's pairs)
                    if (PSUM) {
#pragma unroll
                        for (int k = 0; k < PTMC_PSUM_K; ++k) dS[k] = 0.0;
                    }
                    constexpr int EAM_NK = EAM_REG ? NITER : 1;
                    double nph[EAM_NK], nrh[EAM_NK], nphi_i = 0.0, nrho_i = 0.0;   // EAM: trial components of this lane's atoms / of atom i
                    if (PARK) {
                        // park the moved atom: its own pair term (d2_old = 0) must not enter the row sums
                        __syncwarp(hmask);
                        if (g == 0) { X[i] = PTMC_FAR; Y[i] = PTMC_FAR; Z[i] = PTMC_FAR; }
                        __syncwarp(hmask);
                    }
                    if (!EAM) {
                        // dimer_energy_update! :804-821: sum_j e(new_ij) - sum_j e(old_ij) (the reference forms the same
                        // two row sums, :808-810).
                        double accn = 0.0, acco = 0.0;
                        if (PBC) {
                            // periodic: while some atom of the configuration has never moved, the old row is
                            // dist2_mat[index, :] read from HBM (frozen values: exactly the reference's d2 <= r_cut decisions for
                            // lattice pairs that sit ON the cut-off) and the new row goes to new_dist2_vec; afterwards
                            // ("thawed") no exact tie can exist and the old row is recomputed from the positions.
                            const bool thawed = moved[c * MW + MW - 1] >= N;
                            const double* Drow = P.d2mat + (((size_t)lad * T + c) * N + i) * N;
                            double* NEW = Z + NP;
                            // tangent-dependent potentials with separated volume moves: stale tangents of the old row (sep_trial_energy)
                            const bool tanst = DIRP && P.separated != 0;
                            const double rho_c = tanst ? trho[c] : 1.0;
                            const int* fr = fresh + c * FW;
                            const bool fresh_i = tanst && tan_fresh(fr, i);
                            if (thawed || ties) {
                                // thawed: both rows from the positions in shared memory, pair loop unrolled like the free-space one
                                const double* Xg = X + g;
                                auto pbc_pair = [&](int j, double xj, double yj, double zj, auto ps_tag) {
                                    constexpr bool PS = decltype(ps_tag)::value;
                                    double vn0 = 0, vn1 = 0, vn2 = 0, vo0 = 0, vo1 = 0, vo2 = 0;
                                    const double d2n = DIRP ? dist2_pbc_vec<rhombic>(bx, tx, ty, tz, xj, yj, zj, vn0, vn1, vn2)
                                                            : dist2_pbc<rhombic>(bx, tx, ty, tz, xj, yj, zj);
                                    const double d2o = DIRP ? dist2_pbc_vec<rhombic>(bx, xi, yi, zi, xj, yj, zj, vo0, vo1, vo2)
                                                            : dist2_pbc<rhombic>(bx, xi, yi, zi, xj, yj, zj);
                                    const bool valid = (j != i) && (j < N);
                                    if (POT == PTMC_POT_ELJ_EVEN) {
                                        // beyond the cut-off (:323-336), self or padding: the accumulation is predicated off instead of
                                        // selecting a far distance (no FSEL; a NaN of the self pair, d2o = 0, never reaches the sums)
                                        const bool okn = valid && (!npt || d2n <= rcut), oko = valid && (!npt || d2o <= rcut);
                                        const double y = rcp_pos3(d2n * d2o);
                                        const double xn = y * d2o, xo = y * d2n;
                                        if (PS) {
                                            // the row difference of the six power sums instead of the two Horner forms: they give dE (below)
                                            // and, on acceptance, carry the configuration's S_k along
                                            if (PSUM) { psum_acc<true>(xn, okn, dS); psum_acc<false>(xo, oko, dS); }
                                        } else {
                                            const double pn = elj_even_poly<NC>(P, xn), po = elj_even_poly<NC>(P, xo);
                                            if (okn) accn = fma(pn, xn * xn * xn, accn);
                                            if (oko) acco = fma(po, xo * xo * xo, acco);
                                        }
                                    } else {
                                        const double dn = valid ? d2n : 1.0, dold = valid ? d2o : 1.0;
                                        double en_j, eo_j;
                                        if (DIRP && tanst) {
                                            en_j = pair_energy_dir<POT, NC>(P, dn, vn0, vn1, vn2);
                                            eo_j = pair_energy_stale<POT, NC>(P, dold, vo0, vo1, vo2, rho_c, !(fresh_i || (j < N && tan_fresh(fr, j))));
                                        } else {
                                            en_j = DIRP ? pair_energy_dir<POT, NC>(P, dn, vn0, vn1, vn2) : pair_energy<POT, NC>(P, dn);
                                            eo_j = DIRP ? pair_energy_dir<POT, NC>(P, dold, vo0, vo1, vo2) : pair_energy<POT, NC>(P, dold);
                                        }
                                        if (npt) { en_j = (dn <= rcut) ? en_j : 0.0; eo_j = (dold <= rcut) ? eo_j : 0.0; }
                                        accn += valid ? en_j : 0.0; acco += valid ? eo_j : 0.0;
                                    }
                                };
                                auto pbc_term = [&](int k, auto ps_tag) { pbc_pair(g + k * G, Xg[k * G], Xg[NP + k * G], Xg[2 * NP + k * G], ps_tag); };
                                auto pbc_row = [&](auto ps_tag) {
                                    if (PBLK) {
                                        // 16 lanes, 14 atoms per lane: lane g owns the block of atoms 14 g .. 14 g + 13 and reads them two at a
                                        // time (LDS.128; the 8 lanes of a quarter warp, 112 bytes apart, cover all banks)
                                        const double* Xb = X + g * NITER;
#pragma unroll
                                        for (int k = 0; k < NITER; k += 2) {
                                            const double2 xx = *reinterpret_cast<const double2*>(Xb + k);
                                            const double2 yy = *reinterpret_cast<const double2*>(Xb + NP + k);
                                            const double2 zz = *reinterpret_cast<const double2*>(Xb + 2 * NP + k);
                                            pbc_pair(g * NITER + k, xx.x, yy.x, zz.x, ps_tag);
                                            pbc_pair(g * NITER + k + 1, xx.y, yy.y, zz.y, ps_tag);
                                        }
                                    } else if (NITER) {
#pragma unroll
                                        for (int k = 0; k < NITER; ++k) pbc_term(k, ps_tag);
                                    } else {
                                        const int nk = NP / G;
                                        for (int k = 0; k < nk; ++k) pbc_term(k, ps_tag);
                                    }
                                };
                                if (PSUM && psum) { pbc_row(std::true_type{}); accn = psum_energy(P, dS); }
                                else pbc_row(std::false_type{});
                            } else
                            for (int j = g; j < N; j += G) {
                                const double xj = X[j], yj = Y[j], zj = Z[j];
                                double vn0 = 0, vn1 = 0, vn2 = 0, vo0 = 0, vo1 = 0, vo2 = 0;
                                double d2n = DIRP ? dist2_pbc_vec<rhombic>(bx, tx, ty, tz, xj, yj, zj, vn0, vn1, vn2) : dist2_pbc<rhombic>(bx, tx, ty, tz, xj, yj, zj);
                                double d2o = __ldcg(Drow + j);
                                if (DIRP) (void)dist2_pbc_vec<rhombic>(bx, xi, yi, zi, xj, yj, zj, vo0, vo1, vo2);   // direction of the old pair
                                const bool valid = (j != i);
                                NEW[j] = valid ? d2n : 0.0;                       // new_dist2_vec[index] = 0, MCMoves.jl:111
                                if (POT == PTMC_POT_ELJ_EVEN) {
                                    // one shared reciprocal for the pair; beyond the cut-off (:323-336) or self -> contributes exactly 0
                                    const double a = (valid && (!npt || d2n <= rcut)) ? d2n : PTMC_PAIR_OFF;
                                    const double b = (valid && (!npt || d2o <= rcut)) ? d2o : PTMC_PAIR_OFF;
                                    elj_even_pair_acc<NC>(P, a, b, accn, acco);
                                } else {
                                    d2n = valid ? d2n : 1.0; d2o = valid ? d2o : 1.0;
                                    double en_j, eo_j;
                                    if (DIRP && tanst) {
                                        en_j = pair_energy_dir<POT, NC>(P, d2n, vn0, vn1, vn2);
                                        eo_j = pair_energy_stale<POT, NC>(P, d2o, vo0, vo1, vo2, rho_c, !(fresh_i || tan_fresh(fr, j)));
                                    } else {
                                        en_j = DIRP ? pair_energy_dir<POT, NC>(P, d2n, vn0, vn1, vn2) : pair_energy<POT, NC>(P, d2n);
                                        eo_j = DIRP ? pair_energy_dir<POT, NC>(P, d2o, vo0, vo1, vo2) : pair_energy<POT, NC>(P, d2o);
                                    }
                                    if (npt) { en_j = (d2n <= rcut) ? en_j : 0.0; eo_j = (d2o <= rcut) ? eo_j : 0.0; }   // :323-336
                                    accn += valid ? en_j : 0.0; acco += valid ? eo_j : 0.0;
                                }
                            }
                            if (ties && !thawed && !((moved[c * MW + (i >> 5)] >> (i & 31)) & 1)) {
                                // never-moved atom of a configuration that still holds cut-off ties: its tie partners that never moved either
                                // take the frozen dist2_mat entry in the old row (the d2 <= r_cut decision of the reference) instead of the
                                // recomputed distance the loop above used
                                const size_t row = ((size_t)lad * T + c) * N + i;
                                const int cnt = __ldg(P.tie_cnt + row);
                                for (int t = g; t < cnt; t += G) {
                                    const int j = __ldg(P.tie_idx + row * PTMC_TIE_K + t);
                                    if ((moved[c * MW + (j >> 5)] >> (j & 31)) & 1) continue;
                                    const double d2f = __ldcg(P.tie_val + row * PTMC_TIE_K + t);
                                    double v0 = 0, v1 = 0, v2 = 0;      // the same distance function as the row loop, so that the recomputed term cancels
                                    const double d2r = DIRP ? dist2_pbc_vec<rhombic>(bx, xi, yi, zi, X[j], Y[j], Z[j], v0, v1, v2)
                                                            : dist2_pbc<rhombic>(bx, xi, yi, zi, X[j], Y[j], Z[j]);
                                    auto old_term = [&](double d) {
                                        if (npt && !(d <= rcut)) return 0.0;
                                        return DIRP ? pair_energy_dir<POT, NC>(P, d, v0, v1, v2) : pair_energy<POT == PTMC_POT_EAM ? PTMC_POT_ELJ_EVEN : POT, NC>(P, d);
                                    };
                                    acco += old_term(d2f) - old_term(d2r);
                                    // the pair was an active tie until now, i.e. outside the power sums: take back what the row loop subtracted
                                    if (PSUM && psum) psum_acc<true>(rcp_pos(d2r), d2r <= rcut, dS);
                                }
                            }
                        } else {
                            // free space: the old row is recomputed from the positions (no dist2_mat is stored)
                            const double* Xg = ROT ? X + (i + 1 + g) : X + g;
                            auto pair_term = [&](int k) {
                                const int j = g + k * G;
                                // ROT: Xg points at atom i + 1 + g of the doubled arrays
                                const double xj = Xg[k * G], yj = Xg[NP + k * G], zj = Xg[2 * NP + k * G];
                                const bool valid = ROT || ((j != i) && (j < N));   // not the moved atom, not padding
                                double bx_ = xi - xj, by_ = yi - yj, bz_ = zi - zj;
                                double d2o = fma(bz_, bz_, fma(by_, by_, bx_ * bx_));
                                if (POT == PTMC_POT_ELJ_EVEN) {
                                    // (forming d2n from d2o, |b + s|^2 = |b|^2 + 2 s.b + |s|^2, saves 2 of the 34 FP64 instructions per pair and
                                    // measured 1.4 % SLOWER: the longer dependency chain costs more `wait` stalls than the issue slots give back,
                                    // profiles/r02_experiments.md)
                                    const double ax = tx - xj, ay = ty - yj, az = tz - zj;
                                    const double d2n = fma(az, az, fma(ay, ay, ax * ax));
                                    elj_even_pair_acc<NC>(P, d2n, d2o, accn, acco);     // parked atoms contribute exactly 0
                                } else {
                                    double ax = tx - xj, ay = ty - yj, az = tz - zj;
                                    double d2n = fma(az, az, fma(ay, ay, ax * ax));
                                    // the moved atom itself (d2o = 0 -> NaN) and the padding atoms are discarded here
                                    const double en_j = POT == PTMC_POT_ELJ_B ? eljb_energy<NC>(P, d2n, az * az)
                                                      : POT == PTMC_POT_LUT ? lut_energy(P, d2n, ax, ay, az) : pair_energy<POT, NC>(P, d2n);
                                    const double eo_j = POT == PTMC_POT_ELJ_B ? eljb_energy<NC>(P, d2o, bz_ * bz_)
                                                      : POT == PTMC_POT_LUT ? lut_energy(P, d2o, bx_, by_, bz_) : pair_energy<POT, NC>(P, d2o);
                                    accn += valid ? en_j : 0.0; acco += valid ? eo_j : 0.0;
                                }
                            };
                            if (BLK) {
                                // lane g owns the block of atoms g NITER .. g NITER + NITER - 1 and reads them two at a time (LDS.128):
                                // 1.5 instead of 3 shared-memory loads per pair (any split of the row over the lanes gives the same sum set)
                                const double* Xb = X + g * NITER;
#pragma unroll
                                for (int k = 0; k < NITER; k += 2) {
                                    const double2 xx = *reinterpret_cast<const double2*>(Xb + k);
                                    const double2 yy = *reinterpret_cast<const double2*>(Xb + NP + k);
                                    const double2 zz = *reinterpret_cast<const double2*>(Xb + 2 * NP + k);
                                    {
                                        const double ax = tx - xx.x, ay = ty - yy.x, az = tz - zz.x, bx_ = xi - xx.x, by_ = yi - yy.x, bz_ = zi - zz.x;
                                        elj_even_pair_acc<NC>(P, fma(az, az, fma(ay, ay, ax * ax)), fma(bz_, bz_, fma(by_, by_, bx_ * bx_)), accn, acco);
                                    }
                                    {
                                        const double ax = tx - xx.y, ay = ty - yy.y, az = tz - zz.y, bx_ = xi - xx.y, by_ = yi - yy.y, bz_ = zi - zz.y;
                                        elj_even_pair_acc<NC>(P, fma(az, az, fma(ay, ay, ax * ax)), fma(bz_, bz_, fma(by_, by_, bx_ * bx_)), accn, acco);
                                    }
                                }
                            } else if (NITER) {
#pragma unroll
                                for (int k = 0; k < NITER; ++k) pair_term(k);
                            } else {
                                const int nk = NP / G;
#pragma unroll 2
                                for (int k = 0; k < nk; ++k) pair_term(k);
                            }
                        }
                        const double accd = accn - acco;
                        const double delta_en = group_sum<G>(accd, hmask);
                        new_en = xadd(delta_en, en);
                    } else {
                        // calc_components :1111-1127 + calc_energies_from_components :1156-1162
                        double* PHI = Z + NP; double* RHO = PHI + NP;
                        double sphi = 0.0, srho = 0.0, se = 0.0;
                        auto eam_term = [&](int j, double& nphi, double& nrho) {
                            const double xj = X[j], yj = Y[j], zj = Z[j];
                            double ax = tx - xj, ay = ty - yj, az = tz - zj;
                            double d2n = fma(az, az, fma(ay, ay, ax * ax));
                            double bx_ = xi - xj, by_ = yi - yj, bz_ = zi - zj;
                            double d2o = fma(bz_, bz_, fma(by_, by_, bx_ * bx_));
                            const bool valid = (j != i) && (j < N);
                            d2n = valid ? d2n : 1.0; d2o = valid ? d2o : 1.0;
                            double pn, rn, po, ro;
                            eam_invrexp(P, eam_tab, P.eam_etab, d2n, pn, rn);
                            eam_invrexp(P, eam_tab, P.eam_etab, d2o, po, ro);
                            double tp = pn - po, tr = rn - ro;          // exactly 0 when !valid
                            nphi = PHI[j] + tp; nrho = RHO[j] + tr;
                            sphi += tp; srho += tr;
                            // masked by a 0 / 1 factor rather than a select of the term: the compiler guarded the square root with a divergent
                            // branch per pair iteration (BSSY / BRA / BSYNC, reconvergence in 46 % of the iterations: ncu, profiles/r02_c3_eam.md)
                            se = fma(valid ? 1.0 : 0.0, fma(-P.eam_ecam, sqrt_pos(nrho), P.eam_ean * nphi), se);
                        };
                        if (EAM_REG) {
#pragma unroll
                            for (int k = 0; k < EAM_NK; ++k) eam_term(g + k * G, nph[k], nrh[k]);
                        } else {
                            double* NPHI = RHO + NP; double* NRHO = NPHI + NP;
                            for (int j = g; j < NP; j += G) {
                                double nphi, nrho;
                                eam_term(j, nphi, nrho);
                                NPHI[j] = nphi; NRHO[j] = nrho;
                            }
                        }
                        sphi = group_sum<G>(sphi, hmask); srho = group_sum<G>(srho, hmask); se = group_sum<G>(se, hmask);
                        nphi_i = PHI[i] + sphi; nrho_i = RHO[i] + srho;
                        new_en = se + fma(-P.eam_ecam, sqrt_pos(nrho_i), P.eam_ean * nphi_i);
                    }
                    // ------------------------------------------------------------ acc_test! src/MCRun.jl:122-132
                    de = xsub(new_en, en);
                    const double u = BATCH ? B.template rand<4>(hmask) : S.rand();
                    acc = metropolis_accept(-xmul(de, beta), u);
                    if (PARK) {
                        if (!acc && g == 0) { X[i] = xi; Y[i] = yi; Z[i] = zi; }        // un-park
                    }
                    if (acc) {
                        // swap_atom_config! src/swap_config.jl:40-51 (+ swap_vars! :208-211 for EAM)
                        if (ROT) { if (g < 2) { const int o = g ? i + N : i; X[o] = tx; Y[o] = ty; Z[o] = tz; } }     // both copies (G >= 2)
                        else if (g == 0) { X[i] = tx; Y[i] = ty; Z[i] = tz; }
                        en = new_en;
                        n_acc_atom += 1; n_adj_atom += 1;
                        if (PBC) {
                            if (DIRP && P.separated != 0) {
                                // tan_mat[index, :] = tan_mat[:, index] = new_tan_vec  (swap_vars! src/swap_config.jl:197-206): row and column are fresh
                                __syncwarp(gmask);
                                if (g == 0) fresh[c * FW + (i >> 5)] |= 1 << (i & 31);
                            }
                            if (PSUM && psum && psv[c]) {
                                double* sp = psp + (size_t)c * PTMC_PSUM_K * G + g;
#pragma unroll
                                for (int k = 0; k < PTMC_PSUM_K; ++k) sp[k * G] += dS[k];
                            }
                            if (moved[c * MW + MW - 1] < N) {
                                if (!ties) {
                                    // dist2_mat[index, :] = dist2_mat[:, index] = new_dist2_vec  (swap_config.jl:45-46)
                                    double* D = P.d2mat + ((size_t)lad * T + c) * N * N;
                                    const double* NEW = Z + NP;
                                    for (int j = g; j < N; j += G) {
                                        const double v = NEW[j];
                                        __stcg(D + (size_t)i * N + j, v);
                                        __stcg(D + (size_t)j * N + i, v);
                                    }
                                }
                                __syncwarp(gmask);
                                if (g == 0) {
                                    const int w = i >> 5, bit = 1 << (i & 31);
                                    if (!(moved[c * MW + w] & bit)) { moved[c * MW + w] |= bit; moved[c * MW + MW - 1] += 1; }
                                }
                            }
                        }
                        if (EAM) {
                            double* PHI = Z + NP; double* RHO = PHI + NP;
                            __syncwarp(gmask);            // inside `if (acc)`: only the accepting groups are here
                            if (EAM_REG) {
#pragma unroll
                                for (int k = 0; k < EAM_NK; ++k) {
                                    const int j = g + k * G;
                                    PHI[j] = (j == i) ? nphi_i : nph[k]; RHO[j] = (j == i) ? nrho_i : nrh[k];
                                }
                            } else {
                                const double* NPHI = RHO + NP; const double* NRHO = NPHI + NP;
                                for (int j = g; j < N; j += G) { PHI[j] = (j == i) ? nphi_i : NPHI[j]; RHO[j] = (j == i) ? nrho_i : NRHO[j]; }
                            }
                        }
                    }
                    __syncwarp(hmask);
                } else {
                    // ------------------------------- volume_change_uniform src/MCMoves.jl:277-287, :169-177
                    // ------------------------------- volume_change_separated :300-343 (NPT.separated_volume)
                    const bool sepv = P.separated != 0;
                    double* sepd = sm + lay.off_sep + 6 * slot; int* sepi = reinterpret_cast<int*>(sepd + 3);
                    double scale_xy, scale_z;
                    int xyz = 0;                                  // NPTVariables.xy_or_z: 0 all directions, 1 xy, 2 z
                    if (sepv) {
                        int ra; double uv;
                        if (BATCH) { ra = (int)__umul64hi(B.template raw<2>(gmask), 6ull) + 1; uv = B.template rand<1>(gmask); }
                        else { ra = S.range(6ull); uv = S.rand(); }
                        const double lh = xdiv(boxL, boxH), lhr = xdiv(max_boxl, sepd[2]);     // lh_ratio src/MCStates.jl:107
                        if (ra == 1) {                            // volume_change_z :256-271
                            xyz = 2;
                            double sc = exp(xmul(xsub(uv, 0.5), sepd[1]));
                            if (lh <= xmul(lhr, 1.1) && sc > 1.0) sc = xdiv(1.0, sc);
                            else if (lh >= xmul(lhr, 0.909) && sc < 1.0) sc = xdiv(1.0, sc);
                            if (boxH >= sepd[2] && sc > 1.0) sc = xdiv(1.0, sc);
                            scale_xy = 1.0; scale_z = sc;
                        } else if (ra <= 3) {                     // volume_change_xy :232-247
                            xyz = 1;
                            double sc = pow(exp(xmul(xsub(uv, 0.5), sepd[0])), 1.0 / 2.0);
                            if (lh >= xmul(lhr, 1.1) && sc > 1.0) sc = xdiv(1.0, sc);
                            else if (lh <= xmul(lhr, 0.909) && sc < 1.0) sc = xdiv(1.0, sc);
                            if (boxL >= max_boxl && sc > 1.0) sc = xdiv(1.0, sc);
                            scale_xy = sc; scale_z = 1.0;
                        } else {                                  // volume_change_xyz :169-177
                            double sc = pow(exp(xmul(xsub(uv, 0.5), vmax)), 1.0 / 3.0);
                            if (boxL >= max_boxl && sc > 1.0) sc = 1.0;
                            scale_xy = sc; scale_z = sc;
                        }
                    } else {
                        const double uv = BATCH ? B.template rand<1>(gmask) : S.rand();
                        double scale = pow(exp(xmul(xsub(uv, 0.5), vmax)), 1.0 / 3.0);
                        if (boxL >= max_boxl && scale > 1.0) scale = 1.0;
                        scale_xy = scale; scale_z = scale;
                    }
                    const double tL = xmul(scale_xy, boxL), tH = xmul(scale_z, boxH);
                    const double new_rcut = get_r_cut(P.bc_kind, tL, tH);
                    const double sc2 = xmul(scale_xy, scale_xy);
                    // dimer_energy_config (src/MCRun.jl:76-85, src/EnergyEvaluation.jl:397-474) on new_dist2_mat: dist2_mat * scale^2
                    // for the uniform move; recomputed from the trial configuration (get_distance2_mat!) for the separated moves.
                    // Rows are dealt to the lanes of the group.
                    double acce = 0.0;
                    double* D = P.d2mat + ((size_t)lad * T + c) * N * N;
                    const bool thawed = moved[c * MW + MW - 1] >= N;
                    const bool recompute = thawed || ties;
                    if (PSUM && psum) {
                        double* sp = psp + (size_t)c * PTMC_PSUM_K * G + g;
                        if (!psv[c]) {
                            // first volume move of this configuration in the launch: build the sums from the positions (one pass of the
                            // size of the loop this replaces), active cut-off ties (pairs of never-moved atoms in the tie table) left out
                            double sk[PTMC_PSUM_K];
#pragma unroll
                            for (int k = 0; k < PTMC_PSUM_K; ++k) sk[k] = 0.0;
                            for (int a = 0; a < N - 1; ++a) {
                                const double xa = X[a], ya = Y[a], za = Z[a];
                                for (int j = a + 1 + g; j < N; j += G) {
                                    const double d2 = dist2_pbc<rhombic>(bx, xa, ya, za, X[j], Y[j], Z[j]);
                                    psum_acc<true>(rcp_pos(d2), d2 <= rcut, sk);
                                }
                            }
                            if (!thawed) {
                                for (int a = g; a < N; a += G) {
                                    if ((moved[c * MW + (a >> 5)] >> (a & 31)) & 1) continue;
                                    const size_t row = ((size_t)lad * T + c) * N + a;
                                    const int cnt = __ldg(P.tie_cnt + row);
                                    for (int t = 0; t < cnt; ++t) {
                                        const int j = __ldg(P.tie_idx + row * PTMC_TIE_K + t);
                                        if (j < a || ((moved[c * MW + (j >> 5)] >> (j & 31)) & 1)) continue;
                                        const double d2r = dist2_pbc<rhombic>(bx, X[a], Y[a], Z[a], X[j], Y[j], Z[j]);
                                        psum_acc<false>(rcp_pos(d2r), d2r <= rcut, sk);
                                    }
                                }
                            }
#pragma unroll
                            for (int k = 0; k < PTMC_PSUM_K; ++k) sp[k * G] = sk[k];
                            __syncwarp(gmask);
                            if (g == 0) psv[c] = 1;
                            __syncwarp(gmask);
                        }
                        // dimer_energy_config on dist2_mat * scale^2 below the scaled cut-off: sum_k c_k (1 / scale^2)^(3 + k) S_k
                        const double r = xdiv(1.0, sc2);
                        double e_ = P.coeff[PTMC_PSUM_K - 1] * sp[(PTMC_PSUM_K - 1) * G];
#pragma unroll
                        for (int k = PTMC_PSUM_K - 2; k >= 0; --k) e_ = fma(e_, r, P.coeff[k] * sp[k * G]);
                        acce = e_ * (r * r * r);
                        if (!thawed) {
                            // active ties: new_dist2_mat entry = frozen dist2_mat entry * scale^2, decided against the new cut-off
                            for (int a = g; a < N; a += G) {
                                if ((moved[c * MW + (a >> 5)] >> (a & 31)) & 1) continue;
                                const size_t row = ((size_t)lad * T + c) * N + a;
                                const int cnt = __ldg(P.tie_cnt + row);
                                for (int t = 0; t < cnt; ++t) {
                                    const int j = __ldg(P.tie_idx + row * PTMC_TIE_K + t);
                                    if (j < a || ((moved[c * MW + (j >> 5)] >> (j & 31)) & 1)) continue;
                                    const double d2 = xmul(__ldcg(P.tie_val + row * PTMC_TIE_K + t), sc2);
                                    if (d2 <= new_rcut) acce += pair_energy<PTMC_POT_ELJ_EVEN, NC>(P, d2);
                                }
                            }
                        }
                    } else if (!sepv) {
                        for (int a = 0; a < N - 1; ++a) {
                            const double* Da = D + (size_t)a * N;
                            const double xa = X[a], ya = Y[a], za = Z[a];
                            for (int j = a + 1 + g; j < N; j += G) {
                                double d2o, v0 = 0, v1 = 0, v2 = 0;
                                if (DIRP) d2o = dist2_pbc_vec<rhombic>(bx, xa, ya, za, X[j], Y[j], Z[j], v0, v1, v2);
                                if (!recompute) d2o = __ldcg(Da + j);
                                else if (!DIRP) d2o = dist2_pbc<rhombic>(bx, xa, ya, za, X[j], Y[j], Z[j]);
                                const double d2 = xmul(d2o, sc2);                 // new_dist2_mat[a, j]
                                if (POT == PTMC_POT_ELJ_EVEN) {
                                    // branch-free: some lane of the warp is inside the cut-off for practically every j, so a branch saves nothing
                                    const double e_ = pair_energy<PTMC_POT_ELJ_EVEN, NC>(P, d2);
                                    acce += (d2 <= new_rcut) ? e_ : 0.0;
                                } else
                                if (d2 <= new_rcut) acce += DIRP ? pair_energy_dir<POT, NC>(P, d2, v0, v1, v2, sc2)      // tangents unchanged by a uniform scaling
                                                                 : pair_energy<POT == PTMC_POT_EAM ? PTMC_POT_ELJ_EVEN : POT, NC>(P, d2);
                            }
                        }
                        if (ties && !thawed) {
                            // pairs of never-moved atoms on the cut-off: new_dist2_mat entry = frozen dist2_mat entry * scale^2
                            for (int a = g; a < N; a += G) {
                                if ((moved[c * MW + (a >> 5)] >> (a & 31)) & 1) continue;
                                const size_t row = ((size_t)lad * T + c) * N + a;
                                const int cnt = __ldg(P.tie_cnt + row);
                                for (int t = 0; t < cnt; ++t) {
                                    const int j = __ldg(P.tie_idx + row * PTMC_TIE_K + t);
                                    if (j < a || ((moved[c * MW + (j >> 5)] >> (j & 31)) & 1)) continue;
                                    double v0 = 0, v1 = 0, v2 = 0;
                                    const double d2r = DIRP ? dist2_pbc_vec<rhombic>(bx, X[a], Y[a], Z[a], X[j], Y[j], Z[j], v0, v1, v2)
                                                            : dist2_pbc<rhombic>(bx, X[a], Y[a], Z[a], X[j], Y[j], Z[j]);
                                    auto new_term = [&](double d) {
                                        if (!(d <= new_rcut)) return 0.0;
                                        return DIRP ? pair_energy_dir<POT, NC>(P, d, v0, v1, v2, sc2) : pair_energy<POT == PTMC_POT_EAM ? PTMC_POT_ELJ_EVEN : POT, NC>(P, d);
                                    };
                                    acce += new_term(xmul(__ldcg(P.tie_val + row * PTMC_TIE_K + t), sc2)) - new_term(xmul(d2r, sc2));
                                }
                            }
                        }
                    } else {
                        acce = sep_trial_energy<G, POT, NC, rhombic>(P, X, Y, Z, N, g, scale_xy, scale_z, tL, tH, new_rcut, thawed,
                                                                     fresh + c * FW, DIRP ? trho[c] : 1.0, xyz >= 1);
                    }
                    double e_new = group_sum<G>(acce, gmask);
                    const double l = lrc_pot<POT>(P, new_rcut);
                    if (P.bc_kind == PTMC_BC_CUBIC) e_new = e_new + l;
                    else if (P.bc_kind == PTMC_BC_RHOMBIC) e_new = e_new + l * 3 / 4 * tL / tH;
                    else e_new = (tL < tH) ? e_new + l * tL / tH : e_new + l * (tH * tH) / (tL * tL);
                    // metropolis_condition NPT src/Exchange.jl:58-71
                    de = xsub(e_new, en);
                    const double vc = get_volume(P.bc_kind, tL, tH), vu = get_volume(P.bc_kind, boxL, boxH);
                    const double delta_h = xadd(de, xmul(P.pressure, xsub(vc, vu)));
                    double p = exp(xadd(xmul(-delta_h, beta), xmul((double)P.n_atoms, log(xdiv(vc, vu)))));
                    p = p > 1.0 ? 1.0 : p;
                    const double u = BATCH ? B.template rand<4>(gmask) : S.rand();
                    acc = p >= u;
                    if (DIRP && sepv && xyz >= 1) {
                        // the trial's tangents are the state now, for every pair: rejected -> rho times the true ones, accepted -> the true ones
                        __syncwarp(gmask);
                        if (g == 0) {
                            trho[c] = acc ? 1.0 : xdiv(scale_xy, scale_z);
                            for (int w = 0; w < FW; ++w) fresh[c * FW + w] = 0;
                        }
                        __syncwarp(gmask);
                    }
                    if (acc) {
                        // swap_config_v! src/swap_config.jl:113-140
                        __syncwarp(gmask);
                        for (int j = g; j < N; j += G) { X[j] = xmul(scale_xy, X[j]); Y[j] = xmul(scale_xy, Y[j]); Z[j] = xmul(scale_z, Z[j]); }
                        if (PSUM && psum) {
                            double* sp = psp + (size_t)c * PTMC_PSUM_K * G + g;
                            const double r = xdiv(1.0, sc2);
                            double rk = r * r * r;
#pragma unroll
                            for (int k = 0; k < PTMC_PSUM_K; ++k) { sp[k * G] *= rk; rk *= r; }
                        }
                        if (!thawed) {
                            if (ties) {
                                for (int a = g; a < N; a += G) {                  // dist2_mat .= new_dist2_mat, the entries that are kept
                                    const size_t row = ((size_t)lad * T + c) * N + a;
                                    const int cnt = __ldg(P.tie_cnt + row);
                                    for (int t = 0; t < cnt; ++t) __stcg(P.tie_val + row * PTMC_TIE_K + t, xmul(__ldcg(P.tie_val + row * PTMC_TIE_K + t), sc2));
                                }
                            } else if (!sepv) {
                                for (int k = g; k < N * N; k += G) __stcg(D + k, xmul(__ldcg(D + k), sc2));   // dist2_mat .= new_dist2_mat
                            } else {
                                __syncwarp(gmask);
                                sep_rewrite_matrix<G>(P, X, Y, Z, N, g, tL, tH, D);
                            }
                        }
                        boxL = tL; boxH = tH; rcut = new_rcut; en = e_new;
                        bx = make_box(P.bc_kind, boxL, boxH);
                        if (xyz == 0) { n_acc_vol += 1; n_adj_vol += 1; }
                        else if (g == 0) { sepi[2 * (xyz - 1)] += 1; sepi[2 * (xyz - 1) + 1] += 1; }
                    }
                    __syncwarp(gmask);
                }
                if (tracing) {      // uniform branch; the lanes of a group store identical values to one address
                    size_t o = ((((size_t)lad * P.trace_cap + P.trace_base + cyc) * T + slot) * P.n_steps) + step;
                    P.tr_index[o] = index; P.tr_de[o] = de; P.tr_acc[o] = acc ? 1 : 0;
                }
                if (DBG) { if (S.exhausted) err = PTMC_ERR_REPLAY; }
            }
            if (g == 0) {
                scal[4 * c] = en;
                if (PBC) { scal[4 * c + 1] = boxL; scal[4 * c + 2] = boxH; scal[4 * c + 3] = rcut; }
            }
        }
        // ========================================= exchange: src/MCRun.jl:203-205, src/Exchange.jl:132-186 ==
        __syncwarp();
        int ex_gate = 0, ex_n = 0;
        double ex_u = 0.0;
        if (phase == 0 && (tid & 31) == 0) {
            Stream<DBG> LS;
            LS.r.s0 = wl[0]; LS.r.s1 = wl[1]; LS.r.s2 = wl[2]; LS.r.s3 = wl[3];
            LS.draws = nullptr; LS.pos = 0; LS.len = 0; LS.exhausted = 0;
            if (DBG) {
                if (P.replay) { LS.draws = P.ladder_draws + lad * P.ladder_len; LS.len = P.ladder_len; LS.pos = (i64)wl[4]; }
            }
            if (LS.rand() < 0.1) {                                  // :203
                ex_gate = 1;
                ex_n = LS.range((u64)(T - 1));                      // Exchange.jl:135
                ex_u = LS.rand();                                   // :145 (drawn whenever an exchange is attempted)
            }
            wl[0] = LS.r.s0; wl[1] = LS.r.s1; wl[2] = LS.r.s2; wl[3] = LS.r.s3;
            if (DBG) {
                wl[4] = (u64)LS.pos;
                if (LS.exhausted) err = PTMC_ERR_REPLAY;
            }
        }
        ex_gate = __shfl_sync(0xffffffffu, ex_gate, 0);
        if (ex_gate) {
            ex_n = __shfl_sync(0xffffffffu, ex_n, 0);
            ex_u = __shfl_sync(0xffffffffu, ex_u, 0);
            // Only the two slots of the attempt have to meet: every warp steps its own copy of the ladder stream, so the other
            // warps run on (replicas differ a lot in work per cycle: volume moves, out-of-sphere redraws).  Slots in two warps
            // meet on the mbarrier of the lower warp; its phase parity is tracked per side in xpar.
            const int a = ex_n - 1, b = ex_n;
            const int wa = (a * G) >> 5, wb = (b * G) >> 5, wme = tid >> 5;
            const bool meet = wa != wb && (wme == wa || wme == wb);
            u64* xbar = lrng + 6 * wa + 5;
            const int side = wme == wa ? 0 : 1;
            if (meet) { pair_barrier(xbar, (xpar >> side) & 1u); xpar ^= 1u << side; }   // both slots have finished this cycle's moves
            else __syncwarp();
            if (tid == a * G) {
                int ex_acc = 0;
                exc_cnt[2 * a] += 1; exc_cnt[2 * b] += 1;
                const int ca = perm[a], cb = perm[b];
                double ea = scal[4 * ca], eb = scal[4 * cb];
                if (npt) {
                    ea = xadd(ea, xmul(P.pressure, get_volume(P.bc_kind, scal[4 * ca + 1], scal[4 * ca + 2])));
                    eb = xadd(eb, xmul(P.pressure, get_volume(P.bc_kind, scal[4 * cb + 1], scal[4 * cb + 2])));
                }
                // exc_acceptance :103-108: min(1, exp(dbeta * dE)) > rand()   (strict)
                double pe = exp(xmul(xsub(P.betas[a], P.betas[b]), xsub(ea, eb)));
                pe = (pe != pe) ? pe : (pe < 1.0 ? pe : 1.0);
                if (pe > ex_u) {
                    ex_acc = 1;
                    exc_cnt[2 * a + 1] += 1; exc_cnt[2 * b + 1] += 1;
                    perm[a] = cb; perm[b] = ca;       // exc_trajectories! :115-124
                }
                if (tracing) {
                    int32_t* e = P.tr_exc + ((size_t)lad * P.trace_cap + P.trace_base + cyc) * 3;
                    e[0] = 1; e[1] = ex_n; e[2] = ex_acc;
                }
            }
            if (meet) { pair_barrier(xbar, (xpar >> side) & 1u); xpar ^= 1u << side; }
            else __syncwarp();
        } else {
            __syncwarp(hmask);                                      // scal[] of this slot is visible to its group
        }
        if (active && phase != 1) {
            // ------------------------------------------- update_max_stepsize! src/Exchange.jl:198-263
            if (adj_phase == 0) {
                double rate = xdiv((double)n_adj_atom, (double)((i64)P.n_adjust * P.n_atom_moves));
                if (rate < P.min_acc) dmax = xmul(dmax, 0.9);
                else if (rate > P.max_acc) dmax = xmul(dmax, 1.1);
                dmax53 = dmax * 0x1.0p-53;
                n_adj_atom = 0;
                if (PBC) {
                    if (npt && !P.separated) {
                        rate = xdiv((double)n_adj_vol, (double)((i64)P.n_adjust * P.n_volume_moves));
                        if (rate < P.min_acc) vmax = xmul(vmax, 0.9);
                        else if (rate > P.max_acc) vmax = xmul(vmax, 1.1);
                        n_adj_vol = 0;
                    } else if (npt) {
                        // src/Exchange.jl:221-245: half of the volume moves are uniform, a third xy, a sixth z
                        const double nv = (double)((i64)P.n_adjust * P.n_volume_moves);
                        double* sepd = sm + lay.off_sep + 6 * slot; int* sepi = reinterpret_cast<int*>(sepd + 3);
                        rate = xdiv((double)n_adj_vol, xdiv(nv, 2.0));
                        if (rate < P.min_acc) vmax = xmul(vmax, 0.9);
                        else if (rate > P.max_acc) vmax = xmul(vmax, 1.1);
                        n_adj_vol = 0;
                        __syncwarp(gmask);
                        if (g == 0) {
                            rate = xdiv((double)sepi[1], xdiv(nv, 3.0));
                            if (rate < P.min_acc) sepd[0] = xmul(sepd[0], 0.9);
                            else if (rate > P.max_acc) sepd[0] = xmul(sepd[0], 1.1);
                            sepi[1] = 0;
                            rate = xdiv((double)sepi[3], xdiv(nv, 6.0));
                            if (rate < P.min_acc) sepd[1] = xmul(sepd[1], 0.9);
                            else if (rate > P.max_acc) sepd[1] = xmul(sepd[1], 1.1);
                            sepi[3] = 0;
                        }
                        __syncwarp(gmask);
                    }
                }
            }
            // ---------------------------------------------------- sampling_step! src/Sampling.jl:330-371
            if (sample && smp_phase == 0 && g == 0) {
                const int c = perm[slot];
                const double e = scal[4 * c];
                const int nb = P.n_bin;
                const double en_min = P.en_min[lad];
                const double delta_en = xdiv(xsub(P.en_max[lad], en_min), (double)(nb - 1));   // :150
                if (npt) {
                    const double L = scal[4 * c + 1], h = scal[4 * c + 2];
                    const double hv = xadd(e, xmul(P.pressure, get_volume(P.bc_kind, L, h)));   // :35-46
                    hams[2 * slot] = xadd(hams[2 * slot], hv);
                    hams[2 * slot + 1] = xadd(hams[2 * slot + 1], xmul(hv, hv));
                    const double v_min = P.v_min[lad];
                    const double delta_v = xdiv(xsub(P.v_max[lad], v_min), (double)nb);
                    double vol_h;   // find_hist_index :72-84, :98-113, :115-130 (rhombic quirk: 3*sqrt(h)/2)
                    if (P.bc_kind == PTMC_BC_CUBIC) vol_h = xmul(xmul(L, L), L);
                    else if (P.bc_kind == PTMC_BC_RECTANGULAR) vol_h = xmul(xmul(L, L), h);
                    else vol_h = xdiv(xmul(xmul(L, L), xmul(3.0, sqrt(h))), 2.0);
                    double fe = floor(xdiv(xsub(e, en_min), delta_en)) + 1.0;
                    double fv = floor(xdiv(xsub(vol_h, v_min), delta_v)) + 1.0;
                    int he = fe < 1.0 ? 1 : (fe > (double)nb ? nb + 2 : (int)fe + 1);
                    int hv_ = fv < 1.0 ? 1 : (fv > (double)nb ? nb + 2 : (int)fv + 1);
                    if (fe != fe) he = nb + 2;   // NaN: Julia's floor(Int, NaN) throws; park it in the overflow bin
                    if (fv != fv) hv_ = nb + 2;
                    P.hist[((size_t)rep * (nb + 2) + (he - 1)) * (nb + 2) + (hv_ - 1)] += 1;
                } else {
                    hams[2 * slot] = xadd(hams[2 * slot], e);                                  // :29-34
                    hams[2 * slot + 1] = xadd(hams[2 * slot + 1], xmul(e, e));
                    const double hi = xadd(xdiv(xsub(e, en_min), delta_en), 1.0);             // :63-66
                    int h = hi < 1.0 ? 1 : (hi > (double)nb ? nb + 2 : (int)floor(hi) + 1);     // :53-62
                    if (hi != hi) h = nb + 2;
                    P.hist[(size_t)rep * (nb + 2) + (h - 1)] += 1;
                }
            }
        }
        if (++adj_phase == P.n_adjust) adj_phase = 0;
        if (++smp_phase == P.mc_sample) smp_phase = 0;
    }

    // ---- store: configurations go back in slot order ----------------------------------------------------
    __syncthreads();
    if (active) {
        const int c = perm[slot];
        const double* X = sm + (size_t)c * lay.cfg_stride;
        const double* Y = X + NP; const double* Z = Y + NP;
        double* gp = P.pos + rep * 3 * N;
        for (int j = g; j < N; j += G) {
            gp[3 * j] = X[j]; gp[3 * j + 1] = Y[j]; gp[3 * j + 2] = Z[j];
            if (EAM) { P.comp[(rep * N + j) * 2] = Z[NP + j]; P.comp[(rep * N + j) * 2 + 1] = Z[2 * NP + j]; }
        }
        if (g == 0) {
            P.en_tot[rep] = scal[4 * c];
            P.max_displ[rep * 4] = dmax;
            P.ham[rep * 2] = hams[2 * slot]; P.ham[rep * 2 + 1] = hams[2 * slot + 1];
            i64* cn = P.counters + rep * PTMC_NCOUNT;
            cn[PTMC_C_ATOM_TOT] += n_acc_atom; cn[PTMC_C_ATOM_ADJ] = n_adj_atom;
            if (PBC) {
                P.perm_g[rep] = c;
                for (int w = 0; w < MW; ++w) P.moved_g[((i64)lad * T + c) * MW + w] = moved[c * MW + w];
                if (DIRP && P.separated) {
                    P.tan_rho[(i64)lad * T + c] = trho[c];
                    for (int w = 0; w < P.moved_words; ++w) P.fresh_g[((i64)lad * T + c) * P.moved_words + w] = fresh[c * FW + w];
                }
                P.box[rep * 2] = scal[4 * c + 1]; P.box[rep * 2 + 1] = scal[4 * c + 2];
                P.max_displ[rep * 4 + 1] = vmax;
                cn[PTMC_C_VOL_TOT] += n_acc_vol; cn[PTMC_C_VOL_ADJ] = n_adj_vol;
                const double* sd = sm + lay.off_sep + 6 * slot; const int* si = reinterpret_cast<const int*>(sd + 3);
                P.max_displ[rep * 4 + 2] = sd[0]; P.max_displ[rep * 4 + 3] = sd[1];
                cn[PTMC_C_VOLXY_TOT] += si[0]; cn[PTMC_C_VOLXY_ADJ] = si[1];
                cn[PTMC_C_VOLZ_TOT] += si[2]; cn[PTMC_C_VOLZ_ADJ] = si[3];
            }
            cn[PTMC_C_EXC_ATT] += exc_cnt[2 * slot]; cn[PTMC_C_EXC_ACC] += exc_cnt[2 * slot + 1];
            if (!BATCH) {
                P.slot_rng[rep * 4] = S.r.s0; P.slot_rng[rep * 4 + 1] = S.r.s1;
                P.slot_rng[rep * 4 + 2] = S.r.s2; P.slot_rng[rep * 4 + 3] = S.r.s3;
                if (P.replay) P.slot_cursor[rep] = S.pos;
            }
        }
        if (BATCH) B.store(P.slot_rng + rep * (4 * PTMC_RNG_LANES), g);
    }
    if (tid == 0) {
        P.ladder_rng[lad * 4] = wl[0]; P.ladder_rng[lad * 4 + 1] = wl[1];
        P.ladder_rng[lad * 4 + 2] = wl[2]; P.ladder_rng[lad * 4 + 3] = wl[3];
        if (DBG) { if (P.replay) P.ladder_cursor[lad] = (i64)wl[4]; }
    }
    if (err) atomicMin(P.error, err);
}

template <int G, int POT, int GEOM, bool DBG, int NITER, int NC, bool FULLW = false, bool ROT = false>
__global__ void __launch_bounds__(sweep_max_threads_c(G), sweep_min_blocks_c(G))
sweep_kernel(const DevParams P, const i64 first_idx, const i64 n_cycles, const int sample) {
    sweep_body<G, POT, GEOM, DBG, NITER, NC, FULLW, ROT>(P, first_idx, n_cycles, sample);
}
// Small ladders on 4 lanes per replica (at most 128 threads): 8 replicas share a warp, so the per-move work that every lane repeats
// (RNG, trial position, Metropolis, commit) is paid once per 8 moves instead of once per 4 (-15.5 % warp instructions for Ne55).  Shared
// memory, not registers, bounds the residency there (6 ladders of 24 x Ne55 per SM = 18 warps), so the kernel takes the registers that
// are free anyway (96 instead of 80) to keep more of the 14 unrolled pair iterations in flight: C2 170.9 -> 155.5 ms per step
// (profiles/r02_experiments.md; caps of 112 / 128 registers: 156.0 / 157.0 ms).
#ifndef PTMC_G4S_MAXREG
#define PTMC_G4S_MAXREG 96
#endif
template <int POT, bool DBG, int NITER, int NC, bool FULLW, bool ROT>
__global__ void __maxnreg__(PTMC_G4S_MAXREG)
sweep_kernel_g4s(const DevParams P, const i64 first_idx, const i64 n_cycles, const int sample) {
    sweep_body<4, POT, GEOM_SPHERE, DBG, NITER, NC, FULLW, ROT>(P, first_idx, n_cycles, sample);
}
// Periodic ELJEven{6} on 16 lanes per replica (two replicas per warp, 256-thread ladders of 16 temperatures, two ladders per SM at
// 128 registers): worth it only since uniform volume moves are O(1) (power sums) - before, a volume move of one replica stalled
// its warp mate for 80 atom moves (profiles/r02_experiments.md)
template <int GEOM, int NITER>
__global__ void __maxnreg__(128)
sweep_kernel_p16(const DevParams P, const i64 first_idx, const i64 n_cycles, const int sample) {
    sweep_body<16, PTMC_POT_ELJ_EVEN, GEOM, false, NITER, 6>(P, first_idx, n_cycles, sample);
}
// EmbeddedAtomPotential: the unrolled loop wants ~170 registers; capped so that two 224-thread ladders (scripts/cu55.jl:
// 28 temperatures x 8 lanes) stay resident per SM (a 96-register cap for three ladders spills and measured slower)
#ifndef PTMC_EAM_MAXREG
#define PTMC_EAM_MAXREG 128
#endif
template <int G, bool DBG, int NITER, bool FULLW = false>
__global__ void __maxnreg__(PTMC_EAM_MAXREG)
sweep_kernel_eam(const DevParams P, const i64 first_idx, const i64 n_cycles, const int sample) {
    sweep_body<G, PTMC_POT_EAM, GEOM_SPHERE, DBG, NITER, 0, FULLW>(P, first_idx, n_cycles, sample);
}

// ------------------------------------------------------------------------------------------- init --
// MCState constructor (src/MCStates.jl:137-162): initial en_tot (dimer_energy_config
// src/EnergyEvaluation.jl:378-474, or EAM set_variables :1800-1811 + :1722-1734) for every replica.
// One warp per replica.
template <int POT>
__global__ void init_kernel(const DevParams P, const i64 n_rep) {
    const i64 rep = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int g = threadIdx.x & 31;
    if (rep >= n_rep) return;
    const int N = P.n_atoms;
    const double* gp = P.pos + rep * 3 * N;
    const double L = P.box[rep * 2], h = P.box[rep * 2 + 1];
    const bool pbc = P.bc_kind != PTMC_BC_SPHERICAL;
    const bool npt = P.ensemble == PTMC_ENS_NPT;
    const double rcut = npt ? get_r_cut(P.bc_kind, L, h) : 0.0;
    if (POT == PTMC_POT_EAM) {
        double se = 0.0;
        for (int a = g; a < N; a += 32) {
            double sp = 0.0, sr = 0.0;
            for (int j = 0; j < N; ++j) {
                if (j == a) continue;
                double dx = gp[3 * a] - gp[3 * j], dy = gp[3 * a + 1] - gp[3 * j + 1], dz = gp[3 * a + 2] - gp[3 * j + 2];
                double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                double p, r; eam_invrexp(P, P.eam_ktab, P.eam_etab, d2, p, r);
                sp += p; sr += r;
            }
            P.comp[(rep * N + a) * 2] = sp; P.comp[(rep * N + a) * 2 + 1] = sr;
            se += fma(-P.eam_ecam, sqrt(sr), P.eam_ean * sp);
        }
        se = group_sum<32>(se, 0xffffffffu);
        if (g == 0) P.en_tot[rep] = se;
    } else {
        double acc = 0.0;
        const bool ties = pbc && P.use_ties != 0;
        double* D = (pbc && !ties) ? P.d2mat + (size_t)rep * N * N : nullptr;   // perm is the identity right after an upload
        for (int a = 0; a < N - 1; ++a) {
            const double xa = gp[3 * a], ya = gp[3 * a + 1], za = gp[3 * a + 2];
            for (int j = a + 1 + g; j < N; j += 32) {
                double d2;
                if (pbc) {
                    // get_distance2_mat! src/Configurations.jl:146-162: dest[i,j] = dest[j,i] = distance2(c[i], c[j], bc)
                    d2 = dist2_exact(P.bc_kind, xa, ya, za, gp[3 * j], gp[3 * j + 1], gp[3 * j + 2], L, h);
                    if (!ties) { D[(size_t)a * N + j] = d2; D[(size_t)j * N + a] = d2; }
                    if (POT == PTMC_POT_ELJ_B || POT == PTMC_POT_LUT) {
                        const Box b = make_box(P.bc_kind, L, h);
                        double v0, v1, v2;
                        if (P.bc_kind == PTMC_BC_RHOMBIC) (void)dist2_pbc_vec<true>(b, xa, ya, za, gp[3 * j], gp[3 * j + 1], gp[3 * j + 2], v0, v1, v2);
                        else (void)dist2_pbc_vec<false>(b, xa, ya, za, gp[3 * j], gp[3 * j + 1], gp[3 * j + 2], v0, v1, v2);
                        if (!npt || d2 <= rcut) acc += pair_energy_dir<POT, 0>(P, d2, v0, v1, v2);
                        continue;
                    }
                } else {
                    double dx = xa - gp[3 * j], dy = ya - gp[3 * j + 1], dz = za - gp[3 * j + 2];
                    d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (POT == PTMC_POT_ELJ_B) { acc += eljb_energy(P, d2, dz * dz); continue; }
                    if (POT == PTMC_POT_LUT) { acc += lut_energy(P, d2, dx, dy, dz); continue; }
                }
                if (!npt || d2 <= rcut) acc += pair_energy<POT == PTMC_POT_EAM ? PTMC_POT_ELJ_EVEN : POT>(P, d2);
            }
        }
        if (pbc && !ties) for (int a = g; a < N; a += 32) D[(size_t)a * N + a] = 0.0;
        if (ties) {
            // tie table: per atom, the partners whose dist2_mat entry (dest[i,j] = dest[j,i] = distance2(c[i], c[j]), i < j) sits on the cut-off
            int total = 0;
            for (int a = g; a < N; a += 32) {
                const size_t row = (size_t)rep * N + a;
                int cnt = 0;
                for (int j = 0; j < N; ++j) {
                    if (j == a) continue;
                    const int lo = a < j ? a : j, hi = a < j ? j : a;
                    const double d2 = dist2_exact(P.bc_kind, gp[3 * lo], gp[3 * lo + 1], gp[3 * lo + 2], gp[3 * hi], gp[3 * hi + 1], gp[3 * hi + 2], L, h);
                    if (fabs(d2 - rcut) <= PTMC_TIE_REL * rcut) {
                        if (cnt < PTMC_TIE_K) { P.tie_idx[row * PTMC_TIE_K + cnt] = (unsigned char)j; P.tie_val[row * PTMC_TIE_K + cnt] = d2; }
                        ++cnt;
                    }
                }
                if (cnt > PTMC_TIE_K) { atomicMin(P.error, (int)PTMC_ERR_TIES); cnt = PTMC_TIE_K; }
                P.tie_cnt[row] = cnt;
                total += cnt;
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
            if (total == 0 && g == 0) {
                // nothing on the cut-off: the configuration starts thawed
                int* mv = P.moved_g + rep * (P.moved_words + 1);
                for (int w = 0; w < P.moved_words; ++w) {
                    const int nb = N - 32 * w;
                    mv[w] = nb >= 32 ? -1 : (int)((1u << nb) - 1u);
                }
                mv[P.moved_words] = N;
            }
        }
        double e = group_sum<32>(acc, 0xffffffffu);
        if (npt) {
            const double l = lrc_pot<POT>(P, rcut);
            if (P.bc_kind == PTMC_BC_CUBIC) e = e + l;
            else if (P.bc_kind == PTMC_BC_RHOMBIC) e = e + l * 3 / 4 * L / h;
            else e = (L < h) ? e + l * L / h : e + l * (h * h) / (L * L);
        }
        if (g == 0) P.en_tot[rep] = e;
        if (g == 0 && pbc && (POT == PTMC_POT_ELJ_B || POT == PTMC_POT_LUT) && P.separated) {
            P.tan_rho[rep] = 1.0;                     // tan_mat = get_tantheta_mat(config): nothing stale
            for (int w = 0; w < P.moved_words; ++w) P.fresh_g[rep * P.moved_words + w] = 0;
        }
    }
}

// Host-visible launchers (one per instantiation file).
typedef void (*sweep_fn)(const DevParams, const i64, const i64, const int);

}  // namespace ptmc
