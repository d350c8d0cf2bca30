/*
 * ptmc_b200.h — C ABI of libptmc_b200.so: the B200-native (sm_100a) sweep hot path of
 * ParallelTemperingMonteCarlo.jl.
 *
 * The reference has no FFI for the sweep (extension is by Julia multiple dispatch); its only FFI is
 * `forward` (MachineLearningPotential/ForwardPass.jl:60-83).  This header is therefore the boundary a
 * maintainer binds with `ccall` at mc_step!/mc_cycle! (src/MCRun.jl:169,192,215): "run K cycles on a
 * batch of ladders".  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C, extern "C", host pointers unless a name says `dev`; caller owns every host buffer.
 *   - every function returns 0 on success or a negative ptmc_status; ptmc_last_error() gives the text.
 *     Nothing aborts and no exception crosses the boundary.
 *   - one handle = one CUDA device, used from one host thread at a time.
 *   - there is NO CPU fallback: ptmc_create fails (PTMC_ERR_CUDA) when no sm_100 device is usable.
 *   - array layouts are row-major ("C") with the ladder index slowest:
 *       pos      [n_ladders][n_traj][n_atoms][3]   (slot = temperature index of the ladder)
 *       box      [n_ladders][n_traj][2]            spherical: {radius^2, unused}; cubic: {L, L};
 *                                                  rhombic/rectangular: {box_length, box_height}
 *       max_displ[n_ladders][n_traj][4]            MCState.max_displ (src/MCStates.jl:43,104)
 *       counters [n_ladders][n_traj][PTMC_NCOUNT]  see ptmc_counter
 *       ham      [n_ladders][n_traj][2]            MCState.ham  (sum E, sum E^2; NPT: enthalpy)
 *   - indices that the reference keeps 1-based (atom index, n_exc) are 1-based in traces too.
 */
#ifndef PTMC_B200_H
#define PTMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Every entry point is exported explicitly; the library is built with -fvisibility=hidden. */
#if defined(__GNUC__)
#define PTMC_API __attribute__((visibility("default")))
#else
#define PTMC_API
#endif

#define PTMC_ABI_VERSION 4
#define PTMC_MAX_COEFF 16
#define PTMC_NCOUNT 10

typedef enum ptmc_status {
    PTMC_OK = 0,
    PTMC_ERR_ARG = -1,         /* bad argument / unsupported configuration                       */
    PTMC_ERR_CUDA = -2,        /* CUDA runtime error (text in ptmc_last_error)                   */
    PTMC_ERR_STATE = -3,       /* call order (e.g. run before upload)                            */
    PTMC_ERR_SPHERE = -4,      /* "too many moves out of binding sphere" (src/MCMoves.jl:132)    */
    PTMC_ERR_REPLAY = -5,      /* replay stream exhausted                                        */
    PTMC_ERR_TIES = -6         /* periodic NPT: more than 32 neighbours of one atom sit on the cut-off sphere at upload */
} ptmc_status;

/* src/EnergyEvaluation.jl:891 (ELJPotential), :138 (ELJPotentialEven), :1059 (EmbeddedAtomPotential),
 * :1182 (RuNNerPotential) */
/* :909-921 (ELJPotentialB: extended Lennard-Jones in a magnetic field, pair energy depends on the angle to the z axis) */
typedef enum ptmc_pot_kind { PTMC_POT_ELJ = 0, PTMC_POT_ELJ_EVEN = 1, PTMC_POT_EAM = 2, PTMC_POT_RUNNER = 3, PTMC_POT_ELJ_B = 4,
                             PTMC_POT_LUT = 5 /* :214-295 LookupTablePotential: energy tabulated over (angle to z, distance) */,
                             PTMC_POT_RUNNER2 = 6 /* :1287-1314 RuNNerPotential2Atom: two species, one network per species */ } ptmc_pot_kind;
/* src/BoundaryConditions.jl:44,72,89,110 */
typedef enum ptmc_bc_kind { PTMC_BC_SPHERICAL = 0, PTMC_BC_CUBIC = 1, PTMC_BC_RHOMBIC = 2, PTMC_BC_RECTANGULAR = 3 } ptmc_bc_kind;
/* src/Ensembles.jl:42 (NVT), :75 (NPT; separated_volume in ptmc_config), :138 (NNVT: two species, atom moves + atom swaps) */
typedef enum ptmc_ens_kind { PTMC_ENS_NVT = 0, PTMC_ENS_NPT = 1, PTMC_ENS_NNVT = 2 } ptmc_ens_kind;
typedef enum ptmc_rng_mode {
    PTMC_RNG_XOSHIRO = 0, /* Julia's default RNG as ONE sequential stream per slot: Xoshiro256++,
                             rand()=(u>>11)*2^-53, rand(1:n)=Lemire.  Compatibility mode (slower kernel).  */
    PTMC_RNG_REPLAY = 1,  /* consume pre-drawn raw UInt64 draws from ptmc_set_replay (same derivations)   */
    PTMC_RNG_BATCH8 = 2   /* production mode: every slot owns PTMC_RNG_LANES Xoshiro256++ substreams that
                             step together once per move ("batch"); position k of the batch comes from
                             substream k: 0 index (1 + mulhi(u, n_steps)), 1-3 displacement x,y,z (volume move: 1 scale
                             factor, 2 the rand(1:6) of volume_change_separated as 1 + mulhi(u, 6)), 4 acceptance, 5-7 first out-of-sphere redraw; every further redraw steps a
                             new batch and uses positions 0-2 (the acceptance uniform is position 4 of the
                             latest batch).  Same generator and same uniform construction as mode 0.          */
} ptmc_rng_mode;
#define PTMC_RNG_LANES 8

/* MCState counters (src/MCStates.jl:48-52), flattened in this order. */
typedef enum ptmc_counter {
    PTMC_C_ATOM_TOT = 0, PTMC_C_ATOM_ADJ = 1,   /* count_atom[1], [2] */
    PTMC_C_VOL_TOT = 2, PTMC_C_VOL_ADJ = 3,     /* count_vol[1], [2]  */
    PTMC_C_VOLXY_TOT = 4, PTMC_C_VOLXY_ADJ = 5, /* count_vol_xy (separated volume moves)                 */
    PTMC_C_VOLZ_TOT = 6, PTMC_C_VOLZ_ADJ = 7,   /* count_vol_z                                         */
    PTMC_C_EXC_ATT = 8, PTMC_C_EXC_ACC = 9      /* count_exc[1] (attempts), [2] (accepted)             */
} ptmc_counter;

/* RuNNer single-species potential (src/EnergyEvaluation.jl:1182-1218,
 * MachineLearningPotential/SymmFunc.jl:24-44,86-112, ForwardPass.jl:23-45). */
typedef struct ptmc_runner_desc {
    int32_t n_rad, n_ang;          /* Nrad, Nang (10, 78 for scripts/cu55.jl)                        */
    const double* rad_eta;         /* [n_rad]                                                         */
    const int32_t* rad_type;       /* [n_rad] type_vec as integer (11 is the only evaluated one)      */
    const double* rad_gmax_gmin;   /* [n_rad][2] {G_max, G_min} from scaling.data                     */
    const double* ang_eta;         /* [n_ang]                                                         */
    const double* ang_lambda;      /* [n_ang]                                                         */
    const double* ang_zeta;        /* [n_ang]                                                         */
    const int32_t* ang_type;       /* [n_ang] (111 is the only evaluated one)                         */
    const double* ang_gmax_gmin;   /* [n_ang][2]                                                      */
    double r_cut;                  /* pot.r_cut                                                       */
    double boundary2;              /* pot.boundary (already squared, EnergyEvaluation.jl:1216)        */
    int32_t n_layers;              /* nnp.n_layers                                                    */
    const int32_t* num_nodes;      /* [n_layers]                                                      */
    const int32_t* activation;     /* [n_layers] 1 = linear, 2 = tanh                                 */
    int32_t n_params;
    const double* params;          /* weights.XXX.data order                         */
} ptmc_runner_desc;

/* RuNNer two-species potential (RuNNerPotential2Atom src/EnergyEvaluation.jl:1287-1314; RadialType2a / AngularType3a
 * MachineLearningPotential/SymmFunc.jl:53-84,113-171): atoms 0 .. n1-1 are species 1, n1 .. n1+n2-1 species 2.  Every radial
 * function owns two rows of the g matrix (neighbour of species 1 / 2), every angular function three (0 / 1 / 2 neighbours of
 * species 2): n_sym = 2 n_rad + 3 n_ang rows, radial function k at rows 2k, 2k+1, angular function k at 2 n_rad + 3k + {0,1,2}
 * (total_symm_calc SymmFunc.jl:511-545).  All functions are evaluated (no type filter, DeltaMatrix.jl:256-316). */
typedef struct ptmc_runner2_desc {
    int32_t n_rad, n_ang;          /* Nrad, Nang (5, 26 for scripts/brass55.jl)                       */
    int32_t n1, n2;                /* N1, N2 (50, 5)                                                  */
    const double* rad_eta;         /* [n_rad]                                                         */
    const double* ang_eta;         /* [n_ang]                                                         */
    const double* ang_lambda;      /* [n_ang]                                                         */
    const double* ang_zeta;        /* [n_ang]                                                         */
    const double* gmax_gmin_1;     /* [n_sym][2] {G_max, G_min} of species 1 in g-matrix row order (scaling.data rows 1-88)  */
    const double* gmax_gmin_2;     /* [n_sym][2] the same for species 2 (rows 89-176)                 */
    double r_cut;                  /* pot.r_cut                                                       */
    double boundary2;              /* pot.boundary (squared, :1305-1314; energy_update! :1623-1643 does not use it) */
    int32_t n_layers;              /* both networks share the topology (scripts/brass55.jl:151-164)   */
    const int32_t* num_nodes;      /* [n_layers]                                                      */
    const int32_t* activation;     /* [n_layers]                                                      */
    int32_t n_params;
    const double* params_1;        /* nnp1 (weights.029.data order)                       */
    const double* params_2;        /* nnp2                                                            */
} ptmc_runner2_desc;

/* Everything that is fixed for the lifetime of a handle.  Mirrors what ptmc_run! receives:
 * MCParams (src/InputParams.jl:35), TempGrid (:77), the potential, the ensemble and the BC kind. */
typedef struct ptmc_config {
    int32_t abi_version;      /* = PTMC_ABI_VERSION                                                  */
    int32_t device;           /* CUDA device ordinal                                                 */
    int64_t n_ladders;        /* independent PT ladders batched on this device                       */
    int32_t n_traj;           /* MCParams.n_traj                                                     */
    int32_t n_atoms;          /* MCParams.n_atoms                                                    */
    int32_t ensemble;         /* ptmc_ens_kind                                                       */
    int32_t n_atom_moves;     /* NVT/NPT.n_atom_moves                                                */
    int32_t n_volume_moves;   /* NPT.n_volume_moves (0 for NVT)                                      */
    int32_t bc_kind;          /* ptmc_bc_kind                                                        */
    double pressure;          /* NPT.pressure                                                        */
    int32_t pot_kind;         /* ptmc_pot_kind                                                       */
    int32_t n_coeff;          /* ELJ: number of coefficients (<= PTMC_MAX_COEFF)                     */
    double coeff[PTMC_MAX_COEFF];
    double eam_n, eam_m, eam_ean, eam_ecam;  /* EmbeddedAtomPotential fields                         */
    const ptmc_runner_desc* runner;          /* PTMC_POT_RUNNER only                                 */
    const double* betas;      /* [n_traj] TempGrid.beta_grid                                         */
    int32_t n_adjust;         /* MCParams.n_adjust                                                   */
    int32_t mc_sample;        /* MCParams.mc_sample                                                  */
    int32_t n_bin;            /* MCParams.n_bin                                                      */
    int32_t rng_mode;         /* ptmc_rng_mode                                                       */
    double min_acc, max_acc;  /* MCParams.min_acc / max_acc                                          */
    uint64_t seed;            /* job seed: ptmc_create seeds with ptmc_seed_job(h, seed, 0)           */
    int32_t lanes_per_replica;/* 0 = auto; else 2 (free-space ELJEven only), 4, 8, 16, 32 lanes cooperating on one replica */
    int32_t separated_volume; /* NPT.separated_volume (src/Ensembles.jl:81): volume_change_separated, RhombicBC / RectangularBC */
    /* ABI 2: PTMC_POT_ELJ_B only - coeff_a, coeff_b [n_coeff] (coeff_c travels in coeff[])                */
    const double* eljb_a;
    const double* eljb_b;
    /* ABI 3: PTMC_POT_LUT only - LookupTablePotential fields (src/EnergyEvaluation.jl:214-223); table[l_dist][l_angle]
     * in the file's order of the ANGLE index slowest is table[(i_angle - 1) * l_dist + (j_dist - 1)]                     */
    const double* lut_table;
    int32_t lut_l_dist, lut_l_angle;
    double lut_start_dist, lut_d_dist, lut_start_angle, lut_d_angle, lut_c6coeff;
    /* ABI 4: PTMC_ENS_NNVT - NNVT.n_atom_swaps (src/Ensembles.jl:138-151): move indices n_atom_moves + 1 .. n_atom_moves +
     * n_atom_swaps of MoveStrategy(::NNVT) (:270-282) are "atomswap"; PTMC_POT_RUNNER2 - the two-species potential       */
    int32_t n_atom_swaps;
    const ptmc_runner2_desc* runner2;
} ptmc_config;

typedef struct ptmc_handle ptmc_handle;

/* ---- lifetime ------------------------------------------------------------------------------------ */
PTMC_API int ptmc_abi_version(void);
/* Text of the last error on this host thread (handle may be NULL for create-time errors). */
PTMC_API const char* ptmc_last_error(const ptmc_handle* h);
/* Replaces `initialisation` (src/Initialization.jl:79-101) for the device side: allocates all
 * per-replica state for n_ladders x n_traj replicas. */
PTMC_API int ptmc_create(const ptmc_config* cfg, ptmc_handle** out);
PTMC_API void ptmc_destroy(ptmc_handle* h);
/* Launch all work of this handle on an existing cudaStream_t (e.g. torch's current stream). */
PTMC_API int ptmc_set_stream(ptmc_handle* h, void* cuda_stream);

/* ---- state in ------------------------------------------------------------------------------------ */
/* MCState constructor (src/MCStates.jl:137-162): uploads positions/box, recomputes on the device every
 * derived quantity with the reference's *initialisation* formulas (dimer_energy_config
 * src/EnergyEvaluation.jl:378-445, EAM set_variables :1800-1811 + :1722-1734, RuNNer :1812-1822 +
 * :1735-1747, get_r_cut src/Ensembles.jl:118-125), sets max_boxlength = 1.8*box_length
 * (src/MCStates.jl:63-87).  max_displ may be NULL (=> 0.1 everywhere, src/MCStates.jl:104). Counters and
 * ham are zeroed. */
PTMC_API int ptmc_upload_states(ptmc_handle* h, const double* pos, const double* box, const double* max_displ);
/* Optional restart data (src/ReadSave.jl:436 rebuild_states): any pointer may be NULL. */
PTMC_API int ptmc_upload_counters(ptmc_handle* h, const int64_t* counters, const double* ham);
/* reset_counters (src/MCRun.jl:270-276) on every replica. */
PTMC_API int ptmc_reset_counters(ptmc_handle* h);

/* ---- random streams ------------------------------------------------------------------------------ */
/* Consecutive Julia seeds (compatibility / parity rule; jobs that must not share streams use ptmc_seed_job): Random.seed!(n) is the
 * SHA-256 of the seed words.  With W = ptmc_rng_words()/4
 * substreams per slot (1, or PTMC_RNG_LANES in BATCH8 mode): substream k of (ladder l, slot i) =
 * Random.seed!(seed + (l*n_traj + i)*W + k), ladder ("driver task") stream of ladder l =
 * Random.seed!(seed + n_ladders*n_traj*W + l). */
PTMC_API int ptmc_seed(ptmc_handle* h, uint64_t seed);
/* Job seeding (what ptmc_create does with cfg.seed, first_ladder = 0): every stream is the SHA-256 of the PAIR (seed, stream id), so
 * two jobs with different seeds share no stream (with ptmc_seed's consecutive rule, seeds 1234 and 1235 share all but one), and a
 * ladder's streams depend on its GLOBAL number only: local ladder l is ladder first_ladder + l of the job, whatever the number of GPUs.
 *   substream k of (global ladder gl, slot i): id = (gl * n_traj + i) * W + k + 1;   driver-task stream of ladder gl: id = 2^63 + gl
 *   state = Random.seed!(seed + id * 2^64)  (Julia hashes the little-endian UInt32 words of the integer, trailing zero words dropped) */
PTMC_API int ptmc_seed_job(ptmc_handle* h, uint64_t seed, int64_t first_ladder);
/* Number of UInt64 words of generator state per slot: 4 (XOSHIRO, REPLAY) or 4*PTMC_RNG_LANES (BATCH8). */
PTMC_API int ptmc_rng_words(const ptmc_handle* h);
/* Raw Xoshiro256++ states: slot_states [n_ladders][n_traj][ptmc_rng_words()], ladder_states [n_ladders][4]. */
PTMC_API int ptmc_set_rng_states(ptmc_handle* h, const uint64_t* slot_states, const uint64_t* ladder_states);
PTMC_API int ptmc_get_rng_states(ptmc_handle* h, uint64_t* slot_states, uint64_t* ladder_states);
/* Replay mode: pre-drawn raw UInt64 draws. slot_draws [n_ladders][n_traj][slot_len] is consumed by
 * mc_move! of that slot (rand(1:N), 3 uniforms (+3 per redraw), acceptance uniform: src/MCRun.jl:145,
 * src/MCMoves.jl:34-38,170, src/MCRun.jl:123); ladder_draws [n_ladders][ladder_len] by the exchange
 * gate / pair / uniform (src/MCRun.jl:203, src/Exchange.jl:135,145).  Cursors restart at 0. */
PTMC_API int ptmc_set_replay(ptmc_handle* h, const uint64_t* slot_draws, int64_t slot_len,
                    const uint64_t* ladder_draws, int64_t ladder_len);

/* ---- histograms ---------------------------------------------------------------------------------- */
/* initialise_histograms! (src/Sampling.jl:141-223): fixes en_min/en_max (and v_min/v_max for NPT) and
 * zeroes the histograms.  With per_ladder = 0 the four pointers address one value used by every ladder,
 * otherwise arrays of n_ladders.  delta_en_hist = (en_max-en_min)/(n_bin-1), delta_v_hist =
 * (v_max-v_min)/n_bin are derived exactly as the reference does. */
PTMC_API int ptmc_set_hist_bounds(ptmc_handle* h, const double* en_min, const double* en_max,
                         const double* v_min, const double* v_max, int per_ladder);

/* ---- the hot path -------------------------------------------------------------------------------- */
/* n_cycles calls of mc_cycle! (src/MCRun.jl:192-253) for every ladder, cycle index running from
 * first_idx (the `i`/`idx` argument: step-size adaptation fires when idx % n_adjust == 0, sampling when
 * idx % mc_sample == 0).  sample = 0 gives the 7-argument (equilibration) method, sample = 1 the
 * 10-argument method with sampling_step! (src/Sampling.jl:330-371).  Asynchronous on the handle's
 * stream; errors raised inside the kernel surface at the next ptmc_sync / download. */
PTMC_API int ptmc_run_cycles(ptmc_handle* h, int64_t first_idx, int64_t n_cycles, int sample);
/* The same with mc_cycle! split at the exchange, for a ladder whose temperatures are spread over several handles / GPUs
 * (SURVEY §8e): phase 1 = mc_step! only (src/MCRun.jl:197-202), phase 2 = update_max_stepsize! + sampling_step! only
 * (:206-212, :248-250), phase 0 = ptmc_run_cycles.  Between phase 1 and phase 2 of a cycle the host performs
 * parallel_tempering_exchange! (src/Exchange.jl:132-155) on the energies of all handles and moves configurations with
 * ptmc_export_config / ptmc_import_config; the handle's own exchange step and ladder stream are not touched. */
PTMC_API int ptmc_run_phase(ptmc_handle* h, int64_t first_idx, int64_t n_cycles, int sample, int phase);
/* What exc_trajectories! (src/Exchange.jl:115-124) swaps, for one slot of a SphericalBC handle: buf holds
 * ptmc_config_doubles(h) doubles = positions [N][3], box[2], en_tot, EAM components [N][2]. */
PTMC_API int64_t ptmc_config_doubles(const ptmc_handle* h);
PTMC_API int ptmc_export_config(ptmc_handle* h, int64_t ladder, int32_t slot, double* buf);
PTMC_API int ptmc_import_config(ptmc_handle* h, int64_t ladder, int32_t slot, const double* buf);
/* Wait for the stream and report in-kernel errors (PTMC_ERR_SPHERE, PTMC_ERR_REPLAY). */
PTMC_API int ptmc_sync(ptmc_handle* h);

/* ---- parity trace (test / replay mode) ----------------------------------------------------------- */
/* Record per-move (index, new_en - en_tot, accepted) and per-cycle exchange (gate, n_exc, accepted) for
 * the next run_cycles calls; capacity in cycles.  capacity 0 disables tracing. */
PTMC_API int ptmc_trace_enable(ptmc_handle* h, int64_t capacity_cycles);
/* move_index [n_ladders][cycles][n_traj][n_steps] (1-based, > n_atoms = volume move; an NNVT atom swap records
 * index + (i1 << 10) + (i2 << 20) with the 1-based swap_indices of swap_atoms, src/MCMoves.jl:363-370),
 * delta_en same shape, accepted same shape (0/1); exchange [n_ladders][cycles][3] = {gate, n_exc, acc}.
 * Any pointer may be NULL.  *n_cycles_out = cycles recorded. */
PTMC_API int ptmc_trace_download(ptmc_handle* h, int32_t* move_index, double* delta_en, uint8_t* accepted,
                        int32_t* exchange, int64_t* n_cycles_out);

/* ---- state out ----------------------------------------------------------------------------------- */
/* Everything the reference's driver touches between cycles (checkpoint src/ReadSave.jl:178-218,
 * finalise_results src/Sampling.jl:377-399).  Any pointer may be NULL. Results are per *slot*. */
PTMC_API int ptmc_download_states(ptmc_handle* h, double* pos, double* box, double* en_tot, double* max_displ,
                         int64_t* counters, double* ham);
/* en_hist [n_ladders][n_traj][n_bin+2] (NVT) / ev_hist [n_ladders][n_traj][n_bin+2][n_bin+2] (NPT,
 * E index slowest), as Float64 counts like Output.en_histogram (src/InputParams.jl:126-127).
 * With reduce_ladders != 0 the ladder axis is summed on the device first and the arrays lose it. */
PTMC_API int ptmc_download_hist(ptmc_handle* h, double* en_hist, double* ev_hist, int reduce_ladders);
/* Device pointer + element count of the ladder-reduced histogram (uint64 counts) so that the host
 * plumbing (torch.distributed / NCCL all-reduce) can sum it across GPUs in place. */
PTMC_API int ptmc_reduced_hist_dev(ptmc_handle* h, void** dev_ptr, int64_t* n_elems);

/* ---- radial distribution function ------------------------------------------------------------------
 * update_rdf! (src/Sampling.jl:282-316), called by sampling_step! when rdfsave is set and rem(idx, 1000) == 0
 * (:343-347, :363-367): for every slot and every unordered pair, idx = floor(d2 / delta_r2); idx != 0 && idx <=
 * 5 n_bin  =>  rdf[slot][idx] += 1.  delta_r2 is results.delta_r2 of initialise_histograms! (:151,185,213,240).
 * The host loop ends a ptmc_run_cycles call on those cycles and calls ptmc_rdf_sample (distances are recomputed from
 * the positions with the reference's distance2 arithmetic; the reference reads its stored dist2_mat). */
PTMC_API int ptmc_rdf_enable(ptmc_handle* h, double delta_r2);
PTMC_API int ptmc_rdf_sample(ptmc_handle* h);
/* rdf [n_ladders][n_traj][5 n_bin] as Float64 counts like Output.rdf (src/InputParams.jl:128) */
PTMC_API int ptmc_download_rdf(ptmc_handle* h, double* rdf);

/* ---- introspection ------------------------------------------------------------------------------- */
/* Number of kernel launches issued so far by this handle, and the lanes-per-replica actually used. */
PTMC_API int64_t ptmc_launch_count(const ptmc_handle* h);
PTMC_API int ptmc_lanes_per_replica(const ptmc_handle* h);
/* Measured FP64 FMA-pipe peak of `device` in TFLOP/s (independent DFMA chains on every SM, best of
 * `iters` timed launches).  The sweep is bound by this pipe, not by HBM or the tensor cores, so bench.py
 * uses it as the roofline denominator (MEASURED_PEAKS.json holds no FP64 figure). */
PTMC_API int ptmc_fp64_peak(int device, int iters, double* tflops);

/* Issue-port probe (DESIGN.md cost model): SM cycles per warp-DFMA per SM sub-partition when 0, 8, 16 and 32 independent
 * integer instructions are interleaved with every 8 DFMAs... i.e. 0, 1, 2, 4 per DFMA; sm_mhz = clock to convert time. */
PTMC_API int ptmc_issue_probe(int device, double sm_mhz, double* cycles_per_dfma4);

/* Accuracy probe (tests): the pair arithmetic of the sweep kernel for n pairs of squared distances, exactly as
 * it is accumulated per move: e_new[k] = e(d2_new[k]), e_old[k] = e(d2_old[k]) for the handle's potential
 * (ELJ / ELJEven: dimer_energy src/EnergyEvaluation.jl:923-942; EAM: e_new = the r^-n term of invrexp :1088-1095 for d2_new,
 * e_old = the kernel's square root, :1156-1162, of d2_old). */
PTMC_API int ptmc_debug_pair_energy(ptmc_handle* h, int n, const double* d2_new, const double* d2_old, double* e_new, double* e_old);

/* RuNNer (tests): the g_matrix [n_atoms][n_rad + n_ang] (= Julia's column-major (n_sym x N) NNPVariables.g_matrix) and
 * the per-atom energies en_atom_vec of the configuration currently held by (ladder, slot). */
PTMC_API int ptmc_debug_runner_state(ptmc_handle* h, int64_t ladder, int32_t slot, double* g_matrix, double* en_atom);

/* ---- RuNNer forward pass: drop-in for librunnerjulia.so ------------------------------------------ */
/* Exact ABI of the Fortran symbol the reference ccall()s (ForwardPass.jl:60-83): every argument by
 * reference, caller owns eatom.  Re-entrant. The Float64 return value is the last atom's energy. */
PTMC_API double forward(const double* input, const int32_t* n_in, const int32_t* n_layers, const int32_t* nodes,
               const int32_t* batch, const int32_t* n_params, const double* params, const int32_t* act,
               double* eatom);

#ifdef __cplusplus
}
#endif
#endif /* PTMC_B200_H */
