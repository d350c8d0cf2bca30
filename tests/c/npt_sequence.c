/*
 * npt_sequence.c — INTEGRATION.md's call sequence for an NPT deck from plain C, through include/ptmc_b200.h only (the closest
 * stand-in for the Julia `ccall` path this image allows: no Python, no ctypes, no torch).
 *
 *   create -> upload -> seed -> equilibrate (eq_cycles / 2) -> download en_tot -> check_e_bounds / initialise_histograms! on the
 *   host -> reset_counters -> set_hist_bounds (per ladder) -> main loop in chunks -> download states + histograms -> destroy
 *   (src/MCRun.jl:363-432 with src/Sampling.jl:159-194 for CubicBC).
 *
 * Checks what julia/PTMCB200.jl relies on:
 *   1. the downloaded box belongs to the downloaded positions: a second handle that is given exactly (pos, box) recomputes en_tot
 *      (dimer_energy_config + lrc with r_cut from THAT box) and must reproduce the running energy; with the initial box instead it
 *      must not (the bug a binding makes when it forgets to write the box back);
 *   2. volume moves happened and were accepted (count_vol), boxes differ between slots, positions were scaled with them;
 *   3. every sampled cycle landed in the (E, V) histogram: sum == n_traj * mc_cycles per ladder, and the ladder-reduced download
 *      equals the sum of the per-ladder ones;
 *   4. errors come back as codes + text (bad argument), nothing aborts.
 *
 * usage: npt_sequence /path/to/libptmc_b200.so      (dlopen: the test is built with gcc alone, -ldl)
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptmc_b200.h"

#define FN(ret, name, ...) typedef ret (*name##_t)(__VA_ARGS__); static name##_t p_##name
FN(int, ptmc_abi_version, void);
FN(const char*, ptmc_last_error, const ptmc_handle*);
FN(int, ptmc_create, const ptmc_config*, ptmc_handle**);
FN(void, ptmc_destroy, ptmc_handle*);
FN(int, ptmc_upload_states, ptmc_handle*, const double*, const double*, const double*);
FN(int, ptmc_reset_counters, ptmc_handle*);
FN(int, ptmc_seed, ptmc_handle*, uint64_t);
FN(int, ptmc_set_hist_bounds, ptmc_handle*, const double*, const double*, const double*, const double*, int);
FN(int, ptmc_run_cycles, ptmc_handle*, int64_t, int64_t, int);
FN(int, ptmc_sync, ptmc_handle*);
FN(int, ptmc_download_states, ptmc_handle*, double*, double*, double*, double*, int64_t*, double*);
FN(int, ptmc_download_hist, ptmc_handle*, double*, double*, int);

#define LOAD(name) do { p_##name = (name##_t)dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; } } while (0)
#define OK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_ptmc_last_error(h)); return 3; } } while (0)
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "CHECK FAILED %s:%d: %s\n  ", __FILE__, __LINE__, #cond); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 4; } } while (0)

enum { L = 3, T = 4, N = 32, NBIN = 100, CYCLES = 120, EQ = 40 };

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s libptmc_b200.so\n", argv[0]); return 1; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(ptmc_abi_version); LOAD(ptmc_last_error); LOAD(ptmc_create); LOAD(ptmc_destroy); LOAD(ptmc_upload_states);
    LOAD(ptmc_reset_counters); LOAD(ptmc_seed); LOAD(ptmc_set_hist_bounds); LOAD(ptmc_run_cycles); LOAD(ptmc_sync);
    LOAD(ptmc_download_states); LOAD(ptmc_download_hist);
    ptmc_handle* h = NULL;
    CHECK(p_ptmc_abi_version() == PTMC_ABI_VERSION, "library ABI %d, header %d", p_ptmc_abi_version(), PTMC_ABI_VERSION);

    /* scripts/pbc32-short.jl in spirit: 32-atom fcc block (2 x 2 x 2 cells), CubicBC, NPT, ELJPotentialEven of neon, 10-25 K */
    const double a2b = 1.8897261259077824, box0 = 8.7674 * a2b, cell = box0 / 2;
    static double pos[L][T][N][3], box[L][T][2];
    const double basis[4][3] = {{0, 0, 0}, {0.5, 0.5, 0}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};
    for (int l = 0; l < L; ++l)
        for (int t = 0; t < T; ++t) {
            int k = 0;
            for (int ix = 0; ix < 2; ++ix) for (int iy = 0; iy < 2; ++iy) for (int iz = 0; iz < 2; ++iz)
                for (int b = 0; b < 4; ++b, ++k) {
                    pos[l][t][k][0] = (ix + basis[b][0]) * cell - box0 / 2 + 0.01 * sin(1.0 + k);     /* slightly off the lattice */
                    pos[l][t][k][1] = (iy + basis[b][1]) * cell - box0 / 2 + 0.01 * sin(2.0 + 3 * k);
                    pos[l][t][k][2] = (iz + basis[b][2]) * cell - box0 / 2 + 0.01 * sin(3.0 + 7 * k);
                }
            box[l][t][0] = box0; box[l][t][1] = box0;
        }
    double betas[T];
    for (int i = 0; i < T; ++i) betas[i] = 1.0 / (3.16681196e-6 * 10.0 * pow(25.0 / 10.0, (double)i / (T - 1)));   /* TempGrid geometric */

    ptmc_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = PTMC_ABI_VERSION; cfg.device = 0; cfg.n_ladders = L; cfg.n_traj = T; cfg.n_atoms = N;
    cfg.ensemble = PTMC_ENS_NPT; cfg.n_atom_moves = N; cfg.n_volume_moves = 2; cfg.bc_kind = PTMC_BC_CUBIC;
    cfg.pressure = 101325 * 2.2937122783969076e-13 / (a2b * a2b);
    cfg.pot_kind = PTMC_POT_ELJ_EVEN; cfg.n_coeff = 6;
    const double c[6] = {-10.5097942564988, 989.725135614556, -101383.865938807, 3918846.12841668, -56234083.4334278, 288738837.441765};
    memcpy(cfg.coeff, c, sizeof c);
    cfg.betas = betas; cfg.n_adjust = 10; cfg.mc_sample = 1; cfg.n_bin = NBIN; cfg.rng_mode = PTMC_RNG_BATCH8;
    cfg.min_acc = 0.4; cfg.max_acc = 0.6; cfg.seed = 1234;

    /* 4. errors are codes + text */
    ptmc_config bad = cfg; bad.n_volume_moves = 0;
    CHECK(p_ptmc_create(&bad, &h) == PTMC_ERR_ARG && h == NULL && strstr(p_ptmc_last_error(NULL), "n_volume_moves"), "bad config accepted: %s", p_ptmc_last_error(NULL));

    OK(p_ptmc_create(&cfg, &h));
    OK(p_ptmc_upload_states(h, &pos[0][0][0][0], &box[0][0][0], NULL));
    OK(p_ptmc_seed(h, 1234));
    static double en0[L][T], en[L][T], out_pos[L][T][N][3], out_box[L][T][2], md[L][T][4], ham[L][T][2];
    static int64_t cnt[L][T][PTMC_NCOUNT];
    OK(p_ptmc_download_states(h, NULL, NULL, &en0[0][0], NULL, NULL, NULL));
    /* equilibration_cycle!: eq_cycles / 2 cycles of MC, then the bounds from the current energies (src/MCRun.jl:282-313) */
    OK(p_ptmc_run_cycles(h, 1, EQ / 2, 0));
    OK(p_ptmc_download_states(h, NULL, NULL, &en[0][0], NULL, NULL, NULL));
    double emin[L], emax[L], vmin[L], vmax[L];
    for (int l = 0; l < L; ++l) {
        double lo = 100.0, hi = -100.0;                      /* check_e_bounds :258-265 (note the else-if) */
        for (int rep = EQ / 2 + 1; rep <= EQ; ++rep)
            for (int t = 0; t < T; ++t) { if (en[l][t] < lo) lo = en[l][t]; else if (en[l][t] > hi) hi = en[l][t]; }
        emin[l] = lo; emax[l] = hi;
        vmin[l] = box0 * box0 * box0 * 0.8; vmax[l] = box0 * box0 * box0 * 2.0;       /* initialise_histograms! CubicBC :171-172 */
    }
    OK(p_ptmc_reset_counters(h));
    OK(p_ptmc_set_hist_bounds(h, emin, emax, vmin, vmax, 1));
    for (int64_t i = 1; i <= CYCLES; i += 50) {              /* chunks end where a driver would checkpoint */
        int64_t k = CYCLES - i + 1 < 50 ? CYCLES - i + 1 : 50;
        OK(p_ptmc_run_cycles(h, i, k, 1));
    }
    OK(p_ptmc_sync(h));
    OK(p_ptmc_download_states(h, &out_pos[0][0][0][0], &out_box[0][0][0], &en[0][0], &md[0][0][0], &cnt[0][0][0], &ham[0][0][0]));
    static double ev[L][T][NBIN + 2][NBIN + 2], evsum[T][NBIN + 2][NBIN + 2];
    OK(p_ptmc_download_hist(h, NULL, &ev[0][0][0][0], 0));
    OK(p_ptmc_download_hist(h, NULL, &evsum[0][0][0], 1));

    /* 2. volume moves happened; boxes changed and differ */
    int64_t vol_acc = 0; int changed = 0;
    for (int l = 0; l < L; ++l) for (int t = 0; t < T; ++t) {
        vol_acc += cnt[l][t][PTMC_C_VOL_TOT];
        if (out_box[l][t][0] != box0) ++changed;
        CHECK(out_box[l][t][0] == out_box[l][t][1], "CubicBC keeps {L, L}");
        CHECK(cnt[l][t][PTMC_C_ATOM_TOT] > 0 && md[l][t][0] > 0 && md[l][t][1] > 0, "counters / step sizes of slot %d %d", l, t);
    }
    CHECK(vol_acc > 20 && changed >= L * T - 1, "accepted volume moves %lld, changed boxes %d", (long long)vol_acc, changed);
    /* 3. histograms */
    for (int l = 0; l < L; ++l) {
        double s = 0;
        for (int t = 0; t < T; ++t) for (int a = 0; a < NBIN + 2; ++a) for (int b = 0; b < NBIN + 2; ++b) s += ev[l][t][a][b];
        CHECK(s == (double)T * CYCLES, "ladder %d histogram holds %.0f samples, expected %d", l, s, T * CYCLES);
    }
    for (int t = 0; t < T; ++t) for (int a = 0; a < NBIN + 2; ++a) for (int b = 0; b < NBIN + 2; ++b) {
        double s = 0;
        for (int l = 0; l < L; ++l) s += ev[l][t][a][b];
        CHECK(s == evsum[t][a][b], "reduced histogram [%d][%d][%d]", t, a, b);
    }
    /* 1. box <-> positions consistency through a second handle */
    ptmc_handle* h1 = h;
    ptmc_handle* h2 = NULL;
    h = NULL;
    OK(p_ptmc_create(&cfg, &h2));
    h = h2;
    static double en2[L][T], en3[L][T];
    OK(p_ptmc_upload_states(h2, &out_pos[0][0][0][0], &out_box[0][0][0], &md[0][0][0]));
    OK(p_ptmc_download_states(h2, NULL, NULL, &en2[0][0], NULL, NULL, NULL));
    OK(p_ptmc_upload_states(h2, &out_pos[0][0][0][0], &box[0][0][0], &md[0][0][0]));       /* the initial box: wrong on purpose */
    OK(p_ptmc_download_states(h2, NULL, NULL, &en3[0][0], NULL, NULL, NULL));
    double worst = 0, wrong = 0;
    for (int l = 0; l < L; ++l) for (int t = 0; t < T; ++t) {
        double d = fabs(en2[l][t] - en[l][t]) / fabs(en[l][t]);
        if (d > worst) worst = d;
        d = fabs(en3[l][t] - en[l][t]) / fabs(en[l][t]);
        if (d > wrong) wrong = d;
    }
    CHECK(worst < 1e-10, "energy recomputed from the downloaded (positions, box) differs by %.3e", worst);
    CHECK(wrong > 1e-6, "the initial box reproduces the energy too (%.3e): the boxes did not matter?", wrong);
    p_ptmc_destroy(h2);
    p_ptmc_destroy(h1);
    printf("npt_sequence ok: %d ladders x %d slots, %lld accepted volume moves, |dE|/|E| of the recomputed energy %.2e, with the initial box %.2e\n",
           L, T, (long long)vol_acc, worst, wrong);
    return 0;
}
