/*
 * runner_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE) for the RuNNer neural-network potential of
 * ParallelTemperingMonteCarlo.jl: cut-off function, symmetry functions (initialisation and per-move delta
 * update) and the network forward pass, restated in the reference's operation order.
 *
 * Reference (relative to /root/reference):
 *   MachineLearningPotential/Cutoff.jl:48-58,79-86      all_angular_measure, cutoff_function
 *   MachineLearningPotential/SymmFunc.jl:24-44,86-112   RadialType2 / AngularType3 constructors (G_norm, G_offset, tpz)
 *   MachineLearningPotential/SymmFunc.jl:184-231,259-312,494-510   per-pair / per-triple values, total_symm_calc
 *   MachineLearningPotential/DeltaMatrix.jl:636-975     delta update after moving one atom
 *   MachineLearningPotential/ForwardPass.jl:50-96 + SURVEY A.2   forward (Fortran binary, semantics by probing)
 *   src/EnergyEvaluation.jl:1238-1275,1593-1621,1735-1747,1812-1822
 *
 * Pinning (tests/test_runner_oracle.py): forward == 21 for the all-ones network (test/test_runner_forward.jl:44-75);
 * forward against the reference's own librunnerjulia.so (oracle/_ref) on real weights; the 6-atom Cu4Zn2 energy
 * -0.261652899 of test/test_potentials.jl:67-87 through orc_runner2_energy, which shares every per-pair and
 * per-triple function with the single-species path; the 2-species atom-move energy change -7.5971516e-5 of
 * test/test_states.jl:148-159 through orc_runner2_move_delta (calc_delta_matrix for two species) and the swap energy change
 * -0.0017084901948 of :166-175 through orc_runner2_swap_delta (calc_swap_matrix) (the 2-species sweep itself is not part of the
 * product yet).
 */
#include "../include/ptmc_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct orc_runner {
    int n_rad, n_ang, n_sym;
    double *rad_eta, *rad_norm, *rad_off;
    int* rad_type;
    double *ang_eta, *ang_lambda, *ang_zeta, *ang_tpz, *ang_off;
    int* ang_type;
    double r_cut, boundary2;
    int n_layers, n_params;
    int *nodes, *act;
    double* params;
} orc_runner;

static void* xc(size_t n, size_t s) { void* p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

orc_runner* orc_runner_new(const ptmc_runner_desc* d) {
    orc_runner* r = (orc_runner*)xc(1, sizeof *r);
    r->n_rad = d->n_rad; r->n_ang = d->n_ang; r->n_sym = d->n_rad + d->n_ang;
    r->rad_eta = (double*)xc(d->n_rad, 8); r->rad_norm = (double*)xc(d->n_rad, 8); r->rad_off = (double*)xc(d->n_rad, 8);
    r->rad_type = (int*)xc(d->n_rad, sizeof(int));
    for (int k = 0; k < d->n_rad; ++k) {
        /* RadialType2{T}(eta, r_cut, type, [G_max, G_min]) SymmFunc.jl:40-44 */
        double gmax = d->rad_gmax_gmin[2 * k], gmin = d->rad_gmax_gmin[2 * k + 1];
        r->rad_eta[k] = d->rad_eta[k]; r->rad_type[k] = d->rad_type[k];
        r->rad_norm[k] = 1 / (gmax - gmin);
        r->rad_off[k] = -gmin * r->rad_norm[k];
    }
    r->ang_eta = (double*)xc(d->n_ang, 8); r->ang_lambda = (double*)xc(d->n_ang, 8); r->ang_zeta = (double*)xc(d->n_ang, 8);
    r->ang_tpz = (double*)xc(d->n_ang, 8); r->ang_off = (double*)xc(d->n_ang, 8);
    r->ang_type = (int*)xc(d->n_ang, sizeof(int));
    for (int k = 0; k < d->n_ang; ++k) {
        /* AngularType3{T}(eta, lambda, zeta, r_cut, type, [G_max, G_min]) SymmFunc.jl:107-112 */
        double gmax = d->ang_gmax_gmin[2 * k], gmin = d->ang_gmax_gmin[2 * k + 1];
        double g_norm = 1 / (gmax - gmin);
        r->ang_eta[k] = d->ang_eta[k]; r->ang_lambda[k] = d->ang_lambda[k]; r->ang_zeta[k] = d->ang_zeta[k];
        r->ang_type[k] = d->ang_type[k];
        r->ang_off[k] = -gmin * g_norm;
        r->ang_tpz[k] = pow(2.0, 1 - d->ang_zeta[k]) * g_norm;
    }
    r->r_cut = d->r_cut; r->boundary2 = d->boundary2;
    r->n_layers = d->n_layers; r->n_params = d->n_params;
    r->nodes = (int*)xc(d->n_layers, sizeof(int)); r->act = (int*)xc(d->n_layers, sizeof(int));
    for (int l = 0; l < d->n_layers; ++l) { r->nodes[l] = d->num_nodes[l]; r->act[l] = d->activation[l]; }
    r->params = (double*)xc(d->n_params, 8);
    memcpy(r->params, d->params, 8 * (size_t)d->n_params);
    return r;
}
void orc_runner_free(orc_runner* r) {
    if (!r) return;
    free(r->rad_eta); free(r->rad_norm); free(r->rad_off); free(r->rad_type);
    free(r->ang_eta); free(r->ang_lambda); free(r->ang_zeta); free(r->ang_tpz); free(r->ang_off); free(r->ang_type);
    free(r->nodes); free(r->act); free(r->params); free(r);
}
int orc_runner_nsym(const orc_runner* r) { return r->n_sym; }
double orc_runner_boundary2(const orc_runner* r) { return r->boundary2; }
double orc_runner_rcut(const orc_runner* r) { return r->r_cut; }

/* cutoff_function(r_ij, r_cut) Cutoff.jl:79-86: 0.5 (cos(pi r/rc) + 1) for r < rc, else 0; f(0) = 1 */
double orc_cutoff(double r_ij, double r_cut) {
    if (r_ij < r_cut) return 0.5 * (cos(M_PI * (r_ij / r_cut)) + 1);
    return 0.0;
}
/* f_matrix = cutoff_function.(sqrt.(dist2_mat), r_cut)  EnergyEvaluation.jl:1816 */
void orc_runner_fmatrix(const orc_runner* r, const double* d2, int n, double* f) {
    for (size_t k = 0; k < (size_t)n * n; ++k) f[k] = orc_cutoff(sqrt(d2[k]), r->r_cut);
}

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* all_angular_measure Cutoff.jl:48-58: cos(theta) centred on a, b, c */
static void all_angular_measure(const double* a, const double* b, const double* c, double r2ab, double r2ac, double r2bc, double th[3]) {
    double ab[3], ac[3], bc[3], nab[3], nac[3], nbc[3];
    for (int k = 0; k < 3; ++k) { ab[k] = a[k] - b[k]; ac[k] = a[k] - c[k]; bc[k] = b[k] - c[k]; nab[k] = -ab[k]; nac[k] = -ac[k]; nbc[k] = -bc[k]; }
    double dab = sqrt(r2ab), dac = sqrt(r2ac), dbc = sqrt(r2bc);
    th[0] = dot3(ab, ac) / (dab * dac);       /* thetacalc(xy, xz, disxy, disxz) = xy.xz / (disxy * disxz) */
    th[1] = dot3(nab, bc) / (dab * dbc);
    th[2] = dot3(nac, nbc) / (dac * dbc);
}
/* exponential_part(eta, r2_ij, r2_ik, r2_jk, f_ij, f_ik, f_jk) SymmFunc.jl:184-186 */
static double exp_part6(double eta, double r2ij, double r2ik, double r2jk, double fij, double fik, double fjk) {
    return exp(-eta * (r2ij + r2ik + r2jk)) * fij * fik * fjk;
}
/* theta_part(theta, lambda, zeta) = (1 + lambda theta)^zeta  SymmFunc.jl:192 */
static double theta_part(double th, double lambda, double zeta) { return pow(1 + lambda * th, zeta); }
/* calc_one_symm_val(r2_ij, fc_ij, g_norm, eta) SymmFunc.jl:231-233 */
static double radial_val(double r2, double fc, double g_norm, double eta) {
    return (fc != 0.0 && fc != 1.0) ? fc * exp(-eta * r2) * g_norm : 0.0;
}

/* total_symm_calc (single species) SymmFunc.jl:494-510 with calc_symm_vals! :259-312.  g is [atom][n_sym]
 * (= Julia's column-major (n_sym x N) g_matrix).  Quirk 3: the angular loop runs (Nrad+1):Nang. */
void orc_runner_init_g(const orc_runner* r, const double* pos, const double* d2, const double* f, int n, double* g) {
    const int ns = r->n_sym;
    double* row = (double*)xc(n, 8);
    memset(g, 0, 8 * (size_t)n * ns);
    for (int gi = 1; gi <= r->n_rad; ++gi) {
        const int k = gi - 1;
        memset(row, 0, 8 * (size_t)n);
        if (r->rad_type[k] == 11) {
            for (int a = 0; a < n; ++a)
                for (int b = a + 1; b < n; ++b) {
                    double v = radial_val(d2[(size_t)a * n + b], f[(size_t)a * n + b], r->rad_norm[k], r->rad_eta[k]);
                    row[a] += v; row[b] += v;
                }
            for (int a = 0; a < n; ++a) row[a] = row[a] + r->rad_off[k];
        }
        for (int a = 0; a < n; ++a) g[(size_t)a * ns + (gi - 1)] = row[a];
    }
    for (int gi = r->n_rad + 1; gi <= r->n_ang; ++gi) {              /* sic: (Nrad+1):Nang */
        const int k = gi - r->n_rad - 1;
        memset(row, 0, 8 * (size_t)n);
        if (r->ang_type[k] == 111) {
            for (int a = 0; a < n; ++a)
                for (int b = a + 1; b < n; ++b)
                    for (int c = b + 1; c < n; ++c) {
                        double th[3];
                        double r2ab = d2[(size_t)a * n + b], r2ac = d2[(size_t)a * n + c], r2bc = d2[(size_t)b * n + c];
                        all_angular_measure(pos + 3 * a, pos + 3 * b, pos + 3 * c, r2ab, r2ac, r2bc, th);
                        double ep = exp_part6(r->ang_eta[k], r2ab, r2ac, r2bc, f[(size_t)a * n + b], f[(size_t)a * n + c], f[(size_t)b * n + c]);
                        double g1 = ep * theta_part(th[0], r->ang_lambda[k], r->ang_zeta[k]);
                        double g2 = ep * theta_part(th[1], r->ang_lambda[k], r->ang_zeta[k]);
                        double g3 = ep * theta_part(th[2], r->ang_lambda[k], r->ang_zeta[k]);
                        row[a] += g1 * r->ang_tpz[k]; row[b] += g2 * r->ang_tpz[k]; row[c] += g3 * r->ang_tpz[k];
                    }
            for (int a = 0; a < n; ++a) row[a] = row[a] + r->ang_off[k];
        }
        for (int a = 0; a < n; ++a) g[(size_t)a * ns + (gi - 1)] = row[a];
    }
    free(row);
}

/* forward: for l = 2..L: Z = b_l 1^T + W_l A (W_l column-major n_l x n_{l-1}), A = act(Z).
 * input [batch][n_in] (= column-major n_in x batch), eatom[batch]. */
void orc_runner_forward_raw(const double* input, int n_layers, const int* nodes, const int* act, const double* params, int batch, double* eatom) {
    int maxn = 0;
    for (int l = 0; l < n_layers; ++l) if (nodes[l] > maxn) maxn = nodes[l];
    double* a = (double*)xc(maxn, 8);
    double* z = (double*)xc(maxn, 8);
    for (int b = 0; b < batch; ++b) {
        memcpy(a, input + (size_t)b * nodes[0], 8 * (size_t)nodes[0]);
        size_t off = 0;
        for (int l = 1; l < n_layers; ++l) {
            const int ni = nodes[l - 1], no = nodes[l];
            const double* W = params + off;
            const double* bias = W + (size_t)no * ni;
            for (int o = 0; o < no; ++o) z[o] = bias[o];
            for (int i = 0; i < ni; ++i)                       /* dgemm 'N','N': column sweep, k outer */
                for (int o = 0; o < no; ++o) z[o] += W[o + (size_t)i * no] * a[i];
            for (int o = 0; o < no; ++o) a[o] = act[l] == 2 ? tanh(z[o]) : z[o];
            off += (size_t)no * ni + no;
        }
        eatom[b] = a[0];
    }
    free(a); free(z);
}
void orc_runner_forward(const orc_runner* r, const double* g, int n, double* eatom) {
    orc_runner_forward_raw(g, r->n_layers, r->nodes, r->act, r->params, n, eatom);
}

/* total_thr_symm! DeltaMatrix.jl:933-975 for atom i (0-based) moving to newpos: g ([atom][n_sym]) is updated in
 * place (the caller passes a copy of g_matrix, EnergyEvaluation.jl:1248).  Per row the (j) / (j,k) contributions
 * arrive in the reference's order; the loop over symmetry functions is innermost, which does not change any
 * per-element summation order. */
void orc_runner_delta(const orc_runner* r, const double* pos, const double* newpos, const double* d2, const double* new_d2,
                      const double* f, const double* new_f, int i, int n, double* g) {
    const int ns = r->n_sym;
    /* radial_symmetry_calculation! :794-822 -> adjust_radial_symm_val! :648-657 */
    for (int k = 0; k < r->n_rad; ++k) {
        if (r->rad_type[k] != 11) continue;
        const double eta = r->rad_eta[k], gn = r->rad_norm[k];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            double* gi = &g[(size_t)i * ns + k];
            double* gj = &g[(size_t)j * ns + k];
            double tn = exp(-eta * new_d2[j]) * (new_f[j] * gn);                 /* exponential_part(eta, r_sum, f_prod*g_norm) */
            *gi += tn; *gj += tn;
            double to = exp(-eta * d2[(size_t)i * n + j]) * (-f[(size_t)i * n + j] * gn);
            *gi += to; *gj += to;
        }
    }
    /* angular_symmetry_calculation! :827-872 -> calc_new_symmetry_value! :702-747 */
    for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        for (int k = j + 1; k < n; ++k) {
            if (k == i) continue;
            const double rn_ij = new_d2[j], rn_ik = new_d2[k];
            const double r2ij = d2[(size_t)i * n + j], r2ik = d2[(size_t)i * n + k], r2jk = d2[(size_t)j * n + k];
            double thn[3], tho[3];
            all_angular_measure(newpos, pos + 3 * j, pos + 3 * k, rn_ij, rn_ik, r2jk, thn);
            all_angular_measure(pos + 3 * i, pos + 3 * j, pos + 3 * k, r2ij, r2ik, r2jk, tho);
            const double fjk = f[(size_t)j * n + k];
            for (int a = 0; a < r->n_ang; ++a) {
                if (r->ang_type[a] != 111) continue;
                const double eta = r->ang_eta[a], la = r->ang_lambda[a], ze = r->ang_zeta[a], tpz = r->ang_tpz[a];
                const double en = exp_part6(eta, rn_ij, rn_ik, r2jk, new_f[j], new_f[k], fjk);
                const double eo = exp_part6(eta, r2ij, r2ik, r2jk, f[(size_t)i * n + j], f[(size_t)i * n + k], fjk);
                const int row = r->n_rad + a;
                const int idx[3] = {i, j, k};
                for (int c = 0; c < 3; ++c) {
                    double* gv = &g[(size_t)idx[c] * ns + row];
                    double tvo = theta_part(tho[c], la, ze), tvn = theta_part(thn[c], la, ze);
                    *gv += en * tvn * tpz;                                         /* adjust_angular_symm_val! :667-672 */
                    *gv -= eo * tvo * tpz;
                }
            }
        }
    }
}

/* ================================================================== two species (RuNNerPotential2Atom) ==
 * src/EnergyEvaluation.jl:1287-1448,1623-1643,1748-1768,1823-1857; MachineLearningPotential/SymmFunc.jl:53-84,113-171
 * (RadialType2a / AngularType3a), :320-470 (calc_symm_vals! for two species), :511-545 (total_symm_calc);
 * DeltaMatrix.jl:23-316 (calc_delta_matrix), :319-632 (calc_swap_matrix).
 * Atoms 0 .. n1-1 are species 1 ("Cu"), n1 .. n-1 species 2 ("Zn").  g is [atom][n_sym] (= Julia's column-major
 * (n_sym x N) matrix): radial function k owns rows 2k (species-1 neighbour) and 2k+1 (species-2 neighbour), angular
 * function k rows 2 n_rad + 3k + {0, 1, 2} = number of species-2 atoms among the two neighbours. */
typedef struct orc_runner2 {
    int n_rad, n_ang, n_sym, n1, n2;
    double* rad_eta;
    double* rad_norm;   /* [n_rad][4] G_norm: species-1 rows 1,2 then species-2 rows 1,2 (SymmFunc.jl:64-83) */
    double* rad_off;    /* [n_rad][4] G_offset */
    double *ang_eta, *ang_lambda, *ang_zeta;
    double* ang_tpz;    /* [n_ang][6] 2^(1-zeta) * G_norm (:136-158) */
    double* ang_off;    /* [n_ang][6] */
    double r_cut, boundary2;
    int n_layers, n_params;
    int *nodes, *act;
    double *params1, *params2;
} orc_runner2;

orc_runner2* orc_runner2_new(const ptmc_runner2_desc* d) {
    orc_runner2* r = (orc_runner2*)xc(1, sizeof *r);
    r->n_rad = d->n_rad; r->n_ang = d->n_ang; r->n_sym = 2 * d->n_rad + 3 * d->n_ang; r->n1 = d->n1; r->n2 = d->n2;
    r->rad_eta = (double*)xc(d->n_rad, 8); r->rad_norm = (double*)xc(4 * d->n_rad, 8); r->rad_off = (double*)xc(4 * d->n_rad, 8);
    for (int k = 0; k < d->n_rad; ++k) {
        r->rad_eta[k] = d->rad_eta[k];
        for (int q = 0; q < 4; ++q) {                       /* RadialType2a{T}(eta, r_cut, type, G_vals_a, G_vals_b) :64-83 */
            const double* tab = q < 2 ? d->gmax_gmin_1 : d->gmax_gmin_2;
            const int row = 2 * k + (q & 1);
            const double g_norm = 1 / (tab[2 * row] - tab[2 * row + 1]);
            r->rad_norm[4 * k + q] = g_norm;
            r->rad_off[4 * k + q] = -tab[2 * row + 1] * g_norm;
        }
    }
    r->ang_eta = (double*)xc(d->n_ang, 8); r->ang_lambda = (double*)xc(d->n_ang, 8); r->ang_zeta = (double*)xc(d->n_ang, 8);
    r->ang_tpz = (double*)xc(6 * d->n_ang, 8); r->ang_off = (double*)xc(6 * d->n_ang, 8);
    for (int k = 0; k < d->n_ang; ++k) {
        r->ang_eta[k] = d->ang_eta[k]; r->ang_lambda[k] = d->ang_lambda[k]; r->ang_zeta[k] = d->ang_zeta[k];
        for (int q = 0; q < 6; ++q) {                       /* AngularType3a{T}(..., G_valsa, G_valsb) :133-171 */
            const double* tab = q < 3 ? d->gmax_gmin_1 : d->gmax_gmin_2;
            const int row = 2 * d->n_rad + 3 * k + q % 3;
            const double g_norm = 1 / (tab[2 * row] - tab[2 * row + 1]);
            r->ang_off[6 * k + q] = -tab[2 * row + 1] * g_norm;
            r->ang_tpz[6 * k + q] = pow(2.0, 1 - d->ang_zeta[k]) * g_norm;
        }
    }
    r->r_cut = d->r_cut; r->boundary2 = d->boundary2;
    r->n_layers = d->n_layers; r->n_params = d->n_params;
    r->nodes = (int*)xc(d->n_layers, sizeof(int)); r->act = (int*)xc(d->n_layers, sizeof(int));
    for (int l = 0; l < d->n_layers; ++l) { r->nodes[l] = d->num_nodes[l]; r->act[l] = d->activation[l]; }
    r->params1 = (double*)xc(d->n_params, 8); r->params2 = (double*)xc(d->n_params, 8);
    memcpy(r->params1, d->params_1, 8 * (size_t)d->n_params);
    memcpy(r->params2, d->params_2, 8 * (size_t)d->n_params);
    return r;
}
void orc_runner2_free(orc_runner2* r) {
    if (!r) return;
    free(r->rad_eta); free(r->rad_norm); free(r->rad_off); free(r->ang_eta); free(r->ang_lambda); free(r->ang_zeta);
    free(r->ang_tpz); free(r->ang_off); free(r->nodes); free(r->act); free(r->params1); free(r->params2); free(r);
}
int orc_runner2_nsym(const orc_runner2* r) { return r->n_sym; }
int orc_runner2_n1(const orc_runner2* r) { return r->n1; }
double orc_runner2_rcut(const orc_runner2* r) { return r->r_cut; }

/* calc_one_symm_val(r2_ij, fc_ij, eta) SymmFunc.jl:229-231: 0 when the cut-off function is 0 or 1 */
static double one_radial(double r2, double fc, double eta) { return (fc != 0.0 && fc != 1.0) ? fc * exp(-eta * r2) : 0.0; }
/* calc_one_symm_val(positions..., eta, lambda, zeta) :233-241 -> symmfunc_calc :196-202: exp_part * theta_part(theta) */
static void one_angular(const double* pos, const double* d2, const double* f, int n, int a, int b, int c, double eta, double lambda, double zeta, double gv[3]) {
    double th[3];
    const double r2ab = d2[(size_t)a * n + b], r2ac = d2[(size_t)a * n + c], r2bc = d2[(size_t)b * n + c];
    all_angular_measure(pos + 3 * a, pos + 3 * b, pos + 3 * c, r2ab, r2ac, r2bc, th);
    const double ep = exp_part6(eta, r2ab, r2ac, r2bc, f[(size_t)a * n + b], f[(size_t)a * n + c], f[(size_t)b * n + c]);
    for (int m = 0; m < 3; ++m) gv[m] = ep * theta_part(th[m], lambda, zeta);
}
#define G2(atom, row) g[(size_t)(atom) * ns + (row)]
/* rows 1:n1 / (n1+1):N of a function's block are scaled by the species' factors: row q of species s uses entry s * nrow + q */
static void scale_block(double* g, int ns, int n1, int n, int base, int nrow, const double* mul, const double* add) {
    for (int q = 0; q < nrow; ++q) for (int a = 0; a < n1; ++a) G2(a, base + q) = G2(a, base + q) * mul[q] + (add ? add[q] : 0.0);
    for (int q = 0; q < nrow; ++q) for (int a = n1; a < n; ++a) G2(a, base + q) = G2(a, base + q) * mul[nrow + q] + (add ? add[nrow + q] : 0.0);
}
static void scale_block_mul(double* g, int ns, int n1, int n, int base, int nrow, const double* mul) {
    for (int q = 0; q < nrow; ++q) for (int a = 0; a < n1; ++a) G2(a, base + q) = G2(a, base + q) * mul[q];
    for (int q = 0; q < nrow; ++q) for (int a = n1; a < n; ++a) G2(a, base + q) = G2(a, base + q) * mul[nrow + q];
}

/* total_symm_calc (two species) SymmFunc.jl:511-545 with calc_symm_vals! :320-367 (radial: CuCu, ZnZn, CuZn) and :369-470
 * (angular: CuCuCu, ZnZnZn, CuCuZn, CuZnZn); g [atom][n_sym] is overwritten. */
void orc_runner2_init_g(const orc_runner2* r, const double* pos, const double* d2, const double* f, int n, double* g) {
    const int ns = r->n_sym, n1 = r->n1;
    memset(g, 0, 8 * (size_t)n * ns);
    for (int k = 0; k < r->n_rad; ++k) {
        const int base = 2 * k;
        const double eta = r->rad_eta[k];
        for (int a = 0; a < n1; ++a)
            for (int b = a + 1; b < n1; ++b) { const double v = one_radial(d2[(size_t)a * n + b], f[(size_t)a * n + b], eta); G2(a, base) += v; G2(b, base) += v; }
        for (int a = n1; a < n; ++a)
            for (int b = a + 1; b < n; ++b) { const double v = one_radial(d2[(size_t)a * n + b], f[(size_t)a * n + b], eta); G2(a, base + 1) += v; G2(b, base + 1) += v; }
        for (int a = 0; a < n1; ++a)
            for (int b = n1; b < n; ++b) { const double v = one_radial(d2[(size_t)a * n + b], f[(size_t)a * n + b], eta); G2(a, base + 1) += v; G2(b, base) += v; }
        scale_block(g, ns, n1, n, base, 2, r->rad_norm + 4 * k, r->rad_off + 4 * k);
    }
    for (int k = 0; k < r->n_ang; ++k) {
        const int base = 2 * r->n_rad + 3 * k;
        const double eta = r->ang_eta[k], la = r->ang_lambda[k], ze = r->ang_zeta[k];
        double gv[3];
        for (int a = 0; a < n1; ++a) for (int b = a + 1; b < n1; ++b) for (int c = b + 1; c < n1; ++c) {
            one_angular(pos, d2, f, n, a, b, c, eta, la, ze, gv); G2(a, base) += gv[0]; G2(b, base) += gv[1]; G2(c, base) += gv[2]; }
        for (int a = n1; a < n; ++a) for (int b = a + 1; b < n; ++b) for (int c = b + 1; c < n; ++c) {
            one_angular(pos, d2, f, n, a, b, c, eta, la, ze, gv); G2(a, base + 2) += gv[0]; G2(b, base + 2) += gv[1]; G2(c, base + 2) += gv[2]; }
        for (int a = 0; a < n1; ++a) for (int b = a + 1; b < n1; ++b) for (int c = n1; c < n; ++c) {
            one_angular(pos, d2, f, n, a, b, c, eta, la, ze, gv); G2(a, base + 1) += gv[0]; G2(b, base + 1) += gv[1]; G2(c, base) += gv[2]; }
        for (int a = 0; a < n1; ++a) for (int b = n1; b < n; ++b) for (int c = b + 1; c < n; ++c) {
            one_angular(pos, d2, f, n, a, b, c, eta, la, ze, gv); G2(a, base + 2) += gv[0]; G2(b, base + 1) += gv[1]; G2(c, base + 1) += gv[2]; }
        scale_block(g, ns, n1, n, base, 3, r->ang_tpz + 6 * k, r->ang_off + 6 * k);
    }
}

/* forward_pass on the two species' columns, initialise_energy :1748-1768 / calc_new_runner_energy! :1418-1433 */
void orc_runner2_forward(const orc_runner2* r, const double* g, int n, double* eatom) {
    orc_runner_forward_raw(g, r->n_layers, r->nodes, r->act, r->params1, r->n1, eatom);
    if (r->n2 != 0) orc_runner_forward_raw(g + (size_t)r->n1 * r->n_sym, r->n_layers, r->nodes, r->act, r->params2, n - r->n1, eatom + r->n1);
}

/* new_angular_symm_vals DeltaMatrix.jl:31-62: theta_part(new) * exp_new - theta_part(old) * exp_old for the centres i, j, k */
static void new_angular(const double* trial, const double* pos, const double* d2, const double* nd2, const double* f, const double* nf,
                        int n, int i, int j, int c, double eta, double la, double ze, double gv[3]) {
    double thn[3], tho[3];
    const double r2jk = d2[(size_t)j * n + c];
    all_angular_measure(trial, pos + 3 * j, pos + 3 * c, nd2[j], nd2[c], r2jk, thn);
    all_angular_measure(pos + 3 * i, pos + 3 * j, pos + 3 * c, d2[(size_t)i * n + j], d2[(size_t)i * n + c], r2jk, tho);
    const double en = exp_part6(eta, nd2[j], nd2[c], r2jk, nf[j], nf[c], f[(size_t)j * n + c]);
    const double eo = exp_part6(eta, d2[(size_t)i * n + j], d2[(size_t)i * n + c], r2jk, f[(size_t)i * n + j], f[(size_t)i * n + c], f[(size_t)j * n + c]);
    for (int m = 0; m < 3; ++m) gv[m] = theta_part(thn[m], la, ze) * en - theta_part(tho[m], la, ze) * eo;
}

/* calc_delta_matrix DeltaMatrix.jl:256-316: g (zero-filled by the caller, EnergyEvaluation.jl:1361) receives the change of
 * every symmetry value when atom i moves to `trial`. */
void orc_runner2_delta(const orc_runner2* r, const double* pos, const double* trial, const double* d2, const double* nd2,
                       const double* f, const double* nf, int i, int n, double* g) {
    const int ns = r->n_sym, n1 = r->n1;
    const int ind = i < n1 ? 0 : 1;
    for (int k = 0; k < r->n_rad; ++k) {                     /* calc_delta_symm_val! (radial) :69-122 */
        const int base = 2 * k;
        const double eta = r->rad_eta[k];
        for (int b = 0; b < n1; ++b) {
            if (b == i) continue;
            /* new_radial_symm_val! :23-25 = exponential_part(eta, new) - exponential_part(eta, old) */
            const double gn = exp(-eta * nd2[b]) * nf[b] - exp(-eta * d2[(size_t)i * n + b]) * f[(size_t)i * n + b];
            G2(i, base) += gn;
            G2(b, base + ind) = gn;                          /* sic: assignment (:97); the entry is zero before */
        }
        for (int b = n1; b < n; ++b) {
            if (b == i) continue;
            const double gn = exp(-eta * nd2[b]) * nf[b] - exp(-eta * d2[(size_t)i * n + b]) * f[(size_t)i * n + b];
            G2(i, base + 1) += gn;
            G2(b, base + ind) += gn;
        }
        scale_block_mul(g, ns, n1, n, base, 2, r->rad_norm + 4 * k);
    }
    for (int k = 0; k < r->n_ang; ++k) {                     /* calc_delta_symm_val! (angular) :124-250 */
        const int base = 2 * r->n_rad + 3 * k;
        const double eta = r->ang_eta[k], la = r->ang_lambda[k], ze = r->ang_zeta[k];
        const int iv0 = ind, iv1 = 1 + ind, iv2 = 1 + ind, iv3 = ind;      /* i_val = (1, 2, 2, 1) (+1 for a species-2 mover) */
        double gv[3];
        for (int j = 0; j < n1; ++j) {
            if (j == i) continue;
            for (int c = j + 1; c < n1; ++c) {
                if (c == i) continue;
                new_angular(trial, pos, d2, nd2, f, nf, n, i, j, c, eta, la, ze, gv);
                G2(i, base) += gv[0]; G2(j, base + iv0) += gv[1]; G2(c, base + iv0) += gv[2];
            }
        }
        for (int j = n1; j < n; ++j) {
            if (j == i) continue;
            for (int c = j + 1; c < n; ++c) {
                if (c == i) continue;
                new_angular(trial, pos, d2, nd2, f, nf, n, i, j, c, eta, la, ze, gv);
                G2(i, base + 2) += gv[0]; G2(j, base + iv1) += gv[1]; G2(c, base + iv1) += gv[2];
            }
        }
        for (int j = 0; j < n1; ++j) {
            if (j == i) continue;
            for (int c = n1; c < n; ++c) {
                if (c == i) continue;
                new_angular(trial, pos, d2, nd2, f, nf, n, i, j, c, eta, la, ze, gv);
                G2(i, base + 1) += gv[0]; G2(j, base + iv2) += gv[1]; G2(c, base + iv3) += gv[2];
            }
        }
        scale_block_mul(g, ns, n1, n, base, 3, r->ang_tpz + 6 * k);
    }
}

/* calc_swap_matrix DeltaMatrix.jl:583-632 with calc_swap_symm_val :319-371 (radial) and :373-580 (angular): g (zero-filled by the
 * caller, EnergyEvaluation.jl:1392) receives the change of every symmetry value when atom a1 (species 1) and atom a2 (species 2)
 * exchange their positions. */
void orc_runner2_swap(const orc_runner2* r, const double* pos, const double* d2, const double* f, int a1, int a2, int n, double* g) {
    const int ns = r->n_sym, n1 = r->n1;
    for (int k = 0; k < r->n_rad; ++k) {
        const int base = 2 * k;
        const double eta = r->rad_eta[k];
        for (int x = 0; x < n1; ++x) {
            if (x == a1) continue;
            const double g1 = one_radial(d2[(size_t)a1 * n + x], f[(size_t)a1 * n + x], eta);
            const double g2 = one_radial(d2[(size_t)a2 * n + x], f[(size_t)a2 * n + x], eta);
            G2(x, base) += g2 - g1; G2(x, base + 1) += g1 - g2;
            G2(a1, base) += g2 - g1;
            G2(a2, base) += g1 - g2;
        }
        for (int x = n1; x < n; ++x) {
            if (x == a2) continue;
            const double g1 = one_radial(d2[(size_t)a1 * n + x], f[(size_t)a1 * n + x], eta);
            const double g2 = one_radial(d2[(size_t)a2 * n + x], f[(size_t)a2 * n + x], eta);
            G2(x, base) += g2 - g1; G2(x, base + 1) += g1 - g2;
            G2(a1, base + 1) += g2 - g1;
            G2(a2, base + 1) += g1 - g2;
        }
        scale_block_mul(g, ns, n1, n, base, 2, r->rad_norm + 4 * k);
    }
    for (int k = 0; k < r->n_ang; ++k) {
        const int base = 2 * r->n_rad + 3 * k;
        const double eta = r->ang_eta[k], la = r->ang_lambda[k], ze = r->ang_zeta[k];
        double v1[3], v2[3];
        for (int j = 0; j < n1; ++j) {                                          /* both neighbours of species 1 :377-424 */
            if (j == a1) continue;
            for (int c = j + 1; c < n1; ++c) {
                if (c == a1) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                G2(a1, base) += v2[0] - v1[0]; G2(a2, base) += v1[0] - v2[0];
                G2(j, base) += v2[1] - v1[1]; G2(c, base) += v2[2] - v1[2];
                G2(j, base + 1) += v1[1] - v2[1]; G2(c, base + 1) += v1[2] - v2[2];
            }
        }
        for (int j = n1; j < n; ++j) {                                          /* both neighbours of species 2 */
            if (j == a2) continue;
            for (int c = j + 1; c < n; ++c) {
                if (c == a2) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                G2(a1, base + 2) += v2[0] - v1[0]; G2(a2, base + 2) += v1[0] - v2[0];
                G2(j, base + 1) += v2[1] - v1[1]; G2(c, base + 1) += v2[2] - v1[2];
                G2(j, base + 2) += v1[1] - v2[1]; G2(c, base + 2) += v1[2] - v2[2];
            }
        }
        for (int j = 0; j < n1; ++j) {                                          /* one neighbour of each species */
            if (j == a1) continue;
            for (int c = n1; c < n; ++c) {
                if (c == a2) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                G2(a1, base + 1) += v2[0] - v1[0]; G2(a2, base + 1) += v1[0] - v2[0];
                G2(j, base + 1) += v2[1] - v1[1]; G2(c, base) += v2[2] - v1[2];
                G2(j, base + 2) += v1[1] - v2[1]; G2(c, base + 1) += v1[2] - v2[2];
            }
        }
        for (int c = 0; c < n1; ++c) {                                          /* triples that hold both swapped atoms */
            if (c == a1) continue;
            one_angular(pos, d2, f, n, a1, a2, c, eta, la, ze, v1);
            G2(a1, base + 1) -= v1[0]; G2(a2, base) += v1[0];
            G2(a2, base) -= v1[1]; G2(a1, base + 1) += v1[1];
        }
        for (int c = n1; c < n; ++c) {
            if (c == a2) continue;
            one_angular(pos, d2, f, n, a1, a2, c, eta, la, ze, v1);
            G2(a1, base + 2) -= v1[0]; G2(a2, base + 1) += v1[0];
            G2(a2, base + 1) -= v1[1]; G2(a1, base + 2) += v1[1];
        }
        scale_block_mul(g, ns, n1, n, base, 3, r->ang_tpz + 6 * k);
    }
}
#undef G2

/* ---- known-answer checkers (tests/test_runner_oracle.py): the engine functions above on one configuration -----------------
 * what = 0: initial energy (test/test_potentials.jl:67-87); 1: energy change of moving atom `a` to `trial`
 * (test/test_states.jl:148-159, get_new_state_vars! + calc_new_runner_energy! EnergyEvaluation.jl:1347-1433);
 * 2: energy change of swapping atoms a and b (test/test_states.jl:166-175, swap_energy_update :1437-1448). */
double orc_runner2_check(const ptmc_runner2_desc* d, const double* pos, int what, int a, int b, const double* trial, double* f12_out) {
    orc_runner2* r = orc_runner2_new(d);
    const int n = r->n1 + r->n2, ns = r->n_sym;
    double* d2 = (double*)xc((size_t)n * n, 8);
    double* f = (double*)xc((size_t)n * n, 8);
    double* g = (double*)xc((size_t)n * ns, 8);
    double* dg = (double*)xc((size_t)n * ns, 8);
    double* e = (double*)xc(n, 8);
    for (int p = 0; p < n; ++p)
        for (int q = 0; q < n; ++q) {
            const double x = pos[3 * p] - pos[3 * q], y = pos[3 * p + 1] - pos[3 * q + 1], z = pos[3 * p + 2] - pos[3 * q + 2];
            d2[(size_t)p * n + q] = x * x + y * y + z * z;
        }
    for (size_t k = 0; k < (size_t)n * n; ++k) f[k] = orc_cutoff(sqrt(d2[k]), r->r_cut);   /* set_variables :1838-1839 */
    if (f12_out) *f12_out = f[1];
    orc_runner2_init_g(r, pos, d2, f, n, g);
    orc_runner2_forward(r, g, n, e);
    double e_old = 0.0;
    for (int p = 0; p < n; ++p) e_old += e[p];
    double out = e_old;
    if (what != 0) {
        if (what == 1) {
            double* nd2 = (double*)xc(n, 8);
            double* nf = (double*)xc(n, 8);
            for (int q = 0; q < n; ++q) {
                const double x = trial[0] - pos[3 * q], y = trial[1] - pos[3 * q + 1], z = trial[2] - pos[3 * q + 2];
                nd2[q] = x * x + y * y + z * z;
            }
            nd2[a] = 0.0;
            for (int q = 0; q < n; ++q) nf[q] = orc_cutoff(sqrt(nd2[q]), r->r_cut);
            orc_runner2_delta(r, pos, trial, d2, nd2, f, nf, a, n, dg);
            free(nd2); free(nf);
        } else {
            orc_runner2_swap(r, pos, d2, f, a, b, n, dg);
        }
        for (size_t q = 0; q < (size_t)n * ns; ++q) g[q] = g[q] + dg[q];        /* new_g_matrix = g_matrix .+ new_g_matrix */
        orc_runner2_forward(r, g, n, e);
        double e_new = 0.0;
        for (int p = 0; p < n; ++p) e_new += e[p];
        out = e_new - e_old;
    }
    free(d2); free(f); free(g); free(dg); free(e);
    orc_runner2_free(r);
    return out;
}
