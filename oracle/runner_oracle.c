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

/* ---- 2-species initial energy (KAT checker only; test/test_potentials.jl:67-87, potentialfile.jl) -------------
 * n1 atoms of type 1 then n2 of type 2; n_rad radial functions -> 2 rows each, n_ang angular -> 3 rows each
 * (SymmFunc.jl:320-470, 511-545); scaling rows: type-1 network rows then type-2 (scaling.data 1..88, 89..176).
 * gmaxmin_a/b: [n_rows][2] {G_max, G_min} for the two element networks. */
static void runner2_gmatrix(const double* pos, int n1, int n2, int n_rad, const double* rad_eta, int n_ang, const double* ang_eta,
                            const double* ang_lambda, const double* ang_zeta, double r_cut, const double* gmaxmin_a, const double* gmaxmin_b,
                            double* d2, double* f, double* g /* [atom][row], zeroed */) {
    const int n = n1 + n2, ns = 2 * n_rad + 3 * n_ang;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            double x = pos[3 * a] - pos[3 * b], y = pos[3 * a + 1] - pos[3 * b + 1], z = pos[3 * a + 2] - pos[3 * b + 2];
            d2[(size_t)a * n + b] = x * x + y * y + z * z;
            f[(size_t)a * n + b] = orc_cutoff(sqrt(d2[(size_t)a * n + b]), r_cut);
        }
#define GM(tab, row) (1 / ((tab)[2 * (row)] - (tab)[2 * (row) + 1]))
#define GO(tab, row) (-(tab)[2 * (row) + 1] * GM(tab, row))
#define TYPE(a) ((a) < n1 ? 0 : 1)
    for (int k = 0; k < n_rad; ++k) {
        const int base = 2 * k;
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b) {
                double fc = f[(size_t)a * n + b];
                double v = (fc != 0.0 && fc != 1.0) ? fc * exp(-rad_eta[k] * d2[(size_t)a * n + b]) : 0.0;
                /* row of the NEIGHBOUR's species: Cu-neighbour -> row 1, Zn-neighbour -> row 2 */
                g[(size_t)a * ns + base + TYPE(b)] += v;
                g[(size_t)b * ns + base + TYPE(a)] += v;
            }
        for (int a = 0; a < n; ++a)
            for (int s = 0; s < 2; ++s) {
                const double* tab = a < n1 ? gmaxmin_a : gmaxmin_b;
                double* gv = &g[(size_t)a * ns + base + s];
                *gv = *gv * GM(tab, base + s) + GO(tab, base + s);
            }
    }
    for (int k = 0; k < n_ang; ++k) {
        const int base = 2 * n_rad + 3 * k;
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b)
                for (int c = b + 1; c < n; ++c) {
                    double th[3];
                    double r2ab = d2[(size_t)a * n + b], r2ac = d2[(size_t)a * n + c], r2bc = d2[(size_t)b * n + c];
                    all_angular_measure(pos + 3 * a, pos + 3 * b, pos + 3 * c, r2ab, r2ac, r2bc, th);
                    double ep = exp_part6(ang_eta[k], r2ab, r2ac, r2bc, f[(size_t)a * n + b], f[(size_t)a * n + c], f[(size_t)b * n + c]);
                    /* row = number of Zn atoms among the two neighbours (CuCu 1, CuZn 2, ZnZn 3) */
                    g[(size_t)a * ns + base + TYPE(b) + TYPE(c)] += ep * theta_part(th[0], ang_lambda[k], ang_zeta[k]);
                    g[(size_t)b * ns + base + TYPE(a) + TYPE(c)] += ep * theta_part(th[1], ang_lambda[k], ang_zeta[k]);
                    g[(size_t)c * ns + base + TYPE(a) + TYPE(b)] += ep * theta_part(th[2], ang_lambda[k], ang_zeta[k]);
                }
        for (int a = 0; a < n; ++a)
            for (int s = 0; s < 3; ++s) {
                const double* tab = a < n1 ? gmaxmin_a : gmaxmin_b;
                double tpz = pow(2.0, 1 - ang_zeta[k]) * GM(tab, base + s);
                double* gv = &g[(size_t)a * ns + base + s];
                *gv = *gv * tpz + GO(tab, base + s);
            }
    }
}
static double runner2_total(const double* g, int n1, int n2, int ns, int n_layers, const int* nodes, const int* act,
                            const double* params1, const double* params2) {
    const int n = n1 + n2;
    double* e = (double*)xc(n, 8);
    orc_runner_forward_raw(g, n_layers, nodes, act, params1, n1, e);
    if (n2) orc_runner_forward_raw(g + (size_t)n1 * ns, n_layers, nodes, act, params2, n2, e + n1);
    double tot = 0.0;
    for (int a = 0; a < n; ++a) tot += e[a];
    free(e);
    return tot;
}
double orc_runner2_energy(const double* pos, int n1, int n2, int n_rad, const double* rad_eta, int n_ang, const double* ang_eta,
                          const double* ang_lambda, const double* ang_zeta, double r_cut, const double* gmaxmin_a, const double* gmaxmin_b,
                          int n_layers, const int* nodes, const int* act, const double* params1, const double* params2, double* f12_out) {
    const int n = n1 + n2, ns = 2 * n_rad + 3 * n_ang;
    double* d2 = (double*)xc((size_t)n * n, 8);
    double* f = (double*)xc((size_t)n * n, 8);
    double* g = (double*)xc((size_t)n * ns, 8);
    runner2_gmatrix(pos, n1, n2, n_rad, rad_eta, n_ang, ang_eta, ang_lambda, ang_zeta, r_cut, gmaxmin_a, gmaxmin_b, d2, f, g);
    if (f12_out) *f12_out = f[1];
    const double tot = runner2_total(g, n1, n2, ns, n_layers, nodes, act, params1, params2);
    free(d2); free(f); free(g);
    return tot;
}

/* calc_one_symm_val SymmFunc.jl:229-241 (radial: 0 when the cut-off function is 0 or 1; angular: symmfunc_calc :196-202) */
static double one_radial(double r2, double fc, double eta) { return (fc != 0.0 && fc != 1.0) ? fc * exp(-eta * r2) : 0.0; }
static void one_angular(const double* pos, const double* d2, const double* f, int n, int a, int b, int c, double eta, double lambda, double zeta, double gv[3]) {
    double th[3];
    const double r2ab = d2[(size_t)a * n + b], r2ac = d2[(size_t)a * n + c], r2bc = d2[(size_t)b * n + c];
    all_angular_measure(pos + 3 * a, pos + 3 * b, pos + 3 * c, r2ab, r2ac, r2bc, th);
    const double ep = exp_part6(eta, r2ab, r2ac, r2bc, f[(size_t)a * n + b], f[(size_t)a * n + c], f[(size_t)b * n + c]);
    for (int m = 0; m < 3; ++m) gv[m] = ep * theta_part(th[m], lambda, zeta);
}
/* ---- 2-species atom swap (KAT checker only; test/test_states.jl:166-175): energy change of exchanging the positions of atom a1
 * (species 1) and a2 (species 2) through calc_swap_matrix (MachineLearningPotential/DeltaMatrix.jl:319-632) as swap_energy_update uses
 * it (src/EnergyEvaluation.jl:1381-1448).  Rows are 0-based here: radial row = species of the neighbour, angular row = number of
 * species-2 atoms among the two neighbours; the statements keep the reference's order. */
double orc_runner2_swap_delta(const double* pos, int n1, int n2, int n_rad, const double* rad_eta, int n_ang, const double* ang_eta,
                              const double* ang_lambda, const double* ang_zeta, double r_cut, const double* gmaxmin_a, const double* gmaxmin_b,
                              int n_layers, const int* nodes, const int* act, const double* params1, const double* params2, int a1, int a2) {
    const int n = n1 + n2, ns = 2 * n_rad + 3 * n_ang;
    double* d2 = (double*)xc((size_t)n * n, 8);
    double* f = (double*)xc((size_t)n * n, 8);
    double* g = (double*)xc((size_t)n * ns, 8);
    double* dg = (double*)xc((size_t)n * ns, 8);
    runner2_gmatrix(pos, n1, n2, n_rad, rad_eta, n_ang, ang_eta, ang_lambda, ang_zeta, r_cut, gmaxmin_a, gmaxmin_b, d2, f, g);
    const double e_old = runner2_total(g, n1, n2, ns, n_layers, nodes, act, params1, params2);
#define DG(row, atom) dg[(size_t)(atom) * ns + base + (row)]
    for (int k = 0; k < n_rad; ++k) {                                           /* :322-371 */
        const int base = 2 * k;
        const double eta = rad_eta[k];
        for (int x = 0; x < n1; ++x) {
            if (x == a1) continue;
            const double g1 = one_radial(d2[(size_t)a1 * n + x], f[(size_t)a1 * n + x], eta);
            const double g2 = one_radial(d2[(size_t)a2 * n + x], f[(size_t)a2 * n + x], eta);
            DG(0, x) += g2 - g1; DG(1, x) += g1 - g2;
            DG(0, a1) += g2 - g1;
            DG(0, a2) += g1 - g2;
        }
        for (int x = n1; x < n; ++x) {
            if (x == a2) continue;
            const double g1 = one_radial(d2[(size_t)a1 * n + x], f[(size_t)a1 * n + x], eta);
            const double g2 = one_radial(d2[(size_t)a2 * n + x], f[(size_t)a2 * n + x], eta);
            DG(0, x) += g2 - g1; DG(1, x) += g1 - g2;
            DG(1, a1) += g2 - g1;
            DG(1, a2) += g1 - g2;
        }
        for (int a = 0; a < n; ++a)
            for (int s_ = 0; s_ < 2; ++s_) DG(s_, a) *= GM(a < n1 ? gmaxmin_a : gmaxmin_b, base + s_);
    }
    for (int k = 0; k < n_ang; ++k) {                                           /* :373-580 */
        const int base = 2 * n_rad + 3 * k;
        const double eta = ang_eta[k], la = ang_lambda[k], ze = ang_zeta[k];
        double v1[3], v2[3];
        for (int j = 0; j < n1; ++j) {                                          /* both neighbours of species 1 */
            if (j == a1) continue;
            for (int c = j + 1; c < n1; ++c) {
                if (c == a1) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                DG(0, a1) += v2[0] - v1[0]; DG(0, a2) += v1[0] - v2[0];
                DG(0, j) += v2[1] - v1[1]; DG(0, c) += v2[2] - v1[2];
                DG(1, j) += v1[1] - v2[1]; DG(1, c) += v1[2] - v2[2];
            }
        }
        for (int j = n1; j < n; ++j) {                                          /* both neighbours of species 2 */
            if (j == a2) continue;
            for (int c = j + 1; c < n; ++c) {
                if (c == a2) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                DG(2, a1) += v2[0] - v1[0]; DG(2, a2) += v1[0] - v2[0];
                DG(1, j) += v2[1] - v1[1]; DG(1, c) += v2[2] - v1[2];
                DG(2, j) += v1[1] - v2[1]; DG(2, c) += v1[2] - v2[2];
            }
        }
        for (int j = 0; j < n1; ++j) {                                          /* one neighbour of each species */
            if (j == a1) continue;
            for (int c = n1; c < n; ++c) {
                if (c == a2) continue;
                one_angular(pos, d2, f, n, a1, j, c, eta, la, ze, v1);
                one_angular(pos, d2, f, n, a2, j, c, eta, la, ze, v2);
                DG(1, a1) += v2[0] - v1[0]; DG(1, a2) += v1[0] - v2[0];
                DG(1, j) += v2[1] - v1[1]; DG(0, c) += v2[2] - v1[2];
                DG(2, j) += v1[1] - v2[1]; DG(1, c) += v1[2] - v2[2];
            }
        }
        for (int c = 0; c < n1; ++c) {                                          /* triples that hold both swapped atoms */
            if (c == a1) continue;
            one_angular(pos, d2, f, n, a1, a2, c, eta, la, ze, v1);
            DG(1, a1) -= v1[0]; DG(0, a2) += v1[0];
            DG(0, a2) -= v1[1]; DG(1, a1) += v1[1];
        }
        for (int c = n1; c < n; ++c) {
            if (c == a2) continue;
            one_angular(pos, d2, f, n, a1, a2, c, eta, la, ze, v1);
            DG(2, a1) -= v1[0]; DG(1, a2) += v1[0];
            DG(1, a2) -= v1[1]; DG(2, a1) += v1[1];
        }
        for (int a = 0; a < n; ++a)
            for (int s_ = 0; s_ < 3; ++s_) DG(s_, a) *= pow(2.0, 1 - ze) * GM(a < n1 ? gmaxmin_a : gmaxmin_b, base + s_);
    }
#undef DG
    for (size_t q = 0; q < (size_t)n * ns; ++q) g[q] += dg[q];
    const double e_new = runner2_total(g, n1, n2, ns, n_layers, nodes, act, params1, params2);
    free(d2); free(f); free(g); free(dg);
    return e_new - e_old;
}

/* ---- 2-species atom move (KAT checker only; test/test_states.jl:148-159): energy change of moving atom `atom` to `trial`,
 * through calc_delta_matrix (MachineLearningPotential/DeltaMatrix.jl:69-316) as get_new_state_vars! /
 * calc_new_runner_energy! use it (src/EnergyEvaluation.jl:1347-1418): new_g_matrix = g_matrix + delta, two forward passes. */
double orc_runner2_move_delta(const double* pos, int n1, int n2, int n_rad, const double* rad_eta, int n_ang, const double* ang_eta,
                              const double* ang_lambda, const double* ang_zeta, double r_cut, const double* gmaxmin_a, const double* gmaxmin_b,
                              int n_layers, const int* nodes, const int* act, const double* params1, const double* params2,
                              int atom, const double* trial) {
    const int n = n1 + n2, ns = 2 * n_rad + 3 * n_ang, i = atom;
    double* d2 = (double*)xc((size_t)n * n, 8);
    double* f = (double*)xc((size_t)n * n, 8);
    double* g = (double*)xc((size_t)n * ns, 8);
    double* dg = (double*)xc((size_t)n * ns, 8);
    double* nd2 = (double*)xc(n, 8);
    double* nf = (double*)xc(n, 8);
    runner2_gmatrix(pos, n1, n2, n_rad, rad_eta, n_ang, ang_eta, ang_lambda, ang_zeta, r_cut, gmaxmin_a, gmaxmin_b, d2, f, g);
    const double e_old = runner2_total(g, n1, n2, ns, n_layers, nodes, act, params1, params2);
    for (int b = 0; b < n; ++b) {
        double x = trial[0] - pos[3 * b], y = trial[1] - pos[3 * b + 1], z = trial[2] - pos[3 * b + 2];
        nd2[b] = x * x + y * y + z * z;
        nf[b] = orc_cutoff(sqrt(nd2[b]), r_cut);            /* new_f_vec = cutoff_function.(sqrt.(new_dist2_vec), r_cut) :1357-1359 */
    }
    nd2[i] = 0.0; nf[i] = 1.0;
    const int ti = TYPE(i);
    /* radial rows, calc_delta_symm_val! :69-122: the centre gains in the row of the neighbour's species, the neighbour in the row of
     * the centre's species; new_radial_symm_val! :23-25 = exponential_part(new) - exponential_part(old) */
    for (int k = 0; k < n_rad; ++k) {
        const int base = 2 * k;
        for (int b = 0; b < n; ++b) {                       /* Cu neighbours first, then Zn (loop order of the reference) */
            if (b == i) continue;
            const double gn = exp(-rad_eta[k] * nd2[b]) * nf[b] - exp(-rad_eta[k] * d2[(size_t)i * n + b]) * f[(size_t)i * n + b];
            dg[(size_t)i * ns + base + TYPE(b)] += gn;
            dg[(size_t)b * ns + base + ti] += gn;
        }
        for (int a = 0; a < n; ++a)
            for (int s_ = 0; s_ < 2; ++s_) dg[(size_t)a * ns + base + s_] *= GM(a < n1 ? gmaxmin_a : gmaxmin_b, base + s_);
    }
    /* angular rows, calc_delta_symm_val! :124-250: pairs (j, k) of the other atoms in the order CuCu, ZnZn, CuZn; centre i gains
     * in row (#Zn among j, k), j in row (#Zn among i, k), k in row (#Zn among i, j); new_angular_symm_vals :31-62 */
    for (int k = 0; k < n_ang; ++k) {
        const int base = 2 * n_rad + 3 * k;
        for (int cls = 0; cls < 3; ++cls) {
            const int j0 = cls == 1 ? n1 : 0, j1 = cls == 1 ? n : n1;
            for (int j = j0; j < j1; ++j) {
                if (j == i) continue;
                const int k0 = cls == 2 ? n1 : j + 1, k1 = cls == 0 ? n1 : n;
                for (int c = k0; c < k1; ++c) {
                    if (c == i) continue;
                    double thn[3], tho[3];
                    const double r2jk = d2[(size_t)j * n + c];
                    all_angular_measure(trial, pos + 3 * j, pos + 3 * c, nd2[j], nd2[c], r2jk, thn);
                    all_angular_measure(pos + 3 * i, pos + 3 * j, pos + 3 * c, d2[(size_t)i * n + j], d2[(size_t)i * n + c], r2jk, tho);
                    const double en = exp_part6(ang_eta[k], nd2[j], nd2[c], r2jk, nf[j], nf[c], f[(size_t)j * n + c]);
                    const double eo = exp_part6(ang_eta[k], d2[(size_t)i * n + j], d2[(size_t)i * n + c], r2jk, f[(size_t)i * n + j], f[(size_t)i * n + c], f[(size_t)j * n + c]);
                    double gv[3];
                    for (int m = 0; m < 3; ++m)
                        gv[m] = theta_part(thn[m], ang_lambda[k], ang_zeta[k]) * en - theta_part(tho[m], ang_lambda[k], ang_zeta[k]) * eo;
                    dg[(size_t)i * ns + base + TYPE(j) + TYPE(c)] += gv[0];
                    dg[(size_t)j * ns + base + ti + TYPE(c)] += gv[1];
                    dg[(size_t)c * ns + base + ti + TYPE(j)] += gv[2];
                }
            }
        }
        for (int a = 0; a < n; ++a)
            for (int s_ = 0; s_ < 3; ++s_)
                dg[(size_t)a * ns + base + s_] *= pow(2.0, 1 - ang_zeta[k]) * GM(a < n1 ? gmaxmin_a : gmaxmin_b, base + s_);
    }
    for (size_t q = 0; q < (size_t)n * ns; ++q) g[q] += dg[q];          /* new_g_matrix = g_matrix .+ delta :1379-1380 */
    const double e_new = runner2_total(g, n1, n2, ns, n_layers, nodes, act, params1, params2);
    free(d2); free(f); free(g); free(dg); free(nd2); free(nf);
    return e_new - e_old;
}
