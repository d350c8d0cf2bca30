// runner_sweep.cu — the sweep for the single-species RuNNer neural-network potential (BASELINE C5):
// energy_update!(::RuNNerPotential) src/EnergyEvaluation.jl:1593-1621 = boundary guard + symmetry-function delta
// update (MachineLearningPotential/DeltaMatrix.jl:636-975) + network forward pass (ForwardPass.jl:50-96), fused with
// the move generation, the Metropolis test and the commit in ONE kernel; nothing leaves the SM between them.
//
// Work decomposition: one CTA (256 threads) = one replica for the n_steps moves of a cycle; a second,
// tiny kernel per cycle does the exchange / step-size / sampling of a ladder.  Per move the delta update touches
// every triple (i, j, k): thread (j, s) owns the centre-j terms for k = s, s+4, ... and, for k > j, the centre-i
// terms, so every accumulator has exactly one writer and the result is reproducible bit for bit.
// The reference evaluates exp() and three pow() per triple PER symmetry function; here the radial part of the
// exponential is factorised per pair and cached (ef = exp(-eta d2) f, [n_eta][N][N] in HBM/L2), the angular part
// is shared by all functions with the same (lambda, zeta) and computed with repeated squaring.
#include "runner_dev.cuh"

namespace ptmc {

struct RnSmem {          // offsets in doubles into the dynamic shared memory of the move kernel
    int pos, g0, g1, ea0, ea1, rn, ro, red, net, h1, h2, misc, pair, ptab, total;
};
__host__ __device__ inline RnSmem rn_smem(int N, const RunnerDev& R) {
    RnSmem s; int o = 0;
    s.pos = o; o += 3 * N;
    s.g0 = o; o += N * R.n_act; s.g1 = o; o += N * R.n_act;
    s.ea0 = o; o += N; s.ea1 = o; o += N;
    s.rn = o; o += (3 + R.n_eta) * N;            // new row: d2, 1/d, f, ef[n_eta]
    s.ro = o; o += (3 + R.n_eta) * N;            // old row
    s.red = o; o += (RN_THREADS / 32) * (RN_MAX_ANG + RN_MAX_RAD);
    s.net = o; o += R.h1 * R.n_act + R.h1 + R.h2 * R.h1 + R.h2 + R.h2 + 1;
    s.h1 = o; o += N * R.h1; s.h2 = o; o += N * R.h2;
    s.misc = o; o += 16;
    s.pair = o;
#ifndef PTMC_RN_GLOBAL_TABLES
    o += (3 + R.n_eta) * (N * (N - 1) / 2);      // d2 | 1/d | f | ef[n_eta], upper triangle
#endif
    s.ptab = o; o += (N * (N - 1) / 2 + 1) / 2;              // int (j << 16 | k) per pair
    s.total = o;
    return s;
}

// One dense layer for all atoms: out[a][o] = act(b[o] + sum_r W[o][r] in[a][r]) as (n_out x n_in) * (n_in x N) tiles of
// 8 output nodes x 8 atoms, one tile per warp and round, k in steps of 4.  W is column-major [r * n_out + o].
__device__ __forceinline__ void rn_layer_dmma(const double* W, const double* b, int n_out, int n_in, int act, const double* in, int in_stride,
                                              double* out, int N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int mt = (n_out + 7) >> 3, nt = (N + 7) >> 3, kt = (n_in + 3) >> 2;
    for (int t = warp; t < mt * nt; t += RN_THREADS / 32) {
        const int m = t % mt, n = t / mt;
        const int o = 8 * m + gid, ab = 8 * n + gid;
        const bool ok_o = o < n_out, ok_a = ab < N;
        double d0 = ok_o ? b[o] : 0.0, d1 = d0;
        const double* wp = W + (ok_o ? o : 0);
        const double* ip = in + (ok_a ? ab : 0) * in_stride;
        for (int k = 0; k < kt; ++k) {
            const int r = 4 * k + tig;
            const bool ok_r = r < n_in;
            const double av = (ok_o && ok_r) ? wp[r * n_out] : 0.0;
            const double bv = (ok_a && ok_r) ? ip[r] : 0.0;
            rn_dmma(d0, d1, av, bv);
        }
        const int a0 = 8 * n + 2 * tig;
        if (ok_o) {
            if (a0 < N) out[a0 * n_out + o] = rn_act(d0, act);
            if (a0 + 1 < N) out[(a0 + 1) * n_out + o] = rn_act(d1, act);
        }
    }
}
// Network forward pass for all atoms on the compact g matrix in shared memory (SURVEY A.2 semantics).
// Returns, in every thread, the sum of the per-atom energies (calc_new_runner_energy! :1269-1275).
// The two hidden layers are FP64 tensor-core GEMMs (rn_layer_dmma); the output node is a dot product split over the RN_SUB
// lanes of an atom.
__device__ double rn_forward(const RunnerDev& R, int N, const double* g, const double* net, double* h1, double* h2, double* eatom,
                             double* red) {
    const int tid = threadIdx.x;
    const double* W1 = net; const double* b1 = W1 + R.h1 * R.n_act;
    const double* W2 = b1 + R.h1; const double* b2 = W2 + R.h2 * R.h1;
    const double* W3 = b2 + R.h2; const double* b3 = W3 + R.h2;
    rn_layer_dmma(W1, b1, R.h1, R.n_act, R.act1, g, R.n_act, h1, N);
    __syncthreads();
    rn_layer_dmma(W2, b2, R.h2, R.h1, R.act2, h1, R.h1, h2, N);
    __syncthreads();
    const int a = tid / RN_SUB, q = tid % RN_SUB;
    const bool on = a < N;
    double part = 0.0;
    if (on) for (int o = q; o < R.h2; o += RN_SUB) part = fma(W3[o], h2[a * R.h2 + o], part);
#pragma unroll
    for (int m = 1; m < RN_SUB; m <<= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    double e = 0.0;
    if (on && q == 0) {
        e = rn_act(part + b3[0], R.act3);
        eatom[a] = e;
    }
    e = group_sum<32>(e, 0xffffffffu);
    if ((tid & 31) == 0) red[tid >> 5] = e;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < RN_THREADS / 32; ++w) tot += red[w];
    __syncthreads();
    return tot;
}

// -------------------------------------------------------------------------------------------- init --
// set_variables(::RuNNerPotential) src/EnergyEvaluation.jl:1812-1822 + initialise_energy :1735-1747 for one replica
// per CTA: dist2 / 1/d / f / ef matrices, the g matrix from scratch (total_symm_calc SymmFunc.jl:494-510, including
// its (Nrad+1):Nang loop bound) and the first energies.
__global__ void __launch_bounds__(RN_THREADS) runner_init_kernel(const DevParams P, const RunnerDev R) {
    extern __shared__ double sm[];
    const int N = P.n_atoms, tid = threadIdx.x;
    const i64 rep = blockIdx.x;
    const RnSmem L = rn_smem(N, R);
    double* pos = sm + L.pos; double* g = sm + L.g0; double* net = sm + L.net;
    const int NT = N * (N - 1) / 2;
#ifdef PTMC_RN_GLOBAL_TABLES
    double* D2 = R.pair_c + rep * (size_t)(3 + R.n_eta) * NT;
#else
    double* D2 = sm + L.pair;
#endif
    double* INV = D2 + NT; double* F = INV + NT; double* EF = F + NT;
    for (int k = tid; k < 3 * N; k += RN_THREADS) { pos[k] = P.pos[rep * 3 * N + k]; R.pos_c[rep * 3 * N + k] = pos[k]; }
    for (int k = tid; k < L.h1 - L.net; k += RN_THREADS) net[k] = R.net[k];
    __syncthreads();
    for (int p = tid; p < N * N; p += RN_THREADS) {
        const int a = p / N, b = p - a * N;
        if (a >= b) continue;
        const double x = xsub(pos[3 * a], pos[3 * b]), y = xsub(pos[3 * a + 1], pos[3 * b + 1]), z = xsub(pos[3 * a + 2], pos[3 * b + 2]);
        const double d2 = xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));     // distance2 src/Configurations.jl:101
        const double r = sqrt(d2);
        const double f = rn_cutoff(r, R.r_cut);
        const int t = rn_tri(a, b, N);
        D2[t] = d2; INV[t] = 1.0 / r; F[t] = f;
        for (int e = 0; e < R.n_eta; ++e) EF[e * NT + t] = exp(-R.eta[e] * d2) * f;
    }
    __syncthreads();
    // one thread per (atom a, symmetry function): every contribution with centre a, in the (b, c) order of calc_symm_vals!
    // (SymmFunc.jl:259-312) - 55 x 31 work items instead of 55 threads walking all functions (the upload was 4 cycles' worth of time)
    for (int w = tid; w < N * R.n_act_rad; w += RN_THREADS) {
        const int a = w / R.n_act_rad, r = w - a * R.n_act_rad;
        double s = 0.0;
        for (int b = 0; b < N; ++b) {
            if (b == a) continue;
            const int t = rn_tri(a, b, N);
            const double f = F[t];
            if (f != 0.0 && f != 1.0) s += f * exp(-R.rad_eta[r] * D2[t]) * R.rad_norm[r];   // calc_one_symm_val :231
        }
        g[a * R.n_act + r] = s + R.rad_off[r];
    }
    for (int w = tid; w < N * R.n_act_ang; w += RN_THREADS) {
        const int a = w / R.n_act_ang, q = w - a * R.n_act_ang;
        double s = 0.0;
        if (R.ang_init[q]) {
            const int e = R.ang_eta_idx[q], z = R.ang_lz_idx[q];
            const double* EFe = EF + e * NT;
            const double la = R.lz_lambda[z], ze = R.lz_zeta[z], tpz = R.ang_tpz[q];
            const int zi = R.lz_zint[z];
            for (int b = 0; b < N; ++b) {
                if (b == a) continue;
                const int tab = rn_tri(a, b, N);
                const double dab = D2[tab], iab = INV[tab], eab = EFe[tab];
                for (int c = b + 1; c < N; ++c) {
                    if (c == a) continue;
                    const int tac = rn_tri(a, c, N), tbc = rn_tri(b, c, N);
                    const double ca = (dab + D2[tac] - D2[tbc]) * 0.5 * iab * INV[tac];
                    const double ep = eab * EFe[tac] * EFe[tbc];
                    s += ep * rn_theta(ca, la, ze, zi) * tpz;
                }
            }
            s += R.ang_off[q];
        }
        g[a * R.n_act + R.n_act_rad + q] = s;
    }
#ifndef PTMC_RN_GLOBAL_TABLES
    for (int k = tid; k < (3 + R.n_eta) * NT; k += RN_THREADS) R.pair_c[rep * (size_t)(3 + R.n_eta) * NT + k] = D2[k];
#endif
    __syncthreads();
    double* ea = sm + L.ea0;
    const double e = rn_forward(R, N, g, net, sm + L.h1, sm + L.h2, ea, sm + L.red);
    for (int k = tid; k < N * R.n_act; k += RN_THREADS) R.g_c[rep * N * R.n_act + k] = g[k];
    for (int k = tid; k < N; k += RN_THREADS) R.en_atom_c[rep * N + k] = ea[k];
    if (tid == 0) { R.en_c[rep] = e; P.en_tot[rep] = e; }
}

// ------------------------------------------------------------------------------------------- moves --
// mc_step! for one replica slot: n_steps x mc_move! (src/MCRun.jl:142-182).  NA = number of evaluated angular
// symmetry functions (unused; the accumulators are dense over (eta, lambda-zeta)).
template <bool DBG, unsigned long long PRESENT, unsigned NSQ>
__global__ void __launch_bounds__(RN_THREADS, RN_THREADS <= 256 ? 2 : 1)
runner_moves_kernel(const DevParams P, const RunnerDev R, const i64 cyc_in_launch) {
    extern __shared__ double sm[];
    constexpr bool BATCH = !DBG;
    constexpr bool SPEC = PRESENT != 0;               // structure of the angular function set known at compile time
    const unsigned long long present = SPEC ? PRESENT : R.present;
    const int N = P.n_atoms, T = P.n_traj, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 rep = blockIdx.x;                       // replica slot [ladder][slot]
    const int lad = (int)(rep / T), slot = (int)(rep - (i64)lad * T);
    const i64 cfg = (i64)lad * T + P.perm_g[rep];    // configuration this slot holds
    const RnSmem L = rn_smem(N, R);
    double* pos = sm + L.pos;
    double* gbuf[2] = {sm + L.g0, sm + L.g1};
    double* eabuf[2] = {sm + L.ea0, sm + L.ea1};
    double* rn_d2 = sm + L.rn; double* rn_inv = rn_d2 + N; double* rn_f = rn_inv + N; double* rn_ef = rn_f + N;
    double* ro_d2 = sm + L.ro; double* ro_inv = ro_d2 + N; double* ro_f = ro_inv + N; double* ro_ef = ro_f + N;
    double* red = sm + L.red; double* net = sm + L.net;
    double* misc = sm + L.misc;        // 0 index, 1-3 trial, 4 u_acc, 5 accept, 6 error
    const int NT = N * (N - 1) / 2;
    double* pair_g = R.pair_c + cfg * (size_t)(3 + R.n_eta) * NT;
#ifdef PTMC_RN_GLOBAL_TABLES      // experiment: pair tables read from HBM / L1 instead of shared memory
    double* D2 = pair_g; double* INV = D2 + NT; double* F = INV + NT; double* EF = F + NT;
#else
    double* D2 = sm + L.pair; double* INV = D2 + NT; double* F = INV + NT; double* EF = F + NT;   // pair tables, shared memory
#endif
    int cur = 0;
#ifndef PTMC_RN_GLOBAL_TABLES
    for (int k = tid; k < (3 + R.n_eta) * NT; k += RN_THREADS) D2[k] = pair_g[k];
#endif
    int* ptab = reinterpret_cast<int*>(sm + L.ptab);
    for (int p = tid; p < N * N; p += RN_THREADS) {
        const int a = p / N, b = p - a * N;
        if (a < b) ptab[rn_tri(a, b, N)] = a << 16 | b;
    }

    for (int k = tid; k < 3 * N; k += RN_THREADS) pos[k] = R.pos_c[cfg * 3 * N + k];
    for (int k = tid; k < N * R.n_act; k += RN_THREADS) gbuf[0][k] = R.g_c[cfg * N * R.n_act + k];
    for (int k = tid; k < N; k += RN_THREADS) eabuf[0][k] = R.en_atom_c[cfg * N + k];
    for (int k = tid; k < L.h1 - L.net; k += RN_THREADS) net[k] = R.net[k];
    double en = R.en_c[cfg];
    const double beta = P.betas[slot];
    double dmax = P.max_displ[rep * 4];
    const double R2 = P.box[rep * 2];
    int n_acc = 0;
    // random streams: warp 0 only
    Stream<DBG> S; BatchRng<8> B;
    S.draws = nullptr; S.pos = 0; S.len = 0; S.exhausted = 0;
    if (warp == 0) {
        if (BATCH) B.load(P.slot_rng + rep * (4 * PTMC_RNG_LANES), lane & 7);
        else {
            S.r.s0 = P.slot_rng[rep * 4]; S.r.s1 = P.slot_rng[rep * 4 + 1]; S.r.s2 = P.slot_rng[rep * 4 + 2]; S.r.s3 = P.slot_rng[rep * 4 + 3];
            if (P.replay) { S.draws = P.slot_draws + rep * P.slot_len; S.len = P.slot_len; S.pos = P.slot_cursor[rep]; }
        }
    }
    const bool tracing = P.trace_cap > 0;
    int err = 0;
    __syncthreads();

    for (int step = 0; step < P.n_steps; ++step) {
        double* g = gbuf[cur]; double* ng = gbuf[cur ^ 1];
        // ============================================================ phase 0: draws + trial position (warp 0) ==
        if (warp == 0) {
            const unsigned full = 0xffffffffu;
            int index;
            if (BATCH) { B.begin(); index = (int)__umul64hi(B.template raw<0>(full), (u64)P.n_steps) + 1; }
            else index = S.range((u64)P.n_steps);
            const int i = index - 1;
            double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
            double tx, ty, tz;
            if (BATCH) {
                tx = xadd(xi, xmul(xsub(B.template rand<1>(full), 0.5), dmax));
                ty = xadd(yi, xmul(xsub(B.template rand<2>(full), 0.5), dmax));
                tz = xadd(zi, xmul(xsub(B.template rand<3>(full), 0.5), dmax));
            } else {
                tx = xadd(xi, xmul(xsub(S.rand(), 0.5), dmax)); ty = xadd(yi, xmul(xsub(S.rand(), 0.5), dmax)); tz = xadd(zi, xmul(xsub(S.rand(), 0.5), dmax));
            }
            int count = 0, redraws = 0;
            while (xadd(xadd(xmul(tx, tx), xmul(ty, ty)), xmul(tz, tz)) > R2) {      // src/MCMoves.jl:115-145
                ++count;
                if (count == 50) {
                    __syncwarp();
                    double cx = 0, cy = 0, cz = 0;
                    for (int j = 0; j < N; ++j) cx = xadd(cx, pos[3 * j]);
                    for (int j = 0; j < N; ++j) cy = xadd(cy, pos[3 * j + 1]);
                    for (int j = 0; j < N; ++j) cz = xadd(cz, pos[3 * j + 2]);
                    cx = xdiv(cx, (double)N); cy = xdiv(cy, (double)N); cz = xdiv(cz, (double)N);
                    __syncwarp();
                    for (int j = lane; j < N; j += 32) { pos[3 * j] = xsub(pos[3 * j], cx); pos[3 * j + 1] = xsub(pos[3 * j + 1], cy); pos[3 * j + 2] = xsub(pos[3 * j + 2], cz); }
                    __syncwarp();
                    xi = pos[3 * i]; yi = pos[3 * i + 1]; zi = pos[3 * i + 2];
                } else {
                    if (BATCH) {
                        if (++redraws == 1) {
                            tx = xadd(xi, xmul(xsub(B.template rand<5>(full), 0.5), dmax));
                            ty = xadd(yi, xmul(xsub(B.template rand<6>(full), 0.5), dmax));
                            tz = xadd(zi, xmul(xsub(B.template rand<7>(full), 0.5), dmax));
                        } else {
                            B.begin();
                            tx = xadd(xi, xmul(xsub(B.template rand<0>(full), 0.5), dmax));
                            ty = xadd(yi, xmul(xsub(B.template rand<1>(full), 0.5), dmax));
                            tz = xadd(zi, xmul(xsub(B.template rand<2>(full), 0.5), dmax));
                        }
                    } else {
                        tx = xadd(xi, xmul(xsub(S.rand(), 0.5), dmax)); ty = xadd(yi, xmul(xsub(S.rand(), 0.5), dmax)); tz = xadd(zi, xmul(xsub(S.rand(), 0.5), dmax));
                    }
                    if (count == 100) { err = PTMC_ERR_SPHERE; break; }
                }
                if (DBG) { if (S.exhausted) break; }
            }
            const double u = BATCH ? B.template rand<4>(full) : S.rand();
            if (lane == 0) { misc[0] = (double)i; misc[1] = tx; misc[2] = ty; misc[3] = tz; misc[4] = u; }
            if (DBG) { if (S.exhausted) err = PTMC_ERR_REPLAY; }
        }
        __syncthreads();
        const int i = (int)misc[0];
        const double tx = misc[1], ty = misc[2], tz = misc[3];
        // =============================================== phase 1: new / old row of atom i, radial functions ==
        int guard = 0;
        double prad[RN_MAX_RAD];
#pragma unroll
        for (int r = 0; r < RN_MAX_RAD; ++r) prad[r] = 0.0;
        // thread (j, s) = (tid / RN_SUB, tid % RN_SUB): the RN_SUB lanes of atom j share its exponentials (angular eta e
        // on lane e % RN_SUB, radial function r on lane (RN_SUB - 1 - r) % RN_SUB); the scalars are computed by every lane
        if (tid < N * RN_SUB) {
            const int j = tid / RN_SUB, ls = tid % RN_SUB;
            const double x = xsub(tx, pos[3 * j]), y = xsub(ty, pos[3 * j + 1]), z = xsub(tz, pos[3 * j + 2]);
            double d2n = xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));
            const bool self = j == i;
            if (self) d2n = 0.0;                                              // new_dist2_vec[index] = 0
            guard = (!self && d2n < R.boundary2) ? 1 : 0;                     // EnergyEvaluation.jl:1602-1606
            const double rr = sqrt(d2n);
            const double fn = self ? 1.0 : rn_cutoff(rr, R.r_cut);            // cutoff_function.(sqrt.(new_dist2_vec))
            const int tij = self ? 0 : rn_tri(i, j, N);
            const double d2o = self ? 0.0 : D2[tij], fo = self ? 1.0 : F[tij];
            if (ls == 0) {
                rn_d2[j] = d2n; rn_inv[j] = self ? 0.0 : 1.0 / rr; rn_f[j] = fn;
                ro_d2[j] = d2o; ro_inv[j] = self ? 0.0 : INV[tij]; ro_f[j] = fo;
            }
            for (int e = ls; e < R.n_eta; e += RN_SUB) {
                rn_ef[e * N + j] = self ? 0.0 : exp(-R.eta[e] * d2n) * fn;
                ro_ef[e * N + j] = self ? 0.0 : EF[e * NT + tij];
            }
            if (!self) {
                // radial_symmetry_calculation! DeltaMatrix.jl:794-822 / adjust_radial_symm_val! :648-657
#pragma unroll
                for (int r = 0; r < RN_MAX_RAD; ++r) {
                    if (r < R.n_act_rad && (RN_SUB - 1 - r % RN_SUB) == ls) {
                        const double tn = exp(-R.rad_eta[r] * d2n) * (fn * R.rad_norm[r]);
                        const double to = exp(-R.rad_eta[r] * d2o) * (-fo * R.rad_norm[r]);
                        ng[j * R.n_act + r] = (g[j * R.n_act + r] + tn) + to;
                        prad[r] = tn + to;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RN_MAX_RAD; ++r) {
            if (r < R.n_act_rad) {
                const double v = group_sum<32>(prad[r], 0xffffffffu);
                if (lane == 0) red[warp * (RN_MAX_ANG + RN_MAX_RAD) + r] = v;
            }
        }
        const int any_guard = __syncthreads_or(guard);
        double new_en;
        if (any_guard) {
            new_en = 100.0;
        } else {
            if (tid < R.n_act_rad) {
                double v = 0.0;
                for (int w = 0; w < RN_THREADS / 32; ++w) v += red[w * (RN_MAX_ANG + RN_MAX_RAD) + tid];
                ng[i * R.n_act + tid] = g[i * R.n_act + tid] + v;
            }
            // ======================================== phase 2: angular functions, triples (i, j, k) ==========
            // angular_symmetry_calculation! DeltaMatrix.jl:827-872 / calc_new_symmetry_value! :702-747
            // Accumulators are dense over (eta, (lambda, zeta)) so that every register index is a compile-time constant;
            // combinations without a symmetry function are skipped through the uniform bit mask R.present.
            double acc[RN_MAX_ETA][RN_MAX_LZ];
#pragma unroll
            for (int e = 0; e < RN_MAX_ETA; ++e)
#pragma unroll
                for (int z = 0; z < RN_MAX_LZ; ++z) acc[e][z] = 0.0;
            // The two passes are independent and run side by side: the first nwa warps take the centre-j terms (pass A), the
            // remaining warps the centre-i terms (pass B); N = 55: 7 warps x 13 k-iterations next to 5 warps x 10 pair iterations.
            const int nwa = (N * RN_SUB + 31) >> 5;
            constexpr bool SPLIT = RN_THREADS >= 384;       // enough warps to run the two passes side by side
            if (!SPLIT || warp < nwa) {
            // ---- pass A: centre j, thread (j, s) takes k = s, s + RN_SUB, ...
            const int j = tid / RN_SUB, s = tid % RN_SUB;
            const bool jok = j < N && j != i;
            if (jok) {
                const double an = rn_d2[j], ao = ro_d2[j], ijn = rn_inv[j], ijo = ro_inv[j];
                for (int k = s; k < N; k += RN_SUB) {
                    if (k == i || k == j) continue;
                    const int tjk = rn_tri(j, k, N);
                    const double c = D2[tjk], ijk = INV[tjk];
                    // cos(theta) at centre j, new and old geometry (law of cosines on the cached d2, 1/d)
                    const double cn = (an + c - rn_d2[k]) * 0.5 * ijn * ijk, co = (ao + c - ro_d2[k]) * 0.5 * ijo * ijk;
                    double epn[RN_MAX_ETA], epo[RN_MAX_ETA];
#pragma unroll
                    for (int e = 0; e < RN_MAX_ETA; ++e) {
                        epn[e] = 0.0; epo[e] = 0.0;
                        if (SPEC ? ((PRESENT >> (e * RN_MAX_LZ)) & 0xffull) != 0 : e < R.n_eta) {
                            const double ejk = EF[e * NT + tjk];
                            epn[e] = rn_ef[e * N + j] * rn_ef[e * N + k] * ejk;      // exponential_part, new geometry
                            epo[e] = ro_ef[e * N + j] * ro_ef[e * N + k] * ejk;      // exponential_part, old geometry
                        }
                    }
                    rn_angular_acc<SPEC, PRESENT, NSQ>(R, present, cn, co, epn, epo, acc);
                }
            }
#pragma unroll
            for (int e = 0; e < RN_MAX_ETA; ++e)
#pragma unroll
                for (int z = 0; z < RN_MAX_LZ; ++z)
                    if (present >> (e * RN_MAX_LZ + z) & 1ull) {
                        double v = acc[e][z];
#pragma unroll
                        for (int m = 1; m < RN_SUB; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                        const int q = R.qmap[e * RN_MAX_LZ + z];
                        if (s == 0 && jok) ng[j * R.n_act + R.n_act_rad + q] = g[j * R.n_act + R.n_act_rad + q] + v * R.ang_tpz[q];
                        acc[e][z] = 0.0;
                    }
            }
            if (!SPLIT || warp >= nwa) {
            // ---- pass B: centre i, the pairs (j < k) dealt round-robin to the threads of the remaining warps
            for (int p = SPLIT ? tid - 32 * nwa : tid; p < NT; p += SPLIT ? RN_THREADS - 32 * nwa : RN_THREADS) {
                const int jk = ptab[p], pj = jk >> 16, pk = jk & 0xffff;
                if (pj == i || pk == i) continue;
                const double c = D2[p];
                const double an = rn_d2[pj], ao = ro_d2[pj], bn = rn_d2[pk], bo = ro_d2[pk];
                const double cn = (an + bn - c) * 0.5 * rn_inv[pj] * rn_inv[pk], co = (ao + bo - c) * 0.5 * ro_inv[pj] * ro_inv[pk];
                double epn[RN_MAX_ETA], epo[RN_MAX_ETA];
#pragma unroll
                for (int e = 0; e < RN_MAX_ETA; ++e) {
                    epn[e] = 0.0; epo[e] = 0.0;
                    if (SPEC ? ((PRESENT >> (e * RN_MAX_LZ)) & 0xffull) != 0 : e < R.n_eta) {
                        const double ejk = EF[e * NT + p];
                        epn[e] = rn_ef[e * N + pj] * rn_ef[e * N + pk] * ejk;
                        epo[e] = ro_ef[e * N + pj] * ro_ef[e * N + pk] * ejk;
                    }
                }
                rn_angular_acc<SPEC, PRESENT, NSQ>(R, present, cn, co, epn, epo, acc);
            }
#pragma unroll
            for (int e = 0; e < RN_MAX_ETA; ++e)
#pragma unroll
                for (int z = 0; z < RN_MAX_LZ; ++z)
                    if (present >> (e * RN_MAX_LZ + z) & 1ull) {
                        const double w = group_sum<32>(acc[e][z], 0xffffffffu);
                        if (lane == 0) red[warp * (RN_MAX_ANG + RN_MAX_RAD) + RN_MAX_RAD + R.qmap[e * RN_MAX_LZ + z]] = w;
                    }
            }
            __syncthreads();
            if (tid < R.n_act_ang) {
                double v = 0.0;
                for (int w = SPLIT ? nwa : 0; w < RN_THREADS / 32; ++w) v += red[w * (RN_MAX_ANG + RN_MAX_RAD) + RN_MAX_RAD + tid];
                ng[i * R.n_act + R.n_act_rad + tid] = g[i * R.n_act + R.n_act_rad + tid] + v * R.ang_tpz[tid];
            }
            __syncthreads();
            // ======================================== phase 3: network forward pass on new_g_matrix ==========
            new_en = rn_forward(R, N, ng, net, sm + L.h1, sm + L.h2, eabuf[cur ^ 1], red);
        }
        // ============================================================ phase 4: acc_test! src/MCRun.jl:122-132 ==
        const double de = xsub(new_en, en);
        const bool acc = metropolis_accept(-xmul(de, beta), misc[4]);
        {
            if (tracing && tid == 0) {
                size_t o = ((((size_t)lad * P.trace_cap + P.trace_base + cyc_in_launch) * T + slot) * P.n_steps) + step;
                P.tr_index[o] = i + 1; P.tr_de[o] = de; P.tr_acc[o] = acc ? 1 : 0;
            }
        }
        __syncthreads();       // everybody has read misc[] and the row arrays are final
        if (acc) {
            // swap_atom_config! src/swap_config.jl:40-51 + swap_vars!(::NNPVariables) :213-221
            if (tid == 0) { pos[3 * i] = tx; pos[3 * i + 1] = ty; pos[3 * i + 2] = tz; }
            if (tid < N && tid != i) {
                const int j = tid, tij = rn_tri(i, j, N);
                D2[tij] = rn_d2[j]; INV[tij] = rn_inv[j]; F[tij] = rn_f[j];     // row and column i of dist2_mat / f_matrix
                for (int e = 0; e < R.n_eta; ++e) EF[e * NT + tij] = rn_ef[e * N + j];
            }
            en = new_en; n_acc += 1;
            if (!any_guard) cur ^= 1;     // g_matrix <-> new_g_matrix, en_atom_vec <-> new_en_atom
            // (a guarded move that is accepted would swap in stale matrices in the reference; with new_en = 100
            //  that needs u == 0 exactly and is not reproduced)
        }
        __syncthreads();
    }

    // ---- store ------------------------------------------------------------------------------------------
    for (int k = tid; k < 3 * N; k += RN_THREADS) R.pos_c[cfg * 3 * N + k] = pos[k];
    for (int k = tid; k < N * R.n_act; k += RN_THREADS) R.g_c[cfg * N * R.n_act + k] = gbuf[cur][k];
    for (int k = tid; k < N; k += RN_THREADS) R.en_atom_c[cfg * N + k] = eabuf[cur][k];
#ifndef PTMC_RN_GLOBAL_TABLES
    if (n_acc) for (int k = tid; k < (3 + R.n_eta) * NT; k += RN_THREADS) pair_g[k] = D2[k];
#endif
    if (tid == 0) {
        R.en_c[cfg] = en;
        P.counters[rep * PTMC_NCOUNT + PTMC_C_ATOM_TOT] += n_acc;
        P.counters[rep * PTMC_NCOUNT + PTMC_C_ATOM_ADJ] += n_acc;
        if (!BATCH) {
            P.slot_rng[rep * 4] = S.r.s0; P.slot_rng[rep * 4 + 1] = S.r.s1; P.slot_rng[rep * 4 + 2] = S.r.s2; P.slot_rng[rep * 4 + 3] = S.r.s3;
            if (P.replay) P.slot_cursor[rep] = S.pos;
        }
    }
    if (BATCH && warp == 0 && lane < 8) B.store(P.slot_rng + rep * (4 * PTMC_RNG_LANES), lane);
    if (err) atomicMin(P.error, err);
}

// ------------------------------------------------------------------------------------ end of cycle --
// Second half of mc_cycle! (src/MCRun.jl:203-212, 248-250) for one ladder per warp: exchange attempt
// (src/Exchange.jl:132-155), update_max_stepsize! (:250-263), sampling_step! (src/Sampling.jl:330-349).
template <bool DBG>
__global__ void runner_cycle_end_kernel(const DevParams P, const double* __restrict__ en_c, const i64 idx, const i64 cyc_in_launch, const int sample) {
    const int lad = blockIdx.x, T = P.n_traj, lane = threadIdx.x;
    int* perm = P.perm_g + (i64)lad * T;
    if (lane == 0) {
        Stream<DBG> LS;
        LS.r.s0 = P.ladder_rng[lad * 4]; LS.r.s1 = P.ladder_rng[lad * 4 + 1]; LS.r.s2 = P.ladder_rng[lad * 4 + 2]; LS.r.s3 = P.ladder_rng[lad * 4 + 3];
        LS.draws = nullptr; LS.pos = 0; LS.len = 0; LS.exhausted = 0;
        if (DBG) { if (P.replay) { LS.draws = P.ladder_draws + lad * P.ladder_len; LS.len = P.ladder_len; LS.pos = P.ladder_cursor[lad]; } }
        if (LS.rand() < 0.1) {
            const int n_exc = LS.range((u64)(T - 1));
            const double u = LS.rand();
            const int a = n_exc - 1, b = n_exc;
            i64* ca = P.counters + ((i64)lad * T + a) * PTMC_NCOUNT; i64* cb = P.counters + ((i64)lad * T + b) * PTMC_NCOUNT;
            ca[PTMC_C_EXC_ATT] += 1; cb[PTMC_C_EXC_ATT] += 1;
            const double ea = en_c[(i64)lad * T + perm[a]], eb = en_c[(i64)lad * T + perm[b]];
            double pe = exp(xmul(xsub(P.betas[a], P.betas[b]), xsub(ea, eb)));
            pe = (pe != pe) ? pe : (pe < 1.0 ? pe : 1.0);
            int ex_acc = 0;
            if (pe > u) {
                ex_acc = 1;
                ca[PTMC_C_EXC_ACC] += 1; cb[PTMC_C_EXC_ACC] += 1;
                const int t = perm[a]; perm[a] = perm[b]; perm[b] = t;
            }
            {
                if (P.trace_cap > 0) {
                    int32_t* e = P.tr_exc + ((size_t)lad * P.trace_cap + P.trace_base + cyc_in_launch) * 3;
                    e[0] = 1; e[1] = n_exc; e[2] = ex_acc;
                }
            }
        }
        P.ladder_rng[lad * 4] = LS.r.s0; P.ladder_rng[lad * 4 + 1] = LS.r.s1; P.ladder_rng[lad * 4 + 2] = LS.r.s2; P.ladder_rng[lad * 4 + 3] = LS.r.s3;
        if (DBG) {
            if (P.replay) P.ladder_cursor[lad] = LS.pos;
            if (LS.exhausted) atomicMin(P.error, (int)PTMC_ERR_REPLAY);
        }
    }
    __syncwarp();
    for (int slot = lane; slot < T; slot += 32) {
        const i64 rep = (i64)lad * T + slot;
        i64* cn = P.counters + rep * PTMC_NCOUNT;
        if (idx % P.n_adjust == 0) {
            const double rate = xdiv((double)cn[PTMC_C_ATOM_ADJ], (double)((i64)P.n_adjust * P.n_atom_moves));
            double dmax = P.max_displ[rep * 4];
            if (rate < P.min_acc) dmax = xmul(dmax, 0.9);
            else if (rate > P.max_acc) dmax = xmul(dmax, 1.1);
            P.max_displ[rep * 4] = dmax;
            cn[PTMC_C_ATOM_ADJ] = 0;
        }
        if (sample && idx % P.mc_sample == 0) {
            const double e = en_c[(i64)lad * T + perm[slot]];
            const int nb = P.n_bin;
            const double en_min = P.en_min[lad];
            const double delta_en = xdiv(xsub(P.en_max[lad], en_min), (double)(nb - 1));
            P.ham[rep * 2] = xadd(P.ham[rep * 2], e);
            P.ham[rep * 2 + 1] = xadd(P.ham[rep * 2 + 1], xmul(e, e));
            const double hi = xadd(xdiv(xsub(e, en_min), delta_en), 1.0);
            int h = hi < 1.0 ? 1 : (hi > (double)nb ? nb + 2 : (int)floor(hi) + 1);
            if (hi != hi) h = nb + 2;
            P.hist[(size_t)rep * (nb + 2) + (h - 1)] += 1;
        }
    }
}

// slot-ordered copies of the configuration-side state for ptmc_download_states
__global__ void runner_gather_kernel(const DevParams P, const double* __restrict__ pos_c, const double* __restrict__ en_c) {
    const i64 rep = blockIdx.x;
    const int T = P.n_traj, N = P.n_atoms;
    const i64 cfg = (rep / T) * T + P.perm_g[rep];
    for (int k = threadIdx.x; k < 3 * N; k += blockDim.x) P.pos[rep * 3 * N + k] = pos_c[cfg * 3 * N + k];
    if (threadIdx.x == 0) P.en_tot[rep] = en_c[cfg];
}
// compact g matrix / per-atom energies of the configuration in one slot (tests): g_out [N][n_sym], e_out [N]
__global__ void runner_export_kernel(const DevParams P, const RunnerDev R, i64 rep, double* g_out, double* e_out) {
    const int T = P.n_traj, N = P.n_atoms;
    const i64 cfg = (rep / T) * T + P.perm_g[rep];
    for (int k = threadIdx.x; k < N * R.n_sym; k += blockDim.x) g_out[k] = 0.0;
    __syncthreads();
    for (int k = threadIdx.x; k < N * R.n_act; k += blockDim.x) {
        const int a = k / R.n_act, q = k - a * R.n_act;
        g_out[a * R.n_sym + R.act_row[q]] = R.g_c[cfg * N * R.n_act + k];
    }
    for (int k = threadIdx.x; k < N; k += blockDim.x) e_out[k] = R.en_atom_c[cfg * N + k];
}

// -------------------------------------------------------------------------------------- host side --
size_t runner_smem_bytes(int N, const RunnerDev& R) { return (size_t)rn_smem(N, R).total * sizeof(double); }

typedef void (*rn_moves_fn)(const DevParams, const RunnerDev, const i64);
// specialised on the reference's standard angular set when the potential has exactly that structure
static bool is_standard_set(const RunnerDev& R) {
    if (R.present != RN_PRESENT_V || R.n_eta != 7 || R.n_lz != 8) return false;
    for (int z = 0; z < 8; ++z) if (R.lz_nsq[z] != (int)((RN_NSQ_V >> (4 * z)) & 15u)) return false;
    // the squaring chains of RN_ZSEQ_V: every slot has the lambda of its chain's root
    int root = 0;
    for (int q = 0; q < 8; ++q) {
        const int z = (int)((RN_ZSEQ_V >> (4 * q)) & 15u);
        if ((RN_ZROOT_V >> q) & 1u) root = z;
        else if (R.lz_lambda[z] != R.lz_lambda[root]) return false;
    }
    return true;
}
static rn_moves_fn pick_moves(const RunnerDev& R, bool dbg) {
    if (is_standard_set(R))
        return dbg ? (rn_moves_fn)runner_moves_kernel<true, RN_PRESENT_V, RN_NSQ_V> : (rn_moves_fn)runner_moves_kernel<false, RN_PRESENT_V, RN_NSQ_V>;
    return dbg ? (rn_moves_fn)runner_moves_kernel<true, 0ull, 0u> : (rn_moves_fn)runner_moves_kernel<false, 0ull, 0u>;
}

cudaError_t runner_prepare(int N, const RunnerDev& R) {
    const int bytes = (int)runner_smem_bytes(N, R);
    for (int dbg = 0; dbg < 2; ++dbg) {
        rn_moves_fn f = pick_moves(R, dbg != 0);
        if (!f) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute((const void*)runner_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
void runner_launch_init(const DevParams& P, const RunnerDev& R, i64 n_rep, cudaStream_t st) {
    runner_init_kernel<<<(unsigned)n_rep, RN_THREADS, runner_smem_bytes(P.n_atoms, R), st>>>(P, R);
}
// one cycle = moves for every slot + end-of-cycle step for every ladder; returns the number of launches
int runner_launch_cycle(const DevParams& P, const RunnerDev& R, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st) {
    rn_moves_fn f = pick_moves(R, dbg);
    f<<<(unsigned)(n_ladders * P.n_traj), RN_THREADS, runner_smem_bytes(P.n_atoms, R), st>>>(P, R, cyc_in_launch);
    runner_launch_cycle_end(P, R.en_c, n_ladders, idx, cyc_in_launch, sample, dbg, st);
    return 2;
}
void runner_launch_cycle_end(const DevParams& P, const double* en_c, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st) {
    if (dbg) runner_cycle_end_kernel<true><<<(unsigned)n_ladders, 32, 0, st>>>(P, en_c, idx, cyc_in_launch, sample);
    else runner_cycle_end_kernel<false><<<(unsigned)n_ladders, 32, 0, st>>>(P, en_c, idx, cyc_in_launch, sample);
}
void runner_launch_gather(const DevParams& P, const RunnerDev& R, i64 n_rep, cudaStream_t st) {
    runner_launch_gather_ptr(P, R.pos_c, R.en_c, n_rep, st);
}
void runner_launch_gather_ptr(const DevParams& P, const double* pos_c, const double* en_c, i64 n_rep, cudaStream_t st) {
    runner_gather_kernel<<<(unsigned)n_rep, 64, 0, st>>>(P, pos_c, en_c);
}
void runner_launch_export(const DevParams& P, const RunnerDev& R, i64 rep, double* g_out, double* e_out, cudaStream_t st) {
    runner_export_kernel<<<1, 256, 0, st>>>(P, R, rep, g_out, e_out);
}

}  // namespace ptmc
