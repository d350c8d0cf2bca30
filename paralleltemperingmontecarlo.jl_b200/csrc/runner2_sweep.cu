// runner2_sweep.cu — the sweep for the two-species RuNNer potential with atom moves and atom swaps (scripts/brass55.jl):
// energy_update!(::RuNNerPotential2Atom) src/EnergyEvaluation.jl:1623-1643 and swap_energy_update :1437-1448, i.e.
// calc_delta_matrix / calc_swap_matrix (MachineLearningPotential/DeltaMatrix.jl:256-316, 583-632) + one forward pass per species
// (:1418-1433), fused with move generation (atom_displacement, swap_atoms src/MCMoves.jl:363-370), the Metropolis test and the
// commit (swap_atom_config! / swap_move_config! src/swap_config.jl:40-51,241-276) in one kernel, one CTA per replica.
//
// Same work decomposition as runner_sweep.cu (exponential factorised per pair, angular factor shared by the functions with equal
// (lambda, zeta), pass A = centre-j terms / pass B = centre-i terms side by side, DMMA hidden layers), with what two species add:
//  * a term lands in the row of its neighbour class (radial: species of the neighbour; angular: number of species-2 atoms among
//    the two neighbours), so the loops run class by class (atoms are ordered by species) and flush their accumulators between
//    classes; every accumulator still has one writer, results are reproducible bit for bit;
//  * all 2 n_rad + 3 n_ang rows are live: the two g matrices (77 KB) take the place of the pair tables in shared memory, the pair
//    tables (d2, 1/d, f, exp(-eta d2) f) are read from HBM / L1;
//  * an atom swap exchanges the positions of a1 (species 1) and a2 (species 2).  No distance changes, the rows of the two atoms
//    trade places: it runs as two masked "virtual moves" (a1 takes the table row of a2 with a2 left out, and vice versa) plus the
//    triples that hold both atoms, where only the roles of a1 and a2 change (DeltaMatrix.jl:532-577).
#include "runner_dev.cuh"

namespace ptmc {

constexpr int RN2_RED = 2 * RN_MAX_RAD + 3 * RN_MAX_ANG;      // cross-warp reduction slots: radial [r][class 0..1], angular [q][class 0..2]

struct Rn2Smem {          // offsets in doubles into the dynamic shared memory
    int gs;               // row stride of the g matrices (odd: the DMMA loads of eight atoms spread over the banks)
    int pos, g0, g1, ea0, ea1, rn, ro, red, net, h1, h2, misc, ptab, total;
};
__host__ __device__ inline Rn2Smem rn2_smem(int N, const Runner2Dev& R) {
    Rn2Smem s; int o = 0;
    s.gs = R.n_sym | 1;
    s.pos = o; o += 3 * N;
    s.g0 = o; o += N * s.gs; s.g1 = o; o += N * s.gs;
    s.ea0 = o; o += N; s.ea1 = o; o += N;
    s.rn = o; o += (3 + R.n_eta) * N;            // new row: d2, 1/d, f, ef[n_eta]
    s.ro = o; o += (3 + R.n_eta) * N;            // old row
    s.red = o; o += (RN_THREADS / 32) * RN2_RED;
    s.net = o; o += 2 * R.net_len;
    s.h1 = o; o += N * R.h1; s.h2 = o; o += N * R.h2;
    s.misc = o; o += 16;
    s.ptab = o; o += (N * (N - 1) / 2 + 1) / 2;   // int (j << 16 | k) per pair, sorted by class
    s.total = o;
    return s;
}

// One dense layer for all atoms, the weights of an atom's species: tiles of 8 output nodes x 8 atoms of ONE species per warp and round.
__device__ __forceinline__ void rn2_layer_dmma(const double* net, int net_len, int w_off, int b_off, int n_out, int n_in, int act,
                                               const double* in, int in_stride, double* out, int N, int n1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int mt = (n_out + 7) >> 3, nt1 = (n1 + 7) >> 3, nt2 = (N - n1 + 7) >> 3, kt = (n_in + 3) >> 2;
    for (int t = warp; t < mt * (nt1 + nt2); t += RN_THREADS / 32) {
        const int m = t % mt, n = t / mt;
        const int sp = n >= nt1 ? 1 : 0;
        const int base = sp ? n1 + 8 * (n - nt1) : 8 * n, end = sp ? N : n1;
        const double* W = net + sp * net_len + w_off;
        const double* b = net + sp * net_len + b_off;
        const int o = 8 * m + gid, ab = base + gid;
        const bool ok_o = o < n_out, ok_a = ab < end;
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
        const int a0 = base + 2 * tig;
        if (ok_o) {
            if (a0 < end) out[a0 * n_out + o] = rn_act(d0, act);
            if (a0 + 1 < end) out[(a0 + 1) * n_out + o] = rn_act(d1, act);
        }
    }
}
// forward_pass of both species on the g matrix in shared memory; returns, in every thread, the sum of the per-atom energies
// (calc_new_runner_energy! src/EnergyEvaluation.jl:1418-1433)
__device__ double rn2_forward(const Runner2Dev& R, int N, const double* g, int gs, const double* net, double* h1, double* h2, double* eatom, double* red) {
    const int tid = threadIdx.x;
    const int oW1 = 0, ob1 = R.h1 * R.n_sym, oW2 = ob1 + R.h1, ob2 = oW2 + R.h2 * R.h1, oW3 = ob2 + R.h2, ob3 = oW3 + R.h2;
    rn2_layer_dmma(net, R.net_len, oW1, ob1, R.h1, R.n_sym, R.act1, g, gs, h1, N, R.n1);
    __syncthreads();
    rn2_layer_dmma(net, R.net_len, oW2, ob2, R.h2, R.h1, R.act2, h1, R.h1, h2, N, R.n1);
    __syncthreads();
    const int a = tid / RN_SUB, q = tid % RN_SUB;
    const bool on = a < N;
    const double* ns = net + ((on && a >= R.n1) ? R.net_len : 0);
    double part = 0.0;
    if (on) for (int o = q; o < R.h2; o += RN_SUB) part = fma(ns[oW3 + o], h2[a * R.h2 + o], part);
#pragma unroll
    for (int m = 1; m < RN_SUB; m <<= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    double e = 0.0;
    if (on && q == 0) {
        e = rn_act(part + ns[ob3], R.act3);
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
// set_variables(::RuNNerPotential2Atom) src/EnergyEvaluation.jl:1823-1857 + initialise_energy :1748-1768 for one replica per CTA:
// pair tables, the g matrix from scratch (total_symm_calc for two species, SymmFunc.jl:511-545), the first energies.
__global__ void __launch_bounds__(RN_THREADS) runner2_init_kernel(const DevParams P, const Runner2Dev R) {
    extern __shared__ double sm[];
    const int N = P.n_atoms, tid = threadIdx.x, n1 = R.n1;
    const i64 rep = blockIdx.x;
    const Rn2Smem L = rn2_smem(N, R);
    const int gs = L.gs, nr2 = 2 * R.n_rad;
    double* pos = sm + L.pos; double* g = sm + L.g0; double* net = sm + L.net;
    const int NT = N * (N - 1) / 2;
    double* D2 = R.pair_c + rep * (size_t)(3 + R.n_eta) * NT;
    double* INV = D2 + NT; double* F = INV + NT; double* EF = F + NT;
    for (int k = tid; k < 3 * N; k += RN_THREADS) { pos[k] = P.pos[rep * 3 * N + k]; R.pos_c[rep * 3 * N + k] = pos[k]; }
    for (int k = tid; k < 2 * R.net_len; k += RN_THREADS) net[k] = R.net[k];
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
    // one thread per (atom a, symmetry function): every contribution with centre a, summed per neighbour class, then scaled
    // (calc_symm_vals! SymmFunc.jl:320-470)
    for (int w = tid; w < N * R.n_rad; w += RN_THREADS) {
        const int a = w / R.n_rad, r = w - a * R.n_rad, sa = a >= n1 ? 1 : 0;
        double s[2] = {0.0, 0.0};
        for (int b = 0; b < N; ++b) {
            if (b == a) continue;
            const int t = rn_tri(a, b, N);
            const double f = F[t];
            if (f != 0.0 && f != 1.0) s[b >= n1 ? 1 : 0] += f * exp(-R.rad_eta[r] * D2[t]);   // calc_one_symm_val :229-231
        }
        for (int c = 0; c < 2; ++c) g[a * gs + 2 * r + c] = s[c] * R.rad_norm[r][2 * sa + c] + R.rad_off[r][2 * sa + c];
    }
    for (int w = tid; w < N * R.n_ang; w += RN_THREADS) {
        const int a = w / R.n_ang, q = w - a * R.n_ang, sa = a >= n1 ? 1 : 0;
        const int e = R.ang_eta_idx[q], z = R.ang_lz_idx[q];
        const double* EFe = EF + e * NT;
        const double la = R.lz_lambda[z], ze = R.lz_zeta[z];
        const int zi = R.lz_zint[z];
        double s[3] = {0.0, 0.0, 0.0};
        for (int b = 0; b < N; ++b) {
            if (b == a) continue;
            const int tab = rn_tri(a, b, N);
            const double dab = D2[tab], iab = INV[tab], eab = EFe[tab];
            const int sb = b >= n1 ? 1 : 0;
            for (int c = b + 1; c < N; ++c) {
                if (c == a) continue;
                const int tac = rn_tri(a, c, N), tbc = rn_tri(b, c, N);
                const double ca = (dab + D2[tac] - D2[tbc]) * 0.5 * iab * INV[tac];
                s[sb + (c >= n1 ? 1 : 0)] += eab * EFe[tac] * EFe[tbc] * rn_theta(ca, la, ze, zi);
            }
        }
        for (int c = 0; c < 3; ++c) g[a * gs + nr2 + 3 * q + c] = s[c] * R.ang_tpz[q][3 * sa + c] + R.ang_off[q][3 * sa + c];
    }
    __syncthreads();
    double* ea = sm + L.ea0;
    const double e = rn2_forward(R, N, g, gs, net, sm + L.h1, sm + L.h2, ea, sm + L.red);
    for (int k = tid; k < N * R.n_sym; k += RN_THREADS) { const int a = k / R.n_sym, q = k - a * R.n_sym; R.g_c[rep * N * R.n_sym + k] = g[a * gs + q]; }
    for (int k = tid; k < N; k += RN_THREADS) R.en_atom_c[rep * N + k] = ea[k];
    if (tid == 0) { R.en_c[rep] = e; P.en_tot[rep] = e; }
}

// ------------------------------------------------------------------------------------------- moves --
// Everything one "virtual move" of atom i needs.  src < 0: i moves to (tx, ty, tz).  src >= 0 (half of a swap): i takes the place of
// atom src, i.e. its new row is the table row of src; `ex` (= src) is left out of every sum (its terms do not change, or belong to
// the triples that hold both swapped atoms).
struct Rn2Move { int i, ex, src; double tx, ty, tz; };

template <bool SPEC, unsigned long long PRESENT, unsigned NSQ>
__device__ __forceinline__ void rn2_submove(const Runner2Dev& R, const Rn2Smem& L, double* sm, int N, const Rn2Move mv, double* ng,
                                            const double* D2, const double* INV, const double* F, const double* EF, const int* ptab, const int (&seg)[4]) {
    const unsigned long long present = SPEC ? PRESENT : R.present;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = R.n1, gs = L.gs, nr2 = 2 * R.n_rad, NT = N * (N - 1) / 2;
    const double* pos = sm + L.pos;
    double* rn_d2 = sm + L.rn; double* rn_inv = rn_d2 + N; double* rn_f = rn_inv + N; double* rn_ef = rn_f + N;
    double* ro_d2 = sm + L.ro; double* ro_inv = ro_d2 + N; double* ro_f = ro_inv + N; double* ro_ef = ro_f + N;
    double* red = sm + L.red;
    const int i = mv.i, ex = mv.ex;
    const int spi = i >= n1 ? 1 : 0;
    // =============================================== phase 1: new / old row of atom i, radial functions ==
    double prad[2][RN_MAX_RAD];
#pragma unroll
    for (int r = 0; r < RN_MAX_RAD; ++r) { prad[0][r] = 0.0; prad[1][r] = 0.0; }
    if (tid < N * RN_SUB) {
        const int j = tid / RN_SUB, ls = tid % RN_SUB;
        const bool self = j == i, skip = self || j == ex;
        const int spj = j >= n1 ? 1 : 0;
        const int tij = self ? 0 : rn_tri(i, j, N);
        int tsrc = 0;
        double d2n, invn, fn;
        if (mv.src < 0) {
            const double x = xsub(mv.tx, pos[3 * j]), y = xsub(mv.ty, pos[3 * j + 1]), z = xsub(mv.tz, pos[3 * j + 2]);
            d2n = self ? 0.0 : xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));     // new_dist2_vec[index] = 0
            const double rr = sqrt(d2n);
            fn = self ? 1.0 : rn_cutoff(rr, R.r_cut);                             // cutoff_function.(sqrt.(new_dist2_vec)) :1357-1359
            invn = self ? 0.0 : 1.0 / rr;
        } else {
            tsrc = self ? 0 : (j == mv.src ? tij : rn_tri(mv.src, j, N));         // the distance between the two swapped atoms stays
            d2n = self ? 0.0 : D2[tsrc]; invn = self ? 0.0 : INV[tsrc]; fn = self ? 1.0 : F[tsrc];
        }
        const double d2o = self ? 0.0 : D2[tij], fo = self ? 1.0 : F[tij];
        if (ls == 0) {
            rn_d2[j] = d2n; rn_inv[j] = invn; rn_f[j] = fn;
            ro_d2[j] = d2o; ro_inv[j] = self ? 0.0 : INV[tij]; ro_f[j] = fo;
        }
        for (int e = ls; e < R.n_eta; e += RN_SUB) {
            rn_ef[e * N + j] = self ? 0.0 : (mv.src < 0 ? exp(-R.eta[e] * d2n) * fn : EF[e * NT + tsrc]);
            ro_ef[e * N + j] = self ? 0.0 : EF[e * NT + tij];
        }
        if (!skip) {
            // calc_delta_symm_val! (radial) DeltaMatrix.jl:69-122 / calc_swap_symm_val :322-371: the neighbour j gains in the row of
            // i's species, the centre i in the row of j's species
#pragma unroll
            for (int r = 0; r < RN_MAX_RAD; ++r) {
                if (r < R.n_rad && (RN_SUB - 1 - r % RN_SUB) == ls) {
                    const double gn = exp(-R.rad_eta[r] * d2n) * fn - exp(-R.rad_eta[r] * d2o) * fo;      // new_radial_symm_val! :23-25
                    ng[j * gs + 2 * r + spi] += gn * R.rad_norm[r][2 * spj + spi];
                    prad[0][r] = spj ? 0.0 : gn;
                    prad[1][r] = spj ? gn : 0.0;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RN_MAX_RAD; ++r) {
        if (r < R.n_rad) {
            const double v0 = group_sum<32>(prad[0][r], 0xffffffffu), v1 = group_sum<32>(prad[1][r], 0xffffffffu);
            if (lane == 0) { red[warp * RN2_RED + 2 * r] = v0; red[warp * RN2_RED + 2 * r + 1] = v1; }
        }
    }
    __syncthreads();
    if (tid < nr2) {
        double v = 0.0;
        for (int w = 0; w < RN_THREADS / 32; ++w) v += red[w * RN2_RED + tid];
        ng[i * gs + tid] += v * R.rad_norm[tid >> 1][2 * spi + (tid & 1)];
    }
    // ======================================== phase 2: angular functions, triples (i, j, k) ==========
    // calc_delta_symm_val! (angular) DeltaMatrix.jl:124-250 / calc_swap_symm_val :373-531
    double acc[RN_MAX_ETA][RN_MAX_LZ];
#pragma unroll
    for (int e = 0; e < RN_MAX_ETA; ++e)
#pragma unroll
        for (int z = 0; z < RN_MAX_LZ; ++z) acc[e][z] = 0.0;
    const int nwa = (N * RN_SUB + 31) >> 5;
    if (warp < nwa) {
        // ---- pass A: centre j, thread (j, s) takes k = k0 + s, k0 + s + RN_SUB, ... of one species at a time; class of the
        // term = (species of i) + (species of k)
        const int j = tid / RN_SUB, s = tid % RN_SUB;
        const bool jok = j < N && j != i && j != ex;
        const int spj = (j < N && j >= n1) ? 1 : 0;
        const double an = jok ? rn_d2[j] : 0.0, ao = jok ? ro_d2[j] : 0.0, ijn = jok ? rn_inv[j] : 0.0, ijo = jok ? ro_inv[j] : 0.0;
        for (int sk = 0; sk < 2; ++sk) {
            const int k0 = sk ? n1 : 0, k1 = sk ? N : n1;
            if (jok) {
                for (int k = k0 + s; k < k1; k += RN_SUB) {
                    if (k == i || k == j || k == ex) continue;
                    const int tjk = rn_tri(j, k, N);
                    const double c = D2[tjk], ijk = INV[tjk];
                    const double cn = (an + c - rn_d2[k]) * 0.5 * ijn * ijk, co = (ao + c - ro_d2[k]) * 0.5 * ijo * ijk;
                    double epn[RN_MAX_ETA], epo[RN_MAX_ETA];
#pragma unroll
                    for (int e = 0; e < RN_MAX_ETA; ++e) {
                        epn[e] = 0.0; epo[e] = 0.0;
                        if (SPEC ? ((PRESENT >> (e * RN_MAX_LZ)) & 0xffull) != 0 : e < R.n_eta) {
                            const double ejk = EF[e * NT + tjk];
                            epn[e] = rn_ef[e * N + j] * rn_ef[e * N + k] * ejk;
                            epo[e] = ro_ef[e * N + j] * ro_ef[e * N + k] * ejk;
                        }
                    }
                    rn_angular_acc<SPEC, PRESENT, NSQ>(R, present, cn, co, epn, epo, acc);
                }
            }
            const int cls = spi + sk;
#pragma unroll
            for (int e = 0; e < RN_MAX_ETA; ++e)
#pragma unroll
                for (int z = 0; z < RN_MAX_LZ; ++z)
                    if (present >> (e * RN_MAX_LZ + z) & 1ull) {
                        double v = acc[e][z];
#pragma unroll
                        for (int m = 1; m < RN_SUB; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                        const int q = R.qmap[e * RN_MAX_LZ + z];
                        if (s == 0 && jok) ng[j * gs + nr2 + 3 * q + cls] += v * R.ang_tpz[q][3 * spj + cls];
                        acc[e][z] = 0.0;
                    }
        }
    } else {
        // ---- pass B: centre i, the pairs (j < k) of one class at a time dealt round-robin to the threads of the remaining warps;
        // class of the term = (species of j) + (species of k)
        const int tb = tid - 32 * nwa, nb = RN_THREADS - 32 * nwa;
        for (int cls = 0; cls < 3; ++cls) {
            for (int p = seg[cls] + tb; p < seg[cls + 1]; p += nb) {
                const int jk = ptab[p], pj = jk >> 16, pk = jk & 0xffff;
                if (pj == i || pk == i || pj == ex || pk == ex) continue;
                const int t = rn_tri(pj, pk, N);
                const double c = D2[t];
                const double an = rn_d2[pj], ao = ro_d2[pj], bn = rn_d2[pk], bo = ro_d2[pk];
                const double cn = (an + bn - c) * 0.5 * rn_inv[pj] * rn_inv[pk], co = (ao + bo - c) * 0.5 * ro_inv[pj] * ro_inv[pk];
                double epn[RN_MAX_ETA], epo[RN_MAX_ETA];
#pragma unroll
                for (int e = 0; e < RN_MAX_ETA; ++e) {
                    epn[e] = 0.0; epo[e] = 0.0;
                    if (SPEC ? ((PRESENT >> (e * RN_MAX_LZ)) & 0xffull) != 0 : e < R.n_eta) {
                        const double ejk = EF[e * NT + t];
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
                        if (lane == 0) red[warp * RN2_RED + 2 * RN_MAX_RAD + 3 * R.qmap[e * RN_MAX_LZ + z] + cls] = w;
                        acc[e][z] = 0.0;
                    }
        }
    }
    __syncthreads();
    if (tid < 3 * R.n_ang) {
        double v = 0.0;
        for (int w = nwa; w < RN_THREADS / 32; ++w) v += red[w * RN2_RED + 2 * RN_MAX_RAD + tid];
        ng[i * gs + nr2 + tid] += v * R.ang_tpz[tid / 3][3 * spi + tid % 3];
    }
    __syncthreads();
}

// mc_step! for one replica slot: n_steps x mc_move! (src/MCRun.jl:142-182) with MoveStrategy(::NNVT) (src/Ensembles.jl:270-282)
template <bool DBG, unsigned long long PRESENT, unsigned NSQ>
__global__ void __launch_bounds__(RN_THREADS, 1)
runner2_moves_kernel(const DevParams P, const Runner2Dev R, const i64 cyc_in_launch) {
    static_assert(RN_THREADS >= 384, "the two passes of the angular update run side by side");
    extern __shared__ double sm[];
    constexpr bool BATCH = !DBG;
    constexpr bool SPEC = PRESENT != 0;
    const int N = P.n_atoms, T = P.n_traj, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = R.n1, n2 = N - n1;
    const i64 rep = blockIdx.x;
    const int lad = (int)(rep / T), slot = (int)(rep - (i64)lad * T);
    const i64 cfg = (i64)lad * T + P.perm_g[rep];
    const Rn2Smem L = rn2_smem(N, R);
    const int gs = L.gs;
    double* pos = sm + L.pos;
    double* gbuf[2] = {sm + L.g0, sm + L.g1};
    double* eabuf[2] = {sm + L.ea0, sm + L.ea1};
    double* rn_d2 = sm + L.rn; double* rn_inv = rn_d2 + N; double* rn_f = rn_inv + N; double* rn_ef = rn_f + N;
    double* red = sm + L.red; double* net = sm + L.net;
    double* misc = sm + L.misc;        // 0 atom (or a1), 1-3 trial, 4 u_acc, 7 kind (0 atom move, 1 swap), 8 a2, 9 trace index
    const int NT = N * (N - 1) / 2;
    double* D2 = R.pair_c + cfg * (size_t)(3 + R.n_eta) * NT;
    double* INV = D2 + NT; double* F = INV + NT; double* EF = F + NT;
    int* ptab = reinterpret_cast<int*>(sm + L.ptab);
    // pairs sorted by class: (1,1) | (1,2) | (2,2)
    const int seg[4] = {0, n1 * (n1 - 1) / 2, n1 * (n1 - 1) / 2 + n1 * n2, NT};
    for (int p = tid; p < N * N; p += RN_THREADS) {
        const int a = p / N, b = p - a * N;
        if (a >= b) continue;
        int o;
        if (b < n1) o = seg[0] + a * n1 - (a * (a + 1)) / 2 + (b - a - 1);
        else if (a < n1) o = seg[1] + a * n2 + (b - n1);
        else { const int a2 = a - n1, b2 = b - n1; o = seg[2] + a2 * n2 - (a2 * (a2 + 1)) / 2 + (b2 - a2 - 1); }
        ptab[o] = a << 16 | b;
    }
    int cur = 0;
    for (int k = tid; k < 3 * N; k += RN_THREADS) pos[k] = R.pos_c[cfg * 3 * N + k];
    for (int k = tid; k < N * R.n_sym; k += RN_THREADS) { const int a = k / R.n_sym, q = k - a * R.n_sym; gbuf[0][a * gs + q] = R.g_c[cfg * N * R.n_sym + k]; }
    for (int k = tid; k < N; k += RN_THREADS) eabuf[0][k] = R.en_atom_c[cfg * N + k];
    for (int k = tid; k < 2 * R.net_len; k += RN_THREADS) net[k] = R.net[k];
    double en = R.en_c[cfg];
    const double beta = P.betas[slot];
    const double dmax = P.max_displ[rep * 4];
    const double R2 = P.box[rep * 2];
    int n_acc = 0;
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
        // ============================================================ phase 0: draws (warp 0) ==
        if (warp == 0) {
            const unsigned full = 0xffffffffu;
            int index;
            if (BATCH) { B.begin(); index = (int)__umul64hi(B.template raw<0>(full), (u64)P.n_steps) + 1; }
            else index = S.range((u64)P.n_steps);
            if (index <= P.n_atom_moves) {
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
                if (lane == 0) { misc[0] = (double)i; misc[1] = tx; misc[2] = ty; misc[3] = tz; misc[4] = u; misc[7] = 0.0; misc[9] = (double)index; }
            } else {
                // swap_atoms src/MCMoves.jl:363-370: i1 = rand(1:N1), i2 = rand((N1+1):N); BATCH8: positions 1 and 2 of the batch
                int i1, i2;
                if (BATCH) {
                    i1 = (int)__umul64hi(B.template raw<1>(full), (u64)n1) + 1;
                    i2 = n1 + (int)__umul64hi(B.template raw<2>(full), (u64)n2) + 1;
                } else {
                    i1 = S.range((u64)n1);
                    i2 = n1 + S.range((u64)n2);
                }
                const double u = BATCH ? B.template rand<4>(full) : S.rand();
                if (lane == 0) { misc[0] = (double)(i1 - 1); misc[8] = (double)(i2 - 1); misc[4] = u; misc[7] = 1.0; misc[9] = (double)(index + (i1 << 10) + (i2 << 20)); }
            }
            if (DBG) { if (S.exhausted) err = PTMC_ERR_REPLAY; }
        }
        // new_g_matrix starts as a copy of g_matrix; every change of the move is added to it
        for (int k = tid; k < N * gs; k += RN_THREADS) ng[k] = g[k];
        __syncthreads();
        const bool is_swap = misc[7] != 0.0;
        const int i = (int)misc[0];
        const int a2 = is_swap ? (int)misc[8] : -1;
        const double tx = misc[1], ty = misc[2], tz = misc[3];
        if (!is_swap) {
            const Rn2Move mv = {i, -1, -1, tx, ty, tz};
            rn2_submove<SPEC, PRESENT, NSQ>(R, L, sm, N, mv, ng, D2, INV, F, EF, ptab, seg);
        } else {
            const Rn2Move m1 = {i, a2, a2, 0.0, 0.0, 0.0}, m2 = {a2, i, i, 0.0, 0.0, 0.0};
            rn2_submove<SPEC, PRESENT, NSQ>(R, L, sm, N, m1, ng, D2, INV, F, EF, ptab, seg);
            rn2_submove<SPEC, PRESENT, NSQ>(R, L, sm, N, m2, ng, D2, INV, F, EF, ptab, seg);
            // triples that hold both swapped atoms (DeltaMatrix.jl:532-577): the geometry stays, a1 and a2 trade the angles centred
            // on their two sites; thread (q, species of the third atom c) walks c in order
            if (tid < 2 * R.n_ang) {
                const int q = tid >> 1, sc = tid & 1, a1 = i;
                const int e = R.ang_eta_idx[q], z = R.ang_lz_idx[q];
                const int t12 = rn_tri(a1, a2, N);
                const double d12 = D2[t12], i12 = INV[t12], e12 = EF[e * NT + t12];
                double s1 = 0.0, s2 = 0.0;
                for (int c = sc ? n1 : 0; c < (sc ? N : n1); ++c) {
                    if (c == a1 || c == a2) continue;
                    const int t1c = rn_tri(a1, c, N), t2c = rn_tri(a2, c, N);
                    const double ep = e12 * EF[e * NT + t1c] * EF[e * NT + t2c];
                    const double c1 = (d12 + D2[t1c] - D2[t2c]) * 0.5 * i12 * INV[t1c], c2 = (d12 + D2[t2c] - D2[t1c]) * 0.5 * i12 * INV[t2c];
                    const double v1 = ep * rn_theta_n(c1, R.lz_lambda[z], R.lz_zeta[z], R.lz_nsq[z]);
                    const double v2 = ep * rn_theta_n(c2, R.lz_lambda[z], R.lz_zeta[z], R.lz_nsq[z]);
                    s1 += v2 - v1; s2 += v1 - v2;
                }
                ng[a1 * gs + 2 * R.n_rad + 3 * q + 1 + sc] += s1 * R.ang_tpz[q][1 + sc];        // a1: species 1, neighbours a2 (species 2) and c
                ng[a2 * gs + 2 * R.n_rad + 3 * q + sc] += s2 * R.ang_tpz[q][3 + sc];            // a2: species 2, neighbours a1 (species 1) and c
            }
            __syncthreads();
        }
        // ======================================== phase 3: network forward passes on new_g_matrix ==========
        const double new_en = rn2_forward(R, N, ng, gs, net, sm + L.h1, sm + L.h2, eabuf[cur ^ 1], red);
        // ============================================================ phase 4: acc_test! src/MCRun.jl:122-132 ==
        const double de = xsub(new_en, en);
        const bool acc = metropolis_accept(-xmul(de, beta), misc[4]);
        {
            if (tracing && tid == 0) {
                size_t o = ((((size_t)lad * P.trace_cap + P.trace_base + cyc_in_launch) * T + slot) * P.n_steps) + step;
                P.tr_index[o] = (int)misc[9]; P.tr_de[o] = de; P.tr_acc[o] = acc ? 1 : 0;
            }
        }
        __syncthreads();       // everybody has read misc[] and the row arrays are final
        if (acc) {
            if (!is_swap) {
                // swap_atom_config! src/swap_config.jl:40-51 + swap_vars!(::NNPVariables2a) :223-235
                if (tid == 0) { pos[3 * i] = tx; pos[3 * i + 1] = ty; pos[3 * i + 2] = tz; }
                if (tid < N && tid != i) {
                    const int j = tid, tij = rn_tri(i, j, N);
                    D2[tij] = rn_d2[j]; INV[tij] = rn_inv[j]; F[tij] = rn_f[j];
                    for (int e = 0; e < R.n_eta; ++e) EF[e * NT + tij] = rn_ef[e * N + j];
                }
                n_acc += 1;
            } else {
                // swap_move_config! src/swap_config.jl:241-276: positions, rows and columns of dist2_mat / f_matrix (no counter)
                if (tid < 3) { const double t = pos[3 * i + tid]; pos[3 * i + tid] = pos[3 * a2 + tid]; pos[3 * a2 + tid] = t; }
                if (tid < N && tid != i && tid != a2) {
                    const int t1 = rn_tri(i, tid, N), t2 = rn_tri(a2, tid, N);
                    for (int k = 0; k < 3 + R.n_eta; ++k) { const double t = D2[k * NT + t1]; D2[k * NT + t1] = D2[k * NT + t2]; D2[k * NT + t2] = t; }
                }
            }
            en = new_en;
            cur ^= 1;     // g_matrix <-> new_g_matrix, en_atom_vec <-> new_en_atom
        }
        __syncthreads();
    }

    // ---- store ------------------------------------------------------------------------------------------
    for (int k = tid; k < 3 * N; k += RN_THREADS) R.pos_c[cfg * 3 * N + k] = pos[k];
    for (int k = tid; k < N * R.n_sym; k += RN_THREADS) { const int a = k / R.n_sym, q = k - a * R.n_sym; R.g_c[cfg * N * R.n_sym + k] = gbuf[cur][a * gs + q]; }
    for (int k = tid; k < N; k += RN_THREADS) R.en_atom_c[cfg * N + k] = eabuf[cur][k];
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

// g matrix [N][n_sym] / per-atom energies of the configuration in one slot (tests)
__global__ void runner2_export_kernel(const DevParams P, const Runner2Dev R, i64 rep, double* g_out, double* e_out) {
    const int T = P.n_traj, N = P.n_atoms;
    const i64 cfg = (rep / T) * T + P.perm_g[rep];
    for (int k = threadIdx.x; k < N * R.n_sym; k += blockDim.x) g_out[k] = R.g_c[cfg * N * R.n_sym + k];
    for (int k = threadIdx.x; k < N; k += blockDim.x) e_out[k] = R.en_atom_c[cfg * N + k];
}

// -------------------------------------------------------------------------------------- host side --
size_t runner2_smem_bytes(int N, const Runner2Dev& R) { return (size_t)rn2_smem(N, R).total * sizeof(double); }

typedef void (*rn2_moves_fn)(const DevParams, const Runner2Dev, const i64);
static bool is_standard_set2(const Runner2Dev& R) {
    if (R.present != RN_PRESENT_V || R.n_eta != 7 || R.n_lz != 8) return false;
    for (int z = 0; z < 8; ++z) if (R.lz_nsq[z] != (int)((RN_NSQ_V >> (4 * z)) & 15u)) return false;
    int root = 0;
    for (int q = 0; q < 8; ++q) {
        const int z = (int)((RN_ZSEQ_V >> (4 * q)) & 15u);
        if ((RN_ZROOT_V >> q) & 1u) root = z;
        else if (R.lz_lambda[z] != R.lz_lambda[root]) return false;
    }
    return true;
}
static rn2_moves_fn pick_moves2(const Runner2Dev& R, bool dbg) {
    if (is_standard_set2(R))
        return dbg ? (rn2_moves_fn)runner2_moves_kernel<true, RN_PRESENT_V, RN_NSQ_V> : (rn2_moves_fn)runner2_moves_kernel<false, RN_PRESENT_V, RN_NSQ_V>;
    return dbg ? (rn2_moves_fn)runner2_moves_kernel<true, 0ull, 0u> : (rn2_moves_fn)runner2_moves_kernel<false, 0ull, 0u>;
}
cudaError_t runner2_prepare(int N, const Runner2Dev& R) {
    const int bytes = (int)runner2_smem_bytes(N, R);
    for (int dbg = 0; dbg < 2; ++dbg) {
        cudaError_t e = cudaFuncSetAttribute((const void*)pick_moves2(R, dbg != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute((const void*)runner2_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
void runner2_launch_init(const DevParams& P, const Runner2Dev& R, i64 n_rep, cudaStream_t st) {
    runner2_init_kernel<<<(unsigned)n_rep, RN_THREADS, runner2_smem_bytes(P.n_atoms, R), st>>>(P, R);
}
int runner2_launch_cycle(const DevParams& P, const Runner2Dev& R, i64 n_ladders, i64 idx, i64 cyc_in_launch, int sample, bool dbg, cudaStream_t st) {
    pick_moves2(R, dbg)<<<(unsigned)(n_ladders * P.n_traj), RN_THREADS, runner2_smem_bytes(P.n_atoms, R), st>>>(P, R, cyc_in_launch);
    runner_launch_cycle_end(P, R.en_c, n_ladders, idx, cyc_in_launch, sample, dbg, st);
    return 2;
}
void runner2_launch_export(const DevParams& P, const Runner2Dev& R, i64 rep, double* g_out, double* e_out, cudaStream_t st) {
    runner2_export_kernel<<<1, 256, 0, st>>>(P, R, rep, g_out, e_out);
}

}  // namespace ptmc
