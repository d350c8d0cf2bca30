"""Pins the RuNNer part of the CPU oracle (oracle/runner_oracle.c):
  * forward == 21 for the all-ones network (test/test_runner_forward.jl:44-75);
  * forward against the reference's OWN binary (oracle/_ref/librunnerjulia.so) on the real Cu weights;
  * the 6-atom Cu4Zn2 energy of test/test_potentials.jl:67-87 (shares every per-pair / per-triple function with the
    single-species path);
  * consistency of the per-move delta update (DeltaMatrix.jl) with the from-scratch initialisation (SymmFunc.jl)."""
import pathlib

import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.host import initial_arrays, sweep_config

ROOT = pathlib.Path(__file__).resolve().parent.parent
NODES, ACT = [88, 20, 20, 1], [1, 2, 2, 1]


def test_forward_all_ones(oracle):
    npar = 88 * 20 + 20 + 20 * 20 + 20 + 20 + 1
    e = oracle.runner_forward(np.ones((2, 88)), NODES, ACT, np.ones(npar))
    assert np.array_equal(e, [21.0, 21.0])


def test_forward_against_reference_binary(oracle):
    from oracle import ref_forward
    if not ref_forward.available():
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    w = P.host.load_weights()
    rng = np.random.default_rng(1)
    x = rng.uniform(0.0, 1.0, (55, 88))
    ref = ref_forward.ref_forward(x, NODES, ACT, w)
    got = oracle.runner_forward(x, NODES, ACT, w)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)      # dgemm's summation order is unspecified
    # random small networks: layout of W (column-major n_out x n_in) and biases
    nodes, act = [5, 4, 3, 1], [1, 2, 2, 1]
    p = rng.normal(size=5 * 4 + 4 + 4 * 3 + 3 + 3 + 1)
    y = rng.normal(size=(7, 5))
    np.testing.assert_allclose(oracle.runner_forward(y, nodes, act, p), ref_forward.ref_forward(y, nodes, act, p), rtol=1e-13, atol=1e-15)


BRASS6 = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])


def test_brass_six_atom_energy(oracle):
    # test/test_potentials.jl:67-87 with test/potentialfile.jl, through the oracle ENGINE's total_symm_calc + forward passes
    r2 = P.host.brass_runner_potential(4, 2).params
    assert np.array_equal(r2.params_2, np.load(ROOT / "tests" / "golden" / "brass_nnp_inputs.npz")["weights_zn"])
    e, f12 = oracle.runner2_energy(r2, BRASS6)
    d12 = np.sqrt(((BRASS6[0] - BRASS6[1]) ** 2).sum())
    assert f12 == oracle.cutoff(d12, 11.338)                  # test_potentials.jl:82
    assert oracle.cutoff(0.0, 11.338) == 1.0                  # :83
    assert e == pytest.approx(-0.261652899, rel=1e-8)         # :86  (Julia's ≈: rtol sqrt(eps))
    assert e == pytest.approx(-0.26165289933547947, rel=1e-12)   # SURVEY A.3 (independent restatement + native forward)


def test_brass_atom_move_delta_energy(oracle):
    # test/test_states.jl:148-159 ("States NNVT"): Random.seed!(123); atom_displacement(config[1], max_displ[1] = 0.1, SphericalBC);
    # get_energy!(state, runnerpotential, "atommove") through calc_delta_matrix (2 species) -> new_en - en_tot ≈ -7.5971516e-5
    from ptmc_b200 import multi
    r2 = P.host.brass_runner_potential(4, 2).params
    st = [int(x) for x in multi.julia_seed_state(123)]
    trial = BRASS6[0] + (np.array([multi._jl_rand(st) for _ in range(3)]) - 0.5) * 0.1
    de = oracle.runner2_move_delta(r2, BRASS6, 0, trial)
    assert de == pytest.approx(-7.5971516e-5, rel=2e-8)        # Julia's ≈: rtol = sqrt(eps)
    assert de == pytest.approx(-7.597151644539446e-05, abs=1e-13)   # SURVEY §8c anchor (independent full-energy recompute)
    # the delta path against a fresh evaluation of the moved configuration, for a mover of either species
    rng = np.random.default_rng(11)
    for atom in range(6):
        tr = BRASS6[atom] + rng.uniform(-0.4, 0.4, 3)
        moved = BRASS6.copy(); moved[atom] = tr
        fresh = oracle.runner2_energy(r2, moved)[0] - oracle.runner2_energy(r2, BRASS6)[0]
        assert oracle.runner2_move_delta(r2, BRASS6, atom, tr) == pytest.approx(fresh, rel=1e-9, abs=1e-15)


def test_brass_atom_swap_energy(oracle):
    # test/test_states.jl:166-183: Random.seed!(123); swap_atoms -> rand(1:N1), rand(N1+1:N) = (3, 6) (the test's own dist2_mat / config
    # assertions say so); get_energy!(state, runnerpotential, "atomswap") ≈ -0.0017084901948.  The swap exchanges the positions of one
    # atom of each species: checked through a fresh evaluation of the swapped configuration and through calc_swap_matrix
    from ptmc_b200 import multi
    r2 = P.host.brass_runner_potential(4, 2).params
    st = [int(x) for x in multi.julia_seed_state(123)]
    i1, i2 = multi._jl_rand_range(st, 4), 4 + multi._jl_rand_range(st, 2)
    assert (i1, i2) == (3, 6)
    swapped = BRASS6.copy()
    swapped[[i1 - 1, i2 - 1]] = BRASS6[[i2 - 1, i1 - 1]]
    de = oracle.runner2_energy(r2, swapped)[0] - oracle.runner2_energy(r2, BRASS6)[0]
    assert de == pytest.approx(-0.0017084901948, rel=2e-8)
    # ... and through the restatement of calc_swap_matrix (DeltaMatrix.jl:319-632)
    assert oracle.runner2_swap_delta(r2, BRASS6, i1 - 1, i2 - 1) == pytest.approx(-0.0017084901948, rel=2e-8)
    assert oracle.runner2_swap_delta(r2, BRASS6, i1 - 1, i2 - 1) == pytest.approx(-0.0017084901948977715, abs=1e-13)   # SURVEY §8c anchor
    # the full table of the 8 possible Cu <-> Zn swap energies
    table = {(1, 5): 0.01026470570128829, (1, 6): 0.0030146695878782714, (2, 5): -0.001708490194897827, (2, 6): -0.012458164417961493,
             (3, 5): 0.01373553247113346, (3, 6): -0.0017084901948977715, (4, 5): 0.018061971748564276, (4, 6): 0.0023348531090901514}
    for (a1, a2), want in table.items():
        assert oracle.runner2_swap_delta(r2, BRASS6, a1 - 1, a2 - 1) == pytest.approx(want, abs=1e-13)
    rng = np.random.default_rng(3)
    p2 = BRASS6 + rng.uniform(-0.3, 0.3, BRASS6.shape)
    p2[:, 2] = rng.uniform(-1, 1, 6)
    for a1 in range(4):
        for a2 in (4, 5):
            sw = p2.copy()
            sw[[a1, a2]] = p2[[a2, a1]]
            fresh = oracle.runner2_energy(r2, sw)[0] - oracle.runner2_energy(r2, p2)[0]
            assert oracle.runner2_swap_delta(r2, p2, a1, a2) == pytest.approx(fresh, rel=1e-9, abs=1e-14)


def _cu55_engine(oracle, n_traj=3, seed=5, boundary=0.0):
    mcp, temp, start, pot, ens = P.host.deck_cu55_runner(n_traj=n_traj, cycles=2, boundary=boundary)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=seed, rng_mode=abi.RNG_XOSHIRO)
    eng = oracle.OracleEngine(cfg)
    pos, box = initial_arrays(mcp, start, 1)
    return eng, pos, box, cfg


def test_cu55_runner_initial_state(oracle):
    eng, pos, box, cfg = _cu55_engine(oracle)
    eng.upload_states(pos, box)
    g, e = eng.runner_state(0, 0)
    # rows of symmetry functions that are never evaluated stay exactly 0 (types 10, 110, 100), and quirk 3:
    # total_symm_calc loops (Nrad+1):Nang, so rows 79..88 are still 0 at initialisation
    assert (g[:, 1:10:2] == 0).all() and (g[:, 0:10:2] != 0).all()
    assert (g[:, 78:] == 0).all()
    ang_rows = np.arange(10, 78)
    active = ang_rows[(ang_rows - 10) % 3 == 0]
    assert (g[:, active] != 0).all() and (g[:, np.setdiff1d(ang_rows, active)] == 0).all()
    en = eng.download_states(("en_tot",))["en_tot"][0, 0]
    assert en == pytest.approx(e.sum(), rel=1e-15)
    # icosahedral symmetry: atoms of one shell have equal symmetry functions / energies
    assert np.ptp(e[1:13]) < 1e-8 * abs(e[1])       # the script's coordinates carry 8 digits
    eng.close()


def test_cu55_runner_delta_matches_recomputation(oracle):
    # after accepted moves the incrementally updated g_matrix (DeltaMatrix.jl) equals a from-scratch calculation
    # (SymmFunc.jl) on the new positions — except rows 80, 83, 86, which only the delta path ever fills (quirk 3)
    eng, pos, box, cfg = _cu55_engine(oracle)
    eng.upload_states(pos, box)
    eng.seed(5)
    eng.run_cycles(1, 2, False)
    st = eng.download_states()
    assert st["counters"][0, :, abi.C_ATOM_TOT].sum() > 50
    g_inc, e_inc = eng.runner_state(0, 1)
    eng2 = oracle.OracleEngine(cfg)
    eng2.upload_states(st["pos"], st["box"])
    g_new, e_new = eng2.runner_state(0, 1)
    late = [79, 82, 85]
    keep = np.setdiff1d(np.arange(88), late)
    np.testing.assert_allclose(g_inc[:, keep], g_new[:, keep], rtol=1e-10, atol=1e-12)
    assert (g_new[:, late] == 0).all() and (np.abs(g_inc[:, late]) > 0).any()
    eng.close(); eng2.close()


def test_cu55_runner_boundary_guard(oracle):
    # scripts/cu55.jl:153 passes boundary = 28 Angstrom: every pair is closer, new_energy = 100.0, nothing is accepted
    eng, pos, box, cfg = _cu55_engine(oracle, boundary=28 * 1.8897259886)
    eng.upload_states(pos, box)
    eng.seed(1)
    eng.run_cycles(1, 1, False)
    st = eng.download_states()
    assert st["counters"][0, :, abi.C_ATOM_TOT].sum() == 0
    assert np.array_equal(st["pos"], pos)
    eng.close()


# ---- two species through the oracle ENGINE (mc_move! with NNVT: atom moves and atom swaps) ---------------------------------
def _brass6_engine(oracle, n_swaps=1, rng_mode=abi.RNG_REPLAY, n_traj=2):
    pot = P.host.brass_runner_potential(4, 2)
    mcp = P.MCParams(10, n_traj, 6)
    temp = P.TempGrid(500.0, 650.0, n_traj)                          # test/test_states.jl:123
    start = P.Config(BRASS6, P.SphericalBC(radius=8.0))
    ens = P.host.NNVT((4, 2), n_atom_swaps=n_swaps)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=5, rng_mode=rng_mode)
    pos, box = initial_arrays(mcp, start, 1)
    return oracle.OracleEngine(cfg), cfg, pos, box


def _raw_for_index(k, n):
    """a raw UInt64 draw that Julia's rand(1:n) maps to k (1 + high64(x * n), low64 far from the rejection zone)"""
    return (((k - 1) << 64) // n + (1 << 58)) & 0xFFFFFFFFFFFFFFFF


def test_nnvt_engine_reproduces_reference_known_answers(oracle):
    """test/test_states.jl:110-186 ("States NNVT") through mc_move! of the oracle engine: the replay stream is the index draw that
    selects the move type followed by Julia's own Random.seed!(123) stream."""
    from ptmc_b200 import multi
    from parity import xoshiro_streams
    seed123 = xoshiro_streams(np.array([multi.julia_seed_state(123)], dtype=np.uint64), 8)[0]
    accept_all = np.uint64(0)                                        # rand() = 0.0 <= any probability
    # ---- atom move of atom 1: index draw -> 1, then rand() x 3 (x, y, z) from the seeded stream (:148-159)
    eng, cfg, pos, box = _brass6_engine(oracle)
    assert cfg.n_steps == 7
    try:
        eng.upload_states(pos, box)
        e0 = eng.download_states(("en_tot",))["en_tot"][0, 0]
        assert e0 == pytest.approx(-0.26165289933547947, rel=1e-12)  # test/test_potentials.jl:86 through set_variables / initialise_energy
        g0, _ = eng.runner_state(0, 0)
        assert g0.shape == (6, 88)
        slot = np.zeros((1, 2, 64), dtype=np.uint64)
        slot[0, :, 0] = _raw_for_index(1, 7)
        slot[0, :, 1:4] = seed123[:3]
        slot[0, :, 4] = accept_all
        # remaining moves of the cycle: swaps are not wanted here -> all index draws select atom 2 with a zero displacement
        for m in range(1, 7):
            slot[0, :, 5 * m] = _raw_for_index(2, 7)
            slot[0, :, 5 * m + 1:5 * m + 4] = np.uint64(1) << np.uint64(63)       # rand() = 0.5: trial = position
        lad = np.full((1, 8), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)               # exchange gate rand() ~ 1: never taken
        eng.set_replay(slot, lad)
        eng.trace_enable(1)
        eng.set_hist_bounds(-1.0, 0.0)
        eng.run_cycles(1, 1, False)
        tr = eng.trace_download()
        assert tr["move_index"][0, 0, 0, 0] == 1 and tr["accepted"][0, 0, 0, 0] == 1
        assert tr["delta_en"][0, 0, 0, 0] == pytest.approx(-7.5971516e-5, rel=2e-8)         # :159
        assert tr["delta_en"][0, 0, 0, 0] == pytest.approx(-7.597151644539446e-05, abs=1e-13)
        assert np.abs(tr["delta_en"][0, 0, 0, 1:]).max() < 1e-15                            # zero displacements
        st = eng.download_states()
        st123 = [int(x) for x in multi.julia_seed_state(123)]
        trial = BRASS6[0] + (np.array([multi._jl_rand(st123) for _ in range(3)]) - 0.5) * 0.1
        assert np.array_equal(st["pos"][0, 0, 0], trial)                                    # :162
        assert st["en_tot"][0, 0] - e0 == pytest.approx(-7.597151644539446e-05, abs=1e-13)   # :163
        assert st["counters"][0, 0, abi.C_ATOM_TOT] == 7
    finally:
        eng.close()
    # ---- atom swap: index draw -> 7 ("atomswap"), then rand(1:4), rand(5:6) from the seeded stream = (3, 6) (:166-186)
    eng, cfg, pos, box = _brass6_engine(oracle)
    try:
        eng.upload_states(pos, box)
        ref = eng.download_states()
        d2_ref = eng.dist2_mat(0, 0)
        slot = np.zeros((1, 2, 64), dtype=np.uint64)
        slot[0, :, 0] = _raw_for_index(7, 7)
        slot[0, :, 1:3] = seed123[:2]
        slot[0, :, 3] = accept_all
        for m in range(6):
            slot[0, :, 4 + 5 * m] = _raw_for_index(2, 7)
            slot[0, :, 4 + 5 * m + 1:4 + 5 * m + 4] = np.uint64(1) << np.uint64(63)
        eng.set_replay(slot, np.full((1, 8), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64))
        eng.trace_enable(1)
        eng.set_hist_bounds(-1.0, 0.0)
        eng.run_cycles(1, 1, False)
        tr = eng.trace_download()
        assert tr["move_index"][0, 0, 0, 0] == 7 + (3 << 10) + (6 << 20)                    # swap_indices = (3, 6)
        assert tr["accepted"][0, 0, 0, 0] == 1
        assert tr["delta_en"][0, 0, 0, 0] == pytest.approx(-0.0017084901948, rel=2e-8)      # :175
        st = eng.download_states()
        d2 = eng.dist2_mat(0, 0)
        refmat = d2_ref[5].copy(); refmat[5] = d2_ref[5, 2]; refmat[2] = 0.0                # :179-183
        assert np.array_equal(refmat, d2[2])
        assert st["en_tot"][0, 0] - ref["en_tot"][0, 0] == tr["delta_en"][0, 0, 0, 0]        # :184
        assert np.array_equal(st["pos"][0, 0, 2], ref["pos"][0, 0, 5]) and np.array_equal(st["pos"][0, 0, 5], ref["pos"][0, 0, 2])
        assert st["counters"][0, 0, abi.C_ATOM_TOT] == 6                                    # a swap touches no counter (swap_config.jl:241-276)
        # the swapped state is a state: its g matrix equals a fresh initialisation of the swapped configuration
        g_after, e_after = eng.runner_state(0, 0)
        eng2, _, _, _ = _brass6_engine(oracle)
        try:
            eng2.upload_states(st["pos"], st["box"])
            g_fresh, e_fresh = eng2.runner_state(0, 0)
            np.testing.assert_allclose(g_after, g_fresh, rtol=0, atol=1e-13)
            np.testing.assert_allclose(e_after, e_fresh, rtol=0, atol=1e-13)
        finally:
            eng2.close()
    finally:
        eng.close()


@pytest.mark.parametrize("rng_mode", [abi.RNG_XOSHIRO, abi.RNG_BATCH8])
def test_nnvt_engine_bookkeeping(oracle, rng_mode):
    """a production-stream NNVT run (atom moves + swaps + exchanges): the running energy / g matrix equal a fresh evaluation of the
    final configuration, swaps happen and keep the species counts, atoms stay in the sphere."""
    eng, cfg, pos, box = _brass6_engine(oracle, n_swaps=2, rng_mode=rng_mode, n_traj=3)
    try:
        eng.upload_states(pos, box)
        eng.seed(77)
        eng.trace_enable(30)
        eng.set_hist_bounds(-1.0, 0.0)
        eng.run_cycles(1, 30, True)
        tr = eng.trace_download()
        swaps = tr["move_index"] > 8
        assert swaps.sum() > 5 and (tr["accepted"][swaps] == 1).any()
        i1 = (tr["move_index"][swaps] >> 10) & 1023
        i2 = tr["move_index"][swaps] >> 20
        assert ((i1 >= 1) & (i1 <= 4)).all() and ((i2 >= 5) & (i2 <= 6)).all() and len(set(i1)) > 1 and len(set(i2)) > 1
        st = eng.download_states()
        assert ((st["pos"] ** 2).sum(-1) <= 64.0).all()
        eng2, _, _, _ = _brass6_engine(oracle, n_swaps=2, rng_mode=rng_mode, n_traj=3)
        try:
            eng2.upload_states(st["pos"], st["box"])
            np.testing.assert_allclose(st["en_tot"], eng2.download_states(("en_tot",))["en_tot"], rtol=0, atol=1e-12)
            for t in range(3):
                np.testing.assert_allclose(eng.runner_state(0, t)[0], eng2.runner_state(0, t)[0], rtol=0, atol=1e-12)
        finally:
            eng2.close()
    finally:
        eng.close()
