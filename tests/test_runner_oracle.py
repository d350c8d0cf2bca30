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


def test_brass_six_atom_energy(oracle):
    # test/test_potentials.jl:67-87 with test/potentialfile.jl
    pos = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])
    g = P.host.load_scaling()
    rad_eta = [0.001, 0.020, 0.035, 0.100, 0.400]
    ang = np.array(P.host.CU_ANGULAR)
    w_cu = P.host.load_weights()
    w_zn = np.load(ROOT / "tests" / "golden" / "brass_nnp_inputs.npz")["weights_zn"]
    e, f12 = oracle.runner2_energy(pos, 4, 2, rad_eta, ang[:, 0], ang[:, 1], ang[:, 2], 11.338, g[:88], g[88:176], NODES, ACT, w_cu, w_zn)
    d12 = np.sqrt(((pos[0] - pos[1]) ** 2).sum())
    assert f12 == oracle.cutoff(d12, 11.338)                  # test_potentials.jl:82
    assert oracle.cutoff(0.0, 11.338) == 1.0                  # :83
    assert e == pytest.approx(-0.261652899, rel=1e-8)         # :86  (Julia's ≈: rtol sqrt(eps))
    assert e == pytest.approx(-0.26165289933547947, rel=1e-12)   # SURVEY A.3 (independent restatement + native forward)


def test_brass_atom_move_delta_energy(oracle):
    # test/test_states.jl:148-159 ("States NNVT"): Random.seed!(123); atom_displacement(config[1], max_displ[1] = 0.1, SphericalBC);
    # get_energy!(state, runnerpotential, "atommove") through calc_delta_matrix (2 species) -> new_en - en_tot ≈ -7.5971516e-5
    from ptmc_b200 import multi
    pos = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])
    g = P.host.load_scaling()
    rad_eta = [0.001, 0.020, 0.035, 0.100, 0.400]
    ang = np.array(P.host.CU_ANGULAR)
    w_cu = P.host.load_weights()
    w_zn = np.load(ROOT / "tests" / "golden" / "brass_nnp_inputs.npz")["weights_zn"]
    st = [int(x) for x in multi.julia_seed_state(123)]
    trial = pos[0] + (np.array([multi._jl_rand(st) for _ in range(3)]) - 0.5) * 0.1
    de = oracle.runner2_move_delta(pos, 4, 2, rad_eta, ang[:, 0], ang[:, 1], ang[:, 2], 11.338, g[:88], g[88:176], NODES, ACT, w_cu, w_zn, 0, trial)
    assert de == pytest.approx(-7.5971516e-5, rel=2e-8)        # Julia's ≈: rtol = sqrt(eps)
    assert de == pytest.approx(-7.597151644539446e-05, abs=1e-13)   # SURVEY §8c anchor (independent full-energy recompute)
    # the delta path against a fresh evaluation of the moved configuration
    moved = pos.copy(); moved[0] = trial
    e0, _ = oracle.runner2_energy(pos, 4, 2, rad_eta, ang[:, 0], ang[:, 1], ang[:, 2], 11.338, g[:88], g[88:176], NODES, ACT, w_cu, w_zn)
    e1, _ = oracle.runner2_energy(moved, 4, 2, rad_eta, ang[:, 0], ang[:, 1], ang[:, 2], 11.338, g[:88], g[88:176], NODES, ACT, w_cu, w_zn)
    assert de == pytest.approx(e1 - e0, rel=1e-9, abs=1e-15)


def test_brass_atom_swap_energy(oracle):
    # test/test_states.jl:166-183: Random.seed!(123); swap_atoms -> rand(1:N1), rand(N1+1:N) = (3, 6) (the test's own dist2_mat / config
    # assertions say so); get_energy!(state, runnerpotential, "atomswap") ≈ -0.0017084901948.  The swap exchanges the positions of one
    # atom of each species: checked through a fresh evaluation of the swapped configuration and through calc_swap_matrix
    from ptmc_b200 import multi
    pos = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])
    g = P.host.load_scaling()
    ang = np.array(P.host.CU_ANGULAR)
    w_zn = np.load(ROOT / "tests" / "golden" / "brass_nnp_inputs.npz")["weights_zn"]
    args = (4, 2, [0.001, 0.020, 0.035, 0.100, 0.400], ang[:, 0], ang[:, 1], ang[:, 2], 11.338, g[:88], g[88:176], NODES, ACT, P.host.load_weights(), w_zn)
    st = [int(x) for x in multi.julia_seed_state(123)]
    i1, i2 = multi._jl_rand_range(st, 4), 4 + multi._jl_rand_range(st, 2)
    assert (i1, i2) == (3, 6)
    swapped = pos.copy()
    swapped[[i1 - 1, i2 - 1]] = pos[[i2 - 1, i1 - 1]]
    de = oracle.runner2_energy(swapped, *args)[0] - oracle.runner2_energy(pos, *args)[0]
    assert de == pytest.approx(-0.0017084901948, rel=2e-8)
    # ... and through the restatement of calc_swap_matrix (DeltaMatrix.jl:319-632)
    assert oracle.runner2_swap_delta(pos, *args, i1 - 1, i2 - 1) == pytest.approx(-0.0017084901948, rel=2e-8)
    assert oracle.runner2_swap_delta(pos, *args, i1 - 1, i2 - 1) == pytest.approx(-0.0017084901948977715, abs=1e-13)   # SURVEY §8c anchor
    rng = np.random.default_rng(3)
    p2 = pos + rng.uniform(-0.3, 0.3, pos.shape)
    p2[:, 2] = rng.uniform(-1, 1, 6)
    for a1 in range(4):
        for a2 in (4, 5):
            sw = p2.copy()
            sw[[a1, a2]] = p2[[a2, a1]]
            fresh = oracle.runner2_energy(sw, *args)[0] - oracle.runner2_energy(p2, *args)[0]
            assert oracle.runner2_swap_delta(p2, *args, a1, a2) == pytest.approx(fresh, rel=1e-9, abs=1e-14)


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
