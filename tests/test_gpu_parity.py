"""Parity tests proper: the CUDA sweep (through the C ABI) against the CPU oracle on identical replayed
streams.  Needs a B200: run with `-m gpu`."""
import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.engine import Engine, PtmcError
from ptmc_b200.host import initial_arrays, sweep_config

from parity import compare_traces, run_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lanes", [0, 2, 4, 8, 16, 32])
def test_ne13_elj_nvt_replay(lanes):
    # BASELINE C1 at the script's own 25 temperatures (examples/ne13.jl:82): 2 ladders, 60 cycles incl. one
    # step-size adjustment (n_adjust=25) and two launches (chunks) so cursors/RNG/counters persist correctly.
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=25, cycles=60)
    n_t = {16: 24, 32: 16}.get(lanes, 25)   # 25 x 16 (x 32) lanes exceed the 384 (512) threads of one CTA
    if n_t != 25:
        mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=n_t, cycles=60)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=50, eq_cycles=10, seed=1234,
                        lanes=lanes, chunks=2, n_adjust=25)
    worst = compare_traces(gpu, cpu)
    assert gpu["trace"]["accepted"].mean() > 0.2
    assert gpu["trace"]["exchange"][:, :, 0].sum() > 0          # some exchanges were attempted
    if lanes:
        assert gpu["lanes"] == lanes
    print("ne13 lanes", lanes, "max |dE err|", worst)


@pytest.mark.parametrize("lanes", [0, 8])
def test_ne55_elj_nvt_replay(lanes):
    # BASELINE C2 geometry and ladder (scripts/ne55.jl), 3 ladders; lanes 0 = auto: 4 lanes per replica for free-space ELJEven
    # ladders of at most 128 threads (8 replicas per warp), 8 otherwise
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=24, cycles=20)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=3, n_cycles=20, seed=99, n_adjust=10, lanes=lanes)
    compare_traces(gpu, cpu)
    assert gpu["lanes"] == (lanes or 4)


def test_elj_odd_potential_replay():
    # ELJPotential (odd powers, src/EnergyEvaluation.jl:923-932) on the Ne13 geometry
    mcp, temp, start, _, ens = P.host.deck_ne13(n_traj=8, cycles=20)
    pot = P.ELJPotential((-10.5097942564988, 0.0, 989.725135614556, 0.0, -101383.865938807))
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=1, n_cycles=20, seed=5)
    compare_traces(gpu, cpu)


@pytest.mark.parametrize("fac,dmax,seed,cycles", [(1.05, 24.0, 11, 3), (1.3, 1.5, 5, 10)])
def test_sphere_redraw_and_recentre(fac, dmax, seed, cycles):
    # tight binding sphere + big steps (src/MCMoves.jl:115-145, quirk 8).  Case 1: ~94 % of the trials leave
    # the sphere, so 5 of the 6 replicas reach count == 50 and recentre! (checked on the CPU oracle: their
    # centre of mass drops from 4e-11 to 1e-16) without reaching the error at 100; case 2: frequent redraws
    # with ~9 % of the moves accepted.
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=6, cycles=cycles)
    pos = start.positions
    rmax = np.sqrt((pos ** 2).sum(axis=1).max())
    tight = P.Config(pos, P.SphericalBC(radius=rmax * fac))
    md = np.full((1, 6, 4), 0.1)
    md[:, :, 0] = dmax
    gpu, cpu = run_pair(mcp, temp, tight, pot, ens, n_ladders=1, n_cycles=cycles, seed=seed, max_displ=md, slack=400)
    compare_traces(gpu, cpu)
    com = np.abs(gpu["states"]["pos"].mean(axis=2)).max(axis=-1)
    if dmax > 20:
        assert (com < 1e-14).sum() >= 4          # recentred replicas


def test_cu55_eam_replay():
    # BASELINE C3 (scripts/cu55.jl): full new energy from the updated (phi, rho) components
    mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=28, cycles=10)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=10, seed=3, n_adjust=5)
    compare_traces(gpu, cpu)


def _small_pbc(bc_cls):
    # 32-atom fcc block (2x2x2 cells x 4) of the neon lattice, cubic / rectangular / rhombic cell
    a = 8.4
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pos = np.array([(b + [i, j, k]) * a for i in range(2) for j in range(2) for k in range(2) for b in base])
    L = 2 * a
    if bc_cls is P.CubicBC:
        bc = P.CubicBC(L)
    elif bc_cls is P.RectangularBC:
        bc = P.RectangularBC(L, L * 1.25)
        pos = pos * [1, 1, 1.25]
    else:
        bc = P.RhombicBC(L, L)
        pos = pos @ np.array([[1, 0, 0], [0.5, 3 ** 0.5 / 2, 0], [0, 0, 1]])
    return P.Config(pos - pos.mean(axis=0) * 0, bc)


@pytest.mark.parametrize("bc_cls", [P.CubicBC, P.RectangularBC, P.RhombicBC])
def test_small_npt_replay(bc_cls):
    start = _small_pbc(bc_cls)
    mcp = P.MCParams(30, 6, 32, n_adjust=10)
    temp = P.TempGrid(20.0, 40.0, 6)
    pot = P.ELJPotentialEven(P.host.ELJ_NE_COEFF)
    ens = P.NPT(32, 101325 * 3.398928944382626e-14, False)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=30, seed=21)
    compare_traces(gpu, cpu, positions_exact=False)
    assert gpu["trace"]["move_index"].max() == 33          # volume moves happened


@pytest.mark.parametrize("lanes", [0, 32])
def test_pbc216_npt_replay(lanes):
    # BASELINE C4 (scripts/pbc216.jl): 216 atoms, rhombic cell, 16 temperatures, fcc/hcp alternating
    # (lanes 0 -> 16 lanes per replica, two replicas per warp; 32: one warp per replica)
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=16, cycles=3)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=1, n_cycles=3, seed=8, lanes=lanes)
    compare_traces(gpu, cpu, positions_exact=False)


@pytest.mark.parametrize("lanes", [0, 32])
def test_pbc216_npt_replay_through_thaw(lanes):
    # long enough for every atom of every configuration to have moved: the kernel stops correcting old rows from the tie table
    # ("thawed"); uniform volume moves take their trial energy from the power sums the atom moves carry;
    # decisions and energies must stay on the oracle's trajectory
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=4, cycles=40)
    md = np.full((1, 4, 4), 0.1)
    md[:, :, 0] = 0.5
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=1, n_cycles=40, seed=12, max_displ=md, n_adjust=20, lanes=lanes)
    compare_traces(gpu, cpu, positions_exact=False)
    acc = gpu["trace"]["accepted"]
    idx = gpu["trace"]["move_index"]
    moved_atoms = [len(np.unique(idx[0, :, s][(acc[0, :, s] == 1) & (idx[0, :, s] <= 216)])) for s in range(4)]
    assert min(moved_atoms) > 200, moved_atoms          # (exchanges permute configurations; every slot saw nearly all atoms move)


def test_pbc216_separated_moves_tangent_potential_replay():
    # (f-3) at C4's size: pbc216 in Angstrom with ELJPotentialB and NPT(separated_volume = true) - the reference's tangent matrix is
    # state (stale after rejected xy / z trials); the kernels carry it as one factor per configuration + one bit per atom,
    # the oracle literally as the 216 x 216 matrix; replayed streams.
    # (lattice perturbed by 1e-3: half-cell ties under a separated scaling are a documented waiver)
    a2b = 1.8897261259077824
    mcp, temp, start, _, ens = P.host.deck_pbc216(n_traj=4, cycles=12)
    bc = start[0].boundary_condition
    rng = np.random.default_rng(5)
    start = [P.Config(c.positions / a2b + rng.uniform(-1e-3, 1e-3, c.positions.shape), P.RhombicBC(bc.box_length / a2b, bc.box_height / a2b))
             for c in start]
    pot = P.ELJPotentialB(P.host.ELJB_NE_A, P.host.ELJB_NE_B, P.host.ELJB_NE_C)
    ens = P.NPT(216, ens.pressure * a2b ** 3, True, n_volume_moves=6)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=1, n_cycles=12, seed=14)
    compare_traces(gpu, cpu, positions_exact=False)
    cnt = gpu["states"]["counters"]
    assert cnt[:, :, abi.C_VOLXY_TOT].sum() > 0 and cnt[:, :, abi.C_VOLZ_TOT].sum() > 0 and cnt[:, :, abi.C_VOL_TOT].sum() > 0


def test_periodic_nvt_replay():
    start = _small_pbc(P.CubicBC)
    mcp = P.MCParams(20, 4, 32)
    gpu, cpu = run_pair(mcp, P.TempGrid(20.0, 40.0, 4), start, P.ELJPotentialEven(P.host.ELJ_NE_COEFF), P.NVT(32),
                        n_ladders=1, n_cycles=20, seed=2)
    compare_traces(gpu, cpu)


def _free_run_pair(deck, n_ladders, seed, rng_mode, cycles, lanes=0, bounds=(-0.01, 0.0)):
    """Device-generated random streams (no replay file): both engines seeded by the same rule."""
    from oracle import oracle
    mcp, temp, start, pot, ens = deck
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=n_ladders, seed=seed, rng_mode=rng_mode, lanes_per_replica=lanes)
    pos, box = initial_arrays(mcp, start, n_ladders)
    res = []
    for cls in (Engine, oracle.OracleEngine):
        eng = cls(cfg)
        eng.upload_states(pos, box)
        eng.seed(seed)
        eng.run_cycles(1, cycles // 2, False)
        eng.reset_counters()
        eng.set_hist_bounds(*bounds)
        eng.run_cycles(1, cycles, True)
        eng.sync()
        res.append((eng.download_states(), eng.download_hist(), eng.get_rng_states()))
        eng.close()
    return res


def _assert_same_trajectory(res, e_rtol=1e-11):
    (sg, hg, rg), (sc, hc, rc) = res
    assert np.array_equal(sg["counters"], sc["counters"]), "counters differ"
    assert np.array_equal(sg["pos"], sc["pos"]), "positions differ"
    assert np.array_equal(sg["max_displ"], sc["max_displ"])
    assert np.array_equal(hg, hc), "histograms differ"
    assert np.array_equal(rg[0], rc[0]) and np.array_equal(rg[1], rc[1]), "generator states differ"
    np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=e_rtol)
    np.testing.assert_allclose(sg["ham"], sc["ham"], rtol=e_rtol)


def test_xoshiro_mode_matches_oracle_seeding():
    # compatibility RNG mode: ONE Julia-compatible Xoshiro256++ stream per slot, seeded like
    # Random.seed!(seed + k) -> same trajectory as the oracle stepping the same generator
    res = _free_run_pair(P.host.deck_ne13(n_traj=8, cycles=30), 4, 4321, abi.RNG_XOSHIRO, 30)
    _assert_same_trajectory(res)


@pytest.mark.parametrize("lanes", [2, 4, 8, 16, 32])
def test_batch8_production_mode_ne13(lanes):
    # production RNG mode (8 substreams per slot, one batch per move) through the specialised kernels:
    # the whole trajectory (positions, counters, step sizes, histograms, generator states) is bit-identical
    # to the oracle following the same draw schedule
    res = _free_run_pair(P.host.deck_ne13(n_traj=12, cycles=60), 3, 77, abi.RNG_BATCH8, 60, lanes=lanes)
    _assert_same_trajectory(res)
    assert res[0][0]["counters"][:, :, abi.C_EXC_ATT].sum() > 0


@pytest.mark.parametrize("lanes", [0, 4, 16])
def test_batch8_production_mode_ne55(lanes):
    # BASELINE C2 geometry, production kernels (lanes 0 -> 8 lanes, 7 unrolled pair iterations)
    res = _free_run_pair(P.host.deck_ne55(n_traj=24, cycles=30), 2, 5, abi.RNG_BATCH8, 30, lanes=lanes, bounds=(-0.1, 0.0))
    _assert_same_trajectory(res)


def test_batch8_production_mode_eam_and_redraws():
    res = _free_run_pair(P.host.deck_cu55_eam(n_traj=28, cycles=10), 1, 9, abi.RNG_BATCH8, 10, bounds=(-10.0, 0.0))
    _assert_same_trajectory(res, e_rtol=1e-10)
    # tight sphere: first redraw from positions 5-7, later ones from fresh batches
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=6, cycles=20)
    rmax = np.sqrt((start.positions ** 2).sum(axis=1).max())
    tight = P.Config(start.positions, P.SphericalBC(radius=rmax * 1.05))
    res = _free_run_pair((mcp, temp, tight, pot, ens), 2, 3, abi.RNG_BATCH8, 20)
    _assert_same_trajectory(res)


def _production_trace_pair(deck, n_ladders, seed, cycles, bounds, *, eq_cycles=0, max_displ=None, lanes=0, positions_exact=True, rel=1e-10):
    """Per-move trace of the PRODUCTION kernels (rng_mode BATCH8: the instantiations bench.py times) against the oracle stepping
    the same eight substreams per slot: move index, accept bit and exchange record bit-identical, per-move delta energy to
    north_star's 1e-10 (src/MCRun.jl:122-162).  Runs on the test-only build lib/libptmc_b200_trace.so (same sources,
    -DPTMC_PROD_TRACE: those instantiations write the trace under a runtime flag; the shipped library leaves the three
    instructions per move out), and checks that the shipped library follows the same trajectory."""
    from oracle import oracle
    trace_lib = abi.load_library(abi.LIB_PATH.parent / "libptmc_b200_trace.so")
    mcp, temp, start, pot, ens = deck
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=n_ladders, seed=seed, rng_mode=abi.RNG_BATCH8, lanes_per_replica=lanes)
    pos, box = initial_arrays(mcp, start, n_ladders)
    out = []
    for make, traced in ((lambda: Engine(cfg, lib=trace_lib), True), (lambda: oracle.OracleEngine(cfg), True), (lambda: Engine(cfg), False)):
        eng = make()
        eng.upload_states(pos, box, max_displ)
        init = eng.download_states(("en_tot",))["en_tot"].copy()
        eng.seed(seed)
        if eq_cycles:
            eng.run_cycles(1, eq_cycles, False)
            eng.reset_counters()
        if traced:
            eng.trace_enable(cycles)
        eng.set_hist_bounds(*bounds)
        eng.run_cycles(1, cycles, True)
        eng.sync()
        out.append({"init_en": init, "states": eng.download_states(), "hist": eng.download_hist(),
                    "hist_reduced": eng.download_hist(reduce_ladders=True), "trace": eng.trace_download() if traced else None,
                    "rng": eng.get_rng_states()})
        eng.close()
    gpu, cpu, shipped = out
    worst = compare_traces(gpu, cpu, positions_exact=positions_exact, rel=rel)
    assert np.array_equal(gpu["rng"][0], cpu["rng"][0]) and np.array_equal(gpu["rng"][1], cpu["rng"][1])
    # the shipped library (no trace code in these kernels) walks the same trajectory bit for bit
    for k in ("pos", "box", "en_tot", "max_displ", "counters", "ham"):
        assert np.array_equal(shipped["states"][k], gpu["states"][k]), k
    assert np.array_equal(shipped["hist"], gpu["hist"]) and np.array_equal(shipped["rng"][0], gpu["rng"][0])
    return gpu, cpu, worst


def test_shipped_library_refuses_production_trace():
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=8, cycles=5)
    eng = Engine(sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=1, rng_mode=abi.RNG_BATCH8))
    pos, box = initial_arrays(mcp, start, 1)
    eng.upload_states(pos, box)
    eng.trace_enable(2)
    with pytest.raises(PtmcError, match="libptmc_b200_trace"):
        eng.run_cycles(1, 2, False)
    eng.close()


def test_production_kernel_trace_c2_ne55():
    # BASELINE C2 at its own ladder shape: 24 temperatures x 4 lanes = sweep_kernel_g4s<ELJ_EVEN, false, 14, 6, FULLW, false>, the
    # kernel of the headline number; step sizes adapted during 100 equilibration cycles as bench.py's pilot does
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=24, cycles=40)
    mcp.n_adjust = 10
    gpu, _, worst = _production_trace_pair((mcp, temp, start, pot, ens), 3, 77, 40, (-0.035, -0.02, None, None), eq_cycles=100)
    # ... and the 8-lane kernel (sweep_kernel<8, ELJ_EVEN, SPHERE, false, 7, 6, FULLW>: ladders of more than 32 temperatures, round 1's headline)
    _production_trace_pair((mcp, temp, start, pot, ens), 2, 78, 20, (-0.035, -0.02, None, None), eq_cycles=50, lanes=8)
    acc = gpu["trace"]["accepted"].mean()
    assert 0.3 < acc < 0.75 and gpu["trace"]["exchange"][..., 0].sum() > 3
    print("C2 production trace: %d moves, acceptance %.3f, max |dE err| %.2e" % (gpu["trace"]["accepted"].size, acc, worst))


def test_production_kernel_trace_c1_c3_c4():
    # C1 Ne13 at 32 temperatures (ROT kernel, 4 lanes), C3 Cu55 EAM at 28 temperatures (unrolled EAM kernel, trial components in
    # registers), C4 pbc216 at 16 temperatures through the frozen -> thawed switch with volume moves
    deck = P.host.deck_ne13(n_traj=32, cycles=60)
    deck[0].n_adjust = 10
    gpu, _, w1 = _production_trace_pair(deck, 4, 5, 60, (-0.006, -0.003, None, None), eq_cycles=100)
    assert 0.3 < gpu["trace"]["accepted"].mean() < 0.8
    deck = P.host.deck_cu55_eam(n_traj=28, cycles=15)
    deck[0].n_adjust = 5
    gpu, _, w3 = _production_trace_pair(deck, 2, 9, 15, (-6.3, -5.6, None, None), eq_cycles=30)
    assert 0.2 < gpu["trace"]["accepted"].mean() < 0.95
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=16, cycles=30)
    md = np.full((1, 16, 4), 0.1)
    md[:, :, 0] = 0.5
    v0 = P.get_volume(start[0].boundary_condition)
    gpu, _, w4 = _production_trace_pair((mcp, temp, start, pot, ens), 1, 3, 30, (-0.5, 1.5, 0.8 * v0, 2.0 * v0), max_displ=md, positions_exact=False)
    # (that was sweep_kernel_p16, the 16-lane kernel the benchmark times; the one-warp-per-replica kernel of the other periodic sizes:)
    _production_trace_pair((mcp, temp, start, pot, ens), 1, 4, 12, (-0.5, 1.5, 0.8 * v0, 2.0 * v0), max_displ=md, positions_exact=False, lanes=32)
    idx, acc = gpu["trace"]["move_index"], gpu["trace"]["accepted"]
    assert (idx == 217).sum() > 100 and acc[idx == 217].sum() > 5                     # volume moves, some accepted
    moved = [len(np.unique(idx[0, :, s][(acc[0, :, s] == 1) & (idx[0, :, s] <= 216)])) for s in range(16)]
    assert min(moved) > 150, moved
    print("production traces: max |dE err| C1 %.2e, C3 %.2e, C4 %.2e" % (w1, w3, w4))


def test_energy_bookkeeping_at_full_size():
    # size-independent property at BASELINE C2's full batch (4096 ladders x 24 temperatures): after a sweep
    # the running en_tot of every replica equals the energy recomputed from its positions (the init kernel).
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=24, cycles=5)
    L = 4096
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=1)
    pos, box = initial_arrays(mcp, start, L)
    eng = Engine(cfg)
    eng.upload_states(pos, box)
    eng.run_cycles(1, 5, False)
    eng.sync()
    st = eng.download_states()
    acc = st["counters"][:, :, abi.C_ATOM_TOT].sum() / (L * 24 * 55 * 5)
    assert 0.2 < acc < 1.0
    eng2 = Engine(cfg)
    eng2.upload_states(st["pos"], st["box"])
    e2 = eng2.download_states(("en_tot",))["en_tot"]
    np.testing.assert_allclose(st["en_tot"], e2, rtol=1e-11)
    r2 = (st["pos"] ** 2).sum(axis=-1)
    assert (r2 <= box[..., :1] + 1e-9).all()                  # nobody left the binding sphere
    assert not np.array_equal(st["pos"][0], st["pos"][1])      # ladders are independent streams
    eng.close(); eng2.close()


def test_cu55_runner_replay(oracle):
    # BASELINE C5 (scripts/cu55.jl, runnerpotential with boundary 0): symmetry-function delta update + network forward
    # pass + Metropolis fused in one kernel, against the oracle's restatement of DeltaMatrix.jl / ForwardPass.jl
    mcp, temp, start, pot, ens = P.host.deck_cu55_runner(n_traj=4, cycles=3)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=3, seed=17, n_adjust=2,
                        max_displ=np.full((2, 4, 4), 0.35))
    worst = compare_traces(gpu, cpu)
    assert 0.1 < gpu["trace"]["accepted"].mean() < 0.9
    print("runner max |dE err|", worst, "E", cpu["states"]["en_tot"][0, 0])


def test_cu55_runner_state_matrices(oracle):
    # initial g_matrix (incl. the (Nrad+1):Nang initialisation quirk), per-atom energies, and the same after moves
    from oracle import oracle as orc
    mcp, temp, start, pot, ens = P.host.deck_cu55_runner(n_traj=3, cycles=2)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=4, rng_mode=abi.RNG_XOSHIRO)
    pos, box = initial_arrays(mcp, start, 1)
    md = np.full((1, 3, 4), 0.3)
    engs = [Engine(cfg), orc.OracleEngine(cfg)]
    for e in engs:
        e.upload_states(pos, box, md)
    g0, e0 = engs[0].runner_state(0, 1)
    g1, e1 = engs[1].runner_state(0, 1)
    assert np.array_equal(g0 == 0, g1 == 0)                   # never-evaluated rows and rows 79..88 are exactly 0
    np.testing.assert_allclose(g0, g1, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(e0, e1, rtol=1e-12)
    for e in engs:
        e.seed(4)
        e.run_cycles(1, 2, False)
        e.sync()
    sg, sc = engs[0].download_states(), engs[1].download_states()
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(sg["pos"], sc["pos"])
    assert sg["counters"][0, :, abi.C_ATOM_TOT].sum() > 30
    for slot in range(3):
        g0, e0 = engs[0].runner_state(0, slot)
        g1, e1 = engs[1].runner_state(0, slot)
        np.testing.assert_allclose(g0, g1, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(e0, e1, rtol=1e-11)
        assert (np.abs(g0[:, [79, 82, 85]]) > 0).any()        # filled by the delta path only
    for e in engs:
        e.close()


def test_runner_generic_symmetry_function_set(oracle):
    # a symmetry-function set that is NOT the reference's standard table: fewer angular functions, zeta = 3 (pow()
    # instead of repeated squaring), an extra radial function, hidden layers 7 and 5, linear second activation ->
    # the generic (runtime-structure) instantiation of the kernel
    rng = np.random.default_rng(3)
    radial = [(11, 0.01, 9.0, 14.0, 0.2), (10, 0.01, 9.0, 1.0, 0.0), (11, 0.2, 9.0, 5.0, 0.1), (11, 0.05, 9.0, 8.0, 0.0)]
    angular = [(0.002, 1, 1, 9.0, 111, 3.0, 0.0), (0.002, -1, 3, 9.0, 111, 2.0, 0.1), (0.002, -1, 3, 9.0, 110, 1.0, 0.0),
               (0.03, -1, 3, 9.0, 111, 1.5, 0.0), (0.03, 1, 8, 9.0, 111, 0.9, 0.05), (0.01, 1, 1, 9.0, 111, 2.5, 0.3)]
    nodes = [10, 7, 5, 1]
    w = rng.normal(scale=0.3, size=10 * 7 + 7 + 7 * 5 + 5 + 5 + 1)
    pot = P.RuNNerPotential(nodes, [1, 2, 1, 1], w, radial, angular, boundary=1.0)
    mcp, temp, start, _, ens = P.host.deck_ne13(n_traj=4, cycles=6)
    temp = P.TempGrid(300.0, 900.0, 4)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=6, seed=23, n_adjust=3, max_displ=np.full((2, 4, 4), 0.8),
                        bounds=(-50.0, 50.0, None, None))
    compare_traces(gpu, cpu)
    assert 0.05 < gpu["trace"]["accepted"].mean() < 0.98


def test_cu55_runner_production_mode_and_guard(oracle):
    res = _free_run_pair(P.host.deck_cu55_runner(n_traj=3, cycles=2), 2, 31, abi.RNG_BATCH8, 2, bounds=(-13.0, -11.0))
    _assert_same_trajectory(res, e_rtol=1e-11)
    # scripts/cu55.jl:153 passes boundary = 28 Angstrom: every trial has a neighbour inside it -> new_energy = 100.0
    mcp, temp, start, pot, ens = P.host.deck_cu55_runner(n_traj=3, cycles=1, boundary=28 * 1.8897259886)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=2)
    pos, box = initial_arrays(mcp, start, 1)
    eng = Engine(cfg)
    eng.upload_states(pos, box)
    e0 = eng.download_states(("en_tot",))["en_tot"].copy()
    eng.run_cycles(1, 1, False)
    st = eng.download_states()
    assert st["counters"][0, :, abi.C_ATOM_TOT].sum() == 0 and np.array_equal(st["pos"], pos) and np.array_equal(st["en_tot"], e0)
    eng.close()


@pytest.mark.parametrize("case", ["ne13", "eam", "cubic_nvt", "rhombic_npt", "runner"])
def test_rdf_matches_oracle(case, oracle):
    # update_rdf! (src/Sampling.jl:301-316): the device bins distances recomputed from the positions, the oracle bins its
    # stored dist2_mat like the reference; same trajectory (production streams) -> identical histograms.  NPT: the reference's
    # matrix is rescaled by scale^2 at volume moves instead of being recomputed (1-ulp differences): a count may move to the
    # neighbouring bin when a distance sits within 1e-16 of a bin edge.
    from oracle import oracle as orc
    if case == "ne13":
        deck, bounds = P.host.deck_ne13(n_traj=6, cycles=30), (-0.01, 0.0)
    elif case == "eam":
        deck, bounds = P.host.deck_cu55_eam(n_traj=4, cycles=10), (-10.0, 0.0)
    elif case == "cubic_nvt":
        deck, bounds = (P.MCParams(20, 4, 32), P.TempGrid(20.0, 40.0, 4), _small_pbc(P.CubicBC), P.ELJPotentialEven(P.host.ELJ_NE_COEFF), P.NVT(32)), (-0.1, 0.0)
    elif case == "rhombic_npt":
        deck, bounds = (P.MCParams(20, 4, 32), P.TempGrid(20.0, 40.0, 4), _small_pbc(P.RhombicBC), P.ELJPotentialEven(P.host.ELJ_NE_COEFF), P.NPT(32, 101325 * 3.398928944382626e-14)), None
    else:
        deck, bounds = P.host.deck_cu55_runner(n_traj=3, cycles=2), (-13.0, -11.0)
    mcp, temp, start, pot, ens = deck
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=2, seed=11, rng_mode=abi.RNG_BATCH8)
    pos, box = initial_arrays(mcp, start, 2)
    bc = (start[0] if isinstance(start, (list, tuple)) else start).boundary_condition
    if bounds is None:
        v0 = P.get_volume(bc)
        bounds = (-1.0, 1.0, 0.8 * v0, 2.0 * v0)
    delta = P.host.rdf_delta_r2(bc, mcp.n_bin)
    out = []
    for cls in (Engine, orc.OracleEngine):
        eng = cls(cfg)
        eng.upload_states(pos, box)
        eng.seed(11)
        eng.set_hist_bounds(*bounds)
        eng.rdf_enable(delta)
        for k in range(3):
            eng.run_cycles(1 + k * mcp.mc_cycles, mcp.mc_cycles, True)
            eng.rdf_sample()
        out.append((eng.download_rdf(), eng.download_states()))
        eng.close()
    (rg, sg), (rc, sc) = out
    if case == "rhombic_npt":
        np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-11, atol=1e-11)   # exp / pow of two libms in the volume scale
    else:
        assert np.array_equal(sg["pos"], sc["pos"])
    n_pairs = mcp.n_atoms * (mcp.n_atoms - 1) // 2
    assert rg.shape == (2, mcp.n_traj, 5 * mcp.n_bin) and 0 < rg.sum() <= 3 * 2 * mcp.n_traj * n_pairs
    if case == "rhombic_npt":
        assert np.abs(rg - rc).sum() <= 8 and abs(rg.sum() - rc.sum()) <= 4
    else:
        assert np.array_equal(rg, rc)


def test_ptmc_run_checkpoint_files_gpu(tmp_path):
    # ptmc_run with save / rdfsave on the device, restart from the written directory (src/MCRun.jl:363-488, src/ReadSave.jl)
    from ptmc_b200 import readsave
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=8, cycles=2000)
    mcp.eq_percentage = 0.05
    st, res = P.host.ptmc_run(mcp, temp, start, pot, ens, n_ladders=3, seed=5, rdfsave=True, save=1000, save_dir=tmp_path)
    for l in range(3):
        cfgs, md, ca, cv, back, rdf, index = readsave.rebuild_states(8, tmp_path / f"ladder{l}")
        assert index == 2000 and np.array_equal(back.en_histogram, res[l].en_histogram) and np.array_equal(rdf, res[l].rdf)
        assert np.array_equal(np.stack([c.positions for c in cfgs]), st["pos"][l]) and np.array_equal(ca, st["counters"][l, :, abi.C_ATOM_TOT])
        assert rdf.sum() > 0 and back.en_histogram.sum() == 8 * 2000
    mcp2, temp2, ens2, pot2 = readsave.read_init(tmp_path / "ladder0")
    mcp2.mc_cycles = 2500
    st2, res2 = P.host.ptmc_run(mcp2, temp2, start, pot2, ens2, n_ladders=3, seed=5, save=1000, save_dir=tmp_path, restart=True)
    assert all(r.en_histogram.sum() == 8 * 2501 for r in res2)      # cycle 2000 runs again: start_counter = index (src/Initialization.jl:108)
    assert readsave.rebuild_states(8, tmp_path / "ladder2")[-1] == 2500


def test_pair_arithmetic_accuracy(oracle):
    # the production pair kernel (one shared reciprocal, Halley step, Horner + FMA) against the reference's
    # own formulation (1 + K divisions, src/EnergyEvaluation.jl:934-942) over the whole physical range
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=4, cycles=1)
    eng = Engine(sweep_config(mcp, temp, start, pot, ens))
    rng = np.random.default_rng(0)
    r = np.concatenate([rng.uniform(3.0, 60.0, 200000), rng.uniform(4.5, 8.0, 200000)])
    d2n, d2o = r ** 2, rng.permutation(r) ** 2
    en, eo = eng.debug_pair_energy(d2n, d2o)
    ref_n = np.array([oracle.dimer_energy_eljeven(pot.coeff, float(x)) for x in d2n[:20000]])
    ref_o = np.array([oracle.dimer_energy_eljeven(pot.coeff, float(x)) for x in d2o[:20000]])
    # the polynomial has a root near r = 5.2 bohr: compare against the scale of its largest term
    scale = lambda d2: np.abs(np.array(pot.coeff)[None, :] * d2[:, None] ** -(np.arange(6)[None, :] + 3.0)).max(axis=1)
    assert (np.abs(en[:20000] - ref_n) <= 16 * np.finfo(float).eps * scale(d2n[:20000])).all()
    assert (np.abs(eo[:20000] - ref_o) <= 16 * np.finfo(float).eps * scale(d2o[:20000])).all()
    # switched-off lanes contribute exactly nothing
    z = eng.debug_pair_energy(np.full(4, 1e150), np.full(4, 1e150))
    assert (z[0] == 0).all() and (z[1] == 0).all()
    eng.close()


def test_eam_power_accuracy():
    # invrexp (src/EnergyEvaluation.jl:1088-1095): the table-driven (r2)^-(n/2) of the sweep kernel against extended
    # precision over the physical range and at the edges of the tabulated exponent range (log/exp path outside it)
    mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=4, cycles=1)
    eng = Engine(sweep_config(mcp, temp, start, pot, ens))
    rng = np.random.default_rng(1)
    d2 = np.concatenate([rng.uniform(1.0, 4000.0, 400000), 2.0 ** rng.uniform(-70, 70, 20000),
                         np.array([1.0, 2.0, 4.0, 1.0 + 2.0 ** -52, 2.0 - 2.0 ** -52, 2.0 ** -64, 2.0 ** 63.5])])
    phi, _ = eng.debug_pair_energy(d2, d2[::-1].copy())
    ref = d2.astype(np.longdouble) ** np.longdouble(-pot.n / 2)
    rel = np.abs((phi.astype(np.longdouble) - ref) / ref).astype(float)
    inside = (d2 >= 2.0 ** -64) & (d2 < 2.0 ** 64)
    assert rel[inside].max() <= 4 * np.finfo(float).eps, rel[inside].max()
    assert rel[~inside].max() <= 1e-12, rel[~inside].max()      # exp(-p log r2): the rounding of p * log r2 ~ 200
    eng.close()


def test_eam_sqrt_accuracy():
    # calc_energies_from_components (src/EnergyEvaluation.jl:1156-1162) takes sqrt(rho_i) for every atom and move: the kernel's branch-free
    # square root (rsqrt seed, one coupled Newton step, one residual step) against the correctly rounded one: <= 1 ulp
    mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=4, cycles=1)
    eng = Engine(sweep_config(mcp, temp, start, pot, ens))
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(1e-3, 1e3, 400000), 2.0 ** rng.uniform(-200, 200, 100000), np.array([1.0, 2.0, 4.0, 1.0 + 2.0 ** -52, 4.0 - 2.0 ** -50])])
    _, got = eng.debug_pair_energy(np.ones_like(x), x)
    ref = np.sqrt(x)
    ulp = np.abs(got - ref) / np.spacing(ref)
    assert ulp.max() <= 1.0, ulp.max()
    assert (ulp == 0).mean() > 0.9
    eng.close()


@pytest.mark.parametrize("case", ["eam", "pbc216", "runner", "eam_full", "pbc216_full", "runner_full"])
def test_energy_bookkeeping_other_potentials(case):
    # size-independent property for the other potentials: the running en_tot (updated move by move, through exchanges,
    # volume moves, the frozen -> thawed switch of the periodic kernel, incremental G-matrix updates of RuNNer) equals the
    # energy a fresh handle computes from the final positions and boxes
    # "_full": the ladder counts of the benchmark configurations (C3 2048 x 28, C4 592 x 16, C5 37 x 28 replicas)
    full = case.endswith("_full")
    case = case.replace("_full", "")
    if case == "eam":
        deck, L, cycles, tol = P.host.deck_cu55_eam(n_traj=28, cycles=40), (2048 if full else 64), (10 if full else 40), 1e-11
    elif case == "pbc216":
        deck, L, cycles, tol = P.host.deck_pbc216(n_traj=16, cycles=40), (592 if full else 8), (12 if full else 40), 1e-10
    else:
        deck, L, cycles, tol = P.host.deck_cu55_runner(n_traj=28 if full else 8, cycles=6), (37 if full else 2), (3 if full else 6), 1e-11
    mcp, temp, start, pot, ens = deck
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=3)
    pos, box = initial_arrays(mcp, start, L)
    md = np.full((L, mcp.n_traj, 4), 0.1)
    md[:, :, 0] = 0.4
    eng = Engine(cfg)
    eng.upload_states(pos, box, md)
    eng.run_cycles(1, cycles, False)
    st = eng.download_states()
    assert st["counters"][:, :, abi.C_ATOM_TOT].sum() > 0.1 * L * mcp.n_traj * mcp.n_atoms * cycles
    if case == "pbc216":
        assert st["counters"][:, :, abi.C_VOL_TOT].sum() > 0 and not np.array_equal(st["box"], box)
    eng2 = Engine(cfg)
    eng2.upload_states(st["pos"], st["box"])
    e2 = eng2.download_states(("en_tot",))["en_tot"]
    if case == "runner":
        # rows 80/83/86 of the G matrix differ between the incremental path and an initialisation from scratch (quirk 3),
        # so compare per-configuration energies through the G matrices' common rows instead: re-derive with the oracle-free
        # property that the sum of per-atom energies equals en_tot
        for l in range(0, L, 9 if full else 1):
            for t in range(mcp.n_traj):
                g, ea = eng.runner_state(l, t)
                assert ea.sum() == pytest.approx(st["en_tot"][l, t], rel=1e-13)
    else:
        np.testing.assert_allclose(st["en_tot"], e2, rtol=tol)
    eng.close(); eng2.close()


def test_ptmc_run_host_api(oracle):
    # the reference-facing call: ptmc_run! (src/MCRun.jl:363-432) with the reference's argument list; the same call through
    # the oracle engine gives the same Output (histogram window from the half-length equilibration, quirk 1; finalise_results)
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=8, cycles=300)
    out = []
    for cls in (Engine, oracle.OracleEngine):
        states, results = P.ptmc_run(mcp, temp, start, pot, ens, n_ladders=3, seed=5, engine_cls=cls)
        out.append((states, results))
    (sg, rg), (sc, rc) = out
    assert len(rg) == 3
    for a, b in zip(rg, rc):
        # the window comes from each engine's own running energies, which agree to the last bits only
        assert a.en_min == pytest.approx(b.en_min, rel=1e-12) and a.en_max == pytest.approx(b.en_max, rel=1e-12)
        assert a.delta_en_hist == (a.en_max - a.en_min) / (a.n_bin - 1)
        assert a.en_histogram.sum() == 8 * 300 and np.abs(a.en_histogram - b.en_histogram).sum() <= 4   # bin-edge ties
        np.testing.assert_allclose(a.en_avg, b.en_avg, rtol=1e-12)
        np.testing.assert_allclose(a.heat_cap, b.heat_cap, rtol=1e-8)
        assert np.array_equal(a.count_stat_atom, b.count_stat_atom)
        assert np.array_equal(np.nan_to_num(a.count_stat_exc), np.nan_to_num(b.count_stat_exc))
    assert np.array_equal(sg["pos"], sc["pos"])
    with pytest.raises(ValueError):                           # SphericalBC + NPT, src/Ensembles.jl:189
        P.ptmc_run(mcp, temp, start, pot, P.NPT(13, 1e-9, False))


def test_error_behaviour():
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=4, cycles=5)
    cfg = sweep_config(mcp, temp, start, pot, ens)
    eng = Engine(cfg)
    with pytest.raises(PtmcError) as e:
        eng.run_cycles(1, 1, False)                           # before upload
    assert e.value.code == abi.ERR_STATE
    pos, box = initial_arrays(mcp, start, 1)
    eng.upload_states(pos, box)
    with pytest.raises(PtmcError) as e:
        eng.run_cycles(1, 1, True)                            # sampling before set_hist_bounds
    assert e.value.code == abi.ERR_STATE
    eng.close()
    with pytest.raises(PtmcError):                            # SphericalBC + NPT (src/Ensembles.jl:189)
        Engine(sweep_config(mcp, temp, start, pot, P.NVT(13), lanes_per_replica=3))
    # replay stream too short -> PTMC_ERR_REPLAY at sync
    with pytest.raises(PtmcError) as e:                       # tracing needs a sequential-stream kernel
        eng = Engine(cfg)
        eng.upload_states(pos, box)
        eng.trace_enable(4)
        eng.run_cycles(1, 1, False)
    assert e.value.code == abi.ERR_STATE
    eng.close()
    cfg = sweep_config(mcp, temp, start, pot, ens, rng_mode=abi.RNG_REPLAY)
    eng = Engine(cfg)
    eng.upload_states(pos, box)
    eng.set_replay(np.zeros((1, 4, 8), dtype=np.uint64) + np.uint64(1 << 62), np.zeros((1, 8), dtype=np.uint64))
    eng.run_cycles(1, 2, False)
    with pytest.raises(PtmcError) as e:
        eng.sync()
    assert e.value.code == abi.ERR_REPLAY
    eng.close()


def test_forward_all_ones_kat():
    # test/test_runner_forward.jl:44-75: all-ones parameters and inputs, nodes [88,20,20,1], acts [1,2,2,1],
    # batch 2 -> 21 per atom
    import ctypes as C
    lib = abi.load_library()
    nodes = np.array([88, 20, 20, 1], dtype=np.int32)
    act = np.array([1, 2, 2, 1], dtype=np.int32)
    npar = 88 * 20 + 20 + 20 * 20 + 20 + 20 + 1
    params = np.ones(npar)
    x = np.ones((2, 88))
    e = np.zeros(2)
    i32 = lambda v: C.byref(C.c_int32(v))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lib.forward(dp(x), i32(88), i32(4), ip(nodes), i32(2), i32(npar), dp(params), ip(act), dp(e))
    assert np.allclose(e, 21.0, rtol=1e-14)


def _device_forward(x, nodes, act, params):
    """the exported `forward` of libptmc_b200.so with the Fortran symbol's calling convention (every argument by reference)"""
    import ctypes as C
    lib = abi.load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    nodes = np.ascontiguousarray(nodes, dtype=np.int32); act = np.ascontiguousarray(act, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    e = np.zeros(x.shape[0])
    i32 = lambda v: C.byref(C.c_int32(v))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    ret = lib.forward(dp(x), i32(int(nodes[0])), i32(len(nodes)), ip(nodes), i32(x.shape[0]), i32(len(params)), dp(params), ip(act), dp(e))
    return e, ret


def test_forward_export_against_reference_binary(oracle):
    """The drop-in `forward` on REAL weights and on random networks: column-major W, bias offsets, the activation index shift and
    the [batch][n_in] input layout are all visible here (the all-ones KAT sees none of them).  Reference = the reference's own
    binary (oracle/_ref/librunnerjulia.so) where it was built, and always the oracle restatement (itself pinned to that binary)."""
    from oracle import ref_forward
    rng = np.random.default_rng(12)
    cases = []
    nodes, act = [88, 20, 20, 1], [1, 2, 2, 1]
    for w in (P.host.load_weights(), P.host.load_weights_zn()):               # weights.029.data / weights.030.data
        cases.append((rng.uniform(0.0, 1.0, (55, 88)), nodes, act, w))        # batch 55 = one Cu55 configuration
    cases.append((rng.normal(size=(7, 5)), [5, 4, 3, 1], [1, 2, 2, 1], rng.normal(size=5 * 4 + 4 + 4 * 3 + 3 + 3 + 1)))
    cases.append((rng.normal(size=(130, 5)), [5, 4, 3, 1], [1, 1, 2, 2], rng.normal(size=5 * 4 + 4 + 4 * 3 + 3 + 3 + 1)))   # mixed activations, tanh output
    cases.append((rng.normal(size=(3, 6)), [6, 9, 2, 5, 1], [1, 2, 1, 2, 1], rng.normal(scale=0.5, size=6 * 9 + 9 + 9 * 2 + 2 + 2 * 5 + 5 + 5 + 1)))
    cases.append((rng.normal(size=(1, 4)), [4, 1], [1, 2], rng.normal(size=5)))                                            # a single layer, batch 1
    for x, nd, ac, w in cases:
        got, ret = _device_forward(x, nd, ac, w)
        want = oracle.runner_forward(x, nd, ac, w)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)
        assert ret == got[-1]                                                  # the Fortran function's return register
        if ref_forward.available():
            np.testing.assert_allclose(got, ref_forward.ref_forward(x, nd, ac, w), rtol=1e-13, atol=1e-15)
    # error behaviour: inconsistent sizes give NaN instead of reading out of bounds
    assert np.isnan(_device_forward(np.ones((2, 4)), [4, 3, 1], [1, 2, 1], np.ones(5))[1])  # too few parameters


def test_eljb_magnetic_field_potential(oracle):
    # ELJPotentialB (src/EnergyEvaluation.jl:943-958), SphericalBC / NVT as scripts/ne13_B.jl: the device takes cos 2theta from
    # the positions (2 dz^2 / r^2 - 1) where the reference keeps a tangent matrix; replayed streams -> same decisions, delta
    # energies to 1e-10; production streams -> identical trajectory
    mcp, temp, start, pot, ens = P.host.deck_ne13_B(n_traj=8, cycles=40)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=40, seed=17)
    compare_traces(gpu, cpu)
    assert 0.05 < gpu["trace"]["accepted"].mean() < 0.999
    res = _free_run_pair(P.host.deck_ne13_B(n_traj=16, cycles=60), 3, 23, abi.RNG_BATCH8, 60, bounds=(-0.02, 0.0))
    _assert_same_trajectory(res, e_rtol=1e-10)
    # energy bookkeeping: running en_tot == energy recomputed from the final positions by a fresh handle
    (sg, _, _), _ = res
    cfg = sweep_config(*P.host.deck_ne13_B(n_traj=16, cycles=60), n_ladders=3, seed=1)
    eng = Engine(cfg)
    eng.upload_states(sg["pos"], sg["box"])
    np.testing.assert_allclose(eng.download_states(("en_tot",))["en_tot"], sg["en_tot"], rtol=1e-11)
    eng.close()


@pytest.mark.parametrize("case", ["ne13", "eam", "eljb"])
def test_split_cycle_phases_equal_whole_cycles(case):
    # ptmc_run_phase + ptmc_export_config / ptmc_import_config (a ladder that spans GPUs): mc_step! and the tail of
    # mc_cycle! as separate launches with the exchange done by the host driver give the trajectory of whole cycles on one handle
    from ptmc_b200 import multi
    deck = {"ne13": P.host.deck_ne13(n_traj=6, cycles=40), "eam": P.host.deck_cu55_eam(n_traj=4, cycles=12),
            "eljb": P.host.deck_ne13_B(n_traj=6, cycles=40)}[case]
    bounds = {"ne13": (-0.0075, -0.0045), "eam": (-10.0, 0.0), "eljb": (-0.02, 0.0)}[case]
    mcp, temp, start, pot, ens = deck
    mcp.n_adjust = 10
    L, seed = 3, 99
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=seed)
    eng = Engine(cfg)
    pos, box = initial_arrays(mcp, start, L)
    eng.upload_states(pos, box)
    multi.seed_global(eng, seed, 0, L)
    eng.run_cycles(1, 6, False)
    eng.reset_counters()
    eng.set_hist_bounds(*bounds)
    eng.run_cycles(1, mcp.mc_cycles, True)
    ref, ref_hist = eng.download_states(), eng.download_hist()
    eng.close()
    st, hist, (t0, cnt) = multi.run_split_ladder(mcp, temp, start, pot, ens, n_ladders=L, seed=seed, bounds=bounds, eq_cycles=6)
    assert (t0, cnt) == (0, mcp.n_traj)
    assert ref["counters"][:, :, abi.C_EXC_ATT].sum() > 0
    for k in ("pos", "counters", "max_displ", "en_tot", "ham"):
        assert np.array_equal(st[k], ref[k]), k
    assert np.array_equal(hist, ref_hist)


@pytest.mark.parametrize("bc_cls", [P.RectangularBC, P.RhombicBC])
def test_separated_volume_moves_replay(bc_cls):
    # NPT(separated_volume = true): volume_change_separated (src/MCMoves.jl:300-343) draws rand(1:6) -> z (1/6), xy (2/6) or
    # uniform (3/6) scaling with its own step size, counter and adaptation rule (src/Exchange.jl:221-245); the trial distance
    # matrix is recomputed (get_distance2_mat!).  Replayed streams: same decisions, dE to 1e-10, same counters / step sizes.
    # (started from a slightly perturbed lattice: on a perfect rhombic lattice the image chosen for pairs that sit exactly half
    #  a cell apart in y is decided by the last bit of the scale factor, i.e. by the libm's exp / pow - DESIGN.md waivers)
    lattice = _small_pbc(bc_cls)
    start = P.Config(lattice.positions + np.random.default_rng(8).uniform(-1e-3, 1e-3, lattice.positions.shape), lattice.boundary_condition)
    mcp = P.MCParams(40, 6, 32, n_adjust=10)
    temp = P.TempGrid(20.0, 40.0, 6)
    pot = P.ELJPotentialEven(P.host.ELJ_NE_COEFF)
    ens = P.NPT(32, 101325 * 3.398928944382626e-14, True, n_volume_moves=4)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=40, seed=33)
    compare_traces(gpu, cpu, positions_exact=False)
    cnt = gpu["states"]["counters"]
    assert cnt[:, :, abi.C_VOLXY_TOT].sum() > 0 and cnt[:, :, abi.C_VOLZ_TOT].sum() > 0 and cnt[:, :, abi.C_VOL_TOT].sum() > 0
    box = gpu["states"]["box"]
    assert not np.allclose(box[:, :, 0] / box[:, :, 1], box[0, 0, 0] / box[0, 0, 1] * 0 + start.boundary_condition.box_length / start.boundary_condition.box_height)
    # production streams: whole trajectory against the oracle
    res = _free_run_pair((mcp, temp, start, pot, ens), 2, 5, abi.RNG_BATCH8, 40, bounds=(-1.0, 1.0, 0.5 * P.get_volume(start.boundary_condition), 3.0 * P.get_volume(start.boundary_condition)))
    (sg, hg, rg), (sc, hc, rc) = res
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(hg, hc) and np.array_equal(rg[0], rc[0])
    np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(sg["max_displ"], sc["max_displ"], rtol=1e-15)
    with pytest.raises((PtmcError, ValueError)):      # scale_xy / scale_z do not exist for CubicBC
        Engine(sweep_config(mcp, temp, _small_pbc(P.CubicBC), pot, ens))


def test_ptmc_run_save_restart_other_potentials(tmp_path):
    # the host driver's checkpoint / restart / rdf paths on the device for NPT (volume_hist.data), RuNNer and EAM
    from ptmc_b200 import readsave
    d = tmp_path / "npt"
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=4, cycles=40)
    st, res = P.ptmc_run(mcp, temp, start, pot, ens, n_ladders=2, seed=1, save=20, save_dir=d)
    cf, md, ca, cv, back, rdf, idx = readsave.rebuild_states(4, d / "ladder1", npt=True)
    assert idx == 40 and np.array_equal(back.ev_histogram, res[1].ev_histogram) and back.ev_histogram.sum() == 4 * 40
    mcp2, temp2, ens2, pot2 = readsave.read_init(d / "ladder0")
    mcp2.mc_cycles = 60
    _, res2 = P.ptmc_run(mcp2, temp2, start, pot2, ens2, n_ladders=2, seed=1, save=20, save_dir=d, restart=True)
    assert res2[0].ev_histogram.sum() == 4 * 61
    d = tmp_path / "runner"
    mcp, temp, start, pot, ens = P.host.deck_cu55_runner(n_traj=3, cycles=1000)
    mcp.eq_percentage = 0.0
    _, res = P.ptmc_run(mcp, temp, start, pot, ens, n_ladders=1, seed=1, save=500, save_dir=d, rdfsave=True)
    assert res[0].rdf.sum() == 3 * 55 * 54 // 2 and (d / "rdf.data").exists()       # one sample at cycle 1000, every pair inside the range
    assert (d / "params.data").read_text().splitlines()[4].strip() == "runnerpotential"
    d = tmp_path / "eam"
    mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=4, cycles=300)
    P.ptmc_run(mcp, temp, start, pot, ens, n_ladders=2, seed=1, save=100, save_dir=d)
    mcp2, temp2, ens2, pot2 = readsave.read_init(d / "ladder0")
    assert (pot2.n, pot2.m, pot2.ean, pot2.eCam) == (pot.n, pot.m, pot.ean, pot.eCam)
    mcp2.mc_cycles = 400
    _, res2 = P.ptmc_run(mcp2, temp2, start, pot2, ens2, n_ladders=2, seed=1, save=100, save_dir=d, restart=True)
    assert res2[0].en_histogram.sum() == 4 * 401


def _synthetic_lookup_table():
    # a table in the reference's layout filled from a smooth anisotropic pair energy (Angstrom units like scripts/ne13-lookup.jl)
    l_dist, start, dd = 1100, 2.0, 0.005
    r = start + dd * np.arange(1, l_dist + 1)
    ang = np.deg2rad(np.arange(91.0))
    e0 = 1.2e-4 * ((3.1 / r) ** 12 - 2 * (3.1 / r) ** 6)
    table = e0[None, :] * (1 + 0.2 * np.cos(2 * ang)[:, None])
    c6 = table[:, 999].mean() * (start + 1000 * dd) ** 6
    return P.LookupTablePotential(table, start_dist=start, start_angle=0.0, d_dist=dd, d_angle=1.0, c6coeff=c6)


def test_lookup_table_potential(tmp_path, oracle):
    # LookupTablePotential (src/EnergyEvaluation.jl:214-295), SphericalBC / NVT as scripts/ne13-lookup.jl (part of the reference's
    # runtests): table indices from the angle to z and the distance; replayed decisions identical, production trajectory identical
    pot = _synthetic_lookup_table()
    # through the file format of read_lookuptable as well
    f = tmp_path / "lut.txt"
    with open(f, "w") as fh:
        fh.write("%d, %r, %r ! distances\n%d, 0.0, 1.0\n" % (pot.l_dist, pot.start_dist, pot.d_dist, pot.l_angle))
        fh.writelines("%r\n" % float(v) for v in pot.table.reshape(-1))
    pot_f = P.LookupTablePotential(f)
    assert np.array_equal(pot_f.table, pot.table) and pot_f.c6coeff == pytest.approx(pot.c6coeff, rel=1e-15) and pot_f.start_dist == pot.start_dist
    mcp, temp, start, _, ens = P.host.deck_ne13_lookup(pot_f, n_traj=8, cycles=40)
    gpu, cpu = run_pair(mcp, temp, start, pot_f, ens, n_ladders=2, n_cycles=40, seed=29)
    compare_traces(gpu, cpu)
    assert 0.05 < gpu["trace"]["accepted"].mean() < 0.999
    res = _free_run_pair((mcp, temp, start, pot_f, ens), 3, 31, abi.RNG_BATCH8, 60, bounds=(-0.02, 0.0))
    _assert_same_trajectory(res, e_rtol=1e-10)
    (sg, _, _), _ = res
    eng = Engine(sweep_config(mcp, temp, start, pot_f, ens, n_ladders=3, seed=1))
    eng.upload_states(sg["pos"], sg["box"])
    np.testing.assert_allclose(eng.download_states(("en_tot",))["en_tot"], sg["en_tot"], rtol=1e-11)
    eng.close()


def test_reference_script_decks_cu38_and_pbc32():
    # the remaining decks of the reference's "scripts run without errors" tests (test/runtests.jl:296-316): scripts/cu38.jl (Cu38 EAM,
    # 36 temperatures: five unrolled pair iterations) and scripts/pbc32-short.jl (32-atom fcc block, CubicBC, NPT, 24 temperatures)
    res = _free_run_pair(P.host.deck_cu38_eam(cycles=12), 2, 41, abi.RNG_BATCH8, 12, bounds=(-5.0, -3.0))
    _assert_same_trajectory(res, e_rtol=1e-10)
    assert res[0][0]["counters"][:, :, abi.C_ATOM_TOT].sum() > 0
    mcp, temp, start, pot, ens = P.host.deck_pbc32(cycles=20)
    v0 = P.get_volume(start.boundary_condition)
    (sg, hg, rg), (sc, hc, rc) = _free_run_pair((mcp, temp, start, pot, ens), 2, 43, abi.RNG_BATCH8, 20, bounds=(-1.0, 1.0, 0.8 * v0, 2.0 * v0))
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(hg, hc) and np.array_equal(rg[0], rc[0])
    np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=1e-10)
    assert sg["counters"][:, :, abi.C_VOL_TOT].sum() > 0


@pytest.mark.parametrize("bc_cls,lanes", [(P.RhombicBC, 0), (P.CubicBC, 16), (P.RhombicBC, 32), (P.RectangularBC, 0)])
def test_uniform_volume_moves_from_power_sums(bc_cls, lanes):
    # periodic NPT, ELJPotentialEven{6}, production streams: the kernels take the trial energy of volume_change_uniform
    # (src/MCMoves.jl:277-287, dimer_energy_config on dist2_mat * scale^2: src/MCRun.jl:76-85) from six power sums that the atom moves
    # carry along - five volume moves per cycle from the perfect lattice (cut-off ties live, then thawed), 256+ cycles so
    # that the periodic rebuild of the sums is crossed: same decisions as the oracle's O(N^2) sums, energies to 1e-10
    start = _small_pbc(bc_cls)
    mcp = P.MCParams(300, 6, 32, n_adjust=20)
    temp = P.TempGrid(20.0, 40.0, 6)
    pot = P.ELJPotentialEven(P.host.ELJ_NE_COEFF)
    ens = P.NPT(32, 101325 * 3.398928944382626e-14, False, n_volume_moves=5)
    v0 = P.get_volume(start.boundary_condition)
    (sg, hg, rg), (sc, hc, rc) = _free_run_pair((mcp, temp, start, pot, ens), 2, 17, abi.RNG_BATCH8, 300, lanes=lanes,
                                                bounds=(-1.0, 1.0, 0.5 * v0, 3.0 * v0))
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(hg, hc) and np.array_equal(rg[0], rc[0])
    assert sg["counters"][:, :, abi.C_VOL_TOT].sum() > 100
    np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-10, atol=1e-10)     # exp / pow of two libms in every accepted scale factor
    np.testing.assert_allclose(sg["box"], sc["box"], rtol=1e-11)
    np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=1e-10)
    np.testing.assert_allclose(sg["ham"], sc["ham"], rtol=1e-10)


@pytest.mark.parametrize("n_traj", [8, 16, 32])
def test_rotated_neighbour_order_kernels(n_traj):
    # (N - 1) % G == 0 and no partly active warp (Ne13 on 4 lanes, n_traj * 4 a multiple of 32): the ROT kernels walk the other
    # 12 atoms in three iterations in rotated order; production streams -> the trajectory of the oracle, bit for bit
    res = _free_run_pair(P.host.deck_ne13(n_traj=n_traj, cycles=60), 3, 61 + n_traj, abi.RNG_BATCH8, 60, lanes=4)
    _assert_same_trajectory(res)
    assert res[0][0]["counters"][:, :, abi.C_ATOM_TOT].sum() > 0
    # tight sphere: redraws and the recentre path inside the rotated kernels
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=8, cycles=20)
    rmax = np.sqrt((start.positions ** 2).sum(axis=1).max())
    tight = P.Config(start.positions, P.SphericalBC(radius=rmax * 1.05))
    _assert_same_trajectory(_free_run_pair((mcp, temp, tight, pot, ens), 2, 3, abi.RNG_BATCH8, 20, lanes=4))
    # two lanes per replica (16 replicas per warp, six rotated iterations, four substreams per lane): n_traj * 2 a multiple of 32
    if n_traj % 16 == 0:
        _assert_same_trajectory(_free_run_pair(P.host.deck_ne13(n_traj=n_traj, cycles=60), 3, 91 + n_traj, abi.RNG_BATCH8, 60, lanes=2))
        mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=16, cycles=20)
        tight = P.Config(start.positions, P.SphericalBC(radius=rmax * 1.05))
        _assert_same_trajectory(_free_run_pair((mcp, temp, tight, pot, ens), 2, 5, abi.RNG_BATCH8, 20, lanes=2))


def _fcc_cluster(n, a):
    # the n lattice points of an fcc crystal closest to the origin
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    r = range(-3, 4)
    pts = np.array([(b + [i, j, k]) * a for i in r for j in r for k in r for b in base])
    pts = pts[np.argsort((pts ** 2).sum(axis=1), kind="stable")][:n]
    return pts - pts.mean(axis=0)


@pytest.mark.parametrize("case", ["eam101", "elj150", "eam20"])
def test_other_cluster_sizes_generic_kernels(case):
    # sizes outside the unrolled instantiations: Cu101 EAM as scripts/cu101.jl (13 pair iterations on 8 lanes -> the generic EAM
    # kernel with the trial components in shared memory), a 150-atom ELJ cluster on 32 lanes, a 20-atom EAM cluster on 4 lanes
    if case == "eam101":
        mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=32, cycles=4)
        mcp.n_atoms, ens = 101, P.NVT(101)
        start = P.Config(_fcc_cluster(101, 0.36258 * 18.8973), P.SphericalBC(radius=14 * 1.8897259886))
        bounds, cyc = (-20.0, 0.0), 4
    elif case == "elj150":
        mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=8, cycles=6)
        mcp.n_atoms, ens = 150, P.NVT(150)
        start = P.Config(_fcc_cluster(150, 8.4), P.SphericalBC(radius=40.0))
        bounds, cyc = (-1.0, 0.0), 6
    else:
        mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=6, cycles=20)
        mcp.n_atoms, ens = 20, P.NVT(20)
        start = P.Config(_fcc_cluster(20, 0.36258 * 18.8973), P.SphericalBC(radius=10 * 1.8897259886))
        bounds, cyc = (-5.0, 0.0), 20
    res = _free_run_pair((mcp, temp, start, pot, ens), 2, 71, abi.RNG_BATCH8, cyc, bounds=bounds)
    _assert_same_trajectory(res, e_rtol=1e-10)
    assert res[0][0]["counters"][:, :, abi.C_ATOM_TOT].sum() > 0


def test_edge_case_sizes_replay():
    # smallest legal job (2 atoms, 2 temperatures, one atom move per cycle), sampling every third cycle, step-size adaptation
    # every cycle; and a ladder of 48 temperatures on 8 lanes (384 threads, the largest ladder of the small-lane kernels)
    start = P.Config(np.array([[0.0, 0.0, 2.9], [0.0, 0.0, -2.9]]), P.SphericalBC(radius=8.0))
    mcp = P.MCParams(60, 2, 2, mc_sample=3, n_adjust=1)
    gpu, cpu = run_pair(mcp, P.TempGrid(5.0, 9.0, 2), start, P.ELJPotentialEven(P.host.ELJ_NE_COEFF), P.NVT(2, n_atom_moves=1),
                        n_ladders=3, n_cycles=60, seed=5)
    compare_traces(gpu, cpu)
    assert gpu["hist"].sum() == 3 * 2 * 20                      # cycles 3, 6, ..., 60
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=48, cycles=6)
    res = _free_run_pair((mcp, temp, start, pot, ens), 2, 9, abi.RNG_BATCH8, 6, bounds=(-0.1, 0.0))
    _assert_same_trajectory(res)
    with pytest.raises(PtmcError):                              # 49 temperatures x 8 lanes do not fit one ladder CTA
        Engine(sweep_config(*P.host.deck_ne55(n_traj=49, cycles=1), lanes_per_replica=8))


def _scaled(cfg, f):
    bc = cfg.boundary_condition
    nbc = P.CubicBC(bc.box_length * f) if isinstance(bc, P.CubicBC) else type(bc)(bc.box_length * f, bc.box_height * f)
    return P.Config(cfg.positions * f, nbc)


@pytest.mark.parametrize("pot_name", ["eljb", "lookup"])
@pytest.mark.parametrize("bc_cls", [P.CubicBC, P.RectangularBC, P.RhombicBC])
def test_periodic_direction_dependent_potentials(pot_name, bc_cls, oracle):
    # ELJPotentialB / LookupTablePotential in a periodic cell, NVT and NPT with uniform volume moves (scripts/pbc150.jl, pbc96.jl run
    # the lookup table this way): tangent from the minimum-image separation, d2 <= r_cut and the potentials' own long-range
    # corrections (src/EnergyEvaluation.jl:494-543,673-789,983-1009).  Replayed streams: decisions identical, dE to 1e-10.
    start = _scaled(_small_pbc(bc_cls), 1 / 1.8897259886)            # both potentials work in Angstrom
    pot = P.ELJPotentialB(P.host.ELJB_NE_A, P.host.ELJB_NE_B, P.host.ELJB_NE_C) if pot_name == "eljb" else _synthetic_lookup_table()
    temp = P.TempGrid(20.0, 40.0, 4)
    # NVT
    gpu, cpu = run_pair(P.MCParams(20, 4, 32, n_adjust=10), temp, start, pot, P.NVT(32), n_ladders=2, n_cycles=20, seed=3)
    compare_traces(gpu, cpu)
    # NPT, uniform volume moves
    ens = P.NPT(32, 101325 * 3.398928944382626e-14 * 1.8897259886 ** 3, False, n_volume_moves=2)
    gpu, cpu = run_pair(P.MCParams(30, 4, 32, n_adjust=10), temp, start, pot, ens, n_ladders=2, n_cycles=30, seed=4)
    compare_traces(gpu, cpu, positions_exact=False)
    assert gpu["trace"]["move_index"].max() == 34 and gpu["states"]["counters"][:, :, abi.C_VOL_TOT].sum() > 0
    if bc_cls is P.CubicBC:
        return
    # NPT, separated volume moves: tan_mat becomes state - xy / z trials write the TRIAL configuration's tangents into it and they
    # stay there when the move is rejected (the reference's new_tan_mat and tan_mat are one array, set_variables
    # src/EnergyEvaluation.jl:1790,1798; get_tantheta_mat!(new_tan_mat, trial) src/MCMoves.jl:333-341); atom moves and the uniform
    # trial then read those entries.  Perturbed lattice as in test_separated_volume_moves_replay.
    pert = P.Config(start.positions + np.random.default_rng(8).uniform(-1e-3, 1e-3, start.positions.shape), start.boundary_condition)
    ens = P.NPT(32, 101325 * 3.398928944382626e-14 * 1.8897259886 ** 3, True, n_volume_moves=4)
    mcp = P.MCParams(40, 4, 32, n_adjust=10)
    gpu, cpu = run_pair(mcp, temp, pert, pot, ens, n_ladders=2, n_cycles=40, seed=6)
    compare_traces(gpu, cpu, positions_exact=False)
    cnt = gpu["states"]["counters"]
    assert cnt[:, :, abi.C_VOLXY_TOT].sum() > 0 and cnt[:, :, abi.C_VOLZ_TOT].sum() > 0 and cnt[:, :, abi.C_VOL_TOT].sum() > 0
    v0 = P.get_volume(pert.boundary_condition)
    res = _free_run_pair((mcp, temp, pert, pot, ens), 2, 7, abi.RNG_BATCH8, 40, bounds=(-1.0, 1.0, 0.5 * v0, 3.0 * v0))
    (sg, hg, rg), (sc, hc, rc) = res
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(hg, hc) and np.array_equal(rg[0], rc[0])
    np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-11, atol=1e-11)


# ---- two species: RuNNerPotential2Atom + NNVT atom swaps (scripts/brass55.jl) ------------------------------------------------
BRASS6 = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])


def _brass6_deck(n_traj=2, n_swaps=1, cycles=10, t=(500.0, 650.0)):
    """test/test_states.jl:110-125 ("States NNVT"): Cu4Zn2, SphericalBC(radius = 8), NNVT([4, 2]), TempGrid{2}(500, 650)"""
    return (P.MCParams(cycles, n_traj, 6), P.TempGrid(t[0], t[1], n_traj), P.Config(BRASS6, P.SphericalBC(radius=8.0)),
            P.host.brass_runner_potential(4, 2), P.host.NNVT((4, 2), n_atom_swaps=n_swaps))


def test_brass_known_answers_on_the_device(oracle):
    """The reference's own two-species known answers (test/test_potentials.jl:86, test/test_states.jl:159,175) through the CUDA
    kernels: the replay stream is the index draw that selects the move type followed by Julia's Random.seed!(123) stream."""
    from ptmc_b200 import multi
    from parity import xoshiro_streams
    from test_runner_oracle import _raw_for_index
    seed123 = xoshiro_streams(np.array([multi.julia_seed_state(123)], dtype=np.uint64), 8)[0]
    mcp, temp, start, pot, ens = _brass6_deck()
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=5, rng_mode=abi.RNG_REPLAY)
    pos, box = initial_arrays(mcp, start, 1)
    half = np.uint64(1) << np.uint64(63)
    for kind in ("move", "swap"):
        slot = np.zeros((1, 2, 64), dtype=np.uint64)
        if kind == "move":
            slot[0, :, 0] = _raw_for_index(1, 7); slot[0, :, 1:4] = seed123[:3]; base = 5
        else:
            slot[0, :, 0] = _raw_for_index(7, 7); slot[0, :, 1:3] = seed123[:2]; base = 4
        for m in range(6):
            slot[0, :, base + 5 * m] = _raw_for_index(2, 7)
            slot[0, :, base + 5 * m + 1:base + 5 * m + 4] = half                  # rand() = 0.5: zero displacement
        lad = np.full((1, 8), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        out = []
        for cls in (Engine, oracle.OracleEngine):
            eng = cls(cfg)
            eng.upload_states(pos, box)
            e0 = eng.download_states(("en_tot",))["en_tot"].copy()
            eng.set_replay(slot, lad)
            eng.trace_enable(1)
            eng.set_hist_bounds(-1.0, 0.0)
            eng.run_cycles(1, 1, False)
            out.append((e0, eng.trace_download(), eng.download_states(), eng.runner_state(0, 0)))
            eng.close()
        (e0g, tg, sg, (gg, eg)), (e0c, tc, sc, (gc, ec)) = out
        assert e0g[0, 0] == pytest.approx(-0.261652899, rel=1e-8)                 # test/test_potentials.jl:86
        np.testing.assert_allclose(e0g, e0c, rtol=1e-13)
        assert np.array_equal(tg["move_index"], tc["move_index"]) and np.array_equal(tg["accepted"], tc["accepted"])
        want = -7.5971516e-5 if kind == "move" else -0.0017084901948                # test/test_states.jl:159 / :175
        assert tg["delta_en"][0, 0, 0, 0] == pytest.approx(want, rel=2e-8)
        np.testing.assert_allclose(tg["delta_en"], tc["delta_en"], rtol=1e-10, atol=1e-16)
        assert tg["move_index"][0, 0, 0, 0] == (1 if kind == "move" else 7 + (3 << 10) + (6 << 20))
        assert np.array_equal(sg["pos"], sc["pos"]) and np.array_equal(sg["counters"], sc["counters"])
        if kind == "swap":
            assert np.array_equal(sg["pos"][0, 0, 2], BRASS6[5]) and np.array_equal(sg["pos"][0, 0, 5], BRASS6[2])   # :185-186
        np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=1e-12)
        np.testing.assert_allclose(gg, gc, rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(eg, ec, rtol=1e-11, atol=1e-15)


def test_brass6_nnvt_replay_and_production(oracle):
    # many cycles of atom moves + swaps + exchanges on the small cluster: per-move trace against the oracle, then the production
    # stream (BATCH8) trajectory
    mcp, temp, start, pot, ens = _brass6_deck(n_traj=4, n_swaps=3, t=(500.0, 2500.0))
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=3, n_cycles=40, seed=29, n_adjust=5, max_displ=np.full((3, 4, 4), 0.6),
                        bounds=(-0.5, 0.0, None, None))
    compare_traces(gpu, cpu)
    swaps = gpu["trace"]["move_index"] > 1024
    assert swaps.sum() > 100 and gpu["trace"]["accepted"][swaps].sum() > 10 and gpu["trace"]["accepted"][~swaps].mean() > 0.2
    assert gpu["trace"]["exchange"][..., 0].sum() > 3
    res = _free_run_pair(_brass6_deck(n_traj=4, n_swaps=3, t=(500.0, 2500.0)), 3, 41, abi.RNG_BATCH8, 40, bounds=(-0.5, 0.0))
    _assert_same_trajectory(res, e_rtol=1e-11)
    res = _free_run_pair(_brass6_deck(n_traj=4, n_swaps=3, t=(500.0, 2500.0)), 2, 43, abi.RNG_XOSHIRO, 20, bounds=(-0.5, 0.0))
    _assert_same_trajectory(res, e_rtol=1e-11)


def test_brass55_nnvt_replay(oracle):
    # scripts/brass55.jl sizes: Cu50Zn5, 5 radial + 26 angular two-species symmetry functions, two 88-20-20-1 networks; atom moves
    # AND swaps (6 swap slots per cycle so that a short run holds many), temperatures high enough that swaps are accepted
    mcp, temp, start, pot, ens = P.host.deck_brass55(n_traj=3, cycles=2, n_swaps=6)
    temp = P.TempGrid(600.0, 3000.0, 3)
    gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=1, n_cycles=2, seed=19, n_adjust=1, max_displ=np.full((1, 3, 4), 0.35))
    worst = compare_traces(gpu, cpu)
    swaps = gpu["trace"]["move_index"] > 1024
    assert swaps.sum() >= 10 and gpu["trace"]["accepted"][swaps].sum() >= 1
    assert 0.1 < gpu["trace"]["accepted"][~swaps].mean() < 0.9
    print("brass55 max |dE err|", worst, "swaps", int(swaps.sum()), "accepted", int(gpu["trace"]["accepted"][swaps].sum()))


def test_brass55_state_matrices_and_production(oracle):
    from oracle import oracle as orc
    mcp, temp, start, pot, ens = P.host.deck_brass55(n_traj=2, cycles=1, n_swaps=4)
    temp = P.TempGrid(800.0, 3000.0, 2)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=1, seed=6, rng_mode=abi.RNG_BATCH8)
    pos, box = initial_arrays(mcp, start, 1)
    md = np.full((1, 2, 4), 0.3)
    engs = [Engine(cfg), orc.OracleEngine(cfg)]
    for e in engs:
        e.upload_states(pos, box, md)
    g0, e0 = engs[0].runner_state(0, 1)
    g1, e1 = engs[1].runner_state(0, 1)
    assert g0.shape == (55, 88)
    np.testing.assert_allclose(g0, g1, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(e0, e1, rtol=1e-12)
    for e in engs:
        e.seed(6)
        e.run_cycles(1, 1, False)
        e.sync()
    sg, sc = engs[0].download_states(), engs[1].download_states()
    assert np.array_equal(sg["counters"], sc["counters"]) and np.array_equal(sg["pos"], sc["pos"])
    np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=1e-11)
    moved = (sg["pos"] != pos).any(axis=-1)
    assert moved.sum() > 20
    for slot in range(2):
        gg, eg = engs[0].runner_state(0, slot)
        gc, ec = engs[1].runner_state(0, slot)
        np.testing.assert_allclose(gg, gc, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(eg, ec, rtol=1e-10, atol=1e-14)
    # the running state is a state: a fresh upload of the final configuration gives the same energies
    eng2 = Engine(cfg)
    eng2.upload_states(sg["pos"], sg["box"])
    np.testing.assert_allclose(sg["en_tot"], eng2.download_states(("en_tot",))["en_tot"], rtol=1e-11)
    for e in engs + [eng2]:
        e.close()
