"""Definitions of the golden replay cases (inputs only; outputs live in tests/golden/<name>.npz, written by
tests/tools/make_golden.py with the CPU oracle)."""
import numpy as np

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.host import initial_arrays, sweep_config

from parity import make_replay

CASES = ("ne13_elj_nvt", "cu55_eam_nvt", "ne32_cubic_npt", "ne216_rhombic_npt", "ne13_eljb_nvt", "ne32_rect_npt_separated",
         "ne32_rhombic_eljb_npt_separated")


def _ne32_cubic():
    a = 8.4
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pos = np.array([(b + [i, j, k]) * a for i in range(2) for j in range(2) for k in range(2) for b in base])
    return P.Config(pos, P.CubicBC(2 * a))


def build_case(name, oracle):
    """-> (SweepConfig, pos, box, slot_draws, ladder_draws, n_cycles, bounds, run_engine kwargs)"""
    if name == "ne13_elj_nvt":      # examples/ne13.jl geometry and potential, 8 of the 4..16 K temperatures
        deck, L, cycles, seed, eq, bounds = P.host.deck_ne13(n_traj=8, cycles=24), 2, 24, 1234, 6, (-0.0075, -0.0040, None, None)
        deck[0].n_adjust = 10
    elif name == "cu55_eam_nvt":    # scripts/cu55.jl EAM deck
        deck, L, cycles, seed, eq, bounds = P.host.deck_cu55_eam(n_traj=6, cycles=6), 1, 6, 7, 2, (-7.1, -6.8, None, None)
        deck[0].n_adjust = 4
    elif name == "ne32_cubic_npt":  # 32-atom fcc block, cubic cell, NPT: pairs sit exactly on the cut-off L^2/4
        start = _ne32_cubic()
        deck = (P.MCParams(16, 4, 32, n_adjust=8), P.TempGrid(20.0, 40.0, 4), start, P.ELJPotentialEven(P.host.ELJ_NE_COEFF),
                P.NPT(32, 101325 * 3.398928944382626e-14, False))
        v0 = P.get_volume(start.boundary_condition)
        L, cycles, seed, eq, bounds = 1, 16, 21, 4, (-0.032, -0.024, 0.8 * v0, 2.0 * v0)
    elif name == "ne216_rhombic_npt":   # scripts/pbc216.jl (BASELINE C4): fcc / hcp start, rhombic cell
        deck = P.host.deck_pbc216(n_traj=4, cycles=2)
        v0 = P.get_volume(deck[2][0].boundary_condition)
        L, cycles, seed, eq, bounds = 1, 2, 8, 0, (0.30, 0.40, 0.8 * v0, 2.0 * v0)
    elif name == "ne13_eljb_nvt":       # scripts/ne13_B.jl: ELJPotentialB (magnetic field), Angstrom units
        deck, L, cycles, seed, eq, bounds = P.host.deck_ne13_B(n_traj=6, cycles=24), 2, 24, 77, 6, (-0.02, 0.0, None, None)
        deck[0].n_adjust = 10
    elif name == "ne32_rect_npt_separated":   # NPT(separated_volume = true): xy / z / uniform volume moves, rectangular cell
        start = _ne32_cubic()
        start = P.Config(start.positions * [1, 1, 1.25], P.RectangularBC(start.boundary_condition.box_length, start.boundary_condition.box_length * 1.25))
        deck = (P.MCParams(16, 4, 32, n_adjust=8), P.TempGrid(20.0, 40.0, 4), start, P.ELJPotentialEven(P.host.ELJ_NE_COEFF),
                P.NPT(32, 101325 * 3.398928944382626e-14, True, n_volume_moves=3))
        v0 = P.get_volume(start.boundary_condition)
        L, cycles, seed, eq, bounds = 1, 16, 55, 4, (-0.05, 0.0, 0.8 * v0, 2.0 * v0)
    elif name == "ne32_rhombic_eljb_npt_separated":
        # ELJPotentialB in a rhombic cell (Angstrom) with separated volume moves: the reference's aliased tan_mat / new_tan_mat leaves
        # the tangents of rejected xy / z trials in the state; lattice perturbed by 1e-3 (no half-cell ties)
        a2b = 1.8897259886
        lat = _ne32_cubic()
        pos = (lat.positions @ np.array([[1, 0, 0], [0.5, 3 ** 0.5 / 2, 0], [0, 0, 1]])) / a2b
        pos = pos + np.random.default_rng(8).uniform(-1e-3, 1e-3, pos.shape)
        Lb = lat.boundary_condition.box_length / a2b
        start = P.Config(pos, P.RhombicBC(Lb, Lb))
        deck = (P.MCParams(16, 4, 32, n_adjust=8), P.TempGrid(20.0, 40.0, 4), start,
                P.ELJPotentialB(P.host.ELJB_NE_A, P.host.ELJB_NE_B, P.host.ELJB_NE_C),
                P.NPT(32, 101325 * 3.398928944382626e-14 * a2b ** 3, True, n_volume_moves=4))
        v0 = P.get_volume(start.boundary_condition)
        L, cycles, seed, eq, bounds = 1, 16, 91, 4, (0.2, 1.0, 0.8 * v0, 2.0 * v0)
    else:
        raise KeyError(name)
    mcp, temp, start, pot, ens = deck
    # start with step sizes that give ~50 % acceptance instead of the constructor's 0.1 (src/MCStates.jl:104)
    md = np.full((L, mcp.n_traj, 4), 0.1)
    md[:, :, 0] = {"ne13_elj_nvt": 0.9, "cu55_eam_nvt": 0.5, "ne32_cubic_npt": 0.7, "ne216_rhombic_npt": 0.6,
                   "ne13_eljb_nvt": 0.4, "ne32_rect_npt_separated": 0.7,
                   "ne32_rhombic_eljb_npt_separated": 0.35}[name]
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=seed, rng_mode=abi.RNG_REPLAY)
    pos, box = initial_arrays(mcp, start, L)
    slot, lad = make_replay(oracle, L, mcp.n_traj, cfg.n_steps, eq + cycles, seed)
    return cfg, pos, box, slot, lad, cycles, bounds, {"eq_cycles": eq, "max_displ": md}
