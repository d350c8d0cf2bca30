"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the sweep path
(SURVEY §8c).  Values are quoted from /root/reference/test/*.jl (file:line in each test)."""
import math

import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.host import sweep_config, initial_arrays


def test_elj_potentials_exact(oracle):
    # test/test_potentials.jl:3-15
    assert oracle.dimer_energy_elj([-2.0, 0.0, 1.0], 1.0) == -1.0
    assert oracle.dimer_energy_eljeven([-2.0, 1.0], 1.0) == -1.0
    assert oracle.dimer_energy_eljeven([-2.0, 1.0], 2.0) == oracle.dimer_energy_elj([-2.0, 0.0, 1.0], 2.0)
    c = [-1.0, 2.0, 3.0, -4.0, 5.0]
    assert oracle.dimer_energy_elj(c, 1.0) == sum(c)
    assert oracle.dimer_energy_eljeven([-2.0, 1.0], 1.0) == sum([-2.0, 1.0])


def test_julia_rng_stream(oracle):
    # Random.seed!(1234); rand() x3 -> the trial position of test/test_states.jl:27-38
    s = oracle.julia_seed(1234)
    u = [oracle.julia_rand(s) for _ in range(3)]
    trial = np.array([1.0, 2.0, 3.0]) + (np.array(u) - 0.5) * 0.1
    assert np.allclose(trial, [0.98259767, 2.00490511, 2.97185867], atol=5e-9)
    # Random.seed!(123); rand(1:4), rand(5:6) -> swap indices (3,6) asserted by test/test_states.jl:179-186
    s = oracle.julia_seed(123)
    assert oracle.julia_rand_range(s, 4) == 3
    assert 4 + oracle.julia_rand_range(s, 2) == 6


def test_xoshiro_reference_vector(oracle):
    # xoshiro256++ with state (1,2,3,4): first outputs from the generator's published reference code
    s = np.array([1, 2, 3, 4], dtype=np.uint64)
    out = [oracle.xoshiro_next(s) for _ in range(3)]
    assert out == [41943041, 58720359, 3588806011781223]


def test_distance2_exact(oracle):
    # test/runtests.jl:86-100 (Cubic), :123-140 (Rhombic)
    v1, v2, v3 = [1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [3.0, 6.0, 9.0]
    assert oracle.distance2(abi.BC_CUBIC, v1, v2, 10.0) == 14.0
    assert oracle.distance2(abi.BC_SPHERICAL, v1, v3) == 56.0
    assert oracle.distance2(abi.BC_CUBIC, v1, v3, 10.0) == 36.0
    assert oracle.distance2(abi.BC_RHOMBIC, v1, v2, 10.0, 10.0) == 14.0
    assert oracle.distance2(abi.BC_RHOMBIC, v1, v3, 10.0, 10.0) == 36.0
    v4 = [15.0, 5.0 * 3 ** 0.5, 2.0]
    assert oracle.distance2(abi.BC_RHOMBIC, v1, v4, 10.0, 10.0) == 6.0


def test_volume_and_rcut(oracle):
    # test/runtests.jl:198-206, :34, :42, :47
    assert oracle.get_volume(abi.BC_CUBIC, 10.0) == 1000.0
    assert oracle.get_volume(abi.BC_RHOMBIC, 10.0, 10.0) == 3 ** 0.5 / 2 * 1000.0
    assert oracle.get_r_cut(abi.BC_CUBIC, 8.7674) == 8.7674 ** 2 / 4
    assert oracle.get_r_cut(abi.BC_RHOMBIC, 10.0, 10.0) == 10.0 ** 2 * 3 / 16
    assert oracle.get_r_cut(abi.BC_RHOMBIC, 10.0, 5.0) == 5.0 ** 2 / 4


def test_tempgrid(oracle):
    # test/runtests.jl:216-235
    t, b = oracle.tempgrid(2, 16, 32)
    assert t[0] == pytest.approx(2.0) and t[-1] == pytest.approx(16.0)
    assert 1.0 / (t[0] * b[0]) == pytest.approx(3.16681196e-6)
    g = P.TempGrid(2, 16, 32)
    assert np.array_equal(g.t_grid, t) and np.array_equal(g.beta_grid, b)
    with pytest.raises(ValueError):
        P.TempGrid(2, 16, 256, tdistr="test")


def _three_atom(bc, ens, pot, oracle, n_traj=2):
    conf = P.Config([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.0, 1.0, 0.0]], bc)
    mcp = P.MCParams(10, n_traj, 3)
    temp = P.TempGrid(10, 15, n_traj)
    cfg = sweep_config(mcp, temp, conf, pot, ens, rng_mode=abi.RNG_XOSHIRO)
    eng = oracle.OracleEngine(cfg)
    pos, box = initial_arrays(mcp, conf, 1)
    eng.upload_states(pos, box)
    return eng


def test_eam_three_atom_energy(oracle):
    # test/test_potentials.jl:41-56
    pot = P.EmbeddedAtomPotential(1.0, 1.0, 1.0, 1.0, 1.0)
    assert pot.ean == 0.5
    assert oracle.eam_ctor(1.0, 1.0, 1.0, 1.0, 1.0) == (pot.ean, pot.eCam)
    eng = _three_atom(P.SphericalBC(radius=2.0), P.NVT(3), pot, oracle)
    en = eng.download_states(("en_tot",))["en_tot"]
    assert en[0, 0] == pytest.approx(-1.3495549581716526, rel=1e-15)


@pytest.mark.parametrize("npt", [False, True])
def test_seeded_atom_move_delta_energy(oracle, npt):
    # test/test_states.jl:27-38 (NVT, SphericalBC(5)) and :84-95 (NPT, CubicBC(10)): Random.seed!(1234)
    pot = P.ELJPotentialEven((-2.0, 1.0))
    if npt:
        eng = _three_atom(P.CubicBC(10.0), P.NPT(3, 101325 * 3.398928944382626e-14, False), pot, oracle)
    else:
        eng = _three_atom(P.SphericalBC(radius=5.0), P.NVT(3), pot, oracle)
    st = np.tile(oracle.julia_seed(1234), (1, 2, 1))
    eng.set_rng_states(st, oracle.julia_seed(1)[None])
    trial, de = eng.probe_atom_move(0, 0, 1)
    assert de == pytest.approx(-4.99895252855e-5, rel=1e-9)  # Julia's ≈ is rtol sqrt(eps); we match to 1e-9
    assert np.allclose(trial, [0.98259767, 2.00490511, 2.97185867], atol=5e-9)
    d2 = eng.dist2_mat(0, 0)
    assert d2[0, 1] == 14.0 and d2[1, 0] == 14.0 and d2[0, 0] == 0.0


def test_lrc_matches_closed_form(oracle):
    # lrc src/EnergyEvaluation.jl:968-982: zero below r_cut^2 = 50, else the uniform-density tail integral
    c = list(P.host.ELJ_NE_COEFF)
    assert oracle.lrc_eljeven(c, 216, 50.0) == 0.0
    rc2 = 151.9
    rc = math.sqrt(rc2)
    expect = sum(ci / rc ** (2 * i + 1 + 2) / (2 * i + 1) for i, ci in enumerate(c, start=1))
    expect *= math.pi * 216 ** 2 / 4 / rc ** 3 * rc ** 2  # rc3 starts at r_cut^(3/2): c_i / r^(2i+1) / (2i+1)
    got = oracle.lrc_eljeven(c, 216, rc2)
    ref = 0.0
    rc3 = rc2 * rc
    for i, ci in enumerate(c, start=1):
        ref += ci / rc3 / (2 * i + 1)
        rc3 *= rc2
    ref *= math.pi * 216 ** 2 / 4 / rc ** 3
    assert got == pytest.approx(ref, rel=1e-14)


def test_eljb_rhombic_three_atom_energy(oracle):
    # test/separated_v_test.jl:90-113: ELJPotentialB, RhombicBC(10, 10), NPT -> tan_mat[1,2], tan_mat[1,3] and
    # state.en_tot == -0.00016263185592172208 (pair energies within r_cut = min(3/16 L^2, h^2/4) + lrc * 3/4 * L/h,
    # src/EnergyEvaluation.jl:519-543)
    from ptmc_b200 import abi
    from ptmc_b200.host import ELJB_NE_A, ELJB_NE_B, ELJB_NE_C
    v = [np.array([1.0, 2.0, 3.0]), np.array([2.0, 4.0, 6.0]), np.array([0.0, 1.0, 0.0])]
    L = h = 10.0
    t12 = oracle.get_tan(abi.BC_RHOMBIC, v[0], v[1], L, h)
    t13 = oracle.get_tan(abi.BC_RHOMBIC, v[0], v[2], L, h)
    assert t12 == pytest.approx(0.7453559924999299, rel=1e-15) and t13 == pytest.approx(0.47140452079103173, rel=1e-15)
    r_cut = oracle.get_r_cut(abi.BC_RHOMBIC, L, h)
    e = 0.0
    for i in range(3):
        for j in range(i + 1, 3):
            d2 = oracle.distance2(abi.BC_RHOMBIC, v[i], v[j], L, h)
            if d2 <= r_cut:
                e += oracle.dimer_energy_eljb(ELJB_NE_A, ELJB_NE_B, ELJB_NE_C, d2, oracle.get_tan(abi.BC_RHOMBIC, v[i], v[j], L, h))
    e = e + oracle.lrc_eljb(3, r_cut) * 3 / 4 * L / h
    assert e == -0.00016263185592172208


REF_LUT = "/root/reference/scripts/lookup-tables"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_LUT), reason="the reference tree (its lookup tables) is not on this machine")
def test_lookup_table_reader_and_energy(oracle):
    # test/test_potentials.jl:58-65: LookupTablePotential(look-up_table-2.txt): table[1][1], start_dist, start_angle, size
    pot = P.LookupTablePotential(REF_LUT + "/look-up_table-2.txt")
    assert pot.table[0, 0] == 282.19449125205114 and pot.start_dist == 0.1 and pot.start_angle == 0
    assert pot.table.size == pot.l_angle * pot.l_dist == 91 * 1491
    # c6coeff (src/EnergyEvaluation.jl:252-258) and the three branches of dimer_energy (:274-295) against a direct evaluation
    avg = 0.0
    for i in range(pot.l_angle):
        avg += pot.table[i, 999]
    assert pot.c6coeff == avg / pot.l_angle * (pot.start_dist + 1000 * pot.d_dist) ** 6
    r2, tan = 3.21 ** 2, 0.75
    ai = int(round(math.atan(tan) / pot.d_angle * 180.0 / math.pi + 1.0))
    di = int(round((r2 ** 0.5 - pot.start_dist) / pot.d_dist))
    assert oracle.dimer_energy_lut(pot, r2, tan) == pot.table[ai - 1, di - 1]
    assert oracle.dimer_energy_lut(pot, 0.01 ** 2, 0.001) == pot.table[0, 0]                       # inside the first distance bin, angle bin 1
    assert oracle.dimer_energy_lut(pot, 4.0, 1e3) == pot.table[90, int(round((2.0 - pot.start_dist) / pot.d_dist)) - 1]   # |tan| > 114.59 -> angle bin 91
    far = (pot.start_dist + pot.l_dist * pot.d_dist + 1.0) ** 2
    assert oracle.dimer_energy_lut(pot, far, 0.3) == pot.c6coeff / far ** 3


def test_tangent_known_answers(oracle):
    # test/runtests.jl:168-197 "Tangent" (get_tan returns absolute values, so the reference marks the signed ones broken)
    v1, v2, v3 = [5.0, 0.0, 0.0], [-3.0, 0.0, 4.0], [-2.0, 0.0, -3.0]
    assert oracle.get_tan(abi.BC_SPHERICAL, v1, v2) == 2.0 and oracle.get_tan(abi.BC_SPHERICAL, v1, v3) == 7 / 3 and oracle.get_tan(abi.BC_SPHERICAL, v2, v3) == 1 / 7
    assert oracle.get_tan(abi.BC_CUBIC, v1, v2, 10.0, 10.0) == 1 / 2 and oracle.get_tan(abi.BC_CUBIC, v1, v3, 10.0, 10.0) == 1.0
    assert oracle.get_tan(abi.BC_CUBIC, v2, v3, 10.0, 10.0) == 1 / 3
    assert oracle.get_tan(abi.BC_RHOMBIC, v1, v2, 5.0, 5.0) == 2.0 and oracle.get_tan(abi.BC_RHOMBIC, v2, v3, 5.0, 5.0) == 0.5
    assert oracle.get_tan(abi.BC_RHOMBIC, v1, v3, 5.0, 5.0) == 1.0


def test_eljb_rhombic_state_energy_through_the_engine(oracle):
    # the same known answer (test/separated_v_test.jl:90-111) through the oracle's MCState construction for a periodic NPT state
    from ptmc_b200.host import ELJB_NE_A, ELJB_NE_B, ELJB_NE_C
    start = P.Config(np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.0, 1.0, 0.0]]), P.RhombicBC(10.0, 10.0))
    mcp, temp = P.MCParams(10, 2, 3), P.TempGrid(10.0, 15.0, 2)
    ens = P.NPT(3, 101325 * 3.398928944382626e-14, False)
    eng = oracle.OracleEngine(sweep_config(mcp, temp, start, P.ELJPotentialB(ELJB_NE_A, ELJB_NE_B, ELJB_NE_C), ens))
    pos, box = initial_arrays(mcp, start, 1)
    eng.upload_states(pos, box)
    assert (eng.download_states(("en_tot",))["en_tot"] == -0.00016263185592172208).all()
    eng.close()
