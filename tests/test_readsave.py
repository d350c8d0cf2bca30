"""Checkpoint wire formats (src/ReadSave.jl) and the save / restart / rdf paths of ptmc_run, on the CPU oracle engine."""
import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi, readsave


def test_julia_float_format():
    # Base.show(::Float64): shortest round-trip digits, fixed for 1e-5 < |x| < 1e6, always a fractional part
    cases = {1.0: "1.0", 0.1: "0.1", 100000.0: "100000.0", 1.0e6: "1.0e6", 1234567.8: "1.2345678e6", 0.0001: "0.0001",
             1.0e-5: "1.0e-5", -2.5e-7: "-2.5e-7", 3.16681196e-6: "3.16681196e-6", 19.0: "19.0", 0.0: "0.0",
             10.0539603053030038: "10.053960305303004", 1e21: "1.0e21", 123456.0: "123456.0", 5.0e10: "5.0e10"}
    for x, s in cases.items():
        assert readsave.jl_float(x) == s, (x, readsave.jl_float(x), s)
        assert float(readsave.jl_float(x)) == x


def test_params_file_matches_reference_layout():
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=25, cycles=1000)
    text = readsave.params_text(pot, ens, mcp, temp)
    lines = text.splitlines()
    # writeparams src/ReadSave.jl:24-32 (value row is a Float64 matrix), writeensemble :40-44, writepotential :67-72
    assert lines[0] == "cycles: sample_rate: n_traj: n_atoms: n_adjust n_bins min_acc: max_acc t_i: t_f"
    assert lines[1] == "1000.0 1.0 25.0 13.0 100.0 100.0 0.4 0.6 4.0 16.0"
    assert lines[2] == "ensemble n_atom_moves n_atom_swaps" and lines[3] == "NVT 13 13 0"
    assert lines[4].startswith("ELJPotentialEven{6, Float64} -10.5097942564988 989.725135614556 ")
    assert lines[4].split()[0][17] == "6"          # readpotential parses potinfovec[1,1][18]
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=4, cycles=10)
    lines = readsave.params_text(pot, ens, mcp, temp).splitlines()
    assert lines[2] == "ensemble n_atom_moves n_volume_moves n_atom_swaps pressure separated_volume"
    assert lines[3].split()[:5] == ["NPT", "216", "216", "1", "0"] and lines[3].split()[6] == "false"
    mcp, temp, start, pot, ens = P.host.deck_cu55_eam(n_traj=4, cycles=10)
    assert readsave.params_text(pot, ens, mcp, temp).splitlines()[4].startswith("EAM: 8.482 4.692 ")


def test_config_file_roundtrip_all_boundaries():
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(5, 3))
    for bc in (P.SphericalBC(10.053960305303004), P.CubicBC(12.5), P.RhombicBC(18.79, 15.35), P.RectangularBC(9.0, 11.0)):
        text = readsave.config_text(pos, bc, [0.1, 0.2, 0.3, 0.4], 1234, 56)
        lines = text.splitlines()
        assert lines[0] == "5" and lines[2].startswith("1  ") and len(lines) == 7
        cfg, md, ca, cv = readsave.read_checkpoint_config(text)
        assert np.array_equal(cfg.positions, pos) and md == [0.1, 0.2, 0.3] and (ca, cv) == (1234, 56)
        assert type(cfg.boundary_condition).__name__.endswith(type(bc).__name__) or isinstance(cfg.boundary_condition, type(bc))
        assert cfg.boundary_condition.box() == pytest.approx(bc.box(), rel=0, abs=0)
    assert readsave.config_text(pos, P.SphericalBC(2.0), [0.1] * 4, 0, 0).splitlines()[1] == "SphericalBC{Float64} r2: 4.0 0.1 0.1 0.1 0 0"
    assert readsave.config_text(pos, P.RhombicBC(2.0, 3.0), [0.1] * 4, 7, 8).splitlines()[1] == "RhombicBC{Float64} box_dims: 2.0, 3.0 0.1 0.1 0.1 7 8"


def test_read_reference_written_files(tmp_path):
    # a checkpoint directory spelled the way Julia's writedlm writes it (tabs between histogram counts, Float64 counts)
    (tmp_path / "params.data").write_text(
        "cycles: sample_rate: n_traj: n_atoms: n_adjust n_bins min_acc: max_acc t_i: t_f\n"
        "200.0 1.0 2.0 3.0 100.0 4.0 0.4 0.6 5.0 8.0\nensemble n_atom_moves n_atom_swaps\nNVT 3 3 0\n"
        "ELJPotentialEven{2, Float64} 1.0 -2.0\n")
    (tmp_path / "hist_info.data").write_text("e_min e_max v_min v_max Δ_E Δ_V Δ_r2\n-1.0 -0.5 0.0 0.0 0.16666666666666666 0.0 2.5\n", encoding="utf-8")
    (tmp_path / "histograms.data").write_text("0.0\t1.0\t2.0\t3.0\t4.0\t5.0\n5.0\t4.0\t3.0\t2.0\t1.0\t0.0\n")
    (tmp_path / "index.txt").write_text("150\n")
    for t in (1, 2):
        (tmp_path / f"configT{t}.xyz").write_text("3\nSphericalBC{Float64} r2: 50.0 0.11 0.1 0.1 %d 0\n1  0.0 0.0 0.0 \n1  1.0 0.0 0.0 \n1  0.0 1.5 0.0 \n" % (40 * t))
    mcp, temp, ens, pot = readsave.read_init(tmp_path)
    assert (mcp.mc_cycles, mcp.n_traj, mcp.n_atoms, mcp.n_bin, mcp.eq_cycles) == (200, 2, 3, 4, 0)
    assert isinstance(pot, P.ELJPotentialEven) and pot.coeff == (1.0, -2.0) and ens.n_atoms == 3
    assert temp.t_grid[0] == 5.0 and temp.t_grid[-1] == 8.0
    cfgs, md, ca, cv, res, rdf, index = readsave.rebuild_states(2, tmp_path)
    assert index == 150 and list(ca) == [40, 80] and md[0, 0] == 0.11 and rdf is None
    assert res.en_histogram.shape == (2, 6) and res.en_histogram[1, 0] == 5.0 and res.delta_r2 == 2.5
    assert cfgs[0].boundary_condition.radius2 == 50.0 and cfgs[1].positions[2, 1] == 1.5


def test_ev_histogram_is_column_major(tmp_path):
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=2, cycles=10)
    nb2 = 4
    res = P.host.Output(nb2 - 2)
    res.ev_histogram = np.arange(2 * nb2 * nb2, dtype=float).reshape(2, nb2, nb2)
    states = {"pos": np.zeros((1, 2, 3, 3)), "box": np.full((1, 2, 2), 5.0), "max_displ": np.full((1, 2, 4), 0.1),
              "counters": np.zeros((1, 2, abi.PTMC_NCOUNT), dtype=np.int64)}
    readsave.checkpoint(10, states, res, ens, abi.BC_RHOMBIC, directory=tmp_path)
    readsave.save_histparams(res, tmp_path)
    row = (tmp_path / "volume_hist.data").read_text().splitlines()[0].split("\t")
    # writedlm flattens evmat column-major; setresults :413-418 rebuilds evmat[:, index] from consecutive blocks
    assert [float(v) for v in row[:4]] == [0.0, 4.0, 8.0, 12.0]
    back, _ = readsave.setresults(tmp_path, npt=True)
    assert np.array_equal(back.ev_histogram, res.ev_histogram)


def test_ptmc_run_save_restart_and_rdf(tmp_path, oracle):
    from oracle import oracle as orc
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=4, cycles=2000)
    mcp.eq_percentage = 0.01
    d = tmp_path / "ck"
    st, res = P.host.ptmc_run(mcp, temp, start, pot, ens, n_ladders=1, seed=5, engine_cls=orc.OracleEngine, rdfsave=True,
                              save=1000, save_dir=d)
    assert sorted(p.name for p in d.iterdir()) == ["configT1.xyz", "configT2.xyz", "configT3.xyz", "configT4.xyz", "hist_info.data",
                                                   "histograms.data", "index.txt", "params.data", "rdf.data"]
    r = res[0]
    # two rdf samples (cycles 1000, 2000): N (N-1) / 2 pairs each, minus those in bin 0 / beyond the last bin
    assert r.rdf.shape == (4, 500) and 0 < r.rdf.sum() <= 2 * 4 * 78 and r.rdf[:, 0].sum() == 0
    assert r.en_histogram.sum() == 4 * 2000
    cfgs, md, ca, cv, back, rdf, index = readsave.rebuild_states(4, d)
    assert index == 2000 and np.array_equal(back.en_histogram, r.en_histogram) and np.array_equal(rdf, r.rdf)
    assert np.array_equal(np.stack([c.positions for c in cfgs]), st["pos"][0])
    assert np.array_equal(ca, st["counters"][0, :, abi.C_ATOM_TOT]) and np.array_equal(md, st["max_displ"][0, :, :3])
    assert back.en_min == r.en_min and back.delta_en_hist == r.delta_en_hist and back.delta_r2 == r.delta_r2
    # restart: 1000 more cycles continue the histograms and the counters
    mcp2, temp2, ens2, pot2 = readsave.read_init(d)
    mcp2.mc_cycles = 3000
    st2, res2 = P.host.ptmc_run(mcp2, temp2, cfgs, pot2, ens2, n_ladders=1, seed=5, engine_cls=orc.OracleEngine, rdfsave=True,
                                save=1000, save_dir=d, restart=True)
    # (start_counter = index: the reference's restart runs cycle 2000 again, src/Initialization.jl:108 -> 1001 sampled cycles)
    assert res2[0].en_histogram.sum() == 4 * 3001 and (res2[0].en_histogram >= r.en_histogram).all()
    assert (d / "hist_info.data").read_text().split()[7:] == [readsave.jl_float(v) for v in (r.en_min, r.en_max, 0.0, 0.0, r.delta_en_hist, 0.0, r.delta_r2)]
    assert (st2["counters"][0, :, abi.C_ATOM_TOT] > st["counters"][0, :, abi.C_ATOM_TOT]).all()
    assert readsave.rebuild_states(4, d)[-1] == 3000


def test_ptmc_run_save_restart_npt_separated(tmp_path, oracle):
    # NPT with separated volume moves on a rectangular cell: volume_hist.data, box_dims in the xyz comment line, restart
    from oracle import oracle as orc
    a = 8.4
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pos = np.array([(b + [i, j, k]) * a for i in range(2) for j in range(2) for k in range(2) for b in base]) * [1, 1, 1.25]
    start = P.Config(pos, P.RectangularBC(2 * a, 2.5 * a))
    mcp = P.MCParams(60, 3, 32, eq_percentage=0.1, n_adjust=10)
    temp = P.TempGrid(20.0, 40.0, 3)
    pot = P.ELJPotentialEven(P.host.ELJ_NE_COEFF)
    ens = P.NPT(32, 101325 * 3.398928944382626e-14, True, n_volume_moves=2)
    st, res = P.host.ptmc_run(mcp, temp, start, pot, ens, n_ladders=1, seed=3, engine_cls=orc.OracleEngine, save=30, save_dir=tmp_path)
    assert (tmp_path / "volume_hist.data").exists() and res[0].ev_histogram.sum() == 3 * 60
    line = (tmp_path / "configT2.xyz").read_text().splitlines()[1]
    assert line.startswith("RectangularBC{Float64} box_dims: ") and line.split()[2].endswith(",")
    cfgs, md, ca, cv, back, rdf, index = readsave.rebuild_states(3, tmp_path, npt=True)
    assert index == 60 and np.array_equal(back.ev_histogram, res[0].ev_histogram)
    assert np.array_equal(np.array([c.boundary_condition.box() for c in cfgs]), st["box"][0])
    assert np.array_equal(cv, st["counters"][0, :, abi.C_VOL_TOT]) and back.v_min == res[0].v_min and back.delta_v_hist == res[0].delta_v_hist
    mcp2, temp2, ens2, pot2 = readsave.read_init(tmp_path)
    assert ens2.separated_volume is True and ens2.n_volume_moves == 2 and ens2.pressure == ens.pressure
    mcp2.mc_cycles = 90
    st2, res2 = P.host.ptmc_run(mcp2, temp2, cfgs, pot2, ens2, n_ladders=1, seed=3, engine_cls=orc.OracleEngine, save=30, save_dir=tmp_path, restart=True)
    assert res2[0].ev_histogram.sum() == 3 * 91 and readsave.rebuild_states(3, tmp_path, npt=True)[-1] == 90   # cycle 60 runs again (start_counter = index)


def test_saveconfigs_files_like_the_reference_test(tmp_path, oracle):
    # test/ne13_test.jl ("Saveconfigs"): ptmc_run!(...; save = 1000, saveconfigs = k, configsname = "Configurations") writes
    # ./checkpoint/Configurations<i>T<t>.xyz whose atom lines are "1  " * join(position, " ") * " "
    from oracle import oracle as orc
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=5, cycles=300)
    st, res = P.host.ptmc_run(mcp, temp, start, pot, ens, n_ladders=1, seed=9, engine_cls=orc.OracleEngine, save=100, save_dir=tmp_path,
                              saveconfigs=150, configsname="Configurations")
    assert (tmp_path / "Configurations150T1.xyz").exists() and not (tmp_path / "Configurations100T1.xyz").exists()
    for t in range(1, 6):
        lines = (tmp_path / f"Configurations300T{t}.xyz").read_text().splitlines()
        assert lines[0] == "13" and lines[1].startswith("SphericalBC{Float64} r2: ")
        for a, line in enumerate(lines[2:]):
            assert line[:3] == "1  " and line[3:] == " ".join(readsave.jl_float(v) for v in st["pos"][0, t - 1, a]) + " "
            assert [float(x) for x in line.split()[1:]] == list(st["pos"][0, t - 1, a])


def test_params_file_eljb_lookup_and_nnvt(tmp_path, oracle):
    # writepotential(::AbstractDimerPotentialB) / (::LookupTablePotential) src/ReadSave.jl:70-76,85-87, writeensemble(::NNVT) :51-56,
    # readpotential's ELJB branch :266-272, readensemble's NNVT branch :238-242
    mcp, temp, start, pot, ens = P.host.deck_ne13_B(n_traj=4, cycles=60)
    text = readsave.params_text(pot, ens, mcp, temp)
    lines = text.splitlines()
    assert lines[4] == "ELJB 6 " and lines[5].split("\t") == [readsave.jl_float(v) for v in pot.coeff_a] and len(lines) == 8
    assert lines[7].split("\t")[2] == "35.6955"
    readsave.save_init(pot, ens, mcp, temp, tmp_path / "b")
    _, _, ens2, pot2 = readsave.read_init(tmp_path / "b")
    assert isinstance(pot2, P.ELJPotentialB) and pot2 == pot and ens2.atom_moves() == ens.atom_moves() == 13
    lut = P.LookupTablePotential(np.ones((91, 4)), start_dist=2.0, start_angle=0.0, d_dist=0.5, d_angle=1.0, c6coeff=-1.0)
    assert readsave.params_text(lut, ens, mcp, temp).splitlines()[4] == "LookupTablePotential "
    readsave.save_init(lut, ens, mcp, temp, tmp_path / "l")
    with pytest.raises(ValueError, match="LookupTablePotential"):
        readsave.read_init(tmp_path / "l")
    nn = P.host.NNVT((50, 5), n_atom_swaps=2)
    t2 = readsave.params_text(P.host.brass_runner_potential(50, 5), nn, P.MCParams(100, 10, 55), P.TempGrid(250.0, 900.0, 10)).splitlines()
    assert t2[2] == "ensemble n_1 n_2 n_atom_moves n_atom_swaps" and t2[3] == "NNVT 50 5 55 2" and t2[4] == "runnerpotential "
    readsave.save_init(P.host.brass_runner_potential(50, 5), nn, P.MCParams(100, 10, 55), P.TempGrid(250.0, 900.0, 10), tmp_path / "n")
    _, _, ens3, pot3 = readsave.read_init(tmp_path / "n")
    assert tuple(ens3.natoms) == (50, 5) and ens3.n_atom_swaps == 2 and ens3.atom_moves() == 55 and pot3 == "runnerpotential"
    # save -> restart of an ELJPotentialB run on the oracle engine
    from oracle import oracle as orc
    d = tmp_path / "run"
    st, res = P.host.ptmc_run(mcp, temp, start, pot, ens, n_ladders=1, seed=8, engine_cls=orc.OracleEngine, save=30, save_dir=d)
    mcp2, temp2, ens2, pot2 = readsave.read_init(d)
    cfgs = readsave.rebuild_states(4, d)[0]
    mcp2.mc_cycles = 90
    st2, res2 = P.host.ptmc_run(mcp2, temp2, cfgs, pot2, ens2, n_ladders=1, seed=8, engine_cls=orc.OracleEngine, save=30, save_dir=d, restart=True)
    assert res2[0].en_histogram.sum() == 4 * 91 and readsave.rebuild_states(4, d)[-1] == 90
