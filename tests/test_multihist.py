"""multihistogram_NVT (host-side restatement of src/multihist_NVT.jl) against the reference's own known-answer vector
(test/multihist_test.jl:4-14, extracted by tools/extract_multihist_kat.py), and — on the GPU — heat-capacity curves
from CUDA histograms against curves from CPU-oracle histograms (north star: "within statistical error")."""
import json
import pathlib

import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.multihist import multihistogram_NVT

GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def test_reference_known_answer():
    k = json.loads((GOLD / "multihist_nvt_kat.json").read_text())
    temp = P.TempGrid(*k["temp"][:2], k["temp"][2])
    t, cv = multihistogram_NVT(temp.t_grid, temp.beta_grid, k["en_histogram"], k["en_min"], k["en_max"], k["n_bin"], k["conv_threshold"])
    assert len(cv) == 320
    for idx, want in k["cv_expected_1based"].items():
        assert cv[int(idx) - 1] == pytest.approx(want, rel=1e-9)      # Julia sums sequentially; we sum with numpy


def test_reference_known_answer_npt():
    from ptmc_b200.multihist import multihistogram_NPT
    k = json.loads((GOLD / "multihist_npt_kat.json").read_text())
    temp = P.TempGrid(*k["temp"][:2], k["temp"][2])
    t, cp = multihistogram_NPT(temp.t_grid, temp.beta_grid, k["ev_histogram"], k["en_min"], k["en_max"], k["v_min"], k["v_max"],
                               k["n_bin"], k["pressure"], k["conv_threshold"])
    assert len(cp) == 30
    for idx, want in k["cp_expected_1based"].items():
        assert cp[int(idx) - 1] == pytest.approx(want, rel=1e-9)


BOUNDS = (-0.00565, -0.0024)     # fixed energy window (the reference's own Ne13 run: -0.005638 .. -0.002419, multihist_test.jl:7)


def _cv_curve(engine_cls, n_ladders, cycles, seed, engine_kwargs=None):
    """Equilibrate, then sample into the SAME histogram window for every engine (ptmc_run! takes the window from a
    snapshot of the energies, src/MCRun.jl:299-310, which would differ between two independent runs)."""
    from ptmc_b200.host import initial_arrays, sweep_config
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=16, cycles=cycles)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=n_ladders, seed=seed)
    eng = engine_cls(cfg, **(engine_kwargs or {}))
    pos, box = initial_arrays(mcp, start, n_ladders)
    eng.upload_states(pos, box)
    eng.seed(seed)
    eng.run_cycles(1, 1000, False)
    eng.reset_counters()
    eng.set_hist_bounds(*BOUNDS)
    eng.run_cycles(1, cycles, True)
    eng.sync()
    hist = eng.download_hist(reduce_ladders=True)
    st = eng.download_states(("ham",))
    eng.close()
    t, cv = multihistogram_NVT(temp.t_grid, temp.beta_grid, hist, BOUNDS[0], BOUNDS[1], mcp.n_bin)
    e1, e2 = st["ham"][..., 0].mean(axis=0) / cycles, st["ham"][..., 1].mean(axis=0) / cycles
    return t, cv, (e2 - e1 ** 2) * temp.beta_grid ** 2, hist


@pytest.mark.gpu
def test_heat_capacity_curves_agree(oracle):
    # Ne13, 16 temperatures 4..16 K (examples/ne13.jl).  CUDA: 256 ladders x 4000 cycles; CPU oracle: 32 ladders x 4000 cycles
    # with different seeds -> two independent estimates of the same curve.
    t, cv_gpu, hc_gpu, h_gpu = _cv_curve(P.Engine, 256, 4000, 11)
    _, cv_cpu, hc_cpu, h_cpu = _cv_curve(oracle.OracleEngine, 32, 4000, 977, {"threads": 8})
    assert h_gpu.sum() == 256 * 16 * 4000 and h_cpu.sum() == 32 * 16 * 4000
    assert h_gpu[:, -1].sum() < 1e-3 * h_gpu.sum()                          # the window holds (nearly) every sample
    peak = cv_cpu.max()
    print("cv peak", peak, "max |diff|", np.abs(cv_gpu - cv_cpu).max())
    assert np.abs(cv_gpu - cv_cpu).max() < 0.08 * peak                      # statistical error of the smaller sample
    assert abs(t[np.argmax(cv_gpu)] - t[np.argmax(cv_cpu)]) < 0.8           # melting peak near 10 K
    assert np.abs(hc_gpu - hc_cpu).max() < 0.1 * hc_cpu.max()               # fluctuation estimate (E, E^2 sums), pooled over ladders


def test_direct_multihistogram_and_postprocess(tmp_path, oracle):
    # src/multihist.jl (the analysis behind postprocess(), which the reference's "Saveconfigs" test runs on a checkpoint): no
    # known-answer vector exists for it, so it is checked for self-consistency on an oracle run with a histogram window that
    # holds every sample: U(T) at the ladder temperatures equals the sampled <E>, the Cv curve agrees with the fluctuation
    # formula and with the iterative multihistogram_NVT
    from oracle import oracle as orc
    from ptmc_b200 import abi, multihist, readsave
    from ptmc_b200.host import initial_arrays, sweep_config, finalise_results
    import ptmc_b200 as P
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=8, cycles=6000)
    L, bounds = 4, (-0.0058, -0.0022)
    eng = orc.OracleEngine(sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=11), threads=4)
    pos, box = initial_arrays(mcp, start, L)
    eng.upload_states(pos, box); eng.seed(11)
    eng.run_cycles(1, 1500, False); eng.reset_counters(); eng.set_hist_bounds(*bounds)
    eng.run_cycles(1, mcp.mc_cycles, True)
    st, hl = eng.download_states(), eng.download_hist()
    eng.close()
    assert hl[:, :, 0].sum() + hl[:, :, -1].sum() <= 2e-4 * hl.sum()   # (practically) nothing outside the window
    hist = hl.sum(axis=0)
    res = P.Output(mcp.n_bin, en_min=bounds[0], en_max=bounds[1])
    res.delta_en_hist = (res.en_max - res.en_min) / (res.n_bin - 1)
    res.en_histogram = hist
    en_avg = st["ham"][:, :, 0].mean(axis=0) / mcp.mc_cycles
    cv_fluct = ((st["ham"][:, :, 1] / mcp.mc_cycles - (st["ham"][:, :, 0] / mcp.mc_cycles) ** 2) * temp.beta_grid ** 2).mean(axis=0) * multihist.K_B
    out = multihist.multihistogram(hist, temp.beta_grid, res.en_min, res.en_max, mcp.n_bin)
    u_at = np.interp(temp.t_grid[1:-1], out["T"], out["U"])
    # (multihist.jl takes the LEFT edge of a bin as its energy, readfile :58, so U(T) sits about half a bin below the sampled mean)
    assert np.abs(u_at - en_avg[1:-1]).max() < 2 * res.delta_en_hist
    cv_at = np.interp(temp.t_grid[1:-1], out["T"], out["Cv"])
    assert np.abs(cv_at - cv_fluct[1:-1]).max() < 0.15 * cv_fluct.max()
    t2, cv2 = multihist.multihistogram_NVT(temp.t_grid, temp.beta_grid, hist, res.en_min, res.en_max, mcp.n_bin)
    assert np.abs(np.interp(t2[5:-5], out["T"], out["Cv"]) - cv2[5:-5]).max() < 0.1 * cv2.max()
    # postprocess() on a checkpoint directory of that run: files and return values
    ckpt = tmp_path / "checkpoint"                       # the reference: pwd()/checkpoint, analysis files in pwd()
    readsave.save_init(pot, ens, mcp, temp, ckpt)
    readsave.save_histparams(res, ckpt)
    readsave.checkpoint(mcp.mc_cycles, st, res, ens, abi.BC_SPHERICAL, directory=ckpt)
    raw = (ckpt / "histograms.data").read_text()
    for _ in range(2):                                   # twice: the analysis must not touch the checkpoint it reads
        energies, histdata, T, Z, Cv, dCv, S = readsave.postprocess(ckpt)
        assert len(histdata) == 8 and len(T) == 1000 and (Z >= 1.0).all() and (Z <= 100.0).all()
        assert np.allclose(Cv, out["Cv"]) and np.allclose(S, out["U"])
    assert (tmp_path / "analysis.NVT").read_text().splitlines()[0].split() == ["T", "Z", "Cv", "dCv", "U(T)"]
    assert (tmp_path / "Sol.X").exists() and (tmp_path / "S.data").exists() and (tmp_path / "histograms.data").exists()
    assert (ckpt / "histograms.data").read_text() == raw and not (ckpt / "analysis.NVT").exists()
    cfgs, md, ca, cv, res2, rdf, index = readsave.rebuild_states(8, ckpt)          # ... and a restart still reads the raw counts
    assert res2.n_bin == mcp.n_bin and np.array_equal(res2.en_histogram, hist) and index == mcp.mc_cycles
    with pytest.raises(ValueError, match="outdir"):
        readsave.postprocess(ckpt, outdir=ckpt)
