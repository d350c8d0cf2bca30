"""Committed golden trajectories (tests/golden/*.npz, written by tests/tools/make_golden.py with the CPU oracle).

CPU (`-m "not gpu"`): the oracle still reproduces them (guards the checker itself against drift).
GPU (`-m gpu`): the CUDA library reproduces them through the C ABI without the oracle in the loop."""
import pathlib

import numpy as np
import pytest

from golden_cases import CASES, build_case
from parity import EPS, run_engine

GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def _check(name, res, exact_positions):
    g = np.load(GOLD / f"{name}.npz")
    tr, st = res["trace"], res["states"]
    assert np.array_equal(tr["move_index"], g["move_index"])
    assert np.array_equal(tr["accepted"], g["accepted"]), "accept/reject decisions differ from the golden trajectory"
    assert np.array_equal(tr["exchange"], g["exchange"]), "exchange decisions differ from the golden trajectory"
    e_scale = max(np.abs(g["en_tot"]).max(), np.abs(g["init_en"]).max())
    tol = 1e-10 * np.abs(g["delta_en"]) + 256 * EPS * e_scale
    assert (np.abs(tr["delta_en"] - g["delta_en"]) <= tol).all()
    assert np.array_equal(st["counters"], g["counters"])
    assert np.array_equal(st["max_displ"], g["max_displ"])
    assert np.array_equal(res["hist"], g["hist"])
    if exact_positions:
        assert np.array_equal(st["pos"], g["pos"])
    else:   # NPT: the volume scale goes through exp/pow of another libm -> last-bit differences
        np.testing.assert_allclose(st["pos"], g["pos"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(st["en_tot"], g["en_tot"], rtol=1e-11, atol=1024 * EPS * e_scale)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name, oracle):
    cfg, pos, box, slot, lad, n_cycles, bounds, kw = build_case(name, oracle)
    res = run_engine(oracle.OracleEngine, cfg, pos, box, slot, lad, n_cycles, bounds, **kw)
    _check(name, res, exact_positions=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(name, oracle):
    from ptmc_b200.engine import Engine
    # the oracle module is used here only to regenerate the replay streams from the seed (Xoshiro256++)
    cfg, pos, box, slot, lad, n_cycles, bounds, kw = build_case(name, oracle)
    res = run_engine(Engine, cfg, pos, box, slot, lad, n_cycles, bounds, **kw)
    _check(name, res, exact_positions="npt" not in name)
