"""Shared parity harness: drives libptmc_b200.so (CUDA) and the CPU oracle through the same Engine calls with
the same pre-drawn random streams (SURVEY §8c "replayed-stream trajectories") and compares what comes back.

Tolerances:
  * decisions (move index, accept bit, exchange gate / pair / accept), counters, step sizes, histograms and
    positions: bit-identical;
  * per-move delta energy: |dE_gpu - dE_cpu| <= 1e-10 * |dE_cpu| + 256 * eps * max|E_tot|
    (north_star: 1e-10 relative; the absolute floor is the rounding of a sum of |E_tot| magnitude, which the
    reference's own new_en - en_tot subtraction already carries)."""
from __future__ import annotations

import numpy as np

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.engine import Engine
from ptmc_b200.host import initial_arrays, sweep_config

EPS = np.finfo(np.float64).eps
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _rotl(x, k):
    return (x << np.uint64(k)) | (x >> np.uint64(64 - k))


def xoshiro_streams(states: np.ndarray, length: int) -> np.ndarray:
    """Vectorised Xoshiro256++: states [n][4] uint64 -> raw draws [n][length]."""
    s = [states[:, k].copy() for k in range(4)]
    out = np.empty((states.shape[0], length), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(length):
            out[:, i] = _rotl(s[0] + s[3], 23) + s[0]
            t = s[1] << np.uint64(17)
            s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = _rotl(s[3], 45)
    return out


def make_replay(oracle, n_ladders, n_traj, n_steps, n_cycles, seed, slack=4):
    """Pre-drawn per-slot and per-ladder streams, seeded like Julia's Random.seed!(seed + k)."""
    R = n_ladders * n_traj
    slot_states = np.stack([oracle.julia_seed(seed + r) for r in range(R)])
    lad_states = np.stack([oracle.julia_seed(seed + R + l) for l in range(n_ladders)])
    slot_len = n_cycles * n_steps * (5 + slack) + 64
    lad_len = n_cycles * 4 + 16
    slot = xoshiro_streams(slot_states, slot_len).reshape(n_ladders, n_traj, slot_len)
    lad = xoshiro_streams(lad_states, lad_len).reshape(n_ladders, lad_len)
    return slot, lad


def run_engine(engine_cls, cfg, pos, box, slot_draws, lad_draws, n_cycles, bounds, *, eq_cycles=0,
               max_displ=None, engine_kwargs=None, chunks=1):
    """upload -> (equilibration cycles) -> histogram bounds -> sampled cycles, all traced."""
    eng = engine_cls(cfg, **(engine_kwargs or {}))
    try:
        eng.upload_states(pos, box, max_displ)
        init = eng.download_states(("en_tot",))["en_tot"].copy()
        eng.set_replay(slot_draws, lad_draws)
        eng.trace_enable(eq_cycles + n_cycles)
        if eq_cycles:
            eng.run_cycles(1, eq_cycles, False)
            eng.reset_counters()
        eng.set_hist_bounds(*bounds)
        per = n_cycles // chunks
        i = 1
        for k in range(chunks):
            n = per if k < chunks - 1 else n_cycles - per * (chunks - 1)
            eng.run_cycles(i, n, True)
            i += n
        eng.sync()
        return {"init_en": init, "states": eng.download_states(), "hist": eng.download_hist(),
                "hist_reduced": eng.download_hist(reduce_ladders=True), "trace": eng.trace_download(),
                "lanes": eng.lanes_per_replica() if engine_cls is Engine else 0,
                "launches": eng.launch_count() if engine_cls is Engine else 0}
    finally:
        eng.close()


def default_bounds(oracle, cfg, pos, box, start, ens, width=0.05):
    """Histogram bounds shared by both engines: a window around the oracle's initial energies (the host
    normally derives them from the equilibrated energies, src/MCRun.jl:293-310), and the reference's
    0.8 V0 .. 2 V0 volume window for NPT (src/Sampling.jl:177-183,203-211)."""
    eng = oracle.OracleEngine(cfg)
    try:
        eng.upload_states(pos, box)
        init = eng.download_states(("en_tot",))["en_tot"]
    finally:
        eng.close()
    lo, hi = float(init.min()), float(init.max())
    pad = width * (abs(lo) + abs(hi)) + 1e-12
    if ens.kind == abi.ENS_NPT:
        v0 = P.get_volume((start[0] if isinstance(start, (list, tuple)) else start).boundary_condition)
        return (lo - pad, hi + pad, 0.8 * v0, 2.0 * v0)
    return (lo - pad, hi + pad, None, None)


def run_pair(mcp, temp, start, pot, ens, *, n_ladders=1, n_cycles=10, seed=1234, eq_cycles=0, bounds=None,
             lanes=0, chunks=1, max_displ=None, n_adjust=None, slack=4):
    """Same deck, same streams, through the CUDA library and through the oracle."""
    from oracle import oracle
    if n_adjust is not None:
        mcp.n_adjust = n_adjust
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=n_ladders, seed=seed, rng_mode=abi.RNG_REPLAY,
                       lanes_per_replica=lanes)
    pos, box = initial_arrays(mcp, start, n_ladders)
    if bounds is None:
        bounds = default_bounds(oracle, cfg, pos, box, start, ens)
    slot, lad = make_replay(oracle, n_ladders, mcp.n_traj, cfg.n_steps, eq_cycles + n_cycles, seed, slack=slack)
    gpu = run_engine(Engine, cfg, pos, box, slot, lad, n_cycles, bounds, eq_cycles=eq_cycles, max_displ=max_displ, chunks=chunks)
    cpu = run_engine(oracle.OracleEngine, cfg, pos, box, slot, lad, n_cycles, bounds, eq_cycles=eq_cycles, max_displ=max_displ, chunks=chunks)
    return gpu, cpu


def compare_traces(gpu, cpu, *, rel=1e-10, positions_exact=True):
    tg, tc = gpu["trace"], cpu["trace"]
    assert tg["n_cycles"] == tc["n_cycles"]
    assert np.array_equal(tg["move_index"], tc["move_index"]), "move indices differ"
    assert np.array_equal(tg["exchange"], tc["exchange"]), "exchange decisions differ"
    bad = np.argwhere(tg["accepted"] != tc["accepted"])
    assert bad.size == 0, f"accept decisions differ first at {bad[0]}"
    e_scale = max(np.abs(cpu["states"]["en_tot"]).max(), np.abs(cpu["init_en"]).max())
    tol = rel * np.abs(tc["delta_en"]) + 256 * EPS * e_scale
    err = np.abs(tg["delta_en"] - tc["delta_en"])
    worst = np.unravel_index(np.argmax(err - tol), err.shape)
    first = np.argwhere(err > tol)
    assert first.size == 0, (f"delta_en mismatch: {len(first)} moves, first at {tuple(first[0])} (move index {tc['move_index'][tuple(first[0])]}): "
                             f"gpu {tg['delta_en'][tuple(first[0])]!r} cpu {tc['delta_en'][tuple(first[0])]!r}; worst at {worst}")
    sg, sc = gpu["states"], cpu["states"]
    assert np.array_equal(sg["counters"], sc["counters"]), "counters differ"
    assert np.array_equal(sg["max_displ"], sc["max_displ"]), "step sizes differ"
    assert np.array_equal(gpu["hist"], cpu["hist"]), "histograms differ"
    assert np.array_equal(gpu["hist_reduced"], cpu["hist_reduced"]), "ladder-reduced histograms differ"
    if positions_exact:
        assert np.array_equal(sg["pos"], sc["pos"]), "positions differ"
        assert np.array_equal(sg["box"], sc["box"])
    else:
        np.testing.assert_allclose(sg["pos"], sc["pos"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(sg["box"], sc["box"], rtol=1e-13)
    np.testing.assert_allclose(sg["en_tot"], sc["en_tot"], rtol=1e-11, atol=1024 * EPS * e_scale)
    np.testing.assert_allclose(sg["ham"], sc["ham"], rtol=1e-11, atol=1024 * EPS * e_scale)
    return float(err.max())
