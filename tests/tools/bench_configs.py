#!/usr/bin/env python3
"""Throughput of every BASELINE config (C1..C5) on one GPU, device-resident, CUDA-event timed.
Not the driver's bench (that is bench.py, C2 only); this feeds profiles/ and DESIGN.md.

    python tests/tools/bench_configs.py [--only c1,c3] > gpurun_out/configs.json"""
import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import ptmc_b200 as P  # noqa: E402
from ptmc_b200.engine import Engine  # noqa: E402
from ptmc_b200.host import initial_arrays, sweep_config  # noqa: E402

import os  # noqa: E402

LUT_FILE = os.environ.get("PTMC_LUT_FILE", str(ROOT / "paralleltemperingmontecarlo.jl_b200" / "data" / "_lut" / "LookupTable_Neon_B0.0_MP2.txt"))

def _pbc216_variant(pot_name):
    """pbc216 the way most scripts/pbc*.jl run it: NPT with separated xy / z volume moves (scripts/pbc216.jl:62-70 names the
    lookup table; the table file is the Neon one of scripts/lookup-tables).  The tangent-dependent potentials work in Angstrom."""
    mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=16, cycles=100)
    if pot_name == "elj":
        return mcp, temp, start, pot, P.NPT(216, ens.pressure, True)
    a2b = 1.8897261259077824
    bc = start[0].boundary_condition
    start = [P.Config(c.positions / a2b, P.RhombicBC(bc.box_length / a2b, bc.box_height / a2b)) for c in start]
    pot = (P.ELJPotentialB(P.host.ELJB_NE_A, P.host.ELJB_NE_B, P.host.ELJB_NE_C) if pot_name == "eljb"
           else P.LookupTablePotential(LUT_FILE))
    return mcp, temp, start, pot, P.NPT(216, ens.pressure * a2b ** 3, True)


CASES = {
    # name: (deck factory, ladders, warm-up cycles, timed cycles per step, steps, histogram window)
    "c1_ne13_elj_nvt_32T": (lambda: P.host.deck_ne13(n_traj=32, cycles=100), 8880, 200, 100, 5, (-0.01, 0.0, None, None)),
    "c1b_ne13_eljb_nvt_16T": (lambda: P.host.deck_ne13_B(n_traj=16, cycles=100), 8192, 200, 100, 5, (-0.02, 0.0, None, None)),
    # scripts/ne13-lookup.jl; needs the reference's scripts/lookup-tables/LookupTable_Neon_B0.0_MP2.txt (not redistributed here):
    # PTMC_LUT_FILE or paralleltemperingmontecarlo.jl_b200/data/_lut/ (git-ignored); the case is skipped without it
    "c1c_ne13_lookup_nvt_25T": (lambda: P.host.deck_ne13_lookup(LUT_FILE, n_traj=25, cycles=100), 8192, 200, 100, 5, (-0.02, 0.0, None, None)),
    "c2_ne55_elj_nvt_24T": (lambda: P.host.deck_ne55(n_traj=24, cycles=100), 4096, 200, 50, 5, (-0.1, 0.0, None, None)),
    "c3_cu55_eam_nvt_28T": (lambda: P.host.deck_cu55_eam(n_traj=28, cycles=100), 2048, 100, 20, 5, (-8.0, -5.0, None, None)),
    "c4_pbc216_elj_npt_16T": (lambda: P.host.deck_pbc216(n_traj=16, cycles=100), 592, 20, 40, 3, None),
    "c5_cu55_runner_nvt_28T": (lambda: P.host.deck_cu55_runner(n_traj=28, cycles=100), 37, 4, 4, 3, (-13.0, -11.0, None, None)),
    # (f-3) rows at C4's size: separated volume moves, tangent matrix as state (not BASELINE configs; run with --only c4)
    "c4b_pbc216_elj_npt_separated_16T": (lambda: _pbc216_variant("elj"), 592, 20, 40, 3, None),
    "c4c_pbc216_eljb_npt_separated_16T": (lambda: _pbc216_variant("eljb"), 592, 20, 40, 3, None),
    "c4d_pbc216_lookup_npt_separated_16T": (lambda: _pbc216_variant("lookup"), 592, 20, 40, 3, None),
}


def run(name):
    deck, L, warm, cyc, steps, bounds = CASES[name]
    mcp, temp, start, pot, ens = deck()
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=99)
    eng = Engine(cfg)
    pos, box = initial_arrays(mcp, start, L)
    eng.upload_states(pos, box)
    if bounds is None:
        v0 = P.get_volume(start[0].boundary_condition)
        bounds = (0.0, 1.0, 0.8 * v0, 2.0 * v0)
    eng.set_hist_bounds(*bounds)
    eng.run_cycles(1, warm, False)
    eng.reset_counters()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    idx = 1
    e0.record(stream)
    for _ in range(steps):
        eng.run_cycles(idx, cyc, True)
        idx += cyc
    e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1)
    moves = L * mcp.n_traj * cfg.n_steps * cyc * steps
    st = eng.download_states(("counters", "max_displ"))
    acc = st["counters"][:, :, 0].sum() / (L * mcp.n_traj * mcp.n_atoms * cyc * steps)
    out = {"config": name, "n_ladders": L, "n_traj": mcp.n_traj, "n_atoms": mcp.n_atoms, "moves": moves, "ms": ms,
           "moves_per_s": moves / (ms * 1e-3), "atom_acceptance": float(acc), "lanes_per_replica": eng.lanes_per_replica(),
           "launches": eng.launch_count()}
    eng.close()
    return out


CPU_SAMPLE = {   # ladders, cycles for the CPU-oracle sample of each config (a few seconds on 16 cores)
    "c1_ne13_elj_nvt_32T": (64, 400), "c1b_ne13_eljb_nvt_16T": (64, 200), "c1c_ne13_lookup_nvt_25T": (64, 200), "c2_ne55_elj_nvt_24T": (64, 100), "c3_cu55_eam_nvt_28T": (32, 40),
    "c4_pbc216_elj_npt_16T": (16, 20), "c5_cu55_runner_nvt_28T": (1, 1),
    "c4b_pbc216_elj_npt_separated_16T": (16, 20), "c4c_pbc216_eljb_npt_separated_16T": (16, 10), "c4d_pbc216_lookup_npt_separated_16T": (16, 10),
}


def run_cpu(name):
    """The oracle (C restatement of the reference's CPU path) with OpenMP on all host cores, bounded sample."""
    import os
    from oracle import oracle
    deck, _, _, _, _, bounds = CASES[name]
    L, cyc = CPU_SAMPLE[name]
    mcp, temp, start, pot, ens = deck()
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=99)
    cores = os.cpu_count() or 1
    eng = oracle.OracleEngine(cfg, threads=cores)
    pos, box = initial_arrays(mcp, start, L)
    eng.upload_states(pos, box)
    if bounds is None:
        v0 = P.get_volume(start[0].boundary_condition)
        bounds = (0.0, 1.0, 0.8 * v0, 2.0 * v0)
    eng.set_hist_bounds(*bounds)
    t0 = time.perf_counter()
    eng.run_cycles(1, cyc, True)
    dt = time.perf_counter() - t0
    eng.close()
    moves = L * mcp.n_traj * cfg.n_steps * cyc
    return {"cpu_moves_per_s": moves / dt, "cpu_cores": cores, "cpu_sample_moves": moves, "cpu_s": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a bounded sample")
    a = ap.parse_args()
    names = [n for n in CASES if (any(n.startswith(p) for p in a.only.split(",")) if a.only else n[:3] not in ("c4b", "c4c", "c4d"))]
    for n in names:
        if "lookup" in n and not os.path.exists(LUT_FILE):
            print(json.dumps({"config": n, "skipped": "lookup table file not present (PTMC_LUT_FILE)"}), flush=True)
            continue
        t0 = time.time()
        r = run(n)
        if a.cpu:
            r.update(run_cpu(n))
            r["gpu_over_cpu"] = r["moves_per_s"] / r["cpu_moves_per_s"]
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
