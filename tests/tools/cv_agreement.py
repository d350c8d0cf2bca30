#!/usr/bin/env python3
"""Third correctness criterion of the north star on several potentials: heat-capacity curves from independent runs of the CUDA
sweep (many ladders) and of the CPU oracle (fewer ladders, other seeds) agree within the statistical error.

Cv(T)/kB at the ladder temperatures from the fluctuation formula (finalise_results, src/Sampling.jl:377-399), pooled over
ladders; the error bar is the ladder-to-ladder standard error of each estimate.  One JSON line per deck.

    python tests/tools/cv_agreement.py > gpurun_out/cv_agreement.json"""
import json
import os
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import ptmc_b200 as P  # noqa: E402
from ptmc_b200.engine import Engine  # noqa: E402
from ptmc_b200.host import initial_arrays, sweep_config  # noqa: E402

LUT = ROOT / "paralleltemperingmontecarlo.jl_b200" / "data" / "_lut" / "LookupTable_Neon_B0.0_MP2.txt"
DECKS = {
    # name: (deck, gpu ladders, cpu ladders, equilibration cycles, sampled cycles)
    "ne13_eljeven": (lambda: P.host.deck_ne13(n_traj=16, cycles=1), 512, 32, 2000, 8000),
    "ne13_eljb": (lambda: P.host.deck_ne13_B(n_traj=16, cycles=1), 512, 32, 2000, 8000),
    "cu55_eam": (lambda: P.host.deck_cu55_eam(n_traj=28, cycles=1), 128, 16, 1000, 3000),
    "pbc216_npt": (lambda: P.host.deck_pbc216(n_traj=16, cycles=1), 148, 16, 400, 1200),       # Cp from H = E + pV (src/Sampling.jl:35-46)
    "ne13_lookup": (lambda: P.host.deck_ne13_lookup(str(LUT), n_traj=16, cycles=1), 256, 32, 2000, 8000),
    # two species with atom swaps (NNVT): the Cu4Zn2 cluster of the reference's "States NNVT" test, 8 temperatures
    "brass6_nnvt": (lambda: brass6_deck(), 256, 32, 1000, 4000),
}


def brass6_deck():
    pos = np.array([[2.36, 2.36, 0.0], [6.99, 2.33, 0.0], [2.33, 6.99, 0.0], [-2.36, 2.36, 0.0], [-6.99, 2.33, 0.0], [-2.33, 6.99, 0.0]])
    return (P.MCParams(1, 8, 6), P.TempGrid(500.0, 2500.0, 8), P.Config(pos, P.SphericalBC(radius=8.0)),
            P.host.brass_runner_potential(4, 2), P.host.NNVT((4, 2), n_atom_swaps=1))


def heat_capacity(engine_cls, deck, L, eq, cycles, seed, **kw):
    mcp, temp, start, pot, ens = deck
    eng = engine_cls(sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=seed), **kw)
    pos, box = initial_arrays(mcp, start, L)
    eng.upload_states(pos, box)
    eng.seed(seed)
    eng.run_cycles(1, eq, False)
    eng.reset_counters()
    eng.set_hist_bounds(-1.0, 1.0, 1.0, 2.0)
    eng.run_cycles(1, cycles, True)
    eng.sync()
    ham = eng.download_states(("ham",))["ham"]
    eng.close()
    e1, e2 = ham[..., 0] / cycles, ham[..., 1] / cycles
    cv = (e2 - e1 ** 2) * temp.beta_grid[None, :] ** 2            # per ladder, per temperature
    return temp.t_grid, cv.mean(axis=0), cv.std(axis=0, ddof=1) / np.sqrt(L)


def main():
    from oracle import oracle
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, (mk, lg, lc, eq, cyc) in DECKS.items():
        if only and name not in only:
            continue
        if name == "ne13_lookup" and not LUT.exists():
            continue
        deck = mk()
        t, cg, sg = heat_capacity(Engine, deck, lg, eq, cyc, 101)
        _, cc, sc = heat_capacity(oracle.OracleEngine, deck, lc, eq, cyc, 90210, threads=os.cpu_count() or 1)
        err = np.sqrt(sg ** 2 + sc ** 2)
        z = np.abs(cg - cc) / err
        print(json.dumps({"deck": name, "gpu_ladders": lg, "cpu_ladders": lc, "cycles": cyc, "T": [round(float(x), 3) for x in t],
                          "cv_gpu": [round(float(x), 3) for x in cg], "cv_cpu": [round(float(x), 3) for x in cc],
                          "sigma": [round(float(x), 3) for x in err], "max_abs_z": round(float(z.max()), 2),
                          "mean_abs_z": round(float(z.mean()), 2), "max_rel_diff_of_peak": round(float(np.abs(cg - cc).max() / cc.max()), 4)}), flush=True)


if __name__ == "__main__":
    main()
