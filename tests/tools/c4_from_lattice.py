#!/usr/bin/env python3
"""pbc216 (BASELINE C4) from the perfect fcc / hcp lattice, the way scripts/pbc216.jl runs it (mc_cycles = 10, :521-529): rate of the
first 10 cycles (every atom still 'never moved', cut-off ties alive) next to the rate once the configurations have thawed."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ptmc_b200 as P  # noqa: E402
from ptmc_b200.engine import Engine  # noqa: E402
from ptmc_b200.host import initial_arrays, sweep_config  # noqa: E402

L = 592
mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=16, cycles=10)
cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=3)
eng = Engine(cfg)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
pos, box = initial_arrays(mcp, start, L)
md = np.full((L, 16, 4), 0.1)
md[:, :, 0] = 0.5
v0 = P.get_volume(start[0].boundary_condition)
out = {}
for name, warm in (("first_10_cycles_from_the_lattice", 0), ("cycles_101_to_110", 100)):
    eng.upload_states(pos, box, md)
    eng.seed(3)
    eng.set_hist_bounds(0.0, 1.0, 0.8 * v0, 2.0 * v0)
    if warm:
        eng.run_cycles(1, warm, False)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.run_cycles(1 + warm, 10, True)
    e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1)
    out[name] = {"ms": ms, "moves_per_s": L * 16 * 217 * 10 / (ms * 1e-3)}
out["ratio"] = out["cycles_101_to_110"]["moves_per_s"] / out["first_10_cycles_from_the_lattice"]["moves_per_s"]
eng.close()
print(json.dumps(out))
