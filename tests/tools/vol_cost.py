#!/usr/bin/env python3
"""Cost of one pbc216 volume move in units of atom moves: time per cycle with 1 / 3 / 5 volume moves per cycle (C4 deck, 592 ladders).
    python tests/tools/vol_cost.py [--ladders 592] [--cycles 40]"""
import argparse
import dataclasses
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

ap = argparse.ArgumentParser()
ap.add_argument("--ladders", type=int, default=592)
ap.add_argument("--cycles", type=int, default=40)
ap.add_argument("--max-displ", type=float, default=0.35)
a = ap.parse_args()
mcp, temp, start, pot, ens = P.host.deck_pbc216(n_traj=16, cycles=100)
out = {}
for nv in (1, 3, 5):
    ens_v = dataclasses.replace(ens, n_volume_moves=nv)
    eng = Engine(sweep_config(mcp, temp, start, pot, ens_v, n_ladders=a.ladders, seed=7))
    pos, box = initial_arrays(mcp, start, a.ladders)
    md = np.full((a.ladders, mcp.n_traj, 4), 0.1)
    md[:, :, 0] = a.max_displ
    eng.upload_states(pos, box, md)
    v0 = P.get_volume(start[0].boundary_condition)
    eng.set_hist_bounds(0.0, 1.0, 0.8 * v0, 2.0 * v0)
    eng.run_cycles(1, 40, False)
    eng.sync()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    eng.run_cycles(41, a.cycles, True)
    eng.sync()
    t1.record(); torch.cuda.synchronize()
    out[nv] = t0.elapsed_time(t1) / a.cycles
    eng.close()
va = (out[3] - out[1]) / 2.0
vb = (out[5] - out[3]) / 2.0
atom = (out[1] - va) / 216.0
print(json.dumps({"ms_per_cycle": out, "ms_per_volume_move": [va, vb], "ms_per_atom_move": atom, "volume_in_atom_moves": [va / atom, vb / atom]}))
