#!/usr/bin/env python3
"""One short run of a BASELINE config for ncu: warm-up launch(es), then ONE profiled launch of `--cycles` cycles.

    ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -o gpurun_out/prof_x \\
        python tests/tools/prof_case.py c3_cu55_eam_nvt_28T --cycles 10"""
import argparse
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "tools"))

import ptmc_b200 as P  # noqa: E402
from bench_configs import CASES  # noqa: E402
from ptmc_b200.engine import Engine  # noqa: E402
from ptmc_b200.host import initial_arrays, sweep_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case")
ap.add_argument("--cycles", type=int, default=10)
ap.add_argument("--warm", type=int, default=100)
ap.add_argument("--ladders", type=int, default=0)
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--max-displ", type=float, default=0.0, help="atom step size (0 = the constructor's 0.1)")
a = ap.parse_args()
deck, L, _, _, _, bounds = CASES[a.case]
L = a.ladders or L
mcp, temp, start, pot, ens = deck()
eng = Engine(sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=99, lanes_per_replica=a.lanes))
pos, box = initial_arrays(mcp, start, L)
md = None
if a.max_displ:
    import numpy as np
    md = np.full((L, mcp.n_traj, 4), 0.1)
    md[:, :, 0] = a.max_displ
eng.upload_states(pos, box, md)
if bounds is None:
    v0 = P.get_volume(start[0].boundary_condition)
    bounds = (0.0, 1.0, 0.8 * v0, 2.0 * v0)
eng.set_hist_bounds(*bounds)
eng.run_cycles(1, a.warm, False)        # launch 1 (sweep kernels: -s 1 skips it ... plus a second warm launch)
eng.run_cycles(1 + a.warm, a.cycles, True)
eng.run_cycles(1 + a.warm + a.cycles, a.cycles, True)   # the profiled one with -s 2 -c 1
eng.sync()
eng.close()
print("done")
