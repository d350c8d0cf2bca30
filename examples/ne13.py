#!/usr/bin/env python3
"""examples/ne13.jl of the reference, run through the B200 sweep: Ne13 with the extended Lennard-Jones potential, NVT,
a parallel-tempering ladder from 4 to 16 K; `--ladders` independent ladders (seeds) are batched on the GPU and their
histograms summed before the multihistogram analysis.

    python examples/ne13.py --cycles 100000 --ladders 64 --save 10000

Prints <E>(T), Cv(T) from the fluctuation formula (finalise_results, src/Sampling.jl:377-399) and the Cv curve of the
multiple-histogram analysis (src/multihist_NVT.jl); with --save the reference's checkpoint files are written to ./checkpoint
(src/ReadSave.jl) so that its own `postprocess()` can read the run."""
import argparse
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

import ptmc_b200 as P  # noqa: E402
from ptmc_b200 import multihist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=100000)
ap.add_argument("--n-traj", type=int, default=25)            # examples/ne13.jl:82
ap.add_argument("--ladders", type=int, default=32)
ap.add_argument("--seed", type=int, default=1234)
ap.add_argument("--save", type=int, default=0)
ap.add_argument("--rdf", action="store_true")
a = ap.parse_args()

mc_params, temp, start_config, pot, ensemble = P.host.deck_ne13(n_traj=a.n_traj, cycles=a.cycles)
states, results = P.ptmc_run(mc_params, temp, start_config, pot, ensemble, n_ladders=a.ladders, seed=a.seed, shared_bounds=True,
                             rdfsave=a.rdf, save=a.save or False)
hist = sum(r.en_histogram for r in results)
en_avg = np.mean([r.en_avg for r in results], axis=0)
heat_cap = np.mean([r.heat_cap for r in results], axis=0)
print("   T / K      <E> / Eh        Cv / kB (fluctuation)   acceptance   exchange")
for t, e, c, sa, se in zip(temp.t_grid, en_avg, heat_cap, results[0].count_stat_atom, results[0].count_stat_exc):
    print("%8.3f  %14.8f  %14.4f  %16.3f  %9.3f" % (t, e, c, sa, se))
tt, cv = multihist.multihistogram_NVT(temp.t_grid, temp.beta_grid, hist, results[0].en_min, results[0].en_max, mc_params.n_bin)
k = int(np.argmax(cv))
print("multihistogram: Cv peak %.3f kB at T = %.3f K (%d ladders x %d cycles)" % (cv[k] / multihist.K_MULTIHIST, tt[k], a.ladders, a.cycles))
