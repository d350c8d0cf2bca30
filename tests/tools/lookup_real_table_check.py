import sys, json, numpy as np
import pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.engine import Engine
from ptmc_b200.host import initial_arrays, sweep_config
from oracle import oracle
from parity import compare_traces, run_pair
f = str(ROOT / 'paralleltemperingmontecarlo.jl_b200' / 'data' / '_lut' / 'LookupTable_Neon_B0.0_MP2.txt')
deck = P.host.deck_ne13_lookup(f, n_traj=25, cycles=200)
mcp, temp, start, pot, ens = deck
gpu, cpu = run_pair(mcp, temp, start, pot, ens, n_ladders=2, n_cycles=100, seed=5)
compare_traces(gpu, cpu)
cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=4, seed=77, rng_mode=abi.RNG_BATCH8)
pos, box = initial_arrays(mcp, start, 4)
out = []
for cls in (Engine, oracle.OracleEngine):
    e = cls(cfg); e.upload_states(pos, box); e.seed(77); e.run_cycles(1, 50, False); e.reset_counters(); e.set_hist_bounds(-0.02, 0.0); e.run_cycles(1, 200, True); e.sync()
    out.append((e.download_states(), e.download_hist())); e.close()
(sg, hg), (sc, hc) = out
res = {"table": "LookupTable_Neon_B0.0_MP2.txt (scripts/ne13-lookup.jl)", "replay_moves": int(gpu["trace"]["accepted"].size),
       "replay_decisions_identical": bool(np.array_equal(gpu["trace"]["accepted"], cpu["trace"]["accepted"])),
       "replay_max_abs_dE_diff": float(np.abs(gpu["trace"]["delta_en"] - cpu["trace"]["delta_en"]).max()),
       "production_positions_identical": bool(np.array_equal(sg["pos"], sc["pos"])), "production_hist_identical": bool(np.array_equal(hg, hc)),
       "production_counters_identical": bool(np.array_equal(sg["counters"], sc["counters"])),
       "acceptance": float(gpu["trace"]["accepted"].mean()), "en_tot_max_rel_diff": float(np.abs(sg["en_tot"] / sc["en_tot"] - 1).max())}
print(json.dumps(res))
