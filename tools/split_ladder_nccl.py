#!/usr/bin/env python3
"""One ladder across GPUs over NCCL: every rank owns n_traj / world temperatures of each ladder, the exchange
step all-gathers the energies and swaps configurations point to point.  Rank 0 also runs the same job on ONE handle and
checks that the split run reproduces it bit for bit.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/split_ladder_nccl.py [--cycles 200]"""
import argparse
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ptmc_b200 as P  # noqa: E402
from ptmc_b200 import abi, multi  # noqa: E402
from ptmc_b200.engine import Engine  # noqa: E402
from ptmc_b200.host import initial_arrays, sweep_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=200)
ap.add_argument("--ladders", type=int, default=8)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=24, cycles=a.cycles)
bounds, seed, eq = (-0.1, 0.0), 4242, 20
t0 = time.perf_counter()
st, hist, (s0, cnt) = multi.run_split_ladder(mcp, temp, start, pot, ens, n_ladders=a.ladders, seed=seed, dist=dist, device=local,
                                             bounds=bounds, eq_cycles=eq, comm_device=torch.device("cuda", local))
wall = time.perf_counter() - t0
gathered = [None] * world
dist.gather_object({k: st[k] for k in ("pos", "counters", "en_tot", "max_displ")} | {"hist": hist}, gathered if rank == 0 else None, dst=0)
if rank == 0:
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=a.ladders, seed=seed, device=local)
    eng = Engine(cfg)
    pos, box = initial_arrays(mcp, start, a.ladders)
    eng.upload_states(pos, box)
    multi.seed_global(eng, seed, 0, a.ladders)
    eng.run_cycles(1, eq, False); eng.reset_counters(); eng.set_hist_bounds(*bounds); eng.run_cycles(1, a.cycles, True)
    ref, ref_hist = eng.download_states(), eng.download_hist()
    eng.close()
    cat = lambda k: np.concatenate([g[k] for g in gathered], axis=1)
    same = {k: bool(np.array_equal(cat(k), ref[k])) for k in ("pos", "counters", "en_tot", "max_displ")}
    same["hist"] = bool(np.array_equal(cat("hist"), ref_hist))
    acc = ref["counters"][:, :, abi.C_EXC_ACC]
    b = 24 // world
    print(json.dumps({"world": world, "ladders": a.ladders, "n_traj": 24, "cycles": a.cycles, "identical_to_one_handle": same,
                      "exchanges_accepted": int(acc.sum() // 2), "accepted_across_rank_boundaries": int(sum(acc[:, k * b - 1].sum() for k in range(1, world))),
                      "split_wall_s": round(wall, 2), "backend": "nccl"}))
    assert all(same.values()), same
dist.destroy_process_group()
