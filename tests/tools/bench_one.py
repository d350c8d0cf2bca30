#!/usr/bin/env python3
"""One BASELINE config through bench.py's own measurement (pilot-adapted step sizes, CUDA events, e2e), with overrides for kernel
experiments:  python tests/tools/bench_one.py c4 --lanes 16 --ladders 592 [--lib libptmc_x.so] [--no-e2e-cpu]
Prints one JSON record per run."""
import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--lanes", type=int, default=None)
    ap.add_argument("--ladders", type=int, default=None)
    ap.add_argument("--cycles", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    if a.lib:
        os.environ["PTMC_B200_LIB"] = a.lib
    import torch
    import bench
    from ptmc_b200 import abi
    import ctypes as C
    args = bench.parse.__wrapped__() if hasattr(bench.parse, "__wrapped__") else None
    sys.argv = [sys.argv[0]]
    args = bench.parse()
    S = bench.specs(args)
    spec = S[a.config]
    if a.lanes is not None:
        spec["lanes"] = a.lanes
    if a.ladders is not None:
        spec["ladders"] = a.ladders
    if a.cycles is not None:
        spec["cycles"] = a.cycles
    ctx = bench.Ctx()
    ctx.rank, ctx.world, ctx.local, ctx.dist = 0, 1, 0, None
    torch.cuda.set_device(0)
    ctx.dev = torch.device("cuda", 0)
    ctx.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=ctx.dev)
    v = C.c_double()
    ctx.peak = v.value if abi.load_library().ptmc_fp64_peak(0, 5, C.byref(v)) == 0 else None
    start_md = None
    if spec["pilot"].get("start_from"):
        start_md = bench.pilot_step_sizes(ctx, S[spec["pilot"]["start_from"]])
    md = bench.pilot_step_sizes(ctx, spec, start_md)
    rec = bench.measure(ctx, a.config, spec, a.steps, 3, md, a.cpu)
    out = {"config": a.config, "lanes": rec["details"]["lanes_per_replica"], "ladders": spec["ladders"], "value": rec["value"],
           "e2e": rec["e2e"]["value"], "frac": rec["roofline"]["frac"], "acceptance": rec["details"]["atom_move_acceptance"],
           "kernel_ms": rec["details"]["sweep_kernel_ms_avg"], "lib": a.lib or "libptmc_b200.so"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
