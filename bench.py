#!/usr/bin/env python3
"""bench.py — MC moves/s of the parallel-tempering sweep hot path (BASELINE.json metric).

Headline workload (config.workload): BASELINE C2 — Ne55 extended-Lennard-Jones NVT, 24-temperature ladder
(scripts/ne55.jl), 4096 independent ladders batched per GPU, ladders sharded across GPUs (weak scaling,
no data-path collective; one NCCL all-reduce of the ladder-summed energy histogram at the end).
One "step" = `--cycles` calls of mc_cycle! (src/MCRun.jl:215) for every ladder of the rank =
n_ladders * 24 * 55 * cycles MC moves, including exchange, step-size adaptation and histogram sampling.

  value    device-resident throughput: states already in HBM, K steps timed with CUDA events on the
           launching stream, max over ranks.
  e2e      the same metric through the reference-facing calls with HOST buffers: every step uploads the
           start configurations from pinned host memory (MCState construction), runs the cycles and
           downloads positions / energies / step sizes / counters / ham and the histogram into pinned host
           memory.  The rank's ladders run as two half-batches (two handles, two streams, two host threads),
           so the copies of one half overlap the sweep of the other; nothing is skipped.
  roofline FP64-pipe bound (BASELINE.json: "% FP64 roofline"): achieved = moves/s * W, W = algorithmic flops
           per move, peak = DFMA-chain peak measured live by ptmc_fp64_peak.
  cpu_baseline / --impl reference: the oracle's C restatement of the reference's CPU path, OpenMP over
           all host cores, on a bounded sample of the same workload (Julia itself is not in the image).
  configs  the other BASELINE configs (C1 Ne13, C3 Cu55 EAM, C4 pbc216 NPT, C5 Cu55 RuNNer) measured in the same
           run with the same method (pilot-adapted step sizes, >= 3 warm-up and >= 3 timed steps, CUDA events,
           e2e, roofline, CPU sample); under --gpus N they shard over the ranks like C2.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MC moves/sec (all replicas)"
UNIT = "moves/s"
SEED = 1234


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ladders", type=int, default=4096, help="independent PT ladders per GPU (C2)")
    ap.add_argument("--cycles", type=int, default=200, help="mc_cycle! calls per step (C2)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per replica (0 = auto)")
    ap.add_argument("--cpu-ladders", type=int, default=0, help="ladders in the CPU sample of C2 (0 = 8 per core)")
    ap.add_argument("--cpu-cycles", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="c1,c3,c4,c5", help="other BASELINE configs measured next to the headline ('' = none)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- workloads --
def specs(args):
    """Every BASELINE config: deck, ladders per GPU (whole waves of CTAs on 148 SMs), cycles per step, pilot that adapts the step
    sizes (src/Exchange.jl:250-263), W = algorithmic flops per move (SURVEY §8d / BASELINE.md §3), CPU sample (ladders, cycles)."""
    import ptmc_b200 as P
    H = P.host
    return {
        "c2": dict(workload="Ne55 ELJ NVT (scripts/ne55.jl), 24-temperature ladder, %d independent PT ladders per GPU" % args.ladders,
                   deck=lambda c: H.deck_ne55(n_traj=24, cycles=c), ladders=args.ladders, cycles=args.cycles, eq=200,
                   pilot=dict(ladders=64, cycles=3000, n_adjust=100), W=36 * 54, W_note="36 (N-1): 8 distance + 14 + 14 energy flops per pair",
                   cpu=(args.cpu_ladders or 0, args.cpu_cycles), lanes=args.lanes),
        "c1": dict(workload="Ne13 ELJ NVT (examples/ne13.jl), 32-temperature ladder, 8880 independent PT ladders per GPU",
                   deck=lambda c: H.deck_ne13(n_traj=32, cycles=c), ladders=8880, cycles=400, eq=200,
                   pilot=dict(ladders=148, cycles=3000, n_adjust=100), W=36 * 12, W_note="36 (N-1)", cpu=(64, 400), lanes=0),
        "c3": dict(workload="Cu55 embedded-atom NVT (scripts/cu55.jl), 28-temperature ladder, 2048 independent PT ladders per GPU",
                   deck=lambda c: H.deck_cu55_eam(n_traj=28, cycles=c), ladders=2048, cycles=40, eq=100,
                   pilot=dict(ladders=148, cycles=1500, n_adjust=50), W=20 * 54 + 5 * 55,
                   W_note="20 flops x 54 pairs + 5 x 55 atoms (+ 324 specials: rsqrt / pow), SURVEY §8d", cpu=(32, 40), lanes=0,
                   pair_evals_per_move=2 * 54),
        "c4": dict(workload="pbc216 ELJ NPT, RhombicBC, volume moves (scripts/pbc216.jl), 16-temperature ladder, 592 independent PT ladders per GPU",
                   deck=lambda c: H.deck_pbc216(n_traj=16, cycles=c), ladders=592, cycles=100, eq=60,
                   pilot=dict(ladders=148, cycles=600, n_adjust=20), W=14125,
                   W_note="(216 x 12470 + 371520) / 217: minimum-image atom moves + one full-energy volume move per cycle", cpu=(16, 20), lanes=0),
        "c5": dict(workload="Cu55 RuNNer neural-network potential NVT (scripts/cu55.jl, boundary 0), 28-temperature ladder, 37 independent PT ladders per GPU",
                   deck=lambda c: H.deck_cu55_runner(n_traj=28, cycles=c), ladders=37, cycles=8, eq=4,
                   pilot=dict(ladders=37, cycles=24, n_adjust=2, start_from="c3"), W=3.6e6,
                   W_note="symmetry functions 3.35 Mflop + forward pass 0.24 Mflop as the reference evaluates them", cpu=(1, 1), lanes=0),
        # not a BASELINE config: --configs brass
        "brass": dict(workload="Cu50Zn5 two-species RuNNer potential, NNVT with 2 atom swaps per cycle (scripts/brass55.jl), 10-temperature ladder, 74 independent PT ladders per GPU",
                      deck=lambda c: H.deck_brass55(n_traj=10, cycles=c), ladders=74, cycles=8, eq=4,
                      pilot=dict(ladders=74, cycles=24, n_adjust=2), W=(55 * 3.6e6 + 2 * 7.2e6) / 57,
                      W_note="atom moves as C5, a swap as two atom moves", cpu=(2, 1), lanes=0),
    }


def workload_config(spec):
    mcp = spec["deck"](spec["cycles"])[0]
    return {"workload": spec["workload"], "n_atoms": mcp.n_atoms, "n_traj": mcp.n_traj, "ladders_per_gpu": spec["ladders"],
            "cycles_per_step": spec["cycles"],
            "rng": "PTMC_RNG_BATCH8: 8 Xoshiro256++ substreams per slot (Julia's generator), each seeded from SHA-256(seed, stream id)",
            "l2": "flushed between timed steps (256 MiB write)"}


# --------------------------------------------------------------------------------------- CPU baseline --
def cpu_sample(spec, ladders, cycles, threads, repeats=1, warm=True):
    """The oracle (C restatement of the reference's CPU path, reference operation order) on `threads` cores."""
    import ptmc_b200 as P
    from ptmc_b200.host import initial_arrays, sweep_config
    from oracle import oracle
    mcp, temp, start, pot, ens = spec["deck"](cycles)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=ladders, seed=SEED)
    eng = oracle.OracleEngine(cfg, threads=threads)
    pos, box = initial_arrays(mcp, start, ladders)
    eng.upload_states(pos, box)
    eng.seed_job(SEED)
    eng.set_hist_bounds(*hist_window(P, start, None))
    if warm:
        eng.run_cycles(1, max(1, cycles // 8), True)
    times = []
    for r in range(repeats):
        t0 = time.perf_counter()
        eng.run_cycles(1 + r * cycles, cycles, True)
        times.append(time.perf_counter() - t0)
    eng.close()
    return ladders * mcp.n_traj * cfg.n_steps * cycles, times


def hist_window(P, start, en):
    """energy window from the equilibrated energies (as the host does, src/MCRun.jl:293-310); NPT: 0.8 V0 .. 2 V0 (src/Sampling.jl:203-211)"""
    s0 = start[0] if isinstance(start, (list, tuple)) else start
    lo, hi = (-1.0, 1.0) if en is None else (float(en.min()), float(en.max()))
    pad = 0.05 * (abs(lo) + abs(hi)) + 1e-6
    if isinstance(s0.boundary_condition, P.SphericalBC):
        return (lo - pad, hi + pad, None, None)
    v0 = P.get_volume(s0.boundary_condition)
    return (lo - pad, hi + pad, 0.8 * v0, 2.0 * v0)


def cpu_record(spec, cores, repeats=3):
    ladders, cycles = spec["cpu"]
    ladders = ladders or 8 * cores
    mv, times = cpu_sample(spec, ladders, cycles, cores, repeats=repeats, warm=repeats > 1)
    best = min(times)
    n_traj = spec["deck"](1)[0].n_traj
    return {"value": mv / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d ladders x %d temperatures x %d cycles (%d moves), best of %d, OpenMP over ladders; oracle/ptmc_oracle.c restates "
                      "the reference's Julia CPU path (Julia is not in the image)" % (ladders, n_traj, cycles, mv, repeats)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; Julia is absent from the
    image) on all host cores; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    S = specs(args)
    spec = S["c2"]
    ladders = args.cpu_ladders or 8 * cores
    moves, times = cpu_sample(spec, ladders, args.cpu_cycles, cores, repeats=args.warmup + args.steps, warm=False)
    timed = times[args.warmup:]
    total = sum(timed)
    value = moves * len(timed) / total
    sample = "%d ladders x 24 temperatures x %d cycles per step (%d moves/step), OpenMP over ladders" % (ladders, args.cpu_cycles, moves)
    others = {}
    for name in [c for c in args.configs.split(",") if c]:
        rec = cpu_record(S[name], cores, repeats=1)
        others[name] = {"config": workload_config(S[name]), "value": rec["value"], "unit": UNIT, "cpu_baseline": rec}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(spec),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "configs": others}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------- clock log --
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=10)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w": statistics.median(pw) if pw else None, "samples": len(self.rows)}


class DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (for the NCCL all-reduce)."""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


# ------------------------------------------------------------------------------------------- B200 arm --
class Ctx:
    pass


def pilot_step_sizes(ctx, spec, start_md=None):
    """Step sizes the timed region works at: a small pilot batch runs enough adaptation periods (update_max_stepsize!,
    src/Exchange.jl:250-263) that every temperature sits in the 40-60 % acceptance window the reference steers to, instead of
    the constructor's max_displ = 0.1 (src/MCStates.jl:104).  The median over the pilot's ladders is used for every ladder."""
    from ptmc_b200.engine import Engine
    from ptmc_b200.host import initial_arrays, sweep_config
    p = spec["pilot"]
    mcp, temp, start, pot, ens = spec["deck"](p["cycles"])
    mcp.n_adjust = p["n_adjust"]
    L = p["ladders"]
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=SEED + 1, device=ctx.local, lanes_per_replica=spec["lanes"])   # its own job seed: no stream shared with the run
    eng = Engine(cfg)
    pos, box = initial_arrays(mcp, start, L)
    md = None if start_md is None else np.ascontiguousarray(np.broadcast_to(start_md, (L, mcp.n_traj, 4)))
    eng.upload_states(pos, box, md)
    eng.run_cycles(1, p["cycles"], False)
    md_slot = np.median(eng.download_states(("max_displ",))["max_displ"], axis=0)      # [n_traj][4]
    eng.close()
    return md_slot


def measure(ctx, name, spec, steps, warmup, md_slot, with_cpu):
    """One BASELINE config on this rank's GPU: device-resident value (CUDA events), e2e with host buffers, roofline, CPU sample."""
    import torch
    import ptmc_b200 as P
    from ptmc_b200 import abi, multi
    from ptmc_b200.engine import Engine
    from ptmc_b200.host import initial_arrays, sweep_config
    dist, dev, rank, world, local = ctx.dist, ctx.dev, ctx.rank, ctx.world, ctx.local
    L, cycles = spec["ladders"], spec["cycles"]
    mcp, temp, start, pot, ens = spec["deck"](cycles)
    T, N = mcp.n_traj, mcp.n_atoms
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=SEED, device=local, lanes_per_replica=spec["lanes"])
    n_steps = cfg.n_steps
    eng = Engine(cfg)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    pos, box = initial_arrays(mcp, start, L)
    max_displ = np.ascontiguousarray(np.broadcast_to(md_slot, (L, T, 4)))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pin_pos, pin_box, pin_md = pin(pos), pin(box), pin(max_displ)
    eng.upload_states(pin_pos.numpy(), pin_box.numpy(), pin_md.numpy())
    # ladder l of rank r is global ladder r*L + l of the job: streams are hashed from (seed, global stream id) (ptmc_seed_job),
    # so per-ladder results do not depend on the number of GPUs and jobs with different seeds share no stream
    eng.seed_job(SEED, rank * L)
    eng.run_cycles(1, spec["eq"], False)                       # equilibrate so that the histogram window is meaningful
    en = eng.download_states(("en_tot",))["en_tot"]
    bounds = hist_window(P, start, en)
    eng.reset_counters()
    eng.set_hist_bounds(*bounds)
    moves_per_step = L * T * n_steps * cycles

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- device-resident timing (value)
    idx = 1
    for _ in range(warmup):
        with torch.cuda.stream(stream):
            ctx.flush.fill_(idx & 0xFF)
        eng.run_cycles(idx, cycles, True); idx += cycles
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ev = []
    with ClockSampler(local) as clk:
        barrier()                                                # all ranks enter the timed region together (after the sampler thread started)
        e0.record(stream)
        for _ in range(steps):
            with torch.cuda.stream(stream):
                ctx.flush.fill_(idx & 0xFF)                      # L2 flush between timed steps
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            eng.run_cycles(idx, cycles, True)
            k1.record(stream)
            idx += cycles
            kernel_ev.append((k0, k1))
        # the one collective of the path: ladder-summed histogram, all-reduced across GPUs
        ptr, n = eng.reduced_hist_dev()
        hist_t = torch.as_tensor(DevArray(ptr, n), device=dev)
        if dist is not None:
            with torch.cuda.stream(stream):
                dist.all_reduce(hist_t)
        e1.record(stream)
        barrier()
    launches = eng.launch_count() - launches0
    eng.sync()
    st_end = eng.download_states(("counters",))["counters"]
    acc_timed = float(st_end[:, :, abi.C_ATOM_TOT].sum()) / (L * T * cfg.atom_moves * cycles * (steps + warmup))
    ms = e0.elapsed_time(e1)
    kms = [a.elapsed_time(b) for a, b in kernel_ev]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * moves_per_step * steps / (ms_max * 1e-3)
    hist_total = int(hist_t.sum().item())
    expect = world * L * T * cycles * (steps + warmup) // mcp.mc_sample
    kernel_ms_avg = sum(kms) / len(kms)
    lanes = eng.lanes_per_replica()
    eng.close()

    # ------------------------------------------------------------------- end-to-end timing (e2e)
    # The reference-facing sequence (upload = MCState construction, set_hist_bounds, run_cycles, download everything the driver
    # reads: positions, box, energies, step sizes, counters, ham, histogram) on HOST buffers, every step.  Two half-batches with
    # their own handle, stream and host thread: the copies of one half overlap the sweep of the other.
    e2e_steps = max(3, min(steps, 5))
    halves = [L - L // 2, L // 2] if L >= 2 else [L]
    jobs = []
    off = 0
    for Lh in halves:
        hcfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=Lh, seed=SEED, device=local, lanes_per_replica=spec["lanes"])
        he = Engine(hcfg)
        hs = torch.cuda.Stream(device=dev)
        he.set_stream(hs.cuda_stream)
        he.seed_job(SEED, rank * L + off)
        inp = (pin(pos[off:off + Lh]), pin(box[off:off + Lh]), pin(max_displ[off:off + Lh]))
        outp = {k: torch.empty(shp, dtype=torch.float64 if dt == np.float64 else torch.int64).pin_memory().numpy()
                for k, (shp, dt) in he.state_shapes().items()}
        jobs.append((he, inp, outp))
        off += Lh
    h2d = sum(sum(x.numel() * 8 for x in inp) for _, inp, _ in jobs)
    d2h_box = [0] * len(jobs)

    errors = []

    def e2e_loop(k, n):
        try:
            _e2e_loop(k, n)
        except BaseException as exc:       # a failed half must fail the run, not shorten its clock
            errors.append(exc)

    def _e2e_loop(k, n):
        he, inp, outp = jobs[k]
        for s in range(n):
            he.upload_states(inp[0].numpy(), inp[1].numpy(), inp[2].numpy())
            he.set_hist_bounds(*bounds)
            he.run_cycles(1 + s * cycles, cycles, True)
            st = he.download_states(out=outp)
            hs_ = he.download_hist(reduce_ladders=True)
            d2h_box[k] = sum(v.nbytes for v in st.values()) + hs_.nbytes

    def run_threads(n):
        th = [threading.Thread(target=e2e_loop, args=(k, n)) for k in range(len(jobs))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize(dev)
        if errors:
            raise errors[0]

    run_threads(1)                                               # warm-up (pins, first launches)
    barrier()
    t0 = time.perf_counter()
    run_threads(e2e_steps)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * moves_per_step * e2e_steps / float(t.item())
    for he, _, _ in jobs:
        he.close()

    achieved = (moves_per_step / (kernel_ms_avg * 1e-3)) * spec["W"] / 1e12     # this rank's sweep kernel
    rec = None
    if rank == 0:
        cpu = None
        if with_cpu and world == 1:
            cpu = cpu_record(spec, os.cpu_count() or 1, repeats=3 if name == "c2" else 1)
        roof = {"bound": "fp64", "achieved": achieved, "peak": ctx.peak, "unit": "TFLOP/s", "frac": (achieved / ctx.peak) if ctx.peak else None,
                "traffic": None,
                "note": "FP64 FMA pipe bound (not hbm/tensor): W = %g algorithmic flops per move (%s); peak = DFMA-chain peak measured live "
                        "(MEASURED_PEAKS.json has no FP64 figure); traffic: not read in this run - state load / store per launch, ncu "
                        "dram bytes in profiles/r01_sweep_ne55_g8.md (C2), r02_*.md" % (spec["W"], spec["W_note"])}
        if "pair_evals_per_move" in spec:
            roof["pair_evals_per_s"] = moves_per_step / (kernel_ms_avg * 1e-3) * spec["pair_evals_per_move"]
        rec = {"value": value, "unit": UNIT, "ms_per_step": ms_max / steps, "steps": steps, "warmup": warmup,
               "config": workload_config(spec), "clocks": clk.summary(),
               "details": {"lanes_per_replica": lanes, "moves_per_step_per_gpu": moves_per_step, "hist_samples_total": hist_total,
                           "hist_samples_expected": expect, "sweep_kernel_ms_avg": kernel_ms_avg, "e2e_steps": e2e_steps,
                           "e2e_mode": "two half-batches (two handles / streams / host threads): copies of one overlap the sweep of the other",
                           "atom_move_acceptance": acc_timed,
                           "max_displ_range": [float(md_slot[:, 0].min()), float(md_slot[:, 0].max())],
                           "pilot": {k: v for k, v in spec["pilot"].items()},
                           "algorithmic_hbm_bytes_per_move": 0, "state_bytes_per_launch_per_gpu": int(pos.nbytes * 2)},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(sum(d2h_box))},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
    return rec


def run_b200(args):
    import torch
    from ptmc_b200 import abi

    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sweep has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(ctx.local)
    ctx.dist = None
    if ctx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist
    ctx.dev = torch.device("cuda", ctx.local)
    ctx.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=ctx.dev)
    import ctypes as C
    v = C.c_double()
    ctx.peak = v.value if abi.load_library().ptmc_fp64_peak(ctx.local, 5, C.byref(v)) == 0 else None

    S = specs(args)
    with_cpu = not args.no_cpu_baseline
    md = {"c2": pilot_step_sizes(ctx, S["c2"])}
    head = measure(ctx, "c2", S["c2"], args.steps, args.warmup, md["c2"], with_cpu)
    others = {}
    names = [c for c in args.configs.split(",") if c]
    for name in names:
        spec = S[name]
        start_md = None
        if spec["pilot"].get("start_from"):
            src = spec["pilot"]["start_from"]
            if src not in md:
                md[src] = pilot_step_sizes(ctx, S[src])
            start_md = md[src]
        md[name] = pilot_step_sizes(ctx, spec, start_md)
        rec = measure(ctx, name, spec, max(3, min(args.steps, 4)), max(3, args.warmup), md[name], with_cpu)
        if rec is not None:
            others[name] = rec
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": head["config"], "clocks": head["clocks"], "details": head["details"],
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
                "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "configs": others}
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
