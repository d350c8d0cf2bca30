#!/usr/bin/env python3
"""bench.py — MC moves/s of the parallel-tempering sweep hot path (BASELINE.json metric).

Workload (config.workload): BASELINE C2 — Ne55 extended-Lennard-Jones NVT, 24-temperature ladder
(scripts/ne55.jl), 4096 independent ladders batched per GPU, ladders sharded across GPUs (weak scaling,
no data-path collective; one NCCL all-reduce of the ladder-summed energy histogram at the end).
One "step" = `--cycles` calls of mc_cycle! (src/MCRun.jl:215) for every ladder of the rank =
n_ladders * 24 * 55 * cycles MC moves, including exchange, step-size adaptation and histogram sampling.

  value    device-resident throughput: states already in HBM, K steps timed with CUDA events on the
           launching stream, max over ranks.
  e2e      the same metric through the reference-facing call with HOST buffers: every step uploads the
           start configurations from pinned host memory (MCState construction), runs the cycles and
           downloads positions / energies / step sizes / counters / ham and the histogram.
  roofline FP64-pipe bound (BASELINE.json: "% FP64 roofline"): achieved = moves/s * W, W = 36*(N-1) = 1944
           algorithmic flops per move, peak = DFMA-chain peak measured live by ptmc_fp64_peak.
  cpu_baseline / --impl reference: the oracle's C restatement of the reference's CPU path, OpenMP over
           all host cores, on a bounded sample of the same workload (Julia itself is not in the image).
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

N_TRAJ, N_ATOMS = 24, 55
W_FLOPS_PER_MOVE = 36 * (N_ATOMS - 1)          # (8 + 14 + 14) flops per pair, N-1 pairs
METRIC = "MC moves/sec (all replicas)"
UNIT = "moves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ladders", type=int, default=4096, help="independent PT ladders per GPU")
    ap.add_argument("--cycles", type=int, default=200, help="mc_cycle! calls per step")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per replica (0 = auto)")
    ap.add_argument("--cpu-ladders", type=int, default=0, help="ladders in the CPU sample (0 = 8 per core)")
    ap.add_argument("--cpu-cycles", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(ladders, cycles):
    return {"workload": "Ne55 ELJ NVT (scripts/ne55.jl), 24-temperature ladder, %d independent PT ladders per GPU" % ladders,
            "n_atoms": N_ATOMS, "n_traj": N_TRAJ, "ladders_per_gpu": ladders, "cycles_per_step": cycles,
            "moves_per_step_per_gpu": ladders * N_TRAJ * N_ATOMS * cycles,
            "rng": "PTMC_RNG_BATCH8: 8 Xoshiro256++ substreams per slot (Julia's generator), seeded Random.seed!(seed + k)",
            "l2": "flushed between timed steps (256 MiB write); resident state 130 MB > 126 MB L2"}


# --------------------------------------------------------------------------------------- CPU baseline --
def cpu_sample(ladders, cycles, threads, repeats=1, warm=True):
    """The oracle (C restatement of the reference's CPU path, reference operation order) on `threads` cores."""
    import ptmc_b200 as P
    from ptmc_b200.host import initial_arrays, sweep_config
    from oracle import oracle
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=N_TRAJ, cycles=cycles)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=ladders, seed=1234)
    eng = oracle.OracleEngine(cfg, threads=threads)
    pos, box = initial_arrays(mcp, start, ladders)
    eng.upload_states(pos, box)
    eng.seed(1234)
    eng.set_hist_bounds(-0.1, 0.0)
    if warm:
        eng.run_cycles(1, max(1, cycles // 8), True)
    times = []
    for r in range(repeats):
        t0 = time.perf_counter()
        eng.run_cycles(1 + r * cycles, cycles, True)
        times.append(time.perf_counter() - t0)
    eng.close()
    moves = ladders * N_TRAJ * N_ATOMS * cycles
    return moves, times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; Julia is absent from the
    image) on all host cores; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ladders = args.cpu_ladders or 8 * cores
    moves, times = cpu_sample(ladders, args.cpu_cycles, cores, repeats=args.warmup + args.steps, warm=False)
    timed = times[args.warmup:]
    total = sum(timed)
    value = moves * len(timed) / total
    sample = "%d ladders x %d temperatures x %d cycles per step (%d moves/step), OpenMP over ladders" % (ladders, N_TRAJ, args.cpu_cycles, moves)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.ladders, args.cycles), cpu_sample=sample),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
def run_b200(args):
    import torch
    import ptmc_b200 as P
    from ptmc_b200 import abi
    from ptmc_b200.engine import Engine
    from ptmc_b200.host import initial_arrays, sweep_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sweep has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    L, cycles = args.ladders, args.cycles
    mcp, temp, start, pot, ens = P.host.deck_ne55(n_traj=N_TRAJ, cycles=cycles)
    # ladder l of rank r is global ladder r*L + l of a job of world*L ladders: streams are indexed globally
    # (ptmc_b200.multi), so per-ladder results do not depend on the number of GPUs
    from ptmc_b200 import multi
    seed = 1234
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=L, seed=seed, device=local, lanes_per_replica=args.lanes)
    eng = Engine(cfg)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    pos, box = initial_arrays(mcp, start, L)
    pin_pos = torch.from_numpy(pos).pin_memory()
    pin_box = torch.from_numpy(box).pin_memory()
    # step sizes: a pilot of 64 ladders runs 3000 cycles (30 adaptations, src/Exchange.jl:250-263) so that the timed
    # region works at the adapted 40-60 % acceptance instead of the constructor's max_displ = 0.1 (~87 %)
    pcfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=64, seed=seed + 7, device=local, lanes_per_replica=args.lanes)
    pilot = Engine(pcfg)
    ppos, pbox = initial_arrays(mcp, start, 64)
    pilot.upload_states(ppos, pbox)
    pilot.run_cycles(1, 3000, False)
    md_slot = np.median(pilot.download_states(("max_displ",))["max_displ"], axis=0)      # [n_traj][4]
    pilot.close()
    max_displ = np.ascontiguousarray(np.broadcast_to(md_slot, (L, N_TRAJ, 4)))
    pin_md = torch.from_numpy(max_displ).pin_memory()
    eng.upload_states(pin_pos.numpy(), pin_box.numpy(), pin_md.numpy())
    if world > 1:
        multi.seed_global(eng, seed, rank * L, world * L)
    else:
        eng.seed(seed)
    # equilibrate so that the histogram window is meaningful
    eng.run_cycles(1, 200, False)
    en = eng.download_states(("en_tot",))["en_tot"]
    eng.reset_counters()
    eng.set_hist_bounds(float(en.min()) - 1e-3, float(en.max()) + 2e-3)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    moves_per_step = L * N_TRAJ * N_ATOMS * cycles

    def step(idx):
        with torch.cuda.stream(stream):
            flush.fill_(idx & 0xFF)
        eng.run_cycles(idx, cycles, True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- device-resident timing (value)
    idx = 1
    for _ in range(args.warmup):
        step(idx); idx += cycles
    barrier()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(idx & 0xFF)
            k0.record(stream)
            eng.run_cycles(idx, cycles, True)
            k1.record(stream)
            idx += cycles
            kernel_ms.append((k0, k1))
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    acc_timed = float(eng.download_states(("counters",))["counters"][:, :, abi.C_ATOM_TOT].sum()) / (L * N_TRAJ * N_ATOMS * cycles * (args.steps + args.warmup))
    ms = e0.elapsed_time(e1)
    kms = [a.elapsed_time(b) for a, b in kernel_ms]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * moves_per_step * args.steps / (ms_max * 1e-3)
    hist_total = int(hist_t.sum().item())
    expect = world * L * N_TRAJ * cycles * (args.steps + args.warmup) // mcp.mc_sample
    kernel_ms_avg = sum(kms) / len(kms)

    # ------------------------------------------------------------------- end-to-end timing (e2e)
    # results land in pinned host buffers owned by the caller (what a Julia Array passed to ccall would be after
    # cudaHostRegister): positions, box, energies, step sizes, counters, ham
    pinned = {k: torch.empty(shp, dtype=torch.float64 if dt == np.float64 else torch.int64).pin_memory().numpy()
              for k, (shp, dt) in eng.state_shapes().items()}
    h2d = pin_pos.numel() * 8 + pin_box.numel() * 8 + pin_md.numel() * 8
    d2h = 0
    e2e_steps = max(3, min(args.steps, 5))

    def e2e_step(i):
        eng.upload_states(pin_pos.numpy(), pin_box.numpy(), pin_md.numpy())
        eng.set_hist_bounds(float(en.min()) - 1e-3, float(en.max()) + 2e-3)
        eng.run_cycles(i, cycles, True)
        st = eng.download_states(out=pinned)
        hs = eng.download_hist(reduce_ladders=True)
        return sum(v.nbytes for v in st.values()) + hs.nbytes

    d2h = e2e_step(1)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(1 + s * cycles)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * moves_per_step * e2e_steps / float(t.item())

    # ------------------------------------------------------------------------------- roofline
    peak = eng_peak = None
    if rank == 0:
        import ctypes as C
        v = C.c_double()
        if abi.load_library().ptmc_fp64_peak(local, 5, C.byref(v)) == 0:
            eng_peak = v.value
    achieved = (moves_per_step / (kernel_ms_avg * 1e-3)) * W_FLOPS_PER_MOVE / 1e12   # this rank's sweep kernel
    peak = eng_peak
    lanes = eng.lanes_per_replica()
    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cl = args.cpu_ladders or 8 * cores
        mv, times = cpu_sample(cl, args.cpu_cycles, cores, repeats=3)
        best = min(times)
        cpu = {"value": mv / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d ladders x %d temperatures x %d cycles (%d moves), best of 3, OpenMP over ladders; "
                         "oracle/ptmc_oracle.c restates the reference's Julia CPU path (Julia is not in the image)" % (cl, N_TRAJ, args.cpu_cycles, mv)}

    if rank == 0:
        conf = workload_config(L, cycles)
        conf.update({"lanes_per_replica": lanes, "hist_samples_total": hist_total, "hist_samples_expected": expect,
                     "sweep_kernel_ms_avg": kernel_ms_avg, "e2e_steps": e2e_steps, "atom_move_acceptance": acc_timed,
                     "max_displ_range": [float(md_slot[:, 0].min()), float(md_slot[:, 0].max())],
                     "algorithmic_hbm_bytes_per_move": 0, "state_bytes_per_launch_per_gpu": int(pos.nbytes * 2)})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": conf, "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": (achieved / peak) if peak else None,
                             # dram__bytes_read + dram__bytes_write of one sweep launch from the ncu --set full capture
                             # profiles/r01_sweep_ne55_g8.md (state load / store; independent of the cycle count)
                             "traffic": 301594368,
                             "note": "FP64 FMA pipe bound (not hbm/tensor): W = %d algorithmic flops per move; peak = DFMA-chain "
                                     "peak measured live (MEASURED_PEAKS.json has no FP64 figure); HBM traffic is ~0 B/move" % W_FLOPS_PER_MOVE},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
