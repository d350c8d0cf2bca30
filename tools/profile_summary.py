#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) of the sweep kernel into markdown for profiles/.

usage: tools/profile_summary.py gpurun_out/prof.ncu-rep [warp_moves_in_launch] > profiles/<name>.md
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source sass` (works
without a GPU)."""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (CTAs/SM)"),
    ("launch__occupancy_limit_barriers", "occupancy limit: hardware barriers (CTAs/SM)"),
    ("launch__occupancy_limit_warps", "occupancy limit: warps (CTAs/SM)"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe instructions (% of peak)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (fp32/int) pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (MUFU, conversions)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU pipe"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcgen05 tensor pipe"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "FP64 tensor sub-pipe (DMMA)"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active (all sub-pipes)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
]


def ncu(rep, *args):
    r = subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def main():
    rep = sys.argv[1]
    warp_moves = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = ncu(rep, "--page", "raw", "--csv")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu summary of `{rep}`\n")
    print(f"kernel: `{d.get('Kernel Name', ('?', ''))[0]}`\n")
    print("| metric | value |\n|---|---|")
    for k, name in KEYS:
        if k in d:
            print(f"| {name} (`{k}`) | {d[k][0]} {d[k][1]} |")
    print("\n## warp stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|")
    for k in sorted(d):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            print(f"| {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {float(d[k][0]):.3f} |")
    src = ncu(rep, "--page", "source", "--csv", "--print-source", "sass")
    h = src[1]
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    byop, samp = collections.Counter(), collections.Counter()
    for r in src[2:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += int(r[ia])
        samp[op] += int(r[isamp]) if r[isamp].isdigit() else 0
    tot = sum(byop.values())
    print(f"\n## dynamic SASS instruction mix ({tot} warp instructions)\n")
    unit = " per warp-move" if warp_moves else ""
    print(f"| opcode | warp instructions{unit} | share | stall samples |\n|---|---|---|---|")
    for op, c in byop.most_common(24):
        v = f"{c / warp_moves:.1f}" if warp_moves else str(c)
        print(f"| {op} | {v} | {100 * c / tot:.1f}% | {samp[op]} |")
    fp64 = sum(byop[o] for o in ("DFMA", "DADD", "DMUL", "DSETP"))
    print(f"\nFP64 instructions: {100 * fp64 / tot:.1f}% of all warp instructions" + (f" ({fp64 / warp_moves:.1f} per warp-move)" if warp_moves else ""))


if __name__ == "__main__":
    main()
