#!/usr/bin/env python3
"""profiles/r02_launches_bench*.md from the ncu launch lists (gpurun_out/r02_launches_bench*.csv: `ncu --metrics gpu__time_duration.sum
--clock-control none -c 400 --csv --log-file ... python bench.py ...`): time and share per kernel, share of the sweep kernel in a timed step."""
import csv,collections,sys
def load(path):
    rows=list(csv.reader(open(path)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
    for r in rows[hdr+1:]:
        if len(r)<15: continue
        agg[r[4]][0]+=1; agg[r[4]][1]+=float(r[14]); seq.append((r[4],float(r[14])))
    return agg,seq
def table(agg):
    tot=sum(v[1] for v in agg.values())
    out=["| kernel | launches | time (ms) | share |","|---|---|---|---|"]
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.2f %% |"%(k.replace('|','\\|')[:150],n,t/1e6,100*t/tot))
    return out,tot
agg,seq=load('/root/repo/gpurun_out/r02_launches_bench.csv')
t,tot=table(agg)
sw=sum(v[1] for k,v in agg.items() if 'sweep_kernel' in k or 'runner_moves' in k)
open('/root/repo/profiles/r02_launches_bench.md','w').write("\n".join(["# Launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (all five BASELINE configs)\n",
"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (profiles/r02_launches_bench.csv; first 400 launches; kernels serialised and cold-cache under the profiler: shares, not absolute times).\n"]+t+
["\nThe sweep kernels (`sweep_kernel*`, `runner_moves_kernel`) are %.1f %% of the GPU time of the command; every kernel in the list is this repository's except torch's `fill` (the L2 flush) and `reduce` (one histogram total).  C2 runs `sweep_kernel_g4s<1, 0, 14, 6, 1, 0>` (4 lanes per replica, 14 unrolled pair iterations), C1 `sweep_kernel_g2<1, 0, 6, 6, 1, 1>` (2 lanes, rotated order), C3 `sweep_kernel_eam<8, 0, 7, 1>`, C4 `sweep_kernel_p16<2, 14>` (16 lanes, power sums), C5 `runner_moves_kernel`."%(100*sw/tot)])+"\n")
agg,seq=load('/root/repo/gpurun_out/r02_launches_bench_c2.csv')
t,tot=table(agg)
hk=[k for k in agg if 'sweep_kernel' in k]
sw=sum(agg[k][1] for k in hk)
fill=[tt for k,tt in seq if 'FillFunctor' in k]
big=sorted([tt for k,tt in seq if 'sweep_kernel' in k])
open('/root/repo/profiles/r02_launches_bench_c2.md','w').write("\n".join(["# Launch list of the headline command `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs \"\"` (C2 only)\n",
"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (profiles/r02_launches_bench_c2.csv; kernels serialised and cold-cache under the profiler: shares, not absolute times).  The command holds the pilot (64 ladders x 3000 cycles, one launch), the equilibration, 1 warm-up + 2 timed steps of 200 cycles, and the 1 + 3 end-to-end steps of the two half-batches.\n"]+t+
["\nIn a timed step (one L2-flush `fill` of %.2f ms + one `%s` launch of 200 cycles) the sweep kernel is %.2f %% of the GPU time, in line with bench.py's CUDA events (sweep kernel 155.7 ms of a 155.9 ms step); over the whole command (pilot, equilibration, warm-up, timed and end-to-end steps) it is %.1f %%."%(fill[0]/1e6, hk[0][:60], 100*big[-2]/(big[-2]+fill[0]), 100*sw/tot)])+"\n")
print(open('/root/repo/profiles/r02_launches_bench_c2.md').read()[-900:])
