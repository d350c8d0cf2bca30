#!/usr/bin/env python3
"""Per-source-line stall samples / executed instructions of one kernel in an ncu report.

usage: tools/profile_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top_n]
Joins `ncu --page source --print-source sass` (samples per SASS address) with `nvdisasm -g` of the cubin
embedded in OBJECT.o (source line per SASS offset).  Works without a GPU."""
import collections
import csv
import io
import pathlib
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    kname = rows[hdr - 1][1] if hdr else ""
    cols = {n: k for k, n in enumerate(rows[hdr])}
    samp = []
    for r in rows[hdr + 1:]:
        if len(r) < len(cols):
            continue
        samp.append((int(r[cols["Address"]], 16), r[cols["Source"]].strip(), int(r[cols["# Samples"]] or 0), int(r[cols["Instructions Executed"]] or 0)))
    base = samp[0][0]
    with tempfile.TemporaryDirectory() as t:
        subprocess.run(["cuobjdump", "-xelf", "all", str(pathlib.Path(obj).resolve())], cwd=t, capture_output=True)
        cubin = next(pathlib.Path(t).glob("*.cubin"))
        dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    # pick the function whose mangled name matches the demangled kernel: use instruction count + substring
    funcs = re.split(r"\n\s*\.section\s+\.text\.", dis)
    best = None
    for f in funcs[1:]:
        name = f.split(",", 1)[0]
        n_ins = len(re.findall(r"/\*[0-9a-f]{4,}\*/", f))
        if sub in name and n_ins == len(samp):
            best = f
    if best is None:
        sys.exit("no function containing %r with %d instructions" % (sub, len(samp)))
    line_of = {}
    cur = None
    for ln in best.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (pathlib.Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_i = 0
    for addr, _, s, n in samp:
        k = line_of.get(addr - base)
        per[k][0] += s; per[k][1] += n
        tot_s += s; tot_i += n
    print("kernel:", kname)
    print("total stall samples %d, warp instructions %d" % (tot_s, tot_i))
    print("| file:line | samples | share | warp instructions | share |")
    print("|---|---|---|---|---|")
    for k, (s, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("| %s:%s | %d | %.1f%% | %d | %.1f%% |" % (k[0] if k else "?", k[1] if k else "?", s, 100.0 * s / max(tot_s, 1), n, 100.0 * n / max(tot_i, 1)))


if __name__ == "__main__":
    main()
