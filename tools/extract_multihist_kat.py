#!/usr/bin/env python3
"""Extract the multihistogram known-answer vector of the reference (test/multihist_test.jl:4-14: a literal Output with the
energy histograms of a Ne13 run on TempGrid{32}(5,16), and four expected heat-capacity values) into
tests/golden/multihist_nvt_kat.json.  Numbers only; run where /root/reference exists."""
import json
import pathlib
import re

REF = pathlib.Path("/root/reference/test/multihist_test.jl")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "multihist_nvt_kat.json"


def main():
    s = REF.read_text()
    a = s.index("results=Output{Float64}(") + len("results=Output{Float64}(")
    b = s.index("cv=mul")
    lit = s[a:s.rindex(")", a, b)]
    lit = re.sub(r"\b\w+(\{[^\[\]]*?\})?\[\]", "[]", lit)          # Float64[] / Vector{Float64}[] -> []
    args = json.loads("[" + lit + "]")
    n_bin, en_min, en_max = args[0], args[1], args[2]
    hist = next(x for x in args if isinstance(x, list) and x and isinstance(x[0], list) and len(x[0]) == n_bin + 2)
    exp = {int(i): float(v) for i, v in re.findall(r"@test cv\[(\d+)\]==([0-9.e+-]+)", s[b:b + 400])}
    OUT.write_text(json.dumps({"source": "test/multihist_test.jl:4-14", "n_bin": n_bin, "en_min": en_min, "en_max": en_max,
                               "temp": [5, 16, 32], "conv_threshold": 1e-3, "en_histogram": hist, "cv_expected_1based": exp}))
    print("wrote", OUT, len(hist), "x", len(hist[0]), exp)


if __name__ == "__main__":
    main()


def main_npt():
    """test/multihist_test.jl:16-30: NPT (E,V) histograms of a 32-atom run on TempGrid{3}(20,40) and four expected Cp values."""
    import numpy as np
    s = REF.read_text()
    t0 = s.index('@testset "multihist_NPT"')
    pres = float(re.search(r"NPT\(32,\s*([0-9.e+-]+)", s[t0:t0 + 300]).group(1))
    a = s.index("results = Output{Float64}(", t0) + len("results = Output{Float64}(")
    b = s.index("cp_3=", a)
    lit = s[a:s.rindex(")", a, b)]
    head = re.match(r"\s*(\d+),\s*([-0-9.e+]+),\s*([-0-9.e+]+),\s*([-0-9.e+]+),\s*([-0-9.e+]+)", lit)
    n_bin = int(head.group(1))
    en_min, en_max, v_min, v_max = (float(head.group(k)) for k in range(2, 6))
    mats = []
    for m in re.finditer(r"\[([-0-9.e+ ;]+)\]", lit):           # Julia matrix literals: rows separated by ';'
        if ";" not in m.group(1):
            continue
        rows = [[float(x) for x in r.split()] for r in m.group(1).split(";")]
        mats.append(rows)
    assert len(mats) == 3 and all(len(x) == n_bin + 2 and len(x[0]) == n_bin + 2 for x in mats), [np.shape(x) for x in mats]
    exp = {int(i): float(v) for i, v in re.findall(r"@test cp_3\[(\d+)\]==([0-9.e+-]+)", s[b:b + 600])}
    out = OUT.parent / "multihist_npt_kat.json"
    out.write_text(json.dumps({"source": "test/multihist_test.jl:16-30", "n_bin": n_bin, "en_min": en_min, "en_max": en_max,
                               "v_min": v_min, "v_max": v_max, "pressure": pres, "temp": [20, 40, 3], "conv_threshold": 1e-3,
                               "ev_histogram": mats, "cp_expected_1based": exp}))
    print("wrote", out, exp, "pressure", pres)


if __name__ == "__main__":
    main_npt()
