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
