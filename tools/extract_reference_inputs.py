#!/usr/bin/env python3
"""Extract the *input data* (start geometries, NNP weights/scaling) that the reference's run scripts
embed, into plain data files under paralleltemperingmontecarlo.jl_b200/data/.

Only numbers are extracted (cluster coordinates, network weights); no reference code is copied.
Run in the build container (needs /root/reference); outputs are committed.

Sources (all under /root/reference):
  examples/ne13.jl:40-54      Ne13 icosahedron (Angstrom)
  scripts/ne55.jl:65-121      Ne55 (used as given)
  scripts/cu55.jl:164-220     ico_55 unit icosahedron
  scripts/pbc216.jl:79-515    fcc / hcp 216-atom lattices (Angstrom)
  scripts/cu38.jl:161-200     Cu38 (bohr), scripts/pbc32-short.jl:54-87  32-atom fcc block (Angstrom)
  scripts/brass55.jl:172-228  ico_55 of the brass script (bohr, used as given :230-233)
  scripts/data/weights.029.data, weights.030.data, scripts/data/scaling.data  (RuNNer Cu / Zn networks)
"""
import json, re, sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "paralleltemperingmontecarlo.jl_b200" / "data"

ROW = re.compile(r"\[\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\]")


def grab(path, name):
    lines = (REF / path).read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if re.match(rf"\s*{name}\s*=\s*\[", l))
    rows = []
    for l in lines[start:]:
        m = ROW.search(l)
        if m:
            rows.append([m.group(1), m.group(2), m.group(3)])
        if l.strip().startswith("]") and rows:
            break
    return rows


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    geo = {
        "ne13": grab("examples/ne13.jl", "pos_ne13"),
        "ne55": grab("scripts/ne55.jl", "pos_ne55"),
        "ico55": grab("scripts/cu55.jl", "ico_55"),
        "ne216_fcc": grab("scripts/pbc216.jl", "pos_ne216_fcc"),
        "ne216_hcp": grab("scripts/pbc216.jl", "pos_ne216_hcp"),
        "cu38": grab("scripts/cu38.jl", "pos_cu38"),
        "ne32": grab("scripts/pbc32-short.jl", "pos_ne32"),
        "brass55": grab("scripts/brass55.jl", "ico_55"),
    }
    want = {"ne13": 13, "ne55": 55, "ico55": 55, "ne216_fcc": 216, "ne216_hcp": 216, "cu38": 38, "ne32": 32, "brass55": 55}
    for k, rows in geo.items():
        assert len(rows) == want[k], (k, len(rows))
        # keep the decimal literals as strings -> float() parses them exactly as Julia does
        (OUT / f"{k}.json").write_text(json.dumps(rows))
    # RuNNer inputs as numbers: the Cu network (2221 parameters, scripts/data/weights.029.data), the scaling table
    # (scripts/data/scaling.data rows 1..176: element, index, G_min, G_max, G_avg), the Zn network of the brass KAT
    import numpy as np
    np.savez_compressed(OUT / "cu_nnp_inputs.npz", weights_cu=np.loadtxt(REF / "scripts" / "data" / "weights.029.data").ravel(),
                        weights_zn=np.loadtxt(REF / "scripts" / "data" / "weights.030.data").ravel(),
                        scaling_rows=np.loadtxt(REF / "scripts" / "data" / "scaling.data", usecols=(0, 1, 2, 3, 4), max_rows=176))
    np.savez_compressed(OUT.parent.parent / "tests" / "golden" / "brass_nnp_inputs.npz",
                        weights_zn=np.loadtxt(REF / "test" / "testing_data" / "weights.030.data").ravel())
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    sys.exit(main())
