#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ with the CPU oracle (oracle/ptmc_oracle.c, itself
pinned to the reference's known-answer tests by tests/test_oracle_kat.py).

Each fixture is one replayed trajectory: the deck and the stream seed define the inputs (the replay streams are
regenerated from the seed by tests/parity.py::make_replay), the file holds the outputs: per-move index / delta
energy / accept bit, per-cycle exchange record, final positions, energies, counters, step sizes, histogram.

    python tests/tools/make_golden.py [case ...]     # rewrites tests/golden/<case>.npz (all cases without arguments)
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import ptmc_b200 as P  # noqa: E402
from golden_cases import CASES, build_case  # noqa: E402
from parity import run_engine  # noqa: E402


def main():
    from oracle import oracle
    out = ROOT / "tests" / "golden"
    out.mkdir(exist_ok=True)
    for name in (sys.argv[1:] or CASES):
        cfg, pos, box, slot, lad, n_cycles, bounds, kw = build_case(name, oracle)
        res = run_engine(oracle.OracleEngine, cfg, pos, box, slot, lad, n_cycles, bounds, **kw)
        tr, st = res["trace"], res["states"]
        np.savez_compressed(out / f"{name}.npz", move_index=tr["move_index"], delta_en=tr["delta_en"], accepted=tr["accepted"],
                            exchange=tr["exchange"], pos=st["pos"], box=st["box"], en_tot=st["en_tot"], max_displ=st["max_displ"],
                            counters=st["counters"], ham=st["ham"], hist=res["hist"], init_en=res["init_en"],
                            bounds=np.array([np.nan if b is None else b for b in bounds]))
        print(name, "moves", tr["accepted"].size, "accept rate %.3f" % tr["accepted"].mean(), "exchanges", int(tr["exchange"][..., 0].sum()),
              "hist under/overflow", (lambda h: float(h[:, :, 0].sum() + h[:, :, -1].sum()) / max(1.0, float(h.sum())))(res["hist"]),
              "E range", float(st["en_tot"].min()), float(st["en_tot"].max()))


if __name__ == "__main__":
    main()
