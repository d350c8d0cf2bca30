"""The reference arm of bench.py (the oracle timed on the host cores) runs without a GPU and prints the contract's JSON line."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-ladders", "8",
                          "--cpu-cycles", "20", "--configs", "c1,c3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "moves/s" and line["higher_is_better"] is True
    assert line["metric"] == "MC moves/sec (all replicas)" and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "ladders" in cb["sample"]
    assert "workload" in line["config"]
    # the other BASELINE configs ride along as sub-records with their own CPU sample
    assert set(line["configs"]) == {"c1", "c3"} and all(r["value"] > 0 and "workload" in r["config"] for r in line["configs"].values())
