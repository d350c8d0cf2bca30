"""tests/c/npt_sequence.c: INTEGRATION.md's call sequence for an NPT deck from plain C (gcc + include/ptmc_b200.h + dlopen, no Python
in the loop) — the closest stand-in for the Julia `ccall` path this image allows.  Built on the CPU; run on the GPU."""
import pathlib
import subprocess

import pytest

from ptmc_b200 import abi

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _build(tmp_path):
    exe = tmp_path / "npt_sequence"
    r = subprocess.run(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "npt_sequence.c"),
                        "-o", str(exe), "-ldl", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    r = subprocess.run([str(exe), str(abi.LIB_PATH)], capture_output=True, text=True)
    # argument validation ran (the bad config was refused with a message), then ptmc_create fails: no CPU fallback
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_call_sequence_npt(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), str(abi.LIB_PATH)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "npt_sequence ok" in r.stdout
    print(r.stdout.strip())
