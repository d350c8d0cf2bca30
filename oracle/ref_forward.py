"""Runs the reference's own `forward` (oracle/_ref/librunnerjulia.so) in a child process whose LD_LIBRARY_PATH
points at oracle/_ref (the binary's NEEDED libblas.so.3 / libgfortran.so.5 are resolved by soname).

    python oracle/ref_forward.py in.npz out.npy      (in.npz: input [batch][n_in], nodes, act, params)
or from Python: ref_forward(input, nodes, act, params) -> eatom.   TEST INFRASTRUCTURE."""
import ctypes as C
import os
import pathlib
import subprocess
import sys
import tempfile

import numpy as np

REF = pathlib.Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF / "librunnerjulia.so").exists() and (REF / "libblas.so.3").exists()


def _child(inp, out):
    d = np.load(inp)
    x = np.ascontiguousarray(d["input"], dtype=np.float64)
    nodes = np.ascontiguousarray(d["nodes"], dtype=np.int32)
    act = np.ascontiguousarray(d["act"], dtype=np.int32)
    params = np.ascontiguousarray(d["params"], dtype=np.float64)
    lib = C.CDLL(str(REF / "librunnerjulia.so"))
    lib.forward.restype = C.c_double
    e = np.zeros(x.shape[0])
    i32 = lambda v: C.byref(C.c_int32(v))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lib.forward(dp(x), i32(int(nodes[0])), i32(len(nodes)), ip(nodes), i32(x.shape[0]), i32(len(params)), dp(params), ip(act), dp(e))
    np.save(out, e)


def ref_forward(inp, nodes, act, params):
    with tempfile.TemporaryDirectory() as t:
        a, b = os.path.join(t, "in.npz"), os.path.join(t, "out.npy")
        np.savez(a, input=inp, nodes=nodes, act=act, params=params)
        env = dict(os.environ, LD_LIBRARY_PATH=str(REF) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), OPENBLAS_NUM_THREADS="1")
        subprocess.run([sys.executable, __file__, a, b], check=True, env=env)
        return np.load(b)


if __name__ == "__main__":
    _child(sys.argv[1], sys.argv[2])
