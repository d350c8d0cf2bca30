"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
the ctypes mirror matches the header, and no compute entry point works without a GPU (no CPU fallback)."""
import ctypes as C
import re

import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi
from ptmc_b200.host import initial_arrays, sweep_config


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    return abi.load_library()


def test_header_symbols_exported(lib):
    header = (abi.PKG_DIR.parent / "include" / "ptmc_b200.h").read_text()
    declared = set(re.findall(r"^PTMC_API [\w\s\*]+?\b(ptmc_\w+|forward)\(", header, flags=re.M))
    assert declared == set(abi.SYMBOLS), declared ^ set(abi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ptmc_abi_version() == abi.PTMC_ABI_VERSION


def test_config_struct_matches_header():
    # sizes the C compiler sees (computed by compiling a probe would need gcc + the header: do exactly that)
    import subprocess, tempfile, pathlib
    src = '#include "ptmc_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(ptmc_config), sizeof(ptmc_runner_desc), sizeof(ptmc_runner2_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = pathlib.Path(d)
        (p / "probe.c").write_text(src)
        subprocess.run(["gcc", "-I", str(abi.PKG_DIR.parent / "include"), str(p / "probe.c"), "-o", str(p / "probe")], check=True)
        out = subprocess.run([str(p / "probe")], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == C.sizeof(abi.Config)
    assert int(out[1]) == C.sizeof(abi.RunnerDesc)
    assert int(out[2]) == C.sizeof(abi.Runner2Desc)


def _header_fields(struct_name):
    header = (abi.PKG_DIR.parent / "include" / "ptmc_b200.h").read_text()
    i = header.index("typedef struct %s" % struct_name)
    body = re.sub(r"/\*.*?\*/", "", header[i:header.index("} %s;" % struct_name, i)], flags=re.S)
    names = []
    for stmt in body.replace("typedef struct %s {" % struct_name, "").split(";"):
        m = re.match(r"\s*(?:const\s+)?\w+\s*\*?\s+(.*)", stmt, re.S)
        if m:
            names += [re.sub(r"\[.*\]", "", d.strip().lstrip("*").strip()) for d in m.group(1).split(",") if d.strip()]
    return names


def test_bindings_list_the_header_fields_in_order():
    # the ctypes mirror (what the tests drive) and the Julia struct a maintainer adds against include/ptmc_b200.h
    for struct_name, ctype, jl_name in (("ptmc_config", abi.Config, "PtmcConfig"), ("ptmc_runner_desc", abi.RunnerDesc, "PtmcRunnerDesc"),
                                         ("ptmc_runner2_desc", abi.Runner2Desc, "PtmcRunner2Desc")):
        fields = _header_fields(struct_name)
        assert fields == [f[0] for f in ctype._fields_], struct_name
        jl = (abi.PKG_DIR / "julia" / "PTMCB200.jl").read_text()
        i = jl.index("struct %s" % jl_name)
        if True:
            assert fields == re.findall(r"^\s*(\w+)::", jl[i:jl.index("\nend", i)], flags=re.M), jl_name


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device failure path cannot be exercised")
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=4, cycles=5)
    with pytest.raises(P.PtmcError) as e:
        P.Engine(sweep_config(mcp, temp, start, pot, ens))
    assert e.value.code == abi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_argument_validation_precedes_device(lib):
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=4, cycles=5)
    cfg = sweep_config(mcp, temp, start, pot, ens)
    c = cfg.to_c()
    c.abi_version = 99
    h = C.c_void_p()
    assert lib.ptmc_create(C.byref(c), C.byref(h)) == abi.ERR_ARG
    assert b"abi_version" in lib.ptmc_last_error(None)
    c = cfg.to_c()
    c.ensemble = abi.ENS_NPT
    c.n_volume_moves = 1
    assert lib.ptmc_create(C.byref(c), C.byref(h)) == abi.ERR_ARG
    assert b"SphericalBC cannot be used in an NPT ensemble" in lib.ptmc_last_error(None)


def test_lookup_table_needs_every_angle_row(lib):
    # d_angle = 0.5 -> the angle index reaches round(90 / 0.5 + 1) = 181: a 91-row table would be read out of bounds (the reference
    # throws a BoundsError); refused at create, before any device is touched
    mcp, temp, start, _, ens = P.host.deck_ne13(n_traj=4, cycles=5)
    lut = P.LookupTablePotential(np.ones((91, 8)), start_dist=2.0, start_angle=0.0, d_dist=0.5, d_angle=0.5, c6coeff=-1.0)
    c = sweep_config(mcp, temp, start, lut, ens).to_c()
    h = C.c_void_p()
    assert lib.ptmc_create(C.byref(c), C.byref(h)) == abi.ERR_ARG
    assert b"needs 181" in lib.ptmc_last_error(None)


def test_two_species_host_mirror():
    # RuNNerPotential2Atom / NNVT (scripts/brass55.jl): the host mirror builds the descriptor in g-matrix row order and refuses
    # inconsistent combinations
    pot = P.host.brass_runner_potential(4, 2)
    assert pot.params.n_sym == 88 and pot.params.gmax_gmin_1.shape == (88, 2) and pot.kind == abi.POT_RUNNER2
    g = P.host.load_scaling()
    assert np.array_equal(pot.params.gmax_gmin_1, g[:88]) and np.array_equal(pot.params.gmax_gmin_2, g[88:176])
    assert P.NNVT((4, 2)).atom_moves() == 6 and P.NNVT((4, 2)).n_atom_swaps == 1          # src/Ensembles.jl:144
    cfg = P.Config(np.zeros((6, 3)) + np.arange(6)[:, None], P.SphericalBC(radius=8.0))
    sc = sweep_config(P.MCParams(10, 2, 6), P.TempGrid(500.0, 650.0, 2), cfg, pot, P.NNVT((4, 2), n_atom_swaps=2))
    assert sc.n_steps == 8 and sc.to_c().n_atom_swaps == 2
    with pytest.raises(ValueError, match="NNVT"):
        sweep_config(P.MCParams(10, 2, 6), P.TempGrid(500.0, 650.0, 2), cfg, pot, P.NVT(6))
    with pytest.raises(ValueError, match="natoms"):
        sweep_config(P.MCParams(10, 2, 6), P.TempGrid(500.0, 650.0, 2), cfg, pot, P.NNVT((3, 3)))
