"""Checkpoint wire formats of the reference (src/ReadSave.jl), so that a run of the B200 sweep writes a
`./checkpoint/` directory the reference's `read_init` / `rebuild_states` / `postprocess` read back unchanged, and
restarts from a directory the reference wrote.

    params.data      writeparams + writeensemble + writepotential   src/ReadSave.jl:24-106
    hist_info.data   save_histparams                                 :115-126
    configT<i>.xyz   checkpoint_config / save_configs                :130-175
    index.txt, histograms.data, volume_hist.data, rdf.data           checkpoint :181-218
    readers          read_init :330-344, read_checkpoint_config :350-374, setresults :398-432, rebuild_states :437-483

All files are whitespace- or tab-delimited text as Julia's `writedlm` produces it; numbers are printed in Julia's
shortest round-trip form (`1.0e-5`, `100000.0`, `1.0e6`), which both `readdlm` and Python's `float` parse.
Host-side only: nothing here touches the device."""
from __future__ import annotations

import math
import pathlib

import numpy as np

from . import abi
from . import host as H


# ----------------------------------------------------------------------------------- number format --
def jl_float(x: float) -> str:
    """`print(::Float64)` of Julia: shortest digits that round-trip, fixed notation for 1e-5 < |x| < 1e6
    (decimal point position -4..6), otherwise `d.ddde±x`; always with a fractional part."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    r = repr(abs(x))
    if "e" in r:
        m, e = r.split("e")
        exp10 = int(e)
    else:
        m, exp10 = r, 0
    ip, _, fp = m.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        first = len(ip.lstrip("0")) - 1 + exp10
    else:
        first = -(len(fp) - len(fp.lstrip("0")) + 1) + exp10
    digits = digits.rstrip("0") or "0"
    pt = first + 1                                        # position of the decimal point relative to the digits
    if -4 < pt <= 6:
        if pt <= 0:
            s = "0." + "0" * (-pt) + digits
        elif pt >= len(digits):
            s = digits + "0" * (pt - len(digits)) + ".0"
        else:
            s = digits[:pt] + "." + digits[pt:]
    else:
        s = digits[0] + "." + (digits[1:] or "0") + "e" + str(first)
    return sign + s


def _jl(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return jl_float(float(v))
    return str(v)


def _row(values, delim=" ") -> str:
    return delim.join(_jl(v) for v in values) + "\n"


# ------------------------------------------------------------------------------------ static files --
def _type_name(obj) -> str:
    """`typeof(x)` as Julia prints it for the types of this path."""
    if isinstance(obj, H.ELJPotentialEven):
        return "ELJPotentialEven{%d, Float64}" % len(obj.coeff)
    if isinstance(obj, H.ELJPotential):
        return "ELJPotential{%d, Float64}" % len(obj.coeff)
    if isinstance(obj, H.SphericalBC):
        return "SphericalBC{Float64}"
    if isinstance(obj, H.CubicBC):
        return "CubicBC{Float64}"
    if isinstance(obj, H.RhombicBC):
        return "RhombicBC{Float64}"
    if isinstance(obj, H.RectangularBC):
        return "RectangularBC{Float64}"
    raise TypeError(type(obj))


def params_text(potential, ensemble, params: H.MCParams, temp: H.TempGrid) -> str:
    """Contents of params.data (save_init :93-106)."""
    out = "cycles: sample_rate: n_traj: n_atoms: n_adjust n_bins min_acc: max_acc t_i: t_f\n"
    # the value row is a Matrix{Float64} in the reference: integers are printed as floats
    out += _row([float(params.mc_cycles), float(params.mc_sample), float(params.n_traj), float(params.n_atoms),
                 float(params.n_adjust), float(params.n_bin), float(params.min_acc), float(params.max_acc),
                 float(temp.t_grid[0]), float(temp.t_grid[-1])])
    if ensemble.kind == abi.ENS_NPT:
        out += "ensemble n_atom_moves n_volume_moves n_atom_swaps pressure separated_volume\n"
        out += _row(["NPT", ensemble.n_atoms, ensemble.atom_moves(), ensemble.n_volume_moves, 0, float(ensemble.pressure),
                     bool(ensemble.separated_volume)])
    elif ensemble.kind == abi.ENS_NNVT:                   # writeensemble(::NNVT) :51-56
        out += "ensemble n_1 n_2 n_atom_moves n_atom_swaps\n"
        out += _row(["NNVT", ensemble.natoms[0], ensemble.natoms[1], ensemble.atom_moves(), ensemble.n_atom_swaps])
    else:
        out += "ensemble n_atom_moves n_atom_swaps\n"
        out += _row(["NVT", ensemble.n_atoms, ensemble.atom_moves(), ensemble.n_atom_swaps])
    if isinstance(potential, (H.ELJPotential, H.ELJPotentialEven)):
        out += _type_name(potential) + " " + _row([float(c) for c in potential.coeff])
    elif isinstance(potential, H.EmbeddedAtomPotential):
        out += "EAM: %s %s %s %s \n" % (jl_float(potential.n), jl_float(potential.m), jl_float(potential.ean), jl_float(potential.eCam))
    elif isinstance(potential, (H.RuNNerPotential, H.RuNNerPotential2Atom)):
        out += "runnerpotential \n"                       # writepotential(::AbstractMachineLearningPotential) :82-84
    elif isinstance(potential, H.ELJPotentialB):
        # writepotential(::AbstractDimerPotentialB) :70-76: header line, then writedlm([a', b', c']) with its default tab delimiter
        out += "ELJB %d \n" % len(potential.coeff_a)
        for row in (potential.coeff_a, potential.coeff_b, potential.coeff_c):
            out += _row([float(v) for v in row], "\t")
    elif isinstance(potential, H.LookupTablePotential):
        out += "LookupTablePotential \n"                  # :85-87 (the table itself is not saved; readpotential has no branch for it)
    else:
        raise TypeError(type(potential))
    return out


def save_init(potential, ensemble, params, temp, directory="./checkpoint"):
    d = pathlib.Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    f = d / "params.data"
    if not f.exists():                                   # :99 `if !ispath(...)`
        f.write_text(params_text(potential, ensemble, params, temp))


def save_histparams(results: H.Output, directory="./checkpoint"):
    text = "e_min e_max v_min v_max Δ_E Δ_V Δ_r2\n"
    text += _row([float(results.en_min), float(results.en_max), float(results.v_min), float(results.v_max),
                  float(results.delta_en_hist), float(results.delta_v_hist), float(results.delta_r2)])
    pathlib.Path(directory, "hist_info.data").write_text(text, encoding="utf-8")


def bc_info(bc) -> str:
    if isinstance(bc, H.SphericalBC):
        return "%s r2: %s" % (_type_name(bc), jl_float(bc.radius2))
    if isinstance(bc, H.CubicBC):
        return "%s box_length: %s" % (_type_name(bc), jl_float(bc.box_length))
    return "%s box_dims: %s, %s" % (_type_name(bc), jl_float(bc.box_length), jl_float(bc.box_height))


def config_text(positions, bc, max_displ, count_atom: int, count_vol: int) -> str:
    """checkpoint_config :141-159: xyz with the boundary condition, max_displ[1:3] and the two total counters in
    the comment line; atom type placeholder 1."""
    pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    out = "%d\n" % len(pos)
    out += "%s %s %s %s %d %d\n" % (bc_info(bc), jl_float(max_displ[0]), jl_float(max_displ[1]), jl_float(max_displ[2]),
                                   int(count_atom), int(count_vol))
    for r in pos:
        out += "1  %s %s %s \n" % (jl_float(r[0]), jl_float(r[1]), jl_float(r[2]))
    return out


def _bc_of(kind: int, box) -> object:
    if kind == abi.BC_SPHERICAL:
        return H.SphericalBC(math.sqrt(float(box[0])))
    if kind == abi.BC_CUBIC:
        return H.CubicBC(float(box[0]))
    if kind == abi.BC_RHOMBIC:
        return H.RhombicBC(float(box[0]), float(box[1]))
    return H.RectangularBC(float(box[0]), float(box[1]))


def save_configs(states: dict, ladder: int, bc_kind: int, directory="./checkpoint", filename="config"):
    """save_configs :168-175 for one ladder of `Engine.download_states()`."""
    d = pathlib.Path(directory)
    n_traj = states["pos"].shape[1]
    for t in range(n_traj):
        bc = _bc_of(bc_kind, states["box"][ladder, t])
        if bc_kind == abi.BC_SPHERICAL:
            # radius2 is stored, not radius: print the stored value exactly
            bc = _ExactSphere(float(states["box"][ladder, t, 0]))
        text = config_text(states["pos"][ladder, t], bc, states["max_displ"][ladder, t],
                           states["counters"][ladder, t, abi.C_ATOM_TOT], states["counters"][ladder, t, abi.C_VOL_TOT])
        (d / f"{filename}T{t + 1}.xyz").write_text(text)


class _ExactSphere(H.SphericalBC):
    """SphericalBC carrying radius^2 verbatim (the reference stores radius2, src/BoundaryConditions.jl:44-47)."""

    def __init__(self, radius2):
        object.__setattr__(self, "radius", math.sqrt(radius2))
        object.__setattr__(self, "_r2", radius2)

    @property
    def radius2(self):
        return self._r2


def checkpoint(index: int, states: dict, results: H.Output, ensemble, bc_kind: int, *, ladder: int = 0, rdf=None,
               directory="./checkpoint"):
    """checkpoint :181-218.  `results.en_histogram` [n_traj][n_bin+2]; NPT also writes `results.ev_histogram`
    [n_traj][n_bin+2][n_bin+2] (every matrix flattened column-major into one row, as writedlm does)."""
    d = pathlib.Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    (d / "index.txt").write_text("%d\n" % index)
    save_configs(states, ladder, bc_kind, directory)
    n_traj = states["pos"].shape[1]
    nb2 = results.n_bin + 2
    enh = results.en_histogram if results.en_histogram is not None else np.zeros((n_traj, nb2))
    with open(d / "histograms.data", "w") as f:
        for row in np.asarray(enh, dtype=np.float64):
            f.write(_row([float(v) for v in row], "\t"))
    if ensemble.kind == abi.ENS_NPT:
        with open(d / "volume_hist.data", "w") as f:
            for m in np.asarray(results.ev_histogram, dtype=np.float64):
                f.write(_row([float(v) for v in m.T.reshape(-1)], "\t"))     # column-major: evmat[:, j] contiguous
    if rdf is not None:
        with open(d / "rdf.data", "w") as f:
            for row in np.asarray(rdf, dtype=np.float64):
                f.write(_row([float(v) for v in row], "\t"))


# ------------------------------------------------------------------------------------------ readers --
def _tokens(line: str):
    """One readdlm row: numbers where they parse, strings otherwise."""
    out = []
    for t in line.split():
        try:
            out.append(float(t))
        except ValueError:
            out.append(t)
    return out


def read_init(directory="./checkpoint", restart=True, eq_cycles=0.2):
    """read_init :330-344 -> (mc_params, temp, ensemble, potential).  `runnerpotential` returns the string marker
    (the reference `include`s ./potentialfile.jl there)."""
    lines = pathlib.Path(directory, "params.data").read_text().splitlines()
    p = _tokens(lines[1])
    cycles = int(p[0])
    eq_pct = 0.0 if restart else math.floor(eq_cycles * cycles) / cycles
    params = H.MCParams(mc_cycles=cycles, n_traj=int(p[2]), n_atoms=int(p[3]), eq_percentage=eq_pct, mc_sample=int(p[1]),
                        n_adjust=int(p[4]), n_bin=int(p[5]), min_acc=float(p[6]), max_acc=float(p[7]))
    temp = H.TempGrid(float(p[8]), float(p[9]), int(p[2]))
    e = _tokens(lines[3])
    if e[0] == "NVT":
        ens = H.NVT(int(e[1]), int(e[2]), int(e[3]))
    elif "NPT" in str(e[0]):
        ens = H.NPT(int(e[1]), float(e[5]), str(e[6]) == "true", n_atom_moves=int(e[2]), n_volume_moves=int(e[3]))
    elif "NNVT" in str(e[0]):                             # readensemble :238-242
        ens = H.NNVT((int(e[1]), int(e[2])), n_atom_swaps=int(e[4]), n_atom_moves=int(e[3]))
    else:
        raise ValueError("ensemble %r is outside the sweep path" % (e[0],))
    q = lines[4].split()
    if "ELJPotentialEven" in q[0]:
        n = int(q[0][17])                                 # potinfovec[1,1][18] (1-based)
        pot = H.ELJPotentialEven(tuple(float(v) for v in q[2:2 + n]))
    elif "ELJPotential{" in q[0]:
        n = int(q[0][13])
        pot = H.ELJPotential(tuple(float(v) for v in q[2:2 + n]))
    elif q[0] == "EAM:":
        pot = H.EmbeddedAtomPotential.__new__(H.EmbeddedAtomPotential)
        pot.n, pot.m, pot.ean, pot.eCam = (float(v) for v in q[1:5])
    elif q[0] == "runnerpotential":
        pot = "runnerpotential"
    elif q[0] == "ELJB":                                  # readpotential :266-272: rows 2-4 of the potential block are a, b, c
        n = int(float(q[1]))
        a, b, c = (tuple(float(v) for v in lines[5 + k].split()[:n]) for k in range(3))
        pot = H.ELJPotentialB(a, b, c)
    elif q[0] == "LookupTablePotential":
        raise ValueError("params.data names a LookupTablePotential: the reference saves only this marker (src/ReadSave.jl:85-87) and its "
                         "readpotential has no branch for it - pass the potential (with its table file) to ptmc_run(restart=True) yourself")
    else:
        raise ValueError("potential %r is outside the sweep path" % (q[0],))
    return params, temp, ens, pot


def read_checkpoint_config(text: str):
    """read_checkpoint_config :350-374 -> (Config, max_displ[3], count_atom, count_vol)."""
    lines = text.splitlines()
    n = int(float(lines[0].split()[0]))
    c = lines[1].split()
    num = lambda s: float(s.rstrip(","))
    if "SphericalBC" in c[0]:
        bc = _ExactSphere(num(c[2])); k = 3
    elif "CubicBC" in c[0]:
        bc = H.CubicBC(num(c[2])); k = 3
    elif "RhombicBC" in c[0]:
        bc = H.RhombicBC(num(c[2]), num(c[3])); k = 4
    elif "RectangularBC" in c[0]:
        bc = H.RectangularBC(num(c[2]), num(c[3])); k = 4     # (the reference's reader has no branch for it)
    else:
        raise ValueError("unknown boundary condition %r" % c[0])
    max_displ = [num(c[k]), num(c[k + 1]), num(c[k + 2])]
    count_atom, count_vol = int(num(c[k + 3])), int(num(c[k + 4]))
    pos = np.array([[float(v) for v in ln.split()[1:4]] for ln in lines[2:2 + n]], dtype=np.float64)
    return H.Config(pos, bc), max_displ, count_atom, count_vol


def setresults(directory="./checkpoint", npt=False) -> tuple[H.Output, np.ndarray | None]:
    """setresults :398-432 from hist_info.data / histograms.data / volume_hist.data / rdf.data."""
    d = pathlib.Path(directory)
    info = [float(v) for v in (d / "hist_info.data").read_text(encoding="utf-8").splitlines()[1].split()]
    enh = np.loadtxt(d / "histograms.data", ndmin=2)
    nb2 = enh.shape[1]
    r = H.Output(nb2 - 2, en_min=info[0], en_max=info[1], v_min=info[2], v_max=info[3])
    r.delta_en_hist, r.delta_v_hist, r.delta_r2 = info[4], info[5], info[6]
    r.en_histogram = enh
    if npt:
        ev = np.loadtxt(d / "volume_hist.data", ndmin=2)
        r.ev_histogram = ev.reshape(ev.shape[0], nb2, nb2).transpose(0, 2, 1).copy()   # rows are column-major matrices
    rdf = np.loadtxt(d / "rdf.data", ndmin=2) if (d / "rdf.data").exists() else None
    return r, rdf


def rebuild_states(n_traj: int, directory="./checkpoint", npt=False):
    """rebuild_states :437-483 -> (configs[n_traj], max_displ[n_traj][3], count_atom[n_traj], count_vol[n_traj],
    results, rdf, index)."""
    d = pathlib.Path(directory)
    results, rdf = setresults(directory, npt)
    cfgs, md, ca, cv = [], [], [], []
    for t in range(1, n_traj + 1):
        c, m, a, v = read_checkpoint_config((d / f"configT{t}.xyz").read_text())
        cfgs.append(c); md.append(m); ca.append(a); cv.append(v)
    index = int(float((d / "index.txt").read_text().split()[0])) if (d / "index.txt").exists() else 0
    return cfgs, np.array(md), np.array(ca), np.array(cv), results, rdf, index


def postprocess(directory="./checkpoint", n_points=1000, outdir=None):
    """postprocess() src/multihist.jl:463-485 on a checkpoint directory: rebuild the results from the checkpoint files, run the
    multihistogram analysis and return (energies, histogramdata, T, Z, Cv, dCv, S) as the reference does (S is the fifth column
    of analysis.NVT, U(T)).  The analysis writes its four files (histograms.data with the bin energies in the first row, Sol.X,
    S.data, analysis.NVT) to `outdir`, by default the PARENT of the checkpoint directory: the reference works in pwd() with the
    checkpoint in ./checkpoint, so the checkpoint's own histograms.data (the raw counts a restart reads) is left alone."""
    from . import multihist
    directory = pathlib.Path(directory)
    outdir = directory.resolve().parent if outdir is None else pathlib.Path(outdir)
    if outdir.resolve() == directory.resolve():
        raise ValueError("postprocess: outdir must differ from the checkpoint directory (the analysis writes its own histograms.data)")
    params, temp, ens, pot = read_init(directory)
    results, _ = setresults(directory, npt=ens.kind == abi.ENS_NPT)
    out = multihist.multihistogram(results.en_histogram, temp.beta_grid, results.en_min, results.en_max, results.n_bin,
                                   outdir=outdir, n_points=n_points)
    hists = [[float(v) for v in ln.split()] for ln in pathlib.Path(outdir, "histograms.data").read_text().splitlines()]
    energies, histogramdata = np.array(hists[0]), [np.array(h) for h in hists[1:1 + params.n_traj]]
    return energies, histogramdata, out["T"], out["Z"], out["Cv"], out["dCv"], out["U"]
