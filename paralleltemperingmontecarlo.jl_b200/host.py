"""Host-side mirror of the reference's public API for the sweep path.

The reference's host language is Julia, which this image does not have; `julia/PTMCB200.jl` is the binding
a maintainer adds.  This module restates the same *interface* in Python (same type names, argument
meaning, defaults and error behaviour) so that tests read like the reference's own tests and so that
bench.py has a reference-facing call (`ptmc_run`) to time end-to-end.  It contains host logic only
(parameter structs, the ptmc_run! driver loop, histogram bounds, finalise_results); all per-move
arithmetic happens in libptmc_b200.so.

Reference: src/InputParams.jl, src/BoundaryConditions.jl, src/Configurations.jl, src/Ensembles.jl,
src/EnergyEvaluation.jl (constructors only), src/MCRun.jl:282-432, src/Sampling.jl:141-223,377-399."""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi
from .engine import Engine, RunnerParams, SweepConfig

kB = 3.16681196e-6  # src/InputParams.jl:19
DATA_DIR = abi.PKG_DIR / "data"


# ------------------------------------------------------------------------------------ InputParams --
@dataclass
class MCParams:
    """src/InputParams.jl:35-63"""
    mc_cycles: int
    n_traj: int
    n_atoms: int
    eq_percentage: float = 0.2
    mc_sample: int = 1
    n_adjust: int = 100
    n_bin: int = 100
    min_acc: float = 0.4
    max_acc: float = 0.6

    @property
    def eq_cycles(self) -> int:
        # round(Int, x): ties to even, like Python's round
        return int(round(self.eq_percentage * self.mc_cycles))


class TempGrid:
    """src/InputParams.jl:77-93 (geometric default)"""

    def __init__(self, ti: float, tf: float, n: int, tdistr: str = "geometric"):
        if tdistr == "equally_spaced":
            delta = (tf - ti) / (n - 1)
            t = [ti + (i - 1) * delta for i in range(1, n + 1)]
        elif tdistr == "geometric":
            t = [ti * (tf / ti) ** ((i - 1) / (n - 1)) for i in range(1, n + 1)]
        else:
            raise ValueError(f"chosen temperature distribution {tdistr} does not exist")
        self.t_grid = np.array(t, dtype=np.float64)
        self.beta_grid = 1.0 / (kB * self.t_grid)

    def __len__(self):
        return len(self.t_grid)


@dataclass
class Output:
    """src/InputParams.jl:114-134 (fields the sweep path fills)"""
    n_bin: int
    en_min: float = 0.0
    en_max: float = 0.0
    v_min: float = 0.0
    v_max: float = 0.0
    delta_en_hist: float = 0.0
    delta_v_hist: float = 0.0
    delta_r2: float = 0.0
    max_displ: list = field(default_factory=list)
    en_avg: np.ndarray | None = None
    heat_cap: np.ndarray | None = None
    en_histogram: np.ndarray | None = None   # [n_traj][n_bin+2]
    ev_histogram: np.ndarray | None = None   # [n_traj][n_bin+2][n_bin+2]
    rdf: np.ndarray | None = None            # [n_traj][5 n_bin]
    count_stat_atom: np.ndarray | None = None
    count_stat_exc: np.ndarray | None = None


# ----------------------------------------------------------------------------- BoundaryConditions --
@dataclass(frozen=True)
class SphericalBC:
    """src/BoundaryConditions.jl:44-47 (stores radius^2)"""
    radius: float

    @property
    def radius2(self):
        return self.radius * self.radius

    kind = abi.BC_SPHERICAL

    def box(self):
        return (self.radius2, 0.0)


@dataclass(frozen=True)
class CubicBC:
    """src/BoundaryConditions.jl:72-77"""
    box_length: float
    kind = abi.BC_CUBIC

    def box(self):
        return (self.box_length, self.box_length)


@dataclass(frozen=True)
class RhombicBC:
    """src/BoundaryConditions.jl:110-115"""
    box_length: float
    box_height: float
    kind = abi.BC_RHOMBIC

    def box(self):
        return (self.box_length, self.box_height)


@dataclass(frozen=True)
class RectangularBC:
    """src/BoundaryConditions.jl:89-92"""
    box_length: float
    box_height: float
    kind = abi.BC_RECTANGULAR

    def box(self):
        return (self.box_length, self.box_height)


def get_volume(bc) -> float:
    """src/Configurations.jl:246-256"""
    if isinstance(bc, CubicBC):
        return bc.box_length * bc.box_length * bc.box_length
    if isinstance(bc, RhombicBC):
        return bc.box_length * bc.box_length * bc.box_height * 3 ** 0.5 / 2
    if isinstance(bc, RectangularBC):
        return bc.box_length * bc.box_length * bc.box_height
    raise TypeError("get_volume is defined for periodic boundary conditions only")


class Config:
    """src/Configurations.jl:42-58"""

    def __init__(self, positions, boundary_condition):
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        self.boundary_condition = boundary_condition

    def __len__(self):
        return len(self.positions)


# --------------------------------------------------------------------------------------- Ensembles --
@dataclass(frozen=True)
class NVT:
    """src/Ensembles.jl:42-49"""
    n_atoms: int
    n_atom_moves: int | None = None
    n_atom_swaps: int = 0
    kind = abi.ENS_NVT

    def atom_moves(self):
        return self.n_atoms if self.n_atom_moves is None else self.n_atom_moves


@dataclass(frozen=True)
class NPT:
    """src/Ensembles.jl:75-85 — NPT(n_atoms, pressure, separated_volume)"""
    n_atoms: int
    pressure: float
    separated_volume: bool = False
    n_atom_moves: int | None = None
    n_volume_moves: int = 1
    kind = abi.ENS_NPT

    def atom_moves(self):
        return self.n_atoms if self.n_atom_moves is None else self.n_atom_moves


# -------------------------------------------------------------------------------------- potentials --
@dataclass(frozen=True)
class ELJPotential:
    """src/EnergyEvaluation.jl:891-907: sum c_i r^-(5+i)"""
    coeff: tuple
    kind = abi.POT_ELJ


@dataclass(frozen=True)
class ELJPotentialEven:
    """src/EnergyEvaluation.jl:138-154: sum c_i r^-(4+2i)"""
    coeff: tuple
    kind = abi.POT_ELJ_EVEN


@dataclass(frozen=True)
class ELJPotentialB:
    """src/EnergyEvaluation.jl:909-921, dimer_energy :943-958: extended Lennard-Jones in a magnetic field,
    sum_i c_i r^-k_i (1 + a_i cos 2theta + b_i cos 4theta), theta = angle of the pair to the z axis."""
    coeff_a: tuple
    coeff_b: tuple
    coeff_c: tuple
    kind = abi.POT_ELJ_B


class LookupTablePotential:
    """src/EnergyEvaluation.jl:214-262: pair energy tabulated over (angle between the pair and the z axis, distance);
    `LookupTablePotential(file)` reads the reference's table format (read_lookuptable :231-260): line 1 `l_dist, start_dist,
    d_dist`, line 2 `l_angle, start_angle, d_angle`, then l_angle x l_dist values with the angle index slowest; `!` starts a
    comment.  c6coeff = mean_i table[i, 1000] * (start_dist + 1000 d_dist)^6 carries the tail beyond the table."""
    kind = abi.POT_LUT

    def __init__(self, source, start_dist=None, start_angle=None, d_dist=None, d_angle=None, c6coeff=None):
        if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__"):
            with open(source) as f:
                first = f.readline().split("!")[0].split(", ")
                l_dist, self.start_dist, self.d_dist = int(first[0]), float(first[1]), float(first[2])
                second = f.readline().split("!")[0].split(", ")
                l_angle, self.start_angle, self.d_angle = int(second[0]), float(second[1]), float(second[2])
                vals = [float(f.readline().split("!")[0]) for _ in range(l_angle * l_dist)]
            self.table = np.array(vals, dtype=np.float64).reshape(l_angle, l_dist)
            avg_lr = 0.0
            for i in range(l_angle):                       # sequential sum, then the division (:253-257)
                avg_lr += self.table[i, 999]
            avg_lr /= l_angle
            self.c6coeff = avg_lr * (self.start_dist + 1000 * self.d_dist) ** 6
        else:
            self.table = np.ascontiguousarray(source, dtype=np.float64)
            self.start_dist, self.start_angle, self.d_dist, self.d_angle, self.c6coeff = start_dist, start_angle, d_dist, d_angle, c6coeff
        self.l_angle, self.l_dist = self.table.shape


class EmbeddedAtomPotential:
    """src/EnergyEvaluation.jl:1059-1073: EmbeddedAtomPotential(n, m, eps, C, a)"""
    kind = abi.POT_EAM

    def __init__(self, n, m, eps, Cc, a):
        self.n, self.m = float(n), float(m)
        self.ean = eps * a ** self.n / 2
        self.eCam = eps * Cc * a ** (self.m / 2)


class RuNNerPotential:
    """src/EnergyEvaluation.jl:1182-1218 with RadialType2 / AngularType3 tables
    (MachineLearningPotential/SymmFunc.jl:24-44,86-112) and NeuralNetworkPotential (ForwardPass.jl:23-45)."""
    kind = abi.POT_RUNNER

    def __init__(self, num_nodes, activation, params, radial, angular, boundary: float = 0.0):
        # radial rows: (type, eta, r_cut, G_max, G_min); angular rows: (eta, lambda, zeta, r_cut, type, G_max, G_min)
        radial = np.asarray(radial, dtype=np.float64).reshape(-1, 5)
        angular = np.asarray(angular, dtype=np.float64).reshape(-1, 7)
        self.r_cut = float(radial[0, 2])
        self.boundary2 = float(boundary) * float(boundary)
        self.params = RunnerParams(
            rad_eta=radial[:, 1].copy(), rad_type=radial[:, 0].astype(np.int32),
            rad_gmax_gmin=radial[:, 3:5].copy(),
            ang_eta=angular[:, 0].copy(), ang_lambda=angular[:, 1].copy(), ang_zeta=angular[:, 2].copy(),
            ang_type=angular[:, 4].astype(np.int32), ang_gmax_gmin=angular[:, 5:7].copy(),
            r_cut=self.r_cut, boundary2=self.boundary2,
            num_nodes=np.asarray(num_nodes, dtype=np.int32), activation=np.asarray(activation, dtype=np.int32),
            params=np.asarray(params, dtype=np.float64))


@dataclass(frozen=True)
class NeuralNetworkPotential:
    """MachineLearningPotential/ForwardPass.jl:23-45 — NeuralNetworkPotential(num_nodes, activation_functions, parameters)"""
    num_nodes: tuple
    activation_functions: tuple
    parameters: np.ndarray


@dataclass(frozen=True)
class RadialType2a:
    """MachineLearningPotential/SymmFunc.jl:53-84 — RadialType2a{T}(eta, r_cut, type_vector, G_vals_a, G_vals_b): G_vals_x are the
    two {G_max, G_min} pairs (species-1 neighbour row, species-2 neighbour row) of the network of species x."""
    eta: float
    r_cut: float
    type_vec: int
    G_vals_a: tuple
    G_vals_b: tuple


@dataclass(frozen=True)
class AngularType3a:
    """MachineLearningPotential/SymmFunc.jl:113-171 — AngularType3a{T}(eta, lambda, zeta, r_cut, type_vector, G_valsa, G_valsb):
    three {G_max, G_min} pairs per species (0 / 1 / 2 neighbours of species 2)."""
    eta: float
    lambda_: float
    zeta: float
    r_cut: float
    type_vec: int
    G_valsa: tuple
    G_valsb: tuple


class RuNNerPotential2Atom:
    """src/EnergyEvaluation.jl:1287-1314 — RuNNerPotential2Atom(nnp1, nnp2, radsymvec, angsymvec, n1, n2[, boundary]): two-species
    network potential; atoms 1..n1 are species 1, n1+1..n1+n2 species 2."""
    kind = abi.POT_RUNNER2

    def __init__(self, nnp1, nnp2, radsymvec, angsymvec, n1, n2, boundary: float = 0.0):
        from .engine import Runner2Params
        if tuple(nnp1.num_nodes) != tuple(nnp2.num_nodes) or tuple(nnp1.activation_functions) != tuple(nnp2.activation_functions):
            raise ValueError("the two networks must share one topology")
        self.nnp1, self.nnp2, self.radsymvec, self.angsymvec = nnp1, nnp2, list(radsymvec), list(angsymvec)
        self.n1, self.n2 = int(n1), int(n2)
        self.r_cut = float(self.radsymvec[0].r_cut)                      # :1297
        self.boundary2 = float(boundary) * float(boundary)               # :1312-1314
        g1 = [row for r in self.radsymvec for row in r.G_vals_a] + [row for a in self.angsymvec for row in a.G_valsa]
        g2 = [row for r in self.radsymvec for row in r.G_vals_b] + [row for a in self.angsymvec for row in a.G_valsb]
        self.params = Runner2Params(
            n1=self.n1, n2=self.n2, rad_eta=np.array([r.eta for r in self.radsymvec], dtype=np.float64),
            ang_eta=np.array([a.eta for a in self.angsymvec], dtype=np.float64),
            ang_lambda=np.array([a.lambda_ for a in self.angsymvec], dtype=np.float64),
            ang_zeta=np.array([a.zeta for a in self.angsymvec], dtype=np.float64),
            gmax_gmin_1=np.array(g1, dtype=np.float64).reshape(-1, 2), gmax_gmin_2=np.array(g2, dtype=np.float64).reshape(-1, 2),
            r_cut=self.r_cut, boundary2=self.boundary2, num_nodes=np.asarray(nnp1.num_nodes, dtype=np.int32),
            activation=np.asarray(nnp1.activation_functions, dtype=np.int32),
            params_1=np.asarray(nnp1.parameters, dtype=np.float64), params_2=np.asarray(nnp2.parameters, dtype=np.float64))


@dataclass
class NNVT:
    """src/Ensembles.jl:138-151 — NNVT(natoms; natomswaps = 1, natommoves = sum(natoms)): two species, atom moves + atom swaps"""
    natoms: tuple
    n_atom_swaps: int = 1
    n_atom_moves: int | None = None
    kind = abi.ENS_NNVT

    def atom_moves(self):
        return sum(self.natoms) if self.n_atom_moves is None else self.n_atom_moves


# ------------------------------------------------------------------------------- config building --
def sweep_config(mc_params: MCParams, temp: TempGrid, start_config, potential, ensemble, *, n_ladders=1,
                 seed=1234, device=0, rng_mode=abi.RNG_BATCH8, lanes_per_replica=0) -> SweepConfig:
    if isinstance(ensemble, NNVT) != isinstance(potential, RuNNerPotential2Atom):
        raise ValueError("NNVT (atom swaps between two species) goes with RuNNerPotential2Atom and vice versa "
                         "(swap_energy_update is defined for that potential only, src/EnergyEvaluation.jl:1437-1448)")
    if isinstance(ensemble, NNVT) and (tuple(ensemble.natoms) != (potential.n1, potential.n2) or sum(ensemble.natoms) != mc_params.n_atoms):
        raise ValueError("NNVT.natoms must equal the potential's (n1, n2) and add up to n_atoms")
    cfgs = start_config if isinstance(start_config, (list, tuple)) else (start_config,)
    bc = cfgs[0].boundary_condition
    if isinstance(ensemble, NPT):
        if isinstance(bc, SphericalBC):
            raise ValueError("SphericalBC cannot be used in an NPT ensemble.")  # src/Ensembles.jl:189
        if ensemble.separated_volume and not isinstance(bc, (RhombicBC, RectangularBC)):
            raise ValueError("separated volume moves need a RhombicBC or RectangularBC (scale_xy / scale_z, src/MCMoves.jl:189-222)")
    kw = dict(n_ladders=n_ladders, n_traj=mc_params.n_traj, n_atoms=mc_params.n_atoms, betas=temp.beta_grid,
              pot_kind=potential.kind, bc_kind=bc.kind, ensemble=ensemble.kind,
              n_atom_moves=ensemble.atom_moves(), n_adjust=mc_params.n_adjust, mc_sample=mc_params.mc_sample,
              n_bin=mc_params.n_bin, min_acc=mc_params.min_acc, max_acc=mc_params.max_acc, seed=seed,
              device=device, rng_mode=rng_mode, lanes_per_replica=lanes_per_replica)
    if isinstance(ensemble, NPT):
        kw.update(n_volume_moves=ensemble.n_volume_moves, pressure=ensemble.pressure, separated_volume=ensemble.separated_volume)
    if potential.kind in (abi.POT_ELJ, abi.POT_ELJ_EVEN):
        kw["coeff"] = tuple(potential.coeff)
    elif potential.kind == abi.POT_ELJ_B:
        kw["coeff"] = tuple(potential.coeff_c)
        kw["eljb_ab"] = (tuple(potential.coeff_a), tuple(potential.coeff_b))
    elif potential.kind == abi.POT_LUT:
        kw["lut"] = potential
    elif potential.kind == abi.POT_EAM:
        kw["eam"] = (potential.n, potential.m, potential.ean, potential.eCam)
    elif potential.kind == abi.POT_RUNNER2:
        kw["runner2"] = potential.params
        kw["n_atom_swaps"] = ensemble.n_atom_swaps
    else:
        kw["runner"] = potential.params
    return SweepConfig(**kw)


def initial_arrays(mc_params: MCParams, start_config, n_ladders: int):
    """initialisation (src/Initialization.jl:79-101): slot i starts from start_config[mod1(i, len)]."""
    cfgs = start_config if isinstance(start_config, (list, tuple)) else (start_config,)
    T, N = mc_params.n_traj, mc_params.n_atoms
    pos = np.zeros((n_ladders, T, N, 3))
    box = np.zeros((n_ladders, T, 2))
    for i in range(T):
        c = cfgs[i % len(cfgs)]
        if len(c) != N:
            raise ValueError("number of atoms and positions not the same - check starting config")
        pos[:, i] = c.positions
        box[:, i] = c.boundary_condition.box()
    return pos, box


def recentre(positions):
    """recentre! src/Configurations.jl:79-85: subtract the centre (sum / n, coordinate by coordinate)"""
    p = np.asarray(positions, dtype=np.float64)
    c = np.array([sum(p[:, k].tolist()) / len(p) for k in range(3)])
    return p - c


def check_e_bounds(energy, ebounds):
    """src/MCRun.jl:258-265 (note the elseif)"""
    if energy < ebounds[0]:
        ebounds[0] = energy
    elif energy > ebounds[1]:
        ebounds[1] = energy
    return ebounds


def histogram_bounds(mc_params: MCParams, en_tot_ladder: np.ndarray, bc):
    """Second half of equilibration_cycle! (src/MCRun.jl:293-310) + initialise_histograms!
    (src/Sampling.jl:141-223) for ONE ladder: returns (en_min, en_max, v_min, v_max)."""
    ebounds = [100.0, -100.0]
    for _ in range(mc_params.eq_cycles // 2 + 1, mc_params.eq_cycles + 1):
        for e in en_tot_ladder:
            ebounds = check_e_bounds(float(e), ebounds)
    v_min = v_max = 0.0
    if not isinstance(bc, SphericalBC):
        v0 = get_volume(bc)
        v_min, v_max = v0 * 0.8, v0 * 2.0
    return ebounds[0], ebounds[1], v_min, v_max


def finalise_results(mc_params: MCParams, betas, ham, counters, results: Output) -> Output:
    """src/Sampling.jl:377-399 for one ladder (ham [n_traj][2], counters [n_traj][NCOUNT])."""
    n_sample = mc_params.mc_cycles / mc_params.mc_sample
    en_avg = ham[:, 0] / n_sample
    en2_avg = ham[:, 1] / n_sample
    results.en_avg = en_avg
    results.heat_cap = (en2_avg - en_avg ** 2) * np.asarray(betas) ** 2
    results.count_stat_atom = counters[:, abi.C_ATOM_TOT] / (mc_params.n_atoms * mc_params.mc_cycles)
    with np.errstate(divide="ignore", invalid="ignore"):
        results.count_stat_exc = counters[:, abi.C_EXC_ACC] / counters[:, abi.C_EXC_ATT]
    return results


def rdf_delta_r2(bc, n_bin: int) -> float:
    """delta_r2 of initialise_histograms! (src/Sampling.jl:151,185,213-214,240-241): rdf bins of the squared distance."""
    if isinstance(bc, SphericalBC):
        return 4 * bc.radius2 / n_bin / 5
    if isinstance(bc, CubicBC):
        return 3 / 4 * bc.box_length ** 2 / n_bin / 5
    if isinstance(bc, RhombicBC):
        return (3 / 8 * bc.box_length ** 2 + 1 / 4 * bc.box_height ** 2) / n_bin / 5
    return (1 / 2 * bc.box_length ** 2 + 1 / 4 * bc.box_height ** 2) / n_bin / 5


def _ladder_dir(save_dir, ladder: int, n_ladders: int):
    import pathlib
    return pathlib.Path(save_dir) if n_ladders == 1 else pathlib.Path(save_dir) / f"ladder{ladder}"


def ptmc_run(mc_params: MCParams, temp: TempGrid, start_config, potential, ensemble, *, n_ladders: int = 1,
             seed: int = 1234, device: int = 0, engine_cls=Engine, engine_kwargs=None, chunk: int = 100000,
             shared_bounds: bool = False, lanes_per_replica: int = 0, rdfsave: bool = False, save=False,
             save_dir="./checkpoint", restart: bool = False, saveconfigs=False, configsname: str = "configuration"):
    """ptmc_run! (src/MCRun.jl:363-432, restart method :434-488) for `n_ladders` independent ladders batched on one device.

    Returns (states, results): `states` is the dict of `Engine.download_states()` (per ladder, per slot),
    `results` a list of one `Output` per ladder.  With shared_bounds=True all ladders bin into the same
    [en_min, en_max] (the min/max over every ladder) so that their histograms can be summed.

    rdfsave: radial distribution histograms every 1000 cycles (src/Sampling.jl:282-316), `results[l].rdf`.
    save: False or a cycle count: every `save` cycles (and at the end) the reference's checkpoint files are written to
    `save_dir` (one sub-directory `ladder<l>` per ladder when n_ladders > 1), readsave.py / src/ReadSave.jl.
    saveconfigs: False or a cycle count: every `saveconfigs` cycles the configurations are written as
    `<save_dir>/<configsname><i>T<k>.xyz` (save_configs(mc_states, string(configsname, i)), src/MCRun.jl:413-415).
    restart: continue from the checkpoint in `save_dir` (configurations, step sizes, total counters, histograms, cycle
    index); equilibration is skipped (src/MCRun.jl:337-339)."""
    from . import readsave
    cfg = sweep_config(mc_params, temp, start_config, potential, ensemble, n_ladders=n_ladders, seed=seed,
                       device=device, lanes_per_replica=lanes_per_replica)
    eng = engine_cls(cfg, **(engine_kwargs or {}))
    cfgs = start_config if isinstance(start_config, (list, tuple)) else (start_config,)
    bc0 = cfgs[0].boundary_condition
    npt = ensemble.kind == abi.ENS_NPT
    nb2 = mc_params.n_bin + 2
    try:
        start_counter = 1
        old_hist = old_rdf = None
        if restart:
            # rebuild_states (src/ReadSave.jl:437-483): every ladder resumes from its own directory
            T, N = mc_params.n_traj, mc_params.n_atoms
            pos = np.empty((n_ladders, T, N, 3)); box = np.empty((n_ladders, T, 2)); md = np.full((n_ladders, T, 4), 0.1)
            counters = np.zeros((n_ladders, T, abi.PTMC_NCOUNT), dtype=np.int64)
            bounds = np.zeros((n_ladders, 4)); ck_delta_r2 = None
            old_hist = np.zeros((n_ladders, T, nb2, nb2) if npt else (n_ladders, T, nb2))
            old_rdf = np.zeros((n_ladders, T, 5 * mc_params.n_bin))
            for l in range(n_ladders):
                cf, m, ca, cv, res, rdf, index = readsave.rebuild_states(T, _ladder_dir(save_dir, l, n_ladders), npt)
                for t in range(T):
                    # recentre!(state.config) of every rebuilt state (src/Initialization.jl:115-117, src/Configurations.jl:79-85)
                    pos[l, t] = recentre(cf[t].positions); box[l, t] = cf[t].boundary_condition.box(); md[l, t, :3] = m[t]
                counters[l, :, abi.C_ATOM_TOT] = ca; counters[l, :, abi.C_VOL_TOT] = cv
                bounds[l] = (res.en_min, res.en_max, res.v_min, res.v_max)
                ck_delta_r2 = res.delta_r2                 # the checkpointed bin width of the rdf (hist_info.data)
                old_hist[l] = res.ev_histogram if npt else res.en_histogram
                if rdf is not None:
                    old_rdf[l] = rdf
                start_counter = index                      # start_counter = Int(readdlm("./checkpoint/index.txt")[1]): cycle `index` runs again (:108)
            eng.upload_states(pos, box, md)
            eng.upload_counters(counters, None)
            eng.seed_job(seed + start_counter)      # the reference restarts with whatever its task-local generator holds; hashed: no stream is reused
        else:
            pos, box = initial_arrays(mc_params, start_config, n_ladders)
            eng.upload_states(pos, box)
            eng.seed_job(seed)
            # equilibration (src/MCRun.jl:282-313): MC for eq_cycles ÷ 2 cycles only, cycle index from 1
            n_eq = mc_params.eq_cycles // 2
            if n_eq > 0:
                eng.run_cycles(1, n_eq, False)
            en = eng.download_states(("en_tot",))["en_tot"]
            bounds = np.array([histogram_bounds(mc_params, en[l], bc0) for l in range(n_ladders)])
            if shared_bounds:
                bounds[:, 0] = bounds[:, 0].min()
                bounds[:, 1] = bounds[:, 1].max()
            eng.reset_counters()
        eng.set_hist_bounds(bounds[:, 0].copy(), bounds[:, 1].copy(), bounds[:, 2].copy(), bounds[:, 3].copy())
        delta_r2 = ck_delta_r2 if (restart and ck_delta_r2) else rdf_delta_r2(bc0, mc_params.n_bin)
        if rdfsave:
            eng.rdf_enable(delta_r2)

        def make_results(states, hist, rdf):
            out = []
            for l in range(n_ladders):
                r = Output(mc_params.n_bin, en_min=bounds[l, 0], en_max=bounds[l, 1], v_min=bounds[l, 2], v_max=bounds[l, 3])
                r.delta_en_hist = (r.en_max - r.en_min) / (r.n_bin - 1)
                r.delta_r2 = delta_r2
                hl = hist[l] + (old_hist[l] if old_hist is not None else 0)
                if npt:
                    r.delta_v_hist = (r.v_max - r.v_min) / r.n_bin
                    r.ev_histogram = hl
                else:
                    r.en_histogram = hl
                if rdf is not None:
                    r.rdf = rdf[l] + (old_rdf[l] if old_rdf is not None else 0)
                r.max_displ = states["max_displ"][l]
                out.append(r)
            return out

        def write_checkpoint(i):
            states = eng.download_states()
            res = make_results(states, eng.download_hist(), eng.download_rdf() if rdfsave else None)
            for l in range(n_ladders):
                readsave.checkpoint(i, states, res[l], ensemble, bc0.kind, ladder=l, rdf=res[l].rdf if rdfsave else None,
                                    directory=_ladder_dir(save_dir, l, n_ladders))

        if save and restart:
            res0 = make_results({"max_displ": np.zeros((n_ladders, mc_params.n_traj, 4))}, np.zeros_like(old_hist), None)
            for l in range(n_ladders):
                readsave.save_histparams(res0[l], _ladder_dir(save_dir, l, n_ladders))   # the restart method rewrites hist_info.data (src/MCRun.jl:452-454)
        if save and not restart:
            res0 = make_results({"max_displ": np.zeros((n_ladders, mc_params.n_traj, 4))},
                                np.zeros((n_ladders,) + ((mc_params.n_traj, nb2, nb2) if npt else (mc_params.n_traj, nb2))), None)
            for l in range(n_ladders):
                d = _ladder_dir(save_dir, l, n_ladders)
                readsave.save_init(potential, ensemble, mc_params, temp, d)       # src/MCRun.jl:378-380
                readsave.save_histparams(res0[l], d)                              # :390-392
        # main loop (src/MCRun.jl:397-421), chunked; launches end where the reference touches the states: checkpoints
        # (rem(i, save) == 0) and rdf samples (rem(i, 1000) == 0 on sampled cycles, src/Sampling.jl:343-347)
        i = start_counter
        last_saved = None
        while i <= mc_params.mc_cycles:
            stop = min(i + chunk - 1, mc_params.mc_cycles)
            if save:
                stop = min(stop, (i + save - 1) // save * save)
            if saveconfigs:
                stop = min(stop, (i + saveconfigs - 1) // saveconfigs * saveconfigs)
            if rdfsave:
                stop = min(stop, (i + 999) // 1000 * 1000)
            eng.run_cycles(i, stop - i + 1, True)
            if rdfsave and stop % 1000 == 0 and stop % mc_params.mc_sample == 0:
                eng.rdf_sample()
            if save and stop % save == 0:
                write_checkpoint(stop)
                last_saved = stop
            if saveconfigs and stop % saveconfigs == 0:
                snap = eng.download_states()
                for l in range(n_ladders):
                    d = _ladder_dir(save_dir, l, n_ladders)
                    d.mkdir(parents=True, exist_ok=True)
                    readsave.save_configs(snap, l, bc0.kind, d, filename="%s%d" % (configsname, stop))
            i = stop + 1
        if save and last_saved != mc_params.mc_cycles:
            write_checkpoint(mc_params.mc_cycles)                                 # :424-427 (also after a restart, where the reference leaves it out, :434-481)
        eng.sync()
        states = eng.download_states()
        results = make_results(states, eng.download_hist(), eng.download_rdf() if rdfsave else None)
        for l in range(n_ladders):
            finalise_results(mc_params, temp.beta_grid, states["ham"][l], states["counters"][l], results[l])
        return states, results
    finally:
        eng.close()


# ------------------------------------------------------------------------- reference input decks --
def _geom(name):
    return np.array([[float(x) for x in row] for row in json.loads((DATA_DIR / f"{name}.json").read_text())])


ELJ_NE_COEFF = (-10.5097942564988, 989.725135614556, -101383.865938807, 3918846.12841668,
                -56234083.4334278, 288738837.441765)  # scripts/ne13.jl:39-47, ne55.jl:39-46, pbc216.jl:45-52


def deck_ne13(n_traj=32, cycles=1000):
    """examples/ne13.jl:24-70 (BASELINE C1: 32 temperatures)"""
    a2b = 1.8897259886
    pos = _geom("ne13") * a2b
    start = Config(pos, SphericalBC(radius=5.32 * a2b))
    return MCParams(cycles, n_traj, 13), TempGrid(4.0, 16.0, n_traj), start, ELJPotentialEven(ELJ_NE_COEFF), NVT(13)


# scripts/ne13_B.jl:40-44 (also ne38_B.jl, ne51_B.jl, ne55.jl): neon in a magnetic field
ELJB_NE_A = (0.0005742, -0.4032, -0.2101, -0.0595, 0.0606, 0.1608)
ELJB_NE_B = (-0.01336, -0.02005, -0.1051, -0.1268, -0.1405, -0.1751)
ELJB_NE_C = (-0.1132, -1.5012, 35.6955, -268.7494, 729.7605, -583.4203)


def deck_ne13_B(n_traj=16, cycles=1000):
    """scripts/ne13_B.jl: Ne13 with ELJPotentialB, NVT, 16 temperatures 9-16 K; the script works in Angstrom
    (positions :46-60 unconverted :64-66, SphericalBC(radius=5.32) :72)."""
    mcp, _, _, _, ens = deck_ne13(n_traj, cycles)
    start = Config(_geom("ne13"), SphericalBC(radius=5.32))
    return mcp, TempGrid(9.0, 16.0, n_traj), start, ELJPotentialB(ELJB_NE_A, ELJB_NE_B, ELJB_NE_C), ens


def deck_ne13_lookup(table, n_traj=25, cycles=1000):
    """scripts/ne13-lookup.jl: Ne13 with a LookupTablePotential (`table`: a file in the reference's format, e.g.
    scripts/lookup-tables/LookupTable_Neon_B0.0_MP2.txt, or a LookupTablePotential), NVT, 25 temperatures 4-16 K, Angstrom."""
    mcp, _, _, _, ens = deck_ne13(n_traj, cycles)
    pot = table if isinstance(table, LookupTablePotential) else LookupTablePotential(table)
    return mcp, TempGrid(4.0, 16.0, n_traj), Config(_geom("ne13"), SphericalBC(radius=5.32)), pot, ens


def deck_ne55(n_traj=24, cycles=1000):
    """scripts/ne55.jl:12-133 (BASELINE C2; positions used as given)"""
    a2b = 1.8897259886
    start = Config(_geom("ne55"), SphericalBC(radius=14 * a2b))
    return MCParams(cycles, n_traj, 55), TempGrid(6.0, 19.0, n_traj), start, ELJPotentialEven(ELJ_NE_COEFF), NVT(55)


def cu55_geometry():
    nmtobohr = 18.8973
    return (0.36258 * nmtobohr) * _geom("ico55")  # scripts/cu55.jl:222-224


def deck_cu55_eam(n_traj=28, cycles=1000):
    """scripts/cu55.jl:14-53,222-234 (BASELINE C3)"""
    evtohartree, nmtobohr, a2b = 0.0367493, 18.8973, 1.8897259886
    pot = EmbeddedAtomPotential(8.482, 4.692, evtohartree * 0.0370, 27.561, 0.25 * nmtobohr)
    start = Config(cu55_geometry(), SphericalBC(radius=14 * a2b))
    return MCParams(cycles, n_traj, 55), TempGrid(400.0, 1200.0, n_traj), start, pot, NVT(55)


def deck_cu38_eam(n_traj=36, cycles=1000):
    """scripts/cu38.jl:14-46,155-208 (one of the reference's script tests): Cu38, EmbeddedAtomPotential, NVT, 36 temperatures
    150-1200 K; positions :161-200 are used as given, SphericalBC(radius = 6.0 Angstrom in bohr) :205."""
    ev2h, nm2b = 0.0367493, 18.8973
    pot = EmbeddedAtomPotential(8.482, 4.692, ev2h * 0.0370, 27.561, 0.25 * nm2b)
    start = Config(_geom("cu38"), SphericalBC(radius=6.0 * 1.8897259886))
    return MCParams(cycles, n_traj, 38), TempGrid(150.0, 1200.0, n_traj), start, pot, NVT(38)


def deck_pbc32(n_traj=24, cycles=1000):
    """scripts/pbc32-short.jl:9-25,46-47,89-91 (one of the reference's script tests): 32-atom fcc block, CubicBC, NPT,
    ELJPotentialEven, 24 temperatures 10-25 K; the script's pressure conversion divides by AtoBohr^2."""
    a2b = 1.8897261259077824
    start = Config(_geom("ne32") * a2b, CubicBC(8.7674 * a2b))
    ens = NPT(32, 101325 * 2.2937122783969076e-13 / a2b ** 2, False)
    return MCParams(cycles, n_traj, 32), TempGrid(10.0, 25.0, n_traj), start, ELJPotentialEven(ELJ_NE_COEFF), ens


def deck_pbc216(n_traj=16, cycles=100):
    """scripts/pbc216.jl:14-70,517-529 (BASELINE C4; the LookupTable lines 62-63 are dropped)"""
    a2b = 1.8897261259077824
    fcc = _geom("ne216_fcc") * a2b * 0.85
    hcp = _geom("ne216_hcp") * a2b * 0.85
    bc = RhombicBC(18.79489824 * a2b * 0.85, 15.34597014 * a2b * 0.85)
    ens = NPT(216, 50000000000 * 2.2937122783969076e-13 / (a2b * a2b * a2b), False)
    return (MCParams(cycles, n_traj, 216), TempGrid(1600.0, 2100.0, n_traj), [Config(fcc, bc), Config(hcp, bc)],
            ELJPotentialEven(ELJ_NE_COEFF), ens)


# scripts/cu55.jl:55-100: radial (type, eta) rows and angular (eta, lambda, zeta) rows of the Cu network
CU_RADIAL = ((11, 0.001), (10, 0.001), (11, 0.020), (10, 0.020), (11, 0.035), (10, 0.035), (11, 0.100), (10, 0.100),
             (11, 0.400), (10, 0.400))
CU_ANGULAR = ((0.0001, 1, 1), (0.0001, -1, 2), (0.003, -1, 1), (0.003, -1, 2), (0.008, -1, 1), (0.008, -1, 2), (0.008, 1, 2),
              (0.015, 1, 1), (0.015, -1, 2), (0.015, -1, 4), (0.015, -1, 16), (0.025, -1, 1), (0.025, 1, 1), (0.025, 1, 2),
              (0.025, -1, 4), (0.025, -1, 16), (0.025, 1, 16), (0.045, 1, 1), (0.045, -1, 2), (0.045, -1, 4), (0.045, 1, 4),
              (0.045, 1, 16), (0.08, 1, 1), (0.08, -1, 2), (0.08, -1, 4), (0.08, 1, 4))
CU_ANGULAR_TYPES = (111, 110, 100)
RUNNER_R_CUT = 11.338


def load_scaling(path=None):
    """scaling.data rows (element, index, G_min, G_max, G_avg) -> [n][2] {G_max, G_min} as scripts/cu55.jl:104-112.
    `path`: a RuNNer scaling.data file; default: the Cu / Zn table of the reference's scripts (data/cu_nnp_inputs.npz)."""
    if path is None:
        rows = np.load(DATA_DIR / "cu_nnp_inputs.npz")["scaling_rows"]
    else:
        rows = np.loadtxt(path, usecols=(0, 1, 2, 3, 4), max_rows=176)   # a 2-number footer follows
    return np.stack([rows[:, 3], rows[:, 2]], axis=1)


def load_weights(path=None):
    """RuNNer weights.XXX.data (one parameter per line, SURVEY A.2 order); default: the Cu network of scripts/cu55.jl."""
    if path is None:
        return np.load(DATA_DIR / "cu_nnp_inputs.npz")["weights_cu"]
    return np.loadtxt(path).ravel()


def cu_runner_potential(boundary: float = 0.0, weights=None, scaling=None):
    """scripts/cu55.jl:55-153: 10 radial + 78 angular symmetry functions, network 88-20-20-1 (weights.029.data).
    The script passes boundary = 28 Angstrom, which rejects every move; the default here is the
    3-argument constructor's 0.0 (scripts/cu38.jl:151)."""
    g = load_scaling(scaling)[:88]
    radial = [(t, eta, RUNNER_R_CUT, g[k, 0], g[k, 1]) for k, (t, eta) in enumerate(CU_RADIAL)]
    angular = []
    n_index = 10
    for eta, lam, zeta in CU_ANGULAR:
        for t in CU_ANGULAR_TYPES:
            angular.append((eta, lam, zeta, RUNNER_R_CUT, t, g[n_index, 0], g[n_index, 1]))
            n_index += 1
    w = load_weights(weights)
    return RuNNerPotential([88, 20, 20, 1], [1, 2, 2, 1], w, radial, angular, boundary)


def deck_cu55_runner(n_traj=28, cycles=100, boundary: float = 0.0):
    """scripts/cu55.jl with `runnerpotential` (BASELINE C5)"""
    a2b = 1.8897259886
    start = Config(cu55_geometry(), SphericalBC(radius=14 * a2b))
    return (MCParams(cycles, n_traj, 55), TempGrid(400.0, 1200.0, n_traj), start, cu_runner_potential(boundary), NVT(55))


def load_weights_zn():
    """weights.030.data: the Zn network of scripts/brass55.jl:160-164 (data/cu_nnp_inputs.npz)"""
    return np.load(DATA_DIR / "cu_nnp_inputs.npz")["weights_zn"]


def brass_runner_potential(n1: int, n2: int, boundary: float = 0.0):
    """scripts/brass55.jl:37-167 (= test/potentialfile.jl): 5 radial + 26 angular two-species symmetry functions, scaling rows 1-88
    (Cu network) / 89-176 (Zn network), networks 88-20-20-1 from weights.029.data / weights.030.data."""
    g = load_scaling()
    ga, gb = g[:88], g[88:176]
    rad = [RadialType2a(eta, RUNNER_R_CUT, 2, (tuple(ga[2 * k]), tuple(ga[2 * k + 1])), (tuple(gb[2 * k]), tuple(gb[2 * k + 1])))
           for k, eta in enumerate((0.001, 0.020, 0.035, 0.100, 0.400))]
    ang = [AngularType3a(eta, lam, zeta, RUNNER_R_CUT, 2, tuple(tuple(ga[10 + 3 * k + q]) for q in range(3)),
                         tuple(tuple(gb[10 + 3 * k + q]) for q in range(3))) for k, (eta, lam, zeta) in enumerate(CU_ANGULAR)]
    nodes, acts = (88, 20, 20, 1), (1, 2, 2, 1)
    return RuNNerPotential2Atom(NeuralNetworkPotential(nodes, acts, load_weights()), NeuralNetworkPotential(nodes, acts, load_weights_zn()),
                                rad, ang, n1, n2, boundary)


def deck_brass55(n_traj=10, cycles=100, n_zn=5, n_swaps=2):
    """scripts/brass55.jl: Cu50Zn5 icosahedron, RuNNerPotential2Atom, NNVT([50, 5]; natomswaps = 2), 10 temperatures 250-900 K,
    SphericalBC(14 Angstrom); its ico_55 (:172-228) is used as given.  The script shuffles the 55 sites with Julia's `shuffle!`
    (Random.seed!(1234), :229-230) before the first 50 become Cu; without Julia that permutation is not reproducible here, so the
    deck takes the sites in a fixed pseudo-random order of its own (numpy default_rng(1234)) — which sites start as Zn is an input,
    not part of the path."""
    a2b = 1.8897259886
    pos = _geom("brass55")[np.random.default_rng(1234).permutation(55)]
    start = Config(pos, SphericalBC(radius=14 * a2b))
    ens = NNVT((55 - n_zn, n_zn), n_atom_swaps=n_swaps)
    return MCParams(cycles, n_traj, 55), TempGrid(250.0, 900.0, n_traj), start, brass_runner_potential(55 - n_zn, n_zn), ens
