"""Thin numpy/ctypes driver over the C ABI of include/ptmc_b200.h.

`Engine` is what the Julia wrapper (julia/PTMCB200.jl) does with `ccall`, written in Python because the
build/test image has no Julia.  It owns no arithmetic: every number comes out of libptmc_b200.so.
The CPU oracle (oracle/oracle.py, test infrastructure) subclasses it with its own library so that parity
tests drive both through the same calls."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


class PtmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ptmc error {code}: {msg}")
        self.code = code


@dataclass
class RunnerParams:
    """RuNNerPotential payload (src/EnergyEvaluation.jl:1182-1218)."""
    rad_eta: np.ndarray
    rad_type: np.ndarray
    rad_gmax_gmin: np.ndarray
    ang_eta: np.ndarray
    ang_lambda: np.ndarray
    ang_zeta: np.ndarray
    ang_type: np.ndarray
    ang_gmax_gmin: np.ndarray
    r_cut: float
    boundary2: float
    num_nodes: np.ndarray
    activation: np.ndarray
    params: np.ndarray


@dataclass
class Runner2Params:
    """RuNNerPotential2Atom payload (src/EnergyEvaluation.jl:1287-1314): gmax_gmin_1 / _2 are [2 n_rad + 3 n_ang][2] {G_max, G_min}
    in g-matrix row order for species 1 / 2."""
    n1: int
    n2: int
    rad_eta: np.ndarray
    ang_eta: np.ndarray
    ang_lambda: np.ndarray
    ang_zeta: np.ndarray
    gmax_gmin_1: np.ndarray
    gmax_gmin_2: np.ndarray
    r_cut: float
    boundary2: float
    num_nodes: np.ndarray
    activation: np.ndarray
    params_1: np.ndarray
    params_2: np.ndarray

    @property
    def n_sym(self) -> int:
        return 2 * len(self.rad_eta) + 3 * len(self.ang_eta)

    def to_c(self, keep: list) -> "abi.Runner2Desc":
        d = abi.Runner2Desc()
        arrs = {}
        for name, dt in (("rad_eta", np.float64), ("ang_eta", np.float64), ("ang_lambda", np.float64), ("ang_zeta", np.float64),
                         ("gmax_gmin_1", np.float64), ("gmax_gmin_2", np.float64), ("num_nodes", np.int32), ("activation", np.int32),
                         ("params_1", np.float64), ("params_2", np.float64)):
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            arrs[name] = a
            setattr(d, name, a.ctypes.data_as(C.POINTER(C.c_double) if dt == np.float64 else C.POINTER(C.c_int32)))
        d.n_rad, d.n_ang, d.n1, d.n2 = len(arrs["rad_eta"]), len(arrs["ang_eta"]), int(self.n1), int(self.n2)
        if arrs["gmax_gmin_1"].shape != (self.n_sym, 2) or arrs["gmax_gmin_2"].shape != (self.n_sym, 2):
            raise ValueError("gmax_gmin_1 / gmax_gmin_2 must be [2 n_rad + 3 n_ang][2]")
        if arrs["params_1"].shape != arrs["params_2"].shape:
            raise ValueError("the two networks must share one topology")
        d.r_cut, d.boundary2 = float(self.r_cut), float(self.boundary2)
        d.n_layers, d.n_params = len(arrs["num_nodes"]), len(arrs["params_1"])
        keep += [arrs, d]
        return d


@dataclass
class SweepConfig:
    n_ladders: int
    n_traj: int
    n_atoms: int
    betas: np.ndarray
    pot_kind: int = abi.POT_ELJ_EVEN
    coeff: tuple = ()
    eam: tuple = (0.0, 0.0, 0.0, 0.0)  # n, m, ean, eCam
    eljb_ab: tuple | None = None       # ELJPotentialB: (coeff_a, coeff_b); coeff_c in `coeff`
    lut: object | None = None           # LookupTablePotential (host.LookupTablePotential)
    runner: RunnerParams | None = None
    runner2: Runner2Params | None = None
    n_atom_swaps: int = 0               # NNVT.n_atom_swaps
    bc_kind: int = abi.BC_SPHERICAL
    ensemble: int = abi.ENS_NVT
    n_atom_moves: int | None = None
    n_volume_moves: int = 0
    pressure: float = 0.0
    separated_volume: bool = False
    n_adjust: int = 100
    mc_sample: int = 1
    n_bin: int = 100
    min_acc: float = 0.4
    max_acc: float = 0.6
    rng_mode: int = abi.RNG_BATCH8
    seed: int = 1234
    device: int = 0
    lanes_per_replica: int = 0
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_steps(self) -> int:
        return (self.atom_moves + (self.n_volume_moves if self.ensemble == abi.ENS_NPT else 0)
                + (self.n_atom_swaps if self.ensemble == abi.ENS_NNVT else 0))

    @property
    def atom_moves(self) -> int:
        return self.n_atoms if self.n_atom_moves is None else self.n_atom_moves

    def to_c(self) -> abi.Config:
        c = abi.Config()
        c.abi_version = abi.PTMC_ABI_VERSION
        c.device = self.device
        c.n_ladders = self.n_ladders
        c.n_traj = self.n_traj
        c.n_atoms = self.n_atoms
        c.ensemble = self.ensemble
        c.n_atom_moves = self.atom_moves
        c.n_volume_moves = self.n_volume_moves if self.ensemble == abi.ENS_NPT else 0
        c.bc_kind = self.bc_kind
        c.pressure = self.pressure
        c.pot_kind = self.pot_kind
        if len(self.coeff) > abi.PTMC_MAX_COEFF:
            raise ValueError("too many ELJ coefficients")
        c.n_coeff = len(self.coeff)
        for i, v in enumerate(self.coeff):
            c.coeff[i] = float(v)
        c.eam_n, c.eam_m, c.eam_ean, c.eam_ecam = (float(x) for x in self.eam)
        betas = np.ascontiguousarray(self.betas, dtype=np.float64)
        if betas.shape != (self.n_traj,):
            raise ValueError("betas must have n_traj entries")
        self._keep = [betas]
        c.betas = betas.ctypes.data_as(C.POINTER(C.c_double))
        if self.eljb_ab is not None:
            ea, eb = (np.ascontiguousarray(v, dtype=np.float64) for v in self.eljb_ab)
            if ea.shape != (len(self.coeff),) or eb.shape != (len(self.coeff),):
                raise ValueError("ELJPotentialB needs coeff_a, coeff_b, coeff_c of equal length")
            self._keep += [ea, eb]
            c.eljb_a = ea.ctypes.data_as(C.POINTER(C.c_double))
            c.eljb_b = eb.ctypes.data_as(C.POINTER(C.c_double))
        if self.lut is not None:
            tab = np.ascontiguousarray(self.lut.table, dtype=np.float64)          # [l_angle][l_dist]
            if tab.shape != (self.lut.l_angle, self.lut.l_dist):
                raise ValueError("lookup table must be [l_angle][l_dist]")
            self._keep.append(tab)
            c.lut_table = tab.ctypes.data_as(C.POINTER(C.c_double))
            c.lut_l_dist, c.lut_l_angle = int(self.lut.l_dist), int(self.lut.l_angle)
            c.lut_start_dist, c.lut_d_dist = float(self.lut.start_dist), float(self.lut.d_dist)
            c.lut_start_angle, c.lut_d_angle = float(self.lut.start_angle), float(self.lut.d_angle)
            c.lut_c6coeff = float(self.lut.c6coeff)
        if self.runner is not None:
            r = self.runner
            d = abi.RunnerDesc()
            arrs = {}
            for name, dt in (("rad_eta", np.float64), ("rad_type", np.int32), ("rad_gmax_gmin", np.float64),
                             ("ang_eta", np.float64), ("ang_lambda", np.float64), ("ang_zeta", np.float64),
                             ("ang_type", np.int32), ("ang_gmax_gmin", np.float64), ("num_nodes", np.int32),
                             ("activation", np.int32), ("params", np.float64)):
                a = np.ascontiguousarray(getattr(r, name), dtype=dt)
                arrs[name] = a
                ptr_t = C.POINTER(C.c_double) if dt == np.float64 else C.POINTER(C.c_int32)
                setattr(d, name, a.ctypes.data_as(ptr_t))
            d.n_rad = len(arrs["rad_eta"])
            d.n_ang = len(arrs["ang_eta"])
            d.r_cut = r.r_cut
            d.boundary2 = r.boundary2
            d.n_layers = len(arrs["num_nodes"])
            d.n_params = len(arrs["params"])
            self._keep += [arrs, d]
            c.runner = C.pointer(d)
        if self.runner2 is not None:
            c.runner2 = C.pointer(self.runner2.to_c(self._keep))
        c.n_atom_swaps = self.n_atom_swaps if self.ensemble == abi.ENS_NNVT else 0
        c.n_adjust = self.n_adjust
        c.mc_sample = self.mc_sample
        c.n_bin = self.n_bin
        c.rng_mode = self.rng_mode
        c.min_acc = self.min_acc
        c.max_acc = self.max_acc
        c.seed = self.seed
        c.lanes_per_replica = self.lanes_per_replica
        c.separated_volume = int(bool(self.separated_volume)) if self.ensemble == abi.ENS_NPT else 0
        return c


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Engine:
    """One device-resident batch of PT ladders (n_ladders x n_traj replicas)."""

    prefix = "ptmc_"

    def __init__(self, cfg: SweepConfig, lib=None):
        self.cfg = cfg
        self.lib = lib if lib is not None else abi.load_library()
        self._c = cfg.to_c()
        self._h = C.c_void_p()
        self._keep = []
        self._create()

    # -- plumbing ------------------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _create(self):
        rc = self.lib.ptmc_create(C.byref(self._c), C.byref(self._h))
        if rc != 0:
            raise PtmcError(rc, (self.lib.ptmc_last_error(None) or b"").decode())

    def _check(self, rc):
        if rc != 0:
            raise PtmcError(rc, self.last_error())
        return rc

    def last_error(self) -> str:
        return (self.lib.ptmc_last_error(self._h) or b"").decode()

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        return (self.cfg.n_ladders, self.cfg.n_traj)

    # -- state in ------------------------------------------------------------------------------------
    def upload_states(self, pos, box, max_displ=None):
        L, T = self.shape
        pos = _f64(pos, (L, T, self.cfg.n_atoms, 3))
        box = _f64(box, (L, T, 2))
        md = None if max_displ is None else _f64(max_displ, (L, T, 4))
        self._check(self._fn("upload_states")(self._h, _p(pos), _p(box), _p(md)))

    def upload_counters(self, counters=None, ham=None):
        L, T = self.shape
        cn = None if counters is None else np.ascontiguousarray(counters, dtype=np.int64).reshape(L, T, abi.PTMC_NCOUNT)
        hm = None if ham is None else _f64(ham, (L, T, 2))
        self._check(self._fn("upload_counters")(self._h, _p(cn, C.c_int64), _p(hm)))

    def reset_counters(self):
        self._check(self._fn("reset_counters")(self._h))

    # -- random streams ------------------------------------------------------------------------------
    def seed(self, seed: int):
        self._check(self._fn("seed")(self._h, C.c_uint64(seed)))

    def seed_job(self, seed: int, first_ladder: int = 0):
        """hashed (seed, stream id) rule with GLOBAL ladder numbers: disjoint streams for different seeds, same per-ladder streams for
        any sharding (ptmc_seed_job)"""
        self._check(self._fn("seed_job")(self._h, C.c_uint64(seed), C.c_int64(first_ladder)))

    def rng_words(self) -> int:
        return int(self._fn("rng_words")(self._h))

    def set_rng_states(self, slot_states, ladder_states):
        L, T = self.shape
        s = np.ascontiguousarray(slot_states, dtype=np.uint64).reshape(L, T, self.rng_words())
        l = np.ascontiguousarray(ladder_states, dtype=np.uint64).reshape(L, 4)
        self._check(self._fn("set_rng_states")(self._h, _p(s, C.c_uint64), _p(l, C.c_uint64)))

    def get_rng_states(self):
        L, T = self.shape
        s = np.zeros((L, T, self.rng_words()), dtype=np.uint64)
        l = np.zeros((L, 4), dtype=np.uint64)
        self._check(self._fn("get_rng_states")(self._h, _p(s, C.c_uint64), _p(l, C.c_uint64)))
        return s, l

    def set_replay(self, slot_draws, ladder_draws):
        L, T = self.shape
        s = np.ascontiguousarray(slot_draws, dtype=np.uint64)
        l = np.ascontiguousarray(ladder_draws, dtype=np.uint64)
        if s.shape[:2] != (L, T) or l.shape[0] != L:
            raise ValueError("replay streams must be [n_ladders][n_traj][len] and [n_ladders][len]")
        self._keep = [s, l]  # the oracle reads them in place
        self._check(self._fn("set_replay")(self._h, _p(s, C.c_uint64), s.shape[2], _p(l, C.c_uint64), l.shape[1]))

    # -- histograms ----------------------------------------------------------------------------------
    def set_hist_bounds(self, en_min, en_max, v_min=None, v_max=None):
        per = int(np.ndim(en_min) > 0)
        n = self.cfg.n_ladders if per else 1
        a = [None if x is None else _f64(np.reshape(x, (n,)), (n,)) for x in (en_min, en_max, v_min, v_max)]
        self._check(self._fn("set_hist_bounds")(self._h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), per))

    # -- hot path ------------------------------------------------------------------------------------
    def run_cycles(self, first_idx: int, n_cycles: int, sample: bool):
        self._check(self._fn("run_cycles")(self._h, first_idx, n_cycles, int(bool(sample))))

    def run_phase(self, first_idx: int, n_cycles: int, sample: bool, phase: int):
        """mc_cycle! split at the exchange: phase 1 = mc_step! only, 2 = step-size adaptation + sampling only (multi.run_split_ladder)"""
        self._check(self._fn("run_phase")(self._h, int(first_idx), int(n_cycles), int(bool(sample)), int(phase)))

    def export_config(self, ladder: int, slot: int) -> np.ndarray:
        buf = np.zeros(int(self._fn("config_doubles")(self._h)))
        self._check(self._fn("export_config")(self._h, int(ladder), int(slot), _p(buf)))
        return buf

    def import_config(self, ladder: int, slot: int, buf):
        buf = _f64(buf, (int(self._fn("config_doubles")(self._h)),))
        self._check(self._fn("import_config")(self._h, int(ladder), int(slot), _p(buf)))

    def sync(self):
        self._check(self._fn("sync")(self._h))

    # -- trace ---------------------------------------------------------------------------------------
    def trace_enable(self, capacity_cycles: int):
        self._trace_cap = capacity_cycles
        self._check(self._fn("trace_enable")(self._h, capacity_cycles))

    def trace_download(self):
        L, T = self.shape
        S = self.cfg.n_steps
        cap = self._trace_cap
        idx = np.zeros((L, cap, T, S), dtype=np.int32)
        de = np.zeros((L, cap, T, S), dtype=np.float64)
        acc = np.zeros((L, cap, T, S), dtype=np.uint8)
        exc = np.zeros((L, cap, 3), dtype=np.int32)
        n = C.c_int64(0)
        self._check(self._fn("trace_download")(self._h, _p(idx, C.c_int32), _p(de), _p(acc, C.c_uint8), _p(exc, C.c_int32), C.byref(n)))
        k = n.value
        # the library writes compacted [L][k][...] arrays
        def cut(a, tail):
            return a.reshape(-1)[: L * k * int(np.prod(tail))].reshape((L, k) + tail).copy()
        return {"move_index": cut(idx, (T, S)), "delta_en": cut(de, (T, S)), "accepted": cut(acc, (T, S)),
                "exchange": cut(exc, (3,)), "n_cycles": k}

    # -- state out -----------------------------------------------------------------------------------
    def state_shapes(self):
        L, T = self.shape
        return {"pos": ((L, T, self.cfg.n_atoms, 3), np.float64), "box": ((L, T, 2), np.float64),
                "en_tot": ((L, T), np.float64), "max_displ": ((L, T, 4), np.float64),
                "counters": ((L, T, abi.PTMC_NCOUNT), np.int64), "ham": ((L, T, 2), np.float64)}

    def download_states(self, what=("pos", "box", "en_tot", "max_displ", "counters", "ham"), out=None):
        """Per-slot state.  `out` may hold caller-owned (e.g. pinned) arrays of the right shape to fill."""
        shapes = self.state_shapes()
        out = {} if out is None else {k: v for k, v in out.items() if k in what}
        for k in what:
            shp, dt = shapes[k]
            if k not in out:
                out[k] = np.zeros(shp, dtype=dt)
            elif out[k].shape != shp or out[k].dtype != dt or not out[k].flags.c_contiguous:
                raise ValueError(f"out[{k!r}] must be a C-contiguous {dt.__name__}{shp} array")
        self._check(self._fn("download_states")(
            self._h, _p(out.get("pos")), _p(out.get("box")), _p(out.get("en_tot")), _p(out.get("max_displ")),
            _p(out.get("counters"), C.c_int64), _p(out.get("ham"))))
        return out

    def download_hist(self, reduce_ladders: bool = False):
        L, T = self.shape
        nb2 = self.cfg.n_bin + 2
        lead = (T,) if reduce_ladders else (L, T)
        if self.cfg.ensemble == abi.ENS_NPT:
            ev = np.zeros(lead + (nb2, nb2))
            self._check(self._fn("download_hist")(self._h, None, _p(ev), int(reduce_ladders)))
            return ev
        en = np.zeros(lead + (nb2,))
        self._check(self._fn("download_hist")(self._h, _p(en), None, int(reduce_ladders)))
        return en

    # -- radial distribution function (src/Sampling.jl:282-316) ---------------------------------------
    def rdf_enable(self, delta_r2: float):
        self._check(self._fn("rdf_enable")(self._h, float(delta_r2)))

    def rdf_sample(self):
        self._check(self._fn("rdf_sample")(self._h))

    def download_rdf(self):
        L, T = self.shape
        out = np.zeros((L, T, 5 * self.cfg.n_bin))
        self._check(self._fn("download_rdf")(self._h, _p(out)))
        return out

    # -- device-side extras (CUDA library only) -------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._check(self._fn("set_stream")(self._h, C.c_void_p(cuda_stream)))

    def launch_count(self) -> int:
        return int(self._fn("launch_count")(self._h))

    def lanes_per_replica(self) -> int:
        return int(self._fn("lanes_per_replica")(self._h))

    def debug_pair_energy(self, d2_new, d2_old):
        a, b = _f64(d2_new), _f64(d2_old)
        en, eo = np.zeros_like(a), np.zeros_like(b)
        self._check(self._fn("debug_pair_energy")(self._h, a.size, _p(a), _p(b), _p(en), _p(eo)))
        return en, eo

    def runner_state(self, ladder: int, slot: int):
        """(g_matrix [atom][n_sym], en_atom [atom]) of the configuration held by one slot (RuNNer)"""
        n = self.cfg.n_atoms
        ns = self.cfg.runner2.n_sym if self.cfg.runner2 is not None else len(self.cfg.runner.rad_eta) + len(self.cfg.runner.ang_eta)
        g, e = np.zeros((n, ns)), np.zeros(n)
        self._check(self._fn("debug_runner_state")(self._h, ladder, slot, _p(g), _p(e)))
        return g, e

    def reduced_hist_dev(self):
        ptr = C.c_void_p()
        n = C.c_int64()
        self._check(self._fn("reduced_hist_dev")(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value
