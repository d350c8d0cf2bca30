"""Python face of the CPU ORACLE (oracle/ptmc_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.
`OracleEngine` exposes the same calls as ptmc_b200.engine.Engine so that a parity test drives the CUDA
library and the oracle with identical inputs."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess
import sys

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(_HERE.parent))
import ptmc_b200  # noqa: E402
from ptmc_b200 import abi  # noqa: E402
from ptmc_b200.engine import Engine, PtmcError, SweepConfig  # noqa: E402,F401

LIB_PATH = _HERE / "libptmc_oracle.so"
_lib = None
_vp, _dp, _ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
_u64p, _i64p, _u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)


def build(force: bool = False) -> pathlib.Path:
    srcs = [_HERE / "ptmc_oracle.c", _HERE / "runner_oracle.c", _HERE.parent / "include" / "ptmc_b200.h"]
    srcs = [s for s in srcs if s.exists()]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(str(LIB_PATH))
    sig = {
        "orc_create": (_vp, [C.POINTER(abi.Config)]),
        "orc_destroy": (None, [_vp]),
        "orc_set_threads": (None, [_vp, C.c_int]),
        "orc_upload_states": (C.c_int, [_vp, _dp, _dp, _dp]),
        "orc_upload_counters": (C.c_int, [_vp, _i64p, _dp]),
        "orc_reset_counters": (C.c_int, [_vp]),
        "orc_seed": (C.c_int, [_vp, C.c_uint64]),
        "orc_seed_job": (C.c_int, [_vp, C.c_uint64, C.c_int64]),
        "orc_rng_words": (C.c_int, [_vp]),
        "orc_set_rng_states": (C.c_int, [_vp, _u64p, _u64p]),
        "orc_get_rng_states": (C.c_int, [_vp, _u64p, _u64p]),
        "orc_set_replay": (C.c_int, [_vp, _u64p, C.c_int64, _u64p, C.c_int64]),
        "orc_set_hist_bounds": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int]),
        "orc_run_cycles": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int]),
        "orc_run_phase": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int, C.c_int]),
        "orc_config_doubles": (C.c_int64, [_vp]),
        "orc_export_config": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp]),
        "orc_import_config": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp]),
        "orc_trace_enable": (C.c_int, [_vp, C.c_int64]),
        "orc_trace_download": (C.c_int, [_vp, _ip, _dp, _u8p, _ip, _i64p]),
        "orc_download_states": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _i64p, _dp]),
        "orc_download_hist": (C.c_int, [_vp, _dp, _dp, C.c_int]),
        "orc_get_dist2_mat": (C.c_int, [_vp, C.c_int64, _dp]),
        "orc_probe_atom_move": (C.c_int, [_vp, C.c_int64, C.c_int, _dp, _dp]),
        "orc_get_runner_state": (C.c_int, [_vp, C.c_int64, _dp, _dp]),
        "orc_cutoff": (C.c_double, [C.c_double, C.c_double]),
        "orc_runner_forward_raw": (None, [_dp, C.c_int, _ip, _ip, _dp, C.c_int, _dp]),
        "orc_runner2_check": (C.c_double, [C.POINTER(abi.Runner2Desc), _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]),
        "orc_julia_seed": (None, [C.c_uint64, _u64p]),
        "orc_xoshiro_next": (C.c_uint64, [_u64p]),
        "orc_rand_from_state": (C.c_double, [_u64p]),
        "orc_range_from_state": (C.c_int64, [_u64p, C.c_uint64]),
        "orc_dimer_energy_elj": (C.c_double, [_dp, C.c_int, C.c_double]),
        "orc_dimer_energy_eljeven": (C.c_double, [_dp, C.c_int, C.c_double]),
        "orc_lrc_eljeven": (C.c_double, [_dp, C.c_int, C.c_int, C.c_double]),
        "orc_eam_ctor": (None, [C.c_double] * 5 + [_dp, _dp]),
        "orc_distance2": (C.c_double, [C.c_int, _dp, _dp, C.c_double, C.c_double]),
        "orc_get_tan": (C.c_double, [C.c_int, _dp, _dp, C.c_double, C.c_double]),
        "orc_dimer_energy_eljb": (C.c_double, [_dp, _dp, _dp, C.c_int, C.c_double, C.c_double]),
        "orc_lrc_eljb": (C.c_double, [C.c_int, C.c_double]),
        "orc_lrc_lut": (C.c_double, [C.c_double, C.c_int, C.c_double]),
        "orc_dimer_energy_lut": (C.c_double, [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
        "orc_get_volume": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "orc_get_r_cut": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "orc_tempgrid": (None, [C.c_double, C.c_double, C.c_int, _dp, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _arr(x):
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


# ---- scalar helpers used by the KAT tests -----------------------------------------------------------
def julia_seed(seed: int) -> np.ndarray:
    s = np.zeros(4, dtype=np.uint64)
    load().orc_julia_seed(seed, s.ctypes.data_as(_u64p))
    return s


def julia_rand(state: np.ndarray) -> float:
    return load().orc_rand_from_state(state.ctypes.data_as(_u64p))


def julia_rand_range(state: np.ndarray, n: int) -> int:
    return load().orc_range_from_state(state.ctypes.data_as(_u64p), n)


def xoshiro_next(state: np.ndarray) -> int:
    return load().orc_xoshiro_next(state.ctypes.data_as(_u64p))


def dimer_energy_elj(c, r2):
    a, p = _arr(c)
    return load().orc_dimer_energy_elj(p, len(a), r2)


def dimer_energy_eljeven(c, r2):
    a, p = _arr(c)
    return load().orc_dimer_energy_eljeven(p, len(a), r2)


def lrc_eljeven(c, n_atoms, r_cut):
    a, p = _arr(c)
    return load().orc_lrc_eljeven(p, len(a), n_atoms, r_cut)


def eam_ctor(n, m, eps, Cc, a):
    ean, ecam = C.c_double(), C.c_double()
    load().orc_eam_ctor(n, m, eps, Cc, a, C.byref(ean), C.byref(ecam))
    return ean.value, ecam.value


def distance2(bc_kind, a, b, L=0.0, h=0.0):
    aa, pa = _arr(a)
    bb, pb = _arr(b)
    return load().orc_distance2(bc_kind, pa, pb, L, h)


def get_tan(bc_kind, a, b, L=0.0, h=0.0):
    (_, pa), (_, pb) = _arr(a), _arr(b)
    return load().orc_get_tan(bc_kind, pa, pb, L, h)


def dimer_energy_eljb(ca, cb, cc, r2, tan):
    (a, pa), (b, pb), (c, pc) = _arr(ca), _arr(cb), _arr(cc)
    return load().orc_dimer_energy_eljb(pa, pb, pc, len(c), float(r2), float(tan))


def dimer_energy_lut(pot, r2, tan):
    """dimer_energy(::LookupTablePotential); pot: ptmc_b200.host.LookupTablePotential"""
    tab, pt = _arr(pot.table)
    return load().orc_dimer_energy_lut(pt, int(pot.l_dist), int(pot.l_angle), float(pot.start_dist), float(pot.d_dist), float(pot.d_angle),
                                       float(pot.c6coeff), float(r2), float(tan))


def lrc_eljb(n_atoms, r_cut):
    return load().orc_lrc_eljb(int(n_atoms), float(r_cut))


def get_volume(bc_kind, L, h=0.0):
    return load().orc_get_volume(bc_kind, L, h)


def get_r_cut(bc_kind, L, h=0.0):
    return load().orc_get_r_cut(bc_kind, L, h)


def tempgrid(ti, tf, n):
    t = np.zeros(n)
    b = np.zeros(n)
    load().orc_tempgrid(ti, tf, n, t.ctypes.data_as(_dp), b.ctypes.data_as(_dp))
    return t, b


def cutoff(r, r_cut):
    return load().orc_cutoff(r, r_cut)


def runner_forward(inp, nodes, act, params):
    """forward: inp [batch][n_in] -> eatom [batch]"""
    x, px = _arr(inp)
    nd = np.ascontiguousarray(nodes, dtype=np.int32)
    ac = np.ascontiguousarray(act, dtype=np.int32)
    pr, pp = _arr(params)
    e = np.zeros(x.shape[0])
    load().orc_runner_forward_raw(px, len(nd), nd.ctypes.data_as(_ip), ac.ctypes.data_as(_ip), pp, x.shape[0], e.ctypes.data_as(_dp))
    return e


def _runner2_check(r2, pos, what, a=0, b=0, trial=None):
    """orc_runner2_check: the oracle engine's two-species functions on one configuration; r2: engine.Runner2Params"""
    keep = []
    d = r2.to_c(keep)
    p, pp = _arr(pos)
    t, pt = _arr(np.zeros(3) if trial is None else trial)
    f12 = C.c_double()
    v = load().orc_runner2_check(C.byref(d), pp, what, int(a), int(b), pt, C.cast(C.byref(f12), _dp))
    return v, f12.value


def runner2_energy(r2, pos):
    """initial energy through total_symm_calc + two forward passes (test/test_potentials.jl:67-87) -> (E, f_matrix[1,2])"""
    return _runner2_check(r2, pos, 0)


def runner2_move_delta(r2, pos, atom, trial):
    """energy change of one atom move through calc_delta_matrix (test/test_states.jl:148-159); atom 0-based"""
    return _runner2_check(r2, pos, 1, atom, 0, trial)[0]


def runner2_swap_delta(r2, pos, a1, a2):
    """energy change of swapping atoms a1 (species 1) and a2 (species 2) through calc_swap_matrix; 0-based indices"""
    return _runner2_check(r2, pos, 2, a1, a2)[0]


class OracleEngine(Engine):
    """Engine-compatible driver of the CPU oracle."""

    prefix = "orc_"

    def __init__(self, cfg: SweepConfig, threads: int = 1):
        self._threads = threads
        super().__init__(cfg, lib=load())

    def _create(self):
        h = self.lib.orc_create(C.byref(self._c))
        if not h:
            raise PtmcError(abi.ERR_ARG, "oracle does not support this configuration")
        self._h = C.c_void_p(h)
        self.lib.orc_set_threads(self._h, self._threads)

    def last_error(self) -> str:
        return "oracle"

    def sync(self):
        return None

    def set_threads(self, n: int):
        self.lib.orc_set_threads(self._h, n)

    def dist2_mat(self, ladder: int, slot: int) -> np.ndarray:
        n = self.cfg.n_atoms
        out = np.zeros((n, n))
        self.lib.orc_get_dist2_mat(self._h, ladder * self.cfg.n_traj + slot, out.ctypes.data_as(_dp))
        return out

    # update_rdf! (src/Sampling.jl:301-316) straight from the oracle's dist2_mat, as the reference does
    def rdf_enable(self, delta_r2: float):
        L, T = self.shape
        self._delta_r2 = float(delta_r2)
        self._rdf = np.zeros((L, T, 5 * self.cfg.n_bin))

    def rdf_sample(self):
        L, T = self.shape
        nrb = 5 * self.cfg.n_bin
        for l in range(L):
            for t in range(T):
                d2 = self.dist2_mat(l, t)
                idx = np.floor(d2[np.tril_indices(self.cfg.n_atoms)] / self._delta_r2).astype(np.int64)   # for i, for k in 1:i
                idx = idx[(idx != 0) & (idx <= nrb)]
                np.add.at(self._rdf[l, t], idx - 1, 1.0)

    def download_rdf(self):
        return self._rdf.copy()

    def runner_state(self, ladder: int, slot: int):
        """(g_matrix [atom][n_sym], en_atom [atom]) of one replica (RuNNer)"""
        n = self.cfg.n_atoms
        ns = self.cfg.runner2.n_sym if self.cfg.runner2 is not None else len(self.cfg.runner.rad_eta) + len(self.cfg.runner.ang_eta)
        g, e = np.zeros((n, ns)), np.zeros(n)
        rc = self.lib.orc_get_runner_state(self._h, ladder * self.cfg.n_traj + slot, g.ctypes.data_as(_dp), e.ctypes.data_as(_dp))
        if rc:
            raise PtmcError(rc, "not a RuNNer handle")
        return g, e

    def probe_atom_move(self, ladder: int, slot: int, index1: int):
        trial = np.zeros(3)
        de = C.c_double()
        self.lib.orc_probe_atom_move(self._h, ladder * self.cfg.n_traj + slot, index1, trial.ctypes.data_as(_dp), C.byref(de))
        return trial, de.value
