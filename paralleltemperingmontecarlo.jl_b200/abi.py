"""ctypes mirror of include/ptmc_b200.h and the loader of libptmc_b200.so.

There is no CPU fallback: `load_library()` raises if the CUDA library has not been built, and
`ptmc_create` fails when no sm_100 device is present."""
from __future__ import annotations

import ctypes as C
import pathlib

PKG_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libptmc_b200.so"

PTMC_ABI_VERSION = 4
PTMC_MAX_COEFF = 16
PTMC_NCOUNT = 10

POT_ELJ, POT_ELJ_EVEN, POT_EAM, POT_RUNNER, POT_ELJ_B, POT_LUT, POT_RUNNER2 = 0, 1, 2, 3, 4, 5, 6
BC_SPHERICAL, BC_CUBIC, BC_RHOMBIC, BC_RECTANGULAR = 0, 1, 2, 3
ENS_NVT, ENS_NPT, ENS_NNVT = 0, 1, 2
RNG_XOSHIRO, RNG_REPLAY, RNG_BATCH8 = 0, 1, 2
RNG_LANES = 8

(C_ATOM_TOT, C_ATOM_ADJ, C_VOL_TOT, C_VOL_ADJ, C_VOLXY_TOT, C_VOLXY_ADJ, C_VOLZ_TOT, C_VOLZ_ADJ,
 C_EXC_ATT, C_EXC_ACC) = range(10)

ERR_ARG, ERR_CUDA, ERR_STATE, ERR_SPHERE, ERR_REPLAY, ERR_TIES = -1, -2, -3, -4, -5, -6

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class RunnerDesc(C.Structure):
    _fields_ = [
        ("n_rad", C.c_int32), ("n_ang", C.c_int32),
        ("rad_eta", _dp), ("rad_type", _ip), ("rad_gmax_gmin", _dp),
        ("ang_eta", _dp), ("ang_lambda", _dp), ("ang_zeta", _dp), ("ang_type", _ip), ("ang_gmax_gmin", _dp),
        ("r_cut", C.c_double), ("boundary2", C.c_double),
        ("n_layers", C.c_int32), ("num_nodes", _ip), ("activation", _ip),
        ("n_params", C.c_int32), ("params", _dp),
    ]


class Runner2Desc(C.Structure):
    _fields_ = [
        ("n_rad", C.c_int32), ("n_ang", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32),
        ("rad_eta", _dp), ("ang_eta", _dp), ("ang_lambda", _dp), ("ang_zeta", _dp),
        ("gmax_gmin_1", _dp), ("gmax_gmin_2", _dp),
        ("r_cut", C.c_double), ("boundary2", C.c_double),
        ("n_layers", C.c_int32), ("num_nodes", _ip), ("activation", _ip),
        ("n_params", C.c_int32), ("params_1", _dp), ("params_2", _dp),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("n_ladders", C.c_int64),
        ("n_traj", C.c_int32), ("n_atoms", C.c_int32),
        ("ensemble", C.c_int32), ("n_atom_moves", C.c_int32), ("n_volume_moves", C.c_int32), ("bc_kind", C.c_int32),
        ("pressure", C.c_double),
        ("pot_kind", C.c_int32), ("n_coeff", C.c_int32),
        ("coeff", C.c_double * PTMC_MAX_COEFF),
        ("eam_n", C.c_double), ("eam_m", C.c_double), ("eam_ean", C.c_double), ("eam_ecam", C.c_double),
        ("runner", C.POINTER(RunnerDesc)),
        ("betas", _dp),
        ("n_adjust", C.c_int32), ("mc_sample", C.c_int32), ("n_bin", C.c_int32), ("rng_mode", C.c_int32),
        ("min_acc", C.c_double), ("max_acc", C.c_double),
        ("seed", C.c_uint64),
        ("lanes_per_replica", C.c_int32), ("separated_volume", C.c_int32),
        ("eljb_a", _dp), ("eljb_b", _dp),
        ("lut_table", _dp), ("lut_l_dist", C.c_int32), ("lut_l_angle", C.c_int32),
        ("lut_start_dist", C.c_double), ("lut_d_dist", C.c_double), ("lut_start_angle", C.c_double), ("lut_d_angle", C.c_double),
        ("lut_c6coeff", C.c_double),
        ("n_atom_swaps", C.c_int32), ("runner2", C.POINTER(Runner2Desc)),
    ]


# name -> (restype, argtypes); every symbol include/ptmc_b200.h declares
_vp = C.c_void_p
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
SYMBOLS = {
    "ptmc_abi_version": (C.c_int, []),
    "ptmc_last_error": (C.c_char_p, [_vp]),
    "ptmc_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "ptmc_destroy": (None, [_vp]),
    "ptmc_set_stream": (C.c_int, [_vp, _vp]),
    "ptmc_upload_states": (C.c_int, [_vp, _dp, _dp, _dp]),
    "ptmc_upload_counters": (C.c_int, [_vp, _i64p, _dp]),
    "ptmc_reset_counters": (C.c_int, [_vp]),
    "ptmc_seed": (C.c_int, [_vp, C.c_uint64]),
    "ptmc_seed_job": (C.c_int, [_vp, C.c_uint64, C.c_int64]),
    "ptmc_rng_words": (C.c_int, [_vp]),
    "ptmc_set_rng_states": (C.c_int, [_vp, _u64p, _u64p]),
    "ptmc_get_rng_states": (C.c_int, [_vp, _u64p, _u64p]),
    "ptmc_set_replay": (C.c_int, [_vp, _u64p, C.c_int64, _u64p, C.c_int64]),
    "ptmc_set_hist_bounds": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_int]),
    "ptmc_run_cycles": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int]),
    "ptmc_run_phase": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "ptmc_config_doubles": (C.c_int64, [_vp]),
    "ptmc_export_config": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp]),
    "ptmc_import_config": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp]),
    "ptmc_sync": (C.c_int, [_vp]),
    "ptmc_trace_enable": (C.c_int, [_vp, C.c_int64]),
    "ptmc_trace_download": (C.c_int, [_vp, _ip, _dp, _u8p, _ip, _i64p]),
    "ptmc_download_states": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _i64p, _dp]),
    "ptmc_download_hist": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "ptmc_reduced_hist_dev": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "ptmc_rdf_enable": (C.c_int, [_vp, C.c_double]),
    "ptmc_rdf_sample": (C.c_int, [_vp]),
    "ptmc_download_rdf": (C.c_int, [_vp, _dp]),
    "ptmc_launch_count": (C.c_int64, [_vp]),
    "ptmc_lanes_per_replica": (C.c_int, [_vp]),
    "ptmc_fp64_peak": (C.c_int, [C.c_int, C.c_int, _dp]),
    "ptmc_issue_probe": (C.c_int, [C.c_int, C.c_double, _dp]),
    "ptmc_debug_pair_energy": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp]),
    "ptmc_debug_runner_state": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp, _dp]),
    "forward": (C.c_double, [_dp, _ip, _ip, _ip, _ip, _ip, _dp, _ip, _dp]),
}

_lib = None


def bind(lib, symbols=SYMBOLS, prefix_from="ptmc_", prefix_to="ptmc_"):
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name.replace(prefix_from, prefix_to, 1) if name.startswith(prefix_from) else name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path: pathlib.Path | None = None):
    """Load libptmc_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if missing.
    PTMC_B200_LIB selects another build of the same library (kernel tuning experiments)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    import os
    if path is None and os.environ.get("PTMC_B200_LIB"):
        p = pathlib.Path(os.environ["PTMC_B200_LIB"])
        if not p.is_absolute():
            p = LIB_PATH.parent / p
    else:
        p = pathlib.Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback for the sweep."
        )
    lib = bind(C.CDLL(str(p)))
    if lib.ptmc_abi_version() != PTMC_ABI_VERSION:
        raise RuntimeError("libptmc_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
