"""ptmc_b200 — B200-native (sm_100a) sweep hot path of ParallelTemperingMonteCarlo.jl.

Layout: csrc/ (CUDA kernels + the C ABI of include/ptmc_b200.h), lib/ (built libptmc_b200.so),
abi.py + engine.py (ctypes driver), host.py (mirror of the reference's public API for this path),
julia/PTMCB200.jl (the ccall binding for the reference package), data/ (input decks of the BASELINE configs).
"""
from . import abi  # noqa: F401
from .engine import Engine, PtmcError, RunnerParams, SweepConfig  # noqa: F401
from .host import (  # noqa: F401
    Config, CubicBC, ELJPotential, ELJPotentialB, ELJPotentialEven, EmbeddedAtomPotential, LookupTablePotential, MCParams, NNVT, NPT, NVT, Output,
    RectangularBC, RhombicBC, RuNNerPotential, RuNNerPotential2Atom, SphericalBC, TempGrid, get_volume, kB, ptmc_run,
)
