"""Multi-GPU host logic: independent ladders are sharded across ranks (one process per GPU), no data-path
collective; the only exchange is one all-reduce (sum) of the ladder-summed histograms at the end
(SURVEY §8e).  Random streams are indexed by the GLOBAL ladder number, so a job gives bit-identical
per-ladder results for any number of GPUs.

Works with any torch.distributed backend: NCCL on the B200 box, gloo in the CPU tests."""
from __future__ import annotations

import hashlib

import numpy as np

from . import abi


def ladder_shard(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of ladders owned by `rank`: (first, count); the first total % world ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def julia_seed_state(seed: int) -> np.ndarray:
    """Random.seed!(seed) -> Xoshiro256++ state: SHA-256 of the little-endian UInt32 words."""
    seed &= (1 << 64) - 1
    msg = seed.to_bytes(8, "little")[: 8 if seed >> 32 else 4]
    return np.frombuffer(hashlib.sha256(msg).digest(), dtype="<u8").copy()


def global_stream_states(seed: int, first_ladder: int, n_ladders: int, total_ladders: int, n_traj: int, words: int):
    """Generator states of ladders [first_ladder, first_ladder + n_ladders) of a job of `total_ladders` ladders,
    following the rule of ptmc_seed (include/ptmc_b200.h) with GLOBAL ladder indices."""
    w = words // 4
    slot = np.empty((n_ladders, n_traj, w, 4), dtype=np.uint64)
    lad = np.empty((n_ladders, 4), dtype=np.uint64)
    for l in range(n_ladders):
        gl = first_ladder + l
        for t in range(n_traj):
            for k in range(w):
                slot[l, t, k] = julia_seed_state(seed + (gl * n_traj + t) * w + k)
        lad[l] = julia_seed_state(seed + total_ladders * n_traj * w + gl)
    return slot.reshape(n_ladders, n_traj, words), lad


def seed_global(engine, seed: int, first_ladder: int, total_ladders: int):
    L, T = engine.shape
    slot, lad = global_stream_states(seed, first_ladder, L, total_ladders, T, engine.rng_words())
    engine.set_rng_states(slot, lad)


def allreduce_histogram(hist, dist=None, device=None):
    """Sum a ladder-reduced histogram over all ranks.  `hist` is a numpy array (host path, any backend) or a
    torch tensor that already lives on the rank's device (NCCL path, in place).  Counts are integers held in
    float64 / int64, so the sum is exact."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return hist
    if isinstance(hist, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(hist))
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def run_sharded(mc_params, temp, start_config, potential, ensemble, *, total_ladders: int, seed: int = 1234,
                dist=None, engine_cls=None, device: int = 0, bounds=None, eq_cycles: int = 0, engine_kwargs=None):
    """ptmc_run!-style driver for a job of `total_ladders` ladders spread over the ranks of `dist`.
    Every rank returns (per-ladder states of its shard, job-wide ladder-summed histogram)."""
    from .engine import Engine
    from .host import initial_arrays, sweep_config
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    first, count = ladder_shard(total_ladders, rank, world)
    cfg = sweep_config(mc_params, temp, start_config, potential, ensemble, n_ladders=count, seed=seed, device=device)
    eng = (engine_cls or Engine)(cfg, **(engine_kwargs or {}))
    try:
        pos, box = initial_arrays(mc_params, start_config, count)
        eng.upload_states(pos, box)
        seed_global(eng, seed, first, total_ladders)
        if eq_cycles:
            eng.run_cycles(1, eq_cycles, False)
            eng.reset_counters()
        eng.set_hist_bounds(*bounds)
        eng.run_cycles(1, mc_params.mc_cycles, True)
        eng.sync()
        states = eng.download_states()
        hist = allreduce_histogram(eng.download_hist(reduce_ladders=True), dist)
        return states, hist, (first, count)
    finally:
        eng.close()
