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


def julia_seed128_state(lo: int, hi: int) -> np.ndarray:
    """Random.seed!(lo + hi * 2^64): SHA-256 of the little-endian UInt32 words of the integer, trailing zero words dropped."""
    msg = (lo & (2 ** 64 - 1)).to_bytes(8, "little") + (hi & (2 ** 64 - 1)).to_bytes(8, "little")
    n = 16
    while n > 4 and msg[n - 4:n] == b"\0\0\0\0":
        n -= 4
    return np.frombuffer(hashlib.sha256(msg[:n]).digest(), dtype="<u8").copy()


def job_stream_states(seed: int, first_ladder: int, n_ladders: int, n_traj: int, words: int):
    """The rule of ptmc_seed_job (include/ptmc_b200.h): stream = hash of (seed, stream id), ids from the GLOBAL ladder number."""
    w = words // 4
    slot = np.empty((n_ladders, n_traj, w, 4), dtype=np.uint64)
    lad = np.empty((n_ladders, 4), dtype=np.uint64)
    for l in range(n_ladders):
        gl = first_ladder + l
        for t in range(n_traj):
            for k in range(w):
                slot[l, t, k] = julia_seed128_state(seed, (gl * n_traj + t) * w + k + 1)
        lad[l] = julia_seed128_state(seed, (1 << 63) | gl)
    return slot.reshape(n_ladders, n_traj, words), lad


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
        eng.seed_job(seed, first)
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


# ------------------------------------------------------------------------- one ladder across several GPUs --
# SURVEY §8e, optional mode: the temperatures of every ladder are split into `world` contiguous blocks, one per rank.
# mc_step!, step-size adaptation and sampling stay on the device (ptmc_run_phase); the exchange step
# (src/MCRun.jl:203-205, src/Exchange.jl:132-155) needs the energies of two neighbouring slots that may live on two
# GPUs: all ranks all-gather the energies, take the identical decision from the shared driver-task stream, and the two
# owners swap the configurations (exc_trajectories! :115-124) point to point.  Results are bit-identical to the same job
# on one handle (streams are indexed globally; the acceptance exp() is evaluated on the host instead of the device).
_M64 = (1 << 64) - 1


def _xo_next(s):
    """Xoshiro256++ on a list of four Python ints."""
    rot = lambda x, k: ((x << k) | (x >> (64 - k))) & _M64
    res = (rot((s[0] + s[3]) & _M64, 23) + s[0]) & _M64
    t = (s[1] << 17) & _M64
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rot(s[3], 45)
    return res


def _jl_rand(s) -> float:
    return (_xo_next(s) >> 11) * 2.0 ** -53


def _jl_rand_range(s, n: int) -> int:
    """rand(1:n): Lemire's method with Julia's rejection rule."""
    x = _xo_next(s); m = x * n; lo = m & _M64
    if lo < n:
        t = ((1 << 64) - n) % n
        while lo < t:
            x = _xo_next(s); m = x * n; lo = m & _M64
    return (m >> 64) + 1


def _swap_with_peer(dist, peer: int, buf: np.ndarray, device):
    import torch
    send = torch.from_numpy(np.ascontiguousarray(buf))
    if device is not None:
        send = send.to(device)
    recv = torch.empty_like(send)
    for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]):
        r.wait()
    return recv.cpu().numpy()


def run_split_ladder(mc_params, temp, start_config, potential, ensemble, *, n_ladders: int, seed: int = 1234, dist=None,
                     engine_cls=None, device: int = 0, bounds=None, eq_cycles: int = 0, comm_device=None, engine_kwargs=None):
    """`n_ladders` ladders whose `mc_params.n_traj` temperatures are spread over the ranks of `dist` (n_traj % world == 0).
    SphericalBC / NVT.  Returns (states of the local slots, histograms of the local slots [L][T_loc][n_bin+2], (first slot, count)).
    `comm_device`: torch device of the communication buffers (the rank's GPU for NCCL, None for gloo)."""
    import copy
    import torch
    from .engine import Engine
    from .host import initial_arrays, sweep_config
    if ensemble.kind != abi.ENS_NVT:
        raise NotImplementedError("split ladders are built for NVT / SphericalBC")
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    T = mc_params.n_traj
    if T % world:
        raise ValueError("n_traj must be a multiple of the number of ranks")
    T_loc, t0 = T // world, rank * (T // world)
    mcp = copy.copy(mc_params); mcp.n_traj = T_loc
    tloc = copy.copy(temp); tloc.t_grid = temp.t_grid[t0:t0 + T_loc]; tloc.beta_grid = temp.beta_grid[t0:t0 + T_loc]
    cfg = sweep_config(mcp, tloc, start_config, potential, ensemble, n_ladders=n_ladders, seed=seed, device=device)
    eng = (engine_cls or Engine)(cfg, **(engine_kwargs or {}))
    betas = np.asarray(temp.beta_grid, dtype=np.float64)
    try:
        pos, box = initial_arrays(mcp, start_config, n_ladders)
        eng.upload_states(pos, box)
        words = eng.rng_words()
        slot, _ = global_stream_states(seed, 0, n_ladders, n_ladders, T, words)
        lad_dummy = np.zeros((n_ladders, 4), dtype=np.uint64); lad_dummy[:, 0] = 1
        eng.set_rng_states(np.ascontiguousarray(slot[:, t0:t0 + T_loc]), lad_dummy)
        lad = [[int(x) for x in julia_seed_state(seed + n_ladders * T * (words // 4) + l)] for l in range(n_ladders)]
        exc = np.zeros((n_ladders, T_loc, 2), dtype=np.int64)      # count_exc of the local slots (attempts, accepted)

        def cycle(i, sample):
            eng.run_phase(i, 1, sample, 1)                                   # mc_step!
            plan = []
            for l in range(n_ladders):                                       # src/MCRun.jl:203, Exchange.jl:135,145
                if _jl_rand(lad[l]) < 0.1:
                    n_exc = _jl_rand_range(lad[l], T - 1)
                    plan.append((l, n_exc - 1, n_exc, _jl_rand(lad[l])))
            if plan:
                en_loc = eng.download_states(("en_tot",))["en_tot"]
                if world > 1:
                    mine = torch.from_numpy(np.ascontiguousarray(en_loc))
                    if comm_device is not None:
                        mine = mine.to(comm_device)
                    parts = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(parts, mine)
                    en = np.concatenate([p.cpu().numpy() for p in parts], axis=1)
                else:
                    en = en_loc
                for l, a, b, u in plan:
                    ra, rb = a // T_loc, b // T_loc
                    for s_, r_ in ((a, ra), (b, rb)):
                        if r_ == rank:
                            exc[l, s_ - t0, 0] += 1
                    with np.errstate(over="ignore"):
                        p = float(np.exp((betas[a] - betas[b]) * (en[l, a] - en[l, b])))       # exc_acceptance :103-108
                    p = p if p != p else min(1.0, p)
                    if not p > u:
                        continue
                    for s_, r_ in ((a, ra), (b, rb)):
                        if r_ == rank:
                            exc[l, s_ - t0, 1] += 1
                    if ra == rank and rb == rank:                            # exc_trajectories! :115-124
                        ca, cb = eng.export_config(l, a - t0), eng.export_config(l, b - t0)
                        eng.import_config(l, a - t0, cb); eng.import_config(l, b - t0, ca)
                    elif ra == rank or rb == rank:
                        s_loc, peer = (a - t0, rb) if ra == rank else (b - t0, ra)
                        eng.import_config(l, s_loc, _swap_with_peer(dist, peer, eng.export_config(l, s_loc), comm_device))
            eng.run_phase(i, 1, sample, 2)                                   # update_max_stepsize!, sampling_step!

        for i in range(1, eq_cycles + 1):
            cycle(i, False)
        if eq_cycles:
            eng.reset_counters()
            exc[:] = 0
        eng.set_hist_bounds(*bounds)
        for i in range(1, mc_params.mc_cycles + 1):
            cycle(i, True)
        eng.sync()
        states = eng.download_states()
        states["counters"][:, :, abi.C_EXC_ATT] += exc[:, :, 0]
        states["counters"][:, :, abi.C_EXC_ACC] += exc[:, :, 1]
        return states, eng.download_hist(), (t0, T_loc)
    finally:
        eng.close()
