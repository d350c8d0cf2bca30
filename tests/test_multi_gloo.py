"""N > 1 host logic on CPU: two gloo ranks shard a job's ladders, run them (through the oracle engine — the
CUDA library needs a GPU) and all-reduce the histograms; the result must equal the single-process job
bit for bit, because random streams are indexed by the global ladder number."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ptmc_b200 as P
from ptmc_b200 import abi, multi

TOTAL, CYCLES, SEED = 5, 12, 2024
BOUNDS = (-0.0075, -0.0045, None, None)


def test_ladder_shard_partitions():
    for total in (1, 5, 8, 4096):
        for world in (1, 2, 3, 8):
            blocks = [multi.ladder_shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_python_seed_matches_oracle(oracle):
    for s in (0, 1234, 2 ** 32 + 5, 2 ** 40):
        assert np.array_equal(multi.julia_seed_state(s), oracle.julia_seed(s))


def _job(dist_mod):
    from oracle import oracle
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=6, cycles=CYCLES)
    return multi.run_sharded(mcp, temp, start, pot, ens, total_ladders=TOTAL, seed=SEED, dist=dist_mod,
                             engine_cls=oracle.OracleEngine, bounds=BOUNDS, eq_cycles=4)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        states, hist, (first, count) = _job(dist)
        np.savez(os.path.join(out, f"rank{rank}.npz"), pos=states["pos"], counters=states["counters"], hist=hist,
                 first=first, count=count)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_equal_single_process(tmp_path, oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    states, hist, _ = _job(None)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert [int(p["count"]) for p in parts] == [3, 2]
    assert np.array_equal(np.concatenate([p["pos"] for p in parts]), states["pos"])
    assert np.array_equal(np.concatenate([p["counters"] for p in parts]), states["counters"])
    for p in parts:
        assert np.array_equal(p["hist"], hist)          # every rank holds the job-wide histogram
    assert hist.sum() == TOTAL * 6 * CYCLES


# ------------------------------------------------------------------ one ladder spanning two ranks --
ST, SL, SCYC = 6, 3, 40
SBOUNDS = (-0.0075, -0.0045, None, None)


def _split_deck():
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=ST, cycles=SCYC)
    mcp.n_adjust = 10
    return mcp, temp, start, pot, ens


def _split_worker(rank, world, port, out):
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        states, hist, (t0, cnt) = multi.run_split_ladder(*_split_deck(), n_ladders=SL, seed=SEED, dist=dist,
                                                         engine_cls=oracle.OracleEngine, bounds=SBOUNDS, eq_cycles=6)
        np.savez(os.path.join(out, f"split{rank}.npz"), pos=states["pos"], counters=states["counters"], hist=hist,
                 en=states["en_tot"], md=states["max_displ"], ham=states["ham"], t0=t0)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ladder_split_over_two_ranks_equals_one_handle(tmp_path, oracle):
    # the reference semantics of a whole ladder on one handle (exchange inside mc_cycle!) ...
    from oracle import oracle as orc
    from ptmc_b200.host import initial_arrays, sweep_config
    mcp, temp, start, pot, ens = _split_deck()
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=SL, seed=SEED)
    eng = orc.OracleEngine(cfg)
    pos, box = initial_arrays(mcp, start, SL)
    eng.upload_states(pos, box)
    multi.seed_global(eng, SEED, 0, SL)
    eng.run_cycles(1, 6, False)
    eng.reset_counters()
    eng.set_hist_bounds(*SBOUNDS)
    eng.run_cycles(1, SCYC, True)
    ref, ref_hist = eng.download_states(), eng.download_hist()
    eng.close()
    assert ref["counters"][:, :, abi.C_EXC_ACC].sum() > 0, "no accepted exchange in the reference run: test is vacuous"
    # ... against the same ladders with their temperatures split over two ranks and the exchange done through
    # all_gather + point-to-point swaps (also across the rank boundary: slots 2 | 3)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"split{r}.npz") for r in range(2)]
    assert [int(p["t0"]) for p in parts] == [0, 3]
    cat = lambda k: np.concatenate([p[k] for p in parts], axis=1)
    assert np.array_equal(cat("pos"), ref["pos"])
    assert np.array_equal(cat("counters"), ref["counters"])
    assert np.array_equal(cat("hist"), ref_hist)
    assert np.array_equal(cat("en"), ref["en_tot"]) and np.array_equal(cat("md"), ref["max_displ"]) and np.array_equal(cat("ham"), ref["ham"])
    # an exchange across the boundary between the two ranks did happen
    assert ref["counters"][:, 2, abi.C_EXC_ACC].sum() + ref["counters"][:, 3, abi.C_EXC_ACC].sum() > 0
