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
from ptmc_b200 import multi

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
