"""Stream seeding rules (include/ptmc_b200.h): ptmc_seed = consecutive Julia seeds (parity / compatibility), ptmc_seed_job = SHA-256 of
(seed, global stream id) so that different jobs share no stream and a ladder's streams do not depend on the sharding."""
import hashlib

import numpy as np
import pytest

import ptmc_b200 as P
from ptmc_b200 import abi, multi
from ptmc_b200.host import sweep_config


def _engine(oracle, n_ladders, seed, rng_mode=abi.RNG_BATCH8):
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=5, cycles=5)
    return oracle.OracleEngine(sweep_config(mcp, temp, start, pot, ens, n_ladders=n_ladders, seed=seed, rng_mode=rng_mode))


def test_julia_seed_of_wide_integers():
    # Random.seed!(n) hashes the little-endian UInt32 words of n with trailing zero words dropped (make_seed): the 128-bit rule
    # reduces to the 64-bit one for hi == 0, and to a 12-byte message for a small stream id
    assert np.array_equal(multi.julia_seed128_state(1234, 0), multi.julia_seed_state(1234))
    assert np.array_equal(multi.julia_seed128_state(2 ** 40 + 7, 0), multi.julia_seed_state(2 ** 40 + 7))
    want = np.frombuffer(hashlib.sha256((1234).to_bytes(8, "little") + (5).to_bytes(4, "little")).digest(), dtype="<u8")
    assert np.array_equal(multi.julia_seed128_state(1234, 5), want)
    want = np.frombuffer(hashlib.sha256((1234).to_bytes(8, "little") + ((1 << 63) | 3).to_bytes(8, "little")).digest(), dtype="<u8")
    assert np.array_equal(multi.julia_seed128_state(1234, (1 << 63) | 3), want)


@pytest.mark.parametrize("rng_mode", [abi.RNG_BATCH8, abi.RNG_XOSHIRO])
def test_job_seeding_rule(oracle, rng_mode):
    eng = _engine(oracle, 4, 1234, rng_mode)
    try:
        s0, l0 = eng.get_rng_states()                        # create seeds with (cfg.seed, first_ladder = 0)
        ws, wl = multi.job_stream_states(1234, 0, 4, 5, eng.rng_words())
        assert np.array_equal(s0, ws) and np.array_equal(l0, wl)
        eng.seed_job(1234, 2)                                # local ladder l is global ladder 2 + l
        s2, l2 = eng.get_rng_states()
        assert np.array_equal(s2[:2], s0[2:]) and np.array_equal(l2[:2], l0[2:])
        eng.seed_job(1235, 0)
        s1, l1 = eng.get_rng_states()
        key = lambda a: {tuple(r) for r in a.reshape(-1, 4).tolist()}
        all0, all1 = key(s0) | key(l0), key(s1) | key(l1)
        assert len(all0) == s0.size // 4 + 4 and not (all0 & all1)          # every stream distinct, nothing shared between the two jobs
        # the consecutive rule (ptmc_seed) DOES share: seeds 1234 and 1235 are the same streams shifted by one
        eng.seed(1234); a, _ = eng.get_rng_states()
        eng.seed(1235); b, _ = eng.get_rng_states()
        assert len(key(a) & key(b)) == a.size // 4 - 1
    finally:
        eng.close()


@pytest.mark.gpu
def test_job_seeding_device_equals_oracle(oracle):
    mcp, temp, start, pot, ens = P.host.deck_ne13(n_traj=5, cycles=5)
    cfg = sweep_config(mcp, temp, start, pot, ens, n_ladders=3, seed=99)
    g, c = P.Engine(cfg), oracle.OracleEngine(cfg)
    for first in (0, 7):
        g.seed_job(4242, first); c.seed_job(4242, first)
        (gs, gl), (cs, cl) = g.get_rng_states(), c.get_rng_states()
        assert np.array_equal(gs, cs) and np.array_equal(gl, cl)
    g.close(); c.close()
