"""Pins of the oracle's generators against published known-answer vectors.

The reference ships no golden vectors; what CAN be pinned is that the oracle's rf() is the reference's
generator (MT19937-64 genrand64_real1, src/mt19937.f90:117-225) and that the Philox stream the CUDA
kernels use is the standard Philox4x32-10.
"""
import numpy as np

from pantera_pic_dsmc_b200 import capi


def test_mt19937_64_cxx_standard_10000th(oracle):
    """std::mt19937_64 default seed 5489: the 10000th output is 9981545732273789042 (C++ standard)."""
    oracle.set_config(capi.make_config(rng_seed_global=5489, rank=0))
    x = 0
    for _ in range(10000):
        x = oracle.mt_next_u64()
    assert x == 9981545732273789042


def test_mt19937_64_reference_seed_first_draw(oracle):
    """Seed 65644 (the reference's example input, input:58), rank 0: first word and first rf()."""
    oracle.set_config(capi.make_config(rng_seed_global=65644, rank=0))
    assert oracle.mt_next_u64() == 16554339988561158890
    oracle.set_config(capi.make_config(rng_seed_global=65644, rank=0))
    assert oracle.rf() == 0.89741256898308364


def test_mt_seed_rule_per_rank(oracle):
    """RNG_SEED_LOCAL = RNG_SEED_GLOBAL*(PROC_ID+1) (src/initialization.f90:2294-2296)."""
    oracle.set_config(capi.make_config(rng_seed_global=65644, rank=2))
    a = oracle.mt_next_u64()
    oracle.set_config(capi.make_config(rng_seed_global=3 * 65644, rank=0))
    assert oracle.mt_next_u64() == a


def test_rf_closed_interval_scaling(oracle):
    """genrand64_real1: (x >> 11) / (2^53 - 1) in [0,1]."""
    oracle.set_config(capi.make_config(rng_seed_global=1, rank=0))
    w = oracle.mt_next_u64()
    oracle.set_config(capi.make_config(rng_seed_global=1, rank=0))
    assert oracle.rf() == (w >> 11) * (1.0 / 9007199254740991.0)


def test_philox4x32_10_random123_kat(oracle):
    """Random123 known-answer vectors for philox4x32-10."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_stream_contract(oracle_philox):
    """Stream contract shared with the device (pgpu_internal.cuh PgRng): key/counter layout and the
    53-bit mapping of two uniforms per block."""
    o = oracle_philox
    o.set_config(capi.make_config(rng_seed_global=65644, rank=1))
    seed = 65644 * 2
    kind, tID, e0, e1 = 2, 7, (5 << 32) | 9, 11
    o.rng_stream(kind, tID, e0, e1)
    key = [(seed & 0xFFFFFFFF) ^ ((kind * 0x632BE5AB) & 0xFFFFFFFF), ((seed >> 32) & 0xFFFFFFFF) ^ tID]
    draws = [o.rf() for _ in range(4)]
    for blk in range(2):
        out = o.philox([blk, e0 & 0xFFFFFFFF, e0 >> 32, e1], key)
        for h in range(2):
            u = (out[2 * h + 1] << 32) | out[2 * h]
            assert draws[2 * blk + h] == (u >> 11) * 2.0 ** -53
    assert all(0.0 <= d < 1.0 for d in draws)


def test_philox_word_stream_pick_contract(oracle_philox, oracle):
    """A collision candidate is ONE block: word 0 -> first index, word 1 -> second index (floor(N*word/2^32) for
    N <= 65536), words 2,3 -> the acceptance draw.  Larger N and the MT19937 mode use INT(N*rf()) as the
    reference does (collisions.f90:263-271)."""
    o = oracle_philox
    o.set_config(capi.make_config(rng_seed_global=65644, rank=0))
    kind, tID, e0, e1 = 2, 3, 1234, 7
    key = [(65644 & 0xFFFFFFFF) ^ ((kind * 0x632BE5AB) & 0xFFFFFFFF), tID]
    blk0 = o.philox([0, e0, 0, e1], key)
    blk1 = o.philox([1, e0, 0, e1], key)
    o.rng_stream(kind, tID, e0, e1)
    n = 40
    assert o.pick(n) == (blk0[0] * n) >> 32 and o.pick(n) == (blk0[1] * n) >> 32
    assert o.rf() == (((blk0[3] << 32) | blk0[2]) >> 11) * 2.0 ** -53
    assert o.rf() == (((blk1[1] << 32) | blk1[0]) >> 11) * 2.0 ** -53       # the next block starts here
    # a uniform may straddle two blocks: pick (word 2 of block 1), then rf = (word 3 of block 1, word 0 of block 2)
    blk2 = o.philox([2, e0, 0, e1], key)
    assert o.pick(7) == (blk1[2] * 7) >> 32
    assert o.rf() == (((blk2[0] << 32) | blk1[3]) >> 11) * 2.0 ** -53
    # large N: two words through rf()
    o.rng_stream(kind, tID, e0, e1)
    big = 10_000_000
    u = (((blk0[1] << 32) | blk0[0]) >> 11) * 2.0 ** -53
    assert o.pick(big) == int(big * u)
    # uniformity of the one-word pick
    o.rng_stream(1, 0, 99, 0)
    k = np.array([o.pick(10) for _ in range(20000)])
    assert k.min() == 0 and k.max() == 9
    assert np.all(np.abs(np.bincount(k, minlength=10) - 2000) < 5 * np.sqrt(2000 * 0.9))
    # sequential reference stream: INT(N*rf()) on the MT19937-64 draws
    oracle.set_config(capi.make_config(rng_seed_global=65644, rank=0))
    r = oracle.rf()
    oracle.set_config(capi.make_config(rng_seed_global=65644, rank=0))
    assert oracle.pick(40) == int(40 * r)


def test_philox_streams_are_independent(oracle_philox):
    o = oracle_philox
    o.set_config(capi.make_config())
    seen = set()
    for ent in range(200):
        o.rng_stream(1, 0, ent, 0)
        seen.add(o.rf())
    assert len(seen) == 200
    o.rng_stream(1, 3, 42, 0)
    a = [o.rf() for _ in range(5)]
    o.rng_stream(1, 3, 42, 0)
    assert a == [o.rf() for _ in range(5)]


def test_uniformity_of_both_generators(oracle, oracle_philox):
    for o in (oracle, oracle_philox):
        o.set_config(capi.make_config())
        o.rng_stream(1, 0, 0, 0)
        x = np.array([o.rf() for _ in range(20000)])
        assert abs(x.mean() - 0.5) < 4 * np.sqrt(1 / 12 / len(x))
        assert abs(x.var() - 1 / 12) < 0.005
