"""JAX PRNG restatement pinned by JAX's own known-answer vectors.

Sources of the constants: jax tests/random_test.py::testThreefry2x32 (the three Random123 vectors),
and the documented outputs for key 0 (`random.split(PRNGKey(0))`, `random.normal(PRNGKey(0))` =
-0.20584226, `random.uniform(PRNGKey(0))` = 0.41845703, `random.normal(key,(3,))`), all for the
legacy (non-partitionable) threefry layout the reference's JAX era used.
"""
import numpy as np

from mbd_b200 import prng


def test_threefry_kat(orc):
    assert [hex(v) for v in orc.threefry2x32([0, 0], [0, 0])] == ["0x6b200159", "0x99ba4efe"]
    assert [hex(v) for v in orc.threefry2x32([0xFFFFFFFF] * 2, [0xFFFFFFFF] * 2)] == ["0x1cb996fc", "0xbb002be7"]
    assert [hex(v) for v in orc.threefry2x32([0x13198A2E, 0x03707344], [0x243F6A88, 0x85A308D3])] == ["0xc4923a9c", "0x483df7a0"]
    o0, o1 = prng.threefry2x32(np.uint32([0x13198A2E, 0x03707344]), np.uint32([0x243F6A88]), np.uint32([0x85A308D3]))
    assert (hex(o0[0]), hex(o1[0])) == ("0xc4923a9c", "0x483df7a0")


def test_key0_documented_values(orc):
    k = orc.prng_key(0)
    assert orc.split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.split(prng.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert np.float32(orc.normal(k, ())) == np.float32(-0.20584226)
    assert np.float32(orc.uniform(k, ())) == np.float32(0.41845703)
    assert np.float32(prng.uniform(prng.PRNGKey(0), ())) == np.float32(0.41845703)
    n3 = orc.normal(k, (3,))
    assert np.allclose(n3, [1.8160863, -0.48262316, 0.33988908], rtol=0, atol=1e-7)


def test_host_prng_matches_oracle(orc):
    rng = np.random.default_rng(0)
    for total in (1, 2, 5, 47, 128, 1001):
        key = rng.integers(0, 2**32, size=2, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(prng.random_bits(key, total), orc.random_bits(key, total))
    key = np.uint32([123, 456])
    assert np.array_equal(prng.split(key, 3), orc.split(key, 3))
    assert np.array_equal(prng.uniform(key, (23,), -0.01, 0.01), orc.uniform(key, (23,), -0.01, 0.01))


def test_normal_slices_are_consistent(orc):
    """element (n, j) depends only on the global index: the basis of sample sharding."""
    key = np.uint32([7, 9])
    full = orc.normal(key, (64 * 50,))
    part = orc.normal(key, (64 * 50,), begin=1000, end=1800)
    assert np.array_equal(full[1000:1800], part)
    Ybar = np.linspace(-0.3, 0.3, 50).astype(np.float32)
    allY = orc.sample_Y0s(key, 64, 50, 0.7, Ybar)
    assert np.array_equal(allY[16:32], orc.sample_Y0s(key, 64, 50, 0.7, Ybar, 16, 32))
    assert np.abs(allY).max() <= 1.0
    z = orc.normal(np.uint32([1, 2]), (200000,))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


def test_partitionable_layout_is_self_consistent():
    """the compatibility switch for JAX >= 0.5 (jax_threefry_partitionable=True) [jax-recalled, unpinned]: host key management,
    the oracle's samplers and (on the GPU, tests/test_rollout_gpu.py) the in-kernel sampler implement the SAME layout:
    bits[i] = o0 ^ o1 of block (0, i); split(key, n)[i] = (o0, o1) of block (0, i)"""
    from mbd_b200 import prng
    from oracle import oracle as orc
    key = prng.PRNGKey(42)
    try:
        prng._PARTITIONABLE = True          # host layout only (no GPU library call in a CPU test)
        orc.set_prng_layout(True)
        b = prng.random_bits(key, 7)
        o0, o1 = prng.threefry2x32(key, np.zeros(7, np.uint32), np.arange(7, dtype=np.uint32))
        assert np.array_equal(b, o0 ^ o1) and np.array_equal(b, orc.random_bits(key, 7))
        assert np.array_equal(prng.split(key, 3), orc.split(key, 3))
        a, c = prng.split2(key)
        assert np.array_equal(np.stack([a, c]), prng.split(key, 2))
        assert np.array_equal(np.clip(orc.normal(key, (5,)), -1, 1).view(np.uint32),
                              orc.sample_Y0s(key, 1, 5, 1.0, np.zeros(5, np.float32)).ravel().view(np.uint32))
    finally:
        prng._PARTITIONABLE = False
        orc.set_prng_layout(False)
    assert not np.array_equal(prng.random_bits(key, 7), b)     # and it differs from the legacy layout
