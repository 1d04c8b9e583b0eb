"""oracle/mt19937_oracle.py (the CPU restatement of numpy's legacy stream in the device kernels' parallel decomposition) against
numpy itself: drawn values and the generator state left behind, bit for bit.  numpy's RandomState is the third-party arithmetic
behind MPCController.get_random_action (policies/mpc_controller.py:67-69) and the CEM draw (:85)."""
import numpy as np
import pytest

from oracle import mt19937_oracle as MT


def _state(rs):
    name, key, pos, has_gauss, cached = rs.get_state()
    assert name == "MT19937"
    return np.array(key, np.uint32), int(pos), int(has_gauss), float(cached)


def _advance(seed, burn):
    rs = np.random.RandomState(seed)
    if burn:
        rs.random_sample(burn)
    return rs


def test_block_refill_decomposition_equals_the_sequential_recurrence():
    key = _state(np.random.RandomState(3))[0]
    for _ in range(3):
        want = MT.next_block_sequential(key)
        got = MT.next_block(key)
        np.testing.assert_array_equal(got, want)
        key = got


def test_raw_stream_words_equal_numpy_random_raw():
    rs = _advance(11, 100)                                            # position mid-block
    key, pos, _, _ = _state(rs)
    raw = MT.raw_blocks(key, pos, 3000)
    assert raw.size == 624 * -(-(pos + 3000) // 624)
    want = rs.randint(0, 2 ** 32, size=3000, dtype=np.uint64).astype(np.uint32)   # one 32-bit output per value
    np.testing.assert_array_equal(MT.temper(raw[pos:pos + 3000]), want)


@pytest.mark.parametrize("seed,burn,rows,env", [(0, 0, 500, "hc"), (5, 311, 1248, "ant"), (9, 1, 37, "arm"), (2, 623, 1, "hc"),
                                                (4, 0, 52, "hc")])
def test_uniform_candidates_and_state_equal_numpy(seed, burn, rows, env):
    low, high = {"hc": (-np.ones(6), np.ones(6)), "ant": (-150.0 * np.ones(8), 150.0 * np.ones(8)),
                 "arm": (np.linspace(-3, -1, 7), np.linspace(0.5, 2, 7))}[env]
    rs = _advance(seed, burn)
    key, pos, _, _ = _state(rs)
    got, (key2, pos2) = MT.uniform(key, pos, low, high, rows)
    want = rs.uniform(low=low, high=high, size=(rows,) + low.shape)   # the reference's draw (:67-69)
    np.testing.assert_array_equal(got, want)
    wkey, wpos, _, _ = _state(rs)
    assert pos2 == wpos
    np.testing.assert_array_equal(key2, wkey)


def test_uniform_draw_that_ends_exactly_on_a_block_boundary_leaves_pos_624():
    rs = np.random.RandomState(4)                                     # pos = 624 after seeding; 52 x 6 doubles = 624 words
    key, pos, _, _ = _state(rs)
    assert pos == 624
    _, (key2, pos2) = MT.uniform(key, pos, -np.ones(6), np.ones(6), 52)
    rs.uniform(-1, 1, size=(52, 6))
    wkey, wpos, _, _ = _state(rs)
    assert pos2 == wpos == 624
    np.testing.assert_array_equal(key2, wkey)


@pytest.mark.parametrize("seed,burn,counts", [(1, 0, (900, 900, 900)), (7, 77, (3, 1, 4, 1, 5)), (8, 500, (1001, 2, 999)),
                                              (6, 0, (0, 1, 0, 2)), (12, 13, (5400,))])
def test_cem_normal_draws_and_state_equal_numpy(seed, burn, counts):
    """Consecutive np.random.normal(size=n) calls (the CEM iterations, :84-85): odd sizes leave the second polar value cached and
    the next call starts with it; the state (key, pos, has_gauss, cached) after every call equals numpy's."""
    rs = _advance(seed, burn)
    key, pos, has_gauss, cached = _state(rs)
    for n in counts:
        got, (key, pos, has_gauss, cached) = MT.legacy_normal(key, pos, has_gauss, cached, n)
        want = rs.normal(size=n)
        np.testing.assert_array_equal(got, want)
        wkey, wpos, wg, wc = _state(rs)
        assert (pos, has_gauss) == (wpos, wg)
        if wg:
            assert cached == wc
        np.testing.assert_array_equal(key, wkey)


def test_normal_draw_does_not_depend_on_the_attempt_budget():
    rs = _advance(21, 5)
    key, pos, g, c = _state(rs)
    a, sa = MT.legacy_normal(key, pos, g, c, 301)
    b, sb = MT.legacy_normal(key, pos, g, c, 301, attempts=4000)
    np.testing.assert_array_equal(a, b)
    assert sa[1:] == sb[1:]
    np.testing.assert_array_equal(sa[0], sb[0])
    with pytest.raises(AssertionError):
        MT.legacy_normal(key, pos, g, c, 301, attempts=100)


def test_uniform_after_normal_keeps_the_cached_gaussian():
    """A uniform draw between two normal draws touches only (key, pos): the cached value survives (what the RS planning call
    relies on when it advances numpy's state struct in place)."""
    rs = _advance(30, 0)
    key, pos, g, c = _state(rs)
    z1, (key, pos, g, c) = MT.legacy_normal(key, pos, g, c, 3)
    u, (key, pos) = MT.uniform(key, pos, -np.ones(6), np.ones(6), 10)
    z2, (key, pos, g, c) = MT.legacy_normal(key, pos, g, c, 2)
    np.testing.assert_array_equal(z1, rs.normal(size=3))
    np.testing.assert_array_equal(u, rs.uniform(-1, 1, size=(10, 6)))
    np.testing.assert_array_equal(z2, rs.normal(size=2))
    wkey, wpos, wg, wc = _state(rs)
    assert (pos, g) == (wpos, wg)
    np.testing.assert_array_equal(key, wkey)
