"""Host-side sampler logic against the fixture produced by the reference's own Sampler.obtain_samples
(tests/golden/make_golden_sampler.py).  No GPU: the 'lists' formulation with a recording dynamics-model stub."""
import os

import numpy as np
import pytest

from learning_to_adapt_b200.samplers.sampler import Sampler
from learning_to_adapt_b200.samplers.vectorized_env_executor import IterativeEnvExecutor
from tests.sampler_stubs import RecordingModel, StubPolicy, make_env

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.npz")


@pytest.fixture(scope="module")
def sampler_golden():
    return np.load(GOLDEN)


def run_case(g, tag, window, model=None):
    num_envs, path_len, M, rounds = [int(v) for v in g["%s_meta" % tag]]
    env = make_env()
    model = model or RecordingModel()
    sampler = Sampler(env, StubPolicy(env, model), num_rollouts=num_envs, max_path_length=path_len, adapt_batch_size=M, window=window)
    for i, e in enumerate(sampler.vec_env.envs):
        e.seed(100 + i)
    sampler.total_samples = rounds * num_envs * path_len
    return sampler.obtain_samples(), model


@pytest.mark.parametrize("tag", ["a", "b"])
def test_lists_mode_matches_reference_sampler(sampler_golden, tag):
    g = sampler_golden
    paths, model = run_case(g, tag, "lists")
    assert model.steps == list(g["%s_adapt_step" % tag])
    assert model.n_switch == int(g["%s_n_pre_adapt" % tag][0])
    np.testing.assert_array_equal(np.stack(model.obs), g["%s_adapt_obs" % tag])
    np.testing.assert_array_equal(np.stack(model.act), g["%s_adapt_act" % tag])
    np.testing.assert_array_equal(np.stack(model.nxt), g["%s_adapt_next" % tag])
    np.testing.assert_array_equal(np.stack([p["observations"] for p in paths]), g["%s_path_obs" % tag])
    np.testing.assert_array_equal(np.stack([p["actions"] for p in paths]), g["%s_path_act" % tag])
    np.testing.assert_array_equal(np.stack([p["rewards"] for p in paths]), g["%s_path_rew" % tag])
    np.testing.assert_array_equal(np.stack([p["dones"] for p in paths]), g["%s_path_done" % tag])


def test_random_sampling_needs_no_model():
    env = make_env()
    np.random.seed(3)

    class NoPolicy(object):
        def reset(self, dones=None):
            pass

        def get_actions(self, obs):
            raise AssertionError("random=True must not query the policy")

    sampler = Sampler(env, NoPolicy(), num_rollouts=4, max_path_length=9, adapt_batch_size=4)
    paths = sampler.obtain_samples(random=True)
    assert len(paths) == 4 and sampler.total_timesteps_sampled == 36
    for p in paths:
        assert p["observations"].shape == (9, 20) and p["actions"].shape == (9, 6) and p["rewards"].shape == (9,)
        assert p["dones"][-1] and not p["dones"][:-1].any()
        assert (np.abs(p["actions"]) <= 1.0).all()


def test_executor_auto_reset():
    env = make_env()
    ex = IterativeEnvExecutor(env, 3, max_path_length=4)
    first = ex.reset()
    assert len(first) == 3 and ex.num_envs == 3
    for t in range(1, 9):
        obs, rew, dones, infos = ex.step([np.zeros(6)] * 3)
        assert bool(dones.all()) == (t % 4 == 0) and bool(dones.any()) == (t % 4 == 0)
        if t % 4 == 0:                                    # the observation of a finished env is the first of its next path
            assert all(np.abs(o).max() < 1.0 for o in obs) and (ex.ts == 0).all()
    with pytest.raises(AssertionError):
        ex.step([np.zeros(6)] * 2)


def test_numpy_global_mt19937_state_is_reachable_in_place():
    """The default sampler hands numpy's OWN generator state struct to the planning call (engine.numpy_mt19937_state): the address
    must be the {uint32 key[624]; int pos;} that np.random.get_state() reports, stay the same object across seeding / drawing,
    and a state written through it must be what np.random draws from next (that is how the advanced state comes back)."""
    import ctypes as C
    import numpy as np
    from learning_to_adapt_b200.engine import numpy_mt19937_state
    fast = numpy_mt19937_state()
    assert fast is not None, "numpy's bit generator no longer exposes the mt19937 state layout this relies on"
    bg, addr = fast
    key = np.ctypeslib.as_array((C.c_uint32 * 624).from_address(addr))
    pos = C.c_int32.from_address(addr + 624 * 4)
    for seed, burn in ((0, 0), (5, 3), (123, 1000)):
        np.random.seed(seed)
        np.random.uniform(size=burn)
        st = np.random.get_state()
        assert np.array_equal(key, st[1]) and pos.value == st[2]
    other = np.random.RandomState(77)
    other.uniform(size=555)
    s2 = other.get_state()
    with bg.lock:
        key[:] = s2[1]
        pos.value = int(s2[2])
    np.testing.assert_array_equal(np.random.uniform(-1, 1, size=(40, 6)), other.uniform(-1, 1, size=(40, 6)))
    assert np.random.get_state()[2] == other.get_state()[2]
