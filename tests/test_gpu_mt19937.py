"""The reference's candidate draw (policies/mpc_controller.py:67-69, 114: np.random.uniform from numpy's global MT19937
stream) regenerated on the device -- oracle = numpy itself, compared BIT FOR BIT: the float32 candidate tensor, the float64
chosen actions and the generator state left behind (np.random.get_state()).  Through the C ABI (l2a_plan_create_ex /
l2a_plan_run_ex) and through MPCController's default sampler."""
import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O
from tests.helpers import assert_argmax_consistent, make_engine

pytestmark = pytest.mark.gpu


def _state_equal(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3] and a[4] == b[4]


@pytest.mark.parametrize("env,n,h,m,burn", [
    ("half_cheetah", 50, 3, 1, 0),          # pos = 624 right after seeding (refill before the first word)
    ("half_cheetah", 52, 1, 1, 7),          # exactly one block: 52 * 6 * 2 = 624 words, starting mid-block
    ("half_cheetah", 104, 1, 1, 0),         # ends exactly on a block boundary (pos stays 624, key = last block)
    ("ant", 333, 7, 2, 12345),              # odd sizes, A = 8, bounds +-150
    ("half_cheetah", 2000, 20, 1, 3),       # the headline draw: 240 000 doubles = 770 blocks
])
def test_device_mt19937_matches_numpy_stream_bit_for_bit(env, n, h, m, burn):
    prob = O.make_problem(env, hidden_sizes=(128, 128), n_sets=1, m=m, seed=5)
    eng = make_engine(prob)
    low, high = prob["low"], prob["high"]
    A = prob["act_dim"]
    for call in range(3):                   # first call direct, then graph capture, then graph replay
        np.random.seed(1000 + call)
        if burn:
            np.random.uniform(size=burn)    # the stream is somewhere in the middle of a block
        state0 = np.random.get_state()
        want = np.random.uniform(low=low, high=high, size=(h * n * m, A)).reshape(h, n * m, A)      # the reference's draw
        state_want = np.random.get_state()
        np.random.set_state(state0)
        acts, ret, idx = eng.plan_rs_host(prob["obs0"], n, h, prob["reward_kind"], prob["dt"], low, high, sampler="mt19937")
        state_got = np.random.get_state()
        cand = eng.last_plan_candidates()
        np.testing.assert_array_equal(cand, want.astype(np.float32))
        assert _state_equal(state_got, state_want), "generator state after the call differs from numpy's"
        # the float64 chosen actions are the reference's float64 candidates of time step 0 (mpc_controller.py:118, 129)
        cand0 = want[0].reshape(m, n, A)
        np.testing.assert_array_equal(acts, cand0[range(m), idx])
        # and the choice itself is the oracle's for these candidates
        returns = O.rollout_returns(prob["obs0"], want, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "shared")
        assert_argmax_consistent(idx, returns)
    assert eng.last_plan_uses_graph()


def test_default_controller_sampler_is_the_device_regenerated_numpy_stream():
    """MPCController() with no sampler argument: same chosen actions, same float64 values and same global numpy state as the
    host-drawn path (sampler="numpy_host"), i.e. as the reference controller."""
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=3, seed=8)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(128, 128))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    dev_ctrl = MPCController("policy", env, model, n_candidates=300, horizon=6)
    host_ctrl = MPCController("policy", env, model, n_candidates=300, horizon=6, sampler="numpy_host")
    assert dev_ctrl.sampler == "numpy"
    for step in range(4):
        obs = prob["obs0"] + 0.01 * step
        np.random.seed(77 + step)
        a_host, _ = host_ctrl.get_actions(obs)
        s_host = np.random.get_state()
        np.random.seed(77 + step)
        a_dev, _ = dev_ctrl.get_actions(obs)
        s_dev = np.random.get_state()
        assert a_dev.dtype == np.float64
        np.testing.assert_array_equal(a_dev, a_host)
        assert _state_equal(s_dev, s_host)
    # consecutive calls without reseeding continue the stream like the host path does
    np.random.seed(5)
    seq_host = [host_ctrl.get_actions(prob["obs0"])[0] for _ in range(3)]
    np.random.seed(5)
    seq_dev = [dev_ctrl.get_actions(prob["obs0"])[0] for _ in range(3)]
    for x, y in zip(seq_host, seq_dev):
        np.testing.assert_array_equal(x, y)
    assert torch.cuda.is_available()
