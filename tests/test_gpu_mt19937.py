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


# ------------------------------------------------------------------------------------------------ CEM: np.random.normal on the device
@pytest.mark.parametrize("env,n,h,m,iters", [("arm_7dof", 3, 1, 1, 3),        # 21 normals per draw: odd -> the cached second value carries over
                                             ("arm_7dof", 37, 3, 2, 2),
                                             ("half_cheetah", 300, 6, 1, 3),
                                             ("half_cheetah", 5000, 30, 1, 3)])     # BASELINE cfg4 draw: 900 000 normals per iteration
def test_cem_plan_call_continues_numpy_normal_stream(env, n, h, m, iters):
    """One l2a_plan_run_ex(CEM, MT19937) call must leave np.random exactly where `iters` x np.random.normal(size=(n, m, h*A))
    (policies/mpc_controller.py:85) leave it -- key, position, has_gauss and the cached gaussian -- and plan like the oracle does
    on numpy's own float64 draws."""
    prob = O.make_problem(env, hidden_sizes=(128, 128), n_sets=1, m=m, seed=6)
    eng = make_engine(prob)
    A = prob["act_dim"]
    ha = h * A
    k = max(int(n * 0.1), 1)
    for call in range(3):                                   # direct, capture, replay; the stream simply continues between calls
        if call == 0:
            np.random.seed(4242)
            np.random.uniform(size=5)                       # somewhere inside a block
        state0 = np.random.get_state()
        zs = [np.random.normal(size=(n, m, ha)) for _ in range(iters)]
        state_want = np.random.get_state()
        np.random.set_state(state0)
        acts, ret, idx, mean, std = eng.plan_cem_host(prob["obs0"], n, h, prob["reward_kind"], prob["dt"], prob["low"], prob["high"],
                                                      iters, k, 0.1, sampler="mt19937", compat=True)
        state_got = np.random.get_state()
        assert _state_equal(state_got, state_want), "numpy generator state after the CEM call differs (call %d)" % call
        chosen, best, returns, mean_w, std_w = O.cem_plan(prob["obs0"], zs, prob["low"], prob["high"], prob["param_sets"], prob["norm"],
                                                          prob["reward_kind"], prob["dt"], h, 0.1, 0.1)
        got_returns = eng.last_plan_returns(m, n)
        std_w = np.broadcast_to(std_w, std.shape)           # the reference's std collapses to [H*A] after the first refit (:104)
        if n <= 300:
            np.testing.assert_allclose(mean, mean_w, rtol=2e-3, atol=2e-3)
            np.testing.assert_allclose(std, std_w, rtol=2e-3, atol=2e-3)
            np.testing.assert_allclose(got_returns, returns, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(returns).max()))
        # (at n = 5000 one near-tie rank swap in an earlier iteration changes the compat elite mask and hence every later sample;
        #  tests/test_gpu_configs.py checks that size iteration by iteration)  Same plan -> the float64 sample itself (:106):
        same_plan = np.allclose(mean, mean_w, rtol=1e-6, atol=1e-9) and np.allclose(std, std_w, rtol=1e-6, atol=1e-9)
        if same_plan and np.array_equal(idx, best):
            np.testing.assert_allclose(acts, chosen, rtol=1e-12, atol=1e-12)
    assert eng.last_plan_uses_graph()


def test_cem_controller_default_sampler_matches_host_draw_path():
    """MPCController(use_cem=True) default sampler (device-regenerated numpy stream, one C call) against sampler="numpy_host"
    (host draw, per-iteration kernel calls): same generator state afterwards, same plan up to the float32 rounding of z."""
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=1, seed=9)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(128, 128))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    kw = dict(use_cem=True, n_candidates=400, horizon=5, num_cem_iters=3, percent_elites=0.1, alpha=0.1)
    dev_ctrl = MPCController("policy", env, model, **kw)
    host_ctrl = MPCController("policy", env, model, sampler="numpy_host", **kw)
    np.random.seed(31)
    a_host, _ = host_ctrl.get_actions(prob["obs0"])
    s_host = np.random.get_state()
    np.random.seed(31)
    a_dev, _ = dev_ctrl.get_actions(prob["obs0"])
    s_dev = np.random.get_state()
    assert _state_equal(s_dev, s_host)
    np.testing.assert_allclose(dev_ctrl.last_cem_state[0], host_ctrl.last_cem_state[0].cpu().numpy(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(dev_ctrl.last_cem_state[1], host_ctrl.last_cem_state[1].cpu().numpy()[0:1], rtol=2e-3, atol=2e-3)
    if int(dev_ctrl.last_plan["best_idx"][0]) == int(host_ctrl.last_plan["best_idx"].cpu().numpy()[0]):
        np.testing.assert_allclose(a_dev, a_host, rtol=1e-6, atol=1e-6)


def test_cem_plan_call_philox_normals_are_standard_normal_and_plan_is_consistent():
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=1, seed=10)
    eng = make_engine(prob)
    n, h = 2000, 5
    acts, ret, idx, mean, std = eng.plan_cem_host(prob["obs0"], n, h, prob["reward_kind"], prob["dt"], prob["low"], prob["high"],
                                                  1, 200, 0.0, sampler="philox", compat=False, seed=5)
    # with one iteration from mean 0 / std 1 the samples ARE the normals: the candidate tensor is N(0, 1)
    z = eng.last_plan_candidates().reshape(-1)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and np.abs(z).max() < 6.5
    returns = eng.last_plan_returns(1, n)
    assert int(idx[0]) == int(np.argmax(returns[0])) and abs(float(ret[0]) - float(returns[0, idx[0]])) == 0.0
    samples = z.reshape(n, 1, h * prob["act_dim"])
    np.testing.assert_allclose(acts[0], samples[idx[0], 0, :prob["act_dim"]], rtol=1e-6, atol=1e-7)
    want = O.rollout_returns(prob["obs0"], np.transpose(samples.reshape(n, h, -1), (1, 0, 2)).astype(np.float64), prob["param_sets"],
                             prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "shared")
    from tests.helpers import assert_returns_close
    assert_returns_close(returns, want)
    # true top-k refit (alpha = 0): mean / std of the clipped elite rows
    elite = np.argsort(-returns[0], kind="stable")[:200]
    clipped = np.clip(samples[:, 0, :], np.tile(prob["low"], h), np.tile(prob["high"], h))[elite]
    np.testing.assert_allclose(mean[0], clipped.mean(axis=0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(std[0], clipped.std(axis=0), rtol=1e-5, atol=1e-6)
