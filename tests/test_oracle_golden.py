"""Pins oracle/ (the CPU restatement) to the golden fixtures produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
from collections import OrderedDict

import numpy as np
import pytest

from oracle import mpc_oracle as O


def _returns_from_step_rewards(step_rewards, discount):
    ret = np.zeros(step_rewards.shape[1])
    for t in range(step_rewards.shape[0]):
        ret += discount ** t * step_rewards[t]       # mpc_controller.py:126
    return ret


@pytest.mark.parametrize("env", ["half_cheetah", "ant", "arm_7dof"])
def test_reward_closed_forms_match_reference(golden, env):
    d, a, lim, dt, kind = O.ENV_SPECS[env]
    r = O.reward_fn(kind, dt)(golden["reward_%s_obs" % env], golden["reward_%s_act" % env],
                              golden["reward_%s_next" % env])
    np.testing.assert_array_equal(r, golden["reward_%s_out" % env])


def test_predict_matches_reference_choreography(golden):
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=3)
    out = O.predict(golden["predict_obs"], golden["predict_act"], prob["param_sets"][0], prob["norm"])
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, golden["predict_out"])


@pytest.mark.parametrize("tag,env,hidden", [
    ("rs_hc", "half_cheetah", (32, 32)),
    ("rs_hc_disc", "half_cheetah", (48,)),
    ("rs_ant", "ant", (32, 32, 32)),
    ("rs_arm", "arm_7dof", (32, 32)),
])
def test_random_shooting_matches_reference_planner(golden, tag, env, hidden):
    n, h, m, seed = [int(v) for v in golden[tag + "_meta"]]
    discount = float(golden[tag + "_discount"])
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=1, m=m, seed=seed)
    # the oracle's sampler reproduces the reference's draw
    actions = O.sample_rs_actions(seed + 100, prob["low"], prob["high"], h, n * m)
    np.testing.assert_array_equal(actions, golden[tag + "_actions"])
    chosen, best, returns = O.rs_plan(prob["obs0"], actions, prob["param_sets"], prob["norm"],
                                      prob["reward_kind"], prob["dt"], discount)
    ref_returns = _returns_from_step_rewards(golden[tag + "_step_rewards"], discount).reshape(m, n)
    np.testing.assert_array_equal(returns, ref_returns)
    np.testing.assert_array_equal(chosen, golden[tag + "_chosen"])


def test_get_action_returns_2d(golden):
    assert tuple(golden["get_action_shape"]) == (1, 6)


@pytest.mark.parametrize("tag", ["cem_m1", "cem_m2"])
def test_cem_bug_compatible_matches_reference(golden, tag):
    n, h, m, iters, seed = [int(v) for v in golden[tag + "_meta"]]
    pct, alpha = [float(v) for v in golden[tag + "_pct_alpha"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=m, seed=seed)
    rng = np.random.RandomState(seed + 100)
    zs = [rng.normal(size=(n, m, h * prob["act_dim"])) for _ in range(iters)]
    chosen, best, returns, mean, std = O.cem_plan(prob["obs0"], zs, prob["low"], prob["high"], prob["param_sets"],
                                                  prob["norm"], prob["reward_kind"], prob["dt"], h,
                                                  percent_elites=pct, alpha=alpha)
    ref_returns = _returns_from_step_rewards(golden[tag + "_last_step_rewards"], 1.0).reshape(m, n)
    np.testing.assert_array_equal(returns, ref_returns)
    np.testing.assert_array_equal(chosen, golden[tag + "_chosen"])


def _golden_adapted(golden, keys, k):
    return OrderedDict((key, golden["adapt_theta%d_%s" % (k, key.replace("/", "."))]) for key in keys)


def test_adapt_and_per_task_predict_match_reference_choreography(golden):
    K, M, MBS, n_per = [int(v) for v in golden["adapt_meta"]]
    lr = float(golden["adapt_lr"])
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32, 32), n_sets=1, m=3, seed=31)
    theta = prob["param_sets"][0]
    ctx = O.make_adapt_context(41, prob, K, M)
    adapted = O.adapt(*ctx, theta, prob["norm"], lr)
    for k in range(K):
        ref = _golden_adapted(golden, theta.keys(), k)
        for key in theta.keys():
            np.testing.assert_array_equal(adapted[k][key], ref[key])
    post = O.predict_per_task(golden["adapt_query_obs"], golden["adapt_query_act"], adapted, prob["norm"])
    np.testing.assert_array_equal(post, golden["adapt_post_predict"])
    pre = O.predict(golden["adapt_query_obs"], golden["adapt_query_act"], theta, prob["norm"])
    np.testing.assert_array_equal(pre, golden["adapt_pre_predict"])


def test_grbal_planning_matches_reference(golden):
    n, h, m, seed = [int(v) for v in golden["grbal_rs_meta"]]
    K, M, MBS, n_per = [int(v) for v in golden["adapt_meta"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32, 32), n_sets=1, m=m, seed=seed)
    adapted = O.adapt(*O.make_adapt_context(41, prob, K, M), prob["param_sets"][0], prob["norm"],
                      float(golden["adapt_lr"]))
    chosen, best, returns = O.rs_plan(prob["obs0"], golden["grbal_rs_actions"], adapted, prob["norm"],
                                      prob["reward_kind"], prob["dt"], 1.0, mode="per_env")
    ref_returns = _returns_from_step_rewards(golden["grbal_rs_step_rewards"], 1.0).reshape(m, n)
    np.testing.assert_array_equal(returns, ref_returns)
    np.testing.assert_array_equal(chosen, golden["grbal_rs_chosen"])


def test_manual_backprop_matches_torch_autograd():
    """The TF1 half is unpinned upstream; cross-check the restated gradient against torch autograd."""
    import torch
    prob = O.make_problem("ant", hidden_sizes=(64, 64, 64), n_sets=1, m=1, seed=5, out_scale=1.0)
    theta = prob["param_sets"][0]
    rng = np.random.RandomState(0)
    x = rng.normal(size=(16, prob["obs_dim"] + prob["act_dim"])).astype(np.float32)
    t = rng.normal(size=(16, prob["obs_dim"])).astype(np.float32)
    lr = 0.1
    got = O.adapt_one_task(x, t, theta, lr)
    tp = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in theta.values()]
    h = torch.tensor(x, dtype=torch.float64)
    for l in range(len(tp) // 2):
        h = h @ tp[2 * l] + tp[2 * l + 1]
        if l < len(tp) // 2 - 1:
            h = torch.relu(h)
    loss = ((torch.tensor(t, dtype=torch.float64) - h) ** 2).mean()
    grads = torch.autograd.grad(loss, tp)
    for (key, v), p, g in zip(theta.items(), tp, grads):
        want = (p - lr * g).detach().numpy()
        np.testing.assert_allclose(got[key], want, rtol=2e-5, atol=2e-6)


def test_ensemble_mean_reduces_to_single_model():
    prob = O.make_problem("half_cheetah", hidden_sizes=(32,), n_sets=1, m=2, seed=1)
    acts = O.sample_rs_actions(3, prob["low"], prob["high"], 3, 10)
    a = O.rollout_returns(prob["obs0"], acts, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"])
    b = O.rollout_returns(prob["obs0"], acts, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"],
                          mode="ensemble")
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)


def test_rebal_recurrent_planner_matches_reference(golden):
    """Verbatim RNNMPCController (random shooting, hidden state carried across three planning calls) vs the oracle's
    LSTM restatement + planner."""
    n, h, m, seed, hs = [int(v) for v in golden["rebal_meta"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32,), n_sets=1, m=m, seed=seed)
    params = O.xavier_rnn_params(np.random.RandomState(62), prob["obs_dim"] + prob["act_dim"], hs, prob["obs_dim"], out_scale=0.1)
    hidden = (np.zeros((m, hs), np.float32), np.zeros((m, hs), np.float32))
    rng = np.random.RandomState(171)
    obs_t = np.array(prob["obs0"])
    for step in range(3):
        actions = rng.uniform(prob["low"], prob["high"], size=(h * n * m, prob["act_dim"])).reshape(h, n * m, -1)
        chosen, best, returns, hidden = O.rnn_rs_plan(obs_t, actions, hidden, params, prob["norm"], prob["reward_kind"], prob["dt"])
        ref_ret = golden["rebal_step_rewards"][step * h:(step + 1) * h].sum(axis=0).reshape(m, n)
        np.testing.assert_array_equal(returns, ref_ret)
        np.testing.assert_array_equal(chosen, golden["rebal_chosen"][step])
        np.testing.assert_array_equal(hidden[0], golden["rebal_hidden_c"][step])
        np.testing.assert_array_equal(hidden[1], golden["rebal_hidden_h"][step])
        obs_t = obs_t + 0.05 * np.random.RandomState(step).normal(size=obs_t.shape)
    assert tuple(golden["rebal_get_action_shape"]) == (1, 6)


def test_philox_restatement_matches_the_published_known_answer_vectors():
    """Random123's known-answer vectors for philox4x32-10 (the generator behind the device candidate sampler)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array([ctr]), key)[0]
        assert tuple(int(v) for v in got) == want
    a = O.sample_rs_actions_device(7, 3, -np.ones(6), np.ones(6), 4, 10)
    assert a.shape == (4, 10, 6) and a.dtype == np.float32 and np.all(a >= -1) and np.all(a < 1)
    b = O.sample_rs_actions_device(7, 4, -np.ones(6), np.ones(6), 4, 10)
    assert not np.array_equal(a, b)
