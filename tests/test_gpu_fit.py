"""On-device fit against the host oracle: a seeded fit() of the product (torch autograd on the engine's
resident parameters, TF-Adam update) and the oracle's restatement (numpy float32 backprop / float64 autograd for MAML) consume the
same numpy draws, see the same batches and must end with the same weights to 1e-4."""
import numpy as np
import pytest

from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu


def _data(seed, n, D=20, A=6):
    rng = np.random.RandomState(seed)
    obs = rng.normal(size=(n, D))
    act = rng.uniform(-1, 1, size=(n, A))
    W = 0.05 * rng.normal(size=(D + A, D))
    nxt = obs + np.tanh(np.concatenate([obs, act], 1) @ W) + 0.01 * rng.normal(size=(n, D))
    return obs, act, nxt


def _assert_params_close(got, want, before, tol=1e-4):
    for k in want:
        moved = np.abs(want[k] - before[k]).max()
        assert moved > 0
        err = np.abs(got[k] - want[k]).max()
        assert err <= tol * max(np.abs(want[k]).max(), 1e-3) + 0.02 * moved, "%s: err %.3e (moved %.3e)" % (k, err, moved)


def test_mlp_fit_matches_oracle_and_aggregates_datasets():
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    model = MLPDynamicsModel("dyn", SyntheticEnv("half_cheetah"), hidden_sizes=(64, 64), batch_size=100, seed=3)
    before = model.get_params()
    obs, act, nxt = _data(0, 700)
    np.random.seed(11)
    model.fit(obs, act, nxt, epochs=3)
    got = model.get_params()
    np.random.seed(11)
    want, valid = O.fit_mlp(before, obs, act, nxt, model.normalization, epochs=3, batch_size=100, learning_rate=1e-3)
    assert model.last_fit_stats["Epochs"] == len(valid) - 1
    _assert_params_close(got, want, before)
    assert abs(model.last_fit_stats["valid_loss"] - valid[-1]) <= 1e-4 * max(1.0, valid[-1])
    # the planner sees the trained weights (tiles refreshed): predict == oracle predict with the trained parameters
    q_obs, q_act, _ = _data(5, 50)
    np.testing.assert_allclose(model.predict(q_obs, q_act), O.predict(q_obs, q_act, got, model.normalization), rtol=1e-4, atol=1e-5)
    # a second fit() trains on BOTH calls' data (mlp_dynamics.py:119-129)
    n_tr, n_te = len(model._dataset_train["obs"]), len(model._dataset_test["obs"])
    assert (n_tr, n_te) == (560, 140)
    obs2, act2, nxt2 = _data(1, 300)
    model.fit(obs2, act2, nxt2, epochs=1, compute_normalization=False)
    assert len(model._dataset_train["obs"]) == 560 + 240 and len(model._dataset_test["obs"]) == 140 + 60
    np.testing.assert_array_equal(model._dataset_train["obs"][:560], model._dataset_train["obs"][:560])


def test_maml_fit_matches_oracle():
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    model = MetaMLPDynamicsModel("dyn", SyntheticEnv("half_cheetah"), hidden_sizes=(64, 64), meta_batch_size=4, batch_size=8,
                                 inner_learning_rate=1e-2, seed=4)
    before = model.get_params()
    obs, act, nxt = _data(2, 10 * 40)
    obs, act, nxt = obs.reshape(10, 40, -1), act.reshape(10, 40, -1), nxt.reshape(10, 40, -1)
    np.random.seed(21)
    model.fit(obs, act, nxt, epochs=2)
    got = model.get_params()
    np.random.seed(21)
    want, valid = O.fit_maml(before, obs, act, nxt, model.normalization, epochs=2, batch_size=8, meta_batch_size=4,
                             learning_rate=1e-3, inner_learning_rate=1e-2)
    _assert_params_close(got, want, before)
    assert abs(model.last_fit_stats["valid_loss"] - valid[-1]) <= 1e-3 * max(1.0, valid[-1])
    assert model.last_fit_stats["Epochs"] == len(valid) - 1
    assert model._dataset_train["obs"].shape == (8, 40, 20) and model._dataset_test["obs"].shape == (2, 40, 20)
