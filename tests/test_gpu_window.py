"""GPU tests of the device-resident adaptation window (include/l2a_b200.h l2a_window_*): the windows the
device forms are bit-identical to what the reference's Sampler + host normalisation feed `adapt` with (fixture from the
reference's own Sampler.obtain_samples), and GrBAL's loop gives identical results with and without it."""
import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O
from tests.sampler_stubs import RecordingModel
from tests.test_sampler_cpu import run_case, sampler_golden  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _norm(seed=3, D=20, A=6):
    rng = np.random.RandomState(seed)
    return {"obs": (0.1 * rng.normal(size=D), 0.5 + rng.uniform(size=D)), "act": (0.1 * rng.normal(size=A), 0.5 + rng.uniform(size=A)),
            "delta": (0.01 * rng.normal(size=D), 0.05 + 0.1 * rng.uniform(size=D))}


class DeviceRecordingModel(RecordingModel):
    """A dynamics-model stub whose adaptation windows live on the device; records the gathered, normalised windows."""

    def __init__(self, normalization):
        RecordingModel.__init__(self)
        from learning_to_adapt_b200.engine import PlanningEngine
        self._engine = PlanningEngine(20, 6, (32,), n_sets=1)
        self.normalization = normalization
        self.x, self.target = [], []

    def make_adapt_window(self, n_envs, M):
        from learning_to_adapt_b200.samplers.window import AdaptWindow
        w = AdaptWindow(self._engine, n_envs, M)
        w.set_normalization(self.normalization)
        return w

    def adapt_from_window(self, window):
        x, t = window.gather()
        self.x.append(x.cpu().numpy())
        self.target.append(t.cpu().numpy())
        self.steps.append(self.step)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_device_windows_bit_identical_to_reference_sampler_feed(sampler_golden, tag):
    g = sampler_golden
    norm = _norm()
    paths, model = run_case(g, tag, "device", model=DeviceRecordingModel(norm))
    assert model.steps == list(g["%s_adapt_step" % tag])
    assert model.n_switch == int(g["%s_n_pre_adapt" % tag][0])
    obs, act, nxt = g["%s_adapt_obs" % tag], g["%s_adapt_act" % tag], g["%s_adapt_next" % tag]
    # the reference's feed: float64 normalisation, float32 placeholders (meta_mlp_dynamics.py:334-345, mlp_dynamics.py:242-251)
    x_want = np.concatenate([O.normalize(obs, *norm["obs"]), O.normalize(act, *norm["act"])], axis=-1).astype(np.float32)
    t_want = O.normalize(nxt - obs, *norm["delta"]).astype(np.float32)
    np.testing.assert_array_equal(np.stack(model.x), x_want)
    np.testing.assert_array_equal(np.stack(model.target), t_want)
    np.testing.assert_array_equal(np.stack([p["observations"] for p in paths]), g["%s_path_obs" % tag])
    np.testing.assert_array_equal(np.stack([p["actions"] for p in paths]), g["%s_path_act" % tag])


def _grbal(window, hidden=(128, 128), num_envs=3, M=6, path_len=14, n=60, h=4):
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import LinearWorldEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    from learning_to_adapt_b200.samplers.sampler import Sampler
    env = LinearWorldEnv("half_cheetah", seed=1)
    prob = O.make_problem("half_cheetah", hidden_sizes=hidden, n_sets=1, m=num_envs, seed=21)
    model = MetaMLPDynamicsModel("dyn", env, hidden_sizes=hidden, meta_batch_size=num_envs, inner_learning_rate=1e-2, seed=0)
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    policy = MPCController("policy", env, model, n_candidates=n, horizon=h)
    sampler = Sampler(env, policy, num_rollouts=num_envs, max_path_length=path_len, adapt_batch_size=M, window=window)
    for i, e in enumerate(sampler.vec_env.envs):
        e.seed(50 + i)
    np.random.seed(9)
    paths = sampler.obtain_samples()
    adapted = [model.get_adapted_params(k) for k in range(num_envs)]
    return paths, adapted, prob


def test_grbal_loop_identical_with_device_window_and_host_lists():
    """Whole GrBAL sampling loop (adapt every step + random-shooting MPC): the device-window path and the reference's host-list
    formulation choose the same actions at every step and end with bit-identical adapted weights; the last adaptation also agrees
    with the oracle's restatement of the inner step."""
    paths_d, adapted_d, prob = _grbal("device")
    paths_l, adapted_l, _ = _grbal("lists")
    for pd, pl in zip(paths_d, paths_l):
        np.testing.assert_array_equal(pd["actions"], pl["actions"])
        np.testing.assert_array_equal(pd["observations"], pl["observations"])
    for ad, al in zip(adapted_d, adapted_l):
        for k in ad:
            np.testing.assert_array_equal(ad[k], al[k])
    M = 6
    for k, p in enumerate(paths_d):
        obs, act = p["observations"], p["actions"]
        # last adapt happened before the final step was appended: path length then was len-1
        o, a, nx = obs[-M - 2:-2], act[-M - 2:-2], obs[-M - 1:-1]
        want = O.adapt([o], [a], [nx], prob["param_sets"][0], prob["norm"], 1e-2)[0]
        for key in want:
            scale = max(1e-3, float(np.abs(want[key]).max()))
            assert np.abs(adapted_d[k][key] - want[key]).max() <= 2e-5 * scale, key


def test_window_ring_wraps_resets_and_rejects_short_paths():
    from learning_to_adapt_b200.engine import PlanningEngine
    from learning_to_adapt_b200.samplers.window import AdaptWindow
    eng = PlanningEngine(20, 6, (32,), n_sets=3)
    M, K = 5, 2
    win = AdaptWindow(eng, K, M)
    rng = np.random.RandomState(0)
    with pytest.raises(RuntimeError, match="normalisation"):
        win.gather()
    norm = _norm(5)
    win.set_normalization(norm)
    obs_hist, act_hist = [], []
    for t in range(23):                                        # wraps the M+1 = 6 slot ring several times
        o, a = rng.normal(size=(K, 20)), rng.uniform(-1, 1, size=(K, 6))
        if t < M + 1:
            assert win.length(0) == t and not win.ready()
            with pytest.raises(RuntimeError, match="running path"):
                win.gather()
        win.push(o, a)
        obs_hist.append(o)
        act_hist.append(a)
        if t >= M:
            x, tg = win.gather()
            oh, ah = np.stack(obs_hist, 1), np.stack(act_hist, 1)          # [K, t+1, .]
            ow, aw, nw = oh[:, -M - 1:-1], ah[:, -M - 1:-1], oh[:, -M:]
            np.testing.assert_array_equal(x.cpu().numpy(), np.concatenate(
                [O.normalize(ow, *norm["obs"]), O.normalize(aw, *norm["act"])], axis=-1).astype(np.float32))
            np.testing.assert_array_equal(tg.cpu().numpy(), O.normalize(nw - ow, *norm["delta"]).astype(np.float32))
    assert win.ready() and win.length(1) == 23
    win.reset(1)                                               # env 1's path ended
    assert win.length(1) == 0 and win.length(0) == 23
    with pytest.raises(RuntimeError, match="env 1 has 0 transitions"):
        win.gather()
    win.reset()
    assert win.length(0) == 0 and not win.ready()
    with pytest.raises(RuntimeError, match="out of range"):
        win.length(5)
    win.close()
    eng.close()


def test_window_model_dimension_mismatch_is_rejected():
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.engine import PlanningEngine
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.samplers.window import AdaptWindow
    model = MetaMLPDynamicsModel("dyn", SyntheticEnv("ant"), hidden_sizes=(64, 64), meta_batch_size=2, seed=0)
    prob = O.make_problem("ant", hidden_sizes=(64, 64), n_sets=1, m=2, seed=2)
    model.set_normalization(prob["norm"])
    other = PlanningEngine(20, 6, (32,), n_sets=1)
    win = AdaptWindow(other, 2, 3)
    for _ in range(5):
        win.push(np.zeros((2, 20)), np.zeros((2, 6)))
    with pytest.raises((RuntimeError, AssertionError)):
        model.adapt_from_window(win)
    big = model.make_adapt_window(3, 3)
    with pytest.raises(ValueError, match="meta_batch_size"):
        model.adapt_from_window(big)
