"""GPU parity of the CTA-pair rollout kernel (tcgen05.mma.cta_group::2, csrc/rollout_tc2.cuh) through the C ABI: against the
CPU oracle at small and BASELINE shapes (1e-4 relative per return, north_star) and against the single-CTA tcgen05 kernel
(the two run the same split-bf16 arithmetic, so their returns agree far inside the parity tolerance).

Covers every weight-set mode (shared / per-env / ensemble mean through the L2 flag exchange), all three candidates-per-CTA
instances (72 / 48 / 32), both state-register instances (obs <= 24 / <= 48), packed and unpacked layer 0, 256- and 512-wide
hidden layers, ragged last tiles (peer CTA without candidates), and the three reward families."""
import os

import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O
from learning_to_adapt_b200 import _native as N
from tests.helpers import assert_argmax_consistent, assert_returns_close, dev, error_report, make_engine

pytestmark = pytest.mark.gpu

PAIR, TC, SIMT = N.KERNEL_TCGEN05_PAIR, N.KERNEL_TCGEN05, N.KERNEL_SIMT
MODES = {"shared": N.SETS_SHARED, "per_env": N.SETS_PER_ENV, "ensemble": N.SETS_ENSEMBLE_MEAN}


def _run(eng, prob, actions, n, h, mode, n_sets, kernel, discount=1.0):
    res = eng.rollout(dev(prob["obs0"]), dev(actions), n, h, prob["reward_kind"], prob["dt"], discount=discount,
                      set_mode=MODES[mode], first_set=0, n_sets=n_sets, kernel=kernel)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in res.items()}


def _with_nc(nc):
    class _Ctx(object):
        def __enter__(self):
            self.old = os.environ.get("L2A_TC2_NC")
            if nc:
                os.environ["L2A_TC2_NC"] = str(nc)
            elif self.old is not None:
                del os.environ["L2A_TC2_NC"]

        def __exit__(self, *a):
            if self.old is None:
                os.environ.pop("L2A_TC2_NC", None)
            else:
                os.environ["L2A_TC2_NC"] = self.old
    return _Ctx()


@pytest.mark.parametrize("env,hidden,mode,n_sets,m,n,h,nc", [
    ("half_cheetah", (256, 256), "shared", 1, 1, 64, 3, 32),          # one full tile, packed layer 0, one M-block per layer
    ("half_cheetah", (512, 512), "shared", 1, 1, 100, 4, 32),         # two tiles, ragged peer (100 = 64 + 32 + 4)
    ("half_cheetah", (512, 512), "shared", 1, 2, 150, 5, 48),         # two envs, NC = 48, peer CTA of the last tile nearly empty
    ("half_cheetah", (512, 512, 512), "shared", 1, 1, 200, 6, 72),    # NC = 72 (a 16-candidate block straddles the two CTAs)
    ("half_cheetah", (512, 256), "shared", 1, 1, 145, 4, 72),         # mixed widths; the peer of the 2nd tile has 1 candidate
    ("ant", (512, 512), "shared", 1, 1, 130, 4, 48),                  # unpacked layer 0, obs 41 (48-wide state instance)
    ("arm_7dof", (256, 512), "shared", 1, 1, 90, 5, 32),
    ("half_cheetah", (512, 512, 512), "per_env", 3, 3, 120, 5, 48),   # GrBAL: env k uses set k
    ("half_cheetah", (512, 512), "ensemble", 5, 1, 150, 6, 72),       # ensemble mean of 5 pairs through the L2 flag exchange
    ("ant", (512, 512, 512), "ensemble", 5, 1, 100, 5, 32),
    ("half_cheetah", (256, 256), "ensemble", 8, 2, 70, 4, 32),        # 8 members, two envs
    ("half_cheetah", (512, 512), "ensemble", 3, 1, 1, 1, 32),         # one candidate, one step
])
def test_pair_kernel_vs_oracle(env, hidden, mode, n_sets, m, n, h, nc):
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=n_sets, m=m, seed=3)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(9, prob["low"], prob["high"], h, n * m)
    want = O.rollout_returns(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 0.97, mode)
    with _with_nc(nc):
        got = _run(eng, prob, actions, n, h, mode, n_sets, PAIR, discount=0.97)
    rep = assert_returns_close(got["returns"], want)
    assert_argmax_consistent(got["best_idx"], want)
    np.testing.assert_array_equal(got["best_ret"], got["returns"][range(m), got["best_idx"]])
    print("pair vs oracle", env, hidden, mode, "nc", nc, rep)


@pytest.mark.parametrize("cfg", ["headline", "cfg1", "cfg2", "cfg3", "cfg4"])
def test_pair_kernel_at_baseline_configs(cfg):
    """Full BASELINE sizes with the automatic tile choice: pair kernel == single-CTA kernel within 2e-5 relative on every return
    (same arithmetic, different accumulation grouping), and vs the oracle on a candidate stride."""
    env, hidden, mode, n_sets, m, n, h = {
        "headline": ("half_cheetah", (512, 512, 512), "ensemble", 5, 1, 2000, 20),
        "cfg1": ("half_cheetah", (512, 512), "shared", 1, 1, 500, 10),
        "cfg2": ("half_cheetah", (512, 512, 512), "per_env", 5, 5, 1000, 15),
        "cfg3": ("ant", (512, 512, 512), "ensemble", 5, 1, 2000, 20),
        "cfg4": ("half_cheetah", (512, 512), "shared", 1, 1, 5000, 30),
    }[cfg]
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=n_sets, m=m, seed=17)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(23, prob["low"], prob["high"], h, n * m)
    with _with_nc(0):
        got = _run(eng, prob, actions, n, h, mode, n_sets, PAIR)
    ref = _run(eng, prob, actions, n, h, mode, n_sets, TC)
    rep = error_report(got["returns"], ref["returns"])
    print("pair vs single-CTA", cfg, rep)
    assert rep["rel"] <= 2e-5, rep
    sub = np.arange(0, n, max(1, n // 40))
    rows = np.concatenate([e * n + sub for e in range(m)])
    want = O.rollout_returns(prob["obs0"], actions[:, rows], prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, mode)
    rep = assert_returns_close(got["returns"][:, sub], want)
    print("pair vs oracle", cfg, rep)
    np.testing.assert_array_equal(got["best_idx"], np.argmax(got["returns"], axis=1))


def test_pair_kernel_repeatable_and_unsupported_shape():
    """Two launches give bit-identical returns (the flag exchange leaves no state behind); a 128-wide layer is refused loudly."""
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=5, m=1, seed=5)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(2, prob["low"], prob["high"], 8, 300)
    a = _run(eng, prob, actions, 300, 8, "ensemble", 5, PAIR)
    b = _run(eng, prob, actions, 300, 8, "ensemble", 5, PAIR)
    np.testing.assert_array_equal(a["returns"], b["returns"])
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=1, seed=5)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(2, prob["low"], prob["high"], 3, 40)
    with pytest.raises(Exception, match="256"):
        _run(eng, prob, actions, 40, 3, "shared", 1, PAIR)


# ------------------------------------------------------------------------------------------------ the host-buffer planning calls
@pytest.mark.parametrize("graph", [True, False])
def test_pair_kernel_inside_the_host_buffer_plan_call(graph, monkeypatch):
    """l2a_plan_run_ex with a pair-eligible model (AUTO -> CTA-pair kernel, including the flag reset of the member exchange inside
    the captured graph): every call's choice equals the oracle's argmax over the candidates that call drew."""
    if not graph:
        monkeypatch.setenv("L2A_NO_GRAPH", "1")
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=5, m=2, seed=21)
    eng = make_engine(prob)
    n, h = 200, 6
    for call in range(4):
        obs = prob["obs0"] + 0.01 * call
        acts, ret, idx = eng.plan_rs_host(obs, n, h, prob["reward_kind"], prob["dt"], prob["low"], prob["high"],
                                          discount=0.95, set_mode=2, first_set=0, n_sets=5, seed=7)
        cand = eng.last_plan_candidates()
        np.testing.assert_array_equal(cand, O.sample_rs_actions_device(7, call, prob["low"], prob["high"], h, 2 * n))
        want = O.rollout_returns(obs.astype(np.float32).astype(np.float64), cand.astype(np.float64), prob["param_sets"],
                                 prob["norm"], prob["reward_kind"], prob["dt"], 0.95, "ensemble")
        assert_argmax_consistent(idx, want)
        assert_returns_close(ret, want[range(2), idx])
        np.testing.assert_array_equal(acts, cand[0].reshape(2, n, -1)[range(2), idx].astype(np.float64))
        assert eng.last_plan_uses_graph() == (graph and call >= 1)


def test_pair_kernel_behind_the_drop_in_controller_numpy_stream():
    """MPCController.get_actions with the package's default sampler (numpy's MT19937 stream regenerated on the device) on a
    512-wide model = the CTA-pair kernel: the chosen actions are the oracle's choice on numpy's own draw, bit for bit."""
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=1, m=3, seed=31)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(512, 512))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctrl = MPCController("policy", env, model, n_candidates=250, horizon=7)
    np.random.seed(12)
    for call in range(2):
        acts, _ = ctrl.get_actions(prob["obs0"])
    np.random.seed(12)
    for call in range(2):
        cand = np.random.uniform(prob["low"], prob["high"], size=(7 * 250 * 3, prob["act_dim"])).reshape(7, 750, -1)
        want_act, best, _ = O.rs_plan(prob["obs0"], cand, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"])
    np.testing.assert_array_equal(acts, want_act)


def test_pair_kernel_grbal_loop_device_window_equals_host_lists():
    """The whole GrBAL sampling loop (window gather -> K2 adapt -> re-tile of BOTH blobs -> K1 per-env sets -> window push in one
    call) on a 256-wide model, i.e. through the pair kernel: same actions and bit-identical adapted weights as the host lists."""
    from tests.test_gpu_window import _grbal
    paths_d, adapted_d, prob = _grbal("device", hidden=(256, 256))
    paths_l, adapted_l, _ = _grbal("lists", hidden=(256, 256))
    for pd, pl in zip(paths_d, paths_l):
        np.testing.assert_array_equal(pd["actions"], pl["actions"])
    for ad, al in zip(adapted_d, adapted_l):
        for k in ad:
            np.testing.assert_array_equal(ad[k], al[k])


def test_pair_kernel_cem_one_call_teacher_forced():
    """CEM in one host-buffer call on a 512-wide model, ONE iteration, bug-compatible mode, numpy's normal stream: the call's
    per-candidate returns agree with the oracle on the same samples, and its refit (mean, std) is exactly the reference's rule
    (mpc_controller.py:101-104) applied to the call's own returns."""
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    n, h, pct, alpha = 1200, 12, 0.1, 0.1
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=1, m=1, seed=16)
    A = prob["act_dim"]
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(512, 512))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctrl = MPCController("policy", env, model, use_cem=True, n_candidates=n, horizon=h, num_cem_iters=1, percent_elites=pct, alpha=alpha)
    ctrl.keep_returns = True
    np.random.seed(28)
    acts, _ = ctrl.get_actions(prob["obs0"])
    got = ctrl.last_plan["returns"]
    z = np.random.RandomState(28).normal(size=(n, 1, h * A))                              # mean 0, std 1: the samples are z itself
    a_roll = np.transpose(z.astype(np.float32).astype(np.float64).reshape(n, h, A), (1, 0, 2))
    want = O.rollout_returns(prob["obs0"], a_roll, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "shared")
    assert_returns_close(got, want)
    k = max(int(n * pct), 1)
    order = np.argsort(-np.asarray(got, np.float64), axis=-1, kind="stable")
    mask = (order < k).T
    clipped = np.clip(z, np.concatenate([prob["low"]] * h), np.concatenate([prob["high"]] * h))
    elites = clipped[mask]
    np.testing.assert_allclose(ctrl.last_cem_state[0][0], (1 - alpha) * elites.mean(axis=0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ctrl.last_cem_state[1][0], elites.std(axis=0), rtol=1e-9, atol=1e-12)
    best = int(np.argmax(got[0]))
    np.testing.assert_array_equal(acts[0], z[best, 0, :A])                              # :106: first action of the best (unclipped) sample
