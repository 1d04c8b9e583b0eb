"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures produced by the
reference's own code.  Tolerance: 1e-4 relative on returns / next states (BASELINE.json north_star); chosen actions
identical unless the oracle's top-2 gap is inside that tolerance."""
import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O
from tests.helpers import RTOL, assert_argmax_consistent, assert_returns_close, dev, make_engine

pytestmark = pytest.mark.gpu

N = None


def _native():
    global N
    if N is None:
        from learning_to_adapt_b200 import _native as n
        N = n
    return N


# ------------------------------------------------------------------------------------------------ tcgen05 tile
@pytest.mark.parametrize("n,k", [(16, 64), (32, 128), (80, 64), (80, 256), (128, 128)])
def test_umma_tile_matches_fp32_matmul(n, k):
    from learning_to_adapt_b200.engine import PlanningEngine
    eng = PlanningEngine(20, 6, (128,), n_sets=1, debug=True)       # l2a_debug_umma_tile lives in the debug build
    rng = np.random.RandomState(n + k)
    A = rng.normal(size=(128, k)).astype(np.float32)
    B = rng.normal(size=(n, k)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = eng.debug_umma_tile(dev(A), dev(B), variant=0).cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 2e-5, "split-bf16 tile error %.3e" % err
    got8 = eng.debug_umma_tile(dev(A), dev(B), variant=8).cpu().numpy()     # accumulator-fragment TMEM loads (16x256b)
    np.testing.assert_array_equal(got8, got)
    if k <= 128:
        got16 = eng.debug_umma_tile(dev(A), dev(B), variant=16).cpu().numpy()  # A operand via tcgen05.cp -> TMEM (TS-mode MMA)
        np.testing.assert_array_equal(got16, got)
    got1 = eng.debug_umma_tile(dev(A), dev(B), variant=4).cpu().numpy()     # single bf16 pass: must be visibly worse
    err1 = np.abs(got1 - want).max() / np.abs(want).max()
    assert 1e-4 < err1 < 3e-2, "single-pass bf16 error %.3e" % err1


# ------------------------------------------------------------------------------------------------ K4 predict
def test_predict_matches_golden_reference_choreography(golden):
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=3)
    model = MLPDynamicsModel("dyn", SyntheticEnv("half_cheetah"), hidden_sizes=(32, 32))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    out = model.predict(golden["predict_obs"], golden["predict_act"])
    assert out.dtype == np.float64 and out.shape == golden["predict_out"].shape
    np.testing.assert_allclose(out, golden["predict_out"], rtol=RTOL, atol=1e-5)
    got = model.get_params()
    for k, v in prob["param_sets"][0].items():
        np.testing.assert_array_equal(got[k], v)


@pytest.mark.parametrize("env,hidden,n", [("ant", (512, 512, 512), 100), ("half_cheetah", (64, 200), 33), ("arm_7dof", (128,), 7)])
def test_predict_matches_oracle(env, hidden, n):
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=3, m=1, seed=4, out_scale=1.0)
    eng = make_engine(prob)
    rng = np.random.RandomState(0)
    obs = prob["norm"]["obs"][0] + prob["norm"]["obs"][1] * rng.normal(size=(3 * n, prob["obs_dim"]))
    act = rng.uniform(prob["low"], prob["high"], size=(3 * n, prob["act_dim"]))
    nat = _native()
    d = eng.predict_delta(dev(obs), dev(act), nat.SETS_SHARED, 1, 1).cpu().numpy()
    want = O.predict(obs, act, prob["param_sets"][1], prob["norm"]) - obs
    np.testing.assert_allclose(d, want, rtol=RTOL, atol=1e-5 * np.abs(want).max())
    d = eng.predict_delta(dev(obs), dev(act), nat.SETS_PER_ENV, 0, 3).cpu().numpy()
    want = O.predict_per_task(obs, act, prob["param_sets"], prob["norm"]) - obs
    np.testing.assert_allclose(d, want, rtol=RTOL, atol=1e-5 * np.abs(want).max())
    d = eng.predict_delta(dev(obs), dev(act), nat.SETS_ENSEMBLE_MEAN, 0, 3).cpu().numpy()
    want = O.predict_ensemble_mean(obs, act, prob["param_sets"], prob["norm"]) - obs
    np.testing.assert_allclose(d, want, rtol=RTOL, atol=1e-5 * np.abs(want).max())


# ------------------------------------------------------------------------------------------------ K1 vs golden (reference planner)
def _rollout(eng, prob, actions, n, h, discount=1.0, set_mode=0, first_set=0, n_sets=1, kernel=0):
    res = eng.rollout(dev(prob["obs0"]), dev(actions), n, h, prob["reward_kind"], prob["dt"], discount=discount,
                      set_mode=set_mode, first_set=first_set, n_sets=n_sets, kernel=kernel)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in res.items()}


@pytest.mark.parametrize("tag,env,hidden", [
    ("rs_hc", "half_cheetah", (32, 32)),
    ("rs_hc_disc", "half_cheetah", (48,)),
    ("rs_ant", "ant", (32, 32, 32)),
    ("rs_arm", "arm_7dof", (32, 32)),
])
def test_rollout_matches_reference_planner_golden(golden, tag, env, hidden):
    n, h, m, seed = [int(v) for v in golden[tag + "_meta"]]
    discount = float(golden[tag + "_discount"])
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=1, m=m, seed=seed)
    eng = make_engine(prob)
    actions = golden[tag + "_actions"]
    res = _rollout(eng, prob, actions, n, h, discount)
    ref_returns = np.zeros(n * m)
    for t in range(h):
        ref_returns += discount ** t * golden[tag + "_step_rewards"][t]
    ref_returns = ref_returns.reshape(m, n)
    assert_returns_close(res["returns"], ref_returns)
    assert_argmax_consistent(res["best_idx"], ref_returns)
    best = res["best_idx"]
    np.testing.assert_allclose(res["best_ret"], res["returns"][range(m), best], rtol=0, atol=0)
    cand = actions[0].reshape(m, n, -1)
    np.testing.assert_allclose(res["best_act"], cand[range(m), best].astype(np.float32), rtol=0, atol=0)
    if all(int(b) == int(w) for b, w in zip(best, np.argmax(ref_returns, axis=1))):
        np.testing.assert_allclose(res["best_act"], golden[tag + "_chosen"], rtol=1e-6)


def test_controller_drop_in_reproduces_reference_actions(golden):
    """np.random.seed + MPCController.get_actions == the reference controller's chosen actions (bit-identical float64)."""
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    n, h, m, seed = [int(v) for v in golden["rs_hc_meta"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=m, seed=seed)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(32, 32))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctrl = MPCController("policy", env, model, n_candidates=n, horizon=h)
    np.random.seed(seed + 100)
    acts, info = ctrl.get_actions(prob["obs0"])
    assert info == {} and acts.dtype == np.float64
    np.testing.assert_array_equal(acts, golden["rs_hc_chosen"])
    # get_action on a single observation returns a 2-D [1, A] action (mpc_controller.py:48-57)
    prob1 = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=12)
    model.set_params(prob1["param_sets"][0])
    model.set_normalization(prob1["norm"])
    ctrl1 = MPCController("policy", env, model, n_candidates=20, horizon=3)
    np.random.seed(112)
    a1, _ = ctrl1.get_action(prob1["obs0"][0])
    assert a1.shape == tuple(golden["get_action_shape"])
    np.testing.assert_array_equal(a1, golden["get_action_out"])


# ------------------------------------------------------------------------------------------------ K1 both kernels vs oracle
CASES = [
    # env, hidden, N, H, m, n_sets, mode
    ("half_cheetah", (512, 512), 500, 10, 1, 1, "shared"),            # BASELINE cfg1
    ("half_cheetah", (512, 512, 512), 200, 15, 3, 3, "per_env"),      # cfg2 reading (i), reduced N
    ("half_cheetah", (512, 512, 512), 333, 8, 1, 5, "ensemble"),      # headline reading (ii), ragged N
    ("ant", (512, 512, 512), 150, 6, 2, 1, "shared"),                 # Ant dims (D=41, A=8, +-150 actions)
    ("ant", (256, 128), 90, 5, 1, 2, "ensemble"),
    ("arm_7dof", (128,), 50, 4, 2, 1, "shared"),
]
MODE = {"shared": 0, "per_env": 1, "ensemble": 2}


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("env,hidden,n,h,m,n_sets,mode", CASES)
def test_rollout_matches_oracle(env, hidden, n, h, m, n_sets, mode, kernel):
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=n_sets, m=m, seed=5)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(17, prob["low"], prob["high"], h, n * m)
    want = O.rollout_returns(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"],
                             0.97, mode)
    res = _rollout(eng, prob, actions, n, h, 0.97, MODE[mode], 0, n_sets, kernel)
    assert_returns_close(res["returns"], want)
    assert_argmax_consistent(res["best_idx"], want)
    best = res["best_idx"]
    np.testing.assert_array_equal(res["best_ret"], res["returns"][range(m), best])
    np.testing.assert_array_equal(res["best_act"], actions[0].reshape(m, n, -1)[range(m), best].astype(np.float32))


# edge shapes: single candidate / single step, exact and just-over tile boundaries, many nearly-empty envs, the largest
# supported ensemble (cluster of 8), discount 0
EDGE_CASES = [
    # hidden, N, H, m, n_sets, mode, discount
    ((128,), 1, 1, 1, 1, "shared", 1.0),
    ((128, 128), 80, 3, 1, 1, "shared", 1.0),
    ((128, 128), 81, 2, 2, 2, "per_env", 0.9),
    ((128,), 3, 4, 37, 1, "shared", 1.0),
    ((128,), 40, 2, 1, 8, "ensemble", 1.0),
    ((256, 256), 33, 5, 1, 1, "shared", 0.0),
]


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("hidden,n,h,m,n_sets,mode,discount", EDGE_CASES)
def test_rollout_edge_shapes_match_oracle(hidden, n, h, m, n_sets, mode, discount, kernel):
    prob = O.make_problem("half_cheetah", hidden_sizes=hidden, n_sets=n_sets, m=m, seed=31)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(5, prob["low"], prob["high"], h, n * m)
    want = O.rollout_returns(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"],
                             discount, mode)
    res = _rollout(eng, prob, actions, n, h, discount, MODE[mode], 0, n_sets, kernel)
    assert_returns_close(res["returns"], want)
    assert_argmax_consistent(res["best_idx"], want)
    best = res["best_idx"]
    np.testing.assert_array_equal(res["best_act"], actions[0].reshape(m, n, -1)[range(m), best].astype(np.float32))


@pytest.mark.parametrize("reward_kind", [0, 1, 2])
def test_rollout_maximum_dimensions_of_the_tensor_core_variant(reward_kind):
    """obs_dim 48 and act_dim 16 (the limits of the tcgen05 variant: state in registers, [obs | act] in one 64-wide chunk),
    every reward family, an ensemble of 3 -- tcgen05 and SIMT kernels against the oracle."""
    D, A, hidden, n_sets, m, n, h = 48, 16, (256, 128), 3, 2, 70, 4
    low, high = -2.0 * np.ones(A), 2.0 * np.ones(A)
    sets = [O.xavier_params(np.random.RandomState(77 + e), D + A, hidden, D, out_scale=0.1) for e in range(n_sets)]
    norm = O.make_normalization(np.random.RandomState(78), D, A, low, high)
    obs0 = norm["obs"][0] + norm["obs"][1] * np.random.RandomState(79).normal(size=(m, D))
    prob = dict(obs_dim=D, act_dim=A, low=low, high=high, dt=0.02, reward_kind=reward_kind, param_sets=sets, norm=norm,
                obs0=obs0, hidden_sizes=hidden)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(6, low, high, h, n * m)
    want = O.rollout_returns(obs0, actions, sets, norm, reward_kind, 0.02, 0.99, "ensemble")
    for kernel in (1, 2):
        res = _rollout(eng, prob, actions, n, h, 0.99, 2, 0, n_sets, kernel)
        assert_returns_close(res["returns"], want)
        assert_argmax_consistent(res["best_idx"], want)


def test_rollout_kernels_agree_at_headline_size():
    """BASELINE headline (HC, N=2000, H=20, E=5): the tcgen05 and fp32 SIMT kernels agree within tolerance, and
    rolling the candidates in a permuted order permutes the returns (size-independent properties)."""
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512, 512), n_sets=5, m=1, seed=0)
    eng = make_engine(prob)
    n, h = 2000, 20
    actions = O.sample_rs_actions(3, prob["low"], prob["high"], h, n)
    tc = _rollout(eng, prob, actions, n, h, 1.0, 2, 0, 5, 2)
    simt = _rollout(eng, prob, actions, n, h, 1.0, 2, 0, 5, 1)
    assert_returns_close(tc["returns"], simt["returns"])
    assert_argmax_consistent(tc["best_idx"], simt["returns"])
    perm = np.random.RandomState(0).permutation(n)
    tc_p = _rollout(eng, prob, actions[:, perm], n, h, 1.0, 2, 0, 5, 2)
    np.testing.assert_allclose(tc_p["returns"][0], tc["returns"][0][perm], rtol=1e-6, atol=1e-5)
    assert int(perm[int(tc_p["best_idx"][0])]) == int(tc["best_idx"][0])
    # oracle on a sample of the candidates
    sub = np.arange(0, n, 25)
    want = O.rollout_returns(prob["obs0"], actions[:, sub], prob["param_sets"], prob["norm"], prob["reward_kind"],
                             prob["dt"], 1.0, "ensemble")
    assert_returns_close(tc["returns"][:, sub], want)


def test_argmax_tie_and_nan_semantics():
    """np.argmax: first maximum wins; NaN beats everything (first NaN wins)."""
    prob = O.make_problem("half_cheetah", hidden_sizes=(32,), n_sets=1, m=2, seed=2)
    eng = make_engine(prob)
    n, h = 70, 2
    actions = O.sample_rs_actions(1, prob["low"], prob["high"], h, n * 2)
    actions[:, 5] = actions[:, 40]          # env 0: candidates 5 and 40 identical -> identical returns
    actions[:, 70 + 9] = np.nan             # env 1: candidate 9 produces NaN
    actions[:, 70 + 30] = np.nan
    res = _rollout(eng, prob, actions, n, h)
    r = res["returns"]
    assert r[0, 5] == r[0, 40]
    assert int(res["best_idx"][0]) == int(np.argmax(r[0]))
    assert np.isnan(r[1, 9]) and int(res["best_idx"][1]) == 9 == int(np.argmax(r[1]))


# ------------------------------------------------------------------------------------------------ K2 adapt
def test_adapt_matches_golden_and_oracle(golden):
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    K, M, MBS, n_per = [int(v) for v in golden["adapt_meta"]]
    lr = float(golden["adapt_lr"])
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32, 32), n_sets=1, m=3, seed=31)
    theta = prob["param_sets"][0]
    model = MetaMLPDynamicsModel("dyn", SyntheticEnv("half_cheetah"), hidden_sizes=(32, 32, 32), meta_batch_size=MBS,
                                 inner_learning_rate=lr)
    model.set_params(theta)
    model.set_normalization(prob["norm"])
    ctx = O.make_adapt_context(41, prob, K, M)
    model.adapt(*ctx)
    for k in range(K):
        got = model.get_adapted_params(k)
        for key in theta.keys():
            ref = golden["adapt_theta%d_%s" % (k, key.replace("/", "."))]
            upd = np.abs(ref - theta[key]).max()
            np.testing.assert_allclose(got[key], ref, rtol=1e-5, atol=1e-4 * upd + 1e-9)
    post = model.predict(golden["adapt_query_obs"], golden["adapt_query_act"])
    np.testing.assert_allclose(post, golden["adapt_post_predict"], rtol=RTOL, atol=1e-5)
    # theta untouched; switch_to_pre_adapt goes back to it (meta_mlp_dynamics.py:347-351)
    np.testing.assert_array_equal(model.get_params()["hidden_0/kernel"], theta["hidden_0/kernel"])
    model.switch_to_pre_adapt()
    pre = model.predict(golden["adapt_query_obs"], golden["adapt_query_act"])
    np.testing.assert_allclose(pre, golden["adapt_pre_predict"], rtol=RTOL, atol=1e-5)


def test_grbal_planning_matches_reference_golden(golden):
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    n, h, m, seed = [int(v) for v in golden["grbal_rs_meta"]]
    K, M, MBS, n_per = [int(v) for v in golden["adapt_meta"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32, 32), n_sets=1, m=m, seed=seed)
    env = SyntheticEnv("half_cheetah")
    model = MetaMLPDynamicsModel("dyn", env, hidden_sizes=(32, 32, 32), meta_batch_size=MBS,
                                 inner_learning_rate=float(golden["adapt_lr"]))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    model.switch_to_pre_adapt()
    model.adapt(*O.make_adapt_context(41, prob, K, M))
    ctrl = MPCController("policy", env, model, n_candidates=n, horizon=h)
    np.random.seed(161)
    acts, _ = ctrl.get_actions(prob["obs0"])
    np.testing.assert_array_equal(acts, golden["grbal_rs_chosen"])


def test_adapt_full_size_matches_oracle():
    """BASELINE cfg2/3 shapes: K=5 tasks, M=16, (512,512,512), lr=1e-3."""
    prob = O.make_problem("ant", hidden_sizes=(512, 512, 512), n_sets=1, m=5, seed=7, out_scale=1.0)
    eng = make_engine(prob, n_sets=6)
    theta = prob["param_sets"][0]
    ctx = O.make_adapt_context(4, prob, 5, 16)
    lr = 1e-2
    want = O.adapt(*ctx, theta, prob["norm"], lr)
    xs, ts = [], []
    for o, a, nx in zip(*ctx):
        xs.append(np.concatenate([O.normalize(o, *prob["norm"]["obs"]), O.normalize(a, *prob["norm"]["act"])], axis=1))
        ts.append(O.normalize(nx - o, *prob["norm"]["delta"]))
    eng.adapt(dev(np.stack(xs)), dev(np.stack(ts)), lr, 0, 1)
    for k in range(5):
        got = eng.get_params(1 + k)
        for key in theta.keys():
            upd = np.abs(want[k][key] - theta[key]).max()
            np.testing.assert_allclose(got[key], want[k][key], rtol=1e-5, atol=2e-4 * upd + 1e-9)
    # and the adapted sets drive the tensor-core rollout
    n, h = 96, 5
    actions = O.sample_rs_actions(9, prob["low"], prob["high"], h, n * 5)
    ref = O.rollout_returns(prob["obs0"], actions, want, prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "per_env")
    res = _rollout(eng, prob, actions, n, h, 1.0, 1, 1, 5, 2)
    assert_returns_close(res["returns"], ref)


@pytest.mark.parametrize("env,hidden,K,M", [("ant", (128, 128), 2, 6), ("ant", (128, 128), 3, 5), ("arm_7dof", (64, 200), 4, 10),
                                            ("ant", (512, 512, 512), 2, 6), ("half_cheetah", (128,), 3, 7)])
def test_adapt_odd_dims_and_context_lengths(env, hidden, K, M):
    """obs dims that are not multiples of 4 (Ant 41, Arm 17) with M not a multiple of 4 and K >= 2: the per-task workspace
    blocks must stay 16-byte aligned for the vectorised loads of the backward chain (a misaligned float4 load faults)."""
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=1, m=K, seed=19, out_scale=1.0)
    eng = make_engine(prob, n_sets=1 + K)
    theta = prob["param_sets"][0]
    ctx = O.make_adapt_context(8, prob, K, M)
    lr = 1e-2
    want = O.adapt(*ctx, theta, prob["norm"], lr)
    xs, ts = [], []
    for o, a, nx in zip(*ctx):
        xs.append(np.concatenate([O.normalize(o, *prob["norm"]["obs"]), O.normalize(a, *prob["norm"]["act"])], axis=1))
        ts.append(O.normalize(nx - o, *prob["norm"]["delta"]))
    eng.adapt(dev(np.stack(xs)), dev(np.stack(ts)), lr, 0, 1)
    torch.cuda.synchronize()
    for k in range(K):
        got = eng.get_params(1 + k)
        for key in theta.keys():
            upd = np.abs(want[k][key] - theta[key]).max()
            np.testing.assert_allclose(got[key], want[k][key], rtol=1e-5, atol=2e-4 * upd + 1e-9)


# ------------------------------------------------------------------------------------------------ K1c CEM
@pytest.mark.parametrize("tag", ["cem_m1", "cem_m2"])
def test_cem_bug_compatible_matches_reference_golden(golden, tag):
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.policies.mpc_controller import MPCController
    n, h, m, iters, seed = [int(v) for v in golden[tag + "_meta"]]
    pct, alpha = [float(v) for v in golden[tag + "_pct_alpha"]]
    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=m, seed=seed)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(32, 32))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctrl = MPCController("policy", env, model, use_cem=True, n_candidates=n, horizon=h, num_cem_iters=iters,
                         percent_elites=pct, alpha=alpha)
    ctrl.keep_returns = True
    np.random.seed(seed + 100)
    acts, _ = ctrl.get_actions(prob["obs0"])
    ref_returns = golden[tag + "_last_step_rewards"].sum(axis=0).reshape(m, n)
    got_returns = ctrl.last_plan["returns"].cpu().numpy()
    assert_returns_close(got_returns, ref_returns, rtol=5e-4)      # returns after `iters` refits of fp32-rounded samples
    assert_argmax_consistent(ctrl.last_plan["best_idx"].cpu().numpy(), ref_returns, rtol=5e-4)
    if np.array_equal(ctrl.last_plan["best_idx"].cpu().numpy(), np.argmax(ref_returns, axis=1)):
        np.testing.assert_allclose(acts, golden[tag + "_chosen"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("compat", [True, False])
def test_cem_refit_matches_oracle(compat):
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=1, seed=3)
    eng = make_engine(prob)
    n, h, iters, pct, alpha = 300, 6, 3, 0.1, 0.1
    ha = h * prob["act_dim"]
    rng = np.random.RandomState(5)
    zs = [rng.normal(size=(n, 1, ha)) for _ in range(iters)]
    chosen, best, returns, mean, std = O.cem_plan(prob["obs0"], zs, prob["low"], prob["high"], prob["param_sets"],
                                                  prob["norm"], prob["reward_kind"], prob["dt"], h, pct, alpha,
                                                  corrected=not compat)
    d_mean = torch.zeros((1, ha), device="cuda", dtype=torch.float64)
    d_std = torch.ones((1, ha), device="cuda", dtype=torch.float64)
    lo, hi = dev(np.concatenate([prob["low"]] * h)), dev(np.concatenate([prob["high"]] * h))
    for z in zs:
        samples, clipped = eng.cem_sample(dev(z), d_mean, d_std, lo, hi)
        res = eng.rollout(dev(prob["obs0"]), samples, n, h, prob["reward_kind"], prob["dt"], layout="nmha")
        eng.cem_refit(res["returns"], clipped, max(int(n * pct), 1), alpha, d_mean, d_std, compat=compat)
    np.testing.assert_allclose(d_mean.cpu().numpy(), mean, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(d_std.cpu().numpy()[0], std, rtol=2e-3, atol=2e-3)
    assert_returns_close(res["returns"].cpu().numpy(), returns, rtol=2e-3)


# ------------------------------------------------------------------------------------------------ K3 shard glue
def test_shard_pack_select_kernels_match_host_logic():
    """The two library kernels around the NCCL all-gather agree with the torch restatement (parallel.pack_best /
    select_best, which the gloo CPU tests pin to np.argmax semantics), including ties, NaN and indices >= 2**24."""
    from learning_to_adapt_b200.engine import PlanningEngine
    from learning_to_adapt_b200.parallel import pack_best, select_best
    eng = PlanningEngine(20, 6, (32,), n_sets=1)
    rng = np.random.RandomState(0)
    G, m, A = 4, 7, 6
    packs_k, packs_t = [], []
    for g in range(G):
        ret = rng.normal(size=m).astype(np.float32)
        idx = rng.randint(0, 1000, size=m).astype(np.int32)
        act = rng.normal(size=(m, A)).astype(np.float32)
        if g in (1, 3):
            ret[2] = 5.0                       # tie across ranks
            ret[4] = np.nan                    # NaN on two ranks: lowest global index wins
        off = g * (2 ** 24 + 7)
        pk = eng.shard_pack(dev(ret), torch.as_tensor(idx, device="cuda"), dev(act), off)
        pt = pack_best(torch.tensor(ret), torch.tensor(idx.astype(np.int64) + off), torch.tensor(act))
        np.testing.assert_array_equal(pk.cpu().numpy(), pt.numpy())
        packs_k.append(pk)
        packs_t.append(pt)
    r_k, i_k, a_k = eng.shard_select(torch.stack(packs_k).contiguous())
    r_t, i_t, a_t = select_best(torch.stack(packs_t))
    np.testing.assert_array_equal(i_k.cpu().numpy(), i_t.numpy())
    np.testing.assert_array_equal(a_k.cpu().numpy(), a_t.numpy())
    np.testing.assert_array_equal(np.isnan(r_k.cpu().numpy()), np.isnan(r_t.numpy()))
    ok = ~np.isnan(r_t.numpy())
    np.testing.assert_array_equal(r_k.cpu().numpy()[ok], r_t.numpy()[ok])
