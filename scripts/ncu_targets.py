"""Workload for the ncu captures under profiles/ (run under `ncu ... python scripts/ncu_targets.py [what]`): a few calls of
every host-buffer planning path so that the launch list shows each kernel of this library with its duration.
  what = headline | cfg4 | grbal | all (default)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel  # noqa: E402
from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel  # noqa: E402
from learning_to_adapt_b200.envs.synthetic import SyntheticEnv  # noqa: E402
from learning_to_adapt_b200.policies.mpc_controller import MPCController  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
os.environ.setdefault("L2A_NO_GRAPH", "1")          # individual launches: ncu names every kernel

if what in ("headline", "all"):
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512, 512), n_sets=5, m=1, seed=0)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(512, 512, 512), ensemble_size=5)
    for e, p in enumerate(prob["param_sets"]):
        model.set_params(p, member=e)
    model.set_normalization(prob["norm"])
    for sampler in ("device", "numpy"):
        ctrl = MPCController("policy", env, model, n_candidates=2000, horizon=20, sampler=sampler)
        np.random.seed(0)
        for _ in range(3):
            ctrl.get_actions(prob["obs0"])

if what in ("cfg4", "all"):
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=1, m=1, seed=0)
    env = SyntheticEnv("half_cheetah")
    model = MLPDynamicsModel("dyn", env, hidden_sizes=(512, 512))
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    for sampler in ("device", "numpy"):
        ctrl = MPCController("policy", env, model, use_cem=True, n_candidates=5000, horizon=30, num_cem_iters=3, percent_elites=0.1,
                             alpha=0.1, sampler=sampler)
        np.random.seed(0)
        for _ in range(2):
            ctrl.get_actions(prob["obs0"])

if what in ("grbal", "all"):
    K, M = 5, 16
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512, 512), n_sets=1, m=K, seed=0)
    env = SyntheticEnv("half_cheetah")
    model = MetaMLPDynamicsModel("dyn", env, hidden_sizes=(512, 512, 512), meta_batch_size=10, inner_learning_rate=1e-3)
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctx = O.make_adapt_context(4, prob, K, M)
    win = model.make_adapt_window(K, M)
    for j in range(M + 2):
        win.push(np.stack([c[min(j, M - 1)] for c in ctx[0]]), np.stack([c[min(j, M - 1)] for c in ctx[1]]))
    ctrl = MPCController("policy", env, model, n_candidates=1000, horizon=15, sampler="device")
    ctrl.push_window = win
    for _ in range(3):
        model.switch_to_pre_adapt()
        model.adapt_from_window(win)
        ctrl.get_actions(prob["obs0"])
print("ncu targets done:", what)
