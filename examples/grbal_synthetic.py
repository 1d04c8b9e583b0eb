#!/usr/bin/env python
"""End-to-end GrBAL loop on a MuJoCo-free stand-in environment, written exactly as run_scripts/run_grbal.py wires the
reference's objects (dynamics model -> MPC controller -> sampler loop -> fit), with this repo's B200-backed classes.

The "environment" (envs/synthetic.py:LinearWorldEnv) is a random stable linear system with a per-episode perturbation (the role the
crippled leg plays in the paper): it only serves to produce transitions; the sampler keeps each env's adaptation window on the device; the planner, the one-step adaptation and the dynamics predict calls all run on the
fused CUDA kernels.  Needs a B200.      python examples/grbal_synthetic.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel  # noqa: E402
from learning_to_adapt_b200.envs.synthetic import LinearWorldEnv  # noqa: E402
from learning_to_adapt_b200.policies.mpc_controller import MPCController  # noqa: E402
from learning_to_adapt_b200.samplers.sampler import Sampler  # noqa: E402


def main():
    env = LinearWorldEnv("half_cheetah", seed=0)
    num_envs, path_len, M = 5, 120, 16
    model = MetaMLPDynamicsModel("dyn_model", env, hidden_sizes=(512, 512, 512), meta_batch_size=10, inner_learning_rate=1e-3,
                                 batch_size=M, learning_rate=1e-3, seed=0)
    policy = MPCController("policy", env, model, n_candidates=500, horizon=10, sampler="device")
    sampler = Sampler(env, policy, num_rollouts=num_envs, max_path_length=path_len, adapt_batch_size=M)      # device window
    for i, e in enumerate(sampler.vec_env.envs):
        e.seed(100 + i)
    for itr in range(3):
        for e in sampler.vec_env.envs:
            e.reset_task()
        paths = sampler.obtain_samples(random=(itr == 0))
        obs = np.stack([p["observations"][:-1] for p in paths])
        act = np.stack([p["actions"][:-1] for p in paths])
        nxt = np.stack([p["observations"][1:] for p in paths])
        t0 = time.time()
        model.fit(obs, act, nxt, epochs=20 if itr == 0 else 8)
        print("itr %d  AverageReturn %8.2f  PolicyExecTime %.3fs (%.2f ms / env step incl. adapt)  Time-ModelFit %.1fs" %
              (itr, np.mean([np.sum(p["rewards"]) for p in paths]), sampler.policy_time, 1e3 * sampler.policy_time / path_len,
               time.time() - t0))


if __name__ == "__main__":
    main()
