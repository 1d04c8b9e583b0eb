#!/usr/bin/env python
"""End-to-end GrBAL loop on a MuJoCo-free stand-in environment, written exactly as run_scripts/run_grbal.py wires the
reference's objects (dynamics model -> MPC controller -> sampler loop -> fit), with this repo's B200-backed classes.

The "environment" is a random stable linear system with a per-episode perturbation (the role the crippled leg plays in the
paper): it only serves to produce transitions; the planner, the one-step adaptation and the dynamics predict calls all run on the
fused CUDA kernels.  Needs a B200.      python examples/grbal_synthetic.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel  # noqa: E402
from learning_to_adapt_b200.envs.synthetic import SyntheticEnv  # noqa: E402
from learning_to_adapt_b200.policies.mpc_controller import MPCController  # noqa: E402


class LinearWorld(object):
    """obs' = obs + dt * (obs @ A_task + act @ B); obs[-3] plays the torso x-position the HalfCheetah reward differentiates."""

    def __init__(self, spec, seed=0):
        self.spec, self.rng = spec, np.random.RandomState(seed)
        D, A = spec.observation_space.shape[0], spec.action_space.shape[0]
        self.A0 = -0.5 * np.eye(D) + 0.1 * self.rng.normal(size=(D, D))
        self.B = 0.5 * self.rng.normal(size=(A, D))
        self.B[:, -3] += 1.0                                    # actions push the "torso" forward
        self.reset_task()

    def reset_task(self):
        self.A = self.A0 + 0.05 * self.rng.normal(size=self.A0.shape)   # the perturbation the model must adapt to

    def reset(self):
        self.obs = 0.1 * self.rng.normal(size=self.A.shape[0])
        return self.obs.copy()

    def step(self, act):
        nxt = self.obs + self.spec.dt * 5.0 * (self.obs @ self.A + act @ self.B)
        r = self.spec.reward(self.obs[None], act[None], nxt[None])[0]
        self.obs = nxt
        return nxt.copy(), r


def collect(envs, policy, model, steps, adapt_batch_size, random=False):
    """The body of Sampler.obtain_samples (samplers/sampler.py:73-127) for `len(envs)` vectorised envs."""
    paths = [dict(observations=[], actions=[], rewards=[]) for _ in envs]
    obses = np.stack([e.reset() for e in envs])
    t_policy = 0.0
    for _ in range(steps):
        t0 = time.time()
        if random:
            actions = np.stack([envs[0].spec.action_space.sample() for _ in envs])
        else:
            M = adapt_batch_size
            if M is not None and len(paths[0]["observations"]) > M + 1:
                model.switch_to_pre_adapt()
                model.adapt([np.stack(p["observations"][-M - 1:-1]) for p in paths],
                            [np.stack(p["actions"][-M - 1:-1]) for p in paths],
                            [np.stack(p["observations"][-M:]) for p in paths])
            actions, _ = policy.get_actions(obses)
        t_policy += time.time() - t0
        nxt, rew = zip(*[e.step(a) for e, a in zip(envs, actions)])
        for p, o, a, r in zip(paths, obses, actions, rew):
            p["observations"].append(o)
            p["actions"].append(a)
            p["rewards"].append(r)
        obses = np.stack(nxt)
    return paths, t_policy


def main():
    spec = SyntheticEnv("half_cheetah")
    num_envs, path_len, M = 5, 120, 16
    model = MetaMLPDynamicsModel("dyn_model", spec, hidden_sizes=(512, 512, 512), meta_batch_size=10, inner_learning_rate=1e-3,
                                 batch_size=M, learning_rate=1e-3, seed=0)
    policy = MPCController("policy", spec, model, n_candidates=500, horizon=10, sampler="device")
    envs = [LinearWorld(spec, seed=i) for i in range(num_envs)]
    for itr in range(3):
        for e in envs:
            e.reset_task()
        paths, t_policy = collect(envs, policy, model, path_len, M, random=(itr == 0))
        obs = np.stack([np.stack(p["observations"][:-1]) for p in paths])
        act = np.stack([np.stack(p["actions"][:-1]) for p in paths])
        nxt = np.stack([np.stack(p["observations"][1:]) for p in paths])
        t0 = time.time()
        model.fit(obs, act, nxt, epochs=20 if itr == 0 else 8)
        print("itr %d  AverageReturn %8.2f  PolicyExecTime %.3fs (%.2f ms / env step incl. adapt)  Time-ModelFit %.1fs" %
              (itr, np.mean([np.sum(p["rewards"]) for p in paths]), t_policy, 1e3 * t_policy / path_len, time.time() - t0))


if __name__ == "__main__":
    main()
