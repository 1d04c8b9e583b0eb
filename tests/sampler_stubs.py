"""Deterministic stand-ins shared by tests/golden/make_golden_sampler.py (driving the reference's Sampler) and the sampler tests
(driving this repo's): a closed-form policy and a dynamics-model stub that records what the sampler passes to adapt()."""
import numpy as np

from learning_to_adapt_b200.envs.synthetic import LinearWorldEnv


def make_env():
    return LinearWorldEnv("half_cheetah", seed=7)


class RecordingModel(object):
    def __init__(self):
        self.obs, self.act, self.nxt, self.steps = [], [], [], []
        self.n_switch = 0
        self.step = 0

    def switch_to_pre_adapt(self):
        self.n_switch += 1

    def adapt(self, obs, act, obs_next):
        self.obs.append(np.stack([np.array(o, np.float64) for o in obs]))
        self.act.append(np.stack([np.array(a, np.float64) for a in act]))
        self.nxt.append(np.stack([np.array(o, np.float64) for o in obs_next]))
        self.steps.append(self.step)


class StubPolicy(object):
    """actions = clip(sin(3 obs[:, :A] + step), low, high): depends on the observation and the env step, no randomness."""

    def __init__(self, env, model):
        self.dynamics_model = model
        self.low, self.high = env.action_space.low, env.action_space.high
        self.A = env.action_space.shape[0]

    def reset(self, dones=None):
        pass

    def get_actions(self, observations):
        observations = np.asarray(observations)
        a = np.clip(np.sin(3.0 * observations[:, :self.A] + self.dynamics_model.step), self.low, self.high)
        self.dynamics_model.step += 1
        return a, {}
