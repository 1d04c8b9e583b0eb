"""In-process vectorised env stepping with the reference executor's contract
(samplers/vectorized_env_executor.py:7-77, IterativeEnvExecutor): `num_rollouts` deep copies of the env, auto-reset on
`done` or after `max_path_length` steps (the observation returned for a finished env is the first one of its next path).
The multi-process ParallelEnvExecutor is CPU env plumbing outside this build's scope (out of scope)."""
import copy

import numpy as np


class IterativeEnvExecutor(object):
    def __init__(self, env, num_rollouts, max_path_length):
        self._num_envs = int(num_rollouts)
        self.envs = [copy.deepcopy(env) for _ in range(self._num_envs)]
        self.ts = np.zeros(self._num_envs, dtype=int)
        self.max_path_length = max_path_length

    @property
    def num_envs(self):
        return self._num_envs

    def reset(self):
        self.ts[:] = 0
        return [env.reset() for env in self.envs]

    def step(self, actions):
        assert len(actions) == self._num_envs
        obs, rewards, dones, env_infos = [], [], [], []
        for env, a in zip(self.envs, actions):
            o, r, d, info = env.step(a)
            obs.append(o)
            rewards.append(r)
            dones.append(d)
            env_infos.append(info)
        self.ts += 1
        dones = np.logical_or(self.ts >= self.max_path_length, np.asarray(dones))
        for i in np.flatnonzero(dones):
            obs[i] = self.envs[i].reset()
            self.ts[i] = 0
        return obs, rewards, dones, env_infos
