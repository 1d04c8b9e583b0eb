"""Sampler: the caller of the planning hot path (samplers/sampler.py:11-143), with the per-env adaptation windows held on
the device.

Same constructor and `obtain_samples(log, log_prefix, random)` contract as the reference: steps `num_rollouts` vectorised
envs until `num_rollouts * max_path_length` samples are collected and returns the list of finished paths (dicts of stacked
observations / actions / rewards / dones / env_infos / agent_infos).  With `adapt_batch_size = M` and a meta-learned model,
every env step after the running path of env 0 is longer than M+1 first re-adapts the model on each env's last M transitions
(:82-90).  `window='device'` (default) keeps those transitions in an AdaptWindow ring in HBM and runs
window-gather -> K2 -> K1 back to back on the stream; `window='lists'` is the reference's host formulation (list slicing,
np.stack, host normalisation, upload) kept for parity tests -- both feed K2 bit-identical inputs.
"""
import time

import numpy as np

from .vectorized_env_executor import IterativeEnvExecutor
from .window import AdaptWindow


def _empty_running_path():
    return dict(observations=[], actions=[], rewards=[], dones=[], env_infos=[], agent_infos=[])


def _stack_dicts(dicts):
    keys = dicts[0].keys() if dicts else []
    return {k: np.asarray([d[k] for d in dicts]) for k in keys}


class Sampler(object):
    def __init__(self, env, policy, num_rollouts, max_path_length, n_parallel=1, adapt_batch_size=None, window="device"):
        assert hasattr(policy, "get_actions") and hasattr(policy, "reset")
        assert window in ("device", "lists")
        self.env, self.policy = env, policy
        self.max_path_length = max_path_length
        self.total_samples = num_rollouts * max_path_length
        self.n_parallel = n_parallel              # envs are stepped in-process whatever n_parallel says
        self.total_timesteps_sampled = 0
        self.adapt_batch_size = adapt_batch_size
        self.vec_env = IterativeEnvExecutor(env, num_rollouts, max_path_length)
        self.window_mode = window
        self._window = None
        self.policy_time = self.env_time = 0.0

    def update_tasks(self):
        pass

    def _device_window(self):
        model = getattr(self.policy, "dynamics_model", None)
        if self.adapt_batch_size is None or self.window_mode != "device" or not hasattr(model, "adapt_from_window"):
            return None
        if self._window is None:
            self._window = model.make_adapt_window( self.vec_env.num_envs, self.adapt_batch_size)
        return self._window

    def obtain_samples(self, log=False, log_prefix="", random=False):
        paths = []
        n_samples = 0
        num_envs = self.vec_env.num_envs
        running = [_empty_running_path() for _ in range(num_envs)]
        policy, model = self.policy, getattr(self.policy, "dynamics_model", None)
        M = self.adapt_batch_size
        win = self._device_window()
        policy_time = env_time = 0.0
        policy.reset(dones=[True] * num_envs)
        if win is not None:
            win.reset()
        if hasattr(policy, "push_window"):
            policy.push_window = win                 # planning calls append (obs, action) to the window inside their own graph
        obses = np.asarray(self.vec_env.reset())
        while n_samples < self.total_samples:
            t = time.time()
            if random:
                actions = np.stack([self.env.action_space.sample() for _ in range(num_envs)], axis=0)
                agent_infos = {}
            else:
                if M is not None and len(running[0]["observations"]) > M + 1:
                    model.switch_to_pre_adapt()
                    if win is not None:
                        model.adapt_from_window(win)                       # gather + normalise on the device, then K2
                    else:
                        model.adapt([np.stack(p["observations"][-M - 1:-1]) for p in running],
                                    [np.stack(p["actions"][-M - 1:-1]) for p in running],
                                    [np.stack(p["observations"][-M:]) for p in running])
                actions, agent_infos = policy.get_actions(obses)
            policy_time += time.time() - t

            t = time.time()
            next_obses, rewards, dones, env_infos = self.vec_env.step(actions)
            env_time += time.time() - t
            if not env_infos:
                env_infos = [dict() for _ in range(num_envs)]
            if not agent_infos:
                agent_infos = [dict() for _ in range(num_envs)]

            if win is not None and not (not random and getattr(policy, "pushed_window", False)):
                win.push(obses, actions)
            new_samples = 0
            for idx in range(num_envs):
                rp = running[idx]
                reward = rewards[idx]
                if isinstance(reward, np.ndarray):
                    reward = reward[0]
                rp["observations"].append(obses[idx])
                rp["actions"].append(actions[idx])
                rp["rewards"].append(reward)
                rp["dones"].append(dones[idx])
                rp["env_infos"].append(env_infos[idx])
                rp["agent_infos"].append(agent_infos[idx])
                if dones[idx]:
                    paths.append(dict(observations=np.asarray(rp["observations"]), actions=np.asarray(rp["actions"]),
                                      rewards=np.asarray(rp["rewards"]), dones=np.asarray(rp["dones"]),
                                      env_infos=_stack_dicts(rp["env_infos"]), agent_infos=_stack_dicts(rp["agent_infos"])))
                    new_samples += len(rp["rewards"])
                    running[idx] = _empty_running_path()
                    if win is not None:
                        win.reset(idx)
            n_samples += new_samples
            obses = np.asarray(next_obses)

        self.total_timesteps_sampled += self.total_samples
        self.policy_time, self.env_time = policy_time, env_time
        if log:
            print("%sPolicyExecTime %.3f  %sEnvExecTime %.3f" % (log_prefix, policy_time, log_prefix, env_time))
        return paths
