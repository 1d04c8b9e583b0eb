"""RNNMPCController (ReBAL) on the B200 engine: same constructor and methods as
learning_to_adapt/policies/rnn_mpc_controller.py:7-195.  Random shooting is one fused kernel call
(``l2a_rnn_rollout``) + one ``l2a_rnn_predict`` to advance the real hidden state (:63); the hidden state is kept per env
across calls and zeroed for finished episodes by ``reset(dones)`` (:139-163)."""
import os

import numpy as np
import torch

from learning_to_adapt_b200.envs.synthetic import reward_kind_of
from learning_to_adapt_b200.policies.base import Policy
from learning_to_adapt_b200.utils.serializable import Serializable


class RNNMPCController(Policy, Serializable):
    def __init__(self, name, env, dynamics_model, reward_model=None, discount=1, use_cem=False, n_candidates=1024, horizon=10,
                 num_cem_iters=8, percent_elites=0.05, use_reward_model=False, sampler=None):
        self.dynamics_model = dynamics_model
        self.reward_model = reward_model
        self.discount = discount
        self.n_candidates = n_candidates
        self.horizon = horizon
        self.use_cem = use_cem
        self.num_cem_iters = num_cem_iters
        self.percent_elites = percent_elites
        self.env = env
        self.use_reward_model = use_reward_model
        self._hidden_state = None
        self.sampler = sampler or os.environ.get("L2A_B200_SAMPLER", "numpy")
        self.unwrapped_env = env
        while hasattr(self.unwrapped_env, "wrapped_env"):
            self.unwrapped_env = self.unwrapped_env.wrapped_env
        if use_reward_model:
            raise NotImplementedError("use_reward_model=True: learned reward models are not on the fused path")
        assert hasattr(self.unwrapped_env, "reward"), "env must have a reward function"
        if not hasattr(dynamics_model, "rollout") or not getattr(dynamics_model, "recurrent", False):
            raise TypeError("dynamics_model must be learning_to_adapt_b200's RNNDynamicsModel; there is no CPU fallback")
        self._reward_kind, self._dt = reward_kind_of(self.unwrapped_env)
        self.last_plan = None
        Serializable.quick_init(self, locals())
        super(RNNMPCController, self).__init__(env=env)

    @property
    def vectorized(self):
        return True

    def get_action(self, observation):
        if observation.ndim == 1:
            observation = observation[None]
        action = self.get_actions(observation)[0]            # rnn_mpc_controller.py:47-55: a [1, A] array
        return action, dict()

    def get_actions(self, observations):
        actions = self.get_cem_action(observations) if self.use_cem else self.get_rs_action(observations)
        _, self._hidden_state = self.dynamics_model.predict(np.array(observations), actions, self._hidden_state)   # :63
        return actions, dict()

    def get_random_action(self, n):
        return np.random.uniform(low=self.action_space.low, high=self.action_space.high, size=(n,) + self.action_space.low.shape)

    def get_rs_action(self, observations):
        observations = np.asarray(observations, np.float64)
        n, m, h = self.n_candidates, len(observations), self.horizon
        dm = self.dynamics_model
        if self._hidden_state is None:
            self.reset(dones=[True] * m)
        a_host = None
        if self.sampler == "numpy":
            a_host = self.get_random_action(h * n * m).reshape((h, n * m, -1))          # :116
            a_dev = dm._f32(a_host)
        else:
            low, high = dm._f32(self.action_space.low), dm._f32(self.action_space.high)
            a_dev = torch.rand((h, n * m, low.shape[0]), device=dm.device, dtype=torch.float32) * (high - low) + low
        res = dm.rollout(dm._f32(observations), self._hidden_state, a_dev, n, h, self._reward_kind, self._dt, discount=self.discount)
        self.last_plan = res
        if a_host is not None:
            best = res["best_idx"].cpu().numpy()
            return a_host[0].reshape((m, n, -1))[range(m), best]                         # :118, :134
        return res["best_act"].cpu().numpy().astype(np.float64)

    def get_cem_action(self, observations):
        """CEM for the recurrent planner (rnn_mpc_controller.py:71-110): as MPCController's, bug-compatible elite mask included
        (:106), but the mean is replaced, not smoothed (:107; alpha = 0) and percent_elites defaults to 0.05."""
        import ctypes as C
        from learning_to_adapt_b200 import _native as N
        from learning_to_adapt_b200.engine import _ptr, _stream
        observations = np.asarray(observations, np.float64)
        n, m, h = self.n_candidates, len(observations), self.horizon
        dm = self.dynamics_model
        if self._hidden_state is None:
            self.reset(dones=[True] * m)
        act_dim = self.action_space.shape[0]
        ha = h * act_dim
        num_elites = max(int(n * self.percent_elites), 1)                                 # :78
        mean = torch.zeros((m, ha), device=dm.device, dtype=torch.float64)
        std = torch.ones((m, ha), device=dm.device, dtype=torch.float64)
        clip_low = dm._f32(np.concatenate([self.action_space.low] * h))
        clip_high = dm._f32(np.concatenate([self.action_space.high] * h))
        obs_dev = dm._f32(observations)
        rank = torch.empty(m, n, device=dm.device, dtype=torch.int32)
        res = None
        for _ in range(self.num_cem_iters):
            if self.sampler == "numpy":
                z = dm._f32(np.random.normal(size=(n, m, ha)))                            # :85
            else:
                z = torch.randn((n, m, ha), device=dm.device, dtype=torch.float32)
            samples = torch.empty(n, m, ha, device=dm.device, dtype=torch.float32)
            clipped = torch.empty_like(samples)
            N.check(dm.lib.l2a_cem_sample(dm._ctx, _ptr(z), _ptr(mean), _ptr(std), _ptr(clip_low), _ptr(clip_high), int(n), int(m),
                                          int(ha), _ptr(samples), _ptr(clipped), _stream()))
            res = dm.rollout(obs_dev, self._hidden_state, samples, n, h, self._reward_kind, self._dt, discount=self.discount,
                             want_returns=True, layout="nmha")
            N.check(dm.lib.l2a_cem_refit(dm._ctx, _ptr(res["returns"]), _ptr(clipped), int(n), int(m), int(ha), int(num_elites),
                                         0.0, 1, _ptr(rank), _ptr(mean), _ptr(std), _stream()))
        self.last_plan = res
        return res["best_act"].cpu().numpy().astype(np.float64)

    def get_params_internal(self, **tags):
        return []

    def reset(self, dones=None):
        """Zero the hidden state of finished episodes (rnn_mpc_controller.py:139-163)."""
        if dones is None:
            dones = [True]
        dones = np.asarray(dones, bool)
        if self._hidden_state is None:
            self._hidden_state = self.dynamics_model.get_initial_hidden(batch_size=len(dones))
        zero = self.dynamics_model.get_initial_hidden(batch_size=1)
        self._hidden_state.c[dones] = zero.c
        self._hidden_state.h[dones] = zero.h
