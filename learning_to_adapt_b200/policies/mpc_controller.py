"""MPCController on the B200 engine: same constructor and methods as
learning_to_adapt/policies/mpc_controller.py:6-135.

The reference's H-iteration python loop (one TF ``sess.run`` + numpy reward per step, :116-127) collapses into ONE
fused kernel call per planning step (random shooting) or per CEM iteration: sample -> K1 rollout/reward/argmax ->
(CEM: rank + refit).  The dynamics model must be one of this package's engine-backed models; there is no CPU
fallback.

Candidate sampling
  sampler="numpy" (default): the reference's own draw from the global numpy MT19937 stream (``np.random.uniform`` :67-69,114),
      regenerated bit-exactly ON THE DEVICE from ``np.random.get_state()`` inside the one host-buffer C call (l2a_plan_run_ex:
      H2D obs + generator state -> MT19937 -> K1 -> D2H action + advanced state, replayed as a CUDA graph);
      ``np.random.set_state()`` leaves the global stream where the reference would have left it -> identical candidates, hence
      identical chosen actions for a fixed ``np.random.seed``, at device speed.  CEM likewise: ``np.random.normal`` (:85, numpy's
      legacy polar method with its cached second value) is regenerated on the device and all iterations run inside one C call.
  sampler="numpy_host": the same stream drawn by numpy on the host and uploaded as fp32 (the round-1 parity path; kept as the
      cross-check of the device generator).
  sampler="device": Philox draws on the GPU (throughput mode; env var L2A_B200_SAMPLER=device selects it without touching
      the run scripts); CEM draws with torch.randn.
"""
import os

import numpy as np
import torch

from learning_to_adapt_b200 import _native as N
from learning_to_adapt_b200.envs.synthetic import reward_kind_of
from learning_to_adapt_b200.policies.base import Policy
from learning_to_adapt_b200.utils.serializable import Serializable


class _HostArray(np.ndarray):
    """Host result that also answers the tensor-style accessors of the device-tensor results (``.cpu().numpy()``)."""

    def cpu(self):
        return self

    def numpy(self):
        return np.asarray(self)


def _host(a):
    return np.asarray(a).view(_HostArray)


class MPCController(Policy, Serializable):
    def __init__(self, name, env, dynamics_model, reward_model=None, discount=1, use_cem=False, n_candidates=1024,
                 horizon=10, num_cem_iters=8, percent_elites=0.1, use_reward_model=False, alpha=0.1, sampler=None,
                 cem_compat=True, kernel=N.KERNEL_AUTO, parallel=None, seed=0):
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
        self.alpha = alpha
        self.sampler = sampler or os.environ.get("L2A_B200_SAMPLER", "numpy")
        assert self.sampler in ("numpy", "numpy_host", "device")
        self.cem_compat = cem_compat
        self.kernel = kernel
        self.parallel = parallel          # optional learning_to_adapt_b200.parallel.CandidateShard
        self.seed = seed                  # Philox seed of sampler="device"

        self.unwrapped_env = env
        while hasattr(self.unwrapped_env, "wrapped_env"):
            self.unwrapped_env = self.unwrapped_env.wrapped_env
        # make sure that env has reward function (mpc_controller.py:38-39)
        assert hasattr(self.unwrapped_env, "reward"), "env must have a reward function"
        if use_reward_model:
            raise NotImplementedError("use_reward_model=True: learned reward models are not on the fused path")
        if not hasattr(dynamics_model, "_engine"):
            raise TypeError("dynamics_model must be an engine-backed learning_to_adapt_b200 model "
                            "(MLPDynamicsModel / MetaMLPDynamicsModel); there is no CPU fallback")
        self._reward_kind, self._dt = reward_kind_of(self.unwrapped_env)
        self.last_plan = None             # results of the most recent planning call (diagnostics / tests)
        self.keep_returns = False         # CEM: also fetch the last iteration's per-candidate returns into last_plan (tests)
        self.push_window = None           # AdaptWindow the planning call appends (obs, action) to (set by the Sampler)

        Serializable.quick_init(self, locals())
        super(MPCController, self).__init__(env=env)

    @property
    def vectorized(self):
        return True

    def get_action(self, observation):
        if observation.ndim == 1:
            observation = observation[None]
        if self.use_cem:
            action = self.get_cem_action(observation)
        else:
            action = self.get_rs_action(observation)
        return action, dict()

    def get_actions(self, observations):
        if self.use_cem:
            actions = self.get_cem_action(observations)
        else:
            actions = self.get_rs_action(observations)
        return actions, dict()

    def get_random_action(self, n):
        return np.random.uniform(low=self.action_space.low, high=self.action_space.high,
                                 size=(n,) + self.action_space.low.shape)

    def _grbal_step_args(self):
        """GrBAL env step folded into the planning call (samplers/sampler.py:81-91): a deferred adapt_from_window() of the model
        runs in front of the planner, and (obs, chosen action) is appended to the Sampler's window behind it."""
        model = self.dynamics_model
        pending = getattr(model, "_pending_window", None)
        window = pending if pending is not None else self.push_window
        if window is None:
            return dict()
        assert pending is None or self.push_window is None or pending is self.push_window
        flags = (N.PLAN_ADAPT if pending is not None else 0) | (N.PLAN_PUSH if self.push_window is not None else 0)
        if pending is not None and window._norm_src is not model.normalization:
            window.set_normalization(model.normalization)
        return dict(window=window, flags=flags, inner_lr=float(getattr(model, "inner_learning_rate", 0.0)))

    def _grbal_step_done(self, kw):
        if kw.get("flags", 0) & N.PLAN_ADAPT:
            self.dynamics_model._pending_window = None
        self.pushed_window = bool(kw.get("flags", 0) & N.PLAN_PUSH)      # the Sampler skips its own push when the call did it

    # ------------------------------------------------------------------ random shooting (mpc_controller.py:108-129)
    def get_rs_action(self, observations):
        observations = np.asarray(observations, np.float64)
        n, m, h = self.n_candidates, len(observations), self.horizon
        eng = self.dynamics_model._engine
        act_dim = self.action_space.shape[0]
        set_mode, first_set, n_sets = self.dynamics_model.planning_sets(m)
        self.pushed_window = False
        if self.sampler in ("device", "numpy"):
            step = self._grbal_step_args()
            shard = None
            if self.parallel is not None:
                shard = dict(rank=self.parallel.rank, world=self.parallel.world_size, all_gather=self.parallel.all_gather_bytes)
            acts, ret, idx = eng.plan_rs_host(observations, n, h, self._reward_kind, self._dt, self.action_space.low,
                                              self.action_space.high, discount=self.discount, set_mode=set_mode,
                                              first_set=first_set, n_sets=n_sets, kernel=self.kernel, seed=self.seed,
                                              sampler="philox" if self.sampler == "device" else "mt19937", shard=shard, **step)
            self._grbal_step_done(step)
            self.last_plan = dict(best_ret=_host(ret), best_idx=_host(idx), best_act=_host(acts), returns=None)    # host arrays
            return acts
        self.dynamics_model._flush_pending_adapt() if hasattr(self.dynamics_model, "_flush_pending_adapt") else None
        obs_dev = eng._f32(observations)
        if self.parallel is not None:
            return self.parallel.plan_rs(self, observations, obs_dev, set_mode, first_set, n_sets)
        # parity mode: the reference's own draw from the global numpy stream, uploaded as float32
        a_host = self.get_random_action(h * n * m).reshape((h, n * m, -1))                # :114
        a_dev = eng._f32(a_host)
        res = eng.rollout(obs_dev, a_dev, n, h, self._reward_kind, self._dt, discount=self.discount, set_mode=set_mode,
                          first_set=first_set, n_sets=n_sets, layout="thra", want_returns=False, kernel=self.kernel)
        self.last_plan = res
        best = res["best_idx"].cpu().numpy()
        cand_a = a_host[0].reshape((m, n, -1))                                            # :118
        return cand_a[range(m), best]                                                     # :129, float64

    # ------------------------------------------------------------------ CEM (mpc_controller.py:71-106)
    def get_cem_action(self, observations):
        observations = np.asarray(observations, np.float64)
        n, m, h = self.n_candidates, len(observations), self.horizon
        eng = self.dynamics_model._engine
        act_dim = self.action_space.shape[0]
        ha = h * act_dim
        set_mode, first_set, n_sets = self.dynamics_model.planning_sets(m)
        num_elites = max(int(self.n_candidates * self.percent_elites), 1)                 # :78
        if not self.cem_compat and m > 1:
            raise NotImplementedError("cem_compat=False is defined for one env per call: for m > 1 the reference's own sample -> env "
                                      "layout is inconsistent (mpc_controller.py:85-102) and only the bug-compatible mode reproduces it")
        self.pushed_window = False
        if self.sampler in ("numpy", "device") and self.parallel is None:
            # all iterations in ONE host-buffer C call (l2a_plan_run_ex, CEM planner), replayed as a CUDA graph
            step = self._grbal_step_args()
            acts, ret, idx, mean, std = eng.plan_cem_host(
                observations, n, h, self._reward_kind, self._dt, self.action_space.low, self.action_space.high,
                self.num_cem_iters, num_elites, self.alpha, discount=self.discount, set_mode=set_mode, first_set=first_set,
                n_sets=n_sets, kernel=self.kernel, seed=self.seed, sampler="philox" if self.sampler == "device" else "mt19937",
                compat=self.cem_compat, **step)
            self._grbal_step_done(step)
            returns = _host(eng.last_plan_returns(m, n)) if self.keep_returns else None
            self.last_plan = dict(best_ret=_host(ret), best_idx=_host(idx), best_act=_host(acts), returns=returns)
            self.last_cem_state = (_host(mean), _host(std))
            return acts
        self.dynamics_model._flush_pending_adapt() if hasattr(self.dynamics_model, "_flush_pending_adapt") else None
        mean = torch.zeros((m, ha), device=eng.device, dtype=torch.float64)               # :79
        std = torch.ones((m, ha), device=eng.device, dtype=torch.float64)                 # :80
        clip_low = eng._f32(np.concatenate([self.action_space.low] * h))                  # :81
        clip_high = eng._f32(np.concatenate([self.action_space.high] * h))                # :82
        obs_dev = eng._f32(observations)
        res = None
        for _ in range(self.num_cem_iters):                                               # :84
            if self.sampler in ("numpy", "numpy_host"):
                z = eng._f32(np.random.normal(size=(n, m, ha)))                           # :85
            else:
                z = torch.randn((n, m, ha), device=eng.device, dtype=torch.float32)
            samples, clipped = eng.cem_sample(z, mean, std, clip_low, clip_high)          # :86-87
            # rows of the (n*m, H, A) view are rolled out UNclipped, row r from env r // n  (:88-99)
            res = eng.rollout(obs_dev, samples, n, h, self._reward_kind, self._dt, discount=self.discount,
                              set_mode=set_mode, first_set=first_set, n_sets=n_sets, layout="nmha", want_returns=True,
                              kernel=self.kernel)
            eng.cem_refit(res["returns"], clipped, num_elites, self.alpha, mean, std, compat=self.cem_compat)  # :101-104
        self.last_plan = res
        self.last_cem_state = (mean, std)
        return res["best_act"].cpu().numpy().astype(np.float64)                           # :106

    def get_params_internal(self, **tags):
        return []

    def reset(self, dones=None):
        pass
