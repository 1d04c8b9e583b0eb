"""AdaptWindow: host handle of the device-resident adaptation window (include/l2a_b200.h, l2a_window_*).

Holds, per env, a ring of the last M+1 (observation, action) pairs of the running path in HBM, so that GrBAL's per-step
`adapt` windows (samplers/sampler.py:82-90) and their float64 normalisation (dynamics/meta_mlp_dynamics.py:334-339) are formed
by a kernel in front of K2 instead of list slicing + np.stack + normalise + upload on the host every env step.
"""
import ctypes as C

import numpy as np
import torch

from .. import _native as N
from ..engine import _stream


class AdaptWindow(object):
    def __init__(self, engine, n_envs, adapt_batch_size):
        self._engine = engine                     # keeps the context alive
        self.lib = engine.lib
        self.n_envs, self.M = int(n_envs), int(adapt_batch_size)
        self.obs_dim, self.act_dim = engine.obs_dim, engine.act_dim
        self._h = C.c_void_p()
        N.check(self.lib.l2a_window_create(engine._ctx, self.n_envs, self.M, self.obs_dim, self.act_dim, C.byref(self._h)))
        self._norm_src = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h and self._engine._ctx:
            self.lib.l2a_window_destroy(self._engine._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_normalization(self, normalization):
        """normalization: the reference's dict {'obs': (mean, std), 'act': ..., 'delta': ...} (mlp_dynamics.py:253-262)."""
        arrs = [np.ascontiguousarray(np.asarray(normalization[k][i], np.float64)) for k in ("obs", "act", "delta") for i in (0, 1)]
        assert arrs[0].shape == (self.obs_dim,) and arrs[2].shape == (self.act_dim,) and arrs[4].shape == (self.obs_dim,)
        N.check(self.lib.l2a_window_set_normalization(self._engine._ctx, self._h, *[C.c_void_p(a.ctypes.data) for a in arrs],
                                                      _stream()))
        self._norm_src = normalization

    def push(self, observations, actions):
        """Append one (observation, action) pair per env (sampler.py:109-110)."""
        obs = np.ascontiguousarray(np.asarray(observations, np.float64).reshape(self.n_envs, self.obs_dim))
        act = np.ascontiguousarray(np.asarray(actions, np.float64).reshape(self.n_envs, self.act_dim))
        both = torch.from_numpy(np.concatenate([obs.ravel(), act.ravel()])).to(self._engine.device)
        N.check(self.lib.l2a_window_push(self._engine._ctx, self._h, C.c_void_p(both.data_ptr()),
                                         C.c_void_p(both.data_ptr() + obs.size * 8), _stream()))
        self._last = both                          # alive until the next push (stream-ordered reuse by the allocator)

    def gather(self):
        """Normalised windows (x [n_envs, M, D+A], target [n_envs, M, D]) as device float32 tensors: what adapt() feeds K2."""
        x = torch.empty(self.n_envs, self.M, self.obs_dim + self.act_dim, device=self._engine.device, dtype=torch.float32)
        target = torch.empty(self.n_envs, self.M, self.obs_dim, device=self._engine.device, dtype=torch.float32)
        N.check(self.lib.l2a_window_gather(self._engine._ctx, self._h, C.c_void_p(x.data_ptr()), C.c_void_p(target.data_ptr()),
                                           _stream()))
        return x, target

    def reset(self, env=None):
        """The running path of `env` ended (sampler.py:128); None: every env."""
        N.check(self.lib.l2a_window_reset(self._engine._ctx, self._h, -1 if env is None else int(env), _stream()))

    def length(self, env=0):
        n = self.lib.l2a_window_length(self._engine._ctx, self._h, int(env))
        if n < 0:
            N.check(n)
        return n

    def ready_for_adapt(self):
        """Every env holds at least M + 1 pairs (what the gather kernel needs)."""
        return all(self.length(e) >= self.M + 1 for e in range(self.n_envs))

    def ready(self):
        """The reference's trigger: `len(running_paths[0]['observations']) > adapt_batch_size + 1` (sampler.py:82)."""
        return self.length(0) > self.M + 1
