"""MetaMLPDynamicsModel (GrBAL) on the B200 engine: same constructor and methods as
learning_to_adapt/dynamics/meta_mlp_dynamics.py:11-445.

Weight set 0 is the prior theta; sets 1..meta_batch_size hold the per-env adapted theta'_k produced by kernel K2.
``adapt`` never touches theta (meta_mlp_dynamics.py:321-345); ``switch_to_pre_adapt`` just forgets the adapted sets
(:347-351).  After ``adapt`` with K tasks, ``predict`` and the fused rollout step row chunk k through theta'_k
(:296-306) -- without re-uploading K x theta' every horizon step as the TF feed_dict did.
"""
import ctypes as C

import numpy as np
import torch

from learning_to_adapt_b200 import _native as N
from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel, normalize
from learning_to_adapt_b200.utils.serializable import Serializable


class MetaMLPDynamicsModel(MLPDynamicsModel):
    def __init__(self, name, env, hidden_sizes=(512, 512), meta_batch_size=10, hidden_nonlinearity="relu",
                 output_nonlinearity=None, batch_size=500, learning_rate=0.001, inner_learning_rate=0.1,
                 normalize_input=True, optimizer=None, valid_split_ratio=0.2, rolling_average_persitency=0.99,
                 device=0, seed=None):
        Serializable.quick_init(self, locals())
        self.meta_batch_size = int(meta_batch_size)
        self.inner_learning_rate = inner_learning_rate
        self._num_adapted_models = None
        self._adapted = False
        MLPDynamicsModel.__init__(self, name, env, hidden_sizes=hidden_sizes, hidden_nonlinearity=hidden_nonlinearity,
                                  output_nonlinearity=output_nonlinearity, batch_size=batch_size,
                                  learning_rate=learning_rate, normalize_input=normalize_input, optimizer=optimizer,
                                  valid_split_ratio=valid_split_ratio,
                                  rolling_average_persitency=rolling_average_persitency, ensemble_size=1, device=device,
                                  seed=seed)

    def _total_sets(self):
        return 1 + self.meta_batch_size

    def planning_sets(self, n_envs):
        if self._adapted:
            k = self._num_adapted_models
            if n_envs != k:
                raise NotImplementedError("adapted to %d tasks but planning for %d envs: the reference's row chunking "
                                          "(meta_mlp_dynamics.py:296-306) only lines up with envs when they are equal" % (k, n_envs))
            return N.SETS_PER_ENV, 1, k
        return N.SETS_SHARED, 0, 1

    def compute_normalization(self, obs, act, obs_next):
        """3-D inputs [tasks, T, dim], statistics over axes (0, 1) (meta_mlp_dynamics.py:396-407)."""
        if obs.ndim == 2:
            return MLPDynamicsModel.compute_normalization(self, obs, act, obs_next)
        D, A = obs.shape[-1], act.shape[-1]
        return MLPDynamicsModel.compute_normalization(self, obs.reshape(-1, D), act.reshape(-1, A), obs_next.reshape(-1, D))

    # ------------------------------------------------------------------ K2
    def adapt(self, obs, act, obs_next):
        """obs/act/obs_next: K lists of [M, .] arrays, the last M transitions per env (samplers/sampler.py:83-90)."""
        k = len(obs)
        assert len(obs) == len(act) == len(obs_next)
        if k > self.meta_batch_size:
            raise ValueError("adapt() got %d tasks but meta_batch_size is %d" % (k, self.meta_batch_size))
        obs = np.stack([np.asarray(o, np.float64) for o in obs])
        act = np.stack([np.asarray(a, np.float64) for a in act])
        obs_next = np.stack([np.asarray(o, np.float64) for o in obs_next])
        assert obs.ndim == 3 and obs.shape[2] == self.obs_space_dims
        assert act.ndim == 3 and act.shape[2] == self.action_space_dims
        assert obs_next.shape == obs.shape
        # host-side float64 normalisation, float32 feed (meta_mlp_dynamics.py:334-339)
        obs_n = normalize(obs, *self.normalization["obs"])
        act_n = normalize(act, *self.normalization["act"])
        delta_n = normalize(obs_next - obs, *self.normalization["delta"])
        eng = self._engine
        x = eng._f32(np.concatenate([obs_n, act_n], axis=2))
        target = eng._f32(delta_n)
        self._pending_window = None
        eng.adapt(x, target, self.inner_learning_rate, src_set=0, dst_first_set=1)
        self._num_adapted_models = k
        self._adapted = True

    def make_adapt_window(self, n_envs, adapt_batch_size):
        from learning_to_adapt_b200.samplers.window import AdaptWindow
        return AdaptWindow(self._engine, n_envs, adapt_batch_size)

    def adapt_from_window(self, window, defer=True):
        """adapt() on the windows held by a device-resident AdaptWindow (samplers/window.py): the gather of
        obs[-M-1:-1] / act[-M-1:-1] / obs[-M:] (samplers/sampler.py:83-88) and the float64 normalisation (:334-339) run in a
        kernel in front of K2; inputs to K2 are bit-identical to adapt()'s.
        defer=True (default): nothing is launched here -- the next MPCController planning call runs window-gather -> K2 ->
        re-tile in front of the planner inside its own C call / CUDA graph (l2a_plan_run_ex with L2A_PLAN_ADAPT); any other use
        of the adapted sets (predict, get_adapted_params) first runs the adaptation on the spot."""
        k = window.n_envs
        if k > self.meta_batch_size:
            raise ValueError("window holds %d envs but meta_batch_size is %d" % (k, self.meta_batch_size))
        if (window.obs_dim, window.act_dim) != (self._engine.obs_dim, self._engine.act_dim) or window._engine is not self._engine:
            raise RuntimeError("the adaptation window belongs to another model (dims %d/%d vs %d/%d)"
                               % (window.obs_dim, window.act_dim, self._engine.obs_dim, self._engine.act_dim))
        if not window.ready_for_adapt():
            raise RuntimeError("an env's running path is shorter than adapt_batch_size + 1 transitions")
        self._num_adapted_models = k
        self._adapted = True
        self._pending_window = window
        if not defer:
            self._flush_pending_adapt()

    def _flush_pending_adapt(self):
        window = getattr(self, "_pending_window", None)
        if window is None:
            return
        self._pending_window = None
        if window._norm_src is not self.normalization:
            window.set_normalization(self.normalization)
        eng = self._engine
        N.check(eng.lib.l2a_adapt_from_window(eng._ctx, eng._model, window._h, float(self.inner_learning_rate), 0, 1,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def switch_to_pre_adapt(self):
        self._adapted = False                       # theta itself was never modified (:347-351)
        self._pending_window = None

    def get_adapted_params(self, k):
        assert self._adapted and k < self._num_adapted_models
        self._flush_pending_adapt()
        return self._engine.get_params(1 + k)

    # ------------------------------------------------------------------ K4 with per-task weights
    def _predict_delta(self, obs, act):
        eng = self._engine
        self._flush_pending_adapt()
        if self._adapted:
            k = self._num_adapted_models
            assert obs.shape[0] % k == 0, "rows must split evenly over the adapted tasks (meta_mlp_dynamics.py:308-312)"
            d = eng.predict_delta(eng._f32(obs), eng._f32(act), N.SETS_PER_ENV, 1, k)
        else:
            d = eng.predict_delta(eng._f32(obs), eng._f32(act), N.SETS_SHARED, 0, 1)
        return d.cpu().numpy()

    def fit(self, obs, act, obs_next, epochs=1000, compute_normalization=True, valid_split_ratio=None,
            rolling_average_persitency=None, verbose=False, log_tabular=False):
        """MAML outer loop (meta_mlp_dynamics.py:167-274) on the engine's resident prior theta (weight set 0): per step a meta
        batch of `meta_batch_size` windows of 2*batch_size consecutive transitions (_get_batch :353-383, same np.random draws),
        Adam on the mean post-update loss differentiated through the one-step inner update (second order, :96-140); paths are
        split into train / validation and aggregated over fit() calls (:191-204); validation = plain MSE of theta (:242)."""
        from learning_to_adapt_b200.dynamics.fit import fit_maml, train_test_split
        assert obs.ndim == 3 and obs.shape[2] == self.obs_space_dims
        assert obs_next.ndim == 3 and obs_next.shape[2] == self.obs_space_dims
        assert act.ndim == 3 and act.shape[2] == self.action_space_dims
        valid_split_ratio = self.valid_split_ratio if valid_split_ratio is None else valid_split_ratio
        rolling_average_persitency = self.rolling_average_persitency if rolling_average_persitency is None else rolling_average_persitency
        assert 1 > valid_split_ratio >= 0
        if compute_normalization or self.normalization is None:
            self.compute_normalization(obs, act, obs_next)
        obs_n = normalize(obs, *self.normalization["obs"])
        act_n = normalize(act, *self.normalization["act"])
        delta_n = normalize(obs_next - obs, *self.normalization["delta"])
        o_tr, a_tr, d_tr, o_te, a_te, d_te = train_test_split(obs_n, act_n, delta_n, test_split_ratio=valid_split_ratio)
        if len(o_te) == 0:
            raise ValueError("valid_split_ratio=%r leaves no validation path out of %d" % (valid_split_ratio, obs.shape[0]))
        self._aggregate(dict(obs=o_tr, act=a_tr, delta=d_tr), dict(obs=o_te, act=a_te, delta=d_te))
        train, test = self._device_dataset(self._dataset_train), self._device_dataset(self._dataset_test)
        if train["x"].shape[1] <= 2 * self.batch_size:
            raise ValueError("paths of %d steps are too short for (pre, post) windows of 2 x batch_size = %d steps (:356)"
                             % (train["x"].shape[1], 2 * self.batch_size))
        params, adam = self._train_views(0)
        stats = fit_maml(params, adam, train, test, epochs=epochs, batch_size=self.batch_size, meta_batch_size=self.meta_batch_size,
                         learning_rate=self.learning_rate, inner_learning_rate=self.inner_learning_rate,
                         rolling_average_persitency=rolling_average_persitency, verbose=verbose)
        self._engine.refresh_sets(0, 1)
        self._adapted = False
        self._pending_window = None
        self._log_fit(stats, log_tabular, {"AvgModelEpochTime": float(np.mean(stats["epoch_times"])), "Post-Loss": stats["post_loss"],
                                           "Pre-Loss": stats["pre_loss"], "Epochs": stats["epochs"]})

    def __getstate__(self):
        state = dict()
        state["init_args"] = Serializable.__getstate__(self)
        state["normalization"] = self.normalization
        state["networks"] = [{"network_params": self.get_params(0)}]
        return state
