"""MLPDynamicsModel on the B200 engine: same constructor and methods as
learning_to_adapt/dynamics/mlp_dynamics.py:11-262, with the TF1 graph replaced by libl2a_b200 kernels.

Host choreography kept from the reference: float64 observations in and out, normalisation statistics in
float64, the network itself in float32 (mlp_dynamics.py:204-222).  ``predict`` is kernel K4; the planner does
not call it H times any more -- ``MPCController`` detects ``_engine`` and makes one fused K1 call.
"""
from collections import OrderedDict

import numpy as np
import torch

from learning_to_adapt_b200 import _native as N
from learning_to_adapt_b200.engine import PlanningEngine, param_keys
from learning_to_adapt_b200.utils.serializable import Serializable

_SUPPORTED_HIDDEN = ("relu",)


def _check_activations(hidden_nonlinearity, output_nonlinearity):
    name = getattr(hidden_nonlinearity, "__name__", hidden_nonlinearity)
    if name not in _SUPPORTED_HIDDEN:
        raise NotImplementedError("hidden_nonlinearity=%r: the fused kernels implement ReLU (what every run script "
                                  "uses, run_grbal.py:97); there is no CPU fallback" % (hidden_nonlinearity,))
    if output_nonlinearity is not None:
        raise NotImplementedError("output_nonlinearity=%r: the fused kernels implement a linear output layer" %
                                  (output_nonlinearity,))


def xavier_uniform_params(rng, in_dim, hidden_sizes, out_dim):
    """tf.contrib.layers.xavier_initializer + zeros bias (dynamics/core/utils.py:81-82), reference key order."""
    sizes = [in_dim] + list(hidden_sizes) + [out_dim]
    keys = param_keys(len(hidden_sizes))
    params = OrderedDict()
    for l in range(len(sizes) - 1):
        lim = np.sqrt(6.0 / (sizes[l] + sizes[l + 1]))
        params[keys[2 * l]] = rng.uniform(-lim, lim, size=(sizes[l], sizes[l + 1])).astype(np.float32)
        params[keys[2 * l + 1]] = np.zeros(sizes[l + 1], np.float32)
    return params


def normalize(data_array, mean, std):
    return (data_array - mean) / (std + 1e-10)        # mlp_dynamics.py:265-266


def denormalize(data_array, mean, std):
    return data_array * (std + 1e-10) + mean          # mlp_dynamics.py:269-270


class MLPDynamicsModel(Serializable):
    """Class for MLP continuous dynamics model (B200 engine).

    Extension over the reference signature: ``ensemble_size`` (default 1) keeps E weight sets and predicts with
    the mean of the E denormalised deltas (BASELINE.json "ensemble"); ``device`` picks the GPU.
    """

    def __init__(self, name, env, hidden_sizes=(512, 512), hidden_nonlinearity="relu", output_nonlinearity=None,
                 batch_size=500, learning_rate=0.001, normalize_input=True, optimizer=None, valid_split_ratio=0.2,
                 rolling_average_persitency=0.99, ensemble_size=1, device=0, seed=None):
        if type(self) is MLPDynamicsModel:
            Serializable.quick_init(self, locals())
        _check_activations(hidden_nonlinearity, output_nonlinearity)
        if not normalize_input:
            raise NotImplementedError("normalize_input=False: the fused kernels always apply the affine maps; "
                                      "use identity statistics instead")
        self.name = name
        self.normalization = None
        self.normalize_input = normalize_input
        self.valid_split_ratio = valid_split_ratio
        self.rolling_average_persitency = rolling_average_persitency
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.hidden_sizes = tuple(hidden_sizes)
        self.obs_space_dims = env.observation_space.shape[0]
        self.action_space_dims = env.action_space.shape[0]
        self.ensemble_size = int(ensemble_size)
        self._engine = PlanningEngine(self.obs_space_dims, self.action_space_dims, self.hidden_sizes,
                                      n_sets=self._total_sets(), device=device)
        rng = np.random.RandomState(seed)
        for e in range(self.ensemble_size):
            self._engine.set_params(e, xavier_uniform_params(rng, self.obs_space_dims + self.action_space_dims,
                                                             self.hidden_sizes, self.obs_space_dims))

    def _total_sets(self):
        return self.ensemble_size

    # ------------------------------------------------------------------ planner hook
    def planning_sets(self, n_envs):
        """(set_mode, first_set, n_sets) the fused rollout must use for ``n_envs`` observations."""
        if self.ensemble_size > 1:
            return N.SETS_ENSEMBLE_MEAN, 0, self.ensemble_size
        return N.SETS_SHARED, 0, 1

    # ------------------------------------------------------------------ parameters / statistics
    def get_params(self, member=0):
        return self._engine.get_params(member)

    def set_params(self, params, member=0):
        self._engine.set_params(member, params)

    def set_normalization(self, normalization):
        self.normalization = normalization
        self._engine.set_normalization(normalization)

    def compute_normalization(self, obs, act, obs_next):
        """Population statistics, float64 (mlp_dynamics.py:253-262)."""
        assert obs.shape[0] == obs_next.shape[0] == act.shape[0]
        delta = obs_next - obs
        assert delta.ndim == 2 and delta.shape[0] == obs_next.shape[0]
        normalization = OrderedDict()
        normalization["obs"] = (np.mean(obs, axis=0), np.std(obs, axis=0))
        normalization["delta"] = (np.mean(delta, axis=0), np.std(delta, axis=0))
        normalization["act"] = (np.mean(act, axis=0), np.std(act, axis=0))
        self.set_normalization(normalization)

    def _normalize_data(self, obs, act, obs_next=None):
        obs_normalized = normalize(obs, *self.normalization["obs"])
        actions_normalized = normalize(act, *self.normalization["act"])
        if obs_next is not None:
            deltas_normalized = normalize(obs_next - obs, *self.normalization["delta"])
            return obs_normalized, actions_normalized, deltas_normalized
        return obs_normalized, actions_normalized

    # ------------------------------------------------------------------ K4
    def _predict_delta(self, obs, act):
        set_mode, first_set, n_sets = self.planning_sets(1)
        eng = self._engine
        d = eng.predict_delta(eng._f32(obs), eng._f32(act), set_mode, first_set, n_sets)
        return d.cpu().numpy()

    def predict(self, obs, act):
        """obs [n, D], act [n, A] float64 -> next_obs [n, D] float64 (mlp_dynamics.py:204-222).  The device returns the
        float32 denormalised delta; the float64 ``obs + delta`` happens on the host as in the reference."""
        assert obs.shape[0] == act.shape[0]
        assert obs.ndim == 2 and obs.shape[1] == self.obs_space_dims
        assert act.ndim == 2 and act.shape[1] == self.action_space_dims
        assert self.normalization is not None, "call fit() / compute_normalization() / set_normalization() first"
        delta = self._predict_delta(obs, act)
        assert delta.ndim == 2
        return np.asarray(obs, np.float64) + delta

    # ------------------------------------------------------------------ fit: on the engine's resident parameters (f2)
    def _aggregate(self, train, test):
        """self._dataset_train / _dataset_test keep every fit() call's normalised split and training runs on everything
        collected so far (mlp_dynamics.py:119-129; the trainer passes only the newest iteration's samples, mb_trainer.py:89-94)."""
        if getattr(self, "_dataset_test", None) is None:
            self._dataset_train, self._dataset_test = train, test
        else:
            for k in ("obs", "act", "delta"):
                self._dataset_test[k] = np.concatenate([self._dataset_test[k], test[k]])
                self._dataset_train[k] = np.concatenate([self._dataset_train[k], train[k]])

    def _device_dataset(self, d):
        dev = self._engine.device
        x = np.concatenate([d["obs"], d["act"]], axis=-1).astype(np.float32)           # the float32 placeholders (:63-68)
        return dict(x=torch.as_tensor(x, device=dev), y=torch.as_tensor(d["delta"].astype(np.float32), device=dev))

    def _train_views(self, member):
        """(leaf tensors onto the resident parameters of `member`, its Adam slots) -- created once, kept like TF variables."""
        from learning_to_adapt_b200.dynamics.fit import AdamState
        if not hasattr(self, "_fit_state"):
            self._fit_state = {}
        if member not in self._fit_state:
            views = [v.requires_grad_(True) for v in self._engine.param_views(member)]
            self._fit_state[member] = (views, AdamState(views))
        return self._fit_state[member]

    def fit(self, obs, act, obs_next, epochs=1000, compute_normalization=True, valid_split_ratio=None,
            rolling_average_persitency=None, verbose=False, log_tabular=False):
        """Adam on mean((delta_n - f(x_n))^2) over everything collected so far, validation-based early stop
        (mlp_dynamics.py:91-202), on the engine's resident parameters (no host round trip of the weights).
        log_tabular: the statistics the reference logs (Epochs, AvgModelEpochTime) are left in ``self.last_fit_stats`` and, when
        a ``logger`` with logkv() has been attached as ``self.logger``, logged through it."""
        from learning_to_adapt_b200.dynamics.fit import fit_mlp, train_test_split
        assert obs.ndim == 2 and obs.shape[1] == self.obs_space_dims
        assert obs_next.ndim == 2 and obs_next.shape[1] == self.obs_space_dims
        assert act.ndim == 2 and act.shape[1] == self.action_space_dims
        valid_split_ratio = self.valid_split_ratio if valid_split_ratio is None else valid_split_ratio
        rolling_average_persitency = self.rolling_average_persitency if rolling_average_persitency is None else rolling_average_persitency
        assert 1 > valid_split_ratio >= 0
        if compute_normalization or self.normalization is None:
            self.compute_normalization(obs, act, obs_next)
        obs_n, act_n, delta_n = self._normalize_data(obs, act, obs_next)
        o_tr, a_tr, d_tr, o_te, a_te, d_te = train_test_split(obs_n, act_n, delta_n, test_split_ratio=valid_split_ratio)
        self._aggregate(dict(obs=o_tr, act=a_tr, delta=d_tr), dict(obs=o_te, act=a_te, delta=d_te))
        train, test = self._device_dataset(self._dataset_train), self._device_dataset(self._dataset_test)
        stats = None
        for e in range(self.ensemble_size):
            params, adam = self._train_views(e)
            stats = fit_mlp(params, adam, train, test, epochs=epochs, batch_size=self.batch_size, learning_rate=self.learning_rate,
                            rolling_average_persitency=rolling_average_persitency, verbose=verbose)
        self._engine.refresh_sets(0, self.ensemble_size)
        self._log_fit(stats, log_tabular, dict(AvgModelEpochTime=float(np.mean(stats["epoch_times"])), Epochs=stats["epochs"]))

    def _log_fit(self, stats, log_tabular, kv):
        self.last_fit_stats = dict(stats, **kv)
        logger = getattr(self, "logger", None)
        if log_tabular and logger is not None:
            for k, v in kv.items():
                logger.logkv(k, v)

    # ------------------------------------------------------------------ pickling: ctor args + statistics + weights
    def __getstate__(self):
        state = dict()
        state["init_args"] = Serializable.__getstate__(self)
        state["normalization"] = self.normalization
        state["networks"] = [{"network_params": self.get_params(e)} for e in range(self.ensemble_size)]
        return state

    def __setstate__(self, state):
        Serializable.__setstate__(self, state["init_args"])
        if state["normalization"] is not None:
            self.set_normalization(state["normalization"])
        for e, net in enumerate(state["networks"]):
            self.set_params(net["network_params"], member=e)
