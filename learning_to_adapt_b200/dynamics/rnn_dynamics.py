"""RNNDynamicsModel (ReBAL) on the B200 engine: same constructor and hot-path methods as
learning_to_adapt/dynamics/rnn_dynamics.py:11-300 -- ``predict(obs, act, hidden) -> (next_obs, hidden)`` and
``get_initial_hidden(batch_size)`` -- for the configuration the run script ships (run_rebal.py:98-99): one LSTM layer.

The hidden state is the reference's ``LSTMStateTuple(c, h)`` of float32 arrays ``[batch, hidden]`` (a namedtuple, so code that
iterates it as ``(c, h)`` -- rnn_mpc_controller.py:136-163 -- keeps working).  Recurrent "adaptation" is just this state.
Training (``fit``: BPTT with Adam, rnn_dynamics.py:95-231) is host-side torch glue, off the planning hot path.
"""
import collections
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from learning_to_adapt_b200 import _native as N
from learning_to_adapt_b200.dynamics.mlp_dynamics import normalize
from learning_to_adapt_b200.engine import EPS, _ptr, _stream
from learning_to_adapt_b200.utils.serializable import Serializable

LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))
RNN_PARAM_KEYS = ["rnn/lstm_cell/kernel", "rnn/lstm_cell/bias", "output/kernel", "output/bias"]


class RNNDynamicsModel(Serializable):
    def __init__(self, name, env, hidden_sizes=(512,), cell_type="lstm", hidden_nonlinearity="tanh", output_nonlinearity=None,
                 batch_size=500, learning_rate=0.001, normalize_input=True, optimizer=None, valid_split_ratio=0.2,
                 rolling_average_persitency=0.99, backprop_steps=50, device=0, seed=None):
        Serializable.quick_init(self, locals())
        if cell_type != "lstm" or len(hidden_sizes) != 1:
            raise NotImplementedError("the fused ReBAL kernel implements one LSTM layer (run_rebal.py:98-99); got cell_type=%r, "
                                      "hidden_sizes=%r -- there is no CPU fallback" % (cell_type, hidden_sizes))
        if getattr(hidden_nonlinearity, "__name__", hidden_nonlinearity) != "tanh" or output_nonlinearity is not None:
            raise NotImplementedError("the fused LSTM cell uses tanh / linear output (the reference defaults)")
        if not torch.cuda.is_available():
            raise RuntimeError("learning_to_adapt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.recurrent = True
        self.name = name
        self.normalization = None
        self.normalize_input = normalize_input
        self.valid_split_ratio, self.rolling_average_persitency = valid_split_ratio, rolling_average_persitency
        self.backprop_steps, self.batch_size, self.learning_rate = backprop_steps, batch_size, learning_rate
        self.hidden_size = int(hidden_sizes[0])
        self.obs_space_dims = env.observation_space.shape[0]
        self.action_space_dims = env.action_space.shape[0]
        self.lib = N.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._ctx = C.c_void_p()
        N.check(self.lib.l2a_ctx_create(device, C.byref(self._ctx)))
        self._model = C.c_void_p()
        N.check(self.lib.l2a_rnn_model_create(self._ctx, self.obs_space_dims, self.action_space_dims, self.hidden_size, C.byref(self._model)))
        self._discount_cache = {}
        rng = np.random.RandomState(seed)
        D, A, Hs = self.obs_space_dims, self.action_space_dims, self.hidden_size
        lim_k = np.sqrt(6.0 / (D + A + Hs + 4 * Hs))
        lim_o = np.sqrt(6.0 / (Hs + D))
        self.set_params(OrderedDict([
            (RNN_PARAM_KEYS[0], rng.uniform(-lim_k, lim_k, size=(D + A + Hs, 4 * Hs)).astype(np.float32)),
            (RNN_PARAM_KEYS[1], np.zeros(4 * Hs, np.float32)),
            (RNN_PARAM_KEYS[2], rng.uniform(-lim_o, lim_o, size=(Hs, D)).astype(np.float32)),
            (RNN_PARAM_KEYS[3], np.zeros(D, np.float32))]))

    def _f32(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=self.device)

    def close(self):
        if getattr(self, "_model", None):
            self.lib.l2a_rnn_model_destroy(self._ctx, self._model)
            self._model = None
        if getattr(self, "_ctx", None):
            self.lib.l2a_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters / statistics
    def set_params(self, params):
        vals = [self._f32(params[k]) for k in RNN_PARAM_KEYS]
        D, A, Hs = self.obs_space_dims, self.action_space_dims, self.hidden_size
        assert tuple(vals[0].shape) == (D + A + Hs, 4 * Hs) and tuple(vals[2].shape) == (Hs, D)
        N.check(self.lib.l2a_rnn_model_set_params(self._ctx, self._model, *[_ptr(v) for v in vals], _stream()))
        torch.cuda.current_stream().synchronize()
        self._params = OrderedDict((k, v.cpu().numpy()) for k, v in zip(RNN_PARAM_KEYS, vals))

    def get_params(self):
        return OrderedDict((k, v.copy()) for k, v in self._params.items())

    def set_normalization(self, normalization):
        n = normalization
        arrs = [np.asarray(n["obs"][0], np.float64), np.asarray(n["obs"][1], np.float64) + EPS,
                np.asarray(n["act"][0], np.float64), np.asarray(n["act"][1], np.float64) + EPS,
                np.asarray(n["delta"][0], np.float64), np.asarray(n["delta"][1], np.float64) + EPS]
        ts = [self._f32(a) for a in arrs]
        N.check(self.lib.l2a_rnn_model_set_normalization(self._ctx, self._model, *[_ptr(t) for t in ts], _stream()))
        torch.cuda.current_stream().synchronize()
        self.normalization = normalization

    def compute_normalization(self, obs, act, obs_next):
        """[paths, T, dim] inputs, statistics over axes (0, 1) (rnn_dynamics.py:290-300)."""
        assert obs.ndim == 3 and obs.shape[:2] == obs_next.shape[:2] == act.shape[:2]
        delta = obs_next - obs
        normalization = OrderedDict()
        normalization["obs"] = (np.mean(obs, axis=(0, 1)), np.std(obs, axis=(0, 1)))
        normalization["delta"] = (np.mean(delta, axis=(0, 1)), np.std(delta, axis=(0, 1)))
        normalization["act"] = (np.mean(act, axis=(0, 1)), np.std(act, axis=(0, 1)))
        self.set_normalization(normalization)

    # ------------------------------------------------------------------ hot path
    def get_initial_hidden(self, batch_size):
        z = np.zeros((batch_size, self.hidden_size), np.float32)         # cell.zero_state (rnn_dynamics.py:266-288)
        return LSTMStateTuple(z.copy(), z.copy())

    def predict(self, obs, act, hidden_state):
        """(obs [n,D], act [n,A]) float64, hidden (c, h) [n,Hs] -> (next_obs float64, LSTMStateTuple)  (rnn_dynamics.py:233-252)."""
        assert obs.shape[0] == act.shape[0]
        assert obs.ndim == 2 and obs.shape[1] == self.obs_space_dims
        assert act.ndim == 2 and act.shape[1] == self.action_space_dims
        n = obs.shape[0]
        c_in, h_in = self._f32(hidden_state[0]), self._f32(hidden_state[1])
        assert tuple(c_in.shape) == (n, self.hidden_size)
        delta = torch.empty(n, self.obs_space_dims, device=self.device, dtype=torch.float32)
        c_out, h_out = torch.empty_like(c_in), torch.empty_like(h_in)
        obs_dev, act_dev = self._f32(obs), self._f32(act)          # keep the staging tensors alive across the call
        N.check(self.lib.l2a_rnn_predict(self._ctx, self._model, _ptr(obs_dev), _ptr(act_dev), _ptr(c_in), _ptr(h_in),
                                         int(n), _ptr(delta), _ptr(c_out), _ptr(h_out), _stream()))
        return np.asarray(obs, np.float64) + delta.cpu().numpy(), LSTMStateTuple(c_out.cpu().numpy(), h_out.cpu().numpy())

    def rollout(self, obs_dev, hidden, actions_dev, n_candidates, horizon, reward_kind, dt, discount=1.0, want_returns=False,
                kernel=N.KERNEL_AUTO, layout="thra"):
        """Fused H-step planner rollout (kernel behind RNNMPCController.get_rs_action)."""
        m, A = obs_dev.shape[0], self.action_space_dims
        key = (float(discount), int(horizon))
        if key not in self._discount_cache:
            self._discount_cache[key] = self._f32(np.array([float(discount) ** t for t in range(horizon)], np.float64))
        p = N.RolloutParams()
        p.n_candidates, p.n_envs, p.horizon = int(n_candidates), int(m), int(horizon)
        p.reward_kind, p.dt, p.kernel = int(reward_kind), float(dt), int(kernel)
        if layout == "thra":                     # [H, m*N, A] (random shooting)
            p.act_stride_t, p.act_stride_row = n_candidates * m * A, A
        else:                                    # 'nmha': [N, m, H*A] viewed as (N*m, H, A) (CEM, rnn_mpc_controller.py:87-88)
            p.act_stride_t, p.act_stride_row = A, horizon * A
        best_ret = torch.empty(m, device=self.device, dtype=torch.float32)
        best_idx = torch.empty(m, device=self.device, dtype=torch.int32)
        best_act = torch.empty(m, A, device=self.device, dtype=torch.float32)
        returns = torch.empty(m, n_candidates, device=self.device, dtype=torch.float32) if want_returns else None
        c_in, h_in = self._f32(hidden[0]), self._f32(hidden[1])
        N.check(self.lib.l2a_rnn_rollout(self._ctx, self._model, C.byref(p), _ptr(obs_dev), _ptr(c_in), _ptr(h_in), _ptr(actions_dev),
                                         _ptr(self._discount_cache[key]), _ptr(returns), _ptr(best_ret), _ptr(best_idx),
                                         _ptr(best_act), _stream()))
        return dict(best_ret=best_ret, best_idx=best_idx, best_act=best_act, returns=returns)

    def fit(self, obs, act, obs_next, epochs=1000, compute_normalization=True, valid_split_ratio=None,
            rolling_average_persitency=None, verbose=False, log_tabular=False):
        """Truncated-BPTT Adam regression of the normalised deltas (rnn_dynamics.py:95-231), torch glue off the hot path."""
        from learning_to_adapt_b200.dynamics.fit import fit_lstm
        assert obs.ndim == 3 and act.ndim == 3 and obs_next.ndim == 3
        if compute_normalization or self.normalization is None:
            self.compute_normalization(obs, act, obs_next)
        obs_n = normalize(obs, *self.normalization["obs"])
        act_n = normalize(act, *self.normalization["act"])
        delta_n = normalize(obs_next - obs, *self.normalization["delta"])
        self.set_params(fit_lstm(self.get_params(), obs_n, act_n, delta_n, epochs=epochs, batch_size=self.batch_size,
                                 learning_rate=self.learning_rate, backprop_steps=self.backprop_steps,
                                 valid_split_ratio=self.valid_split_ratio if valid_split_ratio is None else valid_split_ratio,
                                 rolling_average_persitency=(self.rolling_average_persitency if rolling_average_persitency is None
                                                             else rolling_average_persitency),
                                 device=self.device, verbose=verbose))

    def __getstate__(self):
        return {"init_args": Serializable.__getstate__(self), "normalization": self.normalization,
                "networks": [{"network_params": self.get_params()}]}

    def __setstate__(self, state):
        Serializable.__setstate__(self, state["init_args"])
        if state["normalization"] is not None:
            self.set_normalization(state["normalization"])
        self.set_params(state["networks"][0]["network_params"])
