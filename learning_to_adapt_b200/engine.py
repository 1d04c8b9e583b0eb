"""PlanningEngine: torch tensors in, libl2a_b200.so kernels underneath.

PyTorch is only the container (device memory, streams); every computation on the planning path is one of the
hand-written sm_100a kernels behind the C ABI in include/l2a_b200.h.  No CPU fallback.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _native as N

EPS = 1e-10  # dynamics/mlp_dynamics.py:265-270


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def param_keys(n_hidden):
    """Key order of the reference's parameter dict (dynamics/core/utils.py:241-296)."""
    keys = []
    for i in range(n_hidden):
        keys += ["hidden_%d/kernel" % i, "hidden_%d/bias" % i]
    return keys + ["output/kernel", "output/bias"]


class PlanningEngine(object):
    """One (process, device) planning context + one dynamics model with `n_sets` resident weight sets."""

    def __init__(self, obs_dim, act_dim, hidden_sizes, n_sets=1, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("learning_to_adapt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = N.load()
        self.device = torch.device("cuda", device)
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.hidden_sizes = tuple(int(h) for h in hidden_sizes)
        self.n_sets = int(n_sets)
        torch.cuda.set_device(self.device)
        self._ctx = C.c_void_p()
        N.check(self.lib.l2a_ctx_create(device, C.byref(self._ctx)))
        desc = N.MlpDesc()
        desc.obs_dim, desc.act_dim, desc.n_hidden, desc.n_sets = self.obs_dim, self.act_dim, len(self.hidden_sizes), self.n_sets
        for i, h in enumerate(self.hidden_sizes):
            desc.hidden[i] = h
        self._model = C.c_void_p()
        N.check(self.lib.l2a_model_create(self._ctx, C.byref(desc), C.byref(self._model)))
        self._layer_shapes = []
        sizes = [self.obs_dim + self.act_dim] + list(self.hidden_sizes) + [self.obs_dim]
        for l in range(len(sizes) - 1):
            self._layer_shapes.append((sizes[l], sizes[l + 1]))
        self._discount_cache = {}
        self._norm = None
        self._plans = {}

    # ------------------------------------------------------------------ lifetime
    def close(self):
        for plan in getattr(self, "_plans", {}).values():
            self.lib.l2a_plan_destroy(self._ctx, plan["handle"])
        self._plans = {}
        if getattr(self, "_model", None) is not None and self._model:
            self.lib.l2a_model_destroy(self._ctx, self._model)
            self._model = None
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.l2a_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.l2a_ctx_launch_count(self._ctx))

    def _f32(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=self.device)

    # ------------------------------------------------------------------ parameters
    def set_params(self, set_idx, params):
        """params: OrderedDict in the reference's key order ('hidden_i/kernel' [in,out], 'hidden_i/bias', ..., 'output/*')
        or a flat list [W0, b0, W1, b1, ...]."""
        vals = list(params.values()) if isinstance(params, dict) else list(params)
        assert len(vals) == 2 * len(self._layer_shapes), "expected %d arrays" % (2 * len(self._layer_shapes))
        ws, bs = [], []
        for l, (din, dout) in enumerate(self._layer_shapes):
            w, b = self._f32(vals[2 * l]), self._f32(vals[2 * l + 1])
            assert tuple(w.shape) == (din, dout), "layer %d kernel shape %s != %s" % (l, tuple(w.shape), (din, dout))
            assert tuple(b.shape) == (dout,)
            ws.append(w)
            bs.append(b)
        nl = len(ws)
        wp = (C.c_void_p * nl)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * nl)(*[b.data_ptr() for b in bs])
        N.check(self.lib.l2a_model_set_params(self._ctx, self._model, int(set_idx), wp, bp, _stream()))
        torch.cuda.current_stream().synchronize()   # the staging tensors die with this frame

    def get_params(self, set_idx):
        ws = [torch.empty(s, device=self.device, dtype=torch.float32) for s in self._layer_shapes]
        bs = [torch.empty(s[1], device=self.device, dtype=torch.float32) for s in self._layer_shapes]
        nl = len(ws)
        wp = (C.c_void_p * nl)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * nl)(*[b.data_ptr() for b in bs])
        N.check(self.lib.l2a_model_get_params(self._ctx, self._model, int(set_idx), wp, bp, _stream()))
        out = OrderedDict()
        for key, t in zip(param_keys(len(self.hidden_sizes)), [x for pair in zip(ws, bs) for x in pair]):
            out[key] = t.cpu().numpy()
        return out

    def set_normalization(self, normalization):
        """normalization: the reference's dict {'obs': (mean, std), 'act': (...), 'delta': (...)} in float64
        (mlp_dynamics.py:253-262).  Denominators std + 1e-10 are formed in float64, then rounded to fp32."""
        n = normalization
        arrs = [np.asarray(n["obs"][0], np.float64), np.asarray(n["obs"][1], np.float64) + EPS,
                np.asarray(n["act"][0], np.float64), np.asarray(n["act"][1], np.float64) + EPS,
                np.asarray(n["delta"][0], np.float64), np.asarray(n["delta"][1], np.float64) + EPS]
        assert arrs[0].shape == (self.obs_dim,) and arrs[2].shape == (self.act_dim,) and arrs[4].shape == (self.obs_dim,)
        ts = [self._f32(a) for a in arrs]
        N.check(self.lib.l2a_model_set_normalization(self._ctx, self._model, *[_ptr(t) for t in ts], _stream()))
        torch.cuda.current_stream().synchronize()
        self._norm = normalization

    # ------------------------------------------------------------------ K1
    def _discount_pow(self, discount, horizon):
        key = (float(discount), int(horizon))
        if key not in self._discount_cache:
            pw = np.array([float(discount) ** t for t in range(horizon)], np.float64)   # mpc_controller.py:126
            self._discount_cache[key] = self._f32(pw)
        return self._discount_cache[key]

    def rollout(self, obs0, actions, n_candidates, horizon, reward_kind, dt, discount=1.0, set_mode=N.SETS_SHARED,
                first_set=0, n_sets=1, layout="thra", want_returns=True, kernel=N.KERNEL_AUTO):
        """obs0 [m, D] device fp32; actions: device fp32 candidate tensor.
        layout 'thra': [H, m*N, A] (random shooting, mpc_controller.py:114); 'nmha': [N, m, H*A] (CEM, :85-89).
        Returns dict(best_ret [m], best_idx [m] int32, best_act [m, A], returns [m, N] or None) of device tensors."""
        m = obs0.shape[0]
        A = self.act_dim
        assert obs0.is_cuda and obs0.dtype == torch.float32 and obs0.is_contiguous() and obs0.shape[1] == self.obs_dim
        assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
        p = N.RolloutParams()
        p.n_candidates, p.n_envs, p.horizon = int(n_candidates), int(m), int(horizon)
        p.set_mode, p.first_set, p.n_sets = int(set_mode), int(first_set), int(n_sets)
        p.reward_kind, p.dt, p.kernel = int(reward_kind), float(dt), int(kernel)
        rows = n_candidates * m
        if layout == "thra":
            assert actions.numel() == horizon * rows * A
            p.act_stride_t, p.act_stride_row = rows * A, A
        elif layout == "nmha":
            assert actions.numel() == rows * horizon * A
            p.act_stride_t, p.act_stride_row = A, horizon * A
        else:
            raise ValueError(layout)
        best_ret = torch.empty(m, device=self.device, dtype=torch.float32)
        best_idx = torch.empty(m, device=self.device, dtype=torch.int32)
        best_act = torch.empty(m, A, device=self.device, dtype=torch.float32)
        returns = torch.empty(m, n_candidates, device=self.device, dtype=torch.float32) if want_returns else None
        N.check(self.lib.l2a_rollout(self._ctx, self._model, C.byref(p), _ptr(obs0), _ptr(actions),
                                     _ptr(self._discount_pow(discount, horizon)), _ptr(returns), _ptr(best_ret),
                                     _ptr(best_idx), _ptr(best_act), _stream()))
        return dict(best_ret=best_ret, best_idx=best_idx, best_act=best_act, returns=returns)

    # ------------------------------------------------------------------ host-buffer planning call (l2a_plan_*)
    def plan_rs_host(self, observations, n_candidates, horizon, reward_kind, dt, low, high, discount=1.0,
                     set_mode=N.SETS_SHARED, first_set=0, n_sets=1, kernel=N.KERNEL_AUTO, seed=0):
        """One random-shooting planning call with HOST arrays on both sides (policies/mpc_controller.py:59-65, 108-129):
        observations float64 [m, D] -> (actions float64 [m, A], best_ret float32 [m], best_idx int32 [m]).  The candidates
        are drawn on the device (Philox); after the first call the whole sequence H2D -> sample -> K1 -> D2H is one CUDA
        graph replay inside libl2a_b200."""
        obs = np.ascontiguousarray(observations, dtype=np.float64)
        m = obs.shape[0]
        assert obs.shape == (m, self.obs_dim)
        low32 = np.ascontiguousarray(low, dtype=np.float32)
        high32 = np.ascontiguousarray(high, dtype=np.float32)
        assert low32.shape == (self.act_dim,) and high32.shape == (self.act_dim,)
        key = (m, int(n_candidates), int(horizon), int(reward_kind), float(dt), float(discount), int(set_mode), int(first_set),
               int(n_sets), int(kernel), low32.tobytes(), high32.tobytes(), int(seed))
        plan = self._plans.get(key)
        if plan is None:
            p = N.RolloutParams()
            p.n_candidates, p.n_envs, p.horizon = int(n_candidates), int(m), int(horizon)
            p.set_mode, p.first_set, p.n_sets = int(set_mode), int(first_set), int(n_sets)
            p.reward_kind, p.dt, p.kernel = int(reward_kind), float(dt), int(kernel)
            handle = C.c_void_p()
            N.check(self.lib.l2a_plan_create(self._ctx, self._model, C.byref(p), float(discount),
                                             low32.ctypes.data_as(C.c_void_p), high32.ctypes.data_as(C.c_void_p),
                                             C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.byref(handle)))
            plan = dict(handle=handle, act=np.empty((m, self.act_dim), np.float64), ret=np.empty(m, np.float32),
                        idx=np.empty(m, np.int32), shape=(int(horizon), int(n_candidates) * m, self.act_dim))
            self._plans[key] = plan
        N.check(self.lib.l2a_plan_run(self._ctx, plan["handle"], obs.ctypes.data_as(C.c_void_p),
                                      plan["act"].ctypes.data_as(C.c_void_p), plan["ret"].ctypes.data_as(C.c_void_p),
                                      plan["idx"].ctypes.data_as(C.c_void_p), _stream()))
        self._last_plan = plan
        return plan["act"].copy(), plan["ret"].copy(), plan["idx"].copy()

    def sample_uniform(self, low, high, rows, seed=0, call_index=0, out=None):
        """[rows, A] float32 device tensor of U[low, high) draws (Philox4x32-10, this library's kernel)."""
        key = ("bounds", np.asarray(low, np.float32).tobytes(), np.asarray(high, np.float32).tobytes())
        if key not in self._discount_cache:
            self._discount_cache[key] = (self._f32(low), self._f32(high))
        lo, hi = self._discount_cache[key]
        if out is None:
            out = torch.empty((int(rows), self.act_dim), device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_sample_uniform(self._ctx, _ptr(lo), _ptr(hi), _ptr(out), int(rows), int(self.act_dim),
                                            C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint64(int(call_index)), _stream()))
        return out

    def last_plan_candidates(self):
        """[H, m*N, A] float32 candidates of the most recent plan_rs_host call (tests / diagnostics)."""
        plan = self._last_plan
        out = np.empty(plan["shape"], np.float32)
        N.check(self.lib.l2a_plan_copy_candidates(self._ctx, plan["handle"], out.ctypes.data_as(C.c_void_p)))
        return out

    def last_plan_uses_graph(self):
        return bool(self.lib.l2a_plan_uses_graph(self._last_plan["handle"]))

    # ------------------------------------------------------------------ K4
    def predict_delta(self, obs, act, set_mode=N.SETS_SHARED, first_set=0, n_sets=1, kernel=N.KERNEL_AUTO):
        """obs [n, D], act [n, A] device fp32 (raw).  Returns the denormalised delta [n, D] (device fp32)."""
        n = obs.shape[0]
        delta = torch.empty(n, self.obs_dim, device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_predict(self._ctx, self._model, int(set_mode), int(first_set), int(n_sets), _ptr(obs),
                                     _ptr(act), int(n), _ptr(delta), C.c_void_p(0), int(kernel), _stream()))
        return delta

    # ------------------------------------------------------------------ K2
    def adapt(self, x_norm, target_norm, inner_lr, src_set=0, dst_first_set=1):
        """x_norm [K, M, D+A], target_norm [K, M, D] device fp32 (already normalised, meta_mlp_dynamics.py:334-339)."""
        K, M = x_norm.shape[0], x_norm.shape[1]
        assert x_norm.is_contiguous() and target_norm.is_contiguous()
        N.check(self.lib.l2a_adapt(self._ctx, self._model, _ptr(x_norm), _ptr(target_norm), int(K), int(M),
                                   float(inner_lr), int(src_set), int(dst_first_set), _stream()))

    # ------------------------------------------------------------------ K1c
    def cem_sample(self, z, mean, std, clip_low, clip_high):
        n, m, ha = z.shape
        samples = torch.empty(n, m, ha, device=self.device, dtype=torch.float32)
        clipped = torch.empty(n, m, ha, device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_cem_sample(self._ctx, _ptr(z), _ptr(mean), _ptr(std), _ptr(clip_low), _ptr(clip_high),
                                        int(n), int(m), int(ha), _ptr(samples), _ptr(clipped), _stream()))
        return samples, clipped

    def cem_refit(self, returns, clipped, num_elites, alpha, mean, std, compat=True):
        n, m, ha = clipped.shape
        rank = torch.empty(m, n, device=self.device, dtype=torch.int32)
        N.check(self.lib.l2a_cem_refit(self._ctx, _ptr(returns), _ptr(clipped), int(n), int(m), int(ha), int(num_elites),
                                       float(alpha), 1 if compat else 0, _ptr(rank), _ptr(mean), _ptr(std), _stream()))

    # ------------------------------------------------------------------ K3
    def shard_pack(self, best_ret, best_idx, best_act, idx_offset):
        m = best_ret.shape[0]
        packed = torch.empty(m, 3 + self.act_dim, device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_shard_pack(self._ctx, _ptr(best_ret), _ptr(best_idx), _ptr(best_act), int(idx_offset), int(m),
                                        int(self.act_dim), _ptr(packed), _stream()))
        return packed

    def shard_select(self, gathered):
        G, m, _ = gathered.shape
        best_ret = torch.empty(m, device=self.device, dtype=torch.float32)
        best_idx = torch.empty(m, device=self.device, dtype=torch.int64)
        best_act = torch.empty(m, self.act_dim, device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_shard_select(self._ctx, _ptr(gathered), int(G), int(m), int(self.act_dim), _ptr(best_ret),
                                          _ptr(best_idx), _ptr(best_act), _stream()))
        return best_ret, best_idx, best_act

    # ------------------------------------------------------------------ diagnostics
    def debug_umma_tile(self, A, B, variant=0):
        n, k = B.shape
        Cm = torch.empty(128, n, device=self.device, dtype=torch.float32)
        N.check(self.lib.l2a_debug_umma_tile(self._ctx, _ptr(A), _ptr(B), _ptr(Cm), int(n), int(k), int(variant), _stream()))
        return Cm
