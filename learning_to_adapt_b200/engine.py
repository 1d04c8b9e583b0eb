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


_NUMPY_MT = None


def numpy_mt19937_state():
    """Direct access to numpy's GLOBAL legacy generator state (the stream np.random.uniform / normal draw from): returns
    (bit_generator, address of its `mt19937_state` struct {uint32 key[624]; int pos;}) or None when the layout cannot be confirmed.

    np.random.get_state() + set_state() cost ~150 us per planning call (tuple building, validation, a 624-word copy each way) --
    more than the device spends regenerating the stream.  The bit generator's ctypes interface exposes the struct's address
    (numpy/random/src/mt19937/mt19937.h); the host-buffer planning call reads the key / position from there and writes the advanced
    state back in place, under the generator's own lock.  The layout is checked once against get_state()."""
    global _NUMPY_MT
    if _NUMPY_MT and np.random.mtrand._rand._bit_generator is not _NUMPY_MT[0]:
        _NUMPY_MT = None                  # np.random.set_bit_generator() replaced the global generator: look again
    if _NUMPY_MT is None:
        _NUMPY_MT = False
        try:
            bg = np.random.mtrand._rand._bit_generator
            addr = bg.ctypes.state_address
            addr = int(addr.value if hasattr(addr, "value") else addr)
            with bg.lock:
                key = np.ctypeslib.as_array((C.c_uint32 * 624).from_address(addr))
                pos = C.c_int32.from_address(addr + 624 * 4)
                st = np.random.get_state()
                if st[0] == "MT19937" and np.array_equal(key, st[1]) and int(pos.value) == int(st[2]):
                    _NUMPY_MT = (bg, addr)
        except Exception:
            _NUMPY_MT = False
    return _NUMPY_MT or None


class PlanningEngine(object):
    """One (process, device) planning context + one dynamics model with `n_sets` resident weight sets."""

    def __init__(self, obs_dim, act_dim, hidden_sizes, n_sets=1, device=0, debug=False):
        if not torch.cuda.is_available():
            raise RuntimeError("learning_to_adapt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = N.load(debug=debug)                # debug=True: the build that also carries the l2a_debug_* diagnostics
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
            for peer in plan.get("peers", []):
                self.lib.l2a_ipc_close_handle(self._ctx, peer)
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

    def param_views(self, set_idx):
        """Torch views [W_0, b_0, W_1, b_1, ...] ONTO the resident fp32 parameters of weight set `set_idx` (no copy): training
        updates them in place on the device; call refresh_sets() afterwards so the tensor-core tiles follow."""
        nl = len(self._layer_shapes)
        ptr, count = C.c_void_p(), C.c_int64()
        w_off, b_off = (C.c_int32 * nl)(), (C.c_int32 * nl)()
        N.check(self.lib.l2a_model_param_block(self._ctx, self._model, int(set_idx), C.byref(ptr), C.byref(count), w_off, b_off))

        class _Block(object):
            __cuda_array_interface__ = dict(shape=(int(count.value),), typestr="<f4", data=(int(ptr.value), False), version=2)

        flat = torch.as_tensor(_Block(), device=self.device)
        views = []
        for l, (din, dout) in enumerate(self._layer_shapes):
            views.append(flat[w_off[l]:w_off[l] + din * dout].view(din, dout))
            views.append(flat[b_off[l]:b_off[l] + dout])
        return views

    def refresh_sets(self, first_set=0, n_sets=1):
        N.check(self.lib.l2a_model_refresh(self._ctx, self._model, int(first_set), int(n_sets), _stream()))

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
                     set_mode=N.SETS_SHARED, first_set=0, n_sets=1, kernel=N.KERNEL_AUTO, seed=0, sampler="philox", shard=None,
                     window=None, flags=0, inner_lr=0.0):
        """One random-shooting planning call with HOST arrays on both sides (policies/mpc_controller.py:59-65, 108-129):
        observations float64 [m, D] -> (actions float64 [m, A], best_ret float32 [m], best_idx int64 [m]).  ONE C call; after the
        first call the whole sequence H2D -> sample -> K1 [-> peer exchange] -> D2H is one CUDA graph replay inside libl2a_b200.

        sampler "philox": candidates from the device Philox stream (throughput mode).
        sampler "mt19937": the reference's own draw, np.random.uniform(low, high, (H*N*m, A)) (:67-69, 114), regenerated on the
            device from np.random.get_state(); np.random.set_state() then leaves the global stream exactly where the reference
            would have left it, and the returned actions are the float64 candidates the reference returns.
        shard: None, or dict(rank, world, all_gather) -- this process rolls its slice of the `n_candidates` of every env and the
            ranks' winners are exchanged over peer memory inside the call (`all_gather(bytes) -> list of bytes` is only used once,
            to exchange the IPC handles of the exchange buffers).
        window / flags / inner_lr: the GrBAL env step inside the same call (samplers/sampler.py:81-91): with an AdaptWindow and
            N.PLAN_ADAPT the call first adapts weight set 0 into sets 1.. on the window's last M transitions (K2), with N.PLAN_PUSH
            it appends (observation, chosen action) to the window afterwards."""
        obs = np.ascontiguousarray(observations, dtype=np.float64)
        m = obs.shape[0]
        assert obs.shape == (m, self.obs_dim)
        low64 = np.ascontiguousarray(low, dtype=np.float64)
        high64 = np.ascontiguousarray(high, dtype=np.float64)
        assert low64.shape == (self.act_dim,) and high64.shape == (self.act_dim,)
        rank, world = (int(shard["rank"]), int(shard["world"])) if shard else (0, 1)
        key = (m, int(n_candidates), int(horizon), int(reward_kind), float(dt), float(discount), int(set_mode), int(first_set),
               int(n_sets), int(kernel), low64.tobytes(), high64.tobytes(), int(seed), sampler, rank, world,
               id(window) if window is not None else 0, float(inner_lr))
        plan = self._plans.get(key)
        if plan is None:
            from .parallel import shard_bounds
            lo, hi = shard_bounds(n_candidates, rank, world)
            p = N.RolloutParams()
            p.n_candidates, p.n_envs, p.horizon = int(hi - lo), int(m), int(horizon)
            p.set_mode, p.first_set, p.n_sets = int(set_mode), int(first_set), int(n_sets)
            p.reward_kind, p.dt, p.kernel = int(reward_kind), float(dt), int(kernel)
            o = N.PlanOpts()
            o.sampler = {"philox": N.SAMPLER_PHILOX, "mt19937": N.SAMPLER_MT19937}[sampler]
            o.shard_rank, o.shard_world, o.n_candidates_total, o.shard_offset = rank, world, int(n_candidates), int(lo)
            o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            handle = C.c_void_p()
            N.check(self.lib.l2a_plan_create_ex(self._ctx, self._model, C.byref(p), float(discount),
                                                low64.ctypes.data_as(C.c_void_p), high64.ctypes.data_as(C.c_void_p),
                                                C.byref(o), C.byref(handle)))
            plan = dict(handle=handle, act=np.empty((m, self.act_dim), np.float64), ret=np.empty(m, np.float32),
                        idx=np.empty(m, np.int64), shape=(int(horizon), int(hi - lo) * m, self.act_dim),
                        key=np.empty(624, np.uint32), pos=C.c_int32(0), io=N.PlanIO(), peers=[])
            io = plan["io"]
            io.act_out, io.ret_out, io.idx_out = plan["act"].ctypes.data, plan["ret"].ctypes.data, plan["idx"].ctypes.data
            io.mt_key, io.mt_pos = plan["key"].ctypes.data, C.addressof(plan["pos"])
            if world > 1:
                self._attach_peers(plan, rank, world, shard["all_gather"])
            if window is not None:
                N.check(self.lib.l2a_plan_attach_window(self._ctx, handle, window._h, float(inner_lr), 0, 1))
                plan["window"] = window                    # keeps it alive
            self._plans[key] = plan
        plan["io"].flags = int(flags) if window is not None else 0
        io = plan["io"]
        fast = numpy_mt19937_state() if sampler == "mt19937" else None
        if fast is not None:
            # uniform draws touch only (key, pos): the call reads and advances numpy's own state struct in place
            bg, addr = fast
            io.mt_key, io.mt_pos = addr, addr + 624 * 4
            with bg.lock:
                N.check(self.lib.l2a_plan_run_ex(self._ctx, plan["handle"], obs.ctypes.data_as(C.c_void_p), C.byref(io), _stream()))
        else:
            if sampler == "mt19937":
                st = np.random.get_state()
                assert st[0] == "MT19937"
                plan["key"][:] = st[1]
                plan["pos"].value = int(st[2])
                io.mt_key, io.mt_pos = plan["key"].ctypes.data, C.addressof(plan["pos"])
            N.check(self.lib.l2a_plan_run_ex(self._ctx, plan["handle"], obs.ctypes.data_as(C.c_void_p), C.byref(io), _stream()))
            if sampler == "mt19937":
                np.random.set_state(("MT19937", plan["key"], int(plan["pos"].value), st[3], st[4]))
        self._last_plan = plan
        return plan["act"].copy(), plan["ret"].copy(), plan["idx"].copy()

    def plan_cem_host(self, observations, n_candidates, horizon, reward_kind, dt, low, high, num_cem_iters, num_elites, alpha,
                      discount=1.0, set_mode=N.SETS_SHARED, first_set=0, n_sets=1, kernel=N.KERNEL_AUTO, seed=0, sampler="philox",
                      compat=True, window=None, flags=0, inner_lr=0.0):
        """The whole CEM planning call (policies/mpc_controller.py:71-106, all `num_cem_iters` iterations) as ONE C call / one CUDA
        graph replay: observations float64 [m, D] -> (actions float64 [m, A], best_ret [m], best_idx [m], mean [m, H*A], std [m, H*A]).
        sampler "mt19937" continues numpy's global stream exactly like np.random.normal(size=(n, m, H*A)) per iteration (:85),
        including the cached second value of the polar method; "philox" draws on the device stream."""
        obs = np.ascontiguousarray(observations, dtype=np.float64)
        m = obs.shape[0]
        assert obs.shape == (m, self.obs_dim)
        low64 = np.ascontiguousarray(low, dtype=np.float64)
        high64 = np.ascontiguousarray(high, dtype=np.float64)
        ha = int(horizon) * self.act_dim
        key = ("cem", m, int(n_candidates), int(horizon), int(reward_kind), float(dt), float(discount), int(set_mode), int(first_set),
               int(n_sets), int(kernel), low64.tobytes(), high64.tobytes(), int(seed), sampler, int(num_cem_iters), int(num_elites),
               float(alpha), bool(compat), id(window) if window is not None else 0, float(inner_lr))
        plan = self._plans.get(key)
        if plan is None:
            p = N.RolloutParams()
            p.n_candidates, p.n_envs, p.horizon = int(n_candidates), int(m), int(horizon)
            p.set_mode, p.first_set, p.n_sets = int(set_mode), int(first_set), int(n_sets)
            p.reward_kind, p.dt, p.kernel = int(reward_kind), float(dt), int(kernel)
            o = N.PlanOpts()
            o.sampler = {"philox": N.SAMPLER_PHILOX, "mt19937": N.SAMPLER_MT19937}[sampler]
            o.shard_rank, o.shard_world, o.n_candidates_total, o.shard_offset = 0, 1, int(n_candidates), 0
            o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            o.planner, o.cem_iters, o.cem_num_elites = N.PLANNER_CEM, int(num_cem_iters), int(num_elites)
            o.cem_compat, o.cem_alpha = (1 if compat else 0), float(alpha)
            handle = C.c_void_p()
            N.check(self.lib.l2a_plan_create_ex(self._ctx, self._model, C.byref(p), float(discount),
                                                low64.ctypes.data_as(C.c_void_p), high64.ctypes.data_as(C.c_void_p),
                                                C.byref(o), C.byref(handle)))
            plan = dict(handle=handle, act=np.empty((m, self.act_dim), np.float64), ret=np.empty(m, np.float32),
                        idx=np.empty(m, np.int64), shape=(int(n_candidates), m, ha), key=np.empty(624, np.uint32), pos=C.c_int32(0),
                        has_gauss=C.c_int32(0), cached=C.c_double(0.0), mean=np.empty((m, ha), np.float64),
                        std=np.empty((m, ha), np.float64), io=N.PlanIO(), peers=[])
            io = plan["io"]
            io.act_out, io.ret_out, io.idx_out = plan["act"].ctypes.data, plan["ret"].ctypes.data, plan["idx"].ctypes.data
            io.mt_key, io.mt_pos = plan["key"].ctypes.data, C.addressof(plan["pos"])
            io.mt_has_gauss, io.mt_cached = C.addressof(plan["has_gauss"]), C.addressof(plan["cached"])
            io.cem_mean_out, io.cem_std_out = plan["mean"].ctypes.data, plan["std"].ctypes.data
            if window is not None:
                N.check(self.lib.l2a_plan_attach_window(self._ctx, handle, window._h, float(inner_lr), 0, 1))
                plan["window"] = window
            self._plans[key] = plan
        plan["io"].flags = int(flags) if window is not None else 0
        if sampler == "mt19937":
            st = np.random.get_state()
            assert st[0] == "MT19937"
            plan["key"][:] = st[1]
            plan["pos"].value, plan["has_gauss"].value, plan["cached"].value = int(st[2]), int(st[3]), float(st[4])
        N.check(self.lib.l2a_plan_run_ex(self._ctx, plan["handle"], obs.ctypes.data_as(C.c_void_p), C.byref(plan["io"]), _stream()))
        if sampler == "mt19937":
            np.random.set_state(("MT19937", plan["key"], int(plan["pos"].value), int(plan["has_gauss"].value), float(plan["cached"].value)))
        self._last_plan = plan
        return plan["act"].copy(), plan["ret"].copy(), plan["idx"].copy(), plan["mean"].copy(), plan["std"].copy()

    def exchange_resident(self, res, out=None):
        """Peer-memory exchange of device-resident per-rank rollout results `res` (dict from rollout()) through the sharded plan
        of the most recent plan_rs_host call: float64 [m, 2 + A] winner records (return, global index, action) on every rank."""
        plan = self._last_plan
        m = res["best_ret"].shape[0]
        if out is None:
            out = torch.empty((m, 2 + self.act_dim), device=self.device, dtype=torch.float64)
        N.check(self.lib.l2a_plan_exchange_resident(self._ctx, plan["handle"], _ptr(res["best_ret"]), _ptr(res["best_idx"]),
                                                    _ptr(res["best_act"]), _ptr(out), _stream()))
        return out

    def _attach_peers(self, plan, rank, world, all_gather):
        """Exchange the CUDA IPC handles of the ranks' exchange buffers once and hand the peers' pointers to the plan."""
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        N.check(self.lib.l2a_plan_exchange_buffer(plan["handle"], C.byref(ptr), C.byref(nbytes)))
        mine = C.create_string_buffer(64)
        N.check(self.lib.l2a_ipc_get_handle(self._ctx, ptr, mine))
        handles = all_gather(mine.raw)
        assert len(handles) == world
        ptrs = (C.c_void_p * world)()
        for g, h in enumerate(handles):
            if g == rank:
                ptrs[g] = ptr.value
                continue
            peer = C.c_void_p()
            N.check(self.lib.l2a_ipc_open_handle(self._ctx, C.create_string_buffer(bytes(h), 64), C.byref(peer)))
            ptrs[g] = peer.value
            plan["peers"].append(peer)
        N.check(self.lib.l2a_plan_attach_peers(self._ctx, plan["handle"], ptrs))

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

    def last_plan_returns(self, m, n):
        """[m, N] float32 returns of the last iteration of the most recent plan_cem_host call (tests / diagnostics)."""
        out = np.empty((m, n), np.float32)
        N.check(self.lib.l2a_plan_copy_returns(self._ctx, self._last_plan["handle"], out.ctypes.data_as(C.c_void_p)))
        return out

    def last_plan_io_bytes(self):
        """(host->device, device->host) bytes copied inside every call of the most recent plan."""
        a, b = C.c_uint64(), C.c_uint64()
        N.check(self.lib.l2a_plan_io_bytes(self._last_plan["handle"], C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

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
