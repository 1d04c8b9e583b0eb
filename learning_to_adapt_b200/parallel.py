"""Candidate sharding across the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

The planner is embarrassingly parallel over candidates: rank g rolls its own slice of every env's N candidates
through the fused kernel (weights, statistics and observations replicated), producing per env the triple
(best return, global candidate index, first action).  The only exchange is ONE all-gather of ``m x (3 + A)`` floats
per planning call, followed by a local argmax with lowest-global-index tie-break -- identical to ``np.argmax`` over
the concatenated returns (policies/mpc_controller.py:128-129), so G ranks return exactly what G = 1 returns for the
same candidate tensor.  The reference has no multi-device path at all; this is new.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_candidates, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of the N candidates of every env owned by `rank`."""
    base, rem = divmod(int(n_candidates), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def pack_best(best_ret, best_idx_global, best_act):
    """[m, 3 + A] float32: (return, global index >> 16, global index & 0xFFFF, action...): the index travels as two halves that
    are exact in fp32."""
    m = best_ret.shape[0]
    idx = best_idx_global.to(torch.int64)
    hi = (idx >> 16).to(torch.float32)
    lo = (idx & 0xFFFF).to(torch.float32)
    return torch.cat([best_ret.reshape(m, 1).to(torch.float32), hi.reshape(m, 1), lo.reshape(m, 1),
                      best_act.to(torch.float32)], dim=1).contiguous()


def select_best(gathered):
    """gathered [G, m, 3 + A] -> (best_ret [m], best_idx_global [m] int64, best_act [m, A]) with np.argmax semantics:
    maximum return, NaN beats numbers, ties -> lowest global candidate index."""
    ret = gathered[:, :, 0]
    idx = (gathered[:, :, 1].to(torch.int64) << 16) | gathered[:, :, 2].to(torch.int64)
    G, m = ret.shape
    isnan = torch.isnan(ret)
    any_nan = isnan.any(dim=0)
    key = torch.where(isnan, torch.full_like(ret, float("inf")), ret)             # NaN ranks above every number
    key = torch.where(any_nan.unsqueeze(0) & ~isnan, torch.full_like(ret, -float("inf")), key)
    best_key = key.max(dim=0).values
    cand = key == best_key.unsqueeze(0)
    big = torch.iinfo(torch.int64).max
    idx_masked = torch.where(cand, idx, torch.full_like(idx, big))
    win = idx_masked.argmin(dim=0)                                                 # rank holding the winner, per env
    ar = torch.arange(m, device=gathered.device)
    return ret[win, ar], idx[win, ar], gathered[win, ar, 3:]


class CandidateShard(object):
    """Plug into ``MPCController(parallel=CandidateShard(...))``: every rank calls ``get_actions`` with the same
    observations and gets the same (globally best) actions back."""

    def __init__(self, group=None):
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def all_gather_bytes(self, payload):
        """Host-side exchange of a small bytes object (the CUDA IPC handles of the plans' exchange buffers; once per plan)."""
        out = [None] * self.world_size
        dist.all_gather_object(out, bytes(payload), group=self.group)
        return out

    def all_gather_best(self, packed):
        out = torch.empty((self.world_size * packed.shape[0],) + tuple(packed.shape[1:]), device=packed.device,
                          dtype=packed.dtype)
        dist.all_gather_into_tensor(out, packed, group=self.group)       # concatenated along dim 0 (NCCL and gloo)
        return out.view((self.world_size,) + tuple(packed.shape))

    def combine(self, best_ret, best_idx_local, best_act, shard_lo, engine=None):
        """Per-rank best triples -> global winner on every rank.  With an engine (CUDA tensors) the pack / select steps are
        the library's two small kernels around the one NCCL all-gather; without (CPU tensors, gloo tests) the same logic
        runs as torch ops (pack_best / select_best)."""
        if engine is not None and best_ret.is_cuda:
            packed = engine.shard_pack(best_ret, best_idx_local, best_act, shard_lo)
            return engine.shard_select(self.all_gather_best(packed))
        packed = pack_best(best_ret, best_idx_local.to(torch.int64) + int(shard_lo), best_act)
        return select_best(self.all_gather_best(packed))

    def plan_rs(self, ctrl, observations, obs_dev, set_mode, first_set, n_sets):
        """Random shooting with the candidates of every env split over the ranks.
        sampler="numpy": every rank draws the reference's full tensor from the same global numpy stream and rolls its
        slice (so a seeded run is bit-identical to the single-GPU / reference result);
        sampler="device": every rank draws only its own slice with its own Philox stream."""
        eng = ctrl.dynamics_model._engine
        n, m, h = ctrl.n_candidates, len(observations), ctrl.horizon
        act_dim = ctrl.action_space.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world_size)
        n_loc = hi - lo
        if ctrl.sampler in ("numpy", "numpy_host"):
            a_full = ctrl.get_random_action(h * n * m).reshape((h, m, n, act_dim))
            a_dev = eng._f32(np.ascontiguousarray(a_full[:, :, lo:hi]).reshape(h, m * n_loc, act_dim))
        else:
            self._calls = getattr(self, "_calls", 0) + 1
            a_dev = eng.sample_uniform(ctrl.action_space.low, ctrl.action_space.high, h * m * n_loc,
                                       seed=(int(getattr(ctrl, "seed", 0)) << 8) + self.rank, call_index=self._calls)
            a_dev = a_dev.view(h, m * n_loc, act_dim)
        res = eng.rollout(obs_dev, a_dev, n_loc, h, ctrl._reward_kind, ctrl._dt, discount=ctrl.discount,
                          set_mode=set_mode, first_set=first_set, n_sets=n_sets, layout="thra", want_returns=False,
                          kernel=ctrl.kernel)
        best_ret, best_idx, best_act = self.combine(res["best_ret"], res["best_idx"], res["best_act"], lo, engine=eng)
        ctrl.last_plan = dict(best_ret=best_ret, best_idx=best_idx, best_act=best_act, returns=None)
        if ctrl.sampler in ("numpy", "numpy_host"):
            idx = best_idx.cpu().numpy()
            return a_full[0][range(m), idx]
        return best_act.cpu().numpy().astype(np.float64)
