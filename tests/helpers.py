"""Shared helpers for the GPU parity tests (tests only; the product never imports oracle/)."""
import numpy as np
import torch

from oracle import mpc_oracle as O

RTOL = 1e-4   # north_star: "within 1e-4 rel fp32"


def make_engine(prob, n_sets=None, extra_sets=0):
    from learning_to_adapt_b200.engine import PlanningEngine
    sets = prob["param_sets"]
    eng = PlanningEngine(prob["obs_dim"], prob["act_dim"], prob["hidden_sizes"], n_sets=(n_sets or len(sets)) + extra_sets)
    for i, p in enumerate(sets):
        eng.set_params(i, p)
    eng.set_normalization(prob["norm"])
    return eng


def assert_returns_close(got, want, rtol=RTOL):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale = max(1.0, float(np.max(np.abs(want))))
    err = np.abs(got - want)
    tol = rtol * np.maximum(np.abs(want), scale)
    assert np.all(err <= tol), "max err %.3e (tol %.3e, scale %.3e)" % (err.max(), tol.min(), scale)


def assert_argmax_consistent(best_idx, want_returns, rtol=RTOL):
    """Identical argmax unless the oracle's top-2 gap is inside the tolerance (near-tie rule)."""
    want_returns = np.asarray(want_returns, np.float64)
    for e in range(want_returns.shape[0]):
        want = int(np.argmax(want_returns[e]))
        got = int(best_idx[e])
        if got != want:
            gap = want_returns[e, want] - want_returns[e, got]
            scale = max(1.0, float(np.max(np.abs(want_returns[e]))))
            assert gap <= 2 * rtol * scale, "env %d: argmax %d != %d with gap %.3e" % (e, got, want, gap)


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device="cuda")
