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


def error_report(got, want):
    """Error figures of a batch of returns: `scaled` = max |err| / max(1, max |want|) (the round-1 figure), `rel` = max
    per-element |err| / max(|want|, median |want|) (north_star's "1e-4 rel", with the batch median as the floor that keeps
    near-zero returns meaningful), `rel_raw_p99` = 99th percentile of the unfloored per-element relative error."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    err = np.abs(got - want)
    scale = max(1.0, float(np.max(np.abs(want))))
    floor = max(float(np.median(np.abs(want))), 1e-30)
    rel = err / np.maximum(np.abs(want), floor)
    raw = err / np.maximum(np.abs(want), 1e-30)
    return dict(scaled=float(err.max() / scale), rel=float(rel.max()), rel_raw_p99=float(np.percentile(raw, 99)),
                floor=floor, scale=scale)


def assert_returns_close(got, want, rtol=RTOL):
    """Per-element relative check: |got - want| <= rtol * max(|want|, median |want| of the batch)."""
    rep = error_report(got, want)
    assert np.all(np.isfinite(np.asarray(got, np.float64)) | ~np.isfinite(np.asarray(want, np.float64))), "non-finite returns"
    assert rep["rel"] <= rtol, "max per-element relative error %.3e > %.1e (scaled %.3e, floor %.3e)" % (
        rep["rel"], rtol, rep["scaled"], rep["floor"])
    return rep


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
