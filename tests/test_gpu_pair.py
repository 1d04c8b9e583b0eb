"""GPU parity of the CTA-pair rollout kernel (tcgen05.mma.cta_group::2, csrc/rollout_tc2.cuh) through the C ABI: against the
CPU oracle at small and BASELINE shapes (1e-4 relative per return, north_star) and against the single-CTA tcgen05 kernel
(the two run the same split-bf16 arithmetic, so their returns agree far inside the parity tolerance).

Covers every weight-set mode (shared / per-env / ensemble mean through the L2 flag exchange), all three candidates-per-CTA
instances (72 / 48 / 32), both state-register instances (obs <= 24 / <= 48), packed and unpacked layer 0, 256- and 512-wide
hidden layers, ragged last tiles (peer CTA without candidates), and the three reward families."""
import os

import numpy as np
import pytest
import torch

from oracle import mpc_oracle as O
from learning_to_adapt_b200 import _native as N
from tests.helpers import assert_argmax_consistent, assert_returns_close, dev, error_report, make_engine

pytestmark = pytest.mark.gpu

PAIR, TC, SIMT = N.KERNEL_TCGEN05_PAIR, N.KERNEL_TCGEN05, N.KERNEL_SIMT
MODES = {"shared": N.SETS_SHARED, "per_env": N.SETS_PER_ENV, "ensemble": N.SETS_ENSEMBLE_MEAN}


def _run(eng, prob, actions, n, h, mode, n_sets, kernel, discount=1.0):
    res = eng.rollout(dev(prob["obs0"]), dev(actions), n, h, prob["reward_kind"], prob["dt"], discount=discount,
                      set_mode=MODES[mode], first_set=0, n_sets=n_sets, kernel=kernel)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in res.items()}


def _with_nc(nc):
    class _Ctx(object):
        def __enter__(self):
            self.old = os.environ.get("L2A_TC2_NC")
            if nc:
                os.environ["L2A_TC2_NC"] = str(nc)
            elif self.old is not None:
                del os.environ["L2A_TC2_NC"]

        def __exit__(self, *a):
            if self.old is None:
                os.environ.pop("L2A_TC2_NC", None)
            else:
                os.environ["L2A_TC2_NC"] = self.old
    return _Ctx()


@pytest.mark.parametrize("env,hidden,mode,n_sets,m,n,h,nc", [
    ("half_cheetah", (256, 256), "shared", 1, 1, 64, 3, 32),          # one full tile, packed layer 0, one M-block per layer
    ("half_cheetah", (512, 512), "shared", 1, 1, 100, 4, 32),         # two tiles, ragged peer (100 = 64 + 32 + 4)
    ("half_cheetah", (512, 512), "shared", 1, 2, 150, 5, 48),         # two envs, NC = 48, peer CTA of the last tile nearly empty
    ("half_cheetah", (512, 512, 512), "shared", 1, 1, 200, 6, 72),    # NC = 72 (a 16-candidate block straddles the two CTAs)
    ("half_cheetah", (512, 256), "shared", 1, 1, 145, 4, 72),         # mixed widths; the peer of the 2nd tile has 1 candidate
    ("ant", (512, 512), "shared", 1, 1, 130, 4, 48),                  # unpacked layer 0, obs 41 (48-wide state instance)
    ("arm_7dof", (256, 512), "shared", 1, 1, 90, 5, 32),
    ("half_cheetah", (512, 512, 512), "per_env", 3, 3, 120, 5, 48),   # GrBAL: env k uses set k
    ("half_cheetah", (512, 512), "ensemble", 5, 1, 150, 6, 72),       # ensemble mean of 5 pairs through the L2 flag exchange
    ("ant", (512, 512, 512), "ensemble", 5, 1, 100, 5, 32),
    ("half_cheetah", (256, 256), "ensemble", 8, 2, 70, 4, 32),        # 8 members, two envs
    ("half_cheetah", (512, 512), "ensemble", 3, 1, 1, 1, 32),         # one candidate, one step
])
def test_pair_kernel_vs_oracle(env, hidden, mode, n_sets, m, n, h, nc):
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=n_sets, m=m, seed=3)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(9, prob["low"], prob["high"], h, n * m)
    want = O.rollout_returns(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 0.97, mode)
    with _with_nc(nc):
        got = _run(eng, prob, actions, n, h, mode, n_sets, PAIR, discount=0.97)
    rep = assert_returns_close(got["returns"], want)
    assert_argmax_consistent(got["best_idx"], want)
    np.testing.assert_array_equal(got["best_ret"], got["returns"][range(m), got["best_idx"]])
    print("pair vs oracle", env, hidden, mode, "nc", nc, rep)


@pytest.mark.parametrize("cfg", ["headline", "cfg1", "cfg2", "cfg3", "cfg4"])
def test_pair_kernel_at_baseline_configs(cfg):
    """Full BASELINE sizes with the automatic tile choice: pair kernel == single-CTA kernel within 2e-5 relative on every return
    (same arithmetic, different accumulation grouping), and vs the oracle on a candidate stride."""
    env, hidden, mode, n_sets, m, n, h = {
        "headline": ("half_cheetah", (512, 512, 512), "ensemble", 5, 1, 2000, 20),
        "cfg1": ("half_cheetah", (512, 512), "shared", 1, 1, 500, 10),
        "cfg2": ("half_cheetah", (512, 512, 512), "per_env", 5, 5, 1000, 15),
        "cfg3": ("ant", (512, 512, 512), "ensemble", 5, 1, 2000, 20),
        "cfg4": ("half_cheetah", (512, 512), "shared", 1, 1, 5000, 30),
    }[cfg]
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=n_sets, m=m, seed=17)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(23, prob["low"], prob["high"], h, n * m)
    with _with_nc(0):
        got = _run(eng, prob, actions, n, h, mode, n_sets, PAIR)
    ref = _run(eng, prob, actions, n, h, mode, n_sets, TC)
    rep = error_report(got["returns"], ref["returns"])
    print("pair vs single-CTA", cfg, rep)
    assert rep["rel"] <= 2e-5, rep
    sub = np.arange(0, n, max(1, n // 40))
    rows = np.concatenate([e * n + sub for e in range(m)])
    want = O.rollout_returns(prob["obs0"], actions[:, rows], prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, mode)
    rep = assert_returns_close(got["returns"][:, sub], want)
    print("pair vs oracle", cfg, rep)
    np.testing.assert_array_equal(got["best_idx"], np.argmax(got["returns"], axis=1))


def test_pair_kernel_repeatable_and_unsupported_shape():
    """Two launches give bit-identical returns (the flag exchange leaves no state behind); a 128-wide layer is refused loudly."""
    prob = O.make_problem("half_cheetah", hidden_sizes=(512, 512), n_sets=5, m=1, seed=5)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(2, prob["low"], prob["high"], 8, 300)
    a = _run(eng, prob, actions, 300, 8, "ensemble", 5, PAIR)
    b = _run(eng, prob, actions, 300, 8, "ensemble", 5, PAIR)
    np.testing.assert_array_equal(a["returns"], b["returns"])
    prob = O.make_problem("half_cheetah", hidden_sizes=(128, 128), n_sets=1, m=1, seed=5)
    eng = make_engine(prob)
    actions = O.sample_rs_actions(2, prob["low"], prob["high"], 3, 40)
    with pytest.raises(Exception, match="256"):
        _run(eng, prob, actions, 40, 3, "shared", 1, PAIR)
