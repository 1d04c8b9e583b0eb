"""oracle/device_arith_model.py -- the numpy model of the kernels' arithmetic (float32 state, split-bf16 products with three
tensor-core passes) -- against the reference-faithful oracle on slices of the BASELINE shapes: the choreography stays inside
north_star's 1e-4 relative bar with a wide margin over H dependent steps, and a single bf16 pass would not.  The GPU parity
tests measure the same quantity on the real kernels (profiles/r02_parity_configs.jsonl: 1e-6 .. 1.4e-5)."""
import numpy as np
import pytest
import torch

from oracle import device_arith_model as M
from oracle import mpc_oracle as O
from tests.helpers import RTOL, error_report

SHAPES = {   # env, hidden, E, mode, candidates (slice), H, envs
    "headline": ("half_cheetah", (512, 512, 512), 5, "ensemble", 96, 20, 1),
    "cfg3 ant": ("ant", (512, 512, 512), 5, "ensemble", 64, 20, 1),
    "cfg5 ant H=25": ("ant", (512, 512, 512), 5, "ensemble", 64, 25, 1),
    "cfg4 H=30": ("half_cheetah", (512, 512), 1, "shared", 128, 30, 1),
    "cfg2 per-env": ("half_cheetah", (512, 512, 512), 3, "per_env", 48, 15, 3),
    "arm": ("arm_7dof", (256, 256), 1, "shared", 64, 20, 2),
}


def test_bf16_rounding_is_round_to_nearest_even():
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.normal(size=4096) * 10.0 ** rng.randint(-6, 6, size=4096), [0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.0e38]]).astype(np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(M.to_bf16(x), want)
    hi, lo = M.split(x[:4096])
    assert np.all(np.abs(x[:4096] - (hi + lo)) <= np.abs(x[:4096]) * 2.0 ** -16)        # two halves carry 16 mantissa bits


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_split_bf16_choreography_stays_inside_the_parity_bar(name):
    env, hidden, e, mode, n, h, m = SHAPES[name]
    prob = O.make_problem(env, hidden_sizes=hidden, n_sets=e, m=m, seed=0)
    acts = O.sample_rs_actions(3, prob["low"], prob["high"], h, n * m)
    args = (prob["obs0"], acts, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, mode)
    want = O.rollout_returns(*args)
    rep3 = error_report(M.rollout_returns_model(*args), want)
    rep1 = error_report(M.rollout_returns_model(*args, passes=1), want)
    assert rep3["rel"] <= 0.3 * RTOL, "three-pass model error %.2e" % rep3["rel"]       # measured on the GPU: <= 1.4e-5
    assert rep1["rel"] >= 30 * rep3["rel"], "a single bf16 pass should be visibly worse (%.2e vs %.2e)" % (rep1["rel"], rep3["rel"])
    if name != "arm":
        assert rep1["rel"] > RTOL, "a single bf16 pass would already meet the bar (%.2e)?" % rep1["rel"]
    assert np.array_equal(np.argmax(M.rollout_returns_model(*args), axis=1), np.argmax(want, axis=1))
