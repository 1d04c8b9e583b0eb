"""Development probe (run under gpurun): times both rollout kernels at a few BASELINE configs with CUDA events."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from learning_to_adapt_b200.engine import PlanningEngine  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    out = []
    cfgs = [
        ("cfg1", "half_cheetah", (512, 512), 500, 10, 1, 1, 0),
        ("cfg1p", "half_cheetah", (512, 512), 2000, 20, 10, 1, 0),
        ("cfg2i", "half_cheetah", (512, 512, 512), 1000, 15, 5, 5, 1),
        ("headline", "half_cheetah", (512, 512, 512), 2000, 20, 1, 5, 2),
        ("cfg3", "ant", (512, 512, 512), 2000, 20, 1, 5, 2),
    ]
    for name, env, hidden, n, h, m, nsets, mode in cfgs:
        prob = O.make_problem(env, hidden_sizes=hidden, n_sets=nsets, m=m, seed=0)
        eng = PlanningEngine(prob["obs_dim"], prob["act_dim"], hidden, n_sets=nsets)
        for i, p in enumerate(prob["param_sets"]):
            eng.set_params(i, p)
        eng.set_normalization(prob["norm"])
        obs = eng._f32(prob["obs0"])
        low, high = eng._f32(prob["low"]), eng._f32(prob["high"])
        acts = torch.rand((h, n * m, prob["act_dim"]), device="cuda") * (high - low) + low
        flops = 2.0 * sum(a * b for a, b in zip([prob["obs_dim"] + prob["act_dim"]] + list(hidden), list(hidden) + [prob["obs_dim"]]))
        steps = n * m * h * (nsets if mode == 2 else 1)
        for kernel in (2, 1):
            if kernel == 1 and name in ("cfg1p",):
                continue
            try:
                fn = lambda: eng.rollout(obs, acts, n, h, prob["reward_kind"], prob["dt"], set_mode=mode, first_set=0,
                                         n_sets=nsets, want_returns=False, kernel=kernel)
                med, best = timeit(fn, iters=5 if kernel == 1 else 20)
                rec = dict(cfg=name, kernel=kernel, ms_median=med, ms_min=best, rollouts_per_s=n * m / med * 1e3,
                           dyn_steps_per_s=steps / med * 1e3, tflops=steps * flops / med * 1e-9)
            except Exception as ex:  # noqa
                rec = dict(cfg=name, kernel=kernel, error=str(ex))
            print(json.dumps(rec), flush=True)
            out.append(rec)
        eng.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
