"""Development probe: rollout time per BASELINE config for every candidates-per-CTA variant (L2A_TC_NC override)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import mpc_oracle as O
    from learning_to_adapt_b200.engine import PlanningEngine
    cfgs = [("cfg1", "half_cheetah", (512, 512), 500, 10, 1, 1, 0), ("cfg1p", "half_cheetah", (512, 512), 2000, 20, 10, 1, 0),
            ("cfg2i", "half_cheetah", (512, 512, 512), 1000, 15, 5, 5, 1), ("cfg2ii", "half_cheetah", (512, 512, 512), 1000, 15, 1, 5, 2),
            ("headline", "half_cheetah", (512, 512, 512), 2000, 20, 1, 5, 2), ("cfg3", "ant", (512, 512, 512), 2000, 20, 1, 5, 2),
            ("cfg4rs", "half_cheetah", (512, 512), 5000, 30, 1, 1, 0), ("cfg5pergpu", "ant", (512, 512, 512), 4096, 25, 1, 5, 2)]
    for name, env, hidden, n, h, m, nsets, mode in cfgs:
        prob = O.make_problem(env, hidden_sizes=hidden, n_sets=nsets, m=m, seed=0)
        eng = PlanningEngine(prob["obs_dim"], prob["act_dim"], hidden, n_sets=nsets)
        for i, p in enumerate(prob["param_sets"]):
            eng.set_params(i, p)
        eng.set_normalization(prob["norm"])
        obs = eng._f32(prob["obs0"])
        low, high = eng._f32(prob["low"]), eng._f32(prob["high"])
        acts = torch.rand((h, n * m, prob["act_dim"]), device="cuda") * (high - low) + low
        fn = lambda: eng.rollout(obs, acts, n, h, prob["reward_kind"], prob["dt"], set_mode=mode, first_set=0, n_sets=nsets,
                                 want_returns=False, kernel=2)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(json.dumps(dict(cfg=name, nc=os.environ.get("L2A_TC_NC", "auto"), ms=float(np.median(ts)), rollouts_per_s=n * m / np.median(ts) * 1e3)), flush=True)
        eng.close()
else:
    for nc in ("auto", "80", "64", "48", "32"):
        env = dict(os.environ)
        if nc != "auto":
            env["L2A_TC_NC"] = nc
        else:
            env.pop("L2A_TC_NC", None)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
