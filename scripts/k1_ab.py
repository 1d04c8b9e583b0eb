"""Development probe: A/B of libl2a_b200 build variants on ONE box (box-to-box variation is +-2 %).

    python scripts/k1_ab.py [cfg,cfg,...] lib1.so lib2.so ...      (run under gpurun; `default` = the in-tree library;
                                                                    lib@3 times the CTA-pair kernel, lib@2 the single-CTA one)

Per variant (child process with L2A_B200_LIB set): device-resident time of the tcgen05 rollout at the named configs, CUDA
events, L2 flushed between calls, median of 15, plus the max per-element relative error against the oracle on a strided
sample of the headline's candidates (parity guard for every variant)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CFGS = {
    "headline": ("half_cheetah", (512, 512, 512), 2000, 20, 1, 5, 2),
    "cfg1": ("half_cheetah", (512, 512), 500, 10, 1, 1, 0),
    "cfg1p": ("half_cheetah", (512, 512), 2000, 20, 10, 1, 0),
    "cfg2i": ("half_cheetah", (512, 512, 512), 1000, 15, 5, 5, 1),
    "cfg3": ("ant", (512, 512, 512), 2000, 20, 1, 5, 2),
    "cfg4rs": ("half_cheetah", (512, 512), 5000, 30, 1, 1, 0),
    "cfg5pergpu": ("ant", (512, 512, 512), 4096, 25, 1, 5, 2),
}

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import mpc_oracle as O
    from learning_to_adapt_b200.engine import PlanningEngine
    from tests.helpers import error_report
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name in sys.argv[2].split(","):
        env, hidden, n, h, m, nsets, mode = CFGS[name]
        prob = O.make_problem(env, hidden_sizes=hidden, n_sets=nsets, m=m, seed=0)
        eng = PlanningEngine(prob["obs_dim"], prob["act_dim"], hidden, n_sets=nsets)
        for i, p in enumerate(prob["param_sets"]):
            eng.set_params(i, p)
        eng.set_normalization(prob["norm"])
        obs = eng._f32(prob["obs0"])
        acts_np = O.sample_rs_actions(3, prob["low"], prob["high"], h, n * m)
        acts = eng._f32(acts_np)
        fn = lambda wr=False: eng.rollout(obs, acts, n, h, prob["reward_kind"], prob["dt"], set_mode=mode, first_set=0, n_sets=nsets,
                                          want_returns=wr, kernel=int(os.environ.get('L2A_AB_KERNEL', '2')))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for i in range(15):
            flush.fill_(i)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        row = dict(cfg=name, lib=os.path.basename(os.environ.get("L2A_B200_LIB", "default")), kernel=int(os.environ.get("L2A_AB_KERNEL", "2")), ms=float(np.median(ts)), ms_min=float(np.min(ts)))
        if name == "headline":
            res = fn(True)
            torch.cuda.synchronize()
            sub = np.arange(0, n, 40)
            want = O.rollout_returns(prob["obs0"], acts_np[:, sub], prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "ensemble")
            rep = error_report(res["returns"].cpu().numpy()[:, sub], want)
            row.update(rel_err=rep["rel"], scaled_err=rep["scaled"])
        print(json.dumps(row), flush=True)
        eng.close()
else:
    cfgs = sys.argv[1]
    for lib in sys.argv[2:]:
        env = dict(os.environ)
        if "@" in lib:                                   # lib@kernel: 2 = single-CTA tcgen05, 3 = CTA pair
            lib, env["L2A_AB_KERNEL"] = lib.split("@")
        if lib != "default":
            env["L2A_B200_LIB"] = os.path.abspath(lib)
        else:
            env.pop("L2A_B200_LIB", None)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", cfgs], env=env)
