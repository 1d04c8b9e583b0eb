"""Development probe: ReBAL (LSTM) planner rollout timing at the run_rebal.py shape vs the oracle port on the host."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from learning_to_adapt_b200.dynamics.rnn_dynamics import RNNDynamicsModel  # noqa: E402
from learning_to_adapt_b200.envs.synthetic import SyntheticEnv  # noqa: E402

for (n, h, m, hs) in ((500, 10, 5, 256), (2000, 20, 5, 256)):
    prob = O.make_problem("half_cheetah", hidden_sizes=(32,), n_sets=1, m=m, seed=8)
    params = O.xavier_rnn_params(np.random.RandomState(3), prob["obs_dim"] + prob["act_dim"], hs, prob["obs_dim"], out_scale=0.1)
    model = RNNDynamicsModel("dyn", SyntheticEnv("half_cheetah"), hidden_sizes=(hs,))
    model.set_params(params)
    model.set_normalization(prob["norm"])
    hidden = model.get_initial_hidden(m)
    obs = model._f32(prob["obs0"])
    low, high = model._f32(prob["low"]), model._f32(prob["high"])
    acts = torch.rand((h, n * m, prob["act_dim"]), device="cuda") * (high - low) + low
    fn = lambda: model.rollout(obs, hidden, acts, n, h, prob["reward_kind"], prob["dt"])
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    a_host = acts.cpu().numpy().astype(np.float64)
    t0 = time.perf_counter()
    O.rnn_rollout_returns(prob["obs0"], a_host, (hidden.c, hidden.h), params, prob["norm"], prob["reward_kind"], prob["dt"])
    cpu_s = time.perf_counter() - t0
    flops = 2.0 * ((prob["obs_dim"] + prob["act_dim"] + hs) * 4 * hs + hs * prob["obs_dim"]) * n * m * h
    print(json.dumps(dict(cfg="rebal N=%d H=%d m=%d LSTM(%d)" % (n, h, m, hs), gpu_ms=float(np.median(ts)),
                          rollouts_per_s=n * m / np.median(ts) * 1e3, tflops=flops / np.median(ts) * 1e-9,
                          cpu_oracle_ms=cpu_s * 1e3, cpu_rollouts_per_s=n * m / cpu_s)), flush=True)
