"""Development probe: K2 (GrBAL adapt) and the full GrBAL step (switch_to_pre_adapt + adapt + get_actions) timing."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel  # noqa: E402
from learning_to_adapt_b200.envs.synthetic import SyntheticEnv  # noqa: E402
from learning_to_adapt_b200.policies.mpc_controller import MPCController  # noqa: E402

for env_name, K, M, n, h in (("half_cheetah", 5, 16, 1000, 15), ("ant", 5, 16, 2000, 20), ("half_cheetah", 10, 16, 500, 10)):
    prob = O.make_problem(env_name, hidden_sizes=(512, 512, 512), n_sets=1, m=K, seed=0)
    env = SyntheticEnv(env_name)
    model = MetaMLPDynamicsModel("dyn", env, hidden_sizes=(512, 512, 512), meta_batch_size=10, inner_learning_rate=1e-3)
    model.set_params(prob["param_sets"][0])
    model.set_normalization(prob["norm"])
    ctx = O.make_adapt_context(4, prob, K, M)
    eng = model._engine
    xs, ts = [], []
    for o, a, nx in zip(*ctx):
        xs.append(np.concatenate([O.normalize(o, *prob["norm"]["obs"]), O.normalize(a, *prob["norm"]["act"])], axis=1))
        ts.append(O.normalize(nx - o, *prob["norm"]["delta"]))
    x, t = eng._f32(np.stack(xs)), eng._f32(np.stack(ts))
    fn = lambda: eng.adapt(x, t, 1e-3, 0, 1)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tms = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tms.append(s.elapsed_time(e))
    ctrl = MPCController("policy", env, model, n_candidates=n, horizon=h, sampler="device")
    obs = np.array(prob["obs0"])
    def step():
        model.switch_to_pre_adapt()
        model.adapt(*ctx)
        return ctrl.get_actions(obs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 20
    # the same env step with the adaptation windows resident on the device (f3): push + gather + K2 + K1
    win = model.make_adapt_window(K, M)
    rng = np.random.RandomState(1)
    for j in range(M + 2):
        win.push(np.stack([c[min(j, M - 1)] for c in ctx[0]]), np.stack([c[min(j, M - 1)] for c in ctx[1]]))
    act_prev = np.zeros((K, prob["act_dim"]))
    def step_window():
        win.push(obs, act_prev)
        model.switch_to_pre_adapt()
        model.adapt_from_window(win)
        return ctrl.get_actions(obs)
    for _ in range(3):
        step_window()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step_window()
    torch.cuda.synchronize()
    e2e_win = (time.perf_counter() - t0) / 20
    # ... and with the append of (obs, action) folded into the planning call too: the whole env step is ONE C call / graph replay
    ctrl.push_window = win
    def step_fused():
        model.switch_to_pre_adapt()
        model.adapt_from_window(win)
        return ctrl.get_actions(obs)
    for _ in range(3):
        step_fused()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step_fused()
    torch.cuda.synchronize()
    e2e_fused = (time.perf_counter() - t0) / 20
    ctrl.push_window = None
    # host formulation including the list slicing the reference's sampler does each step
    paths = [dict(observations=list(rng.normal(size=(M + 4, prob["obs_dim"]))), actions=list(rng.normal(size=(M + 4, prob["act_dim"])))) for _ in range(K)]
    def step_lists():
        model.switch_to_pre_adapt()
        model.adapt([np.stack(p["observations"][-M - 1:-1]) for p in paths], [np.stack(p["actions"][-M - 1:-1]) for p in paths],
                    [np.stack(p["observations"][-M:]) for p in paths])
        return ctrl.get_actions(obs)
    for _ in range(3):
        step_lists()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step_lists()
    torch.cuda.synchronize()
    e2e_lists = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    O.adapt(*ctx, prob["param_sets"][0], prob["norm"], 1e-3)
    cpu = time.perf_counter() - t0
    print(json.dumps(dict(cfg="%s K=%d M=%d N=%d H=%d" % (env_name, K, M, n, h), adapt_kernels_ms=float(np.median(tms)),
                          grbal_step_e2e_ms=e2e * 1e3, grbal_step_lists_ms=e2e_lists * 1e3, grbal_step_device_window_ms=e2e_win * 1e3, grbal_step_one_call_ms=e2e_fused * 1e3, cpu_oracle_adapt_ms=cpu * 1e3)), flush=True)
