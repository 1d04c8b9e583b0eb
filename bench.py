#!/usr/bin/env python
"""bench.py -- candidate-rollouts/s of the MPC planning hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config NAME] [--scaling weak|strong]    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W [--config NAME]           # the reference algorithm on the host

A "step" is ONE planning call (one pass of the hot path over one batch of candidates).  --config picks the workload
(BASELINE.md section 2); the default is the north-star headline: HalfCheetah random shooting, N=2000 per GPU, H=20, ensemble
E=5 (mean of deltas), MLP 26->512->512->512->20.
    headline  HC RS N=2000 H=20 E=5 (512^3), m=1                   cfg1   HC RS N=500 H=10 (512^2), m=1 (the reference's CPU case)
    cfg1p     HC RS N=2000 H=20 (512^2), m=10 (run_mb_mpc.py)      cfg2   HC GrBAL: 5 envs x N=1000 H=15 (512^3), adapt M=16 every step
    cfg3      Ant RS N=2000 H=20 E=5 (512^3)                       cfg4   HC CEM N=5000, 500 elites, 3 iterations, H=30 (512^2)
    cfg5      Ant RS N=4096 per GPU H=25 E=5 (512^3) (the 8-GPU shard's per-GPU share; --gpus 8 = 32768 candidates)
Scaling: weak (default; every rank plans its own N candidates, global N = N x gpus) or strong (--scaling strong: the config's N
is split over the ranks).

value  : whole-job candidate-rollouts/s with the inputs already resident in HBM (pre-materialised candidate tensor / normal
         draws), CUDA-event timed per step on the launching stream, L2 flushed (256 MB write) between steps, max over ranks.
         At N>1 the step includes the exchange of the per-rank best triple: the library's peer-memory kernel over NVLink (the
         same kernel the e2e call runs inside its graph); `exchange.ms_per_step_nccl_allgather` is the same step with one NCCL
         all-gather through torch.distributed instead.
e2e    : the same metric through the public API call MPCController.get_actions(obs ndarray) -> ndarray = ONE host-buffer C-ABI
         call (l2a_plan_run_ex): host->device copy of the observations (+ RNG state), candidate sampling on the device, K1
         (cfg4: all CEM iterations; cfg2: window gather + K2 adapt in front, window push behind; N>1: peer-memory exchange of
         the ranks' winners over NVLink), device->host copy of the chosen actions -- all inside the timed region.  Sampler =
         device Philox; `default_sampler_value` = the same with the package's DEFAULT sampler (the reference's numpy stream
         regenerated on the device), which is what unchanged run scripts get.
roofline: dominant kernel rollout_tc2_kernel (the CTA-pair tcgen05 rollout; rollout_tc_kernel with L2A_TC_PAIR=0), tensor-pipe bound;
         achieved = algorithmic FLOPs per launch (rows*H*E*F, counted once although split-bf16 issues 3 MMA passes) / mean launch
         duration measured here with CUDA events.
cpu_baseline / --impl reference: the reference's planner on the host cores: the VERBATIM upstream MPCController (oracle/_ref, made
         by oracle/make_ref.py) when present -- kind "reference" -- else its oracle restatement (kind "port"); the dynamics model
         behind it is the oracle's float32 BLAS port in both cases (TF 1.13.1 cannot be installed; "ensemble" has no upstream code).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

CONFIGS = {
    "headline": dict(env="half_cheetah", hidden=(512, 512, 512), n=2000, h=20, m=1, E=5, mode="ensemble", planner="rs",
                     desc="BASELINE.md headline: HalfCheetah random-shooting MPC, ensemble E=5 (mean of deltas), MLP 26-512-512-512-20"),
    "cfg1": dict(env="half_cheetah", hidden=(512, 512), n=500, h=10, m=1, E=1, mode="shared", planner="rs",
                 desc="BASELINE cfg1: run_mb_mpc.py HalfCheetah random shooting, single MLP 26-512-512-20"),
    "cfg1p": dict(env="half_cheetah", hidden=(512, 512), n=2000, h=20, m=10, E=1, mode="shared", planner="rs",
                  desc="BASELINE cfg1' (run_mb_mpc.py defaults): HalfCheetah random shooting, 10 envs per call, MLP 26-512-512-20"),
    "cfg2": dict(env="half_cheetah", hidden=(512, 512, 512), n=1000, h=15, m=5, E=5, mode="per_env", planner="rs", grbal=True,
                 desc="BASELINE cfg2: HalfCheetah GrBAL, 5 envs each with its own adapted weight set (inner adapt M=16, lr=1e-3, every "
                      "env step), MLP 26-512-512-512-20"),
    "cfg3": dict(env="ant", hidden=(512, 512, 512), n=2000, h=20, m=1, E=5, mode="ensemble", planner="rs",
                 desc="BASELINE cfg3: Ant random shooting, ensemble E=5, MLP 49-512-512-512-41, ctrl +-150"),
    "cfg4": dict(env="half_cheetah", hidden=(512, 512), n=5000, h=30, m=1, E=1, mode="shared", planner="cem", iters=3, pct=0.1, alpha=0.1,
                 desc="BASELINE cfg4: HalfCheetah CEM planner, 5000 candidates x 3 iterations, 500 elites, MLP 26-512-512-20 "
                      "(a call rolls N x iters candidate sequences)"),
    "cfg5": dict(env="ant", hidden=(512, 512, 512), n=4096, h=25, m=1, E=5, mode="ensemble", planner="rs",
                 desc="BASELINE cfg5: Ant candidate shard, N=4096 per GPU (32768 on 8 GPUs), ensemble E=5, MLP 49-512-512-512-41"),
}
METRIC = "candidate-rollouts/sec (N x H dynamics steps) HalfCheetah MPC"
MODE = {"shared": 0, "per_env": 1, "ensemble": 2}


def flops_per_dyn_step(obs_dim, act_dim, hidden):
    sizes = [obs_dim + act_dim] + list(hidden) + [obs_dim]
    return 2.0 * sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))      # HC(512^3) = 1 095 680


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_tflops=float(p["bf16_tflops"]), hbm_gbs=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, burst)")
    return dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


_REAL_STDOUT = None


def claim_stdout():
    """Route everything any library writes to fd 1 (e.g. NCCL's version banner) to stderr; the one JSON line goes to the real
    stdout through emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line_dict):
    data = (json.dumps(line_dict) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def make_problem(cfg):
    from oracle import mpc_oracle as O
    n_sets = cfg["E"] if cfg["mode"] == "ensemble" else 1
    return O, O.make_problem(cfg["env"], hidden_sizes=cfg["hidden"], n_sets=n_sets, m=cfg["m"], seed=0)


def split_candidates(cfg, gpus, scaling):
    """(candidates per GPU, global candidates) of one planning call."""
    gpus = max(1, gpus)
    if scaling == "strong":
        return (cfg["n"] + gpus - 1) // gpus, cfg["n"]
    return cfg["n"], cfg["n"] * gpus


def config_dict(name, cfg, gpus, scaling, sampler):
    n_gpu, n_glob = split_candidates(cfg, gpus, scaling)
    d = dict(workload="%s; N=%d candidates/GPU x H=%d, m=%d env(s)" % (cfg["desc"], n_gpu, cfg["h"], cfg["m"]), name=name,
             n_candidates_per_gpu=n_gpu, global_candidates=n_glob, horizon=cfg["h"], n_envs=cfg["m"], weight_sets=cfg["E"],
             set_mode=cfg["mode"], planner=cfg["planner"],
             dyn_steps_per_call_per_gpu=n_gpu * cfg["m"] * cfg["h"] * (cfg["E"] if cfg["mode"] == "ensemble" else 1) * cfg.get("iters", 1),
             parallelism="candidate-shard x%d, winners exchanged once per call over peer memory (NVLink); NCCL (torch.distributed) for the bench's barriers / max-over-ranks" % max(1, gpus),
             sampler=sampler,
             l2="flushed between timed steps (256 MB write); inputs (candidates + weight tiles, < 40 MB) are smaller than L2")
    if cfg["planner"] == "cem":
        d["cem"] = dict(iters=cfg["iters"], percent_elites=cfg["pct"], alpha=cfg["alpha"], compat=True)
    return d


# =================================================================================================== host (reference) planner
class _OracleModel(object):
    """dynamics_model stand-in for the verbatim reference controller: the oracle's predict for the config's weight-set mode."""

    def __init__(self, O, prob, mode, sets):
        self.O, self.prob, self.mode, self.sets = O, prob, mode, sets

    def predict(self, obs, act):
        O, p = self.O, self.prob
        if self.mode == "ensemble":
            return O.predict_ensemble_mean(obs, act, self.sets, p["norm"])
        if self.mode == "per_env":
            return O.predict_per_task(obs, act, self.sets, p["norm"])
        return O.predict(obs, act, self.sets[0], p["norm"])


def host_planner(cfg, O, prob, n, sets):
    """Returns (callable planning once on the host, kind, description)."""
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from oracle.make_ref import import_reference_controller
    ref_cls = import_reference_controller()
    h, m = cfg["h"], cfg["m"]
    if ref_cls is not None:
        env = SyntheticEnv(cfg["env"])
        kw = dict(n_candidates=n, horizon=h)
        if cfg["planner"] == "cem":
            kw.update(use_cem=True, num_cem_iters=cfg["iters"], percent_elites=cfg["pct"], alpha=cfg["alpha"])
        ctrl = ref_cls("policy", env, _OracleModel(O, prob, cfg["mode"], sets), **kw)
        return (lambda seed: ctrl.get_actions(prob["obs0"]), "reference",
                "verbatim upstream MPCController.get_actions (oracle/_ref) over the oracle's float32 BLAS port of the dynamics model")
    if cfg["planner"] == "cem":
        def plan(seed):
            rng = np.random.RandomState(seed)
            zs = [rng.normal(size=(n, m, h * prob["act_dim"])) for _ in range(cfg["iters"])]
            return O.cem_plan(prob["obs0"], zs, prob["low"], prob["high"], sets, prob["norm"], prob["reward_kind"], prob["dt"], h,
                              cfg["pct"], cfg["alpha"], 1.0, cfg["mode"])
    else:
        def plan(seed):
            actions = O.sample_rs_actions(seed, prob["low"], prob["high"], h, n * m)
            return O.rs_plan(prob["obs0"], actions, sets, prob["norm"], prob["reward_kind"], prob["dt"], 1.0, cfg["mode"])
    return plan, "port", "oracle restatement of MPCController.get_actions (numpy planner) over its float32 BLAS port of the dynamics model"


def host_sets(cfg, O, prob):
    if cfg["mode"] == "per_env":          # GrBAL: per-env adapted sets (the host adapt itself is part of the step below)
        ctx = O.make_adapt_context(4, prob, cfg["m"], 16)
        return O.adapt(*ctx, prob["param_sets"][0], prob["norm"], 1e-3), ctx
    return prob["param_sets"], None


def pick_blas_threads(plan_small):
    """The numpy/OpenBLAS port is not fastest with every host thread (a 128-core box oversubscribes 2000x512 GEMMs): time one
    small planning call per candidate thread count and keep the best, so the CPU baseline is the host planner at its best."""
    from threadpoolctl import threadpool_limits
    cores = os.cpu_count() or 1
    cands = sorted(set([c for c in (4, 8, 16, 32, 64, cores) if c <= cores]))
    best, best_t = cores, float("inf")
    for c in cands:
        with threadpool_limits(limits=c):
            plan_small(0)
            dt = float("inf")
            for i in range(2):
                t0 = time.perf_counter()
                plan_small(1 + i)
                dt = min(dt, time.perf_counter() - t0)
        if dt < best_t * 0.97:
            best, best_t = c, dt
    return best


def time_host_planner(cfg, n, budget_s, min_calls, max_calls, warmup):
    """Times the host planner on a bounded sample: whole planning calls of the workload when a call takes < 2 s, else calls on a
    fraction of the candidates (per-candidate cost is flat in N: every step is N-row GEMMs)."""
    from threadpoolctl import threadpool_limits
    O, prob = make_problem(cfg)
    sets, ctx = host_sets(cfg, O, prob)
    small, _, _ = host_planner(cfg, O, prob, max(64, min(n, 256)), sets)
    cores = pick_blas_threads(small)
    with threadpool_limits(limits=cores):
        t0 = time.perf_counter()
        small(0)
        per_cand = (time.perf_counter() - t0) / max(64, min(n, 256))
        n_s = n
        while n_s > 250 and per_cand * n_s > 2.0:
            n_s //= 2
        plan, kind, what = host_planner(cfg, O, prob, n_s, sets)

        def step(seed):
            if ctx is not None:
                O.adapt(*ctx, prob["param_sets"][0], prob["norm"], 1e-3)          # GrBAL: the env step adapts first
            return plan(seed)
        for i in range(warmup):
            step(i)
        calls = 0
        t0 = time.perf_counter()
        while calls < min_calls or (time.perf_counter() - t0 < budget_s and calls < max_calls):
            step(100 + calls)
            calls += 1
        dt = (time.perf_counter() - t0) / calls
    rollouts = n_s * cfg["m"] * cfg.get("iters", 1)
    sample = "%d planning calls of N=%d%s, H=%d, m=%d, %d weight set(s) (%s); %s; best of {4..%d} BLAS threads = %d; %.1f s" % (
        calls, n_s, "" if n_s == n else " (1/%d of the workload's candidates)" % (n // n_s), cfg["h"], cfg["m"], cfg["E"], cfg["mode"], what,
        os.cpu_count(), cores, dt * calls)
    return dict(value=rollouts / dt, unit="rollouts/s", cores=cores, kind=kind, sample=sample), dt, n_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n_gpu, n_glob = split_candidates(cfg, args.gpus, args.scaling)
    cpu, dt, n_s = time_host_planner(cfg, n_glob, budget_s=1e9, min_calls=args.steps, max_calls=args.steps, warmup=args.warmup)
    line = dict(impl="reference", metric=METRIC, value=cpu["value"], unit="rollouts/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling=args.scaling, vs_baseline=None,
                dtype="f32", data="synthetic", config=config_dict(args.config, cfg, args.gpus, args.scaling, "numpy (host, inside the timed call)"),
                cpu_baseline=cpu, e2e=dict(value=cpu["value"], unit="rollouts/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# =================================================================================================== CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from learning_to_adapt_b200 import _native as N
    from learning_to_adapt_b200.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
    from learning_to_adapt_b200.dynamics.mlp_dynamics import MLPDynamicsModel
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv
    from learning_to_adapt_b200.parallel import CandidateShard
    from learning_to_adapt_b200.policies.mpc_controller import MPCController

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = CONFIGS[args.config]
    if distributed and (cfg["planner"] == "cem" or cfg.get("grbal")):
        raise SystemExit("config %s is a single-GPU workload" % args.config)
    O, prob = make_problem(cfg)
    env = SyntheticEnv(cfg["env"])
    n, n_glob = split_candidates(cfg, world, args.scaling)
    h, m, E = cfg["h"], cfg["m"], cfg["E"]
    A = prob["act_dim"]
    grbal = bool(cfg.get("grbal"))
    if grbal:
        model = MetaMLPDynamicsModel("dyn", env, hidden_sizes=cfg["hidden"], meta_batch_size=10, inner_learning_rate=1e-3, device=local_rank)
        model.set_params(prob["param_sets"][0])
    else:
        model = MLPDynamicsModel("dyn", env, hidden_sizes=cfg["hidden"], ensemble_size=E if cfg["mode"] == "ensemble" else 1, device=local_rank)
        for e, p in enumerate(prob["param_sets"]):
            model.set_params(p, member=e)
    model.set_normalization(prob["norm"])
    eng = model._engine
    shard = CandidateShard() if distributed else None
    low, high = eng._f32(prob["low"]), eng._f32(prob["high"])
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    obs_dev = eng._f32(prob["obs0"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    win = None
    if grbal:
        # a running path of M + 2 transitions per env in the device window, then adapt once so sets 1..m are live
        ctx = O.make_adapt_context(4, prob, m, 16)
        win = model.make_adapt_window(m, 16)
        for j in range(18):
            win.push(np.stack([c[min(j, 15)] for c in ctx[0]]), np.stack([c[min(j, 15)] for c in ctx[1]]))
        model.adapt_from_window(win, defer=False)
    set_mode, first_set, n_sets = model.planning_sets(m)
    iters = cfg.get("iters", 1)
    cem = cfg["planner"] == "cem"
    ha = h * A
    if cem:
        pool = [torch.randn((n, m, ha), device="cuda", generator=gen) for _ in range(iters)]
        num_elites = max(int(n * cfg["pct"]), 1)
        clip_low, clip_high = eng._f32(np.concatenate([prob["low"]] * h)), eng._f32(np.concatenate([prob["high"]] * h))
        mean = torch.zeros((m, ha), device="cuda", dtype=torch.float64)
        std = torch.ones((m, ha), device="cuda", dtype=torch.float64)
    else:
        pool = [torch.rand((h, n * m, A), device="cuda", generator=gen) * (high - low) + low for _ in range(4)]

    def rollout(actions, layout="thra", want_returns=False):
        return eng.rollout(obs_dev, actions, n, h, prob["reward_kind"], prob["dt"], set_mode=set_mode, first_set=first_set,
                           n_sets=n_sets, layout=layout, want_returns=want_returns)

    def plan_resident(i, exchange="peer"):
        """One planning call on device-resident inputs.  N>1: + the exchange of the ranks' winners, either the library's
        peer-memory kernel (what the e2e call uses) or one NCCL all-gather through torch.distributed."""
        if cem:
            for it in range(iters):
                samples, clipped = eng.cem_sample(pool[it], mean, std, clip_low, clip_high)
                res = rollout(samples, "nmha", True)
                eng.cem_refit(res["returns"], clipped, num_elites, cfg["alpha"], mean, std, compat=True)
            return res
        res = rollout(pool[i % len(pool)])
        if shard is not None:
            if exchange == "peer":
                eng.exchange_resident(res, out=xrec)
            else:
                shard.combine(res["best_ret"], res["best_idx"], res["best_act"], rank * n, engine=eng)
        return res

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- the public-API controller (its sharded plan also serves
    # the device-resident exchange below)
    def controller(sampler_name):
        kw = dict(n_candidates=n_glob if distributed else n, horizon=h, sampler=sampler_name, parallel=shard)
        if cem:
            kw.update(use_cem=True, num_cem_iters=iters, percent_elites=cfg["pct"], alpha=cfg["alpha"])
        c = MPCController("policy", env, model, **kw)
        if grbal:
            c.push_window = win
        return c

    obs_host = np.array(prob["obs0"])
    ctrl = controller("device")
    if distributed:
        ctrl.get_actions(obs_host)                               # creates the plan and attaches the peers' exchange buffers
        xrec = torch.empty((m, 2 + A), device="cuda", dtype=torch.float64)

    # ---------------------------------------------------------------- device-resident timing ("value")
    W = max(3, args.warmup)
    for i in range(W):
        if cem:
            mean.zero_(); std.fill_(1.0)
        plan_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count

    def timed_steps(k, exchange):
        """k steps enqueued back to back (the ranks then run in lockstep through the per-step exchange, as consecutive planning
        calls do); every step sits between its own pair of CUDA events, the L2 flush between the steps outside of them."""
        evs = []
        barrier()
        for i in range(k):
            flush.fill_(i & 0xFF)                               # evict weights + candidates from L2 (outside the events)
            if cem:
                mean.zero_(); std.fill_(1.0)                    # mean = 0, std = 1 at the start of a call (:79-80)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            plan_resident(i, exchange)
            e.record()
            evs.append((s, e))
        barrier()
        return [s.elapsed_time(e) for s, e in evs]

    wall0 = time.perf_counter()
    step_ms = timed_steps(args.steps, "peer")
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0                     # kernels of this library inside the timed region
    nccl_ms = None
    if distributed:                                             # the same step with the NCCL all-gather as the exchange
        nccl_ms = max_over_ranks(float(np.mean(timed_steps(3 + min(args.steps, 10), "nccl")[3:])))
    ms_step = max_over_ranks(float(np.mean(step_ms)))
    rollouts_per_call = world * n * m * iters if args.scaling == "weak" else n_glob * m * iters
    value = rollouts_per_call / (ms_step * 1e-3)

    # ---------------------------------------------------------------- dominant kernel alone (roofline)
    kern_ms = []
    k_actions = (eng.cem_sample(pool[0], torch.zeros_like(mean), torch.ones_like(std), clip_low, clip_high)[0] if cem else pool[0])
    for i in range(3 + min(args.steps, 20)):
        flush.fill_(i & 0xFF)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rollout(k_actions, "nmha" if cem else "thra", cem)
        e.record()
        torch.cuda.synchronize()
        if i >= 3:
            kern_ms.append(s.elapsed_time(e))
    ms_kernel = max_over_ranks(float(np.mean(kern_ms)))

    # ---------------------------------------------------------------- end-to-end through the public API
    def e2e_rate(ctrl, steps):
        def step():
            if grbal:                                           # the GrBAL env step: samplers/sampler.py:81-91
                model.switch_to_pre_adapt()
                model.adapt_from_window(win)
            return ctrl.get_actions(obs_host)
        for i in range(W):
            step()
        barrier()
        total = 0.0
        for i in range(steps):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()                            # (no per-step barrier: the previous call's exchange left the ranks aligned)
            t0 = time.perf_counter()
            step()                                              # H2D obs, [adapt], sample, K1, [exchange], D2H actions
            total += time.perf_counter() - t0
        return max_over_ranks(total / steps)

    e2e_s = e2e_rate(ctrl, args.steps)
    h2d, d2h = eng.last_plan_io_bytes()
    e2e_value = rollouts_per_call / e2e_s
    e2e_default = None
    if not distributed:
        np.random.seed(0)
        e2e_default = rollouts_per_call / e2e_rate(controller("numpy"), min(args.steps, 20))
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if world == 1 and not os.environ.get("L2A_BENCH_SKIP_CPU"):      # (development runs may skip the host leg)
        cpu, _, _ = time_host_planner(cfg, n, budget_s=12.0, min_calls=3, max_calls=40, warmup=1)

    peaks = load_peaks()
    F = flops_per_dyn_step(prob["obs_dim"], prob["act_dim"], cfg["hidden"])
    e_eff = E if cfg["mode"] == "ensemble" else 1
    flops_per_launch = n * m * h * e_eff * F
    achieved = flops_per_launch / (ms_kernel * 1e-3) / 1e12
    traffic = None
    kernel_name = "rollout_tc_kernel" if os.environ.get("L2A_TC_PAIR") == "0" else "rollout_tc2_kernel"
    tpath = os.path.join(REPO, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(kernel_name + ":" + args.config)
    line = dict(metric=METRIC, value=value, unit="rollouts/s", n_gpus=world, steps=args.steps, warmup=W,
                ms_per_step=ms_step, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="bf16x3-split (fp32 accumulate)",
                data="synthetic", config=config_dict(args.config, cfg, world, args.scaling,
                                                     "device (Philox, inside the e2e call; pre-materialised for `value`)"),
                dyn_steps_per_s=rollouts_per_call * h * e_eff / (ms_step * 1e-3),
                e2e=dict(value=e2e_value, unit="rollouts/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                         ms_per_call=e2e_s * 1e3, default_sampler_value=e2e_default,
                         note="value: sampler='device' (Philox); default_sampler_value: the package default, the reference's numpy "
                              "stream regenerated on the device (N=1 only)"),
                gpu_launches=int(launches),
                exchange=(dict(value_uses="peer-memory exchange kernel (l2a_plan_exchange_resident)", ms_per_step_peer=ms_step,
                               ms_per_step_nccl_allgather=nccl_ms) if distributed else None),
                roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"],
                              traffic=traffic, kernel=kernel_name, kernel_ms=ms_kernel,
                              flops_per_launch=flops_per_launch, peak_source=peaks["source"],
                              note="algorithmic FLOPs counted once; the kernel issues 3 bf16 MMA passes per product (split-bf16), so the "
                                   "attainable fraction of the bf16 peak is 1/3; traffic = ncu dram bytes per launch (profiles/)"),
                cpu_baseline=cpu, clocks=clocks, wall_s_timed_region=wall)
    emit(line)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
