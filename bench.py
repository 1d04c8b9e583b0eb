#!/usr/bin/env python
"""bench.py -- candidate-rollouts/s of the MPC planning hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

A "step" is ONE planning call (one pass of the hot path over one batch of candidates): sample-free K1 rollout of
N candidates x H steps through the E-member dynamics ensemble + reward + argmax (+ the all-gather of the per-rank best
triple when N_gpus > 1).  Workload = BASELINE.md "headline": HalfCheetah, N=2000 per GPU, H=20, ensemble E=5,
MLP 26->512->512->512->20.  Weak scaling: every rank plans its own 2000 candidates, global N = 2000 x gpus.

value  : whole-job candidate-rollouts/s with the candidate tensor already resident in HBM, CUDA-event timed per step,
         L2 flushed (256 MB write) between steps, max over ranks.
e2e    : the same metric through the public API call MPCController.get_actions(obs ndarray) -> ndarray = ONE host-buffer
         C-ABI call (l2a_plan_run): host->device copy of the observations, Philox candidate sampling on the device, K1,
         device->host copy of the chosen actions -- all inside the timed region (sampler="device"; at N>1 GPUs the
         candidate shard adds the NCCL all-gather).
roofline: dominant kernel rollout_tc_kernel, tensor-pipe bound; achieved = algorithmic FLOPs per launch (N*H*E*F, counted once
         although split-bf16 issues 3 MMA passes) / mean launch duration.
cpu_baseline: oracle port (numpy planner + float32 BLAS MLP) timed on the host cores on a bounded sample of the same workload.
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

WORKLOAD = dict(env="half_cheetah", hidden=(512, 512, 512), n_candidates=2000, horizon=20, ensemble=5, n_envs=1)
METRIC = "candidate-rollouts/sec (N x H dynamics steps) HalfCheetah MPC"


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


def pick_blas_threads(O, prob, horizon):
    """The numpy/OpenBLAS port is not fastest with every host thread (a 128-core box oversubscribes 2000x512 GEMMs):
    time one small planning call per candidate thread count and keep the best, so the CPU baseline is the port at its best."""
    from threadpoolctl import threadpool_limits
    cores = os.cpu_count() or 1
    cands = sorted(set([c for c in (4, 8, 16, 32, 64, cores) if c <= cores]))
    best, best_t = cores, float("inf")
    acts = O.sample_rs_actions(0, prob["low"], prob["high"], 6, 2000)
    for c in cands:
        with threadpool_limits(limits=c):
            O.rs_plan(prob["obs0"], acts, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "ensemble")
            dt = float("inf")
            for _ in range(3):
                t0 = time.perf_counter()
                O.rs_plan(prob["obs0"], acts, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "ensemble")
                dt = min(dt, time.perf_counter() - t0)
        if dt < best_t * 0.97:
            best, best_t = c, dt
    return best


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


def make_problem():
    from oracle import mpc_oracle as O
    w = WORKLOAD
    return O, O.make_problem(w["env"], hidden_sizes=w["hidden"], n_sets=w["ensemble"], m=w["n_envs"], seed=0)


# =================================================================================================== reference arm
def run_reference(args):
    """The reference algorithm on the host: numpy planner loop (policies/mpc_controller.py:108-129 restated in oracle/)
    over a float32 BLAS dense stack, all host threads.  TF1 itself cannot be installed here (no py3.12 wheels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, prob = make_problem()
    w = WORKLOAD
    n = w["n_candidates"] * max(1, args.gpus)        # same whole-job workload as the CUDA arm at this N_gpus
    from threadpoolctl import threadpool_limits
    cores = pick_blas_threads(O, prob, w["horizon"])
    limiter = threadpool_limits(limits=cores)

    def plan(seed):
        actions = O.sample_rs_actions(seed, prob["low"], prob["high"], w["horizon"], n * w["n_envs"])   # the reference samples inside the call
        return O.rs_plan(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "ensemble")

    for i in range(args.warmup):
        plan(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        plan(100 + i)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = n * w["n_envs"] / dt
    line = dict(impl="reference", metric=METRIC, value=value, unit="rollouts/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=config_dict(args.gpus, sampler="numpy (host, inside the timed call)"),
                cpu_baseline=dict(value=value, unit="rollouts/s", cores=cores, kind="port",
                                  sample="%d full planning calls of the workload (numpy planner + fp32 BLAS MLP, best of {4..%d} BLAS threads = %d)" % (args.steps, os.cpu_count(), cores)),
                e2e=dict(value=value, unit="rollouts/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def config_dict(gpus, sampler):
    w = WORKLOAD
    return dict(workload="BASELINE.md headline: HalfCheetah random-shooting MPC, N=%d candidates/GPU x H=%d, ensemble E=%d (mean of deltas), "
                         "MLP 26-512-512-512-20, m=%d env" % (w["n_candidates"], w["horizon"], w["ensemble"], w["n_envs"]),
                n_candidates_per_gpu=w["n_candidates"], global_candidates=w["n_candidates"] * max(1, gpus), horizon=w["horizon"],
                ensemble=w["ensemble"], dyn_steps_per_call_per_gpu=w["n_candidates"] * w["horizon"] * w["ensemble"],
                parallelism="candidate-shard x%d + 1 all-gather of (ret, idx, act)" % max(1, gpus), sampler=sampler,
                l2="flushed between timed steps (256 MB write); inputs (0.96 MB candidates + 11 MB weights) are smaller than L2")


# =================================================================================================== CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from learning_to_adapt_b200 import _native as N
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
    O, prob = make_problem()
    w = WORKLOAD
    env = SyntheticEnv(w["env"])
    model = MLPDynamicsModel("dyn", env, hidden_sizes=w["hidden"], ensemble_size=w["ensemble"], device=local_rank)
    for e, p in enumerate(prob["param_sets"]):
        model.set_params(p, member=e)
    model.set_normalization(prob["norm"])
    eng = model._engine
    shard = CandidateShard() if distributed else None
    n, h, m, E = w["n_candidates"], w["horizon"], w["n_envs"], w["ensemble"]
    A = prob["act_dim"]
    low, high = eng._f32(prob["low"]), eng._f32(prob["high"])
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    pool = [torch.rand((h, n * m, A), device="cuda", generator=gen) * (high - low) + low for _ in range(4)]
    obs_dev = eng._f32(prob["obs0"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    set_mode, first_set, n_sets = model.planning_sets(m)

    def plan_resident(i):
        res = eng.rollout(obs_dev, pool[i % len(pool)], n, h, prob["reward_kind"], prob["dt"], set_mode=set_mode,
                          first_set=first_set, n_sets=n_sets, want_returns=False)
        if shard is not None:
            return shard.combine(res["best_ret"], res["best_idx"], res["best_act"], rank * n, engine=eng)
        return res["best_ret"], res["best_idx"], res["best_act"]

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

    # ---------------------------------------------------------------- device-resident timing ("value", roofline)
    for i in range(max(3, args.warmup)):
        plan_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    step_ms, kern_ms = [], []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                   # evict weights + candidates from L2 (outside the events)
        s, k, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s.record()
        res = eng.rollout(obs_dev, pool[i % len(pool)], n, h, prob["reward_kind"], prob["dt"], set_mode=set_mode,
                          first_set=first_set, n_sets=n_sets, want_returns=False)
        k.record()
        if shard is not None:
            shard.combine(res["best_ret"], res["best_idx"], res["best_act"], rank * n, engine=eng)
        e.record()
        torch.cuda.synchronize()
        kern_ms.append(s.elapsed_time(k))
        step_ms.append(s.elapsed_time(e))
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    ms_step = max_over_ranks(float(np.mean(step_ms)))
    ms_kernel = max_over_ranks(float(np.mean(kern_ms)))
    value = world * n * m / (ms_step * 1e-3)

    # ---------------------------------------------------------------- end-to-end through the public API
    ctrl = MPCController("policy", env, model, n_candidates=n * world if distributed else n, horizon=h, sampler="device",
                         parallel=shard)
    obs_host = np.array(prob["obs0"])
    for i in range(max(3, args.warmup)):
        ctrl.get_actions(obs_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        acts, _ = ctrl.get_actions(obs_host)                    # H2D obs, sample, K1 (+ all-gather), D2H actions
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = world * n * m / e2e_s
    ctrl_np = MPCController("policy", env, model, n_candidates=n, horizon=h, sampler="numpy") if not distributed else None
    if os.environ.get("L2A_BENCH_SKIP_CPU"):
        ctrl_np = None
    e2e_numpy = None
    if ctrl_np is not None:
        np.random.seed(0)
        ctrl_np.get_actions(obs_host)
        t0 = time.perf_counter()
        for i in range(min(args.steps, 5)):
            ctrl_np.get_actions(obs_host)
        e2e_numpy = n * m / ((time.perf_counter() - t0) / min(args.steps, 5))
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if world == 1 and not os.environ.get("L2A_BENCH_SKIP_CPU"):      # (development runs may skip the 15 s host leg)
        from threadpoolctl import threadpool_limits
        cores = pick_blas_threads(O, prob, h)
        calls = 0
        with threadpool_limits(limits=cores):
            t0 = time.perf_counter()
            while calls < 3 or (time.perf_counter() - t0 < 12.0 and calls < 40):
                actions = O.sample_rs_actions(calls, prob["low"], prob["high"], h, n * m)
                O.rs_plan(prob["obs0"], actions, prob["param_sets"], prob["norm"], prob["reward_kind"], prob["dt"], 1.0, "ensemble")
                calls += 1
            dt = (time.perf_counter() - t0) / calls
        cpu = dict(value=n * m / dt, unit="rollouts/s", cores=cores, kind="port",
                   sample="%d full planning calls (N=%d,H=%d,E=%d) of the oracle port: numpy planner + fp32 BLAS MLP, best of {4..%d} BLAS "
                          "threads = %d, %.1f s" % (calls, n, h, E, os.cpu_count(), cores, dt * calls))

    peaks = load_peaks()
    F = flops_per_dyn_step(prob["obs_dim"], prob["act_dim"], w["hidden"])
    flops_per_launch = n * m * h * E * F
    achieved = flops_per_launch / (ms_kernel * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("rollout_tc_kernel")
    line = dict(metric=METRIC, value=value, unit="rollouts/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16x3-split (fp32 accumulate)",
                data="synthetic", config=config_dict(world, "device (Philox, inside the e2e call; pre-materialised for `value`)"),
                dyn_steps_per_s=world * n * m * h * E / (ms_step * 1e-3),
                e2e=dict(value=e2e_value, unit="rollouts/s", h2d_bytes_per_step=int(m * prob["obs_dim"] * 4),
                         d2h_bytes_per_step=int(m * A * 4), ms_per_call=e2e_s * 1e3,
                         numpy_sampler_parity_mode_value=e2e_numpy),
                gpu_launches=int(launches),
                roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"],
                              traffic=traffic, kernel="rollout_tc_kernel<80>", kernel_ms=ms_kernel,
                              flops_per_launch=flops_per_launch, peak_source=peaks["source"],
                              note="algorithmic FLOPs counted once; the kernel issues 3 bf16 MMA passes per product (split-bf16), so the "
                                   "attainable fraction of the bf16 peak is 1/3"),
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
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
