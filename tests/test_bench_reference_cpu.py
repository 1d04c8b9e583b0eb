"""bench.py --impl reference runs on the host alone (the reference algorithm's oracle port): one JSON line with the contract's
keys on the real stdout, nothing else there."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_json_line():
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rollouts/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0           # (a slow host times a fraction of the candidates per step; see `sample`)
    assert "N=2000" in d["config"]["workload"] and "H=20" in d["config"]["workload"] and "E=5" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=REPO, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_bench_configs_and_roofline_numerators_follow_the_survey():
    """bench.py's workloads are BASELINE.json's configs (SURVEY.md section 8(d) "Configs -> concrete shapes") and its roofline
    numerator is the survey's algorithmic FLOP count per dynamics step, counted once (not per MMA pass)."""
    sys.path.insert(0, REPO)
    import bench
    from oracle import mpc_oracle as O
    assert bench.flops_per_dyn_step(20, 6, (512, 512)) == 571392            # HC(512^2)
    assert bench.flops_per_dyn_step(20, 6, (512, 512, 512)) == 1095680      # HC(512^3)
    assert bench.flops_per_dyn_step(41, 8, (512, 512, 512)) == 1140736      # Ant(512^3)
    want = {   # name: (env, N, H, m, E, hidden layers, planner)
        "headline": ("half_cheetah", 2000, 20, 1, 5, 3, "rs"), "cfg1": ("half_cheetah", 500, 10, 1, 1, 2, "rs"),
        "cfg1p": ("half_cheetah", 2000, 20, 10, 1, 2, "rs"), "cfg2": ("half_cheetah", 1000, 15, 5, 5, 3, "rs"),
        "cfg3": ("ant", 2000, 20, 1, 5, 3, "rs"), "cfg4": ("half_cheetah", 5000, 30, 1, 1, 2, "cem"),
        "cfg5": ("ant", 4096, 25, 1, 5, 3, "rs")}
    assert sorted(bench.CONFIGS) == sorted(want)
    for name, (env, n, h, m, e, layers, planner) in want.items():
        c = bench.CONFIGS[name]
        assert (c["env"], c["n"], c["h"], c["m"], c["E"], len(c["hidden"]), c["planner"]) == (env, n, h, m, e, layers, planner), name
        assert all(w == 512 for w in c["hidden"])
    c4 = bench.CONFIGS["cfg4"]
    assert (c4["iters"], c4["pct"], c4["alpha"]) == (3, 0.1, 0.1) and int(c4["n"] * c4["pct"]) == 500
    # totals per call: headline 219 GF over 200 k dynamics steps; cfg5 on 8 GPUs 4.67 TF over 4.1 M steps
    head = bench.config_dict("headline", bench.CONFIGS["headline"], 1, "weak", "x")
    assert head["dyn_steps_per_call_per_gpu"] == 200000
    assert abs(200000 * bench.flops_per_dyn_step(20, 6, (512, 512, 512)) - 219.136e9) < 1e6
    c5 = bench.config_dict("cfg5", bench.CONFIGS["cfg5"], 8, "weak", "x")
    assert c5["global_candidates"] == 32768 and c5["n_candidates_per_gpu"] == 4096
    assert abs(8 * c5["dyn_steps_per_call_per_gpu"] * bench.flops_per_dyn_step(41, 8, (512, 512, 512)) / 1e12 - 4.67) < 0.01
    # weak / strong split of the candidates
    assert bench.split_candidates(bench.CONFIGS["headline"], 8, "weak") == (2000, 16000)
    assert bench.split_candidates(bench.CONFIGS["headline"], 8, "strong") == (250, 2000)
    # the synthetic problem of a config has the env's dimensions (HC D=20 A=6, Ant D=41 A=8)
    _, prob = bench.make_problem(bench.CONFIGS["cfg3"])
    assert (prob["obs_dim"], prob["act_dim"]) == (41, 8) and float(prob["high"][0]) == 150.0 and prob["dt"] == 0.02
    _, prob = bench.make_problem(bench.CONFIGS["headline"])
    assert (prob["obs_dim"], prob["act_dim"]) == (20, 6) and prob["dt"] == 0.01 and len(prob["param_sets"]) == 5
    assert prob["reward_kind"] == O.REWARD_HALF_CHEETAH
