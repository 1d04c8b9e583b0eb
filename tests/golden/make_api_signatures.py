"""Extracts the constructor / method signatures of the reference classes on the hot path from the upstream SOURCE TEXT (ast only:
TensorFlow and MuJoCo are not importable here) into tests/golden/reference_api_signatures.json.  Run in the authoring container
(needs /root/reference); the fixture is what travels.

    python tests/golden/make_api_signatures.py
"""
import ast
import json
import os

REF = "/root/reference/learning_to_adapt"
TARGETS = {
    "policies/mpc_controller.py": {"MPCController": ["__init__", "get_action", "get_actions", "get_random_action", "get_cem_action",
                                                     "get_rs_action", "reset"]},
    "policies/rnn_mpc_controller.py": {"RNNMPCController": ["__init__", "get_action", "get_actions", "get_random_action",
                                                            "get_cem_action", "get_rs_action", "reset"]},
    "dynamics/mlp_dynamics.py": {"MLPDynamicsModel": ["__init__", "fit", "predict", "compute_normalization"]},
    "dynamics/meta_mlp_dynamics.py": {"MetaMLPDynamicsModel": ["__init__", "fit", "predict", "adapt", "switch_to_pre_adapt",
                                                               "compute_normalization"]},
    "dynamics/rnn_dynamics.py": {"RNNDynamicsModel": ["__init__", "fit", "predict", "compute_normalization", "get_initial_hidden"]},
    "samplers/sampler.py": {"Sampler": ["__init__", "obtain_samples"]},
    "samplers/vectorized_env_executor.py": {"IterativeEnvExecutor": ["__init__", "step", "reset"]},
}


def signature(fn):
    """[(name, default source text or None)], positional parameters only (the reference uses neither *args nor kw-only)."""
    args = fn.args
    names = [a.arg for a in args.args]
    defaults = [None] * (len(names) - len(args.defaults)) + [ast.unparse(d) for d in args.defaults]
    return [[n, d] for n, d in zip(names, defaults)]


def main():
    out = {}
    for rel, classes in TARGETS.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                methods = {f.name: f for f in node.body if isinstance(f, ast.FunctionDef)}
                for m in classes[node.name]:
                    if m in methods:
                        out["%s.%s" % (node.name, m)] = dict(file="%s:%d" % (rel, methods[m].lineno), params=signature(methods[m]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api_signatures.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote %s (%d signatures)" % (path, len(out)))


if __name__ == "__main__":
    main()
