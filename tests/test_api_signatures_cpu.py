"""Drop-in boundary: the mirror classes keep the reference's constructor / method signatures -- same parameter
names in the same order with the same defaults (extra keyword parameters may only follow them).  The fixture is extracted from
the upstream source text by tests/golden/make_api_signatures.py."""
import ast
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = json.load(open(os.path.join(HERE, "golden", "reference_api_signatures.json")))

MIRRORS = {
    "MPCController": "learning_to_adapt_b200.policies.mpc_controller",
    "RNNMPCController": "learning_to_adapt_b200.policies.rnn_mpc_controller",
    "MLPDynamicsModel": "learning_to_adapt_b200.dynamics.mlp_dynamics",
    "MetaMLPDynamicsModel": "learning_to_adapt_b200.dynamics.meta_mlp_dynamics",
    "RNNDynamicsModel": "learning_to_adapt_b200.dynamics.rnn_dynamics",
    "Sampler": "learning_to_adapt_b200.samplers.sampler",
    "IterativeEnvExecutor": "learning_to_adapt_b200.samplers.vectorized_env_executor",
}
# reference defaults that are TensorFlow objects map to the string / None forms the run scripts actually pass
# (run_scripts/run_grbal.py:33-46: hidden_nonlinearity='relu', output_nonlinearity=None; the optimizer is never overridden)
TF_DEFAULTS = {"tf.nn.relu": ("relu", "'relu'"), "tf.nn.tanh": ("tanh", "'tanh'"), "tf.train.AdamOptimizer": None}


def _same_default(ref_src, got):
    if ref_src in TF_DEFAULTS:
        allowed = TF_DEFAULTS[ref_src]
        return True if allowed is None else (got in allowed or got is None or callable(got))
    want = ast.literal_eval(ref_src)
    return got == want and type(got) is type(want)


@pytest.mark.parametrize("qualname", sorted(FIXTURE))
def test_mirror_keeps_the_reference_signature(qualname):
    cls_name, method = qualname.split(".")
    cls = getattr(importlib.import_module(MIRRORS[cls_name]), cls_name)
    assert hasattr(cls, method), "%s is missing (reference %s)" % (qualname, FIXTURE[qualname]["file"])
    got = list(inspect.signature(getattr(cls, method)).parameters.values())
    ref = FIXTURE[qualname]["params"]
    assert len(got) >= len(ref), "%s: fewer parameters than the reference (%s)" % (qualname, FIXTURE[qualname]["file"])
    for i, (name, default_src) in enumerate(ref):
        p = got[i]
        assert p.name == name, "%s: parameter %d is %r, reference has %r (%s)" % (qualname, i, p.name, name, FIXTURE[qualname]["file"])
        if default_src is None:
            assert p.default is inspect.Parameter.empty, "%s: %s must stay positional-required" % (qualname, name)
        else:
            assert p.default is not inspect.Parameter.empty, "%s: %s lost its default" % (qualname, name)
            assert _same_default(default_src, p.default), "%s: default of %s is %r, reference %s" % (qualname, name, p.default, default_src)
    for p in got[len(ref):]:
        assert p.default is not inspect.Parameter.empty or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL), \
            "%s: extra parameter %s must be optional" % (qualname, p.name)
