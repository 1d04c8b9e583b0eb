"""Generate tests/golden/reference_snapshot_params.pkl by running the REFERENCE'S OWN pickling code (authoring container only).

What is executed from upstream, unmodified, imported from /root/reference:
  * ``learning_to_adapt.utils.serializable.Serializable.quick_init`` / ``__getstate__`` -- constructor-argument capture;
  * ``MetaMLPDynamicsModel.__getstate__`` (dynamics/meta_mlp_dynamics.py:434-439) and ``Layer.__getstate__``
    (dynamics/core/layers.py:103-108) -- called on instances created with ``__new__`` (their ``__init__`` builds TF1 graphs;
    TensorFlow 1.13.1 cannot be installed here), with ``get_param_values`` returning the OrderedDict of arrays that
    ``sess.run(self._params)`` returns;
  * the snapshot writer ``joblib.dump(params, 'params.pkl', compress=3)`` of logger/logger.py:376-397 (``Trainer`` snapshot dict
    ``dict(itr=..., dynamics_model=...)``, trainers/mb_trainer.py:118-122).
The pickle therefore names the upstream class path ``learning_to_adapt.dynamics.meta_mlp_dynamics.MetaMLPDynamicsModel`` and, via
the constructor default, ``tensorflow.python.training.adam.AdamOptimizer``: tests/test_gpu_api.py loads it through
``learning_to_adapt_b200.dropin`` (which resolves both) into the B200 model.

    python tests/golden/make_golden_snapshot.py
"""
import os
import sys
import types
from collections import OrderedDict

import joblib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from make_golden import _Anything  # noqa: E402

REF = "/root/reference"


def main():
    from oracle import mpc_oracle as O
    from learning_to_adapt_b200.envs.synthetic import SyntheticEnv

    # a tensorflow stand-in whose AdamOptimizer pickles under TF 1.13's class path
    tf = _Anything("tensorflow")
    sys.modules["tensorflow"] = tf
    adam_mod = types.ModuleType("tensorflow.python.training.adam")

    class AdamOptimizer(object):
        pass

    AdamOptimizer.__module__ = "tensorflow.python.training.adam"
    AdamOptimizer.__qualname__ = "AdamOptimizer"
    adam_mod.AdamOptimizer = AdamOptimizer
    for name in ("tensorflow.python", "tensorflow.python.training"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.python.training.adam"] = adam_mod
    tf.train.AdamOptimizer = AdamOptimizer
    for name in ("gym", "gym.spaces", "pyprind", "mpi4py"):
        sys.modules.setdefault(name, _Anything(name))
    sys.path.insert(0, REF)
    from learning_to_adapt.dynamics import meta_mlp_dynamics as ref_mod
    from learning_to_adapt.dynamics.core.layers import MLP as RefMLP
    from learning_to_adapt.utils.serializable import Serializable as RefSerializable

    prob = O.make_problem("half_cheetah", hidden_sizes=(32, 32), n_sets=1, m=1, seed=21)
    env = SyntheticEnv("half_cheetah")
    cls = ref_mod.MetaMLPDynamicsModel
    model = cls.__new__(cls)
    # the first statement of the upstream __init__ (meta_mlp_dynamics.py:41), with the call's locals
    locals_ = dict(self=model, name="dyn", env=env, hidden_sizes=(32, 32), meta_batch_size=7, hidden_nonlinearity="relu",
                   output_nonlinearity=None, batch_size=16, learning_rate=0.001, inner_learning_rate=0.05, normalize_input=True,
                   optimizer=AdamOptimizer, valid_split_ratio=0.2, rolling_average_persitency=0.99)
    RefSerializable.quick_init(model, locals_)
    model.normalization = prob["norm"]
    net = RefMLP.__new__(RefMLP)
    net.get_param_values = lambda: OrderedDict((k, v.copy()) for k, v in prob["param_sets"][0].items())   # = sess.run(self._params)
    model._networks = [net]
    params = dict(itr=3, dynamics_model=model)                                   # Trainer.get_itr_snapshot (mb_trainer.py:118-122)
    out = os.path.join(HERE, "reference_snapshot_params.pkl")
    joblib.dump(params, out, compress=3)                                         # logger.save_itr_params, snapshot_mode='last'
    np.savez(os.path.join(HERE, "reference_snapshot_expect.npz"), **{k.replace("/", "."): v for k, v in prob["param_sets"][0].items()},
             obs_mean=prob["norm"]["obs"][0], obs_std=prob["norm"]["obs"][1])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
