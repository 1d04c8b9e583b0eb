"""``import learning_to_adapt_b200.dropin`` -- makes the reference's own import paths of the hot-path classes resolve to this
package, so ``run_scripts/run_grbal.py`` / ``run_rebal.py`` / ``run_mb_mpc.py`` need no edit at all:

    python -c "import learning_to_adapt_b200.dropin, runpy; runpy.run_path('run_scripts/run_grbal.py', run_name='__main__')"

Only the modules on the planning path are redirected (policies/mpc_controller.py, policies/rnn_mpc_controller.py,
dynamics/{mlp,meta_mlp,rnn}_dynamics.py, samplers/{sampler,vectorized_env_executor}.py); every other ``learning_to_adapt.*``
import (envs, trainers, logger, ...) still comes from the reference checkout.  ``uninstall()`` removes the hook."""
import importlib
import importlib.abc
import importlib.util
import sys

REDIRECTS = {
    "learning_to_adapt.policies.mpc_controller": "learning_to_adapt_b200.policies.mpc_controller",
    "learning_to_adapt.policies.rnn_mpc_controller": "learning_to_adapt_b200.policies.rnn_mpc_controller",
    "learning_to_adapt.dynamics.mlp_dynamics": "learning_to_adapt_b200.dynamics.mlp_dynamics",
    "learning_to_adapt.dynamics.meta_mlp_dynamics": "learning_to_adapt_b200.dynamics.meta_mlp_dynamics",
    "learning_to_adapt.dynamics.rnn_dynamics": "learning_to_adapt_b200.dynamics.rnn_dynamics",
    "learning_to_adapt.samplers.sampler": "learning_to_adapt_b200.samplers.sampler",
    "learning_to_adapt.samplers.vectorized_env_executor": "learning_to_adapt_b200.samplers.vectorized_env_executor",
}


class _Redirect(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in REDIRECTS:
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(REDIRECTS[spec.name])      # the SAME module object under both names

    def exec_module(self, module):
        pass


class _EmptyParents(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """LAST on sys.meta_path: when no reference checkout is importable (e.g. only a snapshot is being loaded), the parent
    packages of the redirected modules resolve to empty packages instead of failing."""
    PARENTS = sorted({name.rsplit(".", k)[0] for name in REDIRECTS for k in (1, 2)})

    def find_spec(self, fullname, path=None, target=None):
        if fullname in self.PARENTS:
            return importlib.util.spec_from_loader(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        module.__path__ = []


_HOOK = _Redirect()
_PARENTS_HOOK = _EmptyParents()


def _install_tensorflow_placeholders():
    """Snapshots written by the reference (logger.save_itr_params -> joblib, logger/logger.py:376-397) pickle the constructor
    arguments of the dynamics model, and the constructor default ``optimizer=tf.train.AdamOptimizer`` (mlp_dynamics.py:34)
    makes them name the class ``tensorflow.python.training.adam.AdamOptimizer``.  Where TensorFlow is not installed, a
    placeholder class under that path lets such a ``params.pkl`` unpickle; the B200 models ignore the argument (their
    optimiser is the TF-Adam update of dynamics/fit.py)."""
    import types
    try:
        import tensorflow  # noqa: F401
        return
    except Exception:
        pass
    names = ["tensorflow", "tensorflow.python", "tensorflow.python.training", "tensorflow.python.training.adam"]
    mods = {}
    for name in names:
        mods[name] = sys.modules.get(name) or types.ModuleType(name)
        mods[name].__dict__.setdefault("__l2a_placeholder__", True)
        sys.modules[name] = mods[name]

    class AdamOptimizer(object):
        """placeholder of tf.train.AdamOptimizer (only ever used as a pickled constructor default)"""

    AdamOptimizer.__module__ = "tensorflow.python.training.adam"
    AdamOptimizer.__qualname__ = "AdamOptimizer"
    mods["tensorflow.python.training.adam"].AdamOptimizer = AdamOptimizer
    train = types.ModuleType("tensorflow.train")
    train.AdamOptimizer = AdamOptimizer
    mods["tensorflow"].train = train
    mods["tensorflow"].python = mods["tensorflow.python"]
    mods["tensorflow.python"].training = mods["tensorflow.python.training"]
    mods["tensorflow.python.training"].adam = mods["tensorflow.python.training.adam"]


def install():
    if _HOOK not in sys.meta_path:
        sys.meta_path.insert(0, _HOOK)
    if _PARENTS_HOOK not in sys.meta_path:
        sys.meta_path.append(_PARENTS_HOOK)
    _install_tensorflow_placeholders()


def uninstall():
    if _HOOK in sys.meta_path:
        sys.meta_path.remove(_HOOK)
    if _PARENTS_HOOK in sys.meta_path:
        sys.meta_path.remove(_PARENTS_HOOK)
    for name in REDIRECTS:
        sys.modules.pop(name, None)


install()
