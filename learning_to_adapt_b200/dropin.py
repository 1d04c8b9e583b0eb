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


_HOOK = _Redirect()


def install():
    if _HOOK not in sys.meta_path:
        sys.meta_path.insert(0, _HOOK)


def uninstall():
    if _HOOK in sys.meta_path:
        sys.meta_path.remove(_HOOK)
    for name in REDIRECTS:
        sys.modules.pop(name, None)


install()
