"""The import hook of learning_to_adapt_b200.dropin: inside a reference checkout the run scripts' own import lines
(run_scripts/run_grbal.py:1-12) pick up the B200 classes for the planning path and the reference's modules for the rest."""
import importlib
import sys
import textwrap


def test_reference_import_paths_resolve_to_the_b200_classes(tmp_path, monkeypatch):
    # a stand-in "reference checkout": real package tree, with a marker module that must NOT be redirected
    root = tmp_path / "checkout"
    for pkg in ("learning_to_adapt", "learning_to_adapt/policies", "learning_to_adapt/dynamics", "learning_to_adapt/samplers",
                "learning_to_adapt/envs"):
        (root / pkg).mkdir(parents=True)
        (root / pkg / "__init__.py").write_text("")
    (root / "learning_to_adapt/envs/marker_env.py").write_text("WHO = 'reference'\n")
    (root / "learning_to_adapt/policies/mpc_controller.py").write_text("raise ImportError('the TF1 planner must not be imported')\n")
    monkeypatch.syspath_prepend(str(root))
    for name in [m for m in sys.modules if m == "learning_to_adapt" or m.startswith("learning_to_adapt.")]:
        monkeypatch.delitem(sys.modules, name)
    import learning_to_adapt_b200.dropin as dropin
    dropin.install()
    try:
        ns = {}
        exec(textwrap.dedent("""
            from learning_to_adapt.dynamics.meta_mlp_dynamics import MetaMLPDynamicsModel
            from learning_to_adapt.dynamics.mlp_dynamics import MLPDynamicsModel
            from learning_to_adapt.dynamics.rnn_dynamics import RNNDynamicsModel
            from learning_to_adapt.policies.mpc_controller import MPCController
            from learning_to_adapt.policies.rnn_mpc_controller import RNNMPCController
            from learning_to_adapt.samplers.sampler import Sampler
            from learning_to_adapt.envs.marker_env import WHO
        """), ns)
        for cls in ("MetaMLPDynamicsModel", "MLPDynamicsModel", "RNNDynamicsModel", "MPCController", "RNNMPCController", "Sampler"):
            assert ns[cls].__module__.startswith("learning_to_adapt_b200."), cls
        assert ns["WHO"] == "reference"
        ours = importlib.import_module("learning_to_adapt_b200.policies.mpc_controller")
        assert importlib.import_module("learning_to_adapt.policies.mpc_controller") is ours
    finally:
        dropin.uninstall()
        for name in [m for m in sys.modules if m == "learning_to_adapt" or m.startswith("learning_to_adapt.")]:
            sys.modules.pop(name, None)
