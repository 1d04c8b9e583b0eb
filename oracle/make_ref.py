"""Recipe for oracle/_ref/: the reference's OWN planner, unmodified, next to the oracle (test / baseline infrastructure only).

    python oracle/make_ref.py          (authoring container only: needs /root/reference)

Copies the three pure-Python files the verbatim `MPCController` consists of --
    learning_to_adapt/policies/mpc_controller.py, learning_to_adapt/policies/base.py, learning_to_adapt/utils/serializable.py
-- byte for byte into oracle/_ref/learning_to_adapt/... and writes EMPTY package __init__ files beside them (the upstream
__init__ files import TensorFlow, which cannot be installed here).  oracle/_ref/ is git-ignored (no reference source enters
the history) but travels with the gpurun snapshot, so `bench.py --impl reference` can time the reference's planner loop on the
GPU box's host cores.  The dynamics model behind it stays the oracle port: the upstream model classes build TF1 graphs.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["learning_to_adapt/policies/mpc_controller.py", "learning_to_adapt/policies/base.py",
         "learning_to_adapt/utils/serializable.py"]


def make(verbose=False):
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for pkg in ("learning_to_adapt", "learning_to_adapt/policies", "learning_to_adapt/utils"):
        open(os.path.join(OUT, pkg, "__init__.py"), "w").close()
    if verbose:
        print("oracle/_ref: %d files from %s" % (len(FILES), REF))
    return True


def import_reference_controller():
    """The verbatim MPCController class, or None when oracle/_ref has not been made."""
    if not os.path.exists(os.path.join(OUT, FILES[0])):
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from learning_to_adapt.policies.mpc_controller import MPCController
    return MPCController


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
