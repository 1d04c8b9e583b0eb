"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol include/l2a_b200.h
declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "l2a_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(l2a_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from learning_to_adapt_b200 import _native
    from learning_to_adapt_b200.build import LIB_PATH, build
    build()
    assert os.path.exists(LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libl2a_b200.so does not export %s" % name
    assert sorted(_native.EXPORTS) == declared
    assert _native.load().l2a_version() >= 100


def test_struct_layouts_match_header():
    from learning_to_adapt_b200 import _native
    assert ctypes.sizeof(_native.MlpDesc) == 4 * (3 + 7 + 1)
    assert ctypes.sizeof(_native.RolloutParams) == 8 * 4 + 2 * 8 + 2 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    from learning_to_adapt_b200 import _native
    from learning_to_adapt_b200.engine import PlanningEngine
    lib = _native.load()
    h = ctypes.c_void_p()
    status = lib.l2a_ctx_create(0, ctypes.byref(h))
    assert status == -4                                  # L2A_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.l2a_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PlanningEngine(20, 6, (32, 32))


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "learning_to_adapt_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, f)
