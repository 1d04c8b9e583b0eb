"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol include/l2a_b200.h
declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(debug=False):
    """l2a_* functions include/l2a_b200.h declares; debug=True: the ones inside #ifdef L2A_DEBUG_KERNELS only."""
    text = open(os.path.join(REPO, "include", "l2a_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    m = re.search(r"#ifdef L2A_DEBUG_KERNELS(.*?)#endif", text, flags=re.S)
    assert m, "the diagnostics must sit behind L2A_DEBUG_KERNELS"
    part = m.group(1) if debug else text.replace(m.group(0), "")
    return sorted(set(re.findall(r"\b(l2a_[a-z0-9_]+)\s*\(", part)))


def test_library_builds_and_exports_every_declared_symbol():
    from learning_to_adapt_b200 import _native
    from learning_to_adapt_b200.build import DEBUG_LIB_PATH, LIB_PATH, build
    build()
    build(debug=True)
    assert os.path.exists(LIB_PATH) and os.path.exists(DEBUG_LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libl2a_b200.so does not export %s" % name
    assert sorted(_native.EXPORTS) == declared
    assert _native.load().l2a_version() >= 100
    # the diagnostics exist only in the debug build
    debug_syms = _declared_symbols(debug=True)
    assert sorted(_native.DEBUG_EXPORTS) == debug_syms and len(debug_syms) == 5
    dbg = ctypes.CDLL(DEBUG_LIB_PATH)
    for name in debug_syms:
        assert hasattr(dbg, name), "libl2a_b200_debug.so does not export %s" % name
        assert not hasattr(lib, name), "the product library must not carry %s" % name
    for name in declared:
        assert hasattr(dbg, name)


def test_struct_layouts_match_header():
    from learning_to_adapt_b200 import _native
    assert ctypes.sizeof(_native.MlpDesc) == 4 * (3 + 7 + 1)
    assert ctypes.sizeof(_native.RolloutParams) == 8 * 4 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(_native.PlanOpts) == 4 * 4 + 8 + 8 + 4 * 4 + 8
    assert ctypes.sizeof(_native.PlanIO) == 9 * 8 + 8


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


def _plan(obs_dim, act_dim, hidden):
    from learning_to_adapt_b200 import _native
    lib = _native.load()
    d = _native.MlpDesc()
    d.obs_dim, d.act_dim, d.n_hidden, d.n_sets = obs_dim, act_dim, len(hidden), 1
    for i, h in enumerate(hidden):
        d.hidden[i] = h
    out = (ctypes.c_int32 * 8)()
    status = lib.l2a_tc_plan_query(ctypes.byref(d), out)
    return status, list(out)


def test_tensor_core_tiling_of_the_baseline_shapes():
    """Host logic of the weight-tile plan (no GPU): ring stages per weight set, output-layer tile geometry, and which shapes
    have a tcgen05 variant at all (the others run on the fp32 SIMT kernel -- or raise when tcgen05 is forced)."""
    # HalfCheetah 26-512-512-512-20: layer 0 (input 30 <= 32 wide: two M-blocks per tile pair) 2 pairs + 2 x 32 pairs;
    # output N = 32, 4 K chunks per 32 KB stage, 2 stages
    st, p = _plan(20, 6, (512, 512, 512))
    assert st == 0 and p[:6] == [1, 66, 32, 4, 2, 68] and p[6] == 68 * 32768 and p[7] == 0
    # Ant 49-512-512-512-41: layer-0 input 56 wide -> unpacked (4 pairs); output N = 48 -> 2 chunks per stage, 4 stages
    st, p = _plan(41, 8, (512, 512, 512))
    assert st == 0 and p[:6] == [1, 68, 48, 2, 4, 72]
    # BASELINE cfg1: two hidden layers
    st, p = _plan(20, 6, (512, 512))
    assert st == 0 and p[:6] == [1, 34, 32, 4, 2, 36]
    # single 128-wide hidden layer (arm_7dof test shape): one pair + one output stage
    st, p = _plan(17, 7, (128,))
    assert st == 0 and p[:6] == [1, 1, 32, 4, 1, 2]
    # no tensor-core variant: width not a multiple of 128, too wide, obs / act beyond the register-resident limits
    for shape in [(20, 6, (100, 128)), (20, 6, (640,)), (49, 6, (128,)), (20, 17, (128,)), (48, 17, (128,))]:
        st, p = _plan(*shape)
        assert st == 0 and p[0] == 0, shape
    st, _ = _plan(20, 6, ())
    assert st == -1


def _plan2(obs_dim, act_dim, hidden, n, m, members, sms=148):
    from learning_to_adapt_b200 import _native
    lib = _native.load()
    d = _native.MlpDesc()
    d.obs_dim, d.act_dim, d.n_hidden, d.n_sets = obs_dim, act_dim, len(hidden), max(1, members)
    for i, h in enumerate(hidden):
        d.hidden[i] = h
    out = (ctypes.c_int32 * 12)()
    status = lib.l2a_tc2_plan_query(ctypes.byref(d), n, m, members, sms, out)
    return status, list(out)


def test_cta_pair_plan_and_tile_choice_of_the_baseline_configs():
    """Host logic of the CTA-pair kernel (no GPU): ring stages per weight set (64 KB = one 32 KB half per CTA) and the tile choice --
    fewest waves of WHOLE tiles (the members of a tile only progress together), then the smaller tile -- on a 148-SM device."""
    # headline: packed layer 0 (1 stage) + 2 x 16 + 2 output stages (5 K chunks of 6 KB per stage); 14 tiles of 144 x 5 members = 140 CTAs
    st, p = _plan2(20, 6, (512, 512, 512), 2000, 1, 5)
    assert st == 0 and p[:7] == [1, 33, 1, 32, 5, 2, 35] and p[7] == 35 * 65536 and p[9:] == [72, 14, 140]
    # cfg3 (Ant): unpacked layer 0 (2 stages), output N = 48 -> 9 KB per K chunk -> 3 per stage -> 3 stages
    st, p = _plan2(41, 8, (512, 512, 512), 2000, 1, 5)
    assert st == 0 and p[:7] == [1, 34, 0, 48, 3, 3, 37] and p[9:] == [72, 14, 140]
    # cfg5's per-GPU share: 29 tiles of 144 would need three waves of 14 whole tiles, 26 tiles of 160 need two
    st, p = _plan2(41, 8, (512, 512, 512), 4096, 1, 5)
    assert st == 0 and p[9:] == [80, 26, 260]
    # cfg1 / cfg2 / cfg4: no ensemble -> 74 pair slots; the smallest tile that still fits one wave
    assert _plan2(20, 6, (512, 512), 500, 1, 1)[1][9:] == [32, 8, 16]
    assert _plan2(20, 6, (512, 512, 512), 1000, 5, 1)[1][9:] == [48, 11, 110]
    assert _plan2(20, 6, (512, 512), 5000, 1, 1)[1][9:] == [48, 53, 106]
    assert _plan2(20, 6, (512, 512), 2000, 10, 1)[1][9:] == [72, 14, 280]
    # a 256-wide layer is one M-block; 128-wide layers, > 8 members: no pair variant (the single-CTA / SIMT kernels serve them)
    st, p = _plan2(20, 6, (256, 256), 300, 1, 1)
    assert st == 0 and p[:7] == [1, 1 + 4, 0, 32, 5, 1, 6]
    assert _plan2(20, 6, (128, 128), 300, 1, 1)[1][0] == 0
    assert _plan2(20, 6, (512, 512), 300, 1, 9)[1][0] == 0
    assert _plan2(20, 6, (), 300, 1, 1)[0] == -1


def test_product_library_carries_the_blackwell_instructions():
    """The built sm_100a library, disassembled: every tensor-core rollout kernel issues tcgen05.mma (UTCHMMA) with its commit
    (UTCBAR) and tcgen05.ld (LDTM); the CTA-pair kernel fetches its weight tiles through tensor-map TMA (UTMALDG), the single-CTA
    and the recurrent kernel through bulk TMA copies (UBLKCP)."""
    import shutil
    import subprocess
    from learning_to_adapt_b200.build import LIB_PATH, build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build()
    sass = subprocess.run([cuobjdump, "-sass", LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "arch = sm_100a" in sass
    census = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            census[name] = {}
            continue
        if name is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            census[name][m.group(1)] = census[name].get(m.group(1), 0) + 1
    def kernels(tag):
        found = {k: v for k, v in census.items() if tag in k}
        assert found, "no %s kernel in the library" % tag
        return found
    for tag, tma in (("rollout_tc2_kernel", "UTMALDG"), ("rollout_tc_kernel", "UBLKCP"), ("rollout_rnn_tc_kernel", "UBLKCP")):
        for k, ops in kernels(tag).items():
            assert ops.get("UTCHMMA", 0) >= 30, (k, ops.get("UTCHMMA"))
            assert ops.get("UTCBAR", 0) >= 1 and ops.get("LDTM", 0) >= 1 and ops.get(tma, 0) >= 1, k
            assert ops.get("SYNCS", 0) >= 10, k                      # mbarrier pipeline
    assert len(kernels("rollout_tc2_kernel")) >= 8                    # NC in {32, 48, 72, 80} x two state-register instances
