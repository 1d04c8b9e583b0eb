"""Builds libl2a_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libl2a_b200.so")
DEBUG_LIB_PATH = os.path.join(LIB_DIR, "libl2a_b200_debug.so")     # + diagnostics / microbenchmarks (-DL2A_DEBUG_KERNELS)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(os.path.dirname(PKG), "include", "l2a_b200.h"))
    return out


def is_stale(path=LIB_PATH):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_variant(name, defines, verbose=False):
    """Development: a second build with extra -D flags at lib/variants/lib<name>.so (A/B runs on one box through the
    L2A_B200_LIB environment variable).  Not used by the product."""
    out_dir = os.path.join(LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    return _compile(os.path.join(out_dir, "lib%s.so" % name), list(defines), verbose)


def _compile(out, defines, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; %s must be prebuilt (it travels with the repo snapshot)" % os.path.basename(out))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "api.cu")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return out


def build(force=False, verbose=False, debug=False):
    """Compile csrc/api.cu (which includes every kernel) into lib/libl2a_b200.so -- or, with debug=True, into
    lib/libl2a_b200_debug.so, the same library plus the l2a_debug_* diagnostics and the in-kernel timeline stamps."""
    path = DEBUG_LIB_PATH if debug else LIB_PATH
    if not force and not is_stale(path):
        return path
    return _compile(path, ["-DL2A_DEBUG_KERNELS"] if debug else [], verbose)


def build_all(force=False):
    """Product and debug library, compiled side by side."""
    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(2) as ex:
        return list(ex.map(lambda d: build(force=force, debug=d), (False, True)))


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose=True, debug="--debug" in sys.argv))
