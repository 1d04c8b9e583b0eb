"""Development probe for the next step of K1: a CTA pair issuing tcgen05.mma.cta_group::2 at M = 256,
N = 160, and the DSMEM bulk copy its epilogue all-to-all would use (run under gpurun; first thing next round).
Each launch is wrapped by the caller's `timeout`: the kernels have the usual 2 s barrier watchdog."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_to_adapt_b200 import _native as N  # noqa: E402

lib = N.load(debug=True)
ctx = C.c_void_p()
N.check(lib.l2a_ctx_create(0, C.byref(ctx)))
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for iters in (1, 2, 8, 64, 400):
    for _ in range(2):
        N.check(lib.l2a_debug_pair(ctx, 0, iters, 16, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
    o = out.cpu().numpy()
    print("cta_group::2 M=256 N=160: %3d tile pairs: %7.1f cycles / pair (12 MMAs; single-CTA N=80 with hints: 563), issued after %d, complete after %d"
          % (iters, o[0] / iters, o[1], o[0]))
for nbytes in (1024, 4096, 10240, 16384):
    iters = min(48, (1 << 20) // nbytes - 1)
    for _ in range(2):
        N.check(lib.l2a_debug_pair(ctx, 1, iters, nbytes, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
    o = out.cpu().numpy()
    print("DSMEM bulk copy %5d B x %2d, both directions: %.1f / %.1f B/clk per receiver" % (nbytes, iters, nbytes * iters / o[2], nbytes * iters / o[3]))

for mode, what in ((2, "512 contiguous bytes per warp instruction"), (3, "four 128-byte pieces 1152 B apart per warp instruction (MN-major epilogue)")):
    for iters in (16, 64, 256):
        for _ in range(2):
            N.check(lib.l2a_debug_pair(ctx, mode, iters, 16, C.c_void_p(out.data_ptr()), None))
            torch.cuda.synchronize()
        o = out.cpu().numpy()
        nbytes = iters * 4 * 512                     # 4 warps x 512 B per instruction
        print("DSMEM st.shared::cluster.v4, %s: %4d instr / warp, %6d B per CTA: %.1f / %.1f B/clk (both directions at once)"
              % (what, iters, nbytes, nbytes / o[2], nbytes / o[3]))
