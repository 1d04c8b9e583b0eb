"""Development probe: tensor-pipe cycles per tile pair, SS vs TS vs tcgen05.cp (run under gpurun)."""
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
names = {0: "SS (A,B smem)", 1: "TS (A tmem)", 2: "cp only", 3: "cp + TS pipelined", 4: "SS + A keep/reuse",
         5: "out M128 N32 hints", 6: "out M128 N32 plain", 7: "out M128 N48 hints", 8: "out M128 N48 plain"}
for nc in (64, 80, 128):
    for mode in ((0, 1, 2, 3, 4, 5, 6, 7, 8) if nc == 80 else (0, 1, 4)):
        for _ in range(2):
            N.check(lib.l2a_debug_mma_rate(ctx, nc, mode, 400, C.c_void_p(out.data_ptr()), None))
            torch.cuda.synchronize()
        print("NC=%3d %-20s %7.1f cycles / tile pair (12 MMAs, ideal %d)" % (nc, names[mode], out[0].item() / 400.0, 12 * nc // 2))
# queue depth of the tensor pipe: how far the issuing thread runs ahead of execution for a short burst of MMAs
for iters in (1, 2, 3, 4, 8):
    N.check(lib.l2a_debug_mma_rate(ctx, 80, 4, iters, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    print("NC= 80 SS + hints, %d pair(s) = %2d MMAs: issued after %5d cycles, complete after %5d" % (iters, 12 * iters, out[1].item(), out[0].item()))

# shared-memory bandwidth: the same MMA stream with a concurrent TMA fill of 32 KB stages into the same CTA's shared memory
for mode, base in ((9, 4), (10, 0)):
    for _ in range(2):
        N.check(lib.l2a_debug_mma_rate(ctx, 80, mode, 2000, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
    cyc, copies = out[0].item(), out[2].item()
    print("NC= 80 %-20s + concurrent bulk-copy stream: %7.1f cycles / tile pair; %d x 32 KB landed in %d cycles = %.1f B/clk of fill writes"
          % (names[base], cyc / 2000.0, copies, cyc, copies * 32768.0 / cyc))

# what the issuer's per-pair handshake instructions cost on always-satisfied barriers, and the real 2-deep weight ring in isolation
hs = {11: "wait + fence + commit per pair", 12: "commit per pair", 13: "wait + fence per pair",
      14: "REAL ring 2 x 32 KB (1 producer)", 15: "REAL split ring hi/lo (2 producers)"}
for mode in (4, 11, 12, 13, 14, 15):
    for _ in range(2):
        N.check(lib.l2a_debug_mma_rate(ctx, 80, mode, 2000, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
    print("NC= 80 SS + hints, %-36s %7.1f cycles / tile pair" % (hs.get(mode, "bare loop"), out[0].item() / 2000.0))
