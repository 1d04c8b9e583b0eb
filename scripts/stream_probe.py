"""Development probe: weight-stream ring throughput vs stage count / tile size / issuing threads (run under gpurun)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_to_adapt_b200 import _native as N  # noqa: E402

lib = N.load(debug=True)
ctx = C.c_void_p()
N.check(lib.l2a_ctx_create(0, C.byref(ctx)))
blob = torch.randint(0, 255, (5, 136 * 16384), dtype=torch.uint8, device="cuda")   # 5 members x 2.2 MB
out = torch.zeros(148, dtype=torch.int64, device="cuda")
print("tile_bytes stages prod cons grid hold  cycles/tile  B/cyc/SM")
for tile_bytes, stage_list in ((32768, (2, 3)), (16384, (2, 4, 6)), (8192, (4, 8))):
    tiles = 136 * 16384 // tile_bytes
    for stages in stage_list:
        for prod, cons in ((1, 1), (2, 1), (1, 2), (2, 2)):
            for grid in (1, 125):
                for hold in (0, 280 * tile_bytes // 16384):
                    for _ in range(2):
                        N.check(lib.l2a_debug_stream(ctx, C.c_void_p(blob.data_ptr()), tiles, 20, stages, tile_bytes, hold, prod, cons,
                                                     grid, C.c_void_p(out.data_ptr()), None))
                        torch.cuda.synchronize()
                    cyc = out[:grid].max().item()
                    per_tile = cyc / (tiles * 20)
                    print("%6d %6d %4d %4d %4d %4d  %9.1f  %7.1f" % (tile_bytes, stages, prod, cons, grid, hold, per_tile, tile_bytes / per_tile))
