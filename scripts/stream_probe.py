"""Development probe: weight-stream ring throughput vs stage count / tile size / CTA count (run under gpurun)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_to_adapt_b200 import _native as N  # noqa: E402

lib = N.load()
ctx = C.c_void_p()
N.check(lib.l2a_ctx_create(0, C.byref(ctx)))
blob = torch.randint(0, 255, (5, 136 * 16384), dtype=torch.uint8, device="cuda")   # 5 members x 2.2 MB
out = torch.zeros(148, dtype=torch.int64, device="cuda")
print("tile_bytes stages grid hold  cycles/tile  B/cyc/SM  total_TB/s@1.9GHz")
for tile_bytes in (16384, 8192, 32768):
    tiles = 136 * 16384 // tile_bytes
    for stages in (2, 3, 4, 6, 8, 12):
        if stages * tile_bytes > 200 * 1024:
            continue
        for grid in (1, 25, 125, 148):
            for hold in (0, 240 * tile_bytes // 16384):
                for _ in range(2):
                    N.check(lib.l2a_debug_stream(ctx, C.c_void_p(blob.data_ptr()), tiles, 20, stages, tile_bytes, hold, grid,
                                                 C.c_void_p(out.data_ptr()), None))
                    torch.cuda.synchronize()
                cyc = out[:grid].max().item()
                per_tile = cyc / (tiles * 20)
                bpc = tile_bytes / per_tile
                print("%6d %3d %4d %4d  %9.1f  %7.1f  %6.2f" % (tile_bytes, stages, grid, hold, per_tile, bpc, bpc * grid * 1.9e9 / 1e12))
