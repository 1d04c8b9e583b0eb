"""Development probe: in-kernel timeline of CTA 0 during one horizon step of the headline rollout (run under gpurun)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpc_oracle as O  # noqa: E402
from learning_to_adapt_b200.engine import PlanningEngine  # noqa: E402

CFG = {"headline": ("half_cheetah", (512, 512, 512), 2000, 20, 1, 5, 2), "cfg1": ("half_cheetah", (512, 512), 500, 10, 1, 1, 0),
       "cfg2i": ("half_cheetah", (512, 512, 512), 1000, 15, 5, 5, 1), "cfg1p": ("half_cheetah", (512, 512), 2000, 20, 10, 1, 0)}
env, hidden, n, h, m, nsets, mode = CFG[sys.argv[1] if len(sys.argv) > 1 else "headline"]
KERNEL = int(sys.argv[2]) if len(sys.argv) > 2 else 2          # 2 = single-CTA tcgen05 kernel, 3 = CTA pair
prob = O.make_problem(env, hidden_sizes=hidden, n_sets=nsets, m=m, seed=0)
eng = PlanningEngine(prob["obs_dim"], prob["act_dim"], hidden, n_sets=nsets, debug=True)
for i, p in enumerate(prob["param_sets"]):
    eng.set_params(i, p)
eng.set_normalization(prob["norm"])
obs = eng._f32(prob["obs0"])
low, high = eng._f32(prob["low"]), eng._f32(prob["high"])
acts = torch.rand((h, n * m, prob["act_dim"]), device="cuda") * (high - low) + low
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
eng.lib.l2a_debug_set_timeline(eng._ctx, C.c_void_p(tl.data_ptr()))
for _ in range(3):
    eng.rollout(obs, acts, n, h, prob["reward_kind"], prob["dt"], set_mode=mode, first_set=0, n_sets=nsets, want_returns=False, kernel=KERNEL)
torch.cuda.synchronize()
t = tl.cpu().numpy()
t0 = t[0]
L = len(hidden) + 1
print("MMA warp (cycles since layer-0 start of step 1):")
for l in range(L - 1):
    b = 4 * l
    print("  layer %d: enter %7d  first-act-ready %7d  phaseA-done %7d  committed %7d" % (l, t[b] - t0, t[b + 3] - t0, t[b + 1] - t0, t[b + 2] - t0))
b = 4 * (L - 1)
print("  output : enter %7d  first-act-ready %7d  committed %7d   | next step's layer 0 enter %7d" % (t[b] - t0, t[b + 3] - t0, t[b + 2] - t0, t[80] - t0))
print("epilogue warp 0:")
for l in range(L - 1):
    b = 32 + 4 * l
    print("  layer %d: layer_full seen %7d  mb0 published %7d  all published %7d" % (l, t[b] - t0, t[b + 1] - t0, t[b + 2] - t0)
          + ("   (pair: first column = early seen; layer_full seen for M-block 1 %7d)" % (t[b + 3] - t0) if KERNEL == 3 else ""))
print("  output: layer_full %7d  deltas in registers %7d  member mean done %7d  env done %7d  x written %7d" % tuple(int(t[i] - t0) for i in (60, 61, 62, 63, 64)))
print("  exchange: rows stored + CTA barrier %7d  release-arrive on the peers issued %7d  all members arrived %7d" % tuple(int(t[i] - t0) for i in (65, 66, 67)))
print("  write_x: stores done %7d  fence.proxy.async done %7d" % (int(t[70] - t0), int(t[71] - t0)))
print("first cluster, per member (SM clocks are per-SM counters; only the differences within a member are comparable):")
for e in range(min(nsets, 5)):
    if KERNEL == 3:
        print("  member %d (globaltimer, ns rel. to member 0's flag): flag released %7d   all flags seen %7d" % (e, int(t[86 + 2 * e] - t[86]), int(t[87 + 2 * e] - t[86])))
    else:
        print("  member %d: deltas ready -> member mean done: %7d cycles" % (e, int(t[87 + 2 * e] - t[86 + 2 * e])))

if KERNEL == 3:
    e0 = t[32 + 4 * 1]
    print("pair epilogue of layer 1, M-block 0, warp 0 (cycles after `early` seen): own groups loaded %d, stored %d | peer groups loaded %d, stored (issued) %d | proxy fence + syncwarp %d | arrived %d"
          % tuple(int(t[i] - e0) for i in (100, 101, 103, 104, 105, 32 + 4 * 1 + 1)))
