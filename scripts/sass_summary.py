"""Per-kernel SASS instruction census of the built library (cuobjdump -sass): the mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine), STSM = stmatrix,
SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, UTCATOMSWS / UTCCP = TMEM alloc / copy) and registers / spills from ptxas.

    python scripts/sass_summary.py [path/to/lib.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "learning_to_adapt_b200", "lib", "libl2a_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "STSM", "SYNCS", "USETMAXREG", "UTCATOMSWS", "UTCCP", "ELECT", "BAR.SYNC",
        "MEMBAR", "LDG", "STG", "LDS", "STS", "FFMA", "DFMA", "DADD", "DMUL", "HMMA", "ATOM", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
counts = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[name]["_total"] += 1
    for k in KEYS:
        if op == k or op.startswith(k + ".") or (k == "BAR.SYNC" and op.startswith("BAR.SYNC")):
            counts[name][k] += 1
demangle = subprocess.run(["c++filt"] + list(counts.keys()), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("SASS census of %s (sm_100a)\n" % os.path.relpath(lib, REPO))
for (mangled, c), pretty in zip(counts.items(), demangle):
    pretty = re.sub(r"\(.*", "", pretty)
    reg, shm = usage.get(mangled, (None, None))
    body = "  ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])
    print("%-58s instr %6d  regs %s  static smem %s\n    %s" % (pretty[:58], c["_total"], reg, shm, body))
