"""Program-order listing of the tcgen05 / TMA / mbarrier / DSMEM / fence instructions of ONE kernel of the built library.
    python scripts/sass_excerpt.py rollout_tc2_kernelILi72ELi24 > profiles/r02_rollout_tc2_sass_excerpt.txt"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, "learning_to_adapt_b200", "lib", "libl2a_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "rollout_tc2_kernelILi72ELi24"
KEYS = ("UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STAS", "UTCATOMSWS", "USETMAXREG", "SYNCS.", "UTMAPF", "UTCCP", "FENCE.VIEW", "ELECT",
        "MEMBAR", "UBLKCP", "REDAS", "BAR.SYNC", "CCTL")
lines = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
start = next(i for i, l in enumerate(lines) if "Function : " in l and tag in l)
end = next((i for i in range(start + 1, len(lines)) if "Function : " in lines[i]), len(lines))
ins = [m for m in (re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l) for l in lines[start:end]) if m]
listed = [m for m in ins if any(k in m.group(2) for k in KEYS)]
print("SASS excerpt of %s from learning_to_adapt_b200/lib/libl2a_b200.so (`cuobjdump -sass`):" % lines[start].split("Function : ")[1].strip())
print("every tcgen05 / TMA / mbarrier / DSMEM / fence / barrier instruction of the kernel in program order (address, instruction);")
print("%d SASS instructions in the kernel, %d listed.  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit," % (len(ins), len(listed)))
print("LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (tensor-map TMA), STAS = st.async (DSMEM), SYNCS = mbarrier, UTCATOMSWS = tcgen05.alloc.\n")
for m in listed:
    print("%s  %s" % (m.group(1), re.sub(r"\s+", " ", m.group(2)).strip()))
