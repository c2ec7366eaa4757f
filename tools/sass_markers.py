"""Per-kernel SASS marker counts of the built library (cuobjdump -sass): which Blackwell / Hopper-class features each kernel
really uses.  python tools/sass_markers.py > profiles/r02_sass_markers.txt

  UBLKCP  cp.async.bulk (1-D TMA copy)        SYNCS   mbarrier arrive / try_wait       STAS   st.async (DSMEM push with complete_tx)
  UCGABAR cluster barrier                      IMMA    integer tensor-core MMA           REDUX  warp-wide integer reduction
  IDP     16x8 / 8x8-bit dot product           LDGSTS  cp.async (LDG -> shared)          UTCMMA / LDTM / STTM / UTMALDG: tcgen05 / TMEM / tensor-map TMA
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stereo_svo_slam_b200", "libstereosvo_b200.so")
MARKS = ["UBLKCP", "SYNCS", "STAS", "UCGABAR", "IMMA", "REDUX", "IDP", "LDGSTS", "PRMT", "SHFL", "DFMA|DMUL|DADD", "UTCMMA|UTC.MMA", "LDTM", "STTM", "UTMALDG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*\)$", "", kern).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if kern and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[kern]["instructions"] += 1
        for mk in MARKS:
            if re.search(r"\b(" + mk + r")[\._\s]", line):
                counts[kern][mk] += 1
print("library:", os.path.relpath(LIB, ROOT), " cubin arch:", sorted(arch))
cols = ["instructions"] + MARKS
print("%-58s" % "kernel" + "".join("%9s" % c[:8] for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    print("%-58s" % k[:57] + "".join("%9d" % c[x] for x in cols))
    tot.update(c)
print("%-58s" % "TOTAL" + "".join("%9d" % tot[x] for x in cols))
