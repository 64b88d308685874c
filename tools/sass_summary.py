"""Opcode counts per kernel of the built library (cuobjdump -sass): the tracked evidence that the hot
kernels are tcgen05 / TMEM / TMA code.   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgp_b200 import _build  # noqa: E402

so = sys.argv[1] if len(sys.argv) > 1 else _build.SO
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS",
         "USETMAXREG", "HMMA", "FFMA2", "FFMA", "MUFU", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        per[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(so)}: instruction counts per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,")
print("# UTMASTG = TMA tensor store, LDGSTS = cp.async, SYNCS = mbarrier ops")
tot = collections.Counter()
for k, c in per.items():
    cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{k[:110]:110s} total={c['total']:6d} {cols}")
    tot.update(c)
print("-" * 60)
print("library total: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
