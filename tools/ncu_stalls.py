"""Aggregate warp-stall samples of one kernel from an .ncu-rep source page (SASS view).
    python tools/ncu_stalls.py REPORT KERNEL_REGEX [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: 0 for i in stall_cols}
samples = hdr.index("# Samples")
for r in body:
    for i in stall_cols:
        tot[hdr[i]] += int(r[i] or 0)
all_s = sum(tot.values())
print(f"kernel {rows[0][1][:100]}  total samples {all_s}, instructions {len(body)}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {k:28s} {v:9d}  {100 * v / all_s:5.1f}%")
print("top instructions by samples:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][samples] or 0))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"  #{i:5d} {int(r[samples]):7d}  {r[1].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in why if v))
