"""Rank CUDA source lines of one kernel in an .ncu-rep by warp-stall samples (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    g = lambda n: r[hdr.index(n)]
    try:
        out.append((float(g("# Samples") or 0), float(g("Instructions Executed") or 0), cur, r[0], r[1].strip()[:95],
                    g("stall_long_sb"), g("stall_lg"), g("stall_wait"), g("stall_short_sb"), g("stall_membar"), g("stall_barrier")))
    except ValueError:
        pass
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"kernel {kern}: {ts:.0f} samples, {ti:.0f} warp instructions")
print(" samp   %s      inst   %i | long_sb lg wait short membar bar | line")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]:5.0f} {100*o[0]/ts:5.1f} {o[1]:9.0f} {100*o[1]/ti:5.1f} | {o[5]:>5} {o[6]:>4} {o[7]:>4} {o[8]:>4} {o[9]:>4} {o[10]:>4} | {o[2]}:{o[3]} {o[4]}")
