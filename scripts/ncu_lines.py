"""Join an ncu source page (per SASS instruction) with nvdisasm's line info: executed instructions and stall samples
per CUDA source line.   python scripts/ncu_lines.py <rep> <cubin> <mangled kernel substring> [top]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
h = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[h:]))))
base = int(rows[0]["Address"], 16)
prof = {}
for r in rows:
    prof[int(r["Address"], 16) - base] = (int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0), r["Source"].strip())
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
cur = None
per = collections.defaultdict(lambda: [0, 0, 0])
for l in sass[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3) or "")
        cur = (m.group(1).split("/")[-1], int(m.group(2)), (inl.group(1).split("/")[-1], int(inl.group(2))) if inl else None)
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur:
        off = int(m.group(1), 16)
        if off in prof:
            p = per[cur]
            p[0] += prof[off][0]; p[1] += prof[off][1]; p[2] += 1
tot_i = sum(p[0] for p in per.values()); tot_s = sum(p[1] for p in per.values())
print(f"total warp instructions {tot_i}, samples {tot_s}, static instructions {sum(p[2] for p in per.values())}")
for k, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} {'<- ' + k[2][0] + ':' + str(k[2][1]) if k[2] else '':32s} inst {100 * p[0] / tot_i:5.1f}%  samples {100 * p[1] / max(tot_s, 1):5.1f}%  static {p[2]}")
