"""Join an ncu source-page CSV (SASS view: per-instruction stall samples) with nvdisasm line info of the same cubin and
aggregate by CUDA source line.   usage: stall_by_line.py <sass.csv> <object.o> <mangled-name substring> [top N]"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

sass_csv, obj, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = {}, ("?", 0), False
for l in dis:
    if l.startswith("//--------------------- .text."):
        active = key in l
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > idx["Instructions Executed"] and r[0].startswith("0x")]
seen, uniq = set(), []
for r in data:
    if r[0] not in seen:
        seen.add(r[0]); uniq.append(r)
base = int(uniq[0][0], 16)
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_i = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in uniq:
    off = int(r[0], 16) - base
    (f, ln), ins = lines.get(off, (("?", 0), r[1]))
    s, ex = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    e = by_line[(f, ln)]
    e[0] += s; e[1] += ex
    for c in stall_cols:
        v = int(r[idx[c]] or 0)
        if v:
            e[2][c[6:]] += v
    tot_s += s; tot_i += ex
print(f"total samples {tot_s}, warp instructions {tot_i}; samples per instruction {tot_s / max(tot_i, 1) * 1e3:.3f} (x1e-3)")
by_file = collections.defaultdict(lambda: [0, 0])
for (f, ln), (s, ex, _) in by_line.items():
    by_file[f][0] += s; by_file[f][1] += ex
for f, (s, ex) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} samples {100 * s / tot_s:5.1f} %  instr {100 * ex / max(tot_i, 1):5.1f} %")
print(f"top {top} lines by samples:")
for (f, ln), (s, ex, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    main = ", ".join(f"{k} {100 * v / max(s, 1):.0f}%" for k, v in st.most_common(3))
    print(f"  {f}:{ln:<5d} samples {100 * s / tot_s:5.2f} %  instr {100 * ex / max(tot_i, 1):5.2f} %  ratio {s / max(ex, 1) * tot_i / tot_s:4.2f}   {main}")
