#!/usr/bin/env python
"""usage: ncu_lines.py report.ncu-rep <cubin-name-substr> <kernel-mangled-substr> [top]
Joins ncu's per-SASS-instruction counters (source page, SASS view) with nvdisasm's line info of the SAME build of
libpantera_gpu.so (instruction order within the kernel) and prints executed instructions / stall samples per CUDA
source line.  The .so must be the build that was profiled."""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, cub, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pantera_pic_dsmc_b200", "csrc", "libpantera_gpu.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
f = [x for x in os.listdir(tmp) if cub in x and x.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
# walk the disassembly: function label, line markers, instructions
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"^(\S+):\s*$", ln)
    if ln.startswith("//--------------------- .text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) > 10]
print(f"# {len(sass)} SASS instructions in the report, {len(lines)} in the disassembly", file=sys.stderr)
n = min(len(sass), len(lines))
agg = defaultdict(lambda: [0, 0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += int(sass[k][ix["Instructions Executed"]])
    a[1] += int(sass[k][ix["# Samples"]])
    a[2] += int(sass[k][ix["Thread Instructions Executed"]])
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
src = {}
for (fn, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if fn not in src:
        p = os.path.join(ROOT, "pantera_pic_dsmc_b200", "csrc", fn)
        src[fn] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[fn][l - 1].strip()[:110] if 0 < l <= len(src[fn]) else ""
    print(f"{a[0] / tot * 100:5.2f}% inst {a[1] / tots * 100:5.2f}% stall  {fn}:{l:<5d} {text}")
