"""Dev-time: per-source-line instruction counts and stall samples for one kernel of an .ncu-rep.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep raster_kernel [top_n] [--sass]

Joins `ncu --page source --csv` (SASS rows in address order) with `nvdisasm -g` line info of the cubins
embedded in rendertoy_b200/librendertoy_b200.so (the library must be the build that was profiled).
"""
import csv
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 30
show_sass = "--sass" in sys.argv
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rendertoy_b200", "librendertoy_b200.so")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
sass = rows[2:]
# stop at the next kernel header, if several launches matched
for i, r in enumerate(sass):
    if r and r[0] == "Kernel Name":
        sass = sass[:i]
        break

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", pat)
lines = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur_fn, cur_line, table = None, None, collections.OrderedDict()
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1); table[cur_fn] = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn:
            table[cur_fn].append((int(m.group(1), 16), cur_line, m.group(2).strip()))
    for fn, ins in table.items():
        if mangled_hint in fn and len(ins) == len(sass):
            lines = ins
            break
    if lines:
        break
if lines is None:
    sys.exit(f"no cubin function matching {pat} with {len(sass)} instructions (is the .so the profiled build?)")

ie, ss = col["Instructions Executed"], col["# Samples"]
stall_cols = [n for n in hdr if n.startswith("stall_")] if any(n.startswith("stall_") for n in hdr) else []
per_line = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r, (off, line, text) in zip(sass, lines):
    i, s = int(r[ie] or 0), int(r[ss] or 0)
    per_line[line][0] += i; per_line[line][1] += s
    tot_i += i; tot_s += s
print(f"{kname}\n  warp instructions {tot_i}  samples {tot_s}")
src_cache = {}
def src(line):
    if line is None: return ""
    f, n = line
    if f not in src_cache:
        p = os.path.join(os.path.dirname(so), "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src_cache[f][n - 1].strip()[:100] if 0 < n <= len(src_cache[f]) else ""
for line, (i, s) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {i / max(tot_i, 1) * 100:5.1f}% instr  {s / max(tot_s, 1) * 100:5.1f}% samples  {line[0] if line else '?'}:{line[1] if line else 0:<4d} {src(line)}")
if show_sass:
    print("  -- hottest SASS")
    order = sorted(range(len(sass)), key=lambda k: -int(sass[k][ss] or 0))[:top]
    for k in order:
        print(f"  {int(sass[k][ss] or 0) / max(tot_s, 1) * 100:5.1f}% samples  {int(sass[k][ie] or 0):9d} exec  L{lines[k][1][1] if lines[k][1] else 0:<4d} {lines[k][2][:90]}")
