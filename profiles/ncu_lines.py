"""Per-source-line samples and executed instructions of one kernel of an .ncu-rep (captured with --import-source on, built with
-lineinfo): joins `ncu --page source --csv` (SASS rows in program order) with `nvdisasm -g` of the library's cubin.
Usage: python profiles/ncu_lines.py <rep> <kernel-symbol-substring> [libdlz4_b200.so] [min-percent]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, sym = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "divortio-lz4_b200", "csrc", "libdlz4_b200.so")
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the function, in order, each with its (file, line)
ins = []
inside = False
cur = ("?", 0)
for ln in dis:
    if ln.startswith(".text."):
        inside = sym in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), cur))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
cs, cn, ce = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [r for r in rows[hi + 1:] if len(r) > ce]
if len(sass) != len(ins):
    print("warning: %d SASS rows in the report, %d in the cubin (different build?)" % (len(sass), len(ins)))
agg = {}
tot_s = tot_i = 0
for r, (_, text, loc) in zip(sass, ins):
    s, n = int(r[cn] or 0), int(r[ce] or 0)
    a = agg.setdefault(loc, [0, 0])
    a[0] += s
    a[1] += n
    tot_s += s
    tot_i += n
print("kernel %s: %d samples, %d warp-instructions" % (sym, tot_s, tot_i))
src_cache = {}
for loc, (s, n) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if 100.0 * s / max(tot_s, 1) < minpct and 100.0 * n / max(tot_i, 1) < minpct:
        continue
    f = os.path.join(os.path.dirname(os.path.abspath(so)), loc[0])
    if f not in src_cache:
        try:
            src_cache[f] = open(f).read().splitlines()
        except OSError:
            src_cache[f] = []
    text = src_cache[f][loc[1] - 1].strip() if 0 < loc[1] <= len(src_cache[f]) else ""
    print("%-18s %4d  %5.1f%% smp %5.1f%% ins | %s" % (loc[0], loc[1], 100.0 * s / max(tot_s, 1), 100.0 * n / max(tot_i, 1), text[:110]))
