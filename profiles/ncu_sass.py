"""SASS rows (program order) of one kernel of an .ncu-rep with their source line, samples and executed count.
Usage: python profiles/ncu_sass.py <rep> <kernel-symbol-substring> <file> <line_lo> <line_hi> [libdlz4_b200.so]"""
import csv, os, re, subprocess, sys, tempfile
rep, sym, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
so = sys.argv[6] if len(sys.argv) > 6 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "divortio-lz4_b200", "csrc", "libdlz4_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
ins, inside, cur = [], False, ("?", 0)
for ln in dis:
    if ln.startswith(".text."):
        inside = sym in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi_]
cn, ce = hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [r for r in rows[hi_ + 1:] if len(r) > ce]
tot = sum(int(r[cn] or 0) for r in sass)
for r, (addr, text, loc) in zip(sass, ins):
    inl = "inlined" in loc[2]
    if loc[0] == fname and lo <= loc[1] <= hi or (len(sys.argv) > 7 and sys.argv[7] == "all"):
        print("%05x %-18s %4d %7d smp %11d exe | %s" % (addr, loc[0][:18], loc[1], int(r[cn] or 0), int(r[ce] or 0), text[:90]))
