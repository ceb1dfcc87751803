"""Cycle breakdown of the producer/walker match finder (needs `make -C divortio-lz4_b200/csrc prof`).
Usage: DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so python divortio-lz4_b200/tools/pw_phases.py [kind] [MiB]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, device as dev  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "log"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = mib << 20
ctx = dl.Context(0)
d = torch.device("cuda", 0)
src = torch.from_numpy({"log": lambda: corpus.log(3, n), "mixed": lambda: corpus.mixed(2, n)}[kind]()).to(d)
block = 65536
stride = (dl.compress_bound(block) + 15) & ~15
off, ln, nblk, coff = dev.uniform_blocks(n, block, d, stride)
comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=d)
clen = torch.zeros(nblk, dtype=torch.int32, device=d)
L = dl.lib()
cnt = (C.c_ulonglong * 16)()
dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
torch.cuda.synchronize()
L.dlz4_phase_counters(cnt)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
e1.record()
torch.cuda.synchronize()
L.dlz4_phase_counters(cnt)
v = list(cnt)
print("%s %d MiB: %.2f ms (parse + encode)" % (kind, mib, e0.elapsed_time(e1)))
wt = v[0] + v[1] + v[2] + v[3] + v[4] + v[5] + v[14]
print("walker: %.0f cycles per block; path steps %d (%.1f B/step), cut steps %d, ring steps %d, batch steps %d, poll retries %d" %
      (wt / nblk, v[9], n / max(v[9], 1), v[10], v[11], v[12], v[15]))
for i, nm in ((14, "path step: loop head, publish, slot address"), (1, "path step: poll / wait for ring entries"), (0, "path step: walk, insert, records"), (4, "cut steps (incl. slow probe)"),
              (2, "ring steps"), (3, "batch steps"), (5, "tail")):
    print("  %-44s %5.1f%%  %8.1f cycles per path step" % (nm, 100.0 * v[i] / max(wt, 1), v[i] / max(v[9], 1)))
pt = v[6] + v[7] + v[8]
print("producers: %d window pairs, %.0f cycles each; work %.1f%%, ring full %.1f%%, idle (sparse / done) %.1f%%" %
      (v[13], v[6] / max(v[13], 1), 100.0 * v[6] / max(pt, 1), 100.0 * v[7] / max(pt, 1), 100.0 * v[8] / max(pt, 1)))
