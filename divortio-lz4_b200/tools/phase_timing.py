"""Per-phase cycle breakdown of the compressor's window path (needs `make -C divortio-lz4_b200/csrc prof`).
Usage: DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so python divortio-lz4_b200/tools/phase_timing.py [kind] [MiB]"""
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
names = ["loop head/other", "forward ring (cp.async refill + wait)", "A: shfl+hash+table gather+issue candidate loads",
         "same-slot test (match.any + ballot)", "verify+extend (waits for the candidate loads)", "B: resolve (walk)",
         "D+C: table scatter + emit", "trailing literals + sync", "batch step (fallback)"]
windows, conflicts, batches = v[9], v[10], v[11]
tot = sum(v[:9]) + v[12] + v[13]
print("%s %d MiB: %.2f ms, %d windows (%.1f B/window), %d conflict windows (%.1f%%), %d batch steps" %
      (kind, mib, e0.elapsed_time(e1), windows, n / max(windows, 1), conflicts, 100.0 * conflicts / max(windows, 1), batches))
for i, nm in enumerate(names):
    x = v[i] + (v[12] + v[13] if i == 2 else 0)
    print("  %-42s %8.1f cycles/window  %5.1f%%" % (nm, x / max(windows, 1), 100.0 * x / tot))
print("  total %.1f cycles/window" % (tot / max(windows, 1)))
print("  inside A: source bytes + hash %.1f, table gather until usable %.1f cycles/window" % (v[12] / max(windows, 1), v[13] / max(windows, 1)))
