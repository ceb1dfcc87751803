"""One compress + one decompress of a small corpus (for ncu).  Usage: prof_one.py [kind] [MiB] [block]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, device as dev  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "log"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 128
block = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
n = mib << 20
ctx = dl.Context(0)
d = torch.device("cuda", 0)
gen = {"log": lambda: corpus.log(3, n), "mixed": lambda: corpus.mixed(2, n)}[kind]
src = torch.from_numpy(gen()).to(d)
stride = (dl.compress_bound(block) + 15) & ~15
off, ln, nblk, coff = dev.uniform_blocks(n, block, d, stride)
comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=d)
clen = torch.zeros(nblk, dtype=torch.int32, device=d)
out = torch.empty(n + 64, dtype=torch.uint8, device=d)
olen = torch.zeros(nblk, dtype=torch.int32, device=d)
st = torch.zeros(nblk, dtype=torch.uint8, device=d)
for _ in range(2):
    dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
    dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, ln, olen, st)
torch.cuda.synchronize()
assert torch.equal(out[:n], src[:n])
print("ok", kind, mib, "MiB ratio", n / int(clen.sum()))
