"""Kernel micro-bench: device-resident compress / decompress of 64 KiB blocks per corpus kind (CUDA events).
Usage: python divortio-lz4_b200/tools/kbench.py [MiB] [block] [kind,kind...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import divortio_lz4_b200 as dl  # noqa: E402
from divortio_lz4_b200 import corpus, device as dev  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    block = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    n = mib << 20
    ctx = dl.Context(0)
    d = torch.device("cuda", 0)
    s = torch.cuda.Stream()
    kinds = {"log": lambda: corpus.log(3, n), "zero": lambda: corpus.zero(n), "rand": lambda: corpus.rand(4, n),
             "mixed": lambda: corpus.mixed(2, n), "bench": lambda: corpus.benchjson(n)}
    stride = (dl.compress_bound(block) + 15) & ~15
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else list(kinds)
    for kind, gen in kinds.items():
        if kind not in only:
            continue
        src = torch.from_numpy(gen()).to(d)
        off, ln, nblk, coff = dev.uniform_blocks(n, block, d, stride)
        comp = torch.empty(nblk * stride + 64, dtype=torch.uint8, device=d)
        clen = torch.zeros(nblk, dtype=torch.int32, device=d)
        out = torch.empty(n + 64, dtype=torch.uint8, device=d)
        olen = torch.zeros(nblk, dtype=torch.int32, device=d)
        st = torch.zeros(nblk, dtype=torch.uint8, device=d)
        with torch.cuda.stream(s):
            for _ in range(2):
                dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
                dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, ln, olen, st)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 3
            tc = td = 0.0
            for _ in range(reps):
                e[0].record()
                dev.compress_blocks_dev(ctx, src, off, ln, block, comp, coff, clen)
                e[1].record()
                dev.decompress_blocks_dev(ctx, comp, coff, clen, out, off, ln, olen, st)
                e[2].record()
                s.synchronize()
                tc += e[0].elapsed_time(e[1])
                td += e[1].elapsed_time(e[2])
        assert torch.equal(out[:n], src[:n]) and int(st.max()) == 0
        c = int(clen.sum())
        print("%-6s %4d MiB blk %7d ratio %6.3f | compress %8.2f GB/s (%.2f ms) | decompress %8.2f GB/s (%.2f ms)" %
              (kind, mib, block, n / c, n / (tc / reps) / 1e6, tc / reps, n / (td / reps) / 1e6, td / reps), flush=True)


if __name__ == "__main__":
    main()
